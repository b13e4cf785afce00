// T path — sparse labelled-pixel cross entropy fused with the final bilinear upsample, and the
// stand-alone align_corners=True bilinear resize.
//
// The reference writes full-resolution logits (deeplab.py:55), overwrites every unlabelled target
// with ignore_index (model.py:108-110) and runs a dense log-softmax + NLL over B*C*H*W values
// (model.py:116) for ~10-100 labelled pixels per image.  Here the labelled pixels arrive as a list
// and one warp per pixel gathers the 4 low-resolution neighbours x C, interpolates, evaluates
// log-softmax/NLL and scatter-adds the gradient into the low-resolution logits.
#include "pp_common.cuh"

namespace pp {

constexpr int kCeThreads = 1024;

__global__ void __launch_bounds__(kCeThreads) sparse_ce_kernel(
    const float* __restrict__ logits, int n_img, int C, int h_in, int w_in, int H, int W, float scale_h,
    float scale_w, const int32_t* __restrict__ px_img, const int32_t* __restrict__ px_idx,
    const int32_t* __restrict__ px_label, int n_px_cap, const int32_t* __restrict__ n_px_dev, float grad_scale,
    float* __restrict__ loss, float* __restrict__ grad, int32_t* __restrict__ pred_at) {
  // the labelled-pixel count may live on the device (CUDA-graph replays with a fixed-capacity list)
  const int n_px = n_px_dev ? min(*n_px_dev, n_px_cap) : n_px_cap;
  const float inv_n = n_px > 0 ? 1.0f / (float)n_px : __int_as_float(0x7FC00000);
  const float grad_coef = grad_scale * inv_n;
  __shared__ float sh_part[kCeThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warps_per_block = kCeThreads / 32;
  const int64_t plane = (int64_t)h_in * w_in;
  float acc = 0.f;  // lane 0 accumulates this warp's NLL in pixel order
  for (int i = blockIdx.x * warps_per_block + warp; i < n_px; i += gridDim.x * warps_per_block) {
    const int img = px_img[i], idx = px_idx[i], label = px_label[i];
    // malformed entry (F.cross_entropy raises "Target out of bounds" here, model.py:116): no out-of-bounds access, no
    // silent training on a wrong class - the entry is skipped and the loss is poisoned with NaN so the step fails loudly
    if ((unsigned)img >= (unsigned)n_img || (unsigned)idx >= (unsigned)(H * W) || (unsigned)label >= (unsigned)C) {
      if (lane == 0) {
        acc = __int_as_float(0x7FC00000);
        if (pred_at) pred_at[i] = -1;
      }
      continue;
    }
    const int y = idx / W, x = idx - y * W;
    const Lerp ly = lerp_ac(y, h_in, H, scale_h);
    const Lerp lx = lerp_ac(x, w_in, W, scale_w);
    const float* b = logits + (int64_t)img * C * plane;
    const int64_t o00 = (int64_t)ly.i0 * w_in + lx.i0, o01 = (int64_t)ly.i0 * w_in + lx.i1;
    const int64_t o10 = (int64_t)ly.i1 * w_in + lx.i0, o11 = (int64_t)ly.i1 * w_in + lx.i1;
    // pass 1: max and argmax (first maximal class, as torch.argmax on CPU)
    float m = -INFINITY;
    int am = 0x7FFFFFFF;
    for (int c = lane; c < C; c += 32) {
      const float* pc = b + c * plane;
      const float v = ly.l0 * (lx.l0 * pc[o00] + lx.l1 * pc[o01]) + ly.l1 * (lx.l0 * pc[o10] + lx.l1 * pc[o11]);
      if (v > m) { m = v; am = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xFFFFFFFFu, m, o);
      const int oa = __shfl_xor_sync(0xFFFFFFFFu, am, o);
      if (om > m || (om == m && oa < am)) { m = om; am = oa; }
    }
    float s = 0.f, xl = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float* pc = b + c * plane;
      const float v = ly.l0 * (lx.l0 * pc[o00] + lx.l1 * pc[o01]) + ly.l1 * (lx.l0 * pc[o10] + lx.l1 * pc[o11]);
      s += expf(v - m);
      if (c == label) xl = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
      xl += __shfl_xor_sync(0xFFFFFFFFu, xl, o);
    }
    const float lse = m + logf(s);
    if (lane == 0) {
      acc += lse - xl;
      if (pred_at) pred_at[i] = am;
    }
    if (grad) {
      float* g = grad + (int64_t)img * C * plane;
      for (int c = lane; c < C; c += 32) {
        const float* pc = b + c * plane;
        const float v = ly.l0 * (lx.l0 * pc[o00] + lx.l1 * pc[o01]) + ly.l1 * (lx.l0 * pc[o10] + lx.l1 * pc[o11]);
        float d = expf(v - lse);
        if (c == label) d -= 1.f;
        d *= grad_coef;
        float* gc = g + c * plane;
        atomicAdd(gc + o00, d * ly.l0 * lx.l0);
        atomicAdd(gc + o01, d * ly.l0 * lx.l1);
        atomicAdd(gc + o10, d * ly.l1 * lx.l0);
        atomicAdd(gc + o11, d * ly.l1 * lx.l1);
      }
    }
  }
  if (lane == 0) sh_part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < warps_per_block; ++w) t += sh_part[w];
    if (gridDim.x == 1) *loss = t * inv_n;  // deterministic for the usual (sparse) case
    else atomicAdd(loss, t * inv_n);
  }
}

__global__ void fill_kernel(float* p, float v) { *p = v; }

__global__ void __launch_bounds__(256) upsample_ac_kernel(const float* __restrict__ in, int h_in, int w_in,
                                                          float* __restrict__ out, int H, int W,
                                                          float scale_h, float scale_w, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int x = (int)(i % W);
    const int64_t t = i / W;
    const int y = (int)(t % H);
    const int64_t nc = t / H;
    const Lerp ly = lerp_ac(y, h_in, H, scale_h);
    const Lerp lx = lerp_ac(x, w_in, W, scale_w);
    const float* pc = in + nc * h_in * w_in;
    const float* r0 = pc + (int64_t)ly.i0 * w_in;
    const float* r1 = pc + (int64_t)ly.i1 * w_in;
    out[i] = ly.l0 * (lx.l0 * __ldg(r0 + lx.i0) + lx.l1 * __ldg(r0 + lx.i1)) +
             ly.l1 * (lx.l0 * __ldg(r1 + lx.i0) + lx.l1 * __ldg(r1 + lx.i1));
  }
}

__global__ void __launch_bounds__(256) upsample_ac_bwd_kernel(const float* __restrict__ gout, int H, int W,
                                                              float* __restrict__ gin, int h_in, int w_in,
                                                              float scale_h, float scale_w, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int x = (int)(i % W);
    const int64_t t = i / W;
    const int y = (int)(t % H);
    const int64_t nc = t / H;
    const Lerp ly = lerp_ac(y, h_in, H, scale_h);
    const Lerp lx = lerp_ac(x, w_in, W, scale_w);
    const float g = __ldg(gout + i);
    float* pc = gin + nc * h_in * w_in;
    atomicAdd(pc + (int64_t)ly.i0 * w_in + lx.i0, g * ly.l0 * lx.l0);
    atomicAdd(pc + (int64_t)ly.i0 * w_in + lx.i1, g * ly.l0 * lx.l1);
    atomicAdd(pc + (int64_t)ly.i1 * w_in + lx.i0, g * ly.l1 * lx.l0);
    atomicAdd(pc + (int64_t)ly.i1 * w_in + lx.i1, g * ly.l1 * lx.l1);
  }
}

}  // namespace pp

namespace pp {
// Validation path (model.py:177-239, eval.py:15-94): pred = argmax_c F.interpolate(logits_lowres, (H, W), bilinear,
// align_corners=True); RunningScore.update(y, pred).  One pass: every thread interpolates its pixel's C logits from the
// 1/4-resolution head output (full-resolution logits and the int64 prediction map never exist), takes the first maximum
// (torch.argmax), and the (label, prediction) pair goes into a per-CTA shared-memory confusion matrix that is flushed with
// one atomic per non-empty cell.  label dtype: 0 int64, 1 int32, 2 uint8.
template <int C>
__global__ void __launch_bounds__(256) eval_confusion_up_kernel(const float* __restrict__ logits, int h_in, int w_in, int H, int W,
                                                                float scale_h, float scale_w, const void* __restrict__ labels,
                                                                int label_dtype, unsigned long long* __restrict__ confusion,
                                                                int32_t* __restrict__ pred_out) {
  __shared__ uint32_t sh[C * C];
  for (int i = threadIdx.x; i < C * C; i += 256) sh[i] = 0;
  __syncthreads();
  const int img = blockIdx.y;
  const int64_t HW = (int64_t)H * W;
  const int64_t plane = (int64_t)h_in * w_in;
  const float* __restrict__ base = logits + (int64_t)img * C * plane;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < HW; i += (int64_t)gridDim.x * 256) {
    const int y = (int)(i / W);
    const int x = (int)(i - (int64_t)y * W);
    const Lerp ly = lerp_ac(y, h_in, H, scale_h);
    const Lerp lx = lerp_ac(x, w_in, W, scale_w);
    const float* r0 = base + (int64_t)ly.i0 * w_in;
    const float* r1 = base + (int64_t)ly.i1 * w_in;
    float best = 0.f;
    int arg = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float v00 = __ldg(r0 + c * plane + lx.i0), v01 = __ldg(r0 + c * plane + lx.i1);
      const float v10 = __ldg(r1 + c * plane + lx.i0), v11 = __ldg(r1 + c * plane + lx.i1);
      const float v = ly.l0 * (lx.l0 * v00 + lx.l1 * v01) + ly.l1 * (lx.l0 * v10 + lx.l1 * v11);
      if (c == 0 || v > best || (v != v && best == best)) {  // first maximum; NaN counts as the maximum (torch.argmax)
        best = v;
        arg = c;
      }
    }
    const int64_t pix = (int64_t)img * HW + i;
    if (pred_out) pred_out[pix] = arg;
    long long lt;
    if (label_dtype == 0) lt = reinterpret_cast<const long long*>(labels)[pix];
    else if (label_dtype == 1) lt = reinterpret_cast<const int32_t*>(labels)[pix];
    else lt = reinterpret_cast<const uint8_t*>(labels)[pix];
    if (lt >= 0 && lt < C) atomicAdd(&sh[(int)lt * C + arg], 1u);  // utils/metrics.py:168-173 (_fast_hist mask)
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * C; i += 256) {
    const uint32_t v = sh[i];
    if (v) atomicAdd(confusion + i, (unsigned long long)v);
  }
}
}  // namespace pp

namespace pp {
__global__ void metrics_accumulate_kernel(const int32_t* __restrict__ labels, const int32_t* __restrict__ preds,
                                          const int32_t* __restrict__ n_valid_dev, int n_max, int n_classes,
                                          const float* __restrict__ loss, unsigned long long* __restrict__ confusion,
                                          double* __restrict__ loss_sum, unsigned long long* __restrict__ n_steps) {
  int n = n_valid_dev ? *n_valid_dev : n_max;
  n = n < n_max ? n : n_max;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int lt = labels[i], lp = preds[i];
    if (lt >= 0 && lt < n_classes && lp >= 0 && lp < n_classes)  // utils/metrics.py:168-173 (_fast_hist mask)
      atomicAdd(confusion + (size_t)lt * n_classes + lp, 1ull);
  }
  if (threadIdx.x == 0) {
    if (loss && loss_sum) *loss_sum += (double)*loss;
    if (n_steps) *n_steps += 1ull;
  }
}
}  // namespace pp

using namespace pp;

extern "C" {

int pp_sparse_ce(const float* logits_lowres, int n_img, int C, int h_in, int w_in, int H, int W,
                 const int32_t* px_img, const int32_t* px_idx, const int32_t* px_label, int n_px,
                 const int32_t* n_px_dev, float grad_scale, float* loss, float* grad_lowres, int32_t* pred_at,
                 void* stream) {
  PP_CHECK_ARG(logits_lowres && loss, "pp_sparse_ce: null pointer");
  PP_CHECK_ARG(n_img > 0 && C >= 2 && h_in > 0 && w_in > 0 && H > 0 && W > 0 && n_px >= 0, "pp_sparse_ce: bad shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n_px == 0 && !n_px_dev) {  // F.cross_entropy over an empty selection is NaN (mean of nothing)
    fill_kernel<<<1, 1, 0, st>>>(loss, __builtin_nanf(""));
    PP_LAUNCH_CHECK();
    return PP_OK;
  }
  PP_CHECK_ARG(px_img && px_idx && px_label, "pp_sparse_ce: null pixel list");
  const int wpb = kCeThreads / 32;
  int grid = 1;
  if (n_px > 4096) {
    grid = (n_px + wpb * 8 - 1) / (wpb * 8);
    if (grid > 148 * 2) grid = 148 * 2;
    PP_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
  }
  sparse_ce_kernel<<<grid, kCeThreads, 0, st>>>(logits_lowres, n_img, C, h_in, w_in, H, W, ac_scale(h_in, H),
                                                 ac_scale(w_in, W), px_img, px_idx, px_label, n_px, n_px_dev,
                                                 grad_scale, loss, grad_lowres, pred_at);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int pp_upsample_bilinear_ac(const float* in, int n_img, int C, int h_in, int w_in, float* out, int H, int W,
                            void* stream) {
  PP_CHECK_ARG(in && out && n_img > 0 && C > 0 && h_in > 0 && w_in > 0 && H > 0 && W > 0, "pp_upsample_bilinear_ac: bad args");
  const int64_t total = (int64_t)n_img * C * H * W;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  upsample_ac_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      in, h_in, w_in, out, H, W, ac_scale(h_in, H), ac_scale(w_in, W), total);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int pp_upsample_bilinear_ac_bwd(const float* grad_out, int n_img, int C, int H, int W, float* grad_in, int h_in,
                                int w_in, void* stream) {
  PP_CHECK_ARG(grad_out && grad_in && n_img > 0 && C > 0 && h_in > 0 && w_in > 0 && H > 0 && W > 0, "pp_upsample_bilinear_ac_bwd: bad args");
  const int64_t total = (int64_t)n_img * C * H * W;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  upsample_ac_bwd_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      grad_out, H, W, grad_in, h_in, w_in, ac_scale(h_in, H), ac_scale(w_in, W), total);
  PP_LAUNCH_CHECK();
  return PP_OK;
}


/* Running train metrics on the device: the reference copies two full int64 maps to the host every
 * step for RunningScore.update (model.py:124-129, utils/metrics.py:162-177).  Only labelled pixels count (every other
 * pixel is ignore_index and _fast_hist drops it), so the confusion matrix is accumulated from the (label, prediction)
 * pairs the sparse-CE kernel already produced — inside the captured step, read back once per epoch. */
int pp_metrics_accumulate(const int32_t* labels, const int32_t* preds, const int32_t* n_valid_dev, int n_max,
                          int n_classes, const float* loss, long long* confusion, double* loss_sum,
                          long long* n_steps, void* stream) {
  PP_CHECK_ARG(labels && preds && confusion && n_max >= 0 && n_classes > 0, "pp_metrics_accumulate: bad args");
  pp::metrics_accumulate_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      labels, preds, n_valid_dev, n_max, n_classes, loss, reinterpret_cast<unsigned long long*>(confusion), loss_sum,
      reinterpret_cast<unsigned long long*>(n_steps));
  PP_LAUNCH_CHECK();
  return PP_OK;
}


int pp_eval_confusion_upsampled(const float* logits_lowres, int n_img, int C, int h_in, int w_in, int H, int W,
                                const void* labels, int label_dtype, long long* confusion, int32_t* pred_out,
                                void* stream) {
  PP_CHECK_ARG(logits_lowres && labels && confusion, "pp_eval_confusion_upsampled: null pointer");
  PP_CHECK_ARG(n_img > 0 && n_img <= 65535 && h_in > 0 && w_in > 0 && H > 0 && W > 0, "pp_eval_confusion_upsampled: bad shape");
  PP_CHECK_ARG(label_dtype >= 0 && label_dtype <= 2, "pp_eval_confusion_upsampled: label_dtype=%d (0 int64, 1 int32, 2 uint8)",
               label_dtype);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t HW = (int64_t)H * W;
  int64_t gx = (HW + 256 * 4 - 1) / (256 * 4);  // ~4 pixels per thread: amortises the shared-memory matrix flush
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, n_img);
  const float sh_ = pp::ac_scale(h_in, H), sw_ = pp::ac_scale(w_in, W);
  unsigned long long* cf = reinterpret_cast<unsigned long long*>(confusion);
  switch (C) {
    case 11: pp::eval_confusion_up_kernel<11><<<grid, 256, 0, st>>>(logits_lowres, h_in, w_in, H, W, sh_, sw_, labels, label_dtype, cf, pred_out); break;
    case 19: pp::eval_confusion_up_kernel<19><<<grid, 256, 0, st>>>(logits_lowres, h_in, w_in, H, W, sh_, sw_, labels, label_dtype, cf, pred_out); break;
    case 21: pp::eval_confusion_up_kernel<21><<<grid, 256, 0, st>>>(logits_lowres, h_in, w_in, H, W, sh_, sw_, labels, label_dtype, cf, pred_out); break;
    default:
      pp::set_error("pp_eval_confusion_upsampled: C=%d not instantiated (11, 19, 21)", C);
      return PP_ERR_UNSUPPORTED;
  }
  PP_LAUNCH_CHECK();
  return PP_OK;
}

}  // extern "C"
