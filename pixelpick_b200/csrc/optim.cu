// T path - the optimiser step of model.py:121 (`self.optimizer.step()`, the Adam utils/utils.py:112-141 builds for `cs`):
// torch.optim.Adam's update for EVERY parameter tensor of the network in ONE launch.
//
// HBM-bound: 28 bytes per parameter (read p, g, m, v; write p, m, v).  torch's multi-tensor kernel needs six launches for
// the ~190 tensors of RN50-DeepLabv3+ (its argument block holds a few dozen tensors) and ran at 2.7 TB/s in the step's
// profile.  Here the whole tensor list travels in the kernel's parameter space (11 KB of the 32 KB CUDA 12 allows; it is
// captured by value into the step's CUDA graph, so nothing on the device can go stale), the work is cut into 16 K-element
// chunks spread over all SMs, and each thread keeps four 16-byte loads per array in flight.
//
// Arithmetic = torch's fused kernel (ATen/native/cuda/fused_adam_utils.cuh, ADAM mode, no amsgrad / maximize / grad scaler),
// all in fp32 with the hyper-parameters rounded to fp32 first and the same fused multiply-adds:
//   g = fma(p, wd, g);  m = fma(b1, m, fma(-b1, g, g));  v = fma(b2, v, fma(-b2, g*g, g*g))
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)            t = *step (already incremented by the caller)
#include "pp_common.cuh"

namespace pp {

constexpr int kAdamMaxTensors = 256;
constexpr int kAdamMaxGroups = 8;
constexpr int kAdamChunk = 16384;  // elements per work item
constexpr int kAdamThreads = 256;

struct AdamArgs {
  float* p[kAdamMaxTensors];
  const float* g[kAdamMaxTensors];
  float* m[kAdamMaxTensors];
  float* v[kAdamMaxTensors];
  long long n[kAdamMaxTensors];
  int chunk_begin[kAdamMaxTensors + 1];  // prefix sum of the tensors' chunk counts
  unsigned char group[kAdamMaxTensors];
  const float* lr[kAdamMaxGroups];  // device scalars (a scheduler rewrites them between graph replays)
  float beta1[kAdamMaxGroups], beta2[kAdamMaxGroups], eps[kAdamMaxGroups], wd[kAdamMaxGroups];
  const float* step;  // device scalar, fp32 (torch's capturable step counter)
  int n_tensors;
};

struct AdamCoef {
  float b1, b2, step_size, bc2_sqrt, eps, wd;
};

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamCoef& c) {
  if (c.wd != 0.f) g = fmaf(p, c.wd, g);
  m = fmaf(c.b1, m, fmaf(-c.b1, g, g));
  const float gg = g * g;
  v = fmaf(c.b2, v, fmaf(-c.b2, gg, gg));
  const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), c.bc2_sqrt), c.eps);
  p = __fsub_rn(p, __fdiv_rn(__fmul_rn(c.step_size, m), denom));
}

__global__ void __launch_bounds__(kAdamThreads) adam_step_multi_kernel(const __grid_constant__ AdamArgs a) {
  const int total_chunks = a.chunk_begin[a.n_tensors];
  __shared__ AdamCoef coef[kAdamMaxGroups];
  if (threadIdx.x < kAdamMaxGroups && a.lr[threadIdx.x] != nullptr) {
    const int gi = threadIdx.x;
    const float t = __ldg(a.step);
    const float bc1 = 1.f - powf(a.beta1[gi], t), bc2 = 1.f - powf(a.beta2[gi], t);
    AdamCoef c;
    c.b1 = a.beta1[gi]; c.b2 = a.beta2[gi];
    c.step_size = __fdiv_rn(__ldg(a.lr[gi]), bc1);
    c.bc2_sqrt = sqrtf(bc2);
    c.eps = a.eps[gi]; c.wd = a.wd[gi];
    coef[gi] = c;
  }
  __syncthreads();
  for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    int lo = 0, hi = a.n_tensors - 1;  // last tensor whose first chunk is <= chunk
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (a.chunk_begin[mid] <= chunk) lo = mid; else hi = mid - 1;
    }
    const int r = lo;
    const AdamCoef c = coef[a.group[r]];
    const long long off = (long long)(chunk - a.chunk_begin[r]) * kAdamChunk;
    const long long rem = a.n[r] - off;
    const int len = rem < kAdamChunk ? (int)rem : kAdamChunk;
    float* p = a.p[r] + off;
    const float* g = a.g[r] + off;
    float* m = a.m[r] + off;
    float* v = a.v[r] + off;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15u) == 0;
    if (vec) {
      const int n4 = len >> 2;
      float4* p4 = reinterpret_cast<float4*>(p);
      const float4* g4 = reinterpret_cast<const float4*>(g);
      float4* m4 = reinterpret_cast<float4*>(m);
      float4* v4 = reinterpret_cast<float4*>(v);
      for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * kAdamThreads) {
        float4 pq[4], gg[4], mm[4], vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * kAdamThreads;
          if (i < n4) { pq[u] = p4[i]; gg[u] = g4[i]; mm[u] = m4[i]; vv[u] = v4[i]; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * kAdamThreads;
          if (i < n4) {
            adam_update(pq[u].x, gg[u].x, mm[u].x, vv[u].x, c);
            adam_update(pq[u].y, gg[u].y, mm[u].y, vv[u].y, c);
            adam_update(pq[u].z, gg[u].z, mm[u].z, vv[u].z, c);
            adam_update(pq[u].w, gg[u].w, mm[u].w, vv[u].w, c);
            p4[i] = pq[u]; m4[i] = mm[u]; v4[i] = vv[u];
          }
        }
      }
      for (int i = (n4 << 2) + threadIdx.x; i < len; i += kAdamThreads) {
        float pv = p[i], mv = m[i], vv = v[i];
        adam_update(pv, g[i], mv, vv, c);
        p[i] = pv; m[i] = mv; v[i] = vv;
      }
    } else {
      for (int i = threadIdx.x; i < len; i += kAdamThreads) {
        float pv = p[i], mv = m[i], vv = v[i];
        adam_update(pv, g[i], mv, vv, c);
        p[i] = pv; m[i] = mv; v[i] = vv;
      }
    }
  }
}

}  // namespace pp

using namespace pp;

extern "C" {

int pp_adam_step_multi(int n_tensors, void* const* params, const void* const* grads, void* const* exp_avg,
                       void* const* exp_avg_sq, const long long* numel, const int* group, int n_groups,
                       const float* const* lr, const double* beta1, const double* beta2, const double* eps,
                       const double* weight_decay, const float* step, void* stream) {
  PP_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && numel && group && lr && beta1 && beta2 && eps && weight_decay && step,
               "pp_adam_step_multi: null pointer");
  PP_CHECK_ARG(n_tensors > 0 && n_groups > 0 && n_groups <= kAdamMaxGroups, "pp_adam_step_multi: %d tensors, %d groups (<= %d)",
               n_tensors, n_groups, kAdamMaxGroups);
  for (int gi = 0; gi < n_groups; ++gi)
    PP_CHECK_ARG(lr[gi] && beta1[gi] >= 0.0 && beta1[gi] < 1.0 && beta2[gi] >= 0.0 && beta2[gi] < 1.0 && eps[gi] >= 0.0,
                 "pp_adam_step_multi: bad hyper-parameters in group %d", gi);
  for (int i = 0; i < n_tensors; ++i)
    PP_CHECK_ARG(params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i] && numel[i] > 0 && group[i] >= 0 && group[i] < n_groups,
                 "pp_adam_step_multi: tensor %d: null pointer, empty tensor or bad group", i);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  for (int base = 0; base < n_tensors; base += kAdamMaxTensors) {
    const int cnt = n_tensors - base < kAdamMaxTensors ? n_tensors - base : kAdamMaxTensors;
    AdamArgs a{};
    long long chunks = 0;
    for (int i = 0; i < cnt; ++i) {
      a.p[i] = reinterpret_cast<float*>(params[base + i]);
      a.g[i] = reinterpret_cast<const float*>(grads[base + i]);
      a.m[i] = reinterpret_cast<float*>(exp_avg[base + i]);
      a.v[i] = reinterpret_cast<float*>(exp_avg_sq[base + i]);
      a.n[i] = numel[base + i];
      a.group[i] = (unsigned char)group[base + i];
      a.chunk_begin[i] = (int)chunks;
      chunks += (numel[base + i] + kAdamChunk - 1) / kAdamChunk;
      PP_CHECK_ARG(chunks < (1ll << 30), "pp_adam_step_multi: too many elements in one launch");
    }
    a.chunk_begin[cnt] = (int)chunks;
    for (int gi = 0; gi < n_groups; ++gi) {
      a.lr[gi] = lr[gi];
      a.beta1[gi] = (float)beta1[gi]; a.beta2[gi] = (float)beta2[gi]; a.eps[gi] = (float)eps[gi]; a.wd[gi] = (float)weight_decay[gi];
    }
    a.step = step;
    a.n_tensors = cnt;
    const long long cap = (long long)sms * 8;
    const int grid = (int)(chunks < cap ? chunks : cap);
    adam_step_multi_kernel<<<grid, kAdamThreads, 0, st>>>(a);
    PP_LAUNCH_CHECK();
  }
  return PP_OK;
}

}  // extern "C"
