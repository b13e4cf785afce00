// T path — the HBM-bound companions of the tensor-core convolutions, NHWC bf16, fused per layer:
//   bn_stats        per-channel sum / sum-of-squares of a raw conv output (train-mode BatchNorm statistics)
//   bn_apply        y = dropout(relu(raw * scale + shift))                       (aspp.py:16-20,75-79; decoders.py:107-114)
//   bn_bwd_reduce   g = dy * dropmask * relu'(.) ;  sum_c g, sum_c g*xhat          (BatchNorm backward, pass 1)
//   bn_bwd_apply    d_raw = scale * (g - mean(g) - xhat * mean(g*xhat))           (BatchNorm backward, pass 2)
//   upsample_nhwc   bilinear align_corners=True into a channel slice of the decoder input (deeplab.py:49-50) + adjoint
//   nchw->nhwc      layout/precision change at the backbone boundary
// One thread owns 8 consecutive channels (one 16-byte vector) of a pixel row; warps read/write whole
// 128-byte lines.  Dropout masks come from a counter-based Philox4x32-10 keyed by (seed, layer offset,
// element index), so the backward pass regenerates them instead of storing them.
#include "pp_common.cuh"
#include <cooperative_groups.h>

namespace pp {

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

// keep-mask bits for the 8 channels starting at element index e (multiple of 8): bit j set = keep
__device__ __forceinline__ uint32_t dropout_keep8(uint64_t seed, uint64_t offset, uint64_t e, float p) {
  const uint64_t ctr = offset + (e >> 3);
  const uint4 a = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint4 b = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 1u, 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint32_t thr = (uint32_t)fminf(p * 4294967296.0f, 4294967295.0f);
  uint32_t m = 0;
  m |= (a.x >= thr) << 0; m |= (a.y >= thr) << 1; m |= (a.z >= thr) << 2; m |= (a.w >= thr) << 3;
  m |= (b.x >= thr) << 4; m |= (b.y >= thr) << 5; m |= (b.z >= thr) << 6; m |= (b.w >= thr) << 7;
  return m;
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&t);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

constexpr int kEwThreads = 256;

// 16-byte read-only streaming load (no L1 allocation: every activation byte is touched once per pass)
__device__ __forceinline__ uint4 ldg_stream_u4(const __nv_bfloat16* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

// ---- BN statistics -------------------------------------------------------------------------
// sums[0][c] += sum_rows raw, sums[1][c] += sum_rows raw^2   (fp32, caller zeroes)
__device__ __forceinline__ void bn_stats_body(const __nv_bfloat16* __restrict__ raw, int64_t M, int ld, int c_off, int C,
                                              float* __restrict__ sums, float* sh /*[kEwThreads][16]*/) {
  const int groups = C >> 3;     // 8-channel groups per row
  const int rows_per_block = kEwThreads / groups;
  const int g = threadIdx.x % groups, r = threadIdx.x / groups;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  if (r < rows_per_block) {
    // 8 independent 16-byte loads in flight per thread (the bounds check of a plain strided loop serialises them)
    const int64_t stride = (int64_t)gridDim.x * rows_per_block;
    const __nv_bfloat16* base = raw + c_off + g * 8;
    int64_t row = (int64_t)blockIdx.x * rows_per_block + r;
    auto acc8 = [&](const uint4& v) {
      float f[8];
      unpack8(v, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += f[i];
        q[i] = fmaf(f[i], f[i], q[i]);
      }
    };
    for (; row + 7 * stride < M; row += 8 * stride) {
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = ldg_stream_u4(base + (row + u * stride) * ld);
#pragma unroll
      for (int u = 0; u < 8; ++u) acc8(v[u]);
    }
    if (row < M) {  // tail: one predicated batch (bf16 zeros add nothing to either sum)
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (row + u * stride < M) ? ldg_stream_u4(base + (row + u * stride) * ld) : z;
#pragma unroll
      for (int u = 0; u < 8; ++u) acc8(v[u]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sh[threadIdx.x * 16 + i] = s[i];
    sh[threadIdx.x * 16 + 8 + i] = q[i];
  }
  __syncthreads();
  // thread t < groups*16 reduces (group, slot) over the rows of the block
  for (int t = threadIdx.x; t < groups * 16; t += kEwThreads) {
    const int gg = t / 16, slot = t % 16;
    float acc = 0.f;
    for (int rr = 0; rr < rows_per_block; ++rr) acc += sh[(rr * groups + gg) * 16 + slot];
    const int c = gg * 8 + (slot & 7);
    atomicAdd(sums + (slot >> 3) * C + c, acc);
  }
}
__global__ void __launch_bounds__(kEwThreads) bn_stats_kernel(const __nv_bfloat16* __restrict__ raw, int64_t M, int ld,
                                                              int c_off, int C, float* __restrict__ sums) {
  extern __shared__ float sh[];
  bn_stats_body(raw, M, ld, c_off, C, sums, sh);
}

// ---- BN apply (+ReLU, +dropout) --------------------------------------------------------------
struct BnApplyParams {
  const __nv_bfloat16* raw;
  int64_t M;
  int ld_in, c_off_in, C;
  const float* scale;
  const float* shift;
  int relu;
  float drop_p;
  uint64_t seed, offset;
  const uint64_t* seed_dev;  // optional device-side addend to the seed (CUDA-graph replays advance it)
  __nv_bfloat16* out;
  int ld_out, c_off_out;
  const __nv_bfloat16* res;  // optional residual added BEFORE the activation (resnet_models.py:88-92), [M][ld_res]
  int ld_res;
};
template <bool RES, bool DROP>
__device__ __forceinline__ void bn_apply_body(const BnApplyParams& p, const float (&sc)[8], const float (&sf)[8]) {
  // thread -> fixed 8-channel group (scale/shift live in registers), rows strided over the grid
  const int groups = p.C >> 3;
  const int rows_per_block = kEwThreads / groups;
  const int g = threadIdx.x % groups, r = threadIdx.x / groups;
  if (r >= rows_per_block) return;
  const float keep_scale = DROP ? 1.f / (1.f - p.drop_p) : 1.f;
  const uint64_t seed = p.seed + ((DROP && p.seed_dev) ? *p.seed_dev : 0ull);
  const int64_t stride = (int64_t)gridDim.x * rows_per_block;
  const __nv_bfloat16* base = p.raw + p.c_off_in + g * 8;
  const __nv_bfloat16* rbase = RES ? p.res + g * 8 : nullptr;
  auto finish = [&](int64_t row, const uint4& v, const uint4& rv) {
    float f[8], rs[8];
    unpack8(v, f);
    if (RES) unpack8(rv, rs);
    uint32_t keep = 0xFFu;
    if (DROP) keep = dropout_keep8(seed, p.offset, (uint64_t)(row * groups + g) * 8, p.drop_p);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float y = fmaf(f[j], sc[j], sf[j]);
      if (RES) y += rs[j];
      if (p.relu) y = fmaxf(y, 0.f);
      if (p.relu == 2) y = fminf(y, 6.f);  // ReLU6 (mobilenet_v2.py:7-12)
      f[j] = ((keep >> j) & 1u) ? y * keep_scale : 0.f;
    }
    *reinterpret_cast<uint4*>(p.out + row * p.ld_out + p.c_off_out + g * 8) = pack8(f);
  };
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  int64_t row = (int64_t)blockIdx.x * rows_per_block + r;
  // 4 rows (x2 tensors with a residual) in flight per thread
  for (; row + 3 * stride < p.M; row += 4 * stride) {
    uint4 v[4], rv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u] = ldg_stream_u4(base + (row + u * stride) * p.ld_in);
      rv[u] = RES ? ldg_stream_u4(rbase + (row + u * stride) * p.ld_res) : z;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) finish(row + u * stride, v[u], rv[u]);
  }
  if (row < p.M) {  // tail: one predicated batch, loads still issued together
    uint4 v[4], rv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool ok = row + u * stride < p.M;
      v[u] = ok ? ldg_stream_u4(base + (row + u * stride) * p.ld_in) : z;
      rv[u] = (RES && ok) ? ldg_stream_u4(rbase + (row + u * stride) * p.ld_res) : z;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (row + u * stride < p.M) finish(row + u * stride, v[u], rv[u]);
  }
}
template <bool RES, bool DROP>
__global__ void __launch_bounds__(kEwThreads, 4) bn_apply_kernel(const BnApplyParams p) {
  const int g = threadIdx.x % (p.C >> 3);
  float sc[8], sf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = __ldg(p.scale + g * 8 + j);
    sf[j] = __ldg(p.shift + g * 8 + j);
  }
  bn_apply_body<RES, DROP>(p, sc, sf);
}

// finalize + apply in one launch: the per-channel sums come from the conv epilogue (pp_conv_igemm_stats); every thread derives the
// scale / shift of its 8 channels itself (same arithmetic as bn_finalize_kernel) and block 0 also publishes (scale, shift, mean,
// rstd) for the backward pass and updates the running statistics - one launch less per BatchNorm layer.
struct BnFinalizeArgs {
  const float* sums;  // [2][C]
  const float* gamma;
  const float* beta;
  float inv_m, unbias, eps, momentum;
  float* running_mean;  // may be null
  float* running_var;
  float* stats_out;  // [4][C]
};
template <bool RES>
__global__ void __launch_bounds__(kEwThreads, 4) bn_apply_stats_kernel(const BnApplyParams p, const BnFinalizeArgs a) {
  const int groups = p.C >> 3;
  const int g = threadIdx.x % groups;
  float sc[8], sf[8];
  const bool publish = blockIdx.x == 0 && threadIdx.x < groups;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = g * 8 + j;
    const float mean = __ldg(a.sums + c) * a.inv_m;
    float var = fmaf(-mean, mean, __ldg(a.sums + p.C + c) * a.inv_m);
    var = fmaxf(var, 0.f);
    const float rstd = rsqrtf(var + a.eps);
    sc[j] = __ldg(a.gamma + c) * rstd;
    sf[j] = fmaf(-mean, sc[j], __ldg(a.beta + c));
    if (publish) {
      a.stats_out[c] = sc[j];
      a.stats_out[p.C + c] = sf[j];
      a.stats_out[2 * p.C + c] = mean;
      a.stats_out[3 * p.C + c] = rstd;
      if (a.running_mean) {
        a.running_mean[c] = fmaf(a.momentum, mean - a.running_mean[c], a.running_mean[c]);
        a.running_var[c] = fmaf(a.momentum, var * a.unbias - a.running_var[c], a.running_var[c]);
      }
    }
  }
  bn_apply_body<RES, false>(p, sc, sf);
}

// ---- BN backward -----------------------------------------------------------------------------
struct BnBwdParams {
  const __nv_bfloat16* dy;  // grad wrt the layer output (after dropout), [M][ld_dy] slice c_off_dy
  int ld_dy, c_off_dy;
  const __nv_bfloat16* raw;  // saved raw conv output
  int ld_raw, c_off_raw;
  int64_t M;
  int C;
  const float* scale;  // gamma * rstd
  const float* shift;  // beta - mean * scale
  const float* mean;
  const float* rstd;
  int relu;
  float drop_p;
  uint64_t seed, offset;
  const uint64_t* seed_dev;
  float* sums;          // [2][C]: sum g, sum g*xhat
  __nv_bfloat16* draw;  // pass 2 output [M][ld_draw] at channel offset c_off_draw
  int ld_draw, c_off_draw;
  const __nv_bfloat16* res;  // optional residual that was added before the activation (gate = act'(bn(x) + res))
  int ld_res;
  __nv_bfloat16* dres;  // optional pass 2 output [M][C]: gradient wrt the residual (= gated upstream gradient)
};

// masked upstream gradient of 8 channels: dropout mask/scale and the ReLU / ReLU6 gate (recomputed, never stored)
struct BnBwdVec {
  uint4 dy, x, res;
};
template <bool RES, bool PRED = false>
__device__ __forceinline__ BnBwdVec bn_bwd_load(const BnBwdParams& p, int64_t row, int g) {
  BnBwdVec v;
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  const bool ok = !PRED || row < p.M;  // tail batches past the end load nothing
  v.dy = ok ? ldg_stream_u4(p.dy + row * p.ld_dy + p.c_off_dy + g * 8) : z;
  v.x = ok ? ldg_stream_u4(p.raw + row * p.ld_raw + p.c_off_raw + g * 8) : z;
  v.res = (RES && ok) ? ldg_stream_u4(p.res + row * p.ld_res + g * 8) : z;
  return v;
}
template <bool RES, bool DROP>
__device__ __forceinline__ void bn_bwd_g8(const BnBwdParams& p, uint64_t seed, float keep_scale, int64_t row, int g,
                                          int groups, const float (&sc)[8], const float (&sf)[8], const BnBwdVec& v,
                                          float (&x)[8], float (&go)[8]) {
  float dy[8], rs[8];
  unpack8(v.dy, dy);
  unpack8(v.x, x);
  if (RES) unpack8(v.res, rs);
  uint32_t keep = 0xFFu;
  if (DROP) keep = dropout_keep8(seed, p.offset, (uint64_t)(row * groups + g) * 8, p.drop_p);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float gg = ((keep >> j) & 1u) ? dy[j] * keep_scale : 0.f;
    float yv = fmaf(x[j], sc[j], sf[j]);
    if (RES) yv += rs[j];
    if (p.relu && !(yv > 0.f)) gg = 0.f;
    if (p.relu == 2 && !(yv < 6.f)) gg = 0.f;
    go[j] = gg;
  }
}

template <bool RES, bool DROP>
__device__ __forceinline__ void bn_bwd_reduce_body(const BnBwdParams& p, float* __restrict__ sums, float* sh) {
  const int groups = p.C >> 3;
  const int rows_per_block = kEwThreads / groups;
  const int g = threadIdx.x % groups, r = threadIdx.x / groups;
  const float keep_scale = DROP ? 1.f / (1.f - p.drop_p) : 1.f;
  const uint64_t seed = p.seed + ((DROP && p.seed_dev) ? *p.seed_dev : 0ull);
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  if (r < rows_per_block) {
    float sc[8], sf[8], mu[8], rs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = __ldg(p.scale + g * 8 + j);
      sf[j] = __ldg(p.shift + g * 8 + j);
      mu[j] = __ldg(p.mean + g * 8 + j);
      rs[j] = __ldg(p.rstd + g * 8 + j);
    }
    const int64_t stride = (int64_t)gridDim.x * rows_per_block;
    auto acc = [&](int64_t row, const BnBwdVec& v) {
      float x[8], go[8];
      bn_bwd_g8<RES, DROP>(p, seed, keep_scale, row, g, groups, sc, sf, v, x, go);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += go[j];
        q[j] = fmaf(go[j], (x[j] - mu[j]) * rs[j], q[j]);
      }
    };
    int64_t row = (int64_t)blockIdx.x * rows_per_block + r;
    for (; row + 3 * stride < p.M; row += 4 * stride) {  // 4 rows x (dy, raw[, res]) in flight per thread
      BnBwdVec v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = bn_bwd_load<RES>(p, row + u * stride, g);
#pragma unroll
      for (int u = 0; u < 4; ++u) acc(row + u * stride, v[u]);
    }
    if (row < p.M) {
      BnBwdVec v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = bn_bwd_load<RES, true>(p, row + u * stride, g);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (row + u * stride < p.M) acc(row + u * stride, v[u]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sh[threadIdx.x * 16 + i] = s[i];
    sh[threadIdx.x * 16 + 8 + i] = q[i];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < groups * 16; t += kEwThreads) {
    const int gg = t / 16, slot = t % 16;
    float acc = 0.f;
    for (int rr = 0; rr < rows_per_block; ++rr) acc += sh[(rr * groups + gg) * 16 + slot];
    atomicAdd(sums + (slot >> 3) * p.C + gg * 8 + (slot & 7), acc);
  }
}
template <bool RES, bool DROP>
__global__ void __launch_bounds__(kEwThreads, 3) bn_bwd_reduce_kernel(const BnBwdParams p) {
  extern __shared__ float sh[];
  bn_bwd_reduce_body<RES, DROP>(p, p.sums, sh);
}

template <bool RES, bool DROP, int BATCH, bool SUMS_SHARED = false>
__device__ __forceinline__ void bn_bwd_apply_body(const BnBwdParams& p, const float* __restrict__ sums) {
  // d_raw = scale * (g - mean(g) - xhat * mean(g xhat)) = A*g + B*x + K with per-channel A, B, K held in registers
  const int groups = p.C >> 3;
  const int rows_per_block = kEwThreads / groups;
  const int g = threadIdx.x % groups, r = threadIdx.x / groups;
  if (r >= rows_per_block) return;
  const float inv_m = 1.f / (float)p.M;
  const float keep_scale = DROP ? 1.f / (1.f - p.drop_p) : 1.f;
  const uint64_t seed = p.seed + ((DROP && p.seed_dev) ? *p.seed_dev : 0ull);
  float sc[8], sf[8], cb[8], ck[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = g * 8 + j;
    sc[j] = __ldg(p.scale + c);
    sf[j] = __ldg(p.shift + c);
    const float mu = __ldg(p.mean + c), rs = __ldg(p.rstd + c);
    const float mg = (SUMS_SHARED ? sums[c] : __ldcg(sums + c)) * inv_m;
    const float mgx = (SUMS_SHARED ? sums[p.C + c] : __ldcg(sums + p.C + c)) * inv_m;
    cb[j] = -sc[j] * rs * mgx;
    ck[j] = -sc[j] * (mg - mu * rs * mgx);
  }
  const int64_t stride = (int64_t)gridDim.x * rows_per_block;
  auto finish = [&](int64_t row, const BnBwdVec& v) {
    float x[8], gg[8], o[8];
    bn_bwd_g8<RES, DROP>(p, seed, keep_scale, row, g, groups, sc, sf, v, x, gg);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(sc[j], gg[j], fmaf(cb[j], x[j], ck[j]));
    *reinterpret_cast<uint4*>(p.draw + row * p.ld_draw + p.c_off_draw + g * 8) = pack8(o);
    if (RES && p.dres) *reinterpret_cast<uint4*>(p.dres + row * p.C + g * 8) = pack8(gg);
  };
  int64_t row = (int64_t)blockIdx.x * rows_per_block + r;
  for (; row + (BATCH - 1) * stride < p.M; row += BATCH * stride) {
    BnBwdVec v[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; ++u) v[u] = bn_bwd_load<RES>(p, row + u * stride, g);
#pragma unroll
    for (int u = 0; u < BATCH; ++u) finish(row + u * stride, v[u]);
  }
  if (row < p.M) {
    BnBwdVec v[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; ++u) v[u] = bn_bwd_load<RES, true>(p, row + u * stride, g);
#pragma unroll
    for (int u = 0; u < BATCH; ++u)
      if (row + u * stride < p.M) finish(row + u * stride, v[u]);
  }
}
template <bool RES, bool DROP>
__global__ void __launch_bounds__(kEwThreads, 4) bn_bwd_apply_kernel(const BnBwdParams p) {
  bn_bwd_apply_body<RES, DROP, 2>(p, p.sums);
}

// ---- single-launch BatchNorm passes (cooperative grid barrier) ------------------------------------------
// The 59-61 BatchNorm layers of a train step are each "reduce over everything, then touch everything again": as
// separate kernels that is memset + stats + finalize + apply (+ memset + reduce + apply backward) = 7 launches per
// layer, and at the reference batch (4 images) the step is launch-bound.  Launched cooperatively (all CTAs
// co-resident), one kernel does both passes around a grid-wide barrier: 2 launches per layer, the second pass of a
// small tensor re-reads L2.  scratch = 2*C fp32 partial sums + 2 u32 counters, zero before the launch; the last CTA to
// leave zeroes it again, so one cudaMemset at allocation serves every later launch.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned n_ctas) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < n_ctas);
  }
  __syncthreads();
}
__device__ __forceinline__ void scratch_release(float* sums, int n_sums, unsigned* bar, unsigned n_ctas) {
  __shared__ unsigned last;
  __syncthreads();  // every thread of this CTA is done reading the sums
  if (threadIdx.x == 0) {
    __threadfence();
    last = (atomicAdd(bar + 1, 1u) == n_ctas - 1u) ? 1u : 0u;
  }
  __syncthreads();
  if (last) {
    for (int i = threadIdx.x; i < n_sums; i += blockDim.x) sums[i] = 0.f;
    if (threadIdx.x == 0) {
      bar[0] = 0u;
      bar[1] = 0u;
    }
  }
}

struct BnFwdFusedParams {
  BnApplyParams a;  // a.scale / a.shift unused
  const float* gamma;
  const float* beta;
  float* running_mean;  // nullable
  float* running_var;
  long long* nbt;  // nullable: num_batches_tracked += 1
  float eps, momentum;
  float* stats_out;  // [4][C]: scale, shift, mean, rstd (what the backward needs)
  float* sums;       // scratch [2][C]
  unsigned* bar;     // scratch [2]
};
__device__ __forceinline__ void bn_finalize_one(const float* sums, int C, int c, float inv_m, float eps, float gamma,
                                                float beta, float& scale, float& shift, float& mean, float& var,
                                                float& rstd) {
  mean = __ldcg(sums + c) * inv_m;
  var = fmaxf(fmaf(-mean, mean, __ldcg(sums + C + c) * inv_m), 0.f);
  rstd = rsqrtf(var + eps);
  scale = gamma * rstd;
  shift = fmaf(-mean, scale, beta);
}
template <bool RES, bool DROP>
__global__ void __launch_bounds__(kEwThreads, 4) bn_fwd_fused_kernel(const BnFwdFusedParams q) {
  extern __shared__ float sh[];
  const BnApplyParams& p = q.a;
  bn_stats_body(p.raw, p.M, p.ld_in, p.c_off_in, p.C, q.sums, sh);
  grid_barrier(q.bar, gridDim.x);
  const float inv_m = 1.f / (float)p.M;
  const int g = threadIdx.x % (p.C >> 3);
  float sc[8], sf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = g * 8 + j;
    float mean, var, rstd;
    bn_finalize_one(q.sums, p.C, c, inv_m, q.eps, __ldg(q.gamma + c), __ldg(q.beta + c), sc[j], sf[j], mean, var, rstd);
  }
  if (blockIdx.x == 0) {
    const float unbias = p.M > 1 ? (float)p.M / (float)(p.M - 1) : 1.f;
    for (int c = threadIdx.x; c < p.C; c += kEwThreads) {
      float scale, shift, mean, var, rstd;
      bn_finalize_one(q.sums, p.C, c, inv_m, q.eps, __ldg(q.gamma + c), __ldg(q.beta + c), scale, shift, mean, var, rstd);
      q.stats_out[c] = scale;
      q.stats_out[p.C + c] = shift;
      q.stats_out[2 * p.C + c] = mean;
      q.stats_out[3 * p.C + c] = rstd;
      if (q.running_mean) {
        q.running_mean[c] = fmaf(q.momentum, mean - q.running_mean[c], q.running_mean[c]);
        q.running_var[c] = fmaf(q.momentum, var * unbias - q.running_var[c], q.running_var[c]);
      }
    }
    if (threadIdx.x == 0 && q.nbt) *q.nbt += 1;
  }
  bn_apply_body<RES, DROP>(p, sc, sf);
  scratch_release(q.sums, 2 * p.C, q.bar, gridDim.x);
}

struct BnBwdFusedParams {
  BnBwdParams b;  // b.sums = OUTPUT copy of (d beta, d gamma) [2][C]
  float* sums;    // scratch [2][C]
  unsigned* bar;  // scratch [2]
};
template <bool RES, bool DROP>
__global__ void __launch_bounds__(kEwThreads, 2) bn_bwd_fused_kernel(const BnBwdFusedParams q) {
  extern __shared__ float sh[];
  bn_bwd_reduce_body<RES, DROP>(q.b, q.sums, sh);
  grid_barrier(q.bar, gridDim.x);
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < 2 * q.b.C; i += kEwThreads) q.b.sums[i] = __ldcg(q.sums + i);
  bn_bwd_apply_body<RES, DROP, 4>(q.b, q.sums);
  scratch_release(q.sums, 2 * q.b.C, q.bar, gridDim.x);
}

// ---- small tensors: the same two passes inside ONE thread-block cluster ---------------------------------------
// For very small tensors (< 0.5 MB: the 1/16-resolution projections at the reference batch of 4) the cooperative
// kernels above cost ~14 us, nearly all of it launch + global-atomic + grid-barrier latency.  A cluster of <= 8 CTAs is
// co-scheduled by the hardware (ordinary launch), reduces its partial sums through distributed shared memory and
// synchronises with the hardware cluster barrier: no global atomics, no scratch, no cooperative launch.  Above that size
// 8 SMs cannot stream fast enough (~50 GB/s each) and the grid-wide version wins.
namespace cg = cooperative_groups;
constexpr int kBnClusterMax = 8;

// every CTA sums the cluster's per-CTA partials [2C] into its own shared tot[2C]
__device__ __forceinline__ void cluster_totals(cg::cluster_group& cluster, float* part, float* tot, int n2c) {
  const unsigned n = cluster.num_blocks();
  for (int i = threadIdx.x; i < n2c; i += blockDim.x) {
    float a = 0.f;
    for (unsigned r = 0; r < n; ++r) a += cluster.map_shared_rank(part, r)[i];
    tot[i] = a;
  }
  __syncthreads();
}

template <bool RES, bool DROP>
__global__ void __launch_bounds__(kEwThreads, 4) bn_fwd_cluster_kernel(const BnFwdFusedParams q) {
  extern __shared__ float sh[];  // [kEwThreads * 16] reduce scratch | part [2C] | tot [2C]
  const BnApplyParams& p = q.a;
  float* part = sh + kEwThreads * 16;
  float* tot = part + 2 * p.C;
  cg::cluster_group cluster = cg::this_cluster();
  for (int i = threadIdx.x; i < 2 * p.C; i += kEwThreads) part[i] = 0.f;
  __syncthreads();
  bn_stats_body(p.raw, p.M, p.ld_in, p.c_off_in, p.C, part, sh);  // the CTA's partial sums land in its shared memory
  cluster.sync();
  cluster_totals(cluster, part, tot, 2 * p.C);
  const float inv_m = 1.f / (float)p.M;
  const int g = threadIdx.x % (p.C >> 3);
  float sc[8], sf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = g * 8 + j;
    const float mean = tot[c] * inv_m;
    const float var = fmaxf(fmaf(-mean, mean, tot[p.C + c] * inv_m), 0.f);
    sc[j] = __ldg(q.gamma + c) * rsqrtf(var + q.eps);
    sf[j] = fmaf(-mean, sc[j], __ldg(q.beta + c));
  }
  if (blockIdx.x == 0) {
    const float unbias = p.M > 1 ? (float)p.M / (float)(p.M - 1) : 1.f;
    for (int c = threadIdx.x; c < p.C; c += kEwThreads) {
      const float mean = tot[c] * inv_m;
      const float var = fmaxf(fmaf(-mean, mean, tot[p.C + c] * inv_m), 0.f);
      const float rstd = rsqrtf(var + q.eps);
      const float scale = __ldg(q.gamma + c) * rstd;
      q.stats_out[c] = scale;
      q.stats_out[p.C + c] = fmaf(-mean, scale, __ldg(q.beta + c));
      q.stats_out[2 * p.C + c] = mean;
      q.stats_out[3 * p.C + c] = rstd;
      if (q.running_mean) {
        q.running_mean[c] = fmaf(q.momentum, mean - q.running_mean[c], q.running_mean[c]);
        q.running_var[c] = fmaf(q.momentum, var * unbias - q.running_var[c], q.running_var[c]);
      }
    }
    if (threadIdx.x == 0 && q.nbt) *q.nbt += 1;
  }
  bn_apply_body<RES, DROP>(p, sc, sf);
  cluster.sync();  // no CTA may exit while a peer can still read its shared memory
}

template <bool RES, bool DROP>
__global__ void __launch_bounds__(kEwThreads, 2) bn_bwd_cluster_kernel(const BnBwdFusedParams q) {
  extern __shared__ float sh[];
  const BnBwdParams& p = q.b;
  float* part = sh + kEwThreads * 16;
  float* tot = part + 2 * p.C;
  cg::cluster_group cluster = cg::this_cluster();
  for (int i = threadIdx.x; i < 2 * p.C; i += kEwThreads) part[i] = 0.f;
  __syncthreads();
  bn_bwd_reduce_body<RES, DROP>(p, part, sh);
  cluster.sync();
  cluster_totals(cluster, part, tot, 2 * p.C);
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < 2 * p.C; i += kEwThreads) p.sums[i] = tot[i];
  bn_bwd_apply_body<RES, DROP, 4, true>(p, tot);
  cluster.sync();
}

// ---- bilinear upsample NHWC bf16 (align_corners=True) -----------------------------------------
__global__ void __launch_bounds__(kEwThreads) upsample_nhwc_kernel(const __nv_bfloat16* __restrict__ in, int N, int h, int w,
                                                                   int C, int ld_in, __nv_bfloat16* __restrict__ out, int H,
                                                                   int W, int ld_out, int c_off, float sh_, float sw_) {
  const int groups = C >> 3;
  const int64_t total = (int64_t)N * H * W * groups;
  for (int64_t i = (int64_t)blockIdx.x * kEwThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kEwThreads) {
    const int g = (int)(i % groups);
    int64_t t = i / groups;
    const int x = (int)(t % W);
    t /= W;
    const int y = (int)(t % H);
    const int n = (int)(t / H);
    const Lerp ly = lerp_ac(y, h, H, sh_), lx = lerp_ac(x, w, W, sw_);
    const __nv_bfloat16* b = in + (int64_t)n * h * w * ld_in + g * 8;
    float v00[8], v01[8], v10[8], v11[8], o[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(b + ((int64_t)ly.i0 * w + lx.i0) * ld_in)), v00);
    unpack8(__ldg(reinterpret_cast<const uint4*>(b + ((int64_t)ly.i0 * w + lx.i1) * ld_in)), v01);
    unpack8(__ldg(reinterpret_cast<const uint4*>(b + ((int64_t)ly.i1 * w + lx.i0) * ld_in)), v10);
    unpack8(__ldg(reinterpret_cast<const uint4*>(b + ((int64_t)ly.i1 * w + lx.i1) * ld_in)), v11);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o[j] = ly.l0 * (lx.l0 * v00[j] + lx.l1 * v01[j]) + ly.l1 * (lx.l0 * v10[j] + lx.l1 * v11[j]);
    *reinterpret_cast<uint4*>(out + (((int64_t)n * H + y) * W + x) * ld_out + c_off + g * 8) = pack8(o);
  }
}

// adjoint, as a GATHER: one thread owns 8 channels of one low-resolution pixel and sums the contributions of every
// high-resolution pixel whose stencil touches it (weights re-evaluated with the forward's own lerp_ac, so the pair is
// an exact adjoint).  No atomics, deterministic; grad_in (f32 [N,h,w,C]) is fully overwritten.
__device__ __forceinline__ void adj_range(int i, int in_size, int out_size, float scale, int& lo, int& hi) {
  if (in_size == out_size) {
    lo = hi = i;
  } else if (!(scale > 0.f)) {
    lo = 0;
    hi = out_size - 1;
  } else {
    lo = (int)floorf((float)(i - 1) / scale) - 1;
    hi = (int)ceilf((float)(i + 1) / scale) + 1;
    lo = lo < 0 ? 0 : lo;
    hi = hi > out_size - 1 ? out_size - 1 : hi;
  }
}
__global__ void __launch_bounds__(kEwThreads) upsample_nhwc_bwd_kernel(const __nv_bfloat16* __restrict__ gout, int N, int H,
                                                                       int W, int ld, int c_off, int C,
                                                                       float* __restrict__ gin, int h, int w, float sh_,
                                                                       float sw_) {
  const int groups = C >> 3;
  const int64_t total = (int64_t)N * h * w * groups;
  for (int64_t i = (int64_t)blockIdx.x * kEwThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kEwThreads) {
    const int g = (int)(i % groups);
    int64_t t = i / groups;
    const int xi = (int)(t % w);
    t /= w;
    const int yi = (int)(t % h);
    const int n = (int)(t / h);
    int y_lo, y_hi, x_lo, x_hi;
    adj_range(yi, h, H, sh_, y_lo, y_hi);
    adj_range(xi, w, W, sw_, x_lo, x_hi);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const __nv_bfloat16* b = gout + (int64_t)n * H * W * ld + c_off + g * 8;
    for (int y = y_lo; y <= y_hi; ++y) {
      const Lerp ly = lerp_ac(y, h, H, sh_);
      const float wy = (ly.i0 == yi ? ly.l0 : 0.f) + (ly.i1 == yi ? ly.l1 : 0.f);
      if (wy == 0.f) continue;
      for (int x = x_lo; x <= x_hi; ++x) {
        const Lerp lx = lerp_ac(x, w, W, sw_);
        const float wx = (lx.i0 == xi ? lx.l0 : 0.f) + (lx.i1 == xi ? lx.l1 : 0.f);
        if (wx == 0.f) continue;
        float gv[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(b + ((int64_t)y * W + x) * ld)), gv);
        const float wgt = wy * wx;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(gv[j], wgt, acc[j]);
      }
    }
    float* o = gin + (((int64_t)n * h + yi) * w + xi) * C + g * 8;
    *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// ---- NCHW (f32 / bf16, any strides) -> NHWC bf16 with channel padding ----------------------------
template <typename T>
__global__ void __launch_bounds__(kEwThreads) to_nhwc_bf16_kernel(const T* __restrict__ in, int64_t sn, int64_t sc, int64_t sh_,
                                                                  int64_t sw_, int N, int C, int H, int W,
                                                                  __nv_bfloat16* __restrict__ out, int ld, int c_off) {
  // tile transpose through shared memory: 32 channels x 32 pixels
  __shared__ float tile[32][33];
  const int64_t HW = (int64_t)H * W;
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * 32;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 256 threads: ty in 0..7
  for (int cc = ty; cc < 32; cc += 8) {
    const int c = c0 + cc;
    const int64_t pix = p0 + tx;
    float v = 0.f;
    if (c < C && pix < HW) {
      const int y = (int)(pix / W), x = (int)(pix - (int64_t)y * W);
      v = (float)in[(int64_t)n * sn + (int64_t)c * sc + (int64_t)y * sh_ + (int64_t)x * sw_];
    }
    tile[cc][tx] = v;
  }
  __syncthreads();
  for (int pp_ = ty; pp_ < 32; pp_ += 8) {
    const int64_t pix = p0 + pp_;
    const int c = c0 + tx;
    if (pix < HW && c < C) out[((int64_t)n * HW + pix) * ld + c_off + c] = __float2bfloat16(tile[tx][pp_]);
  }
}

static inline int ew_grid(int64_t total) {
  int64_t b = (total + kEwThreads - 1) / kEwThreads;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace pp

namespace pp {
// single-cluster launch for small tensors; returns 0 CTAs when the tensor should take the cooperative path
static int bn_cluster_ctas(int64_t M, int C) {
  // measured (scripts/bench_dw.py, B=4): one SM streams only ~50 GB/s, so 8 CTAs beat the 14 us cooperative kernel only
  // below ~0.5 MB (0.26 MB: 12.4 vs 14.1 us per fwd+bwd pair; 3.4 MB: 67 vs 36 us)
  if (C > 1024 || (int64_t)M * C * 2 > (512ll << 10)) return 0;
  const int rows_per_block = kEwThreads / (C / 8);
  const int64_t want = (M + (int64_t)rows_per_block * 8 - 1) / ((int64_t)rows_per_block * 8);  // >= 8 rows per thread
  int n = 1;
  while (n < kBnClusterMax && n < want) n <<= 1;
  return n;
}
template <typename K, typename P>
static int launch_cluster(K kernel, const P& params, int n_ctas, size_t smem, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)n_ctas);
  cfg.blockDim = dim3(kEwThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)n_ctas;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  PP_CUDA(cudaLaunchKernelEx(&cfg, kernel, params));
  PP_LAUNCH_CHECK();
  return PP_OK;
}
// cooperative launch with the grid clamped to what can be co-resident
template <typename K, typename P>
static int launch_coop(K kernel, const P& params, int64_t blocks_wanted, size_t smem, cudaStream_t st, const char* what) {
  int dev = 0, sms = 0, occ = 0;
  PP_CUDA(cudaGetDevice(&dev));
  PP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  PP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kEwThreads, smem));
  if (occ < 1) {
    set_error("%s: kernel cannot be resident", what);
    return PP_ERR_CUDA;
  }
  int64_t grid = (int64_t)occ * sms;
  if (blocks_wanted < grid) grid = blocks_wanted;
  if (grid < 1) grid = 1;
  P local = params;
  void* args[] = {&local};
  PP_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kernel), dim3((unsigned)grid), dim3(kEwThreads), args, smem,
                                      st));
  PP_LAUNCH_CHECK();
  return PP_OK;
}

}  // namespace pp

using namespace pp;

extern "C" {

int pp_bn_stats(const void* raw, int64_t M, int ld, int c_off, int C, float* sums, void* stream) {
  PP_CHECK_ARG(raw && sums && M > 0, "pp_bn_stats: bad args");
  PP_CHECK_ARG(C % 8 == 0 && C >= 8 && C <= 2048 && ld % 8 == 0 && c_off % 8 == 0,
               "pp_bn_stats: C=%d ld=%d c_off=%d must be multiples of 8 (C <= 2048)", C, ld, c_off);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  PP_CUDA(cudaMemsetAsync(sums, 0, 2 * (size_t)C * sizeof(float), st));
  const int rows_per_block = kEwThreads / (C / 8);
  int64_t blocks = (M + rows_per_block * 8 - 1) / (rows_per_block * 8);
  if (blocks > 148 * 4) blocks = 148 * 4;  // one resident wave (4 CTAs / SM), 8 loads in flight per thread
  bn_stats_kernel<<<(int)blocks, kEwThreads, kEwThreads * 16 * sizeof(float), st>>>(
      reinterpret_cast<const __nv_bfloat16*>(raw), M, ld, c_off, C, sums);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int pp_bn_apply(const void* raw, int64_t M, int ld_in, int c_off_in, int C, const float* scale, const float* shift,
                int relu, float drop_p, uint64_t seed, uint64_t offset, const uint64_t* seed_dev, void* out, int ld_out,
                int c_off_out, void* stream) {
  return pp_bn_apply_res(raw, M, ld_in, c_off_in, C, scale, shift, relu, drop_p, seed, offset, seed_dev, nullptr, 0, out,
                         ld_out, c_off_out, stream);
}

int pp_bn_apply_res(const void* raw, int64_t M, int ld_in, int c_off_in, int C, const float* scale, const float* shift,
                    int relu, float drop_p, uint64_t seed, uint64_t offset, const uint64_t* seed_dev, const void* res,
                    int ld_res, void* out, int ld_out, int c_off_out, void* stream) {
  PP_CHECK_ARG(raw && out && scale && shift && M > 0, "pp_bn_apply: bad args");
  PP_CHECK_ARG(!res || (ld_res % 8 == 0 && ld_res >= C), "pp_bn_apply: residual ld=%d", ld_res);
  PP_CHECK_ARG(C % 8 == 0 && C <= 2048 && ld_in % 8 == 0 && c_off_in % 8 == 0 && ld_out % 8 == 0 && c_off_out % 8 == 0,
               "pp_bn_apply: channel counts/offsets must be multiples of 8 (C <= 2048)");
  PP_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "pp_bn_apply: drop_p=%f", drop_p);
  BnApplyParams p;
  p.raw = reinterpret_cast<const __nv_bfloat16*>(raw);
  p.M = M; p.ld_in = ld_in; p.c_off_in = c_off_in; p.C = C; p.scale = scale; p.shift = shift; p.relu = relu;
  p.drop_p = drop_p; p.seed = seed; p.offset = offset; p.seed_dev = seed_dev;
  p.out = reinterpret_cast<__nv_bfloat16*>(out); p.ld_out = ld_out; p.c_off_out = c_off_out;
  p.res = reinterpret_cast<const __nv_bfloat16*>(res); p.ld_res = ld_res;
  {
    const int rows_per_block = kEwThreads / (C / 8);
    int64_t blocks = (M + rows_per_block * 4 - 1) / (rows_per_block * 4);
    if (blocks > 148 * 8) blocks = 148 * 8;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool drop = drop_p > 0.f;
    if (res && drop) bn_apply_kernel<true, true><<<(int)blocks, kEwThreads, 0, st>>>(p);
    else if (res) bn_apply_kernel<true, false><<<(int)blocks, kEwThreads, 0, st>>>(p);
    else if (drop) bn_apply_kernel<false, true><<<(int)blocks, kEwThreads, 0, st>>>(p);
    else bn_apply_kernel<false, false><<<(int)blocks, kEwThreads, 0, st>>>(p);
  }
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int pp_bn_apply_stats(const void* raw, int64_t M, int ld_in, int c_off_in, int C, const float* sums, const float* gamma,
                      const float* beta, float eps, float momentum, float* running_mean, float* running_var, float* stats_out,
                      int relu, const void* res, int ld_res, void* out, int ld_out, int c_off_out, void* stream) {
  PP_CHECK_ARG(raw && out && sums && gamma && beta && stats_out && M > 0, "pp_bn_apply_stats: bad args");
  PP_CHECK_ARG(!res || (ld_res % 8 == 0 && ld_res >= C), "pp_bn_apply_stats: residual ld=%d", ld_res);
  PP_CHECK_ARG(C % 8 == 0 && C <= 2048 && ld_in % 8 == 0 && c_off_in % 8 == 0 && ld_out % 8 == 0 && c_off_out % 8 == 0,
               "pp_bn_apply_stats: channel counts/offsets must be multiples of 8 (C <= 2048)");
  PP_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "pp_bn_apply_stats: running_mean / running_var go together");
  BnApplyParams p;
  p.raw = reinterpret_cast<const __nv_bfloat16*>(raw);
  p.M = M; p.ld_in = ld_in; p.c_off_in = c_off_in; p.C = C; p.scale = nullptr; p.shift = nullptr; p.relu = relu;
  p.drop_p = 0.f; p.seed = 0; p.offset = 0; p.seed_dev = nullptr;
  p.out = reinterpret_cast<__nv_bfloat16*>(out); p.ld_out = ld_out; p.c_off_out = c_off_out;
  p.res = reinterpret_cast<const __nv_bfloat16*>(res); p.ld_res = ld_res;
  BnFinalizeArgs a;
  a.sums = sums; a.gamma = gamma; a.beta = beta;
  a.inv_m = 1.f / (float)M;
  a.unbias = M > 1 ? (float)M / (float)(M - 1) : 1.f;
  a.eps = eps; a.momentum = momentum; a.running_mean = running_mean; a.running_var = running_var; a.stats_out = stats_out;
  const int rows_per_block = kEwThreads / (C / 8);
  int64_t blocks = (M + rows_per_block * 4 - 1) / (rows_per_block * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (res) bn_apply_stats_kernel<true><<<(int)blocks, kEwThreads, 0, st>>>(p, a);
  else bn_apply_stats_kernel<false><<<(int)blocks, kEwThreads, 0, st>>>(p, a);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int pp_bn_bwd(const void* dy, int ld_dy, int c_off_dy, const void* raw, int ld_raw, int c_off_raw, int64_t M, int C,
              const float* scale, const float* shift, const float* mean, const float* rstd, int relu, float drop_p,
              uint64_t seed, uint64_t offset, const uint64_t* seed_dev, float* sums, void* draw, void* stream) {
  return pp_bn_bwd_res(dy, ld_dy, c_off_dy, raw, ld_raw, c_off_raw, M, C, scale, shift, mean, rstd, relu, drop_p, seed,
                       offset, seed_dev, nullptr, 0, nullptr, sums, draw, C, 0, stream);
}

int pp_bn_bwd_res(const void* dy, int ld_dy, int c_off_dy, const void* raw, int ld_raw, int c_off_raw, int64_t M, int C,
                  const float* scale, const float* shift, const float* mean, const float* rstd, int relu, float drop_p,
                  uint64_t seed, uint64_t offset, const uint64_t* seed_dev, const void* res, int ld_res, void* dres,
                  float* sums, void* draw, int ld_draw, int c_off_draw, void* stream) {
  PP_CHECK_ARG(dy && raw && sums && draw && M > 0, "pp_bn_bwd: bad args");
  PP_CHECK_ARG(ld_draw % 8 == 0 && c_off_draw % 8 == 0 && c_off_draw + C <= ld_draw, "pp_bn_bwd: draw slice ld=%d c_off=%d",
               ld_draw, c_off_draw);
  PP_CHECK_ARG(!res || (ld_res % 8 == 0 && ld_res >= C), "pp_bn_bwd: residual ld=%d", ld_res);
  PP_CHECK_ARG(C % 8 == 0 && C <= 2048 && ld_dy % 8 == 0 && c_off_dy % 8 == 0 && ld_raw % 8 == 0 && c_off_raw % 8 == 0,
               "pp_bn_bwd: C=%d must be a multiple of 8 and <= 2048", C);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  BnBwdParams p;
  p.dy = reinterpret_cast<const __nv_bfloat16*>(dy); p.ld_dy = ld_dy; p.c_off_dy = c_off_dy;
  p.raw = reinterpret_cast<const __nv_bfloat16*>(raw); p.ld_raw = ld_raw; p.c_off_raw = c_off_raw;
  p.M = M; p.C = C; p.scale = scale; p.shift = shift; p.mean = mean; p.rstd = rstd; p.relu = relu;
  p.drop_p = drop_p; p.seed = seed; p.offset = offset; p.seed_dev = seed_dev;
  p.sums = sums;
  p.draw = reinterpret_cast<__nv_bfloat16*>(draw); p.ld_draw = ld_draw; p.c_off_draw = c_off_draw;
  p.res = reinterpret_cast<const __nv_bfloat16*>(res); p.ld_res = ld_res;
  p.dres = reinterpret_cast<__nv_bfloat16*>(dres);
  PP_CUDA(cudaMemsetAsync(sums, 0, 2 * (size_t)C * sizeof(float), st));
  const int rows_per_block = kEwThreads / (C / 8);
  int64_t blocks = (M + rows_per_block * 4 - 1) / (rows_per_block * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  const size_t sm = kEwThreads * 16 * sizeof(float);
  const bool drop = drop_p > 0.f;
  const int64_t rblocks = blocks < 148 * 3 ? blocks : 148 * 3;  // reduce pass: one resident wave (3 CTAs / SM)
#define PP_BN_BWD(R, D)                                                      \
  do {                                                                       \
    bn_bwd_reduce_kernel<R, D><<<(int)rblocks, kEwThreads, sm, st>>>(p);     \
    PP_LAUNCH_CHECK();                                                       \
    bn_bwd_apply_kernel<R, D><<<(int)blocks, kEwThreads, 0, st>>>(p);        \
    PP_LAUNCH_CHECK();                                                       \
  } while (0)
  if (res && drop) PP_BN_BWD(true, true);
  else if (res) PP_BN_BWD(true, false);
  else if (drop) PP_BN_BWD(false, true);
  else PP_BN_BWD(false, false);
#undef PP_BN_BWD
  return PP_OK;
}

int pp_bn_scratch_bytes(int C, size_t* out_bytes) {  // (the cluster path for small tensors does not touch the scratch)
  PP_CHECK_ARG(out_bytes && C > 0, "pp_bn_scratch_bytes: bad args");
  *out_bytes = ((size_t)2 * C + 2) * sizeof(float);
  return PP_OK;
}

int pp_bn_fwd_fused(const void* raw, int64_t M, int ld_in, int c_off_in, int C, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, long long* num_batches_tracked, float eps, float momentum,
                    int relu, float drop_p, uint64_t seed, uint64_t offset, const uint64_t* seed_dev, const void* res,
                    int ld_res, void* out, int ld_out, int c_off_out, float* stats_out, void* scratch, void* stream) {
  PP_CHECK_ARG(raw && out && gamma && beta && stats_out && scratch && M > 0, "pp_bn_fwd_fused: bad args");
  PP_CHECK_ARG(C % 8 == 0 && C >= 8 && C <= 2048 && ld_in % 8 == 0 && c_off_in % 8 == 0 && ld_out % 8 == 0 && c_off_out % 8 == 0,
               "pp_bn_fwd_fused: channel counts/offsets must be multiples of 8 (C <= 2048)");
  PP_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "pp_bn_fwd_fused: drop_p=%f", drop_p);
  PP_CHECK_ARG(!res || (ld_res % 8 == 0 && ld_res >= C), "pp_bn_fwd_fused: residual ld=%d", ld_res);
  PP_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "pp_bn_fwd_fused: running stats come in pairs");
  BnFwdFusedParams q;
  BnApplyParams& p = q.a;
  p.raw = reinterpret_cast<const __nv_bfloat16*>(raw);
  p.M = M; p.ld_in = ld_in; p.c_off_in = c_off_in; p.C = C; p.scale = nullptr; p.shift = nullptr; p.relu = relu;
  p.drop_p = drop_p; p.seed = seed; p.offset = offset; p.seed_dev = seed_dev;
  p.out = reinterpret_cast<__nv_bfloat16*>(out); p.ld_out = ld_out; p.c_off_out = c_off_out;
  p.res = reinterpret_cast<const __nv_bfloat16*>(res); p.ld_res = ld_res;
  q.gamma = gamma; q.beta = beta; q.running_mean = running_mean; q.running_var = running_var; q.nbt = num_batches_tracked;
  q.eps = eps; q.momentum = momentum; q.stats_out = stats_out;
  q.sums = reinterpret_cast<float*>(scratch);
  q.bar = reinterpret_cast<unsigned*>(q.sums + 2 * (size_t)C);
  const int rows_per_block = kEwThreads / (C / 8);
  const int64_t blocks = (M + rows_per_block * 8 - 1) / (rows_per_block * 8);  // >= 8 rows per thread: fewer partial-sum atomics
  const size_t sm = kEwThreads * 16 * sizeof(float);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool drop = drop_p > 0.f;
  if (const int nc = bn_cluster_ctas(M, C)) {
    const size_t smc = sm + (size_t)4 * C * sizeof(float);
    if (res && drop) return launch_cluster(bn_fwd_cluster_kernel<true, true>, q, nc, smc, st);
    if (res) return launch_cluster(bn_fwd_cluster_kernel<true, false>, q, nc, smc, st);
    if (drop) return launch_cluster(bn_fwd_cluster_kernel<false, true>, q, nc, smc, st);
    return launch_cluster(bn_fwd_cluster_kernel<false, false>, q, nc, smc, st);
  }
  if (res && drop) return launch_coop(bn_fwd_fused_kernel<true, true>, q, blocks, sm, st, "pp_bn_fwd_fused");
  if (res) return launch_coop(bn_fwd_fused_kernel<true, false>, q, blocks, sm, st, "pp_bn_fwd_fused");
  if (drop) return launch_coop(bn_fwd_fused_kernel<false, true>, q, blocks, sm, st, "pp_bn_fwd_fused");
  return launch_coop(bn_fwd_fused_kernel<false, false>, q, blocks, sm, st, "pp_bn_fwd_fused");
}

int pp_bn_bwd_fused(const void* dy, int ld_dy, int c_off_dy, const void* raw, int ld_raw, int c_off_raw, int64_t M, int C,
                    const float* scale, const float* shift, const float* mean, const float* rstd, int relu, float drop_p,
                    uint64_t seed, uint64_t offset, const uint64_t* seed_dev, const void* res, int ld_res, void* dres,
                    float* sums, void* draw, int ld_draw, int c_off_draw, void* scratch, void* stream) {
  PP_CHECK_ARG(dy && raw && sums && draw && scratch && M > 0, "pp_bn_bwd_fused: bad args");
  PP_CHECK_ARG(C % 8 == 0 && C <= 2048 && ld_dy % 8 == 0 && c_off_dy % 8 == 0 && ld_raw % 8 == 0 && c_off_raw % 8 == 0,
               "pp_bn_bwd_fused: C=%d must be a multiple of 8 and <= 2048", C);
  PP_CHECK_ARG(!res || (ld_res % 8 == 0 && ld_res >= C), "pp_bn_bwd_fused: residual ld=%d", ld_res);
  PP_CHECK_ARG(ld_draw % 8 == 0 && c_off_draw % 8 == 0 && c_off_draw + C <= ld_draw, "pp_bn_bwd_fused: draw slice ld=%d c_off=%d",
               ld_draw, c_off_draw);
  BnBwdFusedParams q;
  BnBwdParams& p = q.b;
  p.dy = reinterpret_cast<const __nv_bfloat16*>(dy); p.ld_dy = ld_dy; p.c_off_dy = c_off_dy;
  p.raw = reinterpret_cast<const __nv_bfloat16*>(raw); p.ld_raw = ld_raw; p.c_off_raw = c_off_raw;
  p.M = M; p.C = C; p.scale = scale; p.shift = shift; p.mean = mean; p.rstd = rstd; p.relu = relu;
  p.drop_p = drop_p; p.seed = seed; p.offset = offset; p.seed_dev = seed_dev;
  p.sums = sums;
  p.draw = reinterpret_cast<__nv_bfloat16*>(draw); p.ld_draw = ld_draw; p.c_off_draw = c_off_draw;
  p.res = reinterpret_cast<const __nv_bfloat16*>(res); p.ld_res = ld_res;
  p.dres = reinterpret_cast<__nv_bfloat16*>(dres);
  q.sums = reinterpret_cast<float*>(scratch);
  q.bar = reinterpret_cast<unsigned*>(q.sums + 2 * (size_t)C);
  const int rows_per_block = kEwThreads / (C / 8);
  const int64_t blocks = (M + rows_per_block * 8 - 1) / (rows_per_block * 8);  // >= 8 rows per thread: fewer partial-sum atomics
  const size_t sm = kEwThreads * 16 * sizeof(float);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool drop = drop_p > 0.f;
  if (const int nc = bn_cluster_ctas(M, C)) {
    const size_t smc = sm + (size_t)4 * C * sizeof(float);
    if (res && drop) return launch_cluster(bn_bwd_cluster_kernel<true, true>, q, nc, smc, st);
    if (res) return launch_cluster(bn_bwd_cluster_kernel<true, false>, q, nc, smc, st);
    if (drop) return launch_cluster(bn_bwd_cluster_kernel<false, true>, q, nc, smc, st);
    return launch_cluster(bn_bwd_cluster_kernel<false, false>, q, nc, smc, st);
  }
  if (res && drop) return launch_coop(bn_bwd_fused_kernel<true, true>, q, blocks, sm, st, "pp_bn_bwd_fused");
  if (res) return launch_coop(bn_bwd_fused_kernel<true, false>, q, blocks, sm, st, "pp_bn_bwd_fused");
  if (drop) return launch_coop(bn_bwd_fused_kernel<false, true>, q, blocks, sm, st, "pp_bn_bwd_fused");
  return launch_coop(bn_bwd_fused_kernel<false, false>, q, blocks, sm, st, "pp_bn_bwd_fused");
}

int pp_upsample_nhwc_bf16(const void* in, int N, int h, int w, int C, int ld_in, void* out, int H, int W, int ld_out,
                          int c_off, void* stream) {
  PP_CHECK_ARG(in && out && N > 0 && h > 0 && w > 0 && H > 0 && W > 0, "pp_upsample_nhwc_bf16: bad args");
  PP_CHECK_ARG(C % 8 == 0 && ld_in % 8 == 0 && ld_out % 8 == 0 && c_off % 8 == 0, "pp_upsample_nhwc_bf16: multiples of 8");
  upsample_nhwc_kernel<<<ew_grid((int64_t)N * H * W * (C / 8)), kEwThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(in), N, h, w, C, ld_in, reinterpret_cast<__nv_bfloat16*>(out), H, W, ld_out,
      c_off, ac_scale(h, H), ac_scale(w, W));
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int pp_upsample_nhwc_bf16_bwd(const void* grad_out, int N, int H, int W, int ld, int c_off, int C, float* grad_in, int h,
                              int w, void* stream) {
  PP_CHECK_ARG(grad_out && grad_in && N > 0 && h > 0 && w > 0 && H > 0 && W > 0, "pp_upsample_nhwc_bf16_bwd: bad args");
  PP_CHECK_ARG(C % 8 == 0 && ld % 8 == 0 && c_off % 8 == 0, "pp_upsample_nhwc_bf16_bwd: multiples of 8");
  upsample_nhwc_bwd_kernel<<<ew_grid((int64_t)N * h * w * (C / 8)), kEwThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(grad_out), N, H, W, ld, c_off, C, grad_in, h, w, ac_scale(h, H), ac_scale(w, W));
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int pp_to_nhwc_bf16(const void* in, int dtype, int64_t sn, int64_t sc, int64_t sh, int64_t sw, int N, int C, int H, int W,
                    void* out, int ld, int c_off, void* stream) {
  PP_CHECK_ARG(in && out && N > 0 && C > 0 && H > 0 && W > 0 && ld >= c_off + C, "pp_to_nhwc_bf16: bad args");
  PP_CHECK_ARG(N <= 65535, "pp_to_nhwc_bf16: N too large");
  dim3 grid((unsigned)(((int64_t)H * W + 31) / 32), (C + 31) / 32, N);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == PP_F32)
    to_nhwc_bf16_kernel<float><<<grid, kEwThreads, 0, st>>>(reinterpret_cast<const float*>(in), sn, sc, sh, sw, N, C, H, W,
                                                             reinterpret_cast<__nv_bfloat16*>(out), ld, c_off);
  else if (dtype == PP_BF16)
    to_nhwc_bf16_kernel<__nv_bfloat16><<<grid, kEwThreads, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(in), sn, sc, sh, sw,
                                                                     N, C, H, W, reinterpret_cast<__nv_bfloat16*>(out), ld, c_off);
  else {
    set_error("pp_to_nhwc_bf16: bad dtype %d", dtype);
    return PP_ERR_INVALID_ARG;
  }
  PP_LAUNCH_CHECK();
  return PP_OK;
}

}  // extern "C"

// ---- BN finalize: statistics -> (scale, shift, mean, rstd) + running-stat update ------------------
// nn.BatchNorm2d train-mode semantics (biased variance to normalise, unbiased for running_var,
// momentum update), one thread per channel.
namespace pp {
__global__ void bn_finalize_kernel(const float* __restrict__ sums, int C, float inv_m, float unbias, float eps,
                                   float momentum, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ out /*[4][Cpad]: scale, shift, mean, rstd*/, int Cpad) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cpad) return;
  float scale = 0.f, shift = 0.f, mean = 0.f, rstd = 0.f;
  if (c < C) {
    mean = sums[c] * inv_m;
    float var = fmaf(-mean, mean, sums[C + c] * inv_m);
    var = fmaxf(var, 0.f);
    rstd = rsqrtf(var + eps);
    scale = gamma[c] * rstd;
    shift = fmaf(-mean, scale, beta[c]);
    if (running_mean) {
      running_mean[c] = fmaf(momentum, mean - running_mean[c], running_mean[c]);
      running_var[c] = fmaf(momentum, var * unbias - running_var[c], running_var[c]);
    }
  }
  out[c] = scale;
  out[Cpad + c] = shift;
  out[2 * Cpad + c] = mean;
  out[3 * Cpad + c] = rstd;
}
}  // namespace pp

extern "C" int pp_bn_finalize(const float* sums, int C, int64_t M, float eps, float momentum, const float* gamma,
                              const float* beta, float* running_mean, float* running_var, float* out, int Cpad,
                              void* stream) {
  PP_CHECK_ARG(sums && gamma && beta && out && C > 0 && Cpad >= C && M > 0, "pp_bn_finalize: bad args");
  const float unbias = M > 1 ? (float)M / (float)(M - 1) : 1.f;
  pp::bn_finalize_kernel<<<(Cpad + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      sums, C, 1.f / (float)M, unbias, eps, momentum, gamma, beta, running_mean, running_var, out, Cpad);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

// ---- conv weight packing: f32 [Cout][Cin_total][kh][kw] -> bf16 K-major operand tensors ------------
// fwd   [taps][Cout_pad][Cin_pad]          (tap = ky*kw + kx)
// dgrad [taps][Cin_rows][Cout_cols]        (tap flipped: the data-gradient conv), either may be NULL
namespace pp {
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int Cin_total,
                                                          int taps, __nv_bfloat16* __restrict__ fwd, int Cout_pad,
                                                          int Cin_pad, __nv_bfloat16* __restrict__ dgr, int Cin_rows,
                                                          int Cout_cols) {
  const int64_t n_f = fwd ? (int64_t)taps * Cout_pad * Cin_pad : 0;
  const int64_t n_d = dgr ? (int64_t)taps * Cin_rows * Cout_cols : 0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_f + n_d; i += (int64_t)gridDim.x * 256) {
    if (i < n_f) {
      const int ci = (int)(i % Cin_pad);
      const int64_t t = i / Cin_pad;
      const int co = (int)(t % Cout_pad), tap = (int)(t / Cout_pad);
      float v = 0.f;
      if (ci < Cin && co < Cout) v = w[((int64_t)co * Cin_total + ci) * taps + tap];
      fwd[i] = __float2bfloat16(v);
    } else {
      const int64_t j = i - n_f;
      const int co = (int)(j % Cout_cols);
      const int64_t t = j / Cout_cols;
      const int ci = (int)(t % Cin_rows), tap = (int)(t / Cin_rows);
      float v = 0.f;
      if (ci < Cin && co < Cout) v = w[((int64_t)co * Cin_total + ci) * taps + (taps - 1 - tap)];
      dgr[j] = __float2bfloat16(v);
    }
  }
}
}  // namespace pp

namespace pp {
// Tiled form (taps = 1 or 9, the only kernel sizes packed in this network): a CTA owns a T x T block of (co, ci) with all
// its taps - T = 64 for 1x1, 32 for 3x3 - reads it as T runs of T * taps CONTIGUOUS floats, keeps it in shared memory as
// bf16 and writes both operand images in contiguous runs of T elements (forward along ci, data gradient along co).  The
// gather form above reads the fp32 master with a stride of taps (forward) or Cin * taps (data gradient) floats per thread
// and ran at ~1 TB/s; this one touches every byte once.
struct PackDesc {
  const float* w;
  __nv_bfloat16* fwd;
  __nv_bfloat16* dgr;
  int Cout, Cin, Cin_total, taps, Cout_pad, Cin_pad, Cin_rows, Cout_cols;
};

__host__ __device__ inline int pack_tile_edge(int taps) { return taps == 1 ? 64 : 32; }
__host__ __device__ inline int pack_co_extent(const PackDesc& d) {
  const int a = d.fwd ? d.Cout_pad : 0, b = d.dgr ? d.Cout_cols : 0;
  return a > b ? a : b;
}
__host__ __device__ inline int pack_ci_extent(const PackDesc& d) {
  const int a = d.fwd ? d.Cin_pad : 0, b = d.dgr ? d.Cin_rows : 0;
  return a > b ? a : b;
}
__host__ __device__ inline int pack_num_tiles(const PackDesc& d) {
  const int T = pack_tile_edge(d.taps);
  return ((pack_co_extent(d) + T - 1) / T) * ((pack_ci_extent(d) + T - 1) / T);
}
static bool pack_tiled_ok(const PackDesc& d) {
  return (d.taps == 1 || d.taps == 9) && (!d.fwd || (d.Cin_pad % 8 == 0 && (reinterpret_cast<uintptr_t>(d.fwd) & 15u) == 0)) &&
         (!d.dgr || (d.Cout_cols % 8 == 0 && (reinterpret_cast<uintptr_t>(d.dgr) & 15u) == 0));
}

constexpr int kPackSmemElems = 9 * (32 * 34 + 2);  // >= 64 * 66

template <int TAPS, int T>
__device__ __forceinline__ void pack_tile(const PackDesc& d, int tile, __nv_bfloat16* sm) {
  constexpr int P = T + 2;  // row pitch: even (4-byte pair reads along ci), odd in 4-byte words (conflict-free reads along co)
  constexpr int TS = T * P + 2;  // tap pitch: one word past a multiple of 32 words, so the 9 taps of a pixel hit 9 banks
  constexpr int RUN = T * TAPS;  // contiguous floats of one output channel inside the tile
  const int n_ci_t = (pack_ci_extent(d) + T - 1) / T;
  const int co0 = (tile / n_ci_t) * T, ci0 = (tile % n_ci_t) * T;
  const bool full = co0 + T <= d.Cout && ci0 + T <= d.Cin && (((int64_t)d.Cin_total * TAPS) & 3) == 0 &&
                    (reinterpret_cast<uintptr_t>(d.w) & 15u) == 0;  // RUN * ci0 / T is a multiple of 4 floats by construction
  if (full) {
    for (int e4 = threadIdx.x; e4 < T * RUN / 4; e4 += 256) {
      const int co_l = e4 / (RUN / 4), r4 = (e4 - co_l * (RUN / 4)) * 4;
      const float4 q = __ldg(reinterpret_cast<const float4*>(d.w + ((int64_t)(co0 + co_l) * d.Cin_total + ci0) * TAPS + r4));
      const float qv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r4 + u, ci_l = r / TAPS, tap = r - ci_l * TAPS;
        sm[tap * TS + co_l * P + ci_l] = __float2bfloat16(qv[u]);
      }
    }
  } else {
    for (int e = threadIdx.x; e < T * RUN; e += 256) {
      const int co_l = e / RUN, r = e - co_l * RUN;
      const int ci_l = r / TAPS, tap = r - ci_l * TAPS;
      const int co = co0 + co_l, ci = ci0 + ci_l;
      float v = 0.f;
      if (co < d.Cout && ci < d.Cin) v = __ldg(d.w + ((int64_t)co * d.Cin_total + ci) * TAPS + tap);
      sm[tap * TS + co_l * P + ci_l] = __float2bfloat16(v);
    }
  }
  __syncthreads();
  // stores: 8 elements (16 bytes) per thread
  if (d.fwd) {
    for (int e = threadIdx.x; e < TAPS * T * (T / 8); e += 256) {
      const int ci_l = (e % (T / 8)) * 8, t2 = e / (T / 8);
      const int co_l = t2 % T, tap = t2 / T;
      const int co = co0 + co_l, ci = ci0 + ci_l;
      if (co < d.Cout_pad && ci < d.Cin_pad) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(sm + tap * TS + co_l * P + ci_l);
        *reinterpret_cast<uint4*>(d.fwd + ((int64_t)tap * d.Cout_pad + co) * d.Cin_pad + ci) = make_uint4(src[0], src[1], src[2], src[3]);
      }
    }
  }
  if (d.dgr) {
    for (int e = threadIdx.x; e < TAPS * T * (T / 8); e += 256) {
      const int co_l = (e % (T / 8)) * 8, t2 = e / (T / 8);
      const int ci_l = t2 % T, tap = t2 / T;
      const int co = co0 + co_l, ci = ci0 + ci_l;
      if (ci < d.Cin_rows && co < d.Cout_cols) {
        const uint16_t* src = reinterpret_cast<const uint16_t*>(sm + tap * TS + co_l * P + ci_l);
        uint32_t o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) o[u] = (uint32_t)src[(2 * u) * P] | ((uint32_t)src[(2 * u + 1) * P] << 16);
        *reinterpret_cast<uint4*>(d.dgr + ((int64_t)(TAPS - 1 - tap) * d.Cin_rows + ci) * d.Cout_cols + co) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

__global__ void __launch_bounds__(256) pack_weight_tiled_kernel(const PackDesc d) {
  __shared__ __align__(16) __nv_bfloat16 sm[kPackSmemElems];
  if (d.taps == 1) pack_tile<1, 64>(d, blockIdx.x, sm);
  else pack_tile<9, 32>(d, blockIdx.x, sm);
}

// every conv of a network in ONE launch: table row = (w, fwd, dgrad pointers, the 8 shapes of PackDesc, first tile of the row);
// blockIdx.x = tile, its row found by bisection
__global__ void __launch_bounds__(256) pack_weight_batched_kernel(const long long* __restrict__ table, int n) {
  __shared__ __align__(16) __nv_bfloat16 sm[kPackSmemElems];
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if ((int)table[(size_t)mid * 12 + 11] <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const long long* r = table + (size_t)lo * 12;
  PackDesc d;
  d.w = reinterpret_cast<const float*>(r[0]);
  d.fwd = reinterpret_cast<__nv_bfloat16*>(r[1]);
  d.dgr = reinterpret_cast<__nv_bfloat16*>(r[2]);
  d.Cout = (int)r[3]; d.Cin = (int)r[4]; d.Cin_total = (int)r[5]; d.taps = (int)r[6];
  d.Cout_pad = (int)r[7]; d.Cin_pad = (int)r[8]; d.Cin_rows = (int)r[9]; d.Cout_cols = (int)r[10];
  const int tile = (int)blockIdx.x - (int)r[11];
  if (tile >= pack_num_tiles(d)) return;
  if (d.taps == 1) pack_tile<1, 64>(d, tile, sm);
  else pack_tile<9, 32>(d, tile, sm);
}
}  // namespace pp

extern "C" int pp_pack_conv_weights_batched(const long long* table_dev, int n, int total_tiles, void* stream) {
  PP_CHECK_ARG(table_dev && n > 0 && n <= 65535 && total_tiles > 0, "pp_pack_conv_weights_batched: bad args");
  pp::pack_weight_batched_kernel<<<(unsigned)total_tiles, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(table_dev, n);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

extern "C" int pp_pack_conv_weights_tiles(int Cout_pad, int Cin_pad, int Cin_rows, int Cout_cols, int taps) {
  pp::PackDesc d{};
  d.fwd = Cout_pad > 0 ? reinterpret_cast<__nv_bfloat16*>(16) : nullptr;
  d.dgr = Cin_rows > 0 ? reinterpret_cast<__nv_bfloat16*>(16) : nullptr;
  d.taps = taps; d.Cout_pad = Cout_pad; d.Cin_pad = Cin_pad; d.Cin_rows = Cin_rows; d.Cout_cols = Cout_cols;
  if (!(taps == 1 || taps == 9) || (d.fwd && Cin_pad % 8) || (d.dgr && Cout_cols % 8) || (!d.fwd && !d.dgr)) {
    pp::set_error("pp_pack_conv_weights_tiles: taps must be 1 or 9 and the inner extents multiples of 8");
    return -1;
  }
  return pp::pack_num_tiles(d);
}

extern "C" int pp_pack_conv_weight(const float* w, int Cout, int Cin, int Cin_total, int taps, void* fwd, int Cout_pad,
                                   int Cin_pad, void* dgrad, int Cin_rows, int Cout_cols, void* stream) {
  PP_CHECK_ARG(w && (fwd || dgrad) && Cout > 0 && Cin > 0 && Cin <= Cin_total && taps > 0, "pp_pack_conv_weight: bad args");
  PP_CHECK_ARG(!fwd || (Cout_pad >= Cout && Cin_pad >= Cin), "pp_pack_conv_weight: fwd padding smaller than the tensor");
  PP_CHECK_ARG(!dgrad || (Cin_rows >= Cin && Cout_cols >= Cout), "pp_pack_conv_weight: dgrad padding smaller than the tensor");
  {
    pp::PackDesc d{w, reinterpret_cast<__nv_bfloat16*>(fwd), reinterpret_cast<__nv_bfloat16*>(dgrad), Cout, Cin, Cin_total, taps,
                   Cout_pad, Cin_pad, Cin_rows, Cout_cols};
    if (pp::pack_tiled_ok(d)) {
      pp::pack_weight_tiled_kernel<<<(unsigned)pp::pack_num_tiles(d), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d);
      PP_LAUNCH_CHECK();
      return PP_OK;
    }
  }
  const int64_t total = (fwd ? (int64_t)taps * Cout_pad * Cin_pad : 0) + (dgrad ? (int64_t)taps * Cin_rows * Cout_cols : 0);
  pp::pack_weight_kernel<<<pp::ew_grid(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      w, Cout, Cin, Cin_total, taps, reinterpret_cast<__nv_bfloat16*>(fwd), Cout_pad, Cin_pad,
      reinterpret_cast<__nv_bfloat16*>(dgrad), Cin_rows, Cout_cols);
  PP_LAUNCH_CHECK();
  return PP_OK;
}
