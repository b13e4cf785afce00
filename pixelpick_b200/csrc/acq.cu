// Q path — fused acquisition scoring and per-image sorted top-k.
//
//   K1  acq_score_*      logits[n,C,H,W] (+masks) -> score[n,H*W] (+ level-0 radix histogram)
//                        one coalesced, vectorised pass over the logits: HBM-bound, C*4+2 B/px.  Large fp32 batches run
//                        acq_score_pf_kernel: a thread's next tile is in flight (cp.async into a private shared-memory slot)
//                        while it scores the current one.
//   K2  pick_bucket0     per image: the level-0 bucket holding the k-th score, and that bucket's ordering-key range.
//       select_l0        ONE pass over the score map: composites (ord_key(score) << 32 | flat idx) below the bucket go
//                        to the candidate list, those inside it to the (small) boundary list.  Level 0 is bucket0()
//                        (pp_common.cuh): the key's leading 11 bits for smallest-first selections, a linear
//                        quantisation of [0, 4) for largest-first ones.  select_l0_staged_kernel: persistent, cp.async
//                        ring per warp, survivors compacted before they are classified.
//       select_rest      MSD radix levels on the boundary list, one CTA per image, until the k-th is isolated (short lists
//                        are ranked directly).
//   K3a pick_ranks_fast  the reference only reads n random RANKS of the sorted list (query.py:63-64): three
//                        table-lookup passes over the k unsorted candidates return exactly those order statistics;
//                        images it cannot finish (> 32 exact ties, n > 12) run the generic 5-level walk in the same launch.
//   K3b bitonic_*        full sort of the k composites when the list itself is wanted (ties -> lower flat index first).
//
// The workspace's histograms / counters are zero before a step (pp_acq_topk_prepare) and zero again after its select:
// select_rest zeroes what the step consumed.
//
// Reference semantics restated (query.py:33-69,190-201,224-247): see include/pixelpick_b200.h.
#include <cooperative_groups.h>
#include "pp_common.cuh"
#include <stdlib.h>

namespace pp {

// ------------------------------------------------------------------------------------------
// loads
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ldg_stream_bf4(const __nv_bfloat16* p) {
  uint32_t a, b;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p));
  float4 r;
  r.x = __uint_as_float(a << 16);
  r.y = __uint_as_float(a & 0xFFFF0000u);
  r.z = __uint_as_float(b << 16);
  r.w = __uint_as_float(b & 0xFFFF0000u);
  return r;
}
template <typename T> struct Ld4;
template <> struct Ld4<float> {
  static __device__ __forceinline__ float4 ld(const float* p) { return ldg_stream_f4(p); }
  static __device__ __forceinline__ float ld1(const float* p) { return __ldg(p); }
};
template <> struct Ld4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 ld(const __nv_bfloat16* p) { return ldg_stream_bf4(p); }
  static __device__ __forceinline__ float ld1(const __nv_bfloat16* p) {
    return __bfloat162float(*p);
  }
};

// ------------------------------------------------------------------------------------------
// per-pixel score from C logits held in registers
// ------------------------------------------------------------------------------------------
// exp(x - m) for the softmax denominator: one FFMA + MUFU.EX2 (ex2.approx, rel. error 2^-22) instead of
// the ~8-instruction expf.  The kernel is co-limited by instruction issue and HBM, so this matters;
// the induced score error (~1e-7 abs) is far inside the 2e-6 / 1e-5 parity tolerance.
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
constexpr float kLog2e = 1.4426950408889634f;

template <int C, int STRAT>
__device__ __forceinline__ float score_from_logits(const float (&x)[C]) {
  const float qnan = __uint_as_float(0x7FC00000u);
  if (STRAT == PP_STRAT_ENTROPY) {
    // H = sum_c -p_c log p_c with log p_c = (x_c - m) - log S  =>  H = log S - (sum_c e_c d_c) / S.
    // The reference evaluates -p*log(p) literally, so a class whose probability underflows to 0
    // gives 0 * -inf = NaN (query.py:230); reproduced through the smallest class probability.
    float m = x[0];
#pragma unroll
    for (int c = 1; c < C; ++c) m = fmaxf(m, x[c]);
    float S = 0.f, T = 0.f, dmin = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float d = x[c] - m;
      const float e = ex2_approx(d * kLog2e);
      S += e;
      T = fmaf(e, d, T);
      dmin = fminf(dmin, d);
    }
    float h = logf(S) - T / S;
    if (dmin < -87.0f) {  // rare: only then can a probability underflow to exactly 0 (precise expf)
      if (expf(dmin) / S == 0.f) h = qnan;
    }
    return h;  // NaN / inf logits propagate through S and T as in torch
  } else if (STRAT == PP_STRAT_LEAST_CONFIDENCE) {
    float m = x[0];
#pragma unroll
    for (int c = 1; c < C; ++c) m = fmaxf(m, x[c]);
    const float mb = -m * kLog2e;
    float S = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) S += ex2_approx(fmaf(x[c], kLog2e, mb));
    return 1.0f - 1.0f / S;  // max_c p_c = exp(0) / S
  } else {
    float m1 = x[0], m2 = -INFINITY;
#pragma unroll
    for (int c = 1; c < C; ++c) {
      m2 = fmaxf(m2, fminf(m1, x[c]));
      m1 = fmaxf(m1, x[c]);
    }
    const float mb = -m1 * kLog2e;
    float S = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) S += ex2_approx(fmaf(x[c], kLog2e, mb));
    const float p1 = 1.0f / S;
    const float p2 = expf(m2 - m1) / S;
    return fabsf(p1 - p2);
  }
}

// runtime-C version (scalar fallback): re-reads the logits through L1/L2
template <int STRAT, typename F>
__device__ __forceinline__ float score_runtime_c(int C, F&& get) {
  const float qnan = __uint_as_float(0x7FC00000u);
  float m1 = get(0), m2 = -INFINITY;
  for (int c = 1; c < C; ++c) {
    const float v = get(c);
    m2 = fmaxf(m2, fminf(m1, v));
    m1 = fmaxf(m1, v);
  }
  float S = 0.f, T = 0.f, dmin = 0.f;
  for (int c = 0; c < C; ++c) {
    const float d = get(c) - m1;
    const float e = expf(d);
    S += e;
    T = fmaf(e, d, T);
    dmin = fminf(dmin, d);
  }
  if (STRAT == PP_STRAT_ENTROPY) {
    float h = logf(S) - T / S;
    if (expf(dmin) / S == 0.f) h = qnan;
    return h;
  } else if (STRAT == PP_STRAT_LEAST_CONFIDENCE) {
    return 1.0f - 1.0f / S;
  } else {
    return fabsf(1.0f / S - expf(m2 - m1) / S);
  }
}

constexpr int kHistBins = 2048;
constexpr int kScoreThreads = 256;

struct ScoreParams {
  const void* logits;
  int64_t sn, sc, sh;
  int n_img, C, H, W;
  const uint8_t* lab;
  const uint8_t* vd;
  const uint8_t* keep;
  float* score;
  uint32_t* hist0;
  float fill;
  int largest;
};

// Vector kernel: one thread = PX (4 or 2) consecutive pixels of one row; class planes are read as
// 16 B / 8 B streaming loads, C of them in flight per thread.  PX and MINB (min resident CTAs per SM,
// i.e. the register cap) trade per-thread memory-level parallelism against occupancy.
template <typename T, int PX> struct LdPx;
template <> struct LdPx<float, 4> {
  static __device__ __forceinline__ void ld(const float* p, float (&o)[4]) {
    const float4 r = ldg_stream_f4(p);
    o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
  }
};
template <> struct LdPx<float, 2> {
  static __device__ __forceinline__ void ld(const float* p, float (&o)[2]) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(o[0]), "=f"(o[1]) : "l"(p));
  }
};
template <> struct LdPx<__nv_bfloat16, 4> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&o)[4]) {
    const float4 r = ldg_stream_bf4(p);
    o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
  }
};
template <> struct LdPx<__nv_bfloat16, 2> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&o)[2]) {
    uint32_t a;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(a) : "l"(p));
    o[0] = __uint_as_float(a << 16);
    o[1] = __uint_as_float(a & 0xFFFF0000u);
  }
};
template <int PX> __device__ __forceinline__ uint32_t ld_mask(const uint8_t* p);
template <> __device__ __forceinline__ uint32_t ld_mask<4>(const uint8_t* p) {
  return __ldg(reinterpret_cast<const uint32_t*>(p));
}
template <> __device__ __forceinline__ uint32_t ld_mask<2>(const uint8_t* p) {
  return __ldg(reinterpret_cast<const uint16_t*>(p));
}

template <int C, int STRAT, typename T, bool HIST, int ITERS, int PX, int MINB>
__global__ void __launch_bounds__(kScoreThreads, MINB) acq_score_vec_kernel(const ScoreParams p) {
  __shared__ uint32_t sh_hist[HIST ? kHistBins : 1];
  if (HIST) {
    for (int i = threadIdx.x; i < kHistBins; i += kScoreThreads) sh_hist[i] = 0;
    __syncthreads();
  }
  const int img = blockIdx.y;
  const int Wv = p.W / PX;
  const int nvec = p.H * Wv;
  const int64_t HW = (int64_t)p.H * p.W;
  const T* __restrict__ base = reinterpret_cast<const T*>(p.logits) + (int64_t)img * p.sn;
  const bool largest = p.largest != 0;

#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    const int q = (blockIdx.x * ITERS + it) * kScoreThreads + threadIdx.x;
    if (q >= nvec) break;
    const int y = q / Wv;
    const int x = (q - y * Wv) * PX;
    const T* __restrict__ src = base + (int64_t)y * p.sh + x;
    float v[C][PX];
#pragma unroll
    for (int c = 0; c < C; ++c) LdPx<T, PX>::ld(src + (int64_t)c * p.sc, v[c]);
    const int64_t pix = (int64_t)img * HW + (int64_t)y * p.W + x;
    uint32_t msk = 0;
    if (p.lab) msk |= ld_mask<PX>(p.lab + pix);
    if (p.vd) msk |= ld_mask<PX>(p.vd + pix);
    if (p.keep) {
      const uint32_t k4 = ld_mask<PX>(p.keep + pix);
#pragma unroll
      for (int j = 0; j < PX; ++j)
        if (((k4 >> (8 * j)) & 0xFFu) == 0) msk |= 0xFFu << (8 * j);  // keep == 0 -> excluded
    }
    float out[PX];
#pragma unroll
    for (int j = 0; j < PX; ++j) {
      float xs[C];
#pragma unroll
      for (int c = 0; c < C; ++c) xs[c] = v[c][j];
      out[j] = score_from_logits<C, STRAT>(xs);
      if ((msk >> (8 * j)) & 0xFFu) out[j] = p.fill;
    }
    if (PX == 4) *reinterpret_cast<float4*>(p.score + pix) = make_float4(out[0], out[1], out[2], out[PX - 1]);
    else *reinterpret_cast<float2*>(p.score + pix) = make_float2(out[0], out[1]);
    if (HIST) {
#pragma unroll
      for (int j = 0; j < PX; ++j) atomicAdd(&sh_hist[bucket0(out[j], largest)], 1u);
    }
  }
  if (HIST) {
    __syncthreads();
    uint32_t* gh = p.hist0 + (size_t)img * kHistBins;
    for (int i = threadIdx.x; i < kHistBins; i += kScoreThreads) {
      const uint32_t c = sh_hist[i];
      if (c) atomicAdd(gh + i, c);
    }
  }
}

// Vector kernel with the NEXT tile in flight while the current one is scored (fp32 logits, 4 pixels per thread).
// acq_score_vec_kernel issues a thread's C loads, waits, then spends ~600 instructions on the softmax with nothing of its
// own in flight: only the other 1-3 warps of the scheduler keep HBM busy meanwhile.  Here a thread's C x 16 bytes travel
// through a private slot in shared memory (cp.async): the slot is drained into registers, refilled at once with the
// thread's next tile (and the next tile's mask words are loaded into registers), and only then is the drained tile
// scored - so every thread has C x 16 bytes in flight during its whole compute phase.  A thread only reads what it copied
// itself, so cp.async.wait_group is the only synchronisation.  Shared memory: C x 4 KB per CTA (76 KB at C = 19).
template <int C, int STRAT, bool HIST, int ITERS, int MINB>
__global__ void __launch_bounds__(kScoreThreads, MINB) acq_score_pf_kernel(const ScoreParams p) {
  extern __shared__ __align__(16) float4 pf_slot[];  // [C][kScoreThreads]
  __shared__ uint32_t sh_hist[HIST ? kHistBins : 1];
  if (HIST) {
    for (int i = threadIdx.x; i < kHistBins; i += kScoreThreads) sh_hist[i] = 0;
    __syncthreads();
  }
  const int img = blockIdx.y;
  const int Wv = p.W / 4;
  const int nvec = p.H * Wv;
  const int64_t HW = (int64_t)p.H * p.W;
  const float* __restrict__ base = reinterpret_cast<const float*>(p.logits) + (int64_t)img * p.sn;
  const bool largest = p.largest != 0;
  const uint32_t slot_s = (uint32_t)__cvta_generic_to_shared(pf_slot + threadIdx.x);

  bool ok_n = false;
  int64_t pix_n = 0;
  uint32_t lab_n = 0, vd_n = 0, keep_n = 0;
  auto issue = [&](int it) {
    const int q = (blockIdx.x * ITERS + it) * kScoreThreads + threadIdx.x;
    ok_n = it < ITERS && q < nvec;
    if (ok_n) {
      const int y = q / Wv;
      const int x = (q - y * Wv) * 4;
      const float* __restrict__ src = base + (int64_t)y * p.sh + x;
#pragma unroll
      for (int c = 0; c < C; ++c)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot_s + (uint32_t)(c * kScoreThreads * 16)),
                     "l"(src + (int64_t)c * p.sc)
                     : "memory");
      pix_n = (int64_t)img * HW + (int64_t)y * p.W + x;
      if (p.lab) lab_n = ld_mask<4>(p.lab + pix_n);
      if (p.vd) vd_n = ld_mask<4>(p.vd + pix_n);
      if (p.keep) keep_n = ld_mask<4>(p.keep + pix_n);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue(0);
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    if (!ok_n) break;
    const int64_t pix = pix_n;
    uint32_t msk = 0;
    if (p.lab) msk |= lab_n;
    if (p.vd) msk |= vd_n;
    if (p.keep) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (((keep_n >> (8 * j)) & 0xFFu) == 0) msk |= 0xFFu << (8 * j);  // keep == 0 -> excluded
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    float v[C][4];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float4 r = pf_slot[c * kScoreThreads + threadIdx.x];
      v[c][0] = r.x, v[c][1] = r.y, v[c][2] = r.z, v[c][3] = r.w;
    }
    issue(it + 1);  // the slot is in registers: refill it before scoring
    float out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float xs[C];
#pragma unroll
      for (int c = 0; c < C; ++c) xs[c] = v[c][j];
      out[j] = score_from_logits<C, STRAT>(xs);
      if ((msk >> (8 * j)) & 0xFFu) out[j] = p.fill;
    }
    *reinterpret_cast<float4*>(p.score + pix) = make_float4(out[0], out[1], out[2], out[3]);
    if (HIST) {
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&sh_hist[bucket0(out[j], largest)], 1u);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (HIST) {
    __syncthreads();
    uint32_t* gh = p.hist0 + (size_t)img * kHistBins;
    for (int i = threadIdx.x; i < kHistBins; i += kScoreThreads) {
      const uint32_t c = sh_hist[i];
      if (c) atomicAdd(gh + i, c);
    }
  }
}

// Scalar fallback: any C, any W, any stride/alignment. One thread = one pixel.
template <int STRAT, typename T>
__global__ void __launch_bounds__(kScoreThreads) acq_score_scalar_kernel(const ScoreParams p) {
  __shared__ uint32_t sh_hist[kHistBins];
  const bool hist = p.hist0 != nullptr;
  if (hist) {
    for (int i = threadIdx.x; i < kHistBins; i += kScoreThreads) sh_hist[i] = 0;
    __syncthreads();
  }
  const int img = blockIdx.y;
  const int64_t HW = (int64_t)p.H * p.W;
  const T* __restrict__ base = reinterpret_cast<const T*>(p.logits) + (int64_t)img * p.sn;
  for (int64_t i = (int64_t)blockIdx.x * kScoreThreads + threadIdx.x; i < HW;
       i += (int64_t)gridDim.x * kScoreThreads) {
    const int y = (int)(i / p.W);
    const int x = (int)(i - (int64_t)y * p.W);
    const T* __restrict__ src = base + (int64_t)y * p.sh + x;
    const int64_t sc = p.sc;
    float s = score_runtime_c<STRAT>(p.C, [&](int c) { return Ld4<T>::ld1(src + (int64_t)c * sc); });
    const int64_t pix = (int64_t)img * HW + i;
    bool masked = false;
    if (p.lab) masked |= p.lab[pix] != 0;
    if (p.vd) masked |= p.vd[pix] != 0;
    if (p.keep) masked |= p.keep[pix] == 0;
    if (masked) s = p.fill;
    p.score[pix] = s;
    if (hist) atomicAdd(&sh_hist[bucket0(s, p.largest != 0)], 1u);
  }
  if (hist) {
    __syncthreads();
    uint32_t* gh = p.hist0 + (size_t)img * kHistBins;
    for (int i = threadIdx.x; i < kHistBins; i += kScoreThreads) {
      const uint32_t c = sh_hist[i];
      if (c) atomicAdd(gh + i, c);
    }
  }
}

// Fused bilinear(align_corners=True) upsample + score: reads 1/s-resolution logits.
struct ScoreUpParams {
  const float* logits;  // [n, C, h_in, w_in]
  int n_img, C, h_in, w_in, H, W;
  float scale_h, scale_w;
  const uint8_t* lab;
  const uint8_t* vd;
  const uint8_t* keep;
  float* score;
  uint32_t* hist0;
  float fill;
  int largest;
};

template <int C, int STRAT>
__global__ void __launch_bounds__(kScoreThreads) acq_score_up_kernel(const ScoreUpParams p) {
  __shared__ uint32_t sh_hist[kHistBins];
  const bool hist = p.hist0 != nullptr;
  if (hist) {
    for (int i = threadIdx.x; i < kHistBins; i += kScoreThreads) sh_hist[i] = 0;
    __syncthreads();
  }
  const int img = blockIdx.y;
  const int64_t HW = (int64_t)p.H * p.W;
  const int64_t plane = (int64_t)p.h_in * p.w_in;
  const float* __restrict__ base = p.logits + (int64_t)img * p.C * plane;
  for (int64_t i = (int64_t)blockIdx.x * kScoreThreads + threadIdx.x; i < HW;
       i += (int64_t)gridDim.x * kScoreThreads) {
    const int y = (int)(i / p.W);
    const int x = (int)(i - (int64_t)y * p.W);
    const Lerp ly = lerp_ac(y, p.h_in, p.H, p.scale_h);
    const Lerp lx = lerp_ac(x, p.w_in, p.W, p.scale_w);
    const int x0 = lx.i0, x1 = lx.i1;
    const float ly0 = ly.l0, ly1 = ly.l1, lx0 = lx.l0, lx1 = lx.l1;
    const float* r0 = base + (int64_t)ly.i0 * p.w_in;
    const float* r1 = base + (int64_t)ly.i1 * p.w_in;
    float xs[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float v00 = __ldg(r0 + c * plane + x0), v01 = __ldg(r0 + c * plane + x1);
      const float v10 = __ldg(r1 + c * plane + x0), v11 = __ldg(r1 + c * plane + x1);
      xs[c] = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
    }
    float s = score_from_logits<C, STRAT>(xs);
    const int64_t pix = (int64_t)img * HW + i;
    bool masked = false;
    if (p.lab) masked |= p.lab[pix] != 0;
    if (p.vd) masked |= p.vd[pix] != 0;
    if (p.keep) masked |= p.keep[pix] == 0;
    if (masked) s = p.fill;
    p.score[pix] = s;
    if (hist) atomicAdd(&sh_hist[bucket0(s, p.largest != 0)], 1u);
  }
  if (hist) {
    __syncthreads();
    uint32_t* gh = p.hist0 + (size_t)img * kHistBins;
    for (int i = threadIdx.x; i < kHistBins; i += kScoreThreads) {
      const uint32_t c = sh_hist[i];
      if (c) atomicAdd(gh + i, c);
    }
  }
}

// ------------------------------------------------------------------------------------------
// radix select
// ------------------------------------------------------------------------------------------
constexpr int kLevels = 5;
__constant__ int c_shift[kLevels + 1] = {53, 42, 32, 11, 0, 0};
__constant__ int c_bits[kLevels + 1] = {11, 11, 10, 11, 11, 0};
constexpr int kSelThreads = 256;
constexpr int kSelItems = 16;
constexpr int kSelTile = kSelThreads * kSelItems;  // 4096

// Per-image output counters live on their OWN 128-byte line (word 0: candidates, word 1: boundary entries).  Packed as two
// [n_img] arrays, the counters of every image being classified at one time sat in one line, i.e. in ONE L2 slice, which then
// served all 131 072 atomics of a 256-image step one after the other (ncu: lts__t_tag_requests max over slices 74 % of
// peak against 28 % on average while the level-0 kernel ran at 3.3 TB/s whatever its occupancy or prefetch depth).
constexpr int kCntStride = 32;

struct SelState {
  uint32_t remaining;
  uint32_t done;
  uint32_t bucket;  // level-0 bucket (written by pick_bucket0_kernel, read by select_l0_kernel)
  uint32_t klo;     // smallest ordering key that falls in `bucket`
  uint32_t khi;     // largest ordering key that falls in `bucket` (inclusive)
  uint32_t pad;
};

struct Workspace {
  uint32_t* hist;        // [n_img][2048] level-0 histogram (the later levels are histogrammed in shared memory)
  SelState* state;       // [kLevels+1][n_img]
  uint32_t* cand_count;  // [n_img][kCntStride]: word 0 = candidates written so far
  uint32_t* filt_count;  // = cand_count + 1: word 1 = entries of the boundary list
  uint64_t* cand;        // [n_img][kpad]
  uint64_t* filt;        // [2][n_img][HW]
  size_t zero_bytes;     // hist..filt_count are contiguous from the workspace base
  size_t total_bytes;
  int kpad;
};

static int next_pow2(int x) {
  int p = 32;
  while (p < x) p <<= 1;
  return p;
}

static Workspace carve(void* base, int n_img, int HW, int k) {
  Workspace w;
  char* p = reinterpret_cast<char*>(base);
  size_t off = 0;
  w.hist = reinterpret_cast<uint32_t*>(p + off);
  off += align_up((size_t)n_img * kHistBins * sizeof(uint32_t), 256);
  w.state = reinterpret_cast<SelState*>(p + off);
  off += align_up((size_t)(kLevels + 1) * n_img * sizeof(SelState), 256);
  w.cand_count = reinterpret_cast<uint32_t*>(p + off);
  w.filt_count = w.cand_count + 1;
  off += align_up((size_t)n_img * kCntStride * sizeof(uint32_t), 256);
  w.zero_bytes = off;
  w.kpad = next_pow2(k);
  w.cand = reinterpret_cast<uint64_t*>(p + off);
  off += align_up((size_t)n_img * w.kpad * sizeof(uint64_t), 256);
  w.filt = reinterpret_cast<uint64_t*>(p + off);
  off += align_up((size_t)2 * n_img * HW * sizeof(uint64_t), 256);
  w.total_bytes = off;
  return w;
}

struct SelParams {
  const float* scores;  // level 0 input [n_img][HW]
  const uint64_t* in_list;
  const uint32_t* in_count;
  uint64_t* out_list;
  uint32_t* out_count;
  uint64_t* cand;
  uint32_t* cand_count;
  const SelState* state_cur;
  SelState* state_next;
  int level, n_img, HW, k, kpad, largest;
};

// standalone level-0 histogram (used when pp_acq_score did not fuse it)
__global__ void __launch_bounds__(kSelThreads) hist0_kernel(const float* __restrict__ scores,
                                                             uint32_t* __restrict__ hist0, int HW,
                                                             int largest) {
  __shared__ uint32_t sh_hist[kHistBins];
  for (int i = threadIdx.x; i < kHistBins; i += kSelThreads) sh_hist[i] = 0;
  __syncthreads();
  const int img = blockIdx.y;
  const float* s = scores + (size_t)img * HW;
  for (int i = blockIdx.x * kSelThreads + threadIdx.x; i < HW; i += gridDim.x * kSelThreads)
    atomicAdd(&sh_hist[bucket0(__ldg(s + i), largest != 0)], 1u);
  __syncthreads();
  uint32_t* gh = hist0 + (size_t)img * kHistBins;
  for (int i = threadIdx.x; i < kHistBins; i += kSelThreads) {
    const uint32_t c = sh_hist[i];
    if (c) atomicAdd(gh + i, c);
  }
}

// One CTA per image: find the bucket of the level-0 histogram that holds the k-th element; state[img] = {rank inside
// the bucket, whole bucket selected?, bucket}.  Done once here instead of in the prologue of every select_l0 CTA.
__global__ void __launch_bounds__(kSelThreads) pick_bucket0_kernel(const uint32_t* __restrict__ hist0, SelState* __restrict__ state,
                                                                    uint32_t k, bool largest) {
  __shared__ uint32_t sh_warp[kSelThreads / 32];
  __shared__ uint32_t sh_bucket;
  const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint4* hc = reinterpret_cast<const uint4*>(hist0 + (size_t)img * kHistBins) + tid * 2;
  const uint4 ha = hc[0], hb = hc[1];
  const uint32_t h[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
  uint32_t mine = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) mine += h[i];
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) sh_warp[warp] = incl;
  __syncthreads();
  uint32_t wbase = 0;
#pragma unroll
  for (int w = 0; w < kSelThreads / 32; ++w)
    if (w < warp) wbase += sh_warp[w];
  uint32_t run = wbase + incl - mine;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (run < k && k <= run + h[i]) {  // exactly one (thread, i) matches
      SelState st;
      st.remaining = k - run;
      st.done = (h[i] == k - run) ? 1u : 0u;
      st.bucket = (uint32_t)(tid * 8 + i);
      st.klo = st.khi = 0u;  // filled by warp 0 below
      st.pad = 0u;
      state[img] = st;
      sh_bucket = st.bucket;
    }
    run += h[i];
  }
  __syncthreads();
  // key range [klo, khi] of the bucket: bucket0(score) is monotone in the ordering key, so "bucket(s) < b" is
  // "key(s) < klo" — select_l0 then classifies with unsigned compares instead of re-quantising every score.  Warp 0 finds
  // the two thresholds by 32-ary search over the key space (7 rounds each).
  if (warp == 0) {
    auto bucket_of_key = [&](uint32_t key) -> uint32_t {
      // keys that no score produces (the NaN payload ranges beyond +-inf) are clamped so the map stays monotone
      uint32_t u = largest ? ~key : key;
      if (u == 0xFFFFFFFFu) return largest ? 0u : 2047u;  // the canonical NaN key
      u = u < 0x007FFFFFu ? 0x007FFFFFu : (u > 0xFF800000u ? 0xFF800000u : u);  // [-inf, +inf]
      const uint32_t bits = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
      return bucket0(__uint_as_float(bits), largest);
    };
    auto first_key_with_bucket_ge = [&](uint32_t b, bool& none) -> uint32_t {
      none = bucket_of_key(0xFFFFFFFFu) < b;
      if (none) return 0u;
      uint64_t lo = 0, hi = 0xFFFFFFFFull;  // invariant: the answer is in [lo, hi]
      while (lo < hi) {
        const uint64_t step = (hi - lo + 32) / 33;  // lanes probe lo + (lane + 1) * step - 1 (clamped to hi)
        uint64_t probe = lo + (uint64_t)(lane + 1) * step - 1;
        probe = probe > hi ? hi : probe;
        const bool ok = bucket_of_key((uint32_t)probe) >= b;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, ok);
        if (m == 0u) {
          lo = lo + 32 * step;  // beyond the last probe (which is < hi here, otherwise ok would hold for hi)
          lo = lo > hi ? hi : lo;
        } else {
          const int first = __ffs(m) - 1;
          const uint64_t p_first = lo + (uint64_t)(first + 1) * step - 1;
          hi = p_first > hi ? hi : p_first;
          if (first > 0) lo = lo + (uint64_t)first * step;  // one past the previous probe
        }
      }
      return (uint32_t)lo;
    };
    const uint32_t b = sh_bucket;
    if (!largest) {  // bucket = the key's leading 11 bits: the range is explicit
      if (lane == 0) {
        state[img].klo = b << 21;
        state[img].khi = (b << 21) | 0x1FFFFFu;
      }
    } else {
      bool none;
      const uint32_t klo = first_key_with_bucket_ge(b, none);
      const uint32_t nxt = first_key_with_bucket_ge(b + 1u, none);
      if (lane == 0) {
        state[img].klo = klo;
        state[img].khi = none ? 0xFFFFFFFFu : nxt - 1u;
      }
    }
  }
}

// Level 0 of the radix select: ONE pass over the score map.  Each CTA owns one contiguous chunk of 8192 scores, keeps
// their 32-bit ordering keys in registers (8 x 16-byte streaming loads per thread, all issued before the first use),
// classifies them against the level-0 bucket with two unsigned range compares, and claims its output ranges with ONE
// pair of global atomics.  ncu on the previous versions: the tile-walking kernel (atomic round trip + 3 barriers per
// 4096 scores) ran at 1.7 TB/s; a first single-pass version was ISSUE-bound (58 instructions per score: keys computed
// twice, per-score range checks, 64-bit masks) — hence the FULL fast path and the key-in-place layout here.
constexpr int kL0Items = 16;
constexpr int kL0WarpChunk = 32 * kL0Items;                       // 512 scores per warp
constexpr int kL0Chunk = (kSelThreads / 32) * kL0WarpChunk;       // 4096 scores per CTA
// Every WARP is autonomous: it owns 512 consecutive scores (four 16-byte streaming loads per lane, all in flight before the
// first use), classifies them, scans its counts with shuffles and claims its output ranges with one pair of warp-aggregated
// global atomics - no block barrier.  Measured on B200 (256 images of 256x512, ncu): one 2048-score chunk per CTA with a block
// scan around one atomic pair 61-70 us; warp-autonomous with per-item predicated blocks 68 us (3.3 IPC, 83 % issue slots busy,
// 12 of 32 lanes active); survivors walked by set bits 56 us; THIS form (float pre-filter, ballot compaction of the survivors
// into shared memory, one survivor per lane) 53 us = 2.5 TB/s; a 4-score group filter with a second 16-byte read of the
// survivors 61 us.  What is left is the chain load -> filter -> atomic round trip -> store per warp at 75 % occupancy.
template <bool FULL /* every score of the chunk is in range and 16-byte loadable */>
__global__ void __launch_bounds__(kSelThreads, 6) select_l0_kernel(const SelParams p) {
  // per warp: the indices of the scores that survive the float pre-filter, then their ordering keys (class in the top bits)
  __shared__ uint32_t sh_idx[kSelThreads / 32][kL0WarpChunk];
  __shared__ uint32_t sh_key[kSelThreads / 32][kL0WarpChunk];
  const int img = blockIdx.y;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const uint32_t n_in = (uint32_t)p.HW;
  const float* sc = p.scores + (size_t)img * p.HW;
  const uint32_t chunk0 = blockIdx.x * (uint32_t)kL0Chunk + (uint32_t)warp * (uint32_t)kL0WarpChunk;
  const bool largest = p.largest != 0;

  // FULL: item i = 4 j + e is score chunk0 + (j * 32 + lane) * 4 + e;  otherwise item i is score chunk0 + i * 32 + lane
  uint32_t raw[kL0Items];
  if (FULL) {
#pragma unroll
    for (int j = 0; j < kL0Items / 4; ++j) {
      const float* src = sc + chunk0 + (uint32_t)(j * 32 + lane) * 4u;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(raw[4 * j]), "=r"(raw[4 * j + 1]), "=r"(raw[4 * j + 2]), "=r"(raw[4 * j + 3])
                   : "l"(src));
    }
  } else {
#pragma unroll
    for (int i = 0; i < kL0Items; ++i) {
      const uint32_t idx = chunk0 + (uint32_t)(i * 32 + lane);
      raw[i] = __float_as_uint((idx < n_in) ? __ldg(sc + idx) : 0.f);
    }
  }
  // ---- level-0 bucket of this image (pick_bucket0_kernel): read while the loads are in flight ----
  const SelState st0 = p.state_next[img];
  const bool take_all = st0.done != 0u;
  const uint32_t klo = st0.klo, khi = st0.khi;
  const bool sel_any = take_all || klo > 0u;
  const uint32_t sel_max = take_all ? khi : klo - 1u;
  const uint32_t idx0 = FULL ? chunk0 + (uint32_t)lane * 4u : chunk0 + (uint32_t)lane;
  // ---- stage 1: ~95 % of the scores are neither selected nor on the boundary and the kernel is bound by the instructions it
  // issues (ncu: 3.3 IPC, 83 % issue slots busy, 12 of 32 lanes active inside per-item predicated blocks), so every score
  // takes ONE float compare against the score that maps to khi - a superset test: NaN and the threshold's ties pass - and
  // the survivors' indices are compacted per warp (ballot + popc) into shared memory ----
  const float f_t = ord_key_inv(khi, largest);
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t n_pass = 0;  // warp-uniform
#pragma unroll
  for (int i = 0; i < kL0Items; ++i) {
    const float v = __uint_as_float(raw[i]);
    const uint32_t idx = FULL ? idx0 + (uint32_t)((i >> 2) * 32 * 4 + (i & 3)) : idx0 + (uint32_t)(i * 32);
    const bool ps = (FULL || idx < n_in) && (largest ? !(v < f_t) : !(v > f_t));
    const uint32_t b = __ballot_sync(0xFFFFFFFFu, ps);
    if (ps) {
      const uint32_t e = n_pass + (uint32_t)__popc(b & lt);
      sh_idx[warp][e] = idx;
      sh_key[warp][e] = raw[i];
    }
    n_pass += (uint32_t)__popc(b);
  }
  if (n_pass == 0u) return;  // warp-uniform
  __syncwarp();
  // ---- stage 2: the survivors, one per lane and round (all lanes busy): exact classification on the ordering key against
  // the bucket's key range [klo, khi]: selected <=> below the bucket (or inside it when the whole bucket is taken),
  // boundary <=> inside it ----
  uint32_t nc = 0, nf = 0;
  for (uint32_t e = (uint32_t)lane; e < n_pass; e += 32u) {
    const uint32_t k = ord_key(__uint_as_float(sh_key[warp][e]), largest);
    const bool s1 = sel_any && k <= sel_max;
    const bool s2 = !take_all && k >= klo && k <= khi;
    sh_key[warp][e] = k;
    sh_idx[warp][e] |= ((uint32_t)s1 << 30) | ((uint32_t)s2 << 31);  // pixel indices stay below 2^30
    nc += s1;
    nf += s2;
  }
  // warp inclusive scan of (nc, nf) packed 16:16 (a warp holds 512 items: each count fits 10 bits)
  const uint32_t packed = nc | (nf << 16);
  uint32_t inc = packed;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= o) inc += t;
  }
  const uint32_t tot = __shfl_sync(0xFFFFFFFFu, inc, 31);
  if (tot == 0u) return;  // warp-uniform: nothing of this chunk is selected or on the boundary
  uint32_t base_c = 0, base_f = 0;
  if (lane == 0) {
    const uint32_t tc = tot & 0xFFFFu, tf = tot >> 16;
    if (tc) base_c = atomicAdd(p.cand_count + (size_t)img * kCntStride, tc);
    if (tf) base_f = atomicAdd(p.out_count + (size_t)img * kCntStride, tf);
  }
  base_c = __shfl_sync(0xFFFFFFFFu, base_c, 0);
  base_f = __shfl_sync(0xFFFFFFFFu, base_f, 0);
  const uint32_t excl = inc - packed;
  uint32_t oc = base_c + (excl & 0xFFFFu);
  uint32_t of = base_f + (excl >> 16);
  uint64_t* cand = p.cand + (size_t)img * p.kpad;
  uint64_t* ol = p.out_list + (size_t)img * p.HW;
  for (uint32_t e = (uint32_t)lane; e < n_pass; e += 32u) {
    const uint32_t ix = sh_idx[warp][e];
    if (ix >> 30) {
      const uint64_t comp = ((uint64_t)sh_key[warp][e] << 32) | (ix & 0x3FFFFFFFu);
      if (ix & 0x40000000u) {
        if (oc < (uint32_t)p.kpad) cand[oc] = comp;
        ++oc;
      } else {
        ol[of++] = comp;
      }
    }
  }
}

// Level 0, staged form (the default when the score map is 16-byte aligned and H*W is a multiple of 512).
// ncu on select_l0_kernel (256 images of 256x512): ~1 warp instruction per score (512 SASS instructions per 512-score
// warp chunk: a ballot + two popc + two predicated shared stores for EVERY score) at 2.3-2.5 IPC - the kernel is bound by the
// instructions it issues, 53 us for 134 MB.  Measured on the way here (same box, select phase = pick_bucket0 + this kernel +
// select_rest, legacy 68.5 us): a persistent kernel that walks the survivors by set bits out of a cp.async ring, per lane, 76.4;
// the same with the bucket state loaded one chunk ahead and the output deferred by one chunk 73.7; that one without atomics 61.8,
// without atomics and stores 53.8 - i.e. neither the atomic round trip nor the scattered stores but the 437 instructions per
// chunk (two divergent per-lane loops of ~3 trips at 0.8 active lanes) and an L2 load of the image's state on every chunk.  Here
//   * the grid is persistent; every warp owns ONE contiguous range of chunks, so the bucket state is reloaded only when the
//     range crosses into the next image, and walks it through a private ring in shared memory filled with cp.async (16 bytes
//     per lane and piece).  Once the output counters had their own cache lines (kCntStride) the kernel became sensitive to
//     occupancy, not to prefetch depth: select phase 54.2 / 50.2 / 48.1 us with 4 / 3 / 2 ring stages at 3 / 4 / 5 CTAs per
//     SM, hence the default of 2 stages;
//   * a score costs one float compare + one mask update; the lanes' survivors (~5 %) are compacted into a per-warp list of
//     9-bit chunk offsets behind one warp scan, and then classified on the exact ordering key and written DENSELY, one survivor
//     per lane and trip, behind one warp-aggregated atomic pair, as in the legacy kernel.
// Output = the same two unordered sets (candidates below the bucket, boundary list inside it).
constexpr int kL0sWarps = kSelThreads / 32;
template <int STAGES> struct L0sCfg {
  static constexpr int kWarpBytes = STAGES * kL0WarpChunk * (int)sizeof(float) + kL0WarpChunk * (int)sizeof(uint16_t);
  static constexpr int kSmemBytes = kL0sWarps * kWarpBytes;           // 4 stages: 72 KB, 3: 56 KB, 2: 40 KB
  static constexpr int kCtasPerSm = STAGES >= 4 ? 3 : (STAGES == 3 ? 4 : 5);
};

struct L0sWalk {
  uint32_t chunks_per_img;  // H*W / 512
  uint32_t total_chunks;    // n_img * chunks_per_img
  uint32_t base, extra;     // warp g of the grid owns base + (g < extra) consecutive chunks
};

template <int kL0sStages>
__global__ void __launch_bounds__(kSelThreads, L0sCfg<kL0sStages>::kCtasPerSm) select_l0_staged_kernel(const SelParams p, const L0sWalk w) {
  constexpr int kL0sWarpBytes = L0sCfg<kL0sStages>::kWarpBytes;
  extern __shared__ __align__(16) unsigned char l0s_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* ring = reinterpret_cast<float*>(l0s_smem + warp * kL0sWarpBytes);
  uint16_t* list = reinterpret_cast<uint16_t*>(ring + kL0sStages * kL0WarpChunk);  // survivors of the current chunk
  const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
  const uint32_t gw = blockIdx.x * (uint32_t)kL0sWarps + (uint32_t)warp;
  const bool largest = p.largest != 0;
  const uint32_t full = 0xFFFFFFFFu;
  const uint32_t first = gw * w.base + (gw < w.extra ? gw : w.extra);
  const uint32_t last = first + w.base + (gw < w.extra ? 1u : 0u);  // exclusive
  if (first >= last) return;

  // lane's four 16-byte pieces of chunk wc -> ring slot `slot` (the ring holds a linear copy of the chunk)
  auto issue = [&](uint32_t wc, int slot) {
    const float* src = p.scores + (size_t)wc * kL0WarpChunk + lane * 4;
    const uint32_t dst = ring_s + (uint32_t)(slot * kL0WarpChunk + lane * 4) * 4u;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)j * 512u), "l"(src + j * 128) : "memory");
  };
#pragma unroll
  for (int s = 0; s < kL0sStages - 1; ++s) {
    if (first + s < last) issue(first + s, s);
    asm volatile("cp.async.commit_group;" ::: "memory");  // one group per stage, empty or not: the count below stays uniform
  }
  uint32_t img = first / w.chunks_per_img, rem = first - img * w.chunks_per_img;  // chunk `rem` of image `img`
  // level-0 bucket of the image (pick_bucket0_kernel)
  bool take_all = false, sel_any = false;
  uint32_t klo = 0, khi = 0, sel_max = 0;
  float f_t = 0.f;
  auto load_state = [&]() {
    const SelState* sp = p.state_next + img;
    take_all = sp->done != 0u;
    klo = sp->klo, khi = sp->khi;
    sel_any = take_all || klo > 0u;
    sel_max = take_all ? khi : klo - 1u;
    f_t = ord_key_inv(khi, largest);
  };
  load_state();
  uint64_t* cand = p.cand + (size_t)img * p.kpad;
  uint64_t* ol = p.out_list + (size_t)img * p.HW;
  int slot = 0;
  for (uint32_t wc = first; wc < last; ++wc) {
    {  // refill the slot consumed by the previous iteration
      int ps = slot + kL0sStages - 1;
      ps = ps >= kL0sStages ? ps - kL0sStages : ps;
      if (wc + (kL0sStages - 1) < last) issue(wc + (kL0sStages - 1), ps);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.wait_group %0;" ::"n"(kL0sStages - 1) : "memory");  // the oldest group (this chunk) has landed
    __syncwarp();  // ... for every lane: the survivors below are read across lanes
    const float* sl = ring + slot * kL0WarpChunk;
    slot = slot + 1 == kL0sStages ? 0 : slot + 1;
    const uint32_t px0 = rem * (uint32_t)kL0WarpChunk;  // flat pixel index of the chunk's first score
    // ---- stage 1: one compare per score against the score that maps to khi (a superset test: NaN and the threshold's
    // ties pass); item i = 4 j + e of a lane is the chunk's score (j * 32 + lane) * 4 + e ----
    uint32_t mask = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 v = *reinterpret_cast<const float4*>(sl + (j * 32 + lane) * 4);
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const bool ps = largest ? !(e[q] < f_t) : !(e[q] > f_t);
        mask |= ps ? (1u << (4 * j + q)) : 0u;
      }
    }
    // ---- compaction: the lanes' survivors -> list[0 .. n_pass) of chunk offsets (lane-major) ----
    const uint32_t mine = (uint32_t)__popc(mask);
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(full, inc, o);
      if (lane >= o) inc += t;
    }
    const uint32_t n_pass = __shfl_sync(full, inc, 31);
    if (n_pass != 0u) {  // warp-uniform
      uint32_t off = inc - mine;
      for (uint32_t m = mask; m; m &= m - 1u) {
        const int i = __ffs((int)m) - 1;
        list[off++] = (uint16_t)(((i >> 2) * 32 + lane) * 4 + (i & 3));
      }
      __syncwarp();
      // ---- stage 2: one survivor per lane and trip, exact classification on the ordering key against the bucket's key range:
      // selected <=> below the bucket (or inside it when the whole bucket is taken), boundary <=> inside it ----
      uint32_t nc = 0, nf = 0;
      for (uint32_t e = (uint32_t)lane; e < n_pass; e += 32u) {
        const uint32_t li = list[e];
        const uint32_t k = ord_key(sl[li], largest);
        const bool s1 = sel_any && k <= sel_max;
        const bool s2 = !take_all && k >= klo && k <= khi;
        list[e] = (uint16_t)(li | ((uint32_t)s1 << 14) | ((uint32_t)s2 << 15));  // read back by this lane only
        nc += s1;
        nf += s2;
      }
      const uint32_t packed = nc | (nf << 16);  // a warp holds 512 scores: each count fits 10 bits
      uint32_t pin = packed;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(full, pin, o);
        if (lane >= o) pin += t;
      }
      const uint32_t tot = __shfl_sync(full, pin, 31);
      if (tot != 0u) {  // warp-uniform
        uint32_t base_c = 0, base_f = 0;
        if (lane == 0) {
          // both counters of the image with ONE 64-bit atomic (they share an aligned 8-byte word, see kCntStride)
          const unsigned long long add = (unsigned long long)(tot & 0xFFFFu) | ((unsigned long long)(tot >> 16) << 32);
          const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long*>(p.cand_count + (size_t)img * kCntStride), add);
          base_c = (uint32_t)old;
          base_f = (uint32_t)(old >> 32);
        }
        base_c = __shfl_sync(full, base_c, 0);
        base_f = __shfl_sync(full, base_f, 0);
        const uint32_t excl = pin - packed;
        uint32_t oc = base_c + (excl & 0xFFFFu);
        uint32_t of = base_f + (excl >> 16);
        for (uint32_t e = (uint32_t)lane; e < n_pass; e += 32u) {
          const uint32_t ent = list[e];
          if (ent >> 14) {
            const uint32_t li = ent & 0x1FFu;
            const uint64_t comp = ((uint64_t)ord_key(sl[li], largest) << 32) | (uint64_t)(px0 + li);
            if (ent & 0x4000u) {
              if (oc < (uint32_t)p.kpad) cand[oc] = comp;
              ++oc;
            } else {
              ol[of++] = comp;
            }
          }
        }
      }
      __syncwarp();  // the list is rewritten by the next chunk
    }
    if (++rem == w.chunks_per_img && wc + 1 < last) {  // warp-uniform: the range crosses into the next image
      rem = 0;
      ++img;
      load_state();
      cand = p.cand + (size_t)img * p.kpad;
      ol = p.out_list + (size_t)img * p.HW;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");  // nothing of this CTA is in flight when its shared memory is released
}

// ------------------------------------------------------------------------------------------------------------------
// Fused scoring + level-0 select: ONE pass over the logits, the score map never touches HBM.
// A thread-block cluster owns one image: every CTA scores HW / cluster pixels (the same 4-pixel vector loop as
// acq_score_vec_kernel), keeps its scores in shared memory and builds its level-0 histogram there; the histograms are added into
// the leader CTA's shared memory through distributed shared memory, the leader finds the bucket of the k-th score and its key
// range (the work of pick_bucket0_kernel), every CTA reads that back and classifies ITS OWN scores out of shared memory -
// selected -> candidate list, inside the bucket -> boundary list for select_rest_kernel.  Against the three-kernel form
// (score: 78 B/px read + 4 B/px written; pick_bucket0; select_l0: 4 B/px re-read) this moves 78 B/px and saves two launches.
// ------------------------------------------------------------------------------------------------------------------
namespace cg = cooperative_groups;
constexpr int kFusedMaxPx = 16384;  // scores per CTA held in shared memory (64 KB)

struct FusedSelParams {
  SelState* state;       // [n_img] level-0 state (what pick_bucket0_kernel writes)
  uint64_t* cand;        // [n_img][kpad]
  uint32_t* cand_count;  // [n_img]
  uint64_t* bnd;         // [n_img][HW] boundary list
  uint32_t* bnd_count;   // [n_img]
  uint32_t k;
  int kpad, pxc;         // pixels per CTA (HW / cluster size, a multiple of 1024)
};

// the body of pick_bucket0_kernel on a histogram in shared memory; every thread of the CTA calls it, result in *out (shared)
__device__ void pick_bucket0_block(const uint32_t* hist, SelState* out, uint32_t k, bool largest, uint32_t* sh_warp, uint32_t* sh_bucket) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t h[8];
  uint32_t mine = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    h[i] = hist[tid * 8 + i];
    mine += h[i];
  }
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) sh_warp[warp] = incl;
  __syncthreads();
  uint32_t wbase = 0;
#pragma unroll
  for (int w = 0; w < kSelThreads / 32; ++w)
    if (w < warp) wbase += sh_warp[w];
  uint32_t run = wbase + incl - mine;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (run < k && k <= run + h[i]) {  // exactly one (thread, i) matches
      out->remaining = k - run;
      out->done = (h[i] == k - run) ? 1u : 0u;
      out->bucket = (uint32_t)(tid * 8 + i);
      out->klo = out->khi = 0u;
      out->pad = 0u;
      *sh_bucket = out->bucket;
    }
    run += h[i];
  }
  __syncthreads();
  if (warp == 0) {
    auto bucket_of_key = [&](uint32_t key) -> uint32_t {
      uint32_t u = largest ? ~key : key;
      if (u == 0xFFFFFFFFu) return largest ? 0u : 2047u;
      u = u < 0x007FFFFFu ? 0x007FFFFFu : (u > 0xFF800000u ? 0xFF800000u : u);
      const uint32_t bits = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
      return bucket0(__uint_as_float(bits), largest);
    };
    auto first_key_with_bucket_ge = [&](uint32_t b, bool& none) -> uint32_t {
      none = bucket_of_key(0xFFFFFFFFu) < b;
      if (none) return 0u;
      uint64_t lo = 0, hi = 0xFFFFFFFFull;
      while (lo < hi) {
        const uint64_t step = (hi - lo + 32) / 33;
        uint64_t probe = lo + (uint64_t)(lane + 1) * step - 1;
        probe = probe > hi ? hi : probe;
        const bool ok = bucket_of_key((uint32_t)probe) >= b;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, ok);
        if (m == 0u) {
          lo = lo + 32 * step;
          lo = lo > hi ? hi : lo;
        } else {
          const int first = __ffs(m) - 1;
          const uint64_t p_first = lo + (uint64_t)(first + 1) * step - 1;
          hi = p_first > hi ? hi : p_first;
          if (first > 0) lo = lo + (uint64_t)first * step;
        }
      }
      return (uint32_t)lo;
    };
    const uint32_t b = *sh_bucket;
    if (!largest) {
      if (lane == 0) {
        out->klo = b << 21;
        out->khi = (b << 21) | 0x1FFFFFu;
      }
    } else {
      bool none;
      const uint32_t klo = first_key_with_bucket_ge(b, none);
      const uint32_t nxt = first_key_with_bucket_ge(b + 1u, none);
      if (lane == 0) {
        out->klo = klo;
        out->khi = none ? 0xFFFFFFFFu : nxt - 1u;
      }
    }
  }
  __syncthreads();
}

template <int C, int STRAT>
__global__ void __launch_bounds__(kScoreThreads, 2) acq_score_select_kernel(const ScoreParams p, const FusedSelParams f) {
  extern __shared__ __align__(16) uint8_t fs_raw[];
  float* sh_score = reinterpret_cast<float*>(fs_raw);                                   // [pxc]
  uint32_t* sh_hist = reinterpret_cast<uint32_t*>(fs_raw + (size_t)kFusedMaxPx * 4);    // [2048] this CTA
  uint32_t* sh_tot = sh_hist + kHistBins;                                               // [2048] cluster total (leader's copy is used)
  uint32_t* sh_idx = sh_tot + kHistBins;                                                // [8 warps][512] survivors of the pre-filter
  __shared__ SelState sh_state;
  __shared__ uint32_t sh_warp[kScoreThreads / 32];
  __shared__ uint32_t sh_bucket;
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kHistBins; i += kScoreThreads) {
    sh_hist[i] = 0;
    sh_tot[i] = 0;
  }
  cluster.sync();  // the leader's total is zero before any peer adds into it

  const int img = blockIdx.y;
  const int Wv = p.W / 4;
  const int64_t HW = (int64_t)p.H * p.W;
  const float* __restrict__ base = reinterpret_cast<const float*>(p.logits) + (int64_t)img * p.sn;
  const bool largest = p.largest != 0;
  const int iters = f.pxc / (kScoreThreads * 4);
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    const int q = (blockIdx.x * iters + it) * kScoreThreads + tid;
    const int y = q / Wv;
    const int x = (q - y * Wv) * 4;
    const float* __restrict__ src = base + (int64_t)y * p.sh + x;
    float v[C][4];
#pragma unroll
    for (int c = 0; c < C; ++c) LdPx<float, 4>::ld(src + (int64_t)c * p.sc, v[c]);
    const int64_t pix = (int64_t)img * HW + (int64_t)y * p.W + x;
    uint32_t msk = 0;
    if (p.lab) msk |= ld_mask<4>(p.lab + pix);
    if (p.vd) msk |= ld_mask<4>(p.vd + pix);
    if (p.keep) {
      const uint32_t k4 = ld_mask<4>(p.keep + pix);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (((k4 >> (8 * j)) & 0xFFu) == 0) msk |= 0xFFu << (8 * j);
    }
    float out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float xs[C];
#pragma unroll
      for (int c = 0; c < C; ++c) xs[c] = v[c][j];
      out[j] = score_from_logits<C, STRAT>(xs);
      if ((msk >> (8 * j)) & 0xFFu) out[j] = p.fill;
    }
    const float4 o4 = make_float4(out[0], out[1], out[2], out[3]);
    reinterpret_cast<float4*>(sh_score)[it * kScoreThreads + tid] = o4;
    if (p.score) *reinterpret_cast<float4*>(p.score + pix) = o4;  // optional: the caller also wants the map
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(&sh_hist[bucket0(out[j], largest)], 1u);
  }
  __syncthreads();
  {  // this CTA's histogram -> the leader's total (distributed shared memory)
    uint32_t* tot0 = cluster.map_shared_rank(sh_tot, 0);
    for (int i = tid; i < kHistBins; i += kScoreThreads) {
      const uint32_t c = sh_hist[i];
      if (c) atomicAdd(tot0 + i, c);
    }
  }
  cluster.sync();
  if (cluster.block_rank() == 0) {
    pick_bucket0_block(sh_tot, &sh_state, f.k, largest, sh_warp, &sh_bucket);
    if (tid == 0) f.state[img] = sh_state;  // select_rest_kernel reads {remaining, done}
  }
  cluster.sync();
  const SelState st0 = *cluster.map_shared_rank(&sh_state, 0);
  // ---- classify this CTA's scores out of shared memory (the scheme of select_l0_kernel: float pre-filter, ballot compaction
  // of the survivors per warp, exact key compares one survivor per lane, one warp-aggregated atomic pair per 512 scores) ----
  const bool take_all = st0.done != 0u;
  const uint32_t klo = st0.klo, khi = st0.khi;
  const bool sel_any = take_all || klo > 0u;
  const uint32_t sel_max = take_all ? khi : klo - 1u;
  const float f_t = ord_key_inv(khi, largest);
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t* my_idx = sh_idx + warp * 512;
  uint64_t* cand = f.cand + (size_t)img * f.kpad;
  uint64_t* ol = f.bnd + (size_t)img * HW;
  const uint32_t cta0 = blockIdx.x * (uint32_t)f.pxc;
  const int per_warp = f.pxc / (kScoreThreads / 32);
  for (int r0 = 0; r0 < per_warp; r0 += 512) {
    const uint32_t wbase = (uint32_t)(warp * per_warp + r0);
    uint32_t n_pass = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 s4 = reinterpret_cast<const float4*>(sh_score + wbase)[j * 32 + lane];
      const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ps = largest ? !(sv[e] < f_t) : !(sv[e] > f_t);
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, ps);
        if (ps) my_idx[n_pass + (uint32_t)__popc(b & lt)] = wbase + (uint32_t)((j * 32 + lane) * 4 + e);
        n_pass += (uint32_t)__popc(b);
      }
    }
    __syncwarp();
    if (n_pass == 0u) continue;  // warp-uniform
    uint32_t nc = 0, nf = 0;
    for (uint32_t e = (uint32_t)lane; e < n_pass; e += 32u) {
      const uint32_t li = my_idx[e];
      const uint32_t kk = ord_key(sh_score[li], largest);
      const bool s1 = sel_any && kk <= sel_max;
      const bool s2 = !take_all && kk >= klo && kk <= khi;
      my_idx[e] = li | ((uint32_t)s1 << 30) | ((uint32_t)s2 << 31);
      nc += s1;
      nf += s2;
    }
    const uint32_t packed = nc | (nf << 16);
    uint32_t inc = packed;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
      if (lane >= o) inc += t;
    }
    const uint32_t tot = __shfl_sync(0xFFFFFFFFu, inc, 31);
    if (tot != 0u) {
      uint32_t base_c = 0, base_f = 0;
      if (lane == 0) {
        const uint32_t tc = tot & 0xFFFFu, tf = tot >> 16;
        if (tc) base_c = atomicAdd(f.cand_count + (size_t)img * kCntStride, tc);
        if (tf) base_f = atomicAdd(f.bnd_count + (size_t)img * kCntStride, tf);
      }
      base_c = __shfl_sync(0xFFFFFFFFu, base_c, 0);
      base_f = __shfl_sync(0xFFFFFFFFu, base_f, 0);
      const uint32_t excl = inc - packed;
      uint32_t oc = base_c + (excl & 0xFFFFu);
      uint32_t of = base_f + (excl >> 16);
      for (uint32_t e = (uint32_t)lane; e < n_pass; e += 32u) {
        const uint32_t ix = my_idx[e];
        if (ix >> 30) {
          const uint32_t li = ix & 0x3FFFFFFFu;
          const uint64_t comp = ((uint64_t)ord_key(sh_score[li], largest) << 32) | (cta0 + li);
          if (ix & 0x40000000u) {
            if (oc < (uint32_t)f.kpad) cand[oc] = comp;
            ++oc;
          } else {
            ol[of++] = comp;
          }
        }
      }
    }
    __syncwarp();
  }
  cluster.sync();  // no CTA may exit while a peer can still read its shared memory
}

// The radix tail for one image in ONE CTA: the boundary bucket of level 0 is small (L2-resident), so all key-digit
// radix levels (histogram in shared memory -> pick -> partition) run back to back without further launches or
// global histogram traffic.  Appends are warp-aggregated shared-memory atomics; order is irrelevant (sorted later).
constexpr int kRestThreads = 1024;
constexpr int kRestDirect = 256;  // lists up to this length are ranked directly
struct RestParams {
  uint64_t* list_a;        // [n_img][HW] boundary list written by level 0
  uint64_t* list_b;        // [n_img][HW] scratch
  uint32_t* count_a;       // boundary-list lengths (kCntStride apart)
  uint32_t* hist0;         // [n_img][2048] level-0 histograms (zeroed here)
  uint64_t* cand;
  uint32_t* cand_count;
  const SelState* state1;  // {remaining, done} after level 0
  int HW, kpad;
  int first_level;  // 1 when level 0 was the key's leading digit (it is constant in the boundary list), else 0
};

__global__ void __launch_bounds__(kRestThreads) select_rest_kernel(const RestParams p) {
  __shared__ uint32_t sh_hist[kHistBins];
  __shared__ uint32_t sh_warp[kRestThreads / 32];
  __shared__ uint32_t sh_pick[3];
  __shared__ uint32_t sh_cnt[2];  // cand count, next-list count
  __shared__ uint64_t sh_direct[kRestDirect];
  const int img = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // The select hands the workspace back "prepared" (pp_acq_topk_prepare): this kernel runs last and zeroes what the step
  // consumed - the image's level-0 histogram here (its only reader, pick_bucket0_kernel, is done; the stores drain under the
  // walk below - zeroing inside pick_bucket0_kernel itself cost that kernel 2 us), the two counters at the exits.
  if (tid < kHistBins / 4) reinterpret_cast<uint4*>(p.hist0 + (size_t)img * kHistBins)[tid] = make_uint4(0u, 0u, 0u, 0u);
  const SelState st = p.state1[img];
  if (st.done) {  // level 0 took a whole bucket: nothing left to resolve
    if (tid == 0) {
      p.cand_count[(size_t)img * kCntStride] = 0u;
      p.count_a[(size_t)img * kCntStride] = 0u;
    }
    return;
  }
  uint32_t rem = st.remaining;
  uint64_t* in = p.list_a + (size_t)img * p.HW;
  uint64_t* out = p.list_b + (size_t)img * p.HW;
  uint64_t* cand = p.cand + (size_t)img * p.kpad;
  uint32_t n = p.count_a[(size_t)img * kCntStride];
  if (tid == 0) sh_cnt[0] = p.cand_count[(size_t)img * kCntStride];
  for (int level = p.first_level; level < kLevels; ++level) {  // from the key's first digit when level 0 was the linear bucket0
    if (n <= (uint32_t)kRestDirect) {
      // A short list (the linear level-0 buckets of the largest-first strategies leave ~50-200 entries; after one radix level
      // any list is this short) is ranked directly: thread t counts the composites below its own - they are unique - and the
      // `rem` smallest are the selection.  One barrier pair instead of 2-4 more levels of histogram / scan / partition.
      if (tid < (int)n) sh_direct[tid] = in[tid];
      __syncthreads();
      if (tid < (int)n) {
        const uint64_t v = sh_direct[tid];
        uint32_t below = 0;
        for (uint32_t j = 0; j < n; ++j) below += (sh_direct[j] < v) ? 1u : 0u;
        if (below < rem) {
          const uint32_t pos = atomicAdd(&sh_cnt[0], 1u);
          if (pos < (uint32_t)p.kpad) cand[pos] = v;
        }
      }
      __syncthreads();
      break;
    }
    const int shift = c_shift[level];
    const uint32_t dmask = (1u << c_bits[level]) - 1u;
    for (int i = tid; i < kHistBins; i += kRestThreads) sh_hist[i] = 0;
    if (tid == 0) sh_cnt[1] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < n; i += kRestThreads) atomicAdd(&sh_hist[(uint32_t)(in[i] >> shift) & dmask], 1u);
    __syncthreads();
    // pick: thread t owns bins 2t, 2t+1
    const uint32_t h0 = sh_hist[2 * tid], h1 = sh_hist[2 * tid + 1];
    uint32_t incl = h0 + h1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) sh_warp[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += sh_warp[w];
    uint32_t run = wbase + incl - (h0 + h1);
    if (run < rem && rem <= run + h0) { sh_pick[0] = 2 * tid; sh_pick[1] = run; sh_pick[2] = h0; }
    run += h0;
    if (run < rem && rem <= run + h1) { sh_pick[0] = 2 * tid + 1; sh_pick[1] = run; sh_pick[2] = h1; }
    __syncthreads();
    const uint32_t bucket = sh_pick[0];
    rem -= sh_pick[1];
    const bool take_all = (sh_pick[2] == rem);
    // partition
    for (uint32_t base = 0; base < n; base += kRestThreads) {
      const uint32_t i = base + tid;
      uint64_t v = 0;
      int cls = 0;
      if (i < n) {
        v = in[i];
        const uint32_t d = (uint32_t)(v >> shift) & dmask;
        if (d < bucket || (d == bucket && take_all)) cls = 1;
        else if (d == bucket) cls = 2;
      }
      const uint32_t m1 = __ballot_sync(0xFFFFFFFFu, cls == 1);
      const uint32_t m2 = __ballot_sync(0xFFFFFFFFu, cls == 2);
      uint32_t b1 = 0, b2 = 0;
      if (lane == 0) {
        if (m1) b1 = atomicAdd(&sh_cnt[0], __popc(m1));
        if (m2) b2 = atomicAdd(&sh_cnt[1], __popc(m2));
      }
      b1 = __shfl_sync(0xFFFFFFFFu, b1, 0);
      b2 = __shfl_sync(0xFFFFFFFFu, b2, 0);
      const uint32_t lt = (1u << lane) - 1u;
      if (cls == 1) {
        const uint32_t pos = b1 + __popc(m1 & lt);
        if (pos < (uint32_t)p.kpad) cand[pos] = v;
      } else if (cls == 2) {
        out[b2 + __popc(m2 & lt)] = v;
      }
    }
    __syncthreads();
    if (take_all) break;
    n = sh_cnt[1];
    uint64_t* t = in; in = out; out = t;
    __syncthreads();
  }
  if (tid == 0) {  // every thread has read its counters before the barriers above: leave them zeroed for the next step
    p.cand_count[(size_t)img * kCntStride] = 0u;
    p.count_a[(size_t)img * kCntStride] = 0u;
  }
}

// Order statistics instead of a sort: the reference draws n random RANKS of the sorted top-k list
// (np.random.choice(ind_queries, n, False), query.py:63-64), so only the elements at those ranks are needed.
// One CTA per image runs a radix walk for up to kPickRanks ranks at once over the k unsorted composites: rank j keeps
// (prefix_j, rem_j); every level histograms the elements matching prefix_j by their next digit, then narrows.
// Composites are unique, so after the last level prefix_j IS the element of rank j.
constexpr int kPickThreads = 512;
constexpr int kPickRanks = 12;   // 12 x 2048 x 4 B = 96 KB of histograms
constexpr int kPickItems = 16;   // candidates cached in registers when k <= 512 * 16
constexpr int kPickUnroll = 8;   // candidates in flight per thread when they are streamed from L2 instead (larger k)

struct PickParams {
  const uint64_t* cand;  // [n_img][kpad]
  int kpad, k, n;
  const int32_t* pos;    // [n_img][n] ranks (nullptr: 0..n-1)
  int32_t* out;          // [n_img][n]
};

// The k candidates of one image through f(v, valid), streamed: kPickUnroll independent 8-byte loads per thread are issued
// before the first is used (5 % of a 1024x2048 image is 205 candidates per thread and pass: one load per trip left every
// pass a chain of 205 L2 round trips).  Every thread calls f the same number of times (warp collectives inside f are legal).
template <typename F>
__device__ __forceinline__ void pick_stream(const uint64_t* __restrict__ c, int k, int tid, F&& f) {
  const unsigned long long* cc = reinterpret_cast<const unsigned long long*>(c);
  int i0 = 0;
  for (; i0 + kPickUnroll * kPickThreads <= k; i0 += kPickUnroll * kPickThreads) {
    uint64_t v[kPickUnroll];
#pragma unroll
    for (int u = 0; u < kPickUnroll; ++u) v[u] = __ldg(cc + i0 + u * kPickThreads + tid);
#pragma unroll
    for (int u = 0; u < kPickUnroll; ++u) f(v[u], true);
  }
  for (; i0 < k; i0 += kPickThreads) {
    const int i = i0 + tid;
    const bool ok = i < k;
    f(ok ? (uint64_t)__ldg(cc + i) : 0ull, ok);
  }
}
template <bool CACHED, typename F>
__device__ __forceinline__ void pick_for_each(const uint64_t (&reg)[kPickItems], const uint64_t* __restrict__ c, int k,
                                              int tid, F&& f) {
  if (CACHED) {
#pragma unroll
    for (int i = 0; i < kPickItems; ++i) f(reg[i], i * kPickThreads + tid < k);
  } else {
    pick_stream(c, k, tid, f);
  }
}

// Generic walk (any number of ranks, any number of exact ties): all five radix levels, every live candidate compared with
// every rank's prefix at every level.  Runs for the images the fast kernel below cannot finish.  `pos` / `out` are the image's.
template <bool CACHED>
__device__ __noinline__ void pick_ranks_generic(const uint64_t* __restrict__ c, int k, int n, const int32_t* __restrict__ pos,
                                                int32_t* __restrict__ out, uint32_t* sh_h) {
  __shared__ uint64_t sh_prefix[kPickRanks];
  __shared__ uint32_t sh_rem[kPickRanks];
  __shared__ uint32_t sh_cnt[kPickRanks];  // elements still matching rank j's prefix after the current level
  __shared__ uint32_t sh_n[kPickRanks];
  __shared__ uint32_t sh_small;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // the k candidates are read ONCE into registers (all loads in flight together) and reused by every level;
  // larger k (e.g. 5 % of a 1024x2048 image) streams them from L2 at each level instead
  uint64_t reg[kPickItems];
  if (CACHED) {
#pragma unroll
    for (int i = 0; i < kPickItems; ++i) {
      const int idx = i * kPickThreads + tid;
      reg[i] = (idx < k) ? c[idx < k ? idx : 0] : ~0ull;  // ~0 never matches a prefix below level 0's bucket
    }
  }
  for (int j0 = 0; j0 < n; j0 += kPickRanks) {
    const int nr = (n - j0) < kPickRanks ? (n - j0) : kPickRanks;
    uint32_t alive = 0;
#pragma unroll
    for (int i = 0; i < kPickItems; ++i) alive |= (i * kPickThreads + tid < k) ? (1u << i) : 0u;
    if (tid < nr) {
      sh_prefix[tid] = 0;
      int r = pos ? pos[j0 + tid] : (j0 + tid);
      r = r < 0 ? 0 : (r >= k ? k - 1 : r);
      sh_rem[tid] = (uint32_t)r + 1u;  // 1-based rank inside the current bucket
    }
    for (int level = 0; level < kLevels; ++level) {
      const int shift = c_shift[level];
      const int bits = c_bits[level];
      const uint32_t dmask = (1u << bits) - 1u;
      const int nh = (level == 0) ? 1 : nr;  // level 0: every rank shares the empty prefix
      for (int i = tid; i < nh * kHistBins; i += kPickThreads) sh_h[i] = 0;
      __syncthreads();
      // prefix of rank j above the current digit, pre-shifted so the per-element test is one 64-bit compare
      const int hs = shift + bits;
      uint64_t preh[kPickRanks];
#pragma unroll
      for (int j = 0; j < kPickRanks; ++j) preh[j] = (j < nr) ? (sh_prefix[j] >> (hs & 63)) : ~0ull;
      auto visit = [&](uint64_t v) -> bool {
        const uint32_t d = (uint32_t)(v >> shift) & dmask;
        const uint64_t vh = v >> (hs & 63);
        bool any = false;
#pragma unroll
        for (int j = 0; j < kPickRanks; ++j)
          if (vh == preh[j]) {
            atomicAdd(&sh_h[j * kHistBins + d], 1u);
            any = true;
          }
        return any;
      };
      if (CACHED && level == 0) {
        // the selected candidates sit in a narrow score range, so their leading digit takes only a few values:
        // aggregate equal digits inside the warp (match.any) and issue ONE shared-memory atomic per distinct digit
#pragma unroll
        for (int i = 0; i < kPickItems; ++i) {
          const bool valid = i * kPickThreads + tid < k;
          const uint32_t d = valid ? ((uint32_t)(reg[i] >> shift) & dmask) : 0xFFFFFFFFu;
          const uint32_t m = __match_any_sync(0xFFFFFFFFu, d);
          if (valid && lane == __ffs(m) - 1) atomicAdd(&sh_h[d], (uint32_t)__popc(m));
        }
      } else if (CACHED) {
        // an element that matches no rank's prefix at this level can never match again: drop it from later levels
        uint32_t still = 0;
#pragma unroll
        for (int i = 0; i < kPickItems; ++i)
          if ((alive >> i) & 1u) still |= visit(reg[i]) ? (1u << i) : 0u;
        alive = still;
      } else if (level == 0) {
        pick_stream(c, k, tid, [&](uint64_t v, bool valid) {
          if (valid) atomicAdd(&sh_h[(uint32_t)(v >> shift) & dmask], 1u);
        });
      } else {
        pick_stream(c, k, tid, [&](uint64_t v, bool valid) {
          if (valid) visit(v);
        });
      }
      __syncthreads();
      // warp j narrows rank j: lane l owns bins [64 l, 64 l + 64); reads are rotated by the lane id so the 32 lanes
      // hit 32 different banks
      for (int j = warp; j < nr; j += kPickThreads / 32) {
        const uint32_t* h = sh_h + (level == 0 ? 0 : j * kHistBins);
        const uint32_t rem = sh_rem[j];
        uint32_t mine = 0;
        for (int b = 0; b < 64; ++b) mine += h[lane * 64 + ((b + lane) & 63)];
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
          if (lane >= o) incl += t;
        }
        uint32_t run = incl - mine;
        const bool here = run < rem && rem <= incl;
        int found = -1;
        uint32_t before = 0, in_bin = 0;
        if (here) {
          for (int b = 0; b < 64; ++b) {
            const uint32_t hb = h[lane * 64 + b];
            if (run < rem && rem <= run + hb) { found = lane * 64 + b; before = run; in_bin = hb; break; }
            run += hb;
          }
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, found >= 0);
        const int src = __ffs(m) - 1;
        found = __shfl_sync(0xFFFFFFFFu, found, src);
        before = __shfl_sync(0xFFFFFFFFu, before, src);
        in_bin = __shfl_sync(0xFFFFFFFFu, in_bin, src);
        if (lane == 0) {
          sh_prefix[j] |= (uint64_t)(uint32_t)found << shift;
          sh_rem[j] = rem - before;
          sh_cnt[j] = in_bin;
        }
      }
      __syncthreads();
      // After two levels (22 key bits) a rank's group is normally a handful of elements: finish by ranking each group
      // directly in one warp instead of walking three more radix levels (each: clear 96 KB of histograms, a pass over
      // the candidates, a 2048-bin scan per rank).  Falls through to the remaining levels if any group exceeds a warp.
      if (CACHED && level == 1) {
        if (tid == 0) {
          uint32_t mx = 0;
          for (int j = 0; j < nr; ++j) mx = sh_cnt[j] > mx ? sh_cnt[j] : mx;
          sh_small = (mx <= 32u) ? 1u : 0u;
        }
        if (tid < kPickRanks) sh_n[tid] = 0;
        __syncthreads();
        if (sh_small) {
          uint64_t* lists = reinterpret_cast<uint64_t*>(sh_h);  // [kPickRanks][32], the histograms are dead here
          uint64_t pre22[kPickRanks];
#pragma unroll
          for (int j = 0; j < kPickRanks; ++j) pre22[j] = (j < nr) ? (sh_prefix[j] >> 42) : ~0ull;
#pragma unroll
          for (int i = 0; i < kPickItems; ++i) {
            if (!((alive >> i) & 1u)) continue;
            const uint64_t vh = reg[i] >> 42;
#pragma unroll
            for (int j = 0; j < kPickRanks; ++j)
              if (vh == pre22[j]) lists[j * 32 + atomicAdd(&sh_n[j], 1u)] = reg[i];
          }
          __syncthreads();
          for (int j = warp; j < nr; j += kPickThreads / 32) {
            const uint32_t cnt = sh_n[j];
            const uint64_t x = (uint32_t)lane < cnt ? lists[j * 32 + lane] : ~0ull;
            uint32_t below = 0;
            for (int o = 0; o < 32; ++o) {
              const uint64_t y = __shfl_sync(0xFFFFFFFFu, x, o);
              below += (y < x) ? 1u : 0u;
            }
            if ((uint32_t)lane < cnt && below + 1u == sh_rem[j]) sh_prefix[j] = x;  // composites are unique
          }
          __syncthreads();
          break;
        }
      }
    }
    if (tid < nr) out[j0 + tid] = (int32_t)(uint32_t)(sh_prefix[tid] & 0xFFFFFFFFull);
    __syncthreads();
  }
}

// Fast path of the order-statistics pick (n <= kPickRanks ranks, the common case n = 10): THREE cheap passes over the
// k candidates instead of five 12-way compare passes.
//   pass 1  histogram of the leading 11-bit digit (shared by all ranks)            -> per rank: bucket b0, rank inside it
//   pass 2  ranks that share b0 form a group; map0[digit0] -> group (one byte table lookup); candidates of a group are
//           histogrammed by their second digit                                     -> per rank: bucket b1, count, rank
//   pass 3  candidates whose 22-bit prefix equals a rank's are appended to that rank's list (<= 32 entries), one warp
//           ranks each list directly
// A candidate costs a load, two shifts and one table lookup per pass (the generic walk compares every live candidate
// with every rank's 64-bit prefix at every level: ~3000 instructions per thread at k = 6553).  If a 22-bit group holds more
// than 32 candidates the remaining 10 key bits are split off as well (pass 2b); exact ties beyond a warp, or n > kPickRanks,
// are finished by the generic walk inside the same launch.
// Narrowing a rank in a 2048-bin histogram is two 32-wide steps: the sums of the 32 groups of 64 bins are made by ALL
// threads (one 16-byte read + one half-warp redux each), then the rank's warp scans the 32 sums and the 64 bins of the one
// group that holds it (two bins per lane) - instead of every lane walking 64 bins twice (ncu: the kernel is a latency
// chain at 0.97 IPC, and those walks were ~1300 of the ~2700 instructions on it).
template <bool CACHED>
__global__ void __launch_bounds__(kPickThreads, CACHED ? 2 : 1) pick_ranks_fast_kernel(const PickParams p) {
  extern __shared__ __align__(16) uint32_t sh_h[];  // [kPickRanks][2048]: hist0 = sh_h[0..2048), hist1[g] = sh_h[g * 2048 ...)
  __shared__ uint8_t map0[kHistBins];
  __shared__ uint32_t sh_coarse[kPickRanks][32];
  __shared__ uint32_t sh_b0[kPickRanks], sh_b1[kPickRanks], sh_rem[kPickRanks], sh_cnt[kPickRanks], sh_grp[kPickRanks];
  __shared__ uint32_t sh_n[kPickRanks];
  __shared__ uint64_t sh_res[kPickRanks];
  __shared__ uint32_t sh_ng, sh_ok;
  const int img = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t* c = p.cand + (size_t)img * p.kpad;
  const int32_t* pos = p.pos ? p.pos + (size_t)img * p.n : nullptr;
  int32_t* out = p.out + (size_t)img * p.n;
  const int nr = p.n;
  const uint32_t full = 0xFFFFFFFFu;
  if (nr > kPickRanks) {  // more ranks than one pass handles
    pick_ranks_generic<CACHED>(c, p.k, p.n, pos, out, sh_h);
    return;
  }
  uint64_t reg[kPickItems];
  if (CACHED) {
#pragma unroll
    for (int i = 0; i < kPickItems; ++i) {
      const int idx = i * kPickThreads + tid;
      reg[i] = (idx < p.k) ? c[idx < p.k ? idx : 0] : ~0ull;
    }
  }
  const int s0 = c_shift[0], s1 = c_shift[1];
  const uint32_t m0 = (1u << c_bits[0]) - 1u, m1 = (1u << c_bits[1]) - 1u;
  for (int i = tid; i < kHistBins; i += kPickThreads) sh_h[i] = 0;
  if (tid < nr) {
    int r = pos ? pos[tid] : tid;
    r = r < 0 ? 0 : (r >= p.k ? p.k - 1 : r);
    sh_rem[tid] = (uint32_t)r + 1u;
  }
  __syncthreads();
  // sums of the 32 groups of 64 bins of the first nh histograms (all threads; the caller synchronises)
  auto coarse_sums = [&](int nh) {
    const uint32_t half = (lane & 16) ? 0xFFFF0000u : 0x0000FFFFu;
    for (int g = 0; g < nh; ++g) {
      const uint4 h4 = *reinterpret_cast<const uint4*>(sh_h + g * kHistBins + tid * 4);
      const uint32_t s = __reduce_add_sync(half, h4.x + h4.y + h4.z + h4.w);
      if ((lane & 15) == 0) sh_coarse[g][tid >> 4] = s;
    }
  };
  // narrowing: warp j finds the bin of `h` (group sums `cs`) that holds the element of 1-based rank sh_rem[j]
  auto narrow = [&](const uint32_t* h, const uint32_t* cs, int j, uint32_t* bucket_out) {
    uint32_t rem = sh_rem[j];
    uint32_t mine = cs[lane];
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(full, incl, o);
      if (lane >= o) incl += t;
    }
    uint32_t run = incl - mine;
    const int grp = __ffs((int)__ballot_sync(full, run < rem && rem <= incl)) - 1;  // exactly one lane (1 <= rem <= total)
    rem -= __shfl_sync(full, run, grp);
    const uint2 hb = *reinterpret_cast<const uint2*>(h + grp * 64 + lane * 2);
    mine = hb.x + hb.y;
    incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(full, incl, o);
      if (lane >= o) incl += t;
    }
    run = incl - mine;
    const int src = __ffs((int)__ballot_sync(full, run < rem && rem <= incl)) - 1;
    const bool second = rem > run + hb.x;  // meaningful on lane src
    const uint32_t found = __shfl_sync(full, (uint32_t)(grp * 64 + lane * 2) + (second ? 1u : 0u), src);
    const uint32_t before = __shfl_sync(full, run + (second ? hb.x : 0u), src);
    const uint32_t in_bin = __shfl_sync(full, second ? hb.y : hb.x, src);
    if (lane == 0) {
      bucket_out[j] = found;
      sh_rem[j] = rem - before;
      sh_cnt[j] = in_bin;
    }
  };
  // ---- pass 1: leading digit (few distinct values: aggregate equal digits inside the warp) ----
  pick_for_each<CACHED>(reg, c, p.k, tid, [&](uint64_t v, bool valid) {
    const uint32_t d = valid ? ((uint32_t)(v >> s0) & m0) : 0xFFFFFFFFu;
    const uint32_t m = __match_any_sync(full, d);
    if (valid && lane == __ffs((int)m) - 1) atomicAdd(&sh_h[d], (uint32_t)__popc(m));
  });
  __syncthreads();
  coarse_sums(1);
  for (int i = tid; i < kHistBins; i += kPickThreads) map0[i] = 255;
  __syncthreads();
  for (int j = warp; j < nr; j += kPickThreads / 32) narrow(sh_h, sh_coarse[0], j, sh_b0);
  __syncthreads();
  // ---- groups of ranks sharing the leading bucket ----
  if (tid == 0) {
    uint32_t ng = 0;
    for (int j = 0; j < nr; ++j) {
      const uint32_t b = sh_b0[j];
      if (map0[b] == 255) map0[b] = (uint8_t)ng++;
      sh_grp[j] = map0[b];
    }
    sh_ng = ng;
  }
  __syncthreads();
  const uint32_t ng = sh_ng;
  for (int i = tid; i < (int)ng * (kHistBins / 4); i += kPickThreads)  // hist0 is dead: reuse as hist1[g]
    reinterpret_cast<uint4*>(sh_h)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid < kPickRanks) sh_n[tid] = 0;
  __syncthreads();
  // ---- pass 2: second digit of the candidates that fall in a rank's leading bucket ----
  pick_for_each<CACHED>(reg, c, p.k, tid, [&](uint64_t v, bool valid) {
    if (!valid) return;
    const uint32_t g = map0[(uint32_t)(v >> s0) & m0];
    if (g != 255u) atomicAdd(&sh_h[g * kHistBins + ((uint32_t)(v >> s1) & m1)], 1u);
  });
  __syncthreads();
  coarse_sums((int)ng);
  __syncthreads();
  for (int j = warp; j < nr; j += kPickThreads / 32) narrow(sh_h + sh_grp[j] * kHistBins, sh_coarse[sh_grp[j]], j, sh_b1);
  __syncthreads();
  // ---- unique 22-bit prefixes of the ranks (g2), and are all their groups at most a warp? ----
  __shared__ uint32_t gp2[kPickRanks], sh_g2[kPickRanks], sh_b2[kPickRanks], sh_lid[kPickRanks], sh_pre[kPickRanks], sh_ng2;
  if (tid == 0) {
    uint32_t mx = 0, n2 = 0;
    for (int j = 0; j < nr; ++j) {
      mx = sh_cnt[j] > mx ? sh_cnt[j] : mx;
      const uint32_t pj = (sh_b0[j] << c_bits[1]) | sh_b1[j];
      uint32_t g = 0;
      while (g < n2 && gp2[g] != pj) ++g;
      if (g == n2) gp2[n2++] = pj;
      sh_g2[j] = g;
    }
    sh_ng2 = n2;
    sh_ok = (mx <= 32u) ? 1u : 0u;
  }
  __syncthreads();
  const uint32_t ng2 = sh_ng2;
  // map1[group][second digit] -> g2 (byte table at the END of the histogram area; hist1 is dead).  With it a candidate's
  // prefix is resolved by two table lookups in the passes below.  (ncu had this kernel waiting for instruction fetch on 45 %
  // of its issue cycles - every CTA runs its code once - while it compared each candidate with the 12 ranks' prefixes in
  // 16 x 12 unrolled blocks: 38 -> 24 us for 256 images with the tables.)
  constexpr int kHistBytes = kPickRanks * kHistBins * (int)sizeof(uint32_t);
  uint8_t* map1 = reinterpret_cast<uint8_t*>(sh_h) + kHistBytes - (int)ng * kHistBins;
  for (int i = tid; i < (int)ng * (kHistBins / 16); i += kPickThreads)
    reinterpret_cast<uint4*>(map1)[i] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
  __syncthreads();
  if (tid < nr) map1[sh_grp[tid] * kHistBins + sh_b1[tid]] = (uint8_t)sh_g2[tid];  // equal prefixes write equal values
  __syncthreads();
  uint64_t* lists = reinterpret_cast<uint64_t*>(sh_h);  // [<= kPickRanks][32] (3 KB), filled in pass 3
  if (sh_ok) {
    // ---- pass 3: gather the candidates of each unique prefix (<= 32) ----
    if (tid < nr) sh_lid[tid] = sh_g2[tid];
    pick_for_each<CACHED>(reg, c, p.k, tid, [&](uint64_t v, bool valid) {
      if (!valid) return;
      const uint32_t g = map0[(uint32_t)(v >> s0) & m0];
      if (g == 255u) return;
      const uint32_t id = map1[g * kHistBins + ((uint32_t)(v >> s1) & m1)];
      if (id != 255u) lists[id * 32 + atomicAdd(&sh_n[id], 1u)] = v;
    });
  } else {
    // ---- pass 2b: a 22-bit group is still larger than a warp (scores packed into a narrow range, e.g. the top 5 % of
    // entropies of a 1024x2048 image): split it by the remaining 10 key bits -> the whole float is resolved ----
    const int s2 = c_shift[2];
    const uint32_t m2 = (1u << c_bits[2]) - 1u;
    const bool fit = (int)ng2 * kHistBins * (int)sizeof(uint32_t) + (int)ng * kHistBins <= kHistBytes;  // hist2 below map1
    for (int i = tid; i < (int)ng2 * (kHistBins / 4); i += kPickThreads)
      reinterpret_cast<uint4*>(sh_h)[i] = make_uint4(0u, 0u, 0u, 0u);  // (overwrites map1 when it does not fit)
    __syncthreads();
    if (fit) {
      pick_for_each<CACHED>(reg, c, p.k, tid, [&](uint64_t v, bool valid) {
        if (!valid) return;
        const uint32_t g = map0[(uint32_t)(v >> s0) & m0];
        if (g == 255u) return;
        const uint32_t g2 = map1[g * kHistBins + ((uint32_t)(v >> s1) & m1)];
        if (g2 != 255u) atomicAdd(&sh_h[g2 * kHistBins + ((uint32_t)(v >> s2) & m2)], 1u);
      });
    } else {  // the table was overwritten by the histograms (> 9 distinct large groups): compare with the prefixes instead
      pick_for_each<CACHED>(reg, c, p.k, tid, [&](uint64_t v, bool valid) {
        if (!valid || map0[(uint32_t)(v >> s0) & m0] == 255u) return;
        const uint32_t key = (uint32_t)(v >> s1);
#pragma unroll 1
        for (uint32_t g = 0; g < ng2; ++g)
          if (key == gp2[g]) {
            atomicAdd(&sh_h[g * kHistBins + ((uint32_t)(v >> s2) & m2)], 1u);
            break;
          }
      });
    }
    __syncthreads();
    coarse_sums((int)ng2);
    __syncthreads();
    for (int j = warp; j < nr; j += kPickThreads / 32) narrow(sh_h + sh_g2[j] * kHistBins, sh_coarse[sh_g2[j]], j, sh_b2);
    __syncthreads();
    if (tid == 0) {
      uint32_t mx = 0;
      for (int j = 0; j < nr; ++j) mx = sh_cnt[j] > mx ? sh_cnt[j] : mx;
      sh_ok = (mx <= 32u) ? 1u : 0u;
    }
    __syncthreads();
    if (!sh_ok) {  // exact ties beyond a warp: the generic walk finishes this image (every thread takes this branch)
      pick_ranks_generic<CACHED>(c, p.k, p.n, pos, out, sh_h);
      return;
    }
    // ---- pass 3 on all 32 key bits ----
    if (fit) {
      // map2[g2][third digit] -> list id (the first rank with that prefix), at byte 4096 of the (dead) histogram area
      uint8_t* map2 = reinterpret_cast<uint8_t*>(sh_h) + 4096;
      const int d2n = 1 << c_bits[2];
      for (int i = tid; i < (int)ng2 * (d2n / 16); i += kPickThreads)
        reinterpret_cast<uint4*>(map2)[i] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
      __syncthreads();
      if (tid == 0) {
        for (int j = 0; j < nr; ++j) {
          uint8_t* e = map2 + sh_g2[j] * d2n + sh_b2[j];
          if (*e == 255) *e = (uint8_t)j;
          sh_lid[j] = *e;
        }
      }
      __syncthreads();
      pick_for_each<CACHED>(reg, c, p.k, tid, [&](uint64_t v, bool valid) {
        if (!valid) return;
        const uint32_t g = map0[(uint32_t)(v >> s0) & m0];
        if (g == 255u) return;
        const uint32_t g2 = map1[g * kHistBins + ((uint32_t)(v >> s1) & m1)];
        if (g2 == 255u) return;
        const uint32_t id = map2[g2 * d2n + ((uint32_t)(v >> s2) & m2)];
        if (id != 255u) lists[id * 32 + atomicAdd(&sh_n[id], 1u)] = v;
      });
    } else {
      if (tid < nr) {
        sh_lid[tid] = (uint32_t)tid;
        sh_pre[tid] = (sh_b0[tid] << (c_bits[1] + c_bits[2])) | (sh_b1[tid] << c_bits[2]) | sh_b2[tid];  // all 32 key bits
      }
      __syncthreads();
      pick_for_each<CACHED>(reg, c, p.k, tid, [&](uint64_t v, bool valid) {
        if (!valid || map0[(uint32_t)(v >> s0) & m0] == 255u) return;
        const uint32_t key = (uint32_t)(v >> s2);
#pragma unroll 1
        for (int j = 0; j < nr; ++j)
          if (key == sh_pre[j]) lists[j * 32 + atomicAdd(&sh_n[j], 1u)] = v;
      });
    }
  }
  __syncthreads();
  for (int j = warp; j < nr; j += kPickThreads / 32) {
    const uint32_t id = sh_lid[j];
    const uint32_t cnt = sh_n[id];
    const uint64_t x = (uint32_t)lane < cnt ? lists[id * 32 + lane] : ~0ull;
    uint32_t below = 0;
    for (int o = 0; o < 32; ++o) {
      const uint64_t y = __shfl_sync(full, x, o);
      below += (y < x) ? 1u : 0u;
    }
    if ((uint32_t)lane < cnt && below + 1u == sh_rem[j]) sh_res[j] = x;  // composites are unique
  }
  __syncthreads();
  if (tid < nr) out[tid] = (int32_t)(uint32_t)(sh_res[tid] & 0xFFFFFFFFull);
}

// ------------------------------------------------------------------------------------------
// bitonic sort of the selected composites
// ------------------------------------------------------------------------------------------
constexpr int kSortThreads = 1024;

struct SortParams {
  uint64_t* cand;  // [n_img][kpad]
  int kpad, k, chunk;
  int size_lo;   // first bitonic size handled by this launch (2 => full local sort)
  int size_hi;   // last bitonic size handled locally (== chunk for FULL, == size for MERGE-FINISH)
  int pad_on_load;
  int write_out;
  int largest;
  int32_t* out_idx;  // [n_img][k]
  float* out_val;    // optional
};

__device__ __forceinline__ void cmpx(uint64_t& a, uint64_t& b, bool asc) {
  if ((a > b) == asc) {
    const uint64_t t = a;
    a = b;
    b = t;
  }
}

// One CTA sorts/merges one chunk in shared memory. For size <= chunk the whole network lives in
// the chunk; for size > chunk only strides < chunk are done here (global strides run before).
__global__ void __launch_bounds__(kSortThreads) bitonic_local_kernel(const SortParams p) {
  extern __shared__ uint64_t sh[];
  const int img = blockIdx.y;
  const int cbase = blockIdx.x * p.chunk;
  uint64_t* g = p.cand + (size_t)img * p.kpad + cbase;
  for (int i = threadIdx.x; i < p.chunk; i += kSortThreads) {
    uint64_t v = g[i];
    if (p.pad_on_load && cbase + i >= p.k) v = ~0ull;
    sh[i] = v;
  }
  __syncthreads();
  for (int size = p.size_lo; size <= p.size_hi; size <<= 1) {
    int stride = size >> 1;
    if (stride >= p.chunk) stride = p.chunk >> 1;
    for (; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (p.chunk >> 1); t += kSortThreads) {
        const int pos = 2 * t - (t & (stride - 1));
        const bool asc = ((cbase + pos) & size) == 0;
        uint64_t a = sh[pos], b = sh[pos + stride];
        cmpx(a, b, asc);
        sh[pos] = a;
        sh[pos + stride] = b;
      }
      __syncthreads();
    }
  }
  if (p.write_out) {
    for (int i = threadIdx.x; i < p.chunk; i += kSortThreads) {
      const int j = cbase + i;
      if (j < p.k) {
        const uint64_t v = sh[i];
        p.out_idx[(size_t)img * p.k + j] = (int32_t)(uint32_t)(v & 0xFFFFFFFFull);
        if (p.out_val) p.out_val[(size_t)img * p.k + j] = ord_key_inv((uint32_t)(v >> 32), p.largest != 0);
      }
    }
  } else {
    for (int i = threadIdx.x; i < p.chunk; i += kSortThreads) g[i] = sh[i];
  }
}

// Register-blocked bitonic network for 8192-element chunks (1024 threads x 8 elements).  Element bit b of a
// compare-exchange decides where it runs: bits 0-2 inside the thread's registers, bits 3-7 across lanes with warp
// shuffles, bits 8-12 through shared memory in groups of up to three bits per round trip (12 shared-memory passes
// for a full sort instead of 91).  Shared-memory position of element e is e ^ (((e >> 4) & 3) << 1), which makes
// both the per-thread 64-byte reads and the strided round accesses bank-conflict free.
constexpr int kB8 = 8192;
__device__ __forceinline__ int sw8(int e) { return e ^ (((e >> 4) & 3) << 1); }
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int mask) {
  const uint32_t lo = __shfl_xor_sync(0xFFFFFFFFu, (uint32_t)v, mask);
  const uint32_t hi = __shfl_xor_sync(0xFFFFFFFFu, (uint32_t)(v >> 32), mask);
  return ((uint64_t)hi << 32) | lo;
}

template <bool FULL>
__global__ void __launch_bounds__(1024) bitonic8k_kernel(const SortParams p) {
  extern __shared__ uint64_t sh[];
  const int img = blockIdx.y;
  const int t = threadIdx.x, lane = t & 31;
  const int cbase = blockIdx.x * kB8;
  uint64_t* g = p.cand + (size_t)img * p.kpad + cbase;
  uint64_t a[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    uint64_t v = g[8 * t + r];
    if (p.pad_on_load && cbase + 8 * t + r >= p.k) v = ~0ull;
    a[r] = v;
  }
  auto smem_phase = [&](int hi, int size) {
#pragma unroll
    for (int r = 0; r < 8; ++r) sh[sw8(8 * t + r)] = a[r];
    __syncthreads();
    int b = hi;
    while (b >= 8) {
      const int gb = (b - 7) < 3 ? (b - 7) : 3;
      const int b0 = b - gb + 1;
      const int n = 1 << gb;
      for (int k = 0; k < (8 >> gb); ++k) {
        const int v = t + 1024 * k;
        const int base = (v & ((1 << b0) - 1)) | ((v >> b0) << (b0 + gb));
        const bool asc = ((cbase + base) & size) == 0;
        uint64_t x[8];
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < n) x[c] = sh[sw8(base | (c << b0))];
#pragma unroll
        for (int j = 2; j >= 0; --j) {
          if (j < gb) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (c < n && !(c & (1 << j))) cmpx(x[c], x[c | (1 << j)], asc);
          }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < n) sh[sw8(base | (c << b0))] = x[c];
      }
      __syncthreads();
      b -= gb;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) a[r] = sh[sw8(8 * t + r)];
    __syncthreads();
  };
  auto merge = [&](int s, int size) {
    int hi = s - 1;
    if (hi >= 8) {
      smem_phase(hi > 12 ? 12 : hi, size);
      hi = 7;
    }
    for (int b = hi; b >= 3; --b) {
      const int mask = 1 << (b - 3);
      const bool lower = (lane & mask) == 0;
      const bool asc = ((cbase + 8 * t) & size) == 0;
      const bool keep_min = lower == asc;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const uint64_t o = shfl_xor_u64(a[r], mask);
        a[r] = keep_min ? (a[r] < o ? a[r] : o) : (a[r] > o ? a[r] : o);
      }
    }
    const int top = hi < 2 ? hi : 2;
#pragma unroll
    for (int b = 2; b >= 0; --b) {
      if (b <= top) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
          if (!(r & (1 << b))) cmpx(a[r], a[r | (1 << b)], ((cbase + 8 * t + r) & size) == 0);
      }
    }
  };
  if (FULL) {
    for (int s = 1; s <= 13; ++s) merge(s, 1 << s);
  } else {
    merge(13, p.size_lo);
  }
  if (p.write_out) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int j = cbase + 8 * t + r;
      if (j < p.k) {
        p.out_idx[(size_t)img * p.k + j] = (int32_t)(uint32_t)(a[r] & 0xFFFFFFFFull);
        if (p.out_val) p.out_val[(size_t)img * p.k + j] = ord_key_inv((uint32_t)(a[r] >> 32), p.largest != 0);
      }
    }
  } else {
#pragma unroll
    for (int r = 0; r < 8; ++r) g[8 * t + r] = a[r];
  }
}

__global__ void __launch_bounds__(256) bitonic_global_step_kernel(uint64_t* cand, int kpad, int size,
                                                                  int stride) {
  const int img = blockIdx.y;
  uint64_t* g = cand + (size_t)img * kpad;
  for (int t = blockIdx.x * 256 + threadIdx.x; t < (kpad >> 1); t += gridDim.x * 256) {
    const int pos = 2 * t - (t & (stride - 1));
    const bool asc = (pos & size) == 0;
    uint64_t a = g[pos], b = g[pos + stride];
    if ((a > b) == asc) {
      g[pos] = b;
      g[pos + stride] = a;
    }
  }
}

__global__ void gather_kernel(const int32_t* __restrict__ topk_idx, int k, const int32_t* __restrict__ pos,
                              int n, int32_t* __restrict__ out, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int img = i / n;
  int ps = pos ? pos[i] : (i - img * n);
  ps = ps < 0 ? 0 : (ps >= k ? k - 1 : ps);  // ranks outside [0, k) are clamped like in the pick kernels, never read out of bounds
  out[i] = topk_idx[(size_t)img * k + ps];
}

template <typename T>
__global__ void entropy_at_kernel(const T* __restrict__ logits, int C, int W, int HW, int64_t sn, int64_t sc,
                                  int64_t sh, const int32_t* __restrict__ px, int n, float* __restrict__ out,
                                  int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int img = i / n;
  int idx = px[i];
  idx = idx < 0 ? 0 : (idx >= HW ? HW - 1 : idx);  // a pixel index outside the image is clamped, never read out of bounds
  const int y = idx / W, x = idx - y * W;
  const T* src = logits + (int64_t)img * sn + (int64_t)y * sh + x;
  out[i] = score_runtime_c<PP_STRAT_ENTROPY>(C, [&](int c) { return Ld4<T>::ld1(src + (int64_t)c * sc); });
}

__global__ void entropy_at_up_kernel(const float* __restrict__ logits, int C, int h_in, int w_in, int H, int W,
                                     float scale_h, float scale_w, const int32_t* __restrict__ px, int n,
                                     float* __restrict__ out, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int img = i / n;
  int idx = px[i];
  idx = idx < 0 ? 0 : (idx >= H * W ? H * W - 1 : idx);
  const int y = idx / W, x = idx - y * W;
  const Lerp ly = lerp_ac(y, h_in, H, scale_h);
  const Lerp lx = lerp_ac(x, w_in, W, scale_w);
  const int64_t plane = (int64_t)h_in * w_in;
  const float* b = logits + (int64_t)img * C * plane;
  const int64_t o00 = (int64_t)ly.i0 * w_in + lx.i0, o01 = (int64_t)ly.i0 * w_in + lx.i1;
  const int64_t o10 = (int64_t)ly.i1 * w_in + lx.i0, o11 = (int64_t)ly.i1 * w_in + lx.i1;
  out[i] = score_runtime_c<PP_STRAT_ENTROPY>(C, [&](int c) {
    const float* pc = b + c * plane;
    return ly.l0 * (lx.l0 * pc[o00] + lx.l1 * pc[o01]) + ly.l1 * (lx.l0 * pc[o10] + lx.l1 * pc[o11]);
  });
}

// ------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------
// Tuning knob (bench/experiments only): PP_SCORE_VARIANT = 0 (4 px, 2 CTAs/SM: default), 1 (4 px, 3 CTAs/SM),
// 2 (2 px, 4 CTAs/SM), 3 (2 px, 6 CTAs/SM).
static int score_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PP_SCORE_VARIANT");
    v = e ? atoi(e) : 0;
    if (v < 0 || v > 6) v = 0;
  }
  return v;
}

// PP_SELECT_L0=1: the one-chunk-per-CTA level-0 kernel instead of the staged persistent one (A/B measurements)
static int select_l0_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PP_SELECT_L0");
    v = e ? atoi(e) : 0;
    if (v < 0 || v > 3) v = 0;
  }
  return v;
}

template <int C, int STRAT, typename T, int PX, int MINB>
static void launch_score_vec_v(const ScoreParams& p, cudaStream_t st) {
  constexpr int ITERS = 4;
  const int nvec = p.H * (p.W / PX);
  dim3 grid((nvec + kScoreThreads * ITERS - 1) / (kScoreThreads * ITERS), p.n_img);
  if (p.hist0) acq_score_vec_kernel<C, STRAT, T, true, ITERS, PX, MINB><<<grid, kScoreThreads, 0, st>>>(p);
  else acq_score_vec_kernel<C, STRAT, T, false, ITERS, PX, MINB><<<grid, kScoreThreads, 0, st>>>(p);
}

template <int C, int STRAT, int ITERS>
static void launch_score_pf(const ScoreParams& p, cudaStream_t st) {
  constexpr int smem = C * kScoreThreads * 16;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(acq_score_pf_kernel<C, STRAT, true, ITERS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(acq_score_pf_kernel<C, STRAT, false, ITERS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr = true;
  }
  const int nvec = p.H * (p.W / 4);
  dim3 grid((nvec + kScoreThreads * ITERS - 1) / (kScoreThreads * ITERS), p.n_img);
  if (p.hist0) acq_score_pf_kernel<C, STRAT, true, ITERS, 2><<<grid, kScoreThreads, smem, st>>>(p);
  else acq_score_pf_kernel<C, STRAT, false, ITERS, 2><<<grid, kScoreThreads, smem, st>>>(p);
}

template <int C, int STRAT, typename T>
static void launch_score_vec(const ScoreParams& p, cudaStream_t st) {
  // fp32 logits, enough CTAs for >= 8 waves: the kernel that keeps the next tile in flight while it scores (256 images of
  // 256x512: 415 -> 402 us = 6.51 TB/s of algorithmic bytes; 32 images of 360x480: 73 -> 79 us, hence the size rule).
  // PP_SCORE_VARIANT=4 / 5 force it (4 / 8 tiles per thread), 6 forces the plain kernel.
  const int sv = score_variant();
  const long long ctas = (long long)((p.H * (p.W / 4) + kScoreThreads * 4 - 1) / (kScoreThreads * 4)) * p.n_img;
  if (sizeof(T) == 4 && p.W % 4 == 0 && (sv == 4 || sv == 5 || (sv == 0 && ctas >= 2368))) {
    if (sv == 5) launch_score_pf<C, STRAT, 8>(p, st);
    else launch_score_pf<C, STRAT, 4>(p, st);
    return;
  }
  switch (score_variant()) {
    case 1: launch_score_vec_v<C, STRAT, T, 4, 3>(p, st); break;
    case 2: launch_score_vec_v<C, STRAT, T, 2, 4>(p, st); break;
    case 3: launch_score_vec_v<C, STRAT, T, 2, 6>(p, st); break;
    default: launch_score_vec_v<C, STRAT, T, 4, 2>(p, st); break;
  }
}

template <int STRAT, typename T>
static bool dispatch_c(const ScoreParams& p, cudaStream_t st) {
  switch (p.C) {
    case 11: launch_score_vec<11, STRAT, T>(p, st); return true;
    case 19: launch_score_vec<19, STRAT, T>(p, st); return true;
    case 21: launch_score_vec<21, STRAT, T>(p, st); return true;
    default: return false;
  }
}

template <int STRAT, typename T>
static void launch_score_scalar(const ScoreParams& p, cudaStream_t st) {
  const int64_t HW = (int64_t)p.H * p.W;
  int gx = (int)((HW + kScoreThreads * 4 - 1) / (kScoreThreads * 4));
  if (gx < 1) gx = 1;
  dim3 grid(gx, p.n_img);
  acq_score_scalar_kernel<STRAT, T><<<grid, kScoreThreads, 0, st>>>(p);
}

template <typename T>
static int score_dispatch(const ScoreParams& p, int strategy, bool vec_ok, cudaStream_t st) {
  bool done = false;
  if (vec_ok) {
    if (strategy == PP_STRAT_ENTROPY) done = dispatch_c<PP_STRAT_ENTROPY, T>(p, st);
    else if (strategy == PP_STRAT_LEAST_CONFIDENCE) done = dispatch_c<PP_STRAT_LEAST_CONFIDENCE, T>(p, st);
    else done = dispatch_c<PP_STRAT_MARGIN, T>(p, st);
  }
  if (!done) {
    if (strategy == PP_STRAT_ENTROPY) launch_score_scalar<PP_STRAT_ENTROPY, T>(p, st);
    else if (strategy == PP_STRAT_LEAST_CONFIDENCE) launch_score_scalar<PP_STRAT_LEAST_CONFIDENCE, T>(p, st);
    else launch_score_scalar<PP_STRAT_MARGIN, T>(p, st);
  }
  PP_LAUNCH_CHECK();
  return PP_OK;
}

template <int STRAT>
static bool dispatch_up(const ScoreUpParams& p, dim3 grid, cudaStream_t st) {
  switch (p.C) {
    case 11: acq_score_up_kernel<11, STRAT><<<grid, kScoreThreads, 0, st>>>(p); return true;
    case 19: acq_score_up_kernel<19, STRAT><<<grid, kScoreThreads, 0, st>>>(p); return true;
    case 21: acq_score_up_kernel<21, STRAT><<<grid, kScoreThreads, 0, st>>>(p); return true;
    default: return false;
  }
}

static int select_impl(const float* score_map, int n_img, int HW, int k, int largest, int hist0_valid,
                       void* workspace, size_t workspace_bytes, cudaStream_t st);
static int sort_impl(const Workspace& w, int n_img, int k, int largest, int32_t* topk_idx, float* topk_val,
                     cudaStream_t st);

static int topk_impl(const float* score_map, int n_img, int HW, int k, int largest, int hist0_valid,
                     int32_t* topk_idx, float* topk_val, void* workspace, size_t workspace_bytes,
                     cudaStream_t st) {
  int rc = select_impl(score_map, n_img, HW, k, largest, hist0_valid, workspace, workspace_bytes, st);
  if (rc != PP_OK) return rc;
  Workspace w = carve(workspace, n_img, HW, k);
  return sort_impl(w, n_img, k, largest, topk_idx, topk_val, st);
}

// phase 1: leaves exactly k unsorted composites per image in the workspace candidate list
static int select_impl(const float* score_map, int n_img, int HW, int k, int largest, int hist0_valid,
                       void* workspace, size_t workspace_bytes, cudaStream_t st) {
  Workspace w = carve(workspace, n_img, HW, k);
  if (workspace_bytes < w.total_bytes) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, w.total_bytes);
    return PP_ERR_WORKSPACE;
  }
  const int tiles = (HW + kSelTile - 1) / kSelTile;
  if (!hist0_valid) {
    int gx = tiles < 64 ? tiles : 64;
    hist0_kernel<<<dim3(gx, n_img), kSelThreads, 0, st>>>(score_map, w.hist, HW, largest);
    PP_LAUNCH_CHECK();
  }
  {
    SelParams p;
    p.scores = score_map;
    p.in_list = nullptr;
    p.in_count = nullptr;
    p.out_list = w.filt;  // boundary bucket of level 0
    p.out_count = w.filt_count;
    p.cand = w.cand;
    p.cand_count = w.cand_count;
    p.state_cur = w.state;
    p.state_next = w.state + (size_t)n_img;
    p.level = 0;
    p.n_img = n_img;
    p.HW = HW;
    p.k = k;
    p.kpad = w.kpad;
    p.largest = largest;
    pick_bucket0_kernel<<<n_img, kSelThreads, 0, st>>>(w.hist, w.state + (size_t)n_img, (uint32_t)k, largest != 0);
    PP_LAUNCH_CHECK();
    const int gx = (HW + kL0Chunk - 1) / kL0Chunk;  // one chunk of scores per CTA
    const bool aligned = (reinterpret_cast<uintptr_t>(score_map) & 15) == 0;
    const bool full = (HW % kL0Chunk == 0) && aligned;
    const int l0v = select_l0_variant();
    if (aligned && HW % kL0WarpChunk == 0 && l0v != 1) {
      static int n_sm = 0;
      if (n_sm == 0) {
        int dev = 0, sms = 0;
        PP_CUDA(cudaGetDevice(&dev));
        PP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        PP_CUDA(cudaFuncSetAttribute(select_l0_staged_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, L0sCfg<4>::kSmemBytes));
        PP_CUDA(cudaFuncSetAttribute(select_l0_staged_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, L0sCfg<3>::kSmemBytes));
        PP_CUDA(cudaFuncSetAttribute(select_l0_staged_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, L0sCfg<2>::kSmemBytes));
        n_sm = sms;
      }
      const int stages = l0v == 2 ? 3 : (l0v == 3 ? 4 : 2);  // default: 2 stages, 5 CTAs / SM (measured 48.1 / 50.2 / 54.2 us select phase)
      const int per_sm = stages == 4 ? L0sCfg<4>::kCtasPerSm : (stages == 3 ? L0sCfg<3>::kCtasPerSm : L0sCfg<2>::kCtasPerSm);
      L0sWalk wk;
      wk.chunks_per_img = (uint32_t)(HW / kL0WarpChunk);
      wk.total_chunks = (uint32_t)n_img * wk.chunks_per_img;
      uint32_t ctas = (wk.total_chunks + kL0sWarps - 1) / kL0sWarps;
      const uint32_t cap = (uint32_t)(n_sm * per_sm);
      ctas = ctas < cap ? ctas : cap;
      wk.base = wk.total_chunks / (ctas * kL0sWarps);
      wk.extra = wk.total_chunks % (ctas * kL0sWarps);
      if (stages == 4) select_l0_staged_kernel<4><<<ctas, kSelThreads, L0sCfg<4>::kSmemBytes, st>>>(p, wk);
      else if (stages == 3) select_l0_staged_kernel<3><<<ctas, kSelThreads, L0sCfg<3>::kSmemBytes, st>>>(p, wk);
      else select_l0_staged_kernel<2><<<ctas, kSelThreads, L0sCfg<2>::kSmemBytes, st>>>(p, wk);
    } else if (full) {
      select_l0_kernel<true><<<dim3(gx, n_img), kSelThreads, 0, st>>>(p);
    } else {
      select_l0_kernel<false><<<dim3(gx, n_img), kSelThreads, 0, st>>>(p);
    }
    PP_LAUNCH_CHECK();
    RestParams r;
    r.list_a = w.filt;
    r.list_b = w.filt + (size_t)n_img * HW;
    r.count_a = w.filt_count;
    r.hist0 = w.hist;
    r.cand = w.cand;
    r.cand_count = w.cand_count;
    r.state1 = w.state + (size_t)n_img;
    r.HW = HW;
    r.kpad = w.kpad;
    r.first_level = largest ? 0 : 1;
    select_rest_kernel<<<n_img, kRestThreads, 0, st>>>(r);
    PP_LAUNCH_CHECK();
  }
  return PP_OK;
}

// phase 2a: full sort of the k composites
static int sort_impl(const Workspace& w, int n_img, int k, int largest, int32_t* topk_idx, float* topk_val,
                     cudaStream_t st) {
  SortParams sp;
  sp.cand = w.cand;
  sp.kpad = w.kpad;
  sp.k = k;
  sp.largest = largest;
  sp.out_idx = topk_idx;
  sp.out_val = topk_val;
  if (w.kpad >= kB8) {
    static bool attr8 = false;
    if (!attr8) {
      PP_CUDA(cudaFuncSetAttribute(bitonic8k_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kB8 * (int)sizeof(uint64_t)));
      PP_CUDA(cudaFuncSetAttribute(bitonic8k_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kB8 * (int)sizeof(uint64_t)));
      attr8 = true;
    }
    const size_t smem8 = (size_t)kB8 * sizeof(uint64_t);
    sp.chunk = kB8;
    sp.size_lo = 2;
    sp.size_hi = kB8;
    sp.pad_on_load = 1;
    sp.write_out = (w.kpad == kB8) ? 1 : 0;
    bitonic8k_kernel<true><<<dim3(w.kpad / kB8, n_img), 1024, smem8, st>>>(sp);
    PP_LAUNCH_CHECK();
    for (int size = kB8 << 1; size <= w.kpad; size <<= 1) {
      for (int stride = size >> 1; stride >= kB8; stride >>= 1) {
        int gx = (w.kpad / 2 + 255) / 256;
        if (gx > 1024) gx = 1024;
        bitonic_global_step_kernel<<<dim3(gx, n_img), 256, 0, st>>>(w.cand, w.kpad, size, stride);
        PP_LAUNCH_CHECK();
      }
      sp.size_lo = size;
      sp.size_hi = size;
      sp.pad_on_load = 0;
      sp.write_out = (size == w.kpad) ? 1 : 0;
      bitonic8k_kernel<false><<<dim3(w.kpad / kB8, n_img), 1024, smem8, st>>>(sp);
      PP_LAUNCH_CHECK();
    }
    return PP_OK;
  }
  const int chunk = w.kpad;  // < 8192: one shared-memory chunk
  sp.chunk = chunk;
  const size_t smem = (size_t)chunk * sizeof(uint64_t);
  sp.size_lo = 2;
  sp.size_hi = chunk;
  sp.pad_on_load = 1;
  sp.write_out = 1;
  bitonic_local_kernel<<<dim3(1, n_img), kSortThreads, smem, st>>>(sp);
  PP_LAUNCH_CHECK();
  return PP_OK;
}


// ---- QueryStats at the picks (query.py:250-308): labels, label histogram, unique labels, spatial spread, wire coordinates ----
// np.mean of a float64 array = pairwise summation (numpy/core/src/umath/loops_utils.h, pairwise_sum): restated so that the
// spatial coverage equals the reference's bit for bit.  Element e of the flattened (n, n-1) off-diagonal distance matrix is
// (i, j) = (e / (n-1), e % (n-1) skipping the diagonal); distances are sqrt of exact integers in double (IEEE sqrt).
struct PickDist {
  const long long* idx;  // n sorted flat indices of one image
  int n, W;
  __device__ double at(int e) const {
    const int i = e / (n - 1);
    int j = e - i * (n - 1);
    j += (j >= i);
    const long long a = idx[i], b = idx[j];
    const long long dy = a / W - b / W, dx = a % W - b % W;
    return sqrt((double)(dy * dy + dx * dx));
  }
};
__device__ double pairwise_sum_np(const PickDist& d, int lo, int n) {
  if (n < 8) {
    double res = 0.;
    for (int i = 0; i < n; ++i) res += d.at(lo + i);
    return res;
  }
  if (n <= 128) {
    double r[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) r[q] = d.at(lo + q);
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int q = 0; q < 8; ++q) r[q] += d.at(lo + i + q);
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += d.at(lo + i);
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return pairwise_sum_np(d, lo, n2) + pairwise_sum_np(d, lo + n2, n - n2);
}

__global__ void __launch_bounds__(128) query_stats_kernel(const long long* __restrict__ sel, int n_img, int n, int W, int HW,
                                                          const uint8_t* __restrict__ labels, int n_classes,
                                                          long long* __restrict__ x_coords, long long* __restrict__ y_coords,
                                                          int32_t* __restrict__ labels_at, unsigned long long* __restrict__ label_hist,
                                                          int32_t* __restrict__ n_unique, double* __restrict__ coverage) {
  const int img = blockIdx.x;
  const long long* s = sel + (size_t)img * n;
  __shared__ int sh_lab[1024];
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const long long idx = s[j];
    x_coords[(size_t)img * n + j] = idx % W;  // np.where order: the picks are sorted row-major (query.py:77)
    y_coords[(size_t)img * n + j] = idx / W;
    int lab = -1;
    if (labels) {
      lab = (idx >= 0 && idx < HW) ? (int)labels[(size_t)img * HW + idx] : -1;
      labels_at[(size_t)img * n + j] = lab;
      if (lab >= 0 && lab < n_classes) atomicAdd(label_hist + lab, 1ull);  // QueryStats._count_labels (query.py:266-268)
    }
    if (j < 1024) sh_lab[j] = lab;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (labels) {  // len(set(labels)) (query.py:300)
      int u = 0;
      const int m = n < 1024 ? n : 1024;
      for (int a = 0; a < m; ++a) {
        bool seen = false;
        for (int b = 0; b < a; ++b) seen = seen || sh_lab[b] == sh_lab[a];
        u += seen ? 0 : 1;
      }
      n_unique[img] = u;
    }
    // QueryStats._spatial_coverage (query.py:270-279): mean pairwise distance over the n (n - 1) ordered pairs; NaN for n < 2
    if (n < 2) {
      coverage[img] = __longlong_as_double(0x7FF8000000000000LL);
    } else {
      PickDist d;
      d.idx = s; d.n = n; d.W = W;
      coverage[img] = pairwise_sum_np(d, 0, n * (n - 1)) / (double)(n * (n - 1));
    }
  }
}


template <int C>
static int launch_score_select(const ScoreParams& p, const FusedSelParams& f, int strategy, int cl, int n_img, cudaStream_t st) {
  const size_t smem = (size_t)kFusedMaxPx * 4 + 2 * kHistBins * 4 + 8 * 512 * 4;
  auto go = [&](auto kernel) -> int {
    PP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (cl > 8) PP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)cl, (unsigned)n_img);
    cfg.blockDim = dim3(kScoreThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cl;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    PP_CUDA(cudaLaunchKernelEx(&cfg, kernel, p, f));
    PP_LAUNCH_CHECK();
    return PP_OK;
  };
  if (strategy == PP_STRAT_ENTROPY) return go(acq_score_select_kernel<C, PP_STRAT_ENTROPY>);
  if (strategy == PP_STRAT_LEAST_CONFIDENCE) return go(acq_score_select_kernel<C, PP_STRAT_LEAST_CONFIDENCE>);
  return go(acq_score_select_kernel<C, PP_STRAT_MARGIN>);
}

}  // namespace pp

using namespace pp;

extern "C" {

int pp_acq_score_select(const void* logits, int dtype, int n_img, int C, int H, int W, int64_t stride_n, int64_t stride_c,
                        int64_t stride_h, const uint8_t* labelled, const uint8_t* void_mask, const uint8_t* keep, int strategy,
                        int k, float* score_map, void* workspace, size_t workspace_bytes, void* stream) {
  PP_CHECK_ARG(logits && workspace, "pp_acq_score_select: null pointer");
  PP_CHECK_ARG(n_img > 0 && n_img <= 65535 && C >= 2 && H > 0 && W > 0, "pp_acq_score_select: bad shape n=%d C=%d H=%d W=%d", n_img, C, H, W);
  PP_CHECK_ARG(strategy >= 0 && strategy <= 2, "pp_acq_score_select: bad strategy %d", strategy);
  PP_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) % 256) == 0, "pp_acq_score_select: workspace must be 256-B aligned");
  const int64_t HW64 = (int64_t)H * W;
  PP_CHECK_ARG(HW64 <= (1 << 22) && k > 0 && k <= HW64, "pp_acq_score_select: bad H*W=%lld k=%d", (long long)HW64, k);
  const int HW = (int)HW64;
  auto al = [](const void* q, size_t a) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % a) == 0; };
  const bool vec_ok = dtype == PP_F32 && (W % 4 == 0) && (stride_n % 4 == 0) && (stride_c % 4 == 0) && (stride_h % 4 == 0) &&
                      stride_h >= W && al(logits, 16) && al(score_map, 16) && al(labelled, 4) && al(void_mask, 4) && al(keep, 4);
  int cl = 0;
  for (int c = 1; c <= 16; c <<= 1)
    if (HW % (c * 4096) == 0 && HW / c <= kFusedMaxPx) { cl = c; break; }
  if (!vec_ok || cl == 0 || !(C == 11 || C == 19 || C == 21)) {
    set_error("pp_acq_score_select: shape not covered by the fused kernel (f32, W %% 4 == 0, 16-byte aligned, C in {11, 19, 21}, H*W a "
              "multiple of 4096 x cluster size <= 16 with <= 16384 pixels per CTA): use pp_acq_score + pp_acq_select");
    return PP_ERR_UNSUPPORTED;
  }
  Workspace w = carve(workspace, n_img, HW, k);
  if (workspace_bytes < w.total_bytes) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, w.total_bytes);
    return PP_ERR_WORKSPACE;
  }
  ScoreParams p;
  p.logits = logits;
  p.sn = stride_n; p.sc = stride_c; p.sh = stride_h;
  p.n_img = n_img; p.C = C; p.H = H; p.W = W;
  p.lab = labelled; p.vd = void_mask; p.keep = keep;
  p.score = score_map; p.hist0 = nullptr;
  p.fill = (strategy == PP_STRAT_MARGIN) ? 1.0f : 0.0f;
  p.largest = (strategy == PP_STRAT_MARGIN) ? 0 : 1;
  FusedSelParams f;
  f.state = w.state + (size_t)n_img;
  f.cand = w.cand; f.cand_count = w.cand_count;
  f.bnd = w.filt; f.bnd_count = w.filt_count;
  f.k = (uint32_t)k; f.kpad = w.kpad; f.pxc = HW / cl;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc;
  switch (C) {
    case 11: rc = launch_score_select<11>(p, f, strategy, cl, n_img, st); break;
    case 19: rc = launch_score_select<19>(p, f, strategy, cl, n_img, st); break;
    default: rc = launch_score_select<21>(p, f, strategy, cl, n_img, st); break;
  }
  if (rc != PP_OK) return rc;
  RestParams r;
  r.list_a = w.filt;
  r.list_b = w.filt + (size_t)n_img * HW;
  r.count_a = w.filt_count;
  r.hist0 = w.hist;
  r.cand = w.cand;
  r.cand_count = w.cand_count;
  r.state1 = w.state + (size_t)n_img;
  r.HW = HW;
  r.kpad = w.kpad;
  r.first_level = p.largest ? 0 : 1;
  select_rest_kernel<<<n_img, kRestThreads, 0, st>>>(r);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int pp_acq_score(const void* logits, int dtype, int n_img, int C, int H, int W, int64_t stride_n,
                 int64_t stride_c, int64_t stride_h, const uint8_t* labelled, const uint8_t* void_mask,
                 const uint8_t* keep, int strategy, float* score_map, uint32_t* hist0, void* stream) {
  PP_CHECK_ARG(logits && score_map, "pp_acq_score: null logits/score_map");
  PP_CHECK_ARG(n_img > 0 && C >= 2 && H > 0 && W > 0, "pp_acq_score: bad shape n=%d C=%d H=%d W=%d", n_img, C, H, W);
  PP_CHECK_ARG((int64_t)H * W <= (1 << 22), "pp_acq_score: H*W=%lld exceeds 2^22", (long long)H * W);
  PP_CHECK_ARG(n_img <= 65535, "pp_acq_score: n_img=%d exceeds 65535 per call", n_img);
  PP_CHECK_ARG(strategy >= 0 && strategy <= 2, "pp_acq_score: bad strategy %d", strategy);
  PP_CHECK_ARG(dtype == PP_F32 || dtype == PP_BF16, "pp_acq_score: bad dtype %d", dtype);
  PP_CHECK_ARG(stride_h >= W, "pp_acq_score: stride_h < W");
  ScoreParams p;
  p.logits = logits;
  p.sn = stride_n; p.sc = stride_c; p.sh = stride_h;
  p.n_img = n_img; p.C = C; p.H = H; p.W = W;
  p.lab = labelled; p.vd = void_mask; p.keep = keep;
  p.score = score_map; p.hist0 = hist0;
  p.fill = (strategy == PP_STRAT_MARGIN) ? 1.0f : 0.0f;
  p.largest = (strategy == PP_STRAT_MARGIN) ? 0 : 1;
  const size_t esz = dtype == PP_F32 ? 4 : 2;
  const size_t valign = esz * 4;
  auto al = [](const void* q, size_t a) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % a) == 0; };
  const bool vec_ok = (W % 4 == 0) && (stride_n % 4 == 0) && (stride_c % 4 == 0) && (stride_h % 4 == 0) &&
                      al(logits, valign) && al(score_map, 16) && al(labelled, 4) && al(void_mask, 4) && al(keep, 4);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == PP_F32) return score_dispatch<float>(p, strategy, vec_ok, st);
  return score_dispatch<__nv_bfloat16>(p, strategy, vec_ok, st);
}

int pp_acq_score_upsampled(const float* logits_lowres, int n_img, int C, int h_in, int w_in, int H, int W,
                           const uint8_t* labelled, const uint8_t* void_mask, const uint8_t* keep,
                           int strategy, float* score_map, uint32_t* hist0, void* stream) {
  PP_CHECK_ARG(logits_lowres && score_map, "pp_acq_score_upsampled: null pointer");
  PP_CHECK_ARG(n_img > 0 && n_img <= 65535 && h_in > 0 && w_in > 0 && H > 0 && W > 0, "pp_acq_score_upsampled: bad shape");
  PP_CHECK_ARG((int64_t)H * W <= (1 << 22), "pp_acq_score_upsampled: H*W exceeds 2^22");
  PP_CHECK_ARG(strategy >= 0 && strategy <= 2, "pp_acq_score_upsampled: bad strategy %d", strategy);
  ScoreUpParams p;
  p.logits = logits_lowres;
  p.n_img = n_img; p.C = C; p.h_in = h_in; p.w_in = w_in; p.H = H; p.W = W;
  // ATen area_pixel_compute_scale(align_corners=True): (in-1)/(out-1), 0 when out == 1
  p.scale_h = ac_scale(h_in, H);
  p.scale_w = ac_scale(w_in, W);
  p.lab = labelled; p.vd = void_mask; p.keep = keep;
  p.score = score_map; p.hist0 = hist0;
  p.fill = (strategy == PP_STRAT_MARGIN) ? 1.0f : 0.0f;
  p.largest = (strategy == PP_STRAT_MARGIN) ? 0 : 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t HW = (int64_t)H * W;
  dim3 grid((unsigned)((HW + kScoreThreads - 1) / kScoreThreads), n_img);
  bool ok;
  if (strategy == PP_STRAT_ENTROPY) ok = dispatch_up<PP_STRAT_ENTROPY>(p, grid, st);
  else if (strategy == PP_STRAT_LEAST_CONFIDENCE) ok = dispatch_up<PP_STRAT_LEAST_CONFIDENCE>(p, grid, st);
  else ok = dispatch_up<PP_STRAT_MARGIN>(p, grid, st);
  if (!ok) {
    set_error("pp_acq_score_upsampled: C=%d not instantiated (11, 19, 21)", C);
    return PP_ERR_UNSUPPORTED;
  }
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int pp_acq_topk_workspace_bytes(int n_img, int HW, int k, size_t* out_bytes) {
  PP_CHECK_ARG(out_bytes, "pp_acq_topk_workspace_bytes: null out");
  PP_CHECK_ARG(n_img > 0 && HW > 0 && k > 0 && k <= HW, "pp_acq_topk_workspace_bytes: bad n_img=%d HW=%d k=%d", n_img, HW, k);
  PP_CHECK_ARG(HW <= (1 << 22), "pp_acq_topk_workspace_bytes: HW=%d exceeds 2^22", HW);
  Workspace w = carve(nullptr, n_img, HW, k);
  *out_bytes = w.total_bytes;
  return PP_OK;
}

int pp_acq_topk_prepare(void* workspace, size_t workspace_bytes, int n_img, int HW, int k, void* stream) {
  PP_CHECK_ARG(workspace, "pp_acq_topk_prepare: null workspace");
  PP_CHECK_ARG(n_img > 0 && HW > 0 && k > 0 && k <= HW && HW <= (1 << 22), "pp_acq_topk_prepare: bad shape");
  PP_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) % 256) == 0, "pp_acq_topk_prepare: workspace must be 256-B aligned");
  Workspace w = carve(workspace, n_img, HW, k);
  if (workspace_bytes < w.total_bytes) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, w.total_bytes);
    return PP_ERR_WORKSPACE;
  }
  PP_CUDA(cudaMemsetAsync(workspace, 0, w.zero_bytes, reinterpret_cast<cudaStream_t>(stream)));
  return PP_OK;
}

uint32_t* pp_acq_topk_hist0(void* workspace) { return reinterpret_cast<uint32_t*>(workspace); }

int pp_acq_topk(const float* score_map, int n_img, int HW, int k, int largest, int hist0_valid,
                int32_t* topk_idx, float* topk_val, void* workspace, size_t workspace_bytes, void* stream) {
  PP_CHECK_ARG(score_map && topk_idx && workspace, "pp_acq_topk: null pointer");
  PP_CHECK_ARG(n_img > 0 && n_img <= 65535 && HW > 0 && k > 0 && k <= HW, "pp_acq_topk: bad n_img=%d HW=%d k=%d", n_img, HW, k);
  PP_CHECK_ARG(HW <= (1 << 22), "pp_acq_topk: HW=%d exceeds 2^22", HW);
  PP_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) % 256) == 0, "pp_acq_topk: workspace must be 256-B aligned");
  return topk_impl(score_map, n_img, HW, k, largest, hist0_valid, topk_idx, topk_val, workspace,
                   workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

int pp_acq_select(const float* score_map, int n_img, int HW, int k, int largest, int hist0_valid, void* workspace,
                  size_t workspace_bytes, void* stream) {
  PP_CHECK_ARG(score_map && workspace, "pp_acq_select: null pointer");
  PP_CHECK_ARG(n_img > 0 && n_img <= 65535 && HW > 0 && k > 0 && k <= HW, "pp_acq_select: bad n_img=%d HW=%d k=%d", n_img, HW, k);
  PP_CHECK_ARG(HW <= (1 << 22), "pp_acq_select: HW=%d exceeds 2^22", HW);
  PP_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) % 256) == 0, "pp_acq_select: workspace must be 256-B aligned");
  return select_impl(score_map, n_img, HW, k, largest, hist0_valid, workspace, workspace_bytes,
                     reinterpret_cast<cudaStream_t>(stream));
}

int pp_acq_pick(void* workspace, size_t workspace_bytes, int n_img, int HW, int k, const int32_t* pos, int n,
                int32_t* out, void* stream) {
  PP_CHECK_ARG(workspace && out, "pp_acq_pick: null pointer");
  PP_CHECK_ARG(n_img > 0 && HW > 0 && k > 0 && k <= HW && n > 0 && n <= k, "pp_acq_pick: bad n_img=%d HW=%d k=%d n=%d", n_img, HW, k, n);
  Workspace w = carve(workspace, n_img, HW, k);
  if (workspace_bytes < w.total_bytes) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, w.total_bytes);
    return PP_ERR_WORKSPACE;
  }
  static bool attr = false;
  const int smem = kPickRanks * kHistBins * (int)sizeof(uint32_t);
  if (!attr) {
    // two 97 KB CTAs per SM (64 registers / thread): 256 images are then ONE wave on 148 SMs instead of two
    PP_CUDA(cudaFuncSetAttribute(pick_ranks_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    PP_CUDA(cudaFuncSetAttribute(pick_ranks_fast_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    PP_CUDA(cudaFuncSetAttribute(pick_ranks_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    PP_CUDA(cudaFuncSetAttribute(pick_ranks_fast_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr = true;
  }
  PickParams p;
  p.cand = w.cand;
  p.kpad = w.kpad;
  p.k = k;
  p.n = n;
  p.pos = pos;
  p.out = out;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // ONE launch: three cheap passes per image; images it cannot finish (heavy ties, n > 12) run the generic walk inside it
  if (k <= kPickThreads * kPickItems) pick_ranks_fast_kernel<true><<<n_img, kPickThreads, smem, st>>>(p);   // candidates in registers
  else pick_ranks_fast_kernel<false><<<n_img, kPickThreads, smem, st>>>(p);                                  // streamed from L2
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int pp_acq_gather(const int32_t* topk_idx, int n_img, int k, const int32_t* pos, int n, int32_t* out,
                  void* stream) {
  PP_CHECK_ARG(topk_idx && out, "pp_acq_gather: null pointer");
  PP_CHECK_ARG(n_img > 0 && k > 0 && n > 0 && n <= k, "pp_acq_gather: bad n_img=%d k=%d n=%d", n_img, k, n);
  const int total = n_img * n;
  gather_kernel<<<(total + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(topk_idx, k, pos, n, out, total);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int pp_acq_entropy_at(const void* logits, int dtype, int n_img, int C, int H, int W, int64_t stride_n,
                      int64_t stride_c, int64_t stride_h, const int32_t* px_idx, int n, float* out,
                      void* stream) {
  PP_CHECK_ARG(logits && px_idx && out, "pp_acq_entropy_at: null pointer");
  PP_CHECK_ARG(n_img > 0 && C >= 2 && H > 0 && W > 0 && n > 0, "pp_acq_entropy_at: bad shape");
  const int total = n_img * n;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == PP_F32)
    entropy_at_kernel<float><<<(total + 127) / 128, 128, 0, st>>>(reinterpret_cast<const float*>(logits), C, W, H * W, stride_n, stride_c, stride_h, px_idx, n, out, total);
  else if (dtype == PP_BF16)
    entropy_at_kernel<__nv_bfloat16><<<(total + 127) / 128, 128, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(logits), C, W, H * W, stride_n, stride_c, stride_h, px_idx, n, out, total);
  else {
    set_error("pp_acq_entropy_at: bad dtype %d", dtype);
    return PP_ERR_INVALID_ARG;
  }
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int pp_acq_entropy_at_upsampled(const float* logits_lowres, int n_img, int C, int h_in, int w_in, int H, int W,
                                const int32_t* px_idx, int n, float* out, void* stream) {
  PP_CHECK_ARG(logits_lowres && px_idx && out, "pp_acq_entropy_at_upsampled: null pointer");
  PP_CHECK_ARG(n_img > 0 && C >= 2 && h_in > 0 && w_in > 0 && H > 0 && W > 0 && n > 0, "pp_acq_entropy_at_upsampled: bad shape");
  const int total = n_img * n;
  entropy_at_up_kernel<<<(total + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      logits_lowres, C, h_in, w_in, H, W, ac_scale(h_in, H), ac_scale(w_in, W), px_idx, n, out, total);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

int pp_query_stats_at(const long long* sel_sorted, int n_img, int n, int W, int HW, const uint8_t* labels, int n_classes,
                      long long* x_coords, long long* y_coords, int32_t* labels_at, long long* label_hist, int32_t* n_unique,
                      double* coverage, void* stream) {
  PP_CHECK_ARG(sel_sorted && x_coords && y_coords && coverage, "pp_query_stats_at: null pointer");
  PP_CHECK_ARG(!labels || (labels_at && label_hist && n_unique), "pp_query_stats_at: labels need labels_at / label_hist / n_unique");
  PP_CHECK_ARG(n_img > 0 && n > 0 && n <= 1024 && W > 0 && HW > 0 && n_classes > 0, "pp_query_stats_at: bad shape");
  query_stats_kernel<<<n_img, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      sel_sorted, n_img, n, W, HW, labels, n_classes, x_coords, y_coords, labels_at,
      reinterpret_cast<unsigned long long*>(label_hist), n_unique, coverage);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

}  // extern "C"
