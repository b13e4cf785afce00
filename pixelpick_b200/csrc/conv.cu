// T path — NHWC bf16 implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
// Serves every dense contraction of the DeepLabv3+ head: the ASPP 1x1 and
// dilated 3x3 branches (aspp.py:49-52), the 1280->256 projection (aspp.py:73-75), the two SegmentHead
// 3x3 convs and the classifier (decoders.py:107-116), and — with transposed/flipped weights — their
// data gradients.  Stride 1, "same" zero padding (pad = dilation), kernel 1x1 or 3x3.
//
// GEMM view: M = output pixels (tile = TH x TW rectangle of one image = 128 rows), N = C_out,
// K = taps x C_in.  No im2col buffer: for tap (dy,dx) the A tile is ONE 4-D TMA box
// {64 ch, TW, TH, 1} at (c0, x0+dx*dil, y0+dy*dil, n); out-of-image pixels are zero-filled by TMA,
// which IS the conv's zero padding.  The box lands in shared memory as 128 rows x 128 B with the
// 128-byte swizzle == the canonical K-major UMMA operand layout.  Taps that fall entirely outside
// the image for a tile are skipped (ASPP at 16x32 with dilation 18: 6 of 9 taps).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
// warps 2-9 = epilogue (TMEM -> registers -> global; two warps per TMEM lane quarter).  Persistent over output tiles; the
// accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cstdlib>
#include <cstring>
#include "tc_common.cuh"

namespace pp {

// ------------------------------------------------------------------------------------------
// tensor-map encode through the driver entry point (no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  return make_tmap_bf16_sw(out, base, rank, dims, strides_bytes, box, 128);
}

int make_tmap_bf16_sw(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not found");
    return PP_ERR_CUDA;
  }
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    e[i] = 1;
    if (i > 0) s[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, e,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: %d (rank %d dims %llu %llu %llu box %u %u %u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
              box[0], box[1], rank > 2 ? box[2] : 0);
    return PP_ERR_CUDA;
  }
  return PP_OK;
}

// ------------------------------------------------------------------------------------------
// forward / dgrad kernel
// ------------------------------------------------------------------------------------------
constexpr int kBM = 128;  // output pixels per tile (UMMA M)
constexpr int kBK = 64;   // channels per k-block (128 B of bf16 = one swizzle row)
constexpr int kConvThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
constexpr int kMaxTaps = 28;

struct ConvParams {
  int N, H, W;       // images, spatial size (output == input)
  int Cin;           // padded to a multiple of 64
  int Cout_pad;      // rows of the packed weight tensor (multiple of BN)
  int Cout;          // valid output channels (<= Cout_pad)
  int taps;          // number of tap entries (K = taps x Cin)
  // tap table: entry t reads the A tile shifted by (tdy, tdx) pixels at channel offset ta of the A tensor and uses
  // weight slice t.  A plain 3x3 dilated conv is 9 entries (dy, dx) = ((t/3-1)*dil, (t%3-1)*dil), ta = 0; the fused
  // ASPP data gradient is 1 + 9 + 9 + 9 entries with per-branch dilations and channel offsets.
  short tdy[kMaxTaps], tdx[kMaxTaps], ta[kMaxTaps];
  int n_off, n_tiles_n;  // this launch covers output channels [n_off, n_off + n_tiles_n * BN)
  int TH, TW;        // tile rectangle, TH * TW == 128
  int tiles_y, tiles_x;
  const float* pre_bias;  // [N][Cout_pad] added before scale/shift (ASPP image-pooling branch) or null
  const float* scale;     // [Cout_pad] or null (=1)
  const float* shift;     // [Cout_pad] or null (=0)
  int relu;      // 0 none, 1 ReLU, 2 ReLU6
  int out_mode;  // 0: bf16 NHWC (ld_out, c_off)   1: f32 NCHW [N][Cout][H][W]
  void* out;
  int ld_out, c_off;
  const __nv_bfloat16* res;  // optional residual [pixel][ld_res], channel c of the output adds res[c]; before the activation
  int ld_res;
  int tma_store;  // bf16 NHWC output leaves through shared-memory staging + TMA store (full-line writes, ragged edges clipped)
  float* stats;   // optional [2][Cout]: per-channel sum / sum of squares of the (bf16-rounded) outputs += this launch
};

template <int BN>
struct ConvCfg {
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr uint32_t A_BYTES = kBM * kBK * 2;
  static constexpr uint32_t B_BYTES = BN * kBK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;  // power of two for BN in {16..256}
  static constexpr uint32_t STAGING_BYTES = 8 * 2 * 2048;  // 8 epilogue warps x 2 buffers x (32 pixels x 32 channels bf16)
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

__device__ __forceinline__ bool tap_skipped(const ConvParams& p, int tap, int y0, int x0, int& dy, int& dx) {
  dy = p.tdy[tap];
  dx = p.tdx[tap];
  return (y0 + dy + p.TH <= 0) || (y0 + dy >= p.H) || (x0 + dx + p.TW <= 0) || (x0 + dx >= p.W);
}

// EPI: the per-element work of the staged (TMA-store) epilogue, fixed at compile time - with 8 epilogue warps on 4 schedulers
// the epilogue of a short-K conv is bound by the instructions it issues (ncu: 2.3 IPC, 58 % issue slots busy on the 64 -> 256
// 1x1), so flag tests, selects and dead arithmetic per element cost as much as the stores:
//   kEpiGeneric : pre-bias, affine, residual, activation all optional at run time (also the only variant with the direct path)
//   kEpiRaw     : out = bf16(acc)                        - training forward (BatchNorm follows) and every data gradient
//   kEpiAffine  : out = clamp(acc * scale + shift)       - inference convs with folded BatchNorm + ReLU / ReLU6
//   kEpiRawRes  : out = bf16(acc + res)                  - a data gradient that lands on a residual fork: the other branch's
//                                                          gradient is added here instead of by a separate pass (Cout % 32 == 0)
constexpr int kEpiGeneric = 0, kEpiRaw = 1, kEpiAffine = 2, kEpiRawRes = 3;

template <int BN, int EPI>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmO, const ConvParams p) {
  using Cfg = ConvCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;  // 1024-aligned: stage sizes are multiples of 4 KB
  const uint32_t bar_base = stg_base + Cfg::STAGING_BYTES;
  auto sA = [&](int s) { return smem_base + (uint32_t)s * Cfg::STAGE_BYTES; };
  auto sB = [&](int s) { return smem_base + (uint32_t)s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - tc::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_n_tiles = p.n_tiles_n;
  const int m_tiles_per_img = p.tiles_y * p.tiles_x;
  const int total_tiles = p.N * m_tiles_per_img * n_n_tiles;
  const int kc_per_tap = p.Cin / kBK;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
    if (p.tma_store) tc::tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      tc::mbar_init(full_bar(s), 1);
      tc::mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      tc::mbar_init(tfull_bar(a), 1);
      tc::mbar_init(tempty_bar(a), 8);
    }
    tc::fence_barrier_init();
  }
  if (warp == 2) {
    tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  auto decode = [&](int tile, int& n0, int& img, int& y0, int& x0) {
    const int nt = tile % n_n_tiles;
    int mt = tile / n_n_tiles;
    n0 = p.n_off + nt * BN;
    img = mt / m_tiles_per_img;
    mt -= img * m_tiles_per_img;
    const int ty = mt / p.tiles_x;
    y0 = ty * p.TH;
    x0 = (mt - ty * p.tiles_x) * p.TW;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int n0, img, y0, x0;
        decode(tile, n0, img, y0, x0);
        for (int tap = 0; tap < p.taps; ++tap) {
          int dy, dx;
          if (tap_skipped(p, tap, y0, x0, dy, dx)) continue;
          const int a0 = p.ta[tap];
          for (int kc = 0; kc < kc_per_tap; ++kc, ++it) {
            const int s = it % Cfg::STAGES;
            const uint32_t ph = (it / Cfg::STAGES) & 1u;
            tc::mbar_wait(empty_bar(s), ph ^ 1u);
            tc::mbar_expect_tx(full_bar(s), Cfg::STAGE_BYTES);
            tc::tma_load_4d(sA(s), &tmA, full_bar(s), a0 + kc * kBK, x0 + dx, y0 + dy, img);
            tc::tma_load_3d(sB(s), &tmB, full_bar(s), kc * kBK, n0, tap);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(kBM, BN, 0, 0);
      uint32_t it = 0, tile_iter = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_iter) {
        int n0, img, y0, x0;
        decode(tile, n0, img, y0, x0);
        const uint32_t acc = tile_iter & 1u, aph = (tile_iter >> 1) & 1u;
        tc::mbar_wait(tempty_bar(acc), aph ^ 1u);
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        uint32_t accumulate = 0;
        for (int tap = 0; tap < p.taps; ++tap) {
          int dy, dx;
          if (tap_skipped(p, tap, y0, x0, dy, dx)) continue;
          for (int kc = 0; kc < kc_per_tap; ++kc, ++it) {
            const int s = it % Cfg::STAGES;
            const uint32_t ph = (it / Cfg::STAGES) & 1u;
            tc::mbar_wait(full_bar(s), ph);
            tc::tc_fence_after();
            const uint64_t da = tc::umma_desc_sw128(sA(s), 16, 1024);
            const uint64_t db = tc::umma_desc_sw128(sB(s), 16, 1024);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              tc::umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, accumulate);
              accumulate = 1;
            }
            tc::umma_commit(empty_bar(s));
          }
        }
        tc::umma_commit(tfull_bar(acc));
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // Two warps share each TMEM lane quarter and split the accumulator's 32-column chunks between them: with one warp
    // per quarter the epilogue (not the MMA pipe) bounded every conv whose K is short relative to its output (ncu /
    // per-layer timings in DESIGN.md).  A shared-memory transposed, fully coalesced store path was also measured here
    // and was SLOWER: the epilogue is instruction/latency-bound, not store-pattern-bound.
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int eh = (warp - 2) >> 2;  // which half of the column chunks this warp converts and stores
    const int row = quarter * 32 + lane;
    const int ty_in = row / p.TW, tx_in = row - ty_in * p.TW;
    uint32_t tile_iter = 0;
    if (p.out_mode == 0 && p.tma_store) {
      // ---- bf16 NHWC output through shared memory + TMA store ----
      // A unit = this warp's 32 pixels x one 32-channel chunk = 2 KB, written to a 64-byte-swizzled staging buffer (each lane
      // its pixel's 64 B as four conflict-free 16-byte stores) and shipped by ONE bulk tensor store: the memory system sees
      // whole 64-byte runs instead of 32 scattered 16-byte pieces per instruction, ragged channel counts / image edges are
      // clipped by the tensor map, and the warp moves on while the store drains (two buffers per warp).
      // With p.stats the per-channel sum and sum of squares of the ROUNDED outputs (what BatchNorm will read back) are
      // accumulated from the staged tile: the BatchNorm forward then needs no statistics pass over the tensor.
      const uint32_t my_stage = stg_base + (uint32_t)(warp - 2) * 4096u;
      const int rows_per_q = 32 / p.TW;  // tile rows covered by one TMEM lane quarter
      const float act_lo = p.relu ? 0.f : -INFINITY, act_hi = p.relu == 2 ? 6.f : INFINITY;  // none / ReLU / ReLU6 as a clamp
      uint32_t unit_iter = 0;
      constexpr int NSLOT = BN >= 64 ? BN / 64 : 1;  // 32-channel chunks per warp and tile (the two warps of a quarter alternate)
      float st_s[NSLOT][2], st_q[NSLOT][2];
#pragma unroll
      for (int i = 0; i < NSLOT; ++i) st_s[i][0] = st_s[i][1] = st_q[i][0] = st_q[i][1] = 0.f;
      int n0_cta = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_iter) {
        int n0, img, y0, x0;
        decode(tile, n0, img, y0, x0);
        n0_cta = n0;
        const uint32_t acc = tile_iter & 1u, aph = (tile_iter >> 1) & 1u;
        const int y = y0 + ty_in, x = x0 + tx_in;
        const bool valid = (y < p.H) && (x < p.W);
        const size_t pix = ((size_t)img * p.H + y) * p.W + x;
        uint4 rq[EPI == kEpiRawRes ? NSLOT : 1][4];
        if constexpr (EPI == kEpiRawRes) {
          // this pixel's 64 B per chunk of the other branch: issued BEFORE waiting for the accumulator, so the loads fly
          // under the tile's main loop instead of stalling every unit for a DRAM round trip
#pragma unroll
          for (int slot = 0; slot < NSLOT; ++slot) {
            const int cb = n0 + (eh + 2 * slot) * 32;
            const bool on = valid && (eh + 2 * slot) < BN / 32 && cb < p.Cout;
#pragma unroll
            for (int u = 0; u < 4; ++u)
              rq[slot][u] = on ? __ldg(reinterpret_cast<const uint4*>(p.res + pix * p.ld_res + cb) + u) : make_uint4(0u, 0u, 0u, 0u);
          }
        }
        tc::mbar_wait(tfull_bar(acc), aph);
        tc::tc_fence_after();
#pragma unroll
        for (int slot = 0; slot < NSLOT; ++slot) {
          const int j = eh + 2 * slot;
          const int cbase = n0 + j * 32;
          if (j >= BN / 32 || cbase >= p.Cout) continue;  // warp-uniform: no such chunk / entirely in the channel padding
          uint32_t r[32];
          tc::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + j * 32, r);
          tc::tmem_ld_wait();
          uint32_t pk[16];
          if constexpr (EPI == kEpiRawRes) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint4 q4 = rq[EPI == kEpiRawRes ? slot : 0][u];
              const uint32_t w4[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int c = 8 * u + 2 * e;
                __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(r[c]) + __uint_as_float(w4[e] << 16),
                                                          __uint_as_float(r[c + 1]) + __uint_as_float(w4[e] & 0xFFFF0000u));
                pk[c / 2] = *reinterpret_cast<uint32_t*>(&t0);
              }
            }
          } else if constexpr (EPI == kEpiRaw) {
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
              __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(r[c]), __uint_as_float(r[c + 1]));
              pk[c / 2] = *reinterpret_cast<uint32_t*>(&t0);
            }
          } else if constexpr (EPI == kEpiAffine) {
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              const float4 sc4 = __ldg(reinterpret_cast<const float4*>(p.scale + cbase + c));
              const float4 sf4 = __ldg(reinterpret_cast<const float4*>(p.shift + cbase + c));
              const float a0 = fminf(fmaxf(fmaf(__uint_as_float(r[c]), sc4.x, sf4.x), act_lo), act_hi);
              const float a1 = fminf(fmaxf(fmaf(__uint_as_float(r[c + 1]), sc4.y, sf4.y), act_lo), act_hi);
              const float a2 = fminf(fmaxf(fmaf(__uint_as_float(r[c + 2]), sc4.z, sf4.z), act_lo), act_hi);
              const float a3 = fminf(fmaxf(fmaf(__uint_as_float(r[c + 3]), sc4.w, sf4.w), act_lo), act_hi);
              __nv_bfloat162 t0 = __floats2bfloat162_rn(a0, a1);
              __nv_bfloat162 t1 = __floats2bfloat162_rn(a2, a3);
              pk[c / 2] = *reinterpret_cast<uint32_t*>(&t0);
              pk[c / 2 + 1] = *reinterpret_cast<uint32_t*>(&t1);
            }
          } else {
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              float4 pb = make_float4(0.f, 0.f, 0.f, 0.f), sc4 = make_float4(1.f, 1.f, 1.f, 1.f), sf4 = pb;
              if (p.pre_bias) pb = __ldg(reinterpret_cast<const float4*>(p.pre_bias + (size_t)img * p.Cout_pad + cbase + c));
              if (p.scale) sc4 = __ldg(reinterpret_cast<const float4*>(p.scale + cbase + c));
              if (p.shift) sf4 = __ldg(reinterpret_cast<const float4*>(p.shift + cbase + c));
              const float pbv[4] = {pb.x, pb.y, pb.z, pb.w}, scv[4] = {sc4.x, sc4.y, sc4.z, sc4.w}, sfv[4] = {sf4.x, sf4.y, sf4.z, sf4.w};
              float rs4[4] = {0.f, 0.f, 0.f, 0.f};
              if (p.res && valid && cbase + c + 4 <= p.Cout) {
                const uint2 q = __ldg(reinterpret_cast<const uint2*>(p.res + pix * p.ld_res + cbase + c));
                rs4[0] = __uint_as_float(q.x << 16);
                rs4[1] = __uint_as_float(q.x & 0xFFFF0000u);
                rs4[2] = __uint_as_float(q.y << 16);
                rs4[3] = __uint_as_float(q.y & 0xFFFF0000u);
              }
              float a4[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                a4[e] = fminf(fmaxf(fmaf(__uint_as_float(r[c + e]) + pbv[e], scv[e], sfv[e]) + rs4[e], act_lo), act_hi);
              __nv_bfloat162 t0 = __floats2bfloat162_rn(a4[0], a4[1]);
              __nv_bfloat162 t1 = __floats2bfloat162_rn(a4[2], a4[3]);
              pk[c / 2] = *reinterpret_cast<uint32_t*>(&t0);
              pk[c / 2 + 1] = *reinterpret_cast<uint32_t*>(&t1);
            }
          }
          if (p.stats && !valid) {
            // a tile that overhangs the image edge: its out-of-image pixels still see in-image taps of a 3x3 conv; they are
            // clipped by the TMA store but must not enter the BatchNorm statistics
#pragma unroll
            for (int c = 0; c < 16; ++c) pk[c] = 0u;
          }
          const uint32_t buf = my_stage + (unit_iter & 1u) * 2048u;
          if (unit_iter >= 2) {  // the store issued from this buffer two units ago must have read it
            if (lane == 0) tc::bulk_wait_group_read<1>();
            __syncwarp();
          }
          const uint32_t rowaddr = buf + (uint32_t)lane * 64u;
          const uint32_t sw = ((uint32_t)lane >> 1) & 3u;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + (((uint32_t)u ^ sw) << 4)), "r"(pk[4 * u]),
                         "r"(pk[4 * u + 1]), "r"(pk[4 * u + 2]), "r"(pk[4 * u + 3])
                         : "memory");
          if (p.stats) {
            __syncwarp();
            // lane -> channel pair w = lane & 15 of the rows with parity lane >> 4: one 4-byte word per row, conflict-free
            const uint32_t w = (uint32_t)lane & 15u, par = (uint32_t)lane >> 4;
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint32_t rr = 2u * i + par;
              uint32_t word;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(word) : "r"(buf + rr * 64u + (((w >> 2) ^ ((rr >> 1) & 3u)) << 4) + (w & 3u) * 4u));
              const float lo = __uint_as_float(word << 16), hi = __uint_as_float(word & 0xFFFF0000u);
              s0 += lo; s1 += hi;
              q0 = fmaf(lo, lo, q0); q1 = fmaf(hi, hi, q1);
            }
            s0 += __shfl_xor_sync(0xFFFFFFFFu, s0, 16);
            s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, 16);
            q0 += __shfl_xor_sync(0xFFFFFFFFu, q0, 16);
            q1 += __shfl_xor_sync(0xFFFFFFFFu, q1, 16);
            st_s[slot][0] += s0; st_s[slot][1] += s1;
            st_q[slot][0] += q0; st_q[slot][1] += q1;
          }
          tc::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tc::tma_store_4d(&tmO, buf, cbase, x0, y0 + quarter * rows_per_q, img);
            tc::bulk_commit_group();
          }
          ++unit_iter;
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(tempty_bar(acc));
      }
      if (p.stats && tile_iter > 0 && lane < 16) {
        // every tile of this CTA has the same channel range (the launch makes the grid a multiple of the N tiles)
#pragma unroll
        for (int slot = 0; slot < NSLOT; ++slot) {
          const int j = eh + 2 * slot;
          const int c = n0_cta + j * 32 + 2 * lane;
          if (j >= BN / 32) continue;
          if (c < p.Cout) {
            atomicAdd(p.stats + c, st_s[slot][0]);
            atomicAdd(p.stats + p.Cout + c, st_q[slot][0]);
          }
          if (c + 1 < p.Cout) {
            atomicAdd(p.stats + c + 1, st_s[slot][1]);
            atomicAdd(p.stats + p.Cout + c + 1, st_q[slot][1]);
          }
        }
      }
      if (lane == 0) tc::bulk_wait_group<0>();
      __syncwarp();
    } else if constexpr (EPI == kEpiGeneric)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_iter) {
      int n0, img, y0, x0;
      decode(tile, n0, img, y0, x0);
      const uint32_t acc = tile_iter & 1u, aph = (tile_iter >> 1) & 1u;
      tc::mbar_wait(tfull_bar(acc), aph);
      tc::tc_fence_after();
      const int y = y0 + ty_in, x = x0 + tx_in;
      const bool valid = (y < p.H) && (x < p.W);
      const size_t pix = ((size_t)img * p.H + y) * p.W + x;
#pragma unroll 1
      for (int j = eh; j < BN / 32; j += 2) {
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + j * 32, r);
        tc::tmem_ld_wait();
        const int cbase = n0 + j * 32;
        float v[32];
        float rs[32];
        if (p.res) {
#pragma unroll
          for (int c = 0; c < 32; ++c) rs[c] = 0.f;
          if (valid) {
            const __nv_bfloat16* rp = p.res + pix * p.ld_res + cbase;
#pragma unroll
            for (int c = 0; c < 32; c += 8) {
              if (cbase + c + 8 <= p.Cout) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(rp + c));
                const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  rs[c + 2 * e] = __uint_as_float(w4[e] << 16);
                  rs[c + 2 * e + 1] = __uint_as_float(w4[e] & 0xFFFF0000u);
                }
              }
            }
          }
        }
        // per-channel vectors are warp-uniform: 16-byte loads (4 channels each) instead of one load per channel
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          float4 pb = make_float4(0.f, 0.f, 0.f, 0.f), sc4 = make_float4(1.f, 1.f, 1.f, 1.f), sf4 = pb;
          if (p.pre_bias) pb = __ldg(reinterpret_cast<const float4*>(p.pre_bias + (size_t)img * p.Cout_pad + cbase + c));
          if (p.scale) sc4 = __ldg(reinterpret_cast<const float4*>(p.scale + cbase + c));
          if (p.shift) sf4 = __ldg(reinterpret_cast<const float4*>(p.shift + cbase + c));
          const float pbv[4] = {pb.x, pb.y, pb.z, pb.w}, scv[4] = {sc4.x, sc4.y, sc4.z, sc4.w}, sfv[4] = {sf4.x, sf4.y, sf4.z, sf4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float a = fmaf(__uint_as_float(r[c + e]) + pbv[e], scv[e], sfv[e]);
            if (p.res) a += rs[c + e];
            if (p.relu) a = fmaxf(a, 0.f);
            if (p.relu == 2) a = fminf(a, 6.f);
            v[c + e] = a;
          }
        }
        if (valid) {
          if (p.out_mode == 0) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.ld_out + p.c_off + cbase;
            if (cbase + 32 <= p.Cout) {
#pragma unroll
              for (int c = 0; c < 32; c += 8) {
                uint4 pk;
                __nv_bfloat162 t0 = __floats2bfloat162_rn(v[c], v[c + 1]);
                __nv_bfloat162 t1 = __floats2bfloat162_rn(v[c + 2], v[c + 3]);
                __nv_bfloat162 t2 = __floats2bfloat162_rn(v[c + 4], v[c + 5]);
                __nv_bfloat162 t3 = __floats2bfloat162_rn(v[c + 6], v[c + 7]);
                pk.x = *reinterpret_cast<uint32_t*>(&t0);
                pk.y = *reinterpret_cast<uint32_t*>(&t1);
                pk.z = *reinterpret_cast<uint32_t*>(&t2);
                pk.w = *reinterpret_cast<uint32_t*>(&t3);
                *reinterpret_cast<uint4*>(o + c) = pk;
              }
            } else {
              // ragged tile (Cout not a multiple of 32): 16-byte stores while 8 channels fit, scalars for the rest
#pragma unroll
              for (int c = 0; c < 32; c += 8) {
                if (cbase + c + 8 <= p.Cout) {
                  uint4 pk;
                  __nv_bfloat162 t0 = __floats2bfloat162_rn(v[c], v[c + 1]);
                  __nv_bfloat162 t1 = __floats2bfloat162_rn(v[c + 2], v[c + 3]);
                  __nv_bfloat162 t2 = __floats2bfloat162_rn(v[c + 4], v[c + 5]);
                  __nv_bfloat162 t3 = __floats2bfloat162_rn(v[c + 6], v[c + 7]);
                  pk.x = *reinterpret_cast<uint32_t*>(&t0);
                  pk.y = *reinterpret_cast<uint32_t*>(&t1);
                  pk.z = *reinterpret_cast<uint32_t*>(&t2);
                  pk.w = *reinterpret_cast<uint32_t*>(&t3);
                  *reinterpret_cast<uint4*>(o + c) = pk;
                } else {
#pragma unroll
                  for (int e = 0; e < 8; ++e)
                    if (cbase + c + e < p.Cout) o[c + e] = __float2bfloat16(v[c + e]);
                }
              }
            }
          } else {
            float* o = reinterpret_cast<float*>(p.out);
            const size_t plane = (size_t)p.H * p.W;
            const size_t off = (size_t)img * p.Cout * plane + (size_t)y * p.W + x;
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (cbase + c < p.Cout) o[off + (size_t)(cbase + c) * plane] = v[c];
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tempty_bar(acc));
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, int EPI>
static int launch_conv_epi(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const ConvParams& p, int sm_count,
                           cudaStream_t st) {
  using Cfg = ConvCfg<BN>;
  static bool attr = false;
  if (!attr) {
    PP_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    attr = true;
  }
  const int total = p.N * p.tiles_y * p.tiles_x * p.n_tiles_n;
  int grid = total < sm_count ? total : sm_count;
  // statistics are kept in registers per CTA: every tile of a CTA must cover the same channel range, i.e. the grid is a
  // multiple of the number of N tiles (tile -> N tile = tile % n_tiles_n; total is a multiple of it by construction)
  if (p.stats && p.n_tiles_n > 1) grid = grid / p.n_tiles_n * p.n_tiles_n;
  if (grid < 1) grid = p.n_tiles_n;
  conv_igemm_kernel<BN, EPI><<<grid, kConvThreads, Cfg::SMEM, st>>>(tmA, tmB, tmO, p);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

template <int BN>
static int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const ConvParams& p, int sm_count,
                       cudaStream_t st) {
  if (p.tma_store && !p.pre_bias && !p.res) {
    if (!p.scale && !p.shift && p.relu == 0) return launch_conv_epi<BN, kEpiRaw>(tmA, tmB, tmO, p, sm_count, st);
    if (p.scale && p.shift) return launch_conv_epi<BN, kEpiAffine>(tmA, tmB, tmO, p, sm_count, st);
  }
  if (p.tma_store && p.res && !p.pre_bias && !p.scale && !p.shift && p.relu == 0 && !p.stats && p.Cout % 32 == 0 &&
      p.ld_res % 8 == 0 && (reinterpret_cast<uintptr_t>(p.res) & 15u) == 0)
    return launch_conv_epi<BN, kEpiRawRes>(tmA, tmB, tmO, p, sm_count, st);
  return launch_conv_epi<BN, kEpiGeneric>(tmA, tmB, tmO, p, sm_count, st);
}

static int sm_count_cached() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace pp

using namespace pp;

extern "C" {

int pp_conv_igemm(const void* x, int N, int H, int W, int Cin, int ld_in, const void* w_packed, int taps, int dil,
                  int Cout_pad, int Cout, const float* pre_bias, const float* scale, const float* shift, int relu,
                  void* out, int out_mode, int ld_out, int c_off, int block_n, void* stream) {
  PP_CHECK_ARG(taps == 1 || taps == 9, "pp_conv_igemm: taps=%d (1 or 9)", taps);
  PP_CHECK_ARG(dil >= 1 && dil < 4096, "pp_conv_igemm: dil=%d", dil);
  int tdy[9], tdx[9], tc0[9];
  for (int t = 0; t < taps; ++t) {
    tdy[t] = taps == 1 ? 0 : (t / 3 - 1) * dil;
    tdx[t] = taps == 1 ? 0 : (t % 3 - 1) * dil;
    tc0[t] = 0;
  }
  return pp_conv_igemm_multi(x, N, H, W, Cin, ld_in, Cin, w_packed, taps, tdy, tdx, tc0, Cout_pad, Cout, pre_bias, scale,
                             shift, relu, nullptr, 0, out, out_mode, ld_out, c_off, block_n, stream);
}

int pp_conv_igemm_multi(const void* x, int N, int H, int W, int a_channels, int ld_in, int Cin, const void* w_packed,
                        int n_entries, const int* tap_dy, const int* tap_dx, const int* tap_c0, int Cout_pad, int Cout,
                        const float* pre_bias, const float* scale, const float* shift, int relu, const void* res,
                        int ld_res, void* out, int out_mode, int ld_out, int c_off, int block_n, void* stream) {
  return pp_conv_igemm_stats(x, N, H, W, a_channels, ld_in, Cin, w_packed, n_entries, tap_dy, tap_dx, tap_c0, Cout_pad, Cout,
                             pre_bias, scale, shift, relu, res, ld_res, out, out_mode, ld_out, c_off, block_n, nullptr, stream);
}

static int g_conv_tma_store = -1;

static int conv_tma_store_enabled() {
  if (g_conv_tma_store < 0) {
    const char* e = getenv("PP_CONV_TMA_STORE");  // "0": the direct register -> global epilogue (kept for A/B measurements)
    g_conv_tma_store = (e && e[0] == '0') ? 0 : 1;
  }
  return g_conv_tma_store;
}

int pp_conv_set_epilogue(int tma_store) {
  const int prev = conv_tma_store_enabled();
  g_conv_tma_store = tma_store ? 1 : 0;
  return prev;
}

int pp_conv_igemm_stats(const void* x, int N, int H, int W, int a_channels, int ld_in, int Cin, const void* w_packed,
                        int n_entries, const int* tap_dy, const int* tap_dx, const int* tap_c0, int Cout_pad, int Cout,
                        const float* pre_bias, const float* scale, const float* shift, int relu, const void* res,
                        int ld_res, void* out, int out_mode, int ld_out, int c_off, int block_n, float* stats, void* stream) {
  PP_CHECK_ARG(x && w_packed && out, "pp_conv_igemm: null pointer");
  PP_CHECK_ARG(!stats || out_mode == 0, "pp_conv_igemm: channel statistics need the bf16 NHWC output mode");
  PP_CHECK_ARG(relu >= 0 && relu <= 2, "pp_conv_igemm: relu=%d (0 none, 1 ReLU, 2 ReLU6)", relu);
  PP_CHECK_ARG(!res || (ld_res % 8 == 0 && ld_res >= Cout && (reinterpret_cast<uintptr_t>(res) % 16) == 0),
               "pp_conv_igemm: residual needs ld_res >= Cout, a multiple of 8, and a 16-byte aligned base");
  PP_CHECK_ARG(N > 0 && H > 0 && W > 0, "pp_conv_igemm: bad shape");
  PP_CHECK_ARG(Cin > 0 && Cin % 64 == 0, "pp_conv_igemm: Cin=%d must be a multiple of 64 (pad the buffer)", Cin);
  // a_channels < Cin is allowed for a single channel group: TMA zero-fills the K padding of a tensor whose channel
  // count is not a multiple of 64 (MobileNetV2: 16, 24, 32, 96, 144, ...), so activations need no padded copies
  PP_CHECK_ARG(a_channels > 0 && ld_in >= a_channels && ld_in % 8 == 0,
               "pp_conv_igemm: ld_in=%d must be >= the A channel count %d and a multiple of 8", ld_in, a_channels);
  PP_CHECK_ARG(n_entries >= 1 && n_entries <= kMaxTaps && tap_dy && tap_dx && tap_c0, "pp_conv_igemm: %d tap entries (1..%d)",
               n_entries, kMaxTaps);
  for (int t = 0; t < n_entries; ++t)
    PP_CHECK_ARG(tap_c0[t] >= 0 && tap_c0[t] % 8 == 0 && (tap_c0[t] + Cin <= a_channels || tap_c0[t] == 0) && tap_dy[t] > -8192 &&
                     tap_dy[t] < 8192 && tap_dx[t] > -8192 && tap_dx[t] < 8192,
                 "pp_conv_igemm: bad tap entry %d (dy %d dx %d c0 %d)", t, tap_dy[t], tap_dx[t], tap_c0[t]);
  const int taps = n_entries;
  PP_CHECK_ARG(Cout > 0 && Cout <= Cout_pad, "pp_conv_igemm: Cout=%d Cout_pad=%d", Cout, Cout_pad);
  PP_CHECK_ARG(out_mode == 0 || out_mode == 1, "pp_conv_igemm: out_mode=%d", out_mode);
  PP_CHECK_ARG((reinterpret_cast<uintptr_t>(x) % 16) == 0 && (reinterpret_cast<uintptr_t>(w_packed) % 16) == 0,
               "pp_conv_igemm: x / w must be 16-byte aligned");
  if (out_mode == 0)
    PP_CHECK_ARG(ld_out % 8 == 0 && c_off % 8 == 0 && (reinterpret_cast<uintptr_t>(out) % 16) == 0,
                 "pp_conv_igemm: bf16 output needs ld_out, c_off multiples of 8 and a 16-byte aligned base");
  int BN = block_n;
  int n_main = Cout_pad;  // channels covered by the first launch; a remainder gets a second, narrower launch
  if (BN == 0) {
    if (Cout_pad > 256 && Cout_pad % 256 != 0 && Cout_pad % 64 == 0) {
      // e.g. the data gradient into the 320-wide decoder input: 256-wide tiles + one 64/128-wide remainder
      // (N=64 tiles alone are shared-memory-bandwidth bound: (128+N)*32 B per N/2 cycles)
      BN = 256;
      n_main = Cout_pad / 256 * 256;
    } else if (Cout_pad % 256 == 0) BN = 256;
    else if (Cout_pad % 128 == 0) BN = 128;
    else if (Cout_pad % 64 == 0) BN = 64;
    else BN = 32;
    // small problems: halve N tiles until the grid can fill the machine
    const int th = 8, tw = 16;
    long tiles = (long)N * ((H + th - 1) / th) * ((W + tw - 1) / tw);
    while (n_main == Cout_pad && BN > 64 && tiles * (Cout_pad / BN) < sm_count_cached()) BN >>= 1;
  }
  PP_CHECK_ARG((BN == 32 || BN == 64 || BN == 128 || BN == 256) && n_main % BN == 0,
               "pp_conv_igemm: block_n=%d does not divide Cout_pad=%d", BN, Cout_pad);
  ConvParams p;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout_pad = Cout_pad; p.Cout = Cout; p.taps = taps;
  for (int t = 0; t < kMaxTaps; ++t) {
    p.tdy[t] = (short)(t < taps ? tap_dy[t] : 0);
    p.tdx[t] = (short)(t < taps ? tap_dx[t] : 0);
    p.ta[t] = (short)(t < taps ? tap_c0[t] : 0);
  }
  p.TW = 16; p.TH = 8;
  if (W <= 8) { p.TW = 8; p.TH = 16; }
  p.tiles_y = (H + p.TH - 1) / p.TH;
  p.tiles_x = (W + p.TW - 1) / p.TW;
  p.pre_bias = pre_bias; p.scale = scale; p.shift = shift; p.relu = relu;
  p.out_mode = out_mode; p.out = out; p.ld_out = ld_out; p.c_off = c_off;
  p.res = reinterpret_cast<const __nv_bfloat16*>(res); p.ld_res = ld_res;
  p.n_off = 0;
  p.n_tiles_n = n_main / BN;
  p.stats = stats;
  p.tma_store = (out_mode == 0 && (stats || conv_tma_store_enabled())) ? 1 : 0;

  CUtensorMap tmA, tmB, tmO;
  if (p.tma_store) {
    // output view: channels [c_off, c_off + Cout) of the ld_out-wide NHWC tensor; one box = 32 channels x the 32 pixels of a
    // TMEM lane quarter (TW x 32/TW), 64-byte swizzle.  Channels >= Cout and pixels outside the image are clipped by TMA.
    const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    const uint64_t strides[3] = {(uint64_t)ld_out * 2, (uint64_t)W * ld_out * 2, (uint64_t)H * W * ld_out * 2};
    const uint32_t box[4] = {32, (uint32_t)p.TW, (uint32_t)(32 / p.TW), 1};
    int rc = make_tmap_bf16_sw(&tmO, reinterpret_cast<const __nv_bfloat16*>(out) + c_off, 4, dims, strides, box, 64);
    if (rc != PP_OK) return rc;
  } else {
    memset(&tmO, 0, sizeof(tmO));
  }
  {
    const uint64_t dims[4] = {(uint64_t)a_channels, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    const uint64_t strides[3] = {(uint64_t)ld_in * 2, (uint64_t)W * ld_in * 2, (uint64_t)H * W * ld_in * 2};
    const uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, 1};
    int rc = make_tmap_bf16(&tmA, x, 4, dims, strides, box);
    if (rc != PP_OK) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout_pad, (uint64_t)taps};
    const uint64_t strides[2] = {(uint64_t)Cin * 2, (uint64_t)Cout_pad * Cin * 2};
    const uint32_t box[3] = {64, (uint32_t)BN, 1};
    int rc = make_tmap_bf16(&tmB, w_packed, 3, dims, strides, box);
    if (rc != PP_OK) return rc;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int sms = sm_count_cached();
  int rc;
  switch (BN) {
    case 256: rc = launch_conv<256>(tmA, tmB, tmO, p, sms, st); break;
    case 128: rc = launch_conv<128>(tmA, tmB, tmO, p, sms, st); break;
    case 64: rc = launch_conv<64>(tmA, tmB, tmO, p, sms, st); break;
    default: rc = launch_conv<32>(tmA, tmB, tmO, p, sms, st); break;
  }
  if (rc != PP_OK || n_main == Cout_pad) return rc;
  // remainder launch: narrower tiles over output channels [n_main, Cout_pad)
  const int rem = Cout_pad - n_main;
  const int BR = (rem % 128 == 0) ? 128 : 64;
  CUtensorMap tmB2;
  {
    const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout_pad, (uint64_t)taps};
    const uint64_t strides[2] = {(uint64_t)Cin * 2, (uint64_t)Cout_pad * Cin * 2};
    const uint32_t box[3] = {64, (uint32_t)BR, 1};
    rc = make_tmap_bf16(&tmB2, w_packed, 3, dims, strides, box);
    if (rc != PP_OK) return rc;
  }
  p.n_off = n_main;
  p.n_tiles_n = rem / BR;
  return BR == 128 ? launch_conv<128>(tmA, tmB2, tmO, p, sms, st) : launch_conv<64>(tmA, tmB2, tmO, p, sms, st);
}

}  // extern "C"
