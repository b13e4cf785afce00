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
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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
};

template <int BN>
struct ConvCfg {
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr uint32_t A_BYTES = kBM * kBK * 2;
  static constexpr uint32_t B_BYTES = BN * kBK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;  // power of two for BN in {16..256}
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

__device__ __forceinline__ bool tap_skipped(const ConvParams& p, int tap, int y0, int x0, int& dy, int& dx) {
  dy = p.tdy[tap];
  dx = p.tdx[tap];
  return (y0 + dy + p.TH <= 0) || (y0 + dy >= p.H) || (x0 + dx + p.TW <= 0) || (x0 + dx >= p.W);
}

template <int BN>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
  using Cfg = ConvCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
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

template <int BN>
static int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, int sm_count, cudaStream_t st) {
  using Cfg = ConvCfg<BN>;
  static bool attr = false;
  if (!attr) {
    PP_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    attr = true;
  }
  const int total = p.N * p.tiles_y * p.tiles_x * p.n_tiles_n;
  const int grid = total < sm_count ? total : sm_count;
  conv_igemm_kernel<BN><<<grid, kConvThreads, Cfg::SMEM, st>>>(tmA, tmB, p);
  PP_LAUNCH_CHECK();
  return PP_OK;
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
  PP_CHECK_ARG(x && w_packed && out, "pp_conv_igemm: null pointer");
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

  CUtensorMap tmA, tmB;
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
    case 256: rc = launch_conv<256>(tmA, tmB, p, sms, st); break;
    case 128: rc = launch_conv<128>(tmA, tmB, p, sms, st); break;
    case 64: rc = launch_conv<64>(tmA, tmB, p, sms, st); break;
    default: rc = launch_conv<32>(tmA, tmB, p, sms, st); break;
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
  return BR == 128 ? launch_conv<128>(tmA, tmB2, p, sms, st) : launch_conv<64>(tmA, tmB2, p, sms, st);
}

}  // extern "C"
