// T path — weight gradient of the NHWC bf16 convolutions on tcgen05 tensor cores.
//
//   dW[tap][ci][co] = sum over pixels p of  X[p + shift(tap)][ci] * dY[p][co]
//
// GEMM view: M = input channels (128 per tile), N = output channels (BN <= 256), K = pixels.
// Both operands are stored pixel-major (NHWC), i.e. "MN-major" for this GEMM: a TMA box
// {64 ch, 8, 8, 1} lands as 64 pixel rows x 128 B with the 128-byte swizzle, which is the canonical
// MN-major UMMA layout (8-row K groups 1024 B apart = SBO, 64-channel chunks one box apart = LBO);
// the instruction descriptor carries a_major = b_major = MN.  The shifted X box is zero-filled by TMA
// outside the image (the conv's padding).  K is split over CTAs; partial tiles are reduced with
// fp32 red.global.add.v4 into dW (zeroed by the caller; rows are 16-byte aligned: Cout_pad is a multiple of 64).
#include <cstdlib>
#include "tc_common.cuh"

namespace pp {

constexpr int kWgThreads = 192;
constexpr int kPB = 8;                       // pixel block is kPB x kPB = 64 pixels (one K block)
constexpr uint32_t kBoxBytes = 64 * 64 * 2;  // one {64 ch, 8, 8} box = 8 KB

constexpr int kWgMaxTaps = 16;

struct WgradParams {
  int N, H, W;
  int Cin;            // valid input channels (rows written)
  int taps;
  // tap table: tap t pairs dY[p] with X[p + (tdy, tdx)] read at channel offset tc0 of the X tensor (tc0 != 0: the
  // space-to-depth phases of a stride-2 convolution live side by side in the channel dimension)
  short tdy[kWgMaxTaps], tdx[kWgMaxTaps], tc0[kWgMaxTaps];
  int pby, pbx;       // pixel blocks per image
  int n_ci_tiles, n_co_tiles, splits;
  float* dw;          // [taps][Cin_rows][ld_dw] fp32, Cin_rows >= n_ci_tiles * 128 or >= Cin, ld_dw = n_co_tiles * BN
  int Cin_rows, ld_dw;
};

// NX = X tiles per CTA that share ONE dY tile.  With NX = 2 a stage carries dY (BN channels) + two X tiles (2 x 128 channels)
// for two 128 x BN x 64 MMAs: 64 KB per 8.4 MFLOP instead of 48 KB per 4.2 - the single-tile form is bound by the fill rate
// of shared memory (ncu on the 512 -> 512 3x3: tensor pipe 57-62 % busy, L2 hit 82 %, DRAM 13 %), the pair form by the
// tensor pipe.  The two X tiles are consecutive entries of the list (tap, ci-tile): two input-channel tiles of one tap, or
// two taps of one channel tile when Cin <= 128.
template <int BN, int NX>
struct WgCfg {
  static constexpr uint32_t A_BYTES = NX * 2 * kBoxBytes;
  static constexpr uint32_t B_BYTES = (BN / 64) * kBoxBytes;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (int)((200u * 1024u) / STAGE_BYTES) > 6 ? 6 : (int)((200u * 1024u) / STAGE_BYTES);
  static constexpr uint32_t TMEM_COLS = (NX * BN) < 32 ? 32 : NX * BN;  // 64 .. 512, a power of two
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + 256;
};

__device__ __forceinline__ bool pb_skipped(int dy, int dx, int y0, int x0, int H, int W) {
  return (y0 + dy + kPB <= 0) || (y0 + dy >= H) || (x0 + dx + kPB <= 0) || (x0 + dx >= W);
}

template <int BN, int NX>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const WgradParams p) {
  using Cfg = WgCfg<BN, NX>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  auto sA = [&](int s, int a) { return smem_base + (uint32_t)s * Cfg::STAGE_BYTES + (uint32_t)a * 2u * kBoxBytes; };
  auto sB = [&](int s) { return smem_base + (uint32_t)s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * Cfg::STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - tc::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item: (x-tile group, co-tile, split); the x-tile group varies fastest so that CTAs resident together read the
  // same dY tile (L2 hits) - x tile = (tap, ci-tile), group g covers tiles NX*g .. NX*g + NX-1
  const int n_xtiles = p.taps * p.n_ci_tiles;
  const int n_groups = (n_xtiles + NX - 1) / NX;
  int item = blockIdx.x;
  const int group = item % n_groups;
  item /= n_groups;
  const int co_tile = item % p.n_co_tiles;
  const int split = item / p.n_co_tiles;
  const int co0 = co_tile * BN;
  int tap[NX], ci0[NX], dy[NX], dx[NX], xc0[NX];
  bool act[NX];
#pragma unroll
  for (int a = 0; a < NX; ++a) {
    const int xt = group * NX + a;
    act[a] = xt < n_xtiles;
    const int xtc = act[a] ? xt : 0;
    tap[a] = xtc / p.n_ci_tiles;
    ci0[a] = (xtc - tap[a] * p.n_ci_tiles) * 128;
    dy[a] = p.tdy[tap[a]];
    dx[a] = p.tdx[tap[a]];
    xc0[a] = p.tc0[tap[a]];
  }
  const int pb_per_img = p.pby * p.pbx;
  const int total_pb = p.N * pb_per_img;
  const int pb_lo = (int)((int64_t)total_pb * split / p.splits);
  const int pb_hi = (int)((int64_t)total_pb * (split + 1) / p.splits);

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmX);
    tc::tma_prefetch_desc(&tmDY);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      tc::mbar_init(full_bar(s), 1);
      tc::mbar_init(empty_bar(s), 1);
    }
    tc::mbar_init(tfull_bar, 1);
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

  auto decode = [&](int pb, int& img, int& y0, int& x0) {
    img = pb / pb_per_img;
    const int r = pb - img * pb_per_img;
    const int ty = r / p.pbx;
    y0 = ty * kPB;
    x0 = (r - ty * p.pbx) * kPB;
  };
  // a pixel block is skipped when the shifted X box of EVERY active tile lies outside the image (its products are zero);
  // a tile that is outside while its partner is not is loaded anyway: TMA zero-fills it
  auto skipped = [&](int y0, int x0) {
    bool sk = true;
#pragma unroll
    for (int a = 0; a < NX; ++a)
      if (act[a]) sk = sk && pb_skipped(dy[a], dx[a], y0, x0, p.H, p.W);
    return sk;
  };
  const uint32_t stage_tx = Cfg::B_BYTES + (uint32_t)((act[0] ? 1 : 0) + (NX > 1 && act[NX - 1] ? 1 : 0)) * 2u * kBoxBytes;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int pb = pb_lo; pb < pb_hi; ++pb) {
        int img, y0, x0;
        decode(pb, img, y0, x0);
        if (skipped(y0, x0)) continue;
        const int s = it % Cfg::STAGES;
        const uint32_t ph = (it / Cfg::STAGES) & 1u;
        tc::mbar_wait(empty_bar(s), ph ^ 1u);
        tc::mbar_expect_tx(full_bar(s), stage_tx);
#pragma unroll
        for (int j = 0; j < BN / 64; ++j)
          tc::tma_load_4d(sB(s) + j * kBoxBytes, &tmDY, full_bar(s), co0 + 64 * j, x0, y0, img);
#pragma unroll
        for (int a = 0; a < NX; ++a)
          if (act[a]) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
              tc::tma_load_4d(sA(s, a) + j * kBoxBytes, &tmX, full_bar(s), xc0[a] + ci0[a] + 64 * j, x0 + dx[a], y0 + dy[a], img);
          }
        ++it;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(128, BN, 1, 1);
      uint32_t it = 0, accumulate = 0;
      for (int pb = pb_lo; pb < pb_hi; ++pb) {
        int img, y0, x0;
        decode(pb, img, y0, x0);
        if (skipped(y0, x0)) continue;
        const int s = it % Cfg::STAGES;
        const uint32_t ph = (it / Cfg::STAGES) & 1u;
        tc::mbar_wait(full_bar(s), ph);
        tc::tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 64 pixels = 4 x K16; 16 pixel rows = 2048 B
          const uint64_t db = tc::umma_desc_sw128(sB(s) + k * 2048, kBoxBytes, 1024);
#pragma unroll
          for (int a = 0; a < NX; ++a)
            if (act[a]) {
              const uint64_t da = tc::umma_desc_sw128(sA(s, a) + k * 2048, kBoxBytes, 1024);
              tc::umma_bf16(tmem_base + (uint32_t)a * BN, da, db, idesc, accumulate);
            }
          accumulate = 1;
        }
        tc::umma_commit(empty_bar(s));
        ++it;
      }
      tc::umma_commit(tfull_bar);
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;  // input channel within the tile
    int n_valid = 0;
    for (int pb = pb_lo; pb < pb_hi; ++pb) {
      int img, y0, x0;
      decode(pb, img, y0, x0);
      n_valid += skipped(y0, x0) ? 0 : 1;
    }
    tc::mbar_wait(tfull_bar, 0);
    tc::tc_fence_after();
    if (n_valid > 0) {
#pragma unroll
      for (int a = 0; a < NX; ++a) {
        if (!act[a]) continue;
        const int ci = ci0[a] + row;
        float* dst = p.dw + ((size_t)tap[a] * p.Cin_rows + ci) * p.ld_dw + co0;
#pragma unroll 1
        for (int j = 0; j < BN / 32; ++j) {
          uint32_t r[32];
          tc::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)a * BN + j * 32, r);
          tc::tmem_ld_wait();
          if (ci < p.Cin) {
            // 16-byte vector reductions (red.global.add.v4.f32, sm_90+): 8 L2 operations per row chunk instead of 32
#pragma unroll
            for (int c = 0; c < 32; c += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j * 32 + c), "f"(__uint_as_float(r[c])),
                           "f"(__uint_as_float(r[c + 1])), "f"(__uint_as_float(r[c + 2])), "f"(__uint_as_float(r[c + 3]))
                           : "memory");
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, int NX>
static int launch_wgrad(const CUtensorMap& tmX, const CUtensorMap& tmDY, const WgradParams& p, cudaStream_t st) {
  using Cfg = WgCfg<BN, NX>;
  static bool attr = false;
  if (!attr) {
    PP_CUDA(cudaFuncSetAttribute(wgrad_kernel<BN, NX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    attr = true;
  }
  const int n_groups = (p.taps * p.n_ci_tiles + NX - 1) / NX;
  const int grid = n_groups * p.n_co_tiles * p.splits;
  wgrad_kernel<BN, NX><<<grid, kWgThreads, Cfg::SMEM, st>>>(tmX, tmDY, p);
  PP_LAUNCH_CHECK();
  return PP_OK;
}

}  // namespace pp

using namespace pp;

extern "C" {

int pp_conv_wgrad(const void* x, int ld_x, int Cin, const void* dy, int ld_dy, int Cout_pad, int N, int H, int W,
                  int taps, int dil, float* dw, int Cin_rows, int splits, void* stream) {
  PP_CHECK_ARG(taps == 1 || taps == 9, "pp_conv_wgrad: taps=%d", taps);
  PP_CHECK_ARG(Cout_pad == 64 || Cout_pad == 128 || Cout_pad == 256, "pp_conv_wgrad: Cout_pad=%d (64, 128 or 256)", Cout_pad);
  int tdy[9], tdx[9], tc0[9];
  for (int t = 0; t < taps; ++t) {
    tdy[t] = taps == 1 ? 0 : (t / 3 - 1) * dil;
    tdx[t] = taps == 1 ? 0 : (t % 3 - 1) * dil;
    tc0[t] = 0;
  }
  return pp_conv_wgrad_multi(x, ld_x, ld_x, Cin, dy, ld_dy, Cout_pad, N, H, W, taps, tdy, tdx, tc0, dw, Cin_rows, Cout_pad,
                             splits, stream);
}

int pp_conv_wgrad_multi(const void* x, int x_channels, int ld_x, int Cin, const void* dy, int ld_dy, int Cout, int N, int H,
                        int W, int n_entries, const int* tap_dy, const int* tap_dx, const int* tap_c0, float* dw, int Cin_rows,
                        int ld_dw, int splits, void* stream) {
  PP_CHECK_ARG(x && dy && dw && tap_dy && tap_dx && tap_c0, "pp_conv_wgrad: null pointer");
  PP_CHECK_ARG(N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "pp_conv_wgrad: bad shape");
  PP_CHECK_ARG(n_entries >= 1 && n_entries <= kWgMaxTaps, "pp_conv_wgrad: %d tap entries (1..%d)", n_entries, kWgMaxTaps);
  PP_CHECK_ARG(ld_x % 8 == 0 && ld_dy % 8 == 0 && ld_x >= x_channels && x_channels >= Cin && ld_dy >= Cout,
               "pp_conv_wgrad: bad leading dims (ld_x %d x_channels %d Cin %d ld_dy %d Cout %d)", ld_x, x_channels, Cin, ld_dy, Cout);
  for (int t = 0; t < n_entries; ++t)
    PP_CHECK_ARG(tap_c0[t] >= 0 && tap_c0[t] % 8 == 0 && tap_c0[t] + Cin <= x_channels && tap_dy[t] > -8192 && tap_dy[t] < 8192 &&
                     tap_dx[t] > -8192 && tap_dx[t] < 8192,
                 "pp_conv_wgrad: bad tap entry %d (dy %d dx %d c0 %d)", t, tap_dy[t], tap_dx[t], tap_c0[t]);
  // output-channel tile: the widest that the valid channel count fills reasonably (ld_dw tells how many were allocated)
  const int BN = Cout > 128 ? 256 : (Cout > 64 ? 128 : 64);
  const int n_co_tiles = (Cout + BN - 1) / BN;
  PP_CHECK_ARG(ld_dw >= n_co_tiles * BN && ld_dw % 4 == 0, "pp_conv_wgrad: ld_dw=%d must be >= %d (Cout rounded up to %d)", ld_dw,
               n_co_tiles * BN, BN);
  const int n_ci_tiles = (Cin + 127) / 128;
  PP_CHECK_ARG(Cin_rows >= Cin, "pp_conv_wgrad: Cin_rows=%d too small", Cin_rows);
  PP_CHECK_ARG((reinterpret_cast<uintptr_t>(x) % 16) == 0 && (reinterpret_cast<uintptr_t>(dy) % 16) == 0 &&
                   (reinterpret_cast<uintptr_t>(dw) % 16) == 0,
               "pp_conv_wgrad: x / dy / dw must be 16-byte aligned");
  WgradParams p;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.taps = n_entries;
  for (int t = 0; t < kWgMaxTaps; ++t) {
    p.tdy[t] = (short)(t < n_entries ? tap_dy[t] : 0);
    p.tdx[t] = (short)(t < n_entries ? tap_dx[t] : 0);
    p.tc0[t] = (short)(t < n_entries ? tap_c0[t] : 0);
  }
  p.pby = (H + kPB - 1) / kPB;
  p.pbx = (W + kPB - 1) / kPB;
  p.n_ci_tiles = n_ci_tiles;
  p.n_co_tiles = n_co_tiles;
  const int total_pb = N * p.pby * p.pbx;
  // two X tiles per CTA (sharing the dY tile) whenever there are at least two of them
  static int force_nx = -1;
  if (force_nx < 0) {
    const char* e = getenv("PP_WGRAD_NX");  // "1": the single-tile form (A/B measurements)
    force_nx = (e && e[0] == '1') ? 1 : 0;
  }
  const int NX = (n_entries * n_ci_tiles >= 2 && !force_nx) ? 2 : 1;
  const int n_groups = (n_entries * n_ci_tiles + NX - 1) / NX;
  if (splits <= 0) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int items = n_groups * n_co_tiles;
    // whole waves: the largest split count that keeps the grid within one wave if the items allow it, else two
    splits = sms / items;
    if (splits < 1) splits = (2 * sms) / items;
    const int max_splits = (total_pb + 7) / 8;  // at least 8 pixel blocks (512 pixels) per CTA
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  if (splits > total_pb) splits = total_pb;
  p.splits = splits;
  p.dw = dw;
  p.Cin_rows = Cin_rows;
  p.ld_dw = ld_dw;
  CUtensorMap tmX, tmDY;
  {
    const uint64_t dims[4] = {(uint64_t)x_channels, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    const uint64_t strides[3] = {(uint64_t)ld_x * 2, (uint64_t)W * ld_x * 2, (uint64_t)H * W * ld_x * 2};
    const uint32_t box[4] = {64, kPB, kPB, 1};
    int rc = make_tmap_bf16(&tmX, x, 4, dims, strides, box);
    if (rc != PP_OK) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    const uint64_t strides[3] = {(uint64_t)ld_dy * 2, (uint64_t)W * ld_dy * 2, (uint64_t)H * W * ld_dy * 2};
    const uint32_t box[4] = {64, kPB, kPB, 1};
    int rc = make_tmap_bf16(&tmDY, dy, 4, dims, strides, box);
    if (rc != PP_OK) return rc;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (NX == 2) {
    switch (BN) {
      case 256: return launch_wgrad<256, 2>(tmX, tmDY, p, st);
      case 128: return launch_wgrad<128, 2>(tmX, tmDY, p, st);
      default: return launch_wgrad<64, 2>(tmX, tmDY, p, st);
    }
  }
  switch (BN) {
    case 256: return launch_wgrad<256, 1>(tmX, tmDY, p, st);
    case 128: return launch_wgrad<128, 1>(tmX, tmDY, p, st);
    default: return launch_wgrad<64, 1>(tmX, tmDY, p, st);
  }
}

}  // extern "C"
