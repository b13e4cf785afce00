/*
 * pixelpick_b200 — C-ABI of the B200-native hot paths of PixelPick.
 *
 * The reference (NoelShin/PixelPick) is pure Python and has NO FFI: its boundary is Python
 * duck-typing.  Every entry point below therefore cites the reference Python
 * call site it replaces; the Python host (`pixelpick_b200/query.py`, `pixelpick_b200/loss.py`, …)
 * mirrors those classes and binds this library through ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; device pointers unless the name says `_host`
 *   - caller owns all memory; the only hidden allocations are inside `pp_acq_session_*`
 *   - every device call takes a `cudaStream_t` passed as `void*` (0 = legacy default stream)
 *   - return 0 on success, negative PP_ERR_* otherwise; `pp_last_error()` gives the text
 *   - thread-compatible: one thread per stream/session at a time
 */
#ifndef PIXELPICK_B200_H_
#define PIXELPICK_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PP_OK 0
#define PP_ERR_INVALID_ARG (-1)
#define PP_ERR_CUDA (-2)
#define PP_ERR_WORKSPACE (-3)
#define PP_ERR_UNSUPPORTED (-4)

/* dtypes of logits / feature maps */
#define PP_F32 0
#define PP_BF16 1

/* acquisition strategies — reference `UncertaintySampler` (query.py:224-247) */
#define PP_STRAT_ENTROPY 0          /* query.py:229-230  sum_c -p log p            (top-k: largest)  */
#define PP_STRAT_LEAST_CONFIDENCE 1 /* query.py:233-234  1 - max_c p              (top-k: largest)  */
#define PP_STRAT_MARGIN 2           /* query.py:237-239  |p_(1) - p_(2)|  (BvSB)  (top-k: smallest) */

int pp_version(void);
const char* pp_last_error(void);
/* number of kernels this library has launched so far in this process (bench.py: gpu_launches) */
long long pp_launch_count(void);
/* number of SMs / device name of the current device: used by bench.py for the grid/roofline note */
/* Host-side (no GPU needed) evaluation of the select's two ordering maps, exported for property tests:
 * pp_host_ord_key  = the uint32 whose ascending order is the selection order (descending score for largest, NaN first;
 *                    ascending score otherwise, NaN last; -0.0 == +0.0);
 * pp_host_bucket0  = the level-0 bucket of the radix select (0..2047), monotone non-decreasing in pp_host_ord_key. */
unsigned int pp_host_ord_key(float score, int largest);
unsigned int pp_host_bucket0(float score, int largest);
int pp_device_info(int* sm_count, int* cc_major, int* cc_minor, char* name, int name_len);

/* ------------------------------------------------------------------------------------------
 * Q path: per-pixel acquisition score.
 * Replaces  prob = softmax(model(x)["pred"][:, :, :h, :w], dim=1)           query.py:190
 *           uc_map = UncertaintySampler(strategy)(prob)                      query.py:192,229-239
 *           uc_map[mask] = fill ; uc_map[mask_void] = fill                   query.py:195-201
 *           (reverse_order) uc_map[~sampling_mask] = fill                    query.py:43-48
 * fill = 0.0 for entropy / least-confidence, 1.0 for margin (query.py:198).
 *
 * logits: [n_img, C, H, W] viewed through element strides (stride_w == 1); a sliced view of a
 *         padded forward (`[:, :, :h, :w]`) is expressed with stride_h > W.
 * labelled / void_mask / keep: uint8 [n_img, H, W] contiguous, each may be NULL.
 *         score <- fill where labelled!=0 or void_mask!=0 or keep==0.
 * score_map: float32 [n_img, H*W] contiguous (required).
 * hist0: optional uint32 [n_img, 2048], ZEROED by the caller: receives the level-0 radix histogram
 *        of the ordering key so `pp_acq_topk` can skip its first pass (pass NULL to skip).
 * ------------------------------------------------------------------------------------------ */
int pp_acq_score(const void* logits, int dtype, int n_img, int C, int H, int W,
                 int64_t stride_n, int64_t stride_c, int64_t stride_h,
                 const uint8_t* labelled, const uint8_t* void_mask, const uint8_t* keep,
                 int strategy, float* score_map, uint32_t* hist0, void* stream);

/* Fused variant: the DeepLab head produces logits at 1/4 resolution and the model's last op is a
 * bilinear align_corners=True upsample to the input size (deeplab.py:55).  This entry point reads
 * the LOW-RES logits [n_img, C, h_in, w_in] (contiguous NCHW f32) and evaluates the upsample
 * on the fly, so the full-resolution logits are never written to HBM.  Same outputs as pp_acq_score. */
int pp_acq_score_upsampled(const float* logits_lowres, int n_img, int C, int h_in, int w_in,
                           int H, int W,
                           const uint8_t* labelled, const uint8_t* void_mask, const uint8_t* keep,
                           int strategy, float* score_map, uint32_t* hist0, void* stream);

/* pp_acq_score + pp_acq_select in ONE pass over the logits (query.py:190-201 softmax / uncertainty / mask fills, then the
 * selection half of uc_map.topk(k) query.py:57-61): a thread-block cluster per image keeps the scores in shared memory, merges
 * the level-0 histograms through distributed shared memory and classifies its own scores - the score map is neither written
 * nor re-read (score_map may be NULL; when given it is also written).  Leaves the k unsorted candidates in the workspace
 * exactly like pp_acq_select; follow with pp_acq_pick / pp_acq_topk's sort.  The workspace must have been prepared
 * (pp_acq_topk_prepare).  Covers f32 logits, W % 4 == 0, C in {11, 19, 21}, H*W = cluster x (a multiple of 4096 <= 16384)
 * with cluster <= 16 (e.g. 256x512); other shapes return PP_ERR_UNSUPPORTED: use the two-call form. */
int pp_acq_score_select(const void* logits, int dtype, int n_img, int C, int H, int W, int64_t stride_n, int64_t stride_c,
                        int64_t stride_h, const uint8_t* labelled, const uint8_t* void_mask, const uint8_t* keep, int strategy,
                        int k, float* score_map, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Q path: per-image sorted top-k of a score map.
 * Replaces  uc_map.flatten().topk(k, largest=strategy in {entropy, LC}).indices   query.py:57-61
 * Order contract: sorted by score (descending if largest else ascending), NaN ranks
 * as the largest value (torch semantics), -0.0 == +0.0, ties broken by LOWER flat index first.
 *
 * topk_idx: int32 [n_img, k]; topk_val: optional float32 [n_img, k].
 * hist0_valid != 0 says the workspace histogram was already filled by pp_acq_score(hist0=...),
 * where hist0 must be the pointer returned by pp_acq_topk_hist0(workspace).
 * ------------------------------------------------------------------------------------------ */
int pp_acq_topk_workspace_bytes(int n_img, int HW, int k, size_t* out_bytes);
/* Zeroes the counters/histograms inside the workspace; must run (on the same stream) before the FIRST
 * pp_acq_score(hist0=...) / pp_acq_topk / pp_acq_select / pp_acq_score_select on it.  A completed select (pp_acq_topk,
 * pp_acq_select, or pp_acq_score_select) leaves the workspace prepared again for the same (n_img, HW, k) - the kernels zero what
 * they consumed - so a host that tracks this may skip the call between batches; calling it every batch stays correct.  A
 * pp_acq_score(hist0=...) that is NOT followed by a select leaves the histogram filled: prepare again before reusing it. */
int pp_acq_topk_prepare(void* workspace, size_t workspace_bytes, int n_img, int HW, int k, void* stream);
uint32_t* pp_acq_topk_hist0(void* workspace);
int pp_acq_topk(const float* score_map, int n_img, int HW, int k, int largest, int hist0_valid,
                int32_t* topk_idx, float* topk_val,
                void* workspace, size_t workspace_bytes, void* stream);

/* The same selection WITHOUT the sort, for callers that only need a few ranks of the sorted list — which is all
 * `np.random.choice(ind_queries, n, False)` (query.py:63-64) reads.  pp_acq_select leaves the k selected (unsorted)
 * composites in the workspace; pp_acq_pick returns out[i, j] = flat index of the element of rank pos[i, j] (0-based, in
 * the order contract above; pos == NULL means ranks 0..n-1) by a radix walk over the k candidates: O(k), bit-identical
 * to topk_idx[i, pos[i, j]]. */
int pp_acq_select(const float* score_map, int n_img, int HW, int k, int largest, int hist0_valid, void* workspace,
                  size_t workspace_bytes, void* stream);
int pp_acq_pick(void* workspace, size_t workspace_bytes, int n_img, int HW, int k, const int32_t* pos, int n,
                int32_t* out, void* stream);

/* out[i, j] = topk_idx[i, pos[i, j]] — the device half of
 *   np.random.choice(ind_queries, n_pixels_by_us, False)                       query.py:63-64
 * (the host draws pos = np.random.permutation(k)[:n] from the global NumPy stream).  pos lives on the device, so it
 * cannot be validated by the call: values outside [0, k) are clamped (here and in pp_acq_pick), never read out of bounds. */
int pp_acq_gather(const int32_t* topk_idx, int n_img, int k, const int32_t* pos, int n,
                  int32_t* out, void* stream);

/* entropy of softmax(logits) at given flat pixel indices — the device half of
 * QueryStats._get_entropy (query.py:260-264), evaluated only at the selected pixels.  Indices outside [0, H*W) are clamped. */
int pp_acq_entropy_at(const void* logits, int dtype, int n_img, int C, int H, int W,
                      int64_t stride_n, int64_t stride_c, int64_t stride_h,
                      const int32_t* px_idx, int n, float* out, void* stream);

/* same, reading the 1/4-resolution head logits through the on-the-fly bilinear upsample */
int pp_acq_entropy_at_upsampled(const float* logits_lowres, int n_img, int C, int h_in, int w_in,
                                int H, int W, const int32_t* px_idx, int n, float* out, void* stream);

/* QueryStats.update at the picks (query.py:296-308) + the wire-format coordinates of encode_query (query.py:72-87), on the
 * device: for every image's n picks (flat indices sorted ascending = np.where order) -> x_coords / y_coords (int64 [n_img][n]),
 * the labels at the picks (labels: uint8 [n_img][HW] or NULL), their histogram (label_hist[n_classes] += , _count_labels), the
 * number of distinct labels per image and the spatial coverage = mean pairwise distance over the n (n - 1) ordered pairs in
 * float64 with NumPy's pairwise summation order (bit-identical to np.mean; NaN for n < 2). */
int pp_query_stats_at(const long long* sel_sorted, int n_img, int n, int W, int HW, const uint8_t* labels, int n_classes,
                      long long* x_coords, long long* y_coords, int32_t* labels_at, long long* label_hist, int32_t* n_unique,
                      double* coverage, void* stream);

/* ------------------------------------------------------------------------------------------
 * Q path through HOST buffers (the call a non-PyTorch host makes; bench.py's `e2e`).
 * A session owns device staging buffers, pinned host buffers and two streams; `run_host` copies
 * the logits and masks H2D in image chunks (double-buffered against the kernels), runs
 * score -> top-k -> gather and copies the selected indices back.
 * ------------------------------------------------------------------------------------------ */
typedef struct pp_acq_session pp_acq_session;
int pp_acq_session_create(pp_acq_session** out, int chunk_imgs, int C, int H, int W, int k, int n_sel);
int pp_acq_session_destroy(pp_acq_session* s);
/* h_logits: float32 [n_img, C, H, W]; h_labelled/h_void: uint8 [n_img, H, W] or NULL;
 * h_pos: int32 [n_img, n_sel] positions in the sorted top-k list (NULL => first n_sel);
 * h_sel_idx: int32 [n_img, n_sel] (out); h_topk_idx: optional int32 [n_img, k] (out, may be NULL).
 * Host pointers may be pageable; pinned memory (cudaHostAlloc / torch pin_memory) is faster. */
/* Split form for hosts that draw the pick positions themselves (query.py:63-64): begin launches H2D + scoring + select
 * and returns immediately, the host draws its np.random.permutation positions while the logits cross PCIe, finish uploads
 * them, picks and returns the selected pixel indices.  n_img <= 2 * chunk_imgs; one begin pending at a time. */
int pp_acq_session_begin_host(pp_acq_session* s, const float* h_logits, const uint8_t* h_labelled, const uint8_t* h_void,
                              int n_img, int strategy);
int pp_acq_session_finish_host(pp_acq_session* s, const int32_t* h_pos, int32_t* h_sel_idx);
int pp_acq_session_run_host(pp_acq_session* s, const float* h_logits, const uint8_t* h_labelled,
                            const uint8_t* h_void, int n_img, int strategy,
                            const int32_t* h_pos, int32_t* h_sel_idx, int32_t* h_topk_idx);

/* ------------------------------------------------------------------------------------------
 * T path: sparse-pixel cross entropy.
 * Replaces  y.flatten()[~mask.flatten()] = ignore_index                        model.py:108-110
 *           F.cross_entropy(logits, y, ignore_index)  (+ its backward)         model.py:116,120-121
 *           pred = F.interpolate(head_logits, size, 'bilinear', align_corners=True)  deeplab.py:55
 * The labelled pixels are given as a list (img, flat index, label); the kernel gathers the four
 * low-res neighbours x C, applies the bilinear weights, log-softmax and NLL, and scatter-adds
 * d(loss)/d(logits_lowres).  Mean over labelled pixels (NaN when n_px == 0, as the reference).
 *
 * logits_lowres: float32 [n_img, C, h_in, w_in] NCHW contiguous (h_in==H, w_in==W => no upsample).
 * px_img/px_idx/px_label: int32 [n_px].  loss: float32 [1].  grad_lowres: float32 same shape as
 * logits_lowres, ZEROED by the caller (may be NULL for forward only).  pred_at: optional int32
 * [n_px] argmax class at each labelled pixel (train-time running metrics, model.py:124).
 * n_px_dev: optional device int32: the number of valid list entries (<= n_px = list capacity), so a CUDA graph
 * captured once can be replayed with a different number of labelled pixels.
 * ------------------------------------------------------------------------------------------ */
int pp_sparse_ce(const float* logits_lowres, int n_img, int C, int h_in, int w_in, int H, int W,
                 const int32_t* px_img, const int32_t* px_idx, const int32_t* px_label, int n_px,
                 const int32_t* n_px_dev, float grad_scale, float* loss, float* grad_lowres, int32_t* pred_at,
                 void* stream);

/* Bilinear resize, align_corners=True (deeplab.py:49,55,58; aspp.py:70), NCHW float32,
 * forward and its adjoint (grad_in must be zeroed by the caller). */
int pp_upsample_bilinear_ac(const float* in, int n_img, int C, int h_in, int w_in,
                            float* out, int H, int W, void* stream);
int pp_upsample_bilinear_ac_bwd(const float* grad_out, int n_img, int C, int H, int W,
                                float* grad_in, int h_in, int w_in, void* stream);

/* ------------------------------------------------------------------------------------------
 * T path: NHWC bf16 implicit-GEMM convolution on tcgen05 tensor cores (stride 1, padding == dilation,
 * 1x1 or 3x3).  Replaces every nn.Conv2d of the DeepLabv3+ head — ASPP branches and projection
 * (aspp.py:49-52,73), SegmentHead convs and classifier (decoders.py:107-116) — with the following
 * BatchNorm(eval)/bias + ReLU folded into the epilogue:  out = relu?((acc + pre_bias[n]) * scale + shift).
 *
 * x: bf16 [N, H, W, ld_in] (first Cin channels used; Cin % 64 == 0).
 * w_packed: bf16 [taps][Cout_pad][Cin] (tap = ky*3+kx; rows >= Cout are zero).
 * pre_bias: f32 [N][Cout_pad] or NULL; scale/shift: f32 [Cout_pad] or NULL.
 * out_mode 0: bf16 NHWC written at out[pixel * ld_out + c_off + c];  1: f32 NCHW [N, Cout, H, W].
 * block_n: 0 = auto, else 32/64/128/256 (must divide Cout_pad).
 * The data gradient of the same convolution is this call with w_packed = flipped/transposed weights.
 * ------------------------------------------------------------------------------------------ */
int pp_conv_igemm(const void* x, int N, int H, int W, int Cin, int ld_in, const void* w_packed, int taps, int dil,
                  int Cout_pad, int Cout, const float* pre_bias, const float* scale, const float* shift, int relu,
                  void* out, int out_mode, int ld_out, int c_off, int block_n, void* stream);

/* Generalised form: K runs over n_entries tap entries; entry t reads the A tile shifted by (tap_dy[t], tap_dx[t])
 * pixels from channels [tap_c0[t], tap_c0[t] + Cin) of x (a_channels wide) and multiplies weight slice t of
 * w_packed [n_entries][Cout_pad][Cin].  One launch computes e.g. the data gradient of all four ASPP branches
 * (aspp.py:49-52,64-68: 1 + 9 + 9 + 9 taps with their own dilations) accumulated in TMEM, no partial sums in HBM.
 * Host arrays; n_entries <= 28.  Extras over pp_conv_igemm: relu = 2 is ReLU6; res (bf16 [pixel][ld_res]) is a residual
 * added before the activation (the inverted-residual / bottleneck skip, mobilenet_v2.py:66, resnet_models.py:88-92);
 * a_channels may be smaller than the K-padded Cin for a single channel group (TMA zero-fills the K padding). */
int pp_conv_igemm_multi(const void* x, int N, int H, int W, int a_channels, int ld_in, int Cin, const void* w_packed,
                        int n_entries, const int* tap_dy, const int* tap_dx, const int* tap_c0, int Cout_pad, int Cout,
                        const float* pre_bias, const float* scale, const float* shift, int relu, const void* res,
                        int ld_res, void* out, int out_mode, int ld_out, int c_off, int block_n, void* stream);

/* pp_conv_igemm_multi + per-channel statistics of the output, for TRAIN-mode BatchNorm (replaces the reduction pass of
 * nn.BatchNorm2d over the conv output: mobilenet_v2.py:7-12,34-58, resnet_models.py:74-94, aspp.py:16-20): stats[0][c] +=
 * sum, stats[1][c] += sum of squares (over every pixel of this launch, of the bf16-rounded values written to `out`);
 * stats = float [2][Cout], zeroed by the caller, may be NULL.  Output mode 0 only. */
int pp_conv_igemm_stats(const void* x, int N, int H, int W, int a_channels, int ld_in, int Cin, const void* w_packed,
                        int n_entries, const int* tap_dy, const int* tap_dx, const int* tap_c0, int Cout_pad, int Cout,
                        const float* pre_bias, const float* scale, const float* shift, int relu, const void* res,
                        int ld_res, void* out, int out_mode, int ld_out, int c_off, int block_n, float* stats, void* stream);

/* Epilogue of the bf16 NHWC output mode of pp_conv_igemm*: 1 (default) = shared-memory staging + TMA store, 0 = direct
 * register -> global stores (the round-1 path; launches that take statistics always stage).  Returns the previous setting.
 * Process-wide, for A/B measurements and tests; the environment variable PP_CONV_TMA_STORE=0 sets the initial value. */
int pp_conv_set_epilogue(int tma_store);

/* Weight gradient of the same convolution on tcgen05 (both operands MN-major, split over pixels):
 *   dw[tap][ci][co] += sum_p x[p + shift(tap)][ci] * dy[p][co]
 * x: bf16 [N,H,W,ld_x] (Cin valid channels); dy: bf16 [N,H,W,ld_dy] (first Cout_pad channels, 64/128/256);
 * dw: f32 [taps][Cin_rows][Cout_pad], ZEROED by the caller (Cin_rows >= Cin); splits: 0 = auto. */
int pp_conv_wgrad(const void* x, int ld_x, int Cin, const void* dy, int ld_dy, int Cout_pad, int N, int H, int W,
                  int taps, int dil, float* dw, int Cin_rows, int splits, void* stream);

/* General form of pp_conv_wgrad for the ENCODER convolutions (mobilenet_v2.py:34-58 1x1 expansions / projections,
 * resnet_models.py:74-94 bottleneck 1x1 / 3x3 / downsample): any Cout (tiled by 64 / 128 / 256 output channels, the tensor
 * map zero-fills the ragged tail), a tap table like pp_conv_igemm_multi's (entry t pairs dY[p] with x[p + (dy, dx)] read at
 * channel offset c0 - the four space-to-depth phases of a stride-2 convolution sit side by side in x's channels), and an
 * explicit row pitch of dw.  dw = float [n_entries][Cin_rows][ld_dw], zeroed by the caller, ld_dw >= Cout rounded up to
 * the tile width (64 if Cout <= 64, 128 if <= 128, else 256). */
int pp_conv_wgrad_multi(const void* x, int x_channels, int ld_x, int Cin, const void* dy, int ld_dy, int Cout, int N, int H,
                        int W, int n_entries, const int* tap_dy, const int* tap_dx, const int* tap_c0, float* dw, int Cin_rows,
                        int ld_dw, int splits, void* stream);

/* nn.MaxPool2d(kernel_size=3, stride=2, padding=1) of the ResNet stem (resnet_models.py:116, resnet_backbone.py:56) on NHWC bf16:
 * y [N][Ho][Wo][C] with Ho = (H - 1) / 2 + 1; `code` (optional, one byte per output element) = position 3 * ky + kx of the
 * maximum inside its window (first maximum in row-major order, NaN wins: ATen's rule).  Backward: dx [N][H][W][C] gathers, for
 * every input pixel, the gradients of the windows whose code points at it (no atomics, no zero-fill).  C % 8 == 0. */
int pp_maxpool3x3s2_fwd(const void* x, int N, int H, int W, int C, void* y, unsigned char* code, void* stream);
int pp_maxpool3x3s2_bwd(const void* dy, const unsigned char* code, int N, int H, int W, int C, void* dx, void* stream);

/* model.py:121 `self.optimizer.step()` with the Adam that utils/utils.py:112-141 builds for `cs` (torch.optim.Adam: L2 weight
 * decay folded into the gradient, bias-corrected moments, no amsgrad): the update of ALL n_tensors fp32 parameter tensors in one
 * launch.  params / grads / exp_avg / exp_avg_sq: HOST arrays of n_tensors DEVICE pointers (tensor i holds numel[i] floats and
 * belongs to parameter group group[i] < n_groups <= 8); lr[g] = DEVICE fp32 scalar of group g (a scheduler may rewrite it between
 * CUDA-graph replays); beta1 / beta2 / eps / weight_decay: HOST arrays of n_groups doubles; step = DEVICE fp32 scalar holding
 * the step count t >= 1 of THIS update (the caller increments it first, as torch does).  Per element, in fp32 with the
 * hyper-parameters rounded to fp32 and the fused multiply-adds of ATen's fused_adam_utils.cuh:
 *   g += wd * p;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps). */
int pp_adam_step_multi(int n_tensors, void* const* params, const void* const* grads, void* const* exp_avg,
                       void* const* exp_avg_sq, const long long* numel, const int* group, int n_groups,
                       const float* const* lr, const double* beta1, const double* beta2, const double* eps,
                       const double* weight_decay, const float* step, void* stream);

/* pp_bn_finalize + pp_bn_apply(_res) in ONE launch, for per-channel sums that already exist (pp_conv_igemm_stats): every thread
 * derives scale / shift of its 8 channels from sums [2][C] (same arithmetic as pp_bn_finalize), block 0 writes stats_out
 * [4][C] = (scale, shift, mean, rstd) for pp_bn_bwd and updates running_mean / running_var (may be NULL) as nn.BatchNorm2d
 * does (momentum, unbiased variance).  out = act(raw * scale + shift [+ res]). */
int pp_bn_apply_stats(const void* raw, int64_t M, int ld_in, int c_off_in, int C, const float* sums, const float* gamma,
                      const float* beta, float eps, float momentum, float* running_mean, float* running_var, float* stats_out,
                      int relu, const void* res, int ld_res, void* out, int ld_out, int c_off_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * T path: the HBM-bound layers between the convolutions, NHWC bf16 (channel counts multiples of 8).
 * Replaces nn.BatchNorm2d (train mode) + nn.ReLU + nn.Dropout (aspp.py:16-20,73-79; decoders.py:107-114;
 * deeplab.py:24-26) and the decoder-input upsample + concat (deeplab.py:49-50).
 * ------------------------------------------------------------------------------------------ */
/* sums[0][c] = sum_rows raw[:, c_off+c], sums[1][c] = sum of squares (f32 [2][C], zeroed inside). */
int pp_bn_stats(const void* raw, int64_t M, int ld, int c_off, int C, float* sums, void* stream);
/* (sum, sum sq) -> out[4][Cpad] = (scale = gamma*rstd, shift = beta - mean*scale, mean, rstd), zero for c >= C;
 * updates running_mean / running_var in place with nn.BatchNorm2d semantics when non-NULL. */
int pp_bn_finalize(const float* sums, int C, int64_t M, float eps, float momentum, const float* gamma,
                   const float* beta, float* running_mean, float* running_var, float* out, int Cpad, void* stream);
/* relu: 0 none, 1 ReLU, 2 ReLU6.  out[:, c_off_out+c] = dropout_p(relu?(raw[:, c_off_in+c] * scale[c] + shift[c])); Philox mask keyed by
 * (seed + *seed_dev, offset, element index) so the backward regenerates it; seed_dev (optional device uint64) lets a
 * captured CUDA graph draw fresh masks on every replay. */
int pp_bn_apply(const void* raw, int64_t M, int ld_in, int c_off_in, int C, const float* scale, const float* shift,
                int relu, float drop_p, uint64_t seed, uint64_t offset, const uint64_t* seed_dev, void* out, int ld_out,
                int c_off_out, void* stream);
/* BatchNorm(train)+ReLU+Dropout backward: draw = scale * (g - mean(g) - xhat * mean(g * xhat)) with
 * g = dy * dropmask/(1-p) * [raw*scale+shift > 0]; sums (f32 [2][C]) returns sum g (= d beta) and
 * sum g*xhat (= d gamma).  draw: bf16 [M][C].  g is recomputed in the second pass, never stored. */
int pp_bn_bwd(const void* dy, int ld_dy, int c_off_dy, const void* raw, int ld_raw, int c_off_raw, int64_t M, int C,
              const float* scale, const float* shift, const float* mean, const float* rstd, int relu, float drop_p,
              uint64_t seed, uint64_t offset, const uint64_t* seed_dev, float* sums, void* draw, void* stream);
/* Residual variants for the ResNet bottleneck tail out = relu(bn3(conv3(.)) + identity) (resnet_models.py:88-92):
 * res (bf16 [M][ld_res], channels 0..C) is added before the activation in ONE pass; the backward gates on
 * bn(raw) + res and also returns dres = gated upstream gradient (bf16 [M][C]) = gradient wrt the identity branch.
 * pp_bn_bwd_res writes draw into the channel slice [c_off_draw, c_off_draw + C) of a [M][ld_draw] buffer (the four ASPP
 * branches share one 1024-wide gradient buffer that pp_conv_igemm_multi then reads in one launch). */
int pp_bn_apply_res(const void* raw, int64_t M, int ld_in, int c_off_in, int C, const float* scale, const float* shift,
                    int relu, float drop_p, uint64_t seed, uint64_t offset, const uint64_t* seed_dev, const void* res,
                    int ld_res, void* out, int ld_out, int c_off_out, void* stream);
int pp_bn_bwd_res(const void* dy, int ld_dy, int c_off_dy, const void* raw, int ld_raw, int c_off_raw, int64_t M, int C,
                  const float* scale, const float* shift, const float* mean, const float* rstd, int relu, float drop_p,
                  uint64_t seed, uint64_t offset, const uint64_t* seed_dev, const void* res, int ld_res, void* dres,
                  float* sums, void* draw, int ld_draw, int c_off_draw, void* stream);
/* Single-launch train-mode BatchNorm (+activation/residual/dropout) forward and backward: both passes of the layer
 * (statistics -> normalise; reduce -> input gradient) in ONE cooperative kernel around a grid-wide barrier, instead of
 * memset + stats + finalize + apply (+ memset + reduce + apply).  Same arithmetic as the separate entry points.
 * scratch: pp_bn_scratch_bytes(C) bytes of device memory, ZERO before the first call; each call leaves it zeroed again
 * (one scratch may serve the forward and the backward of a layer, not two kernels in flight at once).
 * pp_bn_fwd_fused also writes stats_out f32 [4][C] = (scale, shift, mean, rstd), updates running_mean / running_var
 * (nn.BatchNorm2d semantics; NULL to skip) and increments *num_batches_tracked (device int64; NULL to skip). */
int pp_bn_scratch_bytes(int C, size_t* out_bytes);
int pp_bn_fwd_fused(const void* raw, int64_t M, int ld_in, int c_off_in, int C, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, long long* num_batches_tracked, float eps, float momentum,
                    int relu, float drop_p, uint64_t seed, uint64_t offset, const uint64_t* seed_dev, const void* res,
                    int ld_res, void* out, int ld_out, int c_off_out, float* stats_out, void* scratch, void* stream);
int pp_bn_bwd_fused(const void* dy, int ld_dy, int c_off_dy, const void* raw, int ld_raw, int c_off_raw, int64_t M, int C,
                    const float* scale, const float* shift, const float* mean, const float* rstd, int relu, float drop_p,
                    uint64_t seed, uint64_t offset, const uint64_t* seed_dev, const void* res, int ld_res, void* dres,
                    float* sums, void* draw, int ld_draw, int c_off_draw, void* scratch, void* stream);
/* Running train metrics on the device (replaces the per-step D2H of two full maps for RunningScore.update,
 * model.py:124-129 / utils/metrics.py:162-177): confusion[lt * n_classes + lp] += 1 for the first *n_valid_dev (or n_max)
 * (label, prediction) pairs with lt, lp in [0, n_classes); *loss_sum += *loss; *n_steps += 1.  All accumulators are
 * device memory owned and zeroed by the caller; one launch, CUDA-graph capturable. */
int pp_metrics_accumulate(const int32_t* labels, const int32_t* preds, const int32_t* n_valid_dev, int n_max,
                          int n_classes, const float* loss, long long* confusion, double* loss_sum,
                          long long* n_steps, void* stream);
/* Validation metrics in one pass (model.py:177-239 `_val`, eval.py:15-94 `evaluate`): pred = argmax over classes of the
 * x4 bilinear (align_corners=True) upsample of the 1/4-resolution head logits [n, C, h_in, w_in] (first maximum, NaN =
 * maximum, as torch.argmax); confusion[lt * C + pred] += 1 for every pixel whose label lt is in [0, C).  labels
 * [n, H, W] of label_dtype 0 int64 / 1 int32 / 2 uint8 (device); confusion: device int64 [C*C], accumulated (caller zeroes
 * it once per epoch); pred_out: optional int32 [n, H, W].  C in {11, 19, 21}. */
int pp_eval_confusion_upsampled(const float* logits_lowres, int n_img, int C, int h_in, int w_in, int H, int W,
                                const void* labels, int label_dtype, long long* confusion, int32_t* pred_out,
                                void* stream);
/* bilinear align_corners=True resize of bf16 NHWC into a channel slice, and its adjoint (gather form: deterministic,
 * grad_in f32 [N,h,w,C] fully overwritten). */
int pp_upsample_nhwc_bf16(const void* in, int N, int h, int w, int C, int ld_in, void* out, int H, int W, int ld_out,
                          int c_off, void* stream);
int pp_upsample_nhwc_bf16_bwd(const void* grad_out, int N, int H, int W, int ld, int c_off, int C, float* grad_in,
                              int h, int w, void* stream);
/* Depthwise 3x3 convolution of the MobileNetV2 inverted residual (mobilenet_v2.py:33-35,46-48: groups = channels, no
 * bias, padding 0 because fixed_padding (mobilenet_v2.py:15-21) padded the tensor explicitly), NHWC bf16, fp32
 * accumulate.  x [N][Hi][Wi][C], w = the module's f32 [C][1][3][3], y / dy [N][Ho][Wo][C] with
 * Ho = (Hi - 2*dil - 1)/stride + 1.  dgrad writes dx [N][Hi][Wi][C]; wgrad writes dw f32 [C][1][3][3]. */
int pp_dwconv3x3_fwd(const void* x, const float* w, void* y, int N, int Hi, int Wi, int C, int stride, int dil,
                     void* stream);
/* inference form: y = act(conv(x) * scale[c] + shift[c]) — eval-mode BatchNorm folded into the epilogue; act 0/1/2 =
 * none / ReLU / ReLU6 (mobilenet_v2.py:33-37) */
int pp_dwconv3x3_fwd_bnact(const void* x, const float* w, const float* scale, const float* shift, int act, void* y, int N,
                           int Hi, int Wi, int C, int stride, int dil, void* stream);
int pp_dwconv3x3_dgrad(const void* dy, const float* w, void* dx, int N, int Hi, int Wi, int C, int stride, int dil,
                       void* stream);
int pp_dwconv3x3_wgrad(const void* x, const void* dy, float* dw, int N, int Hi, int Wi, int C, int stride, int dil,
                       void* stream);
/* f32 conv weight [Cout][Cin_total][kh][kw] (first Cin input channels) -> bf16 operand tensors of pp_conv_igemm:
 * fwd [taps][Cout_pad][Cin_pad] and/or dgrad [taps][Cin_rows][Cout_cols] (taps flipped); zero padded; either NULL. */
int pp_pack_conv_weight(const float* w, int Cout, int Cin, int Cin_total, int taps, void* fwd, int Cout_pad,
                        int Cin_pad, void* dgrad, int Cin_rows, int Cout_cols, void* stream);

/* Joint geometric augmentation of a training batch on the device - datasets/base_dataset.py:48-127 (random scale: PIL BILINEAR
 * for the image, PIL NEAREST for the label map, torch nearest for the query / human-label masks; pad to the crop size with
 * mean_val / ignore_index / 0; random crop; horizontal flip) fused with TF.to_tensor + TF.normalize (base_dataset.py:183).
 * x uint8 [B][H][W][3]; y / q / lq uint8 [B][H][W] (each may be NULL); header int32 [B][20] = {h_rs, w_rs, start_h, start_w,
 * flip, ksize_x, ksize_y, then the offsets into `tables` of: PIL-nearest x / y indices, torch-nearest x / y indices, and per
 * axis the bilinear filter's first source index, tap count and 22-bit fixed-point taps [out][ksize]} - built by the host in
 * double precision as Pillow does (pixelpick_b200/augment.py).  Outputs: x_out f32 [B][3][crop_h][crop_w] normalised,
 * y_out / q_out (0 / 1) / lq_out uint8 [B][crop_h][crop_w].  The image equals PIL's two-pass resample bit for bit. */
int pp_augment_geometric(const uint8_t* x, const uint8_t* y, const uint8_t* q, const uint8_t* lq, int B, int H, int W,
                         const int32_t* header, const int32_t* tables, int crop_h, int crop_w, const float* mean3,
                         const float* std3, const int* mean_val3, int ignore_index, float* x_out, uint8_t* y_out,
                         uint8_t* q_out, uint8_t* lq_out, void* stream);

/* pp_augment_geometric with the image left as uint8 [B][crop_h][crop_w][3] (HWC, what the PIL image holds after
 * base_dataset.py:48-127) instead of the normalised tensor: the input of pp_augment_photometric. */
int pp_augment_geometric_u8(const uint8_t* x, const uint8_t* y, const uint8_t* q, const uint8_t* lq, int B, int H, int W,
                            const int32_t* header, const int32_t* tables, int crop_h, int crop_w, const int* mean_val3,
                            int ignore_index, uint8_t* x_u8_out, uint8_t* y_out, uint8_t* q_out, uint8_t* lq_out, void* stream);

/* Photometric augmentation of a uint8 batch on the device - datasets/base_dataset.py:129-141: RandomApply([ColorJitter(0.8, 0.8,
 * 0.8, 0.2)], p = 0.8), RandomGrayscale(0.2), GaussianBlur (cv2, p = 0.5; base_dataset.py:192-210) - followed by TF.to_tensor +
 * TF.normalize (base_dataset.py:183).  The arithmetic is Pillow's (ImageEnhance / Blend.c, Convert.c RGB<->HSV, the L conversion)
 * and OpenCV's bit-exact uint8 Gaussian, so the result equals the reference's bit for bit given the same draws.
 * x uint8 [B][H][W][3]; header int32 [B][16] = {jitter_on, the four steps in their drawn order (0 brightness, 1 contrast,
 * 2 saturation, 3 hue), brightness / contrast / saturation factors (fp32 bit patterns), hue shift in [0, 255] (= int32(hue * 255)
 * mod 256), grayscale_on, blur_on, 0...}; blur_taps int32 [B][ksize] = the 8-bit fixed-point Gaussian taps of each image (sum 256;
 * ignored where blur_on = 0; ksize = 0: no image is blurred).  The draws and the taps are made by the host
 * (pixelpick_b200/augment.py: draw_photometric, gaussian_taps_q8).  Outputs (either may be NULL): x_out f32 [B][3][H][W]
 * normalised with mean3 / std3, x_u8_out uint8 [B][H][W][3].  workspace: pp_augment_photometric_workspace_bytes(). */
int pp_augment_photometric_workspace_bytes(int B, int H, int W, size_t* bytes);
int pp_augment_photometric(const uint8_t* x, int B, int H, int W, const int32_t* header, const int32_t* blur_taps, int ksize,
                           const float* mean3, const float* std3, void* workspace, size_t workspace_bytes, float* x_out,
                           uint8_t* x_u8_out, void* stream);

/* The same for EVERY convolution of a network in one launch (a train step re-packs ~50 weights after each optimiser
 * update): table_dev = device array of n rows of 12 int64 {w, fwd, dgrad pointers, Cout, Cin, Cin_total, taps, Cout_pad,
 * Cin_pad, Cin_rows, Cout_cols, first_tile} with the meaning of pp_pack_conv_weight's arguments.  The work is cut into tiles
 * of 64 x 64 (1x1) / 32 x 32 x 9 (3x3) weights, one CTA each: pp_pack_conv_weights_tiles() gives a row's tile count (taps must
 * be 1 or 9, the inner extents Cin_pad / Cout_cols multiples of 8 and the images 16-byte aligned; pass 0, 0 for an image that
 * is not wanted), first_tile is the running
 * sum over the rows before it and total_tiles the sum over all rows. */
int pp_pack_conv_weights_tiles(int Cout_pad, int Cin_pad, int Cin_rows, int Cout_cols, int taps);
int pp_pack_conv_weights_batched(const long long* table_dev, int n, int total_tiles, void* stream);
/* strided [N,C,H,W] f32/bf16 -> bf16 NHWC channel slice (backbone boundary, d(logits) for the classifier) */
int pp_to_nhwc_bf16(const void* in, int dtype, int64_t sn, int64_t sc, int64_t sh, int64_t sw, int N, int C, int H,
                    int W, void* out, int ld, int c_off, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PIXELPICK_B200_H_ */
