// Q path through HOST buffers: pp_acq_session_* (include/pixelpick_b200.h).
// Two slots, one stream each: slot s owns device staging for `chunk_imgs` images; chunk i runs
// H2D -> score(+hist0) -> radix select -> pick -> D2H on stream (i & 1), so the copy engine fills one slot while the
// SMs work on the other.  Measured on the B200 pod: every host<->device copy pays ~0.9 ms before its first byte
// (16 MB: 12.5 GiB/s, 160 MB: 33 GiB/s, 640 MB: 48.7 GiB/s — scripts/bench_h2d.py), and the kernels need ~0.15 ms per
// 64 images, so the session wants FEW, LARGE logits copies and must not queue the three small copies (masks, pick
// positions) behind them: those go out on a side stream, concurrently with the logits copy.
#include "pp_common.cuh"
#include <new>

struct pp_acq_session {
  int chunk, C, H, W, k, n_sel;
  int pending_n, pending_strategy;  // pp_acq_session_begin_host .. finish_host
  size_t ws_bytes;
  struct Slot {
    cudaStream_t st;
    cudaStream_t aux;     // small H2D copies (masks, positions), concurrent with the logits copy
    cudaEvent_t aux_done;
    float* logits;
    uint8_t* lab;
    uint8_t* vd;
    float* score;
    void* ws;
    int32_t* topk;
    int32_t* pos;
    int32_t* sel;
  } slot[2];
};

using namespace pp;

extern "C" {

int pp_acq_session_destroy(pp_acq_session* s) {
  if (!s) return PP_OK;
  for (int i = 0; i < 2; ++i) {
    auto& sl = s->slot[i];
    if (sl.st) cudaStreamSynchronize(sl.st);
    if (sl.aux) cudaStreamSynchronize(sl.aux);
    cudaFree(sl.logits);
    cudaFree(sl.lab);
    cudaFree(sl.vd);
    cudaFree(sl.score);
    cudaFree(sl.ws);
    cudaFree(sl.topk);
    cudaFree(sl.pos);
    cudaFree(sl.sel);
    if (sl.st) cudaStreamDestroy(sl.st);
    if (sl.aux) cudaStreamDestroy(sl.aux);
    if (sl.aux_done) cudaEventDestroy(sl.aux_done);
  }
  delete s;
  return PP_OK;
}

int pp_acq_session_create(pp_acq_session** out, int chunk_imgs, int C, int H, int W, int k, int n_sel) {
  PP_CHECK_ARG(out, "pp_acq_session_create: null out");
  PP_CHECK_ARG(chunk_imgs > 0 && C >= 2 && H > 0 && W > 0, "pp_acq_session_create: bad shape");
  const int64_t HW = (int64_t)H * W;
  PP_CHECK_ARG(HW <= (1 << 22) && k > 0 && k <= HW && n_sel > 0 && n_sel <= k, "pp_acq_session_create: bad k/n_sel");
  pp_acq_session* s = new (std::nothrow) pp_acq_session();
  if (!s) {
    set_error("pp_acq_session_create: out of host memory");
    return PP_ERR_WORKSPACE;
  }
  memset(s, 0, sizeof(*s));
  s->chunk = chunk_imgs; s->C = C; s->H = H; s->W = W; s->k = k; s->n_sel = n_sel;
  int rc = pp_acq_topk_workspace_bytes(chunk_imgs, (int)HW, k, &s->ws_bytes);
  if (rc != PP_OK) { delete s; return rc; }
  for (int i = 0; i < 2; ++i) {
    auto& sl = s->slot[i];
    cudaError_t e = cudaStreamCreateWithFlags(&sl.st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&sl.aux, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.aux_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&sl.logits, (size_t)chunk_imgs * C * HW * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&sl.lab, (size_t)chunk_imgs * HW);
    if (e == cudaSuccess) e = cudaMalloc(&sl.vd, (size_t)chunk_imgs * HW);
    if (e == cudaSuccess) e = cudaMalloc(&sl.score, (size_t)chunk_imgs * HW * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&sl.ws, s->ws_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&sl.topk, (size_t)chunk_imgs * k * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&sl.pos, (size_t)chunk_imgs * n_sel * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&sl.sel, (size_t)chunk_imgs * n_sel * sizeof(int32_t));
    if (e != cudaSuccess) {
      set_error("pp_acq_session_create: %s", cudaGetErrorString(e));
      pp_acq_session_destroy(s);
      return PP_ERR_CUDA;
    }
  }
  *out = s;
  return PP_OK;
}

// Split form of run_host for hosts that draw the pick positions themselves (query.py:63-64: one
// np.random.permutation(k) per image, ~60 us each): begin launches the copies, the scoring and the radix select and
// returns at once; the host draws while the logits are still crossing PCIe; finish uploads the positions, picks and
// returns the pixels.  All chunks must be resident at once: n_img <= 2 * chunk_imgs.
int pp_acq_session_begin_host(pp_acq_session* s, const float* h_logits, const uint8_t* h_labelled, const uint8_t* h_void,
                              int n_img, int strategy) {
  PP_CHECK_ARG(s && h_logits && n_img > 0, "pp_acq_session_begin_host: bad args");
  PP_CHECK_ARG(strategy >= 0 && strategy <= 2, "pp_acq_session_begin_host: bad strategy %d", strategy);
  PP_CHECK_ARG(n_img <= 2 * s->chunk, "pp_acq_session_begin_host: %d images exceed the two resident chunks of %d", n_img, s->chunk);
  PP_CHECK_ARG(s->pending_n == 0, "pp_acq_session_begin_host: a previous begin has not been finished");
  const int64_t HW = (int64_t)s->H * s->W;
  const int largest = strategy == PP_STRAT_MARGIN ? 0 : 1;
  int ci = 0;
  for (int i0 = 0; i0 < n_img; i0 += s->chunk, ++ci) {
    auto& sl = s->slot[ci & 1];
    const int n = (n_img - i0) < s->chunk ? (n_img - i0) : s->chunk;
    PP_CUDA(cudaEventRecord(sl.aux_done, sl.st));
    PP_CUDA(cudaStreamWaitEvent(sl.aux, sl.aux_done, 0));
    PP_CUDA(cudaMemcpyAsync(sl.logits, h_logits + (size_t)i0 * s->C * HW, (size_t)n * s->C * HW * sizeof(float),
                            cudaMemcpyHostToDevice, sl.st));
    if (h_labelled)
      PP_CUDA(cudaMemcpyAsync(sl.lab, h_labelled + (size_t)i0 * HW, (size_t)n * HW, cudaMemcpyHostToDevice, sl.aux));
    if (h_void)
      PP_CUDA(cudaMemcpyAsync(sl.vd, h_void + (size_t)i0 * HW, (size_t)n * HW, cudaMemcpyHostToDevice, sl.aux));
    PP_CUDA(cudaEventRecord(sl.aux_done, sl.aux));
    PP_CUDA(cudaStreamWaitEvent(sl.st, sl.aux_done, 0));
    int rc = pp_acq_topk_prepare(sl.ws, s->ws_bytes, n, (int)HW, s->k, sl.st);
    if (rc != PP_OK) return rc;
    rc = pp_acq_score(sl.logits, PP_F32, n, s->C, s->H, s->W, (int64_t)s->C * HW, HW, s->W,
                      h_labelled ? sl.lab : nullptr, h_void ? sl.vd : nullptr, nullptr, strategy, sl.score,
                      pp_acq_topk_hist0(sl.ws), sl.st);
    if (rc != PP_OK) return rc;
    rc = pp_acq_select(sl.score, n, (int)HW, s->k, largest, 1, sl.ws, s->ws_bytes, sl.st);
    if (rc != PP_OK) return rc;
  }
  s->pending_n = n_img;
  s->pending_strategy = strategy;
  return PP_OK;
}

int pp_acq_session_finish_host(pp_acq_session* s, const int32_t* h_pos, int32_t* h_sel_idx) {
  PP_CHECK_ARG(s && h_sel_idx, "pp_acq_session_finish_host: bad args");
  PP_CHECK_ARG(s->pending_n > 0, "pp_acq_session_finish_host: nothing pending (call begin first)");
  const int64_t HW = (int64_t)s->H * s->W;
  const int n_img = s->pending_n;
  s->pending_n = 0;
  int ci = 0;
  for (int i0 = 0; i0 < n_img; i0 += s->chunk, ++ci) {
    auto& sl = s->slot[ci & 1];
    const int n = (n_img - i0) < s->chunk ? (n_img - i0) : s->chunk;
    if (h_pos)
      PP_CUDA(cudaMemcpyAsync(sl.pos, h_pos + (size_t)i0 * s->n_sel, (size_t)n * s->n_sel * sizeof(int32_t),
                              cudaMemcpyHostToDevice, sl.st));
    int rc = pp_acq_pick(sl.ws, s->ws_bytes, n, (int)HW, s->k, h_pos ? sl.pos : nullptr, s->n_sel, sl.sel, sl.st);
    if (rc != PP_OK) return rc;
    PP_CUDA(cudaMemcpyAsync(h_sel_idx + (size_t)i0 * s->n_sel, sl.sel, (size_t)n * s->n_sel * sizeof(int32_t),
                            cudaMemcpyDeviceToHost, sl.st));
  }
  PP_CUDA(cudaStreamSynchronize(s->slot[0].st));
  PP_CUDA(cudaStreamSynchronize(s->slot[1].st));
  return PP_OK;
}

int pp_acq_session_run_host(pp_acq_session* s, const float* h_logits, const uint8_t* h_labelled,
                            const uint8_t* h_void, int n_img, int strategy, const int32_t* h_pos,
                            int32_t* h_sel_idx, int32_t* h_topk_idx) {
  PP_CHECK_ARG(s && h_logits && h_sel_idx && n_img > 0, "pp_acq_session_run_host: bad args");
  PP_CHECK_ARG(strategy >= 0 && strategy <= 2, "pp_acq_session_run_host: bad strategy %d", strategy);
  PP_CHECK_ARG(s->pending_n == 0, "pp_acq_session_run_host: a begin_host is pending on this session (finish it first)");
  const int64_t HW = (int64_t)s->H * s->W;
  const int largest = strategy == PP_STRAT_MARGIN ? 0 : 1;
  int ci = 0;
  for (int i0 = 0; i0 < n_img; i0 += s->chunk, ++ci) {
    auto& sl = s->slot[ci & 1];
    const int n = (n_img - i0) < s->chunk ? (n_img - i0) : s->chunk;
    // the slot's previous kernels (2 chunks ago, same stream) must be done with lab / vd / pos before aux overwrites them
    PP_CUDA(cudaEventRecord(sl.aux_done, sl.st));
    PP_CUDA(cudaStreamWaitEvent(sl.aux, sl.aux_done, 0));
    PP_CUDA(cudaMemcpyAsync(sl.logits, h_logits + (size_t)i0 * s->C * HW, (size_t)n * s->C * HW * sizeof(float),
                            cudaMemcpyHostToDevice, sl.st));
    if (h_labelled)
      PP_CUDA(cudaMemcpyAsync(sl.lab, h_labelled + (size_t)i0 * HW, (size_t)n * HW, cudaMemcpyHostToDevice, sl.aux));
    if (h_void)
      PP_CUDA(cudaMemcpyAsync(sl.vd, h_void + (size_t)i0 * HW, (size_t)n * HW, cudaMemcpyHostToDevice, sl.aux));
    if (h_pos)
      PP_CUDA(cudaMemcpyAsync(sl.pos, h_pos + (size_t)i0 * s->n_sel, (size_t)n * s->n_sel * sizeof(int32_t),
                              cudaMemcpyHostToDevice, sl.aux));
    PP_CUDA(cudaEventRecord(sl.aux_done, sl.aux));
    PP_CUDA(cudaStreamWaitEvent(sl.st, sl.aux_done, 0));
    int rc = pp_acq_topk_prepare(sl.ws, s->ws_bytes, n, (int)HW, s->k, sl.st);
    if (rc != PP_OK) return rc;
    rc = pp_acq_score(sl.logits, PP_F32, n, s->C, s->H, s->W, (int64_t)s->C * HW, HW, s->W,
                      h_labelled ? sl.lab : nullptr, h_void ? sl.vd : nullptr, nullptr, strategy, sl.score,
                      pp_acq_topk_hist0(sl.ws), sl.st);
    if (rc != PP_OK) return rc;
    if (h_topk_idx) {  // the caller wants the whole sorted list: sort, then gather the picks from it
      rc = pp_acq_topk(sl.score, n, (int)HW, s->k, largest, 1, sl.topk, nullptr, sl.ws, s->ws_bytes, sl.st);
      if (rc != PP_OK) return rc;
      rc = pp_acq_gather(sl.topk, n, s->k, h_pos ? sl.pos : nullptr, s->n_sel, sl.sel, sl.st);
    } else {  // picks only: radix select + order statistics, no sort
      rc = pp_acq_select(sl.score, n, (int)HW, s->k, largest, 1, sl.ws, s->ws_bytes, sl.st);
      if (rc != PP_OK) return rc;
      rc = pp_acq_pick(sl.ws, s->ws_bytes, n, (int)HW, s->k, h_pos ? sl.pos : nullptr, s->n_sel, sl.sel, sl.st);
    }
    if (rc != PP_OK) return rc;
    PP_CUDA(cudaMemcpyAsync(h_sel_idx + (size_t)i0 * s->n_sel, sl.sel, (size_t)n * s->n_sel * sizeof(int32_t),
                            cudaMemcpyDeviceToHost, sl.st));
    if (h_topk_idx)
      PP_CUDA(cudaMemcpyAsync(h_topk_idx + (size_t)i0 * s->k, sl.topk, (size_t)n * s->k * sizeof(int32_t),
                              cudaMemcpyDeviceToHost, sl.st));
  }
  PP_CUDA(cudaStreamSynchronize(s->slot[0].st));
  PP_CUDA(cudaStreamSynchronize(s->slot[1].st));
  return PP_OK;
}

}  // extern "C"
