// C-ABI entry points (include/unmore_b200.h).  Argument checking, workspace carving and
// kernel launches only; no device allocation, no synchronisation, no global mutable state
// other than cached device attributes and the thread-local error string.
#include <cstdio>
#include <cstdarg>
#include <cmath>
#include <cstring>

#include "../../include/unmore_b200.h"
#include "unmore_internal.h"

using namespace unmore;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(int code, const char* what) {
  if (code == 0) return 0;
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString((cudaError_t)code));
  return code;
}

int num_sms() {
  static int cached = 0;
  if (!cached) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

// ws layout: [0] work counter, [1 .. n_img+1] exclusive prefix of counts
int make_worklist(WorkList& w, const int* counts, int n_img, int cap, void* ws, cudaStream_t s) {
  int* wsi = reinterpret_cast<int*>(ws);
  w.counter = wsi;
  w.n_img = n_img;
  w.cap = cap;
  w.total_dense = n_img * cap;
  w.offsets = nullptr;
  int e = (int)cudaMemsetAsync(wsi, 0, sizeof(int), s);
  if (e) return cuda_fail(e, "memset work counter");
  if (counts) {
    w.offsets = wsi + 1;
    e = launch_prefix_counts(counts, n_img, wsi + 1, s);
    if (e) return cuda_fail(e, "prefix_counts");
  }
  return 0;
}

int check_fields(const float* f, int n_img, int C, int H, int W) {
  if (!f || n_img <= 0 || C <= 0 || H <= 0 || W <= 0) return fail(UNMORE_E_INVALID, "bad field tensor");
  return 0;
}

}  // namespace

extern "C" {

const char* unmore_last_error(void) { return g_err; }
int unmore_version(void) { return 100; }
size_t unmore_workspace_bytes(int n_img) { return sizeof(int) * (size_t)(n_img + 2 > 2 ? n_img + 2 : 2); }

int unmore_existence_scores(const float* fields, int n_img, int C, int H, int W, int ch_exist, const void* boxes,
                            int boxes_f64, const int* counts, int cap, float* scores_out, void* ws,
                            unmore_stream_t stream) {
  if (int e = check_fields(fields, n_img, C, H, W)) return e;
  if (!boxes || !scores_out || !ws || cap < 0 || ch_exist < 0 || ch_exist >= C)
    return fail(UNMORE_E_INVALID, "unmore_existence_scores: bad argument");
  if (cap == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  ExistParams p{};
  p.fields = fields; p.C = C; p.H = H; p.W = W; p.ch_exist = ch_exist;
  p.boxes = boxes; p.boxes_f64 = boxes_f64; p.scores = scores_out;
  if (int e = make_worklist(p.work, counts, n_img, cap, ws, s)) return e;
  return cuda_fail(launch_existence(p, num_sms(), s), "existence_kernel");
}

int unmore_center_reasoning(const float* fields, int n_img, int C, int H, int W, int ch_sdf, int ch_center_row,
                            int ch_center_col, const void* boxes, int boxes_f64, const int* counts, int cap,
                            double center_score_max_thres, double* max_values_out, int* argmax_out,
                            double* splits_out, void* ws, unmore_stream_t stream) {
  if (int e = check_fields(fields, n_img, C, H, W)) return e;
  if (!boxes || !max_values_out || !argmax_out || !ws || cap < 0 || ch_sdf < 0 || ch_sdf >= C || ch_center_row < 0 ||
      ch_center_row >= C || ch_center_col < 0 || ch_center_col >= C)
    return fail(UNMORE_E_INVALID, "unmore_center_reasoning: bad argument");
  if (cap == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  CenterParams p{};
  p.fields = fields; p.C = C; p.H = H; p.W = W;
  p.ch_sdf = ch_sdf; p.ch_crow = ch_center_row; p.ch_ccol = ch_center_col;
  p.boxes = boxes; p.boxes_f64 = boxes_f64; p.thr = center_score_max_thres;
  p.max_values = max_values_out; p.argmax = argmax_out; p.splits = splits_out;
  for (int i = 0; i < 5; ++i)
    for (int j = 0; j < 5; ++j) {
      const int di = 2 - i, dj = 2 - j;
      // F.normalize(dim=1): v / max(||v||_2, 1e-12), fp32; the centre tap is 0 / 1e-12 = 0
      const float n = std::sqrt((float)(di * di + dj * dj));
      p.filt[i * 5 + j] = (double)((float)di / (n > 1e-12f ? n : 1e-12f));
    }
  if (int e = make_worklist(p.work, counts, n_img, cap, ws, s)) return e;
  return cuda_fail(launch_center(p, num_sms(), s), "center_kernel");
}

int unmore_boundary_refine(const float* fields, int n_img, int C, int H, int W, int ch_sdf, const void* boxes,
                           int boxes_f64, const int* counts, int cap, int n_round, int apply_small_filter,
                           int early_exit, float proposal_area_thres, float max_sdf_thres, float max_shrink_threshold,
                           float delta_ratio, float* boxes_out, float* labels_out, int* rounds_out, void* ws,
                           unmore_stream_t stream) {
  if (int e = check_fields(fields, n_img, C, H, W)) return e;
  if (!boxes || !boxes_out || !labels_out || !ws || cap < 0 || n_round < 1 || ch_sdf < 0 || ch_sdf >= C)
    return fail(UNMORE_E_INVALID, "unmore_boundary_refine: bad argument");
  if (cap == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  RefineParams p{};
  p.fields = fields; p.C = C; p.H = H; p.W = W; p.ch_sdf = ch_sdf;
  p.boxes = boxes; p.boxes_f64 = boxes_f64;
  p.n_round = n_round; p.apply_small_filter = apply_small_filter; p.early_exit = early_exit;
  p.area_thres = proposal_area_thres; p.max_sdf_thres = max_sdf_thres;
  p.max_shrink_thres = max_shrink_threshold; p.delta_ratio = delta_ratio;
  p.boxes_out = reinterpret_cast<float4*>(boxes_out); p.labels_out = labels_out; p.rounds_out = rounds_out;
  if (int e = make_worklist(p.work, counts, n_img, cap, ws, s)) return e;
  return cuda_fail(launch_refine(p, num_sms(), s), "refine_kernel");
}

int unmore_update_bbox_from_tiles(const float* tiles, int M, float* deltas_out, float* max_out,
                                  unmore_stream_t stream) {
  if (M < 0 || (M > 0 && (!tiles || !deltas_out))) return fail(UNMORE_E_INVALID, "unmore_update_bbox_from_tiles: bad argument");
  TileParams p{tiles, M, reinterpret_cast<float4*>(deltas_out), max_out};
  return cuda_fail(launch_tiles(p, (cudaStream_t)stream), "tiles_kernel");
}

int unmore_compact_boxes(const void* in, int in_f64, const int* counts_in, int cap_in, int group, int mode,
                         const void* pred, float thr, void* out, int out_f64, int cap_out, int* counts_out,
                         int append, int* index_out, int n_img, unmore_stream_t stream) {
  if (!in || !pred || !out || !counts_out || cap_in < 0 || cap_out < 0 || group < 1 || mode < 0 || mode > 4 || n_img < 0)
    return fail(UNMORE_E_INVALID, "unmore_compact_boxes: bad argument");
  CompactParams p{};
  p.in = in; p.in_f64 = in_f64; p.counts_in = counts_in; p.cap_in = cap_in; p.group = group; p.mode = mode;
  p.pred = pred; p.thr = thr; p.out = out; p.out_f64 = out_f64; p.cap_out = cap_out; p.counts_out = counts_out;
  p.append = append; p.index_out = index_out; p.n_img = n_img;
  return cuda_fail(launch_compact(p, (cudaStream_t)stream), "compact_kernel");
}

int unmore_box_nms(const float* boxes, const float* scores, const int* counts, int cap, int n_img, float iou_threshold,
                   int* keep_out, int* keep_counts_out, float* boxes_out, int* order_ws, unmore_stream_t stream) {
  if (!boxes || !keep_out || !keep_counts_out || !order_ws || cap < 0 || n_img < 0)
    return fail(UNMORE_E_INVALID, "unmore_box_nms: bad argument");
  if (cap > 32768) return fail(UNMORE_E_CAPACITY, "unmore_box_nms: cap %d > 32768", cap);
  NmsParams p{};
  p.boxes = reinterpret_cast<const float4*>(boxes); p.scores = scores; p.counts = counts; p.cap = cap; p.n_img = n_img;
  p.iou_thr = iou_threshold; p.keep = keep_out; p.keep_counts = keep_counts_out;
  p.boxes_out = reinterpret_cast<float4*>(boxes_out); p.order_ws = order_ws; p.alive_ws = nullptr;
  return cuda_fail(launch_box_nms(p, (cudaStream_t)stream), "box_nms_kernel");
}

}  // extern "C"
