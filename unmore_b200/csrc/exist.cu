// existence_checking (object_reasoning.py:491-523) under the field-stub bridge: the
// Binary_Classifier stand-in is the mean of the resized crop's existence channel, so the
// score of a proposal is mean_{128x128} bilinear(existence[y1:y2, x1:x2]).
//
// One warp per proposal, lane owns 4 output columns; the horizontally interpolated source
// rows are cached across output rows (resample.cuh), so small crops cost ~in_h row fetches.
#include "resample.cuh"
#include "unmore_internal.h"

namespace unmore {

constexpr int kExistWarps = 8;

__global__ void __launch_bounds__(kExistWarps * 32) existence_kernel(const ExistParams p) {
  const int lane = threadIdx.x & 31;
  const int total = worklist_total(p.work);
  for (;;) {
    const int id = worklist_next_warp(p.work);
    if (id >= total) break;
    int img, k;
    worklist_locate(p.work, id, img, k);
    const size_t row = (size_t)img * p.work.cap + k;
    double x1, y1, x2, y2;
    load_box<double>(p.boxes, p.boxes_f64 != 0, row, x1, y1, x2, y2);
    const Window win = snap_window<double>(x1, y1, x2, y2, p.W, p.H);
    float score = 0.f;  // zero-size crop: the reference raises; defined as "nothing there"
    if (!win.empty()) {
      ColTaps taps;
      taps.init<kBlocked>(lane, win.w());
      PlaneRows plane;
      plane.init(p.fields + ((size_t)img * p.C + p.ch_exist) * p.H * p.W, p.W, win);
      const float scale_y = __fdiv_rn((float)win.h(), (float)kCrop);
      const int in_h = win.h();
      double acc = 0.0;
      float part = 0.f;
      for (int i = 0; i < kCrop; ++i) {
        float v[4];
        plane.row(taps, axis_tap(scale_y, i, in_h), v);
        part += (v[0] + v[1]) + (v[2] + v[3]);
        if ((i & 7) == 7) { acc += (double)part; part = 0.f; }
      }
      acc = warp_sum(acc);
      score = (float)(acc * (1.0 / (kCrop * kCrop)));
    }
    if (lane == 0) p.scores[row] = score;
  }
}

int launch_existence(const ExistParams& p, int num_sms, cudaStream_t stream) {
  existence_kernel<<<num_sms * 4, kExistWarps * 32, 0, stream>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace unmore
