// Greedy box NMS with the exact semantics of torchvision.ops.nms on CPU (called at
// object_reasoning.py:661 with all-equal scores and at object_scoring.py:238 with the
// boundary score): stable descending sort, then for each surviving box i suppress every
// later j with  inter / (area_i + area_j - inter) > thr, all in fp32 with the same
// operation order, so keep-sets are bit-exact.
//
// One CTA per image: rank sort (stable by construction: ties broken by index), then a
// matrix-free greedy sweep with the alive set as a shared-memory bitmask.  Per-image lists
// on this path are a few hundred boxes; the sweep is O(kept * n / 1024) with one barrier
// per kept box.
#include "unmore_internal.h"

namespace unmore {

constexpr int kNmsThreads = 1024;
constexpr int kNmsMaxBoxes = 32768;  // alive bitmask: 4 KB of shared memory

__global__ void __launch_bounds__(kNmsThreads) box_nms_kernel(const NmsParams p) {
  __shared__ uint32_t alive[kNmsMaxBoxes / 32];
  __shared__ int n_keep;
  const int b = blockIdx.x;
  const int n = min(p.counts ? p.counts[b] : p.cap, min(p.cap, kNmsMaxBoxes));
  const float4* boxes = p.boxes + (size_t)b * p.cap;
  int* order = p.order_ws + (size_t)b * p.cap;
  int* keep = p.keep + (size_t)b * p.cap;
  // ---- stable descending order
  if (p.scores) {
    const float* sc = p.scores + (size_t)b * p.cap;
    for (int i = threadIdx.x; i < n; i += kNmsThreads) {
      const float si = sc[i];
      int rank = 0;
      for (int j = 0; j < n; ++j) {
        const float sj = sc[j];
        rank += (sj > si || (sj == si && j < i)) ? 1 : 0;
      }
      order[rank] = i;
    }
  } else {
    for (int i = threadIdx.x; i < n; i += kNmsThreads) order[i] = i;
  }
  for (int w = threadIdx.x; w < (n + 31) / 32; w += kNmsThreads)
    alive[w] = (w * 32 + 32 <= n) ? 0xffffffffu : ((1u << (n - w * 32)) - 1u);
  if (threadIdx.x == 0) n_keep = 0;
  __syncthreads();
  // ---- greedy sweep
  for (int i = 0; i < n; ++i) {
    if (!((alive[i >> 5] >> (i & 31)) & 1u)) continue;  // uniform: everyone reads the same word
    const int oi = order[i];
    const float4 bi = boxes[oi];
    const float iarea = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
    if (threadIdx.x == 0) {
      keep[n_keep] = oi;
      if (p.boxes_out) p.boxes_out[(size_t)b * p.cap + n_keep] = bi;
      n_keep = n_keep + 1;
    }
    for (int j = i + 1 + threadIdx.x; j < n; j += kNmsThreads) {
      if (!((alive[j >> 5] >> (j & 31)) & 1u)) continue;
      const float4 bj = boxes[order[j]];
      const float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
      const float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
      const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
      const float inter = __fmul_rn(w, h);
      const float jarea = __fmul_rn(__fsub_rn(bj.z, bj.x), __fsub_rn(bj.w, bj.y));
      const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, jarea), inter));
      if (ovr > p.iou_thr) atomicAnd(&alive[j >> 5], ~(1u << (j & 31)));
    }
    __syncthreads();
  }
  __syncthreads();
  if (threadIdx.x == 0) p.keep_counts[b] = n_keep;
}

int launch_box_nms(const NmsParams& p, cudaStream_t stream) {
  if (p.n_img <= 0) return 0;
  box_nms_kernel<<<p.n_img, kNmsThreads, 0, stream>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace unmore
