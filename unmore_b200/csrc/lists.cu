// Ragged-list plumbing: exclusive prefix of per-image counts (work scheduling) and the
// order-preserving compactions that replace the reference's boolean-mask indexing
// (e.g. object_reasoning.py:422-426, 541-542, 630, 656).  Order preservation matters: the
// final discovery NMS has all-equal scores, so index order decides the keep-set (:661).
#include "unmore_internal.h"

namespace unmore {

__global__ void prefix_counts_kernel(const int* __restrict__ counts, int n_img, int* __restrict__ offsets) {
  // single CTA; n_img is small (images per launch)
  __shared__ int carry;
  __shared__ int wsum[32];
  if (threadIdx.x == 0) { carry = 0; offsets[0] = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n_img; base += blockDim.x) {
    const int i = base + threadIdx.x;
    int v = i < n_img ? counts[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(kFullMask, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int s = lane < (blockDim.x >> 5) ? wsum[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(kFullMask, s, o);
        if (lane >= o) s += y;
      }
      wsum[lane] = s;  // inclusive over warps
    }
    __syncthreads();
    const int incl = x + (warp ? wsum[warp - 1] : 0) + carry;
    if (i < n_img) offsets[i + 1] = incl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = incl;
    __syncthreads();
  }
}

int launch_prefix_counts(const int* counts, int n_img, int* offsets, cudaStream_t stream) {
  prefix_counts_kernel<<<1, 1024, 0, stream>>>(counts, n_img, offsets);
  return (int)cudaGetLastError();
}

__device__ __forceinline__ bool compact_pred(const CompactParams& p, size_t e) {
  switch (p.mode) {
    case kFlagsU8: return reinterpret_cast<const unsigned char*>(p.pred)[e] != 0;
    case kScoreGE: return reinterpret_cast<const float*>(p.pred)[e] >= p.thr;
    case kLabelEQ: return reinterpret_cast<const float*>(p.pred)[e] == p.thr;
    case kArgmaxGE0: return reinterpret_cast<const int*>(p.pred)[e] >= 0;
    case kArgmaxLT0: return reinterpret_cast<const int*>(p.pred)[e] < 0;
    default: return reinterpret_cast<const unsigned char*>(p.pred)[e] != 0;
  }
}

// One CTA per image: block-wide stable stream compaction; a selected entry carries `group` boxes,
// or group_counts[e] <= group of them when the per-entry counts are given (component boxes).
__global__ void __launch_bounds__(1024) compact_kernel(const CompactParams p) {
  __shared__ int wsum[32];
  __shared__ int carry;
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_in = p.counts_in ? p.counts_in[b] : p.cap_in;
  const int base_out = p.append ? p.counts_out[b] : 0;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int start = 0; start < n_in; start += blockDim.x) {
    const int e = start + threadIdx.x;
    const size_t ge = (size_t)b * p.cap_in + e;
    const bool sel = e < n_in && compact_pred(p, ge);
    const int mine = sel ? (p.group_counts ? (int)p.group_counts[ge] : p.group) : 0;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(kFullMask, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int s = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(kFullMask, s, o);
        if (lane >= o) s += y;
      }
      wsum[lane] = s;
    }
    __syncthreads();
    const int first = base_out + carry + (warp ? wsum[warp - 1] : 0) + incl - mine;
    for (int g = 0; g < mine; ++g) {
      const int orow = first + g;
      if (orow >= p.cap_out) break;
      const size_t src = ge * p.group + g, dst = (size_t)b * p.cap_out + orow;
      double v[4];
      if (p.in_f64) {
        const double4 t = reinterpret_cast<const double4*>(p.in)[src];
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
        const float4 t = reinterpret_cast<const float4*>(p.in)[src];
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      }
      if (p.out_f64) reinterpret_cast<double4*>(p.out)[dst] = make_double4(v[0], v[1], v[2], v[3]);
      else reinterpret_cast<float4*>(p.out)[dst] = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
      if (p.index_out) p.index_out[dst] = e;
    }
    __syncthreads();
    if (threadIdx.x == 0) carry += wsum[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int n = base_out + carry;
    if (n > p.cap_out && p.overflow) atomicAdd(p.overflow, 1);
    p.counts_out[b] = n < p.cap_out ? n : p.cap_out;
  }
}

int launch_compact(const CompactParams& p, cudaStream_t stream) {
  if (p.n_img <= 0) return 0;
  compact_kernel<<<p.n_img, 1024, 0, stream>>>(p);
  return (int)cudaGetLastError();
}

// ---- detection rows for the end-of-run collective ----------------------------------------------------
// Appends the detections of a batch — (image_id, x, y, w, h, score) as six doubles per row, image-major, in
// the order of the scoring NMS — to a fixed-capacity row buffer that is handed to the all-gather as it is:
// rows[0] = (row count, overflow flag, 0, 0, 0, 0) is the header, rows[1 ..] the detections.  The cursor
// lives in the header, so consecutive batches on one stream append without a host round trip, and the
// image ids travel as doubles (exact to 2^53; the float32 column of the round-1 gather collided above 2^24).
// One CTA: a batch holds a few thousand detections.
__global__ void __launch_bounds__(1024) pack_detections_kernel(const long long* __restrict__ image_ids,
                                                              const float4* __restrict__ bbox_xywh,
                                                              const double* __restrict__ out5,
                                                              const int* __restrict__ keep_counts, int cap, int n_img,
                                                              double* __restrict__ rows, int max_rows) {
  __shared__ int carry;
  __shared__ int wsum[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = (int)rows[0];
  __syncthreads();
  for (int base = 0; base < n_img; base += blockDim.x) {
    const int b = base + threadIdx.x;
    const int n = b < n_img ? min(keep_counts[b], cap) : 0;
    int x = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(kFullMask, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int s = lane < (blockDim.x >> 5) ? wsum[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(kFullMask, s, o);
        if (lane >= o) s += y;
      }
      wsum[lane] = s;
    }
    __syncthreads();
    const int first = carry + (warp ? wsum[warp - 1] : 0) + x - n;   // rows before this image
    for (int i = 0; i < n; ++i) {
      const int r = first + i;
      if (r >= max_rows) break;
      const size_t src = (size_t)b * cap + i;
      const float4 bx = bbox_xywh[src];
      double* o = rows + (size_t)(r + 1) * 6;
      o[0] = (double)image_ids[b]; o[1] = (double)bx.x; o[2] = (double)bx.y; o[3] = (double)bx.z; o[4] = (double)bx.w;
      o[5] = out5[src * 5];
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = first + n;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (carry > max_rows) { rows[1] = 1.0; carry = max_rows; }
    rows[0] = (double)carry;
  }
}

int launch_pack_detections(const long long* image_ids, const float4* bbox_xywh, const double* out5, const int* keep_counts,
                           int cap, int n_img, double* rows, int max_rows, cudaStream_t stream) {
  if (n_img <= 0) return 0;
  pack_detections_kernel<<<1, 1024, 0, stream>>>(image_ids, bbox_xywh, out5, keep_counts, cap, n_img, rows, max_rows);
  return (int)cudaGetLastError();
}

}  // namespace unmore
