// Warp-cooperative crop + bilinear resample to 128x128 ("a2", object_reasoning.py:402-410).
//
// Mapping: one warp per proposal; each lane owns four output columns of every output row.
// Two layouts, chosen per kernel from measurement (profiles/r01_column_mapping.md):
//   kBlocked: columns 4l .. 4l+3.  A lane's eight taps fall into 2-3 sectors, and the first tap
//             request of a row already touches every new sector of that row, so all L2 misses
//             of a row are in flight at once.  26 sectors / request but 87% L1 hits; the
//             fastest layout for the existence and refine kernels (-20% / -9% time vs strided).
//   kStrided: columns l, l+32, l+64, l+96.  4.6 sectors / request (3x less L1 work) and a
//             __ballot_sync yields one packed mask word per column group, which is what the
//             mask-building kernels (center reasoning, scoring) want.  Horizontal interpolation of a source row is kept in
// registers and reused while consecutive output rows hit the same source rows (always
// the case when the crop is smaller than 128 px high), so an up-sampled crop costs
// in_h * 8 loads per lane instead of 128 * 16.
#pragma once
#include "common.cuh"

namespace unmore {

enum ColLayout { kBlocked = 0, kStrided = 1 };

// base + idx (elements) as ONE IMAD.WIDE on a 64-bit base held in registers; without this the
// compiler re-derives the window origin per tap (IADD3 + LEA.HI.X.SX32 + LEA + LEA.HI.X).
__device__ __forceinline__ const float* elem_ptr(const float* base, int idx) {
  const float* r;
  asm("mad.wide.s32 %0, %1, 4, %2;" : "=l"(r) : "r"(idx), "l"(base));
  return r;
}
// base + byte offset: with the per-row base warp-uniform and the per-lane tap offsets precomputed in
// bytes, a tap address is this single instruction
__device__ __forceinline__ const float* byte_ptr(const float* base, unsigned off_bytes) {
  const float* r;
  asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(r) : "r"(off_bytes), "l"(base));
  return r;
}

template <int LAYOUT>
__device__ __forceinline__ int lane_column(int lane, int c) {
  return LAYOUT == kStrided ? lane + 32 * c : 4 * lane + c;
}

// Column taps of this lane for the current window, shared by all channels.
struct ColTaps {
  unsigned x0[4], x1[4];   // BYTE offsets of the two taps inside a source row; x1 = x0 + 4, or x0 again on the clamped right edge
  float w0[4], w1[4];
  template <int LAYOUT>
  __device__ __forceinline__ void init(int lane, int in_w) {
    const float scale = __fdiv_rn((float)in_w, (float)kCrop);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      AxisTap t = axis_tap(scale, lane_column<LAYOUT>(lane, c), in_w);
      x0[c] = 4u * (unsigned)t.i0; x1[c] = 4u * (unsigned)t.i1; w0[c] = t.l0; w1[c] = t.l1;
    }
  }
};

// One channel plane of one window, with a two-row cache of horizontally interpolated rows.
struct PlaneRows {
  const float* origin;  // &plane[y1 * W + x1]
  int stride;           // W
  int cy0, cy1;         // source rows held in ra / rb (-1: none)
  float ra[4], rb[4];

  __device__ __forceinline__ void init(const float* plane, int W, const Window& win) {
    origin = plane + (size_t)win.y1 * W + win.x1;
    stride = W;
    cy0 = cy1 = -1;
  }
  __device__ __forceinline__ void hrow(const ColTaps& t, int y, float out[4]) const {
    // 32-bit element offsets from the window origin, one address per tap.  Both taps of all four
    // columns are independent loads (no "v1 = v0 unless ..." dependency), so the eight requests of
    // a source row go out back to back and cost one memory round trip.
    const float* rowp = elem_ptr(origin, y * stride);   // warp-uniform
    float v0[4], v1[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      v0[c] = __ldg(byte_ptr(rowp, t.x0[c]));
      v1[c] = __ldg(byte_ptr(rowp, t.x1[c]));
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) out[c] = lerp_h(v0[c], v1[c], t.w0[c], t.w1[c]);
  }
  // S[i][lane_column(lane, c)], c = 0..3, for the output row whose vertical tap is `v`
  __device__ __forceinline__ void row(const ColTaps& t, const AxisTap& v, float out[4]) {
    if (v.i0 != cy0 || v.i1 != cy1) {          // warp-uniform
      if (v.i0 == cy1) {
#pragma unroll
        for (int c = 0; c < 4; ++c) ra[c] = rb[c];
      } else {
        hrow(t, v.i0, ra);
      }
      if (v.i1 == v.i0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) rb[c] = ra[c];
      } else {
        hrow(t, v.i1, rb);
      }
      cy0 = v.i0; cy1 = v.i1;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) out[c] = lerp_v(ra[c], rb[c], v.l0, v.l1);
  }
};

// P channel planes of one window behind ONE row cache: the (cy0, cy1) bookkeeping and its branches
// run once per output row instead of once per plane, and the taps of all planes of a source row
// are requested together before any is consumed (P x 8 loads in flight).
template <int P>
struct MultiPlaneRows {
  const float* origin[P];
  int stride;
  int cy0, cy1;
  float ra[P][4], rb[P][4];

  __device__ __forceinline__ void init(const float* const planes[P], int W, const Window& win) {
#pragma unroll
    for (int p = 0; p < P; ++p) origin[p] = planes[p] + (size_t)win.y1 * W + win.x1;
    stride = W;
    cy0 = cy1 = -1;
  }
  __device__ __forceinline__ void hrows(const ColTaps& t, int y, float out[P][4]) const {
    const int ro = y * stride;
    float v0[P][4], v1[P][4];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const float* rowp = elem_ptr(origin[p], ro);   // warp-uniform
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        v0[p][c] = __ldg(byte_ptr(rowp, t.x0[c]));
        v1[p][c] = __ldg(byte_ptr(rowp, t.x1[c]));
      }
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
      for (int c = 0; c < 4; ++c) out[p][c] = lerp_h(v0[p][c], v1[p][c], t.w0[c], t.w1[c]);
  }
  __device__ __forceinline__ void row(const ColTaps& t, const AxisTap& v, float out[P][4]) {
    if (v.i0 != cy0 || v.i1 != cy1) {          // warp-uniform
      if (v.i0 == cy1) {
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
          for (int c = 0; c < 4; ++c) ra[p][c] = rb[p][c];
      } else {
        hrows(t, v.i0, ra);
      }
      if (v.i1 == v.i0) {
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
          for (int c = 0; c < 4; ++c) rb[p][c] = ra[p][c];
      } else {
        hrows(t, v.i1, rb);
      }
      cy0 = v.i0; cy1 = v.i1;
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
      for (int c = 0; c < 4; ++c) out[p][c] = lerp_v(ra[p][c], rb[p][c], v.l0, v.l1);
  }
};

}  // namespace unmore
