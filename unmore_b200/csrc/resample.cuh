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

// Column taps of this lane for the current window, shared by all channels.  Weights are kept as packed
// pairs (columns c = 0,1 and c = 2,3 of the lane) so the horizontal stage is two FMUL2 + two FFMA2.
struct ColTaps {
  unsigned x0[4], x1[4];   // BYTE offsets of the two taps inside a source row; x1 = x0 + 4, or x0 again on the clamped right edge
  f32x2 w0[2], w1[2];
  template <int LAYOUT>
  __device__ __forceinline__ void init(int lane, int in_w) {
    const float scale = __fdiv_rn((float)in_w, (float)kCrop);
    float l0[4], l1[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      AxisTap t = axis_tap(scale, lane_column<LAYOUT>(lane, c), in_w);
      x0[c] = 4u * (unsigned)t.i0; x1[c] = 4u * (unsigned)t.i1; l0[c] = t.l0; l1[c] = t.l1;
    }
    w0[0] = pk2(l0[0], l0[1]); w0[1] = pk2(l0[2], l0[3]);
    w1[0] = pk2(l1[0], l1[1]); w1[1] = pk2(l1[2], l1[3]);
  }
};

// One channel plane of one window, with a two-row cache of horizontally interpolated rows
// (held as packed pairs: element h = columns 2h, 2h+1 of the lane).
struct PlaneRows {
  const float* origin;  // &plane[y1 * W + x1]
  int stride;           // W
  int cy0, cy1;         // source rows held in ra / rb (-1: none)
  f32x2 ra[2], rb[2];

  __device__ __forceinline__ void init(const float* plane, int W, const Window& win) {
    origin = plane + (size_t)win.y1 * W + win.x1;
    stride = W;
    cy0 = cy1 = -1;
  }
  __device__ __forceinline__ void hrow(const ColTaps& t, int y, f32x2 out[2]) const {
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
    for (int h = 0; h < 2; ++h)
      out[h] = lerp_h2(pk2(v0[2 * h], v0[2 * h + 1]), pk2(v1[2 * h], v1[2 * h + 1]), t.w0[h], t.w1[h]);
  }
  // S[i][lane_column(lane, 2h)], S[i][lane_column(lane, 2h+1)] packed in out[h], for the output row whose vertical tap is `v`
  __device__ __forceinline__ void row2(const ColTaps& t, const AxisTap& v, f32x2 out[2]) {
    if (v.i0 != cy0 || v.i1 != cy1) {          // warp-uniform
      if (v.i0 == cy1) {
        ra[0] = rb[0]; ra[1] = rb[1];
      } else {
        hrow(t, v.i0, ra);
      }
      if (v.i1 == v.i0) {
        rb[0] = ra[0]; rb[1] = ra[1];
      } else {
        hrow(t, v.i1, rb);
      }
      cy0 = v.i0; cy1 = v.i1;
    }
    out[0] = lerp_v2(ra[0], rb[0], v.l0, v.l1);
    out[1] = lerp_v2(ra[1], rb[1], v.l0, v.l1);
  }
  __device__ __forceinline__ void row(const ColTaps& t, const AxisTap& v, float out[4]) {
    f32x2 o[2];
    row2(t, v, o);
    upk2(o[0], out[0], out[1]);
    upk2(o[1], out[2], out[3]);
  }

  // ---- split-phase form: issue() starts the tap loads an output row still needs (none, one or two
  // source rows; both rows' sixteen requests go out together), finish() turns them into the row.  The
  // caller puts a whole row of arithmetic between the two, so the L1/L2 round trip of row i+1 is hidden
  // behind the work on row i.  The (cy0, cy1) bookkeeping depends only on tap indices, never on data, so
  // it can run ahead at issue time.
  struct Fetch {
    bool load_a, load_b;   // row i0 / row i1 must be fetched (taps in a* / b*)
    bool shift_a, dup_b;   // ra <- rb (the old lower row becomes the upper one); rb <- ra (clamped last row)
    float l0, l1;
    float a0[4], a1[4], b0[4], b1[4];
  };
  __device__ __forceinline__ void load_taps(const ColTaps& t, int y, float v0[4], float v1[4]) const {
    const float* rowp = elem_ptr(origin, y * stride);   // warp-uniform
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      v0[c] = __ldg(byte_ptr(rowp, t.x0[c]));
      v1[c] = __ldg(byte_ptr(rowp, t.x1[c]));
    }
  }
  // After the row: ra = H(i0), rb = H(i1).  H(i0) is the cached upper row, the cached lower row, or new;
  // H(i1) is H(i0) again (clamped), the cached lower row, or new.  Every condition is a function of tap
  // indices that are identical across the warp; routing them through a vote tells the compiler so and
  // keeps the branches free of reconvergence bookkeeping.
  // ROW_ELEMS > 0: the row pitch W is known at compile time; when both source rows of an output row are new
  // (always the case when the crop is more than twice the tile height) they are adjacent, so the lower row's
  // taps reuse the upper row's addresses through the load's immediate offset.
  template <int ROW_ELEMS = 0>
  __device__ __forceinline__ void issue(const ColTaps& t, const AxisTap& v, Fetch& f) {
    f.l0 = v.l0; f.l1 = v.l1;
    const bool keep_a = v.i0 == cy0;
    f.shift_a = warp_uniform(!keep_a && v.i0 == cy1);
    f.load_a = warp_uniform(!keep_a && v.i0 != cy1);
    f.dup_b = warp_uniform(v.i1 == v.i0);
    f.load_b = warp_uniform(v.i1 != v.i0 && v.i1 != cy1);
    if constexpr (ROW_ELEMS > 0) {
      if (f.load_a && f.load_b) {        // i1 == i0 + 1
        const float* rowp = elem_ptr(origin, v.i0 * ROW_ELEMS);   // warp-uniform
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float* p0 = byte_ptr(rowp, t.x0[c]);
          const float* p1 = byte_ptr(rowp, t.x1[c]);
          f.a0[c] = __ldg(p0); f.a1[c] = __ldg(p1);
          f.b0[c] = __ldg(p0 + ROW_ELEMS); f.b1[c] = __ldg(p1 + ROW_ELEMS);
        }
      } else {
        if (f.load_a) load_taps(t, v.i0, f.a0, f.a1);
        if (f.load_b) load_taps(t, v.i1, f.b0, f.b1);
      }
    } else {
      if (f.load_a) load_taps(t, v.i0, f.a0, f.a1);
      if (f.load_b) load_taps(t, v.i1, f.b0, f.b1);
    }
    cy0 = v.i0; cy1 = v.i1;
  }
  __device__ __forceinline__ void finish(const ColTaps& t, const Fetch& f, f32x2 out[2]) {
    if (f.load_a) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
        ra[h] = lerp_h2(pk2(f.a0[2 * h], f.a0[2 * h + 1]), pk2(f.a1[2 * h], f.a1[2 * h + 1]), t.w0[h], t.w1[h]);
    } else if (f.shift_a) {
      ra[0] = rb[0]; ra[1] = rb[1];
    }
    if (f.load_b) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
        rb[h] = lerp_h2(pk2(f.b0[2 * h], f.b0[2 * h + 1]), pk2(f.b1[2 * h], f.b1[2 * h + 1]), t.w0[h], t.w1[h]);
    } else if (f.dup_b) {
      rb[0] = ra[0]; rb[1] = ra[1];
    }
    out[0] = lerp_v2(ra[0], rb[0], f.l0, f.l1);
    out[1] = lerp_v2(ra[1], rb[1], f.l0, f.l1);
  }
};

// P channel planes of one window behind ONE row cache: the (cy0, cy1) bookkeeping and its branches
// run once per output row instead of once per plane, and the taps of all planes of a source row
// are requested together before any is consumed (P x 8 loads in flight).
// PLANE_ELEMS > 0: the P planes are consecutive channels of one image whose size is known at compile time;
// the tap addresses are then formed once (plane 0) and the other planes are reached through the immediate
// offset of the load instruction (P x 16 address instructions per source row become 16).
template <int P, int PLANE_ELEMS = 0>
struct MultiPlaneRows {
  const float* origin[P];
  int stride;
  int cy0, cy1;
  f32x2 ra[P][2], rb[P][2];

  __device__ __forceinline__ void init(const float* const planes[P], int W, const Window& win) {
#pragma unroll
    for (int p = 0; p < P; ++p) origin[p] = planes[p] + (size_t)win.y1 * W + win.x1;
    stride = W;
    cy0 = cy1 = -1;
  }
  __device__ __forceinline__ void hrows(const ColTaps& t, int y, f32x2 out[P][2]) const {
    const int ro = y * stride;
    float v0[P][4], v1[P][4];
    if constexpr (PLANE_ELEMS > 0) {
      const float* rowp = elem_ptr(origin[0], ro);   // warp-uniform
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float* a0 = byte_ptr(rowp, t.x0[c]);
        const float* a1 = byte_ptr(rowp, t.x1[c]);
#pragma unroll
        for (int p = 0; p < P; ++p) {
          v0[p][c] = __ldg(a0 + p * PLANE_ELEMS);
          v1[p][c] = __ldg(a1 + p * PLANE_ELEMS);
        }
      }
    } else {
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const float* rowp = elem_ptr(origin[p], ro);   // warp-uniform
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          v0[p][c] = __ldg(byte_ptr(rowp, t.x0[c]));
          v1[p][c] = __ldg(byte_ptr(rowp, t.x1[c]));
        }
      }
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
      for (int h = 0; h < 2; ++h)
        out[p][h] = lerp_h2(pk2(v0[p][2 * h], v0[p][2 * h + 1]), pk2(v1[p][2 * h], v1[p][2 * h + 1]), t.w0[h], t.w1[h]);
  }
  __device__ __forceinline__ void row(const ColTaps& t, const AxisTap& v, float out[P][4]) {
    if (v.i0 != cy0 || v.i1 != cy1) {          // warp-uniform
      if (v.i0 == cy1) {
#pragma unroll
        for (int p = 0; p < P; ++p) { ra[p][0] = rb[p][0]; ra[p][1] = rb[p][1]; }
      } else {
        hrows(t, v.i0, ra);
      }
      if (v.i1 == v.i0) {
#pragma unroll
        for (int p = 0; p < P; ++p) { rb[p][0] = ra[p][0]; rb[p][1] = ra[p][1]; }
      } else {
        hrows(t, v.i1, rb);
      }
      cy0 = v.i0; cy1 = v.i1;
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const f32x2 o = lerp_v2(ra[p][h], rb[p][h], v.l0, v.l1);
        upk2(o, out[p][2 * h], out[p][2 * h + 1]);
      }
  }
};

}  // namespace unmore
