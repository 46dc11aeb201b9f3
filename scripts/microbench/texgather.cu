// Micro-benchmark: bilinear tap gather of a WxH window -> 128x128 tile, one warp per window, lane = columns l + 32c.
//   ldg : what the kernels do today (two source rows kept in registers, reloaded when the vertical tap moves)
//   tld4: one texture-gather instruction per (column, row) returns the 2x2 footprint from a CUDA array
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o texgather texgather.cu ; run: ./texgather
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int PW = 640, PH = 480, GX = 8, GY = 8, NP = GX * GY;

struct Win { int p, x, y, w, h; };

__device__ __forceinline__ void tap(float scale, int i, int n, int& i0, int& i1, float& l0, float& l1) {
  const float src = fmaxf(fmaf(scale, (float)i + 0.5f, -0.5f), 0.f);
  i0 = min((int)src, n - 1); i1 = min(i0 + 1, n - 1); l1 = src - (float)i0; l0 = 1.f - l1;
}

__global__ void __launch_bounds__(256) k_ldg(const float* __restrict__ src, const Win* __restrict__ wins, int n, float* out) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int wi = blockIdx.x * wpb + (threadIdx.x >> 5); wi < n; wi += gridDim.x * wpb) {
    const Win w = wins[wi];
    const float* base = src + ((size_t)w.p * PH + w.y) * PW + w.x;
    const float sx = (float)w.w / 128.f, sy = (float)w.h / 128.f;
    int x0[4], x1[4]; float a0[4], a1[4];
    for (int c = 0; c < 4; ++c) tap(sx, lane + 32 * c, w.w, x0[c], x1[c], a0[c], a1[c]);
    float ra[4], rb[4], acc = 0.f;
    int cy0 = -1, cy1 = -1;
    for (int i = 0; i < 128; ++i) {
      int y0, y1; float l0, l1;
      tap(sy, i, w.h, y0, y1, l0, l1);
      if (y0 != cy0 || y1 != cy1) {
        if (y0 == cy1) { for (int c = 0; c < 4; ++c) ra[c] = rb[c]; }
        else { const float* r = base + (size_t)y0 * PW; for (int c = 0; c < 4; ++c) ra[c] = fmaf(__ldg(r + x0[c]), a0[c], __ldg(r + x1[c]) * a1[c]); }
        if (y1 == y0) { for (int c = 0; c < 4; ++c) rb[c] = ra[c]; }
        else { const float* r = base + (size_t)y1 * PW; for (int c = 0; c < 4; ++c) rb[c] = fmaf(__ldg(r + x0[c]), a0[c], __ldg(r + x1[c]) * a1[c]); }
        cy0 = y0; cy1 = y1;
      }
      for (int c = 0; c < 4; ++c) acc += fmaf(ra[c], l0, rb[c] * l1);
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(~0u, acc, o);
    if (lane == 0) out[wi] = acc;
  }
}

__global__ void __launch_bounds__(256) k_tld4(cudaTextureObject_t tex, const Win* __restrict__ wins, int n, float* out) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int wi = blockIdx.x * wpb + (threadIdx.x >> 5); wi < n; wi += gridDim.x * wpb) {
    const Win w = wins[wi];
    const int ox = (w.p % GX) * PW + w.x, oy = (w.p / GX) * PH + w.y;
    const float sx = (float)w.w / 128.f, sy = (float)w.h / 128.f;
    float fx[4]; bool xe[4]; float a0[4], a1[4];
    for (int c = 0; c < 4; ++c) { int x0, x1; tap(sx, lane + 32 * c, w.w, x0, x1, a0[c], a1[c]); fx[c] = (float)(ox + x0) + 1.0f; xe[c] = x1 == x0; }
    float acc = 0.f;
    for (int i = 0; i < 128; ++i) {
      int y0, y1; float l0, l1;
      tap(sy, i, w.h, y0, y1, l0, l1);
      const float fy = (float)(oy + y0) + 1.0f;
      const bool ye = y1 == y0;
      for (int c = 0; c < 4; ++c) {
        const float4 g = tex2Dgather<float4>(tex, fx[c], fy, 0);   // w=(x0,y0) z=(x1,y0) x=(x0,y1) y=(x1,y1)
        const float v00 = g.w, v01 = xe[c] ? g.w : g.z;
        const float v10 = ye ? v00 : g.x, v11 = ye ? v01 : (xe[c] ? g.x : g.y);
        acc += fmaf(fmaf(v00, a0[c], v01 * a1[c]), l0, fmaf(v10, a0[c], v11 * a1[c]) * l1);
      }
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(~0u, acc, o);
    if (lane == 0) out[wi] = acc;
  }
}

int main() {
  std::vector<float> h((size_t)NP * PH * PW);
  srand(1);
  for (auto& v : h) v = (float)rand() / RAND_MAX - 0.5f;
  float* d; CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  cudaChannelFormatDesc fd = cudaCreateChannelDesc<float>();
  cudaArray_t arr; CK(cudaMallocArray(&arr, &fd, PW * GX, PH * GY, cudaArrayTextureGather));
  for (int p = 0; p < NP; ++p)
    CK(cudaMemcpy2DToArray(arr, (size_t)(p % GX) * PW * 4, (size_t)(p / GX) * PH, d + (size_t)p * PH * PW, PW * 4, PW * 4, PH, cudaMemcpyDeviceToDevice));
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
  cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint;
  td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
  cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  const int sizes[][2] = {{32, 32}, {128, 128}, {200, 200}, {256, 256}, {320, 400}, {480, 640}};
  const int N = 148 * 8 * 16;
  Win* dw; float *o1, *o2; CK(cudaMalloc(&dw, N * sizeof(Win))); CK(cudaMalloc(&o1, N * 4)); CK(cudaMalloc(&o2, N * 4));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (auto& s : sizes) {
    std::vector<Win> w(N);
    for (int i = 0; i < N; ++i) { w[i].p = (i / 64) % NP; w[i].h = s[0]; w[i].w = s[1]; w[i].x = rand() % (PW - s[1] + 1); w[i].y = rand() % (PH - s[0] + 1); }
    CK(cudaMemcpy(dw, w.data(), N * sizeof(Win), cudaMemcpyHostToDevice));
    float t1 = 1e9, t2 = 1e9;
    for (int r = 0; r < 4; ++r) {
      float t;
      cudaEventRecord(a); k_ldg<<<148 * 4, 256>>>(d, dw, N, o1); cudaEventRecord(b); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&t, a, b); if (r) t1 = fminf(t1, t);
      cudaEventRecord(a); k_tld4<<<148 * 4, 256>>>(tex, dw, N, o2); cudaEventRecord(b); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&t, a, b); if (r) t2 = fminf(t2, t);
    }
    std::vector<float> r1(N), r2(N);
    CK(cudaMemcpy(r1.data(), o1, N * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(r2.data(), o2, N * 4, cudaMemcpyDeviceToHost));
    int bad = 0; for (int i = 0; i < N; ++i) bad += r1[i] != r2[i];
    printf("%3dx%3d  ldg %7.3f ms (%5.1f ns/window)   tld4 %7.3f ms (%5.1f ns/window)   ratio %.2f   mismatches %d\n", s[0], s[1], t1, t1 * 1e6 / N, t2,
           t2 * 1e6 / N, t1 / t2, bad);
  }
  return 0;
}
