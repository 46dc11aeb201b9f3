// Micro-benchmark: L2 -> SM read bandwidth (the roofline of the per-proposal kernels on downscaled windows, whose
// fields are L2-resident).  Every CTA streams the whole buffer with ld.global.cg (L1 bypass) float4 loads, many passes.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2bw.bin l2bw.cu
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void __launch_bounds__(512) k_read(const float4* __restrict__ buf, size_t n, int passes, float* out) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int p = 0; p < passes; ++p) {
    // rotate the start per pass so that an SM does not re-read "its" lines
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x + (size_t)p * 7919 * blockDim.x) % n;
    for (size_t c = 0; c < n / stride; ++c) {
      const float4 v = __ldcg(buf + i);
      acc += v.x + v.y + v.z + v.w;
      i += stride; if (i >= n) i -= n;
    }
  }
  if (acc == 12345.678f) out[0] = acc;
}

int main() {
  float* out; CK(cudaMalloc(&out, 4));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int mbs[] = {8, 16, 32, 48, 64, 96, 160, 512};
  for (int mb : mbs) {
    const size_t bytes = (size_t)mb << 20, n = bytes / 16;
    float4* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
    const int passes = mb <= 96 ? 40 : 8;
    float best = 1e9;
    for (int r = 0; r < 4; ++r) {
      cudaEventRecord(a); k_read<<<148 * 4, 512>>>(buf, n, passes, out); cudaEventRecord(b); CK(cudaEventSynchronize(b));
      float t; cudaEventElapsedTime(&t, a, b); if (r) best = t < best ? t : best;
    }
    const double total = (double)(n / (148 * 4 * 512)) * (148 * 4 * 512) * 16.0 * passes;
    printf("%4d MB buffer: %8.3f ms  %8.1f GB/s\n", mb, best, total / best / 1e6);
    cudaFree(buf);
  }
  return 0;
}
