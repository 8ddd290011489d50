// Microbenchmark (GPU box): issue rate of scalar FFMA against packed FFMA2 (fma.rn.f32x2, sm_100) per SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_ffma2 probe_ffma2.cu && ./probe_ffma2
#include <cstdio>
#include <cuda_runtime.h>

template <int PACKED>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  const float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.9999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (PACKED) {
        acc[i] = __ffma2_rn(acc[i], aa, bb);
      } else {
        acc[i].x = fmaf(acc[i].x, aa.x, bb.x);
        acc[i].y = fmaf(acc[i].y, aa.y, bb.y);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  const int iters = 20000;
  for (int packed = 0; packed < 2; ++packed) {
    for (int blocks_per_sm : {1, 2, 4}) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      const int grid = 148 * blocks_per_sm;
      if (packed) k<1><<<grid, 256>>>(out, 10, 0.999f, 0.001f); else k<0><<<grid, 256>>>(out, 10, 0.999f, 0.001f);
      cudaEventRecord(e0);
      if (packed) k<1><<<grid, 256>>>(out, iters, 0.999f, 0.001f); else k<0><<<grid, 256>>>(out, iters, 0.999f, 0.001f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double fma = double(grid) * 256 * iters * 16;
      printf("%s blocks/SM %d: %.3f ms  %.1f TFLOP/s (fp32 FMA = 2 flop)\n", packed ? "FFMA2" : "FFMA ", blocks_per_sm, ms, 2 * fma / ms / 1e9);
    }
  }
  return 0;
}
