// HBM microbenchmark for the traffic MIX of the fused step: R read streams + W write streams of doubles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/membench tools/membench.cu && /tmp/membench
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

template <int R, int W>
__global__ void __launch_bounds__(256) k_mix(const double* __restrict__ in, double* __restrict__ out, size_t n, size_t stride) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) acc += in[r * stride + i];
#pragma unroll
    for (int w = 0; w < W; ++w) out[w * stride + i] = acc + w;
  }
}
template <int R, int W>
float run(const double* in, double* out, size_t n, size_t stride, int reps) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  int grid = 148 * 8;
  k_mix<R, W><<<grid, 256>>>(in, out, n, stride);
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) k_mix<R, W><<<grid, 256>>>(in, out, n, stride);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double gb = (double)(R + W) * n * 8 / 1e9;
  printf("R=%2d W=%2d : %.3f ms/launch  %.1f GB/s  (%.1f B/elem)\n", R, W, ms / reps, gb / (ms / reps * 1e-3), (R + W) * 8.0);
  return ms / reps;
}
int main() {
  size_t n = (size_t)8192 * 8192, stride = n;
  double *in, *out;
  cudaMalloc(&in, 9 * n * 8); cudaMalloc(&out, 12 * n * 8);
  cudaMemset(in, 0, 9 * n * 8);
  run<1, 1>(in, out, n, stride, 20);
  run<3, 12>(in, out, n, stride, 20);
  run<0, 12>(in, out, n, stride, 20);
  run<0, 1>(in, out, n, stride, 20);
  run<9, 9>(in, out, n, stride, 20);
  run<3, 3>(in, out, n, stride, 20);
  run<12, 12>(in, out, n, stride, 20);
  run<12, 0>(in, out, n, stride, 20);   // (reads only; the compiler keeps them because acc feeds a store guarded below)
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
