// fp64 latency / throughput probe for the Jacobi tournament design: nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/fp64_lat scripts/micro/fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* o, long long* cyc, double a, double b, int n) {
  double x = a + threadIdx.x;
  const long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = fma(x, b, a);          // dependent chain
  const long long t1 = clock64();
  o[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void thr(double* o, long long* cyc, double a, double b, int n) {
  double x0 = a + threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  const long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a);
    x4 = fma(x4, b, a); x5 = fma(x5, b, a); x6 = fma(x6, b, a); x7 = fma(x7, b, a);
  }
  const long long t1 = clock64();
  o[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void latf(float* o, long long* cyc, float a, float b, int n) {
  float x = a + threadIdx.x;
  const long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = fmaf(x, b, a);
  const long long t1 = clock64();
  o[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void syncs(long long* cyc, int n) {
  const long long t0 = clock64();
  for (int i = 0; i < n; ++i) __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  double* o; long long* c; float* of;
  cudaMalloc(&o, 1 << 20); cudaMalloc(&c, 1024); cudaMalloc(&of, 1 << 20);
  long long h;
  const int n = 4096;
  for (int threads : {32, 128, 512, 1024}) {
    lat<<<1, threads>>>(o, c, 1.0, 0.999, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("DFMA dependent chain, %4d threads: %.1f cycles per DFMA\n", threads, (double)h / n);
    thr<<<1, threads>>>(o, c, 1.0, 0.999, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("DFMA 8 chains/thread,  %4d threads: %.2f cycles per warp-DFMA per SM (=> %.1f lane-FMA/clk/SM)\n", threads,
           (double)h / (n * 8.0 * (threads / 32)), 32.0 * n * 8.0 * (threads / 32) / (double)h);
    latf<<<1, threads>>>(of, c, 1.0f, 0.999f, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("FFMA dependent chain, %4d threads: %.1f cycles per FFMA\n", threads, (double)h / n);
    syncs<<<1, threads>>>(c, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("__syncthreads,        %4d threads: %.1f cycles\n", threads, (double)h / n);
  }
  return 0;
}
