// MUFU throughput probe: how many sin.approx / cos.approx / ex2.approx results per clock per SM?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mufu_probe mufu_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, float seed, int iters, long long* cyc) {
  float x[8];
  for (int j = 0; j < 8; ++j) x[j] = seed + threadIdx.x * 1e-3f + j;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float r;
      if (OP == 0) asm volatile("sin.approx.f32 %0, %1;" : "=f"(r) : "f"(x[j]));
      if (OP == 1) asm volatile("ex2.approx.f32 %0, %1;" : "=f"(r) : "f"(x[j]));
      if (OP == 2) { float a, b; asm volatile("sin.approx.f32 %0, %1;" : "=f"(a) : "f"(x[j])); asm volatile("cos.approx.f32 %0, %1;" : "=f"(b) : "f"(x[j])); r = a + b; }
      if (OP == 3) asm volatile("rsqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x[j]));
      x[j] = r * 0.5f + 0.25f;
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int j = 0; j < 8; ++j) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  const char* names[4] = {"sin", "ex2", "sin+cos", "rsqrt"};
  for (int op = 0; op < 4; ++op) {
    for (int threads : {128, 256, 512, 1024}) {
      if (op == 0) k<0><<<148, threads>>>(out, 0.1f, iters, cyc);
      if (op == 1) k<1><<<148, threads>>>(out, 0.1f, iters, cyc);
      if (op == 2) k<2><<<148, threads>>>(out, 0.1f, iters, cyc);
      if (op == 3) k<3><<<148, threads>>>(out, 0.1f, iters, cyc);
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double c = h[0];
      double n = double(iters) * 8 * threads * (op == 2 ? 2 : 1);
      printf("%-8s threads/SM=%4d : %.2f results/clk/SM\n", names[op], threads, n / c);
    }
  }
  return 0;
}
