// Micro-benchmark: cost of executing COLD straight-line code on B200 (instruction fetch), per launch.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/icache.bin tools/ubench/icache.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <vector>
#include <algorithm>
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define R4(x) x x x x
#define R16(x) R4(R4(x))
#define R64(x) R4(R16(x))
#define R256(x) R4(R64(x))
#define R1024(x) R4(R256(x))
// 4 independent chains so that the code is issue/fetch bound, not latency bound
#define STEP "fma.rn.f32 %0, %0, %4, %5;\n fma.rn.f32 %1, %1, %4, %5;\n fma.rn.f32 %2, %2, %4, %5;\n fma.rn.f32 %3, %3, %4, %5;\n"
template <int KI>   // KI * 1024 instructions
__global__ void __launch_bounds__(256) straight(float* out, long long* stamps, float a, float b) {
  const long long t0 = gtime();
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
#pragma unroll
  for (int k = 0; k < KI; ++k) asm volatile(R256(STEP) : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3) : "f"(a), "f"(b));
  const long long t1 = gtime();
  if (x0 + x1 + x2 + x3 == 123.456f) out[0] = x0;
  if (threadIdx.x == 0) { stamps[blockIdx.x * 2] = t0; stamps[blockIdx.x * 2 + 1] = t1 - t0; }
}
__global__ void other(float* p, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = p[i] * 1.5f + 1.f; }
template <int KI> void run(const char* name, float* out, long long* stamps, float* big, int warps_mode) {
  const int grid = 148 * 2;
  for (int mode = 0; mode < 3; ++mode) {   // 0: back-to-back same kernel; 1: another kernel in between; 2: L2 flush in between
    std::vector<float> med;
    for (int it = 0; it < 12; ++it) {
      if (mode == 1) other<<<4096, 256>>>(big, 1 << 20);
      if (mode == 2) cudaMemsetAsync(big, it, 512u << 20);
      straight<KI><<<grid, warps_mode>>>(out, stamps, 1.0001f, 0.5f);
      cudaDeviceSynchronize();
      std::vector<long long> h(grid * 2);
      cudaMemcpy(h.data(), stamps, grid * 16, cudaMemcpyDeviceToHost);
      std::vector<float> d;
      for (int b = 0; b < grid; ++b) d.push_back(h[b * 2 + 1] / 1e3f);
      std::sort(d.begin(), d.end());
      if (it >= 2) med.push_back(d[grid / 2]);
    }
    std::sort(med.begin(), med.end());
    printf("%-10s threads %3d mode %d (%s): per-CTA duration median %.2f us (%.1f ns / 100 instr)\n", name, warps_mode, mode,
           mode == 0 ? "same kernel back to back" : mode == 1 ? "other kernel between" : "512MB memset between",
           med[med.size() / 2], med[med.size() / 2] * 1e3 / (KI * 1024 / 100.0));
  }
}
int main() {
  float *out, *big; long long* stamps;
  cudaMalloc(&out, 4096); cudaMalloc(&stamps, 65536); cudaMalloc(&big, 512u << 20);
  for (int threads : {32, 256}) {
    run<1>("1K instr", out, stamps, big, threads);
    run<4>("4K instr", out, stamps, big, threads);
    run<16>("16K instr", out, stamps, big, threads);
  }
  return 0;
}
