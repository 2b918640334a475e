// Micro-benchmark: how long does a cold 12.6 MB read burst take on B200, by load flavour?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/burst.bin tools/ubench/burst.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>
typedef unsigned long long u64;
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void mbar_init(u64* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, unsigned parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, unsigned bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
constexpr int kChunk = 4096;
// mode 0: TMA bulk, 4 sub-blocks x (x, g); mode 1: LDG.128 x 8 per thread; mode 2: TMA bulk 2 x 16 KB
// mode 3: like 0 but a small dependent load (side[]) issued BEFORE the bulk; mode 4: same, issued AFTER
__global__ void __launch_bounds__(256) burst(const float* x, const float* g, const long long* side, float* out, long long* stamps, int mode) {
  extern __shared__ __align__(128) float sm[];
  __shared__ u64 bar[4];
  const int tid = threadIdx.x, bid = blockIdx.x;
  const long long t0 = gtime();
  const float* xp = x + (size_t)bid * kChunk;
  const float* gp = g + (size_t)bid * kChunk;
  float acc = 0.f;
  long long t_side = 0, sv = 0;
  if (mode == 3 && tid == 32) sv = __ldg(side + bid * 16);
  if (mode == 1) {
    float4 a[4], b[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) { a[v] = __ldcs((const float4*)(xp + v * 1024 + tid * 4)); b[v] = __ldcs((const float4*)(gp + v * 1024 + tid * 4)); }
#pragma unroll
    for (int v = 0; v < 4; ++v) acc += a[v].x + a[v].w + b[v].y + b[v].z;
  } else {
    if (tid == 0) {
      for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (mode == 2) {
        mbar_expect_tx(&bar[0], 32768);
        bulk_load_1d(sm, xp, 16384, &bar[0]);
        bulk_load_1d(sm + kChunk, gp, 16384, &bar[0]);
      } else {
        for (int v = 0; v < 4; ++v) {
          mbar_expect_tx(&bar[v], 8192);
          bulk_load_1d(sm + v * 1024, xp + v * 1024, 4096, &bar[v]);
          bulk_load_1d(sm + kChunk + v * 1024, gp + v * 1024, 4096, &bar[v]);
        }
      }
    }
    if (mode == 4 && tid == 32) sv = __ldg(side + bid * 16);
    if ((mode == 3 || mode == 4) && tid == 32) {
      // dependent second load
      const long long sv2 = __ldg(side + (sv & 1023) * 16 + 8);
      if (sv2 != 0x7fffffffffffll) t_side = gtime();
      acc += (float)sv2;
    }
    __syncthreads();
    for (int v = 0; v < (mode == 2 ? 1 : 4); ++v) {
      mbar_wait(&bar[v], 0);
      if (tid == 0) stamps[bid * 8 + 1 + v] = gtime() - t0;
    }
    for (int v = 0; v < 4; ++v) acc += sm[v * 1024 + tid * 4] + sm[kChunk + v * 1024 + tid * 4 + 1];
  }
  const long long t1 = gtime();
  if (acc == 123.456f) out[bid] = acc;
  if (tid == 0) { stamps[bid * 8 + 0] = t0; stamps[bid * 8 + 5] = t1 - t0; }
  if (tid == 32) stamps[bid * 8 + 6] = t_side ? t_side - t0 : 0;
}
int main() {
  const int n_chunks = 384, sets = 10;
  const size_t n = (size_t)n_chunks * kChunk;
  float *x, *g, *out; long long *side, *stamps;
  cudaMalloc(&x, n * 4 * sets); cudaMalloc(&g, n * 4 * sets); cudaMalloc(&out, 4096 * 4);
  cudaMalloc(&side, 1024 * 16 * 8); cudaMalloc(&stamps, n_chunks * 8 * 8);
  cudaMemset(x, 0, n * 4 * sets); cudaMemset(g, 0, n * 4 * sets); cudaMemset(side, 0, 1024 * 16 * 8);
  float* flush; cudaMalloc(&flush, 512u << 20);
  cudaFuncSetAttribute(burst, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  const char* names[5] = {"tma 8x4KB", "ldg.128 x8", "tma 2x16KB", "tma + side load first", "tma + side load after"};
  for (int mode = 0; mode < 5; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      std::vector<float> starts, ends, sides, first;
      float ms_total = 0;
      for (int it = 0; it < 40; ++it) {
        const int s = it % sets;
        cudaMemsetAsync(flush, it, 512u << 20);     // evict L2
        cudaMemsetAsync(stamps, 0, n_chunks * 64);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        burst<<<n_chunks, 256, 32768>>>(x + n * s, g + n * s, side, out, stamps, mode);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 10) ms_total += ms;
        if (it == 39) {
          std::vector<long long> h(n_chunks * 8);
          cudaMemcpy(h.data(), stamps, n_chunks * 64, cudaMemcpyDeviceToHost);
          long long tmin = h[0];
          for (int b = 0; b < n_chunks; ++b) tmin = std::min(tmin, h[b * 8]);
          for (int b = 0; b < n_chunks; ++b) {
            starts.push_back((h[b * 8] - tmin) / 1e3f);
            ends.push_back((h[b * 8] - tmin + h[b * 8 + 5]) / 1e3f);
            first.push_back((h[b * 8] - tmin + h[b * 8 + 1]) / 1e3f);
            if (h[b * 8 + 6]) sides.push_back((h[b * 8] - tmin + h[b * 8 + 6]) / 1e3f);
          }
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
      }
      std::sort(starts.begin(), starts.end()); std::sort(ends.begin(), ends.end()); std::sort(sides.begin(), sides.end()); std::sort(first.begin(), first.end());
      printf("%-24s rep%d: event %.2f us | start med %.2f max %.2f | first sub med %.2f | all data: min %.2f med %.2f max %.2f us", names[mode], rep,
             ms_total / 30 * 1e3, starts[n_chunks / 2], starts.back(), first[n_chunks / 2], ends[0], ends[n_chunks / 2], ends.back());
      if (!sides.empty()) printf(" | side chain: med %.2f max %.2f", sides[sides.size() / 2], sides.back());
      printf("\n");
    }
  }
  return 0;
}
