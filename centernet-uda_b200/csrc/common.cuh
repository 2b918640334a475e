// Shared helpers for libcnhead_sm100 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "cnhead.h"

namespace cnh {

constexpr int kThreads = 256;            // every streaming kernel uses 8 warps per CTA
constexpr int kWarps = kThreads / 32;
constexpr float kLo = 1e-4f;             // utils/tensor.py:6 clamp bounds
constexpr float kHi = 1.0f - 1e-4f;

// ---- error plumbing ---------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int sm_count();                          // SMs of the current device (cached)
long long* debug_buffer();               // device buffer for in-kernel stage timestamps, or nullptr

#define CNH_REQUIRE(cond, code, ...)      \
  do {                                    \
    if (!(cond)) {                        \
      ::cnh::set_error(__VA_ARGS__);      \
      return (code);                      \
    }                                     \
  } while (0)

#define CNH_CUDA(expr)                                        \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return ::cnh::cuda_fail(_e, #expr); \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- device helpers ---------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum with a fixed reduction tree (lane tree, then warps in order): the result
// depends only on the per-thread inputs, never on scheduling.  `red` holds kWarps values.
// Every thread returns the total.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();                       // protect `red` from the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  T t = red[0];
#pragma unroll
  for (int w = 1; w < kWarps; ++w) t += red[w];
  return t;
}

// MUFU approximations with flush-to-zero: one instruction each (the non-ftz intrinsics expand to
// 4-6 instructions of denormal handling that this path never needs: probabilities are clamped
// to [1e-4, 1-1e-4] and exp(-x) underflowing to 0 still yields sigmoid = 1).
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_ftz(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2eF = 1.4426950408889634f;
constexpr float kLn2F = 0.6931471805599453f;

// sigmoid pieces.  FAST: ex2/rcp/lg2.approx.ftz (3 MUFU ops per heat-map element);
// accurate: libdevice expf/logf and an IEEE divide.
template <bool FAST>
__device__ __forceinline__ float sigmoidf_(float x) {
  if (FAST) return rcp_ftz(1.0f + ex2_ftz(x * -kLog2eF));
  return 1.0f / (1.0f + expf(-x));
}
template <bool FAST>
__device__ __forceinline__ float logf_(float x) {
  return FAST ? lg2_ftz(x) * kLn2F : logf(x);
}
// log in the unit the fast path accumulates in: log2 (FAST; scaled by ln2 once per thread) or ln.
template <bool FAST>
__device__ __forceinline__ float log_unit(float x) {
  return FAST ? lg2_ftz(x) : logf(x);
}
__device__ __forceinline__ float clamp_prob(float s) { return fminf(fmaxf(s, kLo), kHi); }

// stage timestamps for tools/ (enabled only when cnh_debug_set_buffer() installed a buffer)
__device__ __forceinline__ void dbg_stamp(long long* dbg, int slot) {
  if (dbg != nullptr && threadIdx.x == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    dbg[(long long)blockIdx.x * 16 + slot] = t;
  }
}

__device__ __forceinline__ float4 ldg_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ void stg_stream(float4* p, float4 v) { __stcs(p, v); }

// L2 residency control for the two-pass schedule: data read by phase 0 and re-read by phase 1 is
// loaded with an evict_last policy; everything streamed once uses evict_first.
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ldg_hint(const float4* ptr, unsigned long long policy) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr), "l"(policy));
  return v;
}
__device__ __forceinline__ void stg_hint(float4* ptr, float4 v, unsigned long long policy) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
               :: "l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(policy) : "memory");
}

typedef unsigned long long u64;

// ---- TMA / mbarrier PTX ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one non-blocking test of the phase with `parity` (true: completed)
__device__ __forceinline__ bool mbar_test(u64* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0u;
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, u64* bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

// 1-D bulk copy global -> shared (TMA, SASS UBLKCP): dst/src 16-byte aligned, bytes % 16 == 0
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, unsigned bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// start `bytes` (multiple of 16) at a 16-byte aligned global address on their way into L2
__device__ __forceinline__ void l2_prefetch(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void l2_prefetch_bulk(const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

}  // namespace cnh
