// DetectionLoss forward+backward for sm_100a: sigmoid+clamp, penalty-reduced focal loss,
// masked gather-L1 heads (plain / sigmoid-angle / RAPiD-periodic), all gradients, in ONE
// persistent launch.  Replaces losses/centernet.py:7-95,98-133,192-223 and
// utils/tensor.py:5-25 of the reference (chains of ~100 eager ATen kernels).
//
// HBM-bound streaming work: every heat-map element is read once (logit + target) and written
// once (clamped probability + gradient).  A heat-map CHUNK (4096 contiguous elements) is moved
// into shared memory by the TMA unit -- 1-D bulk copies (cp.async.bulk, SASS UBLKCP), four 4 KB
// sub-blocks per operand, each with its own mbarrier so that arithmetic starts as soon as the first
// sub-block lands -- and consumed by a small ROLLED loop (one float4 of logits and targets per
// thread and iteration).  The loop body is ~200 instructions: the previous fully unrolled
// register-staged version spent more issue slots stalled on instruction fetch than on memory
// (ncu: stall_no_instruction 7.2 vs long_scoreboard 5.6 per issue at the batch-16 shape).
//
// The gradient needs the batch-wide num_pos, which is only known after everything has been read:
//   * STASH    (single wave: every chunk has its own shared-memory stage): the raw gradient
//               overwrites the logits in the stage, the probabilities are stored at once; the
//               grid barrier's arrival atomic carries the CTA's num_pos; afterwards the gradient
//               is scaled and stored: 16 B/element, the algorithmic floor.  While the bulk loads
//               are in flight the same CTAs do the regression-head work (one item per warp) and
//               zero-fill the dense regression gradient planes.
//   * PRECOUNT (large problems): phase 0 counts num_pos over the target only and records which
//               4 KB sub-blocks of the target hold anything but zeros (gaussian-splat targets are
//               sparse); grid barrier; phase 1 streams the chunks through a TMA ring in REVERSE
//               order (the tail of the target is still in L2) and does not re-read all-zero
//               sub-blocks of the target: ~16 B/element for sparse targets, <= 20 B/element always.
//   * COUNT + MAIN: the same two phases as separate launches with the batch-wide normalisers
//               supplied by the caller (the all-reduce of a sharded run sits between them).
//   * FWD: no gradients (validation under no_grad): 12 B/element.
// Regression heads are warp-granular work items (a 16 KB piece of a dense gradient plane + the
// object slots whose centre falls into it).
//
// Reductions are EXACT and order-independent: every chunk / item partial (a float produced
// by a fixed-shape tree) is converted to 2^-40 fixed point and added with integer atomics
// into a (hi, lo) pair of 64-bit accumulators.  Integer addition is associative, so the
// totals -- and the loss -- are bit-identical for any grid size, any schedule and any
// sharding of the batch over GPUs (a sharded run all-reduces the 24 integers).
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "cand.cuh"

namespace cg = cooperative_groups;

namespace cnh {

constexpr int kChunk = 4096;                  // heat-map elements per chunk (one shared-memory stage)
constexpr int kSubs = 4;                      // sub-blocks per chunk, one mbarrier each
constexpr int kSub = kChunk / kSubs;          // 1024 elements = one float4 per thread
static_assert(kSub == kThreads * 4, "one float4 per thread and sub-block");
constexpr int kMaxStages = 7;                 // 7 x 32 KB = 224 KB of the 227 KB a CTA may own
constexpr int kStreamStages = 3;              // ring depth of the streaming schedule (2 CTAs per SM)
constexpr int kPiece = 4096;                  // floats of one regression plane per work item
constexpr int kSlotsPerLane = 8;              // object slots per lane and round of the synchronous item path
constexpr int kKeep = 5;                      // STASH keeps <= 5 object slots per lane in registers (M <= 160)
constexpr int kQ = CNH_TOTALS / 2;            // quantities: focal, num_pos, 3 x (l1, angle, count), spare
constexpr int kCountUnroll = 3;               // chunks whose target loads are in flight together (phase 0)

enum Mode { M_PRECOUNT = 1, M_MAIN = 2, M_COUNT = 3, M_FWD = 4 };

struct __align__(128) Stage {
  float x[kChunk];                            // logits; STASH: overwritten by the raw gradient
  float g[kChunk];                            // target
};

// Workspace: header, then one sparsity word per chunk.  `parity` selects the live accumulator /
// barrier set.  Ticket modes leave their set zeroed (the last CTA cleans up); STASH leaves it dirty,
// flips parity and cleans the OTHER set, which the launch before it used -- so nothing has to be
// reset on the critical path.
struct WsHeader {
  unsigned parity;
  unsigned done;
  unsigned next;                              // streaming schedules: next chunk ticket
  unsigned epoch;                             // launches that used the peer exchange so far
  unsigned next0;                             // streaming schedules: next chunk-group ticket of the count phase
  unsigned flags_chunks;                      // COUNT left valid sparsity words for this many chunks ...
  unsigned long long flags_gt;                // ... of this target tensor
  unsigned long long bar[2][1 + CNH_MAX_HEADS];  // single wave [parity]: [0] chunk CTAs: arrivals << 32 | num_pos;
                                              // [1 + h] mask-count units of head h: arrivals << 40 | sum(mask_expanded)
  long long acc[2][CNH_TOTALS];               // [q] = hi word, [kQ + q] = lo word
};
constexpr size_t kHeaderBytes = (sizeof(WsHeader) + 255) / 256 * 256;

// One mailbox slot per (parity, source rank), 32 words:
//   [0]      (tag << 32) | num_pos of the source rank   -- sent first, with [25..27]: all the gradients need
//   [1..24]  the source rank's exact totals (counts included)
//   [25..27] (tag << 32) | mask count of head 0..2 of the source rank
//   [31]     tag, stored (release.sys) after the totals
// Words 0 and 25..27 validate themselves (payload and tag in one 64-bit store), so they are plain relaxed
// system-scope stores: no fence sits on the path of the normalisers.  (A release.sys store there waits for
// the SM's outstanding probability stores: measured +4 us per step at N = 2.)
constexpr int kSlotWords = 32;
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

constexpr int kEpochWord = CNH_MAILBOX_EPOCH_WORD;   // local mailbox: number of exchanges completed on it
static_assert(2 * CNH_MAX_PEERS * kSlotWords <= kEpochWord && (kEpochWord + 1) * 8 <= CNH_MAILBOX_BYTES, "mailbox layout");
enum { kPeerTimeoutNormalisers = 1, kPeerTimeoutTotals = 2 };

__device__ __forceinline__ long long now_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// Poll a self-validating word (tag in the upper half) with relaxed system-scope loads until it carries `tag`;
// false when `deadline` (globaltimer) passes first.  The clock is read every 256 polls.
__device__ __forceinline__ bool poll_tagged(const unsigned long long* p, unsigned long long tag, long long deadline,
                                            unsigned long long& v) {
  for (unsigned spins = 1;; ++spins) {
    v = ld_relaxed_sys(p);
    if ((v >> 32) == tag) return true;
    if ((spins & 255u) == 0u && now_ns() > deadline) return false;
  }
}
// Same for a word that IS the tag, with acquire semantics (the totals written before it become visible).
__device__ __forceinline__ bool poll_acquire(const unsigned long long* p, unsigned long long tag, long long deadline) {
  for (unsigned spins = 1;; ++spins) {
    if (ld_acquire_sys(p) == tag) return true;
    if ((spins & 255u) == 0u && now_ns() > deadline) return false;
  }
}
__device__ __forceinline__ void report_peer_timeout(unsigned* status, unsigned code) {
  if (status != nullptr) {
    *reinterpret_cast<volatile unsigned*>(status) = code;
    __threadfence_system();
  }
}

struct Geo {
  int HW;
  long long CHW;
  int cps;                 // chunks per sample
  int n_chunks;            // B * cps
  int ppp;                 // pieces per regression plane
  int item0[CNH_MAX_HEADS + 1];   // first item of each head; head h owns B*D_h*ppp items
  int n_items;
  int small_items;         // n_items < 2^22: item indices are decomposed with float reciprocals
  float inv_ppp, inv_D[CNH_MAX_HEADS];
  int n_count;             // B * n_heads
  int vec_planes;          // regression planes can be zero-filled with float4 stores
  int stash_slots;         // STASH may keep the slots of an item in registers (M <= 32 * kKeep)
  int n_stages;            // shared-memory stages per CTA (STASH: chunks per CTA; streaming: ring depth)
  int chunk_ctas;          // STASH: worker CTAs (the grid has one more: the finaliser)
  int x_delay_ns;          // STASH: pause between issuing the target and the logit copies
  int world, rank;         // peer exchange (world == 1: none)
  int defer_totals;        // peers: post the totals and return; cnh_detloss_peers_finalize receives and sums
  unsigned long long* mailbox[CNH_MAX_PEERS];
  unsigned* status;        // peers: pinned host word (nullable) that receives a non-zero code when a wait times out
  long long timeout_ns;    // peers: bound of every wait on another rank
  WsHeader* hdr;
  unsigned* sparse;        // [n_chunks] bit v: sub-block v of the chunk's target is not all zero
  unsigned* next_b;        // [B] candidate emission: per-sample chunk tickets (zero between launches)
  CandGeo cand;            // candidate emission (cand.cuh): where the peaks of the probability tiles go
  int cand_K;              // top-K the candidates are pruned for (0: no emission)
  long long* dbg;
};

// ---- exact accumulation ---------------------------------------------------------------------
__device__ __forceinline__ void acc_add_fixed(long long* acc, int q, float v) {
  const long long f = __double2ll_rn((double)v * 1099511627776.0);        // 2^40
  atomicAdd(reinterpret_cast<unsigned long long*>(acc + q), (unsigned long long)(f >> 32));
  atomicAdd(reinterpret_cast<unsigned long long*>(acc + kQ + q), (unsigned long long)(f & 0xffffffffll));
}
// Single-wave schedule: every contribution also bumps a counter packed into the upper bits of BOTH words
// (the sums stay far below: |hi| < 2^40, lo < 2^48 for <= 65535 contributions), so a reader that sees the
// expected count in a word knows that word is complete -- no fence, no second signal.
constexpr int kCntHi = 44, kCntLo = 48;
constexpr int kMaxCounted = 60000;
__device__ __forceinline__ void acc_add_counted(long long* acc, int q, float v) {
  const long long f = __double2ll_rn((double)v * 1099511627776.0);        // 2^40
  atomicAdd(reinterpret_cast<unsigned long long*>(acc + q), (unsigned long long)((f >> 32) + (1ll << kCntHi)));
  atomicAdd(reinterpret_cast<unsigned long long*>(acc + kQ + q), (unsigned long long)((f & 0xffffffffll) + (1ll << kCntLo)));
}
// A CTA's partial totals.  Its lo word is carry-normalised first (lo < 2^32, the carry moves to hi: same
// value), so that the sum of the lo words of up to 2^16 contributions stays below the counter bits.
__device__ __forceinline__ void acc_add_pair(long long* acc, int q, long long hi, long long lo) {
  atomicAdd(reinterpret_cast<unsigned long long*>(acc + q), (unsigned long long)(hi + (lo >> 32)));
  atomicAdd(reinterpret_cast<unsigned long long*>(acc + kQ + q), (unsigned long long)(lo & 0xffffffffll));
}
__device__ __forceinline__ void acc_add_pair_counted(long long* acc, int q, long long hi, long long lo) {
  atomicAdd(reinterpret_cast<unsigned long long*>(acc + q), (unsigned long long)(hi + (lo >> 32) + (1ll << kCntHi)));
  atomicAdd(reinterpret_cast<unsigned long long*>(acc + kQ + q), (unsigned long long)((lo & 0xffffffffll) + (1ll << kCntLo)));
}
// published totals are canonical: 0 <= lo < 2^32 for the fixed-point sums (q = 0, 2+3h, 3+3h)
__device__ __forceinline__ void canonicalise_totals(long long* t) {
  const int qs[7] = {0, 2, 3, 5, 6, 8, 9};
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const int q = qs[i];
    const long long lo = t[kQ + q];
    t[q] += lo >> 32;
    t[kQ + q] = lo & 0xffffffffll;
  }
}
__device__ __forceinline__ void acc_add_int(long long* acc, int q, long long n) {
  atomicAdd(reinterpret_cast<unsigned long long*>(acc + kQ + q), (unsigned long long)n);
}
__device__ __host__ __forceinline__ double fixed_to_double(long long hi, long long lo) {
  return ((double)hi * 4294967296.0 + (double)lo) * (1.0 / 1099511627776.0);
}

// ---- focal element ----------------------------------------------------------------------
// term = log(p)(1-p)^2 [gt==1]  or  log(1-p) p^2 (1-gt)^4 [gt<1]   (losses/centernet.py:82-84)
// graw = d(term)/dx with p = clamp(s): in*s(1-s)*d(term)/dp; for in-range s (p == s) this is
//        +((1-s)^3 - 2 s (1-s)^2 log s)            for gt == 1
//        -(1-gt)^4 (s^3 - 2 s^2 (1-s) log(1-s))    for gt <  1
// so one log and no second reciprocal per element.
// FAST accumulates `term` in log2 units (the caller multiplies the per-thread sum by ln2 once).
// log of an argument whose distance from 1, t = 1 - arg, is known exactly.  FAST: lg2.approx carries 2^-22 ABSOLUTE
// error, which for arg -> 1 (log(1 - p) of the many background pixels with small p) is a large RELATIVE error of a
// value that is itself ~ -t.  For t < 2^-5 the series log(1 - t) = -t (1 + t/2 + t^2/3 + t^3/4) (truncation < 2e-7
// relative) is used instead, in the unit the fast path accumulates in (log2): four fused multiply-adds.
constexpr float kNear1 = 0.03125f;
constexpr float kL1 = -1.4426950408889634f, kL2 = kL1 / 2.0f, kL3 = kL1 / 3.0f, kL4 = kL1 / 4.0f;
template <bool FAST>
__device__ __forceinline__ float log_unit_near1(float arg, float t) {
  if (!FAST) return logf(arg);
  const float series = __fmul_rn(t, __fmaf_rn(t, __fmaf_rn(t, __fmaf_rn(t, kL4, kL3), kL2), kL1));
  return t < kNear1 ? series : lg2_ftz(arg);
}

template <bool FAST>
__device__ __forceinline__ void focal_elem(float x, float gt, float& p, float& term, float& graw,
                                           int& npos) {
  const float s = sigmoidf_<FAST>(x);
  p = clamp_prob(s);
  const bool pos = (gt == 1.0f);
  const bool neg = (gt < 1.0f);
  // explicit _rn intrinsics: no FMA contraction, so every schedule (STASH / PRECOUNT / MAIN / FWD)
  // produces bit-identical terms and gradients
  const float q = __fsub_rn(1.0f, p);
  const float arg = pos ? p : q;              // argument of the log
  const float a = pos ? q : p;                // the squared factor (= 1 - arg)
  const float L = log_unit_near1<FAST>(arg, a);
  const float omg = __fsub_rn(1.0f, gt);
  float w4 = __fmul_rn(omg, omg);
  w4 = __fmul_rn(w4, w4);
  const float wgt = pos ? 1.0f : (neg ? w4 : 0.0f);
  const float a2 = __fmul_rn(a, a);
  term = __fmul_rn(__fmul_rn(L, a2), wgt);
  constexpr float k2 = FAST ? 2.0f * kLn2F : 2.0f;       // d/dp of a^2 log(arg), log in its unit
  const float inner = __fmaf_rn(-__fmul_rn(__fmul_rn(k2, a2), arg), L, __fmul_rn(a2, a));
  const float sw = pos ? 1.0f : -wgt;
  graw = (p == s) ? __fmul_rn(sw, inner) : 0.0f;         // clamp passes gradient inclusively
  npos += pos ? 1 : 0;
}

// Same arithmetic for an element known to have gt < 1 (no selects): bit-identical to focal_elem.
template <bool FAST>
__device__ __forceinline__ void focal_elem_neg(float x, float gt, float& p, float& term, float& graw) {
  const float s = sigmoidf_<FAST>(x);
  p = clamp_prob(s);
  const float q = __fsub_rn(1.0f, p);
  const float L = log_unit_near1<FAST>(q, p);
  const float omg = __fsub_rn(1.0f, gt);
  float w4 = __fmul_rn(omg, omg);
  w4 = __fmul_rn(w4, w4);
  const float a2 = __fmul_rn(p, p);
  term = __fmul_rn(__fmul_rn(L, a2), w4);
  constexpr float k2 = FAST ? 2.0f * kLn2F : 2.0f;
  const float inner = __fmaf_rn(-__fmul_rn(__fmul_rn(k2, a2), q), L, __fmul_rn(a2, p));
  graw = (p == s) ? __fmul_rn(-w4, inner) : 0.0f;
}

// ---- packed fp32 pairs (Blackwell FMUL2 / FADD2 / FFMA2): two IEEE round-to-nearest operations per
// instruction and lane -- bit-identical to the scalar _rn intrinsics, half the issue slots on the fma pipe
__device__ __forceinline__ u64 pk(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

__device__ __forceinline__ long long clock_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// barrier of the 8 compute warps (threads 0..255); the single-wave kernel has a ninth warp that must not join
__device__ __forceinline__ void sync_compute() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Two elements known to have gt < 1, same operation sequence as focal_elem_neg on packed pairs:
// bit-identical results.  `term` and `graw` come back as pairs.
template <bool FAST>
__device__ __forceinline__ void focal_pair_neg(float x0, float x1, float g0, float g1, float& p0, float& p1,
                                               u64& term, u64& graw) {
  float s0, s1, L0, L1;
  if (FAST) {
    float t0, t1;
    upk(mul2(pk(x0, x1), pk(-kLog2eF, -kLog2eF)), t0, t1);
    float u0, u1;
    upk(add2(pk(1.0f, 1.0f), pk(ex2_ftz(t0), ex2_ftz(t1))), u0, u1);
    s0 = rcp_ftz(u0);
    s1 = rcp_ftz(u1);
  } else {
    s0 = sigmoidf_<false>(x0);
    s1 = sigmoidf_<false>(x1);
  }
  p0 = clamp_prob(s0);
  p1 = clamp_prob(s1);
  const u64 p = pk(p0, p1);
  const u64 q = add2(pk(1.0f, 1.0f), pk(-p0, -p1));          // 1 - p   (x + (-y) == x - y, exactly)
  float q0, q1;
  upk(q, q0, q1);
  if (FAST) {                                                // log_unit_near1 on the pair (same operations, same bits)
    const u64 ser = mul2(p, fma2(p, fma2(p, fma2(p, pk(kL4, kL4), pk(kL3, kL3)), pk(kL2, kL2)), pk(kL1, kL1)));
    float s0, s1;
    upk(ser, s0, s1);
    L0 = p0 < kNear1 ? s0 : lg2_ftz(q0);
    L1 = p1 < kNear1 ? s1 : lg2_ftz(q1);
  } else {
    L0 = logf(q0);
    L1 = logf(q1);
  }
  const u64 L = pk(L0, L1);
  const u64 omg = add2(pk(1.0f, 1.0f), pk(-g0, -g1));
  u64 w4 = mul2(omg, omg);
  w4 = mul2(w4, w4);
  const u64 a2 = mul2(p, p);
  term = mul2(mul2(L, a2), w4);
  constexpr float k2 = FAST ? 2.0f * kLn2F : 2.0f;
  const u64 m = mul2(mul2(pk(k2, k2), a2), q);
  float m0, m1;
  upk(m, m0, m1);
  const u64 inner = fma2(pk(-m0, -m1), L, mul2(a2, p));
  float w0, w1;
  upk(w4, w0, w1);
  const u64 gr = mul2(pk(-w0, -w1), inner);
  float r0, r1;
  upk(gr, r0, r1);
  graw = pk((p0 == s0) ? r0 : 0.0f, (p1 == s1) ? r1 : 0.0f);
}

__device__ __forceinline__ int block_sum_compute(int v, int* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  sync_compute();
  if (lane == 0) red[warp] = v;
  sync_compute();
  int t = red[0];
#pragma unroll
  for (int w = 1; w < kWarps; ++w) t += red[w];
  return t;
}

// ---- one chunk: shared-memory stage -> probability, loss terms, gradient ---------------------------
struct ChunkRef {
  long long base;          // element offset of the chunk in the heat-map tensors
  int n;                   // valid elements (the last chunk of a sample may be short)
};
__device__ __forceinline__ ChunkRef chunk_ref(const Geo& g, int chunk) {
  const int b = chunk / g.cps, j = chunk - b * g.cps;
  const long long in_sample = (long long)j * kChunk;
  const long long left = g.CHW - in_sample;
  ChunkRef r;
  r.base = (long long)b * g.CHW + in_sample;
  r.n = left < kChunk ? (int)left : kChunk;
  return r;
}

// thread 0: start the bulk copies of one chunk into `st`.  Sub-block v completes bar[v]; a sub-block of
// the target whose bit in `gmask` is clear is known to be all zero and is not fetched.
__device__ __forceinline__ void issue_chunk(const cnh_detloss_args& a, const ChunkRef& r, Stage& st, u64* bar,
                                            unsigned gmask) {
#pragma unroll
  for (int v = 0; v < kSubs; ++v) {
    const int left = r.n - v * kSub;
    if (left <= 0) {                          // short chunk: keep the phases of all four barriers in step
      mbar_arrive(bar + v);
      continue;
    }
    const unsigned bytes = (unsigned)(left < kSub ? left : kSub) * 4u;
    const bool has_g = (gmask >> v) & 1u;
    mbar_expect_tx(bar + v, has_g ? 2u * bytes : bytes);
    bulk_load_1d(st.x + v * kSub, a.hm_logits + r.base + v * kSub, bytes, bar + v);
    if (has_g) bulk_load_1d(st.g + v * kSub, a.hm_gt + r.base + v * kSub, bytes, bar + v);
  }
}

// Shapes the TMA unit cannot move (C*H*W % 4 != 0 or a misaligned base pointer): every thread copies,
// guarded; elements past the end of the chunk become (x = 0, gt = 2), i.e. neither positive nor negative.
__device__ __forceinline__ void fill_stage_sync(const cnh_detloss_args& a, const ChunkRef& r, Stage& st) {
  sync_compute();
  const float* __restrict__ xp = a.hm_logits + r.base;
  const float* __restrict__ gp = a.hm_gt + r.base;
#pragma unroll 4
  for (int i = threadIdx.x; i < kChunk; i += kThreads) {
    const bool ok = i < r.n;
    st.x[i] = ok ? __ldcs(xp + i) : 0.f;
    st.g[i] = ok ? __ldcs(gp + i) : 2.f;
  }
  sync_compute();
}

// Thread-local loss accumulators.  A thread's partial sum over its 16 elements of a chunk (a float,
// fixed order) is converted to 2^-40 fixed point and split into (hi, lo) at once; from there on everything
// is integer addition, so the totals do not depend on which CTA, wave or GPU processed which chunk.
struct LossAcc {
  long long hi, lo;
  int npos;
};
__device__ __forceinline__ void loss_acc_add(LossAcc& la, float v) {
  const long long f = __double2ll_rn((double)v * 1099511627776.0);        // 2^40
  la.hi += f >> 32;
  la.lo += f & 0xffffffffll;
}

// Consume one staged chunk: the probability goes straight to HBM; the gradient too when its scale is
// known (NEED_GRAD && !KEEP), or (KEEP, single wave) it replaces the logits in the stage, unscaled.
// Loss terms and num_pos go to the thread-local accumulators.  No barrier inside: a thread only
// touches its own slots of the stage (WAIT: the stage is filled by bulk copies, wait for each sub-block).
// NOTE (candidate emission): every pixel whose probability reaches the pruning threshold `pend.thr` is noted in the
// chunk's pending list (its offset in the chunk, 16 bits; one shared-memory atomic per noted pixel -- rare once the
// threshold has tightened; four instructions otherwise).  thr == 0: nothing is noted (the emitter scans the whole tile).
constexpr int kPendCap = 256;                 // pending pixels per list; more: the emitter scans the whole tile
constexpr int kPendSlots = 8;                 // lists in flight between the consumers and the emitter (a ring of its own)
constexpr int kBootRows = 4;                  // rows of the CTA's first chunk each consumer warp tests unpruned (8 warps)
constexpr int kEmitSliceKeys = kSliceCap - kPendSlots * kPendCap * 2 / 8;   // keys of a slice; the lists take its tail
struct Pending {
  unsigned short* list;                       // [kPendCap]
  unsigned* count;
  float thr;                                  // 0: do not note
};
template <bool NEED_GRAD, bool FAST, bool WAIT, bool KEEP, bool NOTE = false>
__device__ __forceinline__ void process_chunk(const cnh_detloss_args& a, const ChunkRef& r, Stage& st, u64* bar,
                                              unsigned parity, unsigned gmask, float scale, LossAcc& la,
                                              const Pending pend = Pending{nullptr, nullptr, 0.f}) {
  constexpr bool VEC = WAIT;                  // bulk-copied stages imply 16-byte aligned tensors and n % 4 == 0
  float* __restrict__ pp = a.prob + r.base;
  float* __restrict__ gq = NEED_GRAD ? a.grad_hm + r.base : nullptr;
  float sum = 0.f;
  int npos = 0;
#pragma unroll 1
  for (int v = 0; v < kSubs; ++v) {
    if (v * kSub >= r.n) break;
    const int off = v * kSub + threadIdx.x * 4;
    if (WAIT) mbar_wait(bar + v, parity);
    float4 x4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool live = !VEC || off < r.n;      // VEC: a float4 is all in or all out
    if (live) {
      x4 = *reinterpret_cast<const float4*>(st.x + off);
      if ((gmask >> v) & 1u) g4 = *reinterpret_cast<const float4*>(st.g + off);
    } else {
      g4 = make_float4(2.f, 2.f, 2.f, 2.f);
    }
    const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, gs[4] = {g4.x, g4.y, g4.z, g4.w};
    // positives are rare (one pixel per object): a warp whose targets are all < 1 takes the
    // select-free packed path (same bits, about half the instructions)
    const bool special = !(gs[0] < 1.0f) || !(gs[1] < 1.0f) || !(gs[2] < 1.0f) || !(gs[3] < 1.0f);
    float ps[4], gr[4];
    if (__any_sync(0xffffffffu, special)) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float term;
        focal_elem<FAST>(xs[e], gs[e], ps[e], term, gr[e], npos);
        sum = __fadd_rn(sum, term);
      }
    } else {
      u64 t01, t23, g01, g23;
      focal_pair_neg<FAST>(xs[0], xs[1], gs[0], gs[1], ps[0], ps[1], t01, g01);
      focal_pair_neg<FAST>(xs[2], xs[3], gs[2], gs[3], ps[2], ps[3], t23, g23);
      float t0, t1, t2, t3;
      upk(t01, t0, t1);
      upk(t23, t2, t3);
      sum = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sum, t0), t1), t2), t3);   // same order as the scalar path
      upk(g01, gr[0], gr[1]);
      upk(g23, gr[2], gr[3]);
    }
    if (VEC) {
      if (live) {
        *reinterpret_cast<float4*>(pp + off) = make_float4(ps[0], ps[1], ps[2], ps[3]);   // decode reads it next: default policy
        if (NEED_GRAD && !KEEP) {
          const u64 sc = pk(scale, scale);
          float o0, o1, o2, o3;
          upk(mul2(pk(gr[0], gr[1]), sc), o0, o1);
          upk(mul2(pk(gr[2], gr[3]), sc), o2, o3);
          stg_stream(reinterpret_cast<float4*>(gq + off), make_float4(o0, o1, o2, o3));
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (off + e < r.n) {
          pp[off + e] = ps[e];
          if (NEED_GRAD && !KEEP) gq[off + e] = __fmul_rn(gr[e], scale);
        }
    }
    if (NEED_GRAD && KEEP) *reinterpret_cast<float4*>(st.x + off) = make_float4(gr[0], gr[1], gr[2], gr[3]);
    if (NOTE) {
      // (no vote, no prefix sum: a lane that holds such a pixel draws its own position -- rare once the threshold has
      // tightened, and the lanes that hold none skip the branch in four instructions)
      if (pend.thr > 0.f && fmaxf(fmaxf(ps[0], ps[1]), fmaxf(ps[2], ps[3])) >= pend.thr) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (ps[e] >= pend.thr) {
            const unsigned pos = atomicAdd(pend.count, 1u);
            if (pos < (unsigned)kPendCap) pend.list[pos] = (unsigned short)(off + e);
          }
      }
    }
  }
  if (FAST) sum = __fmul_rn(sum, kLn2F);      // log2 -> natural log, once per thread and chunk
  loss_acc_add(la, sum);
  la.npos += npos;
}

// Sum the thread-local accumulators over the 8 compute warps (integers: exact, any order).  Thread 0 returns the totals.
__device__ __forceinline__ void block_reduce_loss(LossAcc& la, long long* red_l, int* red_i) {
  la.hi = warp_sum(la.hi);
  la.lo = warp_sum(la.lo);
  la.npos = warp_sum(la.npos);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  sync_compute();
  if (lane == 0) {
    red_l[2 * warp] = la.hi;
    red_l[2 * warp + 1] = la.lo;
    red_i[warp] = la.npos;
  }
  sync_compute();
  if (threadIdx.x == 0) {
    long long h = 0, l = 0;
    int n = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      h += red_l[2 * w];
      l += red_l[2 * w + 1];
      n += red_i[w];
    }
    la.hi = h;
    la.lo = l;
    la.npos = n;
  }
}

// Single wave, after the grid barrier: raw gradient (left in the stage by process_chunk<KEEP>) * scale -> HBM.
template <bool VEC>
__device__ __forceinline__ void store_stash(const cnh_detloss_args& a, const ChunkRef& r, const Stage& st, float scale) {
  float* __restrict__ gq = a.grad_hm + r.base;
  const u64 sc = pk(scale, scale);
#pragma unroll 1
  for (int v = 0; v < kSubs; ++v) {
    if (v * kSub >= r.n) break;
    const int off = v * kSub + threadIdx.x * 4;
    const float4 q = *reinterpret_cast<const float4*>(st.x + off);
    if (VEC) {
      if (off < r.n) {
        float o0, o1, o2, o3;
        upk(mul2(pk(q.x, q.y), sc), o0, o1);
        upk(mul2(pk(q.z, q.w), sc), o2, o3);
        stg_stream(reinterpret_cast<float4*>(gq + off), make_float4(o0, o1, o2, o3));
      }
    } else {
      const float qs[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (off + e < r.n) gq[off + e] = __fmul_rn(qs[e], scale);
    }
  }
}

// Single-wave schedule, first pass: num_pos of a staged chunk from its target alone, sub-block by
// sub-block as they land (the logits are still in flight).  Per thread.
template <bool VEC>
__device__ __forceinline__ int count_stage(const ChunkRef& r, const Stage& st, u64* gbar) {
  int npos = 0;
#pragma unroll 1
  for (int v = 0; v < kSubs; ++v) {
    if (v * kSub >= r.n) break;
    const int off = v * kSub + threadIdx.x * 4;
    if (VEC) mbar_wait(gbar + v, 0u);
    if (!VEC || off < r.n) {
      const float4 t = *reinterpret_cast<const float4*>(st.g + off);
      npos += (t.x == 1.f) + (t.y == 1.f) + (t.z == 1.f) + (t.w == 1.f);
    }
  }
  return npos;
}

// thread 0: bulk copies of one operand of a chunk, sub-block v completes bar[v]
__device__ __forceinline__ void issue_operand(const float* src, float* dst, int n, u64* bar) {
#pragma unroll
  for (int v = 0; v < kSubs; ++v) {
    const int left = n - v * kSub;
    if (left <= 0) break;
    const unsigned bytes = (unsigned)(left < kSub ? left : kSub) * 4u;
    mbar_expect_tx(bar + v, bytes);
    bulk_load_1d(dst + v * kSub, src + v * kSub, bytes, bar + v);
  }
}

// Phase 0 of PRECOUNT / COUNT: num_pos of up to kCountUnroll chunks from the target only (their loads
// are in flight together), and the chunks' sparsity words.  Per thread.
template <bool VEC>
__device__ __forceinline__ int count_chunks(const cnh_detloss_args& a, const Geo& g, int first, int stride) {
  int npos = 0;
  if (VEC) {
    float4 q[kCountUnroll][kSubs];
    int n[kCountUnroll];
#pragma unroll
    for (int u = 0; u < kCountUnroll; ++u) {
      const int chunk = first + u * stride;
      n[u] = 0;
      if (chunk < g.n_chunks) {
        const ChunkRef r = chunk_ref(g, chunk);
        n[u] = r.n;
        const float* __restrict__ gp = a.hm_gt + r.base;
#pragma unroll
        for (int v = 0; v < kSubs; ++v) {
          const int off = v * kSub + threadIdx.x * 4;
          q[u][v] = (off < r.n) ? __ldg(reinterpret_cast<const float4*>(gp + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kCountUnroll; ++u) {
      if (n[u] == 0) continue;
      unsigned nz = 0;
#pragma unroll
      for (int v = 0; v < kSubs; ++v) {
        const float4 t = q[u][v];
        npos += (t.x == 1.f) + (t.y == 1.f) + (t.z == 1.f) + (t.w == 1.f);
        // (NaN != 0 is true: a NaN target marks its sub-block as occupied, as it must)
        nz |= (t.x != 0.f || t.y != 0.f || t.z != 0.f || t.w != 0.f) ? (1u << v) : 0u;
      }
      nz = __reduce_or_sync(0xffffffffu, nz);
      if (nz != 0u && (threadIdx.x & 31) == 0) atomicOr(g.sparse + first + u * stride, nz);
    }
  } else {
    for (int u = 0; u < kCountUnroll; ++u) {
      const int chunk = first + u * stride;
      if (chunk >= g.n_chunks) break;
      const ChunkRef r = chunk_ref(g, chunk);
      const float* __restrict__ gp = a.hm_gt + r.base;
      for (int i = threadIdx.x; i < r.n; i += kThreads) npos += (__ldg(gp + i) == 1.f);
    }
  }
  return npos;
}

// ---- masked gather-L1 heads: warp-granular work items -----------------------------------------
__device__ __forceinline__ const cnh_head& head_of(const cnh_detloss_args& a, int h) {
  return h == 0 ? a.heads[0] : (h == 1 ? a.heads[1] : a.heads[2]);   // no local copy of the params
}

__device__ __host__ __forceinline__ bool head_has_angle(const cnh_head& hd) {
  return hd.D == 3 && (hd.angle_mode == CNH_ANGLE_SIGMOID || hd.angle_mode == CNH_ANGLE_PERIODIC);
}
__device__ __host__ __forceinline__ bool head_has_limb(const cnh_head& hd) {
  return hd.angle_mode == CNH_LIMB_SQRT || hd.angle_mode == CNH_LIMB_L1;
}

struct ItemRef {
  int h, b, d, p0, p1;
};
// x / d for 0 <= x < 2^22 with inv = 1.0f / d (exact: the rounding error of the product stays below 0.5 / d)
__device__ __forceinline__ int fast_div(int x, float inv) { return (int)(((float)x + 0.5f) * inv); }
__device__ __forceinline__ ItemRef decode_item(const cnh_detloss_args& a, const Geo& g, int item) {
  ItemRef r;
  r.h = 0;
#pragma unroll
  for (int h = 1; h < CNH_MAX_HEADS; ++h)
    if (h < a.n_heads && item >= g.item0[h]) r.h = h;
  const int local = item - g.item0[r.h];
  const int D = head_of(a, r.h).D;
  int plane, piece;
  if (g.small_items) {
    plane = fast_div(local, g.inv_ppp);
    piece = local - plane * g.ppp;
    r.b = fast_div(plane, r.h == 0 ? g.inv_D[0] : (r.h == 1 ? g.inv_D[1] : g.inv_D[2]));
    r.d = plane - r.b * D;
  } else {
    piece = local % g.ppp;
    plane = local / g.ppp;
    r.d = plane % D;
    r.b = plane / D;
  }
  r.p0 = piece * kPiece;
  r.p1 = min(g.HW, r.p0 + kPiece);
  return r;
}

__device__ __forceinline__ void l1_zero_fill_warp(const cnh_detloss_args& a, const Geo& g, const ItemRef& r) {
  const cnh_head& hd = head_of(a, r.h);
  const int lane = threadIdx.x & 31;
  float* __restrict__ dst = hd.grad + ((long long)r.b * hd.D + r.d) * g.HW;
  if (g.vec_planes) {
    for (int off = r.p0 + lane * 4; off < r.p1; off += 128)
      *reinterpret_cast<float4*>(dst + off) = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    for (int off = r.p0 + lane; off < r.p1; off += 32) dst[off] = 0.f;
  }
}

__device__ __forceinline__ void l1_zero_fill_block(const cnh_detloss_args& a, const Geo& g, const ItemRef& r) {
  const cnh_head& hd = head_of(a, r.h);
  float* __restrict__ dst = hd.grad + ((long long)r.b * hd.D + r.d) * g.HW;
  if (g.vec_planes) {
    for (int off = r.p0 + threadIdx.x * 4; off < r.p1; off += kThreads * 4)
      *reinterpret_cast<float4*>(dst + off) = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    for (int off = r.p0 + threadIdx.x; off < r.p1; off += kThreads) dst[off] = 0.f;
  }
}

// |.| term and d(term)/d(pred*m) (times m*w) of one object slot of one channel.
template <bool FAST>
__device__ __forceinline__ void l1_slot_math(const cnh_head& hd, bool is_angle, float pred, float tgt, float m,
                                             float& val, float& gv) {
  constexpr float kPi = 3.14159265358979323846f;          // float(np.pi)
  constexpr float kHalfPi = 1.57079632679489661923f;      // float(np.pi / 2)
  constexpr float kDeg = 0.017453292519943295f;           // torch.deg2rad constant
  const float pv = pred * m;                              // pred *= mask   (centernet.py:108)
  const float tv = tgt * m;                               // target *= mask (centernet.py:109)
  float gcoef;
  if (!is_angle) {
    const float diff = pv - tv;
    val = fabsf(diff);
    gcoef = (diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f);
  } else {
    const float s = sigmoidf_<FAST>(pv);
    const float sc = clamp_prob(s);
    const float ds = (sc == s) ? s * (1.f - s) : 0.f;
    if (hd.angle_mode == CNH_ANGLE_SIGMOID) {             // centernet.py:112-126
      const float diff = sc - clamp_prob(sigmoidf_<FAST>(tv));
      val = fabsf(diff);
      gcoef = ((diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f)) * ds;
    } else {                                              // centernet.py:203-220
      const float pa = sc * 2.f * kPi - kPi;
      const float ta = tv * kDeg;
      float rem = fmodf((pa - ta) - kHalfPi, kPi);
      if (rem != 0.f && rem < 0.f) rem += kPi;
      const float diff = rem - kHalfPi;
      val = fabsf(diff);
      gcoef = ((diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f)) * 2.f * kPi * ds;
    }
  }
  gv = gcoef * m * (is_angle ? hd.angle_weight : hd.weight);
}

// Limb-length consistency term of a keypoint head (KPSL1Loss, losses/centernet.py:153-187) as seen from channel d
// of one object slot whose centre is `idx`: for every pair (ka, kb) that contains keypoint d/2,
//   pd = sqrt((pa-pb)^2 summed over x,y + 1e4)   (or |.|_1),   td likewise from the targets,   term = |pd - td|
// with pa/pb/ta/tb the MASKED predictions / targets (pred *= mask; target *= mask, :147-148).  `val` receives the
// terms this channel accounts for (each pair once: by the x channel of its first keypoint), `gcoef` the derivative of
// the sum of terms with respect to the masked prediction of channel d.  Same operation order as the reference's fp32.
__device__ __forceinline__ void limb_slot(const cnh_head& hd, const float* __restrict__ maps_b, const float* __restrict__ tgt_row,
                                          const uint8_t* __restrict__ mask_row, int idx, int d, int HW, float& val, float& gcoef) {
  const int j = d >> 1, c = d & 1;
  val = 0.f;
  gcoef = 0.f;
  for (int p = 0; p < hd.n_pairs; ++p) {
    const int ka = __ldg(hd.pairs + 2 * p), kb = __ldg(hd.pairs + 2 * p + 1);
    if (ka != j && kb != j) continue;
    float pa[2], pb[2], ta[2], tb[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int ca = 2 * ka + e, cb = 2 * kb + e;
      const float ma = (float)__ldg(mask_row + ca), mb = (float)__ldg(mask_row + cb);
      pa[e] = __fmul_rn(__ldg(maps_b + (long long)ca * HW + idx), ma);
      pb[e] = __fmul_rn(__ldg(maps_b + (long long)cb * HW + idx), mb);
      ta[e] = __fmul_rn(__ldg(tgt_row + ca), ma);
      tb[e] = __fmul_rn(__ldg(tgt_row + cb), mb);
    }
    const float dxp = __fsub_rn(pa[0], pb[0]), dyp = __fsub_rn(pa[1], pb[1]);
    const float dxt = __fsub_rn(ta[0], tb[0]), dyt = __fsub_rn(ta[1], tb[1]);
    const float dc = c ? dyp : dxp;
    float pd, td, dcoef;
    if (hd.angle_mode == CNH_LIMB_L1) {
      pd = __fadd_rn(fabsf(dxp), fabsf(dyp));
      td = __fadd_rn(fabsf(dxt), fabsf(dyt));
      dcoef = (dc > 0.f) ? 1.f : ((dc < 0.f) ? -1.f : 0.f);
    } else {
      pd = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dxp, dxp), __fmul_rn(dyp, dyp)), 1e4f));
      td = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dxt, dxt), __fmul_rn(dyt, dyt)), 1e4f));
      dcoef = __fdiv_rn(dc, pd);
    }
    const float diff = __fsub_rn(pd, td);
    const float sg = (diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f);
    if (ka == j && c == 0) val += fabsf(diff);
    if (ka == j) gcoef += sg * dcoef;
    if (kb == j) gcoef -= sg * dcoef;
  }
}

// One warp, one item: zero-fill of the piece (ZERO), forward terms of the slots inside it (added to
// acc), and either the scatter itself (SCATTER, inv_denom known) or the slots' (index, coefficient)
// kept in registers for a scatter after the grid barrier (KEEP).  Every lane owns slots lane,
// lane+32, ...  Two dependent memory round trips: {index, mask, target} then {prediction}; the
// zero-fill stores are issued between them so that they never delay a load.
template <bool ZERO, bool FORWARD, bool SCATTER, bool COUNTED, bool FAST>
__device__ __forceinline__ void l1_item_warp(const cnh_detloss_args& a, const Geo& g, long long* acc,
                                             const ItemRef& r, float inv_denom) {
  const cnh_head& hd = head_of(a, r.h);
  const int lane = threadIdx.x & 31;
  const int D = hd.D;
  const bool is_angle = (r.d == 2 && head_has_angle(hd));
  const bool limb = head_has_limb(hd);
  const float* __restrict__ plane = hd.map + ((long long)r.b * D + r.d) * g.HW;
  float* __restrict__ gplane = hd.grad ? hd.grad + ((long long)r.b * D + r.d) * g.HW : nullptr;
  float l1 = 0.f, ang = 0.f;
  for (int k0 = 0; k0 < a.M || (ZERO && k0 == 0); k0 += 32 * kSlotsPerLane) {
    int idx[kSlotsPerLane];                      // H*W < 2^30 (validated); out-of-range centres are ignored
    float mk[kSlotsPerLane], tg[kSlotsPerLane];
    // round trip 1: centre index, mask and target of up to 8 slots per lane, all in flight together
#pragma unroll
    for (int u = 0; u < kSlotsPerLane; ++u) {
      const int k = k0 + u * 32 + lane;
      idx[u] = -1;
      mk[u] = 0.f;
      tg[u] = 0.f;
      if (k < a.M) {
        const long long slot = (long long)r.b * a.M + k;
        const long long i64 = __ldg(a.ind + slot);
        idx[u] = (i64 >= 0 && i64 < (long long)g.HW) ? (int)i64 : -1;
        mk[u] = (float)(hd.elementwise_mask ? __ldg(hd.mask + slot * D + r.d) : __ldg(hd.mask + slot));
        tg[u] = __ldg(hd.target + slot * D + r.d);
      }
    }
    if (ZERO && k0 == 0) {
      l1_zero_fill_warp(a, g, r);
      if (SCATTER) __syncwarp();
    }
    // round trip 2: prediction at the centre
    float pr[kSlotsPerLane];
#pragma unroll
    for (int u = 0; u < kSlotsPerLane; ++u) {
      const bool in = (idx[u] >= r.p0 && idx[u] < r.p1);
      pr[u] = 0.f;
      if (in) pr[u] = __ldg(plane + idx[u]); else idx[u] = -1;
    }
#pragma unroll
    for (int u = 0; u < kSlotsPerLane; ++u) {
      float val = 0.f, gv = 0.f;
      if (idx[u] >= 0) {
        l1_slot_math<FAST>(hd, is_angle, pr[u], tg[u], mk[u], val, gv);
        if (FORWARD) { if (is_angle) ang += val; else l1 += val; }
        if (limb) {                                              // keypoint head: + the limb-length term
          const long long slot = (long long)r.b * a.M + (k0 + u * 32 + lane);
          float lv, lg;
          limb_slot(hd, hd.map + (long long)r.b * D * g.HW, hd.target + slot * D,
                    hd.mask + (hd.elementwise_mask ? slot * D : slot), idx[u], r.d, g.HW, lv, lg);
          if (FORWARD) ang += lv;
          gv += lg * mk[u] * hd.angle_weight;
        }
        if (SCATTER && gv != 0.f) atomicAdd(gplane + idx[u], gv * inv_denom);   // duplicate centres accumulate
      }
    }
  }
  if (FORWARD) {
    l1 = warp_sum(l1);
    ang = warp_sum(ang);
    if (lane == 0) {
      if (COUNTED) {                                         // exactly one contribution per item and word, zero or not
        if (limb) {
          acc_add_counted(acc, 2 + 3 * r.h, l1);
          acc_add_counted(acc, 3 + 3 * r.h, ang);
        } else {
          acc_add_counted(acc, (is_angle ? 3 : 2) + 3 * r.h, is_angle ? ang : l1);
        }
      } else {
        if (l1 != 0.f) acc_add_fixed(acc, 2 + 3 * r.h, l1);
        if (ang != 0.f) acc_add_fixed(acc, 3 + 3 * r.h, ang);
      }
    }
  }
}


// ---- the same item in three steps, for the single-wave schedule: the loads are issued before the
// heat-map bulk copies (anything issued after them queues behind 12 MB of traffic), their results wait
// in registers while the chunk is processed, and the arithmetic runs in the shadow of the grid barrier.
// step 1: centre index, mask and target of the lane's slots
// (nothing here may USE a loaded value: the warp must get past this point without waiting)
__device__ __forceinline__ void item_load_slots(const cnh_detloss_args& a, const ItemRef& r, long long (&i64)[kKeep],
                                                uint8_t (&mraw)[kKeep], float (&tg)[kKeep]) {
  const cnh_head& hd = head_of(a, r.h);
  const int lane = threadIdx.x & 31, D = hd.D;
#pragma unroll
  for (int u = 0; u < kKeep; ++u) {
    const int k = u * 32 + lane;
    i64[u] = -1;
    mraw[u] = 0;
    tg[u] = 0.f;
    if (k < a.M) {
      const long long slot = (long long)r.b * a.M + k;
      i64[u] = __ldg(a.ind + slot);
      mraw[u] = hd.elementwise_mask ? __ldg(hd.mask + slot * D + r.d) : __ldg(hd.mask + slot);
      tg[u] = __ldg(hd.target + slot * D + r.d);
    }
  }
}
// step 2: prediction at the centres that fall into the item's piece (others are dropped)
__device__ __forceinline__ void item_load_pred(const cnh_detloss_args& a, const ItemRef& r, const long long (&i64)[kKeep],
                                               int (&idx)[kKeep], float (&pr)[kKeep], int HW) {
  const cnh_head& hd = head_of(a, r.h);
  const float* __restrict__ plane = hd.map + ((long long)r.b * hd.D + r.d) * HW;
#pragma unroll
  for (int u = 0; u < kKeep; ++u) {
    const bool in = (i64[u] >= (long long)r.p0 && i64[u] < (long long)r.p1);     // p1 <= H*W: out-of-range centres are ignored
    idx[u] = in ? (int)i64[u] : -1;
    pr[u] = 0.f;
    if (in) pr[u] = __ldg(plane + idx[u]);
  }
}
// step 3: forward terms -> acc; afterwards idx[u] >= 0 marks a slot with a gradient, pr[u] = its coefficient
template <bool FAST>
__device__ __forceinline__ void item_math(const cnh_detloss_args& a, long long* acc, const ItemRef& r, int (&idx)[kKeep],
                                          const float (&mk)[kKeep], const float (&tg)[kKeep], float (&pr)[kKeep]) {
  const cnh_head& hd = head_of(a, r.h);
  const bool is_angle = (r.d == 2 && head_has_angle(hd));
  float l1 = 0.f, ang = 0.f;
#pragma unroll
  for (int u = 0; u < kKeep; ++u) {
    float val = 0.f, gv = 0.f;
    if (idx[u] >= 0) {
      l1_slot_math<FAST>(hd, is_angle, pr[u], tg[u], mk[u], val, gv);
      if (is_angle) ang += val; else l1 += val;
    }
    if (gv == 0.f) idx[u] = -1;
    pr[u] = gv;
  }
  l1 = warp_sum(l1);
  ang = warp_sum(ang);
  if ((threadIdx.x & 31) == 0) acc_add_counted(acc, (is_angle ? 3 : 2) + 3 * r.h, is_angle ? ang : l1);
}

// sum(mask_expanded) of one (head, sample): D * sum(mask[b,:]) or sum(mask[b,:,:]); one warp.
__device__ __forceinline__ int count_unit_value(const cnh_detloss_args& a, int unit) {
  const int h = unit / a.B, b = unit - h * a.B;
  const cnh_head& hd = head_of(a, h);
  const int lane = threadIdx.x & 31;
  const int n = hd.elementwise_mask ? a.M * hd.D : a.M;
  const uint8_t* __restrict__ mp = hd.mask + (long long)b * n;
  int c = 0;
  for (int k = lane; k < n; k += 32) c += __ldg(mp + k);
  c = warp_sum(c);
  if (!hd.elementwise_mask) c *= hd.D;
  return c;
}
__device__ __forceinline__ void count_unit_warp(const cnh_detloss_args& a, long long* acc, int unit) {
  const int c = count_unit_value(a, unit);
  if ((threadIdx.x & 31) == 0 && c) acc_add_int(acc, 4 + 3 * (unit / a.B), c);
}

// ---- finalisation -------------------------------------------------------------------------
// 24 accumulator words -> scalars, mirroring the reference's fp32 arithmetic
// (centernet.py:91-95,119-131,213-222,42).  One thread.
__device__ void scalars_from_totals(const cnh_detloss_args& a, const long long* tot, float* out) {
  double q[kQ];
#pragma unroll
  for (int i = 0; i < kQ; ++i) q[i] = 0.0;
  q[0] = fixed_to_double(tot[0], tot[kQ + 0]);
  q[1] = (double)tot[kQ + 1];
#pragma unroll
  for (int h = 0; h < CNH_MAX_HEADS; ++h) {
    q[2 + 3 * h] = fixed_to_double(tot[2 + 3 * h], tot[kQ + 2 + 3 * h]);
    q[3 + 3 * h] = fixed_to_double(tot[3 + 3 * h], tot[kQ + 3 + 3 * h]);
    q[4 + 3 * h] = (double)tot[kQ + 4 + 3 * h];
  }
  const float fsum = (float)q[0], npos = (float)q[1];
  float hm = (npos == 0.f) ? (0.f - fsum) : (0.f - fsum / npos);
  hm *= a.hm_weight;
  float total = hm;
  out[1] = hm;
#pragma unroll
  for (int h = 0; h < CNH_MAX_HEADS; ++h) {
    float l = 0.f;
    if (h < a.n_heads) {
      const cnh_head& hd = a.heads[h];
      const float denom = (float)q[4 + 3 * h] + 1e-4f;
      l = (float)q[2 + 3 * h] / denom * hd.weight;
      if (head_has_angle(hd) || head_has_limb(hd)) l += (float)q[3 + 3 * h] / denom * hd.angle_weight;
      total += l;
    }
    out[2 + h] = l;
  }
  out[0] = total;
  out[5] = npos;
  out[6] = 0.f;
  out[7] = 0.f;
}

// Read the live accumulator set (written by other CTAs: through L2), publish it to a.totals,
// compute the scalars.  Threads 0..23 load one word each; thread 0 finishes.
// local_only: the totals are this SHARD's (a sharded launch): publish them, leave the scalars to the exchange.
__device__ void finalize_from_acc(const cnh_detloss_args& a, const long long* acc, long long* sh_tot, bool local_only = false) {
  if (threadIdx.x < CNH_TOTALS) sh_tot[threadIdx.x] = __ldcg(acc + threadIdx.x);
  __syncthreads();
  if (threadIdx.x == 0) canonicalise_totals(sh_tot);
  __syncthreads();
  if (threadIdx.x < CNH_TOTALS && a.totals != nullptr) a.totals[threadIdx.x] = sh_tot[threadIdx.x];
  if (threadIdx.x == 0 && a.scalars != nullptr && !local_only) scalars_from_totals(a, sh_tot, a.scalars);
}


// ================================================================================================
// Single wave ("STASH" in the flags): every heat-map chunk has its own shared-memory stage in ONE wave
// of CTAs, so each input byte is read exactly once and nothing intermediate is written: 16 B/element.
// 9 warps per CTA.
//   warps 0-7 of a worker: thread 0 issues the bulk copies of the CTA's chunks, TARGETS FIRST; the warps
//             count num_pos from the target sub-blocks as they land, while the logits are still in
//             flight, and arrive on the grid barrier (the arrival atomic carries the count: no fence);
//             as the logits land: probabilities to HBM, loss sums, raw gradient into the stage; zero-fill
//             of this CTA's pieces of the dense regression gradient planes; by then the barrier has
//             resolved: scaled gradient from shared memory to HBM.
//   warp 8 of a worker: regression units (items, mask counts), one per round.  Its two dependent round
//             trips (centre index -> prediction) run beside the chunk pipeline; warp 0 prefetches the
//             first one's lines into L2 before the bulk copies are issued (requests are served roughly
//             in order: anything issued behind 12 MB of bulk traffic returns after it).
//   finaliser (last CTA): the scalars.  The accumulator words carry contribution counters, the mask
//             counts ride in the per-head barrier words: it polls until every word is complete.
// ================================================================================================
constexpr int kStashThreads = kThreads + 32;
constexpr int kStashBars = 2 * kSubs;            // per stage: target sub-blocks, then logit sub-blocks

template <bool FAST, bool VEC>
__global__ void __launch_bounds__(kStashThreads, 3)
detloss_stash_kernel(const cnh_detloss_args a, const Geo g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Stage* const stages = reinterpret_cast<Stage*>(smem_raw);
  __shared__ u64 mbar[kMaxStages * kStashBars];
  __shared__ long long red_l[2 * kWarps];
  __shared__ int red_i[kWarps];
  __shared__ long long sh_tot[CNH_TOTALS];
  __shared__ int sh_norm[1 + CNH_MAX_HEADS];
  __shared__ unsigned sh_hdr[2];
  const int bid = blockIdx.x, grid = gridDim.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool worker = bid < grid - 1;
  const int S = g.n_stages, W = g.chunk_ctas;
  // regression units (items, then mask-count units) are dealt round-robin over the worker CTAs: every SM
  // gets some, and the CTA that zero-fills a piece of a gradient plane is the one that scatters into it
  const int n_other = g.n_items + g.n_count;
  constexpr unsigned long long kCountMask = (1ull << 40) - 1ull;

  // PDL: a dependent launched with programmatic stream serialisation (cnh_scale_inplace, cnh_decode) may be
  // placed on the SMs now; it still blocks in griddepcontrol.wait until this grid has completed.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // ... and this launch may itself have been placed early (programmatic serialisation behind cnh_decode):
  // nothing is read or written before the grid in front of it has completed
  asm volatile("griddepcontrol.wait;" ::: "memory");
  dbg_stamp(g.dbg, 0);
  // every thread that touches the accumulators reads the header itself (all lanes of warps 0 and 8 load
  // the same two words: one request, no shuffle that would make the warp wait before it has issued the rest)
  unsigned par = 0, epoch = 0;
  if (warp == 0 || warp == kWarps) {
    par = __ldcg(&g.hdr->parity) & 1u;
    if (g.world > 1) epoch = (unsigned)__ldcg(g.mailbox[g.rank] + kEpochWord);
  }
  const long long deadline = g.world > 1 ? now_ns() + g.timeout_ns : 0ll;
  auto wait_for = [&](const unsigned long long* bar, int count) -> unsigned {
    unsigned long long v;
    do { v = ld_acquire_u64(bar); } while ((unsigned)(v >> 32) < (unsigned)count);
    return (unsigned)(v & 0xffffffffull);
  };

  if (worker && warp < kWarps) {
    // ================= compute warps =================
    if (warp == 0 && bid < g.n_items && g.stash_slots) {
      // L2 prefetch of the lines the regression warp's first round trip will touch
      const ItemRef r = decode_item(a, g, bid);
      const cnh_head& hd = head_of(a, r.h);
      const int mask_row = a.M * (hd.elementwise_mask ? hd.D : 1);
      const char* p0 = reinterpret_cast<const char*>(a.ind + (long long)r.b * a.M);
      const char* p1 = reinterpret_cast<const char*>(hd.target + (long long)r.b * a.M * hd.D);
      const char* p2 = reinterpret_cast<const char*>(hd.mask + (long long)r.b * mask_row);
      for (int off = lane * 128; off < a.M * 8; off += 4096) asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + off));
      for (int off = lane * 128; off < a.M * hd.D * 4; off += 4096) asm volatile("prefetch.global.L2 [%0];" ::"l"(p1 + off));
      for (int off = lane * 128; off < mask_row; off += 4096) asm volatile("prefetch.global.L2 [%0];" ::"l"(p2 + off));
    }
    if (tid == 0 && VEC) {
      for (int i = 0; i < S * kStashBars; ++i) mbar_init(&mbar[i], 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      for (int r = 0; r < S; ++r) {                          // targets of every stage first ...
        const int chunk = bid + r * W;
        if (chunk >= g.n_chunks) break;
        const ChunkRef cr = chunk_ref(g, chunk);
        issue_operand(a.hm_gt + cr.base, stages[r].g, cr.n, &mbar[r * kStashBars]);
      }
      if (g.x_delay_ns > 0) __nanosleep(g.x_delay_ns);
      for (int r = 0; r < S; ++r) {                          // ... then the logits
        const int chunk = bid + r * W;
        if (chunk >= g.n_chunks) break;
        const ChunkRef cr = chunk_ref(g, chunk);
        issue_operand(a.hm_logits + cr.base, stages[r].x, cr.n, &mbar[r * kStashBars + kSubs]);
      }
    }
    sync_compute();                                          // mbarriers initialised
    // ---- pass 1: num_pos from the targets -----------------------------------------------------------
    int npos = 0;
#pragma unroll 1
    for (int r = 0; r < S; ++r) {
      const int chunk = bid + r * W;
      if (chunk >= g.n_chunks) break;
      const ChunkRef cr = chunk_ref(g, chunk);
      if (!VEC) fill_stage_sync(a, cr, stages[r]);
      npos += count_stage<VEC>(cr, stages[r], &mbar[r * kStashBars]);
    }
    npos = block_sum_compute(npos, red_i);
    dbg_stamp(g.dbg, 1);
    unsigned long long* bar_chunk = &g.hdr->bar[par][0];
    // the arrival carries the count: nothing else has to be ordered before it (no fence)
    if (tid == 0) atomicAdd(bar_chunk, (1ull << 32) | (unsigned long long)(unsigned)npos);
    // ---- pass 2a, before the barrier resolves: probabilities to HBM, loss sums, raw gradient into the
    // stage (the logits arrive while the other CTAs are still counting)
    LossAcc la = {0ll, 0ll, 0};
#pragma unroll 1
    for (int r = 0; r < S; ++r) {
      const int chunk = bid + r * W;
      if (chunk >= g.n_chunks) break;
      process_chunk<true, FAST, VEC, true>(a, chunk_ref(g, chunk), stages[r], &mbar[r * kStashBars + kSubs], 0u, 0xfu, 0.f, la);
    }
    block_reduce_loss(la, red_l, red_i);
    if (tid == 0) acc_add_pair_counted(g.hdr->acc[par], 0, la.hi, la.lo);   // one contribution per worker CTA
    dbg_stamp(g.dbg, 4);
    // zero-fill of this CTA's pieces of the regression gradient planes; warp 8 scatters after it
    const bool want_grad = a.grad_hm != nullptr;             // forward only (validation): no stores after the barrier
    if (want_grad)
      for (int o = bid; o < g.n_items; o += W) l1_zero_fill_block(a, g, decode_item(a, g, o));
    asm volatile("bar.arrive 2, %0;" ::"n"(kStashThreads) : "memory");
    if (!want_grad) return;
    if (g.world > 1) {
      if (warp == 0) {                                       // sharded: num_pos of every rank, straight from the mailbox;
        const unsigned long long tag = (unsigned long long)epoch + 1ull;     // one lane per source rank polls in parallel
        const unsigned long long* box = g.mailbox[g.rank] + (size_t)(tag & 1ull) * CNH_MAX_PEERS * kSlotWords;
        int mine = 0;
        bool ok = true;
        if (lane < g.world) {
          unsigned long long v;
          ok = poll_tagged(box + (size_t)lane * kSlotWords, tag, deadline, v);
          mine = ok ? (int)(unsigned)(v & 0xffffffffull) : 0;
        }
        ok = __all_sync(0xffffffffu, ok);
        mine = warp_sum(mine);
        if (lane == 0) {
          sh_norm[0] = ok ? mine : -1;                       // -1: a peer never arrived -> NaN scale below
          if (!ok && bid == 0) report_peer_timeout(g.status, kPeerTimeoutNormalisers);
        }
      }
    } else if (tid == 0) {
      sh_norm[0] = (int)wait_for(bar_chunk, W);
    }
    sync_compute();
    dbg_stamp(g.dbg, 3);
    // ---- pass 2b: the gradient, scaled, from shared memory ------------------------------------------------
    const int npos_all = sh_norm[0];
    const float scale = (npos_all < 0) ? __int_as_float(0x7fc00000)
                                       : ((npos_all == 0) ? -a.hm_weight : -a.hm_weight / (float)npos_all);
#pragma unroll 1
    for (int r = 0; r < S; ++r) {
      const int chunk = bid + r * W;
      if (chunk >= g.n_chunks) break;
      store_stash<VEC>(a, chunk_ref(g, chunk), stages[r], scale);
    }
    dbg_stamp(g.dbg, 5);
  } else if (worker) {
    // ================= regression warp =================
    const bool dbg8 = g.dbg != nullptr && lane == 0;
    const int first = bid, step = W;
    const bool kept = g.stash_slots && first < g.n_items;    // the first item lives in registers
    ItemRef kr;
    int idx[kKeep];
    long long i64[kKeep];
    uint8_t mraw[kKeep];
    float mk[kKeep], tg[kKeep], pr[kKeep];
    if (kept) {
      kr = decode_item(a, g, first);
      item_load_slots(a, kr, i64, mraw, tg);
    }
    long long* acc = g.hdr->acc[par];
    unsigned long long* bar_cnt = &g.hdr->bar[par][1];
    if (kept) {
      item_load_pred(a, kr, i64, idx, pr, g.HW);
#pragma unroll
      for (int u = 0; u < kKeep; ++u) mk[u] = (float)mraw[u];
      if (dbg8) g.dbg[(long long)bid * 16 + 8] = clock_ns();
      item_math<FAST>(a, acc, kr, idx, mk, tg, pr);
    }
    for (int o = kept ? first + step : first; o < n_other; o += step) {
      if (o < g.n_items) {
        l1_item_warp<false, true, false, true, FAST>(a, g, acc, decode_item(a, g, o), 0.f);
      } else {                                               // mask count of one (head, sample): rides in the barrier word
        const int unit = o - g.n_items, h = unit / a.B;
        const int c = count_unit_value(a, unit);
        if (lane == 0) atomicAdd(bar_cnt + h, (1ull << 40) | (unsigned long long)(unsigned)c);
      }
    }
    if (dbg8) g.dbg[(long long)bid * 16 + 2] = clock_ns();
    // ---- regression gradients: need the batch-wide mask counts and this CTA's zero-fill -----------------
    asm volatile("bar.sync 2, %0;" ::"n"(kStashThreads) : "memory");
    if (a.grad_hm == nullptr) return;                        // forward only: the loss terms are all there is
    if (first < g.n_items) {
      if (g.world > 1) {                                     // the mask counts of every rank arrive with its num_pos:
        const unsigned long long tag = (unsigned long long)epoch + 1ull;   // lane = (source rank, head) polls in parallel
        const unsigned long long* box = g.mailbox[g.rank] + (size_t)(tag & 1ull) * CNH_MAX_PEERS * kSlotWords;
        const int r = lane >> 2, h = lane & 3;               // 8 ranks x 4 (3 heads used)
        int mine = 0;
        bool ok = true;
        if (r < g.world && h < CNH_MAX_HEADS) {
          unsigned long long v;
          ok = poll_tagged(box + (size_t)r * kSlotWords + 25 + h, tag, deadline, v);                    // self-validating words
          mine = ok ? (int)(unsigned)(v & 0xffffffffull) : 0;
        }
        ok = __all_sync(0xffffffffu, ok);
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);    // sum over ranks, per head
        if (lane < CNH_MAX_HEADS) sh_norm[1 + lane] = ok ? mine : -1;
      } else if (lane == 0) {
        {
          const bool all_heads = first + step < g.n_items || !kept;
          for (int h = 0; h < a.n_heads; ++h) {
            if (!all_heads && h != kr.h) continue;
            unsigned long long v;
            do { v = ld_acquire_u64(bar_cnt + h); } while ((int)(v >> 40) < a.B);
            sh_norm[1 + h] = (int)(v & kCountMask);
          }
        }
      }
      __syncwarp();
    }
    if (dbg8) g.dbg[(long long)bid * 16 + 9] = clock_ns();
    for (int o = first; o < g.n_items; o += step) {
      const ItemRef r = (kept && o == first) ? kr : decode_item(a, g, o);
      const float inv = sh_norm[1 + r.h] < 0 ? __int_as_float(0x7fc00000) : 1.f / ((float)sh_norm[1 + r.h] + 1e-4f);
      if (kept && o == first) {                              // slots kept in registers: no reload
        const cnh_head& hd = head_of(a, r.h);
        float* __restrict__ gplane = hd.grad + ((long long)r.b * hd.D + r.d) * g.HW;
#pragma unroll
        for (int u = 0; u < kKeep; ++u)
          if (idx[u] >= 0) atomicAdd(gplane + idx[u], pr[u] * inv);
      } else {
        l1_item_warp<false, false, true, false, FAST>(a, g, acc, r, inv);
      }
    }
    if (dbg8) g.dbg[(long long)bid * 16 + 6] = clock_ns();
  } else {
    // ================= finaliser: the scalars =================
    if (tid == 0) { sh_hdr[0] = par; sh_hdr[1] = epoch; }
    if (tid < CNH_TOTALS) sh_tot[tid] = 0ll;
    __syncthreads();
    par = sh_hdr[0];
    epoch = sh_hdr[1];
    const unsigned long long tag = (unsigned long long)epoch + 1ull;     // peer exchange: this launch's tag
    const unsigned mpar = (unsigned)(tag & 1ull);                        // mailbox parity follows the exchange count
    long long* acc = g.hdr->acc[par];
    // retire the OTHER accumulator set (the launch before this one used it) while waiting
    if (tid < CNH_TOTALS) g.hdr->acc[par ^ 1u][tid] = 0ll;
    if (tid < 1 + CNH_MAX_HEADS) g.hdr->bar[par ^ 1u][tid] = 0ull;
    if (tid == 0) sh_norm[0] = (int)wait_for(&g.hdr->bar[par][0], W);
    if (g.world > 1 && tid >= 1 && tid <= CNH_MAX_HEADS) {   // mask counts (tiny units, long since arrived)
      int c = 0;
      if (tid - 1 < a.n_heads) {
        unsigned long long v;
        do { v = ld_acquire_u64(&g.hdr->bar[par][tid]); } while ((int)(v >> 40) < a.B);
        c = (int)(v & kCountMask);
      }
      sh_norm[tid] = c;
    }
    __syncthreads();
    // first, and alone on the critical path: this rank's num_pos inside the tag word of every peer's slot,
    // preceded by its mask counts (the regression gradients of the peers wait for nothing else)
    if (g.world > 1 && tid < g.world) {
      unsigned long long* slot = g.mailbox[tid] + ((size_t)mpar * CNH_MAX_PEERS + g.rank) * kSlotWords;
      st_relaxed_sys(slot, (tag << 32) | (unsigned long long)(unsigned)sh_norm[0]);
#pragma unroll
      for (int h = 0; h < CNH_MAX_HEADS; ++h)
        st_relaxed_sys(slot + 25 + h, (tag << 32) | (unsigned long long)(unsigned)sh_norm[1 + h]);
    }
    // every sum word validates itself: poll until it holds the expected number of contributions
    if (tid < 14) {
      const int qi = tid % 7, lo = tid / 7;                  // 7 sums (focal, then l1 / angle per head) x (hi, lo)
      const int h = (qi - 1) >> 1, is_ang = (qi - 1) & 1;
      const int q = qi == 0 ? 0 : 2 + 3 * h + is_ang;
      long long expect = W;                                  // focal: one contribution per worker CTA
      if (qi > 0) {
        expect = 0;
        if (h < a.n_heads) {
          const cnh_head& hd = head_of(a, h);
          const int ang_ch = head_has_angle(hd) ? 1 : 0;       // items that contribute to the angle / limb word
          expect = (long long)a.B * g.ppp * (is_ang ? (head_has_limb(hd) ? hd.D : ang_ch) : hd.D - ang_ch);
        }
      }
      long long clean = 0;
      if (expect > 0) {
        const unsigned long long* word = reinterpret_cast<const unsigned long long*>(acc + (lo ? kQ : 0) + q);
        long long v;
        if (lo) { do { v = (long long)ld_acquire_u64(word); } while ((v >> kCntLo) != expect); }
        else { do { v = (long long)ld_acquire_u64(word); } while (((v + (1ll << (kCntHi - 1))) >> kCntHi) != expect); }
        clean = v - (expect << (lo ? kCntLo : kCntHi));
      }
      sh_tot[(lo ? kQ : 0) + q] = clean;
    } else if (tid < 14 + CNH_MAX_HEADS) {                   // mask counts: all B units of the head have arrived
      const int h = tid - 14;
      if (h < a.n_heads) {
        unsigned long long v;
        do { v = ld_acquire_u64(&g.hdr->bar[par][1 + h]); } while ((int)(v >> 40) < a.B);
        sh_tot[kQ + 4 + 3 * h] = (long long)(v & kCountMask);
      }
    } else if (tid == 14 + CNH_MAX_HEADS) {
      sh_tot[kQ + 1] = (long long)(unsigned)sh_norm[0];      // num_pos
    }
    __syncthreads();
    if (tid == 0) canonicalise_totals(sh_tot);
    __syncthreads();
    dbg_stamp(g.dbg, 3);
    if (g.world > 1 && g.defer_totals) {
      // deferred: this rank's totals stay local; cnh_detloss_peers_finalize posts them, receives the peers'
      // and sums.  (Posting them here costs the launch the NVLink write acknowledgements at its very end:
      // measured 4.5 us per step at N = 2, all of it in front of the next kernel of the stream.)
      if (tid < CNH_TOTALS) a.totals[tid] = sh_tot[tid];
      if (tid == 0) { g.mailbox[g.rank][kEpochWord] = tag; g.hdr->parity = par ^ 1u; }
      return;
    }
    if (g.world > 1) {
      // ---- the rest of the exchange: exact totals (counts included), then the second tag --------
      if (tid < g.world * CNH_TOTALS) {                      // thread = (destination rank, word)
        const int dst = tid / CNH_TOTALS, w = tid % CNH_TOTALS;
        g.mailbox[dst][((size_t)mpar * CNH_MAX_PEERS + g.rank) * kSlotWords + 1 + w] = (unsigned long long)sh_tot[w];
      }
      __threadfence_system();
      __syncthreads();
      if (tid < g.world) st_release_sys(g.mailbox[tid] + ((size_t)mpar * CNH_MAX_PEERS + g.rank) * kSlotWords + 31, tag);
      // wait for every source rank's totals in the LOCAL mailbox, then sum (exact integers)
      if (tid == 0) sh_hdr[0] = 1u;                          // every peer's totals arrived
      __syncthreads();
      if (tid < g.world) {
        const unsigned long long* slot = g.mailbox[g.rank] + ((size_t)mpar * CNH_MAX_PEERS + tid) * kSlotWords;
        if (!poll_acquire(slot + 31, tag, deadline)) sh_hdr[0] = 0u;
      }
      __syncthreads();
      if (tid < CNH_TOTALS) {
        long long sum = 0;
        for (int r = 0; r < g.world; ++r)
          sum += (long long)__ldcv(g.mailbox[g.rank] + ((size_t)mpar * CNH_MAX_PEERS + r) * kSlotWords + 1 + tid);
        sh_tot[tid] = sum;
      }
      __syncthreads();
      if (tid == 0) canonicalise_totals(sh_tot);             // sums of canonical (hi, lo) pairs: carry again
      __syncthreads();
      if (tid == 0) g.mailbox[g.rank][kEpochWord] = tag;
      if (sh_hdr[0] == 0u) {                                 // a peer never delivered: poisoned scalars, error word
        if (tid == 0) report_peer_timeout(g.status, kPeerTimeoutTotals);
        if (tid < CNH_SCALARS && a.scalars != nullptr) a.scalars[tid] = __int_as_float(0x7fc00000);
        if (tid == 0) g.hdr->parity = par ^ 1u;
        return;
      }
    }
    if (tid < CNH_TOTALS && a.totals != nullptr) a.totals[tid] = sh_tot[tid];
    if (tid == 0 && a.scalars != nullptr) scalars_from_totals(a, sh_tot, a.scalars);
    if (tid == 0) g.hdr->parity = par ^ 1u;                  // flip: the next launch uses the set cleaned above
  }
  dbg_stamp(g.dbg, 7);
}

// ================================================================================================
// Streaming schedules (PRECOUNT / MAIN / COUNT / FWD): persistent CTAs of 9 warps.
//   warp 8 (producer): draws chunk tickets from a global counter (dynamic scheduling: SMs do not get
//           the same share of the memory system, a static split ends with a 20 % tail), waits for a free
//           stage of the ring, issues the chunk's bulk copies.  The ticket for the next issue is drawn one
//           step ahead, so its round trip is never waited for.
//   warps 0-7 (consumers): sub-block by sub-block as they land; probability + gradient straight to HBM;
//           loss terms into thread-local exact accumulators; one arrival per warp frees the stage.
//           No block barrier and no atomic per chunk.
// ================================================================================================
// EMIT (candidate emission; needs W == 128, H*W % 4096 == 0: a chunk is 32 full rows of one class plane): a tenth warp.
//   consumers: note, per chunk, the pixels whose probability reaches the sample's pruning threshold (a handful once it
//           has tightened) in a small shared-memory list; they free the stage themselves, as without emission.
//   warp 9 (emitter): runs the 3x3 peak test on the noted pixels, reading the probabilities the consumers have just
//           written to global memory back through L2 (no shared-memory stage is held: the kernel is issue-bound and a
//           lone warp gets a fraction of an SM sub-partition's issue slots -- whatever it does must not sit between a
//           stage and its refill); forwards the peaks to this CTA's slice of the sample's candidate list, counts them
//           into the sample's histogram, keeps the threshold current (cand.cuh: CandEmitter).
//   Every CTA serves ONE sample (CTA i: sample i % B, chunk tickets per sample), so a slice holds one sample's keys.
// The decode (cnh_decode_candidates) then never reads the heat map: 4*C*H*W bytes per sample and a launch less.
constexpr int kEmitThreads = kStashThreads + 32;
template <int MODE, bool FAST, bool VEC, bool EMIT = false>
__global__ void __launch_bounds__(EMIT ? kEmitThreads : kStashThreads, 2)
detloss_stream_kernel(const cnh_detloss_args a, const Geo g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Stage* const stages = reinterpret_cast<Stage*>(smem_raw);
  __shared__ u64 full[kStreamStages * kSubs];   // (static shared memory is kept under 1 KB: two CTAs of three stages
                                               // then fit the 196 KB carve-out, and the L1 keeps 60 KB for the spills)
  __shared__ u64 empty[kStreamStages];
  __shared__ long long red_l[2 * kWarps];
  __shared__ int red_i[kWarps];
  __shared__ long long sh_tot[CNH_TOTALS];
  __shared__ unsigned sh_ticket, sh_parity;
  __shared__ long long sh_gnorm[4];           // peers: batch-wide mask counts of head 0..2, num_pos (-1: a peer timed out)
  __shared__ unsigned long long sh_tag;
  __shared__ unsigned sh_mask[kStreamStages];
  __shared__ int sh_chunk[kStreamStages];
  // EMIT: a ring of kPendSlots pending lists between the consumers and the emitter (slot = chunk count % kPendSlots)
  __shared__ u64 scanned[kPendSlots];          // the eight consumer warps are done with the slot's chunk
  __shared__ u64 pend_free[kPendSlots];        // the emitter has taken the slot's list
  __shared__ int sh_pchunk[kPendSlots];        // the chunk the list belongs to (-1: no more)
  __shared__ unsigned sh_npend[kPendSlots];    // pending pixels of the chunk (may run past kPendCap)
  __shared__ unsigned sh_fullscan[kPendSlots]; // some warp had no threshold yet: the emitter scans the whole tile
  __shared__ float sh_thr;                     // EMIT: the emitter's current threshold, read by the consumers per chunk
  __shared__ unsigned sh_boot_cnt;             // EMIT: keys of the CTA's first chunk (tested and forwarded by the consumers)
  const int bid = blockIdx.x, grid = gridDim.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool consumer = warp < kWarps;
  constexpr bool kGrad = (MODE == M_PRECOUNT || MODE == M_MAIN);
  constexpr int S = kStreamStages;
  // EMIT: this CTA's sample and slice (CTAs beyond B * G draw no chunks)
  const int my_b = EMIT ? bid % a.B : 0, my_j = EMIT ? bid / a.B : 0;
  const bool has_sample = !EMIT || my_j < g.cand.G;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: see detloss_stash_kernel
  asm volatile("griddepcontrol.wait;" ::: "memory");                // (cooperative launches carry the PDL attribute)
  dbg_stamp(g.dbg, 0);
  if (tid == 0) {
    if (MODE != M_COUNT && VEC) {
      for (int i = 0; i < S * kSubs; ++i) mbar_init(&full[i], 1);
      for (int i = 0; i < S; ++i) mbar_init(&empty[i], kWarps);
      if (EMIT) {
        for (int i = 0; i < kPendSlots; ++i) {
          mbar_init(&scanned[i], kWarps);
          mbar_init(&pend_free[i], 1);
          sh_npend[i] = 0u;
          sh_fullscan[i] = 0u;
        }
        sh_thr = 0.f;
        sh_boot_cnt = 0u;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    sh_parity = __ldcg(&g.hdr->parity) & 1u;   // ticket modes leave their set zeroed: parity stays what it is
  }
  __syncthreads();
  long long* acc = g.hdr->acc[sh_parity];

  float scale = 0.f;
  float inv_denom[CNH_MAX_HEADS] = {0.f, 0.f, 0.f};
  if (MODE == M_PRECOUNT || MODE == M_COUNT) {
    // ---- phase 0: normalisers and sparsity words from the targets only --------------------------
    int npos = 0;
    if (consumer) {
      // groups of kCountUnroll consecutive chunks, drawn dynamically (ticket for the next group requested
      // before this group's loads are consumed)
      const unsigned n_groups = (unsigned)((g.n_chunks + kCountUnroll - 1) / kCountUnroll);
      unsigned t_next = 0;
      if (tid == 0) t_next = atomicAdd(&g.hdr->next0, 1u);
      for (;;) {
        if (tid == 0) sh_ticket = t_next;
        sync_compute();
        const unsigned t = sh_ticket;
        sync_compute();
        if (t >= n_groups) break;
        if (tid == 0) t_next = atomicAdd(&g.hdr->next0, 1u);
        npos += count_chunks<VEC>(a, g, (int)t * kCountUnroll, 1);
      }
      npos = block_sum_compute(npos, red_i);
      if (tid == 0 && npos) acc_add_int(acc, 1, npos);
      for (int u = (grid - 1 - bid) * kWarps + warp; u < g.n_count; u += grid * kWarps) count_unit_warp(a, acc, u);
    }
    dbg_stamp(g.dbg, 1);
  }
  if (MODE == M_PRECOUNT) {
    cg::this_grid().sync();
    long long npos = __ldcg(acc + kQ + 1);
    long long cnt[CNH_MAX_HEADS];
#pragma unroll
    for (int h = 0; h < CNH_MAX_HEADS; ++h) cnt[h] = __ldcg(acc + kQ + 4 + 3 * h);
    if (g.world > 1) {
      // ---- sharded: the normalisers of every rank, traded through the peer-mapped mailboxes right here ----
      // CTA 0 stores this shard's four words (self-validating: tag in the upper half, relaxed system-scope
      // stores, no fence) into every peer's mailbox; warp 0 of EVERY CTA polls the local mailbox, one lane per
      // (source rank, word), and sums over the ranks.  The totals follow in cnh_detloss_peers_finalize.
      if (warp == 0) {
        const unsigned long long tag = __ldcg(g.mailbox[g.rank] + kEpochWord) + 1ull;
        const unsigned mpar = (unsigned)(tag & 1ull);
        const long long deadline = now_ns() + g.timeout_ns;
        const int r = lane >> 2, q = lane & 3;               // q = 0..2: mask count of head q (word 25 + q); 3: num_pos (word 0)
        if (bid == 0 && r < g.world) {
          const long long mine = q == 3 ? npos : (q == 0 ? cnt[0] : (q == 1 ? cnt[1] : cnt[2]));
          unsigned long long* slot = g.mailbox[r] + ((size_t)mpar * CNH_MAX_PEERS + g.rank) * kSlotWords;
          st_relaxed_sys(slot + (q == 3 ? 0 : 25 + q), (tag << 32) | (unsigned long long)(unsigned)mine);
        }
        long long got = 0;
        bool ok = true;
        if (r < g.world) {
          const unsigned long long* slot = g.mailbox[g.rank] + ((size_t)mpar * CNH_MAX_PEERS + r) * kSlotWords;
          unsigned long long v;
          ok = poll_tagged(slot + (q == 3 ? 0 : 25 + q), tag, deadline, v);
          got = ok ? (long long)(unsigned)(v & 0xffffffffull) : 0ll;
        }
        ok = __all_sync(0xffffffffu, ok);
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) got += __shfl_xor_sync(0xffffffffu, got, o);      // sum over ranks, per word
        if (lane < 4) sh_gnorm[lane] = ok ? got : -1ll;
        if (lane == 0) sh_tag = tag;
        if (!ok && bid == 0 && lane == 0) report_peer_timeout(g.status, kPeerTimeoutNormalisers);
      }
      __syncthreads();
      npos = sh_gnorm[3];
#pragma unroll
      for (int h = 0; h < CNH_MAX_HEADS; ++h) cnt[h] = sh_gnorm[h];
    }
    const float kNaN = __int_as_float(0x7fc00000);
    scale = npos < 0 ? kNaN : ((npos == 0) ? -a.hm_weight : -a.hm_weight / (float)npos);
#pragma unroll
    for (int h = 0; h < CNH_MAX_HEADS; ++h) inv_denom[h] = cnt[h] < 0 ? kNaN : 1.f / ((float)cnt[h] + 1e-4f);
    dbg_stamp(g.dbg, 2);
  }
  if (MODE == M_MAIN) {
    const float npos = (float)a.norm[0];
    scale = (npos == 0.f) ? -a.hm_weight : -a.hm_weight / npos;
#pragma unroll
    for (int h = 0; h < CNH_MAX_HEADS; ++h) inv_denom[h] = 1.f / ((float)a.norm[1 + h] + 1e-4f);
  }
  if (MODE != M_COUNT) {
    // ---- the streaming pass -----------------------------------------------------------------------
    // ticket t -> chunk: reverse order after a pre-count (the tail of the target is still in L2)
    // (EMIT: tickets are per sample, t counts this sample's chunks)
    const int n_tickets = EMIT ? (has_sample ? g.cps : 0) : g.n_chunks;
    unsigned* const ticket_ctr = EMIT ? g.next_b + my_b : &g.hdr->next;
    auto chunk_of = [&](int t) {
      if (EMIT) return my_b * g.cps + ((MODE == M_PRECOUNT) ? g.cps - 1 - t : t);
      return (MODE == M_PRECOUNT) ? g.n_chunks - 1 - t : t;
    };
    // EMIT: chunk -> its 32 x 128 tile of probabilities: rows [32*ty, 32*ty+32) of class plane c of sample my_b
    const int tiles_per_plane = a.H / kCandRows;
    auto tile_of = [&](int chunk) {
      const int jc = chunk - my_b * g.cps;                   // = c * tiles_per_plane + ty
      GTile t;
      t.p = a.prob + (long long)chunk * kChunk;
      t.flat0 = (unsigned)jc * (unsigned)kChunk;
      t.y0 = (jc % tiles_per_plane) * kCandRows;
      t.H = a.H;
      t.HW = g.HW;
      return t;
    };
    // EMIT: the pending lists live in the last 4 KB of the CTA's slice of the candidate workspace (global memory, read
    // back through L2 like the tile: a handful of 16-bit entries per chunk)
    unsigned short* const pend_lists =
        EMIT ? reinterpret_cast<unsigned short*>(g.cand.slices + ((long long)my_b * g.cand.G + (has_sample ? my_j : 0)) * kSliceCap + kEmitSliceKeys)
             : nullptr;
    LossAcc la = {0ll, 0ll, 0};
    // regression units first (the last CTAs get them): the chunk tickets are dynamic, a CTA that is busy
    // here simply draws fewer chunks -- nothing is left to do after the streaming loop
    if (consumer) {
      const int n_other = g.n_items + (MODE == M_PRECOUNT ? 0 : g.n_count);
      for (int o = (grid - 1 - bid) * kWarps + warp; o < n_other; o += grid * kWarps) {
        if (o < g.n_items) {
          const ItemRef r = decode_item(a, g, o);
          if (kGrad) {
            const float inv = r.h == 0 ? inv_denom[0] : (r.h == 1 ? inv_denom[1] : inv_denom[2]);
            l1_item_warp<true, true, true, false, FAST>(a, g, acc, r, inv);
          } else {
            l1_item_warp<false, true, false, false, FAST>(a, g, acc, r, 0.f);
          }
        } else {
          count_unit_warp(a, acc, o - g.n_items);
        }
      }
    }
    if (!VEC) {
      // shapes the TMA unit cannot move: static split, every consumer thread copies (correctness path)
      if (consumer)
        for (int c = bid; c < g.n_chunks; c += grid) {
          const ChunkRef cr = chunk_ref(g, chunk_of(c));
          fill_stage_sync(a, cr, stages[0]);
          process_chunk<kGrad, FAST, false, false>(a, cr, stages[0], full, 0u, 0xfu, scale, la);
        }
    } else if (warp == kWarps) {
      // ---- producer ----
      if (lane == 0) {
        // sparsity words: PRECOUNT wrote them in phase 0; MAIN may use the ones a preceding COUNT left
        // for this very target; FWD has none.  Whoever reads a word clears it (self-cleaning workspace).
        bool sparse_ok = (MODE == M_PRECOUNT);
        if (MODE == M_MAIN)
          sparse_ok = __ldcg(&g.hdr->flags_chunks) == (unsigned)g.n_chunks &&
                      __ldcg(&g.hdr->flags_gt) == (unsigned long long)(uintptr_t)a.hm_gt;
        // software pipeline: the ticket for issue j+2 and the sparsity word for issue j+1 are requested
        // while issue j is prepared -- no round trip of the (loaded) memory system is ever waited for
        auto load_mask = [&](unsigned t) -> unsigned {
          unsigned m = 0xfu;
          if (sparse_ok && t < (unsigned)n_tickets)
            asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(m) : "l"(g.sparse + chunk_of((int)t)));
          return m;
        };
        // (a CTA that serves no sample must not draw: a ticket drawn is a chunk nobody else will take)
        auto draw = [&]() -> unsigned { return has_sample ? atomicAdd(ticket_ctr, 1u) : 0xffffffffu; };
        unsigned t_a = draw();
        unsigned t_b = draw();
        unsigned m_a = load_mask(t_a);
        for (int j = 0;; ++j) {
          const int s = j % S;
          const unsigned t_c = draw();
          const unsigned m_b = load_mask(t_b);
          if (t_b < (unsigned)n_tickets) {                                   // next issue's chunk: start it towards L2 now
            const ChunkRef nr = chunk_ref(g, chunk_of((int)t_b));
            const unsigned bytes = (unsigned)nr.n * 4u;
            l2_prefetch_bulk(a.hm_logits + nr.base, bytes);
            if (!sparse_ok) l2_prefetch_bulk(a.hm_gt + nr.base, bytes);
          }
          long long tw0 = 0;
          if (g.dbg != nullptr) tw0 = clock_ns();
          if (j >= S) mbar_wait(&empty[s], (unsigned)(((j / S) - 1) & 1));   // every consumer warp is done with the stage
          if (g.dbg != nullptr) g.dbg[(long long)bid * 16 + 10] += clock_ns() - tw0;
          if (t_a >= (unsigned)n_tickets) {                                  // out of work: wake the consumers
            sh_chunk[s] = -1;
            mbar_arrive(&full[s * kSubs]);
            break;
          }
          const int chunk = chunk_of((int)t_a);
          if (sparse_ok) g.sparse[chunk] = 0u;
          sh_chunk[s] = chunk;
          sh_mask[s] = m_a;
          issue_chunk(a, chunk_ref(g, chunk), stages[s], &full[s * kSubs], m_a);
          t_a = t_b;
          t_b = t_c;
          m_a = m_b;
        }
      }
    } else if (EMIT && warp == kWarps + 1) {
      // ---- emitter ----
      // The CTA's FIRST chunk is tested whole, unpruned, by the consumer warps (below: eight warps in parallel; this warp
      // runs at ~10 cycles per dependent instruction, see the kernel's header, and would need 10 us for it): by then the
      // histogram holds one chunk of every CTA of the sample, ~5 % of it, and the first threshold already leaves a few
      // dozen pixels per chunk.  From there on this warp works a ring of kPendSlots pending lists behind the consumers,
      // up to kBatch chunks per pass: the lists of a batch are flattened over the lanes (one noted pixel per lane and
      // trip), so the instructions of a pass are shared by several chunks.
      CandEmitter em;
      em.init(g.cand, my_b, has_sample ? my_j : 0, g.cand_K);
      em.cap = (unsigned)kEmitSliceKeys;
      constexpr int kBatch = 4;
      auto release = [&](int s) {                            // the slot's list is taken: the consumers may reuse it
        if (lane == 0) { sh_npend[s] = 0u; sh_fullscan[s] = 0u; mbar_arrive(&pend_free[s]); }
      };
      bool done = false, repruned = false;
      constexpr int kRepruneAt = 6;
#pragma unroll 1
      for (int i = 0; !done;) {
        // ---- gather a batch: chunks i .. i + nb - 1 whose lists are short; a chunk that needs a full scan goes alone ----
        int nb = 0, slot_q[kBatch];
        GTile tile_q[kBatch];
        unsigned cnt_q[kBatch], total = 0;
        bool alone = false;
#pragma unroll
        for (int q = 0; q < kBatch; ++q) {
          if (nb != q) break;                                // (the batch ended before q)
          const int s = (i + q) % kPendSlots;
          const unsigned ph = (unsigned)(((i + q) / kPendSlots) & 1);
          // (only `scanned` is waited on: the producer may refill a stage -- and complete its `full` barrier a second
          // time -- before this warp gets here; the consumers cannot run more than kPendSlots chunks ahead: pend_free)
          if (q == 0) {
            const long long tw = (g.dbg != nullptr) ? clock64() : 0;
            mbar_wait(&scanned[s], ph);
            if (g.dbg != nullptr && lane == 0) g.dbg[(long long)bid * 16 + 12] += clock64() - tw;
          } else if (!mbar_test(&scanned[s], ph)) {
            break;                                           // later chunks join the batch only if they are ready
          }
          const int chunk = sh_pchunk[s];
          if (chunk < 0) { done = true; break; }
          const unsigned n_pend = *reinterpret_cast<volatile unsigned*>(&sh_npend[s]);
          const bool whole = (i + q) == 0 || *reinterpret_cast<volatile unsigned*>(&sh_fullscan[s]) != 0u ||
                             n_pend > (unsigned)kPendCap;
          if (whole && q > 0) break;                         // (waits for the next pass)
          tile_q[q] = tile_of(chunk);
          slot_q[q] = s;
          cnt_q[q] = whole ? 0u : n_pend;
          total += cnt_q[q];
          nb = q + 1;
          if (whole) { alone = true; break; }
        }
        if (nb == 0) break;                                  // (the sentinel came first)
        const long long tw1 = (g.dbg != nullptr) ? clock64() : 0;
        if (alone && i == 0) {
          // the consumers have tested and forwarded the first chunk: take over the slice, wait for the first threshold
          // (two L2 round trips, once -- the sample's other CTAs are at the same point: let their REDs land first)
          em.local_cnt = min(*reinterpret_cast<volatile unsigned*>(&sh_boot_cnt), (unsigned)kEmitSliceKeys);
          __nanosleep(600);
          em.refresh_blocking();
          if (lane == 0) *reinterpret_cast<volatile float*>(&sh_thr) = __uint_as_float(em.thr);
          release(slot_q[0]);
          if (kBootRows * kWarps < kCandRows) gtile_scan(em, tile_q[0], kBootRows * kWarps);   // the rest of the chunk, pruned
        } else if (alone) {
          release(slot_q[0]);
          gtile_scan(em, tile_q[0], 0);
        } else {
          // ---- the batch's noted pixels, flattened: entry e of the batch = entry (e - first) of chunk q ----
          const float thr_eff = fmaxf(__uint_as_float(em.thr), __uint_as_float(1u));
          for (unsigned e0 = 0; e0 < total; e0 += 32) {
            const unsigned e = e0 + (unsigned)lane;
            GTile t = tile_q[0];
            int slot = slot_q[0];
            unsigned first = 0, end = cnt_q[0];
#pragma unroll
            for (int q = 1; q < kBatch; ++q)
              if (q < nb && e >= end) { first = end; end += cnt_q[q]; t = tile_q[q]; slot = slot_q[q]; }
            const bool has = e < total;
            const int off = has ? (int)__ldcg(pend_lists + slot * kPendCap + (int)(e - first)) : 0;
            gtile_test_push(em, t, has, off, thr_eff);
          }
          __syncwarp();
#pragma unroll
          for (int q = 0; q < kBatch; ++q)
            if (q < nb) release(slot_q[q]);
        }
        if (g.dbg != nullptr && lane == 0) {
          g.dbg[(long long)bid * 16 + (alone ? 13 : 11)] += clock64() - tw1;
          g.dbg[(long long)bid * 16 + 15] += alone ? (1ll << 32) : 1ll;
          g.dbg[(long long)bid * 16 + 14] = (long long)em.local_cnt | ((long long)total << 32);
        }
        i += nb;
        // The slice once against the threshold as soon as that has settled somewhat: nearly all of what the slice holds
        // then are the unpruned keys of the first chunk (the keys added later passed a threshold close to the final one:
        // no second pass on the way out, where the whole CTA would wait for it).
        if (!repruned && i >= kRepruneAt && em.thr != 0u) {
          em.reprune([](u64) {});
          repruned = true;
        }
        // the threshold: one step of the pipelined refresh per pass (loads issued at the previous pass have landed by
        // now: nothing is waited for)
        em.refresh_step<1>(i, true);
        if (lane == 0) *reinterpret_cast<volatile float*>(&sh_thr) = __uint_as_float(em.thr);
        __syncwarp();
      }
      if (has_sample) {
        if (!repruned) em.reprune([](u64) {});
        if (lane == 0) {
          g.cand.cta_cnt[(long long)my_b * g.cand.G + my_j] = em.local_cnt;
          if (em.overflow) g.cand.state[my_b].overflow = 1u;
        }
      }
    } else if (consumer) {
      // ---- consumers ----
      dbg_stamp(g.dbg, 7);
#pragma unroll 1
      for (int i = 0;; ++i) {
        const int s = i % S;
        const unsigned ph = (unsigned)((i / S) & 1);
        long long tw0 = 0;
        if (g.dbg != nullptr && tid == 0) tw0 = clock64();
        mbar_wait(&full[s * kSubs], ph);                     // acquires sh_chunk / sh_mask of this fill
        if (g.dbg != nullptr && tid == 0) { g.dbg[(long long)bid * 16 + 8] += clock64() - tw0; g.dbg[(long long)bid * 16 + 9] += 1; }
        const int chunk = sh_chunk[s];
        const int ps = i % kPendSlots;                       // EMIT: this chunk's slot in the ring of pending lists
        if (EMIT) {
          // the emitter has taken the list this slot held kPendSlots chunks ago (it lags that far only rarely); then the
          // chunk is handed to it by name -- sh_chunk[s] may be overwritten by the next refill before it looks
          if (i >= kPendSlots) {
            const long long tw = (g.dbg != nullptr && tid == 0) ? clock64() : 0;
            mbar_wait(&pend_free[ps], (unsigned)(((i / kPendSlots) - 1) & 1));
            if (g.dbg != nullptr && tid == 0) g.dbg[(long long)bid * 16 + 6] += clock64() - tw;
          }
          if (tid == 0) sh_pchunk[ps] = chunk;
          if (chunk < 0) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&scanned[ps]);        // wake the emitter: no more chunks
          }
        }
        if (chunk < 0) break;
        const unsigned m = sh_mask[s];
        Pending pend = {nullptr, nullptr, 0.f};
        if (EMIT && i > 0) {
          // Every warp takes the emitter's threshold as it stands (no agreement needed: it only rises, so whatever the
          // emitter still wants when it gets to this tile has been noted by every warp).  A warp that finds none (the
          // first threshold is still on its way, or fewer than K peaks have been counted so far) notes nothing and tells
          // the emitter to scan the whole tile.
          pend.list = pend_lists + ps * kPendCap;
          pend.count = &sh_npend[ps];
          pend.thr = *reinterpret_cast<volatile float*>(&sh_thr);
          if (pend.thr == 0.f && lane == 0) sh_fullscan[ps] = 1u;
        }
        const long long tp0 = (g.dbg != nullptr && tid == 0) ? clock64() : 0;
        process_chunk<kGrad, FAST, true, false, EMIT>(a, chunk_ref(g, chunk), stages[s], &full[s * kSubs], ph, m, scale, la, pend);
        if (g.dbg != nullptr && tid == 0) g.dbg[(long long)bid * 16 + 5] += clock64() - tp0;
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (EMIT) {
          if (i == 0) {
            // the CTA's first chunk: once all of it is out, every warp tests four of its rows against no threshold at
            // all and forwards the peaks (read back through L2; positions in the slice from a shared-memory counter)
            const long long tb0 = (g.dbg != nullptr && tid == 0) ? clock64() : 0;
            sync_compute();
            CandEmitter boot;
            boot.init(g.cand, my_b, my_j, g.cand_K);
            boot.shared_cnt = &sh_boot_cnt;
            boot.cap = (unsigned)kEmitSliceKeys;
            gtile_rows_unpruned<kBootRows>(boot, tile_of(chunk), kBootRows * warp);
            if (boot.overflow && lane == 0) g.cand.state[my_b].overflow = 1u;
            __syncwarp();
            if (g.dbg != nullptr && tid == 0) g.dbg[(long long)bid * 16 + 6] += clock64() - tb0;
          }
          if (lane == 0) mbar_arrive(&scanned[ps]);          // this warp's probabilities, pending pixels (and keys) of the chunk are out
        }
      }
    }
    dbg_stamp(g.dbg, 3);
    if (consumer) {
      block_reduce_loss(la, red_l, red_i);
      if (tid == 0) {
        acc_add_pair(acc, 0, la.hi, la.lo);
        if (MODE != M_PRECOUNT && la.npos) acc_add_int(acc, 1, la.npos);     // PRECOUNT counted in phase 0
      }
    }
  }

  // ---- last CTA: totals, scalars, leave the accumulator set zeroed -------------------------
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    sh_ticket = atomicAdd(&g.hdr->done, 1u);
  }
  __syncthreads();
  if (sh_ticket != (unsigned)(grid - 1)) return;
  __threadfence();
  if (MODE == M_COUNT) {
    if (tid == 0) {
      a.norm_out[0] = (double)__ldcg(acc + kQ + 1);
      for (int h = 0; h < CNH_MAX_HEADS; ++h) a.norm_out[1 + h] = (double)__ldcg(acc + kQ + 4 + 3 * h);
      g.hdr->flags_chunks = VEC ? (unsigned)g.n_chunks : 0u;
      g.hdr->flags_gt = (unsigned long long)(uintptr_t)a.hm_gt;
    }
  } else {
    finalize_from_acc(a, acc, sh_tot, /*local_only=*/MODE == M_PRECOUNT && g.world > 1);
    if (tid == 0 && MODE == M_MAIN) g.hdr->flags_chunks = 0u;
    // sharded pre-count launch: the exchange is complete on this rank (cnh_detloss_peers_finalize trades the totals)
    if (tid == 0 && MODE == M_PRECOUNT && g.world > 1) g.mailbox[g.rank][kEpochWord] = sh_tag;
  }
  __syncthreads();
  if (tid < CNH_TOTALS) acc[tid] = 0ll;
  if (tid == 0) {
    g.hdr->done = 0;
    g.hdr->next = 0;
    g.hdr->next0 = 0;
  }
  if (EMIT)
    for (int i = tid; i < a.B; i += blockDim.x) g.next_b[i] = 0u;
  dbg_stamp(g.dbg, 4);
}

__global__ void __launch_bounds__(32)
detloss_finalize_kernel(const cnh_detloss_args a, const long long* __restrict__ totals) {
  if (threadIdx.x == 0) scalars_from_totals(a, totals, a.scalars);
}

// Deferred half of the peer exchange (one warp): post this rank's totals of the LAST cnh_detloss_fused_peers
// launch on this workspace (tag = hdr->epoch; a.totals holds them) into every peer's mailbox, wait for the
// peers', sum (exact integers) and produce the global totals + scalars.
__global__ void __launch_bounds__(32)
detloss_peers_finalize_kernel(const cnh_detloss_args a, const Geo g) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __shared__ long long sh_tot[CNH_TOTALS];
  const int tid = threadIdx.x;
  const unsigned long long tag = __ldcg(g.mailbox[g.rank] + kEpochWord);
  const unsigned mpar = (unsigned)(tag & 1ull);
  const long long deadline = now_ns() + g.timeout_ns;
  if (tid < CNH_TOTALS) {
    const unsigned long long mine = (unsigned long long)a.totals[tid];
    for (int r = 0; r < g.world; ++r)
      g.mailbox[r][((size_t)mpar * CNH_MAX_PEERS + g.rank) * kSlotWords + 1 + tid] = mine;
  }
  __threadfence_system();
  __syncwarp();
  bool ok = true;
  if (tid < g.world) {
    st_release_sys(g.mailbox[tid] + ((size_t)mpar * CNH_MAX_PEERS + g.rank) * kSlotWords + 31, tag);
    const unsigned long long* slot = g.mailbox[g.rank] + ((size_t)mpar * CNH_MAX_PEERS + tid) * kSlotWords;
    ok = poll_acquire(slot + 31, tag, deadline);
  }
  ok = __all_sync(0xffffffffu, ok);
  if (!ok) {                                                 // a peer never delivered: poisoned scalars, error word
    if (tid == 0) report_peer_timeout(g.status, kPeerTimeoutTotals);
    if (tid < CNH_SCALARS && a.scalars != nullptr) a.scalars[tid] = __int_as_float(0x7fc00000);
    return;
  }
  if (tid < CNH_TOTALS) {
    long long sum = 0;
    for (int r = 0; r < g.world; ++r)
      sum += (long long)__ldcv(g.mailbox[g.rank] + ((size_t)mpar * CNH_MAX_PEERS + r) * kSlotWords + 1 + tid);
    sh_tot[tid] = sum;
  }
  __syncwarp();
  if (tid == 0) canonicalise_totals(sh_tot);                 // sums of canonical (hi, lo) pairs: carry again
  __syncwarp();
  if (tid < CNH_TOTALS) a.totals[tid] = sh_tot[tid];
  if (tid == 0 && a.scalars != nullptr) scalars_from_totals(a, sh_tot, a.scalars);
}

// g *= factor (factor read from device scalars); nothing to do when factor == 1.
// Launched with programmatic stream serialisation: the launch overlaps the tail of the loss kernel.
__global__ void __launch_bounds__(kThreads)
scale_inplace_kernel(const cnh_scale_args s) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next PDL launch (decode) may be placed now
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (int t = 0; t < s.n_tensors; ++t) {
    const float f = (s.fa[t] ? __ldg(s.fa[t]) : 0.f) + (s.fb[t] ? __ldg(s.fb[t]) : 0.f);
    if (f == 1.0f) continue;
    float* __restrict__ p = s.data[t];
    const long long n = s.count[t];
    const long long stride = (long long)gridDim.x * kThreads;
    long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
      const long long n4 = n >> 2;
      float4* p4 = reinterpret_cast<float4*>(p);
      long long k = i;
      for (; k + 3 * stride < n4; k += 4 * stride) {       // four independent 16-byte loads in flight per thread
        float4 v0 = p4[k], v1 = p4[k + stride], v2 = p4[k + 2 * stride], v3 = p4[k + 3 * stride];
        v0.x *= f; v0.y *= f; v0.z *= f; v0.w *= f;
        v1.x *= f; v1.y *= f; v1.z *= f; v1.w *= f;
        v2.x *= f; v2.y *= f; v2.z *= f; v2.w *= f;
        v3.x *= f; v3.y *= f; v3.z *= f; v3.w *= f;
        p4[k] = v0; p4[k + stride] = v1; p4[k + 2 * stride] = v2; p4[k + 3 * stride] = v3;
      }
      for (; k < n4; k += stride) {
        float4 v = p4[k];
        v.x *= f; v.y *= f; v.z *= f; v.w *= f;
        p4[k] = v;
      }
      for (long long k = (n4 << 2) + i; k < n; k += stride) p[k] *= f;
    } else {
      for (long long k = i; k < n; k += stride) p[k] *= f;
    }
  }
}

// ---- host side -----------------------------------------------------------------------------
static int validate(const cnh_detloss_args* a, bool need_grad_ptrs) {
  CNH_REQUIRE(a != nullptr, CNH_E_NULL, "detloss: args is NULL");
  CNH_REQUIRE(a->B > 0 && a->C > 0 && a->H > 0 && a->W > 0 && a->M >= 0, CNH_E_SHAPE,
              "detloss: bad dims B=%d C=%d H=%d W=%d M=%d", a->B, a->C, a->H, a->W, a->M);
  CNH_REQUIRE((long long)a->H * a->W < (1ll << 30), CNH_E_SHAPE, "detloss: H*W too large");
  CNH_REQUIRE((long long)a->B * a->C * a->H * a->W / kChunk + a->B < (1ll << 30), CNH_E_SHAPE,
              "detloss: heat map too large");
  CNH_REQUIRE(a->n_heads >= 0 && a->n_heads <= CNH_MAX_HEADS, CNH_E_SHAPE, "detloss: n_heads=%d",
              a->n_heads);
  CNH_REQUIRE(a->hm_logits && a->hm_gt && a->prob, CNH_E_NULL, "detloss: hm_logits/hm_gt/prob is NULL");
  CNH_REQUIRE(a->n_heads == 0 || a->M == 0 || a->ind != nullptr, CNH_E_NULL, "detloss: ind is NULL");
  for (int h = 0; h < a->n_heads; ++h) {
    const cnh_head& hd = a->heads[h];
    CNH_REQUIRE(hd.map != nullptr, CNH_E_NULL, "detloss: head %d map is NULL", h);
    CNH_REQUIRE(a->M == 0 || (hd.target && hd.mask), CNH_E_NULL, "detloss: head %d target/mask is NULL", h);
    CNH_REQUIRE(hd.D > 0 && hd.D <= 1024, CNH_E_SHAPE, "detloss: head %d D=%d", h, hd.D);
    CNH_REQUIRE(hd.angle_mode >= CNH_ANGLE_NONE && hd.angle_mode <= CNH_LIMB_L1, CNH_E_UNSUPPORTED,
                "detloss: head %d angle_mode=%d", h, hd.angle_mode);
    if (head_has_limb(hd))
      CNH_REQUIRE(hd.D % 2 == 0 && hd.n_pairs > 0 && hd.pairs != nullptr, CNH_E_SHAPE,
                  "detloss: head %d limb term needs D = 2*nk (got %d) and a pairs table (n_pairs=%d)", h, hd.D, hd.n_pairs);
    CNH_REQUIRE(!(hd.angle_mode == CNH_ANGLE_PERIODIC && hd.D != 3), CNH_E_SHAPE,
                "detloss: periodic angle loss needs a 3-channel head (got D=%d)", hd.D);
    if (need_grad_ptrs) {
      const bool any = a->grad_hm != nullptr;
      CNH_REQUIRE((hd.grad != nullptr) == any, CNH_E_NULL,
                  "detloss: grad pointers must be all set or all NULL (head %d)", h);
    }
  }
  return CNH_OK;
}

static long long n_chunks_of(const cnh_detloss_args* a) {
  const long long CHW = (long long)a->C * a->H * a->W;
  return (long long)a->B * ((CHW + kChunk - 1) / kChunk);
}

static size_t sparse_bytes(const cnh_detloss_args* a) { return ((size_t)n_chunks_of(a) * sizeof(unsigned) + 255) / 256 * 256; }

static Geo make_geo(const cnh_detloss_args* a, void* ws) {
  Geo g;
  g.HW = a->H * a->W;
  g.CHW = (long long)a->C * g.HW;
  g.cps = (int)((g.CHW + kChunk - 1) / kChunk);
  g.n_chunks = a->B * g.cps;
  g.ppp = (g.HW + kPiece - 1) / kPiece;
  int it = 0;
  bool vec_planes = (g.HW % 4 == 0);
  for (int h = 0; h <= CNH_MAX_HEADS; ++h) {
    g.item0[h] = it;
    if (h < a->n_heads) {
      it += a->B * a->heads[h].D * g.ppp;
      vec_planes = vec_planes && aligned16(a->heads[h].grad);
    }
  }
  g.n_items = it;
  g.small_items = it < (1 << 22) ? 1 : 0;
  g.inv_ppp = 1.0f / (float)g.ppp;
  for (int h = 0; h < CNH_MAX_HEADS; ++h) g.inv_D[h] = h < a->n_heads ? 1.0f / (float)a->heads[h].D : 1.0f;
  g.n_count = a->B * a->n_heads;
  g.vec_planes = vec_planes ? 1 : 0;
  g.stash_slots = (a->M <= 32 * kKeep) ? 1 : 0;
  for (int h = 0; h < a->n_heads; ++h)
    if (head_has_limb(a->heads[h])) g.stash_slots = 0;       // the limb term lives in the generic item path only
  g.n_stages = kStreamStages;
  g.chunk_ctas = 0;
  static const int x_delay = getenv("CNH_X_DELAY_NS") ? atoi(getenv("CNH_X_DELAY_NS")) : 0;
  g.x_delay_ns = x_delay;
  g.world = 1;
  g.defer_totals = 0;
  g.rank = 0;
  g.status = nullptr;
  g.timeout_ns = 2000ll * 1000000ll;
  for (int i = 0; i < CNH_MAX_PEERS; ++i) g.mailbox[i] = nullptr;
  g.hdr = static_cast<WsHeader*>(ws);
  g.sparse = reinterpret_cast<unsigned*>(static_cast<char*>(ws) + kHeaderBytes);
  g.next_b = reinterpret_cast<unsigned*>(static_cast<char*>(ws) + kHeaderBytes + sparse_bytes(a));
  g.cand_K = 0;
  g.cand = cand_geo(nullptr, a->B, 0);
  g.dbg = debug_buffer();
  return g;
}

static size_t ws_bytes(const cnh_detloss_args* a) {
  return kHeaderBytes + sparse_bytes(a) + ((size_t)a->B * sizeof(unsigned) + 255) / 256 * 256;   // + per-sample tickets
}

static bool use_vec(const cnh_detloss_args* a, const Geo& g) {
  return (g.CHW % 4 == 0) && aligned16(a->hm_logits) && aligned16(a->hm_gt) && aligned16(a->prob) &&
         (a->grad_hm == nullptr || aligned16(a->grad_hm));
}

static const void* pick_stash(bool fast, bool vec) {
  if (fast) return vec ? (const void*)detloss_stash_kernel<true, true> : (const void*)detloss_stash_kernel<true, false>;
  return vec ? (const void*)detloss_stash_kernel<false, true> : (const void*)detloss_stash_kernel<false, false>;
}
template <int MODE>
static const void* pick_stream_emit(bool fast) {          // candidate emission: bulk-staged (VEC) shapes only
  return fast ? (const void*)detloss_stream_kernel<MODE, true, true, true> : (const void*)detloss_stream_kernel<MODE, false, true, true>;
}
template <int MODE>
static const void* pick_stream(bool fast, bool vec) {
  if (fast) return vec ? (const void*)detloss_stream_kernel<MODE, true, true> : (const void*)detloss_stream_kernel<MODE, true, false>;
  return vec ? (const void*)detloss_stream_kernel<MODE, false, true> : (const void*)detloss_stream_kernel<MODE, false, false>;
}

// resident CTAs of `kernel` with `stages` shared-memory stages, on the current device (cached)
static int resident_ctas(const void* kernel, int stages, int threads = kThreads) {
  struct Entry { const void* k; int dev, stages, ctas; };
  static thread_local Entry cache[64];
  static thread_local int n_cache = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  for (int i = 0; i < n_cache; ++i)
    if (cache[i].k == kernel && cache[i].dev == dev && cache[i].stages == stages) return cache[i].ctas;
  const size_t smem = (size_t)stages * sizeof(Stage);
  int per_sm = 0;
  // opt in to as much dynamic shared memory as the kernel's static part leaves of the 227 KB a CTA may own
  cudaFuncAttributes fa;
  size_t dyn_max = kMaxStages * sizeof(Stage);
  if (cudaFuncGetAttributes(&fa, kernel) == cudaSuccess && fa.sharedSizeBytes + dyn_max > (size_t)227 * 1024)
    dyn_max = (size_t)227 * 1024 - fa.sharedSizeBytes;
  if (smem > dyn_max ||
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_max) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 0;
  }
  const int ctas = per_sm * sm_count();
  if (n_cache < 64) cache[n_cache++] = Entry{kernel, dev, stages, ctas};
  return ctas;
}

static int launch(const void* kernel, bool cooperative, int grid, int stages, const cnh_detloss_args* a, const Geo& g,
                  cudaStream_t stream, int threads = kThreads) {
  void* params[2] = {const_cast<cnh_detloss_args*>(a), const_cast<Geo*>(&g)};
  const size_t smem = (size_t)stages * sizeof(Stage);
  if (cooperative) {
    // cooperative (co-residency for the grid barrier) + programmatic stream serialisation: the launch latency
    // and the CTA ramp-up hide under the tail of the kernel in front (decode of the previous step); the kernel
    // starts with griddepcontrol.wait.  Falls back to the plain cooperative launch if the driver refuses.
    static bool pdl_ok = (getenv("CNH_NO_PDL") == nullptr);
    if (pdl_ok) {
      cudaLaunchConfig_t lc;
      memset(&lc, 0, sizeof(lc));
      lc.gridDim = dim3(grid);
      lc.blockDim = dim3(threads);
      lc.dynamicSmemBytes = smem;
      lc.stream = stream;
      cudaLaunchAttribute attr[2];
      attr[0].id = cudaLaunchAttributeCooperative;
      attr[0].val.cooperative = 1;
      attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[1].val.programmaticStreamSerializationAllowed = 1;
      lc.attrs = attr;
      lc.numAttrs = 2;
      if (cudaLaunchKernelExC(&lc, kernel, params) == cudaSuccess) return CNH_OK;
      cudaGetLastError();
      pdl_ok = false;
    }
    CNH_CUDA(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(threads), params, smem, stream));
  } else
    CNH_CUDA(cudaLaunchKernel(kernel, dim3(grid), dim3(threads), params, smem, stream));
  return CNH_OK;
}

// STASH plan: the smallest number of stages per CTA with which one wave holds every chunk
static bool plan_stash(const void* kernel, Geo& g) {
  if (g.n_items > kMaxCounted) return false;                 // contribution counters are 16 bits
  for (int S = 1; S <= kMaxStages; ++S) {
    const long long cap = (long long)resident_ctas(kernel, S, kStashThreads) - 1;     // one CTA is the finaliser
    if (cap >= 1 && cap * S >= g.n_chunks) {
      g.n_stages = S;
      g.chunk_ctas = (int)(cap < g.n_chunks ? cap : g.n_chunks);
      return true;
    }
  }
  return false;
}

static size_t cand_ws_upper(int B) {
  int G = (3 * sm_count()) / (B > 0 ? B : 1);
  if (G < 1) G = 1;
  return cand_ws_bytes(B, G);
}

// Candidate emission for a streaming launch of `grid` CTAs: possible when a chunk is 32 full rows of one class plane
// (W == 128, H*W % 4096 == 0), the tensors are bulk-staged, every sample gets at least one CTA and the caller's
// workspace is large enough.  Fills a->cand (host) and g.cand / g.cand_K; false: the launch emits nothing (cand->G = 0).
static bool setup_emission(const cnh_detloss_args* a, Geo& g, bool vec, int grid) {
  cnh_cand* c = a->cand;
  if (c == nullptr) return false;
  c->G = 0;
  c->B = a->B; c->C = a->C; c->H = a->H; c->W = a->W;
  if (!vec || a->W != 128 || g.HW % kChunk != 0 || c->workspace == nullptr || c->K < 1 || c->K > 1024) return false;
  const int G = grid / a->B;
  if (G < 1 || cand_ws_bytes(a->B, G) > c->workspace_bytes) return false;
  g.cand = cand_geo(c->workspace, a->B, G);
  g.cand_K = c->K;
  c->G = G;
  return true;
}

static int stream_grid(const void* kernel, const Geo& g, long long warp_units, int stages = kStreamStages,
                       int threads = kStashThreads) {
  long long want = g.n_chunks + (warp_units + kWarps - 1) / kWarps;
  if (want < 1) want = 1;
  int cap = resident_ctas(kernel, stages, threads);
  if (cap < 1) cap = 1;
  return (int)(want < cap ? want : cap);
}

// one streaming launch (PRECOUNT / MAIN / FWD), with candidate emission when the caller asked for it and it is possible
template <int MODE>
static int launch_stream(const cnh_detloss_args* a, Geo& g, bool fast, bool vec, bool cooperative, cudaStream_t st) {
  const long long units = g.n_items + g.n_count;
  if (a->cand != nullptr) {
    const void* ke = pick_stream_emit<MODE>(fast);
    const int grid = stream_grid(ke, g, units, kStreamStages, kEmitThreads);
    // (exactly G CTAs per sample: a CTA beyond B * G would serve no sample, and the tickets its producer draws ahead
    // would be lost to the sample it is numbered into)
    if (setup_emission(a, g, vec, grid))
      return launch(ke, cooperative, a->cand->G * a->B, kStreamStages, a, g, st, kEmitThreads);
  }
  const void* k = pick_stream<MODE>(fast, vec);
  return launch(k, cooperative, stream_grid(k, g, units), kStreamStages, a, g, st, kStashThreads);
}

}  // namespace cnh

using namespace cnh;

extern "C" size_t cnh_detloss_workspace_bytes(const cnh_detloss_args* a) {
  if (validate(a, false) != CNH_OK) return 0;
  return ws_bytes(a);
}

extern "C" size_t cnh_cand_workspace_bytes(int32_t B) { return B > 0 ? cand_ws_upper(B) : 0; }

extern "C" int cnh_detloss_single_wave(const cnh_detloss_args* a) {
  if (validate(a, false) != CNH_OK) return 0;
  Geo g = make_geo(a, nullptr);
  return plan_stash(pick_stash(!(a->flags & CNH_FLAG_ACCURATE_MATH), use_vec(a, g)), g) ? 1 : 0;
}

static void geo_peers(Geo& g, const cnh_peers* peers) {
  g.world = peers->world;
  g.rank = peers->rank;
  for (int i = 0; i < peers->world; ++i) g.mailbox[i] = static_cast<unsigned long long*>(peers->mailbox[i]);
  g.status = peers->status;
  g.timeout_ns = (long long)(peers->timeout_ms ? peers->timeout_ms : 2000u) * 1000000ll;
}

// an earlier launch on these mailboxes gave up waiting for a peer: refuse to go on (the exchange counters of the
// ranks no longer agree)
static int check_peer_status(const cnh_peers* peers, const char* who) {
  if (peers->status != nullptr) {
    const unsigned code = *reinterpret_cast<volatile const unsigned*>(peers->status);
    CNH_REQUIRE(code == 0u, CNH_E_PEER,
                "%s: an earlier peer exchange on these mailboxes timed out waiting for %s (code %u): a rank died or "
                "skipped a step; re-create the mailboxes on every rank", who,
                code == 1u ? "the normalisers" : "the totals", code);
  }
  return CNH_OK;
}

static int launch_peers_finalize(const cnh_detloss_args* a, const Geo& g, cudaStream_t stream);

static int detloss_fused_impl(const cnh_detloss_args* a, const cnh_peers* peers, void* workspace,
                              size_t workspace_bytes, cnh_stream_t stream) {
  if (int rc = validate(a, true)) return rc;
  CNH_REQUIRE(a->scalars != nullptr, CNH_E_NULL, "detloss_fused: scalars is NULL");
  CNH_REQUIRE(workspace != nullptr && workspace_bytes >= ws_bytes(a), CNH_E_WORKSPACE,
              "detloss_fused: workspace %zu < %zu bytes", workspace_bytes, ws_bytes(a));
  Geo g = make_geo(a, workspace);
  if (peers != nullptr && peers->world > 1) {
    if (int rc = check_peer_status(peers, "detloss_fused_peers")) return rc;
    geo_peers(g, peers);
    g.defer_totals = (a->flags & CNH_FLAG_DEFER_TOTALS) ? 1 : 0;
    CNH_REQUIRE(!g.defer_totals || a->totals != nullptr, CNH_E_NULL, "detloss_fused_peers: CNH_FLAG_DEFER_TOTALS needs a->totals");
    CNH_REQUIRE(a->grad_hm != nullptr, CNH_E_UNSUPPORTED, "detloss_fused_peers: forward-only runs need no exchange before the loss value; use cnh_detloss_fused + an all-reduce of totals");
  }
  const bool fast = !(a->flags & CNH_FLAG_ACCURATE_MATH), vec = use_vec(a, g);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const void* ks = pick_stash(fast, vec);
  if (a->cand != nullptr) a->cand->G = 0;                    // (the single-wave schedule emits no candidates)
  if (a->grad_hm == nullptr) {
    // forward only (torch.no_grad() validation): the single wave when it fits (its CTAs simply stop after the
    // loss terms; measured 17.6 -> see DESIGN at cfg2), else the streaming kernel
    if (!(a->flags & CNH_FLAG_NO_STASH) && plan_stash(ks, g)) {
      static const bool no_coop_f = (getenv("CNH_NO_COOP") != nullptr);
      return launch(ks, !no_coop_f, g.chunk_ctas + 1, g.n_stages, a, g, st, kStashThreads);
    }
    g.n_stages = kStreamStages;
    return launch_stream<M_FWD>(a, g, fast, vec, false, st);
  }
  if (!(a->flags & CNH_FLAG_NO_STASH) && plan_stash(ks, g)) {
    static const bool no_coop = (getenv("CNH_NO_COOP") != nullptr);      // experiment: plain launch of the single wave
    return launch(ks, !no_coop, g.chunk_ctas + 1, g.n_stages, a, g, st, kStashThreads);
  }
  g.n_stages = kStreamStages;
  if (g.world > 1) {
    // sharded pre-count launch: the normalisers are traded after the count phase's grid barrier; the totals always
    // by the one-warp finalize launch (here, unless the caller defers it to overlap it with other work)
    CNH_REQUIRE(a->totals != nullptr, CNH_E_NULL, "detloss_fused_peers: a->totals is NULL");
    const int want_finalize = !g.defer_totals;
    g.defer_totals = 1;
    if (int rc = launch_stream<M_PRECOUNT>(a, g, fast, vec, true, st)) return rc;
    return want_finalize ? launch_peers_finalize(a, g, st) : CNH_OK;
  }
  return launch_stream<M_PRECOUNT>(a, g, fast, vec, true, st);
}

static int launch_peers_finalize(const cnh_detloss_args* a, const Geo& g, cudaStream_t stream) {
  static const bool use_pdl = (getenv("CNH_NO_PDL") == nullptr);
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.gridDim = dim3(1);
  lc.blockDim = dim3(32);
  lc.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = use_pdl ? 1 : 0;
  CNH_CUDA(cudaLaunchKernelEx(&lc, detloss_peers_finalize_kernel, *a, g));
  CNH_CUDA(cudaGetLastError());
  return CNH_OK;
}

extern "C" int cnh_detloss_fused(const cnh_detloss_args* a, void* workspace, size_t workspace_bytes,
                                 cnh_stream_t stream) {
  return detloss_fused_impl(a, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int cnh_detloss_fused_peers(const cnh_detloss_args* a, const cnh_peers* peers, void* workspace,
                                       size_t workspace_bytes, cnh_stream_t stream) {
  CNH_REQUIRE(peers != nullptr, CNH_E_NULL, "detloss_fused_peers: peers is NULL");
  CNH_REQUIRE(peers->world >= 1 && peers->world <= CNH_MAX_PEERS && peers->rank >= 0 && peers->rank < peers->world,
              CNH_E_SHAPE, "detloss_fused_peers: world=%d rank=%d", peers->world, peers->rank);
  for (int i = 0; i < peers->world; ++i)
    CNH_REQUIRE(peers->mailbox[i] != nullptr, CNH_E_NULL, "detloss_fused_peers: mailbox[%d] is NULL", i);
  return detloss_fused_impl(a, peers, workspace, workspace_bytes, stream);
}

extern "C" int cnh_detloss_peers_finalize(const cnh_detloss_args* a, const cnh_peers* peers, void* workspace,
                                          size_t workspace_bytes, cnh_stream_t stream) {
  if (int rc = validate(a, false)) return rc;
  CNH_REQUIRE(peers != nullptr && peers->world > 1 && peers->world <= CNH_MAX_PEERS && peers->rank >= 0 &&
              peers->rank < peers->world, CNH_E_SHAPE, "detloss_peers_finalize: bad peers");
  CNH_REQUIRE(a->totals != nullptr, CNH_E_NULL, "detloss_peers_finalize: a->totals (this rank's totals in, global totals out) is NULL");
  CNH_REQUIRE(workspace != nullptr && workspace_bytes >= ws_bytes(a), CNH_E_WORKSPACE,
              "detloss_peers_finalize: workspace %zu < %zu bytes", workspace_bytes, ws_bytes(a));
  for (int i = 0; i < peers->world; ++i)
    CNH_REQUIRE(peers->mailbox[i] != nullptr, CNH_E_NULL, "detloss_peers_finalize: mailbox[%d] is NULL", i);
  if (int rc = check_peer_status(peers, "detloss_peers_finalize")) return rc;
  Geo g = make_geo(a, workspace);
  geo_peers(g, peers);
  return launch_peers_finalize(a, g, static_cast<cudaStream_t>(stream));
}

extern "C" int cnh_detloss_count(const cnh_detloss_args* a, void* workspace, size_t workspace_bytes,
                                 cnh_stream_t stream) {
  if (int rc = validate(a, false)) return rc;
  CNH_REQUIRE(a->norm_out != nullptr, CNH_E_NULL, "detloss_count: norm_out is NULL");
  CNH_REQUIRE(workspace != nullptr && workspace_bytes >= ws_bytes(a), CNH_E_WORKSPACE,
              "detloss_count: workspace %zu < %zu bytes", workspace_bytes, ws_bytes(a));
  const Geo g = make_geo(a, workspace);
  const void* k = pick_stream<M_COUNT>(true, use_vec(a, g));
  return launch(k, false, stream_grid(k, g, g.n_count, 0), 0, a, g, static_cast<cudaStream_t>(stream), kStashThreads);
}

extern "C" int cnh_detloss_main(const cnh_detloss_args* a, void* workspace, size_t workspace_bytes,
                                cnh_stream_t stream) {
  if (int rc = validate(a, true)) return rc;
  CNH_REQUIRE(a->grad_hm != nullptr, CNH_E_NULL, "detloss_main: grad_hm is NULL (use cnh_detloss_fused for forward only)");
  CNH_REQUIRE(a->norm != nullptr, CNH_E_NULL, "detloss_main: norm is NULL");
  CNH_REQUIRE(a->totals != nullptr, CNH_E_NULL, "detloss_main: totals is NULL");
  CNH_REQUIRE(workspace != nullptr && workspace_bytes >= ws_bytes(a), CNH_E_WORKSPACE,
              "detloss_main: workspace %zu < %zu bytes", workspace_bytes, ws_bytes(a));
  Geo g = make_geo(a, workspace);
  if (a->cand != nullptr) a->cand->G = 0;
  return launch_stream<M_MAIN>(a, g, !(a->flags & CNH_FLAG_ACCURATE_MATH), use_vec(a, g), false, static_cast<cudaStream_t>(stream));
}

extern "C" int cnh_detloss_finalize(const cnh_detloss_args* a, const int64_t* totals, cnh_stream_t stream) {
  CNH_REQUIRE(a != nullptr && totals != nullptr && a->scalars != nullptr, CNH_E_NULL,
              "detloss_finalize: args/totals/scalars is NULL");
  CNH_REQUIRE(a->n_heads >= 0 && a->n_heads <= CNH_MAX_HEADS, CNH_E_SHAPE, "detloss_finalize: n_heads=%d", a->n_heads);
  detloss_finalize_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
      *a, reinterpret_cast<const long long*>(totals));
  CNH_CUDA(cudaGetLastError());
  return CNH_OK;
}

extern "C" int cnh_scale_inplace(const cnh_scale_args* s, cnh_stream_t stream) {
  CNH_REQUIRE(s != nullptr, CNH_E_NULL, "scale_inplace: args is NULL");
  CNH_REQUIRE(s->n_tensors >= 0 && s->n_tensors <= 4, CNH_E_SHAPE, "scale_inplace: n_tensors=%d", s->n_tensors);
  long long most = 0;
  for (int t = 0; t < s->n_tensors; ++t) {
    CNH_REQUIRE(s->data[t] != nullptr && s->count[t] >= 0, CNH_E_NULL, "scale_inplace: tensor %d", t);
    CNH_REQUIRE((reinterpret_cast<uintptr_t>(s->data[t]) & 3u) == 0, CNH_E_ALIGN, "scale_inplace: tensor %d misaligned", t);
    if (s->count[t] > most) most = s->count[t];
  }
  if (s->n_tensors == 0 || most == 0) return CNH_OK;
  long long want = (most / 4 + kThreads - 1) / kThreads;
  const int cap = sm_count() * 2;               // the usual case (factor 1.0) is a no-op: keep the launch small
  const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
  static const bool use_pdl = (getenv("CNH_NO_PDL") == nullptr);
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.gridDim = dim3((unsigned)grid);
  lc.blockDim = dim3(kThreads);
  lc.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = use_pdl ? 1 : 0;
  CNH_CUDA(cudaLaunchKernelEx(&lc, scale_inplace_kernel, *s));
  CNH_CUDA(cudaGetLastError());
  return CNH_OK;
}
