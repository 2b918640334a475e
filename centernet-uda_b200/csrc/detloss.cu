// DetectionLoss forward+backward for sm_100a: sigmoid+clamp, penalty-reduced focal loss,
// masked gather-L1 heads (plain / sigmoid-angle / RAPiD-periodic), all gradients, in ONE
// persistent launch.  Replaces losses/centernet.py:7-95,98-133,192-223 and
// utils/tensor.py:5-25 of the reference (chains of ~100 eager ATen kernels).
//
// HBM-bound streaming work: every heat-map element is read once (logit + target, float4,
// ld.global.cs) and written once (clamped probability + gradient).  The gradient needs
// the batch-wide num_pos, which is only known after everything has been read:
//   * STASH    (small problems): the raw gradient of <= kStash chunks per CTA stays in
//               registers across a grid barrier, then is scaled and stored: 16 B/element,
//               the algorithmic floor.  The barrier's arrival atomic carries the CTA's
//               num_pos, a dedicated CTA computes the scalars while the others store.
//   * PRECOUNT (large problems): phase 0 counts num_pos over the target only (forward
//               order), grid barrier, phase 1 does the full pass in REVERSE order so the
//               tail of the target is still in the 126 MB L2: <= 20 B/element.
//   * COUNT + MAIN: the same two phases as separate launches with the batch-wide
//               normalisers supplied by the caller (the all-reduce of a sharded run sits
//               between them).
//   * FWD: no gradients (validation under no_grad): 12 B/element.
// Regression heads are warp-granular work items (zero-fill of a 16 KB piece of a dense
// gradient plane + the object slots whose centre falls into it), dealt to the CTAs that
// hold the fewest heat-map chunks.
//
// Reductions are EXACT and order-independent: every chunk / item partial (a float produced
// by a fixed-shape tree) is converted to 2^-40 fixed point and added with integer atomics
// into a (hi, lo) pair of 64-bit accumulators.  Integer addition is associative, so the
// totals -- and the loss -- are bit-identical for any grid size, any schedule and any
// sharding of the batch over GPUs (a sharded run all-reduces the 24 integers).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

// L2 eviction-priority hints for the two-pass schedule (evict_last for the pre-counted target,
// evict_first for everything streamed once) measured SLOWER on B200 (98 vs 88 us at the cfg5 shard):
// off by default, kept for experiments.
#ifndef CNH_L2_HINTS
#define CNH_L2_HINTS 0
#endif

namespace cnh {

constexpr int kVec = 4;                       // float4 per thread per chunk
constexpr int kElemsPerThread = kVec * 4;
constexpr int kChunk = kThreads * kElemsPerThread;   // 4096 heat-map elements
constexpr int kStash = 1;                     // chunks a CTA may keep in registers (probability + raw gradient)
constexpr int kPiece = 4096;                  // floats of one regression plane per work item
constexpr int kSlotsPerLane = 8;              // STASH keeps <= 8 object slots per lane (M <= 256)
constexpr int kQ = CNH_TOTALS / 2;            // quantities: focal, num_pos, 3 x (l1, angle, count), spare

enum Mode { M_STASH = 0, M_PRECOUNT = 1, M_MAIN = 2, M_COUNT = 3, M_FWD = 4 };

// Workspace header.  `parity` selects the live accumulator / barrier set.  Ticket modes leave their
// set zeroed (the last CTA cleans up); STASH leaves it dirty, flips parity and cleans the OTHER set,
// which the launch before it used -- so nothing has to be reset on the critical path.
struct WsHeader {
  unsigned parity;
  unsigned done;
  unsigned epoch;                             // launches that used the peer exchange so far
  unsigned pad;
  unsigned long long bar[2][2];               // STASH barriers [parity][0 = chunk CTAs: arrivals << 32 | num_pos, 1 = item CTAs]
  long long acc[2][CNH_TOTALS];               // [q] = hi word, [kQ + q] = lo word
};

// One mailbox slot per (parity, source rank), 32 words:
//   [0]     (tag << 32) | num_pos of the source rank   -- sent first, all a chunk CTA needs
//   [1..24] the source rank's exact totals (counts included)
//   [31]    tag, stored (release.sys) after the totals
constexpr int kSlotWords = 32;
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

struct Geo {
  int HW;
  long long CHW;
  int cps;                 // chunks per sample
  int n_chunks;            // B * cps
  int ppp;                 // pieces per regression plane
  int item0[CNH_MAX_HEADS + 1];   // first item of each head; head h owns B*D_h*ppp items
  int n_items;
  int n_count;             // B * n_heads
  int vec_planes;          // regression planes can be zero-filled with float4 stores
  int stash_slots;         // STASH may keep the slots of an item in registers (M <= 256)
  int chunk_ctas, item_ctas;   // STASH roles
  int world, rank;             // peer exchange (world == 1: none)
  unsigned long long* mailbox[CNH_MAX_PEERS];
  WsHeader* hdr;
  long long* dbg;
};

// ---- exact accumulation ---------------------------------------------------------------------
__device__ __forceinline__ void acc_add_fixed(long long* acc, int q, float v) {
  const long long f = __double2ll_rn((double)v * 1099511627776.0);        // 2^40
  atomicAdd(reinterpret_cast<unsigned long long*>(acc + q), (unsigned long long)(f >> 32));
  atomicAdd(reinterpret_cast<unsigned long long*>(acc + kQ + q), (unsigned long long)(f & 0xffffffffll));
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
  const float a = pos ? q : p;                // the squared factor
  const float L = log_unit<FAST>(arg);
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
  const float L = log_unit<FAST>(q);
  const float omg = __fsub_rn(1.0f, gt);
  float w4 = __fmul_rn(omg, omg);
  w4 = __fmul_rn(w4, w4);
  const float a2 = __fmul_rn(p, p);
  term = __fmul_rn(__fmul_rn(L, a2), w4);
  constexpr float k2 = FAST ? 2.0f * kLn2F : 2.0f;
  const float inner = __fmaf_rn(-__fmul_rn(__fmul_rn(k2, a2), q), L, __fmul_rn(a2, p));
  graw = (p == s) ? __fmul_rn(-w4, inner) : 0.0f;
}

__device__ __forceinline__ void block_reduce2(float& s, int& n, float* red_f, int* red_i) {
  s = warp_sum(s);
  n = warp_sum(n);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) {
    red_f[warp] = s;
    red_i[warp] = n;
  }
  __syncthreads();
  float ts = red_f[0];
  int tn = red_i[0];
#pragma unroll
  for (int w = 1; w < kWarps; ++w) {
    ts += red_f[w];
    tn += red_i[w];
  }
  s = ts;
  n = tn;
}

// One chunk of the heat map.  WRITE_GRAD: scale known, store the gradient now; otherwise (STASH)
// the raw gradient is returned in `graw` and (KEEP_P) the probabilities in `pkeep`, unstored.  Adds the chunk's loss sum to acc; returns its num_pos.
template <bool NEED_GRAD, bool WRITE_GRAD, bool KEEP_P, bool HINT, bool FAST, bool VEC>
__device__ __forceinline__ int focal_chunk(const cnh_detloss_args& a, const Geo& g, long long* acc, int chunk,
                                           float scale, float (&graw)[kElemsPerThread],
                                           float (&pkeep)[kElemsPerThread], float* red_f, int* red_i) {
  const int b = chunk / g.cps, j = chunk - b * g.cps;
  const long long in_sample = (long long)j * kChunk;
  const long long base = (long long)b * g.CHW + in_sample;
  const long long left = g.CHW - in_sample;
  const int n = left < kChunk ? (int)left : kChunk;
  const float* __restrict__ xp = a.hm_logits + base;
  const float* __restrict__ gp = a.hm_gt + base;
  float* __restrict__ pp = a.prob + base;

  float xs[kElemsPerThread], gs[kElemsPerThread];
#pragma unroll
  for (int v = 0; v < kVec; ++v) {
    const int off = v * kThreads * 4 + threadIdx.x * 4;
    if (VEC) {
      float4 x4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = make_float4(2.f, 2.f, 2.f, 2.f);
      if (off < n) {
        if (HINT) {                           // two-pass schedule: everything here is used for the last time
          x4 = ldg_hint(reinterpret_cast<const float4*>(xp + off), l2_policy_evict_first());
          g4 = ldg_hint(reinterpret_cast<const float4*>(gp + off), l2_policy_evict_first());
        } else {
          x4 = ldg_stream(reinterpret_cast<const float4*>(xp + off));
          g4 = ldg_stream(reinterpret_cast<const float4*>(gp + off));
        }
      }
      xs[4 * v + 0] = x4.x; xs[4 * v + 1] = x4.y; xs[4 * v + 2] = x4.z; xs[4 * v + 3] = x4.w;
      gs[4 * v + 0] = g4.x; gs[4 * v + 1] = g4.y; gs[4 * v + 2] = g4.z; gs[4 * v + 3] = g4.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = off + e < n;
        xs[4 * v + e] = ok ? __ldcs(xp + off + e) : 0.f;
        gs[4 * v + e] = ok ? __ldcs(gp + off + e) : 2.f;   // gt = 2: neither pos nor neg
      }
    }
  }
  float sum = 0.f;
  int npos = 0;
  // positives are rare (one pixel per object): a warp whose 512 targets are all < 1 takes the
  // select-free path (same bits, ~30 % fewer instructions)
  bool special = false;
#pragma unroll
  for (int i = 0; i < kElemsPerThread; ++i) special |= !(gs[i] < 1.0f);
  const bool generic = __any_sync(0xffffffffu, special);
#pragma unroll
  for (int v = 0; v < kVec; ++v) {
    const int off = v * kThreads * 4 + threadIdx.x * 4;
    float ps[4], gr[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float term;
      if (generic) focal_elem<FAST>(xs[4 * v + e], gs[4 * v + e], ps[e], term, gr[e], npos);
      else focal_elem_neg<FAST>(xs[4 * v + e], gs[4 * v + e], ps[e], term, gr[e]);
      sum = __fadd_rn(sum, term);
      if (NEED_GRAD) graw[4 * v + e] = WRITE_GRAD ? __fmul_rn(gr[e], scale) : gr[e];
      if (KEEP_P) pkeep[4 * v + e] = ps[e];
    }
    if (KEEP_P) continue;                     // STASH: nothing is stored before the grid barrier
    if (VEC) {
      if (off < n) {
        if (HINT) {
          stg_hint(reinterpret_cast<float4*>(pp + off), make_float4(ps[0], ps[1], ps[2], ps[3]), l2_policy_evict_first());
          if (WRITE_GRAD)
            stg_hint(reinterpret_cast<float4*>(a.grad_hm + base + off),
                     make_float4(graw[4 * v], graw[4 * v + 1], graw[4 * v + 2], graw[4 * v + 3]), l2_policy_evict_first());
        } else {
          *reinterpret_cast<float4*>(pp + off) = make_float4(ps[0], ps[1], ps[2], ps[3]);
          if (WRITE_GRAD)
            stg_stream(reinterpret_cast<float4*>(a.grad_hm + base + off),
                       make_float4(graw[4 * v], graw[4 * v + 1], graw[4 * v + 2], graw[4 * v + 3]));
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (off + e < n) {
          pp[off + e] = ps[e];
          if (WRITE_GRAD) a.grad_hm[base + off + e] = graw[4 * v + e];
        }
    }
  }
  if (FAST) sum = __fmul_rn(sum, kLn2F);      // log2 -> natural log, once per thread
  block_reduce2(sum, npos, red_f, red_i);
  if (threadIdx.x == 0) acc_add_fixed(acc, 0, sum);
  return npos;
}

template <bool VEC>
__device__ __forceinline__ void focal_store_stash(const cnh_detloss_args& a, const Geo& g, int chunk,
                                                  float scale, const float (&graw)[kElemsPerThread],
                                                  const float (&pkeep)[kElemsPerThread]) {
  const int b = chunk / g.cps, j = chunk - b * g.cps;
  const long long in_sample = (long long)j * kChunk;
  const long long base = (long long)b * g.CHW + in_sample;
  const long long left = g.CHW - in_sample;
  const int n = left < kChunk ? (int)left : kChunk;
#pragma unroll
  for (int v = 0; v < kVec; ++v) {
    const int off = v * kThreads * 4 + threadIdx.x * 4;
    if (VEC) {
      if (off < n) {
        *reinterpret_cast<float4*>(a.prob + base + off) =
            make_float4(pkeep[4 * v], pkeep[4 * v + 1], pkeep[4 * v + 2], pkeep[4 * v + 3]);
        stg_stream(reinterpret_cast<float4*>(a.grad_hm + base + off),
                   make_float4(__fmul_rn(graw[4 * v], scale), __fmul_rn(graw[4 * v + 1], scale),
                               __fmul_rn(graw[4 * v + 2], scale), __fmul_rn(graw[4 * v + 3], scale)));
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (off + e < n) {
          a.prob[base + off + e] = pkeep[4 * v + e];
          a.grad_hm[base + off + e] = __fmul_rn(graw[4 * v + e], scale);
        }
    }
  }
}

// num_pos of one chunk from the target only (phase 0 of PRECOUNT / COUNT), per thread.
template <bool VEC, bool HINT>
__device__ __forceinline__ int count_chunk(const cnh_detloss_args& a, const Geo& g, int chunk) {
  const int b = chunk / g.cps, j = chunk - b * g.cps;
  const long long in_sample = (long long)j * kChunk;
  const long long left = g.CHW - in_sample;
  const int n = left < kChunk ? (int)left : kChunk;
  const float* __restrict__ gp = a.hm_gt + (long long)b * g.CHW + in_sample;
  int npos = 0;
#pragma unroll
  for (int v = 0; v < kVec; ++v) {
    const int off = v * kThreads * 4 + threadIdx.x * 4;
    if (VEC) {
      if (off < n) {
        const float4 g4 = HINT ? ldg_hint(reinterpret_cast<const float4*>(gp + off), l2_policy_evict_last())   // stay in L2 for phase 1
                               : __ldg(reinterpret_cast<const float4*>(gp + off));
        npos += (g4.x == 1.f) + (g4.y == 1.f) + (g4.z == 1.f) + (g4.w == 1.f);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (off + e < n) npos += (__ldg(gp + off + e) == 1.f);
    }
  }
  return npos;
}

// ---- masked gather-L1 heads: warp-granular work items -----------------------------------------
__device__ __forceinline__ const cnh_head& head_of(const cnh_detloss_args& a, int h) {
  return h == 0 ? a.heads[0] : (h == 1 ? a.heads[1] : a.heads[2]);   // no local copy of the params
}

struct ItemRef {
  int h, b, d, p0, p1;
};
__device__ __forceinline__ ItemRef decode_item(const cnh_detloss_args& a, const Geo& g, int item) {
  ItemRef r;
  r.h = 0;
#pragma unroll
  for (int h = 1; h < CNH_MAX_HEADS; ++h)
    if (h < a.n_heads && item >= g.item0[h]) r.h = h;
  const int local = item - g.item0[r.h];
  const int piece = local % g.ppp;
  const int plane = local / g.ppp;
  const int D = head_of(a, r.h).D;
  r.d = plane % D;
  r.b = plane / D;
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

// One warp, one item: zero-fill of the piece (ZERO), forward terms of the slots inside it (added to
// acc), and either the scatter itself (SCATTER, inv_denom known) or the slots' (index, coefficient)
// kept in registers for a scatter after the grid barrier (KEEP).  Every lane owns slots lane,
// lane+32, ...  Two dependent memory round trips: {index, mask, target} then {prediction}; the
// zero-fill stores are issued between them so that they never delay a load.
template <bool ZERO, bool FORWARD, bool SCATTER, bool KEEP, bool FAST>
__device__ __forceinline__ void l1_item_warp(const cnh_detloss_args& a, const Geo& g, long long* acc,
                                             const ItemRef& r, float inv_denom, int (&keep_i)[kSlotsPerLane],
                                             float (&keep_g)[kSlotsPerLane]) {
  const cnh_head& hd = head_of(a, r.h);
  const int lane = threadIdx.x & 31;
  const int D = hd.D;
  const bool is_angle = (D == 3 && r.d == 2 && hd.angle_mode != CNH_ANGLE_NONE);
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
        if (SCATTER && gv != 0.f) atomicAdd(gplane + idx[u], gv * inv_denom);   // duplicate centres accumulate
      }
      if (KEEP && k0 == 0) {
        keep_i[u] = (idx[u] >= 0 && gv != 0.f) ? idx[u] : -1;
        keep_g[u] = gv;
      }
    }
  }
  if (FORWARD) {
    l1 = warp_sum(l1);
    ang = warp_sum(ang);
    if (lane == 0) {
      if (l1 != 0.f) acc_add_fixed(acc, 2 + 3 * r.h, l1);
      if (ang != 0.f) acc_add_fixed(acc, 3 + 3 * r.h, ang);
    }
  }
}

// sum(mask_expanded) of one (head, sample): D * sum(mask[b,:]) or sum(mask[b,:,:]); one warp.
__device__ __forceinline__ void count_unit_warp(const cnh_detloss_args& a, long long* acc, int unit) {
  const int h = unit / a.B, b = unit - h * a.B;
  const cnh_head& hd = head_of(a, h);
  const int lane = threadIdx.x & 31;
  const int n = hd.elementwise_mask ? a.M * hd.D : a.M;
  const uint8_t* __restrict__ mp = hd.mask + (long long)b * n;
  int c = 0;
  for (int k = lane; k < n; k += 32) c += __ldg(mp + k);
  c = warp_sum(c);
  if (!hd.elementwise_mask) c *= hd.D;
  if (lane == 0 && c) acc_add_int(acc, 4 + 3 * h, c);
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
      if (hd.D == 3 && hd.angle_mode != CNH_ANGLE_NONE) l += (float)q[3 + 3 * h] / denom * hd.angle_weight;
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
__device__ void finalize_from_acc(const cnh_detloss_args& a, const long long* acc, long long* sh_tot) {
  if (threadIdx.x < CNH_TOTALS) {
    const long long v = __ldcg(acc + threadIdx.x);
    sh_tot[threadIdx.x] = v;
    if (a.totals != nullptr) a.totals[threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0 && a.scalars != nullptr) scalars_from_totals(a, sh_tot, a.scalars);
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

template <int MODE, bool FAST, bool VEC>
__global__ void __launch_bounds__(kThreads, (MODE == M_STASH || MODE == M_PRECOUNT) ? 3 : 4)
detloss_kernel(const cnh_detloss_args a, const Geo g) {
  __shared__ float red_f[kWarps];
  __shared__ int red_i[kWarps];
  __shared__ long long sh_tot[CNH_TOTALS];
  __shared__ unsigned sh_ticket;
  __shared__ unsigned sh_parity;
  __shared__ int sh_norm[1 + CNH_MAX_HEADS];
  const int bid = blockIdx.x, grid = gridDim.x;
  const int warp = threadIdx.x >> 5;
  constexpr bool kGrad = (MODE == M_STASH || MODE == M_PRECOUNT || MODE == M_MAIN);

  dbg_stamp(g.dbg, 0);
  __shared__ unsigned sh_epoch;
  if (threadIdx.x == 0) {
    sh_parity = __ldcg(&g.hdr->parity) & 1u;
    sh_epoch = __ldcg(&g.hdr->epoch);
  }
  __syncthreads();
  const unsigned par = sh_parity;
  const unsigned long long tag = (unsigned long long)sh_epoch + 1ull;   // peer exchange: this launch's tag
  long long* acc = g.hdr->acc[par];
  int keep_i[kSlotsPerLane];
  float keep_g[kSlotsPerLane];

  if (MODE == M_STASH) {
    // Three roles, so that no CTA carries both the gradient stash and the slot registers:
    //   [0, item_ctas)                      regression items, one per warp and round (launched first:
    //                                       their two dependent round trips are the longest chain)
    //   [item_ctas, item_ctas+chunk_ctas)   heat-map chunks (raw gradient kept in registers)
    //   grid-1                              the scalars, while the others store
    // Two barrier words: chunk CTAs only wait for each other (the arrival atomic carries num_pos),
    // item CTAs only for each other (mask counts); the finaliser waits for both.
    unsigned long long* bar_chunk = &g.hdr->bar[par][0];
    unsigned long long* bar_item = &g.hdr->bar[par][1];
    auto arrive = [&](unsigned long long* bar, int payload) {      // call from thread 0 after __syncthreads
      __threadfence();                                             // cumulative: orders the CTA's adds
      atomicAdd(bar, (1ull << 32) | (unsigned long long)(unsigned)payload);
    };
    auto wait_for = [&](unsigned long long* bar, int count) -> unsigned {
      unsigned long long v;
      do { v = ld_acquire_u64(bar); } while ((unsigned)(v >> 32) < (unsigned)count);
      return (unsigned)(v & 0xffffffffull);
    };
    if (bid < g.item_ctas) {
      const int n_other = g.n_items + g.n_count;
      const int first = bid * kWarps + warp;
      const int step = g.item_ctas * kWarps;
      int my_item = -1;
      for (int o = first; o < n_other; o += step) {
        if (o < g.n_items) {
          const ItemRef r = decode_item(a, g, o);
          if (g.stash_slots && o == first) {
            l1_item_warp<false, true, false, true, FAST>(a, g, acc, r, 0.f, keep_i, keep_g);
            my_item = o;
          } else {
            l1_item_warp<false, true, false, false, FAST>(a, g, acc, r, 0.f, keep_i, keep_g);
          }
        } else {
          count_unit_warp(a, acc, o - g.n_items);
        }
      }
      dbg_stamp(g.dbg, 2);
      __syncthreads();
      if (threadIdx.x == 0) {
        arrive(bar_item, 0);
        wait_for(bar_item, g.item_ctas);                   // mask counts are complete
        wait_for(bar_chunk, g.chunk_ctas);                 // the gradient planes are zero-filled
        if (g.world > 1) {                                 // sharded: mask counts of every rank from the mailbox
          const unsigned long long* box = g.mailbox[g.rank] + (size_t)(tag & 1ull) * CNH_MAX_PEERS * kSlotWords;
          long long cnt[CNH_MAX_HEADS] = {0, 0, 0};
          for (int r = 0; r < g.world; ++r) {
            const unsigned long long* slot = box + (size_t)r * kSlotWords;
            while (ld_acquire_sys(slot + 31) != tag) { }
#pragma unroll
            for (int h = 0; h < CNH_MAX_HEADS; ++h) cnt[h] += (long long)__ldcv(slot + 1 + kQ + 4 + 3 * h);
          }
#pragma unroll
          for (int h = 0; h < CNH_MAX_HEADS; ++h) sh_norm[1 + h] = (int)cnt[h];
        } else {
#pragma unroll
          for (int h = 0; h < CNH_MAX_HEADS; ++h) sh_norm[1 + h] = (int)__ldcg(acc + kQ + 4 + 3 * h);
        }
      }
      __syncthreads();
      dbg_stamp(g.dbg, 3);
      for (int o = first; o < g.n_items; o += step) {
        const ItemRef r = decode_item(a, g, o);
        const float inv = 1.f / ((float)sh_norm[1 + r.h] + 1e-4f);
        if (o == my_item) {                                // slots kept in registers: no reload
          const cnh_head& hd = head_of(a, r.h);
          float* __restrict__ gplane = hd.grad + ((long long)r.b * hd.D + r.d) * g.HW;
#pragma unroll
          for (int u = 0; u < kSlotsPerLane; ++u)
            if (keep_i[u] >= 0) atomicAdd(gplane + keep_i[u], keep_g[u] * inv);
        } else {
          l1_item_warp<false, false, true, false, FAST>(a, g, acc, r, inv, keep_i, keep_g);
        }
      }
      dbg_stamp(g.dbg, 6);
    } else if (bid < grid - 1) {
      const int cb = bid - g.item_ctas;
      float stash[kStash][kElemsPerThread], pstash[kStash][kElemsPerThread];
      int cta_npos = 0;
      // the dense regression-gradient planes are zero-filled here, spread over every SM; the item
      // CTAs scatter into them only after this barrier
      for (int o = cb; o < g.n_items; o += g.chunk_ctas) l1_zero_fill_block(a, g, decode_item(a, g, o));
#pragma unroll
      for (int r = 0; r < kStash; ++r) {
        const int chunk = cb + r * g.chunk_ctas;
        if (chunk < g.n_chunks)
          cta_npos += focal_chunk<true, false, true, false, FAST, VEC>(a, g, acc, chunk, 0.f, stash[r], pstash[r], red_f, red_i);
      }
      dbg_stamp(g.dbg, 1);
      if (threadIdx.x == 0) {                              // focal_chunk ends with a block barrier
        arrive(bar_chunk, cta_npos);
        if (g.world > 1) {                                 // sharded: num_pos of every rank, straight from the mailbox
          const unsigned long long* box = g.mailbox[g.rank] + (size_t)(tag & 1ull) * CNH_MAX_PEERS * kSlotWords;
          unsigned total = 0;
          for (int r = 0; r < g.world; ++r) {
            unsigned long long v;
            do { v = ld_acquire_sys(box + (size_t)r * kSlotWords); } while ((v >> 32) != tag);
            total += (unsigned)(v & 0xffffffffull);
          }
          sh_norm[0] = (int)total;
        } else {
          sh_norm[0] = (int)wait_for(bar_chunk, g.chunk_ctas);
        }
      }
      __syncthreads();
      dbg_stamp(g.dbg, 3);
      const int npos = sh_norm[0];
      const float scale = (npos == 0) ? -a.hm_weight : -a.hm_weight / (float)npos;
#pragma unroll
      for (int r = 0; r < kStash; ++r) {
        const int chunk = cb + r * g.chunk_ctas;
        if (chunk < g.n_chunks) focal_store_stash<VEC>(a, g, chunk, scale, stash[r], pstash[r]);
      }
      dbg_stamp(g.dbg, 5);
    } else {
      // the scalars, while the workers store their gradients; then retire the OTHER accumulator set
      const int t = threadIdx.x;
      const unsigned mpar = (unsigned)(tag & 1ull);       // mailbox parity follows the exchange count, not `par`
      if (t == 0) sh_norm[0] = (int)wait_for(bar_chunk, g.chunk_ctas);
      __syncthreads();
      // first, and alone on the critical path: this rank's num_pos inside the tag word of every peer's slot
      if (g.world > 1 && t < g.world)
        st_release_sys(g.mailbox[t] + ((size_t)mpar * CNH_MAX_PEERS + g.rank) * kSlotWords,
                       (tag << 32) | (unsigned long long)(unsigned)sh_norm[0]);
      if (t == 0) {
        wait_for(bar_item, g.item_ctas);
        acc_add_int(acc, 1, (long long)(unsigned)sh_norm[0]);   // num_pos joins the totals
        __threadfence();
      }
      __syncthreads();
      dbg_stamp(g.dbg, 3);
      if (g.world > 1) {
        // ---- the rest of the exchange: exact totals (counts included), then the second tag --------
        if (t < CNH_TOTALS) sh_tot[t] = __ldcg(acc + t);
        __syncthreads();
        if (t < g.world * CNH_TOTALS) {                    // thread = (destination rank, word)
          const int dst = t / CNH_TOTALS, w = t % CNH_TOTALS;
          g.mailbox[dst][((size_t)mpar * CNH_MAX_PEERS + g.rank) * kSlotWords + 1 + w] = (unsigned long long)sh_tot[w];
        }
        __threadfence_system();
        __syncthreads();
        if (t < g.world) st_release_sys(g.mailbox[t] + ((size_t)mpar * CNH_MAX_PEERS + g.rank) * kSlotWords + 31, tag);
        // wait for every source rank's totals in the LOCAL mailbox, then sum (exact integers)
        if (t < g.world) {
          const unsigned long long* slot = g.mailbox[g.rank] + ((size_t)mpar * CNH_MAX_PEERS + t) * kSlotWords;
          while (ld_acquire_sys(slot + 31) != tag) { }
        }
        __syncthreads();
        if (t < CNH_TOTALS) {
          long long sum = 0;
          for (int r = 0; r < g.world; ++r)
            sum += (long long)__ldcv(g.mailbox[g.rank] + ((size_t)mpar * CNH_MAX_PEERS + r) * kSlotWords + 1 + t);
          sh_tot[t] = sum;
          if (a.totals != nullptr) a.totals[t] = sum;
        }
        __syncthreads();
        if (t == 0) {
          if (a.scalars != nullptr) scalars_from_totals(a, sh_tot, a.scalars);
          g.hdr->epoch = (unsigned)tag;
        }
      } else {
        finalize_from_acc(a, acc, sh_tot);
      }
      if (threadIdx.x < CNH_TOTALS) g.hdr->acc[par ^ 1u][threadIdx.x] = 0ll;
      if (threadIdx.x == 0) {
        g.hdr->bar[par ^ 1u][0] = 0ull;
        g.hdr->bar[par ^ 1u][1] = 0ull;
        g.hdr->parity = par ^ 1u;
      }
    }
    dbg_stamp(g.dbg, 7);
    return;
  }

  float scale = 0.f;
  float inv_denom[CNH_MAX_HEADS] = {0.f, 0.f, 0.f};
  if (MODE == M_PRECOUNT || MODE == M_COUNT) {
    // ---- phase 0: normalisers from the targets only ----------------------------------
    int npos = 0;
    for (int chunk = bid; chunk < g.n_chunks; chunk += grid) npos += count_chunk<VEC, (MODE == M_PRECOUNT) && CNH_L2_HINTS>(a, g, chunk);
    npos = block_sum(npos, red_i);
    if (threadIdx.x == 0 && npos) acc_add_int(acc, 1, npos);
    for (int u = (grid - 1 - bid) * kWarps + warp; u < g.n_count; u += grid * kWarps) count_unit_warp(a, acc, u);
  }
  if (MODE == M_PRECOUNT) {
    cg::this_grid().sync();
    const long long npos = __ldcg(acc + kQ + 1);
    scale = (npos == 0) ? -a.hm_weight : -a.hm_weight / (float)npos;
#pragma unroll
    for (int h = 0; h < CNH_MAX_HEADS; ++h) inv_denom[h] = 1.f / ((float)__ldcg(acc + kQ + 4 + 3 * h) + 1e-4f);
  }
  if (MODE == M_MAIN) {
    const float npos = (float)a.norm[0];
    scale = (npos == 0.f) ? -a.hm_weight : -a.hm_weight / npos;
#pragma unroll
    for (int h = 0; h < CNH_MAX_HEADS; ++h) inv_denom[h] = 1.f / ((float)a.norm[1 + h] + 1e-4f);
  }
  if (MODE != M_COUNT) {
    // ---- the streaming pass (reverse chunk order after a pre-count: L2 reuse) ---------
    float unused[kElemsPerThread];
    int npos = 0;
    for (int u = bid; u < g.n_chunks; u += grid) {
      const int chunk = (MODE == M_PRECOUNT) ? g.n_chunks - 1 - u : u;
      npos += focal_chunk<kGrad, kGrad, false, (MODE == M_PRECOUNT) && CNH_L2_HINTS, FAST, VEC>(a, g, acc, chunk, scale, unused, unused, red_f, red_i);
    }
    if (MODE != M_PRECOUNT && threadIdx.x == 0 && npos) acc_add_int(acc, 1, npos);   // PRECOUNT counted in phase 0
    const int n_other = g.n_items + (MODE == M_PRECOUNT ? 0 : g.n_count);
    for (int o = (grid - 1 - bid) * kWarps + warp; o < n_other; o += grid * kWarps) {
      if (o < g.n_items) {
        const ItemRef r = decode_item(a, g, o);
        if (kGrad) {
          const float inv = r.h == 0 ? inv_denom[0] : (r.h == 1 ? inv_denom[1] : inv_denom[2]);
          l1_item_warp<true, true, true, false, FAST>(a, g, acc, r, inv, keep_i, keep_g);
        } else {
          l1_item_warp<false, true, false, false, FAST>(a, g, acc, r, 0.f, keep_i, keep_g);
        }
      } else {
        count_unit_warp(a, acc, o - g.n_items);
      }
    }
  }

  // ---- last CTA: totals, scalars, leave the accumulator set zeroed -------------------------
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    sh_ticket = atomicAdd(&g.hdr->done, 1u);
  }
  __syncthreads();
  if (sh_ticket != (unsigned)(grid - 1)) return;
  __threadfence();
  if (MODE == M_COUNT) {
    if (threadIdx.x == 0) {
      a.norm_out[0] = (double)__ldcg(acc + kQ + 1);
      for (int h = 0; h < CNH_MAX_HEADS; ++h) a.norm_out[1 + h] = (double)__ldcg(acc + kQ + 4 + 3 * h);
    }
  } else {
    finalize_from_acc(a, acc, sh_tot);
  }
  __syncthreads();
  if (threadIdx.x < CNH_TOTALS) acc[threadIdx.x] = 0ll;
  if (threadIdx.x == 0) g.hdr->done = 0;
}

__global__ void __launch_bounds__(32)
detloss_finalize_kernel(const cnh_detloss_args a, const long long* __restrict__ totals) {
  if (threadIdx.x == 0) scalars_from_totals(a, totals, a.scalars);
}

// g *= factor (factor read from device scalars); nothing to do when factor == 1.
__global__ void __launch_bounds__(kThreads)
scale_inplace_kernel(const cnh_scale_args s) {
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
      for (long long k = i; k < n4; k += stride) {
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
    CNH_REQUIRE(hd.angle_mode >= CNH_ANGLE_NONE && hd.angle_mode <= CNH_ANGLE_PERIODIC, CNH_E_UNSUPPORTED,
                "detloss: head %d angle_mode=%d", h, hd.angle_mode);
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
  g.n_count = a->B * a->n_heads;
  g.vec_planes = vec_planes ? 1 : 0;
  g.stash_slots = (a->M <= 32 * kSlotsPerLane) ? 1 : 0;
  g.chunk_ctas = 0;
  g.item_ctas = 0;
  g.world = 1;
  g.rank = 0;
  for (int i = 0; i < CNH_MAX_PEERS; ++i) g.mailbox[i] = nullptr;
  g.hdr = static_cast<WsHeader*>(ws);
  g.dbg = debug_buffer();
  return g;
}

static size_t ws_bytes(const cnh_detloss_args*) { return (sizeof(WsHeader) + 255) / 256 * 256; }

static bool use_vec(const cnh_detloss_args* a, const Geo& g) {
  return (g.CHW % 4 == 0) && aligned16(a->hm_logits) && aligned16(a->hm_gt) && aligned16(a->prob) &&
         (a->grad_hm == nullptr || aligned16(a->grad_hm));
}

template <int MODE>
static const void* pick_kernel(bool fast, bool vec) {
  if (fast) return vec ? (const void*)detloss_kernel<MODE, true, true> : (const void*)detloss_kernel<MODE, true, false>;
  return vec ? (const void*)detloss_kernel<MODE, false, true> : (const void*)detloss_kernel<MODE, false, false>;
}

static int max_resident_ctas(const void* kernel) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 1;
  }
  return per_sm * sm_count();
}

static int launch(const void* kernel, bool cooperative, int grid, const cnh_detloss_args* a, const Geo& g,
                  cudaStream_t stream) {
  void* params[2] = {const_cast<cnh_detloss_args*>(a), const_cast<Geo*>(&g)};
  if (cooperative)
    CNH_CUDA(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(kThreads), params, 0, stream));
  else
    CNH_CUDA(cudaLaunchKernel(kernel, dim3(grid), dim3(kThreads), params, 0, stream));
  return CNH_OK;
}

static int units_to_grid(long long chunks, long long warp_units, int cap) {
  long long want = chunks + (warp_units + kWarps - 1) / kWarps;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

}  // namespace cnh

using namespace cnh;

extern "C" size_t cnh_detloss_workspace_bytes(const cnh_detloss_args* a) {
  if (validate(a, false) != CNH_OK) return 0;
  return ws_bytes(a);
}

static int detloss_fused_impl(const cnh_detloss_args* a, const cnh_peers* peers, void* workspace,
                              size_t workspace_bytes, cnh_stream_t stream) {
  if (int rc = validate(a, true)) return rc;
  CNH_REQUIRE(a->scalars != nullptr, CNH_E_NULL, "detloss_fused: scalars is NULL");
  CNH_REQUIRE(workspace != nullptr && workspace_bytes >= ws_bytes(a), CNH_E_WORKSPACE,
              "detloss_fused: workspace %zu < %zu bytes", workspace_bytes, ws_bytes(a));
  Geo g = make_geo(a, workspace);
  if (peers != nullptr && peers->world > 1) {
    g.world = peers->world;
    g.rank = peers->rank;
    for (int i = 0; i < peers->world; ++i) g.mailbox[i] = static_cast<unsigned long long*>(peers->mailbox[i]);
    CNH_REQUIRE(a->grad_hm != nullptr, CNH_E_UNSUPPORTED, "detloss_fused_peers: forward-only runs need no exchange before the loss value; use cnh_detloss_fused + an all-reduce of totals");
  }
  const bool fast = !(a->flags & CNH_FLAG_ACCURATE_MATH), vec = use_vec(a, g);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->grad_hm == nullptr) {
    const void* k = pick_kernel<M_FWD>(fast, vec);
    return launch(k, false, units_to_grid(g.n_chunks, g.n_items + g.n_count, max_resident_ctas(k)), a, g, st);
  }
  const void* ks = pick_kernel<M_STASH>(fast, vec);
  const int cap = max_resident_ctas(ks);
  {
    // roles: item CTAs (one warp per item and round; at most a quarter of the machine), chunk CTAs
    // (<= kStash chunks each), one finaliser
    const long long warp_units = (long long)g.n_items + g.n_count;
    long long item_ctas = (warp_units + kWarps - 1) / kWarps;
    if (item_ctas > cap / 4) item_ctas = cap / 4;
    if (item_ctas < 1) item_ctas = 1;
    long long chunk_ctas = cap - 1 - item_ctas;
    if (chunk_ctas > g.n_chunks) chunk_ctas = g.n_chunks;
    if (!(a->flags & CNH_FLAG_NO_STASH) && chunk_ctas >= 1 && g.n_chunks <= chunk_ctas * kStash) {
      g.chunk_ctas = (int)chunk_ctas;
      g.item_ctas = (int)item_ctas;
      return launch(ks, true, (int)(chunk_ctas + item_ctas + 1), a, g, st);
    }
  }
  CNH_REQUIRE(g.world == 1, CNH_E_UNSUPPORTED,
              "detloss_fused_peers: problem too large for the register-stash schedule (%d chunks); use count/main", g.n_chunks);
  const void* kp = pick_kernel<M_PRECOUNT>(fast, vec);
  return launch(kp, true, units_to_grid(g.n_chunks, g.n_items + g.n_count, max_resident_ctas(kp)), a, g, st);
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

extern "C" int cnh_detloss_count(const cnh_detloss_args* a, void* workspace, size_t workspace_bytes,
                                 cnh_stream_t stream) {
  if (int rc = validate(a, false)) return rc;
  CNH_REQUIRE(a->norm_out != nullptr, CNH_E_NULL, "detloss_count: norm_out is NULL");
  CNH_REQUIRE(workspace != nullptr && workspace_bytes >= ws_bytes(a), CNH_E_WORKSPACE,
              "detloss_count: workspace %zu < %zu bytes", workspace_bytes, ws_bytes(a));
  const Geo g = make_geo(a, workspace);
  const void* k = pick_kernel<M_COUNT>(true, use_vec(a, g));
  return launch(k, false, units_to_grid(g.n_chunks, g.n_count, max_resident_ctas(k)), a, g,
                static_cast<cudaStream_t>(stream));
}

extern "C" int cnh_detloss_main(const cnh_detloss_args* a, void* workspace, size_t workspace_bytes,
                                cnh_stream_t stream) {
  if (int rc = validate(a, true)) return rc;
  CNH_REQUIRE(a->grad_hm != nullptr, CNH_E_NULL, "detloss_main: grad_hm is NULL (use cnh_detloss_fused for forward only)");
  CNH_REQUIRE(a->norm != nullptr, CNH_E_NULL, "detloss_main: norm is NULL");
  CNH_REQUIRE(a->totals != nullptr, CNH_E_NULL, "detloss_main: totals is NULL");
  CNH_REQUIRE(workspace != nullptr && workspace_bytes >= ws_bytes(a), CNH_E_WORKSPACE,
              "detloss_main: workspace %zu < %zu bytes", workspace_bytes, ws_bytes(a));
  const Geo g = make_geo(a, workspace);
  const void* k = pick_kernel<M_MAIN>(!(a->flags & CNH_FLAG_ACCURATE_MATH), use_vec(a, g));
  return launch(k, false, units_to_grid(g.n_chunks, g.n_items + g.n_count, max_resident_ctas(k)), a, g,
                static_cast<cudaStream_t>(stream));
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
  const int cap = sm_count() * 8;
  const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
  scale_inplace_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(*s);
  CNH_CUDA(cudaGetLastError());
  return CNH_OK;
}
