// Detection decode for sm_100a in ONE persistent launch (replaces backends/decode.py:6-76:
// max_pool2d + 4 element-wise passes + two torch.topk sorts + 3 full-map transpose copies).
//
// Work unit = one tile of 16 rows x <=128 columns of one class plane.  Every CTA owns a contiguous
// range of tiles (sample-major) and walks it with a 2-stage TMA ring: while tile t is processed,
// tile t+1 (+ halo) is already being staged in shared memory by cp.async.bulk.tensor.3d (SASS
// UTMALDG) from a [B*C, H, W] tensor map; out-of-range rows/columns are zero-filled by the TMA
// unit, which equals max-pool's -inf padding because heat >= 0.  (W % 4 != 0 or a misaligned base
// falls back to guarded loads.)
//
// Per tile:
//   * threshold-first scan: a row of the tile is only examined further if some pixel reaches the
//     sample's current pruning threshold (one LDS.128 + 4 compares + a ballot per warp-row), so
//     once the threshold has tightened a tile costs little more than its TMA transfer;
//   * surviving pixels get the 3x3 peak test from shared memory; peaks become 64-bit keys
//     (score_bits << 32) | ~flat_index -- descending key order == score descending, ties to the
//     LOWER flat index c*HW + y*W + x -- and are counted in a LOCAL 4096-bin histogram;
//   * a tile with more than K peaks keeps only the bins >= that of its K-th score (block suffix
//     scan; exact MSB radix select only when a bin is overfull, i.e. massive ties), adds its
//     counts to the sample's GLOBAL 1024-bin histogram and republishes the sample's threshold;
//   * kept keys are staged in shared memory and flushed once per (CTA, sample): one reservation
//     atomic for a dense slice of the sample's candidate list, one ticket atomic.
// The CTA whose ticket completes a sample merges it: final threshold from the global histogram,
// survivors into shared memory, histogram selection, rank sort of the <= K+few keys, zero-score
// filler when the sample has fewer than K peaks (ascending flat index, as a stable sort would),
// gather of reg / wh / angle / keypoints straight from NCHW, box assembly.
// No full sort, no transposes: heat is read from HBM exactly once (4*C*H*W bytes/sample).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <cooperative_groups.h>

#include "common.cuh"
#include "cand.cuh"

namespace cg = cooperative_groups;

namespace cnh {

// Barrier over the first 256 threads of the CTA (warps 0-7).  The selection helpers below are written for
// 256 threads; in the 256-thread kernels this is a full-CTA barrier, in the 1024-thread cluster kernel the
// other warps wait at the next __syncthreads().
__device__ __forceinline__ void group_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

constexpr int kCols = 128;                // max tile columns
constexpr int kPadL = 4;                  // left halo padded to 4 floats: interior is 16B aligned
constexpr int kBoxWMax = kCols + 2 * kPadL;
// tile height: 16 rows (deep TMA ring, many tiles per CTA) or 32 rows (one tile per CTA, small problems)
__host__ __device__ constexpr int tile_floats(int rows) { return ((rows + 2) * kBoxWMax + 31) / 32 * 32; }   // 128-byte multiples
constexpr int kMergeKeyCap = 4096;        // survivors the merge kernel can hold in shared memory
constexpr int kMaxK = 1024;
constexpr int kCoarseBins = 1024;         // per-sample global histogram: fine bin >> 2
constexpr int kSlack = 64;                // a tile forwards at most K + kSlack keys
constexpr int kStageCap = kMaxK + kSlack; // staging buffer (keys) per CTA

struct SampleState {                      // zero between launches
  unsigned cand_cnt;
  unsigned thr_bits;
  unsigned pad[2];
};

struct DecGeo {
  int HW, rows, tiles_x, tiles_y, tiles_per_plane, tiles_per_sample, box_w, use_tma, slot, n_stages;
  long long n_tiles;
  SampleState* state;                     // [B]                        zero between launches
  unsigned* ghist;                        // [B][kCoarseBins]           zero between launches
  u64* cand;                              // [B][tiles_per_sample * slot]  dense per-sample lists
  long long* dbg;
  // candidate lists (cand.cuh): filled by decode_stream_kernel or by the detection-loss kernels, consumed by
  // finish_sample (inside decode_cluster_kernel, only_overflow)
  CandGeo cl;
  int only_overflow;                      // cluster kernel launched as the fallback: samples without the flag exit at once
  int verify_rows;                        // finish: candidates of the first / last row of every `verify_rows`-row tile were
                                          // tested without the row beyond the tile (0: every candidate is a verified peak)
};

// leave a sample's candidate-list state (histograms, slice sizes) zeroed for the next launch; first 256 threads
__device__ __forceinline__ void cand_state_clear(const DecGeo& g, int b) {
  const int tid = threadIdx.x;
  if (tid >= 256) return;
  unsigned* const shist = g.cl.shist + (long long)b * kSuperBins;
  unsigned* const fhist = g.cl.fhist + (long long)b * kFineBins;
  unsigned* const cta_cnt = g.cl.cta_cnt + (long long)b * g.cl.G;
  for (int q = tid; q < kFineBins; q += 256) fhist[q] = 0u;
  if (tid < kSuperBins) shist[tid] = 0u;
  for (int q = tid; q < g.cl.G; q += 256) cta_cnt[q] = 0u;
}

constexpr int kMaxStages = 8;
template <int KEYS>
struct __align__(128) DecSmemT {            // followed in dynamic shared memory by the TMA ring: n_stages tiles
  static constexpr int kKeyCap = KEYS;    // worst case: every pixel of the tile is a peak
  u64 keys[KEYS];
  u64 stage[kStageCap];                   // merge: `sorted`
  unsigned hist[kFineBins / 2];           // 16-bit counters packed in pairs; radix select uses [0,256); merge: `sel`
  u64 mbar[kMaxStages];
  u64 sh_prefix;
  unsigned cnt;
  unsigned cnt2;
  unsigned sh_need;
  unsigned sh_flag;
  unsigned sh_thr;
  unsigned sh_bin;
  unsigned sh_above;
  unsigned sh_inbin;
  unsigned sh_base;
  unsigned warp_tot[kWarps];
};
static_assert(sizeof(unsigned) * (kFineBins / 2) >= sizeof(u64) * kMaxK, "merge `sel` aliases the histogram");

// ---- block-wide helpers ---------------------------------------------------------------------
// Exclusive SUFFIX sum over threads (sum of v of all threads with a higher tid) and the total.
__device__ __forceinline__ unsigned block_suffix_excl(unsigned v, unsigned* warp_tot, unsigned& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_down_sync(0xffffffffu, incl, o);
    if (lane + o < 32) incl += t;
  }
  group_sync();                        // warp_tot may still be read from a previous use
  if (lane == 0) warp_tot[warp] = incl;
  group_sync();
  unsigned higher = 0;
  total = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    if (w > warp) higher += warp_tot[w];
    total += warp_tot[w];
  }
  return higher + (incl - v);
}

__device__ __forceinline__ void hist_add(unsigned* hist, int bin) {
  atomicAdd(&hist[bin >> 1], (bin & 1) ? 0x10000u : 1u);
}

// Highest fine bin t with count(bins >= t) >= need, from the packed 16-bit histogram
// (thread i owns bins [16i, 16i+16)).  Results in s.sh_bin / s.sh_above / s.sh_inbin;
// needs total >= need.  Ends with a barrier.
template <class SM>
__device__ __noinline__ void find_kth_bin(SM& s, unsigned need) {
  const int tid = threadIdx.x;
  unsigned c[16], v = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const unsigned w = s.hist[8 * tid + j];
    c[2 * j] = w & 0xffffu;
    c[2 * j + 1] = w >> 16;
    v += c[2 * j] + c[2 * j + 1];
  }
  unsigned total;
  const unsigned excl = block_suffix_excl(v, s.warp_tot, total);
  if (excl < need && excl + v >= need) {
    unsigned acc = excl;
#pragma unroll
    for (int j = 15; j >= 0; --j) {
      if (acc + c[j] >= need) {
        s.sh_bin = 16 * tid + j;
        s.sh_above = acc;
        s.sh_inbin = c[j];
        break;
      }
      acc += c[j];
    }
  }
  group_sync();
}

// Threshold from the sample's global coarse histogram: score bits of the lower edge of the highest
// coarse bin t with count(bins >= t) >= K, or 0 if fewer than K peaks are known.  Thread i owns
// coarse bins [4i, 4i+4).  Result in s.sh_thr (also returned); ends with a barrier.
template <class SM>
__device__ __noinline__ unsigned global_threshold(const unsigned* ghist, unsigned K, SM& s) {
  const int tid = threadIdx.x;
  const uint4 g4 = __ldcg(reinterpret_cast<const uint4*>(ghist) + tid);
  const unsigned c[4] = {g4.x, g4.y, g4.z, g4.w};
  const unsigned v = c[0] + c[1] + c[2] + c[3];
  if (tid == 0) s.sh_thr = 0u;
  unsigned total;
  const unsigned excl = block_suffix_excl(v, s.warp_tot, total);
  if (excl < K && excl + v >= K) {
    unsigned acc = excl;
#pragma unroll
    for (int j = 3; j >= 0; --j) {
      if (acc + c[j] >= K) {
        s.sh_thr = __float_as_uint((float)(4 * tid + j) * (1.0f / (float)kCoarseBins));
        break;
      }
      acc += c[j];
    }
  }
  group_sync();
  return s.sh_thr;
}

// ---- block-level radix select over 64-bit keys (exact; fallback for overfull bins) ---------------
// for_each(f) must call f(key) for every key, each thread visiting a disjoint subset.
// Returns T such that exactly `need` keys are >= T (keys are unique; #keys > need >= 1).
// MSB-first, 8-bit digits, early exit as soon as the remaining bin is taken whole.
template <class ForEach, class SM>
__device__ u64 radix_select_kth(ForEach for_each, int need, SM& s) {
  u64 prefix = 0, mask = 0;
  unsigned remaining = (unsigned)need;
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += kThreads) s.hist[i] = 0;
    group_sync();
    for_each([&](u64 k) {
      if ((k & mask) == prefix) atomicAdd(&s.hist[(unsigned)(k >> shift) & 255u], 1u);
    });
    group_sync();
    if (threadIdx.x < 32) {
      // lane owns bins [8*lane, 8*lane+8); find the highest digit d with count(>= d) >= remaining
      const int lane = threadIdx.x;
      unsigned c[8], tot = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j] = s.hist[8 * lane + j]; tot += c[j]; }
      unsigned incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += v;
      }
      const unsigned above = incl - tot;
      if (above < remaining && incl >= remaining) {
        unsigned acc = above;
        int d = 8 * lane + 7;
#pragma unroll
        for (int j = 7; j >= 0; --j) {
          if (acc + c[j] >= remaining) { d = 8 * lane + j; break; }
          acc += c[j];
        }
        s.sh_prefix = prefix | ((u64)(unsigned)d << shift);
        s.sh_need = remaining - acc;                            // still needed inside bin d
        s.sh_flag = (s.hist[d] == remaining - acc) ? 1u : 0u;   // whole bin taken: done
      }
    }
    group_sync();
    prefix = s.sh_prefix;
    remaining = s.sh_need;
    mask |= (u64)255u << shift;
    const bool done = s.sh_flag != 0u;
    group_sync();
    if (done) break;
  }
  return prefix;   // lower digits zero: every key of the last bin is >= prefix
}

// Warp-aggregated append of `key` (if keep) to dst[*counter ...]; counter lives in shared memory.
__device__ __forceinline__ void append_if(bool keep, u64 key, u64* dst, unsigned* counter, unsigned cap) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  if (bal == 0u) return;
  unsigned base = 0;
  if (lane == (unsigned)(__ffs(bal) - 1)) base = atomicAdd(counter, (unsigned)__popc(bal));
  base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
  const unsigned pos = base + __popc(bal & ((1u << lane) - 1u));
  if (keep && pos < cap) dst[pos] = key;
}

// is flat position a positive-score peak?  (global-memory version for the filler path)
__device__ __noinline__ bool is_candidate_global(const cnh_decode_args& a, int b, long long flat, int HW) {
  const int c = (int)(flat / HW), pix = (int)(flat - (long long)c * HW);
  const int y = pix / a.W, x = pix - y * a.W;
  const float* plane = a.heat + ((long long)b * a.C + c) * HW;
  auto val = [&](int yy, int xx) -> float {
    if (yy < 0 || yy >= a.H || xx < 0 || xx >= a.W) return 0.f;
    float v = plane[yy * a.W + xx];
    if (a.apply_sigmoid) v = clamp_prob(1.0f / (1.0f + expf(-v)));
    return v;
  };
  const float v = val(y, x);
  if (!(v > 0.f)) return false;
  float m = v;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) m = fmaxf(m, val(y + dy, x + dx));
  return m == v;
}

// ---- threshold-first scan of one staged tile (rows x cols, halo included in `tile`) -------------
// warp w of kNWarps owns rows [w*kRows/kNWarps, (w+1)*kRows/kNWarps), lane owns columns [4*lane, 4*lane+4).  Peaks with score >= thr
// are appended to keys (counter *key_cnt, shared memory) and counted in the packed fine histogram s.hist.
// plane_flat0 = c*H*W; (y0, x0) = tile origin in the plane.
template <int kRows, int kNWarps, class SM>
__device__ __forceinline__ void scan_tile(SM& s, u64* keys, unsigned* key_cnt, const float* tile, int BW, int rows,
                                          int cols, unsigned thr, unsigned plane_flat0, int y0, int x0, int W) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float thr_f = __uint_as_float(thr);
  constexpr int kRowsPerWarp = kRows / kNWarps;
  float cv[kRowsPerWarp][4];
  unsigned flags = 0;                     // bit 4*rr + e
#pragma unroll
  for (int rr = 0; rr < kRowsPerWarp; ++rr) {
    const int r = warp * kRowsPerWarp + rr;
    const float* row = tile + (r + 1) * BW + kPadL;
    const float4 v = *reinterpret_cast<const float4*>(row + 4 * lane);
    cv[rr][0] = v.x; cv[rr][1] = v.y; cv[rr][2] = v.z; cv[rr][3] = v.w;
    const bool row_ok = r < rows;
    unsigned pass = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e)
      pass |= (row_ok && (4 * lane + e < cols) && cv[rr][e] > 0.f && cv[rr][e] >= thr_f) ? (1u << e) : 0u;
    const unsigned hit = __ballot_sync(0xffffffffu, pass != 0u);
    if (hit == 0u) continue;                                       // nothing in this row can matter
    if (__popc(hit) <= 6) {
      // sparse row (the steady state once the threshold has tightened): each lane tests its own
      // few pixels against their 8 neighbours straight from shared memory
      if (pass) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (pass & (1u << e)) {
            const float* p = row + 4 * lane + e;
            const float m = fmaxf(fmaxf(fmaxf(p[-BW - 1], p[-BW]), fmaxf(p[-BW + 1], p[-1])),
                                  fmaxf(fmaxf(p[1], p[BW - 1]), fmaxf(p[BW], p[BW + 1])));
            if (cv[rr][e] >= m) flags |= 1u << (4 * rr + e);
          }
      }
      continue;
    }
    // dense row: 3x3 maximum for the whole row, rows r-1, r, r+1 (tile rows r, r+1, r+2),
    // neighbours by shuffle
    float m[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int dr = 0; dr < 3; ++dr) {
      const float* rw = tile + (r + dr) * BW + kPadL;
      const float4 u = (dr == 1) ? v : *reinterpret_cast<const float4*>(rw + 4 * lane);
      float left = __shfl_up_sync(0xffffffffu, u.w, 1);
      float right = __shfl_down_sync(0xffffffffu, u.x, 1);
      if (lane == 0) left = rw[-1];
      if (lane == 31) right = rw[4 * 32];
      m[0] = fmaxf(m[0], fmaxf(fmaxf(left, u.x), u.y));
      m[1] = fmaxf(m[1], fmaxf(fmaxf(u.x, u.y), u.z));
      m[2] = fmaxf(m[2], fmaxf(fmaxf(u.y, u.z), u.w));
      m[3] = fmaxf(m[3], fmaxf(fmaxf(u.z, u.w), right));
    }
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if ((pass & (1u << e)) && cv[rr][e] == m[e]) flags |= 1u << (4 * rr + e);
  }
  // one warp-aggregated append for the warp's rows
  const int mine = __popc(flags);
  if (__ballot_sync(0xffffffffu, mine != 0) != 0u) {
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned base = 0;
    if (lane == 31) base = atomicAdd(key_cnt, (unsigned)total);
    base = __shfl_sync(0xffffffffu, base, 31);
    unsigned pos = base + (unsigned)(incl - mine);
#pragma unroll
    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
      const unsigned flat0 = plane_flat0 +
                             (unsigned)(y0 + warp * kRowsPerWarp + rr) * (unsigned)W + (unsigned)(x0 + 4 * lane);
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (flags & (1u << (4 * rr + e))) {
          const unsigned bits = __float_as_uint(cv[rr][e]);
          keys[pos++] = ((u64)bits << 32) | (u64)(0xffffffffu - (flat0 + e));
          hist_add(s.hist, fine_bin(bits));
        }
    }
  }
}

template <class SM>
__device__ __forceinline__ void emit_sorted(const cnh_decode_args& a, const DecGeo& g, SM& s, int b, int got,
                                            const float* payload, int n_payload);

// ---- final selection, sort, filler, gather and box assembly of one sample -------------------------
// keys[0..m): candidate keys in shared memory, a superset of the sample's top-K, with s.hist holding their
// packed fine histogram (only read when m > kThreads); if m > key_cap the list in `keys` is partial and
// overflow_keys(f) must call f(key) for every candidate (each thread a disjoint subset, all threads
// participating).  Uses s.stage as `sorted` and aliases s.hist as `sel`; s.cnt2 must be 0.
template <class SM, class Overflow>
__device__ __forceinline__ void select_sort_emit(const cnh_decode_args& a, const DecGeo& g, SM& s, int b,
                                                 const u64* keys, int m, int key_cap, Overflow overflow_keys) {
  const int tid = threadIdx.x;
  const int K = a.K;
  u64* const sel = reinterpret_cast<u64*>(s.hist);
  u64* const sorted = s.stage;
  int got = 0;                               // keys to sort; the first min(got, K) ranks are real detections
  const u64* sort_src = sel;
  if (m <= kThreads || m <= K) {
    sort_src = keys;                         // one key per thread (bitonic) or nothing to cut: sort the survivors directly
    got = m;
  } else if (m <= key_cap) {
    find_kth_bin(s, (unsigned)K);            // ends with a barrier: the histogram is dead afterwards
    const unsigned keep_n = s.sh_above + s.sh_inbin;
    if (keep_n <= (unsigned)kMaxK) {
      const int tbin = (int)s.sh_bin;
      for (int i0 = 0; i0 < m; i0 += kThreads) {
        const int i = i0 + tid;
        const u64 k = (i < m) ? keys[i] : 0ull;
        append_if((i < m) && fine_bin((unsigned)(k >> 32)) >= tbin, k, sel, &s.cnt2, (unsigned)kMaxK);
      }
      got = (int)keep_n;
    } else {
      const u64 T = radix_select_kth([&](auto f) { for (int i = tid; i < m; i += kThreads) f(keys[i]); }, K, s);
      for (int i0 = 0; i0 < m; i0 += kThreads) {
        const int i = i0 + tid;
        const u64 k = (i < m) ? keys[i] : 0ull;
        append_if((i < m) && k >= T, k, sel, &s.cnt2, (unsigned)kMaxK);
      }
      got = K;
    }
  } else {
    // more survivors than shared memory holds (heavy ties): exact radix select straight from the
    // sample's candidate list in global memory
    const u64 T = radix_select_kth(overflow_keys, K, s);
    overflow_keys([&](u64 k) { append_if(k >= T, k, sel, &s.cnt2, (unsigned)kMaxK); });
    got = K;
  }
  group_sync();
  dbg_stamp(g.dbg, 8);
  const float* payload = nullptr;
  int n_payload = 0;
  if (got <= kThreads) {
    // <= 256 keys: one key per thread, rank sort (keys are unique: position = number of larger keys; ~got broadcast
    // LDS.64 per thread, no barrier inside).  The key's reg / wh values are loaded BEFORE the scan (their round trip
    // hides under it) and land at the key's rank in `payload` -- the upper half of the key buffer, free by now.
    float* const pay = reinterpret_cast<float*>(const_cast<u64*>(keys) + key_cap / 2);
    const bool act = tid < got;
    const u64 k = act ? sort_src[tid] : 0ull;
    float pv[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (act) {
      const unsigned pix = (0xffffffffu - (unsigned)(k & 0xffffffffu)) % (unsigned)g.HW;
      if (a.reg) {
        pv[0] = __ldg(a.reg + ((long long)b * 2 + 0) * g.HW + pix);
        pv[1] = __ldg(a.reg + ((long long)b * 2 + 1) * g.HW + pix);
      }
      pv[2] = __ldg(a.wh + ((long long)b * a.D + 0) * g.HW + pix);
      pv[3] = __ldg(a.wh + ((long long)b * a.D + 1) * g.HW + pix);
      if (a.rotated) pv[4] = __ldg(a.wh + ((long long)b * a.D + 2) * g.HW + pix);
    }
    int rank = 0;
    if (act) {
      // (kept compact on purpose: this code runs once per launch, straight from a cold instruction cache -- the
      // fully unrolled form of this loop, 4.6 KB of it, took twice as long)
      int j = 0, r0 = 0, r1 = 0, r2 = 0, r3 = 0;
#pragma unroll 1
      for (; j + 3 < got; j += 4) {                         // four independent loads in flight
        const u64 k0 = sort_src[j], k1 = sort_src[j + 1], k2 = sort_src[j + 2], k3 = sort_src[j + 3];
        r0 += (k0 > k);
        r1 += (k1 > k);
        r2 += (k2 > k);
        r3 += (k3 > k);
      }
      rank = r0 + r1 + r2 + r3;
#pragma unroll 1
      for (; j < got; ++j) rank += (sort_src[j] > k);
    }
    group_sync();                                           // sort_src may be `sorted`'s neighbour: settle the reads
    if (act) {
      sorted[rank] = k;
#pragma unroll
      for (int c = 0; c < 5; ++c) pay[5 * rank + c] = pv[c];
    }
    payload = pay;
    n_payload = got;
  } else {
  // rank sort (keys are unique): position = number of larger keys.  T lanes share a key when there
  // are fewer keys than threads (T = 8, 4, 2 or 1), each scanning every T-th key.
  {
    int T = 1;
    while (T < 8 && got * T * 2 <= kThreads) T *= 2;
    const int per_pass = kThreads / T;
    for (int i0 = 0; i0 < got; i0 += per_pass) {
      const int i = i0 + tid / T, part = tid % T;
      const bool active = i < got;
      const u64 k = active ? sort_src[i] : 0ull;
      int rank = 0;
      if (active) {
        int j = part;
        for (; j + 3 * T < got; j += 4 * T) {                   // 4 independent loads in flight
          const u64 k0 = sort_src[j], k1 = sort_src[j + T], k2 = sort_src[j + 2 * T], k3 = sort_src[j + 3 * T];
          rank += (k0 > k) + (k1 > k) + (k2 > k) + (k3 > k);
        }
        for (; j < got; j += T) rank += (sort_src[j] > k);
      }
      for (int o = 1; o < T; o <<= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
      if (active && part == 0 && rank < kMaxK) sorted[rank] = k;
    }
  }
  }
  dbg_stamp(g.dbg, 11);
  if (g.dbg && tid == 0) { g.dbg[(long long)blockIdx.x * 16 + 12] = m; g.dbg[(long long)blockIdx.x * 16 + 13] = got; }
  group_sync();
  emit_sorted(a, g, s, b, got, payload, n_payload);
}

// ---- the tail of select_sort_emit: s.stage[0..got) sorted descending -> filler, counts, gather, boxes -------------
// payload (nullable, shared memory): for rank r < n_payload the gathered values {reg x, reg y, w, h, angle} at
// payload[5*r ..], loaded while the keys were being sorted (the cluster leader does); other ranks gather here.
template <class SM>
__device__ __forceinline__ void emit_sorted(const cnh_decode_args& a, const DecGeo& g, SM& s, int b, int got,
                                            const float* payload, int n_payload) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = a.K;
  u64* const sorted = s.stage;
  if (got > K) got = K;
  // fewer than K peaks: zero-score filler at the lowest flat indices that are not candidates
  if (got < K) {
    const long long total = (long long)a.C * g.HW;
    int have = got;
    for (long long f0 = 0; f0 < total && have < K; f0 += kThreads) {
      const long long f = f0 + tid;
      const bool fill = (f < total) && !is_candidate_global(a, b, f, g.HW);
      const unsigned bal = __ballot_sync(0xffffffffu, fill);
      if (lane == 0) s.warp_tot[warp] = (unsigned)__popc(bal);
      group_sync();
      unsigned before = 0, all = 0;
      for (int w = 0; w < kWarps; ++w) { if (w < warp) before += s.warp_tot[w]; all += s.warp_tot[w]; }
      const unsigned pos = (unsigned)have + before + (unsigned)__popc(bal & ((1u << lane) - 1u));
      if (fill && pos < (unsigned)K) sorted[pos] = (u64)(0xffffffffu - (unsigned)f);   // score bits 0
      have += (int)all;
      group_sync();
    }
  }
  group_sync();
  dbg_stamp(g.dbg, 9);
  // ---- gather + box assembly (backends/decode.py:44-74) -----------------------------------------
  const int ncol = a.rotated ? 7 : 6;
  if (a.counts_out) {                        // rows with score >= threshold: a prefix of the sorted list
    if (tid == 0) s.cnt2 = 0;
    group_sync();
    int mine = 0;
    for (int r = tid; r < K; r += kThreads) mine += (__uint_as_float((unsigned)(sorted[r] >> 32)) >= a.score_threshold) ? 1 : 0;
    mine = warp_sum(mine);
    if (lane == 0 && mine) atomicAdd(&s.cnt2, (unsigned)mine);
    group_sync();
    if (tid == 0) a.counts_out[b] = (int)s.cnt2;
  }
  for (int r = tid; r < K; r += kThreads) {
    const u64 k = sorted[r];
    const float score = __uint_as_float((unsigned)(k >> 32));
    const unsigned flat = 0xffffffffu - (unsigned)(k & 0xffffffffu);
    const int cls = (int)(flat / (unsigned)g.HW);
    const int pix = (int)(flat - (unsigned)cls * (unsigned)g.HW);
    const int yy = pix / a.W, xx = pix - yy * a.W;
    float xs = (float)xx, ys = (float)yy;
    const bool pre = r < n_payload;           // gathered while sorting
    if (a.reg) {
      xs += pre ? payload[5 * r + 0] : a.reg[((long long)b * 2 + 0) * g.HW + pix];
      ys += pre ? payload[5 * r + 1] : a.reg[((long long)b * 2 + 1) * g.HW + pix];
    } else {
      xs += 0.5f;
      ys += 0.5f;
    }
    const float w = pre ? payload[5 * r + 2] : a.wh[((long long)b * a.D + 0) * g.HW + pix];
    const float h = pre ? payload[5 * r + 3] : a.wh[((long long)b * a.D + 1) * g.HW + pix];
    float* out = a.dets + ((long long)b * K + r) * ncol;
    const float sc = a.box_scale;
    if (!a.rotated) {
      float x1 = xs - w / 2, y1 = ys - h / 2, x2 = xs + w / 2, y2 = ys + h / 2;
      if (sc != 1.0f) { x1 *= sc; y1 *= sc; x2 *= sc; y2 *= sc; }
      out[0] = x1; out[1] = y1; out[2] = x2; out[3] = y2; out[4] = score; out[5] = (float)cls;
    } else {
      const float av = pre ? payload[5 * r + 4] : a.wh[((long long)b * a.D + 2) * g.HW + pix];
      const float ang = clamp_prob(1.0f / (1.0f + expf(-av))) * 360.0f - 180.0f;
      float bx = xs, by = ys, bw = w, bh = h;
      if (sc != 1.0f) { bx *= sc; by *= sc; bw *= sc; bh *= sc; }
      out[0] = bx; out[1] = by; out[2] = bw; out[3] = bh; out[4] = ang; out[5] = score; out[6] = (float)cls;
    }
    if (a.inds_out) a.inds_out[(long long)b * K + r] = (long long)flat;
    if (a.kps && a.kps_out) {
      float* ko = a.kps_out + ((long long)b * K + r) * a.nk * 2;
      for (int j = 0; j < a.nk; ++j) {
        float kx = a.kps[((long long)b * 2 * a.nk + 2 * j) * g.HW + pix] + xs;
        float ky = a.kps[((long long)b * 2 * a.nk + 2 * j + 1) * g.HW + pix] + ys;
        if (sc != 1.0f) { kx *= sc; ky *= sc; }
        ko[2 * j] = kx;
        ko[2 * j + 1] = ky;
      }
    }
  }
}

// ---- stage 2: merge one sample (run by the CTA whose ticket completed it) ---------------------
typedef DecSmemT<kMergeKeyCap> MergeSmem;
__device__ void merge_sample(const cnh_decode_args& a, const DecGeo& g, MergeSmem& s, int b) {
  constexpr int kKeyCap = MergeSmem::kKeyCap;
  const int tid = threadIdx.x;
  const int K = a.K;
  unsigned* ghist = g.ghist + (long long)b * kCoarseBins;
  const u64* cand = g.cand + (long long)b * g.tiles_per_sample * g.slot;
  dbg_stamp(g.dbg, 5);
  for (int j = 0; j < kFineBins / 2 / kThreads; ++j) s.hist[tid + j * kThreads] = 0u;
  if (tid == 0) { s.cnt = 0; s.cnt2 = 0; }
  asm volatile("griddepcontrol.wait;" ::: "memory");     // PDL: the tile kernel's writes are visible from here
  const unsigned nc = __ldcg(&g.state[b].cand_cnt);                         // same round trip as the histogram
  const unsigned thr_final = global_threshold(ghist, (unsigned)K, s);        // barriers inside
  dbg_stamp(g.dbg, 6);
  // survivors (score >= final threshold) -> shared memory keys + fine histogram; dense list, 8
  // independent loads in flight per thread
  auto for_each_survivor = [&](auto f) {
    constexpr int kB = 8;
    for (unsigned e0 = 0; e0 < nc; e0 += kB * kThreads) {
      u64 k[kB];
#pragma unroll
      for (int j = 0; j < kB; ++j) {
        const unsigned e = e0 + j * kThreads + tid;
        k[j] = (e < nc) ? __ldcg(cand + e) : 0ull;
      }
#pragma unroll
      for (int j = 0; j < kB; ++j) {
        if (e0 + j * kThreads >= nc) break;                                 // block-uniform
        f(e0 + j * kThreads + tid < nc && (unsigned)(k[j] >> 32) >= thr_final, k[j]);
      }
    }
  };
  for_each_survivor([&](bool ok, u64 k) {
    append_if(ok, k, s.keys, &s.cnt, (unsigned)kKeyCap);
    if (ok) hist_add(s.hist, fine_bin((unsigned)(k >> 32)));
  });
  __syncthreads();
  dbg_stamp(g.dbg, 7);
  select_sort_emit(a, g, s, b, s.keys, (int)s.cnt, kKeyCap,
                   [&](auto f) { for_each_survivor([&](bool ok, u64 k) { f(ok ? k : 0ull); }); });   // convergent: f may vote
  dbg_stamp(g.dbg, 10);
  // ---- leave the per-sample state zeroed for the next launch ---------------------------------
  __syncthreads();
  for (int j = 0; j < kCoarseBins / kThreads; ++j) ghist[tid + j * kThreads] = 0u;
  if (tid == 0) {
    g.state[b].cand_cnt = 0u;
    g.state[b].thr_bits = 0u;
  }
}

__global__ void __launch_bounds__(kThreads)
decode_merge_kernel(const cnh_decode_args a, const DecGeo g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  merge_sample(a, g, *reinterpret_cast<MergeSmem*>(smem_raw), (int)blockIdx.x);
}

template <int kRows>
__global__ void __launch_bounds__(kThreads)
decode_tiles_kernel(const __grid_constant__ CUtensorMap tmap, const cnh_decode_args a, const DecGeo g) {
  typedef DecSmemT<kRows * kCols> DecSmem;
  constexpr int kTileFloats = tile_floats(kRows);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  DecSmem& s = *reinterpret_cast<DecSmem*>(smem_raw);
  float* const ring = reinterpret_cast<float*>(smem_raw + sizeof(DecSmem));
  const int S = g.n_stages;
  const int tid = threadIdx.x;
  const int K = a.K;
  const int BW = g.box_w;
  // this CTA's contiguous tile range
  const long long lo = g.n_tiles * blockIdx.x / gridDim.x, hi = g.n_tiles * (blockIdx.x + 1) / gridDim.x;
  dbg_stamp(g.dbg, 0);
  if (lo >= hi) return;

  // tile cursor: (sample, class, tile row, tile column), advanced without divisions
  struct Cursor { int b, c, ty, tx; };
  auto cursor_at = [&](long long t) {
    Cursor q;
    q.b = (int)(t / g.tiles_per_sample);
    const int ts = (int)(t - (long long)q.b * g.tiles_per_sample);
    q.c = ts / g.tiles_per_plane;
    const int tp = ts - q.c * g.tiles_per_plane;
    q.ty = tp / g.tiles_x;
    q.tx = tp - q.ty * g.tiles_x;
    return q;
  };
  auto advance = [&](Cursor& q) {
    if (++q.tx == g.tiles_x) {
      q.tx = 0;
      if (++q.ty == g.tiles_y) {
        q.ty = 0;
        if (++q.c == a.C) { q.c = 0; ++q.b; }
      }
    }
  };
  auto issue_tile = [&](const Cursor& q, int buf) {       // thread 0 only
    mbar_expect_tx(&s.mbar[buf], (unsigned)(BW * (kRows + 2) * sizeof(float)));
    tma_load_3d(ring + (size_t)buf * kTileFloats, &tmap, &s.mbar[buf], q.tx * kCols - kPadL, q.ty * kRows - 1,
                q.b * a.C + q.c);
  };
  Cursor cur = cursor_at(lo), pre = cur;                  // tile being processed / next tile to prefetch
  long long pre_t = lo;

  if (tid == 0) {
    s.cnt = 0;
    s.cnt2 = 0;
    if (g.use_tma) {
      for (int i = 0; i < S; ++i) mbar_init(&s.mbar[i], 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      for (int i = 0; i < S && pre_t < hi; ++i, ++pre_t, advance(pre)) issue_tile(pre, i);
    }
  }
#pragma unroll
  for (int j = 0; j < kFineBins / 2 / kThreads; ++j) s.hist[tid + j * kThreads] = 0u;

  int cur_b = cur.b;
  long long dbg_heavy = 0, dbg_cands = 0;
  unsigned stage_n = 0;                     // keys staged for cur_b (uniform)
  unsigned tiles_staged = 0;                // tiles of cur_b processed since the last flush (uniform)
  unsigned thr = 0;                         // pruning threshold (score bits) for cur_b
  if (tid == 0) s.sh_thr = __ldcg(&g.state[cur_b].thr_bits);
  __syncthreads();
  thr = s.sh_thr;

  // flush the staged keys of sample b into a dense slice of its candidate list
  auto flush = [&](int b) {
    if (stage_n == 0) return;
    if (tid == 0) s.sh_base = atomicAdd(&g.state[b].cand_cnt, stage_n);
    __syncthreads();
    u64* dst = g.cand + (long long)b * g.tiles_per_sample * g.slot + s.sh_base;
    for (unsigned i = tid; i < stage_n; i += kThreads) dst[i] = s.stage[i];
    __syncthreads();
    stage_n = 0;
  };

  for (long long t = lo; t < hi; ++t) {
    const bool dbg_it = (t == lo + 2) && ((int)blockIdx.x >= a.B);
    if (dbg_it) dbg_stamp(g.dbg, 6);
    const int buf = (int)((t - lo) % S), phase = (int)(((t - lo) / S) & 1);
    const int b = cur.b, c = cur.c, y0 = cur.ty * kRows, x0 = cur.tx * kCols;
    if (b != cur_b) {
      flush(cur_b);
      cur_b = b;
      tiles_staged = 0;
      if (tid == 0) s.sh_thr = __ldcg(&g.state[b].thr_bits);
      __syncthreads();
      thr = s.sh_thr;
    }
    const int rows = min(kRows, a.H - y0), cols = min(kCols, a.W - x0);
    float* tile = ring + (size_t)buf * kTileFloats;
    if (g.use_tma) {
      mbar_wait(&s.mbar[buf], (unsigned)phase);
    } else {
      const float* src = a.heat + ((long long)b * a.C + c) * g.HW;
      for (int i = tid; i < (kRows + 2) * BW; i += kThreads) {
        const int r = i / BW, cc = i - r * BW;
        const int gy = y0 - 1 + r, gx = x0 - kPadL + cc;
        tile[i] = (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) ? __ldcs(src + (long long)gy * a.W + gx) : 0.f;
      }
      __syncthreads();
    }
    if (a.apply_sigmoid) {                  // export.py:31-33: logits in, clamp(sigmoid) fused
      for (int i = tid; i < (kRows + 2) * BW; i += kThreads) {
        const int r = i / BW, cc = i - r * BW;
        const int gy = y0 - 1 + r, gx = x0 - kPadL + cc;
        const bool in = (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W);
        tile[i] = in ? clamp_prob(1.0f / (1.0f + expf(-tile[i]))) : 0.f;
      }
      __syncthreads();
    }
    if (t == lo) dbg_stamp(g.dbg, 1);
    if (dbg_it) dbg_stamp(g.dbg, 7);

    scan_tile<kRows, kWarps>(s, s.keys, &s.cnt, tile, BW, rows, cols, thr, (unsigned)c * (unsigned)g.HW, y0, x0, a.W);
    if (dbg_it) dbg_stamp(g.dbg, 8);
    __syncthreads();                          // tile[buf] is free; keys / histogram complete
    const int n = (int)s.cnt;
    if (tid == 0 && g.use_tma && pre_t < hi) { issue_tile(pre, buf); ++pre_t; advance(pre); }   // refill the ring
    dbg_heavy += (n > K); dbg_cands += n;
    if (t == lo) dbg_stamp(g.dbg, 3);
    if (dbg_it) dbg_stamp(g.dbg, 9);

    if (n > 0) {
      // publish this tile's counts to the sample's global histogram (fire-and-forget REDs)
      unsigned cc[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const unsigned w = s.hist[8 * tid + j];
        cc[j >> 1] += (w & 0xffffu) + (w >> 16);
      }
      unsigned* ghist = g.ghist + (long long)b * kCoarseBins;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (cc[j]) atomicAdd(&ghist[4 * tid + j], cc[j]);
      // make room in the staging buffer
      const unsigned incoming = (unsigned)min(n, g.slot);
      if (stage_n + incoming > (unsigned)kStageCap) flush(b);
      if (n > K) {
        find_kth_bin(s, (unsigned)K);
        const unsigned keep_n = s.sh_above + s.sh_inbin;
        if (tid == 0) s.cnt2 = 0;
        __syncthreads();
        if (keep_n <= (unsigned)g.slot) {
          const int tbin = (int)s.sh_bin;
          for (int i0 = 0; i0 < n; i0 += kThreads) {
            const int i = i0 + tid;
            const u64 k = (i < n) ? s.keys[i] : 0ull;
            append_if((i < n) && fine_bin((unsigned)(k >> 32)) >= tbin, k, s.stage + stage_n, &s.cnt2,
                      (unsigned)g.slot);
          }
          stage_n += keep_n;
        } else {                               // massive ties inside one bin: exact selection
          const u64* keys = s.keys;
          const u64 T = radix_select_kth([&](auto f) { for (int i = tid; i < n; i += kThreads) f(keys[i]); }, K, s);
          for (int i0 = 0; i0 < n; i0 += kThreads) {
            const int i = i0 + tid;
            const u64 k = (i < n) ? s.keys[i] : 0ull;
            append_if((i < n) && k >= T, k, s.stage + stage_n, &s.cnt2, (unsigned)g.slot);
          }
          stage_n += (unsigned)K;
        }
        // a heavy tile republishes the sample's threshold from everything known so far
        __syncthreads();
        if (t + 1 < hi) {                      // (pointless if this CTA has no further tile to prune)
          const unsigned t_new = global_threshold(ghist, (unsigned)K, s);
          if (tid == 0 && t_new > thr) atomicMax(&g.state[b].thr_bits, t_new);
          if (t_new > thr) thr = t_new;
        }
      } else {
        for (int i = tid; i < n; i += kThreads) s.stage[stage_n + i] = s.keys[i];
        stage_n += (unsigned)n;
      }
      __syncthreads();
      // clean the local histogram and counters for the next tile
#pragma unroll
      for (int j = 0; j < kFineBins / 2 / kThreads; ++j) s.hist[tid + j * kThreads] = 0u;
      if (tid == 0) { s.cnt = 0; s.cnt2 = 0; }
    }
    if (dbg_it) dbg_stamp(g.dbg, 10);
    ++tiles_staged;
    advance(cur);
    // keep the pruning threshold current: for the first tiles of a sample just read the published
    // word; every 8th tile recompute it from the sample's global histogram and republish
    if (((t - lo) & 7) == 7) {
      const unsigned t_new = global_threshold(g.ghist + (long long)b * kCoarseBins, (unsigned)K, s);
      if (t_new > thr) {
        thr = t_new;
        if (tid == 0) atomicMax(&g.state[b].thr_bits, t_new);
      }
    } else if (tiles_staged <= 2) {
      if (tid == 0) s.sh_thr = __ldcg(&g.state[b].thr_bits);
      __syncthreads();
      if (s.sh_thr > thr) thr = s.sh_thr;
    }
  }
  dbg_stamp(g.dbg, 4);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (g.dbg && tid == 0) { long long* d = g.dbg + (long long)blockIdx.x * 16; d[12] = dbg_heavy; d[13] = dbg_cands; d[14] = thr; d[15] = hi - lo; }
  flush(cur_b);
}

// ---- lean scan for the cluster kernel: peaks >= thr of one staged tile -> keys, no histogram ---------
// The tile is staged WITHOUT halo columns: rows of W floats (W <= 128, row stride W), first row = the row above
// the tile; rows outside the image hold 0.  One row per warp, lane owns columns [4*lane, 4*lane+4); colmask
// (4 bits) masks the lane's columns beyond W (those lanes read the following row: harmless).  3x3 maximum =
// vertical max of three LDS.128, then horizontal max with the neighbours' edge columns by shuffle; the
// columns left of 0 and right of W-1 are the max-pool padding (0 is neutral: heat >= 0).
__device__ __forceinline__ void scan_rows_lean(u64* keys, unsigned* key_cnt, const float* tile, int W, unsigned colmask,
                                               bool last_lane, unsigned thr, unsigned flat_row0) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* row = tile + (warp + 1) * W + 4 * lane;
  const float4 mid = *reinterpret_cast<const float4*>(row);
  const float thr_eff = fmaxf(__uint_as_float(thr), __uint_as_float(1u));      // >= thr and > 0
  unsigned pass = (mid.x >= thr_eff ? 1u : 0u) | (mid.y >= thr_eff ? 2u : 0u) | (mid.z >= thr_eff ? 4u : 0u) |
                  (mid.w >= thr_eff ? 8u : 0u);
  pass &= colmask;
  if (__ballot_sync(0xffffffffu, pass != 0u) == 0u) return;                      // nothing in this row can matter
  const float4 up = *reinterpret_cast<const float4*>(row - W);
  const float4 dn = *reinterpret_cast<const float4*>(row + W);
  const float v0 = fmaxf(fmaxf(up.x, mid.x), dn.x), v1 = fmaxf(fmaxf(up.y, mid.y), dn.y);
  const float v2 = fmaxf(fmaxf(up.z, mid.z), dn.z), v3 = fmaxf(fmaxf(up.w, mid.w), dn.w);
  float left = __shfl_up_sync(0xffffffffu, v3, 1), right = __shfl_down_sync(0xffffffffu, v0, 1);
  if (lane == 0) left = 0.f;
  if (last_lane) right = 0.f;
  const float h0 = fmaxf(fmaxf(left, v0), v1), h1 = fmaxf(fmaxf(v0, v1), v2);
  const float h2 = fmaxf(fmaxf(v1, v2), v3), h3 = fmaxf(fmaxf(v2, v3), right);
  const unsigned flags = pass & ((mid.x == h0 ? 1u : 0u) | (mid.y == h1 ? 2u : 0u) | (mid.z == h2 ? 4u : 0u) |
                                 (mid.w == h3 ? 8u : 0u));
  const unsigned any = __ballot_sync(0xffffffffu, flags != 0u);
  if (any == 0u) return;
  const int mine = __popc(flags);
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    incl += (lane >= o) ? v : 0;
  }
  unsigned base = 0;
  if (lane == 31) base = atomicAdd(key_cnt, (unsigned)incl);
  base = __shfl_sync(0xffffffffu, base, 31);
  u64* dst = keys + base + (unsigned)(incl - mine);
  const unsigned nflat = 0xffffffffu - (flat_row0 + 4u * (unsigned)lane);       // ~flat of the lane's first column
  if (flags & 1u) *dst++ = ((u64)__float_as_uint(mid.x) << 32) | (u64)nflat;
  if (flags & 2u) *dst++ = ((u64)__float_as_uint(mid.y) << 32) | (u64)(nflat - 1u);
  if (flags & 4u) *dst++ = ((u64)__float_as_uint(mid.z) << 32) | (u64)(nflat - 2u);
  if (flags & 8u) *dst++ = ((u64)__float_as_uint(mid.w) << 32) | (u64)(nflat - 3u);
}

// RW rows per warp (the 256-thread group scan: four tiles of a round are scanned concurrently, one per
// group of 8 warps; RW = tile rows / 8).  Same test as scan_rows_lean with a rolling three-row window (6 LDS.128 for 4 rows);
// the scores of flagged pixels are re-read from shared memory when the keys are written.  Appends are
// bounded by `cap`: the counter may run past it (the caller detects that and rescans the round serially).
template <int RW>
__device__ __forceinline__ void scan_rows_group(u64* keys, unsigned* key_cnt, unsigned cap, const float* tile, int W,
                                                unsigned colmask, bool last_lane, unsigned thr, unsigned flat_tile0, int gwarp) {
  const int lane = threadIdx.x & 31;
  const float* base_row = tile + (RW * gwarp) * W + 4 * lane;                     // staged row above the warp's first row
  const float thr_eff = fmaxf(__uint_as_float(thr), __uint_as_float(1u));      // >= thr and > 0
  float4 up = *reinterpret_cast<const float4*>(base_row);
  float4 mid = *reinterpret_cast<const float4*>(base_row + W);
  unsigned flags = 0;                                                             // bit 4*rr + e
#pragma unroll
  for (int rr = 0; rr < RW; ++rr) {
    const float4 dn = *reinterpret_cast<const float4*>(base_row + (rr + 2) * W);
    // row-level early-out on the lane's maximum (3 instructions); the per-pixel bits only when a lane passes
    const bool any = colmask != 0u && fmaxf(fmaxf(mid.x, mid.y), fmaxf(mid.z, mid.w)) >= thr_eff;
    if (__ballot_sync(0xffffffffu, any) != 0u) {
      unsigned pass = (mid.x >= thr_eff ? 1u : 0u) | (mid.y >= thr_eff ? 2u : 0u) | (mid.z >= thr_eff ? 4u : 0u) |
                      (mid.w >= thr_eff ? 8u : 0u);
      pass &= colmask;
      const float v0 = fmaxf(fmaxf(up.x, mid.x), dn.x), v1 = fmaxf(fmaxf(up.y, mid.y), dn.y);
      const float v2 = fmaxf(fmaxf(up.z, mid.z), dn.z), v3 = fmaxf(fmaxf(up.w, mid.w), dn.w);
      float left = __shfl_up_sync(0xffffffffu, v3, 1), right = __shfl_down_sync(0xffffffffu, v0, 1);
      if (lane == 0) left = 0.f;
      if (last_lane) right = 0.f;
      const float h0 = fmaxf(fmaxf(left, v0), v1), h1 = fmaxf(fmaxf(v0, v1), v2);
      const float h2 = fmaxf(fmaxf(v1, v2), v3), h3 = fmaxf(fmaxf(v2, v3), right);
      const unsigned f = pass & ((mid.x == h0 ? 1u : 0u) | (mid.y == h1 ? 2u : 0u) | (mid.z == h2 ? 4u : 0u) |
                                 (mid.w == h3 ? 8u : 0u));
      flags |= f << (4 * rr);
    }
    up = mid;
    mid = dn;
  }
  if (__ballot_sync(0xffffffffu, flags != 0u) == 0u) return;
  const int mine = __popc(flags);
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    incl += (lane >= o) ? v : 0;
  }
  unsigned base = 0;
  if (lane == 31) base = atomicAdd(key_cnt, (unsigned)incl);
  base = __shfl_sync(0xffffffffu, base, 31);
  unsigned pos = base + (unsigned)(incl - mine);
  const unsigned flat_lane0 = flat_tile0 + (unsigned)(RW * gwarp) * (unsigned)W + 4u * (unsigned)lane;
  while (flags) {
    const int bit = __ffs(flags) - 1;
    flags &= flags - 1u;
    const int rr = bit >> 2, e = bit & 3;
    const float v = base_row[(rr + 1) * W + e];
    if (pos < cap) keys[pos] = ((u64)__float_as_uint(v) << 32) | (u64)(0xffffffffu - (flat_lane0 + (unsigned)(rr * W + e)));
    ++pos;
  }
}

__device__ __forceinline__ void finish_sample(const cnh_decode_args& a, const DecGeo& g, MergeSmem& s, int b);

// ---- cluster path: one thread-block cluster per sample, ONE launch, no global scratch ------------------
// The CS CTAs of a cluster split the sample's tiles (tile t -> CTA t % CS), each walking its share
// with a TMA ring and keeping a running candidate set in shared memory: whenever the set outgrows
// K + kSlack keys it is cut back to the bins >= that of its K-th score (exact radix select on heavy
// ties) and the cut becomes the pruning threshold of the following tiles.  At the end every CTA
// pushes its <= K + kSlack keys into the leader's shared memory through DSMEM (one remote atomic for
// the slice, then remote stores), a cluster barrier publishes them, and the leader runs the final
// selection, sort, filler, gather and box assembly.  Versus the two-kernel path: no candidate lists,
// histograms or tickets in global memory, no second launch, nothing to leave zeroed.
constexpr int kClGroups = 4;               // tiles scanned concurrently (one per group of 8 warps) = one round
constexpr int kClThreads = 1024;
// Two shapes of the same kernel.  R = 32: tiles of 32 rows, ring of 2 rounds -- the fewest rounds, for short
// walks (latency-bound, cfg2).  R = 16: tiles of 16 rows, ring of 4 rounds -- twice the rounds of copies in
// flight, for long walks (cfg5: a round lasts about one copy latency, so bytes in flight are throughput).
template <int R>
struct ClCfg {
  static constexpr int kRows = R;
  static constexpr int kDepth = R == 32 ? 2 : 4;                 // rounds in the ring
  static constexpr int kStages = kDepth * kClGroups;             // ring slots
  static constexpr int kTileFloats = (R + 2) * kCols + 128;      // R + 2 rows of <= 128 floats + slack for masked lanes
  static constexpr int kTile = R * kCols;                        // most peaks one tile can add
  static constexpr int kCap = R == 32 ? 2 * kTile : 3 * kTile;   // running candidate set; cut whenever it exceeds kTile
  static constexpr int kRowsPerWarp = R / kWarps;                // group scan: 8 warps per tile
};
template <int R>
struct __align__(128) ClSmemT {             // followed by the TMA ring (the leader's doubles as the inbox)
  u64 stage[ClCfg<R>::kCap];               // the set (the scan appends to it); final stage: `sorted`
  unsigned hist[kFineBins / 2];            // packed fine histogram of the set, built per cut; final stage: `sel`
  u64 mbar[ClCfg<R>::kStages];
  u64 sh_prefix;
  unsigned cnt;                            // size of the set
  unsigned cnt2;
  unsigned sh_need;
  unsigned sh_flag;
  unsigned sh_thr;
  unsigned sh_bin;
  unsigned sh_above;
  unsigned sh_inbin;
  unsigned sh_base;
  unsigned fin_cnt;                        // leader: keys received from the cluster
  unsigned warp_tot[kWarps];
};
static_assert(ClCfg<32>::kCap >= kMaxK && ClCfg<16>::kCap >= kMaxK, "stage also holds the sorted output");

template <int R>
__global__ void __launch_bounds__(kClThreads, 1)
decode_cluster_kernel(const __grid_constant__ cnh_decode_args a, const __grid_constant__ DecGeo g) {
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int b = (int)(blockIdx.x / (unsigned)CS);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  typedef ClCfg<R> Cfg;
  constexpr int kClRows = Cfg::kRows, kClStages = Cfg::kStages, kClTileFloats = Cfg::kTileFloats;
  constexpr int kClTile = Cfg::kTile, kClCap = Cfg::kCap, kDepth = Cfg::kDepth;
  typedef ClSmemT<R> ClSmem;
  ClSmem& s = *reinterpret_cast<ClSmem*>(smem_raw);
  float* const ring = reinterpret_cast<float*>(smem_raw + sizeof(ClSmem));
  u64* const inbox = reinterpret_cast<u64*>(ring);          // the leader's ring becomes the inbox once every CTA has scanned
  const int tid = threadIdx.x;
  const bool group0 = tid < kThreads;                       // warps 0-7 run the 256-thread selection helpers
  const int K = a.K, W = a.W, H = a.H;
  const unsigned slot = (unsigned)g.slot;
  const int n_my = rank < g.tiles_per_sample ? (g.tiles_per_sample - rank + CS - 1) / CS : 0;
  const float* const sample = a.heat + (long long)b * a.C * g.HW;
  dbg_stamp(g.dbg, 0);

  // Tile t of the sample = rows [32*ty, 32*ty+32) of class plane c, t = c*tiles_y + ty (W <= 128: one tile
  // column).  This CTA walks t = rank, rank+CS, ...; every thread keeps two cursors, advanced without
  // divisions: the tile being scanned and the tile being prefetched kClStages ahead.
  struct Cursor { int c, ty; };
  struct Stride { int c, ty; };                             // a tile-index step split by tiles_y, computed once
  auto stride_of = [&](int by) { Stride d; d.c = by / g.tiles_y; d.ty = by - d.c * g.tiles_y; return d; };
  auto advance = [&](Cursor& q, const Stride& d) {
    q.c += d.c;
    q.ty += d.ty;
    if (q.ty >= g.tiles_y) { q.ty -= g.tiles_y; ++q.c; }
  };
  auto cursor_at = [&](int t) { Cursor q; q.c = t / g.tiles_y; q.ty = t - q.c * g.tiles_y; return q; };
  const Stride step1 = stride_of(CS), step_round = stride_of(kClGroups * CS);
  // Stage tile q into ring slot buf: the rows of the tile plus one halo row above and below are contiguous
  // in global memory (one bulk copy, SASS UBLKCP); halo/tail rows outside the image are zero-filled by the
  // calling warp (max-pool padding: 0 is neutral because heat >= 0).  Called by ONE warp.
  auto stage_tile = [&](const Cursor& q, int buf) {
    float* dst = ring + (size_t)buf * kClTileFloats;
    const int lane = tid & 31;
    const int y0 = q.ty * kClRows;
    const int ylo = max(y0 - 1, 0), yhi = min(y0 + kClRows + 1, H);       // image rows [ylo, yhi) are copied
    const int r_lo = ylo - (y0 - 1), r_hi = yhi - (y0 - 1);               // -> tile rows [r_lo, r_hi) of 34
    if (r_lo > 0 || r_hi < kClRows + 2) {
      for (int i = lane; i < r_lo * W; i += 32) dst[i] = 0.f;
      for (int i = r_hi * W + lane; i < (kClRows + 2) * W; i += 32) dst[i] = 0.f;
    }
    if (lane == 0) {
      const unsigned bytes = (unsigned)((yhi - ylo) * W) * 4u;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // earlier generic accesses of the slot
      mbar_expect_tx(&s.mbar[buf], bytes);
      bulk_load_1d(dst + r_lo * W, sample + (long long)q.c * g.HW + (long long)ylo * W, bytes, &s.mbar[buf]);
    }
  };

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next PDL launch (the next step's loss) may be placed
  asm volatile("griddepcontrol.wait;" ::: "memory");        // PDL: the producer of `heat` has completed
  // Launched behind a kernel that left per-sample candidate lists (the streaming decode kernel, or a detection-loss
  // launch): the usual sample is FINISHED from its lists by the first 256 threads of its cluster's leader (the other
  // CTAs and warps exit at once); a sample whose candidate buffers ran over is redone from the heat map by the whole
  // cluster, exactly (every CTA of the cluster reads the same word, before anyone can have cleared it: the leader does
  // so after the last cluster barrier).
  if (g.only_overflow) {
    if (__ldcg(&g.cl.state[b].overflow) == 0u) {
      if (rank != 0 || tid >= kThreads) return;
      finish_sample(a, g, *reinterpret_cast<MergeSmem*>(smem_raw), b);
      return;
    }
  }
  if (tid == 0) {
    s.cnt = 0;
    s.cnt2 = 0;
    s.fin_cnt = 0;
    for (int i = 0; i < kClStages; ++i) mbar_init(&s.mbar[i], 1);
  }
  __syncthreads();
  const int gq = tid >> 8, gwarp = (tid >> 5) & 7;           // scan group (one tile of the round each) and warp in it
  const int wid = tid >> 5;
  // warp w < kClStages stages ring slot w: tiles w, w + kClStages, ... of this CTA's walk (j = w + 8k)
  Cursor pre = cursor_at(rank + min(wid, kClStages - 1) * CS);
  const Stride step_ring = stride_of(kClStages * CS);
  if (wid < kClStages && wid < n_my) stage_tile(pre, wid);
  advance(pre, step_ring);
  __syncthreads();                                          // the staging warps' zero-filled halo rows (generic stores: the
                                                            // mbarrier only covers the bulk copy) before anyone scans
  Cursor mine = cursor_at(rank + gq * CS);                  // the tile this thread's group scans in the current round
  Cursor round0 = cursor_at(rank);                          // first tile of the current round (uniform)

  // In-place compaction of arr[0..n) by the whole CTA, 1024 keys per step: a step's keys are in registers
  // before anything is written, and writes only go below the step's first key.  New size -> *out_cnt.
  auto compact_arr = [&](u64* arr, const unsigned n, unsigned* out_cnt, auto pred) {
    if (tid == 0) s.cnt2 = 0;
    for (unsigned i0 = 0; i0 < n; i0 += kClThreads) {
      const unsigned i = i0 + tid;
      const u64 k = (i < n) ? arr[i] : 0ull;
      const bool keep = (i < n) && pred(k);
      __syncthreads();
      append_if(keep, k, arr, &s.cnt2, n);
    }
    __syncthreads();
    if (tid == 0) *out_cnt = s.cnt2;
  };
  auto compact_all = [&](const unsigned n, auto pred) { compact_arr(s.stage, n, &s.cnt, pred); };
  // packed fine histogram of stage[0..n) (built only when a cut needs it)
  auto build_hist_of = [&](const u64* arr, const unsigned n) {
    for (int w = tid; w < kFineBins / 2; w += kClThreads) s.hist[w] = 0u;
    __syncthreads();
    for (unsigned i = tid; i < n; i += kClThreads) hist_add(s.hist, fine_bin((unsigned)(arr[i] >> 32)));
    __syncthreads();
  };
  auto build_hist = [&](const unsigned n) { build_hist_of(s.stage, n); };
  // K-th key of the set, exactly (heavy ties inside one histogram bin); clobbers s.hist.  Everyone gets T.
  auto exact_cut_key = [&](const unsigned n) {
    if (group0) {
      const u64 T = radix_select_kth([&](auto f) { for (unsigned i = tid; i < n; i += kThreads) f(s.stage[i]); }, K, s);
      if (tid == 0) s.sh_prefix = T;
    }
    __syncthreads();
    return s.sh_prefix;
  };

  unsigned thr = 0;                                         // pruning threshold, score bits (uniform)
  const int lane_col = 4 * (tid & 31);
  const unsigned colmask = W - lane_col >= 4 ? 15u : (W - lane_col <= 0 ? 0u : ((1u << (W - lane_col)) - 1u));
  const bool last_lane = lane_col + 4 >= W;                 // the column right of this lane's is outside the image
  // Cut the set back to (a superset of) its top K; the cut becomes the pruning threshold.
  auto cut = [&](const unsigned n) {
    build_hist(n);
    if (group0) find_kth_bin(s, (unsigned)K);
    __syncthreads();
    const unsigned keep_n = s.sh_above + s.sh_inbin;
    const int tbin = (int)s.sh_bin;
    unsigned thr_new;
    if (keep_n <= (unsigned)kClTile) {
      compact_all(n, [&](u64 k) { return fine_bin((unsigned)(k >> 32)) >= tbin; });
      thr_new = __float_as_uint((float)tbin * (1.0f / (float)kFineBins));
    } else {                                                // a plateau of ties fills the cut bin: exact selection
      const u64 T = exact_cut_key(n);
      compact_all(n, [&](u64 k) { return k >= T; });
      thr_new = (unsigned)(T >> 32);
    }
    if (thr_new > thr) thr = thr_new;
    __syncthreads();
  };
  auto tile_rows = [&](int y0, int& r_lo, int& r_hi) {      // staged rows that hold image rows
    r_lo = max(y0 - 1, 0) - (y0 - 1);
    r_hi = min(y0 + kClRows + 1, H) - (y0 - 1);
  };

  const int n_rounds = (n_my + kClGroups - 1) / kClGroups;
  for (int r = 0; r < n_rounds; ++r) {
    const int half = (r % kDepth) * kClGroups, phase = (r / kDepth) & 1;
    const int in_round = min(kClGroups, n_my - r * kClGroups);
    const unsigned n_before = s.cnt;                        // <= kClTile (invariant)
    if (gq < in_round) {
      const Cursor q = mine;
      const int y0 = q.ty * kClRows;
      float* tile = ring + (size_t)(half + gq) * kClTileFloats;
      mbar_wait(&s.mbar[half + gq], (unsigned)phase);
      if (a.apply_sigmoid) {                                // export.py:31-33: logits in, clamp(sigmoid) fused
        int r_lo, r_hi;
        tile_rows(y0, r_lo, r_hi);
        for (int i = r_lo * W + (tid & 255); i < r_hi * W; i += 256) tile[i] = clamp_prob(1.0f / (1.0f + expf(-tile[i])));
        asm volatile("bar.sync %0, 256;" ::"r"(2 + gq) : "memory");
      }
      scan_rows_group<Cfg::kRowsPerWarp>(s.stage, &s.cnt, (unsigned)kClCap, tile, W, colmask, last_lane, thr,
                 (unsigned)q.c * (unsigned)g.HW + (unsigned)y0 * (unsigned)W, gwarp);
    }
    __syncthreads();                                        // the round's tiles are scanned
    if (r == 0) dbg_stamp(g.dbg, 1);
    unsigned n = s.cnt;
    if (n > (unsigned)kClCap) {
      // plateaus: the round found more peaks than the buffer holds.  Rescan its tiles one at a time with
      // the whole CTA (one row per warp), cutting in between.
      __syncthreads();
      if (tid == 0) s.cnt = n_before;
      __syncthreads();
      Cursor q = round0;
      for (int t = 0; t < in_round; ++t, advance(q, step1)) {
        if ((tid >> 5) < kClRows)                             // one row per warp
          scan_rows_lean(s.stage, &s.cnt, ring + (size_t)(half + t) * kClTileFloats, W, colmask, last_lane, thr,
                         (unsigned)q.c * (unsigned)g.HW + (unsigned)(q.ty * kClRows + (tid >> 5)) * (unsigned)W);
        __syncthreads();
        n = s.cnt;
        if (n > (unsigned)kClTile) { cut(n); n = s.cnt; }
      }
    }
    advance(mine, step_round);
    advance(round0, step_round);
    if (wid >= half && wid < half + kClGroups) {            // refill the freed slots of the ring: round r + kDepth
      if ((r + kDepth) * kClGroups + (wid - half) < n_my) stage_tile(pre, wid);
      advance(pre, step_ring);
    }
    // Cut when the next round might overflow the invariant and, on long walks, after rounds 1, 2, 4, 8, ...
    // so that the pruning threshold tightens early.
    if (r + 1 < n_rounds && (n > (unsigned)kClTile || (n_rounds >= 4 && n > slot && ((r + 1) & r) == 0))) cut(n);
  }
  dbg_stamp(g.dbg, 2);
  cluster.barrier_arrive();                                 // this CTA no longer reads its ring (the leader's is the inbox)

  // ---- final cut fused with the push into the leader's inbox (DSMEM) --------------------------------
  const unsigned n = s.cnt;
  int mode = 0, tbin = 0;                                   // 0: send everything, 1: bins >= tbin, 2: keys >= T
  u64 T = 0;
  unsigned send_n = n;
  // What travels to the leader is at most K + kSlack keys per CTA (measured: letting the leader cut the
  // raw sets of a whole cluster is slower than one cut per CTA in parallel).
  if (n > slot) {
    build_hist(n);
    if (group0) find_kth_bin(s, (unsigned)K);
    __syncthreads();
    send_n = s.sh_above + s.sh_inbin;
    tbin = (int)s.sh_bin;
    mode = 1;
    if (send_n > slot) {
      T = exact_cut_key(n);
      send_n = (unsigned)K;
      mode = 2;
    }
  }
  dbg_stamp(g.dbg, 5);
  cluster.barrier_wait();                                   // every CTA, the leader included, is done with its ring
  dbg_stamp(g.dbg, 6);
  unsigned* const r_cnt = cluster.map_shared_rank(&s.fin_cnt, 0);
  u64* const r_inbox = cluster.map_shared_rank(inbox, 0);
  if (tid == 0) {
    s.sh_base = send_n ? atomicAdd(r_cnt, send_n) : 0u;     // one remote atomic reserves the CTA's slice
    s.cnt2 = 0;
  }
  __syncthreads();
  {
    u64* const dst = r_inbox + s.sh_base;
    for (unsigned i0 = 0; i0 < n; i0 += kClThreads) {
      const unsigned i = i0 + tid;
      const u64 k = (i < n) ? s.stage[i] : 0ull;
      const bool keep = (i < n) && (mode == 0 || (mode == 1 ? fine_bin((unsigned)(k >> 32)) >= tbin : k >= T));
      append_if(keep, k, dst, &s.cnt2, send_n);
    }
  }
  dbg_stamp(g.dbg, 7);
  cluster.sync();                                           // release/acquire: the inbox is complete
  if (rank != 0) return;
  if (g.only_overflow) {                                    // redone here: the lists are void, the flag comes down
    if (tid == 0) g.cl.state[b].overflow = 0u;
    cand_state_clear(g, b);
  }
  dbg_stamp(g.dbg, 3);

  // ---- leader: final selection (all warps) + sort + gather (warps 0-7) -------------------------------------
  int m = (int)*reinterpret_cast<volatile unsigned*>(&s.fin_cnt);
  if (m > kThreads && m > K) {
    build_hist_of(inbox, (unsigned)m);
    if (group0) find_kth_bin(s, (unsigned)K);
    __syncthreads();
    const unsigned keep_n = s.sh_above + s.sh_inbin;
    const int tbin2 = (int)s.sh_bin;
    if (keep_n <= (unsigned)kMaxK) {                        // (else: ties, select_sort_emit selects exactly)
      // the histogram stays valid for select_sort_emit: it only looks at the bins from the top down to tbin2
      compact_arr(inbox, (unsigned)m, &s.fin_cnt, [&](u64 k) { return fine_bin((unsigned)(k >> 32)) >= tbin2; });
      __syncthreads();
      m = (int)keep_n;
    }
  }
  if (tid == 0) s.cnt2 = 0;
  __syncthreads();
  if (m <= kThreads) {
    // the usual case: <= 256 keys left.  Rank sort by the whole CTA (keys are unique: position = number of
    // larger keys), four lanes per key each scanning a quarter of the list.  The key's reg / wh values are
    // loaded BEFORE the scan (their latency hides under it) and land in shared memory at the key's rank.
    const int i = tid >> 2, part = tid & 3;
    const u64 k = (i < m) ? inbox[i] : 0ull;
    float pv[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    const bool loader = i < m && part == 0;
    if (loader) {
      const unsigned pix = (0xffffffffu - (unsigned)(k & 0xffffffffu)) % (unsigned)g.HW;
      if (a.reg) {
        pv[0] = __ldg(a.reg + ((long long)b * 2 + 0) * g.HW + pix);
        pv[1] = __ldg(a.reg + ((long long)b * 2 + 1) * g.HW + pix);
      }
      pv[2] = __ldg(a.wh + ((long long)b * a.D + 0) * g.HW + pix);
      pv[3] = __ldg(a.wh + ((long long)b * a.D + 1) * g.HW + pix);
      if (a.rotated) pv[4] = __ldg(a.wh + ((long long)b * a.D + 2) * g.HW + pix);
    }
    int rank = 0;
    if (i < m) {
      int j = part;
      for (; j + 12 < m; j += 16) {                         // four independent loads in flight
        const u64 k0 = inbox[j], k1 = inbox[j + 4], k2 = inbox[j + 8], k3 = inbox[j + 12];
        rank += (k0 > k) + (k1 > k) + (k2 > k) + (k3 > k);
      }
      for (; j < m; j += 4) rank += (inbox[j] > k);
    }
    rank += __shfl_xor_sync(0xffffffffu, rank, 1);
    rank += __shfl_xor_sync(0xffffffffu, rank, 2);
    float* const payload = reinterpret_cast<float*>(s.hist);      // the histogram is dead: 256 ranks x 5 floats
    if (loader) {
      s.stage[rank] = k;
#pragma unroll
      for (int c = 0; c < 5; ++c) payload[5 * rank + c] = pv[c];
    }
    __syncthreads();
    if (!group0) return;
    dbg_stamp(g.dbg, 11);
    emit_sorted(a, g, s, b, m, payload, m);
    dbg_stamp(g.dbg, 4);
    return;
  }
  if (!group0) return;
  if (m <= kThreads && tid < m) {                           // start the gather's cache lines on their way before sorting
    const unsigned flat = 0xffffffffu - (unsigned)(inbox[tid] & 0xffffffffu);
    const unsigned pix = flat % (unsigned)g.HW;
    if (a.reg) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a.reg + ((long long)b * 2 + 0) * g.HW + pix));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a.reg + ((long long)b * 2 + 1) * g.HW + pix));
    }
    for (int d = 0; d < (a.rotated ? 3 : 2); ++d)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a.wh + ((long long)b * a.D + d) * g.HW + pix));
  }
  select_sort_emit(a, g, s, b, inbox, m, kClStages * kClTileFloats / 2,
                   [&](auto f) { for (int i = tid; i < m; i += kThreads) f(inbox[i]); });
  dbg_stamp(g.dbg, 4);
}

// ================================================================================================
// Streaming path (long walks: many tiles per SM, e.g. the 80-class COCO-scale shard).  Same skeleton as the
// streaming detection-loss kernel: persistent CTAs of 9 warps, two per SM, NO block-wide barrier in steady state.
//   warp 8 (producer): walks this CTA's tiles of ITS sample (tile t = j, j + G, ... for CTA j of the sample's G
//           CTAs): waits for a free ring stage, forwards the candidates the consumers left in the stage's buffer
//           to the CTA's private slice of the sample's candidate list in global memory (plain stores, no atomic
//           round trip) while counting them into the sample's two-level global histogram (fire-and-forget REDs),
//           refreshes the pruning threshold from that histogram (loads issued one tile ahead of their use), and
//           issues the bulk copy (cp.async.bulk, SASS UBLKCP) of the next tile + halo rows.
//   warps 0-7 (consumers): each scans 4 rows of the staged tile (threshold-first, 3x3 test from shared memory)
//           and appends its peaks to the stage's buffer; one mbarrier arrival per warp frees the stage.
// The histogram counts only keys already forwarded, so every threshold derived from it is valid (at least K real
// peaks reach it) however far the CTAs of a sample have drifted apart.  ONE cluster launch follows (programmatic
// stream serialisation): the leader CTA of each sample's cluster selects, sorts and emits from the lists
// (finish_sample; the other CTAs exit at once).  A stage buffer or slice that runs over (plateaus of ties: thousands of
// equal scores) raises the sample's overflow flag; that sample's whole cluster then redoes it from the heat map, exactly.
// ================================================================================================
constexpr int kStRows = 32;                               // tile rows
constexpr int kStStages = 4;                              // ring depth
constexpr int kStCap = 1024;                              // candidate keys per stage buffer
constexpr int kStTileFloats = (kStRows + 2) * kCols + 128;   // + slack for the masked lanes of narrow maps
constexpr int kStThreads = kThreads + 32;
struct __align__(128) StSmem {
  u64 cand[kStStages][kStCap];
  u64 full[kStStages];
  u64 empty[kStStages];
  unsigned cnt[kStStages];
  unsigned thr[kStStages];
  unsigned thr_start;                                     // threshold after the first tile (start-up tiles were issued with 0)
  unsigned go;                                            // set by the producer once thr_start is valid
  int tile_c[kStStages];                                  // class plane of the staged tile, -1 = no more tiles
  int tile_ty[kStStages];
};
constexpr size_t kStSmemBytes = sizeof(StSmem) + (size_t)kStStages * kStTileFloats * sizeof(float);

__global__ void __launch_bounds__(kStThreads, 2)
decode_stream_kernel(const __grid_constant__ cnh_decode_args a, const __grid_constant__ DecGeo g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  StSmem& s = *reinterpret_cast<StSmem*>(smem_raw);
  float* const ring = reinterpret_cast<float*>(smem_raw + sizeof(StSmem));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = g.cl.G;
  const int b = (int)blockIdx.x / G, j = (int)blockIdx.x - b * G;
  const int W = a.W, H = a.H, K = a.K;
  const int n_mine = j < g.tiles_per_sample ? (g.tiles_per_sample - j + G - 1) / G : 0;
  const float* const sample = a.heat + (long long)b * a.C * g.HW;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");        // PDL: the producer of `heat` has completed
  dbg_stamp(g.dbg, 0);
  if (tid == 0) {
    for (int i = 0; i < kStStages; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], kWarps);
      s.cnt[i] = 0u;
    }
    s.go = 0u;
    s.thr_start = 0u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp == kWarps) {
    // ================= producer =================
    CandEmitter em;
    em.init(g.cl, b, j, K);
    auto every_key = [](u64) { return true; };               // every candidate of this kernel passed the full 3x3 test
    auto flush = [&](int st) {
      em.forward(s.cand[st], *reinterpret_cast<volatile unsigned*>(&s.cnt[st]), (unsigned)kStCap, every_key);
      if (lane == 0) s.cnt[st] = 0u;
    };
    int c = 0, ty = 0;                                       // cursor of the next tile to issue: t = j + i * G
    { const int t0 = j; c = t0 / g.tiles_y; ty = t0 - c * g.tiles_y; }
    const int dc = G / g.tiles_y, dty = G - dc * g.tiles_y;
    // Start-up: the ring is filled with the first kStStages tiles at once, but only the first is scanned without a
    // threshold.  The consumers then wait (s.go) until this warp has forwarded that tile's peaks and read back the
    // sample's histogram -- by then it holds the first tile of most CTAs of the sample -- while the other copies
    // keep landing.  Costs two L2 round trips once; saves every CTA thousands of useless candidates.
    for (int i = 0; i < n_mine + kStStages; ++i) {
      const int st = i % kStStages;
      if (i >= kStStages) {
        // ---- stage st is free once the eight consumer warps have arrived: forward its candidates ----
        mbar_wait(&s.empty[st], (unsigned)(((i / kStStages) - 1) & 1));
        flush(st);
        em.refresh_step(i, i < n_mine);
      }
      if (i < n_mine) {
        // ---- stage tile (c, ty): rows [y0-1, y0+33) of the plane, contiguous; rows outside the image are zero ----
        float* dst = ring + (size_t)st * kStTileFloats;
        const int y0 = ty * kStRows;
        const int ylo = max(y0 - 1, 0), yhi = min(y0 + kStRows + 1, H);
        const int r_lo = ylo - (y0 - 1), r_hi = yhi - (y0 - 1);
        if (r_lo > 0 || r_hi < kStRows + 2) {
          for (int q = lane; q < r_lo * W; q += 32) dst[q] = 0.f;
          for (int q = r_hi * W + lane; q < (kStRows + 2) * W; q += 32) dst[q] = 0.f;
        }
        __syncwarp();
        if (lane == 0) {
          s.tile_c[st] = c;
          s.tile_ty[st] = ty;
          s.thr[st] = em.thr;
          const unsigned bytes = (unsigned)((yhi - ylo) * W) * 4u;
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_expect_tx(&s.full[st], bytes);
          bulk_load_1d(dst + r_lo * W, sample + (long long)c * g.HW + (long long)ylo * W, bytes, &s.full[st]);
        }
        c += dc;
        ty += dty;
        if (ty >= g.tiles_y) { ty -= g.tiles_y; ++c; }
      } else if (i == n_mine) {
        if (lane == 0) {                                     // out of tiles: wake the consumers
          s.tile_c[st] = -1;
          mbar_arrive(&s.full[st]);
        }
      }
      if (n_mine > 0 && i == min(kStStages - 1, n_mine)) {
        // ---- every start-up copy is on its way: the first threshold (n_mine >= 1: G <= tiles per sample) ----
        mbar_wait(&s.empty[0], 0u);                          // tile 0 scanned
        flush(0);
        __threadfence();
        __nanosleep(600);                                    // the other CTAs of the sample are at the same point: let their REDs land
        em.refresh_blocking();
        if (lane == 0) {
          s.thr_start = em.thr;
          __threadfence_block();
          *reinterpret_cast<volatile unsigned*>(&s.go) = 1u;
        }
        __syncwarp();
      }
    }
    // ---- the slice once more against the latest threshold: most of what it holds is the first tile, scanned
    // without one; what the finish kernel has to read shrinks from thousands of keys per CTA to a few dozen ----
    // (the survivors' reg / wh cache lines are started towards L2: the finish kernel's gather then hits L2)
    em.reprune([&](u64 key) {
      const unsigned pix = (0xffffffffu - (unsigned)(key & 0xffffffffu)) % (unsigned)g.HW;
      if (a.reg) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.reg + ((long long)b * 2 + 0) * g.HW + pix));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.reg + ((long long)b * 2 + 1) * g.HW + pix));
      }
      for (int d = 0; d < (a.rotated ? 3 : 2); ++d)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.wh + ((long long)b * a.D + d) * g.HW + pix));
    });
    if (lane == 0) {
      g.cl.cta_cnt[(long long)b * G + j] = em.local_cnt;
      if (em.overflow) g.cl.state[b].overflow = 1u;
    }
    dbg_stamp(g.dbg, 2);
    return;
  }

  // ================= consumers =================
  const int lane_col = 4 * lane;
  const unsigned colmask = W - lane_col >= 4 ? 15u : (W - lane_col <= 0 ? 0u : ((1u << (W - lane_col)) - 1u));
  const bool last_lane = lane_col + 4 >= W;
#pragma unroll 1
  for (int i = 0;; ++i) {
    const int st = i % kStStages;
    mbar_wait(&s.full[st], (unsigned)((i / kStStages) & 1));    // acquires the tile's meta data
    const int c = s.tile_c[st];
    if (c < 0) break;
    if (i == 1) {                                            // start-up: wait for the threshold of the first tiles
      while (*reinterpret_cast<volatile unsigned*>(&s.go) == 0u) __nanosleep(64);
      __syncwarp();
    }
    const int ty = s.tile_ty[st];
    unsigned thr = s.thr[st];
    if (i >= 1 && i < kStStages) thr = max(thr, *reinterpret_cast<volatile unsigned*>(&s.thr_start));
    scan_rows_group<kStRows / kWarps>(s.cand[st], &s.cnt[st], (unsigned)kStCap, ring + (size_t)st * kStTileFloats, W,
                                      colmask, last_lane, thr,
                                      (unsigned)c * (unsigned)g.HW + (unsigned)(ty * kStRows) * (unsigned)W, warp);
    __syncwarp();
    if (lane == 0) mbar_arrive(&s.empty[st]);
  }
  if (tid == 0) dbg_stamp(g.dbg, 1);
}

// Finish of one sample from its candidate lists (256 threads; barriers are the named 256-thread barrier, so this runs
// as a kernel of its own or inside the first 256 threads of the cluster kernel's leader): threshold = lower edge of the
// fine bin of the K-th counted key (two dependent L2 loads, issued beside the load of the slice sizes), survivors of
// the G slices into shared memory (slices dealt to warps, 16-byte loads, four in flight per lane), selection + sort +
// filler + gather + boxes (select_sort_emit), and the sample's global state left zeroed for the next launch.
__device__ __forceinline__ void finish_sample(const cnh_decode_args& a, const DecGeo& g, MergeSmem& s, int b) {
  constexpr int kKeyCap = MergeSmem::kKeyCap;
  const int tid = threadIdx.x, lane = tid & 31;
  const int K = a.K, G = g.cl.G;
  unsigned* const shist = g.cl.shist + (long long)b * kSuperBins;
  unsigned* const fhist = g.cl.fhist + (long long)b * kFineBins;
  unsigned* const cta_cnt = g.cl.cta_cnt + (long long)b * G;
  const u64* const slices = g.cl.slices + (long long)b * G * kSliceCap;
  for (int q = 0; q < kFineBins / 2 / kThreads; ++q) s.hist[tid + q * kThreads] = 0u;
  if (tid == 0) { s.cnt = 0; s.cnt2 = 0; s.sh_thr = 0u; }
  dbg_stamp(g.dbg, 5);
  // ---- round trip 1: the slice sizes and (warp 0) the super bins, side by side ----
  unsigned* const n_slice = reinterpret_cast<unsigned*>(s.stage);           // [G] keys per slice (`stage` is free until the sort)
  for (int q = tid; q < G; q += kThreads) n_slice[q] = __ldcg(cta_cnt + q);
  uint2 sv = make_uint2(0u, 0u);
  if (tid < 32) sv = __ldcg(reinterpret_cast<const uint2*>(shist) + lane);
  group_sync();
  unsigned thr_final = 0u;                                   // 0: fewer than K counted peaks, keep everything
  // Candidates that came out of a tile scanned WITHOUT its halo rows (the detection-loss kernels emit them from their
  // own 32-row chunks) were tested against the neighbours inside the tile only: those of a tile's first / last row
  // are checked here against the three pixels of the row beyond it (read from the heat map: a few dozen keys).
  auto verified = [&](u64 key) -> bool {
    if (g.verify_rows == 0 || key == 0ull) return true;
    const unsigned flat = 0xffffffffu - (unsigned)(key & 0xffffffffu);
    const unsigned cls = flat / (unsigned)g.HW, pix = flat - cls * (unsigned)g.HW;
    const int y = (int)(pix / (unsigned)a.W), x = (int)(pix - (unsigned)y * (unsigned)a.W);
    const int r = y % g.verify_rows;
    int yy = -1;
    if (r == 0 && y > 0) yy = y - 1;
    else if (r == g.verify_rows - 1 && y < a.H - 1) yy = y + 1;
    if (yy < 0) return true;
    const float* row = a.heat + ((long long)b * a.C + cls) * g.HW + (long long)yy * a.W + x;
    // (clamped at the borders: a pixel read twice does not change the maximum; three independent loads)
    const float m = fmaxf(fmaxf(__ldcg(row - (x > 0 ? 1 : 0)), __ldcg(row)), __ldcg(row + (x < a.W - 1 ? 1 : 0)));
    return m <= __uint_as_float((unsigned)(key >> 32));
  };
  // Slices are dealt to the warps, kSl of them per warp and round: 16-byte loads (two keys each) of all of them are in
  // flight together -- one L2 round trip for G <= 24 slices of <= 128 keys, the usual case.  f(in_range_and_above_thr,
  // key) may vote: trip counts are warp-uniform.
  constexpr int kSl = 3, kLd = 2;
  struct Round { ulonglong2 k[kSl][kLd]; };
  auto load_round = [&](int j0, unsigned p0, Round& r) {     // pairs [p0, p0 + kLd * 32) of slices j0, j0 + kWarps, ...
#pragma unroll
    for (int sl = 0; sl < kSl; ++sl) {
      const int jj = j0 + sl * kWarps;
      const unsigned np = jj < G ? (n_slice[jj] + 1u) >> 1 : 0u;
      const ulonglong2* cand = reinterpret_cast<const ulonglong2*>(slices + (long long)jj * kSliceCap);
#pragma unroll
      for (int q = 0; q < kLd; ++q) {
        const unsigned e = p0 + (unsigned)(q * 32 + lane);
        r.k[sl][q] = (e < np) ? __ldcg(cand + e) : make_ulonglong2(0ull, 0ull);
      }
    }
  };
  auto pairs_of_round = [&](int j0) {                        // the longest of the round's slices, in pairs (warp-uniform)
    unsigned np = 0;
#pragma unroll
    for (int sl = 0; sl < kSl; ++sl) {
      const int jj = j0 + sl * kWarps;
      if (jj < G) np = max(np, (n_slice[jj] + 1u) >> 1);
    }
    return np;
  };
  auto consume_round = [&](int j0, unsigned p0, const Round& r, auto f) {
#pragma unroll
    for (int sl = 0; sl < kSl; ++sl) {
      const int jj = j0 + sl * kWarps;
      if (jj >= G) break;                                    // warp-uniform
      const unsigned nc = n_slice[jj], np = (nc + 1u) >> 1;
#pragma unroll
      for (int q = 0; q < kLd; ++q) {
        if (p0 + (unsigned)(q * 32) >= np) break;            // warp-uniform
        const unsigned e = p0 + (unsigned)(q * 32 + lane);
        f(2 * e < nc && (unsigned)(r.k[sl][q].x >> 32) >= thr_final, r.k[sl][q].x);
        f(2 * e + 1 < nc && (unsigned)(r.k[sl][q].y >> 32) >= thr_final, r.k[sl][q].y);
      }
    }
  };
  // skip_first: the round (j0 = warp, p0 = 0) has been taken care of by the caller
  auto for_each_round = [&](bool skip_first, auto f) {
    for (int j0 = tid >> 5; j0 < G; j0 += kWarps * kSl) {
      const unsigned np = pairs_of_round(j0);
      for (unsigned p0 = 0; p0 < np; p0 += (unsigned)(kLd * 32)) {
        if (skip_first && j0 == (tid >> 5) && p0 == 0u) continue;
        Round r;
        load_round(j0, p0, r);
        consume_round(j0, p0, r, f);
      }
    }
  };
  // ---- round trip 2: every warp's first round of keys, and (warp 0) the fine bins of the K-th counted key's super bin
  // -- the same two-level walk as the producers' ----
  Round first;
  load_round(tid >> 5, 0u, first);
  if (tid < 32) {
    unsigned mine = sv.x + sv.y, incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned u = __shfl_down_sync(0xffffffffu, incl, o);
      if (lane + o < 32) incl += u;
    }
    unsigned above = incl - mine;
    int sel = -1;
    unsigned ab = 0;
    if (above < (unsigned)K && incl >= (unsigned)K) {
      if (above + sv.y >= (unsigned)K) { sel = 2 * lane + 1; ab = above; }
      else { sel = 2 * lane; ab = above + sv.y; }
    }
    const unsigned who = __ballot_sync(0xffffffffu, sel >= 0);
    if (who != 0u) {
      const int src = __ffs(who) - 1;
      const int sb = __shfl_sync(0xffffffffu, sel, src);
      const unsigned above_sb = __shfl_sync(0xffffffffu, ab, src);
      const uint2 f = __ldcg(reinterpret_cast<const uint2*>(fhist + sb * 64) + lane);
      mine = f.x + f.y;
      incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += u;
      }
      above = above_sb + incl - mine;
      if (above < (unsigned)K && above + mine >= (unsigned)K)
        s.sh_thr = __float_as_uint((float)(sb * 64 + ((above + f.y >= (unsigned)K) ? 2 * lane + 1 : 2 * lane)) *
                                   (1.0f / (float)kFineBins));
    }
  }
  group_sync();
  thr_final = s.sh_thr;
  dbg_stamp(g.dbg, 6);
  // ---- the keys at or above the threshold -> shared memory ----
  auto take = [&](bool ok, u64 k) { append_if(ok, k, s.keys, &s.cnt, (unsigned)kKeyCap); };
  consume_round(tid >> 5, 0u, first, take);
  for_each_round(true, take);
  group_sync();
  // ---- the rows beyond a tile for those that need them (all in flight together: one more round trip), the fine
  // histogram of what stays, and the lines the boxes will be gathered from started towards L2.  The list is compacted
  // in place, 256 keys per wave (a wave is read before anything at or below it is written).  A list that ran over
  // shared memory (plateaus of ties) is left alone: select_sort_emit then goes through the slices again, slowly. ----
  const unsigned n_taken = s.cnt;
  if (n_taken <= (unsigned)kKeyCap) {
    group_sync();                                            // (everyone has read s.cnt)
    if (tid == 0) s.cnt = 0u;
    for (unsigned i0 = 0; i0 < n_taken; i0 += kThreads) {
      const unsigned i = i0 + (unsigned)tid;
      const u64 k = i < n_taken ? s.keys[i] : 0ull;
      const bool good = i < n_taken && verified(k);
      group_sync();
      append_if(good, k, s.keys, &s.cnt, (unsigned)kKeyCap);
      if (good) {
        hist_add(s.hist, fine_bin((unsigned)(k >> 32)));
        const unsigned pix = (0xffffffffu - (unsigned)(k & 0xffffffffu)) % (unsigned)g.HW;
        if (a.reg) {
          l2_prefetch(a.reg + ((long long)b * 2 + 0) * g.HW + pix);
          l2_prefetch(a.reg + ((long long)b * 2 + 1) * g.HW + pix);
        }
        l2_prefetch(a.wh + ((long long)b * a.D + 0) * g.HW + pix);
        l2_prefetch(a.wh + ((long long)b * a.D + 1) * g.HW + pix);
        if (a.rotated) l2_prefetch(a.wh + ((long long)b * a.D + 2) * g.HW + pix);
      }
    }
  }
  auto for_each_survivor = [&](auto f) { for_each_round(false, [&](bool ok, u64 k) { f(ok && verified(k), k); }); };
  group_sync();
  dbg_stamp(g.dbg, 7);
  select_sort_emit(a, g, s, b, s.keys, (int)s.cnt, kKeyCap,
                   [&](auto f) { for_each_survivor([&](bool ok, u64 k) { f(ok ? k : 0ull); }); });   // convergent: f may vote
  dbg_stamp(g.dbg, 10);
  group_sync();
  cand_state_clear(g, b);
}

// ---- host ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}

static int validate(const cnh_decode_args* a) {
  CNH_REQUIRE(a != nullptr, CNH_E_NULL, "decode: args is NULL");
  CNH_REQUIRE(a->B > 0 && a->C > 0 && a->H > 0 && a->W > 0, CNH_E_SHAPE, "decode: bad dims B=%d C=%d H=%d W=%d",
              a->B, a->C, a->H, a->W);
  CNH_REQUIRE(a->K >= 1 && a->K <= kMaxK, CNH_E_UNSUPPORTED, "decode: K=%d outside [1,%d]", a->K, kMaxK);
  CNH_REQUIRE((long long)a->K <= (long long)a->H * a->W, CNH_E_SHAPE,
              "decode: K=%d > H*W=%lld (torch.topk over H*W would fail, backends/decode.py:19)", a->K,
              (long long)a->H * a->W);
  CNH_REQUIRE((long long)a->C * a->H * a->W < (1ll << 32), CNH_E_SHAPE, "decode: C*H*W does not fit 32 bits");
  CNH_REQUIRE((long long)a->B * a->C < (1ll << 31), CNH_E_SHAPE, "decode: B*C too large");
  CNH_REQUIRE(a->heat && a->wh && a->dets, CNH_E_NULL, "decode: heat/wh/dets is NULL");
  CNH_REQUIRE(a->D >= 2 && (!a->rotated || a->D >= 3), CNH_E_SHAPE, "decode: wh has D=%d channels (rotated=%d)",
              a->D, a->rotated);
  CNH_REQUIRE(a->kps == nullptr || (a->nk > 0 && a->kps_out != nullptr), CNH_E_NULL,
              "decode: kps given without nk/kps_out");
  return CNH_OK;
}

static size_t up128(size_t v) { return (v + 127) / 128 * 128; }

// keys of one sample's candidate list of the two-kernel path (16-row tiling has the most tiles)
static size_t cand_keys_per_sample(const cnh_decode_args* a) {
  const size_t tiles16 = (size_t)a->C * ((a->W + kCols - 1) / kCols) * ((a->H + 15) / 16);
  return tiles16 * (size_t)(a->K + kSlack);
}

static DecGeo make_geo(const cnh_decode_args* a, void* ws, int rows, int stream_ctas_per_sm = 2) {
  DecGeo g;
  g.rows = rows;
  g.HW = a->H * a->W;
  g.tiles_x = (a->W + kCols - 1) / kCols;
  g.tiles_y = (a->H + g.rows - 1) / g.rows;
  g.tiles_per_plane = g.tiles_x * g.tiles_y;
  g.tiles_per_sample = g.tiles_per_plane * a->C;
  g.n_tiles = (long long)a->B * g.tiles_per_sample;
  const int tw = a->W < kCols ? a->W : kCols;
  g.box_w = ((tw + 3) / 4) * 4 + 2 * kPadL;
  g.use_tma = 0;
  g.dbg = debug_buffer();
  g.slot = a->K + kSlack;
  char* p = static_cast<char*>(ws);
  g.state = reinterpret_cast<SampleState*>(p);
  p += up128((size_t)a->B * sizeof(SampleState));
  g.ghist = reinterpret_cast<unsigned*>(p);
  p += up128((size_t)a->B * kCoarseBins * sizeof(unsigned));
  g.cand = reinterpret_cast<u64*>(p);
  // streaming path: G CTAs per sample (two CTAs per SM over the batch, at most one per 32-row tile)
  const int tiles32 = a->C * ((a->H + kStRows - 1) / kStRows);
  int G = (stream_ctas_per_sm * sm_count()) / a->B;        // (the workspace is sized for two CTAs per SM)
  if (G > tiles32) G = tiles32;
  if (G < 1) G = 1;
  g.only_overflow = 0;
  g.verify_rows = 0;
  p += up128((size_t)a->B * cand_keys_per_sample(a) * sizeof(u64));
  g.cl = cand_geo(p, a->B, G);
  return g;
}

static size_t decode_ws_bytes(const cnh_decode_args* a) {
  DecGeo g = make_geo(a, nullptr, 16);
  return up128((size_t)a->B * sizeof(SampleState)) + up128((size_t)a->B * kCoarseBins * sizeof(unsigned)) +
         up128((size_t)a->B * cand_keys_per_sample(a) * sizeof(u64)) + cand_ws_bytes(a->B, g.cl.G);
}

// environment switches of the tests / tools, read once per process... unless CNH_DECODE_ENV_RELOAD is set (the
// parity tests flip the switches between cases inside one process)
struct DecodeEnv { bool rows16, two_kernel, reload; int stream; };
static const DecodeEnv& decode_env() {
  static DecodeEnv e = {false, false, true, -1};
  if (e.reload) {
    const char* rows = getenv("CNH_DECODE_ROWS");
    e.rows16 = rows != nullptr && atoi(rows) == 16;
    e.two_kernel = getenv("CNH_DECODE_TWO_KERNEL") != nullptr;
    const char* st = getenv("CNH_DECODE_STREAM");            // 1 / 0 force the streaming path on / off; unset: by size
    e.stream = st != nullptr ? atoi(st) : -1;
    e.reload = getenv("CNH_DECODE_ENV_RELOAD") != nullptr;
  }
  return e;
}

constexpr int kClusterUnavailable = -999;
constexpr int kClMaxSmem = 227 * 1024;

// co-resident clusters of cs CTAs at one CTA per SM on device `dev` (cached; -1 = unavailable)
static int active_clusters(int dev, int cs) {
  static int max_clusters[64][9] = {};            // 0 = not queried
  if (max_clusters[dev][cs] == 0) {
    cudaLaunchConfig_t q;
    memset(&q, 0, sizeof(q));
    q.gridDim = dim3((unsigned)cs * 64u);
    q.blockDim = dim3(kClThreads);
    q.dynamicSmemBytes = kClMaxSmem - 1024;        // conservative: one CTA per SM
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = (unsigned)cs;
    qa[0].val.clusterDim.y = 1;
    qa[0].val.clusterDim.z = 1;
    q.attrs = qa;
    q.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, decode_cluster_kernel<32>, &q) != cudaSuccess || n < 1) {
      cudaGetLastError();
      n = -1;
    }
    max_clusters[dev][cs] = n;
  }
  return max_clusters[dev][cs];
}

// Largest cluster size <= 8 whose B clusters are co-resident, else 1 (samples then run in waves).
// (On a B200 fifteen 8-CTA clusters fit at one CTA per SM, so a batch of 16 runs as 16 clusters of 7.)
static int pick_cluster_size(int B, int tiles_per_sample, int dev) {
  for (int cs = 8; cs > 1; --cs) {
    if (cs > tiles_per_sample) continue;
    if ((long long)B <= (long long)active_clusters(dev, cs)) return cs;
  }
  return active_clusters(dev, 1) < 1 ? -1 : 1;
}

template <int R>
static int launch_cluster_rows(const cnh_decode_args* a, const DecGeo& g0, int cs, cudaStream_t st, int only_overflow = 0) {
  typedef ClCfg<R> Cfg;
  DecGeo g = g0;
  g.n_stages = Cfg::kStages;
  g.use_tma = 1;
  g.only_overflow = only_overflow;
  const size_t smem = sizeof(ClSmemT<R>) + (size_t)Cfg::kStages * Cfg::kTileFloats * sizeof(float);
  static_assert((size_t)Cfg::kStages * Cfg::kTileFloats * sizeof(float) >= (size_t)8 * kStageCap * sizeof(u64), "the ring holds the inbox");
  static_assert(sizeof(ClSmemT<R>) + (size_t)Cfg::kStages * Cfg::kTileFloats * sizeof(float) <= (size_t)kClMaxSmem, "shared memory");
  static const bool use_pdl = (getenv("CNH_NO_PDL") == nullptr);
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.gridDim = dim3((unsigned)a->B * (unsigned)cs);
  lc.blockDim = dim3(kClThreads);
  lc.dynamicSmemBytes = smem;
  lc.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = use_pdl ? 2 : 1;
  if (cudaLaunchKernelEx(&lc, decode_cluster_kernel<R>, *a, g) != cudaSuccess) {
    cudaGetLastError();                                  // e.g. a partitioned device that cannot co-schedule the
    return kClusterUnavailable;                          // cluster: the caller takes the two-kernel path
  }
  return CNH_OK;
}

static bool cluster_attrs(int dev) {
  static bool attr_set[64] = {false};
  if (dev < 0 || dev >= 64) return false;
  if (!attr_set[dev]) {
    if (cudaFuncSetAttribute(decode_cluster_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kClMaxSmem) != cudaSuccess ||
        cudaFuncSetAttribute(decode_cluster_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kClMaxSmem) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    attr_set[dev] = true;
  }
  return true;
}

// 32-row tiles by default.  The 16-row / four-round shape exists for experiments (CNH_DECODE_ROWS=16) and is
// covered by the parity tests; measured at the cfg5 shard it is SLOWER (58 vs 44 us): a round is bounded by
// its fixed cost (scan issue + barrier), not by the bytes in flight -- mbarrier waits return at once.
static int launch_cluster(const cnh_decode_args* a, void* workspace, int dev, cudaStream_t st) {
  if (!cluster_attrs(dev)) return kClusterUnavailable;
  DecGeo g32 = make_geo(a, workspace, 32);
  const int cs = pick_cluster_size(a->B, g32.tiles_per_sample, dev);
  if (cs < 1) return kClusterUnavailable;
  const bool rows16 = decode_env().rows16;                  // tests: force one shape
  if (!rows16) return launch_cluster_rows<32>(a, g32, cs, st);
  return launch_cluster_rows<16>(a, make_geo(a, workspace, 16), cs, st);
}

static int launch_finish(const cnh_decode_args* a, const DecGeo& g, int cs, cudaStream_t st);

// Streaming path: stream kernel (scan + candidate lists + thresholds) -> finish kernel (one CTA per sample) -> the
// cluster kernel as the fallback for samples whose buffers ran over (it exits at once for all others).  All three
// carry the programmatic-stream-serialisation attribute: each starts with griddepcontrol.wait.
static int launch_stream(const cnh_decode_args* a, void* workspace, int dev, cudaStream_t st) {
  if (!cluster_attrs(dev)) return kClusterUnavailable;
  static int occ[64] = {0};                                 // resident stream CTAs per SM (0 = not set up yet)
  if (dev < 0 || dev >= 64) return kClusterUnavailable;
  if (occ[dev] == 0) {
    int n = 0;
    if (cudaFuncSetAttribute(decode_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStSmemBytes) != cudaSuccess ||
        cudaFuncSetAttribute(decode_stream_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, decode_stream_kernel, kStThreads, kStSmemBytes) != cudaSuccess || n < 1) {
      cudaGetLastError();
      return kClusterUnavailable;
    }
    occ[dev] = n > 2 ? 2 : n;
  }
  DecGeo g = make_geo(a, workspace, kStRows, occ[dev]);
  const int cs = pick_cluster_size(a->B, g.tiles_per_sample, dev);
  if (cs < 1) return kClusterUnavailable;
  static const bool use_pdl = (getenv("CNH_NO_PDL") == nullptr);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.stream = st;
  lc.attrs = attr;
  lc.numAttrs = use_pdl ? 1 : 0;
  lc.gridDim = dim3((unsigned)a->B * (unsigned)g.cl.G);
  lc.blockDim = dim3(kStThreads);
  lc.dynamicSmemBytes = kStSmemBytes;
  CNH_CUDA(cudaLaunchKernelEx(&lc, decode_stream_kernel, *a, g));
  // (a separate finish kernel with the cluster kernel as a pure fallback BEHIND it ended the step 4 us later, between the
  // two 2-3 us later than this single launch: measured)
  return launch_finish(a, g, cs, st);
}

// ONE cluster launch behind whatever left the candidate lists in g.cl: finishes the usual samples from their lists,
// redoes the overflowed ones from the heat map (decode_cluster_kernel, only_overflow)
static int launch_finish(const cnh_decode_args* a, const DecGeo& g, int cs, cudaStream_t st) {
  const int rc = launch_cluster_rows<32>(a, g, cs, st, 1);
  CNH_REQUIRE(rc != kClusterUnavailable, CNH_E_UNSUPPORTED, "decode: the cluster launch that finishes the candidate lists was refused");
  return rc;
}

static bool want_stream(const cnh_decode_args* a) {
  const int force = decode_env().stream;
  if (a->apply_sigmoid || force == 0) return false;
  if (force == 1) return true;
  // long walks only: at least four 32-row tiles for each of the two CTAs per SM
  const long long tiles = (long long)a->B * a->C * ((a->H + kStRows - 1) / kStRows);
  return tiles >= 8ll * sm_count();
}

}  // namespace cnh

using namespace cnh;

// not part of the public ABI (tools/): cluster size the one-launch path would use * 1000 + co-resident 8-CTA clusters
extern "C" int cnh_debug_decode_cluster(const cnh_decode_args* a) {
  int dev = 0;
  if (validate(a) != CNH_OK || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  if (!cluster_attrs(dev)) return -1;
  DecGeo g = make_geo(a, nullptr, 32);
  const int cs = pick_cluster_size(a->B, g.tiles_per_sample, dev);
  return (cs < 0 ? 0 : cs) * 1000 + active_clusters(dev, 8);
}

// not part of the public ABI (tools/): co-resident clusters of `cs` CTAs of the decode kernel on the current device
extern "C" int cnh_debug_active_clusters(int cs) {
  int dev = 0;
  if (cs < 1 || cs > 8 || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || !cluster_attrs(dev)) return -1;
  return active_clusters(dev, cs);
}

extern "C" int cnh_decode_candidates(const cnh_decode_args* a, const cnh_cand* cand, cnh_stream_t stream) {
  if (int rc = validate(a)) return rc;
  CNH_REQUIRE(cand != nullptr && cand->workspace != nullptr, CNH_E_NULL, "decode_candidates: cand / its workspace is NULL");
  CNH_REQUIRE(cand->G >= 1, CNH_E_UNSUPPORTED, "decode_candidates: the loss launch emitted no candidates (G = 0): use cnh_decode");
  CNH_REQUIRE(cand->B == a->B && cand->C == a->C && cand->H == a->H && cand->W == a->W, CNH_E_SHAPE,
              "decode_candidates: candidates of a %dx%dx%dx%d heat map, decode of %dx%dx%dx%d", cand->B, cand->C, cand->H,
              cand->W, a->B, a->C, a->H, a->W);
  CNH_REQUIRE(a->K <= cand->K, CNH_E_SHAPE, "decode_candidates: K=%d but the candidates were pruned for K=%d", a->K, cand->K);
  CNH_REQUIRE(!a->apply_sigmoid, CNH_E_UNSUPPORTED, "decode_candidates: heat must be the probability map the loss launch wrote");
  CNH_REQUIRE(a->W <= kCols && a->W % 4 == 0 && aligned16(a->heat), CNH_E_UNSUPPORTED, "decode_candidates: unsupported heat map layout");
  CNH_REQUIRE(cand_ws_bytes(cand->B, cand->G) <= cand->workspace_bytes, CNH_E_WORKSPACE, "decode_candidates: candidate workspace too small");
  int dev = 0;
  CNH_CUDA(cudaGetDevice(&dev));
  CNH_REQUIRE(cluster_attrs(dev), CNH_E_UNSUPPORTED, "decode_candidates: kernel attributes refused");
  DecGeo g = make_geo(a, nullptr, kStRows);
  const int cs = pick_cluster_size(a->B, g.tiles_per_sample, dev);
  CNH_REQUIRE(cs >= 1, CNH_E_UNSUPPORTED, "decode_candidates: no cluster launch possible on this device");
  g.cl = cand_geo(cand->workspace, cand->B, cand->G);
  g.verify_rows = kCandRows;
  return launch_finish(a, g, cs, static_cast<cudaStream_t>(stream));
}

extern "C" size_t cnh_cand_state_bytes(const cnh_cand* cand) {
  if (cand == nullptr || cand->B < 1 || cand->G < 1) return 0;
  const CandGeo c = cand_geo(nullptr, cand->B, cand->G);
  return (size_t)(reinterpret_cast<const char*>(c.slices) - static_cast<const char*>(nullptr));
}

extern "C" size_t cnh_decode_workspace_bytes(const cnh_decode_args* a) {
  if (validate(a) != CNH_OK) return 0;
  return decode_ws_bytes(a);
}

extern "C" int cnh_decode(const cnh_decode_args* a, void* workspace, size_t workspace_bytes, cnh_stream_t stream) {
  if (int rc = validate(a)) return rc;
  CNH_REQUIRE(workspace != nullptr && workspace_bytes >= cnh_decode_workspace_bytes(a), CNH_E_WORKSPACE,
              "decode: workspace %zu < %zu bytes", workspace_bytes, cnh_decode_workspace_bytes(a));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EncodeTiledFn enc = encode_fn();
  const bool tma_ok = enc != nullptr && a->W % 4 == 0 && aligned16(a->heat);
  int dev = 0;
  CNH_CUDA(cudaGetDevice(&dev));
  // ---- cluster path (one launch, no global scratch) when a tile's rows are contiguous and 16-byte aligned ----
  if (a->W <= kCols && a->W % 4 == 0 && aligned16(a->heat) && !decode_env().two_kernel) {
    if (want_stream(a)) {                                    // long walks: streaming kernel + finish kernel (+ fallback)
      const int rc = launch_stream(a, workspace, dev, st);
      if (rc != kClusterUnavailable) return rc;
    }
    const int rc = launch_cluster(a, workspace, dev, st);
    if (rc != kClusterUnavailable) return rc;
  }
  // ---- configuration: 32-row tiles + single buffer when every CTA gets at most one tile (small
  // problems: latency matters), else 16-row tiles walked by persistent CTAs with a 4-deep TMA ring.
  struct Cfg { int rows, stages; const void* kernel; size_t smem; int ctas_per_sm; };
  static Cfg cfgs[2] = {
      {32, 1, (const void*)decode_tiles_kernel<32>, sizeof(DecSmemT<32 * kCols>) + 1 * sizeof(float) * tile_floats(32), 0},
      {16, 4, (const void*)decode_tiles_kernel<16>, sizeof(DecSmemT<16 * kCols>) + 4 * sizeof(float) * tile_floats(16), 0}};
  static bool attr_set[64] = {false};
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    for (Cfg& c : cfgs) CNH_CUDA(cudaFuncSetAttribute(c.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
    CNH_CUDA(cudaFuncSetAttribute(decode_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MergeSmem)));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  for (Cfg& c : cfgs)
    if (c.ctas_per_sm == 0) {
      int n = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, c.kernel, kThreads, c.smem) != cudaSuccess || n < 1) {
        cudaGetLastError();
        n = 1;
      }
      c.ctas_per_sm = n;
    }
  const int sms = sm_count();
  const long long tiles32 = (long long)a->B * a->C * ((a->H + 31) / 32) * ((a->W + kCols - 1) / kCols);
  const Cfg& cfg = (tiles32 <= (long long)cfgs[0].ctas_per_sm * sms) ? cfgs[0] : cfgs[1];
  DecGeo g = make_geo(a, workspace, cfg.rows);
  g.n_stages = cfg.stages;
  CNH_REQUIRE(g.n_tiles < (1ll << 31), CNH_E_SHAPE, "decode: too many tiles");
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (tma_ok) {
    const cuuint64_t dims[3] = {(cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->B * (cuuint64_t)a->C};
    const cuuint64_t strides[2] = {(cuuint64_t)a->W * 4, (cuuint64_t)a->W * (cuuint64_t)a->H * 4};
    const cuuint32_t box[3] = {(cuuint32_t)g.box_w, (cuuint32_t)(cfg.rows + 2), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(a->heat), dims, strides,
                           box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    g.use_tma = (r == CUDA_SUCCESS) ? 1 : 0;
  }
  long long grid = (long long)cfg.ctas_per_sm * sms;          // persistent: every CTA walks a tile range
  if (grid > g.n_tiles) grid = g.n_tiles;
  {
    void* params[3] = {&tmap, const_cast<cnh_decode_args*>(a), &g};
    CNH_CUDA(cudaLaunchKernel(cfg.kernel, dim3((unsigned)grid), dim3(kThreads), params, cfg.smem, st));
  }
  // merge: one CTA per sample; programmatic dependent launch lets its prologue overlap the tile kernel's tail
  static const bool use_pdl = (getenv("CNH_NO_PDL") == nullptr);
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.gridDim = dim3((unsigned)a->B);
  lc.blockDim = dim3(kThreads);
  lc.dynamicSmemBytes = sizeof(MergeSmem);
  lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = use_pdl ? 1 : 0;
  CNH_CUDA(cudaLaunchKernelEx(&lc, decode_merge_kernel, *a, g));
  CNH_CUDA(cudaGetLastError());
  return CNH_OK;
}
