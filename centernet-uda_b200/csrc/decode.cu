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

#include "common.cuh"

namespace cnh {

constexpr int kCols = 128;                // max tile columns
constexpr int kPadL = 4;                  // left halo padded to 4 floats: interior is 16B aligned
constexpr int kBoxWMax = kCols + 2 * kPadL;
// tile height: 16 rows (deep TMA ring, many tiles per CTA) or 32 rows (one tile per CTA, small problems)
__host__ __device__ constexpr int tile_floats(int rows) { return ((rows + 2) * kBoxWMax + 31) / 32 * 32; }   // 128-byte multiples
constexpr int kMergeKeyCap = 4096;        // survivors the merge kernel can hold in shared memory
constexpr int kMaxK = 1024;
constexpr int kFineBins = 4096;           // local histogram: bin = min(4095, int(score * 4096))
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
};

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
  __syncthreads();                        // warp_tot may still be read from a previous use
  if (lane == 0) warp_tot[warp] = incl;
  __syncthreads();
  unsigned higher = 0;
  total = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    if (w > warp) higher += warp_tot[w];
    total += warp_tot[w];
  }
  return higher + (incl - v);
}

__device__ __forceinline__ int fine_bin(unsigned score_bits) {
  const int b = (int)(__uint_as_float(score_bits) * (float)kFineBins);   // exact: power-of-two scale
  return b < kFineBins - 1 ? b : kFineBins - 1;
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
  __syncthreads();
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
  __syncthreads();
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
    __syncthreads();
    for_each([&](u64 k) {
      if ((k & mask) == prefix) atomicAdd(&s.hist[(unsigned)(k >> shift) & 255u], 1u);
    });
    __syncthreads();
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
    __syncthreads();
    prefix = s.sh_prefix;
    remaining = s.sh_need;
    mask |= (u64)255u << shift;
    const bool done = s.sh_flag != 0u;
    __syncthreads();
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

// ---- stage 2: merge one sample (run by the CTA whose ticket completed it) ---------------------
typedef DecSmemT<kMergeKeyCap> MergeSmem;
__device__ void merge_sample(const cnh_decode_args& a, const DecGeo& g, MergeSmem& s, int b) {
  constexpr int kKeyCap = MergeSmem::kKeyCap;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = a.K;
  unsigned* ghist = g.ghist + (long long)b * kCoarseBins;
  const u64* cand = g.cand + (long long)b * g.tiles_per_sample * g.slot;
  u64* const sel = reinterpret_cast<u64*>(s.hist);
  u64* const sorted = s.stage;
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
  const int m = (int)s.cnt;                  // survivors (may exceed kKeyCap: then s.keys is partial)
  int got = 0;                               // keys to sort; the first min(got, K) ranks are real detections
  const u64* sort_src = sel;
  if (m <= kMaxK) {
    sort_src = s.keys;                       // few enough: rank-sort the survivors directly
    got = m;
  } else if (m <= kKeyCap) {
    const u64* keys = s.keys;
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
    const u64 T = radix_select_kth([&](auto f) { for_each_survivor([&](bool ok, u64 k) { if (ok) f(k); }); }, K, s);
    for_each_survivor([&](bool ok, u64 k) { append_if(ok && k >= T, k, sel, &s.cnt2, (unsigned)kMaxK); });
    got = K;
  }
  __syncthreads();
  dbg_stamp(g.dbg, 8);
  if (got <= kThreads) {
    // <= 256 keys: one key per thread, bitonic sort (descending; padding 0 sorts last).  Exchange
    // distances below 32 are warp shuffles, the rest go through shared memory.
    u64 k = (tid < got) ? sort_src[tid] : 0ull;
    __syncthreads();                                       // sort_src may alias `sorted`'s neighbours: settle reads
    for (int size = 2; size <= kThreads; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        u64 other;
        if (stride >= 32) {
          sorted[tid] = k;
          __syncthreads();
          other = sorted[tid ^ stride];
          __syncthreads();
        } else {
          other = __shfl_xor_sync(0xffffffffu, k, stride);
        }
        const bool up = ((tid & size) == 0);               // this block sorts descending
        const bool lower = ((tid & stride) == 0);
        const bool take_max = (up == lower);
        const u64 mx = k > other ? k : other, mn = k > other ? other : k;
        k = take_max ? mx : mn;
      }
    }
    sorted[tid] = k;
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
  __syncthreads();
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
      __syncthreads();
      unsigned before = 0, all = 0;
      for (int w = 0; w < kWarps; ++w) { if (w < warp) before += s.warp_tot[w]; all += s.warp_tot[w]; }
      const unsigned pos = (unsigned)have + before + (unsigned)__popc(bal & ((1u << lane) - 1u));
      if (fill && pos < (unsigned)K) sorted[pos] = (u64)(0xffffffffu - (unsigned)f);   // score bits 0
      have += (int)all;
      __syncthreads();
    }
  }
  __syncthreads();
  dbg_stamp(g.dbg, 9);
  // ---- gather + box assembly (backends/decode.py:44-74) -----------------------------------------
  const int ncol = a.rotated ? 7 : 6;
  for (int r = tid; r < K; r += kThreads) {
    const u64 k = sorted[r];
    const float score = __uint_as_float((unsigned)(k >> 32));
    const unsigned flat = 0xffffffffu - (unsigned)(k & 0xffffffffu);
    const int cls = (int)(flat / (unsigned)g.HW);
    const int pix = (int)(flat - (unsigned)cls * (unsigned)g.HW);
    const int yy = pix / a.W, xx = pix - yy * a.W;
    float xs = (float)xx, ys = (float)yy;
    if (a.reg) {
      xs += a.reg[((long long)b * 2 + 0) * g.HW + pix];
      ys += a.reg[((long long)b * 2 + 1) * g.HW + pix];
    } else {
      xs += 0.5f;
      ys += 0.5f;
    }
    const float w = a.wh[((long long)b * a.D + 0) * g.HW + pix];
    const float h = a.wh[((long long)b * a.D + 1) * g.HW + pix];
    float* out = a.dets + ((long long)b * K + r) * ncol;
    const float sc = a.box_scale;
    if (!a.rotated) {
      float x1 = xs - w / 2, y1 = ys - h / 2, x2 = xs + w / 2, y2 = ys + h / 2;
      if (sc != 1.0f) { x1 *= sc; y1 *= sc; x2 *= sc; y2 *= sc; }
      out[0] = x1; out[1] = y1; out[2] = x2; out[3] = y2; out[4] = score; out[5] = (float)cls;
    } else {
      const float av = a.wh[((long long)b * a.D + 2) * g.HW + pix];
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
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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

    // ---- threshold-first scan: warp w owns rows {2w, 2w+1}, lane owns columns [4*lane, 4*lane+4) ----
    {
      const float thr_f = __uint_as_float(thr);
      constexpr int kRowsPerWarp = kRows / kWarps;
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
        if (lane == 31) base = atomicAdd(&s.cnt, (unsigned)total);
        base = __shfl_sync(0xffffffffu, base, 31);
        unsigned pos = base + (unsigned)(incl - mine);
#pragma unroll
        for (int rr = 0; rr < kRowsPerWarp; ++rr) {
          const unsigned flat0 = (unsigned)c * (unsigned)g.HW +
                                 (unsigned)(y0 + warp * kRowsPerWarp + rr) * (unsigned)a.W + (unsigned)(x0 + 4 * lane);
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (flags & (1u << (4 * rr + e))) {
              const unsigned bits = __float_as_uint(cv[rr][e]);
              s.keys[pos++] = ((u64)bits << 32) | (u64)(0xffffffffu - (flat0 + e));
              hist_add(s.hist, fine_bin(bits));
            }
        }
      }
    }
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

static DecGeo make_geo(const cnh_decode_args* a, void* ws, int rows) {
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
  return g;
}

static size_t decode_ws_bytes(const cnh_decode_args* a) {
  DecGeo g = make_geo(a, nullptr, 16);     // 16-row tiling has the most tiles: sizes the candidate lists
  return up128((size_t)a->B * sizeof(SampleState)) + up128((size_t)a->B * kCoarseBins * sizeof(unsigned)) +
         (size_t)a->B * g.tiles_per_sample * g.slot * sizeof(u64);
}

}  // namespace cnh

using namespace cnh;

extern "C" size_t cnh_decode_workspace_bytes(const cnh_decode_args* a) {
  if (validate(a) != CNH_OK) return 0;
  return decode_ws_bytes(a);
}

extern "C" int cnh_decode(const cnh_decode_args* a, void* workspace, size_t workspace_bytes, cnh_stream_t stream) {
  if (int rc = validate(a)) return rc;
  CNH_REQUIRE(workspace != nullptr && workspace_bytes >= cnh_decode_workspace_bytes(a), CNH_E_WORKSPACE,
              "decode: workspace %zu < %zu bytes", workspace_bytes, cnh_decode_workspace_bytes(a));
  // ---- configuration: 32-row tiles + single buffer when every CTA gets at most one tile (small
  // problems: latency matters), else 16-row tiles walked by persistent CTAs with a 4-deep TMA ring.
  struct Cfg { int rows, stages; const void* kernel; size_t smem; int ctas_per_sm; };
  static Cfg cfgs[2] = {
      {32, 1, (const void*)decode_tiles_kernel<32>, sizeof(DecSmemT<32 * kCols>) + 1 * sizeof(float) * tile_floats(32), 0},
      {16, 4, (const void*)decode_tiles_kernel<16>, sizeof(DecSmemT<16 * kCols>) + 4 * sizeof(float) * tile_floats(16), 0}};
  static bool attr_set[64] = {false};
  int dev = 0;
  CNH_CUDA(cudaGetDevice(&dev));
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
  EncodeTiledFn enc = encode_fn();
  if (enc != nullptr && a->W % 4 == 0 && aligned16(a->heat)) {
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
  cudaStream_t st = static_cast<cudaStream_t>(stream);
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
