// Detection decode for sm_100a in ONE launch (replaces backends/decode.py:6-76: max_pool2d
// + 4 element-wise passes + two torch.topk sorts + 3 full-map transpose copies).
//
// Stage 1 (every CTA = one tile of 32 rows x <=128 columns of one class plane):
//   * the tile plus a 1-pixel halo is staged in shared memory by ONE TMA bulk-tensor copy
//     (cp.async.bulk.tensor.3d, SASS UTMALDG) from a [B*C, H, W] tensor map; out-of-range
//     rows/columns are zero-filled by the TMA unit, which equals max-pool's -inf padding
//     because heat >= 0.  (W % 4 != 0 or a misaligned base falls back to guarded loads.)
//   * while the copy is in flight the CTA reads the sample's GLOBAL score histogram (1024
//     linear bins over [0,1], filled by the tiles that already finished) and derives a
//     threshold below which no peak can reach the sample's top K any more.
//   * 3x3 peak test with a rolling 3-row window in registers: one LDS.128 per row per
//     lane, left/right neighbours by warp shuffle.
//   * peaks >= threshold are compacted to 64-bit keys (score_bits << 32) | ~flat_index --
//     descending key order == score descending, ties to the LOWER flat index c*HW+y*W+x --
//     and counted in a LOCAL 4096-bin histogram.  If the tile holds more than K of them, a
//     block suffix scan over the histogram finds the bin of its K-th score and only keys
//     in bins >= it are forwarded (a few more than K; exact radix select only if a bin is
//     overfull, i.e. massive ties).  The tile adds its counts to the global histogram.
// Stage 2 (the last tile CTA of each sample, elected by an atomic ticket): final threshold
//   from the complete global histogram, survivors of all tiles into shared memory, the same
//   histogram selection, rank sort of the <= K+few keys, zero-score filler when the sample
//   has fewer than K peaks (ascending flat index, as a stable sort would), gather of
//   reg / wh / angle / keypoints straight from NCHW, box assembly.
// No full sort, no transposes: heat is read from HBM exactly once (4*C*H*W bytes/sample).
#include <cuda.h>
#include <string.h>

#include "common.cuh"

namespace cnh {

typedef unsigned long long u64;

constexpr int kRows = 32;                 // tile rows
constexpr int kCols = 128;                // max tile columns
constexpr int kPadL = 4;                  // left halo padded to 4 floats: interior is 16B aligned
constexpr int kBoxWMax = kCols + 2 * kPadL;
constexpr int kTileFloats = (kRows + 2) * kBoxWMax;
constexpr int kKeyCap = kRows * kCols;    // worst case: every pixel of the tile is a peak
constexpr int kMaxK = 1024;
constexpr int kFineBins = 4096;           // local histogram: bin = min(4095, int(score * 4096))
constexpr int kCoarseBins = 1024;         // per-sample global histogram: fine bin >> 2
constexpr int kSlack = 64;                // a tile forwards at most K + kSlack keys

struct DecGeo {
  int HW, tiles_x, tiles_y, tiles_per_plane, tiles_per_sample, box_w, use_tma, slot;
  unsigned* tiles_done;                   // [B]                        zero between launches
  unsigned* ghist;                        // [B][kCoarseBins]           zero between launches
  unsigned* tile_cnt;                     // [B][tiles_per_sample]      rewritten by every launch
  u64* cand;                              // [B][tiles_per_sample][slot]
  long long* dbg;
};

struct __align__(128) DecSmem {
  float tile[kTileFloats];                // stage 2 reuses it: sel = [0,kMaxK), sorted = [kMaxK,2*kMaxK)
  u64 keys[kKeyCap];
  unsigned hist[kFineBins / 2];           // 16-bit counters packed in pairs; radix select uses [0,256)
  u64 mbar;
  u64 sh_prefix;
  unsigned cnt;
  unsigned cnt2;
  unsigned sh_need;
  unsigned sh_flag;
  unsigned sh_thr;
  unsigned sh_bin;
  unsigned sh_above;
  unsigned sh_inbin;
  unsigned warp_tot[kWarps];
};

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
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, u64* bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

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
__device__ void find_kth_bin(DecSmem& s, unsigned need) {
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
__device__ unsigned global_threshold(const unsigned* ghist, unsigned K, DecSmem& s) {
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
template <class ForEach>
__device__ u64 radix_select_kth(ForEach for_each, int need, DecSmem& s) {
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
__device__ bool is_candidate_global(const cnh_decode_args& a, int b, long long flat, int HW) {
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

__global__ void __launch_bounds__(kThreads)
decode_kernel(const __grid_constant__ CUtensorMap tmap, const cnh_decode_args a, const DecGeo g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  DecSmem& s = *reinterpret_cast<DecSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = a.K;

  // ---- which tile ------------------------------------------------------------------------
  const int t = blockIdx.x;
  const int b = t / g.tiles_per_sample;
  const int ts = t - b * g.tiles_per_sample;
  const int c = ts / g.tiles_per_plane;
  const int tp = ts - c * g.tiles_per_plane;
  const int ty = tp / g.tiles_x, tx = tp - ty * g.tiles_x;
  const int y0 = ty * kRows, x0 = tx * kCols;
  const int rows = min(kRows, a.H - y0), cols = min(kCols, a.W - x0);
  const int plane = b * a.C + c;
  const int BW = g.box_w;
  unsigned* ghist = g.ghist + (long long)b * kCoarseBins;

  dbg_stamp(g.dbg, 0);
  if (tid == 0) {
    s.cnt = 0;
    s.cnt2 = 0;
    if (g.use_tma) {
      mbar_init(&s.mbar, 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&s.mbar, (unsigned)(BW * (kRows + 2) * sizeof(float)));
      tma_load_3d(s.tile, &tmap, &s.mbar, x0 - kPadL, y0 - 1, plane);
    }
  }
  // while the tile is in flight: clear the local histogram, derive the pruning threshold from what
  // the finished tiles of this sample have published
#pragma unroll
  for (int j = 0; j < kFineBins / 2 / kThreads; ++j) s.hist[tid + j * kThreads] = 0u;
  const unsigned thr = global_threshold(ghist, (unsigned)K, s);     // contains barriers
  dbg_stamp(g.dbg, 1);
  if (g.use_tma) {
    mbar_wait(&s.mbar, 0);
  } else {
    const float* src = a.heat + (long long)plane * g.HW;
    for (int i = tid; i < (kRows + 2) * BW; i += kThreads) {
      const int r = i / BW, cc = i - r * BW;
      const int gy = y0 - 1 + r, gx = x0 - kPadL + cc;
      s.tile[i] = (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) ? __ldcs(src + (long long)gy * a.W + gx) : 0.f;
    }
    __syncthreads();
  }
  if (a.apply_sigmoid) {                  // export.py:31-33: logits in, clamp(sigmoid) fused
    for (int i = tid; i < (kRows + 2) * BW; i += kThreads) {
      const int r = i / BW, cc = i - r * BW;
      const int gy = y0 - 1 + r, gx = x0 - kPadL + cc;
      const bool in = (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W);
      s.tile[i] = in ? clamp_prob(1.0f / (1.0f + expf(-s.tile[i]))) : 0.f;
    }
    __syncthreads();
  }

  dbg_stamp(g.dbg, 2);
  // ---- 3x3 peaks: warp w owns rows [4w, 4w+4), lane owns columns [4*lane, 4*lane+4) -------
  constexpr int kRowsPerWarp = kRows / kWarps;
  {
    const int r_begin = warp * kRowsPerWarp;
    float hm[3][4];                        // horizontal 3-max of rows r-1, r, r+1
    float cv[kRowsPerWarp][4];             // centre values of the warp's rows
    float4 ctr = make_float4(0.f, 0.f, 0.f, 0.f), nxt = ctr;
    auto load_row = [&](int tr, float (&h)[4], float4& centre) {   // tr: tile row incl. halo
      const float* row = s.tile + tr * BW + kPadL;
      const float4 v = *reinterpret_cast<const float4*>(row + 4 * lane);
      float left = __shfl_up_sync(0xffffffffu, v.w, 1);
      float right = __shfl_down_sync(0xffffffffu, v.x, 1);
      if (lane == 0) left = row[-1];
      if (lane == 31) right = row[4 * 32];
      h[0] = fmaxf(fmaxf(left, v.x), v.y);
      h[1] = fmaxf(fmaxf(v.x, v.y), v.z);
      h[2] = fmaxf(fmaxf(v.y, v.z), v.w);
      h[3] = fmaxf(fmaxf(v.z, v.w), right);
      centre = v;
    };
    float4 dummy;
    load_row(r_begin + 0, hm[0], dummy);       // halo row above
    load_row(r_begin + 1, hm[1], ctr);
    unsigned flags = 0;                        // bit 4*rr + e
#pragma unroll
    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
      const int r = r_begin + rr;
      load_row(r + 2, hm[2], nxt);
      cv[rr][0] = ctr.x; cv[rr][1] = ctr.y; cv[rr][2] = ctr.z; cv[rr][3] = ctr.w;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float m = fmaxf(fmaxf(hm[0][e], hm[1][e]), hm[2][e]);
        const bool ok = (r < rows) && (4 * lane + e < cols) && (cv[rr][e] == m) && (cv[rr][e] > 0.f) &&
                        (__float_as_uint(cv[rr][e]) >= thr);
        flags |= ok ? (1u << (4 * rr + e)) : 0u;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) { hm[0][e] = hm[1][e]; hm[1][e] = hm[2][e]; }
      ctr = nxt;
    }
    // one warp-aggregated append for the warp's 4 x 128 pixels
    const int mine = __popc(flags);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total) {
      unsigned base = 0;
      if (lane == 31) base = atomicAdd(&s.cnt, (unsigned)total);
      base = __shfl_sync(0xffffffffu, base, 31);
      unsigned pos = base + (unsigned)(incl - mine);
#pragma unroll
      for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        const unsigned flat0 = (unsigned)c * (unsigned)g.HW + (unsigned)(y0 + r_begin + rr) * (unsigned)a.W +
                               (unsigned)(x0 + 4 * lane);
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
  __syncthreads();
  const int n = (int)s.cnt;
  dbg_stamp(g.dbg, 3);

  // ---- publish this tile's counts to the sample's global histogram (fire-and-forget REDs) -------
  if (n > 0) {
    unsigned cc[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const unsigned w = s.hist[8 * tid + j];
      cc[j >> 1] += (w & 0xffffu) + (w >> 16);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (cc[j]) atomicAdd(&ghist[4 * tid + j], cc[j]);
  }

  // ---- forward the tile's best keys into its slot of the sample's candidate list ------------------
  u64* slot = g.cand + ((long long)b * g.tiles_per_sample + ts) * g.slot;
  unsigned forwarded = (unsigned)n;
  if (n > K) {
    find_kth_bin(s, (unsigned)K);
    const unsigned keep_n = s.sh_above + s.sh_inbin;
    if (keep_n <= (unsigned)g.slot) {
      const int tbin = (int)s.sh_bin;
      for (int i0 = 0; i0 < n; i0 += kThreads) {
        const int i = i0 + tid;
        const u64 k = (i < n) ? s.keys[i] : 0ull;
        append_if((i < n) && fine_bin((unsigned)(k >> 32)) >= tbin, k, slot, &s.cnt2, (unsigned)g.slot);
      }
      forwarded = keep_n;
    } else {                               // massive ties inside one bin: exact selection
      const u64* keys = s.keys;
      const u64 T = radix_select_kth([&](auto f) { for (int i = tid; i < n; i += kThreads) f(keys[i]); }, K, s);
      for (int i0 = 0; i0 < n; i0 += kThreads) {
        const int i = i0 + tid;
        const u64 k = (i < n) ? s.keys[i] : 0ull;
        append_if((i < n) && k >= T, k, slot, &s.cnt2, (unsigned)g.slot);
      }
      forwarded = (unsigned)K;
    }
  } else {
    for (int i = tid; i < n; i += kThreads) slot[i] = s.keys[i];
  }

  dbg_stamp(g.dbg, 4);
  // ---- elect the last tile of this sample ---------------------------------------------------
  __syncthreads();                         // the CTA's slot writes and histogram REDs are issued ...
  if (tid == 0) {
    g.tile_cnt[(long long)b * g.tiles_per_sample + ts] = forwarded;
    __threadfence();                       // ... and ordered before the ticket (cumulative fence)
    s.sh_flag = (atomicAdd(&g.tiles_done[b], 1u) == (unsigned)(g.tiles_per_sample - 1)) ? 1u : 0u;
  }
  __syncthreads();
  dbg_stamp(g.dbg, 5);
  if (!s.sh_flag) return;
  __threadfence();

  // ---- stage 2: merge ------------------------------------------------------------------------
  static_assert(sizeof(float) * kTileFloats >= 2 * kMaxK * sizeof(u64), "stage-2 buffers alias the tile");
  u64* const sel = reinterpret_cast<u64*>(s.tile);
  u64* const sorted = sel + kMaxK;
  unsigned* const sh_cnt = reinterpret_cast<unsigned*>(sorted + kMaxK);      // per-tile counts, if they fit
  constexpr int kCntCap = (int)((sizeof(float) * kTileFloats - 2 * kMaxK * sizeof(u64)) / sizeof(unsigned));
  const unsigned* tcnt = g.tile_cnt + (long long)b * g.tiles_per_sample;
  const u64* cand = g.cand + (long long)b * g.tiles_per_sample * g.slot;
  const bool cnt_in_smem = g.tiles_per_sample <= kCntCap;
#pragma unroll
  for (int j = 0; j < kFineBins / 2 / kThreads; ++j) s.hist[tid + j * kThreads] = 0u;
  if (tid == 0) { s.cnt = 0; s.cnt2 = 0; }
  if (cnt_in_smem)                                  // same round trip as the histogram read below
    for (int i = tid; i < g.tiles_per_sample; i += kThreads) sh_cnt[i] = __ldcg(tcnt + i);
  const unsigned thr_final = global_threshold(ghist, (unsigned)K, s);        // barriers inside
  dbg_stamp(g.dbg, 6);
  // survivors (score >= final threshold) of every tile -> shared memory keys + fine histogram.
  // One warp per tile, 8 keys per lane per batch, the next batch in flight while this one is used.
  auto for_each_survivor = [&](auto f) {
    constexpr int kB = 8;
    const int batches_per_tile = (g.slot + 32 * kB - 1) / (32 * kB);
    const int n_batches = ((g.tiles_per_sample - warp + kWarps - 1) / kWarps) * batches_per_tile;   // this warp's
    u64 cur[kB], nxt[kB];
    unsigned cur_n = 0, nxt_n = 0;
    auto issue = [&](int q, u64 (&k)[kB], unsigned& valid) {
      const int tile = warp + kWarps * (q / batches_per_tile);
      const unsigned i0 = (unsigned)(q % batches_per_tile) * 32u * kB;
      const unsigned nt = cnt_in_smem ? sh_cnt[tile] : __ldcg(tcnt + tile);
      valid = nt > i0 ? nt - i0 : 0u;
      const u64* src = cand + (long long)tile * g.slot + i0;
#pragma unroll
      for (int j = 0; j < kB; ++j) {
        const unsigned i = (unsigned)lane + 32u * j;
        k[j] = (i < valid) ? __ldcg(src + i) : 0ull;
      }
    };
    if (n_batches > 0) issue(0, cur, cur_n);
    for (int q = 0; q < n_batches; ++q) {
      if (q + 1 < n_batches) issue(q + 1, nxt, nxt_n);
      if (cur_n)
#pragma unroll
        for (int j = 0; j < kB; ++j) {
          if (32u * j >= cur_n) break;                         // warp-uniform
          const unsigned i = (unsigned)lane + 32u * j;
          f(i < cur_n && (unsigned)(cur[j] >> 32) >= thr_final, cur[j]);
        }
#pragma unroll
      for (int j = 0; j < kB; ++j) cur[j] = nxt[j];
      cur_n = nxt_n;
    }
  };
  for_each_survivor([&](bool ok, u64 k) {
    append_if(ok, k, s.keys, &s.cnt, (unsigned)kKeyCap);
    if (ok) hist_add(s.hist, fine_bin((unsigned)(k >> 32)));
  });
  __syncthreads();
  dbg_stamp(g.dbg, 7);
  const int m = (int)s.cnt;                  // survivors (may exceed kKeyCap: then s.keys is partial)
  dbg_stamp(g.dbg, 11);
  int got = 0;                               // keys to sort; the first min(got, K) ranks are real detections
  const u64* sort_src = sel;
  if (m <= kMaxK) {
    sort_src = s.keys;                       // few enough: rank-sort the survivors directly
    got = m;
  } else if (m <= kKeyCap) {
    const u64* keys = s.keys;
    find_kth_bin(s, (unsigned)K);
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
    // candidate slots in global memory
    const u64 T = radix_select_kth([&](auto f) { for_each_survivor([&](bool ok, u64 k) { if (ok) f(k); }); }, K, s);
    for_each_survivor([&](bool ok, u64 k) { append_if(ok && k >= T, k, sel, &s.cnt2, (unsigned)kMaxK); });
    got = K;
  }
  __syncthreads();
  dbg_stamp(g.dbg, 8);
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
  dbg_stamp(g.dbg, 14);
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
  if (g.dbg && tid == 0) { g.dbg[(long long)blockIdx.x * 16 + 12] = m; g.dbg[(long long)blockIdx.x * 16 + 13] = got; }
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
#pragma unroll
  for (int j = 0; j < kCoarseBins / kThreads; ++j) ghist[tid + j * kThreads] = 0u;
  if (tid == 0) g.tiles_done[b] = 0u;
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

static DecGeo make_geo(const cnh_decode_args* a, void* ws) {
  DecGeo g;
  g.HW = a->H * a->W;
  g.tiles_x = (a->W + kCols - 1) / kCols;
  g.tiles_y = (a->H + kRows - 1) / kRows;
  g.tiles_per_plane = g.tiles_x * g.tiles_y;
  g.tiles_per_sample = g.tiles_per_plane * a->C;
  const int tw = a->W < kCols ? a->W : kCols;
  g.box_w = ((tw + 3) / 4) * 4 + 2 * kPadL;
  g.use_tma = 0;
  g.dbg = debug_buffer();
  g.slot = a->K + kSlack;
  char* p = static_cast<char*>(ws);
  g.tiles_done = reinterpret_cast<unsigned*>(p);
  p += up128((size_t)a->B * sizeof(unsigned));
  g.ghist = reinterpret_cast<unsigned*>(p);
  p += up128((size_t)a->B * kCoarseBins * sizeof(unsigned));
  g.tile_cnt = reinterpret_cast<unsigned*>(p);
  p += up128((size_t)a->B * g.tiles_per_sample * sizeof(unsigned));
  g.cand = reinterpret_cast<u64*>(p);
  return g;
}

static size_t decode_ws_bytes(const cnh_decode_args* a) {
  DecGeo g = make_geo(a, nullptr);
  return up128((size_t)a->B * sizeof(unsigned)) + up128((size_t)a->B * kCoarseBins * sizeof(unsigned)) +
         up128((size_t)a->B * g.tiles_per_sample * sizeof(unsigned)) +
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
  DecGeo g = make_geo(a, workspace);
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  EncodeTiledFn enc = encode_fn();
  if (enc != nullptr && a->W % 4 == 0 && aligned16(a->heat)) {
    const cuuint64_t dims[3] = {(cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->B * (cuuint64_t)a->C};
    const cuuint64_t strides[2] = {(cuuint64_t)a->W * 4, (cuuint64_t)a->W * (cuuint64_t)a->H * 4};
    const cuuint32_t box[3] = {(cuuint32_t)g.box_w, (cuuint32_t)(kRows + 2), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(a->heat), dims, strides,
                           box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    g.use_tma = (r == CUDA_SUCCESS) ? 1 : 0;
  }
  static bool attr_set[64] = {false};
  int dev = 0;
  CNH_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    CNH_CUDA(cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecSmem)));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const long long tiles = (long long)a->B * g.tiles_per_sample;
  CNH_REQUIRE(tiles < (1ll << 31), CNH_E_SHAPE, "decode: too many tiles");
  decode_kernel<<<(unsigned)tiles, kThreads, sizeof(DecSmem), static_cast<cudaStream_t>(stream)>>>(tmap, *a, g);
  CNH_CUDA(cudaGetLastError());
  return CNH_OK;
}
