/*
 * cnhead.h -- C ABI of libcnhead_sm100.so: hand-written sm_100a CUDA kernels for the
 * CenterNet-UDA per-pixel head path (detection loss fwd+bwd, UDA target-domain losses,
 * detection decode).  This is the drop-in boundary: the Python plugin modules
 * (losses.centernet.DetectionLoss, losses.entropy.EntropyLoss, losses.max_square.
 * MaxSquareLoss, losses.advent.AdventLoss, utils.image.entropy_map,
 * backends.decode.decode_detection) bind exactly these entry points through ctypes.
 * Reference interface each one replaces is cited as file:line relative to the
 * scheckmedia/centernet-uda checkout.
 *
 * Conventions (all entry points)
 *   - plain C: raw DEVICE pointers, int sizes, float parameters; no torch types.
 *   - the caller owns every buffer including the workspace; the library never
 *     allocates, frees or synchronises.  All work is enqueued on `stream`
 *     (a cudaStream_t passed as void*), so calls compose with CUDA graphs.
 *   - tensors are fp32, contiguous, NCHW; indices are int64; masks are uint8.
 *   - return value: 0 = success, < 0 = argument error (CNH_E_*), > 0 = cudaError_t.
 *     cnh_last_error() returns a thread-local message for the last non-zero return.
 *   - workspaces must be zero-filled ONCE by the caller (cudaMemset) and may then be
 *     reused by later calls on the same stream: kernels leave their counters zeroed.
 *     cnh_decode's workspace layout depends on (B,C,H,W,K): reuse it only for calls with
 *     the same dimensions, or zero it again.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef CNHEAD_H_
#define CNHEAD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNH_VERSION 103 /* major*100 + minor */

typedef void* cnh_stream_t; /* cudaStream_t */

enum {
  CNH_OK = 0,
  CNH_E_NULL = -1,       /* required pointer is NULL                      */
  CNH_E_SHAPE = -2,      /* unsupported / inconsistent dimensions         */
  CNH_E_ALIGN = -3,      /* pointer not aligned as required               */
  CNH_E_WORKSPACE = -4,  /* workspace too small                           */
  CNH_E_UNSUPPORTED = -5, /* parameter combination not implemented        */
  CNH_E_PEER = -6         /* an earlier peer exchange timed out (see cnh_peers.status) */
};

/* second term of a regression head: angle handling of channel 2 of a 3-channel size head, or the limb-length
 * consistency term of a keypoint head (its weight is cnh_head.angle_weight in every case) */
enum {
  CNH_ANGLE_NONE = 0,     /* plain L1 on every channel (losses/centernet.py:128-131)          */
  CNH_ANGLE_SIGMOID = 1,  /* |sc(pred) - sc(target)|   (losses/centernet.py:112-126)           */
  CNH_ANGLE_PERIODIC = 2, /* RAPiD periodic L1         (losses/centernet.py:192-223)           */
  CNH_LIMB_SQRT = 3,      /* keypoint head, D = 2*nk: + sum over pairs |sqrt(|pa-pb|^2 + 1e4) - sqrt(|ta-tb|^2 + 1e4)|
                             (KPSL1Loss, losses/centernet.py:153-187; the 1e4 is the reference's literal)       */
  CNH_LIMB_L1 = 4         /* same with |pa-pb|_1 (use_l1, losses/centernet.py:173-175)                            */
};

/* flags */
enum {
  CNH_FLAG_ACCURATE_MATH = 1, /* expf/logf/IEEE divide instead of ex2/lg2/rcp.approx */
  CNH_FLAG_NO_STASH = 2,      /* force the two-pass (pre-count) schedule             */
  CNH_FLAG_DEFER_TOTALS = 4   /* cnh_detloss_fused_peers: see cnh_detloss_peers_finalize */
};

/* One masked gather-L1 regression head (replaces RegL1Loss / PeriodicRegL1Loss /
 * KPSL1Loss main term, losses/centernet.py:98-133, 192-223, 136-151). */
typedef struct cnh_head {
  const float* map;      /* [B,D,H,W] predictions                                    */
  const float* target;   /* [B,M,D]                                                  */
  const uint8_t* mask;   /* [B,M] or, if elementwise_mask, [B,M,D]                   */
  float* grad;           /* [B,D,H,W] out: dLoss/dmap (dense, zero-filled) or NULL   */
  int32_t D;
  int32_t angle_mode;    /* CNH_ANGLE_* (used when D == 3) or CNH_LIMB_* (D = 2*nk)  */
  int32_t elementwise_mask;
  float weight;          /* wh_weight / off_weight / kp_weight                       */
  float angle_weight;    /* angle_weight, or kp_distance_weight for CNH_LIMB_*       */
  int32_t n_pairs;       /* CNH_LIMB_*: rows of `pairs`                              */
  const int32_t* pairs;  /* CNH_LIMB_*: [n_pairs,2] keypoint indices (device), the reference's kps_weight_indices */
} cnh_head;

/* Peak candidates for the decode, emitted by the detection-loss launch itself (nullable member of
 * cnh_detloss_args).  The loss kernels have every clamped probability in shared memory anyway: with `cand` set they
 * also run the 3x3 peak test there and keep, per sample, a superset of the top-K peaks in `workspace`
 * (thresholds from a running histogram; cnh_cand_workspace_bytes(B); zero-filled ONCE by the caller, left clean by
 * cnh_decode_candidates).  cnh_decode_candidates(decode args, cand) then produces the detections of
 * backends/decode.py:35-76 without reading the heat map again: 4*C*H*W bytes per sample and most of a launch less.
 * The loss launch fills G (0: it could not emit -- shape other than W == 128 with H*W % 4096 == 0, misaligned
 * tensors, or the single-wave schedule -- and the caller uses cnh_decode) and the heat map's dimensions. */
typedef struct cnh_cand {
  void* workspace;
  size_t workspace_bytes;
  int32_t K;             /* in: largest K a later cnh_decode_candidates may ask for (1..1024)          */
  int32_t G;             /* out: candidate slices per sample written by the loss launch (0 = none)    */
  int32_t B, C, H, W;    /* out: heat map the candidates belong to                                     */
} cnh_cand;

#define CNH_MAX_HEADS 3
#define CNH_TOTALS 24  /* int64 words, see below */
#define CNH_SCALARS 8  /* floats, see below      */

/* totals (int64[CNH_TOTALS]): exact batch sums as (hi, lo) pairs, word q = hi, word 12+q = lo;
 * value = (hi * 2^32 + lo) * 2^-40 for the fixed-point sums, lo alone for the integer counts.
 *   q = 0 sum(pos_loss + neg_loss)   q = 1 num_pos (count)
 *   q = 2+3h sum|l1| of head h       q = 3+3h angle / limb sum of head h    q = 4+3h sum(mask_expanded) (count)
 * Integer addition is associative: totals of batch shards may simply be added (all-reduce SUM)
 * and give bit-identical scalars to a single launch over the whole batch.
 * scalar block (float[CNH_SCALARS]):
 *   [0] total loss  [1] hm_loss  [2..4] head losses  [5] num_pos  [6] reserved [7] reserved */
typedef struct cnh_detloss_args {
  int32_t B, C, H, W, M;
  int32_t n_heads;
  int32_t flags;
  int32_t B_global;        /* sharded runs: global batch (only documents intent) */
  const float* hm_logits;  /* [B,C,H,W] raw logits (NOT modified)                                */
  const float* hm_gt;      /* [B,C,H,W] gaussian-splat target                                    */
  float* prob;             /* [B,C,H,W] out: clamp(sigmoid(x),1e-4,1-1e-4) (utils/tensor.py:5-7) */
  float* grad_hm;          /* [B,C,H,W] out: dLoss/dlogits, or NULL for forward only             */
  const int64_t* ind;      /* [B,M] flat y*W+x                                                   */
  float hm_weight;
  int32_t _pad;
  cnh_head heads[CNH_MAX_HEADS];
  float* scalars;          /* [CNH_SCALARS] out (fused / finalize)                               */
  int64_t* totals;         /* [CNH_TOTALS] out: this launch's exact sums (nullable for fused)    */
  const double* norm;      /* [4] in (cnh_detloss_main): global num_pos, mask counts per head    */
  double* norm_out;        /* [4] out (cnh_detloss_count): this shard's num_pos, mask counts     */
  cnh_cand* cand;          /* nullable: emit peak candidates for cnh_decode_candidates (see above) */
} cnh_detloss_args;

int cnh_version(void);
const char* cnh_last_error(void);

/* ---- DetectionLoss (losses/centernet.py:7-95,98-133,192-223) ---------------------------
 * cnh_detloss_fused: ONE cooperative launch = sigmoid+clamp, penalty-reduced focal loss,
 * masked gather-L1 heads, all gradients (upstream gradient 1.0), scalars and totals.
 * Heat-map chunks are staged in shared memory by TMA bulk copies.  Problems that fit one wave
 * of CTAs keep the raw heat-map gradient in shared memory across the grid barrier (16 B per
 * heat-map element of HBM traffic); larger ones pre-count num_pos over the target and skip
 * its all-zero 4 KB sub-blocks in the second pass (16 B for sparse targets, <= 20 B). */
size_t cnh_detloss_workspace_bytes(const cnh_detloss_args* a);
/* 1 if cnh_detloss_fused would run this problem as a single wave, else 0 (the pre-count schedule).
 * Needs a current CUDA device. */
int cnh_detloss_single_wave(const cnh_detloss_args* a);
int cnh_detloss_fused(const cnh_detloss_args* a, void* workspace, size_t workspace_bytes,
                      cnh_stream_t stream);
/* Sharded, fused exchange (one process per GPU, peer-mapped mailboxes over NVLink/NVSwitch): the same
 * launch as cnh_detloss_fused; the normalisers (num_pos, mask counts) of every rank are stored into every
 * peer's mailbox and read from the local one INSIDE the kernel -- no collective library call on the step
 * path.  Single wave: the finaliser CTA posts, the chunk CTAs poll.  Larger problems (pre-count schedule):
 * CTA 0 posts after the count phase's grid barrier, every CTA polls before its streaming pass.
 * Every rank must issue the call.  mailbox[r]: device pointer, valid on THIS device, to rank r's mailbox
 * (CNH_MAILBOX_BYTES, zeroed once, symmetric allocation); mailbox[rank] is the local one.  The exchange
 * counter lives in the local mailbox (word CNH_MAILBOX_EPOCH_WORD), so any workspace may be used with it.
 * Waits on peers are BOUNDED: after timeout_ms (0 = 2000) without the peers' words a kernel stores a
 * non-zero code to *status (nullable; PINNED HOST memory, zeroed by the caller: the kernel writes it through
 * its unified address), poisons what it was about to produce (NaN gradient scale / NaN scalars) and
 * terminates; every later peers call that sees *status != 0 returns CNH_E_PEER until the caller has
 * re-created the mailboxes (all ranks) and cleared the word. */
#define CNH_MAX_PEERS 8
#define CNH_MAILBOX_BYTES 8192
#define CNH_MAILBOX_EPOCH_WORD 512 /* uint64 index inside the local mailbox */
typedef struct cnh_peers {
  int32_t world, rank;
  void* mailbox[CNH_MAX_PEERS];
  uint32_t* status;      /* nullable: pinned host word, see above */
  uint32_t timeout_ms;   /* 0 = default (2000)                    */
  uint32_t _pad;
} cnh_peers;
int cnh_detloss_fused_peers(const cnh_detloss_args* a, const cnh_peers* peers, void* workspace,
                            size_t workspace_bytes, cnh_stream_t stream);
/* With CNH_FLAG_DEFER_TOTALS the fused launch only trades the normalisers (what the gradients need),
 * leaves THIS rank's exact totals in a->totals and writes no scalars; this one-warp launch -- any time
 * later, and complete BEFORE the next cnh_detloss_fused_peers on the workspace starts -- posts them to the
 * peers, waits for theirs, sums, and writes the global a->totals / a->scalars.  It may run on another
 * stream next to decode (bench.py does): the NVLink round trip of the totals and the write
 * acknowledgements then cost the step nothing (measured at N = 2: 33.8 -> 27 us). */
int cnh_detloss_peers_finalize(const cnh_detloss_args* a, const cnh_peers* peers, void* workspace,
                               size_t workspace_bytes, cnh_stream_t stream);
/* Sharded (one process per GPU) schedule: count -> all-reduce(norm_out) -> main ->
 * all-reduce(totals) -> finalize.  Gradients are final after cnh_detloss_main.
 * cnh_detloss_count also leaves, in the workspace, one sparsity word per 4096-element chunk of hm_gt
 * (which 4 KB sub-blocks hold anything but zeros); a cnh_detloss_main on the SAME workspace and the
 * SAME hm_gt pointer consumes them (all-zero sub-blocks of the target are not re-read) and clears
 * them.  Do not modify hm_gt between the two calls. */
int cnh_detloss_count(const cnh_detloss_args* a, void* workspace, size_t workspace_bytes,
                      cnh_stream_t stream);
int cnh_detloss_main(const cnh_detloss_args* a, void* workspace, size_t workspace_bytes,
                     cnh_stream_t stream);
/* totals: int64[CNH_TOTALS] on the device (e.g. the all-reduced sum over shards) -> a->scalars.
 * Only weights / D / angle_mode / n_heads of `a` are read. */
int cnh_detloss_finalize(const cnh_detloss_args* a, const int64_t* totals, cnh_stream_t stream);

/* In-place gradient rescale by an upstream gradient that lives on the device:
 * g[i] *= (fa ? *fa : 0) + (fb ? *fb : 0); exits immediately when the factor is 1.0
 * (autograd's grad_output for `loss.backward()`).  Up to 4 tensors per launch. */
typedef struct cnh_scale_args {
  int32_t n_tensors;
  int32_t _pad;
  float* data[4];
  int64_t count[4];
  const float* fa[4];
  const float* fb[4];
} cnh_scale_args;
int cnh_scale_inplace(const cnh_scale_args* a, cnh_stream_t stream);

/* Plumbing for steps captured into CUDA graphs (cnhead/graphed.py): one cudaMemcpyAsync(cudaMemcpyDefault) on
 * `stream` -- the `.cpu()` of uda/base.py:84-88 into PINNED host memory as a copy node of the captured step.  No data
 * is touched; dst/src may be device or page-locked host pointers. */
int cnh_copy_async(void* dst, const void* src, size_t bytes, cnh_stream_t stream);

/* ---- UDA target-domain losses over a channel softmax ---------------------------------
 * mode: 0 entropy (losses/entropy.py:24-25), 1 entropy with eta (losses/entropy.py:18-22),
 *       2 max-squares (losses/max_square.py:6-14).
 * logits [N,C,H,W]; grad (nullable) receives dLoss/dlogits for upstream gradient 1.0;
 * n_total = N of the GLOBAL batch (normaliser); loss_out[0] = this shard's contribution
 * (sum over shards == reference loss on the concatenated batch). */
enum { CNH_SOFTMAX_ENTROPY = 0, CNH_SOFTMAX_ENTROPY_ETA = 1, CNH_SOFTMAX_MAX_SQUARE = 2 };
size_t cnh_softmax_workspace_bytes(int32_t N, int32_t C, int32_t H, int32_t W);
int cnh_softmax_loss(const float* logits, float* grad, float* loss_out, int32_t N, int32_t C,
                     int32_t H, int32_t W, int32_t n_total, int32_t mode, float eta,
                     void* workspace, size_t workspace_bytes, cnh_stream_t stream);
/* utils/image.py:121-124 entropy_map: out = -p*log2(p+1e-30)/log2(C) */
int cnh_entropy_map_fwd(const float* logits, float* out, int32_t N, int32_t C, int32_t H,
                        int32_t W, cnh_stream_t stream);
int cnh_entropy_map_bwd(const float* logits, const float* grad_out, float* grad_in, int32_t N,
                        int32_t C, int32_t H, int32_t W, cnh_stream_t stream);
/* losses/advent.py:10-18: mean BCE-with-logits against a constant label; grad nullable. */
int cnh_bce_const(const float* y, float* grad, float* loss_out, int64_t n, float label,
                  cnh_stream_t stream);

/* ---- decode (backends/decode.py:6-76) ---------------------------------------------------
 * heat [B,C,H,W] probabilities in [0,1]; wh [B,D,H,W] (D = 2, or 3 when rotated);
 * reg [B,2,H,W] or NULL (+0.5); kps [B,2*nk,H,W] or NULL.
 * dets [B,K,6] = (x1,y1,x2,y2,score,class) or, rotated, [B,K,7] = (x,y,w,h,angle,score,class),
 * sorted by score descending, ties broken by LOWER flat index c*H*W + y*W + x.
 * inds_out (nullable) [B,K] int64 receives that flat index; kps_out (nullable) [B,K,nk,2].
 * Two launch shapes, chosen by the library: when a tile's rows are contiguous and 16-byte aligned
 * (W <= 128, W % 4 == 0) ONE launch of thread-block clusters (a cluster per sample, candidates exchanged
 * through distributed shared memory, no global scratch); otherwise a persistent tile kernel + a merge
 * kernel that use the workspace. */
typedef struct cnh_decode_args {
  int32_t B, C, H, W, K, D, nk, rotated;
  const float* heat;
  const float* wh;
  const float* reg;
  const float* kps;
  float* dets;
  int64_t* inds_out;
  float* kps_out;
  int32_t apply_sigmoid;   /* 1: heat holds raw logits; clamp(sigmoid) fused in (export.py:31-33) */
  float box_scale;         /* multiply the 4 box columns (uda/base.py:90 down_ratio); 1.0 = off   */
  int32_t* counts_out;     /* nullable [B]: detections with score >= score_threshold -- a prefix of the
                              sorted rows (the filter of evaluation/coco.py:266-267, folded in)     */
  float score_threshold;
  int32_t _pad;
} cnh_decode_args;
size_t cnh_decode_workspace_bytes(const cnh_decode_args* a);
int cnh_decode(const cnh_decode_args* a, void* workspace, size_t workspace_bytes,
               cnh_stream_t stream);
/* Decode from the candidates a detection-loss launch left (cnh_cand): a->heat must be the probability map that
 * launch wrote (a few dozen candidates are re-checked against it; samples whose candidate buffers ran over --
 * plateaus of thousands of equal scores -- are redone from it exactly), a->B/C/H/W must match cand, a->K <= cand->K,
 * apply_sigmoid must be 0.  Same outputs, bit for bit, as cnh_decode.  cand->G == 0: CNH_E_UNSUPPORTED. */
size_t cnh_cand_workspace_bytes(int32_t B);
int cnh_decode_candidates(const cnh_decode_args* a, const cnh_cand* cand, cnh_stream_t stream);
/* cnh_decode_candidates leaves the workspace clean for the next emitting loss launch.  If the candidates of a loss
 * launch are NOT decoded, the caller must zero the first cnh_cand_state_bytes(cand) bytes of the workspace (counters
 * and histograms; the key slices behind them need no clearing) before the next emitting launch. */
size_t cnh_cand_state_bytes(const cnh_cand* cand);

/* ---- target rasteriser (datasets/coco.py:168-215, utils/image.py:8-57) -----------
 * Builds the dense targets of DetectionLoss on the device from per-sample object lists.
 * boxes [B,M,4] fp32 (x1,y1,x2,y2) in HEAT-MAP pixels (i.e. after the dataset's resize to the output
 * grid, coco.py:189-196), classes [B,M] int32, n_obj [B] int32 (only the first n_obj[b] slots of sample
 * b are read).  Out: hm [B,C,H,W] (zero-filled, then one gaussian per object max-blended into its class
 * plane, exact 1.0 at the integer centre), wh [B,M,2], reg [B,M,2], ind [B,M] int64, reg_mask [B,M] u8;
 * slots without a valid object are zero, as in the reference.  Arithmetic is the reference's float64
 * sequence; min_overlap is passed as a ratio of integers (7/10: the reference's 0.7 literal). */
typedef struct cnh_raster_args {
  int32_t B, C, H, W, M;
  int32_t min_overlap_num, min_overlap_den;
  int32_t _pad;
  const float* boxes;
  const int32_t* classes;
  const int32_t* n_obj;
  float* hm;
  float* wh;
  float* reg;
  int64_t* ind;
  uint8_t* reg_mask;
} cnh_raster_args;
int cnh_raster_targets(const cnh_raster_args* a, cnh_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CNHEAD_H_ */
