/*
 * rsb200.h -- C ABI of librsb200.so, the B200 (sm_100a) implementation of
 * RecStudio's retriever training-step hot path.
 *
 * The reference (ustcml/RecStudio) is pure Python/PyTorch and has NO native
 * boundary of its own; the entry points below are what a binding for this path
 * would call.  Each one names the reference code it replaces (paths relative
 * to the reference root).  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *  - every function returns 0 on success, a negative RSB200_E* code for an
 *    argument error, or a positive cudaError_t; rsb200_last_error() returns a
 *    thread-local message for the last non-zero return on this thread;
 *  - the CALLER owns every buffer: the library never calls cudaMalloc.  All
 *    pointers are device pointers on the current device unless marked host;
 *  - tables are row-major fp32 with d % 4 == 0 and 16-byte aligned rows;
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*), no
 *    host synchronisation, no global mutable state: calls are re-entrant
 *    (the reference may call plugins from one thread per GPU,
 *    recstudio/utils/data_parallel.py:84-92);
 *  - ids are int64 where the reference produces int64 (batch ids, sampler
 *    output) and int32 inside the library (N < 2^31).
 *  - row 0 of every table is the padding row: it is scored like any other
 *    row but never receives gradient (nn.Embedding(padding_idx=0),
 *    recstudio/model/basemodel/baseretriever.py:84,104).
 */
#ifndef RSB200_H
#define RSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSB200_VERSION 200

/* error codes (negative); positive return values are cudaError_t */
#define RSB200_OK            0
#define RSB200_EINVAL       -1   /* bad argument (null pointer, bad size, misaligned) */
#define RSB200_EUNSUPPORTED -2   /* valid but not implemented shape / mode */
#define RSB200_ENOCUDA      -3   /* no usable CUDA device / wrong architecture */
#define RSB200_EWORKSPACE   -4   /* workspace too small */

/* enums (plain ints in the ABI) */
#define RSB200_LOSS_BPR      0   /* recstudio/model/loss_func.py:50-59  BPRLoss(dns=False) */
#define RSB200_LOSS_SSM      1   /* recstudio/model/loss_func.py:80-90  SampledSoftmaxLoss */
#define RSB200_LOSS_FULL     2   /* recstudio/model/loss_func.py:39-42  SoftmaxLoss (rsb200_pair_loss only: neg_score = all_score) */
#define RSB200_SCORE_IP      0   /* recstudio/model/scorer.py:5-17     InnerProductScorer  */
#define RSB200_SCORE_EUCLID  1   /* recstudio/model/scorer.py:28-34    EuclideanScorer     */

/* phases of rsb200_pair_step (bit mask) */
#define RSB200_PHASE_COUNT   1   /* group touched rows: histogram + slot assignment      */
#define RSB200_PHASE_SCAN    2   /* exclusive scan -> CSR offsets + unique-row list       */
#define RSB200_PHASE_FWD     4   /* fused gather -> score -> loss -> coefficient emit + loss sum */
#define RSB200_PHASE_SCATTER 8   /* segmented (row-grouped) gradient accumulate           */
#define RSB200_PHASE_ALL     15

/* gradient sinks for PHASE_SCATTER */
#define RSB200_SINK_COMPACT  0   /* rows[R] (int64, ascending, unique) + vals[R,d]: a coalesced sparse-COO gradient */
#define RSB200_SINK_DENSE    1   /* vals is a dense [num_rows, d] buffer; touched rows are OVERWRITTEN (accumulate=0) or += (accumulate=1); rows[] still written */
#define RSB200_SINK_APPLY    2   /* no gradient output: the optimizer update (opt_* fields) is applied to the touched rows of w_*_rw; rows[] still written */

/* ------------------------------------------------------------------------- */
int32_t     rsb200_version(void);
/* number of kernels this library has launched in this process (statistics only) */
uint64_t    rsb200_launch_count(void);
const char* rsb200_last_error(void);
/* sm_count / max_threads_per_sm of the CURRENT device (they fix the Philox
 * element<->counter mapping of torch's CUDA generator, see below). */
int32_t     rsb200_device_info(int32_t* sm_count, int32_t* max_threads_per_sm, int32_t* cc_major, int32_t* cc_minor);

/* -------------------------------------------------------------------------
 * S1  UniformSampler.forward        recstudio/ann/sampler.py:86-114
 * S2  PopularSamplerModel.forward   recstudio/ann/sampler.py:243-258
 *
 * Bit-identical to torch.randint(1, num_items, (num_queries, num_neg),
 * device='cuda') / torch.rand(...) -> searchsorted for the same
 * (seed, philox_offset) of the CUDA generator: Philox4x32-10,
 * curand_init(seed, thread, offset), 256-thread blocks,
 * grid = min(sm_count * (max_threads_per_sm/256), ceil(numel/256)), unroll 4
 * (ATen/native/cuda/DistributionTemplates.h:50-89,318-346,485-506).
 * The host must advance the generator by rsb200_philox_counter_offset().
 * Either output pointer may be NULL.  32-bit draw path only
 * (num_items - 1 < 2^28) and numel * 8 < 2^31 (torch's 32-bit-indexing case).
 */
int64_t rsb200_philox_counter_offset(int64_t numel, int32_t sm_count, int32_t max_threads_per_sm);

int32_t rsb200_sample_uniform(uint64_t seed, uint64_t philox_offset,
                              int64_t num_items /* table rows incl. padding row 0 */,
                              int64_t num_queries, int64_t num_neg,
                              int32_t sm_count, int32_t max_threads_per_sm,
                              int64_t* neg_out_i64 /* [num_queries, num_neg] or NULL */,
                              int32_t* neg_out_i32 /* [num_queries, num_neg] or NULL */,
                              void* stream);

/* Same draw with the generator state in DEVICE memory (state_dev[0] = seed, state_dev[1] = philox offset, a multiple of
 * 4): the launch can be captured in a CUDA graph and draws fresh ids on every replay; state_dev[1] is advanced on the
 * device by rsb200_philox_counter_offset(...) after the draw (the host advances torch's generator by the same amount). */
int32_t rsb200_sample_uniform_dev(uint64_t* state_dev, int64_t num_items, int64_t num_queries, int64_t num_neg,
                                  int32_t sm_count, int32_t max_threads_per_sm,
                                  int64_t* neg_out_i64, int32_t* neg_out_i32, void* stream);

/* guide table: guide[k] = searchsorted(table, k / 2^guide_bits), k in [0, 2^guide_bits],
 * turns the reference's log2(N)-step bisection into an O(1)-expected search with
 * identical results.  guide has 2^guide_bits + 1 int32 entries. */
int32_t rsb200_popular_build_guide(const float* table, int64_t num_items, int32_t guide_bits,
                                   int32_t* guide_out, void* stream);

int32_t rsb200_sample_popular(uint64_t seed, uint64_t philox_offset,
                              const float* table /* [num_items] cumsum, sampler.py:240 */,
                              const float* pop_prob /* [num_items], sampler.py:239,241 */,
                              int64_t num_items, int64_t num_queries, int64_t num_neg,
                              int32_t sm_count, int32_t max_threads_per_sm,
                              const int32_t* guide /* or NULL: plain bisection */, int32_t guide_bits,
                              int64_t* neg_out_i64, int32_t* neg_out_i32,
                              float* logq_out /* log(pop_prob[neg]) or NULL */,
                              void* stream);

/* compute_item_p: out[i] = log(pop_prob[ids[i]])   sampler.py:257-258 */
int32_t rsb200_popular_logq(const float* pop_prob, int64_t num_items, const int64_t* ids, int64_t numel,
                            float* out, void* stream);

/* S3  MaskedUniformSampler.forward / uniform_sample_masked_hist   recstudio/ann/sampler.py:117-147,187-214
 * Uniform over the items a user has NOT interacted with: user_hist is [num_users, hist_len] int64,
 * right- or left-padded with 0, hist_len <= 4096.  Draws per_user ids per row (per_user =
 * num_query_per_user * num_neg; the [.., n_q, num_neg] reshape of the reference is a view).
 * Bit-identical to the reference on a CUDA device for the same generator (seed, offset): the seeds are
 * torch.rand(num_users, per_user), the history row is sorted, shifted and searched exactly as
 * sampler.py:137-144 does (ATen upper_bound loop, so duplicate history items behave the same too).
 * num_items = table rows INCLUDING the padding row (the sampler draws from [1, num_items-1] minus history).
 * Workspace: adj_ws int64 [num_users*hist_len], cnt_ws int32 [num_users]. */
int32_t rsb200_masked_workspace_elems(int64_t num_users, int64_t hist_len, int64_t* adj_elems, int64_t* cnt_elems);
int32_t rsb200_sample_uniform_masked(uint64_t seed, uint64_t philox_offset, int64_t num_items,
                                     const int64_t* user_hist, int64_t num_users, int64_t hist_len,
                                     int64_t per_user, int32_t sm_count, int32_t max_threads_per_sm,
                                     int64_t* adj_ws, int32_t* cnt_ws,
                                     int64_t* neg_out_i64, int32_t* neg_out_i32, void* stream);

/* -------------------------------------------------------------------------
 * R1 + E1 + Q1/Q2 + L1/L2 + E2: the fused retriever training step
 *   BaseRetriever.forward (sampler branch) + training_step + loss.backward()
 *   recstudio/model/basemodel/baseretriever.py:142-176,399-404,
 *   recstudio/model/basemodel/recommender.py:638
 *
 * One call = for B interactions (user, pos, neg[n]):
 *   gather W_user[user], W_item[pos], W_item[neg]  (nn.Embedding fwd)
 *   score (IP | Euclid), loss (BPR | SampledSoftmax, mean reduction),
 *   and the gradient of the loss w.r.t. both tables as sparse rows
 *   (what the reference materialises as embedding_dense_backward).
 *
 * All pointers are device pointers.  Workspace arrays are caller-allocated
 * with the sizes given in the comments (rsb200_pair_workspace_sizes fills them).
 */
typedef struct rsb200_pair_args {
    /* tables */
    const float*   w_item;      /* [num_items, d] */
    const float*   w_user;      /* [num_users, d] */
    /* batch */
    const int64_t* user;        /* [B] */
    const int64_t* pos;         /* [B] */
    const int64_t* neg_i64;     /* [B, n] or NULL ... exactly one of neg_i64 / neg_i32 */
    const int32_t* neg_i32;     /* [B, n] or NULL */
    const float*   logq_pos;    /* [B]    or NULL (= 0, UniformSampler)              */
    const float*   logq_neg;    /* [B, n] or NULL (= 0)                               */
    const float*   grad_scale_dev; /* [1] device-resident upstream gradient multiplied into every gradient row
                                      by PHASE_SCATTER (autograd's grad_output), or NULL (= 1)             */
    /* outputs */
    float*         loss;        /* [1]  mean loss                                     */
    float*         pos_score;   /* [B]    or NULL                                     */
    float*         neg_score;   /* [B, n] or NULL                                     */
    int64_t*       item_rows;   /* [cap_item] unique touched rows, ascending          */
    float*         item_vals;   /* COMPACT: [cap_item, d]; DENSE: [num_items, d]      */
    int64_t*       user_rows;   /* [cap_user]                                         */
    float*         user_vals;   /* COMPACT: [cap_user, d]; DENSE: [num_users, d]      */
    uint32_t*      totals;      /* [4] = {item entries, item unique rows, user entries, user unique rows} */
    /* workspace */
    uint32_t*      off_item;    /* [num_items + 1]  histogram -> CSR offsets          */
    uint32_t*      off_user;    /* [num_users + 1]                                    */
    int32_t*       neg32_buf;   /* [B * n] int32 copy of neg_i64 (unused when neg_i32 is given) */
    uint32_t*      slot_neg;    /* [B * n]  COUNT: position inside the row's segment; SCAN turns it into the absolute entry position */
    uint32_t*      slot_pos;    /* [B]                                                */
    uint32_t*      slot_user;   /* [B]                                                */
    uint64_t*      ent_item;    /* [B * (n + 1)]  (query, coefficient) entries by row */
    uint64_t*      ent_user;    /* [B]                                                */
    uint32_t*      urow_item;   /* [cap_item]                                         */
    uint32_t*      urow_user;   /* [cap_user]                                         */
    float*         q_buf;       /* [B, d] gathered queries                            */
    float*         dq_buf;      /* [B, d] d loss / d query                            */
    float*         loss_part;   /* [B]                                                */
    float*         lse;         /* [B]   (SSM)                                        */
    uint64_t*      scan_tmp;    /* [scan_tmp_elems]                                   */
    uint32_t*      err_flag;    /* [1] set non-zero if an id was out of range         */
    /* sizes */
    int64_t num_items, num_users, B, n, d;
    int64_t cap_item, cap_user;      /* capacity of *_rows / urow_* (>= min(touches, rows)) */
    int64_t scan_tmp_elems;
    float   grad_scale;              /* upstream gradient (loss.backward() => 1.0)     */
    int32_t loss_kind, score_kind, sink, accumulate;
    int32_t variant;                 /* 0 = default kernels.  A/B switches kept for the measurements quoted in DESIGN.md 2
                                      * (same results, different memory schedule): 1 software-pipelined forward, 2 TMA
                                      * (cp.async.bulk) ring, 3 L2 prefetch, 4 forward compiled for 4 CTAs/SM, 6 CSR offsets
                                      * looked up inside the forward kernel (no resolve pass), 16..31 L2 eviction-priority
                                      * hints (bit 0 rows evict_first, bit 1 offsets, bit 2 entries evict_last, bit 3 scatter), 7 staged
                                      * entries + permute pass (needs cstage),
                                      * 40..44 scatter occupancy / unroll.  32 = timing diagnostic WITHOUT the entry list:
                                      * the gradients it produces are invalid. */
    /* RSB200_SINK_APPLY ("fuse into the scatter epilogue"): PHASE_SCATTER applies the optimizer update of
     * every touched row straight from the accumulated gradient, so the gradient rows are neither written nor re-read.
     * w_*_rw normally alias w_item / w_user; state arrays have the tables' shapes. */
    float*  w_item_rw;
    float*  w_user_rw;
    float*  item_state1;             /* Adagrad: sum of squares; SparseAdam: exp_avg        */
    float*  item_state2;             /* SparseAdam: exp_avg_sq                              */
    float*  user_state1;
    float*  user_state2;
    int32_t opt_kind;                /* 0 SGD, 1 Adagrad, 2 SparseAdam (semantics of rsb200_rows_update) */
    float   opt_lr, opt_beta1, opt_beta2, opt_eps;
    float   opt_step_size;           /* SparseAdam: lr * sqrt(1 - beta2^t) / (1 - beta1^t)   */
    float*  cstage;                  /* [B * n] or NULL; used by variant 7 only (A/B): PHASE_FWD writes each negative's
                                      * coefficient / logit here in touch order (coalesced) and PHASE_SCATTER starts with a
                                      * permute pass that moves them to their row-grouped entry positions -- instead of one
                                      * random 8-byte write per touch inside the row stream.  Measured at config 2: forward
                                      * 0.786 -> 0.703 ms, permute pass +0.097 ms: the random writes cost the same wherever
                                      * they run, so the direct write stays the default. */
    /* Binned grouping of the touches (grouping = 1; csrc/bins.cu).  Each table is cut into bins of 2^bin_shift consecutive
     * rows; PHASE_COUNT histograms the touches per bin, PHASE_SCAN turns the histograms into list offsets and append
     * cursors, PHASE_FWD appends each touch's entry to its bin's list, and PHASE_SCATTER groups one bin's entries by row
     * inside shared memory and writes every touched row once (deterministic: a row's entries are summed in ascending order
     * of their encoding, whatever order the appends landed in).  off_* / slot_* / urow_* / scan_tmp are then unused and may be
     * NULL.  The bin arrays hold the item table's bins first, then the user table's (nbins_item + nbins_user elements;
     * bin_off has one more per table).  grouping = 0 is the N-bucket counting sort of round 1 (csrc/group.cu + scatter.cu),
     * kept for A/B runs and for shapes the bins do not cover (rsb200_pair_workspace_sizes reports bin_shift = 0 for those). */
    int32_t   grouping;
    int32_t   bin_shift;             /* item table: rows per bin = 2^bin_shift                               */
    int32_t   bin_shift_user;        /* user table                                                           */
    uint32_t* bin_cnt;               /* [nbins_item + nbins_user], nbins = ceil(rows / 2^shift)              */
    uint32_t* bin_off;               /* [nbins_item + 1 + nbins_user + 1]                                    */
    uint32_t* bin_cursor;            /* [(nbins_item + nbins_user) * 8]  one 32-byte sector per bin          */
    uint64_t* bin_status;            /* [nbins_item + nbins_user]  look-back words of the compact sink       */
    uint32_t* bin_ticket;            /* [2]                                                                  */
    uint32_t* bin_heavy;             /* [bin_heavy elements] per-CTA scratch for bins with more than 4096 touches */
} rsb200_pair_args;

/* Fills the size fields a caller needs to allocate the workspace of a
 * (num_items, num_users, B, n, d) problem.  All counts are in ELEMENTS. */
typedef struct rsb200_pair_sizes {
    int64_t off_item, off_user, neg32_buf, slot_neg, slot_pos, slot_user, ent_item, ent_user,
            urow_item, urow_user, q_buf, dq_buf, loss_part, lse, scan_tmp, cap_item, cap_user,
            bin_shift /* suggested rows-per-bin exponent of the item table, 0 = use grouping 0 */, bin_shift_user,
            nbins /* item table */, nbins_user, bin_cnt, bin_off, bin_cursor, bin_status, bin_heavy;
} rsb200_pair_sizes;
int32_t rsb200_pair_workspace_sizes(int64_t num_items, int64_t num_users, int64_t B, int64_t n, int64_t d,
                                    rsb200_pair_sizes* out);

int32_t rsb200_pair_step(const rsb200_pair_args* args, int32_t phases, void* stream);
/* UniformSampler draw fused with RSB200_PHASE_COUNT (grouping 1): the ids torch.randint(1, num_items, (B, n), device='cuda')
 * would return for the generator state -- (seed, philox_offset), or state_dev[0..1] in DEVICE memory when state_dev != NULL
 * (advanced on the device afterwards, as rsb200_sample_uniform_dev) -- are written to neg_out_i32 (which must be args->neg_i32)
 * and histogrammed per bin while they are in registers, together with args->pos / args->user.  Follow it with
 * rsb200_pair_step(args, RSB200_PHASE_SCAN | RSB200_PHASE_FWD | RSB200_PHASE_SCATTER, stream).
 * Replaces UniformSampler.forward (recstudio/ann/sampler.py:86-114) + the grouping pass of the embedding backward. */
int32_t rsb200_pair_draw_count(const rsb200_pair_args* args, uint64_t* state_dev, uint64_t seed, uint64_t philox_offset,
                               int32_t sm_count, int32_t max_threads_per_sm, int32_t* neg_out_i32, void* stream);
/* sizeof(rsb200_pair_args) as compiled into the library (binding sanity check) */
size_t  rsb200_sizeof_pair_args(void);

/* -------------------------------------------------------------------------
 * E1 / E2 standalone: nn.Embedding forward / backward
 *   F.embedding(ids, W)                       baseretriever.py:154,168,211
 *   embedding_dense_backward (autograd)       recommender.py:638
 */
int32_t rsb200_gather_rows(const float* w, int64_t num_rows, int64_t d,
                           const int64_t* ids, int64_t numel, float* out /* [numel, d] */, void* stream);
/* dW[ids[i], :] += d_out[i, :] for ids[i] != 0, dW dense [num_rows, d] (caller zero-fills) */
int32_t rsb200_scatter_add_rows(float* dw, int64_t num_rows, int64_t d,
                                const int64_t* ids, int64_t numel, const float* d_out, void* stream);

/* Sum value rows that target the same table row: (ids[M], vals[M,d]) -> (rows_out[R] ascending
 * unique, vals_out[R,d]) (sink COMPACT) or vals_out dense [num_rows,d] (sink DENSE, overwrite or
 * accumulate).  Owner-side accumulate of the row-sharded table (the reference's
 * embedding_dense_backward into one dense gradient, recommender.py:638).  skip_row0 != 0 drops
 * ids equal to 0 (padding row).  Workspace: off[num_rows+1] u32, slot[M] u32, ent[M] u64,
 * urow[cap] u32 (cap >= min(M, num_rows)), scan_tmp[scan_tmp_elems >= num_rows/4096 + 2] u64. */
int32_t rsb200_rows_coalesce(const int64_t* ids, const float* vals, int64_t M, int64_t num_rows, int64_t d,
                             int32_t skip_row0, int64_t* rows_out, float* vals_out, int32_t sink, int32_t accumulate,
                             uint32_t* totals /* [2] = {entries, unique rows} */, uint32_t* off, uint32_t* slot,
                             uint64_t* ent, uint32_t* urow, int64_t cap, uint64_t* scan_tmp, int64_t scan_tmp_elems,
                             uint32_t* err_flag, void* stream);

/* Optimizer step on the touched rows only (next-row 8(f)-1; the reference: dense optimizer over the whole
 * table, recommender.py:445-474,646).  rows[R] / vals[R,d] as produced by rsb200_pair_step (SINK_COMPACT);
 * R is read from DEVICE memory (*count_dev, e.g. &totals[1]) -- no host sync.  kind: 0 SGD, 1 Adagrad
 * (state1 = sum of squares), 2 SparseAdam (state1 = exp_avg, state2 = exp_avg_sq, step >= 1). */
int32_t rsb200_rows_update(int32_t kind, float* w, float* state1, float* state2, int64_t num_rows, int64_t d,
                           const int64_t* rows, const float* vals, const uint32_t* count_dev, int64_t cap,
                           int64_t step, float lr, float beta1, float beta2, float eps, void* stream);

/* -------------------------------------------------------------------------
 * Q1 / Q2 standalone on ids (no [B,n,d] materialisation):
 *   score[b, j] = score_func(q[b], W[ids[b, j]])       scorer.py:10-14 / 28-34
 */
int32_t rsb200_score_ids(int32_t score_kind, const float* q /* [B,d] */, const float* w, int64_t num_rows,
                         int64_t d, const int64_t* ids /* [B, n] */, int64_t B, int64_t n,
                         float* out /* [B, n] */, void* stream);

/* Q1 / Q2 standalone on an already gathered [B, n, d] item tensor (per-row branches of
 * scorer.py:10-14; n = 1 covers ([B,D],[B,D])) and its backward. */
int32_t rsb200_score_dense(int32_t score_kind, const float* q /* [B,d] */, const float* items /* [B,n,d] */,
                           int64_t B, int64_t n, int64_t d, float* out /* [B,n] */, void* stream);
int32_t rsb200_score_dense_bwd(int32_t score_kind, const float* q, const float* items, const float* g /* [B,n] */,
                               int64_t B, int64_t n, int64_t d, float* dq /* [B,d] */, float* ditems /* [B,n,d] */,
                               void* stream);

/* -------------------------------------------------------------------------
 * L1 / L2 standalone on score tensors (for mixing with reference plugins):
 *   loss and d loss / d score in one pass.     loss_func.py:55-59, 80-90
 */
int32_t rsb200_pair_loss(int32_t loss_kind, const float* pos_score /* [B] */, const float* neg_score /* [B,n] */,
                         const float* logq_pos, const float* logq_neg, int64_t B, int64_t n,
                         float* loss /* [1] */, float* d_pos /* [B] */, float* d_neg /* [B,n] */,
                         float* loss_part /* [B] workspace */, void* stream);

/* -------------------------------------------------------------------------
 * T1  BaseRetriever.topk (no ANN index, InnerProduct | Euclid)
 *   recstudio/model/basemodel/baseretriever.py:374-397
 * Scores every item row 1..num_items-1 against each query, masks the ids in
 * hist (0 = padding), returns the k best (score desc, id asc on ties),
 * ids 1-based.  [Be, num_items] is never written to HBM (only one max per 8 items).
 * Limits: k + H + 8 <= 1024.
 */
size_t  rsb200_topk_workspace_bytes(int64_t Be, int64_t num_items, int64_t k, int64_t H);
int32_t rsb200_topk_full(int32_t score_kind, const float* q /* [Be,d] */, const float* w_item, int64_t num_items,
                         int64_t d, int64_t Be, int64_t k, const int64_t* hist /* [Be,H] or NULL */, int64_t H,
                         float* score_out /* [Be,k] */, int64_t* id_out /* [Be,k] */,
                         void* workspace, size_t workspace_bytes, void* stream);

/* -------------------------------------------------------------------------
 * L3 + Q1-full  SoftmaxLoss over the whole catalog (forward + backward)
 *   baseretriever.py:177-186, loss_func.py:39-42
 */
size_t  rsb200_fullsoftmax_workspace_bytes(int64_t B, int64_t num_items, int64_t d);
int32_t rsb200_fullsoftmax_fwd_bwd(const float* q /* [B,d] */, const float* w_item, const int64_t* pos /* [B] */,
                                   int64_t num_items, int64_t B, int64_t d,
                                   float* loss /* [1] */, float* dq /* [B,d] */, float* dw /* [num_items,d] dense, overwritten */,
                                   void* workspace, size_t workspace_bytes, void* stream);

/* -------------------------------------------------------------------------
 * A1  attention core of SASRecQueryEncoder / BERT4Rec on tcgen05 tensor cores (bf16 operands,
 *     fp32 accumulation in TMEM).   recstudio/model/seq/sasrec.py:19-32,44-53
 *     (nn.TransformerEncoderLayer self-attention: softmax(Q K^T / sqrt(dh) + masks) V per head;
 *      mask = causal triu(1) (unless bidirectional) AND key padding hist == 0)
 * q, k, v, out: fp32 [B, L, heads * head_dim] (head h occupies columns [h*dh, (h+1)*dh));
 * hist: int64 [B, L] item ids (0 = padding key) or NULL; lse: [B, heads, L] or NULL.
 * Limits: head_dim == 64, L <= 256.  err_flag[0] is set if a tensor-core barrier timed out.
 * bf16 operands: results match an fp32 evaluation to ~1e-2 relative (not the 1e-5 of the fp32 paths).
 * p_drop / drop_key: attention-probability dropout (nn.MultiheadAttention(dropout=p): out = (softmax(.) o M / (1 - p)) V with
 * M ~ Bernoulli(1 - p)).  M is a counter-based hash of (drop_key, sequence, head, query, key): stateless, so rsb200_attn_bwd
 * regenerates it from the same drop_key; p_drop = 0 disables it.  (The mask is NOT torch's Philox mask: dropout is random
 * by definition; tests rebuild this hash on the host to check the arithmetic against a torch evaluation with the same mask.)
 */
int32_t rsb200_attn_fwd(const float* q, const float* k, const float* v, const int64_t* hist, int64_t B, int64_t L,
                        int64_t heads, int64_t head_dim, int32_t causal, float p_drop, uint64_t drop_key, float* out, float* lse,
                        uint32_t* err_flag, void* stream);
/* backward of rsb200_attn_fwd: o / lse are its outputs, d_o the upstream gradient [B, L, heads*head_dim];
 * writes dq, dk, dv (same layout, overwritten).  dV = P~^T dO, dS = P o (dP - rowsum(dO o O)) / sqrt(dh) with
 * P~ = P o M / (1 - p), dP = (dO V^T) o M / (1 - p);  dQ = dS K, dK = dS^T Q, all five contractions on tcgen05. */
int32_t rsb200_attn_bwd(const float* q, const float* k, const float* v, const float* o, const float* d_o, const float* lse,
                        const int64_t* hist, int64_t B, int64_t L, int64_t heads, int64_t head_dim, int32_t causal,
                        float p_drop, uint64_t drop_key, float* dq, float* dk, float* dv, uint32_t* err_flag, void* stream);
/* validation hook for the tcgen05 plumbing: D[128,N] = bf16(A[128,K]) * bf16(B[N,K])^T, fp32 accumulate */
int32_t rsb200_tc_gemm_test(const float* A, const float* B, float* D, int64_t N, int64_t K, uint32_t* err_flag, void* stream);

/* -------------------------------------------------------------------------
 * 8(f)-4  Index build of the model-based samplers (Sampler.update, once per epoch:
 *         recstudio/model/basemodel/recommender.py:561-570)
 *
 * kmeans (recstudio/ann/sampler.py:9-36), one Lloyd iteration = assign + update:
 *   assign[i] = argmin_k |x_i - c_k|^2 evaluated as |x|^2 - 2 x.c + |c|^2 (first minimum on ties), and
 *   *loss_out = sum_i |x_i - c_assign(i)|^2 (double).  x is [num_points, d] with row stride ldx floats (a column
 *   chunk of the item table: torch.chunk(item_embs, 2, -1), sampler.py:275); cnorm_ws: float [num_clusters].
 *   update: sums_out[k,:] = sum of the points assigned to k, counts_out[k] = their number (as float, like
 *   assign_m.sum(0)); the caller divides and re-seeds empty clusters (sampler.py:31-35).
 *   Neither the [N,K] distance matrix nor the [N,K] one-hot matrix of the reference is materialised. */
int32_t rsb200_kmeans_assign(const float* x, int64_t ldx, int64_t num_points, int64_t d, const float* centers /* [K,d] */,
                             int64_t num_clusters, float* cnorm_ws, int64_t* assign_out, double* loss_out, void* stream);
int32_t rsb200_kmeans_update(const float* x, int64_t ldx, int64_t num_points, int64_t d, const int64_t* assign,
                             int64_t num_clusters, float* sums_out /* [K,d] */, float* counts_out /* [K] */, void* stream);
/* construct_index (sampler.py:39-45): indices = argsort(codes, stable), indptr[c] = first position of bucket c
 * (indptr has num_buckets + 1 entries).  codes in [0, num_buckets), num_buckets <= 65536. */
size_t  rsb200_index_workspace_bytes(int64_t num_points, int64_t num_buckets);
int32_t rsb200_index_build(const int64_t* codes, int64_t num_points, int64_t num_buckets, int64_t* indices_out,
                           int64_t* indptr_out, void* workspace, size_t workspace_bytes, void* stream);
/* per-bucket normalised cumulative weights (sampler.py:300-306,417-423): cp_out[e] = cumsum within the bucket of
 * weight[indices[e]] divided by the bucket total; total_out[c] (or NULL) = bucket total (= an entry of wkk). */
int32_t rsb200_segment_cdf(const float* weight /* [num_points] */, const int64_t* indices, const int64_t* indptr,
                           int64_t num_buckets, float* cp_out /* [num_points] */, float* total_out /* [num_buckets] */, void* stream);
/* _sample_item_with_pop (sampler.py:348-365) for num_draws (bucket, uniform seed) pairs: inverse-CDF search inside
 * the bucket.  neg_out = indices[start + idx] and logp_out = log(p[start + idx + 1]) exactly as the reference indexes
 * them; no [num_q, neg, max_bucket] tensor. */
int32_t rsb200_segment_search(const int64_t* k01, const float* u, int64_t num_draws, const float* cp, const int64_t* indices,
                              const int64_t* indptr, int64_t num_buckets, const float* p /* [num_points + 1] */,
                              int64_t* neg_out, float* logp_out, void* stream);

/* -------------------------------------------------------------------------
 * 8(e)  Row-sharded item table, owner-compute ("ship queries, not rows") training step.
 *   The same step as rsb200_pair_step (baseretriever.py:142-176,399-404 + loss.backward(),
 *   recommender.py:638) for a GLOBAL batch of G = world x B interactions, restricted to the rows ONE
 *   owner holds (contiguous block [row0, row0 + local_rows) of the [num_items, d] table).  No torch /
 *   NCCL types: the three small exchanges between the phases are the caller's (NCCL, gloo, or a loop
 *   over owners on one device):
 *
 *     RSB200_SHARD_PREP    -> all-reduce SUM  sp[G]                 (4 B per query)
 *     RSB200_SHARD_FWD     -> all-gather      stats_all[rank] slices (8 B per query)
 *     RSB200_SHARD_FINISH  -> all-reduce SUM  dq[G, d]              (d loss / d query, 4d B per query)
 *     RSB200_SHARD_SCATTER    gradient rows of the OWNED rows (never leave the owner)
 *
 *   loss is the mean over the global batch (identical on every owner); gradients are those of that
 *   global mean.  Touches of global row 0 (padding) are scored but get no gradient row.
 */
#define RSB200_SHARD_PREP     1
#define RSB200_SHARD_FWD      2
#define RSB200_SHARD_FINISH   4
#define RSB200_SHARD_SCATTER  8
/* PREP may be issued in two halves (grouping 1): PREP_NEG filters / compacts the negatives -- it needs the draw (neg or
 * regen_state) but NOT q_all / pos, so it can run while the queries are still being all-gathered -- and PREP_POS scores the
 * owned positives and closes the bin histogram.  RSB200_SHARD_PREP = both. */
#define RSB200_SHARD_PREP_NEG 16
#define RSB200_SHARD_PREP_POS 32

typedef struct rsb200_shard_args {
    /* owner's table block and the global batch (device pointers) */
    const float*   w_local;     /* [local_rows, d] rows row0 .. row0+local_rows-1 of the table */
    const float*   q_all;       /* [G, d] query vectors of the global batch                    */
    const int64_t* pos;         /* [G]    GLOBAL positive item ids                             */
    const int32_t* neg;         /* [G, n] GLOBAL negative item ids                             */
    const float*   logq_pos;    /* [G]    or NULL (SampledSoftmax)                             */
    const float*   logq_neg;    /* [G, n] or NULL                                              */
    const float*   grad_scale_dev; /* [1] or NULL, multiplied into the gradient rows by SCATTER */
    /* exchanged between owners by the caller */
    float*         sp;          /* [G]    PREP: positive score if owned, else 0                */
    float*         stats_all;   /* [world, G, 2]  FWD writes slice [rank]                      */
    float*         dq;          /* [G, d] FWD: raw partial; FINISH: this owner's share of d loss / d query */
    /* outputs */
    float*         loss;        /* [1]  global mean loss (after FINISH)                        */
    int64_t*       item_rows;   /* [cap] LOCAL row ids, ascending unique (SCATTER)             */
    float*         item_vals;   /* COMPACT: [cap, d]; DENSE: [local_rows, d]                   */
    uint32_t*      totals;      /* [2] = {owned touches with a gradient, unique owned rows}    */
    /* workspace */
    int32_t*       neg_c;       /* [G * n] compacted LOCAL ids of the owned negatives (stride n) */
    uint32_t*      slot_neg;    /* [G * n]                                                     */
    float*         lq_c;        /* [G * n] or NULL when logq_neg is NULL                       */
    int32_t*       ncount;      /* [G]   owned negatives per query                             */
    int32_t*       pos_local;   /* [G]   local row of the positive or -1                       */
    uint32_t*      slot_pos;    /* [G]                                                         */
    uint32_t*      off;         /* [local_rows + 1] histogram -> CSR offsets                   */
    uint32_t*      urow;        /* [cap]                                                       */
    uint64_t*      ent;         /* [G * (n + 1)] worst case: every touch owned here            */
    float*         loss_part;   /* [G]                                                         */
    float*         lse;         /* [G]                                                         */
    uint64_t*      scan_tmp;    /* [scan_tmp_elems >= rsb200_scan_tmp_elems(local_rows)]       */
    uint32_t*      err_flag;    /* [1] set non-zero if an id was outside [0, num_items)        */
    /* sizes */
    int64_t G, n, d, num_items, row0, local_rows;
    int64_t cap;                /* capacity of urow / item_rows: >= min(G*(n+1), local_rows)   */
    int64_t scan_tmp_elems;
    float   grad_scale;         /* upstream gradient (loss.backward() => 1.0)                  */
    int32_t world, rank;        /* number of owners, index of this owner in stats_all          */
    int32_t loss_kind, score_kind, sink, accumulate;
    /* Owner-side regeneration of the UniformSampler draw (optional; neg may then be NULL): instead of all-gathering the
     * 4-byte negative ids of every rank, every owner recomputes them from the ranks' generator states -- rank r's ids are
     * exactly what rsb200_sample_uniform(seed[r], offset[r], num_items, regen_B, n) would write, i.e. what
     * torch.randint(1, num_items, (regen_B, n), device='cuda') returns on rank r.  Query g belongs to rank g / regen_B. */
    const uint64_t* regen_state;   /* DEVICE [world, 2] = (seed, philox offset) of every rank, or NULL (ids given in neg) */
    int64_t regen_B;               /* queries per rank (G == world * regen_B)                   */
    int32_t regen_sm_count, regen_max_threads_per_sm;   /* ATen draw policy, as for rsb200_sample_uniform */
    /* The same for PopularSamplerModel (regen_kind = 1; recstudio/ann/sampler.py:243-258): rank r's ids are
     * searchsorted(table, torch.rand(regen_B, n, device='cuda')) for its (seed, offset), and an owner needs only ITS slices of
     * the sampler's buffers: a draw u lands on an owned row iff pop_cdf_lo < u <= pop_cdf_hi (pop_cdf_lo = table[row0 - 1],
     * -inf for the first owner; pop_cdf_hi = table[row0 + local_rows - 1], +inf for the last), the search runs inside the
     * slice, and log Q(neg) = log(pop_prob_local[id]) is written to lq_c.  pop_guide_local (optional) holds
     * first-i-with-table_local[i] >= k / 2^pop_guide_bits for k = pop_guide_k0 ... (rsb200_popular_build_guide_range). */
    int32_t        regen_kind;         /* 0 UniformSampler, 1 PopularSamplerModel                                */
    int32_t        pop_guide_bits;
    const float*   pop_table_local;    /* [local_rows] table[row0 .. row0 + local_rows)                          */
    const float*   pop_prob_local;     /* [local_rows] pop_prob[row0 .. row0 + local_rows)                       */
    const int32_t* pop_guide_local;    /* [pop_guide_len] or NULL                                                */
    int64_t        pop_guide_k0;
    float          pop_cdf_lo, pop_cdf_hi;
    float*         lq_pos_out;         /* [G] or NULL: PREP writes log(pop_prob_local[pos]) for owned positives, else 0
                                          (all-reduce SUM it together with sp to obtain logq_pos for FINISH)     */
    /* Binned grouping of the owned touches (grouping = 1), exactly as in rsb200_pair_args: bins of 2^bin_shift LOCAL rows;
     * slot_neg / slot_pos / off / urow / scan_tmp are then unused and may be NULL. */
    int32_t   grouping;
    int32_t   bin_shift;
    uint32_t* bin_cnt;                 /* [nbins + 1], nbins = ceil(local_rows / 2^bin_shift); the last word is    */
                                       /* PREP_NEG's work counter (zeroed by the library together with the counts) */
    uint32_t* bin_off;                 /* [nbins + 1]                                                            */
    uint32_t* bin_cursor;              /* [nbins * 8]                                                            */
    uint64_t* bin_status;              /* [nbins]                                                                */
    uint32_t* bin_ticket;              /* [1]                                                                    */
    uint32_t* bin_heavy;               /* [rsb200_bin_heavy_elems()]                                             */
} rsb200_shard_args;

/* rows-per-bin exponent the library suggests for a table block of num_rows rows that receives `touches` gradient touches
 * per step from num_queries queries (0 = not supported: use grouping 0), and the size of the bin_heavy scratch */
int32_t rsb200_bin_shift(int64_t num_rows, int64_t touches, int64_t num_queries);
int64_t rsb200_bin_heavy_elems(void);
/* host arithmetic behind the owner-side regeneration of UniformSampler draws (sampler.py:86-111: ids = randint(1, N)): a Philox
 * word v is the id v mod (N - 1) + 1; it lands on rows [row0, row0 + local_rows) iff lo <= low64(magic * v) <= hi (exact for
 * every 32-bit v; magic = ceil(2^64 / (N - 1))).  An empty block gives lo = 1, hi = 0.  No device work. */
int32_t rsb200_uniform_owner_range(int64_t num_items, int64_t row0, int64_t local_rows, uint64_t* magic, uint64_t* lo, uint64_t* hi);
/* guide entries k0 .. k0 + len - 1 of a table slice: out[i] = first j in [0, num_rows - 1] with table_local[j] >= (k0 + i) / 2^bits
 * (num_rows - 1 if none) */
int32_t rsb200_popular_build_guide_range(const float* table_local, int64_t num_rows, int32_t guide_bits, int64_t k0, int64_t len,
                                         int32_t* guide_out, void* stream);

int32_t rsb200_shard_step(const rsb200_shard_args* args, int32_t phases, void* stream);
size_t  rsb200_sizeof_shard_args(void);
int64_t rsb200_scan_tmp_elems(int64_t num_rows);

#ifdef __cplusplus
}
#endif
#endif /* RSB200_H */
