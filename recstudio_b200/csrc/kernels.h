// kernels.h -- internal launch interfaces shared by the .cu translation units.
#pragma once
#include "common.cuh"

namespace rsb {

struct FwdParams {
    const float* w_item; const float* w_user;
    const int64_t* user; const int64_t* pos; const int32_t* neg;
    const float* logq_pos; const float* logq_neg;
    const uint32_t* off_item; const uint32_t* off_user;
    const uint32_t* slot_neg; const uint32_t* slot_pos; const uint32_t* slot_user;
    uint64_t* ent_item; uint64_t* ent_user;
    float* q_buf; float* dq_buf; float* loss_part; float* lse;
    float* pos_score; float* neg_score;
    int num_items, num_users, B, n, D;
    float coef_scale;   // grad_scale / (B n) for BPR, grad_scale / B for SSM
    float loss_scale;   // 1 / (B n)              for BPR, 1 / B              for SSM
    int prefetch;       // variant 3: L2-prefetch the next batch's rows
    int slot_abs;       // slot_neg already holds absolute entry positions (group.cu resolve_kernel)
    float* cstage;      // [B, n] staged coefficients / logits in touch order (entries are permuted later), or null
    int hint;           // L2 eviction hints: bit 0 rows evict_first, bit 1 offsets evict_last, bit 2 entries evict_last
    // owner-compute (PARTIAL) mode of the row-sharded step, see shard.cu
    // binned grouping (bins.cu): entries are appended to the list of their row's bin with a cursor atomic
    uint32_t* bin_cursor;    // [nbins * kCursorStride] or null (= row-sorted positions from slot_neg / off_item)
    int bin_shift, bin_bbits;
    int pad_row;             // (binned grouping) id of the padding row in `neg`: 0, or -1 for an owner whose block does not hold global row 0
    uint32_t* bin_cursor_user; int bin_shift_user;    // the same for the user table's entries (null in the owner-compute step)
    const int32_t* ncount;   // [B] length of each query's compacted negative list (stride n)
    const float* sp_in;      // [B] positive score (computed by the positive's owner)
    float* stats_part;       // [B, 2] {csum, loss} (BPR) | {m, l} (SSM)
};

struct ScatterParams {
    const uint32_t* off;      // [num_rows + 1] CSR offsets
    const uint32_t* urow;     // [R] unique touched rows, ascending
    const uint32_t* totals;   // totals[1] = R
    const uint64_t* ent;      // entries grouped by row
    const float* src;         // [B, D]
    const float* lse;         // [B] or null
    const float* w;           // table (Euclid only)
    const float* gscale;      // [1] device upstream gradient or null
    int64_t* rows_out;        // [R]
    float* vals;              // compact [R, D] or dense [num_rows, D]
    int64_t cap;              // capacity of urow / rows_out
    int D;
    float ssm_scale;
    int dense, accumulate, euclid;
    int hint;                 // L2 eviction hints (PLAIN kernel): entries / output rows evict_first, src evict_last
    // optimizer fused into the epilogue (RSB200_SINK_APPLY): opt < 0 = off
    int opt;                  // 0 SGD, 1 Adagrad, 2 SparseAdam
    float* w_rw; float* s1; float* s2;
    float lr, b1, b2, eps, step_size;
};

// bins.cu: two-level grouping (bins of 2^shift rows, entries appended per bin, grouped by row in shared memory)
constexpr int kMinBinShift = 4, kMaxBinShift = 12;
constexpr int kCursorStride = 8;      // u32 words between two bins' append cursors: one 32-byte sector each
struct BinScatterParams {
    const uint64_t* ent;      // entries grouped by bin (arbitrary order inside a bin)
    const uint32_t* bin_off;  // [nbins + 1]
    uint64_t* status;         // [nbins] look-back words (cleared by bin_scan)
    uint32_t* ticket;         // [1]
    uint32_t* totals;         // totals[1] <- number of touched rows
    uint32_t* heavy_counts;   // [grid, 4096] per-CTA scratch for bins that exceed one chunk
    const float* src;         // [B, D]
    const float* lse;         // [B] or null
    const float* w;           // table (Euclid only)
    const float* gscale;      // [1] device upstream gradient or null
    int64_t* rows_out;        // [cap] or null
    float* vals;              // compact [cap, D] or dense [num_rows, D]
    int64_t cap;
    int nbins, shift, bbits, D;
    float ssm_scale;
    int dense, accumulate, euclid;
    int opt;                  // < 0: gradient sink; 0 SGD, 1 Adagrad, 2 SparseAdam applied in the epilogue
    float* w_rw; float* s1; float* s2;
    float lr, b1, b2, eps, step_size;
    int tune;                 // A/B switch of the hot configuration's (loads in flight, CTAs/SM); 0 = default
};
int bin_shift_for(int64_t num_rows, int64_t touches, int64_t num_queries);
int64_t bin_scatter_grid();
struct BinTable {             // the bin arrays of one table
    uint32_t* cnt;            // [nbins] touches per bin
    uint32_t* off;            // [nbins + 1] entry offsets
    uint32_t* cursor;         // [nbins * kCursorStride] append cursors
    uint64_t* status;         // [nbins]
    uint32_t* ticket;         // [1]
    uint32_t* totals;         // totals[0] <- entries, totals[1] <- touched rows
    int nbins, shift;
    int64_t num_rows;
};
template <typename IdT>
int32_t launch_bin_count(const IdT* ids, int64_t M, const int64_t* pos, int64_t B, const BinTable& t0, const int64_t* ids1,
                         int64_t B1, const BinTable& t1, int32_t* ids32_out, uint32_t* err_flag, cudaStream_t st);
int32_t launch_bin_scan(const BinTable& t0, const BinTable& t1, cudaStream_t st);
int32_t launch_draw_bin_count(uint64_t* state_dev, uint64_t seed, uint64_t philox_offset, int64_t num_queries, int64_t num_neg,
                              int32_t sm_cnt, int32_t max_tpsm, int32_t* out32, const int64_t* pos, int64_t B, const BinTable& t0,
                              const int64_t* ids1, int64_t B1, const BinTable& t1, uint32_t* err_flag, cudaStream_t st);
int32_t launch_bin_scatter(const BinScatterParams& p, cudaStream_t st);
// group.cu
int64_t scan_tmp_elems(int64_t num_rows);
template <typename IdT>
int32_t launch_count(const IdT* ids, int64_t M, int64_t num_rows, uint32_t* cnt, uint32_t* slot,
                     int32_t* ids32_out, uint32_t* err_flag, cudaStream_t st);
int32_t launch_resolve(const int32_t* ids, uint32_t* slot, const uint32_t* off, int64_t M, cudaStream_t st);
int32_t launch_permute_entries(const uint32_t* epos, const float* cstage, int64_t M, int64_t n, uint32_t flag, uint64_t* ent,
                               cudaStream_t st);
int32_t launch_scan(uint32_t* cnt_off, int64_t num_rows, uint32_t* urow, int64_t cap, uint32_t* totals,
                    uint64_t* tmp, int64_t tmp_elems, cudaStream_t st);
// pair_fwd.cu
int32_t launch_pair_fwd(const FwdParams& p, int loss, int score, int variant, cudaStream_t st);
int32_t launch_pair_fwd_partial(const FwdParams& p, int loss, int score, cudaStream_t st);
// scatter.cu
int32_t launch_scatter(const ScatterParams& p, int64_t cap_rows, cudaStream_t st);
int32_t launch_loss_sum(const float* part, int B, float* loss, cudaStream_t st);

}  // namespace rsb
