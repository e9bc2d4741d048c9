// bins.cu -- two-level ("binned") grouping of gradient touches and the per-bin gradient scatter.
//
// Replaces, for the item table, the N-bucket counting sort of group.cu + scatter.cu (the sparse-row
// form of embedding_dense_backward, autograd of nn.Embedding at recommender.py:638):
//
//   level 1  a table of num_rows rows is cut into bins of 2^shift consecutive rows.  BIN_COUNT builds the
//            histogram of touches per bin (shared-memory aggregated: one global atomic per CTA and bin),
//            BIN_SCAN turns it into entry offsets and per-bin append cursors.  The forward kernel
//            (pair_fwd.cu) appends one 8-byte entry per touch to its bin's list with a cursor atomic:
//            the tail sector of a list is completed within microseconds, so entries leave L2 as
//            full sectors (the row-sorted layout of group.cu cost one 32-byte read-modify-write per entry),
//            and the 10M-counter histogram, its scan, the slot / resolve arrays are gone.
//   level 2  bin_scatter_kernel: one CTA per bin loads the bin's entries into shared memory, groups them
//            by row there (counting sort over the 2^shift local rows), and forms every touched row's
//            gradient  g[row] = sum_e c_e * src[b_e]  exactly once: no fp atomics, no [N,d] zero-fill.
//            Entries of one row are summed in ascending order of their 64-bit encoding, so the result does
//            not depend on the order in which the forward kernel's atomics landed: same inputs => same bits.
//            The compact sink needs the rank of a bin's first unique row among all unique rows: bins are
//            taken in ticket order and chained with a decoupled look-back over per-bin status words.
//
// Entry: low word = kDirect flag (bit 31) | local row (shift bits) | query index (bbits = 31 - shift bits),
//        high word = coefficient (flag set) or logit (SampledSoftmax negatives), as in scatter.cu.
#include "common.cuh"
#include "kernels.h"
#include "rowopt.cuh"

namespace rsb {

constexpr uint64_t kStAgg = 1ull << 62;          // status word: own unique-row count published
constexpr uint64_t kStPre = 2ull << 62;          // status word: inclusive prefix published
constexpr uint32_t kLast = 0x8000u;              // idx flag: last entry of its row
constexpr int kBsDepth = 8;                      // query-row loads in flight per warp (d <= 128)
constexpr int kBsOcc = 3;                        // CTAs per SM (measured at config 2: 8 x 3 0.617 ms, 16 x 2 0.671, 8 x 2 0.70, 4 x 3 0.66)

// ------------------------------------------------------------------------------------------ policy
// bins of 2^shift rows, sized so that a bin receives ~kBinTarget touches on average (one chunk), limited by the
// bits left for the query index in the entry's low word.  shift < kMinBinShift => the caller uses group.cu.
int bin_shift_for(int64_t num_rows, int64_t touches, int64_t num_queries) {
    constexpr double kBinTarget = 3400.0;
    int qbits = 1;
    while (((int64_t)1 << qbits) < num_queries) ++qbits;
    int shift = kMinBinShift;
    const double want = kBinTarget * (double)num_rows / (double)(touches > 0 ? touches : 1);
    while (shift < kMaxBinShift && (double)((int64_t)1 << (shift + 1)) <= want) ++shift;
    if (shift > 31 - qbits) shift = 31 - qbits;
    return shift;
}

// ------------------------------------------------------------------------------------------ BIN_COUNT
// Histogram of the touches per bin for up to two tables in one launch: table 0 is touched by ids[M] and pos[B]
// (negatives and positives of the item table), table 1 by ids1[B1] (the user table).  Every CTA aggregates in shared
// memory and flushes once: one global atomic per CTA and non-empty bin.
template <typename IdT>
__global__ void __launch_bounds__(256)
bin_count_kernel(const IdT* __restrict__ ids, int64_t M, const int64_t* __restrict__ pos, int64_t B, const BinTable t0,
                 const int64_t* __restrict__ ids1, int64_t B1, const BinTable t1, int use_smem, int32_t* __restrict__ ids32_out,
                 uint32_t* __restrict__ err_flag) {
    extern __shared__ uint32_t s_hist[];
    const int nb0 = t0.nbins, nb1 = t1.nbins;
    if (use_smem) {
        for (int i = threadIdx.x; i < nb0 + nb1; i += blockDim.x) s_hist[i] = 0u;
        __syncthreads();
    }
    uint32_t* h0 = use_smem ? s_hist : t0.cnt;
    uint32_t* h1 = use_smem ? s_hist + nb0 : t1.cnt;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool bad = false;
    constexpr int kPer = 4;
    for (int64_t base = i0; base < M; base += stride * kPer) {
        int64_t id[kPer];
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int64_t i = base + k * stride;
            id[k] = i < M ? (int64_t)ids[i] : 0;
        }
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int64_t i = base + k * stride;
            if (id[k] > 0 && id[k] < t0.num_rows) atomicAdd(h0 + (id[k] >> t0.shift), 1u);
            else if (id[k] != 0) bad = true;
            if (ids32_out && i < M) ids32_out[i] = (id[k] >= 0 && id[k] < t0.num_rows) ? (int32_t)id[k] : 0;
        }
    }
    for (int64_t i = i0; i < B; i += stride) {
        const int64_t id = pos[i];
        if (id > 0 && id < t0.num_rows) atomicAdd(h0 + (id >> t0.shift), 1u);
        else if (id != 0) bad = true;
    }
    for (int64_t i = i0; i < B1; i += stride) {
        const int64_t id = ids1[i];
        if (id > 0 && id < t1.num_rows) atomicAdd(h1 + (id >> t1.shift), 1u);
        else if (id != 0) bad = true;
    }
    if (bad) *err_flag = 1u;
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < nb0 + nb1; i += blockDim.x) {
            const uint32_t c = s_hist[i];
            if (c) atomicAdd(i < nb0 ? t0.cnt + i : t1.cnt + (i - nb0), c);
        }
    }
}

template <typename IdT>
int32_t launch_bin_count(const IdT* ids, int64_t M, const int64_t* pos, int64_t B, const BinTable& t0, const int64_t* ids1,
                         int64_t B1, const BinTable& t1, int32_t* ids32_out, uint32_t* err_flag, cudaStream_t st) {
    RSB_CUDA(cudaMemsetAsync(t0.cnt, 0, sizeof(uint32_t) * (size_t)t0.nbins, st));
    if (t1.nbins) RSB_CUDA(cudaMemsetAsync(t1.cnt, 0, sizeof(uint32_t) * (size_t)t1.nbins, st));
    if (M + B + B1 == 0) return 0;
    const int use_smem = t0.nbins + t1.nbins <= 8192;
    int64_t blocks = cdiv(M + B + B1, 256 * 16);
    const int64_t cap = (int64_t)sm_count() * 2;           // few CTAs: one flush of the shared histogram per CTA
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const size_t smem = use_smem ? sizeof(uint32_t) * (size_t)(t0.nbins + t1.nbins) : 0;
    bin_count_kernel<IdT><<<(unsigned)blocks, 256, smem, st>>>(ids, M, pos, B, t0, ids1, B1, t1, use_smem, ids32_out, err_flag);
    RSB_LAUNCH_CHECK();
    return 0;
}
template int32_t launch_bin_count<int32_t>(const int32_t*, int64_t, const int64_t*, int64_t, const BinTable&, const int64_t*, int64_t,
                                           const BinTable&, int32_t*, uint32_t*, cudaStream_t);
template int32_t launch_bin_count<int64_t>(const int64_t*, int64_t, const int64_t*, int64_t, const BinTable&, const int64_t*, int64_t,
                                           const BinTable&, int32_t*, uint32_t*, cudaStream_t);

// ------------------------------------------------------------------------------------------ DRAW + BIN_COUNT
// The UniformSampler draw (sampler.cu: torch.randint's Philox element map) with the bin histogram taken while the ids are
// still in registers: the 33 MB id matrix is written once and not re-read by a separate count pass.  (seed, offset) come
// from device memory when state != null (CUDA-graph replays draw fresh ids), else from the arguments.
__global__ void __launch_bounds__(256)
draw_bin_count_kernel(const uint64_t* __restrict__ state, uint64_t seed_h, uint64_t ctr_h, int64_t T, int64_t rounds, int64_t numel,
                      uint32_t range, uint64_t mod_magic, int32_t* __restrict__ out32, const int64_t* __restrict__ pos, int64_t B,
                      const BinTable t0, const int64_t* __restrict__ ids1, int64_t B1, const BinTable t1, int use_smem,
                      uint32_t* __restrict__ err_flag) {
    extern __shared__ uint32_t s_hist[];
    const int nb0 = t0.nbins, nb1 = t1.nbins;
    if (use_smem) {
        for (int i = threadIdx.x; i < nb0 + nb1; i += blockDim.x) s_hist[i] = 0u;
        __syncthreads();
    }
    uint32_t* h0 = use_smem ? s_hist : t0.cnt;
    uint32_t* h1 = use_smem ? s_hist + nb0 : t1.cnt;
    const uint64_t seed = state ? state[0] : seed_h, ctr_base = state ? state[1] / 4 : ctr_h;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t t = i0; t < T * rounds; t += stride) {
        const int64_t r = t / T, idx = t - r * T;
        const uint4 w = Philox::gen(seed, (uint64_t)idx, ctr_base + (uint64_t)r);
        const uint32_t words[4] = {w.x, w.y, w.z, w.w};
        int64_t li = r * 4 * T + idx;
#pragma unroll
        for (int ii = 0; ii < 4; ++ii, li += T) {
            if (li < numel) {
                const uint32_t v = (uint32_t)__umul64hi(mod_magic * (uint64_t)words[ii], (uint64_t)range) + 1u;   // words % range + 1
                out32[li] = (int32_t)v;
                atomicAdd(h0 + (v >> t0.shift), 1u);
            }
        }
    }
    bool bad = false;
    for (int64_t i = i0; i < B; i += stride) {
        const int64_t id = pos[i];
        if (id > 0 && id < t0.num_rows) atomicAdd(h0 + (id >> t0.shift), 1u);
        else if (id != 0) bad = true;
    }
    for (int64_t i = i0; i < B1; i += stride) {
        const int64_t id = ids1[i];
        if (id > 0 && id < t1.num_rows) atomicAdd(h1 + (id >> t1.shift), 1u);
        else if (id != 0) bad = true;
    }
    if (bad) *err_flag = 1u;
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < nb0 + nb1; i += blockDim.x) {
            const uint32_t c = s_hist[i];
            if (c) atomicAdd(i < nb0 ? t0.cnt + i : t1.cnt + (i - nb0), c);
        }
    }
}

__global__ void advance_philox_state_kernel(uint64_t* state, uint64_t inc) { state[1] += inc; }

int32_t launch_draw_bin_count(uint64_t* state_dev, uint64_t seed, uint64_t philox_offset, int64_t num_queries, int64_t num_neg,
                              int32_t sm_cnt, int32_t max_tpsm, int32_t* out32, const int64_t* pos, int64_t B, const BinTable& t0,
                              const int64_t* ids1, int64_t B1, const BinTable& t1, uint32_t* err_flag, cudaStream_t st) {
    RSB_CUDA(cudaMemsetAsync(t0.cnt, 0, sizeof(uint32_t) * (size_t)t0.nbins, st));
    if (t1.nbins) RSB_CUDA(cudaMemsetAsync(t1.cnt, 0, sizeof(uint32_t) * (size_t)t1.nbins, st));
    const int64_t numel = num_queries * num_neg;
    // ATen's launch policy fixes the element <-> (thread, round, word) map (sampler.cu)
    int64_t grid = (int64_t)sm_cnt * (max_tpsm / 256);
    if (cdiv(numel, 256) < grid) grid = cdiv(numel, 256);
    const int64_t T = 256 * (grid > 0 ? grid : 1);
    const int64_t rounds = numel > 0 ? (numel - 1) / (T * 4) + 1 : 0;
    const int use_smem = t0.nbins + t1.nbins <= 8192;
    int64_t blocks = cdiv(T * rounds + B + B1, 256 * 4);
    const int64_t cap = (int64_t)sm_count() * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const size_t smem = use_smem ? sizeof(uint32_t) * (size_t)(t0.nbins + t1.nbins) : 0;
    const uint32_t range = (uint32_t)(t0.num_rows - 1);
    const uint64_t magic = (~(uint64_t)0) / (uint64_t)range + 1;
    draw_bin_count_kernel<<<(unsigned)blocks, 256, smem, st>>>(state_dev, seed, philox_offset / 4, T, rounds, numel, range, magic, out32,
                                                               pos, B, t0, ids1, B1, t1, use_smem, err_flag);
    RSB_LAUNCH_CHECK();
    if (state_dev && rounds > 0) {
        advance_philox_state_kernel<<<1, 1, 0, st>>>(state_dev, (uint64_t)(rounds * 4));
        RSB_LAUNCH_CHECK();
    }
    return 0;
}

// ------------------------------------------------------------------------------------------ BIN_SCAN
// one CTA per table: exclusive scan of the bin counts -> off[nbins + 1]; arms the append cursors (cursor[b * stride] =
// off[b]) and clears the look-back status words and the bin ticket for bin_scatter_kernel.
__global__ void __launch_bounds__(1024)
bin_scan_kernel(const BinTable t0, const BinTable t1) {
    const BinTable& t = blockIdx.x == 0 ? t0 : t1;
    __shared__ uint32_t warp_tot[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nbins = t.nbins;
    const int per = (nbins + 1023) / 1024;                 // consecutive bins per thread: one pass, one block scan
    const int b0 = threadIdx.x * per;
    uint32_t local = 0;
    for (int k = 0; k < per; ++k)
        if (b0 + k < nbins) local += t.cnt[b0 + k];
    uint32_t inc = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t x = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += x;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    uint32_t wv = warp_tot[lane];                          // 32 warps: scan their totals with one more warp scan
    uint32_t winc = wv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t x = __shfl_up_sync(kFull, winc, o);
        if (lane >= o) winc += x;
    }
    const uint32_t wbase = __shfl_sync(kFull, winc - wv, w);
    const uint32_t total = __shfl_sync(kFull, winc, 31);
    uint32_t ex = wbase + inc - local;
    for (int k = 0; k < per; ++k) {
        const int i = b0 + k;
        if (i < nbins) {
            const uint32_t c = t.cnt[i];
            t.off[i] = ex;
            t.cursor[(size_t)i * kCursorStride] = ex;
            t.status[i] = 0ull;
            ex += c;
        }
    }
    if (threadIdx.x == 0) {
        t.off[nbins] = total;
        t.totals[0] = total;
        t.ticket[0] = 0u;
    }
}

int32_t launch_bin_scan(const BinTable& t0, const BinTable& t1, cudaStream_t st) {
    bin_scan_kernel<<<t1.nbins ? 2 : 1, 1024, 0, st>>>(t0, t1);
    RSB_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------ BIN_SCATTER
// Two shapes of the kernel: A = 256 threads, chunks of 4096 entries, bins of up to 4096 rows (72 KB of shared memory, 3 CTAs/SM);
// B = 128 threads, chunks of 2048 entries, bins of up to 1024 rows (31 KB, 6 CTAs/SM): twice as many independent CTAs per SM,
// so that the barrier-separated grouping phases of one bin overlap the accumulation phases of five others.
struct BsCfgA { static constexpr int kThreads = 256, kWarps = 8, kEcap = 4096, kMaxRows = 4096; };
struct BsCfgB { static constexpr int kThreads = 128, kWarps = 4, kEcap = 2048, kMaxRows = 1024; };
constexpr uint32_t kShortSeg = 24;

template <class CF>
struct BinSmemT {
    unsigned long long stash[CF::kEcap];   // entries of the current row range, arrival order          (A: 32 KB)
    uint32_t cur[CF::kMaxRows];            // row histogram -> exclusive starts -> (after placement) row ends  (16 KB)
    uint16_t urank[CF::kMaxRows];          // rank of a row among the bin's touched rows               ( 8 KB)
    uint16_t idx[CF::kEcap];               // row-sorted position -> stash index | kLast               ( 8 KB)
    uint16_t scratch[CF::kEcap];           // permutation buffer of the long-segment sort              ( 8 KB)
    uint32_t warp_tot[CF::kWarps];
    uint32_t long_rows[CF::kEcap / kShortSeg + 8];   // rows with more than kShortSeg entries in the current range
    uint32_t bin, base, uniq, nst, nlong, ra, rb, giant;
};

// exclusive scan over the R rows of sm.cur, packed as (count | touched << 16): cur[r] <- start offset of row r,
// urank[r] <- number of touched rows before r (RANK).  Returns the packed total.  Counts must sum to < 65536.
template <class CF, bool RANK>
__device__ __forceinline__ uint32_t scan_rows(BinSmemT<CF>& sm, int R) {
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int rpt = R >= CF::kThreads ? R / CF::kThreads : 1;      // rows per thread: 1, 2, 4, 8 or 16 (R is a power of two)
    const int r0 = t * rpt;
    uint32_t c[16];                                            // fully unrolled below: stays in registers
#pragma unroll
    for (int k = 0; k < 16; k += 4) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (rpt >= 4) {
            if (k < rpt) v = *reinterpret_cast<const uint4*>(&sm.cur[r0 + k]);
        } else if (k == 0) {
            v.x = (r0 < R) ? sm.cur[r0] : 0u;
            v.y = (rpt == 2) ? sm.cur[r0 + 1] : 0u;
        }
        c[k] = v.x; c[k + 1] = v.y; c[k + 2] = v.z; c[k + 3] = v.w;
    }
    uint32_t local = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) local += c[k] + ((c[k] != 0u) ? 0x10000u : 0u);
    uint32_t inc = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) sm.warp_tot[w] = inc;
    __syncthreads();
    uint32_t wbase = 0, total = 0;
#pragma unroll
    for (int k = 0; k < CF::kWarps; ++k) {
        const uint32_t x = sm.warp_tot[k];
        if (k < w) wbase += x;
        total += x;
    }
    uint32_t ex = wbase + inc - local;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        if (k < rpt && r0 + k < R) {
            sm.cur[r0 + k] = ex & 0xFFFFu;
            if (RANK) sm.urank[r0 + k] = (uint16_t)(ex >> 16);
        }
        ex += c[k] + ((c[k] != 0u) ? 0x10000u : 0u);
    }
    __syncthreads();
    return total;
}

// ascending sort of sm.idx[lo, hi) by the 64-bit entry it points to: one thread, short segments
template <class CF>
__device__ __forceinline__ void sort_segment_small(BinSmemT<CF>& sm, uint32_t lo, uint32_t hi) {
    for (uint32_t i = lo + 1; i < hi; ++i) {
        const uint16_t xi = sm.idx[i];
        const unsigned long long key = sm.stash[xi];
        uint32_t j = i;
        while (j > lo && sm.stash[sm.idx[j - 1]] > key) { sm.idx[j] = sm.idx[j - 1]; --j; }
        sm.idx[j] = xi;
    }
}

// the same for a long segment (a hot row), by one warp: final position = number of smaller keys (ties by stash index;
// equal keys are equal entries, so their order cannot change the sum)
template <class CF>
__device__ __forceinline__ void sort_segment_warp(BinSmemT<CF>& sm, uint32_t lo, uint32_t hi, int lane) {
    const uint32_t n = hi - lo;
    for (uint32_t i = lane; i < n; i += 32) {
        const uint16_t xi = sm.idx[lo + i];
        const unsigned long long key = sm.stash[xi];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < n; ++j) {
            const uint16_t xj = sm.idx[lo + j];
            const unsigned long long kj = sm.stash[xj];
            rank += (kj < key || (kj == key && xj < xi)) ? 1u : 0u;
        }
        sm.scratch[lo + rank] = xi;
    }
    __syncwarp();
    for (uint32_t i = lane; i < n; i += 32) sm.idx[lo + i] = sm.scratch[lo + i];
    __syncwarp();
}

// Decoupled look-back by ONE WARP: sm.base <- number of touched rows in all bins before `bin`; publishes this bin's
// inclusive prefix.  The warp inspects 32 predecessors per round trip (all resident CTAs took their tickets at about the
// same time, so a one-thread walk would cross ~300 unresolved bins at one L2 round trip each).  Bins are taken in ticket
// order, so every predecessor is resident and publishes its own count without waiting for anybody: no deadlock.
template <class CF>
__device__ __forceinline__ void resolve_base(const BinScatterParams& p, BinSmemT<CF>& sm, uint32_t bin, int lane) {
    uint32_t prefix = 0;
    int64_t j = (int64_t)bin - 1;
    while (j >= 0) {
        const int64_t k = j - lane;
        // the virtual predecessor of bin 0 is an inclusive prefix of 0
        const unsigned long long v = k >= 0 ? *reinterpret_cast<volatile unsigned long long*>(p.status + k) : kStPre;
        const uint32_t flag = (uint32_t)(v >> 62);
        const uint32_t pmask = __ballot_sync(kFull, flag == 2u), zmask = __ballot_sync(kFull, flag == 0u);
        const int first_p = pmask ? __ffs(pmask) - 1 : 32;
        const uint32_t need = first_p >= 31 ? 0xFFFFFFFFu : ((2u << first_p) - 1u);   // lanes 0 .. first_p
        if (zmask & need) continue;                                   // somebody in the window has not published yet
        uint32_t x = (lane <= first_p) ? (uint32_t)v : 0u;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
        prefix += x;
        if (first_p < 32) break;
        j -= 32;
    }
    if (lane == 0) {
        if (bin != 0) *reinterpret_cast<volatile unsigned long long*>(p.status + bin) = kStPre | (unsigned long long)(prefix + sm.uniq);
        if (bin == (uint32_t)p.nbins - 1) p.totals[1] = prefix + sm.uniq;
        sm.base = prefix;
    }
}

// write (or apply) the finished gradient of local row lr
template <class CF, int VPL, bool FULL, int OPT>
__device__ __forceinline__ void flush_row(const BinScatterParams& p, const BinSmemT<CF>& sm, float4 (&acc)[VPL], float& csum, uint32_t lr,
                                          uint32_t row0, int lane, const bool (&act)[VPL]) {
    const int D = p.D;
    const uint32_t row = row0 + lr;
    const size_t orow = p.dense ? (size_t)row : (size_t)sm.base + sm.urank[lr];
#pragma unroll
    for (int x = 0; x < VPL; ++x) {
        if (FULL || act[x]) {
            const int col = lane * 4 + x * 128;
            float4 a = acc[x];
            if (p.euclid) {
                const float4 wv = ldg128(p.w + (size_t)row * D + col);
                a.x = 2.f * (a.x - csum * wv.x); a.y = 2.f * (a.y - csum * wv.y);
                a.z = 2.f * (a.z - csum * wv.z); a.w = 2.f * (a.w - csum * wv.w);
            }
            if (OPT >= 0) {
                const OptParams o = {p.lr, p.b1, p.b2, p.eps, p.step_size};
                opt_update4<(OPT >= 0 ? OPT : 0)>(p.w_rw, p.s1, p.s2, (size_t)row * D + col, a, o);
            } else {
                float* dst = p.vals + orow * D + col;
                if (p.accumulate) {
                    const float4 o = *reinterpret_cast<const float4*>(dst);
                    a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
                }
                stg128_stream(dst, a);
            }
        }
        acc[x] = make_float4(0, 0, 0, 0);
    }
    csum = 0.f;
}

// stash[0, m) holds the entries of rows [ra, rb) and cur[] their per-row counts: group, order and sum them
// (WHOLE: the range is the whole bin, cur already holds the row starts, and the touched rows are reported here)
// PLAIN: compact sink, inner product, overwrite, no optimizer -- the touched rows of a range have consecutive ranks, so
// the output is a running pointer and a finished row costs one 16-byte store per lane.
template <class CF, int VPL, bool FULL, int OPT, bool WHOLE, bool PLAIN, int DEPTH>
__device__ __forceinline__ void process_range(const BinScatterParams& p, BinSmemT<CF>& sm, uint32_t m, uint32_t row0, float gs,
                                              const bool (&act)[VPL]) {
    // query rows in flight per warp.  The loop is latency-bound (entry -> query row from L2 -> FMA, ~1.2 us per
    // round trip): throughput = rows in flight per SM / latency, so registers are spent on depth (2 CTAs/SM, 128 regs)
    constexpr int UNR = VPL == 1 ? DEPTH : (VPL == 2 ? DEPTH / 2 : DEPTH / 4);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int D = p.D;
    const int R = 1 << p.shift;
    const uint32_t rmask = (uint32_t)R - 1u, bmask = (1u << p.bbits) - 1u;
    if (!WHOLE) scan_rows<CF, false>(sm, R);                            // counts -> starts
    for (uint32_t i = t; i < m; i += CF::kThreads) {                  // placement: sorted position -> stash index
        const uint32_t lr = ((uint32_t)sm.stash[i] >> p.bbits) & rmask;
        sm.idx[atomicAdd(&sm.cur[lr], 1u)] = (uint16_t)i;
    }
    if (t == 0) sm.nlong = 0u;
    __syncthreads();
    // per row: order its entries by their 64-bit encoding -- the arrival order of the forward kernel's atomics must not
    // reach the floating-point sums -- and flag its last entry
    for (int r = t; r < R; r += CF::kThreads) {
        const uint32_t e1 = sm.cur[r], e0 = r ? sm.cur[r - 1] : 0u;
        if (e1 != e0) {
            if (e1 - e0 == 2u) {                                      // the common multi-entry case: one compare-exchange
                const uint16_t a = sm.idx[e0], b = sm.idx[e0 + 1];
                if (sm.stash[a] > sm.stash[b]) { sm.idx[e0] = b; sm.idx[e0 + 1] = a; }
            } else if (e1 - e0 > kShortSeg) sm.long_rows[atomicAdd(&sm.nlong, 1u)] = (uint32_t)r;
            else if (e1 - e0 > 2u) sort_segment_small<CF>(sm, e0, e1);
        }
    }
    if (WHOLE && warp == 0) {                                       // as late as possible: predecessors have published by now
        __syncwarp();
        resolve_base<CF>(p, sm, sm.bin, lane);
    }
    __syncthreads();
    for (uint32_t k = warp; k < sm.nlong; k += CF::kWarps) {
        const uint32_t r = sm.long_rows[k];
        sort_segment_warp<CF>(sm, r ? sm.cur[r - 1] : 0u, sm.cur[r], lane);
    }
    __syncthreads();
    for (int r = t; r < R; r += CF::kThreads) {
        const uint32_t e1 = sm.cur[r], e0 = r ? sm.cur[r - 1] : 0u;
        if (e1 != e0) {
            sm.idx[e1 - 1] |= kLast;
            if (WHOLE && p.rows_out && (int64_t)sm.base + sm.urank[r] < p.cap)
                p.rows_out[(size_t)sm.base + sm.urank[r]] = (int64_t)(row0 + r);
        }
    }
    __syncthreads();
    // warp w sums the sorted positions [bound(w), bound(w + 1)): equal shares of the range snapped to row ends
    auto bound = [&](int w) -> uint32_t {
        if (w >= CF::kWarps) return m;
        uint32_t b = (uint32_t)(((uint64_t)m * (uint32_t)w) / CF::kWarps);
        if (b > 0) b = sm.cur[((uint32_t)sm.stash[sm.idx[b - 1] & 0x7FFFu] >> p.bbits) & rmask];
        return b;
    };
    const uint32_t p0 = bound(warp), p1 = bound(warp + 1);
    const char* src_lane = reinterpret_cast<const char*>(p.src) + lane * 16;   // 32-bit byte offsets below: B * D * 4 < 2^32
    const uint32_t row_bytes = (uint32_t)D * 4u;
    float4 acc[VPL];
#pragma unroll
    for (int x = 0; x < VPL; ++x) acc[x] = make_float4(0, 0, 0, 0);
    float csum = 0.f;
    float* dst_run = nullptr;                                          // PLAIN: output row of the current (unfinished) row
    if (PLAIN && p0 < p1) {
        const uint32_t lr0 = ((uint32_t)sm.stash[sm.idx[p0] & 0x7FFFu] >> p.bbits) & rmask;
        dst_run = p.vals + ((size_t)sm.base + sm.urank[lr0]) * D + lane * 4;
    }
    for (uint32_t q0 = p0; q0 < p1; q0 += 32) {
        // lane l decodes entry q0 + l of the sorted list (coefficient incl. the softmax denominator and the upstream gradient)
        uint32_t bq = 0u, meta = 0u;                                   // meta = local row | last-of-row << 31
        float c = 0.f;
        if (q0 + lane < p1) {
            const uint32_t ix = sm.idx[q0 + lane];
            const unsigned long long e = sm.stash[ix & 0x7FFFu];
            const uint32_t lo = (uint32_t)e;
            const float val = __uint_as_float((uint32_t)(e >> 32));
            bq = (lo & bmask) * row_bytes;                             // byte offset of the query row
            meta = ((lo >> p.bbits) & rmask) | ((ix & kLast) ? 0x80000000u : 0u);
            c = ((lo & kDirect) ? val : expf(val - __ldg(p.lse + (lo & bmask))) * p.ssm_scale) * gs;
        }
        const uint32_t lastmask = __ballot_sync(kFull, (meta >> 31) != 0u);
        const int cnt = (int)min(32u, p1 - q0);
        for (int t0 = 0; t0 < cnt; t0 += UNR) {                        // lanes >= cnt hold (offset 0, c = 0): harmless
            float4 v[UNR][VPL];
            float cc[UNR];
#pragma unroll
            for (int k = 0; k < UNR; ++k) {
                const uint32_t off = __shfl_sync(kFull, bq, t0 + k);
                cc[k] = __shfl_sync(kFull, c, t0 + k);
                const float* srow = reinterpret_cast<const float*>(src_lane + off);
#pragma unroll
                for (int x = 0; x < VPL; ++x) v[k][x] = (FULL || act[x]) ? ldg128(srow + x * 128) : make_float4(0, 0, 0, 0);
            }
            const uint32_t lm = lastmask >> t0;
#pragma unroll
            for (int k = 0; k < UNR; ++k) {
#pragma unroll
                for (int x = 0; x < VPL; ++x) fma4(acc[x], cc[k], v[k][x]);
                if (!PLAIN) csum += cc[k];
                if ((lm >> k) & 1u) {                                  // warp-uniform: this row is complete
                    if (PLAIN) {
#pragma unroll
                        for (int x = 0; x < VPL; ++x) {
                            if (FULL || act[x]) stg128_stream(dst_run + x * 128, acc[x]);
                            acc[x] = make_float4(0, 0, 0, 0);
                        }
                        dst_run += D;
                    } else {
                        flush_row<CF, VPL, FULL, OPT>(p, sm, acc, csum, __shfl_sync(kFull, meta, t0 + k) & 0x7FFFFFFFu, row0, lane, act);
                    }
                }
            }
        }
    }
    __syncthreads();
}

// a single row with more than CF::kEcap entries: all warps stream the bin, sum the entries of row `lr` (arrival order),
// the eight partial sums are added in warp order
template <class CF, int VPL, bool FULL, int OPT>
__device__ __forceinline__ void process_giant_row(const BinScatterParams& p, BinSmemT<CF>& sm, uint32_t beg, uint32_t cnt, uint32_t lr,
                                                  uint32_t row0, float gs, const bool (&act)[VPL]) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int D = p.D;
    const uint32_t rmask = (1u << p.shift) - 1u, bmask = (1u << p.bbits) - 1u;
    const float* src_lane = p.src + lane * 4;
    float4 acc[VPL];
#pragma unroll
    for (int x = 0; x < VPL; ++x) acc[x] = make_float4(0, 0, 0, 0);
    float csum = 0.f;
    for (uint32_t i0 = warp * 32; i0 < cnt; i0 += CF::kThreads) {
        const uint32_t i = i0 + lane;
        unsigned long long e = 0ull;
        bool mine = false;
        if (i < cnt) {
            e = __ldg(reinterpret_cast<const unsigned long long*>(p.ent) + beg + i);
            mine = (((uint32_t)e >> p.bbits) & rmask) == lr;
        }
        const uint32_t lo = (uint32_t)e;
        const uint32_t bq = lo & bmask;
        float c = 0.f;
        if (mine) c = ((lo & kDirect) ? __uint_as_float((uint32_t)(e >> 32))
                                      : expf(__uint_as_float((uint32_t)(e >> 32)) - __ldg(p.lse + bq)) * p.ssm_scale) * gs;
        uint32_t mask = __ballot_sync(kFull, mine);
        while (mask) {
            const int l = __ffs(mask) - 1;
            mask &= mask - 1u;
            const uint32_t bb = __shfl_sync(kFull, bq, l);
            const float cl = __shfl_sync(kFull, c, l);
            const float* srow = src_lane + (size_t)bb * D;
#pragma unroll
            for (int x = 0; x < VPL; ++x)
                if (FULL || act[x]) fma4(acc[x], cl, ldg128(srow + x * 128));
            csum += cl;
        }
    }
    float* part = reinterpret_cast<float*>(sm.stash);                 // [CF::kWarps][D + 1]
#pragma unroll
    for (int x = 0; x < VPL; ++x)
        if (FULL || act[x]) *reinterpret_cast<float4*>(part + (size_t)warp * 516 + lane * 4 + x * 128) = acc[x];
    if (lane == 0) part[(size_t)warp * 516 + 512] = csum;
    __syncthreads();
    if (warp == 0) {
        csum = 0.f;
#pragma unroll
        for (int x = 0; x < VPL; ++x) acc[x] = make_float4(0, 0, 0, 0);
        for (int w = 0; w < CF::kWarps; ++w) {
#pragma unroll
            for (int x = 0; x < VPL; ++x)
                if (FULL || act[x]) {
                    const float4 a = *reinterpret_cast<const float4*>(part + (size_t)w * 516 + lane * 4 + x * 128);
                    acc[x].x += a.x; acc[x].y += a.y; acc[x].z += a.z; acc[x].w += a.w;
                }
            csum += part[(size_t)w * 516 + 512];
        }
        flush_row<CF, VPL, FULL, OPT>(p, sm, acc, csum, lr, row0, lane, act);
    }
    __syncthreads();
}

template <class CF, int VPL, bool FULL, int OPT, bool PLAIN, int DEPTH, int OCC>
__global__ void __launch_bounds__(CF::kThreads, OCC)
bin_scatter_kernel(const BinScatterParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BinSmemT<CF>& sm = *reinterpret_cast<BinSmemT<CF>*>(smem_raw);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int R = 1 << p.shift;
    const uint32_t rmask = (uint32_t)R - 1u;
    const float gs = p.gscale ? __ldg(p.gscale) : 1.0f;
    bool act[VPL];
#pragma unroll
    for (int x = 0; x < VPL; ++x) act[x] = FULL || (lane * 4 + x * 128) < p.D;
    constexpr int UNR = 4;
    uint32_t* gcount = p.heavy_counts + (size_t)blockIdx.x * CF::kMaxRows;   // whole-bin row counts of a heavy bin

    for (;;) {
        __syncthreads();                                              // sm.bin of the previous bin is no longer read
        if (t == 0) sm.bin = atomicAdd(p.ticket, 1u);
        __syncthreads();
        const uint32_t bin = sm.bin;
        if (bin >= (uint32_t)p.nbins) break;
        const uint32_t beg = __ldg(p.bin_off + bin), end = __ldg(p.bin_off + bin + 1);
        const uint32_t cnt = end - beg;
        const uint32_t row0 = bin << p.shift;
        const bool heavy = cnt > (uint32_t)CF::kEcap;

        // ---- whole-bin pass: per-row touch counts (and, for bins that fit one chunk, the entries themselves)
        for (int r = t * 4; r < R; r += CF::kThreads * 4) *reinterpret_cast<uint4*>(&sm.cur[r]) = make_uint4(0, 0, 0, 0);
        __syncthreads();
        for (uint32_t i0 = t; i0 < cnt; i0 += CF::kThreads * UNR) {
            unsigned long long e[UNR];
#pragma unroll
            for (int k = 0; k < UNR; ++k) {
                const uint32_t i = i0 + k * CF::kThreads;
                e[k] = i < cnt ? __ldcs(reinterpret_cast<const unsigned long long*>(p.ent) + beg + i) : 0ull;
            }
#pragma unroll
            for (int k = 0; k < UNR; ++k) {
                const uint32_t i = i0 + k * CF::kThreads;
                if (i < cnt) {
                    if (!heavy) sm.stash[i] = e[k];
                    atomicAdd(&sm.cur[((uint32_t)e[k] >> p.bbits) & rmask], 1u);
                }
            }
        }
        __syncthreads();
        if (heavy) {                                                  // keep the counts, scan the presence flags
            for (int r = t; r < R; r += CF::kThreads) {
                const uint32_t c = sm.cur[r];
                gcount[r] = c;
                sm.cur[r] = c ? 1u : 0u;
            }
            __syncthreads();
        }
        // ---- rank of every touched row inside the bin; publish the bin's unique-row count, then look back for the
        //      rank of its first row among all touched rows of the table (decoupled look-back over bins in ticket order)
        const uint32_t tot = scan_rows<CF, true>(sm, R);                  // heavy: cur is scratch after this
        if (t == 0) {
            sm.uniq = tot >> 16;
            *reinterpret_cast<volatile unsigned long long*>(p.status + bin) = (bin == 0 ? kStPre : kStAgg) | (unsigned long long)sm.uniq;
        }
        if (warp == 0 && (cnt == 0 || heavy)) {                       // else: resolved late, inside process_range
            __syncwarp();
            resolve_base<CF>(p, sm, bin, lane);
        }
        if (cnt == 0) continue;
        if (!heavy) {                                                 // cur already holds the row starts
            process_range<CF, VPL, FULL, OPT, true, PLAIN, DEPTH>(p, sm, cnt, row0, gs, act);
            continue;
        }
        __syncthreads();                                              // sm.base
        // ---- heavy bin: consecutive row ranges of <= CF::kEcap entries, each gathered from the bin's list by its own pass
        if (p.rows_out) {
            for (int r = t; r < R; r += CF::kThreads)
                if (gcount[r] && (int64_t)sm.base + sm.urank[r] < p.cap) p.rows_out[(size_t)sm.base + sm.urank[r]] = (int64_t)(row0 + r);
        }
        uint32_t ra = 0;
        while (ra < (uint32_t)R) {
            if (warp == 0) {                                          // longest range [ra, rb) with <= CF::kEcap entries
                uint32_t run = 0, rb = ra;
                bool done = false;
                while (!done && rb < (uint32_t)R) {
                    const uint32_t r = rb + lane;
                    const uint32_t c = r < (uint32_t)R ? gcount[r] : 0u;
                    uint32_t inc = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t v = __shfl_up_sync(kFull, inc, o);
                        if (lane >= o) inc += v;
                    }
                    const uint32_t over = __ballot_sync(kFull, run + inc > (uint32_t)CF::kEcap);
                    if (over) {
                        const int l = __ffs(over) - 1;                // first row that does not fit
                        rb += l;
                        run += l ? __shfl_sync(kFull, inc, l - 1) : 0u;
                        done = true;
                    } else {
                        run += __shfl_sync(kFull, inc, 31);
                        rb += 32;
                    }
                }
                if (rb > (uint32_t)R) rb = R;
                if (lane == 0) {
                    sm.giant = (rb == ra) ? 1u : 0u;                  // row ra alone exceeds a chunk
                    sm.ra = ra; sm.rb = (rb == ra) ? ra + 1 : rb; sm.nst = 0u;
                }
            }
            for (int r = t * 4; r < R; r += CF::kThreads * 4) *reinterpret_cast<uint4*>(&sm.cur[r]) = make_uint4(0, 0, 0, 0);
            __syncthreads();
            const uint32_t rb = sm.rb;
            if (sm.giant) {
                process_giant_row<CF, VPL, FULL, OPT>(p, sm, beg, cnt, ra, row0, gs, act);
            } else {
                for (uint32_t i = t; i < cnt; i += CF::kThreads) {
                    const unsigned long long e = __ldg(reinterpret_cast<const unsigned long long*>(p.ent) + beg + i);
                    const uint32_t lr = ((uint32_t)e >> p.bbits) & rmask;
                    if (lr >= ra && lr < rb) {
                        sm.stash[atomicAdd(&sm.nst, 1u)] = e;
                        atomicAdd(&sm.cur[lr], 1u);
                    }
                }
                __syncthreads();
                const uint32_t m = sm.nst;
                if (m) process_range<CF, VPL, FULL, OPT, false, PLAIN, DEPTH>(p, sm, m, row0, gs, act);
            }
            __syncthreads();
            ra = rb;
        }
    }
}

int64_t bin_scatter_grid() { return (int64_t)sm_count() * 6; }          // upper bound over the configurations below (sizes bin_heavy)

template <class CF, int VPL, bool FULL, int OPT, bool PLAIN, int DEPTH, int OCC>
static int32_t launch_bs(const BinScatterParams& p, cudaStream_t st) {
    RSB_REQUIRE((1 << p.shift) <= CF::kMaxRows, RSB200_EINVAL, "bin_shift %d exceeds the %d rows per bin of this kernel shape", p.shift, CF::kMaxRows);
    const size_t smem = sizeof(BinSmemT<CF>);
    int64_t blocks = (int64_t)sm_count() * OCC;
    if (blocks > p.nbins) blocks = p.nbins;
    if (blocks < 1) blocks = 1;
    RSB_CUDA(cudaFuncSetAttribute(bin_scatter_kernel<CF, VPL, FULL, OPT, PLAIN, DEPTH, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bin_scatter_kernel<CF, VPL, FULL, OPT, PLAIN, DEPTH, OCC><<<(unsigned)blocks, CF::kThreads, smem, st>>>(p);
    RSB_LAUNCH_CHECK();
    return 0;
}

template <int VPL>
static int32_t launch_bin_scatter_v(const BinScatterParams& p, cudaStream_t st) {
    using A = BsCfgA;
    const bool full = p.D == 128 * VPL;
    const bool plain = !p.dense && !p.accumulate && !p.euclid && p.opt < 0;
    if (plain && full) {
        if constexpr (VPL == 1) {                                     // the hot configuration; p.tune: A/B of shape x depth x occupancy
            if (p.tune == 1) return launch_bs<A, VPL, true, -1, true, 16, 2>(p, st);
            if (p.tune == 2) return launch_bs<A, VPL, true, -1, true, 8, 2>(p, st);
            if (p.tune == 3) return launch_bs<A, VPL, true, -1, true, 4, 3>(p, st);
            if (p.tune == 4 && p.shift <= 10) return launch_bs<BsCfgB, VPL, true, -1, true, 8, 6>(p, st);
            if (p.tune == 5 && p.shift <= 10) return launch_bs<BsCfgB, VPL, true, -1, true, 4, 6>(p, st);
        }
        return launch_bs<A, VPL, true, -1, true, kBsDepth, kBsOcc>(p, st);
    }
    if (plain) return launch_bs<A, VPL, false, -1, true, kBsDepth, kBsOcc>(p, st);
    if (p.opt == 0) return full ? launch_bs<A, VPL, true, 0, false, kBsDepth, kBsOcc>(p, st) : launch_bs<A, VPL, false, 0, false, kBsDepth, kBsOcc>(p, st);
    if (p.opt == 1) return full ? launch_bs<A, VPL, true, 1, false, kBsDepth, kBsOcc>(p, st) : launch_bs<A, VPL, false, 1, false, kBsDepth, kBsOcc>(p, st);
    if (p.opt == 2) return full ? launch_bs<A, VPL, true, 2, false, kBsDepth, kBsOcc>(p, st) : launch_bs<A, VPL, false, 2, false, kBsDepth, kBsOcc>(p, st);
    return full ? launch_bs<A, VPL, true, -1, false, kBsDepth, kBsOcc>(p, st) : launch_bs<A, VPL, false, -1, false, kBsDepth, kBsOcc>(p, st);
}

int32_t launch_bin_scatter(const BinScatterParams& p, cudaStream_t st) {
    if (p.nbins <= 0) return 0;
    if (p.D <= 128) return launch_bin_scatter_v<1>(p, st);
    if (p.D <= 256) return launch_bin_scatter_v<2>(p, st);
    if (p.D <= 512) return launch_bin_scatter_v<4>(p, st);
    set_error("embedding dim %d > 512 is not supported", p.D);
    return RSB200_EUNSUPPORTED;
}

}  // namespace rsb
