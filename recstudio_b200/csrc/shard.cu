// shard.cu -- owner-compute training step of the ROW-SHARDED item table ("ship
// queries, not rows").
//
// The reference has no table sharding (its only multi-GPU mode replicates every parameter each
// step, recstudio/utils/data_parallel.py:151).  Row lookup over NVLink (sharded.py: ShardedRows)
// moves 512 B per touched row each way; here nothing row-sized ever leaves its owner:
//
//   every owner sees the GLOBAL batch (G = world x B queries: query vectors, positive ids, negative
//   ids -- all-gathered by the caller) and does the part of the reference step
//   (baseretriever.py:142-176,399-404 + loss.backward(), recommender.py:638) that touches ITS rows:
//
//   PREP    per query: keep the negatives whose row this owner holds (stable ballot compaction to
//           LOCAL row ids), count touches per local row (the COUNT phase of group.cu fused in),
//           score the positive if it is owned;  then the usual exclusive scan -> CSR offsets.
//           --> caller: all-reduce(SUM) of sp[G] (every positive has exactly one owner)
//   FWD     pair_fwd_kernel<PARTIAL>: the unchanged streaming gather/score/loss loop over the owned
//           negatives; leaves per query the raw dq accumulator and {sum c, sum loss} (BPR) or the
//           online-softmax state {m, l} (SampledSoftmax) in stats_all[rank].
//           --> caller: all-gather of the [G, 2] stats slices
//   FINISH  per query: merge the owners' stats (same arithmetic on every owner, in owner order ->
//           identical lse / loss everywhere), turn the raw accumulator into this owner's share of
//           d loss / d query, add the positive's term and emit the positive's gradient entry if owned.
//           --> caller: all-reduce(SUM) of dq[G, d]
//   SCATTER the unchanged segmented scatter over the owner's rows (scatter.cu): gradient rows of
//           OWNED rows only, so dV never crosses NVLink.
//
// Bytes on the wire per rank and step: ids (4 B per negative) + 8 B + 4 B + d*4 B per query, instead
// of 2 x 512 B per touched row.
#include "common.cuh"
#include "kernels.h"

namespace rsb {

constexpr int kPrepWarps = 8;

__device__ __forceinline__ float warp_sum_s(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

struct PrepParams {
    const float* w_local; const float* q_all;
    const int64_t* pos; const int32_t* neg; const float* logq_neg;
    uint32_t* cnt;                 // [local_rows + 1] zeroed histogram
    int32_t* neg_c; uint32_t* slot_neg; float* lq_c; int32_t* ncount;
    int32_t* pos_local; uint32_t* slot_pos; float* sp;
    uint32_t* err;
    int64_t num_items, row0, local_rows;
    int G, n, D, euclid;
    // owner-side regeneration of the uniform draw (regen_state != null): ids are recomputed, `neg` is not read
    const uint64_t* regen_state;   // [world, 2] (seed, philox offset)
    int64_t regen_T;               // 256 * grid of the ATen kernel for numel = regen_B * n
    int regen_B;
};

// id of element li of the draw torch.randint(1, num_items, (B, n)) for generator state (seed, offset): ATen thread
// idx = li mod T draws curand4 in round (li / T) / 4 and hands word (li / T) % 4 to this element (sampler.cu).
__device__ __forceinline__ int32_t regen_uniform_id(uint64_t seed, uint64_t offset, int64_t li, int64_t T, uint32_t range) {
    const int64_t q = li / T, idx = li - q * T;
    const uint4 w = Philox::gen(seed, (uint64_t)idx, offset / 4 + (uint64_t)(q >> 2));
    const int ii = (int)(q & 3);
    const uint32_t word = ii == 0 ? w.x : (ii == 1 ? w.y : (ii == 2 ? w.z : w.w));
    return (int32_t)(word % range + 1u);
}

// one warp per query
__global__ void __launch_bounds__(kPrepWarps * 32)
shard_prep_kernel(const PrepParams p) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * kPrepWarps + (threadIdx.x >> 5);
    if (b >= p.G) return;
    const int D = p.D;

    // ---- positive: owned? -> score + slot ---------------------------------------------------
    int64_t gp = __ldg(p.pos + b);
    if (gp < 0 || gp >= p.num_items) { if (lane == 0) atomicOr(p.err, 1u); gp = 0; }
    const int64_t lp = gp - p.row0;
    const bool own = lp >= 0 && lp < p.local_rows;        // warp-uniform
    float sp = 0.f;
    if (own) {
        for (int c = lane * 4; c < D; c += 128) {
            const float4 qv = ldg128(p.q_all + (size_t)b * D + c);
            const float4 vv = ldg128(p.w_local + (size_t)lp * D + c);
            sp += p.euclid ? sqdist4(qv, vv) : dot4(qv, vv);
        }
        sp = warp_sum_s(sp);
        if (p.euclid) sp = -sp;
    }
    if (lane == 0) {
        p.sp[b] = own ? sp : 0.f;
        p.pos_local[b] = own ? (int32_t)lp : -1;
        p.slot_pos[b] = (own && gp != 0) ? atomicAdd(p.cnt + lp, 1u) : kNoSlot;     // padding row: no gradient
    }

    // ---- negatives: stable compaction of the owned ids + per-row slot --------------------------
    const size_t base = (size_t)b * p.n;
    int kept = 0;
    bool bad = false;
    uint64_t rg_seed = 0, rg_off = 0;
    int64_t rg_base = 0;
    if (p.regen_state) {
        const int r = b / p.regen_B;
        rg_seed = __ldg(p.regen_state + 2 * r);
        rg_off = __ldg(p.regen_state + 2 * r + 1);
        rg_base = (int64_t)(b - r * p.regen_B) * p.n;        // first element of this query inside rank r's draw
    }
    // kU batches of 32 ids per iteration: the loads, then the atomics, of kU batches are independent and in
    // flight together (the one-batch loop was latency-bound: 0.37 ms for 67 M ids at 8 owners)
    constexpr int kU = 4;
    for (int jb = 0; jb < p.n; jb += 32 * kU) {
        int32_t raw[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int j = jb + u * 32 + lane;
            if (p.regen_state) raw[u] = (j < p.n) ? regen_uniform_id(rg_seed, rg_off, rg_base + j, p.regen_T, (uint32_t)(p.num_items - 1)) : -1;
            else raw[u] = (j < p.n) ? __ldg(p.neg + base + j) : -1;
        }
        int kpos[kU]; int64_t lids[kU]; bool mines[kU]; bool zero[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int j = jb + u * 32 + lane;
            const bool valid = j < p.n;
            int64_t gid = raw[u];
            if (valid && (gid < 0 || gid >= p.num_items)) { bad = true; gid = 0; }
            const int64_t lid = gid - p.row0;
            const bool mine = valid && lid >= 0 && lid < p.local_rows;
            const uint32_t mask = __ballot_sync(kFull, mine);
            kpos[u] = kept + __popc(mask & ((1u << lane) - 1u));
            kept += __popc(mask);
            lids[u] = lid; mines[u] = mine; zero[u] = gid == 0;
        }
        uint32_t sl[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) sl[u] = (mines[u] && !zero[u]) ? atomicAdd(p.cnt + lids[u], 1u) : kNoSlot;
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            if (mines[u]) {
                p.neg_c[base + kpos[u]] = (int32_t)lids[u];
                p.slot_neg[base + kpos[u]] = sl[u];
                if (p.lq_c) p.lq_c[base + kpos[u]] = __ldg(p.logq_neg + base + jb + u * 32 + lane);
            }
        }
    }
    if (bad) atomicOr(p.err, 1u);
    if (lane == 0) p.ncount[b] = kept;
}

// slot -> absolute entry position for the compacted lists (same purpose as group.cu resolve_kernel): the offsets are
// looked up here, right after the scan wrote them, instead of inside the forward kernel's row stream.
__global__ void __launch_bounds__(kPrepWarps * 32)
shard_resolve_kernel(const int32_t* __restrict__ neg_c, uint32_t* __restrict__ slot_neg, const int32_t* __restrict__ ncount,
                     const uint32_t* __restrict__ off, int G, int n) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * kPrepWarps + (threadIdx.x >> 5);
    if (b >= G) return;
    const size_t base = (size_t)b * n;
    const int cnt = min(n, max(0, __ldg(ncount + b)));
    for (int j = lane; j < cnt; j += 32) {
        const uint32_t s = slot_neg[base + j];
        if (s != kNoSlot) slot_neg[base + j] = __ldg(off + neg_c[base + j]) + s;
    }
}

// ------------------------------------------------------------------------------------------ PREP, binned grouping
// One warp per query: candidates -> owned? -> stable compaction to LOCAL ids (+ log Q) + per-bin touch histogram
// (bins.cu; shared-memory aggregated, persistent CTAs).  Where do the candidates come from (MODE):
//   0  neg[G, n] global ids all-gathered by the caller;
//   1  owner-side regeneration of every rank's UniformSampler draw:  torch.randint(1, N, (B, n)) of rank r is
//      element li = lb * n + j -> ATen thread idx = li mod T, round (li / T) / 4, word (li / T) mod 4.  With T = t * n
//      the four words of one Philox call belong to the SAME position j of four queries t apart, so a CTA of four warps
//      takes the queries {4 t R + x + t ii, ii = 0..3} of one rank: every thread computes whole Philox blocks, the ids go
//      through shared memory, and no random bit is computed twice (the per-query formulation recomputed each block 4x);
//   2  the same for PopularSamplerModel (sampler.py:243-258): u = torch.rand(B, n) of rank r, id = searchsorted(table, u).
//      An owner only needs ITS slice of the cumulative table: id lands in [row0, row0 + L) iff cdf_lo < u <= cdf_hi
//      (cdf_lo = table[row0 - 1], cdf_hi = table[row0 + L - 1]; -inf / +inf at the ends), and the search runs inside the
//      slice (guide table + short bisection), log Q = log(pop_prob_local[id]).
struct PrepBinParams {
    const float* w_local; const float* q_all; const int64_t* pos;
    const int32_t* neg; const float* logq_neg;
    int32_t* neg_c; float* lq_c; int32_t* ncount; int32_t* pos_local; float* sp; float* lq_pos_out;
    uint32_t* err;
    BinTable bt;
    int64_t num_items, row0, local_rows;
    int G, n, D, euclid, use_smem;
    uint64_t mod_magic;            // ceil(2^64 / (num_items - 1)): division-free v mod (num_items - 1)
    uint32_t* work;                // PREP's work counter: bin_cnt[nbins] (zeroed with the counts)
    uint64_t own_lo, own_hi;       // uniform regeneration: word v lands on a row of this owner  <=>  own_lo <= mod_magic * v <= own_hi
    int do_pos, do_neg;            // PREP may be split: negatives (independent of the batch's queries) | positives
    // regeneration
    const uint64_t* regen_state; int regen_B; int64_t regen_T; int t_per; int n_round_blocks;
    const float* pop_table; const float* pop_prob; const int32_t* pop_guide; int pop_bits; int64_t pop_k0;
    float cdf_lo, cdf_hi;
};

__device__ __forceinline__ int local_lower_bound(const float* __restrict__ table, int lo, int hi, float u) {
    // first i in [lo, hi] with table[i] >= u (hi if none): bisection down to <= 8 entries, then independent loads
    constexpr int kLinear = 8;
    while (hi - lo > kLinear) {
        const int mid = lo + ((hi - lo) >> 1);
        if (__ldg(table + mid) < u) lo = mid + 1; else hi = mid;
    }
    int id = lo;
#pragma unroll
    for (int t = 0; t < kLinear; ++t)
        if (lo + t < hi) id += (__ldg(table + lo + t) < u) ? 1 : 0;
    return id;
}

constexpr int kPrepSeg = 1024;                       // candidates of one query staged per pass
constexpr int kPrepSegWords = (kPrepSeg / 32) * 33;  // lane-major staging with a 33-word pitch (conflict-free both ways)

// One CTA = four warps = four queries.  Per segment of <= 1024 candidates of each query:
//   A  stage the candidates in shared memory (ids from global memory, or random words from Philox -- in the shared-block
//      formulation every thread computes whole Philox blocks and feeds all four queries);
//   B  every LANE walks its own contiguous chunk of the segment, keeps what this owner holds (compacted in place,
//      lane-private column), then one warp scan turns the per-lane counts into offsets: the compaction stays stable
//      (query order) without a ballot per 32 candidates, and all 32 lanes test candidates all the time;
//   C  every lane resolves and writes its kept candidates (the popularity search runs here, on owned draws only).
template <int MODE>
__global__ void __launch_bounds__(128, 6)
shard_prep_bins_kernel(const PrepBinParams p) {
    extern __shared__ uint32_t s_dyn[];
    uint32_t* s_hist = s_dyn;                                         // [nbins] (use_smem)
    uint32_t* s_cand = s_dyn + (p.use_smem ? p.bt.nbins : 0);         // [4][kPrepSegWords]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = p.D, n = p.n;
    if (p.use_smem) {
        for (int i = threadIdx.x; i < p.bt.nbins; i += blockDim.x) s_hist[i] = 0u;
        __syncthreads();
    }
    uint32_t* hist = p.use_smem ? s_hist : p.bt.cnt;
    uint32_t* my = s_cand + warp * kPrepSegWords;
    bool bad = false;
    // shared Philox blocks need T = t_per * n; any other shape regenerates per id (t_per = 0: every block computed 4x)
    const bool shared = MODE != 0 && p.t_per > 0;
    const int64_t nitems = !shared ? ((int64_t)p.G + 3) / 4 : (int64_t)(p.G / p.regen_B) * p.n_round_blocks * p.t_per;
    // Work distribution: every CTA's first item is its block index; with negatives to prepare, the following ones come from a
    // counter (p.work, zeroed by the launcher), fetched one item ahead so the atomic's latency is off the path.  A CTA that
    // becomes resident late (the SMs are shared with an exchange kernel in the look-ahead schedule) then only takes what is
    // left, instead of a fixed 1/gridDim share that would stretch the kernel to twice its length.
    const bool dynamic = p.do_neg && n > 0 && p.work != nullptr;
    __shared__ uint32_t s_next;
    int64_t item = blockIdx.x;
    while (item < nitems) {
        uint32_t nxt = 0;
        if (dynamic && threadIdx.x == 0) nxt = gridDim.x + atomicAdd(p.work, 1u);
        int g;                                                        // global query of this warp, -1 = none
        uint64_t seed = 0, off = 0;
        int64_t li0 = 0;                                              // per-id regeneration: first element of the query
        int R = 0, x = 0;
        if (!shared) {
            g = (int)(item * 4 + warp);
            if (g >= p.G) g = -1;
            if (MODE != 0 && g >= 0) {
                const int r = g / p.regen_B;
                seed = __ldg(p.regen_state + 2 * r); off = __ldg(p.regen_state + 2 * r + 1);
                li0 = (int64_t)(g - r * p.regen_B) * n;
            }
        } else {
            const uint32_t per_rank = (uint32_t)(p.n_round_blocks * p.t_per), it32 = (uint32_t)item;     // nitems < 2^31
            const int r = (int)(it32 / per_rank), rem = (int)(it32 - (uint32_t)r * per_rank);
            R = (int)((uint32_t)rem / (uint32_t)p.t_per); x = rem - R * p.t_per;
            const int lb = 4 * p.t_per * R + x + p.t_per * warp;      // this warp's query inside rank r's batch
            g = lb < p.regen_B ? r * p.regen_B + lb : -1;
            seed = __ldg(p.regen_state + 2 * r); off = __ldg(p.regen_state + 2 * r + 1);
        }

        // ---- positive: owned? -> score, local row, histogram, log Q
        if (g >= 0 && p.do_pos) {
            int64_t gp = __ldg(p.pos + g);
            if (gp < 0 || gp >= p.num_items) { bad = true; gp = 0; }
            const int64_t lp = gp - p.row0;
            const bool own = lp >= 0 && lp < p.local_rows;
            float sp = 0.f;
            if (own) {
                for (int c = lane * 4; c < D; c += 128) {
                    const float4 qv = ldg128(p.q_all + (size_t)g * D + c);
                    const float4 vv = ldg128(p.w_local + (size_t)lp * D + c);
                    sp += p.euclid ? sqdist4(qv, vv) : dot4(qv, vv);
                }
                sp = warp_sum_s(sp);
                if (p.euclid) sp = -sp;
            }
            if (lane == 0) {
                p.sp[g] = own ? sp : 0.f;
                p.pos_local[g] = own ? (int32_t)lp : -1;
                if (own && gp != 0) atomicAdd(hist + (lp >> p.bt.shift), 1u);
                if (p.lq_pos_out) p.lq_pos_out[g] = own ? logf(__ldg(p.pop_prob + lp)) : 0.f;
            }
        }

        // ---- negatives, one segment of <= kPrepSeg candidates at a time
        const size_t base = g >= 0 ? (size_t)g * n : 0;
        int kept_total = 0;
        for (int seg0 = 0; seg0 < (p.do_neg ? n : 0); seg0 += kPrepSeg) {
            const int m = min(kPrepSeg, n - seg0);
            const int c = (m + 31) >> 5;                              // candidates per lane
            const int cs = (c & (c - 1)) == 0 ? 31 - __clz(c) : -1;   // c a power of two (every full segment): shifts, no division
            __syncthreads();                                          // the previous segment's staging area is consumed
            // -- A: stage
            if (shared) {
                for (int j = threadIdx.x; j < m; j += blockDim.x) {
                    const uint4 w = Philox::gen(seed, (uint64_t)((int64_t)x * n + seg0 + j), off / 4 + (uint64_t)R);
                    const int slot = cs >= 0 ? (j & (c - 1)) * 33 + (j >> cs) : (j % c) * 33 + j / c;
                    s_cand[0 * kPrepSegWords + slot] = w.x; s_cand[1 * kPrepSegWords + slot] = w.y;
                    s_cand[2 * kPrepSegWords + slot] = w.z; s_cand[3 * kPrepSegWords + slot] = w.w;
                }
            } else if (g >= 0) {
                for (int j = lane; j < m; j += 32) {
                    uint32_t v;
                    if (MODE == 0) {
                        v = (uint32_t)__ldg(p.neg + base + seg0 + j);
                    } else {
                        const int64_t li = li0 + seg0 + j, qq = li / p.regen_T, idx = li - qq * p.regen_T;
                        const uint4 w = Philox::gen(seed, (uint64_t)idx, off / 4 + (uint64_t)(qq >> 2));
                        const int ii = (int)(qq & 3);
                        v = ii == 0 ? w.x : (ii == 1 ? w.y : (ii == 2 ? w.z : w.w));
                    }
                    my[cs >= 0 ? (j & (c - 1)) * 33 + (j >> cs) : (j % c) * 33 + j / c] = v;
                }
            }
            __syncthreads();
            if (g < 0) continue;
            // -- B: lane-private filter + in-place compaction
            int cnt = 0;
#pragma unroll 4
            for (int i = 0; i < c; ++i) {
                const int j = lane * c + i;
                if (j >= m) break;
                const uint32_t v = my[i * 33 + lane];
                bool mine;
                uint32_t keep;
                if (MODE == 2) {
                    float u = curand_uniform_from_u32(v);             // (0, 1]
                    u = u * 1.0f + 0.0f;
                    if (u == 1.0f) u = 0.0f;                          // -> [0, 1)   (ATen uniform_, DistributionTemplates.h:485-506)
                    mine = u > p.cdf_lo && u <= p.cdf_hi;
                    keep = __float_as_uint(u);
                } else if (MODE == 1) {
                    // the id is v mod (N - 1) + 1 (ATen random_from_to, 32-bit path), by Lemire's fastmod (exact for every 32-bit
                    // v:  M = ceil(2^64 / d),  v mod d = mulhi64((M * v) mod 2^64, d)).  Ownership only needs the low product
                    // (monotone, see own_lo / own_hi); the kept words are turned into local rows in C.
                    const uint64_t lb = p.mod_magic * (uint64_t)v;
                    mine = lb >= p.own_lo && lb <= p.own_hi;
                    keep = v;
                } else {
                    uint32_t gid = v;
                    if (gid >= (uint32_t)p.num_items) { bad = true; gid = 0u; }          // also catches negative ids
                    const uint32_t l = gid - (uint32_t)p.row0;        // wraps for rows below the block: fails the test
                    mine = l < (uint32_t)p.local_rows;
                    keep = (uint32_t)j;                               // keeps the position: log Q is looked up in C
                }
                if (mine) { my[cnt * 33 + lane] = keep; ++cnt; }
            }
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(kFull, inc, o);
                if (lane >= o) inc += t;
            }
            const int total = __shfl_sync(kFull, inc, 31);
            const size_t out0 = base + kept_total + (inc - cnt);
            // -- C: resolve + write
            if (MODE == 2) {
                // popularity search on the kept draws, four at a time: the guide lookups, the bisection steps and the final
                // <= 8-entry counts of four draws are independent loads, so their latencies overlap instead of adding up
                constexpr int kW = 4, kLinear = 8;
                for (int t0 = 0; t0 < cnt; t0 += kW) {
                    float u[kW]; int lo[kW], hi[kW]; bool ok[kW];
#pragma unroll
                    for (int k = 0; k < kW; ++k) {
                        ok[k] = t0 + k < cnt;
                        u[k] = ok[k] ? __uint_as_float(my[(t0 + k) * 33 + lane]) : 0.f;
                        lo[k] = 0; hi[k] = ok[k] ? (int)p.local_rows - 1 : 0;
                    }
                    if (p.pop_guide) {
#pragma unroll
                        for (int k = 0; k < kW; ++k) {
                            if (ok[k]) {
                                const int64_t g0 = (int64_t)(u[k] * (float)(1 << p.pop_bits)) - p.pop_k0;   // exact: power-of-two scale
                                lo[k] = __ldg(p.pop_guide + g0); hi[k] = __ldg(p.pop_guide + g0 + 1);
                            }
                        }
                    }
                    bool more = true;
                    while (more) {                                    // bisection in lockstep, down to brackets of <= 8 entries
                        more = false;
                        float tv[kW]; int mid[kW];
#pragma unroll
                        for (int k = 0; k < kW; ++k) {
                            mid[k] = lo[k] + ((hi[k] - lo[k]) >> 1);
                            tv[k] = (hi[k] - lo[k] > kLinear) ? __ldg(p.pop_table + mid[k]) : 0.f;
                        }
#pragma unroll
                        for (int k = 0; k < kW; ++k) {
                            if (hi[k] - lo[k] > kLinear) {
                                if (tv[k] < u[k]) lo[k] = mid[k] + 1; else hi[k] = mid[k];
                                more |= hi[k] - lo[k] > kLinear;
                            }
                        }
                    }
                    int lid[kW];
#pragma unroll
                    for (int k = 0; k < kW; ++k) {
                        lid[k] = lo[k];
#pragma unroll
                        for (int t = 0; t < kLinear; ++t)
                            if (lo[k] + t < hi[k]) lid[k] += (__ldg(p.pop_table + lo[k] + t) < u[k]) ? 1 : 0;
                    }
                    float pr[kW];
#pragma unroll
                    for (int k = 0; k < kW; ++k) pr[k] = ok[k] ? __ldg(p.pop_prob + lid[k]) : 1.f;
#pragma unroll
                    for (int k = 0; k < kW; ++k) {
                        if (ok[k]) {
                            p.neg_c[out0 + t0 + k] = lid[k];
                            if (p.lq_c) p.lq_c[out0 + t0 + k] = logf(pr[k]);
                            if (p.row0 + lid[k] != 0) atomicAdd(hist + ((uint32_t)lid[k] >> p.bt.shift), 1u);
                        }
                    }
                }
            } else {
                for (int t = 0; t < cnt; ++t) {
                    const uint32_t v = my[t * 33 + lane];
                    int lid;
                    float lq = 0.f;
                    if (MODE == 0) {
                        int64_t gid = (int64_t)__ldg(p.neg + base + seg0 + v);
                        if (gid < 0 || gid >= p.num_items) gid = 0;
                        lid = (int)(gid - p.row0);
                        if (p.logq_neg) lq = __ldg(p.logq_neg + base + seg0 + v);
                    } else {
                        lid = (int)((uint32_t)__umul64hi(p.mod_magic * (uint64_t)v, (uint64_t)(uint32_t)(p.num_items - 1)) + 1u
                                    - (uint32_t)p.row0);
                    }
                    p.neg_c[out0 + t] = lid;
                    if (p.lq_c) p.lq_c[out0 + t] = lq;
                    if (p.row0 + lid != 0) atomicAdd(hist + ((uint32_t)lid >> p.bt.shift), 1u);   // padding row: scored, no gradient
                }
            }
            kept_total += total;
        }
        if (g >= 0 && lane == 0 && p.do_neg) p.ncount[g] = kept_total;
        if (dynamic) {
            if (threadIdx.x == 0) s_next = nxt;
            __syncthreads();                                          // (the next write of s_next is behind the segment loop's barriers)
            item = s_next;
        } else {
            item += gridDim.x;
        }
    }
    if (bad) atomicOr(p.err, 1u);
    if (p.use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < p.bt.nbins; i += blockDim.x) {
            const uint32_t c = s_hist[i];
            if (c) atomicAdd(p.bt.cnt + i, c);
        }
    }
}

struct FinishParams {
    const float* w_local; const float* q_all;
    const float* stats_all;        // [world, G, 2]
    const float* sp; const float* logq_pos;
    const int32_t* pos_local; const uint32_t* slot_pos; const uint32_t* off;
    uint64_t* ent; float* dq; float* loss_part; float* lse;
    int G, D, world, rank, ssm, euclid;
    float coef_scale, loss_scale;
    uint32_t* bin_cursor; int bin_shift;       // binned grouping: the positive's entry is appended to its bin's list
    int64_t row0;
};

// one warp per query
__global__ void __launch_bounds__(kPrepWarps * 32)
shard_finish_kernel(const FinishParams p) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * kPrepWarps + (threadIdx.x >> 5);
    if (b >= p.G) return;
    const int D = p.D;
    const size_t G = (size_t)p.G;

    // ---- merge the owners' statistics (every owner runs exactly this sequence) -------------------
    float cpos, mul_own, cs_own, loss_b, lse_b = 0.f;
    const float2 mine = *reinterpret_cast<const float2*>(p.stats_all + 2 * ((size_t)p.rank * G + b));
    if (!p.ssm) {
        float CS = 0.f, LS = 0.f;
        for (int o = 0; o < p.world; ++o) {
            const float2 s = *reinterpret_cast<const float2*>(p.stats_all + 2 * ((size_t)o * G + b));
            CS += s.x; LS += s.y;
        }
        cpos = -CS; mul_own = 1.f; cs_own = mine.x;
        loss_b = LS * p.loss_scale;
    } else {
        float M = -INFINITY;
        for (int o = 0; o < p.world; ++o) M = fmaxf(M, p.stats_all[2 * ((size_t)o * G + b)]);
        float L = 0.f;
        for (int o = 0; o < p.world; ++o) {
            const float2 s = *reinterpret_cast<const float2*>(p.stats_all + 2 * ((size_t)o * G + b));
            if (s.x != -INFINITY) L += s.y * expf(s.x - M);
        }
        const float z0 = p.sp[b] - (p.logq_pos ? p.logq_pos[b] : 0.f);
        const float M2 = fmaxf(M, z0);
        const float L2 = ((M == -INFINITY) ? 0.f : L * expf(M - M2)) + expf(z0 - M2);
        lse_b = M2 + logf(L2);
        cpos = (expf(z0 - lse_b) - 1.f) * p.coef_scale;
        mul_own = (mine.x == -INFINITY) ? 0.f : expf(mine.x - lse_b) * p.coef_scale;    // accumulator is relative to m_own
        cs_own = mine.y * mul_own;
        loss_b = (lse_b - z0) * p.loss_scale;
    }

    const int lp = p.pos_local[b];
    for (int c = lane * 4; c < D; c += 128) {
        float4 a = *reinterpret_cast<const float4*>(p.dq + (size_t)b * D + c);
        a.x *= mul_own; a.y *= mul_own; a.z *= mul_own; a.w *= mul_own;
        if (p.euclid) {
            const float4 q = ldg128(p.q_all + (size_t)b * D + c);
            a.x = 2.f * (a.x - cs_own * q.x); a.y = 2.f * (a.y - cs_own * q.y);
            a.z = 2.f * (a.z - cs_own * q.z); a.w = 2.f * (a.w - cs_own * q.w);
            if (lp >= 0) {
                const float4 v = ldg128(p.w_local + (size_t)lp * D + c);
                const float t = 2.f * cpos;
                a.x += t * (v.x - q.x); a.y += t * (v.y - q.y); a.z += t * (v.z - q.z); a.w += t * (v.w - q.w);
            }
        } else if (lp >= 0) {
            fma4(a, cpos, ldg128(p.w_local + (size_t)lp * D + c));
        }
        *reinterpret_cast<float4*>(p.dq + (size_t)b * D + c) = a;
    }
    if (lane == 0) {
        p.loss_part[b] = loss_b;
        p.lse[b] = lse_b;
        if (lp >= 0 && p.bin_cursor) {
            if (p.row0 + lp != 0) {                                   // global row 0 is the padding row: no gradient
                const uint32_t at = atomicAdd(p.bin_cursor + (size_t)((uint32_t)lp >> p.bin_shift) * kCursorStride, 1u);
                const uint32_t low = (uint32_t)b | kDirect | (((uint32_t)lp & ((1u << p.bin_shift) - 1u)) << (31 - p.bin_shift));
                p.ent[at] = (uint64_t)low | ((uint64_t)__float_as_uint(cpos) << 32);
            }
        } else if (lp >= 0) {
            const uint32_t sl = p.slot_pos[b];
            if (sl != kNoSlot)
                p.ent[__ldg(p.off + lp) + sl] = (uint64_t)((uint32_t)b | kDirect) | ((uint64_t)__float_as_uint(cpos) << 32);
        }
    }
}

}  // namespace rsb

using namespace rsb;

extern "C" size_t rsb200_sizeof_shard_args(void) { return sizeof(rsb200_shard_args); }
extern "C" int32_t rsb200_bin_shift(int64_t num_rows, int64_t touches, int64_t num_queries) {
    const int s = bin_shift_for(num_rows, touches, num_queries);
    return s >= kMinBinShift ? s : 0;
}
extern "C" int64_t rsb200_bin_heavy_elems(void) { return bin_scatter_grid() * ((int64_t)1 << kMaxBinShift); }
extern "C" int64_t rsb200_scan_tmp_elems(int64_t num_rows) { return scan_tmp_elems(num_rows); }

// Host arithmetic of the owner-side regeneration of UniformSampler draws (no device work): a 32-bit Philox word v becomes the
// id  v mod d + 1,  d = num_items - 1  (ATen random_from_to).  With M = ceil(2^64 / d), Lemire's fastmod gives
// v mod d = floor(low64(M v) d / 2^64), which is monotone in low64(M v): "row0 <= id < row0 + local_rows" is the range test
// lo <= low64(M v) <= hi, and only the draws that pass it pay for the 64 x 64 high product.  Empty block: lo = 1, hi = 0.
extern "C" int32_t rsb200_uniform_owner_range(int64_t num_items, int64_t row0, int64_t local_rows, uint64_t* magic, uint64_t* lo,
                                              uint64_t* hi) {
    RSB_REQUIRE(magic && lo && hi, RSB200_EINVAL, "null output");
    RSB_REQUIRE(num_items >= 1 && row0 >= 0 && local_rows >= 0 && num_items - 1 <= (int64_t)0xffffffffLL, RSB200_EINVAL,
                "bad sizes (ids are drawn from 32-bit words: num_items - 1 <= 2^32 - 1)");
    const int64_t d = num_items - 1, a = row0 - 1, b = row0 + local_rows - 1;          // a <= v mod d < b
    *magic = d > 0 ? (~(uint64_t)0) / (uint64_t)d + 1 : 0;
    *lo = 1; *hi = 0;
    if (d > 0 && b > 0 && b > a) {
        const unsigned __int128 one = (unsigned __int128)1 << 64;
        *lo = a <= 0 ? 0 : (uint64_t)(((unsigned __int128)a * one + (unsigned __int128)(d - 1)) / (unsigned __int128)d);
        *hi = b >= d ? ~(uint64_t)0 : (uint64_t)(((unsigned __int128)b * one + (unsigned __int128)(d - 1)) / (unsigned __int128)d - 1);
    }
    return RSB200_OK;
}

extern "C" int32_t rsb200_shard_step(const rsb200_shard_args* a, int32_t phases, void* stream) {
    RSB_REQUIRE(a != nullptr, RSB200_EINVAL, "null args");
    RSB_REQUIRE(a->d >= 4 && a->d % 4 == 0 && a->d <= 512, a->d > 512 ? RSB200_EUNSUPPORTED : RSB200_EINVAL,
                "embedding dim must be a multiple of 4 in [4, 512], got %lld", (long long)a->d);
    RSB_REQUIRE(a->G >= 0 && a->n >= 0 && a->num_items >= 1 && a->local_rows >= 1 && a->row0 >= 0 &&
                a->row0 + a->local_rows <= a->num_items, RSB200_EINVAL, "bad sizes / row block");
    RSB_REQUIRE(a->world >= 1 && a->rank >= 0 && a->rank < a->world, RSB200_EINVAL, "bad world / rank");
    RSB_REQUIRE(a->G * (a->n + 1) < ((int64_t)1 << 31), RSB200_EUNSUPPORTED, "G*(n+1) must be < 2^31");
    RSB_REQUIRE(a->num_items < ((int64_t)1 << 31), RSB200_EUNSUPPORTED, "table rows must be < 2^31");
    RSB_REQUIRE(a->loss_kind == RSB200_LOSS_BPR || a->loss_kind == RSB200_LOSS_SSM, RSB200_EINVAL, "bad loss_kind");
    RSB_REQUIRE(a->score_kind == RSB200_SCORE_IP || a->score_kind == RSB200_SCORE_EUCLID, RSB200_EINVAL, "bad score_kind");
    RSB_REQUIRE(a->sink == RSB200_SINK_COMPACT || a->sink == RSB200_SINK_DENSE, RSB200_EINVAL, "bad sink");
    RSB_REQUIRE(a->w_local && a->q_all && a->pos && (a->n == 0 || a->neg || a->regen_state) && a->sp && a->stats_all && a->dq && a->loss,
                RSB200_EINVAL, "null table / batch / exchange pointer");
    if (a->regen_state) {
        RSB_REQUIRE(a->regen_B >= 1 && a->regen_B * a->world == a->G, RSB200_EINVAL, "regen_B * world must equal G");
        RSB_REQUIRE(a->regen_sm_count > 0 && a->regen_max_threads_per_sm >= 256, RSB200_EINVAL, "bad draw policy");
        RSB_REQUIRE(a->logq_neg == nullptr, RSB200_EINVAL, "owner-side regeneration computes log Q itself: logq_neg must be NULL");
        RSB_REQUIRE(a->regen_kind == 0 || a->regen_kind == 1, RSB200_EINVAL, "regen_kind must be 0 (uniform) or 1 (popularity)");
        RSB_REQUIRE(a->num_items - 1 < ((int64_t)1 << 28) && a->regen_B * a->n * 8 < ((int64_t)1 << 31), RSB200_EUNSUPPORTED,
                    "draw outside ATen's 32-bit path (see rsb200_sample_uniform)");
    }
    RSB_REQUIRE(a->neg_c && a->ncount && a->pos_local && a->ent && a->loss_part && a->lse && a->totals && a->err_flag, RSB200_EINVAL,
                "null workspace pointer");
    RSB_REQUIRE((a->logq_neg == nullptr && !(a->regen_state && a->regen_kind == 1)) || a->lq_c != nullptr, RSB200_EINVAL,
                "logq_neg / the regenerated popularity draw need the lq_c workspace");
    RSB_REQUIRE(aligned16(a->w_local) && aligned16(a->q_all) && aligned16(a->dq) && aligned16(a->item_vals) &&
                (reinterpret_cast<uintptr_t>(a->stats_all) & 7u) == 0, RSB200_EINVAL, "row buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t G = a->G, n = a->n;
    const bool ssm = a->loss_kind == RSB200_LOSS_SSM, eu = a->score_kind == RSB200_SCORE_EUCLID;
    const double denom = ssm ? (double)(G > 0 ? G : 1) : (double)(G > 0 ? G : 1) * (double)(n > 0 ? n : 1);
    const float coef_scale = (float)((double)a->grad_scale / denom), loss_scale = (float)(1.0 / denom);
    const unsigned qblocks = (unsigned)cdiv(G > 0 ? G : 1, kPrepWarps);
    int32_t rc;

    const bool bins = a->grouping == 1;
    BinTable bt = {}, none = {};
    if (bins) {
        RSB_REQUIRE(a->bin_shift >= kMinBinShift && a->bin_shift <= kMaxBinShift, RSB200_EINVAL, "bin_shift must be in [%d, %d]",
                    kMinBinShift, kMaxBinShift);
        RSB_REQUIRE(G <= ((int64_t)1 << (31 - a->bin_shift)), RSB200_EUNSUPPORTED, "G = %lld does not fit the query bits of a binned entry",
                    (long long)G);
        RSB_REQUIRE(G * a->d < ((int64_t)1 << 30), RSB200_EUNSUPPORTED, "G * d must be < 2^30");
        RSB_REQUIRE(a->bin_cnt && a->bin_off && a->bin_cursor && a->bin_status && a->bin_ticket && a->bin_heavy, RSB200_EINVAL,
                    "null bin workspace pointer (grouping 1)");
        bt.cnt = a->bin_cnt; bt.off = a->bin_off; bt.cursor = a->bin_cursor; bt.status = a->bin_status; bt.ticket = a->bin_ticket;
        bt.totals = a->totals; bt.shift = a->bin_shift; bt.num_rows = a->local_rows;
        bt.nbins = (int)cdiv(a->local_rows, (int64_t)1 << a->bin_shift);
    } else {
        RSB_REQUIRE(a->slot_neg && a->slot_pos && a->off && a->urow && a->scan_tmp, RSB200_EINVAL, "null workspace pointer (grouping 0)");
        RSB_REQUIRE(a->regen_kind == 0, RSB200_EUNSUPPORTED, "owner-side regeneration of the popularity draw needs grouping 1");
    }
    if (a->regen_state && a->regen_kind == 1) {
        RSB_REQUIRE(a->pop_table_local && a->pop_prob_local, RSB200_EINVAL, "regen_kind 1 needs this owner's slices of table / pop_prob");
        RSB_REQUIRE(a->pop_guide_local == nullptr || (a->pop_guide_bits >= 1 && a->pop_guide_bits <= 24), RSB200_EINVAL, "bad guide_bits");
    }

    // PREP = PREP_NEG (negatives: needs only the draw; may run while the queries are still being gathered) + PREP_POS
    // (positives + the scan of the bin histogram)
    const bool prep_neg = (phases & (RSB200_SHARD_PREP | RSB200_SHARD_PREP_NEG)) != 0;
    const bool prep_pos = (phases & (RSB200_SHARD_PREP | RSB200_SHARD_PREP_POS)) != 0;
    RSB_REQUIRE(bins || !(phases & (RSB200_SHARD_PREP_NEG | RSB200_SHARD_PREP_POS)), RSB200_EUNSUPPORTED, "split PREP needs grouping 1");
    if ((prep_neg || prep_pos) && bins) {
        // counts of this step's touches per bin (PREP_NEG starts them) + the word behind them: PREP's work counter
        if (prep_neg) RSB_CUDA(cudaMemsetAsync(bt.cnt, 0, sizeof(uint32_t) * ((size_t)bt.nbins + 1), st));
        if (G > 0) {
            PrepBinParams p;
            p.w_local = a->w_local; p.q_all = a->q_all; p.pos = a->pos; p.neg = a->neg; p.logq_neg = a->logq_neg;
            p.neg_c = a->neg_c; p.lq_c = (a->logq_neg || (a->regen_state && a->regen_kind == 1)) ? a->lq_c : nullptr;
            p.ncount = a->ncount; p.pos_local = a->pos_local; p.sp = a->sp; p.lq_pos_out = a->lq_pos_out; p.err = a->err_flag;
            p.work = bt.cnt + bt.nbins;
            p.bt = bt; p.num_items = a->num_items; p.row0 = a->row0; p.local_rows = a->local_rows;
            p.G = (int)G; p.n = (int)n; p.D = (int)a->d; p.euclid = eu;
            p.regen_state = a->regen_state; p.regen_B = (int)a->regen_B; p.regen_T = 0; p.t_per = 0; p.n_round_blocks = 0;
            p.pop_table = a->pop_table_local; p.pop_prob = a->pop_prob_local; p.pop_guide = a->pop_guide_local;
            p.pop_bits = a->pop_guide_bits; p.pop_k0 = a->pop_guide_k0; p.cdf_lo = a->pop_cdf_lo; p.cdf_hi = a->pop_cdf_hi;
            RSB_REQUIRE(p.lq_pos_out == nullptr || p.pop_prob != nullptr, RSB200_EINVAL, "lq_pos_out needs pop_prob_local");
            int mode = 0;
            if (a->regen_state) {        // ATen policy: grid = min(sm * (max_threads / 256), ceil(numel / 256)), T = 256 grid
                const int64_t numel = a->regen_B * n;
                int64_t grid = (int64_t)a->regen_sm_count * (a->regen_max_threads_per_sm / 256);
                if (cdiv(numel, 256) < grid) grid = cdiv(numel, 256);
                p.regen_T = 256 * (grid > 0 ? grid : 1);
                if (n > 0 && p.regen_T % n == 0) {                    // four queries share every Philox block
                    p.t_per = (int)(p.regen_T / n);
                    p.n_round_blocks = (int)cdiv(a->regen_B, 4 * (int64_t)p.t_per);
                }
                mode = a->regen_kind == 1 ? 2 : 1;
            }
            p.do_neg = prep_neg ? 1 : 0; p.do_pos = prep_pos ? 1 : 0;       // one launch does whatever halves are requested
            p.mod_magic = 0; p.own_lo = 1; p.own_hi = 0;
            if (mode == 1) {                                          // uniform regeneration: ownership thresholds of this block
                const int32_t rc = rsb200_uniform_owner_range(a->num_items, a->row0, a->local_rows, &p.mod_magic, &p.own_lo, &p.own_hi);
                if (rc != RSB200_OK) return rc;
            }
            if (!p.do_neg) { p.t_per = 0; p.n_round_blocks = 0; mode = 0; }      // positives only: plain query-major mapping
            p.use_smem = bt.nbins <= 8192;
            const size_t smem = sizeof(uint32_t) * ((p.use_smem ? (size_t)bt.nbins : 0) + 4 * (size_t)kPrepSegWords);
            const int64_t nitems = p.t_per ? (G / a->regen_B) * p.n_round_blocks * p.t_per : cdiv(G, 4);
            int64_t blocks = (int64_t)sm_count() * 6;
            if (blocks > nitems) blocks = nitems;
#define RSB_PREP(M)                                                                                                       \
    do {                                                                                                                  \
        RSB_CUDA(cudaFuncSetAttribute(shard_prep_bins_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        shard_prep_bins_kernel<M><<<(unsigned)blocks, 128, smem, st>>>(p);                                                \
    } while (0)
            if (mode == 0) RSB_PREP(0); else if (mode == 1) RSB_PREP(1); else RSB_PREP(2);
#undef RSB_PREP
            RSB_LAUNCH_CHECK();
        }
        if (prep_pos) {
            rc = launch_bin_scan(bt, none, st);
            if (rc) return rc;
        }
    }
    if ((phases & RSB200_SHARD_PREP) && !bins) {
        RSB_CUDA(cudaMemsetAsync(a->off, 0, sizeof(uint32_t) * (size_t)(a->local_rows + 1), st));
        if (G > 0) {
            PrepParams p;
            p.w_local = a->w_local; p.q_all = a->q_all; p.pos = a->pos; p.neg = a->neg; p.logq_neg = a->logq_neg;
            p.cnt = a->off; p.neg_c = a->neg_c; p.slot_neg = a->slot_neg; p.lq_c = a->logq_neg ? a->lq_c : nullptr;
            p.ncount = a->ncount; p.pos_local = a->pos_local; p.slot_pos = a->slot_pos; p.sp = a->sp; p.err = a->err_flag;
            p.num_items = a->num_items; p.row0 = a->row0; p.local_rows = a->local_rows;
            p.G = (int)G; p.n = (int)n; p.D = (int)a->d; p.euclid = eu;
            p.regen_state = a->regen_state; p.regen_B = (int)a->regen_B; p.regen_T = 0;
            if (a->regen_state) {        // ATen policy: grid = min(sm * (max_threads / 256), ceil(numel / 256)), T = 256 grid
                const int64_t numel = a->regen_B * n;
                int64_t grid = (int64_t)a->regen_sm_count * (a->regen_max_threads_per_sm / 256);
                if (cdiv(numel, 256) < grid) grid = cdiv(numel, 256);
                p.regen_T = 256 * (grid > 0 ? grid : 1);
            }
            shard_prep_kernel<<<qblocks, kPrepWarps * 32, 0, st>>>(p);
            RSB_LAUNCH_CHECK();
        }
        rc = launch_scan(a->off, a->local_rows, a->urow, a->cap, a->totals, a->scan_tmp, a->scan_tmp_elems, st);
        if (rc) return rc;
        if (G > 0 && n > 0) {
            shard_resolve_kernel<<<qblocks, kPrepWarps * 32, 0, st>>>(a->neg_c, a->slot_neg, a->ncount, a->off, (int)G, (int)n);
            RSB_LAUNCH_CHECK();
        }
    }
    if ((phases & RSB200_SHARD_FWD) && G > 0) {
        FwdParams p;
        p.w_item = a->w_local; p.w_user = a->q_all; p.user = nullptr; p.pos = nullptr; p.neg = a->neg_c;
        p.logq_pos = nullptr; p.logq_neg = a->logq_neg ? a->lq_c : nullptr;
        p.off_item = a->off; p.off_user = nullptr; p.slot_neg = a->slot_neg; p.slot_pos = nullptr; p.slot_user = nullptr;
        p.ent_item = a->ent; p.ent_user = nullptr; p.q_buf = nullptr; p.dq_buf = a->dq; p.loss_part = nullptr; p.lse = nullptr;
        p.pos_score = nullptr; p.neg_score = nullptr;
        p.num_items = (int)a->local_rows; p.num_users = (int)G; p.B = (int)G; p.n = (int)n; p.D = (int)a->d;
        p.coef_scale = coef_scale; p.loss_scale = loss_scale; p.prefetch = 0; p.hint = 0; p.slot_abs = 1; p.cstage = nullptr;
        p.ncount = a->ncount; p.sp_in = a->sp; p.stats_part = a->stats_all + 2 * (size_t)a->rank * (size_t)G;
        p.bin_cursor = bins ? bt.cursor : nullptr; p.bin_shift = a->bin_shift; p.bin_bbits = 31 - a->bin_shift;
        p.bin_cursor_user = nullptr; p.bin_shift_user = 0; p.pad_row = a->row0 == 0 ? 0 : -1;
        if (a->regen_state && a->regen_kind == 1) p.logq_neg = a->lq_c;            // log Q of the regenerated popularity draw
        rc = launch_pair_fwd_partial(p, a->loss_kind, a->score_kind, st);
        if (rc) return rc;
    }
    if (phases & RSB200_SHARD_FINISH) {
        if (G > 0) {
            FinishParams f;
            f.w_local = a->w_local; f.q_all = a->q_all; f.stats_all = a->stats_all; f.sp = a->sp; f.logq_pos = a->logq_pos;
            f.pos_local = a->pos_local; f.slot_pos = a->slot_pos; f.off = a->off; f.ent = a->ent; f.dq = a->dq;
            f.loss_part = a->loss_part; f.lse = a->lse; f.G = (int)G; f.D = (int)a->d; f.world = a->world; f.rank = a->rank;
            f.ssm = ssm; f.euclid = eu; f.coef_scale = coef_scale; f.loss_scale = loss_scale;
            f.bin_cursor = bins ? bt.cursor : nullptr; f.bin_shift = a->bin_shift; f.row0 = a->row0;
            shard_finish_kernel<<<qblocks, kPrepWarps * 32, 0, st>>>(f);
            RSB_LAUNCH_CHECK();
        }
        rc = launch_loss_sum(a->loss_part, (int)G, a->loss, st);
        if (rc) return rc;
    }
    if (phases & RSB200_SHARD_SCATTER) {
        RSB_REQUIRE(a->item_rows && a->item_vals, RSB200_EINVAL, "SCATTER needs item_rows / item_vals");
        if (bins) {
            BinScatterParams b;
            b.ent = a->ent; b.bin_off = bt.off; b.status = bt.status; b.ticket = bt.ticket; b.totals = bt.totals;
            b.heavy_counts = a->bin_heavy; b.src = a->q_all; b.lse = a->lse; b.w = a->w_local; b.gscale = a->grad_scale_dev;
            b.rows_out = a->item_rows; b.vals = a->item_vals; b.cap = a->cap;
            b.nbins = bt.nbins; b.shift = bt.shift; b.bbits = 31 - bt.shift; b.D = (int)a->d;
            b.ssm_scale = coef_scale; b.dense = a->sink == RSB200_SINK_DENSE; b.accumulate = a->accumulate; b.euclid = eu;
            b.opt = -1; b.w_rw = nullptr; b.s1 = nullptr; b.s2 = nullptr; b.lr = b.b1 = b.b2 = b.eps = b.step_size = 0.f; b.tune = 0;
            return launch_bin_scatter(b, st);
        }
        ScatterParams s;
        s.off = a->off; s.urow = a->urow; s.totals = a->totals; s.ent = a->ent; s.src = a->q_all; s.lse = a->lse;
        s.w = a->w_local; s.gscale = a->grad_scale_dev; s.rows_out = a->item_rows; s.vals = a->item_vals; s.cap = a->cap;
        s.D = (int)a->d; s.ssm_scale = coef_scale;
        s.dense = a->sink == RSB200_SINK_DENSE; s.accumulate = a->accumulate; s.euclid = eu; s.hint = 0; s.opt = -1;
        rc = launch_scatter(s, a->cap, st);
        if (rc) return rc;
    }
    return 0;
}
