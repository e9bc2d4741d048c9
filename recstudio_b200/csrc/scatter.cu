// scatter.cu -- segmented (row-grouped) gradient accumulate: the sparse-row replacement of
// embedding_dense_backward (E2; autograd of nn.Embedding, recommender.py:638), plus the
// final loss reduction.
//
// Input: the entry list written by pair_fwd.cu, grouped by table row through the CSR
// offsets of group.cu.  Entry = (query index b | kDirect flag, value).  The gradient row is
//     g[row] = sum_e  c_e * src[b_e, :]            (src = q_buf for items, dq_buf for users)
//     c_e    = value                               if the entry is flagged kDirect
//            = exp(value - lse[b_e]) * ssm_scale   otherwise (SampledSoftmax negatives: the
//              logit is stored and the softmax denominator is applied here, so the forward
//              kernel never needs a second sweep over its negatives)
//     Euclid: g[row] = 2 * (g[row] - (sum_e c_e) * W[row, :])
// Every unique row is produced by exactly ONE warp: no floating-point atomics, no
// zero-fill of an [N,d] buffer, and each output row is written once with 16-byte stores.
// A warp owns 32 consecutive unique rows; because the entry list is sorted by row, their
// entries are one contiguous range that is streamed with coalesced loads.  Row boundaries
// inside a 32-entry batch are a bit mask built with one warp reduction (REDUX.OR), so the
// per-entry work is 2 shuffles + one 16-byte load + 4 FMA + a bit test.
#include "common.cuh"
#include "kernels.h"
#include "rowopt.cuh"

namespace rsb {

constexpr int kScatWarps = 8;

// optimizer update of 4 consecutive elements of one row: the very same code as rows_update_kernel (rowopt.cuh)
template <int OPT>
__device__ __forceinline__ void apply_update(const ScatterParams& p, size_t off, float4 g) {
    const OptParams o = {p.lr, p.b1, p.b2, p.eps, p.step_size};
    opt_update4<OPT>(p.w_rw, p.s1, p.s2, off, g, o);
}

// PLAIN = compact sink, overwrite, inner-product: the hot configuration with all options compiled out
// HINT (PLAIN only) = L2 eviction priorities: the entry list and the gradient rows are touched once
// (evict_first), the query matrix is re-read by every entry (evict_last).
// OCC = CTAs per SM the kernel is compiled for (register cap), UNR = entries (query-row loads) in flight per warp.
// FULL = every lane owns a live float4 of the row (D == 128 VPL): the column predicates compile out.
// OPT >= 0 = RSB200_SINK_APPLY: the row's optimizer update (rowopt.cu arithmetic) replaces the gradient store.
template <int VPL, bool PLAIN, bool HINT = false, int OCC = 4, int UNR = 4, bool FULL = false, int OPT = -1>
__global__ void __launch_bounds__(kScatWarps * 32, OCC)
scatter_kernel(const ScatterParams p) {
    uint64_t pol_first = 0, pol_last = 0;
    if (HINT) { pol_first = l2_policy(1); pol_last = l2_policy(2); }
    const int lane = threadIdx.x & 31;
    const int D = p.D;
    const uint32_t R = (uint32_t)min((int64_t)p.totals[1], p.cap);
    const uint32_t nchunks = (R + 31) / 32;
    const float gs = p.gscale ? __ldg(p.gscale) : 1.0f;
    const uint32_t warps_total = gridDim.x * kScatWarps;
    const bool dense = !PLAIN && p.dense, accumulate = !PLAIN && p.accumulate, euclid = !PLAIN && p.euclid;
    bool act[VPL];
#pragma unroll
    for (int t = 0; t < VPL; ++t) act[t] = FULL || (lane * 4 + t * 128) < D;
    const float* src_lane = p.src + lane * 4;          // 32-bit row offsets below: B * D < 2^31

    for (uint32_t chunk = blockIdx.x * kScatWarps + (threadIdx.x >> 5); chunk < nchunks; chunk += warps_total) {
        const uint32_t u = chunk * 32 + lane;
        const bool have = u < R;
        uint32_t r = 0, beg = 0, end = 0;
        if (have) {
            r = __ldg(p.urow + u);
            beg = __ldg(p.off + r);
            end = __ldg(p.off + r + 1);
            p.rows_out[u] = (int64_t)r;
        }
        const int nrows = min(32u, R - chunk * 32);
        const uint32_t e_begin = __shfl_sync(kFull, beg, 0);
        const uint32_t e_end = __shfl_sync(kFull, end, nrows - 1);
        const uint32_t last = end - 1;                 // index of this lane's row's last entry (rows have >= 1 entry)

        int cur = 0;                                   // row of the chunk being accumulated
        float* dst_run = p.vals + (size_t)chunk * 32 * D + lane * 4;   // PLAIN && FULL: output pointer of the current row
        float4 acc[VPL];
#pragma unroll
        for (int t = 0; t < VPL; ++t) acc[t] = make_float4(0, 0, 0, 0);
        float csum = 0.f;
        // Euclid needs W[row] at flush time (2.9 GB of extra DRAM reads at config 2): every lane asks L2 to
        // fetch the table row of ITS unique row now, so the flush-time load below is an L2 hit.
        // (A register-resident rolling prefetch was measured slower: 1.71 ms vs 1.29 ms.)
        if (euclid && have) {
            const char* wr = reinterpret_cast<const char*>(p.w + (size_t)r * D);
            for (int o = 0; o < D * 4; o += 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(wr + o));
        }

        for (uint32_t eb = e_begin; eb < e_end; eb += 32) {
            const uint32_t e = eb + lane;
            uint32_t bq = 0; float c = 0.f;
            if (e < e_end) {
                const uint64_t en = HINT ? ldg64_stream_hint(p.ent + e, pol_first)
                                         : __ldg(reinterpret_cast<const unsigned long long*>(p.ent) + e);
                const uint32_t lo = (uint32_t)en;
                const float val = __uint_as_float((uint32_t)(en >> 32));
                bq = lo & 0x7FFFFFFFu;
                c = (lo & kDirect) ? val : expf(val - __ldg(p.lse + bq)) * p.ssm_scale;
                if (PLAIN) c *= gs;            // upstream gradient folded into the coefficient: no per-row scaling at flush time
            }
            // bit t set <=> entry eb+t is the last entry of its row
            const uint32_t contrib = (have && last >= eb && last < eb + 32) ? (1u << (last - eb)) : 0u;
            const uint32_t lastmask = __reduce_or_sync(kFull, contrib);
            const int cnt = min(32u, e_end - eb);
            for (int t0 = 0; t0 < cnt; t0 += UNR) {    // lanes >= cnt hold (b = 0, c = 0): harmless
                float4 v[UNR][VPL];
                float cc[UNR];
#pragma unroll
                for (int k = 0; k < UNR; ++k) {
                    const uint32_t bt = __shfl_sync(kFull, bq, t0 + k);
                    cc[k] = __shfl_sync(kFull, c, t0 + k);
                    const float* srow = src_lane + bt * (uint32_t)D;
#pragma unroll
                    for (int x = 0; x < VPL; ++x)
                        v[k][x] = act[x] ? (HINT ? ldg128_hint(srow + x * 128, pol_last) : ldg128(srow + x * 128))
                                         : make_float4(0, 0, 0, 0);
                }
#pragma unroll
                for (int k = 0; k < UNR; ++k) {
#pragma unroll
                    for (int x = 0; x < VPL; ++x) fma4(acc[x], cc[k], v[k][x]);
                    csum += cc[k];
                    if (PLAIN && FULL) {
                        if ((lastmask >> (t0 + k)) & 1u) {             // warp-uniform: the current row is complete
                            if (OPT >= 0) {                            // RSB200_SINK_APPLY: update W[row] in place
                                const uint32_t rr = __shfl_sync(kFull, r, cur);
#pragma unroll
                                for (int x = 0; x < VPL; ++x) {
                                    apply_update<OPT>(p, (size_t)rr * D + lane * 4 + x * 128, acc[x]);
                                    acc[x] = make_float4(0, 0, 0, 0);
                                }
                                ++cur;
                            } else {
#pragma unroll
                                for (int x = 0; x < VPL; ++x) {
                                    if (HINT) stg128_stream_hint(dst_run + x * 128, acc[x], pol_first);
                                    else stg128_stream(dst_run + x * 128, acc[x]);
                                    acc[x] = make_float4(0, 0, 0, 0);
                                }
                                dst_run += D;                          // compact sink: rows of the chunk are consecutive
                            }
                        }
                    } else if ((lastmask >> (t0 + k)) & 1u) {          // warp-uniform: row `cur` is complete
                        uint32_t rr = 0;
                        if (!PLAIN) rr = __shfl_sync(kFull, r, cur);
                        const size_t orow = dense ? (size_t)rr : (size_t)chunk * 32 + cur;
#pragma unroll
                        for (int x = 0; x < VPL; ++x) {
                            if (act[x]) {
                                const int col = lane * 4 + x * 128;
                                float4 a = acc[x];
                                if (euclid) {
                                    const float4 wv = ldg128(p.w + (size_t)rr * D + col);
                                    a.x = 2.f * (a.x - csum * wv.x); a.y = 2.f * (a.y - csum * wv.y);
                                    a.z = 2.f * (a.z - csum * wv.z); a.w = 2.f * (a.w - csum * wv.w);
                                }
                                if (!PLAIN) { a.x *= gs; a.y *= gs; a.z *= gs; a.w *= gs; }   // PLAIN: folded into c
                                if (OPT >= 0) {
                                    apply_update<OPT>(p, (size_t)rr * D + col, a);
                                } else {
                                    float* dst = p.vals + orow * D + col;
                                    if (accumulate) {
                                        const float4 o = *reinterpret_cast<const float4*>(dst);
                                        a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
                                    }
                                    if (HINT) stg128_stream_hint(dst, a, pol_first);
                                    else stg128_stream(dst, a);
                                }
                            }
                            acc[x] = make_float4(0, 0, 0, 0);
                        }
                        csum = 0.f;
                        ++cur;
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256)
loss_sum_kernel(const float* __restrict__ part, int B, float* __restrict__ loss) {
    __shared__ double sh[256];
    double a = 0.0;
    for (int i = threadIdx.x; i < B; i += 256) a += (double)part[i];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = (float)sh[0];
}

template <int VPL>
static void launch_scatter_v(const ScatterParams& p, unsigned blocks, cudaStream_t st) {
    if (p.opt >= 0 && !p.euclid && p.D == 128 * VPL) {      // RSB200_SINK_APPLY on full rows: the specialised body
        if (p.opt == 0) scatter_kernel<VPL, true, false, 4, 4, true, 0><<<blocks, kScatWarps * 32, 0, st>>>(p);
        else if (p.opt == 1) scatter_kernel<VPL, true, false, 4, 4, true, 1><<<blocks, kScatWarps * 32, 0, st>>>(p);
        else scatter_kernel<VPL, true, false, 4, 4, true, 2><<<blocks, kScatWarps * 32, 0, st>>>(p);
        return;
    }
    if (p.opt >= 0) {               // RSB200_SINK_APPLY: generic (non-PLAIN) body with the update in the epilogue
        if (p.opt == 0) scatter_kernel<VPL, false, false, 4, 4, false, 0><<<blocks, kScatWarps * 32, 0, st>>>(p);
        else if (p.opt == 1) scatter_kernel<VPL, false, false, 4, 4, false, 1><<<blocks, kScatWarps * 32, 0, st>>>(p);
        else scatter_kernel<VPL, false, false, 4, 4, false, 2><<<blocks, kScatWarps * 32, 0, st>>>(p);
        return;
    }
    const bool plain = !p.dense && !p.accumulate && !p.euclid;
    if (plain && VPL == 1 && p.hint >= 2) {          // occupancy / unroll experiments (rsb200_pair_args.variant 40..44)
        const unsigned chunks = blocks;              // caller passes the uncapped block count for these
        auto grid = [&](int occ) { return (unsigned)min((int64_t)chunks, (int64_t)sm_count() * occ); };
        switch (p.hint) {
            case 2: scatter_kernel<1, true, false, 4, 4, false><<<grid(8), kScatWarps * 32, 0, st>>>(p); break;   // previous default
            case 3: scatter_kernel<1, true, false, 5, 4><<<grid(5), kScatWarps * 32, 0, st>>>(p); break;
            case 4: scatter_kernel<1, true, false, 6, 4><<<grid(6), kScatWarps * 32, 0, st>>>(p); break;
            case 5: scatter_kernel<1, true, false, 6, 2><<<grid(6), kScatWarps * 32, 0, st>>>(p); break;
            default: scatter_kernel<1, true, false, 5, 8><<<grid(5), kScatWarps * 32, 0, st>>>(p); break;
        }
        return;
    }
    if (plain && p.hint && VPL == 1) scatter_kernel<VPL, true, true><<<blocks, kScatWarps * 32, 0, st>>>(p);
    else if (plain && p.D == 128 * VPL) scatter_kernel<VPL, true, false, 4, 4, true><<<blocks, kScatWarps * 32, 0, st>>>(p);
    else if (plain) scatter_kernel<VPL, true><<<blocks, kScatWarps * 32, 0, st>>>(p);
    else scatter_kernel<VPL, false><<<blocks, kScatWarps * 32, 0, st>>>(p);
}

int32_t launch_scatter(const ScatterParams& p, int64_t cap_rows, cudaStream_t st) {
    if (cap_rows <= 0) return 0;
    int64_t chunks = cdiv(cap_rows, 32);
    int64_t blocks = cdiv(chunks, kScatWarps);
    int64_t max_blocks = (int64_t)sm_count() * 8;
    if (blocks > max_blocks && p.hint < 2) blocks = max_blocks;
    if (p.D <= 128) launch_scatter_v<1>(p, (unsigned)blocks, st);
    else if (p.D <= 256) launch_scatter_v<2>(p, (unsigned)blocks, st);
    else if (p.D <= 512) launch_scatter_v<4>(p, (unsigned)blocks, st);
    else { set_error("embedding dim %d > 512 is not supported", p.D); return RSB200_EUNSUPPORTED; }
    RSB_LAUNCH_CHECK();
    return 0;
}

int32_t launch_loss_sum(const float* part, int B, float* loss, cudaStream_t st) {
    loss_sum_kernel<<<1, 256, 0, st>>>(part, B, loss);
    RSB_LAUNCH_CHECK();
    return 0;
}

}  // namespace rsb
