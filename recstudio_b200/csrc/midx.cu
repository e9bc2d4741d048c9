// midx.cu -- index build of the model-based samplers: Sampler.update() of
// MIDXSamplerUniform / MIDXSamplerPop / ClusterSamplerUniform / ClusterSamplerPop, called once per epoch
// (recstudio/model/basemodel/recommender.py:561-570).
//
//   kmeans           recstudio/ann/sampler.py:9-36    Lloyd iterations: [N,K] distances -> argmin -> centroid means
//   construct_index  recstudio/ann/sampler.py:39-45   stable sort of the cluster codes -> (indices, indptr)
//   _update          recstudio/ann/sampler.py:296-306,410-423  per-cluster weight wkk and the normalised per-cluster
//                    cumulative distribution cp (a Python `for c in range(K**2)` loop in the reference)
//   _sample_item_with_pop  sampler.py:348-365         per-draw inverse-CDF search inside the chosen cluster
//
// The reference materialises the [N,K] distance matrix, the [N,K] one-hot assignment matrix (a GEMM against
// it produces the centroid sums and wkk) and, per draw, a [num_q, neg, max_cluster_size] gather.  Here:
//   * kmeans_assign: argmin and the exact sum of squared residuals in the epilogue of the distance contraction -- no
//     [N,K] matrix.  Small codebooks (the MIDX case, D <= 64): one thread per point, centroids broadcast from shared
//     memory; otherwise an fp32 FFMA tile GEMM (128 points x 128 centroids per CTA);
//   * kmeans_update: per-CTA shared-memory centroid accumulators (shared atomics), one global flush per CTA;
//   * index build: stable LSD radix sort by 8-bit digits (warp match_any ranks, no atomics on the order)
//     -> identical to torch.sort(stable=True);
//   * segment_cdf: one warp per cluster, warp-scan cumulative sums, normalised in place;
//   * segment_search: one thread per draw, bisection inside the cluster's slice of cp.
#include "common.cuh"
#include "kernels.h"
#include "tile_gemm.cuh"

namespace rsb {

// ------------------------------------------------------------------------------------------ kmeans
__global__ void __launch_bounds__(256)
row_sqnorm_kernel(const float* __restrict__ c, int K, int D, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (k >= K) return;
    float a = 0.f;
    for (int j = lane; j < D; j += 32) { const float v = c[(size_t)k * D + j]; a = fmaf(v, v, a); }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) a += __shfl_xor_sync(kFull, a, o);
    if (lane == 0) out[k] = a;
}

// assign[i] = argmin_k ( |x_i|^2 - 2 x_i.c_k + |c_k|^2 )  (first minimum on ties, sampler.py:19-21),
// loss += sum_i |x_i - c_assign(i)|^2 evaluated directly from the residuals (sampler.py:24).
__global__ void __launch_bounds__(256, 2)
kmeans_assign_kernel(const float* __restrict__ x, int ldx, int N, int D, const float* __restrict__ c, int K,
                     const float* __restrict__ cn, int64_t* __restrict__ assign, double* __restrict__ loss) {
    using namespace tg;
    __shared__ __align__(16) float As[2][TK][LDS_];
    __shared__ __align__(16) float Bs[2][TK][LDS_];
    __shared__ float s_xn[TM];
    __shared__ int s_assign[TM];
    __shared__ double s_red[8];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.x * TM;

    // |x_i|^2 of the tile's points: two threads per point
    {
        const int p = tid >> 1, h = tid & 1, row = m0 + p;
        float a = 0.f;
        if (row < N)
            for (int j = h; j < D; j += 2) { const float v = x[(size_t)row * ldx + j]; a = fmaf(v, v, a); }
        a += __shfl_xor_sync(kFull, a, 1);
        if (h == 0) s_xn[p] = a;
    }
    __syncthreads();

    float bestv[8]; int besti[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { bestv[i] = INFINITY; besti[i] = 0; }
    for (int n0 = 0; n0 < K; n0 += TN) {
        float acc[8][8];
        zero_acc(acc);
        gemm_nt(acc, x, N, m0, c, K, n0, D, As, Bs, ldx, D);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = n0 + ty * 8 + j;
            if (col < K) {
                const float cj = __ldg(cn + col);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float dist = (s_xn[row_of(tx, i)] - 2.f * acc[i][j]) + cj;
                    if (dist < bestv[i]) { bestv[i] = dist; besti[i] = col; }
                }
            }
        }
    }
    // reduce over the 16 threads (ty) that share a point: staging buffers are free after gemm_nt's barrier
    float* rv = &As[0][0][0];                       // [128][16]
    int* ri = reinterpret_cast<int*>(&Bs[0][0][0]);
#pragma unroll
    for (int i = 0; i < 8; ++i) { rv[row_of(tx, i) * 16 + ty] = bestv[i]; ri[row_of(tx, i) * 16 + ty] = besti[i]; }
    __syncthreads();
    if (tid < TM) {
        float bv = rv[tid * 16]; int bi = ri[tid * 16];
        for (int t = 1; t < 16; ++t) {
            const float v = rv[tid * 16 + t]; const int ix = ri[tid * 16 + t];
            if (v < bv || (v == bv && ix < bi)) { bv = v; bi = ix; }
        }
        s_assign[tid] = bi;
        if (m0 + tid < N) assign[m0 + tid] = bi;
    }
    __syncthreads();
    double part = 0.0;
    {
        const int p = tid >> 1, h = tid & 1, row = m0 + p;
        if (row < N) {
            const float* cr = c + (size_t)s_assign[p] * D;
            float a = 0.f;
            for (int j = h; j < D; j += 2) { const float df = x[(size_t)row * ldx + j] - cr[j]; a = fmaf(df, df, a); }
            part = (double)a;
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) part += __shfl_xor_sync(kFull, part, o);
    if (lane == 0) s_red[warp] = part;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_red[w];
        atomicAdd(loss, t);
    }
}

// Small-codebook form (D in {16, 32, 64}, K * D floats in shared memory): the MIDX case, where each codebook sees one
// HALF of the item vector (d = 128 -> D = 64) and K is a few dozen.  One thread per point: the point's D values live in
// registers, the centroids are read from shared memory with broadcast 16-byte loads (4 FMAs per load, 4 centroids
// interleaved), so the kernel runs on the FMA pipe instead of paying a 128 x 128 tile's prologue / epilogue per 4 k-steps.
template <int DV>     // DV = D / 4
__global__ void __launch_bounds__(128)
kmeans_assign_small_kernel(const float* __restrict__ x, int ldx, int N, const float* __restrict__ c, int K,
                           const float* __restrict__ cn, int64_t* __restrict__ assign, double* __restrict__ loss) {
    constexpr int D = DV * 4, XS = D + 4;            // padded row stride of the staged point tile (16-byte aligned rows)
    extern __shared__ __align__(16) float sm[];
    float* s_c = sm;                                  // [K][D]
    float* s_x = sm + (size_t)K * D;                  // [128][XS]
    __shared__ double s_red[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.x * 128;
    for (int i = tid * 4; i < K * D; i += 128 * 4) *reinterpret_cast<float4*>(s_c + i) = ldg128(c + i);
    // coalesced staging of the tile: consecutive threads read consecutive 16-byte pieces of a row
    for (int i = tid; i < 128 * DV; i += 128) {
        const int r = i / DV, v = i - r * DV;
        float4 val = make_float4(0, 0, 0, 0);
        if (m0 + r < N) val = ldg128_stream(x + (size_t)(m0 + r) * ldx + v * 4);
        *reinterpret_cast<float4*>(s_x + (size_t)r * XS + v * 4) = val;
    }
    __syncthreads();
    float4 xv[DV];
    float xn = 0.f;
#pragma unroll
    for (int v = 0; v < DV; ++v) {
        xv[v] = *reinterpret_cast<const float4*>(s_x + (size_t)tid * XS + v * 4);
        xn = fmaf(xv[v].x, xv[v].x, xn); xn = fmaf(xv[v].y, xv[v].y, xn);
        xn = fmaf(xv[v].z, xv[v].z, xn); xn = fmaf(xv[v].w, xv[v].w, xn);
    }
    float bestv = INFINITY; int besti = 0;
    for (int k0 = 0; k0 < K; k0 += 4) {
        float dot[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int v = 0; v < DV; ++v) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = min(k0 + u, K - 1);
                const float4 cv = *reinterpret_cast<const float4*>(s_c + (size_t)k * D + v * 4);
                dot[u] = fmaf(xv[v].x, cv.x, dot[u]); dot[u] = fmaf(xv[v].y, cv.y, dot[u]);
                dot[u] = fmaf(xv[v].z, cv.z, dot[u]); dot[u] = fmaf(xv[v].w, cv.w, dot[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (k0 + u < K) {
                const float dist = (xn - 2.f * dot[u]) + __ldg(cn + k0 + u);
                if (dist < bestv) { bestv = dist; besti = k0 + u; }
            }
        }
    }
    double part = 0.0;
    if (m0 + tid < N) {
        assign[m0 + tid] = besti;
        float a = 0.f;
#pragma unroll
        for (int v = 0; v < DV; ++v) {
            const float4 cv = *reinterpret_cast<const float4*>(s_c + (size_t)besti * D + v * 4);
            const float dx = xv[v].x - cv.x, dy = xv[v].y - cv.y, dz = xv[v].z - cv.z, dw = xv[v].w - cv.w;
            a = fmaf(dx, dx, a); a = fmaf(dy, dy, a); a = fmaf(dz, dz, a); a = fmaf(dw, dw, a);
        }
        part = (double)a;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) part += __shfl_xor_sync(kFull, part, o);
    if (lane == 0) s_red[warp] = part;
    __syncthreads();
    if (tid == 0) atomicAdd(loss, s_red[0] + s_red[1] + s_red[2] + s_red[3]);
}

// sums[k, :] += x_i for assign[i] == k, counts[k] += 1  (the reference's assign_m.T @ X and assign_m.sum(0),
// sampler.py:30-31); per-CTA accumulators in shared memory, one flush per CTA.
// One accumulator set per CTA in shared memory (shared fp32 atomics; two warps rarely meet on the same (cluster, column)),
// kept small so that several CTAs fit an SM, and kUpdRows independent row loads in flight per warp: the kernel streams
// the points once and is bound by DRAM latency x bytes in flight, not by the atomics.
constexpr int kUpdRows = 8;

__global__ void __launch_bounds__(256)
kmeans_update_kernel(const float* __restrict__ x, int ldx, int N, int D, const int64_t* __restrict__ assign, int K,
                     float* __restrict__ sums, float* __restrict__ counts) {
    extern __shared__ float sh[];                   // [K*D] sums, [K] counts
    float* s_sum = sh;
    float* s_cnt = sh + (size_t)K * D;
    for (int i = threadIdx.x; i < K * D + K; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    for (int row0 = (blockIdx.x * wpb + warp) * kUpdRows; row0 < N; row0 += gridDim.x * wpb * kUpdRows) {
        int a[kUpdRows];
#pragma unroll
        for (int r = 0; r < kUpdRows; ++r) {
            const int64_t v = (row0 + r < N) ? assign[row0 + r] : -1;
            a[r] = (v >= 0 && v < K) ? (int)v : -1;
        }
        for (int j = lane; j < D; j += 32) {
            float v[kUpdRows];
#pragma unroll
            for (int r = 0; r < kUpdRows; ++r) v[r] = (a[r] >= 0) ? x[(size_t)(row0 + r) * ldx + j] : 0.f;
#pragma unroll
            for (int r = 0; r < kUpdRows; ++r)
                if (a[r] >= 0) atomicAdd(&s_sum[(size_t)a[r] * D + j], v[r]);
        }
#pragma unroll
        for (int r = 0; r < kUpdRows; ++r)
            if (lane == r && a[r] >= 0) atomicAdd(&s_cnt[a[r]], 1.0f);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * D + K; i += blockDim.x) {
        const float t = sh[i];
        if (t != 0.f) atomicAdd(i < K * D ? &sums[i] : &counts[i - K * D], t);
    }
}

// ------------------------------------------------------------------------------------------ construct_index
constexpr int kRadixThreads = 256, kRadixPer = 8, kRadixTile = kRadixThreads * kRadixPer;   // 2048 keys per CTA

__global__ void __launch_bounds__(256)
index_keys_kernel(const int64_t* __restrict__ codes, int64_t N, int nb, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                  uint32_t* __restrict__ counts, uint32_t* __restrict__ err_flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int64_t c = codes[i];
    if (c < 0 || c >= nb) { *err_flag = 1u; c = 0; }
    keys[i] = (uint32_t)c;
    vals[i] = (uint32_t)i;
    atomicAdd(counts + c, 1u);
}

__global__ void __launch_bounds__(kRadixThreads)
radix_hist_kernel(const uint32_t* __restrict__ keys, int64_t N, int shift, uint32_t* __restrict__ tile_hist, int num_tiles) {
    __shared__ uint32_t hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kRadixTile;
#pragma unroll
    for (int r = 0; r < kRadixPer; ++r) {
        const int64_t i = base + (int64_t)r * kRadixThreads + threadIdx.x;
        if (i < N) atomicAdd(&hist[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    tile_hist[(size_t)threadIdx.x * num_tiles + blockIdx.x] = hist[threadIdx.x];     // digit-major: scan order = (digit, tile)
}

// stable scatter of one 8-bit digit: warp w owns keys [base + 256 w, +256) in index order; the rank of a key among
// the equal digits before it = (earlier warps) + (earlier rounds of this warp) + (lower lanes of this round)
__global__ void __launch_bounds__(kRadixThreads)
radix_scatter_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t N, int shift,
                     const uint32_t* __restrict__ tile_off, int num_tiles, uint32_t* __restrict__ keys_out,
                     uint32_t* __restrict__ vals_out) {
    __shared__ uint32_t whist[8][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 256; i += kRadixThreads) (&whist[0][0])[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kRadixTile + (int64_t)w * 256;
    uint32_t key[kRadixPer], val[kRadixPer], rank[kRadixPer];
#pragma unroll
    for (int r = 0; r < kRadixPer; ++r) {
        const int64_t i = base + r * 32 + lane;
        const bool valid = i < N;
        key[r] = valid ? keys[i] : 0u;
        val[r] = valid ? vals[i] : 0u;
        const uint32_t digit = (key[r] >> shift) & 255u;
        const uint32_t mask = __match_any_sync(kFull, valid ? digit : 0x1000u);
        const uint32_t before = valid ? whist[w][digit] : 0u;
        __syncwarp();
        rank[r] = before + __popc(mask & ((1u << lane) - 1u));
        if (valid && lane == __ffs(mask) - 1) whist[w][digit] = before + __popc(mask);
        __syncwarp();
    }
    __syncthreads();
    {
        const int d = threadIdx.x;
        uint32_t run = tile_off[(size_t)d * num_tiles + blockIdx.x];
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) { const uint32_t t = whist[ww][d]; whist[ww][d] = run; run += t; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRadixPer; ++r) {
        const int64_t i = base + r * 32 + lane;
        if (i < N) {
            const uint32_t dst = whist[w][(key[r] >> shift) & 255u] + rank[r];
            keys_out[dst] = key[r];
            vals_out[dst] = val[r];
        }
    }
}

// indptr = exclusive scan of the bucket counts (single CTA), indices = sorted item positions as int64
__global__ void __launch_bounds__(1024)
index_finish_kernel(const uint32_t* __restrict__ counts, int nb, int64_t* __restrict__ indptr) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < nb ? counts[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_warp[w] = inc;
        __syncthreads();
        if (w == 0) {
            uint32_t s = s_warp[lane], si = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(kFull, si, o); if (lane >= o) si += t; }
            s_warp[lane] = si - s;
        }
        __syncthreads();
        const uint32_t excl = s_carry + s_warp[w] + inc - v;
        if (i < nb) indptr[i] = (int64_t)excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) indptr[nb] = (int64_t)s_carry;
}

__global__ void __launch_bounds__(256)
widen_kernel(const uint32_t* __restrict__ in, int64_t N, int64_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) out[i] = (int64_t)in[i];
}

// ------------------------------------------------------------------------------------------ per-cluster CDF
// cp[e] = cumsum_{e' <= e in cluster}(weight[indices[e']]) / cluster total, total[c] = cluster total
// (sampler.py:300-306: `for c in range(K**2): cumsum = cp[start:end].cumsum(0); cp[start:end] = cumsum / cumsum[-1]`)
__global__ void __launch_bounds__(256)
segment_cdf_kernel(const float* __restrict__ weight, const int64_t* __restrict__ indices, const int64_t* __restrict__ indptr,
                   int nb, float* __restrict__ cp, float* __restrict__ total) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (c >= nb) return;
    const int64_t start = indptr[c], end = indptr[c + 1];
    float carry = 0.f;
    for (int64_t e0 = start; e0 < end; e0 += 32) {
        const int64_t e = e0 + lane;
        float v = e < end ? __ldg(weight + indices[e]) : 0.f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(kFull, v, o); if (lane >= o) v += t; }
        v += carry;
        if (e < end) cp[e] = v;
        carry = __shfl_sync(kFull, v, 31);
    }
    if (lane == 0 && total) total[c] = carry;
    __syncwarp();
    for (int64_t e = start + lane; e < end; e += 32) cp[e] = cp[e] / carry;
}

// _sample_item_with_pop (sampler.py:348-365): item_idx = first position in the cluster with cp >= u,
// neg = indices[start + item_idx] (no +1: the reference's convention here), logp = log(p[start + item_idx + 1])
__global__ void __launch_bounds__(256)
segment_search_kernel(const int64_t* __restrict__ k01, const float* __restrict__ u, int64_t M, const float* __restrict__ cp,
                      const int64_t* __restrict__ indices, const int64_t* __restrict__ indptr, int nb,
                      const float* __restrict__ p, int64_t* __restrict__ neg, float* __restrict__ logp) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    int64_t c = k01[i];
    if (c < 0 || c >= nb) c = 0;
    const int64_t start = indptr[c], last = indptr[c + 1] - 1;
    const float uu = u[i];
    int64_t lo = 0, hi = last - start + 1;              // local positions
    if (hi < 0) hi = 0;
    const int64_t cnt = hi;
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(cp + start + mid) < uu) lo = mid + 1; else hi = mid;
    }
    if (lo >= cnt && cnt > 0) lo = cnt - 1;              // the clamped tail of `fullrange` repeats the last entry
    int64_t idx = lo < last ? lo : last;                 // torch.minimum(item_idx, last)   (:361)
    if (idx < 0) idx = 0;
    neg[i] = indices[idx + start];
    logp[i] = logf(__ldg(p + idx + start + 1));
}

}  // namespace rsb

using namespace rsb;

extern "C" int32_t rsb200_kmeans_assign(const float* x, int64_t ldx, int64_t num_points, int64_t d, const float* centers,
                                        int64_t num_clusters, float* cnorm_ws, int64_t* assign_out, double* loss_out,
                                        void* stream) {
    RSB_REQUIRE(x && centers && cnorm_ws && assign_out && loss_out, RSB200_EINVAL, "null pointer");
    RSB_REQUIRE(d >= 4 && d % 4 == 0 && ldx >= d && ldx % 4 == 0 && aligned16(x) && aligned16(centers), RSB200_EINVAL,
                "kmeans needs d %% 4 == 0 and 16-byte aligned rows (d=%lld ldx=%lld)", (long long)d, (long long)ldx);
    RSB_REQUIRE(num_points >= 1 && num_points < ((int64_t)1 << 31) && num_clusters >= 1 && num_clusters <= 65536, RSB200_EINVAL,
                "bad kmeans shape");
    cudaStream_t st = (cudaStream_t)stream;
    RSB_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(double), st));
    row_sqnorm_kernel<<<(unsigned)cdiv(num_clusters, 8), 256, 0, st>>>(centers, (int)num_clusters, (int)d, cnorm_ws);
    RSB_LAUNCH_CHECK();
    const size_t small_smem = (size_t)(num_clusters * d + 128 * (d + 4)) * sizeof(float);
    if ((d == 16 || d == 32 || d == 64) && small_smem <= 96 * 1024) {
        const unsigned grid = (unsigned)cdiv(num_points, 128);
#define RSB_SMALL(DV)                                                                                                       \
        do {                                                                                                                \
            RSB_CUDA(cudaFuncSetAttribute(kmeans_assign_small_kernel<DV>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                          (int)small_smem));                                                                \
            kmeans_assign_small_kernel<DV><<<grid, 128, small_smem, st>>>(x, (int)ldx, (int)num_points, centers,           \
                                                                         (int)num_clusters, cnorm_ws, assign_out, loss_out); \
        } while (0)
        if (d == 16) RSB_SMALL(4); else if (d == 32) RSB_SMALL(8); else RSB_SMALL(16);
#undef RSB_SMALL
        RSB_LAUNCH_CHECK();
        return 0;
    }
    kmeans_assign_kernel<<<(unsigned)cdiv(num_points, tg::TM), 256, 0, st>>>(x, (int)ldx, (int)num_points, (int)d, centers,
                                                                            (int)num_clusters, cnorm_ws, assign_out, loss_out);
    RSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int32_t rsb200_kmeans_update(const float* x, int64_t ldx, int64_t num_points, int64_t d, const int64_t* assign,
                                        int64_t num_clusters, float* sums_out, float* counts_out, void* stream) {
    RSB_REQUIRE(x && assign && sums_out && counts_out, RSB200_EINVAL, "null pointer");
    RSB_REQUIRE(d >= 1 && ldx >= d && num_points >= 1 && num_points < ((int64_t)1 << 31) && num_clusters >= 1, RSB200_EINVAL,
                "bad kmeans shape");
    const size_t set_bytes = (size_t)(num_clusters * d + num_clusters) * sizeof(float);
    RSB_REQUIRE(set_bytes <= 200 * 1024, RSB200_EUNSUPPORTED, "num_clusters * d = %lld floats exceed the shared-memory accumulator",
                (long long)(num_clusters * d));
    const size_t smem = set_bytes;
    cudaStream_t st = (cudaStream_t)stream;
    RSB_CUDA(cudaMemsetAsync(sums_out, 0, sizeof(float) * (size_t)(num_clusters * d), st));
    RSB_CUDA(cudaMemsetAsync(counts_out, 0, sizeof(float) * (size_t)num_clusters, st));
    int64_t per_sm = (int64_t)((200 * 1024) / smem);
    if (per_sm > 8) per_sm = 8;
    int64_t blocks = cdiv(num_points, 8 * 64);            // >= 64 points per warp before a flush
    if (blocks > (int64_t)sm_count() * per_sm) blocks = (int64_t)sm_count() * per_sm;
    if (blocks < 1) blocks = 1;
    if (smem > 48 * 1024)
        RSB_CUDA(cudaFuncSetAttribute(kmeans_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kmeans_update_kernel<<<(unsigned)blocks, 256, smem, st>>>(x, (int)ldx, (int)num_points, (int)d, assign, (int)num_clusters,
                                                             sums_out, counts_out);
    RSB_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t rsb200_index_workspace_bytes(int64_t num_points, int64_t num_buckets) {
    if (num_points < 0 || num_buckets < 1) return 0;
    const int64_t tiles = cdiv(num_points > 0 ? num_points : 1, kRadixTile);
    const int64_t hist = 256 * tiles + 1;
    // keys[2][N] vals[2][N] counts[nb] tile_hist[hist] totals[4] err[1] + scan_tmp (u64)
    size_t words = (size_t)(4 * num_points + num_buckets + hist + 8);
    words = (words + 3) & ~(size_t)3;
    return words * 4 + (size_t)scan_tmp_elems(hist) * 8 + 64;
}

extern "C" int32_t rsb200_index_build(const int64_t* codes, int64_t num_points, int64_t num_buckets, int64_t* indices_out,
                                      int64_t* indptr_out, void* workspace, size_t workspace_bytes, void* stream) {
    RSB_REQUIRE(codes && indices_out && indptr_out && workspace, RSB200_EINVAL, "null pointer");
    RSB_REQUIRE(num_points >= 1 && num_points < ((int64_t)1 << 31) && num_buckets >= 1 && num_buckets <= 65536, RSB200_EUNSUPPORTED,
                "index build supports up to 2^31 points and 65536 buckets (got %lld, %lld)", (long long)num_points, (long long)num_buckets);
    RSB_REQUIRE(workspace_bytes >= rsb200_index_workspace_bytes(num_points, num_buckets), RSB200_EWORKSPACE, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t N = num_points;
    const int tiles = (int)cdiv(N, kRadixTile);
    const int64_t hist = 256 * (int64_t)tiles + 1;
    uint32_t* w32 = reinterpret_cast<uint32_t*>(workspace);
    uint32_t* keys[2] = {w32, w32 + N};
    uint32_t* vals[2] = {w32 + 2 * N, w32 + 3 * N};
    uint32_t* counts = w32 + 4 * N;
    uint32_t* tile_hist = counts + num_buckets;
    uint32_t* totals = tile_hist + hist;
    uint32_t* err = totals + 4;
    size_t words = (size_t)(4 * N + num_buckets + hist + 8);
    words = (words + 3) & ~(size_t)3;
    uint64_t* scan_tmp = reinterpret_cast<uint64_t*>(w32 + words);
    const int64_t tmp_elems = scan_tmp_elems(hist);
    RSB_CUDA(cudaMemsetAsync(counts, 0, sizeof(uint32_t) * (size_t)num_buckets, st));
    RSB_CUDA(cudaMemsetAsync(err, 0, sizeof(uint32_t), st));
    index_keys_kernel<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(codes, N, (int)num_buckets, keys[0], vals[0], counts, err);
    RSB_LAUNCH_CHECK();
    int cur = 0;
    for (int shift = 0; shift < 16 && (num_buckets - 1) >> shift; shift += 8) {
        radix_hist_kernel<<<tiles, kRadixThreads, 0, st>>>(keys[cur], N, shift, tile_hist, tiles);
        RSB_LAUNCH_CHECK();
        int32_t rc = launch_scan(tile_hist, hist - 1, nullptr, 0, totals, scan_tmp, tmp_elems, st);
        if (rc) return rc;
        radix_scatter_kernel<<<tiles, kRadixThreads, 0, st>>>(keys[cur], vals[cur], N, shift, tile_hist, tiles, keys[cur ^ 1],
                                                              vals[cur ^ 1]);
        RSB_LAUNCH_CHECK();
        cur ^= 1;
    }
    widen_kernel<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(vals[cur], N, indices_out);
    RSB_LAUNCH_CHECK();
    index_finish_kernel<<<1, 1024, 0, st>>>(counts, (int)num_buckets, indptr_out);
    RSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int32_t rsb200_segment_cdf(const float* weight, const int64_t* indices, const int64_t* indptr, int64_t num_buckets,
                                      float* cp_out, float* total_out, void* stream) {
    RSB_REQUIRE(weight && indices && indptr && cp_out, RSB200_EINVAL, "null pointer");
    RSB_REQUIRE(num_buckets >= 1 && num_buckets < ((int64_t)1 << 31), RSB200_EINVAL, "bad num_buckets");
    segment_cdf_kernel<<<(unsigned)cdiv(num_buckets, 8), 256, 0, (cudaStream_t)stream>>>(weight, indices, indptr, (int)num_buckets,
                                                                                        cp_out, total_out);
    RSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int32_t rsb200_segment_search(const int64_t* k01, const float* u, int64_t num_draws, const float* cp,
                                         const int64_t* indices, const int64_t* indptr, int64_t num_buckets, const float* p,
                                         int64_t* neg_out, float* logp_out, void* stream) {
    RSB_REQUIRE(k01 && u && cp && indices && indptr && p && neg_out && logp_out, RSB200_EINVAL, "null pointer");
    if (num_draws == 0) return 0;
    segment_search_kernel<<<(unsigned)cdiv(num_draws, 256), 256, 0, (cudaStream_t)stream>>>(k01, u, num_draws, cp, indices, indptr,
                                                                                           (int)num_buckets, p, neg_out, logp_out);
    RSB_LAUNCH_CHECK();
    return 0;
}
