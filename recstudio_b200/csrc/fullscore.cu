// fullscore.cu -- full-catalog kernels.
//
// T1  BaseRetriever.topk (recstudio/model/basemodel/baseretriever.py:374-397):
//       score = score_func(query, item_vector)        [Be, N-1]   (scorer.py:15-16)
//       torch.topk(k + H) -> +1 id shift -> mask user history -> topk(k)
//     Here the [Be, N-1] score matrix never reaches HBM.  Three kernels:
//       1. score_gmax : fp32 FFMA tile GEMM (128 queries x 128 items per CTA, 8x8 per thread);
//                       the epilogue keeps only the max of every group of 8 consecutive items
//                       -> gmax[Be, ceil((N-1)/8)]  (1/8 of the score matrix, 4 B per group)
//       2. select_groups : per query, exact radix select of the K' = k + H + 8 largest group
//                       maxima.  The k + H best ITEMS of a query all live in those groups
//                       (each selected group holds at least one item >= the K'-th group max),
//                       which is exactly the candidate set the reference's topk(k + H) needs.
//       3. topk_final : re-score the 8 K' candidates exactly (fp32, one warp per row), drop
//                       the ids found in the user's history, bitonic-sort by (score desc,
//                       id asc) and emit k (score, 1-based id) pairs.
//     The contraction runs on the fp32 pipe on purpose: ranks must match the reference's
//     fp32 scores and the north star keeps tensor cores for the attention block only.
//
// L3 (full softmax) lives in fullsoftmax.cu.
#include "common.cuh"
#include "kernels.h"

namespace rsb {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;
constexpr int kGroup = 8;                       // items per group-max

// C[m][n] = sum_k A[m][k] * B[n][k]  (both operands K-contiguous), 128x128 tile, 256 threads.
// thread (tx = tid % 16, ty = tid / 16): queries {tx*4+i, 64+tx*4+i}, items {ty*8 + j}.
template <bool EUCLID>
__global__ void __launch_bounds__(256, 2)
score_gmax_kernel(const float* __restrict__ q, const float* __restrict__ w1 /* first ITEM row (table row 1) */,
                  int M, int Nit, int D, float* __restrict__ gmax, int ngroups) {
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int lr = tid >> 2, lk = (tid & 3) * 4;          // loader: rows lr, lr+64; k offset lk

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float qn[8], vn[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { qn[i] = 0.f; vn[i] = 0.f; }

    auto gload = [&](const float* base, int rows, int r, int k0) -> float4 {
        // 16-byte load of base[r][k0 + lk .. +3], zero outside the matrix
        if (r < rows && k0 + lk < D) return ldg128(base + (size_t)r * D + k0 + lk);
        return make_float4(0, 0, 0, 0);
    };
    float4 ra0 = gload(q, M, m0 + lr, 0), ra1 = gload(q, M, m0 + lr + 64, 0);
    float4 rb0 = gload(w1, Nit, n0 + lr, 0), rb1 = gload(w1, Nit, n0 + lr + 64, 0);
    const int ksteps = (D + BK - 1) / BK;
    for (int ks = 0; ks < ksteps; ++ks) {
        const int buf = ks & 1;
        As[buf][lk + 0][lr] = ra0.x; As[buf][lk + 1][lr] = ra0.y; As[buf][lk + 2][lr] = ra0.z; As[buf][lk + 3][lr] = ra0.w;
        As[buf][lk + 0][lr + 64] = ra1.x; As[buf][lk + 1][lr + 64] = ra1.y; As[buf][lk + 2][lr + 64] = ra1.z; As[buf][lk + 3][lr + 64] = ra1.w;
        Bs[buf][lk + 0][lr] = rb0.x; Bs[buf][lk + 1][lr] = rb0.y; Bs[buf][lk + 2][lr] = rb0.z; Bs[buf][lk + 3][lr] = rb0.w;
        Bs[buf][lk + 0][lr + 64] = rb1.x; Bs[buf][lk + 1][lr + 64] = rb1.y; Bs[buf][lk + 2][lr + 64] = rb1.z; Bs[buf][lk + 3][lr + 64] = rb1.w;
        __syncthreads();                                    // one barrier per k-step: buffers alternate
        if (ks + 1 < ksteps) {
            const int k0 = (ks + 1) * BK;
            ra0 = gload(q, M, m0 + lr, k0); ra1 = gload(q, M, m0 + lr + 64, k0);
            rb0 = gload(w1, Nit, n0 + lr, k0); rb1 = gload(w1, Nit, n0 + lr + 64, k0);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][tx * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + tx * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][ty * 8]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][ty * 8 + 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            if (EUCLID) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { qn[i] = fmaf(a[i], a[i], qn[i]); vn[i] = fmaf(b[i], b[i], vn[i]); }
            }
        }
    }
    // epilogue: max over this thread's 8 consecutive items, per query
    const int g = n0 / kGroup + ty;
    if (g >= ngroups) return;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ((i < 4) ? tx * 4 + i : 64 + tx * 4 + (i - 4));
        if (m >= M) continue;
        float best = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int nidx = n0 + ty * 8 + j;
            float s = acc[i][j];
            if (EUCLID) s = 2.f * s - vn[j] - qn[i];        // -(|q|^2 - 2 q.v + |v|^2)
            if (nidx < Nit) best = fmaxf(best, s);
        }
        gmax[(size_t)m * ngroups + g] = best;
    }
}

// order-preserving float -> uint (ascending)
__device__ __forceinline__ uint32_t fkey(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// one CTA per query: indices of the K largest values of vals[0..n) -> out[0..K) (any order)
__global__ void __launch_bounds__(1024)
select_groups_kernel(const float* __restrict__ gmax, int ngroups, int K, int32_t* __restrict__ cand) {
    __shared__ uint32_t hist[2048];
    __shared__ uint32_t s_prefix, s_need, s_cnt_gt, s_cnt_eq;
    const float* vals = gmax + (size_t)blockIdx.x * ngroups;
    int32_t* out = cand + (size_t)blockIdx.x * K;
    const int tid = threadIdx.x;
    if (K >= ngroups) {                                   // everything is a candidate
        for (int i = tid; i < K; i += 1024) out[i] = (i < ngroups) ? i : -1;
        return;
    }
    uint32_t prefix = 0, pmask = 0, need = (uint32_t)K;
    const int shifts[3] = {21, 10, 0};
    const int bits[3] = {11, 11, 10};
    for (int pass = 0; pass < 3; ++pass) {
        const int sh = shifts[pass], nb = 1 << bits[pass];
        for (int i = tid; i < 2048; i += 1024) hist[i] = 0;
        __syncthreads();
        for (int i = tid; i < ngroups; i += 1024) {
            const uint32_t key = fkey(vals[i]);
            if ((key & pmask) == prefix) atomicAdd(&hist[(key >> sh) & (nb - 1)], 1u);
        }
        __syncthreads();
        if (tid < 32) {
            // find the bin (from the top) where the running count reaches `need`: lane L owns the L-th highest
            // block of nb/32 bins; one warp scan over the block sums, then one lane walks its <= 64 bins
            const int per = nb >> 5, hi = nb - tid * per, lo = hi - per;
            uint32_t local = 0;
            for (int bn = lo; bn < hi; ++bn) local += hist[bn];
            uint32_t pre = local;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(kFull, pre, o);
                if (tid >= o) pre += t;
            }
            const unsigned m = __ballot_sync(kFull, pre >= need);      // non-empty: the total count is >= need
            if (tid == __ffs(m) - 1) {
                uint32_t acc = pre - local; int bsel = lo;
                for (int bn = hi - 1; bn >= lo; --bn) {
                    if (acc + hist[bn] >= need) { bsel = bn; break; }
                    acc += hist[bn];
                }
                s_prefix = prefix | ((uint32_t)bsel << sh);
                s_need = need - acc;
            }
        }
        __syncthreads();
        prefix = s_prefix; need = s_need;
        pmask |= (uint32_t)(nb - 1) << sh;
        __syncthreads();
    }
    // prefix is now the exact key of the K-th largest value; `need` of the values equal to it are wanted
    if (tid == 0) { s_cnt_gt = 0; s_cnt_eq = 0; }
    __syncthreads();
    const uint32_t n_gt = (uint32_t)K - need;
    for (int i = tid; i < ngroups; i += 1024) {
        const uint32_t key = fkey(vals[i]);
        if (key > prefix) {
            const uint32_t pos = atomicAdd(&s_cnt_gt, 1u);
            if (pos < n_gt) out[pos] = i;
        } else if (key == prefix) {
            const uint32_t pos = atomicAdd(&s_cnt_eq, 1u);
            if (pos < need) out[n_gt + pos] = i;
        }
    }
}

// one CTA (256 threads) per query: exact scores of the candidates, history mask, sort, emit k
template <bool EUCLID>
__global__ void __launch_bounds__(256)
topk_final_kernel(const float* __restrict__ q, const float* __restrict__ w_item, int num_items, int D,
                  const int32_t* __restrict__ cand, int K, const int64_t* __restrict__ hist, int H, int k,
                  float* __restrict__ score_out, int64_t* __restrict__ id_out, int ncand_pow2) {
    extern __shared__ unsigned long long skeys[];          // [ncand_pow2] sort keys
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* qr = q + (size_t)b * D;
    const int64_t* hr = hist ? hist + (size_t)b * H : nullptr;
    const int ncand = K * kGroup;
    for (int i = tid; i < ncand_pow2; i += 256) skeys[i] = ~0ull;         // sorts last
    __syncthreads();
    for (int c = warp; c < ncand; c += 8) {
        const int g = cand[(size_t)b * K + c / kGroup];
        if (g < 0) continue;
        const int id = g * kGroup + (c % kGroup) + 1;                     // 1-based item id == table row
        if (id >= num_items) continue;
        const float* vr = w_item + (size_t)id * D;
        float a = 0.f;
        for (int col = lane * 4; col < D; col += 128) {
            const float4 x = ldg128(qr + col), y = ldg128(vr + col);
            a += EUCLID ? sqdist4(x, y) : dot4(x, y);
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) a += __shfl_xor_sync(kFull, a, o);
        float s = EUCLID ? -a : a;
        bool seen = false;
        for (int h = lane; h < H; h += 32) seen |= (hr[h] == (int64_t)id);
        if (__any_sync(kFull, seen)) s = -INFINITY;                       // baseretriever.py:390
        if (lane == 0) skeys[c] = ((unsigned long long)(~fkey(s)) << 32) | (uint32_t)id;   // score desc, id asc
    }
    __syncthreads();
    // bitonic sort ascending
    for (int size = 2; size <= ncand_pow2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < ncand_pow2 / 2; i += 256) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const unsigned long long x = skeys[lo], y = skeys[hi];
                if ((x > y) == up) { skeys[lo] = y; skeys[hi] = x; }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < k; i += 256) {
        const unsigned long long key = (i < ncand_pow2) ? skeys[i] : ~0ull;
        float s = -INFINITY; int64_t id = 0;
        if (key != ~0ull) {
            uint32_t u = ~(uint32_t)(key >> 32);                          // back to the ordered key
            u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;               // inverse of fkey
            s = __uint_as_float(u);
            id = (int64_t)(uint32_t)key;
        }
        score_out[(size_t)b * k + i] = s;
        id_out[(size_t)b * k + i] = id;
    }
}

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

}  // namespace rsb

using namespace rsb;

static int64_t topk_K(int64_t k, int64_t H, int64_t ngroups) {
    int64_t K = k + H + 8;
    return K < ngroups ? K : ngroups;
}

extern "C" size_t rsb200_topk_workspace_bytes(int64_t Be, int64_t num_items, int64_t k, int64_t H) {
    if (Be <= 0 || num_items < 2 || k <= 0 || H < 0) return 0;
    const int64_t ngroups = cdiv(num_items - 1, kGroup);
    const int64_t K = topk_K(k, H, ngroups);
    size_t a = (size_t)Be * ngroups * sizeof(float);
    a = (a + 255) & ~(size_t)255;
    return a + (size_t)Be * K * sizeof(int32_t);
}

extern "C" int32_t rsb200_topk_full(int32_t score_kind, const float* q, const float* w_item, int64_t num_items,
                                    int64_t d, int64_t Be, int64_t k, const int64_t* hist, int64_t H,
                                    float* score_out, int64_t* id_out, void* workspace, size_t workspace_bytes,
                                    void* stream) {
    RSB_REQUIRE(q && w_item && score_out && id_out && workspace, RSB200_EINVAL, "null pointer");
    RSB_REQUIRE(aligned16(q) && aligned16(w_item) && d >= 4 && d % 4 == 0, RSB200_EINVAL, "rows must be 16-byte aligned, d %% 4 == 0");
    RSB_REQUIRE(score_kind == RSB200_SCORE_IP || score_kind == RSB200_SCORE_EUCLID, RSB200_EINVAL, "bad score_kind");
    RSB_REQUIRE(Be >= 1 && k >= 1 && H >= 0 && num_items >= 2 && num_items < ((int64_t)1 << 31), RSB200_EINVAL, "bad shape");
    RSB_REQUIRE(hist || H == 0, RSB200_EINVAL, "H > 0 needs a history pointer");
    RSB_REQUIRE(k <= num_items - 1, RSB200_EINVAL, "k = %lld exceeds the %lld items", (long long)k, (long long)(num_items - 1));
    const int64_t Nit = num_items - 1, ngroups = cdiv(Nit, kGroup), K = topk_K(k, H, ngroups);
    RSB_REQUIRE(K * kGroup <= 8192, RSB200_EUNSUPPORTED, "k + H = %lld too large for the in-shared-memory final sort", (long long)(k + H));
    RSB_REQUIRE(workspace_bytes >= rsb200_topk_workspace_bytes(Be, num_items, k, H), RSB200_EWORKSPACE, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    float* gmax = (float*)workspace;
    size_t a = ((size_t)Be * ngroups * sizeof(float) + 255) & ~(size_t)255;
    int32_t* cand = (int32_t*)((char*)workspace + a);
    dim3 grid((unsigned)cdiv(Nit, BN), (unsigned)cdiv(Be, BM));
    const float* w1 = w_item + d;                                   // item id 1 (row 0 is padding)
    if (score_kind == RSB200_SCORE_IP) score_gmax_kernel<false><<<grid, 256, 0, st>>>(q, w1, (int)Be, (int)Nit, (int)d, gmax, (int)ngroups);
    else score_gmax_kernel<true><<<grid, 256, 0, st>>>(q, w1, (int)Be, (int)Nit, (int)d, gmax, (int)ngroups);
    RSB_LAUNCH_CHECK();
    select_groups_kernel<<<(unsigned)Be, 1024, 0, st>>>(gmax, (int)ngroups, (int)K, cand);
    RSB_LAUNCH_CHECK();
    const int np2 = next_pow2((int)(K * kGroup));
    const size_t smem = (size_t)np2 * sizeof(unsigned long long);
    if (score_kind == RSB200_SCORE_IP) {
        RSB_CUDA(cudaFuncSetAttribute(topk_final_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        topk_final_kernel<false><<<(unsigned)Be, 256, smem, st>>>(q, w_item, (int)num_items, (int)d, cand, (int)K, hist, (int)H, (int)k, score_out, id_out, np2);
    } else {
        RSB_CUDA(cudaFuncSetAttribute(topk_final_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        topk_final_kernel<true><<<(unsigned)Be, 256, smem, st>>>(q, w_item, (int)num_items, (int)d, cand, (int)K, hist, (int)H, (int)k, score_out, id_out, np2);
    }
    RSB_LAUNCH_CHECK();
    return 0;
}
