// fullsoftmax.cu -- L3: SoftmaxLoss over the whole catalog, forward + backward.
//   baseretriever.py:177-186 (full-score branch: all_score = query @ weight[1:].T)
//   loss_func.py:39-42       (mean_b( logsumexp_i all_score_bi - pos_score_b ))
//   autograd: d/d all_score = softmax/B, d/d pos_score = -1/B; dQ = dS W, dW = dS^T Q
//
// The [B, N-1] score matrix (30 GB at B = 8192, N = 1M) is never written to HBM:
//   pass 1  fs_lse_kernel   : S tile = Q_mb W_tile^T (fp32 FFMA), per-(query, tile) (max, sum exp)
//           fs_finish_kernel: per query combine the tile partials -> lse_b, pos_score_b, loss part
//   pass 2  fs_bwd_kernel   : one CTA per item tile, looping over the query blocks: recompute S,
//           dS = (exp(S - lse) - onehot(pos)) / B into shared memory in both layouts, then
//             dW_tile += dS^T Q_mb   (accumulated in registers over all query blocks; every dW row
//                                     is written exactly once -- dense, no atomics)
//             dQ_mb   += dS   W_tile (vector atomics into the small [B, d] buffer)
// 4 GEMM-equivalents (reference: 3, plus two [B,N] round trips through HBM).  fp32 on the CUDA
// cores by design (north star keeps tensor cores for attention; SURVEY 8(d) C4 discussion).
// Limits of this version: d <= 128.
#include "common.cuh"
#include "kernels.h"
#include "tile_gemm.cuh"

namespace rsb {
using namespace tg;

// ---------------------------------------------------------------------------- pass 1
__global__ void __launch_bounds__(256, 2)
fs_lse_kernel(const float* __restrict__ q, const float* __restrict__ w1, int B, int Nit, int D, float2* __restrict__ part,
              int ntiles) {
    __shared__ __align__(16) float As[2][TK][LDS_];
    __shared__ __align__(16) float Bs[2][TK][LDS_];
    float2 (*red)[TM] = reinterpret_cast<float2 (*)[TM]>(&As[0][0][0]);   // reused after the GEMM (16 KB <= sizeof(As))
    static_assert(sizeof(float2) * 16 * TM <= sizeof(float) * 2 * TK * LDS_, "red must fit in As");
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int tile = blockIdx.x, m0 = blockIdx.y * TM, n0 = tile * TN;
    float acc[8][8];
    zero_acc(acc);
    gemm_nt(acc, q, B, m0, w1, Nit, n0, D, As, Bs);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (n0 + ty * 8 + j < Nit) mx = fmaxf(mx, acc[i][j]);
        float l = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (n0 + ty * 8 + j < Nit) l += expf(acc[i][j] - mx);
        red[ty][row_of(tx, i)] = make_float2(mx, l);
    }
    __syncthreads();
    if (tid < TM && m0 + tid < B) {
        float mx = -INFINITY;
#pragma unroll
        for (int t = 0; t < 16; ++t) mx = fmaxf(mx, red[t][tid].x);
        float l = 0.f;
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const float2 r = red[t][tid];
            if (r.x != -INFINITY) l += r.y * expf(r.x - mx);
        }
        part[(size_t)(m0 + tid) * ntiles + tile] = make_float2(mx, l);
    }
}

// one warp per query: lse, pos_score, loss part
__global__ void __launch_bounds__(256)
fs_finish_kernel(const float2* __restrict__ part, int ntiles, const float* __restrict__ q, const float* __restrict__ w,
                 const int64_t* __restrict__ pos, int B, int N, int D, float* __restrict__ lse, float* __restrict__ loss_part) {
    const int lane = threadIdx.x & 31, b = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= B) return;
    float mx = -INFINITY;
    for (int t = lane; t < ntiles; t += 32) mx = fmaxf(mx, part[(size_t)b * ntiles + t].x);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
    float l = 0.f;
    for (int t = lane; t < ntiles; t += 32) {
        const float2 r = part[(size_t)b * ntiles + t];
        if (r.x != -INFINITY) l += r.y * expf(r.x - mx);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) l += __shfl_xor_sync(kFull, l, o);
    int64_t p = pos[b];
    if (p < 0 || p >= N) p = 0;
    float ps = 0.f;
    for (int c = lane * 4; c < D; c += 128) ps += dot4(ldg128(q + (size_t)b * D + c), ldg128(w + (size_t)p * D + c));
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) ps += __shfl_xor_sync(kFull, ps, o);
    if (lane == 0) {
        const float ls = mx + logf(l);
        lse[b] = ls;
        loss_part[b] = (ls - ps) / (float)B;
    }
}

// ---------------------------------------------------------------------------- pass 2
__global__ void __launch_bounds__(256, 1)
fs_bwd_kernel(const float* __restrict__ q, const float* __restrict__ w, const int64_t* __restrict__ pos,
              const float* __restrict__ lse, int B, int N, int D, float* __restrict__ dq, float* __restrict__ dw) {
    extern __shared__ __align__(16) float sm[];
    float (*As)[TK][LDS_] = reinterpret_cast<float (*)[TK][LDS_]>(sm);
    float (*Bs)[TK][LDS_] = reinterpret_cast<float (*)[TK][LDS_]>(sm + 2 * TK * LDS_);
    float (*Bc)[TK][LDS_] = reinterpret_cast<float (*)[TK][LDS_]>(sm + 4 * TK * LDS_);
    float (*Ps)[LDS_] = reinterpret_cast<float (*)[LDS_]>(sm + 6 * TK * LDS_);               // dS[m][n]
    float (*Pt)[LDS_] = reinterpret_cast<float (*)[LDS_]>(sm + 6 * TK * LDS_ + TM * LDS_);   // dS[n][m]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int Nit = N - 1, n0 = blockIdx.x * TN;
    const float* w1 = w + D;                                   // item id 1
    const float invB = 1.0f / (float)B;
    float dwacc[8][8];
    zero_acc(dwacc);
    for (int m0 = 0; m0 < B; m0 += TM) {
        float s[8][8];
        zero_acc(s);
        gemm_nt(s, q, B, m0, w1, Nit, n0, D, As, Bs);
        // dS = (softmax - onehot) / B, written in both layouts
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int ml = row_of(tx, i), m = m0 + ml;
            const bool mv = m < B;
            const float ls = mv ? __ldg(lse + m) : 0.f;
            const int64_t pm = mv ? pos[m] : 0;
            float pv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = n0 + ty * 8 + j;
                float v = 0.f;
                if (mv && n < Nit) {
                    v = expf(s[i][j] - ls) * invB;
                    if (pm == (int64_t)n + 1) v -= invB;       // onehot(pos); pos = 0 (padding) never matches
                }
                pv[j] = v;
            }
            *reinterpret_cast<float4*>(&Ps[ml][ty * 8]) = make_float4(pv[0], pv[1], pv[2], pv[3]);
            *reinterpret_cast<float4*>(&Ps[ml][ty * 8 + 4]) = make_float4(pv[4], pv[5], pv[6], pv[7]);
#pragma unroll
            for (int j = 0; j < 8; ++j) Pt[ty * 8 + j][ml] = pv[j];
        }
        __syncthreads();
        // dW_tile[n][d] += sum_m dS[m][n] Q[m0+m][d]        (A = Ps: k = m, row = n)
        gemm_sa(dwacc, Ps, TM, q, B, m0, D, D, Bc);
        // dQ[m0+m][d] += sum_n dS[m][n] W1[n0+n][d]         (A = Pt: k = n, row = m)
        float t[8][8];
        zero_acc(t);
        gemm_sa(t, Pt, TN, w1, Nit, n0, D, D, Bc);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = m0 + row_of(tx, i);
            if (m < B && ty * 8 < D) {
                float* dst = dq + (size_t)m * D + ty * 8;
                atomicAdd(reinterpret_cast<float4*>(dst), make_float4(t[i][0], t[i][1], t[i][2], t[i][3]));
                if (ty * 8 + 4 < D) atomicAdd(reinterpret_cast<float4*>(dst + 4), make_float4(t[i][4], t[i][5], t[i][6], t[i][7]));
            }
        }
        __syncthreads();       // Ps / Pt are rewritten by the next query block
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = n0 + row_of(tx, i);
        if (n < Nit && ty * 8 < D) {
            float* dst = dw + (size_t)(n + 1) * D + ty * 8;
            stg128_stream(dst, make_float4(dwacc[i][0], dwacc[i][1], dwacc[i][2], dwacc[i][3]));
            if (ty * 8 + 4 < D) stg128_stream(dst + 4, make_float4(dwacc[i][4], dwacc[i][5], dwacc[i][6], dwacc[i][7]));
        }
    }
    if (blockIdx.x == 0)                                            // padding row 0 never receives gradient
        for (int c = tid; c < D; c += 256) dw[c] = 0.f;
}

constexpr size_t kBwdSmem = (size_t)(6 * TK * LDS_ + 2 * TM * LDS_) * sizeof(float);

}  // namespace rsb

using namespace rsb;

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" size_t rsb200_fullsoftmax_workspace_bytes(int64_t B, int64_t num_items, int64_t d) {
    if (B <= 0 || num_items < 2 || d <= 0) return 0;
    const int64_t ntiles = cdiv(num_items - 1, tg::TN);
    return align256((size_t)B * ntiles * sizeof(float2)) + align256((size_t)B * sizeof(float)) * 2;
}

extern "C" int32_t rsb200_fullsoftmax_fwd_bwd(const float* q, const float* w_item, const int64_t* pos, int64_t num_items,
                                              int64_t B, int64_t d, float* loss, float* dq, float* dw, void* workspace,
                                              size_t workspace_bytes, void* stream) {
    RSB_REQUIRE(q && w_item && pos && loss && dq && dw && workspace, RSB200_EINVAL, "null pointer");
    RSB_REQUIRE(aligned16(q) && aligned16(w_item) && aligned16(dq) && aligned16(dw), RSB200_EINVAL, "pointers must be 16-byte aligned");
    RSB_REQUIRE(d >= 4 && d % 4 == 0, RSB200_EINVAL, "d must be a positive multiple of 4");
    RSB_REQUIRE(d <= 128, RSB200_EUNSUPPORTED, "full-softmax kernels support d <= 128 (got %lld)", (long long)d);
    RSB_REQUIRE(B >= 1 && B < ((int64_t)1 << 24) && num_items >= 2 && num_items < ((int64_t)1 << 31), RSB200_EINVAL, "bad shape");
    RSB_REQUIRE(workspace_bytes >= rsb200_fullsoftmax_workspace_bytes(B, num_items, d), RSB200_EWORKSPACE, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int Nit = (int)(num_items - 1), ntiles = (int)cdiv(Nit, tg::TN), mblocks = (int)cdiv(B, tg::TM);
    RSB_REQUIRE(mblocks <= 65535, RSB200_EUNSUPPORTED, "B too large");
    char* wsp = (char*)workspace;
    float2* part = (float2*)wsp;
    float* lse = (float*)(wsp + align256((size_t)B * ntiles * sizeof(float2)));
    float* loss_part = (float*)((char*)lse + align256((size_t)B * sizeof(float)));
    const float* w1 = w_item + d;
    fs_lse_kernel<<<dim3(ntiles, mblocks), 256, 0, st>>>(q, w1, (int)B, Nit, (int)d, part, ntiles);
    RSB_LAUNCH_CHECK();
    fs_finish_kernel<<<(unsigned)cdiv(B, 8), 256, 0, st>>>(part, ntiles, q, w_item, pos, (int)B, (int)num_items, (int)d, lse, loss_part);
    RSB_LAUNCH_CHECK();
    int32_t rc = launch_loss_sum(loss_part, (int)B, loss, st);
    if (rc) return rc;
    RSB_CUDA(cudaMemsetAsync(dq, 0, sizeof(float) * (size_t)B * d, st));
    RSB_CUDA(cudaFuncSetAttribute(fs_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem));
    fs_bwd_kernel<<<ntiles, 256, kBwdSmem, st>>>(q, w_item, pos, lse, (int)B, (int)num_items, (int)d, dq, dw);
    RSB_LAUNCH_CHECK();
    return 0;
}
