// ops.cu -- standalone plugin kernels (E1, E2, Q1/Q2 on ids, L1/L2 on score tensors).
// They exist so that each replacement plugin is correct on its own when it is mixed with
// reference plugins; the fused path (pair_fwd.cu + scatter.cu) is what a recognised
// plugin combination actually runs.
#include "common.cuh"
#include "kernels.h"

namespace rsb {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// E1: out[i, :] = w[ids[i], :]      (F.embedding, baseretriever.py:154,168,211)
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ w, int64_t num_rows, int D, const int64_t* __restrict__ ids,
                   int64_t numel, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= numel) return;
    int64_t id = ids[i];
    if (id < 0 || id >= num_rows) id = 0;
    const float* src = w + (size_t)id * D;
    float* dst = out + (size_t)i * D;
    for (int c = lane * 4; c < D; c += 128) stg128_stream(dst + c, ldg128(src + c));
}

// E2 (dense sink, atomics): dw[ids[i], :] += d_out[i, :], row 0 skipped (padding_idx = 0)
__global__ void __launch_bounds__(256)
scatter_add_rows_kernel(float* __restrict__ dw, int64_t num_rows, int D, const int64_t* __restrict__ ids,
                        int64_t numel, const float* __restrict__ d_out) {
    const int lane = threadIdx.x & 31;
    int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= numel) return;
    int64_t id = ids[i];
    if (id <= 0 || id >= num_rows) return;
    const float* src = d_out + (size_t)i * D;
    float* dst = dw + (size_t)id * D;
    for (int c = lane * 4; c < D; c += 128) {
        float4 v = ldg128_stream(src + c);
        atomicAdd(reinterpret_cast<float4*>(dst + c), v);      // red.global.add.v4.f32 (sm_90+)
    }
}

// Q1/Q2 on ids: out[b, j] = score(q[b], w[ids[b, j]])
template <int SCORE>
__global__ void __launch_bounds__(256)
score_ids_kernel(const float* __restrict__ q, const float* __restrict__ w, int64_t num_rows, int D,
                 const int64_t* __restrict__ ids, int64_t B, int64_t n, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= B * n) return;
    int64_t b = i / n;
    int64_t id = ids[i];
    if (id < 0 || id >= num_rows) id = 0;
    const float* qr = q + (size_t)b * D;
    const float* vr = w + (size_t)id * D;
    float a = 0.f;
    for (int c = lane * 4; c < D; c += 128) {
        float4 x = ldg128(qr + c), y = ldg128_stream(vr + c);
        a += (SCORE == RSB200_SCORE_IP) ? dot4(x, y) : sqdist4(x, y);
    }
    a = warp_sum_f(a);
    if (lane == 0) out[i] = (SCORE == RSB200_SCORE_IP) ? a : -a;
}

// Q1/Q2 on ids, d <= 128, streaming form (the pool scoring of sampling_method dns / sir, baseretriever.py:
// 323-324): a warp takes 32 ids of one query, keeps 8 row requests (8 x 512 B) in flight, and reduces the 8
// partial dot products with a transposed butterfly (lane L ends up with row L / 4 of the group).
template <int SCORE>
__global__ void __launch_bounds__(256)
score_ids_stream_kernel(const float* __restrict__ q, const float* __restrict__ w, int64_t num_rows, int D,
                        const int64_t* __restrict__ ids, int64_t B, int64_t n, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t chunks = (n + 31) / 32;
    const int64_t wid = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (wid >= B * chunks) return;
    const int64_t b = wid / chunks;
    const int64_t jb = (wid - b * chunks) * 32;
    const int cnt = (int)min((int64_t)32, n - jb);
    const bool act = lane * 4 < D;
    const float4 qv = act ? ldg128(q + (size_t)b * D + lane * 4) : make_float4(0, 0, 0, 0);
    int64_t id = (lane < cnt) ? ids[b * n + jb + lane] : 0;
    if (id < 0 || id >= num_rows) id = 0;
    float mine = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        if (g * 8 < cnt) {                                  // warp-uniform
            float4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int64_t rid = __shfl_sync(kFull, id, g * 8 + k);
                v[k] = act ? ldg128_stream(w + (size_t)rid * D + lane * 4) : make_float4(0, 0, 0, 0);
            }
            float pr[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) pr[k] = (SCORE == RSB200_SCORE_IP) ? dot4(qv, v[k]) : sqdist4(qv, v[k]);
            // transposed butterfly: 8 partials per lane -> lane L holds the full sum of row L / 4
#pragma unroll
            for (int half = 4, off = 16; half >= 1; half >>= 1, off >>= 1) {
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (i < half) {
                        const float send = up ? pr[i] : pr[i + half];
                        const float keep = up ? pr[i + half] : pr[i];
                        pr[i] = keep + __shfl_xor_sync(kFull, send, off);
                    }
                }
            }
            float s = pr[0];
            s += __shfl_xor_sync(kFull, s, 2);
            s += __shfl_xor_sync(kFull, s, 1);
            const float ts = __shfl_sync(kFull, s, (lane & 7) * 4);
            if ((lane >> 3) == g) mine = ts;
        }
    }
    if (lane < cnt) out[b * n + jb + lane] = (SCORE == RSB200_SCORE_IP) ? mine : -mine;
}

// L1/L2 on score tensors: one warp per query; loss and d loss/d score in one pass.
template <int LOSS>
__global__ void __launch_bounds__(256)
pair_loss_kernel(const float* __restrict__ pos, const float* __restrict__ neg, const float* __restrict__ lqp,
                 const float* __restrict__ lqn, int64_t B, int64_t n, float* __restrict__ d_pos,
                 float* __restrict__ d_neg, float* __restrict__ loss_part) {
    const int lane = threadIdx.x & 31;
    int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= B) return;
    const float sp = pos[b];
    const float* nr = neg + (size_t)b * n;
    float* dn = d_neg ? d_neg + (size_t)b * n : nullptr;
    if (LOSS == RSB200_LOSS_BPR) {
        const float inv = 1.0f / ((float)B * (float)(n > 0 ? n : 1));
        float ls = 0.f, cs = 0.f;
        for (int64_t j = lane; j < n; j += 32) {
            float x = nr[j] - sp;
            float c = sigmoidf(x) * inv;
            ls += softplusf(x);
            cs += c;
            if (dn) dn[j] = c;
        }
        ls = warp_sum_f(ls); cs = warp_sum_f(cs);
        if (lane == 0) { loss_part[b] = ls * inv; if (d_pos) d_pos[b] = -cs; }
    } else if (LOSS == 2) {
        // SoftmaxLoss on a materialised all_score row (loss_func.py:41-42): lse(all) - pos
        const float invB = 1.0f / (float)B;
        float m = -INFINITY;
        for (int64_t j = lane; j < n; j += 32) m = fmaxf(m, nr[j]);
        m = warp_max_f(m);
        float l = 0.f;
        for (int64_t j = lane; j < n; j += 32) l += expf(nr[j] - m);
        l = warp_sum_f(l);
        const float lse = m + logf(l);
        if (dn)
            for (int64_t j = lane; j < n; j += 32) dn[j] = expf(nr[j] - lse) * invB;
        if (lane == 0) { loss_part[b] = (lse - sp) * invB; if (d_pos) d_pos[b] = -invB; }
    } else {
        const float invB = 1.0f / (float)B;
        const float z0 = sp - (lqp ? lqp[b] : 0.f);
        float m = z0;
        for (int64_t j = lane; j < n; j += 32) m = fmaxf(m, nr[j] - (lqn ? lqn[(size_t)b * n + j] : 0.f));
        m = warp_max_f(m);
        float l = 0.f;
        for (int64_t j = lane; j < n; j += 32) l += expf(nr[j] - (lqn ? lqn[(size_t)b * n + j] : 0.f) - m);
        l = warp_sum_f(l) + expf(z0 - m);
        const float lse = m + logf(l);
        if (dn)
            for (int64_t j = lane; j < n; j += 32)
                dn[j] = expf(nr[j] - (lqn ? lqn[(size_t)b * n + j] : 0.f) - lse) * invB;
        if (lane == 0) { loss_part[b] = (lse - z0) * invB; if (d_pos) d_pos[b] = (expf(z0 - lse) - 1.f) * invB; }
    }
}

}  // namespace rsb

using namespace rsb;

static int32_t check_rows(const void* w, int64_t num_rows, int64_t d) {
    RSB_REQUIRE(w != nullptr && aligned16(w), RSB200_EINVAL, "table pointer null or not 16-byte aligned");
    RSB_REQUIRE(num_rows >= 1 && d >= 4 && d % 4 == 0, RSB200_EINVAL, "bad table shape [%lld, %lld]", (long long)num_rows, (long long)d);
    return 0;
}

extern "C" int32_t rsb200_gather_rows(const float* w, int64_t num_rows, int64_t d, const int64_t* ids, int64_t numel,
                                      float* out, void* stream) {
    int32_t rc = check_rows(w, num_rows, d);
    if (rc) return rc;
    RSB_REQUIRE(ids && out && aligned16(out), RSB200_EINVAL, "null / misaligned pointer");
    if (numel == 0) return 0;
    gather_rows_kernel<<<(unsigned)cdiv(numel, 8), 256, 0, (cudaStream_t)stream>>>(w, num_rows, (int)d, ids, numel, out);
    RSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int32_t rsb200_scatter_add_rows(float* dw, int64_t num_rows, int64_t d, const int64_t* ids, int64_t numel,
                                           const float* d_out, void* stream) {
    int32_t rc = check_rows(dw, num_rows, d);
    if (rc) return rc;
    RSB_REQUIRE(ids && d_out && aligned16(d_out), RSB200_EINVAL, "null / misaligned pointer");
    if (numel == 0) return 0;
    scatter_add_rows_kernel<<<(unsigned)cdiv(numel, 8), 256, 0, (cudaStream_t)stream>>>(dw, num_rows, (int)d, ids, numel, d_out);
    RSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int32_t rsb200_score_ids(int32_t score_kind, const float* q, const float* w, int64_t num_rows, int64_t d,
                                    const int64_t* ids, int64_t B, int64_t n, float* out, void* stream) {
    int32_t rc = check_rows(w, num_rows, d);
    if (rc) return rc;
    RSB_REQUIRE(q && ids && out && aligned16(q), RSB200_EINVAL, "null / misaligned pointer");
    RSB_REQUIRE(score_kind == RSB200_SCORE_IP || score_kind == RSB200_SCORE_EUCLID, RSB200_EINVAL, "bad score_kind");
    if (B * n == 0) return 0;
    if (d <= 128 && n >= 8) {
        unsigned sgrid = (unsigned)cdiv(B * cdiv(n, 32), 8);
        if (score_kind == RSB200_SCORE_IP)
            score_ids_stream_kernel<RSB200_SCORE_IP><<<sgrid, 256, 0, (cudaStream_t)stream>>>(q, w, num_rows, (int)d, ids, B, n, out);
        else
            score_ids_stream_kernel<RSB200_SCORE_EUCLID><<<sgrid, 256, 0, (cudaStream_t)stream>>>(q, w, num_rows, (int)d, ids, B, n, out);
        RSB_LAUNCH_CHECK();
        return 0;
    }
    unsigned grid = (unsigned)cdiv(B * n, 8);
    if (score_kind == RSB200_SCORE_IP)
        score_ids_kernel<RSB200_SCORE_IP><<<grid, 256, 0, (cudaStream_t)stream>>>(q, w, num_rows, (int)d, ids, B, n, out);
    else
        score_ids_kernel<RSB200_SCORE_EUCLID><<<grid, 256, 0, (cudaStream_t)stream>>>(q, w, num_rows, (int)d, ids, B, n, out);
    RSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int32_t rsb200_pair_loss(int32_t loss_kind, const float* pos_score, const float* neg_score,
                                    const float* logq_pos, const float* logq_neg, int64_t B, int64_t n, float* loss,
                                    float* d_pos, float* d_neg, float* loss_part, void* stream) {
    RSB_REQUIRE(pos_score && (neg_score || n == 0) && loss && loss_part, RSB200_EINVAL, "null pointer");
    RSB_REQUIRE(loss_kind == RSB200_LOSS_BPR || loss_kind == RSB200_LOSS_SSM || loss_kind == RSB200_LOSS_FULL, RSB200_EINVAL, "bad loss_kind");
    RSB_REQUIRE(B >= 1 && n >= 0 && B < ((int64_t)1 << 31), RSB200_EINVAL, "bad shape");
    unsigned grid = (unsigned)cdiv(B, 8);
    if (loss_kind == RSB200_LOSS_BPR)
        pair_loss_kernel<RSB200_LOSS_BPR><<<grid, 256, 0, (cudaStream_t)stream>>>(pos_score, neg_score, logq_pos, logq_neg, B, n, d_pos, d_neg, loss_part);
    else if (loss_kind == RSB200_LOSS_FULL)
        pair_loss_kernel<RSB200_LOSS_FULL><<<grid, 256, 0, (cudaStream_t)stream>>>(pos_score, neg_score, logq_pos, logq_neg, B, n, d_pos, d_neg, loss_part);
    else
        pair_loss_kernel<RSB200_LOSS_SSM><<<grid, 256, 0, (cudaStream_t)stream>>>(pos_score, neg_score, logq_pos, logq_neg, B, n, d_pos, d_neg, loss_part);
    RSB_LAUNCH_CHECK();
    return launch_loss_sum(loss_part, (int)B, loss, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------
// Q1/Q2 on an already gathered [B, n, d] tensor (standalone scorer plugins) + backward.
namespace rsb {

template <int SCORE>
__global__ void __launch_bounds__(256)
score_dense_kernel(const float* __restrict__ q, const float* __restrict__ items, int64_t B, int64_t n, int D,
                   float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= B * n) return;
    const float* qr = q + (size_t)(i / n) * D;
    const float* vr = items + (size_t)i * D;
    float a = 0.f;
    for (int c = lane * 4; c < D; c += 128) {
        float4 x = ldg128(qr + c), y = ldg128_stream(vr + c);
        a += (SCORE == RSB200_SCORE_IP) ? dot4(x, y) : sqdist4(x, y);
    }
    a = warp_sum_f(a);
    if (lane == 0) out[i] = (SCORE == RSB200_SCORE_IP) ? a : -a;
}

// one warp per query: ditems[b,j,:] = g q (IP) | 2 g (q - v) (EU); dq[b,:] = sum_j g v | 2 g (v - q)
template <int SCORE>
__global__ void __launch_bounds__(256)
score_dense_bwd_kernel(const float* __restrict__ q, const float* __restrict__ items, const float* __restrict__ g,
                       int64_t B, int64_t n, int D, float* __restrict__ dq, float* __restrict__ ditems) {
    const int lane = threadIdx.x & 31;
    int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= B) return;
    for (int c = lane * 4; c < D; c += 128) {
        float4 x = ldg128(q + (size_t)b * D + c);
        float4 acc = make_float4(0, 0, 0, 0);
        for (int64_t j = 0; j < n; ++j) {
            const float gj = g[(size_t)b * n + j];
            const size_t o = ((size_t)b * n + j) * D + c;
            float4 v = ldg128_stream(items + o);
            float4 dv;
            if (SCORE == RSB200_SCORE_IP) {
                dv = make_float4(gj * x.x, gj * x.y, gj * x.z, gj * x.w);
                fma4(acc, gj, v);
            } else {
                float4 df = make_float4(x.x - v.x, x.y - v.y, x.z - v.z, x.w - v.w);
                dv = make_float4(2.f * gj * df.x, 2.f * gj * df.y, 2.f * gj * df.z, 2.f * gj * df.w);
                fma4(acc, -2.f * gj, df);
            }
            if (ditems) stg128_stream(ditems + o, dv);
        }
        if (dq) *reinterpret_cast<float4*>(dq + (size_t)b * D + c) = acc;
    }
}

}  // namespace rsb

extern "C" int32_t rsb200_score_dense(int32_t score_kind, const float* q, const float* items, int64_t B, int64_t n,
                                      int64_t d, float* out, void* stream) {
    RSB_REQUIRE(q && items && out && aligned16(q) && aligned16(items), RSB200_EINVAL, "null / misaligned pointer");
    RSB_REQUIRE(d >= 4 && d % 4 == 0 && B >= 0 && n >= 0, RSB200_EINVAL, "bad shape");
    RSB_REQUIRE(score_kind == RSB200_SCORE_IP || score_kind == RSB200_SCORE_EUCLID, RSB200_EINVAL, "bad score_kind");
    if (B * n == 0) return 0;
    unsigned grid = (unsigned)cdiv(B * n, 8);
    if (score_kind == RSB200_SCORE_IP) score_dense_kernel<RSB200_SCORE_IP><<<grid, 256, 0, (cudaStream_t)stream>>>(q, items, B, n, (int)d, out);
    else score_dense_kernel<RSB200_SCORE_EUCLID><<<grid, 256, 0, (cudaStream_t)stream>>>(q, items, B, n, (int)d, out);
    RSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int32_t rsb200_score_dense_bwd(int32_t score_kind, const float* q, const float* items, const float* g,
                                          int64_t B, int64_t n, int64_t d, float* dq, float* ditems, void* stream) {
    RSB_REQUIRE(q && items && g && aligned16(q) && aligned16(items) && aligned16(dq) && aligned16(ditems), RSB200_EINVAL,
                "null / misaligned pointer");
    RSB_REQUIRE(d >= 4 && d % 4 == 0 && B >= 0 && n >= 0, RSB200_EINVAL, "bad shape");
    RSB_REQUIRE(score_kind == RSB200_SCORE_IP || score_kind == RSB200_SCORE_EUCLID, RSB200_EINVAL, "bad score_kind");
    if (B == 0) return 0;
    unsigned grid = (unsigned)cdiv(B, 8);
    if (score_kind == RSB200_SCORE_IP) score_dense_bwd_kernel<RSB200_SCORE_IP><<<grid, 256, 0, (cudaStream_t)stream>>>(q, items, g, B, n, (int)d, dq, ditems);
    else score_dense_bwd_kernel<RSB200_SCORE_EUCLID><<<grid, 256, 0, (cudaStream_t)stream>>>(q, items, g, B, n, (int)d, dq, ditems);
    RSB_LAUNCH_CHECK();
    return 0;
}
