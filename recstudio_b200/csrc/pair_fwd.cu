// pair_fwd.cu -- the fused gather -> score -> loss -> gradient-coefficient kernel (PHASE_FWD).
//
// Replaces, for B interactions (user, pos, neg[n]):
//   F.embedding x3               recstudio/model/basemodel/baseretriever.py:154,168,211
//   score_func(query, pos|neg)   recstudio/model/scorer.py:10-14 (IP), :28-34 (Euclid)
//   BPRLoss / SampledSoftmaxLoss recstudio/model/loss_func.py:55-59 / :80-90
//   and the part of loss.backward() (recommender.py:638) that turns the loss into
//   per-touch gradient coefficients and the query gradient.
//
// Because the loss is a mean with a constant upstream gradient, forward and backward
// collapse into ONE pass over the gathered rows:
//   BPR : c_bj = sigmoid(s-_bj - s+_b) / (B n),  c+_b = -sum_j c_bj
//   SSM : z = s - logQ, c_bj = softmax(z_b)_j / B, c+_b = (softmax(z_b)_0 - 1) / B
//   IP  : dq_b = sum_j c_bj v_bj + c+_b v+_b ;   dV[row] += c * q_b
//   EU  : dq_b = 2 (sum_j c_bj (v_bj - q_b) + c+_b (v+_b - q_b)) ; dV[row] += 2 c (q_b - v_row)
// The [B,n,d] gathered tensor, the [B,n] score/probability temporaries and the dense
// [N,d] gradient of the reference never exist.  Per touched row the kernel emits one
// 8-byte entry (query index, coefficient | logit) into the row-grouped list built by
// group.cu; scatter.cu turns the list into gradient rows.
//
// Mapping: one query per CTA (8 warps x n/8 negatives) when n >= 128, else one query
// per warp.  A warp streams its negatives in batches of 32 ids; rows are fetched LOADS
// at a time with one 16-byte load per lane (a 512-B row at d = 128 is one fully
// coalesced warp request), partial dot products of LOADS rows are reduced with a
// transposed butterfly (LOADS + log2(32) - 1 shuffles instead of 5 per row), and the
// online-softmax state (m, l, acc) of SSM is kept per warp and merged in shared memory.
#include "common.cuh"
#include "kernels.h"

namespace rsb {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

__device__ __forceinline__ uint64_t pack_entry(uint32_t b_flag, float v) {
    return (uint64_t)b_flag | ((uint64_t)__float_as_uint(v) << 32);
}

template <int K, int OFF>
struct TransposeReduce {
    // K partial sums per lane -> full sums; lane L ends up owning element L / (32 / K0)
    __device__ __forceinline__ static float run(float* p, int lane) {
        constexpr int half = K / 2;
        const bool up = (lane & OFF) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            float send = up ? p[i] : p[i + half];
            float keep = up ? p[i + half] : p[i];
            p[i] = keep + __shfl_xor_sync(kFull, send, OFF);
        }
        return TransposeReduce<half, OFF / 2>::run(p, lane);
    }
};
template <int OFF>
struct TransposeReduce<1, OFF> {
    __device__ __forceinline__ static float run(float* p, int) {
        float r = p[0];
#pragma unroll
        for (int o = OFF; o >= 1; o >>= 1) r += __shfl_xor_sync(kFull, r, o);
        return r;
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// VPL = float4 per lane per row (D <= 128*VPL); lanes whose columns are >= D are idle.
template <int VPL, int LOSS, int SCORE, bool MULTI>
__global__ void __launch_bounds__(kThreads, (VPL == 1) ? 3 : 2)
pair_fwd_kernel(const FwdParams p) {
    constexpr int LOADS = (VPL == 1) ? 8 : (VPL == 2 ? 4 : 2);   // rows fetched per group
    constexpr int REP = 32 / LOADS;                               // lanes sharing one reduced row
    constexpr int NG = 32 / LOADS;                                // groups per 32-id batch
    constexpr float kRepInv = 1.0f / REP;

    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = p.D;
    // query handled by this warp / CTA
    const int b = MULTI ? blockIdx.x : blockIdx.x * kWarps + warp;
    if (!MULTI && b >= p.B) return;     // whole warp exits; no CTA-wide barrier on this path

    // shared layout per query group: q[D] vp[D] acc[nw][D] stats[nw][4]
    const int nw = MULTI ? kWarps : 1;
    float* sm_base = MULTI ? smem : smem + (size_t)warp * (size_t)(3 * D + 4);
    float* s_q = sm_base;
    float* s_vp = s_q + D;
    float* s_acc = s_vp + D;                       // [nw][D]
    float* s_stat = s_acc + (size_t)nw * D;        // [nw][4] = {m, l, csum, loss}

    int64_t uid = p.user[b], pid = p.pos[b];
    if (uid < 0 || uid >= p.num_users) uid = 0;
    if (pid < 0 || pid >= p.num_items) pid = 0;

    float4 q[VPL], vp[VPL];
    bool act[VPL];
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
        int col = lane * 4 + t * 128;
        act[t] = col < D;
        q[t] = act[t] ? ldg128(p.w_user + (size_t)uid * D + col) : make_float4(0, 0, 0, 0);
        vp[t] = act[t] ? ldg128(p.w_item + (size_t)pid * D + col) : make_float4(0, 0, 0, 0);
    }
    float sp = 0.f;
#pragma unroll
    for (int t = 0; t < VPL; ++t) sp += (SCORE == RSB200_SCORE_IP) ? dot4(q[t], vp[t]) : sqdist4(q[t], vp[t]);
    sp = warp_sum(sp);
    if (SCORE == RSB200_SCORE_EUCLID) sp = -sp;

    if (!MULTI || warp == 0) {
#pragma unroll
        for (int t = 0; t < VPL; ++t) {
            int col = lane * 4 + t * 128;
            if (act[t]) {
                *reinterpret_cast<float4*>(s_q + col) = q[t];
                *reinterpret_cast<float4*>(s_vp + col) = vp[t];
                *reinterpret_cast<float4*>(p.q_buf + (size_t)b * D + col) = q[t];
            }
        }
    }

    // this warp's slice of the negatives
    const int n = p.n;
    int j0 = 0, j1 = n;
    if (MULTI) {
        int per = ((n + kWarps - 1) / kWarps + 31) & ~31;   // multiple of 32 so batches stay aligned
        j0 = min(n, warp * per);
        j1 = min(n, j0 + per);
    }
    const size_t rowbase = (size_t)b * n;

    float4 acc[VPL];
#pragma unroll
    for (int t = 0; t < VPL; ++t) acc[t] = make_float4(0, 0, 0, 0);
    float csum = 0.f, lossacc = 0.f;          // per-lane partials (each element counted REP times)
    float m_run = -INFINITY, l_run = 0.f;     // SSM online softmax (warp-uniform)

    for (int jb = j0; jb < j1; jb += 32) {
        const int j = jb + lane;
        const bool valid = j < j1;
        int id = valid ? p.neg[rowbase + j] : 0;
        if ((unsigned)id >= (unsigned)p.num_items) id = 0;
        const uint32_t slot = valid ? p.slot_neg[rowbase + j] : kNoSlot;
        float lq = 0.f;
        if (LOSS == RSB200_LOSS_SSM && p.logq_neg != nullptr && valid) lq = p.logq_neg[rowbase + j];
        uint32_t epos = 0;
        if (slot != kNoSlot) epos = __ldg(p.off_item + id) + slot;
        float val_out = 0.f, sc_out = 0.f;

#pragma unroll
        for (int g = 0; g < NG; ++g) {
            if (jb + g * LOADS < j1) {            // warp-uniform
                float4 v[LOADS][VPL];
#pragma unroll
                for (int k = 0; k < LOADS; ++k) {
                    int rid = __shfl_sync(kFull, id, g * LOADS + k);
                    const float* row = p.w_item + (size_t)rid * D + lane * 4;
#pragma unroll
                    for (int t = 0; t < VPL; ++t)
                        v[k][t] = act[t] ? ldg128_stream(row + t * 128) : make_float4(0, 0, 0, 0);
                }
                float pr[LOADS];
#pragma unroll
                for (int k = 0; k < LOADS; ++k) {
                    float a = 0.f;
#pragma unroll
                    for (int t = 0; t < VPL; ++t)
                        a += (SCORE == RSB200_SCORE_IP) ? dot4(q[t], v[k][t]) : sqdist4(q[t], v[k][t]);
                    pr[k] = a;
                }
                float s = TransposeReduce<LOADS, 16>::run(pr, lane);
                if (SCORE == RSB200_SCORE_EUCLID) s = -s;
                const int e_own = g * LOADS + lane / REP;       // element (within the batch) this lane owns
                const bool ev = (jb + e_own) < j1;
                float wgt;                                        // weight of v in the dq accumulator
                float value;                                      // what goes into the entry
                if (LOSS == RSB200_LOSS_BPR) {
                    float x = s - sp;
                    wgt = ev ? sigmoidf(x) * p.coef_scale : 0.f;
                    lossacc += ev ? softplusf(x) : 0.f;
                    csum += wgt;
                    value = wgt;
                } else {
                    float lq_e = __shfl_sync(kFull, lq, e_own);
                    float z = ev ? (s - lq_e) : -INFINITY;
                    float gm = warp_max(z);
                    float m_new = fmaxf(m_run, gm);
                    // m_new == -inf only if nothing valid has been seen: keep everything at zero
                    float scale = (m_new == -INFINITY) ? 1.f : expf(m_run - m_new);
                    wgt = (ev && m_new != -INFINITY) ? expf(z - m_new) : 0.f;
                    l_run = l_run * scale + warp_sum(wgt) * kRepInv;
#pragma unroll
                    for (int t = 0; t < VPL; ++t) {
                        acc[t].x *= scale; acc[t].y *= scale; acc[t].z *= scale; acc[t].w *= scale;
                    }
                    m_run = m_new;
                    value = z;
                }
#pragma unroll
                for (int k = 0; k < LOADS; ++k) {
                    float wk = __shfl_sync(kFull, wgt, k * REP);
#pragma unroll
                    for (int t = 0; t < VPL; ++t) fma4(acc[t], wk, v[k][t]);
                }
                // hand each element's value back to the lane that holds its id
                float tv = __shfl_sync(kFull, value, (lane % LOADS) * REP);
                float ts = __shfl_sync(kFull, s, (lane % LOADS) * REP);
                if (lane / LOADS == g) { val_out = tv; sc_out = ts; }
            }
        }
        if (valid) {
            if (p.neg_score) p.neg_score[rowbase + j] = sc_out;
            if (slot != kNoSlot)
                p.ent_item[epos] = pack_entry((uint32_t)b | (LOSS == RSB200_LOSS_BPR ? kDirect : 0u), val_out);
        }
    }

    // ---- per-warp -> per-query merge ------------------------------------------------------
    csum = warp_sum(csum) * kRepInv;
    lossacc = warp_sum(lossacc) * kRepInv;
    {
        float* a = s_acc + (size_t)(MULTI ? warp : 0) * D;
#pragma unroll
        for (int t = 0; t < VPL; ++t) {
            int col = lane * 4 + t * 128;
            if (act[t]) *reinterpret_cast<float4*>(a + col) = acc[t];
        }
        if (lane == 0) {
            float* st = s_stat + (MULTI ? warp : 0) * 4;
            st[0] = m_run; st[1] = l_run; st[2] = csum; st[3] = lossacc;
        }
    }
    if (MULTI) __syncthreads(); else __syncwarp();

    // every thread of the group recomputes the (tiny) scalar merge
    float M = -INFINITY, L = 0.f, CS = 0.f, LS = 0.f;
    for (int w = 0; w < nw; ++w) M = fmaxf(M, s_stat[w * 4 + 0]);
    float wscale[kWarps];
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        wscale[w] = 1.f;
        if (w < nw) {
            if (LOSS == RSB200_LOSS_SSM) {
                float mw = s_stat[w * 4 + 0];
                wscale[w] = (mw == -INFINITY) ? 0.f : expf(mw - M);
                L += s_stat[w * 4 + 1] * wscale[w];
            }
            CS += s_stat[w * 4 + 2];
            LS += s_stat[w * 4 + 3];
        }
    }
    float cpos, neg_mul, loss_b, lse_b = 0.f;
    if (LOSS == RSB200_LOSS_BPR) {
        cpos = -CS;
        neg_mul = 1.f;                       // acc already carries the final coefficients
        loss_b = LS * p.loss_scale;
    } else {
        float z0 = sp - (p.logq_pos ? p.logq_pos[b] : 0.f);
        float M2 = fmaxf(M, z0);
        float L2 = ((M == -INFINITY) ? 0.f : L * expf(M - M2)) + expf(z0 - M2);
        lse_b = M2 + logf(L2);
        float p0 = expf(z0 - lse_b);
        cpos = (p0 - 1.f) * p.coef_scale;
        neg_mul = (M == -INFINITY) ? 0.f : expf(M - lse_b) * p.coef_scale;   // acc, L are relative to M
        CS = L * neg_mul;                    // = sum_j c_bj
        loss_b = (lse_b - z0) * p.loss_scale;
    }

    const int gsize = MULTI ? kThreads : 32;
    const int gtid = MULTI ? threadIdx.x : lane;
    for (int c = gtid; c < D; c += gsize) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w)
            if (w < nw) a += s_acc[(size_t)w * D + c] * wscale[w];
        a *= neg_mul;
        float qc = s_q[c], vc = s_vp[c];
        float dq;
        if (SCORE == RSB200_SCORE_IP) dq = a + cpos * vc;
        else dq = 2.f * (a - CS * qc) + 2.f * cpos * (vc - qc);
        p.dq_buf[(size_t)b * D + c] = dq;
    }
    if (gtid == 0) {
        p.loss_part[b] = loss_b;
        if (LOSS == RSB200_LOSS_SSM) p.lse[b] = lse_b;
        if (p.pos_score) p.pos_score[b] = sp;
        uint32_t sl = p.slot_pos[b];
        if (sl != kNoSlot) p.ent_item[__ldg(p.off_item + pid) + sl] = pack_entry((uint32_t)b | kDirect, cpos);
        uint32_t su = p.slot_user[b];
        if (su != kNoSlot) p.ent_user[__ldg(p.off_user + uid) + su] = pack_entry((uint32_t)b | kDirect, 1.0f);
    }
}

template <int VPL, int LOSS, int SCORE>
static int32_t launch_fwd_vls(const FwdParams& p, cudaStream_t st) {
    const bool multi = p.n >= 256;
    if (multi) {
        size_t smem = (size_t)(2 * p.D + kWarps * p.D + kWarps * 4) * sizeof(float);
        pair_fwd_kernel<VPL, LOSS, SCORE, true><<<p.B, kThreads, smem, st>>>(p);
    } else {
        size_t smem = (size_t)kWarps * (3 * p.D + 4) * sizeof(float);
        pair_fwd_kernel<VPL, LOSS, SCORE, false><<<(unsigned)cdiv(p.B, kWarps), kThreads, smem, st>>>(p);
    }
    RSB_LAUNCH_CHECK();
    return 0;
}

template <int VPL>
static int32_t launch_fwd_v(const FwdParams& p, int loss, int score, cudaStream_t st) {
    if (loss == RSB200_LOSS_BPR) {
        return score == RSB200_SCORE_IP ? launch_fwd_vls<VPL, RSB200_LOSS_BPR, RSB200_SCORE_IP>(p, st)
                                        : launch_fwd_vls<VPL, RSB200_LOSS_BPR, RSB200_SCORE_EUCLID>(p, st);
    }
    return score == RSB200_SCORE_IP ? launch_fwd_vls<VPL, RSB200_LOSS_SSM, RSB200_SCORE_IP>(p, st)
                                    : launch_fwd_vls<VPL, RSB200_LOSS_SSM, RSB200_SCORE_EUCLID>(p, st);
}

int32_t launch_pair_fwd(const FwdParams& p, int loss, int score, int variant, cudaStream_t st) {
    (void)variant;
    if (p.B == 0) return 0;
    if (p.D <= 128) return launch_fwd_v<1>(p, loss, score, st);
    if (p.D <= 256) return launch_fwd_v<2>(p, loss, score, st);
    if (p.D <= 512) return launch_fwd_v<4>(p, loss, score, st);
    set_error("embedding dim %d > 512 is not supported", p.D);
    return RSB200_EUNSUPPORTED;
}

}  // namespace rsb
