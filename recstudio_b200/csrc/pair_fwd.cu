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
// Mapping: one query per CTA (8 warps x n/8 negatives) when n >= 256, else one query
// per warp.  A warp streams its negatives in batches of 32 ids; rows are fetched LOADS
// at a time with one 16-byte load per lane (a 512-B row at d = 128 is one fully
// coalesced warp request), partial dot products of LOADS rows are reduced with a
// transposed butterfly (LOADS + log2(32) - 1 shuffles instead of 5 per row), and the
// online-softmax state (m, l, acc) of SSM is kept per warp and merged in shared memory.
// PIPE = true software-pipelines the stream: the rows of group g+1 (and the ids of the
// next batch) are in flight while group g is reduced, so every warp keeps 8-16 row
// requests outstanding at all times (the kernel is bound by DRAM latency x parallelism).
#include "common.cuh"
#include "kernels.h"

namespace rsb {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

__device__ __forceinline__ uint64_t pack_entry(uint32_t b_flag, float v) {
    return (uint64_t)b_flag | ((uint64_t)__float_as_uint(v) << 32);
}

// binned grouping (bins.cu): reserve the next position of the row's bin list; the entry's low word carries the local row
__device__ __forceinline__ uint32_t bin_reserve(const FwdParams& p, int id) {
    return atomicAdd(p.bin_cursor + (size_t)((uint32_t)id >> p.bin_shift) * kCursorStride, 1u);
}
__device__ __forceinline__ uint32_t bin_low(const FwdParams& p, uint32_t b_flag, int id) {
    return b_flag | (((uint32_t)id & ((1u << p.bin_shift) - 1u)) << p.bin_bbits);
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

template <int VPL>
struct Cfg {
    static constexpr int LOADS = (VPL == 1) ? 8 : (VPL == 2 ? 4 : 2);   // rows fetched per group
    static constexpr int REP = 32 / LOADS;                               // lanes sharing one reduced row
    static constexpr int NG = 32 / LOADS;                                // groups per 32-id batch
};

// per-lane metadata of one batch of 32 negatives
struct BatchMeta {
    int id;           // row id (0 for lanes past the end)
    uint32_t slot;    // position inside the row's entry segment, kNoSlot = no gradient entry
    float lq;         // log Q(neg) (SSM)
};

template <int LOSS, bool HINT = false>
__device__ __forceinline__ BatchMeta load_meta(const FwdParams& p, size_t rowbase, int jb, int j1, int lane, uint64_t pol = 0) {
    BatchMeta m;
    const int j = jb + lane;
    const bool valid = j < j1;
    if (HINT) {        // ids / slots are read exactly once: same eviction priority as the rows
        m.id = valid ? (int)ldg32_hint(reinterpret_cast<const uint32_t*>(p.neg) + rowbase + j, pol) : 0;
        if ((unsigned)m.id >= (unsigned)p.num_items) m.id = 0;
        if (p.bin_cursor) m.slot = (valid && m.id != p.pad_row) ? 0u : kNoSlot;
        else m.slot = valid ? ldg32_hint(p.slot_neg + rowbase + j, pol) : kNoSlot;
    } else {
        m.id = valid ? __ldg(p.neg + rowbase + j) : 0;
        if ((unsigned)m.id >= (unsigned)p.num_items) m.id = 0;
        if (p.bin_cursor) m.slot = (valid && m.id != p.pad_row) ? 0u : kNoSlot;      // binned grouping: no per-touch slot array
        else m.slot = valid ? __ldg(p.slot_neg + rowbase + j) : kNoSlot;
    }
    m.lq = 0.f;
    if (LOSS == RSB200_LOSS_SSM && p.logq_neg != nullptr && valid) m.lq = __ldg(p.logq_neg + rowbase + j);
    return m;
}

template <int VPL>
struct RowBuf {
    float4 v[Cfg<VPL>::LOADS][VPL];
};

// fetch the LOADS rows of group g of a batch (ids live one per lane in `id`)
template <int VPL, bool HINT = false>
__device__ __forceinline__ void load_group(RowBuf<VPL>& buf, const FwdParams& p, int id, int g, int lane,
                                           const bool (&act)[VPL], uint64_t pol = 0) {
    constexpr int LOADS = Cfg<VPL>::LOADS;
#pragma unroll
    for (int k = 0; k < LOADS; ++k) {
        const int rid = __shfl_sync(kFull, id, g * LOADS + k);
        const float* row = p.w_item + (size_t)rid * p.D + lane * 4;
#pragma unroll
        for (int t = 0; t < VPL; ++t)
            buf.v[k][t] = act[t] ? (HINT ? ldg128_stream_hint(row + t * 128, pol) : ldg128_stream(row + t * 128))
                                 : make_float4(0, 0, 0, 0);
    }
}

template <int VPL>
struct WarpState {
    float4 acc[VPL];
    float csum, lossacc;     // per-lane partials (every element is counted REP times)
    float m_run, l_run;      // SSM online softmax (warp-uniform)
    float val_out, sc_out;   // per-lane value / score of the element whose id this lane holds
};

// score + loss + dq accumulation of one group whose rows are in `buf`
template <int VPL, int LOSS, int SCORE>
__device__ __forceinline__ void compute_group(const RowBuf<VPL>& buf, WarpState<VPL>& st, const FwdParams& p,
                                              const float4 (&q)[VPL], float sp, float lq, int g, int jb, int j1,
                                              int lane) {
    constexpr int LOADS = Cfg<VPL>::LOADS, REP = Cfg<VPL>::REP;
    constexpr float kRepInv = 1.0f / REP;
    float pr[LOADS];
#pragma unroll
    for (int k = 0; k < LOADS; ++k) {
        float a = 0.f;
#pragma unroll
        for (int t = 0; t < VPL; ++t)
            a += (SCORE == RSB200_SCORE_IP) ? dot4(q[t], buf.v[k][t]) : sqdist4(q[t], buf.v[k][t]);
        pr[k] = a;
    }
    float s = TransposeReduce<LOADS, 16>::run(pr, lane);
    if (SCORE == RSB200_SCORE_EUCLID) s = -s;
    const int e_own = g * LOADS + lane / REP;       // element (within the batch) this lane owns
    const bool ev = (jb + e_own) < j1;
    float wgt;                                        // weight of v in the dq accumulator
    float value;                                      // what goes into the entry
    if (LOSS == RSB200_LOSS_BPR) {
        const float x = s - sp;
        wgt = ev ? sigmoidf(x) * p.coef_scale : 0.f;
        st.lossacc += ev ? softplusf(x) : 0.f;
        st.csum += wgt;
        value = wgt;
    } else {
        const float lq_e = __shfl_sync(kFull, lq, e_own);
        const float z = ev ? (s - lq_e) : -INFINITY;
        const float gm = warp_max(z);
        const float m_new = fmaxf(st.m_run, gm);
        // m_new == -inf only if nothing valid has been seen: keep everything at zero
        const float scale = (m_new == -INFINITY) ? 1.f : expf(st.m_run - m_new);
        wgt = (ev && m_new != -INFINITY) ? expf(z - m_new) : 0.f;
        st.l_run = st.l_run * scale + warp_sum(wgt) * kRepInv;
#pragma unroll
        for (int t = 0; t < VPL; ++t) {
            st.acc[t].x *= scale; st.acc[t].y *= scale; st.acc[t].z *= scale; st.acc[t].w *= scale;
        }
        st.m_run = m_new;
        value = z;
    }
#pragma unroll
    for (int k = 0; k < LOADS; ++k) {
        const float wk = __shfl_sync(kFull, wgt, k * REP);
#pragma unroll
        for (int t = 0; t < VPL; ++t) fma4(st.acc[t], wk, buf.v[k][t]);
    }
    // hand each element's value back to the lane that holds its id
    const float tv = __shfl_sync(kFull, value, (lane % LOADS) * REP);
    const float ts = __shfl_sync(kFull, s, (lane % LOADS) * REP);
    if (lane / LOADS == g) { st.val_out = tv; st.sc_out = ts; }
}

// merge the per-warp partial state of one query and write dq, loss, lse and the positive / user entries
template <int VPL, int LOSS, int SCORE, bool MULTI>
__device__ __forceinline__ void finish_query(const FwdParams& p, WarpState<VPL>& st, float* s_q, float* s_vp,
                                             float* s_acc, float* s_stat, const bool (&act)[VPL], float sp, int b,
                                             int64_t uid, int64_t pid, int lane, int warp) {
    constexpr float kRepInv = 1.0f / Cfg<VPL>::REP;
    const int D = p.D;
    const int nw = MULTI ? kWarps : 1;
    // ---- per-warp -> per-query merge ------------------------------------------------------
    const float csum = warp_sum(st.csum) * kRepInv;
    const float lossacc = warp_sum(st.lossacc) * kRepInv;
    {
        float* a = s_acc + (size_t)(MULTI ? warp : 0) * D;
#pragma unroll
        for (int t = 0; t < VPL; ++t) {
            int col = lane * 4 + t * 128;
            if (act[t]) *reinterpret_cast<float4*>(a + col) = st.acc[t];
        }
        if (lane == 0) {
            float* sst = s_stat + (MULTI ? warp : 0) * 4;
            sst[0] = st.m_run; sst[1] = st.l_run; sst[2] = csum; sst[3] = lossacc;
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
        if (p.bin_cursor) {
            if (pid != 0) p.ent_item[bin_reserve(p, (int)pid)] = pack_entry(bin_low(p, (uint32_t)b | kDirect, (int)pid), cpos);
        } else {
            uint32_t sl = p.slot_pos[b];
            if (sl != kNoSlot) p.ent_item[__ldg(p.off_item + pid) + sl] = pack_entry((uint32_t)b | kDirect, cpos);
        }
        if (p.bin_cursor_user) {
            if (uid != 0) {
                const uint32_t at = atomicAdd(p.bin_cursor_user + (size_t)((uint32_t)uid >> p.bin_shift_user) * kCursorStride, 1u);
                const uint32_t lr = ((uint32_t)uid & ((1u << p.bin_shift_user) - 1u)) << (31 - p.bin_shift_user);
                p.ent_user[at] = pack_entry((uint32_t)b | kDirect | lr, 1.0f);
            }
        } else {
            uint32_t su = p.slot_user[b];
            if (su != kNoSlot) p.ent_user[__ldg(p.off_user + uid) + su] = pack_entry((uint32_t)b | kDirect, 1.0f);
        }
    }
}

// VPL = float4 per lane per row (D <= 128*VPL); lanes whose columns are >= D are idle.
//
// PARTIAL = owner-compute step of the row-sharded table (shard.cu): the "user table" is the all-gathered
// query matrix (query b = row b), the negatives are this owner's compacted sub-list of ncount[b] LOCAL row
// ids, the positive score comes from its owner through sp_in, and instead of the final loss / dq the warp
// leaves its partial state (raw accumulator + {csum, loss} | {m, l}) for shard_finish_kernel.
//
// MODE: 0 = load-then-reduce loop; 1 = software-pipelined stream (PIPE); 2 = mode 0 with L2 eviction-priority
// hints (HINT): table rows evict_first, CSR offsets / entry list evict_last (p.hint selects which); 3 = mode 0
// compiled for 4 CTAs/SM (64 registers, 32 warps/SM instead of 24).
template <int VPL, int LOSS, int SCORE, bool MULTI, int MODE, bool PARTIAL = false>
__global__ void __launch_bounds__(kThreads, (VPL == 1 && MODE == 3) ? 4 : ((VPL == 1 && MODE != 1) ? 3 : 2))
pair_fwd_kernel(const FwdParams p) {
    constexpr bool PIPE = MODE == 1, HINT = MODE == 2;
    static_assert(!PARTIAL || (!MULTI && !PIPE), "the owner-compute step runs one query per warp");
    uint64_t pol_row = 0, pol_off = 0, pol_ent = 0;
    if (HINT) {
        pol_row = l2_policy((p.hint & 1) ? 1 : 0);
        pol_off = l2_policy((p.hint & 2) ? 2 : 0);
        pol_ent = l2_policy((p.hint & 4) ? 2 : 0);
    }
    constexpr int LOADS = Cfg<VPL>::LOADS, REP = Cfg<VPL>::REP, NG = Cfg<VPL>::NG;
    constexpr float kRepInv = 1.0f / REP;
    static_assert(NG % 2 == 0, "the pipelined loop alternates two row buffers");

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

    int64_t uid = b, pid = 0;
    if (!PARTIAL) {
        uid = p.user[b]; pid = p.pos[b];
        if (uid < 0 || uid >= p.num_users) uid = 0;
        if (pid < 0 || pid >= p.num_items) pid = 0;
    }

    float4 q[VPL], vp[VPL];
    bool act[VPL];
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
        int col = lane * 4 + t * 128;
        act[t] = col < D;
        q[t] = act[t] ? ldg128(p.w_user + (size_t)uid * D + col) : make_float4(0, 0, 0, 0);
        vp[t] = (act[t] && !PARTIAL) ? ldg128(p.w_item + (size_t)pid * D + col) : make_float4(0, 0, 0, 0);
    }

    // this warp's slice of the negatives
    const int n = p.n;
    int j0 = 0, j1 = n;
    if (PARTIAL) j1 = min(n, max(0, __ldg(p.ncount + b)));
    if (MULTI) {
        int per = ((n + kWarps - 1) / kWarps + 31) & ~31;   // multiple of 32 so batches stay aligned
        j0 = min(n, warp * per);
        j1 = min(n, j0 + per);
    }
    const size_t rowbase = (size_t)b * n;

    // start the stream before the (dependent) positive-score reduction
    BatchMeta cur = load_meta<LOSS, HINT>(p, rowbase, j0, j1, lane, pol_row);
    BatchMeta nxt = cur;
    RowBuf<VPL> bufA, bufB;
    if (PIPE) {
        if (j0 < j1) load_group<VPL>(bufA, p, cur.id, 0, lane, act);
        if (j0 + 32 < j1) nxt = load_meta<LOSS>(p, rowbase, j0 + 32, j1, lane);
    }

    float sp = 0.f;
#pragma unroll
    for (int t = 0; t < VPL; ++t) sp += (SCORE == RSB200_SCORE_IP) ? dot4(q[t], vp[t]) : sqdist4(q[t], vp[t]);
    sp = warp_sum(sp);
    if (SCORE == RSB200_SCORE_EUCLID) sp = -sp;
    if (PARTIAL) sp = __ldg(p.sp_in + b);

    if (!PARTIAL && (!MULTI || warp == 0)) {
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

    WarpState<VPL> st;
#pragma unroll
    for (int t = 0; t < VPL; ++t) st.acc[t] = make_float4(0, 0, 0, 0);
    st.csum = 0.f; st.lossacc = 0.f; st.m_run = -INFINITY; st.l_run = 0.f;

    for (int jb = j0; jb < j1; jb += 32) {
        const bool valid = (jb + lane) < j1;
        uint32_t epos = 0;
        if (HINT && (p.hint & 8)) cur.slot = kNoSlot;        // variant 32 (timing diagnostic only): no grouping metadata
        if (cur.slot != kNoSlot)                                                  // consumed after the groups
            epos = p.bin_cursor ? bin_reserve(p, cur.id)
                 : p.slot_abs ? cur.slot
                              : (HINT ? ldg32_hint(p.off_item + cur.id, pol_off) : __ldg(p.off_item + cur.id)) + cur.slot;
        st.val_out = 0.f; st.sc_out = 0.f;

        if (PIPE) {
            BatchMeta nn = nxt;                                   // ids of batch jb+64 (arrive a batch ahead)
            if (jb + 64 < j1) nn = load_meta<LOSS>(p, rowbase, jb + 64, j1, lane);
#pragma unroll
            for (int g = 0; g < NG; g += 2) {
                // group g lives in bufA; fetch g+1 into bufB, then reduce g
                if (jb + (g + 1) * LOADS < j1) load_group<VPL>(bufB, p, cur.id, g + 1, lane, act);
                if (jb + g * LOADS < j1) compute_group<VPL, LOSS, SCORE>(bufA, st, p, q, sp, cur.lq, g, jb, j1, lane);
                // fetch g+2 (or group 0 of the next batch) into bufA, then reduce g+1
                if (g + 2 < NG) {
                    if (jb + (g + 2) * LOADS < j1) load_group<VPL>(bufA, p, cur.id, g + 2, lane, act);
                } else if (jb + 32 < j1) {
                    load_group<VPL>(bufA, p, nxt.id, 0, lane, act);
                }
                if (jb + (g + 1) * LOADS < j1) compute_group<VPL, LOSS, SCORE>(bufB, st, p, q, sp, cur.lq, g + 1, jb, j1, lane);
            }
            if (valid) {
                if (p.neg_score) p.neg_score[rowbase + jb + lane] = st.sc_out;
                if (cur.slot != kNoSlot) {
                    uint32_t low = (uint32_t)b | (LOSS == RSB200_LOSS_BPR ? kDirect : 0u);
                    if (p.bin_cursor) low = bin_low(p, low, cur.id);
                    p.ent_item[epos] = pack_entry(low, st.val_out);
                }
            }
            cur = nxt; nxt = nn;
        } else {
            // variant 3: fetch the ids of the NEXT batch now and ask L2 for its rows (one prefetch per
            // 128-byte line, issued by the lane that holds the id), so that next batch's loads are L2 hits
            if (p.prefetch && jb + 32 < j1) {
                nxt = load_meta<LOSS>(p, rowbase, jb + 32, j1, lane);
                if (jb + 32 + lane < j1) {
                    const char* rp = reinterpret_cast<const char*>(p.w_item + (size_t)nxt.id * D);
                    for (int o = 0; o < D * 4; o += 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(rp + o));
                }
            }
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                if (jb + g * LOADS < j1) {            // warp-uniform
                    load_group<VPL, HINT>(bufA, p, cur.id, g, lane, act, pol_row);
                    compute_group<VPL, LOSS, SCORE>(bufA, st, p, q, sp, cur.lq, g, jb, j1, lane);
                }
            }
            if (valid) {
                if (p.neg_score) p.neg_score[rowbase + jb + lane] = st.sc_out;
                if (p.cstage) {
                    p.cstage[rowbase + jb + lane] = st.val_out;          // touch order: coalesced; permuted into ent later
                } else if (cur.slot != kNoSlot) {
                    uint32_t low = (uint32_t)b | (LOSS == RSB200_LOSS_BPR ? kDirect : 0u);
                    if (p.bin_cursor) low = bin_low(p, low, cur.id);
                    const uint64_t en = pack_entry(low, st.val_out);
                    if (HINT) stg64_hint(p.ent_item + epos, en, pol_ent);
                    else p.ent_item[epos] = en;
                }
            }
            if (p.prefetch) cur = nxt;
            else if (jb + 32 < j1) cur = load_meta<LOSS, HINT>(p, rowbase, jb + 32, j1, lane, pol_row);
        }
    }

    if (PARTIAL) {
        const float csum = warp_sum(st.csum) * kRepInv, lossacc = warp_sum(st.lossacc) * kRepInv;
#pragma unroll
        for (int t = 0; t < VPL; ++t)
            if (act[t]) *reinterpret_cast<float4*>(p.dq_buf + (size_t)b * D + lane * 4 + t * 128) = st.acc[t];
        if (lane == 0)
            *reinterpret_cast<float2*>(p.stats_part + 2 * (size_t)b) =
                (LOSS == RSB200_LOSS_BPR) ? make_float2(csum, lossacc) : make_float2(st.m_run, st.l_run);
        return;
    }
    finish_query<VPL, LOSS, SCORE, MULTI>(p, st, s_q, s_vp, s_acc, s_stat, act, sp, b, uid, pid, lane, warp);
}

// ------------------------------------------------------------------------------------------
// TMA variant (variant 2, d <= 128): every warp owns a private ring of kStages shared-memory
// stages; the 8 rows of a group are fetched by 8 lanes with cp.async.bulk (UBLKCP, one 512-B
// bulk copy per row) that complete on the stage's mbarrier, so the bytes in flight per warp are
// bounded by shared memory (kStages x 4 KB) instead of registers.  The consumer side (same warp)
// waits on the mbarrier phase, pulls the rows into registers with conflict-free 16-byte LDS and
// immediately re-arms the stage for the group kStages ahead.
constexpr int kStages = 3;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int LOSS, int SCORE, bool MULTI>
__global__ void __launch_bounds__(kThreads, 2)
pair_fwd_tma_kernel(const FwdParams p) {
    constexpr int VPL = 1;
    constexpr int LOADS = Cfg<VPL>::LOADS, NG = Cfg<VPL>::NG;
    extern __shared__ __align__(128) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = p.D;
    const int b = MULTI ? blockIdx.x : blockIdx.x * kWarps + warp;
    if (!MULTI && b >= p.B) return;

    // layout: [ring: kWarps x kStages x LOADS x D floats][mbarriers: kWarps x kStages x u64][merge area]
    float* ring = smem + (size_t)warp * kStages * LOADS * D;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + (size_t)kWarps * kStages * LOADS * D) + warp * kStages;
    float* merge = smem + (size_t)kWarps * kStages * LOADS * D + 2 * kWarps * kStages;
    const int nw = MULTI ? kWarps : 1;
    float* sm_base = MULTI ? merge : merge + (size_t)warp * (size_t)(3 * D + 4);
    float* s_q = sm_base;
    float* s_vp = s_q + D;
    float* s_acc = s_vp + D;
    float* s_stat = s_acc + (size_t)nw * D;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(smem_u32(bars + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    int64_t uid = p.user[b], pid = p.pos[b];
    if (uid < 0 || uid >= p.num_users) uid = 0;
    if (pid < 0 || pid >= p.num_items) pid = 0;
    float4 q[VPL], vp[VPL];
    bool act[VPL];
    act[0] = lane * 4 < D;
    q[0] = act[0] ? ldg128(p.w_user + (size_t)uid * D + lane * 4) : make_float4(0, 0, 0, 0);
    vp[0] = act[0] ? ldg128(p.w_item + (size_t)pid * D + lane * 4) : make_float4(0, 0, 0, 0);

    const int n = p.n;
    int j0 = 0, j1 = n;
    if (MULTI) {
        int per = ((n + kWarps - 1) / kWarps + 31) & ~31;
        j0 = min(n, warp * per);
        j1 = min(n, j0 + per);
    }
    const size_t rowbase = (size_t)b * n;
    const int total_groups = (j1 - j0 + LOADS - 1) / LOADS;
    const uint32_t row_bytes = (uint32_t)D * 4u;

    BatchMeta cur = load_meta<LOSS>(p, rowbase, j0, j1, lane);
    BatchMeta nxt = cur, nn = cur;
    if (j0 + 32 < j1) nxt = load_meta<LOSS>(p, rowbase, j0 + 32, j1, lane);
    if (j0 + 64 < j1) nn = load_meta<LOSS>(p, rowbase, j0 + 64, j1, lane);

    // arm stage (gi % kStages) with the LOADS rows of global group gi; ids come from the batch it belongs to
    auto issue = [&](int gi, int cur_batch) {
        const int bi = gi / NG;                                  // batch of this group: cur_batch or cur_batch + 1
        const int src = (gi % NG) * LOADS + (lane & (LOADS - 1));
        const int rid_cur = __shfl_sync(kFull, cur.id, src);
        const int rid_nxt = __shfl_sync(kFull, nxt.id, src);
        const int rid = (bi == cur_batch) ? rid_cur : rid_nxt;
        const int s = gi % kStages;
        const uint32_t bar = smem_u32(bars + s);
        if (lane == 0) mbar_expect_tx(bar, row_bytes * LOADS);
        if (lane < LOADS)
            bulk_g2s(smem_u32(ring + ((size_t)s * LOADS + lane) * D), p.w_item + (size_t)rid * D, row_bytes, bar);
    };
#pragma unroll
    for (int s = 0; s < kStages; ++s)
        if (s < total_groups) issue(s, 0);

    float sp = (SCORE == RSB200_SCORE_IP) ? dot4(q[0], vp[0]) : sqdist4(q[0], vp[0]);
    sp = warp_sum(sp);
    if (SCORE == RSB200_SCORE_EUCLID) sp = -sp;
    if (!MULTI || warp == 0) {
        if (act[0]) {
            *reinterpret_cast<float4*>(s_q + lane * 4) = q[0];
            *reinterpret_cast<float4*>(s_vp + lane * 4) = vp[0];
            *reinterpret_cast<float4*>(p.q_buf + (size_t)b * D + lane * 4) = q[0];
        }
    }

    WarpState<VPL> st;
    st.acc[0] = make_float4(0, 0, 0, 0);
    st.csum = 0.f; st.lossacc = 0.f; st.m_run = -INFINITY; st.l_run = 0.f;
    st.val_out = 0.f; st.sc_out = 0.f;
    uint32_t epos = 0;
    if (cur.slot != kNoSlot) epos = p.bin_cursor ? bin_reserve(p, cur.id) : (p.slot_abs ? cur.slot : __ldg(p.off_item + cur.id) + cur.slot);

    int batch = 0;
    for (int gi = 0; gi < total_groups; ++gi) {
        const int g = gi % NG, s = gi % kStages;
        const int jb = j0 + batch * 32;
        mbar_wait(smem_u32(bars + s), (uint32_t)((gi / kStages) & 1));
        RowBuf<VPL> buf;
#pragma unroll
        for (int k = 0; k < LOADS; ++k)
            buf.v[k][0] = act[0] ? *reinterpret_cast<const float4*>(ring + ((size_t)s * LOADS + k) * D + lane * 4)
                                 : make_float4(0, 0, 0, 0);
        __syncwarp();                                            // all lanes have read the stage
        if (gi + kStages < total_groups) issue(gi + kStages, batch);
        compute_group<VPL, LOSS, SCORE>(buf, st, p, q, sp, cur.lq, g, jb, j1, lane);
        if (g == NG - 1 || gi == total_groups - 1) {            // batch complete: emit and rotate metadata
            if (jb + lane < j1) {
                if (p.neg_score) p.neg_score[rowbase + jb + lane] = st.sc_out;
                if (cur.slot != kNoSlot) {
                    uint32_t low = (uint32_t)b | (LOSS == RSB200_LOSS_BPR ? kDirect : 0u);
                    if (p.bin_cursor) low = bin_low(p, low, cur.id);
                    p.ent_item[epos] = pack_entry(low, st.val_out);
                }
            }
            ++batch;
            cur = nxt; nxt = nn;
            const int jn = j0 + (batch + 2) * 32;
            if (jn < j1) nn = load_meta<LOSS>(p, rowbase, jn, j1, lane);
            epos = 0;
            if (j0 + batch * 32 >= j1) cur.slot = kNoSlot;          // past the last batch: `cur` is stale -- reserve nothing
            if (cur.slot != kNoSlot) epos = p.bin_cursor ? bin_reserve(p, cur.id) : (p.slot_abs ? cur.slot : __ldg(p.off_item + cur.id) + cur.slot);
            st.val_out = 0.f; st.sc_out = 0.f;
        }
    }
    finish_query<VPL, LOSS, SCORE, MULTI>(p, st, s_q, s_vp, s_acc, s_stat, act, sp, b, uid, pid, lane, warp);
}

template <int LOSS, int SCORE>
static int32_t launch_fwd_tma_ls(const FwdParams& p, cudaStream_t st) {
    const bool multi = p.n >= 256;
    const size_t ring = (size_t)kWarps * kStages * Cfg<1>::LOADS * p.D * sizeof(float) + (size_t)kWarps * kStages * 8;
    if (multi) {
        size_t smem = ring + (size_t)(2 * p.D + kWarps * p.D + kWarps * 4) * sizeof(float);
        RSB_CUDA(cudaFuncSetAttribute(pair_fwd_tma_kernel<LOSS, SCORE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pair_fwd_tma_kernel<LOSS, SCORE, true><<<p.B, kThreads, smem, st>>>(p);
    } else {
        size_t smem = ring + (size_t)kWarps * (3 * p.D + 4) * sizeof(float);
        RSB_CUDA(cudaFuncSetAttribute(pair_fwd_tma_kernel<LOSS, SCORE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pair_fwd_tma_kernel<LOSS, SCORE, false><<<(unsigned)cdiv(p.B, kWarps), kThreads, smem, st>>>(p);
    }
    RSB_LAUNCH_CHECK();
    return 0;
}

static int32_t launch_fwd_tma(const FwdParams& p, int loss, int score, cudaStream_t st) {
    if (loss == RSB200_LOSS_BPR) {
        return score == RSB200_SCORE_IP ? launch_fwd_tma_ls<RSB200_LOSS_BPR, RSB200_SCORE_IP>(p, st)
                                        : launch_fwd_tma_ls<RSB200_LOSS_BPR, RSB200_SCORE_EUCLID>(p, st);
    }
    return score == RSB200_SCORE_IP ? launch_fwd_tma_ls<RSB200_LOSS_SSM, RSB200_SCORE_IP>(p, st)
                                    : launch_fwd_tma_ls<RSB200_LOSS_SSM, RSB200_SCORE_EUCLID>(p, st);
}

template <int VPL, int LOSS, int SCORE, int PIPE>
static int32_t launch_fwd_vlsp(const FwdParams& p, cudaStream_t st) {
    const bool multi = p.n >= 256;
    if (multi) {
        size_t smem = (size_t)(2 * p.D + kWarps * p.D + kWarps * 4) * sizeof(float);
        pair_fwd_kernel<VPL, LOSS, SCORE, true, PIPE><<<p.B, kThreads, smem, st>>>(p);
    } else {
        size_t smem = (size_t)kWarps * (3 * p.D + 4) * sizeof(float);
        if (smem > 48 * 1024)            // d = 512: 49 KB, above the default dynamic shared-memory limit
            RSB_CUDA(cudaFuncSetAttribute(pair_fwd_kernel<VPL, LOSS, SCORE, false, PIPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pair_fwd_kernel<VPL, LOSS, SCORE, false, PIPE><<<(unsigned)cdiv(p.B, kWarps), kThreads, smem, st>>>(p);
    }
    RSB_LAUNCH_CHECK();
    return 0;
}

template <int VPL, int PIPE>
static int32_t launch_fwd_v(const FwdParams& p, int loss, int score, cudaStream_t st) {
    if (loss == RSB200_LOSS_BPR) {
        return score == RSB200_SCORE_IP ? launch_fwd_vlsp<VPL, RSB200_LOSS_BPR, RSB200_SCORE_IP, PIPE>(p, st)
                                        : launch_fwd_vlsp<VPL, RSB200_LOSS_BPR, RSB200_SCORE_EUCLID, PIPE>(p, st);
    }
    return score == RSB200_SCORE_IP ? launch_fwd_vlsp<VPL, RSB200_LOSS_SSM, RSB200_SCORE_IP, PIPE>(p, st)
                                    : launch_fwd_vlsp<VPL, RSB200_LOSS_SSM, RSB200_SCORE_EUCLID, PIPE>(p, st);
}

// variant 0 (default): load-then-reduce loop, 3 CTAs/SM.  variant 1: software-pipelined stream,
// 2 CTAs/SM -- measured SLOWER on B200 (1.12 ms vs 0.875 ms at config 2: the occupancy loss
// outweighs the deeper per-warp pipeline), kept for A/B runs.
int32_t launch_pair_fwd(const FwdParams& p, int loss, int score, int variant, cudaStream_t st) {
    if (p.B == 0) return 0;
    const bool pipe = variant == 1;
    if (variant == 2 && p.D <= 128) return launch_fwd_tma(p, loss, score, st);   // TMA (cp.async.bulk) ring
    if (p.D <= 128) {
        if (p.hint) return launch_fwd_v<1, 2>(p, loss, score, st);               // variants 16..31: L2 eviction hints
        if (variant == 4) return launch_fwd_v<1, 3>(p, loss, score, st);         // 4 CTAs/SM
        return pipe ? launch_fwd_v<1, 1>(p, loss, score, st) : launch_fwd_v<1, 0>(p, loss, score, st);
    }
    if (p.D <= 256) return launch_fwd_v<2, 0>(p, loss, score, st);
    if (p.D <= 512) return launch_fwd_v<4, 0>(p, loss, score, st);
    set_error("embedding dim %d > 512 is not supported", p.D);
    return RSB200_EUNSUPPORTED;
}

template <int VPL>
static int32_t launch_partial_v(const FwdParams& p, int loss, int score, cudaStream_t st) {
    const size_t smem = (size_t)kWarps * (3 * p.D + 4) * sizeof(float);
    const unsigned grid = (unsigned)cdiv(p.B, kWarps);
#define RSB_PARTIAL(LOSS, SCORE)                                                                                      \
    do {                                                                                                              \
        if (smem > 48 * 1024)                                                                                         \
            RSB_CUDA(cudaFuncSetAttribute(pair_fwd_kernel<VPL, LOSS, SCORE, false, 0, true>,                      \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                   \
        pair_fwd_kernel<VPL, LOSS, SCORE, false, 0, true><<<grid, kThreads, smem, st>>>(p);                       \
    } while (0)
    if (loss == RSB200_LOSS_BPR) {
        if (score == RSB200_SCORE_IP) RSB_PARTIAL(RSB200_LOSS_BPR, RSB200_SCORE_IP);
        else RSB_PARTIAL(RSB200_LOSS_BPR, RSB200_SCORE_EUCLID);
    } else {
        if (score == RSB200_SCORE_IP) RSB_PARTIAL(RSB200_LOSS_SSM, RSB200_SCORE_IP);
        else RSB_PARTIAL(RSB200_LOSS_SSM, RSB200_SCORE_EUCLID);
    }
#undef RSB_PARTIAL
    RSB_LAUNCH_CHECK();
    return 0;
}

// owner-compute forward of the row-sharded step (one query per warp, see PARTIAL above)
int32_t launch_pair_fwd_partial(const FwdParams& p, int loss, int score, cudaStream_t st) {
    if (p.B == 0) return 0;
    if (p.D <= 128) return launch_partial_v<1>(p, loss, score, st);
    if (p.D <= 256) return launch_partial_v<2>(p, loss, score, st);
    if (p.D <= 512) return launch_partial_v<4>(p, loss, score, st);
    set_error("embedding dim %d > 512 is not supported", p.D);
    return RSB200_EUNSUPPORTED;
}

}  // namespace rsb
