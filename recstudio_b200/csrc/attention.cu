// attention.cu -- A1: the attention core of SASRecQueryEncoder / BERT4Rec on the 5th-gen tensor
// cores (tcgen05.mma, accumulators in TMEM), the one place on this path that is a dense
// bf16 contraction.
//   reference: recstudio/model/seq/sasrec.py:19-32,44-53 (nn.TransformerEncoderLayer self-attention,
//   causal mask triu(ones(L,L),1) AND key-padding mask hist == 0, softmax(QK^T/sqrt(dh)) V per head)
//
// One CTA (128 threads = 4 warps = 128 TMEM lanes) per (sequence, head); L <= 256, head_dim = 64.
//   S  = Q_tile K^T    : tcgen05.mma M128 N256 K16 x4   -> TMEM columns [0, 256)
//   P  = softmax(mask(S / 8)) : every thread owns one query row (tcgen05.ld 32x32b), no shuffles;
//        P is written to shared memory as bf16 in the UMMA K-major operand layout
//   O  = P V           : tcgen05.mma M128 N64 K16 x16   -> TMEM columns [256, 320)
// Operands are staged with plain loads (fp32 -> bf16 conversion on the fly) into the
// interleaved no-swizzle layout of tc05.cuh; V is transposed while staging so that both GEMMs
// use K-major operands.
#include "common.cuh"
#include "kernels.h"
#include "tc05.cuh"

namespace rsb {
using namespace tc;

// ---------------------------------------------------------------------------------------------
// Debug / validation entry: D[128, N] = A[128, K] * B[N, K]^T with bf16 operands, fp32 accumulate.
__global__ void __launch_bounds__(128)
tc_gemm_test_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int K,
                    uint32_t* __restrict__ err) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem_raw);                       // [128 x K]
    __nv_bfloat16* sB = reinterpret_cast<__nv_bfloat16*>(smem_raw + (size_t)128 * K * 2); // [N x K]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / K, c = i % K;
        *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<unsigned char*>(sA) + kmajor_off(r, c, 128)) = __float2bfloat16(A[i]);
    }
    for (int i = tid; i < N * K; i += 128) {
        const int r = i / K, c = i % K;
        *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<unsigned char*>(sB) + kmajor_off(r, c, N)) = __float2bfloat16(B[i]);
    }
    uint32_t ncols = 32;
    while ((int)ncols < N) ncols <<= 1;
    if (warp == 0) {
        tmem_alloc(&tmem_slot, ncols);
        if (lane == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t taddr = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N);
        for (int k16 = 0; k16 < K / 16; ++k16) {
            const uint64_t ad = make_desc(smem_addr(sA) + k16 * 2 * (128 * 16), 128 * 16, 128);
            const uint64_t bd = make_desc(smem_addr(sB) + k16 * 2 * (N * 16), N * 16, 128);
            mma_bf16(taddr, ad, bd, idesc, k16 > 0);
        }
        mma_commit(&bar);
    }
    const bool ok = mbar_wait(&bar, 0);
    fence_after_sync();
    if (!ok) {
        if (tid == 0) *err = 1u;
    } else {
        for (int c0 = 0; c0 < N; c0 += 32) {
            float v[32];
            tmem_ld32(taddr + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (c0 + i < N) D[(size_t)(warp * 32 + lane) * N + c0 + i] = v[i];
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(taddr, ncols);
}

// ---------------------------------------------------------------------------------------------
constexpr int kLP = 256;       // padded sequence length (keys)
constexpr int kDH = 64;        // head dim
constexpr uint32_t kTmemCols = 512;
constexpr size_t kAttnSmem = (size_t)kLP * kDH * 2 * 3 + (size_t)128 * kLP * 2 + kLP;   // Q, K, Vt, P, key-valid bytes

// ---------------------------------------------------------------------------------------------
// Operand staging: fp32 global -> bf16 shared memory in the interleaved K-major layout, one 16-byte
// shared-memory store per 8 elements (a whole 16-byte chunk of the layout), conflict-free.
__device__ __forceinline__ uint4 pack8(const float4 a, const float4 b) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
    u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
    return u;
}
// natural tile: dst[r][c] = src[(row0 + r) * ld + c], r < R (rows >= L are zero), c < 64
__device__ __forceinline__ void stage_natural(unsigned char* dst, const float* __restrict__ src, int row0, int L, int ld, int R, int tid, int nt) {
#pragma unroll 4                                               // 8 independent 16-byte loads in flight per thread
    for (int i = tid; i < R * (kDH / 8); i += nt) {
        const int r = i % R, cg = i / R;                     // consecutive threads -> consecutive rows: contiguous 16-B stores
        float4 a = make_float4(0, 0, 0, 0), b = a;
        if (row0 + r < L) {
            const float* p = src + (size_t)(row0 + r) * ld + cg * 8;
            a = ldg128(p); b = ldg128(p + 4);
        }
        *reinterpret_cast<uint4*>(dst + (size_t)cg * (R * 16) + r * 16) = pack8(a, b);
    }
}
// transposed tile [64 x 256]: dst[c][j] = src[j * ld + c]  (rows = head dim, K = sequence positions)
__device__ __forceinline__ void stage_transposed(unsigned char* dst, const float* __restrict__ src, int L, int ld, int tid, int nt) {
#pragma unroll 2
    for (int i = tid; i < (kLP / 8) * (kDH / 4); i += nt) {
        const int c4 = i % (kDH / 4), jg = i / (kDH / 4);    // 16 consecutive threads read 256 contiguous bytes of one row
        float4 f[8];
#pragma unroll
        for (int x = 0; x < 8; ++x) {
            const int j = jg * 8 + x;
            f[x] = (j < L) ? ldg128(src + (size_t)j * ld + c4 * 4) : make_float4(0, 0, 0, 0);
        }
        unsigned char* base = dst + (size_t)jg * (kDH * 16) + (c4 * 4) * 16;
        *reinterpret_cast<uint4*>(base) = pack8(make_float4(f[0].x, f[1].x, f[2].x, f[3].x), make_float4(f[4].x, f[5].x, f[6].x, f[7].x));
        *reinterpret_cast<uint4*>(base + 16) = pack8(make_float4(f[0].y, f[1].y, f[2].y, f[3].y), make_float4(f[4].y, f[5].y, f[6].y, f[7].y));
        *reinterpret_cast<uint4*>(base + 32) = pack8(make_float4(f[0].z, f[1].z, f[2].z, f[3].z), make_float4(f[4].z, f[5].z, f[6].z, f[7].z));
        *reinterpret_cast<uint4*>(base + 48) = pack8(make_float4(f[0].w, f[1].w, f[2].w, f[3].w), make_float4(f[4].w, f[5].w, f[6].w, f[7].w));
    }
}

// Thread layout of the attention kernels: kNW warps share every TMEM lane quarter (a warp can only touch
// lanes 32*(warp % 4) .. +31) and split the COLUMNS, so a CTA has 4*kNW warps working on the softmax /
// elementwise phases instead of 4 (ncu of the 4-warp version: tensor pipe 3 % active, 6 % warp occupancy,
// 57 % of the samples in the per-row softmax).
constexpr int kNW = 4;
constexpr int kNT = 128 * kNW;

// Attention-probability dropout (nn.MultiheadAttention(dropout=p), sasrec.py:19-32 / seq/config/sasrec.yaml:5): the keep mask
// of element (sequence-head bh, query i, key j) is a counter-based hash of (key, bh, i, j) -- no state, so the forward
// kernel and BOTH roles of the backward kernel regenerate the very same mask from the 64-bit key (drawn from torch's CUDA
// generator by the host).  lowbias32 (two xorshift-multiply rounds, full avalanche) twice: ~14 integer ops per element.
struct DropSpec {
    uint32_t key_lo, key_hi;
    uint32_t thresh;        // drop iff hash < thresh;  thresh = round(p * 2^32), 0 = no dropout
    float keep_scale;       // 1 / (1 - p)
};
__device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ bool drop_keep(const DropSpec& ds, uint32_t bh, int qi, int kj) {
    uint32_t x = lowbias32(((uint32_t)qi << 8 | (uint32_t)kj) ^ ds.key_lo);      // qi, kj < 256
    x = lowbias32(x ^ (bh * 0x9E3779B1u) ^ ds.key_hi);
    return x >= ds.thresh;
}

__global__ void __launch_bounds__(kNT, 1)
attn_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                const int64_t* __restrict__ hist, int L, int heads, int causal, float scale, const DropSpec drop,
                float* __restrict__ out, float* __restrict__ lse_out, uint32_t* __restrict__ err) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sQ = smem_raw;                                  // 2 x [128 x 64] bf16 (one block per M tile)
    unsigned char* sK = sQ + (size_t)kLP * kDH * 2;                // [256 x 64]
    unsigned char* sVt = sK + (size_t)kLP * kDH * 2;               // [64 x 256]  (rows = head dim, K = keys)
    unsigned char* sP = sVt + (size_t)kLP * kDH * 2;               // [128 x 256]
    unsigned char* keyok = sP + (size_t)128 * kLP * 2;             // [256]
    __shared__ __align__(8) uint64_t bar_s, bar_o;
    __shared__ uint32_t tmem_slot;
    __shared__ float s_max[kNW][128], s_sum[kNW][128];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cq = warp >> 2;                                      // column slice of this warp
    const int row = (warp & 3) * 32 + lane;                        // TMEM lane = query row inside the M tile
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const int d = heads * kDH;
    const float* qb = q + (size_t)b * L * d + h * kDH;
    const float* kb = k + (size_t)b * L * d + h * kDH;
    const float* vb = v + (size_t)b * L * d + h * kDH;
    constexpr int CW = kLP / kNW;                                  // score columns per warp slice (64)
    constexpr int OW = kDH / kNW;                                  // output columns per warp slice (16)

    // ---- stage operands (fp32 -> bf16), zero beyond L ---------------------------------------
    stage_natural(sQ, qb, 0, L, d, 128, tid, kNT);
    stage_natural(sQ + 128 * kDH * 2, qb, 128, L, d, 128, tid, kNT);
    stage_natural(sK, kb, 0, L, d, kLP, tid, kNT);
    stage_transposed(sVt, vb, L, d, tid, kNT);
    for (int j = tid; j < kLP; j += kNT) keyok[j] = (j < L && (hist == nullptr || hist[(size_t)b * L + j] != 0)) ? 1 : 0;
    if (warp == 0) {
        tmem_alloc(&tmem_slot, kTmemCols);
        if (lane == 0) { mbar_init(&bar_s, 1); mbar_init(&bar_o, 1); fence_mbar_init(); }
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tS = tmem_slot, tO = tmem_slot + 256;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    bool failed = false;

    const int mtiles = (L + 127) / 128;
    for (int mt = 0; mt < mtiles; ++mt) {
        if (tid == 0) {
            const uint32_t idesc = make_idesc_bf16(128, kLP);
            const uint32_t qa = smem_addr(sQ) + mt * (128 * kDH * 2), ka = smem_addr(sK);
#pragma unroll
            for (int k16 = 0; k16 < kDH / 16; ++k16)
                mma_bf16(tS, make_desc(qa + k16 * 2 * (128 * 16), 128 * 16, 128),
                         make_desc(ka + k16 * 2 * (kLP * 16), kLP * 16, 128), idesc, k16 > 0);
            mma_commit(&bar_s);
        }
        if (!mbar_wait(&bar_s, mt & 1)) failed = true;
        fence_after_sync();
        const int gi = mt * 128 + row;                               // this thread's query row
        // pass 1: max of the masked, scaled scores over this warp's column slice, then across the slices
        float mx = -INFINITY;
        for (int c0 = cq * CW; c0 < (cq + 1) * CW; c0 += 32) {
            float s[32];
            tmem_ld32(tS + lane_base + c0, s);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int j = c0 + i;
                const bool ok = keyok[j] && !(causal && j > gi);
                if (ok) mx = fmaxf(mx, s[i] * scale);
            }
        }
        s_max[cq][row] = mx;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < kNW; ++w) mx = fmaxf(mx, s_max[w][row]);
        // pass 2: P = exp(s - max) (bf16, UMMA K-major layout), partial row sum
        float l = 0.f;
        for (int c0 = cq * CW; c0 < (cq + 1) * CW; c0 += 32) {
            float s[32];
            tmem_ld32(tS + lane_base + c0, s);
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
                __align__(16) __nv_bfloat16 pk[8];
#pragma unroll
                for (int x = 0; x < 8; ++x) {
                    const int j = c0 + g8 * 8 + x;
                    const bool ok = keyok[j] && !(causal && j > gi);
                    const float p = (ok && mx != -INFINITY) ? __expf(s[g8 * 8 + x] * scale - mx) : 0.f;
                    pk[x] = __float2bfloat16(p);
                    l += __bfloat162float(pk[x]);            // normalise by what the tensor core will actually sum
                    // dropout acts on the NORMALISED probabilities: the denominator keeps every element, the P V product
                    // only the kept ones (scaled by 1 / (1 - p) together with 1 / l below)
                    if (drop.thresh && !drop_keep(drop, (uint32_t)blockIdx.x, gi, j)) pk[x] = __float2bfloat16(0.f);
                }
                *reinterpret_cast<uint4*>(sP + kmajor_off(row, c0 + g8 * 8, 128)) = *reinterpret_cast<const uint4*>(pk);
            }
        }
        s_sum[cq][row] = l;
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        if (tid == 0) {
            const uint32_t idesc = make_idesc_bf16(128, kDH);
            const uint32_t pa = smem_addr(sP), va = smem_addr(sVt);
#pragma unroll
            for (int k16 = 0; k16 < kLP / 16; ++k16)
                mma_bf16(tO, make_desc(pa + k16 * 2 * (128 * 16), 128 * 16, 128),
                         make_desc(va + k16 * 2 * (kDH * 16), kDH * 16, 128), idesc, k16 > 0);
            mma_commit(&bar_o);
        }
        l = 0.f;
#pragma unroll
        for (int w = 0; w < kNW; ++w) l += s_sum[w][row];
        if (!mbar_wait(&bar_o, mt & 1)) failed = true;
        fence_after_sync();
        const float inv = l > 0.f ? drop.keep_scale / l : 0.f;
        {
            float o[16];
            tmem_ld16(tO + lane_base + cq * OW, o);
            if (gi < L) {
                float* dst = out + ((size_t)b * L + gi) * d + h * kDH + cq * OW;
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    *reinterpret_cast<float4*>(dst + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
            }
        }
        if (cq == 0 && gi < L && lse_out) lse_out[((size_t)b * heads + h) * L + gi] = (l > 0.f) ? mx + logf(l) : -INFINITY;
        fence_before_sync();
        __syncthreads();                 // sP / TMEM / s_max / s_sum are reused by the next M tile
        fence_after_sync();
    }
    if (failed && tid == 0) *err = 1u;
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// Backward.  With P = softmax(mask(S / 8)), D_i = sum_c dO_ic O_ic:
//     dV = P^T dO,   dP = dO V^T,   dS = P o (dP - D) / 8,   dQ = dS K,   dK = dS^T Q.
// One kernel, two roles (template KEYSIDE), so that every tensor-core A operand is written by the
// thread that OWNS its row (no transposed scatter through shared memory):
//   query side (KEYSIDE = 0): rows = queries, columns = keys  ->  dQ = dS  K
//   key   side (KEYSIDE = 1): rows = keys,    columns = queries -> dV = P^T dO,  dK = dS^T Q
// (the key side recomputes S^T = K Q^T and dP^T = V dO^T directly).  Columns are processed in
// halves of 128 so that S, dP (2 x 128 TMEM columns) and the output accumulators (2 x 64) coexist.
constexpr size_t kBwdSmem = (size_t)128 * kDH * 2 * 4      // X tile, U tile, Y half, W half
                          + (size_t)128 * 128 * 2 * 2      // E_P, E_dS
                          + (size_t)kDH * kLP * 2 * 2      // Yt, Wt
                          + (size_t)kLP * 4 * 2 + kLP;     // lse, D, key-valid

template <bool KEYSIDE>
__global__ void __launch_bounds__(kNT, 1)
attn_bwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                const float* __restrict__ o, const float* __restrict__ d_o, const float* __restrict__ lse,
                const int64_t* __restrict__ hist, int L, int heads, int causal, float scale, const DropSpec drop,
                float* __restrict__ out_ds /* dQ | dK */, float* __restrict__ out_p /* - | dV */, uint32_t* __restrict__ err) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sX = smem_raw;                                  // [128 x 64] rows operand of S
    unsigned char* sU = sX + 128 * kDH * 2;                        // [128 x 64] rows operand of dP
    unsigned char* sYh = sU + 128 * kDH * 2;                       // [128 x 64] column-half operand of S
    unsigned char* sWh = sYh + 128 * kDH * 2;                      // [128 x 64] column-half operand of dP
    unsigned char* sEP = sWh + 128 * kDH * 2;                      // [128 x 128] P   (rows x column half)
    unsigned char* sEdS = sEP + 128 * 128 * 2;                     // [128 x 128] dS
    unsigned char* sYt = sEdS + 128 * 128 * 2;                     // [64 x 256]  Y^T (B operand of dS * Y)
    unsigned char* sWt = sYt + kDH * kLP * 2;                      // [64 x 256]  W^T (B operand of P * W), key side only
    float* s_lse = reinterpret_cast<float*>(sWt + kDH * kLP * 2);  // [256] per query
    float* s_D = s_lse + kLP;                                      // [256] per query
    unsigned char* keyok = reinterpret_cast<unsigned char*>(s_D + kLP);
    __shared__ __align__(8) uint64_t bar1, bar2;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const int d = heads * kDH;
    const size_t base = (size_t)b * L * d + h * kDH;
    const float* X = (KEYSIDE ? k : q) + base;        // rows operand of S
    const float* Y = (KEYSIDE ? q : k) + base;        // cols operand of S
    const float* U = (KEYSIDE ? v : d_o) + base;      // rows operand of dP
    const float* W = (KEYSIDE ? d_o : v) + base;      // cols operand of dP

    auto stage_rows = [&](unsigned char* dst, const float* src, int row0) { stage_natural(dst, src, row0, L, d, 128, tid, kNT); };
    auto stage_t = [&](unsigned char* dst, const float* src) { stage_transposed(dst, src, L, d, tid, kNT); };
    stage_t(sYt, Y);
    if (KEYSIDE) stage_t(sWt, W);
    for (int i = tid; i < kLP; i += kNT) {
        float dsum = 0.f, ls = 0.f;
        if (i < L) {
            const float* orow = o + base + (size_t)i * d;
            const float* grow = d_o + base + (size_t)i * d;
            for (int c = 0; c < kDH; c += 4) dsum += dot4(ldg128(orow + c), ldg128(grow + c));
            ls = lse[((size_t)b * heads + h) * L + i];
        }
        s_D[i] = dsum; s_lse[i] = ls;
        keyok[i] = (i < L && (hist == nullptr || hist[(size_t)b * L + i] != 0)) ? 1 : 0;
    }
    if (warp == 0) {
        tmem_alloc(&tmem_slot, kTmemCols);
        if (lane == 0) { mbar_init(&bar1, 1); mbar_init(&bar2, 1); fence_mbar_init(); }
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tS = tmem_slot, tdP = tmem_slot + 128, tOdS = tmem_slot + 256, tOP = tmem_slot + 320;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    bool failed = false;
    uint32_t ph1 = 0, ph2 = 0;
    const int tiles = (L + 127) / 128;
    const int row = (warp & 3) * 32 + lane;                    // TMEM lane of this thread
    const int cq = warp >> 2;                                  // column slice of this warp (see kNW)
    constexpr int CW = 128 / kNW, OW = kDH / kNW;
    static_assert(CW == 32 && OW == 16, "column slices are one tmem_ld32 / tmem_ld16 wide");

    for (int rt = 0; rt < tiles; ++rt) {
        stage_rows(sX, X, rt * 128);
        stage_rows(sU, U, rt * 128);
        const int gr = rt * 128 + row;                        // global row index of this thread
        for (int ch = 0; ch < tiles; ++ch) {
            stage_rows(sYh, Y, ch * 128);
            stage_rows(sWh, W, ch * 128);
            fence_async_smem();
            fence_before_sync();
            __syncthreads();
            fence_after_sync();
            if (tid == 0) {
                const uint32_t idesc = make_idesc_bf16(128, 128);
#pragma unroll
                for (int k16 = 0; k16 < kDH / 16; ++k16) {
                    const uint32_t koff = k16 * 2 * (128 * 16);
                    mma_bf16(tS, make_desc(smem_addr(sX) + koff, 128 * 16, 128), make_desc(smem_addr(sYh) + koff, 128 * 16, 128), idesc, k16 > 0);
                    mma_bf16(tdP, make_desc(smem_addr(sU) + koff, 128 * 16, 128), make_desc(smem_addr(sWh) + koff, 128 * 16, 128), idesc, k16 > 0);
                }
                mma_commit(&bar1);
            }
            if (!mbar_wait(&bar1, ph1)) failed = true;
            ph1 ^= 1;
            fence_after_sync();
            {
                const int c0 = cq * CW;
                float s[32], dp[32];
                tmem_ld32(tS + lane_base + c0, s);
                tmem_ld32(tdP + lane_base + c0, dp);
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                    __align__(16) __nv_bfloat16 pk[8], dk[8];
#pragma unroll
                    for (int x = 0; x < 8; ++x) {
                        const int gc = ch * 128 + c0 + g8 * 8 + x;         // global column index
                        const int qi = KEYSIDE ? gc : gr, kj = KEYSIDE ? gr : gc;
                        const bool ok = qi < L && kj < kLP && keyok[kj] && !(causal && kj > qi);
                        float pv = 0.f, dsv = 0.f;
                        if (ok) {
                            pv = __expf(s[g8 * 8 + x] * scale - s_lse[qi]);
                            // dropout: O = (P o M / (1 - p)) V  =>  dP = (dO V^T) o M / (1 - p), dV = (P o M / (1 - p))^T dO,
                            // and D_i = sum_j P_ij dP_ij = dO_i . O_i is unchanged (O is the dropped output)
                            const float m = (!drop.thresh || drop_keep(drop, (uint32_t)blockIdx.x, qi, kj)) ? drop.keep_scale : 0.f;
                            dsv = pv * (dp[g8 * 8 + x] * m - s_D[qi]) * scale;
                            pv *= m;
                        }
                        pk[x] = __float2bfloat16(pv); dk[x] = __float2bfloat16(dsv);
                    }
                    *reinterpret_cast<uint4*>(sEP + kmajor_off(row, c0 + g8 * 8, 128)) = *reinterpret_cast<const uint4*>(pk);
                    *reinterpret_cast<uint4*>(sEdS + kmajor_off(row, c0 + g8 * 8, 128)) = *reinterpret_cast<const uint4*>(dk);
                }
            }
            fence_async_smem();
            fence_before_sync();
            __syncthreads();
            fence_after_sync();
            if (tid == 0) {
                const uint32_t idesc = make_idesc_bf16(128, kDH);
#pragma unroll
                for (int k16 = 0; k16 < 128 / 16; ++k16) {
                    const uint32_t aoff = k16 * 2 * (128 * 16);
                    const uint32_t boff = (ch * 16 + k16 * 2) * (kDH * 16);       // column half ch, k-group pair k16
                    const uint32_t acc = (ch > 0 || k16 > 0) ? 1u : 0u;
                    mma_bf16(tOdS, make_desc(smem_addr(sEdS) + aoff, 128 * 16, 128), make_desc(smem_addr(sYt) + boff, kDH * 16, 128), idesc, acc);
                    if (KEYSIDE)
                        mma_bf16(tOP, make_desc(smem_addr(sEP) + aoff, 128 * 16, 128), make_desc(smem_addr(sWt) + boff, kDH * 16, 128), idesc, acc);
                }
                mma_commit(&bar2);
            }
            if (!mbar_wait(&bar2, ph2)) failed = true;     // E / Yh / Wh buffers are free again
            ph2 ^= 1;
            fence_after_sync();
        }
        {
            const int c0 = cq * OW;
            float a[16];
            tmem_ld16(tOdS + lane_base + c0, a);
            if (gr < L) {
                float* dst = out_ds + base + (size_t)gr * d + c0;
#pragma unroll
                for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(a[i], a[i + 1], a[i + 2], a[i + 3]);
            }
            if (KEYSIDE) {
                tmem_ld16(tOP + lane_base + c0, a);
                if (gr < L) {
                    float* dst = out_p + base + (size_t)gr * d + c0;
#pragma unroll
                    for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(a[i], a[i + 1], a[i + 2], a[i + 3]);
                }
            }
        }
        fence_before_sync();
        __syncthreads();              // TMEM accumulators and the X / U tiles are reused by the next row tile
        fence_after_sync();
    }
    if (failed && tid == 0) *err = 1u;
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, kTmemCols);
}

}  // namespace rsb

using namespace rsb;

static int32_t make_drop(float p_drop, uint64_t drop_key, DropSpec* ds) {
    RSB_REQUIRE(p_drop >= 0.f && p_drop < 1.f, RSB200_EINVAL, "attention dropout probability must be in [0, 1), got %g", (double)p_drop);
    ds->key_lo = (uint32_t)drop_key; ds->key_hi = (uint32_t)(drop_key >> 32);
    const double t = (double)p_drop * 4294967296.0;
    ds->thresh = p_drop > 0.f ? (uint32_t)(t < 1.0 ? 1.0 : (t > 4294967295.0 ? 4294967295.0 : t)) : 0u;
    ds->keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
    return 0;
}

extern "C" int32_t rsb200_attn_bwd(const float* q, const float* k, const float* v, const float* o, const float* d_o,
                                   const float* lse, const int64_t* hist, int64_t B, int64_t L, int64_t heads,
                                   int64_t head_dim, int32_t causal, float p_drop, uint64_t drop_key, float* dq, float* dk, float* dv,
                                   uint32_t* err_flag, void* stream) {
    DropSpec drop;
    { int32_t rc = make_drop(p_drop, drop_key, &drop); if (rc) return rc; }
    RSB_REQUIRE(q && k && v && o && d_o && lse && dq && dk && dv && err_flag, RSB200_EINVAL, "null pointer");
    RSB_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o) && aligned16(d_o) && aligned16(dq) &&
                aligned16(dk) && aligned16(dv), RSB200_EINVAL, "pointers must be 16-byte aligned");
    RSB_REQUIRE(head_dim == kDH, RSB200_EUNSUPPORTED, "attention kernel is built for head_dim = 64 (got %lld)", (long long)head_dim);
    RSB_REQUIRE(L >= 1 && L <= kLP, RSB200_EUNSUPPORTED, "sequence length must be in [1, 256] (got %lld)", (long long)L);
    RSB_REQUIRE(B >= 1 && heads >= 1 && B * heads < ((int64_t)1 << 31), RSB200_EINVAL, "bad shape");
    const float scale = 1.0f / sqrtf((float)head_dim);
    cudaStream_t st = (cudaStream_t)stream;
    RSB_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem));
    RSB_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem));
    attn_bwd_kernel<false><<<(unsigned)(B * heads), kNT, kBwdSmem, st>>>(q, k, v, o, d_o, lse, hist, (int)L, (int)heads, causal, scale, drop, dq, nullptr, err_flag);
    RSB_LAUNCH_CHECK();
    attn_bwd_kernel<true><<<(unsigned)(B * heads), kNT, kBwdSmem, st>>>(q, k, v, o, d_o, lse, hist, (int)L, (int)heads, causal, scale, drop, dk, dv, err_flag);
    RSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int32_t rsb200_tc_gemm_test(const float* A, const float* B, float* D, int64_t N, int64_t K, uint32_t* err_flag,
                                       void* stream) {
    RSB_REQUIRE(A && B && D && err_flag, RSB200_EINVAL, "null pointer");
    RSB_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && K >= 16 && K <= 256 && K % 16 == 0, RSB200_EINVAL, "N, K must be multiples of 16 in [16, 256]");
    const size_t smem = (size_t)(128 + N) * K * 2;
    RSB_CUDA(cudaFuncSetAttribute(tc_gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_gemm_test_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, (int)N, (int)K, err_flag);
    RSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int32_t rsb200_attn_fwd(const float* q, const float* k, const float* v, const int64_t* hist, int64_t B, int64_t L,
                                   int64_t heads, int64_t head_dim, int32_t causal, float p_drop, uint64_t drop_key, float* out,
                                   float* lse, uint32_t* err_flag, void* stream) {
    DropSpec drop;
    { int32_t rc = make_drop(p_drop, drop_key, &drop); if (rc) return rc; }
    RSB_REQUIRE(q && k && v && out && err_flag, RSB200_EINVAL, "null pointer");
    RSB_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(out), RSB200_EINVAL, "pointers must be 16-byte aligned");
    RSB_REQUIRE(head_dim == kDH, RSB200_EUNSUPPORTED, "attention kernel is built for head_dim = 64 (got %lld)", (long long)head_dim);
    RSB_REQUIRE(L >= 1 && L <= kLP, RSB200_EUNSUPPORTED, "sequence length must be in [1, 256] (got %lld)", (long long)L);
    RSB_REQUIRE(B >= 1 && heads >= 1 && B * heads < ((int64_t)1 << 31), RSB200_EINVAL, "bad shape");
    RSB_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAttnSmem));
    attn_fwd_kernel<<<(unsigned)(B * heads), kNT, kAttnSmem, (cudaStream_t)stream>>>(
        q, k, v, hist, (int)L, (int)heads, causal, 1.0f / sqrtf((float)head_dim), drop, out, lse, err_flag);
    RSB_LAUNCH_CHECK();
    return 0;
}
