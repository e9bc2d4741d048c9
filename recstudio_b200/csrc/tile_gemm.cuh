// tile_gemm.cuh -- fp32 FFMA tile-GEMM building blocks (128 x 128 tile, 256 threads, 8 x 8 per thread).
// Thread (tx = tid % 16, ty = tid / 16) owns output rows {tx*4+i, 64+tx*4+i} (i < 4) and columns
// {ty*8 + j} (j < 8): the row operand is read from shared memory with two conflict-free 16-byte loads,
// the column operand with two broadcast 16-byte loads.
#pragma once
#include "common.cuh"

namespace rsb {
namespace tg {

constexpr int TM = 128, TN = 128, TK = 16, TPAD = 4, LDS_ = TM + TPAD;

__device__ __forceinline__ int row_of(int tx, int i) { return (i < 4) ? tx * 4 + i : 64 + tx * 4 + (i - 4); }

__device__ __forceinline__ void zero_acc(float (&acc)[8][8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
}

// one k-slab (TK deep) of the micro-kernel: acc[i][j] += A[k][row_i] * B[k][col_j]
__device__ __forceinline__ void mma_slab(float (&acc)[8][8], const float (*As)[LDS_], const float (*Bs)[LDS_], int tx, int ty,
                                         int kmax = TK) {
#pragma unroll
    for (int k = 0; k < TK; ++k) {
        if (k < kmax) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][tx * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + tx * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][ty * 8]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][ty * 8 + 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
}

// C[m][n] += sum_k A[m0+m][k] * B[n0+n][k]   (A: [rowsA, K], B: [rowsB, K], both K-contiguous, K % 4 == 0)
// As / Bs: shared staging [2][TK][LDS_].  All 256 threads must call; ends with a __syncthreads().
__device__ __forceinline__ void gemm_nt(float (&acc)[8][8], const float* __restrict__ A, int rowsA, int m0,
                                        const float* __restrict__ B, int rowsB, int n0, int K,
                                        float (*As)[TK][LDS_], float (*Bs)[TK][LDS_], int lda = 0, int ldb = 0) {
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    if (lda == 0) lda = K;                 // row strides (a column slice of a wider matrix has ld > K)
    if (ldb == 0) ldb = K;
    auto gload = [&](const float* base, int rows, int ld, int r, int k0) -> float4 {
        if (r < rows && k0 + lk < K) return ldg128(base + (size_t)r * ld + k0 + lk);
        return make_float4(0, 0, 0, 0);
    };
    float4 ra0 = gload(A, rowsA, lda, m0 + lr, 0), ra1 = gload(A, rowsA, lda, m0 + lr + 64, 0);
    float4 rb0 = gload(B, rowsB, ldb, n0 + lr, 0), rb1 = gload(B, rowsB, ldb, n0 + lr + 64, 0);
    const int ksteps = (K + TK - 1) / TK;
    for (int ks = 0; ks < ksteps; ++ks) {
        const int buf = ks & 1;
        As[buf][lk + 0][lr] = ra0.x; As[buf][lk + 1][lr] = ra0.y; As[buf][lk + 2][lr] = ra0.z; As[buf][lk + 3][lr] = ra0.w;
        As[buf][lk + 0][lr + 64] = ra1.x; As[buf][lk + 1][lr + 64] = ra1.y; As[buf][lk + 2][lr + 64] = ra1.z; As[buf][lk + 3][lr + 64] = ra1.w;
        Bs[buf][lk + 0][lr] = rb0.x; Bs[buf][lk + 1][lr] = rb0.y; Bs[buf][lk + 2][lr] = rb0.z; Bs[buf][lk + 3][lr] = rb0.w;
        Bs[buf][lk + 0][lr + 64] = rb1.x; Bs[buf][lk + 1][lr + 64] = rb1.y; Bs[buf][lk + 2][lr + 64] = rb1.z; Bs[buf][lk + 3][lr + 64] = rb1.w;
        __syncthreads();
        if (ks + 1 < ksteps) {
            const int k0 = (ks + 1) * TK;
            ra0 = gload(A, rowsA, lda, m0 + lr, k0); ra1 = gload(A, rowsA, lda, m0 + lr + 64, k0);
            rb0 = gload(B, rowsB, ldb, n0 + lr, k0); rb1 = gload(B, rowsB, ldb, n0 + lr + 64, k0);
        }
        mma_slab(acc, As[buf], Bs[buf], tx, ty);
    }
    __syncthreads();      // staging buffers may be reused by the caller
}

// C[r][c] += sum_k Asm[k][r] * B[k0_row + k][c]   A already in shared memory, k-major: Asm[k][r], k < K (<= 128);
// B: global [rowsB, ldb] row-major, rows b_row0 + k, columns 0..TN-1 (c < ncols valid, ncols % 4 == 0).
// Bc: shared staging [2][TK][LDS_].
__device__ __forceinline__ void gemm_sa(float (&acc)[8][8], const float (*Asm)[LDS_], int K, const float* __restrict__ B,
                                        int rowsB, int b_row0, int ldb, int ncols, float (*Bc)[TK][LDS_]) {
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    // loader: 16 rows x 128 cols per slab = 512 float4 -> 2 per thread: row = (tid >> 5) + 8*h, col4 = (tid & 31) * 4
    const int rr = tid >> 5, cc = (tid & 31) * 4;
    auto gload = [&](int k) -> float4 {
        const int r = b_row0 + k;
        if (k < K && r < rowsB && cc < ncols) return ldg128(B + (size_t)r * ldb + cc);
        return make_float4(0, 0, 0, 0);
    };
    float4 r0 = gload(rr), r1 = gload(rr + 8);
    const int ksteps = (K + TK - 1) / TK;
    for (int ks = 0; ks < ksteps; ++ks) {
        const int buf = ks & 1;
        *reinterpret_cast<float4*>(&Bc[buf][rr][cc]) = r0;
        *reinterpret_cast<float4*>(&Bc[buf][rr + 8][cc]) = r1;
        __syncthreads();
        if (ks + 1 < ksteps) { r0 = gload((ks + 1) * TK + rr); r1 = gload((ks + 1) * TK + rr + 8); }
        mma_slab(acc, Asm + ks * TK, Bc[buf], tx, ty, min(TK, K - ks * TK));
    }
    __syncthreads();
}

}  // namespace tg
}  // namespace rsb
