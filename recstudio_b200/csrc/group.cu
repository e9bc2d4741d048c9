// group.cu -- grouping of touched rows for the sparse-row gradient (E2).
//
// The reference accumulates embedding gradients with embedding_dense_backward
// (autograd of F.embedding, recommender.py:638): a dense [N,d] zero-fill plus
// scatter-add.  Here the touched rows are grouped instead (counting sort keyed
// by row id, N buckets) so that every gradient row is produced exactly once by
// one warp in scatter.cu with no atomics on fp32 data:
//   count : cnt[row]++ (integer atomics), slot[touch] = position inside the row
//   scan  : off[row] = exclusive prefix sum of cnt (CSR offsets, in place) and
//           urow[rank] = row for rows with cnt > 0 (ascending unique rows)
// Touches of the padding row 0 (and invalid ids) get slot = kNoSlot: they are
// scored but receive no gradient (nn.Embedding(padding_idx=0)).
#include "common.cuh"
#include "kernels.h"

namespace rsb {

// Every thread owns kCountPer elements a grid-stride apart (coalesced) and issues their atomics back to back, so four
// returning atomics (~1 us each) are in flight per thread instead of one: the kernel is latency-, not throughput-bound.
constexpr int kCountPer = 4;

template <typename IdT>
__global__ void __launch_bounds__(256)
count_kernel(const IdT* __restrict__ ids, int64_t M, int64_t num_rows, uint32_t* __restrict__ cnt,
             uint32_t* __restrict__ slot, int32_t* __restrict__ ids32_out, uint32_t* __restrict__ err_flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t id[kCountPer];
    uint32_t s[kCountPer];
#pragma unroll
    for (int k = 0; k < kCountPer; ++k) {
        const int64_t i = i0 + k * stride;
        id[k] = i < M ? (int64_t)ids[i] : 0;
    }
#pragma unroll
    for (int k = 0; k < kCountPer; ++k) {
        s[k] = kNoSlot;
        if (id[k] > 0 && id[k] < num_rows) s[k] = atomicAdd(cnt + id[k], 1u);
        else if (id[k] != 0) *err_flag = 1u;
    }
#pragma unroll
    for (int k = 0; k < kCountPer; ++k) {
        const int64_t i = i0 + k * stride;
        if (i < M) {
            slot[i] = s[k];
            if (ids32_out) ids32_out[i] = (id[k] >= 0 && id[k] < num_rows) ? (int32_t)id[k] : 0;   // i64 -> i32 copy for pair_fwd
        }
    }
}

// slot[i] (position inside the row's segment) -> absolute entry position off[id] + slot.  Runs right after the scan,
// while the freshly written offsets are still L2-resident, so that the forward kernel -- whose 4 GB row stream
// evicts them -- does not have to look them up with one random DRAM sector read per touch.
__global__ void __launch_bounds__(256)
resolve_kernel(const int32_t* __restrict__ ids, uint32_t* __restrict__ slot, const uint32_t* __restrict__ off, int64_t M) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s[kCountPer], o[kCountPer];
#pragma unroll
    for (int k = 0; k < kCountPer; ++k) {
        const int64_t i = i0 + k * stride;
        s[k] = i < M ? slot[i] : kNoSlot;
        o[k] = (s[k] != kNoSlot) ? __ldg(off + ids[i]) : 0u;
    }
#pragma unroll
    for (int k = 0; k < kCountPer; ++k) {
        const int64_t i = i0 + k * stride;
        if (s[k] != kNoSlot) slot[i] = o[k] + s[k];
    }
}

int32_t launch_resolve(const int32_t* ids, uint32_t* slot, const uint32_t* off, int64_t M, cudaStream_t st) {
    if (M == 0) return 0;
    resolve_kernel<<<(unsigned)cdiv(M, 256 * kCountPer), 256, 0, st>>>(ids, slot, off, M);
    RSB_LAUNCH_CHECK();
    return 0;
}

// ent[epos[i]] = (query of touch i | flag, staged value): the forward kernel leaves the per-touch coefficient / logit
// in touch order (a coalesced store inside its row stream); this pass moves them to their row-grouped positions.  Run on
// its own, the 67 MB entry list stays L2-resident, so the 8-byte random writes merge in L2 and reach DRAM as full lines --
// inside the forward kernel every one of them was a read-modify-write of a 32-byte sector.
__global__ void __launch_bounds__(256)
permute_entries_kernel(const uint32_t* __restrict__ epos, const float* __restrict__ cstage, int64_t M, int64_t n, uint32_t flag,
                       uint64_t* __restrict__ ent) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s[kCountPer]; float v[kCountPer];
#pragma unroll
    for (int k = 0; k < kCountPer; ++k) {
        const int64_t i = i0 + k * stride;
        s[k] = i < M ? __ldg(epos + i) : kNoSlot;
        v[k] = i < M ? __ldg(cstage + i) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < kCountPer; ++k) {
        const int64_t i = i0 + k * stride;
        if (s[k] != kNoSlot)
            ent[s[k]] = (uint64_t)((uint32_t)(i / n) | flag) | ((uint64_t)__float_as_uint(v[k]) << 32);
    }
}

int32_t launch_permute_entries(const uint32_t* epos, const float* cstage, int64_t M, int64_t n, uint32_t flag, uint64_t* ent,
                               cudaStream_t st) {
    if (M == 0) return 0;
    permute_entries_kernel<<<(unsigned)cdiv(M, 256 * kCountPer), 256, 0, st>>>(epos, cstage, M, n, flag, ent);
    RSB_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------- scan
// 3-phase scan over u32 counts.  Each block owns a tile of kTile elements; the packed
// u64 carries (sum of counts) in the low and (number of non-empty rows) in the high word.
constexpr int kScanThreads = 512;
constexpr int kScanPer = 8;
constexpr int kTile = kScanThreads * kScanPer;   // 4096

__device__ __forceinline__ uint64_t warp_incl_scan(uint64_t v) {
    int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t t = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across the block; returns the exclusive prefix and the block total
__device__ __forceinline__ uint64_t block_excl_scan(uint64_t v, uint64_t* total) {
    __shared__ uint64_t warp_sums[kScanThreads / 32];
    __shared__ uint64_t block_total;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint64_t inc = warp_incl_scan(v);
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint64_t s = (lane < kScanThreads / 32) ? warp_sums[lane] : 0;
        uint64_t si = warp_incl_scan(s);
        if (lane < kScanThreads / 32) warp_sums[lane] = si - s;
        if (lane == kScanThreads / 32 - 1) block_total = si;
    }
    __syncthreads();
    uint64_t excl = inc - v + warp_sums[w];
    *total = block_total;
    __syncthreads();   // shared arrays are reused by the next call
    return excl;
}

__global__ void __launch_bounds__(kScanThreads)
scan_reduce_kernel(const uint32_t* __restrict__ cnt, int64_t num_rows, uint64_t* __restrict__ block_sums) {
    int64_t base = (int64_t)blockIdx.x * kTile;
    uint64_t acc = 0;
#pragma unroll
    for (int k = 0; k < kScanPer; ++k) {
        int64_t i = base + (int64_t)k * kScanThreads + threadIdx.x;
        if (i < num_rows) {
            uint32_t c = cnt[i];
            acc += (uint64_t)c + ((uint64_t)(c != 0) << 32);
        }
    }
    uint64_t total;
    block_excl_scan(acc, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of the block sums in place; writes totals and off[num_rows]
__global__ void __launch_bounds__(kScanThreads)
scan_spine_kernel(uint64_t* __restrict__ block_sums, int64_t nblocks, uint32_t* __restrict__ off_last,
                  uint32_t* __restrict__ totals /* [2]: entries, unique rows */) {
    uint64_t carry = 0;
    for (int64_t base = 0; base < nblocks; base += kScanThreads) {
        int64_t i = base + threadIdx.x;
        uint64_t v = (i < nblocks) ? block_sums[i] : 0;
        uint64_t total;
        uint64_t ex = block_excl_scan(v, &total);
        if (i < nblocks) block_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) {
        *off_last = (uint32_t)carry;
        totals[0] = (uint32_t)carry;
        totals[1] = (uint32_t)(carry >> 32);
    }
}

// rescan each tile with its block offset; write offsets IN PLACE and the unique-row list
__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(uint32_t* __restrict__ cnt_off, int64_t num_rows, const uint64_t* __restrict__ block_sums,
                  uint32_t* __restrict__ urow, int64_t cap) {
    // thread t owns kScanPer CONSECUTIVE elements so that its local prefix is a serial sum
    int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kScanPer;
    static_assert(kScanPer == 8, "vector path loads two uint4");
    uint32_t c[kScanPer];
    uint64_t acc = 0;
    // cnt_off comes from a >=256-byte aligned allocation and base is a multiple of 8 elements
    const bool full = base + kScanPer <= num_rows && (reinterpret_cast<uintptr_t>(cnt_off) & 15u) == 0;
    if (full) {
        const uint4 a = *reinterpret_cast<const uint4*>(cnt_off + base);
        const uint4 b = *reinterpret_cast<const uint4*>(cnt_off + base + 4);
        c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < kScanPer; ++k) c[k] = (base + k < num_rows) ? cnt_off[base + k] : 0u;
    }
#pragma unroll
    for (int k = 0; k < kScanPer; ++k) acc += (uint64_t)c[k] + ((uint64_t)(c[k] != 0) << 32);
    uint64_t total;
    uint64_t ex = block_excl_scan(acc, &total) + block_sums[blockIdx.x];
    uint32_t o[kScanPer];
#pragma unroll
    for (int k = 0; k < kScanPer; ++k) {
        o[k] = (uint32_t)ex;
        if (c[k] != 0 && base + k < num_rows) {
            const int64_t rank = (int64_t)(ex >> 32);
            if (rank < cap) urow[rank] = (uint32_t)(base + k);
        }
        ex += (uint64_t)c[k] + ((uint64_t)(c[k] != 0) << 32);
    }
    if (full) {
        *reinterpret_cast<uint4*>(cnt_off + base) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(cnt_off + base + 4) = make_uint4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
        for (int k = 0; k < kScanPer; ++k)
            if (base + k < num_rows) cnt_off[base + k] = o[k];
    }
}

int64_t scan_tmp_elems(int64_t num_rows) { return cdiv(num_rows, kTile) + 1; }

template <typename IdT>
int32_t launch_count(const IdT* ids, int64_t M, int64_t num_rows, uint32_t* cnt, uint32_t* slot,
                     int32_t* ids32_out, uint32_t* err_flag, cudaStream_t st) {
    if (M == 0) return 0;
    count_kernel<IdT><<<(unsigned)cdiv(M, 256 * kCountPer), 256, 0, st>>>(ids, M, num_rows, cnt, slot, ids32_out, err_flag);
    RSB_LAUNCH_CHECK();
    return 0;
}
template int32_t launch_count<int32_t>(const int32_t*, int64_t, int64_t, uint32_t*, uint32_t*, int32_t*, uint32_t*, cudaStream_t);
template int32_t launch_count<int64_t>(const int64_t*, int64_t, int64_t, uint32_t*, uint32_t*, int32_t*, uint32_t*, cudaStream_t);

int32_t launch_scan(uint32_t* cnt_off, int64_t num_rows, uint32_t* urow, int64_t cap, uint32_t* totals,
                    uint64_t* tmp, int64_t tmp_elems, cudaStream_t st) {
    int64_t nblocks = cdiv(num_rows, kTile);
    RSB_REQUIRE(tmp_elems >= nblocks + 1, RSB200_EWORKSPACE, "scan_tmp too small: %lld < %lld",
                (long long)tmp_elems, (long long)(nblocks + 1));
    scan_reduce_kernel<<<(unsigned)nblocks, kScanThreads, 0, st>>>(cnt_off, num_rows, tmp);
    RSB_LAUNCH_CHECK();
    scan_spine_kernel<<<1, kScanThreads, 0, st>>>(tmp, nblocks, cnt_off + num_rows, totals);
    RSB_LAUNCH_CHECK();
    scan_apply_kernel<<<(unsigned)nblocks, kScanThreads, 0, st>>>(cnt_off, num_rows, tmp, urow, cap);
    RSB_LAUNCH_CHECK();
    return 0;
}

}  // namespace rsb
