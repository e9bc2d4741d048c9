// coalesce.cu -- sum gradient rows that target the same table row (sparse-row "coalesce").
//
// Used on the owner side of the row-sharded table (8(e)): gradient rows arrive from every rank
// as (local row id, 512-B value row) lists and must be accumulated per owned row -- the
// reference would do this with embedding_dense_backward into a dense [N,d] gradient
// (recommender.py:638).  Same machinery as the fused step: integer count -> scan -> CSR, then
// the segmented scatter kernel with src = the incoming value rows and coefficient 1.
#include "common.cuh"
#include "kernels.h"

namespace rsb {

__global__ void __launch_bounds__(256)
count_any_kernel(const int64_t* __restrict__ ids, int64_t M, int64_t num_rows, int skip_row0, uint32_t* __restrict__ cnt,
                 uint32_t* __restrict__ slot, uint32_t* __restrict__ err_flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int64_t id = ids[i];
    uint32_t s = kNoSlot;
    if (id >= (skip_row0 ? 1 : 0) && id < num_rows) s = atomicAdd(cnt + id, 1u);
    else if (!(skip_row0 && id == 0)) *err_flag = 1u;
    slot[i] = s;
}

__global__ void __launch_bounds__(256)
fill_entries_kernel(const int64_t* __restrict__ ids, int64_t M, const uint32_t* __restrict__ off,
                    const uint32_t* __restrict__ slot, uint64_t* __restrict__ ent) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const uint32_t s = slot[i];
    if (s == kNoSlot) return;
    ent[off[ids[i]] + s] = (uint64_t)((uint32_t)i | kDirect) | ((uint64_t)__float_as_uint(1.0f) << 32);
}

}  // namespace rsb

using namespace rsb;

extern "C" int32_t rsb200_rows_coalesce(const int64_t* ids, const float* vals, int64_t M, int64_t num_rows, int64_t d,
                                        int32_t skip_row0, int64_t* rows_out, float* vals_out, int32_t sink,
                                        int32_t accumulate, uint32_t* totals, uint32_t* off, uint32_t* slot,
                                        uint64_t* ent, uint32_t* urow, int64_t cap, uint64_t* scan_tmp,
                                        int64_t scan_tmp_elems, uint32_t* err_flag, void* stream) {
    RSB_REQUIRE(M >= 0 && M < ((int64_t)1 << 31) && num_rows >= 1 && num_rows < ((int64_t)1 << 31), RSB200_EINVAL, "bad sizes");
    RSB_REQUIRE(d >= 4 && d % 4 == 0 && d <= 512, RSB200_EINVAL, "d must be a multiple of 4 in [4, 512]");
    RSB_REQUIRE(rows_out && vals_out && totals && off && slot && ent && urow && scan_tmp && err_flag, RSB200_EINVAL, "null pointer");
    RSB_REQUIRE(M == 0 || (ids && vals && aligned16(vals)), RSB200_EINVAL, "null / misaligned input");
    RSB_REQUIRE(aligned16(vals_out), RSB200_EINVAL, "vals_out must be 16-byte aligned");
    RSB_REQUIRE(sink == RSB200_SINK_COMPACT || sink == RSB200_SINK_DENSE, RSB200_EINVAL, "bad sink");
    cudaStream_t st = (cudaStream_t)stream;
    RSB_CUDA(cudaMemsetAsync(off, 0, sizeof(uint32_t) * (size_t)(num_rows + 1), st));
    if (M > 0) {
        count_any_kernel<<<(unsigned)cdiv(M, 256), 256, 0, st>>>(ids, M, num_rows, skip_row0, off, slot, err_flag);
        RSB_LAUNCH_CHECK();
    }
    int32_t rc = launch_scan(off, num_rows, urow, cap, totals, scan_tmp, scan_tmp_elems, st);
    if (rc) return rc;
    if (M > 0) {
        fill_entries_kernel<<<(unsigned)cdiv(M, 256), 256, 0, st>>>(ids, M, off, slot, ent);
        RSB_LAUNCH_CHECK();
    }
    ScatterParams s;
    s.off = off; s.urow = urow; s.totals = totals; s.ent = ent; s.src = vals; s.lse = nullptr; s.w = nullptr; s.gscale = nullptr;
    s.rows_out = rows_out; s.vals = vals_out; s.cap = cap; s.D = (int)d; s.ssm_scale = 1.f;
    s.dense = sink == RSB200_SINK_DENSE; s.accumulate = accumulate; s.euclid = 0; s.hint = 0; s.opt = -1;
    return launch_scatter(s, cap, st);
}
