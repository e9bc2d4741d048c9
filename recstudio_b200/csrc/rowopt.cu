// rowopt.cu -- optimizer step on the touched rows only (the step right after the path:
// recommender.py:445-474,646).  The reference runs a dense optimizer over the whole [N, d] table
// (~19 GiB touched per step at config 2 for Adam); with sparse-row gradients only the R unique rows
// are read and written, and R comes from DEVICE memory (totals), so no host synchronisation is needed
// between the fused step and the update.
//   kind 0  SGD         : w -= lr * g                                        (torch.optim.SGD on a sparse grad)
//   kind 1  Adagrad     : s += g*g ; w -= lr * g / (sqrt(s) + eps)           (torch.optim.Adagrad, sparse path)
//   kind 2  SparseAdam  : m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g*g ;
//                         w -= lr * sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps)   (torch.optim.SparseAdam: moments
//                         of untouched rows do not decay -- NOT the reference's dense Adam)
// Row 0 (padding) never appears in `rows`.  One warp per row, 16-byte accesses.
#include "common.cuh"
#include "kernels.h"
#include "rowopt.cuh"

namespace rsb {

template <int KIND>
__global__ void __launch_bounds__(256)
rows_update_kernel(float* __restrict__ w, float* __restrict__ s1, float* __restrict__ s2, const int64_t* __restrict__ rows,
                   const float* __restrict__ vals, const uint32_t* __restrict__ count, int64_t cap, int D, float lr,
                   float b1, float b2, float eps, float step_size /* Adam: lr*sqrt(bc2)/bc1 */) {
    const int lane = threadIdx.x & 31;
    const int64_t R = min((int64_t)*count, cap);
    const int64_t warps = (int64_t)gridDim.x * 8;
    const OptParams o = {lr, b1, b2, eps, step_size};
    for (int64_t u = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); u < R; u += warps) {
        const int64_t r = rows[u];
        for (int c = lane * 4; c < D; c += 128)
            opt_update4<KIND>(w, s1, s2, (size_t)r * D + c, ldg128_stream(vals + (size_t)u * D + c), o);   // rowopt.cuh
    }
}

}  // namespace rsb

using namespace rsb;

extern "C" int32_t rsb200_rows_update(int32_t kind, float* w, float* state1, float* state2, int64_t num_rows, int64_t d,
                                      const int64_t* rows, const float* vals, const uint32_t* count_dev, int64_t cap,
                                      int64_t step, float lr, float beta1, float beta2, float eps, void* stream) {
    RSB_REQUIRE(w && rows && vals && count_dev, RSB200_EINVAL, "null pointer");
    RSB_REQUIRE(aligned16(w) && aligned16(vals) && d >= 4 && d % 4 == 0 && num_rows >= 1, RSB200_EINVAL, "bad table");
    RSB_REQUIRE(kind >= 0 && kind <= 2, RSB200_EINVAL, "kind must be 0 (sgd), 1 (adagrad) or 2 (sparse_adam)");
    RSB_REQUIRE(kind == 0 || (state1 && aligned16(state1)), RSB200_EINVAL, "optimizer state missing");
    RSB_REQUIRE(kind != 2 || (state2 && aligned16(state2) && step >= 1), RSB200_EINVAL, "sparse_adam needs state2 and step >= 1");
    if (cap <= 0) return 0;
    int64_t blocks = cdiv(cap, 8);
    const int64_t maxb = (int64_t)sm_count() * 8;
    if (blocks > maxb) blocks = maxb;
    cudaStream_t st = (cudaStream_t)stream;
    float step_size = lr;
    if (kind == 2) {
        const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
        step_size = (float)((double)lr * sqrt(bc2) / bc1);
    }
    if (kind == 0) rows_update_kernel<0><<<(unsigned)blocks, 256, 0, st>>>(w, state1, state2, rows, vals, count_dev, cap, (int)d, lr, beta1, beta2, eps, step_size);
    else if (kind == 1) rows_update_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(w, state1, state2, rows, vals, count_dev, cap, (int)d, lr, beta1, beta2, eps, step_size);
    else rows_update_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(w, state1, state2, rows, vals, count_dev, cap, (int)d, lr, beta1, beta2, eps, step_size);
    RSB_LAUNCH_CHECK();
    return 0;
}
