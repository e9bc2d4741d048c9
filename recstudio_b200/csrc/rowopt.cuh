// rowopt.cuh -- the per-element optimizer arithmetic shared by rows_update_kernel (rowopt.cu) and the fused
// scatter epilogue (scatter.cu, RSB200_SINK_APPLY).  Written with explicit round-to-nearest intrinsics so that
// the compiler cannot contract the two call sites differently: both paths produce bit-identical weights and
// state.  The operation order follows the torch optimizers they are tested against
// (torch/optim/sgd.py, adagrad.py (sparse path), sparse_adam.py).
#pragma once
#include "common.cuh"

namespace rsb {

struct OptParams {
    float lr, b1, b2, eps, step_size;   // step_size: SparseAdam lr * sqrt(1 - b2^t) / (1 - b1^t)
};

// KIND 0 SGD, 1 Adagrad (s1 = sum of squares), 2 SparseAdam (s1 = exp_avg, s2 = exp_avg_sq)
template <int KIND>
__device__ __forceinline__ float opt_elem(float x, float g, float& s1, float& s2, const OptParams& o) {
    if (KIND == 0) return __fmaf_rn(-o.lr, g, x);
    if (KIND == 1) {
        s1 = __fmaf_rn(g, g, s1);
        return __fmaf_rn(-o.lr, __fdiv_rn(g, __fadd_rn(__fsqrt_rn(s1), o.eps)), x);
    }
    s1 = __fadd_rn(__fmul_rn(o.b1, s1), __fmul_rn(1.f - o.b1, g));
    s2 = __fadd_rn(__fmul_rn(o.b2, s2), __fmul_rn(__fmul_rn(g, g), 1.f - o.b2));
    return __fmaf_rn(-o.step_size, __fdiv_rn(s1, __fadd_rn(__fsqrt_rn(s2), o.eps)), x);
}

// update of 4 consecutive elements at offset `off` of the table / state arrays
template <int KIND>
__device__ __forceinline__ void opt_update4(float* __restrict__ w, float* __restrict__ st1, float* __restrict__ st2, size_t off,
                                            float4 g, const OptParams& o) {
    float4 x = *reinterpret_cast<const float4*>(w + off);
    float4 a = make_float4(0, 0, 0, 0), b = make_float4(0, 0, 0, 0);
    if (KIND >= 1) a = *reinterpret_cast<const float4*>(st1 + off);
    if (KIND == 2) b = *reinterpret_cast<const float4*>(st2 + off);
    x.x = opt_elem<KIND>(x.x, g.x, a.x, b.x, o); x.y = opt_elem<KIND>(x.y, g.y, a.y, b.y, o);
    x.z = opt_elem<KIND>(x.z, g.z, a.z, b.z, o); x.w = opt_elem<KIND>(x.w, g.w, a.w, b.w, o);
    if (KIND >= 1) *reinterpret_cast<float4*>(st1 + off) = a;
    if (KIND == 2) *reinterpret_cast<float4*>(st2 + off) = b;
    *reinterpret_cast<float4*>(w + off) = x;
}

}  // namespace rsb
