// fullscore.cu -- full-catalog kernels (T1 top-k, L3 full softmax).  PLACEHOLDER: the
// entry points are exported so the ABI is complete, the kernels land in a later commit.
#include "common.cuh"
#include "kernels.h"

using namespace rsb;

extern "C" size_t rsb200_topk_workspace_bytes(int64_t, int64_t, int64_t) { return 0; }
extern "C" int32_t rsb200_topk_full(int32_t, const float*, const float*, int64_t, int64_t, int64_t, int64_t,
                                    const int64_t*, int64_t, float*, int64_t*, void*, size_t, void*) {
    set_error("rsb200_topk_full: not implemented yet");
    return RSB200_EUNSUPPORTED;
}
extern "C" size_t rsb200_fullsoftmax_workspace_bytes(int64_t, int64_t, int64_t) { return 0; }
extern "C" int32_t rsb200_fullsoftmax_fwd_bwd(const float*, const float*, const int64_t*, int64_t, int64_t, int64_t,
                                              float*, float*, float*, void*, size_t, void*) {
    set_error("rsb200_fullsoftmax_fwd_bwd: not implemented yet");
    return RSB200_EUNSUPPORTED;
}
