// sampler.cu -- S1 UniformSampler / S2 PopularSamplerModel draws, bit-identical to the
// reference's torch.randint / torch.rand + searchsorted on a CUDA device.
//   reference: recstudio/ann/sampler.py:86-114 (uniform), :243-258 (popularity)
//   ATen:      native/cuda/DistributionTemplates.h:50-89 (policy + grid-stride kernel),
//              :318-346 (random_from_to, 32-bit path), :485-506 (uniform_, (0,1] -> [0,1))
//
// ATen thread `idx` (of T = 256*grid threads) in round r draws curand4 once and hands word ii
// to element li = idx + T*(4r + ii).  Here one thread owns one (r, idx) pair, computes the
// same Philox block and writes its four elements; stores of a warp are 32 consecutive
// elements for every ii, i.e. fully coalesced.
#include "common.cuh"

namespace rsb {

struct DrawPolicy {
    int64_t T;        // 256 * grid of the ATen kernel
    int64_t rounds;   // grid-stride iterations per ATen thread
};

static DrawPolicy make_policy(int64_t numel, int32_t sm_count, int32_t max_threads_per_sm) {
    int64_t blocks = cdiv(numel, 256);
    int64_t grid = (int64_t)sm_count * (max_threads_per_sm / 256);
    if (blocks < grid) grid = blocks;
    DrawPolicy p;
    p.T = 256 * grid;
    p.rounds = (numel - 1) / (p.T * 4) + 1;
    return p;
}

__global__ void __launch_bounds__(256)
uniform_kernel(uint64_t seed, uint64_t ctr_base, int64_t T, int64_t rounds, int64_t numel,
               uint32_t range, int64_t* __restrict__ out64, int32_t* __restrict__ out32) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T * rounds) return;
    int64_t r = t / T, idx = t - r * T;
    uint4 w = Philox::gen(seed, (uint64_t)idx, ctr_base + (uint64_t)r);
    uint32_t words[4] = {w.x, w.y, w.z, w.w};
    int64_t li = r * 4 * T + idx;
#pragma unroll
    for (int ii = 0; ii < 4; ++ii, li += T) {
        if (li < numel) {
            uint32_t v = words[ii] % range + 1u;   // uniform_int_from_to<int64>(V, range, base=1)
            if (out64) out64[li] = (int64_t)v;
            if (out32) out32[li] = (int32_t)v;
        }
    }
}

// CUDA-graph-capturable form: (seed, philox offset) live in DEVICE memory, so a captured launch draws fresh ids on every
// replay; advance_state_kernel moves the offset on by what the draw consumed (the host does the same to torch's generator).
__global__ void __launch_bounds__(256)
uniform_dev_state_kernel(const uint64_t* __restrict__ state, int64_t T, int64_t rounds, int64_t numel, uint32_t range,
                         int64_t* __restrict__ out64, int32_t* __restrict__ out32) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T * rounds) return;
    const uint64_t seed = state[0], ctr_base = state[1] / 4;
    int64_t r = t / T, idx = t - r * T;
    uint4 w = Philox::gen(seed, (uint64_t)idx, ctr_base + (uint64_t)r);
    uint32_t words[4] = {w.x, w.y, w.z, w.w};
    int64_t li = r * 4 * T + idx;
#pragma unroll
    for (int ii = 0; ii < 4; ++ii, li += T) {
        if (li < numel) {
            uint32_t v = words[ii] % range + 1u;
            if (out64) out64[li] = (int64_t)v;
            if (out32) out32[li] = (int32_t)v;
        }
    }
}

__global__ void advance_state_kernel(uint64_t* state, uint64_t inc) { state[1] += inc; }

// first i in [lo, hi] with table[i] >= u  (hi is returned if none)
__device__ __forceinline__ int lower_bound(const float* __restrict__ table, int lo, int hi, float u) {
    while (lo < hi) {
        int mid = lo + ((hi - lo) >> 1);
        if (__ldg(table + mid) < u) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
popular_kernel(uint64_t seed, uint64_t ctr_base, int64_t T, int64_t rounds, int64_t numel,
               const float* __restrict__ table, const float* __restrict__ pop_prob, int num_items,
               const int32_t* __restrict__ guide, int guide_bits,
               int64_t* __restrict__ out64, int32_t* __restrict__ out32, float* __restrict__ logq) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T * rounds) return;
    int64_t r = t / T, idx = t - r * T;
    uint4 w = Philox::gen(seed, (uint64_t)idx, ctr_base + (uint64_t)r);
    uint32_t words[4] = {w.x, w.y, w.z, w.w};
    int64_t li = r * 4 * T + idx;
#pragma unroll
    for (int ii = 0; ii < 4; ++ii, li += T) {
        if (li < numel) {
            float u = curand_uniform_from_u32(words[ii]);   // (0, 1]
            u = u * 1.0f + 0.0f;                            // rand * range + from
            if (u == 1.0f) u = 0.0f;                        // reverse bounds -> [0, 1)
            int lo = 0, hi = num_items - 1;                 // searchsorted result N is clamped to N-1
            if (guide) {
                int k = (int)(u * (float)(1 << guide_bits));  // exact: power-of-two scale of a 24-bit float
                lo = guide[k]; hi = guide[k + 1];
            }
            // bisect down to a bracket of <= kLinear entries, then count the entries below u with INDEPENDENT loads:
            // every bisection step is a dependent random DRAM access (~1 us each at a 400 MB table), the final
            // bracket is one or two adjacent sectors.  Same result as a full bisection on a non-decreasing table.
            constexpr int kLinear = 8;
            while (hi - lo > kLinear) {
                const int mid = lo + ((hi - lo) >> 1);
                if (__ldg(table + mid) < u) lo = mid + 1; else hi = mid;
            }
            int id = lo;
#pragma unroll
            for (int t = 0; t < kLinear; ++t)
                if (lo + t < hi) id += (__ldg(table + lo + t) < u) ? 1 : 0;
            if (out64) out64[li] = id;
            if (out32) out32[li] = id;
            if (logq) logq[li] = logf(__ldg(pop_prob + id));
        }
    }
}

__global__ void __launch_bounds__(256)
guide_kernel(const float* __restrict__ table, int num_items, int guide_bits, int32_t* __restrict__ guide) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    int K = 1 << guide_bits;
    if (k > K) return;
    float u = (float)k / (float)K;            // exact
    guide[k] = lower_bound(table, 0, num_items - 1, u);
}

__global__ void __launch_bounds__(256)
logq_kernel(const float* __restrict__ pop_prob, int64_t num_items, const int64_t* __restrict__ ids, int64_t numel,
            float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numel) return;
    int64_t id = ids[i];
    if (id < 0 || id >= num_items) id = 0;
    out[i] = logf(__ldg(pop_prob + id));
}


// ------------------------------------------------------------------------------------------
// S3  MaskedUniformSampler (sampler.py:117-147,187-214): uniform over the items a user has NOT
// interacted with.  Reference arithmetic per user row b (H = history width, c_b = nonzero count):
//   neg   = floor(rand * float(num_items - c_b)) + 1                      (:133-136)
//   adj   = sort(hist_b) - max(i - (H - c_b), 0)                          (:137-141)
//   neg  += searchsorted(adj, neg, right=True) - (H - c_b)                (:142-144)
// masked_prep_kernel sorts one history row per CTA (bitonic, shared memory) and stores `adj`;
// masked_draw_kernel reproduces torch.rand's Philox element map and runs ATen's upper_bound loop
// (ATen/native/cuda/Bucketization.cu) over the FULL adjusted row, so results match the reference
// even for histories with duplicate items (where `adj` is not monotone).
constexpr int kMaskMaxHist = 4096;

__global__ void __launch_bounds__(256)
masked_prep_kernel(const int64_t* __restrict__ hist, int H, int P /* pow2 >= H */, int64_t* __restrict__ adj,
                   int32_t* __restrict__ cnt_out) {
    extern __shared__ long long sh_hist[];
    __shared__ int sh_cnt;
    const int b = blockIdx.x;
    if (threadIdx.x == 0) sh_cnt = 0;
    __syncthreads();
    int local = 0;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        long long v = (i < H) ? (long long)hist[(size_t)b * H + i] : 0x7FFFFFFFFFFFFFFFll;
        sh_hist[i] = v;
        local += (i < H && v != 0) ? 1 : 0;
    }
    if (local) atomicAdd(&sh_cnt, local);
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    long long a = sh_hist[i], c = sh_hist[ixj];
                    bool up = (i & k) == 0;
                    if ((a > c) == up) { sh_hist[i] = c; sh_hist[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    const int c = sh_cnt;
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
        int off = i - (H - c);
        adj[(size_t)b * H + i] = (int64_t)sh_hist[i] - (off > 0 ? off : 0);
    }
    if (threadIdx.x == 0) cnt_out[b] = c;
}

__global__ void __launch_bounds__(256)
masked_draw_kernel(uint64_t seed, uint64_t ctr_base, int64_t T, int64_t rounds, int64_t numel, int64_t per_user,
                   int64_t num_items /* real items, padding excluded (Sampler.num_items) */,
                   const int64_t* __restrict__ adj, const int32_t* __restrict__ cnt, int H,
                   int64_t* __restrict__ out64, int32_t* __restrict__ out32) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T * rounds) return;
    int64_t r = t / T, idx = t - r * T;
    uint4 w = Philox::gen(seed, (uint64_t)idx, ctr_base + (uint64_t)r);
    uint32_t words[4] = {w.x, w.y, w.z, w.w};
    int64_t li = r * 4 * T + idx;
#pragma unroll
    for (int ii = 0; ii < 4; ++ii, li += T) {
        if (li < numel) {
            float u = curand_uniform_from_u32(words[ii]);
            u = u * 1.0f + 0.0f;
            if (u == 1.0f) u = 0.0f;                                   // torch.rand: [0, 1)
            const int64_t b = li / per_user;
            const int c = __ldg(cnt + b);
            const float span = (float)(num_items - (int64_t)c);        // int64 -> fp32 promotion of the reference
            int64_t neg = (int64_t)floorf(__fmul_rn(u, span)) + 1;
            const int64_t* row = adj + (size_t)b * H;
            int start = 0, end = H;                                    // upper_bound: first adj[i] > neg
            while (start < end) {
                const int mid = start + ((end - start) >> 1);
                if (!(__ldg(row + mid) > neg)) start = mid + 1; else end = mid;
            }
            neg += (int64_t)start - (int64_t)(H - c);
            if (out64) out64[li] = neg;
            if (out32) out32[li] = (int32_t)neg;
        }
    }
}

}  // namespace rsb

using namespace rsb;

extern "C" int64_t rsb200_philox_counter_offset(int64_t numel, int32_t sm_count, int32_t max_threads_per_sm) {
    if (numel <= 0 || sm_count <= 0 || max_threads_per_sm < 256) return 0;
    DrawPolicy p = make_policy(numel, sm_count, max_threads_per_sm);
    return p.rounds * 4;
}

static int32_t check_draw(uint64_t off, int64_t num_items, int64_t nq, int64_t nn, int32_t sm, int32_t mt) {
    RSB_REQUIRE(nq >= 0 && nn >= 0, RSB200_EINVAL, "negative sampler shape");
    RSB_REQUIRE(num_items >= 2, RSB200_EINVAL, "num_items must include the padding row and at least one item");
    RSB_REQUIRE(sm > 0 && mt >= 256, RSB200_EINVAL, "bad device policy (sm_count=%d max_threads_per_sm=%d)", sm, mt);
    RSB_REQUIRE(off % 4 == 0, RSB200_EINVAL, "philox offset must be a multiple of 4 (torch invariant)");
    RSB_REQUIRE(nq * nn * 8 < ((int64_t)1 << 31), RSB200_EUNSUPPORTED,
                "numel*8 >= 2^31: torch splits such draws into 32-bit-indexable sub-iterators; not restated");
    RSB_REQUIRE(num_items < ((int64_t)1 << 31), RSB200_EUNSUPPORTED, "num_items >= 2^31");
    return 0;
}

extern "C" int32_t rsb200_sample_uniform(uint64_t seed, uint64_t philox_offset, int64_t num_items,
                                         int64_t num_queries, int64_t num_neg, int32_t sm_cnt, int32_t max_tpsm,
                                         int64_t* neg64, int32_t* neg32, void* stream) {
    int32_t rc = check_draw(philox_offset, num_items, num_queries, num_neg, sm_cnt, max_tpsm);
    if (rc) return rc;
    RSB_REQUIRE(num_items - 1 < ((int64_t)1 << 28), RSB200_EUNSUPPORTED,
                "range >= 2^28 takes ATen's 64-bit draw path (DistributionTemplates.h:319); not implemented");
    int64_t numel = num_queries * num_neg;
    if (numel == 0) return 0;
    DrawPolicy p = make_policy(numel, sm_cnt, max_tpsm);
    int64_t threads = p.T * p.rounds;
    uniform_kernel<<<(unsigned)cdiv(threads, 256), 256, 0, (cudaStream_t)stream>>>(
        seed, philox_offset / 4, p.T, p.rounds, numel, (uint32_t)(num_items - 1), neg64, neg32);
    RSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int32_t rsb200_popular_build_guide(const float* table, int64_t num_items, int32_t guide_bits,
                                              int32_t* guide_out, void* stream) {
    RSB_REQUIRE(table && guide_out, RSB200_EINVAL, "null pointer");
    RSB_REQUIRE(guide_bits >= 1 && guide_bits <= 24, RSB200_EINVAL, "guide_bits must be in [1, 24]");
    RSB_REQUIRE(num_items >= 1 && num_items < ((int64_t)1 << 31), RSB200_EINVAL, "bad num_items");
    int K = 1 << guide_bits;
    guide_kernel<<<(unsigned)cdiv(K + 1, 256), 256, 0, (cudaStream_t)stream>>>(table, (int)num_items, guide_bits, guide_out);
    RSB_LAUNCH_CHECK();
    return 0;
}

__global__ void __launch_bounds__(256)
guide_range_kernel(const float* __restrict__ table, int num_rows, int guide_bits, int64_t k0, int64_t len, int32_t* __restrict__ guide) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    const float u = (float)(k0 + i) / (float)(1 << guide_bits);            // exact (k <= 2^24)
    guide[i] = lower_bound(table, 0, num_rows - 1, u);
}

extern "C" int32_t rsb200_popular_build_guide_range(const float* table_local, int64_t num_rows, int32_t guide_bits, int64_t k0,
                                                    int64_t len, int32_t* guide_out, void* stream) {
    RSB_REQUIRE(table_local && guide_out, RSB200_EINVAL, "null pointer");
    RSB_REQUIRE(guide_bits >= 1 && guide_bits <= 24, RSB200_EINVAL, "guide_bits must be in [1, 24]");
    RSB_REQUIRE(num_rows >= 1 && num_rows < ((int64_t)1 << 31) && k0 >= 0 && len >= 1 && k0 + len - 1 <= ((int64_t)1 << guide_bits) + 1,
                RSB200_EINVAL, "bad slice / guide range");
    guide_range_kernel<<<(unsigned)cdiv(len, 256), 256, 0, (cudaStream_t)stream>>>(table_local, (int)num_rows, guide_bits, k0, len, guide_out);
    RSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int32_t rsb200_sample_popular(uint64_t seed, uint64_t philox_offset, const float* table,
                                         const float* pop_prob, int64_t num_items, int64_t num_queries,
                                         int64_t num_neg, int32_t sm_cnt, int32_t max_tpsm,
                                         const int32_t* guide, int32_t guide_bits,
                                         int64_t* neg64, int32_t* neg32, float* logq, void* stream) {
    int32_t rc = check_draw(philox_offset, num_items, num_queries, num_neg, sm_cnt, max_tpsm);
    if (rc) return rc;
    RSB_REQUIRE(table && (pop_prob || !logq), RSB200_EINVAL, "null table / pop_prob");
    RSB_REQUIRE(!guide || (guide_bits >= 1 && guide_bits <= 24), RSB200_EINVAL, "guide_bits must be in [1, 24]");
    int64_t numel = num_queries * num_neg;
    if (numel == 0) return 0;
    DrawPolicy p = make_policy(numel, sm_cnt, max_tpsm);
    int64_t threads = p.T * p.rounds;
    popular_kernel<<<(unsigned)cdiv(threads, 256), 256, 0, (cudaStream_t)stream>>>(
        seed, philox_offset / 4, p.T, p.rounds, numel, table, pop_prob, (int)num_items, guide, guide_bits,
        neg64, neg32, logq);
    RSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int32_t rsb200_popular_logq(const float* pop_prob, int64_t num_items, const int64_t* ids, int64_t numel,
                                       float* out, void* stream) {
    RSB_REQUIRE(pop_prob && ids && out, RSB200_EINVAL, "null pointer");
    if (numel == 0) return 0;
    logq_kernel<<<(unsigned)cdiv(numel, 256), 256, 0, (cudaStream_t)stream>>>(pop_prob, num_items, ids, numel, out);
    RSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int32_t rsb200_masked_workspace_elems(int64_t num_users, int64_t hist_len, int64_t* adj_elems, int64_t* cnt_elems) {
    RSB_REQUIRE(adj_elems && cnt_elems && num_users >= 0 && hist_len >= 0, RSB200_EINVAL, "bad arguments");
    *adj_elems = num_users * hist_len;
    *cnt_elems = num_users;
    return 0;
}

extern "C" int32_t rsb200_sample_uniform_masked(uint64_t seed, uint64_t philox_offset, int64_t num_items,
                                                const int64_t* user_hist, int64_t num_users, int64_t hist_len,
                                                int64_t per_user, int32_t sm_cnt, int32_t max_tpsm,
                                                int64_t* adj_ws, int32_t* cnt_ws,
                                                int64_t* neg64, int32_t* neg32, void* stream) {
    int32_t rc = check_draw(philox_offset, num_items, num_users, per_user, sm_cnt, max_tpsm);
    if (rc) return rc;
    RSB_REQUIRE(user_hist && adj_ws && cnt_ws, RSB200_EINVAL, "null history / workspace pointer");
    RSB_REQUIRE(hist_len >= 1 && hist_len <= kMaskMaxHist, RSB200_EUNSUPPORTED,
                "history width %lld outside [1, %d]", (long long)hist_len, kMaskMaxHist);
    int64_t numel = num_users * per_user;
    if (numel == 0) return 0;
    int P = 1;
    while (P < hist_len) P <<= 1;
    cudaStream_t st = (cudaStream_t)stream;
    masked_prep_kernel<<<(unsigned)num_users, 256, (size_t)P * sizeof(long long), st>>>(user_hist, (int)hist_len, P, adj_ws, cnt_ws);
    RSB_LAUNCH_CHECK();
    DrawPolicy p = make_policy(numel, sm_cnt, max_tpsm);
    int64_t threads = p.T * p.rounds;
    masked_draw_kernel<<<(unsigned)cdiv(threads, 256), 256, 0, st>>>(seed, philox_offset / 4, p.T, p.rounds, numel, per_user,
                                                                     num_items - 1, adj_ws, cnt_ws, (int)hist_len, neg64, neg32);
    RSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int32_t rsb200_sample_uniform_dev(uint64_t* state_dev, int64_t num_items, int64_t num_queries, int64_t num_neg,
                                             int32_t sm_cnt, int32_t max_tpsm, int64_t* neg64, int32_t* neg32, void* stream) {
    RSB_REQUIRE(state_dev != nullptr, RSB200_EINVAL, "null generator state");
    int32_t rc = check_draw(0, num_items, num_queries, num_neg, sm_cnt, max_tpsm);
    if (rc) return rc;
    RSB_REQUIRE(num_items - 1 < ((int64_t)1 << 28), RSB200_EUNSUPPORTED,
                "range >= 2^28 takes ATen's 64-bit draw path (DistributionTemplates.h:319); not implemented");
    int64_t numel = num_queries * num_neg;
    if (numel == 0) return 0;
    DrawPolicy p = make_policy(numel, sm_cnt, max_tpsm);
    int64_t threads = p.T * p.rounds;
    cudaStream_t st = (cudaStream_t)stream;
    uniform_dev_state_kernel<<<(unsigned)cdiv(threads, 256), 256, 0, st>>>(state_dev, p.T, p.rounds, numel,
                                                                           (uint32_t)(num_items - 1), neg64, neg32);
    RSB_LAUNCH_CHECK();
    advance_state_kernel<<<1, 1, 0, st>>>(state_dev, (uint64_t)(p.rounds * 4));
    RSB_LAUNCH_CHECK();
    return 0;
}
