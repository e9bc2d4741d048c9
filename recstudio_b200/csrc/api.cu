// api.cu -- C-ABI entry points of librsb200.so (see include/rsb200.h) and host-side glue.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"
#include "kernels.h"

namespace rsb {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static thread_local int cached_dev = -1, cached_sm = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached_dev = dev; cached_sm = v;
    }
    return cached_sm;
}

}  // namespace rsb

using namespace rsb;

extern "C" int32_t rsb200_version(void) { return RSB200_VERSION; }
extern "C" uint64_t rsb200_launch_count(void) { return g_launches.load(); }
extern "C" size_t rsb200_sizeof_pair_args(void) { return sizeof(rsb200_pair_args); }
extern "C" const char* rsb200_last_error(void) { return g_err; }

extern "C" int32_t rsb200_device_info(int32_t* sm, int32_t* max_threads_per_sm, int32_t* cc_major, int32_t* cc_minor) {
    int dev = 0, n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return RSB200_ENOCUDA;
    }
    RSB_CUDA(cudaGetDevice(&dev));
    int v = 0;
    if (sm) { RSB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm = v; }
    if (max_threads_per_sm) { RSB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxThreadsPerMultiProcessor, dev)); *max_threads_per_sm = v; }
    if (cc_major) { RSB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *cc_major = v; }
    if (cc_minor) { RSB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *cc_minor = v; }
    return 0;
}

// ------------------------------------------------------------------------------------------
extern "C" int32_t rsb200_pair_workspace_sizes(int64_t num_items, int64_t num_users, int64_t B, int64_t n, int64_t d,
                                               rsb200_pair_sizes* o) {
    RSB_REQUIRE(o != nullptr, RSB200_EINVAL, "null output");
    RSB_REQUIRE(num_items >= 1 && num_users >= 1 && B >= 0 && n >= 0 && d >= 4 && d % 4 == 0, RSB200_EINVAL,
                "bad problem shape (num_items=%lld num_users=%lld B=%lld n=%lld d=%lld)", (long long)num_items,
                (long long)num_users, (long long)B, (long long)n, (long long)d);
    RSB_REQUIRE(B * (n + 1) < ((int64_t)1 << 31), RSB200_EUNSUPPORTED, "B*(n+1) must be < 2^31");
    RSB_REQUIRE(num_items < ((int64_t)1 << 31) && num_users < ((int64_t)1 << 31), RSB200_EUNSUPPORTED, "table rows must be < 2^31");
    o->off_item = num_items + 1;
    o->off_user = num_users + 1;
    o->neg32_buf = B * n;
    o->slot_neg = B * n;
    o->slot_pos = B;
    o->slot_user = B;
    o->ent_item = B * (n + 1);
    o->ent_user = B;
    int64_t ci = B * (n + 1) < num_items ? B * (n + 1) : num_items;
    int64_t cu = B < num_users ? B : num_users;
    o->cap_item = ci > 0 ? ci : 1;
    o->cap_user = cu > 0 ? cu : 1;
    o->urow_item = o->cap_item;
    o->urow_user = o->cap_user;
    o->q_buf = B * d;
    o->dq_buf = B * d;
    o->loss_part = B;
    o->lse = B;
    int64_t a = scan_tmp_elems(num_items), b = scan_tmp_elems(num_users);
    o->scan_tmp = a > b ? a : b;
    // binned grouping (bins.cu)
    const int shift = bin_shift_for(num_items, B * (n + 1), B), shift_u = bin_shift_for(num_users, B, B);
    const bool ok = shift >= kMinBinShift && shift_u >= kMinBinShift && B * d < ((int64_t)1 << 30);
    o->bin_shift = ok ? shift : 0;
    o->bin_shift_user = ok ? shift_u : 0;
    o->nbins = ok ? cdiv(num_items, (int64_t)1 << shift) : 0;
    o->nbins_user = ok ? cdiv(num_users, (int64_t)1 << shift_u) : 0;
    o->bin_cnt = o->nbins + o->nbins_user;
    o->bin_off = o->bin_cnt + 2;
    o->bin_cursor = o->bin_cnt * kCursorStride;
    o->bin_status = o->bin_cnt;
    o->bin_heavy = ok ? bin_scatter_grid() * ((int64_t)1 << kMaxBinShift) : 0;
    return 0;
}

static int32_t check_pair(const rsb200_pair_args* a) {
    RSB_REQUIRE(a != nullptr, RSB200_EINVAL, "null args");
    RSB_REQUIRE(a->d >= 4 && a->d % 4 == 0 && a->d <= 512, a->d > 512 ? RSB200_EUNSUPPORTED : RSB200_EINVAL,
                "embedding dim must be a multiple of 4 in [4, 512], got %lld", (long long)a->d);
    RSB_REQUIRE(a->B >= 0 && a->n >= 0 && a->num_items >= 1 && a->num_users >= 1, RSB200_EINVAL, "bad sizes");
    RSB_REQUIRE(a->B * (a->n + 1) < ((int64_t)1 << 31), RSB200_EUNSUPPORTED, "B*(n+1) must be < 2^31");
    RSB_REQUIRE(a->num_items < ((int64_t)1 << 31) && a->num_users < ((int64_t)1 << 31), RSB200_EUNSUPPORTED, "table rows must be < 2^31");
    RSB_REQUIRE(a->w_item && a->w_user && a->user && a->pos, RSB200_EINVAL, "null table / batch pointer");
    RSB_REQUIRE(a->n == 0 || ((a->neg_i64 != nullptr) != (a->neg_i32 != nullptr)), RSB200_EINVAL,
                "exactly one of neg_i64 / neg_i32 must be given");
    RSB_REQUIRE(a->n == 0 || a->neg_i32 || a->neg32_buf, RSB200_EINVAL, "neg_i64 needs the neg32_buf workspace");
    RSB_REQUIRE(aligned16(a->w_item) && aligned16(a->w_user) && aligned16(a->q_buf) && aligned16(a->dq_buf) &&
                aligned16(a->item_vals) && aligned16(a->user_vals), RSB200_EINVAL, "tables / row buffers must be 16-byte aligned");
    RSB_REQUIRE(a->loss_kind == RSB200_LOSS_BPR || a->loss_kind == RSB200_LOSS_SSM, RSB200_EINVAL, "bad loss_kind");
    RSB_REQUIRE(a->score_kind == RSB200_SCORE_IP || a->score_kind == RSB200_SCORE_EUCLID, RSB200_EINVAL, "bad score_kind");
    RSB_REQUIRE(a->sink == RSB200_SINK_COMPACT || a->sink == RSB200_SINK_DENSE || a->sink == RSB200_SINK_APPLY, RSB200_EINVAL, "bad sink");
    if (a->sink == RSB200_SINK_APPLY) {
        RSB_REQUIRE(a->opt_kind >= 0 && a->opt_kind <= 2, RSB200_EINVAL, "opt_kind must be 0 (sgd), 1 (adagrad) or 2 (sparse_adam)");
        RSB_REQUIRE(a->w_item_rw && a->w_user_rw && aligned16(a->w_item_rw) && aligned16(a->w_user_rw), RSB200_EINVAL,
                    "SINK_APPLY needs writable tables (w_item_rw / w_user_rw)");
        RSB_REQUIRE(a->opt_kind == 0 || (a->item_state1 && a->user_state1 && aligned16(a->item_state1) && aligned16(a->user_state1)),
                    RSB200_EINVAL, "optimizer state1 missing");
        RSB_REQUIRE(a->opt_kind != 2 || (a->item_state2 && a->user_state2 && aligned16(a->item_state2) && aligned16(a->user_state2)),
                    RSB200_EINVAL, "optimizer state2 missing");
    }
    RSB_REQUIRE(a->grouping == 0 || a->grouping == 1, RSB200_EINVAL, "grouping must be 0 (counting sort) or 1 (bins)");
    RSB_REQUIRE(a->ent_item && a->ent_user && a->q_buf && a->dq_buf && a->loss_part && a->lse && a->err_flag && a->totals && a->loss,
                RSB200_EINVAL, "null workspace pointer");
    if (a->grouping == 0) {
        RSB_REQUIRE(a->off_item && a->slot_neg && a->slot_pos && a->urow_item && a->off_user && a->slot_user && a->urow_user &&
                    a->scan_tmp, RSB200_EINVAL, "null workspace pointer (grouping 0)");
    } else {
        RSB_REQUIRE(a->bin_shift >= kMinBinShift && a->bin_shift <= kMaxBinShift && a->bin_shift_user >= kMinBinShift &&
                    a->bin_shift_user <= kMaxBinShift, RSB200_EINVAL, "bin_shift / bin_shift_user must be in [%d, %d]", kMinBinShift, kMaxBinShift);
        RSB_REQUIRE(a->B <= ((int64_t)1 << (31 - (a->bin_shift > a->bin_shift_user ? a->bin_shift : a->bin_shift_user))), RSB200_EUNSUPPORTED,
                    "B = %lld does not fit the query bits of a binned entry", (long long)a->B);
        RSB_REQUIRE(a->bin_cnt && a->bin_off && a->bin_cursor && a->bin_status && a->bin_ticket && a->bin_heavy, RSB200_EINVAL,
                    "null bin workspace pointer (grouping 1)");
        RSB_REQUIRE(a->variant != 7 && a->variant != 6, RSB200_EUNSUPPORTED, "variants 6 / 7 belong to grouping 0");
        RSB_REQUIRE(a->B * a->d < ((int64_t)1 << 30), RSB200_EUNSUPPORTED, "B * d must be < 2^30 (32-bit byte offsets into the query matrix)");
    }
    return 0;
}

static void pair_bin_tables(const rsb200_pair_args* a, BinTable& ti, BinTable& tu) {
    ti = BinTable{}; tu = BinTable{};
    ti.nbins = (int)cdiv(a->num_items, (int64_t)1 << a->bin_shift); ti.shift = a->bin_shift; ti.num_rows = a->num_items;
    tu.nbins = (int)cdiv(a->num_users, (int64_t)1 << a->bin_shift_user); tu.shift = a->bin_shift_user; tu.num_rows = a->num_users;
    ti.cnt = a->bin_cnt; ti.off = a->bin_off; ti.cursor = a->bin_cursor; ti.status = a->bin_status; ti.ticket = a->bin_ticket;
    ti.totals = a->totals;
    tu.cnt = a->bin_cnt + ti.nbins; tu.off = a->bin_off + ti.nbins + 1; tu.cursor = a->bin_cursor + (size_t)ti.nbins * kCursorStride;
    tu.status = a->bin_status + ti.nbins; tu.ticket = a->bin_ticket + 1; tu.totals = a->totals + 2;
}

extern "C" int32_t rsb200_pair_draw_count(const rsb200_pair_args* a, uint64_t* state_dev, uint64_t seed, uint64_t philox_offset,
                                          int32_t sm_count_, int32_t max_threads_per_sm, int32_t* neg_out_i32, void* stream) {
    int32_t rc = check_pair(a);
    if (rc) return rc;
    RSB_REQUIRE(a->grouping == 1, RSB200_EUNSUPPORTED, "rsb200_pair_draw_count needs the binned grouping (grouping = 1)");
    RSB_REQUIRE(neg_out_i32 != nullptr && a->neg_i32 == neg_out_i32, RSB200_EINVAL, "neg_out_i32 must be the args' neg_i32 buffer");
    RSB_REQUIRE(sm_count_ > 0 && max_threads_per_sm >= 256 && philox_offset % 4 == 0, RSB200_EINVAL, "bad draw policy / philox offset");
    RSB_REQUIRE(a->num_items >= 2 && a->num_items - 1 < ((int64_t)1 << 28) && a->B * a->n * 8 < ((int64_t)1 << 31), RSB200_EUNSUPPORTED,
                "draw outside ATen's 32-bit path (see rsb200_sample_uniform)");
    BinTable ti, tu;
    pair_bin_tables(a, ti, tu);
    return launch_draw_bin_count(state_dev, seed, philox_offset, a->B, a->n, sm_count_, max_threads_per_sm, neg_out_i32, a->pos, a->B, ti,
                                 a->user, a->B, tu, a->err_flag, (cudaStream_t)stream);
}

extern "C" int32_t rsb200_pair_step(const rsb200_pair_args* a, int32_t phases, void* stream) {
    int32_t rc = check_pair(a);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t B = a->B, n = a->n;
    const int32_t* neg32 = a->neg_i32 ? a->neg_i32 : a->neg32_buf;

    const bool bins = a->grouping == 1;
    BinTable ti = {}, tu = {};
    if (bins) pair_bin_tables(a, ti, tu);
    if (phases & RSB200_PHASE_COUNT) {
        if (bins) {
            if (a->neg_i32) rc = launch_bin_count<int32_t>(a->neg_i32, B * n, a->pos, B, ti, a->user, B, tu, nullptr, a->err_flag, st);
            else            rc = launch_bin_count<int64_t>(a->neg_i64, B * n, a->pos, B, ti, a->user, B, tu, a->neg32_buf, a->err_flag, st);
            if (rc) return rc;
        } else {
            RSB_CUDA(cudaMemsetAsync(a->off_item, 0, sizeof(uint32_t) * (size_t)(a->num_items + 1), st));
            RSB_CUDA(cudaMemsetAsync(a->off_user, 0, sizeof(uint32_t) * (size_t)(a->num_users + 1), st));
            if (a->neg_i32) rc = launch_count<int32_t>(a->neg_i32, B * n, a->num_items, a->off_item, a->slot_neg, nullptr, a->err_flag, st);
            else            rc = launch_count<int64_t>(a->neg_i64, B * n, a->num_items, a->off_item, a->slot_neg, a->neg32_buf, a->err_flag, st);
            if (rc) return rc;
            rc = launch_count<int64_t>(a->pos, B, a->num_items, a->off_item, a->slot_pos, nullptr, a->err_flag, st);
            if (rc) return rc;
            rc = launch_count<int64_t>(a->user, B, a->num_users, a->off_user, a->slot_user, nullptr, a->err_flag, st);
            if (rc) return rc;
        }
    }
    if (phases & RSB200_PHASE_SCAN) {
        if (bins) {
            rc = launch_bin_scan(ti, tu, st);
            if (rc) return rc;
        } else {
            rc = launch_scan(a->off_item, a->num_items, a->urow_item, a->cap_item, a->totals, a->scan_tmp, a->scan_tmp_elems, st);
            if (rc) return rc;
            rc = launch_scan(a->off_user, a->num_users, a->urow_user, a->cap_user, a->totals + 2, a->scan_tmp, a->scan_tmp_elems, st);
            if (rc) return rc;
            if (a->variant != 6) {       // slot -> absolute entry position while the offsets are L2-resident (variant 6: legacy lookup in FWD)
                rc = launch_resolve(neg32, a->slot_neg, a->off_item, B * n, st);
                if (rc) return rc;
            }
        }
    }
    if (phases & RSB200_PHASE_FWD) {
        FwdParams p;
        p.w_item = a->w_item; p.w_user = a->w_user; p.user = a->user; p.pos = a->pos; p.neg = neg32;
        p.logq_pos = a->logq_pos; p.logq_neg = a->logq_neg;
        p.off_item = a->off_item; p.off_user = a->off_user;
        p.slot_neg = a->slot_neg; p.slot_pos = a->slot_pos; p.slot_user = a->slot_user;
        p.ent_item = a->ent_item; p.ent_user = a->ent_user;
        p.q_buf = a->q_buf; p.dq_buf = a->dq_buf; p.loss_part = a->loss_part; p.lse = a->lse;
        p.pos_score = a->pos_score; p.neg_score = a->neg_score;
        p.num_items = (int)a->num_items; p.num_users = (int)a->num_users; p.B = (int)B; p.n = (int)n; p.D = (int)a->d;
        const double denom = (a->loss_kind == RSB200_LOSS_BPR) ? (double)B * (double)(n > 0 ? n : 1) : (double)B;
        p.loss_scale = (float)(1.0 / (denom > 0 ? denom : 1.0));
        p.prefetch = (a->variant == 3) ? 1 : 0;
        p.slot_abs = (a->variant != 6) ? 1 : 0;
        // variant 7: staged entries + permute pass (measured: forward -0.08 ms, permute +0.10 ms -- not the default)
        p.cstage = (a->cstage && a->variant == 7) ? a->cstage : nullptr;
        p.hint = (a->variant >= 16 && a->variant < 32) ? (a->variant & 7) : 0;   // variants 16..31: L2 eviction hints
        if (a->variant == 32) p.hint = 8;    // timing diagnostic: skip the offset lookups / entry writes (gradients invalid)
        p.ncount = nullptr; p.sp_in = nullptr; p.stats_part = nullptr;
        p.bin_cursor = bins ? ti.cursor : nullptr; p.bin_shift = a->bin_shift; p.bin_bbits = 31 - a->bin_shift;
        p.bin_cursor_user = bins ? tu.cursor : nullptr; p.bin_shift_user = a->bin_shift_user; p.pad_row = 0;
        p.coef_scale = (float)((double)a->grad_scale / (denom > 0 ? denom : 1.0));
        rc = launch_pair_fwd(p, a->loss_kind, a->score_kind, a->variant, st);
        if (rc) return rc;
        rc = launch_loss_sum(a->loss_part, (int)B, a->loss, st);
        if (rc) return rc;
    }
    if (phases & RSB200_PHASE_SCATTER) {
        if (a->cstage && a->variant == 7) {
            rc = launch_permute_entries(a->slot_neg, a->cstage, B * n, n, a->loss_kind == RSB200_LOSS_BPR ? kDirect : 0u,
                                        a->ent_item, st);
            if (rc) return rc;
        }
        ScatterParams s;
        s.off = a->off_item; s.urow = a->urow_item; s.totals = a->totals; s.ent = a->ent_item; s.src = a->q_buf;
        s.lse = a->lse; s.w = a->w_item; s.gscale = a->grad_scale_dev; s.rows_out = a->item_rows; s.vals = a->item_vals; s.D = (int)a->d;
        s.cap = a->cap_item;
        s.ssm_scale = (float)((double)a->grad_scale / (double)(B > 0 ? B : 1));
        s.dense = a->sink == RSB200_SINK_DENSE; s.accumulate = a->accumulate; s.euclid = a->score_kind == RSB200_SCORE_EUCLID;
        s.hint = (a->variant >= 16 && a->variant < 32) ? ((a->variant >> 3) & 1) : 0;
        if (a->variant >= 40 && a->variant <= 44) s.hint = a->variant - 38;   // scatter occupancy / unroll experiments
        s.opt = -1; s.w_rw = nullptr; s.s1 = nullptr; s.s2 = nullptr; s.lr = s.b1 = s.b2 = s.eps = s.step_size = 0.f;
        if (a->sink == RSB200_SINK_APPLY) {
            s.opt = a->opt_kind; s.w_rw = a->w_item_rw; s.s1 = a->item_state1; s.s2 = a->item_state2;
            s.lr = a->opt_lr; s.b1 = a->opt_beta1; s.b2 = a->opt_beta2; s.eps = a->opt_eps; s.step_size = a->opt_step_size;
            s.dense = 0; s.accumulate = 0;
        }
        if (bins) {
            BinScatterParams b;
            b.ent = a->ent_item; b.bin_off = ti.off; b.status = ti.status; b.ticket = ti.ticket; b.totals = ti.totals;
            b.heavy_counts = a->bin_heavy; b.src = a->q_buf; b.lse = a->lse; b.w = a->w_item; b.gscale = a->grad_scale_dev;
            b.rows_out = a->item_rows; b.vals = a->item_vals; b.cap = a->cap_item;
            b.nbins = ti.nbins; b.shift = ti.shift; b.bbits = 31 - ti.shift; b.D = (int)a->d;
            b.ssm_scale = s.ssm_scale; b.dense = s.dense; b.accumulate = s.accumulate; b.euclid = s.euclid;
            b.opt = s.opt; b.w_rw = s.w_rw; b.s1 = s.s1; b.s2 = s.s2;
            b.lr = s.lr; b.b1 = s.b1; b.b2 = s.b2; b.eps = s.eps; b.step_size = s.step_size;
            b.tune = (a->variant >= 50 && a->variant < 60) ? a->variant - 50 : 0;      // variants 51..53: scatter depth / occupancy A/B
            rc = launch_bin_scatter(b, st);
            if (rc) return rc;
            // user table: every entry is (query b, 1.0) and the source rows are d loss / d query
            b.ent = a->ent_user; b.bin_off = tu.off; b.status = tu.status; b.ticket = tu.ticket; b.totals = tu.totals;
            b.src = a->dq_buf; b.lse = nullptr; b.w = a->w_user; b.rows_out = a->user_rows; b.vals = a->user_vals; b.cap = a->cap_user;
            b.nbins = tu.nbins; b.shift = tu.shift; b.bbits = 31 - tu.shift; b.euclid = 0;
            if (a->sink == RSB200_SINK_APPLY) { b.w_rw = a->w_user_rw; b.s1 = a->user_state1; b.s2 = a->user_state2; }
            return launch_bin_scatter(b, st);
        }
        rc = launch_scatter(s, a->cap_item, st);
        if (rc) return rc;
        ScatterParams u = s;
        u.off = a->off_user; u.urow = a->urow_user; u.totals = a->totals + 2; u.ent = a->ent_user; u.src = a->dq_buf;
        u.lse = nullptr; u.w = a->w_user; u.rows_out = a->user_rows; u.vals = a->user_vals; u.cap = a->cap_user;
        u.euclid = 0;
        if (a->sink == RSB200_SINK_APPLY) { u.w_rw = a->w_user_rw; u.s1 = a->user_state1; u.s2 = a->user_state2; }
        rc = launch_scatter(u, a->cap_user, st);
        if (rc) return rc;
    }
    return 0;
}
