/* C-only consumer of the rsb200 C ABI (include/rsb200.h): links librsb200.so with no Python / torch in the process.
 * Exercises the entry points that need no GPU -- version, size queries, argument validation, error strings -- which is
 * what a foreign-language binding (cgo / JNI / ctypes) relies on before it ever launches a kernel.
 * Built and run by tests/test_host_cpu.py::test_c_abi_from_plain_c. */
#include <stdio.h>
#include <string.h>
#include "rsb200.h"

#define CHECK(cond)                                                     \
    do {                                                                \
        if (!(cond)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); return 1; } \
    } while (0)

int main(void) {
    CHECK(rsb200_version() == RSB200_VERSION);

    rsb200_pair_sizes sz;
    memset(&sz, 0, sizeof sz);
    CHECK(rsb200_pair_workspace_sizes(10000001, 1000001, 8192, 1024, 128, &sz) == RSB200_OK);
    CHECK(sz.off_item == 10000002 && sz.ent_item == 8192LL * 1025 && sz.q_buf == 8192LL * 128);
    CHECK(sz.cap_item == 8192LL * 1025);                 /* min(B (n + 1), num_items) */
    CHECK(rsb200_pair_workspace_sizes(10, 10, 4, 3, 6, &sz) == RSB200_EINVAL);    /* d must be a multiple of 4 */
    CHECK(strstr(rsb200_last_error(), "bad problem shape") != NULL);
    CHECK(rsb200_pair_workspace_sizes(10, 10, 1 << 20, 1 << 12, 8, &sz) == RSB200_EUNSUPPORTED);   /* B (n + 1) >= 2^31 */

    /* null / malformed arguments are rejected before anything touches the device */
    CHECK(rsb200_pair_step(NULL, RSB200_PHASE_ALL, NULL) == RSB200_EINVAL);
    rsb200_pair_args a;
    memset(&a, 0, sizeof a);
    a.d = 6; a.B = 1; a.n = 1; a.num_items = 2; a.num_users = 2;
    CHECK(rsb200_pair_step(&a, RSB200_PHASE_ALL, NULL) == RSB200_EINVAL);
    a.d = 1024;
    CHECK(rsb200_pair_step(&a, RSB200_PHASE_ALL, NULL) == RSB200_EUNSUPPORTED);
    CHECK(rsb200_sample_uniform(0, 2, 10, 1, 1, 148, 2048, NULL, NULL, NULL) == RSB200_EINVAL);   /* philox offset % 4 */
    CHECK(rsb200_sample_uniform(0, 0, ((int64_t)1 << 28) + 2, 1, 1, 148, 2048, NULL, NULL, NULL) == RSB200_EUNSUPPORTED);
    CHECK(rsb200_philox_counter_offset(8192 * 1024, 148, 2048) == ((8192LL * 1024 - 1) / (256LL * 1184 * 4) + 1) * 4);
    CHECK(rsb200_topk_workspace_bytes(128, 1000001, 10, 64) > 0);
    CHECK(rsb200_index_workspace_bytes(1000000, 4096) > 4u * 4u * 1000000u);
    CHECK(rsb200_kmeans_assign(NULL, 64, 10, 64, NULL, 4, NULL, NULL, NULL, NULL) == RSB200_EINVAL);
    CHECK(rsb200_launch_count() == 0);                   /* nothing was launched */
    printf("abi ok\n");
    return 0;
}
