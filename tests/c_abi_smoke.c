/* Plain-C consumer of include/pqperm.h: proves the boundary is a C ABI (no C++
 * or Python types in the signatures).  Built and run by tests/test_host.py.
 * Without a GPU only the device-free paths are exercised. */
#include <stdio.h>
#include <string.h>

#include "pqperm.h"

int main(void)
{
    const double eye[2 * 4] = {1, 0, 0, 0, 0, 0, 1, 0};
    int32_t zeros[2] = {0, 0}, ones[2] = {1, 1}, bad[2] = {1, 0};
    double out[2] = {0, 0};
    pq_plan_info info;

    if (pq_perm_c128(eye, 2, 2, zeros, zeros, out) != PQ_OK || out[0] != 1.0 || out[1] != 0.0)
        return 1; /* src/permanent.cpp:106-108: empty problem -> 1 */
    if (pq_perm_c128(eye, 2, 2, ones, bad, out) != PQ_ERR_SUM_MISMATCH)
        return 2;
    if (strlen(pq_last_error()) == 0)
        return 3;
    if (pq_perm_plan(2, 2, ones, ones, &info) != PQ_OK || info.idx_max != 2 || info.sum_rows != 2)
        return 4;
    {
        int32_t g[2] = {-1, -1};
        if (pq_perm_gray_of_offset(2, ones, 1, g) != PQ_OK || g[0] != 0 || g[1] != 1)
            return 5;
    }
    if (pq_device_count() > 0) {
        if (pq_perm_c128(eye, 2, 2, ones, ones, out) != PQ_OK)
            return 6;
        if (out[0] < 0.999999999 || out[0] > 1.000000001 || out[1] != 0.0)
            return 7; /* perm(I_2) = 1 */
    } else if (pq_perm_c128(eye, 2, 2, ones, ones, out) != PQ_ERR_NO_DEVICE) {
        return 8; /* no CPU fallback */
    }
    printf("c abi ok, devices=%d\n", pq_device_count());
    return 0;
}
