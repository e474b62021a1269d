/* Plain C client of the C ABI: proves include/cuspmm_b200.h is valid C (no C++ types in any signature)
 * and that a C program links against libcuspmm_b200.so.  Built and run by tests/test_abi.py (no GPU needed:
 * it only exercises argument validation and the error string). */
#include <stdio.h>
#include <string.h>

#include "cuspmm_b200.h"

int main(void) {
    int ok = 1;
    if (cuspmm_version() != CUSPMM_B200_VERSION) { printf("version mismatch\n"); ok = 0; }
    /* unknown variant: rejected before any CUDA call, like Engine::runKernel's "Not implemented" */
    int rc = cuspmm_spmm_csr(NULL, NULL, NULL, 4, 4, 0, NULL, 4, 4, NULL, 4, 99, NULL);
    if (rc != CUSPMM_ERR_INVALID || strstr(cuspmm_last_error(), "variant") == NULL) { printf("variant check failed\n"); ok = 0; }
    rc = cuspmm_spmm_bsr_f32(NULL, NULL, NULL, 1, 0, 0, 4, NULL, 4, 4, NULL, 4, NULL);
    if (rc != CUSPMM_ERR_INVALID) { printf("block shape check failed\n"); ok = 0; }
    if (cuspmm_spmm_coo_workspace(100, 1000, 64, 1) != 0 || cuspmm_spmm_coo_workspace(100, 1000, 64, 2) != 101 * 4) {
        printf("workspace query failed\n"); ok = 0;
    }
    cuspmmBsrTcPlan plan = NULL;
    rc = cuspmm_bsr_tc_plan_create(&plan, (const uint32_t *)&ok, NULL, NULL, 1, 0, 8, 8, 8, CUSPMM_BLK_BF16, NULL);
    if (rc != CUSPMM_ERR_UNSUPPORTED) { printf("tensor-core block size check failed (%d)\n", rc); ok = 0; }
    cuspmm_reset_launch_count();
    if (cuspmm_launch_count() != 0ULL) ok = 0;
    printf(ok ? "abi_c_client ok\n" : "abi_c_client FAILED\n");
    return ok ? 0 : 1;
}
