/*
 * alloc_failure_main.c -- TEST INFRASTRUCTURE.  Every device allocation the host layer makes is failed once, in turn
 * (host_model_fail_malloc_in), in AecmB200_Create, AecmB200_GroupCreate, WebRtcAecm_Create and inside AecmB200_SubmitStream's
 * on-demand staging growth; the process runs under AddressSanitizer with leak checking, so whatever a failure path forgets to
 * release is reported at exit (tests/test_host_orchestration_model.py).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "aecm_b200.h"
#include "echo_control_mobile.h"

void host_model_fail_malloc_in(int n);

int main(void) {
    int failed_create = 0, failed_group = 0, failed_scalar = 0, failed_submit = 0, k;
    int32_t devices[2] = {0, 1};
    for (k = 1; k < 60; k++) {
        AecmB200Batch *b;
        host_model_fail_malloc_in(k);
        b = AecmB200_Create(9, 0);
        host_model_fail_malloc_in(0);
        if (b) { AecmB200_Free(b); break; }
        failed_create++;
    }
    for (k = 1; k < 120; k++) {
        AecmB200Group *g;
        host_model_fail_malloc_in(k);
        g = AecmB200_GroupCreate(9, 2, devices);
        host_model_fail_malloc_in(0);
        if (g) { AecmB200_GroupFree(g); break; }
        failed_group++;
    }
    for (k = 1; k < 60; k++) {
        void *h;
        host_model_fail_malloc_in(k);
        h = WebRtcAecm_Create();
        host_model_fail_malloc_in(0);
        if (h) { WebRtcAecm_Free(h); break; }
        failed_scalar++;
    }
    {
        AecmB200Batch *b = AecmB200_Create(4, 0);
        int16_t far[4 * 160] = {0}, near_[4 * 160] = {0}, clean[4 * 160] = {0}, out[4 * 160];
        uint8_t mask[4] = {1, 1, 0, 1};
        int32_t status[4];
        AecmB200StreamArgs a;
        if (!b || AecmB200_Init(b, 0, 4, 16000) != 0) return 2;
        memset(&a, 0, sizeof(a));
        a.farend = far; a.nearendNoisy = near_; a.nearendClean = clean; a.out = out; a.rowStride = 160; a.count = 4; a.nrOfSamples = 160;
        a.nCalls = 1; a.msUniform = 40; a.farendMask = mask; a.status = status;
        for (k = 1; k < 12; k++) { /* staging planes, the clean plane and the mask buffer are allocated on first use */
            int32_t rc;
            host_model_fail_malloc_in(k);
            rc = AecmB200_SubmitStream(b, &a, 3);
            host_model_fail_malloc_in(0);
            if (rc == 0) {
                if (AecmB200_WaitStream(b, 3) != 0) return 3;
                break;
            }
            if (rc != 12000 || AecmB200_WaitStream(b, 3) != 0) return 4; /* AECM_UNSPECIFIED_ERROR, and the slot is free */
            failed_submit++;
        }
        if (AecmB200_ProcessStream(b, &a) != 0 || status[0] != 0) return 5; /* the batch is still usable */
        AecmB200_Free(b);
    }
    printf("allocation failures survived: Create %d, GroupCreate %d, WebRtcAecm_Create %d, SubmitStream %d\n", failed_create, failed_group,
           failed_scalar, failed_submit);
    return failed_create >= 10 && failed_group >= 20 && failed_scalar >= 10 && failed_submit >= 3 ? 0 : 1;
}
