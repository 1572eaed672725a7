/* delay_estimator_options.c -- the delay estimator's optional machinery from plain C: robust validation,
 * near-end lookahead, a soft reset and the estimate/quality query, on a small batch fed with an echo of known delay.
 *
 *   gcc -O2 -I include examples/delay_estimator_options.c -L webrtc_aecm_b200 -laecm_b200 \
 *       -Wl,-rpath,$PWD/webrtc_aecm_b200 -o delay_estimator_options && ./delay_estimator_options
 *
 * Needs a CUDA device (the library has no CPU path). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "aecm_b200.h"

enum { N = 8, SPC = 160, CALLS = 400 };

int main(void) {
    AecmB200Batch *b = AecmB200_Create(N, 0);
    AecmB200DelayEstimatorOptions opt = {1 /* robustValidation */, 2 /* allowedOffset */, 4 /* maxLookahead */, 1 /* lookahead */};
    AecmConfig cfg = {AecmTrue, 3};
    int16_t *far, *near, *out;
    int32_t status[N], delay[N], applied[N];
    float quality[N];
    uint32_t rng = 12345u;
    int c, i, k;
    if (!b) {
        fprintf(stderr, "AecmB200_Create failed: no CUDA device (this library has no CPU path)\n");
        return 1;
    }
    far = (int16_t *)AecmB200_HostAlloc((size_t)N * SPC * sizeof(int16_t));
    near = (int16_t *)AecmB200_HostAlloc((size_t)N * SPC * sizeof(int16_t));
    out = (int16_t *)AecmB200_HostAlloc((size_t)N * SPC * sizeof(int16_t));
#define CHECK(call)                                                                              \
    do {                                                                                         \
        const int32_t rc_ = (call);                                                              \
        if (rc_) {                                                                               \
            fprintf(stderr, "%s -> %d (%s)\n", #call, (int)rc_, AecmB200_LastError(b));         \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)
    if (!far || !near || !out) {
        fprintf(stderr, "AecmB200_HostAlloc failed\n");
        return 1;
    }
    CHECK(AecmB200_Init(b, 0, N, 16000));
    CHECK(AecmB200_SetConfig(b, 0, N, cfg, status));
    /* right after Init, as the reference collects the validation statistics from Init on */
    CHECK(AecmB200_SetDelayEstimatorOptions(b, 0, N, &opt, status));
    for (i = 0; i < N; i++)
        if (status[i]) {
            fprintf(stderr, "SetDelayEstimatorOptions: instance %d refused with %d\n", i, (int)status[i]);
            return 1;
        }
    {
        static int16_t hist[N][SPC * 64]; /* far-end history for the synthetic echo: instance i hears the far end i*3 ms late */
        memset(hist, 0, sizeof(hist));
        for (c = 0; c < CALLS; c++) {
            AecmB200StreamArgs a;
            for (i = 0; i < N; i++) {
                const int d = i * 48; /* samples */
                memmove(hist[i], hist[i] + SPC, sizeof(hist[i]) - SPC * sizeof(int16_t));
                for (k = 0; k < SPC; k++) {
                    /* bursty noise (a new level every 30 ms): the binary-spectrum estimator needs a far end that changes */
                    static const int level[8] = {0, 8, 1, 6, 0, 3, 8, 2};
                    rng = rng * 1664525u + 1013904223u;
                    hist[i][SPC * 63 + k] = (int16_t)(((int32_t)(rng >> 16) - 32768) / 16 * level[(c / 3 + c / 7) & 7] / 2);
                    far[i * SPC + k] = hist[i][SPC * 63 + k];
                    near[i * SPC + k] = (int16_t)(hist[i][SPC * 63 + k - d] / 2);
                }
            }
            memset(&a, 0, sizeof(a));
            a.farend = far; a.nearendNoisy = near; a.out = out;
            a.rowStride = SPC; a.count = N; a.nrOfSamples = SPC; a.nCalls = 1; a.msUniform = 40; a.status = status;
            CHECK(AecmB200_ProcessStream(b, &a));
            if (c == CALLS / 2) { /* the caller shifted its far-end stream by two blocks: tell the estimator */
                CHECK(AecmB200_SoftResetDelayEstimator(b, 0, N, 2, applied));
                printf("soft reset by 2 blocks: lookahead shift applied per instance = %d\n", applied[0]);
            }
        }
    }
    CHECK(AecmB200_GetDelayEstimate(b, 0, N, delay, quality));
    for (i = 0; i < N; i++) printf("instance %d: delay estimate %d blocks, quality %.3f\n", i, delay[i], quality[i]);
    AecmB200_HostFree(far); AecmB200_HostFree(near); AecmB200_HostFree(out);
    AecmB200_Free(b);
    return 0;
}
