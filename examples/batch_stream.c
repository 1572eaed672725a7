/* batch_stream.c -- the throughput path of INTEGRATION.md as a complete program.
 *
 * N AECM instances on one GPU, fed every 10 ms from pinned host planes through the pipelined
 * AecmB200_SubmitStream / AecmB200_WaitStream calls (four slots round-robin, far|near adjacent so that each tick
 * costs one host->device and one device->host transfer).  The audio here is synthetic; a real caller fills the planes
 * from its jitter buffers.  Per instance this is exactly the loop of the reference's main.cc:128-144.
 *
 *   gcc -O2 -I include examples/batch_stream.c -L webrtc_aecm_b200 -laecm_b200 -Wl,-rpath,$PWD/webrtc_aecm_b200 -o batch_stream
 *   ./batch_stream [instances=4096] [ticks=1000]
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "aecm_b200.h"

#define SPC 160 /* samples per 10 ms at 16 kHz */

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 4096, ticks = argc > 2 ? atoi(argv[2]) : 1000;
    const int slots = AecmB200_NumStreamSlots();
    const size_t plane = (size_t)n * SPC; /* samples per plane and tick */
    AecmB200Batch *b = AecmB200_Create(n, 0);
    AecmConfig cfg = {AecmTrue, 3};
    int16_t **in = (int16_t **)calloc((size_t)slots, sizeof(*in)), **out = (int16_t **)calloc((size_t)slots, sizeof(*out));
    int32_t **status = (int32_t **)calloc((size_t)slots, sizeof(*status));
    uint32_t rng = 1;
    long nonzero = 0;
    double t0;
    int s, k;
    if (!b) { fprintf(stderr, "no CUDA device: the engine has no CPU path\n"); return 1; }
    if (AecmB200_Init(b, 0, n, 16000) || AecmB200_SetConfig(b, 0, n, cfg, NULL)) { fprintf(stderr, "%s\n", AecmB200_LastError(b)); return 1; }
    for (s = 0; s < slots; s++) {
        in[s] = (int16_t *)AecmB200_HostAlloc(2 * plane * sizeof(int16_t)); /* [far plane | near plane] */
        out[s] = (int16_t *)AecmB200_HostAlloc(plane * sizeof(int16_t));
        status[s] = (int32_t *)calloc((size_t)n, sizeof(int32_t));
        if (!in[s] || !out[s] || !status[s]) { fprintf(stderr, "out of pinned memory\n"); return 1; }
    }
    t0 = now_s();
    for (k = 0; k < ticks + slots - 1; k++) {
        if (k >= slots - 1) { /* the oldest tick in flight: its results are needed now */
            const int done = (k - (slots - 1)) % slots;
            size_t i;
            if (AecmB200_WaitStream(b, done)) { fprintf(stderr, "%s\n", AecmB200_LastError(b)); return 1; }
            for (i = 0; i < (size_t)n; i++) nonzero += status[done][i] != 0;
            /* ... out[done] holds [n][160] echo-cancelled samples of that tick: hand them to the senders ... */
        }
        if (k < ticks) {
            AecmB200StreamArgs a;
            size_t i;
            s = k % slots;
            for (i = 0; i < plane; i++) { /* stand-in for real audio: far = noise, near = its echo */
                rng = rng * 1664525u + 1013904223u;
                in[s][i] = (int16_t)((int32_t)(rng >> 16) % 6000 - 3000);
                in[s][plane + i] = (int16_t)(in[s][i] / 3);
            }
            memset(&a, 0, sizeof(a));
            a.farend = in[s];
            a.nearendNoisy = in[s] + plane;
            a.out = out[s];
            a.rowStride = SPC;
            a.count = n;
            a.nrOfSamples = SPC;
            a.nCalls = 1;
            a.msUniform = 40;
            a.status = status[s];
            if (AecmB200_SubmitStream(b, &a, s)) { fprintf(stderr, "%s\n", AecmB200_LastError(b)); return 1; }
        }
    }
    printf("%d instances x %d ticks of 10 ms in %.3f s (%.0f x real time per instance), %ld non-zero statuses\n", n, ticks,
           now_s() - t0, ticks * 0.01 / (now_s() - t0), nonzero);
    for (s = 0; s < slots; s++) { AecmB200_HostFree(in[s]); AecmB200_HostFree(out[s]); free(status[s]); }
    free(in); free(out); free(status);
    AecmB200_Free(b);
    return 0;
}
