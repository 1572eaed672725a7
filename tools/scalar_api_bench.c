/* Throughput of the reference's scalar API (WebRtcAecm_BufferFarend + WebRtcAecm_Process per 10 ms) when T host
 * threads drive T instances at the same time -- the way a media server built on the reference would.  Shows what the
 * call combiner in echo_control_mobile.c buys: concurrent calls share a launch.
 *   usage: scalar_api_bench [threads=32] [calls=500]      (links against libaecm_b200.so) */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "aecm_b200.h"

typedef struct { void *h; int calls; uint32_t seed; int failed; } Worker;
static pthread_barrier_t g_start;

static void *work(void *arg) {
    Worker *w = (Worker *)arg;
    int16_t far[160], near[160], out[160];
    uint32_t s = w->seed;
    int c, i;
    pthread_barrier_wait(&g_start);
    for (c = 0; c < w->calls; c++) {
        for (i = 0; i < 160; i++) {
            s = s * 1664525u + 1013904223u;
            far[i] = (int16_t)((int32_t)(s >> 16) % 4000 - 2000);
            near[i] = (int16_t)(far[i] / 2 + (int32_t)(s >> 8) % 64);
        }
        if (WebRtcAecm_BufferFarend(w->h, far, 160) || WebRtcAecm_Process(w->h, near, NULL, out, 160, 40)) { w->failed = 1; break; }
    }
    return NULL;
}

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

int main(int argc, char **argv) {
    const int T = argc > 1 ? atoi(argv[1]) : 32, calls = argc > 2 ? atoi(argv[2]) : 500;
    Worker *w = (Worker *)calloc((size_t)T, sizeof(Worker));
    pthread_t *th = (pthread_t *)calloc((size_t)T, sizeof(pthread_t));
    int64_t r0 = 0, q0 = 0, r1 = 0, q1 = 0;
    AecmConfig cfg = {AecmTrue, 3};
    double t0, dt;
    int i, bad = 0;
    if (T < 1 || T > 256) { fprintf(stderr, "threads must be 1..256 (one pool)\n"); return 2; }
    pthread_barrier_init(&g_start, NULL, (unsigned)T + 1);
    for (i = 0; i < T; i++) {
        w[i].h = WebRtcAecm_Create();
        w[i].calls = calls;
        w[i].seed = 12345u + 977u * (uint32_t)i;
        if (!w[i].h || WebRtcAecm_Init(w[i].h, 16000) || WebRtcAecm_set_config(w[i].h, cfg)) { fprintf(stderr, "setup failed\n"); return 1; }
    }
    AecmB200_ScalarApiStats(w[0].h, &r0, &q0);
    for (i = 0; i < T; i++) pthread_create(&th[i], NULL, work, &w[i]);
    pthread_barrier_wait(&g_start);
    t0 = now_s();
    for (i = 0; i < T; i++) { pthread_join(th[i], NULL); bad |= w[i].failed; }
    dt = now_s() - t0;
    AecmB200_ScalarApiStats(w[0].h, &r1, &q1);
    printf("{\"threads\": %d, \"calls_per_thread\": %d, \"seconds\": %.4f, \"ten_ms_calls_per_s\": %.0f, \"realtime_channels\": %.0f, "
           "\"requests\": %lld, \"rounds\": %lld, \"requests_per_round\": %.2f, \"us_per_round\": %.1f, \"failed\": %d}\n",
           T, calls, dt, (double)T * calls / dt, (double)T * calls / dt / 100.0, (long long)(q1 - q0), (long long)(r1 - r0),
           (double)(q1 - q0) / (double)(r1 - r0 ? r1 - r0 : 1), 1e6 * dt / (double)(r1 - r0 ? r1 - r0 : 1), bad);
    for (i = 0; i < T; i++) WebRtcAecm_Free(w[i].h);
    return bad;
}
