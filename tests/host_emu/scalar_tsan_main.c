/*
 * scalar_tsan_main.c -- TEST INFRASTRUCTURE.  Drives the product's scalar-API layer (csrc/echo_control_mobile.c: the ten
 * WebRtcAecm_* symbols, handle pools, request combiner) from many threads over the CPU stand-in for the device batch
 * (fake_batch.c), built with -fsanitize=thread by tests/test_scalar_combiner_tsan.py.  Every handle follows a script of
 * BufferFarend / Process calls with underruns, overruns, live set_config and echo-path calls; the expected output of
 * every call is computed beforehand by stepping the oracle directly through the same script, single-threaded.
 *
 *   scalar_tsan T H N   T threads, H handles per thread (T*H > 256 spans two pools), N 10 ms calls per handle
 *
 * Phase 1: all handles live, all threads calling at once (the combiner packs their requests into shared rounds).
 * Phase 2: create / use / free churn from all threads (pools are torn down when their last handle goes and rebuilt).
 * Exit code 0 = every output sample and return code as expected; ThreadSanitizer reports go to stderr.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "aecm_b200.h"
#include "echo_control_mobile.h"
#include "../../oracle/aecm_oracle.h"

void AecmB200_ScalarApiStats(void *inst, int64_t *rounds, int64_t *requests);
long fake_batches_created(void) __attribute__((weak)); /* only the stand-in batch (fake_batch.c) counts them */
long fake_batches_freed(void) __attribute__((weak));

typedef struct {
    void *h;
    int32_t (*init)(void *, int32_t);
    int32_t (*set_config)(void *, int cng, int mode);
    int32_t (*buffer_farend)(void *, const int16_t *, size_t);
    int32_t (*process)(void *, const int16_t *, const int16_t *, int16_t *, size_t, int16_t);
    int32_t (*init_echo_path)(void *, const void *, size_t);
    int32_t (*get_echo_path)(void *, void *, size_t);
} Api;

/* adapters: the oracle ... */
static int32_t o_init(void *h, int32_t fs) { return oracle_aecm_init((OracleAecm *)h, fs); }
static int32_t o_cfg(void *h, int c, int m) { return oracle_aecm_set_config((OracleAecm *)h, (int16_t)c, (int16_t)m); }
static int32_t o_far(void *h, const int16_t *f, size_t n) { return oracle_aecm_buffer_farend((OracleAecm *)h, f, n); }
static int32_t o_proc(void *h, const int16_t *x, const int16_t *c, int16_t *o, size_t n, int16_t ms) {
    return oracle_aecm_process((OracleAecm *)h, x, c, o, n, ms);
}
static int32_t o_ipath(void *h, const void *p, size_t n) { return oracle_aecm_init_echo_path((OracleAecm *)h, p, n); }
static int32_t o_gpath(void *h, void *p, size_t n) { return oracle_aecm_get_echo_path((OracleAecm *)h, p, n); }
/* ... and the product's scalar API */
static int32_t p_cfg(void *h, int c, int m) {
    AecmConfig cfg;
    cfg.cngMode = (int16_t)c;
    cfg.echoMode = (int16_t)m;
    return WebRtcAecm_set_config(h, cfg);
}

static uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

typedef struct {
    int id, fs, n, use_clean, n_calls;
    int16_t *far, *near_, *clean, *want, *got;
    int32_t *want_rc, *got_rc;
    int16_t want_path[65], got_path[65];
} Stream;

static void make_stream(Stream *s, int id, int n_calls) {
    const int total = n_calls * 160;
    s->id = id;
    s->fs = id % 3 == 0 ? 8000 : 16000;
    s->n = (id % 5 == 0) ? 240 - (s->fs == 8000 ? 80 : 160) : (s->fs == 8000 ? 80 : 160); /* sometimes the other frame size */
    s->use_clean = id % 4 == 1;
    s->n_calls = n_calls;
    s->far = (int16_t *)calloc((size_t)total, 2); s->near_ = (int16_t *)calloc((size_t)total, 2);
    s->clean = (int16_t *)calloc((size_t)total, 2); s->want = (int16_t *)calloc((size_t)total, 2);
    s->got = (int16_t *)calloc((size_t)total, 2);
    s->want_rc = (int32_t *)calloc((size_t)n_calls * 4, 4); s->got_rc = (int32_t *)calloc((size_t)n_calls * 4, 4);
    const int delay = 40 + (id * 37) % 700;
    for (int t = 0; t < total; t++) {
        const uint32_t r = mix((uint32_t)id * 0x9e3779b9u + (uint32_t)t);
        const int burst = (mix((uint32_t)id * 77u + (uint32_t)(t / 4000)) % 3) ? 6000 : 300;
        s->far[t] = (int16_t)(((int)(r & 0xffff) - 32768) * burst / 32768);
        const int echo = t >= delay ? s->far[t - delay] / 2 : 0;
        s->near_[t] = (int16_t)(echo + ((int)((r >> 16) & 0xff) - 128));
        s->clean[t] = (int16_t)(echo / 2 + ((int)((r >> 24) & 0x3f) - 32));
    }
}

/* the script of one 10 ms call `c` of stream s on `api`; appends up to four return codes at rc[] */
static void step(const Api *a, Stream *s, int c, int16_t *out, int32_t *rc, int16_t *path_out) {
    const int n = s->n;
    const int16_t *far = s->far + (size_t)c * 160, *near_ = s->near_ + (size_t)c * 160, *clean = s->clean + (size_t)c * 160;
    const int k = c + s->id;
    rc[0] = rc[1] = rc[2] = rc[3] = 0;
    if (k % 37 == 5) rc[0] = a->set_config(a->h, k % 2, (k / 37) % 5);
    if (k % 89 == 11) {
        int16_t p[65];
        for (int i = 0; i < 65; i++) p[i] = (int16_t)(mix((uint32_t)(k * 65 + i)) & 0x3fff);
        rc[0] |= a->init_echo_path(a->h, p, sizeof(p));
    }
    if (k % 53 != 7) rc[1] = a->buffer_farend(a->h, far, (size_t)n);          /* else: underrun */
    if (k % 41 == 3) rc[1] |= a->buffer_farend(a->h, far, (size_t)n);         /* overrun: the frame twice */
    rc[2] = a->process(a->h, near_, s->use_clean ? clean : NULL, out + (size_t)c * 160, (size_t)n, (int16_t)(20 + (s->id * 13) % 200));
    if (k % 101 == 17) rc[3] = a->get_echo_path(a->h, path_out, 130);
}

typedef struct {
    int tid, n_handles, n_calls, churn;
    Stream *streams;
    pthread_barrier_t *bar;
    int bad;
} Worker;

static void *worker(void *arg) {
    Worker *w = (Worker *)arg;
    Api *api = (Api *)calloc((size_t)w->n_handles, sizeof(Api));
    for (int i = 0; i < w->n_handles; i++) {
        Stream *s = &w->streams[i];
        api[i].h = WebRtcAecm_Create();
        api[i].init = WebRtcAecm_Init; api[i].set_config = p_cfg; api[i].buffer_farend = WebRtcAecm_BufferFarend;
        api[i].process = WebRtcAecm_Process; api[i].init_echo_path = WebRtcAecm_InitEchoPath; api[i].get_echo_path = WebRtcAecm_GetEchoPath;
        if (!api[i].h || WebRtcAecm_Init(api[i].h, s->fs) != 0) w->bad++;
    }
    pthread_barrier_wait(w->bar);
    for (int c = 0; c < w->n_calls; c++)
        for (int i = 0; i < w->n_handles; i++) step(&api[i], &w->streams[i], c, w->streams[i].got, w->streams[i].got_rc + 4 * c, w->streams[i].got_path);
    pthread_barrier_wait(w->bar);
    if (w->tid == 0) {
        int64_t rounds = 0, reqs = 0;
        AecmB200_ScalarApiStats(api[0].h, &rounds, &reqs);
        printf("pool of handle 0: %lld rounds, %lld requests (%.2f per round)\n", (long long)rounds, (long long)reqs,
               rounds ? (double)reqs / (double)rounds : 0.0);
    }
    for (int i = 0; i < w->n_handles; i++) WebRtcAecm_Free(api[i].h);
    /* phase 2: churn -- a few short-lived handles at a time, pools come and go underneath */
    for (int round = 0; round < w->churn; round++) {
        void *h[3];
        int16_t far[160] = {0}, near_[160], out[160];
        for (int j = 0; j < 3; j++) {
            h[j] = WebRtcAecm_Create();
            if (!h[j] || WebRtcAecm_Init(h[j], 16000) != 0) w->bad++;
        }
        for (int c = 0; c < 4; c++)
            for (int j = 0; j < 3; j++) {
                for (int t = 0; t < 160; t++) near_[t] = (int16_t)(mix((uint32_t)(w->tid * 1000 + round * 10 + c * 160 + t)) & 0x3ff);
                if (WebRtcAecm_BufferFarend(h[j], far, 160) != 0 || WebRtcAecm_Process(h[j], near_, NULL, out, 160, 40) != 0) w->bad++;
                if (c < 4 && memcmp(out, near_, sizeof(out)) != 0) w->bad++; /* start-up: the first calls pass the near end through */
            }
        for (int j = 0; j < 3; j++) WebRtcAecm_Free(h[j]);
    }
    free(api);
    return NULL;
}

int main(int argc, char **argv) {
    const int T = argc > 1 ? atoi(argv[1]) : 8, H = argc > 2 ? atoi(argv[2]) : 8, N = argc > 3 ? atoi(argv[3]) : 120;
    const int churn = argc > 4 ? atoi(argv[4]) : 20;
    Stream *streams = (Stream *)calloc((size_t)T * H, sizeof(Stream));
    int bad = 0;
    /* expected results: the oracle, stepped directly, one stream after the other */
    for (int i = 0; i < T * H; i++) {
        Stream *s = &streams[i];
        Api a = {oracle_aecm_create(), o_init, o_cfg, o_far, o_proc, o_ipath, o_gpath};
        make_stream(s, i, N);
        a.init(a.h, s->fs);
        for (int c = 0; c < N; c++) step(&a, s, c, s->want, s->want_rc + 4 * c, s->want_path);
        oracle_aecm_free((OracleAecm *)a.h);
    }
    pthread_t *th = (pthread_t *)calloc((size_t)T, sizeof(pthread_t));
    Worker *w = (Worker *)calloc((size_t)T, sizeof(Worker));
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, (unsigned)T);
    for (int t = 0; t < T; t++) {
        w[t].tid = t; w[t].n_handles = H; w[t].n_calls = N; w[t].churn = churn; w[t].streams = streams + (size_t)t * H; w[t].bar = &bar;
        pthread_create(&th[t], NULL, worker, &w[t]);
    }
    for (int t = 0; t < T; t++) {
        pthread_join(th[t], NULL);
        bad += w[t].bad;
    }
    for (int i = 0; i < T * H; i++) {
        Stream *s = &streams[i];
        if (memcmp(s->want, s->got, (size_t)N * 160 * 2) != 0) { printf("stream %d: output differs\n", i); bad++; }
        if (memcmp(s->want_rc, s->got_rc, (size_t)N * 16) != 0) { printf("stream %d: return codes differ\n", i); bad++; }
        if (memcmp(s->want_path, s->got_path, sizeof(s->want_path)) != 0) { printf("stream %d: exported echo path differs\n", i); bad++; }
    }
    printf("%d threads x %d handles x %d calls, churn %d: %d problems", T, H, N, churn, bad);
    if (fake_batches_created) printf("; device batches created %ld, freed %ld", fake_batches_created(), fake_batches_freed());
    printf("\n");
    return bad ? 1 : 0;
}
