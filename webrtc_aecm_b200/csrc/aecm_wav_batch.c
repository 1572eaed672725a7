/* aecm_wav_batch: the reference's command-line front-end (main.cc) as a multi-file batch.
 *
 * The reference CLI takes ONE far/near pair of mono WAV files, runs the near file through one AECM instance in 10 ms
 * calls (main.cc:97-147: cng on, echoMode 1, msInSndCardBuf 40) and writes "<near>_out.wav" (main.cc:172-196).  This
 * tool does the same for MANY pairs at once: every pair becomes one instance of a batch on the GPU, all instances
 * advance in lockstep, many 10 ms calls per kernel launch.  Every output file is byte-identical to what the
 * reference CLI writes for that pair (tests/test_gpu_wav_batch.py).
 *
 * Behaviour taken over from main.cc, per pair:
 *   - sample rate and length come from the NEAR file (main.cc:156, 166); 10 ms = rate/100 samples, at most 160;
 *   - only whole 10 ms frames are processed, the tail of the near file is written back untouched (main.cc:105,169);
 *   - a pair that is not mono is skipped without output (main.cc:46-51, 158-162);
 *   - a rate the canceller rejects (not 8000/16000) leaves the near signal unchanged but still writes it
 *     (main.cc:109-113: Init fails, aecProcess returns, AECM() writes the buffer anyway).
 * Deliberate difference: a far file shorter than the near file is padded with silence (the reference reads past the
 * end of its buffer).
 *
 * Ragged lengths: the pairs of one rate are sorted longest first, so the set of instances that still have frames
 * left is always a prefix [0, live) of the batch.  The batch advances in chunks of --chunk-calls 10 ms calls, one
 * kernel launch per chunk; a stream that ends inside a chunk runs on into silence for the rest of that chunk (its
 * output there is dropped) and is gone from the next one.  Chunks are staged through two sets of pinned planes, so the
 * host-side packing of chunk k+1, the PCIe copies and the kernel of chunk k overlap.  With several devices the sorted
 * pairs are dealt round-robin, one host thread per device
 * (AecmB200_SubmitStream / AecmB200_WaitStream).  No CPU fallback: without a CUDA device the tool fails. */
#define _GNU_SOURCE
#include <errno.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "aecm_b200.h"

typedef struct {
    char *far_path, *near_path, *out_path;
    int16_t *far, *near; /* near doubles as the output buffer, as in main.cc:141 */
    uint32_t rate;
    uint64_t n_far, n_near;
    int spc;          /* samples per 10 ms call, 0 = rate not supported */
    int64_t n_calls;  /* whole 10 ms frames in the near file */
    int skip;         /* unreadable or not mono: no output */
    int failed;       /* unreadable input or unwritable output */
    int unsupported;  /* rate the canceller rejects: written back unchanged */
    char msg[2][320]; /* what the reader / writer has to say about this pair (printed in input order) */
} Pair;

typedef struct {
    int echo_mode, cng, ms, chunk_calls, quiet, passthrough, n_devices, io_threads;
    int devices[64];
    const char *suffix;
} Options;

static double now_s(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

/* ------------------------------------------------------------------------------------------------------------------
 * WAV in / out.  RIFF little-endian; "fmt " may be plain PCM (1) or WAVE_FORMAT_EXTENSIBLE (0xFFFE) wrapping PCM.
 * Integer PCM of 8/16/24/32 bits is reduced to int16 the way dr_wav does it (top 16 bits; unsigned 8-bit recentred).
 * ---------------------------------------------------------------------------------------------------------------- */
static uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
static uint32_t rd16(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8; }

static int16_t *wav_read_mono16(const char *path, uint32_t *rate, uint64_t *count, int *channels, char *why, size_t why_len) {
    FILE *f = fopen(path, "rb");
    uint8_t hdr[12], ck[8], fmt[40];
    uint32_t tag = 0, ch = 0, sr = 0, bits = 0, align = 0;
    int have_fmt = 0;
    int16_t *out = NULL;
    *rate = 0; *count = 0; *channels = 0;
    if (!f) { snprintf(why, why_len, "%s", strerror(errno)); return NULL; }
    if (fread(hdr, 1, 12, f) != 12 || memcmp(hdr, "RIFF", 4) || memcmp(hdr + 8, "WAVE", 4)) {
        snprintf(why, why_len, "not a RIFF/WAVE file");
        goto fail;
    }
    while (fread(ck, 1, 8, f) == 8) {
        uint32_t size = rd32(ck + 4);
        if (!memcmp(ck, "fmt ", 4)) {
            uint32_t take = size < sizeof(fmt) ? size : (uint32_t)sizeof(fmt);
            if (size < 16 || fread(fmt, 1, take, f) != take) { snprintf(why, why_len, "short fmt chunk"); goto fail; }
            tag = rd16(fmt); ch = rd16(fmt + 2); sr = rd32(fmt + 4); align = rd16(fmt + 12); bits = rd16(fmt + 14);
            if (tag == 0xFFFEu && take >= 26) tag = rd16(fmt + 24); /* sub-format GUID starts with the real tag */
            have_fmt = 1;
            if (fseek(f, (long)(size - take) + (long)(size & 1u), SEEK_CUR)) break;
        } else if (!memcmp(ck, "data", 4)) {
            long pos = ftell(f), end;
            uint64_t bytes, frames, i;
            uint8_t *raw;
            uint32_t bps;
            if (!have_fmt) { snprintf(why, why_len, "data chunk before fmt chunk"); goto fail; }
            if (tag != 1) { snprintf(why, why_len, "format tag %u (only integer PCM is supported)", tag); goto fail; }
            if (bits != 8 && bits != 16 && bits != 24 && bits != 32) { snprintf(why, why_len, "%u bits per sample", bits); goto fail; }
            bps = bits / 8;
            if (ch == 0 || align != ch * bps) { snprintf(why, why_len, "inconsistent block align"); goto fail; }
            fseek(f, 0, SEEK_END);
            end = ftell(f);
            fseek(f, pos, SEEK_SET);
            bytes = size;
            if (bytes == 0 || bytes > (uint64_t)(end - pos)) bytes = (uint64_t)(end - pos); /* streamed or truncated file */
            frames = bytes / align;
            *channels = (int)ch;
            *rate = sr;
            if (ch != 1) { fclose(f); return NULL; } /* caller skips the pair (main.cc:46-51) */
            raw = (uint8_t *)malloc(frames * bps + 1);
            out = (int16_t *)malloc((frames + 1) * sizeof(int16_t));
            if (!raw || !out || fread(raw, bps, frames, f) != frames) {
                free(raw);
                snprintf(why, why_len, "short read");
                goto fail;
            }
            for (i = 0; i < frames; i++) {
                const uint8_t *p = raw + i * bps;
                switch (bps) {
                    case 1: out[i] = (int16_t)(((int)p[0] << 8) - 32768); break;
                    case 2: out[i] = (int16_t)rd16(p); break;
                    case 3: out[i] = (int16_t)rd16(p + 1); break;
                    default: out[i] = (int16_t)rd16(p + 2); break;
                }
            }
            free(raw);
            *count = frames;
            fclose(f);
            return out;
        } else if (fseek(f, (long)size + (long)(size & 1u), SEEK_CUR)) {
            break;
        }
    }
    snprintf(why, why_len, "no data chunk");
fail:
    free(out);
    fclose(f);
    return NULL;
}

static void wr32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

/* 44-byte canonical header, mono 16-bit PCM: what main.cc:21-37 produces through dr_wav */
static int wav_write_mono16(const char *path, const int16_t *s, uint64_t count, uint32_t rate) {
    uint8_t h[44];
    const uint64_t bytes = count * 2;
    FILE *f = fopen(path, "wb");
    size_t ok;
    if (!f) return -1;
    memcpy(h, "RIFF", 4); wr32(h + 4, (uint32_t)(36 + bytes)); memcpy(h + 8, "WAVEfmt ", 8); wr32(h + 16, 16);
    h[20] = 1; h[21] = 0; h[22] = 1; h[23] = 0; wr32(h + 24, rate); wr32(h + 28, rate * 2); h[32] = 2; h[33] = 0; h[34] = 16; h[35] = 0;
    memcpy(h + 36, "data", 4); wr32(h + 40, (uint32_t)bytes);
    ok = fwrite(h, 1, 44, f) == 44 && fwrite(s, 2, count, f) == count; /* hosts are little-endian (x86-64, aarch64) */
    return (fclose(f) == 0 && ok) ? 0 : -1;
}

/* "<dir><name>_out<ext>" of the near path (main.cc:189-190; the extension is the last dot of the last component) */
static char *derive_out_path(const char *near_path, const char *suffix) {
    const char *slash = strrchr(near_path, '/'), *base = slash ? slash + 1 : near_path, *dot = strrchr(base, '.');
    size_t stem = dot ? (size_t)(dot - near_path) : strlen(near_path);
    char *o = (char *)malloc(strlen(near_path) + strlen(suffix) + 1);
    memcpy(o, near_path, stem);
    strcpy(o + stem, suffix);
    strcat(o, near_path + stem);
    return o;
}

/* ------------------------------------------------------------------------------------------------------------------
 * one device's share of one rate group
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct {
    const Options *opt;
    Pair **pairs; /* sorted longest first */
    int n, device, spc;
    uint32_t rate;
    int64_t blocks_done;
    int launches;
    double init_s; /* CUDA context + batch + pinned staging: seconds on a cold GPU */
    char err[256];
    int failed;
} Shard;

/* wait for the submission in `slot`; any non-zero per-instance status ends the run (main.cc:135-139) */
static int collect(Shard *s, AecmB200Batch *b, int slot, const int32_t *status, int *pending) {
    const int count = *pending;
    int i, rc = AecmB200_WaitStream(b, slot);
    if (rc) { snprintf(s->err, sizeof(s->err), "WaitStream failed with %d: %s", rc, AecmB200_LastError(b)); return -1; }
    for (i = 0; i < count; i++)
        if (status[i]) { snprintf(s->err, sizeof(s->err), "%s: WebRtcAecm_Process returned %d", s->pairs[i]->near_path, status[i]); return -1; }
    return 0;
}

/* copy the chunk [t, t + calls) of the live rows [i0, i1) into (fill != 0) or out of the slot's pinned staging planes */
typedef struct {
    Shard *s;
    int i0, i1, calls, fill;
    int64_t t;
    int16_t *far, *near;
    const int16_t *out;
} StageJob;

static void *stage_rows(void *arg) {
    const StageJob *j = (const StageJob *)arg;
    const size_t spc = (size_t)j->s->spc, width = (size_t)j->calls * spc;
    int i;
    for (i = j->i0; i < j->i1; i++) {
        Pair *p = j->s->pairs[i];
        const int64_t left = p->n_calls - j->t; /* > 0 for a live row */
        const size_t have = (size_t)(left < j->calls ? left : j->calls) * spc, at = (size_t)j->t * spc;
        if (j->fill) {
            size_t have_far = p->n_far > at ? (size_t)(p->n_far - at) : 0; /* a short far file is continued with silence */
            if (have_far > have) have_far = have;
            memcpy(j->far + i * width, p->far + at, have_far * 2);
            memset(j->far + i * width + have_far, 0, (width - have_far) * 2);
            memcpy(j->near + i * width, p->near + at, have * 2);
            memset(j->near + i * width + have, 0, (width - have) * 2); /* a row that ends inside the chunk runs on into silence */
        } else {
            memcpy(p->near + at, j->out + i * width, have * 2); /* main.cc:141 */
        }
    }
    return NULL;
}

/* the rows are independent: a few threads move them (a chunk of 1 000 pairs is tens of megabytes each way) */
static void stage_chunk(Shard *s, int live, int64_t t, int calls, int16_t *far, int16_t *near, const int16_t *out, int fill) {
    StageJob job[16];
    pthread_t th[16];
    int T = s->opt->io_threads, k;
    if (T > 16) T = 16;
    if (T > live / 64 + 1) T = live / 64 + 1; /* not worth a thread below 64 rows */
    for (k = 0; k < T; k++) {
        job[k].s = s; job[k].calls = calls; job[k].fill = fill; job[k].t = t; job[k].far = far; job[k].near = near; job[k].out = out;
        job[k].i0 = (int)((int64_t)live * k / T);
        job[k].i1 = (int)((int64_t)live * (k + 1) / T);
    }
    for (k = 1; k < T; k++) pthread_create(&th[k], NULL, stage_rows, &job[k]);
    stage_rows(&job[0]);
    for (k = 1; k < T; k++) pthread_join(th[k], NULL);
}

static void *run_shard(void *arg) {
    Shard *s = (Shard *)arg;
    const Options *o = s->opt;
    const int n = s->n, spc = s->spc, K = o->chunk_calls;
    const int64_t max_calls = n ? s->pairs[0]->n_calls : 0;
    const size_t plane = (size_t)n * (size_t)K * (size_t)spc * sizeof(int16_t);
    AecmB200Batch *b = NULL;
    int16_t *far[2] = {NULL, NULL}, *near[2] = {NULL, NULL}, *out[2] = {NULL, NULL};
    int32_t *status[2] = {NULL, NULL};
    int64_t t_of[2] = {0, 0}, t = 0;
    AecmConfig cfg;
    int live = n, k = 0, i, rc, pending[2] = {0, 0}; /* pending[slot] = live rows of the chunk in flight there */
    const double t_start = now_s();
    if (n == 0 || max_calls == 0) return NULL;
    b = AecmB200_Create(n, s->device);
    if (!b) { snprintf(s->err, sizeof(s->err), "AecmB200_Create(%d instances, device %d) failed", n, s->device); s->failed = 1; return NULL; }
    for (i = 0; i < 2; i++) { /* pinned staging for two chunks in flight; far and near adjacent: one transfer for both */
        far[i] = (int16_t *)AecmB200_HostAlloc(2 * plane);
        out[i] = (int16_t *)AecmB200_HostAlloc(plane);
        status[i] = (int32_t *)calloc((size_t)n, sizeof(int32_t));
        if (!far[i] || !out[i] || !status[i]) { snprintf(s->err, sizeof(s->err), "out of pinned host memory (%zu bytes per plane)", plane); goto fail; }
    }
    cfg.cngMode = (int16_t)o->cng;
    cfg.echoMode = (int16_t)o->echo_mode;
    rc = AecmB200_Init(b, 0, n, (int32_t)s->rate);
    if (!rc) rc = AecmB200_SetConfig(b, 0, n, cfg, NULL);
    if (rc) { snprintf(s->err, sizeof(s->err), "init/set_config failed with %d: %s", rc, AecmB200_LastError(b)); goto fail; }
    s->init_s = now_s() - t_start;
    for (t = 0; t < max_calls; t += K, k++) {
        AecmB200StreamArgs a;
        const int slot = k & 1;
        while (live > 0 && s->pairs[live - 1]->n_calls <= t) live--; /* sorted longest first: the live rows are a prefix */
        if (pending[slot]) { /* results of the chunk before last */
            int count = pending[slot];
            if (collect(s, b, slot, status[slot], &count)) goto fail;
            stage_chunk(s, pending[slot], t_of[slot], K, NULL, NULL, out[slot], 0);
            pending[slot] = 0;
        }
        near[slot] = far[slot] + (size_t)live * (size_t)K * (size_t)spc;
        stage_chunk(s, live, t, K, far[slot], near[slot], NULL, 1);
        memset(&a, 0, sizeof(a));
        a.farend = far[slot];
        a.nearendNoisy = near[slot];
        a.out = out[slot];
        a.rowStride = (int64_t)K * spc;
        a.first = 0;
        a.count = live;
        a.nrOfSamples = spc;
        a.nCalls = K;
        a.msUniform = (int16_t)o->ms;
        a.status = status[slot];
        rc = AecmB200_SubmitStream(b, &a, slot);
        if (rc) { snprintf(s->err, sizeof(s->err), "SubmitStream failed with %d: %s", rc, AecmB200_LastError(b)); goto fail; }
        pending[slot] = live;
        t_of[slot] = t;
        s->launches++;
    }
    for (i = 0; i < 2; i++) {
        const int slot = (k + i) & 1; /* oldest first */
        int count = pending[slot];
        if (!count) continue;
        if (collect(s, b, slot, status[slot], &count)) goto fail;
        stage_chunk(s, pending[slot], t_of[slot], K, NULL, NULL, out[slot], 0);
        pending[slot] = 0;
    }
    for (i = 0; i < n; i++) s->blocks_done += s->pairs[i]->n_calls * spc / 64;
    goto done;
fail:
    s->failed = 1;
done:
    for (i = 0; i < 2; i++) { AecmB200_HostFree(far[i]); AecmB200_HostFree(out[i]); free(status[i]); }
    AecmB200_Free(b);
    return NULL;
}

/* ------------------------------------------------------------------------------------------------------------------
 * file I/O on a few threads: the pairs are independent, and with thousands of small files the open/read/close round
 * trips dominate the run otherwise
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct {
    Pair *pairs;
    int n, next, quiet, writing;
    pthread_mutex_t lock;
} IoJob;

static void read_pair(Pair *p, int quiet) {
    char why[128] = "";
    uint32_t far_rate;
    int ch_near, ch_far;
    p->near = wav_read_mono16(p->near_path, &p->rate, &p->n_near, &ch_near, why, sizeof(why));
    if (!p->near) {
        if (ch_near > 1) snprintf(p->msg[0], sizeof(p->msg[0]), "skipped (near has %d channels, mono only): %.200s\n", ch_near, p->near_path);
        else { snprintf(p->msg[0], sizeof(p->msg[0]), "error: %.200s: %s\n", p->near_path, why); p->failed = 1; }
        p->skip = 1;
        return;
    }
    p->far = wav_read_mono16(p->far_path, &far_rate, &p->n_far, &ch_far, why, sizeof(why));
    if (!p->far) {
        if (ch_far > 1) snprintf(p->msg[0], sizeof(p->msg[0]), "skipped (far has %d channels, mono only): %.200s\n", ch_far, p->far_path);
        else { snprintf(p->msg[0], sizeof(p->msg[0]), "error: %.200s: %s\n", p->far_path, why); p->failed = 1; }
        p->skip = 1;
        return;
    }
    if ((p->rate == 8000 || p->rate == 16000) && p->n_near > 0) { /* WebRtcAecm_Init, echo_control_mobile.cc:150-152 */
        p->spc = (int)(p->rate / 100);
        p->n_calls = (int64_t)(p->n_near / (uint64_t)p->spc);
    } else {
        p->unsupported = 1;
        if (!quiet) snprintf(p->msg[0], sizeof(p->msg[0]), "not processed (rate %u Hz): %.200s is written back unchanged\n", p->rate, p->near_path);
    }
    if (p->n_far < (uint64_t)p->n_calls * (uint64_t)p->spc && !quiet)
        snprintf(p->msg[0], sizeof(p->msg[0]), "note: %.200s is shorter than its near file; padded with silence\n", p->far_path);
}

static void *io_worker(void *arg) {
    IoJob *j = (IoJob *)arg;
    for (;;) {
        int i;
        Pair *p;
        pthread_mutex_lock(&j->lock);
        i = j->next++;
        pthread_mutex_unlock(&j->lock);
        if (i >= j->n) return NULL;
        p = &j->pairs[i];
        if (!j->writing) {
            read_pair(p, j->quiet);
        } else if (!p->skip && wav_write_mono16(p->out_path, p->near, p->n_near, p->rate)) {
            snprintf(p->msg[1], sizeof(p->msg[1]), "error: cannot write %.200s\n", p->out_path);
            p->failed = 1;
        }
    }
}

static void run_io(Pair *pairs, int n, int threads, int quiet, int writing) {
    IoJob j;
    pthread_t th[64];
    int t;
    j.pairs = pairs; j.n = n; j.next = 0; j.quiet = quiet; j.writing = writing;
    pthread_mutex_init(&j.lock, NULL);
    if (threads > n) threads = n;
    if (threads > 64) threads = 64;
    for (t = 1; t < threads; t++) pthread_create(&th[t], NULL, io_worker, &j);
    io_worker(&j);
    for (t = 1; t < threads; t++) pthread_join(th[t], NULL);
    pthread_mutex_destroy(&j.lock);
    for (t = 0; t < n; t++)
        if (pairs[t].msg[writing][0]) fputs(pairs[t].msg[writing], stderr); /* in input order */
}

static int by_calls_desc(const void *x, const void *y) {
    const Pair *a = *(Pair *const *)x, *b = *(Pair *const *)y;
    if (a->n_calls != b->n_calls) return a->n_calls < b->n_calls ? 1 : -1;
    return a < b ? -1 : (a > b);
}

static void usage(void) {
    fprintf(stderr,
            "usage: aecm_wav_batch [options] far1.wav near1.wav [far2.wav near2.wav ...]\n"
            "  writes <near>_out.wav next to every near file, exactly what the reference CLI writes per pair\n"
            "  --list FILE        additional pairs, one per line: far near [out]\n"
            "  --echo-mode M      0..4 (default 1, as main.cc)      --cng 0|1   comfort noise (default 1)\n"
            "  --ms MS            msInSndCardBuf for every call (default 40, as main.cc)\n"
            "  --gpus N | --devices a,b,..   CUDA devices to spread the pairs over (default: device 0)\n"
            "  --chunk-calls K    10 ms calls per kernel launch (default 100)\n"
            "  --suffix S         output suffix (default _out)     -q  quiet     --io-threads T  file readers/writers (default 8)\n"
            "  --passthrough      I/O self-test: read, do not process, write (no GPU needed)\n");
}

static void add_pair(Pair **v, int *n, int *cap, const char *far, const char *near, const char *out, const char *suffix) {
    Pair *p;
    if (*n == *cap) { *cap = *cap ? *cap * 2 : 64; *v = (Pair *)realloc(*v, (size_t)*cap * sizeof(Pair)); }
    p = &(*v)[(*n)++];
    memset(p, 0, sizeof(*p));
    p->far_path = strdup(far);
    p->near_path = strdup(near);
    p->out_path = out ? strdup(out) : derive_out_path(near, suffix);
}

int main(int argc, char **argv) {
    Options o;
    Pair *pairs = NULL;
    int n = 0, cap = 0, i, d, bad = 0, read_bad = 0, proc_bad = 0, written = 0, skipped = 0, unsupported = 0;
    const char *lists[64];
    int n_lists = 0;
    double t0, t_read, t_proc, t_write, t_init = 0;
    int64_t blocks = 0;
    int launches = 0;
    double audio_s = 0;
    memset(&o, 0, sizeof(o));
    o.echo_mode = 1; o.cng = 1; o.ms = 40; o.chunk_calls = 100; o.suffix = "_out"; o.n_devices = 1; o.devices[0] = 0; o.io_threads = 8;
    for (i = 1; i < argc; i++) {
        const char *a = argv[i];
        if (!strcmp(a, "--echo-mode") && i + 1 < argc) o.echo_mode = atoi(argv[++i]);
        else if (!strcmp(a, "--cng") && i + 1 < argc) o.cng = atoi(argv[++i]);
        else if (!strcmp(a, "--ms") && i + 1 < argc) o.ms = atoi(argv[++i]);
        else if (!strcmp(a, "--chunk-calls") && i + 1 < argc) o.chunk_calls = atoi(argv[++i]);
        else if (!strcmp(a, "--suffix") && i + 1 < argc) o.suffix = argv[++i];
        else if (!strcmp(a, "--io-threads") && i + 1 < argc) o.io_threads = atoi(argv[++i]);
        else if (!strcmp(a, "--list") && i + 1 < argc && n_lists < 64) lists[n_lists++] = argv[++i];
        else if (!strcmp(a, "--gpus") && i + 1 < argc) {
            o.n_devices = atoi(argv[++i]);
            if (o.n_devices < 1 || o.n_devices > 64) { usage(); return 2; }
            for (d = 0; d < o.n_devices; d++) o.devices[d] = d;
        } else if (!strcmp(a, "--devices") && i + 1 < argc) {
            char *s = strdup(argv[++i]), *tok;
            o.n_devices = 0;
            for (tok = strtok(s, ","); tok && o.n_devices < 64; tok = strtok(NULL, ",")) o.devices[o.n_devices++] = atoi(tok);
            free(s);
            if (o.n_devices < 1) { usage(); return 2; }
        } else if (!strcmp(a, "-q")) o.quiet = 1;
        else if (!strcmp(a, "--passthrough")) o.passthrough = 1;
        else if (!strcmp(a, "-h") || !strcmp(a, "--help")) { usage(); return 0; }
        else if (a[0] == '-' && a[1] == '-') { fprintf(stderr, "unknown option %s\n", a); usage(); return 2; }
        else if (i + 1 < argc) { add_pair(&pairs, &n, &cap, argv[i], argv[i + 1], NULL, o.suffix); i++; }
        else { fprintf(stderr, "far file %s has no near file\n", a); return 2; }
    }
    if (o.echo_mode < 0 || o.echo_mode > 4 || (o.cng != 0 && o.cng != 1) || o.chunk_calls < 1 || o.ms < 0 || o.ms > 500 || o.io_threads < 1) { usage(); return 2; }
    for (i = 0; i < n_lists; i++) {
        FILE *f = fopen(lists[i], "r");
        char line[4096];
        if (!f) { fprintf(stderr, "%s: %s\n", lists[i], strerror(errno)); return 2; }
        while (fgets(line, sizeof(line), f)) {
            char *far = strtok(line, " \t\r\n"), *near = far ? strtok(NULL, " \t\r\n") : NULL, *out = near ? strtok(NULL, " \t\r\n") : NULL;
            if (!far || far[0] == '#') continue;
            if (!near) { fprintf(stderr, "%s: line without a near file: %s\n", lists[i], far); fclose(f); return 2; }
            add_pair(&pairs, &n, &cap, far, near, out, o.suffix);
        }
        fclose(f);
    }
    if (n == 0) { usage(); return 2; }
    if (!o.passthrough && AecmB200_DeviceCount() < 1) { fprintf(stderr, "aecm_wav_batch: no CUDA device; this tool has no CPU path\n"); return 3; }

    t0 = now_s();
    run_io(pairs, n, o.io_threads, o.quiet, 0);
    for (i = 0; i < n; i++) { /* an unreadable / non-WAV input costs its own pair only (the reference CLI skips it the same way) */
        read_bad += pairs[i].failed;
        unsupported += pairs[i].unsupported;
    }
    t_read = now_s() - t0;

    t0 = now_s();
    if (!o.passthrough) {
        static const uint32_t rates[2] = {16000, 8000};
        int r;
        for (r = 0; r < 2; r++) {
            Pair **sel = (Pair **)malloc((size_t)n * sizeof(Pair *));
            Shard shards[64];
            pthread_t th[64];
            int m = 0;
            for (i = 0; i < n; i++)
                if (!pairs[i].skip && pairs[i].rate == rates[r] && pairs[i].n_calls > 0) sel[m++] = &pairs[i];
            if (m == 0) { free(sel); continue; }
            qsort(sel, (size_t)m, sizeof(Pair *), by_calls_desc);
            memset(shards, 0, sizeof(shards));
            for (d = 0; d < o.n_devices; d++) {
                Shard *s = &shards[d];
                int j;
                s->opt = &o; s->device = o.devices[d]; s->rate = rates[r]; s->spc = (int)(rates[r] / 100);
                s->pairs = (Pair **)malloc((size_t)(m / o.n_devices + 1) * sizeof(Pair *));
                for (j = d; j < m; j += o.n_devices) s->pairs[s->n++] = sel[j]; /* dealt round-robin: still longest first */
            }
            for (d = 0; d < o.n_devices; d++) pthread_create(&th[d], NULL, run_shard, &shards[d]);
            for (d = 0; d < o.n_devices; d++) {
                pthread_join(th[d], NULL);
                if (shards[d].failed) { /* only the pairs of THIS shard lose their output; every other pair is still written */
                    int j;
                    fprintf(stderr, "error (device %d): %s; %d pair(s) of this shard get no output\n", shards[d].device, shards[d].err,
                            shards[d].n);
                    for (j = 0; j < shards[d].n; j++) shards[d].pairs[j]->skip = 1;
                    proc_bad += shards[d].n;
                }
                blocks += shards[d].blocks_done;
                if (shards[d].init_s > t_init) t_init = shards[d].init_s;
                launches += shards[d].launches;
                free(shards[d].pairs);
            }
            free(sel);
        }
    }
    t_proc = now_s() - t0;

    t0 = now_s();
    for (i = 0; i < n; i++) pairs[i].failed = 0;
    run_io(pairs, n, o.io_threads, o.quiet, 1);
    for (i = 0; i < n; i++) {
        Pair *p = &pairs[i];
        if (p->skip) { skipped++; continue; }
        if (p->failed) { bad++; continue; }
        written++;
        if (p->spc) audio_s += (double)p->n_calls / 100.0;
    }
    t_write = now_s() - t0;
    if (!o.quiet)
        printf("aecm_wav_batch: %d pairs written, %d skipped, %d unchanged (unsupported rate); %.1f s of audio, %lld blocks in %d launches on %d device(s); "
               "read %.3f s, device set-up %.3f s, process %.3f s (%.3g blocks/s, %.0fx real time), write %.3f s\n",
               written, skipped, unsupported, audio_s, (long long)blocks, launches, o.n_devices, t_read, t_init, t_proc - t_init,
               t_proc > t_init ? (double)blocks / (t_proc - t_init) : 0.0, t_proc > t_init ? audio_s / (t_proc - t_init) : 0.0, t_write);
    for (i = 0; i < n; i++) { free(pairs[i].far); free(pairs[i].near); free(pairs[i].far_path); free(pairs[i].near_path); free(pairs[i].out_path); }
    free(pairs);
    if (read_bad || proc_bad || bad)
        fprintf(stderr, "aecm_wav_batch: %d pair(s) unreadable, %d lost to a failed device shard, %d not writable; the other %d were written\n",
                read_bad, proc_bad, bad, written);
    return (read_bad || proc_bad || bad) ? 1 : 0;
}
