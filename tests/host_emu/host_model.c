/*
 * host_model.c -- TEST INFRASTRUCTURE.  A model of everything BELOW the product's host layer, so that the host layer
 * itself -- webrtc_aecm_b200/csrc/aecm_batch.c and csrc/echo_control_mobile.c, compiled unchanged -- can be run on a machine
 * without a GPU and under sanitizers (tests/test_host_orchestration_model.py):
 *
 *   * the CUDA runtime calls aecm_batch.c makes (tests/host_emu/fake_cuda/cuda_runtime_api.h).  A stream is an in-order
 *     queue that is executed LAZILY: work runs only when something forces it (a synchronize on its stream, an event another
 *     forced stream waits for, a cudaFree).  Any order the real device may legally pick is thereby replaced by the latest
 *     one, which is the order that exposes a missing dependency (a kernel launched without waiting for its H2D copy reads
 *     poisoned staging memory) or a host buffer reused before its call was waited for.  FAKE_CUDA_ORDER=eager runs
 *     everything at once instead, =random:<seed> runs random prefixes at random moments.  Copies from/to host memory that
 *     did not come from cudaMallocHost behave like the real runtime's pageable path (H2D snapshots at the call, D2H
 *     completes inside the call); a pinned copy that runs past the end of its allocation is refused with
 *     cudaErrorInvalidValue, as the real runtime refuses it.  "Device" memory is host memory filled with a poison pattern.
 *   * the launch interface of aecm_kernels.cu (csrc/aecm_state.h: aecm_cuda_launch_*): every launch becomes a stream
 *     operation that steps the CPU ORACLE for the rows it names -- the arithmetic is not under test here, the plumbing is.
 *     The two state fields the host reads straight from the device record (init_flag, ch_stored: AecmB200_GetEchoPath)
 *     and a handful of scalars are mirrored into the record after every operation; the rest of a record stays poison.
 *
 * Nothing here is shipped, measured or reachable from the product: the resulting library is built by the test into a
 * temporary directory and announces itself on stderr when it is loaded.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cuda_runtime_api.h"

#include "aecm_state.h"
#include "../../oracle/aecm_oracle.h"

/* ------------------------------------------------------------------ bookkeeping ---------------------------------- */
typedef enum { OP_COPY, OP_COPY2D, OP_MEMSET, OP_RECORD, OP_WAIT, OP_KERNEL } OpKind;
typedef enum { K_INIT, K_SET_CONFIG, K_INIT_ECHO_PATH, K_CALL, K_DLY_OPTIONS, K_DLY_SOFT_RESET, K_DLY_QUERY } KernelKind;

struct FakeEvent {
    long gen_issued, gen_done;
    struct FakeStream *last_stream;
    int refs, destroyed;
};

typedef struct Op {
    OpKind kind;
    struct Op *next;
    void *dst;
    const void *src;
    void *staged; /* snapshot of a pageable H2D source, owned by the op */
    size_t bytes, dpitch, spitch, width, height;
    int value;
    struct FakeEvent *ev;
    long gen;
    struct FakeStream *wait_stream;
    KernelKind kk;
    AecmBatchPlanes planes;
    AecmCallArgs call;
    int a[6];
    void *p0, *p1;
} Op;

struct FakeStream {
    Op *head, *tail;
    struct FakeStream *next_all;
};

typedef struct Plane { /* the oracle instances behind one state plane */
    AecmInstanceState *base;
    int n;
    OracleAecm **inst;
    struct Plane *next;
} Plane;

typedef struct Pinned {
    char *p;
    size_t bytes;
    struct Pinned *next;
} Pinned;

static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static struct FakeStream *g_streams;
static struct FakeStream g_null_stream; /* the legacy default stream is not used by the product; kept for completeness */
static Plane *g_planes;
static Pinned *g_pinned;
static __thread int t_device;
static __thread cudaError_t t_last_error;
static int g_launches;
static int g_mode = -1; /* 0 lazy, 1 eager, 2 random */
static uint32_t g_rng = 1;
static long g_ops_run, g_ops_forced_by_wait;

static void lock(void) { pthread_mutex_lock(&g_mu); }
static void unlock(void) { pthread_mutex_unlock(&g_mu); }
static cudaError_t err(cudaError_t e) {
    if (e != cudaSuccess) t_last_error = e;
    return e;
}

__attribute__((constructor)) static void announce(void) {
    const char *m = getenv("FAKE_CUDA_ORDER");
    g_mode = 0;
    if (m && !strcmp(m, "eager")) g_mode = 1;
    if (m && !strncmp(m, "random", 6)) {
        g_mode = 2;
        g_rng = m[6] == ':' ? (uint32_t)atoi(m + 7) * 2654435761u + 1u : 12345u;
    }
    if (getenv("FAKE_CUDA_FAIL_MALLOC_IN")) { /* for programs that cannot call host_model_fail_malloc_in() themselves */
        extern void host_model_fail_malloc_in(int n);
        host_model_fail_malloc_in(atoi(getenv("FAKE_CUDA_FAIL_MALLOC_IN")));
    }
    fprintf(stderr, "aecm host model loaded: TEST INFRASTRUCTURE (oracle behind a modelled CUDA runtime, %s execution) -- not the product\n",
            g_mode == 0 ? "lazy" : g_mode == 1 ? "eager" : "random");
}

static uint32_t rnd(void) {
    g_rng ^= g_rng << 13; g_rng ^= g_rng >> 17; g_rng ^= g_rng << 5;
    return g_rng;
}

static void die(const char *what) {
    fprintf(stderr, "host model: %s\n", what);
    abort();
}

/* ------------------------------------------------------------------ the oracle behind the launches ---------------- */
static Plane *plane_of(const AecmBatchPlanes *p) {
    Plane *pl;
    for (pl = g_planes; pl; pl = pl->next)
        if (pl->base == p->state) return pl;
    pl = (Plane *)calloc(1, sizeof(*pl));
    pl->base = p->state;
    pl->n = p->n_instances;
    pl->inst = (OracleAecm **)calloc((size_t)pl->n, sizeof(OracleAecm *));
    for (int i = 0; i < pl->n; i++) pl->inst[i] = oracle_aecm_create();
    pl->next = g_planes;
    g_planes = pl;
    return pl;
}

static void mirror(Plane *pl, int i) { /* what the host reads from the record without a kernel */
    AecmInstanceState *rec = pl->base + i;
    rec->s.init_flag = pl->inst[i]->init_flag;
    rec->s.tot_count = pl->inst[i]->tot_count; /* a few more scalars that tests read through AecmB200_ReadState */
    rec->s.fs = pl->inst[i]->fs;
    rec->s.echo_mode = pl->inst[i]->echo_mode;
    rec->s.cng_mode = pl->inst[i]->cng_mode;
    rec->s.last_delay = pl->inst[i]->last_delay;
    memcpy(rec->ch_stored, pl->inst[i]->ch_stored, AECM_BINS * sizeof(int16_t));
}

static void plane_memset(char *dst, size_t bytes) { /* a memset that covers whole records = AecmB200_Create / _Reset */
    for (Plane *pl = g_planes; pl; pl = pl->next) {
        char *b = (char *)pl->base, *e = b + (size_t)pl->n * sizeof(AecmInstanceState);
        if (dst < b || dst >= e) continue;
        if ((size_t)(dst - b) % sizeof(AecmInstanceState) || bytes % sizeof(AecmInstanceState) || dst + bytes > e)
            die("memset of a state plane that does not cover whole records");
        for (size_t i = (size_t)(dst - b) / sizeof(AecmInstanceState), k = 0; k < bytes / sizeof(AecmInstanceState); k++, i++) {
            oracle_aecm_free(pl->inst[i]);
            pl->inst[i] = oracle_aecm_create();
        }
    }
}

static void plane_free(void *p) {
    for (Plane **pp = &g_planes; *pp; pp = &(*pp)->next)
        if ((void *)(*pp)->base == p) {
            Plane *pl = *pp;
            *pp = pl->next;
            for (int i = 0; i < pl->n; i++) oracle_aecm_free(pl->inst[i]);
            free(pl->inst);
            free(pl);
            return;
        }
}

static void run_kernel(Op *op) {
    Plane *pl = plane_of(&op->planes);
    const int first = op->a[0], count = op->a[1];
    switch (op->kk) {
    case K_INIT:
        for (int i = first; i < first + count; i++) { oracle_aecm_init(pl->inst[i], op->a[2]); mirror(pl, i); }
        break;
    case K_SET_CONFIG:
        for (int r = 0; r < count; r++) {
            const int32_t rc = oracle_aecm_set_config(pl->inst[first + r], (int16_t)op->a[2], (int16_t)op->a[3]);
            if (op->p0) ((int32_t *)op->p0)[r] = rc;
        }
        break;
    case K_INIT_ECHO_PATH:
        for (int r = 0; r < count; r++) {
            const int32_t rc = oracle_aecm_init_echo_path(pl->inst[first + r], (const int16_t *)op->p1 + (size_t)r * (size_t)op->a[2],
                                                          AECM_BINS * sizeof(int16_t));
            if (op->p0) ((int32_t *)op->p0)[r] = rc;
            mirror(pl, first + r);
        }
        break;
    case K_DLY_OPTIONS:
        for (int r = 0; r < count; r++) {
            OracleAecm *o = pl->inst[first + r];
            const int32_t rc = o->init_flag != AECM_INIT_MAGIC ? AECM_UNINITIALIZED_ERROR
                                                              : oracle_aecm_set_delay_estimator_options(o, op->a[2], op->a[3], op->a[4], op->a[5]);
            if (op->p0) ((int32_t *)op->p0)[r] = rc;
        }
        break;
    case K_DLY_SOFT_RESET:
        for (int r = 0; r < count; r++) {
            const int32_t rc = oracle_aecm_soft_reset_delay_estimator(pl->inst[first + r], op->a[2]);
            if (op->p0) ((int32_t *)op->p0)[r] = rc;
        }
        break;
    case K_DLY_QUERY:
        for (int r = 0; r < count; r++) {
            if (op->p0) ((int32_t *)op->p0)[r] = oracle_aecm_last_delay(pl->inst[first + r]);
            if (op->p1) ((float *)op->p1)[r] = oracle_aecm_last_delay_quality(pl->inst[first + r]);
        }
        break;
    case K_CALL: {
        const AecmCallArgs *a = &op->call;
        for (int r = 0; r < a->count; r++) {
            const int i = a->index ? a->index[r] : a->first + r;
            int32_t st = 0;
            if (i < 0 || i >= pl->n) { /* the kernel's own range check of a device index array */
                if (a->status) a->status[r] = AECM_BAD_PARAMETER_ERROR;
                continue;
            }
            for (int c = 0; c < a->n_calls; c++) {
                const size_t off = (size_t)r * (size_t)a->row_stride + (size_t)c * (size_t)a->samples_per_call;
                int32_t rc;
                if (a->far && (!a->far_mask || a->far_mask[(size_t)c * (size_t)a->mask_stride + (size_t)r])) {
                    rc = oracle_aecm_buffer_farend(pl->inst[i], a->far + off, (size_t)a->samples_per_call);
                    if (!st) st = rc;
                }
                if (a->near_noisy) {
                    rc = oracle_aecm_process(pl->inst[i], a->near_noisy + off, a->near_clean ? a->near_clean + off : NULL, a->out + off,
                                             (size_t)a->samples_per_call, a->ms ? a->ms[r] : (int16_t)a->ms_uniform);
                    if (!st) st = rc;
                }
            }
            if (a->status) a->status[r] = st;
            mirror(pl, i);
        }
        break;
    }
    }
}

/* ------------------------------------------------------------------ streams -------------------------------------- */
static void run_one(struct FakeStream *s);

static void ensure(struct FakeEvent *e, long gen, struct FakeStream *where) {
    while (e->gen_done < gen) {
        if (!where || !where->head) die("a stream waits for an event whose record is not queued anywhere (lost dependency)");
        g_ops_forced_by_wait++;
        run_one(where);
    }
}

static void event_unref(struct FakeEvent *e) {
    if (--e->refs == 0 && e->destroyed) free(e);
}

static void run_one(struct FakeStream *s) {
    Op *op = s->head;
    if (!op) return;
    if (op->kind == OP_WAIT) ensure(op->ev, op->gen, op->wait_stream); /* may run other streams; ours still starts with op */
    s->head = op->next;
    if (!s->head) s->tail = NULL;
    g_ops_run++;
    switch (op->kind) {
    case OP_COPY: memcpy(op->dst, op->src, op->bytes); break;
    case OP_COPY2D:
        for (size_t r = 0; r < op->height; r++) memcpy((char *)op->dst + r * op->dpitch, (const char *)op->src + r * op->spitch, op->width);
        break;
    case OP_MEMSET:
        memset(op->dst, op->value, op->bytes);
        if (op->value == 0) plane_memset((char *)op->dst, op->bytes);
        break;
    case OP_RECORD:
        if (op->gen > op->ev->gen_done) op->ev->gen_done = op->gen;
        event_unref(op->ev);
        break;
    case OP_WAIT: event_unref(op->ev); break;
    case OP_KERNEL: run_kernel(op); break;
    }
    free(op->staged);
    free(op);
}

static void drain(struct FakeStream *s) {
    while (s->head) run_one(s);
}
static void drain_through(struct FakeStream *s, const Op *last) {
    for (;;) {
        const int was_last = s->head == last;
        if (!s->head) die("drain_through: operation not queued");
        run_one(s);
        if (was_last) return;
    }
}
static void drain_all(void) {
    for (struct FakeStream *s = g_streams; s; s = s->next_all) drain(s);
    drain(&g_null_stream);
}

static struct FakeStream *stream_of(cudaStream_t s) { return s ? s : &g_null_stream; }

static Op *enqueue(struct FakeStream *s, OpKind kind) {
    Op *op = (Op *)calloc(1, sizeof(*op));
    op->kind = kind;
    if (s->tail) s->tail->next = op; else s->head = op;
    s->tail = op;
    return op;
}

/* after every enqueue: the execution-order policy */
static void policy(struct FakeStream *s) {
    if (g_mode == 1) {
        drain(s);
    } else if (g_mode == 2) {
        for (int k = (int)(rnd() % 4); k > 0; k--) {
            int n = 0, pick;
            for (struct FakeStream *q = g_streams; q; q = q->next_all) n += q->head != NULL;
            if (!n) break;
            pick = (int)(rnd() % (uint32_t)n);
            for (struct FakeStream *q = g_streams; q; q = q->next_all)
                if (q->head && pick-- == 0) { run_one(q); break; }
        }
    }
}

/* ------------------------------------------------------------------ the runtime API ------------------------------ */
const char *cudaGetErrorString(cudaError_t e) {
    switch (e) {
    case cudaSuccess: return "no error";
    case cudaErrorInvalidValue: return "invalid argument";
    case cudaErrorMemoryAllocation: return "out of memory";
    case cudaErrorInvalidDevice: return "invalid device ordinal";
    default: return "unknown error";
    }
}
cudaError_t cudaGetLastError(void) {
    const cudaError_t e = t_last_error;
    t_last_error = cudaSuccess;
    return e;
}
static int n_devices(void) {
    const char *e = getenv("FAKE_CUDA_DEVICES");
    return e ? atoi(e) : 2;
}
cudaError_t cudaGetDeviceCount(int *n) { *n = n_devices(); return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = t_device; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) {
    if (d < 0 || d >= n_devices()) return err(cudaErrorInvalidDevice);
    t_device = d;
    return cudaSuccess;
}

int host_model_take_malloc_failure(void);
cudaError_t cudaMalloc(void **p, size_t bytes) {
    if (host_model_take_malloc_failure()) {
        *p = NULL;
        return err(cudaErrorMemoryAllocation);
    }
    *p = malloc(bytes ? bytes : 1);
    if (!*p) return err(cudaErrorMemoryAllocation);
    memset(*p, 0xA5, bytes); /* never-written device memory must not look like silence */
    return cudaSuccess;
}
cudaError_t cudaFree(void *p) {
    if (!p) return cudaSuccess;
    lock();
    drain_all(); /* cudaFree synchronises the device */
    plane_free(p);
    unlock();
    free(p);
    return cudaSuccess;
}
cudaError_t cudaMallocHost(void **p, size_t bytes) {
    Pinned *h;
    if (host_model_take_malloc_failure()) {
        *p = NULL;
        return err(cudaErrorMemoryAllocation);
    }
    h = (Pinned *)calloc(1, sizeof(*h));
    *p = malloc(bytes ? bytes : 1);
    if (!*p || !h) { free(h); return err(cudaErrorMemoryAllocation); }
    memset(*p, 0x5A, bytes);
    h->p = (char *)*p;
    h->bytes = bytes;
    lock();
    h->next = g_pinned;
    g_pinned = h;
    unlock();
    return cudaSuccess;
}
cudaError_t cudaFreeHost(void *p) {
    if (!p) return cudaSuccess;
    lock();
    drain_all();
    for (Pinned **pp = &g_pinned; *pp; pp = &(*pp)->next)
        if ((*pp)->p == (char *)p) {
            Pinned *h = *pp;
            *pp = h->next;
            free(h);
            break;
        }
    unlock();
    free(p);
    return cudaSuccess;
}
/* 1 = inside one pinned allocation, 0 = pageable, -1 = starts in a pinned allocation and runs past its end */
static int pinned_kind(const void *p, size_t extent) {
    for (Pinned *h = g_pinned; h; h = h->next)
        if ((const char *)p >= h->p && (const char *)p < h->p + h->bytes) return (const char *)p + extent <= h->p + h->bytes ? 1 : -1;
    return 0;
}

cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags) {
    (void)flags;
    *s = (struct FakeStream *)calloc(1, sizeof(struct FakeStream));
    lock();
    (*s)->next_all = g_streams;
    g_streams = *s;
    unlock();
    return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t s) {
    if (!s) return err(cudaErrorInvalidResourceHandle);
    lock();
    drain(s);
    for (struct FakeStream **pp = &g_streams; *pp; pp = &(*pp)->next_all)
        if (*pp == s) { *pp = s->next_all; break; }
    unlock();
    free(s);
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t s) {
    lock();
    drain(stream_of(s));
    unlock();
    return cudaSuccess;
}
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags) {
    (void)flags;
    *e = (struct FakeEvent *)calloc(1, sizeof(struct FakeEvent));
    return *e ? cudaSuccess : err(cudaErrorMemoryAllocation);
}
cudaError_t cudaEventCreate(cudaEvent_t *e) { return cudaEventCreateWithFlags(e, 0); }
cudaError_t cudaEventDestroy(cudaEvent_t e) {
    if (!e) return err(cudaErrorInvalidResourceHandle);
    lock();
    e->destroyed = 1; /* work already queued keeps it alive, as on the device */
    if (e->refs == 0) free(e);
    unlock();
    return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
    Op *op;
    if (!e) return err(cudaErrorInvalidResourceHandle);
    lock();
    op = enqueue(stream_of(s), OP_RECORD);
    op->ev = e;
    op->gen = ++e->gen_issued;
    e->last_stream = stream_of(s);
    e->refs++;
    policy(stream_of(s));
    unlock();
    return cudaSuccess;
}
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags) {
    Op *op;
    (void)flags;
    if (!e) return err(cudaErrorInvalidResourceHandle);
    lock();
    if (e->gen_issued > e->gen_done) { /* the record it refers to: the latest one issued so far */
        op = enqueue(stream_of(s), OP_WAIT);
        op->ev = e;
        op->gen = e->gen_issued;
        op->wait_stream = e->last_stream;
        e->refs++;
        policy(stream_of(s));
    }
    unlock();
    return cudaSuccess;
}
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
    (void)a; (void)b;
    *ms = 0.f;
    return cudaSuccess;
}

static cudaError_t copy_common(Op *proto, const void *host, size_t host_extent, enum cudaMemcpyKind kind, cudaStream_t s) {
    struct FakeStream *st = stream_of(s);
    int pk = 1;
    Op *op;
    lock();
    if (kind == cudaMemcpyHostToDevice || kind == cudaMemcpyDeviceToHost) {
        pk = pinned_kind(host, host_extent);
        if (pk < 0) { /* one DMA cannot cross from one pinned allocation into whatever lies behind it */
            unlock();
            return err(cudaErrorInvalidValue);
        }
    }
    op = enqueue(st, proto->kind);
    {
        Op *next = op->next;
        *op = *proto;
        op->next = next;
    }
    if (kind == cudaMemcpyHostToDevice && pk == 0) { /* pageable source: the runtime stages it before returning */
        op->staged = malloc(host_extent ? host_extent : 1);
        memcpy(op->staged, host, host_extent);
        op->src = op->staged;
    }
    if (kind == cudaMemcpyDeviceToHost && pk == 0) drain_through(st, op); /* pageable destination: done when the call returns */
    else policy(st);
    unlock();
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, enum cudaMemcpyKind kind, cudaStream_t s) {
    Op p;
    memset(&p, 0, sizeof(p));
    p.kind = OP_COPY;
    p.dst = dst; p.src = src; p.bytes = bytes;
    return copy_common(&p, kind == cudaMemcpyHostToDevice ? src : dst, bytes, kind, s);
}
cudaError_t cudaMemcpy2DAsync(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height,
                              enum cudaMemcpyKind kind, cudaStream_t s) {
    Op p;
    const size_t hp = kind == cudaMemcpyHostToDevice ? spitch : dpitch;
    if (width > dpitch || width > spitch) return err(cudaErrorInvalidValue);
    memset(&p, 0, sizeof(p));
    p.kind = OP_COPY2D;
    p.dst = dst; p.src = src; p.dpitch = dpitch; p.spitch = spitch; p.width = width; p.height = height;
    return copy_common(&p, kind == cudaMemcpyHostToDevice ? src : dst, height ? (height - 1) * hp + width : 0, kind, s);
}
cudaError_t cudaMemsetAsync(void *dst, int value, size_t bytes, cudaStream_t s) {
    Op *op;
    lock();
    op = enqueue(stream_of(s), OP_MEMSET);
    op->dst = dst; op->value = value; op->bytes = bytes;
    policy(stream_of(s));
    unlock();
    return cudaSuccess;
}

/* ------------------------------------------------------------------ the launch interface (csrc/aecm_state.h) ----- */
static Op *launch(void *stream, KernelKind kk, const AecmBatchPlanes *p) {
    Op *op = enqueue(stream_of((cudaStream_t)stream), OP_KERNEL);
    op->kk = kk;
    op->planes = *p;
    g_launches++;
    return op;
}
#define LAUNCH_END(stream)                     \
    policy(stream_of((cudaStream_t)(stream))); \
    unlock();                                  \
    return 0

int aecm_cuda_launch_init(const AecmBatchPlanes *p, int first, int count, int fs, void *stream) {
    Op *op;
    if (count <= 0) return 0;
    lock();
    op = launch(stream, K_INIT, p);
    op->a[0] = first; op->a[1] = count; op->a[2] = fs;
    LAUNCH_END(stream);
}
int aecm_cuda_launch_set_config(const AecmBatchPlanes *p, int first, int count, int cng, int mode, int32_t *status, void *stream) {
    Op *op;
    if (count <= 0) return 0;
    lock();
    op = launch(stream, K_SET_CONFIG, p);
    op->a[0] = first; op->a[1] = count; op->a[2] = cng; op->a[3] = mode; op->p0 = status;
    LAUNCH_END(stream);
}
int aecm_cuda_launch_init_echo_path(const AecmBatchPlanes *p, int first, int count, const int16_t *d_path, int stride, int32_t *status,
                                    void *stream) {
    Op *op;
    if (count <= 0) return 0;
    lock();
    op = launch(stream, K_INIT_ECHO_PATH, p);
    op->a[0] = first; op->a[1] = count; op->a[2] = stride; op->p0 = status; op->p1 = (void *)(uintptr_t)d_path;
    LAUNCH_END(stream);
}
int aecm_cuda_launch_call(const AecmBatchPlanes *p, const AecmCallArgs *a, void *stream) {
    Op *op;
    if (a->count <= 0 || a->n_calls <= 0) return 0;
    lock();
    op = launch(stream, K_CALL, p);
    op->call = *a;
    LAUNCH_END(stream);
}
int aecm_cuda_launch_dly_options(const AecmBatchPlanes *p, int first, int count, int robust, int allowed_offset, int max_lookahead,
                                 int lookahead, int32_t *status, void *stream) {
    Op *op;
    if (count <= 0) return 0;
    lock();
    op = launch(stream, K_DLY_OPTIONS, p);
    op->a[0] = first; op->a[1] = count; op->a[2] = robust; op->a[3] = allowed_offset; op->a[4] = max_lookahead; op->a[5] = lookahead;
    op->p0 = status;
    LAUNCH_END(stream);
}
int aecm_cuda_launch_dly_soft_reset(const AecmBatchPlanes *p, int first, int count, int shift, int32_t *applied, void *stream) {
    Op *op;
    if (count <= 0) return 0;
    lock();
    op = launch(stream, K_DLY_SOFT_RESET, p);
    op->a[0] = first; op->a[1] = count; op->a[2] = shift; op->p0 = applied;
    LAUNCH_END(stream);
}
int aecm_cuda_launch_dly_query(const AecmBatchPlanes *p, int first, int count, int32_t *last_delay, float *quality, void *stream) {
    Op *op;
    if (count <= 0) return 0;
    lock();
    op = launch(stream, K_DLY_QUERY, p);
    op->a[0] = first; op->a[1] = count; op->p0 = last_delay; op->p1 = quality;
    LAUNCH_END(stream);
}
long long aecm_cuda_selftest_sqrt(uint32_t first, uint32_t count) {
    unsigned long long bad = 0;
    const uint64_t end = (uint64_t)first + count;
    for (uint64_t v = first; v < end && v < 0x80000000ull; v += 65537) { /* a sparse sweep: the oracle's, not the device's, arithmetic */
        const int64_t r = oracle_sqrt_floor((int32_t)v);
        bad += !(r * r <= (int64_t)v && (r + 1) * (r + 1) > (int64_t)v);
    }
    return (long long)bad;
}
int aecm_cuda_selftest_fft(int mode, const int16_t *in, int16_t *out, int32_t *scale, int n) {
    for (int v = 0; v < n; v++) {
        if (mode == 2) {
            const int sc = oracle_real_fft128_inverse(in + (size_t)v * 130, out + (size_t)v * 128);
            if (scale) scale[v] = sc;
        } else {
            oracle_real_fft128_forward(in + (size_t)v * 128, out + (size_t)v * 130);
        }
    }
    return 0;
}
int aecm_cuda_kernel_launch_count(void) { return g_launches; }

/* for the tests: how much work ran, and how much of it only because another stream's wait demanded it */
void host_model_stats(long *ops_run, long *ops_forced_by_wait) {
    lock();
    if (ops_run) *ops_run = g_ops_run;
    if (ops_forced_by_wait) *ops_forced_by_wait = g_ops_forced_by_wait;
    unlock();
}

/* ------------------------------------------------------------------ hooks for tests/host_emu/model_cases.py -------- */
/* two pinned "allocations" of `bytes` each that touch in the address space (what two cudaMallocHost calls can produce):
 * *a ends exactly where *b begins.  Free both with host_model_free_touching(*a). */
int host_model_alloc_touching(size_t bytes, void **a, void **b) {
    char *p = (char *)malloc(2 * bytes);
    Pinned *ha = (Pinned *)calloc(1, sizeof(Pinned)), *hb = (Pinned *)calloc(1, sizeof(Pinned));
    if (!p || !ha || !hb) return 1;
    ha->p = p; ha->bytes = bytes;
    hb->p = p + bytes; hb->bytes = bytes;
    lock();
    ha->next = g_pinned; hb->next = ha; g_pinned = hb;
    unlock();
    *a = p;
    *b = p + bytes;
    return 0;
}
static size_t unpin(const char *p) {
    for (Pinned **pp = &g_pinned; *pp; pp = &(*pp)->next)
        if ((*pp)->p == p) {
            Pinned *h = *pp;
            const size_t bytes = h->bytes;
            *pp = h->next;
            free(h);
            return bytes;
        }
    return 0;
}
void host_model_free_touching(void *a) {
    lock();
    drain_all();
    unpin((const char *)a + unpin((const char *)a));
    unlock();
    free(a);
}
static int g_fail_malloc_in = 0; /* > 0: the n-th device / pinned allocation from now fails once (allocations come from many threads) */
void host_model_fail_malloc_in(int n) { __atomic_store_n(&g_fail_malloc_in, n, __ATOMIC_SEQ_CST); }
int host_model_take_malloc_failure(void) {
    if (__atomic_load_n(&g_fail_malloc_in, __ATOMIC_SEQ_CST) <= 0) return 0;
    return __atomic_sub_fetch(&g_fail_malloc_in, 1, __ATOMIC_SEQ_CST) == 0;
}
int host_model_current_device(void) { return t_device; }
long host_model_pending_ops(void) {
    long n = 0;
    lock();
    for (struct FakeStream *s = g_streams; s; s = s->next_all)
        for (Op *op = s->head; op; op = op->next) n++;
    unlock();
    return n;
}
