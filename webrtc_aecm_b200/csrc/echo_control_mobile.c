/*
 * echo_control_mobile.c -- the reference's ten scalar entry points (aecm/echo_control_mobile.h:46-202) as
 * thin views over slots of device-resident batches.  Handles are small host descriptors; the AECM state and
 * every bit of arithmetic live on the GPU (aecm_kernels.cu).  Error codes and their precedence follow
 * aecm/echo_control_mobile.cc line by line (cited below).
 *
 * Threading.  The reference keeps all state inside the instance, so a program may drive different instances from
 * different threads at the same time; the same holds here.  Instances share a batch (256 to a pool), and a batch
 * takes one caller at a time, so every entry point serialises on its pool.  BufferFarend/Process calls go through a
 * COMBINER: a caller queues its request; whoever finds no round in progress becomes the leader, takes everything
 * queued so far (requests of all threads, for different instances), packs the audio rows into pinned staging
 * planes and issues them as ONE launch per kind (rows name their instances through
 * AecmB200StreamArgs.instanceIndex), hands each result back to its owner and repeats until its own request has been
 * served; then the oldest request still queued is woken to lead the next round.  T threads
 * calling concurrently therefore cost about one launch + one PCIe round trip per round, not T of them
 * ("batching of scalar WebRtcAecm_Process calls into lockstep launches").
 *
 * Deferred far end.  WebRtcAecm_BufferFarend can only fail on its arguments or on an un-initialised instance
 * (echo_control_mobile.cc:195-213), all of which the host knows (initFlag is mirrored in the handle: only Init sets
 * it, only Create clears it).  So the call validates, keeps the frame in the handle and returns; the frame travels
 * with the NEXT request of that instance -- normally the WebRtcAecm_Process that follows, which then runs
 * BufferFarend + Process in ONE launch, in the reference's order.  Anything else that touches the instance first
 * (a second BufferFarend, a different frame size, set_config, the echo-path calls) sends the pending frame ahead
 * on its own; Init and Free drop it, as the reference's Init drops everything buffered (:163-166).
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "aecm_b200.h"

#define POOL_SLOTS 256
#define HANDLE_MAGIC 0x41454342u /* "AECB" */

#define MAX_SAMPLES 160


typedef struct Req {
    int32_t slot;
    const int16_t *far, *near_noisy, *near_clean; /* far only = BufferFarend; near (+ far) = [BufferFarend +] Process */
    int16_t *out;
    int n;
    int16_t ms;
    int32_t result;
    int done;
    pthread_cond_t cond; /* signalled when done, or when this request's owner should take over as leader */
    struct Req *next;
} Req;

typedef struct Pool {
    AecmB200Batch *batch;
    unsigned char used[POOL_SLOTS];
    int n_used;
    struct Pool *next;
    /* the combiner */
    pthread_mutex_t q_lock; /* queue, leader flag, Req.done */
    Req *head, *tail;
    int leader;
    pthread_mutex_t exec;   /* held by whoever is talking to the batch */
    int16_t *h_far, *h_near, *h_clean, *h_out; /* pinned, [POOL_SLOTS][MAX_SAMPLES] */
    int32_t index[POOL_SLOTS], status[POOL_SLOTS];
    int16_t ms[POOL_SLOTS];
    long rounds, combined;  /* statistics: rounds run, requests served */
} Pool;

typedef struct {
    uint32_t magic;
    Pool *pool;
    int32_t slot;
    int initialized;            /* host mirror of the device's initFlag */
    int pending_n;              /* samples of a far-end frame accepted by BufferFarend but not yet sent (0 = none) */
    int16_t pending[MAX_SAMPLES];
} Handle;

static Pool *g_pools = NULL;
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;

/* Device of the next pool: AECM_B200_DEVICE pins every pool to one GPU; otherwise pools are dealt round-robin over
 * the visible GPUs (256 handles per pool), so a process with thousands of scalar handles uses the whole box. */
static int g_pools_created = 0;
static int next_pool_device(void) {
    const char *e = getenv("AECM_B200_DEVICE");
    int n;
    if (e) return atoi(e);
    n = AecmB200_DeviceCount();
    return n > 0 ? g_pools_created % n : 0;
}

static void destroy_pool(Pool *p) { /* g_lock held; no handle of the pool is left, so nobody can be inside it */
    Pool **pp;
    for (pp = &g_pools; *pp && *pp != p; pp = &(*pp)->next) {}
    if (*pp) *pp = p->next;
    AecmB200_HostFree(p->h_far); AecmB200_HostFree(p->h_near); AecmB200_HostFree(p->h_clean); AecmB200_HostFree(p->h_out);
    AecmB200_Free(p->batch);
    pthread_mutex_destroy(&p->q_lock);
    pthread_mutex_destroy(&p->exec);
    free(p);
}

void *WebRtcAecm_Create(void) { /* echo_control_mobile.cc:89-117 */
    Handle *h = NULL;
    Pool *p;
    int s;
    pthread_mutex_lock(&g_lock);
    for (p = g_pools; p && p->n_used == POOL_SLOTS; p = p->next) {}
    if (!p) {
        p = (Pool *)calloc(1, sizeof(Pool));
        if (p) {
            const size_t plane = (size_t)POOL_SLOTS * MAX_SAMPLES * sizeof(int16_t);
            p->batch = AecmB200_Create(POOL_SLOTS, next_pool_device());
            if (p->batch) {
                p->h_far = (int16_t *)AecmB200_HostAlloc(plane);
                p->h_near = (int16_t *)AecmB200_HostAlloc(plane);
                p->h_clean = (int16_t *)AecmB200_HostAlloc(plane);
                p->h_out = (int16_t *)AecmB200_HostAlloc(plane);
            }
            if (!p->batch || !p->h_far || !p->h_near || !p->h_clean || !p->h_out) {
                AecmB200_HostFree(p->h_far); AecmB200_HostFree(p->h_near); AecmB200_HostFree(p->h_clean); AecmB200_HostFree(p->h_out);
                AecmB200_Free(p->batch);
                free(p);
                p = NULL;
            } else {
                pthread_mutex_init(&p->q_lock, NULL);
                pthread_mutex_init(&p->exec, NULL);
                p->next = g_pools;
                g_pools = p;
                g_pools_created++;
            }
        }
    }
    if (p) {
        for (s = 0; p->used[s]; s++) {}
        h = (Handle *)calloc(1, sizeof(Handle));
        /* a fresh instance is zero-filled and un-initialised, like the reference's calloc (:91) */
        if (h) {
            pthread_mutex_lock(&p->exec);
            if (AecmB200_Reset(p->batch, s, 1) != 0) {
                free(h);
                h = NULL;
            }
            pthread_mutex_unlock(&p->exec);
        }
        if (h) {
            p->used[s] = 1;
            p->n_used++;
            h->magic = HANDLE_MAGIC;
            h->pool = p;
            h->slot = s;
        }
    }
    pthread_mutex_unlock(&g_lock);
    return h;
}

void WebRtcAecm_Free(void *inst) { /* echo_control_mobile.cc:119-140 */
    Handle *h = (Handle *)inst;
    if (!h || h->magic != HANDLE_MAGIC) return;
    pthread_mutex_lock(&g_lock);
    h->pool->used[h->slot] = 0;
    h->magic = 0;
    /* the last handle of a pool takes the pool with it -- its 256-instance device batch, pinned planes, streams and
     * events -- unless it is the only pool left: one empty pool is kept so that Create/Free cycles of a single handle
     * (the reference CLI's pattern, main.cc:111 / :146) do not pay a device allocation each time */
    if (--h->pool->n_used == 0 && (g_pools != h->pool || h->pool->next)) destroy_pool(h->pool);
    pthread_mutex_unlock(&g_lock);
    free(h);
}

int32_t WebRtcAecm_Init(void *inst, int32_t sampFreq) { /* echo_control_mobile.cc:142-191 */
    Handle *h = (Handle *)inst;
    int32_t rc;
    if (!h) return -1;
    if (sampFreq != 8000 && sampFreq != 16000) return AECM_BAD_PARAMETER_ERROR;
    h->pending_n = 0; /* Init empties the far-end buffer (:163-166): a frame not yet sent is dropped with it */
    pthread_mutex_lock(&h->pool->exec);
    rc = AecmB200_Init(h->pool->batch, h->slot, 1, sampFreq);
    if (!rc) rc = AecmB200_Synchronize(h->pool->batch) ? AECM_UNSPECIFIED_ERROR : 0;
    pthread_mutex_unlock(&h->pool->exec);
    if (!rc) h->initialized = 1;
    return rc;
}

/* One launch for every request of the same shape (far / near / both, frame size, clean-or-not) in the round.  Rows are packed densely from row `r0` of
 * the pinned planes; `slot` alternates so that the copies of one group overlap the kernel of another. */
static void issue_group(Pool *p, Req **reqs, int count, int r0, int slot) {
    AecmB200StreamArgs a;
    const Req *q0 = reqs[0];
    const int n = q0->n;
    const size_t base = (size_t)r0 * MAX_SAMPLES;
    int32_t rc;
    int i;
    for (i = 0; i < count; i++) {
        const Req *q = reqs[i];
        p->index[r0 + i] = q->slot;
        p->ms[r0 + i] = q->ms;
        p->status[r0 + i] = 0;
        if (q->far) memcpy(p->h_far + base + (size_t)i * n, q->far, (size_t)n * sizeof(int16_t));
        if (q->near_noisy) memcpy(p->h_near + base + (size_t)i * n, q->near_noisy, (size_t)n * sizeof(int16_t));
        if (q->near_clean) memcpy(p->h_clean + base + (size_t)i * n, q->near_clean, (size_t)n * sizeof(int16_t));
    }
    memset(&a, 0, sizeof(a));
    if (q0->far) a.farend = p->h_far + base;
    if (q0->near_noisy) {
        a.nearendNoisy = p->h_near + base;
        a.nearendClean = q0->near_clean ? p->h_clean + base : NULL;
        a.out = p->h_out + base;
    }
    a.rowStride = n;
    a.count = count;
    a.nrOfSamples = n;
    a.nCalls = 1;
    a.msInSndCardBuf = p->ms + r0;
    a.status = p->status + r0;
    a.instanceIndex = p->index + r0;
    rc = AecmB200_WaitStream(p->batch, slot);
    if (!rc) rc = AecmB200_SubmitStream(p->batch, &a, slot);
    for (i = 0; i < count; i++) reqs[i]->result = rc; /* 0 = in flight: the per-row status follows in collect_group */
}

static void collect_group(Pool *p, Req **reqs, int count, int r0, int slot) {
    const int32_t rc = AecmB200_WaitStream(p->batch, slot);
    int i;
    for (i = 0; i < count; i++) {
        Req *q = reqs[i];
        if (q->result) continue; /* the submission itself failed */
        if (rc) { q->result = rc; continue; }
        q->result = p->status[r0 + i];
        if (q->near_noisy) memcpy(q->out, p->h_out + (size_t)r0 * MAX_SAMPLES + (size_t)i * q->n, (size_t)q->n * sizeof(int16_t));
    }
}

/* One round: the leader owns `list` (at most POOL_SLOTS requests on distinct instances). */
static void run_round(Pool *p, Req *list) {
    Req *group[12][POOL_SLOTS];
    int count[12] = {0}, r0[12], pending[16], g, rows = 0, issued = 0, s;
    const int n_slots = AecmB200_NumStreamSlots() < 16 ? AecmB200_NumStreamSlots() : 16;
    Req *q;
    for (q = list; q; q = q->next) { /* far-only / near-only / both x frame size x clean-or-not: a launch is uniform in all */
        g = (q->near_noisy ? (q->far ? 2 : 1) : 0) * 4 + (q->n == 160 ? 2 : 0) + (q->near_clean ? 1 : 0);
        group[g][count[g]++] = q;
    }
    for (s = 0; s < n_slots; s++) pending[s] = -1;
    pthread_mutex_lock(&p->exec);
    for (g = 0; g < 12; g++) { /* the copies of one group overlap the kernel of another */
        if (!count[g]) continue;
        s = issued++ % n_slots;
        if (pending[s] >= 0) collect_group(p, group[pending[s]], count[pending[s]], r0[pending[s]], s);
        r0[g] = rows;
        rows += count[g];
        issue_group(p, group[g], count[g], r0[g], s);
        pending[s] = g;
    }
    for (s = 0; s < n_slots; s++)
        if (pending[s] >= 0) collect_group(p, group[pending[s]], count[pending[s]], r0[pending[s]], s);
    pthread_mutex_unlock(&p->exec);
}

/* Queue the request and wait for its result; lead a round if nobody is. */
static int32_t combine(Pool *p, Req *r) {
    r->done = 0;
    r->next = NULL;
    pthread_cond_init(&r->cond, NULL);
    pthread_mutex_lock(&p->q_lock);
    if (p->tail) p->tail->next = r; else p->head = r;
    p->tail = r;
    while (!r->done) {
        if (p->leader) { /* sleeps until its result is in, or until leadership is handed to it */
            pthread_cond_wait(&r->cond, &p->q_lock);
            continue;
        }
        p->leader = 1;
        while (!r->done && p->head) { /* lead until the own request has been served (it is still queued until then) */
            /* take up to POOL_SLOTS requests on distinct instances, in arrival order; a second request for an instance
             * already in the round (two threads on one instance) stays queued for the next round */
            unsigned char seen[POOL_SLOTS] = {0};
            Req *list = NULL, *list_tail = NULL, *keep = NULL, *keep_tail = NULL, *q = p->head, *nx;
            int taken = 0;
            for (; q; q = nx) {
                nx = q->next;
                q->next = NULL;
                if (taken < POOL_SLOTS && !seen[q->slot]) {
                    seen[q->slot] = 1;
                    taken++;
                    if (list_tail) list_tail->next = q; else list = q;
                    list_tail = q;
                } else {
                    if (keep_tail) keep_tail->next = q; else keep = q;
                    keep_tail = q;
                }
            }
            p->head = keep;
            p->tail = keep_tail;
            pthread_mutex_unlock(&p->q_lock);
            run_round(p, list);
            pthread_mutex_lock(&p->q_lock);
            p->rounds++;
            for (q = list; q; q = nx) {
                nx = q->next; /* the owner may release the request as soon as it sees done */
                q->done = 1;
                p->combined++;
                if (q != r) pthread_cond_signal(&q->cond); /* only the owner wakes up */
            }
        }
        p->leader = 0;
        if (p->head) pthread_cond_signal(&p->head->cond); /* requests that arrived meanwhile: the oldest one leads next */
    }
    pthread_mutex_unlock(&p->q_lock);
    pthread_cond_destroy(&r->cond);
    return r->result;
}

/* send a far-end frame that was accepted earlier on its own (it cannot ride with a Process call) */
static int32_t flush_pending(Handle *h) {
    Req r;
    if (!h->pending_n) return 0;
    memset(&r, 0, sizeof(r));
    r.slot = h->slot;
    r.far = h->pending;
    r.n = h->pending_n;
    h->pending_n = 0;
    return combine(h->pool, &r);
}

/* diagnostics for the tests: rounds run and requests served by the combiner of the pool `inst` lives in */
void AecmB200_ScalarApiStats(void *inst, int64_t *rounds, int64_t *requests) {
    Handle *h = (Handle *)inst;
    if (!h) return;
    pthread_mutex_lock(&h->pool->q_lock);
    if (rounds) *rounds = h->pool->rounds;
    if (requests) *requests = h->pool->combined;
    pthread_mutex_unlock(&h->pool->q_lock);
}

int32_t WebRtcAecm_GetBufferFarendError(void *inst, const int16_t *farend, size_t n) { /* :195-213 */
    Handle *h = (Handle *)inst;
    if (!h) return -1;
    if (!farend) return AECM_NULL_POINTER_ERROR;
    if (!h->initialized) return AECM_UNINITIALIZED_ERROR;
    if (n != 80 && n != 160) return AECM_BAD_PARAMETER_ERROR;
    return 0;
}

int32_t WebRtcAecm_BufferFarend(void *inst, const int16_t *farend, size_t n) { /* :215-234 */
    Handle *h = (Handle *)inst;
    int32_t rc = WebRtcAecm_GetBufferFarendError(inst, farend, n); /* :221-225: the only ways this call can fail */
    if (rc) return rc;
    rc = flush_pending(h); /* a frame is already waiting (two BufferFarend in a row): it goes first, on its own */
    if (rc) return rc;
    memcpy(h->pending, farend, n * sizeof(int16_t)); /* rides with the next Process of this instance */
    h->pending_n = (int)n;
    return 0;
}

int32_t WebRtcAecm_Process(void *inst, const int16_t *near_noisy, const int16_t *near_clean, int16_t *out, size_t n,
                           int16_t ms) { /* :236-408 */
    Handle *h = (Handle *)inst;
    Req r;
    int32_t rc;
    if (!h) return -1;
    if (!near_noisy) return AECM_NULL_POINTER_ERROR;
    if (!out) return AECM_NULL_POINTER_ERROR;
    if (!h->initialized) return AECM_UNINITIALIZED_ERROR; /* :264-266 precedes :268-270 */
    if (n != 80 && n != 160) return AECM_BAD_PARAMETER_ERROR;
    memset(&r, 0, sizeof(r));
    r.slot = h->slot;
    if (h->pending_n == (int)n) {
        r.far = h->pending; /* BufferFarend + Process in one launch, in the reference's order */
    } else {
        rc = flush_pending(h); /* another frame size: a launch has one */
        if (rc) return rc;
    }
    r.near_noisy = near_noisy;
    r.near_clean = near_clean;
    r.out = out;
    r.n = (int)n;
    r.ms = ms;
    rc = combine(h->pool, &r);
    h->pending_n = 0;
    return rc;
}

int32_t WebRtcAecm_set_config(void *inst, AecmConfig config) { /* :410-479 */
    Handle *h = (Handle *)inst;
    int32_t status = 0, rc;
    if (!h) return -1;
    rc = flush_pending(h); /* keep the caller's order of operations on this instance */
    if (rc) return rc;
    pthread_mutex_lock(&h->pool->exec);
    rc = AecmB200_SetConfig(h->pool->batch, h->slot, 1, config, &status);
    pthread_mutex_unlock(&h->pool->exec);
    return rc ? rc : status;
}

int32_t WebRtcAecm_InitEchoPath(void *inst, const void *echo_path, size_t size_bytes) { /* :481-503 */
    Handle *h = (Handle *)inst;
    int32_t rc;
    if (!h) return -1;
    rc = flush_pending(h); /* keep the caller's order of operations on this instance */
    if (rc) return rc;
    pthread_mutex_lock(&h->pool->exec);
    rc = AecmB200_InitEchoPath(h->pool->batch, h->slot, echo_path, size_bytes);
    pthread_mutex_unlock(&h->pool->exec);
    return rc;
}

int32_t WebRtcAecm_GetEchoPath(void *inst, void *echo_path, size_t size_bytes) { /* :505-528 */
    Handle *h = (Handle *)inst;
    int32_t rc;
    if (!h) return -1;
    rc = flush_pending(h); /* keep the caller's order of operations on this instance */
    if (rc) return rc;
    pthread_mutex_lock(&h->pool->exec);
    rc = AecmB200_GetEchoPath(h->pool->batch, h->slot, echo_path, size_bytes);
    pthread_mutex_unlock(&h->pool->exec);
    return rc;
}
