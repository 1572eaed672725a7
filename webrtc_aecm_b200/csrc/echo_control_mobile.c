/*
 * echo_control_mobile.c -- the reference's ten scalar entry points (aecm/echo_control_mobile.h:46-202) as
 * thin views over slots of device-resident batches.  Handles are small host descriptors; the AECM state and
 * every bit of arithmetic live on the GPU (aecm_kernels.cu).  Error codes and their precedence follow
 * aecm/echo_control_mobile.cc line by line (cited below).
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "aecm_b200.h"

#define POOL_SLOTS 256
#define HANDLE_MAGIC 0x41454342u /* "AECB" */

typedef struct Pool {
    AecmB200Batch *batch;
    unsigned char used[POOL_SLOTS];
    int n_used;
    struct Pool *next;
} Pool;

typedef struct {
    uint32_t magic;
    Pool *pool;
    int32_t slot;
} Handle;

static Pool *g_pools = NULL;
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;

static int default_device(void) {
    const char *e = getenv("AECM_B200_DEVICE");
    return e ? atoi(e) : 0;
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
            p->batch = AecmB200_Create(POOL_SLOTS, default_device());
            if (!p->batch) {
                free(p);
                p = NULL;
            } else {
                p->next = g_pools;
                g_pools = p;
            }
        }
    }
    if (p) {
        for (s = 0; p->used[s]; s++) {}
        h = (Handle *)calloc(1, sizeof(Handle));
        /* a fresh instance is zero-filled and un-initialised, like the reference's calloc (:91) */
        if (h && AecmB200_Reset(p->batch, s, 1) != 0) {
            free(h);
            h = NULL;
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
    h->pool->n_used--;
    h->magic = 0;
    pthread_mutex_unlock(&g_lock);
    free(h);
}

int32_t WebRtcAecm_Init(void *inst, int32_t sampFreq) { /* echo_control_mobile.cc:142-191 */
    Handle *h = (Handle *)inst;
    int32_t rc;
    if (!h) return -1;
    if (sampFreq != 8000 && sampFreq != 16000) return AECM_BAD_PARAMETER_ERROR;
    rc = AecmB200_Init(h->pool->batch, h->slot, 1, sampFreq);
    if (rc) return rc;
    return AecmB200_Synchronize(h->pool->batch) ? AECM_UNSPECIFIED_ERROR : 0;
}

static int32_t one_call(Handle *h, const int16_t *far, const int16_t *near_noisy, const int16_t *near_clean, int16_t *out,
                        size_t n, int16_t ms) {
    AecmB200StreamArgs a;
    int32_t status = 0, rc;
    memset(&a, 0, sizeof(a));
    a.farend = far;
    a.nearendNoisy = near_noisy;
    a.nearendClean = near_clean;
    a.out = out;
    a.rowStride = (int64_t)n;
    a.first = h->slot;
    a.count = 1;
    a.nrOfSamples = (int32_t)n;
    a.nCalls = 1;
    a.msUniform = ms;
    a.status = &status;
    rc = AecmB200_ProcessStream(h->pool->batch, &a);
    return rc ? rc : status;
}

int32_t WebRtcAecm_GetBufferFarendError(void *inst, const int16_t *farend, size_t n) { /* :195-213 */
    Handle *h = (Handle *)inst;
    int16_t path[65];
    if (!h) return -1;
    if (!farend) return AECM_NULL_POINTER_ERROR;
    /* initFlag lives on the device: GetEchoPath reports AECM_UNINITIALIZED_ERROR exactly when it is unset */
    if (AecmB200_GetEchoPath(h->pool->batch, h->slot, path, sizeof(path)) == AECM_UNINITIALIZED_ERROR)
        return AECM_UNINITIALIZED_ERROR;
    if (n != 80 && n != 160) return AECM_BAD_PARAMETER_ERROR;
    return 0;
}

int32_t WebRtcAecm_BufferFarend(void *inst, const int16_t *farend, size_t n) { /* :215-234 */
    Handle *h = (Handle *)inst;
    if (!h) return -1;
    if (!farend) return AECM_NULL_POINTER_ERROR;
    if (n != 80 && n != 160) {
        /* precedence: the reference reports UNINITIALIZED before BAD_PARAMETER (:205-210) */
        const int32_t e = WebRtcAecm_GetBufferFarendError(inst, farend, n);
        return e ? e : AECM_BAD_PARAMETER_ERROR;
    }
    return one_call(h, farend, NULL, NULL, NULL, n, 0);
}

int32_t WebRtcAecm_Process(void *inst, const int16_t *near_noisy, const int16_t *near_clean, int16_t *out, size_t n,
                           int16_t ms) { /* :236-408 */
    Handle *h = (Handle *)inst;
    if (!h) return -1;
    if (!near_noisy) return AECM_NULL_POINTER_ERROR;
    if (!out) return AECM_NULL_POINTER_ERROR;
    if (n != 80 && n != 160) {
        int16_t path[65];
        if (AecmB200_GetEchoPath(h->pool->batch, h->slot, path, sizeof(path)) == AECM_UNINITIALIZED_ERROR)
            return AECM_UNINITIALIZED_ERROR; /* :264-266 precedes :268-270 */
        return AECM_BAD_PARAMETER_ERROR;
    }
    return one_call(h, NULL, near_noisy, near_clean, out, n, ms);
}

int32_t WebRtcAecm_set_config(void *inst, AecmConfig config) { /* :410-479 */
    Handle *h = (Handle *)inst;
    int32_t status = 0, rc;
    if (!h) return -1;
    rc = AecmB200_SetConfig(h->pool->batch, h->slot, 1, config, &status);
    return rc ? rc : status;
}

int32_t WebRtcAecm_InitEchoPath(void *inst, const void *echo_path, size_t size_bytes) { /* :481-503 */
    Handle *h = (Handle *)inst;
    if (!h) return -1;
    return AecmB200_InitEchoPath(h->pool->batch, h->slot, echo_path, size_bytes);
}

int32_t WebRtcAecm_GetEchoPath(void *inst, void *echo_path, size_t size_bytes) { /* :505-528 */
    Handle *h = (Handle *)inst;
    if (!h) return -1;
    return AecmB200_GetEchoPath(h->pool->batch, h->slot, echo_path, size_bytes);
}
