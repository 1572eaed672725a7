/*
 * fake_batch.c -- TEST INFRASTRUCTURE.  The subset of the batched ABI (include/aecm_b200.h) that the product's two host
 * programs above the batch call -- the scalar-API layer csrc/echo_control_mobile.c and the WAV tool csrc/aecm_wav_batch.c --
 * implemented on the CPU oracle, so that their host logic (handle pools, the request combiner with its leader election and
 * hand-over, the deferred far end, pool teardown; reader/writer threads, sharding, lockstep chunks, the Submit/Wait
 * pipeline) can run under ThreadSanitizer / AddressSanitizer on a machine without a GPU
 * (tests/test_scalar_combiner_tsan.py, tests/test_wav_batch_on_stand_in.py).  Nothing here is shipped or measured; the
 * product's batch is csrc/aecm_batch.c on CUDA.
 *
 * Semantics kept from the real batch: a call is applied to rows in lockstep; status[r] = first non-zero return code of
 * the row's calls; Submit on a busy slot is refused; Wait on an idle slot is a no-op.  Like the real batch, this one is
 * NOT thread safe: the layer under test must serialise its callers, and the `in_call` guard below aborts if it does not.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "aecm_b200.h"
#include "../../oracle/aecm_oracle.h"

#define FAKE_SLOTS 4

struct AecmB200Batch {
    int32_t n;
    OracleAecm **inst;
    int busy[FAKE_SLOTS];
    volatile int in_call; /* deliberately unsynchronised: two callers inside the batch at once are the bug to catch (TSan sees it too) */
};

static long g_created, g_freed;
long fake_batches_created(void) { return g_created; }
long fake_batches_freed(void) { return g_freed; }

static void enter(AecmB200Batch *b) {
    if (b->in_call) {
        fprintf(stderr, "fake_batch: two callers inside one batch\n");
        abort();
    }
    b->in_call = 1;
}
static void leave(AecmB200Batch *b) { b->in_call = 0; }

int32_t AecmB200_DeviceCount(void) { return 1; }
int32_t AecmB200_NumStreamSlots(void) { return FAKE_SLOTS; }
void *AecmB200_HostAlloc(size_t bytes) { return malloc(bytes); }
void AecmB200_HostFree(void *p) { free(p); }

AecmB200Batch *AecmB200_Create(int32_t n, int32_t device) {
    AecmB200Batch *b = (AecmB200Batch *)calloc(1, sizeof(*b));
    (void)device;
    b->n = n;
    b->inst = (OracleAecm **)calloc((size_t)n, sizeof(OracleAecm *));
    for (int i = 0; i < n; i++) b->inst[i] = oracle_aecm_create();
    __atomic_add_fetch(&g_created, 1, __ATOMIC_RELAXED);
    return b;
}

void AecmB200_Free(AecmB200Batch *b) {
    if (!b) return;
    for (int i = 0; i < b->n; i++) oracle_aecm_free(b->inst[i]);
    free(b->inst);
    free(b);
    __atomic_add_fetch(&g_freed, 1, __ATOMIC_RELAXED);
}

static int range_ok(const AecmB200Batch *b, int32_t first, int32_t count) {
    return first >= 0 && count >= 0 && (int64_t)first + count <= b->n;
}

int32_t AecmB200_Reset(AecmB200Batch *b, int32_t first, int32_t count) {
    if (!range_ok(b, first, count)) return AECM_BAD_PARAMETER_ERROR;
    enter(b);
    for (int i = first; i < first + count; i++) {
        oracle_aecm_free(b->inst[i]);
        b->inst[i] = oracle_aecm_create();
    }
    leave(b);
    return 0;
}

int32_t AecmB200_Init(AecmB200Batch *b, int32_t first, int32_t count, int32_t fs) {
    if (fs != 8000 && fs != 16000) return AECM_BAD_PARAMETER_ERROR;
    if (!range_ok(b, first, count)) return AECM_BAD_PARAMETER_ERROR;
    enter(b);
    for (int i = first; i < first + count; i++) oracle_aecm_init(b->inst[i], fs);
    leave(b);
    return 0;
}

int32_t AecmB200_Synchronize(AecmB200Batch *b) {
    enter(b);
    leave(b);
    return 0;
}

int32_t AecmB200_SetConfig(AecmB200Batch *b, int32_t first, int32_t count, AecmConfig cfg, int32_t *status) {
    if (!range_ok(b, first, count)) return AECM_BAD_PARAMETER_ERROR;
    enter(b);
    for (int i = 0; i < count; i++) {
        const int32_t r = oracle_aecm_set_config(b->inst[first + i], cfg.cngMode, cfg.echoMode);
        if (status) status[i] = r;
    }
    leave(b);
    return 0;
}

int32_t AecmB200_InitEchoPath(AecmB200Batch *b, int32_t instance, const void *path, size_t bytes) {
    int32_t r;
    if (!path) return AECM_NULL_POINTER_ERROR;
    if (bytes != oracle_aecm_echo_path_size_bytes()) return AECM_BAD_PARAMETER_ERROR;
    if (!range_ok(b, instance, 1)) return AECM_BAD_PARAMETER_ERROR;
    enter(b);
    r = oracle_aecm_init_echo_path(b->inst[instance], path, bytes);
    leave(b);
    return r;
}

int32_t AecmB200_GetEchoPath(AecmB200Batch *b, int32_t instance, void *path, size_t bytes) {
    int32_t r;
    if (!path) return AECM_NULL_POINTER_ERROR;
    if (bytes != oracle_aecm_echo_path_size_bytes()) return AECM_BAD_PARAMETER_ERROR;
    if (!range_ok(b, instance, 1)) return AECM_BAD_PARAMETER_ERROR;
    enter(b);
    r = oracle_aecm_get_echo_path(b->inst[instance], path, bytes);
    leave(b);
    return r;
}

/* the whole submission runs inside Submit; Wait only releases the slot */
int32_t AecmB200_SubmitStream(AecmB200Batch *b, const AecmB200StreamArgs *a, int32_t slot) {
    if (slot < 0 || slot >= FAKE_SLOTS || !a) return AECM_BAD_PARAMETER_ERROR;
    if (a->nrOfSamples != 80 && a->nrOfSamples != 160) return AECM_BAD_PARAMETER_ERROR;
    if (b->busy[slot]) return AECM_BAD_PARAMETER_ERROR;
    enter(b);
    for (int r = 0; r < a->count; r++) { /* instanceIndex: distinct and in range, as the real batch checks */
        const int32_t i = a->instanceIndex ? a->instanceIndex[r] : a->first + r;
        if (i < 0 || i >= b->n) { leave(b); return AECM_BAD_PARAMETER_ERROR; }
        for (int q = 0; a->instanceIndex && q < r; q++)
            if (a->instanceIndex[q] == i) { leave(b); return AECM_BAD_PARAMETER_ERROR; }
    }
    for (int r = 0; r < a->count; r++) {
        OracleAecm *o = b->inst[a->instanceIndex ? a->instanceIndex[r] : a->first + r];
        int32_t st = 0;
        for (int c = 0; c < a->nCalls; c++) {
            const size_t off = (size_t)r * (size_t)a->rowStride + (size_t)c * (size_t)a->nrOfSamples;
            int32_t rc;
            if (a->farend && (!a->farendMask || a->farendMask[(size_t)c * a->count + r])) {
                rc = oracle_aecm_buffer_farend(o, a->farend + off, (size_t)a->nrOfSamples);
                if (!st) st = rc;
            }
            if (a->nearendNoisy) {
                rc = oracle_aecm_process(o, a->nearendNoisy + off, a->nearendClean ? a->nearendClean + off : NULL, a->out + off,
                                         (size_t)a->nrOfSamples, a->msInSndCardBuf ? a->msInSndCardBuf[r] : a->msUniform);
                if (!st) st = rc;
            }
        }
        if (a->status) a->status[r] = st;
    }
    b->busy[slot] = 1;
    leave(b);
    return 0;
}

int32_t AecmB200_WaitStream(AecmB200Batch *b, int32_t slot) {
    if (slot < 0 || slot >= FAKE_SLOTS) return AECM_BAD_PARAMETER_ERROR;
    enter(b);
    b->busy[slot] = 0;
    leave(b);
    return 0;
}

const char *AecmB200_LastError(const AecmB200Batch *b) {
    (void)b;
    return "stand-in batch: no device errors";
}
