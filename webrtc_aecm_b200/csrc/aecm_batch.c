/*
 * aecm_batch.c -- host side (plain C) of the batched AECM engine: device arenas, streams, argument
 * validation and host<->device staging around the CUDA launchers of aecm_kernels.cu.  There is no CPU
 * implementation of the algorithm here: if CUDA is unavailable every entry point fails loudly.
 */
#include "aecm_b200.h"

#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "aecm_state.h"

#define AECM_SLOTS 4 /* in-flight host-plane submissions: H2D, kernel and D2H of different calls overlap (a call holds
                        its slot for H2D + kernel + D2H, so keeping the H2D engine busy takes more than two) */

/* One in-flight host-plane submission: device staging planes, a copy stream and the events that order
 * H2D -> kernel (on the batch's compute stream, so kernels stay in submission order) -> D2H. */
typedef struct {
    int16_t *d_far, *d_near, *d_clean, *d_out;
    size_t plane_samples;
    int16_t *d_ms;
    int32_t *d_index;
    uint8_t *d_mask;
    size_t mask_bytes;
    int32_t *d_status, *h_status; /* h_status is pinned */
    cudaStream_t copy;      /* the slot's own stream: H2D -> kernel -> D2H all run on it */
    cudaEvent_t h2d_done, kernel_done;
    cudaEvent_t t[5];       /* optional timing marks (AECM_B200_TRACE=1): start, after H2D, kernel start, kernel end, end */
    int32_t *user_status;
    int32_t count;
    int busy;
} Slot;

struct AecmB200Batch {
    int32_t n, device;
    AecmBatchPlanes planes;
    cudaStream_t stream;
    int owns_stream;
    Slot slot[AECM_SLOTS];  /* staging for the host-plane calls, grown on demand */
    cudaEvent_t last_kernel; /* kernel_done of the most recent submission (kernels must run in submission order) */
    cudaEvent_t ev_main;
    int main_dirty;          /* the main stream received work that later submissions must follow */
    int slots_dirty;         /* a submission's kernel has not yet been ordered before later main-stream work */
    int trace;
    int32_t *d_status;      /* scratch for the control calls (set_config, echo path) */
    int16_t *d_path;
    int ext_active;         /* the delay estimator's optional machinery is in use: launches run the kExt kernel */
    float *d_quality;       /* scratch for AecmB200_GetDelayEstimate */
    uint64_t *index_seen;   /* bitmap over the instances, for the duplicate check of instanceIndex (allocated on first use) */
    char err[256];
};

static int use_main_stream(AecmB200Batch *b);

static int fail(AecmB200Batch *b, const char *what, cudaError_t e) {
    if (b) snprintf(b->err, sizeof(b->err), "%s: %s", what, cudaGetErrorString(e));
    return AECM_UNSPECIFIED_ERROR;
}
#define CU(b, call)                                         \
    do {                                                    \
        cudaError_t e_ = (call);                            \
        if (e_ != cudaSuccess) return fail((b), #call, e_); \
    } while (0)

/* Every entry point makes the batch's device current for its own CUDA calls and puts the caller's device back on the way
 * out (all return paths: the scope object's cleanup handler runs when the function leaves), so a host application that
 * uses CUDA on another device is not silently switched. */
typedef struct { int prev; } DeviceScope;
static void device_scope_exit(DeviceScope *s) {
    if (s->prev >= 0) cudaSetDevice(s->prev);
}
static cudaError_t device_scope_enter(DeviceScope *s, int device) {
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur == device) return cudaSuccess; /* already there: nothing to restore */
    s->prev = cur;
    return cudaSetDevice(device);
}
#define ENTER_DEVICE(b)                                                           \
    DeviceScope scope_ __attribute__((cleanup(device_scope_exit))) = {-1};        \
    CU((b), device_scope_enter(&scope_, (b)->device))

int32_t AecmB200_DeviceCount(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

AecmB200Batch *AecmB200_Create(int32_t n, int32_t device) {
    AecmB200Batch *b;
    DeviceScope scope_ __attribute__((cleanup(device_scope_exit))) = {-1};
    if (n <= 0) return NULL;
    if (device_scope_enter(&scope_, device) != cudaSuccess) {
        fprintf(stderr, "aecm_b200: CUDA device %d unavailable (%s); this library has no CPU path\n", device,
                cudaGetErrorString(cudaGetLastError()));
        return NULL;
    }
    b = (AecmB200Batch *)calloc(1, sizeof(*b));
    if (!b) return NULL;
    b->n = n;
    b->device = device;
    b->planes.n_instances = n;
    if (cudaMalloc((void **)&b->planes.state, (size_t)n * sizeof(AecmInstanceState)) != cudaSuccess ||
        cudaMalloc((void **)&b->planes.far_hist, (size_t)n * AECM_MAX_DELAY * AECM_HIST_ROW * sizeof(uint16_t)) != cudaSuccess ||
        cudaMalloc((void **)&b->planes.far_ring, (size_t)n * AECM_FAR_RING * sizeof(int16_t)) != cudaSuccess ||
        cudaMalloc((void **)&b->planes.dly_ext, (size_t)n * sizeof(AecmDelayExt)) != cudaSuccess ||
        cudaMalloc((void **)&b->d_quality, (size_t)n * sizeof(float)) != cudaSuccess ||
        cudaMalloc((void **)&b->d_status, (size_t)n * sizeof(int32_t)) != cudaSuccess ||
        cudaMalloc((void **)&b->d_path, AECM_BINS * sizeof(int16_t)) != cudaSuccess ||
        cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess) {
        fprintf(stderr, "aecm_b200: device allocation failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        AecmB200_Free(b);
        return NULL;
    }
    for (int i = 0; i < AECM_SLOTS; i++) {
        Slot *sl = &b->slot[i];
        if (cudaMalloc((void **)&sl->d_ms, (size_t)n * sizeof(int16_t)) != cudaSuccess ||
            cudaMalloc((void **)&sl->d_index, (size_t)n * sizeof(int32_t)) != cudaSuccess ||
            cudaMalloc((void **)&sl->d_status, (size_t)n * sizeof(int32_t)) != cudaSuccess ||
            cudaMallocHost((void **)&sl->h_status, (size_t)n * sizeof(int32_t)) != cudaSuccess ||
            cudaStreamCreateWithFlags(&sl->copy, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&sl->h2d_done, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&sl->kernel_done, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreate(&sl->t[0]) != cudaSuccess || cudaEventCreate(&sl->t[1]) != cudaSuccess ||
            cudaEventCreate(&sl->t[2]) != cudaSuccess || cudaEventCreate(&sl->t[3]) != cudaSuccess ||
            cudaEventCreate(&sl->t[4]) != cudaSuccess) {
            fprintf(stderr, "aecm_b200: staging allocation failed: %s\n", cudaGetErrorString(cudaGetLastError()));
            AecmB200_Free(b);
            return NULL;
        }
    }
    if (cudaEventCreateWithFlags(&b->ev_main, cudaEventDisableTiming) != cudaSuccess) {
        AecmB200_Free(b);
        return NULL;
    }
    b->trace = getenv("AECM_B200_TRACE") != NULL;
    b->owns_stream = 1;
    /* calloc semantics of the reference's Create: an un-Init'ed instance reports AECM_UNINITIALIZED_ERROR */
    if (cudaMemsetAsync(b->planes.state, 0, (size_t)n * sizeof(AecmInstanceState), b->stream) != cudaSuccess ||
        cudaMemsetAsync(b->planes.dly_ext, 0, (size_t)n * sizeof(AecmDelayExt), b->stream) != cudaSuccess ||
        cudaStreamSynchronize(b->stream) != cudaSuccess) {
        AecmB200_Free(b);
        return NULL;
    }
    return b;
}

void AecmB200_Free(AecmB200Batch *b) {
    DeviceScope scope_ __attribute__((cleanup(device_scope_exit))) = {-1};
    if (!b) return;
    device_scope_enter(&scope_, b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    cudaFree(b->planes.state);
    cudaFree(b->planes.far_hist);
    cudaFree(b->planes.far_ring);
    cudaFree(b->planes.dly_ext);
    cudaFree(b->d_quality);
    for (int i = 0; i < AECM_SLOTS; i++) {
        Slot *sl = &b->slot[i];
        if (sl->copy) cudaStreamSynchronize(sl->copy);
        cudaFree(sl->d_far); cudaFree(sl->d_clean); cudaFree(sl->d_out); /* d_near lives inside d_far's allocation */
        cudaFree(sl->d_ms); cudaFree(sl->d_index); cudaFree(sl->d_mask); cudaFree(sl->d_status);
        cudaFreeHost(sl->h_status);
        if (sl->copy) cudaStreamDestroy(sl->copy);
        if (sl->h2d_done) cudaEventDestroy(sl->h2d_done);
        if (sl->kernel_done) cudaEventDestroy(sl->kernel_done);
        for (int j = 0; j < 5; j++)
            if (sl->t[j]) cudaEventDestroy(sl->t[j]);
    }
    if (b->ev_main) cudaEventDestroy(b->ev_main);
    cudaFree(b->d_status);
    cudaFree(b->d_path);
    if (b->stream && b->owns_stream) cudaStreamDestroy(b->stream);
    free(b->index_seen);
    free(b);
}

int32_t AecmB200_NumInstances(const AecmB200Batch *b) { return b ? b->n : -1; }
int32_t AecmB200_Device(const AecmB200Batch *b) { return b ? b->device : -1; }
const char *AecmB200_LastError(const AecmB200Batch *b) { return b ? b->err : "null batch"; }
void *AecmB200_CudaStream(const AecmB200Batch *b) { return b ? (void *)b->stream : NULL; }

int32_t AecmB200_AdoptCudaStream(AecmB200Batch *b, void *s) {
    if (!b) return -1;
    ENTER_DEVICE(b);
    CU(b, cudaStreamSynchronize(b->stream));
    if (b->owns_stream) cudaStreamDestroy(b->stream);
    b->stream = (cudaStream_t)s;
    b->owns_stream = 0;
    return 0;
}

int32_t AecmB200_Synchronize(AecmB200Batch *b) {
    if (!b) return -1;
    ENTER_DEVICE(b);
    if (use_main_stream(b)) return AECM_UNSPECIFIED_ERROR;
    CU(b, cudaStreamSynchronize(b->stream));
    return 0;
}

void *AecmB200_HostAlloc(size_t bytes) {
    void *p = NULL;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return NULL;
    return p;
}
void AecmB200_HostFree(void *p) {
    if (p) cudaFreeHost(p);
}

size_t AecmB200_StateBytes(void) { return sizeof(AecmInstanceState); }
int64_t AecmB200_KernelLaunchCount(void) { return aecm_cuda_kernel_launch_count(); }
int32_t AecmB200_NumStreamSlots(void) { return AECM_SLOTS; }
int64_t AecmB200_SelfTestSqrtFloor(uint32_t first, uint32_t count) { return aecm_cuda_selftest_sqrt(first, count); }
int32_t AecmB200_SelfTestFftForward(const int16_t *in, int16_t *out, int32_t n, int32_t twoSignals) {
    if (!in || !out) return AECM_NULL_POINTER_ERROR;
    if (n < 0) return AECM_BAD_PARAMETER_ERROR;
    return aecm_cuda_selftest_fft(twoSignals ? 1 : 0, in, out, NULL, n) == 0 ? 0 : AECM_UNSPECIFIED_ERROR;
}
int32_t AecmB200_SelfTestFftInverse(const int16_t *in, int16_t *out, int32_t *scale, int32_t n) {
    if (!in || !out || !scale) return AECM_NULL_POINTER_ERROR;
    if (n < 0) return AECM_BAD_PARAMETER_ERROR;
    return aecm_cuda_selftest_fft(2, in, out, scale, n) == 0 ? 0 : AECM_UNSPECIFIED_ERROR;
}

/* Called before enqueuing anything on the main stream: order it after the kernels of earlier submissions and
 * remember that later submissions have to follow it. */
static int use_main_stream(AecmB200Batch *b) {
    if (b->slots_dirty && b->last_kernel) {
        CU(b, cudaStreamWaitEvent(b->stream, b->last_kernel, 0));
        b->slots_dirty = 0;
    }
    b->main_dirty = 1;
    return 0;
}

static int range_ok(const AecmB200Batch *b, int32_t first, int32_t count) {
    return first >= 0 && count >= 0 && (int64_t)first + count <= b->n;
}

int32_t AecmB200_ReadState(AecmB200Batch *b, int32_t instance, void *dst, size_t bytes) {
    if (!b) return -1;
    if (!dst) return AECM_NULL_POINTER_ERROR;
    if (!range_ok(b, instance, 1) || bytes > sizeof(AecmInstanceState)) return AECM_BAD_PARAMETER_ERROR;
    ENTER_DEVICE(b);
    if (use_main_stream(b)) return AECM_UNSPECIFIED_ERROR;
    CU(b, cudaMemcpyAsync(dst, b->planes.state + instance, bytes, cudaMemcpyDeviceToHost, b->stream));
    CU(b, cudaStreamSynchronize(b->stream));
    return 0;
}

int32_t AecmB200_Reset(AecmB200Batch *b, int32_t first, int32_t count) {
    if (!b) return -1;
    if (!range_ok(b, first, count)) return AECM_BAD_PARAMETER_ERROR;
    ENTER_DEVICE(b);
    if (use_main_stream(b)) return AECM_UNSPECIFIED_ERROR;
    CU(b, cudaMemsetAsync(b->planes.state + first, 0, (size_t)count * sizeof(AecmInstanceState), b->stream));
    CU(b, cudaMemsetAsync(b->planes.dly_ext + first, 0, (size_t)count * sizeof(AecmDelayExt), b->stream));
    CU(b, cudaStreamSynchronize(b->stream));
    return 0;
}

int32_t AecmB200_Init(AecmB200Batch *b, int32_t first, int32_t count, int32_t fs) {
    if (!b) return -1;
    if (fs != 8000 && fs != 16000) return AECM_BAD_PARAMETER_ERROR; /* echo_control_mobile.cc:150-152 */
    if (!range_ok(b, first, count)) return AECM_BAD_PARAMETER_ERROR;
    ENTER_DEVICE(b);
    if (use_main_stream(b)) return AECM_UNSPECIFIED_ERROR;
    CU(b, (cudaError_t)aecm_cuda_launch_init(&b->planes, first, count, fs, b->stream));
    return 0;
}

static int copy_status_out(AecmB200Batch *b, int32_t *status, int32_t count) {
    if (status) CU(b, cudaMemcpyAsync(status, b->d_status, (size_t)count * sizeof(int32_t), cudaMemcpyDeviceToHost, b->stream));
    CU(b, cudaStreamSynchronize(b->stream));
    return 0;
}

int32_t AecmB200_SetConfig(AecmB200Batch *b, int32_t first, int32_t count, AecmConfig cfg, int32_t *status) {
    if (!b) return -1;
    if (!range_ok(b, first, count)) return AECM_BAD_PARAMETER_ERROR;
    ENTER_DEVICE(b);
    if (use_main_stream(b)) return AECM_UNSPECIFIED_ERROR;
    CU(b, (cudaError_t)aecm_cuda_launch_set_config(&b->planes, first, count, cfg.cngMode, cfg.echoMode, b->d_status, b->stream));
    return copy_status_out(b, status, count);
}

size_t WebRtcAecm_echo_path_size_bytes(void) { return AECM_BINS * sizeof(int16_t); } /* echo_control_mobile.cc:530-532 */

int32_t AecmB200_InitEchoPath(AecmB200Batch *b, int32_t instance, const void *path, size_t size_bytes) {
    int32_t st = 0, rc;
    if (!b) return -1;
    if (!path) return AECM_NULL_POINTER_ERROR;
    if (size_bytes != WebRtcAecm_echo_path_size_bytes()) return AECM_BAD_PARAMETER_ERROR;
    if (!range_ok(b, instance, 1)) return AECM_BAD_PARAMETER_ERROR;
    ENTER_DEVICE(b);
    if (use_main_stream(b)) return AECM_UNSPECIFIED_ERROR;
    CU(b, cudaMemcpyAsync(b->d_path, path, size_bytes, cudaMemcpyHostToDevice, b->stream));
    CU(b, (cudaError_t)aecm_cuda_launch_init_echo_path(&b->planes, instance, 1, b->d_path, AECM_BINS, b->d_status, b->stream));
    rc = copy_status_out(b, &st, 1);
    return rc ? rc : st;
}

int32_t AecmB200_GetEchoPath(AecmB200Batch *b, int32_t instance, void *path, size_t size_bytes) {
    int32_t flag = 0;
    const char *base;
    if (!b) return -1;
    if (!path) return AECM_NULL_POINTER_ERROR;
    if (size_bytes != WebRtcAecm_echo_path_size_bytes()) return AECM_BAD_PARAMETER_ERROR;
    if (!range_ok(b, instance, 1)) return AECM_BAD_PARAMETER_ERROR;
    ENTER_DEVICE(b);
    if (use_main_stream(b)) return AECM_UNSPECIFIED_ERROR;
    base = (const char *)(b->planes.state + instance);
    CU(b, cudaMemcpyAsync(&flag, base + offsetof(AecmInstanceState, s) + offsetof(AecmScalars, init_flag), sizeof(flag),
                          cudaMemcpyDeviceToHost, b->stream));
    CU(b, cudaStreamSynchronize(b->stream));
    if (flag != AECM_INIT_MAGIC) return AECM_UNINITIALIZED_ERROR;
    CU(b, cudaMemcpyAsync(path, base + offsetof(AecmInstanceState, ch_stored), size_bytes, cudaMemcpyDeviceToHost, b->stream));
    CU(b, cudaStreamSynchronize(b->stream));
    return 0;
}

/* ---- the delay estimator's optional machinery (robust validation, lookahead, soft reset) ---- */
int32_t AecmB200_SetDelayEstimatorOptions(AecmB200Batch *b, int32_t first, int32_t count, const AecmB200DelayEstimatorOptions *o,
                                          int32_t *status) {
    if (!b) return -1;
    if (!o) return AECM_NULL_POINTER_ERROR;
    if (!range_ok(b, first, count)) return AECM_BAD_PARAMETER_ERROR;
    /* what the reference's setters reject for any estimator (delay_estimator_wrapper.cc:401-434) or its Create cannot hold */
    if ((o->robustValidation != 0 && o->robustValidation != 1) || o->allowedOffset < 0 || o->maxLookahead > AECM_MAX_LOOKAHEAD ||
        (o->maxLookahead >= 0 && o->lookahead > o->maxLookahead))
        return AECM_BAD_PARAMETER_ERROR;
    ENTER_DEVICE(b);
    if (use_main_stream(b)) return AECM_UNSPECIFIED_ERROR;
    b->ext_active = 1;
    CU(b, (cudaError_t)aecm_cuda_launch_dly_options(&b->planes, first, count, o->robustValidation, o->allowedOffset, o->maxLookahead,
                                                    o->lookahead, b->d_status, b->stream));
    return copy_status_out(b, status, count);
}

int32_t AecmB200_SoftResetDelayEstimator(AecmB200Batch *b, int32_t first, int32_t count, int32_t delayShift, int32_t *applied) {
    if (!b) return -1;
    if (!range_ok(b, first, count)) return AECM_BAD_PARAMETER_ERROR;
    if (delayShift <= -AECM_MAX_DELAY || delayShift >= AECM_MAX_DELAY) return AECM_BAD_PARAMETER_ERROR; /* delay_estimator.cc:346-347 */
    ENTER_DEVICE(b);
    if (use_main_stream(b)) return AECM_UNSPECIFIED_ERROR;
    b->ext_active = 1;
    CU(b, (cudaError_t)aecm_cuda_launch_dly_soft_reset(&b->planes, first, count, delayShift, b->d_status, b->stream));
    return copy_status_out(b, applied, count);
}

int32_t AecmB200_GetDelayEstimate(AecmB200Batch *b, int32_t first, int32_t count, int32_t *lastDelay, float *quality) {
    if (!b) return -1;
    if (!range_ok(b, first, count)) return AECM_BAD_PARAMETER_ERROR;
    ENTER_DEVICE(b);
    if (use_main_stream(b)) return AECM_UNSPECIFIED_ERROR;
    CU(b, (cudaError_t)aecm_cuda_launch_dly_query(&b->planes, first, count, b->d_status, b->d_quality, b->stream));
    if (quality) CU(b, cudaMemcpyAsync(quality, b->d_quality, (size_t)count * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
    return copy_status_out(b, lastDelay, count);
}

static int32_t validate_stream_args(const AecmB200Batch *b, const AecmB200StreamArgs *a) {
    if (!a) return AECM_NULL_POINTER_ERROR;
    if (!a->farend && !a->nearendNoisy) return AECM_NULL_POINTER_ERROR;
    if (a->nearendNoisy && !a->out) return AECM_NULL_POINTER_ERROR;            /* echo_control_mobile.cc:256-262 */
    if (a->nrOfSamples != 80 && a->nrOfSamples != 160) return AECM_BAD_PARAMETER_ERROR; /* :209, :268 */
    if (a->nCalls < 1) return AECM_BAD_PARAMETER_ERROR;
    if (a->instanceIndex ? (a->count < 0 || a->count > b->n) : !range_ok(b, a->first, a->count)) return AECM_BAD_PARAMETER_ERROR;
    if (a->rowStride < (int64_t)a->nCalls * a->nrOfSamples) return AECM_BAD_PARAMETER_ERROR;
    return 0;
}

static void fill_call_args(AecmCallArgs *k, const AecmB200StreamArgs *a) {
    memset(k, 0, sizeof(*k));
    k->far = a->farend;
    k->near_noisy = a->nearendNoisy;
    k->near_clean = a->nearendNoisy ? a->nearendClean : NULL;
    k->out = a->out;
    k->ms = a->msInSndCardBuf;
    k->far_mask = a->farendMask;
    k->status = a->status;
    k->row_stride = a->rowStride;
    k->first = a->first;
    k->count = a->count;
    k->samples_per_call = a->nrOfSamples;
    k->n_calls = a->nCalls;
    k->ms_uniform = a->msUniform;
    k->mask_stride = a->count;
    k->index = a->instanceIndex;
}

int32_t AecmB200_ProcessStreamDevice(AecmB200Batch *b, const AecmB200StreamArgs *a) {
    AecmCallArgs k;
    int32_t rc;
    if (!b) return -1;
    rc = validate_stream_args(b, a);
    if (rc) return rc;
    if (a->count == 0) return 0;
    ENTER_DEVICE(b);
    if (use_main_stream(b)) return AECM_UNSPECIFIED_ERROR;
    fill_call_args(&k, a);
    k.ext = b->ext_active;
    CU(b, (cudaError_t)aecm_cuda_launch_call(&b->planes, &k, b->stream));
    return 0;
}

/* The far and near staging planes are carved out of ONE allocation, far first, so that a caller who keeps its own
 * far and near planes adjacent in host memory gets them across PCIe in a single DMA (each transfer costs a fixed
 * ~10 us on top of its bytes). */
static int grow_planes(AecmB200Batch *b, Slot *sl, size_t samples, int need_clean) {
    if (samples > sl->plane_samples) {
        cudaFree(sl->d_far); cudaFree(sl->d_clean); cudaFree(sl->d_out);
        sl->d_far = sl->d_near = sl->d_clean = sl->d_out = NULL;
        sl->plane_samples = 0;
        CU(b, cudaMalloc((void **)&sl->d_far, 2 * samples * sizeof(int16_t)));
        CU(b, cudaMalloc((void **)&sl->d_out, samples * sizeof(int16_t)));
        sl->plane_samples = samples;
    }
    if (need_clean && !sl->d_clean) CU(b, cudaMalloc((void **)&sl->d_clean, sl->plane_samples * sizeof(int16_t)));
    return 0;
}

/* plane copy: one flat DMA when the host rows are dense, a pitched copy otherwise */
static cudaError_t copy_plane(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows,
                              enum cudaMemcpyKind kind, cudaStream_t s) {
    if (dpitch == width && spitch == width) return cudaMemcpyAsync(dst, src, width * rows, kind, s);
    return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, kind, s);
}

/* Host planes, asynchronous.  H2D on the slot's copy stream, the kernel on the batch's compute stream (so kernels of
 * successive submissions run in order: every call depends on the state left by the previous one), D2H back on the
 * copy stream.  With two slots the copies of call k+1 overlap the kernel of call k. */
static int32_t enqueue_stream(AecmB200Batch *b, const AecmB200StreamArgs *h, Slot *sl);

/* rows must name distinct instances: two warps working on one record would race (a host array here; the device array of
 * AecmB200_ProcessStreamDevice is the caller's responsibility, the kernel only range-checks it) */
static int32_t check_instance_index(AecmB200Batch *b, const int32_t *index, int32_t count) {
    const size_t words = ((size_t)b->n + 63) / 64;
    int32_t r, rc = 0;
    if (!b->index_seen) {
        b->index_seen = (uint64_t *)malloc(words * sizeof(uint64_t));
        if (!b->index_seen) return AECM_UNSPECIFIED_ERROR;
    }
    memset(b->index_seen, 0, words * sizeof(uint64_t));
    for (r = 0; r < count && !rc; r++) {
        const int32_t i = index[r];
        if (i < 0 || i >= b->n) rc = AECM_BAD_PARAMETER_ERROR;
        else if (b->index_seen[i >> 6] & (1ull << (i & 63))) rc = AECM_BAD_PARAMETER_ERROR; /* duplicate */
        else b->index_seen[i >> 6] |= 1ull << (i & 63);
    }
    return rc;
}

int32_t AecmB200_SubmitStream(AecmB200Batch *b, const AecmB200StreamArgs *h, int32_t slot) {
    Slot *sl;
    int32_t rc;
    if (!b) return -1;
    if (slot < 0 || slot >= AECM_SLOTS) return AECM_BAD_PARAMETER_ERROR;
    rc = validate_stream_args(b, h);
    if (rc) return rc;
    if (h->instanceIndex) {
        rc = check_instance_index(b, h->instanceIndex, h->count);
        if (rc) return rc;
    }
    sl = &b->slot[slot];
    if (sl->busy) return AECM_BAD_PARAMETER_ERROR; /* the previous submission on this slot was not waited for */
    ENTER_DEVICE(b);
    sl->user_status = NULL;
    sl->count = 0;
    if (h->count > 0) {
        rc = enqueue_stream(b, h, sl);
        if (rc) { /* nothing of a failed submission may linger: drain what was enqueued, leave the slot free */
            cudaStreamSynchronize(sl->copy);
            return rc;
        }
        sl->user_status = h->status;
        sl->count = h->count;
    }
    sl->busy = 1; /* only a fully enqueued submission occupies the slot */
    return 0;
}

static int32_t enqueue_stream(AecmB200Batch *b, const AecmB200StreamArgs *h, Slot *sl) {
    size_t width, rows, pitch_src, pitch_dst;
    AecmCallArgs k;
    int32_t rc;
    int adjacent;
    width = (size_t)h->nCalls * (size_t)h->nrOfSamples; /* samples actually used per row */
    rows = (size_t)h->count;
    rc = grow_planes(b, sl, rows * width, h->nearendClean != NULL);
    if (rc) return rc;
    pitch_src = (size_t)h->rowStride * sizeof(int16_t);
    pitch_dst = width * sizeof(int16_t);
    /* earlier work on the batch's main stream (Init, set_config, device-plane calls) comes first */
    if (b->main_dirty) {
        int i;
        CU(b, cudaEventRecord(b->ev_main, b->stream));
        for (i = 0; i < AECM_SLOTS; i++) CU(b, cudaStreamWaitEvent(b->slot[i].copy, b->ev_main, 0));
        b->main_dirty = 0;
    }
    if (b->trace) CU(b, cudaEventRecord(sl->t[0], sl->copy));
    memset(&k, 0, sizeof(k));
    sl->d_near = sl->d_far + rows * width; /* right behind the far plane of this call */
    /* dense far plane immediately followed by the dense near plane: one DMA for both */
    adjacent = h->farend && h->nearendNoisy && pitch_src == pitch_dst && h->nearendNoisy == h->farend + rows * width;
    if (adjacent) {
        /* Two planes that merely touch in the address space but come from two pinned allocations cannot be moved by one
         * copy: the runtime refuses it up front (invalid value, nothing enqueued) and the planes go separately. */
        const cudaError_t e = cudaMemcpyAsync(sl->d_far, h->farend, 2 * rows * pitch_dst, cudaMemcpyHostToDevice, sl->copy);
        if (e == cudaErrorInvalidValue) {
            (void)cudaGetLastError();
            adjacent = 0;
        } else if (e != cudaSuccess) {
            return fail(b, "cudaMemcpyAsync(far|near)", e);
        } else {
            k.far = sl->d_far;
        }
    }
    if (!adjacent && h->farend) {
        CU(b, copy_plane(sl->d_far, pitch_dst, h->farend, pitch_src, pitch_dst, rows, cudaMemcpyHostToDevice, sl->copy));
        k.far = sl->d_far;
    }
    if (h->nearendNoisy) {
        if (!adjacent)
            CU(b, copy_plane(sl->d_near, pitch_dst, h->nearendNoisy, pitch_src, pitch_dst, rows, cudaMemcpyHostToDevice, sl->copy));
        k.near_noisy = sl->d_near;
        k.out = sl->d_out;
        if (h->nearendClean) {
            CU(b, copy_plane(sl->d_clean, pitch_dst, h->nearendClean, pitch_src, pitch_dst, rows, cudaMemcpyHostToDevice, sl->copy));
            k.near_clean = sl->d_clean;
        }
    }
    if (h->msInSndCardBuf) {
        CU(b, cudaMemcpyAsync(sl->d_ms, h->msInSndCardBuf, rows * sizeof(int16_t), cudaMemcpyHostToDevice, sl->copy));
        k.ms = sl->d_ms;
    }
    if (h->instanceIndex) {
        CU(b, cudaMemcpyAsync(sl->d_index, h->instanceIndex, rows * sizeof(int32_t), cudaMemcpyHostToDevice, sl->copy));
        k.index = sl->d_index;
    }
    if (h->farendMask) {
        const size_t mb = rows * (size_t)h->nCalls;
        if (mb > sl->mask_bytes) {
            cudaFree(sl->d_mask);
            sl->d_mask = NULL;
            sl->mask_bytes = 0;
            CU(b, cudaMalloc((void **)&sl->d_mask, mb));
            sl->mask_bytes = mb;
        }
        CU(b, cudaMemcpyAsync(sl->d_mask, h->farendMask, mb, cudaMemcpyHostToDevice, sl->copy));
        k.far_mask = sl->d_mask;
    }
    k.mask_stride = h->count;
    k.status = sl->d_status;
    k.row_stride = (int64_t)width;
    k.first = h->first;
    k.count = h->count;
    k.samples_per_call = h->nrOfSamples;
    k.n_calls = h->nCalls;
    k.ms_uniform = h->msUniform;
    k.ext = b->ext_active;
    if (b->trace) CU(b, cudaEventRecord(sl->t[1], sl->copy));
    /* the state recurrence: this kernel must follow the previous submission's kernel (on the other slot's stream) */
    if (b->last_kernel) CU(b, cudaStreamWaitEvent(sl->copy, b->last_kernel, 0));
    if (b->trace) CU(b, cudaEventRecord(sl->t[2], sl->copy));
    CU(b, (cudaError_t)aecm_cuda_launch_call(&b->planes, &k, sl->copy));
    CU(b, cudaEventRecord(sl->kernel_done, sl->copy));
    if (b->trace) CU(b, cudaEventRecord(sl->t[3], sl->copy));
    b->last_kernel = sl->kernel_done;
    b->slots_dirty = 1; /* later main-stream work must see this call's state: see use_main_stream() */
    if (h->nearendNoisy)
        CU(b, copy_plane(h->out, pitch_src, sl->d_out, pitch_dst, pitch_dst, rows, cudaMemcpyDeviceToHost, sl->copy));
    if (h->status)
        CU(b, cudaMemcpyAsync(sl->h_status, sl->d_status, rows * sizeof(int32_t), cudaMemcpyDeviceToHost, sl->copy));
    if (b->trace) CU(b, cudaEventRecord(sl->t[4], sl->copy));
    return 0;
}

int32_t AecmB200_WaitStream(AecmB200Batch *b, int32_t slot) {
    Slot *sl;
    if (!b) return -1;
    if (slot < 0 || slot >= AECM_SLOTS) return AECM_BAD_PARAMETER_ERROR;
    sl = &b->slot[slot];
    if (!sl->busy) return 0;
    ENTER_DEVICE(b);
    CU(b, cudaStreamSynchronize(sl->copy));
    if (b->trace && sl->count > 0) {
        float h2d = 0, gap = 0, ker = 0, d2h = 0;
        cudaEventElapsedTime(&h2d, sl->t[0], sl->t[1]);
        cudaEventElapsedTime(&gap, sl->t[1], sl->t[2]);
        cudaEventElapsedTime(&ker, sl->t[2], sl->t[3]);
        cudaEventElapsedTime(&d2h, sl->t[3], sl->t[4]);
        fprintf(stderr, "aecm_b200 trace slot %d: h2d %.1f us, wait-for-previous-kernel %.1f us, kernel %.1f us, d2h %.1f us\n",
                slot, h2d * 1e3f, gap * 1e3f, ker * 1e3f, d2h * 1e3f);
    }
    if (sl->user_status && sl->count > 0) memcpy(sl->user_status, sl->h_status, (size_t)sl->count * sizeof(int32_t));
    sl->busy = 0;
    return 0;
}

int32_t AecmB200_ProcessStream(AecmB200Batch *b, const AecmB200StreamArgs *h) {
    int32_t rc;
    if (!b) return -1;
    rc = AecmB200_WaitStream(b, 0); /* keep submission order if the caller mixes the two styles */
    if (rc) return rc;
    rc = AecmB200_SubmitStream(b, h, 0); /* a failed submission leaves the slot free */
    if (rc) return rc;
    return AecmB200_WaitStream(b, 0);
}

int32_t AecmB200_BufferFarend(AecmB200Batch *b, const int16_t *farend, size_t n, int32_t *status) {
    AecmB200StreamArgs a;
    if (!b) return -1;
    if (!farend) return AECM_NULL_POINTER_ERROR; /* echo_control_mobile.cc:203 */
    memset(&a, 0, sizeof(a));
    a.farend = farend;
    a.rowStride = (int64_t)n;
    a.count = b->n;
    a.nrOfSamples = (int32_t)n;
    a.nCalls = 1;
    a.status = status;
    if (n != 80 && n != 160) return AECM_BAD_PARAMETER_ERROR;
    return AecmB200_ProcessStream(b, &a);
}

int32_t AecmB200_Process(AecmB200Batch *b, const int16_t *near_noisy, const int16_t *near_clean, int16_t *out, size_t n,
                         const int16_t *ms, int16_t ms_uniform, int32_t *status) {
    AecmB200StreamArgs a;
    if (!b) return -1;
    if (!near_noisy || !out) return AECM_NULL_POINTER_ERROR; /* echo_control_mobile.cc:256-262 */
    if (n != 80 && n != 160) return AECM_BAD_PARAMETER_ERROR;
    memset(&a, 0, sizeof(a));
    a.nearendNoisy = near_noisy;
    a.nearendClean = near_clean;
    a.out = out;
    a.rowStride = (int64_t)n;
    a.count = b->n;
    a.nrOfSamples = (int32_t)n;
    a.nCalls = 1;
    a.msInSndCardBuf = ms;
    a.msUniform = ms_uniform;
    a.status = status;
    return AecmB200_ProcessStream(b, &a);
}

/* ------------------------------------------------------------------------------------------------------------
 * Multi-GPU group: N instances split into contiguous shards, one batch per device, driven from one host thread.
 * Instances share nothing, so there is no exchange between shards: a group call submits the
 * shard-local part of the planes to every device (asynchronously) and then waits for all of them.
 * ------------------------------------------------------------------------------------------------------------ */
struct AecmB200Group {
    int32_t n, n_shards;
    AecmB200Batch **shard;
    int32_t *first; /* n_shards + 1 offsets */
};

AecmB200Group *AecmB200_GroupCreate(int32_t n, int32_t n_devices, const int32_t *devices) {
    AecmB200Group *g;
    int i;
    if (n <= 0 || n_devices <= 0 || n_devices > n) return NULL;
    g = (AecmB200Group *)calloc(1, sizeof(*g));
    if (!g) return NULL;
    g->n = n;
    g->n_shards = n_devices;
    g->shard = (AecmB200Batch **)calloc((size_t)n_devices, sizeof(*g->shard));
    g->first = (int32_t *)calloc((size_t)n_devices + 1, sizeof(int32_t));
    if (!g->shard || !g->first) {
        AecmB200_GroupFree(g);
        return NULL;
    }
    for (i = 0; i < n_devices; i++) { /* ceil-split: the first n % n_devices shards get one more instance */
        const int32_t base = n / n_devices, extra = n % n_devices;
        g->first[i + 1] = g->first[i] + base + (i < extra ? 1 : 0);
        g->shard[i] = AecmB200_Create(g->first[i + 1] - g->first[i], devices ? devices[i] : i);
        if (!g->shard[i]) {
            AecmB200_GroupFree(g);
            return NULL;
        }
    }
    return g;
}

void AecmB200_GroupFree(AecmB200Group *g) {
    int i;
    if (!g) return;
    for (i = 0; g->shard && i < g->n_shards; i++) AecmB200_Free(g->shard[i]);
    free(g->shard);
    free(g->first);
    free(g);
}

int32_t AecmB200_GroupNumShards(const AecmB200Group *g) { return g ? g->n_shards : -1; }

AecmB200Batch *AecmB200_GroupShard(const AecmB200Group *g, int32_t i, int32_t *first, int32_t *count) {
    if (!g || i < 0 || i >= g->n_shards) return NULL;
    if (first) *first = g->first[i];
    if (count) *count = g->first[i + 1] - g->first[i];
    return g->shard[i];
}

int32_t AecmB200_GroupInit(AecmB200Group *g, int32_t fs) {
    int i;
    int32_t rc;
    if (!g) return -1;
    for (i = 0; i < g->n_shards; i++) {
        rc = AecmB200_Init(g->shard[i], 0, g->first[i + 1] - g->first[i], fs);
        if (rc) return rc;
    }
    return 0;
}

int32_t AecmB200_GroupSetConfig(AecmB200Group *g, AecmConfig cfg, int32_t *status) {
    int i;
    int32_t rc;
    if (!g) return -1;
    for (i = 0; i < g->n_shards; i++) {
        rc = AecmB200_SetConfig(g->shard[i], 0, g->first[i + 1] - g->first[i], cfg, status ? status + g->first[i] : NULL);
        if (rc) return rc;
    }
    return 0;
}

/* the shard-local slice of the caller's range goes to every device on stream slot `slot` */
int32_t AecmB200_GroupSubmitStream(AecmB200Group *g, const AecmB200StreamArgs *h, int32_t slot) {
    int i;
    int32_t rc = 0, worst = 0;
    if (!g) return -1;
    if (!h) return AECM_NULL_POINTER_ERROR;
    if (h->first < 0 || h->count < 0 || (int64_t)h->first + h->count > g->n) return AECM_BAD_PARAMETER_ERROR;
    if (h->farendMask || h->instanceIndex) return AECM_UNSUPPORTED_FUNCTION_ERROR; /* both are laid out over the whole range */
    for (i = 0; i < g->n_shards; i++) {
        const int32_t lo = h->first > g->first[i] ? h->first : g->first[i];
        const int32_t hi = h->first + h->count < g->first[i + 1] ? h->first + h->count : g->first[i + 1];
        AecmB200StreamArgs a = *h;
        const int64_t row = (int64_t)lo - h->first; /* row of the caller's planes where this slice starts */
        if (hi <= lo) continue;
        a.first = lo - g->first[i];
        a.count = hi - lo;
        if (h->farend) a.farend = h->farend + row * h->rowStride;
        if (h->nearendNoisy) a.nearendNoisy = h->nearendNoisy + row * h->rowStride;
        if (h->nearendClean) a.nearendClean = h->nearendClean + row * h->rowStride;
        if (h->out) a.out = h->out + row * h->rowStride;
        if (h->msInSndCardBuf) a.msInSndCardBuf = h->msInSndCardBuf + row;
        if (h->status) a.status = h->status + row;
        rc = AecmB200_SubmitStream(g->shard[i], &a, slot);
        if (rc) worst = rc;
    }
    return worst;
}

int32_t AecmB200_GroupWaitStream(AecmB200Group *g, int32_t slot) {
    int i;
    int32_t rc, worst = 0;
    if (!g) return -1;
    for (i = 0; i < g->n_shards; i++) {
        rc = AecmB200_WaitStream(g->shard[i], slot);
        if (rc) worst = rc;
    }
    return worst;
}

/* submit to all devices, then collect */
int32_t AecmB200_GroupProcessStream(AecmB200Group *g, const AecmB200StreamArgs *h) {
    int32_t rc, rc2;
    if (!g) return -1;
    rc = AecmB200_GroupWaitStream(g, 0); /* keep submission order if the caller mixes the two styles */
    if (rc) return rc;
    rc = AecmB200_GroupSubmitStream(g, h, 0);
    rc2 = AecmB200_GroupWaitStream(g, 0);
    return rc ? rc : rc2;
}
