/*
 * aecm_b200.h -- batched C ABI of the B200-native AECM engine (libaecm_b200.so).
 *
 * A batch is N independent AECM instances resident on one GPU (one warp per instance, state in HBM).
 * Every entry point below applies the *same* reference API call to a range of instances in lockstep, in a
 * single kernel launch.  The per-instance semantics -- arguments, state machine, int16 output and return
 * code -- are exactly those of the reference call named beside each function (reference file
 * aecm/echo_control_mobile.h / .cc in cpuimage/WebRTC_AECM).  Plain C: pointers, sizes and integers only.
 *
 * Return value of every int32_t function: 0 when the call was issued; -1 for a NULL batch;
 * AECM_NULL_POINTER_ERROR / AECM_BAD_PARAMETER_ERROR for bad arguments that the reference would reject for
 * every instance alike; AECM_UNSPECIFIED_ERROR when CUDA failed (AecmB200_LastError() has the text).
 * Per-instance return codes (0, AECM_UNINITIALIZED_ERROR, AECM_BAD_PARAMETER_WARNING, ...) are delivered
 * in the optional status[] array, one int32 per instance of the range.
 *
 * Threads: a batch (and a group) takes one caller at a time -- there is no locking inside; different batches may be
 * driven from different threads, also on the same GPU.  The ten WebRtcAecm_* symbols of echo_control_mobile.h ARE
 * safe to call concurrently for different instances, as with the reference (they serialise and combine their calls
 * internally).  There is no CPU implementation behind any of this: without a CUDA device Create returns NULL.
 */
#ifndef AECM_B200_H_
#define AECM_B200_H_

#include <stddef.h>
#include <stdint.h>

#include "echo_control_mobile.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct AecmB200Batch AecmB200Batch;

/* N x { WebRtcAecm_Create (echo_control_mobile.h:46) } on CUDA device `device`.  NULL on failure. */
AecmB200Batch *AecmB200_Create(int32_t nInstances, int32_t device);
/* N x { WebRtcAecm_Free (:55) }.  NULL is a no-op. */
void AecmB200_Free(AecmB200Batch *batch);
int32_t AecmB200_NumInstances(const AecmB200Batch *batch);
int32_t AecmB200_Device(const AecmB200Batch *batch);
const char *AecmB200_LastError(const AecmB200Batch *batch);

/* Returns instances [first, first+count) to the freshly created (zero-filled, un-initialised) state. */
int32_t AecmB200_Reset(AecmB200Batch *batch, int32_t first, int32_t count);
/* count x { WebRtcAecm_Init (:70) } on instances [first, first+count). */
int32_t AecmB200_Init(AecmB200Batch *batch, int32_t first, int32_t count, int32_t sampFreq);
/* count x { WebRtcAecm_set_config (:156) }. status: host array [count] or NULL. */
int32_t AecmB200_SetConfig(AecmB200Batch *batch, int32_t first, int32_t count, AecmConfig config, int32_t *status);
/* WebRtcAecm_InitEchoPath (:172) / WebRtcAecm_GetEchoPath (:191) for one instance (host buffers). */
int32_t AecmB200_InitEchoPath(AecmB200Batch *batch, int32_t instance, const void *echo_path, size_t size_bytes);
int32_t AecmB200_GetEchoPath(AecmB200Batch *batch, int32_t instance, void *echo_path, size_t size_bytes);

/*
 * The streaming call.  For every instance r of [first, first+count) and every call c of [0, nCalls):
 *      if (farend && (!farendMask || farendMask[c*count + r]))
 *          WebRtcAecm_BufferFarend(inst, farend + r*rowStride + c*nrOfSamples, nrOfSamples);        (:87)
 *      if (nearendNoisy)
 *          WebRtcAecm_Process(inst, nearendNoisy + r*rowStride + c*nrOfSamples,
 *                             nearendClean ? nearendClean + ... : NULL, out + ..., nrOfSamples,
 *                             msInSndCardBuf ? msInSndCardBuf[r] : msUniform);                        (:135-140)
 * status[r] receives the first non-zero return code among those calls (0 if none).  Where WebRtcAecm_Process refuses a call
 * (an error code such as AECM_UNINITIALIZED_ERROR; not the AECM_BAD_PARAMETER_WARNING of a clamped msInSndCardBuf, which is
 * returned after processing) the reference leaves `out` untouched; here the row's `out` samples of that call are unspecified.
 * nCalls > 1 fuses consecutive 10 ms calls into one launch with the instance state held on-chip.
 */
typedef struct {
    const int16_t *farend;        /* [count][rowStride] plane or NULL */
    const int16_t *nearendNoisy;  /* [count][rowStride] plane or NULL */
    const int16_t *nearendClean;  /* optional plane.  Per instance: on every Process call of a stream or on none -- switching it
                                     on and off in mid-stream leaves the reference reading stale / uninitialised memory
                                     (aecm_core.cc:537-546 with ring_buffer.c:113-131), so no behaviour is defined for it */
    int16_t *out;                 /* [count][rowStride] plane (required with nearendNoisy) */
    int64_t rowStride;            /* samples between instance rows; >= nCalls * nrOfSamples */
    int32_t first, count;         /* instance range */
    int32_t nrOfSamples;          /* 80 or 160, as in the reference */
    int32_t nCalls;               /* >= 1 */
    const int16_t *msInSndCardBuf;/* [count] or NULL */
    int16_t msUniform;            /* used when msInSndCardBuf is NULL */
    const uint8_t *farendMask;    /* optional [nCalls][count] */
    int32_t *status;              /* optional [count] */
    const int32_t *instanceIndex; /* optional [count]: plane row r is instance instanceIndex[r] instead of first + r;
                                     the indices must be distinct (one warp per row, no locking); `first` is ignored.
                                     The host-plane calls reject an out-of-range or repeated index with
                                     AECM_BAD_PARAMETER_ERROR before anything is enqueued; for the device array of
                                     AecmB200_ProcessStreamDevice distinctness is the caller's responsibility (an
                                     out-of-range row only gets AECM_BAD_PARAMETER_ERROR in its status) */
} AecmB200StreamArgs;

/* Host-memory planes: copies in, runs, copies out and waits (pinned planes avoid the staging copy). */
int32_t AecmB200_ProcessStream(AecmB200Batch *batch, const AecmB200StreamArgs *hostArgs);
/* The same call split in two for pipelining: Submit enqueues the H2D copies, the kernel and the D2H copies on one of
 * AecmB200_NumStreamSlots() slots (0 .. n-1) and returns at once; Wait blocks until that slot's results (out planes,
 * status) are in host memory, after which the slot may be submitted to again.  Kernels execute in submission order;
 * the copies of one slot overlap the kernels and copies of the others.  A call occupies its slot for H2D + kernel +
 * D2H, so keeping the PCIe link busy takes three or four calls in flight (round-robin over the slots, waiting for
 * the oldest).  The host planes (pinned for real overlap) and the status array must stay valid until Wait.  Tip: if
 * the dense near plane starts exactly where the dense far plane ends, both cross PCIe in one transfer.
 * A Submit that fails (bad arguments, CUDA error) leaves its slot free: whatever it had enqueued is drained first and
 * nothing is delivered for it; Wait on a free slot returns 0 and touches no status array. */
int32_t AecmB200_SubmitStream(AecmB200Batch *batch, const AecmB200StreamArgs *hostArgs, int32_t slot);
int32_t AecmB200_WaitStream(AecmB200Batch *batch, int32_t slot);
int32_t AecmB200_NumStreamSlots(void);
/* Device-memory planes on the batch's GPU: enqueues on the batch stream and returns immediately. */
int32_t AecmB200_ProcessStreamDevice(AecmB200Batch *batch, const AecmB200StreamArgs *deviceArgs);

/* Lockstep mirrors of the two reference calls over the whole batch; planes are [N][nrOfSamples], host memory. */
int32_t AecmB200_BufferFarend(AecmB200Batch *batch, const int16_t *farend, size_t nrOfSamples, int32_t *status);
int32_t AecmB200_Process(AecmB200Batch *batch, const int16_t *nearendNoisy, const int16_t *nearendClean, int16_t *out,
                         size_t nrOfSamples, const int16_t *msInSndCardBuf, int16_t msUniform, int32_t *status);

/*
 * SURVEY 8-f4 -- the delay estimator's optional machinery.  The reference's estimator object carries a robust-validation
 * mode (a float histogram of candidate delays, delay_estimator.cc:96-266), a near-end lookahead (:536-543) and a soft
 * reset (:336-367, 506-519); AECM wires it with robust validation off and max_lookahead 0 (aecm_core.cc:218, 225) and
 * has no public call that reaches the estimator, so these three entry points have no counterpart among the ten
 * WebRtcAecm_* symbols: they do to the estimator of every instance in the range what the cited setters of
 * delay_estimator_wrapper.h do to one estimator handle.  Bit-exact against the reference driven through the same
 * setters (float arithmetic included).  The first call of either of the first two switches the batch to the kernel
 * instantiation that keeps the robust-validation statistics, which the reference keeps from Init on
 * (delay_estimator.cc:628-633): to follow the reference exactly, set the options right after AecmB200_Init.  Options
 * survive AecmB200_Init, as the properties of the reference's estimator object do; AecmB200_Reset clears them.
 */
typedef struct {
    int32_t robustValidation; /* WebRtc_enable_robust_validation (delay_estimator_wrapper.cc:421-434): 0 or 1 */
    int32_t allowedOffset;    /* WebRtc_set_allowed_offset (:401-409): >= 0 */
    int32_t maxLookahead;     /* >= 0: the estimator is re-created with this max_lookahead (WebRtc_CreateDelayEstimator, :289-336; at
                                 most 15; lookahead starts at maxLookahead, robust validation off, offset 0) and initialised
                                 (WebRtc_InitDelayEstimator, :338-355) before the other fields are applied; -1: leave it */
    int32_t lookahead;        /* WebRtc_set_lookahead (:384-393): 0 .. max_lookahead; -1: leave it */
} AecmB200DelayEstimatorOptions;
/* status: optional host array [count]: 0, AECM_UNINITIALIZED_ERROR, or -1 where WebRtc_set_lookahead refuses the value */
int32_t AecmB200_SetDelayEstimatorOptions(AecmB200Batch *batch, int32_t first, int32_t count,
                                          const AecmB200DelayEstimatorOptions *options, int32_t *status);
/* WebRtc_SoftResetDelayEstimatorFarend + WebRtc_SoftResetDelayEstimator (:227-231, 357-361) with delayShift in (-100, 100);
 * applied: optional host array [count] receiving what the latter returns (the lookahead shift actually applied). */
int32_t AecmB200_SoftResetDelayEstimator(AecmB200Batch *batch, int32_t first, int32_t count, int32_t delayShift, int32_t *applied);
/* WebRtc_last_delay (:498-506; -2 until the first valid estimate) and WebRtc_last_delay_quality (:508-512) per instance;
 * either host array may be NULL. */
int32_t AecmB200_GetDelayEstimate(AecmB200Batch *batch, int32_t first, int32_t count, int32_t *lastDelay, float *quality);

/* Multi-GPU group: nInstances split into contiguous shards (ceil-split), one batch per entry of devices[] (NULL =
 * devices 0..nDevices-1), driven from the calling thread.  Instances are independent, so shards never exchange
 * data: a group call submits every shard's slice of the host planes asynchronously, then waits for all of them.
 * first/count in the args address the group's global instance numbering; farendMask and instanceIndex are not supported
 * here (both are laid out over the whole range).  Everything else -- InitEchoPath, the delay-estimator options, device
 * planes -- goes through the shard's own batch handle (AecmB200_GroupShard). */
typedef struct AecmB200Group AecmB200Group;
AecmB200Group *AecmB200_GroupCreate(int32_t nInstances, int32_t nDevices, const int32_t *devices);
void AecmB200_GroupFree(AecmB200Group *group);
int32_t AecmB200_GroupNumShards(const AecmB200Group *group);
AecmB200Batch *AecmB200_GroupShard(const AecmB200Group *group, int32_t shard, int32_t *first, int32_t *count);
int32_t AecmB200_GroupInit(AecmB200Group *group, int32_t sampFreq);
int32_t AecmB200_GroupSetConfig(AecmB200Group *group, AecmConfig config, int32_t *status);
int32_t AecmB200_GroupProcessStream(AecmB200Group *group, const AecmB200StreamArgs *hostArgs);
/* the pipelined form, as AecmB200_SubmitStream / AecmB200_WaitStream: every shard works on its slice in slot `slot` */
int32_t AecmB200_GroupSubmitStream(AecmB200Group *group, const AecmB200StreamArgs *hostArgs, int32_t slot);
int32_t AecmB200_GroupWaitStream(AecmB200Group *group, int32_t slot);

/* stream plumbing: the batch owns a CUDA stream unless one is adopted (e.g. the caller's current stream) */
int32_t AecmB200_Synchronize(AecmB200Batch *batch);
void *AecmB200_CudaStream(const AecmB200Batch *batch);
int32_t AecmB200_AdoptCudaStream(AecmB200Batch *batch, void *cudaStream);

/* pinned host memory for the host-plane calls */
void *AecmB200_HostAlloc(size_t bytes);
void AecmB200_HostFree(void *p);

/* introspection (tests, tooling) */
size_t AecmB200_StateBytes(void);
int32_t AecmB200_ReadState(AecmB200Batch *batch, int32_t instance, void *dst, size_t bytes);
int64_t AecmB200_KernelLaunchCount(void);
/* Diagnostics of the scalar API's call combiner (echo_control_mobile.c): rounds run and BufferFarend/Process requests
 * served so far by the pool the instance `aecmInst` (a WebRtcAecm_Create handle) lives in. */
void AecmB200_ScalarApiStats(void *aecmInst, int64_t *rounds, int64_t *requests);
/* Diagnostics: runs the device's floor-sqrt (the magnitude step, signal_processing_library.cc:76-105) over the
 * arguments [first, first + count) on the current device and returns how many differ from the reference's
 * bit-by-bit construction (0 expected; negative = CUDA error). */
int64_t AecmB200_SelfTestSqrtFloor(uint32_t first, uint32_t count);
/* Diagnostics: runs the kernel's own warp FFT (the code and tables of the hot path) on n host vectors on the current
 * device, for known-answer tests against the reference's WebRtcSpl_RealForwardFFT / WebRtcSpl_RealInverseFFT
 * (real_fft.c:47-102 over complex_fft.c:241-491, order 7).
 *   Forward: in [n][128] real samples (already windowed) -> out [n][130] = bins 0..64 as interleaved re, im, exactly
 *            what WebRtcSpl_RealForwardFFT returns; twoSignals != 0 runs the far+near interleaved variant on vector pairs.
 *   Inverse: in [n][130] = bins 0..64 interleaved, as handed to WebRtcSpl_RealInverseFFT -> out [n][128] real samples
 *            and scale[n] = the value WebRtcSpl_RealInverseFFT returns (block-floating shifts, complex_fft.c:381-396).
 * Return: 0, AECM_NULL_POINTER_ERROR, AECM_BAD_PARAMETER_ERROR, or AECM_UNSPECIFIED_ERROR when CUDA failed. */
int32_t AecmB200_SelfTestFftForward(const int16_t *in, int16_t *out, int32_t n, int32_t twoSignals);
int32_t AecmB200_SelfTestFftInverse(const int16_t *in, int16_t *out, int32_t *scale, int32_t n);
int32_t AecmB200_DeviceCount(void);

#ifdef __cplusplus
}
#endif
#endif /* AECM_B200_H_ */
