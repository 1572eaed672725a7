/*
 * cuda_runtime_api.h -- TEST INFRASTRUCTURE, found before the real header only by tests/test_host_orchestration_model.py.
 * The subset of the CUDA runtime API that webrtc_aecm_b200/csrc/aecm_batch.c uses, declared for the host-side model in
 * tests/host_emu/host_model.c (streams as lazily executed in-order queues, events, poisoned "device" memory).
 */
#ifndef FAKE_CUDA_RUNTIME_API_H_
#define FAKE_CUDA_RUNTIME_API_H_
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum cudaError {
    cudaSuccess = 0,
    cudaErrorInvalidValue = 1,
    cudaErrorMemoryAllocation = 2,
    cudaErrorInvalidDevice = 101,
    cudaErrorInvalidResourceHandle = 400,
    cudaErrorUnknown = 999
};
typedef enum cudaError cudaError_t;

enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };

typedef struct FakeStream *cudaStream_t;
typedef struct FakeEvent *cudaEvent_t;

#define cudaStreamNonBlocking 0x01
#define cudaEventDisableTiming 0x02

const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError(void);
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaGetDevice(int *d);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaMalloc(void **p, size_t bytes);
cudaError_t cudaFree(void *p);
cudaError_t cudaMallocHost(void **p, size_t bytes);
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags);
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, enum cudaMemcpyKind kind, cudaStream_t s);
cudaError_t cudaMemcpy2DAsync(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height,
                              enum cudaMemcpyKind kind, cudaStream_t s);
cudaError_t cudaMemsetAsync(void *dst, int value, size_t bytes, cudaStream_t s);

#ifdef __cplusplus
}
#endif
#endif
