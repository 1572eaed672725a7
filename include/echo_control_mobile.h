/*
 * echo_control_mobile.h -- the reference's public C API, re-exported by libaecm_b200.so.
 *
 * Drop-in replacement for cpuimage/WebRTC_AECM's aecm/echo_control_mobile.h: same ten symbols, same
 * signatures, same error conventions, so the reference's only caller (main.cc:97-147) links against
 * libaecm_b200.so unmodified.  Each handle is a slot of a device-resident batch; every call is executed by
 * the CUDA kernel (there is no CPU implementation behind these symbols).  For throughput use the batched
 * ABI in aecm_b200.h instead: these scalar calls pay one host<->device round trip per 10 ms frame.
 *
 * Interface replaced, entry by entry (reference file aecm/echo_control_mobile.h):
 *   WebRtcAecm_Create :46            WebRtcAecm_Free :55            WebRtcAecm_Init :70
 *   WebRtcAecm_BufferFarend :87      WebRtcAecm_GetBufferFarendError :106
 *   WebRtcAecm_Process :135-140      WebRtcAecm_set_config :156
 *   WebRtcAecm_InitEchoPath :172     WebRtcAecm_GetEchoPath :191   WebRtcAecm_echo_path_size_bytes :202
 */
#ifndef AECM_B200_ECHO_CONTROL_MOBILE_H_
#define AECM_B200_ECHO_CONTROL_MOBILE_H_

#include <stddef.h>
#include <stdint.h>

enum { AecmFalse = 0, AecmTrue }; /* reference :18-20 */

/* error / warning codes, reference :22-30 */
#define AECM_UNSPECIFIED_ERROR 12000
#define AECM_UNSUPPORTED_FUNCTION_ERROR 12001
#define AECM_UNINITIALIZED_ERROR 12002
#define AECM_NULL_POINTER_ERROR 12003
#define AECM_BAD_PARAMETER_ERROR 12004
#define AECM_BAD_PARAMETER_WARNING 12100

typedef struct {
    int16_t cngMode;  /* AecmFalse, AecmTrue (default) */
    int16_t echoMode; /* 0, 1, 2, 3 (default), 4 */
} AecmConfig; /* reference :32-35 */

#ifdef __cplusplus
extern "C" {
#endif

/* Allocates one AECM instance (a slot of the process-wide device batch). NULL on failure. */
void *WebRtcAecm_Create(void);
/* Releases the instance; NULL is a no-op. */
void WebRtcAecm_Free(void *aecmInst);
/* (Re)initialises the instance for sampFreq 8000 or 16000 and applies the default config
 * {cngMode = AecmTrue, echoMode = 3}. 0, -1 (NULL instance) or AECM_BAD_PARAMETER_ERROR. */
int32_t WebRtcAecm_Init(void *aecmInst, int32_t sampFreq);
/* Buffers 80 or 160 far-end samples. */
int32_t WebRtcAecm_BufferFarend(void *aecmInst, const int16_t *farend, size_t nrOfSamples);
/* The error WebRtcAecm_BufferFarend would return for these arguments, without buffering. */
int32_t WebRtcAecm_GetBufferFarendError(void *aecmInst, const int16_t *farend, size_t nrOfSamples);
/* Processes 80 or 160 near-end samples; nearendClean may be NULL (then for every call of the stream, as with the
 * reference: its clean-frame FIFO only keeps step with the noisy one that way); out may alias the input.
 * Returns 0, an error code, or AECM_BAD_PARAMETER_WARNING when msInSndCardBuf was clamped to [0, 500]. */
int32_t WebRtcAecm_Process(void *aecmInst, const int16_t *nearendNoisy, const int16_t *nearendClean, int16_t *out,
                           size_t nrOfSamples, int16_t msInSndCardBuf);
int32_t WebRtcAecm_set_config(void *aecmInst, AecmConfig config);
/* Imports / exports the stored echo channel: 65 x int16 = WebRtcAecm_echo_path_size_bytes() bytes. */
int32_t WebRtcAecm_InitEchoPath(void *aecmInst, const void *echo_path, size_t size_bytes);
int32_t WebRtcAecm_GetEchoPath(void *aecmInst, void *echo_path, size_t size_bytes);
size_t WebRtcAecm_echo_path_size_bytes(void);

#ifdef __cplusplus
}
#endif
#endif /* AECM_B200_ECHO_CONTROL_MOBILE_H_ */
