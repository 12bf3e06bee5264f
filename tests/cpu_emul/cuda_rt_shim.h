// cuda_rt_shim.h -- TEST ONLY.  A synchronous stand-in for the few CUDA runtime calls pbk_api.cu and the launch wrappers
// of pbk_kernels.cu make, so that the WHOLE C ABI (context, batching, table growth, sharding, key exchange, lookup) can be
// compiled for the host and driven by the tests in the GPU-less build container: "device" memory is host memory, every
// asynchronous call completes before it returns (a valid execution of a stream program whose dependencies all point
// backwards in issue order), kernels run as one sequential thread (cuda_shim.h).  Never shipped, never timed, not a
// fallback: the product library has no code path that reaches this.
#pragma once
#include <chrono>
#include <cstdlib>
#include <cstring>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorPeerAccessAlreadyEnabled = 704 };
struct emul_stream;
typedef emul_stream *cudaStream_t;
struct emul_event { double ms; };
typedef emul_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaLimitMaxL2FetchGranularity = 5 };
struct cudaDeviceProp { int multiProcessorCount; };

static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : e == cudaErrorMemoryAllocation ? "out of memory" : "invalid value"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { p->multiProcessorCount = 4; return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = *t = (size_t)48 << 30; return cudaSuccess; }
static inline cudaError_t cudaDeviceSetLimit(int, size_t) { return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)malloc(n ? n : 8); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <typename T> static inline cudaError_t cudaMallocHost(T **p, size_t n) { *p = (T *)malloc(n ? n : 8); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)(size_t)1; return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = (cudaStream_t)(size_t)1; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline double emul_now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emul_event{0}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new emul_event{0}; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->ms = emul_now_ms(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->ms - a->ms); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
// peer mapping: "device memory" is host memory of this process, so contexts of one process see each other's buffers as they are;
// there is no second process to hand an IPC handle to
struct cudaIpcMemHandle_t { char reserved[64]; };
static const unsigned cudaIpcMemLazyEnablePeerAccess = 1;
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { memset(h, 0, sizeof *h); memcpy(h, &p, sizeof p); return cudaSuccess; }
static inline cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, &h, sizeof *p); return cudaSuccess; }
static inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
static inline cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) { *can = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaMemcpyPeerAsync(void *d, int, const void *s, int, size_t n, cudaStream_t) { if (n) memmove(d, s, n); return cudaSuccess; }
