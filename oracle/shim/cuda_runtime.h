// Host-CPU stand-in for the few pieces of <cuda_runtime.h> that the reference translation unit
// (/root/reference/mpm/csrc/integrator.cu and its headers) touches.  TEST INFRASTRUCTURE ONLY: it lets
// oracle/build_ref.sh compile the UNMODIFIED reference sources with g++ into oracle/_ref/libmaniskill_mpm_cpu.so
// so that the reference's own kernels can be executed on host cores (one "thread" per loop iteration).
// Nothing in dexdeform_b200/ includes this file.
#pragma once
#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <stddef.h>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))

struct shim_dim3 { unsigned x = 1, y = 1, z = 1; };
extern thread_local shim_dim3 blockIdx, threadIdx;
extern shim_dim3 blockDim, gridDim;

struct float3 { float x, y, z; };

typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef struct shim_stream_st *cudaStream_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };

inline const char *cudaGetErrorString(cudaError_t) { return "shim"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return cudaSuccess; }
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2D(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind) {
  for (size_t i = 0; i < h; ++i) memcpy((char *)d + i * dp, (const char *)s + i * sp, w);
  return cudaSuccess;
}
inline cudaError_t cudaMemset(void *p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 0; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) { *d = -1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = *t = 0; return cudaSuccess; }

// texture / array API used only by the renderer's create_volume/destroy_volume (out of scope): inert stubs
typedef void *cudaArray_t;
typedef unsigned long long cudaTextureObject_t;
enum cudaChannelFormatKind { cudaChannelFormatKindFloat };
struct cudaChannelFormatDesc { int x, y, z, w; cudaChannelFormatKind f; };
inline cudaChannelFormatDesc cudaCreateChannelDesc(int x, int y, int z, int w, cudaChannelFormatKind f) { return {x, y, z, w, f}; }
struct cudaExtent { size_t width, height, depth; };
inline cudaExtent make_cudaExtent(size_t w, size_t h, size_t d) { return {w, h, d}; }
struct cudaPitchedPtr { void *ptr; size_t pitch, xsize, ysize; };
inline cudaPitchedPtr make_cudaPitchedPtr(void *p, size_t pitch, size_t xs, size_t ys) { return {p, pitch, xs, ys}; }
struct cudaMemcpy3DParms { cudaPitchedPtr srcPtr; cudaArray_t dstArray; cudaExtent extent; cudaMemcpyKind kind; };
inline cudaError_t cudaMalloc3DArray(cudaArray_t *a, const cudaChannelFormatDesc *, cudaExtent) { *a = nullptr; return cudaSuccess; }
inline cudaError_t cudaMemcpy3D(const cudaMemcpy3DParms *) { return cudaSuccess; }
enum cudaResourceType { cudaResourceTypeArray };
struct cudaResourceDesc { cudaResourceType resType; struct { struct { cudaArray_t array; } array; } res; };
enum cudaTextureAddressMode { cudaAddressModeClamp };
enum cudaTextureFilterMode { cudaFilterModeLinear };
enum cudaTextureReadMode { cudaReadModeElementType };
struct cudaTextureDesc { cudaTextureAddressMode addressMode[3]; cudaTextureFilterMode filterMode; cudaTextureReadMode readMode; int normalizedCoords; };
inline cudaError_t cudaCreateTextureObject(cudaTextureObject_t *t, const cudaResourceDesc *, const cudaTextureDesc *, const void *) { *t = 0; return cudaSuccess; }
inline cudaError_t cudaDestroyTextureObject(cudaTextureObject_t) { return cudaSuccess; }
inline cudaError_t cudaFreeArray(cudaArray_t) { return cudaSuccess; }

// CUDA's global-namespace min/max overload set (math_functions.hpp): mixed float/double promotes to double
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline float min(float a, float b) { return fminf(a, b); }
inline float max(float a, float b) { return fmaxf(a, b); }
inline double min(double a, double b) { return fmin(a, b); }
inline double max(double a, double b) { return fmax(a, b); }
inline double min(float a, double b) { return fmin((double)a, b); }
inline double min(double a, float b) { return fmin(a, (double)b); }
inline double max(float a, double b) { return fmax((double)a, b); }
inline double max(double a, float b) { return fmax(a, (double)b); }

// atomics: real atomics so the emulated grid may be run by several host threads
inline float atomicAdd(float *a, float b) {
  float old;
#pragma omp atomic capture
  { old = *a; *a += b; }
  return old;
}
inline int atomicMin(int *a, int b) {
  int old = __atomic_load_n(a, __ATOMIC_RELAXED);
  while (b < old && !__atomic_compare_exchange_n(a, &old, b, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
inline int atomicMax(int *a, int b) {
  int old = __atomic_load_n(a, __ATOMIC_RELAXED);
  while (b > old && !__atomic_compare_exchange_n(a, &old, b, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
inline long long atomicMin(long long *a, long long b) {
  long long old = __atomic_load_n(a, __ATOMIC_RELAXED);
  while (b < old && !__atomic_compare_exchange_n(a, &old, b, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
