/* tests/emu/cuda_runtime.h -- a host stand-in for the few CUDA runtime calls and kernel-language keywords that
 * mpas-seaice_b200/csrc/ir_kernels.cu uses, so that tests can compile that file with g++ and step through its kernels
 * one thread at a time (sequentially, block by block) where no GPU is available.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing here is shipped or loaded by the product: the library built from it lives under
 * tests/_build/ and is opened by tests/test_ir_parity.py alone.  It checks the kernels' LOGIC against the oracle
 * (indexing, operation order, the tracer-row tables); it says nothing about races, device memory bounds or speed --
 * the `cuda` leg of the same tests does that on a B200.
 */
#ifndef TESTS_EMU_CUDA_RUNTIME_H
#define TESTS_EMU_CUDA_RUNTIME_H

#include <cmath>
#include <cstdlib>
#include <cstring>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
/* shared memory: the kernels use it as per-thread scratch columns only (no thread reads another's slot and there
 * is no barrier), so one static array reused by the sequentially executed threads behaves the same */
#define __shared__ static

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static thread_local dim3 blockIdx, threadIdx, blockDim, gridDim;

template <typename F>
static void emu_launch(dim3 grid, dim3 block, F body)
{
    gridDim = grid;
    blockDim = block;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++)
                for (unsigned tz = 0; tz < block.z; tz++)
                    for (unsigned ty = 0; ty < block.y; ty++)
                        for (unsigned tx = 0; tx < block.x; tx++) {
                            blockIdx = dim3(bx, by, bz);
                            threadIdx = dim3(tx, ty, tz);
                            body();
                        }
}
#define IR_LAUNCH(kernel, grid, block, stream, ...) emu_launch(dim3(grid), dim3(block), [&] { kernel(__VA_ARGS__); })

typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef int cudaStream_t;
typedef int cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1 };

static inline const char *cudaGetErrorString(cudaError_t) { return "emulation"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? cudaSuccess : 2; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = 1; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = 1; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
enum { cudaHostRegisterDefault = 0 };
static inline cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
static inline int atomicOr(int *p, int v) { const int old = *p; *p = old | v; return old; }

#endif
