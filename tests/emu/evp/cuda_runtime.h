/* tests/emu/evp/cuda_runtime.h -- host stand-in for the CUDA runtime and kernel language as used by the EVP library
 * (mpas-seaice_b200/csrc/evp_*.cu), on top of tests/emu/cuda_runtime.h (fibers, dim3, the basic runtime calls).
 *
 * TEST INFRASTRUCTURE ONLY.  tests/emu/evp_emu.py rewrites the triple-chevron launches of the shipped sources into
 * emu_submit(...) calls and the handful of inline-PTX helpers into plain C++, compiles the result with g++ into
 * tests/_build/libevp_b200_emu.so, and tests/test_evp_emulation.py drives that library through the same C ABI and the
 * same host code as the product.  What it checks: the kernels' logic (indexing, operation order, masks, the launch
 * sequence, graph capture and replay) against the oracle where no GPU exists, dependence on the thread order (a race
 * on the device) through IR_EMU_ORDER=reverse (the switch of tests/emu/cuda_runtime.h), and out-of-bounds accesses under AddressSanitizer.  What it cannot:
 * timing, memory-model questions, more than one rank.
 */
#ifndef TESTS_EMU_EVP_CUDA_RUNTIME_H
#define TESTS_EMU_EVP_CUDA_RUNTIME_H

#include "../cuda_runtime.h"

#include <stdint.h>

#include <memory>
#include <tuple>
#include <utility>

#define __align__(n) alignas(n)
#define __constant__ static
#define __ldcg(p) (*(p))

/* aligned like the device's vector types: a misaligned double2 / int2 access faults on a GPU and is reported by
 * -fsanitize=undefined in the AddressSanitizer build (evp_emu.library(asan=True)) */
struct alignas(16) double2 { double x, y; };
struct alignas(8) int2 { int x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
static inline int2 make_int2(int x, int y) { int2 r; r.x = x; r.y = y; return r; }

/* dynamic shared memory of the block being executed */
alignas(128) static unsigned char emu_evp_dyn_smem_one[256 * 1024];
static unsigned char *emu_evp_dyn_smem = emu_evp_dyn_smem_one;      /* (a cooperative launch points it at the running block's) */

/* ---- warp collectives: every thread of the block calls them (true for the kernels under test); implemented as a block
 * barrier over per-thread predicate slots, two generations alternating like __syncthreads_or ---- */
static int emu_pred[2][1024];
static inline unsigned emu_ballot_sync(unsigned, int pred)
{
    EmuFiber &f = emu_fibers[emu_current];
    const long gen = f.orCount++;
    const int slot = (int)(gen & 1);
    const unsigned tid = (unsigned)emu_current;
    emu_pred[slot][tid] = pred ? 1 : 0;
    emu_syncthreads();
    const unsigned nt = blockDim.x * blockDim.y * blockDim.z;
    const unsigned w0 = tid & ~31u;
    unsigned b = 0;
    for (unsigned l = 0; l < 32 && w0 + l < nt; l++)
        if (emu_pred[slot][w0 + l]) b |= 1u << l;
    return b;
}
#define __ballot_sync(mask, pred) emu_ballot_sync((mask), (pred))
#define __any_sync(mask, pred) (emu_ballot_sync((mask), (pred)) != 0u)
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline void __threadfence() {}
static inline void __threadfence_system() {}
static inline void __nanosleep(unsigned) {}
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline unsigned atomicAdd(unsigned *p, unsigned v) { const unsigned old = *p; *p = old + v; return old; }
static inline int atomicAdd(int *p, int v) { const int old = *p; *p = old + v; return old; }
static inline int atomicMax(int *p, int v) { const int old = *p; if (v > old) *p = v; return old; }
static inline int atomicExch(int *p, int v) { const int old = *p; *p = v; return old; }

/* ---- launches, streams, graphs.  One in-order "stream"; a launch runs at once unless the stream is being captured,
 * then it is recorded (arguments copied at that moment, like a kernel node) and runs at every cudaGraphLaunch. ---- */
struct EmuGraph { std::vector<std::function<void()>> nodes; };
typedef EmuGraph *cudaGraph_t;
typedef EmuGraph *cudaGraphExec_t;
inline EmuGraph *emu_capture = nullptr;      /* shared by the translation units of the library */
inline long emu_launches = 0;
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal, cudaStreamCaptureModeThreadLocal, cudaStreamCaptureModeRelaxed };

static inline void emu_enqueue(std::function<void()> fn)
{
    if (emu_capture) emu_capture->nodes.push_back(std::move(fn));
    else fn();
}
static inline cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode)
{
    if (emu_capture) return 900;
    emu_capture = new EmuGraph();
    return cudaSuccess;
}
static inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t *g)
{
    *g = emu_capture;
    emu_capture = nullptr;
    return *g ? cudaSuccess : 901;
}
static inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t *e, cudaGraph_t g, unsigned long long)
{
    *e = new EmuGraph(*g);
    return cudaSuccess;
}
static inline cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
static inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t g) { delete g; return cudaSuccess; }
static inline cudaError_t cudaGraphLaunch(cudaGraphExec_t g, cudaStream_t)
{
    for (auto &n : g->nodes) n();
    return cudaSuccess;
}

struct EmuCfg { dim3 grid, block; };
static inline EmuCfg emu_cfg(dim3 g, dim3 b, size_t = 0, cudaStream_t = 0) { EmuCfg c; c.grid = g; c.block = b; return c; }

template <typename... P, typename... A>
static void emu_submit(EmuCfg cfg, void (*kernel)(P...), A &&...args)
{
    /* the parameters are converted and copied NOW (a launch evaluates its arguments when it is issued) */
    auto pack = std::make_shared<std::tuple<std::decay_t<P>...>>(static_cast<std::decay_t<P>>(std::forward<A>(args))...);
    emu_enqueue([cfg, kernel, pack] {
        emu_launches++;
        if (cfg.grid.x == 0 || cfg.grid.y == 0 || cfg.grid.z == 0) return;     /* (an error on the device; never issued) */
        emu_launch_sync(cfg.grid, cfg.block, [&] { std::apply(kernel, *pack); });
    });
}

/* memory operations that can sit inside a captured region are recorded too */
static inline cudaError_t emu_memset_async(void *d, int v, size_t n)
{
    emu_enqueue([d, v, n] { memset(d, v, n); });
    return cudaSuccess;
}
static inline cudaError_t emu_memcpy_async(void *d, const void *s, size_t n)
{
    emu_enqueue([d, s, n] { memmove(d, s, n); });
    return cudaSuccess;
}
#define cudaMemsetAsync(d, v, n, stream) emu_memset_async((d), (v), (n))
#define cudaMemcpyAsync(d, s, n, kind, stream) emu_memcpy_async((d), (s), (n))
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
#define cudaMemcpyToSymbolAsync(sym, src, n, off, kind, stream) (memcpy((char *)(sym) + (off), (src), (n)), cudaSuccess)

static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }

static inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? cudaSuccess : 2; }
enum { cudaHostAllocMapped = 2, cudaHostAllocDefault = 0, cudaHostAllocPortable = 1 };
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMallocHost(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostGetDevicePointer(void **d, void *h, unsigned) { *d = h; return cudaSuccess; }
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void *devicePointer, *hostPointer; };
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *)
{
    a->type = cudaMemoryTypeUnregistered;      /* plain pageable memory: the library stages it itself */
    a->device = 0;
    a->devicePointer = a->hostPointer = nullptr;
    return cudaSuccess;
}

/* one device with 148 multiprocessors, four resident blocks of the persistent kernel each */
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrCooperativeLaunch = 95 };
static inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int) { *v = a == cudaDevAttrMultiProcessorCount ? 148 : 0; return cudaSuccess; }
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename K> static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
template <typename K> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, K, int, size_t) { *n = 4; return cudaSuccess; }
/* cudaLaunchCooperativeKernel((const void *)kernel, ...) is rewritten into this by tests/emu/evp_emu.py (the typed pointer) */
template <typename A>
static cudaError_t emu_launch_cooperative(void (*kernel)(A), dim3 grid, dim3 block, void **args, size_t smem, cudaStream_t)
{
    auto pack = std::make_shared<std::decay_t<A>>(*reinterpret_cast<std::decay_t<A> *>(args[0]));
    emu_enqueue([=] {
        emu_launches++;
        emu_launch_grid(grid, block, smem, &emu_evp_dyn_smem, [&] { kernel(*pack); });
    });
    return cudaSuccess;
}

/* peer-to-peer halo exchange: needs a second process with a device of its own -- not emulated */
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return 903; }
static inline cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return 903; }
static inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
static inline cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) { *can = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }

#endif
