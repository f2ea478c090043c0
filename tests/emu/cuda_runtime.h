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

#include <ucontext.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
/* shared memory: one static array per declaration, shared by the threads of the block being executed (blocks run one
 * after the other).  Kernels launched with IR_LAUNCH use it as per-thread scratch columns only; kernels that exchange
 * data between threads through it are launched with IR_LAUNCH_SYNC (below), where __syncthreads() is a real barrier. */
#define __shared__ static
/* dynamic shared memory: one buffer, big enough for any launch of the file under test */
static double emu_dyn_smem[64 * 1024];
#define IR_DYN_SHARED(type, name) type *name = reinterpret_cast<type *>(emu_dyn_smem)
#define __ldg(p) (*(p))

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static thread_local dim3 blockIdx, threadIdx, blockDim, gridDim;

/* Threads of a launch run one after the other.  IR_EMU_ORDER=reverse runs blocks and threads last to first: a kernel
 * whose threads depend on one another within a launch (a race on the device) gives different results in the two
 * orders, so comparing them is a cheap race detector (tests/test_ir_parity.py). */
template <typename F>
static void emu_launch(dim3 grid, dim3 block, F body)
{
    gridDim = grid;
    blockDim = block;
    const char *order = getenv("IR_EMU_ORDER");
    const bool reverse = order != nullptr && order[0] == 'r';
    const unsigned long long nb = (unsigned long long)grid.x * grid.y * grid.z, nt = (unsigned long long)block.x * block.y * block.z;
    for (unsigned long long ib = 0; ib < nb; ib++) {
        const unsigned long long b = reverse ? nb - 1 - ib : ib;
        for (unsigned long long it = 0; it < nt; it++) {
            const unsigned long long t = reverse ? nt - 1 - it : it;
            blockIdx = dim3((unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((unsigned long long)grid.x * grid.y)));
            threadIdx = dim3((unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / ((unsigned long long)block.x * block.y)));
            body();
        }
    }
}
#define IR_LAUNCH(kernel, grid, block, stream, ...) emu_launch(dim3(grid), dim3(block), [&] { kernel(__VA_ARGS__); })

/* Kernels with __syncthreads(): every thread of a block is a fiber (ucontext) with a stack of its own; a fiber runs
 * until it reaches a barrier or returns, then the next one runs; when all living fibers of the block wait at the
 * barrier the round starts over.  Same order switch as above (IR_EMU_ORDER=reverse). */
struct EmuFiber {
    ucontext_t ctx;
    char *stack = nullptr;
    bool done = true;
    long orCount = 0;      /* number of __syncthreads_or() this fiber has passed in the current block */
};
static std::vector<EmuFiber> emu_fibers;
static ucontext_t emu_sched_ctx;
static std::function<void()> *emu_body = nullptr;
static unsigned long long emu_current = 0;
static const size_t EMU_STACK = 256 * 1024;

static void emu_trampoline()
{
    (*emu_body)();
    emu_fibers[emu_current].done = true;      /* uc_link brings control back to the scheduler */
}
/* grid mode (a cooperative launch, emu_launch_grid below): the fibers of ALL blocks are alive at the same time */
struct EmuGridFiber {
    ucontext_t ctx;
    bool done = true, atBarrier = false;
    unsigned block = 0, thread = 0;
};
static bool emu_grid_mode = false;
static std::vector<EmuGridFiber> emu_gfibers;
static unsigned long long emu_gcurrent = 0;
static inline void emu_syncthreads()
{
    if (emu_grid_mode) {
        emu_gfibers[emu_gcurrent].atBarrier = true;
        swapcontext(&emu_gfibers[emu_gcurrent].ctx, &emu_sched_ctx);
        return;
    }
    swapcontext(&emu_fibers[emu_current].ctx, &emu_sched_ctx);
}
/* a thread that polls memory written by another block lets the others run */
static inline void emu_yield()
{
    if (emu_grid_mode) swapcontext(&emu_gfibers[emu_gcurrent].ctx, &emu_sched_ctx);
}
#define __syncthreads() emu_syncthreads()
/* __syncthreads_or: two accumulators used alternately; the first fiber to arrive for a generation clears its slot */
static int emu_or_acc[2] = {0, 0};
static long emu_or_gen[2] = {-1, -1};
static inline int emu_syncthreads_or(int pred)
{
    EmuFiber &f = emu_fibers[emu_current];
    const long gen = f.orCount++;
    const int slot = (int)(gen & 1);
    if (emu_or_gen[slot] != gen) { emu_or_gen[slot] = gen; emu_or_acc[slot] = 0; }
    if (pred) emu_or_acc[slot] = 1;
    emu_syncthreads();
    return emu_or_acc[slot];
}
#define __syncthreads_or(p) emu_syncthreads_or(p)

template <typename F>
static void emu_launch_sync(dim3 grid, dim3 block, F body)
{
    gridDim = grid;
    blockDim = block;
    const char *order = getenv("IR_EMU_ORDER");
    const bool reverse = order != nullptr && order[0] == 'r';
    const unsigned long long nb = (unsigned long long)grid.x * grid.y * grid.z, nt = (unsigned long long)block.x * block.y * block.z;
    if (emu_fibers.size() < nt) emu_fibers.resize(nt);
    for (unsigned long long t = 0; t < nt; t++)
        if (!emu_fibers[t].stack) emu_fibers[t].stack = (char *)malloc(EMU_STACK);
    std::function<void()> fn = body;
    emu_body = &fn;
    for (unsigned long long ib = 0; ib < nb; ib++) {
        const unsigned long long b = reverse ? nb - 1 - ib : ib;
        blockIdx = dim3((unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((unsigned long long)grid.x * grid.y)));
        for (unsigned long long t = 0; t < nt; t++) {
            EmuFiber &f = emu_fibers[t];
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = EMU_STACK;
            f.ctx.uc_link = &emu_sched_ctx;
            makecontext(&f.ctx, emu_trampoline, 0);
            f.done = false;
            f.orCount = 0;
        }
        emu_or_gen[0] = emu_or_gen[1] = -1;
        unsigned long long alive = nt;
        while (alive > 0) {
            for (unsigned long long it = 0; it < nt; it++) {
                const unsigned long long t = reverse ? nt - 1 - it : it;
                EmuFiber &f = emu_fibers[t];
                if (f.done) continue;
                threadIdx = dim3((unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / ((unsigned long long)block.x * block.y)));
                emu_current = t;
                swapcontext(&emu_sched_ctx, &f.ctx);
                if (f.done) alive--;
            }
        }
    }
    emu_body = nullptr;
}
#define IR_LAUNCH_SYNC(kernel, grid, block, smem, stream, ...) \
    emu_launch_sync(dim3(grid), dim3(block), [&] { kernel(__VA_ARGS__); })

/* Cooperative launch: every thread of every block is a fiber of its own, resumed round-robin (IR_EMU_ORDER=reverse: last
 * to first).  __syncthreads() holds a fiber until all living fibers of ITS block have arrived; a fiber spinning on a word
 * another block will write calls emu_yield() in the loop (the acquire-load helpers do).  Dynamic shared memory: one
 * buffer per block, *smemVar points at the running block's before every resumption (kernels take the address once, at
 * their start).  A launch that makes no progress for a long time aborts: a grid barrier that cannot complete. */
static void emu_gtrampoline()
{
    (*emu_body)();
    emu_gfibers[emu_gcurrent].done = true;
}
template <typename F>
static void emu_launch_grid(dim3 grid, dim3 block, size_t smemBytes, unsigned char **smemVar, F body)
{
    gridDim = grid;
    blockDim = block;
    const char *order = getenv("IR_EMU_ORDER");
    const bool reverse = order != nullptr && order[0] == 'r';
    const unsigned long long nb = (unsigned long long)grid.x * grid.y * grid.z, nt = (unsigned long long)block.x * block.y * block.z;
    const unsigned long long total = nb * nt;
    const size_t stackBytes = 64 * 1024, smemPitch = (smemBytes + 127) & ~(size_t)127;
    char *stacks = (char *)malloc(total * stackBytes);
    unsigned char *smem = (unsigned char *)aligned_alloc(128, smemPitch * nb + 128);
    if (!stacks || !smem) abort();
    {   /* shared memory of a fresh block is not clear on a device either: 0xFF bytes under IR_EMU_POISON=1 */
        const char *e = getenv("IR_EMU_POISON");
        memset(smem, (e && e[0] == '1') ? 0xFF : 0, smemPitch * nb + 128);
    }
    emu_gfibers.assign(total, EmuGridFiber());
    std::function<void()> fn = body;
    emu_body = &fn;
    std::vector<unsigned long long> alive(nb, nt), arrived(nb, 0);
    for (unsigned long long i = 0; i < total; i++) {
        EmuGridFiber &f = emu_gfibers[i];
        f.block = (unsigned)(i / nt);
        f.thread = (unsigned)(i % nt);
        f.done = false;
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = stacks + i * stackBytes;
        f.ctx.uc_stack.ss_size = stackBytes;
        f.ctx.uc_link = &emu_sched_ctx;
        makecontext(&f.ctx, emu_gtrampoline, 0);
    }
    unsigned char *savedSmem = *smemVar;
    emu_grid_mode = true;
    unsigned long long remaining = total, idle = 0;
    while (remaining > 0) {
        bool ran = false, changed = false;
        for (unsigned long long it = 0; it < total; it++) {
            const unsigned long long i = reverse ? total - 1 - it : it;
            EmuGridFiber &f = emu_gfibers[i];
            if (f.done || f.atBarrier) continue;
            const unsigned long long b = f.block, t = f.thread;
            blockIdx = dim3((unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((unsigned long long)grid.x * grid.y)));
            threadIdx = dim3((unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / ((unsigned long long)block.x * block.y)));
            *smemVar = smem + smemPitch * b;
            emu_gcurrent = i;
            swapcontext(&emu_sched_ctx, &f.ctx);
            ran = true;
            if (f.done) { remaining--; alive[b]--; changed = true; }
            else if (f.atBarrier) { arrived[b]++; changed = true; }
        }
        for (unsigned long long b = 0; b < nb; b++) {
            if (alive[b] > 0 && arrived[b] == alive[b]) {
                for (unsigned long long t = 0; t < nt; t++) emu_gfibers[b * nt + t].atBarrier = false;
                arrived[b] = 0;
                changed = true;
            }
        }
        idle = changed ? 0 : idle + 1;
        if (!ran || idle > 100000) {
            fprintf(stderr, "emu_launch_grid: no progress (%llu threads left): a barrier that cannot complete\n", remaining);
            abort();
        }
    }
    emu_grid_mode = false;
    *smemVar = savedSmem;
    emu_body = nullptr;
    emu_gfibers.clear();
    free(stacks);
    free(smem);
}

typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1 };

static inline const char *cudaGetErrorString(cudaError_t) { return "emulation"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
/* cudaMalloc does not clear memory.  IR_EMU_POISON=1 fills every allocation with 0xFF bytes (NaN as a double, -1 as an int):
 * code that counts on fresh device memory being zero gives wrong results here instead of passing by luck. */
static inline cudaError_t cudaMalloc(void **p, size_t n)
{
    *p = calloc(n ? n : 1, 1);
    if (*p) {
        static const int poison = [] { const char *e = getenv("IR_EMU_POISON"); return (e && e[0] == '1') ? 1 : 0; }();
        if (poison) memset(*p, 0xFF, n ? n : 1);
    }
    return *p ? cudaSuccess : 2;
}
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
enum { cudaHostRegisterDefault = 0 };
static inline cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
static inline int atomicOr(int *p, int v) { const int old = *p; *p = old | v; return old; }
static inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) { const unsigned long long old = *p; if (v < old) *p = v; return old; }

#endif
