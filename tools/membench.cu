// tools/membench.cu -- layout experiment behind DESIGN.md section 2: how fast can a B200 stream R
// "rows" per cell when the rows are (a) R separate SoA arrays (stride nC) or (b) tiled so that the
// R rows of a tile of T cells are contiguous.  One thread per (cell, slot) like evp_cell_kernel:
// blockDim = (CB, S); slot s reads rows s*R/S .. (s+1)*R/S-1.  Read-only (one double2 store per thread).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membench tools/membench.cu && ./membench
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int T>
__device__ __forceinline__ size_t ix(int row, size_t c, int R, size_t nC)
{
    if (MODE == 0) return (size_t)row * nC + c;
    return ((c / T) * R + row) * T + (c % T);
}
template <int MODE, int T, int CB, int S, int RPS>
__global__ void __launch_bounds__(CB *S) k_read(const double2 *__restrict__ a, double2 *__restrict__ out, size_t nC)
{
    const size_t c = (size_t)blockIdx.x * CB + threadIdx.x;
    const int s = threadIdx.y;
    if (c >= nC) return;
    double x = 0, y = 0;
#pragma unroll
    for (int r = 0; r < RPS; r++) {
        const double2 v = a[ix<MODE, T>(s * RPS + r, c, S * RPS, nC)];
        x += v.x; y += v.y;
    }
    out[ix<MODE, T>(s, c, S, nC)] = make_double2(x, y);
}
template <int MODE, int T, int CB, int S, int RPS>
void run(const char *name, const double2 *a, double2 *out, size_t nC)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dim3 block(CB, S);
    const unsigned grid = (unsigned)((nC + CB - 1) / CB);
    for (int i = 0; i < 3; i++) k_read<MODE, T, CB, S, RPS><<<grid, block>>>(a, out, nC);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int i = 0; i < reps; i++) k_read<MODE, T, CB, S, RPS><<<grid, block>>>(a, out, nC);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    const double bytes = (double)nC * 16.0 * (S * RPS + S);
    printf("%-44s rows=%3d  %.3f ms  %.0f GB/s  (%s)\n", name, S * RPS, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    const size_t nC = 10485760;            // multiple of 64
    const int R = 96;
    double2 *a, *out;
    cudaMalloc(&a, nC * R * sizeof(double2));
    cudaMalloc(&out, nC * 8 * sizeof(double2));
    cudaMemset(a, 0, nC * R * sizeof(double2));
    run<0, 32, 64, 6, 16>("SoA rows, block (64,6), 16 rows/thread", a, out, nC);
    run<1, 32, 64, 6, 16>("tiled T=32, block (64,6), 16 rows/thread", a, out, nC);
    run<1, 64, 64, 6, 16>("tiled T=64, block (64,6), 16 rows/thread", a, out, nC);
    run<0, 32, 32, 6, 16>("SoA rows, block (32,6), 16 rows/thread", a, out, nC);
    run<1, 32, 32, 6, 16>("tiled T=32, block (32,6), 16 rows/thread", a, out, nC);
    run<0, 32, 64, 6, 8>("SoA rows, block (64,6), 8 rows/thread", a, out, nC);
    run<1, 32, 64, 6, 8>("tiled T=32, block (64,6), 8 rows/thread", a, out, nC);
    run<0, 32, 64, 6, 2>("SoA rows, block (64,6), 2 rows/thread", a, out, nC);
    run<1, 32, 64, 6, 2>("tiled T=32, block (64,6), 2 rows/thread", a, out, nC);
    run<0, 32, 256, 1, 1>("SoA rows, block (256,1), 1 row (stream copy)", a, out, nC);
    run<0, 32, 256, 1, 8>("SoA rows, block (256,1), 8 rows/thread", a, out, nC);
    run<1, 32, 256, 1, 8>("tiled T=32, block (256,1), 8 rows/thread", a, out, nC);
    run<0, 32, 128, 1, 32>("SoA rows, block (128,1), 32 rows/thread", a, out, nC);
    run<1, 32, 128, 1, 32>("tiled T=32, block (128,1), 32 rows/thread", a, out, nC);
    return 0;
}
