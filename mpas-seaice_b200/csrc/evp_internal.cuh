// evp_internal.cuh -- handle layout and helpers shared by the translation units of libevp_b200.so.
//
// Device data layout.  The cell / vertex index is always the FASTEST dimension, so a warp of 32
// consecutive cells reads 32 consecutive elements of every row.
//
// Static basis arrays are TILED: the rows of a tile of EVP_TILE = 32 cells are contiguous
// (element (row, c) lives at ((c/32)*nRows + row)*32 + c%32, see evp_tix), so that one block of the
// cell kernel fetches the whole basis of its 32 cells with three bulk copies (cp.async.bulk, 9-18 KB
// each) into shared memory:
//   Gb   [tile][k*M+j][32] double2, k = 0..2: the three non-zero (basisGradientU,V)(i,j,c), i in
//        {j-1, j, j+1} cyclic (Wachspress, wachspress.F:1178-1191) in increasing-i order   48*M B / cell
//   G    [tile][j*M+i][32] double2 = (basisGradientU(i,j,c), basisGradientV(i,j,c))     16*M*M B / cell
//        -- only while the basis is being built, and kept instead of Gb when the gradients are dense (PWL)
//   Suv  [tile][j*M+i][32] double2 = (basisIntegralsU(i,j,c), basisIntegralsV(i,j,c))   16*M*M B / cell
//   Sm   [tile][j*M+i][32] double  =  basisIntegralsMetric(i,j,c)                         8*M*M B / cell
// State and connectivity are plain row-SoA with row stride nCp:
//   sig  [i][c]    double2 = (stress11(i,c), stress22(i,c));  sig12[i][c] double
//   contrib[j][c]  double2 = per-cell partial sums (stressDivergenceUCell, stressDivergenceVCell)
//                            of reference variational.F:1151-1173 for velocity vertex slot j
//   voc  [i][c]    int     = verticesOnCell(i,c)-1
//   uv   [v]       double2 = (uVelocity(v), vVelocity(v))
//   vertex fields are packed in double2 pairs (U,V) / (mass, mass*f) / (iceArea, denominator).
//
// c strides are padded to a multiple of 64 elements (nCp, nVp) so every row starts 512-B aligned.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/evp_b200.h"

void evp_set_error(const char *fmt, ...);

constexpr int EVP_TILE = 32;
// element index of (row, cell) in a tiled basis array with nRows rows per tile
__host__ __device__ __forceinline__ size_t evp_tix(int row, size_t c, int nRows)
{
    return ((c / EVP_TILE) * (size_t)nRows + (size_t)row) * EVP_TILE + (c % EVP_TILE);
}

#define EVP_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            evp_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return EVP_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

#define EVP_REQUIRE(cond, msg)                                     \
    do {                                                           \
        if (!(cond)) {                                             \
            evp_set_error("%s:%d: %s", __FILE__, __LINE__, msg);   \
            return EVP_ERR_ARGUMENT;                               \
        }                                                          \
    } while (0)

struct evp_halo;  // evp_halo.cu
int evp_enqueue_weak_cell_pass(evp_handle *h, bool diag, cudaStream_t s);    // strain [+ stress] on cells
int evp_enqueue_weak_to_variational(evp_handle *h, cudaStream_t s);          // interpolate_strains_weak_to_variational

// evp_halo.cu

struct evp_dev {
    // static
    uint8_t *nEdges = nullptr;
    int *voc = nullptr;
    double2 *G = nullptr, *Gb = nullptr, *Suv = nullptr;
    double *Sm = nullptr;
    double *tanLat = nullptr;
    int *gidx = nullptr;          // [D][nVp] index into contrib (j*nCp + c) or -1
    int *cov = nullptr;           // [D][nVp] cellsOnVertex - 1, or -1 when not a local cell
    // pre-/post-subcycle mesh extension (evp_set_mesh_ext)
    int *coc = nullptr;           // [M][nCp] cellsOnCell - 1 or -1
    uint8_t *vflags = nullptr;    // bit 0 interiorVertex, bit 1 landIceMaskVertex
    double *areaCell = nullptr, *areaTri = nullptr, *fVertex = nullptr;
    double2 *airCell = nullptr;   // (airStressCellU, airStressCellV) of the current step
    double2 *osFinal = nullptr;   // ocean_stress_final's oceanStressU/V
    uint8_t *solveVelPrev = nullptr;
    // evp_aggregate: aggregate_mass_and_area and the Hibler strength on the device (allocated on first use)
    double *aggArea = nullptr, *aggVolIce = nullptr, *aggVolSnow = nullptr, *aggMass = nullptr, *aggP = nullptr;
    // weak operators (evp_set_weak_mesh): per cell-slot / vertex-slot edge data, one stress point per cell
    int2 *wEdgeV = nullptr;       // [M][nCp] the two vertices of edge k of the cell (0-based)
    double2 *wNp = nullptr;       // [M][nCp] normalVectorPolygon
    double *wDv = nullptr;        // [M][nCp] dvEdge of that edge
    int2 *wEdgeC = nullptr;       // [D][nVp] the two cells of edge s of the vertex (0-based, -1 = none)
    double2 *wNt = nullptr;       // [D][nVp] normalVectorTriangle
    double *wDc = nullptr;        // [D][nVp] dcEdge
    double *wTanC = nullptr, *wTanV = nullptr;     // tan(latCellRotated), tan(latVertexRotated) (host libm)
    double *wAreaC = nullptr, *wAreaT = nullptr;   // areaCell, areaTriangle
    double2 *sigW = nullptr;      // (stress11Weak, stress22Weak)
    double *sigW12 = nullptr, *eW11 = nullptr, *eW22 = nullptr, *eW12 = nullptr, *repPW = nullptr;
    double *eV11 = nullptr, *eV22 = nullptr, *eV12 = nullptr;   // strainXXVertex of interpolate_strains_weak_to_variational
    double wRadius = 1.0;
    // per step
    uint8_t *solveStress = nullptr, *solveVel = nullptr;
    uint8_t *tileWork = nullptr;  // per tile of EVP_TILE cells: 1 = some cell is solved or holds a non-zero stress
    int *tileList = nullptr;      // the tiles with work, compacted (cell kernel grid = their number)
    int *tileCount = nullptr;     // device counter behind tileList
    unsigned *gridBar = nullptr;  // persistent whole-loop kernel: release word + one arrival slot per block
    double2 *contrib2 = nullptr;  // ... and its second buffer of the divergence sums (allocated on first use)
    uint8_t *vblockWork = nullptr;  // per block of 256 owned vertices: 1 = some vertex is solved
    int *vblockList = nullptr, *vblockCount = nullptr;
    double *P = nullptr;
    double2 *uv = nullptr, *sig = nullptr, *contrib = nullptr;
    double *sig12 = nullptr;
    double2 *massf = nullptr, *air = nullptr, *tilt = nullptr, *ocnStress = nullptr, *ocnVel = nullptr;
    double2 *areaDen = nullptr;   // (iceAreaVertex, variationalDenominator)
    double2 *uvInit = nullptr;
    // outputs of the last subcycle
    double *e11 = nullptr, *e22 = nullptr, *e12 = nullptr, *repP = nullptr;
    double2 *sdiv = nullptr;
    double *ocoef = nullptr;
    // special boundaries (resolved to pre-loop sources, see evp_abi.cu)
    int nSB = 0;
    int *sbDst = nullptr, *sbSrc = nullptr;
    double *sbSign = nullptr;
    double2 *sbTmp = nullptr;
    // staging
    void *stage = nullptr;
    size_t stageBytes = 0;
};

struct evp_handle {
    int device = 0;
    int nCells = 0, nCellsSolve = 0, nVertices = 0, nVerticesSolve = 0, D = 0;
    int Mh = 0;                   // maxEdges of the host arrays
    int M = 0;                    // slots per cell of the device layout / kernel instantiation (4, 6 or 8)
    size_t nCp = 0, nVp = 0;
    int nActiveTiles = -1;        // tiles with work (host copy); -1 = not computed yet
    int nActiveVBlocks = -1;      // 256-vertex blocks with a solved vertex; -1 = not computed yet
    evp_options opt{};
    bool metric = false;          // any tanLatVertexRotatedOverRadius != 0
    bool haveExt = false, haveWeak = false;
    bool haveAgg = false, haveAggP = false;   // evp_aggregate ran (with the Hibler strength)
    bool haveBasis = false, haveStep = false, haveSB = false, useGraph = true, pinHost = false, timed = false;
    evp_dev d;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaGraphExec_t graphExec = nullptr;
    int graphN = -1;
    float lastMs = 0.f;
    unsigned long long devBytes = 0;
    std::vector<void *> allocs;
    std::vector<void *> pinned;   // host ranges registered with cudaHostRegister
    void *pinStage[2] = {nullptr, nullptr};
    cudaEvent_t pinEv[2] = {nullptr, nullptr};
    int pinNext = 0;
    evp_halo *halo = nullptr;
};

// evp_kernels.cu
int evp_enqueue_subcycles(evp_handle *h, int nSub, cudaStream_t s);
int evp_count_launches(evp_handle *h, int nSub);
// all nSub subcycles as ONE cooperative launch when the whole mesh can be resident (small meshes); 0 = launched (or
// would be, with probeOnly), -1 = not eligible: use the two-kernel graph
int evp_persistent_run(evp_handle *h, int nSub, cudaStream_t s, bool probeOnly);
int evp_enqueue_cell_pass(evp_handle *h, bool diag, cudaStream_t s);
int evp_enqueue_vertex_pass(evp_handle *h, bool diag, cudaStream_t s);
int evp_enqueue_special_boundaries(evp_handle *h, cudaStream_t s);
// recompute tileWork from the masks and stresses on the device (and zero contrib of tiles without work); must
// follow every change of solveStress / sig outside the subcycle kernels
int evp_refresh_tile_flags(evp_handle *h, cudaStream_t s);

// evp_weak.cu
int evp_enqueue_weak_cell_pass(evp_handle *h, bool diag, cudaStream_t s);    // strain [+ stress] on cells
int evp_enqueue_weak_to_variational(evp_handle *h, cudaStream_t s);          // interpolate_strains_weak_to_variational

// evp_halo.cu
// The peer-to-peer halo exchange as the kernels see it (all-default without it, see evp_halo.cu for the protocol).
struct evp_halo_view {
    const int *ctr = nullptr;       // completed vertex passes of this rank (device counter)
    const int *flagsIn = nullptr;   // [nNb] completed vertex passes of each neighbour, stored by them over NVLink
    int *err = nullptr;             // mapped host flag: a wait ran into its time limit
    int nNb = 0;
    int haloFirst = 0x7fffffff;     // = nVerticesSolve: the first halo vertex
    int shift0 = 0;                 // halo vertex v of an even pass lives at element v + shift0 ...
    int stride = 0;                 // ... of an odd pass at v + shift0 + stride
};
struct evp_push_view {              // what the vertex kernel needs to store boundary-owned (u,v) into the neighbours
    int *ctr = nullptr;
    unsigned *done = nullptr;       // block tickets: the last block of the pass publishes the flags
    const int *bStart = nullptr;    // [vertex blocks + 1] boundary vertices before each block of 256 owned vertices
    const int *pushStart = nullptr; // [nBoundary + 1] CSR over the boundary vertices in ascending order
    const int2 *push = nullptr;     // (neighbour slot, element index in that neighbour's array for an even pass)
    double2 *const *peerUv = nullptr;   // [nNb] the neighbour's velocity array (peer mapping)
    const int *peerStride = nullptr;    // [nNb] its distance between the two halo buffers
    int *const *peerFlag = nullptr;     // [nNb] this rank's slot among the neighbour's incoming flags
    int nNb = 0;
    int nPushBlocks = 0;                // blocks of 256 owned vertices that hold a boundary vertex (the ticket count)
    const int *order = nullptr;         // [vertex blocks] launch order: the pushing blocks first, so that their stores, fences
                                        // and the publication of the pass overlap the interior vertices' work
    int dbg = 0;                        // EVP_B200_P2P_DEBUG: timing experiments only (results are NOT valid with it)
};
int evp_halo_enqueue(evp_handle *h, cudaStream_t s);             // in-loop exchange of d.uv on the NCCL path
int evp_halo_exchange(evp_handle *h, cudaStream_t s, double2 *field);   // any (nVp) double2 vertex field, NCCL
int evp_halo_launches(evp_handle *h);
int evp_halo_mark_masks(evp_handle *h);
int evp_halo_boundary_count(evp_handle *h);        // boundary-owned vertices (unique send-list entries)
bool evp_halo_p2p_active(evp_handle *h);
evp_halo_view evp_halo_get_view(evp_handle *h);
evp_push_view evp_halo_get_push(evp_handle *h);                  // ctr == nullptr when not active
int evp_halo_begin_run(evp_handle *h, cudaStream_t s);           // around every run of subcycles (no-ops on the NCCL path)
int evp_halo_end_run(evp_handle *h, cudaStream_t s);
int evp_halo_check(evp_handle *h);                 // EVP_ERR_NCCL after a timed-out wait
void evp_halo_destroy(evp_handle *h);

constexpr unsigned long long EVP_HALO_WAIT_NS = 300ull * 1000ull * 1000ull * 1000ull;
#ifdef __CUDACC__
__device__ __forceinline__ int evp_ld_acquire_sys(const int *p)
{
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void evp_st_relaxed_sys(int *p, int v)      // ordered by a preceding __threadfence_system()
{
    asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long evp_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Block until every neighbour has published vertex pass c (flags only grow; the comparison survives wrap-around).
// Bounded: after EVP_HALO_WAIT_NS (5 minutes: ranks of a real host may be skewed by I/O far longer than a bench's) the mapped
// error flag is raised and the caller proceeds; the host reports it at the next evp_synchronize / evp_fetch.
__device__ __forceinline__ void evp_halo_wait(const evp_halo_view &hv, int c)
{
    for (int i = 0; i < hv.nNb; i++) {
        if (evp_ld_acquire_sys(hv.flagsIn + i) - c >= 0) continue;
        if (*(volatile int *)hv.err) return;              // an earlier wait already gave up: do not queue further long waits
        const unsigned long long t0 = evp_globaltimer();
        while (evp_ld_acquire_sys(hv.flagsIn + i) - c < 0) {
            __nanosleep(100);
            if (evp_globaltimer() - t0 > EVP_HALO_WAIT_NS) {
                *(volatile int *)hv.err = 1;
                __threadfence_system();
                return;
            }
        }
    }
}
#endif

inline unsigned grid_for(size_t n, int block) { return (unsigned)((n + block - 1) / block); }

struct Stage {   // bump allocator over the device staging area
    char *base; size_t cap, off;
    void *take(size_t bytes) {
        size_t o = (off + 255) & ~(size_t)255;
        if (o + bytes > cap) return nullptr;
        off = o + bytes;
        return base + o;
    }
};

// layout kernels and transfers (evp_abi.cu)
int evp_h2d(evp_handle *h, void *dst, const void *src, size_t bytes);
int evp_d2h(evp_handle *h, void *dst, const void *src, size_t bytes);
int evp_upload_rows(evp_handle *h, const double *host, double *dst, int dims, int ncomp, int comp);
int evp_download_rows(evp_handle *h, double *host, const double *soa, int dims, int ncomp, int comp);
int evp_dev_alloc(evp_handle *h, void **p, size_t bytes);
void evp_dev_free(evp_handle *h, void *p, size_t bytes);
// dense G (re)allocated and zeroed, ready to be filled; then band-compressed when the pattern allows
int evp_basis_begin(evp_handle *h);
int evp_basis_finalize(evp_handle *h);
