// ir_kernels.cu -- incremental-remapping transport on B200 (sm_100a): kernels and the C ABI of include/ir_b200.h.
//
// What one ir_run does (reference: incremental_remap_block, src/shared/mpas_seaice_advection_incremental_remap.F:2740):
//   k_prepare           per cell: ice mask, volume -> thickness                                    (:2462-2480, make_masks :3404)
//   k_reconstruct_coop  block = 32 cells x all rows: gradient, limiter, centre value, barycentre   (:3580-5250)
//   k_triangles         per edge: departure triangles and their quadrature points                  (:5255-6665)
//   k_fluxes_coop       block = 32 edges x the rows of one category: mass * tracer over the triangles (:6667-6980)
//   k_update_coop       block = 32 cells x all rows: new mass and tracers, zap, thickness -> volume (:6982-7540, :8764-8895, :2680-2700)
// Five launches per step whatever the number of tracers.
//
// Layout: the host keeps a tracer as (nLayers, nCategories, nCells) -- all components of a cell together.  On the
// device every (tracer, category, layer) is a ROW of one matrix val[nRows][nCp] with the cell index fastest, so a warp
// of 32 cells reads 256 contiguous bytes of a row; the same for every per-cell / per-edge geometry array
// ([slot][nCp]).  Parents are resolved to row numbers once, in ir_set_tracers.  Categories never mix, and within a
// category a row depends on its parents in the same cell only.  The three tracer kernels are cooperative: a warp is
// 32 consecutive cells (or edges) of ONE row, the rows of the tracer hierarchy are spread over the thread rows of the
// block and walked depth by depth with a barrier in between, and what all rows share -- the cell's geometry, the
// departure triangles' quadrature points, the mass reconstruction at those points -- is staged or computed once per
// block in shared memory.  (Round 1 ran one thread per (cell or edge, category) walking its 23 rows serially: 8.4 ms
// per QU60 step, 78 % of it in the flux kernel at 27 % occupancy and 9 % L1 hit rate; now 2.85 ms, profiles/ir_r02_*.)
//
// All of it is gather-heavy FP64 streaming work bounded by HBM / L2 and, in the flux kernel, the FP64 pipe -- not tensor
// work.  Built with --fmad=false and in the reference's operation order so that the results are bit-identical to
// oracle/ir_oracle.c (tests/test_ir_parity.py).
#include <cuda_runtime.h>

#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "../../include/ir_b200.h"

// Kernel launches go through one macro so that tests/emu can compile this file for the host and step through the
// kernels thread by thread (tests/test_ir_parity.py, "emulation" leg: a check of the kernel logic where no GPU is
// available; it is test infrastructure and never part of the shipped library).
#ifndef IR_LAUNCH
#define IR_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
// kernels whose threads cooperate through shared memory and __syncthreads() (the emulation runs them as fibers)
#define IR_LAUNCH_SYNC(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define IR_DYN_SHARED(type, name)                                  \
    extern __shared__ __align__(16) unsigned char ir_dyn_smem_raw[]; \
    type *name = reinterpret_cast<type *>(ir_dyn_smem_raw)
#define IR_DEVICE_BUILD 1
#endif

namespace {

constexpr double EPS11 = 1.0e-11;
constexpr double W1QP = 1.09951743655321885e-01, W2QP = 2.23381589678011389e-01;
constexpr double Q1QP = 9.15762135097710761e-02, Q2QP = 8.16847572980458514e-01;
constexpr double Q3QP = 1.08103018168070275e-01, Q4QP = 4.45948490915965612e-01;
constexpr int NTRI = IR_N_TRI_PER_EDGE, NCER = 6, NEER = 6, NVER = 8;
constexpr int MAXM = 8;        // maxEdges supported
constexpr int MAX_DEPTH = 4;   // mass + three parents (incremental_remap.F:6745)

enum { FLAG_NEG_QP = 1, FLAG_NEG_MASS = 2, FLAG_PARALLEL = 4, FLAG_MANY_TRI = 8 };

thread_local char g_err[512] = "";
void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

#define IR_CUDA(call)                                                                        \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return IR_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)
#define IR_REQUIRE(cond, msg)                                  \
    do {                                                       \
        if (!(cond)) {                                         \
            set_error("%s:%d: %s", __FILE__, __LINE__, msg);   \
            return IR_ERR_ARGUMENT;                            \
        }                                                      \
    } while (0)

struct RowInfo {
    int chain[MAX_DEPTH];  // rows of the mass field .. this row (mass first); chain[depth] == this row
    int depth;             // number of parents
    int hasChild;
    int cat;               // category of the row
    int volumeLike;
    int slot;              // row in the arrays kept for parents only (xBary, yBary, mtpNew), -1 for a childless row
    int parentSlot;        // the parent's slot, -1 for the mass-like field
};

// One row of a category in parents-first order; the table is the same for every category k, whose device row is
// baseRow + k * layers.  anc[q] is the position IN THIS TABLE of the ancestor of depth q (anc[depth] = the row itself).
struct CatRow {
    int baseRow, layers, depth;
    int cls;                  // index among the category's rows of the same depth (kept for depth 0 and 1), else -1
    int anc[MAX_DEPTH];
    int hasChild, volumeLike;
    int slotJ, parentSlotJ;   // index among the category's rows that have children: this row's / its parent's, or -1
    int vpIdx;                // k_fluxes_coop: which shared partial product the row starts from (its own for depth 0 / 1,
                              // its depth-1 ancestor's below)
    int base2, layers2;       // ... and, for depth >= 2, the device row (category 0) / layer stride of its depth-2 ancestor
};

struct Dev {
    // sizes
    int nC, nCS, nV, nE, M, D, nK, nQP, sphere, rotate;
    size_t nCp, nEp, nVp;
    // mesh / geometry, [slot][pitch]
    int *nEdgesOnCell, *edgesOnCell, *cellsOnCell, *verticesOnCell, *cellsOnEdge, *verticesOnEdge;
    int *remapEdge, *coer, *eoer;
    double *areaCell, *sdc /* signed dcEdge per (slot, cell) */, *coef /* [3*M] */, *trans /* rows 1,2 of transGlobalToCell: [6] */;
    double *xvc, *yvc, *xve, *yve, *geom /* [14] */;
    int *fluxSign;   // [M][nCp]: +1 if the cell is cellsOnEdge(1) of its k-th edge, else -1
    // per step
    double *u, *v;
    int *maskCell, *maskEdge, *iCellTri /* [NTRI][nEp] */;
    double *xq, *yq /* [NTRI*6][nEp] */, *triArea /* [NTRI][nEp] */;
    // tracer state, [nRows][nCp] unless noted
    int nRows, nRowsPerCat;
    RowInfo *rows;
    int *catBaseRow, *catLayers;   // rows of category k in parents-first order: catBaseRow[j] + k * catLayers[j]
    CatRow *catRows;               // the same list with the tree structure resolved (k_fluxes)
    int nD0, nD1, nSlotsPerCat;    // rows of depth 0 / depth 1 / rows with children, per category
    double *val, *valNew, *center, *xGrad, *yGrad, *xBary, *yBary, *mtpNew;
    double *edgeFlux;   // [nRows][nEp]
    int *flags;
    double *stage;      // raw host-layout staging
    size_t stageBytes;
    // optional checks (ir_set_checks): allocated when first used
    double *lmin, *lmax;            // [nRows][nCp]: tracer_local_min_max of the old values
    double *sums;                   // [2][nRows]: sum_tracers before / after the update
    int *rowT, *rowKL;              // row -> tracer index, category * maxLayers + layer (the reference's loop order)
    unsigned long long *monoKey;    // smallest (tracer, cell, category, layer) key among the violations
    double *monoDetail;             // [3]: new value, lower bound, upper bound of one (row, cell)
    int maxLayers;
};

// device state of the upwind option (ir_upwind.cuh)
struct Upw {
    int *interiorEdge;
    double *dvEdge, *nve /* normalVectorEdge, [2 * slot + component][nCp] */;
    double *oldv, *newv;               // [nVars * nK][nCp]: time levels 1 and 2
    double *eflux;                     // [nVars * nK][nEp]: <variable>EdgeFlux
    double *fluxUp;                    // [nK][nEp]: flux_upwind of the variable being advected
    double *pOldNone, *pNewNone;       // [nK][nCp]: the stand-in parent of a variable without one
    double *pFluxNone;                 // [nK][nEp]
    double *edgeVel;                   // [nEp]
};

}  // namespace

struct ir_handle {
    int device;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    cudaEvent_t evK[4];   // between the five kernels of a step (ir_last_kernel_ms)
    float kernelMs[5];
    Dev d;
    std::vector<RowInfo> rows;
    std::vector<int> tracerRow0, tracerLayers, tracerParent, tracerVolume, tracerDepth;
    std::vector<int> depthRow0;   // rows of depth q are [depthRow0[q], depthRow0[q+1])  (rows sorted by depth)
    std::vector<int> rowOrder;    // device row -> (tracer, k, l) linear index on the host side
    std::vector<void *> allocs;
    std::vector<std::pair<const void *, size_t>> pinned;   // host ranges registered under IR_B200_PIN_HOST
    bool pinHost;
    int nSlots;           // rows of the parent-only arrays
    float lastMs;
    long long launches;
    bool haveTracers;
    int checkConservation, checkMonotonicity;   // ir_set_checks
    ir_check_report report;                     // of the last ir_run
    std::vector<double> sumsHost;               // [2][nRows] of the last ir_run
    Upw up;                                     // ir_set_upwind_mesh / ir_run_upwind
    int upVars;
    bool upMesh;
};

namespace {

// ------------------------------------------------------------------------------------------------ small device helpers

__device__ __forceinline__ double cross2(double ax, double ay, double bx, double by) { return ax * by - ay * bx; }

// point_in_half_plane as the reference executes it (its dummy arguments are (point, lineStart, lineEnd) while every
// caller passes (V1, V2, P)): cross product of (P - V2) and (V1 - V2), incremental_remap.F:9200-9235
__device__ __forceinline__ bool in_half_plane(double v1x, double v1y, double v2x, double v2y, double px, double py)
{
    return cross2(px - v2x, py - v2y, v1x - v2x, v1y - v2y) >= 0.0;
}

__device__ __forceinline__ double tri_area(double x1, double y1, double x2, double y2, double x3, double y3)
{
    return fabs(0.5 * ((x2 - x1) * (y3 - y1) - (y2 - y1) * (x3 - x1)));
}

// find_line_intersection, incremental_remap.F:8934
__device__ bool line_intersection(double x1, double y1, double x2, double y2, double x3, double y3, double x4, double y4,
                                  double &ipx, double &ipy)
{
    const double rx = x2 - x1, ry = y2 - y1, sx = x4 - x3, sy = y4 - y3;
    const double rsCross = rx * sy - ry * sx;
    const double rsCrossMin = EPS11 * sqrt((rx * rx + ry * ry) * (sx * sx + sy * sy));
    if (fabs(rsCross) > rsCrossMin) {
        const double t1 = (sy * (x3 - x1) - sx * (y3 - y1)) / rsCross;
        const double t2 = (ry * (x3 - x1) - rx * (y3 - y1)) / rsCross;
        ipx = x1 + t1 * rx;
        ipy = y1 + t1 * ry;
        return t1 > 0.0 && t1 < 1.0 && t2 > 0.0 && t2 < 1.0;
    }
    ipx = DBL_MAX;
    ipy = DBL_MAX;
    return false;
}

// ------------------------------------------------------------------------------------------------------- layout kernels

// host (n, w) row-major -> device [w][pitch]
template <typename T>
__global__ void k_rows_in(const T *__restrict__ raw, T *__restrict__ dst, size_t n, int w, size_t pitch)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int j = 0; j < w; j++) dst[(size_t)j * pitch + i] = raw[i * w + j];
}

// tracer: host (n, w) -> rows row0 .. row0+w-1 of val, and back
__global__ void k_tracer_out(double *__restrict__ raw, const double *__restrict__ val, size_t n, int w, size_t pitch)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int j = 0; j < w; j++) raw[i * w + j] = val[(size_t)j * pitch + i];
}

// per (slot, cell): signed dcEdge and the flux sign (compute_gradient :4330-4340, update_mass_and_tracers :7260-7270)
__global__ void k_cell_edge_signs(Dev d, const double *__restrict__ dcEdge)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (size_t)d.nC) return;
    for (int k = 0; k < d.M; k++) {
        const int e = d.edgesOnCell[k * d.nCp + c];
        double s = 1.0;
        int fs = 1;
        if (k < d.nEdgesOnCell[c] && e >= 1 && e <= d.nE) {
            const bool first = ((int)c + 1 == d.cellsOnEdge[2 * ((size_t)e - 1)]);   // cellsOnEdge(1, e), host layout
            fs = first ? 1 : -1;
            s = (first ? 1.0 : -1.0) * dcEdge[e - 1];
        }
        d.sdc[k * d.nCp + c] = s;
        d.fluxSign[k * d.nCp + c] = fs;
    }
}

// -------------------------------------------------------------------------------------------------------------- step

// maskCell (make_masks :3455-3470: sum over the mass field's categories and layers > 0) and volume -> thickness
// (volume_to_thickness :9248, every column including the extra one)
__global__ void k_prepare(Dev d, int massRow0, int massRows)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > (size_t)d.nC) return;
    if (c < (size_t)d.nC) {
        double massSumCell = 0.0;
        for (int r = 0; r < massRows; r++) massSumCell = massSumCell + d.val[(size_t)(massRow0 + r) * d.nCp + c];
        d.maskCell[c] = massSumCell > 0.0 ? 1 : 0;
    }
    for (int r = 0; r < d.nRows; r++) {
        const RowInfo &ri = d.rows[r];
        if (!ri.volumeLike) continue;
        const double area = d.val[(size_t)(massRow0 + ri.cat) * d.nCp + c];   // mass field with one layer: row = category
        double *v = &d.val[(size_t)r * d.nCp + c];
        *v = (area > 0.0) ? *v / area : 0.0;
    }
}

__device__ __forceinline__ bool row_mask(const Dev &d, const RowInfo &ri, size_t c)
{
    if (c >= (size_t)d.nC) return false;                                  // the extra slot never holds ice
    if (ri.depth == 0) return true;                                       // make_masks :3450
    return d.val[(size_t)ri.chain[ri.depth - 1] * d.nCp + c] > EPS11;     // parent value above the threshold, :3480-3510
}

// compute_barycenter_coordinates (:4658): centre of mass * tracer chain of `n` linear fields
__device__ void barycenter(const Dev &d, size_t c, int n, const double *mean, const double *cen, const double *gx,
                           const double *gy, double &xB, double &yB)
{
    const double *G = d.geom;
    const size_t p = d.nCp;
    const double ax = G[0 * p + c], ay = G[1 * p + c], axx = G[2 * p + c], axy = G[3 * p + c], ayy = G[4 * p + c];
    if (n == 1) {
        const double c0 = cen[0], cx = gx[0], cy = gy[0];
        const double reciprocal = (fabs(mean[0]) > 0.0) ? 1.0 / mean[0] : 0.0;
        xB = (c0 * ax + cx * axx + cy * axy) * reciprocal;
        yB = (c0 * ay + cx * axy + cy * ayy) * reciprocal;
        return;
    }
    const double axxx = G[5 * p + c], axxy = G[6 * p + c], axyy = G[7 * p + c], ayyy = G[8 * p + c];
    if (n == 2) {
        const double c0 = cen[0] * cen[1];
        const double cx = cen[0] * gx[1] + gx[0] * cen[1];
        const double cy = cen[0] * gy[1] + gy[0] * cen[1];
        const double cxx = gx[0] * gx[1];
        const double cxy = gx[0] * gy[1] + gy[0] * gx[1];
        const double cyy = gy[0] * gy[1];
        const double prod = mean[0] * mean[1];
        const double reciprocal = (fabs(prod) > 0.0) ? 1.0 / prod : 0.0;
        xB = (c0 * ax + cx * axx + cy * axy + cxx * axxx + cxy * axxy + cyy * axyy) * reciprocal;
        yB = (c0 * ay + cx * axy + cy * ayy + cxx * axxy + cxy * axyy + cyy * ayyy) * reciprocal;
        return;
    }
    const double axxxx = G[9 * p + c], axxxy = G[10 * p + c], axxyy = G[11 * p + c], axyyy = G[12 * p + c], ayyyy = G[13 * p + c];
    const double c0 = cen[0] * cen[1] * cen[2];
    const double cx = cen[0] * cen[1] * gx[2] + cen[0] * gx[1] * cen[2] + gx[0] * cen[1] * cen[2];
    const double cy = cen[0] * cen[1] * gy[2] + cen[0] * gy[1] * cen[2] + gy[0] * cen[1] * cen[2];
    const double cxx = cen[0] * gx[1] * gx[2] + gx[0] * cen[1] * gx[2] + gx[0] * gx[1] * cen[2];
    const double cxy = cen[0] * gx[1] * gy[2] + gx[0] * gy[1] * cen[2] + gy[0] * cen[1] * gx[2] + cen[0] * gy[1] * gx[2] +
                       gx[0] * cen[1] * gy[2] + gy[0] * gx[1] * cen[2];
    const double cyy = cen[0] * gy[1] * gy[2] + gy[0] * cen[1] * gy[2] + gy[0] * gy[1] * cen[2];
    const double cxxx = gx[0] * gx[1] * gx[2];
    const double cxxy = gx[0] * gx[1] * gy[2] + gx[0] * gy[1] * gx[2] + gy[0] * gx[1] * gx[2];
    const double cxyy = gy[0] * gy[1] * gx[2] + gy[0] * gx[1] * gy[2] + gx[0] * gy[1] * gy[2];
    const double cyyy = gy[0] * gy[1] * gy[2];
    const double prod = mean[0] * mean[1] * mean[2];
    const double reciprocal = (fabs(prod) > 0.0) ? 1.0 / prod : 0.0;
    xB = (c0 * ax + cx * axx + cy * axy + cxx * axxx + cxy * axxy + cyy * axyy + cxxx * axxxx + cxxy * axxxy + cxyy * axxyy +
          cyyy * axyyy) * reciprocal;
    yB = (c0 * ay + cx * axy + cy * ayy + cxx * axxy + cxy * axyy + cyy * ayyy + cxxx * axxxy + cxxy * axxyy + cxyy * axyyy +
          cyyy * ayyyy) * reciprocal;
}

// construct_linear_tracer_fields (:3580): compute_gradient (:4204), limit_tracer_gradient (:4802), centre value
// (:3735), barycentre (:3750-3840).  One block = 32 cells (lane = cell) x CW thread rows and ALL rows of all categories.  The cell's geometry (reconstruction coefficients, signed dcEdge, vertex coordinates,
// neighbour cells: 6 * maxEdges doubles) is staged in shared memory once per block and used by every row; the rows are
// processed depth by depth with a barrier in between, because a row needs the barycentre of its parent (and, for its
// own barycentre, the reconstructions of all its ancestors) in the same cell -- written by other threads of the block.
// A row needs neighbour cells' INPUT values only, so the whole hierarchy is one launch.
constexpr int CL = 32;   // cells per block
constexpr int CW = 12;   // thread rows per block

inline size_t ir_reconstruct_smem_bytes() { return sizeof(double) * ((size_t)CL * (6 * MAXM + 8)) + sizeof(int) * (size_t)CL * (MAXM + 2); }

__global__ void __launch_bounds__(CL *CW, 3) k_reconstruct_coop(Dev d, int maxDepth)
{
    IR_DYN_SHARED(double, sm);
    double *S = sm;                                  // [6 * MAXM][CL]: coef (3M), sdc (M), xvc (M), yvc (M)
    double *TR = S + 6 * MAXM * CL;                  // [6][CL] rows 1, 2 of transGlobalToCell
    double *AVG = TR + 6 * CL;                       // [2][CL] geometric centre
    int *NB = reinterpret_cast<int *>(AVG + 2 * CL); // [MAXM][CL] cellsOnCell (1-based)
    int *NE = NB + MAXM * CL;                        // [CL] nEdgesOnCell
    int *ICE = NE + CL;                              // [CL] maskCell
    const int lane = threadIdx.x, ty = threadIdx.y;
    const size_t c = (size_t)blockIdx.x * CL + lane;
    const bool valid = c < (size_t)d.nC;
    const size_t p = d.nCp;
    const int M = d.M;
    // ---- stage the geometry ----
    for (int row = ty; row < 6 * M; row += CW) {
        double v = 0.0;
        if (valid) {
            if (row < 3 * M) v = d.coef[(size_t)row * p + c];
            else if (row < 4 * M) v = d.sdc[(size_t)(row - 3 * M) * p + c];
            else if (row < 5 * M) v = d.xvc[(size_t)(row - 4 * M) * p + c];
            else v = d.yvc[(size_t)(row - 5 * M) * p + c];
        }
        S[row * CL + lane] = v;
    }
    for (int row = ty; row < M; row += CW) NB[row * CL + lane] = valid ? d.cellsOnCell[(size_t)row * p + c] : 0;
    for (int row = ty; row < 6; row += CW) TR[row * CL + lane] = (valid && d.sphere) ? d.trans[(size_t)row * p + c] : 0.0;
    for (int row = ty; row < 2; row += CW) AVG[row * CL + lane] = valid ? d.geom[(size_t)row * p + c] : 0.0;
    if (ty == 0) {
        NE[lane] = valid ? d.nEdgesOnCell[c] : 0;
        ICE[lane] = valid ? d.maskCell[c] : 0;
    }
    __syncthreads();
    const int n = NE[lane];
    const bool ice = ICE[lane] == 1;
    const double xAvg = AVG[lane], yAvg = AVG[CL + lane];
    int j0 = 0;
    for (int q = 0; q <= maxDepth; q++) {
        int j1 = j0;
        while (j1 < d.nRowsPerCat && d.catRows[j1].depth == q) j1++;
        const int cnt = j1 - j0;
        for (int it = ty; valid && it < d.nK * cnt; it += CW) {
            const int cat = it / cnt, j = j0 + (it - cat * cnt);
            const int r = d.catRows[j].baseRow + cat * d.catRows[j].layers;
            const RowInfo ri = d.rows[r];
            const double *field = d.val + (size_t)r * p;
            const double f0 = field[c];
            const int pr = ri.depth > 0 ? ri.chain[ri.depth - 1] : -1;
            double xB, yB;   // barycentre of the parent: where this row's value sits
            if (pr >= 0) { xB = d.xBary[(size_t)ri.parentSlot * p + c]; yB = d.yBary[(size_t)ri.parentSlot * p + c]; }
            else { xB = xAvg; yB = yAvg; }
            double xg = 0.0, yg = 0.0;
            if (ice) {
                const bool m0 = row_mask(d, ri, c);
                double g1 = 0.0, g2 = 0.0, g3 = 0.0;
                double maxNeighbor = f0, minNeighbor = f0;
                for (int k = 0; k < n; k++) {
                    const int nbk = NB[k * CL + lane];
                    const bool mn = row_mask(d, ri, (size_t)nbk - 1);
                    double normalGrad = 0.0;
                    const double fn = (nbk >= 1 && nbk <= d.nC + 1) ? field[nbk - 1] : 0.0;
                    if (nbk >= 1 && nbk <= d.nC && m0 && mn) normalGrad = (fn - f0) / S[(3 * M + k) * CL + lane];
                    g1 = g1 + S[(3 * k + 0) * CL + lane] * normalGrad;
                    g2 = g2 + S[(3 * k + 1) * CL + lane] * normalGrad;
                    g3 = g3 + S[(3 * k + 2) * CL + lane] * normalGrad;
                    if (mn) {
                        maxNeighbor = (maxNeighbor > fn) ? maxNeighbor : fn;
                        minNeighbor = (minNeighbor < fn) ? minNeighbor : fn;
                    }
                }
                if (d.rotate && d.sphere) { const double t = g1; g1 = -g3; g3 = t; }
                if (d.sphere) {
                    xg = TR[0 * CL + lane] * g1 + TR[1 * CL + lane] * g2 + TR[2 * CL + lane] * g3;
                    yg = TR[3 * CL + lane] * g1 + TR[4 * CL + lane] * g2 + TR[5 * CL + lane] * g3;
                } else {
                    xg = g1;
                    yg = g2;
                }
                maxNeighbor = maxNeighbor - f0;
                minNeighbor = minNeighbor - f0;
                double maxLocal = 0.0, minLocal = 0.0;
                for (int k = 0; k < n; k++) {
                    const double dev = xg * (S[(4 * M + k) * CL + lane] - xB) + yg * (S[(5 * M + k) * CL + lane] - yB);
                    maxLocal = (maxLocal > dev) ? maxLocal : dev;
                    minLocal = (minLocal < dev) ? minLocal : dev;
                }
                double f1 = 1.0, f2 = 1.0;
                if (fabs(maxLocal) > fabs(maxNeighbor)) { f1 = maxNeighbor / maxLocal; if (!(f1 > 0.0)) f1 = 0.0; }
                if (fabs(minLocal) > fabs(minNeighbor)) { f2 = minNeighbor / minLocal; if (!(f2 > 0.0)) f2 = 0.0; }
                double gradFactor = (f1 < f2) ? f1 : f2;
                gradFactor = gradFactor - EPS11;
                if (!(gradFactor > 0.0)) gradFactor = 0.0;
                xg = xg * gradFactor;
                yg = yg * gradFactor;
            }
            const double cen = f0 - xg * xB - yg * yB;
            d.xGrad[(size_t)r * p + c] = xg;
            d.yGrad[(size_t)r * p + c] = yg;
            d.center[(size_t)r * p + c] = cen;
            if (ri.hasChild) {
                double bx = 0.0, by = 0.0;
                if (ice) {
                    double mean[3], ce[3], gx[3], gy[3];
                    const int nn = ri.depth + 1;
                    for (int z = 0; z < nn - 1; z++) {
                        const size_t a = (size_t)ri.chain[z] * p + c;
                        mean[z] = d.val[a]; ce[z] = d.center[a]; gx[z] = d.xGrad[a]; gy[z] = d.yGrad[a];
                    }
                    mean[nn - 1] = f0; ce[nn - 1] = cen; gx[nn - 1] = xg; gy[nn - 1] = yg;
                    barycenter(d, c, nn, mean, ce, gx, gy, bx, by);
                }
                d.xBary[(size_t)ri.slot * p + c] = bx;
                d.yBary[(size_t)ri.slot * p + c] = by;
            }
        }
        j0 = j1;
        __syncthreads();       // the next depth reads what this one stored (same cell, other threads of the block)
    }
}

// find_departure_points + find_departure_triangles + get_triangle_quadrature_points for one edge per thread.
struct Tri {
    double x[3], y[3];
    double e1x, e1y, e2x, e2y;
    int cell, vOnEdge, vOnCell, sign;
};

__device__ int vertex_on_cell(const Dev &d, int iCell, int vGlobal, int prev)
{
    int r = prev;
    if (iCell < 1 || iCell > d.nC) return r;
    const int n = d.nEdgesOnCell[iCell - 1];
    for (int k = 0; k < n; k++)
        if (d.verticesOnCell[k * d.nCp + (iCell - 1)] == vGlobal) r = k + 1;
    return r;
}

__global__ void __launch_bounds__(128) k_triangles(Dev d, double dt)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)d.nE) return;
    const size_t pe = d.nEp, pc = d.nCp;
    for (int t = 0; t < NTRI; t++) {
        d.triArea[t * pe + e] = 0.0;
        d.iCellTri[t * pe + e] = 0;
    }
    const int v1 = d.verticesOnEdge[e], v2 = d.verticesOnEdge[pe + e];
    // find_departure_points (:5255): departurePoint = -velocity * dt
    double dpx[2], dpy[2];
    bool any = false;
    if (d.remapEdge[e] == 1) {
        const int vv[2] = {v1, v2};
        for (int k = 0; k < 2; k++) {
            dpx[k] = -d.u[vv[k] - 1] * dt;
            dpy[k] = -d.v[vv[k] - 1] * dt;
            if (dpx[k] * dpx[k] + dpy[k] * dpy[k] > 0.0) any = true;
        }
    }
    d.maskEdge[e] = any ? 1 : 0;
    if (!any) return;

    double XE[NVER], YE[NVER];
#pragma unroll
    for (int k = 0; k < NVER; k++) { XE[k] = d.xve[k * pe + e]; YE[k] = d.yve[k * pe + e]; }
    int EO[NEER], CO[NCER];
#pragma unroll
    for (int k = 0; k < NEER; k++) { EO[k] = d.eoer[k * pe + e]; CO[k] = d.coer[k * pe + e]; }
    const int vg[2] = {v1, v2};
    const double evx[2] = {XE[0], XE[1]}, evy[2] = {YE[0], YE[1]};
    for (int k = 0; k < 2; k++) { dpx[k] = XE[k] + dpx[k]; dpy[k] = YE[k] + dpy[k]; }

    Tri T[NTRI];
    int count = 0;
    const int D = d.D;
    const bool sphere = d.sphere != 0;
#define NEW_TRI(name)                                                   \
    if (count >= NTRI) { atomicOr(d.flags, FLAG_MANY_TRI); return; }    \
    Tri &name = T[count];                                               \
    count++;                                                            \
    name.e1x = name.e1y = name.e2x = name.e2y = 0.0;                    \
    name.vOnCell = 0

    // side triangles: does the segment D1-D2 cut one of the side edges E1..E4 (:5700-5900)?
    for (int iv = 0; iv < 2; iv++) {
        const double n0x = evx[iv], n0y = evy[iv];
        for (int side = 0; side < 2; side++) {
            const int ieo = (iv + 1) + 2 * side;   // E1/E3 at V1, E2/E4 at V2 (1-based)
            const int ivr = ieo + 2;               // V3..V6 (1-based)
            const int en = EO[ieo - 1];
            bool hit = false;
            double ipx = 0.0, ipy = 0.0, n1x = 0.0, n1y = 0.0;
            if (en >= 1 && en <= d.nE) {
                n1x = XE[ivr - 1];
                n1y = YE[ivr - 1];
                hit = line_intersection(dpx[0], dpy[0], dpx[1], dpy[1], n0x, n0y, n1x, n1y, ipx, ipy);
            }
            if (!hit) continue;
            NEW_TRI(t);
            t.x[0] = evx[iv]; t.y[0] = evy[iv];
            t.x[1] = dpx[iv]; t.y[1] = dpy[iv];
            t.x[2] = ipx;     t.y[2] = ipy;
            t.vOnEdge = iv + 1;
            t.cell = CO[iv + 2];
            t.vOnCell = vertex_on_cell(d, t.cell, vg[iv], 0);
            if (sphere) {
                t.e1x = n1x - n0x; t.e1y = n1y - n0y;
                int iOtherEdge;
                if (D == 3) { iOtherEdge = ieo + 2; if (iOtherEdge > 4) iOtherEdge = iOtherEdge - 4; }
                else iOtherEdge = iv + 1 + 4;
                const int iOtherVertex = iOtherEdge + 2;
                t.e2x = XE[iOtherVertex - 1] - n0x; t.e2y = YE[iOtherVertex - 1] - n0y;
            }
            t.sign = (side == 0) ? 1 : -1;
            if (D == 4) {
                const int en5 = EO[iv + 4];        // E5 at V1, E6 at V2
                bool hitMain = false;
                double ipmx = 0.0, ipmy = 0.0;
                if (en5 >= 1 && en5 <= d.nE)
                    hitMain = line_intersection(dpx[0], dpy[0], dpx[1], dpy[1], n0x, n0y, XE[iv + 6], YE[iv + 6], ipmx, ipmy);
                if (hitMain) {
                    t.x[2] = ipmx; t.y[2] = ipmy;
                    if (side == 0) {
                        t.cell = CO[iv + 4];
                        t.vOnCell = vertex_on_cell(d, t.cell, vg[iv], t.vOnCell);
                        if (sphere) { t.e1x = XE[ivr + 2 - 1] - n0x; t.e1y = YE[ivr + 2 - 1] - n0y; }
                    } else if (sphere) {
                        t.e1x = XE[ivr - 2 - 1] - n0x; t.e1y = YE[ivr - 2 - 1] - n0y;
                    }
                    const int prevOnEdge = t.vOnEdge, prevSign = t.sign;
                    const double pe2x = t.e2x, pe2y = t.e2y;
                    NEW_TRI(t2);
                    t2.x[0] = evx[iv]; t2.y[0] = evy[iv];
                    t2.x[1] = ipmx;    t2.y[1] = ipmy;
                    t2.x[2] = ipx;     t2.y[2] = ipy;
                    t2.cell = (side == 0) ? CO[iv + 2] : CO[iv + 4];
                    t2.vOnEdge = prevOnEdge;
                    t2.vOnCell = vertex_on_cell(d, t2.cell, vg[iv], 0);
                    if (sphere) { t2.e1x = XE[ivr - 1] - n0x; t2.e1y = YE[ivr - 1] - n0y; t2.e2x = pe2x; t2.e2y = pe2y; }
                    t2.sign = prevSign;
                } else if (side != 0) {
                    t.cell = CO[iv + 4];
                    t.vOnCell = vertex_on_cell(d, t.cell, vg[iv], t.vOnCell);
                }
            }
            dpx[iv] = ipx;   // the departure point moves to the side intersection for what follows (:5898)
            dpy[iv] = ipy;
        }
    }

    // central triangles in C1 / C2 (:5905-6075)
    {
        double ipmx, ipmy;
        const bool hitMain = line_intersection(dpx[0], dpy[0], dpx[1], dpy[1], evx[0], evy[0], evx[1], evy[1], ipmx, ipmy);
        bool two = hitMain;
        if (!hitMain) {
            const double quadArea = tri_area(evx[0], evy[0], evx[1], evy[1], dpx[1], dpy[1]) +
                                    tri_area(evx[0], evy[0], dpx[1], dpy[1], dpx[0], dpy[0]);
            two = quadArea > 0.0;
        }
        if (two) {
            for (int iv = 0; iv < 2; iv++) {
                NEW_TRI(t);
                if (hitMain) {
                    t.x[0] = evx[iv]; t.y[0] = evy[iv]; t.x[1] = dpx[iv]; t.y[1] = dpy[iv]; t.x[2] = ipmx; t.y[2] = ipmy;
                } else if (iv == 0) {
                    t.x[0] = evx[0]; t.y[0] = evy[0]; t.x[1] = evx[1]; t.y[1] = evy[1]; t.x[2] = dpx[0]; t.y[2] = dpy[0];
                } else {
                    t.x[0] = evx[1]; t.y[0] = evy[1]; t.x[1] = dpx[0]; t.y[1] = dpy[0]; t.x[2] = dpx[1]; t.y[2] = dpy[1];
                }
                t.vOnEdge = iv + 1;
                const bool inHP = in_half_plane(evx[0], evy[0], evx[1], evy[1], dpx[iv], dpy[iv]);
                t.cell = inHP ? CO[0] : CO[1];
                t.sign = inHP ? 1 : -1;
                t.vOnCell = vertex_on_cell(d, t.cell, vg[iv], 0);
                if (sphere) {
                    const int other = 1 - iv;
                    t.e1x = evx[other] - evx[iv]; t.e1y = evy[other] - evy[iv];
                    const int iOtherEdge = inHP ? (iv + 1) : (iv + 3);
                    t.e2x = XE[iOtherEdge + 2 - 1] - evx[iv]; t.e2y = YE[iOtherEdge + 2 - 1] - evy[iv];
                }
            }
        }
    }
#undef NEW_TRI

    // shift_vertices_of_departure_triangle (:6270), signed area, quadrature points (:6546)
    for (int it = 0; it < NTRI; it++) {
        double x[3] = {0.0, 0.0, 0.0}, y[3] = {0.0, 0.0, 0.0};
        if (it < count) {
            Tri &t = T[it];
            for (int q = 0; q < 3; q++) { x[q] = t.x[q]; y[q] = t.y[q]; }
            d.iCellTri[it * pe + e] = t.cell;
            if (t.cell >= 1 && t.cell <= d.nC) {
                const size_t ic = (size_t)t.cell - 1;
                const int kk = t.vOnCell < 1 ? 1 : t.vOnCell;
                const double xv = d.xvc[(kk - 1) * pc + ic], yv = d.yvc[(kk - 1) * pc + ic];
                const double oex = XE[t.vOnEdge - 1], oey = YE[t.vOnEdge - 1];
                if (sphere) {
                    double e1x = t.e1x, e1y = t.e1y, e2x = t.e2x, e2y = t.e2y;
                    const double crossProduct = cross2(e1x, e1y, e2x, e2y);
                    if (fabs(crossProduct) < EPS11) atomicOr(d.flags, FLAG_PARALLEL);
                    if (crossProduct > 0.0) {
                        const double tx = e1x, ty = e1y;
                        e1x = e2x; e1y = e2y; e2x = tx; e2y = ty;
                    }
                    const int n = d.nEdgesOnCell[ic];
                    int km1 = kk - 1; if (km1 < 1) km1 = km1 + n;
                    int kp1 = kk + 1; if (kp1 > n) kp1 = kp1 - n;
                    const double c1x = d.xvc[(km1 - 1) * pc + ic] - xv, c1y = d.yvc[(km1 - 1) * pc + ic] - yv;
                    const double c2x = d.xvc[(kp1 - 1) * pc + ic] - xv, c2y = d.yvc[(kp1 - 1) * pc + ic] - yv;
                    const double denom = e1x * e2y - e2x * e1y;
                    for (int q = 0; q < 3; q++) {
                        x[q] = x[q] - oex;
                        y[q] = y[q] - oey;
                        const double coeff_a = (x[q] * e2y - y[q] * e2x) / denom;
                        const double coeff_b = (y[q] * e1x - x[q] * e1y) / denom;
                        x[q] = xv + coeff_a * c1x + coeff_b * c2x;
                        y[q] = yv + coeff_a * c1y + coeff_b * c2y;
                    }
                } else {
                    for (int q = 0; q < 3; q++) {
                        x[q] = x[q] - oex + xv;
                        y[q] = y[q] - oey + yv;
                    }
                }
                const double area = fabs(0.5 * ((x[1] - x[0]) * (y[2] - y[0]) - (y[1] - y[0]) * (x[2] - x[0])));
                d.triArea[it * pe + e] = area * t.sign;
            }
        }
        double *xo = d.xq + (size_t)it * 6 * pe + e, *yo = d.yq + (size_t)it * 6 * pe + e;
        if (d.nQP == 3) {
            const double xMid = (x[0] + x[1] + x[2]) / 3.0;
            const double yMid = (x[0] + x[1] + x[2]) / 3.0;   // as the reference has it (:6598)
            for (int q = 0; q < 3; q++) {
                xo[q * pe] = 0.5 * (x[q] + xMid);
                yo[q * pe] = 0.5 * (y[q] + yMid);
            }
        } else {
            xo[0 * pe] = Q1QP * x[0] + Q1QP * x[1] + Q2QP * x[2]; yo[0 * pe] = Q1QP * y[0] + Q1QP * y[1] + Q2QP * y[2];
            xo[1 * pe] = Q1QP * x[0] + Q2QP * x[1] + Q1QP * x[2]; yo[1 * pe] = Q1QP * y[0] + Q2QP * y[1] + Q1QP * y[2];
            xo[2 * pe] = Q2QP * x[0] + Q1QP * x[1] + Q1QP * x[2]; yo[2 * pe] = Q2QP * y[0] + Q1QP * y[1] + Q1QP * y[2];
            xo[3 * pe] = Q3QP * x[0] + Q4QP * x[1] + Q4QP * x[2]; yo[3 * pe] = Q3QP * y[0] + Q4QP * y[1] + Q4QP * y[2];
            xo[4 * pe] = Q4QP * x[0] + Q3QP * x[1] + Q4QP * x[2]; yo[4 * pe] = Q4QP * y[0] + Q3QP * y[1] + Q4QP * y[2];
            xo[5 * pe] = Q4QP * x[0] + Q4QP * x[1] + Q3QP * x[2]; yo[5 * pe] = Q4QP * y[0] + Q4QP * y[1] + Q3QP * y[2];
        }
    }
}

// integrate_fluxes_over_triangles (:6667): one block = 32 edges (lane = edge) x FR thread rows, looping over
// the categories (the departure triangles and their quadrature points are the same for all of them).  The reference evaluates, for every row, at every quadrature point of every
// departure triangle, the product down the row's chain of parents of the linear reconstructions,
//     value = ((1 * v_0) * v_1) * ... * v_depth,   v_s = center_s + xGrad_s * x + yGrad_s * y  at the source cell.
// All rows of a category share v_0 (the mass field), and the children of a depth-1 row share v_0 * v_1, so the
// block computes those once into shared memory (1 * v_0 is v_0 exactly; (v_0 * v_1) is the reference's own
// intermediate: the results stay bit-identical) and a row of depth 2 costs one reconstruction and one product per
// point instead of three and three.  Phases (barriers between them):
//   A    area, source cell and ice mask of the edge's triangles -> shared memory; a block without any edge whose
//        triangles touch ice writes its zeros and leaves; every thread compacts its edge's non-empty triangles for
//        itself (triangle order is kept, it is the order the reference sums in);
//   B+C  thread (lane, triangle, point): the quadrature point, v_0 for the category's depth-0 rows (negative-mass
//        check, :6895) and v_0 * v_1 for its depth-1 rows -> shared memory;
//   D    thread (lane, row): the row's flux through the edge -> edgeFlux[row][edge]  (one warp = 32 consecutive
//        edges of one row: coalesced stores; shared memory is [..][lane]: conflict-free).
// One barrier for A, two per category; 12 thread rows, three blocks per SM (58 KB of shared memory each).
constexpr int FL = 32;   // edges per block
constexpr int FR = 12;   // thread rows per block
constexpr int FNS = NTRI * 6;   // (triangle, point) slots per edge

inline size_t ir_flux_smem_bytes(int nD0, int nD1)
{
    return sizeof(double) * ((size_t)FL * (2 * FNS + (size_t)(nD0 + nD1) * FNS + NTRI) + (size_t)FL * NTRI + 8);
}

template <int NQ>
__global__ void __launch_bounds__(FL *FR, 3) k_fluxes_coop(Dev d)
{
    IR_DYN_SHARED(double, sm);
    const int lane = threadIdx.x, ty = threadIdx.y;
    const size_t e = (size_t)blockIdx.x * FL + lane;
    const bool inRange = e < (size_t)d.nE;
    const size_t pe = d.nEp, pc = d.nCp;
    constexpr int SL = FNS * FL;                       // doubles per (row, all slots, all lanes) plane
    double *xq = sm;                                   // [FNS][FL]
    double *yq = xq + SL;                              // [FNS][FL]
    double *vp = yq + SL;                              // [(nD0 + nD1)][FNS][FL]: v_0 of the depth-0 rows, then v_0 * v_1
    double *areaRaw = vp + (d.nD0 + d.nD1) * SL;       // [NTRI][FL] signed area of the edge's triangles (0 = empty)
    int *cellRaw = reinterpret_cast<int *>(areaRaw + NTRI * FL);   // [NTRI][FL] source cell (0-based)
    int *iceRaw = cellRaw + NTRI * FL;                 // [NTRI][FL] its ice mask

    // ---- A: one thread per (edge, triangle) fetches area, source cell and its ice mask ----
    bool mine = false;
    if (ty < NTRI) {
        double ar = 0.0;
        int cl = 0, ic = 0;
        if (inRange && d.maskEdge[e] == 1) {
            ar = d.triArea[ty * pe + e];
            if (ar != 0.0) {
                cl = d.iCellTri[ty * pe + e] - 1;
                ic = d.maskCell[cl];
            }
        }
        areaRaw[ty * FL + lane] = ar; cellRaw[ty * FL + lane] = cl; iceRaw[ty * FL + lane] = ic;
        mine = ar != 0.0 && ic == 1;
    }
    // an edge has work when one of its non-empty triangles lies in a cell with ice: ice-free or motionless blocks
    // (most of an ocean mesh) write their zeros and leave
    if (!__syncthreads_or(mine)) {
        if (inRange)
            for (int it = ty; it < d.nK * d.nRowsPerCat; it += FR) {
                const int cat = it / d.nRowsPerCat, j = it - cat * d.nRowsPerCat;
                d.edgeFlux[(size_t)(d.catRows[j].baseRow + cat * d.catRows[j].layers) * pe + e] = 0.0;
            }
        return;
    }
    // every thread compacts the non-empty triangles of its edge for itself (triangle order kept: it is the order the
    // reference sums in): tmap holds the triangle index of compacted slot k in bits 3k .. 3k+2
    int nt = 0;
    unsigned tmap = 0;
    {
        bool ice = false;
#pragma unroll
        for (int t = 0; t < NTRI; t++) {
            if (areaRaw[t * FL + lane] == 0.0) continue;
            tmap |= (unsigned)t << (3 * nt);
            ice = ice || iceRaw[t * FL + lane] == 1;
            nt++;
        }
        // source cells without ice in any category: the mass reconstruction is identically zero there, every product
        // down the chain is a zero and the flux is the +0.0 written in phase D
        if (!ice) nt = 0;
    }
    // the quadrature points: fetched once, used by every category
    for (int s0 = ty; s0 < NTRI * NQ; s0 += FR) {
        const int k = s0 / NQ, q = s0 - NQ * k;
        if (k < nt) {
            const int t = (tmap >> (3 * k)) & 7, so = (k * 6 + q) * FL + lane;
            xq[so] = d.xq[(size_t)(t * 6 + q) * pe + e];
            yq[so] = d.yq[(size_t)(t * 6 + q) * pe + e];
        }
    }
    // (each thread reads back below exactly the points it stored: no barrier needed before B + C)
    for (int cat = 0; cat < d.nK; cat++) {
        // ---- B + C: v_0 of the depth-0 rows and v_0 * v_1 of the depth-1 rows, one (triangle, point) per thread ----
        for (int s0 = ty; s0 < NTRI * NQ; s0 += FR) {
            const int k = s0 / NQ, q = s0 - NQ * k;
            if (k < nt) {
                const int t = (tmap >> (3 * k)) & 7, so = (k * 6 + q) * FL + lane;
                const double x = xq[so], y = yq[so];
                const size_t cl = (size_t)cellRaw[t * FL + lane];
                for (int j = 0; j < d.nD0; j++) {          // depth-0 rows come first in the table
                    const size_t a = (size_t)(d.catRows[j].baseRow + cat * d.catRows[j].layers) * pc + cl;
                    const double value = 1.0 * (d.center[a] + d.xGrad[a] * x + d.yGrad[a] * y);
                    if (value < 0.0) atomicOr(d.flags, FLAG_NEG_QP);        // an abort condition (:6895): rare
                    vp[j * SL + so] = value;
                }
                for (int j1 = 0; j1 < d.nD1; j1++) {       // depth-1 rows follow
                    const CatRow cr = d.catRows[d.nD0 + j1];
                    const size_t a = (size_t)(cr.baseRow + cat * cr.layers) * pc + cl;
                    vp[(d.nD0 + j1) * SL + so] = vp[d.catRows[cr.anc[0]].cls * SL + so] * (d.center[a] + d.xGrad[a] * x + d.yGrad[a] * y);
                }
            }
        }
        __syncthreads();
        // ---- D: the fluxes, one (edge, row) per thread ----
        for (int j = ty; j < d.nRowsPerCat; j += FR) {
            const CatRow cr = d.catRows[j];
            const int r = cr.baseRow + cat * cr.layers;
            double flux = 0.0;
            if (nt > 0) {
                const int base = cr.vpIdx * SL;
                const size_t row2 = (size_t)(cr.base2 + cat * cr.layers2) * pc, row3 = (size_t)r * pc;
                for (int k = 0; k < nt; k++) {
                    const int t = (tmap >> (3 * k)) & 7;
                    const size_t cl = (size_t)cellRaw[t * FL + lane];
                    double c2 = 0.0, gx2 = 0.0, gy2 = 0.0, c3 = 0.0, gx3 = 0.0, gy3 = 0.0;
                    if (cr.depth >= 2) { c2 = d.center[row2 + cl]; gx2 = d.xGrad[row2 + cl]; gy2 = d.yGrad[row2 + cl]; }
                    if (cr.depth >= 3) { c3 = d.center[row3 + cl]; gx3 = d.xGrad[row3 + cl]; gy3 = d.yGrad[row3 + cl]; }
                    double tracerIntegral = 0.0;
#pragma unroll
                    for (int q = 0; q < NQ; q++) {
                        const int so = (k * 6 + q) * FL + lane;
                        double value = vp[base + so];
                        if (cr.depth >= 2) {
                            const double x = xq[so], y = yq[so];
                            value = value * (c2 + gx2 * x + gy2 * y);
                            if (cr.depth >= 3) value = value * (c3 + gx3 * x + gy3 * y);
                        }
                        const double w = (NQ == 3) ? (1.0 / 3.0) : (q < 3 ? W1QP : W2QP);
                        tracerIntegral = tracerIntegral + w * value;
                    }
                    flux = flux + areaRaw[t * FL + lane] * tracerIntegral;
                }
            }
            if (inRange) d.edgeFlux[(size_t)r * pe + e] = flux;
        }
        __syncthreads();       // the next category overwrites the shared partial products
    }
}

// compute_mass_tracer_products (:6982), update_mass_and_tracers (:7125), zap_small_mass (:8764; one-layer mass field)
// and thickness -> volume (:9295): one block = 32 cells x UW thread rows and all rows of all categories, depth by depth
// (a row needs the NEW mass * tracer product of its parent in the same cell), the cell's edges and flux signs staged in
// shared memory.
constexpr int UW = 12;

__global__ void __launch_bounds__(CL *UW, 4) k_update_coop(Dev d, int massOneLayer, int maxDepth)
{
    __shared__ int EDGE[MAXM][CL], SIGN[MAXM][CL], NE[CL];
    __shared__ double AREA[CL];
    const int lane = threadIdx.x, ty = threadIdx.y;
    const size_t c = (size_t)blockIdx.x * CL + lane;
    const bool valid = c <= (size_t)d.nC;            // the extra slot takes part (it keeps its values)
    const bool owned = c < (size_t)d.nCS;
    const size_t pe = d.nEp, pc = d.nCp;
    for (int k = ty; k < d.M; k += UW) {
        EDGE[k][lane] = owned ? d.edgesOnCell[(size_t)k * pc + c] : 1;
        SIGN[k][lane] = owned ? d.fluxSign[(size_t)k * pc + c] : 0;
    }
    if (ty == 0) {
        NE[lane] = owned ? d.nEdgesOnCell[c] : 0;
        AREA[lane] = owned ? d.areaCell[c] : 1.0;
    }
    __syncthreads();
    const int n = NE[lane];
    const double area = AREA[lane];
    int j0 = 0;
    for (int q = 0; q <= maxDepth; q++) {
        int j1 = j0;
        while (j1 < d.nRowsPerCat && d.catRows[j1].depth == q) j1++;
        const int cnt = j1 - j0;
        for (int it = ty; valid && it < d.nK * cnt; it += UW) {
            const int cat = it / cnt, j = j0 + (it - cat * cnt);
            const int r = d.catRows[j].baseRow + cat * d.catRows[j].layers;
            if (!owned) {                               // halo cells and the extra slot keep their values
                d.valNew[(size_t)r * pc + c] = d.val[(size_t)r * pc + c];
                continue;
            }
            const RowInfo ri = d.rows[r];
            double fluxFromCell = 0.0;
            for (int k = 0; k < n; k++)
                fluxFromCell = fluxFromCell + d.edgeFlux[(size_t)r * pe + (EDGE[k][lane] - 1)] * (double)SIGN[k][lane];
            double mtpOld = 1.0;
            for (int z = 0; z <= ri.depth; z++) mtpOld = mtpOld * d.val[(size_t)ri.chain[z] * pc + c];
            const double pm = ri.depth > 0 ? d.mtpNew[(size_t)ri.parentSlot * pc + c] : 1.0;
            double v = 0.0;
            if (pm > 0.0) v = (mtpOld - (fluxFromCell / area)) / pm;
            if (ri.slot >= 0) d.mtpNew[(size_t)ri.slot * pc + c] = pm * v;   // only children read it
            if (ri.depth == 0) {
                constexpr double puny2 = 1.0e-11 * 1.0e-11;   // seaicePuny**2
                if (v < -puny2) atomicOr(d.flags, FLAG_NEG_MASS);
                else if (v < 0.0) v = 0.0;
            }
            d.valNew[(size_t)r * pc + c] = v;
        }
        j0 = j1;
        __syncthreads();
    }
    if (!massOneLayer) return;
    // zap_small_mass (:8764) and thickness -> volume (:9295); the mass row of category k is row k.  A row is rewritten
    // only if it changes: the volume-like rows (value -> mass * value) and, in a zapped category, every row (-> 0).
    // Every thread reads the category's new mass before any row (the mass row included) is rewritten.
    const int total = d.nK * d.nRowsPerCat;
    const int myItems = (total - ty + UW - 1) / UW;          // items ty, ty + UW, ...
    const int maxItems = (total + UW - 1) / UW;              // the same for every thread: the loop holds barriers
    double vNew[16];
    unsigned changed = 0;
    for (int base = 0; base < maxItems; base += 16) {
        const int m = myItems - base < 16 ? (myItems - base > 0 ? myItems - base : 0) : 16;
        changed = 0;
        for (int i = 0; valid && i < m; i++) {
            const int it = ty + (base + i) * UW, cat = it / d.nRowsPerCat, j = it - cat * d.nRowsPerCat;
            const double mass = d.valNew[(size_t)cat * pc + c];
            const bool zap = owned && mass > 0.0 && mass < 1.0e-22;
            const bool vol = d.catRows[j].volumeLike != 0;
            if (!zap && !vol) continue;
            double v = zap ? 0.0 : d.valNew[(size_t)(d.catRows[j].baseRow + cat * d.catRows[j].layers) * pc + c];
            if (vol) v = (zap ? 0.0 : mass) * v;
            vNew[i] = v;
            changed |= 1u << i;
        }
        __syncthreads();
        for (int i = 0; valid && i < m; i++) {
            if (!(changed & (1u << i))) continue;
            const int it = ty + (base + i) * UW, cat = it / d.nRowsPerCat, j = it - cat * d.nRowsPerCat;
            d.valNew[(size_t)(d.catRows[j].baseRow + cat * d.catRows[j].layers) * pc + c] = vNew[i];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------- optional checks (default off)
// config_conservation_check / config_monotonicity_check (incremental_remap.F:2574-2600, :2999-3015, :3259-3310).  With
// a check switched on, ir_run launches the kernels below around the update and runs zap / thickness -> volume as
// kernels of their own (the fused tail of k_update_coop is skipped), because the reference sums the new products before
// zap_small_mass and tests monotonicity after it but before the volumes are restored.  None of this is on the default path.

// tracer_local_min_max (:8268) on the old values, masks of make_masks(threshold 0) (:3001): one thread per (cell, row).
// Evaluated for every cell of the block; the reference computes the owned cells and fills the halo by an exchange.
__global__ void __launch_bounds__(128) k_check_minmax(Dev d)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= (size_t)d.nC) return;
    const RowInfo ri = d.rows[r];
    const double *val = d.val + (size_t)r * d.nCp;
    const double *par = ri.depth > 0 ? d.val + (size_t)ri.chain[ri.depth - 1] * d.nCp : nullptr;
    double lo = 0.0, hi = 0.0;
    if (par == nullptr || par[c] > 0.0) { lo = val[c]; hi = val[c]; }
    const int n = d.nEdgesOnCell[c];
    for (int k = 0; k < n; k++) {
        const int nb = d.cellsOnCell[(size_t)k * d.nCp + c];
        if (nb >= 1 && nb <= d.nC) {
            const size_t cn = (size_t)nb - 1;
            if (par == nullptr || par[cn] > 0.0) {
                if (val[cn] < lo) lo = val[cn];
                if (val[cn] > hi) hi = val[cn];
            }
        }
    }
    d.lmin[(size_t)r * d.nCp + c] = lo;
    d.lmax[(size_t)r * d.nCp + c] = hi;
}

// sum_tracers (:7998): area-weighted sum of mass * tracer products over the owned cells, one block per row.  The
// reference adds cell after cell (and then over the ranks); here every thread adds its cells in increasing order and the
// 256 partial sums are combined in a fixed tree, so the result is reproducible but not the serial sum's last bits.
__global__ void __launch_bounds__(256) k_check_sums(Dev d, const double *__restrict__ src, double *__restrict__ out)
{
    __shared__ double part[256];
    const int r = blockIdx.x, tid = threadIdx.x;
    const RowInfo ri = d.rows[r];
    double s = 0.0;
    for (size_t c = tid; c < (size_t)d.nCS; c += 256) {
        double mtp = 1.0;
        for (int z = 0; z <= ri.depth; z++) mtp = mtp * src[(size_t)ri.chain[z] * d.nCp + c];
        s = s + d.areaCell[c] * mtp;
    }
    part[tid] = s;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (tid < w) part[tid] = part[tid] + part[tid + w];
        __syncthreads();
    }
    if (tid == 0) out[r] = part[0];
}

// zap_small_mass (:8764) as a kernel of its own: one thread per owned cell (one-layer mass field: row = category)
__global__ void __launch_bounds__(128) k_zap(Dev d)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (size_t)d.nCS) return;
    for (int cat = 0; cat < d.nK; cat++) {
        const double mass = d.valNew[(size_t)cat * d.nCp + c];
        if (mass > 0.0 && mass < 1.0e-22)
            for (int j = 0; j < d.nRowsPerCat; j++) d.valNew[(size_t)(d.catRows[j].baseRow + cat * d.catRows[j].layers) * d.nCp + c] = 0.0;
    }
}

// thickness -> volume (:9295) as a kernel of its own: every column including the extra one
__global__ void __launch_bounds__(128) k_thickness_to_volume(Dev d)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > (size_t)d.nC) return;
    for (int r = 0; r < d.nRows; r++) {
        if (!d.rows[r].volumeLike) continue;
        double *v = &d.valNew[(size_t)r * d.nCp + c];
        *v = d.valNew[(size_t)d.rows[r].cat * d.nCp + c] * *v;
    }
}

// check_tracer_monotonicity (:8416), extendedMinMax = .true.: the bounds of (row, cell) widened by those of its edge
// neighbours, masks of make_masks(threshold eps11) on the OLD values.  The neighbours are read in their un-extended state
// (the reference widens in place cell after cell, so there a cell may also see what a lower-numbered neighbour has already
// gathered: bounds that depend on the numbering and are never tighter than these).
__device__ bool mono_bounds(const Dev &d, int r, size_t c, double &lo, double &hi)
{
    const RowInfo ri = d.rows[r];
    if (!row_mask(d, ri, c)) return false;
    lo = d.lmin[(size_t)r * d.nCp + c];
    hi = d.lmax[(size_t)r * d.nCp + c];
    const int n = d.nEdgesOnCell[c];
    for (int k = 0; k < n; k++) {
        const int nb = d.cellsOnCell[(size_t)k * d.nCp + c];
        if (nb >= 1 && nb <= d.nC && row_mask(d, ri, (size_t)nb - 1)) {
            const double nlo = d.lmin[(size_t)r * d.nCp + nb - 1], nhi = d.lmax[(size_t)r * d.nCp + nb - 1];
            if (nlo < lo) lo = nlo;
            if (nhi > hi) hi = nhi;
        }
    }
    return true;
}

__global__ void __launch_bounds__(128) k_check_mono(Dev d)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= (size_t)d.nCS || d.rows[r].depth == 0) return;   // monotonicity holds for tracers, not for the mass-like field
    double lo, hi;
    if (!mono_bounds(d, r, c, lo, hi)) return;
    const double v = d.valNew[(size_t)r * d.nCp + c];
    const double toleranceMin = EPS11 * fmax(1.0, fabs(lo)), toleranceMax = EPS11 * fmax(1.0, fabs(hi));
    if (v < lo - toleranceMin || v > hi + toleranceMax) {
        const unsigned long long key = ((unsigned long long)d.rowT[r] * ((unsigned long long)d.nC + 1) + c) *
                                           (unsigned long long)(d.nK * d.maxLayers) + (unsigned long long)d.rowKL[r];
        atomicMin(d.monoKey, key);
    }
}

// the numbers of the first violation, taken before thickness -> volume rewrites the volume-like rows (one thread)
__global__ void k_check_mono_detail(Dev d)
{
    const unsigned long long key = *d.monoKey;
    if (key == ~0ull) return;
    const unsigned long long KL = (unsigned long long)(d.nK * d.maxLayers);
    const int kl = (int)(key % KL);
    const unsigned long long tc = key / KL;
    const size_t c = (size_t)(tc % ((unsigned long long)d.nC + 1));
    const int t = (int)(tc / ((unsigned long long)d.nC + 1));
    for (int r = 0; r < d.nRows; r++) {
        if (d.rowT[r] != t || d.rowKL[r] != kl) continue;
        double lo = 0.0, hi = 0.0;
        mono_bounds(d, r, c, lo, hi);
        d.monoDetail[0] = d.valNew[(size_t)r * d.nCp + c];
        d.monoDetail[1] = lo;
        d.monoDetail[2] = hi;
    }
}

// rows 1 and 2 of transGlobalToCell (3,3,nCells): trans(i,j,c) at c*9 + j*3 + i
__global__ void k_trans_in(const double *__restrict__ raw, double *__restrict__ dst, size_t n, size_t pitch)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 3; j++) dst[(size_t)(i * 3 + j) * pitch + c] = raw[c * 9 + j * 3 + i];
}

// ------------------------------------------------------------------------------------------------ init-time geometry
// seaice_init_advection_incremental_remap's geometry (:446-711) for hosts that do not run the Fortran init: local
// frames, vertex coordinates in cell and edge frames, remap stencils, geometric cell averages.  Init-time work, so the
// arrays keep the host layout on the device.

struct Geo {
    int nC, nCS, nV, nE, M, D, sphere, rotate;
    const int *nEdgesOnCell, *edgesOnCell, *verticesOnCell, *cellsOnEdge, *verticesOnEdge, *edgesOnVertex;
    const double *xC, *yC, *zC, *xV, *yV, *zV, *xE, *yE, *zE, *dcEdge, *dvEdge;
    double *trans, *xvc, *yvc, *xve, *yve, *minLen, *geom[14];
    int *remapEdge, *coer, *eoer, *flags;
};
enum { GEO_BAD_EDGE = 1, GEO_BAD_CELL = 2 };

// rotate_global_vectors (:948): (x, y, z) -> (-z, y, x) when the Cartesian grid is rotated
__device__ __forceinline__ void point(const Geo &g, const double *x, const double *y, const double *z, int i, double *o)
{
    if (g.rotate && g.sphere) { o[0] = -z[i]; o[1] = y[i]; o[2] = x[i]; }
    else { o[0] = x[i]; o[1] = y[i]; o[2] = z[i]; }
}

// define_local_to_global_transformations (:990): rows east, north, up; t(i,j) at t[(j-1)*3 + (i-1)]
__device__ void global_to_local(const double *pt, double *t)
{
    double e1[3], e2[3], e3[3] = {pt[0], pt[1], pt[2]};
    double mag = sqrt(e3[0] * e3[0] + e3[1] * e3[1] + e3[2] * e3[2]);
    if (mag > 0.0) { e3[0] = e3[0] / mag; e3[1] = e3[1] / mag; e3[2] = e3[2] / mag; } else { e3[0] = e3[1] = e3[2] = 0.0; }
    if (fabs(pt[0] * pt[0] + pt[1] * pt[1]) > EPS11) {
        e1[0] = -pt[1]; e1[1] = pt[0]; e1[2] = 0.0;
        mag = sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
        if (mag > 0.0) { e1[0] = e1[0] / mag; e1[1] = e1[1] / mag; e1[2] = e1[2] / mag; } else { e1[0] = e1[1] = e1[2] = 0.0; }
        e2[0] = e3[1] * e1[2] - e3[2] * e1[1];
        e2[1] = e3[2] * e1[0] - e3[0] * e1[2];
        e2[2] = e3[0] * e1[1] - e3[1] * e1[0];
    } else if (pt[2] > 0.0) {
        e1[0] = 1.0; e1[1] = 0.0; e1[2] = 0.0; e2[0] = 0.0; e2[1] = 1.0; e2[2] = 0.0;
    } else {
        e1[0] = 0.0; e1[1] = 1.0; e1[2] = 0.0; e2[0] = 1.0; e2[1] = 0.0; e2[2] = 0.0;
    }
    for (int j = 0; j < 3; j++) { t[j * 3 + 0] = e1[j]; t[j * 3 + 1] = e2[j]; t[j * 3 + 2] = e3[j]; }
}

// get_vertex_on_cell_coordinates (:1823) + compute_geometric_cell_averages (:2097)
__global__ void k_geo_cells(Geo g)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.nC) return;
    const int M = g.M, n = g.nEdgesOnCell[c];
    double pc[3], t[9];
    point(g, g.xC, g.yC, g.zC, c, pc);
    if (g.sphere) {
        global_to_local(pc, t);
        for (int q = 0; q < 9; q++) g.trans[(size_t)c * 9 + q] = t[q];
    }
    double xv[MAXM], yv[MAXM];
    for (int k = 0; k < n; k++) {
        const int v = g.verticesOnCell[(size_t)c * M + k] - 1;
        double pv[3];
        point(g, g.xV, g.yV, g.zV, v, pv);
        if (g.sphere) {
            const double w0 = pv[0] - pc[0], w1 = pv[1] - pc[1], w2 = pv[2] - pc[2];
            xv[k] = t[0] * w0 + t[3] * w1 + t[6] * w2;
            yv[k] = t[1] * w0 + t[4] * w1 + t[7] * w2;
        } else {
            xv[k] = pv[0] - pc[0];
            yv[k] = pv[1] - pc[1];
        }
        g.xvc[(size_t)c * M + k] = xv[k];
        g.yvc[(size_t)c * M + k] = yv[k];
    }
    for (int k = 0; k < n; k++) {
        const int k1 = (k + 1 >= n) ? 0 : k + 1;
        if (!(cross2(xv[k], yv[k], xv[k1], yv[k1]) >= 0.0)) atomicOr(g.flags, GEO_BAD_CELL);
    }
    double frac[MAXM], sumArea = 0.0;
    for (int k = 0; k < n; k++) {
        const int e = g.edgesOnCell[(size_t)c * M + k] - 1;
        frac[k] = 0.25 * g.dcEdge[e] * g.dvEdge[e];
        sumArea = sumArea + frac[k];
    }
    for (int k = 0; k < n; k++) frac[k] = frac[k] / sumArea;
    double acc[14];
    for (int q = 0; q < 14; q++) acc[q] = 0.0;
    for (int k = 0; k < n; k++) {
        const int k1 = (k + 1 >= n) ? 0 : k + 1;
        const double x1 = 0.0, y1 = 0.0, x2 = xv[k], y2 = yv[k], x3 = xv[k1], y3 = yv[k1];
        double xq[6], yq[6];
        xq[0] = Q1QP * x1 + Q1QP * x2 + Q2QP * x3; yq[0] = Q1QP * y1 + Q1QP * y2 + Q2QP * y3;
        xq[1] = Q1QP * x1 + Q2QP * x2 + Q1QP * x3; yq[1] = Q1QP * y1 + Q2QP * y2 + Q1QP * y3;
        xq[2] = Q2QP * x1 + Q1QP * x2 + Q1QP * x3; yq[2] = Q2QP * y1 + Q1QP * y2 + Q1QP * y3;
        xq[3] = Q3QP * x1 + Q4QP * x2 + Q4QP * x3; yq[3] = Q3QP * y1 + Q4QP * y2 + Q4QP * y3;
        xq[4] = Q4QP * x1 + Q3QP * x2 + Q4QP * x3; yq[4] = Q4QP * y1 + Q3QP * y2 + Q4QP * y3;
        xq[5] = Q4QP * x1 + Q4QP * x2 + Q3QP * x3; yq[5] = Q4QP * y1 + Q4QP * y2 + Q3QP * y3;
        double a[14];
        for (int q = 0; q < 14; q++) a[q] = 0.0;
        for (int q = 0; q < 6; q++) {
            const double x = xq[q], y = yq[q], w = (q < 3) ? W1QP : W2QP;
            const double x2p = x * x, y2p = y * y, x3p = x * x * x, y3p = y * y * y, x4p = (x * x) * (x * x), y4p = (y * y) * (y * y);
            a[0] = a[0] + w * x;        a[1] = a[1] + w * y;
            a[2] = a[2] + w * x2p;      a[3] = a[3] + w * x * y;      a[4] = a[4] + w * y2p;
            a[5] = a[5] + w * x3p;      a[6] = a[6] + w * x2p * y;    a[7] = a[7] + w * x * y2p;   a[8] = a[8] + w * y3p;
            a[9] = a[9] + w * x4p;      a[10] = a[10] + w * x3p * y;  a[11] = a[11] + w * x2p * y2p;
            a[12] = a[12] + w * x * y3p; a[13] = a[13] + w * y4p;
        }
        for (int q = 0; q < 14; q++) acc[q] = acc[q] + frac[k] * a[q];
    }
    for (int q = 0; q < 14; q++) g.geom[q][c] = acc[q];
}

// get_geometry_incremental_remap (:1105), first pass per edge: remapEdge, orientation check, the stencils
// cellsOnEdgeRemap / edgesOnEdgeRemap, the edge's own two vertices in its frame
__global__ void k_geo_edges(Geo g)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= g.nE) return;
    const int M = g.M, D = g.D, nC = g.nC, nE = g.nE;
    const int c1 = g.cellsOnEdge[(size_t)e * 2], c2 = g.cellsOnEdge[(size_t)e * 2 + 1];
    const int v1 = g.verticesOnEdge[(size_t)e * 2], v2 = g.verticesOnEdge[(size_t)e * 2 + 1];
    // an edge of an owned cell with a cell on both sides (:1235-1262)
    const bool both = c1 >= 1 && c1 <= nC && c2 >= 1 && c2 <= nC;
    const bool remap = both && (c1 <= g.nCS || c2 <= g.nCS);
    g.remapEdge[e] = remap ? 1 : 0;
    double pe[3], t[9], p1[3], p2[3];
    point(g, g.xE, g.yE, g.zE, e, pe);
    const bool haveV = v1 >= 1 && v1 <= g.nV && v2 >= 1 && v2 <= g.nV;
    if (haveV) { point(g, g.xV, g.yV, g.zV, v1 - 1, p1); point(g, g.xV, g.yV, g.zV, v2 - 1, p2); }
    double ex[2] = {0.0, 0.0}, ey[2] = {0.0, 0.0};
    if (g.sphere) {
        global_to_local(pe, t);
        if (haveV) {
            const double w0 = p2[0] - p1[0], w1 = p2[1] - p1[1], w2 = p2[2] - p1[2];
            const double xVector = t[0] * w0 + t[3] * w1 + t[6] * w2, yVector = t[1] * w0 + t[4] * w1 + t[7] * w2;
            g.xve[(size_t)e * NVER + 0] = -0.5 * xVector; g.yve[(size_t)e * NVER + 0] = -0.5 * yVector;
            g.xve[(size_t)e * NVER + 1] = 0.5 * xVector;  g.yve[(size_t)e * NVER + 1] = 0.5 * yVector;
        }
    } else if (remap) {
        g.xve[(size_t)e * NVER + 0] = p1[0] - pe[0]; g.yve[(size_t)e * NVER + 0] = p1[1] - pe[1];
        g.xve[(size_t)e * NVER + 1] = p2[0] - pe[0]; g.yve[(size_t)e * NVER + 1] = p2[1] - pe[1];
    }
    if (!remap) return;
    // C1 must lie to the left of V1 -> V2, tested in the edge frame with the vertices relative to the edge point (:1270-1330)
    {
        double cc[2], pcell[3];
        point(g, g.xC, g.yC, g.zC, c1 - 1, pcell);
        const double *pp[2] = {p1, p2};
        for (int k = 0; k < 2; k++) {
            const double w0 = pp[k][0] - pe[0], w1 = pp[k][1] - pe[1], w2 = pp[k][2] - pe[2];
            if (g.sphere) { ex[k] = t[0] * w0 + t[3] * w1 + t[6] * w2; ey[k] = t[1] * w0 + t[4] * w1 + t[7] * w2; }
            else { ex[k] = w0; ey[k] = w1; }
        }
        const double w0 = pcell[0] - pe[0], w1 = pcell[1] - pe[1], w2 = pcell[2] - pe[2];
        if (g.sphere) { cc[0] = t[0] * w0 + t[3] * w1 + t[6] * w2; cc[1] = t[1] * w0 + t[4] * w1 + t[7] * w2; }
        else { cc[0] = w0; cc[1] = w1; }
        if (!in_half_plane(ex[0], ey[0], ex[1], ey[1], cc[0], cc[1])) atomicOr(g.flags, GEO_BAD_EDGE);
    }
    int EO[NEER] = {0, 0, 0, 0, 0, 0}, CO[NCER] = {0, 0, 0, 0, 0, 0};
    CO[0] = c1; CO[1] = c2;
    for (int side = 0; side < 2; side++) {
        const int cell = side == 0 ? c1 : c2;
        const int n = g.nEdgesOnCell[cell - 1];
        int iMain = 0;
        for (int k = 1; k <= n; k++)
            if (g.edgesOnCell[(size_t)(cell - 1) * M + k - 1] == e + 1) { iMain = k; break; }
        int km = iMain - 1; if (km < 1) km = km + n;
        int kp = iMain + 1; if (kp > n) kp = kp - n;
        const int em = g.edgesOnCell[(size_t)(cell - 1) * M + km - 1], ep = g.edgesOnCell[(size_t)(cell - 1) * M + kp - 1];
        if (side == 0) { EO[0] = em; EO[1] = ep; } else { EO[2] = ep; EO[3] = em; }
    }
    if (D == 4) {
        const int vv[2] = {v1, v2};
        for (int iv = 0; iv < 2; iv++)
            for (int k = 0; k < D; k++) {
                const int en = g.edgesOnVertex[(size_t)(vv[iv] - 1) * D + k];
                if (en >= 1 && en <= nE) {
                    bool isNew = true;
                    for (int q = 0; q < 4; q++)
                        if (en == EO[q] || en == e + 1) { isNew = false; break; }
                    if (isNew) { EO[iv + 4] = en; break; }
                }
            }
    }
    if (D == 3) {
        CO[2] = nC + 1; CO[3] = nC + 1;
        for (int iv = 0; iv < 2; iv++) {
            int en = EO[iv];
            if (en < 1 || en > nE) en = EO[iv + 2];
            if (en < 1 || en > nE) continue;
            for (int k = 0; k < 2; k++) {
                const int cn = g.cellsOnEdge[(size_t)(en - 1) * 2 + k];
                if (cn >= 1 && cn <= nC && cn != CO[0] && cn != CO[1]) CO[iv + 2] = cn;
            }
        }
    } else {
        for (int q = 2; q < 6; q++) CO[q] = nC + 1;
        for (int iv = 0; iv < 2; iv++) {
            int en = EO[iv];
            if (en >= 1 && en <= nE)
                for (int k = 0; k < 2; k++) {
                    const int cn = g.cellsOnEdge[(size_t)(en - 1) * 2 + k];
                    if (cn >= 1 && cn <= nC && cn != CO[0]) CO[iv + 2] = cn;
                }
            en = EO[iv + 2];
            if (en >= 1 && en <= nE)
                for (int k = 0; k < 2; k++) {
                    const int cn = g.cellsOnEdge[(size_t)(en - 1) * 2 + k];
                    if (cn >= 1 && cn <= nC && cn != CO[1]) CO[iv + 4] = cn;
                }
        }
    }
    for (int q = 0; q < NEER; q++) { g.eoer[(size_t)e * NEER + q] = EO[q]; g.coer[(size_t)e * NCER + q] = CO[q]; }
}

// second pass per edge: the far vertices of the side edges in this edge's frame (:1690-1780)
__global__ void k_geo_side_vertices(Geo g)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= g.nE || g.remapEdge[e] != 1) return;
    const int nE = g.nE, nSide = (g.D == 3) ? 4 : 6;
    const int ve[2] = {g.verticesOnEdge[(size_t)e * 2], g.verticesOnEdge[(size_t)e * 2 + 1]};
    double pe[3];
    point(g, g.xE, g.yE, g.zE, e, pe);
    int count = 2;
    for (int q = 0; q < nSide; q++) {
        const int en = g.eoer[(size_t)e * NEER + q];
        if (en < 1 || en > nE) continue;
        const int vn[2] = {g.verticesOnEdge[(size_t)(en - 1) * 2], g.verticesOnEdge[(size_t)(en - 1) * 2 + 1]};
        int n1 = 1, m1 = 1, m2 = 2, far = 0;
        for (int n = 1; n <= 2; n++)
            for (int m = 1; m <= 2; m++)
                if (vn[m - 1] == ve[n - 1]) {
                    n1 = n;
                    if (m == 1) { m1 = 1; m2 = 2; } else { m1 = 2; m2 = 1; }
                    far = vn[m2 - 1];
                    if (!g.sphere) count = count + 1;
                    break;
                }
        if (g.sphere) {
            // this edge's shared vertex plus the side edge's own vector, taken in the side edge's frame as it is
            g.xve[(size_t)e * NVER + q + 2] = g.xve[(size_t)e * NVER + n1 - 1] +
                                              (g.xve[(size_t)(en - 1) * NVER + m2 - 1] - g.xve[(size_t)(en - 1) * NVER + m1 - 1]);
            g.yve[(size_t)e * NVER + q + 2] = g.yve[(size_t)e * NVER + n1 - 1] +
                                              (g.yve[(size_t)(en - 1) * NVER + m2 - 1] - g.yve[(size_t)(en - 1) * NVER + m1 - 1]);
        } else if (count <= NVER && far >= 1) {
            // on a plane the reference fills the slots by counting the side edges that exist (:1745-1772)
            double pf[3];
            point(g, g.xV, g.yV, g.zV, far - 1, pf);
            g.xve[(size_t)e * NVER + count - 1] = pf[0] - pe[0];
            g.yve[(size_t)e * NVER + count - 1] = pf[1] - pe[1];
        }
    }
}

// minLengthEdgesOnVertex (:1785-1805)
__global__ void k_geo_vertices(Geo g)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v > g.nV) return;
    double mn = DBL_MAX;
    if (v < g.nV)
        for (int k = 0; k < g.D; k++) {
            const int e = g.edgesOnVertex[(size_t)v * g.D + k];
            if (e >= 1 && e <= g.nE) {
                double a[3], b[3];
                point(g, g.xV, g.yV, g.zV, g.verticesOnEdge[(size_t)(e - 1) * 2] - 1, a);
                point(g, g.xV, g.yV, g.zV, g.verticesOnEdge[(size_t)(e - 1) * 2 + 1] - 1, b);
                const double w0 = b[0] - a[0], w1 = b[1] - a[1], w2 = b[2] - a[2];
                const double len = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
                if (len < mn) mn = len;
            }
        }
    g.minLen[v] = mn;
}

inline unsigned grid_for(size_t n, int block) { return (unsigned)((n + block - 1) / block); }
inline size_t round_up(size_t n, size_t m) { return (n + m - 1) / m * m; }

template <typename T>
int dev_alloc(ir_handle *h, T **p, size_t count)
{
    void *q = nullptr;
    IR_CUDA(cudaMalloc(&q, sizeof(T) * (count ? count : 1)));
    IR_CUDA(cudaMemsetAsync(q, 0, sizeof(T) * (count ? count : 1), h->stream));
    h->allocs.push_back(q);
    *p = (T *)q;
    return IR_OK;
}

int ensure_stage(ir_handle *h, size_t bytes)
{
    if (h->d.stageBytes >= bytes) return IR_OK;
    if (h->d.stage) {
        IR_CUDA(cudaStreamSynchronize(h->stream));
        IR_CUDA(cudaFree(h->d.stage));
        h->d.stage = nullptr;
        h->d.stageBytes = 0;
    }
    void *q = nullptr;
    IR_CUDA(cudaMalloc(&q, bytes));
    h->d.stage = (double *)q;
    h->d.stageBytes = bytes;
    return IR_OK;
}

// host (n, w) -> device [w][pitch]
template <typename T>
int upload_rows(ir_handle *h, T *dst, const T *host, size_t n, int w, size_t pitch)
{
    int rc = ensure_stage(h, n * w * sizeof(T));
    if (rc) return rc;
    IR_CUDA(cudaMemcpyAsync(h->d.stage, host, n * w * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    IR_LAUNCH((k_rows_in<T>), grid_for(n, 256), 256, h->stream, (const T *)h->d.stage, dst, n, w, pitch);
    h->launches++;
    IR_CUDA(cudaGetLastError());
    IR_CUDA(cudaStreamSynchronize(h->stream));   // the staging area is reused by the next upload
    return IR_OK;
}

// Opt-in (environment IR_B200_PIN_HOST=1 at ir_create): page-lock the host arrays the first time ir_run sees them, so
// the per-step uploads and downloads run at the full PCIe rate.  The pool arrays of the host model live as long as
// the model; a host that frees them earlier calls ir_release_host_memory first.  A range that cannot be registered
// (already registered by someone else, not page-lockable) is simply copied as pageable memory.
void pin_host(ir_handle *h, const void *p, size_t bytes)
{
    if (!h->pinHost || p == nullptr || bytes == 0) return;
    for (auto &r : h->pinned)
        if (r.first == p && r.second >= bytes) return;
    if (cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterDefault) == cudaSuccess) h->pinned.emplace_back(p, bytes);
    else (void)cudaGetLastError();
}

void unpin_all(ir_handle *h)
{
    for (auto &r : h->pinned) cudaHostUnregister(const_cast<void *>(r.first));
    h->pinned.clear();
}

void upwind_free_all(ir_handle *h);   // ir_upwind.cuh

void free_check_buffers(ir_handle *h)
{
    Dev &d = h->d;
    void *bufs[] = {d.lmin, d.lmax, d.sums, d.rowT, d.rowKL, d.monoKey, d.monoDetail};
    for (void *b : bufs)
        if (b) cudaFree(b);
    d.lmin = d.lmax = d.sums = d.monoDetail = nullptr;
    d.rowT = d.rowKL = nullptr;
    d.monoKey = nullptr;
}

// buffers of the optional checks, allocated by the first ir_run that needs them (freed when the tracer set changes)
int ensure_check_buffers(ir_handle *h)
{
    Dev &d = h->d;
    const size_t nRows = (size_t)d.nRows;
    if (h->checkConservation && !d.sums) IR_CUDA(cudaMalloc((void **)&d.sums, sizeof(double) * 2 * nRows));
    if (h->checkMonotonicity && !d.lmin) {
        IR_CUDA(cudaMalloc((void **)&d.lmin, sizeof(double) * nRows * d.nCp));
        IR_CUDA(cudaMalloc((void **)&d.lmax, sizeof(double) * nRows * d.nCp));
        IR_CUDA(cudaMalloc((void **)&d.rowT, sizeof(int) * nRows));
        IR_CUDA(cudaMalloc((void **)&d.rowKL, sizeof(int) * nRows));
        IR_CUDA(cudaMalloc((void **)&d.monoKey, sizeof(unsigned long long)));
        IR_CUDA(cudaMalloc((void **)&d.monoDetail, sizeof(double) * 3));
        int maxLayers = 1;
        for (int nl : h->tracerLayers) maxLayers = nl > maxLayers ? nl : maxLayers;
        d.maxLayers = maxLayers;
        std::vector<int> rowT(nRows), rowKL(nRows);
        for (size_t t = 0; t < h->tracerRow0.size(); t++)
            for (int k = 0; k < d.nK; k++)
                for (int l = 0; l < h->tracerLayers[t]; l++) {
                    const size_t r = (size_t)h->tracerRow0[t] + (size_t)k * h->tracerLayers[t] + l;
                    rowT[r] = (int)t;
                    rowKL[r] = k * maxLayers + l;
                }
        IR_CUDA(cudaMemcpyAsync(d.rowT, rowT.data(), sizeof(int) * nRows, cudaMemcpyHostToDevice, h->stream));
        IR_CUDA(cudaMemcpyAsync(d.rowKL, rowKL.data(), sizeof(int) * nRows, cudaMemcpyHostToDevice, h->stream));
        IR_CUDA(cudaStreamSynchronize(h->stream));     // the vectors go out of scope
    }
    return IR_OK;
}

}  // namespace

// =============================================================================================================== ABI

extern "C" const char *ir_last_error_string(void) { return g_err; }

extern "C" int ir_create(ir_handle **out, const ir_mesh_desc *m, int device)
{
    IR_REQUIRE(out != nullptr && m != nullptr, "handle/mesh is NULL");
    *out = nullptr;
    IR_REQUIRE(m->nCells >= 0 && m->nVertices >= 0 && m->nEdges >= 0, "negative dimension");
    IR_REQUIRE(m->nCellsSolve >= 0 && m->nCellsSolve <= m->nCells, "nCellsSolve out of range");
    IR_REQUIRE(m->maxEdges >= 3 && m->maxEdges <= MAXM, "maxEdges must be 3..8");
    IR_REQUIRE(m->vertexDegree == 3 || m->vertexDegree == 4, "vertexDegree must be 3 or 4");
    IR_REQUIRE(m->nCategories >= 1, "nCategories must be positive");
    IR_REQUIRE(m->nQuadPoints == 3 || m->nQuadPoints == 6, "nQuadPoints must be 3 or 6 (incremental_remap.F:780-787)");
    IR_REQUIRE(m->nEdgesOnCell && m->edgesOnCell && m->cellsOnCell && m->verticesOnCell && m->cellsOnEdge && m->verticesOnEdge,
               "connectivity arrays must not be NULL");
    IR_REQUIRE(m->areaCell && m->dcEdge && m->coeffs_reconstruct, "areaCell / dcEdge / coeffs_reconstruct must not be NULL");
    IR_REQUIRE(m->xVertexOnCell && m->yVertexOnCell && m->xVertexOnEdge && m->yVertexOnEdge && m->remapEdge &&
                   m->cellsOnEdgeRemap && m->edgesOnEdgeRemap,
               "incremental_remap pool arrays must not be NULL");
    IR_REQUIRE(!m->on_a_sphere || m->transGlobalToCell, "transGlobalToCell is needed on a sphere");
    for (int k = 0; k < 14; k++) IR_REQUIRE(m->geomAvgCell[k] != nullptr, "geomAvgCell arrays must not be NULL");
    int dev = device;
    if (dev < 0) IR_CUDA(cudaGetDevice(&dev));
    IR_CUDA(cudaSetDevice(dev));
    ir_handle *h = new ir_handle();
    h->device = dev;
    h->lastMs = 0.f;
    h->launches = 0;
    h->haveTracers = false;
    {
        const char *e = getenv("IR_B200_PIN_HOST");
        h->pinHost = e != nullptr && e[0] != '\0' && e[0] != '0';
    }
    memset(&h->d, 0, sizeof h->d);
    cudaError_t ce = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (ce != cudaSuccess) { set_error("cudaStreamCreate -> %s", cudaGetErrorString(ce)); delete h; return IR_ERR_CUDA; }
    cudaEventCreate(&h->ev0);
    cudaEventCreate(&h->ev1);
    for (cudaEvent_t &ev : h->evK) cudaEventCreate(&ev);
    Dev &d = h->d;
    d.nC = m->nCells; d.nCS = m->nCellsSolve; d.nV = m->nVertices; d.nE = m->nEdges; d.M = m->maxEdges; d.D = m->vertexDegree;
    d.nK = m->nCategories; d.nQP = m->nQuadPoints; d.sphere = m->on_a_sphere ? 1 : 0; d.rotate = m->rotate_cartesian_grid ? 1 : 0;
    d.nCp = round_up((size_t)d.nC + 1, 32); d.nEp = round_up((size_t)d.nE + 1, 32); d.nVp = round_up((size_t)d.nV + 1, 32);
    const size_t nC1 = (size_t)d.nC + 1, nE1 = (size_t)d.nE + 1, nV1 = (size_t)d.nV + 1;
    const int M = d.M;
    int rc = IR_OK;
#define TRY(x) do { if ((rc = (x)) != IR_OK) { ir_destroy(h); return rc; } } while (0)
    TRY(dev_alloc(h, &d.nEdgesOnCell, d.nCp));
    TRY(dev_alloc(h, &d.edgesOnCell, M * d.nCp));
    TRY(dev_alloc(h, &d.cellsOnCell, M * d.nCp));
    TRY(dev_alloc(h, &d.verticesOnCell, M * d.nCp));
    TRY(dev_alloc(h, &d.cellsOnEdge, 2 * nE1));           // host layout (nEdges+1, 2): used by k_cell_edge_signs only
    TRY(dev_alloc(h, &d.verticesOnEdge, 2 * d.nEp));
    TRY(dev_alloc(h, &d.remapEdge, d.nEp));
    TRY(dev_alloc(h, &d.coer, NCER * d.nEp));
    TRY(dev_alloc(h, &d.eoer, NEER * d.nEp));
    TRY(dev_alloc(h, &d.areaCell, d.nCp));
    TRY(dev_alloc(h, &d.sdc, M * d.nCp));
    TRY(dev_alloc(h, &d.fluxSign, M * d.nCp));
    TRY(dev_alloc(h, &d.coef, 3 * M * d.nCp));
    TRY(dev_alloc(h, &d.trans, 6 * d.nCp));
    TRY(dev_alloc(h, &d.xvc, M * d.nCp));
    TRY(dev_alloc(h, &d.yvc, M * d.nCp));
    TRY(dev_alloc(h, &d.xve, NVER * d.nEp));
    TRY(dev_alloc(h, &d.yve, NVER * d.nEp));
    TRY(dev_alloc(h, &d.geom, 14 * d.nCp));
    TRY(dev_alloc(h, &d.u, d.nVp));
    TRY(dev_alloc(h, &d.v, d.nVp));
    TRY(dev_alloc(h, &d.maskCell, d.nCp));
    TRY(dev_alloc(h, &d.maskEdge, d.nEp));
    TRY(dev_alloc(h, &d.iCellTri, NTRI * d.nEp));
    TRY(dev_alloc(h, &d.xq, NTRI * 6 * d.nEp));
    TRY(dev_alloc(h, &d.yq, NTRI * 6 * d.nEp));
    TRY(dev_alloc(h, &d.triArea, NTRI * d.nEp));
    TRY(dev_alloc(h, &d.flags, 1));
    cudaStream_t s = h->stream;
    auto copy1 = [&](void *dst, const void *src, size_t bytes) -> int {
        IR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
        return IR_OK;
    };
    TRY(copy1(d.nEdgesOnCell, m->nEdgesOnCell, nC1 * 4));
    TRY(copy1(d.areaCell, m->areaCell, nC1 * 8));
    TRY(copy1(d.remapEdge, m->remapEdge, nE1 * 4));
    TRY(copy1(d.cellsOnEdge, m->cellsOnEdge, 2 * nE1 * 4));
    for (int k = 0; k < 14; k++) TRY(copy1(d.geom + (size_t)k * d.nCp, m->geomAvgCell[k], nC1 * 8));
    TRY(upload_rows<int>(h, d.edgesOnCell, m->edgesOnCell, nC1, M, d.nCp));
    TRY(upload_rows<int>(h, d.cellsOnCell, m->cellsOnCell, nC1, M, d.nCp));
    TRY(upload_rows<int>(h, d.verticesOnCell, m->verticesOnCell, nC1, M, d.nCp));
    TRY(upload_rows<int>(h, d.verticesOnEdge, m->verticesOnEdge, nE1, 2, d.nEp));
    TRY(upload_rows<int>(h, d.coer, m->cellsOnEdgeRemap, nE1, NCER, d.nEp));
    TRY(upload_rows<int>(h, d.eoer, m->edgesOnEdgeRemap, nE1, NEER, d.nEp));
    TRY(upload_rows<double>(h, d.coef, m->coeffs_reconstruct, nC1, 3 * M, d.nCp));
    TRY(upload_rows<double>(h, d.xvc, m->xVertexOnCell, nC1, M, d.nCp));
    TRY(upload_rows<double>(h, d.yvc, m->yVertexOnCell, nC1, M, d.nCp));
    TRY(upload_rows<double>(h, d.xve, m->xVertexOnEdge, nE1, NVER, d.nEp));
    TRY(upload_rows<double>(h, d.yve, m->yVertexOnEdge, nE1, NVER, d.nEp));
    if (d.sphere && d.nC > 0) {
        TRY(ensure_stage(h, (size_t)d.nC * 9 * 8));
        TRY(copy1(d.stage, m->transGlobalToCell, (size_t)d.nC * 9 * 8));
        IR_LAUNCH((k_trans_in), grid_for(d.nC, 256), 256, s, d.stage, d.trans, (size_t)d.nC, d.nCp);
        h->launches++;
    }
    {   // signed dcEdge and flux signs per (slot, cell)
        double *dc = nullptr;
        TRY(dev_alloc(h, &dc, nE1));
        TRY(copy1(dc, m->dcEdge, nE1 * 8));
        if (d.nC > 0) {
            IR_LAUNCH((k_cell_edge_signs), grid_for(d.nC, 256), 256, s, d, dc);
            h->launches++;
        }
    }
    (void)nV1;
    cudaError_t e2 = cudaStreamSynchronize(s);
    if (e2 == cudaSuccess) e2 = cudaGetLastError();
    if (e2 != cudaSuccess) { set_error("ir_create: %s", cudaGetErrorString(e2)); ir_destroy(h); return IR_ERR_CUDA; }
#undef TRY
    *out = h;
    return IR_OK;
}

extern "C" int ir_init_geometry(const ir_geometry_in *in, const ir_geometry_out *out, int device)
{
    IR_REQUIRE(in != nullptr && out != nullptr, "NULL argument");
    IR_REQUIRE(in->nCells >= 0 && in->nVertices >= 0 && in->nEdges >= 0, "negative dimension");
    IR_REQUIRE(in->nCellsSolve >= 0 && in->nCellsSolve <= in->nCells, "nCellsSolve out of range");
    IR_REQUIRE(in->maxEdges >= 3 && in->maxEdges <= MAXM, "maxEdges must be 3..8");
    IR_REQUIRE(in->vertexDegree == 3 || in->vertexDegree == 4, "vertexDegree must be 3 or 4");
    IR_REQUIRE(in->nEdgesOnCell && in->edgesOnCell && in->verticesOnCell && in->cellsOnEdge && in->verticesOnEdge && in->edgesOnVertex,
               "connectivity arrays must not be NULL");
    IR_REQUIRE(in->xCell && in->yCell && in->zCell && in->xVertex && in->yVertex && in->zVertex && in->xEdge && in->yEdge &&
                   in->zEdge && in->dcEdge && in->dvEdge,
               "coordinate arrays must not be NULL");
    IR_REQUIRE(out->xVertexOnCell && out->yVertexOnCell && out->remapEdge && out->cellsOnEdgeRemap && out->edgesOnEdgeRemap &&
                   out->xVertexOnEdge && out->yVertexOnEdge && out->minLengthEdgesOnVertex,
               "output arrays must not be NULL");
    IR_REQUIRE(!in->on_a_sphere || out->transGlobalToCell, "transGlobalToCell is needed on a sphere");
    for (int k = 0; k < 14; k++) IR_REQUIRE(out->geomAvgCell[k] != nullptr, "geomAvgCell arrays must not be NULL");
    int dev = device;
    if (dev < 0) IR_CUDA(cudaGetDevice(&dev));
    IR_CUDA(cudaSetDevice(dev));
    const size_t nC1 = (size_t)in->nCells + 1, nE1 = (size_t)in->nEdges + 1, nV1 = (size_t)in->nVertices + 1;
    const int M = in->maxEdges, D = in->vertexDegree;
    std::vector<void *> bufs;
    int rc = IR_OK;
    auto up = [&](const void *host, size_t bytes, void **devp) -> int {
        IR_CUDA(cudaMalloc(devp, bytes ? bytes : 1));
        bufs.push_back(*devp);
        if (host) IR_CUDA(cudaMemcpyAsync(*devp, host, bytes, cudaMemcpyHostToDevice, 0));
        else IR_CUDA(cudaMemsetAsync(*devp, 0, bytes ? bytes : 1, 0));
        return IR_OK;
    };
    auto cleanup = [&]() { for (void *b : bufs) cudaFree(b); };
#define TRYG(x) do { if ((rc = (x)) != IR_OK) { cleanup(); return rc; } } while (0)
    Geo g;
    memset(&g, 0, sizeof g);
    g.nC = in->nCells; g.nCS = in->nCellsSolve; g.nV = in->nVertices; g.nE = in->nEdges; g.M = M; g.D = D;
    g.sphere = in->on_a_sphere ? 1 : 0; g.rotate = in->rotate_cartesian_grid ? 1 : 0;
    TRYG(up(in->nEdgesOnCell, nC1 * 4, (void **)&g.nEdgesOnCell));
    TRYG(up(in->edgesOnCell, nC1 * M * 4, (void **)&g.edgesOnCell));
    TRYG(up(in->verticesOnCell, nC1 * M * 4, (void **)&g.verticesOnCell));
    TRYG(up(in->cellsOnEdge, nE1 * 2 * 4, (void **)&g.cellsOnEdge));
    TRYG(up(in->verticesOnEdge, nE1 * 2 * 4, (void **)&g.verticesOnEdge));
    TRYG(up(in->edgesOnVertex, nV1 * D * 4, (void **)&g.edgesOnVertex));
    TRYG(up(in->xCell, nC1 * 8, (void **)&g.xC)); TRYG(up(in->yCell, nC1 * 8, (void **)&g.yC)); TRYG(up(in->zCell, nC1 * 8, (void **)&g.zC));
    TRYG(up(in->xVertex, nV1 * 8, (void **)&g.xV)); TRYG(up(in->yVertex, nV1 * 8, (void **)&g.yV)); TRYG(up(in->zVertex, nV1 * 8, (void **)&g.zV));
    TRYG(up(in->xEdge, nE1 * 8, (void **)&g.xE)); TRYG(up(in->yEdge, nE1 * 8, (void **)&g.yE)); TRYG(up(in->zEdge, nE1 * 8, (void **)&g.zE));
    TRYG(up(in->dcEdge, nE1 * 8, (void **)&g.dcEdge)); TRYG(up(in->dvEdge, nE1 * 8, (void **)&g.dvEdge));
    const size_t nCt = in->nCells > 0 ? (size_t)in->nCells : 1;
    TRYG(up(nullptr, nCt * 9 * 8, (void **)&g.trans));
    TRYG(up(nullptr, nC1 * M * 8, (void **)&g.xvc)); TRYG(up(nullptr, nC1 * M * 8, (void **)&g.yvc));
    TRYG(up(nullptr, nE1 * NVER * 8, (void **)&g.xve)); TRYG(up(nullptr, nE1 * NVER * 8, (void **)&g.yve));
    TRYG(up(nullptr, nV1 * 8, (void **)&g.minLen));
    for (int k = 0; k < 14; k++) TRYG(up(nullptr, nC1 * 8, (void **)&g.geom[k]));
    TRYG(up(nullptr, nE1 * 4, (void **)&g.remapEdge));
    TRYG(up(nullptr, nE1 * NCER * 4, (void **)&g.coer)); TRYG(up(nullptr, nE1 * NEER * 4, (void **)&g.eoer));
    TRYG(up(nullptr, 4, (void **)&g.flags));
    cudaStream_t s0 = 0;
    if (g.nC > 0) IR_LAUNCH((k_geo_cells), grid_for((size_t)g.nC, 128), 128, s0, g);
    if (g.nE > 0) {
        IR_LAUNCH((k_geo_edges), grid_for((size_t)g.nE, 128), 128, s0, g);
        IR_LAUNCH((k_geo_side_vertices), grid_for((size_t)g.nE, 128), 128, s0, g);
    }
    IR_LAUNCH((k_geo_vertices), grid_for(nV1, 128), 128, s0, g);
    auto down = [&](void *host, const void *devp, size_t bytes) -> int {
        IR_CUDA(cudaMemcpyAsync(host, devp, bytes, cudaMemcpyDeviceToHost, 0));
        return IR_OK;
    };
    if (g.sphere && in->nCells > 0) TRYG(down(out->transGlobalToCell, g.trans, (size_t)in->nCells * 9 * 8));
    TRYG(down(out->xVertexOnCell, g.xvc, nC1 * M * 8)); TRYG(down(out->yVertexOnCell, g.yvc, nC1 * M * 8));
    TRYG(down(out->xVertexOnEdge, g.xve, nE1 * NVER * 8)); TRYG(down(out->yVertexOnEdge, g.yve, nE1 * NVER * 8));
    TRYG(down(out->minLengthEdgesOnVertex, g.minLen, nV1 * 8));
    for (int k = 0; k < 14; k++) TRYG(down(out->geomAvgCell[k], g.geom[k], nC1 * 8));
    TRYG(down(out->remapEdge, g.remapEdge, nE1 * 4));
    TRYG(down(out->cellsOnEdgeRemap, g.coer, nE1 * NCER * 4)); TRYG(down(out->edgesOnEdgeRemap, g.eoer, nE1 * NEER * 4));
    int flags = 0;
    TRYG(down(&flags, g.flags, 4));
    cudaError_t ce = cudaStreamSynchronize(s0);
    if (ce == cudaSuccess) ce = cudaGetLastError();
    cleanup();
#undef TRYG
    if (ce != cudaSuccess) { set_error("ir_init_geometry: %s", cudaGetErrorString(ce)); return IR_ERR_CUDA; }
    if (flags & GEO_BAD_EDGE) {
        set_error("IR geometry: cellsOnEdge(1) is not to the left of verticesOnEdge(1) -> (2) (incremental_remap.F:1296)");
        return IR_ERR_MESH;
    }
    if (flags & GEO_BAD_CELL) {
        set_error("IR geometry: the vertices of a cell do not run counter-clockwise (incremental_remap.F:2010)");
        return IR_ERR_MESH;
    }
    return IR_OK;
}

extern "C" int ir_set_tracers(ir_handle *h, int nTracers, const ir_tracer_desc *tr)
{
    IR_REQUIRE(h != nullptr && tr != nullptr, "handle/tracers is NULL");
    IR_REQUIRE(nTracers >= 1, "at least the mass-like field is needed");
    IR_REQUIRE(tr[0].parent == -1, "the first tracer must be the mass-like field (parent = -1)");
    IR_CUDA(cudaSetDevice(h->device));
    Dev &d = h->d;
    const int nK = d.nK;
    std::vector<int> depth(nTracers, 0), hasChild(nTracers, 0);
    int maxLayers = 1;
    for (int t = 0; t < nTracers; t++) {
        IR_REQUIRE(tr[t].nLayers >= 1, "nLayers must be positive");
        if (tr[t].nLayers > maxLayers) maxLayers = tr[t].nLayers;
        if (t == 0) continue;
        IR_REQUIRE(tr[t].parent >= 0 && tr[t].parent < t, "a parent must come before its children; only tracer 0 has none");
        depth[t] = depth[tr[t].parent] + 1;
        IR_REQUIRE(depth[t] < MAX_DEPTH, "at most three parents (incremental_remap.F:6745)");
        const int pl = tr[tr[t].parent].nLayers;
        IR_REQUIRE(pl == 1 || pl == tr[t].nLayers, "a layered parent must have the child's number of layers");
        hasChild[tr[t].parent] = 1;
        IR_REQUIRE(!tr[t].volumeLike || (tr[t].nLayers == 1 && tr[t].parent == 0 && tr[0].nLayers == 1),
                   "volume-like tracers are one-layer children of a one-layer mass field");
    }
    for (int t = 0; t < nTracers; t++)
        IR_REQUIRE(!(hasChild[t] && depth[t] >= 3), "a tracer with three parents cannot have children (incremental_remap.F:3840)");
    // rows grouped by depth, tracers in list order within a depth
    h->tracerRow0.assign(nTracers, 0);
    h->tracerLayers.assign(nTracers, 1);
    h->tracerParent.assign(nTracers, -1);
    h->tracerVolume.assign(nTracers, 0);
    h->tracerDepth = depth;
    h->depthRow0.assign(MAX_DEPTH + 1, 0);
    int row = 0;
    for (int q = 0; q < MAX_DEPTH; q++) {
        h->depthRow0[q] = row;
        for (int t = 0; t < nTracers; t++)
            if (depth[t] == q) { h->tracerRow0[t] = row; row += nK * tr[t].nLayers; }
    }
    h->depthRow0[MAX_DEPTH] = row;
    const int nRows = row;
    h->rows.assign(nRows, RowInfo());
    for (int t = 0; t < nTracers; t++) {
        h->tracerLayers[t] = tr[t].nLayers;
        h->tracerParent[t] = tr[t].parent;
        h->tracerVolume[t] = tr[t].volumeLike ? 1 : 0;
        for (int k = 0; k < nK; k++)
            for (int l = 0; l < tr[t].nLayers; l++) {
                RowInfo &ri = h->rows[h->tracerRow0[t] + k * tr[t].nLayers + l];
                ri.depth = depth[t];
                ri.hasChild = hasChild[t];
                ri.cat = k;
                ri.volumeLike = tr[t].volumeLike ? 1 : 0;
                int q = t, s = depth[t];
                while (q >= 0) {
                    const int ql = tr[q].nLayers;
                    ri.chain[s] = h->tracerRow0[q] + k * ql + (ql == 1 ? 0 : l);
                    q = tr[q].parent;
                    s--;
                }
                for (int z = depth[t] + 1; z < MAX_DEPTH; z++) ri.chain[z] = 0;
            }
    }
    // barycentres and new mass * tracer products are kept for rows that have children only
    int nSlots = 0;
    for (int r = 0; r < nRows; r++) h->rows[r].slot = h->rows[r].hasChild ? nSlots++ : -1;
    for (int r = 0; r < nRows; r++) h->rows[r].parentSlot = h->rows[r].depth > 0 ? h->rows[h->rows[r].chain[h->rows[r].depth - 1]].slot : -1;
    h->nSlots = nSlots;
    // (re)allocate the tracer state; a failure from here on leaves the handle without tracers
    h->haveTracers = false;
    h->sumsHost.clear();
    IR_CUDA(cudaStreamSynchronize(h->stream));
    double **bufs[] = {&d.val, &d.valNew, &d.center, &d.xGrad, &d.yGrad, &d.xBary, &d.yBary, &d.mtpNew, &d.edgeFlux};
    for (double **b : bufs)
        if (*b) { cudaFree(*b); *b = nullptr; }
    free_check_buffers(h);
    if (d.rows) { cudaFree(d.rows); d.rows = nullptr; }
    if (d.catBaseRow) { cudaFree(d.catBaseRow); d.catBaseRow = nullptr; }
    if (d.catLayers) { cudaFree(d.catLayers); d.catLayers = nullptr; }
    d.nRows = nRows;
    // rows of one category, parents first: tracers by depth (list order within a depth), layers innermost
    std::vector<int> baseRow, layers;
    for (int q = 0; q < MAX_DEPTH; q++)
        for (int t = 0; t < nTracers; t++)
            if (depth[t] == q)
                for (int l = 0; l < tr[t].nLayers; l++) { baseRow.push_back(h->tracerRow0[t] + l); layers.push_back(tr[t].nLayers); }
    d.nRowsPerCat = (int)baseRow.size();
    {   // the same list with the tree structure resolved: positions of the ancestors, classes, parent slots
        std::vector<CatRow> cr(baseRow.size());
        std::vector<int> rowToJ(nRows, -1);
        for (size_t j = 0; j < baseRow.size(); j++) rowToJ[baseRow[j]] = (int)j;      // category 0's rows
        int nCls[MAX_DEPTH] = {0, 0, 0, 0}, nSlotJ = 0;
        for (size_t j = 0; j < baseRow.size(); j++) {
            const RowInfo &ri = h->rows[baseRow[j]];
            CatRow &c = cr[j];
            c.baseRow = baseRow[j]; c.layers = layers[j]; c.depth = ri.depth;
            c.cls = ri.depth <= 1 ? nCls[ri.depth] : -1;
            nCls[ri.depth]++;
            for (int q = 0; q < MAX_DEPTH; q++) c.anc[q] = q <= ri.depth ? rowToJ[ri.chain[q]] : -1;
            c.hasChild = ri.hasChild; c.volumeLike = ri.volumeLike;
            c.slotJ = ri.hasChild ? nSlotJ++ : -1;
            c.parentSlotJ = -1;
        }
        for (size_t j = 0; j < cr.size(); j++) {
            CatRow &c = cr[j];
            if (c.depth > 0) c.parentSlotJ = cr[c.anc[c.depth - 1]].slotJ;
            c.vpIdx = c.depth == 0 ? c.cls : nCls[0] + cr[c.anc[1]].cls;
            c.base2 = c.depth >= 2 ? cr[c.anc[2]].baseRow : 0;
            c.layers2 = c.depth >= 2 ? cr[c.anc[2]].layers : 0;
        }
        d.nD0 = nCls[0]; d.nD1 = nCls[1]; d.nSlotsPerCat = nSlotJ;
        if (d.catRows) { cudaFree(d.catRows); d.catRows = nullptr; }
        IR_CUDA(cudaMalloc((void **)&d.catRows, sizeof(CatRow) * cr.size()));
        IR_CUDA(cudaMemcpyAsync(d.catRows, cr.data(), sizeof(CatRow) * cr.size(), cudaMemcpyHostToDevice, h->stream));
        IR_CUDA(cudaStreamSynchronize(h->stream));
    }
    IR_CUDA(cudaMalloc((void **)&d.catBaseRow, sizeof(int) * baseRow.size()));
    IR_CUDA(cudaMalloc((void **)&d.catLayers, sizeof(int) * layers.size()));
    IR_CUDA(cudaMemcpyAsync(d.catBaseRow, baseRow.data(), sizeof(int) * baseRow.size(), cudaMemcpyHostToDevice, h->stream));
    IR_CUDA(cudaMemcpyAsync(d.catLayers, layers.data(), sizeof(int) * layers.size(), cudaMemcpyHostToDevice, h->stream));
    IR_CUDA(cudaStreamSynchronize(h->stream));     // baseRow / layers go out of scope
    const size_t cellBytes = sizeof(double) * (size_t)nRows * d.nCp, edgeBytes = sizeof(double) * (size_t)nRows * d.nEp;
    const size_t slotBytes = sizeof(double) * (size_t)(nSlots > 0 ? nSlots : 1) * d.nCp;
    for (double **b : bufs) {
        const bool parentOnly = (b == &d.xBary || b == &d.yBary || b == &d.mtpNew);
        const size_t bytes = (b == &d.edgeFlux) ? edgeBytes : (parentOnly ? slotBytes : cellBytes);
        IR_CUDA(cudaMalloc((void **)b, bytes));
        IR_CUDA(cudaMemsetAsync(*b, 0, bytes, h->stream));
    }
    IR_CUDA(cudaMalloc((void **)&d.rows, sizeof(RowInfo) * nRows));
    IR_CUDA(cudaMemcpyAsync(d.rows, h->rows.data(), sizeof(RowInfo) * nRows, cudaMemcpyHostToDevice, h->stream));
    int rc = ensure_stage(h, sizeof(double) * ((size_t)d.nC + 1) * nK * maxLayers);
    if (rc) return rc;
    IR_CUDA(cudaStreamSynchronize(h->stream));
    h->haveTracers = true;
    return IR_OK;
}

extern "C" int ir_run(ir_handle *h, int nTracers, const ir_tracer_desc *tr, const double *u, const double *v, double dt)
{
    IR_REQUIRE(h != nullptr && tr != nullptr && u != nullptr && v != nullptr, "NULL argument");
    if (!h->haveTracers) { set_error("ir_run before ir_set_tracers"); return IR_ERR_STATE; }
    IR_REQUIRE(nTracers == (int)h->tracerRow0.size(), "tracer table differs from the one given to ir_set_tracers");
    for (int t = 0; t < nTracers; t++)
        IR_REQUIRE(tr[t].array != nullptr && tr[t].nLayers == h->tracerLayers[t] && tr[t].parent == h->tracerParent[t] &&
                       (tr[t].volumeLike ? 1 : 0) == h->tracerVolume[t],
                   "tracer table differs from the one given to ir_set_tracers");
    IR_CUDA(cudaSetDevice(h->device));
    Dev &d = h->d;
    cudaStream_t s = h->stream;
    const size_t nC1 = (size_t)d.nC + 1;
    const int nK = d.nK;
    // in: tracers (host (nCells+1, nK*nL) -> rows) and velocities
    for (int t = 0; t < nTracers; t++) pin_host(h, tr[t].array, nC1 * nK * tr[t].nLayers * sizeof(double));
    pin_host(h, u, ((size_t)d.nV + 1) * 8);
    pin_host(h, v, ((size_t)d.nV + 1) * 8);
    for (int t = 0; t < nTracers; t++) {
        int rc = upload_rows<double>(h, d.val + (size_t)h->tracerRow0[t] * d.nCp, tr[t].array, nC1, nK * tr[t].nLayers, d.nCp);
        if (rc) return rc;
    }
    IR_CUDA(cudaMemcpyAsync(d.u, u, ((size_t)d.nV + 1) * 8, cudaMemcpyHostToDevice, s));
    IR_CUDA(cudaMemcpyAsync(d.v, v, ((size_t)d.nV + 1) * 8, cudaMemcpyHostToDevice, s));
    IR_CUDA(cudaMemsetAsync(d.flags, 0, sizeof(int), s));
    const bool checks = h->checkConservation != 0 || h->checkMonotonicity != 0;
    memset(&h->report, 0, sizeof h->report);
    if (checks) {
        int rc = ensure_check_buffers(h);
        if (rc) return rc;
        if (h->checkMonotonicity) IR_CUDA(cudaMemsetAsync(d.monoKey, 0xff, sizeof(unsigned long long), s));
    }
    IR_CUDA(cudaEventRecord(h->ev0, s));
    const int massRows = nK * h->tracerLayers[0];
    const unsigned gc = grid_for(nC1, 128), ge = grid_for((size_t)d.nE, 128);
    IR_LAUNCH((k_prepare), gc, 128, s, d, 0, massRows);
    h->launches++;
    if (h->checkMonotonicity && d.nC > 0) {
        IR_LAUNCH((k_check_minmax), dim3(grid_for((size_t)d.nC, 128), (unsigned)d.nRows), 128, s, d);
        h->launches++;
    }
    if (h->checkConservation) {
        IR_LAUNCH_SYNC((k_check_sums), (unsigned)d.nRows, 256, 0, s, d, (const double *)d.val, d.sums);
        h->launches++;
    }
    IR_CUDA(cudaEventRecord(h->evK[0], s));
    if (d.nC > 0) {
        int maxDepth = 0;
        for (int dq : h->tracerDepth) maxDepth = dq > maxDepth ? dq : maxDepth;
        IR_LAUNCH_SYNC((k_reconstruct_coop), grid_for((size_t)d.nC, CL), dim3(CL, CW), ir_reconstruct_smem_bytes(), s, d, maxDepth);
        h->launches++;
    }
    IR_CUDA(cudaEventRecord(h->evK[1], s));
    if (d.nE > 0) {
        IR_LAUNCH((k_triangles), ge, 128, s, d, dt);
        IR_CUDA(cudaEventRecord(h->evK[2], s));
        {
            const size_t smem = ir_flux_smem_bytes(d.nD0, d.nD1);
#ifdef IR_DEVICE_BUILD
            IR_CUDA(cudaFuncSetAttribute(k_fluxes_coop<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            IR_CUDA(cudaFuncSetAttribute(k_fluxes_coop<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#endif
            if (d.nQP == 3) IR_LAUNCH_SYNC((k_fluxes_coop<3>), grid_for((size_t)d.nE, FL), dim3(FL, FR), smem, s, d);
            else IR_LAUNCH_SYNC((k_fluxes_coop<6>), grid_for((size_t)d.nE, FL), dim3(FL, FR), smem, s, d);
        }
        h->launches += 2;
    } else {
        IR_CUDA(cudaEventRecord(h->evK[2], s));
    }
    IR_CUDA(cudaEventRecord(h->evK[3], s));
    {
        int maxDepth = 0;
        for (int dq : h->tracerDepth) maxDepth = dq > maxDepth ? dq : maxDepth;
        const int massOneLayer = h->tracerLayers[0] == 1 ? 1 : 0;
        // with a check on, zap / thickness -> volume run as kernels of their own around the checks
        IR_LAUNCH_SYNC((k_update_coop), grid_for(nC1, CL), dim3(CL, UW), 0, s, d, checks ? 0 : massOneLayer, maxDepth);
        h->launches++;
        if (checks) {
            if (h->checkConservation) {
                IR_LAUNCH_SYNC((k_check_sums), (unsigned)d.nRows, 256, 0, s, d, (const double *)d.valNew, d.sums + d.nRows);
                h->launches++;
            }
            if (massOneLayer && d.nCS > 0) {
                IR_LAUNCH((k_zap), grid_for((size_t)d.nCS, 128), 128, s, d);
                h->launches++;
            }
            if (h->checkMonotonicity && d.nCS > 0) {
                IR_LAUNCH((k_check_mono), dim3(grid_for((size_t)d.nCS, 128), (unsigned)d.nRows), 128, s, d);
                IR_LAUNCH((k_check_mono_detail), 1, 1, s, d);
                h->launches += 2;
            }
            if (massOneLayer) {
                IR_LAUNCH((k_thickness_to_volume), gc, 128, s, d);
                h->launches++;
            }
        }
    }
    IR_CUDA(cudaEventRecord(h->ev1, s));
    IR_CUDA(cudaGetLastError());
    // out
    for (int t = 0; t < nTracers; t++) {
        const int w = nK * tr[t].nLayers;
        IR_LAUNCH((k_tracer_out), grid_for(nC1, 256), 256, s, d.stage, d.valNew + (size_t)h->tracerRow0[t] * d.nCp, nC1, w, d.nCp);
        h->launches++;
        IR_CUDA(cudaMemcpyAsync(tr[t].array, d.stage, nC1 * w * 8, cudaMemcpyDeviceToHost, s));
        IR_CUDA(cudaStreamSynchronize(s));
    }
    int flags = 0;
    IR_CUDA(cudaMemcpyAsync(&flags, d.flags, sizeof(int), cudaMemcpyDeviceToHost, s));
    IR_CUDA(cudaStreamSynchronize(s));
    IR_CUDA(cudaEventElapsedTime(&h->lastMs, h->ev0, h->ev1));
    {
        cudaEvent_t marks[6] = {h->ev0, h->evK[0], h->evK[1], h->evK[2], h->evK[3], h->ev1};
        for (int i = 0; i < 5; i++) IR_CUDA(cudaEventElapsedTime(&h->kernelMs[i], marks[i], marks[i + 1]));
    }
    if (flags & FLAG_NEG_MASS) { set_error("IR: negative mass in a cell (incremental_remap.F:7465)"); return IR_ERR_NEGATIVE_MASS; }
    if (flags & FLAG_NEG_QP) { set_error("IR: negative mass at a quadrature point (incremental_remap.F:6895)"); return IR_ERR_NEGATIVE_MASS_QP; }
    if (flags & FLAG_PARALLEL) { set_error("IR: parallel basis edges in shift_vertices (incremental_remap.F:6415)"); return IR_ERR_PARALLEL_EDGES; }
    if (flags & FLAG_MANY_TRI) { set_error("IR: more than nTriPerEdgeRemap departure triangles on an edge"); return IR_ERR_TOO_MANY_TRIANGLES; }
    if (h->checkConservation) {
        // check_tracer_conservation (:8126): the reference's loop order (tracer, category, layer), first violation
        h->sumsHost.assign(2 * (size_t)d.nRows, 0.0);
        IR_CUDA(cudaMemcpyAsync(h->sumsHost.data(), d.sums, sizeof(double) * 2 * d.nRows, cudaMemcpyDeviceToHost, s));
        IR_CUDA(cudaStreamSynchronize(s));
        for (int t = 0; h->checkConservation == 1 && t < nTracers && !h->report.conservationViolated; t++)
            for (int k = 0; k < nK && !h->report.conservationViolated; k++)
                for (int l = 0; l < tr[t].nLayers; l++) {
                    const size_t r = (size_t)h->tracerRow0[t] + (size_t)k * tr[t].nLayers + l;
                    const double si = h->sumsHost[r], sf = h->sumsHost[d.nRows + r];
                    if (fabs(si) > EPS11) {
                        const double difference = sf - si;
                        const double ratio = difference / si;
                        if (fabs(ratio) > EPS11) {
                            h->report.conservationViolated = 1;
                            h->report.consTracer = t; h->report.consCategory = k + 1; h->report.consLayer = l + 1;
                            h->report.sumInit = si; h->report.sumFinal = sf;
                            break;
                        }
                    }
                }
        if (h->report.conservationViolated) {
            set_error("IR: tracer conservation error (incremental_remap.F:8170): tracer %d, category %d, layer %d: %.17g -> %.17g",
                      h->report.consTracer, h->report.consCategory, h->report.consLayer, h->report.sumInit, h->report.sumFinal);
            return IR_ERR_CONSERVATION;       // the reference aborts before the monotonicity check (:2577-2581)
        }
    }
    if (h->checkMonotonicity) {
        unsigned long long key = ~0ull;
        IR_CUDA(cudaMemcpyAsync(&key, d.monoKey, sizeof key, cudaMemcpyDeviceToHost, s));
        IR_CUDA(cudaStreamSynchronize(s));
        if (key != ~0ull) {
            const unsigned long long KL = (unsigned long long)nK * d.maxLayers;
            const int kl = (int)(key % KL);
            const unsigned long long tc = key / KL;
            const size_t c = (size_t)(tc % ((unsigned long long)d.nC + 1));
            const int t = (int)(tc / ((unsigned long long)d.nC + 1)), k = kl / d.maxLayers, l = kl % d.maxLayers;
            double det[3] = {0.0, 0.0, 0.0};
            IR_CUDA(cudaMemcpyAsync(det, d.monoDetail, sizeof det, cudaMemcpyDeviceToHost, s));
            IR_CUDA(cudaStreamSynchronize(s));
            const double tolMin = EPS11 * fmax(1.0, fabs(det[1])), tolMax = EPS11 * fmax(1.0, fabs(det[2]));
            const bool below = det[0] < det[1] - tolMin;
            h->report.monotonicityViolated = below ? 1 : 2;
            h->report.monoTracer = t; h->report.monoCategory = k + 1; h->report.monoLayer = l + 1; h->report.monoCell = (int)c + 1;
            h->report.newValue = det[0]; h->report.bound = below ? det[1] : det[2]; h->report.tolerance = below ? tolMin : tolMax;
            set_error("IR: monotonicity violation (incremental_remap.F:8530): tracer %d, layer %d, category %d, cell %d: new value %.17g, old %s %.17g",
                      t, l + 1, k + 1, (int)c + 1, det[0], below ? "minimum" : "maximum", h->report.bound);
            return IR_ERR_MONOTONICITY;
        }
    }
    return IR_OK;
}

extern "C" int ir_set_checks(ir_handle *h, int conservation, int monotonicity)
{
    IR_REQUIRE(h != nullptr, "handle is NULL");
    IR_REQUIRE(conservation >= 0 && conservation <= 2, "conservation must be 0 (off), 1 (sums and check) or 2 (sums only)");
    IR_REQUIRE(monotonicity == 0 || monotonicity == 1, "monotonicity must be 0 or 1");
    h->checkConservation = conservation;
    h->checkMonotonicity = monotonicity;
    return IR_OK;
}

extern "C" int ir_fetch_check_report(ir_handle *h, ir_check_report *out)
{
    IR_REQUIRE(h != nullptr && out != nullptr, "NULL argument");
    *out = h->report;
    return IR_OK;
}

extern "C" int ir_fetch_conservation_sums(ir_handle *h, int tracer, double *sumInit, double *sumFinal)
{
    IR_REQUIRE(h != nullptr, "handle is NULL");
    if (!h->haveTracers || h->sumsHost.empty()) { set_error("no conservation sums: ir_set_checks(conservation) and ir_run first"); return IR_ERR_STATE; }
    IR_REQUIRE(tracer >= 0 && tracer < (int)h->tracerRow0.size(), "tracer index out of range");
    const size_t n = (size_t)h->d.nK * h->tracerLayers[tracer], r0 = (size_t)h->tracerRow0[tracer], nRows = (size_t)h->d.nRows;
    for (size_t i = 0; i < n; i++) {      // rows of a tracer: category-major, layer fastest == Fortran (nLayers, nCategories)
        if (sumInit) sumInit[i] = h->sumsHost[r0 + i];
        if (sumFinal) sumFinal[i] = h->sumsHost[nRows + r0 + i];
    }
    return IR_OK;
}

extern "C" int ir_fetch_diagnostics(ir_handle *h, double *xTriangle, double *yTriangle, double *triangleArea,
                                    int *iCellTriangle, int *maskEdge, double *edgeFluxMass)
{
    IR_REQUIRE(h != nullptr, "handle is NULL");
    IR_CUDA(cudaSetDevice(h->device));
    Dev &d = h->d;
    const size_t nE = (size_t)d.nE;
    if (nE == 0) return IR_OK;
    cudaStream_t s = h->stream;
    std::vector<double> tmp;
    std::vector<int> itmp;
    auto fetch = [&](const double *src, size_t rows) -> int {
        tmp.resize(rows * d.nEp);
        IR_CUDA(cudaMemcpyAsync(tmp.data(), src, rows * d.nEp * 8, cudaMemcpyDeviceToHost, s));
        IR_CUDA(cudaStreamSynchronize(s));
        return IR_OK;
    };
    int rc;
    for (int pass = 0; pass < 2; pass++) {
        double *dst = pass == 0 ? xTriangle : yTriangle;
        if (!dst) continue;
        if ((rc = fetch(pass == 0 ? d.xq : d.yq, NTRI * 6))) return rc;
        for (size_t e = 0; e < nE; e++)
            for (int t = 0; t < NTRI; t++)
                for (int q = 0; q < d.nQP; q++) dst[(e * NTRI + t) * d.nQP + q] = tmp[(size_t)(t * 6 + q) * d.nEp + e];
    }
    if (triangleArea) {
        if ((rc = fetch(d.triArea, NTRI))) return rc;
        for (size_t e = 0; e < nE; e++)
            for (int t = 0; t < NTRI; t++) triangleArea[e * NTRI + t] = tmp[(size_t)t * d.nEp + e];
    }
    if (iCellTriangle) {
        itmp.resize(NTRI * d.nEp);
        IR_CUDA(cudaMemcpyAsync(itmp.data(), d.iCellTri, NTRI * d.nEp * 4, cudaMemcpyDeviceToHost, s));
        IR_CUDA(cudaStreamSynchronize(s));
        for (size_t e = 0; e < nE; e++)
            for (int t = 0; t < NTRI; t++) iCellTriangle[e * NTRI + t] = itmp[(size_t)t * d.nEp + e];
    }
    if (maskEdge) {
        IR_CUDA(cudaMemcpyAsync(maskEdge, d.maskEdge, nE * 4, cudaMemcpyDeviceToHost, s));
        IR_CUDA(cudaStreamSynchronize(s));
    }
    if (edgeFluxMass && h->haveTracers) {
        const int w = d.nK * h->tracerLayers[0];
        if ((rc = fetch(d.edgeFlux, (size_t)w))) return rc;
        for (size_t e = 0; e < nE; e++)
            for (int j = 0; j < w; j++) edgeFluxMass[e * w + j] = tmp[(size_t)j * d.nEp + e];
    }
    return IR_OK;
}

extern "C" int ir_fetch_tracer_field(ir_handle *h, int which, int tracer, double *out)
{
    IR_REQUIRE(h != nullptr && out != nullptr, "NULL argument");
    if (!h->haveTracers) { set_error("ir_fetch_tracer_field before ir_set_tracers"); return IR_ERR_STATE; }
    IR_REQUIRE(tracer >= 0 && tracer < (int)h->tracerRow0.size(), "tracer index out of range");
    IR_CUDA(cudaSetDevice(h->device));
    Dev &d = h->d;
    const double *src = nullptr;
    bool onEdges = false;
    switch (which) {
        case IR_FIELD_CENTER: src = d.center; break;
        case IR_FIELD_XGRAD: src = d.xGrad; break;
        case IR_FIELD_YGRAD: src = d.yGrad; break;
        case IR_FIELD_XBARYCENTER: src = d.xBary; break;
        case IR_FIELD_YBARYCENTER: src = d.yBary; break;
        case IR_FIELD_MASS_TRACER_PRODUCT: src = d.mtpNew; break;
        case IR_FIELD_EDGE_FLUX: src = d.edgeFlux; onEdges = true; break;
        default: IR_REQUIRE(false, "unknown field");
    }
    const int w = d.nK * h->tracerLayers[tracer];
    const size_t n = onEdges ? (size_t)d.nE + 1 : (size_t)d.nC + 1, pitch = onEdges ? d.nEp : d.nCp;
    size_t firstRow = (size_t)h->tracerRow0[tracer];
    if (which == IR_FIELD_XBARYCENTER || which == IR_FIELD_YBARYCENTER || which == IR_FIELD_MASS_TRACER_PRODUCT) {
        IR_REQUIRE(h->rows[firstRow].slot >= 0, "barycentres and products are kept for tracers that have children only");
        firstRow = (size_t)h->rows[firstRow].slot;     // the rows of one tracer have consecutive slots
    }
    int rc = ensure_stage(h, n * w * sizeof(double));
    if (rc) return rc;
    IR_LAUNCH((k_tracer_out), grid_for(n, 256), 256, h->stream, d.stage, src + firstRow * pitch, n, w, pitch);
    h->launches++;
    IR_CUDA(cudaGetLastError());
    IR_CUDA(cudaMemcpyAsync(out, d.stage, n * w * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    IR_CUDA(cudaStreamSynchronize(h->stream));
    return IR_OK;
}

extern "C" int ir_last_run_ms(ir_handle *h, float *ms)
{
    IR_REQUIRE(h != nullptr && ms != nullptr, "NULL argument");
    *ms = h->lastMs;
    return IR_OK;
}

extern "C" int ir_last_kernel_ms(ir_handle *h, float *ms)
{
    IR_REQUIRE(h != nullptr && ms != nullptr, "NULL argument");
    for (int i = 0; i < 5; i++) ms[i] = h->kernelMs[i];
    return IR_OK;
}

extern "C" int ir_launch_count(ir_handle *h, long long *n)
{
    IR_REQUIRE(h != nullptr && n != nullptr, "NULL argument");
    *n = h->launches;
    return IR_OK;
}

extern "C" int ir_release_host_memory(ir_handle *h)
{
    IR_REQUIRE(h != nullptr, "handle is NULL");
    IR_CUDA(cudaSetDevice(h->device));
    IR_CUDA(cudaStreamSynchronize(h->stream));
    unpin_all(h);
    return IR_OK;
}

extern "C" int ir_destroy(ir_handle *h)
{
    if (!h) return IR_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    unpin_all(h);
    Dev &d = h->d;
    double *bufs[] = {d.val, d.valNew, d.center, d.xGrad, d.yGrad, d.xBary, d.yBary, d.mtpNew, d.edgeFlux, d.stage};
    for (double *b : bufs)
        if (b) cudaFree(b);
    if (d.rows) cudaFree(d.rows);
    if (d.catBaseRow) cudaFree(d.catBaseRow);
    if (d.catLayers) cudaFree(d.catLayers);
    if (d.catRows) cudaFree(d.catRows);
    free_check_buffers(h);
    upwind_free_all(h);
    for (void *p : h->allocs) cudaFree(p);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (cudaEvent_t ev : h->evK)
        if (ev) cudaEventDestroy(ev);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return IR_OK;
}

// the non-default options: seaice_normal_vectors on the device and the upwind transport
#include "ir_upwind.cuh"
