// evp_kernels.cu -- the two fused FP64 kernels of the EVP subcycle (sm_100a), plus the tiny
// special-boundary kernels, and the launch sequence of subcycle_velocity_solver
// (reference: src/shared/mpas_seaice_velocity_solver.F:2404-2592).
//
// Built with --fmad=false: the operation order below is the reference's source order evaluated
// with separate IEEE multiply / add / divide / sqrt, which is what the CPU oracle executes too
// (gcc -ffp-contract=off), so results are comparable bit for bit.  Both kernels are bound by HBM
// bandwidth (about 0.5 flop/B), not by the FP64 pipe -- see DESIGN.md.
//
// cell kernel  = seaice_strain_tensor_variational (variational.F:575-670)
//              + seaice_stress_tensor_variational / constitutive relation (variational.F:777-975,
//                constitutive_relation.F:178-373)
//              + the per-cell inner sums of seaice_stress_divergence_variational (variational.F:1151-1170)
// vertex kernel = outer sum + denominator of the stress divergence (variational.F:1139-1178)
//              + ocean_stress_coefficient (velocity_solver.F:2986-3082)
//              + solve_velocity / solve_velocity_revised (velocity_solver.F:3096-3342)
#include <stdlib.h>
#include "evp_internal.cuh"

namespace {

// constitutive_relation.F:41-59
constexpr double kEccentricitySquared = 2.0 * 2.0;
constexpr double kPuny = 1.0e-11;
constexpr double kDampingRatioDenominator = 0.86;
constexpr double kDampingRatio = 5.5e-3;
// velocity_solver.F:61-63
constexpr double kSinOceanTurningAngle = 0.0;
constexpr double kCosOceanTurningAngle = 1.0;
// mpas_seaice_constants.F / ice_constants_colpkg.F90:  dragio * rhow, evaluated in FP64 like the reference
constexpr double kDragio = 0.00536;
constexpr double kRhow = 1026.0;

struct CellArgs {
    int nCells;
    int nTiles;                         // grid size: tiles with work
    size_t nCp;
    const uint8_t *__restrict__ nEdges;
    const uint8_t *__restrict__ solveStress;
    const int *__restrict__ tileList;   // compacted list of the tiles with work, or nullptr when every tile has work
    const int *__restrict__ voc;
    const double2 *__restrict__ G;      // dense [j][i][c] (PWL / unknown pattern) or nullptr
    const double2 *__restrict__ Gb;     // banded [k][j][c], k = 0..2 (Wachspress) or nullptr
    const double2 *__restrict__ Suv;
    const double *__restrict__ Sm;
    const double2 *__restrict__ uv;
    const double *__restrict__ tanLat;
    const double *__restrict__ P;
    double2 *__restrict__ sig;
    double *__restrict__ sig12;
    double2 *__restrict__ contrib;
    double *__restrict__ e11;
    double *__restrict__ e22;
    double *__restrict__ e12;
    double *__restrict__ repP;
    double dte, damping;
    evp_halo_view hv;                   // peer-to-peer halo exchange: where the current pass' halo velocities live
};

template <int CR>
__device__ __forceinline__ void constitutive(double &s11, double &s22, double &s12, double e11, double e22,
                                             double e12, double P, double &rep, double dte, double T)
{
    if (CR == EVP_CR_EVP || CR == EVP_CR_EVP_REVISED) {
        const double sd = e11 + e22;
        const double st = e11 - e22;
        const double ss = e12 * 2.0;
        double s1 = s11 + s22;
        double s2 = s11 - s22;
        const double Delta = sqrt(sd * sd + (st * st + ss * ss) / kEccentricitySquared);
        double pc = P / fmax(Delta, kPuny);
        rep = pc * Delta;
        double den;
        if (CR == EVP_CR_EVP) {
            pc = (pc * dte) / (2.0 * T);
            den = 1.0 + (0.5 * dte) / T;
        } else {
            pc = (pc * 2.0 * kDampingRatio) / kDampingRatioDenominator;
            den = 1.0 + (2.0 * kDampingRatio) / kDampingRatioDenominator;
        }
        s1 = (s1 + pc * (sd - Delta)) / den;
        s2 = (s2 + (pc / kEccentricitySquared) * st) / den;
        s12 = (s12 + (pc / kEccentricitySquared) * ss * 0.5) / den;
        s11 = 0.5 * (s1 + s2);
        s22 = 0.5 * (s1 - s2);
    } else if (CR == EVP_CR_LINEAR) {
        s11 = 1.0 * e11;
        s22 = 1.0 * e22;
        s12 = 1.0 * e12;
    }
}

// ---- mbarrier + bulk-copy PTX (cp.async.bulk = SASS UBLKCP on sm_100a; completion counted in bytes on
// an mbarrier in shared memory) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// Shared memory of one block of the cell kernel = one tile of EVP_TILE cells.
template <int M, bool METRIC, bool GBAND>
struct CellSmem {
    static constexpr int GR = GBAND ? 3 * M : M * M;
    double2 G[GR][EVP_TILE];                         // bulk-copied basis gradients of the tile
    double2 S[M * M][EVP_TILE];                      // bulk-copied basisIntegralsU/V
    double Sm[METRIC ? M * M : 1][EVP_TILE];         // bulk-copied basisIntegralsMetric
    double u[M][EVP_TILE], v[M][EVP_TILE];           // velocity at the cell's vertices
    double s11[M][EVP_TILE], s22[M][EVP_TILE], s12[M][EVP_TILE];   // stress after the update
    unsigned long long barG, barS;
};

// ---------------------------------------------------------------------------------------------
// Cell kernel: one block per tile of 32 cells, one thread per (cell, cell-vertex slot);
// blockDim = (32 cells, M slots), i.e. warp j handles slot j of the 32 cells, so every row is read as
// one contiguous segment.
//   start    one thread arms two mbarriers and issues three bulk copies (cp.async.bulk) that bring the
//            tile's basis arrays (contiguous in the tiled layout, 9-18 KB each) into shared memory;
//            they are in flight while phases 0/1 run, and several blocks are resident per SM, so the
//            HBM pipe stays full without holding the data in registers;
//   phase 0  slot j gathers (u,v) and tan(lat)/R of its vertex and stages them in shared memory;
//   phase 1  slot j = stress point j: strain (variational.F:633-668) + constitutive relation
//            (constitutive_relation.F:178-373); new stress to HBM and to shared memory;
//   phase 2  slot j = velocity vertex j: the cell's partial sum of the stress divergence
//            (variational.F:1151-1173) -> contrib[j][c].
// Sums run over i = 1..nEdgesOnCell in increasing i exactly like the reference loops.  With the
// Wachspress basis basisGradientU/V(i,j,c) is exactly zero unless i is j-1, j or j+1 (cyclic;
// wachspress.F:1178-1191); GBAND reads only those three entries, stored in increasing-i order.
// Dropping the +-0 products leaves every partial sum bit-identical for finite velocities (x + +-0 == x,
// and a partial sum that starts at +0 never becomes -0 in round-to-nearest).
// Tiles without any solved cell issue no bulk copy (ice-free ocean costs ~26 B per cell).
// ---------------------------------------------------------------------------------------------
// PHASE 0: the fused kernel.  config_average_variational_strain splits it around the vertex averaging of
// the strains (seaice_average_strains_on_vertex, variational.F:684-763): PHASE 1 stops after the strain and
// stores it, PHASE 2 starts from the stored (averaged) strain.
template <int M, bool METRIC, int CR, bool DIAG, bool GBAND, int PHASE>
__global__ void __launch_bounds__(EVP_TILE *M) evp_cell_kernel(const CellArgs a)
{
    using Smem = CellSmem<M, METRIC, GBAND>;
    extern __shared__ __align__(128) unsigned char evp_smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(evp_smem_raw);
    constexpr int GR = Smem::GR;

    const int cx = threadIdx.x;
    const int j = threadIdx.y;
    // ice-free ocean: tiles without a solved cell and without left-over stress are not launched at all -- the grid
    // runs over the compacted list of tiles with work (evp_refresh_tile_flags), their contrib rows hold zeros
    const size_t tile = a.tileList ? (size_t)a.tileList[blockIdx.x] : (size_t)blockIdx.x;
    const size_t c = tile * EVP_TILE + cx;
    const size_t nCp = a.nCp;
    const bool leader = (cx == 0) && (j == 0);
    int n = 0;
    bool solve = false;
    if (c < (size_t)a.nCells) {
        n = a.nEdges[c];
        solve = a.solveStress[c] == 1;
    }
    if (leader) {
        mbar_init(&sm.barG, 1);
        mbar_init(&sm.barS, 1);
        mbar_fence_init();
    }
    const bool staged = __syncthreads_or(solve) != 0;      // block-uniform; also publishes the mbarrier init
    if (staged && leader) {
        if (PHASE != 2) {
            const double2 *gsrc = (GBAND ? a.Gb : a.G) + tile * (size_t)(GR * EVP_TILE);
            mbar_expect_tx(&sm.barG, (uint32_t)sizeof(sm.G));
            bulk_g2s(&sm.G[0][0], gsrc, (uint32_t)sizeof(sm.G), &sm.barG);
        }
        if (PHASE != 1) {
            mbar_expect_tx(&sm.barS, (uint32_t)(sizeof(sm.S) + (METRIC ? sizeof(sm.Sm) : 0)));
            bulk_g2s(&sm.S[0][0], a.Suv + tile * (size_t)(M * M * EVP_TILE), (uint32_t)sizeof(sm.S), &sm.barS);
            if (METRIC) bulk_g2s(&sm.Sm[0][0], a.Sm + tile * (size_t)(M * M * EVP_TILE), (uint32_t)sizeof(sm.Sm), &sm.barS);
        }
    }

    const bool act = j < n;                       // false for every slot of an out-of-range cell
    const size_t q = (size_t)j * nCp + c;
    double uj = 0.0, vj = 0.0, tj = 0.0, x11 = 0.0, x22 = 0.0, x12 = 0.0;
    if (act) {
        const int vi = a.voc[q];
        if (solve && PHASE != 2) {
            double2 w;
            if (vi >= a.hv.haloFirst) {
                // peer-to-peer halo exchange (evp_halo.cu): the neighbours' vertex kernels stored the velocity of this
                // halo vertex into the buffer of the current pass' parity; make sure their pass is complete.  Only
                // the few threads that gather a halo vertex ever come here, the flags are in local memory
                const int c = *(const volatile int *)a.hv.ctr;
                evp_halo_wait(a.hv, c);
                w = __ldcg(&a.uv[(size_t)vi + a.hv.shift0 + (size_t)(c & 1) * a.hv.stride]);
            } else {
                w = a.uv[vi];
            }
            uj = w.x; vj = w.y;
        }
        if (METRIC) tj = a.tanLat[vi];
        if (PHASE != 1) {
            const double2 s = a.sig[q];
            x11 = s.x; x22 = s.y; x12 = a.sig12[q];
        }
    }
    sm.u[j][cx] = uj;
    sm.v[j][cx] = vj;
    __syncthreads();

    if (act && solve) {
        double e11 = 0.0, e22 = 0.0, e12 = 0.0;
        const double P = (PHASE == 1) ? 0.0 : a.P[c];
        if (PHASE == 2) {
            e11 = a.e11[q]; e22 = a.e22[q]; e12 = a.e12[q];
        } else {
        mbar_wait(&sm.barG, 0);
        if (GBAND) {
            int i0 = j - 1, i1 = j, i2 = j + 1;
            if (j == 0) { i0 = 0; i1 = 1; i2 = n - 1; }
            else if (j == n - 1) { i0 = 0; i1 = n - 2; i2 = n - 1; }
            const double2 g0 = sm.G[0 * M + j][cx];
            const double2 g1 = sm.G[1 * M + j][cx];
            const double2 g2 = sm.G[2 * M + j][cx];
            double u, v;
            u = sm.u[i0][cx]; v = sm.v[i0][cx];
            e11 = e11 + u * g0.x; e22 = e22 + v * g0.y; e12 = e12 + 0.5 * (u * g0.y + v * g0.x);
            u = sm.u[i1][cx]; v = sm.v[i1][cx];
            e11 = e11 + u * g1.x; e22 = e22 + v * g1.y; e12 = e12 + 0.5 * (u * g1.y + v * g1.x);
            u = sm.u[i2][cx]; v = sm.v[i2][cx];
            e11 = e11 + u * g2.x; e22 = e22 + v * g2.y; e12 = e12 + 0.5 * (u * g2.y + v * g2.x);
        } else {
#pragma unroll
            for (int i = 0; i < M; i++) {
                if (i < n) {
                    const double2 g = sm.G[(GBAND ? 0 : j * M) + i][cx];
                    const double u = sm.u[i][cx], v = sm.v[i][cx];
                    e11 = e11 + u * g.x;
                    e22 = e22 + v * g.y;
                    e12 = e12 + 0.5 * (u * g.y + v * g.x);
                }
            }
        }
        // metric terms (variational.F:658-662); tj == 0 when METRIC is off
        e11 = e11 - vj * tj;
        e12 = e12 + uj * tj * 0.5;
        }
        if (PHASE == 1) {
            a.e11[q] = e11; a.e22[q] = e22; a.e12[q] = e12;
            return;                       // no barrier follows in this instantiation
        }
        double rep = 0.0;
        constitutive<CR>(x11, x22, x12, e11, e22, e12, P, rep, a.dte, a.damping);
        if (CR != EVP_CR_NONE) {
            a.sig[q] = make_double2(x11, x22);
            a.sig12[q] = x12;
        }
        if (DIAG) {
            if (PHASE == 0) { a.e11[q] = e11; a.e22[q] = e22; a.e12[q] = e12; }
            if (CR == EVP_CR_EVP || CR == EVP_CR_EVP_REVISED) a.repP[q] = rep;
        }
    } else if (act) {
        // cell not solved: the stress is whatever the host left there (zero after
        // init_subcycle_variables, velocity_solver.F:2335-2345) and still enters the divergence
        if (DIAG && CR == EVP_CR_EVP) a.repP[q] = 0.0;   // variational.F:862
    }
    if (PHASE == 1) return;
    sm.s11[j][cx] = x11;
    sm.s22[j][cx] = x22;
    sm.s12[j][cx] = x12;
    __syncthreads();
    if (!act) return;

    if (!solve) {
        bool anyNonZero = false;
#pragma unroll
        for (int i = 0; i < M; i++)
            if (i < n) anyNonZero |= (sm.s11[i][cx] != 0.0) | (sm.s22[i][cx] != 0.0) | (sm.s12[i][cx] != 0.0);
        if (!anyNonZero) {
            // all-zero stress: every partial sum is exactly +0 (0 - 0*S), skip reading the integrals
            a.contrib[q] = make_double2(0.0, 0.0);
            return;
        }
    }
    if (staged) mbar_wait(&sm.barS, 0);
    double cU = 0.0, cV = 0.0;
#pragma unroll
    for (int i = 0; i < M; i++) {
        if (i < n) {
            const double s11 = sm.s11[i][cx], s22 = sm.s22[i][cx], s12 = sm.s12[i][cx];
            double2 S;
            double m = 0.0;
            if (staged) {
                S = sm.S[j * M + i][cx];
                if (METRIC) m = sm.Sm[j * M + i][cx];
            } else {   // unsolved cell with left-over stress in a tile without any solved cell
                const size_t b = evp_tix(j * M + i, c, M * M);
                S = a.Suv[b];
                if (METRIC) m = a.Sm[b];
            }
            if (METRIC) {
                cU = cU - s11 * S.x - s12 * S.y - s12 * m * tj;
                cV = cV - s22 * S.y - s12 * S.x + s11 * m * tj;
            } else {
                cU = cU - s11 * S.x - s12 * S.y;
                cV = cV - s22 * S.y - s12 * S.x;
            }
        }
    }
    a.contrib[q] = make_double2(cU, cV);
}

// seaice_average_strains_on_vertex (variational.F:684-763): for every OWNED vertex the cell-area weighted
// mean of the strains its vertexDegree cells hold at that vertex, written back to each of them.  A
// (cell, slot) entry belongs to exactly one vertex, so one thread per vertex is race-free; the reference runs
// this loop serially.  gidx is the (slot, cell) index already resolved for the divergence gather.
template <int D>
__global__ void __launch_bounds__(256) evp_average_strain_kernel(int nVerticesSolve, size_t nVp, size_t nCp,
                                                                  const int *__restrict__ gidx,
                                                                  const int *__restrict__ cov,
                                                                  const double *__restrict__ areaCell,
                                                                  double *__restrict__ e11, double *__restrict__ e22,
                                                                  double *__restrict__ e12)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nVerticesSolve) return;
    int idx[D];
    double a11 = 0.0, a22 = 0.0, a12 = 0.0, denominator = 0.0;
#pragma unroll
    for (int s = 0; s < D; s++) {
        idx[s] = gidx[(size_t)s * nVp + v];
        const int c = cov[(size_t)s * nVp + v];
        if (c >= 0) {
            // a cell that does not list the vertex (cellVerticesAtVertex = 0) cannot occur on a valid mesh
            const double area = areaCell[c];
            if (idx[s] >= 0) {
                a11 = a11 + e11[idx[s]] * area;
                a22 = a22 + e22[idx[s]] * area;
                a12 = a12 + e12[idx[s]] * area;
            }
            denominator = denominator + area;
        }
    }
    a11 = a11 / denominator;
    a22 = a22 / denominator;
    a12 = a12 / denominator;
#pragma unroll
    for (int s = 0; s < D; s++) {
        if (idx[s] >= 0) {
            e11[idx[s]] = a11;
            e22[idx[s]] = a22;
            e12[idx[s]] = a12;
        }
    }
}

struct VertexArgs {
    int nVerticesSolve;                 // owned vertices
    const int *__restrict__ vblockList; // compacted list of the 256-vertex blocks with work, or nullptr
    int nBlocks;                        // blocks with work (the grid, except that P2P launches at least one block)
    evp_push_view pv;                   // P2P: peer-to-peer halo exchange fused into this kernel (evp_halo.cu)
    size_t nVp;
    const uint8_t *__restrict__ solveVel;
    const int *__restrict__ gidx;
    const double2 *__restrict__ contrib;
    const double2 *__restrict__ areaDen;
    const double2 *__restrict__ massf;
    const double2 *__restrict__ air;
    const double2 *__restrict__ tilt;
    const double2 *__restrict__ ocnStress;
    const double2 *__restrict__ ocnVel;
    const double2 *__restrict__ uvInit;
    double2 *__restrict__ uv;
    double2 *__restrict__ sdiv;
    double *__restrict__ ocoef;
    double dte, dtDyn, beta;
    int useOcean, oceanType;
    // weak stress divergence (seaice_stress_divergence_weak, weak.F:493-640)
    const int *__restrict__ cov;
    const int2 *__restrict__ wEdgeC;
    const double2 *__restrict__ wNt;
    const double *__restrict__ wDc, *__restrict__ wTanV, *__restrict__ wAreaT;
    const double2 *__restrict__ sigW;
    const double *__restrict__ sigW12;
    double wRadius;
};

// one owned, solved vertex: stress divergence (variational gather or weak line integral), drag coefficient, 2x2 solve;
// returns the vertex' (u,v) after the pass
// ocean_stress_coefficient (velocity_solver.F:2986-3082)
__device__ __forceinline__ double evp_drag_coefficient(int useOcean, int oceanType, double iceAreaVertex, double2 o, double2 w)
{
    double coef = 0.0;
    if (useOcean) {
        if (oceanType == EVP_OCEAN_QUADRATIC) {
            const double du = o.x - w.x, dv = o.y - w.y;
            coef = kDragio * kRhow * iceAreaVertex * sqrt(du * du + dv * dv);
        } else {
            coef = kDragio * kRhow * iceAreaVertex;
        }
    }
    return coef;
}
// solve_velocity / solve_velocity_revised (velocity_solver.F:3096-3342): Cramer's rule in the reference's order
template <int CR>
__device__ __forceinline__ double2 evp_momentum_solve(double sdU, double sdV, double coef, double2 w, double2 mf, double2 air,
                                                      double2 tilt, double2 os, double2 w0, double dte, double dtDyn, double beta)
{
    const double sgn = copysign(1.0, mf.y);
    double l11, l22, r1, r2;
    if (CR == EVP_CR_EVP) {
        l11 = mf.x / dte + coef * kCosOceanTurningAngle;
        l22 = mf.x / dte + coef * kCosOceanTurningAngle;
        r1 = sdU + air.x + tilt.x + coef * os.x + (mf.x * w.x) / dte;
        r2 = sdV + air.y + tilt.y + coef * os.y + (mf.x * w.y) / dte;
    } else {
        l11 = (beta + 1.0) * (mf.x / dtDyn) + coef * kCosOceanTurningAngle;
        l22 = (beta + 1.0) * (mf.x / dtDyn) + coef * kCosOceanTurningAngle;
        r1 = sdU + air.x + tilt.x + coef * os.x + (mf.x * (beta * w.x + w0.x)) / dtDyn;
        r2 = sdV + air.y + tilt.y + coef * os.y + (mf.x * (beta * w.y + w0.y)) / dtDyn;
    }
    const double l12 = -mf.y - coef * kSinOceanTurningAngle * sgn;
    const double l21 = mf.y + coef * kSinOceanTurningAngle * sgn;
    const double den = l11 * l22 - l12 * l21;
    return make_double2((l22 * r1 - l12 * r2) / den, (l11 * r2 - l21 * r1) / den);
}

// COHERENT: contrib and (u,v) were written by OTHER blocks of the same (persistent) kernel: read them from L2.
template <int D, int CR, bool DIAG, bool WEAK, bool COHERENT = false>
__device__ __forceinline__ double2 evp_vertex_solve(const VertexArgs &a, const int v)
{
    double sdU = 0.0, sdV = 0.0;
    const double2 ad = a.areaDen[v];
    if (WEAK) {
        // line integral around the dual triangle: the stress on edge s is the mean of its two cells; a cell
        // that does not exist is the junk cell with zero stress (only at vertices that are never solved)
        double s11V = 0.0, s22V = 0.0, s12V = 0.0;
#pragma unroll
        for (int s = 0; s < D; s++) {
            const size_t q = (size_t)s * a.nVp + v;
            const int c = a.cov[q];
            if (c >= 0) {
                const double2 x = a.sigW[c];
                s11V = s11V + x.x; s22V = s22V + x.y; s12V = s12V + a.sigW12[c];
            }
            const int2 ec = a.wEdgeC[q];
            double e11 = 0.0, e22 = 0.0, e12 = 0.0;
            if (ec.x >= 0) { const double2 x = a.sigW[ec.x]; e11 = e11 + x.x; e22 = e22 + x.y; e12 = e12 + a.sigW12[ec.x]; }
            if (ec.y >= 0) { const double2 x = a.sigW[ec.y]; e11 = e11 + x.x; e22 = e22 + x.y; e12 = e12 + a.sigW12[ec.y]; }
            e11 = e11 / 2.0; e22 = e22 / 2.0; e12 = e12 / 2.0;
            const double2 nv = a.wNt[q];
            const double dc = a.wDc[q];
            sdU = sdU + (e11 * nv.x + e12 * nv.y) * dc;
            sdV = sdV + (e22 * nv.y + e12 * nv.x) * dc;
        }
        s11V = s11V / (double)D; s22V = s22V / (double)D; s12V = s12V / (double)D;
        const double areaT = a.wAreaT[v], t = a.wTanV[v];
        sdU = sdU / areaT;
        sdV = sdV / areaT;
        sdU = sdU - (t * s12V * 2.0) / a.wRadius;
        sdV = sdV + (t * (s11V - s22V)) / a.wRadius;
    } else {
#pragma unroll
        for (int s = 0; s < D; s++) {
            const int idx = a.gidx[(size_t)s * a.nVp + v];
            double2 cc = make_double2(0.0, 0.0);
            if (idx >= 0) cc = COHERENT ? __ldcg(&a.contrib[idx]) : a.contrib[idx];
            sdU = sdU + cc.x;
            sdV = sdV + cc.y;
        }
        sdU = sdU / ad.y;
        sdV = sdV / ad.y;
    }

    const double2 w = COHERENT ? __ldcg(&a.uv[v]) : a.uv[v];
    const double coef = evp_drag_coefficient(a.useOcean, a.oceanType, ad.x, a.ocnVel[v], w);
    if (DIAG) {
        a.sdiv[v] = make_double2(sdU, sdV);
        a.ocoef[v] = coef;
    }
    if (CR == EVP_CR_EVP || CR == EVP_CR_EVP_REVISED) {
        const double2 w0 = (CR == EVP_CR_EVP_REVISED) ? a.uvInit[v] : make_double2(0.0, 0.0);
        const double2 wNew = evp_momentum_solve<CR>(sdU, sdV, coef, w, a.massf[v], a.air[v], a.tilt[v], a.ocnStress[v], w0,
                                                    a.dte, a.dtDyn, a.beta);
        a.uv[v] = wNew;
        return wNew;
    }
    return w;       // linear / none: no velocity update (velocity_solver.F:2529-2541)
}

// solveVel[v]: bit 0 = solveVelocity(v) == 1, bit 1 = boundary-owned vertex (some neighbour rank holds a halo copy;
// set only while the peer-to-peer halo exchange is established).
// P2P: the halo exchange is part of this kernel.  A boundary-owned vertex stores its (u,v) -- new, or unchanged when
// it is not solved -- into the halo buffer of the NEXT pass' parity in every neighbour that holds a copy (NVLink
// stores into peer-mapped memory); the last block to finish publishes "pass c+1 complete" into each neighbour's
// flag word and advances the local pass counter.  The (neighbour, slot) list of a vertex is found from its rank
// among the boundary vertices of its block of 256 (bStart) -- no per-vertex table is read by the other vertices, and
// blocks without a boundary vertex (nearly all of them) leave after one look at bStart: no barrier, no fence, no atomic.
template <int D, int CR, bool DIAG, bool P2P, bool WEAK>
__device__ __forceinline__ void evp_vertex_body(const VertexArgs &a)
{
    int blk;
    if (P2P && (int)blockIdx.x >= a.nBlocks) blk = -1;      // the extra block: publishes the pass when no block pushes
    else if (a.vblockList) blk = a.vblockList[blockIdx.x];
    else if (P2P) blk = a.pv.order[blockIdx.x];             // pushing blocks first
    else blk = (int)blockIdx.x;
    const int v = blk * (int)blockDim.x + (int)threadIdx.x;
    uint8_t mask = 0;
    if (blk >= 0 && v < a.nVerticesSolve) mask = a.solveVel[v];
    double2 wPush = make_double2(0.0, 0.0);
    if (mask & 1) wPush = evp_vertex_solve<D, CR, DIAG, WEAK>(a, v);
    if (P2P) {
        // blocks of 256 vertices without a boundary-owned vertex (nearly all) are done here
        const int e0 = blk >= 0 ? a.pv.bStart[blk] : 0;
        const int nPushers = blk >= 0 ? a.pv.bStart[blk + 1] - e0 : 0;
        if (blk >= 0 && (nPushers == 0 || (a.pv.dbg & 4))) return;
        if (blk < 0 && a.pv.nPushBlocks > 0) return;
        __shared__ int warpCount[8];
        const bool bnd = (mask & 2) != 0;
        if (bnd && !(mask & 1)) wPush = a.uv[v];            // not solved: its unchanged velocity fills the other parity
        const unsigned ballot = __ballot_sync(0xffffffffu, bnd);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) warpCount[warp] = __popc(ballot);
        __syncthreads();
        if (bnd) {
            int e = e0 + __popc(ballot & ((1u << lane) - 1u));
            for (int w = 0; w < warp; w++) e += warpCount[w];
            const int next = (*(const volatile int *)a.pv.ctr + 1) & 1;
            const int t1 = a.pv.pushStart[e + 1];
            for (int t = a.pv.pushStart[e]; t < t1; t++) {
                const int2 pd = a.pv.push[t];
                if (a.pv.dbg & 1) a.sdiv[a.nVp - 1 - (t & 31)] = wPush;
                else a.pv.peerUv[pd.x][(size_t)pd.y + (size_t)next * a.pv.peerStride[pd.x]] = wPush;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            // one system-scope fence per pushing block: cumulative over the stores of the threads that arrived at the
            // barrier above; then the ticket -- the last pushing block publishes the pass
            if (a.pv.dbg & 8) __threadfence();
            else if (!(a.pv.dbg & 2)) __threadfence_system();
            const unsigned last = a.pv.nPushBlocks > 0 ? (unsigned)a.pv.nPushBlocks - 1u : 0u;
            const unsigned ticket = atomicAdd(a.pv.done, 1u);
            if (ticket == last) {
                // ONE system-scope fence (cumulative over the tickets observed, i.e. over every pushing block's
                // stores), then the flags as plain system-scope stores: a release store per neighbour would repeat it
                __threadfence_system();
                *(volatile unsigned *)a.pv.done = 0;
                const int c1 = *(volatile int *)a.pv.ctr + 1;
                for (int k = 0; k < a.pv.nNb; k++) evp_st_relaxed_sys(a.pv.peerFlag[k], c1);
                *(volatile int *)a.pv.ctr = c1;
            }
        }
    }
}

template <int D, int CR, bool DIAG, bool WEAK>
__global__ void __launch_bounds__(256) evp_vertex_kernel(const VertexArgs a)
{
    evp_vertex_body<D, CR, DIAG, false, WEAK>(a);
}
// 5 blocks per SM like the plain kernel (48 registers): the exchange code must not cost the interior blocks occupancy
template <int D, int CR, bool DIAG>
__global__ void __launch_bounds__(256, 5) evp_vertex_p2p_kernel(const VertexArgs a)
{
    evp_vertex_body<D, CR, DIAG, true, false>(a);
}

// ---------------------------------------------------------------------------------------------
// Persistent whole-loop kernel for meshes small enough that every tile of 32 cells can have a block of its own
// resident at the same time (square 7 708 cells, QU240 10 242 cells: BASELINE configs[1] and [2]).  There the
// graph of 2 x nSub kernel nodes is bound by launch latency (about 3.5 us per node against about 1 us of work), so
// ONE cooperative launch runs all nSub subcycles and nothing but the per-cell divergence sums travels through L2:
//   * the tile's basis arrays are bulk-copied into shared memory once and stay there for the whole loop;
//   * thread (cell, slot j) keeps in registers, from the first subcycle to the last, the stresses of its stress
//     point AND the complete state of the vertex in slot j -- its velocity, mass, forcing, the positions of the
//     vertexDegree divergence sums it gathers.  The vertexDegree threads that share a vertex each solve it, from
//     identical inputs with identical instructions, so their copies never differ and no velocity is exchanged;
//   * per subcycle: (u,v) -> shared -> strain, stress -> per-cell divergence sums -> L2 (double-buffered) |
//     ONE grid barrier | gather vertexDegree sums from L2 -> drag coefficient, 2x2 solve -> (u,v) in registers.
// The barrier has no atomics (same-address atomics from hundreds of blocks serialise in L2): every block stores its
// epoch into a slot of its own, block 0 polls the slots and stores the release word the others poll.
// Arithmetic, operation order and summation order are those of evp_cell_kernel / evp_vertex_solve: results are
// bit-identical to the two-kernel path (tests/test_gpu_parity.py::test_graph_and_stream_paths_agree).
// ---------------------------------------------------------------------------------------------
struct PersistArgs {
    CellArgs c;
    VertexArgs v;
    int nSub;
    double2 *contrib2;     // second buffer of the divergence sums
    unsigned *arrive;      // [grid] epoch every block has reached, zero at launch
    unsigned *release;     // epoch everybody may pass, zero at launch
};

__device__ __forceinline__ unsigned evp_ld_acquire_gpu_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void evp_st_release_gpu_u32(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// all blocks of the (cooperative) grid have finished epoch `e` (1, 2, ...)
__device__ __forceinline__ void evp_grid_barrier(unsigned *arrive, unsigned *release, unsigned e)
{
    const unsigned tid = threadIdx.y * blockDim.x + threadIdx.x, nThreads = blockDim.x * blockDim.y;
    __syncthreads();
    if (tid == 0) evp_st_release_gpu_u32(arrive + blockIdx.x, e);     // cumulative over the block's stores before the barrier
    if (blockIdx.x == 0) {
        for (unsigned b = tid; b < gridDim.x; b += nThreads)
            while (evp_ld_acquire_gpu_u32(arrive + b) < e) { }
        __syncthreads();
        if (tid == 0) evp_st_release_gpu_u32(release, e);
    }
    if (tid == 0)
        while (evp_ld_acquire_gpu_u32(release) < e) { }
    __syncthreads();
}

template <int M, bool METRIC, int CR, int D>
__global__ void __launch_bounds__(EVP_TILE *M, 3) evp_persistent_kernel(const PersistArgs p)
{
    using Smem = CellSmem<M, METRIC, true>;
    extern __shared__ __align__(128) unsigned char evp_smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(evp_smem_raw);
    const CellArgs &a = p.c;
    const VertexArgs &va = p.v;
    const int cx = threadIdx.x, j = threadIdx.y;
    const size_t tile = blockIdx.x;
    const size_t c = tile * EVP_TILE + cx;
    const size_t nCp = a.nCp;
    const bool leader = (cx == 0) && (j == 0);
    int n = 0;
    bool solve = false;
    if (c < (size_t)a.nCells) {
        n = a.nEdges[c];
        solve = a.solveStress[c] == 1;
    }
    if (leader) {
        mbar_init(&sm.barG, 1);
        mbar_init(&sm.barS, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (leader) {       // the basis of the tile: fetched once, resident for all subcycles
        mbar_expect_tx(&sm.barG, (uint32_t)sizeof(sm.G));
        bulk_g2s(&sm.G[0][0], a.Gb + tile * (size_t)(Smem::GR * EVP_TILE), (uint32_t)sizeof(sm.G), &sm.barG);
        mbar_expect_tx(&sm.barS, (uint32_t)(sizeof(sm.S) + (METRIC ? sizeof(sm.Sm) : 0)));
        bulk_g2s(&sm.S[0][0], a.Suv + tile * (size_t)(M * M * EVP_TILE), (uint32_t)sizeof(sm.S), &sm.barS);
        if (METRIC) bulk_g2s(&sm.Sm[0][0], a.Sm + tile * (size_t)(M * M * EVP_TILE), (uint32_t)sizeof(sm.Sm), &sm.barS);
    }
    const bool act = j < n;
    const size_t q = (size_t)j * nCp + c;
    // ---- the stress point (cell, j) ----
    double tj = 0.0, x11 = 0.0, x22 = 0.0, x12 = 0.0, P = 0.0;
    // ---- the vertex in slot j: everything the momentum solve reads, and where its divergence sums come from ----
    int vi = 0, g[D];
    bool vsolve = false, writer = false;
    double2 w = make_double2(0.0, 0.0), ad = w, mf = w, air = w, tilt = w, os = w, ov = w, w0 = w;
#pragma unroll
    for (int s = 0; s < D; s++) g[s] = -1;
    if (act) {
        vi = a.voc[q];
        if (METRIC) tj = a.tanLat[vi];
        const double2 s0 = a.sig[q];
        x11 = s0.x; x22 = s0.y; x12 = a.sig12[q];
        P = a.P[c];
        w = a.uv[vi];
        vsolve = vi < va.nVerticesSolve && (va.solveVel[vi] & 1);
        if (vsolve) {
#pragma unroll
            for (int s = 0; s < D; s++) g[s] = va.gidx[(size_t)s * va.nVp + vi];
            // one of the threads that share the vertex stores its results: the one sitting on the first divergence sum
            writer = false;
#pragma unroll
            for (int s = D - 1; s >= 0; s--) if (g[s] >= 0) writer = (size_t)g[s] == q;
            ad = va.areaDen[vi]; mf = va.massf[vi]; air = va.air[vi]; tilt = va.tilt[vi];
            os = va.ocnStress[vi]; ov = va.ocnVel[vi];
            if (CR == EVP_CR_EVP_REVISED) w0 = va.uvInit[vi];
        }
    }
    int i0 = j - 1, i1 = j, i2 = j + 1;
    if (j == 0) { i0 = 0; i1 = 1; i2 = n - 1; }
    else if (j == n - 1) { i0 = 0; i1 = n - 2; i2 = n - 1; }
    mbar_wait(&sm.barG, 0);
    mbar_wait(&sm.barS, 0);

    for (int k = 0; k < p.nSub; k++) {
        const bool diag = k == p.nSub - 1;
        // the last subcycle leaves its sums in the regular buffer
        double2 *contrib = ((p.nSub - 1 - k) & 1) ? p.contrib2 : a.contrib;
        const double uj = (act && solve) ? w.x : 0.0, vj = (act && solve) ? w.y : 0.0;
        sm.u[j][cx] = uj;
        sm.v[j][cx] = vj;
        __syncthreads();
        if (act && solve) {
            double e11 = 0.0, e22 = 0.0, e12 = 0.0;
            const double2 g0 = sm.G[0 * M + j][cx];
            const double2 g1 = sm.G[1 * M + j][cx];
            const double2 g2 = sm.G[2 * M + j][cx];
            double u, v;
            u = sm.u[i0][cx]; v = sm.v[i0][cx];
            e11 = e11 + u * g0.x; e22 = e22 + v * g0.y; e12 = e12 + 0.5 * (u * g0.y + v * g0.x);
            u = sm.u[i1][cx]; v = sm.v[i1][cx];
            e11 = e11 + u * g1.x; e22 = e22 + v * g1.y; e12 = e12 + 0.5 * (u * g1.y + v * g1.x);
            u = sm.u[i2][cx]; v = sm.v[i2][cx];
            e11 = e11 + u * g2.x; e22 = e22 + v * g2.y; e12 = e12 + 0.5 * (u * g2.y + v * g2.x);
            e11 = e11 - vj * tj;
            e12 = e12 + uj * tj * 0.5;
            double rep = 0.0;
            constitutive<CR>(x11, x22, x12, e11, e22, e12, P, rep, a.dte, a.damping);
            if (diag) {
                a.e11[q] = e11; a.e22[q] = e22; a.e12[q] = e12;
                a.repP[q] = rep;
            }
        } else if (act && diag && CR == EVP_CR_EVP) {
            a.repP[q] = 0.0;                             // variational.F:862
        }
        sm.s11[j][cx] = x11;
        sm.s22[j][cx] = x22;
        sm.s12[j][cx] = x12;
        __syncthreads();
        if (act) {
            double cU = 0.0, cV = 0.0;
#pragma unroll
            for (int i = 0; i < M; i++) {
                if (i < n) {
                    const double s11 = sm.s11[i][cx], s22 = sm.s22[i][cx], s12 = sm.s12[i][cx];
                    const double2 S = sm.S[j * M + i][cx];
                    if (METRIC) {
                        const double m = sm.Sm[j * M + i][cx];
                        cU = cU - s11 * S.x - s12 * S.y - s12 * m * tj;
                        cV = cV - s22 * S.y - s12 * S.x + s11 * m * tj;
                    } else {
                        cU = cU - s11 * S.x - s12 * S.y;
                        cV = cV - s22 * S.y - s12 * S.x;
                    }
                }
            }
            contrib[q] = make_double2(cU, cV);
        }
        evp_grid_barrier(p.arrive, p.release, (unsigned)(k + 1));
        if (vsolve) {
            double sdU = 0.0, sdV = 0.0;
#pragma unroll
            for (int s = 0; s < D; s++) {
                double2 cc = make_double2(0.0, 0.0);
                if (g[s] >= 0) cc = __ldcg(&contrib[g[s]]);      // written by other blocks: from L2
                sdU = sdU + cc.x;
                sdV = sdV + cc.y;
            }
            sdU = sdU / ad.y;
            sdV = sdV / ad.y;
            const double coef = evp_drag_coefficient(va.useOcean, va.oceanType, ad.x, ov, w);
            if (diag && writer) {
                va.sdiv[vi] = make_double2(sdU, sdV);
                va.ocoef[vi] = coef;
            }
            w = evp_momentum_solve<CR>(sdU, sdV, coef, w, mf, air, tilt, os, w0, va.dte, va.dtDyn, va.beta);
        }
    }
    if (act && solve) {
        a.sig[q] = make_double2(x11, x22);
        a.sig12[q] = x12;
    }
    if (vsolve && writer) va.uv[vi] = w;
}

// seaice_set_special_boundaries_velocity (special_boundaries.F:301-324) with the sequential
// in-place semantics resolved on the host into (source-before-the-loop, sign) pairs.
__global__ void evp_sb_gather(int n, const int *__restrict__ src, const double *__restrict__ sign,
                              const double2 *__restrict__ uv, double2 *__restrict__ tmp)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double sg = sign[k];
    double2 w = make_double2(0.0, 0.0);
    if (sg != 0.0) {
        w = uv[src[k]];
        if (sg < 0.0) { w.x = -w.x; w.y = -w.y; }
    }
    tmp[k] = w;
}
__global__ void evp_sb_scatter(int n, const int *__restrict__ dst, const double2 *__restrict__ tmp,
                               double2 *__restrict__ uv)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uv[dst[k]] = tmp[k];
}

template <int M, bool METRIC, int CR, bool DIAG, bool GBAND, int PHASE>
int launch_cell_k(const CellArgs &a, cudaStream_t s)
{
    auto kern = evp_cell_kernel<M, METRIC, CR, DIAG, GBAND, PHASE>;
    constexpr size_t smem = sizeof(CellSmem<M, METRIC, GBAND>);
    // opt in to > 48 KB of dynamic shared memory: per device and cheap, so set on every launch rather than cached
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1;
    const dim3 block(EVP_TILE, M);
    const unsigned grid = (unsigned)a.nTiles;
    if (grid) kern<<<grid, block, smem, s>>>(a);
    return 0;
}
template <int M, bool METRIC, int CR, bool DIAG, bool GBAND>
int launch_cell_p(const CellArgs &a, int phase, cudaStream_t s)
{
    if (phase == 1) return launch_cell_k<M, METRIC, EVP_CR_NONE, false, GBAND, 1>(a, s);   // strain only
    if (phase == 2) return launch_cell_k<M, METRIC, CR, DIAG, GBAND, 2>(a, s);
    return launch_cell_k<M, METRIC, CR, DIAG, GBAND, 0>(a, s);
}
template <int M, bool METRIC, int CR>
int launch_cell_d(const CellArgs &a, bool diag, int phase, cudaStream_t s)
{
    const bool band = a.Gb != nullptr;
    if (diag) return band ? launch_cell_p<M, METRIC, CR, true, true>(a, phase, s) : launch_cell_p<M, METRIC, CR, true, false>(a, phase, s);
    return band ? launch_cell_p<M, METRIC, CR, false, true>(a, phase, s) : launch_cell_p<M, METRIC, CR, false, false>(a, phase, s);
}
template <int M, bool METRIC>
int launch_cell_cr(const CellArgs &a, int cr, bool diag, int phase, cudaStream_t s)
{
    switch (cr) {
    case EVP_CR_EVP: return launch_cell_d<M, METRIC, EVP_CR_EVP>(a, diag, phase, s);
    case EVP_CR_EVP_REVISED: return launch_cell_d<M, METRIC, EVP_CR_EVP_REVISED>(a, diag, phase, s);
    case EVP_CR_LINEAR: return launch_cell_d<M, METRIC, EVP_CR_LINEAR>(a, diag, phase, s);
    default: return launch_cell_d<M, METRIC, EVP_CR_NONE>(a, diag, phase, s);
    }
}
template <int M>
int launch_cell_m(const CellArgs &a, bool metric, int cr, bool diag, int phase, cudaStream_t s)
{
    return metric ? launch_cell_cr<M, true>(a, cr, diag, phase, s) : launch_cell_cr<M, false>(a, cr, diag, phase, s);
}

template <int D, int CR, bool WEAK>
int launch_vertex_w(const VertexArgs &a, bool diag, cudaStream_t s)
{
    const int block = 256;
    const bool p2p = !WEAK && a.pv.ctr != nullptr;
    // peer-to-peer: one extra block, which publishes the pass when this rank has no boundary-owned vertex to push
    const int grid = p2p ? a.nBlocks + 1 : a.nBlocks;
    if (grid == 0) return 0;
    if (p2p) {
        if (diag) evp_vertex_p2p_kernel<D, CR, true><<<grid, block, 0, s>>>(a);
        else      evp_vertex_p2p_kernel<D, CR, false><<<grid, block, 0, s>>>(a);
    } else {
        if (diag) evp_vertex_kernel<D, CR, true, WEAK><<<grid, block, 0, s>>>(a);
        else      evp_vertex_kernel<D, CR, false, WEAK><<<grid, block, 0, s>>>(a);
    }
    return 0;
}
template <int D, int CR>
int launch_vertex_d(const VertexArgs &a, bool diag, cudaStream_t s)
{
    return a.sigW ? launch_vertex_w<D, CR, true>(a, diag, s) : launch_vertex_w<D, CR, false>(a, diag, s);
}
template <int D>
int launch_vertex_cr(const VertexArgs &a, int cr, bool diag, cudaStream_t s)
{
    switch (cr) {
    case EVP_CR_EVP: return launch_vertex_d<D, EVP_CR_EVP>(a, diag, s);
    case EVP_CR_EVP_REVISED: return launch_vertex_d<D, EVP_CR_EVP_REVISED>(a, diag, s);
    default: return launch_vertex_d<D, EVP_CR_NONE>(a, diag, s);   // linear / none: no velocity update
    }
}

}  // namespace

static int enqueue_cell_phase(evp_handle *h, bool diag, int phase, cudaStream_t s);

// strain -> [vertex averaging] -> stress -> per-cell divergence sums
int evp_enqueue_cell_pass(evp_handle *h, bool diag, cudaStream_t s)
{
    if (h->nCells == 0) return EVP_OK;
    int rc;
    if (h->opt.strain_scheme == EVP_SCHEME_WEAK) {
        if ((rc = evp_enqueue_weak_cell_pass(h, diag, s))) return rc;
        if (h->opt.stress_divergence_scheme == EVP_SCHEME_WEAK) return EVP_OK;
        // weak strain + variational divergence: interpolate_strains_weak_to_variational, then the stress /
        // divergence half of the cell kernel from the stored strain
        if ((rc = evp_enqueue_weak_to_variational(h, s))) return rc;
        return enqueue_cell_phase(h, diag, 2, s);
    }
    if (!h->opt.average_variational_strain) return enqueue_cell_phase(h, diag, 0, s);
    if ((rc = enqueue_cell_phase(h, diag, 1, s))) return rc;
    if (h->nVerticesSolve) {
        const int block = 256, grid = (h->nVerticesSolve + block - 1) / block;
        if (h->D == 3)
            evp_average_strain_kernel<3><<<grid, block, 0, s>>>(h->nVerticesSolve, h->nVp, h->nCp, h->d.gidx, h->d.cov,
                                                                h->d.areaCell, h->d.e11, h->d.e22, h->d.e12);
        else
            evp_average_strain_kernel<4><<<grid, block, 0, s>>>(h->nVerticesSolve, h->nVp, h->nCp, h->d.gidx, h->d.cov,
                                                                h->d.areaCell, h->d.e11, h->d.e22, h->d.e12);
        EVP_CUDA(cudaGetLastError());
    }
    return enqueue_cell_phase(h, diag, 2, s);
}

static void fill_cell_args(evp_handle *h, CellArgs &a)
{
    a.nCells = h->nCells; a.nCp = h->nCp;
    const int allTiles = (h->nCells + EVP_TILE - 1) / EVP_TILE;
    a.nTiles = h->nActiveTiles < 0 ? allTiles : h->nActiveTiles;
    a.tileList = (a.nTiles == allTiles) ? nullptr : h->d.tileList;
    a.nEdges = h->d.nEdges; a.solveStress = h->d.solveStress; a.voc = h->d.voc;
    a.G = h->d.G; a.Gb = h->d.Gb; a.Suv = h->d.Suv; a.Sm = h->d.Sm;
    a.uv = h->d.uv; a.tanLat = h->d.tanLat; a.P = h->d.P;
    a.sig = h->d.sig; a.sig12 = h->d.sig12; a.contrib = h->d.contrib;
    a.e11 = h->d.e11; a.e22 = h->d.e22; a.e12 = h->d.e12; a.repP = h->d.repP;
    a.dte = h->opt.elasticTimeStep; a.damping = h->opt.dampingTimescale;
    a.hv = evp_halo_get_view(h);
}

static int enqueue_cell_phase(evp_handle *h, bool diag, int phase, cudaStream_t s)
{
    CellArgs a;
    fill_cell_args(h, a);
    const int cr = h->opt.constitutive_relation_type;
    int rc = 0;
    switch (h->M) {
    case 4: rc = launch_cell_m<4>(a, h->metric, cr, diag, phase, s); break;
    case 6: rc = launch_cell_m<6>(a, h->metric, cr, diag, phase, s); break;
    case 7:
    case 8: rc = launch_cell_m<8>(a, h->metric, cr, diag, phase, s); break;
    default: evp_set_error("unsupported maxEdges %d", h->M); return EVP_ERR_ARGUMENT;
    }
    if (rc) { evp_set_error("cell kernel: cannot opt in to its dynamic shared memory size"); return EVP_ERR_CUDA; }
    EVP_CUDA(cudaGetLastError());
    return EVP_OK;
}

static void fill_vertex_args(evp_handle *h, VertexArgs &a)
{
    a.pv = evp_halo_get_push(h);
    a.nVerticesSolve = h->nVerticesSolve; a.nVp = h->nVp;
    const int allBlocks = (h->nVerticesSolve + 255) / 256;
    a.nBlocks = h->nActiveVBlocks < 0 ? allBlocks : h->nActiveVBlocks;
    a.vblockList = (a.nBlocks == allBlocks) ? nullptr : h->d.vblockList;
    a.solveVel = h->d.solveVel; a.gidx = h->d.gidx; a.contrib = h->d.contrib;
    a.areaDen = h->d.areaDen; a.massf = h->d.massf; a.air = h->d.air; a.tilt = h->d.tilt;
    a.ocnStress = h->d.ocnStress; a.ocnVel = h->d.ocnVel; a.uvInit = h->d.uvInit;
    a.uv = h->d.uv; a.sdiv = h->d.sdiv; a.ocoef = h->d.ocoef;
    a.dte = h->opt.elasticTimeStep; a.dtDyn = h->opt.dynamicsTimeStep;
    a.beta = h->opt.numericalInertiaCoefficient;
    a.useOcean = h->opt.use_ocean_stress; a.oceanType = h->opt.ocean_stress_type;
    const bool weak = h->opt.stress_divergence_scheme == EVP_SCHEME_WEAK;
    a.cov = h->d.cov; a.wEdgeC = h->d.wEdgeC; a.wNt = h->d.wNt; a.wDc = h->d.wDc; a.wTanV = h->d.wTanV;
    a.wAreaT = h->d.wAreaT; a.sigW = weak ? h->d.sigW : nullptr; a.sigW12 = h->d.sigW12; a.wRadius = h->d.wRadius;
}

int evp_enqueue_vertex_pass(evp_handle *h, bool diag, cudaStream_t s)
{
    VertexArgs a;
    fill_vertex_args(h, a);
    if (h->nVerticesSolve == 0 && !a.pv.ctr) return EVP_OK;
    const int cr = h->opt.constitutive_relation_type;
    switch (h->D) {
    case 3: launch_vertex_cr<3>(a, cr, diag, s); break;
    case 4: launch_vertex_cr<4>(a, cr, diag, s); break;
    default: evp_set_error("unsupported vertexDegree %d", h->D); return EVP_ERR_ARGUMENT;
    }
    EVP_CUDA(cudaGetLastError());
    return EVP_OK;
}

namespace {
// one warp per tile: lane = cell.  work = solved, or a non-solved cell still carrying stress (it enters the
// divergence, variational.F:1151-1170).  Tiles without work get their contrib rows zeroed once, here.
__global__ void __launch_bounds__(256) k_tile_flags(int nCells, size_t nCp, int M, const uint8_t *__restrict__ nEdges,
                                                    const uint8_t *__restrict__ solveStress, const double2 *__restrict__ sig,
                                                    const double *__restrict__ sig12, double2 *__restrict__ contrib,
                                                    uint8_t *__restrict__ tileWork, size_t nTiles)
{
    const size_t tile = (size_t)blockIdx.x * (blockDim.x / EVP_TILE) + threadIdx.x / EVP_TILE;
    if (tile >= nTiles) return;
    const int lane = threadIdx.x % EVP_TILE;
    const size_t c = tile * EVP_TILE + lane;
    bool work = false;
    if (c < (size_t)nCells) {
        work = solveStress[c] == 1;
        if (!work) {
            const int n = nEdges[c];
            for (int j = 0; j < n; j++) {
                const double2 s = sig[(size_t)j * nCp + c];
                work |= (s.x != 0.0) | (s.y != 0.0) | (sig12[(size_t)j * nCp + c] != 0.0);
            }
        }
    }
    const bool any = __any_sync(0xffffffffu, work);
    if (lane == 0) tileWork[tile] = any ? 1 : 0;
    if (!any)
        for (int j = 0; j < M; j++) contrib[(size_t)j * nCp + c] = make_double2(0.0, 0.0);
}
}  // namespace

namespace {
// ordered inside a block of 256 tiles (warp ballots + a shared prefix), blocks land in atomicAdd order: the list
// keeps the mesh's locality in runs of 256 tiles, and the order has no influence on any result
__global__ void __launch_bounds__(256) k_tile_compact(int nTiles, const uint8_t *__restrict__ tileWork, int *__restrict__ list,
                                                      int *__restrict__ count)
{
    __shared__ int warpSum[8];
    __shared__ int base;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = t < nTiles && tileWork[t] != 0;
    const unsigned ballot = __ballot_sync(0xffffffffu, on);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warpSum[warp] = __popc(ballot);
    __syncthreads();
    if (threadIdx.x == 0) {
        int total = 0;
        for (int w = 0; w < 8; w++) { const int v = warpSum[w]; warpSum[w] = total; total += v; }
        base = total ? atomicAdd(count, total) : 0;
    }
    __syncthreads();
    if (on) list[base + warpSum[warp] + __popc(ballot & ((1u << lane) - 1u))] = t;
}
}  // namespace

namespace {
__global__ void __launch_bounds__(256) k_vblock_flags(int nVerticesSolve, const uint8_t *__restrict__ solveVel,
                                                      uint8_t *__restrict__ vblockWork)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    // bit 1 (boundary-owned, peer-to-peer exchange): the vertex stores its velocity into the neighbours every pass
    const int on = v < nVerticesSolve && solveVel[v] != 0;
    const int any = __syncthreads_or(on);
    if (threadIdx.x == 0) vblockWork[blockIdx.x] = any ? 1 : 0;
}
}  // namespace

int evp_refresh_tile_flags(evp_handle *h, cudaStream_t s)
{
    const size_t nTiles = h->nCp / EVP_TILE;
    const int tilesPerBlock = 256 / EVP_TILE;
    k_tile_flags<<<(unsigned)((nTiles + tilesPerBlock - 1) / tilesPerBlock), 256, 0, s>>>(
        h->nCells, h->nCp, h->M, h->d.nEdges, h->d.solveStress, h->d.sig, h->d.sig12, h->d.contrib, h->d.tileWork, nTiles);
    EVP_CUDA(cudaMemsetAsync(h->d.tileCount, 0, sizeof(int), s));
    const int realTiles = (h->nCells + EVP_TILE - 1) / EVP_TILE;
    if (realTiles) k_tile_compact<<<(realTiles + 255) / 256, 256, 0, s>>>(realTiles, h->d.tileWork, h->d.tileList, h->d.tileCount);
    EVP_CUDA(cudaGetLastError());
    // the same for the vertex pass: blocks of 256 owned vertices without a solved vertex are not launched
    const int vBlocks = (h->nVerticesSolve + 255) / 256;
    EVP_CUDA(cudaMemsetAsync(h->d.vblockCount, 0, sizeof(int), s));
    if (vBlocks) {
        k_vblock_flags<<<vBlocks, 256, 0, s>>>(h->nVerticesSolve, h->d.solveVel, h->d.vblockWork);
        k_tile_compact<<<(vBlocks + 255) / 256, 256, 0, s>>>(vBlocks, h->d.vblockWork, h->d.vblockList, h->d.vblockCount);
    }
    EVP_CUDA(cudaGetLastError());
    int count = 0, vcount = 0;
    EVP_CUDA(cudaMemcpyAsync(&count, h->d.tileCount, sizeof(int), cudaMemcpyDeviceToHost, s));
    EVP_CUDA(cudaMemcpyAsync(&vcount, h->d.vblockCount, sizeof(int), cudaMemcpyDeviceToHost, s));
    EVP_CUDA(cudaStreamSynchronize(s));
    if (count != h->nActiveTiles || vcount != h->nActiveVBlocks) {
        // the grids are baked into the graph nodes: a different number of tiles / vertex blocks needs a new graph
        h->nActiveTiles = count;
        h->nActiveVBlocks = vcount;
        if (h->graphExec) { cudaGraphExecDestroy(h->graphExec); h->graphExec = nullptr; }
        h->graphN = -1;
    }
    return EVP_OK;
}

int evp_enqueue_special_boundaries(evp_handle *h, cudaStream_t s)
{
    if (!h->opt.use_special_boundaries_velocity || h->d.nSB == 0) return EVP_OK;
    const int block = 256, grid = (h->d.nSB + block - 1) / block;
    evp_sb_gather<<<grid, block, 0, s>>>(h->d.nSB, h->d.sbSrc, h->d.sbSign, h->d.uv, h->d.sbTmp);
    evp_sb_scatter<<<grid, block, 0, s>>>(h->d.nSB, h->d.sbDst, h->d.sbTmp, h->d.uv);
    EVP_CUDA(cudaGetLastError());
    return EVP_OK;
}

static inline bool diag_always(const evp_handle *h)
{
    const int cr = h->opt.constitutive_relation_type;
    return cr == EVP_CR_LINEAR || cr == EVP_CR_NONE;
}

// subcycle_velocity_solver (velocity_solver.F:2442-2458): special boundaries, then nSub x
// { internal stress, drag coefficient, velocity solve, halo exchange, special boundaries }.
// Strain, replacementPressure, stressDivergence and oceanStressCoeff are only consumed after the
// loop (velocity_solver.F:3360-3380), so only the last subcycle stores them (DIAG).
// One subcycle after the cell pass: vertex pass + halo exchange.  With the peer-to-peer exchange the vertex kernel
// is the exchange (evp_halo.cu); otherwise pack / ncclSend+ncclRecv / [unpack] follow on the same stream.
int evp_enqueue_vertex_and_halo(evp_handle *h, bool diag, cudaStream_t s)
{
    int rc;
    if ((rc = evp_enqueue_vertex_pass(h, diag, s))) return rc;
    return evp_halo_enqueue(h, s);          // no-op without neighbours or with the peer-to-peer exchange
}

int evp_enqueue_subcycles(evp_handle *h, int nSub, cudaStream_t s)
{
    int rc;
    if ((rc = evp_enqueue_special_boundaries(h, s))) return rc;
    if (nSub > 0 && (rc = evp_halo_begin_run(h, s))) return rc;
    for (int k = 0; k < nSub; k++) {
        const bool diag = (k == nSub - 1) || diag_always(h);
        if ((rc = evp_enqueue_cell_pass(h, diag, s))) return rc;
        if ((rc = evp_enqueue_vertex_and_halo(h, diag, s))) return rc;
        if ((rc = evp_enqueue_special_boundaries(h, s))) return rc;
    }
    if (nSub > 0 && (rc = evp_halo_end_run(h, s))) return rc;
    return EVP_OK;
}

// ---- the persistent whole-loop kernel: eligibility and launch ----
namespace {
template <int M, bool METRIC, int CR, int D>
int persistent_launch(evp_handle *h, const PersistArgs &p, unsigned grid, cudaStream_t s, bool probeOnly)
{
    auto kern = evp_persistent_kernel<M, METRIC, CR, D>;
    constexpr size_t smem = sizeof(CellSmem<M, METRIC, true>);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return -1; }
    int perSM = 0, nSM = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, EVP_TILE * M, smem) != cudaSuccess) { cudaGetLastError(); return -1; }
    if (cudaDeviceGetAttribute(&nSM, cudaDevAttrMultiProcessorCount, h->device) != cudaSuccess) { cudaGetLastError(); return -1; }
    if ((long long)perSM * nSM < (long long)grid) return -1;          // not every tile can be resident at once
    if (probeOnly) return 0;
    void *args[] = {(void *)&p};
    if (cudaLaunchCooperativeKernel((const void *)kern, dim3(grid), dim3(EVP_TILE, M), args, smem, s) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return 0;
}
template <int M, bool METRIC, int CR>
int persistent_d(evp_handle *h, const PersistArgs &p, unsigned grid, cudaStream_t s, bool probe)
{
    return h->D == 3 ? persistent_launch<M, METRIC, CR, 3>(h, p, grid, s, probe) : persistent_launch<M, METRIC, CR, 4>(h, p, grid, s, probe);
}
template <int M, bool METRIC>
int persistent_cr(evp_handle *h, const PersistArgs &p, unsigned grid, cudaStream_t s, bool probe)
{
    return h->opt.constitutive_relation_type == EVP_CR_EVP ? persistent_d<M, METRIC, EVP_CR_EVP>(h, p, grid, s, probe)
                                                           : persistent_d<M, METRIC, EVP_CR_EVP_REVISED>(h, p, grid, s, probe);
}
template <int M>
int persistent_m(evp_handle *h, const PersistArgs &p, unsigned grid, cudaStream_t s, bool probe)
{
    return h->metric ? persistent_cr<M, true>(h, p, grid, s, probe) : persistent_cr<M, false>(h, p, grid, s, probe);
}
}  // namespace

static void fill_cell_args(evp_handle *h, CellArgs &a);
static void fill_vertex_args(evp_handle *h, VertexArgs &a);

// Can (and should) nSub subcycles of this handle run as ONE cooperative launch?  Default configuration only:
// variational operators without vertex averaging, EVP / revised EVP, banded (Wachspress) gradients, one rank, no
// special boundaries -- everything else keeps the two-kernel graph.  Used whenever every tile can have a resident
// block of its own; EVP_B200_PERSISTENT=0 in the environment switches it off (the tests compare both paths).
static bool persistent_configured(evp_handle *h)
{
    const char *e = getenv("EVP_B200_PERSISTENT");
    if (e && e[0] == '0') return false;
    if (!h->useGraph || h->nCells == 0 || h->nVerticesSolve == 0) return false;
    if (h->opt.strain_scheme == EVP_SCHEME_WEAK || h->opt.stress_divergence_scheme == EVP_SCHEME_WEAK) return false;
    if (h->opt.average_variational_strain) return false;
    const int cr = h->opt.constitutive_relation_type;
    if (cr != EVP_CR_EVP && cr != EVP_CR_EVP_REVISED) return false;
    if (!h->d.Gb) return false;
    if (h->opt.use_special_boundaries_velocity && h->d.nSB) return false;
    if (evp_halo_launches(h) || evp_halo_p2p_active(h)) return false;
    if (h->D != 3 && h->D != 4) return false;
    if (h->nVerticesSolve != h->nVertices) return false;        // one rank: every vertex of a local cell is owned
    return true;
}

// 0 = launched (or, with probeOnly, would be), -1 = not eligible
int evp_persistent_run(evp_handle *h, int nSub, cudaStream_t s, bool probeOnly)
{
    if (nSub <= 0 || !persistent_configured(h)) return -1;
    const unsigned grid = (unsigned)((h->nCells + EVP_TILE - 1) / EVP_TILE);
    if (grid > 4096) return -1;                       // far beyond what can be co-resident: skip the occupancy query
    PersistArgs p;
    fill_cell_args(h, p.c);
    fill_vertex_args(h, p.v);
    p.nSub = nSub;
    if (!probeOnly) {
        if (!h->d.contrib2) {      // second buffer of the divergence sums + the barrier words, on first use
            if (evp_dev_alloc(h, (void **)&h->d.contrib2, sizeof(double2) * h->M * h->nCp)) return -1;
            if (evp_dev_alloc(h, (void **)&h->d.gridBar, sizeof(unsigned) * (grid + 64))) return -1;
        }
        if (cudaMemsetAsync(h->d.gridBar, 0, sizeof(unsigned) * (grid + 64), s) != cudaSuccess) { cudaGetLastError(); return -1; }
    }
    p.contrib2 = h->d.contrib2;
    p.arrive = h->d.gridBar + 32;
    p.release = h->d.gridBar;
    switch (h->M) {
    case 4: return persistent_m<4>(h, p, grid, s, probeOnly);
    case 6: return persistent_m<6>(h, p, grid, s, probeOnly);
    case 8: return persistent_m<8>(h, p, grid, s, probeOnly);
    default: return -1;
    }
}

int evp_count_launches(evp_handle *h, int nSub)
{
    if (evp_persistent_run(h, nSub, h->stream, true) == 0) return 1;
    const int sb = (h->opt.use_special_boundaries_velocity && h->d.nSB) ? 2 : 0;
    const bool cells = h->nCells > 0, work = cells && h->nActiveTiles != 0, verts = h->nVerticesSolve > 0;
    const bool vwork = verts && h->nActiveVBlocks != 0;
    int cellKernels;                                   // per subcycle, before the vertex pass
    if (h->opt.strain_scheme == EVP_SCHEME_WEAK) {
        cellKernels = cells ? 1 : 0;                   // k_weak_cells runs over every cell
        if (h->opt.stress_divergence_scheme != EVP_SCHEME_WEAK)   // strain -> vertex, vertex -> stress points, PHASE 2
            cellKernels += (verts ? 1 : 0) + (cells ? 1 : 0) + (work ? 1 : 0);
    } else if (h->opt.average_variational_strain) {
        cellKernels = (work ? 1 : 0) + (verts ? 1 : 0) + (work ? 1 : 0);   // PHASE 1, vertex average, PHASE 2
    } else {
        cellKernels = work ? 1 : 0;
    }
    const bool p2p = evp_halo_p2p_active(h);
    const int perSub = cellKernels + ((vwork || p2p) ? 1 : 0) + evp_halo_launches(h) + sb;
    return sb + nSub * perSub + ((p2p && nSub > 0) ? 2 : 0);      // + evp_halo_begin_run / evp_halo_end_run
}

// Instrumentation for bench.py: average device time of the cell pass, the vertex pass and the rest
// (halo + special boundaries) over nSub subcycles, CUDA events on the launching stream, no graph.
extern "C" int evp_profile_passes(evp_handle *h, int nSub, float *cellMs, float *vertexMs, float *otherMs)
{
    EVP_REQUIRE(h != nullptr && nSub > 0, "bad argument");
    if (!h->haveBasis || !h->haveStep) { evp_set_error("evp_profile_passes before create/update_step"); return EVP_ERR_STATE; }
    EVP_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    cudaEvent_t e[4];
    for (auto &x : e) EVP_CUDA(cudaEventCreate(&x));
    double acc[3] = {0, 0, 0};
    int rc = evp_halo_begin_run(h, s);
    for (int k = 0; k < nSub && rc == EVP_OK; k++) {
        EVP_CUDA(cudaEventRecord(e[0], s));
        if ((rc = evp_enqueue_cell_pass(h, false, s))) break;
        EVP_CUDA(cudaEventRecord(e[1], s));
        if ((rc = evp_enqueue_vertex_and_halo(h, false, s))) break;
        EVP_CUDA(cudaEventRecord(e[2], s));
        if ((rc = evp_enqueue_special_boundaries(h, s))) break;
        EVP_CUDA(cudaEventRecord(e[3], s));
        EVP_CUDA(cudaEventSynchronize(e[3]));
        for (int i = 0; i < 3; i++) {
            float ms = 0.f;
            EVP_CUDA(cudaEventElapsedTime(&ms, e[i], e[i + 1]));
            acc[i] += ms;
        }
    }
    for (auto &x : e) cudaEventDestroy(x);
    if (rc) return rc;
    if ((rc = evp_halo_end_run(h, s))) return rc;
    EVP_CUDA(cudaStreamSynchronize(s));
    if (cellMs) *cellMs = (float)(acc[0] / nSub);
    if (vertexMs) *vertexMs = (float)(acc[1] / nSub);
    if (otherMs) *otherMs = (float)(acc[2] / nSub);
    return EVP_OK;
}
