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
    size_t nCp;
    const uint8_t *__restrict__ nEdges;
    const uint8_t *__restrict__ solveStress;
    const int *__restrict__ voc;
    const double2 *__restrict__ G;
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

template <int M, bool METRIC, int CR, bool DIAG>
__global__ void __launch_bounds__(128) evp_cell_kernel(const CellArgs a)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.nCells) return;
    const size_t nCp = a.nCp;
    const int n = a.nEdges[c];
    const bool solve = a.solveStress[c] == 1;

    double s11[M], s22[M], s12[M], tv[M];

    if (solve) {
        double u[M], v[M];
#pragma unroll
        for (int i = 0; i < M; i++) {
            u[i] = 0.0; v[i] = 0.0; tv[i] = 0.0;
            if (i < n) {
                const int vi = a.voc[(size_t)i * nCp + c];
                const double2 w = a.uv[vi];
                u[i] = w.x; v[i] = w.y;
                if (METRIC) tv[i] = a.tanLat[vi];
            }
        }
        const double P = a.P[c];
#pragma unroll
        for (int j = 0; j < M; j++) {
            s11[j] = 0.0; s22[j] = 0.0; s12[j] = 0.0;
            if (j < n) {
                double e11 = 0.0, e22 = 0.0, e12 = 0.0;
#pragma unroll
                for (int i = 0; i < M; i++) {
                    if (i < n) {
                        const double2 g = a.G[(size_t)(j * M + i) * nCp + c];
                        e11 = e11 + u[i] * g.x;
                        e22 = e22 + v[i] * g.y;
                        e12 = e12 + 0.5 * (u[i] * g.y + v[i] * g.x);
                    }
                }
                // metric terms (variational.F:658-662); tv == 0 when METRIC is off
                e11 = e11 - v[j] * tv[j];
                e12 = e12 + u[j] * tv[j] * 0.5;
                const size_t q = (size_t)j * nCp + c;
                const double2 s = a.sig[q];
                double x11 = s.x, x22 = s.y, x12 = a.sig12[q], rep = 0.0;
                constitutive<CR>(x11, x22, x12, e11, e22, e12, P, rep, a.dte, a.damping);
                if (CR != EVP_CR_NONE) {
                    a.sig[q] = make_double2(x11, x22);
                    a.sig12[q] = x12;
                }
                s11[j] = x11; s22[j] = x22; s12[j] = x12;
                if (DIAG) {
                    a.e11[q] = e11; a.e22[q] = e22; a.e12[q] = e12;
                    if (CR == EVP_CR_EVP || CR == EVP_CR_EVP_REVISED) a.repP[q] = rep;
                }
            }
        }
    } else {
        // cell not solved: the stress is whatever the host left there (zero after
        // init_subcycle_variables, velocity_solver.F:2335-2345) and still enters the divergence.
        bool anyNonZero = false;
#pragma unroll
        for (int j = 0; j < M; j++) {
            s11[j] = 0.0; s22[j] = 0.0; s12[j] = 0.0; tv[j] = 0.0;
            if (j < n) {
                const size_t q = (size_t)j * nCp + c;
                const double2 s = a.sig[q];
                s11[j] = s.x; s22[j] = s.y; s12[j] = a.sig12[q];
                anyNonZero |= (s.x != 0.0) | (s.y != 0.0) | (s12[j] != 0.0);
                if (DIAG && CR == EVP_CR_EVP) a.repP[q] = 0.0;   // variational.F:862
            }
        }
        if (!anyNonZero) {
            // all-zero stress: every partial sum is exactly +0 (0 - 0*S), skip reading the integrals
#pragma unroll
            for (int j = 0; j < M; j++)
                if (j < n) a.contrib[(size_t)j * nCp + c] = make_double2(0.0, 0.0);
            return;
        }
        if (METRIC) {
#pragma unroll
            for (int i = 0; i < M; i++)
                if (i < n) tv[i] = a.tanLat[a.voc[(size_t)i * nCp + c]];
        }
    }

    // per-cell partial sums of the stress divergence for each velocity vertex slot jv
#pragma unroll
    for (int jv = 0; jv < M; jv++) {
        if (jv < n) {
            double cU = 0.0, cV = 0.0;
#pragma unroll
            for (int i = 0; i < M; i++) {
                if (i < n) {
                    const size_t b = (size_t)(jv * M + i) * nCp + c;
                    const double2 S = a.Suv[b];
                    if (METRIC) {
                        const double sm = a.Sm[b];
                        cU = cU - s11[i] * S.x - s12[i] * S.y - s12[i] * sm * tv[jv];
                        cV = cV - s22[i] * S.y - s12[i] * S.x + s11[i] * sm * tv[jv];
                    } else {
                        cU = cU - s11[i] * S.x - s12[i] * S.y;
                        cV = cV - s22[i] * S.y - s12[i] * S.x;
                    }
                }
            }
            a.contrib[(size_t)jv * nCp + c] = make_double2(cU, cV);
        }
    }
}

struct VertexArgs {
    int nVerticesSolve;
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
};

template <int D, int CR, bool DIAG>
__global__ void __launch_bounds__(256) evp_vertex_kernel(const VertexArgs a)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= a.nVerticesSolve) return;
    if (a.solveVel[v] != 1) return;

    double sdU = 0.0, sdV = 0.0;
#pragma unroll
    for (int s = 0; s < D; s++) {
        const int idx = a.gidx[(size_t)s * a.nVp + v];
        double2 cc = make_double2(0.0, 0.0);
        if (idx >= 0) cc = a.contrib[idx];
        sdU = sdU + cc.x;
        sdV = sdV + cc.y;
    }
    const double2 ad = a.areaDen[v];
    sdU = sdU / ad.y;
    sdV = sdV / ad.y;

    const double2 w = a.uv[v];
    double coef = 0.0;
    if (a.useOcean) {
        if (a.oceanType == EVP_OCEAN_QUADRATIC) {
            const double2 o = a.ocnVel[v];
            const double du = o.x - w.x, dv = o.y - w.y;
            coef = kDragio * kRhow * ad.x * sqrt(du * du + dv * dv);
        } else {
            coef = kDragio * kRhow * ad.x;
        }
    }
    if (DIAG) {
        a.sdiv[v] = make_double2(sdU, sdV);
        a.ocoef[v] = coef;
    }
    if (CR == EVP_CR_EVP || CR == EVP_CR_EVP_REVISED) {
        const double2 mf = a.massf[v];
        const double2 air = a.air[v];
        const double2 tilt = a.tilt[v];
        const double2 os = a.ocnStress[v];
        const double sgn = copysign(1.0, mf.y);
        double l11, l22, r1, r2;
        if (CR == EVP_CR_EVP) {
            l11 = mf.x / a.dte + coef * kCosOceanTurningAngle;
            l22 = mf.x / a.dte + coef * kCosOceanTurningAngle;
            r1 = sdU + air.x + tilt.x + coef * os.x + (mf.x * w.x) / a.dte;
            r2 = sdV + air.y + tilt.y + coef * os.y + (mf.x * w.y) / a.dte;
        } else {
            const double2 w0 = a.uvInit[v];
            l11 = (a.beta + 1.0) * (mf.x / a.dtDyn) + coef * kCosOceanTurningAngle;
            l22 = (a.beta + 1.0) * (mf.x / a.dtDyn) + coef * kCosOceanTurningAngle;
            r1 = sdU + air.x + tilt.x + coef * os.x + (mf.x * (a.beta * w.x + w0.x)) / a.dtDyn;
            r2 = sdV + air.y + tilt.y + coef * os.y + (mf.x * (a.beta * w.y + w0.y)) / a.dtDyn;
        }
        const double l12 = -mf.y - coef * kSinOceanTurningAngle * sgn;
        const double l21 = mf.y + coef * kSinOceanTurningAngle * sgn;
        const double den = l11 * l22 - l12 * l21;
        a.uv[v] = make_double2((l22 * r1 - l12 * r2) / den, (l11 * r2 - l21 * r1) / den);
    }
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

template <int M, bool METRIC, int CR>
int launch_cell_d(const CellArgs &a, bool diag, cudaStream_t s)
{
    const int block = 128;
    const int grid = (a.nCells + block - 1) / block;
    if (diag) evp_cell_kernel<M, METRIC, CR, true><<<grid, block, 0, s>>>(a);
    else      evp_cell_kernel<M, METRIC, CR, false><<<grid, block, 0, s>>>(a);
    return 0;
}
template <int M, bool METRIC>
int launch_cell_cr(const CellArgs &a, int cr, bool diag, cudaStream_t s)
{
    switch (cr) {
    case EVP_CR_EVP: return launch_cell_d<M, METRIC, EVP_CR_EVP>(a, diag, s);
    case EVP_CR_EVP_REVISED: return launch_cell_d<M, METRIC, EVP_CR_EVP_REVISED>(a, diag, s);
    case EVP_CR_LINEAR: return launch_cell_d<M, METRIC, EVP_CR_LINEAR>(a, diag, s);
    default: return launch_cell_d<M, METRIC, EVP_CR_NONE>(a, diag, s);
    }
}
template <int M>
int launch_cell_m(const CellArgs &a, bool metric, int cr, bool diag, cudaStream_t s)
{
    return metric ? launch_cell_cr<M, true>(a, cr, diag, s) : launch_cell_cr<M, false>(a, cr, diag, s);
}

template <int D, int CR>
int launch_vertex_d(const VertexArgs &a, bool diag, cudaStream_t s)
{
    const int block = 256;
    const int grid = (a.nVerticesSolve + block - 1) / block;
    if (diag) evp_vertex_kernel<D, CR, true><<<grid, block, 0, s>>>(a);
    else      evp_vertex_kernel<D, CR, false><<<grid, block, 0, s>>>(a);
    return 0;
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

int evp_enqueue_cell_pass(evp_handle *h, bool diag, cudaStream_t s)
{
    if (h->nCells == 0) return EVP_OK;
    CellArgs a;
    a.nCells = h->nCells; a.nCp = h->nCp;
    a.nEdges = h->d.nEdges; a.solveStress = h->d.solveStress; a.voc = h->d.voc;
    a.G = h->d.G; a.Suv = h->d.Suv; a.Sm = h->d.Sm;
    a.uv = h->d.uv; a.tanLat = h->d.tanLat; a.P = h->d.P;
    a.sig = h->d.sig; a.sig12 = h->d.sig12; a.contrib = h->d.contrib;
    a.e11 = h->d.e11; a.e22 = h->d.e22; a.e12 = h->d.e12; a.repP = h->d.repP;
    a.dte = h->opt.elasticTimeStep; a.damping = h->opt.dampingTimescale;
    const int cr = h->opt.constitutive_relation_type;
    switch (h->M) {
    case 4: launch_cell_m<4>(a, h->metric, cr, diag, s); break;
    case 6: launch_cell_m<6>(a, h->metric, cr, diag, s); break;
    case 7:
    case 8: launch_cell_m<8>(a, h->metric, cr, diag, s); break;
    default: evp_set_error("unsupported maxEdges %d", h->M); return EVP_ERR_ARGUMENT;
    }
    EVP_CUDA(cudaGetLastError());
    return EVP_OK;
}

int evp_enqueue_vertex_pass(evp_handle *h, bool diag, cudaStream_t s)
{
    if (h->nVerticesSolve == 0) return EVP_OK;
    VertexArgs a;
    a.nVerticesSolve = h->nVerticesSolve; a.nVp = h->nVp;
    a.solveVel = h->d.solveVel; a.gidx = h->d.gidx; a.contrib = h->d.contrib;
    a.areaDen = h->d.areaDen; a.massf = h->d.massf; a.air = h->d.air; a.tilt = h->d.tilt;
    a.ocnStress = h->d.ocnStress; a.ocnVel = h->d.ocnVel; a.uvInit = h->d.uvInit;
    a.uv = h->d.uv; a.sdiv = h->d.sdiv; a.ocoef = h->d.ocoef;
    a.dte = h->opt.elasticTimeStep; a.dtDyn = h->opt.dynamicsTimeStep;
    a.beta = h->opt.numericalInertiaCoefficient;
    a.useOcean = h->opt.use_ocean_stress; a.oceanType = h->opt.ocean_stress_type;
    const int cr = h->opt.constitutive_relation_type;
    switch (h->D) {
    case 3: launch_vertex_cr<3>(a, cr, diag, s); break;
    case 4: launch_vertex_cr<4>(a, cr, diag, s); break;
    default: evp_set_error("unsupported vertexDegree %d", h->D); return EVP_ERR_ARGUMENT;
    }
    EVP_CUDA(cudaGetLastError());
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
int evp_enqueue_subcycles(evp_handle *h, int nSub, cudaStream_t s)
{
    int rc;
    if ((rc = evp_enqueue_special_boundaries(h, s))) return rc;
    for (int k = 0; k < nSub; k++) {
        const bool diag = (k == nSub - 1) || diag_always(h);
        if ((rc = evp_enqueue_cell_pass(h, diag, s))) return rc;
        if ((rc = evp_enqueue_vertex_pass(h, diag, s))) return rc;
        if ((rc = evp_halo_enqueue(h, s))) return rc;
        if ((rc = evp_enqueue_special_boundaries(h, s))) return rc;
    }
    return EVP_OK;
}

int evp_count_launches(evp_handle *h, int nSub)
{
    const int sb = (h->opt.use_special_boundaries_velocity && h->d.nSB) ? 2 : 0;
    const int perSub = (h->nCells ? 1 : 0) + (h->nVerticesSolve ? 1 : 0) + evp_halo_launches(h) + sb;
    return sb + nSub * perSub;
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
    int rc = EVP_OK;
    for (int k = 0; k < nSub && rc == EVP_OK; k++) {
        EVP_CUDA(cudaEventRecord(e[0], s));
        if ((rc = evp_enqueue_cell_pass(h, false, s))) break;
        EVP_CUDA(cudaEventRecord(e[1], s));
        if ((rc = evp_enqueue_vertex_pass(h, false, s))) break;
        EVP_CUDA(cudaEventRecord(e[2], s));
        if ((rc = evp_halo_enqueue(h, s))) break;
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
    if (cellMs) *cellMs = (float)(acc[0] / nSub);
    if (vertexMs) *vertexMs = (float)(acc[1] / nSub);
    if (otherMs) *otherMs = (float)(acc[2] / nSub);
    return EVP_OK;
}
