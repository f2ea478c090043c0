// evp_prepost.cu -- the steps on either side of the subcycle, on the device:
//   evp_pre_subcycle  = velocity_solver_pre_subcycle  (reference: src/shared/mpas_seaice_velocity_solver.F:613-671)
//   evp_post_subcycle = velocity_solver_post_subcycle (velocity_solver.F:3360-3380)
// One dynamics step then moves a dozen CELL fields in and a handful of fields out, and u, v, the stresses
// and solveVelocityPrevious never leave the device.  Built --fmad=false like the subcycle kernels: the
// expressions below keep the reference's operation order, so results are bit-identical to the oracle.
// What is NOT here on purpose: exp() of the Hibler strength and the category sums (host, see the header).
#include <math.h>
#include <algorithm>
#include "evp_internal.cuh"

namespace {

// velocity_solver.F:61-65
constexpr double kSinOceanTurningAngle = 0.0;
constexpr double kCosOceanTurningAngle = 1.0;
constexpr double kAreaMinimum = 0.001;
constexpr double kMassMinimum = 0.01;
// mpas_seaice_constants.F:43-92, ice_constants_colpkg.F90:22-63
constexpr double kGravity = 9.80616;
constexpr double kDragio = 0.00536;
constexpr double kRhow = 1026.0;
constexpr double kPuny = 1.0e-11;
constexpr double kEccentricitySquared = 2.0 * 2.0;
// the junk cell every invalid neighbour points to (src/shared/mpas_seaice_initialize.F:214-234)
constexpr double kJunkArea = -1.0e34;

struct PreArgs {
    int nCells, nVerticesSolve, nVertices, M, D;
    size_t nCp, nVp;
    // options
    int useAir, constantAir, useOcean, tiltMode /* 0 none, 1 geostrophic, 2 ssh gradient */, calcMasks, coldStart, cr;
    // static
    const uint8_t *__restrict__ nEdges;
    const int *__restrict__ coc;
    const int *__restrict__ cov;
    const uint8_t *__restrict__ vflags;
    const double *__restrict__ areaCell;
    const double *__restrict__ fVertex;
    // staged cell inputs (plain (nCells) arrays)
    const double *__restrict__ areaInit, *__restrict__ areaNow, *__restrict__ mass, *__restrict__ Pin;
    const double *__restrict__ uOcn, *__restrict__ vOcn;
    const double *__restrict__ airU, *__restrict__ airV, *__restrict__ uAir, *__restrict__ vAir, *__restrict__ rhoAir;
    const double *__restrict__ tiltU, *__restrict__ tiltV;
    const int *__restrict__ landIce;
    const int *__restrict__ ssIn, *__restrict__ svIn;
    // device state / outputs
    uint8_t *__restrict__ solveStress, *__restrict__ solveVel, *__restrict__ solveVelPrev;
    double *__restrict__ P;
    double2 *__restrict__ airCell;
    double2 *__restrict__ sig;
    double *__restrict__ sig12;
    double2 *__restrict__ sigW;          // weak stress divergence scheme: stress11/22Weak, stress12Weak (else nullptr)
    double *__restrict__ sigW12;
    double2 *__restrict__ uv, *__restrict__ uvInit, *__restrict__ areaDen, *__restrict__ massf, *__restrict__ air,
        *__restrict__ tilt, *__restrict__ ocnStress, *__restrict__ ocnVel;
};

__device__ __forceinline__ bool enough_ice(const PreArgs &a, int c)
{
    return a.areaInit[c] > kAreaMinimum && a.mass[c] > kMassMinimum && (a.landIce == nullptr || a.landIce[c] == 0);
}

// aggregate_mass_and_area (velocity_solver.F:685-752) + the Hibler strength before its mask (:1419-1436).
// cat arrays: (nCategories, nCells), category fastest.  The sums start from 0 and run in category order like sum().
constexpr double kDensityIce = 917.0, kDensitySnow = 330.0;                  // ice_constants_colpkg.F90
constexpr double kHiblerP = 2.75e4, kHiblerC = 20.0;                         // mpas_seaice_constants.F
__global__ void __launch_bounds__(256) k_aggregate(int nCells, int nCat, const double *__restrict__ aCat,
                                                   const double *__restrict__ viCat, const double *__restrict__ vsCat,
                                                   double *__restrict__ area, double *__restrict__ volIce,
                                                   double *__restrict__ volSnow, double *__restrict__ mass,
                                                   double *__restrict__ P, int hibler)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    double a = 0.0, vi = 0.0, vs = 0.0;
    for (int k = 0; k < nCat; k++) {
        a = a + aCat[(size_t)c * nCat + k];
        vi = vi + viCat[(size_t)c * nCat + k];
        vs = vs + vsCat[(size_t)c * nCat + k];
    }
    area[c] = a;
    volIce[c] = vi;
    volSnow[c] = vs;
    mass[c] = vi * kDensityIce + vs * kDensitySnow;
    if (hibler) P[c] = kHiblerP * vi * exp(-kHiblerC * (1.0 - a));
}

// cells: stress_calculation_mask (:961-1059), ice_strength mask (:1419-1436), air stress at cells
// (:1560-1580, constant_air_stress :1716-1723), stress reset of init_subcycle_variables (:2335-2345)
__global__ void __launch_bounds__(256) k_pre_cells(const PreArgs a)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.nCells) return;
    int solve;
    if (a.calcMasks) {
        solve = 0;
        if (enough_ice(a, c)) {
            solve = 1;
        } else {
            const int n = a.nEdges[c];
            for (int k = 0; k < n; k++) {
                const int nb = a.coc[(size_t)k * a.nCp + c];
                if (nb >= 0 && enough_ice(a, nb)) { solve = 1; break; }
            }
        }
    } else {
        solve = a.ssIn[c] == 1;
    }
    a.solveStress[c] = (uint8_t)solve;
    a.P[c] = solve ? a.Pin[c] : 0.0;

    double2 air = make_double2(0.0, 0.0);
    if (a.useAir) {
        if (a.constantAir) {
            const double airStressCoeff = 0.0012;
            const double ua = a.uAir[c], va = a.vAir[c];
            const double windSpeed = sqrt(ua * ua + va * va);
            air.x = a.rhoAir[c] * windSpeed * airStressCoeff * ua * a.areaNow[c];
            air.y = a.rhoAir[c] * windSpeed * airStressCoeff * va * a.areaNow[c];
        } else {
            air = make_double2(a.airU[c], a.airV[c]);
        }
    }
    a.airCell[c] = air;

    if (!solve || a.coldStart) {
        if (a.sigW) {                     // init_subcycle_variables, weak branch (:2350-2365)
            a.sigW[c] = make_double2(0.0, 0.0);
            a.sigW12[c] = 0.0;
        } else {
            for (int j = 0; j < a.M; j++) {
                a.sig[(size_t)j * a.nCp + c] = make_double2(0.0, 0.0);
                a.sig12[(size_t)j * a.nCp + c] = 0.0;
            }
        }
    }
}

// seaice_interpolate_cell_to_vertex, cell-area weights, no validity test (mesh.F:2835-2851): an invalid
// neighbour is the junk cell (area -1e34, value 0 here), which only reaches non-interior vertices
struct C2V {
    int c[4];
    double w[4];
    int D;
    __device__ __forceinline__ double operator()(const double *__restrict__ f) const
    {
        double acc = 0.0, tot = 0.0;
        for (int k = 0; k < D; k++) {
            acc = acc + w[k] * (c[k] >= 0 ? f[c[k]] : 0.0);
            tot = tot + w[k];
        }
        return acc / tot;
    }
    __device__ __forceinline__ double2 operator()(const double2 *__restrict__ f) const
    {
        double ax = 0.0, ay = 0.0, tot = 0.0;
        for (int k = 0; k < D; k++) {
            const double2 x = c[k] >= 0 ? f[c[k]] : make_double2(0.0, 0.0);
            ax = ax + w[k] * x.x;
            ay = ay + w[k] * x.y;
            tot = tot + w[k];
        }
        return make_double2(ax / tot, ay / tot);
    }
};

// owned vertices: the interpolations and vertex loops of calculation_masks (:860-947), new_ice_velocities
// (:1236-1279), air_stress (:1635-1650), coriolis_force_coefficient (:1775-1783), ocean_stress (:1846-1878),
// surface_tilt (:1941-2213) and init_subcycle_variables (:2287-2310)
__global__ void __launch_bounds__(256) k_pre_vertices(const PreArgs a)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= a.nVertices) return;
    if (v >= a.nVerticesSolve) {          // halo vertices: masks 0 here, (u,v) arrive by the halo exchange
        a.solveVel[v] = 0;
        a.solveVelPrev[v] = 0;
        return;
    }
    C2V c2v;
    c2v.D = a.D;
    for (int k = 0; k < a.D; k++) {
        c2v.c[k] = a.cov[(size_t)k * a.nVp + v];
        c2v.w[k] = c2v.c[k] >= 0 ? a.areaCell[c2v.c[k]] : kJunkArea;
    }
    const double areaV = c2v(a.areaInit);
    const double massV = c2v(a.mass);
    const double uo = c2v(a.uOcn), vo = c2v(a.vOcn);
    const double2 airV = c2v(a.airCell);
    const double f = a.fVertex[v];

    int solve;
    if (a.calcMasks) {
        const uint8_t fl = a.vflags[v];
        solve = ((fl & 1) != 0) && ((fl & 2) == 0) && areaV > kAreaMinimum && massV > kMassMinimum;
    } else {
        solve = a.svIn[v] == 1;
    }

    double2 w = a.coldStart ? make_double2(0.0, 0.0) : a.uv[v];
    const int prev = a.coldStart == EVP_START_FROM_REST ? solve : (a.coldStart == EVP_START_FIRST_STEP ? 0 : a.solveVelPrev[v]);
    if (solve) {
        if (prev == 0) w = make_double2(uo, vo);
    } else {
        w = make_double2(0.0, 0.0);
    }
    a.solveVelPrev[v] = (uint8_t)solve;
    a.uvInit[v] = w;
    a.uv[v] = w;            // init_subcycle_variables changes nothing more: w is already 0 where not solved
    a.solveVel[v] = (uint8_t)solve;

    double2 ad = a.areaDen[v];
    ad.x = areaV;
    a.areaDen[v] = ad;
    a.massf[v] = make_double2(massV, massV * f);
    a.air[v] = airV;
    a.ocnVel[v] = make_double2(uo, vo);

    const double sgn = copysign(1.0, f);
    double2 os = make_double2(0.0, 0.0);
    if (a.useOcean && solve) {
        os.x = uo * kCosOceanTurningAngle - vo * kSinOceanTurningAngle * sgn;
        os.y = uo * kSinOceanTurningAngle * sgn + vo * kCosOceanTurningAngle;
    }
    a.ocnStress[v] = os;

    double2 t = make_double2(0.0, 0.0);
    if (a.tiltMode == 1) {
        if (solve) { t.x = -f * massV * vo; t.y = f * massV * uo; }
    } else if (a.tiltMode == 2) {
        const double tu = c2v(a.tiltU), tv = c2v(a.tiltV);
        if (solve) { t.x = -kGravity * massV * tu; t.y = -kGravity * massV * tv; }
    }
    a.tilt[v] = t;
}

// ---- post-subcycle --------------------------------------------------------------------------------
struct PostArgs {
    int nCells, nCellsSolve, nVerticesSolve, M;
    size_t nCp, nVp;
    int useOcean, oceanType;
    const uint8_t *__restrict__ nEdges, *__restrict__ solveStress, *__restrict__ solveVel, *__restrict__ vflags;
    const int *__restrict__ voc;
    const double *__restrict__ e11, *__restrict__ e22, *__restrict__ e12, *__restrict__ repP, *__restrict__ sig12;
    const double2 *__restrict__ sig;
    const double2 *__restrict__ uv, *__restrict__ ocnVel, *__restrict__ areaDen;
    const double *__restrict__ fVertex, *__restrict__ areaTri;
    double *__restrict__ div, *__restrict__ shear, *__restrict__ ridgeConv, *__restrict__ ridgeShear;
    double2 *__restrict__ principal;        // [M][nCp] (principalStress1Var, principalStress2Var)
    double2 *__restrict__ osFinal;          // [nVp]
    double *__restrict__ ocoef;
    double *__restrict__ oscU, *__restrict__ oscV;
};

// seaice_final_divergence_shear_variational (variational.F:1198-1330), unit change :1324-1325
__global__ void __launch_bounds__(256) k_post_div_shear(const PostArgs a)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.nCells) return;
    double divergence = 0.0, shear = 0.0, rc = 0.0, rs = 0.0;
    if (a.solveStress[c] == 1) {
        double dSum = 0.0, tSum = 0.0, sSum = 0.0, DeltaAverage = 0.0;
        const int n = a.nEdges[c];
        for (int j = 0; j < n; j++) {
            const size_t q = (size_t)j * a.nCp + c;
            const double e11 = a.e11[q], e22 = a.e22[q], e12 = a.e12[q];
            const double sd = e11 + e22;
            const double st = e11 - e22;
            const double ss = e12 * 2.0;
            const double Delta = sqrt(sd * sd + (st * st + ss * ss) / kEccentricitySquared);
            dSum = dSum + sd;
            tSum = tSum + st;
            sSum = sSum + ss;
            DeltaAverage = DeltaAverage + Delta;
        }
        divergence = dSum / (double)n;
        shear = sqrt(tSum * tSum + sSum * sSum) / (double)n;
        DeltaAverage = DeltaAverage / (double)n;
        rc = -fmin(divergence, 0.0);
        rs = 0.5 * (DeltaAverage - fabs(divergence));
    }
    a.div[c] = divergence * 100.0 * 86400.0;
    a.shear[c] = shear * 100.0 * 86400.0;
    a.ridgeConv[c] = rc;
    a.ridgeShear[c] = rs;
}

// principal_stresses (velocity_solver.F:3565-3610) at every stress point of the owned cells (:3520-3540)
__global__ void __launch_bounds__(256) k_post_principal(const PostArgs a)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.nCellsSolve) return;
    const int n = a.nEdges[c];
    for (int j = 0; j < a.M; j++) {
        const size_t q = (size_t)j * a.nCp + c;
        double2 p = make_double2(0.0, 0.0);
        if (j < n) {
            const double rep = a.repP[q];
            if (rep > kPuny) {
                const double2 s = a.sig[q];
                const double s12 = a.sig12[q];
                const double sqrtContents = (s.x + s.y) * (s.x + s.y) - 4.0 * s.x * s.y + 4.0 * (s12 * s12);
                const double p1 = 0.5 * (s.x + s.y) + 0.5 * sqrt(sqrtContents);
                const double p2 = 0.5 * (s.x + s.y) - 0.5 * sqrt(sqrtContents);
                p = make_double2(p1 / rep, p2 / rep);
            } else {
                p = make_double2(1.0e30, 1.0e30);
            }
        }
        a.principal[q] = p;
    }
}

// seaice_final_divergence_shear_weak (weak.F:651-751) over the owned cells.  As written in the reference the
// whole Delta work array is assigned inside the loop (:729), so ridgeShear sees the Delta of the LAST owned cell.
__global__ void __launch_bounds__(256) k_post_div_shear_weak(int nCellsSolve, const double *__restrict__ e11,
                                                             const double *__restrict__ e22, const double *__restrict__ e12,
                                                             double *__restrict__ div, double *__restrict__ shear,
                                                             double *__restrict__ ridgeConv, double *__restrict__ ridgeShear)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCellsSolve) return;
    const int last = nCellsSolve - 1;
    const double ld = e11[last] + e22[last], lt = e11[last] - e22[last], ls = e12[last] * 2.0;
    const double Delta = sqrt(ld * ld + (lt * lt + ls * ls) / kEccentricitySquared);
    const double sd = e11[c] + e22[c], st = e11[c] - e22[c], ss = e12[c] * 2.0;
    div[c] = sd;
    shear[c] = sqrt(st * st + ss * ss);
    ridgeConv[c] = -fmin(sd, 0.0);
    ridgeShear[c] = 0.5 * (Delta - fabs(sd));
}

// principal_stresses (velocity_solver.F:3565-3610) at the single stress point of every owned cell (:3500-3515)
__global__ void __launch_bounds__(256) k_post_principal_weak(int nCellsSolve, const double2 *__restrict__ sigW,
                                                             const double *__restrict__ sigW12,
                                                             const double *__restrict__ repPW,
                                                             double *__restrict__ p1, double *__restrict__ p2)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCellsSolve) return;
    const double rep = repPW[c];
    if (rep > kPuny) {
        const double2 s = sigW[c];
        const double s12 = sigW12[c];
        const double sqrtContents = (s.x + s.y) * (s.x + s.y) - 4.0 * s.x * s.y + 4.0 * (s12 * s12);
        p1[c] = (0.5 * (s.x + s.y) + 0.5 * sqrt(sqrtContents)) / rep;
        p2[c] = (0.5 * (s.x + s.y) - 0.5 * sqrt(sqrtContents)) / rep;
    } else {
        p1[c] = 1.0e30;
        p2[c] = 1.0e30;
    }
}

// ocean_stress_final, vertex part (:3690-3720) after ocean_stress_coefficient (:3046-3075) on the final (u,v)
__global__ void __launch_bounds__(256) k_post_ocean_vertices(const PostArgs a)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= a.nVerticesSolve) return;
    double2 os = make_double2(0.0, 0.0);
    if (a.useOcean) {
        if (a.solveVel[v] & 1) {
            const double2 w = a.uv[v], o = a.ocnVel[v];
            const double areaV = a.areaDen[v].x;
            double coef;
            if (a.oceanType == EVP_OCEAN_QUADRATIC) {
                const double du = o.x - w.x, dv = o.y - w.y;
                coef = kDragio * kRhow * areaV * sqrt(du * du + dv * dv);
            } else {
                coef = kDragio * kRhow * areaV;
            }
            a.ocoef[v] = coef;
            const double sgn = copysign(1.0, a.fVertex[v]);
            os.x = coef * ((o.x - w.x) * kCosOceanTurningAngle - (o.y - w.y) * kSinOceanTurningAngle * sgn);
            os.y = coef * ((o.y - w.y) * kCosOceanTurningAngle + (o.x - w.x) * kSinOceanTurningAngle * sgn);
            os.x = os.x / areaV;
            os.y = os.y / areaV;
        }
    } else {
        a.ocoef[v] = 0.0;
    }
    a.osFinal[v] = os;
}

// seaice_interpolate_vertex_to_cell (mesh.F:2958-2971) of the two ocean stress components
__global__ void __launch_bounds__(256) k_post_ocean_cells(const PostArgs a)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.nCellsSolve) return;
    double cu = 0.0, cv = 0.0;
    if (a.useOcean) {
        double totalArea = 0.0;
        const int n = a.nEdges[c];
        for (int j = 0; j < n; j++) {
            const int v = a.voc[(size_t)j * a.nCp + c];
            const double ri = (double)(a.vflags[v] & 1);
            const double at = a.areaTri[v];
            const double2 os = a.osFinal[v];
            cu = cu + at * os.x * ri;
            cv = cv + at * os.y * ri;
            totalArea = totalArea + at * ri;
        }
        if (totalArea > 0.0) { cu = cu / totalArea; cv = cv / totalArea; }
    }
    a.oscU[c] = cu;
    a.oscV[c] = cv;
}

// the last loop of ocean_stress_final (:3790-3800): back to a stress per unit grid area
__global__ void __launch_bounds__(256) k_post_ocean_rescale(const PostArgs a)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= a.nVerticesSolve) return;
    if (a.useOcean && (a.solveVel[v] & 1)) {
        double2 os = a.osFinal[v];
        const double areaV = a.areaDen[v].x;
        os.x = os.x * areaV;
        os.y = os.y * areaV;
        a.osFinal[v] = os;
    }
}

__global__ void k_int_in_coc(const int *__restrict__ src, int *__restrict__ dst, const uint8_t *__restrict__ nEdges, int Mh,
                             size_t count, size_t c0, size_t stride, int nCells)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= count) return;
    const int n = nEdges[c0 + c];
    for (int r = 0; r < Mh; r++) {
        int nb = src[(size_t)Mh * c + r] - 1;
        if (r >= n || nb < 0 || nb >= nCells) nb = -1;
        dst[(size_t)r * stride + c0 + c] = nb;
    }
}
__global__ void k_vflags(const int *__restrict__ interior, const int *__restrict__ landIce, uint8_t *__restrict__ out, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = (uint8_t)((interior[i] == 1 ? 1 : 0) | ((landIce && landIce[i] != 0) ? 2 : 0));
}
__global__ void k_u8_to_int(const uint8_t *__restrict__ src, int *__restrict__ dst, size_t n, int mask)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] & mask;
}
__global__ void k_int_to_u8(const int *__restrict__ src, uint8_t *__restrict__ dst, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (uint8_t)(src[i] == 1);
}
__global__ void k_split(const double2 *__restrict__ src, double *__restrict__ a, double *__restrict__ b, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 w = src[i];
    if (a) a[i] = w.x;
    if (b) b[i] = w.y;
}
__global__ void k_join(const double *__restrict__ a, const double *__restrict__ b, double2 *__restrict__ dst, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = make_double2(a[i], b[i]);
}

}  // namespace

extern "C" int evp_set_mesh_ext(evp_handle *h, const evp_mesh_ext *e)
{
    EVP_REQUIRE(h != nullptr && e != nullptr, "handle/ext is NULL");
    EVP_REQUIRE(e->cellsOnCell && e->interiorVertex && e->areaCell && e->areaTriangle && e->fVertex,
                "cellsOnCell, interiorVertex, areaCell, areaTriangle and fVertex must not be NULL");
    EVP_CUDA(cudaSetDevice(h->device));
    evp_dev &d = h->d;
    const size_t nC = h->nCells, nV = h->nVertices, nCp = h->nCp, nVp = h->nVp;
    int rc;
    if (!d.coc) {
        if ((rc = evp_dev_alloc(h, (void **)&d.coc, sizeof(int) * h->M * nCp))) return rc;
        if ((rc = evp_dev_alloc(h, (void **)&d.vflags, nVp))) return rc;
        if ((rc = evp_dev_alloc(h, (void **)&d.areaCell, sizeof(double) * nCp))) return rc;
        if ((rc = evp_dev_alloc(h, (void **)&d.areaTri, sizeof(double) * nVp))) return rc;
        if ((rc = evp_dev_alloc(h, (void **)&d.fVertex, sizeof(double) * nVp))) return rc;
        if ((rc = evp_dev_alloc(h, (void **)&d.airCell, sizeof(double2) * nCp))) return rc;
        if ((rc = evp_dev_alloc(h, (void **)&d.osFinal, sizeof(double2) * nVp))) return rc;
    }
    cudaStream_t s = h->stream;
    EVP_CUDA(cudaStreamSynchronize(s));
    EVP_CUDA(cudaMemsetAsync(d.coc, 0xff, sizeof(int) * h->M * nCp, s));
    EVP_CUDA(cudaMemsetAsync(d.vflags, 0, nVp, s));
    EVP_CUDA(cudaMemsetAsync(d.areaCell, 0, sizeof(double) * nCp, s));
    EVP_CUDA(cudaMemsetAsync(d.areaTri, 0, sizeof(double) * nVp, s));
    EVP_CUDA(cudaMemsetAsync(d.fVertex, 0, sizeof(double) * nVp, s));
    EVP_CUDA(cudaMemsetAsync(d.airCell, 0, sizeof(double2) * nCp, s));
    EVP_CUDA(cudaMemsetAsync(d.osFinal, 0, sizeof(double2) * nVp, s));
    const bool savedPin = h->pinHost;
    h->pinHost = false;                       // one-shot transfers never page-lock caller memory
    rc = EVP_OK;
    if (nC) {
        const int Mh = h->Mh;
        const size_t chunk = std::max<size_t>(1, std::min(nC, (h->d.stageBytes - 4096) / ((size_t)Mh * 4)));
        for (size_t c0 = 0; c0 < nC && !rc; c0 += chunk) {
            const size_t cnt = std::min(chunk, nC - c0);
            rc = evp_h2d(h, d.stage, e->cellsOnCell + c0 * Mh, cnt * Mh * 4);
            if (!rc) k_int_in_coc<<<grid_for(cnt, 128), 128, 0, s>>>((const int *)d.stage, d.coc, d.nEdges, Mh, cnt, c0, nCp, (int)nC);
        }
        if (!rc) rc = evp_h2d(h, d.areaCell, e->areaCell, nC * 8);
    }
    if (nV && !rc) {
        Stage st{(char *)d.stage, d.stageBytes, 0};
        EVP_CUDA(cudaStreamSynchronize(s));
        int *ri = (int *)st.take(nV * 4), *rl = e->landIceMaskVertex ? (int *)st.take(nV * 4) : nullptr;
        rc = evp_h2d(h, ri, e->interiorVertex, nV * 4);
        if (!rc && rl) rc = evp_h2d(h, rl, e->landIceMaskVertex, nV * 4);
        if (!rc) k_vflags<<<grid_for(nV, 256), 256, 0, s>>>(ri, rl, d.vflags, nV);
        if (!rc) rc = evp_h2d(h, d.areaTri, e->areaTriangle, nV * 8);
        if (!rc) rc = evp_h2d(h, d.fVertex, e->fVertex, nV * 8);
    }
    h->pinHost = savedPin;
    if (rc) return rc;
    EVP_CUDA(cudaGetLastError());
    EVP_CUDA(cudaStreamSynchronize(s));
    h->haveExt = true;
    return EVP_OK;
}

extern "C" int evp_set_state(evp_handle *h, const double *u, const double *v, const double *s11, const double *s22,
                             const double *s12, const int *svPrev)
{
    EVP_REQUIRE(h != nullptr, "handle is NULL");
    EVP_REQUIRE((u == nullptr) == (v == nullptr), "uVelocity and vVelocity go together");
    EVP_REQUIRE((s11 == nullptr) == (s22 == nullptr) && (s11 == nullptr) == (s12 == nullptr), "the three stresses go together");
    EVP_CUDA(cudaSetDevice(h->device));
    evp_dev &d = h->d;
    const size_t nC = h->nCells, nV = h->nVertices;
    cudaStream_t s = h->stream;
    EVP_CUDA(cudaStreamSynchronize(s));
    Stage st{(char *)d.stage, d.stageBytes, 0};
    int rc;
    if (u && nV) {
        double *ra = (double *)st.take(nV * 8), *rb = (double *)st.take(nV * 8);
        EVP_REQUIRE(ra && rb, "staging area exhausted");
        if ((rc = evp_h2d(h, ra, u, nV * 8))) return rc;
        if ((rc = evp_h2d(h, rb, v, nV * 8))) return rc;
        k_join<<<grid_for(nV, 256), 256, 0, s>>>(ra, rb, d.uv, nV);
    }
    if (svPrev && nV) {
        int *ri = (int *)st.take(nV * 4);
        EVP_REQUIRE(ri, "staging area exhausted");
        if ((rc = evp_h2d(h, ri, svPrev, nV * 4))) return rc;
        k_int_to_u8<<<grid_for(nV, 256), 256, 0, s>>>(ri, d.solveVelPrev, nV);
    }
    EVP_CUDA(cudaGetLastError());
    EVP_CUDA(cudaStreamSynchronize(s));
    if (s11 && nC) {
        if ((rc = evp_upload_rows(h, s11, (double *)d.sig, 1, 2, 0))) return rc;
        if ((rc = evp_upload_rows(h, s22, (double *)d.sig, 1, 2, 1))) return rc;
        if ((rc = evp_upload_rows(h, s12, d.sig12, 1, 1, 0))) return rc;
        if ((rc = evp_refresh_tile_flags(h, s))) return rc;
    }
    EVP_CUDA(cudaStreamSynchronize(s));
    return EVP_OK;
}

extern "C" int evp_pre_subcycle(evp_handle *h, const evp_pre_fields *f, const evp_pre_options *o)
{
    EVP_REQUIRE(h != nullptr && f != nullptr && o != nullptr, "NULL argument");
    if (!h->haveExt) { evp_set_error("evp_pre_subcycle needs evp_set_mesh_ext first"); return EVP_ERR_STATE; }
    EVP_REQUIRE(f->uOceanVelocity && f->vOceanVelocity, "u/vOceanVelocity must not be NULL");
    if (!f->iceAreaCell || !f->totalMassCell) {
        if (!h->haveAgg) { evp_set_error("iceAreaCell / totalMassCell are NULL and evp_aggregate has not run"); return EVP_ERR_STATE; }
    }
    if (!f->icePressure && !h->haveAggP) { evp_set_error("icePressure is NULL and evp_aggregate has not computed the Hibler strength"); return EVP_ERR_STATE; }
    EVP_REQUIRE(f->iceAreaCellInitial || f->iceAreaCell || h->haveAgg, "iceAreaCellInitial must not be NULL");
    const bool constantAir = o->use_air_stress && !f->airStressCellU;
    if (o->use_air_stress) {
        if (constantAir) EVP_REQUIRE(f->uAirVelocity && f->vAirVelocity && f->airDensity && (f->iceAreaCell || f->iceAreaCellInitial || h->haveAgg),
                                     "air stress: give airStressCellU/V or u/vAirVelocity + airDensity");
        else EVP_REQUIRE(f->airStressCellV, "airStressCellV is NULL");
    }
    const int tiltMode = !o->use_surface_tilt ? 0 : (o->geostrophic_surface_tilt ? 1 : 2);
    if (tiltMode == 2) EVP_REQUIRE(f->seaSurfaceTiltU && f->seaSurfaceTiltV, "ssh-gradient surface tilt needs seaSurfaceTiltU/V");
    if (!o->calc_velocity_masks) EVP_REQUIRE(f->solveStress && f->solveVelocity, "config_calc_velocity_masks = false needs the masks");
    EVP_CUDA(cudaSetDevice(h->device));
    evp_dev &d = h->d;
    const size_t nC = h->nCells, nV = h->nVertices;
    cudaStream_t s = h->stream;
    EVP_CUDA(cudaStreamSynchronize(s));          // the staging area may still feed kernels of a previous call
    Stage st{(char *)d.stage, d.stageBytes, 0};
    int rc = EVP_OK;
    auto stage_d = [&](const double *src, size_t n) -> const double * {
        if (!src || rc) return nullptr;
        double *p = (double *)st.take(n * 8 + 8);
        if (!p) { evp_set_error("staging area exhausted"); rc = EVP_ERR_ARGUMENT; return nullptr; }
        rc = evp_h2d(h, p, src, n * 8);
        return p;
    };
    auto stage_i = [&](const int *src, size_t n) -> const int * {
        if (!src || rc) return nullptr;
        int *p = (int *)st.take(n * 4 + 8);
        if (!p) { evp_set_error("staging area exhausted"); rc = EVP_ERR_ARGUMENT; return nullptr; }
        rc = evp_h2d(h, p, src, n * 4);
        return p;
    };
    PreArgs a{};
    a.nCells = h->nCells; a.nVerticesSolve = h->nVerticesSolve; a.nVertices = h->nVertices; a.M = h->M; a.D = h->D;
    a.nCp = h->nCp; a.nVp = h->nVp;
    a.useAir = o->use_air_stress != 0; a.constantAir = constantAir; a.useOcean = h->opt.use_ocean_stress != 0;
    a.tiltMode = tiltMode; a.calcMasks = o->calc_velocity_masks != 0; a.coldStart = o->cold_start;
    a.cr = h->opt.constitutive_relation_type;
    a.nEdges = d.nEdges; a.coc = d.coc; a.cov = d.cov; a.vflags = d.vflags; a.areaCell = d.areaCell; a.fVertex = d.fVertex;
    // iceAreaCell: the caller's, else the aggregate on the device; iceAreaCellInitial: the caller's, else iceAreaCell
    const double *areaNow = f->iceAreaCell ? nullptr : (h->haveAgg ? d.aggArea : nullptr);
    if (f->iceAreaCellInitial) a.areaInit = stage_d(f->iceAreaCellInitial, nC);
    if (f->iceAreaCell) areaNow = (f->iceAreaCell == f->iceAreaCellInitial) ? a.areaInit : stage_d(f->iceAreaCell, nC);
    if (!areaNow) areaNow = a.areaInit;
    if (!f->iceAreaCellInitial) a.areaInit = areaNow;
    a.areaNow = areaNow;
    a.mass = f->totalMassCell ? stage_d(f->totalMassCell, nC) : d.aggMass;
    a.Pin = f->icePressure ? stage_d(f->icePressure, nC) : d.aggP;
    a.uOcn = stage_d(f->uOceanVelocity, nC);
    a.vOcn = stage_d(f->vOceanVelocity, nC);
    if (a.useAir && !constantAir) { a.airU = stage_d(f->airStressCellU, nC); a.airV = stage_d(f->airStressCellV, nC); }
    if (a.useAir && constantAir) {
        a.uAir = stage_d(f->uAirVelocity, nC); a.vAir = stage_d(f->vAirVelocity, nC); a.rhoAir = stage_d(f->airDensity, nC);
    }
    if (tiltMode == 2) { a.tiltU = stage_d(f->seaSurfaceTiltU, nC); a.tiltV = stage_d(f->seaSurfaceTiltV, nC); }
    a.landIce = stage_i(f->landIceMask, nC);
    if (!a.calcMasks) { a.ssIn = stage_i(f->solveStress, nC); a.svIn = stage_i(f->solveVelocity, nV); }
    if (rc) return rc;
    a.solveStress = d.solveStress; a.solveVel = d.solveVel; a.solveVelPrev = d.solveVelPrev;
    a.P = d.P; a.airCell = d.airCell; a.sig = d.sig; a.sig12 = d.sig12;
    const bool weakDiv = h->opt.stress_divergence_scheme == EVP_SCHEME_WEAK;
    if (weakDiv && !h->haveWeak) { evp_set_error("the weak schemes need evp_set_weak_mesh first"); return EVP_ERR_STATE; }
    a.sigW = weakDiv ? d.sigW : nullptr; a.sigW12 = weakDiv ? d.sigW12 : nullptr;
    a.uv = d.uv; a.uvInit = d.uvInit; a.areaDen = d.areaDen; a.massf = d.massf; a.air = d.air; a.tilt = d.tilt;
    a.ocnStress = d.ocnStress; a.ocnVel = d.ocnVel;

    if (nC) k_pre_cells<<<grid_for(nC, 256), 256, 0, s>>>(a);
    if (nV) k_pre_vertices<<<grid_for(nV, 256), 256, 0, s>>>(a);
    EVP_CUDA(cudaGetLastError());
    // what the reference zeroes at the start of every dynamics step (init_subcycle_variables :2287-2324,
    // oceanStressCoeff :2303, replacementPressure variational.F:862)
    const size_t rowBytes = sizeof(double) * h->M * h->nCp;
    EVP_CUDA(cudaMemsetAsync(d.e11, 0, rowBytes, s));
    EVP_CUDA(cudaMemsetAsync(d.e22, 0, rowBytes, s));
    EVP_CUDA(cudaMemsetAsync(d.e12, 0, rowBytes, s));
    EVP_CUDA(cudaMemsetAsync(d.repP, 0, rowBytes, s));
    EVP_CUDA(cudaMemsetAsync(d.sdiv, 0, sizeof(double2) * h->nVp, s));
    EVP_CUDA(cudaMemsetAsync(d.ocoef, 0, sizeof(double) * h->nVp, s));
    if (h->haveWeak) {
        EVP_CUDA(cudaMemsetAsync(d.eW11, 0, sizeof(double) * h->nCp, s));
        EVP_CUDA(cudaMemsetAsync(d.eW22, 0, sizeof(double) * h->nCp, s));
        EVP_CUDA(cudaMemsetAsync(d.eW12, 0, sizeof(double) * h->nCp, s));
    }
    if ((rc = evp_halo_mark_masks(h))) return rc;
    if ((rc = evp_refresh_tile_flags(h, s))) return rc;
    // the velocity halo exchange that closes new_ice_velocities (:1281-1320)
    if ((rc = evp_halo_exchange(h, s, d.uv))) return rc;
    EVP_CUDA(cudaStreamSynchronize(s));           // pinned sources were copied asynchronously
    h->haveStep = true;
    return EVP_OK;
}

extern "C" int evp_aggregate(evp_handle *h, const evp_category_fields *cf, int hibler)
{
    EVP_REQUIRE(h != nullptr && cf != nullptr, "handle/categories is NULL");
    EVP_REQUIRE(cf->nCategories >= 1 && cf->iceAreaCategory && cf->iceVolumeCategory && cf->snowVolumeCategory,
                "nCategories >= 1 and the three category arrays are needed");
    EVP_CUDA(cudaSetDevice(h->device));
    evp_dev &d = h->d;
    const size_t nC = h->nCells, nK = (size_t)cf->nCategories;
    cudaStream_t s = h->stream;
    int rc;
    if (!d.aggArea) {
        double **bufs[] = {&d.aggArea, &d.aggVolIce, &d.aggVolSnow, &d.aggMass, &d.aggP};
        for (double **b : bufs) {
            if ((rc = evp_dev_alloc(h, (void **)b, sizeof(double) * h->nCp))) return rc;
            EVP_CUDA(cudaMemsetAsync(*b, 0, sizeof(double) * h->nCp, s));
        }
    }
    if (nC == 0) { h->haveAgg = true; h->haveAggP = hibler != 0; return EVP_OK; }
    EVP_CUDA(cudaStreamSynchronize(s));          // the staging area may still feed kernels of a previous call
    // the category arrays go through the staging area in chunks of cells
    const size_t perCell = 3 * nK * sizeof(double);
    const size_t chunk = std::max<size_t>(1, std::min(nC, (d.stageBytes - 1024) / perCell));
    for (size_t c0 = 0; c0 < nC; c0 += chunk) {
        const size_t cnt = std::min(chunk, nC - c0);
        double *sa = (double *)d.stage, *svi = sa + cnt * nK, *svs = svi + cnt * nK;
        if ((rc = evp_h2d(h, sa, cf->iceAreaCategory + c0 * nK, cnt * nK * 8))) return rc;
        if ((rc = evp_h2d(h, svi, cf->iceVolumeCategory + c0 * nK, cnt * nK * 8))) return rc;
        if ((rc = evp_h2d(h, svs, cf->snowVolumeCategory + c0 * nK, cnt * nK * 8))) return rc;
        k_aggregate<<<grid_for(cnt, 256), 256, 0, s>>>((int)cnt, (int)nK, sa, svi, svs, d.aggArea + c0, d.aggVolIce + c0,
                                                        d.aggVolSnow + c0, d.aggMass + c0, d.aggP + c0, hibler);
        EVP_CUDA(cudaGetLastError());
        if (c0 + chunk < nC) EVP_CUDA(cudaStreamSynchronize(s));     // the staging area is reused by the next chunk
    }
    EVP_CUDA(cudaStreamSynchronize(s));           // pinned sources were copied asynchronously
    h->haveAgg = true;
    h->haveAggP = hibler != 0;
    return EVP_OK;
}

extern "C" int evp_fetch_aggregate(evp_handle *h, double *iceAreaCell, double *iceVolumeCell, double *snowVolumeCell,
                                   double *totalMassCell, double *icePressure)
{
    EVP_REQUIRE(h != nullptr, "handle is NULL");
    if (!h->haveAgg) { evp_set_error("evp_fetch_aggregate before evp_aggregate"); return EVP_ERR_STATE; }
    if (icePressure && !h->haveAggP) { evp_set_error("evp_aggregate did not compute the Hibler strength"); return EVP_ERR_STATE; }
    EVP_CUDA(cudaSetDevice(h->device));
    evp_dev &d = h->d;
    const size_t nC = h->nCells;
    int rc;
    struct { double *host; const double *dev; } copies[] = {{iceAreaCell, d.aggArea}, {iceVolumeCell, d.aggVolIce},
        {snowVolumeCell, d.aggVolSnow}, {totalMassCell, d.aggMass}, {icePressure, d.aggP}};
    for (auto &c : copies)
        if (c.host && nC && (rc = evp_d2h(h, c.host, c.dev, nC * 8))) return rc;
    EVP_CUDA(cudaStreamSynchronize(h->stream));
    return EVP_OK;
}

extern "C" int evp_post_subcycle(evp_handle *h, const evp_post_fields *o)
{
    EVP_REQUIRE(h != nullptr && o != nullptr, "handle/out is NULL");
    if (!h->haveExt) { evp_set_error("evp_post_subcycle needs evp_set_mesh_ext first"); return EVP_ERR_STATE; }
    if (!h->haveStep) { evp_set_error("evp_post_subcycle before any dynamics step"); return EVP_ERR_STATE; }
    EVP_CUDA(cudaSetDevice(h->device));
    evp_dev &d = h->d;
    const size_t nC = h->nCells, nV = h->nVertices;
    cudaStream_t s = h->stream;
    int rc;
    PostArgs a{};
    a.nCells = h->nCells; a.nCellsSolve = h->nCellsSolve; a.nVerticesSolve = h->nVerticesSolve; a.M = h->M;
    a.nCp = h->nCp; a.nVp = h->nVp;
    a.useOcean = h->opt.use_ocean_stress != 0; a.oceanType = h->opt.ocean_stress_type;
    a.nEdges = d.nEdges; a.solveStress = d.solveStress; a.solveVel = d.solveVel; a.vflags = d.vflags; a.voc = d.voc;
    a.e11 = d.e11; a.e22 = d.e22; a.e12 = d.e12; a.repP = d.repP; a.sig = d.sig; a.sig12 = d.sig12;
    a.uv = d.uv; a.ocnVel = d.ocnVel; a.areaDen = d.areaDen; a.fVertex = d.fVertex; a.areaTri = d.areaTri;
    a.principal = d.contrib;                  // free between dynamics steps: rewritten by the next cell pass
    a.osFinal = d.osFinal; a.ocoef = d.ocoef;

    Stage st{(char *)d.stage, d.stageBytes, 0};
    double *cellOut[8];
    for (auto &p : cellOut) { p = (double *)st.take(nC * 8 + 8); EVP_REQUIRE(p, "staging area exhausted"); }
    a.div = cellOut[0]; a.shear = cellOut[1]; a.ridgeConv = cellOut[2]; a.ridgeShear = cellOut[3];
    a.oscU = cellOut[4]; a.oscV = cellOut[5];
    double *vtxOut[2];
    for (auto &p : vtxOut) { p = (double *)st.take(nV * 8 + 8); EVP_REQUIRE(p, "staging area exhausted"); }

    const bool wantDiv = o->divergence || o->shear || o->ridgeConvergence || o->ridgeShear;
    const bool wantOcean = o->oceanStressCellU || o->oceanStressCellV || o->oceanStressU || o->oceanStressV || o->oceanStressCoeff;
    const bool weakDiv = h->opt.stress_divergence_scheme == EVP_SCHEME_WEAK;
    if (weakDiv && !h->haveWeak) { evp_set_error("the weak schemes need evp_set_weak_mesh first"); return EVP_ERR_STATE; }
    if (wantDiv && nC) {
        if (weakDiv) {
            // the reference leaves the halo entries untouched; here they read 0
            for (int i = 0; i < 4; i++) EVP_CUDA(cudaMemsetAsync(cellOut[i], 0, nC * 8, s));
            if (h->nCellsSolve)
                k_post_div_shear_weak<<<grid_for(h->nCellsSolve, 256), 256, 0, s>>>(h->nCellsSolve, d.eW11, d.eW22, d.eW12, a.div,
                                                                                   a.shear, a.ridgeConv, a.ridgeShear);
        } else {
            k_post_div_shear<<<grid_for(nC, 256), 256, 0, s>>>(a);
        }
    }
    if ((o->principalStress1Weak || o->principalStress2Weak) && nC) {
        if (!weakDiv) { evp_set_error("principalStress*Weak needs the weak stress divergence scheme"); return EVP_ERR_ARGUMENT; }
        EVP_CUDA(cudaMemsetAsync(cellOut[6], 0, nC * 8, s));
        EVP_CUDA(cudaMemsetAsync(cellOut[7], 0, nC * 8, s));
        if (h->nCellsSolve)
            k_post_principal_weak<<<grid_for(h->nCellsSolve, 256), 256, 0, s>>>(h->nCellsSolve, d.sigW, d.sigW12, d.repPW,
                                                                               cellOut[6], cellOut[7]);
    }
    if (wantOcean) {
        if (nV) k_post_ocean_vertices<<<grid_for(nV, 256), 256, 0, s>>>(a);
        EVP_CUDA(cudaGetLastError());
        if ((rc = evp_halo_exchange(h, s, d.osFinal))) return rc;      // the oceanStress halo exchange (:3735-3760)
        if (nC) k_post_ocean_cells<<<grid_for(nC, 256), 256, 0, s>>>(a);
        if (nV) k_post_ocean_rescale<<<grid_for(nV, 256), 256, 0, s>>>(a);
    }
    EVP_CUDA(cudaGetLastError());
    struct { double *host; const double *dev; size_t n; } copies[] = {
        {o->divergence, a.div, nC}, {o->shear, a.shear, nC}, {o->ridgeConvergence, a.ridgeConv, nC},
        {o->ridgeShear, a.ridgeShear, nC}, {o->oceanStressCellU, a.oscU, nC}, {o->oceanStressCellV, a.oscV, nC},
        {o->oceanStressCoeff, d.ocoef, nV}, {o->principalStress1Weak, cellOut[6], nC}, {o->principalStress2Weak, cellOut[7], nC},
    };
    for (auto &c : copies)
        if (c.host && c.n && (rc = evp_d2h(h, c.host, c.dev, c.n * 8))) return rc;
    if ((o->oceanStressU || o->oceanStressV) && nV) {
        k_split<<<grid_for(nV, 256), 256, 0, s>>>(d.osFinal, vtxOut[0], vtxOut[1], nV);
        EVP_CUDA(cudaGetLastError());
        if (o->oceanStressU && (rc = evp_d2h(h, o->oceanStressU, vtxOut[0], nV * 8))) return rc;
        if (o->oceanStressV && (rc = evp_d2h(h, o->oceanStressV, vtxOut[1], nV * 8))) return rc;
        EVP_CUDA(cudaStreamSynchronize(s));      // vtxOut is reused below
    }
    if ((o->uVelocity || o->vVelocity) && nV) {
        k_split<<<grid_for(nV, 256), 256, 0, s>>>(d.uv, vtxOut[0], vtxOut[1], nV);
        EVP_CUDA(cudaGetLastError());
        if (o->uVelocity && (rc = evp_d2h(h, o->uVelocity, vtxOut[0], nV * 8))) return rc;
        if (o->vVelocity && (rc = evp_d2h(h, o->vVelocity, vtxOut[1], nV * 8))) return rc;
    }
    EVP_CUDA(cudaStreamSynchronize(s));
    if ((o->principalStress1Var || o->principalStress2Var) && nC) {
        k_post_principal<<<grid_for(h->nCellsSolve ? h->nCellsSolve : 1, 256), 256, 0, s>>>(a);
        EVP_CUDA(cudaGetLastError());
        if (o->principalStress1Var && (rc = evp_download_rows(h, o->principalStress1Var, (const double *)d.contrib, 1, 2, 0))) return rc;
        if (o->principalStress2Var && (rc = evp_download_rows(h, o->principalStress2Var, (const double *)d.contrib, 1, 2, 1))) return rc;
        // contrib served as scratch: tiles without work rely on their contrib rows being zero
        if ((rc = evp_refresh_tile_flags(h, s))) return rc;
    }
    EVP_CUDA(cudaStreamSynchronize(s));
    return EVP_OK;
}

extern "C" int evp_fetch_pre(evp_handle *h, const evp_pre_out_fields *o)
{
    EVP_REQUIRE(h != nullptr && o != nullptr, "handle/out is NULL");
    if (!h->haveStep) { evp_set_error("evp_fetch_pre before any dynamics step"); return EVP_ERR_STATE; }
    EVP_CUDA(cudaSetDevice(h->device));
    evp_dev &d = h->d;
    const size_t nC = h->nCells, nV = h->nVertices;
    cudaStream_t s = h->stream;
    int rc;
    EVP_CUDA(cudaStreamSynchronize(s));
    Stage st{(char *)d.stage, d.stageBytes, 0};
    struct { int *host; const uint8_t *dev; size_t n; } ints[] = {
        {o->solveStress, d.solveStress, nC}, {o->solveVelocity, d.solveVel, nV}, {o->solveVelocityPrevious, d.solveVelPrev, nV}};
    for (auto &q : ints) {
        if (!q.host || !q.n) continue;
        int *tmp = (int *)st.take(q.n * 4 + 8);
        EVP_REQUIRE(tmp, "staging area exhausted");
        k_u8_to_int<<<grid_for(q.n, 256), 256, 0, s>>>(q.dev, tmp, q.n, 1);
        EVP_CUDA(cudaGetLastError());
        if ((rc = evp_d2h(h, q.host, tmp, q.n * 4))) return rc;
    }
    if (o->icePressure && nC && (rc = evp_d2h(h, o->icePressure, d.P, nC * 8))) return rc;
    struct { double *a, *b; const double2 *src; } pairs[] = {
        {o->iceAreaVertex, nullptr, d.areaDen}, {o->totalMassVertex, o->totalMassVertexfVertex, d.massf},
        {o->airStressVertexU, o->airStressVertexV, d.air}, {o->surfaceTiltForceU, o->surfaceTiltForceV, d.tilt},
        {o->oceanStressU, o->oceanStressV, d.ocnStress}, {o->uOceanVelocityVertex, o->vOceanVelocityVertex, d.ocnVel},
        {o->uVelocityInitial, o->vVelocityInitial, d.uvInit}};
    for (auto &p : pairs) {
        if ((!p.a && !p.b) || !nV) continue;
        double *ra = (double *)st.take(nV * 8 + 8), *rb = (double *)st.take(nV * 8 + 8);
        EVP_REQUIRE(ra && rb, "staging area exhausted");
        k_split<<<grid_for(nV, 256), 256, 0, s>>>(p.src, ra, rb, nV);
        EVP_CUDA(cudaGetLastError());
        if (p.a && (rc = evp_d2h(h, p.a, ra, nV * 8))) return rc;
        if (p.b && (rc = evp_d2h(h, p.b, rb, nV * 8))) return rc;
    }
    EVP_CUDA(cudaStreamSynchronize(s));
    return EVP_OK;
}
