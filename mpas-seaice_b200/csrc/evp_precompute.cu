// evp_precompute.cu -- device version of seaice_init_velocity_solver_wachspress
// (reference: src/shared/mpas_seaice_velocity_solver_wachspress.F:46-161): Wachspress coefficients
// (:535-614), basis gradients at the cell vertices (:1083-1206) and the basis integrals by
// sub-triangle quadrature (:179-467), written straight into the SoA device layout.
//
// One thread per cell.  Compiled with --fmad=false and written in the reference's operation order, so
// the result is bit-identical to an FP64 non-FMA CPU evaluation (checked against the oracle in
// tests/test_gpu_parity.py).  At 10.5 M cells this replaces ~15 GB of host->device traffic and minutes
// of host time by well under a second of device time.
#include "evp_internal.cuh"

namespace {

constexpr int QMAX = 64;
__constant__ double cQu[QMAX], cQv[QMAX], cQw[QMAX];

// Quadrature tables: D. A. Dunavant, Int. J. Num. Meth. Engng 21 (1985) 1129-1148, with the
// reference's truncated literals (wachspress.F:1441-1597) -- the digits must not be "improved".
struct Rule { int n; double norm; double u[QMAX], v[QMAX], w[QMAX]; };

// dunavant orders 9, 10, 12 and the 'fekete' rules (wachspress.F:1599-1941), generated from the reference's routines
#include "evp_quadrature_tables.inc"

int make_rule(int type, int order, Rule &r)
{
    if (type == 1) {   // trapezoidal (wachspress.F:1301-1387)
        const int nT = order;
        const int np = ((nT + 1) * (nT + 1) + (nT + 1)) / 2;
        if (nT < 1 || np > QMAX) return 1;
        r.n = np;
        int ij = 0;
        for (int i = 0; i <= nT; i++)
            for (int j = 0; j <= nT - i; j++) {
                r.u[ij] = (double)i / (double)nT;
                r.v[ij] = (double)j / (double)nT;
                double w = 0.0;
                if (i <= nT - j) {
                    if (i == nT || j == nT || (i == 0 && j == 0)) w = 1.0;
                    else if ((j == 0 && i != 0 && i != nT) || (i == 0 && j != 0 && j != nT) ||
                             (i == nT - j && i != 0 && j != 0)) w = 3.0;
                    else w = 6.0;
                }
                r.w[ij++] = w;
            }
        r.norm = 6.0 * ((double)nT * (double)nT);
        return 0;
    }
    if (type != 0 && type != 2) return 1;
    {
        int n = 0;
        double norm = 0.0;
        const double *t = generated_rule(type, order, n, norm);
        if (t != nullptr) {
            if (n > QMAX) return 1;
            r.n = n;
            r.norm = norm;
            for (int k = 0; k < n; k++) {
                r.u[k] = t[3 * k];
                r.v[k] = t[3 * k + 1];
                r.w[k] = t[3 * k + 2];
            }
            return 0;
        }
    }
    if (type != 0) return 1;
    r.norm = 2.0;
    // each rule: a list of (u, v, w) triplets
    static const double d1[] = {0.33333333333333, 0.33333333333333, 1.00000000000000};
    static const double d2[] = {0.16666666666667, 0.16666666666667, 0.33333333333333,
                                0.16666666666667, 0.66666666666667, 0.33333333333333,
                                0.66666666666667, 0.16666666666667, 0.33333333333333};
    static const double d3[] = {0.33333333333333, 0.33333333333333, -0.56250000000000,
                                0.20000000000000, 0.20000000000000, 0.52083333333333,
                                0.20000000000000, 0.60000000000000, 0.52083333333333,
                                0.60000000000000, 0.20000000000000, 0.52083333333333};
    static const double d4[] = {0.44594849091597, 0.44594849091597, 0.22338158967801,
                                0.44594849091597, 0.10810301816807, 0.22338158967801,
                                0.10810301816807, 0.44594849091597, 0.22338158967801,
                                0.09157621350977, 0.09157621350977, 0.10995174365532,
                                0.09157621350977, 0.81684757298046, 0.10995174365532,
                                0.81684757298046, 0.09157621350977, 0.10995174365532};
    static const double d5[] = {0.33333333333333, 0.33333333333333, 0.22500000000000,
                                0.47014206410511, 0.47014206410511, 0.13239415278851,
                                0.47014206410511, 0.05971587178977, 0.13239415278851,
                                0.05971587178977, 0.47014206410511, 0.13239415278851,
                                0.10128650732346, 0.10128650732346, 0.12593918054483,
                                0.10128650732346, 0.79742698535309, 0.12593918054483,
                                0.79742698535309, 0.10128650732346, 0.12593918054483};
    static const double d6[] = {0.24928674517091, 0.24928674517091, 0.11678627572638,
                                0.24928674517091, 0.50142650965818, 0.11678627572638,
                                0.50142650965818, 0.24928674517091, 0.11678627572638,
                                0.06308901449150, 0.06308901449150, 0.05084490637021,
                                0.06308901449150, 0.87382197101700, 0.05084490637021,
                                0.87382197101700, 0.06308901449150, 0.05084490637021,
                                0.31035245103378, 0.63650249912140, 0.08285107561837,
                                0.63650249912140, 0.05314504984482, 0.08285107561837,
                                0.05314504984482, 0.31035245103378, 0.08285107561837,
                                0.63650249912140, 0.31035245103378, 0.08285107561837,
                                0.31035245103378, 0.05314504984482, 0.08285107561837,
                                0.05314504984482, 0.63650249912140, 0.08285107561837};
    static const double d7[] = {0.33333333333333, 0.33333333333333, -0.14957004446768,
                                0.26034596607904, 0.26034596607904, 0.17561525743321,
                                0.26034596607904, 0.47930806784192, 0.17561525743321,
                                0.47930806784192, 0.26034596607904, 0.17561525743321,
                                0.06513010290222, 0.06513010290222, 0.05334723560884,
                                0.06513010290222, 0.86973979419557, 0.05334723560884,
                                0.86973979419557, 0.06513010290222, 0.05334723560884,
                                0.31286549600487, 0.63844418856981, 0.07711376089026,
                                0.63844418856981, 0.04869031542532, 0.07711376089026,
                                0.04869031542532, 0.31286549600487, 0.07711376089026,
                                0.63844418856981, 0.31286549600487, 0.07711376089026,
                                0.31286549600487, 0.04869031542532, 0.07711376089026,
                                0.04869031542532, 0.63844418856981, 0.07711376089026};
    static const double d8[] = {0.33333333333333, 0.33333333333333, 0.14431560767779,
                                0.45929258829272, 0.45929258829272, 0.09509163426728,
                                0.45929258829272, 0.08141482341455, 0.09509163426728,
                                0.08141482341455, 0.45929258829272, 0.09509163426728,
                                0.17056930775176, 0.17056930775176, 0.10321737053472,
                                0.17056930775176, 0.65886138449648, 0.10321737053472,
                                0.65886138449648, 0.17056930775176, 0.10321737053472,
                                0.05054722831703, 0.05054722831703, 0.03245849762320,
                                0.05054722831703, 0.89890554336594, 0.03245849762320,
                                0.89890554336594, 0.05054722831703, 0.03245849762320,
                                0.26311282963464, 0.72849239295540, 0.02723031417443,
                                0.72849239295540, 0.00839477740996, 0.02723031417443,
                                0.00839477740996, 0.26311282963464, 0.02723031417443,
                                0.72849239295540, 0.26311282963464, 0.02723031417443,
                                0.26311282963464, 0.00839477740996, 0.02723031417443,
                                0.00839477740996, 0.72849239295540, 0.02723031417443};
    const double *tabs[] = {nullptr, d1, d2, d3, d4, d5, d6, d7, d8};
    const int counts[] = {0, 1, 3, 4, 6, 7, 12, 13, 16};
    if (order < 1 || order > 8) return 1;
    r.n = counts[order];
    for (int k = 0; k < r.n; k++) {
        r.u[k] = tabs[order][3 * k];
        r.v[k] = tabs[order][3 * k + 1];
        r.w[k] = tabs[order][3 * k + 2];
    }
    return 0;
}

template <int M>
struct WCell {
    int n;
    double x[M], y[M], A[M], B[M], kappa[M];
};

__device__ __forceinline__ int wrap1(int input, int n)   // seaice_wrapped_index, 1-based
{
    int m = (input - 1) % n;
    if (m < 0) m += n;
    return m + 1;
}

// numerators (wachspress.F:864-925), their derivatives (:939-1035), then all basis functions
// (:682-749) and derivatives (:763-850) at one point.  kappa(j,i) of the reference does not depend
// on i, so the numerators are shared by all basis indices.
template <int M>
__device__ void eval_all(const WCell<M> &c, double x, double y, double *phi, double *dpx, double *dpy)
{
    const int n = c.n;
    double num[M], dnx[M], dny[M], e[M];
    double denominator = 0.0, sdx = 0.0, sdy = 0.0;
    for (int k = 0; k < n; k++) e[k] = 1.0 - c.A[k] * x - c.B[k] * y;     // wachspress_edge_equation (:1049-1069)
    for (int j = 1; j <= n; j++) {
        const int i1 = j, i2 = wrap1(j + 1, n);
        int sub[M], ns = 0;
        for (int k = 1; k <= n; k++)
            if (k != i1 && k != i2) sub[ns++] = k - 1;                    // wachspress_indexes (:628-668)
        double numerator = 1.0;
        for (int k = 0; k < ns; k++) numerator = numerator * e[sub[k]];
        numerator = numerator * c.kappa[j - 1];
        num[j - 1] = numerator;
        denominator = denominator + numerator;
        double spx = 0.0, spy = 0.0;
        for (int k = 0; k < ns; k++) {
            double px = 1.0, py = 1.0;
            for (int l = 0; l < k; l++) { px = px * e[sub[l]]; py = py * e[sub[l]]; }
            px = px * (-c.A[sub[k]]);
            py = py * (-c.B[sub[k]]);
            for (int l = k + 1; l < ns; l++) { px = px * e[sub[l]]; py = py * e[sub[l]]; }
            spx = spx + px;
            spy = spy + py;
        }
        dnx[j - 1] = spx * c.kappa[j - 1];
        dny[j - 1] = spy * c.kappa[j - 1];
        sdx = sdx + dnx[j - 1];
        sdy = sdy + dny[j - 1];
    }
    for (int i = 0; i < n; i++) {
        if (phi) phi[i] = num[i] / denominator;
        dpx[i] = dnx[i] / denominator - (num[i] / (denominator * denominator)) * sdx;
        dpy[i] = dny[i] / denominator - (num[i] / (denominator * denominator)) * sdy;
    }
}

template <int M>
__global__ void __launch_bounds__(64) k_wachspress(int nCells, size_t nCp, int Mh, const uint8_t *__restrict__ nEdges,
                                                    const double *__restrict__ xl, const double *__restrict__ yl,
                                                    int nq, double norm, double2 *__restrict__ G,
                                                    double2 *__restrict__ Suv, double *__restrict__ Sm)
{
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= nCells) return;
    const int n = nEdges[cell];
    if (n < 3 || n > M) return;
    WCell<M> c;
    c.n = n;
    for (int i = 0; i < n; i++) { c.x[i] = xl[(size_t)Mh * cell + i]; c.y[i] = yl[(size_t)Mh * cell + i]; }
    // calc_wachspress_coefficients (:535-614)
    for (int iv = 1; iv <= n; iv++) {
        int i1 = iv - 1, i2 = iv;
        if (i1 < 1) i1 = i1 + n;
        const double den = c.x[i1 - 1] * c.y[i2 - 1] - c.x[i2 - 1] * c.y[i1 - 1];
        c.A[iv - 1] = (c.y[i2 - 1] - c.y[i1 - 1]) / den;
        c.B[iv - 1] = (c.x[i1 - 1] - c.x[i2 - 1]) / den;
    }
    c.kappa[0] = 1.0;
    for (int j = 2; j <= n; j++) {
        int i0 = j - 1, i1 = j, i2 = j + 1;
        if (i2 > n) i2 = i2 - n;
        c.kappa[j - 1] = c.kappa[j - 2] *
            (c.A[i2 - 1] * (c.x[i0 - 1] - c.x[i1 - 1]) + c.B[i2 - 1] * (c.y[i0 - 1] - c.y[i1 - 1])) /
            (c.A[i0 - 1] * (c.x[i1 - 1] - c.x[i0 - 1]) + c.B[i0 - 1] * (c.y[i1 - 1] - c.y[i0 - 1]));
    }
    // gradients at the vertices: kept only for iGradientVertex in {i-1, i, i+1} (:1178-1191)
    for (int jg = 1; jg <= n; jg++) {
        double dx[M], dy[M];
        eval_all<M>(c, c.x[jg - 1], c.y[jg - 1], nullptr, dx, dy);
        for (int ib = 1; ib <= n; ib++) {
            double2 g = make_double2(0.0, 0.0);
            if (jg == ib || jg == wrap1(ib - 1, n) || jg == wrap1(ib + 1, n)) g = make_double2(dx[ib - 1], dy[ib - 1]);
            G[evp_tix((jg - 1) * M + (ib - 1), cell, M * M)] = g;
        }
    }
    // integrals (:304-467): per pair (iStress, iVelocity): sum over sub-triangles of (sum over points) / norm
    double bU[M][M], bV[M][M], bM[M][M];     // [iVel][iStr]
    for (int a = 0; a < n; a++)
        for (int b = 0; b < n; b++) { bU[a][b] = 0.0; bV[a][b] = 0.0; bM[a][b] = 0.0; }
    for (int s = 1; s <= n; s++) {
        const int i1 = s, i2 = wrap1(s + 1, n);
        // get_triangle_mapping (:485-517) with (x1,y1) = (1,0), (x2,y2) = (0,1)
        const double x1 = 1.0, y1 = 0.0, x2 = 0.0, y2 = 1.0;
        const double u1 = c.x[i1 - 1], v1 = c.y[i1 - 1], u2 = c.x[i2 - 1], v2 = c.y[i2 - 1];
        const double m11 = (u2 * y1 - u1 * y2) / (x2 * y1 - x1 * y2);
        const double m12 = (u1 * x2 - u2 * x1) / (y1 * x2 - y2 * x1);
        const double m21 = (v2 * y1 - v1 * y2) / (x2 * y1 - x1 * y2);
        const double m22 = (v1 * x2 - v2 * x1) / (y1 * x2 - y2 * x1);
        const double jac = m11 * m22 - m12 * m21;
        double sU[M][M], sV[M][M], sM[M][M];
        for (int a = 0; a < n; a++)
            for (int b = 0; b < n; b++) { sU[a][b] = 0.0; sV[a][b] = 0.0; sM[a][b] = 0.0; }
        for (int p = 0; p < nq; p++) {
            const double x = m11 * cQu[p] + m12 * cQv[p];
            const double y = m21 * cQu[p] + m22 * cQv[p];
            double phi[M], dpx[M], dpy[M];
            eval_all<M>(c, x, y, phi, dpx, dpy);
            for (int is = 0; is < n; is++) {
                const double tmp = jac * cQw[p] * phi[is];
                for (int iv = 0; iv < n; iv++) {
                    sU[iv][is] = sU[iv][is] + tmp * dpx[iv];
                    sV[iv][is] = sV[iv][is] + tmp * dpy[iv];
                    sM[iv][is] = sM[iv][is] + tmp * phi[iv];
                }
            }
        }
        for (int a = 0; a < n; a++)
            for (int b = 0; b < n; b++) {
                bU[a][b] = bU[a][b] + sU[a][b] / norm;
                bV[a][b] = bV[a][b] + sV[a][b] / norm;
                bM[a][b] = bM[a][b] + sM[a][b] / norm;
            }
    }
    for (int iv = 0; iv < n; iv++)
        for (int is = 0; is < n; is++) {
            const size_t q = evp_tix(iv * M + is, cell, M * M);
            Suv[q] = make_double2(bU[iv][is], bV[iv][is]);
            Sm[q] = bM[iv][is];
        }
}

// ------------------------------------------------------------------------------------------
// PWL basis (reference: src/shared/mpas_seaice_velocity_solver_pwl.F:44-373) with the 3x3 LU solve of
// src/shared/mpas_seaice_numerics.F:44-212 (Crout, implicit scaling, partial pivoting).  Only + - * / sqrt
// and comparisons: bit-identical to the non-FMA host evaluation.
// ------------------------------------------------------------------------------------------
__device__ void lu_decomposition3(double a[3][3], int indices[3])
{
    const double tiny = 1.0e-20;
    double maxa[3];
    for (int i = 0; i < 3; i++) {
        double m = 0.0;
        for (int j = 0; j < 3; j++) if (fabs(a[i][j]) > m) m = fabs(a[i][j]);
        maxa[i] = 1.0 / m;
    }
    for (int j = 0; j < 3; j++) {
        int jmax = j;
        double best = maxa[j] * fabs(a[j][j]);
        for (int i = j + 1; i < 3; i++) {          // maxloc: first maximum
            const double val = maxa[i] * fabs(a[i][j]);
            if (val > best) { best = val; jmax = i; }
        }
        if (j != jmax) {
            for (int k = 0; k < 3; k++) { const double t = a[jmax][k]; a[jmax][k] = a[j][k]; a[j][k] = t; }
            maxa[jmax] = maxa[j];
        }
        indices[j] = jmax;
        if (a[j][j] == 0.0) a[j][j] = tiny;
        for (int i = j + 1; i < 3; i++) a[i][j] = a[i][j] / a[j][j];
        for (int i = j + 1; i < 3; i++)
            for (int k = j + 1; k < 3; k++) a[i][k] = a[i][k] - a[i][j] * a[j][k];
    }
}
__device__ void lu_back_substitution3(double a[3][3], const int indices[3], double b[3])
{
    int j = -1;
    for (int i = 0; i < 3; i++) {
        const int k = indices[i];
        double sums = b[k];
        b[k] = b[i];
        if (j != -1) {
            double dot = 0.0;
            for (int l = j; l <= i - 1; l++) dot = dot + a[i][l] * b[l];
            sums = sums - dot;
        } else if (sums != 0.0) {
            j = i;
        }
        b[i] = sums;
    }
    for (int i = 2; i >= 0; i--) {
        double dot = 0.0;
        for (int l = i + 1; l < 3; l++) dot = dot + a[i][l] * b[l];
        b[i] = (b[i] - dot) / a[i][i];
    }
}
__device__ void solve_linear_basis_system3(const double left[3][3], const double rhs[3], double sol[3])
{
    double a[3][3];
    int indices[3];
    for (int i = 0; i < 3; i++) {
        sol[i] = rhs[i];
        for (int j = 0; j < 3; j++) a[i][j] = left[i][j];
    }
    lu_decomposition3(a, indices);
    lu_back_substitution3(a, indices, sol);
}

template <int M>
__global__ void __launch_bounds__(64) k_pwl(int nCells, int Mh, const uint8_t *__restrict__ nEdges,
                                             const double *__restrict__ xlAll, const double *__restrict__ ylAll,
                                             const int *__restrict__ edgesOnCell, const double *__restrict__ dvEdge,
                                             int nEdgesTotal, const double *__restrict__ areaCell,
                                             double2 *__restrict__ G, double2 *__restrict__ Suv, double *__restrict__ Sm)
{
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= nCells) return;
    const int n = nEdges[cell];
    if (n < 3 || n > M) return;
    double xl[M], yl[M];
    for (int i = 0; i < n; i++) { xl[i] = xlAll[(size_t)Mh * cell + i]; yl[i] = ylAll[(size_t)Mh * cell + i]; }
    const double alphaPWL = 1.0 / (double)n;
    double xC = 0.0, yC = 0.0;
    for (int j = 0; j < n; j++) { xC = xC + alphaPWL * xl[j]; yC = yC + alphaPWL * yl[j]; }
    // sub-triangle areas by Heron with dvEdge as the outer side, rescaled to areaCell (:159-183)
    double subArea[M], subAreaSum = 0.0;
    for (int s = 1; s <= n; s++) {
        int iEdge = edgesOnCell[(size_t)Mh * cell + (s - 1)];
        if (iEdge < 1 || iEdge > nEdgesTotal) iEdge = 1;
        const int v1 = s, v2 = wrap1(s + 1, n);
        const double c = dvEdge[iEdge - 1];
        const double a = sqrt((xl[v1 - 1] - xC) * (xl[v1 - 1] - xC) + (yl[v1 - 1] - yC) * (yl[v1 - 1] - yC));
        const double b = sqrt((xl[v2 - 1] - xC) * (xl[v2 - 1] - xC) + (yl[v2 - 1] - yC) * (yl[v2 - 1] - yC));
        const double sp = (a + b + c) * 0.5;
        subArea[s - 1] = sqrt(sp * (sp - a) * (sp - b) * (sp - c));
        subAreaSum = subAreaSum + subArea[s - 1];
    }
    {
        const double scale = areaCell[cell] / subAreaSum;
        for (int s = 0; s < n; s++) subArea[s] = subArea[s] * scale;
    }
    // linear basis on each sub-triangle: two 3x3 solves (:186-231)
    double sbU[M][3], sbV[M][3];
    for (int s = 1; s <= n; s++) {
        const int v1 = s, v2 = wrap1(s + 1, n);
        const double left[3][3] = {{xl[v1 - 1] - xC, yl[v1 - 1] - yC, 1.0},
                                   {xl[v2 - 1] - xC, yl[v2 - 1] - yC, 1.0},
                                   {0.0, 0.0, 1.0}};
        const double rhs1[3] = {1.0, 0.0, 0.0}, rhs2[3] = {0.0, 1.0, 0.0};
        double sol[3];
        solve_linear_basis_system3(left, rhs1, sol);
        sbU[s - 1][0] = sol[0]; sbV[s - 1][0] = sol[1];
        solve_linear_basis_system3(left, rhs2, sol);
        sbU[s - 1][1] = sol[0]; sbV[s - 1][1] = sol[1];
        sbU[s - 1][2] = -sbU[s - 1][0] - sbU[s - 1][1];
        sbV[s - 1][2] = -sbV[s - 1][0] - sbV[s - 1][1];
    }
    // gradient of basis function ib on sub-cell s (:234-257)
    double scU[M][M], scV[M][M];
    for (int ib = 1; ib <= n; ib++) {
        for (int s = 1; s <= n; s++) {
            scU[ib - 1][s - 1] = sbU[s - 1][2] * alphaPWL;
            scV[ib - 1][s - 1] = sbV[s - 1][2] * alphaPWL;
            if (s == ib) {
                scU[ib - 1][s - 1] = scU[ib - 1][s - 1] + sbU[s - 1][0];
                scV[ib - 1][s - 1] = scV[ib - 1][s - 1] + sbV[s - 1][0];
            } else if (s == wrap1(ib - 1, n)) {
                scU[ib - 1][s - 1] = scU[ib - 1][s - 1] + sbU[s - 1][1];
                scV[ib - 1][s - 1] = scV[ib - 1][s - 1] + sbV[s - 1][1];
            }
        }
    }
    // gradient at a vertex = mean of its two sub-cells (:260-274): dense in (ib, ig)
    for (int ib = 1; ib <= n; ib++) {
        for (int ig = 1; ig <= n; ig++) {
            const int s1 = ig, s2 = wrap1(ig - 1, n);
            G[evp_tix((ig - 1) * M + (ib - 1), cell, M * M)] =
                make_double2(0.5 * (scU[ib - 1][s1 - 1] + scU[ib - 1][s2 - 1]), 0.5 * (scV[ib - 1][s1 - 1] + scV[ib - 1][s2 - 1]));
        }
    }
    // integrals (:277-362)
    for (int is = 1; is <= n; is++) {
        for (int iv = 1; iv <= n; iv++) {
            double bU = 0.0, bV = 0.0, bM = 0.0;
            for (int s = 1; s <= n; s++) {
                double basisIntegral;
                if (s == is || s == wrap1(is - 1, n)) basisIntegral = ((alphaPWL + 1) * subArea[s - 1]) / 3.0;
                else basisIntegral = (alphaPWL * subArea[s - 1]) / 3.0;
                bU = bU + scU[iv - 1][s - 1] * basisIntegral;
                bV = bV + scV[iv - 1][s - 1] * basisIntegral;
            }
            for (int s = 1; s <= n; s++) {
                const int tS = (s == is) ? 1 : (s == wrap1(is - 1, n)) ? 2 : 3;
                const int tV = (s == iv) ? 1 : (s == wrap1(iv - 1, n)) ? 2 : 3;
                double val = 0.0;
                if ((tS == 1 && tV == 1) || (tS == 2 && tV == 2)) val = 2.0 * (alphaPWL * alphaPWL) + 2.0 * alphaPWL + 2.0;
                else if ((tS == 1 && tV == 2) || (tS == 2 && tV == 1)) val = 2.0 * (alphaPWL * alphaPWL) + 2.0 * alphaPWL + 1.0;
                else if (tS == 3 && tV == 3) val = 2.0 * (alphaPWL * alphaPWL);
                else val = 2.0 * (alphaPWL * alphaPWL) + alphaPWL;
                val = val * subArea[s - 1] / 12.0;
                bM = bM + val;
            }
            const size_t q = evp_tix((iv - 1) * M + (is - 1), cell, M * M);
            Suv[q] = make_double2(bU, bV);
            Sm[q] = bM;
        }
    }
}

}  // namespace

// The rule evp_precompute_wachspress integrates with (get_integration_factors, wachspress.F:1224-1287): host-only, no
// device needed.  u, v, w: room for 64 points each.
extern "C" int evp_integration_rule(int integrationType, int integrationOrder, int *nPoints, double *u, double *v,
                                    double *w, double *normalizationFactor)
{
    EVP_REQUIRE(nPoints != nullptr && u != nullptr && v != nullptr && w != nullptr && normalizationFactor != nullptr,
                "NULL argument");
    Rule r;
    if (make_rule(integrationType, integrationOrder, r)) {
        evp_set_error("unsupported integration rule (type %d, order %d)", integrationType, integrationOrder);
        return EVP_ERR_ARGUMENT;
    }
    *nPoints = r.n;
    *normalizationFactor = r.norm;
    for (int k = 0; k < r.n; k++) {
        u[k] = r.u[k];
        v[k] = r.v[k];
        w[k] = r.w[k];
    }
    return EVP_OK;
}

extern "C" int evp_precompute_wachspress(evp_handle *h, const double *xLocal, const double *yLocal,
                                         int integrationType, int integrationOrder)
{
    EVP_REQUIRE(h != nullptr && xLocal != nullptr && yLocal != nullptr, "NULL argument");
    Rule r;
    if (make_rule(integrationType, integrationOrder, r)) {
        evp_set_error("unsupported integration rule (type %d, order %d)", integrationType, integrationOrder);
        return EVP_ERR_ARGUMENT;
    }
    EVP_CUDA(cudaSetDevice(h->device));
    const size_t nC = h->nCells;
    if (nC == 0) { h->haveBasis = true; return EVP_OK; }
    {
        int rc = evp_basis_begin(h);
        if (rc) return rc;
    }
    const size_t bytes = (size_t)h->Mh * nC * sizeof(double);
    EVP_REQUIRE(2 * bytes + 512 <= h->d.stageBytes, "staging area too small for the local coordinates");
    EVP_CUDA(cudaStreamSynchronize(h->stream));
    EVP_CUDA(cudaMemcpyToSymbolAsync(cQu, r.u, sizeof(double) * r.n, 0, cudaMemcpyHostToDevice, h->stream));
    EVP_CUDA(cudaMemcpyToSymbolAsync(cQv, r.v, sizeof(double) * r.n, 0, cudaMemcpyHostToDevice, h->stream));
    EVP_CUDA(cudaMemcpyToSymbolAsync(cQw, r.w, sizeof(double) * r.n, 0, cudaMemcpyHostToDevice, h->stream));
    double *dx = (double *)h->d.stage;
    double *dy = (double *)((char *)h->d.stage + ((bytes + 255) & ~(size_t)255));
    EVP_CUDA(cudaMemcpyAsync(dx, xLocal, bytes, cudaMemcpyHostToDevice, h->stream));
    EVP_CUDA(cudaMemcpyAsync(dy, yLocal, bytes, cudaMemcpyHostToDevice, h->stream));
    const int block = 64;
    const unsigned grid = (unsigned)((nC + block - 1) / block);
    switch (h->M) {
    case 4: k_wachspress<4><<<grid, block, 0, h->stream>>>((int)nC, h->nCp, h->Mh, h->d.nEdges, dx, dy, r.n, r.norm, h->d.G, h->d.Suv, h->d.Sm); break;
    case 6: k_wachspress<6><<<grid, block, 0, h->stream>>>((int)nC, h->nCp, h->Mh, h->d.nEdges, dx, dy, r.n, r.norm, h->d.G, h->d.Suv, h->d.Sm); break;
    default: k_wachspress<8><<<grid, block, 0, h->stream>>>((int)nC, h->nCp, h->Mh, h->d.nEdges, dx, dy, r.n, r.norm, h->d.G, h->d.Suv, h->d.Sm); break;
    }
    EVP_CUDA(cudaGetLastError());
    EVP_CUDA(cudaStreamSynchronize(h->stream));
    return evp_basis_finalize(h);
}

extern "C" int evp_precompute_pwl(evp_handle *h, const double *xLocal, const double *yLocal, const int *edgesOnCell,
                                  const double *dvEdge, int nEdges, const double *areaCell)
{
    EVP_REQUIRE(h != nullptr && xLocal && yLocal && edgesOnCell && dvEdge && areaCell, "NULL argument");
    EVP_REQUIRE(nEdges >= 1, "nEdges must be >= 1");
    EVP_CUDA(cudaSetDevice(h->device));
    const size_t nC = h->nCells;
    if (nC == 0) { h->haveBasis = true; return EVP_OK; }
    {
        int rc = evp_basis_begin(h);
        if (rc) return rc;
    }
    EVP_CUDA(cudaMemsetAsync(h->d.Suv, 0, sizeof(double2) * h->M * h->M * h->nCp, h->stream));
    EVP_CUDA(cudaMemsetAsync(h->d.Sm, 0, sizeof(double) * h->M * h->M * h->nCp, h->stream));
    EVP_CUDA(cudaStreamSynchronize(h->stream));
    Stage st{(char *)h->d.stage, h->d.stageBytes, 0};
    const size_t rowBytes = (size_t)h->Mh * nC * sizeof(double);
    double *dx = (double *)st.take(rowBytes), *dy = (double *)st.take(rowBytes);
    int *de = (int *)st.take((size_t)h->Mh * nC * sizeof(int));
    double *dv = (double *)st.take((size_t)nEdges * sizeof(double)), *da = (double *)st.take(nC * sizeof(double));
    EVP_REQUIRE(dx && dy && de && dv && da, "staging area too small for the PWL inputs");
    EVP_CUDA(cudaMemcpyAsync(dx, xLocal, rowBytes, cudaMemcpyHostToDevice, h->stream));
    EVP_CUDA(cudaMemcpyAsync(dy, yLocal, rowBytes, cudaMemcpyHostToDevice, h->stream));
    EVP_CUDA(cudaMemcpyAsync(de, edgesOnCell, (size_t)h->Mh * nC * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    EVP_CUDA(cudaMemcpyAsync(dv, dvEdge, (size_t)nEdges * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    EVP_CUDA(cudaMemcpyAsync(da, areaCell, nC * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    const int block = 64;
    const unsigned grid = (unsigned)((nC + block - 1) / block);
    switch (h->M) {
    case 4: k_pwl<4><<<grid, block, 0, h->stream>>>((int)nC, h->Mh, h->d.nEdges, dx, dy, de, dv, nEdges, da, h->d.G, h->d.Suv, h->d.Sm); break;
    case 6: k_pwl<6><<<grid, block, 0, h->stream>>>((int)nC, h->Mh, h->d.nEdges, dx, dy, de, dv, nEdges, da, h->d.G, h->d.Suv, h->d.Sm); break;
    default: k_pwl<8><<<grid, block, 0, h->stream>>>((int)nC, h->Mh, h->d.nEdges, dx, dy, de, dv, nEdges, da, h->d.G, h->d.Suv, h->d.Sm); break;
    }
    EVP_CUDA(cudaGetLastError());
    EVP_CUDA(cudaStreamSynchronize(h->stream));
    return evp_basis_finalize(h);
}
