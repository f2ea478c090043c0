/*
 * oracle/evp_precompute_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the one-time precompute that feeds the EVP subcycle
 * (seaice_init_velocity_solver_variational and callees).  Same conventions as evp_oracle.c.  Pinned by outputs of
 * the reference's own source executed here (tests/golden/fortran_subset.py interprets
 * init_velocity_solver_variational_primary_mesh with the Wachspress / PWL routines below it; fixtures
 * tests/golden/init/refexec_init_*.npz, reproduced bit for bit: tests/test_refexec_init.py), by the exact reproduction properties
 * of the bases (tests/test_oracle_kat.py) and by an independent closed form (tests/test_wachspress_independent.py).
 *
 * Arrays: Fortran column-major, 1-based index values, junk slot at the end.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define IDX2(i, c, M) ((size_t)((i) - 1) + (size_t)(M) * (size_t)((c) - 1))
#define IDX3(i, j, c, M) ((size_t)((i) - 1) + (size_t)(M) * ((size_t)((j) - 1) + (size_t)(M) * (size_t)((c) - 1)))
#define MAXE 12

/* seaice_wrapped_index (src/shared/mpas_seaice_velocity_solver_variational_shared.F:372-385) */
static inline int wrapped_index(int input, int nelements)
{
    int m = (input - 1) % nelements;
    if (m < 0) m += nelements;
    return m + 1;
}

/* seaice_grid_rotation_forward (src/shared/mpas_seaice_mesh.F:2350-2381) */
static inline void grid_rotation_forward(double *xp, double *yp, double *zp, double x, double y, double z, int rotate)
{
    if (rotate) { *xp = -z; *yp = y; *zp = x; }
    else        { *xp = x;  *yp = y; *zp = z; }
}

/* seaice_calc_variational_metric_terms (variational_shared.F:293-358) */
void orc_calc_variational_metric_terms(double *tanLatVertexRotatedOverRadius, int nVertices,
                                       const double *xVertex, const double *yVertex, const double *zVertex,
                                       double sphereRadius, int rotateCartesianGrid, int includeMetricTerms)
{
    if (includeMetricTerms) {
        for (int i = 0; i < nVertices; i++) {
            double xr, yr, zr;
            grid_rotation_forward(&xr, &yr, &zr, xVertex[i], yVertex[i], zVertex[i], rotateCartesianGrid);
            const double latVertexRotated = asin(zr / sphereRadius);
            tanLatVertexRotatedOverRadius[i] = tan(latVertexRotated) / sphereRadius;
        }
    } else {
        for (int i = 0; i < nVertices; i++) tanLatVertexRotatedOverRadius[i] = 0.0;
    }
}

/* seaice_cell_vertices_at_vertex (src/shared/mpas_seaice_mesh.F:632-685).  nEdgesOnCell of the junk
 * slot must be 0. */
void orc_cell_vertices_at_vertex(int *cellVerticesAtVertex, int nVertices, int vertexDegree, int maxEdges,
                                 const int *nEdgesOnCell, const int *verticesOnCell, const int *cellsOnVertex)
{
    const int M = maxEdges, D = vertexDegree;
    for (int iVertex = 1; iVertex <= nVertices; iVertex++) {
        for (int k = 1; k <= D; k++) {
            cellVerticesAtVertex[IDX2(k, iVertex, D)] = 0;
            const int iCell = cellsOnVertex[IDX2(k, iVertex, D)];
            for (int j = 1; j <= nEdgesOnCell[iCell - 1]; j++) {
                if (verticesOnCell[IDX2(j, iCell, M)] == iVertex) cellVerticesAtVertex[IDX2(k, iVertex, D)] = j;
            }
        }
    }
}

/* interior_vertices (mesh.F:423-488) */
void orc_interior_vertices(int *interiorVertex, int nVerticesSolve, int vertexDegree, int nCells,
                           const int *cellsOnVertex)
{
    const int D = vertexDegree;
    for (int iVertex = 1; iVertex <= nVerticesSolve; iVertex++) {
        interiorVertex[iVertex - 1] = 0;
        int nInterior = 0;
        for (int k = 1; k <= D; k++) {
            const int iCell = cellsOnVertex[IDX2(k, iVertex, D)];
            if (iCell >= 1 && iCell <= nCells) nInterior++;
        }
        if (nInterior == D) interiorVertex[iVertex - 1] = 1;
    }
}

/* local_eastern_and_northern_unit_vectors + seaice_project_3D_vector_onto_local_2D (mesh.F:2021-2332) */
static void project_3D_vector_onto_local_2D(double out[2], const double vec[3], double xP, double yP, double zP)
{
    double e[3], nrt[3], mag;
    e[0] = -yP; e[1] = xP; e[2] = 0.0;
    mag = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    e[0] = e[0] / mag; e[1] = e[1] / mag; e[2] = e[2] / mag;
    if (zP != 0.0) {
        nrt[0] = -xP; nrt[1] = -yP; nrt[2] = (xP * xP + yP * yP) / zP;
        mag = sqrt(nrt[0] * nrt[0] + nrt[1] * nrt[1] + nrt[2] * nrt[2]);
        nrt[0] = nrt[0] / mag; nrt[1] = nrt[1] / mag; nrt[2] = nrt[2] / mag;
        if (zP < 0.0) { nrt[0] = -nrt[0]; nrt[1] = -nrt[1]; nrt[2] = -nrt[2]; }
    } else {
        nrt[0] = 0.0; nrt[1] = 0.0; nrt[2] = 1.0;
    }
    /* seaice_dot_product_3space: x1*x2 + y1*y2 + z1*z2 */
    out[0] = vec[0] * e[0] + vec[1] * e[1] + vec[2] * e[2];
    out[1] = vec[0] * nrt[0] + vec[1] * nrt[1] + vec[2] * nrt[2];
}

/* seaice_calc_local_coords (variational_shared.F:42-279) */
void orc_calc_local_coords(double *xLocal, double *yLocal, int nCells, int maxEdges,
                           const int *nEdgesOnCell, const int *verticesOnCell,
                           const double *xVertex, const double *yVertex, const double *zVertex,
                           const double *xCell, const double *yCell, const double *zCell,
                           int rotateCartesianGrid, int onASphere)
{
    const int M = maxEdges;
#pragma omp parallel for schedule(static)
    for (int iCell = 1; iCell <= nCells; iCell++) {
        if (onASphere) {
            double xc, yc, zc;
            grid_rotation_forward(&xc, &yc, &zc, xCell[iCell - 1], yCell[iCell - 1], zCell[iCell - 1], rotateCartesianGrid);
            for (int j = 1; j <= nEdgesOnCell[iCell - 1]; j++) {
                const int iVertex = verticesOnCell[IDX2(j, iCell, M)];
                double v3[3], v2[2];
                grid_rotation_forward(&v3[0], &v3[1], &v3[2], xVertex[iVertex - 1], yVertex[iVertex - 1],
                                      zVertex[iVertex - 1], rotateCartesianGrid);
                project_3D_vector_onto_local_2D(v2, v3, xc, yc, zc);
                xLocal[IDX2(j, iCell, M)] = v2[0];
                yLocal[IDX2(j, iCell, M)] = v2[1];
            }
        } else {
            for (int j = 1; j <= nEdgesOnCell[iCell - 1]; j++) {
                const int iVertex = verticesOnCell[IDX2(j, iCell, M)];
                xLocal[IDX2(j, iCell, M)] = xVertex[iVertex - 1] - xCell[iCell - 1];
                yLocal[IDX2(j, iCell, M)] = yVertex[iVertex - 1] - yCell[iCell - 1];
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * quadrature rules (src/shared/mpas_seaice_velocity_solver_wachspress.F:1224-1712).
 * The literals are the reference's TRUNCATED literals, digit for digit:
 * Dunavant (1985) symmetric Gaussian rules for the triangle.
 * ------------------------------------------------------------------------------------------ */
#define QMAX 512
typedef struct { int n; double u[QMAX], v[QMAX], w[QMAX]; double norm; } quad_rule;

/* dunavant 9, 10, 12 and the 'fekete' rules (wachspress.F:1599-1941): generated by executing the reference's routines */
#include "quadrature_tables.inc"

static int rule_dunavant(int order, quad_rule *q)
{
    q->norm = 2.0;
#define SETQ(N, ...) do { static const double t_[] = { __VA_ARGS__ }; q->n = (N); \
        for (int i_ = 0; i_ < (N); i_++) { q->u[i_] = t_[i_]; q->v[i_] = t_[(N) + i_]; q->w[i_] = t_[2 * (N) + i_]; } } while (0)
    switch (order) {
    case 1: SETQ(1, 0.33333333333333, 0.33333333333333, 1.00000000000000); break;
    case 2: SETQ(3, 0.16666666666667, 0.16666666666667, 0.66666666666667,
                    0.16666666666667, 0.66666666666667, 0.16666666666667,
                    0.33333333333333, 0.33333333333333, 0.33333333333333); break;
    case 3: SETQ(4, 0.33333333333333, 0.20000000000000, 0.20000000000000, 0.60000000000000,
                    0.33333333333333, 0.20000000000000, 0.60000000000000, 0.20000000000000,
                    -0.56250000000000, 0.52083333333333, 0.52083333333333, 0.52083333333333); break;
    case 4: SETQ(6, 0.44594849091597, 0.44594849091597, 0.10810301816807, 0.09157621350977, 0.09157621350977, 0.81684757298046,
                    0.44594849091597, 0.10810301816807, 0.44594849091597, 0.09157621350977, 0.81684757298046, 0.09157621350977,
                    0.22338158967801, 0.22338158967801, 0.22338158967801, 0.10995174365532, 0.10995174365532, 0.10995174365532); break;
    case 5: SETQ(7, 0.33333333333333, 0.47014206410511, 0.47014206410511, 0.05971587178977, 0.10128650732346, 0.10128650732346, 0.79742698535309,
                    0.33333333333333, 0.47014206410511, 0.05971587178977, 0.47014206410511, 0.10128650732346, 0.79742698535309, 0.10128650732346,
                    0.22500000000000, 0.13239415278851, 0.13239415278851, 0.13239415278851, 0.12593918054483, 0.12593918054483, 0.12593918054483); break;
    case 6: SETQ(12, 0.24928674517091, 0.24928674517091, 0.50142650965818, 0.06308901449150, 0.06308901449150, 0.87382197101700,
                     0.31035245103378, 0.63650249912140, 0.05314504984482, 0.63650249912140, 0.31035245103378, 0.05314504984482,
                     0.24928674517091, 0.50142650965818, 0.24928674517091, 0.06308901449150, 0.87382197101700, 0.06308901449150,
                     0.63650249912140, 0.05314504984482, 0.31035245103378, 0.31035245103378, 0.05314504984482, 0.63650249912140,
                     0.11678627572638, 0.11678627572638, 0.11678627572638, 0.05084490637021, 0.05084490637021, 0.05084490637021,
                     0.08285107561837, 0.08285107561837, 0.08285107561837, 0.08285107561837, 0.08285107561837, 0.08285107561837); break;
    case 7: SETQ(13, 0.33333333333333, 0.26034596607904, 0.26034596607904, 0.47930806784192, 0.06513010290222, 0.06513010290222, 0.86973979419557,
                     0.31286549600487, 0.63844418856981, 0.04869031542532, 0.63844418856981, 0.31286549600487, 0.04869031542532,
                     0.33333333333333, 0.26034596607904, 0.47930806784192, 0.26034596607904, 0.06513010290222, 0.86973979419557, 0.06513010290222,
                     0.63844418856981, 0.04869031542532, 0.31286549600487, 0.31286549600487, 0.04869031542532, 0.63844418856981,
                     -0.14957004446768, 0.17561525743321, 0.17561525743321, 0.17561525743321, 0.05334723560884, 0.05334723560884, 0.05334723560884,
                     0.07711376089026, 0.07711376089026, 0.07711376089026, 0.07711376089026, 0.07711376089026, 0.07711376089026); break;
    case 8: SETQ(16, 0.33333333333333, 0.45929258829272, 0.45929258829272, 0.08141482341455, 0.17056930775176, 0.17056930775176, 0.65886138449648, 0.05054722831703,
                     0.05054722831703, 0.89890554336594, 0.26311282963464, 0.72849239295540, 0.00839477740996, 0.72849239295540, 0.26311282963464, 0.00839477740996,
                     0.33333333333333, 0.45929258829272, 0.08141482341455, 0.45929258829272, 0.17056930775176, 0.65886138449648, 0.17056930775176, 0.05054722831703,
                     0.89890554336594, 0.05054722831703, 0.72849239295540, 0.00839477740996, 0.26311282963464, 0.26311282963464, 0.00839477740996, 0.72849239295540,
                     0.14431560767779, 0.09509163426728, 0.09509163426728, 0.09509163426728, 0.10321737053472, 0.10321737053472, 0.10321737053472, 0.03245849762320,
                     0.03245849762320, 0.03245849762320, 0.02723031417443, 0.02723031417443, 0.02723031417443, 0.02723031417443, 0.02723031417443, 0.02723031417443); break;
    default: return rule_generated(0, order, q);       /* orders 9, 10, 12 */
    }
#undef SETQ
    return 0;
}

/* get_integration_factors_trapezoidal (wachspress.F:1301-1387) */
static int rule_trapezoidal(int order, quad_rule *q)
{
    const int nT = order;
    const int npts = ((nT + 1) * (nT + 1) + (nT + 1)) / 2;
    if (npts > QMAX || nT < 1) return 1;
    q->n = npts;
    int ij = 0;
    for (int i = 0; i <= nT; i++) {
        for (int j = 0; j <= nT - i; j++) {
            q->u[ij] = (double)i / (double)nT;
            q->v[ij] = (double)j / (double)nT;
            double w = 0.0;
            if (i <= nT - j) {
                if (i == nT || j == nT || (i == 0 && j == 0)) w = 1.0;
                else if ((j == 0 && i != 0 && i != nT) || (i == 0 && j != 0 && j != nT) ||
                         (i == nT - j && i != 0 && j != 0)) w = 3.0;
                else w = 6.0;
            }
            q->w[ij] = w;
            ij++;
        }
    }
    q->norm = 6.0 * ((double)nT * (double)nT);
    return 0;
}

/* 0 = dunavant, 1 = trapezoidal, 2 = fekete */
int orc_get_integration_factors(int integrationType, int integrationOrder, int *n, double *u, double *v,
                                double *w, double *norm)
{
    quad_rule q;
    int err = integrationType == 0 ? rule_dunavant(integrationOrder, &q)
            : integrationType == 1 ? rule_trapezoidal(integrationOrder, &q)
            : integrationType == 2 ? rule_generated(2, integrationOrder, &q) : 1;
    if (err) return err;
    *n = q.n; *norm = q.norm;
    memcpy(u, q.u, sizeof(double) * q.n);
    memcpy(v, q.v, sizeof(double) * q.n);
    memcpy(w, q.w, sizeof(double) * q.n);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Wachspress basis (wachspress.F:535-1206)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int n;
    double x[MAXE], y[MAXE], A[MAXE], B[MAXE], kappa[MAXE];
    int nsub[MAXE], sub[MAXE][MAXE];
} wcell;

/* calc_wachspress_coefficients (:535-614) + wachspress_indexes (:628-668), one cell.
 * kappa(j,i) of the reference is identical for every i, so one column is kept. */
static void wachspress_cell_setup(wcell *c, int n, const double *xl, const double *yl)
{
    c->n = n;
    for (int i = 0; i < n; i++) { c->x[i] = xl[i]; c->y[i] = yl[i]; }
    for (int iVertex = 1; iVertex <= n; iVertex++) {
        int i1 = iVertex - 1, i2 = iVertex;
        if (i1 < 1) i1 = i1 + n;
        const double den = c->x[i1 - 1] * c->y[i2 - 1] - c->x[i2 - 1] * c->y[i1 - 1];
        c->A[iVertex - 1] = (c->y[i2 - 1] - c->y[i1 - 1]) / den;
        c->B[iVertex - 1] = (c->x[i1 - 1] - c->x[i2 - 1]) / den;
    }
    c->kappa[0] = 1.0;
    for (int j = 2; j <= n; j++) {
        int i0 = j - 1, i1 = j, i2 = j + 1;
        if (i2 > n) i2 = i2 - n;
        c->kappa[j - 1] = c->kappa[j - 2] *
            (c->A[i2 - 1] * (c->x[i0 - 1] - c->x[i1 - 1]) + c->B[i2 - 1] * (c->y[i0 - 1] - c->y[i1 - 1])) /
            (c->A[i0 - 1] * (c->x[i1 - 1] - c->x[i0 - 1]) + c->B[i0 - 1] * (c->y[i1 - 1] - c->y[i0 - 1]));
    }
    for (int j = 1; j <= n; j++) {
        const int i1 = j, i2 = wrapped_index(j + 1, n);
        c->nsub[j - 1] = 0;
        for (int k = 1; k <= n; k++)
            if (k != i1 && k != i2) c->sub[j - 1][c->nsub[j - 1]++] = k;
    }
}

/* wachspress_edge_equation (:1049-1069) */
static inline double edge_equation(const wcell *c, int k, double x, double y)
{
    return 1.0 - c->A[k - 1] * x - c->B[k - 1] * y;
}

/* At one point: all numerators (wachspress_numerator :864-925), their derivatives
 * (wachspress_numerator_derivative :939-1035), denominator and derivative sums, then every basis
 * function (:682-749) and basis derivative (:763-850).  The reference recomputes these per basis
 * index; the values are identical because kappa(j,i) does not depend on i. */
static void wachspress_eval_all(const wcell *c, double x, double y, double *phi, double *dphix, double *dphiy)
{
    const int n = c->n;
    double num[MAXE], dnx[MAXE], dny[MAXE];
    double denominator = 0.0, sdx = 0.0, sdy = 0.0;
    for (int j = 1; j <= n; j++) {
        const int ns = c->nsub[j - 1];
        double numerator = 1.0;
        for (int k = 0; k < ns; k++) numerator = numerator * edge_equation(c, c->sub[j - 1][k], x, y);
        numerator = numerator * c->kappa[j - 1];
        num[j - 1] = numerator;
        denominator = denominator + numerator;
        double spx = 0.0, spy = 0.0;
        for (int k = 0; k < ns; k++) {
            double px = 1.0, py = 1.0;
            for (int l = 0; l < k; l++) {
                const double e = edge_equation(c, c->sub[j - 1][l], x, y);
                px = px * e; py = py * e;
            }
            px = px * (-c->A[c->sub[j - 1][k] - 1]);
            py = py * (-c->B[c->sub[j - 1][k] - 1]);
            for (int l = k + 1; l < ns; l++) {
                const double e = edge_equation(c, c->sub[j - 1][l], x, y);
                px = px * e; py = py * e;
            }
            spx = spx + px; spy = spy + py;
        }
        dnx[j - 1] = spx * c->kappa[j - 1];
        dny[j - 1] = spy * c->kappa[j - 1];
        sdx = sdx + dnx[j - 1];
        sdy = sdy + dny[j - 1];
    }
    for (int i = 0; i < n; i++) {
        if (phi) phi[i] = num[i] / denominator;
        if (dphix) {
            dphix[i] = dnx[i] / denominator - (num[i] / (denominator * denominator)) * sdx;
            dphiy[i] = dny[i] / denominator - (num[i] / (denominator * denominator)) * sdy;
        }
    }
}

/* seaice_init_velocity_solver_wachspress (:46-161): calculate_wachspress_derivatives (:1083-1206)
 * and integrate_wachspress (:179-467).  Entries with an index beyond nEdgesOnCell are left untouched
 * (the caller zero-fills, like MPAS pool allocation does). */
int orc_init_velocity_solver_wachspress(int nCells, int maxEdges, const int *nEdgesOnCell,
                                        const double *xLocal, const double *yLocal,
                                        int integrationType, int integrationOrder,
                                        double *basisGradientU, double *basisGradientV,
                                        double *basisIntegralsU, double *basisIntegralsV,
                                        double *basisIntegralsMetric)
{
    const int M = maxEdges;
    if (M > MAXE) return 2;
    quad_rule q;
    int err = integrationType == 0 ? rule_dunavant(integrationOrder, &q)
            : integrationType == 1 ? rule_trapezoidal(integrationOrder, &q)
            : integrationType == 2 ? rule_generated(2, integrationOrder, &q) : 1;
    if (err) return err;
    const int nq = q.n;

#pragma omp parallel
    {
        double *phi = (double *)malloc(sizeof(double) * MAXE * MAXE * nq);
        double *dpx = (double *)malloc(sizeof(double) * MAXE * MAXE * nq);
        double *dpy = (double *)malloc(sizeof(double) * MAXE * MAXE * nq);
#pragma omp for schedule(static)
        for (int iCell = 1; iCell <= nCells; iCell++) {
            const int n = nEdgesOnCell[iCell - 1];
            if (n < 3) continue;
            wcell c;
            wachspress_cell_setup(&c, n, &xLocal[IDX2(1, iCell, M)], &yLocal[IDX2(1, iCell, M)]);

            /* gradients at the cell vertices; only i-1, i, i+1 are kept (:1178-1191) */
            for (int iBasis = 1; iBasis <= n; iBasis++)
                for (int j = 1; j <= M; j++) {
                    basisGradientU[IDX3(iBasis, j, iCell, M)] = 0.0;
                    basisGradientV[IDX3(iBasis, j, iCell, M)] = 0.0;
                }
            for (int iGrad = 1; iGrad <= n; iGrad++) {
                double dx[MAXE], dy[MAXE];
                wachspress_eval_all(&c, c.x[iGrad - 1], c.y[iGrad - 1], NULL, dx, dy);
                for (int iBasis = 1; iBasis <= n; iBasis++) {
                    if (iGrad == iBasis || iGrad == wrapped_index(iBasis - 1, n) || iGrad == wrapped_index(iBasis + 1, n)) {
                        basisGradientU[IDX3(iBasis, iGrad, iCell, M)] = dx[iBasis - 1];
                        basisGradientV[IDX3(iBasis, iGrad, iCell, M)] = dy[iBasis - 1];
                    }
                }
            }

            /* basis values at every quadrature point of every sub-triangle */
            double jacobian[MAXE];
            for (int s = 1; s <= n; s++) {
                const int i1 = s, i2 = wrapped_index(s + 1, n);
                /* get_triangle_mapping (:485-517) with (x1,y1)=(1,0), (x2,y2)=(0,1) */
                const double x1 = 1.0, y1 = 0.0, x2 = 0.0, y2 = 1.0;
                const double u1 = c.x[i1 - 1], v1 = c.y[i1 - 1], u2 = c.x[i2 - 1], v2 = c.y[i2 - 1];
                const double m11 = (u2 * y1 - u1 * y2) / (x2 * y1 - x1 * y2);
                const double m12 = (u1 * x2 - u2 * x1) / (y1 * x2 - y2 * x1);
                const double m21 = (v2 * y1 - v1 * y2) / (x2 * y1 - x1 * y2);
                const double m22 = (v1 * x2 - v2 * x1) / (y1 * x2 - y2 * x1);
                jacobian[s - 1] = m11 * m22 - m12 * m21;
                for (int p = 0; p < nq; p++) {
                    const double x = m11 * q.u[p] + m12 * q.v[p];
                    const double y = m21 * q.u[p] + m22 * q.v[p];
                    const size_t o = ((size_t)(s - 1) * nq + p) * MAXE;
                    wachspress_eval_all(&c, x, y, &phi[o], &dpx[o], &dpy[o]);
                }
            }
            /* integrate_wachspress_polygon (:304-467) for every (iStressVertex, iVelocityVertex) */
            for (int iVel = 1; iVel <= n; iVel++) {
                for (int iStr = 1; iStr <= n; iStr++) {
                    double bU = 0.0, bV = 0.0, bM = 0.0;
                    for (int s = 1; s <= n; s++) {
                        double sU = 0.0, sV = 0.0, sM = 0.0;
                        for (int p = 0; p < nq; p++) {
                            const size_t o = ((size_t)(s - 1) * nq + p) * MAXE;
                            const double tmp = jacobian[s - 1] * q.w[p] * phi[o + iStr - 1];
                            sU = sU + tmp * dpx[o + iVel - 1];
                            sV = sV + tmp * dpy[o + iVel - 1];
                            sM = sM + tmp * phi[o + iVel - 1];
                        }
                        bU = bU + sU / q.norm;
                        bV = bV + sV / q.norm;
                        bM = bM + sM / q.norm;
                    }
                    basisIntegralsU[IDX3(iStr, iVel, iCell, M)] = bU;
                    basisIntegralsV[IDX3(iStr, iVel, iCell, M)] = bV;
                    basisIntegralsMetric[IDX3(iStr, iVel, iCell, M)] = bM;
                }
            }
        }
        free(phi); free(dpx); free(dpy);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * PWL basis (src/shared/mpas_seaice_velocity_solver_pwl.F:44-373) with the dense LU solve of
 * src/shared/mpas_seaice_numerics.F:44-212 (Crout, implicit scaling, partial pivoting).
 * ------------------------------------------------------------------------------------------ */
static void lu_decomposition(double a[3][3], int indices[3])
{
    const int n = 3;
    const double tiny = 1.0e-20;
    double maxa[3];
    for (int i = 0; i < n; i++) {
        double m = 0.0;
        for (int j = 0; j < n; j++) if (fabs(a[i][j]) > m) m = fabs(a[i][j]);
        maxa[i] = 1.0 / m;
    }
    for (int j = 0; j < n; j++) {
        int jmax = j;
        double best = maxa[j] * fabs(a[j][j]);
        for (int i = j + 1; i < n; i++) {          /* maxloc: first maximum */
            const double val = maxa[i] * fabs(a[i][j]);
            if (val > best) { best = val; jmax = i; }
        }
        if (j != jmax) {
            for (int k = 0; k < n; k++) { const double t = a[jmax][k]; a[jmax][k] = a[j][k]; a[j][k] = t; }
            maxa[jmax] = maxa[j];
        }
        indices[j] = jmax;
        if (a[j][j] == 0.0) a[j][j] = tiny;
        for (int i = j + 1; i < n; i++) a[i][j] = a[i][j] / a[j][j];
        for (int i = j + 1; i < n; i++)
            for (int k = j + 1; k < n; k++) a[i][k] = a[i][k] - a[i][j] * a[j][k];
    }
}

static void lu_back_substitution(double a[3][3], const int indices[3], double b[3])
{
    const int n = 3;
    int j = -1;
    for (int i = 0; i < n; i++) {
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
    for (int i = n - 1; i >= 0; i--) {
        double dot = 0.0;
        for (int l = i + 1; l < n; l++) dot = dot + a[i][l] * b[l];
        b[i] = (b[i] - dot) / a[i][i];
    }
}

static void solve_linear_basis_system(const double left[3][3], const double rhs[3], double sol[3])
{
    double a[3][3];
    int indices[3];
    memcpy(a, left, sizeof(a));
    memcpy(sol, rhs, 3 * sizeof(double));
    lu_decomposition(a, indices);
    lu_back_substitution(a, indices, sol);
}

int orc_init_velocity_solver_pwl(int nCells, int maxEdges, const int *nEdgesOnCell, const int *edgesOnCell,
                                 const double *dvEdge, const double *areaCell,
                                 const double *xLocal, const double *yLocal,
                                 double *basisGradientU, double *basisGradientV,
                                 double *basisIntegralsMetric, double *basisIntegralsU, double *basisIntegralsV)
{
    const int M = maxEdges;
    if (M > MAXE) return 2;
#pragma omp parallel for schedule(static)
    for (int iCell = 1; iCell <= nCells; iCell++) {
        const int n = nEdgesOnCell[iCell - 1];
        if (n < 3) continue;
        const double *xl = &xLocal[IDX2(1, iCell, M)], *yl = &yLocal[IDX2(1, iCell, M)];
        const double alphaPWL = 1.0 / (double)n;
        double xC = 0.0, yC = 0.0;
        for (int j = 0; j < n; j++) { xC = xC + alphaPWL * xl[j]; yC = yC + alphaPWL * yl[j]; }
        double basisSubArea[MAXE], basisSubAreaSum = 0.0;
        for (int s = 1; s <= n; s++) {
            const int iEdge = edgesOnCell[IDX2(s, iCell, M)];
            const int v1 = s, v2 = wrapped_index(s + 1, n);
            const double c = dvEdge[iEdge - 1];
            const double a = sqrt((xl[v1 - 1] - xC) * (xl[v1 - 1] - xC) + (yl[v1 - 1] - yC) * (yl[v1 - 1] - yC));
            const double b = sqrt((xl[v2 - 1] - xC) * (xl[v2 - 1] - xC) + (yl[v2 - 1] - yC) * (yl[v2 - 1] - yC));
            const double sp = (a + b + c) * 0.5;
            basisSubArea[s - 1] = sqrt(sp * (sp - a) * (sp - b) * (sp - c));
            basisSubAreaSum = basisSubAreaSum + basisSubArea[s - 1];
        }
        {
            const double scale = areaCell[iCell - 1] / basisSubAreaSum;
            for (int s = 0; s < n; s++) basisSubArea[s] = basisSubArea[s] * scale;
        }
        double sbU[MAXE][3], sbV[MAXE][3];
        for (int s = 1; s <= n; s++) {
            const int v1 = s, v2 = wrapped_index(s + 1, n);
            double left[3][3] = {{xl[v1 - 1] - xC, yl[v1 - 1] - yC, 1.0},
                                 {xl[v2 - 1] - xC, yl[v2 - 1] - yC, 1.0},
                                 {0.0, 0.0, 1.0}};
            double rhs1[3] = {1.0, 0.0, 0.0}, rhs2[3] = {0.0, 1.0, 0.0}, sol[3];
            solve_linear_basis_system(left, rhs1, sol);
            sbU[s - 1][0] = sol[0]; sbV[s - 1][0] = sol[1];
            solve_linear_basis_system(left, rhs2, sol);
            sbU[s - 1][1] = sol[0]; sbV[s - 1][1] = sol[1];
            sbU[s - 1][2] = -sbU[s - 1][0] - sbU[s - 1][1];
            sbV[s - 1][2] = -sbV[s - 1][0] - sbV[s - 1][1];
        }
        double scU[MAXE][MAXE], scV[MAXE][MAXE]; /* [iBasisVertex][iSubCell] */
        for (int ib = 1; ib <= n; ib++) {
            for (int s = 1; s <= n; s++) {
                scU[ib - 1][s - 1] = sbU[s - 1][2] * alphaPWL;
                scV[ib - 1][s - 1] = sbV[s - 1][2] * alphaPWL;
                if (s == ib) {
                    scU[ib - 1][s - 1] = scU[ib - 1][s - 1] + sbU[s - 1][0];
                    scV[ib - 1][s - 1] = scV[ib - 1][s - 1] + sbV[s - 1][0];
                } else if (s == wrapped_index(ib - 1, n)) {
                    scU[ib - 1][s - 1] = scU[ib - 1][s - 1] + sbU[s - 1][1];
                    scV[ib - 1][s - 1] = scV[ib - 1][s - 1] + sbV[s - 1][1];
                }
            }
        }
        for (int ib = 1; ib <= n; ib++) {
            for (int ig = 1; ig <= n; ig++) {
                const int s1 = ig, s2 = wrapped_index(ig - 1, n);
                basisGradientU[IDX3(ib, ig, iCell, M)] = 0.5 * (scU[ib - 1][s1 - 1] + scU[ib - 1][s2 - 1]);
                basisGradientV[IDX3(ib, ig, iCell, M)] = 0.5 * (scV[ib - 1][s1 - 1] + scV[ib - 1][s2 - 1]);
            }
        }
        for (int is = 1; is <= n; is++) {
            for (int iv = 1; iv <= n; iv++) {
                double bU = 0.0, bV = 0.0;
                for (int s = 1; s <= n; s++) {
                    double basisIntegral;
                    if (s == is || s == wrapped_index(is - 1, n))
                        basisIntegral = ((alphaPWL + 1) * basisSubArea[s - 1]) / 3.0;
                    else
                        basisIntegral = (alphaPWL * basisSubArea[s - 1]) / 3.0;
                    bU = bU + scU[iv - 1][s - 1] * basisIntegral;
                    bV = bV + scV[iv - 1][s - 1] * basisIntegral;
                }
                basisIntegralsU[IDX3(is, iv, iCell, M)] = bU;
                basisIntegralsV[IDX3(is, iv, iCell, M)] = bV;
            }
        }
        for (int is = 1; is <= n; is++) {
            for (int iv = 1; iv <= n; iv++) {
                double bM = 0.0;
                for (int s = 1; s <= n; s++) {
                    const int tS = (s == is) ? 1 : (s == wrapped_index(is - 1, n)) ? 2 : 3;
                    const int tV = (s == iv) ? 1 : (s == wrapped_index(iv - 1, n)) ? 2 : 3;
                    double val = 0.0;
                    if ((tS == 1 && tV == 1) || (tS == 2 && tV == 2))
                        val = 2.0 * (alphaPWL * alphaPWL) + 2.0 * alphaPWL + 2.0;
                    else if ((tS == 1 && tV == 2) || (tS == 2 && tV == 1))
                        val = 2.0 * (alphaPWL * alphaPWL) + 2.0 * alphaPWL + 1.0;
                    else if ((tS == 1 && tV == 3) || (tS == 3 && tV == 1) || (tS == 2 && tV == 3) || (tS == 3 && tV == 2))
                        val = 2.0 * (alphaPWL * alphaPWL) + alphaPWL;
                    else if (tS == 3 && tV == 3)
                        val = 2.0 * (alphaPWL * alphaPWL);
                    val = val * basisSubArea[s - 1] / 12.0;
                    bM = bM + val;
                }
                basisIntegralsMetric[IDX3(is, iv, iCell, M)] = bM;
            }
        }
    }
    return 0;
}

/* variational_denominator (src/shared/mpas_seaice_velocity_solver_variational.F:358-445);
 * type 0 = 'original', 1 = 'alternate' */
void orc_variational_denominator(int nVertices, int vertexDegree, int maxEdges, const int *nEdgesOnCell,
                                 const double *areaTriangle, const int *cellsOnVertex,
                                 const int *cellVerticesAtVertex, const double *basisIntegralsMetric,
                                 int denominatorType, double *variationalDenominator)
{
    const int M = maxEdges, D = vertexDegree;
    if (denominatorType == 1) {
        for (int iVertex = 1; iVertex <= nVertices; iVertex++) {
            double acc = 0.0;
            for (int k = 1; k <= D; k++) {
                const int iCell = cellsOnVertex[IDX2(k, iVertex, D)];
                const int iVel = cellVerticesAtVertex[IDX2(k, iVertex, D)];
                for (int is = 1; is <= nEdgesOnCell[iCell - 1]; is++)
                    acc = acc + basisIntegralsMetric[IDX3(is, iVel, iCell, M)];
            }
            variationalDenominator[iVertex - 1] = acc;
        }
    } else {
        for (int i = 0; i < nVertices; i++) variationalDenominator[i] = areaTriangle[i];
    }
}
