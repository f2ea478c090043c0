/* oracle/ir_oracle.c -- CPU restatement of the incremental-remapping (IR) transport of MPAS-Seaice,
 * the consumer of the EVP velocities.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may call this;
 * the product never does.
 *
 * Restates, in the reference's operation order (build with -ffp-contract=off):
 *   src/shared/mpas_seaice_advection_incremental_remap.F
 *     define_local_to_global_transformations :990-1090    get_geometry_incremental_remap :1105-1810
 *     get_vertex_on_cell_coordinates        :1823-2040    compute_geometric_cell_averages :2097-2330
 *     seaice_run_advection_incremental_remap :2338-2730   incremental_remap_block        :2740-3400
 *     make_masks :3404-3570   construct_linear_tracer_fields :3580-4200   compute_gradient_2d/3d :4204-4650
 *     compute_barycenter_coordinates :4658-4800   limit_tracer_gradient_2d/3d :4802-5250
 *     find_departure_points :5255-5360   find_departure_triangles :5365-6265
 *     shift_vertices_of_departure_triangle :6270-6545   get_triangle_quadrature_points :6546-6665
 *     integrate_fluxes_over_triangles :6667-6980   compute_mass_tracer_products :6982-7120
 *     update_mass_and_tracers :7125-7540   zap_small_mass :8764-8895   helpers :8900-9326
 *     the optional checks (config_conservation_check, config_monotonicity_check; default off):
 *     sum_tracers :7998-8120   check_tracer_conservation :8126-8260   tracer_local_min_max :8268-8410
 *     check_tracer_monotonicity :8416-8760
 *
 * The reference keeps "2D" (nCategories, nCells) and "3D" (nLayers, nCategories, nCells) tracers in separate
 * code paths that differ only in the layer index; here every tracer is (nLayers, nCategories, nCells) with
 * nLayers = 1 for the 2D ones, and a parent with one layer is addressed with layer 1 whatever the child's layer
 * (the reference's "parentTracer % ndims == 2" branches).
 *
 * coeffs_reconstruct (the gradient reconstruction coefficients) comes from the MPAS framework
 * (mpas_init_reconstruct, an un-vendored dependency: /root/reference/src/Makefile:6-7) and is an INPUT here, as it
 * is for the reference (incremental_remap.F:744-746).
 *
 * Parity status: the reference cannot be built in this image (no Fortran compiler, no MPAS framework) and its
 * advection test case (testing_and_setup/testcases/advection) stores no numbers, only plots.  This file is pinned by
 * OUTPUTS OF THE REFERENCE'S OWN SOURCE EXECUTED HERE: the Fortran-subset interpreter tests/golden/fortran_subset.py runs
 * seaice_init_advection_incremental_remap and seaice_run_advection_incremental_remap (with incremental_remap_block and
 * everything below it, the optional checks included) from the file under /root/reference; the fixtures
 * (tests/golden/ir/refexec_ir*.npz: the full tracer hierarchy with layers on planar hexagons and quadrilaterals and the sphere,
 * rotated and not) are reproduced by this file bit for bit (tests/test_ir_parity.py) -- every tracer, the conservation
 * sums, the abort decisions of both checks, the whole geometry pool.  tests/test_oracle_ir.py adds the properties the
 * scheme guarantees (conservation, monotonicity, uniform fields, exact translation of linear fields, an independent
 * exact remap) and the reference's own test case (solid-body rotation on the sphere, second-order convergence).
 *
 * Array conventions (as everywhere in this repo): 1-based index VALUES, numpy C-order == Fortran column-major with
 * the dimensions reversed, a junk slot n+1 at the end of cell / vertex / edge arrays.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define EPS11 1.0e-11
#define W1QP 1.09951743655321885e-01
#define W2QP 2.23381589678011389e-01
#define Q1QP 9.15762135097710761e-02
#define Q2QP 8.16847572980458514e-01
#define Q3QP 1.08103018168070275e-01
#define Q4QP 4.45948490915965612e-01

#define NTRI 6  /* nTriPerEdgeRemap,        Registry.xml:63-66 */
#define NCER 6  /* maxCellsPerEdgeRemap,    Registry.xml:67-70 */
#define NEER 6  /* maxEdgesPerEdgeRemap,    Registry.xml:71-74 */
#define NVER 8  /* maxVerticesPerEdgeRemap, Registry.xml:75-78 */

/* error codes */
#define IR_OK 0
#define IR_ERR_EDGE_ORIENTATION 1   /* cellsOnEdge(1) not left of V1->V2 (incremental_remap.F:1290-1330) */
#define IR_ERR_CELL_ORIENTATION 2   /* cell vertices not counter-clockwise (incremental_remap.F:1960-2030) */
#define IR_ERR_PARALLEL_EDGES   3   /* shift_vertices: basis edges parallel (incremental_remap.F:6420) */
#define IR_ERR_NEGATIVE_MASS_QP 4   /* negative mass at a quadrature point (incremental_remap.F:6900) */
#define IR_ERR_NEGATIVE_MASS    5   /* new mass < -puny^2 (incremental_remap.F:7470) */
#define IR_ERR_TOO_MANY_PARENTS 6
#define IR_ERR_TOO_MANY_TRIANGLES 7
#define IR_ERR_BAD_ARGUMENT     8
#define IR_ERR_CONSERVATION     9   /* check_tracer_conservation (:8126): relative change of a global sum > eps11 */
#define IR_ERR_MONOTONICITY     10  /* check_tracer_monotonicity (:8416): new value outside the old neighbourhood range */

/* 1-based accessors */
#define A2(a, i, j, n1) ((a)[((size_t)(j) - 1) * (size_t)(n1) + ((size_t)(i) - 1)])            /* a(i,j), first dim n1 */
#define T33(t, i, j, c) ((t)[((size_t)(c) - 1) * 9 + ((size_t)(j) - 1) * 3 + ((size_t)(i) - 1)]) /* t(i,j,c) */

typedef struct {
    int nCells, nCellsSolve, nVertices, nEdges, maxEdges, vertexDegree;
    int on_a_sphere, rotate_cartesian_grid;
    const int *nEdgesOnCell, *edgesOnCell, *verticesOnCell, *cellsOnEdge, *verticesOnEdge, *edgesOnVertex;
    const double *xCell, *yCell, *zCell, *xVertex, *yVertex, *zVertex, *xEdge, *yEdge, *zEdge;
    const double *dcEdge, *dvEdge;
    /* outputs */
    double *transGlobalToCell;                 /* (nCells, 3, 3); untouched on a plane */
    double *xVertexOnCell, *yVertexOnCell;     /* (nCells+1, maxEdges) */
    int *remapEdge;                            /* (nEdges+1) */
    int *cellsOnEdgeRemap, *edgesOnEdgeRemap;  /* (nEdges+1, 6) */
    double *xVertexOnEdge, *yVertexOnEdge;     /* (nEdges+1, 8) */
    double *minLengthEdgesOnVertex;            /* (nVertices+1) */
    double *geomAvg[14];                       /* x y xx xy yy xxx xxy xyy yyy xxxx xxxy xxyy xyyy yyyy, (nCells+1) each */
} orc_ir_geometry_args;

/* ------------------------------------------------------------------------------------------------ helpers */

static void unit_vector_3d(double *v) /* :8900 */
{
    const double magnitude = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (magnitude > 0.0) { v[0] = v[0] / magnitude; v[1] = v[1] / magnitude; v[2] = v[2] / magnitude; }
    else { v[0] = v[1] = v[2] = 0.0; }
}

static void cross_product_3d(const double *a, const double *b, double *o) /* :9109 */
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

static double cross_product_2d(const double *a, const double *b) { return a[0] * b[1] - a[1] * b[0]; } /* :9083 */

/* point_in_half_plane (:9200): the dummy arguments are (point, lineStart, lineEnd); every caller passes
 * (edge vertex 1, edge vertex 2, the point tested), so the cross product evaluated is
 * (third - second) x (first - second).  Restated with the dummy-argument roles, i.e. as executed. */
static int point_in_half_plane(const double *point, const double *lineStart, const double *lineEnd)
{
    const double v1[2] = {lineEnd[0] - lineStart[0], lineEnd[1] - lineStart[1]};
    const double v2[2] = {point[0] - lineStart[0], point[1] - lineStart[1]};
    return cross_product_2d(v1, v2) >= 0.0;
}

static double triangle_area(const double *v1, const double *v2, const double *v3) /* :9172 */
{
    return fabs(0.5 * ((v2[0] - v1[0]) * (v3[1] - v1[1]) - (v2[1] - v1[1]) * (v3[0] - v1[0])));
}

static double quadrilateral_area(const double *v1, const double *v2, const double *v3, const double *v4) /* :9137 */
{
    const double a1 = triangle_area(v1, v2, v3), a2 = triangle_area(v1, v3, v4);
    return a1 + a2;
}

static int find_line_intersection(const double *p1, const double *p2, const double *p3, const double *p4, double *ip) /* :8934 */
{
    int lineIntersect = 0;
    const double x1 = p1[0], y1 = p1[1], x2 = p2[0], y2 = p2[1], x3 = p3[0], y3 = p3[1], x4 = p4[0], y4 = p4[1];
    const double rx = x2 - x1, ry = y2 - y1, sx = x4 - x3, sy = y4 - y3;
    const double rsCross = rx * sy - ry * sx;
    const double rsCrossMin = EPS11 * sqrt((rx * rx + ry * ry) * (sx * sx + sy * sy));
    if (fabs(rsCross) > rsCrossMin) {
        const double t1 = (sy * (x3 - x1) - sx * (y3 - y1)) / rsCross;
        const double t2 = (ry * (x3 - x1) - rx * (y3 - y1)) / rsCross;
        ip[0] = x1 + t1 * rx;
        ip[1] = y1 + t1 * ry;
        if (t1 > 0.0 && t1 < 1.0 && t2 > 0.0 && t2 < 1.0) lineIntersect = 1;
    } else {
        ip[0] = DBL_MAX;
        ip[1] = DBL_MAX;
    }
    return lineIntersect;
}

/* define_local_to_global_transformations (:990): transGlobalToLocal = matmul(transpose(unitVectorLocal), I), i.e.
 * row 1 = local east, row 2 = local north, row 3 = radial unit vector. */
static void global_to_local(double x, double y, double z, double *t /* t(i,j) at t[(j-1)*3+(i-1)] */)
{
    double e1[3], e2[3], e3[3] = {x, y, z};
    unit_vector_3d(e3);
    if (fabs(x * x + y * y) > EPS11) {
        e1[0] = -y; e1[1] = x; e1[2] = 0.0;
        unit_vector_3d(e1);
        cross_product_3d(e3, e1, e2);
    } else if (z > 0.0) {
        e1[0] = 1.0; e1[1] = 0.0; e1[2] = 0.0;
        e2[0] = 0.0; e2[1] = 1.0; e2[2] = 0.0;
    } else {
        e1[0] = 0.0; e1[1] = 1.0; e1[2] = 0.0;
        e2[0] = 1.0; e2[1] = 0.0; e2[2] = 0.0;
    }
    for (int j = 0; j < 3; j++) { t[j * 3 + 0] = e1[j]; t[j * 3 + 1] = e2[j]; t[j * 3 + 2] = e3[j]; }
}

static inline double sq(double x) { return x * x; }

/* ------------------------------------------------------------------------------------------ init geometry */

int orc_ir_init_geometry(orc_ir_geometry_args *g)
{
    const int nC = g->nCells, nV = g->nVertices, nE = g->nEdges, M = g->maxEdges, D = g->vertexDegree;
    if (D != 3 && D != 4) return IR_ERR_BAD_ARGUMENT;
    int err = IR_OK;
    const double *xC = g->xCell, *yC = g->yCell, *zC = g->zCell, *xV = g->xVertex, *yV = g->yVertex, *zV = g->zVertex;
    const double *xE = g->xEdge, *yE = g->yEdge, *zE = g->zEdge;
    double *rot = NULL, *tE = NULL;
    /* rotate_global_vectors (:948): (x, y, z) -> (-z, y, x) */
    if (g->rotate_cartesian_grid && g->on_a_sphere) {
        const size_t tot = 3 * ((size_t)nC + nV + nE);
        rot = (double *)malloc(sizeof(double) * (tot + 3));
        double *p = rot;
        const double *src[9] = {xC, yC, zC, xV, yV, zV, xE, yE, zE};
        const int cnt[3] = {nC, nV, nE};
        const double *dst[9];
        for (int k = 0; k < 3; k++) {
            double *rx = p, *ry = p + cnt[k], *rz = p + 2 * (size_t)cnt[k];
            for (int i = 0; i < cnt[k]; i++) { rx[i] = -src[3 * k + 2][i]; ry[i] = src[3 * k + 1][i]; rz[i] = src[3 * k][i]; }
            dst[3 * k] = rx; dst[3 * k + 1] = ry; dst[3 * k + 2] = rz;
            p += 3 * (size_t)cnt[k];
        }
        xC = dst[0]; yC = dst[1]; zC = dst[2]; xV = dst[3]; yV = dst[4]; zV = dst[5]; xE = dst[6]; yE = dst[7]; zE = dst[8];
    }
    if (g->on_a_sphere) {
        for (int c = 1; c <= nC; c++) global_to_local(xC[c - 1], yC[c - 1], zC[c - 1], &T33(g->transGlobalToCell, 1, 1, c));
        tE = (double *)malloc(sizeof(double) * 9 * (size_t)(nE > 0 ? nE : 1));
        for (int e = 1; e <= nE; e++) global_to_local(xE[e - 1], yE[e - 1], zE[e - 1], &T33(tE, 1, 1, e));
    }

    /* get_vertex_on_cell_coordinates (:1823) */
    memset(g->xVertexOnCell, 0, sizeof(double) * ((size_t)nC + 1) * M);
    memset(g->yVertexOnCell, 0, sizeof(double) * ((size_t)nC + 1) * M);
    for (int iCell = 1; iCell <= nC; iCell++) {
        for (int k = 1; k <= g->nEdgesOnCell[iCell - 1]; k++) {
            const int iVertex = A2(g->verticesOnCell, k, iCell, M);
            if (g->on_a_sphere) {
                const double w[3] = {xV[iVertex - 1] - xC[iCell - 1], yV[iVertex - 1] - yC[iCell - 1], zV[iVertex - 1] - zC[iCell - 1]};
                const double *t = g->transGlobalToCell;
                A2(g->xVertexOnCell, k, iCell, M) = T33(t, 1, 1, iCell) * w[0] + T33(t, 1, 2, iCell) * w[1] + T33(t, 1, 3, iCell) * w[2];
                A2(g->yVertexOnCell, k, iCell, M) = T33(t, 2, 1, iCell) * w[0] + T33(t, 2, 2, iCell) * w[1] + T33(t, 2, 3, iCell) * w[2];
            } else {
                A2(g->xVertexOnCell, k, iCell, M) = xV[iVertex - 1] - xC[iCell - 1];
                A2(g->yVertexOnCell, k, iCell, M) = yV[iVertex - 1] - yC[iCell - 1];
            }
        }
    }
    for (int iCell = 1; iCell <= nC; iCell++) {
        const int n = g->nEdgesOnCell[iCell - 1];
        for (int k = 1; k <= n; k++) {
            const int kp1 = (k + 1 > n) ? 1 : k + 1;
            const double a[2] = {A2(g->xVertexOnCell, k, iCell, M), A2(g->yVertexOnCell, k, iCell, M)};
            const double b[2] = {A2(g->xVertexOnCell, kp1, iCell, M), A2(g->yVertexOnCell, kp1, iCell, M)};
            if (!(cross_product_2d(a, b) >= 0.0)) err = IR_ERR_CELL_ORIENTATION;
        }
    }

    /* get_geometry_incremental_remap (:1105) */
    int *remapEdge = g->remapEdge;
    for (int e = 0; e <= nE; e++) remapEdge[e] = 0;
    for (int iCell = 1; iCell <= g->nCellsSolve; iCell++)
        for (int k = 1; k <= g->nEdgesOnCell[iCell - 1]; k++) {
            const int iEdge = A2(g->edgesOnCell, k, iCell, M);
            if (iEdge >= 1 && iEdge <= nE) remapEdge[iEdge - 1] = 1;
        }
    for (int iEdge = 1; iEdge <= nE; iEdge++)
        if (remapEdge[iEdge - 1] == 1)
            for (int k = 1; k <= 2; k++) {
                const int iCell = A2(g->cellsOnEdge, k, iEdge, 2);
                if (iCell < 1 || iCell > nC) { remapEdge[iEdge - 1] = 0; break; }
            }
    /* orientation check: C1 must lie in the left half-plane of V1 -> V2 */
    for (int iEdge = 1; iEdge <= nE; iEdge++) {
        if (remapEdge[iEdge - 1] != 1) continue;
        double ev[2][2], cc[2];
        const int iCell = A2(g->cellsOnEdge, 1, iEdge, 2);
        for (int k = 1; k <= 2; k++) {
            const int iVertex = A2(g->verticesOnEdge, k, iEdge, 2);
            if (g->on_a_sphere) {
                const double w[3] = {xV[iVertex - 1] - xE[iEdge - 1], yV[iVertex - 1] - yE[iEdge - 1], zV[iVertex - 1] - zE[iEdge - 1]};
                ev[k - 1][0] = T33(tE, 1, 1, iEdge) * w[0] + T33(tE, 1, 2, iEdge) * w[1] + T33(tE, 1, 3, iEdge) * w[2];
                ev[k - 1][1] = T33(tE, 2, 1, iEdge) * w[0] + T33(tE, 2, 2, iEdge) * w[1] + T33(tE, 2, 3, iEdge) * w[2];
            } else {
                ev[k - 1][0] = xV[iVertex - 1] - xE[iEdge - 1];
                ev[k - 1][1] = yV[iVertex - 1] - yE[iEdge - 1];
            }
        }
        if (iCell >= 1 && iCell <= nC) {
            if (g->on_a_sphere) {
                const double w[3] = {xC[iCell - 1] - xE[iEdge - 1], yC[iCell - 1] - yE[iEdge - 1], zC[iCell - 1] - zE[iEdge - 1]};
                cc[0] = T33(tE, 1, 1, iEdge) * w[0] + T33(tE, 1, 2, iEdge) * w[1] + T33(tE, 1, 3, iEdge) * w[2];
                cc[1] = T33(tE, 2, 1, iEdge) * w[0] + T33(tE, 2, 2, iEdge) * w[1] + T33(tE, 2, 3, iEdge) * w[2];
            } else {
                cc[0] = xC[iCell - 1] - xE[iEdge - 1];
                cc[1] = yC[iCell - 1] - yE[iEdge - 1];
            }
            if (!point_in_half_plane(ev[0], ev[1], cc)) err = IR_ERR_EDGE_ORIENTATION;
        }
    }

    int *EOER = g->edgesOnEdgeRemap, *COER = g->cellsOnEdgeRemap;
    memset(EOER, 0, sizeof(int) * ((size_t)nE + 1) * NEER);
    memset(COER, 0, sizeof(int) * ((size_t)nE + 1) * NCER);
    for (int iEdge = 1; iEdge <= nE; iEdge++) {
        if (remapEdge[iEdge - 1] != 1) continue;
        for (int k = 1; k <= 2; k++) A2(COER, k, iEdge, NCER) = A2(g->cellsOnEdge, k, iEdge, 2);
        for (int iCellOnEdge = 1; iCellOnEdge <= 2; iCellOnEdge++) {
            const int iCell = A2(g->cellsOnEdge, iCellOnEdge, iEdge, 2);
            if (iCell >= 1 && iCell <= nC) {
                const int n = g->nEdgesOnCell[iCell - 1];
                int iMain = 0, k;
                for (k = 1; k <= n; k++)
                    if (A2(g->edgesOnCell, k, iCell, M) == iEdge) { iMain = k; break; }
                if (iCellOnEdge == 1) {
                    k = iMain - 1; if (k < 1) k = k + n;
                    A2(EOER, 1, iEdge, NEER) = A2(g->edgesOnCell, k, iCell, M);
                    k = iMain + 1; if (k > n) k = k - n;
                    A2(EOER, 2, iEdge, NEER) = A2(g->edgesOnCell, k, iCell, M);
                } else {
                    k = iMain + 1; if (k > n) k = k - n;
                    A2(EOER, 3, iEdge, NEER) = A2(g->edgesOnCell, k, iCell, M);
                    k = iMain - 1; if (k < 1) k = k + n;
                    A2(EOER, 4, iEdge, NEER) = A2(g->edgesOnCell, k, iCell, M);
                }
            }
        }
        if (D == 4) {
            for (int iVertexOnEdge = 1; iVertexOnEdge <= 2; iVertexOnEdge++) {
                const int iVertex = A2(g->verticesOnEdge, iVertexOnEdge, iEdge, 2);
                for (int k = 1; k <= D; k++) {
                    const int iEdgeNeighbor = A2(g->edgesOnVertex, k, iVertex, D);
                    if (iEdgeNeighbor >= 1 && iEdgeNeighbor <= nE) {
                        int newRemapEdge = 1;
                        for (int q = 1; q <= 4; q++)
                            if (iEdgeNeighbor == A2(EOER, q, iEdge, NEER) || iEdgeNeighbor == iEdge) { newRemapEdge = 0; break; }
                        if (newRemapEdge) { A2(EOER, iVertexOnEdge + 4, iEdge, NEER) = iEdgeNeighbor; break; }
                    }
                }
            }
        }
        if (D == 3) {
            A2(COER, 3, iEdge, NCER) = nC + 1;
            A2(COER, 4, iEdge, NCER) = nC + 1;
            for (int iVertexOnEdge = 1; iVertexOnEdge <= 2; iVertexOnEdge++) {
                int iEdgeNeighbor = A2(EOER, iVertexOnEdge, iEdge, NEER);
                if (iEdgeNeighbor < 1 || iEdgeNeighbor > nE) iEdgeNeighbor = A2(EOER, iVertexOnEdge + 2, iEdge, NEER);
                if (iEdgeNeighbor < 1 || iEdgeNeighbor > nE) continue;   /* (the reference would index out of bounds) */
                for (int k = 1; k <= 2; k++) {
                    const int iCellNeighbor = A2(g->cellsOnEdge, k, iEdgeNeighbor, 2);
                    if (iCellNeighbor >= 1 && iCellNeighbor <= nC)
                        if (iCellNeighbor != A2(COER, 1, iEdge, NCER) && iCellNeighbor != A2(COER, 2, iEdge, NCER))
                            A2(COER, iVertexOnEdge + 2, iEdge, NCER) = iCellNeighbor;
                }
            }
        } else {
            for (int q = 3; q <= 6; q++) A2(COER, q, iEdge, NCER) = nC + 1;
            for (int iVertexOnEdge = 1; iVertexOnEdge <= 2; iVertexOnEdge++) {
                int iEdgeNeighbor = A2(EOER, iVertexOnEdge, iEdge, NEER);
                if (iEdgeNeighbor >= 1 && iEdgeNeighbor <= nE)
                    for (int k = 1; k <= 2; k++) {
                        const int iCellNeighbor = A2(g->cellsOnEdge, k, iEdgeNeighbor, 2);
                        if (iCellNeighbor >= 1 && iCellNeighbor <= nC && iCellNeighbor != A2(COER, 1, iEdge, NCER))
                            A2(COER, iVertexOnEdge + 2, iEdge, NCER) = iCellNeighbor;
                    }
                iEdgeNeighbor = A2(EOER, iVertexOnEdge + 2, iEdge, NEER);
                if (iEdgeNeighbor >= 1 && iEdgeNeighbor <= nE)
                    for (int k = 1; k <= 2; k++) {
                        const int iCellNeighbor = A2(g->cellsOnEdge, k, iEdgeNeighbor, 2);
                        if (iCellNeighbor >= 1 && iCellNeighbor <= nC && iCellNeighbor != A2(COER, 2, iEdge, NCER))
                            A2(COER, iVertexOnEdge + 4, iEdge, NCER) = iCellNeighbor;
                    }
            }
        }
    }

    const int nEdgesOnEdgeRemap = (D == 3) ? 4 : 6;
    double *XE = g->xVertexOnEdge, *YE = g->yVertexOnEdge;
    memset(XE, 0, sizeof(double) * ((size_t)nE + 1) * NVER);
    memset(YE, 0, sizeof(double) * ((size_t)nE + 1) * NVER);
    if (g->on_a_sphere) {
        for (int iEdge = 1; iEdge <= nE; iEdge++) {
            const int iVertex1 = A2(g->verticesOnEdge, 1, iEdge, 2), iVertex2 = A2(g->verticesOnEdge, 2, iEdge, 2);
            if (iVertex1 < 1 || iVertex1 > nV || iVertex2 < 1 || iVertex2 > nV) continue;
            const double w[3] = {xV[iVertex2 - 1] - xV[iVertex1 - 1], yV[iVertex2 - 1] - yV[iVertex1 - 1], zV[iVertex2 - 1] - zV[iVertex1 - 1]};
            const double xVector = T33(tE, 1, 1, iEdge) * w[0] + T33(tE, 1, 2, iEdge) * w[1] + T33(tE, 1, 3, iEdge) * w[2];
            const double yVector = T33(tE, 2, 1, iEdge) * w[0] + T33(tE, 2, 2, iEdge) * w[1] + T33(tE, 2, 3, iEdge) * w[2];
            A2(XE, 1, iEdge, NVER) = -0.5 * xVector;
            A2(YE, 1, iEdge, NVER) = -0.5 * yVector;
            A2(XE, 2, iEdge, NVER) = 0.5 * xVector;
            A2(YE, 2, iEdge, NVER) = 0.5 * yVector;
        }
        for (int iEdge = 1; iEdge <= nE; iEdge++) {
            if (remapEdge[iEdge - 1] != 1) continue;
            for (int iEdgeOnEdge = 1; iEdgeOnEdge <= nEdgesOnEdgeRemap; iEdgeOnEdge++) {
                const int iEdgeNeighbor = A2(EOER, iEdgeOnEdge, iEdge, NEER);
                if (iEdgeNeighbor >= 1 && iEdgeNeighbor <= nE) {
                    int n1 = 1, m1 = 1, m2 = 2;
                    for (int n = 1; n <= 2; n++)
                        for (int m = 1; m <= 2; m++)
                            if (A2(g->verticesOnEdge, m, iEdgeNeighbor, 2) == A2(g->verticesOnEdge, n, iEdge, 2)) {
                                n1 = n;
                                if (m == 1) { m1 = 1; m2 = 2; } else { m1 = 2; m2 = 1; }
                                break;
                            }
                    const int iVertexOnEdge = 2 + iEdgeOnEdge;
                    A2(XE, iVertexOnEdge, iEdge, NVER) =
                        A2(XE, n1, iEdge, NVER) + (A2(XE, m2, iEdgeNeighbor, NVER) - A2(XE, m1, iEdgeNeighbor, NVER));
                    A2(YE, iVertexOnEdge, iEdge, NVER) =
                        A2(YE, n1, iEdge, NVER) + (A2(YE, m2, iEdgeNeighbor, NVER) - A2(YE, m1, iEdgeNeighbor, NVER));
                }
            }
        }
    } else {
        for (int iEdge = 1; iEdge <= nE; iEdge++) {
            if (remapEdge[iEdge - 1] != 1) continue;
            for (int k = 1; k <= 2; k++) {
                const int iVertex = A2(g->verticesOnEdge, k, iEdge, 2);
                A2(XE, k, iEdge, NVER) = xV[iVertex - 1] - xE[iEdge - 1];
                A2(YE, k, iEdge, NVER) = yV[iVertex - 1] - yE[iEdge - 1];
            }
            /* the reference counts the neighbour edges that exist instead of using slot iEdgeOnEdge + 2 here
             * (:1740-1775); with every neighbour edge present the two are the same */
            int count = 2, iVertex = 0;
            for (int iEdgeOnEdge = 1; iEdgeOnEdge <= nEdgesOnEdgeRemap; iEdgeOnEdge++) {
                const int iEdgeNeighbor = A2(EOER, iEdgeOnEdge, iEdge, NEER);
                if (iEdgeNeighbor >= 1 && iEdgeNeighbor <= nE) {
                    for (int n = 1; n <= 2; n++)
                        for (int m = 1; m <= 2; m++)
                            if (A2(g->verticesOnEdge, m, iEdgeNeighbor, 2) == A2(g->verticesOnEdge, n, iEdge, 2)) {
                                const int m2 = (m == 1) ? 2 : 1;
                                iVertex = A2(g->verticesOnEdge, m2, iEdgeNeighbor, 2);
                                count = count + 1;
                                break;
                            }
                    if (count <= NVER && iVertex >= 1) {
                        A2(XE, count, iEdge, NVER) = xV[iVertex - 1] - xE[iEdge - 1];
                        A2(YE, count, iEdge, NVER) = yV[iVertex - 1] - yE[iEdge - 1];
                    }
                }
            }
        }
    }

    for (int iVertex = 1; iVertex <= nV; iVertex++) {
        double mn = DBL_MAX;
        for (int k = 1; k <= D; k++) {
            const int iEdge = A2(g->edgesOnVertex, k, iVertex, D);
            if (iEdge >= 1 && iEdge <= nE) {
                const int iVertex1 = A2(g->verticesOnEdge, 1, iEdge, 2), iVertex2 = A2(g->verticesOnEdge, 2, iEdge, 2);
                const double w[3] = {xV[iVertex2 - 1] - xV[iVertex1 - 1], yV[iVertex2 - 1] - yV[iVertex1 - 1], zV[iVertex2 - 1] - zV[iVertex1 - 1]};
                const double edgeLength = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
                if (edgeLength < mn) mn = edgeLength;
            }
        }
        g->minLengthEdgesOnVertex[iVertex - 1] = mn;
    }
    g->minLengthEdgesOnVertex[nV] = DBL_MAX;

    /* compute_geometric_cell_averages (:2097) */
    for (int k = 0; k < 14; k++) memset(g->geomAvg[k], 0, sizeof(double) * ((size_t)nC + 1));
    for (int iCell = 1; iCell <= nC; iCell++) {
        const int n = g->nEdgesOnCell[iCell - 1];
        double fracEdgeArea[16] = {0};
        double sumArea = 0.0;
        for (int k = 1; k <= n; k++) {
            const int iEdge = A2(g->edgesOnCell, k, iCell, M);
            fracEdgeArea[k - 1] = 0.25 * g->dcEdge[iEdge - 1] * g->dvEdge[iEdge - 1];
            sumArea = sumArea + fracEdgeArea[k - 1];
        }
        for (int k = 1; k <= n; k++) fracEdgeArea[k - 1] = fracEdgeArea[k - 1] / sumArea;
        double acc[14] = {0};
        for (int k = 1; k <= n; k++) {
            const double x1 = 0.0, y1 = 0.0;
            const double x2 = A2(g->xVertexOnCell, k, iCell, M), y2 = A2(g->yVertexOnCell, k, iCell, M);
            const int kp1 = (k + 1 > n) ? 1 : k + 1;
            const double x3 = A2(g->xVertexOnCell, kp1, iCell, M), y3 = A2(g->yVertexOnCell, kp1, iCell, M);
            double xq[6], yq[6], wq[6] = {W1QP, W1QP, W1QP, W2QP, W2QP, W2QP};
            xq[0] = Q1QP * x1 + Q1QP * x2 + Q2QP * x3; yq[0] = Q1QP * y1 + Q1QP * y2 + Q2QP * y3;
            xq[1] = Q1QP * x1 + Q2QP * x2 + Q1QP * x3; yq[1] = Q1QP * y1 + Q2QP * y2 + Q1QP * y3;
            xq[2] = Q2QP * x1 + Q1QP * x2 + Q1QP * x3; yq[2] = Q2QP * y1 + Q1QP * y2 + Q1QP * y3;
            xq[3] = Q3QP * x1 + Q4QP * x2 + Q4QP * x3; yq[3] = Q3QP * y1 + Q4QP * y2 + Q4QP * y3;
            xq[4] = Q4QP * x1 + Q3QP * x2 + Q4QP * x3; yq[4] = Q4QP * y1 + Q3QP * y2 + Q4QP * y3;
            xq[5] = Q4QP * x1 + Q4QP * x2 + Q3QP * x3; yq[5] = Q4QP * y1 + Q4QP * y2 + Q3QP * y3;
            double a[14] = {0};
            for (int q = 0; q < 6; q++) {
                /* integer powers as gfortran expands them: x**3 = x*x*x, x**4 = (x*x)*(x*x) */
                const double x = xq[q], y = yq[q], w = wq[q];
                const double x2p = x * x, y2p = y * y, x3p = x * x * x, y3p = y * y * y, x4p = (x * x) * (x * x), y4p = (y * y) * (y * y);
                a[0] = a[0] + w * x;
                a[1] = a[1] + w * y;
                a[2] = a[2] + w * x2p;
                a[3] = a[3] + w * x * y;
                a[4] = a[4] + w * y2p;
                a[5] = a[5] + w * x3p;
                a[6] = a[6] + w * x2p * y;
                a[7] = a[7] + w * x * y2p;
                a[8] = a[8] + w * y3p;
                a[9] = a[9] + w * x4p;
                a[10] = a[10] + w * x3p * y;
                a[11] = a[11] + w * x2p * y2p;
                a[12] = a[12] + w * x * y3p;
                a[13] = a[13] + w * y4p;
            }
            for (int q = 0; q < 14; q++) acc[q] = acc[q] + fracEdgeArea[k - 1] * a[q];
        }
        for (int q = 0; q < 14; q++) g->geomAvg[q][iCell - 1] = acc[q];
    }
    free(rot);
    free(tE);
    return err;
}

/* ------------------------------------------------------------------------------------------------------ run */

typedef struct {
    int nLayers;     /* 1 for the reference's "2D" tracers */
    int parent;      /* index into the tracer table (parents come first), -1 for the mass-like field */
    int nParents;    /* 0 (mass) .. 3 */
    int hasChild;
    int volumeLike;  /* 1: volume on entry / exit, thickness while transported (iceVolumeCategory, snowVolumeCategory; :2470, :2690) */
    double *array;   /* (nCells+1, nCategories, nLayers), IN/OUT */
} orc_ir_tracer;

typedef struct {
    int nCells, nCellsSolve, nVertices, nEdges, maxEdges, vertexDegree, nCategories, nQuadPoints;
    int on_a_sphere, rotate_cartesian_grid;
    double dt;
    const int *nEdgesOnCell, *edgesOnCell, *cellsOnCell, *verticesOnCell, *cellsOnEdge, *verticesOnEdge;
    const double *areaCell, *dcEdge;
    const double *coeffsReconstruct;   /* (nCells+1, maxEdges, 3) */
    const double *transGlobalToCell, *xVertexOnCell, *yVertexOnCell, *xVertexOnEdge, *yVertexOnEdge;
    const int *remapEdge, *cellsOnEdgeRemap, *edgesOnEdgeRemap;
    const double *geomAvg[14];
    const double *uVelocity, *vVelocity;
    int nTracers;
    orc_ir_tracer *tracers;
    /* optional diagnostics (may be NULL) */
    double *xTriangleOut, *yTriangleOut;  /* (nEdges, 6, nQuadPoints): quadrature points */
    double *triangleAreaOut;              /* (nEdges, 6) */
    int *iCellTriangleOut;                /* (nEdges, 6) */
    double *edgeFluxMassOut;              /* (nEdges, nCategories, nLayers of tracer 0) */
    int *maskEdgeOut;                     /* (nEdges) */
    double *xGradOut, *yGradOut;          /* tracer `gradTracerOut`: limited gradients (nCells+1, nCategories, nLayers) */
    int gradTracerOut;
    /* optional checks (config_conservation_check / config_monotonicity_check, both default .false.) */
    int conservationCheck;                /* 1: sums and the check on this block's sums; 2: sums only (the caller adds the ranks) */
    int monotonicityCheck;                /* 1: as the reference (in-place extension of the bounds); 2: order-independent extension */
    double *sumInitOut, *sumFinalOut;     /* tracer after tracer, (nCategories, nLayers) each: globalSumInit / globalSumFinal */
    int *consErrOut;                      /* [4]: violated, tracer (0-based), iCat, iLayer (1-based) of the first violation */
    int *monoErrOut;                      /* [5]: 0 none / 1 below the minimum / 2 above the maximum, tracer, iLayer, iCat, iCell */
    double *monoValOut;                   /* [3]: new value, the bound it crossed, the tolerance */
} orc_ir_run_args;

typedef struct {
    int *mask;
    double *center, *xGrad, *yGrad, *xBary, *yBary, *mtp, *edgeFlux;
} ir_work;

#define TIX(l, k, c, nL, nK) ((((size_t)(c) - 1) * (size_t)(nK) + ((size_t)(k) - 1)) * (size_t)(nL) + ((size_t)(l) - 1))

/* compute_barycenter_coordinates (:4658) */
static void barycenter(const double *const *G, int iCell, double *xB, double *yB, int nTerms,
                       const double *mean, const double *center, const double *xGrad, const double *yGrad)
{
    const size_t c = (size_t)iCell - 1;
    const double gx = G[0][c], gy = G[1][c], gxx = G[2][c], gxy = G[3][c], gyy = G[4][c], gxxx = G[5][c], gxxy = G[6][c],
                 gxyy = G[7][c], gyyy = G[8][c], gxxxx = G[9][c], gxxxy = G[10][c], gxxyy = G[11][c], gxyyy = G[12][c],
                 gyyyy = G[13][c];
    double reciprocal;
    if (nTerms == 0) {
        *xB = gx;
        *yB = gy;
    } else if (nTerms == 1) {
        const double c0 = center[0], cx = xGrad[0], cy = yGrad[0];
        reciprocal = (fabs(mean[0]) > 0.0) ? 1.0 / mean[0] : 0.0;
        *xB = (c0 * gx + cx * gxx + cy * gxy) * reciprocal;
        *yB = (c0 * gy + cx * gxy + cy * gyy) * reciprocal;
    } else if (nTerms == 2) {
        const double center0 = center[0], center1 = center[1], xGrad0 = xGrad[0], xGrad1 = xGrad[1], yGrad0 = yGrad[0], yGrad1 = yGrad[1];
        const double c0 = center0 * center1;
        const double cx = center0 * xGrad1 + xGrad0 * center1;
        const double cy = center0 * yGrad1 + yGrad0 * center1;
        const double cxx = xGrad0 * xGrad1;
        const double cxy = xGrad0 * yGrad1 + yGrad0 * xGrad1;
        const double cyy = yGrad0 * yGrad1;
        const double massTracerProd = mean[0] * mean[1];
        reciprocal = (fabs(massTracerProd) > 0.0) ? 1.0 / massTracerProd : 0.0;
        *xB = (c0 * gx + cx * gxx + cy * gxy + cxx * gxxx + cxy * gxxy + cyy * gxyy) * reciprocal;
        *yB = (c0 * gy + cx * gxy + cy * gyy + cxx * gxxy + cxy * gxyy + cyy * gyyy) * reciprocal;
    } else {
        const double center0 = center[0], center1 = center[1], center2 = center[2];
        const double xGrad0 = xGrad[0], xGrad1 = xGrad[1], xGrad2 = xGrad[2], yGrad0 = yGrad[0], yGrad1 = yGrad[1], yGrad2 = yGrad[2];
        const double c0 = center0 * center1 * center2;
        const double cx = center0 * center1 * xGrad2 + center0 * xGrad1 * center2 + xGrad0 * center1 * center2;
        const double cy = center0 * center1 * yGrad2 + center0 * yGrad1 * center2 + yGrad0 * center1 * center2;
        const double cxx = center0 * xGrad1 * xGrad2 + xGrad0 * center1 * xGrad2 + xGrad0 * xGrad1 * center2;
        const double cxy = center0 * xGrad1 * yGrad2 + xGrad0 * yGrad1 * center2 + yGrad0 * center1 * xGrad2 +
                           center0 * yGrad1 * xGrad2 + xGrad0 * center1 * yGrad2 + yGrad0 * xGrad1 * center2;
        const double cyy = center0 * yGrad1 * yGrad2 + yGrad0 * center1 * yGrad2 + yGrad0 * yGrad1 * center2;
        const double cxxx = xGrad0 * xGrad1 * xGrad2;
        const double cxxy = xGrad0 * xGrad1 * yGrad2 + xGrad0 * yGrad1 * xGrad2 + yGrad0 * xGrad1 * xGrad2;
        const double cxyy = yGrad0 * yGrad1 * xGrad2 + yGrad0 * xGrad1 * yGrad2 + xGrad0 * yGrad1 * yGrad2;
        const double cyyy = yGrad0 * yGrad1 * yGrad2;
        const double massTracerProd = mean[0] * mean[1] * mean[2];
        reciprocal = (fabs(massTracerProd) > 0.0) ? 1.0 / massTracerProd : 0.0;
        *xB = (c0 * gx + cx * gxx + cy * gxy + cxx * gxxx + cxy * gxxy + cyy * gxyy + cxxx * gxxxx + cxxy * gxxxy + cxyy * gxxyy +
               cyyy * gxyyy) * reciprocal;
        *yB = (c0 * gy + cx * gxy + cy * gyy + cxx * gxxy + cxy * gxyy + cyy * gyyy + cxxx * gxxxy + cxxy * gxxyy + cxyy * gxyyy +
               cyyy * gyyyy) * reciprocal;
    }
}

/* shift_vertices_of_departure_triangle (:6270) */
static int shift_vertices(const orc_ir_run_args *a, int iEdge, int iCell, double *ev1, double *ev2, double *xT, double *yT,
                          int tvOnEdge, int tvOnCell, double *area)
{
    const int M = a->maxEdges;
    int err = IR_OK;
    if (a->on_a_sphere) {
        const double crossProduct = cross_product_2d(ev1, ev2);
        if (fabs(crossProduct) < EPS11) err = IR_ERR_PARALLEL_EDGES;
        if (crossProduct > 0.0) {
            const double t0 = ev1[0], t1 = ev1[1];
            ev1[0] = ev2[0]; ev1[1] = ev2[1];
            ev2[0] = t0; ev2[1] = t1;
        }
        const int n = a->nEdgesOnCell[iCell - 1];
        const int k = tvOnCell;
        int km1 = k - 1; if (km1 < 1) km1 = km1 + n;
        int kp1 = k + 1; if (kp1 > n) kp1 = kp1 - n;
        const double e1c[2] = {A2(a->xVertexOnCell, km1, iCell, M) - A2(a->xVertexOnCell, k, iCell, M),
                               A2(a->yVertexOnCell, km1, iCell, M) - A2(a->yVertexOnCell, k, iCell, M)};
        const double e2c[2] = {A2(a->xVertexOnCell, kp1, iCell, M) - A2(a->xVertexOnCell, k, iCell, M),
                               A2(a->yVertexOnCell, kp1, iCell, M) - A2(a->yVertexOnCell, k, iCell, M)};
        const double denom = ev1[0] * ev2[1] - ev2[0] * ev1[1];
        for (int t = 0; t < 3; t++) {
            xT[t] = xT[t] - A2(a->xVertexOnEdge, tvOnEdge, iEdge, NVER);
            yT[t] = yT[t] - A2(a->yVertexOnEdge, tvOnEdge, iEdge, NVER);
            const double coeff_a = (xT[t] * ev2[1] - yT[t] * ev2[0]) / denom;
            const double coeff_b = (yT[t] * ev1[0] - xT[t] * ev1[1]) / denom;
            xT[t] = A2(a->xVertexOnCell, k, iCell, M) + coeff_a * e1c[0] + coeff_b * e2c[0];
            yT[t] = A2(a->yVertexOnCell, k, iCell, M) + coeff_a * e1c[1] + coeff_b * e2c[1];
        }
    } else {
        for (int t = 0; t < 3; t++) {
            xT[t] = xT[t] - A2(a->xVertexOnEdge, tvOnEdge, iEdge, NVER) + A2(a->xVertexOnCell, tvOnCell, iCell, M);
            yT[t] = yT[t] - A2(a->yVertexOnEdge, tvOnEdge, iEdge, NVER) + A2(a->yVertexOnCell, tvOnCell, iCell, M);
        }
    }
    *area = fabs(0.5 * ((xT[1] - xT[0]) * (yT[2] - yT[0]) - (yT[1] - yT[0]) * (xT[2] - xT[0])));
    return err;
}

static int vertex_on_cell(const orc_ir_run_args *a, int iCell, int iVertexGlobal, int prev)
{
    /* the reference's search loops have no exit: the last match wins; no match keeps the previous value */
    int r = prev;
    if (iCell < 1 || iCell > a->nCells) return r;
    for (int k = 1; k <= a->nEdgesOnCell[iCell - 1]; k++)
        if (iVertexGlobal == A2(a->verticesOnCell, k, iCell, a->maxEdges)) r = k;
    return r;
}

/* find_departure_triangles (:5365) for one edge.  xT / yT: [NTRI][nQP] (vertices in the first three entries). */
static int departure_triangles_edge(const orc_ir_run_args *a, int iEdge, const double *dpIn /* (nVertices+1, 2) */,
                                    double *xT, double *yT, int nQP, int *iCellTri, double *triArea)
{
    const int D = a->vertexDegree, nE = a->nEdges, nC = a->nCells;
    const double *XE = a->xVertexOnEdge, *YE = a->yVertexOnEdge;
    const int *EOER = a->edgesOnEdgeRemap, *COER = a->cellsOnEdgeRemap;
    int err = IR_OK;
    int tvOnEdge[NTRI + 2] = {0}, tvOnCell[NTRI + 2] = {0}, fluxSign[NTRI + 2] = {0};
    double ev1[NTRI + 2][2], ev2[NTRI + 2][2];
    memset(ev1, 0, sizeof ev1);
    memset(ev2, 0, sizeof ev2);
    double dp[2][2], edgeVertex[2][2], enb[2][2], ip[2] = {0, 0}, ipMain[2] = {0, 0};
    int dpInHalfPlane[2];
#define XT(q, t) xT[((t) - 1) * nQP + ((q) - 1)]
#define YT(q, t) yT[((t) - 1) * nQP + ((q) - 1)]
#define NEWTRI() do { triangleCount = triangleCount + 1; if (triangleCount > NTRI) return IR_ERR_TOO_MANY_TRIANGLES; } while (0)
    for (int k = 1; k <= 2; k++) {
        const int iVertex = A2(a->verticesOnEdge, k, iEdge, 2);
        dp[k - 1][0] = A2(XE, k, iEdge, NVER) + dpIn[((size_t)iVertex - 1) * 2 + 0];
        dp[k - 1][1] = A2(YE, k, iEdge, NVER) + dpIn[((size_t)iVertex - 1) * 2 + 1];
    }
    edgeVertex[0][0] = A2(XE, 1, iEdge, NVER); edgeVertex[0][1] = A2(YE, 1, iEdge, NVER);
    edgeVertex[1][0] = A2(XE, 2, iEdge, NVER); edgeVertex[1][1] = A2(YE, 2, iEdge, NVER);
    for (int k = 0; k < 2; k++) dpInHalfPlane[k] = point_in_half_plane(edgeVertex[0], edgeVertex[1], dp[k]);
    int triangleCount = 0;

    for (int iVertexOnEdge = 1; iVertexOnEdge <= 2; iVertexOnEdge++) {
        enb[0][0] = edgeVertex[iVertexOnEdge - 1][0];
        enb[0][1] = edgeVertex[iVertexOnEdge - 1][1];
        for (int iSideIndex = 0; iSideIndex <= 1; iSideIndex++) {
            int iEdgeOnEdgeRemap = iVertexOnEdge + 2 * iSideIndex;
            const int iVertexOnEdgeRemap = iEdgeOnEdgeRemap + 2;
            int iEdgeNeighbor = A2(EOER, iEdgeOnEdgeRemap, iEdge, NEER);
            int edgeIntersect;
            if (iEdgeNeighbor >= 1 && iEdgeNeighbor <= nE) {
                enb[1][0] = A2(XE, iVertexOnEdgeRemap, iEdge, NVER);
                enb[1][1] = A2(YE, iVertexOnEdgeRemap, iEdge, NVER);
                edgeIntersect = find_line_intersection(dp[0], dp[1], enb[0], enb[1], ip);
            } else {
                edgeIntersect = 0;
            }
            if (!edgeIntersect) continue;
            NEWTRI();
            XT(1, triangleCount) = edgeVertex[iVertexOnEdge - 1][0]; YT(1, triangleCount) = edgeVertex[iVertexOnEdge - 1][1];
            XT(2, triangleCount) = dp[iVertexOnEdge - 1][0];         YT(2, triangleCount) = dp[iVertexOnEdge - 1][1];
            XT(3, triangleCount) = ip[0];                            YT(3, triangleCount) = ip[1];
            tvOnEdge[triangleCount] = iVertexOnEdge;
            iCellTri[triangleCount - 1] = A2(COER, iVertexOnEdge + 2, iEdge, NCER);
            const int vGlobal = A2(a->verticesOnEdge, iVertexOnEdge, iEdge, 2);
            tvOnCell[triangleCount] = vertex_on_cell(a, iCellTri[triangleCount - 1], vGlobal, tvOnCell[triangleCount]);
            if (a->on_a_sphere) {
                ev1[triangleCount][0] = enb[1][0] - enb[0][0];
                ev1[triangleCount][1] = enb[1][1] - enb[0][1];
                int iOtherEdge;
                if (D == 3) { iOtherEdge = iEdgeOnEdgeRemap + 2; if (iOtherEdge > 4) iOtherEdge = iOtherEdge - 4; }
                else iOtherEdge = iVertexOnEdge + 4;
                const int iOtherVertex = iOtherEdge + 2;
                enb[1][0] = A2(XE, iOtherVertex, iEdge, NVER);
                enb[1][1] = A2(YE, iOtherVertex, iEdge, NVER);
                ev2[triangleCount][0] = enb[1][0] - enb[0][0];
                ev2[triangleCount][1] = enb[1][1] - enb[0][1];
            }
            fluxSign[triangleCount] = (iSideIndex == 0) ? 1 : -1;
            if (D == 4) {
                int edgeIntersectMain;
                iEdgeOnEdgeRemap = iVertexOnEdge + 4;
                iEdgeNeighbor = A2(EOER, iEdgeOnEdgeRemap, iEdge, NEER);
                if (iEdgeNeighbor >= 1 && iEdgeNeighbor <= nE) {
                    enb[1][0] = A2(XE, iVertexOnEdge + 6, iEdge, NVER);
                    enb[1][1] = A2(YE, iVertexOnEdge + 6, iEdge, NVER);
                    edgeIntersectMain = find_line_intersection(dp[0], dp[1], enb[0], enb[1], ipMain);
                } else {
                    edgeIntersectMain = 0;
                }
                if (edgeIntersectMain) {
                    XT(3, triangleCount) = ipMain[0];
                    YT(3, triangleCount) = ipMain[1];
                    if (iSideIndex == 0) {
                        iCellTri[triangleCount - 1] = A2(COER, iVertexOnEdge + 4, iEdge, NCER);
                        tvOnCell[triangleCount] = vertex_on_cell(a, iCellTri[triangleCount - 1], vGlobal, tvOnCell[triangleCount]);
                        if (a->on_a_sphere) {
                            const int iOtherVertex = iVertexOnEdgeRemap + 2;
                            enb[1][0] = A2(XE, iOtherVertex, iEdge, NVER);
                            enb[1][1] = A2(YE, iOtherVertex, iEdge, NVER);
                            ev1[triangleCount][0] = enb[1][0] - enb[0][0];
                            ev1[triangleCount][1] = enb[1][1] - enb[0][1];
                        }
                    } else if (a->on_a_sphere) {
                        const int iOtherVertex = iVertexOnEdgeRemap - 2;
                        enb[1][0] = A2(XE, iOtherVertex, iEdge, NVER);
                        enb[1][1] = A2(YE, iOtherVertex, iEdge, NVER);
                        ev1[triangleCount][0] = enb[1][0] - enb[0][0];
                        ev1[triangleCount][1] = enb[1][1] - enb[0][1];
                    }
                    NEWTRI();
                    XT(1, triangleCount) = edgeVertex[iVertexOnEdge - 1][0]; YT(1, triangleCount) = edgeVertex[iVertexOnEdge - 1][1];
                    XT(2, triangleCount) = ipMain[0];                        YT(2, triangleCount) = ipMain[1];
                    XT(3, triangleCount) = ip[0];                            YT(3, triangleCount) = ip[1];
                    iCellTri[triangleCount - 1] = (iSideIndex == 0) ? A2(COER, iVertexOnEdge + 2, iEdge, NCER)
                                                                    : A2(COER, iVertexOnEdge + 4, iEdge, NCER);
                    tvOnEdge[triangleCount] = tvOnEdge[triangleCount - 1];
                    tvOnCell[triangleCount] = vertex_on_cell(a, iCellTri[triangleCount - 1], vGlobal, tvOnCell[triangleCount]);
                    if (a->on_a_sphere) {
                        enb[1][0] = A2(XE, iVertexOnEdgeRemap, iEdge, NVER);
                        enb[1][1] = A2(YE, iVertexOnEdgeRemap, iEdge, NVER);
                        ev1[triangleCount][0] = enb[1][0] - enb[0][0];
                        ev1[triangleCount][1] = enb[1][1] - enb[0][1];
                        ev2[triangleCount][0] = ev2[triangleCount - 1][0];
                        ev2[triangleCount][1] = ev2[triangleCount - 1][1];
                    }
                    fluxSign[triangleCount] = fluxSign[triangleCount - 1];
                } else if (iSideIndex != 0) {
                    iCellTri[triangleCount - 1] = A2(COER, iVertexOnEdge + 4, iEdge, NCER);
                    tvOnCell[triangleCount] = vertex_on_cell(a, iCellTri[triangleCount - 1], vGlobal, tvOnCell[triangleCount]);
                }
            }
            dp[iVertexOnEdge - 1][0] = ip[0];
            dp[iVertexOnEdge - 1][1] = ip[1];
        }
    }

    const int edgeIntersectMain = find_line_intersection(dp[0], dp[1], edgeVertex[0], edgeVertex[1], ipMain);
    if (edgeIntersectMain) {
        for (int iVertexOnEdge = 1; iVertexOnEdge <= 2; iVertexOnEdge++) {
            NEWTRI();
            XT(1, triangleCount) = edgeVertex[iVertexOnEdge - 1][0]; YT(1, triangleCount) = edgeVertex[iVertexOnEdge - 1][1];
            XT(2, triangleCount) = dp[iVertexOnEdge - 1][0];         YT(2, triangleCount) = dp[iVertexOnEdge - 1][1];
            XT(3, triangleCount) = ipMain[0];                        YT(3, triangleCount) = ipMain[1];
            tvOnEdge[triangleCount] = iVertexOnEdge;
            dpInHalfPlane[iVertexOnEdge - 1] = point_in_half_plane(edgeVertex[0], edgeVertex[1], dp[iVertexOnEdge - 1]);
            if (dpInHalfPlane[iVertexOnEdge - 1]) { iCellTri[triangleCount - 1] = A2(COER, 1, iEdge, NCER); fluxSign[triangleCount] = 1; }
            else { iCellTri[triangleCount - 1] = A2(COER, 2, iEdge, NCER); fluxSign[triangleCount] = -1; }
            tvOnCell[triangleCount] = vertex_on_cell(a, iCellTri[triangleCount - 1], A2(a->verticesOnEdge, iVertexOnEdge, iEdge, 2),
                                                     tvOnCell[triangleCount]);
            if (a->on_a_sphere) {
                const int iOtherVertex = (iVertexOnEdge == 1) ? 2 : 1;
                ev1[triangleCount][0] = edgeVertex[iOtherVertex - 1][0] - edgeVertex[iVertexOnEdge - 1][0];
                ev1[triangleCount][1] = edgeVertex[iOtherVertex - 1][1] - edgeVertex[iVertexOnEdge - 1][1];
                const int iOtherEdge = dpInHalfPlane[iVertexOnEdge - 1] ? iVertexOnEdge : iVertexOnEdge + 2;
                ev2[triangleCount][0] = A2(XE, iOtherEdge + 2, iEdge, NVER) - edgeVertex[iVertexOnEdge - 1][0];
                ev2[triangleCount][1] = A2(YE, iOtherEdge + 2, iEdge, NVER) - edgeVertex[iVertexOnEdge - 1][1];
            }
        }
    } else {
        const double quadArea = quadrilateral_area(edgeVertex[0], edgeVertex[1], dp[1], dp[0]);
        if (quadArea > 0.0) {
            for (int iVertexOnEdge = 1; iVertexOnEdge <= 2; iVertexOnEdge++) {
                NEWTRI();
                if (iVertexOnEdge == 1) {
                    XT(1, triangleCount) = edgeVertex[0][0]; YT(1, triangleCount) = edgeVertex[0][1];
                    XT(2, triangleCount) = edgeVertex[1][0]; YT(2, triangleCount) = edgeVertex[1][1];
                    XT(3, triangleCount) = dp[0][0];         YT(3, triangleCount) = dp[0][1];
                } else {
                    XT(1, triangleCount) = edgeVertex[1][0]; YT(1, triangleCount) = edgeVertex[1][1];
                    XT(2, triangleCount) = dp[0][0];         YT(2, triangleCount) = dp[0][1];
                    XT(3, triangleCount) = dp[1][0];         YT(3, triangleCount) = dp[1][1];
                }
                tvOnEdge[triangleCount] = iVertexOnEdge;
                dpInHalfPlane[iVertexOnEdge - 1] = point_in_half_plane(edgeVertex[0], edgeVertex[1], dp[iVertexOnEdge - 1]);
                if (dpInHalfPlane[iVertexOnEdge - 1]) { iCellTri[triangleCount - 1] = A2(COER, 1, iEdge, NCER); fluxSign[triangleCount] = 1; }
                else { iCellTri[triangleCount - 1] = A2(COER, 2, iEdge, NCER); fluxSign[triangleCount] = -1; }
                tvOnCell[triangleCount] = vertex_on_cell(a, iCellTri[triangleCount - 1], A2(a->verticesOnEdge, iVertexOnEdge, iEdge, 2),
                                                         tvOnCell[triangleCount]);
                if (a->on_a_sphere) {
                    const int iOtherVertex = (iVertexOnEdge == 1) ? 2 : 1;
                    ev1[triangleCount][0] = edgeVertex[iOtherVertex - 1][0] - edgeVertex[iVertexOnEdge - 1][0];
                    ev1[triangleCount][1] = edgeVertex[iOtherVertex - 1][1] - edgeVertex[iVertexOnEdge - 1][1];
                    const int iOtherEdge = dpInHalfPlane[iVertexOnEdge - 1] ? iVertexOnEdge : iVertexOnEdge + 2;
                    ev2[triangleCount][0] = A2(XE, iOtherEdge + 2, iEdge, NVER) - edgeVertex[iVertexOnEdge - 1][0];
                    ev2[triangleCount][1] = A2(YE, iOtherEdge + 2, iEdge, NVER) - edgeVertex[iVertexOnEdge - 1][1];
                }
            }
        }
    }

    for (int iTri = 1; iTri <= triangleCount; iTri++) {
        const int iCell = iCellTri[iTri - 1];
        if (iCell >= 1 && iCell <= nC) {
            double x3[3] = {XT(1, iTri), XT(2, iTri), XT(3, iTri)}, y3[3] = {YT(1, iTri), YT(2, iTri), YT(3, iTri)};
            double area;
            const int e = shift_vertices(a, iEdge, iCell, ev1[iTri], ev2[iTri], x3, y3, tvOnEdge[iTri], tvOnCell[iTri], &area);
            if (e) err = e;
            for (int t = 0; t < 3; t++) { XT(t + 1, iTri) = x3[t]; YT(t + 1, iTri) = y3[t]; }
            triArea[iTri - 1] = area * fluxSign[iTri];
        }
    }
#undef NEWTRI
    return err;
}

/* get_triangle_quadrature_points (:6546) for one edge; the nQuadPoints == 3 branch keeps the reference's
 * yMidpoint, which is computed from xTriangle (:6598) */
static void quadrature_points_edge(double *xT, double *yT, int nQP)
{
    for (int iTri = 1; iTri <= NTRI; iTri++) {
        if (nQP == 3) {
            const double xMid = (XT(1, iTri) + XT(2, iTri) + XT(3, iTri)) / 3.0;
            const double yMid = (XT(1, iTri) + XT(2, iTri) + XT(3, iTri)) / 3.0;
            for (int q = 1; q <= 3; q++) {
                XT(q, iTri) = 0.5 * (XT(q, iTri) + xMid);
                YT(q, iTri) = 0.5 * (YT(q, iTri) + yMid);
            }
        } else {
            const double x1 = XT(1, iTri), y1 = YT(1, iTri), x2 = XT(2, iTri), y2 = YT(2, iTri), x3 = XT(3, iTri), y3 = YT(3, iTri);
            XT(1, iTri) = Q1QP * x1 + Q1QP * x2 + Q2QP * x3; YT(1, iTri) = Q1QP * y1 + Q1QP * y2 + Q2QP * y3;
            XT(2, iTri) = Q1QP * x1 + Q2QP * x2 + Q1QP * x3; YT(2, iTri) = Q1QP * y1 + Q2QP * y2 + Q1QP * y3;
            XT(3, iTri) = Q2QP * x1 + Q1QP * x2 + Q1QP * x3; YT(3, iTri) = Q2QP * y1 + Q1QP * y2 + Q1QP * y3;
            XT(4, iTri) = Q3QP * x1 + Q4QP * x2 + Q4QP * x3; YT(4, iTri) = Q3QP * y1 + Q4QP * y2 + Q4QP * y3;
            XT(5, iTri) = Q4QP * x1 + Q3QP * x2 + Q4QP * x3; YT(5, iTri) = Q4QP * y1 + Q3QP * y2 + Q4QP * y3;
            XT(6, iTri) = Q4QP * x1 + Q4QP * x2 + Q3QP * x3; YT(6, iTri) = Q4QP * y1 + Q4QP * y2 + Q3QP * y3;
        }
    }
}
#undef XT
#undef YT

int orc_ir_run(orc_ir_run_args *a)
{
    const int nC = a->nCells, nV = a->nVertices, nE = a->nEdges, M = a->maxEdges, nK = a->nCategories, nQP = a->nQuadPoints;
    const int nT = a->nTracers;
    if (nQP != 3 && nQP != 6) return IR_ERR_BAD_ARGUMENT;
    if (nT < 1 || a->tracers[0].nParents != 0 || a->tracers[0].parent != -1) return IR_ERR_BAD_ARGUMENT;
    for (int t = 1; t < nT; t++) {
        const orc_ir_tracer *tr = &a->tracers[t];
        if (tr->parent < 0 || tr->parent >= t) return IR_ERR_BAD_ARGUMENT;
        const orc_ir_tracer *p = &a->tracers[tr->parent];
        if (tr->nParents != p->nParents + 1) return IR_ERR_BAD_ARGUMENT;
        if (tr->nParents > 3) return IR_ERR_TOO_MANY_PARENTS;
        if (p->nLayers != 1 && p->nLayers != tr->nLayers) return IR_ERR_BAD_ARGUMENT;
    }
    int err = IR_OK;
    double weightQuadPoint[6];
    if (nQP == 3) for (int q = 0; q < 3; q++) weightQuadPoint[q] = 1.0 / 3.0;
    else { for (int q = 0; q < 3; q++) weightQuadPoint[q] = W1QP; for (int q = 3; q < 6; q++) weightQuadPoint[q] = W2QP; }
    orc_ir_tracer *mass = &a->tracers[0];

    /* volume -> thickness (:2462-2480); loops over every column including the junk one */
    for (int t = 0; t < nT; t++) {
        orc_ir_tracer *tr = &a->tracers[t];
        if (!tr->volumeLike) continue;
        for (int c = 1; c <= nC + 1; c++)
            for (int k = 1; k <= nK; k++) {
                const double area = mass->array[TIX(1, k, c, mass->nLayers, nK)];
                double *v = &tr->array[TIX(1, k, c, tr->nLayers, nK)];
                if (area > 0.0) *v = *v / area; else *v = 0.0;
            }
    }

    ir_work *W = (ir_work *)calloc((size_t)nT, sizeof(ir_work));
    for (int t = 0; t < nT; t++) {
        const size_t n = ((size_t)nC + 1) * nK * a->tracers[t].nLayers;
        W[t].mask = (int *)calloc(n, sizeof(int));
        W[t].center = (double *)calloc(n, sizeof(double));
        W[t].xGrad = (double *)calloc(n, sizeof(double));
        W[t].yGrad = (double *)calloc(n, sizeof(double));
        W[t].xBary = (double *)calloc(n, sizeof(double));
        W[t].yBary = (double *)calloc(n, sizeof(double));
        W[t].mtp = (double *)calloc(n, sizeof(double));
        W[t].edgeFlux = (double *)calloc(((size_t)nE + 1) * nK * a->tracers[t].nLayers, sizeof(double));
    }
    int *maskCell = (int *)calloc((size_t)nC + 1, sizeof(int));
    int *maskEdge = (int *)calloc((size_t)nE + 1, sizeof(int));
    double *dpIn = (double *)calloc(((size_t)nV + 1) * 2, sizeof(double));
    double *xTri = (double *)calloc((size_t)(nE > 0 ? nE : 1) * NTRI * nQP, sizeof(double));
    double *yTri = (double *)calloc((size_t)(nE > 0 ? nE : 1) * NTRI * nQP, sizeof(double));
    double *triArea = (double *)calloc((size_t)(nE > 0 ? nE : 1) * NTRI, sizeof(double));
    int *iCellTri = (int *)calloc((size_t)(nE > 0 ? nE : 1) * NTRI, sizeof(int));

    /* config_monotonicity_check (:2999-3015): make_masks with threshold 0, then tracer_local_min_max (:8268) on the
     * old values.  The reference fills localMin / localMax for the owned cells and brings the halo cells in with a
     * halo exchange (:8450); this single-block restatement evaluates the same loop for every cell of the block
     * instead -- for the halo cells next to owned ones (all the extension below reads) their neighbours are present
     * with the two halo layers the scheme requires (:829), so the values are those the owner would send. */
    double **lmin = NULL, **lmax = NULL;
    if (a->monotonicityCheck) {
        lmin = (double **)calloc((size_t)nT, sizeof(double *));
        lmax = (double **)calloc((size_t)nT, sizeof(double *));
        for (int t = 0; t < nT; t++) {
            const orc_ir_tracer *tr = &a->tracers[t];
            const int nL = tr->nLayers;
            const size_t n = ((size_t)nC + 1) * nK * nL;
            lmin[t] = (double *)calloc(n, sizeof(double));
            lmax[t] = (double *)calloc(n, sizeof(double));
            int *mask0 = (int *)calloc(n, sizeof(int));
            if (tr->nParents == 0) {
                for (int c = 1; c <= nC; c++)
                    for (int k = 1; k <= nK; k++)
                        for (int l = 1; l <= nL; l++) mask0[TIX(l, k, c, nL, nK)] = 1;
            } else {
                const orc_ir_tracer *p = &a->tracers[tr->parent];
                for (int c = 1; c <= nC; c++)
                    for (int k = 1; k <= nK; k++)
                        for (int l = 1; l <= nL; l++)
                            if (p->array[TIX((p->nLayers == 1) ? 1 : l, k, c, p->nLayers, nK)] > 0.0) mask0[TIX(l, k, c, nL, nK)] = 1;
            }
            for (int iCell = 1; iCell <= nC; iCell++)
                for (int k = 1; k <= nK; k++)
                    for (int l = 1; l <= nL; l++) {
                        const size_t q = TIX(l, k, iCell, nL, nK);
                        if (mask0[q] == 1) { lmin[t][q] = tr->array[q]; lmax[t][q] = tr->array[q]; }
                        for (int iCellOnCell = 1; iCellOnCell <= a->nEdgesOnCell[iCell - 1]; iCellOnCell++) {
                            const int iCellNeighbor = A2(a->cellsOnCell, iCellOnCell, iCell, M);
                            if (iCellNeighbor >= 1 && iCellNeighbor <= nC) {
                                const size_t qn = TIX(l, k, iCellNeighbor, nL, nK);
                                if (mask0[qn] == 1) {
                                    if (tr->array[qn] < lmin[t][q]) lmin[t][q] = tr->array[qn];
                                    if (tr->array[qn] > lmax[t][q]) lmax[t][q] = tr->array[qn];
                                }
                            }
                        }
                    }
            free(mask0);
        }
    }

    /* make_masks, threshold eps11 (:3404; called at :2995) */
    for (int t = 0; t < nT; t++) {
        const orc_ir_tracer *tr = &a->tracers[t];
        const int nL = tr->nLayers;
        if (tr->nParents == 0) {
            for (int c = 1; c <= nC; c++) {
                double massSumCell = 0.0;
                for (int k = 1; k <= nK; k++)
                    for (int l = 1; l <= nL; l++) {
                        W[t].mask[TIX(l, k, c, nL, nK)] = 1;
                        massSumCell = massSumCell + tr->array[TIX(l, k, c, nL, nK)];
                    }
                maskCell[c - 1] = (massSumCell > 0.0) ? 1 : 0;
            }
        } else {
            const orc_ir_tracer *p = &a->tracers[tr->parent];
            for (int c = 1; c <= nC; c++)
                for (int k = 1; k <= nK; k++)
                    for (int l = 1; l <= nL; l++) {
                        const int lp = (p->nLayers == 1) ? 1 : l;
                        if (p->array[TIX(lp, k, c, p->nLayers, nK)] > EPS11) W[t].mask[TIX(l, k, c, nL, nK)] = 1;
                    }
        }
    }

    /* construct_linear_tracer_fields (:3580) */
    for (int t = 0; t < nT; t++) {
        const orc_ir_tracer *tr = &a->tracers[t];
        const int nL = tr->nLayers;
        const int ip = tr->parent;
        const orc_ir_tracer *p = (ip >= 0) ? &a->tracers[ip] : NULL;
        const int pL = p ? p->nLayers : 1;
        const double *field = tr->array;
        const int *mask = W[t].mask;
        double *xGrad = W[t].xGrad, *yGrad = W[t].yGrad;
#pragma omp parallel for schedule(static)
        for (int iCell = 1; iCell <= nC; iCell++) {
            if (maskCell[iCell - 1] != 1) continue;
            const int n = a->nEdgesOnCell[iCell - 1];
            for (int k = 1; k <= nK; k++)
                for (int l = 1; l <= nL; l++) {
                    /* compute_gradient_2d / _3d (:4204, :4429) */
                    double g1 = 0.0, g2 = 0.0, g3 = 0.0;
                    for (int iEdgeOnCell = 1; iEdgeOnCell <= n; iEdgeOnCell++) {
                        double normalGrad = 0.0;
                        const int iCellNeighbor = A2(a->cellsOnCell, iEdgeOnCell, iCell, M);
                        if (iCellNeighbor >= 1 && iCellNeighbor <= nC) {
                            const int iEdge = A2(a->edgesOnCell, iEdgeOnCell, iCell, M);
                            if (mask[TIX(l, k, iCell, nL, nK)] == 1 && mask[TIX(l, k, iCellNeighbor, nL, nK)] == 1) {
                                const double signGradient = (iCell == A2(a->cellsOnEdge, 1, iEdge, 2)) ? 1.0 : -1.0;
                                normalGrad = signGradient * (field[TIX(l, k, iCellNeighbor, nL, nK)] - field[TIX(l, k, iCell, nL, nK)]) /
                                             a->dcEdge[iEdge - 1];
                            }
                        }
                        const double *cr = &a->coeffsReconstruct[(((size_t)iCell - 1) * M + (iEdgeOnCell - 1)) * 3];
                        g1 = g1 + cr[0] * normalGrad;
                        g2 = g2 + cr[1] * normalGrad;
                        g3 = g3 + cr[2] * normalGrad;
                    }
                    if (a->rotate_cartesian_grid && a->on_a_sphere) {
                        const double tempGrad = g1;
                        g1 = -g3;
                        g3 = tempGrad;
                    }
                    double xg, yg;
                    if (a->on_a_sphere) {
                        const double *tt = a->transGlobalToCell;
                        xg = T33(tt, 1, 1, iCell) * g1 + T33(tt, 1, 2, iCell) * g2 + T33(tt, 1, 3, iCell) * g3;
                        yg = T33(tt, 2, 1, iCell) * g1 + T33(tt, 2, 2, iCell) * g2 + T33(tt, 2, 3, iCell) * g3;
                    } else {
                        xg = g1;
                        yg = g2;
                    }
                    /* limit_tracer_gradient_2d / _3d (:4802, :4999) */
                    const int lp = (pL == 1) ? 1 : l;
                    const double xB = p ? W[ip].xBary[TIX(lp, k, iCell, pL, nK)] : a->geomAvg[0][iCell - 1];
                    const double yB = p ? W[ip].yBary[TIX(lp, k, iCell, pL, nK)] : a->geomAvg[1][iCell - 1];
                    const double f0 = field[TIX(l, k, iCell, nL, nK)];
                    double maxNeighbor = f0, minNeighbor = f0;
                    for (int iEdgeOnCell = 1; iEdgeOnCell <= n; iEdgeOnCell++) {
                        const int iCellNeighbor = A2(a->cellsOnCell, iEdgeOnCell, iCell, M);
                        if (mask[TIX(l, k, iCellNeighbor, nL, nK)] == 1) {
                            const double fn = field[TIX(l, k, iCellNeighbor, nL, nK)];
                            maxNeighbor = (maxNeighbor > fn) ? maxNeighbor : fn;
                            minNeighbor = (minNeighbor < fn) ? minNeighbor : fn;
                        }
                    }
                    maxNeighbor = maxNeighbor - f0;
                    minNeighbor = minNeighbor - f0;
                    double maxLocal = 0.0, minLocal = 0.0;
                    for (int iVertex = 1; iVertex <= n; iVertex++) {
                        const double deviationAtVertex = xg * (A2(a->xVertexOnCell, iVertex, iCell, M) - xB) +
                                                         yg * (A2(a->yVertexOnCell, iVertex, iCell, M) - yB);
                        maxLocal = (maxLocal > deviationAtVertex) ? maxLocal : deviationAtVertex;
                        minLocal = (minLocal < deviationAtVertex) ? minLocal : deviationAtVertex;
                    }
                    double gradFactor1, gradFactor2;
                    if (fabs(maxLocal) > fabs(maxNeighbor)) { gradFactor1 = maxNeighbor / maxLocal; if (!(gradFactor1 > 0.0)) gradFactor1 = 0.0; }
                    else gradFactor1 = 1.0;
                    if (fabs(minLocal) > fabs(minNeighbor)) { gradFactor2 = minNeighbor / minLocal; if (!(gradFactor2 > 0.0)) gradFactor2 = 0.0; }
                    else gradFactor2 = 1.0;
                    double gradFactor = (gradFactor1 < gradFactor2) ? gradFactor1 : gradFactor2;
                    gradFactor = gradFactor - EPS11;
                    if (!(gradFactor > 0.0)) gradFactor = 0.0;
                    xGrad[TIX(l, k, iCell, nL, nK)] = xg * gradFactor;
                    yGrad[TIX(l, k, iCell, nL, nK)] = yg * gradFactor;
                }
        }
        /* value at the cell centre (:3735, :3850, :3880), every cell */
        for (int iCell = 1; iCell <= nC; iCell++)
            for (int k = 1; k <= nK; k++)
                for (int l = 1; l <= nL; l++) {
                    const int lp = (pL == 1) ? 1 : l;
                    const double xB = p ? W[ip].xBary[TIX(lp, k, iCell, pL, nK)] : a->geomAvg[0][iCell - 1];
                    const double yB = p ? W[ip].yBary[TIX(lp, k, iCell, pL, nK)] : a->geomAvg[1][iCell - 1];
                    const size_t q = TIX(l, k, iCell, nL, nK);
                    W[t].center[q] = field[q] - xGrad[q] * xB - yGrad[q] * yB;
                }
        /* barycentre of mass * tracer chain, where a child needs it (:3750-3840, :3895-4170) */
        if (tr->hasChild) {
            if (tr->nParents >= 3) { err = IR_ERR_TOO_MANY_PARENTS; continue; }
            int chain[3], nChain = 0;   /* mass-like field first */
            for (int q = t; q >= 0; q = a->tracers[q].parent) chain[nChain++] = q;
            for (int iCell = 1; iCell <= nC; iCell++) {
                if (maskCell[iCell - 1] != 1) continue;
                for (int k = 1; k <= nK; k++)
                    for (int l = 1; l <= nL; l++) {
                        double mean[3], center[3], xg[3], yg[3];
                        for (int s = 0; s < nChain; s++) {
                            const int q = chain[nChain - 1 - s];
                            const int qL = a->tracers[q].nLayers;
                            const size_t ix = TIX((qL == 1) ? 1 : l, k, iCell, qL, nK);
                            mean[s] = a->tracers[q].array[ix];
                            center[s] = W[q].center[ix];
                            xg[s] = W[q].xGrad[ix];
                            yg[s] = W[q].yGrad[ix];
                        }
                        const size_t ix = TIX(l, k, iCell, nL, nK);
                        barycenter(a->geomAvg, iCell, &W[t].xBary[ix], &W[t].yBary[ix], nChain, mean, center, xg, yg);
                    }
            }
        }
    }
    if (a->xGradOut && a->gradTracerOut >= 0 && a->gradTracerOut < nT) {
        const size_t n = ((size_t)nC + 1) * nK * a->tracers[a->gradTracerOut].nLayers;
        memcpy(a->xGradOut, W[a->gradTracerOut].xGrad, n * sizeof(double));
        memcpy(a->yGradOut, W[a->gradTracerOut].yGrad, n * sizeof(double));
    }

    /* find_departure_points (:5255) */
    for (int v = 1; v <= nV; v++) {
        dpIn[((size_t)v - 1) * 2 + 0] = -a->uVelocity[v - 1] * a->dt;
        dpIn[((size_t)v - 1) * 2 + 1] = -a->vVelocity[v - 1] * a->dt;
    }

    /* find_departure_triangles (:5365): mask of edges, then the triangles and their quadrature points */
    for (int iEdge = 1; iEdge <= nE; iEdge++) {
        maskEdge[iEdge - 1] = 0;
        if (a->remapEdge[iEdge - 1] == 1)
            for (int k = 1; k <= 2; k++) {
                const int iVertex = A2(a->verticesOnEdge, k, iEdge, 2);
                const double lenSquared = sq(dpIn[((size_t)iVertex - 1) * 2]) + sq(dpIn[((size_t)iVertex - 1) * 2 + 1]);
                if (lenSquared > 0.0) maskEdge[iEdge - 1] = 1;
            }
    }
#pragma omp parallel for schedule(static)
    for (int iEdge = 1; iEdge <= nE; iEdge++) {
        if (maskEdge[iEdge - 1] != 1) continue;
        double *xT = &xTri[((size_t)iEdge - 1) * NTRI * nQP], *yT = &yTri[((size_t)iEdge - 1) * NTRI * nQP];
        const int e = departure_triangles_edge(a, iEdge, dpIn, xT, yT, nQP, &iCellTri[((size_t)iEdge - 1) * NTRI],
                                               &triArea[((size_t)iEdge - 1) * NTRI]);
        if (e) {
#pragma omp critical
            err = e;
        }
        quadrature_points_edge(xT, yT, nQP);
    }

    /* integrate_fluxes_over_triangles (:6667).  triangleValue of a tracer is the parent's triangleValue times the
     * tracer's linear reconstruction; instead of storing it per (edge, triangle, point) it is rebuilt down the chain of
     * parents, in the same order of multiplications. */
    for (int t = 0; t < nT; t++) {
        const orc_ir_tracer *tr = &a->tracers[t];
        const int nL = tr->nLayers;
        int chain[4], nChain = 0;
        for (int q = t; q >= 0; q = a->tracers[q].parent) chain[nChain++] = q;
        double *edgeFlux = W[t].edgeFlux;
        int negative = 0;
#pragma omp parallel for schedule(static) reduction(| : negative)
        for (int iEdge = 1; iEdge <= nE; iEdge++) {
            if (maskEdge[iEdge - 1] != 1) continue;
            for (int iTri = 1; iTri <= NTRI; iTri++) {
                const double area = triArea[((size_t)iEdge - 1) * NTRI + iTri - 1];
                if (area == 0.0) continue;
                const int iCell = iCellTri[((size_t)iEdge - 1) * NTRI + iTri - 1];
                const double *xq = &xTri[(((size_t)iEdge - 1) * NTRI + iTri - 1) * nQP];
                const double *yq = &yTri[(((size_t)iEdge - 1) * NTRI + iTri - 1) * nQP];
                for (int k = 1; k <= nK; k++)
                    for (int l = 1; l <= nL; l++) {
                        double tracerIntegral = 0.0;
                        for (int iqp = 0; iqp < nQP; iqp++) {
                            double value = 1.0;
                            for (int s = nChain - 1; s >= 0; s--) {
                                const int q = chain[s];
                                const int qL = a->tracers[q].nLayers;
                                const size_t ix = TIX((qL == 1) ? 1 : l, k, iCell, qL, nK);
                                value = value * (W[q].center[ix] + W[q].xGrad[ix] * xq[iqp] + W[q].yGrad[ix] * yq[iqp]);
                            }
                            if (t == 0 && value < 0.0) negative = 1;
                            tracerIntegral = tracerIntegral + weightQuadPoint[iqp] * value;
                        }
                        edgeFlux[TIX(l, k, iEdge, nL, nK)] = edgeFlux[TIX(l, k, iEdge, nL, nK)] + area * tracerIntegral;
                    }
            }
        }
        if (negative) err = IR_ERR_NEGATIVE_MASS_QP;
    }

    /* compute_mass_tracer_products (:6982) */
    for (int t = 0; t < nT; t++) {
        const orc_ir_tracer *tr = &a->tracers[t];
        const int nL = tr->nLayers, ip = tr->parent;
        const int pL = (ip >= 0) ? a->tracers[ip].nLayers : 1;
        for (int c = 1; c <= nC; c++)
            for (int k = 1; k <= nK; k++)
                for (int l = 1; l <= nL; l++) {
                    const double pm = (ip >= 0) ? W[ip].mtp[TIX((pL == 1) ? 1 : l, k, c, pL, nK)] : 1.0;
                    W[t].mtp[TIX(l, k, c, nL, nK)] = pm * tr->array[TIX(l, k, c, nL, nK)];
                }
    }

    /* config_conservation_check: sum_tracers(init = .true.) (:3259-3268, :7998) */
    size_t *sumOff = (size_t *)calloc((size_t)nT + 1, sizeof(size_t));
    for (int t = 0; t < nT; t++) sumOff[t + 1] = sumOff[t] + (size_t)nK * a->tracers[t].nLayers;
    double *sumInit = (double *)calloc(sumOff[nT], sizeof(double)), *sumFinal = (double *)calloc(sumOff[nT], sizeof(double));
    if (a->conservationCheck)
        for (int t = 0; t < nT; t++) {
            const int nL = a->tracers[t].nLayers;
            for (int iCell = 1; iCell <= a->nCellsSolve; iCell++)
                for (int k = 1; k <= nK; k++)
                    for (int l = 1; l <= nL; l++)
                        sumInit[sumOff[t] + (size_t)(k - 1) * nL + (l - 1)] =
                            sumInit[sumOff[t] + (size_t)(k - 1) * nL + (l - 1)] + a->areaCell[iCell - 1] * W[t].mtp[TIX(l, k, iCell, nL, nK)];
        }

    /* update_mass_and_tracers (:7125) */
    for (int t = 0; t < nT && err != IR_ERR_NEGATIVE_MASS; t++) {
        orc_ir_tracer *tr = &a->tracers[t];
        const int nL = tr->nLayers, ip = tr->parent;
        const int pL = (ip >= 0) ? a->tracers[ip].nLayers : 1;
#pragma omp parallel for schedule(static)
        for (int iCell = 1; iCell <= a->nCellsSolve; iCell++) {
            for (int k = 1; k <= nK; k++)
                for (int l = 1; l <= nL; l++) {
                    double fluxFromCell = 0.0;
                    for (int iEdgeOnCell = 1; iEdgeOnCell <= a->nEdgesOnCell[iCell - 1]; iEdgeOnCell++) {
                        const int iEdge = A2(a->edgesOnCell, iEdgeOnCell, iCell, M);
                        const int edgeSignOnCell = (iCell == A2(a->cellsOnEdge, 1, iEdge, 2)) ? 1 : -1;
                        fluxFromCell = fluxFromCell + W[t].edgeFlux[TIX(l, k, iEdge, nL, nK)] * edgeSignOnCell;
                    }
                    const double pm = (ip >= 0) ? W[ip].mtp[TIX((pL == 1) ? 1 : l, k, iCell, pL, nK)] : 1.0;
                    const size_t q = TIX(l, k, iCell, nL, nK);
                    if (pm > 0.0) tr->array[q] = (W[t].mtp[q] - (fluxFromCell / a->areaCell[iCell - 1])) / pm;
                    else tr->array[q] = 0.0;
                    W[t].mtp[q] = pm * tr->array[q];
                }
        }
        if (tr->nParents == 0) {
            const double puny2 = 1.0e-11 * 1.0e-11; /* seaicePuny**2, constants.F */
            for (int iCell = 1; iCell <= a->nCellsSolve && err != IR_ERR_NEGATIVE_MASS; iCell++)
                for (int k = 1; k <= nK; k++)
                    for (int l = 1; l <= nL; l++) {
                        double *v = &tr->array[TIX(l, k, iCell, nL, nK)];
                        if (*v < -puny2) { err = IR_ERR_NEGATIVE_MASS; break; }
                        else if (*v >= -puny2 && *v < 0.0) *v = 0.0;
                    }
        }
    }

    /* config_conservation_check: compute_mass_tracer_products on the new values, sum_tracers(init = .false.) (:3298-3310) */
    if (a->conservationCheck && err != IR_ERR_NEGATIVE_MASS)
        for (int t = 0; t < nT; t++) {
            const orc_ir_tracer *tr = &a->tracers[t];
            const int nL = tr->nLayers, ip = tr->parent;
            const int pL = (ip >= 0) ? a->tracers[ip].nLayers : 1;
            for (int c = 1; c <= nC; c++)
                for (int k = 1; k <= nK; k++)
                    for (int l = 1; l <= nL; l++) {
                        const double pm = (ip >= 0) ? W[ip].mtp[TIX((pL == 1) ? 1 : l, k, c, pL, nK)] : 1.0;
                        W[t].mtp[TIX(l, k, c, nL, nK)] = pm * tr->array[TIX(l, k, c, nL, nK)];
                    }
            for (int iCell = 1; iCell <= a->nCellsSolve; iCell++)
                for (int k = 1; k <= nK; k++)
                    for (int l = 1; l <= nL; l++)
                        sumFinal[sumOff[t] + (size_t)(k - 1) * nL + (l - 1)] =
                            sumFinal[sumOff[t] + (size_t)(k - 1) * nL + (l - 1)] + a->areaCell[iCell - 1] * W[t].mtp[TIX(l, k, iCell, nL, nK)];
        }

    /* zap_small_mass (:8764); the reference handles a one-layer mass-like field only */
    if (mass->nLayers == 1 && err != IR_ERR_NEGATIVE_MASS) {
        const double smallMassThreshold = 1.0e-22;
        for (int iCell = 1; iCell <= a->nCellsSolve; iCell++)
            for (int k = 1; k <= nK; k++) {
                double *m = &mass->array[TIX(1, k, iCell, 1, nK)];
                if (*m > 0.0 && *m < smallMassThreshold) {
                    *m = 0.0;
                    for (int t = 1; t < nT; t++)
                        for (int l = 1; l <= a->tracers[t].nLayers; l++)
                            a->tracers[t].array[TIX(l, k, iCell, a->tracers[t].nLayers, nK)] = 0.0;
                }
            }
    }

    /* check_tracer_conservation (:8126; the sums of the ranks are added first, :8150 -- the caller's job with
     * conservationCheck = 2).  First violation in the reference's loop order: tracer, category, layer. */
    if (a->sumInitOut) memcpy(a->sumInitOut, sumInit, sumOff[nT] * sizeof(double));
    if (a->sumFinalOut) memcpy(a->sumFinalOut, sumFinal, sumOff[nT] * sizeof(double));
    if (a->consErrOut) a->consErrOut[0] = a->consErrOut[1] = a->consErrOut[2] = a->consErrOut[3] = 0;
    if (a->monoErrOut) a->monoErrOut[0] = a->monoErrOut[1] = a->monoErrOut[2] = a->monoErrOut[3] = a->monoErrOut[4] = 0;
    int consViolated = 0;
    if (a->conservationCheck == 1 && err == IR_OK) {
        for (int t = 0; t < nT && !consViolated; t++) {
            const int nL = a->tracers[t].nLayers;
            for (int k = 1; k <= nK && !consViolated; k++)
                for (int l = 1; l <= nL; l++) {
                    const double si = sumInit[sumOff[t] + (size_t)(k - 1) * nL + (l - 1)], sf = sumFinal[sumOff[t] + (size_t)(k - 1) * nL + (l - 1)];
                    if (fabs(si) > EPS11) {
                        const double difference = sf - si;
                        const double ratio = difference / si;
                        if (fabs(ratio) > EPS11) {
                            consViolated = 1;
                            if (a->consErrOut) { a->consErrOut[0] = 1; a->consErrOut[1] = t; a->consErrOut[2] = k; a->consErrOut[3] = l; }
                            break;
                        }
                    }
                }
        }
        if (consViolated) err = IR_ERR_CONSERVATION;
    }
    /* check_tracer_monotonicity (:8416), not reached when the conservation check has aborted (:2577-2581).  The
     * masks are those of the last make_masks call (threshold eps11, OLD values); extendedMinMax = .true. */
    if (a->monotonicityCheck && err == IR_OK) {
        int found = 0;
        for (int t = 0; t < nT && !found; t++) {
            const orc_ir_tracer *tr = &a->tracers[t];
            if (tr->nParents == 0) continue;      /* monotonicity holds for tracers but not for the mass-like field */
            const int nL = tr->nLayers;
            const int *mask = W[t].mask;
            const size_t n = ((size_t)nC + 1) * nK * nL;
            double *emin = (double *)malloc(n * sizeof(double)), *emax = (double *)malloc(n * sizeof(double));
            memcpy(emin, lmin[t], n * sizeof(double));
            memcpy(emax, lmax[t], n * sizeof(double));
            /* the reference extends in place, cell after cell (:8478-8500): a cell updated earlier in the loop is seen
             * in its extended state by a later neighbour, so its bounds depend on the cell numbering (and are never
             * tighter than the two-ring bounds).  monotonicityCheck = 2 is the order-independent variant the parallel
             * device kernel computes: neighbours are read in their un-extended state. */
            for (int iCell = 1; iCell <= a->nCellsSolve; iCell++)
                for (int k = 1; k <= nK; k++)
                    for (int l = 1; l <= nL; l++) {
                        const size_t q = TIX(l, k, iCell, nL, nK);
                        if (mask[q] != 1) continue;
                        for (int iCellOnCell = 1; iCellOnCell <= a->nEdgesOnCell[iCell - 1]; iCellOnCell++) {
                            const int iCellNeighbor = A2(a->cellsOnCell, iCellOnCell, iCell, M);
                            if (iCellNeighbor >= 1 && iCellNeighbor <= nC) {
                                const size_t qn = TIX(l, k, iCellNeighbor, nL, nK);
                                if (mask[qn] == 1) {
                                    const double nmin = (a->monotonicityCheck == 2) ? lmin[t][qn] : emin[qn];
                                    const double nmax = (a->monotonicityCheck == 2) ? lmax[t][qn] : emax[qn];
                                    if (nmin < emin[q]) emin[q] = nmin;
                                    if (nmax > emax[q]) emax[q] = nmax;
                                }
                            }
                        }
                    }
            for (int iCell = 1; iCell <= a->nCellsSolve && !found; iCell++)
                for (int k = 1; k <= nK && !found; k++)
                    for (int l = 1; l <= nL; l++) {
                        const size_t q = TIX(l, k, iCell, nL, nK);
                        if (mask[q] != 1) continue;
                        const double toleranceMin = EPS11 * fmax(1.0, fabs(emin[q]));
                        const double toleranceMax = EPS11 * fmax(1.0, fabs(emax[q]));
                        int which = 0;
                        if (tr->array[q] < emin[q] - toleranceMin) which = 1;
                        else if (tr->array[q] > emax[q] + toleranceMax) which = 2;
                        if (which) {
                            found = 1;
                            if (a->monoErrOut) { a->monoErrOut[0] = which; a->monoErrOut[1] = t; a->monoErrOut[2] = l; a->monoErrOut[3] = k; a->monoErrOut[4] = iCell; }
                            if (a->monoValOut) { a->monoValOut[0] = tr->array[q]; a->monoValOut[1] = which == 1 ? emin[q] : emax[q]; a->monoValOut[2] = which == 1 ? toleranceMin : toleranceMax; }
                            break;
                        }
                    }
            free(emin); free(emax);
        }
        if (found) err = IR_ERR_MONOTONICITY;
    }
    if (lmin) { for (int t = 0; t < nT; t++) { free(lmin[t]); free(lmax[t]); } free(lmin); free(lmax); }
    free(sumOff); free(sumInit); free(sumFinal);

    /* thickness -> volume (:2680-2700) with the new area */
    for (int t = 0; t < nT; t++) {
        orc_ir_tracer *tr = &a->tracers[t];
        if (!tr->volumeLike) continue;
        for (int c = 1; c <= nC + 1; c++)
            for (int k = 1; k <= nK; k++) {
                double *v = &tr->array[TIX(1, k, c, tr->nLayers, nK)];
                *v = mass->array[TIX(1, k, c, mass->nLayers, nK)] * *v;
            }
    }

    if (a->xTriangleOut) memcpy(a->xTriangleOut, xTri, sizeof(double) * (size_t)nE * NTRI * nQP);
    if (a->yTriangleOut) memcpy(a->yTriangleOut, yTri, sizeof(double) * (size_t)nE * NTRI * nQP);
    if (a->triangleAreaOut) memcpy(a->triangleAreaOut, triArea, sizeof(double) * (size_t)nE * NTRI);
    if (a->iCellTriangleOut) memcpy(a->iCellTriangleOut, iCellTri, sizeof(int) * (size_t)nE * NTRI);
    if (a->maskEdgeOut) memcpy(a->maskEdgeOut, maskEdge, sizeof(int) * (size_t)nE);
    if (a->edgeFluxMassOut) memcpy(a->edgeFluxMassOut, W[0].edgeFlux, sizeof(double) * (size_t)nE * nK * mass->nLayers);

    for (int t = 0; t < nT; t++) {
        free(W[t].mask); free(W[t].center); free(W[t].xGrad); free(W[t].yGrad);
        free(W[t].xBary); free(W[t].yBary); free(W[t].mtp); free(W[t].edgeFlux);
    }
    free(W); free(maskCell); free(maskEdge); free(dpIn); free(xTri); free(yTri); free(triArea); free(iCellTri);
    return err;
}
