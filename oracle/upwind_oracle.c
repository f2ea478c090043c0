/* oracle/upwind_oracle.c -- CPU restatement of two non-default corners next to the hot path:
 *
 *   (1) seaice_normal_vectors            src/shared/mpas_seaice_mesh.F:703-846
 *         normal_vectors_planar_polygon  :858-943      normal_vectors_planar_triangle            :957-1024
 *         normal_vectors_spherical_polygon_metric :1038-1241   normal_vectors_spherical_triangle_metric :1393-1606
 *         seaice_dot_product_3space :1758   cross_product_3space :1783   seaice_grid_rotation_forward :2350
 *       -- the init-time geometry of the weak operators (weak.F:84-96, removeMetricTerms = .true.) and of the upwind
 *       transport (advection_upwind.F:122-126, removeMetricTerms = .false.);
 *   (2) seaice_run_advection_upwind      src/shared/mpas_seaice_advection_upwind.F:385-520
 *         prepare_advection :1549   initialize_timelevel_variables :1633   edge_from_vertex_velocity :1403
 *         prepare_tracers :1890   run_advection_variable_3D :638   prepare_none_parent_tracer :819
 *         run_advection_subvariable :1048   upwind_tendencies :1242   finalize_tracers :1989
 *         scale_tracers_back_3D :2199
 *       -- config_advection_type = 'upwind', one block, the tracer halo exchange (:1786) left to the caller.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may call this; the product never
 * does.  Build with -ffp-contract=off: the reference's operation order, one IEEE operation at a time.
 *
 * Parity status: the reference cannot be compiled in this image; both parts are pinned by OUTPUTS OF THE
 * REFERENCE'S OWN SOURCE EXECUTED HERE by the Fortran-subset interpreter tests/golden/fortran_subset.py (mesh.F's
 * seaice_normal_vectors; advection_upwind.F's define_tracer_connectivities and seaice_run_advection_upwind with the
 * module's own table and parameters): fixtures tests/golden/options/refexec_*.npz, reproduced by this file bit for bit
 * (tests/test_transport_options.py).  Also: an independent vectorised reading of the normal vectors
 * (mpas-seaice_b200/weakmesh.py), the reference's analytic operator fields through the weak operators
 * (tests/test_analytic_golden.py), conservation, uniform tracers and the closed-form donor-cell update.
 *
 * The upwind module is restated AS EXECUTED, with the tracer connectivity table as an input.  What the reference's own
 * table does (define_tracer_connectivities :145-170) is noted in tests/test_transport_options.py: the chain
 * iceAreaCategory -> surfaceTemperature -> iceVolumeCategory -> snowVolumeCategory makes the (negative) area-weighted
 * surface temperature the "mass" that carries the ice volume, so `parentTracerNew > 0` never holds and the volumes come
 * out zero.  Quirks kept: the update loop runs over nCells (the variable named nCellsSolve is read from the dimension
 * "nCells", :1136), the tendencies over nCellsSolve (:1310); volume -> thickness only where the area exceeds
 * iceAreaMinimum (:1952), the values elsewhere are transported as they are.
 *
 * Array conventions as everywhere in this repo: 1-based index VALUES, numpy C order == Fortran order with the
 * dimensions reversed, a junk slot n+1 at the end of cell / vertex / edge arrays.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define A2(a, i, j, n1) ((a)[((size_t)(j) - 1) * (size_t)(n1) + ((size_t)(i) - 1)]) /* a(i,j), first dim n1 */
#define A3(a, i, j, k, n1, n2) ((a)[(((size_t)(k) - 1) * (size_t)(n2) + ((size_t)(j) - 1)) * (size_t)(n1) + ((size_t)(i) - 1)])

/* ------------------------------------------------------------------------------------------ (1) normal vectors */

typedef struct {
    int nCells, nVertices, nVerticesSolve, nEdges, maxEdges, vertexDegree;
    int on_a_sphere, rotate_cartesian_grid, removeMetricTerms;
    double sphere_radius;
    const int *nEdgesOnCell, *edgesOnCell, *verticesOnEdge, *cellsOnEdge, *edgesOnVertex, *interiorVertex;
    const double *xCell, *yCell, *zCell, *xVertex, *yVertex, *zVertex, *xEdge, *yEdge, *zEdge;
    /* outputs; the triangle part is skipped when normalVectorTriangle is NULL (seaice_normal_vectors_polygon alone) */
    double *normalVectorPolygon;  /* (nCells+1, maxEdges, 2) */
    double *normalVectorTriangle; /* (nVertices+1, vertexDegree, 2) */
    double *latCellRotated;       /* (nCells+1), may be NULL */
    double *latVertexRotated;     /* (nVertices+1), may be NULL */
} orc_normals_args;

static void grid_rotation_forward(double *p, double x, double y, double z, int rotate) /* mesh.F:2350 */
{
    if (rotate) { p[0] = -z; p[1] = y; p[2] = x; }
    else { p[0] = x; p[1] = y; p[2] = z; }
}

/* matmul(yRotationMatrix, matmul(zRotationMatrix, p)), mesh.F:1176 */
static void to_equator(const double yR[3][3], const double zR[3][3], const double *p, double *o)
{
    double t[3];
    for (int i = 0; i < 3; i++) t[i] = zR[i][0] * p[0] + zR[i][1] * p[1] + zR[i][2] * p[2];
    for (int i = 0; i < 3; i++) o[i] = yR[i][0] * t[0] + yR[i][1] * t[1] + yR[i][2] * t[2];
}

static void rotation_matrices(double yR[3][3], double zR[3][3], double lat, double lon, int removeMetricTerms)
{
    memset(yR, 0, 9 * sizeof(double));
    memset(zR, 0, 9 * sizeof(double));
    yR[1][1] = 1.0;
    zR[2][2] = 1.0;
    if (removeMetricTerms) {
        yR[0][0] = cos(lat); yR[0][2] = sin(lat); yR[2][0] = -sin(lat); yR[2][2] = cos(lat);
        zR[0][0] = cos(-lon); zR[0][1] = -sin(-lon); zR[1][0] = sin(-lon); zR[1][1] = cos(-lon);
    } else {
        yR[0][0] = 1.0; yR[2][2] = 1.0;
        zR[0][0] = 1.0; zR[1][1] = 1.0;
    }
}

/* the common tail of the two spherical routines (:1190-1225, :1560-1595): the unit normal of the great circle through
 * the side, flipped on request, expressed by its eastward component and the signed remainder */
static void side_components(const double *sideVector, const double *edgeEquator, int flip, double *n1, double *n2)
{
    double g[3];
    g[0] = sideVector[1] * edgeEquator[2] - sideVector[2] * edgeEquator[1];
    g[1] = sideVector[2] * edgeEquator[0] - sideVector[0] * edgeEquator[2];
    g[2] = sideVector[0] * edgeEquator[1] - sideVector[1] * edgeEquator[0];
    if (flip) { g[0] = -1.0 * g[0]; g[1] = -1.0 * g[1]; g[2] = -1.0 * g[2]; }
    const double norm = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    g[0] = g[0] / norm; g[1] = g[1] / norm; g[2] = g[2] / norm;
    double e[3] = {-edgeEquator[1], edgeEquator[0], 0.0};
    const double en = sqrt(e[0] * e[0] + e[1] * e[1]);
    e[0] = e[0] / en; e[1] = e[1] / en; e[2] = e[2] / en;
    *n1 = g[0] * e[0] + g[1] * e[1] + g[2] * e[2];
    const double clipped = fmax(fmin(*n1, 1.0), -1.0);
    *n2 = copysign(1.0, g[2]) * sqrt(1.0 - clipped * clipped);
}

int orc_normal_vectors(orc_normals_args *a)
{
    const int nC = a->nCells, nV = a->nVertices, M = a->maxEdges, D = a->vertexDegree;
    memset(a->normalVectorPolygon, 0, sizeof(double) * ((size_t)nC + 1) * M * 2);
    if (a->normalVectorTriangle) memset(a->normalVectorTriangle, 0, sizeof(double) * ((size_t)nV + 1) * D * 2);
    if (a->latCellRotated) memset(a->latCellRotated, 0, sizeof(double) * ((size_t)nC + 1));
    if (a->latVertexRotated) memset(a->latVertexRotated, 0, sizeof(double) * ((size_t)nV + 1));
    if (!a->on_a_sphere) {
        /* normal_vectors_planar_polygon (:858) */
        for (int iCell = 1; iCell <= nC; iCell++)
            for (int iEdgeOnCell = 1; iEdgeOnCell <= a->nEdgesOnCell[iCell - 1]; iEdgeOnCell++) {
                const int iEdge = A2(a->edgesOnCell, iEdgeOnCell, iCell, M);
                const int iVertex1 = A2(a->verticesOnEdge, 1, iEdge, 2), iVertex2 = A2(a->verticesOnEdge, 2, iEdge, 2);
                double tx = a->xVertex[iVertex2 - 1] - a->xVertex[iVertex1 - 1];
                double ty = a->yVertex[iVertex2 - 1] - a->yVertex[iVertex1 - 1];
                const double tmag = sqrt(tx * tx + ty * ty);
                tx = tx / tmag;
                ty = ty / tmag;
                const double nx = a->xEdge[iEdge - 1] - a->xCell[iCell - 1], ny = a->yEdge[iEdge - 1] - a->yCell[iCell - 1];
                if ((nx * ty - ny * tx) < 0.0) { tx = -tx; ty = -ty; }
                A3(a->normalVectorPolygon, 1, iEdgeOnCell, iCell, 2, M) = ty;
                A3(a->normalVectorPolygon, 2, iEdgeOnCell, iCell, 2, M) = -tx;
            }
        /* normal_vectors_planar_triangle (:957) */
        if (a->normalVectorTriangle)
            for (int iVertex = 1; iVertex <= nV; iVertex++) {
                if (a->interiorVertex[iVertex - 1] != 1) continue;
                for (int iVertexDegree = 1; iVertexDegree <= D; iVertexDegree++) {
                    const int iEdge = A2(a->edgesOnVertex, iVertexDegree, iVertex, D);
                    const double dx = a->xEdge[iEdge - 1] - a->xVertex[iVertex - 1], dy = a->yEdge[iEdge - 1] - a->yVertex[iVertex - 1];
                    A3(a->normalVectorTriangle, 1, iVertexDegree, iVertex, 2, D) = dx / sqrt(dx * dx + dy * dy);
                    A3(a->normalVectorTriangle, 2, iVertexDegree, iVertex, 2, D) = dy / sqrt(dx * dx + dy * dy);
                }
            }
        return 0;
    }
    double yR[3][3], zR[3][3];
    const int rot = a->rotate_cartesian_grid;
    /* normal_vectors_spherical_polygon_metric (:1038) */
    for (int iCell = 1; iCell <= nC; iCell++) {
        double cellCentreRotated[3];
        grid_rotation_forward(cellCentreRotated, a->xCell[iCell - 1], a->yCell[iCell - 1], a->zCell[iCell - 1], rot);
        const double lonCellRotated = atan2(cellCentreRotated[1], cellCentreRotated[0]);
        const double latCellRotated = asin(cellCentreRotated[2] / a->sphere_radius);
        rotation_matrices(yR, zR, latCellRotated, lonCellRotated, a->removeMetricTerms);
        for (int iEdgeOnCell = 1; iEdgeOnCell <= a->nEdgesOnCell[iCell - 1]; iEdgeOnCell++) {
            const int iEdge = A2(a->edgesOnCell, iEdgeOnCell, iCell, M);
            const int iVertex1 = A2(a->verticesOnEdge, 1, iEdge, 2), iVertex2 = A2(a->verticesOnEdge, 2, iEdge, 2);
            double edgeRotated[3], vertexRotated1[3], vertexRotated2[3], edgeEquator[3], vertexEquator1[3], vertexEquator2[3];
            grid_rotation_forward(edgeRotated, a->xEdge[iEdge - 1], a->yEdge[iEdge - 1], a->zEdge[iEdge - 1], rot);
            grid_rotation_forward(vertexRotated1, a->xVertex[iVertex1 - 1], a->yVertex[iVertex1 - 1], a->zVertex[iVertex1 - 1], rot);
            grid_rotation_forward(vertexRotated2, a->xVertex[iVertex2 - 1], a->yVertex[iVertex2 - 1], a->zVertex[iVertex2 - 1], rot);
            to_equator(yR, zR, edgeRotated, edgeEquator);
            to_equator(yR, zR, vertexRotated1, vertexEquator1);
            to_equator(yR, zR, vertexRotated2, vertexEquator2);
            const double vertexVector[3] = {vertexEquator2[0] - vertexEquator1[0], vertexEquator2[1] - vertexEquator1[1],
                                            vertexEquator2[2] - vertexEquator1[2]};
            side_components(vertexVector, edgeEquator, iCell == A2(a->cellsOnEdge, 2, iEdge, 2),
                            &A3(a->normalVectorPolygon, 1, iEdgeOnCell, iCell, 2, M),
                            &A3(a->normalVectorPolygon, 2, iEdgeOnCell, iCell, 2, M));
        }
        if (a->latCellRotated) a->latCellRotated[iCell - 1] = latCellRotated;
    }
    /* normal_vectors_spherical_triangle_metric (:1393) */
    if (a->normalVectorTriangle)
        for (int iVertex = 1; iVertex <= a->nVerticesSolve; iVertex++) {
            if (a->interiorVertex[iVertex - 1] != 1) continue;
            double vertexRotated[3];
            grid_rotation_forward(vertexRotated, a->xVertex[iVertex - 1], a->yVertex[iVertex - 1], a->zVertex[iVertex - 1], rot);
            const double lonVertexRotated = atan2(vertexRotated[1], vertexRotated[0]);
            const double latVertexRotated = asin(vertexRotated[2] / a->sphere_radius);
            rotation_matrices(yR, zR, latVertexRotated, lonVertexRotated, a->removeMetricTerms);
            for (int iVertexDegree = 1; iVertexDegree <= D; iVertexDegree++) {
                const int iEdge = A2(a->edgesOnVertex, iVertexDegree, iVertex, D);
                const int iCell1 = A2(a->cellsOnEdge, 1, iEdge, 2), iCell2 = A2(a->cellsOnEdge, 2, iEdge, 2);
                double edgeRotated[3], cellRotated1[3], cellRotated2[3], edgeEquator[3], cellEquator1[3], cellEquator2[3];
                grid_rotation_forward(edgeRotated, a->xEdge[iEdge - 1], a->yEdge[iEdge - 1], a->zEdge[iEdge - 1], rot);
                grid_rotation_forward(cellRotated1, a->xCell[iCell1 - 1], a->yCell[iCell1 - 1], a->zCell[iCell1 - 1], rot);
                grid_rotation_forward(cellRotated2, a->xCell[iCell2 - 1], a->yCell[iCell2 - 1], a->zCell[iCell2 - 1], rot);
                to_equator(yR, zR, edgeRotated, edgeEquator);
                to_equator(yR, zR, cellRotated1, cellEquator1);
                to_equator(yR, zR, cellRotated2, cellEquator2);
                const double cellVector[3] = {cellEquator2[0] - cellEquator1[0], cellEquator2[1] - cellEquator1[1],
                                              cellEquator2[2] - cellEquator1[2]};
                side_components(cellVector, edgeEquator, iVertex == A2(a->verticesOnEdge, 1, iEdge, 2),
                                &A3(a->normalVectorTriangle, 1, iVertexDegree, iVertex, 2, D),
                                &A3(a->normalVectorTriangle, 2, iVertexDegree, iVertex, 2, D));
            }
            if (a->latVertexRotated) a->latVertexRotated[iVertex - 1] = latVertexRotated;
        }
    return 0;
}

/* ------------------------------------------------------------------------------------------ (2) upwind transport */

#define ICE_AREA_MINIMUM 1.0e-11 /* iceAreaMinimum = seaicePuny, src/shared/mpas_seaice_constants.F:90 */

typedef struct {
    int parent;          /* index into the table, -1 = 'none' (:710); a parent comes before its children (:480 runs the table in order) */
    double childMinimum; /* childTracerMinimum (:235), 0 in the reference's table */
    int volumeLike;      /* iceVolumeCategory / snowVolumeCategory: divided by variable 0 before, multiplied after (:1949, :2063) */
    double *array;       /* (nCells+1, nCategories): time level 1 on entry, the new time level on exit */
    double *edgeFluxOut; /* (nEdges+1, nCategories) or NULL: <name>EdgeFlux */
} orc_upwind_var;

typedef struct {
    int nCells, nCellsSolve, nVertices, nEdges, maxEdges, nCategories;
    double dt;
    const int *nEdgesOnCell, *edgesOnCell, *cellsOnCell, *cellsOnEdge, *verticesOnEdge, *interiorEdge;
    const double *areaCell, *dvEdge;
    const double *normalVectorEdge; /* (nCells+1, maxEdges, 2) */
    const double *uVelocity, *vVelocity;
    int nVars;
    orc_upwind_var *vars;
    double *edgeVelocityOut; /* (nEdges+1) or NULL */
} orc_upwind_args;

#define K2(a, k, c) ((a)[((size_t)(c) - 1) * (size_t)nK + ((size_t)(k) - 1)]) /* a(1,k,c) */

int orc_upwind_run(orc_upwind_args *a)
{
    const int nC = a->nCells, nCS = a->nCellsSolve, nE = a->nEdges, M = a->maxEdges, nK = a->nCategories, nT = a->nVars;
    if (nT < 1 || a->vars[0].parent != -1) return 8;
    for (int t = 0; t < nT; t++)
        if (a->vars[t].parent >= t || a->vars[t].parent < -1) return 8;
    const double dt = a->dt;
    const size_t nCell = ((size_t)nC + 1) * nK, nEdge = ((size_t)nE + 1) * nK;

    /* prepare_advection (:1549): the new time level, tendencies and edge fluxes start from zero (:1633) */
    double **newv = (double **)calloc((size_t)nT, sizeof(double *)), **tend = (double **)calloc((size_t)nT, sizeof(double *));
    double **eflux = (double **)calloc((size_t)nT, sizeof(double *));
    for (int t = 0; t < nT; t++) {
        newv[t] = (double *)calloc(nCell, sizeof(double));
        tend[t] = (double *)calloc(nCell, sizeof(double));
        eflux[t] = (double *)calloc(nEdge, sizeof(double));
    }
    /* edge_from_vertex_velocity (:1403) */
    double *edgeVelocity = (double *)calloc((size_t)nE + 1, sizeof(double));
    for (int iCell = 1; iCell <= nC; iCell++)
        for (int iEdgeOnCell = 1; iEdgeOnCell <= a->nEdgesOnCell[iCell - 1]; iEdgeOnCell++) {
            const int iEdge = A2(a->edgesOnCell, iEdgeOnCell, iCell, M);
            if (A2(a->cellsOnEdge, 1, iEdge, 2) == iCell) {
                double uVelocityEdge = 0.0, vVelocityEdge = 0.0;
                for (int iVertexOnEdge = 1; iVertexOnEdge <= 2; iVertexOnEdge++) {
                    const int iVertex = A2(a->verticesOnEdge, iVertexOnEdge, iEdge, 2);
                    uVelocityEdge = uVelocityEdge + a->uVelocity[iVertex - 1];
                    vVelocityEdge = vVelocityEdge + a->vVelocity[iVertex - 1];
                }
                uVelocityEdge = uVelocityEdge / 2.0;
                vVelocityEdge = vVelocityEdge / 2.0;
                edgeVelocity[iEdge - 1] = uVelocityEdge * A3(a->normalVectorEdge, 1, iEdgeOnCell, iCell, 2, M) +
                                          vVelocityEdge * A3(a->normalVectorEdge, 2, iEdgeOnCell, iCell, 2, M);
            }
        }
    /* prepare_tracers (:1890): volume -> thickness where the area exceeds iceAreaMinimum */
    for (int t = 1; t < nT; t++)
        if (a->vars[t].volumeLike)
            for (int iCell = 1; iCell <= nC; iCell++)
                for (int k = 1; k <= nK; k++)
                    if (K2(a->vars[0].array, k, iCell) > ICE_AREA_MINIMUM)
                        K2(a->vars[t].array, k, iCell) = K2(a->vars[t].array, k, iCell) / K2(a->vars[0].array, k, iCell);
    /* (halo_exchange_advection :1786 is the caller's) */

    double *pOldNone = (double *)calloc(nCell, sizeof(double)), *pNewNone = (double *)calloc(nCell, sizeof(double));
    double *pFluxNone = (double *)calloc(nEdge, sizeof(double));
    for (int t = 0; t < nT; t++) {
        orc_upwind_var *var = &a->vars[t];
        double *cOld = var->array, *cNew = newv[t], *cTend = tend[t], *cFlux = eflux[t];
        const double *pOld, *pNew, *pFlux;
        double parentTracerMinimum;
        if (var->parent < 0) {
            /* prepare_none_parent_tracer (:819) */
            for (int iCell = 1; iCell <= nC; iCell++)
                for (int k = 1; k <= nK; k++) {
                    K2(pOldNone, k, iCell) = 0.0;
                    K2(pNewNone, k, iCell) = 0.0;
                    if (K2(cOld, k, iCell) > ICE_AREA_MINIMUM) { K2(pOldNone, k, iCell) = 1.0; K2(pNewNone, k, iCell) = 1.0; }
                    for (int iEdgeOnCell = 1; iEdgeOnCell <= a->nEdgesOnCell[iCell - 1]; iEdgeOnCell++)
                        if (K2(cOld, k, A2(a->cellsOnCell, iEdgeOnCell, iCell, M)) > ICE_AREA_MINIMUM) {
                            K2(pOldNone, k, iCell) = 1.0;
                            K2(pNewNone, k, iCell) = 1.0;
                            break;
                        }
                }
            for (int iEdge = 1; iEdge <= nE; iEdge++)
                for (int k = 1; k <= nK; k++) {
                    if (K2(cOld, k, A2(a->cellsOnEdge, 1, iEdge, 2)) > ICE_AREA_MINIMUM ||
                        K2(cOld, k, A2(a->cellsOnEdge, 2, iEdge, 2)) > ICE_AREA_MINIMUM)
                        K2(pFluxNone, k, iEdge) = edgeVelocity[iEdge - 1];
                    else
                        K2(pFluxNone, k, iEdge) = 0.0;
                }
            pOld = pOldNone; pNew = pNewNone; pFlux = pFluxNone;
            parentTracerMinimum = 0.0;                       /* add_parent_tracer_minimums (:342) */
        } else {
            pOld = a->vars[var->parent].array;               /* time level 1 of the parent: already scaled by ITS parent (:1232) */
            pNew = newv[var->parent];
            pFlux = eflux[var->parent];
            parentTracerMinimum = a->vars[var->parent].childMinimum;
        }
        /* run_advection_subvariable (:1048) -> upwind_tendencies (:1242) */
        for (int iCell = 1; iCell <= nCS; iCell++) {
            const double invAreaCell1 = 1.0 / a->areaCell[iCell - 1];
            for (int iEdgeOnCell = 1; iEdgeOnCell <= a->nEdgesOnCell[iCell - 1]; iEdgeOnCell++) {
                const int iEdge = A2(a->edgesOnCell, iEdgeOnCell, iCell, M);
                const int cell1 = A2(a->cellsOnEdge, 1, iEdge, 2), cell2 = A2(a->cellsOnEdge, 2, iEdge, 2);
                const int edgeSignOnCell = (iCell == cell1) ? -1 : 1;
                if (a->interiorEdge[iEdge - 1] == 1)
                    for (int k = 1; k <= nK; k++)
                        if (K2(pOld, k, cell1) > parentTracerMinimum || K2(pOld, k, cell2) > parentTracerMinimum) {
                            const double flux_upwind = a->dvEdge[iEdge - 1] * (fmax(0.0, K2(pFlux, k, iEdge)) * K2(cOld, k, cell1) +
                                                                            fmin(0.0, K2(pFlux, k, iEdge)) * K2(cOld, k, cell2));
                            K2(cTend, k, iCell) = K2(cTend, k, iCell) + edgeSignOnCell * flux_upwind * invAreaCell1;
                            K2(cFlux, k, iEdge) = flux_upwind / a->dvEdge[iEdge - 1];
                        }
            }
        }
        /* the update (:1215-1236): over nCells, see the header */
        for (int iCell = 1; iCell <= nC; iCell++)
            for (int k = 1; k <= nK; k++)
                if (K2(pNew, k, iCell) > parentTracerMinimum) {
                    K2(cNew, k, iCell) = K2(cOld, k, iCell) * K2(pOld, k, iCell) + K2(cTend, k, iCell) * dt;
                    K2(cOld, k, iCell) = K2(cOld, k, iCell) * K2(pOld, k, iCell);
                }
    }
    /* finalize_tracers (:1989): the new time level becomes the current one; scale_tracers_back (:2141), last variable
     * first; thickness -> volume on the owned cells (:2063) */
    for (int t = nT - 1; t >= 1; t--) {
        const double *pCur = newv[a->vars[t].parent];
        for (int iCell = 1; iCell <= nCS; iCell++)
            for (int k = 1; k <= nK; k++) {
                if (K2(pCur, k, iCell) > 0.0) K2(newv[t], k, iCell) = K2(newv[t], k, iCell) / K2(pCur, k, iCell);
                else K2(newv[t], k, iCell) = 0.0;
            }
    }
    for (int t = 1; t < nT; t++)
        if (a->vars[t].volumeLike)
            for (int iCell = 1; iCell <= nCS; iCell++)
                for (int k = 1; k <= nK; k++)
                    if (K2(newv[0], k, iCell) > ICE_AREA_MINIMUM) K2(newv[t], k, iCell) = K2(newv[t], k, iCell) * K2(newv[0], k, iCell);

    for (int t = 0; t < nT; t++) {
        memcpy(a->vars[t].array, newv[t], nCell * sizeof(double));
        if (a->vars[t].edgeFluxOut) memcpy(a->vars[t].edgeFluxOut, eflux[t], nEdge * sizeof(double));
        free(newv[t]); free(tend[t]); free(eflux[t]);
    }
    if (a->edgeVelocityOut) memcpy(a->edgeVelocityOut, edgeVelocity, ((size_t)nE + 1) * sizeof(double));
    free(newv); free(tend); free(eflux); free(edgeVelocity); free(pOldNone); free(pNewNone); free(pFluxNone);
    return 0;
}
