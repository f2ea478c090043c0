/* examples/minimal_host.c -- the C ABI of include/evp_b200.h driven from plain C (what the ISO_C_BINDING shim
 * does from Fortran): one quadrilateral cell mesh of 3 x 3 cells, linear constitutive relation, one subcycle.
 *
 *   gcc -std=c99 -Iinclude examples/minimal_host.c -Lmpas-seaice_b200/csrc -levp_b200 \
 *       -Wl,-rpath,$PWD/mpas-seaice_b200/csrc -o /tmp/minimal_host && /tmp/minimal_host
 *
 * Arrays are laid out exactly as c_loc() of the MPAS pool arrays gives them: column-major, the first dimension
 * maxEdges / vertexDegree, 1-based index values, nCells+1 / nVertices+1 as the invalid neighbour.
 * Without a CUDA device evp_create fails loudly (EVP_ERR_CUDA) -- there is no CPU fallback; the program then
 * prints the library's message and exits with status 3.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "evp_b200.h"

#define NX 3
#define NC (NX * NX)
#define NVX (NX + 1)
#define NV (NVX * NVX)
#define M 4
#define D 4

static const double dc = 1000.0;

/* mesh + basis arrays of the example (also called from tests/test_examples.py, which runs the oracle on them) */
void example_fill(int *nEdgesOnCell, int *verticesOnCell, int *cellsOnVertex, int *cellVerticesAtVertex,
                  double *gU, double *gV, double *sU, double *sV, double *sM, double *tanLat, double *denom)
{
    /* connectivity of the structured quad mesh: cell (i,j) has vertices (i,j) (i+1,j) (i+1,j+1) (i,j+1), counter-clockwise */
    for (int j = 0; j < NX; j++)
        for (int i = 0; i < NX; i++) {
            const int c = j * NX + i;
            const int v[4] = {j * NVX + i, j * NVX + i + 1, (j + 1) * NVX + i + 1, (j + 1) * NVX + i};
            nEdgesOnCell[c] = 4;
            for (int k = 0; k < 4; k++) verticesOnCell[c * M + k] = v[k] + 1;
        }
    for (int k = 0; k < (NV + 1) * D; k++) { cellsOnVertex[k] = NC + 1; cellVerticesAtVertex[k] = 0; }
    for (int c = 0; c < NC; c++)
        for (int k = 0; k < 4; k++) {
            const int v = verticesOnCell[c * M + k] - 1;
            for (int s = 0; s < D; s++)
                if (cellsOnVertex[v * D + s] == NC + 1) { cellsOnVertex[v * D + s] = c + 1; cellVerticesAtVertex[v * D + s] = k + 1; break; }
        }
    /* bilinear basis on a square of side dc: gradients at the vertices and exact integrals (the host model fills
     * these in seaice_init_velocity_solver_variational; any consistent basis will do for the example) */
    const double xs[4] = {-0.5, 0.5, 0.5, -0.5}, ys[4] = {-0.5, -0.5, 0.5, 0.5};
    for (int c = 0; c < NC; c++)
        for (int jg = 0; jg < 4; jg++)           /* second index: gradient / velocity vertex */
            for (int ib = 0; ib < 4; ib++) {     /* first index: basis / stress vertex */
                const size_t q = (size_t)c * M * M + (size_t)jg * M + ib;
                gU[q] = 2.0 * xs[ib] * (0.5 + 2.0 * ys[ib] * ys[jg]) / dc;
                gV[q] = 2.0 * ys[ib] * (0.5 + 2.0 * xs[ib] * xs[jg]) / dc;
                sU[q] = dc * xs[jg] * (0.5 + ys[ib] * ys[jg] * 2.0 / 3.0) * 0.5;
                sV[q] = dc * ys[jg] * (0.5 + xs[ib] * xs[jg] * 2.0 / 3.0) * 0.5;
                sM[q] = dc * dc * (0.25 + xs[ib] * xs[jg] / 3.0) * (0.25 + ys[ib] * ys[jg] / 3.0);
            }
    for (int v = 0; v < NV; v++) { tanLat[v] = 0.0; denom[v] = dc * dc; }
}

int main(void)
{
    static int nEdgesOnCell[NC + 1], verticesOnCell[(NC + 1) * M], cellsOnVertex[(NV + 1) * D], cellVerticesAtVertex[(NV + 1) * D];
    static double gU[(NC + 1) * M * M], gV[(NC + 1) * M * M], sU[(NC + 1) * M * M], sV[(NC + 1) * M * M], sM[(NC + 1) * M * M];
    static double tanLat[NV + 1], denom[NV + 1];
    example_fill(nEdgesOnCell, verticesOnCell, cellsOnVertex, cellVerticesAtVertex, gU, gV, sU, sV, sM, tanLat, denom);

    evp_mesh_desc mesh;
    memset(&mesh, 0, sizeof mesh);
    mesh.nCells = NC; mesh.nCellsSolve = NC; mesh.nVertices = NV; mesh.nVerticesSolve = NV; mesh.maxEdges = M; mesh.vertexDegree = D;
    mesh.nEdgesOnCell = nEdgesOnCell; mesh.verticesOnCell = verticesOnCell; mesh.cellsOnVertex = cellsOnVertex;
    mesh.cellVerticesAtVertex = cellVerticesAtVertex;
    mesh.basisGradientU = gU; mesh.basisGradientV = gV; mesh.basisIntegralsU = sU; mesh.basisIntegralsV = sV;
    mesh.basisIntegralsMetric = sM; mesh.tanLatVertexRotatedOverRadius = tanLat; mesh.variationalDenominator = denom;

    evp_options opt;
    memset(&opt, 0, sizeof opt);
    opt.constitutive_relation_type = EVP_CR_LINEAR;
    opt.ocean_stress_type = EVP_OCEAN_QUADRATIC;
    opt.use_ocean_stress = 1;
    opt.device = -1;
    opt.elasticTimeStep = 30.0; opt.dynamicsTimeStep = 3600.0; opt.dampingTimescale = 1296.0;

    evp_handle *h = NULL;
    int rc = evp_create(&h, &mesh, &opt);
    if (rc != EVP_OK) {
        fprintf(stderr, "evp_create failed (%d): %s\n", rc, evp_last_error_string());
        return rc == EVP_ERR_CUDA ? 3 : 1;
    }

    static int solveStress[NC + 1], solveVelocity[NV + 1];
    static double zc[NC + 1], zcm[(NC + 1) * M], zv[NV + 1], u[NV + 1], v[NV + 1];
    for (int c = 0; c < NC; c++) solveStress[c] = 1;
    for (int i = 0; i < NV; i++) {
        const int ix = i % NVX, iy = i / NVX;
        solveVelocity[i] = (ix > 0 && ix < NX && iy > 0 && iy < NX);       /* interior vertices */
        u[i] = 1.0e-6 * dc * ix;                                             /* u = 1e-6 * x: strain11 = 1e-6 */
        v[i] = 0.0;
    }
    evp_step_fields f;
    memset(&f, 0, sizeof f);
    f.solveStress = solveStress; f.solveVelocity = solveVelocity; f.icePressure = zc; f.uVelocity = u; f.vVelocity = v;
    f.stress11 = zcm; f.stress22 = zcm; f.stress12 = zcm;
    f.totalMassVertex = zv; f.totalMassVertexfVertex = zv; f.iceAreaVertex = zv; f.airStressVertexU = zv; f.airStressVertexV = zv;
    f.surfaceTiltForceU = zv; f.surfaceTiltForceV = zv; f.oceanStressU = zv; f.oceanStressV = zv;
    f.uOceanVelocityVertex = zv; f.vOceanVelocityVertex = zv;
    if ((rc = evp_update_step(h, &f)) || (rc = evp_run_subcycles(h, 1))) {
        fprintf(stderr, "subcycle failed (%d): %s\n", rc, evp_last_error_string());
        evp_destroy(h);
        return 1;
    }
    static double e11[(NC + 1) * M];
    evp_out_fields o;
    memset(&o, 0, sizeof o);
    o.strain11 = e11;
    rc = evp_fetch(h, &o);
    float ms = 0.f;
    evp_last_run_ms(h, &ms);
    printf("strain11 at cell 5, vertex 1: %.6e (expected 1.000000e-06), %.3f ms\n", e11[4 * M + 0], ms);
    evp_destroy(h);
    return rc == EVP_OK ? 0 : 1;
}
