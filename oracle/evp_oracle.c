/*
 * oracle/evp_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, FP64, no FMA contraction: build with -ffp-contract=off) of the
 * MPAS-Seaice EVP momentum path, loop-for-loop and in the reference's floating-point operation
 * order.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (libevp_b200.so) never does.
 *
 * PARITY PINNING: the reference (Fortran + un-vendored MPAS framework) cannot be compiled in this
 * image (no Fortran compiler, no MPI, no netCDF) and ships NO stored numeric outputs
 * (SURVEY.md section 8c).  This oracle is pinned by OUTPUTS OF THE REFERENCE'S OWN SOURCE EXECUTED HERE: an
 * interpreter for the Fortran subset of these routines (tests/golden/fortran_subset.py) runs
 * subcycle_velocity_solver, velocity_solver_pre_subcycle and velocity_solver_post_subcycle with everything below them
 * from the files under /root/reference, one IEEE operation per operator in the written order; the fixtures
 * (tests/golden/refexec_*.npz, tests/golden/step/refexec_step_*.npz) are reproduced by this file bit for bit
 * (tests/test_golden.py, tests/test_refexec_step.py).  Also: the reference's analytic known answers
 * (testing_and_setup/testcases/square/operators_strain_stress_divergence/create_ics.py:12-48,
 * src/shared/mpas_seaice_testing.F:726-839; tests/test_oracle_kat.py), analytic fields generated with the reference's
 * own Python test-case scripts (tests/test_analytic_golden.py) and closed-form recurrences (tests/test_closed_forms.py).
 * Not pinned: what a production Fortran compiler does beyond the source semantics (FMA contraction, vectorised sums).
 *
 * Array conventions are the reference's: Fortran column-major, 1-based index VALUES, one junk
 * element at the end of every mesh array.  A Fortran A(i,j,c) with leading dims (M,M) is
 * A[(i-1) + M*((j-1) + M*(c-1))].
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define IDX2(i, c, M) ((size_t)((i) - 1) + (size_t)(M) * (size_t)((c) - 1))
#define IDX3(i, j, c, M) ((size_t)((i) - 1) + (size_t)(M) * ((size_t)((j) - 1) + (size_t)(M) * (size_t)((c) - 1)))

/* src/shared/mpas_seaice_velocity_solver_constitutive_relation.F:41-59 */
static const double eccentricitySquared = 2.0 * 2.0;
static const double puny = 1.0e-11;
static const double dampingRatioDenominator = 0.86;
static const double dampingRatio = 5.5e-3;
/* src/shared/mpas_seaice_velocity_solver.F:61-65 */
static const double sinOceanTurningAngle = 0.0;
static const double cosOceanTurningAngle = 1.0;
static const double seaiceAreaMinimum = 0.001;
static const double seaiceMassMinimum = 0.01;
/* src/shared/mpas_seaice_constants.F:43-92 with src/column/constants/cice/ice_constants_colpkg.F90:22-63 */
static const double seaiceDensitySeaWater = 1026.0;
static const double seaiceIceOceanDragCoefficient = 0.00536;
static const double seaiceDensityIce = 917.0;
static const double seaiceDensitySnow = 330.0;
static const double seaiceIceStrengthConstantHiblerP = 2.75e4;
static const double seaiceIceStrengthConstantHiblerC = 20.0;
static const double seaicePuny = 1.0e-11;

enum { EVP = 1, EVP_REVISED = 2, LINEAR = 3, NONE = 4 };   /* constitutive_relation.F:34-38 */
enum { QUADRATIC_OCEAN_STRESS = 1, LINEAR_OCEAN_STRESS = 2 }; /* velocity_solver.F:55-57 */

/* ------------------------------------------------------------------------------------------
 * seaice_strain_tensor_variational  (src/shared/mpas_seaice_velocity_solver_variational.F:575-670)
 * ------------------------------------------------------------------------------------------ */
void orc_strain_tensor_variational(int nCells, int maxEdges, const int *nEdgesOnCell,
                                   const int *verticesOnCell, const int *solveStress,
                                   const double *uVelocity, const double *vVelocity,
                                   const double *basisGradientU, const double *basisGradientV,
                                   const double *tanLatVertexRotatedOverRadius,
                                   double *strain11, double *strain22, double *strain12)
{
    const int M = maxEdges;
#pragma omp parallel for schedule(static)
    for (int iCell = 1; iCell <= nCells; iCell++) {
        if (solveStress[iCell - 1] == 1) {
            const int n = nEdgesOnCell[iCell - 1];
            for (int iGradientVertex = 1; iGradientVertex <= n; iGradientVertex++) {
                double strain11Tmp = 0.0, strain22Tmp = 0.0, strain12Tmp = 0.0;
                for (int iBasisVertex = 1; iBasisVertex <= n; iBasisVertex++) {
                    const int iVertex = verticesOnCell[IDX2(iBasisVertex, iCell, M)];
                    const double gu = basisGradientU[IDX3(iBasisVertex, iGradientVertex, iCell, M)];
                    const double gv = basisGradientV[IDX3(iBasisVertex, iGradientVertex, iCell, M)];
                    const double u = uVelocity[iVertex - 1], v = vVelocity[iVertex - 1];
                    strain11Tmp = strain11Tmp + u * gu;
                    strain22Tmp = strain22Tmp + v * gv;
                    strain12Tmp = strain12Tmp + 0.5 * (u * gv + v * gu);
                }
                const int jVertex = verticesOnCell[IDX2(iGradientVertex, iCell, M)];
                const double t = tanLatVertexRotatedOverRadius[jVertex - 1];
                strain11[IDX2(iGradientVertex, iCell, M)] = strain11Tmp - vVelocity[jVertex - 1] * t;
                strain12[IDX2(iGradientVertex, iCell, M)] = strain12Tmp + uVelocity[jVertex - 1] * t * 0.5;
                strain22[IDX2(iGradientVertex, iCell, M)] = strain22Tmp;
            }
        }
    }
}

/* seaice_average_strains_on_vertex (variational.F:684-763); serial in the reference */
void orc_average_strains_on_vertex(int nCells, int nVerticesSolve, int vertexDegree, int maxEdges,
                                   const int *cellsOnVertex, const int *cellVerticesAtVertex,
                                   const double *areaCell,
                                   double *strain11, double *strain22, double *strain12)
{
    const int M = maxEdges, D = vertexDegree;
    for (int iVertex = 1; iVertex <= nVerticesSolve; iVertex++) {
        double s11 = 0.0, s22 = 0.0, s12 = 0.0, denominator = 0.0;
        for (int k = 1; k <= D; k++) {
            const int iCell = cellsOnVertex[IDX2(k, iVertex, D)];
            if (iCell <= nCells) {
                const int j = cellVerticesAtVertex[IDX2(k, iVertex, D)];
                s11 = s11 + strain11[IDX2(j, iCell, M)] * areaCell[iCell - 1];
                s22 = s22 + strain22[IDX2(j, iCell, M)] * areaCell[iCell - 1];
                s12 = s12 + strain12[IDX2(j, iCell, M)] * areaCell[iCell - 1];
                denominator = denominator + areaCell[iCell - 1];
            }
        }
        s11 = s11 / denominator;
        s22 = s22 / denominator;
        s12 = s12 / denominator;
        for (int k = 1; k <= D; k++) {
            const int iCell = cellsOnVertex[IDX2(k, iVertex, D)];
            if (iCell <= nCells) {
                const int j = cellVerticesAtVertex[IDX2(k, iVertex, D)];
                strain11[IDX2(j, iCell, M)] = s11;
                strain22[IDX2(j, iCell, M)] = s22;
                strain12[IDX2(j, iCell, M)] = s12;
            }
        }
    }
}

/* seaice_evp_constitutive_relation (constitutive_relation.F:178-248) */
static inline void evp_constitutive_relation(double *stress11, double *stress22, double *stress12,
                                             double strain11, double strain22, double strain12,
                                             double icePressure, double *replacementPressure,
                                             double dtElastic, double dampingTimescale)
{
    const double strainDivergence = strain11 + strain22;
    const double strainTension = strain11 - strain22;
    const double strainShearing = strain12 * 2.0;
    double stress1 = *stress11 + *stress22;
    double stress2 = *stress11 - *stress22;
    const double Delta = sqrt(strainDivergence * strainDivergence +
                              (strainTension * strainTension + strainShearing * strainShearing) / eccentricitySquared);
    double pressureCoefficient = icePressure / fmax(Delta, puny);
    *replacementPressure = pressureCoefficient * Delta;
    pressureCoefficient = (pressureCoefficient * dtElastic) / (2.0 * dampingTimescale);
    const double denominator = 1.0 + (0.5 * dtElastic) / dampingTimescale;
    stress1 = (stress1 + pressureCoefficient * (strainDivergence - Delta)) / denominator;
    stress2 = (stress2 + (pressureCoefficient / eccentricitySquared) * strainTension) / denominator;
    *stress12 = (*stress12 + (pressureCoefficient / eccentricitySquared) * strainShearing * 0.5) / denominator;
    *stress11 = 0.5 * (stress1 + stress2);
    *stress22 = 0.5 * (stress1 - stress2);
}

/* seaice_evp_constitutive_relation_revised (constitutive_relation.F:262-330) */
static inline void evp_constitutive_relation_revised(double *stress11, double *stress22, double *stress12,
                                                     double strain11, double strain22, double strain12,
                                                     double icePressure, double *replacementPressure)
{
    const double strainDivergence = strain11 + strain22;
    const double strainTension = strain11 - strain22;
    const double strainShearing = strain12 * 2.0;
    double stress1 = *stress11 + *stress22;
    double stress2 = *stress11 - *stress22;
    const double Delta = sqrt(strainDivergence * strainDivergence +
                              (strainTension * strainTension + strainShearing * strainShearing) / eccentricitySquared);
    double pressureCoefficient = icePressure / fmax(Delta, puny);
    *replacementPressure = pressureCoefficient * Delta;
    pressureCoefficient = (pressureCoefficient * 2.0 * dampingRatio) / dampingRatioDenominator;
    const double denominator = 1.0 + (2.0 * dampingRatio) / dampingRatioDenominator;
    stress1 = (stress1 + pressureCoefficient * (strainDivergence - Delta)) / denominator;
    stress2 = (stress2 + (pressureCoefficient / eccentricitySquared) * strainTension) / denominator;
    *stress12 = (*stress12 + (pressureCoefficient / eccentricitySquared) * strainShearing * 0.5) / denominator;
    *stress11 = 0.5 * (stress1 + stress2);
    *stress22 = 0.5 * (stress1 - stress2);
}

/* seaice_stress_tensor_variational (variational.F:777-975).  Note the EVP branch zeroes
 * replacementPressure(:,iCell) for EVERY cell (:862); the revised and linear branches do not. */
void orc_stress_tensor_variational(int nCells, int maxEdges, const int *nEdgesOnCell,
                                   const int *solveStress, int constitutiveRelationType,
                                   double dtElastic, double dampingTimescale,
                                   const double *icePressure,
                                   const double *strain11, const double *strain22, const double *strain12,
                                   double *stress11, double *stress22, double *stress12,
                                   double *replacementPressure)
{
    const int M = maxEdges;
    if (constitutiveRelationType == EVP) {
#pragma omp parallel for schedule(static)
        for (int iCell = 1; iCell <= nCells; iCell++) {
            for (int k = 1; k <= M; k++) replacementPressure[IDX2(k, iCell, M)] = 0.0;
            if (solveStress[iCell - 1] == 1) {
                for (int j = 1; j <= nEdgesOnCell[iCell - 1]; j++) {
                    const size_t q = IDX2(j, iCell, M);
                    evp_constitutive_relation(&stress11[q], &stress22[q], &stress12[q],
                                              strain11[q], strain22[q], strain12[q],
                                              icePressure[iCell - 1], &replacementPressure[q],
                                              dtElastic, dampingTimescale);
                }
            }
        }
    } else if (constitutiveRelationType == EVP_REVISED) {
#pragma omp parallel for schedule(static)
        for (int iCell = 1; iCell <= nCells; iCell++) {
            if (solveStress[iCell - 1] == 1) {
                for (int j = 1; j <= nEdgesOnCell[iCell - 1]; j++) {
                    const size_t q = IDX2(j, iCell, M);
                    evp_constitutive_relation_revised(&stress11[q], &stress22[q], &stress12[q],
                                                      strain11[q], strain22[q], strain12[q],
                                                      icePressure[iCell - 1], &replacementPressure[q]);
                }
            }
        }
    } else if (constitutiveRelationType == LINEAR) {
        /* seaice_linear_constitutive_relation (constitutive_relation.F:344-373), lambda = 1 */
#pragma omp parallel for schedule(static)
        for (int iCell = 1; iCell <= nCells; iCell++) {
            if (solveStress[iCell - 1] == 1) {
                for (int j = 1; j <= nEdgesOnCell[iCell - 1]; j++) {
                    const size_t q = IDX2(j, iCell, M);
                    stress11[q] = 1.0 * strain11[q];
                    stress22[q] = 1.0 * strain22[q];
                    stress12[q] = 1.0 * strain12[q];
                }
            }
        }
    }
}

/* seaice_stress_divergence_variational (variational.F:1064-1184) */
void orc_stress_divergence_variational(int nVerticesSolve, int vertexDegree, int maxEdges,
                                       const int *nEdgesOnCell, const int *cellsOnVertex,
                                       const int *cellVerticesAtVertex, const int *solveVelocity,
                                       const double *stress11, const double *stress22, const double *stress12,
                                       const double *basisIntegralsU, const double *basisIntegralsV,
                                       const double *basisIntegralsMetric,
                                       const double *variationalDenominator,
                                       const double *tanLatVertexRotatedOverRadius,
                                       double *stressDivergenceU, double *stressDivergenceV)
{
    const int M = maxEdges, D = vertexDegree;
#pragma omp parallel for schedule(static)
    for (int iVertex = 1; iVertex <= nVerticesSolve; iVertex++) {
        if (solveVelocity[iVertex - 1] == 1) {
            double sdU = 0.0, sdV = 0.0;
            const double t = tanLatVertexRotatedOverRadius[iVertex - 1];
            for (int iSurroundingCell = 1; iSurroundingCell <= D; iSurroundingCell++) {
                const int iCell = cellsOnVertex[IDX2(iSurroundingCell, iVertex, D)];
                const int iVelocityVertex = cellVerticesAtVertex[IDX2(iSurroundingCell, iVertex, D)];
                double cU = 0.0, cV = 0.0;
                for (int iStressVertex = 1; iStressVertex <= nEdgesOnCell[iCell - 1]; iStressVertex++) {
                    const size_t q = IDX2(iStressVertex, iCell, M);
                    const size_t b = IDX3(iStressVertex, iVelocityVertex, iCell, M);
                    cU = cU - stress11[q] * basisIntegralsU[b] - stress12[q] * basisIntegralsV[b] -
                         stress12[q] * basisIntegralsMetric[b] * t;
                    cV = cV - stress22[q] * basisIntegralsV[b] - stress12[q] * basisIntegralsU[b] +
                         stress11[q] * basisIntegralsMetric[b] * t;
                }
                sdU = sdU + cU;
                sdV = sdV + cV;
            }
            stressDivergenceU[iVertex - 1] = sdU / variationalDenominator[iVertex - 1];
            stressDivergenceV[iVertex - 1] = sdV / variationalDenominator[iVertex - 1];
        }
    }
}

/* ocean_stress_coefficient (src/shared/mpas_seaice_velocity_solver.F:2986-3082) */
void orc_ocean_stress_coefficient(int nVerticesSolve, int nVertices, int useOceanStress, int oceanStressType,
                                  const int *solveVelocity, const double *iceAreaVertex,
                                  const double *uOceanVelocityVertex, const double *vOceanVelocityVertex,
                                  const double *uVelocity, const double *vVelocity,
                                  double *oceanStressCoeff)
{
    if (useOceanStress) {
        if (oceanStressType == QUADRATIC_OCEAN_STRESS) {
#pragma omp parallel for schedule(static)
            for (int i = 0; i < nVerticesSolve; i++) {
                if (solveVelocity[i] == 1) {
                    const double du = uOceanVelocityVertex[i] - uVelocity[i];
                    const double dv = vOceanVelocityVertex[i] - vVelocity[i];
                    oceanStressCoeff[i] = seaiceIceOceanDragCoefficient * seaiceDensitySeaWater * iceAreaVertex[i] *
                                          sqrt(du * du + dv * dv);
                }
            }
        } else if (oceanStressType == LINEAR_OCEAN_STRESS) {
#pragma omp parallel for schedule(static)
            for (int i = 0; i < nVerticesSolve; i++) {
                if (solveVelocity[i] == 1)
                    oceanStressCoeff[i] = seaiceIceOceanDragCoefficient * seaiceDensitySeaWater * iceAreaVertex[i];
            }
        }
    } else {
        for (int i = 0; i < nVertices + 1; i++) oceanStressCoeff[i] = 0.0; /* whole array (:3075) */
    }
}

/* solve_velocity (velocity_solver.F:3096-3208) */
void orc_solve_velocity(int nVerticesSolve, const int *solveVelocity, double elasticTimeStep,
                        const double *totalMassVertex, const double *totalMassVertexfVertex,
                        const double *stressDivergenceU, const double *stressDivergenceV,
                        const double *airStressVertexU, const double *airStressVertexV,
                        const double *surfaceTiltForceU, const double *surfaceTiltForceV,
                        const double *oceanStressU, const double *oceanStressV,
                        const double *oceanStressCoeff, double *uVelocity, double *vVelocity)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nVerticesSolve; i++) {
        if (solveVelocity[i] == 1) {
            const double sgn = copysign(1.0, totalMassVertexfVertex[i]);
            const double l11 = totalMassVertex[i] / elasticTimeStep + oceanStressCoeff[i] * cosOceanTurningAngle;
            const double l12 = -totalMassVertexfVertex[i] - oceanStressCoeff[i] * sinOceanTurningAngle * sgn;
            const double l21 = totalMassVertexfVertex[i] + oceanStressCoeff[i] * sinOceanTurningAngle * sgn;
            const double l22 = totalMassVertex[i] / elasticTimeStep + oceanStressCoeff[i] * cosOceanTurningAngle;
            const double r1 = stressDivergenceU[i] + airStressVertexU[i] + surfaceTiltForceU[i] +
                              oceanStressCoeff[i] * oceanStressU[i] +
                              (totalMassVertex[i] * uVelocity[i]) / elasticTimeStep;
            const double r2 = stressDivergenceV[i] + airStressVertexV[i] + surfaceTiltForceV[i] +
                              oceanStressCoeff[i] * oceanStressV[i] +
                              (totalMassVertex[i] * vVelocity[i]) / elasticTimeStep;
            const double solutionDenominator = l11 * l22 - l12 * l21;
            uVelocity[i] = (l22 * r1 - l12 * r2) / solutionDenominator;
            vVelocity[i] = (l11 * r2 - l21 * r1) / solutionDenominator;
        }
    }
}

/* solve_velocity_revised (velocity_solver.F:3222-3342) */
void orc_solve_velocity_revised(int nVerticesSolve, const int *solveVelocity, double dynamicsTimeStep,
                                double numericalInertiaCoefficient,
                                const double *totalMassVertex, const double *totalMassVertexfVertex,
                                const double *stressDivergenceU, const double *stressDivergenceV,
                                const double *airStressVertexU, const double *airStressVertexV,
                                const double *surfaceTiltForceU, const double *surfaceTiltForceV,
                                const double *oceanStressU, const double *oceanStressV,
                                const double *oceanStressCoeff,
                                const double *uVelocityInitial, const double *vVelocityInitial,
                                double *uVelocity, double *vVelocity)
{
    const double beta = numericalInertiaCoefficient;
    for (int i = 0; i < nVerticesSolve; i++) {
        if (solveVelocity[i] == 1) {
            const double sgn = copysign(1.0, totalMassVertexfVertex[i]);
            const double l11 = (beta + 1.0) * (totalMassVertex[i] / dynamicsTimeStep) + oceanStressCoeff[i] * cosOceanTurningAngle;
            const double l12 = -totalMassVertexfVertex[i] - oceanStressCoeff[i] * sinOceanTurningAngle * sgn;
            const double l21 = totalMassVertexfVertex[i] + oceanStressCoeff[i] * sinOceanTurningAngle * sgn;
            const double l22 = (beta + 1.0) * (totalMassVertex[i] / dynamicsTimeStep) + oceanStressCoeff[i] * cosOceanTurningAngle;
            const double r1 = stressDivergenceU[i] + airStressVertexU[i] + surfaceTiltForceU[i] +
                              oceanStressCoeff[i] * oceanStressU[i] +
                              (totalMassVertex[i] * (beta * uVelocity[i] + uVelocityInitial[i])) / dynamicsTimeStep;
            const double r2 = stressDivergenceV[i] + airStressVertexV[i] + surfaceTiltForceV[i] +
                              oceanStressCoeff[i] * oceanStressV[i] +
                              (totalMassVertex[i] * (beta * vVelocity[i] + vVelocityInitial[i])) / dynamicsTimeStep;
            const double solutionDenominator = l11 * l22 - l12 * l21;
            uVelocity[i] = (l22 * r1 - l12 * r2) / solutionDenominator;
            vVelocity[i] = (l11 * r2 - l21 * r1) / solutionDenominator;
        }
    }
}

/* seaice_set_special_boundaries_velocity (src/shared/mpas_seaice_special_boundaries.F:253-331);
 * sequential, in place, in vertex order -- exactly as the reference. */
void orc_set_special_boundaries_velocity(int nVertices, const int *vertexBoundaryType,
                                         const int *vertexBoundarySourceLocal,
                                         double *uVelocity, double *vVelocity)
{
    for (int i = 0; i < nVertices; i++) {
        if (vertexBoundaryType[i] == 1) {
            const int s = vertexBoundarySourceLocal[i] - 1;
            uVelocity[i] = uVelocity[s];
            vVelocity[i] = vVelocity[s];
        } else if (vertexBoundaryType[i] == 2) {
            const int s = vertexBoundarySourceLocal[i] - 1;
            uVelocity[i] = -uVelocity[s];
            vVelocity[i] = -vVelocity[s];
        } else if (vertexBoundaryType[i] == 3) {
            uVelocity[i] = 0.0;
            vVelocity[i] = 0.0;
        }
    }
}

/* ==========================================================================================
 * weak operators (src/shared/mpas_seaice_velocity_solver_weak.F): one stress point per cell
 * ========================================================================================== */
#define NV2(d, k, c, M) ((size_t)((d) - 1) + 2 * ((size_t)((k) - 1) + (size_t)(M) * (size_t)((c) - 1)))

/* seaice_strain_tensor_weak (weak.F:112-253) */
void orc_strain_tensor_weak(int nCells, int maxEdges, const int *nEdgesOnCell, const int *verticesOnCell,
                            const int *edgesOnCell, const int *verticesOnEdge, const double *dvEdge,
                            const double *areaCell, double sphere_radius, const int *solveStress,
                            const double *uVelocity, const double *vVelocity,
                            const double *normalVectorPolygon, const double *latCellRotated,
                            double *strain11, double *strain22, double *strain12)
{
    const int M = maxEdges;
    double sphereRadius = sphere_radius;
    if (sphereRadius == 0.0) sphereRadius = 1.0;
    for (int iCell = 1; iCell <= nCells; iCell++) {
        strain11[iCell - 1] = 0.0;
        strain22[iCell - 1] = 0.0;
        strain12[iCell - 1] = 0.0;
        if (solveStress[iCell - 1] == 1) {
            double uCellCentre = 0.0, vCellCentre = 0.0;
            for (int k = 1; k <= nEdgesOnCell[iCell - 1]; k++) {
                int iVertex = verticesOnCell[IDX2(k, iCell, M)];
                uCellCentre = uCellCentre + uVelocity[iVertex - 1];
                vCellCentre = vCellCentre + vVelocity[iVertex - 1];
                const int iEdge = edgesOnCell[IDX2(k, iCell, M)];
                double uVelocityEdge = 0.0, vVelocityEdge = 0.0;
                for (int e = 1; e <= 2; e++) {
                    iVertex = verticesOnEdge[IDX2(e, iEdge, 2)];
                    uVelocityEdge = uVelocityEdge + uVelocity[iVertex - 1];
                    vVelocityEdge = vVelocityEdge + vVelocity[iVertex - 1];
                }
                uVelocityEdge = uVelocityEdge / 2.0;
                vVelocityEdge = vVelocityEdge / 2.0;
                const double nx = normalVectorPolygon[NV2(1, k, iCell, M)], ny = normalVectorPolygon[NV2(2, k, iCell, M)];
                strain11[iCell - 1] = strain11[iCell - 1] + uVelocityEdge * nx * dvEdge[iEdge - 1];
                strain22[iCell - 1] = strain22[iCell - 1] + vVelocityEdge * ny * dvEdge[iEdge - 1];
                strain12[iCell - 1] = strain12[iCell - 1] + 0.5 * (uVelocityEdge * ny + vVelocityEdge * nx) * dvEdge[iEdge - 1];
            }
            uCellCentre = uCellCentre / (double)nEdgesOnCell[iCell - 1];
            vCellCentre = vCellCentre / (double)nEdgesOnCell[iCell - 1];
            strain11[iCell - 1] = strain11[iCell - 1] / areaCell[iCell - 1];
            strain22[iCell - 1] = strain22[iCell - 1] / areaCell[iCell - 1];
            strain12[iCell - 1] = strain12[iCell - 1] / areaCell[iCell - 1];
            strain11[iCell - 1] = strain11[iCell - 1] - (vCellCentre * tan(latCellRotated[iCell - 1])) / sphereRadius;
            strain12[iCell - 1] = strain12[iCell - 1] + (uCellCentre * tan(latCellRotated[iCell - 1]) * 0.5) / sphereRadius;
        }
    }
}

/* seaice_stress_tensor_weak (weak.F:267-385): unlike the variational routine it zeroes the stress of
 * cells that are not solved, and it never touches replacementPressure of those cells */
void orc_stress_tensor_weak(int nCells, const int *solveStress, int constitutiveRelationType,
                            double dtElastic, double dampingTimescale, const double *icePressure,
                            const double *strain11, const double *strain22, const double *strain12,
                            double *stress11, double *stress22, double *stress12, double *replacementPressure)
{
    if (constitutiveRelationType != EVP && constitutiveRelationType != EVP_REVISED && constitutiveRelationType != LINEAR)
        return;
    for (int i = 0; i < nCells; i++) {
        if (solveStress[i] == 1) {
            if (constitutiveRelationType == EVP)
                evp_constitutive_relation(&stress11[i], &stress22[i], &stress12[i], strain11[i], strain22[i], strain12[i],
                                          icePressure[i], &replacementPressure[i], dtElastic, dampingTimescale);
            else if (constitutiveRelationType == EVP_REVISED)
                evp_constitutive_relation_revised(&stress11[i], &stress22[i], &stress12[i], strain11[i], strain22[i],
                                                  strain12[i], icePressure[i], &replacementPressure[i]);
            else {
                stress11[i] = 1.0 * strain11[i];
                stress22[i] = 1.0 * strain22[i];
                stress12[i] = 1.0 * strain12[i];
            }
        } else {
            stress11[i] = 0.0;
            stress22[i] = 0.0;
            stress12[i] = 0.0;
        }
    }
}

/* seaice_stress_divergence_weak (weak.F:493-640) */
void orc_stress_divergence_weak(int nVerticesSolve, int vertexDegree, const int *cellsOnVertex, const int *edgesOnVertex,
                                const int *cellsOnEdge, const double *dcEdge, const double *areaTriangle,
                                double sphere_radius, const int *solveVelocity,
                                const double *stress11, const double *stress22, const double *stress12,
                                const double *normalVectorTriangle, const double *latVertexRotated,
                                double *stressDivergenceU, double *stressDivergenceV)
{
    const int D = vertexDegree;
    double sphereRadius = sphere_radius;
    if (sphereRadius == 0.0) sphereRadius = 1.0;
    for (int iVertex = 1; iVertex <= nVerticesSolve; iVertex++) {
        stressDivergenceU[iVertex - 1] = 0.0;
        stressDivergenceV[iVertex - 1] = 0.0;
        if (solveVelocity[iVertex - 1] == 1) {
            double stress11Vertex = 0.0, stress22Vertex = 0.0, stress12Vertex = 0.0;
            for (int k = 1; k <= D; k++) {
                int iCell = cellsOnVertex[IDX2(k, iVertex, D)];
                stress11Vertex = stress11Vertex + stress11[iCell - 1];
                stress22Vertex = stress22Vertex + stress22[iCell - 1];
                stress12Vertex = stress12Vertex + stress12[iCell - 1];
                const int iEdge = edgesOnVertex[IDX2(k, iVertex, D)];
                double stress11Edge = 0.0, stress22Edge = 0.0, stress12Edge = 0.0;
                for (int e = 1; e <= 2; e++) {
                    iCell = cellsOnEdge[IDX2(e, iEdge, 2)];
                    stress11Edge = stress11Edge + stress11[iCell - 1];
                    stress22Edge = stress22Edge + stress22[iCell - 1];
                    stress12Edge = stress12Edge + stress12[iCell - 1];
                }
                stress11Edge = stress11Edge / 2.0;
                stress22Edge = stress22Edge / 2.0;
                stress12Edge = stress12Edge / 2.0;
                const double nx = normalVectorTriangle[NV2(1, k, iVertex, D)], ny = normalVectorTriangle[NV2(2, k, iVertex, D)];
                stressDivergenceU[iVertex - 1] = stressDivergenceU[iVertex - 1] +
                    (stress11Edge * nx + stress12Edge * ny) * dcEdge[iEdge - 1];
                stressDivergenceV[iVertex - 1] = stressDivergenceV[iVertex - 1] +
                    (stress22Edge * ny + stress12Edge * nx) * dcEdge[iEdge - 1];
            }
            stress11Vertex = stress11Vertex / (double)D;
            stress22Vertex = stress22Vertex / (double)D;
            stress12Vertex = stress12Vertex / (double)D;
            stressDivergenceU[iVertex - 1] = stressDivergenceU[iVertex - 1] / areaTriangle[iVertex - 1];
            stressDivergenceV[iVertex - 1] = stressDivergenceV[iVertex - 1] / areaTriangle[iVertex - 1];
            stressDivergenceU[iVertex - 1] = stressDivergenceU[iVertex - 1] -
                (tan(latVertexRotated[iVertex - 1]) * stress12Vertex * 2.0) / sphereRadius;
            stressDivergenceV[iVertex - 1] = stressDivergenceV[iVertex - 1] +
                (tan(latVertexRotated[iVertex - 1]) * (stress11Vertex - stress22Vertex)) / sphereRadius;
        }
    }
}

/* interpolate_strains_weak_to_variational (velocity_solver.F:2877-2972); strainXXVertex are work arrays
 * (nVertices+1) that keep whatever they held at the halo vertices, as in the reference */
void orc_interpolate_strains_weak_to_variational(int nCells, int nVerticesSolve, int vertexDegree, int maxEdges,
                                                 const int *nEdgesOnCell, const int *verticesOnCell,
                                                 const int *cellsOnVertex, const double *areaCell,
                                                 const double *strain11weak, const double *strain22weak,
                                                 const double *strain12weak,
                                                 double *strain11Vertex, double *strain22Vertex, double *strain12Vertex,
                                                 double *strain11var, double *strain22var, double *strain12var)
{
    const int M = maxEdges, D = vertexDegree;
    for (int iVertex = 1; iVertex <= nVerticesSolve; iVertex++) {
        double denom = 0.0;
        strain11Vertex[iVertex - 1] = 0.0;
        strain22Vertex[iVertex - 1] = 0.0;
        strain12Vertex[iVertex - 1] = 0.0;
        for (int k = 1; k <= D; k++) {
            const int iCell = cellsOnVertex[IDX2(k, iVertex, D)];
            if (iCell >= 1 && iCell <= nCells) {
                strain11Vertex[iVertex - 1] = strain11Vertex[iVertex - 1] + areaCell[iCell - 1] * strain11weak[iCell - 1];
                strain22Vertex[iVertex - 1] = strain22Vertex[iVertex - 1] + areaCell[iCell - 1] * strain22weak[iCell - 1];
                strain12Vertex[iVertex - 1] = strain12Vertex[iVertex - 1] + areaCell[iCell - 1] * strain12weak[iCell - 1];
                denom = denom + areaCell[iCell - 1];
            }
        }
        strain11Vertex[iVertex - 1] = strain11Vertex[iVertex - 1] / denom;
        strain22Vertex[iVertex - 1] = strain22Vertex[iVertex - 1] / denom;
        strain12Vertex[iVertex - 1] = strain12Vertex[iVertex - 1] / denom;
    }
    for (int iCell = 1; iCell <= nCells; iCell++) {
        for (int k = 1; k <= nEdgesOnCell[iCell - 1]; k++) {
            const int iVertex = verticesOnCell[IDX2(k, iCell, M)];
            strain11var[IDX2(k, iCell, M)] = strain11Vertex[iVertex - 1];
            strain22var[IDX2(k, iCell, M)] = strain22Vertex[iVertex - 1];
            strain12var[IDX2(k, iCell, M)] = strain12Vertex[iVertex - 1];
        }
    }
}

/* seaice_final_divergence_shear_weak (weak.F:651-751).  NOTE the reference assigns the WHOLE work array
 * ("Delta = sqrt(...)", weak.F:729) inside the cell loop, so after the loop every Delta(i) holds the value of
 * the last owned cell; ridgeShear is computed from that.  Restated as written. */
void orc_final_divergence_shear_weak(int nCellsSolve, const double *strain11, const double *strain22,
                                     const double *strain12, double *divergence, double *shear,
                                     double *ridgeConvergence, double *ridgeShear)
{
    double Delta = 0.0;
    for (int i = 0; i < nCellsSolve; i++) {
        const double strainDivergence = strain11[i] + strain22[i];
        const double strainTension = strain11[i] - strain22[i];
        const double strainShearing = strain12[i] * 2.0;
        Delta = sqrt(strainDivergence * strainDivergence +
                     (strainTension * strainTension + strainShearing * strainShearing) / eccentricitySquared);
        divergence[i] = strainDivergence;
        shear[i] = sqrt(strainTension * strainTension + strainShearing * strainShearing);
    }
    if (ridgeConvergence) {
        for (int i = 0; i < nCellsSolve; i++) {
            ridgeConvergence[i] = -fmin(divergence[i], 0.0);
            ridgeShear[i] = 0.5 * (Delta - fabs(divergence[i]));
        }
    }
}

/* The argument block of one dynamics step of the subcycle; mirrors module seaice_mesh_pool
 * (src/shared/mpas_seaice_mesh_pool.F:23-57) plus the host-side vertex fields of a8/a9. */
typedef struct {
    int nCells, nVertices, nVerticesSolve, maxEdges, vertexDegree;
    const int *nEdgesOnCell, *verticesOnCell, *cellsOnVertex, *cellVerticesAtVertex;
    const double *basisGradientU, *basisGradientV, *basisIntegralsU, *basisIntegralsV, *basisIntegralsMetric;
    const double *tanLatVertexRotatedOverRadius, *variationalDenominator, *areaCell;
    int constitutiveRelationType, oceanStressType, useOceanStress, averageVariationalStrains;
    int useSpecialBoundariesVelocity, useSpecialBoundariesVelocityMasks;
    double elasticTimeStep, dynamicsTimeStep, dampingTimescale, numericalInertiaCoefficient;
    int *solveStress, *solveVelocity;
    const int *vertexBoundaryType, *vertexBoundarySourceLocal;
    const int *solveStressSpecialBoundaries, *solveVelocitySpecialBoundaries;
    const double *icePressure, *totalMassVertex, *totalMassVertexfVertex, *iceAreaVertex;
    const double *airStressVertexU, *airStressVertexV, *surfaceTiltForceU, *surfaceTiltForceV;
    const double *oceanStressU, *oceanStressV, *uOceanVelocityVertex, *vOceanVelocityVertex;
    const double *uVelocityInitial, *vVelocityInitial;
    double *uVelocity, *vVelocity, *stress11, *stress22, *stress12;
    double *strain11, *strain22, *strain12, *replacementPressure;
    double *stressDivergenceU, *stressDivergenceV, *oceanStressCoeff;
    /* weak operators (config_strain_scheme / config_stress_divergence_scheme = 'weak'); 1 = variational, 2 = weak */
    int strainScheme, stressDivergenceScheme;
    double sphere_radius;
    const int *edgesOnCell, *verticesOnEdge, *edgesOnVertex, *cellsOnEdge;
    const double *dvEdge, *dcEdge, *areaTriangle, *normalVectorPolygon, *normalVectorTriangle;
    const double *latCellRotated, *latVertexRotated;
    double *stress11Weak, *stress22Weak, *stress12Weak, *strain11Weak, *strain22Weak, *strain12Weak;
    double *replacementPressureWeak, *strain11Vertex, *strain22Vertex, *strain12Vertex;
} orc_subcycle_args;

static void special_boundaries(const orc_subcycle_args *a)
{
    if (a->useSpecialBoundariesVelocity)
        orc_set_special_boundaries_velocity(a->nVertices, a->vertexBoundaryType, a->vertexBoundarySourceLocal,
                                            a->uVelocity, a->vVelocity);
    /* seaice_set_special_boundaries_velocity_masks (special_boundaries.F:345-401) */
    if (a->useSpecialBoundariesVelocityMasks) {
        for (int i = 0; i < a->nVertices; i++) a->solveVelocity[i] = a->solveVelocitySpecialBoundaries[i];
        for (int i = 0; i < a->nCells; i++) a->solveStress[i] = a->solveStressSpecialBoundaries[i];
    }
}

/* single_subcycle_velocity_solver (velocity_solver.F:2478-2592) with seaice_internal_stress
 * (:2606-2863, variational/variational branch).  Single block, so the halo exchange is a no-op. */
void orc_single_subcycle(const orc_subcycle_args *a)
{
    const int weakStrain = a->strainScheme == 2, weakDivergence = a->stressDivergenceScheme == 2;
    /* seaice_internal_stress (velocity_solver.F:2606-2863) */
    if (weakStrain)
        orc_strain_tensor_weak(a->nCells, a->maxEdges, a->nEdgesOnCell, a->verticesOnCell, a->edgesOnCell,
                               a->verticesOnEdge, a->dvEdge, a->areaCell, a->sphere_radius, a->solveStress,
                               a->uVelocity, a->vVelocity, a->normalVectorPolygon, a->latCellRotated,
                               a->strain11Weak, a->strain22Weak, a->strain12Weak);
    else
        orc_strain_tensor_variational(a->nCells, a->maxEdges, a->nEdgesOnCell, a->verticesOnCell, a->solveStress,
                                      a->uVelocity, a->vVelocity, a->basisGradientU, a->basisGradientV,
                                      a->tanLatVertexRotatedOverRadius, a->strain11, a->strain22, a->strain12);
    if (!weakStrain && a->averageVariationalStrains)
        orc_average_strains_on_vertex(a->nCells, a->nVerticesSolve, a->vertexDegree, a->maxEdges, a->cellsOnVertex,
                                      a->cellVerticesAtVertex, a->areaCell, a->strain11, a->strain22, a->strain12);
    if (weakStrain && !weakDivergence)
        orc_interpolate_strains_weak_to_variational(a->nCells, a->nVerticesSolve, a->vertexDegree, a->maxEdges,
                                                    a->nEdgesOnCell, a->verticesOnCell, a->cellsOnVertex, a->areaCell,
                                                    a->strain11Weak, a->strain22Weak, a->strain12Weak,
                                                    a->strain11Vertex, a->strain22Vertex, a->strain12Vertex,
                                                    a->strain11, a->strain22, a->strain12);
    if (weakDivergence) {
        orc_stress_tensor_weak(a->nCells, a->solveStress, a->constitutiveRelationType, a->elasticTimeStep,
                               a->dampingTimescale, a->icePressure, a->strain11Weak, a->strain22Weak, a->strain12Weak,
                               a->stress11Weak, a->stress22Weak, a->stress12Weak, a->replacementPressureWeak);
        orc_stress_divergence_weak(a->nVerticesSolve, a->vertexDegree, a->cellsOnVertex, a->edgesOnVertex,
                                   a->cellsOnEdge, a->dcEdge, a->areaTriangle, a->sphere_radius, a->solveVelocity,
                                   a->stress11Weak, a->stress22Weak, a->stress12Weak, a->normalVectorTriangle,
                                   a->latVertexRotated, a->stressDivergenceU, a->stressDivergenceV);
    } else {
        orc_stress_tensor_variational(a->nCells, a->maxEdges, a->nEdgesOnCell, a->solveStress,
                                      a->constitutiveRelationType, a->elasticTimeStep, a->dampingTimescale,
                                      a->icePressure, a->strain11, a->strain22, a->strain12,
                                      a->stress11, a->stress22, a->stress12, a->replacementPressure);
        orc_stress_divergence_variational(a->nVerticesSolve, a->vertexDegree, a->maxEdges, a->nEdgesOnCell,
                                          a->cellsOnVertex, a->cellVerticesAtVertex, a->solveVelocity,
                                          a->stress11, a->stress22, a->stress12,
                                          a->basisIntegralsU, a->basisIntegralsV, a->basisIntegralsMetric,
                                          a->variationalDenominator, a->tanLatVertexRotatedOverRadius,
                                          a->stressDivergenceU, a->stressDivergenceV);
    }
    orc_ocean_stress_coefficient(a->nVerticesSolve, a->nVertices, a->useOceanStress, a->oceanStressType,
                                 a->solveVelocity, a->iceAreaVertex, a->uOceanVelocityVertex, a->vOceanVelocityVertex,
                                 a->uVelocity, a->vVelocity, a->oceanStressCoeff);
    if (a->constitutiveRelationType == EVP) {
        orc_solve_velocity(a->nVerticesSolve, a->solveVelocity, a->elasticTimeStep, a->totalMassVertex,
                           a->totalMassVertexfVertex, a->stressDivergenceU, a->stressDivergenceV,
                           a->airStressVertexU, a->airStressVertexV, a->surfaceTiltForceU, a->surfaceTiltForceV,
                           a->oceanStressU, a->oceanStressV, a->oceanStressCoeff, a->uVelocity, a->vVelocity);
    } else if (a->constitutiveRelationType == EVP_REVISED) {
        orc_solve_velocity_revised(a->nVerticesSolve, a->solveVelocity, a->dynamicsTimeStep,
                                   a->numericalInertiaCoefficient, a->totalMassVertex, a->totalMassVertexfVertex,
                                   a->stressDivergenceU, a->stressDivergenceV, a->airStressVertexU, a->airStressVertexV,
                                   a->surfaceTiltForceU, a->surfaceTiltForceV, a->oceanStressU, a->oceanStressV,
                                   a->oceanStressCoeff, a->uVelocityInitial, a->vVelocityInitial,
                                   a->uVelocity, a->vVelocity);
    }
}

/* subcycle_velocity_solver (velocity_solver.F:2404-2464) */
void orc_subcycle_velocity_solver(const orc_subcycle_args *a, int nElasticSubcycle)
{
    special_boundaries(a);
    for (int k = 1; k <= nElasticSubcycle; k++) {
        orc_single_subcycle(a);
        special_boundaries(a);
    }
}

/* ==========================================================================================
 * pre-subcycle (host side of the boundary; used by the tests to build identical inputs)
 * ========================================================================================== */

/* seaice_interpolate_cell_to_vertex, the "#if 1 cell area" variant (src/shared/mpas_seaice_mesh.F:2835-2851).
 * No validity test on iCell: the junk slot (areaCell = -1e34) contaminates boundary vertices by design. */
void orc_interpolate_cell_to_vertex(int nVerticesSolve, int vertexDegree, const int *cellsOnVertex,
                                    const double *areaCell, const double *variableCell, double *variableVertex)
{
    const int D = vertexDegree;
    for (int iVertex = 1; iVertex <= nVerticesSolve; iVertex++) {
        double acc = 0.0, totalArea = 0.0;
        for (int k = 1; k <= D; k++) {
            const int iCell = cellsOnVertex[IDX2(k, iVertex, D)];
            acc = acc + areaCell[iCell - 1] * variableCell[iCell - 1];
            totalArea = totalArea + areaCell[iCell - 1];
        }
        variableVertex[iVertex - 1] = acc / totalArea;
    }
}

/* stress_calculation_mask (velocity_solver.F:961-1059) */
void orc_stress_calculation_mask(int nCells, int maxEdges, const int *nEdgesOnCell, const int *cellsOnCell,
                                 const double *iceAreaCellInitial, const double *totalMassCell,
                                 const int *landIceMask, int *solveStress)
{
    const int M = maxEdges;
    for (int iCell = 1; iCell <= nCells; iCell++) {
        solveStress[iCell - 1] = 0;
        if (iceAreaCellInitial[iCell - 1] > seaiceAreaMinimum && totalMassCell[iCell - 1] > seaiceMassMinimum &&
            landIceMask[iCell - 1] == 0) {
            solveStress[iCell - 1] = 1;
        } else {
            for (int k = 1; k <= nEdgesOnCell[iCell - 1]; k++) {
                const int nb = cellsOnCell[IDX2(k, iCell, M)];
                if (iceAreaCellInitial[nb - 1] > seaiceAreaMinimum && totalMassCell[nb - 1] > seaiceMassMinimum &&
                    landIceMask[nb - 1] == 0) {
                    solveStress[iCell - 1] = 1;
                    break;
                }
            }
        }
    }
}

/* velocity_calculation_mask (velocity_solver.F:1073-1150) */
void orc_velocity_calculation_mask(int nVerticesSolve, int nVertices, const int *interiorVertex,
                                   const int *landIceMaskVertex, const double *iceAreaVertex,
                                   const double *totalMassVertex, int *solveVelocity)
{
    for (int i = 0; i < nVerticesSolve; i++) {
        solveVelocity[i] = 0;
        if (interiorVertex[i] == 1 && landIceMaskVertex[i] == 0 && iceAreaVertex[i] > seaiceAreaMinimum &&
            totalMassVertex[i] > seaiceMassMinimum)
            solveVelocity[i] = 1;
    }
    for (int i = nVerticesSolve; i < nVertices; i++) solveVelocity[i] = 0;
}

/* init_special_boundaries_velocity / init_special_boundaries_tracers (special_boundaries.F:83-150, 164-250): the local
 * index of every special entity's source, from the global IDs the stream delivers.  globalToLocalID has nEntities slots
 * in the reference: it serves a block whose global IDs are 1..nEntities (one block per rank, one rank).  Returns 1 when
 * an ID falls outside (the reference would index out of bounds), leaving the output as far as written. */
int orc_boundary_source_local(int nEntities, const int *indexToID, const int *boundaryType, const int *boundarySource,
                              int *boundarySourceLocal)
{
    int *globalToLocalID = (int *)calloc((size_t)(nEntities > 0 ? nEntities : 1), sizeof(int));
    int rc = 0;
    for (int i = 1; i <= nEntities; i++) {
        const int id = indexToID[i - 1];
        if (id < 1 || id > nEntities) { rc = 1; break; }
        globalToLocalID[id - 1] = i;
    }
    for (int i = 1; i <= nEntities && !rc; i++) {
        if (boundaryType[i - 1] != 0) {
            const int src = boundarySource[i - 1];
            if (src < 1 || src > nEntities) { rc = 1; break; }
            boundarySourceLocal[i - 1] = globalToLocalID[src - 1];
        }
    }
    free(globalToLocalID);
    return rc;
}

/* seaice_set_special_boundaries_tracers (special_boundaries.F:415-485): in cell order, in place -- a source changed
 * earlier in the loop is read changed.  Arrays (nCells, n) with n = nCategories * (the ONE layer). */
void orc_set_special_boundaries_tracers(int nCells, int n, const int *tracerBoundaryType, const int *tracerBoundarySourceLocal,
                                        double *iceAreaCategory, double *iceVolumeCategory, double *snowVolumeCategory)
{
    double *f[3] = {iceAreaCategory, iceVolumeCategory, snowVolumeCategory};
    for (int iCell = 1; iCell <= nCells; iCell++) {
        if (tracerBoundaryType[iCell - 1] == 1) {
            for (int t = 0; t < 3; t++)
                for (int k = 0; k < n; k++) f[t][(size_t)(iCell - 1) * n + k] = 0.0;
        } else if (tracerBoundaryType[iCell - 1] == 2) {
            const int s = tracerBoundarySourceLocal[iCell - 1];
            for (int t = 0; t < 3; t++)
                for (int k = 0; k < n; k++) f[t][(size_t)(iCell - 1) * n + k] = f[t][(size_t)(s - 1) * n + k];
        }
    }
}

/* init_ice_shelve_vertex_mask (velocity_solver.F:481-544): a vertex of the owned range touching a land-ice cell */
void orc_ice_shelve_vertex_mask(int nVertices, int nVerticesSolve, int vertexDegree, const int *cellsOnVertex,
                                const int *landIceMask, int *landIceMaskVertex)
{
    const int D = vertexDegree;
    for (int i = 0; i < nVertices; i++) landIceMaskVertex[i] = 0;
    for (int iVertex = 1; iVertex <= nVerticesSolve; iVertex++) {
        for (int k = 1; k <= D; k++) {
            const int iCell = cellsOnVertex[IDX2(k, iVertex, D)];
            if (landIceMask[iCell - 1] == 1) landIceMaskVertex[iVertex - 1] = 1;
        }
    }
}

/* dynamically_locked_cell_mask (velocity_solver.F:402-467): 1 where no vertex of the cell is an interior vertex */
void orc_dynamically_locked_cell_mask(int nCells, int maxEdges, const int *nEdgesOnCell, const int *verticesOnCell,
                                      const int *interiorVertex, int *dynamicallyLockedCellsMask)
{
    const int M = maxEdges;
    for (int iCell = 1; iCell <= nCells; iCell++) {
        dynamicallyLockedCellsMask[iCell - 1] = 1;
        for (int k = 1; k <= nEdgesOnCell[iCell - 1]; k++) {
            const int iVertex = verticesOnCell[IDX2(k, iCell, M)];
            if (interiorVertex[iVertex - 1] == 1) {
                dynamicallyLockedCellsMask[iCell - 1] = 0;
                break;
            }
        }
    }
}

/* ice_strength, Hibler branch (velocity_solver.F:1419-1436) */
void orc_ice_strength_hibler(int nCellsSolve, const int *solveStress, const double *iceVolumeCell,
                             const double *iceAreaCell, double *icePressure)
{
    for (int i = 0; i < nCellsSolve; i++) {
        if (solveStress[i] == 1)
            icePressure[i] = seaiceIceStrengthConstantHiblerP * iceVolumeCell[i] *
                             exp(-seaiceIceStrengthConstantHiblerC * (1.0 - iceAreaCell[i]));
        else
            icePressure[i] = 0.0;
    }
}

/* constant_air_stress (velocity_solver.F:1665-1728), airStressCoeff = 0.0012 */
void orc_constant_air_stress(int nCellsSolve, const double *uAirVelocity, const double *vAirVelocity,
                             const double *airDensity, const double *iceAreaCell,
                             double *airStressCellU, double *airStressCellV)
{
    const double airStressCoeff = 0.0012;
    for (int i = 0; i < nCellsSolve; i++) {
        const double windSpeed = sqrt(uAirVelocity[i] * uAirVelocity[i] + vAirVelocity[i] * vAirVelocity[i]);
        airStressCellU[i] = airDensity[i] * windSpeed * airStressCoeff * uAirVelocity[i] * iceAreaCell[i];
        airStressCellV[i] = airDensity[i] * windSpeed * airStressCoeff * vAirVelocity[i] * iceAreaCell[i];
    }
}

/* coriolis_force_coefficient (velocity_solver.F:1742-1788) */
void orc_coriolis_force_coefficient(int nVerticesSolve, const double *totalMassVertex, const double *fVertex,
                                    double *totalMassVertexfVertex)
{
    for (int i = 0; i < nVerticesSolve; i++) totalMassVertexfVertex[i] = totalMassVertex[i] * fVertex[i];
}

/* ocean_stress (velocity_solver.F:1802-1883) */
void orc_ocean_stress(int nVerticesSolve, int nVertices, int useOceanStress, const int *solveVelocity,
                      const double *uOceanVelocityVertex, const double *vOceanVelocityVertex,
                      const double *fVertex, double *oceanStressU, double *oceanStressV)
{
    if (useOceanStress) {
        for (int i = 0; i < nVerticesSolve; i++) {
            if (solveVelocity[i] == 1) {
                const double sgn = copysign(1.0, fVertex[i]);
                oceanStressU[i] = uOceanVelocityVertex[i] * cosOceanTurningAngle -
                                  vOceanVelocityVertex[i] * sinOceanTurningAngle * sgn;
                oceanStressV[i] = uOceanVelocityVertex[i] * sinOceanTurningAngle * sgn +
                                  vOceanVelocityVertex[i] * cosOceanTurningAngle;
            } else {
                oceanStressU[i] = 0.0;
                oceanStressV[i] = 0.0;
            }
        }
    } else {
        for (int i = 0; i < nVertices + 1; i++) { oceanStressU[i] = 0.0; oceanStressV[i] = 0.0; }
    }
}

/* surface_tilt_geostrophic (velocity_solver.F:1941-2010); no_surface_tilt zeroes both (:2183-2213) */
void orc_surface_tilt(int nVerticesSolve, int nVertices, int useSurfaceTilt, const int *solveVelocity,
                      const double *fVertex, const double *totalMassVertex,
                      const double *uOceanVelocityVertex, const double *vOceanVelocityVertex,
                      double *surfaceTiltForceU, double *surfaceTiltForceV)
{
    if (useSurfaceTilt) {
        for (int i = 0; i < nVerticesSolve; i++) {
            if (solveVelocity[i] == 1) {
                surfaceTiltForceU[i] = -fVertex[i] * totalMassVertex[i] * vOceanVelocityVertex[i];
                surfaceTiltForceV[i] = fVertex[i] * totalMassVertex[i] * uOceanVelocityVertex[i];
            } else {
                surfaceTiltForceU[i] = 0.0;
                surfaceTiltForceV[i] = 0.0;
            }
        }
    } else {
        for (int i = 0; i < nVertices + 1; i++) { surfaceTiltForceU[i] = 0.0; surfaceTiltForceV[i] = 0.0; }
    }
}

/* new_ice_velocities, the vertex loop (velocity_solver.F:1250-1279) */
void orc_new_ice_velocities(int nVerticesSolve, int nVertices, const int *solveVelocity, int *solveVelocityPrevious,
                            const double *uOceanVelocityVertex, const double *vOceanVelocityVertex,
                            double *uVelocity, double *vVelocity, double *stressDivergenceU,
                            double *stressDivergenceV, double *oceanStressU, double *oceanStressV,
                            double *uVelocityInitial, double *vVelocityInitial)
{
    for (int i = 0; i < nVerticesSolve; i++) {
        if (solveVelocity[i] == 1) {
            if (solveVelocityPrevious[i] == 0) {
                uVelocity[i] = uOceanVelocityVertex[i];
                vVelocity[i] = vOceanVelocityVertex[i];
            }
        } else {
            uVelocity[i] = 0.0;
            vVelocity[i] = 0.0;
            stressDivergenceU[i] = 0.0;
            stressDivergenceV[i] = 0.0;
            oceanStressU[i] = 0.0;
            oceanStressV[i] = 0.0;
        }
    }
    for (int i = 0; i < nVertices + 1; i++) {
        solveVelocityPrevious[i] = solveVelocity[i];
        uVelocityInitial[i] = uVelocity[i];
        vVelocityInitial[i] = vVelocity[i];
    }
}

/* init_subcycle_variables, variational branch (velocity_solver.F:2227-2386) */
void orc_init_subcycle_variables(int nCells, int nVertices, int nVerticesSolve, int maxEdges,
                                 const int *solveStress, const int *solveVelocity,
                                 double *stressDivergenceU, double *stressDivergenceV,
                                 double *uVelocity, double *vVelocity, double *oceanStressCoeff,
                                 double *strain11, double *strain22, double *strain12,
                                 double *stress11, double *stress22, double *stress12)
{
    const int M = maxEdges;
    for (int i = 0; i < nVertices + 1; i++) { stressDivergenceU[i] = 0.0; stressDivergenceV[i] = 0.0; }
    for (int i = 0; i < nVerticesSolve; i++) {
        if (solveVelocity[i] != 1) { uVelocity[i] = 0.0; vVelocity[i] = 0.0; oceanStressCoeff[i] = 0.0; }
    }
    for (size_t q = 0; q < (size_t)M * (size_t)(nCells + 1); q++) { strain11[q] = 0.0; strain22[q] = 0.0; strain12[q] = 0.0; }
    for (int iCell = 1; iCell <= nCells; iCell++) {
        if (solveStress[iCell - 1] != 1) {
            for (int k = 1; k <= M; k++) {
                stress11[IDX2(k, iCell, M)] = 0.0;
                stress22[IDX2(k, iCell, M)] = 0.0;
                stress12[IDX2(k, iCell, M)] = 0.0;
            }
        }
    }
}

/* ==========================================================================================
 * post-subcycle
 * ========================================================================================== */

/* seaice_final_divergence_shear_variational (variational.F:1198-1330), incl. the unit change :1324-1325 */
void orc_final_divergence_shear_variational(int nCells, int maxEdges, const int *nEdgesOnCell,
                                            const int *solveStress,
                                            const double *strain11, const double *strain22, const double *strain12,
                                            double *divergence, double *shear,
                                            double *ridgeConvergence, double *ridgeShear)
{
    const int M = maxEdges;
    for (int iCell = 1; iCell <= nCells; iCell++) {
        if (solveStress[iCell - 1] == 1) {
            double dSum = 0.0, tSum = 0.0, sSum = 0.0, DeltaAverage = 0.0;
            const int n = nEdgesOnCell[iCell - 1];
            for (int j = 1; j <= n; j++) {
                const size_t q = IDX2(j, iCell, M);
                const double sd = strain11[q] + strain22[q];
                const double st = strain11[q] - strain22[q];
                const double ss = strain12[q] * 2.0;
                const double Delta = sqrt(sd * sd + (st * st + ss * ss) / eccentricitySquared);
                dSum = dSum + sd;
                tSum = tSum + st;
                sSum = sSum + ss;
                DeltaAverage = DeltaAverage + Delta;
            }
            divergence[iCell - 1] = dSum / (double)n;
            shear[iCell - 1] = sqrt(tSum * tSum + sSum * sSum) / (double)n;
            DeltaAverage = DeltaAverage / (double)n;
            if (ridgeConvergence) {
                ridgeConvergence[iCell - 1] = -fmin(divergence[iCell - 1], 0.0);
                ridgeShear[iCell - 1] = 0.5 * (DeltaAverage - fabs(divergence[iCell - 1]));
            }
        } else {
            divergence[iCell - 1] = 0.0;
            shear[iCell - 1] = 0.0;
            if (ridgeConvergence) { ridgeConvergence[iCell - 1] = 0.0; ridgeShear[iCell - 1] = 0.0; }
        }
    }
    for (int i = 0; i < nCells; i++) {
        divergence[i] = divergence[i] * 100.0 * 86400.0;
        shear[i] = shear[i] * 100.0 * 86400.0;
    }
}

/* principal_stresses (velocity_solver.F:3565-3610) over the variational stress points (:3520-3540) */
void orc_principal_stresses_variational(int nCellsSolve, int maxEdges, const int *nEdgesOnCell,
                                        const double *stress11, const double *stress22, const double *stress12,
                                        const double *replacementPressure,
                                        double *principalStress1, double *principalStress2)
{
    const int M = maxEdges;
    for (int iCell = 1; iCell <= nCellsSolve; iCell++) {
        for (int j = 1; j <= nEdgesOnCell[iCell - 1]; j++) {
            const size_t q = IDX2(j, iCell, M);
            if (replacementPressure[q] > seaicePuny) {
                const double sqrtContents = (stress11[q] + stress22[q]) * (stress11[q] + stress22[q]) -
                                            4.0 * stress11[q] * stress22[q] + 4.0 * (stress12[q] * stress12[q]);
                double p1 = 0.5 * (stress11[q] + stress22[q]) + 0.5 * sqrt(sqrtContents);
                double p2 = 0.5 * (stress11[q] + stress22[q]) - 0.5 * sqrt(sqrtContents);
                principalStress1[q] = p1 / replacementPressure[q];
                principalStress2[q] = p2 / replacementPressure[q];
            } else {
                principalStress1[q] = 1.0e30;
                principalStress2[q] = 1.0e30;
            }
        }
    }
}

/* seaice_interpolate_vertex_to_cell (src/shared/mpas_seaice_mesh.F:2906-2976) */
void orc_interpolate_vertex_to_cell(int nCellsSolve, int maxEdges, const int *nEdgesOnCell, const int *verticesOnCell,
                                    const double *areaTriangle, const int *interiorVertex,
                                    const double *variableVertex, double *variableCell)
{
    const int M = maxEdges;
    for (int iCell = 1; iCell <= nCellsSolve; iCell++) {
        double totalArea = 0.0;
        variableCell[iCell - 1] = 0.0;
        for (int k = 1; k <= nEdgesOnCell[iCell - 1]; k++) {
            const int iVertex = verticesOnCell[IDX2(k, iCell, M)];
            variableCell[iCell - 1] = variableCell[iCell - 1] +
                areaTriangle[iVertex - 1] * variableVertex[iVertex - 1] * (double)interiorVertex[iVertex - 1];
            totalArea = totalArea + areaTriangle[iVertex - 1] * (double)interiorVertex[iVertex - 1];
        }
        if (totalArea > 0.0) variableCell[iCell - 1] = variableCell[iCell - 1] / totalArea;
    }
}

/* ocean_stress_final (src/shared/mpas_seaice_velocity_solver.F:3624-3848); the caller has already run
 * ocean_stress_coefficient on the final velocities (:3686) and exchanged the halo where there is one */
void orc_ocean_stress_final(int nVerticesSolve, int nVertices, int nCellsSolve, int nCells, int maxEdges,
                            int useOceanStress, const int *solveVelocity, const double *oceanStressCoeff,
                            const double *uOceanVelocityVertex, const double *vOceanVelocityVertex,
                            const double *uVelocity, const double *vVelocity, const double *fVertex,
                            const double *iceAreaVertex, const int *nEdgesOnCell, const int *verticesOnCell,
                            const double *areaTriangle, const int *interiorVertex,
                            double *oceanStressU, double *oceanStressV, double *oceanStressCellU, double *oceanStressCellV)
{
    if (useOceanStress) {
        for (int i = 0; i < nVerticesSolve; i++) {
            if (solveVelocity[i] == 1) {
                const double sgn = copysign(1.0, fVertex[i]);
                oceanStressU[i] = oceanStressCoeff[i] *
                    ((uOceanVelocityVertex[i] - uVelocity[i]) * cosOceanTurningAngle -
                     (vOceanVelocityVertex[i] - vVelocity[i]) * sinOceanTurningAngle * sgn);
                oceanStressV[i] = oceanStressCoeff[i] *
                    ((vOceanVelocityVertex[i] - vVelocity[i]) * cosOceanTurningAngle +
                     (uOceanVelocityVertex[i] - uVelocity[i]) * sinOceanTurningAngle * sgn);
                oceanStressU[i] = oceanStressU[i] / iceAreaVertex[i];
                oceanStressV[i] = oceanStressV[i] / iceAreaVertex[i];
            } else {
                oceanStressU[i] = 0.0;
                oceanStressV[i] = 0.0;
            }
        }
        orc_interpolate_vertex_to_cell(nCellsSolve, maxEdges, nEdgesOnCell, verticesOnCell, areaTriangle, interiorVertex,
                                       oceanStressU, oceanStressCellU);
        orc_interpolate_vertex_to_cell(nCellsSolve, maxEdges, nEdgesOnCell, verticesOnCell, areaTriangle, interiorVertex,
                                       oceanStressV, oceanStressCellV);
        for (int i = 0; i < nVerticesSolve; i++) {
            if (solveVelocity[i] == 1) {
                oceanStressU[i] = oceanStressU[i] * iceAreaVertex[i];
                oceanStressV[i] = oceanStressV[i] * iceAreaVertex[i];
            }
        }
    } else {
        for (int i = 0; i < nVertices + 1; i++) { oceanStressU[i] = 0.0; oceanStressV[i] = 0.0; }
        for (int i = 0; i < nCells + 1; i++) { oceanStressCellU[i] = 0.0; oceanStressCellV[i] = 0.0; }
    }
}

/* surface_tilt_ssh_gradient, the vertex loop (velocity_solver.F:2140-2160); seaiceGravity = 9.80616 */
void orc_surface_tilt_ssh_gradient(int nVerticesSolve, const int *solveVelocity, const double *totalMassVertex,
                                   const double *seaSurfaceTiltVertexU, const double *seaSurfaceTiltVertexV,
                                   double *surfaceTiltForceU, double *surfaceTiltForceV)
{
    const double seaiceGravity = 9.80616;
    for (int i = 0; i < nVerticesSolve; i++) {
        if (solveVelocity[i] == 1) {
            surfaceTiltForceU[i] = -seaiceGravity * totalMassVertex[i] * seaSurfaceTiltVertexU[i];
            surfaceTiltForceV[i] = -seaiceGravity * totalMassVertex[i] * seaSurfaceTiltVertexV[i];
        } else {
            surfaceTiltForceU[i] = 0.0;
            surfaceTiltForceV[i] = 0.0;
        }
    }
}

/* seaice_init_evp scalars (constitutive_relation.F:125, 154-162) */
double orc_damping_timescale(double dynamicsTimeStep) { return 0.36 * dynamicsTimeStep; }
double orc_numerical_inertia_coefficient(double dynamicsTimeStep, double dvEdgeMinGlobal)
{
    const double gamma = 0.25 * 1.0e11 * dynamicsTimeStep;
    return (2.0 * dampingRatioDenominator * dampingRatio * gamma) / (dvEdgeMinGlobal * dvEdgeMinGlobal);
}

/* aggregate_mass_and_area for one category (velocity_solver.F:738-746) */
/* aggregate_mass_and_area (velocity_solver.F:685-752): sums over the categories in category order (layer 1 of the
 * (1, nCategories, nCells) tracer arrays), then the total mass */
void orc_aggregate_mass_and_area(int nCells, int nCategories, const double *iceAreaCategory,
                                 const double *iceVolumeCategory, const double *snowVolumeCategory, double *iceAreaCell,
                                 double *iceVolumeCell, double *snowVolumeCell, double *totalMassCell)
{
    for (int i = 0; i < nCells; i++) {
        double a = 0.0, vi = 0.0, vs = 0.0;
        for (int k = 0; k < nCategories; k++) {
            a = a + iceAreaCategory[(size_t)i * nCategories + k];
            vi = vi + iceVolumeCategory[(size_t)i * nCategories + k];
            vs = vs + snowVolumeCategory[(size_t)i * nCategories + k];
        }
        iceAreaCell[i] = a;
        iceVolumeCell[i] = vi;
        snowVolumeCell[i] = vs;
        totalMassCell[i] = vi * seaiceDensityIce + vs * seaiceDensitySnow;
    }
}

void orc_total_mass(int nCells, const double *iceVolumeCell, const double *snowVolumeCell, double *totalMassCell)
{
    for (int i = 0; i < nCells; i++)
        totalMassCell[i] = iceVolumeCell[i] * seaiceDensityIce + snowVolumeCell[i] * seaiceDensitySnow;
}
