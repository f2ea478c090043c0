/*
 * evp_b200.h -- C ABI of the B200-native EVP momentum subcycle for MPAS-Seaice.
 *
 * Drop-in boundary: the body of subroutine subcycle_velocity_solver
 * (reference: src/shared/mpas_seaice_velocity_solver.F:2404-2464) and the device-residency
 * lifecycle the reference already has in module seaice_mesh_pool
 * (src/shared/mpas_seaice_mesh_pool.F:76-177 create, :261-281 update, :188-250 destroy).
 * The Fortran host binds these symbols with ISO_C_BINDING (fortran/seaice_evp_b200.F90,
 * INTEGRATION.md); nothing here mentions torch, C++ or CUDA types.
 *
 * Conventions for every array argument (what c_loc() of an MPAS pool array gives):
 *   - contiguous, Fortran column-major, first dimension maxEdges / vertexDegree;
 *   - index VALUES are 1-based; an invalid neighbour is any value outside 1..n (MPAS uses n+1);
 *   - cell arrays hold at least nCells (owned + halo) columns, vertex arrays at least nVertices
 *     entries; the MPAS junk element n+1 may be present and is never dereferenced;
 *   - the library copies what it needs during the call and never keeps or frees a host pointer.
 *
 * Every function returns 0 on success, non-zero on error (EVP_ERR_*); evp_last_error_string()
 * gives the text.  No function throws.  One handle <-> one GPU <-> one rank (block); a handle is not
 * re-entrant, matching the single main thread of an MPAS rank (mesh_pool.F:98-105).
 */
#ifndef EVP_B200_H
#define EVP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct evp_handle evp_handle;

enum {
    EVP_OK = 0,
    EVP_ERR_ARGUMENT = 1,     /* NULL / out-of-range / unsupported combination */
    EVP_ERR_CUDA = 2,         /* a CUDA runtime call failed (no device, OOM, launch failure) */
    EVP_ERR_NCCL = 3,         /* NCCL missing or a NCCL call failed */
    EVP_ERR_STATE = 4         /* call order violated (e.g. run before update_step) */
};

/* config_constitutive_relation_type
 * (src/shared/mpas_seaice_velocity_solver_constitutive_relation.F:34-38) */
enum { EVP_CR_EVP = 1, EVP_CR_EVP_REVISED = 2, EVP_CR_LINEAR = 3, EVP_CR_NONE = 4 };
/* config_ocean_stress_type (src/shared/mpas_seaice_velocity_solver.F:55-57) */
enum { EVP_OCEAN_QUADRATIC = 1, EVP_OCEAN_LINEAR = 2 };
/* vertexBoundaryType (src/shared/mpas_seaice_special_boundaries.F:35-39) */
enum { EVP_VB_NONE = 0, EVP_VB_PERIODIC = 1, EVP_VB_REVERSE = 2, EVP_VB_ZERO = 3 };

/* config_strain_scheme / config_stress_divergence_scheme (src/shared/mpas_seaice_velocity_solver.F:168-198);
 * 0 is read as variational.  weak strain + variational divergence is allowed, the reverse is rejected
 * like the reference does (:195-198). */
enum { EVP_SCHEME_VARIATIONAL = 1, EVP_SCHEME_WEAK = 2 };

/* evp_options.flags */
enum {
    EVP_FLAG_OVERLAP_HALO = 2,/* accepted and ignored (kept for ABI stability): round 1's forked pack / NCCL / unpack branch
                                 measured no gain and is superseded by the peer-to-peer exchange, in which the vertex
                                 kernel itself stores boundary velocities into the neighbours (evp_set_halo) */
    EVP_FLAG_PIN_HOST = 1     /* host arrays passed to update_step / fetch live at stable addresses for the
                                 life of the handle (true for MPAS pool arrays): page-lock them once with
                                 cudaHostRegister so the per-step copies run at full PCIe speed.  The arrays must
                                 NOT be freed while registered: call evp_release_host_memory() first (or
                                 evp_destroy) -- a stale registration can collide with later allocations */
};

/* Static description of one block: the variable list of module seaice_mesh_pool
 * (mesh_pool.F:23-57, :146-171) plus variationalDenominator (passed as an argument at
 * velocity_solver.F:2852) and the special-boundary maps (special_boundaries.F:295-296). */
typedef struct {
    int nCells;              /* owned + halo: loop bound of the strain / stress loops (variational.F:633,860) */
    int nCellsSolve;         /* owned cells (informational) */
    int nVertices;           /* owned + halo */
    int nVerticesSolve;      /* owned: loop bound of divergence / drag / solve (variational.F:1135) */
    int maxEdges;            /* 4, 6 or 8 supported (7 is handled by the 8 instantiation) */
    int vertexDegree;        /* 3 or 4 */
    const int *nEdgesOnCell;             /* (nCells) */
    const int *verticesOnCell;           /* (maxEdges, nCells) */
    const int *cellsOnVertex;            /* (vertexDegree, nVertices) */
    const int *cellVerticesAtVertex;     /* (vertexDegree, nVertices); 0 = vertex not in that cell */
    const double *basisGradientU;        /* (maxEdges, maxEdges, nCells)  (iBasisVertex, iGradientVertex, iCell) */
    const double *basisGradientV;
    const double *basisIntegralsU;       /* (maxEdges, maxEdges, nCells)  (iStressVertex, iVelocityVertex, iCell) */
    const double *basisIntegralsV;
    const double *basisIntegralsMetric;
    const double *tanLatVertexRotatedOverRadius; /* (nVertices) */
    const double *variationalDenominator;        /* (nVertices) */
    const int *vertexBoundaryType;        /* (nVertices) or NULL when special boundaries are off */
    const int *vertexBoundarySourceLocal; /* (nVertices) or NULL */
} evp_mesh_desc;

/* Namelist options that shape the subcycle (src/Registry.xml:566-647) and the scalars of
 * seaice_init_evp (constitutive_relation.F:125,154-162). */
typedef struct {
    int constitutive_relation_type;           /* EVP_CR_*  */
    int ocean_stress_type;                    /* EVP_OCEAN_* */
    int use_ocean_stress;                     /* config_use_ocean_stress */
    int use_special_boundaries_velocity;      /* config_use_special_boundaries_velocity */
    int device;                               /* CUDA device ordinal, -1 = current device */
    int flags;                                /* EVP_FLAG_* */
    int average_variational_strain;           /* config_average_variational_strain (seaice_average_strains_on_vertex,
                                                 variational.F:684-763); needs evp_set_mesh_ext (areaCell) */
    int strain_scheme;                        /* EVP_SCHEME_*; weak needs evp_set_weak_mesh */
    int stress_divergence_scheme;             /* EVP_SCHEME_* */
    double elasticTimeStep;                   /* velocity_solver.F:157 */
    double dynamicsTimeStep;                  /* velocity_solver.F:155 */
    double dampingTimescale;                  /* constitutive_relation.F:125 */
    double numericalInertiaCoefficient;       /* constitutive_relation.F:159 (evp_revised only) */
} evp_options;

/* Per-dynamics-step inputs: what seaice_mesh_pool_update refreshes (mesh_pool.F:269-278) plus the
 * vertex fields read by ocean_stress_coefficient / solve_velocity that the reference kept on the
 * host (velocity_solver.F:3036-3042, 3152-3167).  All (nVertices) / (nCells) / (maxEdges, nCells). */
typedef struct {
    const int *solveStress;               /* (nCells) */
    const int *solveVelocity;             /* (nVertices) */
    const double *icePressure;            /* (nCells) */
    const double *uVelocity;              /* (nVertices) */
    const double *vVelocity;
    const double *stress11;               /* (maxEdges, nCells) */
    const double *stress22;
    const double *stress12;
    const double *totalMassVertex;        /* (nVertices) */
    const double *totalMassVertexfVertex;
    const double *iceAreaVertex;
    const double *airStressVertexU;
    const double *airStressVertexV;
    const double *surfaceTiltForceU;
    const double *surfaceTiltForceV;
    const double *oceanStressU;
    const double *oceanStressV;
    const double *uOceanVelocityVertex;
    const double *vOceanVelocityVertex;
    const double *uVelocityInitial;       /* evp_revised only, else may be NULL */
    const double *vVelocityInitial;
} evp_step_fields;

/* Outputs consumed by velocity_solver_post_subcycle (velocity_solver.F:3360-3380) and by the
 * restart stream (src/Registry.xml:1937-1957).  Any pointer may be NULL (= not wanted). */
typedef struct {
    double *uVelocity;                    /* (nVertices) */
    double *vVelocity;
    double *stress11;                     /* (maxEdges, nCells) */
    double *stress22;
    double *stress12;
    double *strain11;                     /* (maxEdges, nCells): strain of the LAST subcycle */
    double *strain22;
    double *strain12;
    double *replacementPressure;          /* (maxEdges, nCells) */
    double *stressDivergenceU;            /* (nVertices) */
    double *stressDivergenceV;
    double *oceanStressCoeff;             /* (nVertices) */
} evp_out_fields;

/* seaice_mesh_pool_create equivalent.  Must be called AFTER seaice_init_velocity_solver filled the
 * basis arrays (initialize.F:121).  The five basis pointers may all be NULL if
 * evp_precompute_wachspress() is going to fill them on the device. */
int evp_create(evp_handle **handle, const evp_mesh_desc *mesh, const evp_options *options);

/* Change the scalars / switches without rebuilding the mesh (e.g. config_dt changed). */
int evp_set_options(evp_handle *handle, const evp_options *options);

/* Device version of seaice_init_velocity_solver_wachspress
 * (src/shared/mpas_seaice_velocity_solver_wachspress.F:46-161) writing straight into the device layout.
 * xLocal, yLocal: (maxEdges, nCells) from seaice_calc_local_coords.  config_wachspress_integration_type /
 * _order (Registry.xml:603-610): integrationType 0 = 'dunavant' (orders 1..10, 12), 1 = 'trapezoidal' (orders 1..9:
 * at most 64 points), 2 = 'fekete' (orders 1..6, 8, 9).  Bit-identical to the FP64 non-FMA evaluation of the
 * reference formulas. */
int evp_precompute_wachspress(evp_handle *handle, const double *xLocal, const double *yLocal,
                              int integrationType, int integrationOrder);

/* The points, weights and normalisation evp_precompute_wachspress uses for (integrationType, integrationOrder):
 * get_integration_factors (wachspress.F:1224-1287).  Host-only (no device needed); u, v, w: room for 64 doubles each. */
int evp_integration_rule(int integrationType, int integrationOrder, int *nPoints, double *u, double *v, double *w,
                         double *normalizationFactor);

/* Device version of seaice_init_velocity_solver_pwl (src/shared/mpas_seaice_velocity_solver_pwl.F:44-373,
 * config_variational_basis = 'pwl'), incl. the 3x3 LU solves of src/shared/mpas_seaice_numerics.F:44-212.
 * edgesOnCell (maxEdges, nCells), dvEdge (nEdges), areaCell (nCells).  Bit-identical to the non-FMA host
 * evaluation; the gradients are dense, so the cell kernel takes its dense-gradient path. */
int evp_precompute_pwl(evp_handle *handle, const double *xLocal, const double *yLocal, const int *edgesOnCell,
                       const double *dvEdge, int nEdges, const double *areaCell);

/* Copy the basis arrays back to host arrays in the Registry layout (any pointer may be NULL). */
int evp_fetch_basis(evp_handle *handle, double *basisGradientU, double *basisGradientV,
                    double *basisIntegralsU, double *basisIntegralsV, double *basisIntegralsMetric);

/* seaice_mesh_pool_update equivalent: upload one dynamics step's inputs. */
int evp_update_step(evp_handle *handle, const evp_step_fields *fields);

/* Replace the masks by the special-boundary masks (seaice_set_special_boundaries_velocity_masks,
 * special_boundaries.F:345-401).  Optional. */
int evp_set_masks(evp_handle *handle, const int *solveStress, const int *solveVelocity);

/* subcycle_velocity_solver: nSubcycles x (strain -> stress -> divergence -> drag -> solve -> halo),
 * special boundaries before the loop and after every subcycle.  Asynchronous on the handle's
 * stream; evp_fetch / evp_synchronize block. */
int evp_run_subcycles(evp_handle *handle, int nSubcycles);

int evp_synchronize(evp_handle *handle);

/* Blocking copy of the results into host arrays. */
int evp_fetch(evp_handle *handle, const evp_out_fields *out);

/* Undo every cudaHostRegister done under EVP_FLAG_PIN_HOST (before the host frees or reallocates arrays it
 * passed to update_step / fetch / pre_subcycle / post_subcycle). */
int evp_release_host_memory(evp_handle *handle);

/* seaice_mesh_pool_destroy equivalent. */
int evp_destroy(evp_handle *handle);

const char *evp_last_error_string(void);

/* ---- multi-GPU: the per-subcycle uVelocity/vVelocity halo exchange (velocity_solver.F:2543-2584) ----
 * One rank per GPU.  Exchange lists come from the host's decomposition (in MPAS: the
 * mpas_dmpar exchange lists of the 'velocityHaloExchangeGroup', velocity_solver.F:259-349):
 * for neighbour k, sendIndex[sendOffset[k] .. sendOffset[k+1]) are the local 1-based owned vertices
 * whose (u,v) this rank sends to rank neighbourRank[k]; recvIndex likewise are the local halo
 * vertices filled from that rank, in the sender's send order.
 * evp_comm_init, evp_set_halo and (once a halo is attached) evp_destroy are COLLECTIVE: every rank of the
 * communicator must call them.  evp_set_halo chooses the exchange:
 *   EVP_HALO_P2P   every neighbour's velocity array could be mapped (cudaIpc*, one process per GPU on one NVLink
 *                  node) and all ranks agreed: the vertex kernel stores the (u,v) of boundary-owned vertices straight
 *                  into double-buffered halo slots of the neighbours and publishes a per-pass flag; the cell kernel
 *                  waits on the flags only where it gathers a halo vertex.  No communication kernel, no NCCL node in
 *                  the captured graph.  All ranks must then call evp_run_subcycles with the same counts.
 *   EVP_HALO_NCCL  otherwise (or EVP_B200_HALO=nccl in the environment; also whenever the weak operators or the
 *                  special boundaries are switched on): pack kernel + grouped ncclSend/ncclRecv on the handle's stream
 *                  inside the graph, received in place when each neighbour's halo vertices are one contiguous run of the
 *                  local numbering.  set_halo performs one eager warm-up exchange so that NCCL's lazy connection
 *                  set-up never happens inside the captured subcycle graph.
 * EVP_B200_HALO=p2p makes evp_set_halo fail instead of falling back.  Results are bit-identical either way. */
enum { EVP_HALO_NONE = 0, EVP_HALO_NCCL = 1, EVP_HALO_P2P = 2 };
int evp_comm_get_unique_id(char *id128);   /* rank 0 calls this, host broadcasts the 128 bytes (MPI_Bcast) */
int evp_comm_init(evp_handle *handle, int rank, int nRanks, const char *id128);
int evp_set_halo(evp_handle *handle, int nNeighbours, const int *neighbourRank,
                 const int *sendOffset, const int *sendIndex,
                 const int *recvOffset, const int *recvIndex);
/* Which exchange the next evp_run_subcycles uses; `why` (may be NULL) receives the reason for a fallback. */
int evp_halo_mode(evp_handle *handle, int *mode, char *why, int whyLen);


/* ======================================================================================================
 * Widening: the steps on either side of the subcycle on the device, so that one
 * dynamics step moves CELL fields in and a handful of fields out instead of ~20 vertex fields each way,
 * and u, v, stress11/22/12 and solveVelocityPrevious stay resident between steps.
 *   evp_pre_subcycle  = velocity_solver_pre_subcycle  (src/shared/mpas_seaice_velocity_solver.F:613-671)
 *   evp_post_subcycle = velocity_solver_post_subcycle (velocity_solver.F:3360-3380)
 * aggregate_mass_and_area (:685-752, sums over the ice categories) and the Hibler ice strength (:1419-1436) may
 * run on the host (the caller passes iceAreaCell, totalMassCell and an unmasked icePressure) or on the device
 * (evp_aggregate below, then evp_pre_subcycle with those pointers NULL).  colpkg_ice_strength (the Rothrock
 * strength of the column package, :1438-1460) is column physics and stays with the host.
 * ====================================================================================================== */

/* aggregate_mass_and_area on the device: the three category tracers of the `tracers` pool, layer 1, stored
 * (nCategories, nCells) = the Registry's (ONE, nCategories, nCells).  The sums run in category order like the
 * reference's sum() (bit-identical to a sequential host sum); totalMassCell = iceVolumeCell * rho_i + snowVolumeCell *
 * rho_s.  With hibler_strength != 0 the unmasked Hibler strength P* h exp(-C (1 - a)) is evaluated too -- with the
 * DEVICE's exp(), which is specified to 1 ulp and therefore need not equal the host libm's in the last bit.  Measured
 * on QU240 (10 242 cells, polar caps in 5 categories) it does: no cell's icePressure differs from glibc's and a full
 * dynamics step is bit-identical either way (tests/test_gpu_prepost.py::test_device_hibler_strength, which asserts
 * <= 1 ulp and <= 1e-9 after 120 subcycles and prints the measured numbers).  Hosts that must not depend on that
 * libm pass icePressure themselves.  The results stay on the device and feed the next evp_pre_subcycle whose
 * iceAreaCell / totalMassCell / icePressure (and iceAreaCellInitial, if it is to be the same field) pointers are NULL;
 * evp_fetch_aggregate copies them to the host arrays of the tracers_aggregate / icestate / velocity_solver pools. */
typedef struct {
    int nCategories;
    const double *iceAreaCategory;
    const double *iceVolumeCategory;
    const double *snowVolumeCategory;
} evp_category_fields;
int evp_aggregate(evp_handle *handle, const evp_category_fields *categories, int hibler_strength);
/* any pointer may be NULL; icePressure is the UNMASKED strength (the mask is applied by evp_pre_subcycle) */
int evp_fetch_aggregate(evp_handle *handle, double *iceAreaCell, double *iceVolumeCell, double *snowVolumeCell,
                        double *totalMassCell, double *icePressure);

/* Mesh fields the pre-/post-subcycle read in addition to evp_mesh_desc (src/Registry.xml:2251-2367,
 * boundary pool interiorVertex, ocean_coupling landIceMaskVertex). */
typedef struct {
    const int *cellsOnCell;          /* (maxEdges, nCells)  stress_calculation_mask, velocity_solver.F:1030-1040 */
    const int *interiorVertex;       /* (nVertices)  src/shared/mpas_seaice_mesh.F:423-488 */
    const int *landIceMaskVertex;    /* (nVertices) or NULL = no land ice (velocity_solver.F:481-544) */
    const double *areaCell;          /* (nCells)     weights of seaice_interpolate_cell_to_vertex, mesh.F:2835-2851 */
    const double *areaTriangle;      /* (nVertices)  weights of seaice_interpolate_vertex_to_cell, mesh.F:2958-2971 */
    const double *fVertex;           /* (nVertices) */
} evp_mesh_ext;

/* Cell inputs of one dynamics step, all (nCells).  Optional groups are NULL when unused. */
typedef struct {
    const double *iceAreaCellInitial;   /* masks + vertex interpolation (velocity_solver.F:860-880); NULL = iceAreaCell */
    const double *iceAreaCell;          /* constant_air_stress (:1716-1723); may alias iceAreaCellInitial;
                                           NULL = the result of evp_aggregate */
    const double *totalMassCell;        /* aggregate_mass_and_area (:742-744); NULL = the result of evp_aggregate */
    const double *icePressure;          /* ice strength, unmasked; the device zeroes it where solveStress /= 1;
                                           NULL = the Hibler strength of evp_aggregate */
    const double *uOceanVelocity;       /* ocean_coupling pool */
    const double *vOceanVelocity;
    const double *airStressCellU;       /* either the coupler's stresses ...                                   */
    const double *airStressCellV;
    const double *uAirVelocity;         /* ... or the inputs of constant_air_stress (used when airStressCellU is NULL) */
    const double *vAirVelocity;
    const double *airDensity;
    const double *seaSurfaceTiltU;      /* surface_tilt_ssh_gradient (:2024-2170) only */
    const double *seaSurfaceTiltV;
    const int *landIceMask;             /* (nCells) or NULL = no land ice */
    const int *solveStress;             /* config_calc_velocity_masks = false: masks given by the host ...   */
    const int *solveVelocity;           /* ... (nCells) / (nVertices), else NULL                              */
} evp_pre_fields;

#define EVP_START_RESIDENT 0
#define EVP_START_FROM_REST 1
#define EVP_START_FIRST_STEP 2

/* Namelist switches of the pre-subcycle (src/Registry.xml:566-647). */
typedef struct {
    int use_air_stress;                 /* config_use_air_stress */
    int use_surface_tilt;               /* config_use_surface_tilt */
    int geostrophic_surface_tilt;       /* config_geostrophic_surface_tilt */
    int calc_velocity_masks;            /* config_calc_velocity_masks */
    int cold_start;                     /* EVP_START_RESIDENT (0): use the state resident on the device (previous step,
                                           or seeded with evp_update_step / evp_set_state);
                                           EVP_START_FROM_REST (1): u = v = 0, stresses = 0 and solveVelocityPrevious =
                                           solveVelocity -- ice at rest that is NOT treated as new ice (what a restart
                                           file with zero velocities gives; not the reference's first step);
                                           EVP_START_FIRST_STEP (2): the reference's first step without a restart file:
                                           stresses = 0 and solveVelocityPrevious = 0 (no Registry default, assigned only
                                           at velocity_solver.F:1274), so every solved vertex is new ice and starts at the
                                           interpolated ocean velocity (velocity_solver.F:1252-1258) */
} evp_pre_options;

/* Outputs of one dynamics step; any pointer may be NULL (= not wanted, nothing computed for it beyond what
 * others need).  Cell arrays (nCells), vertex arrays (nVertices), *Var (maxEdges, nCells). */
typedef struct {
    double *uVelocity;                  /* -> advection */
    double *vVelocity;
    double *divergence;                 /* seaice_final_divergence_shear_variational, variational.F:1198-1330, or with
                                           the weak divergence scheme seaice_final_divergence_shear_weak, weak.F:651-751
                                           (no unit change there, and ridgeShear from the last owned cell's Delta --
                                           the reference assigns its whole Delta work array inside the loop, :729) */
    double *shear;
    double *ridgeConvergence;           /* -> ridging */
    double *ridgeShear;
    double *principalStress1Var;        /* principal_stresses_driver, velocity_solver.F:3443-3610 */
    double *principalStress2Var;
    double *oceanStressCellU;           /* ocean_stress_final, velocity_solver.F:3624-3848 -> coupler */
    double *oceanStressCellV;
    double *oceanStressU;               /* (nVertices) as left by ocean_stress_final */
    double *oceanStressV;
    double *oceanStressCoeff;
    double *principalStress1Weak;       /* (nCells) weak stress divergence scheme only (velocity_solver.F:3500-3515) */
    double *principalStress2Weak;
} evp_post_fields;

int evp_set_mesh_ext(evp_handle *handle, const evp_mesh_ext *ext);

/* Seed the state that is carried between dynamics steps (restart: src/Registry.xml:1937-1957). */
int evp_set_state(evp_handle *handle, const double *uVelocity, const double *vVelocity,
                  const double *stress11, const double *stress22, const double *stress12,
                  const int *solveVelocityPrevious);

/* velocity_solver_pre_subcycle on the device; afterwards the handle is ready for evp_run_subcycles.
 * Multi-rank: ends with the uVelocity/vVelocity halo exchange of new_ice_velocities (:1281-1320). */
int evp_pre_subcycle(evp_handle *handle, const evp_pre_fields *fields, const evp_pre_options *options);

/* velocity_solver_post_subcycle on the device + copy of the wanted results. */
int evp_post_subcycle(evp_handle *handle, const evp_post_fields *out);

/* The fields evp_pre_subcycle produced, for hosts that still need them (and for the parity tests):
 * same struct as evp_update_step consumes, every non-NULL pointer is filled. */
typedef struct {
    int *solveStress; int *solveVelocity; int *solveVelocityPrevious;
    double *icePressure;
    double *iceAreaVertex; double *totalMassVertex; double *totalMassVertexfVertex;
    double *airStressVertexU; double *airStressVertexV;
    double *surfaceTiltForceU; double *surfaceTiltForceV;
    double *oceanStressU; double *oceanStressV;
    double *uOceanVelocityVertex; double *vOceanVelocityVertex;
    double *uVelocityInitial; double *vVelocityInitial;
} evp_pre_out_fields;
int evp_fetch_pre(evp_handle *handle, const evp_pre_out_fields *out);


/* ======================================================================================================
 * Weak operators: config_strain_scheme / config_stress_divergence_scheme = 'weak'
 * (src/shared/mpas_seaice_velocity_solver_weak.F): line-integral strain on cells with ONE stress point per
 * cell, line-integral stress divergence on vertices.  The normal vectors come from the host
 * (seaice_normal_vectors, src/shared/mpas_seaice_mesh.F:703-2007, is not part of this library).
 * ====================================================================================================== */
typedef struct {
    int nEdges;
    double sphere_radius;                  /* mesh attribute; 0 is read as 1 like weak.F:208 */
    const int *edgesOnCell;                /* (maxEdges, nCells) */
    const int *verticesOnEdge;             /* (2, nEdges) */
    const int *edgesOnVertex;              /* (vertexDegree, nVertices) */
    const int *cellsOnEdge;                /* (2, nEdges) */
    const double *dvEdge;                  /* (nEdges) */
    const double *dcEdge;                  /* (nEdges) */
    const double *areaCell;                /* (nCells) */
    const double *areaTriangle;            /* (nVertices) */
    const double *normalVectorPolygon;     /* (2, maxEdges, nCells) */
    const double *normalVectorTriangle;    /* (2, vertexDegree, nVertices) */
    const double *latCellRotated;          /* (nCells)    tan() is taken on the host at this call */
    const double *latVertexRotated;        /* (nVertices) */
} evp_weak_mesh;

/* velocity_weak pool fields (src/Registry.xml, var_struct velocity_weak), all (nCells); any may be NULL */
typedef struct {
    double *stress11Weak;
    double *stress22Weak;
    double *stress12Weak;
    double *strain11Weak;
    double *strain22Weak;
    double *strain12Weak;
    double *replacementPressureWeak;
} evp_weak_fields;

int evp_set_weak_mesh(evp_handle *handle, const evp_weak_mesh *mesh);
/* Upload the carried weak stresses (the three stress pointers of `fields`; the rest is ignored). */
int evp_update_weak_state(evp_handle *handle, const evp_weak_fields *fields);
/* Blocking copy of the weak fields into host arrays. */
int evp_fetch_weak(evp_handle *handle, const evp_weak_fields *fields);

/* ---- host-side helper for non-Fortran hosts (plain CPU code, no device needed) ----
 * seaice_calc_variational_metric_terms (src/shared/mpas_seaice_velocity_solver_variational_shared.F:293-358):
 * tanLatVertexRotatedOverRadius[v] = tan(asin(zVertexRotated[v] / sphereRadius)) / sphereRadius, evaluated
 * element by element with the scalar libm, so the value of a vertex does not depend on where it sits in
 * the array (vectorised math libraries give position-dependent last bits, which would break the
 * bit-equality of owned results across rank counts).  A Fortran host computes this field itself. */
int evp_host_metric_terms(int nVertices, const double *zVertexRotated, double sphereRadius,
                          double *tanLatVertexRotatedOverRadius);

/* ---- instrumentation (bench.py, tests) ---- */
/* Time of the last evp_run_subcycles in milliseconds (CUDA events on the handle's stream). */
int evp_last_run_ms(evp_handle *handle, float *ms);
/* Number of kernel launches (graph kernel nodes) one evp_run_subcycles(nSubcycles) issues. */
int evp_launch_count(evp_handle *handle, int nSubcycles, int *count);
/* Average device time (ms) of the cell kernel, the vertex kernel and the rest (halo + special
 * boundaries) over nSubcycles un-graphed subcycles; CUDA events on the handle's stream. */
int evp_profile_passes(evp_handle *handle, int nSubcycles, float *cellMs, float *vertexMs, float *otherMs);
/* Raw stream / device pointers for profiling harnesses; not needed by the Fortran host. */
int evp_get_stream(evp_handle *handle, void **cudaStream);
/* Bytes of device memory owned by the handle. */
int evp_device_bytes(evp_handle *handle, unsigned long long *bytes);
/* Use (1, default) or bypass (0) CUDA-graph replay of the subcycle loop. */
int evp_set_use_graph(evp_handle *handle, int useGraph);

#ifdef __cplusplus
}
#endif
#endif /* EVP_B200_H */
