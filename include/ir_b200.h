/* ir_b200.h -- C ABI of the B200 incremental-remapping (IR) transport of MPAS-Seaice: the consumer of the EVP
 * velocities.
 *
 * Replaces the work of one call of
 *     seaice_run_advection_incremental_remap        src/shared/mpas_seaice_advection_incremental_remap.F:2338-2730
 * on one block, i.e. incremental_remap_block (:2740-3400) with the volume <-> thickness conversions around it
 * (:2462-2480, :2680-2700).  Initialisation (seaice_init_advection_incremental_remap, :165-816) stays with the host
 * model: it fills the `incremental_remap` pool (Registry.xml var_struct incremental_remap) once, and those pool arrays
 * are what ir_create receives -- the same contract evp_b200.h has with the velocity solver's pools.
 *
 * STATUS: bit-identical to oracle/ir_oracle.c on a B200 and under host emulation (tests/test_ir_parity.py); timings
 * and ncu summaries under profiles/ir_r02_*.  One block per handle: the tracer halo update after the call
 * (seaice_update_tracer_halo, :2710) is the host's.
 *
 * Conventions as in evp_b200.h: host pointers, Fortran (column-major) layout passed with c_loc(), 1-based index
 * values, arrays dimensioned nCells / nEdges / nVertices carry MPAS's extra slot n+1.  Every entry point returns
 * IR_OK or an error code; ir_last_error_string() describes the last failure of the calling thread.
 */
#ifndef IR_B200_H
#define IR_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ir_handle ir_handle;

enum {
    IR_OK = 0,
    IR_ERR_ARGUMENT = 1,
    IR_ERR_CUDA = 2,
    IR_ERR_STATE = 3,
    IR_ERR_MESH = 4,                /* ir_init_geometry: orientation checks of the reference failed (:1296, :2010) */
    /* conditions the reference aborts on (MPAS_LOG_CRIT), reported after the step has run */
    IR_ERR_NEGATIVE_MASS_QP = 10,   /* negative mass at a quadrature point   incremental_remap.F:6895-6935 */
    IR_ERR_NEGATIVE_MASS = 11,      /* new mass below -puny**2               incremental_remap.F:7465-7480 */
    IR_ERR_PARALLEL_EDGES = 12,     /* degenerate basis in shift_vertices    incremental_remap.F:6415-6425 */
    IR_ERR_TOO_MANY_TRIANGLES = 13, /* more than nTriPerEdgeRemap departure triangles on an edge */
    /* the optional checks (ir_set_checks), reported like the abort conditions above */
    IR_ERR_CONSERVATION = 14,       /* check_tracer_conservation             incremental_remap.F:8126-8260 */
    IR_ERR_MONOTONICITY = 15        /* check_tracer_monotonicity             incremental_remap.F:8416-8760 */
};

/* Dimensions fixed by Registry.xml:59-78. */
#define IR_N_TRI_PER_EDGE 6
#define IR_MAX_CELLS_PER_EDGE_REMAP 6
#define IR_MAX_EDGES_PER_EDGE_REMAP 6
#define IR_MAX_VERTICES_PER_EDGE_REMAP 8

/* Mesh pool + incremental_remap pool arrays (incremental_remap_block's pointer list, :2860-2925). */
typedef struct ir_mesh_desc {
    int nCells, nCellsSolve, nVertices, nEdges, maxEdges, vertexDegree;
    int nCategories;
    int nQuadPoints;                 /* 3 or 6 (Registry.xml:59-62) */
    int on_a_sphere;                 /* mesh pool config on_a_sphere */
    int rotate_cartesian_grid;       /* config_rotate_cartesian_grid */
    /* mesh pool */
    const int *nEdgesOnCell;         /* (nCells+1) */
    const int *edgesOnCell;          /* (maxEdges, nCells+1) */
    const int *cellsOnCell;          /* (maxEdges, nCells+1) */
    const int *verticesOnCell;       /* (maxEdges, nCells+1) */
    const int *cellsOnEdge;          /* (2, nEdges+1) */
    const int *verticesOnEdge;       /* (2, nEdges+1) */
    const double *areaCell;          /* (nCells+1) */
    const double *dcEdge;            /* (nEdges+1) */
    const double *coeffs_reconstruct;/* (3, maxEdges, nCells+1)   mpas_init_reconstruct, :744-746 */
    /* incremental_remap pool */
    const double *transGlobalToCell; /* (3, 3, nCells); may be NULL on a plane */
    const double *xVertexOnCell, *yVertexOnCell;   /* (maxEdges, nCells+1) */
    const double *xVertexOnEdge, *yVertexOnEdge;   /* (8, nEdges+1) */
    const int *remapEdge;            /* (nEdges+1) */
    const int *cellsOnEdgeRemap;     /* (6, nEdges+1) */
    const int *edgesOnEdgeRemap;     /* (6, nEdges+1) */
    /* xAvgCell yAvgCell xxAvgCell xyAvgCell yyAvgCell xxxAvgCell xxyAvgCell xyyAvgCell yyyAvgCell
     * xxxxAvgCell xxxyAvgCell xxyyAvgCell xyyyAvgCell yyyyAvgCell, (nCells+1) each (:2051-2090) */
    const double *geomAvgCell[14];
} ir_mesh_desc;

/* One element of the reference's tracer linked list (incremental_remap_tracers.F:26-110), parents before children
 * (the order of seaice_add_tracers_to_linked_list after sorting, the first being the mass-like field). */
typedef struct ir_tracer_desc {
    int nLayers;      /* 1 for (nCategories, nCells) tracers, else the first dimension of (nLayers, nCategories, nCells) */
    int parent;       /* index of the parent tracer in the table, -1 for the mass-like field (iceAreaCategory) */
    int volumeLike;   /* 1 for iceVolumeCategory / snowVolumeCategory: volume in and out, thickness while transported */
    double *array;    /* (nLayers, nCategories, nCells+1), IN/OUT */
} ir_tracer_desc;

/* For hosts that do not run seaice_init_advection_incremental_remap (:165-816): its geometry part (:446-711) --
 * local frames, vertex coordinates in cell and edge frames, remap stencils, minimum edge length at the vertices,
 * geometric cell averages -- computed on the device from the mesh-file arrays into the arrays of the
 * incremental_remap pool.  A Fortran host keeps its own init and never calls this. */
typedef struct ir_geometry_in {
    int nCells, nCellsSolve, nVertices, nEdges, maxEdges, vertexDegree;
    int on_a_sphere, rotate_cartesian_grid;
    const int *nEdgesOnCell;     /* (nCells+1) */
    const int *edgesOnCell;      /* (maxEdges, nCells+1) */
    const int *verticesOnCell;   /* (maxEdges, nCells+1) */
    const int *cellsOnEdge;      /* (2, nEdges+1) */
    const int *verticesOnEdge;   /* (2, nEdges+1) */
    const int *edgesOnVertex;    /* (vertexDegree, nVertices+1) */
    const double *xCell, *yCell, *zCell;         /* (nCells+1) */
    const double *xVertex, *yVertex, *zVertex;   /* (nVertices+1) */
    const double *xEdge, *yEdge, *zEdge;         /* (nEdges+1) */
    const double *dcEdge, *dvEdge;               /* (nEdges+1) */
} ir_geometry_in;

typedef struct ir_geometry_out {
    double *transGlobalToCell;                   /* (3, 3, nCells); may be NULL on a plane */
    double *xVertexOnCell, *yVertexOnCell;       /* (maxEdges, nCells+1) */
    int *remapEdge;                              /* (nEdges+1) */
    int *cellsOnEdgeRemap, *edgesOnEdgeRemap;    /* (6, nEdges+1) */
    double *xVertexOnEdge, *yVertexOnEdge;       /* (8, nEdges+1) */
    double *minLengthEdgesOnVertex;              /* (nVertices+1) */
    double *geomAvgCell[14];                     /* as in ir_mesh_desc */
} ir_geometry_out;

int ir_init_geometry(const ir_geometry_in *in, const ir_geometry_out *out, int device);

/* Create the device copy of the mesh and geometry.  device < 0: the current CUDA device. */
int ir_create(ir_handle **out, const ir_mesh_desc *mesh, int device);

/* Declare the tracer hierarchy (sizes and parents; the array pointers are not read here).  Allocates the resident
 * device state.  May be called again when the set of active tracers changes. */
int ir_set_tracers(ir_handle *h, int nTracers, const ir_tracer_desc *tracers);

/* One transport step: uploads the tracers and the vertex velocities ((nVertices+1) each), runs the step, writes the
 * tracers back.  Returns one of the IR_ERR_* abort conditions if the reference would have aborted; the arrays then
 * hold whatever the step produced, as after the reference's abort write. */
int ir_run(ir_handle *h, int nTracers, const ir_tracer_desc *tracers, const double *uVelocity, const double *vVelocity,
           double dt);

/* config_conservation_check / config_monotonicity_check (Registry.xml, both default .false.; incremental_remap.F
 * :2574-2600, :2999-3015, :3259-3310).  Switched on, ir_run also computes
 *   conservation  the area-weighted sums of every mass * tracer product over the owned cells before and after the update
 *                 (sum_tracers :7998).  1: ir_run tests them itself (check_tracer_conservation :8126: a relative change
 *                 above 1e-11 returns IR_ERR_CONSERVATION) -- for a block that holds the whole mesh; 2: sums only, for
 *                 decomposed runs, where the host adds the ranks' sums first (mpas_dmpar_sum_real, :8150) and tests them;
 *   monotonicity  for every tracer with a parent, the range of the old values over a cell and its edge neighbours
 *                 (tracer_local_min_max :8268), widened by one more ring, against the new value
 *                 (check_tracer_monotonicity :8416): outside by more than 1e-11 * max(1, |bound|) returns
 *                 IR_ERR_MONOTONICITY.  Needs the halo layers the scheme itself requires (two; three on quadrilaterals, :829-852).
 * The transported fields are the same with and without the checks.  0 / 0 switches them off again. */
int ir_set_checks(ir_handle *h, int conservation, int monotonicity);

/* What the checks of the last ir_run found.  Indices as the reference logs them: tracer = position in the tracer table
 * (0-based), category / layer / cell 1-based (cell = local index).  The first violation in the reference's loop order. */
typedef struct ir_check_report {
    int conservationViolated;      /* 0 / 1 */
    int consTracer, consCategory, consLayer;
    double sumInit, sumFinal;      /* of that (tracer, category, layer) */
    int monotonicityViolated;      /* 0 none, 1 below the old minimum, 2 above the old maximum */
    int monoTracer, monoCategory, monoLayer, monoCell;
    double newValue, bound, tolerance;
} ir_check_report;
int ir_fetch_check_report(ir_handle *h, ir_check_report *out);

/* The sums of one tracer, (nLayers, nCategories) each in Fortran order: globalSumInit / globalSumFinal before the ranks
 * are added (incremental_remap_tracers.F:54-57). */
int ir_fetch_conservation_sums(ir_handle *h, int tracer, double *sumInit, double *sumFinal);

/* Diagnostics of the last step (any pointer may be NULL):
 * xTriangle / yTriangle (nQuadPoints, 6, nEdges) quadrature points, triangleArea (6, nEdges), iCellTriangle (6, nEdges),
 * maskEdge (nEdges), edgeFluxMass (nLayers of the mass field, nCategories, nEdges). */
int ir_fetch_diagnostics(ir_handle *h, double *xTriangle, double *yTriangle, double *triangleArea, int *iCellTriangle,
                         int *maskEdge, double *edgeFluxMass);

/* Work fields of the last step for one tracer, in the host layout of the tracer itself -- (nLayers, nCategories,
 * nCells+1), or (nLayers, nCategories, nEdges+1) for IR_FIELD_EDGE_FLUX: what the reference keeps in the
 * tracer_reconstruction / tracer_barycenter / tracer_products / tracer_edge_fluxes pools (center2D/3D, xGrad, yGrad,
 * xBarycenter, yBarycenter, massTracerProduct, edgeFlux; incremental_remap_tracers.F:40-110).  Debugging aid. */
enum {
    IR_FIELD_CENTER = 0,
    IR_FIELD_XGRAD = 1,
    IR_FIELD_YGRAD = 2,
    IR_FIELD_XBARYCENTER = 3,
    IR_FIELD_YBARYCENTER = 4,
    IR_FIELD_MASS_TRACER_PRODUCT = 5,   /* after the update */
    IR_FIELD_EDGE_FLUX = 6
};
int ir_fetch_tracer_field(ir_handle *h, int which, int tracer, double *out);

/* ---- seaice_normal_vectors (src/shared/mpas_seaice_mesh.F:703-846) on the device, for hosts that do not run the Fortran
 * init: the unit normals of the cell sides in the cell's local east / north frame (normalVectorPolygon, (2, maxEdges,
 * nCells+1)), those of the dual triangles' sides at interior vertices (normalVectorTriangle, (2, vertexDegree,
 * nVertices+1)) and the rotated latitudes.  Planar :858-1024, spherical :1038-1241 / :1393-1606.  Callers in the
 * reference: seaice_init_velocity_solver_weak (weak.F:84-96, remove_metric_terms = 1) -- the arrays evp_set_weak_mesh of
 * evp_b200.h takes -- and seaice_init_advection_upwind (advection_upwind.F:122-126, polygons only, remove_metric_terms
 * = 0) -- the normalVectorEdge of ir_set_upwind_mesh below.  The device's sin / cos / asin / atan2 are accurate to 1-2
 * ulp, not correctly rounded, so the result agrees with a host libm to round-off, not bit for bit; the second component
 * is sign(n3) * sqrt(1 - n1**2), which loses half the digits where a side points due east or west (n1 = +-1) -- in the
 * reference as here. */
typedef struct ir_normals_in {
    int nCells, nVertices, nVerticesSolve, nEdges, maxEdges, vertexDegree;
    int on_a_sphere, rotate_cartesian_grid, remove_metric_terms;
    double sphere_radius;
    const int *nEdgesOnCell;     /* (nCells+1) */
    const int *edgesOnCell;      /* (maxEdges, nCells+1) */
    const int *verticesOnEdge;   /* (2, nEdges+1) */
    const int *cellsOnEdge;      /* (2, nEdges+1) */
    const int *edgesOnVertex;    /* (vertexDegree, nVertices+1); triangles only */
    const int *interiorVertex;   /* (nVertices+1); triangles only */
    const double *xCell, *yCell, *zCell;         /* (nCells+1); z may be NULL on a plane */
    const double *xVertex, *yVertex, *zVertex;   /* (nVertices+1) */
    const double *xEdge, *yEdge, *zEdge;         /* (nEdges+1) */
} ir_normals_in;

typedef struct ir_normals_out {
    double *normalVectorPolygon;   /* (2, maxEdges, nCells+1) */
    double *normalVectorTriangle;  /* (2, vertexDegree, nVertices+1); NULL: seaice_normal_vectors_polygon alone */
    double *latCellRotated;        /* (nCells+1), may be NULL */
    double *latVertexRotated;      /* (nVertices+1), may be NULL */
} ir_normals_out;

int ir_normal_vectors(const ir_normals_in *in, const ir_normals_out *out, int device);

/* ---- config_advection_type = 'upwind': seaice_run_advection_upwind (src/shared/mpas_seaice_advection_upwind.F:385-520)
 * on one block, restated AS EXECUTED with the tracer connectivity table (tracerConnectivities, :37-56, filled at
 * :145-170) as an argument.  The handle is the one of ir_create (its mesh arrays are shared; the incremental_remap pool
 * arrays are not read by this path).  Left to the host: the tracer halo exchange between prepare_advection and the
 * variable loop (halo_exchange_advection :1786 -- with the values unchanged by prepare_tracers' volume -> thickness
 * conversion only if the host exchanges thicknesses, so decomposed hosts call with halos already current), and the time-
 * level shift of the pool (:2032).
 *
 * ir_set_upwind_mesh: interiorEdge (nEdges+1) of the boundary pool (mesh.F:567), dvEdge (nEdges+1), normalVectorEdge
 * (2, maxEdges, nCells+1) of the velocity_solver pool (advection_upwind.F:122).  Once, before the first step. */
int ir_set_upwind_mesh(ir_handle *h, const int *interiorEdge, const double *dvEdge, const double *normalVectorEdge);

typedef struct ir_upwind_var {
    int parent;            /* index of the parent in the table, -1 for 'none' (the first variable: the ice area) */
    int volumeLike;        /* 1: divided by variable 0 before and multiplied by its new value after the step where it
                            * exceeds iceAreaMinimum (config_convert_volume_to_thickness, :57, :1949, :2063) */
    double childMinimum;   /* childTracerMinimum (:235); the threshold its children use as parentTracerMinimum */
    double *array;         /* (nCategories, nCells+1): time level 1 in, the new time level out */
} ir_upwind_var;

/* One step IN PLACE on the variables' arrays, in table order (a parent before its children).  uVelocity / vVelocity
 * (nVertices+1).  Owned cells hold the new values; halo cells hold old * parent as after the reference's update loop
 * over nCells (:1136, :1215) and want the host's next halo exchange. */
int ir_run_upwind(ir_handle *h, int nVars, const ir_upwind_var *vars, const double *uVelocity, const double *vVelocity,
                  double dt);

/* <variable>EdgeFlux (nCategories, nEdges+1) of the tracer_edge_fluxes pool and the edge velocity (nEdges+1) of the
 * last ir_run_upwind; either pointer may be NULL. */
int ir_fetch_upwind_fluxes(ir_handle *h, int var, double *edgeFlux, double *edgeVelocity);

/* With IR_B200_PIN_HOST=1 in the environment at ir_create, ir_run page-locks the tracer and velocity arrays the first
 * time it sees them (cudaHostRegister) so that the per-step copies run at the full PCIe rate.  Call this before the
 * host frees or reallocates such an array (e.g. mpas_pool_destroy_pool); ir_destroy does it too.  No-op otherwise. */
int ir_release_host_memory(ir_handle *h);

int ir_last_run_ms(ir_handle *h, float *ms);       /* device time of the last ir_run / ir_run_upwind, kernels only */
/* ... of the last ir_run by kernel, ms[5]: prepare, reconstruct, triangles, fluxes, update (each with the check kernels
 * launched next to it, if any): CUDA events on the handle's stream */
int ir_last_kernel_ms(ir_handle *h, float *ms);
int ir_launch_count(ir_handle *h, long long *n);   /* kernels launched by this handle so far */
int ir_destroy(ir_handle *h);
const char *ir_last_error_string(void);

#ifdef __cplusplus
}
#endif
#endif
