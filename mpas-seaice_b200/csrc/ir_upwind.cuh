// ir_upwind.cuh -- second part of the translation unit ir_kernels.cu (included at its end): two non-default corners of
// the transport row behind include/ir_b200.h.
//
//   ir_normal_vectors   seaice_normal_vectors, src/shared/mpas_seaice_mesh.F:703-846 (planar :858-1024, spherical
//                       :1038-1241 / :1393-1606): init-time geometry of the weak operators and of the upwind transport,
//                       for hosts that do not run the Fortran init.  One thread per cell / per vertex, arrays in the
//                       host layout (init-time work, like ir_init_geometry).
//   ir_run_upwind       seaice_run_advection_upwind, src/shared/mpas_seaice_advection_upwind.F:385-520, AS EXECUTED
//                       and with the tracer connectivity table as an argument: edge_from_vertex_velocity (:1403),
//                       prepare_tracers (:1890), prepare_none_parent_tracer (:819), upwind_tendencies (:1242), the
//                       update of run_advection_subvariable (:1215-1236), scale_tracers_back (:2141), thickness ->
//                       volume (:2063).  Rows [variable * nCategories + category][cell], one thread per cell or edge.
//                       The reference's cell loop that scatters flux_upwind / dvEdge into the edge array becomes an
//                       edge kernel (every edge of an owned cell gets the same value from either side), and the
//                       tendency a gather over the cell's edges in edgesOnCell order -- the same sums, bit for bit
//                       (oracle/upwind_oracle.c, tests/test_transport_options.py).
//
// Both are HBM-bound streaming / gather work with a handful of FP64 operations per element; no tensor work.

namespace {

constexpr double ICE_AREA_MINIMUM = 1.0e-11;   // iceAreaMinimum = seaicePuny, src/shared/mpas_seaice_constants.F:90
constexpr int UP_MAX_VARS = 16;

// ------------------------------------------------------------------------------------------------ normal vectors

struct Nrm {
    int nC, nV, nVS, nE, M, D, sphere, rotate, removeMetric;
    double radius;
    const int *nEdgesOnCell, *edgesOnCell, *verticesOnEdge, *cellsOnEdge, *edgesOnVertex, *interiorVertex;
    const double *xC, *yC, *zC, *xV, *yV, *zV, *xE, *yE, *zE;
    double *nvp, *nvt, *latC, *latV;
};

// seaice_grid_rotation_forward, mesh.F:2350
__device__ __forceinline__ void nrm_point(const Nrm &g, const double *x, const double *y, const double *z, int i, double *o)
{
    if (g.rotate) { o[0] = -z[i]; o[1] = y[i]; o[2] = x[i]; }
    else { o[0] = x[i]; o[1] = y[i]; o[2] = z[i]; }
}

// yRotationMatrix and zRotationMatrix of mesh.F:1118-1145 as the four numbers they hold
struct NrmRot { double cy, sy, cz, sz; };
__device__ __forceinline__ NrmRot nrm_rotation(const Nrm &g, double lat, double lon)
{
    NrmRot r;
    if (g.removeMetric) { r.cy = cos(lat); r.sy = sin(lat); r.cz = cos(-lon); r.sz = sin(-lon); }
    else { r.cy = 1.0; r.sy = 0.0; r.cz = 1.0; r.sz = 0.0; }
    return r;
}

// matmul(yRotationMatrix, matmul(zRotationMatrix, p)) with the zero entries of the matrices kept as multiplications
// (x * 0 adds a signed zero and nothing else; the sums run in matmul's order)
__device__ __forceinline__ void nrm_to_equator(const NrmRot &r, const double *p, double *o)
{
    const double t0 = r.cz * p[0] + -r.sz * p[1] + 0.0 * p[2];
    const double t1 = r.sz * p[0] + r.cz * p[1] + 0.0 * p[2];
    const double t2 = 0.0 * p[0] + 0.0 * p[1] + 1.0 * p[2];
    o[0] = r.cy * t0 + 0.0 * t1 + r.sy * t2;
    o[1] = 0.0 * t0 + 1.0 * t1 + 0.0 * t2;
    o[2] = -r.sy * t0 + 0.0 * t1 + r.cy * t2;
}

// the common tail of the two spherical routines (mesh.F:1190-1225, :1560-1595)
__device__ __forceinline__ void nrm_side(const double *side, const double *edgeEq, bool flip, double &n1, double &n2)
{
    double g0 = side[1] * edgeEq[2] - side[2] * edgeEq[1];
    double g1 = side[2] * edgeEq[0] - side[0] * edgeEq[2];
    double g2 = side[0] * edgeEq[1] - side[1] * edgeEq[0];
    if (flip) { g0 = -1.0 * g0; g1 = -1.0 * g1; g2 = -1.0 * g2; }
    const double norm = sqrt(g0 * g0 + g1 * g1 + g2 * g2);
    g0 = g0 / norm; g1 = g1 / norm; g2 = g2 / norm;
    double e0 = -edgeEq[1], e1 = edgeEq[0], e2 = 0.0;
    const double en = sqrt(e0 * e0 + e1 * e1);
    e0 = e0 / en; e1 = e1 / en; e2 = e2 / en;
    n1 = g0 * e0 + g1 * e1 + g2 * e2;
    const double clipped = fmax(fmin(n1, 1.0), -1.0);
    n2 = copysign(1.0, g2) * sqrt(1.0 - clipped * clipped);
}

// normal_vectors_planar_polygon (:858) / normal_vectors_spherical_polygon_metric (:1038): one thread per cell
__global__ void __launch_bounds__(128) k_normals_cells(Nrm g)
{
    const int c = (int)((size_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (c >= g.nC) return;
    const int n = g.nEdgesOnCell[c];
    double *out = g.nvp + (size_t)c * g.M * 2;
    if (!g.sphere) {
        for (int k = 0; k < n; k++) {
            const int e = g.edgesOnCell[(size_t)c * g.M + k] - 1;
            const int v1 = g.verticesOnEdge[2 * (size_t)e] - 1, v2 = g.verticesOnEdge[2 * (size_t)e + 1] - 1;
            double tx = g.xV[v2] - g.xV[v1], ty = g.yV[v2] - g.yV[v1];
            const double tmag = sqrt(tx * tx + ty * ty);
            tx = tx / tmag;
            ty = ty / tmag;
            const double nx = g.xE[e] - g.xC[c], ny = g.yE[e] - g.yC[c];
            if ((nx * ty - ny * tx) < 0.0) { tx = -tx; ty = -ty; }
            out[2 * k] = ty;
            out[2 * k + 1] = -tx;
        }
        return;
    }
    double centre[3];
    nrm_point(g, g.xC, g.yC, g.zC, c, centre);
    const double lon = atan2(centre[1], centre[0]);
    const double lat = asin(centre[2] / g.radius);
    const NrmRot r = nrm_rotation(g, lat, lon);
    for (int k = 0; k < n; k++) {
        const int e = g.edgesOnCell[(size_t)c * g.M + k] - 1;
        const int v1 = g.verticesOnEdge[2 * (size_t)e] - 1, v2 = g.verticesOnEdge[2 * (size_t)e + 1] - 1;
        double pe[3], p1[3], p2[3], qe[3], q1[3], q2[3];
        nrm_point(g, g.xE, g.yE, g.zE, e, pe);
        nrm_point(g, g.xV, g.yV, g.zV, v1, p1);
        nrm_point(g, g.xV, g.yV, g.zV, v2, p2);
        nrm_to_equator(r, pe, qe);
        nrm_to_equator(r, p1, q1);
        nrm_to_equator(r, p2, q2);
        const double side[3] = {q2[0] - q1[0], q2[1] - q1[1], q2[2] - q1[2]};
        nrm_side(side, qe, c + 1 == g.cellsOnEdge[2 * (size_t)e + 1], out[2 * k], out[2 * k + 1]);
    }
    if (g.latC) g.latC[c] = lat;
}

// normal_vectors_planar_triangle (:957, all vertices) / normal_vectors_spherical_triangle_metric (:1393, nVerticesSolve)
__global__ void __launch_bounds__(128) k_normals_vertices(Nrm g)
{
    const int v = (int)((size_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (v >= (g.sphere ? g.nVS : g.nV) || g.interiorVertex[v] != 1) return;
    double *out = g.nvt + (size_t)v * g.D * 2;
    if (!g.sphere) {
        for (int k = 0; k < g.D; k++) {
            const int e = g.edgesOnVertex[(size_t)v * g.D + k] - 1;
            const double dx = g.xE[e] - g.xV[v], dy = g.yE[e] - g.yV[v];
            out[2 * k] = dx / sqrt(dx * dx + dy * dy);
            out[2 * k + 1] = dy / sqrt(dx * dx + dy * dy);
        }
        return;
    }
    double pv[3];
    nrm_point(g, g.xV, g.yV, g.zV, v, pv);
    const double lon = atan2(pv[1], pv[0]);
    const double lat = asin(pv[2] / g.radius);
    const NrmRot r = nrm_rotation(g, lat, lon);
    for (int k = 0; k < g.D; k++) {
        const int e = g.edgesOnVertex[(size_t)v * g.D + k] - 1;
        const int c1 = g.cellsOnEdge[2 * (size_t)e] - 1, c2 = g.cellsOnEdge[2 * (size_t)e + 1] - 1;
        double pe[3], p1[3], p2[3], qe[3], q1[3], q2[3];
        nrm_point(g, g.xE, g.yE, g.zE, e, pe);
        nrm_point(g, g.xC, g.yC, g.zC, c1, p1);
        nrm_point(g, g.xC, g.yC, g.zC, c2, p2);
        nrm_to_equator(r, pe, qe);
        nrm_to_equator(r, p1, q1);
        nrm_to_equator(r, p2, q2);
        const double side[3] = {q2[0] - q1[0], q2[1] - q1[1], q2[2] - q1[2]};
        nrm_side(side, qe, v + 1 == g.verticesOnEdge[2 * (size_t)e], out[2 * k], out[2 * k + 1]);
    }
    if (g.latV) g.latV[v] = lat;
}

// ------------------------------------------------------------------------------------------------ upwind transport

struct UpTable {
    int n;
    int parent[UP_MAX_VARS], volumeLike[UP_MAX_VARS];
};

// edge_from_vertex_velocity (:1403): the normal of cellsOnEdge(1, e) for its copy of the edge
__global__ void __launch_bounds__(128) k_up_edge_velocity(Dev d, Upw p)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)d.nE) return;
    double out = 0.0;
    const int c1 = d.cellsOnEdge[2 * e];
    if (c1 >= 1 && c1 <= d.nC) {
        const size_t c = (size_t)c1 - 1;
        const int n = d.nEdgesOnCell[c];
        for (int j = 0; j < n; j++) {
            if (d.edgesOnCell[(size_t)j * d.nCp + c] != (int)e + 1) continue;
            double uE = 0.0, vE = 0.0;
            for (int i = 0; i < 2; i++) {
                const int iv = d.verticesOnEdge[(size_t)i * d.nEp + e] - 1;
                uE = uE + d.u[iv];
                vE = vE + d.v[iv];
            }
            uE = uE / 2.0;
            vE = vE / 2.0;
            out = uE * p.nve[(size_t)(2 * j) * d.nCp + c] + vE * p.nve[(size_t)(2 * j + 1) * d.nCp + c];
        }
    }
    p.edgeVel[e] = out;
}

// prepare_tracers (:1890): volume -> thickness where the area exceeds iceAreaMinimum; cells 1 .. nCells
__global__ void __launch_bounds__(128) k_up_prepare(Dev d, Upw p, UpTable tb)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (size_t)d.nC) return;
    for (int t = 1; t < tb.n; t++) {
        if (!tb.volumeLike[t]) continue;
        for (int k = 0; k < d.nK; k++) {
            const double area = p.oldv[(size_t)k * d.nCp + c];
            if (area > ICE_AREA_MINIMUM) p.oldv[(size_t)(t * d.nK + k) * d.nCp + c] = p.oldv[(size_t)(t * d.nK + k) * d.nCp + c] / area;
        }
    }
}

// prepare_none_parent_tracer (:819), the cell loop: 1 where the cell or an edge neighbour holds ice
__global__ void __launch_bounds__(128) k_up_none_cells(Dev d, Upw p, int t)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (size_t)d.nC) return;
    const int n = d.nEdgesOnCell[c];
    for (int k = 0; k < d.nK; k++) {
        const double *cOld = p.oldv + (size_t)(t * d.nK + k) * d.nCp;
        double one = 0.0;
        if (cOld[c] > ICE_AREA_MINIMUM) one = 1.0;
        for (int j = 0; j < n; j++)
            if (cOld[d.cellsOnCell[(size_t)j * d.nCp + c] - 1] > ICE_AREA_MINIMUM) { one = 1.0; break; }
        p.pOldNone[(size_t)k * d.nCp + c] = one;
        p.pNewNone[(size_t)k * d.nCp + c] = one;
    }
}

// ... and its edge loop: the edge velocity where either cell holds ice
__global__ void __launch_bounds__(128) k_up_none_edges(Dev d, Upw p, int t)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)d.nE) return;
    const size_t c1 = (size_t)d.cellsOnEdge[2 * e] - 1, c2 = (size_t)d.cellsOnEdge[2 * e + 1] - 1;
    for (int k = 0; k < d.nK; k++) {
        const double *cOld = p.oldv + (size_t)(t * d.nK + k) * d.nCp;
        p.pFluxNone[(size_t)k * d.nEp + e] = (cOld[c1] > ICE_AREA_MINIMUM || cOld[c2] > ICE_AREA_MINIMUM) ? p.edgeVel[e] : 0.0;
    }
}

// upwind_tendencies (:1242), the part that belongs to the edge: flux_upwind, and <child>EdgeFlux = flux_upwind / dvEdge.
// Only edges of owned cells are visited by the reference's loop (:1310).
__global__ void __launch_bounds__(128) k_up_edge_flux(Dev d, Upw p, int t, const double *__restrict__ pOld,
                                                      const double *__restrict__ pFlux, double parentMinimum)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)d.nE) return;
    const int i1 = d.cellsOnEdge[2 * e], i2 = d.cellsOnEdge[2 * e + 1];
    const bool live = p.interiorEdge[e] == 1 && (i1 <= d.nCS || i2 <= d.nCS);
    const size_t c1 = (size_t)i1 - 1, c2 = (size_t)i2 - 1;
    const double dv = p.dvEdge[e];
    for (int k = 0; k < d.nK; k++) {
        double fu = 0.0;
        if (live && (pOld[(size_t)k * d.nCp + c1] > parentMinimum || pOld[(size_t)k * d.nCp + c2] > parentMinimum)) {
            const double *cOld = p.oldv + (size_t)(t * d.nK + k) * d.nCp;
            const double pf = pFlux[(size_t)k * d.nEp + e];
            fu = dv * (fmax(0.0, pf) * cOld[c1] + fmin(0.0, pf) * cOld[c2]);
            p.eflux[(size_t)(t * d.nK + k) * d.nEp + e] = fu / dv;
        }
        p.fluxUp[(size_t)k * d.nEp + e] = fu;
    }
}

// the tendency of an owned cell (the gather side of upwind_tendencies) and the update of run_advection_subvariable
// (:1215-1236), which runs over nCells and rescales the old time level in place
__global__ void __launch_bounds__(128) k_up_update(Dev d, Upw p, int t, const double *__restrict__ pOld,
                                                   const double *__restrict__ pNew, double parentMinimum, double dt)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (size_t)d.nC) return;
    const bool owned = c < (size_t)d.nCS;
    const int n = owned ? d.nEdgesOnCell[c] : 0;
    const double invAreaCell1 = owned ? 1.0 / d.areaCell[c] : 0.0;
    for (int k = 0; k < d.nK; k++) {
        double tend = 0.0;
        for (int j = 0; j < n; j++) {
            const size_t e = (size_t)d.edgesOnCell[(size_t)j * d.nCp + c] - 1;
            const double edgeSignOnCell = (double)(-d.fluxSign[(size_t)j * d.nCp + c]);   // -1 for cellsOnEdge(1, e)
            tend = tend + edgeSignOnCell * p.fluxUp[(size_t)k * d.nEp + e] * invAreaCell1;
        }
        if (pNew[(size_t)k * d.nCp + c] > parentMinimum) {
            double *cOld = p.oldv + (size_t)(t * d.nK + k) * d.nCp;
            const double scaled = cOld[c] * pOld[(size_t)k * d.nCp + c];
            p.newv[(size_t)(t * d.nK + k) * d.nCp + c] = scaled + tend * dt;
            cOld[c] = scaled;
        }
    }
}

// finalize_tracers (:1989) on the owned cells: scale_tracers_back (:2141, last variable first), thickness -> volume (:2063)
__global__ void __launch_bounds__(128) k_up_finalize(Dev d, Upw p, UpTable tb)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (size_t)d.nCS) return;
    for (int t = tb.n - 1; t >= 1; t--)
        for (int k = 0; k < d.nK; k++) {
            const double pc = p.newv[(size_t)(tb.parent[t] * d.nK + k) * d.nCp + c];
            double *v = &p.newv[(size_t)(t * d.nK + k) * d.nCp + c];
            *v = pc > 0.0 ? *v / pc : 0.0;
        }
    for (int t = 1; t < tb.n; t++) {
        if (!tb.volumeLike[t]) continue;
        for (int k = 0; k < d.nK; k++) {
            const double area = p.newv[(size_t)k * d.nCp + c];
            if (area > ICE_AREA_MINIMUM) p.newv[(size_t)(t * d.nK + k) * d.nCp + c] = p.newv[(size_t)(t * d.nK + k) * d.nCp + c] * area;
        }
    }
}

void upwind_free_state(ir_handle *h)
{
    Upw &p = h->up;
    void *bufs[] = {p.oldv, p.newv, p.eflux, p.fluxUp, p.pOldNone, p.pNewNone, p.pFluxNone, p.edgeVel};
    for (void *b : bufs)
        if (b) cudaFree(b);
    p.oldv = p.newv = p.eflux = p.fluxUp = p.pOldNone = p.pNewNone = p.pFluxNone = p.edgeVel = nullptr;
    h->upVars = 0;
}

void upwind_free_all(ir_handle *h)
{
    upwind_free_state(h);
    Upw &p = h->up;
    if (p.interiorEdge) cudaFree(p.interiorEdge);
    if (p.dvEdge) cudaFree(p.dvEdge);
    if (p.nve) cudaFree(p.nve);
    p.interiorEdge = nullptr;
    p.dvEdge = p.nve = nullptr;
}

}  // namespace

extern "C" int ir_normal_vectors(const ir_normals_in *in, const ir_normals_out *out, int device)
{
    IR_REQUIRE(in != nullptr && out != nullptr, "NULL argument");
    IR_REQUIRE(in->nCells >= 0 && in->nVertices >= 0 && in->nEdges >= 0, "negative dimension");
    IR_REQUIRE(in->nVerticesSolve >= 0 && in->nVerticesSolve <= in->nVertices, "nVerticesSolve out of range");
    IR_REQUIRE(in->maxEdges >= 3 && in->maxEdges <= MAXM, "maxEdges must be 3..8");
    IR_REQUIRE(in->vertexDegree == 3 || in->vertexDegree == 4, "vertexDegree must be 3 or 4");
    IR_REQUIRE(in->nEdgesOnCell && in->edgesOnCell && in->verticesOnEdge && in->cellsOnEdge, "connectivity arrays must not be NULL");
    IR_REQUIRE(in->xCell && in->yCell && in->xVertex && in->yVertex && in->xEdge && in->yEdge, "coordinate arrays must not be NULL");
    IR_REQUIRE(!in->on_a_sphere || (in->zCell && in->zVertex && in->zEdge && in->sphere_radius > 0.0),
               "z coordinates and a positive sphere_radius are needed on a sphere");
    IR_REQUIRE(out->normalVectorPolygon != nullptr, "normalVectorPolygon must not be NULL");
    IR_REQUIRE(out->normalVectorTriangle == nullptr || (in->edgesOnVertex && in->interiorVertex),
               "edgesOnVertex and interiorVertex are needed for normalVectorTriangle");
    int dev = device;
    if (dev < 0) IR_CUDA(cudaGetDevice(&dev));
    IR_CUDA(cudaSetDevice(dev));
    const size_t nC1 = (size_t)in->nCells + 1, nE1 = (size_t)in->nEdges + 1, nV1 = (size_t)in->nVertices + 1;
    const int M = in->maxEdges, D = in->vertexDegree;
    const bool tri = out->normalVectorTriangle != nullptr;
    std::vector<void *> bufs;
    int rc = IR_OK;
    auto up = [&](const void *host, size_t bytes, void **devp) -> int {
        IR_CUDA(cudaMalloc(devp, bytes ? bytes : 1));
        bufs.push_back(*devp);
        if (host) IR_CUDA(cudaMemcpyAsync(*devp, host, bytes, cudaMemcpyHostToDevice, 0));
        else IR_CUDA(cudaMemsetAsync(*devp, 0, bytes ? bytes : 1, 0));
        return IR_OK;
    };
    auto cleanup = [&]() { for (void *b : bufs) cudaFree(b); };
#define TRYN(x) do { if ((rc = (x)) != IR_OK) { cleanup(); return rc; } } while (0)
    Nrm g;
    memset(&g, 0, sizeof g);
    g.nC = in->nCells; g.nV = in->nVertices; g.nVS = in->nVerticesSolve; g.nE = in->nEdges; g.M = M; g.D = D;
    g.sphere = in->on_a_sphere ? 1 : 0; g.rotate = in->rotate_cartesian_grid ? 1 : 0; g.removeMetric = in->remove_metric_terms ? 1 : 0;
    g.radius = in->sphere_radius;
    TRYN(up(in->nEdgesOnCell, nC1 * 4, (void **)&g.nEdgesOnCell));
    TRYN(up(in->edgesOnCell, nC1 * M * 4, (void **)&g.edgesOnCell));
    TRYN(up(in->verticesOnEdge, nE1 * 2 * 4, (void **)&g.verticesOnEdge));
    TRYN(up(in->cellsOnEdge, nE1 * 2 * 4, (void **)&g.cellsOnEdge));
    TRYN(up(in->xCell, nC1 * 8, (void **)&g.xC)); TRYN(up(in->yCell, nC1 * 8, (void **)&g.yC));
    TRYN(up(in->xVertex, nV1 * 8, (void **)&g.xV)); TRYN(up(in->yVertex, nV1 * 8, (void **)&g.yV));
    TRYN(up(in->xEdge, nE1 * 8, (void **)&g.xE)); TRYN(up(in->yEdge, nE1 * 8, (void **)&g.yE));
    if (g.sphere) {
        TRYN(up(in->zCell, nC1 * 8, (void **)&g.zC)); TRYN(up(in->zVertex, nV1 * 8, (void **)&g.zV)); TRYN(up(in->zEdge, nE1 * 8, (void **)&g.zE));
    }
    TRYN(up(nullptr, nC1 * M * 2 * 8, (void **)&g.nvp));
    if (out->latCellRotated) TRYN(up(nullptr, nC1 * 8, (void **)&g.latC));
    if (tri) {
        TRYN(up(in->edgesOnVertex, nV1 * D * 4, (void **)&g.edgesOnVertex));
        TRYN(up(in->interiorVertex, nV1 * 4, (void **)&g.interiorVertex));
        TRYN(up(nullptr, nV1 * D * 2 * 8, (void **)&g.nvt));
        if (out->latVertexRotated) TRYN(up(nullptr, nV1 * 8, (void **)&g.latV));
    }
    cudaStream_t s0 = 0;
    if (g.nC > 0) IR_LAUNCH((k_normals_cells), grid_for((size_t)g.nC, 128), 128, s0, g);
    if (tri && g.nV > 0) IR_LAUNCH((k_normals_vertices), grid_for((size_t)g.nV, 128), 128, s0, g);
    auto down = [&](void *host, const void *devp, size_t bytes) -> int {
        IR_CUDA(cudaMemcpyAsync(host, devp, bytes, cudaMemcpyDeviceToHost, 0));
        return IR_OK;
    };
    TRYN(down(out->normalVectorPolygon, g.nvp, nC1 * M * 2 * 8));
    if (out->latCellRotated) TRYN(down(out->latCellRotated, g.latC, nC1 * 8));
    if (tri) {
        TRYN(down(out->normalVectorTriangle, g.nvt, nV1 * D * 2 * 8));
        if (out->latVertexRotated) TRYN(down(out->latVertexRotated, g.latV, nV1 * 8));
    }
    cudaError_t ce = cudaStreamSynchronize(s0);
    if (ce == cudaSuccess) ce = cudaGetLastError();
    cleanup();
#undef TRYN
    if (ce != cudaSuccess) { set_error("ir_normal_vectors: %s", cudaGetErrorString(ce)); return IR_ERR_CUDA; }
    return IR_OK;
}

extern "C" int ir_set_upwind_mesh(ir_handle *h, const int *interiorEdge, const double *dvEdge, const double *normalVectorEdge)
{
    IR_REQUIRE(h != nullptr && interiorEdge != nullptr && dvEdge != nullptr && normalVectorEdge != nullptr, "NULL argument");
    IR_CUDA(cudaSetDevice(h->device));
    Dev &d = h->d;
    Upw &p = h->up;
    IR_CUDA(cudaStreamSynchronize(h->stream));
    if (!p.interiorEdge) IR_CUDA(cudaMalloc((void **)&p.interiorEdge, sizeof(int) * d.nEp));
    if (!p.dvEdge) IR_CUDA(cudaMalloc((void **)&p.dvEdge, sizeof(double) * d.nEp));
    if (!p.nve) IR_CUDA(cudaMalloc((void **)&p.nve, sizeof(double) * 2 * d.M * d.nCp));
    const size_t nE1 = (size_t)d.nE + 1, nC1 = (size_t)d.nC + 1;
    IR_CUDA(cudaMemsetAsync(p.interiorEdge, 0, sizeof(int) * d.nEp, h->stream));
    IR_CUDA(cudaMemsetAsync(p.dvEdge, 0, sizeof(double) * d.nEp, h->stream));
    IR_CUDA(cudaMemcpyAsync(p.interiorEdge, interiorEdge, nE1 * 4, cudaMemcpyHostToDevice, h->stream));
    IR_CUDA(cudaMemcpyAsync(p.dvEdge, dvEdge, nE1 * 8, cudaMemcpyHostToDevice, h->stream));
    IR_CUDA(cudaStreamSynchronize(h->stream));      // pageable sources
    int rc = upload_rows<double>(h, p.nve, normalVectorEdge, nC1, 2 * d.M, d.nCp);   // (nCells+1, maxEdges, 2) -> [2 j + i][cell]
    if (rc) return rc;
    h->upMesh = true;
    return IR_OK;
}

extern "C" int ir_run_upwind(ir_handle *h, int nVars, const ir_upwind_var *vars, const double *u, const double *v, double dt)
{
    IR_REQUIRE(h != nullptr && vars != nullptr && u != nullptr && v != nullptr, "NULL argument");
    if (!h->upMesh) { set_error("ir_run_upwind before ir_set_upwind_mesh"); return IR_ERR_STATE; }
    IR_REQUIRE(nVars >= 1 && nVars <= UP_MAX_VARS, "1 .. 16 variables");
    IR_REQUIRE(vars[0].parent == -1, "the first variable must have no parent ('none')");
    UpTable tb;
    memset(&tb, 0, sizeof tb);
    tb.n = nVars;
    for (int t = 0; t < nVars; t++) {
        IR_REQUIRE(vars[t].array != nullptr, "variable array is NULL");
        IR_REQUIRE(vars[t].parent >= -1 && vars[t].parent < t, "a parent must come before its children");
        IR_REQUIRE(!(vars[t].volumeLike && t == 0), "the first variable is the area; it cannot be volume-like");
        tb.parent[t] = vars[t].parent;
        tb.volumeLike[t] = vars[t].volumeLike ? 1 : 0;
    }
    IR_CUDA(cudaSetDevice(h->device));
    Dev &d = h->d;
    Upw &p = h->up;
    cudaStream_t s = h->stream;
    const int nK = d.nK;
    const size_t nC1 = (size_t)d.nC + 1;
    if (h->upVars != nVars) {
        IR_CUDA(cudaStreamSynchronize(s));
        upwind_free_state(h);
        const size_t cells = sizeof(double) * (size_t)nVars * nK * d.nCp, edges = sizeof(double) * (size_t)nVars * nK * d.nEp;
        IR_CUDA(cudaMalloc((void **)&p.oldv, cells));
        IR_CUDA(cudaMalloc((void **)&p.newv, cells));
        IR_CUDA(cudaMalloc((void **)&p.eflux, edges));
        IR_CUDA(cudaMalloc((void **)&p.fluxUp, sizeof(double) * nK * d.nEp));
        IR_CUDA(cudaMalloc((void **)&p.pOldNone, sizeof(double) * nK * d.nCp));
        IR_CUDA(cudaMalloc((void **)&p.pNewNone, sizeof(double) * nK * d.nCp));
        IR_CUDA(cudaMalloc((void **)&p.pFluxNone, sizeof(double) * nK * d.nEp));
        IR_CUDA(cudaMalloc((void **)&p.edgeVel, sizeof(double) * d.nEp));
        IR_CUDA(cudaMemsetAsync(p.oldv, 0, cells, s));
        IR_CUDA(cudaMemsetAsync(p.edgeVel, 0, sizeof(double) * d.nEp, s));
        h->upVars = nVars;
    }
    for (int t = 0; t < nVars; t++) pin_host(h, vars[t].array, nC1 * nK * sizeof(double));
    pin_host(h, u, ((size_t)d.nV + 1) * 8);
    pin_host(h, v, ((size_t)d.nV + 1) * 8);
    for (int t = 0; t < nVars; t++) {
        int rc = upload_rows<double>(h, p.oldv + (size_t)t * nK * d.nCp, vars[t].array, nC1, nK, d.nCp);
        if (rc) return rc;
    }
    IR_CUDA(cudaMemcpyAsync(d.u, u, ((size_t)d.nV + 1) * 8, cudaMemcpyHostToDevice, s));
    IR_CUDA(cudaMemcpyAsync(d.v, v, ((size_t)d.nV + 1) * 8, cudaMemcpyHostToDevice, s));
    IR_CUDA(cudaEventRecord(h->ev0, s));
    // initialize_timelevel_variables (:1633): the new time level and the edge fluxes start from zero
    IR_CUDA(cudaMemsetAsync(p.newv, 0, sizeof(double) * (size_t)nVars * nK * d.nCp, s));
    IR_CUDA(cudaMemsetAsync(p.eflux, 0, sizeof(double) * (size_t)nVars * nK * d.nEp, s));
    const unsigned gc = grid_for((size_t)d.nC, 128), ge = grid_for((size_t)d.nE, 128), gs = grid_for((size_t)d.nCS, 128);
    if (d.nE > 0) { IR_LAUNCH((k_up_edge_velocity), ge, 128, s, d, p); h->launches++; }
    if (d.nC > 0) { IR_LAUNCH((k_up_prepare), gc, 128, s, d, p, tb); h->launches++; }
    for (int t = 0; t < nVars; t++) {
        const double *pOld, *pNew, *pFlux;
        double parentMinimum = 0.0;
        if (vars[t].parent < 0) {
            if (d.nC > 0) { IR_LAUNCH((k_up_none_cells), gc, 128, s, d, p, t); h->launches++; }
            if (d.nE > 0) { IR_LAUNCH((k_up_none_edges), ge, 128, s, d, p, t); h->launches++; }
            pOld = p.pOldNone; pNew = p.pNewNone; pFlux = p.pFluxNone;
        } else {
            const size_t pr = (size_t)vars[t].parent * nK;
            pOld = p.oldv + pr * d.nCp; pNew = p.newv + pr * d.nCp; pFlux = p.eflux + pr * d.nEp;
            parentMinimum = vars[vars[t].parent].childMinimum;     // add_parent_tracer_minimums (:342)
        }
        if (d.nE > 0) { IR_LAUNCH((k_up_edge_flux), ge, 128, s, d, p, t, pOld, pFlux, parentMinimum); h->launches++; }
        if (d.nC > 0) { IR_LAUNCH((k_up_update), gc, 128, s, d, p, t, pOld, pNew, parentMinimum, dt); h->launches++; }
    }
    if (d.nCS > 0) { IR_LAUNCH((k_up_finalize), gs, 128, s, d, p, tb); h->launches++; }
    IR_CUDA(cudaEventRecord(h->ev1, s));
    IR_CUDA(cudaGetLastError());
    for (int t = 0; t < nVars; t++) {
        IR_LAUNCH((k_tracer_out), grid_for(nC1, 256), 256, s, d.stage, p.newv + (size_t)t * nK * d.nCp, nC1, nK, d.nCp);
        h->launches++;
        IR_CUDA(cudaMemcpyAsync(vars[t].array, d.stage, nC1 * nK * 8, cudaMemcpyDeviceToHost, s));
        IR_CUDA(cudaStreamSynchronize(s));
    }
    IR_CUDA(cudaEventElapsedTime(&h->lastMs, h->ev0, h->ev1));
    return IR_OK;
}

extern "C" int ir_fetch_upwind_fluxes(ir_handle *h, int var, double *edgeFlux, double *edgeVelocity)
{
    IR_REQUIRE(h != nullptr, "handle is NULL");
    if (h->upVars == 0) { set_error("ir_fetch_upwind_fluxes before ir_run_upwind"); return IR_ERR_STATE; }
    IR_REQUIRE(var >= 0 && var < h->upVars, "variable index out of range");
    IR_CUDA(cudaSetDevice(h->device));
    Dev &d = h->d;
    const size_t nE1 = (size_t)d.nE + 1;
    if (edgeFlux) {
        int rc = ensure_stage(h, nE1 * d.nK * sizeof(double));
        if (rc) return rc;
        IR_LAUNCH((k_tracer_out), grid_for(nE1, 256), 256, h->stream, d.stage, h->up.eflux + (size_t)var * d.nK * d.nEp, nE1, d.nK, d.nEp);
        h->launches++;
        IR_CUDA(cudaGetLastError());
        IR_CUDA(cudaMemcpyAsync(edgeFlux, d.stage, nE1 * d.nK * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    if (edgeVelocity) IR_CUDA(cudaMemcpyAsync(edgeVelocity, h->up.edgeVel, nE1 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    IR_CUDA(cudaStreamSynchronize(h->stream));
    return IR_OK;
}
