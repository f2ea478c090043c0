// evp_weak.cu -- the weak (line-integral) operators of config_strain_scheme / config_stress_divergence_scheme
// = 'weak' (reference: src/shared/mpas_seaice_velocity_solver_weak.F) on the device: one stress point per cell.
//   cells    seaice_strain_tensor_weak (weak.F:112-253) [+ seaice_stress_tensor_weak (:267-385)]
//   vertices seaice_stress_divergence_weak (:493-640) lives inside evp_vertex_kernel<..., WEAK> (evp_kernels.cu)
//            so that it stays fused with the drag coefficient and the 2x2 solve
//   mixed    interpolate_strains_weak_to_variational (src/shared/mpas_seaice_velocity_solver.F:2877-2972)
// Same rules as the variational kernels: --fmad=false, the reference's operation order, tan() taken on
// the host once (transcendentals are not bit-reproducible between math libraries).
#include <math.h>
#include <algorithm>
#include <vector>
#include "evp_internal.cuh"

namespace {

constexpr double kEccentricitySquared = 2.0 * 2.0;
constexpr double kPuny = 1.0e-11;
constexpr double kDampingRatioDenominator = 0.86;
constexpr double kDampingRatio = 5.5e-3;

struct WeakCellArgs {
    int nCells;
    size_t nCp;
    const uint8_t *__restrict__ nEdges, *__restrict__ solveStress;
    const int *__restrict__ voc;
    const int2 *__restrict__ edgeV;
    const double2 *__restrict__ np;
    const double *__restrict__ dv, *__restrict__ areaC, *__restrict__ tanC, *__restrict__ P;
    const double2 *__restrict__ uv;
    double2 *__restrict__ sigW;
    double *__restrict__ sigW12, *__restrict__ e11, *__restrict__ e22, *__restrict__ e12, *__restrict__ repP;
    double radius, dte, damping;
    int cr, withStress;
};

__global__ void __launch_bounds__(128) k_weak_cells(const WeakCellArgs a)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.nCells) return;
    const bool solve = a.solveStress[c] == 1;
    double e11 = 0.0, e22 = 0.0, e12 = 0.0;
    if (solve) {
        const int n = a.nEdges[c];
        double uC = 0.0, vC = 0.0;
        for (int k = 0; k < n; k++) {
            const size_t q = (size_t)k * a.nCp + c;
            const double2 w = a.uv[a.voc[q]];
            uC = uC + w.x;
            vC = vC + w.y;
            const int2 ev = a.edgeV[q];
            const double2 w1 = a.uv[ev.x], w2 = a.uv[ev.y];
            double uE = 0.0, vE = 0.0;
            uE = uE + w1.x; vE = vE + w1.y;
            uE = uE + w2.x; vE = vE + w2.y;
            uE = uE / 2.0;
            vE = vE / 2.0;
            const double2 nv = a.np[q];
            const double dv = a.dv[q];
            e11 = e11 + uE * nv.x * dv;
            e22 = e22 + vE * nv.y * dv;
            e12 = e12 + 0.5 * (uE * nv.y + vE * nv.x) * dv;
        }
        uC = uC / (double)n;
        vC = vC / (double)n;
        const double area = a.areaC[c], t = a.tanC[c];
        e11 = e11 / area;
        e22 = e22 / area;
        e12 = e12 / area;
        e11 = e11 - (vC * t) / a.radius;
        e12 = e12 + (uC * t * 0.5) / a.radius;
    }
    a.e11[c] = e11;
    a.e22[c] = e22;
    a.e12[c] = e12;
    if (!a.withStress || a.cr == EVP_CR_NONE) return;
    // seaice_stress_tensor_weak: cells that are not solved get zero stress (weak.F:318-322)
    double2 s = make_double2(0.0, 0.0);
    double s12 = 0.0;
    if (solve) {
        s = a.sigW[c];
        s12 = a.sigW12[c];
        if (a.cr == EVP_CR_LINEAR) {
            s.x = 1.0 * e11; s.y = 1.0 * e22; s12 = 1.0 * e12;
        } else {
            const double sd = e11 + e22, st = e11 - e22, ss = e12 * 2.0;
            double s1 = s.x + s.y, s2 = s.x - s.y;
            const double Delta = sqrt(sd * sd + (st * st + ss * ss) / kEccentricitySquared);
            double pc = a.P[c] / fmax(Delta, kPuny);
            a.repP[c] = pc * Delta;
            double den;
            if (a.cr == EVP_CR_EVP) {
                pc = (pc * a.dte) / (2.0 * a.damping);
                den = 1.0 + (0.5 * a.dte) / a.damping;
            } else {
                pc = (pc * 2.0 * kDampingRatio) / kDampingRatioDenominator;
                den = 1.0 + (2.0 * kDampingRatio) / kDampingRatioDenominator;
            }
            s1 = (s1 + pc * (sd - Delta)) / den;
            s2 = (s2 + (pc / kEccentricitySquared) * st) / den;
            s12 = (s12 + (pc / kEccentricitySquared) * ss * 0.5) / den;
            s.x = 0.5 * (s1 + s2);
            s.y = 0.5 * (s1 - s2);
        }
    }
    a.sigW[c] = s;
    a.sigW12[c] = s12;
}

// interpolate_strains_weak_to_variational, vertex loop (velocity_solver.F:2930-2955): owned vertices only
template <int D>
__global__ void __launch_bounds__(256) k_weak_strain_to_vertex(int nVerticesSolve, size_t nVp, const int *__restrict__ cov,
                                                                const double *__restrict__ areaC,
                                                                const double *__restrict__ e11, const double *__restrict__ e22,
                                                                const double *__restrict__ e12, double *__restrict__ v11,
                                                                double *__restrict__ v22, double *__restrict__ v12)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nVerticesSolve) return;
    double a11 = 0.0, a22 = 0.0, a12 = 0.0, denom = 0.0;
#pragma unroll
    for (int s = 0; s < D; s++) {
        const int c = cov[(size_t)s * nVp + v];
        if (c >= 0) {
            const double area = areaC[c];
            a11 = a11 + area * e11[c];
            a22 = a22 + area * e22[c];
            a12 = a12 + area * e12[c];
            denom = denom + area;
        }
    }
    v11[v] = a11 / denom;
    v22[v] = a22 / denom;
    v12[v] = a12 / denom;
}

// ... and its cell loop (:2957-2968): every stress point takes the value of its vertex
__global__ void __launch_bounds__(128) k_weak_vertex_to_var(int nCells, size_t nCp, int M, const uint8_t *__restrict__ nEdges,
                                                             const int *__restrict__ voc, const double *__restrict__ v11,
                                                             const double *__restrict__ v22, const double *__restrict__ v12,
                                                             double *__restrict__ e11, double *__restrict__ e22,
                                                             double *__restrict__ e12)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    const int n = nEdges[c];
    for (int k = 0; k < n; k++) {
        const size_t q = (size_t)k * nCp + c;
        const int v = voc[q];
        e11[q] = v11[v];
        e22[q] = v22[v];
        e12[q] = v12[v];
    }
}

template <typename T>
int upload_vec(evp_handle *h, T **dst, const std::vector<T> &src)
{
    int rc = evp_dev_alloc(h, (void **)dst, sizeof(T) * std::max<size_t>(src.size(), 1));
    if (rc) return rc;
    // on h->stream through the pinned bounce buffers (the handle's stream is non-blocking: a plain cudaMemcpy from
    // pageable memory is not ordered before the kernels launched on it); the source is consumed when this returns
    if (!src.empty()) {
        const bool pin = h->pinHost;
        h->pinHost = false;
        rc = evp_h2d(h, *dst, src.data(), sizeof(T) * src.size());
        h->pinHost = pin;
        if (rc) return rc;
    }
    return EVP_OK;
}

}  // namespace

extern "C" int evp_set_weak_mesh(evp_handle *h, const evp_weak_mesh *m)
{
    EVP_REQUIRE(h != nullptr && m != nullptr, "handle/mesh is NULL");
    EVP_REQUIRE(!h->haveWeak, "evp_set_weak_mesh was already called for this handle");
    EVP_REQUIRE(m->nEdges >= 0, "negative nEdges");
    EVP_REQUIRE(m->edgesOnCell && m->verticesOnEdge && m->edgesOnVertex && m->cellsOnEdge && m->dvEdge && m->dcEdge &&
                m->areaCell && m->areaTriangle && m->normalVectorPolygon && m->normalVectorTriangle &&
                m->latCellRotated && m->latVertexRotated, "weak mesh arrays must not be NULL");
    EVP_CUDA(cudaSetDevice(h->device));
    const size_t nC = h->nCells, nV = h->nVertices, nCp = h->nCp, nVp = h->nVp;
    const int Mh = h->Mh, Mk = h->M, D = h->D, nE = m->nEdges;
    // the host arrays are indexed through edges; resolve that indirection once, on the host
    std::vector<uint8_t> nEd(nCp, 0);
    EVP_CUDA(cudaMemcpy(nEd.data(), h->d.nEdges, nCp, cudaMemcpyDeviceToHost));
    std::vector<int2> edgeV((size_t)Mk * nCp, make_int2(0, 0));
    std::vector<double2> np((size_t)Mk * nCp, make_double2(0.0, 0.0));
    std::vector<double> dv((size_t)Mk * nCp, 0.0), areaC(nCp, 0.0), tanC(nCp, 0.0);
    for (size_t c = 0; c < nC; c++) {
        areaC[c] = m->areaCell[c];
        tanC[c] = tan(m->latCellRotated[c]);
        for (int k = 0; k < (int)nEd[c] && k < Mh; k++) {
            const int e = m->edgesOnCell[c * Mh + k];
            EVP_REQUIRE(e >= 1 && e <= nE, "edgesOnCell entry out of range");
            const int v1 = m->verticesOnEdge[2 * (size_t)(e - 1)], v2 = m->verticesOnEdge[2 * (size_t)(e - 1) + 1];
            EVP_REQUIRE(v1 >= 1 && v1 <= (int)nV && v2 >= 1 && v2 <= (int)nV, "verticesOnEdge entry out of range");
            const size_t q = (size_t)k * nCp + c;
            edgeV[q] = make_int2(v1 - 1, v2 - 1);
            np[q] = make_double2(m->normalVectorPolygon[2 * (c * Mh + k)], m->normalVectorPolygon[2 * (c * Mh + k) + 1]);
            dv[q] = m->dvEdge[e - 1];
        }
    }
    std::vector<int2> edgeC((size_t)D * nVp, make_int2(-1, -1));
    std::vector<double2> nt((size_t)D * nVp, make_double2(0.0, 0.0));
    std::vector<double> dc((size_t)D * nVp, 0.0), areaT(nVp, 0.0), tanV(nVp, 0.0);
    for (size_t v = 0; v < nV; v++) {
        areaT[v] = m->areaTriangle[v];
        tanV[v] = tan(m->latVertexRotated[v]);
        for (int s = 0; s < D; s++) {
            const int e = m->edgesOnVertex[v * D + s];
            if (e < 1 || e > nE) continue;                 // boundary vertex: never solved
            int c1 = m->cellsOnEdge[2 * (size_t)(e - 1)] - 1, c2 = m->cellsOnEdge[2 * (size_t)(e - 1) + 1] - 1;
            if (c1 < 0 || c1 >= (int)nC) c1 = -1;
            if (c2 < 0 || c2 >= (int)nC) c2 = -1;
            const size_t q = (size_t)s * nVp + v;
            edgeC[q] = make_int2(c1, c2);
            nt[q] = make_double2(m->normalVectorTriangle[2 * (v * D + s)], m->normalVectorTriangle[2 * (v * D + s) + 1]);
            dc[q] = m->dcEdge[e - 1];
        }
    }
    evp_dev &d = h->d;
    int rc;
    if ((rc = upload_vec(h, &d.wEdgeV, edgeV))) return rc;
    if ((rc = upload_vec(h, &d.wNp, np))) return rc;
    if ((rc = upload_vec(h, &d.wDv, dv))) return rc;
    if ((rc = upload_vec(h, &d.wAreaC, areaC))) return rc;
    if ((rc = upload_vec(h, &d.wTanC, tanC))) return rc;
    if ((rc = upload_vec(h, &d.wEdgeC, edgeC))) return rc;
    if ((rc = upload_vec(h, &d.wNt, nt))) return rc;
    if ((rc = upload_vec(h, &d.wDc, dc))) return rc;
    if ((rc = upload_vec(h, &d.wAreaT, areaT))) return rc;
    if ((rc = upload_vec(h, &d.wTanV, tanV))) return rc;
    struct { void **p; size_t bytes; } state[] = {
        {(void **)&d.sigW, sizeof(double2) * nCp}, {(void **)&d.sigW12, sizeof(double) * nCp},
        {(void **)&d.eW11, sizeof(double) * nCp}, {(void **)&d.eW22, sizeof(double) * nCp},
        {(void **)&d.eW12, sizeof(double) * nCp}, {(void **)&d.repPW, sizeof(double) * nCp},
        {(void **)&d.eV11, sizeof(double) * nVp}, {(void **)&d.eV22, sizeof(double) * nVp},
        {(void **)&d.eV12, sizeof(double) * nVp}};
    for (auto &x : state) {
        if ((rc = evp_dev_alloc(h, x.p, x.bytes))) return rc;
        EVP_CUDA(cudaMemsetAsync(*x.p, 0, x.bytes, h->stream));
    }
    EVP_CUDA(cudaStreamSynchronize(h->stream));
    d.wRadius = m->sphere_radius == 0.0 ? 1.0 : m->sphere_radius;
    h->haveWeak = true;
    return EVP_OK;
}

extern "C" int evp_update_weak_state(evp_handle *h, const evp_weak_fields *f)
{
    EVP_REQUIRE(h != nullptr && f != nullptr, "handle/fields is NULL");
    if (!h->haveWeak) { evp_set_error("evp_update_weak_state needs evp_set_weak_mesh first"); return EVP_ERR_STATE; }
    EVP_REQUIRE(f->stress11Weak && f->stress22Weak && f->stress12Weak, "the three weak stresses must not be NULL");
    EVP_CUDA(cudaSetDevice(h->device));
    evp_dev &d = h->d;
    const size_t nC = h->nCells, nCp = h->nCp;
    std::vector<double2> s(nCp, make_double2(0.0, 0.0));
    for (size_t c = 0; c < nC; c++) s[c] = make_double2(f->stress11Weak[c], f->stress22Weak[c]);
    // everything on the handle's stream (a non-blocking stream is not ordered after the legacy default stream), then one
    // synchronisation before the staging vector goes out of scope
    EVP_CUDA(cudaMemcpyAsync(d.sigW, s.data(), sizeof(double2) * nCp, cudaMemcpyHostToDevice, h->stream));
    if (nC) EVP_CUDA(cudaMemcpyAsync(d.sigW12, f->stress12Weak, sizeof(double) * nC, cudaMemcpyHostToDevice, h->stream));
    // init_subcycle_variables, weak branch (velocity_solver.F:2350-2365): strains start from zero
    EVP_CUDA(cudaMemsetAsync(d.eW11, 0, sizeof(double) * nCp, h->stream));
    EVP_CUDA(cudaMemsetAsync(d.eW22, 0, sizeof(double) * nCp, h->stream));
    EVP_CUDA(cudaMemsetAsync(d.eW12, 0, sizeof(double) * nCp, h->stream));
    EVP_CUDA(cudaStreamSynchronize(h->stream));
    return EVP_OK;
}

extern "C" int evp_fetch_weak(evp_handle *h, const evp_weak_fields *f)
{
    EVP_REQUIRE(h != nullptr && f != nullptr, "handle/fields is NULL");
    if (!h->haveWeak) { evp_set_error("evp_fetch_weak needs evp_set_weak_mesh first"); return EVP_ERR_STATE; }
    EVP_CUDA(cudaSetDevice(h->device));
    evp_dev &d = h->d;
    const size_t nC = h->nCells, nCp = h->nCp;
    EVP_CUDA(cudaStreamSynchronize(h->stream));
    if ((f->stress11Weak || f->stress22Weak) && nC) {
        std::vector<double2> s(nCp);
        EVP_CUDA(cudaMemcpy(s.data(), d.sigW, sizeof(double2) * nCp, cudaMemcpyDeviceToHost));
        for (size_t c = 0; c < nC; c++) {
            if (f->stress11Weak) f->stress11Weak[c] = s[c].x;
            if (f->stress22Weak) f->stress22Weak[c] = s[c].y;
        }
    }
    struct { double *host; const double *dev; } plain[] = {
        {f->stress12Weak, d.sigW12}, {f->strain11Weak, d.eW11}, {f->strain22Weak, d.eW22}, {f->strain12Weak, d.eW12},
        {f->replacementPressureWeak, d.repPW}};
    for (auto &p : plain)
        if (p.host && nC) EVP_CUDA(cudaMemcpy(p.host, p.dev, sizeof(double) * nC, cudaMemcpyDeviceToHost));
    return EVP_OK;
}

int evp_enqueue_weak_cell_pass(evp_handle *h, bool diag, cudaStream_t s)
{
    (void)diag;     // the weak fields are the reference's working arrays: written every subcycle
    if (h->nCells == 0) return EVP_OK;
    evp_dev &d = h->d;
    WeakCellArgs a;
    a.nCells = h->nCells; a.nCp = h->nCp;
    a.nEdges = d.nEdges; a.solveStress = d.solveStress; a.voc = d.voc; a.edgeV = d.wEdgeV; a.np = d.wNp; a.dv = d.wDv;
    a.areaC = d.wAreaC; a.tanC = d.wTanC; a.P = d.P; a.uv = d.uv;
    a.sigW = d.sigW; a.sigW12 = d.sigW12; a.e11 = d.eW11; a.e22 = d.eW22; a.e12 = d.eW12; a.repP = d.repPW;
    a.radius = d.wRadius; a.dte = h->opt.elasticTimeStep; a.damping = h->opt.dampingTimescale;
    a.cr = h->opt.constitutive_relation_type;
    a.withStress = h->opt.stress_divergence_scheme == EVP_SCHEME_WEAK;
    k_weak_cells<<<grid_for(h->nCells, 128), 128, 0, s>>>(a);
    EVP_CUDA(cudaGetLastError());
    return EVP_OK;
}

int evp_enqueue_weak_to_variational(evp_handle *h, cudaStream_t s)
{
    evp_dev &d = h->d;
    if (h->nVerticesSolve) {
        const unsigned grid = grid_for(h->nVerticesSolve, 256);
        if (h->D == 3)
            k_weak_strain_to_vertex<3><<<grid, 256, 0, s>>>(h->nVerticesSolve, h->nVp, d.cov, d.wAreaC, d.eW11, d.eW22, d.eW12,
                                                            d.eV11, d.eV22, d.eV12);
        else
            k_weak_strain_to_vertex<4><<<grid, 256, 0, s>>>(h->nVerticesSolve, h->nVp, d.cov, d.wAreaC, d.eW11, d.eW22, d.eW12,
                                                            d.eV11, d.eV22, d.eV12);
    }
    if (h->nCells)
        k_weak_vertex_to_var<<<grid_for(h->nCells, 128), 128, 0, s>>>(h->nCells, h->nCp, h->M, d.nEdges, d.voc, d.eV11, d.eV22,
                                                                      d.eV12, d.e11, d.e22, d.e12);
    EVP_CUDA(cudaGetLastError());
    return EVP_OK;
}
