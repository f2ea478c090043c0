// evp_abi.cu -- the C-ABI of libevp_b200.so: handle lifecycle, host<->device marshalling between the
// Registry (Fortran column-major, AoS-in-cell) layout and the SoA device layout, CUDA-graph replay
// of the subcycle loop.  Mirrors the lifecycle of module seaice_mesh_pool
// (reference: src/shared/mpas_seaice_mesh_pool.F:76-281).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <new>
#include "evp_internal.cuh"

static thread_local char g_err[512] = "no error";

void evp_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *evp_last_error_string(void) { return g_err; }

// ------------------------------------------------------------------------------------------
// layout kernels: one thread per cell / vertex so that the SoA side is always coalesced
// ------------------------------------------------------------------------------------------
namespace {

constexpr size_t kPinChunk = 32u << 20;   // pinned bounce buffers (2 x 32 MiB)

// Host block of `count` cells, each (Mh) or (Mh, Mh) doubles with the first index fastest, to the device
// layout: 1-D -> row-SoA dst[(i*stride + c0 + c)*ncomp + comp]; 2-D (basis arrays) -> tiled,
// dst[evp_tix(j*Mk + i, c0 + c, Mk*Mk)*ncomp + comp].
__global__ void k_rows_in(const double *__restrict__ src, double *__restrict__ dst, int Mh, int Mk, int dims,
                          size_t count, size_t c0, size_t stride, int ncomp, int comp)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= count) return;
    const int nj = dims == 2 ? Mh : 1;
    const double *s = src + (size_t)Mh * nj * c;
    for (int j = 0; j < nj; j++)
        for (int i = 0; i < Mh; i++)
            dst[(dims == 2 ? evp_tix(j * Mk + i, c0 + c, Mk * Mk) : (size_t)i * stride + c0 + c) * ncomp + comp] = s[j * Mh + i];
}
__global__ void k_rows_out(double *__restrict__ raw, const double *__restrict__ soa, int Mh, int Mk, int dims,
                           size_t count, size_t c0, size_t stride, int ncomp, int comp)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= count) return;
    const int nj = dims == 2 ? Mh : 1;
    double *s = raw + (size_t)Mh * nj * c;
    for (int j = 0; j < nj; j++)
        for (int i = 0; i < Mh; i++)
            s[j * Mh + i] = soa[(dims == 2 ? evp_tix(j * Mk + i, c0 + c, Mk * Mk) : (size_t)i * stride + c0 + c) * ncomp + comp];
}
__global__ void k_voc_in(const int *__restrict__ src, int *__restrict__ dst, const int *__restrict__ nEdgesRaw,
                         int Mh, size_t count, size_t c0, size_t stride, int nVertices)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= count) return;
    const int n = nEdgesRaw[c];
    for (int r = 0; r < Mh; r++) {
        int v = src[(size_t)Mh * c + r] - 1;
        if (r >= n || v < 0 || v >= nVertices) v = 0;   // slots beyond nEdgesOnCell are never used; keep them in range
        dst[(size_t)r * stride + c0 + c] = v;
    }
}
__global__ void k_u8_in(const int *__restrict__ src, uint8_t *__restrict__ dst, size_t count, int isMask, int cap)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int x = src[i];
    dst[i] = isMask ? (uint8_t)(x == 1) : (uint8_t)(x < 0 ? 0 : (x > cap ? cap : x));
}
__global__ void k_pair_in(const double *__restrict__ a, const double *__restrict__ b, double2 *__restrict__ dst, size_t count)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double2 w = dst[i];
    if (a) w.x = a[i];
    if (b) w.y = b[i];
    dst[i] = w;
}
__global__ void k_pair_out(double *__restrict__ a, double *__restrict__ b, const double2 *__restrict__ src, size_t count)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double2 w = src[i];
    if (a) a[i] = w.x;
    if (b) b[i] = w.y;
}
__global__ void k_gidx(const int *__restrict__ cov, const int *__restrict__ cvav, const uint8_t *__restrict__ nEdges,
                       int *__restrict__ gidx, int *__restrict__ cov0, int D, size_t nV, size_t nVp, int nCells, size_t nCp)
{
    const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nV) return;
    for (int s = 0; s < D; s++) {
        const int c = cov[(size_t)D * v + s];
        const int j = cvav[(size_t)D * v + s];
        int g = -1;
        // the reference's inner loop runs 1..nEdgesOnCell(iCell); the junk cell has 0 edges
        if (c >= 1 && c <= nCells && j >= 1 && j <= (int)nEdges[c - 1]) g = (int)((size_t)(j - 1) * nCp + (size_t)(c - 1));
        gidx[(size_t)s * nVp + v] = g;
        cov0[(size_t)s * nVp + v] = (c >= 1 && c <= nCells) ? c - 1 : -1;     // pre-/post-subcycle interpolation
    }
}

// increasing-i order of the three band entries of gradient vertex j (0-based) in a cell with n vertices
__device__ __forceinline__ void band_rows(int j, int n, int &i0, int &i1, int &i2)
{
    i0 = j - 1; i1 = j; i2 = j + 1;
    if (j == 0) { i0 = 0; i1 = 1; i2 = n - 1; }
    else if (j == n - 1) { i0 = 0; i1 = n - 2; i2 = n - 1; }
}
// flag[0] |= 1 when some basisGradient(i,j,c) outside the cyclic band i in {j-1,j,j+1} is non-zero
__global__ void k_band_check(const double2 *__restrict__ G, const uint8_t *__restrict__ nEdges, int M, size_t nC,
                             size_t nCp, int *flag)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC) return;
    const int n = nEdges[c];
    bool bad = n < 3 && n > 0;
    for (int j = 0; j < n; j++) {
        int i0, i1, i2;
        band_rows(j, n, i0, i1, i2);
        for (int i = 0; i < n; i++) {
            if (i == i0 || i == i1 || i == i2) continue;
            const double2 g = G[evp_tix(j * M + i, c, M * M)];
            bad |= (g.x != 0.0) | (g.y != 0.0);
        }
    }
    if (bad) atomicOr(flag, 1);
}
__global__ void k_band_pack(const double2 *__restrict__ G, double2 *__restrict__ Gb, const uint8_t *__restrict__ nEdges,
                            int M, size_t nC, size_t nCp)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC) return;
    const int n = nEdges[c];
    for (int j = 0; j < M; j++) {
        double2 g[3] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
        if (j < n) {
            int i[3];
            band_rows(j, n, i[0], i[1], i[2]);
            for (int k = 0; k < 3; k++) g[k] = G[evp_tix(j * M + i[k], c, M * M)];
        }
        for (int k = 0; k < 3; k++) Gb[evp_tix(k * M + j, c, 3 * M)] = g[k];
    }
}
__global__ void k_band_unpack(const double2 *__restrict__ Gb, double2 *__restrict__ G, const uint8_t *__restrict__ nEdges,
                              int M, size_t nC, size_t nCp)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC) return;
    const int n = nEdges[c];
    for (int j = 0; j < n; j++) {
        int i[3];
        band_rows(j, n, i[0], i[1], i[2]);
        for (int k = 0; k < 3; k++) G[evp_tix(j * M + i[k], c, M * M)] = Gb[evp_tix(k * M + j, c, 3 * M)];
    }
}


struct NoPin {   // one-shot transfers (static data, basis read-back) never page-lock caller memory
    evp_handle *h; bool saved;
    explicit NoPin(evp_handle *h_) : h(h_), saved(h_->pinHost) { h->pinHost = false; }
    ~NoPin() { h->pinHost = saved; }
};

}  // namespace

int evp_dev_alloc(evp_handle *h, void **p, size_t bytes)
{
    if (bytes == 0) bytes = 256;
    EVP_CUDA(cudaMalloc(p, bytes));
    h->allocs.push_back(*p);
    h->devBytes += bytes;
    return EVP_OK;
}

void evp_dev_free(evp_handle *h, void *p, size_t bytes)
{
    if (!p) return;
    for (size_t i = 0; i < h->allocs.size(); i++)
        if (h->allocs[i] == p) { h->allocs.erase(h->allocs.begin() + i); break; }
    cudaFree(p);
    h->devBytes -= std::min<unsigned long long>(h->devBytes, bytes);
}

static void invalidate_graph(evp_handle *h);

int evp_basis_begin(evp_handle *h)
{
    invalidate_graph(h);
    const size_t bytes = sizeof(double2) * h->M * h->M * h->nCp;
    if (!h->d.G) {
        int rc = evp_dev_alloc(h, (void **)&h->d.G, bytes);
        if (rc) return rc;
    }
    EVP_CUDA(cudaMemsetAsync(h->d.G, 0, bytes, h->stream));
    if (h->d.Gb) {
        EVP_CUDA(cudaStreamSynchronize(h->stream));
        evp_dev_free(h, h->d.Gb, sizeof(double2) * 3 * h->M * h->nCp);
        h->d.Gb = nullptr;
    }
    h->haveBasis = false;
    return EVP_OK;
}

// Wachspress gradients are band-sparse: keep only the band (saves 16*(M*M - 3*M) B per cell and per
// subcycle of HBM traffic).  Any other pattern (PWL: dense, pwl.F:259-274) keeps the dense array.
int evp_basis_finalize(evp_handle *h)
{
    h->haveBasis = true;
    if (h->nCells == 0 || getenv("EVP_B200_DENSE_GRADIENT")) return EVP_OK;
    const size_t nC = h->nCells;
    int *flag = (int *)h->d.stage, bad = 0;
    EVP_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), h->stream));
    k_band_check<<<grid_for(nC, 128), 128, 0, h->stream>>>(h->d.G, h->d.nEdges, h->M, nC, h->nCp, flag);
    EVP_CUDA(cudaGetLastError());
    EVP_CUDA(cudaMemcpyAsync(&bad, flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    EVP_CUDA(cudaStreamSynchronize(h->stream));
    if (bad) return EVP_OK;
    const size_t bytes = sizeof(double2) * 3 * h->M * h->nCp;
    int rc = evp_dev_alloc(h, (void **)&h->d.Gb, bytes);
    if (rc) return rc;
    EVP_CUDA(cudaMemsetAsync(h->d.Gb, 0, bytes, h->stream));
    k_band_pack<<<grid_for(nC, 128), 128, 0, h->stream>>>(h->d.G, h->d.Gb, h->d.nEdges, h->M, nC, h->nCp);
    EVP_CUDA(cudaGetLastError());
    EVP_CUDA(cudaStreamSynchronize(h->stream));
    evp_dev_free(h, h->d.G, sizeof(double2) * h->M * h->M * h->nCp);
    h->d.G = nullptr;
    return EVP_OK;
}

static bool host_is_pinned(evp_handle *h, const void *p, size_t bytes)
{
    if (!h->pinHost || bytes < (1u << 20)) return false;
    // both ends: a range that only overlaps an older (possibly stale) registration must not be treated as pinned
    cudaPointerAttributes a0, a1;
    if (cudaPointerGetAttributes(&a0, p) == cudaSuccess && a0.type == cudaMemoryTypeHost &&
        cudaPointerGetAttributes(&a1, (const char *)p + bytes - 1) == cudaSuccess && a1.type == cudaMemoryTypeHost)
        return true;
    cudaGetLastError();
    if (cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterDefault) == cudaSuccess) {
        h->pinned.push_back(const_cast<void *>(p));
        return true;
    }
    cudaGetLastError();   // e.g. a page shared with an already registered range: use the bounce path
    return false;
}

// host -> device on h->stream; pageable sources go through two pinned bounce buffers
int evp_h2d(evp_handle *h, void *dst, const void *src, size_t bytes)
{
    if (bytes == 0) return EVP_OK;
    if (host_is_pinned(h, src, bytes)) {
        if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream) == cudaSuccess) return EVP_OK;
        cudaGetLastError();      // e.g. a stale registration of freed memory under this address: take the bounce path
    }
    size_t off = 0;
    while (off < bytes) {
        const int k = h->pinNext;
        const size_t n = std::min(kPinChunk, bytes - off);
        EVP_CUDA(cudaEventSynchronize(h->pinEv[k]));     // the previous copy out of this bounce buffer is done
        memcpy(h->pinStage[k], (const char *)src + off, n);
        EVP_CUDA(cudaMemcpyAsync((char *)dst + off, h->pinStage[k], n, cudaMemcpyHostToDevice, h->stream));
        EVP_CUDA(cudaEventRecord(h->pinEv[k], h->stream));
        off += n;
        h->pinNext ^= 1;
    }
    return EVP_OK;
}

// device -> host; blocking for pageable destinations, stream-ordered for pinned ones
int evp_d2h(evp_handle *h, void *dst, const void *src, size_t bytes)
{
    if (bytes == 0) return EVP_OK;
    if (host_is_pinned(h, dst, bytes)) {
        if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream) == cudaSuccess) return EVP_OK;
        cudaGetLastError();      // e.g. a stale registration of freed memory under this address: take the bounce path
    }
    size_t off = 0, pendOff[2] = {0, 0}, pendN[2] = {0, 0};
    bool pend[2] = {false, false};
    EVP_CUDA(cudaEventSynchronize(h->pinEv[0]));
    EVP_CUDA(cudaEventSynchronize(h->pinEv[1]));
    int k = 0;
    while (off < bytes) {
        const size_t n = std::min(kPinChunk, bytes - off);
        if (pend[k]) {
            EVP_CUDA(cudaEventSynchronize(h->pinEv[k]));
            memcpy((char *)dst + pendOff[k], h->pinStage[k], pendN[k]);
        }
        EVP_CUDA(cudaMemcpyAsync(h->pinStage[k], (const char *)src + off, n, cudaMemcpyDeviceToHost, h->stream));
        EVP_CUDA(cudaEventRecord(h->pinEv[k], h->stream));
        pend[k] = true; pendOff[k] = off; pendN[k] = n;
        off += n;
        k ^= 1;
    }
    for (int i = 0; i < 2; i++) {
        const int kk = (k + i) & 1;
        if (pend[kk]) {
            EVP_CUDA(cudaEventSynchronize(h->pinEv[kk]));
            memcpy((char *)dst + pendOff[kk], h->pinStage[kk], pendN[kk]);
        }
    }
    return EVP_OK;
}

// upload a (Mh[,Mh], nCells) host array into SoA rows, chunked through the device staging area
int evp_upload_rows(evp_handle *h, const double *host, double *dst, int dims, int ncomp, int comp)
{
    const size_t nC = (size_t)h->nCells;
    const size_t perCell = (size_t)h->Mh * (dims == 2 ? h->Mh : 1) * sizeof(double);
    const size_t chunkCells = std::max<size_t>(1, std::min(nC, h->d.stageBytes / perCell));
    for (size_t c0 = 0; c0 < nC; c0 += chunkCells) {
        const size_t cnt = std::min(chunkCells, nC - c0);
        int rc = evp_h2d(h, h->d.stage, (const char *)host + c0 * perCell, cnt * perCell);
        if (rc) return rc;
        k_rows_in<<<grid_for(cnt, 128), 128, 0, h->stream>>>((const double *)h->d.stage, dst, h->Mh, h->M, dims, cnt,
                                                              c0, h->nCp, ncomp, comp);
        EVP_CUDA(cudaGetLastError());
        // the staging area is reused by the next chunk: its copy is stream-ordered behind this kernel
    }
    return EVP_OK;
}

int evp_download_rows(evp_handle *h, double *host, const double *soa, int dims, int ncomp, int comp)
{
    const size_t nC = (size_t)h->nCells;
    const size_t perCell = (size_t)h->Mh * (dims == 2 ? h->Mh : 1) * sizeof(double);
    const size_t chunkCells = std::max<size_t>(1, std::min(nC, h->d.stageBytes / perCell));
    for (size_t c0 = 0; c0 < nC; c0 += chunkCells) {
        const size_t cnt = std::min(chunkCells, nC - c0);
        k_rows_out<<<grid_for(cnt, 128), 128, 0, h->stream>>>((double *)h->d.stage, soa, h->Mh, h->M, dims, cnt, c0,
                                                               h->nCp, ncomp, comp);
        EVP_CUDA(cudaGetLastError());
        int rc = evp_d2h(h, (char *)host + c0 * perCell, h->d.stage, cnt * perCell);
        if (rc) return rc;
        // pinned destinations: the copy is asynchronous and the staging area is about to be reused
        EVP_CUDA(cudaStreamSynchronize(h->stream));
    }
    return EVP_OK;
}

static void invalidate_graph(evp_handle *h)
{
    if (h->graphExec) {
        cudaGraphExecDestroy(h->graphExec);
        h->graphExec = nullptr;
    }
    h->graphN = -1;
}

static int check_options(const evp_options *o)
{
    EVP_REQUIRE(o != nullptr, "options is NULL");
    EVP_REQUIRE(o->constitutive_relation_type >= EVP_CR_EVP && o->constitutive_relation_type <= EVP_CR_NONE,
                "constitutive_relation_type must be 1..4");
    EVP_REQUIRE(o->ocean_stress_type == EVP_OCEAN_QUADRATIC || o->ocean_stress_type == EVP_OCEAN_LINEAR,
                "ocean_stress_type must be 1 or 2");
    EVP_REQUIRE(o->strain_scheme >= 0 && o->strain_scheme <= EVP_SCHEME_WEAK && o->stress_divergence_scheme >= 0 &&
                o->stress_divergence_scheme <= EVP_SCHEME_WEAK, "strain / stress divergence scheme must be 0..2");
    // velocity_solver.F:195-198: "variational strain scheme with weak stress divergence scheme" is rejected
    EVP_REQUIRE(!(o->strain_scheme != EVP_SCHEME_WEAK && o->stress_divergence_scheme == EVP_SCHEME_WEAK),
                "variational strain with weak stress divergence is not a valid combination");
    EVP_REQUIRE(!(o->strain_scheme == EVP_SCHEME_WEAK && o->average_variational_strain),
                "average_variational_strain applies to the variational strain scheme only");
    if (o->constitutive_relation_type == EVP_CR_EVP)
        EVP_REQUIRE(o->elasticTimeStep > 0.0 && o->dampingTimescale > 0.0,
                    "elasticTimeStep and dampingTimescale must be > 0");
    if (o->constitutive_relation_type == EVP_CR_EVP_REVISED)
        EVP_REQUIRE(o->dynamicsTimeStep > 0.0, "dynamicsTimeStep must be > 0");
    return EVP_OK;
}

extern "C" int evp_set_options(evp_handle *h, const evp_options *o)
{
    EVP_REQUIRE(h != nullptr, "handle is NULL");
    int rc = check_options(o);
    if (rc) return rc;
    if (o->use_special_boundaries_velocity && !h->haveSB) {
        evp_set_error("special boundaries were not described at evp_create");
        return EVP_ERR_ARGUMENT;
    }
    const int dev = h->device;
    evp_options next = *o;
    next.device = dev;
    // the Fortran shim refreshes the options every dynamics step (config_dt may change in coupled runs): keep the
    // instantiated graph when nothing changed -- kernel arguments such as the time steps are baked into its nodes
    const evp_options &p = h->opt;      // field by field: the struct has padding
    const bool same = next.constitutive_relation_type == p.constitutive_relation_type &&
                      next.ocean_stress_type == p.ocean_stress_type && next.use_ocean_stress == p.use_ocean_stress &&
                      next.use_special_boundaries_velocity == p.use_special_boundaries_velocity &&
                      next.flags == p.flags && next.average_variational_strain == p.average_variational_strain &&
                      next.strain_scheme == p.strain_scheme && next.stress_divergence_scheme == p.stress_divergence_scheme &&
                      next.elasticTimeStep == p.elasticTimeStep && next.dynamicsTimeStep == p.dynamicsTimeStep &&
                      next.dampingTimescale == p.dampingTimescale &&
                      next.numericalInertiaCoefficient == p.numericalInertiaCoefficient;
    h->opt = next;
    h->pinHost = (o->flags & EVP_FLAG_PIN_HOST) != 0;
    if (same) return EVP_OK;
    invalidate_graph(h);
    if (h->haveStep) {      // keep the boundary bit of the velocity mask in step with the options
        EVP_CUDA(cudaSetDevice(h->device));
        int rc2 = evp_halo_mark_masks(h);
        if (rc2) return rc2;
        EVP_CUDA(cudaStreamSynchronize(h->stream));
    }
    return EVP_OK;
}

extern "C" int evp_create(evp_handle **out, const evp_mesh_desc *m, const evp_options *o)
{
    EVP_REQUIRE(out != nullptr && m != nullptr, "handle/mesh is NULL");
    *out = nullptr;
    int rc = check_options(o);
    if (rc) return rc;
    EVP_REQUIRE(m->nCells >= 0 && m->nVertices >= 0, "negative dimension");
    EVP_REQUIRE(m->nVerticesSolve >= 0 && m->nVerticesSolve <= m->nVertices, "nVerticesSolve out of range");
    EVP_REQUIRE(m->maxEdges >= 3 && m->maxEdges <= 8, "maxEdges must be 3..8");
    EVP_REQUIRE(m->vertexDegree == 3 || m->vertexDegree == 4, "vertexDegree must be 3 or 4");
    EVP_REQUIRE(m->nEdgesOnCell && m->verticesOnCell && m->cellsOnVertex, "connectivity arrays must not be NULL");
    // a pure weak configuration (pkgVariational inactive in the host) has no velocity_variational fields at all
    const bool pureWeak = o->stress_divergence_scheme == EVP_SCHEME_WEAK;
    if (!pureWeak) {
        EVP_REQUIRE(m->cellVerticesAtVertex, "cellVerticesAtVertex must not be NULL");
        EVP_REQUIRE(m->tanLatVertexRotatedOverRadius && m->variationalDenominator,
                    "tanLatVertexRotatedOverRadius / variationalDenominator must not be NULL");
    }
    const bool anyBasis = m->basisGradientU || m->basisGradientV || m->basisIntegralsU || m->basisIntegralsV ||
                          m->basisIntegralsMetric;
    const bool allBasis = m->basisGradientU && m->basisGradientV && m->basisIntegralsU && m->basisIntegralsV &&
                          m->basisIntegralsMetric;
    EVP_REQUIRE(!anyBasis || allBasis, "basis arrays must be all given or all NULL");
    const bool haveSB = m->vertexBoundaryType && m->vertexBoundarySourceLocal;
    if (o->use_special_boundaries_velocity)
        EVP_REQUIRE(haveSB, "special boundaries need vertexBoundaryType / vertexBoundarySourceLocal");

    int dev = o->device;
    if (dev < 0) EVP_CUDA(cudaGetDevice(&dev));
    EVP_CUDA(cudaSetDevice(dev));

    evp_handle *h = new (std::nothrow) evp_handle();
    EVP_REQUIRE(h != nullptr, "out of host memory");
    h->device = dev;
    h->opt = *o;
    h->opt.device = dev;
    h->pinHost = (o->flags & EVP_FLAG_PIN_HOST) != 0;
    h->nCells = m->nCells; h->nCellsSolve = m->nCellsSolve;
    h->nVertices = m->nVertices; h->nVerticesSolve = m->nVerticesSolve;
    h->Mh = m->maxEdges; h->D = m->vertexDegree;
    // the kernels are instantiated for 4, 6 and 8 slots per cell
    h->M = h->Mh <= 4 ? 4 : (h->Mh <= 6 ? 6 : 8);
    const int Mh = h->Mh, Mk = h->M, D = h->D;
    h->nCp = ((size_t)h->nCells + 63) / 64 * 64;
    h->nVp = ((size_t)h->nVertices + 63) / 64 * 64;
    if (h->nCp == 0) h->nCp = 64;
    if (h->nVp == 0) h->nVp = 64;
    const size_t nCp = h->nCp, nVp = h->nVp, nC = h->nCells, nV = h->nVertices;

#define FAIL_IF(x) do { int rc_ = (x); if (rc_) { evp_destroy(h); return rc_; } } while (0)
#define CUDA_FAIL(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
        evp_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); evp_destroy(h); return EVP_ERR_CUDA; } } while (0)

    if ((size_t)Mk * nCp >= (size_t)0x7fffffff) {
        evp_set_error("mesh too large for 32-bit gather indices");
        evp_destroy(h);
        return EVP_ERR_ARGUMENT;
    }
    CUDA_FAIL(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CUDA_FAIL(cudaEventCreate(&h->ev0));
    CUDA_FAIL(cudaEventCreate(&h->ev1));
    CUDA_FAIL(cudaEventCreateWithFlags(&h->pinEv[0], cudaEventDisableTiming));
    CUDA_FAIL(cudaEventCreateWithFlags(&h->pinEv[1], cudaEventDisableTiming));
    CUDA_FAIL(cudaMallocHost(&h->pinStage[0], kPinChunk));
    CUDA_FAIL(cudaMallocHost(&h->pinStage[1], kPinChunk));

    evp_dev &d = h->d;
    FAIL_IF(evp_dev_alloc(h, (void **)&d.nEdges, nCp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.voc, sizeof(int) * Mk * nCp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.Suv, sizeof(double2) * Mk * Mk * nCp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.Sm, sizeof(double) * Mk * Mk * nCp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.tanLat, sizeof(double) * nVp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.gidx, sizeof(int) * D * nVp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.cov, sizeof(int) * D * nVp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.solveVelPrev, nVp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.solveStress, nCp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.tileWork, nCp / EVP_TILE + 1));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.tileList, sizeof(int) * (nCp / EVP_TILE + 1)));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.tileCount, sizeof(int)));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.vblockWork, nVp / 256 + 2));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.vblockList, sizeof(int) * (nVp / 256 + 2)));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.vblockCount, sizeof(int)));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.solveVel, nVp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.P, sizeof(double) * nCp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.uv, sizeof(double2) * nVp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.sig, sizeof(double2) * Mk * nCp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.sig12, sizeof(double) * Mk * nCp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.contrib, sizeof(double2) * Mk * nCp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.massf, sizeof(double2) * nVp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.air, sizeof(double2) * nVp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.tilt, sizeof(double2) * nVp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.ocnStress, sizeof(double2) * nVp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.ocnVel, sizeof(double2) * nVp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.areaDen, sizeof(double2) * nVp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.uvInit, sizeof(double2) * nVp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.e11, sizeof(double) * Mk * nCp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.e22, sizeof(double) * Mk * nCp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.e12, sizeof(double) * Mk * nCp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.repP, sizeof(double) * Mk * nCp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.sdiv, sizeof(double2) * nVp));
    FAIL_IF(evp_dev_alloc(h, (void **)&d.ocoef, sizeof(double) * nVp));
    {   // staging: one dynamics step of raw inputs (3 stresses, pressure, masks, 15 vertex fields), >= 64 MiB
        size_t need = 3 * (size_t)Mh * nC * 8 + nC * 8 + (nC + nV) * 4 + 17 * nV * 8 + 64 * 256;
        need = std::max(need, (size_t)(64u << 20));
        d.stageBytes = need;
        FAIL_IF(evp_dev_alloc(h, &d.stage, need));
    }
    // zero everything a kernel may read before the host wrote it
    CUDA_FAIL(cudaMemsetAsync(d.voc, 0, sizeof(int) * Mk * nCp, h->stream));
    CUDA_FAIL(cudaMemsetAsync(d.nEdges, 0, nCp, h->stream));
    CUDA_FAIL(cudaMemsetAsync(d.solveVelPrev, 0, nVp, h->stream));
    CUDA_FAIL(cudaMemsetAsync(d.solveVel, 0, nVp, h->stream));
    CUDA_FAIL(cudaMemsetAsync(d.solveStress, 0, nCp, h->stream));
    CUDA_FAIL(cudaMemsetAsync(d.tileWork, 1, nCp / EVP_TILE + 1, h->stream));
    CUDA_FAIL(cudaMemsetAsync(d.uv, 0, sizeof(double2) * nVp, h->stream));
    CUDA_FAIL(cudaMemsetAsync(d.Suv, 0, sizeof(double2) * Mk * Mk * nCp, h->stream));
    CUDA_FAIL(cudaMemsetAsync(d.Sm, 0, sizeof(double) * Mk * Mk * nCp, h->stream));
    CUDA_FAIL(cudaMemsetAsync(d.contrib, 0, sizeof(double2) * Mk * nCp, h->stream));
    CUDA_FAIL(cudaMemsetAsync(d.sig, 0, sizeof(double2) * Mk * nCp, h->stream));
    CUDA_FAIL(cudaMemsetAsync(d.sig12, 0, sizeof(double) * Mk * nCp, h->stream));
    CUDA_FAIL(cudaMemsetAsync(d.uvInit, 0, sizeof(double2) * nVp, h->stream));
    CUDA_FAIL(cudaMemsetAsync(d.areaDen, 0, sizeof(double2) * nVp, h->stream));
    CUDA_FAIL(cudaMemsetAsync(d.tanLat, 0, sizeof(double) * nVp, h->stream));

    // ---- static uploads -------------------------------------------------------------------
    NoPin noPin(h);
    Stage st{(char *)d.stage, d.stageBytes, 0};
    if (nC > 0) {
        int *rawN = (int *)st.take(nC * 4);
        FAIL_IF(evp_h2d(h, rawN, m->nEdgesOnCell, nC * 4));
        k_u8_in<<<grid_for(nC, 256), 256, 0, h->stream>>>(rawN, d.nEdges, nC, 0, Mh);
        const size_t chunk = std::max<size_t>(1, std::min(nC, (st.cap - st.off - 4096) / ((size_t)Mh * 4)));
        int *rawV = (int *)st.take(chunk * Mh * 4);
        for (size_t c0 = 0; c0 < nC; c0 += chunk) {
            const size_t cnt = std::min(chunk, nC - c0);
            FAIL_IF(evp_h2d(h, rawV, m->verticesOnCell + c0 * Mh, cnt * Mh * 4));
            k_voc_in<<<grid_for(cnt, 128), 128, 0, h->stream>>>(rawV, d.voc, rawN + c0, Mh, cnt, c0, nCp, (int)nV);
        }
        CUDA_FAIL(cudaGetLastError());
        CUDA_FAIL(cudaStreamSynchronize(h->stream));
    }
    if (nV > 0) {
        st.off = 0;
        const size_t chunk = std::max<size_t>(1, std::min(nV, (st.cap - 8192) / ((size_t)D * 8)));
        int *rawCov = (int *)st.take(chunk * D * 4);
        int *rawCv = (int *)st.take(chunk * D * 4);
        for (size_t v0 = 0; v0 < nV; v0 += chunk) {
            const size_t cnt = std::min(chunk, nV - v0);
            FAIL_IF(evp_h2d(h, rawCov, m->cellsOnVertex + v0 * D, cnt * D * 4));
            if (m->cellVerticesAtVertex) FAIL_IF(evp_h2d(h, rawCv, m->cellVerticesAtVertex + v0 * D, cnt * D * 4));
            else CUDA_FAIL(cudaMemsetAsync(rawCv, 0, cnt * D * 4, h->stream));      // slot 0 = "not in that cell"
            k_gidx<<<grid_for(cnt, 256), 256, 0, h->stream>>>(rawCov, rawCv, d.nEdges, d.gidx + v0, d.cov + v0, D, cnt, nVp,
                                                               (int)nC, nCp);
        }
        CUDA_FAIL(cudaGetLastError());
        if (m->tanLatVertexRotatedOverRadius) FAIL_IF(evp_h2d(h, d.tanLat, m->tanLatVertexRotatedOverRadius, nV * 8));
        if (m->variationalDenominator) {
            st.off = 0;
            double *rawD = (double *)st.take(nV * 8);
            FAIL_IF(evp_h2d(h, rawD, m->variationalDenominator, nV * 8));
            k_pair_in<<<grid_for(nV, 256), 256, 0, h->stream>>>(nullptr, rawD, d.areaDen, nV);   // .y = denominator
        }
        CUDA_FAIL(cudaGetLastError());
        CUDA_FAIL(cudaStreamSynchronize(h->stream));
        h->metric = false;
        for (size_t i = 0; m->tanLatVertexRotatedOverRadius && i < nV; i++)
            if (m->tanLatVertexRotatedOverRadius[i] != 0.0) { h->metric = true; break; }
    }
    if (allBasis && nC > 0) {
        FAIL_IF(evp_basis_begin(h));
        FAIL_IF(evp_upload_rows(h, m->basisGradientU, (double *)d.G, 2, 2, 0));
        FAIL_IF(evp_upload_rows(h, m->basisGradientV, (double *)d.G, 2, 2, 1));
        FAIL_IF(evp_upload_rows(h, m->basisIntegralsU, (double *)d.Suv, 2, 2, 0));
        FAIL_IF(evp_upload_rows(h, m->basisIntegralsV, (double *)d.Suv, 2, 2, 1));
        FAIL_IF(evp_upload_rows(h, m->basisIntegralsMetric, d.Sm, 2, 1, 0));
        CUDA_FAIL(cudaStreamSynchronize(h->stream));
        FAIL_IF(evp_basis_finalize(h));
    }
    if (nC == 0) h->haveBasis = true;      // an empty block (a rank without cells) has nothing to precompute

    // ---- special boundaries: resolve the sequential in-place loop (special_boundaries.F:301-324) ----
    if (haveSB && nV > 0) {
        std::vector<int> curSrc(nV), dst, src;
        std::vector<double> curSign(nV, 1.0), sign;
        for (size_t i = 0; i < nV; i++) curSrc[i] = (int)i;
        for (size_t i = 0; i < nV; i++) {
            const int t = m->vertexBoundaryType[i];
            if (t == EVP_VB_PERIODIC || t == EVP_VB_REVERSE) {
                const int s = m->vertexBoundarySourceLocal[i] - 1;
                if (s < 0 || s >= (int)nV) {
                    evp_set_error("vertexBoundarySourceLocal(%zu) = %d out of range", i + 1, s + 1);
                    evp_destroy(h);
                    return EVP_ERR_ARGUMENT;
                }
                // vertices before i already hold their updated value, later ones their old value
                curSrc[i] = curSrc[s];
                curSign[i] = (t == EVP_VB_REVERSE) ? -curSign[s] : curSign[s];
            } else if (t == EVP_VB_ZERO) {
                curSign[i] = 0.0;
            }
            if (t != EVP_VB_NONE) { dst.push_back((int)i); src.push_back(curSrc[i]); sign.push_back(curSign[i]); }
        }
        // a source that is itself a later boundary vertex must be read BEFORE it is overwritten:
        // the gather/scatter kernel pair does exactly that, but the resolution above assumed
        // "later = old value", which is what the sequential loop sees.  Nothing else to do.
        d.nSB = (int)dst.size();
        h->haveSB = true;
        if (d.nSB) {
            FAIL_IF(evp_dev_alloc(h, (void **)&d.sbDst, sizeof(int) * d.nSB));
            FAIL_IF(evp_dev_alloc(h, (void **)&d.sbSrc, sizeof(int) * d.nSB));
            FAIL_IF(evp_dev_alloc(h, (void **)&d.sbSign, sizeof(double) * d.nSB));
            FAIL_IF(evp_dev_alloc(h, (void **)&d.sbTmp, sizeof(double2) * d.nSB));
            // through the bounce buffers on h->stream (a non-blocking stream is not ordered behind a plain cudaMemcpy)
            FAIL_IF(evp_h2d(h, d.sbDst, dst.data(), sizeof(int) * d.nSB));
            FAIL_IF(evp_h2d(h, d.sbSrc, src.data(), sizeof(int) * d.nSB));
            FAIL_IF(evp_h2d(h, d.sbSign, sign.data(), sizeof(double) * d.nSB));
        }
    }
    CUDA_FAIL(cudaStreamSynchronize(h->stream));
#undef FAIL_IF
#undef CUDA_FAIL
    *out = h;
    return EVP_OK;
}

extern "C" int evp_set_masks(evp_handle *h, const int *solveStress, const int *solveVelocity)
{
    EVP_REQUIRE(h != nullptr && solveStress && solveVelocity, "NULL argument");
    EVP_CUDA(cudaSetDevice(h->device));
    evp_dev &d = h->d;
    const size_t nC = h->nCells, nV = h->nVertices;
    Stage st{(char *)d.stage, d.stageBytes, 0};
    int *rawMs = (int *)st.take(nC * 4), *rawMv = (int *)st.take(nV * 4);
    int rc;
    EVP_CUDA(cudaStreamSynchronize(h->stream));
    if ((rc = evp_h2d(h, rawMs, solveStress, nC * 4))) return rc;
    if ((rc = evp_h2d(h, rawMv, solveVelocity, nV * 4))) return rc;
    if (nC) k_u8_in<<<grid_for(nC, 256), 256, 0, h->stream>>>(rawMs, d.solveStress, nC, 1, 1);
    if (nV) k_u8_in<<<grid_for(nV, 256), 256, 0, h->stream>>>(rawMv, d.solveVel, nV, 1, 1);
    EVP_CUDA(cudaGetLastError());
    if ((rc = evp_halo_mark_masks(h))) return rc;
    if ((rc = evp_refresh_tile_flags(h, h->stream))) return rc;
    EVP_CUDA(cudaStreamSynchronize(h->stream));
    return EVP_OK;
}

extern "C" int evp_update_step(evp_handle *h, const evp_step_fields *f)
{
    EVP_REQUIRE(h != nullptr && f != nullptr, "handle/fields is NULL");
    EVP_REQUIRE(f->solveStress && f->solveVelocity && f->icePressure && f->uVelocity && f->vVelocity,
                "mesh-pool step fields must not be NULL");
    // the variational stresses do not exist in a pure weak configuration (evp_update_weak_state carries its state)
    const bool needVarStress = h->opt.stress_divergence_scheme != EVP_SCHEME_WEAK;
    EVP_REQUIRE((f->stress11 && f->stress22 && f->stress12) || (!needVarStress && !f->stress11 && !f->stress22 && !f->stress12),
                "stress11/22/12 must not be NULL (all three may be NULL only with the weak stress divergence scheme)");
    EVP_REQUIRE(f->totalMassVertex && f->totalMassVertexfVertex && f->iceAreaVertex && f->airStressVertexU &&
                f->airStressVertexV && f->surfaceTiltForceU && f->surfaceTiltForceV && f->oceanStressU &&
                f->oceanStressV && f->uOceanVelocityVertex && f->vOceanVelocityVertex,
                "vertex forcing fields must not be NULL");
    if (h->opt.constitutive_relation_type == EVP_CR_EVP_REVISED)
        EVP_REQUIRE(f->uVelocityInitial && f->vVelocityInitial, "evp_revised needs u/vVelocityInitial");
    EVP_CUDA(cudaSetDevice(h->device));
    evp_dev &d = h->d;
    const size_t nC = h->nCells, nV = h->nVertices, nCp = h->nCp, nVp = h->nVp;
    const int Mk = h->M, Mh = h->Mh;
    cudaStream_t s = h->stream;
    // the staging area may still feed kernels of a previous call
    EVP_CUDA(cudaStreamSynchronize(s));
    Stage st{(char *)d.stage, d.stageBytes, 0};
    int rc;

    int *rawMs = (int *)st.take(nC * 4), *rawMv = (int *)st.take(nV * 4);
    EVP_REQUIRE(rawMs && rawMv, "staging area exhausted");
    if ((rc = evp_h2d(h, rawMs, f->solveStress, nC * 4))) return rc;
    if ((rc = evp_h2d(h, rawMv, f->solveVelocity, nV * 4))) return rc;
    if (nC) k_u8_in<<<grid_for(nC, 256), 256, 0, s>>>(rawMs, d.solveStress, nC, 1, 1);
    if (nV) k_u8_in<<<grid_for(nV, 256), 256, 0, s>>>(rawMv, d.solveVel, nV, 1, 1);
    if ((rc = evp_halo_mark_masks(h))) return rc;
    if ((rc = evp_h2d(h, d.P, f->icePressure, nC * 8))) return rc;

    struct { const double *a, *b; double2 *dst; } pairs[] = {
        {f->uVelocity, f->vVelocity, d.uv},
        {f->totalMassVertex, f->totalMassVertexfVertex, d.massf},
        {f->airStressVertexU, f->airStressVertexV, d.air},
        {f->surfaceTiltForceU, f->surfaceTiltForceV, d.tilt},
        {f->oceanStressU, f->oceanStressV, d.ocnStress},
        {f->uOceanVelocityVertex, f->vOceanVelocityVertex, d.ocnVel},
        {f->uVelocityInitial, f->vVelocityInitial, d.uvInit},
        {f->iceAreaVertex, nullptr, d.areaDen},     // .y keeps the static denominator
    };
    for (auto &p : pairs) {
        if (!p.a || nV == 0) continue;
        double *ra = (double *)st.take(nV * 8), *rb = p.b ? (double *)st.take(nV * 8) : nullptr;
        EVP_REQUIRE(ra && (rb || !p.b), "staging area exhausted");
        if ((rc = evp_h2d(h, ra, p.a, nV * 8))) return rc;
        if (p.b && (rc = evp_h2d(h, rb, p.b, nV * 8))) return rc;
        k_pair_in<<<grid_for(nV, 256), 256, 0, s>>>(ra, rb, p.dst, nV);
    }
    if (nC && f->stress11) {
        const double *src[3] = {f->stress11, f->stress22, f->stress12};
        double *dst[3] = {(double *)d.sig, (double *)d.sig, d.sig12};
        const int ncomp[3] = {2, 2, 1}, comp[3] = {0, 1, 0};
        for (int a = 0; a < 3; a++) {
            double *raw = (double *)st.take((size_t)Mh * nC * 8);
            EVP_REQUIRE(raw, "staging area exhausted");
            if ((rc = evp_h2d(h, raw, src[a], (size_t)Mh * nC * 8))) return rc;
            k_rows_in<<<grid_for(nC, 128), 128, 0, s>>>(raw, dst[a], Mh, Mk, 1, nC, 0, nCp, ncomp[a], comp[a]);
        }
    }
    // outputs that the reference zeroes at the start of every dynamics step
    // (init_subcycle_variables, velocity_solver.F:2287-2288, 2322-2324; oceanStressCoeff :2303)
    EVP_CUDA(cudaMemsetAsync(d.e11, 0, sizeof(double) * Mk * nCp, s));
    EVP_CUDA(cudaMemsetAsync(d.e22, 0, sizeof(double) * Mk * nCp, s));
    EVP_CUDA(cudaMemsetAsync(d.e12, 0, sizeof(double) * Mk * nCp, s));
    EVP_CUDA(cudaMemsetAsync(d.repP, 0, sizeof(double) * Mk * nCp, s));
    EVP_CUDA(cudaMemsetAsync(d.sdiv, 0, sizeof(double2) * nVp, s));
    EVP_CUDA(cudaMemsetAsync(d.ocoef, 0, sizeof(double) * nVp, s));
    EVP_CUDA(cudaGetLastError());
    if ((rc = evp_refresh_tile_flags(h, s))) return rc;
    // pinned sources are copied asynchronously: do not return before the host may touch them again
    EVP_CUDA(cudaStreamSynchronize(s));
    h->haveStep = true;
    return EVP_OK;
}

extern "C" int evp_run_subcycles(evp_handle *h, int nSub)
{
    EVP_REQUIRE(h != nullptr, "handle is NULL");
    EVP_REQUIRE(nSub >= 0, "nSubcycles must be >= 0");
    if (h->opt.strain_scheme == EVP_SCHEME_WEAK && !h->haveWeak) {
        evp_set_error("the weak schemes need evp_set_weak_mesh first");
        return EVP_ERR_STATE;
    }
    if (h->opt.average_variational_strain && !h->haveExt) {
        evp_set_error("average_variational_strain needs areaCell: call evp_set_mesh_ext first");
        return EVP_ERR_STATE;
    }
    if (!h->haveBasis && h->opt.stress_divergence_scheme != EVP_SCHEME_WEAK) {
        evp_set_error("basis arrays were neither given to evp_create nor precomputed");
        return EVP_ERR_STATE;
    }
    if (!h->haveStep) { evp_set_error("evp_update_step must be called before evp_run_subcycles"); return EVP_ERR_STATE; }
    EVP_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    int rc;
    if (h->useGraph && nSub > 0 && evp_persistent_run(h, nSub, s, true) == 0) {
        // small mesh: the whole loop is one cooperative launch (evp_persistent_kernel)
        EVP_CUDA(cudaEventRecord(h->ev0, s));
        if (evp_persistent_run(h, nSub, s, false) != 0) { evp_set_error("persistent kernel launch failed"); return EVP_ERR_CUDA; }
        EVP_CUDA(cudaEventRecord(h->ev1, s));
    } else if (h->useGraph && nSub > 0) {
        if (!h->graphExec || h->graphN != nSub) {
            invalidate_graph(h);
            cudaGraph_t g = nullptr;
            EVP_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
            rc = evp_enqueue_subcycles(h, nSub, s);
            cudaError_t ce = cudaStreamEndCapture(s, &g);
            if (rc) { if (g) cudaGraphDestroy(g); return rc; }
            if (ce != cudaSuccess) { evp_set_error("graph capture failed: %s", cudaGetErrorString(ce)); return EVP_ERR_CUDA; }
            ce = cudaGraphInstantiate(&h->graphExec, g, 0);
            cudaGraphDestroy(g);
            if (ce != cudaSuccess) { h->graphExec = nullptr; evp_set_error("graph instantiate failed: %s", cudaGetErrorString(ce)); return EVP_ERR_CUDA; }
            h->graphN = nSub;
        }
        EVP_CUDA(cudaEventRecord(h->ev0, s));
        EVP_CUDA(cudaGraphLaunch(h->graphExec, s));
        EVP_CUDA(cudaEventRecord(h->ev1, s));
    } else {
        EVP_CUDA(cudaEventRecord(h->ev0, s));
        if ((rc = evp_enqueue_subcycles(h, nSub, s))) return rc;
        EVP_CUDA(cudaEventRecord(h->ev1, s));
    }
    h->timed = true;
    return EVP_OK;
}

extern "C" int evp_synchronize(evp_handle *h)
{
    EVP_REQUIRE(h != nullptr, "handle is NULL");
    EVP_CUDA(cudaSetDevice(h->device));
    EVP_CUDA(cudaStreamSynchronize(h->stream));
    return evp_halo_check(h);
}

extern "C" int evp_last_run_ms(evp_handle *h, float *ms)
{
    EVP_REQUIRE(h != nullptr && ms != nullptr, "NULL argument");
    if (!h->timed) { evp_set_error("no evp_run_subcycles yet"); return EVP_ERR_STATE; }
    EVP_CUDA(cudaSetDevice(h->device));
    EVP_CUDA(cudaEventSynchronize(h->ev1));
    EVP_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return EVP_OK;
}

extern "C" int evp_fetch(evp_handle *h, const evp_out_fields *o)
{
    EVP_REQUIRE(h != nullptr && o != nullptr, "handle/out is NULL");
    EVP_CUDA(cudaSetDevice(h->device));
    evp_dev &d = h->d;
    const size_t nV = h->nVertices;
    cudaStream_t s = h->stream;
    int rc;
    struct { double *a, *b; const double2 *src; } pairs[] = {
        {o->uVelocity, o->vVelocity, d.uv},
        {o->stressDivergenceU, o->stressDivergenceV, d.sdiv},
    };
    for (auto &p : pairs) {
        if ((!p.a && !p.b) || nV == 0) continue;
        Stage st{(char *)d.stage, d.stageBytes, 0};
        double *ra = (double *)st.take(nV * 8), *rb = (double *)st.take(nV * 8);
        k_pair_out<<<grid_for(nV, 256), 256, 0, s>>>(ra, rb, p.src, nV);
        EVP_CUDA(cudaGetLastError());
        if (p.a && (rc = evp_d2h(h, p.a, ra, nV * 8))) return rc;
        if (p.b && (rc = evp_d2h(h, p.b, rb, nV * 8))) return rc;
        EVP_CUDA(cudaStreamSynchronize(s));
    }
    if (o->oceanStressCoeff && nV) {
        if ((rc = evp_d2h(h, o->oceanStressCoeff, d.ocoef, nV * 8))) return rc;
    }
    struct { double *host; const double *soa; int ncomp, comp; } rows[] = {
        {o->stress11, (const double *)d.sig, 2, 0}, {o->stress22, (const double *)d.sig, 2, 1},
        {o->stress12, d.sig12, 1, 0},
        {o->strain11, d.e11, 1, 0}, {o->strain22, d.e22, 1, 0}, {o->strain12, d.e12, 1, 0},
        {o->replacementPressure, d.repP, 1, 0},
    };
    for (auto &r : rows) {
        if (!r.host || h->nCells == 0) continue;
        if ((rc = evp_download_rows(h, r.host, r.soa, 1, r.ncomp, r.comp))) return rc;
    }
    EVP_CUDA(cudaStreamSynchronize(s));
    return evp_halo_check(h);
}

extern "C" int evp_fetch_basis(evp_handle *h, double *gu, double *gv, double *su, double *sv, double *sm)
{
    EVP_REQUIRE(h != nullptr, "handle is NULL");
    if (!h->haveBasis) { evp_set_error("no basis on the device"); return EVP_ERR_STATE; }
    EVP_CUDA(cudaSetDevice(h->device));
    NoPin noPin(h);
    evp_dev &d = h->d;
    int rc;
    if (h->nCells == 0) return EVP_OK;
    if (gu || gv) {
        double2 *G = d.G;
        const size_t gBytes = sizeof(double2) * h->M * h->M * h->nCp;
        if (!G) {   // banded on the device: expand into a temporary dense array
            EVP_CUDA(cudaMalloc((void **)&G, gBytes));
            cudaMemsetAsync(G, 0, gBytes, h->stream);
            k_band_unpack<<<grid_for(h->nCells, 128), 128, 0, h->stream>>>(d.Gb, G, d.nEdges, h->M, h->nCells, h->nCp);
        }
        rc = EVP_OK;
        if (gu) rc = evp_download_rows(h, gu, (const double *)G, 2, 2, 0);
        if (!rc && gv) rc = evp_download_rows(h, gv, (const double *)G, 2, 2, 1);
        if (!d.G) { cudaStreamSynchronize(h->stream); cudaFree(G); }
        if (rc) return rc;
    }
    if (su && (rc = evp_download_rows(h, su, (const double *)d.Suv, 2, 2, 0))) return rc;
    if (sv && (rc = evp_download_rows(h, sv, (const double *)d.Suv, 2, 2, 1))) return rc;
    if (sm && (rc = evp_download_rows(h, sm, d.Sm, 2, 1, 0))) return rc;
    EVP_CUDA(cudaStreamSynchronize(h->stream));
    return EVP_OK;
}

extern "C" int evp_release_host_memory(evp_handle *h)
{
    EVP_REQUIRE(h != nullptr, "handle is NULL");
    EVP_CUDA(cudaSetDevice(h->device));
    EVP_CUDA(cudaStreamSynchronize(h->stream));
    for (void *p : h->pinned) cudaHostUnregister(p);
    h->pinned.clear();
    cudaGetLastError();
    return EVP_OK;
}

extern "C" int evp_destroy(evp_handle *h)
{
    if (!h) return EVP_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    // the graph first: ncclCommDestroy waits for every captured graph that references the communicator
    invalidate_graph(h);
    evp_halo_destroy(h);
    for (void *p : h->pinned) cudaHostUnregister(p);
    for (void *p : h->allocs) cudaFree(p);
    for (int i = 0; i < 2; i++) {
        if (h->pinStage[i]) cudaFreeHost(h->pinStage[i]);
        if (h->pinEv[i]) cudaEventDestroy(h->pinEv[i]);
    }
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();
    delete h;
    return EVP_OK;
}

extern "C" int evp_host_metric_terms(int nVertices, const double *zRot, double radius, double *out)
{
    EVP_REQUIRE(nVertices >= 0 && (nVertices == 0 || (zRot && out)), "NULL argument");
    EVP_REQUIRE(radius > 0.0, "sphereRadius must be > 0");
    for (int v = 0; v < nVertices; v++) {
        const double lat = asin(zRot[v] / radius);      // variational_shared.F:342
        out[v] = tan(lat) / radius;                     // :344
    }
    return EVP_OK;
}

extern "C" int evp_launch_count(evp_handle *h, int nSub, int *count)
{
    EVP_REQUIRE(h != nullptr && count != nullptr, "NULL argument");
    *count = evp_count_launches(h, nSub);
    return EVP_OK;
}

extern "C" int evp_get_stream(evp_handle *h, void **stream)
{
    EVP_REQUIRE(h != nullptr && stream != nullptr, "NULL argument");
    *stream = (void *)h->stream;
    return EVP_OK;
}

extern "C" int evp_device_bytes(evp_handle *h, unsigned long long *bytes)
{
    EVP_REQUIRE(h != nullptr && bytes != nullptr, "NULL argument");
    *bytes = h->devBytes;
    return EVP_OK;
}

extern "C" int evp_set_use_graph(evp_handle *h, int useGraph)
{
    EVP_REQUIRE(h != nullptr, "handle is NULL");
    h->useGraph = useGraph != 0;
    return EVP_OK;
}
