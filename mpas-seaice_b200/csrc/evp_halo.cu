// evp_halo.cu -- the per-subcycle uVelocity/vVelocity halo exchange
// (reference: src/shared/mpas_seaice_velocity_solver.F:2543-2584, the 'velocityHaloExchangeGroup'
// built at :259-349) as grouped ncclSend/ncclRecv over NVLink, enqueued on the same stream (and so
// captured in the same CUDA graph) as the two compute kernels.
//
// NCCL is resolved at run time with dlopen so that the single-GPU library has no NCCL dependency;
// inside a process that imported torch the bundled libnccl.so.2 is found by soname.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <algorithm>
#include "evp_internal.cuh"

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl()
{
    if (g_nccl.lib) return EVP_OK;
    const char *names[] = {getenv("EVP_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *n : names) {
        if (!n) continue;
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) {
        evp_set_error("NCCL not found (set EVP_B200_NCCL_LIB): %s", dlerror());
        return EVP_ERR_NCCL;
    }
#define SYM(field, name)                                                          \
    *(void **)(&g_nccl.field) = dlsym(lib, name);                                 \
    if (!g_nccl.field) { evp_set_error("NCCL symbol %s missing", name); return EVP_ERR_NCCL; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = lib;
    return EVP_OK;
}

#define EVP_NCCL(call)                                                                              \
    do {                                                                                            \
        ncclResult_t r_ = (call);                                                                   \
        if (r_ != ncclSuccess) {                                                                    \
            evp_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
            return EVP_ERR_NCCL;                                                                    \
        }                                                                                           \
    } while (0)

__global__ void k_pack(int n, const int *__restrict__ idx, const double2 *__restrict__ uv, double2 *__restrict__ buf)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) buf[k] = uv[idx[k]];
}
// bit 1 of the velocity mask byte marks the boundary-owned vertices (see evp_vertex_kernel)
__global__ void k_mark_boundary(int n, const int *__restrict__ idx, uint8_t *__restrict__ mask)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) mask[idx[k]] |= 2;
}
__global__ void k_clear_boundary(size_t n, uint8_t *__restrict__ mask)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) mask[k] &= 1;
}
__global__ void k_unpack(int n, const int *__restrict__ idx, const double2 *__restrict__ buf, double2 *__restrict__ uv)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) uv[idx[k]] = buf[k];
}

}  // namespace

struct evp_halo {
    ncclComm_t comm = nullptr;
    int rank = 0, nRanks = 1;
    int nNb = 0, nSend = 0, nRecv = 0;
    std::vector<int> nbRank, sendOff, recvOff;
    int nBoundary = 0;            // unique send-list vertices
    int *dBoundary = nullptr;
    int *dSendIdx = nullptr, *dRecvIdx = nullptr;
    double2 *dSendBuf = nullptr, *dRecvBuf = nullptr;
};

extern "C" int evp_comm_get_unique_id(char *id128)
{
    EVP_REQUIRE(id128 != nullptr, "id buffer is NULL");
    int rc = load_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    EVP_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return EVP_OK;
}

extern "C" int evp_comm_init(evp_handle *h, int rank, int nRanks, const char *id128)
{
    EVP_REQUIRE(h != nullptr && id128 != nullptr, "NULL argument");
    EVP_REQUIRE(nRanks >= 1 && rank >= 0 && rank < nRanks, "bad rank / nRanks");
    int rc = load_nccl();
    if (rc) return rc;
    EVP_CUDA(cudaSetDevice(h->device));
    if (!h->halo) h->halo = new evp_halo();
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    EVP_NCCL(g_nccl.CommInitRank(&h->halo->comm, nRanks, id, rank));
    h->halo->rank = rank;
    h->halo->nRanks = nRanks;
    return EVP_OK;
}

extern "C" int evp_set_halo(evp_handle *h, int nNb, const int *nbRank, const int *sendOff, const int *sendIdx,
                            const int *recvOff, const int *recvIdx)
{
    EVP_REQUIRE(h != nullptr, "handle is NULL");
    EVP_REQUIRE(nNb >= 0, "negative neighbour count");
    if (nNb > 0) EVP_REQUIRE(nbRank && sendOff && sendIdx && recvOff && recvIdx, "NULL halo list");
    EVP_CUDA(cudaSetDevice(h->device));
    if (!h->halo) h->halo = new evp_halo();
    evp_halo &H = *h->halo;
    H.nNb = nNb;
    H.nbRank.assign(nbRank, nbRank + nNb);
    H.sendOff.assign(sendOff, sendOff + nNb + 1);
    H.recvOff.assign(recvOff, recvOff + nNb + 1);
    H.nSend = nNb ? sendOff[nNb] : 0;
    H.nRecv = nNb ? recvOff[nNb] : 0;
    std::vector<int> s0(H.nSend), r0(H.nRecv);
    for (int k = 0; k < H.nSend; k++) {
        s0[k] = sendIdx[k] - 1;
        EVP_REQUIRE(s0[k] >= 0 && s0[k] < h->nVerticesSolve, "send index is not an owned vertex");
    }
    for (int k = 0; k < H.nRecv; k++) {
        r0[k] = recvIdx[k] - 1;
        EVP_REQUIRE(r0[k] >= h->nVerticesSolve && r0[k] < h->nVertices, "recv index is not a halo vertex");
    }
    int rc;
    if ((rc = evp_dev_alloc(h, (void **)&H.dSendIdx, sizeof(int) * (H.nSend + 1)))) return rc;
    if ((rc = evp_dev_alloc(h, (void **)&H.dRecvIdx, sizeof(int) * (H.nRecv + 1)))) return rc;
    if ((rc = evp_dev_alloc(h, (void **)&H.dSendBuf, sizeof(double2) * (H.nSend + 1)))) return rc;
    if ((rc = evp_dev_alloc(h, (void **)&H.dRecvBuf, sizeof(double2) * (H.nRecv + 1)))) return rc;
    {   // boundary-owned vertices = the distinct entries of the send lists, ascending
        std::vector<int> b(s0);
        std::sort(b.begin(), b.end());
        b.erase(std::unique(b.begin(), b.end()), b.end());
        H.nBoundary = (int)b.size();
        if ((rc = evp_dev_alloc(h, (void **)&H.dBoundary, sizeof(int) * (b.size() + 1)))) return rc;
        if (!b.empty()) EVP_CUDA(cudaMemcpy(H.dBoundary, b.data(), sizeof(int) * b.size(), cudaMemcpyHostToDevice));
        if ((rc = evp_halo_mark_masks(h))) return rc;
    }
    if (H.nSend) EVP_CUDA(cudaMemcpy(H.dSendIdx, s0.data(), sizeof(int) * H.nSend, cudaMemcpyHostToDevice));
    if (H.nRecv) EVP_CUDA(cudaMemcpy(H.dRecvIdx, r0.data(), sizeof(int) * H.nRecv, cudaMemcpyHostToDevice));
    if (h->graphExec) { cudaGraphExecDestroy(h->graphExec); h->graphExec = nullptr; h->graphN = -1; }
    // NCCL opens its point-to-point connections lazily inside the first ncclGroupEnd that uses them: a
    // host-side handshake between the ranks plus allocations, none of which may happen while the stream is
    // being captured into the subcycle graph.  Do one eager exchange of the (still meaningless) pack
    // buffers now -- evp_set_halo is therefore COLLECTIVE over the ranks of the communicator.
    if (H.comm && nNb > 0) {
        EVP_CUDA(cudaMemsetAsync(H.dSendBuf, 0, sizeof(double2) * (H.nSend + 1), h->stream));
        EVP_NCCL(g_nccl.GroupStart());
        for (int k = 0; k < H.nNb; k++) {
            const int ns = H.sendOff[k + 1] - H.sendOff[k], nr = H.recvOff[k + 1] - H.recvOff[k];
            if (ns) EVP_NCCL(g_nccl.Send(H.dSendBuf + H.sendOff[k], (size_t)2 * ns, ncclDouble, H.nbRank[k], H.comm, h->stream));
            if (nr) EVP_NCCL(g_nccl.Recv(H.dRecvBuf + H.recvOff[k], (size_t)2 * nr, ncclDouble, H.nbRank[k], H.comm, h->stream));
        }
        EVP_NCCL(g_nccl.GroupEnd());
        EVP_CUDA(cudaStreamSynchronize(h->stream));
    }
    return EVP_OK;
}

int evp_halo_boundary_count(evp_handle *h)
{
    if (!h->halo || !h->halo->comm || h->halo->nNb == 0 || !(h->opt.flags & EVP_FLAG_OVERLAP_HALO)) return 0;
    return h->halo->nBoundary;
}
const int *evp_halo_boundary_list(evp_handle *h) { return h->halo ? h->halo->dBoundary : nullptr; }

// (re)apply the boundary bit to the velocity mask; called after every mask upload
int evp_halo_mark_masks(evp_handle *h)
{
    if (!h->halo || h->nVertices == 0) return EVP_OK;
    k_clear_boundary<<<(unsigned)((h->nVp + 255) / 256), 256, 0, h->stream>>>(h->nVp, h->d.solveVel);
    if (h->halo->nBoundary && (h->opt.flags & EVP_FLAG_OVERLAP_HALO))
        k_mark_boundary<<<(h->halo->nBoundary + 255) / 256, 256, 0, h->stream>>>(h->halo->nBoundary, h->halo->dBoundary,
                                                                                 h->d.solveVel);
    EVP_CUDA(cudaGetLastError());
    return EVP_OK;
}

int evp_halo_launches(evp_handle *h)
{
    if (!h->halo || !h->halo->comm || h->halo->nNb == 0) return 0;
    return (h->halo->nSend ? 1 : 0) + (h->halo->nRecv ? 1 : 0);   // pack + unpack (NCCL's own kernel not counted)
}

int evp_halo_enqueue(evp_handle *h, cudaStream_t s) { return evp_halo_exchange(h, s, h->d.uv); }

int evp_halo_exchange(evp_handle *h, cudaStream_t s, double2 *field)
{
    if (!h->halo || h->halo->nNb == 0) return EVP_OK;
    evp_halo &H = *h->halo;
    if (!H.comm) { evp_set_error("evp_set_halo without evp_comm_init"); return EVP_ERR_STATE; }
    if (H.nSend) k_pack<<<(H.nSend + 255) / 256, 256, 0, s>>>(H.nSend, H.dSendIdx, field, H.dSendBuf);
    EVP_NCCL(g_nccl.GroupStart());
    for (int k = 0; k < H.nNb; k++) {
        const int ns = H.sendOff[k + 1] - H.sendOff[k], nr = H.recvOff[k + 1] - H.recvOff[k];
        if (ns) EVP_NCCL(g_nccl.Send(H.dSendBuf + H.sendOff[k], (size_t)2 * ns, ncclDouble, H.nbRank[k], H.comm, s));
        if (nr) EVP_NCCL(g_nccl.Recv(H.dRecvBuf + H.recvOff[k], (size_t)2 * nr, ncclDouble, H.nbRank[k], H.comm, s));
    }
    EVP_NCCL(g_nccl.GroupEnd());
    if (H.nRecv) k_unpack<<<(H.nRecv + 255) / 256, 256, 0, s>>>(H.nRecv, H.dRecvIdx, H.dRecvBuf, field);
    EVP_CUDA(cudaGetLastError());
    return EVP_OK;
}

void evp_halo_destroy(evp_handle *h)
{
    if (!h->halo) return;
    if (h->halo->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->halo->comm);
    delete h->halo;
    h->halo = nullptr;
}
