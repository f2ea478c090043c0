// evp_halo.cu -- the per-subcycle uVelocity/vVelocity halo exchange
// (reference: src/shared/mpas_seaice_velocity_solver.F:2543-2584, the 'velocityHaloExchangeGroup'
// built at :259-349), one rank per GPU on one NVLink / NVSwitch node.
//
// Two implementations behind evp_set_halo():
//
//  * PEER-TO-PEER (default when every neighbour's memory can be mapped): no communication kernel at all.
//    Every rank exports its velocity array with cudaIpcGetMemHandle; the handles travel once, at set-up, in a
//    grouped ncclSend/ncclRecv.  In the subcycle the VERTEX kernel stores the new (u,v) of a boundary-owned
//    vertex straight into the halo slots of the neighbours that hold a copy (NVLink stores from the SM), its last
//    block publishes "vertex pass c done" into a flag word in each neighbour's memory, and the CELL kernel of the
//    next subcycle makes the threads that gather a halo vertex wait for the flags of pass c (usually set long
//    before).  The halo slots are double-buffered by the parity of the pass counter, which is all the
//    write-after-read protection needed: a neighbour's pass c+1 stores come after its cell pass c+1, which waited
//    for this rank's flag c, which was published after this rank's cell pass c finished reading.
//    Layout of the exported array (double2 elements):
//        [0, nVp)                       the velocity array as every other kernel sees it (owned, then halo)
//        [nVp, nVp + nHp)               halo buffer of even passes (halo vertex v at nVp + v - nVerticesSolve)
//        [nVp + nHp, nVp + 2 nHp)       halo buffer of odd passes
//        then 64 ints                   the incoming flags, one per neighbour
//    evp_halo_begin_run copies the canonical halo entries into the current parity's buffer, evp_halo_end_run
//    waits for the last pass of the neighbours and copies the result back, so everything outside the subcycle
//    loop (pre-/post-subcycle kernels, evp_fetch, evp_update_step) keeps using the canonical entries.
//  * NCCL (fallback: a neighbour in the same process, IPC refused, special boundaries or the weak operators in
//    use, EVP_B200_HALO=nccl): pack kernel -> grouped ncclSend/ncclRecv on the handle's stream inside the
//    captured graph, received straight into the contiguous halo slice of the field when the host's numbering
//    allows (partition.py orders halo vertices by owner), else into a buffer + unpack kernel.
//    Also used for the two once-per-step exchanges of the pre-/post-subcycle (evp_halo_exchange).
//
// NCCL is resolved at run time with dlopen so that the single-GPU library has no NCCL dependency;
// inside a process that imported torch the bundled libnccl.so.2 is found by soname.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
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
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
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
    SYM(AllReduce, "ncclAllReduce")
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
// inside ncclGroupStart .. ncclGroupEnd: close the group before returning the error
#define EVP_NCCL_G(call)                                                                            \
    do {                                                                                            \
        ncclResult_t r_ = (call);                                                                   \
        if (r_ != ncclSuccess) {                                                                    \
            evp_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
            g_nccl.GroupEnd();                                                                      \
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

// canonical halo entries -> the halo buffer of the current parity (start of a run of subcycles)
__global__ void k_halo_begin(const evp_halo_view hv, double2 *__restrict__ uvx, int nHalo)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nHalo) return;
    const int c = *(const volatile int *)hv.ctr;
    uvx[(size_t)hv.haloFirst + hv.shift0 + (size_t)(c & 1) * hv.stride + i] = uvx[(size_t)hv.haloFirst + i];
}
// wait for the neighbours' last vertex pass, then halo buffer of the current parity -> canonical entries
__global__ void k_halo_end(const evp_halo_view hv, double2 *__restrict__ uvx, int nHalo)
{
    const int c = *(const volatile int *)hv.ctr;
    if (threadIdx.x == 0) evp_halo_wait(hv, c);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nHalo) return;
    uvx[(size_t)hv.haloFirst + i] = __ldcg(&uvx[(size_t)hv.haloFirst + hv.shift0 + (size_t)(c & 1) * hv.stride + i]);
}

// what a rank tells each neighbour at set-up (one grouped ncclSend/ncclRecv of 128 bytes per neighbour)
struct HaloHello {
    cudaIpcMemHandle_t mem;          // 64 bytes: the exported velocity array
    long long flagsOffset;           // byte offset of the incoming flags inside it
    long long nVp, nHp;              // element index of halo buffer 0, distance to halo buffer 1
    int recvBase;                    // where the receiver of this message must store: its j-th send entry for me
                                     //   goes to element nVp + parity * nHp + recvBase + j
    int flagSlot;                    // ... and its flag for me is flags[flagSlot]
    int pid;
    int hostHash;
    int ok;                          // this rank is willing and able to run the peer-to-peer exchange
    int pad[5];
};
static_assert(sizeof(HaloHello) == 128, "HaloHello is 128 bytes");

}  // namespace

struct evp_halo {
    ncclComm_t comm = nullptr;
    int rank = 0, nRanks = 1;
    int nNb = 0, nSend = 0, nRecv = 0;
    std::vector<int> nbRank, sendOff, recvOff, recvBase;   // recvBase[k]: first halo vertex filled by neighbour k, or -1
    bool recvContiguous = false;  // every neighbour's recv list is one ascending run of halo vertices
    int nBoundary = 0;            // unique send-list vertices
    int *dBoundary = nullptr;
    int *dSendIdx = nullptr, *dRecvIdx = nullptr;
    double2 *dSendBuf = nullptr, *dRecvBuf = nullptr;
    size_t bytesSendIdx = 0, bytesRecvIdx = 0, bytesSendBuf = 0, bytesRecvBuf = 0, bytesBoundary = 0;
    // ---- peer-to-peer ----
    bool p2p = false;             // established with every neighbour (agreed by all ranks)
    std::string p2pWhyNot = "not attempted";
    void *xBase = nullptr;        // the exported allocation: velocity array + 2 halo buffers + flags
    size_t xBytes = 0;
    size_t nHp = 0;
    int *ctr = nullptr;           // device: completed vertex passes
    unsigned *done = nullptr;     // device: block tickets of the vertex kernel
    int *errHost = nullptr, *errDev = nullptr;   // mapped host flag: a wait timed out
    std::vector<void *> peerBase; // cudaIpcOpenMemHandle results
    int *dBStart = nullptr, *dPushStart = nullptr, *dOrder = nullptr;
    int2 *dPush = nullptr;
    double2 **dPeerUv = nullptr;
    int *dPeerStride = nullptr;
    int **dPeerFlag = nullptr;
    evp_push_view pushView{};
    std::vector<std::pair<void *, size_t>> p2pAllocs;
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
    EVP_REQUIRE(h->halo->comm == nullptr, "evp_comm_init was already called for this handle");
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    EVP_NCCL(g_nccl.CommInitRank(&h->halo->comm, nRanks, id, rank));
    h->halo->rank = rank;
    h->halo->nRanks = nRanks;
    return EVP_OK;
}

static void invalidate_graph_(evp_handle *h)
{
    if (h->graphExec) { cudaGraphExecDestroy(h->graphExec); h->graphExec = nullptr; }
    h->graphN = -1;
}

static int p2p_alloc(evp_handle *h, void **p, size_t bytes)
{
    int rc = evp_dev_alloc(h, p, bytes);
    if (rc) return rc;
    h->halo->p2pAllocs.push_back({*p, bytes ? bytes : 256});
    return EVP_OK;
}

// undo a previous evp_set_halo (lists, buffers, peer mappings); the communicator stays
static void halo_release(evp_handle *h, bool collective)
{
    evp_halo &H = *h->halo;
    cudaStreamSynchronize(h->stream);
    for (void *p : H.peerBase)
        if (p) cudaIpcCloseMemHandle(p);
    H.peerBase.clear();
    if (H.xBase) {
        // the exported allocation may only be freed after every importer closed its mapping
        if (collective && H.comm && H.p2p && H.done) {
            g_nccl.AllReduce(H.done, H.done, 1, ncclInt, ncclMax, H.comm, h->stream);
            cudaStreamSynchronize(h->stream);
        }
        // the velocity array moves back into an allocation of its own
        double2 *uv = nullptr;
        if (cudaMalloc((void **)&uv, sizeof(double2) * h->nVp) == cudaSuccess) {
            cudaMemcpy(uv, H.xBase, sizeof(double2) * h->nVp, cudaMemcpyDeviceToDevice);
            h->allocs.push_back(uv);
            h->devBytes += sizeof(double2) * h->nVp;
            h->d.uv = uv;
        }
        evp_dev_free(h, H.xBase, H.xBytes);
        H.xBase = nullptr;
    }
    for (auto &a : H.p2pAllocs) evp_dev_free(h, a.first, a.second);
    H.p2pAllocs.clear();
    if (H.errHost) { cudaFreeHost(H.errHost); H.errHost = nullptr; H.errDev = nullptr; }
    H.p2p = false;
    H.ctr = nullptr; H.done = nullptr;
    H.dBStart = H.dPushStart = H.dOrder = nullptr; H.dPush = nullptr; H.dPeerUv = nullptr; H.dPeerStride = nullptr;
    H.dPeerFlag = nullptr; H.pushView = evp_push_view{};
    evp_dev_free(h, H.dSendIdx, H.bytesSendIdx); H.dSendIdx = nullptr;
    evp_dev_free(h, H.dRecvIdx, H.bytesRecvIdx); H.dRecvIdx = nullptr;
    evp_dev_free(h, H.dSendBuf, H.bytesSendBuf); H.dSendBuf = nullptr;
    evp_dev_free(h, H.dRecvBuf, H.bytesRecvBuf); H.dRecvBuf = nullptr;
    evp_dev_free(h, H.dBoundary, H.bytesBoundary); H.dBoundary = nullptr;
    H.nNb = H.nSend = H.nRecv = H.nBoundary = 0;
    cudaGetLastError();
}

// Export the velocity array, trade handles and slots with the neighbours, map their arrays, build the push
// tables.  Collective; leaves H.p2p = true only when EVERY rank of the communicator succeeded.
static int p2p_setup(evp_handle *h, const std::vector<int> &s0)
{
    evp_halo &H = *h->halo;
    const char *mode = getenv("EVP_B200_HALO");
    const bool want = !(mode && strcmp(mode, "nccl") == 0);
    const size_t nVs = (size_t)h->nVerticesSolve, nHalo = (size_t)h->nVertices - nVs;
    int ok = 1;
    if (!want) { ok = 0; H.p2pWhyNot = "EVP_B200_HALO=nccl"; }
    else if (!H.recvContiguous) { ok = 0; H.p2pWhyNot = "a neighbour's halo vertices are not one contiguous run of the local numbering"; }
    else if (H.nNb > 64) { ok = 0; H.p2pWhyNot = "more than 64 neighbours"; }
    int rc;
    cudaStream_t s = h->stream;

    // ---- the exported allocation: [uv nVp][halo buffer 0][halo buffer 1][64 flags] ----
    H.nHp = (nHalo + 63) / 64 * 64;
    if (H.nHp == 0) H.nHp = 64;
    const size_t flagsOffset = sizeof(double2) * (h->nVp + 2 * H.nHp);
    H.xBytes = flagsOffset + 64 * sizeof(int);
    HaloHello mine{};
    if (ok) {
        if ((rc = evp_dev_alloc(h, &H.xBase, H.xBytes))) return rc;
        EVP_CUDA(cudaMemsetAsync(H.xBase, 0, H.xBytes, s));
        EVP_CUDA(cudaMemcpyAsync(H.xBase, h->d.uv, sizeof(double2) * h->nVp, cudaMemcpyDeviceToDevice, s));
        EVP_CUDA(cudaStreamSynchronize(s));
        if (cudaIpcGetMemHandle(&mine.mem, H.xBase) != cudaSuccess) {
            cudaGetLastError();
            ok = 0; H.p2pWhyNot = "cudaIpcGetMemHandle refused";
        }
    }
    char host[256] = {0};
    gethostname(host, sizeof(host) - 1);
    int hh = 0;
    for (const char *c = host; *c; c++) hh = hh * 131 + *c;
    mine.flagsOffset = (long long)flagsOffset; mine.nVp = (long long)h->nVp; mine.nHp = (long long)H.nHp;
    mine.pid = (int)getpid(); mine.hostHash = hh; mine.ok = ok;

    // ---- one 128-byte message each way per neighbour ----
    std::vector<HaloHello> out(std::max(H.nNb, 1)), in(std::max(H.nNb, 1));
    for (int k = 0; k < H.nNb; k++) {
        out[k] = mine;
        out[k].recvBase = H.recvBase[k] >= 0 ? H.recvBase[k] - (int)nVs : 0;
        out[k].flagSlot = k;
    }
    HaloHello *dOut = nullptr, *dIn = nullptr;
    int *dOk = nullptr;
    EVP_CUDA(cudaMalloc((void **)&dOut, sizeof(HaloHello) * out.size()));
    EVP_CUDA(cudaMalloc((void **)&dIn, sizeof(HaloHello) * in.size()));
    EVP_CUDA(cudaMalloc((void **)&dOk, sizeof(int)));
    auto cleanup = [&]() { cudaFree(dOut); cudaFree(dIn); cudaFree(dOk); };
    if ((rc = evp_h2d(h, dOut, out.data(), sizeof(HaloHello) * out.size()))) { cleanup(); return rc; }
    if (H.nNb) {
        ncclResult_t r = g_nccl.GroupStart();
        for (int k = 0; k < H.nNb && r == ncclSuccess; k++) {
            r = g_nccl.Send(dOut + k, sizeof(HaloHello), ncclChar, H.nbRank[k], H.comm, s);
            if (r == ncclSuccess) r = g_nccl.Recv(dIn + k, sizeof(HaloHello), ncclChar, H.nbRank[k], H.comm, s);
        }
        ncclResult_t r2 = g_nccl.GroupEnd();
        if (r == ncclSuccess) r = r2;
        if (r != ncclSuccess) { evp_set_error("halo set-up exchange: %s", g_nccl.GetErrorString(r)); cleanup(); return EVP_ERR_NCCL; }
    }
    if (cudaMemcpyAsync(in.data(), dIn, sizeof(HaloHello) * in.size(), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) { cleanup(); evp_set_error("halo set-up exchange: copy back failed"); return EVP_ERR_CUDA; }

    // ---- map the neighbours ----
    H.peerBase.assign(H.nNb, nullptr);
    for (int k = 0; k < H.nNb && ok; k++) {
        if (!in[k].ok) { ok = 0; H.p2pWhyNot = "a neighbour cannot"; break; }
        if (in[k].hostHash != hh) { ok = 0; H.p2pWhyNot = "a neighbour runs on another host"; break; }
        if (in[k].pid == mine.pid) { ok = 0; H.p2pWhyNot = "a neighbour lives in the same process (no IPC mapping)"; break; }
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, in[k].mem, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = 0; H.p2pWhyNot = "cudaIpcOpenMemHandle refused (no peer access?)";
            break;
        }
        H.peerBase[k] = p;
    }
    // ---- agree: all ranks or none ----
    {
        cudaError_t e = cudaMemcpyAsync(dOk, &ok, sizeof(int), cudaMemcpyHostToDevice, s);
        ncclResult_t r = e == cudaSuccess ? g_nccl.AllReduce(dOk, dOk, 1, ncclInt, ncclMin, H.comm, s) : ncclSystemError;
        int all = 0;
        if (r == ncclSuccess) e = cudaMemcpyAsync(&all, dOk, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (r != ncclSuccess || e != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
            cleanup(); evp_set_error("halo set-up: agreement all-reduce failed"); return EVP_ERR_NCCL;
        }
        if (ok && !all) H.p2pWhyNot = "another rank cannot";
        ok = all;
    }
    cleanup();
    if (!ok) {
        for (void *&p : H.peerBase) { if (p) cudaIpcCloseMemHandle(p); p = nullptr; }
        if (H.xBase) { evp_dev_free(h, H.xBase, H.xBytes); H.xBase = nullptr; }
        cudaGetLastError();
        if (mode && strcmp(mode, "p2p") == 0) {
            evp_set_error("EVP_B200_HALO=p2p but the peer-to-peer halo exchange is not possible: %s", H.p2pWhyNot.c_str());
            return EVP_ERR_NCCL;
        }
        return EVP_OK;
    }

    // ---- the velocity array now lives in the exported allocation ----
    evp_dev_free(h, h->d.uv, sizeof(double2) * h->nVp);
    h->d.uv = (double2 *)H.xBase;

    // ---- push tables: CSR over the boundary vertices (ascending), entries (neighbour slot, element index) ----
    std::vector<int> b(s0);
    std::sort(b.begin(), b.end());
    b.erase(std::unique(b.begin(), b.end()), b.end());
    const int nB = (int)b.size();
    std::vector<int> pushStart(nB + 1, 0);
    for (int k = 0; k < H.nNb; k++)
        for (int t = H.sendOff[k]; t < H.sendOff[k + 1]; t++)
            pushStart[(std::lower_bound(b.begin(), b.end(), s0[t]) - b.begin()) + 1]++;
    for (int i = 0; i < nB; i++) pushStart[i + 1] += pushStart[i];
    std::vector<int2> push(std::max(H.nSend, 1));
    {
        std::vector<int> fill(pushStart.begin(), pushStart.end() - 1);
        for (int k = 0; k < H.nNb; k++)
            for (int t = H.sendOff[k]; t < H.sendOff[k + 1]; t++) {
                const int e = (int)(std::lower_bound(b.begin(), b.end(), s0[t]) - b.begin());
                const long long idx = in[k].nVp + in[k].recvBase + (t - H.sendOff[k]);
                if (idx + in[k].nHp >= 0x7fffffffLL) { evp_set_error("neighbour array too large for 32-bit push indices"); return EVP_ERR_ARGUMENT; }
                push[fill[e]++] = make_int2(k, (int)idx);
            }
    }
    const int nVBlocks = (h->nVerticesSolve + 255) / 256;
    std::vector<int> bStart(nVBlocks + 2, 0);
    for (int v : b) bStart[v / 256 + 1]++;
    for (int i = 0; i <= nVBlocks; i++) bStart[i + 1] += bStart[i];
    std::vector<double2 *> peerUv(std::max(H.nNb, 1), nullptr);
    std::vector<int> peerStride(std::max(H.nNb, 1), 0);
    std::vector<int *> peerFlag(std::max(H.nNb, 1), nullptr);
    for (int k = 0; k < H.nNb; k++) {
        peerUv[k] = (double2 *)H.peerBase[k];
        peerStride[k] = (int)in[k].nHp;
        peerFlag[k] = (int *)((char *)H.peerBase[k] + in[k].flagsOffset) + in[k].flagSlot;
    }
    if ((rc = p2p_alloc(h, (void **)&H.ctr, sizeof(int)))) return rc;
    if ((rc = p2p_alloc(h, (void **)&H.done, sizeof(unsigned)))) return rc;
    if ((rc = p2p_alloc(h, (void **)&H.dBStart, sizeof(int) * bStart.size()))) return rc;
    if ((rc = p2p_alloc(h, (void **)&H.dPushStart, sizeof(int) * pushStart.size()))) return rc;
    std::vector<int> order;
    order.reserve(nVBlocks + 1);
    for (int i = 0; i < nVBlocks; i++) if (bStart[i + 1] > bStart[i]) order.push_back(i);
    const int nPushBlocks = (int)order.size();
    for (int i = 0; i < nVBlocks; i++) if (bStart[i + 1] == bStart[i]) order.push_back(i);
    if (order.empty()) order.push_back(0);
    if ((rc = p2p_alloc(h, (void **)&H.dOrder, sizeof(int) * order.size()))) return rc;
    if ((rc = evp_h2d(h, H.dOrder, order.data(), sizeof(int) * order.size()))) return rc;
    if ((rc = p2p_alloc(h, (void **)&H.dPush, sizeof(int2) * push.size()))) return rc;
    if ((rc = p2p_alloc(h, (void **)&H.dPeerUv, sizeof(double2 *) * peerUv.size()))) return rc;
    if ((rc = p2p_alloc(h, (void **)&H.dPeerStride, sizeof(int) * peerStride.size()))) return rc;
    if ((rc = p2p_alloc(h, (void **)&H.dPeerFlag, sizeof(int *) * peerFlag.size()))) return rc;
    EVP_CUDA(cudaMemsetAsync(H.ctr, 0, sizeof(int), s));
    EVP_CUDA(cudaMemsetAsync(H.done, 0, sizeof(unsigned), s));
    if ((rc = evp_h2d(h, H.dBStart, bStart.data(), sizeof(int) * bStart.size()))) return rc;
    if ((rc = evp_h2d(h, H.dPushStart, pushStart.data(), sizeof(int) * pushStart.size()))) return rc;
    if ((rc = evp_h2d(h, H.dPush, push.data(), sizeof(int2) * push.size()))) return rc;
    if ((rc = evp_h2d(h, H.dPeerUv, peerUv.data(), sizeof(double2 *) * peerUv.size()))) return rc;
    if ((rc = evp_h2d(h, H.dPeerStride, peerStride.data(), sizeof(int) * peerStride.size()))) return rc;
    if ((rc = evp_h2d(h, H.dPeerFlag, peerFlag.data(), sizeof(int *) * peerFlag.size()))) return rc;
    EVP_CUDA(cudaHostAlloc((void **)&H.errHost, sizeof(int), cudaHostAllocMapped));
    *H.errHost = 0;
    EVP_CUDA(cudaHostGetDevicePointer((void **)&H.errDev, H.errHost, 0));
    evp_push_view pv{};
    pv.ctr = H.ctr; pv.done = H.done; pv.bStart = H.dBStart; pv.pushStart = H.dPushStart; pv.push = H.dPush;
    pv.peerUv = H.dPeerUv; pv.peerStride = H.dPeerStride; pv.peerFlag = H.dPeerFlag; pv.nNb = H.nNb;
    pv.nPushBlocks = nPushBlocks;
    pv.order = H.dOrder;
    H.pushView = pv;
    EVP_CUDA(cudaStreamSynchronize(s));
    H.p2p = true;
    H.p2pWhyNot.clear();
    return EVP_OK;
}

extern "C" int evp_set_halo(evp_handle *h, int nNb, const int *nbRank, const int *sendOff, const int *sendIdx,
                            const int *recvOff, const int *recvIdx)
{
    EVP_REQUIRE(h != nullptr, "handle is NULL");
    EVP_REQUIRE(nNb >= 0, "negative neighbour count");
    if (nNb > 0) EVP_REQUIRE(nbRank && sendOff && recvOff, "NULL halo list");
    EVP_CUDA(cudaSetDevice(h->device));
    if (!h->halo) h->halo = new evp_halo();
    evp_halo &H = *h->halo;
    // ---- validate everything before anything is allocated or changed ----
    for (int k = 0; k < nNb; k++) {
        EVP_REQUIRE(nbRank[k] >= 0 && (H.comm == nullptr || nbRank[k] < H.nRanks), "neighbour rank out of range");
        EVP_REQUIRE(H.comm == nullptr || nbRank[k] != H.rank, "a rank cannot be its own neighbour");
        EVP_REQUIRE(sendOff[k + 1] >= sendOff[k] && recvOff[k + 1] >= recvOff[k], "halo offsets must not decrease");
    }
    if (nNb > 0) EVP_REQUIRE(sendOff[0] == 0 && recvOff[0] == 0, "halo offsets must start at 0");
    const int nSend = nNb ? sendOff[nNb] : 0, nRecv = nNb ? recvOff[nNb] : 0;
    EVP_REQUIRE((nSend == 0 || sendIdx) && (nRecv == 0 || recvIdx), "NULL halo index list");
    std::vector<int> s0(nSend), r0(nRecv);
    for (int k = 0; k < nSend; k++) {
        s0[k] = sendIdx[k] - 1;
        EVP_REQUIRE(s0[k] >= 0 && s0[k] < h->nVerticesSolve, "send index is not an owned vertex");
    }
    for (int k = 0; k < nRecv; k++) {
        r0[k] = recvIdx[k] - 1;
        EVP_REQUIRE(r0[k] >= h->nVerticesSolve && r0[k] < h->nVertices, "recv index is not a halo vertex");
    }
    if (H.nNb || H.dSendIdx || H.xBase) halo_release(h, true);      // a repeated call replaces the previous lists
    invalidate_graph_(h);
    H.nNb = nNb;
    H.nbRank.assign(nbRank, nbRank + nNb);
    H.sendOff.assign(1, 0); H.recvOff.assign(1, 0);
    if (nNb) { H.sendOff.assign(sendOff, sendOff + nNb + 1); H.recvOff.assign(recvOff, recvOff + nNb + 1); }
    H.nSend = nSend;
    H.nRecv = nRecv;
    H.recvBase.assign(nNb, -1);
    H.recvContiguous = true;
    for (int k = 0; k < nNb; k++) {
        const int a = H.recvOff[k], b = H.recvOff[k + 1];
        if (b > a) H.recvBase[k] = r0[a];
        for (int t = a + 1; t < b; t++)
            if (r0[t] != r0[t - 1] + 1) H.recvContiguous = false;
    }
    int rc;
    H.bytesSendIdx = sizeof(int) * (H.nSend + 1); H.bytesRecvIdx = sizeof(int) * (H.nRecv + 1);
    H.bytesSendBuf = sizeof(double2) * (H.nSend + 1); H.bytesRecvBuf = sizeof(double2) * (H.nRecv + 1);
    if ((rc = evp_dev_alloc(h, (void **)&H.dSendIdx, H.bytesSendIdx))) return rc;
    if ((rc = evp_dev_alloc(h, (void **)&H.dRecvIdx, H.bytesRecvIdx))) return rc;
    if ((rc = evp_dev_alloc(h, (void **)&H.dSendBuf, H.bytesSendBuf))) return rc;
    if ((rc = evp_dev_alloc(h, (void **)&H.dRecvBuf, H.bytesRecvBuf))) return rc;
    {   // boundary-owned vertices = the distinct entries of the send lists, ascending
        std::vector<int> b(s0);
        std::sort(b.begin(), b.end());
        b.erase(std::unique(b.begin(), b.end()), b.end());
        H.nBoundary = (int)b.size();
        H.bytesBoundary = sizeof(int) * (b.size() + 1);
        if ((rc = evp_dev_alloc(h, (void **)&H.dBoundary, H.bytesBoundary))) return rc;
        if (!b.empty() && (rc = evp_h2d(h, H.dBoundary, b.data(), sizeof(int) * b.size()))) return rc;
    }
    if (H.nSend && (rc = evp_h2d(h, H.dSendIdx, s0.data(), sizeof(int) * H.nSend))) return rc;
    if (H.nRecv && (rc = evp_h2d(h, H.dRecvIdx, r0.data(), sizeof(int) * H.nRecv))) return rc;
    EVP_CUDA(cudaStreamSynchronize(h->stream));
    if (nNb == 0 && !H.comm) return evp_halo_mark_masks(h);
    // NCCL opens its point-to-point connections lazily inside the first ncclGroupEnd that uses them: a
    // host-side handshake between the ranks plus allocations, none of which may happen while the stream is
    // being captured into the subcycle graph.  Do one eager exchange of the (still meaningless) pack
    // buffers now -- evp_set_halo is therefore COLLECTIVE over the ranks of the communicator.
    if (H.comm && nNb > 0) {
        EVP_CUDA(cudaMemsetAsync(H.dSendBuf, 0, sizeof(double2) * (H.nSend + 1), h->stream));
        EVP_NCCL(g_nccl.GroupStart());
        for (int k = 0; k < H.nNb; k++) {
            const int ns = H.sendOff[k + 1] - H.sendOff[k], nr = H.recvOff[k + 1] - H.recvOff[k];
            if (ns) EVP_NCCL_G(g_nccl.Send(H.dSendBuf + H.sendOff[k], (size_t)2 * ns, ncclDouble, H.nbRank[k], H.comm, h->stream));
            if (nr) EVP_NCCL_G(g_nccl.Recv(H.dRecvBuf + H.recvOff[k], (size_t)2 * nr, ncclDouble, H.nbRank[k], H.comm, h->stream));
        }
        EVP_NCCL(g_nccl.GroupEnd());
        EVP_CUDA(cudaStreamSynchronize(h->stream));
    }
    if (H.comm && (rc = p2p_setup(h, s0))) return rc;
    if ((rc = evp_halo_mark_masks(h))) return rc;
    // the lists of vertex blocks with work depend on the boundary bit just (re)applied
    if (h->haveStep && (rc = evp_refresh_tile_flags(h, h->stream))) return rc;
    EVP_CUDA(cudaStreamSynchronize(h->stream));
    return EVP_OK;
}

// the peer-to-peer exchange is established AND usable with the current options (all ranks share the options, so
// they take the same branch): the kernels that know the double-buffered halo slots are the variational cell
// kernel and the vertex kernel; the special-boundary kernels and the weak operators read the canonical entries
bool evp_halo_p2p_active(evp_handle *h)
{
    if (!h->halo || !h->halo->p2p) return false;
    if (h->opt.strain_scheme == EVP_SCHEME_WEAK) return false;
    if (h->opt.use_special_boundaries_velocity && h->d.nSB) return false;
    return true;
}

evp_halo_view evp_halo_get_view(evp_handle *h)
{
    evp_halo_view v{};
    if (!evp_halo_p2p_active(h)) return v;
    evp_halo &H = *h->halo;
    v.ctr = H.ctr;
    v.flagsIn = (const int *)((char *)H.xBase + sizeof(double2) * (h->nVp + 2 * H.nHp));
    v.err = H.errDev;
    v.nNb = H.nNb;
    v.haloFirst = h->nVerticesSolve;
    v.shift0 = (int)(h->nVp - (size_t)h->nVerticesSolve);
    v.stride = (int)H.nHp;
    return v;
}

evp_push_view evp_halo_get_push(evp_handle *h)
{
    if (!evp_halo_p2p_active(h)) return evp_push_view{};
    evp_push_view pv = h->halo->pushView;
    // timing experiments (tools/gpu_job_p2p_dbg.sh): 1 = boundary stores go to local memory, 2 = no system fence in the
    // pushing blocks, 4 = no block pushes at all (the extra block publishes), 8 = device- instead of system-scope fence
    const char *e = getenv("EVP_B200_P2P_DEBUG");
    if (e) {
        pv.dbg = atoi(e);
        if (pv.dbg & 4) pv.nPushBlocks = 0;
    }
    return pv;
}

int evp_halo_begin_run(evp_handle *h, cudaStream_t s)
{
    if (!evp_halo_p2p_active(h)) return EVP_OK;
    const int nHalo = h->nVertices - h->nVerticesSolve;
    if (nHalo > 0) k_halo_begin<<<(nHalo + 255) / 256, 256, 0, s>>>(evp_halo_get_view(h), h->d.uv, nHalo);
    EVP_CUDA(cudaGetLastError());
    return EVP_OK;
}

int evp_halo_end_run(evp_handle *h, cudaStream_t s)
{
    if (!evp_halo_p2p_active(h)) return EVP_OK;
    const int nHalo = h->nVertices - h->nVerticesSolve;
    // launched even without halo vertices: the wait is what keeps this rank from leaving (and re-entering) the
    // loop while a neighbour still stores into its buffers
    k_halo_end<<<std::max((nHalo + 255) / 256, 1), 256, 0, s>>>(evp_halo_get_view(h), h->d.uv, nHalo);
    EVP_CUDA(cudaGetLastError());
    return EVP_OK;
}

// a wait on a neighbour's flag ran into its time limit (a rank died or left the loop): results are not valid
int evp_halo_check(evp_handle *h)
{
    if (!h->halo || !h->halo->errHost || *(volatile int *)h->halo->errHost == 0) return EVP_OK;
    evp_set_error("peer-to-peer halo exchange: a wait for a neighbour's vertex pass timed out");
    return EVP_ERR_NCCL;
}

extern "C" int evp_halo_mode(evp_handle *h, int *mode, char *why, int whyLen)
{
    EVP_REQUIRE(h != nullptr && mode != nullptr, "NULL argument");
    *mode = EVP_HALO_NONE;
    if (h->halo && h->halo->comm && h->halo->nNb > 0) *mode = evp_halo_p2p_active(h) ? EVP_HALO_P2P : EVP_HALO_NCCL;
    if (why && whyLen > 0) {
        const char *w = (h->halo && *mode == EVP_HALO_NCCL)
                            ? (h->halo->p2p ? "options in use need the canonical halo entries (weak operators / special boundaries)"
                                            : h->halo->p2pWhyNot.c_str())
                            : "";
        strncpy(why, w, (size_t)whyLen - 1);
        why[whyLen - 1] = 0;
    }
    return EVP_OK;
}

int evp_halo_boundary_count(evp_handle *h) { return h->halo ? h->halo->nBoundary : 0; }

// (re)apply the boundary bit to the velocity mask; called after every mask upload.  The bit is only used by the
// peer-to-peer exchange (the vertex kernel pushes the vertices that carry it)
int evp_halo_mark_masks(evp_handle *h)
{
    if (!h->halo || h->nVertices == 0) return EVP_OK;
    k_clear_boundary<<<(unsigned)((h->nVp + 255) / 256), 256, 0, h->stream>>>(h->nVp, h->d.solveVel);
    if (h->halo->nBoundary && h->halo->p2p)
        k_mark_boundary<<<(h->halo->nBoundary + 255) / 256, 256, 0, h->stream>>>(h->halo->nBoundary, h->halo->dBoundary,
                                                                                 h->d.solveVel);
    EVP_CUDA(cudaGetLastError());
    return EVP_OK;
}

// kernels of our own per exchange on the NCCL path (NCCL's kernel is not counted)
int evp_halo_launches(evp_handle *h)
{
    if (!h->halo || !h->halo->comm || h->halo->nNb == 0 || evp_halo_p2p_active(h)) return 0;
    return (h->halo->nSend ? 1 : 0) + ((h->halo->nRecv && !h->halo->recvContiguous) ? 1 : 0);
}

// the in-loop exchange of d.uv on the NCCL path; a no-op when the vertex kernel pushes by itself
int evp_halo_enqueue(evp_handle *h, cudaStream_t s)
{
    if (evp_halo_p2p_active(h)) return EVP_OK;
    return evp_halo_exchange(h, s, h->d.uv);
}

int evp_halo_exchange(evp_handle *h, cudaStream_t s, double2 *field)
{
    if (!h->halo || h->halo->nNb == 0) return EVP_OK;
    evp_halo &H = *h->halo;
    if (!H.comm) { evp_set_error("evp_set_halo without evp_comm_init"); return EVP_ERR_STATE; }
    if (H.nSend) k_pack<<<(H.nSend + 255) / 256, 256, 0, s>>>(H.nSend, H.dSendIdx, field, H.dSendBuf);
    EVP_NCCL(g_nccl.GroupStart());
    for (int k = 0; k < H.nNb; k++) {
        const int ns = H.sendOff[k + 1] - H.sendOff[k], nr = H.recvOff[k + 1] - H.recvOff[k];
        if (ns) EVP_NCCL_G(g_nccl.Send(H.dSendBuf + H.sendOff[k], (size_t)2 * ns, ncclDouble, H.nbRank[k], H.comm, s));
        // partition.py groups the halo vertices by owner: what a neighbour sends is one contiguous slice of the field
        double2 *dst = H.recvContiguous ? field + H.recvBase[k] : H.dRecvBuf + H.recvOff[k];
        if (nr) EVP_NCCL_G(g_nccl.Recv(dst, (size_t)2 * nr, ncclDouble, H.nbRank[k], H.comm, s));
    }
    EVP_NCCL(g_nccl.GroupEnd());
    if (H.nRecv && !H.recvContiguous) k_unpack<<<(H.nRecv + 255) / 256, 256, 0, s>>>(H.nRecv, H.dRecvIdx, H.dRecvBuf, field);
    EVP_CUDA(cudaGetLastError());
    return EVP_OK;
}

void evp_halo_destroy(evp_handle *h)
{
    if (!h->halo) return;
    halo_release(h, true);
    if (h->halo->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->halo->comm);
    delete h->halo;
    h->halo = nullptr;
}
