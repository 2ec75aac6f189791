// comm.cu -- multi-GPU execution, one process per GPU (new design; the reference has no distributed code,
// SURVEY.md 8e).  NCCL is loaded at run time with dlopen("libnccl.so.2") so that the library the host
// process already uses (e.g. torch's bundled NCCL) is shared; the communicator is bootstrapped from a
// 128-byte ncclUniqueId that the host language distributes (torch.distributed / MPI.jl).
//
//  SHARD_BATCH : rank r owns transforms [r*B/P, (r+1)*B/P) of a batched plan; nodes/permutation/tables are
//                replicated at nodes!; no collective on the data path.
//  SHARD_NODES : the tile-sorted node list is cut into P tile-aligned ranges of ~M/P nodes.
//     adjoint : local spread of the own range onto a full-grid replica
//               -> ncclReduceScatter(sum) over slabs of the slowest grid dimension
//               -> slab-decomposed FFT (D >= 2): local FFT over dims 1..D-1, all-to-all transpose, FFT over dim D
//               -> crop/apodise on the slab owners -> all-reduce of the (small) image
//     forward : apodise/zero-pad on the transposed slabs -> FFT over dim D -> all-to-all back -> local FFT
//               over dims 1..D-1 -> ncclAllGather ("broadcast the grid") -> interpolate the own node range
//     For D == 1 (and when the grid does not divide) the FFT is replicated after an all-gather.
//  Fused spread + reduce-scatter over peer memory (3-D, default when it applies; kernel mode 6 selects the NCCL
//  reduce-scatter baseline): every rank spreads its tile range into a per-tile scratch that is exported with
//  CUDA IPC; after one stream-ordered barrier each rank's gather pass builds ITS z-slab directly, reading the
//  tiles that cover it from whichever rank owns them (P2P loads over NVLink).  Only the tiles straddling a slab
//  boundary cross the link instead of (P-1)/P of the grid, nothing is atomically added, the full-grid replica is
//  never written, and the summation order is fixed.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace {

struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& api()
{
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        a.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.h) a.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!a.h) return;
#define LOAD(field, sym) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.h, sym)); if (!a.field) return;
        LOAD(GetUniqueId, "ncclGetUniqueId") LOAD(CommInitRank, "ncclCommInitRank") LOAD(CommDestroy, "ncclCommDestroy")
        LOAD(ReduceScatter, "ncclReduceScatter") LOAD(AllGather, "ncclAllGather") LOAD(AllReduce, "ncclAllReduce")
        LOAD(Send, "ncclSend") LOAD(Recv, "ncclRecv") LOAD(GroupStart, "ncclGroupStart") LOAD(GroupEnd, "ncclGroupEnd")
        LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
        a.ok = true;
    });
    return a;
}

#define NCCL_TRY(p, call)                                                                      \
    do {                                                                                       \
        ncclResult_t r__ = (call);                                                             \
        if (r__ != ncclSuccess)                                                                \
            return nfftb_fail((p), NFFTB200_NCCL_ERROR, std::string(#call) + ": " + api().GetErrorString(r__)); \
    } while (0)

struct CommState {
    ncclComm_t comm = nullptr;
    // slab FFT resources (D >= 2)
    bool slab = false;
    cufftHandle fft_local = 0, fft_last = 0;
    bool have_local = false, have_last = false;
    void* d_a = nullptr;      // z-slab / transposed slab work buffer  (gsz/P complex)
    void* d_b = nullptr;      // pack / unpack buffer                  (gsz/P complex)
    int64_t inner = 1, mid = 1, outer = 1, Ms = 1, Os = 1;
    // fused spread + reduce-scatter over peer memory
    bool fused = false;
    void* d_peerbuf = nullptr; int64_t cap_peerbuf = 0;   // own tile scratch (IPC-exported)
    void* peer_ptr[NFFTB_MAX_PEERS] = {};                 // mappings of the other ranks' scratch
    PeerTab tab{};
    void* d_xchg = nullptr;                               // handle exchange / barrier word
    // forward: interpolation straight from the ranks' z-slabs instead of an all-gather of the grid
    bool fused_fwd = false;
    void* slab_ptr[NFFTB_MAX_PEERS] = {};                 // mappings of the other ranks' d_a
    SlabTab slabs{};
};

struct PeerMsg {
    cudaIpcMemHandle_t h;
    int ok;
    int pad;
};

CommState* cs(nfftb200_plan* p) { return reinterpret_cast<CommState*>(p->nccl_comm); }

ncclDataType_t nccl_real(const nfftb200_plan* p) { return p->dtype == NFFTB200_F32 ? ncclFloat : ncclDouble; }

// send[q][o][m][i] = slab[o][q*Ms + m][i]   (z-slab -> blocks for the all-to-all);  inverse with unpack = true
template <typename C>
__global__ void k_pack_blocks(const C* __restrict__ src, C* __restrict__ dst, long long inner, long long Ms,
                              long long Os, int P, bool unpack)
{
    const long long n = inner * Ms * Os * P;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e % inner;
        long long r = e / inner;
        const long long m = r % Ms; r /= Ms;
        const long long o = r % Os;
        const long long q = r / Os;
        const long long slab_idx = (o * (Ms * P) + q * Ms + m) * inner + i;
        if (unpack) dst[slab_idx] = src[e]; else dst[e] = src[slab_idx];
    }
}

int all_to_all(nfftb200_plan* p, const void* send, void* recv, size_t block_reals)
{
    CommState* c = cs(p);
    const size_t esz = p->esz();
    NCCL_TRY(p, api().GroupStart());
    for (int q = 0; q < p->nranks; q++) {
        NCCL_TRY(p, api().Send((const char*)send + (size_t)q * block_reals * esz, block_reals, nccl_real(p), q, c->comm, p->stream));
        NCCL_TRY(p, api().Recv((char*)recv + (size_t)q * block_reals * esz, block_reals, nccl_real(p), q, c->comm, p->stream));
    }
    NCCL_TRY(p, api().GroupEnd());
    p->launches++;
    return NFFTB200_OK;
}

int exec_fft(nfftb200_plan* p, cufftHandle h, void* buf, int dir)
{
    const int cdir = dir < 0 ? CUFFT_FORWARD : CUFFT_INVERSE;
    if (p->dtype == NFFTB200_F32) CUFFT_TRY(p, cufftExecC2C(h, (cufftComplex*)buf, (cufftComplex*)buf, cdir));
    else CUFFT_TRY(p, cufftExecZ2Z(h, (cufftDoubleComplex*)buf, (cufftDoubleComplex*)buf, cdir));
    p->launches++;
    return NFFTB200_OK;
}

int plan_fft_replicated(nfftb200_plan* p, void* grid, int dir)
{
    if (p->dtype == NFFTB200_F32) CUFFT_TRY(p, cufftExecC2C(p->fft, (cufftComplex*)grid, (cufftComplex*)grid, dir < 0 ? CUFFT_FORWARD : CUFFT_INVERSE));
    else CUFFT_TRY(p, cufftExecZ2Z(p->fft, (cufftDoubleComplex*)grid, (cufftDoubleComplex*)grid, dir < 0 ? CUFFT_FORWARD : CUFFT_INVERSE));
    p->launches++;
    return NFFTB200_OK;
}

int setup_slab(nfftb200_plan* p)
{
    CommState* c = cs(p);
    const int D = p->D, P = p->nranks;
    c->slab = false;
    if (D < 2) return NFFTB200_OK;
    c->outer = p->Nt[D - 1]; c->mid = p->Nt[D - 2]; c->inner = 1;
    for (int d = 0; d < D - 2; d++) c->inner *= p->Nt[d];
    if (c->outer % P || c->mid % P) return NFFTB200_OK;            // falls back to the replicated FFT
    c->Os = c->outer / P; c->Ms = c->mid / P;
    const size_t slab_bytes = (size_t)(p->gsz / P) * 2 * p->esz();
    CUDA_TRY(p, cudaMalloc(&c->d_a, slab_bytes));
    CUDA_TRY(p, cudaMalloc(&c->d_b, slab_bytes));
    const cufftType ty = p->dtype == NFFTB200_F32 ? CUFFT_C2C : CUFFT_Z2Z;
    size_t ws = 0;
    {   // local FFT over dims 1..D-1 of the z-slab [Os][mid][inner]: rank D-1, batch Os, contiguous
        long long n[2];
        for (int d = 0; d < D - 1; d++) n[d] = p->Nt[D - 2 - d];
        const long long dist = c->mid * c->inner;
        CUFFT_TRY(p, cufftCreate(&c->fft_local)); c->have_local = true;
        CUFFT_TRY(p, cufftMakePlanMany64(c->fft_local, D - 1, n, nullptr, 1, dist, nullptr, 1, dist, ty, c->Os, &ws));
        CUFFT_TRY(p, cufftSetStream(c->fft_local, p->stream));
    }
    {   // FFT over the last dim on the transposed slab [outer][Ms][inner]: stride Ms*inner, batch Ms*inner
        long long n[1] = {c->outer};
        long long emb[1] = {c->outer};
        const long long stride = c->Ms * c->inner;
        CUFFT_TRY(p, cufftCreate(&c->fft_last)); c->have_last = true;
        CUFFT_TRY(p, cufftMakePlanMany64(c->fft_last, 1, n, emb, stride, 1, emb, stride, 1, ty, stride, &ws));
        CUFFT_TRY(p, cufftSetStream(c->fft_last, p->stream));
    }
    c->slab = true;
    return NFFTB200_OK;
}

}  // namespace

// tile-aligned cut of the sorted node list into nranks ranges of ~M/nranks nodes (host logic, also used by
// the CPU tests): out[r] .. out[r+1] is rank r's half-open tile range
extern "C" int nfftb200_partition_tiles(const int64_t* tile_start, int64_t ntiles, int nranks, int64_t* out)
{
    if (!tile_start || !out || nranks < 1 || ntiles < 0) return NFFTB200_BAD_ARGUMENT;
    const int64_t M = tile_start[ntiles];
    out[0] = 0;
    for (int r = 1; r < nranks; r++) {
        const int64_t target = (M * r) / nranks;
        const int64_t* it = std::lower_bound(tile_start, tile_start + ntiles + 1, target);
        int64_t t = it - tile_start;
        if (t > ntiles) t = ntiles;
        out[r] = std::max(out[r - 1], t);
    }
    out[nranks] = ntiles;
    return NFFTB200_OK;
}

// nfftb200_set_stream: the slab FFT plans follow the plan's stream
void nfftb_comm_set_stream(nfftb200_plan* p)
{
    CommState* c = cs(p);
    if (!c) return;
    if (c->have_local) cufftSetStream(c->fft_local, p->stream);
    if (c->have_last) cufftSetStream(c->fft_last, p->stream);
}

void nfftb_comm_destroy(nfftb200_plan* p)
{
    CommState* c = cs(p);
    if (!c) return;
    if (c->have_local) cufftDestroy(c->fft_local);
    if (c->have_last) cufftDestroy(c->fft_last);
    if (c->d_a) cudaFree(c->d_a);
    if (c->d_b) cudaFree(c->d_b);
    for (int r = 0; r < NFFTB_MAX_PEERS; r++) {
        if (c->peer_ptr[r]) cudaIpcCloseMemHandle(c->peer_ptr[r]);
        if (c->slab_ptr[r]) cudaIpcCloseMemHandle(c->slab_ptr[r]);
    }
    if (c->d_peerbuf) cudaFree(c->d_peerbuf);
    if (c->d_xchg) cudaFree(c->d_xchg);
    if (c->comm && api().ok) api().CommDestroy(c->comm);
    delete c;
    p->nccl_comm = nullptr;
}

static void own_tiles(nfftb200_plan* p, int64_t& t_lo, int64_t& t_hi)
{
    std::vector<int64_t> ts(p->h_tile_start.begin(), p->h_tile_start.end()), cut((size_t)p->nranks + 1);
    nfftb200_partition_tiles(ts.data(), p->ntiles, p->nranks, cut.data());
    t_lo = cut[(size_t)p->rank]; t_hi = cut[(size_t)p->rank + 1];
}

// all-gather of one PeerMsg per rank through the communicator (host-synchronous; plan-time only)
static int exchange_msgs(nfftb200_plan* p, const PeerMsg& mine, std::vector<PeerMsg>& msg)
{
    CommState* c = cs(p);
    const int P = p->nranks;
    msg.resize((size_t)P);
    if (!c->d_xchg) CUDA_TRY(p, cudaMalloc(&c->d_xchg, sizeof(PeerMsg) * NFFTB_MAX_PEERS));
    CUDA_TRY(p, cudaMemcpyAsync((char*)c->d_xchg + sizeof(PeerMsg) * p->rank, &mine, sizeof(PeerMsg), cudaMemcpyHostToDevice, p->stream));
    NCCL_TRY(p, api().AllGather((char*)c->d_xchg + sizeof(PeerMsg) * p->rank, c->d_xchg, sizeof(PeerMsg), ncclChar, c->comm, p->stream));
    CUDA_TRY(p, cudaMemcpyAsync(msg.data(), c->d_xchg, sizeof(PeerMsg) * P, cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(p, cudaStreamSynchronize(p->stream));
    return NFFTB200_OK;
}

static bool all_ok(const std::vector<PeerMsg>& msg)
{
    for (const PeerMsg& m : msg) if (!m.ok) return false;
    return true;
}

// Collective (comm_init): export the z-slab work buffer and map everyone else's, so that the forward transform can
// interpolate straight from the slabs.
static int export_slabs(nfftb200_plan* p)
{
    CommState* c = cs(p);
    const int P = p->nranks;
    c->fused_fwd = false;
    if (P > NFFTB_MAX_PEERS || !c->slab || p->D != 3) return NFFTB200_OK;        // same decision on every rank
    std::vector<PeerMsg> msg;
    PeerMsg mine{};
    mine.ok = cudaIpcGetMemHandle(&mine.h, c->d_a) == cudaSuccess ? 1 : 0;
    if (!mine.ok) cudaGetLastError();
    ST_TRY(exchange_msgs(p, mine, msg));
    if (!all_ok(msg)) return NFFTB200_OK;
    const std::vector<PeerMsg> handles = msg;
    mine.ok = 1;
    for (int r = 0; r < P && mine.ok; r++) {
        if (r == p->rank) continue;
        if (cudaIpcOpenMemHandle(&c->slab_ptr[r], handles[(size_t)r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError(); c->slab_ptr[r] = nullptr; mine.ok = 0;
        }
    }
    ST_TRY(exchange_msgs(p, mine, msg));
    if (!all_ok(msg)) return NFFTB200_OK;
    c->slabs.n = P;
    c->slabs.planes = (int)c->Os;
    for (int r = 0; r < P; r++) c->slabs.base[r] = r == p->rank ? c->d_a : c->slab_ptr[r];
    c->fused_fwd = true;
    return NFFTB200_OK;
}

// Collective (every rank calls it from nodes!): size the own tile scratch for the new node set, export it and map
// the scratch of all other ranks.  Any rank that cannot take part switches the whole group to the NCCL path.
int nfftb_comm_after_nodes(nfftb200_plan* p)
{
    CommState* c = cs(p);
    if (!c || !c->comm || p->shard_mode != NFFTB200_SHARD_NODES) return NFFTB200_OK;
    const int P = p->nranks;
    c->fused = false;
    if (P > NFFTB_MAX_PEERS) return NFFTB200_OK;                          // same decision on every rank
    CUDA_TRY(p, cudaStreamSynchronize(p->stream));
    for (int r = 0; r < NFFTB_MAX_PEERS; r++)
        if (c->peer_ptr[r]) { cudaIpcCloseMemHandle(c->peer_ptr[r]); c->peer_ptr[r] = nullptr; }
    std::vector<PeerMsg> msg;
    auto exchange = [&](const PeerMsg& mine) -> int { return exchange_msgs(p, mine, msg); };
    PeerMsg mine{};
    // round 1: everyone has closed its mappings (so buffers may be re-allocated) and says whether the path applies
    const int PN = nfftb_peer_tile_cells(p);
    mine.ok = (PN > 0 && c->slab && c->Os % 16 == 0) ? 1 : 0;
    ST_TRY(exchange(mine));
    bool all = true;
    for (int r = 0; r < P; r++) all = all && msg[(size_t)r].ok;
    if (!all) return NFFTB200_OK;
    // round 2: allocate, export
    std::vector<int64_t> ts(p->h_tile_start.begin(), p->h_tile_start.end()), cut((size_t)P + 1);
    nfftb200_partition_tiles(ts.data(), p->ntiles, P, cut.data());
    c->tab.n = P;
    for (int r = 0; r <= P; r++) c->tab.cut[r] = (int)cut[(size_t)r];
    for (int r = 0; r < P; r++) c->tab.item_lo[r] = p->h_tile_items[(size_t)cut[(size_t)r]];
    const int64_t items = p->h_tile_items[(size_t)cut[(size_t)p->rank + 1]] - p->h_tile_items[(size_t)cut[(size_t)p->rank]];
    const int64_t need = std::max<int64_t>(items, 1) * PN * 2 * (int64_t)p->esz();
    mine.ok = 1;
    if (need > c->cap_peerbuf) {
        if (c->d_peerbuf) cudaFree(c->d_peerbuf);
        c->d_peerbuf = nullptr; c->cap_peerbuf = 0;
        if (cudaMalloc(&c->d_peerbuf, (size_t)need) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; }
        else c->cap_peerbuf = need;
    }
    if (mine.ok && cudaIpcGetMemHandle(&mine.h, c->d_peerbuf) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; }
    ST_TRY(exchange(mine));
    for (int r = 0; r < P; r++) all = all && msg[(size_t)r].ok;
    if (!all) return NFFTB200_OK;
    // round 3: map, agree
    const std::vector<PeerMsg> handles = msg;
    mine.ok = 1;
    for (int r = 0; r < P && mine.ok; r++) {
        if (r == p->rank) continue;
        if (cudaIpcOpenMemHandle(&c->peer_ptr[r], handles[(size_t)r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError(); c->peer_ptr[r] = nullptr; mine.ok = 0;
        }
    }
    ST_TRY(exchange(mine));
    for (int r = 0; r < P; r++) all = all && msg[(size_t)r].ok;
    if (!all) return NFFTB200_OK;
    for (int r = 0; r < P; r++) c->tab.base[r] = r == p->rank ? c->d_peerbuf : c->peer_ptr[r];
    c->fused = true;
    return NFFTB200_OK;
}

int nfftb_slab_deconvolve(nfftb200_plan* p, const void* d_f, void* d_slab, int64_t mid_off, int64_t Ms);            // deconv.cu
int nfftb_slab_deconvolve_transpose(nfftb200_plan* p, const void* d_slab, void* d_f, int64_t mid_off, int64_t Ms);

int nfftb_comm_exec_adjoint(nfftb200_plan* p, const void* d_fhat, void* d_f)
{
    CommState* c = cs(p);
    if (!c || !c->comm) return nfftb_fail(p, NFFTB200_NCCL_ERROR, "node sharding not initialised (call nfftb200_comm_init)");
    const int P = p->nranks;
    const size_t csz = 2 * p->esz();
    const size_t slab_reals = (size_t)(p->gsz / P) * 2;
    int64_t t_lo, t_hi;
    own_tiles(p, t_lo, t_hi);
    if (p->timing) cudaEventRecord(p->ev[0], p->stream);
    const bool fused = c->fused && c->slab && p->kernel_mode != 6 && p->kernel_mode != 1 && p->kernel_mode != 2;
    if (fused) {
        // spread the own tiles into the exported scratch; the barrier orders every rank's spread before any gather
        ST_TRY(nfftb_peer_spread(p, d_fhat, c->d_peerbuf, t_lo, t_hi));
        NCCL_TRY(p, api().AllReduce(c->d_xchg, c->d_xchg, 1, ncclFloat, ncclSum, c->comm, p->stream));
        ST_TRY(nfftb_peer_gather(p, c->d_a, (int)(p->rank * c->Os / 16), (int)(c->Os / 16), c->tab));
        p->launches++;
    } else {
        ST_TRY(nfftb_spread(p, d_fhat, p->d_grid, 1, 1, t_lo, t_hi));
    }
    if (p->timing) cudaEventRecord(p->ev[1], p->stream);
    if (c->slab) {
        if (!fused) {
            // reduce-scatter over slabs of the slowest grid dimension
            NCCL_TRY(p, api().ReduceScatter(p->d_grid, c->d_a, slab_reals, nccl_real(p), ncclSum, c->comm, p->stream));
            p->launches++;
        }
        ST_TRY(exec_fft(p, c->fft_local, c->d_a, +1));
        const long long n = p->gsz / P;
        const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 32);
        if (p->dtype == NFFTB200_F32) k_pack_blocks<float2><<<blocks, 256, 0, p->stream>>>((const float2*)c->d_a, (float2*)c->d_b, c->inner, c->Ms, c->Os, P, false);
        else k_pack_blocks<double2><<<blocks, 256, 0, p->stream>>>((const double2*)c->d_a, (double2*)c->d_b, c->inner, c->Ms, c->Os, P, false);
        p->launches++;
        ST_TRY(all_to_all(p, c->d_b, c->d_a, slab_reals / P));       // d_a is now [outer][Ms][inner]
        ST_TRY(exec_fft(p, c->fft_last, c->d_a, +1));
        if (p->timing) cudaEventRecord(p->ev[2], p->stream);
        CUDA_TRY(p, cudaMemsetAsync(d_f, 0, (size_t)p->fsz * csz, p->stream));
        ST_TRY(nfftb_slab_deconvolve_transpose(p, c->d_a, d_f, (int64_t)p->rank * c->Ms, c->Ms));
        NCCL_TRY(p, api().AllReduce(d_f, d_f, (size_t)p->fsz * 2, nccl_real(p), ncclSum, c->comm, p->stream));
        p->launches += 2;
    } else {
        char* mine = (char*)p->d_grid + (size_t)p->rank * slab_reals * p->esz();
        NCCL_TRY(p, api().ReduceScatter(p->d_grid, mine, slab_reals, nccl_real(p), ncclSum, c->comm, p->stream));
        NCCL_TRY(p, api().AllGather(mine, p->d_grid, slab_reals, nccl_real(p), c->comm, p->stream));
        p->launches += 2;
        ST_TRY(plan_fft_replicated(p, p->d_grid, +1));
        if (p->timing) cudaEventRecord(p->ev[2], p->stream);
        ST_TRY(nfftb_deconvolve_transpose(p, p->d_grid, d_f, 1));
    }
    if (p->timing) { cudaEventRecord(p->ev[3], p->stream); p->pending = 2; }
    return NFFTB200_OK;
}

int nfftb_comm_exec_forward(nfftb200_plan* p, const void* d_f, void* d_fhat)
{
    CommState* c = cs(p);
    if (!c || !c->comm) return nfftb_fail(p, NFFTB200_NCCL_ERROR, "node sharding not initialised (call nfftb200_comm_init)");
    const int P = p->nranks;
    const size_t slab_reals = (size_t)(p->gsz / P) * 2;
    int64_t t_lo, t_hi;
    own_tiles(p, t_lo, t_hi);
    if (p->timing) cudaEventRecord(p->ev[0], p->stream);
    if (c->slab) {
        ST_TRY(nfftb_slab_deconvolve(p, d_f, c->d_a, (int64_t)p->rank * c->Ms, c->Ms));   // [outer][Ms][inner]
        if (p->timing) cudaEventRecord(p->ev[1], p->stream);
        ST_TRY(exec_fft(p, c->fft_last, c->d_a, -1));
        ST_TRY(all_to_all(p, c->d_a, c->d_b, slab_reals / P));       // blocks [q][Os][Ms][inner]
        const long long n = p->gsz / P;
        const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 32);
        if (p->dtype == NFFTB200_F32) k_pack_blocks<float2><<<blocks, 256, 0, p->stream>>>((const float2*)c->d_b, (float2*)c->d_a, c->inner, c->Ms, c->Os, P, true);
        else k_pack_blocks<double2><<<blocks, 256, 0, p->stream>>>((const double2*)c->d_b, (double2*)c->d_a, c->inner, c->Ms, c->Os, P, true);
        p->launches++;
        ST_TRY(exec_fft(p, c->fft_local, c->d_a, -1));
        if (c->fused_fwd && p->kernel_mode != 6 && p->kernel_mode != 1) {
            // no "broadcast of the grid": after a barrier (every slab is final) the own tiles are staged plane by
            // plane from the slab that holds them; the trailing barrier keeps the slabs alive until every rank is done
            NCCL_TRY(p, api().AllReduce(c->d_xchg, c->d_xchg, 1, ncclFloat, ncclSum, c->comm, p->stream));
            if (p->timing) cudaEventRecord(p->ev[2], p->stream);
            CUDA_TRY(p, cudaMemsetAsync(d_fhat, 0, (size_t)p->M * 2 * p->esz(), p->stream));
            const int r = nfftb_peer_interp(p, c->slabs, d_fhat, t_lo, t_hi);
            if (r >= 0) {
                ST_TRY(r);
                NCCL_TRY(p, api().AllReduce(c->d_xchg, c->d_xchg, 1, ncclFloat, ncclSum, c->comm, p->stream));
                p->launches += 3;
                if (p->timing) { cudaEventRecord(p->ev[3], p->stream); p->pending = 1; }
                return NFFTB200_OK;
            }
        }
        // "broadcast the grid": all-gather of the z-slabs
        NCCL_TRY(p, api().AllGather(c->d_a, p->d_grid, slab_reals, nccl_real(p), c->comm, p->stream));
        p->launches++;
    } else {
        ST_TRY(nfftb_deconvolve(p, d_f, p->d_grid, 1));
        if (p->timing) cudaEventRecord(p->ev[1], p->stream);
        ST_TRY(plan_fft_replicated(p, p->d_grid, -1));
    }
    if (p->timing) cudaEventRecord(p->ev[2], p->stream);
    // entries of nodes owned by other ranks are defined to be zero (a sum all-reduce assembles the full vector)
    CUDA_TRY(p, cudaMemsetAsync(d_fhat, 0, (size_t)p->M * 2 * p->esz(), p->stream));
    ST_TRY(nfftb_interp(p, p->d_grid, d_fhat, 1, 1, t_lo, t_hi));
    if (p->timing) { cudaEventRecord(p->ev[3], p->stream); p->pending = 1; }
    return NFFTB200_OK;
}

extern "C" {

int nfftb200_comm_unique_id(void* out128)
{
    if (!out128) return NFFTB200_BAD_ARGUMENT;
    if (!api().ok) return nfftb_fail(nullptr, NFFTB200_NCCL_ERROR, "libnccl.so.2 could not be loaded");
    ncclUniqueId id;
    if (api().GetUniqueId(&id) != ncclSuccess) return nfftb_fail(nullptr, NFFTB200_NCCL_ERROR, "ncclGetUniqueId failed");
    std::memcpy(out128, &id, sizeof(id));
    return NFFTB200_OK;
}

int nfftb200_comm_is_fused(nfftb200_plan* p)
{
    CommState* c = p ? cs(p) : nullptr;
    if (!c || p->shard_mode != NFFTB200_SHARD_NODES) return 0;
    return (c->fused ? 1 : 0) | (c->fused_fwd ? 2 : 0);
}

int nfftb200_comm_init(nfftb200_plan* p, const void* nccl_unique_id, int rank, int nranks, int mode)
{
    if (!p) return NFFTB200_BAD_ARGUMENT;
    if (nranks < 1 || rank < 0 || rank >= nranks) return nfftb_fail(p, NFFTB200_BAD_ARGUMENT, "bad rank / nranks");
    if (mode == NFFTB200_SHARD_NONE || nranks == 1) { p->shard_mode = NFFTB200_SHARD_NONE; p->rank = 0; p->nranks = 1; return NFFTB200_OK; }
    if (mode == NFFTB200_SHARD_BATCH) {       // no collective on the data path: only bookkeeping
        p->rank = rank; p->nranks = nranks; p->shard_mode = NFFTB200_SHARD_BATCH;
        return NFFTB200_OK;
    }
    if (mode != NFFTB200_SHARD_NODES) return nfftb_fail(p, NFFTB200_BAD_ARGUMENT, "unknown sharding mode");
    if (p->device < 0) return nfftb_fail(p, NFFTB200_CUDA_ERROR, "host-only plan (device < 0): no CPU fallback exists");
    if (p->B != 1) return nfftb_fail(p, NFFTB200_UNSUPPORTED, "node sharding needs ntransforms == 1");
    if (p->D > 3) return nfftb_fail(p, NFFTB200_UNSUPPORTED, "node sharding: only D = 1, 2, 3 are supported");
    if (p->gsz % nranks) return nfftb_fail(p, NFFTB200_UNSUPPORTED, "grid size must be divisible by the number of ranks");
    if (!nccl_unique_id) return nfftb_fail(p, NFFTB200_BAD_ARGUMENT, "nccl_unique_id == NULL");
    if (!api().ok) return nfftb_fail(p, NFFTB200_NCCL_ERROR, "libnccl.so.2 could not be loaded");
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(p->device);
    nfftb_comm_destroy(p);
    CommState* c = new CommState();
    p->nccl_comm = c;
    ncclUniqueId id;
    std::memcpy(&id, nccl_unique_id, sizeof(id));
    p->rank = rank; p->nranks = nranks;
    NCCL_TRY(p, api().CommInitRank(&c->comm, nranks, id, rank));
    int st = setup_slab(p);
    if (st == NFFTB200_OK) st = export_slabs(p);
    if (st == NFFTB200_OK) {
        p->shard_mode = NFFTB200_SHARD_NODES;
        if (p->have_nodes) st = nfftb_comm_after_nodes(p);
    }
    if (prev >= 0 && prev != p->device) cudaSetDevice(prev);
    return st;
}

}  // extern "C"
