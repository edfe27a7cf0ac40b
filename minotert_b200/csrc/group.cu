// group.cu -- multi-GPU groups behind the C ABI (SURVEY.md 8b "mrt_group_*", 8e).
//
// The render never communicates: every rank holds the whole scene (replicated BVH) and renders either its own row
// slabs of the image (tile mode, mrt_set_partition) or whole frames for its own frame counters (sample sets).  The
// one exchange step is of the finished per-pixel buffer:
//     mrt_group_gather : tile mode  -- every rank's compact slabs travel to the root (NCCL send/recv, grouped) into a
//                        staging area and k_scatter_rows (ours) puts the rows back into image order;
//     mrt_group_reduce : sample sets -- ncclReduce(sum) of the fp32 accumulators onto the root.
// Two ways to form a group: one process driving n devices (ncclCommInitAll) or one process per GPU with a
// caller-distributed ncclUniqueId (ncclCommInitRank; bench.py hands the 128 bytes round with torch.distributed).
// NCCL is bound at run time (dlopen "libnccl.so.2": inside a PyTorch process that is the copy torch already loaded, so
// two NCCL builds never meet in one address space); a group whose transport is MRT_GROUP_P2P uses plain device-to-device
// copies instead (single process only; also the NCCL-free baseline SURVEY 5 names, and it allows several contexts of
// one device -- how the N-rank tile logic is tested on a single GPU).
#include <dlfcn.h>
#include <nccl.h>
#include <stdarg.h>

#include <vector>

#include "context.cuh"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi g_nccl;
char g_group_error[512] = "";

const char* nccl_load() {
    if (g_nccl.handle) return nullptr;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return "libnccl.so.2 not found (dlopen)";
#define SYM(field, name)                                               \
    *(void**)(&g_nccl.field) = dlsym(h, name);                         \
    if (!g_nccl.field) { dlclose(h); return "libnccl: symbol " name " missing"; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommInitAll, "ncclCommInitAll")
    SYM(CommDestroy, "ncclCommDestroy") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
    SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(Reduce, "ncclReduce") SYM(GetErrorString, "ncclGetErrorString")
    SYM(GetVersion, "ncclGetVersion")
#undef SYM
    g_nccl.handle = h;
    return nullptr;
}

// rows of `bytes_per_row` bytes: staging holds rank r's rows compactly from row offset `first`; row k of that block
// belongs at image row rows[first + k]
template <typename T>
__global__ void __launch_bounds__(256) k_scatter_rows(const T* __restrict__ staging, T* __restrict__ full,
                                                      const uint32_t* __restrict__ rows, uint32_t nrows, uint32_t row_elems) {
    const size_t total = (size_t)nrows * row_elems;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t k = (uint32_t)(i / row_elems), c = (uint32_t)(i - (size_t)k * row_elems);
        full[(size_t)__ldg(&rows[k]) * row_elems + c] = staging[i];
    }
}

uint32_t buffer_bpp(int id) {
    switch (id) {
    case MRT_BUF_VISIBILITY: case MRT_BUF_MOTION: case MRT_BUF_LDR: case MRT_BUF_HIT_T: case MRT_BUF_DENOISED: return 4;
    case MRT_BUF_DEPTH: return 2;
    case MRT_BUF_NORMAL: case MRT_BUF_COLOR: return 8;
    case MRT_BUF_ACCUM: return 16;
    default: return 0;
    }
}

}  // namespace

struct mrt_group {
    uint32_t nranks = 0;                 // ranks of the whole group
    int transport = MRT_GROUP_NCCL;
    bool multi_process = false;
    std::vector<mrt_context*> ctx;       // local contexts ...
    std::vector<uint32_t> rank;          // ... and their ranks
    std::vector<ncclComm_t> comm;        // one per local context (NCCL transport)
    std::vector<cudaStream_t> comm_stream;  // exchange stream of each local context (ordered by events)
    std::vector<cudaEvent_t> ready, done;
    // frames in flight (mrt_group_set_frames_in_flight): frame contexts of each local rank, used round-robin by renders
    // that carry MRT_SECONDARY_FRAME_SUM and committed to ctx[i] in call order; empty: ctx[i] renders itself
    std::vector<std::vector<mrt_context*>> fctx;
    uint32_t frames_in_flight = 1;
    uint64_t frame_index = 0;
    std::vector<cudaEvent_t> copies;     // outstanding mrt_group_readback_async copies, oldest first
    uint32_t slab_rows = 0;              // 0: no tile partition set
    // root side (allocated on the root's device on first use)
    int root_local = -1;                 // index into ctx of the root of the last gather, -1 if the root is remote
    DevArray<unsigned char> staging, full;
    DevArray<uint32_t> row_table;        // image row of every staged row, rank after rank
    uint32_t table_w = 0, table_h = 0, table_slab = 0;
    std::vector<uint32_t> rank_rows, rank_first;  // rows owned by each rank, and their offset in the staging area
    size_t full_bytes = 0;
    bool gather_pending = false;
    char err[512] = {0};
};

static int group_fail(mrt_group* g, int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g ? g->err : g_group_error, 512, fmt, ap);
    va_end(ap);
    return code;
}
#define GRP_NCCL(g, call)                                                                                   \
    do {                                                                                                    \
        ncclResult_t _r = (call);                                                                           \
        if (_r != ncclSuccess) return group_fail((g), MRT_ERR_CUDA, "%s: %s", #call, g_nccl.GetErrorString(_r)); \
    } while (0)
#define GRP_CUDA(g, call)                                                                              \
    do {                                                                                               \
        cudaError_t _e = (call);                                                                       \
        if (_e != cudaSuccess) return group_fail((g), MRT_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(_e)); \
    } while (0)
#define GRP_CTX(g, i, call)                                                                                       \
    do {                                                                                                          \
        int _s = (call);                                                                                          \
        if (_s != MRT_OK) return group_fail((g), _s, "rank %u: %s", (g)->rank[i], mrt_last_error((g)->ctx[i]));   \
    } while (0)

static int group_add_streams(mrt_group* g) {
    for (size_t i = 0; i < g->ctx.size(); i++) {
        GRP_CUDA(g, cudaSetDevice(g->ctx[i]->device));
        cudaStream_t s;
        cudaEvent_t a, b;
        GRP_CUDA(g, cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        GRP_CUDA(g, cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        GRP_CUDA(g, cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        g->comm_stream.push_back(s);
        g->ready.push_back(a);
        g->done.push_back(b);
    }
    return MRT_OK;
}

// partition tables for a (w, h, slab) image: rows per rank, staging offsets, device row table on the root
static int group_tables(mrt_group* g, mrt_context* root, uint32_t w, uint32_t h) {
    if (g->table_w == w && g->table_h == h && g->table_slab == g->slab_rows && g->row_table.p) return MRT_OK;
    std::vector<uint32_t> table;
    g->rank_rows.assign(g->nranks, 0);
    g->rank_first.assign(g->nranks, 0);
    for (uint32_t r = 0; r < g->nranks; r++) {
        Partition p{r, g->nranks, g->slab_rows};
        const uint32_t n = partition_local_rows(p, h);
        g->rank_first[r] = (uint32_t)table.size();
        g->rank_rows[r] = n;
        for (uint32_t lr = 0; lr < n; lr++) table.push_back(partition_local_to_y(p, lr));
    }
    if (table.size() != h) return group_fail(g, MRT_ERR_INVALID, "partition tables cover %zu of %u rows", table.size(), h);
    GRP_CUDA(g, cudaSetDevice(root->device));
    if (dev_reserve(root, g->row_table, h) != MRT_OK) return group_fail(g, MRT_ERR_OOM, "%s", root->err);
    GRP_CUDA(g, cudaMemcpyAsync(g->row_table.p, table.data(), sizeof(uint32_t) * h, cudaMemcpyHostToDevice, root->stream));
    GRP_CUDA(g, cudaStreamSynchronize(root->stream));
    g->table_w = w; g->table_h = h; g->table_slab = g->slab_rows;
    return MRT_OK;
}

extern "C" {

const char* mrt_group_last_error(const mrt_group* g) { return g ? g->err : g_group_error; }

int mrt_group_unique_id(uint8_t id_out[128]) {
    if (!id_out) return MRT_ERR_INVALID;
    if (const char* e = nccl_load()) return group_fail(nullptr, MRT_ERR_STATE, "%s", e);
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    GRP_NCCL(nullptr, g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, 128);
    return MRT_OK;
}

int mrt_group_create(const int* devices, uint32_t ndev, int transport, mrt_group** out) {
    if (!out) return MRT_ERR_INVALID;
    *out = nullptr;
    if (!devices || ndev == 0 || ndev > 64) return group_fail(nullptr, MRT_ERR_INVALID, "mrt_group_create: %u devices", ndev);
    if (transport != MRT_GROUP_NCCL && transport != MRT_GROUP_P2P) return group_fail(nullptr, MRT_ERR_INVALID, "unknown transport %d", transport);
    if (transport == MRT_GROUP_NCCL) {
        if (const char* e = nccl_load()) return group_fail(nullptr, MRT_ERR_STATE, "%s", e);
        for (uint32_t i = 0; i < ndev; i++)
            for (uint32_t j = 0; j < i; j++)
                if (devices[i] == devices[j])
                    return group_fail(nullptr, MRT_ERR_INVALID, "NCCL transport: device %d appears twice (use MRT_GROUP_P2P for several contexts of one GPU)", devices[i]);
    }
    mrt_group* g = new mrt_group();
    g->nranks = ndev;
    g->transport = transport;
    for (uint32_t i = 0; i < ndev; i++) {
        mrt_context* c = nullptr;
        int s = mrt_create(devices[i], &c);
        if (s != MRT_OK) {
            group_fail(nullptr, s, "mrt_group_create: device %d: %s", devices[i], mrt_last_error(nullptr));
            for (mrt_context* k : g->ctx) mrt_destroy(k);
            delete g;
            return s;
        }
        g->ctx.push_back(c);
        g->rank.push_back(i);
    }
    int s = group_add_streams(g);
    if (s == MRT_OK && transport == MRT_GROUP_NCCL) {
        g->comm.resize(ndev);
        ncclResult_t r = g_nccl.CommInitAll(g->comm.data(), (int)ndev, devices);
        if (r != ncclSuccess) { g->comm.clear(); s = group_fail(nullptr, MRT_ERR_CUDA, "ncclCommInitAll: %s", g_nccl.GetErrorString(r)); }
    }
    if (s == MRT_OK && transport == MRT_GROUP_P2P)
        for (uint32_t i = 0; i < ndev; i++)
            for (uint32_t j = 0; j < ndev; j++)
                if (devices[i] != devices[j]) {
                    int can = 0;
                    cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
                    if (can) { cudaSetDevice(devices[i]); cudaDeviceEnablePeerAccess(devices[j], 0); cudaGetLastError(); }
                }
    if (s != MRT_OK) {
        memcpy(g_group_error, g->err[0] ? g->err : g_group_error, 512);
        mrt_group_destroy(g);
        return s;
    }
    *out = g;
    return MRT_OK;
}

int mrt_group_create_rank(int device, uint32_t rank, uint32_t nranks, const uint8_t unique_id[128], mrt_group** out) {
    if (!out) return MRT_ERR_INVALID;
    *out = nullptr;
    if (!unique_id || nranks == 0 || rank >= nranks) return group_fail(nullptr, MRT_ERR_INVALID, "mrt_group_create_rank: rank %u of %u", rank, nranks);
    if (const char* e = nccl_load()) return group_fail(nullptr, MRT_ERR_STATE, "%s", e);
    mrt_context* c = nullptr;
    int s = mrt_create(device, &c);
    if (s != MRT_OK) return group_fail(nullptr, s, "mrt_group_create_rank: %s", mrt_last_error(nullptr));
    mrt_group* g = new mrt_group();
    g->nranks = nranks;
    g->multi_process = true;
    g->ctx.push_back(c);
    g->rank.push_back(rank);
    s = group_add_streams(g);
    if (s == MRT_OK) {
        ncclUniqueId id;
        memcpy(&id, unique_id, 128);
        g->comm.resize(1);
        ncclResult_t r = g_nccl.CommInitRank(&g->comm[0], (int)nranks, id, (int)rank);
        if (r != ncclSuccess) { g->comm.clear(); s = group_fail(nullptr, MRT_ERR_CUDA, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); }
    }
    if (s != MRT_OK) { mrt_group_destroy(g); return s; }
    *out = g;
    return MRT_OK;
}

void mrt_group_destroy(mrt_group* g) {
    if (!g) return;
    for (size_t i = 0; i < g->ctx.size(); i++) {
        cudaSetDevice(g->ctx[i]->device);
        if (i < g->comm_stream.size()) { cudaStreamSynchronize(g->comm_stream[i]); }
        cudaStreamSynchronize(g->ctx[i]->stream);
    }
    for (auto& slots : g->fctx)  // the frame contexts borrow the scenes of the ranks' contexts: they go first
        for (mrt_context* c : slots) mrt_destroy(c);
    g->fctx.clear();
    for (cudaEvent_t e : g->copies) cudaEventDestroy(e);
    g->copies.clear();
    for (ncclComm_t c : g->comm) g_nccl.CommDestroy(c);
    if (g->root_local >= 0 || g->staging.p || g->full.p || g->row_table.p) {
        // root-side buffers live on the device of the context that was root
        if (g->root_local >= 0) cudaSetDevice(g->ctx[g->root_local]->device);
        dev_free(g->staging); dev_free(g->full); dev_free(g->row_table);
    }
    for (size_t i = 0; i < g->ctx.size(); i++) {
        cudaSetDevice(g->ctx[i]->device);
        if (i < g->comm_stream.size()) cudaStreamDestroy(g->comm_stream[i]);
        if (i < g->ready.size()) cudaEventDestroy(g->ready[i]);
        if (i < g->done.size()) cudaEventDestroy(g->done[i]);
        mrt_destroy(g->ctx[i]);
    }
    delete g;
}

int mrt_group_size(const mrt_group* g, uint32_t* nranks, uint32_t* nlocal) {
    if (!g) return MRT_ERR_INVALID;
    if (nranks) *nranks = g->nranks;
    if (nlocal) *nlocal = (uint32_t)g->ctx.size();
    return MRT_OK;
}

int mrt_group_context(mrt_group* g, uint32_t local_index, mrt_context** ctx_out, uint32_t* rank_out) {
    if (!g || local_index >= g->ctx.size() || !ctx_out) return MRT_ERR_INVALID;
    *ctx_out = g->ctx[local_index];
    if (rank_out) *rank_out = g->rank[local_index];
    return MRT_OK;
}

int mrt_group_set_tiles(mrt_group* g, uint32_t slab_rows) {
    if (!g) return MRT_ERR_INVALID;
    if (slab_rows == 0) return group_fail(g, MRT_ERR_INVALID, "slab_rows = 0");
    for (size_t i = 0; i < g->ctx.size(); i++) {
        GRP_CTX(g, i, mrt_set_partition(g->ctx[i], g->rank[i], g->nranks, slab_rows));
        if (i < g->fctx.size())
            for (mrt_context* c : g->fctx[i])
                if (mrt_set_partition(c, g->rank[i], g->nranks, slab_rows) != MRT_OK)
                    return group_fail(g, MRT_ERR_INVALID, "rank %u frame context: %s", g->rank[i], mrt_last_error(c));
    }
    g->slab_rows = slab_rows;
    return MRT_OK;
}

int mrt_group_set_frames_in_flight(mrt_group* g, uint32_t frames) {
    if (!g) return MRT_ERR_INVALID;
    if (frames < 1 || frames > MRT_GROUP_MAX_FRAMES) return group_fail(g, MRT_ERR_INVALID, "frames in flight: %u (1..%d)", frames, MRT_GROUP_MAX_FRAMES);
    if (int s0 = mrt_group_sync(g)) return s0;
    if (frames == 1) {
        for (auto& slots : g->fctx)
            for (mrt_context* c : slots) mrt_destroy(c);
        g->fctx.clear();
        for (size_t i = 0; i < g->ctx.size(); i++) GRP_CTX(g, i, mrt_set_option(g->ctx[i], "trace_ctas_per_sm", 0));
    } else {
        g->fctx.resize(g->ctx.size());
        for (size_t i = 0; i < g->ctx.size(); i++) {
            while (g->fctx[i].size() > frames) { mrt_destroy(g->fctx[i].back()); g->fctx[i].pop_back(); }
            while (g->fctx[i].size() < frames) {
                mrt_context* c = nullptr;
                int s = mrt_create(g->ctx[i]->device, &c);
                if (s != MRT_OK) return group_fail(g, s, "frame context on device %d: %s", g->ctx[i]->device, mrt_last_error(nullptr));
                g->fctx[i].push_back(c);
            }
            for (mrt_context* c : g->fctx[i]) {
                // co-running frames share the SMs: cap each frame's persistent traversal grid so that ~8 CTAs per SM are
                // offered in all (6 are resident; measured on 1/8-image slabs, tools/bench_slab_inflight.py)
                mrt_set_option(c, "trace_ctas_per_sm", (int)((8 + frames - 1) / frames));
                if (g->slab_rows) mrt_set_partition(c, g->rank[i], g->nranks, g->slab_rows);
            }
        }
    }
    g->frames_in_flight = frames;
    g->frame_index = 0;
    return MRT_OK;
}

int mrt_group_frame_context(mrt_group* g, uint32_t local_index, uint32_t slot, mrt_context** ctx_out) {
    if (!g || !ctx_out || local_index >= g->ctx.size()) return MRT_ERR_INVALID;
    if (g->frames_in_flight == 1) {
        if (slot != 0) return group_fail(g, MRT_ERR_INVALID, "frame context %u of 1", slot);
        *ctx_out = g->ctx[local_index];
        return MRT_OK;
    }
    if (slot >= g->fctx[local_index].size()) return group_fail(g, MRT_ERR_INVALID, "frame context %u of %zu", slot, g->fctx[local_index].size());
    *ctx_out = g->fctx[local_index][slot];
    return MRT_OK;
}

int mrt_group_render(mrt_group* g, uint32_t w, uint32_t h, const mrt_primary_constants* pc, const mrt_secondary_constants* sc,
                     uint32_t spp, uint32_t bounces, uint32_t flags, uint32_t frame_stride) {
    if (!g || !pc || !sc) return MRT_ERR_INVALID;
    const bool frame_sum = (flags & MRT_SECONDARY_FRAME_SUM) != 0;
    for (size_t i = 0; i < g->ctx.size(); i++) {
        mrt_primary_constants p = *pc;
        mrt_secondary_constants s = *sc;
        p.frameCounter += g->rank[i] * frame_stride;  // sample sets: rank r renders its own frame counters
        s.frameCounter += g->rank[i] * frame_stride;
        // frames in flight: this frame renders on the next frame context (its previous frame has been committed; the
        // commit below is stream-ordered behind the render and behind the commit of the frame before)
        mrt_context* c = (frame_sum && g->frames_in_flight > 1) ? g->fctx[i][g->frame_index % g->frames_in_flight] : g->ctx[i];
        if (mrt_primary_rays(c, w, h, &p) != MRT_OK || mrt_secondary_rays(c, &s, spp, bounces, flags) != MRT_OK)
            return group_fail(g, MRT_ERR_STATE, "rank %u: %s", g->rank[i], mrt_last_error(c));
        if (frame_sum) GRP_CTX(g, i, mrt_accum_commit(g->ctx[i], c, flags & MRT_SECONDARY_ACCUMULATE));
    }
    if (frame_sum) g->frame_index++;
    return MRT_OK;
}

int mrt_group_tonemap(mrt_group* g, int mode, float exposure, const float* params, uint32_t nparams, int source) {
    if (!g) return MRT_ERR_INVALID;
    for (size_t i = 0; i < g->ctx.size(); i++) GRP_CTX(g, i, mrt_tonemap(g->ctx[i], mode, exposure, params, nparams, source));
    return MRT_OK;
}

// Tile mode exchange.  Issued on each context's exchange stream behind an event on its render stream, so the next
// frame can be issued at once; MRT_BUF_LDR is double-buffered (the next tonemap writes the other half), for every other
// buffer the render stream is made to wait for the exchange.
int mrt_group_gather(mrt_group* g, int buffer_id, uint32_t root) {
    if (!g) return MRT_ERR_INVALID;
    if (root >= g->nranks) return group_fail(g, MRT_ERR_INVALID, "gather: root %u of %u", root, g->nranks);
    if (g->slab_rows == 0) return group_fail(g, MRT_ERR_STATE, "gather before mrt_group_set_tiles");
    const uint32_t bpp = buffer_bpp(buffer_id);
    if (!bpp) return group_fail(g, MRT_ERR_INVALID, "gather: buffer %d is not a per-pixel image", buffer_id);
    const size_t nloc = g->ctx.size();
    int root_local = -1;
    for (size_t i = 0; i < nloc; i++) if (g->rank[i] == root) root_local = (int)i;
    const uint32_t W = g->ctx[0]->W, H = g->ctx[0]->H;
    if (W == 0 || H == 0) return group_fail(g, MRT_ERR_STATE, "gather before the first render");
    const size_t row_bytes = (size_t)W * bpp;
    std::vector<void*> src(nloc);
    std::vector<size_t> src_bytes(nloc);
    for (size_t i = 0; i < nloc; i++) {
        GRP_CTX(g, i, mrt_buffer(g->ctx[i], buffer_id, &src[i], &src_bytes[i]));
        if (g->ctx[i]->W != W || g->ctx[i]->H != H) return group_fail(g, MRT_ERR_STATE, "gather: ranks rendered different image sizes");
    }
    mrt_context* rc = root_local >= 0 ? g->ctx[root_local] : nullptr;
    if (rc) {
        if (g->root_local >= 0 && g->root_local != root_local) {  // the root moved: its buffers live on another device
            cudaSetDevice(g->ctx[g->root_local]->device);
            dev_free(g->staging); dev_free(g->full); dev_free(g->row_table);
            g->table_w = 0;
        }
        g->root_local = root_local;
        MRT_TRY(group_tables(g, rc, W, H));
        GRP_CUDA(g, cudaSetDevice(rc->device));
        // staging and the full image are reused by every gather: with NCCL all their readers and writers (recv, scatter,
        // readback copies) are on the root's exchange stream, in order; with P2P the peers' copies are made to wait for the
        // root's previous scatter below.  Nothing blocks the host: frames in flight stay in flight.
        if (dev_reserve(rc, g->staging, row_bytes * H) != MRT_OK || dev_reserve(rc, g->full, row_bytes * H) != MRT_OK)
            return group_fail(g, MRT_ERR_OOM, "%s", rc->err);
        g->full_bytes = row_bytes * H;
    } else {
        // remote root: this process only needs its own row counts
        g->rank_rows.assign(g->nranks, 0);
        for (size_t i = 0; i < nloc; i++) g->rank_rows[g->rank[i]] = partition_local_rows(Partition{g->rank[i], g->nranks, g->slab_rows}, H);
    }
    for (size_t i = 0; i < nloc; i++) {
        GRP_CUDA(g, cudaSetDevice(g->ctx[i]->device));
        GRP_CUDA(g, cudaEventRecord(g->ready[i], g->ctx[i]->stream));
        GRP_CUDA(g, cudaStreamWaitEvent(g->comm_stream[i], g->ready[i], 0));
        if (src_bytes[i] != (size_t)g->rank_rows[g->rank[i]] * row_bytes)
            return group_fail(g, MRT_ERR_STATE, "gather: rank %u holds %zu bytes, its partition says %zu", g->rank[i], src_bytes[i],
                              (size_t)g->rank_rows[g->rank[i]] * row_bytes);
    }
    if (g->transport == MRT_GROUP_NCCL) {
        GRP_NCCL(g, g_nccl.GroupStart());
        for (size_t i = 0; i < nloc; i++) {
            if ((int)i == root_local) {
                for (uint32_t r = 0; r < g->nranks; r++)
                    if (r != root && g->rank_rows[r])
                        GRP_NCCL(g, g_nccl.Recv(g->staging.p + (size_t)g->rank_first[r] * row_bytes, (size_t)g->rank_rows[r] * row_bytes,
                                                ncclUint8, (int)r, g->comm[i], g->comm_stream[i]));
            } else if (src_bytes[i]) {
                GRP_NCCL(g, g_nccl.Send(src[i], src_bytes[i], ncclUint8, (int)root, g->comm[i], g->comm_stream[i]));
            }
        }
        GRP_NCCL(g, g_nccl.GroupEnd());
    } else {
        if (!rc) return group_fail(g, MRT_ERR_STATE, "P2P transport needs the root in this process");
        // peers' exchange streams copy straight into the root's staging area; the root's exchange stream waits for them
        for (size_t i = 0; i < nloc; i++) {
            if ((int)i == root_local || !src_bytes[i]) continue;
            GRP_CUDA(g, cudaSetDevice(g->ctx[i]->device));
            if (g->gather_pending) GRP_CUDA(g, cudaStreamWaitEvent(g->comm_stream[i], g->done[root_local], 0));  // previous scatter has read staging
            GRP_CUDA(g, cudaMemcpyPeerAsync(g->staging.p + (size_t)g->rank_first[g->rank[i]] * row_bytes, rc->device, src[i],
                                            g->ctx[i]->device, src_bytes[i], g->comm_stream[i]));
            GRP_CUDA(g, cudaEventRecord(g->done[i], g->comm_stream[i]));
            GRP_CUDA(g, cudaStreamWaitEvent(g->comm_stream[root_local], g->done[i], 0));
        }
    }
    if (rc) {
        GRP_CUDA(g, cudaSetDevice(rc->device));
        cudaStream_t cs = g->comm_stream[root_local];
        if (src_bytes[root_local])  // the root's own slabs
            GRP_CUDA(g, cudaMemcpyAsync(g->staging.p + (size_t)g->rank_first[root] * row_bytes, src[root_local], src_bytes[root_local],
                                        cudaMemcpyDeviceToDevice, cs));
        const unsigned grid = 148 * 8;
        if (row_bytes % 16 == 0)
            k_scatter_rows<uint4><<<grid, 256, 0, cs>>>((const uint4*)g->staging.p, (uint4*)g->full.p, g->row_table.p, H, (uint32_t)(row_bytes / 16));
        else if (row_bytes % 4 == 0)
            k_scatter_rows<uint32_t><<<grid, 256, 0, cs>>>((const uint32_t*)g->staging.p, (uint32_t*)g->full.p, g->row_table.p, H, (uint32_t)(row_bytes / 4));
        else
            k_scatter_rows<uint16_t><<<grid, 256, 0, cs>>>((const uint16_t*)g->staging.p, (uint16_t*)g->full.p, g->row_table.p, H, (uint32_t)(row_bytes / 2));
        MRT_LAUNCHED(rc);
        GRP_CUDA(g, cudaGetLastError());
        g->gather_pending = true;
    }
    for (size_t i = 0; i < nloc; i++) {
        GRP_CUDA(g, cudaSetDevice(g->ctx[i]->device));
        GRP_CUDA(g, cudaEventRecord(g->done[i], g->comm_stream[i]));
        if (buffer_id != MRT_BUF_LDR) GRP_CUDA(g, cudaStreamWaitEvent(g->ctx[i]->stream, g->done[i], 0));
        else {
            // the LDR framebuffer is double-buffered and gathers do not block the host: the tonemap after next, which
            // overwrites this half, waits for this gather (the context's own async-readback guard, tonemap.cu)
            mrt_context* c = g->ctx[i];
            if (!c->copy_done[c->ldr_cur]) GRP_CUDA(g, cudaEventCreateWithFlags(&c->copy_done[c->ldr_cur], cudaEventDisableTiming));
            GRP_CUDA(g, cudaEventRecord(c->copy_done[c->ldr_cur], g->comm_stream[i]));
            c->copy_pending[c->ldr_cur] = true;
        }
    }
    return MRT_OK;
}

// Sample-set exchange: in-place sum of the fp32 accumulators (xyz radiance sums, w sample counts) onto the root.
int mrt_group_reduce(mrt_group* g, uint32_t root) {
    if (!g) return MRT_ERR_INVALID;
    if (root >= g->nranks) return group_fail(g, MRT_ERR_INVALID, "reduce: root %u of %u", root, g->nranks);
    if (g->nranks == 1) return MRT_OK;
    if (g->transport != MRT_GROUP_NCCL) return group_fail(g, MRT_ERR_STATE, "mrt_group_reduce needs the NCCL transport");
    const size_t nloc = g->ctx.size();
    std::vector<void*> buf(nloc);
    std::vector<size_t> bytes(nloc);
    for (size_t i = 0; i < nloc; i++) GRP_CTX(g, i, mrt_buffer(g->ctx[i], MRT_BUF_ACCUM, &buf[i], &bytes[i]));
    GRP_NCCL(g, g_nccl.GroupStart());
    for (size_t i = 0; i < nloc; i++)  // on the render stream: the accumulator is the next frame's input
        GRP_NCCL(g, g_nccl.Reduce(buf[i], buf[i], bytes[i] / 4, ncclFloat32, ncclSum, (int)root, g->comm[i], g->ctx[i]->stream));
    GRP_NCCL(g, g_nccl.GroupEnd());
    return MRT_OK;
}

// the gathered image on the root (borrowed device pointer, valid until the next gather); ordered on `stream_out`
int mrt_group_result(mrt_group* g, void** device_ptr, size_t* bytes, void** stream_out) {
    if (!g || !device_ptr || !bytes) return MRT_ERR_INVALID;
    if (g->root_local < 0 || !g->full.p) return group_fail(g, MRT_ERR_STATE, "no gathered image in this process (remote root, or no gather yet)");
    *device_ptr = g->full.p;
    *bytes = g->full_bytes;
    if (stream_out) *stream_out = (void*)g->comm_stream[g->root_local];
    return MRT_OK;
}

int mrt_group_readback(mrt_group* g, void* host, size_t bytes) {
    if (!g || !host) return MRT_ERR_INVALID;
    if (g->root_local < 0 || !g->full.p) return group_fail(g, MRT_ERR_STATE, "no gathered image in this process (remote root, or no gather yet)");
    if (bytes > g->full_bytes) return group_fail(g, MRT_ERR_INVALID, "readback of %zu bytes from a %zu-byte image", bytes, g->full_bytes);
    GRP_CUDA(g, cudaSetDevice(g->ctx[g->root_local]->device));
    GRP_CUDA(g, cudaMemcpyAsync(host, g->full.p, bytes, cudaMemcpyDeviceToHost, g->comm_stream[g->root_local]));
    GRP_CUDA(g, cudaStreamSynchronize(g->comm_stream[g->root_local]));
    g->gather_pending = false;
    return MRT_OK;
}

// Asynchronous readback of the gathered image: the copy is queued on the root's exchange stream behind the gather (and
// ahead of the next one, which reuses the image), so the host can go on issuing frames; mrt_group_readback_wait blocks
// until at most `keep_in_flight` of the queued copies are still outstanding (0: all have landed).
int mrt_group_readback_async(mrt_group* g, void* host, size_t bytes) {
    if (!g || !host) return MRT_ERR_INVALID;
    if (g->root_local < 0 || !g->full.p) return group_fail(g, MRT_ERR_STATE, "no gathered image in this process (remote root, or no gather yet)");
    if (bytes > g->full_bytes) return group_fail(g, MRT_ERR_INVALID, "readback of %zu bytes from a %zu-byte image", bytes, g->full_bytes);
    GRP_CUDA(g, cudaSetDevice(g->ctx[g->root_local]->device));
    cudaStream_t cs = g->comm_stream[g->root_local];
    GRP_CUDA(g, cudaMemcpyAsync(host, g->full.p, bytes, cudaMemcpyDeviceToHost, cs));
    cudaEvent_t e;
    GRP_CUDA(g, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    GRP_CUDA(g, cudaEventRecord(e, cs));
    g->copies.push_back(e);
    return MRT_OK;
}

int mrt_group_readback_wait(mrt_group* g, uint32_t keep_in_flight) {
    if (!g) return MRT_ERR_INVALID;
    while (g->copies.size() > keep_in_flight) {
        cudaEvent_t e = g->copies.front();
        g->copies.erase(g->copies.begin());
        cudaError_t r = cudaEventSynchronize(e);
        cudaEventDestroy(e);
        if (r != cudaSuccess) return group_fail(g, MRT_ERR_CUDA, "readback wait: %s", cudaGetErrorString(r));
    }
    return MRT_OK;
}

int mrt_group_sync(mrt_group* g) {
    if (!g) return MRT_ERR_INVALID;
    for (size_t i = 0; i < g->ctx.size(); i++) {
        if (i < g->fctx.size())
            for (mrt_context* c : g->fctx[i])
                if (mrt_sync(c) != MRT_OK) return group_fail(g, MRT_ERR_CUDA, "rank %u frame context: %s", g->rank[i], mrt_last_error(c));
        GRP_CTX(g, i, mrt_sync(g->ctx[i]));
        GRP_CUDA(g, cudaStreamSynchronize(g->comm_stream[i]));
    }
    g->gather_pending = false;
    return mrt_group_readback_wait(g, 0);
}

}  // extern "C"
