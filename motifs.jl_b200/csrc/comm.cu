// Multi-GPU plumbing behind the C ABI: one process per GPU, one NCCL communicator per ctx.
//
// Reference seams: train.jl:33-52 (the training loop whose gradients are averaged over ranks) and
// render.jl:70-85 (per-motif counts, summed over ranks once per scan).  The reference itself is single-GPU; this file is
// what lets a Julia (or any) host get data parallelism through `ccall` alone, without torch.distributed.
//
// NCCL is bound at run time with dlopen("libnccl.so.2") — preferring a copy the process has already loaded (e.g. the one
// bundled with PyTorch) — so the library has no link-time dependency on it and single-GPU users never load it.  There is no
// fallback transport: without NCCL mb200_comm_init fails.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};
NcclApi g_nccl;

bool nccl_load(std::string* why) {
    if (g_nccl.handle) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);      // already in the process (PyTorch's bundled copy)?
    if (!h) { const char* e = getenv("MB200_NCCL_LIBRARY"); if (e && *e) h = dlopen(e, RTLD_NOW | RTLD_GLOBAL); }
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { if (why) *why = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
#define MB_SYM(field, name) do { *(void**)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) { if (why) *why = std::string("libnccl lacks ") + name; dlclose(h); return false; } } while (0)
    MB_SYM(GetUniqueId, "ncclGetUniqueId");
    MB_SYM(CommInitRank, "ncclCommInitRank");
    MB_SYM(CommDestroy, "ncclCommDestroy");
    MB_SYM(AllReduce, "ncclAllReduce");
    MB_SYM(Broadcast, "ncclBroadcast");
    MB_SYM(AllGather, "ncclAllGather");
    MB_SYM(GetErrorString, "ncclGetErrorString");
    MB_SYM(GetVersion, "ncclGetVersion");
#undef MB_SYM
    g_nccl.handle = h;
    return true;
}

}  // namespace

#define MB_NCCL(ctx, expr) do { ncclResult_t _r = (expr); if (_r != ncclSuccess) { \
    MB_FAIL(ctx, MB200_E_COMM, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); } } while (0)

static_assert(sizeof(ncclUniqueId) == MB200_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");

extern "C" int32_t mb200_comm_unique_id(uint8_t* id_out) {
    if (!id_out) return MB200_E_INVALID;
    if (!nccl_load(nullptr)) return MB200_E_COMM;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return MB200_E_COMM;
    memcpy(id_out, &id, sizeof id);
    return MB200_OK;
}

extern "C" int32_t mb200_comm_init(mb200_ctx* ctx, const uint8_t* id, int32_t rank, int32_t world) {
    if (!ctx || !id) return MB200_E_INVALID;
    if (world < 1 || rank < 0 || rank >= world) MB_FAIL(ctx, MB200_E_INVALID, "comm_init: rank %d of %d", rank, world);
    if (ctx->comm) MB_FAIL(ctx, MB200_E_INVALID, "comm_init: this ctx already has a communicator");
    std::string why;
    if (!nccl_load(&why)) MB_FAIL(ctx, MB200_E_COMM, "comm_init: %s", why.c_str());
    MB_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    ncclComm_t c = nullptr;
    MB_NCCL(ctx, g_nccl.CommInitRank(&c, world, uid, rank));
    ctx->comm = c; ctx->rank = rank; ctx->world = world;
    return MB200_OK;
}

extern "C" int32_t mb200_comm_destroy(mb200_ctx* ctx) {
    if (!ctx) return MB200_E_INVALID;
    if (ctx->comm) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        g_nccl.CommDestroy((ncclComm_t)ctx->comm);
        ctx->comm = nullptr;
    }
    ctx->rank = 0; ctx->world = 1;
    return MB200_OK;
}

extern "C" int32_t mb200_comm_info(const mb200_ctx* ctx, int32_t* rank, int32_t* world, int32_t* nccl_version) {
    if (!ctx) return MB200_E_INVALID;
    if (rank) *rank = ctx->rank;
    if (world) *world = ctx->world;
    if (nccl_version) { int v = 0; if (g_nccl.handle) g_nccl.GetVersion(&v); *nccl_version = v; }
    return MB200_OK;
}

// ---- collectives on device buffers, enqueued on the ctx stream (no host synchronisation) -----------------------------------
int mb_comm_allreduce_f32(mb200_ctx* ctx, float* buf, size_t n, bool average) {
    if (!ctx->comm || ctx->world == 1) return MB200_OK;
    MB_NCCL(ctx, g_nccl.AllReduce(buf, buf, n, ncclFloat32, average ? ncclAvg : ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    return MB200_OK;
}
int mb_comm_allreduce_u64(mb200_ctx* ctx, unsigned long long* buf, size_t n) {
    if (!ctx->comm || ctx->world == 1) return MB200_OK;
    MB_NCCL(ctx, g_nccl.AllReduce(buf, buf, n, ncclUint64, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    return MB200_OK;
}
int mb_comm_allreduce_u32(mb200_ctx* ctx, unsigned int* buf, size_t n) {
    if (!ctx->comm || ctx->world == 1) return MB200_OK;
    MB_NCCL(ctx, g_nccl.AllReduce(buf, buf, n, ncclUint32, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    return MB200_OK;
}
int mb_comm_broadcast_bytes(mb200_ctx* ctx, void* buf, size_t bytes, int root) {
    if (!ctx->comm || ctx->world == 1) return MB200_OK;
    MB_NCCL(ctx, g_nccl.Broadcast(buf, buf, bytes, ncclUint8, root, (ncclComm_t)ctx->comm, ctx->stream));
    return MB200_OK;
}
int mb_comm_allgather_bytes(mb200_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank) {
    if (!ctx->comm || ctx->world == 1) { if (send != recv) cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream); return MB200_OK; }
    MB_NCCL(ctx, g_nccl.AllGather(send, recv, bytes_per_rank, ncclUint8, (ncclComm_t)ctx->comm, ctx->stream));
    return MB200_OK;
}

// Host-buffer helpers for the host framework's own small exchanges (an epoch seed, a stop flag, code counts): staged through the
// ctx's scratch buffer.  Blocking.
extern "C" int32_t mb200_comm_broadcast(mb200_ctx* ctx, void* host_buf, int64_t bytes, int32_t root) {
    if (!ctx || !host_buf || bytes < 0) return MB200_E_INVALID;
    if (!ctx->comm || ctx->world == 1 || bytes == 0) return MB200_OK;
    if (root < 0 || root >= ctx->world) MB_FAIL(ctx, MB200_E_INVALID, "comm_broadcast: root %d of %d", root, ctx->world);
    MB_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = mb_ensure_scratch(ctx, (size_t)bytes); if (rc) return rc;
    if (ctx->rank == root) MB_CUDA(ctx, cudaMemcpyAsync(ctx->scratch, host_buf, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
    rc = mb_comm_broadcast_bytes(ctx, ctx->scratch, (size_t)bytes, root); if (rc) return rc;
    if (ctx->rank != root) MB_CUDA(ctx, cudaMemcpyAsync(host_buf, ctx->scratch, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MB200_OK;
}

extern "C" int32_t mb200_comm_allreduce_i64(mb200_ctx* ctx, int64_t* host_buf, int64_t n) {
    if (!ctx || !host_buf || n < 0) return MB200_E_INVALID;
    if (!ctx->comm || ctx->world == 1 || n == 0) return MB200_OK;
    MB_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = mb_ensure_scratch(ctx, (size_t)n * 8); if (rc) return rc;
    MB_CUDA(ctx, cudaMemcpyAsync(ctx->scratch, host_buf, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    MB_NCCL(ctx, g_nccl.AllReduce(ctx->scratch, ctx->scratch, (size_t)n, ncclInt64, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    MB_CUDA(ctx, cudaMemcpyAsync(host_buf, ctx->scratch, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MB200_OK;
}

extern "C" int32_t mb200_comm_allgather(mb200_ctx* ctx, const void* host_send, void* host_recv, int64_t bytes_per_rank) {
    if (!ctx || !host_send || !host_recv || bytes_per_rank < 0) return MB200_E_INVALID;
    if (bytes_per_rank == 0) return MB200_OK;
    if (!ctx->comm || ctx->world == 1) { if (host_send != host_recv) memcpy(host_recv, host_send, (size_t)bytes_per_rank); return MB200_OK; }
    MB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t b = (size_t)bytes_per_rank, tot = b * (size_t)(ctx->world + 1);
    int rc = mb_ensure_scratch(ctx, tot); if (rc) return rc;
    uint8_t* s = (uint8_t*)ctx->scratch;
    MB_CUDA(ctx, cudaMemcpyAsync(s, host_send, b, cudaMemcpyHostToDevice, ctx->stream));
    rc = mb_comm_allgather_bytes(ctx, s, s + b, b); if (rc) return rc;
    MB_CUDA(ctx, cudaMemcpyAsync(host_recv, s + b, b * (size_t)ctx->world, cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MB200_OK;
}
