// dist.cu -- data-parallel plumbing: one process per GPU, an NCCL communicator owned by the library so that the
// all-reduces of the step (BN batch statistics, BN backward sums, gradients, loss accumulators) are enqueued on the
// state's stream between the kernels -- and captured into the step's CUDA graph with them.
// The reference has no multi-GPU code; equivalence target = the single-process step at the global batch.
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy PyTorch already loaded, else the system one), so the
// library has no link-time dependency and single-GPU users never touch it.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "common.cuh"

namespace {
struct NcclUniqueId { char internal[128]; };
typedef void *NcclComm;
typedef int (*fn_get_id)(NcclUniqueId *);
typedef int (*fn_init_rank)(NcclComm *, int, NcclUniqueId, int);
typedef int (*fn_all_reduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_broadcast)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_all_gather)(const void *, void *, size_t, int, NcclComm, cudaStream_t);
typedef int (*fn_destroy)(NcclComm);
typedef const char *(*fn_errstr)(int);
typedef int (*fn_group)(void);
struct NcclApi {
    void *handle = nullptr;
    fn_get_id get_id = nullptr; fn_init_rank init_rank = nullptr; fn_all_reduce all_reduce = nullptr; fn_broadcast broadcast = nullptr; fn_all_gather all_gather = nullptr;
    fn_destroy destroy = nullptr; fn_errstr errstr = nullptr; fn_group group_start = nullptr, group_end = nullptr;
} g_nccl;
enum { NCCL_SUM = 0, NCCL_UINT8 = 1, NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8, NCCL_BFLOAT16 = 9 };

int load_nccl() {
    if (g_nccl.handle) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    void *h = nullptr;
    for (const char *n : names) { h = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (h) break; }
    if (!h) { cenn_set_error("NCCL not found (dlopen libnccl.so.2: %s)", dlerror()); return 1; }
    g_nccl.get_id = (fn_get_id)dlsym(h, "ncclGetUniqueId");
    g_nccl.init_rank = (fn_init_rank)dlsym(h, "ncclCommInitRank");
    g_nccl.all_reduce = (fn_all_reduce)dlsym(h, "ncclAllReduce");
    g_nccl.broadcast = (fn_broadcast)dlsym(h, "ncclBroadcast");
    g_nccl.all_gather = (fn_all_gather)dlsym(h, "ncclAllGather");
    g_nccl.destroy = (fn_destroy)dlsym(h, "ncclCommDestroy");
    g_nccl.errstr = (fn_errstr)dlsym(h, "ncclGetErrorString");
    g_nccl.group_start = (fn_group)dlsym(h, "ncclGroupStart");
    g_nccl.group_end = (fn_group)dlsym(h, "ncclGroupEnd");
    if (!g_nccl.get_id || !g_nccl.init_rank || !g_nccl.all_reduce || !g_nccl.broadcast || !g_nccl.destroy || !g_nccl.errstr) {
        cenn_set_error("libnccl.so.2 lacks an expected symbol"); dlclose(h); return 1;
    }
    g_nccl.handle = h;
    return 0;
}
int nccl_check(int rc, const char *what) {
    if (rc == 0) return 0;
    cenn_set_error("NCCL error in %s: %s", what, g_nccl.errstr ? g_nccl.errstr(rc) : "?");
    return 1;
}
}  // namespace

int cenn_dist_all_reduce_on(cenn_state *s, void *buf, int64_t count, int is_double, cudaStream_t stream) {
    if (!s->comm) { cenn_set_error("cenn_dist_all_reduce: no communicator (call cenn_dist_init)"); return 1; }
    if (count <= 0) return 0;
    return nccl_check(g_nccl.all_reduce(buf, buf, (size_t)count, is_double ? NCCL_FLOAT64 : NCCL_FLOAT32, NCCL_SUM, s->comm, stream), "ncclAllReduce");
}

// several small all-reduces as ONE NCCL launch (the leftover ranges of a gradient vector)
int cenn_dist_group(int begin) {
    if (!g_nccl.handle || !g_nccl.group_start || !g_nccl.group_end) return 0;
    return nccl_check(begin ? g_nccl.group_start() : g_nccl.group_end(), begin ? "ncclGroupStart" : "ncclGroupEnd");
}
int cenn_dist_all_reduce_bulk(cenn_state *s, float *buf, int64_t count) {
    if (!s->comm2 || !s->comm_stream) { cenn_set_error("cenn_dist_all_reduce_bulk: no bulk communicator"); return 1; }
    if (count <= 0) return 0;
    return nccl_check(g_nccl.all_reduce(buf, buf, (size_t)count, NCCL_FLOAT32, NCCL_SUM, s->comm2, s->comm_stream), "ncclAllReduce (bulk)");
}

// bf16 gradient bucket (in-place sum) on the bulk communicator
int cenn_dist_all_reduce_bulk_bf16(cenn_state *s, void *buf, int64_t count) {
    if (!s->comm2 || !s->comm_stream) { cenn_set_error("cenn_dist_all_reduce_bulk_bf16: no bulk communicator"); return 1; }
    if (count <= 0) return 0;
    return nccl_check(g_nccl.all_reduce(buf, buf, (size_t)count, NCCL_BFLOAT16, NCCL_SUM, s->comm2, s->comm_stream), "ncclAllReduce (bulk, bf16)");
}

// Map one device allocation of every rank into this process (CUDA IPC over NVLink): peers[r] = rank r's `base` (peers[rank] = base).
// Collective over the first communicator (every rank calls it in the same order); returns 1 and leaves nothing mapped if any rank failed.
int cenn_dist_ipc_map(cenn_state *s, void *base, void **peers) {
    if (!s->comm || !g_nccl.all_gather || s->world > XR_MAX_WORLD) { cenn_set_error("cenn_dist_ipc_map: no communicator"); return 1; }
    const int world = s->world, rank = s->rank;
    cudaIpcMemHandle_t mine;
    bool ok = cudaIpcGetMemHandle(&mine, base) == cudaSuccess;
    if (!ok) { cudaGetLastError(); memset(&mine, 0, sizeof(mine)); }
    std::vector<cudaIpcMemHandle_t> all(world);
    void *hbuf = nullptr;
    if (cudaMalloc(&hbuf, sizeof(mine) * (world + 1)) != cudaSuccess) { cenn_set_error("cenn_dist_ipc_map: allocation failed"); return 1; }
    char *send = (char *)hbuf + sizeof(mine) * world;
    cudaMemcpy(send, &mine, sizeof(mine), cudaMemcpyHostToDevice);
    if (g_nccl.all_gather(send, hbuf, sizeof(mine), NCCL_UINT8, s->comm, s->stream) != 0) ok = false;
    cudaStreamSynchronize(s->stream);
    cudaMemcpy(all.data(), hbuf, sizeof(mine) * world, cudaMemcpyDeviceToHost);
    cudaFree(hbuf);
    for (int r = 0; r < world; ++r) peers[r] = nullptr;
    for (int r = 0; r < world && ok; ++r) {
        if (r == rank) { peers[r] = base; continue; }
        if (cudaIpcOpenMemHandle(&peers[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); peers[r] = nullptr; ok = false; }
    }
    float *flag_dev = nullptr; float flag_host = ok ? 0.f : 1.f;
    cudaMalloc(&flag_dev, sizeof(float)); cudaMemcpy(flag_dev, &flag_host, sizeof(float), cudaMemcpyHostToDevice);
    g_nccl.all_reduce(flag_dev, flag_dev, 1, NCCL_FLOAT32, NCCL_SUM, s->comm, s->stream);
    cudaStreamSynchronize(s->stream);
    cudaMemcpy(&flag_host, flag_dev, sizeof(float), cudaMemcpyDeviceToHost); cudaFree(flag_dev);
    if (flag_host != 0.f) { cenn_dist_ipc_unmap(s, peers); cenn_set_error("cenn_dist_ipc_map: CUDA IPC mapping failed on at least one rank"); return 1; }
    return 0;
}
void cenn_dist_ipc_unmap(cenn_state *s, void **peers) {
    for (int r = 0; r < s->world && r < XR_MAX_WORLD; ++r) { if (r != s->rank && peers[r]) cudaIpcCloseMemHandle(peers[r]); peers[r] = nullptr; }
}
// host-level barrier over the ranks (a one-element all-reduce + stream synchronisation); 0 if there is no communicator
int cenn_dist_barrier(cenn_state *s) {
    if (!s->comm) return 0;
    float *d = nullptr;
    if (cudaMalloc(&d, sizeof(float)) != cudaSuccess) return 1;
    cudaMemset(d, 0, sizeof(float));
    int rc = g_nccl.all_reduce(d, d, 1, NCCL_FLOAT32, NCCL_SUM, s->comm, s->stream);
    cudaStreamSynchronize(s->stream);
    cudaFree(d);
    return rc != 0;
}


extern "C" {

int cenn_dist_unique_id(void *id128_host) {
    if (!id128_host) { cenn_set_error("null id buffer"); return 1; }
    if (load_nccl()) return 1;
    NcclUniqueId id;
    if (nccl_check(g_nccl.get_id(&id), "ncclGetUniqueId")) return 1;
    memcpy(id128_host, &id, sizeof(id));
    return 0;
}

int cenn_dist_init(cenn_state *s, const void *id128_host, int world_size, int rank) {
    API_BEGIN(s);
    REQUIRE(id128_host && world_size >= 1 && rank >= 0 && rank < world_size, "cenn_dist_init: bad arguments (world %d, rank %d)", world_size, rank);
    REQUIRE(!s->comm, "cenn_dist_init: communicator already initialised");
    if (load_nccl()) return 1;
    NcclUniqueId id;
    memcpy(&id, id128_host, sizeof(id));
    NcclComm comm = nullptr;
    if (nccl_check(g_nccl.init_rank(&comm, world_size, id, rank), "ncclCommInitRank")) return 1;
    s->comm = comm; s->world = world_size; s->rank = rank;
    // second communicator for the gradient buckets (its own channels: a 285 MB all-reduce must not queue in front of the
    // BN-statistics reductions): rank 0 draws another id and broadcasts it over the first communicator
    NcclUniqueId id2;
    void *dbuf = nullptr;
    CK(cudaMalloc(&dbuf, sizeof(id2)));
    if (rank == 0) { if (nccl_check(g_nccl.get_id(&id2), "ncclGetUniqueId")) return 1; CK(cudaMemcpy(dbuf, &id2, sizeof(id2), cudaMemcpyHostToDevice)); }
    if (nccl_check(g_nccl.broadcast(dbuf, dbuf, sizeof(id2), NCCL_UINT8, 0, s->comm, s->stream), "ncclBroadcast")) return 1;
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaMemcpy(&id2, dbuf, sizeof(id2), cudaMemcpyDeviceToHost));
    cudaFree(dbuf);
    NcclComm comm2 = nullptr;
    if (nccl_check(g_nccl.init_rank(&comm2, world_size, id2, rank), "ncclCommInitRank (bulk)")) return 1;
    s->comm2 = comm2;
    {   // CENN_COMM_PRIO=1: highest stream priority for the bucket all-reduces (measured at N = 2: no effect on the step, so the default stays 0)
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        const char *e = getenv("CENN_COMM_PRIO");
        CK(cudaStreamCreateWithPriority(&s->comm_stream, cudaStreamNonBlocking, (e && atoi(e) != 0) ? prio_hi : 0));
    }
    // peer mailboxes for the latency-bound BN-statistics exchanges (35 per step): CUDA IPC over NVLink.  Any failure
    // here just leaves xr_enabled = false and those exchanges stay on NCCL.
    do {
        if (!g_nccl.all_gather || world_size > XR_MAX_WORLD || getenv("CENN_NO_XR")) break;
        const size_t half = (size_t)2 * world_size * XR_MAXF * 8 + 256;   // one mailbox: 2 parity slots x one row per source rank of {value, tag} words (+ spare)
        const size_t bytes = 3 * half;                                  // three independent mailbox sequences (main chain, second chain, communication stream)
        void *own = nullptr, *hbuf = nullptr;
        if (cudaMalloc(&own, bytes) != cudaSuccess) { cudaGetLastError(); break; }
        cudaMemset(own, 0, bytes);
        cudaIpcMemHandle_t mine;
        if (cudaIpcGetMemHandle(&mine, own) != cudaSuccess) { cudaGetLastError(); cudaFree(own); break; }
        std::vector<cudaIpcMemHandle_t> all(world_size);
        if (cudaMalloc(&hbuf, sizeof(mine) * (world_size + 1)) != cudaSuccess) { cudaGetLastError(); cudaFree(own); break; }
        char *send = (char *)hbuf + sizeof(mine) * world_size;
        cudaMemcpy(send, &mine, sizeof(mine), cudaMemcpyHostToDevice);
        cudaDeviceSynchronize();
        int rc = g_nccl.all_gather(send, hbuf, sizeof(mine), NCCL_UINT8, s->comm, s->stream);
        cudaStreamSynchronize(s->stream);
        cudaMemcpy(all.data(), hbuf, sizeof(mine) * world_size, cudaMemcpyDeviceToHost);
        cudaFree(hbuf);
        bool ok = rc == 0;
        XrCtx x = {}, x2 = {}, x3 = {};
        x.world = x2.world = x3.world = world_size; x.rank = x2.rank = x3.rank = rank;
        x.push = x2.push = x3.push = getenv("CENN_XR_PULL") ? 0 : 1;
        { const char *e = getenv("CENN_XR_TIMEOUT_S"); const double sec = e ? atof(e) : 120.0; x.timeout_cycles = x2.timeout_cycles = x3.timeout_cycles = (long long)(sec * 2.0e9); }
        for (int r = 0; r < world_size && ok; ++r) {
            void *base = own;
            if (r != rank && cudaIpcOpenMemHandle(&base, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
            x.data[r] = reinterpret_cast<float *>(base);
            x.flags[r] = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(base) + (size_t)2 * world_size * XR_MAXF * 8);
            x2.data[r] = reinterpret_cast<float *>(reinterpret_cast<char *>(base) + half);
            x2.flags[r] = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(base) + half + (size_t)2 * world_size * XR_MAXF * 8);
            x3.data[r] = reinterpret_cast<float *>(reinterpret_cast<char *>(base) + 2 * half);
            x3.flags[r] = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(base) + 2 * half + (size_t)2 * world_size * XR_MAXF * 8);
        }
        // every rank must learn whether ALL ranks succeeded (a partial set-up would deadlock the exchange)
        float *flag_dev = nullptr; float flag_host = ok ? 0.f : 1.f;
        cudaMalloc(&flag_dev, sizeof(float)); cudaMemcpy(flag_dev, &flag_host, sizeof(float), cudaMemcpyHostToDevice);
        g_nccl.all_reduce(flag_dev, flag_dev, 1, NCCL_FLOAT32, NCCL_SUM, s->comm, s->stream);
        cudaStreamSynchronize(s->stream);
        cudaMemcpy(&flag_host, flag_dev, sizeof(float), cudaMemcpyDeviceToHost); cudaFree(flag_dev);
        if (flag_host != 0.f) { for (int r = 0; r < world_size; ++r) if (r != rank && x.data[r]) cudaIpcCloseMemHandle(x.data[r]); cudaFree(own); break; }
        unsigned long long *ep = nullptr;
        cudaMalloc(&ep, 3 * sizeof(*ep)); cudaMemset(ep, 0, 3 * sizeof(*ep));
        x.epoch = ep; x2.epoch = ep + 1; x3.epoch = ep + 2;
        s->xr = x; s->xr2 = x2; s->xr3 = x3; s->xr_own = own; s->xr_enabled = true;
    } while (0);
    return 0;
}

int cenn_dist_all_reduce(cenn_state *s, void *buf_dev, int64_t count, int is_double) {
    API_BEGIN(s);
    return cenn_dist_all_reduce_on(s, buf_dev, count, is_double, s->stream);
}

int cenn_dist_broadcast(cenn_state *s, void *buf_dev, int64_t bytes, int root) {
    API_BEGIN(s);
    REQUIRE(s->comm, "cenn_dist_broadcast: no communicator (call cenn_dist_init)");
    return nccl_check(g_nccl.broadcast(buf_dev, buf_dev, (size_t)bytes, NCCL_UINT8, root, s->comm, s->stream), "ncclBroadcast");
}

int cenn_dist_shutdown(cenn_state *s) {
    API_BEGIN(s);
    if (s->xr_enabled) {
        cudaDeviceSynchronize();
        for (int r = 0; r < s->xr.world; ++r) if (r != s->xr.rank && s->xr.data[r]) cudaIpcCloseMemHandle(s->xr.data[r]);
        cudaFree(s->xr_own); cudaFree(s->xr.epoch); s->xr_enabled = false; s->xr_own = nullptr;
    }
    if (s->comm_stream) { cudaStreamSynchronize(s->comm_stream); }
    if (s->comm2) { g_nccl.destroy(s->comm2); s->comm2 = nullptr; }
    if (s->comm_stream) { cudaStreamDestroy(s->comm_stream); s->comm_stream = nullptr; }
    if (s->comm) { cudaStreamSynchronize(s->stream); g_nccl.destroy(s->comm); s->comm = nullptr; s->world = 1; s->rank = 0; }
    return 0;
}

}  // extern "C"
