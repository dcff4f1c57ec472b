// sharded.cu — the row-sharded SMC weight update over the GPUs of one box, driven from C (SURVEY.md §8 row e).
//
// Reference: ABC::weight_predictive_prior, src/AbcUtil.cpp:547-586 (caller: AbcSmc.cpp:1053-1064). Rows i of the new set are
// independent: GPU g owns rows [g per, (g+1) per), per = ceil(N_new / G). What is exchanged:
//   ncclBroadcast   the previous set (theta_old, w_old, dv_old) from the member that holds it          (NVLink / NVSwitch)
//   ncclAllReduce   MAX of one double: the conditioning maximum max |a_i|^2 over ALL rows, so that every member picks the
//                   same formulation (DMMA inner product or pairwise differences) whatever the number of GPUs
//   ncclAllReduce   SUM of one double: the squared norm for Eigen's normalize() (AbcUtil.cpp:583)
//   ncclAllGather   the slices (device-buffer flavour only; the host flavour copies each slice straight to its place)
// Two kinds of group: one host process driving all GPUs (the reference's host is ONE C++ process: ncclCommInitAll, the
// collectives of the members are issued inside ncclGroupStart/End), or one process per GPU (torchrun / MPI launchers:
// ncclCommInitRank with an id the caller passes round). NCCL is resolved with dlopen at first use, so the library has no
// link-time dependency on it (and shares the copy a host such as PyTorch has already loaded).
#include <dlfcn.h>
#include <nccl.h>

#include <new>
#include <vector>

#include "kernels.cuh"

namespace {

struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    char why[256] = {0};
};

NcclApi load_nccl() {
    NcclApi a;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { a.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (a.h) break; }
    if (!a.h) { snprintf(a.why, sizeof(a.why), "libnccl.so.2 could not be loaded: %s", dlerror()); return a; }
#define NCCL_SYM(field, sym)                                                                                       \
    do {                                                                                                           \
        *(void**)(&a.field) = dlsym(a.h, sym);                                                                     \
        if (!a.field) { snprintf(a.why, sizeof(a.why), "%s missing from libnccl", sym); a.h = nullptr; return a; } \
    } while (0)
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId"); NCCL_SYM(CommInitRank, "ncclCommInitRank"); NCCL_SYM(CommInitAll, "ncclCommInitAll");
    NCCL_SYM(CommDestroy, "ncclCommDestroy"); NCCL_SYM(GroupStart, "ncclGroupStart"); NCCL_SYM(GroupEnd, "ncclGroupEnd");
    NCCL_SYM(Broadcast, "ncclBroadcast"); NCCL_SYM(AllReduce, "ncclAllReduce"); NCCL_SYM(AllGather, "ncclAllGather");
    NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef NCCL_SYM
    return a;
}
NcclApi& nccl() {
    static NcclApi a = load_nccl();      // resolved once, thread-safe (C++11 static local)
    return a;
}

}  // namespace

struct abcb200_group {
    int world = 0;                       // members over all processes
    int nlocal = 0;                      // members this process drives
    bool owns_ctx = false;
    std::vector<abcb200_ctx*> ctx;       // [nlocal]
    std::vector<ncclComm_t> comm;        // [nlocal]
    std::vector<int> rank;               // [nlocal] global rank of each local member
    char err[512] = {0};
};

#define GRP_FAIL(g, code, ...)                              \
    do {                                                    \
        snprintf((g)->err, sizeof((g)->err), __VA_ARGS__);  \
        return (code);                                      \
    } while (0)
#define NCCL_TRY(g, call)                                                                                                  \
    do {                                                                                                                   \
        ncclResult_t _r = (call);                                                                                          \
        if (_r != ncclSuccess) GRP_FAIL(g, ABCB200_ECUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, nccl().GetErrorString(_r)); \
    } while (0)
// inside ncclGroupStart / ncclGroupEnd: the open group is closed before the error is returned
#define NCCL_TRY_G(g, call)                                                                                                \
    do {                                                                                                                   \
        ncclResult_t _r = (call);                                                                                          \
        if (_r != ncclSuccess) {                                                                                           \
            nccl().GroupEnd();                                                                                             \
            GRP_FAIL(g, ABCB200_ECUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, nccl().GetErrorString(_r));           \
        }                                                                                                                  \
    } while (0)
#define GRP_CTX_TRY(g, c, call)                                                                  \
    do {                                                                                         \
        int _r = (call);                                                                         \
        if (_r != ABCB200_OK) { snprintf((g)->err, sizeof((g)->err), "%s", (c)->err); return _r; } \
    } while (0)

extern "C" const char* abcb200_group_last_error(abcb200_group* g) { return g ? g->err : "null group"; }
extern "C" int abcb200_group_size(const abcb200_group* g) { return g ? g->world : 0; }
extern "C" int abcb200_group_local_size(const abcb200_group* g) { return g ? g->nlocal : 0; }
extern "C" abcb200_ctx* abcb200_group_ctx(abcb200_group* g, int local_index) {
    return (g && local_index >= 0 && local_index < g->nlocal) ? g->ctx[(size_t)local_index] : nullptr;
}

extern "C" int abcb200_group_create(int n_gpus, const int* device_ids, abcb200_group** out) {
    if (!out) return ABCB200_EINVAL;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return ABCB200_ENODEV;
    if (n_gpus <= 0) n_gpus = ndev;
    if (n_gpus > ndev) return ABCB200_ENODEV;
    if (!nccl().h) return ABCB200_ENODEV;
    abcb200_group* g = new (std::nothrow) abcb200_group();
    if (!g) return ABCB200_ENOMEM;
    g->world = g->nlocal = n_gpus;
    g->owns_ctx = true;
    std::vector<int> devs((size_t)n_gpus);
    for (int i = 0; i < n_gpus; i++) devs[(size_t)i] = device_ids ? device_ids[i] : i;
    g->ctx.assign((size_t)n_gpus, nullptr);
    g->comm.assign((size_t)n_gpus, nullptr);
    g->rank.resize((size_t)n_gpus);
    int rc = ABCB200_OK;
    for (int i = 0; i < n_gpus && rc == ABCB200_OK; i++) { g->rank[(size_t)i] = i; rc = abcb200_create(devs[(size_t)i], &g->ctx[(size_t)i]); }
    if (rc == ABCB200_OK && nccl().CommInitAll(g->comm.data(), n_gpus, devs.data()) != ncclSuccess) rc = ABCB200_ECUDA;
    if (rc != ABCB200_OK) {
        for (auto c : g->ctx) if (c) abcb200_destroy(c);
        delete g;
        return rc;
    }
    *out = g;
    return ABCB200_OK;
}

extern "C" int abcb200_group_unique_id(void* id_out, size_t bytes) {
    if (!id_out || bytes < sizeof(ncclUniqueId) || !nccl().h) return ABCB200_EINVAL;
    ncclUniqueId id;
    if (nccl().GetUniqueId(&id) != ncclSuccess) return ABCB200_ECUDA;
    memcpy(id_out, &id, sizeof(id));
    return ABCB200_OK;
}

extern "C" int abcb200_group_create_rank(abcb200_ctx* ctx, const void* id, int rank, int world, abcb200_group** out) {
    if (!out) return ABCB200_EINVAL;
    *out = nullptr;
    if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return ABCB200_EINVAL;
    if (!nccl().h) { snprintf(ctx->err, sizeof(ctx->err), "%s", nccl().why); return ABCB200_ENODEV; }
    abcb200_group* g = new (std::nothrow) abcb200_group();
    if (!g) return ABCB200_ENOMEM;
    g->world = world; g->nlocal = 1; g->owns_ctx = false;
    g->ctx.assign(1, ctx); g->comm.assign(1, nullptr); g->rank.assign(1, rank);
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    if (cudaSetDevice(ctx->device) != cudaSuccess || nccl().CommInitRank(&g->comm[0], world, uid, rank) != ncclSuccess) {
        snprintf(ctx->err, sizeof(ctx->err), "ncclCommInitRank(rank %d of %d) failed", rank, world);
        delete g;
        return ABCB200_ECUDA;
    }
    *out = g;
    return ABCB200_OK;
}

extern "C" int abcb200_group_destroy(abcb200_group* g) {
    if (!g) return ABCB200_OK;
    for (int i = 0; i < g->nlocal; i++) {
        cudaSetDevice(g->ctx[(size_t)i]->device);
        cudaStreamSynchronize(g->ctx[(size_t)i]->stream);
        if (g->comm[(size_t)i]) nccl().CommDestroy(g->comm[(size_t)i]);
        if (g->owns_ctx) abcb200_destroy(g->ctx[(size_t)i]);
    }
    delete g;
    return ABCB200_OK;
}

namespace {

struct Member {                 // what one local member works on (device pointers on its own device)
    const double *numer, *th_new;        // full-length buffers (N_new rows, ld_new), or slice-only with row0 = lo (host flavour)
    int64_t ld_new, row0;                // th_new[(p * ld_new) + (i - row0)] is row i
    double *th_old, *w_old, *dv_old;     // the previous set (filled by the broadcast when bcast_root >= 0)
    int64_t ld_old;
    double* w_slice;                     // per entries: the member's un-normalised, then normalised rows
    double* ss;                          // one double
    WeightsJob job;
    int64_t lo, hi;
};

// The exchange + compute sequence shared by both flavours. per = ceil(N_new / world).
int sharded_core(abcb200_group* g, std::vector<Member>& mem, int64_t N_new, int64_t N_old, int P, int algo, int bcast_root, double* const* gather_out) {
    NcclApi& nc = nccl();
    const int64_t per = (N_new + g->world - 1) / g->world;
    if (bcast_root >= 0 && g->world > 1) {   // the previous set travels GPU to GPU, not once per GPU over PCIe
        NCCL_TRY(g, nc.GroupStart());
        for (int m = 0; m < g->nlocal; m++) {
            Member& me = mem[(size_t)m];
            abcb200_ctx* c = g->ctx[(size_t)m];
            cudaSetDevice(c->device);
            NCCL_TRY_G(g, nc.Broadcast(me.th_old, me.th_old, (size_t)me.ld_old * P, ncclDouble, bcast_root, g->comm[(size_t)m], c->stream));
            NCCL_TRY_G(g, nc.Broadcast(me.w_old, me.w_old, (size_t)N_old, ncclDouble, bcast_root, g->comm[(size_t)m], c->stream));
            NCCL_TRY_G(g, nc.Broadcast(me.dv_old, me.dv_old, (size_t)P, ncclDouble, bcast_root, g->comm[(size_t)m], c->stream));
        }
        NCCL_TRY(g, nc.GroupEnd());
    }
    // phase 1 on every member: constants, packed operands, the slice's conditioning maximum
    for (int m = 0; m < g->nlocal; m++) {
        Member& me = mem[(size_t)m];
        abcb200_ctx* c = g->ctx[(size_t)m];
        cudaSetDevice(c->device);
        me.lo = std::min<int64_t>((int64_t)g->rank[(size_t)m] * per, N_new);
        me.hi = std::min<int64_t>(me.lo + per, N_new);
        if (me.hi > me.lo)
            GRP_CTX_TRY(g, c, weights_pack(c, me.th_new + (me.lo - me.row0), me.ld_new, me.hi - me.lo, me.th_old, me.ld_old, N_old, me.w_old, me.dv_old, P, algo, &me.job));
        else {       // an empty slice still takes part in the collectives
            me.job.scal = ws_new<double>(c, 4);
            if (!me.job.scal) GRP_FAIL(g, ABCB200_ENOMEM, "workspace exhausted in weights_sharded");
            if (cudaMemsetAsync(me.job.scal, 0, 4 * sizeof(double), c->stream) != cudaSuccess) GRP_FAIL(g, ABCB200_ECUDA, "memset failed");
        }
    }
    if (g->world > 1) {      // one gate for everybody: max |a|^2 over all rows (bit patterns of non-negative doubles order like the values)
        NCCL_TRY(g, nc.GroupStart());
        for (int m = 0; m < g->nlocal; m++) {
            abcb200_ctx* c = g->ctx[(size_t)m];
            cudaSetDevice(c->device);
            NCCL_TRY_G(g, nc.AllReduce(mem[(size_t)m].job.scal + 1, mem[(size_t)m].job.scal + 1, 1, ncclDouble, ncclMax, g->comm[(size_t)m], c->stream));
        }
        NCCL_TRY(g, nc.GroupEnd());
    }
    for (int m = 0; m < g->nlocal; m++) {
        Member& me = mem[(size_t)m];
        abcb200_ctx* c = g->ctx[(size_t)m];
        cudaSetDevice(c->device);
        if (me.hi > me.lo) GRP_CTX_TRY(g, c, weights_eval(c, &me.job, me.numer ? me.numer + (me.lo - me.row0) : nullptr, me.w_slice, me.ss));
        else if (cudaMemsetAsync(me.ss, 0, sizeof(double), c->stream) != cudaSuccess) GRP_FAIL(g, ABCB200_ECUDA, "memset failed");
    }
    if (g->world > 1) {      // NaN rows poison the sum as they poison squaredNorm()
        NCCL_TRY(g, nc.GroupStart());
        for (int m = 0; m < g->nlocal; m++) {
            abcb200_ctx* c = g->ctx[(size_t)m];
            cudaSetDevice(c->device);
            NCCL_TRY_G(g, nc.AllReduce(mem[(size_t)m].ss, mem[(size_t)m].ss, 1, ncclDouble, ncclSum, g->comm[(size_t)m], c->stream));
        }
        NCCL_TRY(g, nc.GroupEnd());
    }
    for (int m = 0; m < g->nlocal; m++) {
        Member& me = mem[(size_t)m];
        abcb200_ctx* c = g->ctx[(size_t)m];
        cudaSetDevice(c->device);
        if (me.hi > me.lo) GRP_CTX_TRY(g, c, launch_scale_weights(c, me.w_slice, me.hi - me.lo, me.ss));
    }
    if (gather_out) {
        if (g->world > 1) {
            NCCL_TRY(g, nc.GroupStart());
            for (int m = 0; m < g->nlocal; m++) {
                abcb200_ctx* c = g->ctx[(size_t)m];
                cudaSetDevice(c->device);
                NCCL_TRY_G(g, nc.AllGather(mem[(size_t)m].w_slice, gather_out[m], (size_t)per, ncclDouble, g->comm[(size_t)m], c->stream));
            }
            NCCL_TRY(g, nc.GroupEnd());
        } else if (cudaMemcpyAsync(gather_out[0], mem[0].w_slice, sizeof(double) * (size_t)per, cudaMemcpyDeviceToDevice, g->ctx[0]->stream) != cudaSuccess)
            GRP_FAIL(g, ABCB200_ECUDA, "copy of the gathered weights failed");
    }
    return ABCB200_OK;
}

inline int64_t pad32(int64_t n) { return (n + 31) / 32 * 32; }

}  // namespace

// Device buffers: called by every process of a process-per-GPU group (one local member each). A group whose members all live in
// this process is driven through abcb200_weights_sharded (host buffers) instead.
extern "C" int abcb200_weights_sharded_dev(abcb200_group* g, const double* numer, const double* theta_new, int64_t ld_new, int64_t N_new,
                                           double* theta_old, int64_t ld_old, int64_t N_old, double* w_old, double* dv_old, int P, int algo,
                                           int bcast_root, double* w_gathered, double* w_slice_out) {
    if (!g) return ABCB200_EINVAL;
    g->err[0] = 0;
    if (g->nlocal != 1) GRP_FAIL(g, ABCB200_EINVAL, "weights_sharded_dev: the group drives %d GPUs from this process; use abcb200_weights_sharded", g->nlocal);
    if (!theta_new || !theta_old || !w_old || !dv_old || N_new < 1 || N_old < 1 || P < 1 || ld_new < N_new || ld_old < N_old)
        GRP_FAIL(g, ABCB200_EINVAL, "weights_sharded_dev: bad argument");
    if (bcast_root >= g->world) GRP_FAIL(g, ABCB200_EINVAL, "weights_sharded_dev: broadcast root %d outside the group", bcast_root);
    abcb200_ctx* c = g->ctx[0];
    if (cudaSetDevice(c->device) != cudaSuccess) GRP_FAIL(g, ABCB200_ENODEV, "cudaSetDevice failed");
    const int64_t per = (N_new + g->world - 1) / g->world;
    GRP_CTX_TRY(g, c, ws_reserve(c, weights_ws_bytes(c, per, N_old, P) + 2 * align_up((size_t)per * g->world * 8, 256) + 4096));
    std::vector<Member> mem(1);
    Member& me = mem[0];
    me.numer = numer; me.th_new = theta_new; me.ld_new = ld_new; me.row0 = 0;
    me.th_old = theta_old; me.w_old = w_old; me.dv_old = dv_old; me.ld_old = ld_old;
    me.w_slice = w_slice_out ? w_slice_out : ws_new<double>(c, (size_t)per);
    me.ss = ws_new<double>(c, 1);
    double* gathered = nullptr;
    if (w_gathered) gathered = ((int64_t)per * g->world == N_new) ? w_gathered : ws_new<double>(c, (size_t)per * g->world);
    if (!me.w_slice || !me.ss || (w_gathered && !gathered)) GRP_FAIL(g, ABCB200_ENOMEM, "workspace exhausted in weights_sharded_dev");
    if (cudaMemsetAsync(me.w_slice, 0, sizeof(double) * (size_t)per, c->stream) != cudaSuccess) GRP_FAIL(g, ABCB200_ECUDA, "memset failed");
    stage_begin(c, 7);
    double* go[1] = {gathered};
    const int rc = sharded_core(g, mem, N_new, N_old, P, algo, bcast_root, w_gathered ? go : nullptr);
    if (rc != ABCB200_OK) return rc;
    if (w_gathered && gathered != w_gathered &&
        cudaMemcpyAsync(w_gathered, gathered, sizeof(double) * (size_t)N_new, cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess)
        GRP_FAIL(g, ABCB200_ECUDA, "copy of the gathered weights failed");
    stage_end(c, 7);
    return ABCB200_OK;
}

// Host buffers, one process driving every GPU of the group (what a C++ AbcSmc host calls instead of abcb200_weights).
extern "C" int abcb200_weights_sharded(abcb200_group* g, const double* numer, const double* theta_new, int64_t ld_new, int64_t N_new,
                                       const double* theta_old, int64_t ld_old, int64_t N_old, const double* w_old, const double* dv_old,
                                       int P, int algo, double* w_out) {
    if (!g) return ABCB200_EINVAL;
    g->err[0] = 0;
    if (g->nlocal != g->world) GRP_FAIL(g, ABCB200_EINVAL, "weights_sharded: this process drives %d of the group's %d GPUs; use abcb200_weights_sharded_dev", g->nlocal, g->world);
    if (!theta_new || !theta_old || !w_old || !dv_old || !w_out) GRP_FAIL(g, ABCB200_EINVAL, "weights_sharded: null argument");
    if (N_new < 1 || N_old < 1 || P < 1 || ld_new < N_new || ld_old < N_old) GRP_FAIL(g, ABCB200_EINVAL, "weights_sharded: bad shape");
    const int G = g->world;
    const int64_t per = (N_new + G - 1) / G, ldo = pad32(N_old), lds = pad32(per);
    std::vector<Member> mem((size_t)G);
    for (int m = 0; m < G; m++) {
        abcb200_ctx* c = g->ctx[(size_t)m];
        Member& me = mem[(size_t)m];
        if (cudaSetDevice(c->device) != cudaSuccess) GRP_FAIL(g, ABCB200_ENODEV, "cudaSetDevice failed");
        const size_t need = weights_ws_bytes(c, per, N_old, P) + align_up((size_t)lds * P * 8, 256) + align_up((size_t)ldo * P * 8, 256) +
                            3 * align_up((size_t)per * 8, 256) + align_up((size_t)N_old * 8, 256) + align_up((size_t)P * 8, 256) + 8192;
        GRP_CTX_TRY(g, c, ws_reserve(c, need));
        const int64_t lo = std::min<int64_t>((int64_t)m * per, N_new), hi = std::min<int64_t>(lo + per, N_new);
        double* d_new = ws_new<double>(c, (size_t)lds * P);
        double* d_numer = numer ? ws_new<double>(c, (size_t)per) : nullptr;
        me.th_old = ws_new<double>(c, (size_t)ldo * P);
        me.w_old = ws_new<double>(c, (size_t)N_old);
        me.dv_old = ws_new<double>(c, (size_t)P);
        me.w_slice = ws_new<double>(c, (size_t)per);
        me.ss = ws_new<double>(c, 1);
        if (!d_new || !me.th_old || !me.w_old || !me.dv_old || !me.w_slice || !me.ss || (numer && !d_numer)) GRP_FAIL(g, ABCB200_ENOMEM, "workspace exhausted in weights_sharded");
        me.th_new = d_new; me.numer = d_numer; me.ld_new = lds; me.row0 = lo; me.ld_old = ldo;
        stage_begin(c, 8);
        if (hi > lo) {       // every GPU receives only ITS rows of the new set
            if (cudaMemcpy2DAsync(d_new, (size_t)lds * 8, theta_new + lo, (size_t)ld_new * 8, (size_t)(hi - lo) * 8, (size_t)P, cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
                GRP_FAIL(g, ABCB200_ECUDA, "H2D of the new rows failed");
            if (numer && cudaMemcpyAsync(d_numer, numer + lo, sizeof(double) * (size_t)(hi - lo), cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
                GRP_FAIL(g, ABCB200_ECUDA, "H2D of the numerators failed");
        }
        if (m == 0) {        // the previous set crosses PCIe once; the other GPUs get it by ncclBroadcast
            if (cudaMemcpy2DAsync(me.th_old, (size_t)ldo * 8, theta_old, (size_t)ld_old * 8, (size_t)N_old * 8, (size_t)P, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
                cudaMemcpyAsync(me.w_old, w_old, sizeof(double) * (size_t)N_old, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
                cudaMemcpyAsync(me.dv_old, dv_old, sizeof(double) * (size_t)P, cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
                GRP_FAIL(g, ABCB200_ECUDA, "H2D of the previous set failed");
        }
        stage_end(c, 8);
        stage_begin(c, 7);
    }
    const int rc = sharded_core(g, mem, N_new, N_old, P, algo, 0, nullptr);
    if (rc != ABCB200_OK) return rc;
    for (int m = 0; m < G; m++) {
        abcb200_ctx* c = g->ctx[(size_t)m];
        Member& me = mem[(size_t)m];
        cudaSetDevice(c->device);
        stage_end(c, 7);
        stage_begin(c, 9);
        if (me.hi > me.lo && cudaMemcpyAsync(w_out + me.lo, me.w_slice, sizeof(double) * (size_t)(me.hi - me.lo), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
            GRP_FAIL(g, ABCB200_ECUDA, "D2H of the weights failed");
        stage_end(c, 9);
    }
    for (int m = 0; m < G; m++) {
        abcb200_ctx* c = g->ctx[(size_t)m];
        cudaSetDevice(c->device);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) GRP_FAIL(g, ABCB200_ECUDA, "stream synchronisation failed on GPU %d", c->device);
    }
    return ABCB200_OK;
}
