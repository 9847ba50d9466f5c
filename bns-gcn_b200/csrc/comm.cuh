// comm.cuh -- the collectives of the path behind the C ABI (included by bnsgcn.cu).
//
// SURVEY section 8(b) lists bns_ctx_create / bns_alltoallv_* / bns_allreduce_sum_f32 among the exports a non-torch host
// needs to bind the whole path; round 1 left them in torch.distributed.  They are thin: NCCL does the work.  NCCL is
// resolved at RUN time (dlopen "libnccl.so.2": inside a PyTorch process that is the copy torch already loaded, so both
// share one NCCL; elsewhere the system library), so libbnsgcn.so itself has no link-time dependency on it and still loads
// on a box without NCCL -- only these entry points then fail, loudly.
//   helper/reducer.py:28-49          one all-reduce per parameter          -> bns_allreduce_sum_f32 (one flat bucket)
//   helper/utils.py:187-213          data_transfer(..., tag=NODE), gloo    -> bns_alltoallv_i64
//   helper/feature_buffer.py:101-153 __gloo_all_to_all / __mpi_all_to_all  -> bns_alltoallv_f32 (staged transport)
#include <dlfcn.h>
#include <nccl.h>

struct bns_ctx {
    int32_t rank = 0, world = 0;
    ncclComm_t comm = nullptr;
};

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi &nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle) break;
        }
        if (!a.handle) return a;
#define BNS_NCCL_SYM(field, sym) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, sym))
        BNS_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
        BNS_NCCL_SYM(CommInitRank, "ncclCommInitRank");
        BNS_NCCL_SYM(CommDestroy, "ncclCommDestroy");
        BNS_NCCL_SYM(AllReduce, "ncclAllReduce");
        BNS_NCCL_SYM(Send, "ncclSend");
        BNS_NCCL_SYM(Recv, "ncclRecv");
        BNS_NCCL_SYM(GroupStart, "ncclGroupStart");
        BNS_NCCL_SYM(GroupEnd, "ncclGroupEnd");
        BNS_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef BNS_NCCL_SYM
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.Send && a.Recv && a.GroupStart &&
               a.GroupEnd && a.GetErrorString;
        return a;
    }();
    return api;
}

#define BNS_NCCL(call)                                                                                          \
    do {                                                                                                        \
        ncclResult_t r_ = (call);                                                                               \
        if (r_ != ncclSuccess) return fail(BNS_E_CUDA, "%s failed: %s", #call, nccl_api().GetErrorString(r_)); \
    } while (0)

int require_nccl(const char *who) {
    if (!nccl_api().ok) return fail(BNS_E_UNSUPPORTED, "%s: libnccl.so.2 could not be loaded (%s)", who, dlerror());
    return BNS_OK;
}

template <typename T>
int alltoallv(bns_ctx *c, const T *send, const int64_t *send_counts, const int64_t *send_offsets, T *recv,
              const int64_t *recv_counts, const int64_t *recv_offsets, int64_t width, ncclDataType_t dt, cudaStream_t st,
              const char *who) {
    BNS_REQUIRE(c && c->comm, "%s: NULL context", who);
    BNS_REQUIRE(send_counts && send_offsets && recv_counts && recv_offsets && width > 0, "%s: NULL table", who);
    NcclApi &n = nccl_api();
    BNS_NCCL(n.GroupStart());
    // the reference's ring order (helper/feature_buffer.py:111-113): right = rank + i, left = rank - i
    for (int i = 1; i < c->world; ++i) {
        const int right = (c->rank + i) % c->world, left = (c->rank - i + c->world) % c->world;
        if (send_counts[right] > 0) {
            BNS_REQUIRE(send, "%s: NULL send buffer", who);
            BNS_NCCL(n.Send(send + send_offsets[right] * width, (size_t)(send_counts[right] * width), dt, right, c->comm, st));
        }
        if (recv_counts[left] > 0) {
            BNS_REQUIRE(recv, "%s: NULL receive buffer", who);
            BNS_NCCL(n.Recv(recv + recv_offsets[left] * width, (size_t)(recv_counts[left] * width), dt, left, c->comm, st));
        }
    }
    BNS_NCCL(n.GroupEnd());
    return BNS_OK;
}

}  // namespace

extern "C" int bns_comm_unique_id(void *id_out) {
    BNS_REQUIRE(id_out, "bns_comm_unique_id: NULL output");
    static_assert(sizeof(ncclUniqueId) <= BNS_COMM_ID_BYTES, "unique id size");
    int rc = require_nccl("bns_comm_unique_id");
    if (rc) return rc;
    ncclUniqueId id;
    BNS_NCCL(nccl_api().GetUniqueId(&id));
    memset(id_out, 0, BNS_COMM_ID_BYTES);
    memcpy(id_out, &id, sizeof(id));
    return BNS_OK;
}

extern "C" int bns_ctx_create(bns_ctx_t **out, int32_t rank, int32_t world, const void *unique_id) {
    BNS_REQUIRE(out && unique_id, "bns_ctx_create: NULL argument");
    BNS_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bns_ctx_create: bad rank / world");
    int rc = require_nccl("bns_ctx_create");
    if (rc) return rc;
    bns_ctx *c = new (std::nothrow) bns_ctx();
    if (!c) return fail(BNS_E_INVALID, "bns_ctx_create: out of host memory");
    c->rank = rank; c->world = world;
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    ncclResult_t r = nccl_api().CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        delete c;
        return fail(BNS_E_CUDA, "ncclCommInitRank failed: %s", nccl_api().GetErrorString(r));
    }
    *out = c;
    return BNS_OK;
}

extern "C" int bns_ctx_destroy(bns_ctx_t *c) {
    if (!c) return BNS_OK;
    if (c->comm && nccl_api().ok) nccl_api().CommDestroy(c->comm);
    delete c;
    return BNS_OK;
}

// buf[i] = sum over ranks of buf[i], in place (helper/reducer.py:37 / :46 -- for ONE flat bucket instead of per parameter)
extern "C" int bns_allreduce_sum_f32(bns_ctx_t *c, float *buf, int64_t n, void *stream) {
    BNS_REQUIRE(c && c->comm, "bns_allreduce_sum_f32: NULL context");
    BNS_REQUIRE(n >= 0 && (n == 0 || buf), "bns_allreduce_sum_f32: bad buffer");
    if (n == 0 || c->world == 1) return BNS_OK;
    BNS_NCCL(nccl_api().AllReduce(buf, buf, (size_t)n, ncclFloat32, ncclSum, c->comm, as_stream(stream)));
    return BNS_OK;
}

// rows of `width` elements: rank r sends rows [send_offsets[j], +send_counts[j]) of `send` to j and receives
// rows [recv_offsets[j], +recv_counts[j]) of `recv` from j, for every j != r, as ONE grouped NCCL operation
extern "C" int bns_alltoallv_f32(bns_ctx_t *c, const float *send, const int64_t *send_counts, const int64_t *send_offsets,
                                 float *recv, const int64_t *recv_counts, const int64_t *recv_offsets, int64_t width,
                                 void *stream) {
    return alltoallv<float>(c, send, send_counts, send_offsets, recv, recv_counts, recv_offsets, width, ncclFloat32,
                            as_stream(stream), "bns_alltoallv_f32");
}

// the same exchange for buffers that are separate allocations: peer j gets send_ptrs[j][0 .. send_bytes[j]) and its
// message lands in recv_ptrs[j][0 .. recv_bytes[j])  (host arrays [world] of device pointers / byte counts)
extern "C" int bns_alltoallv_bytes(bns_ctx_t *c, const void *const *send_ptrs, const int64_t *send_bytes, void *const *recv_ptrs,
                                   const int64_t *recv_bytes, void *stream) {
    BNS_REQUIRE(c && c->comm, "bns_alltoallv_bytes: NULL context");
    BNS_REQUIRE(send_ptrs && send_bytes && recv_ptrs && recv_bytes, "bns_alltoallv_bytes: NULL table");
    NcclApi &n = nccl_api();
    cudaStream_t st = as_stream(stream);
    BNS_NCCL(n.GroupStart());
    for (int i = 1; i < c->world; ++i) {
        const int right = (c->rank + i) % c->world, left = (c->rank - i + c->world) % c->world;
        if (send_bytes[right] > 0) {
            BNS_REQUIRE(send_ptrs[right], "bns_alltoallv_bytes: NULL send buffer for peer %d", right);
            BNS_NCCL(n.Send(send_ptrs[right], (size_t)send_bytes[right], ncclInt8, right, c->comm, st));
        }
        if (recv_bytes[left] > 0) {
            BNS_REQUIRE(recv_ptrs[left], "bns_alltoallv_bytes: NULL receive buffer for peer %d", left);
            BNS_NCCL(n.Recv(recv_ptrs[left], (size_t)recv_bytes[left], ncclInt8, left, c->comm, st));
        }
    }
    BNS_NCCL(n.GroupEnd());
    return BNS_OK;
}

extern "C" int bns_alltoallv_i64(bns_ctx_t *c, const int64_t *send, const int64_t *send_counts, const int64_t *send_offsets,
                                 int64_t *recv, const int64_t *recv_counts, const int64_t *recv_offsets, void *stream) {
    return alltoallv<int64_t>(c, send, send_counts, send_offsets, recv, recv_counts, recv_offsets, 1, ncclInt64,
                              as_stream(stream), "bns_alltoallv_i64");
}
