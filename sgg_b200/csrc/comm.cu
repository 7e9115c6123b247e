// Data-parallel gradient exchange: a thin NCCL communicator owned by the library.
// The reference is single-GPU (train:412-418); this is the one exchange step data parallelism adds:
// sum of the flat gradient bucket over ranks before every optimiser step, over NVLink 5 / NVSwitch.
// libnccl.so.2 is resolved at run time (dlopen) so that the library loads on machines without NCCL and uses the
// same NCCL build as the hosting PyTorch process.
#include <dlfcn.h>

#include "common.cuh"
#include "../../include/sgg_b200.h"

namespace sgg {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[SGG_COMM_ID_BYTES]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclFloat32 = 7, ncclSum = 0 };

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*);   // NCCL >= 2.18 (optional)
  ncclResult_t (*ReduceScatter)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char* (*GetErrorString)(ncclResult_t);
  bool ok;
};

static NcclApi* nccl() {
  static NcclApi api{};
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
      api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
      api.CommSplit = (decltype(api.CommSplit))dlsym(h, "ncclCommSplit");
      api.ReduceScatter = (decltype(api.ReduceScatter))dlsym(h, "ncclReduceScatter");
      api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
      api.Send = (decltype(api.Send))dlsym(h, "ncclSend");
      api.Recv = (decltype(api.Recv))dlsym(h, "ncclRecv");
      api.GroupStart = (decltype(api.GroupStart))dlsym(h, "ncclGroupStart");
      api.GroupEnd = (decltype(api.GroupEnd))dlsym(h, "ncclGroupEnd");
      api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString &&
               api.ReduceScatter && api.AllGather && api.Send && api.Recv && api.GroupStart && api.GroupEnd;
    }
  }
  return api.ok ? &api : nullptr;
}

#define SGG_NCCL(call)                                                                       \
  do {                                                                                       \
    ncclResult_t r__ = (call);                                                               \
    if (r__ != 0) {                                                                          \
      ::sgg::set_error("%s failed: %s", #call, nccl()->GetErrorString(r__));                 \
      return -3;                                                                             \
    }                                                                                        \
  } while (0)

// Two communicators over the same ranks: collectives on one communicator must be issued in one order on one stream,
// so the side stream of an optimiser step (plan.cu) gets its own "lane".
struct Comm {
  ncclComm_t nccl;
  ncclComm_t side;   // may be null (ncclCommSplit unavailable): callers then serialise on lane 0
  int rank, world;
};

bool comm_has_side_lane(void* comm) {
  Comm* c = reinterpret_cast<Comm*>(comm);
  return c && (c->world == 1 || c->side != nullptr);
}

int comm_allreduce(void* comm, float* buf, long long n, cudaStream_t st, int lane) {
  Comm* c = reinterpret_cast<Comm*>(comm);
  SGG_CHECK(c && buf && n >= 0, "comm_allreduce: bad argument");
  if (c->world == 1 || n == 0) return 0;
  SGG_CHECK(lane == 0 || c->side, "comm_allreduce: no side communicator");
  SGG_NCCL(nccl()->AllReduce(buf, buf, (size_t)n, ncclFloat32, ncclSum, lane == 0 ? c->nccl : c->side, st));
  return 0;
}

int comm_rank(void* comm) { return comm ? reinterpret_cast<Comm*>(comm)->rank : 0; }
int comm_world(void* comm) { return comm ? reinterpret_cast<Comm*>(comm)->world : 1; }

// out[n_per_rank] = sum over ranks of their in[rank * n_per_rank ...]
int comm_reduce_scatter(void* comm, const float* in, float* out, long long n_per_rank, cudaStream_t st, int lane) {
  Comm* c = reinterpret_cast<Comm*>(comm);
  SGG_CHECK(c && in && out && n_per_rank >= 0, "comm_reduce_scatter: bad argument");
  SGG_CHECK(lane == 0 || c->side, "comm_reduce_scatter: no side communicator");
  SGG_NCCL(nccl()->ReduceScatter(in, out, (size_t)n_per_rank, ncclFloat32, ncclSum, lane == 0 ? c->nccl : c->side, st));
  return 0;
}
// out[rank * n_per_rank ...] = every rank's in[n_per_rank]
int comm_all_gather(void* comm, const float* in, float* out, long long n_per_rank, cudaStream_t st, int lane) {
  Comm* c = reinterpret_cast<Comm*>(comm);
  SGG_CHECK(c && in && out && n_per_rank >= 0, "comm_all_gather: bad argument");
  SGG_CHECK(lane == 0 || c->side, "comm_all_gather: no side communicator");
  SGG_NCCL(nccl()->AllGather(in, out, (size_t)n_per_rank, ncclFloat32, lane == 0 ? c->nccl : c->side, st));
  return 0;
}
// Block p (bytes_per_rank bytes) of `send` goes to rank p; block p of `recv` comes from rank p.
int comm_all_to_all(void* comm, const void* send, void* recv, long long bytes_per_rank, cudaStream_t st, int lane) {
  Comm* c = reinterpret_cast<Comm*>(comm);
  SGG_CHECK(c && send && recv && bytes_per_rank >= 0, "comm_all_to_all: bad argument");
  SGG_CHECK(lane == 0 || c->side, "comm_all_to_all: no side communicator");
  ncclComm_t nc = lane == 0 ? c->nccl : c->side;
  SGG_NCCL(nccl()->GroupStart());
  for (int p = 0; p < c->world; ++p) {
    const char* sp = reinterpret_cast<const char*>(send) + (long long)p * bytes_per_rank;
    char* rp = reinterpret_cast<char*>(recv) + (long long)p * bytes_per_rank;
    ncclResult_t r1 = nccl()->Send(sp, (size_t)bytes_per_rank, ncclInt8, p, nc, st);
    ncclResult_t r2 = nccl()->Recv(rp, (size_t)bytes_per_rank, ncclInt8, p, nc, st);
    if (r1 != 0 || r2 != 0) {
      nccl()->GroupEnd();
      set_error("ncclSend/ncclRecv failed: %s", nccl()->GetErrorString(r1 != 0 ? r1 : r2));
      return -3;
    }
  }
  SGG_NCCL(nccl()->GroupEnd());
  return 0;
}

}  // namespace sgg

using namespace sgg;

extern "C" int sgg_comm_unique_id(void* id_out) {
  SGG_CHECK(id_out != nullptr, "sgg_comm_unique_id: null output");
  SGG_CHECK(nccl() != nullptr, "sgg_comm_unique_id: libnccl.so.2 not found in this process");
  SGG_NCCL(nccl()->GetUniqueId(reinterpret_cast<ncclUniqueId*>(id_out)));
  return 0;
}

extern "C" int sgg_comm_init(const void* id, int32_t rank, int32_t world, void** comm_out) {
  SGG_CHECK(id && comm_out, "sgg_comm_init: null argument");
  SGG_CHECK(world >= 1 && rank >= 0 && rank < world, "sgg_comm_init: bad rank %d / world %d", rank, world);
  SGG_CHECK(nccl() != nullptr, "sgg_comm_init: libnccl.so.2 not found in this process");
  Comm* c = new Comm{nullptr, nullptr, rank, world};
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  ncclResult_t r = nccl()->CommInitRank(&c->nccl, world, uid, rank);
  if (r != 0) {
    set_error("ncclCommInitRank failed: %s", nccl()->GetErrorString(r));
    delete c;
    return -3;
  }
  if (world > 1 && nccl()->CommSplit) {
    if (nccl()->CommSplit(c->nccl, 0, rank, &c->side, nullptr) != 0) c->side = nullptr;
  }
  *comm_out = c;
  return 0;
}

extern "C" int sgg_comm_destroy(void* comm) {
  Comm* c = reinterpret_cast<Comm*>(comm);
  if (!c) return 0;
  if (c->side && nccl()) nccl()->CommDestroy(c->side);
  if (c->nccl && nccl()) nccl()->CommDestroy(c->nccl);
  delete c;
  return 0;
}

extern "C" int sgg_comm_allreduce_sum(void* comm, float* buf, int64_t n, sgg_stream_t stream) {
  return comm_allreduce(comm, buf, (long long)n, reinterpret_cast<cudaStream_t>(stream), 0);
}
