// Engine lifetime, error reporting and the NCCL communicator (loaded lazily
// with dlopen so that a single-GPU deployment needs no libnccl at all).
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "comm.cuh"
#include "common.cuh"

namespace ss {

static thread_local std::string g_error;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
}

// ---- NCCL through dlopen ---------------------------------------------------
struct NcclApi {
  void* handle = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // RTLD_NOLOAD first: inside a torch process reuse the libnccl it already mapped.
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return;
    api.handle = h;
#define SS_SYM(name) api.name = (decltype(api.name))dlsym(h, "nccl" #name)
    SS_SYM(GetUniqueId);
    SS_SYM(CommInitRank);
    SS_SYM(CommDestroy);
    SS_SYM(AllReduce);
    SS_SYM(Broadcast);
    SS_SYM(AllGather);
    SS_SYM(Send);
    SS_SYM(Recv);
    SS_SYM(GroupStart);
    SS_SYM(GroupEnd);
    SS_SYM(GetErrorString);
#undef SS_SYM
  });
  if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.Broadcast ||
      !api.GroupStart || !api.GroupEnd)
    return nullptr;
  return &api;
}

#define SS_NCCL(api, call)                                                         \
  do {                                                                             \
    ncclResult_t _r = (call);                                                      \
    if (_r != ncclSuccess) {                                                       \
      ss::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,                   \
                    (api)->GetErrorString ? (api)->GetErrorString(_r) : "nccl error"); \
      return SS_ERR_NCCL;                                                          \
    }                                                                              \
  } while (0)

}  // namespace ss

struct CommState {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};

void comm_state_free(CommState* c) {
  if (!c) return;
  if (c->comm) {
    auto* api = ss::nccl_api();
    if (api && api->CommDestroy) api->CommDestroy(c->comm);
  }
  delete c;
}

int comm_rank(const ss_engine* e) { return e->comm ? e->comm->rank : 0; }
int comm_world(const ss_engine* e) { return e->comm ? e->comm->world : 1; }

int comm_allreduce_sum_f64_on(ss_engine* e, cudaStream_t st, double* dev_buf, size_t count) {
  if (!e->comm || e->comm->world == 1) return SS_OK;
  auto* api = ss::nccl_api();
  SS_REQUIRE(api, SS_ERR_NCCL, "libnccl not loadable");
  SS_NCCL(api, api->AllReduce(dev_buf, dev_buf, count, ncclDouble, ncclSum, e->comm->comm, st));
  return SS_OK;
}
int comm_allreduce_sum_f64(ss_engine* e, double* dev_buf, size_t count) {
  return comm_allreduce_sum_f64_on(e, e->stream, dev_buf, count);
}

int comm_allgatherv_bytes(ss_engine* e, void* dev_buf, const size_t* byte_off, const size_t* byte_cnt) {
  return comm_allgatherv_bytes_on(e, e->stream, dev_buf, byte_off, byte_cnt);
}
int comm_allgatherv_bytes_on(ss_engine* e, cudaStream_t st, void* dev_buf, const size_t* byte_off,
                             const size_t* byte_cnt) {
  if (!e->comm || e->comm->world == 1) return SS_OK;
  auto* api = ss::nccl_api();
  SS_REQUIRE(api, SS_ERR_NCCL, "libnccl not loadable");
  SS_NCCL(api, api->GroupStart());
  ncclResult_t first = ncclSuccess;
  for (int r = 0; r < e->comm->world; ++r) {
    if (byte_cnt[r] == 0) continue;
    char* p = (char*)dev_buf + byte_off[r];
    const ncclResult_t rc = api->Broadcast(p, p, byte_cnt[r], ncclChar, r, e->comm->comm, st);
    if (rc != ncclSuccess && first == ncclSuccess) first = rc;
  }
  const ncclResult_t end = api->GroupEnd();  // always closed, also on the error path
  SS_NCCL(api, first);
  SS_NCCL(api, end);
  return SS_OK;
}

int comm_allreduce_sum_u32(ss_engine* e, uint32_t* dev_buf, size_t count) {
  if (!e->comm || e->comm->world == 1) return SS_OK;
  auto* api = ss::nccl_api();
  SS_REQUIRE(api, SS_ERR_NCCL, "libnccl not loadable");
  SS_NCCL(api, api->AllReduce(dev_buf, dev_buf, count, ncclUint32, ncclSum, e->comm->comm, e->stream));
  return SS_OK;
}

int comm_alltoallv_u32(ss_engine* e, int n, const uint32_t* const* send, const size_t* send_off, const size_t* send_cnt,
                       uint32_t* const* recv, const size_t* recv_off, const size_t* recv_cnt) {
  const int world = comm_world(e), rank = comm_rank(e);
  if (world == 1) {
    for (int i = 0; i < n; ++i)
      if (send_cnt[0])
        SS_CUDA(cudaMemcpyAsync(recv[i] + recv_off[0], send[i] + send_off[0], send_cnt[0] * 4, cudaMemcpyDeviceToDevice,
                                e->stream));
    return SS_OK;
  }
  auto* api = ss::nccl_api();
  SS_REQUIRE(api && api->Send && api->Recv, SS_ERR_NCCL, "libnccl not loadable (Send/Recv)");
  (void)rank;
  SS_NCCL(api, api->GroupStart());
  ncclResult_t first = ncclSuccess;
  for (int i = 0; i < n; ++i)
    for (int r = 0; r < world; ++r) {
      ncclResult_t rc = ncclSuccess;
      if (send_cnt[r]) rc = api->Send(send[i] + send_off[r], send_cnt[r], ncclUint32, r, e->comm->comm, e->stream);
      if (rc != ncclSuccess && first == ncclSuccess) first = rc;
      rc = ncclSuccess;
      if (recv_cnt[r]) rc = api->Recv(recv[i] + recv_off[r], recv_cnt[r], ncclUint32, r, e->comm->comm, e->stream);
      if (rc != ncclSuccess && first == ncclSuccess) first = rc;
    }
  const ncclResult_t end = api->GroupEnd();
  SS_NCCL(api, first);
  SS_NCCL(api, end);
  return SS_OK;
}

int comm_allgather_dev(ss_engine* e, int n, const void* const* in, void* const* out, const size_t* bytes) {
  if (!e->comm || e->comm->world == 1) {
    for (int i = 0; i < n; ++i)
      if (bytes[i]) SS_CUDA(cudaMemcpyAsync(out[i], in[i], bytes[i], cudaMemcpyDeviceToDevice, e->stream));
    return SS_OK;
  }
  auto* api = ss::nccl_api();
  SS_REQUIRE(api && api->AllGather, SS_ERR_NCCL, "libnccl not loadable");
  SS_NCCL(api, api->GroupStart());
  ncclResult_t first = ncclSuccess;
  for (int i = 0; i < n; ++i) {
    if (!bytes[i]) continue;
    const ncclResult_t r = api->AllGather(in[i], out[i], bytes[i], ncclChar, e->comm->comm, e->stream);
    if (r != ncclSuccess && first == ncclSuccess) first = r;
  }
  const ncclResult_t end = api->GroupEnd();  // always closed, also on the error path
  SS_NCCL(api, first);
  SS_NCCL(api, end);
  return SS_OK;
}

int comm_allgather_host_bytes(ss_engine* e, const void* in, size_t bytes, void* out) {
  const int world = comm_world(e);
  if (world == 1) {
    memcpy(out, in, bytes);
    return SS_OK;
  }
  auto* api = ss::nccl_api();
  SS_REQUIRE(api && api->AllGather, SS_ERR_NCCL, "libnccl not loadable");
  ss::DevBuf<char> d_in, d_out;
  SS_TRY(d_in.alloc(bytes));
  SS_TRY(d_out.alloc(bytes * world));
  SS_CUDA(cudaMemcpyAsync(d_in.p, in, bytes, cudaMemcpyHostToDevice, e->stream));
  SS_NCCL(api, api->AllGather(d_in.p, d_out.p, bytes, ncclChar, e->comm->comm, e->stream));
  SS_CUDA(cudaMemcpyAsync(out, d_out.p, bytes * world, cudaMemcpyDeviceToHost, e->stream));
  SS_CUDA(cudaStreamSynchronize(e->stream));
  return SS_OK;
}

extern "C" {

SS_API int ss_version(void) { return 100; }

SS_API const char* ss_last_error(void) { return ss::g_error.c_str(); }

SS_API int ss_create(const ss_config* cfg, ss_engine** out) {
  SS_REQUIRE(out, SS_ERR_INVALID, "ss_create: out is NULL");
  *out = nullptr;
  int dev = cfg ? cfg->device : 0;
  int count = 0;
  cudaError_t err = cudaGetDeviceCount(&count);
  if (err != cudaSuccess || count == 0) {
    cudaGetLastError();
    ss::set_error("ss_create: no CUDA device (%s); this engine has no CPU path",
                  err == cudaSuccess ? "device count 0" : cudaGetErrorString(err));
    return SS_ERR_NO_DEVICE;
  }
  SS_REQUIRE(dev >= 0 && dev < count, SS_ERR_INVALID, "ss_create: device %d of %d", dev, count);
  cudaDeviceProp prop;
  SS_CUDA(cudaGetDeviceProperties(&prop, dev));
  SS_REQUIRE(prop.major == 10, SS_ERR_NO_DEVICE,
             "ss_create: device %d is sm_%d%d; kernels are built for sm_100a only", dev, prop.major,
             prop.minor);
  DeviceGuard guard(dev);
  ss_engine* e = new (std::nothrow) ss_engine();
  SS_REQUIRE(e, SS_ERR_OOM, "ss_create: host allocation failed");
  e->device = dev;
  e->flags = cfg ? cfg->flags : 0;
  e->sm_count = prop.multiProcessorCount;
  err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
  if (err != cudaSuccess) {
    ss::set_error("cudaStreamCreate -> %s", cudaGetErrorString(err));
    delete e;
    return SS_ERR_CUDA;
  }
  *out = e;
  return SS_OK;
}

SS_API void ss_destroy(ss_engine* e) {
  if (!e) return;
  DeviceGuard guard(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  pagerank_state_free(e->pr);
  index_state_free(e->idx);
  comm_state_free(e->comm);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

SS_API void* ss_stream_handle(ss_engine* e) { return e ? (void*)e->stream : nullptr; }

SS_API int ss_comm_unique_id(void* id128) {
  SS_REQUIRE(id128, SS_ERR_INVALID, "ss_comm_unique_id: NULL");
  auto* api = ss::nccl_api();
  SS_REQUIRE(api, SS_ERR_NCCL, "libnccl.so.2 not loadable: %s", dlerror());
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  SS_NCCL(api, api->GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return SS_OK;
}

SS_API int ss_comm_init(ss_engine* e, const void* id128, int32_t rank, int32_t world) {
  SS_REQUIRE(e && id128, SS_ERR_INVALID, "ss_comm_init: NULL argument");
  SS_REQUIRE(world >= 1 && rank >= 0 && rank < world, SS_ERR_INVALID, "ss_comm_init: rank %d of %d",
             rank, world);
  std::lock_guard<std::mutex> lock(e->mu);
  DeviceGuard guard(e->device);
  SS_REQUIRE(!e->comm, SS_ERR_STATE, "ss_comm_init: already initialised");
  SS_REQUIRE(!e->pr && !e->idx, SS_ERR_STATE, "ss_comm_init must precede the loads");
  CommState* c = new (std::nothrow) CommState();
  SS_REQUIRE(c, SS_ERR_OOM, "host allocation failed");
  c->rank = rank;
  c->world = world;
  if (world > 1) {
    auto* api = ss::nccl_api();
    if (!api) {
      delete c;
      ss::set_error("libnccl.so.2 not loadable");
      return SS_ERR_NCCL;
    }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclResult_t r = api->CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
      ss::set_error("ncclCommInitRank -> %s", api->GetErrorString ? api->GetErrorString(r) : "error");
      delete c;
      return SS_ERR_NCCL;
    }
  }
  e->comm = c;
  return SS_OK;
}

}  // extern "C"
