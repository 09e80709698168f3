// device_comm.cu -- the multi-GPU group of the C ABI (include/yart.h): yart_comm_* and yart_film_*.
//
// The reference gathers its 64 tiles over an mpsc channel (main.rs:629-646, 746-760).  Here every GPU renders a
// SAMPLE RANGE of the whole frame into its own f64 XYZ film and the films are summed onto the root
// with ONE collective: ncclReduce, in place, on the context's own stream -- the film the render kernels accumulate
// into IS the buffer NCCL reduces, there is no staging copy and no host round trip.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"), not at link time: a host that never calls yart_comm_* needs no
// NCCL at all, and inside a process that already carries a libnccl (PyTorch bundles its own) the SAME copy is used
// instead of a second one.  There is no non-NCCL fallback: when the library cannot be loaded yart_comm_* fail with
// YART_ERR_UNSUPPORTED and say why.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "host_common.h"

namespace yart {
// accessors of the opaque context (defined in yart_device.cu)
cudaStream_t ctx_stream(yart_ctx* ctx);
int ctx_device(const yart_ctx* ctx);
void ctx_set_error(yart_ctx* ctx, const std::string& msg);
} // namespace yart

namespace {

struct NcclApi {
  void* handle = nullptr;
  std::string error;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi api; // process-wide; filled once

NcclApi* nccl() {
  static std::once_flag once;
  std::call_once(once, [] {
    // 1. a libnccl the process already carries (PyTorch's bundled copy, the host application's): use THAT one --
    //    two different NCCL builds under one SONAME in a process break whichever comes second;
    // 2. $YART_NCCL_LIB;  3. the system's libnccl.so.2.
    api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    const char* names[] = {getenv("YART_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (api.handle) break;
      if (!n || !*n) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (!api.handle) api.error = dlerror();
    }
    if (!api.handle) {
      api.error = "cannot load NCCL (" + api.error + "); set YART_NCCL_LIB to a libnccl.so.2";
      return;
    }
    bool ok = true;
    auto sym = [&](const char* name) {
      void* p = dlsym(api.handle, name);
      if (!p) {
        ok = false;
        api.error = std::string("NCCL library lacks ") + name;
      }
      return p;
    };
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.Reduce = reinterpret_cast<decltype(api.Reduce)>(sym("ncclReduce"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    if (!ok) {
      dlclose(api.handle);
      api.handle = nullptr;
    }
  });
  return api.handle ? &api : nullptr;
}

} // namespace

// A communicator: one NCCL rank per local context.  One-process-per-GPU hosts hold one local rank
// (yart_comm_init_rank); a single process driving N GPUs holds N (yart_comm_init).
struct yart_comm {
  std::vector<yart_ctx*> ctxs;
  std::vector<ncclComm_t> comms;
  std::vector<int> ranks;
  int n_ranks = 0;
  std::string err;
};

namespace {

int fail(yart_comm* c, yart_ctx* ctx, int code, const std::string& msg) {
  if (c) c->err = msg;
  if (ctx) yart::ctx_set_error(ctx, msg);
  yart::set_global_error(msg);
  return code;
}

#define NCCL_TRY(c, ctx, expr)                                                                          \
  do {                                                                                                  \
    ncclResult_t r__ = (expr);                                                                          \
    if (r__ != ncclSuccess)                                                                             \
      return fail(c, ctx, YART_ERR_CUDA, std::string(#expr) + ": " + nccl()->GetErrorString(r__));      \
  } while (0)

} // namespace

extern "C" {

int yart_comm_unique_id(uint8_t id[YART_COMM_ID_BYTES]) {
  if (!id) return fail(nullptr, nullptr, YART_ERR_INVALID, "yart_comm_unique_id: null argument");
  NcclApi* N = nccl();
  if (!N) return fail(nullptr, nullptr, YART_ERR_UNSUPPORTED, "yart_comm_unique_id: " + api.error);
  static_assert(sizeof(ncclUniqueId) <= YART_COMM_ID_BYTES, "ncclUniqueId must fit the ABI's id buffer");
  ncclUniqueId u;
  NCCL_TRY(nullptr, nullptr, N->GetUniqueId(&u));
  memset(id, 0, YART_COMM_ID_BYTES);
  memcpy(id, &u, sizeof(u));
  return YART_OK;
}

int yart_comm_init_rank(yart_ctx* ctx, const uint8_t id[YART_COMM_ID_BYTES], int rank, int n_ranks, yart_comm** out) {
  if (!ctx || !id || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks)
    return fail(nullptr, ctx, YART_ERR_INVALID, "yart_comm_init_rank: bad argument");
  NcclApi* N = nccl();
  if (!N) return fail(nullptr, ctx, YART_ERR_UNSUPPORTED, "yart_comm_init_rank: " + api.error);
  if (cudaSetDevice(yart::ctx_device(ctx)) != cudaSuccess) return fail(nullptr, ctx, YART_ERR_CUDA, "yart_comm_init_rank: cudaSetDevice failed");
  yart_comm* c = new (std::nothrow) yart_comm();
  if (!c) return YART_ERR_NOMEM;
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  ncclComm_t comm = nullptr;
  ncclResult_t r = N->CommInitRank(&comm, n_ranks, u, rank);
  if (r != ncclSuccess) {
    delete c;
    return fail(nullptr, ctx, YART_ERR_CUDA, std::string("ncclCommInitRank: ") + N->GetErrorString(r));
  }
  c->ctxs.push_back(ctx);
  c->comms.push_back(comm);
  c->ranks.push_back(rank);
  c->n_ranks = n_ranks;
  *out = c;
  return YART_OK;
}

int yart_comm_init(yart_ctx* const* ctxs, int n, yart_comm** out) {
  if (!ctxs || !out || n < 1) return fail(nullptr, nullptr, YART_ERR_INVALID, "yart_comm_init: bad argument");
  NcclApi* N = nccl();
  if (!N) return fail(nullptr, ctxs[0], YART_ERR_UNSUPPORTED, "yart_comm_init: " + api.error);
  std::vector<int> devs(n);
  for (int i = 0; i < n; ++i) {
    if (!ctxs[i]) return fail(nullptr, nullptr, YART_ERR_INVALID, "yart_comm_init: null context");
    devs[i] = yart::ctx_device(ctxs[i]);
    for (int j = 0; j < i; ++j)
      if (devs[j] == devs[i]) return fail(nullptr, ctxs[i], YART_ERR_INVALID, "yart_comm_init: two contexts on the same GPU (one rank per GPU)");
  }
  yart_comm* c = new (std::nothrow) yart_comm();
  if (!c) return YART_ERR_NOMEM;
  c->comms.resize(n);
  ncclResult_t r = N->CommInitAll(c->comms.data(), n, devs.data());
  if (r != ncclSuccess) {
    delete c;
    return fail(nullptr, ctxs[0], YART_ERR_CUDA, std::string("ncclCommInitAll: ") + N->GetErrorString(r));
  }
  for (int i = 0; i < n; ++i) {
    c->ctxs.push_back(ctxs[i]);
    c->ranks.push_back(i);
  }
  c->n_ranks = n;
  *out = c;
  return YART_OK;
}

void yart_comm_destroy(yart_comm* c) {
  if (!c) return;
  NcclApi* N = nccl();
  for (size_t i = 0; i < c->comms.size(); ++i) {
    cudaSetDevice(yart::ctx_device(c->ctxs[i]));
    cudaStreamSynchronize(yart::ctx_stream(c->ctxs[i]));
    if (N && c->comms[i]) N->CommDestroy(c->comms[i]);
  }
  delete c;
}

int yart_comm_info(const yart_comm* c, int* n_ranks, int* n_local, int* first_local_rank, int* nccl_version) {
  if (!c) return YART_ERR_INVALID;
  if (n_ranks) *n_ranks = c->n_ranks;
  if (n_local) *n_local = (int)c->comms.size();
  if (first_local_rank) *first_local_rank = c->ranks.empty() ? -1 : c->ranks[0];
  if (nccl_version) {
    *nccl_version = 0;
    if (NcclApi* N = nccl()) N->GetVersion(nccl_version);
  }
  return YART_OK;
}

const char* yart_comm_last_error(const yart_comm* c) { return c ? c->err.c_str() : yart_last_error_global(); }

// Sum of the per-rank films onto `root` (or onto every rank when root < 0), in place, asynchronous on each local
// context's stream: a later yart_film_read / yart_film_finalize on the same context is ordered after it.
int yart_film_reduce(yart_comm* c, double* const* dev_films, uint32_t width, uint32_t height, int root) {
  if (!c || !dev_films || !width || !height || root >= c->n_ranks)
    return fail(c, nullptr, YART_ERR_INVALID, "yart_film_reduce: bad argument");
  NcclApi* N = nccl();
  if (!N) return fail(c, nullptr, YART_ERR_UNSUPPORTED, "yart_film_reduce: " + api.error);
  const size_t count = (size_t)width * height * 3;
  const size_t n_local = c->comms.size();
  for (size_t i = 0; i < n_local; ++i)
    if (!dev_films[i]) return fail(c, c->ctxs[i], YART_ERR_INVALID, "yart_film_reduce: null film");
  if (n_local > 1) NCCL_TRY(c, c->ctxs[0], N->GroupStart());
  for (size_t i = 0; i < n_local; ++i) {
    cudaStream_t s = yart::ctx_stream(c->ctxs[i]);
    ncclResult_t r = root < 0 ? N->AllReduce(dev_films[i], dev_films[i], count, ncclDouble, ncclSum, c->comms[i], s)
                              : N->Reduce(dev_films[i], dev_films[i], count, ncclDouble, ncclSum, root, c->comms[i], s);
    if (r != ncclSuccess) {
      if (n_local > 1) N->GroupEnd();
      return fail(c, c->ctxs[i], YART_ERR_CUDA, std::string("ncclReduce: ") + N->GetErrorString(r));
    }
  }
  if (n_local > 1) NCCL_TRY(c, c->ctxs[0], N->GroupEnd());
  return YART_OK;
}

// ---- device-resident films for hosts that have no CUDA of their own (the Rust / C callers) ----
int yart_film_create(yart_ctx* ctx, uint32_t width, uint32_t height, double** dev_film) {
  if (!ctx || !dev_film || !width || !height) return fail(nullptr, ctx, YART_ERR_INVALID, "yart_film_create: bad argument");
  const size_t bytes = (size_t)width * height * 3 * sizeof(double);
  if (cudaSetDevice(yart::ctx_device(ctx)) != cudaSuccess) return fail(nullptr, ctx, YART_ERR_CUDA, "yart_film_create: cudaSetDevice failed");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return fail(nullptr, ctx, e == cudaErrorMemoryAllocation ? YART_ERR_NOMEM : YART_ERR_CUDA, std::string("yart_film_create: ") + cudaGetErrorString(e));
  e = cudaMemsetAsync(p, 0, bytes, yart::ctx_stream(ctx));
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(nullptr, ctx, YART_ERR_CUDA, std::string("yart_film_create: ") + cudaGetErrorString(e));
  }
  *dev_film = reinterpret_cast<double*>(p);
  return YART_OK;
}

int yart_film_clear(yart_ctx* ctx, double* dev_film, uint32_t width, uint32_t height) {
  if (!ctx || !dev_film || !width || !height) return fail(nullptr, ctx, YART_ERR_INVALID, "yart_film_clear: bad argument");
  if (cudaSetDevice(yart::ctx_device(ctx)) != cudaSuccess) return fail(nullptr, ctx, YART_ERR_CUDA, "yart_film_clear: cudaSetDevice failed");
  cudaError_t e = cudaMemsetAsync(dev_film, 0, (size_t)width * height * 3 * sizeof(double), yart::ctx_stream(ctx));
  if (e != cudaSuccess) return fail(nullptr, ctx, YART_ERR_CUDA, std::string("yart_film_clear: ") + cudaGetErrorString(e));
  return YART_OK;
}

int yart_film_read(yart_ctx* ctx, const double* dev_film, uint32_t width, uint32_t height, double* host_film) {
  if (!ctx || !dev_film || !host_film || !width || !height) return fail(nullptr, ctx, YART_ERR_INVALID, "yart_film_read: bad argument");
  if (cudaSetDevice(yart::ctx_device(ctx)) != cudaSuccess) return fail(nullptr, ctx, YART_ERR_CUDA, "yart_film_read: cudaSetDevice failed");
  cudaStream_t s = yart::ctx_stream(ctx);
  cudaError_t e = cudaMemcpyAsync(host_film, dev_film, (size_t)width * height * 3 * sizeof(double), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return fail(nullptr, ctx, YART_ERR_CUDA, std::string("yart_film_read: ") + cudaGetErrorString(e));
  return YART_OK;
}

void yart_film_destroy(yart_ctx* ctx, double* dev_film) {
  if (!ctx || !dev_film) return;
  cudaSetDevice(yart::ctx_device(ctx));
  cudaStreamSynchronize(yart::ctx_stream(ctx));
  cudaFree(dev_film);
}

} // extern "C"
