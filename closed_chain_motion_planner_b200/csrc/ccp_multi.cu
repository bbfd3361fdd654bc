// ccp_multi.cu — multi-GPU entry points of the C ABI for a C++ host (include/ccp.h, "multi-GPU" section).
//
// The reference planner is ONE C++ process (src/main.cpp:27-63).  Two ways for such a host to use every GPU of the box:
//   * ccp_peer_group_*: one process, one handle per GPU.  Every device holds a pool double[world][capacity][n] that
//     its peers can write (cudaDeviceEnablePeerAccess); the projection kernels' epilogues store each converged state
//     into every pool (ccp_set_gather_peers, P2P stores over NVLink), the 8-byte counts follow (ccp_publish_count): the
//     all-gather of SURVEY §8e happens inside the projection kernels, no collective library involved.
//   * ccp_allgather_converged: one process per GPU with an NCCL communicator the host already owns — counts, then the
//     padded compacted states.  NCCL is resolved at run time from the process image (so the communicator and the
//     functions come from the same library the host linked) or from libnccl.so.2; libccp.so itself does not link it.
// Built on the public entry points only (no access to the handle's internals).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <new>

#include "ccp.h"

#define CCP_MAX_GROUP 8

struct ccp_peer_group {
  int world;
  int n;  // doubles per state
  int64_t capacity;
  ccp_handle* h[CCP_MAX_GROUP];
  int dev[CCP_MAX_GROUP];
  double* pool[CCP_MAX_GROUP];     // on device r: double[world][capacity][n]
  int64_t* counts[CCP_MAX_GROUP];  // on device r: int64[world]
  int64_t* n_ok[CCP_MAX_GROUP];    // on device r: this rank's converged count
  cudaStream_t stream[CCP_MAX_GROUP];
  char err[256];
};

static thread_local char g_multi_err[256] = "";

static int multi_err(ccp_peer_group* g, int code, const char* what, const char* detail = "") {
  snprintf(g ? g->err : g_multi_err, 256, "%s%s%s", what, detail[0] ? ": " : "", detail);
  return code;
}

struct dev_scope {
  int prev = -1;
  explicit dev_scope(int d) {
    cudaGetDevice(&prev);
    if (prev != d) cudaSetDevice(d);
  }
  ~dev_scope() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

extern "C" {

const char* ccp_peer_group_last_error(const ccp_peer_group* g) { return g ? g->err : g_multi_err; }

void ccp_peer_group_destroy(ccp_peer_group* g) {
  if (!g) return;
  for (int r = 0; r < g->world; ++r) {
    dev_scope s(g->dev[r]);
    if (g->stream[r]) cudaStreamSynchronize(g->stream[r]);
    if (g->h[r]) ccp_set_gather_peers(g->h[r], 0, 0, nullptr, 0);
    if (g->pool[r]) cudaFree(g->pool[r]);
    if (g->counts[r]) cudaFree(g->counts[r]);
    if (g->n_ok[r]) cudaFree(g->n_ok[r]);
    if (g->stream[r]) cudaStreamDestroy(g->stream[r]);
  }
  delete g;
}

int ccp_peer_group_create(ccp_handle* const* handles, int32_t world, int64_t capacity, ccp_peer_group** out) {
  if (!handles || !out) return multi_err(nullptr, CCP_ERR_INVALID, "null argument");
  *out = nullptr;
  if (world < 1 || world > CCP_MAX_GROUP) return multi_err(nullptr, CCP_ERR_INVALID, "peer group: 1 <= world <= 8");
  if (capacity < 1) return multi_err(nullptr, CCP_ERR_INVALID, "peer group: capacity >= 1");
  ccp_peer_group* g = new (std::nothrow) ccp_peer_group();
  if (!g) return multi_err(nullptr, CCP_ERR_INVALID, "out of memory");
  memset(g, 0, sizeof *g);
  g->world = world;
  g->capacity = capacity;
  for (int r = 0; r < world; ++r) {
    if (!handles[r]) {
      delete g;
      return multi_err(nullptr, CCP_ERR_INVALID, "peer group: null handle");
    }
    g->h[r] = handles[r];
    g->dev[r] = ccp_device(handles[r]);
    const int n = 7 * ccp_n_arms(handles[r]);
    if (r == 0) g->n = n;
    bool dup = false;
    for (int q = 0; q < r; ++q) dup = dup || g->dev[q] == g->dev[r];
    if (n != g->n || dup) {
      delete g;
      return multi_err(nullptr, CCP_ERR_INVALID, "peer group: handles must share the model size and sit on distinct devices");
    }
  }
  cudaError_t e = cudaSuccess;
  for (int r = 0; r < world && e == cudaSuccess; ++r) {
    dev_scope s(g->dev[r]);
    for (int q = 0; q < world && e == cudaSuccess; ++q) {
      if (q == r) continue;
      int can = 0;
      e = cudaDeviceCanAccessPeer(&can, g->dev[r], g->dev[q]);
      if (e == cudaSuccess && !can) {
        ccp_peer_group_destroy(g);
        return multi_err(nullptr, CCP_ERR_CUDA, "peer group: no peer access between two of the devices");
      }
      if (e == cudaSuccess) {
        e = cudaDeviceEnablePeerAccess(g->dev[q], 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) {
          cudaGetLastError();
          e = cudaSuccess;
        }
      }
    }
    const size_t bytes = sizeof(double) * (size_t)world * (size_t)capacity * (size_t)g->n;
    if (e == cudaSuccess) e = cudaMalloc((void**)&g->pool[r], bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&g->counts[r], sizeof(int64_t) * world);
    if (e == cudaSuccess) e = cudaMemset(g->counts[r], 0, sizeof(int64_t) * world);
    if (e == cudaSuccess) e = cudaMalloc((void**)&g->n_ok[r], sizeof(int64_t));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g->stream[r], cudaStreamNonBlocking);
  }
  if (e != cudaSuccess) {
    multi_err(nullptr, CCP_ERR_CUDA, "peer group", cudaGetErrorString(e));
    ccp_peer_group_destroy(g);
    return CCP_ERR_CUDA;
  }
  *out = g;
  return CCP_OK;
}

int32_t ccp_peer_group_world(const ccp_peer_group* g) { return g ? g->world : CCP_ERR_INVALID; }

// contiguous split of [0, total) whose slice sizes differ by at most one
static void shard(int64_t total, int rank, int world, int64_t* first, int64_t* count) {
  const int64_t base = total / world, rem = total % world;
  *count = base + (rank < rem ? 1 : 0);
  *first = rank * base + (rank < rem ? rank : rem);
}

int ccp_peer_group_sample_project(ccp_peer_group* g, const ccp_sampler_args* a, int64_t total, int64_t* counts_host) {
  if (!g) return CCP_ERR_INVALID;
  if (!a || total < 0) return multi_err(g, CCP_ERR_INVALID, "peer group: bad sampler arguments");
  uint64_t pools[CCP_MAX_GROUP], cnts[CCP_MAX_GROUP];
  for (int r = 0; r < g->world; ++r) {
    pools[r] = (uint64_t)(uintptr_t)g->pool[r];
    cnts[r] = (uint64_t)(uintptr_t)g->counts[r];
  }
  // every device: zero its count, project its slice with the fused gather on, publish the count — all asynchronous
  for (int r = 0; r < g->world; ++r) {
    dev_scope s(g->dev[r]);
    int64_t first = 0, count = 0;
    shard(total, r, g->world, &first, &count);
    ccp_sampler_args ar = *a;
    ar.first_index = a->first_index + first;
    cudaError_t e = cudaMemsetAsync(g->n_ok[r], 0, sizeof(int64_t), g->stream[r]);
    if (e != cudaSuccess) return multi_err(g, CCP_ERR_CUDA, "peer group", cudaGetErrorString(e));
    int rc = ccp_set_gather_peers(g->h[r], g->world, r, pools, g->capacity);
    if (rc == CCP_OK)
      rc = ccp_sample_project_batch(g->h[r], &ar, count, CCP_LAYOUT_AOS, nullptr, nullptr, nullptr, nullptr, g->n_ok[r],
                                    g->stream[r]);
    if (rc == CCP_OK) rc = ccp_publish_count(g->h[r], g->n_ok[r], g->world, r, cnts, g->stream[r]);
    ccp_set_gather_peers(g->h[r], 0, 0, nullptr, 0);  // host-side state: the launches above carry their copy
    if (rc != CCP_OK) return multi_err(g, rc, "peer group", ccp_last_error(g->h[r]));
  }
  // completion of every device's stream = every row and count is in every pool
  for (int r = 0; r < g->world; ++r) {
    dev_scope s(g->dev[r]);
    cudaError_t e = cudaStreamSynchronize(g->stream[r]);
    if (e != cudaSuccess) return multi_err(g, CCP_ERR_CUDA, "peer group", cudaGetErrorString(e));
  }
  if (counts_host) {
    dev_scope s(g->dev[0]);
    cudaError_t e = cudaMemcpy(counts_host, g->counts[0], sizeof(int64_t) * g->world, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return multi_err(g, CCP_ERR_CUDA, "peer group", cudaGetErrorString(e));
    for (int r = 0; r < g->world; ++r)
      if (counts_host[r] > g->capacity) return multi_err(g, CCP_ERR_INVALID, "peer group: a rank converged more states than the capacity");
  }
  return CCP_OK;
}

int ccp_peer_group_pool(const ccp_peer_group* g, int32_t rank, const double** pool_dev, const int64_t** counts_dev,
                        int64_t* capacity) {
  if (!g || rank < 0 || rank >= g->world) return CCP_ERR_INVALID;
  if (pool_dev) *pool_dev = g->pool[rank];
  if (counts_dev) *counts_dev = g->counts[rank];
  if (capacity) *capacity = g->capacity;
  return CCP_OK;
}

int ccp_peer_group_gather_host(ccp_peer_group* g, int32_t rank, double* states_host, int64_t max_rows, int64_t* counts_host,
                               int64_t* rows_out) {
  if (!g || rank < 0 || rank >= g->world) return CCP_ERR_INVALID;
  if (!states_host || !counts_host) return multi_err(g, CCP_ERR_INVALID, "peer group: null host buffer");
  dev_scope s(g->dev[rank]);
  cudaError_t e = cudaMemcpy(counts_host, g->counts[rank], sizeof(int64_t) * g->world, cudaMemcpyDeviceToHost);
  int64_t rows = 0;
  for (int r = 0; r < g->world && e == cudaSuccess; ++r) {
    const int64_t k = counts_host[r];
    if (k > g->capacity) return multi_err(g, CCP_ERR_INVALID, "peer group: a rank converged more states than the capacity");
    if (rows + k > max_rows) return multi_err(g, CCP_ERR_INVALID, "peer group: host buffer too small");
    if (k > 0)
      e = cudaMemcpy(states_host + rows * g->n, g->pool[rank] + (size_t)r * g->capacity * g->n, sizeof(double) * g->n * (size_t)k,
                     cudaMemcpyDeviceToHost);
    rows += k;
  }
  if (e != cudaSuccess) return multi_err(g, CCP_ERR_CUDA, "peer group", cudaGetErrorString(e));
  if (rows_out) *rows_out = rows;
  return CCP_OK;
}

// ---- NCCL fallback: counts + padded compacted states, on a communicator the host owns --------------------------------
typedef int (*nccl_allgather_fn)(const void*, void*, size_t, int /*ncclDataType_t*/, void* /*ncclComm_t*/, cudaStream_t);
typedef const char* (*nccl_errstr_fn)(int);
static nccl_allgather_fn g_allgather = nullptr;
static nccl_errstr_fn g_errstr = nullptr;
static std::once_flag g_nccl_once;

static void resolve_nccl() {
  // the library the host process already carries (its communicator was made by it), else the system one
  void* fn = dlsym(RTLD_DEFAULT, "ncclAllGather");
  void* es = dlsym(RTLD_DEFAULT, "ncclGetErrorString");
  if (!fn) {
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (lib) {
      fn = dlsym(lib, "ncclAllGather");
      es = dlsym(lib, "ncclGetErrorString");
    }
  }
  g_allgather = (nccl_allgather_fn)fn;
  g_errstr = (nccl_errstr_fn)es;
}

int ccp_allgather_converged(ccp_handle* h, void* nccl_comm, int32_t world, const double* compact_dev, const int64_t* n_ok_dev,
                            int64_t capacity, double* pool_dev, int64_t* counts_dev, void* stream) {
  if (!h) return CCP_ERR_INVALID;
  if (!nccl_comm || !compact_dev || !n_ok_dev || !pool_dev || !counts_dev || capacity < 1 || world < 1) {
    multi_err(nullptr, CCP_ERR_INVALID, "allgather_converged: null argument / bad capacity");
    return CCP_ERR_INVALID;
  }
  std::call_once(g_nccl_once, resolve_nccl);
  if (!g_allgather) {
    multi_err(nullptr, CCP_ERR_NCCL, "allgather_converged: no NCCL in this process and libnccl.so.2 not loadable");
    return CCP_ERR_NCCL;
  }
  const int n = 7 * ccp_n_arms(h);
  const int ncclInt64 = 4, ncclFloat64 = 8;  // nccl.h: ncclInt64 = 4, ncclDouble = ncclFloat64 = 8
  dev_scope s(ccp_device(h));
  int rc = g_allgather(n_ok_dev, counts_dev, 1, ncclInt64, nccl_comm, (cudaStream_t)stream);
  if (rc == 0) rc = g_allgather(compact_dev, pool_dev, (size_t)capacity * n, ncclFloat64, nccl_comm, (cudaStream_t)stream);
  if (rc != 0) {
    multi_err(nullptr, CCP_ERR_NCCL, "allgather_converged", g_errstr ? g_errstr(rc) : "NCCL error");
    return CCP_ERR_NCCL;
  }
  return CCP_OK;
}

}  // extern "C"
