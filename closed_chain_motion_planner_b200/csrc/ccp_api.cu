// ccp_api.cu — the small batched kernels and the C ABI (include/ccp.h) of the batched closed-chain
// constraint-projection engine.
//
// Kernel design (DESIGN.md has the long form):
//   * one THREAD owns one sample: the matrices are 3-vectors and quaternions, nothing is a dense
//     contraction, so every lane does useful FP64 work and no shuffles sit on the critical path;
//   * the Newton loop is a persistent LANE-REFILL loop (ccp_project.cu): the moment a lane's sample converges
//     (or hits the 250-iteration cap) the lane writes its result and takes the next work number from its
//     warp's private chunk, so the 0..250-iteration spread does not idle the warp; the launch tail is packed
//     (complete launches) or parked for the next launch (pipelined launches);
//   * the model (DH constants, frames, tolerances; 2.2 KB) is a __grid_constant__ kernel parameter:
//     every use is a constant-bank operand, never a global load;
//   * all arithmetic is the shared header ccp_core.h (explicit fma, --fmad=false).
//
// Reference behaviour replaced: KinematicChainConstraint::{function,jacobian,project,isSatisfied,
// jointValid,setInitialPosition} (include/.../base/constraints/ConstraintFunction.h:31-120) and
// PandaModel FK/Jacobian (src/kinematics/panda_rbdl.cpp:9-42).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <new>
#include <type_traits>

#include "ccp.h"
#include "ccp_core.h"
#include "ccp_device.cuh"
#include "ccp_internal.h"
#include "ccp_flops.h"
#include "ccp_ik.h"
#include "ccp_pack.h"

#define CCP_VERSION_STRING "ccp-b200 0.1 (sm_100a)"

// ------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------
#define CCP_NUM_COUNTERS 64
#define CCP_MAX_AGE 40          /* < CCP_NUM_DESC: a launch's descriptor slot outlives every sample it parked */
#define CCP_HOST_LAG_DEFAULT 6
#define CCP_ZERO_COPY_MAX 512   /* host batches up to this many states run in place in page-locked host memory */
/* batch size up to which the two-lanes-per-sample kernel is the faster one: one cooperative warp (16 samples) per
 * scheduler, 3/4 full — measured crossover on B200 between 4 000 and 10 000 samples (tools/coop_probe.py) */
// Complete launches up to this many samples per SM take the cooperative kernel (measured crossovers on B200, DESIGN.md
// 4.1b): two arms 128 (two resident blocks x 4 warps x 16 samples), three arms 72 (8 samples per warp; a little beyond
// the 64 with a static place).  The geodesic walk switches much later (4.2): its thresholds are these times a ratio.
#define CCP_COOP_MAX_PER_SM 128
#define CCP_COOP3_MAX_PER_SM 72
#define CCP_GEO_COOP_PER_SM 448    // two arms: ~66 000 edges on B200
#define CCP_GEO_COOP3_PER_SM 512   // three arms: ~76 000 edges
#define CCP_HOST_MAX_CHUNKS 24  /* < CCP_NUM_DESC / 2: every chunk launch of a host call stays pipelined */

// per-launch device record (ring of CCP_NUM_COUNTERS): zeroed by ONE stream-ordered memset before the launch
struct ccp_launch_rec {
  unsigned long long work;  // dynamic work counter of the projection / geodesic kernel
  unsigned parked;          // samples the (pipelined) launch parked
  unsigned pad;
};

// One submitted host batch of the streaming host path (ccp_project_batch_host_submit / _wait)
// cuStreamWaitValue32(stream, device address, value, flags): flags 0 = wait until (int32)(*addr - value) >= 0
typedef int (*ccp_wait32_fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);

struct ccp_host_call {
  int64_t ticket;      // 0 = never used
  bool finished;       // results are complete in the caller's host buffers (or the slot is free)
  int64_t count, chunk;
  int parts, copied;    // chunks of the batch / chunks whose D2H copies are enqueued
  bool use_wait;        // copies wait on the launch slots' completion counts (pinned outputs, cuStreamWaitValue32)
  bool enqueued;        // use_wait: the waits and copies are on the D2H stream
  unsigned slot[CCP_HOST_MAX_CHUNKS], expect[CCP_HOST_MAX_CHUNKS];  // per chunk: launch slot, completion count to wait for
  int64_t first_launch; // number (ccp_handle::host_launches) of the launch that projected chunk 0
  double* x_out_host;
  uint8_t* ok_host;
  uint8_t* conv_host;
  int32_t* iters_host;
  double* resid_host;
  // compact outputs (ccp_host_batch::compact_host): the ok states of the batch packed by the kernels' epilogues; their
  // count is copied to `pin_nok` behind the last chunk, the rows themselves are copied by the wait (sized by the count)
  bool compact;
  double* compact_host;
  int32_t* cidx_host;
  int64_t ccap;          // rows the caller's buffers hold
  int64_t n_ok;          // valid after the wait
  double* dcompact;
  int32_t* dcidx;
  unsigned long long* dnok;
  long long* pin_nok;    // page-locked, one per slot
  void* stage;
  size_t stage_bytes;
  double* dx;
  double* dres;
  int32_t* dit;
  uint8_t* dok;
  uint8_t* dcv;
  cudaEvent_t done;
};

struct ccp_handle {
  int device;
  int sm_count;
  bool has_ref;
  ccp_model model;  // host image; passed BY VALUE to every kernel
  long long launches;
  ccp_launch_rec* d_counters;  // 2 * CCP_NUM_COUNTERS launch records (work counter + parked-sample counter): the first
                               // ring belongs to the projection launches (a parked sample's launch must keep its
                               // record and descriptor slot until it is adopted), the second to the geodesic launches
  unsigned launch_seq, geo_seq;
  // grow-only device staging for the *_host entry points
  void* d_stage;
  size_t d_stage_bytes;
  void* pin;               // page-locked host buffer of the zero-copy small-batch path
  cudaStream_t hstream[3];
  cudaEvent_t ev0, ev1;
  cudaEvent_t ev_chunk_in[CCP_HOST_MAX_CHUNKS], ev_chunk_k[CCP_HOST_MAX_CHUNKS];  // host path: chunk landed / projected
  // pipelined projection launches: two park buffers (the launch adopts from one and parks into the other),
  // their record counters, and the table of output descriptors by launch slot
  ccp_park_rec* d_park[2];
  unsigned prev_slot;      // launch slot whose record counts the parked samples in d_park[park_cur]
  ccp_out_desc* d_desc;    // [CCP_NUM_DESC]
  unsigned* d_done;        // [CCP_NUM_DESC] samples finished per launch slot (cumulative; ccp_project_args::done)
  unsigned done_total[CCP_NUM_DESC];  // what d_done[slot] reads once every launch that used the slot is complete
  unsigned last_slot;      // descriptor slot of the most recent projection launch
  ccp_wait32_fn wait32;    // cuStreamWaitValue32, or nullptr (then the host path copies a chunk `lag` launches late)
  size_t park_capacity;    // records per buffer
  int park_cur;            // buffer holding the parked samples of the last pipelined launch
  bool pipeline_open;      // parked samples exist
  int pipe_sig;            // layout | gen << 1 of the launches in the pipeline (same kernel instantiation)
  int pipe_launches;       // pipelined launches since the pipeline opened (slot ring safety)
  ccp_host_call hcall[2];  // streaming host path: two batches in flight, each with its own device stage
  int64_t next_ticket;
  int64_t host_launches;   // chunk launches of the host path so far
  cudaMemPool_t pool;  // stream-ordered scratch of the sampler path; keeps its memory between calls
  // fused all-gather target (ccp_set_gather_peers)
  double* peer_pool[CCP_MAX_PEERS];
  double* peer_mc;
  int peer_world, peer_rank;
  long long peer_cap;
  long long coop_max;  // complete launches of at most this many samples take the cooperative kernel
  std::mutex mu;
  std::mutex host_mu;  // the *_host entry points share the stage buffer and the private streams: one at a time
  char err[512];
};

static thread_local char g_create_err[512] = "";

static int set_err(ccp_handle* h, int code, const char* fmt, const char* a = "", const char* b = "") {
  char* dst = h ? h->err : g_create_err;
  snprintf(dst, 512, fmt, a, b);
  return code;
}
#define CCP_CUDA(call)                                                                             \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return set_err(h, CCP_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

// ------------------------------------------------------------------------------------------
// small batched kernels (thread per state, grid-stride)
// ------------------------------------------------------------------------------------------
template <int K, int PANDA, bool SOA>
__global__ void __launch_bounds__(128)
ccp_function_kernel(const __grid_constant__ ccp_model M, const double* __restrict__ xin, long long count,
                    double* __restrict__ f, uint8_t* __restrict__ satisfied) {
  constexpr int n = CCPC_DOF * K, m = 2 * (K - 1);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < count;
       idx += (long long)gridDim.x * blockDim.x) {
    double x[n];
#pragma unroll
    for (int j = 0; j < n; ++j) x[j] = ld_elem<SOA>(xin, idx, j, count, n);
    ccp_fwd<K> F;
    ccp_sc_local<K> S;
    ccp_forward<K, PANDA>(M, x, S, F);
    double fv[m];
    ccp_residual<K>(F, fv, nullptr);
    if (f) {
#pragma unroll
      for (int k = 0; k < m; ++k) st_elem<SOA>(f, idx, k, count, m, fv[k]);
    }
    if (satisfied) satisfied[idx] = ccp_is_satisfied<K>(M, fv);
  }
}

template <int K, int PANDA, bool SOA>
__global__ void __launch_bounds__(128)
ccp_jacobian_kernel(const __grid_constant__ ccp_model M, const double* __restrict__ xin, long long count,
                    double* __restrict__ Jout) {
  constexpr int n = CCPC_DOF * K, m = 2 * (K - 1);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < count;
       idx += (long long)gridDim.x * blockDim.x) {
    double x[n];
#pragma unroll
    for (int j = 0; j < n; ++j) x[j] = ld_elem<SOA>(xin, idx, j, count, n);
    ccp_fwd<K> F;
    ccp_jac<K> J;
    ccp_sc_local<K> S;
    ccp_forward<K, PANDA>(M, x, S, F);
    ccp_jacobian<K, PANDA>(M, S, F, J);
    double D[m * n];
    ccp_jac_dense<K>(F, J, D);
#pragma unroll
    for (int k = 0; k < m * n; ++k) st_elem<SOA>(Jout, idx, k, count, m * n, D[k]);
  }
}

template <int K, bool SOA>
__global__ void __launch_bounds__(128)
ccp_joint_valid_kernel(const __grid_constant__ ccp_model M, const double* __restrict__ xin, long long count,
                       uint8_t* __restrict__ out) {
  constexpr int n = CCPC_DOF * K;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < count;
       idx += (long long)gridDim.x * blockDim.x) {
    double x[n];
#pragma unroll
    for (int j = 0; j < n; ++j) x[j] = ld_elem<SOA>(xin, idx, j, count, n);
    out[idx] = ccp_joint_valid<K>(M, x);
  }
}

template <bool SOA>
__global__ void __launch_bounds__(128)
ccp_arm_fk_kernel(const __grid_constant__ ccp_model M, int arm, const double* __restrict__ qin, long long count,
                  double* __restrict__ T, double* __restrict__ Jac) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < count;
       idx += (long long)gridDim.x * blockDim.x) {
    double q[CCPC_DOF];
#pragma unroll
    for (int j = 0; j < CCPC_DOF; ++j) q[j] = ld_elem<SOA>(qin, idx, j, count, CCPC_DOF);
    double Tl[12], Jl[42];
    // `arm` is uniform; select the arm with a switch so the model stays in the constant bank
    switch (arm) {
      case 0: ccp_arm_fk(M.arm[0], q, Tl, Jl); break;
      case 1: ccp_arm_fk(M.arm[1], q, Tl, Jl); break;
      default: ccp_arm_fk(M.arm[2], q, Tl, Jl); break;
    }
    if (T) {
#pragma unroll
      for (int k = 0; k < 12; ++k) st_elem<SOA>(T, idx, k, count, 12, Tl[k]);
    }
    if (Jac) {
#pragma unroll
      for (int k = 0; k < 42; ++k) st_elem<SOA>(Jac, idx, k, count, 42, Jl[k]);
    }
  }
}

template <int K, bool SOA>
__global__ void __launch_bounds__(128)
ccp_seed_kernel(const __grid_constant__ ccp_model M, const __grid_constant__ ccp_project_args A,
                double* __restrict__ out) {
  // one thread per ELEMENT (sample, joint): every element of the counter-based stream is an independent hash of
  // (seed, sample, joint), and element order = memory order in both layouts, so the stores are fully coalesced (a
  // thread per sample wrote 14 doubles at a 112-byte stride: 0.76 TB/s, 6 % of a Wine_Bottle pool refill)
  constexpr int n = CCPC_DOF * K;
  const long long total = A.count * n;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long idx = SOA ? e % A.count : e / n;
    const int j = (int)(SOA ? e / A.count : e % n);
    out[e] = make_seed<K>(M, A, idx, j);
  }
}

__global__ void __launch_bounds__(256) ccp_wrap_kernel(double* __restrict__ x, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x)
    x[i] = ccp_wrap_pi(x[i]);
}

template <int K, int PANDA>
__global__ void ccp_reference_kernel(const __grid_constant__ ccp_model M, const double* __restrict__ q_start,
                                     ccp_pair_ref* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  ccp_model L = M;
  double q[CCPC_DOF * K];
  for (int j = 0; j < CCPC_DOF * K; ++j) q[j] = q_start[j];
  ccp_reference_chain<K, PANDA>(L, q);
  for (int p = 0; p < K - 1; ++p) out[p] = L.ref[p];
}

// Register-only DFMA chains: 8 independent accumulators per thread, `inner` x 8 x 2 FLOP per thread.
__global__ void __launch_bounds__(256) ccp_dfma_probe_kernel(double* __restrict__ sink, int inner, double a, double b) {
  double r0 = threadIdx.x * 1e-3, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6,
         r7 = r0 + 7;
#pragma unroll 1
  for (int i = 0; i < inner; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
      r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
    }
  }
  double s = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
  if (s == 123.456) sink[0] = s;  // never true; keeps the chain alive
}

// one warp: this rank's converged count -> slot `rank` of every rank's count array (peer-mapped memory)
struct ccp_count_peers {
  long long* counts[CCP_MAX_PEERS];
};
__global__ void ccp_publish_count_kernel(const long long* __restrict__ n_ok, const __grid_constant__ ccp_count_peers P,
                                         int world, int rank) {
  const long long v = *n_ok;
  if ((int)threadIdx.x < world) {
    P.counts[threadIdx.x][rank] = v;
    __threadfence_system();
  }
}

// dispatch helper: (arms, link-code mode 0 generic / 1 structured alpha / 2 stock) -> template arguments
#define CCP_DISPATCH_KP(h, CALL)                                                  \
  do {                                                                            \
    if ((h)->model.n_arms == 2) {                                                 \
      if ((h)->model.stock) { CALL(2, 2); } else if ((h)->model.panda_alpha) { CALL(2, 1); } else { CALL(2, 0); } \
    } else {                                                                      \
      if ((h)->model.stock) { CALL(3, 2); } else if ((h)->model.panda_alpha) { CALL(3, 1); } else { CALL(3, 0); } \
    }                                                                             \
  } while (0)

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static inline int grid_for(const ccp_handle* h, long long count, int block, int per_sm) {
  long long need = (count + block - 1) / block;
  long long cap = (long long)h->sm_count * per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

struct device_guard {
  int prev;
  bool ok;
  explicit device_guard(int dev) : prev(-1), ok(false) {
    if (cudaGetDevice(&prev) != cudaSuccess) return;
    ok = (prev == dev) || (cudaSetDevice(dev) == cudaSuccess);
  }
  ~device_guard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

// Pipelined launches need the park buffers: every lane can park its live sample plus (its share of) the unstarted
// rest of its warp's private chunk, so 2 records per resident thread bound a launch's parking.
static int ensure_pipeline(ccp_handle* h) {
  if (h->d_park[0]) return CCP_OK;
  const size_t cap = (size_t)h->sm_count * 512 * 2;
  CCP_CUDA(cudaMalloc(&h->d_park[0], cap * sizeof(ccp_park_rec)));
  CCP_CUDA(cudaMalloc(&h->d_park[1], cap * sizeof(ccp_park_rec)));
  CCP_CUDA(cudaMalloc(&h->d_desc, CCP_NUM_DESC * sizeof(ccp_out_desc)));
  CCP_CUDA(cudaMemset(h->d_desc, 0, CCP_NUM_DESC * sizeof(ccp_out_desc)));
  CCP_CUDA(cudaMalloc(&h->d_done, CCP_NUM_DESC * sizeof(unsigned)));
  CCP_CUDA(cudaMemset(h->d_done, 0, CCP_NUM_DESC * sizeof(unsigned)));
  h->park_capacity = cap;
  return CCP_OK;
}

static int dispatch_project(ccp_handle* h, ccp_project_args& A, bool soa, cudaStream_t st) {
  cudaError_t e = cudaSuccess;
  const int mode = h->model.stock ? 2 : (h->model.panda_alpha ? 1 : 0);
  if (h->model.n_arms == 2)
    e = mode == 2 ? ccp_launch_project_K2_P2(h->sm_count, h->model, A, soa, st)
        : mode == 1 ? ccp_launch_project_K2_P1(h->sm_count, h->model, A, soa, st)
                    : ccp_launch_project_K2_P0(h->sm_count, h->model, A, soa, st);
  else
    e = mode == 2 ? ccp_launch_project_K3_P2(h->sm_count, h->model, A, soa, st)
        : mode == 1 ? ccp_launch_project_K3_P1(h->sm_count, h->model, A, soa, st)
                    : ccp_launch_project_K3_P0(h->sm_count, h->model, A, soa, st);
  if (e != cudaSuccess) return set_err(h, CCP_ERR_CUDA, "project kernel launch: %s", cudaGetErrorString(e));
  return CCP_OK;
}

// One projection launch.  defer = true: pipelined (the samples still iterating when the work runs dry are parked
// for the next launch).  A launch that finds the pipeline open adopts the parked samples first; a launch that is
// not itself deferred completes them and thereby closes the pipeline.
static int launch_project(ccp_handle* h, ccp_project_args& A, int layout, cudaStream_t st, bool defer = false) {
  if (!h->has_ref) return set_err(h, CCP_ERR_STATE, "%s", "project before ccp_set_reference (setInitialPosition)");
  if (A.count > 0x7fff0000LL)
    return set_err(h, CCP_ERR_INVALID, "%s", "more than 2^31 - 65536 samples in one call: split the batch");
  const bool soa = layout == CCP_LAYOUT_SOA;
  const int sig = soa ? 1 : 0;
  A.stage_seeds = (A.stage_seeds >= 0 && A.seeds && (((uintptr_t)A.seeds) & 15u) == 0) ? 1 : 0;
  if (A.seed_stride == 0) A.seed_stride = A.count;
  if (A.out_stride == 0) A.out_stride = A.count;
  if (h->pipeline_open && sig != h->pipe_sig) {
    // parked samples belong to the other kernel instantiation (layout): complete them first
    ccp_project_args F;
    memset(&F, 0, sizeof F);
    int rc = launch_project(h, F, (h->pipe_sig & 1) ? CCP_LAYOUT_SOA : CCP_LAYOUT_AOS, st, false);
    if (rc) return rc;
  }
  if (A.count == 0 && !h->pipeline_open) return CCP_OK;
  unsigned slot;
  {
    std::lock_guard<std::mutex> lk(h->mu);
    slot = h->launch_seq++ % CCP_NUM_COUNTERS;
    h->launches++;
  }
  A.counter = &h->d_counters[slot].work;
  A.slot = slot % CCP_NUM_DESC;
  h->last_slot = A.slot;
  if (A.max_age == 0 || A.max_age > CCP_MAX_AGE) A.max_age = CCP_MAX_AGE;
  if (h->peer_world > 0 && A.n_ok) {
    A.peer_world = h->peer_world;
    A.peer_row0 = (long long)h->peer_rank * h->peer_cap;
    A.peer_cap = h->peer_cap;
    A.peer_mc = h->peer_mc;
    for (int p = 0; p < h->peer_world; ++p) A.peer_pool[p] = h->peer_pool[p];
  }
  // a small complete launch: latency, not throughput, is what it costs — two lanes per sample (ccp_coop.cu).  When every
  // sample has its static place in the grid (always, below the default threshold) the work counter is never touched and
  // its stream-ordered zeroing — one more operation on the path of a single project() call — is skipped.
  const long long spw = h->model.n_arms == 2 ? 16 : 8;  // samples per warp of the cooperative kernels
  const bool coop = !defer && !h->pipeline_open && A.count <= h->coop_max &&
                    A.peer_world == 0 && !A.own_n_ok;
  const bool coop_static = coop && A.count <= 4 * spw * (long long)ccp_coop_grid(h->sm_count, A.count, (int)spw);
  if (coop_static) A.counter = nullptr;
  else CCP_CUDA(cudaMemsetAsync(h->d_counters + slot, 0, sizeof(ccp_launch_rec), st));
  if (defer || h->pipeline_open) {
    int rc = ensure_pipeline(h);
    if (rc) return rc;
    A.desc_table = h->d_desc;
    A.done = h->d_done;
    h->done_total[A.slot] += (unsigned)A.count;
    if (h->pipeline_open) {
      A.adopt = h->d_park[h->park_cur];
      A.adopt_count = &h->d_counters[h->prev_slot].parked;
    }
    if (defer) {
      A.park = h->d_park[1 - h->park_cur];
      A.park_count = &h->d_counters[slot].parked;
    }
  }
  int rc;
  if (coop) {
    cudaError_t e = ccp_launch_project_coop(h->sm_count, h->model, A, soa, st);
    rc = (e == cudaSuccess) ? CCP_OK : set_err(h, CCP_ERR_CUDA, "cooperative project kernel launch: %s", cudaGetErrorString(e));
  } else {
    rc = dispatch_project(h, A, soa, st);
  }
  if (rc) return rc;
  if (defer) {
    h->pipeline_open = true;
    h->pipe_sig = sig;
    h->pipe_launches++;
    h->park_cur = 1 - h->park_cur;
    h->prev_slot = slot;
  } else {
    h->pipeline_open = false;
    h->pipe_launches = 0;
  }
  return CCP_OK;
}

static int check_common(ccp_handle* h, const void* p, int64_t count, int32_t layout) {
  if (!h) return CCP_ERR_INVALID;
  if (count < 0) return set_err(h, CCP_ERR_INVALID, "%s", "negative count");
  if (count > 0 && !p) return set_err(h, CCP_ERR_INVALID, "%s", "null state pointer");
  if (layout != CCP_LAYOUT_AOS && layout != CCP_LAYOUT_SOA) return set_err(h, CCP_ERR_INVALID, "%s", "bad layout");
  return CCP_OK;
}

// page-locked host buffer of the copy-free small-call paths (the kernels read and write it in place over PCIe)
#define CCP_PIN_BYTES ((size_t)CCP_ZERO_COPY_MAX * (sizeof(double) * (CCPC_DOF * CCPC_MAX_ARMS + 2 * CCPC_MAX_ARMS) + 16) + 256)
static int ensure_pin(ccp_handle* h) {
  if (h->pin) return CCP_OK;
  CCP_CUDA(cudaHostAlloc(&h->pin, CCP_PIN_BYTES, cudaHostAllocDefault));
  return CCP_OK;
}

static int ensure_stage(ccp_handle* h, size_t bytes) {
  if (bytes <= h->d_stage_bytes) return CCP_OK;
  if (h->d_stage) cudaFree(h->d_stage);
  h->d_stage = nullptr;
  h->d_stage_bytes = 0;
  size_t want = bytes + bytes / 4;
  CCP_CUDA(cudaMalloc(&h->d_stage, want));
  h->d_stage_bytes = want;
  return CCP_OK;
}

extern "C" {

const char* ccp_version(void) { return CCP_VERSION_STRING; }

int ccp_default_model(int32_t n_arms, const int32_t* arm_index, ccp_model_desc* d) {
  return ccp_fill_default_model(n_arms, arm_index, d) == 0 ? CCP_OK : CCP_ERR_INVALID;
}

int ccp_create(const ccp_model_desc* model, int32_t device, ccp_handle** out) {
  ccp_handle* h = nullptr;
  if (!model || !out) return set_err(nullptr, CCP_ERR_INVALID, "%s", "null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return set_err(nullptr, CCP_ERR_CUDA, "%s", "no CUDA device: this engine has no CPU fallback");
  if (device < 0 || device >= ndev) return set_err(nullptr, CCP_ERR_INVALID, "%s", "device index out of range");
  ccp_model M;
  if (ccp_pack_model(model, &M) != 0) return set_err(nullptr, CCP_ERR_INVALID, "%s", "invalid model (n_arms must be 2 or 3)");
  ccp_handle* nh = new (std::nothrow) ccp_handle();
  if (!nh) return set_err(nullptr, CCP_ERR_INVALID, "%s", "out of memory");
  nh->device = device;
  nh->model = M;
  nh->has_ref = false;
  nh->launches = 0;
  nh->launch_seq = 0;
  nh->geo_seq = 0;
  nh->d_stage = nullptr;
  nh->pin = nullptr;
  nh->d_stage_bytes = 0;
  nh->d_park[0] = nh->d_park[1] = nullptr;
  nh->prev_slot = 0;
  {
    const char* e = getenv("CCP_COOP_MAX");
    nh->coop_max = e ? atoll(e) : -1;  // -1: default, resolved once the SM count is known
  }
  nh->peer_world = 0;
  nh->peer_mc = nullptr;
  nh->peer_rank = 0;
  nh->peer_cap = 0;
  nh->d_desc = nullptr;
  nh->d_done = nullptr;
  memset(nh->done_total, 0, sizeof nh->done_total);
  nh->last_slot = 0;
  nh->wait32 = nullptr;
  {
    const char* off = getenv("CCP_HOST_WAITVALUE");
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (!(off && atoi(off) == 0) &&
        cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      nh->wait32 = (ccp_wait32_fn)fn;
    cudaGetLastError();
  }
  nh->park_capacity = 0;
  nh->park_cur = 0;
  nh->pipeline_open = false;
  nh->pipe_sig = 0;
  nh->pipe_launches = 0;
  nh->err[0] = 0;
  device_guard g(device);
  cudaError_t e = g.ok ? cudaSuccess : cudaErrorInvalidDevice;
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&nh->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (e == cudaSuccess && nh->coop_max < 0)
    nh->coop_max = (long long)nh->sm_count * (model->n_arms == 2 ? CCP_COOP_MAX_PER_SM : CCP_COOP3_MAX_PER_SM);
  if (e == cudaSuccess) e = cudaMalloc(&nh->d_counters, 2 * CCP_NUM_COUNTERS * sizeof(ccp_launch_rec));
  for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&nh->hstream[i], cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&nh->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&nh->ev1);
  nh->next_ticket = 1;
  nh->host_launches = 0;
  for (int i = 0; i < 2; ++i) {
    memset(&nh->hcall[i], 0, sizeof nh->hcall[i]);
    nh->hcall[i].finished = true;
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&nh->hcall[i].done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&nh->hcall[i].pin_nok, 64, cudaHostAllocDefault);
  }
  nh->pool = nullptr;
  if (e == cudaSuccess) {
    cudaMemPoolProps props;
    memset(&props, 0, sizeof props);
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    e = cudaMemPoolCreate(&nh->pool, &props);
    if (e == cudaSuccess) {
      unsigned long long keep = ~0ULL;  // never trim at synchronisation points: the scratch is reused by every call
      e = cudaMemPoolSetAttribute(nh->pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  for (int i = 0; i < CCP_HOST_MAX_CHUNKS && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&nh->ev_chunk_in[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&nh->ev_chunk_k[i], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    set_err(nullptr, CCP_ERR_CUDA, "ccp_create: %s", cudaGetErrorString(e));
    delete nh;
    return CCP_ERR_CUDA;
  }
  *out = nh;
  return CCP_OK;
}

void ccp_destroy(ccp_handle* h) {
  if (!h) return;
  device_guard g(h->device);
  cudaDeviceSynchronize();
  if (h->d_counters) cudaFree(h->d_counters);
  if (h->d_stage) cudaFree(h->d_stage);
  if (h->pin) cudaFreeHost(h->pin);
  for (int i = 0; i < 2; ++i) {
    if (h->hcall[i].stage) cudaFreeAsync(h->hcall[i].stage, h->hstream[0]);
    if (h->hcall[i].done) cudaEventDestroy(h->hcall[i].done);
    if (h->hcall[i].pin_nok) cudaFreeHost(h->hcall[i].pin_nok);
  }
  if (h->pool) cudaMemPoolDestroy(h->pool);
  if (h->d_park[0]) cudaFree(h->d_park[0]);
  if (h->d_park[1]) cudaFree(h->d_park[1]);
  if (h->d_desc) cudaFree(h->d_desc);
  if (h->d_done) cudaFree(h->d_done);
  for (int i = 0; i < CCP_HOST_MAX_CHUNKS; ++i) {
    cudaEventDestroy(h->ev_chunk_in[i]);
    cudaEventDestroy(h->ev_chunk_k[i]);
  }
  for (int i = 0; i < 3; ++i) cudaStreamDestroy(h->hstream[i]);
  cudaEventDestroy(h->ev0);
  cudaEventDestroy(h->ev1);
  delete h;
}

const char* ccp_last_error(const ccp_handle* h) { return h ? h->err : g_create_err; }
int ccp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
int ccp_n_arms(const ccp_handle* h) { return h ? h->model.n_arms : CCP_ERR_INVALID; }
int ccp_device(const ccp_handle* h) { return h ? h->device : CCP_ERR_INVALID; }
int64_t ccp_launch_count(const ccp_handle* h) { return h ? h->launches : 0; }

#define CCP_NO_OPEN_PIPELINE(h)                                                                                   \
  do {                                                                                                             \
    if ((h)->pipeline_open)                                                                                        \
      return set_err((h), CCP_ERR_STATE, "%s", "pipelined projections are in flight: call ccp_project_flush first"); \
  } while (0)


// Device-pointer projection calls share the handle's launch pipeline with the streaming host path, whose launches run on
// the handle's private (non-blocking) streams: a device launch on the caller's stream would adopt the host batch's
// parked samples with nothing ordering it after them.  So they are refused while a host ticket is unfinished
// (ccp.h: "do not issue other projection calls on the handle while a ticket is pending").
#define CCP_NO_PENDING_TICKETS(h)                                                                                  \
  do {                                                                                                              \
    if (!(h)->hcall[0].finished || !(h)->hcall[1].finished)                                                         \
      return set_err((h), CCP_ERR_STATE, "%s",                                                                      \
                     "a host batch is pending (ccp_project_batch_host_submit): wait for its ticket first");         \
  } while (0)

int ccp_set_reference(ccp_handle* h, const double* q_start_host) {
  if (!h || !q_start_host) return h ? set_err(h, CCP_ERR_INVALID, "%s", "null q_start") : CCP_ERR_INVALID;
  CCP_NO_OPEN_PIPELINE(h);
  device_guard g(h->device);
  const int n = CCPC_DOF * h->model.n_arms;
  double* dq = nullptr;
  ccp_pair_ref* dref = nullptr;
  CCP_CUDA(cudaMalloc(&dq, sizeof(double) * n));
  cudaError_t e = cudaMalloc(&dref, sizeof(ccp_pair_ref) * (CCPC_MAX_ARMS - 1));
  if (e == cudaSuccess) e = cudaMemcpy(dq, q_start_host, sizeof(double) * n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
#define CCP_CALL(K, P) ccp_reference_kernel<K, P><<<1, 32>>>(h->model, dq, dref)
    CCP_DISPATCH_KP(h, CCP_CALL);
#undef CCP_CALL
    e = cudaGetLastError();
    h->launches++;
  }
  ccp_pair_ref ref[CCPC_MAX_ARMS - 1];
  if (e == cudaSuccess)
    e = cudaMemcpy(ref, dref, sizeof(ccp_pair_ref) * (h->model.n_arms - 1), cudaMemcpyDeviceToHost);
  cudaFree(dq);
  if (dref) cudaFree(dref);
  if (e != cudaSuccess) return set_err(h, CCP_ERR_CUDA, "ccp_set_reference: %s", cudaGetErrorString(e));
  for (int p = 0; p < h->model.n_arms - 1; ++p) h->model.ref[p] = ref[p];
  h->has_ref = true;
  return CCP_OK;
}

int ccp_get_reference(const ccp_handle* h, int32_t pair, double* t0_host, double* q0_host) {
  if (!h || pair < 0 || pair >= h->model.n_arms - 1 || !h->has_ref) return CCP_ERR_INVALID;
  if (t0_host) memcpy(t0_host, h->model.ref[pair].t0, 3 * sizeof(double));
  if (q0_host) memcpy(q0_host, h->model.ref[pair].q0, 4 * sizeof(double));
  return CCP_OK;
}

int ccp_set_tolerance(ccp_handle* h, double tol_position, double tol_rotation) {
  if (!h) return CCP_ERR_INVALID;
  if (!(tol_position > 0) || !(tol_rotation > 0))
    return set_err(h, CCP_ERR_INVALID, "%s", "setTolerance: tolerance must be positive.");
  CCP_NO_OPEN_PIPELINE(h);
  ccp_model_set_tolerance(&h->model, tol_position, tol_rotation);
  return CCP_OK;
}

int ccp_set_options(ccp_handle* h, const ccp_options* opt) {
  if (!h || !opt) return CCP_ERR_INVALID;
  if (!(opt->step > 0) || opt->max_iter < 0 || opt->max_iter > 65535 || !(opt->joint_margin >= 0))
    return set_err(h, CCP_ERR_INVALID, "%s", "bad options (step > 0, 0 <= max_iter <= 65535, joint_margin >= 0)");
  if (!(opt->damping >= 0) || (opt->clamp != 0 && opt->clamp != 1))
    return set_err(h, CCP_ERR_INVALID, "%s", "bad options (damping >= 0, clamp 0 or 1)");
  CCP_NO_OPEN_PIPELINE(h);
  h->model.step = opt->step;
  h->model.max_iter = opt->max_iter;
  ccp_model_set_margin(&h->model, opt->joint_margin);
  h->model.damping = opt->damping;
  h->model.clamp = opt->clamp;
  return CCP_OK;
}

int ccp_set_coop_threshold(ccp_handle* h, int64_t max_count) {
  if (!h) return CCP_ERR_INVALID;
  h->coop_max = max_count < 0 ? (long long)h->sm_count * (h->model.n_arms == 2 ? CCP_COOP_MAX_PER_SM : CCP_COOP3_MAX_PER_SM) : max_count;
  return CCP_OK;
}

int ccp_get_options(const ccp_handle* h, ccp_options* opt, double* tol_position, double* tol_rotation) {
  if (!h) return CCP_ERR_INVALID;
  if (opt) {
    opt->step = h->model.step;
    opt->max_iter = h->model.max_iter;
    opt->clamp = h->model.clamp;
    opt->joint_margin = h->model.margin;
    opt->damping = h->model.damping;
  }
  if (tol_position) *tol_position = h->model.tol_p;
  if (tol_rotation) *tol_rotation = h->model.tol_r;
  return CCP_OK;
}

int ccp_algorithmic_flops(const ccp_handle* h, double* per_iteration, double* per_tail) {
  if (!h) return CCP_ERR_INVALID;
  const bool k2 = h->model.n_arms == 2;
  if (per_iteration) *per_iteration = k2 ? CCP_FLOPS_ITER_K2 : CCP_FLOPS_ITER_K3;
  if (per_tail) *per_tail = k2 ? CCP_FLOPS_TAIL_K2 : CCP_FLOPS_TAIL_K3;
  return CCP_OK;
}

int ccp_algorithmic_flops_ik(double* per_iteration, double* per_tail) {
  if (per_iteration) *per_iteration = CCP_FLOPS_IK_ITER;
  if (per_tail) *per_tail = CCP_FLOPS_IK_TAIL;
  return CCP_OK;
}

static int function_impl(ccp_handle* h, const double* x_dev, int64_t count, int32_t layout, double* f_dev,
                         uint8_t* sat_dev, void* stream) {
  int rc = check_common(h, x_dev, count, layout);
  if (rc) return rc;
  if (!h->has_ref) return set_err(h, CCP_ERR_STATE, "%s", "function before ccp_set_reference (setInitialPosition)");
  if (count == 0) return CCP_OK;
  device_guard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(h, count, 128, 8);
  const bool soa = layout == CCP_LAYOUT_SOA;
#define CCP_CALL(K, P)                                                                                   \
  do {                                                                                                   \
    if (soa) ccp_function_kernel<K, P, true><<<grid, 128, 0, st>>>(h->model, x_dev, count, f_dev, sat_dev);   \
    else ccp_function_kernel<K, P, false><<<grid, 128, 0, st>>>(h->model, x_dev, count, f_dev, sat_dev);      \
  } while (0)
  CCP_DISPATCH_KP(h, CCP_CALL);
#undef CCP_CALL
  h->launches++;
  CCP_CUDA(cudaGetLastError());
  return CCP_OK;
}

int ccp_function_batch(ccp_handle* h, const double* x_dev, int64_t count, int32_t layout, double* f_dev,
                       void* stream) {
  if (h && count > 0 && !f_dev) return set_err(h, CCP_ERR_INVALID, "%s", "null output");
  return function_impl(h, x_dev, count, layout, f_dev, nullptr, stream);
}

int ccp_is_satisfied_batch(ccp_handle* h, const double* x_dev, int64_t count, int32_t layout, uint8_t* out_dev,
                           void* stream) {
  if (h && count > 0 && !out_dev) return set_err(h, CCP_ERR_INVALID, "%s", "null output");
  return function_impl(h, x_dev, count, layout, nullptr, out_dev, stream);
}

int ccp_jacobian_batch(ccp_handle* h, const double* x_dev, int64_t count, int32_t layout, double* J_dev,
                       void* stream) {
  int rc = check_common(h, x_dev, count, layout);
  if (rc) return rc;
  if (count > 0 && !J_dev) return set_err(h, CCP_ERR_INVALID, "%s", "null output");
  if (!h->has_ref) return set_err(h, CCP_ERR_STATE, "%s", "jacobian before ccp_set_reference (setInitialPosition)");
  if (count == 0) return CCP_OK;
  device_guard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(h, count, 128, 4);
  const bool soa = layout == CCP_LAYOUT_SOA;
#define CCP_CALL(K, P)                                                                          \
  do {                                                                                          \
    if (soa) ccp_jacobian_kernel<K, P, true><<<grid, 128, 0, st>>>(h->model, x_dev, count, J_dev);   \
    else ccp_jacobian_kernel<K, P, false><<<grid, 128, 0, st>>>(h->model, x_dev, count, J_dev);      \
  } while (0)
  CCP_DISPATCH_KP(h, CCP_CALL);
#undef CCP_CALL
  h->launches++;
  CCP_CUDA(cudaGetLastError());
  return CCP_OK;
}

int ccp_joint_valid_batch(ccp_handle* h, const double* x_dev, int64_t count, int32_t layout, uint8_t* out_dev,
                          void* stream) {
  int rc = check_common(h, x_dev, count, layout);
  if (rc) return rc;
  if (count > 0 && !out_dev) return set_err(h, CCP_ERR_INVALID, "%s", "null output");
  if (count == 0) return CCP_OK;
  device_guard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(h, count, 128, 8);
  const bool soa = layout == CCP_LAYOUT_SOA;
  if (h->model.n_arms == 2) {
    if (soa) ccp_joint_valid_kernel<2, true><<<grid, 128, 0, st>>>(h->model, x_dev, count, out_dev);
    else ccp_joint_valid_kernel<2, false><<<grid, 128, 0, st>>>(h->model, x_dev, count, out_dev);
  } else {
    if (soa) ccp_joint_valid_kernel<3, true><<<grid, 128, 0, st>>>(h->model, x_dev, count, out_dev);
    else ccp_joint_valid_kernel<3, false><<<grid, 128, 0, st>>>(h->model, x_dev, count, out_dev);
  }
  h->launches++;
  CCP_CUDA(cudaGetLastError());
  return CCP_OK;
}

static int arm_fk_impl(ccp_handle* h, int32_t arm, const double* q_dev, int64_t count, int32_t layout, double* T_dev,
                       double* J_dev, void* stream) {
  int rc = check_common(h, q_dev, count, layout);
  if (rc) return rc;
  if (arm < 0 || arm >= h->model.n_arms) return set_err(h, CCP_ERR_INVALID, "%s", "arm index out of range");
  if (count == 0) return CCP_OK;
  device_guard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(h, count, 128, 4);
  if (layout == CCP_LAYOUT_SOA) ccp_arm_fk_kernel<true><<<grid, 128, 0, st>>>(h->model, arm, q_dev, count, T_dev, J_dev);
  else ccp_arm_fk_kernel<false><<<grid, 128, 0, st>>>(h->model, arm, q_dev, count, T_dev, J_dev);
  h->launches++;
  CCP_CUDA(cudaGetLastError());
  return CCP_OK;
}

int ccp_fk_batch(ccp_handle* h, int32_t arm, const double* q_dev, int64_t count, int32_t layout, double* T_dev,
                 void* stream) {
  if (h && count > 0 && !T_dev) return set_err(h, CCP_ERR_INVALID, "%s", "null output");
  return arm_fk_impl(h, arm, q_dev, count, layout, T_dev, nullptr, stream);
}

int ccp_arm_jacobian_batch(ccp_handle* h, int32_t arm, const double* q_dev, int64_t count, int32_t layout,
                           double* J_dev, void* stream) {
  if (h && count > 0 && !J_dev) return set_err(h, CCP_ERR_INVALID, "%s", "null output");
  return arm_fk_impl(h, arm, q_dev, count, layout, nullptr, J_dev, stream);
}

int ccp_project_batch(ccp_handle* h, const double* seeds_dev, int64_t count, int32_t layout, double* x_out_dev,
                      uint8_t* ok_dev, uint8_t* converged_dev, int32_t* iters_dev, double* resid_dev,
                      double* compact_dev, int64_t* n_ok_dev, void* stream) {
  int rc = check_common(h, seeds_dev, count, layout);
  if (rc) return rc;
  if (compact_dev && !n_ok_dev) return set_err(h, CCP_ERR_INVALID, "%s", "compact output needs n_ok");
  CCP_NO_PENDING_TICKETS(h);
  device_guard g(h->device);
  ccp_project_args A;
  memset(&A, 0, sizeof A);
  A.seeds = seeds_dev;
  A.x_out = x_out_dev;
  A.ok = ok_dev;
  A.conv = converged_dev;
  A.iters = iters_dev;
  A.resid = resid_dev;
  A.compact = compact_dev;
  A.n_ok = (unsigned long long*)n_ok_dev;
  A.count = count;
  return launch_project(h, A, layout, (cudaStream_t)stream);
}

static int fill_sampler_args(ccp_handle* h, const ccp_sampler_args* a, int64_t count, ccp_project_args* A);

int ccp_project_batch_pipelined(ccp_handle* h, const double* seeds_dev, int64_t count, int32_t layout, double* x_out_dev,
                                uint8_t* ok_dev, uint8_t* converged_dev, int32_t* iters_dev, double* resid_dev,
                                double* compact_dev, int64_t* n_ok_dev, void* stream) {
  int rc = check_common(h, seeds_dev, count, layout);
  if (rc) return rc;
  if (compact_dev && !n_ok_dev) return set_err(h, CCP_ERR_INVALID, "%s", "compact output needs n_ok");
  CCP_NO_PENDING_TICKETS(h);
  device_guard g(h->device);
  ccp_project_args A;
  memset(&A, 0, sizeof A);
  A.seeds = seeds_dev;
  A.x_out = x_out_dev;
  A.ok = ok_dev;
  A.conv = converged_dev;
  A.iters = iters_dev;
  A.resid = resid_dev;
  A.compact = compact_dev;
  A.n_ok = (unsigned long long*)n_ok_dev;
  A.count = count;
  return launch_project(h, A, layout, (cudaStream_t)stream, true);
}

static int sample_project_impl(ccp_handle* h, const ccp_sampler_args* a, int64_t count, int32_t layout, double* x_out_dev,
                               uint8_t* ok_dev, int32_t* iters_dev, double* compact_dev, int64_t* n_ok_dev, void* stream,
                               bool defer);

int ccp_sample_project_batch_pipelined(ccp_handle* h, const ccp_sampler_args* a, int64_t count, int32_t layout,
                                       double* x_out_dev, uint8_t* ok_dev, int32_t* iters_dev, double* compact_dev,
                                       int64_t* n_ok_dev, void* stream) {
  if (!h) return CCP_ERR_INVALID;
  CCP_NO_PENDING_TICKETS(h);
  return sample_project_impl(h, a, count, layout, x_out_dev, ok_dev, iters_dev, compact_dev, n_ok_dev, stream, true);
}

int ccp_project_flush(ccp_handle* h, double* compact_dev, int64_t* n_ok_dev, void* stream) {
  if (!h) return CCP_ERR_INVALID;
  if (compact_dev && !n_ok_dev) return set_err(h, CCP_ERR_INVALID, "%s", "compact output needs n_ok");
  CCP_NO_PENDING_TICKETS(h);
  if (!h->pipeline_open) return CCP_OK;
  device_guard g(h->device);
  ccp_project_args A;
  memset(&A, 0, sizeof A);
  A.compact = compact_dev;
  A.n_ok = (unsigned long long*)n_ok_dev;
  return launch_project(h, A, (h->pipe_sig & 1) ? CCP_LAYOUT_SOA : CCP_LAYOUT_AOS, (cudaStream_t)stream, false);
}

int ccp_set_gather_peers(ccp_handle* h, int32_t world, int32_t rank, const uint64_t* pool_dev_ptrs, int64_t capacity) {
  if (!h) return CCP_ERR_INVALID;
  h->peer_mc = nullptr;
  if (world == 0) {
    h->peer_world = 0;
    return CCP_OK;
  }
  if (world < 1 || world > CCP_MAX_PEERS || rank < 0 || rank >= world || !pool_dev_ptrs || capacity < 1)
    return set_err(h, CCP_ERR_INVALID, "%s", "gather peers: 1 <= world <= 8, 0 <= rank < world, capacity >= 1");
  for (int p = 0; p < world; ++p) {
    if (!pool_dev_ptrs[p]) return set_err(h, CCP_ERR_INVALID, "%s", "gather peers: null pool pointer");
    if (pool_dev_ptrs[p] & 15u) return set_err(h, CCP_ERR_INVALID, "%s", "gather peers: pools must be 16-byte aligned");
    h->peer_pool[p] = (double*)(uintptr_t)pool_dev_ptrs[p];
  }
  h->peer_world = world;
  h->peer_rank = rank;
  h->peer_cap = capacity;
  return CCP_OK;
}

int ccp_set_gather_multicast(ccp_handle* h, uint64_t pool_multicast_dev_ptr) {
  if (!h) return CCP_ERR_INVALID;
  if (pool_multicast_dev_ptr && h->peer_world < 1)
    return set_err(h, CCP_ERR_STATE, "%s", "gather multicast: call ccp_set_gather_peers first");
  if (pool_multicast_dev_ptr & 15u) return set_err(h, CCP_ERR_INVALID, "%s", "gather multicast: the mapping must be 16-byte aligned");
  h->peer_mc = (double*)(uintptr_t)pool_multicast_dev_ptr;
  return CCP_OK;
}

int ccp_publish_count(ccp_handle* h, const int64_t* n_ok_dev, int32_t world, int32_t rank, const uint64_t* counts_dev_ptrs,
                      void* stream) {
  if (!h) return CCP_ERR_INVALID;
  if (world < 1 || world > CCP_MAX_PEERS || rank < 0 || rank >= world || !n_ok_dev || !counts_dev_ptrs)
    return set_err(h, CCP_ERR_INVALID, "%s", "publish count: 1 <= world <= 8, 0 <= rank < world, non-null pointers");
  ccp_count_peers P;
  for (int p = 0; p < CCP_MAX_PEERS; ++p) P.counts[p] = p < world ? (long long*)(uintptr_t)counts_dev_ptrs[p] : nullptr;
  device_guard g(h->device);
  ccp_publish_count_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const long long*)n_ok_dev, P, world, rank);
  h->launches++;
  CCP_CUDA(cudaGetLastError());
  return CCP_OK;
}

int ccp_project_pipeline_open(const ccp_handle* h) { return h ? (h->pipeline_open ? 1 : 0) : CCP_ERR_INVALID; }

int ccp_project_batch_timed(ccp_handle* h, const double* seeds_dev, int64_t count, int32_t layout, double* x_out_dev,
                            uint8_t* ok_dev, uint8_t* converged_dev, int32_t* iters_dev, double* resid_dev,
                            double* compact_dev, int64_t* n_ok_dev, void* stream, float* kernel_ms) {
  if (!h) return CCP_ERR_INVALID;
  device_guard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  CCP_CUDA(cudaEventRecord(h->ev0, st));
  int rc = ccp_project_batch(h, seeds_dev, count, layout, x_out_dev, ok_dev, converged_dev, iters_dev, resid_dev,
                             compact_dev, n_ok_dev, stream);
  if (rc) return rc;
  CCP_CUDA(cudaEventRecord(h->ev1, st));
  CCP_CUDA(cudaEventSynchronize(h->ev1));
  float ms = 0.f;
  CCP_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  if (kernel_ms) *kernel_ms = ms;
  return CCP_OK;
}

static int fill_sampler_args(ccp_handle* h, const ccp_sampler_args* a, int64_t count, ccp_project_args* A) {
  if (!a) return set_err(h, CCP_ERR_INVALID, "%s", "null sampler args");
  if (a->mode < 0 || a->mode > 2) return set_err(h, CCP_ERR_INVALID, "%s", "sampler mode must be 0, 1 or 2");
  if (a->mode != 0 && !a->near_host) return set_err(h, CCP_ERR_INVALID, "%s", "near state required for modes 1, 2");
  memset(A, 0, sizeof *A);
  A->count = count;
  A->gen_mode = a->mode;
  A->wrap = a->wrap_bounds;
  A->rng_seed = a->rng_seed;
  A->first_index = a->first_index;
  A->distance = a->distance;
  if (a->near_host)
    for (int j = 0; j < CCPC_DOF * h->model.n_arms; ++j) A->near[j] = a->near_host[j];
  return CCP_OK;
}

int ccp_generate_seeds(ccp_handle* h, const ccp_sampler_args* a, int64_t count, int32_t layout, double* seeds_dev,
                       void* stream) {
  int rc = check_common(h, seeds_dev, count, layout);
  if (rc) return rc;
  ccp_project_args A;
  rc = fill_sampler_args(h, a, count, &A);
  if (rc) return rc;
  if (count == 0) return CCP_OK;
  device_guard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(h, count * CCPC_DOF * h->model.n_arms, 128, 16);
  const bool soa = layout == CCP_LAYOUT_SOA;
  if (h->model.n_arms == 2) {
    if (soa) ccp_seed_kernel<2, true><<<grid, 128, 0, st>>>(h->model, A, seeds_dev);
    else ccp_seed_kernel<2, false><<<grid, 128, 0, st>>>(h->model, A, seeds_dev);
  } else {
    if (soa) ccp_seed_kernel<3, true><<<grid, 128, 0, st>>>(h->model, A, seeds_dev);
    else ccp_seed_kernel<3, false><<<grid, 128, 0, st>>>(h->model, A, seeds_dev);
  }
  h->launches++;
  CCP_CUDA(cudaGetLastError());
  return CCP_OK;
}

}  // extern "C"
template <int K, bool SOA>
static void launch_seed_kernel(const ccp_handle* h, const ccp_project_args& A, double* out, cudaStream_t st) {
  const int grid = grid_for(h, A.count * CCPC_DOF * K, 128, 16);
  ccp_seed_kernel<K, SOA><<<grid, 128, 0, st>>>(h->model, A, out);
}
extern "C" {

// sample -> project -> (wrap) -> compact.  The seeds of a chunk are generated by the seed kernel into a
// stream-ordered scratch buffer (their HBM round trip is ~0.1 % of the projection time) and projected by the same
// kernel as caller-provided seeds; a batch larger than CCP_SAMPLE_CHUNK is cut into pipelined launches so the
// scratch stays small and there is one tail for the whole call.
#define CCP_SAMPLE_CHUNK (2 * 1024 * 1024)
static int sample_project_impl(ccp_handle* h, const ccp_sampler_args* a, int64_t count, int32_t layout, double* x_out_dev,
                               uint8_t* ok_dev, int32_t* iters_dev, double* compact_dev, int64_t* n_ok_dev, void* stream,
                               bool defer) {
  if (!h) return CCP_ERR_INVALID;
  if (count < 0) return set_err(h, CCP_ERR_INVALID, "%s", "negative count");
  if (layout != CCP_LAYOUT_AOS && layout != CCP_LAYOUT_SOA) return set_err(h, CCP_ERR_INVALID, "%s", "bad layout");
  if (compact_dev && !n_ok_dev) return set_err(h, CCP_ERR_INVALID, "%s", "compact output needs n_ok");
  if (!h->has_ref) return set_err(h, CCP_ERR_STATE, "%s", "project before ccp_set_reference (setInitialPosition)");
  ccp_project_args S;
  int rc = fill_sampler_args(h, a, count, &S);
  if (rc) return rc;
  device_guard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (count == 0) {
    if (!defer && h->pipeline_open) {
      ccp_project_args F;
      memset(&F, 0, sizeof F);
      F.compact = compact_dev;
      F.n_ok = (unsigned long long*)n_ok_dev;
      return launch_project(h, F, layout, st, false);
    }
    return CCP_OK;
  }
  const int n = CCPC_DOF * h->model.n_arms;
  const bool soa = layout == CCP_LAYOUT_SOA;
  const int64_t chunk = count < CCP_SAMPLE_CHUNK ? count : CCP_SAMPLE_CHUNK;
  double* scratch = nullptr;
  CCP_CUDA(cudaMallocFromPoolAsync((void**)&scratch, sizeof(double) * n * (size_t)chunk, h->pool, st));
  for (int64_t off = 0; off < count && rc == CCP_OK; off += chunk) {
    const int64_t cn = (count - off < chunk) ? (count - off) : chunk;
    S.count = cn;
    S.first_index = a->first_index + off;
    if (h->model.n_arms == 2) {
      if (soa) launch_seed_kernel<2, true>(h, S, scratch, st); else launch_seed_kernel<2, false>(h, S, scratch, st);
    } else {
      if (soa) launch_seed_kernel<3, true>(h, S, scratch, st); else launch_seed_kernel<3, false>(h, S, scratch, st);
    }
    h->launches++;
    ccp_project_args A;
    memset(&A, 0, sizeof A);
    A.seeds = scratch;
    A.seed_stride = cn;
    A.count = cn;
    A.out_stride = count;
    A.wrap = a->wrap_bounds;
    A.x_out = x_out_dev ? (soa ? x_out_dev + off : x_out_dev + off * n) : nullptr;
    A.ok = ok_dev ? ok_dev + off : nullptr;
    A.iters = iters_dev ? iters_dev + off : nullptr;
    A.compact = compact_dev;
    A.n_ok = (unsigned long long*)n_ok_dev;
    const bool last = off + cn >= count;
    rc = launch_project(h, A, layout, st, defer || !last);
  }
  cudaError_t e = cudaFreeAsync(scratch, st);
  if (rc) return rc;
  if (e != cudaSuccess) return set_err(h, CCP_ERR_CUDA, "cudaFreeAsync: %s", cudaGetErrorString(e));
  return CCP_OK;
}

int ccp_sample_project_batch(ccp_handle* h, const ccp_sampler_args* a, int64_t count, int32_t layout,
                             double* x_out_dev, uint8_t* ok_dev, int32_t* iters_dev, double* compact_dev,
                             int64_t* n_ok_dev, void* stream) {
  if (!h) return CCP_ERR_INVALID;
  CCP_NO_PENDING_TICKETS(h);
  return sample_project_impl(h, a, count, layout, x_out_dev, ok_dev, iters_dev, compact_dev, n_ok_dev, stream, false);
}

// the geodesic kernels' crossover, scaled with the handle's cooperative threshold (ccp_set_coop_threshold: 0 = never)
static long long geo_coop_max(const ccp_handle* h) {
  const bool two = h->model.n_arms == 2;
  const double ratio = two ? (double)CCP_GEO_COOP_PER_SM / CCP_COOP_MAX_PER_SM : (double)CCP_GEO_COOP3_PER_SM / CCP_COOP3_MAX_PER_SM;
  const double v = (double)h->coop_max * ratio;
  return v > 4.0e18 ? (long long)4e18 : (long long)v;
}

int ccp_geodesic_batch(ccp_handle* h, const double* from_dev, const double* to_dev, int64_t edges, double delta,
                       double lambda, int32_t max_states, double* states_dev, int32_t* n_states_dev,
                       uint8_t* reached_dev, int32_t* iters_dev, void* stream) {
  int rc = check_common(h, from_dev, edges, CCP_LAYOUT_AOS);
  if (rc) return rc;
  if (edges > 0 && (!to_dev || !states_dev || !n_states_dev || !reached_dev))
    return set_err(h, CCP_ERR_INVALID, "%s", "null geodesic buffer");
  if (!(delta > 0) || !(lambda > 0) || max_states < 1) return set_err(h, CCP_ERR_INVALID, "%s", "bad geodesic parameters");
  if (!h->has_ref) return set_err(h, CCP_ERR_STATE, "%s", "geodesic before ccp_set_reference (setInitialPosition)");
  if (edges == 0) return CCP_OK;
  device_guard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  unsigned slot;
  {
    std::lock_guard<std::mutex> lk(h->mu);
    slot = CCP_NUM_COUNTERS + h->geo_seq++ % CCP_NUM_COUNTERS;
    h->launches++;
  }
  unsigned long long* counter = &h->d_counters[slot].work;
  CCP_CUDA(cudaMemsetAsync(h->d_counters + slot, 0, sizeof(ccp_launch_rec), st));
  cudaError_t e = ccp_launch_geodesic(h->sm_count, h->model, from_dev, to_dev, edges, delta, lambda, max_states, states_dev,
                                      n_states_dev, reached_dev, iters_dev, counter,
                                      geo_coop_max(h) /* a walk is a serial chain of projections with frequent bookkeeping: several lanes per edge win up to ~66 000 (two arms) / ~76 000 (three arms) edges on B200 */, st);
  if (e != cudaSuccess) return set_err(h, CCP_ERR_CUDA, "geodesic kernel launch: %s", cudaGetErrorString(e));
  return CCP_OK;
}

// ---- batched pose IK (goal sampling) -------------------------------------------------------
void ccp_ik_default_options(ccp_ik_options* o) {
  if (!o) return;
  o->max_iter = 200;
  o->reserved = 0;
  o->eps_pos = 1e-5;  // TRAC-IK's default eps on every twist component
  o->eps_rot = 1e-5;
  o->damping = 1e-4;
  o->joint_margin = 1e-3;  // TrackIKAdaptor::isValid, panda_tracik.cpp:99-108
}

static int ik_options(ccp_handle* h, const ccp_ik_options* opt, ccp_ik_opt* O) {
  ccp_ik_options d;
  ccp_ik_default_options(&d);
  if (opt) d = *opt;
  if (d.max_iter < 0 || !(d.eps_pos > 0) || !(d.eps_rot > 0) || !(d.damping >= 0) || !(d.joint_margin >= 0))
    return set_err(h, CCP_ERR_INVALID, "%s", "bad IK options");
  O->max_iter = d.max_iter;
  O->pad = 0;
  O->eps_p = d.eps_pos;
  O->eps_r = d.eps_rot;
  O->lambda2 = d.damping;
  O->margin = d.joint_margin;
  return CCP_OK;
}

int ccp_ik_batch(ccp_handle* h, int32_t arm, const double* T_target_dev, const double* q_seed_dev, int64_t count,
                 const ccp_ik_options* opt, double* q_out_dev, uint8_t* ok_dev, int32_t* iters_dev, double* err_dev,
                 void* stream) {
  if (!h) return CCP_ERR_INVALID;
  if (count < 0) return set_err(h, CCP_ERR_INVALID, "%s", "negative count");
  if (arm < 0 || arm >= h->model.n_arms) return set_err(h, CCP_ERR_INVALID, "%s", "arm index out of range");
  if (count > 0 && (!T_target_dev || !q_seed_dev || !q_out_dev)) return set_err(h, CCP_ERR_INVALID, "%s", "null IK buffer");
  ccp_ik_opt O;
  int rc = ik_options(h, opt, &O);
  if (rc) return rc;
  if (count == 0) return CCP_OK;
  device_guard g(h->device);
  unsigned slot;
  {  // the geodesic ring of launch records doubles as the IK kernels' work counters
    std::lock_guard<std::mutex> lk(h->mu);
    slot = CCP_NUM_COUNTERS + h->geo_seq++ % CCP_NUM_COUNTERS;
  }
  CCP_CUDA(cudaMemsetAsync(h->d_counters + slot, 0, sizeof(ccp_launch_rec), (cudaStream_t)stream));
  cudaError_t e = ccp_launch_ik(h->sm_count, h->model, arm, T_target_dev, q_seed_dev, count, O, q_out_dev, ok_dev, iters_dev,
                                err_dev, &h->d_counters[slot].work, (cudaStream_t)stream);
  h->launches++;
  if (e != cudaSuccess) return set_err(h, CCP_ERR_CUDA, "IK kernel launch: %s", cudaGetErrorString(e));
  return CCP_OK;
}

int ccp_ik_sample_batch(ccp_handle* h, int32_t arm, const double* T_target_dev, int64_t n_targets, int32_t restarts,
                        uint64_t rng_seed, double sigma, const double* q_ref_dev, const ccp_ik_options* opt,
                        double* q_best_dev, uint8_t* ok_dev, int32_t* n_success_dev, void* stream) {
  if (!h) return CCP_ERR_INVALID;
  if (n_targets < 0) return set_err(h, CCP_ERR_INVALID, "%s", "negative count");
  if (arm < 0 || arm >= h->model.n_arms) return set_err(h, CCP_ERR_INVALID, "%s", "arm index out of range");
  if (restarts < 1 || restarts > 32) return set_err(h, CCP_ERR_INVALID, "%s", "IK restarts must be 1..32");
  if (!(sigma >= 0)) return set_err(h, CCP_ERR_INVALID, "%s", "IK sigma must be >= 0");
  if (n_targets > 0 && (!T_target_dev || !q_best_dev || !ok_dev)) return set_err(h, CCP_ERR_INVALID, "%s", "null IK buffer");
  ccp_ik_opt O;
  int rc = ik_options(h, opt, &O);
  if (rc) return rc;
  if (n_targets == 0) return CCP_OK;
  if (n_targets > CCP_IK_SAMPLE_CHUNK * CCP_IK_SAMPLE_MAX_LAUNCHES)
    return set_err(h, CCP_ERR_INVALID, "%s", "more than 16 M IK targets in one call: split the batch");
  device_guard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  // scratch (selection keys + candidate solutions + per-target done counts) and the chunk launches' work counters
  const size_t sbytes = (ccp_ik_sample_scratch_bytes(n_targets, restarts) + 255) & ~(size_t)255;
  const size_t cbytes = sizeof(unsigned long long) * CCP_IK_SAMPLE_MAX_LAUNCHES;
  char* scratch = nullptr;
  CCP_CUDA(cudaMallocFromPoolAsync((void**)&scratch, sbytes + cbytes, h->pool, st));
  unsigned long long* counters = (unsigned long long*)(scratch + sbytes);
  cudaError_t e = cudaMemsetAsync(counters, 0, cbytes, st);
  if (e == cudaSuccess)
    e = ccp_launch_ik_sample(h->sm_count, h->model, arm, T_target_dev, q_ref_dev, n_targets, restarts, rng_seed, sigma, O,
                             q_best_dev, ok_dev, n_success_dev, scratch, counters, CCPC_DOF, st);
  h->launches += (n_targets + CCP_IK_SAMPLE_CHUNK - 1) / CCP_IK_SAMPLE_CHUNK;
  cudaFreeAsync(scratch, st);
  if (e != cudaSuccess) return set_err(h, CCP_ERR_CUDA, "IK sample kernel launch: %s", cudaGetErrorString(e));
  return CCP_OK;
}

int ccp_goal_sample_batch(ccp_handle* h, const double* T_obj_dev, int64_t n, const double* t_o7_host, const double* q_ref_dev,
                          int32_t restarts, uint64_t rng_seed, double sigma, const ccp_ik_options* opt, double* q_out_dev,
                          uint8_t* ok_dev, void* stream) {
  if (!h) return CCP_ERR_INVALID;
  if (n < 0) return set_err(h, CCP_ERR_INVALID, "%s", "negative count");
  if (restarts < 1 || restarts > 32) return set_err(h, CCP_ERR_INVALID, "%s", "IK restarts must be 1..32");
  if (!(sigma >= 0)) return set_err(h, CCP_ERR_INVALID, "%s", "IK sigma must be >= 0");
  if (!t_o7_host || (n > 0 && (!T_obj_dev || !q_out_dev || !ok_dev))) return set_err(h, CCP_ERR_INVALID, "%s", "null goal-sampling buffer");
  if (n > CCP_IK_SAMPLE_CHUNK * CCP_IK_SAMPLE_MAX_LAUNCHES)
    return set_err(h, CCP_ERR_INVALID, "%s", "more than 16 M object poses in one call: split the batch");
  ccp_ik_opt O;
  int rc = ik_options(h, opt, &O);
  if (rc) return rc;
  if (n == 0) return CCP_OK;
  device_guard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int K = h->model.n_arms, nq = CCPC_DOF * K;
  // scratch: per-arm targets [K][n][12] | per-arm verdicts [K][n] | IK selection scratch | one work counter per chunk launch and arm
  const size_t tb = (sizeof(double) * 12 * (size_t)K * (size_t)n + 255) & ~(size_t)255;
  const size_t ob = ((size_t)K * (size_t)n + 255) & ~(size_t)255;
  const size_t sb = (ccp_ik_sample_scratch_bytes(n, restarts) + 255) & ~(size_t)255;
  const size_t cb = sizeof(unsigned long long) * CCP_IK_SAMPLE_MAX_LAUNCHES * CCPC_MAX_ARMS;
  char* base = nullptr;
  CCP_CUDA(cudaMallocFromPoolAsync((void**)&base, tb + ob + sb + cb, h->pool, st));
  double* Tt = (double*)base;
  uint8_t* ok_arm = (uint8_t*)(base + tb);
  void* scratch = base + tb + ob;
  unsigned long long* counters = (unsigned long long*)(base + tb + ob + sb);
  cudaError_t e = cudaMemsetAsync(counters, 0, cb, st);
  if (e == cudaSuccess) e = ccp_launch_goal_targets(h->sm_count, h->model, t_o7_host, T_obj_dev, n, Tt, st);
  h->launches++;
  for (int a = 0; a < K && e == cudaSuccess; ++a) {
    // every arm draws its own restarts from the stream; its seven joints sit at column 7 a of the 7K-wide rows
    e = ccp_launch_ik_sample(h->sm_count, h->model, a, Tt + (size_t)a * n * 12, q_ref_dev ? q_ref_dev + a * CCPC_DOF : nullptr, n,
                             restarts, rng_seed + 0x9E3779B97F4A7C15ull * (uint64_t)a, sigma, O, q_out_dev + a * CCPC_DOF,
                             ok_arm + (size_t)a * n, nullptr, scratch, counters + a * CCP_IK_SAMPLE_MAX_LAUNCHES, nq, st);
    h->launches += (n + CCP_IK_SAMPLE_CHUNK - 1) / CCP_IK_SAMPLE_CHUNK;
  }
  if (e == cudaSuccess) e = ccp_launch_goal_combine(h->sm_count, ok_arm, n, K, ok_dev, st);
  h->launches++;
  cudaFreeAsync(base, st);
  if (e != cudaSuccess) return set_err(h, CCP_ERR_CUDA, "goal sampling launch: %s", cudaGetErrorString(e));
  return CCP_OK;
}

int ccp_goal_sample_batch_host(ccp_handle* h, const double* T_obj_host, int64_t n, const double* t_o7_host,
                               const double* q_ref_host, int32_t restarts, uint64_t rng_seed, double sigma,
                               const ccp_ik_options* opt, double* q_out_host, uint8_t* ok_host) {
  if (!h) return CCP_ERR_INVALID;
  if (n < 0) return set_err(h, CCP_ERR_INVALID, "%s", "negative count");
  if (n > 0 && (!T_obj_host || !t_o7_host || !q_out_host || !ok_host)) return set_err(h, CCP_ERR_INVALID, "%s", "null goal-sampling buffer");
  if (n == 0) return CCP_OK;
  std::lock_guard<std::mutex> host_lock(h->host_mu);
  device_guard g(h->device);
  cudaStream_t st = h->hstream[1];
  const int nq = CCPC_DOF * h->model.n_arms;
  const size_t tb = sizeof(double) * 12 * (size_t)n, qb = sizeof(double) * nq * (size_t)n;
  char* base = nullptr;
  CCP_CUDA(cudaMallocFromPoolAsync((void**)&base, tb + 2 * qb + (size_t)n + 64, h->pool, st));
  double* dT = (double*)base;
  double* dref = (double*)(base + tb);
  double* dq = (double*)(base + tb + qb);
  uint8_t* dok = (uint8_t*)(base + tb + 2 * qb);
  cudaError_t e = cudaMemcpyAsync(dT, T_obj_host, tb, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && q_ref_host) e = cudaMemcpyAsync(dref, q_ref_host, qb, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(dq, 0, qb, st);
  int rc = CCP_OK;
  if (e == cudaSuccess)
    rc = ccp_goal_sample_batch(h, dT, n, t_o7_host, q_ref_host ? dref : nullptr, restarts, rng_seed, sigma, opt, dq, dok, st);
  if (rc == CCP_OK && e == cudaSuccess) e = cudaMemcpyAsync(q_out_host, dq, qb, cudaMemcpyDeviceToHost, st);
  if (rc == CCP_OK && e == cudaSuccess) e = cudaMemcpyAsync(ok_host, dok, (size_t)n, cudaMemcpyDeviceToHost, st);
  cudaFreeAsync(base, st);
  cudaError_t e2 = cudaStreamSynchronize(st);
  if (rc) return rc;
  if (e != cudaSuccess || e2 != cudaSuccess)
    return set_err(h, CCP_ERR_CUDA, "ccp_goal_sample_batch_host: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
  return CCP_OK;
}

int ccp_enforce_bounds_batch(ccp_handle* h, double* x_dev, int64_t count, int32_t layout, void* stream) {
  int rc = check_common(h, x_dev, count, layout);
  if (rc) return rc;
  if (count == 0) return CCP_OK;
  device_guard g(h->device);
  const long long total = (long long)count * CCPC_DOF * h->model.n_arms;
  ccp_wrap_kernel<<<grid_for(h, total, 256, 8), 256, 0, (cudaStream_t)stream>>>(x_dev, total);
  h->launches++;
  CCP_CUDA(cudaGetLastError());
  return CCP_OK;
}

// ---- host-buffer entry points ------------------------------------------------------------
// ccp_project_batch_host: the batch is cut into chunks that flow through three private streams,
//   C: H2D chunk 0 | H2D chunk 1 | ...
//   K:      project 0 | project 1 | ... | flush            (PIPELINED launches: a chunk's stragglers are parked and
//   D:                 D2H chunk 0 | ...        | D2H last   adopted by the next chunk's launch, so there is ONE tail
// for the whole call however fine the chunks are, and chunk c's outputs are complete after launch c + 1).  Everything
// is ordered by events; the host only waits at the end.  With pinned caller buffers the copies run at PCIe rate
// behind the kernels; with pageable buffers CUDA stages them and the call is still correct.
// ---- host path ----------------------------------------------------------------------------------------------------
// A host batch is cut into chunks that flow through three private streams — H2D chunk c | pipelined projection launch c
// | D2H of the chunk launch c has COMPLETED — ordered by events.  A pipelined launch parks what is still iterating when
// its seed list runs dry and the next launch adopts it, so a sample may be carried through several launches; the
// launches of the host path carry max_age = lag, so that launch c + lag finishes whatever chunk c still has in flight
// and chunk c's outputs are copied behind launch c + lag.  Whatever is pending at the end is completed by one flush
// launch.  ccp_project_batch_host_submit returns a ticket without waiting and ccp_project_batch_host_wait blocks until
// that batch's results are in the caller's buffers: with two batches in flight the next batch's launches complete the
// previous batch's last chunks, no launch tail is paid and the copies of one batch run behind the kernels of the other.
static bool zero_copy_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("CCP_HOST_ZERO_COPY");
    on = (e && atoi(e) == 0) ? 0 : 1;
  }
  return on != 0;
}

static int host_lag() {
  static int lag = -1;
  if (lag < 0) {
    const char* e = getenv("CCP_HOST_LAG");
    lag = e ? atoi(e) : CCP_HOST_LAG_DEFAULT;
    if (lag < 1) lag = 1;
    if (lag > CCP_MAX_AGE) lag = CCP_MAX_AGE;
  }
  return lag;
}

static int host_copy_out(ccp_handle* h, const ccp_host_call& c, int chunk_index, cudaStream_t sD) {
  const int n = CCPC_DOF * h->model.n_arms, m = 2 * (h->model.n_arms - 1);
  const int64_t off = (int64_t)chunk_index * c.chunk;
  const int64_t cn = (c.count - off < c.chunk) ? (c.count - off) : c.chunk;
  if (c.x_out_host)
    CCP_CUDA(cudaMemcpyAsync(c.x_out_host + off * n, c.dx + off * n, sizeof(double) * n * cn, cudaMemcpyDeviceToHost, sD));
  if (c.ok_host) CCP_CUDA(cudaMemcpyAsync(c.ok_host + off, c.dok + off, (size_t)cn, cudaMemcpyDeviceToHost, sD));
  if (c.conv_host) CCP_CUDA(cudaMemcpyAsync(c.conv_host + off, c.dcv + off, (size_t)cn, cudaMemcpyDeviceToHost, sD));
  if (c.iters_host) CCP_CUDA(cudaMemcpyAsync(c.iters_host + off, c.dit + off, sizeof(int32_t) * cn, cudaMemcpyDeviceToHost, sD));
  if (c.resid_host)
    CCP_CUDA(cudaMemcpyAsync(c.resid_host + off * m, c.dres + off * m, sizeof(double) * m * cn, cudaMemcpyDeviceToHost, sD));
  return CCP_OK;
}

// Behind a batch's last chunk copy: the number of packed ok rows (compact outputs), then the batch's `done` event.
static int host_finish_copies(ccp_handle* h, ccp_host_call& q, cudaStream_t sD) {
  if (q.compact)
    CCP_CUDA(cudaMemcpyAsync(q.pin_nok, q.dnok, sizeof(long long), cudaMemcpyDeviceToHost, sD));
  CCP_CUDA(cudaEventRecord(q.done, sD));
  return CCP_OK;
}

// Copy out every chunk of q that the host-path launch numbered `g` (just enqueued, event `ev`) has completed.
static int host_copy_ready(ccp_handle* h, ccp_host_call& q, int64_t g, cudaEvent_t ev, bool& waited) {
  cudaStream_t sD = h->hstream[2];
  while (q.copied < q.parts && q.first_launch + q.copied + host_lag() <= g) {
    if (!waited) {
      CCP_CUDA(cudaStreamWaitEvent(sD, ev, 0));
      waited = true;
    }
    int rc = host_copy_out(h, q, q.copied, sD);
    if (rc) return rc;
    if (++q.copied == q.parts) {
      rc = host_finish_copies(h, q, sD);
      if (rc) return rc;
    }
  }
  return CCP_OK;
}

// Complete every submitted batch: one flush launch, then the chunks still on the device (older batch first).
// The caller holds host_mu.
static int host_drain_locked(ccp_handle* h) {
  ccp_host_call* q[2] = {&h->hcall[0], &h->hcall[1]};
  if (q[0]->ticket > q[1]->ticket) std::swap(q[0], q[1]);
  if (q[0]->copied == q[0]->parts && q[1]->copied == q[1]->parts) return CCP_OK;
  cudaStream_t sK = h->hstream[1], sD = h->hstream[2];
  if (h->pipeline_open) {
    ccp_project_args F;
    memset(&F, 0, sizeof F);
    int rc = launch_project(h, F, CCP_LAYOUT_AOS, sK, false);
    if (rc) return rc;
  }
  CCP_CUDA(cudaEventRecord(h->ev1, sK));
  CCP_CUDA(cudaStreamWaitEvent(sD, h->ev1, 0));
  for (int i = 0; i < 2; ++i)
    while (q[i]->copied < q[i]->parts) {
      int rc = host_copy_out(h, *q[i], q[i]->copied, sD);
      if (rc) return rc;
      if (++q[i]->copied == q[i]->parts) {
        rc = host_finish_copies(h, *q[i], sD);
        if (rc) return rc;
      }
    }
  return CCP_OK;
}

// use_wait scheme: put the batch's copies on the D2H stream, each behind a wait until the completion count of its
// chunk's launch slot says every sample of the chunk has finished, in whichever launch that happened.  Called only when
// the launches that make every one of these waits come true are already enqueued (see host_submit_locked), so nothing
// that cannot complete by itself is ever left pending on the device.
static int host_enqueue_copies(ccp_handle* h, ccp_host_call& c) {
  cudaStream_t sD = h->hstream[2];
  for (int k = 0; k < c.parts; ++k) {
    if (h->wait32(sD, (unsigned long long)(uintptr_t)(h->d_done + c.slot[k]), c.expect[k], 0u) != 0)
      return set_err(h, CCP_ERR_CUDA, "%s", "cuStreamWaitValue32 failed");
    int rc = host_copy_out(h, c, k, sD);
    if (rc) return rc;
  }
  int rc = host_finish_copies(h, c, sD);
  if (rc) return rc;
  c.copied = c.parts;
  c.enqueued = true;
  return CCP_OK;
}

static int host_wait_locked(ccp_handle* h, ccp_host_call& c) {
  if (c.finished) return CCP_OK;
  if (!c.use_wait) {
    if (c.copied < c.parts) {
      int rc = host_drain_locked(h);
      if (rc) return rc;
    }
  } else {
    const bool newest = c.first_launch + c.parts == h->host_launches;
    if (newest && h->pipeline_open) {  // no later batch will finish what this one parked: flush
      ccp_project_args F;
      memset(&F, 0, sizeof F);
      int rc = launch_project(h, F, CCP_LAYOUT_AOS, h->hstream[1], false);
      if (rc) return rc;
    }
    // (not the newest: the later batch's last launch finishes whatever this one still had in flight)
    if (!c.enqueued) {
      int rc = host_enqueue_copies(h, c);
      if (rc) return rc;
    }
  }
  CCP_CUDA(cudaEventSynchronize(c.done));
  if (c.compact) {
    // the packed rows, sized by the count that just arrived; the device is already busy with the next batch
    const int n = CCPC_DOF * h->model.n_arms;
    c.n_ok = *c.pin_nok;
    const int64_t rows = c.n_ok < c.ccap ? c.n_ok : c.ccap;
    cudaStream_t sD = h->hstream[2];
    if (rows > 0) {
      if (c.compact_host)
        CCP_CUDA(cudaMemcpyAsync(c.compact_host, c.dcompact, sizeof(double) * n * (size_t)rows, cudaMemcpyDeviceToHost, sD));
      if (c.cidx_host)
        CCP_CUDA(cudaMemcpyAsync(c.cidx_host, c.dcidx, sizeof(int32_t) * (size_t)rows, cudaMemcpyDeviceToHost, sD));
      CCP_CUDA(cudaEventRecord(c.done, sD));
      CCP_CUDA(cudaEventSynchronize(c.done));
    }
  }
  c.finished = true;
  return CCP_OK;
}

// The caller holds host_mu and has checked the arguments.
static int host_submit_locked(ccp_handle* h, const ccp_host_batch& B, int64_t* ticket_out) {
  const double* seeds_host = B.seeds_host;
  const int64_t count = B.count;
  double* x_out_host = B.x_out_host;
  uint8_t* ok_host = B.ok_host;
  uint8_t* converged_host = B.converged_host;
  int32_t* iters_host = B.iters_host;
  double* resid_host = B.resid_host;
  const bool compact = B.compact_host != nullptr || B.compact_index_host != nullptr || B.compact_capacity > 0;
  ccp_project_args S;  // seeds generated on the device
  if (B.sampler) {
    int src = fill_sampler_args(h, B.sampler, count, &S);
    if (src) return src;
  }
  ccp_host_call& c = h->hcall[h->next_ticket & 1];
  ccp_host_call& prev = h->hcall[(h->next_ticket & 1) ^ 1];
  int rc = host_wait_locked(h, c);  // the batch that used this slot two submits ago
  if (rc) return rc;
  // A D2H copy into pageable memory blocks the host until it has run, so it can only be issued once the launch that
  // completes its chunk is enqueued (the `lag` scheme).  With pinned outputs the copy is enqueued at once behind a
  // stream wait on the chunk's completion count.
  bool use_wait = h->wait32 != nullptr;
  {
    const void* outs[7] = {x_out_host, ok_host, converged_host, iters_host, resid_host, B.compact_host, B.compact_index_host};
    for (int i = 0; i < 7 && use_wait; ++i) {
      if (!outs[i]) continue;
      cudaPointerAttributes pa;
      if (cudaPointerGetAttributes(&pa, outs[i]) != cudaSuccess || pa.type == cudaMemoryTypeUnregistered) use_wait = false;
    }
    cudaGetLastError();
  }
  if (!prev.finished && prev.use_wait != use_wait) {  // the two schemes do not share a pipeline
    rc = host_wait_locked(h, prev);
    if (rc) return rc;
  }
  if (h->pipeline_open && (prev.finished || (!prev.use_wait && prev.copied == prev.parts)))
    return set_err(h, CCP_ERR_STATE, "%s", "pipelined projections are in flight: call ccp_project_flush first");
  const int n = CCPC_DOF * h->model.n_arms, m = 2 * (h->model.n_arms - 1);
  const size_t per = sizeof(double) * n + sizeof(double) * m + 8 /* ok, conv, pad */ + sizeof(int32_t) + 4;
  // packed rows: as many as the caller's buffers hold, never more than the batch
  const int64_t ccap = compact ? (B.compact_capacity < count ? B.compact_capacity : count) : 0;
  const size_t cbytes = compact ? (sizeof(double) * n + sizeof(int32_t)) * (size_t)ccap + 64 : 0;
  const size_t need = per * (size_t)count + 1024 + cbytes;
  cudaStream_t sC = h->hstream[0], sK = h->hstream[1];
  if (need > c.stage_bytes) {
    // Stream-ordered, from the handle's pool: no device-wide synchronisation.  (cudaFree would wait for everything
    // pending on the device — and the previous batch's last chunks are only finished by THIS batch's launches.)  The
    // old stage was last touched by the batch two submits back, which host_wait_locked above has seen complete.
    if (c.stage) CCP_CUDA(cudaFreeAsync(c.stage, sC));
    c.stage = nullptr;
    c.stage_bytes = 0;
    CCP_CUDA(cudaMallocFromPoolAsync(&c.stage, need + need / 4, h->pool, sC));
    c.stage_bytes = need + need / 4;
  }
  // Every fallible host-side check is behind us.  The previous batch's copies (use_wait scheme) go on the D2H stream
  // only AFTER this batch's launches are enqueued (below): their stream waits are satisfied by these launches, and a
  // wait whose launches never came would block the D2H stream — and ccp_destroy's synchronisation — for ever.
  c.use_wait = use_wait;
  c.enqueued = false;
  // carve the stage: x [count][n] | resid [count][m] | iters | ok | conv
  char* base = (char*)c.stage;
  c.dx = (double*)base;
  c.dres = (double*)(base + sizeof(double) * n * (size_t)count);
  c.dit = (int32_t*)((char*)c.dres + sizeof(double) * m * (size_t)count);
  c.dok = (uint8_t*)((char*)c.dit + sizeof(int32_t) * (size_t)count);
  c.dcv = c.dok + (size_t)count;
  c.compact = compact;
  c.compact_host = B.compact_host;
  c.cidx_host = B.compact_index_host;
  c.ccap = ccap;
  c.n_ok = 0;
  if (compact) {
    char* cb = base + ((per * (size_t)count + 1024 - 64) & ~(size_t)63);
    c.dnok = (unsigned long long*)cb;
    c.dcompact = (double*)(cb + 64);
    c.dcidx = (int32_t*)(cb + 64 + sizeof(double) * n * (size_t)ccap);
    CCP_CUDA(cudaMemsetAsync(c.dnok, 0, 64, sC));  // on the copy stream: behind the stage's allocation, ahead of the first chunk event
  }
  c.count = count;
  c.x_out_host = x_out_host;
  c.ok_host = ok_host;
  c.conv_host = converged_host;
  c.iters_host = iters_host;
  c.resid_host = resid_host;
  // chunking.  A blocking call wants many chunks — its first H2D and the D2H of its last chunks are hidden by nothing
  // and shrink with the chunk — but every launch boundary costs a drain and a refill of the machine: ~2x the resident
  // lane count per chunk (1 M seeds: 8 chunks 4.74 ms, 4 chunks 5.34, 24 chunks 5.26).  With a previous batch still in
  // flight (the streaming use) that batch hides this one's first copy and the next one its last, and fewer, larger
  // chunks win: ~5x the lane count (1 M seeds per batch: 3 chunks 3.42 ms, 8 chunks 3.54, 2 chunks 4.00;
  // tools/e2e_stream.py).
  int parts;
  {
    static int env_parts = -1;
    if (env_parts < 0) {
      const char* e = getenv("CCP_HOST_CHUNKS");
      env_parts = e ? atoi(e) : 0;
    }
    const int64_t lanes = (int64_t)h->sm_count * 384;
    const bool streaming = !prev.finished;
    parts = env_parts > 0 ? env_parts : (int)(count / ((streaming ? 5 : 2) * lanes));
    if (streaming && env_parts <= 0 && parts < 3 && count >= 6 * lanes) parts = 3;
    if (parts > CCP_HOST_MAX_CHUNKS) parts = CCP_HOST_MAX_CHUNKS;
    if (parts < 1) parts = 1;
  }
  int64_t chunk = (count + parts - 1) / parts;
  chunk = (chunk + 15) / 16 * 16;  // chunks never share a 128 B line of any output array
  parts = (int)((count + chunk - 1) / chunk);
  c.chunk = chunk;
  c.parts = parts;
  c.copied = 0;
  c.first_launch = h->host_launches;
  c.finished = false;
  c.ticket = h->next_ticket++;
  for (int k = 0; k < parts; ++k) {
    const int64_t off = (int64_t)k * chunk;
    const int64_t cn = (count - off < chunk) ? (count - off) : chunk;
    if (B.sampler) {
      // the chunk's seeds come from the counter-based seed kernel on the projection stream: nothing crosses PCIe
      if (k == 0) {  // the stage (re)allocation on the copy stream precedes the first kernel that touches it
        CCP_CUDA(cudaEventRecord(h->ev_chunk_in[0], sC));
        CCP_CUDA(cudaStreamWaitEvent(sK, h->ev_chunk_in[0], 0));
      }
      S.count = cn;
      S.first_index = B.sampler->first_index + off;
      if (h->model.n_arms == 2) launch_seed_kernel<2, false>(h, S, c.dx + off * n, sK);
      else launch_seed_kernel<3, false>(h, S, c.dx + off * n, sK);
      h->launches++;
    } else {
      CCP_CUDA(cudaMemcpyAsync(c.dx + off * n, seeds_host + off * n, sizeof(double) * n * cn, cudaMemcpyHostToDevice, sC));
      CCP_CUDA(cudaEventRecord(h->ev_chunk_in[k], sC));
      CCP_CUDA(cudaStreamWaitEvent(sK, h->ev_chunk_in[k], 0));
    }
    ccp_project_args A;
    memset(&A, 0, sizeof A);
    A.seeds = c.dx + off * n;
    A.x_out = x_out_host ? c.dx + off * n : nullptr;
    A.ok = ok_host ? c.dok + off : nullptr;
    A.conv = converged_host ? c.dcv + off : nullptr;
    A.iters = iters_host ? c.dit + off : nullptr;
    A.resid = resid_host ? c.dres + off * m : nullptr;
    A.wrap = B.sampler ? B.sampler->wrap_bounds : 0;
    if (compact) {
      A.own_n_ok = c.dnok;
      A.own_compact = B.compact_host ? c.dcompact : nullptr;
      A.own_compact_idx = B.compact_index_host ? c.dcidx : nullptr;
      A.own_compact_cap = ccap;
      A.idx_base = (unsigned)off;
    }
    A.count = cn;
    // use_wait: the batch's LAST launch finishes every sample older than the batch (age >= parts), so that once a
    // batch is submitted the one before it is complete without any further call; nothing else is ever forced.
    A.max_age = !use_wait ? (unsigned)host_lag() : (k == parts - 1 ? (unsigned)parts : (unsigned)CCP_MAX_AGE);
    rc = launch_project(h, A, CCP_LAYOUT_AOS, sK, /*defer=*/true);
    if (rc) return rc;
    const int64_t g = h->host_launches++;
    if (use_wait) {
      c.slot[k] = h->last_slot;
      c.expect[k] = h->done_total[h->last_slot];
    } else {
      CCP_CUDA(cudaEventRecord(h->ev_chunk_k[k], sK));
      bool waited = false;
      rc = host_copy_ready(h, prev, g, h->ev_chunk_k[k], waited);
      if (rc) return rc;
      rc = host_copy_ready(h, c, g, h->ev_chunk_k[k], waited);
      if (rc) return rc;
    }
  }
  if (use_wait && !prev.finished && !prev.enqueued) {  // this batch's launches complete the previous one
    rc = host_enqueue_copies(h, prev);
    if (rc) return rc;
  }
  if (ticket_out) *ticket_out = c.ticket;
  return CCP_OK;
}

int ccp_project_batch_host(ccp_handle* h, const double* seeds_host, int64_t count, double* x_out_host,
                           uint8_t* ok_host, uint8_t* converged_host, int32_t* iters_host, double* resid_host) {
  int rc = check_common(h, seeds_host, count, CCP_LAYOUT_AOS);
  if (rc) return rc;
  if (!h->has_ref) return set_err(h, CCP_ERR_STATE, "%s", "project before ccp_set_reference (setInitialPosition)");
  if (count == 0) return CCP_OK;
  std::lock_guard<std::mutex> host_lock(h->host_mu);
  device_guard g(h->device);
  for (int i = 0; i < 2; ++i) {  // batches submitted through the streaming form complete first
    rc = host_wait_locked(h, h->hcall[i]);
    if (rc) return rc;
  }
  CCP_NO_OPEN_PIPELINE(h);  // this call runs its own pipeline on the handle's private streams
  if (count >= 4 * (int64_t)h->sm_count * 384) {  // at least two chunks
    int64_t ticket = 0;
    ccp_host_batch B;
    memset(&B, 0, sizeof B);
    B.seeds_host = seeds_host;
    B.count = count;
    B.x_out_host = x_out_host;
    B.ok_host = ok_host;
    B.converged_host = converged_host;
    B.iters_host = iters_host;
    B.resid_host = resid_host;
    rc = host_submit_locked(h, B, &ticket);
    if (rc) return rc;
    return host_wait_locked(h, h->hcall[ticket & 1]);
  }
  const int n = CCPC_DOF * h->model.n_arms, m = 2 * (h->model.n_arms - 1);
  const size_t per = sizeof(double) * n + sizeof(double) * m + 8 /* ok, conv, pad */ + sizeof(int32_t) + 4;
  if (count <= CCP_ZERO_COPY_MAX && zero_copy_enabled()) {
    // A handful of states — what a planner projecting state by state sends: no copy operations at all.  The seeds are
    // put into a page-locked buffer of the handle that the kernel reads and writes in place over PCIe (in place, as
    // project() does); the only device work is the launch record's memset and the kernel.
    rc = ensure_pin(h);
    if (rc) return rc;
    char* pb = (char*)h->pin;
    double* px = (double*)pb;
    double* pres = (double*)(pb + sizeof(double) * n * (size_t)count);
    int32_t* pit = (int32_t*)((char*)pres + sizeof(double) * m * (size_t)count);
    uint8_t* pok = (uint8_t*)((char*)pit + sizeof(int32_t) * (size_t)count);
    uint8_t* pcv = pok + (size_t)count;
    memcpy(px, seeds_host, sizeof(double) * n * (size_t)count);
    cudaStream_t st = h->hstream[1];
    ccp_project_args A;
    memset(&A, 0, sizeof A);
    A.seeds = px;
    A.x_out = px;
    A.ok = pok;
    A.conv = converged_host ? pcv : nullptr;
    A.iters = iters_host ? pit : nullptr;
    A.resid = resid_host ? pres : nullptr;
    A.count = count;
    A.stage_seeds = -1;  // no bulk copies out of host memory: the lanes load their seeds directly
    rc = launch_project(h, A, CCP_LAYOUT_AOS, st, false);
    if (rc) return rc;
    CCP_CUDA(cudaStreamSynchronize(st));
    if (x_out_host) memcpy(x_out_host, px, sizeof(double) * n * (size_t)count);
    if (ok_host) memcpy(ok_host, pok, (size_t)count);
    if (converged_host) memcpy(converged_host, pcv, (size_t)count);
    if (iters_host) memcpy(iters_host, pit, sizeof(int32_t) * (size_t)count);
    if (resid_host) memcpy(resid_host, pres, sizeof(double) * m * (size_t)count);
    return CCP_OK;
  }
  // a batch too small to chunk: copy-in, one complete launch and copy-out on one stream, no events
  rc = ensure_stage(h, per * (size_t)count + 1024);
  if (rc) return rc;
  char* base = (char*)h->d_stage;
  double* dx = (double*)base;
  double* dres = (double*)(base + sizeof(double) * n * (size_t)count);
  int32_t* dit = (int32_t*)((char*)dres + sizeof(double) * m * (size_t)count);
  uint8_t* dok = (uint8_t*)((char*)dit + sizeof(int32_t) * (size_t)count);
  uint8_t* dcv = dok + (size_t)count;
  cudaStream_t st = h->hstream[1];
  CCP_CUDA(cudaMemcpyAsync(dx, seeds_host, sizeof(double) * n * count, cudaMemcpyHostToDevice, st));
  ccp_project_args A;
  memset(&A, 0, sizeof A);
  A.seeds = dx;
  A.x_out = dx;
  A.ok = dok;
  A.conv = dcv;
  A.iters = dit;
  A.resid = resid_host ? dres : nullptr;
  A.count = count;
  rc = launch_project(h, A, CCP_LAYOUT_AOS, st, false);
  if (rc) return rc;
  if (x_out_host) CCP_CUDA(cudaMemcpyAsync(x_out_host, dx, sizeof(double) * n * count, cudaMemcpyDeviceToHost, st));
  if (ok_host) CCP_CUDA(cudaMemcpyAsync(ok_host, dok, (size_t)count, cudaMemcpyDeviceToHost, st));
  if (converged_host) CCP_CUDA(cudaMemcpyAsync(converged_host, dcv, (size_t)count, cudaMemcpyDeviceToHost, st));
  if (iters_host) CCP_CUDA(cudaMemcpyAsync(iters_host, dit, sizeof(int32_t) * count, cudaMemcpyDeviceToHost, st));
  if (resid_host) CCP_CUDA(cudaMemcpyAsync(resid_host, dres, sizeof(double) * m * count, cudaMemcpyDeviceToHost, st));
  CCP_CUDA(cudaStreamSynchronize(st));
  return CCP_OK;
}

int ccp_project_batch_host_submit(ccp_handle* h, const double* seeds_host, int64_t count, double* x_out_host,
                                  uint8_t* ok_host, uint8_t* converged_host, int32_t* iters_host, double* resid_host,
                                  int64_t* ticket_out) {
  int rc = check_common(h, seeds_host, count, CCP_LAYOUT_AOS);
  if (rc) return rc;
  if (!ticket_out) return set_err(h, CCP_ERR_INVALID, "%s", "null ticket");
  if (count < 1) return set_err(h, CCP_ERR_INVALID, "%s", "submit needs at least one state");
  if (!h->has_ref) return set_err(h, CCP_ERR_STATE, "%s", "project before ccp_set_reference (setInitialPosition)");
  ccp_host_batch B;
  memset(&B, 0, sizeof B);
  B.seeds_host = seeds_host;
  B.count = count;
  B.x_out_host = x_out_host;
  B.ok_host = ok_host;
  B.converged_host = converged_host;
  B.iters_host = iters_host;
  B.resid_host = resid_host;
  std::lock_guard<std::mutex> host_lock(h->host_mu);
  device_guard g(h->device);
  return host_submit_locked(h, B, ticket_out);
}

int ccp_host_batch_submit(ccp_handle* h, const ccp_host_batch* b, int64_t* ticket_out) {
  if (!h) return CCP_ERR_INVALID;
  if (!b || !ticket_out) return set_err(h, CCP_ERR_INVALID, "%s", "null batch / ticket");
  if (b->count < 1) return set_err(h, CCP_ERR_INVALID, "%s", "submit needs at least one state");
  if ((b->seeds_host != nullptr) == (b->sampler != nullptr))
    return set_err(h, CCP_ERR_INVALID, "%s", "a host batch has either seeds_host or sampler arguments");
  if (b->count > 0x7fff0000LL) return set_err(h, CCP_ERR_INVALID, "%s", "more than 2^31 - 65536 samples in one batch");
  if ((b->compact_host || b->compact_index_host) && b->compact_capacity < 1)
    return set_err(h, CCP_ERR_INVALID, "%s", "compact outputs need compact_capacity >= 1");
  if (b->compact_capacity < 0) return set_err(h, CCP_ERR_INVALID, "%s", "negative compact_capacity");
  if (!h->has_ref) return set_err(h, CCP_ERR_STATE, "%s", "project before ccp_set_reference (setInitialPosition)");
  std::lock_guard<std::mutex> host_lock(h->host_mu);
  device_guard g(h->device);
  return host_submit_locked(h, *b, ticket_out);
}

int ccp_host_batch_wait(ccp_handle* h, int64_t ticket, int64_t* n_ok_out) {
  if (!h) return CCP_ERR_INVALID;
  std::lock_guard<std::mutex> host_lock(h->host_mu);
  if (ticket < 1 || ticket >= h->next_ticket) return set_err(h, CCP_ERR_INVALID, "%s", "unknown ticket");
  device_guard g(h->device);
  for (int i = 0; i < 2; ++i)
    if (h->hcall[i].ticket == ticket) {
      int rc = host_wait_locked(h, h->hcall[i]);
      if (rc) return rc;
      if (n_ok_out) *n_ok_out = h->hcall[i].compact ? h->hcall[i].n_ok : -1;
      return CCP_OK;
    }
  return set_err(h, CCP_ERR_INVALID, "%s", "ticket already retired (its slot was reused by a later submit)");
}

int ccp_project_batch_host_wait(ccp_handle* h, int64_t ticket) {
  if (!h) return CCP_ERR_INVALID;
  std::lock_guard<std::mutex> host_lock(h->host_mu);
  if (ticket < 1 || ticket >= h->next_ticket) return set_err(h, CCP_ERR_INVALID, "%s", "unknown ticket");
  device_guard g(h->device);
  for (int i = 0; i < 2; ++i)
    if (h->hcall[i].ticket == ticket) return host_wait_locked(h, h->hcall[i]);
  return CCP_OK;  // an older ticket: its slot was recycled, which waited for it
}

// Host-buffer forms of the sampler, the geodesic walk and the IK sampler: what a C++ planner that never touches CUDA
// calls (include/closed_chain_motion_planner_b200/ProjectedStateSpace.hpp).  Device scratch comes from the handle's
// stream-ordered pool; synchronous.
int ccp_sample_project_batch_host(ccp_handle* h, const ccp_sampler_args* a, int64_t count, double* x_out_host,
                                  uint8_t* ok_host, int32_t* iters_host, double* compact_host, int64_t* n_ok_host) {
  if (!h) return CCP_ERR_INVALID;
  if (count < 0) return set_err(h, CCP_ERR_INVALID, "%s", "negative count");
  if (compact_host && !n_ok_host) return set_err(h, CCP_ERR_INVALID, "%s", "compact output needs n_ok");
  if (count == 0) {
    if (n_ok_host) *n_ok_host = 0;
    return CCP_OK;
  }
  std::lock_guard<std::mutex> host_lock(h->host_mu);
  device_guard g(h->device);
  for (int i = 0; i < 2; ++i) {  // batches submitted through the streaming projection form complete first
    int wrc = host_wait_locked(h, h->hcall[i]);
    if (wrc) return wrc;
  }
  CCP_NO_OPEN_PIPELINE(h);
  const int n = CCPC_DOF * h->model.n_arms;
  cudaStream_t st = h->hstream[1];
  const size_t xb = sizeof(double) * n * (size_t)count;
  char* base = nullptr;
  const size_t total = (x_out_host ? xb : 0) + (compact_host ? xb : 0) + sizeof(int32_t) * (size_t)count + (size_t)count + 64;
  CCP_CUDA(cudaMallocFromPoolAsync((void**)&base, total, h->pool, st));
  char* p = base;
  double* dx = x_out_host ? (double*)p : nullptr;       p += x_out_host ? xb : 0;
  double* dc = compact_host ? (double*)p : nullptr;     p += compact_host ? xb : 0;
  long long* dn = (long long*)p;                         p += 16;
  int32_t* dit = (int32_t*)p;                            p += sizeof(int32_t) * (size_t)count;
  uint8_t* dok = (uint8_t*)p;
  cudaError_t e = cudaMemsetAsync(dn, 0, 16, st);
  int rc = CCP_OK;
  if (e == cudaSuccess)
    rc = sample_project_impl(h, a, count, CCP_LAYOUT_AOS, dx, ok_host ? dok : nullptr, iters_host ? dit : nullptr, dc,
                             (compact_host || n_ok_host) ? (int64_t*)dn : nullptr, st, false);
  long long nk = 0;
  if (rc == CCP_OK && e == cudaSuccess && (compact_host || n_ok_host)) {
    e = cudaMemcpyAsync(&nk, dn, sizeof nk, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess && n_ok_host) *n_ok_host = nk;
    if (e == cudaSuccess && compact_host && nk > 0)
      e = cudaMemcpyAsync(compact_host, dc, sizeof(double) * n * (size_t)nk, cudaMemcpyDeviceToHost, st);
  }
  if (rc == CCP_OK && e == cudaSuccess && x_out_host) e = cudaMemcpyAsync(x_out_host, dx, xb, cudaMemcpyDeviceToHost, st);
  if (rc == CCP_OK && e == cudaSuccess && ok_host) e = cudaMemcpyAsync(ok_host, dok, (size_t)count, cudaMemcpyDeviceToHost, st);
  if (rc == CCP_OK && e == cudaSuccess && iters_host)
    e = cudaMemcpyAsync(iters_host, dit, sizeof(int32_t) * (size_t)count, cudaMemcpyDeviceToHost, st);
  cudaFreeAsync(base, st);
  cudaError_t e2 = cudaStreamSynchronize(st);
  if (rc) return rc;
  if (e != cudaSuccess || e2 != cudaSuccess)
    return set_err(h, CCP_ERR_CUDA, "ccp_sample_project_batch_host: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
  return CCP_OK;
}

int ccp_geodesic_batch_host(ccp_handle* h, const double* from_host, const double* to_host, int64_t edges, double delta,
                            double lambda, int32_t max_states, double* states_host, int32_t* n_states_host,
                            uint8_t* reached_host, int32_t* iters_host) {
  int rc = check_common(h, from_host, edges, CCP_LAYOUT_AOS);
  if (rc) return rc;
  if (edges > 0 && (!to_host || !states_host || !n_states_host || !reached_host))
    return set_err(h, CCP_ERR_INVALID, "%s", "null geodesic buffer");
  if (max_states < 1) return set_err(h, CCP_ERR_INVALID, "%s", "bad geodesic parameters");
  if (edges == 0) return CCP_OK;
  std::lock_guard<std::mutex> host_lock(h->host_mu);
  device_guard g(h->device);
  const int n = CCPC_DOF * h->model.n_arms;
  cudaStream_t st = h->hstream[1];
  const size_t eb = sizeof(double) * n * (size_t)edges, sb = eb * (size_t)max_states;
  char* base = nullptr;
  CCP_CUDA(cudaMallocFromPoolAsync((void**)&base, 2 * eb + sb + (2 * sizeof(int32_t) + 1) * (size_t)edges + 64, h->pool, st));
  double* dfrom = (double*)base;
  double* dto = (double*)(base + eb);
  double* dst = (double*)(base + 2 * eb);
  int32_t* dns = (int32_t*)(base + 2 * eb + sb);
  int32_t* dit = dns + edges;
  uint8_t* drc = (uint8_t*)(dit + edges);
  cudaError_t e = cudaMemcpyAsync(dfrom, from_host, eb, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dto, to_host, eb, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) rc = ccp_geodesic_batch(h, dfrom, dto, edges, delta, lambda, max_states, dst, dns, drc, dit, st);
  if (rc == CCP_OK && e == cudaSuccess) e = cudaMemcpyAsync(states_host, dst, sb, cudaMemcpyDeviceToHost, st);
  if (rc == CCP_OK && e == cudaSuccess) e = cudaMemcpyAsync(n_states_host, dns, sizeof(int32_t) * (size_t)edges, cudaMemcpyDeviceToHost, st);
  if (rc == CCP_OK && e == cudaSuccess) e = cudaMemcpyAsync(reached_host, drc, (size_t)edges, cudaMemcpyDeviceToHost, st);
  if (rc == CCP_OK && e == cudaSuccess && iters_host)
    e = cudaMemcpyAsync(iters_host, dit, sizeof(int32_t) * (size_t)edges, cudaMemcpyDeviceToHost, st);
  cudaFreeAsync(base, st);
  cudaError_t e2 = cudaStreamSynchronize(st);
  if (rc) return rc;
  if (e != cudaSuccess || e2 != cudaSuccess)
    return set_err(h, CCP_ERR_CUDA, "ccp_geodesic_batch_host: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
  return CCP_OK;
}

int ccp_ik_sample_batch_host(ccp_handle* h, int32_t arm, const double* T_target_host, int64_t n_targets, int32_t restarts,
                             uint64_t rng_seed, double sigma, const double* q_ref_host, const ccp_ik_options* opt,
                             double* q_best_host, uint8_t* ok_host, int32_t* n_success_host) {
  if (!h) return CCP_ERR_INVALID;
  if (n_targets < 0) return set_err(h, CCP_ERR_INVALID, "%s", "negative count");
  if (n_targets > 0 && (!T_target_host || !q_best_host || !ok_host)) return set_err(h, CCP_ERR_INVALID, "%s", "null IK buffer");
  if (n_targets == 0) return CCP_OK;
  std::lock_guard<std::mutex> host_lock(h->host_mu);
  device_guard g(h->device);
  cudaStream_t st = h->hstream[1];
  const size_t tb = sizeof(double) * 12 * (size_t)n_targets, qb = sizeof(double) * 7 * (size_t)n_targets;
  char* base = nullptr;
  CCP_CUDA(cudaMallocFromPoolAsync((void**)&base, tb + 2 * qb + (sizeof(int32_t) + 1) * (size_t)n_targets + 64, h->pool, st));
  double* dT = (double*)base;
  double* dref = (double*)(base + tb);
  double* dq = (double*)(base + tb + qb);
  int32_t* dns = (int32_t*)(base + tb + 2 * qb);
  uint8_t* dok = (uint8_t*)(dns + n_targets);
  cudaError_t e = cudaMemcpyAsync(dT, T_target_host, tb, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && q_ref_host) e = cudaMemcpyAsync(dref, q_ref_host, qb, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(dq, 0, qb, st);
  int rc = CCP_OK;
  if (e == cudaSuccess)
    rc = ccp_ik_sample_batch(h, arm, dT, n_targets, restarts, rng_seed, sigma, q_ref_host ? dref : nullptr, opt, dq, dok, dns, st);
  if (rc == CCP_OK && e == cudaSuccess) e = cudaMemcpyAsync(q_best_host, dq, qb, cudaMemcpyDeviceToHost, st);
  if (rc == CCP_OK && e == cudaSuccess) e = cudaMemcpyAsync(ok_host, dok, (size_t)n_targets, cudaMemcpyDeviceToHost, st);
  if (rc == CCP_OK && e == cudaSuccess && n_success_host)
    e = cudaMemcpyAsync(n_success_host, dns, sizeof(int32_t) * (size_t)n_targets, cudaMemcpyDeviceToHost, st);
  cudaFreeAsync(base, st);
  cudaError_t e2 = cudaStreamSynchronize(st);
  if (rc) return rc;
  if (e != cudaSuccess || e2 != cudaSuccess)
    return set_err(h, CCP_ERR_CUDA, "ccp_ik_sample_batch_host: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
  return CCP_OK;
}

}  // extern "C"

// Host-buffer form of the small evaluation kernels: in -> device kernel -> out.  Small calls (the single-state
// function()/jacobian()/isSatisfied() of the reference's interface) run in place in the handle's page-locked buffer with
// no copy operations; larger ones go through the device stage.  `run(in_dev, out_dev, stream)` enqueues the kernel.
template <class Run>
static int host_eval(ccp_handle* h, const void* in_host, size_t in_bytes, void* const* out_host, const size_t* out_bytes,
                     int n_out, Run run) {
  std::lock_guard<std::mutex> host_lock(h->host_mu);
  device_guard g(h->device);
  size_t total = (in_bytes + 15) & ~(size_t)15;
  size_t off[4];
  for (int i = 0; i < n_out; ++i) {
    off[i] = total;
    total += (out_bytes[i] + 15) & ~(size_t)15;
  }
  cudaStream_t st = h->hstream[0];
  if (total <= CCP_PIN_BYTES && zero_copy_enabled()) {
    int rc = ensure_pin(h);
    if (rc) return rc;
    char* pb = (char*)h->pin;
    memcpy(pb, in_host, in_bytes);
    rc = run(pb, pb, off, st);
    if (rc) return rc;
    CCP_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < n_out; ++i)
      if (out_host[i]) memcpy(out_host[i], pb + off[i], out_bytes[i]);
    return CCP_OK;
  }
  int rc = ensure_stage(h, total);
  if (rc) return rc;
  char* db = (char*)h->d_stage;
  CCP_CUDA(cudaMemcpyAsync(db, in_host, in_bytes, cudaMemcpyHostToDevice, st));
  rc = run(db, db, off, st);
  if (rc) return rc;
  for (int i = 0; i < n_out; ++i)
    if (out_host[i]) CCP_CUDA(cudaMemcpyAsync(out_host[i], db + off[i], out_bytes[i], cudaMemcpyDeviceToHost, st));
  CCP_CUDA(cudaStreamSynchronize(st));
  return CCP_OK;
}

extern "C" {

int ccp_function_batch_host(ccp_handle* h, const double* x_host, int64_t count, double* f_host) {
  int rc = check_common(h, x_host, count, CCP_LAYOUT_AOS);
  if (rc) return rc;
  if (count > 0 && !f_host) return set_err(h, CCP_ERR_INVALID, "%s", "null output");
  if (count == 0) return CCP_OK;
  const int n = CCPC_DOF * h->model.n_arms, m = 2 * (h->model.n_arms - 1);
  void* outs[1] = {f_host};
  const size_t ob[1] = {sizeof(double) * m * (size_t)count};
  return host_eval(h, x_host, sizeof(double) * n * (size_t)count, outs, ob, 1,
                   [&](char* in, char* out, const size_t* off, cudaStream_t st) {
                     return ccp_function_batch(h, (const double*)in, count, CCP_LAYOUT_AOS, (double*)(out + off[0]), st);
                   });
}

int ccp_jacobian_batch_host(ccp_handle* h, const double* x_host, int64_t count, double* J_host) {
  int rc = check_common(h, x_host, count, CCP_LAYOUT_AOS);
  if (rc) return rc;
  if (count > 0 && !J_host) return set_err(h, CCP_ERR_INVALID, "%s", "null output");
  if (count == 0) return CCP_OK;
  const int n = CCPC_DOF * h->model.n_arms, m = 2 * (h->model.n_arms - 1);
  void* outs[1] = {J_host};
  const size_t ob[1] = {sizeof(double) * m * n * (size_t)count};
  return host_eval(h, x_host, sizeof(double) * n * (size_t)count, outs, ob, 1,
                   [&](char* in, char* out, const size_t* off, cudaStream_t st) {
                     return ccp_jacobian_batch(h, (const double*)in, count, CCP_LAYOUT_AOS, (double*)(out + off[0]), st);
                   });
}

int ccp_arm_fk_batch_host(ccp_handle* h, int32_t arm, const double* q_host, int64_t count, double* T_host,
                          double* J_host) {
  int rc = check_common(h, q_host, count, CCP_LAYOUT_AOS);
  if (rc) return rc;
  if (count == 0) return CCP_OK;
  void* outs[2] = {T_host, J_host};
  const size_t ob[2] = {sizeof(double) * 12 * (size_t)count, sizeof(double) * 42 * (size_t)count};
  return host_eval(h, q_host, sizeof(double) * 7 * (size_t)count, outs, ob, 2,
                   [&](char* in, char* out, const size_t* off, cudaStream_t st) {
                     return arm_fk_impl(h, arm, (const double*)in, count, CCP_LAYOUT_AOS,
                                        T_host ? (double*)(out + off[0]) : nullptr,
                                        J_host ? (double*)(out + off[1]) : nullptr, st);
                   });
}

int ccp_host_alloc(void** out, size_t bytes) {
  if (!out) return CCP_ERR_INVALID;
  *out = nullptr;
  if (bytes == 0) return CCP_OK;
  cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
  if (e != cudaSuccess) {
    *out = nullptr;
    return set_err(nullptr, CCP_ERR_CUDA, "cudaHostAlloc: %s", cudaGetErrorString(e));
  }
  return CCP_OK;
}

void ccp_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int ccp_host_register(void* p, size_t bytes) {
  if (!p || bytes == 0) return CCP_ERR_INVALID;
  cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) return set_err(nullptr, CCP_ERR_CUDA, "cudaHostRegister: %s", cudaGetErrorString(e));
  return CCP_OK;
}

int ccp_host_unregister(void* p) {
  if (!p) return CCP_ERR_INVALID;
  cudaError_t e = cudaHostUnregister(p);
  if (e != cudaSuccess) return set_err(nullptr, CCP_ERR_CUDA, "cudaHostUnregister: %s", cudaGetErrorString(e));
  return CCP_OK;
}

int ccp_fp64_peak_probe(ccp_handle* h, int32_t repeats, double* flops_per_s, double* ms_out) {
  if (!h) return CCP_ERR_INVALID;
  device_guard g(h->device);
  if (repeats < 1) repeats = 1;
  const int inner = 4096;
  const int grid = h->sm_count * 8, block = 256;
  double* sink = nullptr;
  CCP_CUDA(cudaMalloc(&sink, sizeof(double)));
  cudaStream_t st = h->hstream[0];
  ccp_dfma_probe_kernel<<<grid, block, 0, st>>>(sink, 64, 1.0000001, 1e-9);  // warm-up
  float best = 1e30f;
  for (int r = 0; r < repeats; ++r) {
    cudaEventRecord(h->ev0, st);
    ccp_dfma_probe_kernel<<<grid, block, 0, st>>>(sink, inner, 1.0000001, 1e-9);
    cudaEventRecord(h->ev1, st);
    cudaError_t e = cudaEventSynchronize(h->ev1);
    if (e != cudaSuccess) {
      cudaFree(sink);
      return set_err(h, CCP_ERR_CUDA, "fp64 probe: %s", cudaGetErrorString(e));
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    if (ms < best) best = ms;
    h->launches++;
  }
  cudaFree(sink);
  const double flops = (double)grid * block * (double)inner * 16.0 * 8.0 * 2.0;
  if (flops_per_s) *flops_per_s = flops / (best * 1e-3);
  if (ms_out) *ms_out = best;
  return CCP_OK;
}

}  // extern "C"
