// ccp_device.cuh — device-side pieces shared by the translation units of libccp.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ccp_core.h"

// ------------------------------------------------------------------------------------------
// state access (AOS: [count][n], SOA: [n][count])
// ------------------------------------------------------------------------------------------
template <bool SOA>
__device__ __forceinline__ double ld_elem(const double* __restrict__ base, long long idx, int j, long long count,
                                          int n) {
  return SOA ? __ldg(base + (long long)j * count + idx) : __ldg(base + idx * n + j);
}
template <bool SOA>
__device__ __forceinline__ void st_elem(double* __restrict__ base, long long idx, int j, long long count, int n,
                                        double v) {
  if (SOA) base[(long long)j * count + idx] = v;
  else base[idx * n + j] = v;
}

// ------------------------------------------------------------------------------------------
// projection kernel
// ------------------------------------------------------------------------------------------
// Output arrays of one launch, kept in a device table indexed by launch slot: a sample that a pipelined launch
// carried over to its successor still reports into the arrays of the launch it came from.
struct ccp_out_desc {
  double* x_out;
  uint8_t* ok;
  uint8_t* conv;
  int32_t* iters;
  double* resid;
  long long count;  // SOA stride of those arrays
  int wrap;
  unsigned idx_base;            // index of the launch's first seed within its host batch
  // per-BATCH compaction (host path): the ok states of the batch this launch belongs to, whichever launch finishes
  // them.  n_ok == nullptr: the launch's own stream-compaction arguments apply (ccp_project_args::compact / n_ok).
  double* compact;              // AOS [cap][n]
  unsigned long long* n_ok;
  int32_t* compact_idx;         // [cap] seed index (within the batch) of each packed row, or nullptr
  long long compact_cap;
};
#define CCP_NUM_DESC 64
#define CCP_MAX_PEERS 8

// A sample between two trips is (x, iteration count, index, launch slot): what a pipelined launch parks when
// its seed list runs dry and what the next launch adopts.
struct ccp_park_rec {
  unsigned idx;
  int it_slot;  // iterations so far | slot << 16
  double x[CCPC_DOF * CCPC_MAX_ARMS];
};
#define CCP_NO_SAMPLE 0xffffffffu

struct ccp_project_args {
  const double* seeds;  // nullptr when gen_mode >= 0
  double* x_out;
  uint8_t* ok;
  uint8_t* conv;
  int32_t* iters;
  double* resid;
  double* compact;             // AOS [<=count][n] of ok states, or nullptr
  unsigned long long* n_ok;    // appended-to counter for `compact`
  unsigned long long* counter; // work counter (zeroed before launch)
  long long count;
  long long seed_stride;  // SOA: element stride between joints of `seeds` (= count unless the launch is a chunk)
  long long out_stride;   // SOA: same for x_out / resid
  int stage_seeds;  // 1: `seeds` is 16-byte aligned, chunks of it may be bulk-copied to shared memory
  int gen_mode;  // seed kernel only: 0 uniform; 1 uniform-near; 2 gaussian
  int wrap;
  unsigned long long rng_seed;
  long long first_index;
  double distance;
  double near[CCPC_DOF * CCPC_MAX_ARMS];
  // fused all-gather of the converged states (ccp_set_gather_peers): the epilogue also stores an ok state into row
  // peer_row0 + slot of EVERY rank's pool (peer-mapped device memory: P2P stores over NVLink)
  double* peer_pool[CCP_MAX_PEERS];
  double* peer_mc;              // multicast mapping of the pool (one store reaches every rank's pool), or nullptr
  int peer_world;               // 0 = off
  long long peer_row0;          // rank * capacity
  long long peer_cap;           // rows per rank in a pool
  // pipelined launches (ccp_project_batch_pipelined): adopt the samples the previous launch parked, park the
  // samples still iterating when this launch's work runs dry instead of idling the machine on them
  const ccp_park_rec* adopt;    // nullptr = nothing to adopt
  const unsigned* adopt_count;
  ccp_park_rec* park;           // nullptr = run every sample to completion
  unsigned* park_count;
  ccp_out_desc* desc_table;     // [CCP_NUM_DESC]
  unsigned slot;                // this launch's entry of desc_table
  unsigned max_age;             // a sample adopted from a launch this many slots back is not parked again
  unsigned* done;               // [CCP_NUM_DESC] samples finished so far per launch slot, cumulative over the slot's
                                // reuses (never reset); nullptr = do not count.  The host path's D2H stream waits on it.
  // per-batch compaction of the host path (see ccp_out_desc): travels with the launch's descriptor
  double* own_compact;
  unsigned long long* own_n_ok;
  int32_t* own_compact_idx;
  long long own_compact_cap;
  unsigned idx_base;
};

// warp-aggregated claim of the next sample index by the lanes currently finishing
__device__ __forceinline__ long long claim_next(unsigned long long* counter) {
  const unsigned mask = __activemask();
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(mask) - 1;
  unsigned long long base = 0;
  if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(mask));
  base = __shfl_sync(mask, base, leader);
  return (long long)(base + __popc(mask & ((1u << lane) - 1u)));
}

// Box-Muller pair member from two uniforms (mode 2).  log() is CUDA libm: the gaussian stream is
// engine-defined; parity tests read the generated seeds back instead of regenerating them.
__device__ __forceinline__ double gauss01(unsigned long long seed, unsigned long long sample, unsigned j) {
  double u1 = ccp_uniform01(seed, sample, 2u * j + 64u);
  double u2 = ccp_uniform01(seed, sample, 2u * j + 65u);
  u1 = (u1 <= 0.0) ? 0x1.0p-53 : u1;
  double s, c;
  ccp_sincos(6.283185307179586476925 * u2, &s, &c);
  return sqrt(-2.0 * log(u1)) * c;
}

// out-of-line enforceBounds wrap: the epilogue calls it per joint instead of inlining fmod's slow path 7K times
static __device__ __noinline__ double ccp_wrap_pi_call(double v) { return ccp_wrap_pi(v); }

template <int K>
__device__ __forceinline__ double make_seed(const ccp_model& M, const ccp_project_args& A, long long idx, int j) {
  const unsigned long long sample = (unsigned long long)(A.first_index + idx);
  const int i = j % CCPC_DOF;
  if (A.gen_mode == 0) return ccp_seed_uniform(M, A.rng_seed, sample, j);
  if (A.gen_mode == 1) {
    // RealVectorStateSampler::sampleUniformNear: U[max(lb, near-d), min(ub, near+d)]
    double lo = A.near[j] - A.distance, hi = A.near[j] + A.distance;
    lo = (lo < M.lb[i]) ? M.lb[i] : lo;
    hi = (hi > M.ub[i]) ? M.ub[i] : hi;
    return CCP_FMA(ccp_uniform01(A.rng_seed, sample, (unsigned)j), hi - lo, lo);
  }
  // RealVectorStateSampler::sampleGaussian: N(mean, stddev) clipped to the bounds
  double v = CCP_FMA(gauss01(A.rng_seed, sample, (unsigned)j), A.distance, A.near[j]);
  v = (v < M.lb[i]) ? M.lb[i] : v;
  v = (v > M.ub[i]) ? M.ub[i] : v;
  return v;
}

