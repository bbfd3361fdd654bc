// ccp_ik.cu — batched pose IK kernels (ccp_ik.h): goal sampling for the planner without TRAC-IK's one-call-at-a-time
// solver (ik_task.cpp:16-49, jy_ConstrainedValidStateSampler.h:63-189).
#include "ccp_device.cuh"
#include "ccp_ik.h"
#include "ccp_internal.h"

#ifndef CCP_IK_BLOCKS_PER_SM
#define CCP_IK_BLOCKS_PER_SM 3  // 168 registers: measured 5 % / 12 % faster than 2 blocks at 212 / 226 registers
#endif

// `arm` is uniform.  The model is a __grid_constant__ parameter, so M.arm[arm] is an indexed read of the constant bank:
// ONE copy of the trip's code (a switch over the arm made three: 102 KB of SASS, and ncu showed 28 % of the sampling
// kernel's stall samples waiting for instructions).
template <int PANDA>
__device__ __forceinline__ bool ik_trip(const ccp_model& M, int arm, const double* T, double* q, const ccp_ik_opt& O, int32_t& it,
                                        bool& conv, double& ep, double& er) {
  return ccp_ik_trip<PANDA>(M.arm[arm], M.lb, M.ub, T, q, O, it, conv, ep, er);
}
// The explicit-seed kernel's loop is short enough for three copies (one per arm, constants at fixed offsets): 2 % faster
// there than the indexed read.
template <int PANDA>
__device__ __forceinline__ bool ik_trip_per_arm(const ccp_model& M, int arm, const double* T, double* q, const ccp_ik_opt& O,
                                                int32_t& it, bool& conv, double& ep, double& er) {
  switch (arm) {
    case 0: return ccp_ik_trip<PANDA>(M.arm[0], M.lb, M.ub, T, q, O, it, conv, ep, er);
    case 1: return ccp_ik_trip<PANDA>(M.arm[1], M.lb, M.ub, T, q, O, it, conv, ep, er);
    default: return ccp_ik_trip<PANDA>(M.arm[2], M.lb, M.ub, T, q, O, it, conv, ep, er);
  }
}

// one out-of-line copy of the Box-Muller draw (log, sincos) instead of seven inlined ones in the refill path
static __device__ __noinline__ double ik_gauss01(unsigned long long seed, unsigned long long sample, unsigned j) {
  return gauss01(seed, sample, j);
}

// Explicit seeds: one lane owns one (target, seed) pair at a time.  Iteration counts spread 0 .. max_iter (a solve
// that fails runs all 200 while the average success takes ~30), so this is a persistent LANE-REFILL loop like the
// projection kernel: one loop pass = one Newton trip of the lane's current solve, and a lane whose solve finished
// writes it out and takes the next pair (first pair static and interleaved over the blocks, the rest from a global
// counter, one warp-aggregated atomic per refill event).  Measured before: 4.8 of 32 lanes active per instruction.
template <int PANDA>
__global__ void __launch_bounds__(128, CCP_IK_BLOCKS_PER_SM)
ccp_ik_kernel(const __grid_constant__ ccp_model M, int arm, const double* __restrict__ Tt, const double* __restrict__ qseed,
              long long count, const __grid_constant__ ccp_ik_opt O, double* __restrict__ qout, uint8_t* __restrict__ ok,
              int32_t* __restrict__ iters, double* __restrict__ err, unsigned long long* __restrict__ counter) {
  double T[12], q[CCPC_DOF];
  int32_t it = 0;
  const long long static_items = (long long)gridDim.x * blockDim.x;
  long long i = ((long long)(threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32 + (threadIdx.x & 31);
  bool load = true;
  for (;;) {
    if (load) {
      if (i >= count) break;
#pragma unroll
      for (int k = 0; k < 12; ++k) T[k] = __ldg(Tt + i * 12 + k);
#pragma unroll
      for (int k = 0; k < CCPC_DOF; ++k) q[k] = __ldg(qseed + i * CCPC_DOF + k);
      it = 0;
      load = false;
    }
    bool conv;
    double ep, er;
    if (ik_trip_per_arm<PANDA>(M, arm, T, q, O, it, conv, ep, er)) {
#pragma unroll
      for (int k = 0; k < CCPC_DOF; ++k) qout[i * CCPC_DOF + k] = q[k];
      if (ok) ok[i] = ccp_ik_accept(M.lb, M.ub, q, O, conv);
      if (iters) iters[i] = it;
      if (err) {
        err[2 * i] = ep;
        err[2 * i + 1] = er;
      }
      i = static_items + claim_next(counter);
      load = true;
    }
  }
}

// Goal sampling: every (target, restart) pair is one work item of the same lane-refill loop (a restart that fails runs
// all max_iter trips, one that succeeds ~30: with a lane group per target and no refill 11 of 32 lanes were active).
// Restart 0 starts from the reference configuration when there is one (the seeded solve of sampleCalibGoal), the others
// from N(nominal, sigma) draws clipped to the limits (getRandomConfig).  A finished restart leaves its selection key
// (and, when it succeeded, its solution) in scratch memory and counts itself done; whichever lane finishes a target's
// LAST restart picks the winner: the seeded solution if it succeeded, else the successful restart nearest to the
// reference (sampleCalibGoal) — or, without a reference, the lowest-numbered successful restart.  Ties go to the lower
// restart number, so the result does not depend on which lane ran what.  Restarts that can no longer win are abandoned
// (see the loop): n_success counts the successful restarts that got to finish.
struct ccp_ik_sample_scratch {
  double* key;        // [n_targets][restarts] selection key, 1e300 = failed
  double* q;          // [n_targets][restarts][7] solutions of the successful restarts
  unsigned* done;     // [n_targets] restarts finished (zeroed before the launch)
  unsigned* decided;  // [n_targets] restarts - r of the lowest-numbered restart that has decided the target (zeroed with `done`)
};

template <int PANDA>
__global__ void __launch_bounds__(128, CCP_IK_BLOCKS_PER_SM)
ccp_ik_sample_kernel(const __grid_constant__ ccp_model M, int arm, const double* __restrict__ Tt,
                     const double* __restrict__ qref, int q_stride, long long n_targets, int restarts,
                     unsigned long long rng_seed, long long first_target, double sigma, const __grid_constant__ ccp_ik_opt O,
                     const __grid_constant__ ccp_ik_sample_scratch W, double* __restrict__ qbest, uint8_t* __restrict__ ok,
                     int32_t* __restrict__ n_success, unsigned long long* __restrict__ counter) {
  double T[12], q[CCPC_DOF], ref[CCPC_DOF];
  int32_t it = 0;
  const long long items = n_targets * restarts;
  const long long static_items = (long long)gridDim.x * blockDim.x;
  long long w = ((long long)(threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32 + (threadIdx.x & 31);
  long long t = 0;
  int r = 0;
  bool load = true;
  for (;;) {
    if (load) {
      if (w >= items) break;
      // restart-major: every target's restart 0 is handed out before any restart 1, ...: in a launch larger than the
      // machine a restart that can no longer win (below) finds its target decided at once
      r = (int)(w / n_targets);
      t = w - (long long)r * n_targets;
#pragma unroll
      for (int k = 0; k < 12; ++k) T[k] = __ldg(Tt + t * 12 + k);
#pragma unroll
      for (int k = 0; k < CCPC_DOF; ++k) ref[k] = qref ? __ldg(qref + t * q_stride + k) : 0.5 * (M.lb[k] + M.ub[k]);
      if (r == 0 && qref) {
#pragma unroll
        for (int k = 0; k < CCPC_DOF; ++k) q[k] = ref[k];
      } else {
#pragma unroll
        for (int k = 0; k < CCPC_DOF; ++k)
          q[k] = ccp_ik_random_joint(M.lb[k], M.ub[k], sigma,
                                     ik_gauss01(rng_seed, (unsigned long long)(first_target + t) * 64ull + (unsigned)r, (unsigned)k));
      }
      it = 0;
      load = false;
    }
    bool conv;
    double ep, er;
    // Restarts that can no longer win are abandoned.  sampleCalibGoal draws random restarts only when the seeded solve
    // failed (jy_ConstrainedValidStateSampler.h:80-101), so with a reference a successful restart 0 decides the target.
    // sampleRandomGoal (:126-136) runs all its draws and keeps the last success; the draws are i.i.d., so any successful
    // one is distributed alike — here the lowest-numbered success is the answer and higher-numbered draws give up once a
    // lower one has succeeded.  W.decided[t] holds restarts - r of the lowest-numbered restart that has decided target t;
    // restart r gives up when that is above its own restarts - r.  (The look is issued before the trip and read after
    // it, off the critical path.)
    const unsigned decided = (r > 0) ? __ldcg(W.decided + t) : 0u;
    bool fin = ik_trip<PANDA>(M, arm, T, q, O, it, conv, ep, er);
    if (decided > (unsigned)(restarts - r)) {
      fin = true;
      conv = false;
    }
    if (fin) {
      const bool okk = ccp_ik_accept(M.lb, M.ub, q, O, conv);
      if (okk && (r == 0 || !qref)) atomicMax(W.decided + t, (unsigned)(restarts - r));
      double dist2 = 0.0;
#pragma unroll
      for (int k = 0; k < CCPC_DOF; ++k) {
        const double d = q[k] - ref[k];
        dist2 = CCP_FMA(d, d, dist2);
      }
      // selection key: seeded success wins outright, then distance to the reference (or the restart number)
      const double key = okk ? ((r == 0 && qref) ? -1.0 : (qref ? dist2 : (double)r)) : 1e300;
      W.key[t * restarts + r] = key;
      if (okk) {
#pragma unroll
        for (int k = 0; k < CCPC_DOF; ++k) W.q[(t * restarts + r) * CCPC_DOF + k] = q[k];
      }
      __threadfence();  // key and solution before the count
      if (atomicAdd(W.done + t, 1u) == (unsigned)restarts - 1u) {
        __threadfence();
        double best = 1e300;
        int who = 0, cnt = 0;
        for (int rr = 0; rr < restarts; ++rr) {
          const double k2 = __ldcg(W.key + t * restarts + rr);
          cnt += k2 < 1e300;
          if (k2 < best) {
            best = k2;
            who = rr;
          }
        }
        const bool any = best < 1e300;
        if (any) {
#pragma unroll
          for (int k = 0; k < CCPC_DOF; ++k) qbest[t * q_stride + k] = __ldcg(W.q + (t * restarts + who) * CCPC_DOF + k);
        }
        ok[t] = any;
        if (n_success) n_success[t] = cnt;
      }
      w = static_items + claim_next(counter);
      load = true;
    }
  }
}

cudaError_t ccp_launch_ik(int sm_count, const ccp_model& M, int arm, const double* Tt, const double* qseed, long long count,
                          const ccp_ik_opt& O, double* qout, uint8_t* ok, int32_t* iters, double* err,
                          unsigned long long* counter, cudaStream_t st) {
  // persistent grid: 2 blocks of 128 per SM at ~250 registers; a small batch is spread one warp's worth per block
  long long need = (count + 31) / 32, cap = (long long)sm_count * CCP_IK_BLOCKS_PER_SM;
  const int grid = (int)(need < cap ? (need < 1 ? 1 : need) : cap);
  if (M.stock) ccp_ik_kernel<2><<<grid, 128, 0, st>>>(M, arm, Tt, qseed, count, O, qout, ok, iters, err, counter);
  else if (M.panda_alpha) ccp_ik_kernel<1><<<grid, 128, 0, st>>>(M, arm, Tt, qseed, count, O, qout, ok, iters, err, counter);
  else ccp_ik_kernel<0><<<grid, 128, 0, st>>>(M, arm, Tt, qseed, count, O, qout, ok, iters, err, counter);
  return cudaGetLastError();
}

size_t ccp_ik_sample_scratch_bytes(long long n_targets, int restarts) {
  const long long chunk = n_targets < CCP_IK_SAMPLE_CHUNK ? n_targets : CCP_IK_SAMPLE_CHUNK;
  return (size_t)chunk * ((size_t)restarts * (8 + 8 * CCPC_DOF) + 8) + 256;
}

cudaError_t ccp_launch_ik_sample(int sm_count, const ccp_model& M, int arm, const double* Tt, const double* qref,
                                 long long n_targets, int restarts, unsigned long long rng_seed, double sigma,
                                 const ccp_ik_opt& O, double* qbest, uint8_t* ok, int32_t* n_success, void* scratch,
                                 unsigned long long* counters, int q_stride, cudaStream_t st) {
  // targets go through in chunks so that the scratch stays bounded; every chunk has its own work counter
  int launch = 0;
  for (long long first = 0; first < n_targets; first += CCP_IK_SAMPLE_CHUNK, ++launch) {
    const long long nt = (n_targets - first < CCP_IK_SAMPLE_CHUNK) ? (n_targets - first) : CCP_IK_SAMPLE_CHUNK;
    ccp_ik_sample_scratch W;
    char* p = (char*)scratch;
    W.key = (double*)p;
    p += sizeof(double) * (size_t)nt * restarts;
    W.q = (double*)p;
    p += sizeof(double) * CCPC_DOF * (size_t)nt * restarts;
    W.done = (unsigned*)p;
    W.decided = W.done + nt;
    cudaError_t e = cudaMemsetAsync(W.done, 0, 2 * sizeof(unsigned) * (size_t)nt, st);
    if (e != cudaSuccess) return e;
    if (launch >= CCP_IK_SAMPLE_MAX_LAUNCHES) return cudaErrorInvalidValue;
    const long long items = nt * restarts;
    long long need = (items + 31) / 32, cap = (long long)sm_count * CCP_IK_BLOCKS_PER_SM;
    const int grid = (int)(need < cap ? (need < 1 ? 1 : need) : cap);
#define CCP_IK_SAMPLE_LAUNCH(P)                                                                                                \
  ccp_ik_sample_kernel<P><<<grid, 128, 0, st>>>(M, arm, Tt + first * 12, qref ? qref + first * q_stride : nullptr, q_stride, nt, \
                                                 restarts, rng_seed, first, sigma, O, W, qbest + first * q_stride, ok + first,     \
                                                 n_success ? n_success + first : nullptr, counters + launch)
    if (M.stock) CCP_IK_SAMPLE_LAUNCH(2);
    else if (M.panda_alpha) CCP_IK_SAMPLE_LAUNCH(1);
    else CCP_IK_SAMPLE_LAUNCH(0);
#undef CCP_IK_SAMPLE_LAUNCH
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// ---- goal sampling for the whole closed chain (jy_ConstrainedValidStateSampler.h:63-189) ----
// IK target of arm a for object pose T_obj: t_b7 = t_wb_a^-1 * T_obj * t_o7_a (IKTask::solve, ik_task.cpp:16-27), 3x4 row-major
struct ccp_goal_frames {
  double to7[CCPC_MAX_ARMS][12];
};
__global__ void __launch_bounds__(128)
ccp_goal_targets_kernel(const __grid_constant__ ccp_model M, const __grid_constant__ ccp_goal_frames F, const double* __restrict__ Tobj,
                        long long n, int n_arms, double* __restrict__ Tt /*[arms][n][12]*/) {
  const long long total = n * n_arms;
  for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < total; w += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(w / n);
    const long long i = w - (long long)a * n;
    double To[12], X[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) To[k] = __ldg(Tobj + i * 12 + k);
    // X = T_obj * t_o7
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        double acc = (c == 3) ? To[4 * r + 3] : 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) acc = CCP_FMA(To[4 * r + k], F.to7[a][4 * k + c], acc);
        X[4 * r + c] = acc;
      }
    }
    // t_wb^-1 * X = [R^T | -R^T p] X
    const ccp_arm& A = M.arm[a];
    double* out = Tt + ((long long)a * n + i) * 12;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) acc = CCP_FMA(A.Rwb[3 * k + r], (c == 3) ? X[4 * k + 3] - A.pwb[k] : X[4 * k + c], acc);
        out[4 * r + c] = acc;
      }
    }
  }
}
__global__ void __launch_bounds__(256)
ccp_goal_combine_kernel(const uint8_t* __restrict__ ok_arm /*[arms][n]*/, long long n, int n_arms, uint8_t* __restrict__ ok) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint8_t all = 1;
    for (int a = 0; a < n_arms; ++a) all &= ok_arm[(long long)a * n + i];
    ok[i] = all;
  }
}

cudaError_t ccp_launch_goal_targets(int sm_count, const ccp_model& M, const double* to7 /*[arms][12] host*/, const double* Tobj,
                                    long long n, double* Tt, cudaStream_t st) {
  ccp_goal_frames F;
  for (int a = 0; a < M.n_arms; ++a)
    for (int k = 0; k < 12; ++k) F.to7[a][k] = to7[a * 12 + k];
  long long need = (n * M.n_arms + 127) / 128, cap = (long long)sm_count * 8;
  const int grid = (int)(need < cap ? (need < 1 ? 1 : need) : cap);
  ccp_goal_targets_kernel<<<grid, 128, 0, st>>>(M, F, Tobj, n, M.n_arms, Tt);
  return cudaGetLastError();
}
cudaError_t ccp_launch_goal_combine(int sm_count, const uint8_t* ok_arm, long long n, int n_arms, uint8_t* ok, cudaStream_t st) {
  long long need = (n + 255) / 256, cap = (long long)sm_count * 8;
  const int grid = (int)(need < cap ? (need < 1 ? 1 : need) : cap);
  ccp_goal_combine_kernel<<<grid, 256, 0, st>>>(ok_arm, n, n_arms, ok);
  return cudaGetLastError();
}
