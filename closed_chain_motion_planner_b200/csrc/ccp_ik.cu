// ccp_ik.cu — batched pose IK kernels (ccp_ik.h): goal sampling for the planner without TRAC-IK's one-call-at-a-time
// solver (ik_task.cpp:16-49, jy_ConstrainedValidStateSampler.h:63-189).
#include "ccp_device.cuh"
#include "ccp_ik.h"
#include "ccp_internal.h"

// explicit seeds: thread per (target, seed) pair
__global__ void __launch_bounds__(128)
ccp_ik_kernel(const __grid_constant__ ccp_model M, int arm, const double* __restrict__ Tt, const double* __restrict__ qseed,
              long long count, const __grid_constant__ ccp_ik_opt O, double* __restrict__ qout, uint8_t* __restrict__ ok,
              int32_t* __restrict__ iters, double* __restrict__ err) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    double T[12], q[CCPC_DOF], e[2];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = __ldg(Tt + i * 12 + k);
#pragma unroll
    for (int k = 0; k < CCPC_DOF; ++k) q[k] = __ldg(qseed + i * CCPC_DOF + k);
    int32_t it;
    bool okk;
    switch (arm) {  // `arm` is uniform; the switch keeps the model in the constant bank
      case 0: ccp_ik_solve_one(M.arm[0], M.lb, M.ub, T, q, O, &it, &okk, e); break;
      case 1: ccp_ik_solve_one(M.arm[1], M.lb, M.ub, T, q, O, &it, &okk, e); break;
      default: ccp_ik_solve_one(M.arm[2], M.lb, M.ub, T, q, O, &it, &okk, e); break;
    }
#pragma unroll
    for (int k = 0; k < CCPC_DOF; ++k) qout[i * CCPC_DOF + k] = q[k];
    if (ok) ok[i] = okk;
    if (iters) iters[i] = it;
    if (err) {
      err[2 * i] = e[0];
      err[2 * i + 1] = e[1];
    }
  }
}

// Goal sampling: a group of G = 2^g lanes (G >= restarts) owns one target.  Lane 0 of the group starts from the
// reference configuration (the seeded solve of sampleCalibGoal), lanes 1.. from N(nominal, sigma) draws clipped to the
// limits (getRandomConfig); the group then picks, with shuffles, the seeded solution if it succeeded, else the
// successful restart nearest to the reference (sampleCalibGoal) — or, without a reference, the lowest-numbered
// successful restart.
template <int G>
__global__ void __launch_bounds__(128)
ccp_ik_sample_kernel(const __grid_constant__ ccp_model M, int arm, const double* __restrict__ Tt,
                     const double* __restrict__ qref, long long n_targets, int restarts, unsigned long long rng_seed,
                     double sigma, const __grid_constant__ ccp_ik_opt O, double* __restrict__ qbest,
                     uint8_t* __restrict__ ok, int32_t* __restrict__ n_success) {
  const int lane_in_group = threadIdx.x % G;
  const long long groups_per_block = blockDim.x / G;
  for (long long t0 = blockIdx.x * groups_per_block; t0 < n_targets; t0 += (long long)gridDim.x * groups_per_block) {
    const long long t = t0 + threadIdx.x / G;
    const bool have_target = t < n_targets;
    const bool active = have_target && lane_in_group < restarts;
    double T[12], q[CCPC_DOF], ref[CCPC_DOF];
    bool okk = false;
    double dist2 = 0.0;
    if (have_target) {
#pragma unroll
      for (int k = 0; k < CCPC_DOF; ++k) ref[k] = qref ? __ldg(qref + t * CCPC_DOF + k) : 0.5 * (M.lb[k] + M.ub[k]);
    }
    if (active) {
#pragma unroll
      for (int k = 0; k < 12; ++k) T[k] = __ldg(Tt + t * 12 + k);
      if (lane_in_group == 0 && qref) {
#pragma unroll
        for (int k = 0; k < CCPC_DOF; ++k) q[k] = ref[k];
      } else {
#pragma unroll
        for (int k = 0; k < CCPC_DOF; ++k)
          q[k] = ccp_ik_random_joint(M.lb[k], M.ub[k], sigma,
                                     gauss01(rng_seed, (unsigned long long)t * 64ull + (unsigned)lane_in_group, (unsigned)k));
      }
      int32_t it;
      switch (arm) {
        case 0: ccp_ik_solve_one(M.arm[0], M.lb, M.ub, T, q, O, &it, &okk, nullptr); break;
        case 1: ccp_ik_solve_one(M.arm[1], M.lb, M.ub, T, q, O, &it, &okk, nullptr); break;
        default: ccp_ik_solve_one(M.arm[2], M.lb, M.ub, T, q, O, &it, &okk, nullptr); break;
      }
#pragma unroll
      for (int k = 0; k < CCPC_DOF; ++k) {
        const double d = q[k] - ref[k];
        dist2 = CCP_FMA(d, d, dist2);
      }
    }
    // selection key: seeded success wins outright, then distance to the reference (or the restart number)
    double key = okk ? ((lane_in_group == 0 && qref) ? -1.0 : (qref ? dist2 : (double)lane_in_group)) : 1e300;
    int who = lane_in_group;
    int cnt = okk ? 1 : 0;
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) {
      const double k2 = __shfl_xor_sync(0xffffffffu, key, off, G);
      const int w2 = __shfl_xor_sync(0xffffffffu, who, off, G);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, off, G);
      if (k2 < key || (k2 == key && w2 < who)) {
        key = k2;
        who = w2;
      }
    }
    if (have_target) {
      const bool any = key < 1e300;
      if (lane_in_group == who && any) {
#pragma unroll
        for (int k = 0; k < CCPC_DOF; ++k) qbest[t * CCPC_DOF + k] = q[k];
      }
      if (lane_in_group == 0) {
        ok[t] = any;
        if (n_success) n_success[t] = cnt;
      }
    }
  }
}

cudaError_t ccp_launch_ik(int sm_count, const ccp_model& M, int arm, const double* Tt, const double* qseed, long long count,
                          const ccp_ik_opt& O, double* qout, uint8_t* ok, int32_t* iters, double* err, cudaStream_t st) {
  long long need = (count + 127) / 128, cap = (long long)sm_count * 8;
  const int grid = (int)(need < cap ? (need < 1 ? 1 : need) : cap);
  ccp_ik_kernel<<<grid, 128, 0, st>>>(M, arm, Tt, qseed, count, O, qout, ok, iters, err);
  return cudaGetLastError();
}

cudaError_t ccp_launch_ik_sample(int sm_count, const ccp_model& M, int arm, const double* Tt, const double* qref,
                                 long long n_targets, int restarts, unsigned long long rng_seed, double sigma,
                                 const ccp_ik_opt& O, double* qbest, uint8_t* ok, int32_t* n_success, cudaStream_t st) {
  int G = 1;
  while (G < restarts) G <<= 1;
  const long long per_block = 128 / G;
  long long need = (n_targets + per_block - 1) / per_block, cap = (long long)sm_count * 8;
  const int grid = (int)(need < cap ? (need < 1 ? 1 : need) : cap);
#define CCP_IK_CASE(GG) \
  case GG: ccp_ik_sample_kernel<GG><<<grid, 128, 0, st>>>(M, arm, Tt, qref, n_targets, restarts, rng_seed, sigma, O, qbest, ok, n_success); break
  switch (G) {
    CCP_IK_CASE(1);
    CCP_IK_CASE(2);
    CCP_IK_CASE(4);
    CCP_IK_CASE(8);
    CCP_IK_CASE(16);
    CCP_IK_CASE(32);
    default: return cudaErrorInvalidValue;
  }
#undef CCP_IK_CASE
  return cudaGetLastError();
}
