// ccp_internal.h — host-side declarations shared by the translation units of libccp.so.
#pragma once
#include <cuda_runtime.h>

#include "ccp_core.h"

struct ccp_project_args;
cudaError_t ccp_launch_project_K2_P0(int sm_count, const ccp_model& M, const ccp_project_args& A, bool soa, cudaStream_t st);
cudaError_t ccp_launch_project_K2_P1(int sm_count, const ccp_model& M, const ccp_project_args& A, bool soa, cudaStream_t st);
cudaError_t ccp_launch_project_K3_P0(int sm_count, const ccp_model& M, const ccp_project_args& A, bool soa, cudaStream_t st);
cudaError_t ccp_launch_project_K3_P1(int sm_count, const ccp_model& M, const ccp_project_args& A, bool soa, cudaStream_t st);
cudaError_t ccp_launch_project_K2_P2(int sm_count, const ccp_model& M, const ccp_project_args& A, bool soa, cudaStream_t st);
cudaError_t ccp_launch_project_K3_P2(int sm_count, const ccp_model& M, const ccp_project_args& A, bool soa, cudaStream_t st);

// cooperative kernels (two lanes per sample for K = 2, four for K = 3; complete launches without pipelining / peers):
// ccp_coop.cu.  Grid of a launch of `count` samples, `spw` samples per warp, 4 warps per block: one warp per SM first,
// then the other three warps of those blocks (one per scheduler) before any SM gets a second block; at most blocks_per_sm
// blocks per SM.  Every sample of a launch with count <= grid * 4 * spw has a static place (the work counter is not touched).
static inline int ccp_coop_grid(int sm_count, long long count, int spw, int blocks_per_sm = 2) {
  const long long need = (count + spw - 1) / spw;
  long long grid = need <= 4LL * sm_count ? (need < sm_count ? need : sm_count) : (need + 3) / 4;
  if (grid > (long long)blocks_per_sm * sm_count) grid = (long long)blocks_per_sm * sm_count;
  return (int)(grid < 1 ? 1 : grid);
}
cudaError_t ccp_launch_project_coop(int sm_count, const ccp_model& M, const ccp_project_args& A, bool soa, cudaStream_t st);

cudaError_t ccp_launch_geodesic(int sm_count, const ccp_model& M, const double* from, const double* to, long long edges,
                                double delta, double lambda, int max_states, double* states, int32_t* n_states,
                                uint8_t* reached, int32_t* total_iters, unsigned long long* counter, long long coop_max,
                                cudaStream_t st);

struct ccp_ik_opt;
cudaError_t ccp_launch_ik(int sm_count, const ccp_model& M, int arm, const double* Tt, const double* qseed, long long count,
                          const ccp_ik_opt& O, double* qout, uint8_t* ok, int32_t* iters, double* err,
                          unsigned long long* counter, cudaStream_t st);
// goal sampling: targets run in chunks of CCP_IK_SAMPLE_CHUNK (bounded scratch); `counters` has one zeroed work counter
// per chunk launch (at most CCP_IK_SAMPLE_MAX_LAUNCHES), `scratch` ccp_ik_sample_scratch_bytes() of device memory
#define CCP_IK_SAMPLE_CHUNK 262144LL
#define CCP_IK_SAMPLE_MAX_LAUNCHES 64
size_t ccp_ik_sample_scratch_bytes(long long n_targets, int restarts);
cudaError_t ccp_launch_ik_sample(int sm_count, const ccp_model& M, int arm, const double* Tt, const double* qref,
                                 long long n_targets, int restarts, unsigned long long rng_seed, double sigma,
                                 const ccp_ik_opt& O, double* qbest, uint8_t* ok, int32_t* n_success, void* scratch,
                                 unsigned long long* counters, int q_stride /* doubles between rows of qref / qbest */,
                                 cudaStream_t st);
// goal sampling of the whole chain: per-arm IK targets from object poses, and the AND of the arms' verdicts
cudaError_t ccp_launch_goal_targets(int sm_count, const ccp_model& M, const double* to7_host, const double* Tobj, long long n,
                                    double* Tt, cudaStream_t st);
cudaError_t ccp_launch_goal_combine(int sm_count, const uint8_t* ok_arm, long long n, int n_arms, uint8_t* ok, cudaStream_t st);
