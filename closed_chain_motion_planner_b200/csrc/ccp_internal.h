// ccp_internal.h — host-side declarations shared by the translation units of libccp.so.
#pragma once
#include <cuda_runtime.h>

#include "ccp_core.h"

struct ccp_project_args;
cudaError_t ccp_launch_project_K2_P0(int sm_count, const ccp_model& M, const ccp_project_args& A, bool soa, cudaStream_t st);
cudaError_t ccp_launch_project_K2_P1(int sm_count, const ccp_model& M, const ccp_project_args& A, bool soa, cudaStream_t st);
cudaError_t ccp_launch_project_K3_P0(int sm_count, const ccp_model& M, const ccp_project_args& A, bool soa, cudaStream_t st);
cudaError_t ccp_launch_project_K3_P1(int sm_count, const ccp_model& M, const ccp_project_args& A, bool soa, cudaStream_t st);

cudaError_t ccp_launch_geodesic(int sm_count, const ccp_model& M, const double* from, const double* to, long long edges,
                                double delta, double lambda, int max_states, double* states, int32_t* n_states,
                                uint8_t* reached, int32_t* total_iters, unsigned long long* counter, cudaStream_t st);

struct ccp_ik_opt;
cudaError_t ccp_launch_ik(int sm_count, const ccp_model& M, int arm, const double* Tt, const double* qseed, long long count,
                          const ccp_ik_opt& O, double* qout, uint8_t* ok, int32_t* iters, double* err, cudaStream_t st);
cudaError_t ccp_launch_ik_sample(int sm_count, const ccp_model& M, int arm, const double* Tt, const double* qref,
                                 long long n_targets, int restarts, unsigned long long rng_seed, double sigma,
                                 const ccp_ik_opt& O, double* qbest, uint8_t* ok, int32_t* n_success, cudaStream_t st);
