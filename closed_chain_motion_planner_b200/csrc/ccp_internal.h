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
