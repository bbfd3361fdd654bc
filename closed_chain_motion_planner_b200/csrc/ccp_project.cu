// ccp_project.cu — the persistent lane-refill projection kernel (the hot path) and its launchers.
// Compiled once per (K arms, PANDA structured-alpha) pair: -DCCP_TU_K=2|3 -DCCP_TU_PANDA=0|1, so the four
// translation units build in parallel.  Each exports ccp_launch_project_K<k>_P<p>().
//
// Replaces KinematicChainConstraint::project (ConstraintFunction.h:57-82) for a whole batch.
#include <stdlib.h>

#include <type_traits>

#include "ccp_device.cuh"
#include "ccp_internal.h"

// (sin, cos) of the 7K joints of this thread's sample, in shared memory: [slot][thread] so that a warp's
// access to one slot is 32 consecutive doubles (conflict-free), and ~56 registers are freed.
template <int K, int BLOCK>
struct ccp_sc_smem {
  double* base;  // &smem[threadIdx.x]
  __device__ __forceinline__ double& at(int a, int i, int k) const { return base[((a * CCPC_DOF + i) * 4 + k) * BLOCK]; }
  __device__ __forceinline__ double& s(int a, int i) { return at(a, i, 0); }
  __device__ __forceinline__ double& c(int a, int i) { return at(a, i, 1); }
  __device__ __forceinline__ double& rx(int a, int i) { return at(a, i, 2); }
  __device__ __forceinline__ double& ry(int a, int i) { return at(a, i, 3); }
  __device__ __forceinline__ double s(int a, int i) const { return at(a, i, 0); }
  __device__ __forceinline__ double c(int a, int i) const { return at(a, i, 1); }
  __device__ __forceinline__ double rx(int a, int i) const { return at(a, i, 2); }
  __device__ __forceinline__ double ry(int a, int i) const { return at(a, i, 3); }
};

// the Jacobian rows (28 (K-1) doubles) and the state x (7K doubles) can live there too
template <int K, int BLOCK>
struct ccp_jac_smem {
  double* base;
  __device__ __forceinline__ double& a(int p, int r, int i) { return base[(((p * 2 + r) * 2 + 0) * CCPC_DOF + i) * BLOCK]; }
  __device__ __forceinline__ double& z(int p, int r, int i) { return base[(((p * 2 + r) * 2 + 1) * CCPC_DOF + i) * BLOCK]; }
  __device__ __forceinline__ double a(int p, int r, int i) const { return base[(((p * 2 + r) * 2 + 0) * CCPC_DOF + i) * BLOCK]; }
  __device__ __forceinline__ double z(int p, int r, int i) const { return base[(((p * 2 + r) * 2 + 1) * CCPC_DOF + i) * BLOCK]; }
};
template <int BLOCK>
struct ccp_x_smem {
  double* base;
  __device__ __forceinline__ double& operator[](int j) { return base[j * BLOCK]; }
  __device__ __forceinline__ double operator[](int j) const { return base[j * BLOCK]; }
};
// what goes to shared memory: bit 0 = sincos, bit 1 = Jacobian rows, bit 2 = state x
#define CCP_SM_SC 1
#define CCP_SM_J 2
#define CCP_SM_X 4
template <int K, int BLOCK, int SM>
constexpr size_t ccp_proj_smem_bytes() {
  return sizeof(double) * BLOCK *
         (((SM & CCP_SM_SC) ? 4 * CCPC_DOF * K : 0) + ((SM & CCP_SM_J) ? 4 * CCPC_DOF * (K - 1) : 0) +
          ((SM & CCP_SM_X) ? CCPC_DOF * K : 0));
}

// GEN = false: seeds are read from memory (project).  GEN = true: seeds come from the counter-based
// generator and the sampler epilogue (wrap) is compiled in (sample_project).
// PANDA: structured stock-Panda link code (ccp_core.h).  SM: which per-sample arrays are staged in
// shared memory (CCP_SM_* bits) instead of registers.
template <int K, bool PANDA, bool SOA, bool GEN, int BLOCK, int MINB, int SM>
__global__ void __launch_bounds__(BLOCK, MINB)
ccp_project_kernel(const __grid_constant__ ccp_model M, const __grid_constant__ ccp_project_args A) {
  constexpr int n = CCPC_DOF * K, m = 2 * (K - 1);
  extern __shared__ double ccp_smem[];
  double* sm_next = ccp_smem + threadIdx.x;
  typename std::conditional<(SM & CCP_SM_SC) != 0, ccp_sc_smem<K, BLOCK>, ccp_sc_local<K>>::type S;
  if constexpr ((SM & CCP_SM_SC) != 0) {
    S.base = sm_next;
    sm_next += 4 * CCPC_DOF * K * BLOCK;
  }
  typename std::conditional<(SM & CCP_SM_J) != 0, ccp_jac_smem<K, BLOCK>, ccp_jac<K>>::type J;
  if constexpr ((SM & CCP_SM_J) != 0) {
    J.base = sm_next;
    sm_next += 4 * CCPC_DOF * (K - 1) * BLOCK;
  }
  typename std::conditional<(SM & CCP_SM_X) != 0, ccp_x_smem<BLOCK>, double[n]>::type x;
  if constexpr ((SM & CCP_SM_X) != 0) x.base = sm_next;
  int it = 0;
  long long idx = claim_next(A.counter);
  if (!GEN && !SOA && A.ready) wait_chunk_ready(A, idx, idx < A.count);
  if (idx < A.count) {
    if (!GEN) {
      if (!SOA && A.ready) {
#pragma unroll
        for (int j = 0; j < n; ++j) x[j] = __ldcg(A.seeds + idx * n + j);
      } else {
#pragma unroll
        for (int j = 0; j < n; ++j) x[j] = ld_elem<SOA>(A.seeds, idx, j, A.count, n);
      }
    } else {
#pragma unroll
      for (int j = 0; j < n; ++j) x[j] = make_seed<K>(M, A, idx, j);
    }
  }
  while (idx < A.count) {
    ccp_fwd<K> F;
    ccp_forward<K, PANDA>(M, x, S, F);
    const bool cont = ccp_needs_step<K>(M, F.f) && it < M.max_iter;
    if (cont) {
      ++it;
      ccp_jacobian<K, PANDA>(M, S, F, J);
      ccp_newton_step<K>(M, F, J, x);
    } else {
      // ---- epilogue of this sample (ConstraintFunction.h:75-81), then refill the lane ----
      const bool cv = ccp_converged<K>(M, F.f);
      const bool okk = cv && ccp_joint_valid<K>(M, x);
      if (GEN && A.wrap) {
#pragma unroll
        for (int j = 0; j < n; ++j) x[j] = ccp_wrap_pi(x[j]);
      }
      if (A.x_out) {
#pragma unroll
        for (int j = 0; j < n; ++j) st_elem<SOA>(A.x_out, idx, j, A.count, n, x[j]);
      }
      if (A.ok) A.ok[idx] = okk;
      if (A.conv) A.conv[idx] = cv;
      if (A.iters) A.iters[idx] = it;
      if (A.resid) {
#pragma unroll
        for (int k = 0; k < m; ++k) st_elem<SOA>(A.resid, idx, k, A.count, m, F.f[k]);
      }
      if (A.n_ok && okk) {
        const unsigned long long slot = atomicAdd(A.n_ok, 1ULL);
        if (A.compact) {
#pragma unroll
          for (int j = 0; j < n; ++j) A.compact[slot * n + j] = x[j];
        }
      }
      if (!GEN && !SOA && A.done) {
        // streaming: count this sample into its chunk; the last one tells the host the chunk can be copied out
        __threadfence();
        const long long c = idx / A.chunk;
        const long long in_chunk = (A.count - c * A.chunk < A.chunk) ? (A.count - c * A.chunk) : A.chunk;
        if ((long long)atomicAdd(A.done + c, 1u) + 1 == in_chunk) {
          __threadfence_system();
          A.host_done[c] = 1;
        }
      }
      idx = claim_next(A.counter);
      it = 0;
      if (!GEN && !SOA && A.ready) wait_chunk_ready(A, idx, idx < A.count);
      if (idx < A.count) {
        if (!GEN) {
          if (!SOA && A.ready) {
#pragma unroll
            for (int j = 0; j < n; ++j) x[j] = __ldcg(A.seeds + idx * n + j);
          } else {
#pragma unroll
            for (int j = 0; j < n; ++j) x[j] = ld_elem<SOA>(A.seeds, idx, j, A.count, n);
          }
        } else {
#pragma unroll
          for (int j = 0; j < n; ++j) x[j] = make_seed<K>(M, A, idx, j);
        }
      }
    }
  }
}


// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
// projection launch configuration: BLOCK threads, MINB resident blocks per SM (persistent grid), and
// which per-sample arrays live in shared memory.  CCP_PROJ_VARIANT (environment, read once) selects
// among the compiled configurations for tuning; the default is the best one measured on B200.
template <int K, bool PANDA, bool SOA, bool GEN, int BLOCK, int MINB, int SM>
static cudaError_t launch_project_v(int sm_count, const ccp_model& M, const ccp_project_args& A, cudaStream_t st) {
  long long need = (A.count + BLOCK - 1) / BLOCK;
  long long cap = (long long)sm_count * MINB;
  int grid = (int)(need < cap ? need : cap);
  if (grid < 1) grid = 1;
  constexpr size_t smem = ccp_proj_smem_bytes<K, BLOCK, SM>();
  auto kern = ccp_project_kernel<K, PANDA, SOA, GEN, BLOCK, MINB, SM>;
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  kern<<<grid, BLOCK, smem, st>>>(M, A);
  return cudaGetLastError();
}

static int proj_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CCP_PROJ_VARIANT");
    v = e ? atoi(e) : 0;
  }
  return v;
}

template <int K, bool PANDA, bool SOA, bool GEN>
static cudaError_t launch_project_g(int sm_count, const ccp_model& M, const ccp_project_args& A, cudaStream_t st) {
#ifdef CCP_TUNE
  if constexpr (K == 2) {
    switch (proj_variant()) {
      case 1: return launch_project_v<K, PANDA, SOA, GEN, 128, 3, CCP_SM_SC>(sm_count, M, A, st);
      case 2: return launch_project_v<K, PANDA, SOA, GEN, 128, 4, CCP_SM_SC>(sm_count, M, A, st);
      case 3: return launch_project_v<K, PANDA, SOA, GEN, 128, 3, CCP_SM_SC | CCP_SM_J>(sm_count, M, A, st);
      case 4: return launch_project_v<K, PANDA, SOA, GEN, 64, 6, 0>(sm_count, M, A, st);
      case 7: return launch_project_v<K, PANDA, SOA, GEN, 32, 12, 0>(sm_count, M, A, st);
      case 5: return launch_project_v<K, PANDA, SOA, GEN, 256, 2, CCP_SM_SC | CCP_SM_J>(sm_count, M, A, st);
      case 6: return launch_project_v<K, PANDA, SOA, GEN, 128, 2, 0>(sm_count, M, A, st);
      default: break;
    }
  }
#endif
  // measured on B200 (profiles/): 3 resident blocks of 128 threads per SM, everything in registers
  // (168 regs, no spills) beats every shared-memory staging variant for K = 2
  if constexpr (K == 2) return launch_project_v<K, PANDA, SOA, GEN, 128, 3, 0>(sm_count, M, A, st);
  else return launch_project_v<K, PANDA, SOA, GEN, 128, 2, CCP_SM_SC>(sm_count, M, A, st);
}

#define CCP_CAT_(a, b, c, d) a##b##c##d
#define CCP_CAT(a, b, c, d) CCP_CAT_(a, b, c, d)
cudaError_t CCP_CAT(ccp_launch_project_K, CCP_TU_K, _P, CCP_TU_PANDA)(int sm_count, const ccp_model& M,
                                                                     const ccp_project_args& A, bool soa,
                                                                     cudaStream_t st) {
  constexpr int K = CCP_TU_K;
  constexpr bool PANDA = CCP_TU_PANDA != 0;
  const bool gen = A.gen_mode >= 0;
  if (soa) return gen ? launch_project_g<K, PANDA, true, true>(sm_count, M, A, st) : launch_project_g<K, PANDA, true, false>(sm_count, M, A, st);
  return gen ? launch_project_g<K, PANDA, false, true>(sm_count, M, A, st) : launch_project_g<K, PANDA, false, false>(sm_count, M, A, st);
}

