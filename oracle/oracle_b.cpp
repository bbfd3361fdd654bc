// oracle_b.cpp — ORACLE-B: the ENGINE arithmetic (csrc/ccp_core.h) compiled for the host.
//
// TEST INFRASTRUCTURE ONLY.  The product never links this file.  Its one job is the bitwise
// host/device reproducibility check: the CUDA kernels and this file include the same
// __host__ __device__ header, are built with FP contraction off (nvcc --fmad=false,
// g++ -ffp-contract=off) and explicit fma(), so projections must agree BIT FOR BIT
// (tests/test_parity_gpu.py).  Correctness of the algorithm itself is pinned by ORACLE-A
// (oracle_a.c), which shares no code with the engine.
//
// Follows: ConstraintFunction.h:31-40 (setInitialPosition), :43-55 (jointValid), :57-82 (project),
// :84-102 (function), :114-120 (isSatisfied); panda_rbdl.cpp:9-42 (arm FK / Jacobian).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ccp.h"
#include "ccp_core.h"
#include "ccp_ik.h"
#include "ccp_pack.h"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {
// dispatch on (arms, link-code mode: generic / structured alpha / stock) exactly like the CUDA library does
#define OB_DISPATCH(M, CALL)                                  \
  do {                                                        \
    if ((M)->n_arms == 2) {                                   \
      if ((M)->stock) { CALL(2, 2); } else if ((M)->panda_alpha) { CALL(2, 1); } else { CALL(2, 0); } \
    } else {                                                  \
      if ((M)->stock) { CALL(3, 2); } else if ((M)->panda_alpha) { CALL(3, 1); } else { CALL(3, 0); } \
    }                                                         \
  } while (0)

template <int K, int P>
static void function_batch(const ccp_model* M, const double* x, int64_t count, double* f, uint8_t* sat) {
  constexpr int n = 7 * K, m = 2 * (K - 1);
#pragma omp parallel for schedule(static)
  for (int64_t s = 0; s < count; ++s) {
    ccp_fwd<K> F;
    ccp_sc_local<K> S;
    ccp_forward<K, P>(*M, x + s * n, S, F);
    double fv[m];
    ccp_residual<K>(F, fv, nullptr);
    if (f) for (int k = 0; k < m; ++k) f[s * m + k] = fv[k];
    if (sat) sat[s] = ccp_is_satisfied<K>(*M, fv);
  }
}

template <int K, int P>
static void jacobian_batch(const ccp_model* M, const double* x, int64_t count, double* Jout) {
  constexpr int n = 7 * K, m = 2 * (K - 1);
#pragma omp parallel for schedule(static)
  for (int64_t s = 0; s < count; ++s) {
    ccp_fwd<K> F;
    ccp_jac<K> J;
    ccp_sc_local<K> S;
    ccp_forward<K, P>(*M, x + s * n, S, F);
    ccp_jacobian<K, P>(*M, S, F, J);
    ccp_jac_dense<K>(F, J, Jout + s * m * n);
  }
}

template <int K, int P>
static void project_batch(const ccp_model* M, double* x, int64_t count, uint8_t* ok, uint8_t* conv,
                          int32_t* iters, double* resid, int nthreads) {
  constexpr int n = 7 * K, m = 2 * (K - 1);
  (void)nthreads;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads > 0 ? nthreads : 1)
  for (int64_t s = 0; s < count; ++s) {
    double f[m];
    int32_t it;
    bool c, o;
    ccp_project_one<K, P>(*M, x + s * n, f, &it, &c, &o);
    if (ok) ok[s] = o;
    if (conv) conv[s] = c;
    if (iters) iters[s] = it;
    if (resid) for (int k = 0; k < m; ++k) resid[s * m + k] = f[k];
  }
}

// host twin of ccp_geodesic_kernel (same calls, same order)
template <int K, int P>
static void geodesic_batch(const ccp_model* M, const double* from, const double* to, int64_t edges, double delta,
                           double lambda, int max_states, double* states, int32_t* n_states, uint8_t* reached,
                           int32_t* total_iters) {
  constexpr int n = 7 * K, m = 2 * (K - 1);
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t e = 0; e < edges; ++e) {
    const double* fr = from + e * n;
    const double* tov = to + e * n;
    double* out = states + (size_t)e * max_states * n;
    double x[n];
    double acc = 0.0;
    for (int j = 0; j < n; ++j) {
      out[j] = fr[j];
      const double d = fr[j] - tov[j];
      acc = CCP_FMA(d, d, acc);
      x[j] = fr[j];
    }
    ccp_geo_state g;
    g.dist = sqrt(acc);
    g.total = 0.0;
    g.max = g.dist * lambda;
    int ns = 1, iters_sum = 0;
    bool overflow = false;
    if (g.dist <= delta) {
      n_states[e] = 1;
      reached[e] = 1;
      if (total_iters) total_iters[e] = 0;
      continue;
    }
    for (;;) {
      const double t = delta / g.dist;
      for (int j = 0; j < n; ++j) x[j] = ccp_interpolate_joint(x[j], tov[j], t);
      double f[m];
      int32_t it;
      bool cv, okk;
      ccp_project_one<K, P>(*M, x, f, &it, &cv, &okk);
      iters_sum += it;
      const double* prev = out + (size_t)(ns - 1) * n;
      int code = ccp_geodesic_advance<n>(g, okk, prev, x, tov, delta, lambda);
      if (code != 2) {
        if (ns < max_states) {
          for (int j = 0; j < n; ++j) out[(size_t)ns * n + j] = x[j];
          ++ns;
        } else {
          code = 2;
          overflow = true;
        }
      }
      if (code != 0) break;
    }
    n_states[e] = ns;
    reached[e] = (!overflow && g.dist <= delta) ? 1 : 0;
    if (total_iters) total_iters[e] = iters_sum;
  }
}
}  // namespace

extern "C" {

int ob_model_size(void) { return (int)sizeof(ccp_model); }

int ob_create(const ccp_model_desc* d, ccp_model* M) { return ccp_pack_model(d, M); }

int ob_default_desc(int32_t n_arms, const int32_t* arm_index, ccp_model_desc* d) {
  return ccp_fill_default_model(n_arms, arm_index, d);
}

void ob_set_reference(ccp_model* M, const double* q_start) {
#define OB_CALL(K, P) ccp_reference_chain<K, P>(*M, q_start)
  OB_DISPATCH(M, OB_CALL);
#undef OB_CALL
}
int ob_is_panda_alpha(const ccp_model* M) { return M->panda_alpha; }
void ob_get_reference(const ccp_model* M, int pair, double* t0, double* q0) {
  memcpy(t0, M->ref[pair].t0, 3 * sizeof(double));
  memcpy(q0, M->ref[pair].q0, 4 * sizeof(double));
}
void ob_set_tolerance(ccp_model* M, double t1, double t2) { ccp_model_set_tolerance(M, t1, t2); }
void ob_set_options(ccp_model* M, double step, int max_iter, double margin) {
  M->step = step; M->max_iter = max_iter; ccp_model_set_margin(M, margin);
}
void ob_set_modes(ccp_model* M, double damping, int clamp) { M->damping = damping; M->clamp = clamp; }

void ob_function_batch(const ccp_model* M, const double* x, int64_t count, double* f) {
#define OB_CALL(K, P) function_batch<K, P>(M, x, count, f, nullptr)
  OB_DISPATCH(M, OB_CALL);
#undef OB_CALL
}

void ob_jacobian_batch(const ccp_model* M, const double* x, int64_t count, double* J) {
#define OB_CALL(K, P) jacobian_batch<K, P>(M, x, count, J)
  OB_DISPATCH(M, OB_CALL);
#undef OB_CALL
}

// x: AOS count x n, updated in place
void ob_project_batch(const ccp_model* M, double* x, int64_t count, uint8_t* ok, uint8_t* conv,
                      int32_t* iters, double* resid, int nthreads) {
#define OB_CALL(K, P) project_batch<K, P>(M, x, count, ok, conv, iters, resid, nthreads)
  OB_DISPATCH(M, OB_CALL);
#undef OB_CALL
}

void ob_joint_valid_batch(const ccp_model* M, const double* x, int64_t count, uint8_t* out) {
  const int n = 7 * M->n_arms;
  for (int64_t s = 0; s < count; ++s)
    out[s] = M->n_arms == 2 ? ccp_joint_valid<2>(*M, x + s * n) : ccp_joint_valid<3>(*M, x + s * n);
}
void ob_is_satisfied_batch(const ccp_model* M, const double* x, int64_t count, uint8_t* out) {
#define OB_CALL(K, P) function_batch<K, P>(M, x, count, nullptr, out)
  OB_DISPATCH(M, OB_CALL);
#undef OB_CALL
}

void ob_geodesic_batch(const ccp_model* M, const double* from, const double* to, int64_t edges, double delta,
                       double lambda, int max_states, double* states, int32_t* n_states, uint8_t* reached,
                       int32_t* total_iters) {
#define OB_CALL(K, P) geodesic_batch<K, P>(M, from, to, edges, delta, lambda, max_states, states, n_states, reached, total_iters)
  OB_DISPATCH(M, OB_CALL);
#undef OB_CALL
}

// host twin of ccp_ik_kernel (explicit seeds)
void ob_ik_batch(const ccp_model* M, int arm, const double* Tt, const double* qseed, int64_t count, int max_iter,
                 double eps_p, double eps_r, double lambda2, double margin, double* qout, uint8_t* ok, int32_t* iters,
                 double* err) {
  ccp_ik_opt O;
  O.max_iter = max_iter; O.pad = 0; O.eps_p = eps_p; O.eps_r = eps_r; O.lambda2 = lambda2; O.margin = margin;
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t s = 0; s < count; ++s) {
    double q[7], e[2];
    for (int k = 0; k < 7; ++k) q[k] = qseed[7 * s + k];
    int32_t it; bool okk;
    // the link-code mode the CUDA library picks for this model (ccp_launch_ik)
    if (M->stock) ccp_ik_solve_one<2>(M->arm[arm], M->lb, M->ub, Tt + 12 * s, q, O, &it, &okk, e);
    else if (M->panda_alpha) ccp_ik_solve_one<1>(M->arm[arm], M->lb, M->ub, Tt + 12 * s, q, O, &it, &okk, e);
    else ccp_ik_solve_one<0>(M->arm[arm], M->lb, M->ub, Tt + 12 * s, q, O, &it, &okk, e);
    for (int k = 0; k < 7; ++k) qout[7 * s + k] = q[k];
    if (ok) ok[s] = okk;
    if (iters) iters[s] = it;
    if (err) { err[2 * s] = e[0]; err[2 * s + 1] = e[1]; }
  }
}

void ob_arm_fk_batch(const ccp_model* M, int arm, const double* q, int64_t count, double* T, double* J) {
  for (int64_t s = 0; s < count; ++s)
    ccp_arm_fk(M->arm[arm], q + 7 * s, T ? T + 12 * s : nullptr, J ? J + 42 * s : nullptr);
}

void ob_seeds_uniform(const ccp_model* M, uint64_t seed, int64_t first, int64_t count, double* x) {
  const int n = 7 * M->n_arms;
  for (int64_t s = 0; s < count; ++s)
    for (int j = 0; j < n; ++j) x[s * n + j] = ccp_seed_uniform(*M, seed, (uint64_t)(first + s), j);
}

void ob_enforce_bounds(double* x, int64_t n) {
  for (int64_t i = 0; i < n; ++i) x[i] = ccp_wrap_pi(x[i]);
}

void ob_sincos(const double* x, int64_t n, double* s, double* c) {
  for (int64_t i = 0; i < n; ++i) ccp_sincos(x[i], s + i, c + i);
}
void ob_atan2_pos(const double* y, const double* x, int64_t n, double* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = ccp_atan2_pos(y[i], x[i]);
}

}  // extern "C"
