// ccp_geodesic.cu — batched discreteGeodesic (jy_ProjectedStateSpace.cpp:32-96): the planner's real consumer of
// project (stefanBiPRM.cpp:315; checkMotion at :397,:463).  One thread walks one edge: interpolate by delta toward
// `to` (KinematicChain.h:145-171), project, check deviation / arc length / progress, repeat.  The steps of an edge
// are sequential, so the parallelism is across edges; like the projection kernel this is a persistent lane-refill
// loop whose trip is ONE Newton iteration, so edges of different length and projections of different difficulty
// do not idle the warp.  The state validity check (MoveIt collision, jy_ProjectedStateSpace.cpp:66) is a host
// concern: this is the `interpolate = true` walk; callers validate the returned states lazily.
#include "ccp_coop.cuh"
#include "ccp_device.cuh"
#include "ccp_internal.h"

struct ccp_geodesic_args {
  const double* from;  // AOS [edges][n]
  const double* to;    // AOS [edges][n]
  double* states;      // AOS [edges][max_states][n]; states[e][0] = from
  int32_t* n_states;   // [edges]
  uint8_t* reached;    // [edges] discreteGeodesic's return value (dist <= delta)
  int32_t* total_iters;  // [edges] Newton iterations spent on the edge (may be null)
  unsigned long long* counter;
  long long edges;
  int max_states;
  double delta, lambda;
};

#ifndef CCP_GEO_BLOCKS_K3
#define CCP_GEO_BLOCKS_K3 2  // three arms: 255 registers and 144 B of spill; 3 blocks (168 registers, 680 B) measured 1.3-1.6x slower
#endif
#ifndef CCP_GEO_BLOCKS_K2
#define CCP_GEO_BLOCKS_K2 3
#endif
template <int K, int PANDA>
__global__ void __launch_bounds__(128, K == 2 ? CCP_GEO_BLOCKS_K2 : CCP_GEO_BLOCKS_K3)
ccp_geodesic_kernel(const __grid_constant__ ccp_model M, const __grid_constant__ ccp_geodesic_args A) {
  constexpr int n = CCPC_DOF * K;
  double x[n];
  ccp_sc_local<K> S;
  ccp_jac<K> J;
  ccp_geo_state g;
  g.dist = g.total = g.max = 0.0;
  int it = 0, ns = 0, iters_sum = 0;
  long long e = -1;
  bool need_edge = true;
  // The first edge of every lane is static and interleaved over the blocks (edge group g -> block g % grid, warp
  // g / grid), so a batch smaller than the machine — the planner's usual k x new-vertices edges — spreads one warp
  // per SM before any SM gets a second one; the global counter hands out what lies beyond.
  const long long static_edges = (long long)gridDim.x * (blockDim.x / 32) * 32;
  bool first = true;
  for (;;) {
    if (need_edge) {
      // ---- start the next edge (jy_ProjectedStateSpace.cpp:35-50) ----
      if (first) {
        first = false;
        e = ((long long)(threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32 + (threadIdx.x & 31);
      } else {
        e = static_edges + claim_next(A.counter);
      }
      if (e >= A.edges) break;
      const double* fr = A.from + e * n;
      const double* to = A.to + e * n;
      double* out = A.states + (e * A.max_states) * n;
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < n; ++j) {
        const double a = __ldg(fr + j), b = __ldg(to + j);
        out[j] = a;
        const double d = a - b;
        acc = CCP_FMA(d, d, acc);
        x[j] = a;
      }
      g.dist = sqrt(acc);
      g.total = 0.0;
      g.max = g.dist * A.lambda;
      ns = 1;
      iters_sum = 0;
      if (g.dist <= A.delta) {  // already there
        A.n_states[e] = 1;
        A.reached[e] = 1;
        if (A.total_iters) A.total_iters[e] = 0;
        continue;
      }
      const double t = A.delta / g.dist;
#pragma unroll
      for (int j = 0; j < n; ++j) x[j] = ccp_interpolate_joint(x[j], __ldg(to + j), t);
      it = 0;
      need_edge = false;
    }
    ccp_fwd<K> F;
    ccp_forward<K, PANDA>(M, x, S, F);
    const bool cont = ccp_needs_step<K>(M, F) && it < M.max_iter;
    if (cont) {
      ++it;
      ccp_jacobian<K, PANDA>(M, S, F, J);
      ccp_newton_step<K>(M, F, J, x);
      if (M.clamp) ccp_clamp_to_limits<K>(M, x);
    } else {
      // ---- the projection of this step finished: bookkeeping of the walk ----
      iters_sum += it;
      const bool okk = ccp_converged<K>(M, F) && ccp_joint_valid<K>(M, x);
      const double* to = A.to + e * n;
      double* out = A.states + (e * A.max_states) * n;
      const double* prev = out + (long long)(ns - 1) * n;
      double tov[n], pv[n];
#pragma unroll
      for (int j = 0; j < n; ++j) {
        tov[j] = __ldg(to + j);
        pv[j] = prev[j];
      }
      int code = ccp_geodesic_advance<n>(g, okk, pv, x, tov, A.delta, A.lambda);
      bool overflow = false;
      if (code != 2) {
        if (ns < A.max_states) {
#pragma unroll
          for (int j = 0; j < n; ++j) out[(long long)ns * n + j] = x[j];
          ++ns;
        } else {
          code = 2;  // out of room: report as not reached
          overflow = true;
        }
      }
      if (code == 0) {
        const double t = A.delta / g.dist;
#pragma unroll
        for (int j = 0; j < n; ++j) x[j] = ccp_interpolate_joint(x[j], tov[j], t);
        it = 0;
      } else {
        A.n_states[e] = ns;
        A.reached[e] = (!overflow && g.dist <= A.delta) ? 1 : 0;  // return dist <= tolerance (jy_ProjectedStateSpace.cpp:95)
        if (A.total_iters) A.total_iters[e] = iters_sum;
        need_edge = true;
      }
    }
  }
}

// ---- the same walk with TWO lanes per edge (lane a owns arm a; ccp_coop.cuh) ---------------------------------------------
// A batch of edges is bounded by its longest edge — a serial chain of projections — so for the planner's
// batch sizes (the k = 5 edges of a new vertex, stefanBiPRM.cpp:315; a few thousand at most) what matters is the time of
// one Newton trip of a lone warp, which the cooperative mapping shortens.  Pairs run independently (pair-masked
// shuffles); distances are accumulated across the pair in joint order, so every number equals the one-thread walk's.
template <int PANDA>
__global__ void __launch_bounds__(128, 2)
ccp_geodesic_coop_kernel(const __grid_constant__ ccp_model M, const __grid_constant__ ccp_geodesic_args A) {
  constexpr int n = 2 * CCPC_DOF, H = CCPC_DOF;
  const ccp_pair P = ccp_make_pair();
  const int a = P.a;
  double x[H], prev[H], tov[H];
  ccp_geo_state g;
  g.dist = g.total = g.max = 0.0;
  int it = 0, ns = 0, iters_sum = 0;
  // the first edge of every pair is static and interleaved over the blocks; the counter hands out the rest
  const long long static_edges = (long long)gridDim.x * (blockDim.x / 2);
  long long e = ((long long)(threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 16 + ((threadIdx.x & 31) >> 1);
  bool need_edge = true, first = true;
  for (;;) {
    if (need_edge) {
      if (!first) {
        unsigned long long c = 0;
        if (a == 0) c = atomicAdd(A.counter, 1ULL);
        const unsigned lo = __shfl_sync(P.mask, (unsigned)c, P.lane0), hi = __shfl_sync(P.mask, (unsigned)(c >> 32), P.lane0);
        e = static_edges + (long long)(((unsigned long long)hi << 32) | lo);
      }
      first = false;
      if (e >= A.edges) break;
      const double* fr = A.from + e * n + a * H;
      const double* to = A.to + e * n + a * H;
      double* out = A.states + (e * A.max_states) * n + a * H;
#pragma unroll
      for (int j = 0; j < H; ++j) {
        prev[j] = __ldg(fr + j);
        tov[j] = __ldg(to + j);
        out[j] = prev[j];
        x[j] = prev[j];
      }
      g.dist = ccp_pair_distance(P, prev, tov);
      g.total = 0.0;
      g.max = g.dist * A.lambda;
      ns = 1;
      iters_sum = 0;
      if (g.dist <= A.delta) {  // already there
        if (a == 0) {
          A.n_states[e] = 1;
          A.reached[e] = 1;
          if (A.total_iters) A.total_iters[e] = 0;
        }
        continue;
      }
      const double t = A.delta / g.dist;
#pragma unroll
      for (int j = 0; j < H; ++j) x[j] = ccp_interpolate_joint(x[j], tov[j], t);
      it = 0;
      need_edge = false;
    }
    ccp_sc_local<1> S;
    double w[3], m[3], e2, sv2, d0;
    ccp_pair_forward<PANDA>(M, P, x, S, w, m, e2, sv2, d0);
    const double dw2 = M.tan2_r * (d0 * d0);
    const bool cont = ((e2 > M.tol_p2) || (sv2 > dw2)) && it < M.max_iter;
    if (cont) {
      ++it;
      ccp_pair_step<PANDA>(M, P, S, w, m, e2, sv2, d0, x);
    } else {
      // ---- the projection of this step finished: bookkeeping of the walk (ccp_geodesic_advance, term by term) ----
      iters_sum += it;
      const bool cv = (e2 <= M.tol_p2) && (sv2 < dw2);
      const bool okk = ccp_pair_joint_valid(M, P, x) && cv;
      int code = 2;
      if (okk) {
        const double step = ccp_pair_distance(P, prev, x);
        if (!(step > A.lambda * A.delta)) {
          g.total += step;
          if (!(g.total > g.max)) {
            const double newDist = ccp_pair_distance(P, x, tov);
            if (!(newDist >= g.dist)) {
              g.dist = newDist;
              code = (g.dist >= A.delta) ? 0 : 1;
            }
          }
        }
      }
      bool overflow = false;
      if (code != 2) {
        if (ns < A.max_states) {
          double* out = A.states + (e * A.max_states + ns) * n + a * H;
#pragma unroll
          for (int j = 0; j < H; ++j) {
            out[j] = x[j];
            prev[j] = x[j];
          }
          ++ns;
        } else {
          code = 2;  // out of room: report as not reached
          overflow = true;
        }
      }
      if (code == 0) {
        const double t = A.delta / g.dist;
#pragma unroll
        for (int j = 0; j < H; ++j) x[j] = ccp_interpolate_joint(x[j], tov[j], t);
        it = 0;
      } else {
        if (a == 0) {
          A.n_states[e] = ns;
          A.reached[e] = (!overflow && g.dist <= A.delta) ? 1 : 0;
          if (A.total_iters) A.total_iters[e] = iters_sum;
        }
        need_edge = true;
      }
    }
  }
}

// ---- three arms: FOUR lanes per edge (ccp_coop.cuh, ccp_quad: lanes 0, 1, 2 carry arms 0, 1, 2; lane 3 repeats arm 0) ----
template <int PANDA>
__global__ void __launch_bounds__(128, 2)
ccp_geodesic_coop3_kernel(const __grid_constant__ ccp_model M, const __grid_constant__ ccp_geodesic_args A) {
  constexpr int n = 3 * CCPC_DOF, H = CCPC_DOF;
  const ccp_quad Q = ccp_make_quad();
  const int arm = Q.arm;
  const bool writer = Q.g != 3;  // lane 3's copy of arm 0 is never stored
  double x[H], prev[H], tov[H];
  ccp_geo_state g;
  g.dist = g.total = g.max = 0.0;
  int it = 0, ns = 0, iters_sum = 0;
  const long long static_edges = (long long)gridDim.x * (blockDim.x / 4);
  long long e = ((long long)(threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 8 + ((threadIdx.x & 31) >> 2);
  bool need_edge = true, first = true;
  for (;;) {
    if (need_edge) {
      if (!first) {
        unsigned long long c = 0;
        if (Q.g == 0) c = atomicAdd(A.counter, 1ULL);
        const unsigned lo = __shfl_sync(Q.mask, (unsigned)c, Q.lane0), hi = __shfl_sync(Q.mask, (unsigned)(c >> 32), Q.lane0);
        e = static_edges + (long long)(((unsigned long long)hi << 32) | lo);
      }
      first = false;
      if (e >= A.edges) break;
      const double* fr = A.from + e * n + arm * H;
      const double* to = A.to + e * n + arm * H;
      double* out = A.states + (e * A.max_states) * n + arm * H;
#pragma unroll
      for (int j = 0; j < H; ++j) {
        prev[j] = __ldg(fr + j);
        tov[j] = __ldg(to + j);
        if (writer) out[j] = prev[j];
        x[j] = prev[j];
      }
      g.dist = ccp_quad_distance(Q, prev, tov);
      g.total = 0.0;
      g.max = g.dist * A.lambda;
      ns = 1;
      iters_sum = 0;
      if (g.dist <= A.delta) {  // already there
        if (Q.g == 0) {
          A.n_states[e] = 1;
          A.reached[e] = 1;
          if (A.total_iters) A.total_iters[e] = 0;
        }
        continue;
      }
      const double t = A.delta / g.dist;
#pragma unroll
      for (int j = 0; j < H; ++j) x[j] = ccp_interpolate_joint(x[j], tov[j], t);
      it = 0;
      need_edge = false;
    }
    ccp_sc_local<1> S;
    double w[3], m[3], e2[2], sv2[2], d0[2];
    ccp_quad_forward<PANDA>(M, Q, x, S, w, m, e2, sv2, d0);
    const double dw2a = M.tan2_r * (d0[0] * d0[0]), dw2b = M.tan2_r * (d0[1] * d0[1]);
    const bool cont = ((e2[0] > M.tol_p2) || (sv2[0] > dw2a) || (e2[1] > M.tol_p2) || (sv2[1] > dw2b)) && it < M.max_iter;
    if (cont) {
      ++it;
      ccp_quad_step<PANDA>(M, Q, S, w, m, e2, sv2, d0, x);
    } else {
      // ---- the projection of this step finished: bookkeeping of the walk (ccp_geodesic_advance, term by term) ----
      iters_sum += it;
      const bool cv = (e2[0] <= M.tol_p2) && (sv2[0] < dw2a) && (e2[1] <= M.tol_p2) && (sv2[1] < dw2b);
      const bool okk = ccp_quad_joint_valid(M, Q, x) && cv;
      int code = 2;
      if (okk) {
        const double step = ccp_quad_distance(Q, prev, x);
        if (!(step > A.lambda * A.delta)) {
          g.total += step;
          if (!(g.total > g.max)) {
            const double newDist = ccp_quad_distance(Q, x, tov);
            if (!(newDist >= g.dist)) {
              g.dist = newDist;
              code = (g.dist >= A.delta) ? 0 : 1;
            }
          }
        }
      }
      bool overflow = false;
      if (code != 2) {
        if (ns < A.max_states) {
          double* out = A.states + (e * A.max_states + ns) * n + arm * H;
#pragma unroll
          for (int j = 0; j < H; ++j) {
            if (writer) out[j] = x[j];
            prev[j] = x[j];
          }
          ++ns;
        } else {
          code = 2;  // out of room: report as not reached
          overflow = true;
        }
      }
      if (code == 0) {
        const double t = A.delta / g.dist;
#pragma unroll
        for (int j = 0; j < H; ++j) x[j] = ccp_interpolate_joint(x[j], tov[j], t);
        it = 0;
      } else {
        if (Q.g == 0) {
          A.n_states[e] = ns;
          A.reached[e] = (!overflow && g.dist <= A.delta) ? 1 : 0;
          if (A.total_iters) A.total_iters[e] = iters_sum;
        }
        need_edge = true;
      }
    }
  }
}

cudaError_t ccp_launch_geodesic(int sm_count, const ccp_model& M, const double* from, const double* to, long long edges,
                                double delta, double lambda, int max_states, double* states, int32_t* n_states,
                                uint8_t* reached, int32_t* total_iters, unsigned long long* counter, long long coop_max,
                                cudaStream_t st) {
  ccp_geodesic_args A;
  A.from = from;
  A.to = to;
  A.states = states;
  A.n_states = n_states;
  A.reached = reached;
  A.total_iters = total_iters;
  A.counter = counter;
  A.edges = edges;
  A.max_states = max_states;
  A.delta = delta;
  A.lambda = lambda;
  if (edges <= coop_max) {
    // few edges: several lanes per edge (two arms: 2 lanes, 16 edges per warp; three arms: 4 lanes, 8 edges per warp), one
    // cooperative warp per scheduler before any SM gets a second block
    if (M.n_arms == 2) {
      const int gridc = ccp_coop_grid(sm_count, edges, 16);
      if (M.stock) ccp_geodesic_coop_kernel<2><<<gridc, 128, 0, st>>>(M, A);
      else if (M.panda_alpha) ccp_geodesic_coop_kernel<1><<<gridc, 128, 0, st>>>(M, A);
      else ccp_geodesic_coop_kernel<0><<<gridc, 128, 0, st>>>(M, A);
    } else {
      const int gridc = ccp_coop_grid(sm_count, edges, 8);
      if (M.stock) ccp_geodesic_coop3_kernel<2><<<gridc, 128, 0, st>>>(M, A);
      else if (M.panda_alpha) ccp_geodesic_coop3_kernel<1><<<gridc, 128, 0, st>>>(M, A);
      else ccp_geodesic_coop3_kernel<0><<<gridc, 128, 0, st>>>(M, A);
    }
    return cudaGetLastError();
  }
  long long need = (edges + 31) / 32;  // one warp's worth of edges per block before any block gets more
  long long cap = (long long)sm_count * (M.n_arms == 2 ? CCP_GEO_BLOCKS_K2 : CCP_GEO_BLOCKS_K3);
  int grid = (int)(need < cap ? need : cap);
  if (grid < 1) grid = 1;
  if (M.n_arms == 2) {
    if (M.stock) ccp_geodesic_kernel<2, 2><<<grid, 128, 0, st>>>(M, A);
    else if (M.panda_alpha) ccp_geodesic_kernel<2, 1><<<grid, 128, 0, st>>>(M, A);
    else ccp_geodesic_kernel<2, 0><<<grid, 128, 0, st>>>(M, A);
  } else {
    if (M.stock) ccp_geodesic_kernel<3, 2><<<grid, 128, 0, st>>>(M, A);
    else if (M.panda_alpha) ccp_geodesic_kernel<3, 1><<<grid, 128, 0, st>>>(M, A);
    else ccp_geodesic_kernel<3, 0><<<grid, 128, 0, st>>>(M, A);
  }
  return cudaGetLastError();
}
