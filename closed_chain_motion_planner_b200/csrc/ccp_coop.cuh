// ccp_coop.cuh — one Newton trip of a two-arm sample on a PAIR of lanes (lane a owns arm a), as device functions with
// pair-masked shuffles, for kernels whose pairs run independently of one another (ccp_geodesic.cu).  The operations are
// those of ccp_core.h on the same operands in the same order (see ccp_coop.cu, whose projection kernel keeps its own
// warp-uniform copy of the same sequence): results are bit-identical to the thread-per-sample mapping.
#pragma once
#include "ccp_device.cuh"

struct ccp_pair {
  int a;          // 0 / 1: which arm this lane owns
  int lane0;      // the pair's arm-0 lane
  unsigned mask;  // the pair's two lanes
  __device__ __forceinline__ double from0(double v) const { return __shfl_sync(mask, v, lane0); }
  __device__ __forceinline__ double from1(double v) const { return __shfl_sync(mask, v, lane0 | 1); }
  __device__ __forceinline__ double other(double v) const { return __shfl_xor_sync(mask, v, 1); }
};

__device__ __forceinline__ ccp_pair ccp_make_pair() {
  ccp_pair P;
  const int lane = threadIdx.x & 31;
  P.a = lane & 1;
  P.lane0 = lane & ~1;
  P.mask = 3u << P.lane0;
  return P;
}

// forward evaluation: the pair's residual (e2, sv2, d0 on BOTH lanes) and the own arm's gradient start vectors (w, m)
template <int PANDA>
__device__ __forceinline__ void ccp_pair_forward(const ccp_model& M, const ccp_pair& P, const double* x, ccp_sc_local<1>& S,
                                                 double* w, double* m, double& e2, double& sv2, double& d0) {
  const int a = P.a;
  double q[4];
  if (a == 0) {
    const double* s0 = PANDA ? M.arm[1].qrel_scaled : M.arm[1].qrel;
    q[0] = s0[0]; q[1] = s0[1]; q[2] = s0[2]; q[3] = s0[3];
    ccp_fwd_link_quat<PANDA, 0>(M.arm[0], 0, x, q, S);
  } else {
    ccp_fwd_link_quat<PANDA, 0, true>(M.arm[1], 0, x, q, S);
  }
  ccp_fwd_quat_links_1_6<PANDA>(M.arm[a], 0, x, q, S);
  double r[3] = {0.0, 0.0, 0.0};
  if (a == 0) {
    ccp_fwd_down_arm0<PANDA>(M.arm[0], 0, S, r);
    S.rx(0, 6) = 0.0; S.ry(0, 6) = 0.0;  // the EE-0 origin lies on joint 7's own axis (never read into a result)
  }
  double q0[4];
#pragma unroll
  for (int k = 0; k < 3; ++k) r[k] = P.from0(r[k]);
#pragma unroll
  for (int k = 0; k < 4; ++k) q0[k] = P.from0(q[k]);
  double w0[3] = {0.0, 0.0, 0.0}, m0[3] = {0.0, 0.0, 0.0};
  e2 = 0.0; sv2 = 0.0; d0 = 0.0;
  if (a == 1) {
    double v[3], tc[3], qc[4], d[4], e[3];
    ccp_fwd_up_arm<PANDA>(M.arm[1], 0, S, r, v);
    ccp_fwd_pair<2, PANDA>(M.ref[0], v, q, q0, tc, qc, d, e, e2, sv2);
    d0 = d[0];
    w[0] = e[0]; w[1] = e[1]; w[2] = e[2];
    m[0] = d[1]; m[1] = d[2]; m[2] = d[3];
    w0[0] = e[0]; w0[1] = e[1]; w0[2] = e[2];
    ccp_qrot_inv(qc, w0);
    double dq[4];
    ccp_qmul_conj_left(M.ref[0].q0, qc, dq);
    m0[0] = dq[1]; m0[1] = dq[2]; m0[2] = dq[3];
  }
  const double t0 = P.from1(w0[0]), t1 = P.from1(w0[1]), t2 = P.from1(w0[2]);
  const double t3 = P.from1(m0[0]), t4 = P.from1(m0[1]), t5 = P.from1(m0[2]);
  e2 = P.from1(e2);
  sv2 = P.from1(sv2);
  d0 = P.from1(d0);
  if (a == 0) {
    w[0] = t0; w[1] = t1; w[2] = t2;
    m[0] = t3; m[1] = t4; m[2] = t5;
  }
}

// one Newton step of the pair (ccp_jacobian + ccp_newton_step + optional clamp); every lane of the pair calls it
template <int PANDA>
__device__ __forceinline__ void ccp_pair_step(const ccp_model& M, const ccp_pair& P, const ccp_sc_local<1>& S, double* w, double* m,
                                              double e2, double sv2, double d0, double* x) {
  constexpr int H = CCPC_DOF;
  const int a = P.a;
  ccp_jac<2> Jl;
  ccp_jac_arm<PANDA, false>(M.arm[a], 0, 0, S, w, m, Jl);
  if (a == 0) Jl.Ja[0][0][6] = 0.0;  // ARM0, joint 7: no lever arm
  double g00 = ccp_row_dot7(Jl.Ja[0][0], Jl.Ja[0][0]);
  double g10 = ccp_row_dot7(Jl.Ja[0][1], Jl.Ja[0][0]);
  double g11 = ccp_row_dot7(Jl.Ja[0][1], Jl.Ja[0][1]);
  g00 = g00 + P.other(g00);  // arm 0's sum + arm 1's sum (ccp_newton_step: acc + acca; IEEE addition commutes)
  g10 = g10 + P.other(g10);
  g11 = g11 + P.other(g11);
  g00 = CCP_FMA(M.damping, e2, g00);
  g11 = CCP_FMA(M.damping, sv2, g11);
  double rhs[2], y[2];
  ccp_step_rhs(e2, sv2, d0, rhs);
  ccp_solve_2x2(g00, g10, g11, rhs, y);
  if (a == 0) {
#pragma unroll
    for (int i = 0; i < H; ++i) {
      double dx = 0.0;
      dx = CCP_FMA(Jl.Ja[0][0][i], y[0], dx);
      dx = CCP_FMA(Jl.Ja[0][1][i], y[1], dx);
      x[i] = CCP_FMA(-M.step, dx, x[i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < H; ++i) {
      const double dx = CCP_FMA(Jl.Ja[0][1][i], y[1], Jl.Ja[0][0][i] * y[0]);
      x[i] = CCP_FMA(M.step, dx, x[i]);  // the arm-1 rows hold -J
    }
  }
  if (M.clamp) {
#pragma unroll
    for (int i = 0; i < H; ++i) {
      double v = x[i];
      v = (v < M.lb[i]) ? M.lb[i] : v;
      v = (v > M.ub[i]) ? M.ub[i] : v;
      x[i] = v;
    }
  }
}

// jointValid of the whole state: each lane checks its arm, the verdicts are combined across the pair
__device__ __forceinline__ bool ccp_pair_joint_valid(const ccp_model& M, const ccp_pair& P, const double* x) {
  unsigned lo = 0u, hi = 0u;
#pragma unroll
  for (int i = 0; i < CCPC_DOF; ++i) {
    lo |= (unsigned)(x[i] < M.lbm[i]);
    hi |= (unsigned)(x[i] > M.ubm[i]);
  }
  unsigned bad = lo | hi;
  bad |= __shfl_xor_sync(P.mask, bad, 1);
  return bad == 0u;
}

// Euclidean distance of two 14-vectors held half and half by the pair, accumulated in joint order 0..13 exactly as
// ccp_distance does on one thread: lane 0 runs its seven terms, hands the partial sum to lane 1, which runs the other seven
__device__ __forceinline__ double ccp_pair_distance(const ccp_pair& P, const double* u, const double* v) {
  double acc = 0.0;
  if (P.a == 0) {
#pragma unroll
    for (int j = 0; j < CCPC_DOF; ++j) {
      const double d = u[j] - v[j];
      acc = CCP_FMA(d, d, acc);
    }
  }
  acc = P.from0(acc);
  if (P.a == 1) {
#pragma unroll
    for (int j = 0; j < CCPC_DOF; ++j) {
      const double d = u[j] - v[j];
      acc = CCP_FMA(d, d, acc);
    }
  }
  acc = P.from1(acc);
  return sqrt(acc);
}

// ------------------------------------------------------------------------------------------
// Three arms: a QUAD of lanes per sample (ccp_coop.cu, ccp_project_coop3_kernel, keeps its own warp-uniform copy of this
// sequence).  Lane g of a quad: g = 0 arm 0 for residual pair (0,1); g = 1 arm 1; g = 2 arm 2; g = 3 arm 0 AGAIN, for
// pair (0,2) — it repeats arm 0's forward pass in lane 0's instruction stream and runs arm 0's second gradient pass.
// ------------------------------------------------------------------------------------------
struct ccp_quad {
  int g;          // 0..3
  int arm;        // 0, 1, 2, 0
  int pair;       // 0, 0, 1, 1: the residual pair the lane works for
  int lane0;      // the quad's first lane
  unsigned mask;  // the quad's four lanes
  bool is0;       // carries arm 0
  __device__ __forceinline__ double from(double v, int k) const { return __shfl_sync(mask, v, lane0 | k); }
  __device__ __forceinline__ double xor1(double v) const { return __shfl_xor_sync(mask, v, 1); }
};

__device__ __forceinline__ ccp_quad ccp_make_quad() {
  ccp_quad Q;
  const int lane = threadIdx.x & 31;
  Q.g = lane & 3;
  Q.arm = (Q.g == 3) ? 0 : Q.g;
  Q.pair = Q.g >> 1;
  Q.lane0 = lane & ~3;
  Q.mask = 15u << Q.lane0;
  Q.is0 = Q.arm == 0;
  return Q;
}

// forward evaluation: both pairs' residuals (e2, sv2, d0) on EVERY lane, and the own (arm, pair)'s gradient start vectors
template <int PANDA>
__device__ __forceinline__ void ccp_quad_forward(const ccp_model& M, const ccp_quad& Q, const double* x, ccp_sc_local<1>& S,
                                                 double* w, double* m, double* e2, double* sv2, double* d0) {
  const ccp_arm& Arm = M.arm[Q.arm];
  const ccp_pair_ref& Ref = M.ref[Q.pair];
  double q[4];
  q[0] = Arm.qwb[0]; q[1] = Arm.qwb[1]; q[2] = Arm.qwb[2]; q[3] = Arm.qwb[3];
  ccp_fwd_link_quat<PANDA, 0>(Arm, 0, x, q, S);
  ccp_fwd_quat_links_1_6<PANDA>(Arm, 0, x, q, S);
  double r[3] = {0.0, 0.0, 0.0};
  if (Q.is0) {
    ccp_fwd_down_arm0<PANDA>(M.arm[0], 0, S, r);
    S.rx(0, 6) = 0.0; S.ry(0, 6) = 0.0;  // the EE-0 origin lies on joint 7's own axis (never read into a result)
  }
  double q0[4];
#pragma unroll
  for (int k = 0; k < 3; ++k) r[k] = Q.from(r[k], 0);
#pragma unroll
  for (int k = 0; k < 4; ++k) q0[k] = Q.from(q[k], 0);
  double w0[3] = {0.0, 0.0, 0.0}, m0[3] = {0.0, 0.0, 0.0}, e2o = 0.0, sv2o = 0.0, d0o = 0.0;
  if (!Q.is0) {
    double v[3], tc[3], qc[4], d[4], e[3];
    ccp_fwd_up_arm<PANDA>(Arm, 0, S, r, v);
    ccp_fwd_pair<3, PANDA>(Ref, v, q, q0, tc, qc, d, e, e2o, sv2o);
    d0o = d[0];
    w[0] = e[0]; w[1] = e[1]; w[2] = e[2];
    m[0] = d[1]; m[1] = d[2]; m[2] = d[3];
    w0[0] = e[0]; w0[1] = e[1]; w0[2] = e[2];
    ccp_qrot_inv(qc, w0);
    double dq[4];
    ccp_qmul_conj_left(Ref.q0, qc, dq);
    m0[0] = dq[1]; m0[1] = dq[2]; m0[2] = dq[3];
  }
  {
    const double t0 = Q.xor1(w0[0]), t1 = Q.xor1(w0[1]), t2 = Q.xor1(w0[2]);
    const double t3 = Q.xor1(m0[0]), t4 = Q.xor1(m0[1]), t5 = Q.xor1(m0[2]);
    if (Q.is0) {
      w[0] = t0; w[1] = t1; w[2] = t2;
      m[0] = t3; m[1] = t4; m[2] = t5;
    }
  }
  e2[0] = Q.from(e2o, 1); e2[1] = Q.from(e2o, 2);
  sv2[0] = Q.from(sv2o, 1); sv2[1] = Q.from(sv2o, 2);
  d0[0] = Q.from(d0o, 1); d0[1] = Q.from(d0o, 2);
}

// one Newton step of the quad (ccp_jacobian + ccp_newton_step + optional clamp); every lane of the quad calls it
template <int PANDA>
__device__ __forceinline__ void ccp_quad_step(const ccp_model& M, const ccp_quad& Q, const ccp_sc_local<1>& S, double* w, double* m,
                                              const double* e2, const double* sv2, const double* d0, double* x) {
  constexpr int H = CCPC_DOF;
  ccp_jac<2> Jl;
  ccp_jac_arm<PANDA, false>(M.arm[Q.arm], 0, 0, S, w, m, Jl);
  if (Q.is0) Jl.Ja[0][0][6] = 0.0;  // ARM0, joint 7: no lever arm
  double g00 = ccp_row_dot7(Jl.Ja[0][0], Jl.Ja[0][0]);
  double g10 = ccp_row_dot7(Jl.Ja[0][1], Jl.Ja[0][0]);
  double g11 = ccp_row_dot7(Jl.Ja[0][1], Jl.Ja[0][1]);
  g00 = g00 + Q.xor1(g00);  // arm 0's sum + arm p+1's sum (ccp_newton_step: acc + acca)
  g10 = g10 + Q.xor1(g10);
  g11 = g11 + Q.xor1(g11);
  double P0[2][H], P1[2][H];  // arm 0's rows of pair 0 (lane 0) and of pair 1 (lane 3)
#pragma unroll
  for (int rr = 0; rr < 2; ++rr)
#pragma unroll
    for (int i = 0; i < H; ++i) {
      P0[rr][i] = Q.from(Jl.Ja[0][rr][i], 0);
      P1[rr][i] = Q.from(Jl.Ja[0][rr][i], 3);
    }
  double rhs[4];
  {
    double ro[2];
    ccp_step_rhs(Q.pair ? e2[1] : e2[0], Q.pair ? sv2[1] : sv2[0], Q.pair ? d0[1] : d0[0], ro);
    rhs[0] = Q.from(ro[0], 0); rhs[1] = Q.from(ro[1], 0);
    rhs[2] = Q.from(ro[0], 2); rhs[3] = Q.from(ro[1], 2);
  }
  double G[4][4];
  G[0][0] = Q.from(g00, 0); G[1][0] = Q.from(g10, 0); G[1][1] = Q.from(g11, 0);
  G[2][2] = Q.from(g00, 2); G[3][2] = Q.from(g10, 2); G[3][3] = Q.from(g11, 2);
  G[2][0] = ccp_row_dot7(P1[0], P0[0]); G[2][1] = ccp_row_dot7(P1[0], P0[1]);
  G[3][0] = ccp_row_dot7(P1[1], P0[0]); G[3][1] = ccp_row_dot7(P1[1], P0[1]);
  G[0][0] = CCP_FMA(M.damping, e2[0], G[0][0]); G[1][1] = CCP_FMA(M.damping, sv2[0], G[1][1]);
  G[2][2] = CCP_FMA(M.damping, e2[1], G[2][2]); G[3][3] = CCP_FMA(M.damping, sv2[1], G[3][3]);
  double y[4];
  ccp_solve_ldlt<4>(G, rhs, y);
  if (Q.is0) {
#pragma unroll
    for (int i = 0; i < H; ++i) {
      double dx = 0.0;
      dx = CCP_FMA(P0[0][i], y[0], dx);
      dx = CCP_FMA(P0[1][i], y[1], dx);
      dx = CCP_FMA(P1[0][i], y[2], dx);
      dx = CCP_FMA(P1[1][i], y[3], dx);
      x[i] = CCP_FMA(-M.step, dx, x[i]);
    }
  } else {
    const double ya = Q.pair ? y[2] : y[0], yb = Q.pair ? y[3] : y[1];
#pragma unroll
    for (int i = 0; i < H; ++i) {
      const double dx = CCP_FMA(Jl.Ja[0][1][i], yb, Jl.Ja[0][0][i] * ya);
      x[i] = CCP_FMA(M.step, dx, x[i]);  // the arm-a rows hold -J
    }
  }
  if (M.clamp) {
#pragma unroll
    for (int i = 0; i < H; ++i) {
      double v = x[i];
      v = (v < M.lb[i]) ? M.lb[i] : v;
      v = (v > M.ub[i]) ? M.ub[i] : v;
      x[i] = v;
    }
  }
}

__device__ __forceinline__ bool ccp_quad_joint_valid(const ccp_model& M, const ccp_quad& Q, const double* x) {
  unsigned lo = 0u, hi = 0u;
#pragma unroll
  for (int i = 0; i < CCPC_DOF; ++i) {
    lo |= (unsigned)(x[i] < M.lbm[i]);
    hi |= (unsigned)(x[i] > M.ubm[i]);
  }
  unsigned bad = lo | hi;
  bad |= __shfl_xor_sync(Q.mask, bad, 1);
  bad |= __shfl_xor_sync(Q.mask, bad, 2);
  return bad == 0u;
}

// Euclidean distance of two 21-vectors held arm by arm on lanes 0, 1, 2 of the quad, accumulated in joint order 0..20
// exactly as ccp_distance does on one thread: the partial sum travels lane 0 -> 1 -> 2
__device__ __forceinline__ double ccp_quad_distance(const ccp_quad& Q, const double* u, const double* v) {
  double acc = 0.0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (Q.g == a) {
#pragma unroll
      for (int j = 0; j < CCPC_DOF; ++j) {
        const double d = u[j] - v[j];
        acc = CCP_FMA(d, d, acc);
      }
    }
    acc = Q.from(acc, a);
  }
  return sqrt(acc);
}
