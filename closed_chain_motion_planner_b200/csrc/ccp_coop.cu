// ccp_coop.cu — the COOPERATIVE projection kernels: two lanes per sample, one arm each (K = 2); four lanes per sample
// for three arms (further down).
//
// The thread-per-sample kernel (ccp_project.cu) is the throughput design: every lane does useful FP64 work and the FP64
// pipe is the limit.  But a lone sample is a chain of ~1100 dependent-ish instructions per Newton iteration on ONE
// lane (3 800 clk per trip measured), and that chain is all there is when the batch is smaller than the machine —
// a planner projecting state by state, a pool refill of a few thousand seeds, the 5 edges of a new roadmap vertex
// (jy_ProjectedStateSpace.cpp:32-96, stefanBiPRM.cpp:315).  Here lane a of a pair owns arm a:
//   * both lanes run their arm's quaternion chain and (sin, cos) in the same instruction stream;
//   * lane 0 pushes the EE-0 origin down arm 0, hands it (and its chain quaternion) to lane 1 by shuffle, lane 1 carries
//     it up arm 1 and forms the pair's residual and the start vectors of both arms' gradient passes;
//   * both lanes run their arm's gradient pass and their share of the Gram sums in one stream again; the shares are
//     added across the pair by shuffle (IEEE addition commutes: both lanes hold the same bits);
//   * the 2x2 solve runs redundantly on both lanes (no extra issue slots), each lane updates its own 7 joints.
// Every arithmetic operation is the one ccp_core.h performs for the thread-per-sample mapping, on the same operands in
// the same order — results are BIT-IDENTICAL (tests/test_coop_gpu.py), so the launcher can pick either kernel by batch
// size.  Cost: ~1.4x the issue slots per sample (the serial down/up passes run at half lane use); gain: a trip takes
// about a third of the cycles.  Used for complete (non-pipelined) launches of small batches only.
#include "ccp_device.cuh"
#include "ccp_internal.h"

#define CCP_COOP_BLOCK 128
#define CCP_FULL 0xffffffffu

__device__ __forceinline__ double shfl_from(double v, int src) { return __shfl_sync(CCP_FULL, v, src); }
__device__ __forceinline__ double shfl_xor1(double v) { return __shfl_xor_sync(CCP_FULL, v, 1); }

template <int PANDA, bool SOA>
__global__ void __launch_bounds__(CCP_COOP_BLOCK, 2)
ccp_project_coop_kernel(const __grid_constant__ ccp_model M, const __grid_constant__ ccp_project_args A) {
  constexpr int n = 2 * CCPC_DOF, H = CCPC_DOF;
  const int lane = threadIdx.x & 31, a = lane & 1, lane0 = lane & ~1;
  const ccp_arm& Arm = M.arm[a];
  // sample numbers: the first one of every pair is static and interleaved over the blocks (warp w of block b takes
  // samples (w * grid + b) * 16 ..), so a small batch spreads one warp per SM; the rest come from the work counter
  const unsigned static_samples = gridDim.x * (CCP_COOP_BLOCK / 2);
  unsigned u = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 16u + (unsigned)(lane >> 1);
  const unsigned total = (unsigned)A.count;
  double x[H];
  unsigned idx = CCP_NO_SAMPLE;
  int it = 0;
  bool fresh = true;
  for (;;) {
    if (fresh) {
      // ---- (re)fill: every lane of the warp passes here together ----
      u = __shfl_sync(CCP_FULL, u, lane0);
      if (idx == CCP_NO_SAMPLE && u < total) {
        idx = u;
        it = 0;
#pragma unroll
        for (int j = 0; j < H; ++j) x[j] = ld_elem<SOA>(A.seeds, idx, a * H + j, A.seed_stride, n);
      } else if (idx == CCP_NO_SAMPLE) {
#pragma unroll
        for (int j = 0; j < H; ++j) x[j] = 0.0;  // an idle pair iterates on a harmless dummy until the warp is done
      }
      fresh = false;
      if (__all_sync(CCP_FULL, idx == CCP_NO_SAMPLE)) break;
    }
    const bool live = idx != CCP_NO_SAMPLE;
    // ---- forward: own chain quaternion and (sin, cos); link 0 differs between the arms (ccp_forward) ----
    ccp_sc_local<1> S;
    double q[4];
    if (a == 0) {
      const double* s0 = PANDA ? M.arm[1].qrel_scaled : M.arm[1].qrel;
      q[0] = s0[0]; q[1] = s0[1]; q[2] = s0[2]; q[3] = s0[3];
      ccp_fwd_link_quat<PANDA, 0>(M.arm[0], 0, x, q, S);
    } else {
      ccp_fwd_link_quat<PANDA, 0, true>(M.arm[1], 0, x, q, S);
    }
    ccp_fwd_quat_links_1_6<PANDA>(Arm, 0, x, q, S);
    double r[3] = {0.0, 0.0, 0.0};
    if (a == 0) {
      ccp_fwd_down_arm0<PANDA>(M.arm[0], 0, S, r);
      S.rx(0, 6) = 0.0; S.ry(0, 6) = 0.0;  // the EE-0 origin lies on joint 7's own axis (never read into a result)
    }
    double q0[4];
#pragma unroll
    for (int k = 0; k < 3; ++k) r[k] = shfl_from(r[k], lane0);
#pragma unroll
    for (int k = 0; k < 4; ++k) q0[k] = shfl_from(q[k], lane0);
    // lane 1: the pair's residual, and the start vectors of both gradient passes
    double w[3], m[3], w0[3] = {0.0, 0.0, 0.0}, m0[3] = {0.0, 0.0, 0.0}, e2 = 0.0, sv2 = 0.0, d0 = 0.0;
    if (a == 1) {
      double v[3], tc[3], qc[4], d[4], e[3];
      ccp_fwd_up_arm<PANDA>(M.arm[1], 0, S, r, v);
      ccp_fwd_pair<2, PANDA>(M.ref[0], v, q, q0, tc, qc, d, e, e2, sv2);
      d0 = d[0];
      w[0] = e[0]; w[1] = e[1]; w[2] = e[2];
      m[0] = d[1]; m[1] = d[2]; m[2] = d[3];
      // arm 0 (ccp_jacobian): e rotated into EE_0's frame; vec d likewise = vec(conj(q_ref) q_c)
      w0[0] = e[0]; w0[1] = e[1]; w0[2] = e[2];
      ccp_qrot_inv(qc, w0);
      double dq[4];
      ccp_qmul_conj_left(M.ref[0].q0, qc, dq);
      m0[0] = dq[1]; m0[1] = dq[2]; m0[2] = dq[3];
    }
    {
      const int lane1 = lane | 1;
      const double t0 = shfl_from(w0[0], lane1), t1 = shfl_from(w0[1], lane1), t2 = shfl_from(w0[2], lane1);
      const double t3 = shfl_from(m0[0], lane1), t4 = shfl_from(m0[1], lane1), t5 = shfl_from(m0[2], lane1);
      e2 = shfl_from(e2, lane1);
      sv2 = shfl_from(sv2, lane1);
      d0 = shfl_from(d0, lane1);
      if (a == 0) {
        w[0] = t0; w[1] = t1; w[2] = t2;
        m[0] = t3; m[1] = t4; m[2] = t5;
      }
    }
    // loop test / success test of project() on the pair's residual (ccp_needs_step / ccp_converged), both lanes alike
    const double dw2 = M.tan2_r * (d0 * d0);
    const bool needs = (e2 > M.tol_p2) || (sv2 > dw2);
    const bool cont = live && needs && it < M.max_iter;
    // ---- gradient pass of the own arm and the own share of the Gram sums ----
    ccp_jac<2> Jl;
    double g00 = 0.0, g10 = 0.0, g11 = 0.0;
    if (cont) {
      ccp_jac_arm<PANDA, false>(Arm, 0, 0, S, w, m, Jl);
      if (a == 0) Jl.Ja[0][0][6] = 0.0;  // ARM0, joint 7: no lever arm
      g00 = ccp_row_dot7(Jl.Ja[0][0], Jl.Ja[0][0]);
      g10 = ccp_row_dot7(Jl.Ja[0][1], Jl.Ja[0][0]);
      g11 = ccp_row_dot7(Jl.Ja[0][1], Jl.Ja[0][1]);
    }
    g00 = g00 + shfl_xor1(g00);  // arm 0's sum + arm 1's sum (ccp_newton_step: acc + acca)
    g10 = g10 + shfl_xor1(g10);
    g11 = g11 + shfl_xor1(g11);
    // jointValid of the own arm when the sample is finishing; the partner's verdict by shuffle
    unsigned bad = 0u;
    if (live && !cont) {
      unsigned lo = 0u, hi = 0u;
#pragma unroll
      for (int i = 0; i < H; ++i) {
        lo |= (unsigned)(x[i] < M.lbm[i]);
        hi |= (unsigned)(x[i] > M.ubm[i]);
      }
      bad = lo | hi;
    }
    bad |= __shfl_xor_sync(CCP_FULL, bad, 1);
    unsigned long long slot = ~0ULL;
    bool packed = false;
    if (cont) {
      ++it;
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
    } else if (live) {
      // ---- the sample is finished (ConstraintFunction.h:75-81) ----
      const bool cv = (e2 <= M.tol_p2) && (sv2 < dw2);
      const bool okk = cv && bad == 0u;
      if (A.wrap) {
#pragma unroll
        for (int j = 0; j < H; ++j) x[j] = ccp_wrap_pi_call(x[j]);
      }
      if (A.x_out) {
#pragma unroll
        for (int j = 0; j < H; ++j) st_elem<SOA>(A.x_out, idx, a * H + j, A.out_stride, n, x[j]);
      }
      if (a == 0) {
        if (A.ok) A.ok[idx] = okk;
        if (A.conv) A.conv[idx] = cv;
        if (A.iters) A.iters[idx] = it;
        if (A.resid) {
          ccp_fwd<2> F;
          F.e2[0] = e2; F.sv2[0] = sv2; F.d[0][0] = d0;
          double fv[2];
          ccp_residual<2>(F, fv, nullptr);
          st_elem<SOA>(A.resid, idx, 0, A.out_stride, 2, fv[0]);
          st_elem<SOA>(A.resid, idx, 1, A.out_stride, 2, fv[1]);
        }
        if (A.n_ok && okk) slot = atomicAdd(A.n_ok, 1ULL);
        // the pair's next sample (no counter: every sample had its static place, nothing is left)
        u = A.counter ? static_samples + atomicAdd((unsigned*)A.counter, 1u) : total;
      }
      packed = A.n_ok && okk && A.compact;
      idx = CCP_NO_SAMPLE;
      fresh = true;
    }
    // the compacted row of a finishing ok sample: lane 0 took the slot, each lane stores its own arm's joints
    if (__any_sync(CCP_FULL, packed)) {
      const unsigned lo = __shfl_sync(CCP_FULL, (unsigned)slot, lane0);
      const unsigned hi = __shfl_sync(CCP_FULL, (unsigned)(slot >> 32), lane0);
      if (packed) {
        const unsigned long long s2 = ((unsigned long long)hi << 32) | lo;
#pragma unroll
        for (int j = 0; j < H; ++j) A.compact[s2 * n + a * H + j] = x[j];
      }
    }
    fresh = __any_sync(CCP_FULL, fresh);  // a refill anywhere in the warp: everyone passes the refill block together
  }
}

// ------------------------------------------------------------------------------------------
// K = 3 (21 DoF): FOUR lanes per sample.  Lane g of a group: g = 0 arm 0, g = 1 arm 1, g = 2 arm 2, g = 3 arm 0 AGAIN.
// The thread-per-sample code runs arm 0's gradient pass twice (once for each residual pair, ccp_jacobian); here lane 0
// runs it for pair (0,1) and lane 3 for pair (0,2) — lane 3 repeats arm 0's forward pass in lane 0's instruction stream,
// which costs no issue slot.  Lanes 1 and 2 carry the EE-0 origin up their arms and form their pair's residual in ONE
// stream (same role), so a trip is about as long as the two-arm cooperative trip, while the thread-per-sample trip for
// three arms is 1.6x the two-arm one.  The 4x4 Gram matrix: same-pair blocks are lane sums across xor 1 (arm 0's share +
// arm p+1's share, as ccp_newton_step adds them), the cross-pair block needs both of arm 0's row pairs — every lane
// fetches them (from lanes 0 and 3) and runs the four dot products, the L D L^T solve and, on the arm-0 lanes, the
// update chain over both pairs in ccp_newton_step's order.  Bit-identical to the thread-per-sample kernel.
template <int PANDA, bool SOA>
__global__ void __launch_bounds__(CCP_COOP_BLOCK, 2)
ccp_project_coop3_kernel(const __grid_constant__ ccp_model M, const __grid_constant__ ccp_project_args A) {
  constexpr int n = 3 * CCPC_DOF, H = CCPC_DOF;
  const int lane = threadIdx.x & 31, g = lane & 3, lane0 = lane & ~3;
  const int arm = (g == 3) ? 0 : g;  // the arm this lane carries
  const int pair = g >> 1;           // the residual pair this lane works for: (0,1) on lanes 0,1; (0,2) on lanes 2,3
  const bool is0 = arm == 0;
  const ccp_arm& Arm = M.arm[arm];
  const ccp_pair_ref& Ref = M.ref[pair];
  const unsigned static_samples = gridDim.x * (CCP_COOP_BLOCK / 4);
  unsigned u = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 8u + (unsigned)(lane >> 2);
  const unsigned total = (unsigned)A.count;
  double x[H];
  unsigned idx = CCP_NO_SAMPLE;
  int it = 0;
  bool fresh = true;
  for (;;) {
    if (fresh) {
      u = __shfl_sync(CCP_FULL, u, lane0);
      if (idx == CCP_NO_SAMPLE && u < total) {
        idx = u;
        it = 0;
#pragma unroll
        for (int j = 0; j < H; ++j) x[j] = ld_elem<SOA>(A.seeds, idx, arm * H + j, A.seed_stride, n);
      } else if (idx == CCP_NO_SAMPLE) {
#pragma unroll
        for (int j = 0; j < H; ++j) x[j] = 0.0;
      }
      fresh = false;
      if (__all_sync(CCP_FULL, idx == CCP_NO_SAMPLE)) break;
    }
    const bool live = idx != CCP_NO_SAMPLE;
    // ---- forward: own chain quaternion from the arm's base rotation (ccp_forward, K != 2), own (sin, cos) ----
    ccp_sc_local<1> S;
    double q[4];
    q[0] = Arm.qwb[0]; q[1] = Arm.qwb[1]; q[2] = Arm.qwb[2]; q[3] = Arm.qwb[3];
    ccp_fwd_link_quat<PANDA, 0>(Arm, 0, x, q, S);
    ccp_fwd_quat_links_1_6<PANDA>(Arm, 0, x, q, S);
    double r[3] = {0.0, 0.0, 0.0};
    if (is0) {
      ccp_fwd_down_arm0<PANDA>(M.arm[0], 0, S, r);
      S.rx(0, 6) = 0.0; S.ry(0, 6) = 0.0;  // the EE-0 origin lies on joint 7's own axis (never read into a result)
    }
    double q0[4];
#pragma unroll
    for (int k = 0; k < 3; ++k) r[k] = shfl_from(r[k], lane0);
#pragma unroll
    for (int k = 0; k < 4; ++k) q0[k] = shfl_from(q[k], lane0);
    // lanes 1, 2: their pair's residual and the start vectors of both of its gradient passes
    double w[3], m[3], w0[3] = {0.0, 0.0, 0.0}, m0[3] = {0.0, 0.0, 0.0}, e2o = 0.0, sv2o = 0.0, d0o = 0.0;
    if (!is0) {
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
      // arm 0's start vectors go to the pair's arm-0 lane (0 <- 1, 3 <- 2)
      const double t0 = shfl_xor1(w0[0]), t1 = shfl_xor1(w0[1]), t2 = shfl_xor1(w0[2]);
      const double t3 = shfl_xor1(m0[0]), t4 = shfl_xor1(m0[1]), t5 = shfl_xor1(m0[2]);
      if (is0) {
        w[0] = t0; w[1] = t1; w[2] = t2;
        m[0] = t3; m[1] = t4; m[2] = t5;
      }
    }
    // both pairs' residuals on every lane
    double e2[2], sv2[2], d0[2];
    e2[0] = shfl_from(e2o, lane0 | 1); e2[1] = shfl_from(e2o, lane0 | 2);
    sv2[0] = shfl_from(sv2o, lane0 | 1); sv2[1] = shfl_from(sv2o, lane0 | 2);
    d0[0] = shfl_from(d0o, lane0 | 1); d0[1] = shfl_from(d0o, lane0 | 2);
    const double dw2a = M.tan2_r * (d0[0] * d0[0]), dw2b = M.tan2_r * (d0[1] * d0[1]);
    const bool needs = (e2[0] > M.tol_p2) || (sv2[0] > dw2a) || (e2[1] > M.tol_p2) || (sv2[1] > dw2b);
    const bool cont = live && needs && it < M.max_iter;
    // ---- gradient pass of the own (arm, pair) and the own share of the pair's Gram block ----
    ccp_jac<2> Jl;
#pragma unroll
    for (int i = 0; i < H; ++i) { Jl.Ja[0][0][i] = 0.0; Jl.Ja[0][1][i] = 0.0; }
    double g00 = 0.0, g10 = 0.0, g11 = 0.0;
    if (cont) {
      ccp_jac_arm<PANDA, false>(Arm, 0, 0, S, w, m, Jl);
      if (is0) Jl.Ja[0][0][6] = 0.0;  // ARM0, joint 7: no lever arm
      g00 = ccp_row_dot7(Jl.Ja[0][0], Jl.Ja[0][0]);
      g10 = ccp_row_dot7(Jl.Ja[0][1], Jl.Ja[0][0]);
      g11 = ccp_row_dot7(Jl.Ja[0][1], Jl.Ja[0][1]);
    }
    g00 = g00 + shfl_xor1(g00);  // arm 0's sum + arm p+1's sum (ccp_newton_step: acc + acca)
    g10 = g10 + shfl_xor1(g10);
    g11 = g11 + shfl_xor1(g11);
    // arm 0's rows of both pairs, on every lane: the cross-pair Gram block and arm 0's update need both
    double P0[2][H], P1[2][H];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
      for (int i = 0; i < H; ++i) {
        P0[rr][i] = shfl_from(Jl.Ja[0][rr][i], lane0);
        P1[rr][i] = shfl_from(Jl.Ja[0][rr][i], lane0 | 3);
      }
    // jointValid of the own arm when the sample is finishing; the group's verdict by shuffle
    unsigned bad = 0u;
    if (live && !cont) {
      unsigned lo = 0u, hi = 0u;
#pragma unroll
      for (int i = 0; i < H; ++i) {
        lo |= (unsigned)(x[i] < M.lbm[i]);
        hi |= (unsigned)(x[i] > M.ubm[i]);
      }
      bad = lo | hi;
    }
    bad |= __shfl_xor_sync(CCP_FULL, bad, 1);
    bad |= __shfl_xor_sync(CCP_FULL, bad, 2);
    // right-hand side of the own pair; both pairs' on every lane
    double rhs[4];
    {
      double ro[2];
      ccp_step_rhs(pair ? e2[1] : e2[0], pair ? sv2[1] : sv2[0], pair ? d0[1] : d0[0], ro);
      rhs[0] = shfl_from(ro[0], lane0); rhs[1] = shfl_from(ro[1], lane0);
      rhs[2] = shfl_from(ro[0], lane0 | 2); rhs[3] = shfl_from(ro[1], lane0 | 2);
    }
    double G[4][4];
    G[0][0] = shfl_from(g00, lane0); G[1][0] = shfl_from(g10, lane0); G[1][1] = shfl_from(g11, lane0);
    G[2][2] = shfl_from(g00, lane0 | 2); G[3][2] = shfl_from(g10, lane0 | 2); G[3][3] = shfl_from(g11, lane0 | 2);
    unsigned long long slot = ~0ULL;
    bool packed = false;
    if (cont) {
      ++it;
      G[2][0] = ccp_row_dot7(P1[0], P0[0]); G[2][1] = ccp_row_dot7(P1[0], P0[1]);
      G[3][0] = ccp_row_dot7(P1[1], P0[0]); G[3][1] = ccp_row_dot7(P1[1], P0[1]);
      G[0][0] = CCP_FMA(M.damping, e2[0], G[0][0]); G[1][1] = CCP_FMA(M.damping, sv2[0], G[1][1]);
      G[2][2] = CCP_FMA(M.damping, e2[1], G[2][2]); G[3][3] = CCP_FMA(M.damping, sv2[1], G[3][3]);
      double y[4];
      ccp_solve_ldlt<4>(G, rhs, y);
      if (is0) {
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
        const double ya = pair ? y[2] : y[0], yb = pair ? y[3] : y[1];
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
    } else if (live) {
      // ---- the sample is finished (ConstraintFunction.h:75-81) ----
      const bool cv = (e2[0] <= M.tol_p2) && (sv2[0] < dw2a) && (e2[1] <= M.tol_p2) && (sv2[1] < dw2b);
      const bool okk = cv && bad == 0u;
      if (A.wrap) {
#pragma unroll
        for (int j = 0; j < H; ++j) x[j] = ccp_wrap_pi_call(x[j]);
      }
      if (A.x_out && g != 3) {
#pragma unroll
        for (int j = 0; j < H; ++j) st_elem<SOA>(A.x_out, idx, arm * H + j, A.out_stride, n, x[j]);
      }
      if (g == 0) {
        if (A.ok) A.ok[idx] = okk;
        if (A.conv) A.conv[idx] = cv;
        if (A.iters) A.iters[idx] = it;
        if (A.resid) {
          ccp_fwd<3> F;
          F.e2[0] = e2[0]; F.sv2[0] = sv2[0]; F.d[0][0] = d0[0];
          F.e2[1] = e2[1]; F.sv2[1] = sv2[1]; F.d[1][0] = d0[1];
          double fv[4];
          ccp_residual<3>(F, fv, nullptr);
#pragma unroll
          for (int k = 0; k < 4; ++k) st_elem<SOA>(A.resid, idx, k, A.out_stride, 4, fv[k]);
        }
        if (A.n_ok && okk) slot = atomicAdd(A.n_ok, 1ULL);
        u = A.counter ? static_samples + atomicAdd((unsigned*)A.counter, 1u) : total;
      }
      packed = A.n_ok && okk && A.compact;
      idx = CCP_NO_SAMPLE;
      fresh = true;
    }
    if (__any_sync(CCP_FULL, packed)) {
      const unsigned lo = __shfl_sync(CCP_FULL, (unsigned)slot, lane0);
      const unsigned hi = __shfl_sync(CCP_FULL, (unsigned)(slot >> 32), lane0);
      if (packed && g != 3) {
        const unsigned long long s2 = ((unsigned long long)hi << 32) | lo;
#pragma unroll
        for (int j = 0; j < H; ++j) A.compact[s2 * n + arm * H + j] = x[j];
      }
    }
    fresh = __any_sync(CCP_FULL, fresh);
  }
}

template <int PANDA>
static cudaError_t launch_coop3(int sm_count, const ccp_model& M, const ccp_project_args& A, bool soa, cudaStream_t st) {
  const int grid = ccp_coop_grid(sm_count, A.count, 8);  // 8 samples per warp
  if (soa) ccp_project_coop3_kernel<PANDA, true><<<grid, CCP_COOP_BLOCK, 0, st>>>(M, A);
  else ccp_project_coop3_kernel<PANDA, false><<<grid, CCP_COOP_BLOCK, 0, st>>>(M, A);
  return cudaGetLastError();
}

template <int PANDA>
static cudaError_t launch_coop(int sm_count, const ccp_model& M, const ccp_project_args& A, bool soa, cudaStream_t st) {
  const int grid = ccp_coop_grid(sm_count, A.count, 16);  // 16 samples per warp
  if (soa) ccp_project_coop_kernel<PANDA, true><<<grid, CCP_COOP_BLOCK, 0, st>>>(M, A);
  else ccp_project_coop_kernel<PANDA, false><<<grid, CCP_COOP_BLOCK, 0, st>>>(M, A);
  return cudaGetLastError();
}

cudaError_t ccp_launch_project_coop(int sm_count, const ccp_model& M, const ccp_project_args& A, bool soa, cudaStream_t st) {
  if (M.n_arms == 3) {
    if (M.stock) return launch_coop3<2>(sm_count, M, A, soa, st);
    return M.panda_alpha ? launch_coop3<1>(sm_count, M, A, soa, st) : launch_coop3<0>(sm_count, M, A, soa, st);
  }
  if (M.n_arms != 2) return cudaErrorInvalidValue;
  if (M.stock) return launch_coop<2>(sm_count, M, A, soa, st);
  return M.panda_alpha ? launch_coop<1>(sm_count, M, A, soa, st) : launch_coop<0>(sm_count, M, A, soa, st);
}
