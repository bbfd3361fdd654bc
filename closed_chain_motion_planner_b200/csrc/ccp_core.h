// ccp_core.h — the engine arithmetic of the closed-chain projection, written once as
// __host__ __device__ code.  The CUDA kernels (ccp_project.cu, ccp_geodesic.cu, ccp_api.cu) are the product; the SAME header
// compiled by g++ (oracle/oracle_b.cpp, test infrastructure only) reproduces the device results
// bit for bit, because
//   * every fused multiply-add is an explicit fma() and nothing else may be contracted
//     (nvcc --fmad=false, g++ -ffp-contract=off),
//   * sqrt and division are IEEE-754 correctly rounded on both sides,
//   * sin/cos/atan2 are our own polynomial kernels (glibc and CUDA libm differ by >= 1 ulp).
//
// What it computes (reference: jkw0701/closed_chain_motion_planner)
//   function()  ConstraintFunction.h:84-102     relative pose of arm a's EE against arm 0's EE
//   jacobian()  called at ConstraintFunction.h:70 (OMPL finite differences there; analytic here)
//   project()   ConstraintFunction.h:57-82      fixed-step min-norm Newton iteration
//   FK          panda_rbdl.cpp:24-42,73-161     modified-DH chain, flange 0.107, Rz(-pi/4)
//
// Formulation (one thread owns one sample; all state is a handful of 3-vectors/quaternions):
//   rotation    world quaternion of each arm's EE by right-multiplying the link quaternions
//               q_Rx(alpha_i) (x) q_Rz(theta_i/2): half-angle sincos, no rot->quat conversion.
//   position    the EE-0 origin is pushed DOWN arm 0 (tip -> base), through the world, and UP
//               arm a (base -> tip): 3-vector recursions only, no 3x3 products.
//   residual    e = t_c - t_0, d = q_c (x) conj(q_0):  f0 = |e|,  f1 = 2 atan2(|vec d|, |d_w|).
//   loop tests  on the squares, so no sqrt / atan2 sits between an evaluation and the decision to step:
//               f0 > tol1  <=>  |e|^2 > tol1^2 ;  f1 > tol2  <=>  |vec d|^2 > tan^2(tol2/2) d_w^2.
//   jacobian    UNNORMALISED gradient rows  g0 = f0 grad f0 = e . de/dq ,  g1 = s |vec d| grad f1
//               (s = sign d_w):  carry (r, w, m) = (lever arm to EE-0, e, vec d) down each arm; at
//               joint i  g0_i = -+ (r x w)_z,  g1_i = -+ m_z  (+ on arm 0, - on arm a).  No division by
//               f0 or |vec d|: with D = diag(1/f0, s/|vec d|), J = D g and
//   step        x -= step J^T (J J^T)^-1 f = step g^T (g g^T)^-1 (f0^2, s |vec d| f1): the same minimum-norm
//               Newton step; (g g^T) is solved by Cramer (2x2) or L D L^T (4x4) in registers; rows that
//               vanish or make D_k <= 0 are dropped (what JacobiSVD's rank threshold does).  sqrt and
//               atan2 are only needed for the right-hand side, so they overlap the Jacobian pass.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define CCP_HD __host__ __device__ __forceinline__
#else
#define CCP_HD inline __attribute__((always_inline))
#endif

#define CCP_FMA(a, b, c) fma((a), (b), (c))

#define CCPC_DOF 7
#define CCPC_MAX_ARMS 3

// ------------------------------------------------------------------------------------------
// Packed model (device constant / shared memory image).  Built by ccp_pack_model() on the host.
// ------------------------------------------------------------------------------------------
struct ccp_link {
  double tx, ty, tz;  // link translation (a, -sin(alpha) d, cos(alpha) d), panda_rbdl.cpp:159
  double sa, ca;      // sin/cos alpha
  double sha, cha;    // sin/cos alpha/2
  double qoff;        // theta offset (calibration dh.col(2))
  double hqoff;       // half of it (the quaternion chain works on half angles)
};

struct ccp_arm {
  ccp_link link[CCPC_DOF];
  double qwb[4];  // base orientation in world, (w,x,y,z)
  double Rwb[9];  // same, row-major
  double pwb[3];
  double fl;            // flange offset along z7
  double sphi, cphi;    // sin/cos of the EE yaw (-pi/4)
  double shphi, chphi;  // half angle
  // Joint 7 turns about the axis the flange offset and the EE yaw also use, so in the closed-chain code the yaw is
  // folded into joint 7's angle (frame 7' = the EE frame turned back by nothing: Rz(theta7) Tz(fl) Rz(phi) =
  // Tz(fl) Rz(theta7 + phi)) and the EE origin seen from frame 6 is a constant:
  // base of arm 0 seen from THIS arm's base (arms a >= 1): x_a = Rrel x_0 + prel, and q_rel = conj(qwb_a) qwb_0
  double Rrel[9], prel[3], qrel[4];
  double qrel_scaled[4];  // qrel / 64: the unscaled structured link quaternions (1, +-1, 0, 0) of two arms leave 2^6
  double hq7;           // link[6].hqoff + phi / 2
  double r6[3];         // (0, 0, fl) pushed down link 7: (t_x, -sin(alpha) fl + t_y, cos(alpha) fl + t_z) of link[6]
};

struct ccp_pair_ref {
  double t0[3];  // init_chain_ translation           (ConstraintFunction.h:39)
  double q0[4];  // init_chain_ rotation, (w,x,y,z)
};

struct ccp_model {
  int32_t n_arms;
  int32_t max_iter;
  int32_t panda_alpha;  // 1: every arm has the stock Panda alpha pattern (0,-pi/2,pi/2,pi/2,-pi/2,pi/2,pi/2)
                        //    EXACTLY (no alpha calibration): the kernels use the structured link code
  int32_t stock;        // 1 (implies panda_alpha): additionally no theta calibration, exact zeros where the stock a / d
                        //    tables have zeros, and every base-to-base rotation Rrel a diagonal of +-1 (the reference's
                        //    grasping_point frames): the kernels skip those terms (template argument PANDA = 2)
  double tol_p, tol_r;  // tolerance1_, tolerance2_
  double tol_p2;        // tol_p^2
  double tan2_r;        // tan^2(tol_r / 2): f1 > tol_r  <=>  |vec d|^2 > tan2_r d_w^2   (0 < tol_r < pi)
  double step;          // 0.30
  double margin;        // 1e-3
  // opt-in modes of the north star's description, OFF in the reference and by default (parity):
  double damping;       // lambda^2 added to the diagonal of the rows' Gram matrix (damped least squares); 0 = min-norm
  int32_t clamp;        // 1: clamp every iterate to [lb, ub] (the reference only CHECKS the limits at the end)
  int32_t reserved2;
  double lb[CCPC_DOF], ub[CCPC_DOF];
  double lbm[CCPC_DOF], ubm[CCPC_DOF];  // lb + margin, ub - margin (jointValid's thresholds, precomputed)
  ccp_arm arm[CCPC_MAX_ARMS];
  ccp_pair_ref ref[CCPC_MAX_ARMS - 1];
};

// joint margin (ConstraintFunction.h:45) + jointValid's thresholds
static inline void ccp_model_set_margin(ccp_model* M, double margin) {
  M->margin = margin;
  for (int i = 0; i < CCPC_DOF; ++i) {
    M->lbm[i] = M->lb[i] + margin;
    M->ubm[i] = M->ub[i] - margin;
  }
}

// setTolerance (ConstraintFunction.h:104-112) + the derived thresholds of the loop tests.  Host only (libm tan).
static inline void ccp_model_set_tolerance(ccp_model* M, double tol_p, double tol_r) {
  M->tol_p = tol_p;
  M->tol_r = tol_r;
  M->tol_p2 = tol_p * tol_p;
  const double t = tan(0.5 * tol_r);
  M->tan2_r = (0.5 * tol_r < 1.5707963267948966) ? t * t : HUGE_VAL;  // f1 <= pi always
}

// ------------------------------------------------------------------------------------------
// Elementary functions
// ------------------------------------------------------------------------------------------
CCP_HD int32_t ccp_lo32(double t) {
#if defined(__CUDA_ARCH__)
  return __double2loint(t);
#else
  uint64_t u;
  memcpy(&u, &t, sizeof u);
  return (int32_t)(uint32_t)u;
#endif
}

// sin and cos of x by table + short series: x = k pi/512 + r with |r| <= pi/1024 (Cody-Waite in two pieces: k P1 is
// exact for |k| < 2^20, the second product is rounded once); (sin, cos)(k pi/512) come from a 1024-entry table of
// correctly rounded values (16 KB, L1-resident), sin r / cos r - 1 from their Taylor series to r^5 / r^4
// (truncation < 2e-18), and
//     sin x = S + (S (cos r - 1) + C sin r),   cos x = C + (C (cos r - 1) - S sin r).
// 14 FP64 instructions and one 16-byte table load; no quadrant selects.  <= 2 ulp for |x| < ~1e5; degrades
// gracefully (and identically on host and device) beyond.
#include "ccp_sincos_table.h"

static const double ccp_sc_table_host[1 << CCP_SC_TABLE_BITS][2] = {CCP_SC_TABLE_ENTRIES};
#if defined(__CUDACC__)
static __device__ __align__(16) const double ccp_sc_table_dev[1 << CCP_SC_TABLE_BITS][2] = {CCP_SC_TABLE_ENTRIES};
#endif

// Constants of the elementary functions.  On the device they sit in the constant bank (pairs come in with one
// LDCU.128); written as literals each of them costs two UMOVs per use site.
#define CCP_ELEM_CONSTS                                                                                     \
  {0x1.45f306dc9c883p+7,  /*  0 512 / pi                        */                                          \
   0x1.921fb54400000p-8,  /*  1 pi/512, leading 33 bits         */                                          \
   0x1.0b4611a626331p-42, /*  2 pi/512 - [1] to 53 bits         */                                          \
   0x1.1111111111111p-7,  /*  3 1/5!                            */                                          \
   -0x1.5555555555555p-3, /*  4 -1/3!                           */                                          \
   0x1.5555555555555p-5,  /*  5 1/4!                            */                                          \
   0x1.561b82ab7f990p-1,  /*  6 tan(3 pi/16)                    */                                          \
   0x1.975f5e0553158p-3,  /*  7 tan(pi/16)                      */                                          \
   0x1.a827999fcef32p-2,  /*  8 tan(pi/8)                       */                                          \
   0x1.921fb54442d18p-1,  /*  9 pi/4                            */                                          \
   0x1.921fb54442d18p-2,  /* 10 pi/8                            */                                          \
   0x1.921fb54442d18p+0,  /* 11 pi/2                            */                                          \
   -0x1.642c8590b2164p-5, /* 12 -1/23                           */                                          \
   0x1.8618618618618p-5,  /* 13 1/21                            */                                          \
   -0x1.af286bca1af28p-5, /* 14 -1/19                           */                                          \
   0x1.e1e1e1e1e1e1ep-5,  /* 15 1/17                            */                                          \
   -0x1.1111111111111p-4, /* 16 -1/15                           */                                          \
   0x1.3b13b13b13b14p-4,  /* 17 1/13                            */                                          \
   -0x1.745d1745d1746p-4, /* 18 -1/11                           */                                          \
   0x1.c71c71c71c71cp-4,  /* 19 1/9                             */                                          \
   -0x1.2492492492492p-3, /* 20 -1/7                            */                                          \
   0x1.999999999999ap-3,  /* 21 1/5                             */                                          \
   -0x1.5555555555555p-2, /* 22 -1/3                            */                                          \
   0.0}
static const double ccp_elem_host[24] = CCP_ELEM_CONSTS;
#if defined(__CUDACC__)
static __constant__ __align__(16) double ccp_elem_dev[24] = CCP_ELEM_CONSTS;
#endif
#if defined(__CUDA_ARCH__)
#define CCP_EC(i) ccp_elem_dev[i]
#else
#define CCP_EC(i) ccp_elem_host[i]
#endif

CCP_HD void ccp_sincos(double x, double* s_out, double* c_out) {
  const double SCALE = CCP_EC(0);            // 512 / pi
  const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52: round-to-nearest-integer trick
  const double P1 = CCP_EC(1);               // pi/512, leading 33 bits
  const double P2 = CCP_EC(2);               // pi/512 - P1 to 53 bits
  const double t = CCP_FMA(x, SCALE, MAGIC);
  const uint32_t idx = (uint32_t)ccp_lo32(t) & ((1u << CCP_SC_TABLE_BITS) - 1u);
  const double k = t - MAGIC;
  double r = CCP_FMA(-k, P1, x);
  r = CCP_FMA(-k, P2, r);
#if defined(__CUDA_ARCH__)
  const double2 sc = __ldg(reinterpret_cast<const double2*>(&ccp_sc_table_dev[idx][0]));
  const double S = sc.x, C = sc.y;
#else
  const double S = ccp_sc_table_host[idx][0], C = ccp_sc_table_host[idx][1];
#endif
  const double z = r * r;
  const double ps = CCP_FMA(z, CCP_EC(3), CCP_EC(4));                         // 1/5!, -1/3!
  const double sr = CCP_FMA(r * z, ps, r);                                    // sin r
  const double pc = CCP_FMA(z, CCP_EC(5), -0.5);                              // 1/4!, -1/2!
  const double cm1 = z * pc;                                                  // cos r - 1
  *s_out = CCP_FMA(C, sr, CCP_FMA(S, cm1, S));
  *c_out = CCP_FMA(-S, sr, CCP_FMA(C, cm1, C));
}

// atan2(y, x) for y >= 0, x >= 0 (the only case angularDistance needs).  Result in [0, pi/2].
// min/max -> t in [0,1]; shift by atan(0), pi/8 or pi/4 so |t'| <= tan(pi/16); 12-term series; one division.
CCP_HD double ccp_atan2_pos(double y, double x) {
  const bool inv = y > x;
  const double num = inv ? x : y;
  const double den = inv ? y : x;
  // t = num/den in [0,1] is never formed: the region tests and the shifted argument
  // t' = (t - t0)/(1 + t t0) = (num - t0 den)/(den + t0 num) need ONE division in total
  const bool hi = num > CCP_EC(6) * den;   // t > tan(3 pi/16)
  const bool mid = num > CCP_EC(7) * den;  // t > tan(pi/16)
  const double t0 = hi ? 1.0 : (mid ? CCP_EC(8) : 0.0);          // 1, tan(pi/8), 0
  const double off = hi ? CCP_EC(9) : (mid ? CCP_EC(10) : 0.0);  // pi/4, pi/8, 0
  double tr = (den > 0.0) ? CCP_FMA(-t0, den, num) / CCP_FMA(t0, num, den) : 0.0;
  double z = tr * tr;
  double p = CCP_FMA(z, CCP_EC(12), CCP_EC(13));  // -1/23, 1/21
  p = CCP_FMA(z, p, CCP_EC(14));                  // -1/19
  p = CCP_FMA(z, p, CCP_EC(15));                  // 1/17
  p = CCP_FMA(z, p, CCP_EC(16));                  // -1/15
  p = CCP_FMA(z, p, CCP_EC(17));                  // 1/13
  p = CCP_FMA(z, p, CCP_EC(18));                  // -1/11
  p = CCP_FMA(z, p, CCP_EC(19));                  // 1/9
  p = CCP_FMA(z, p, CCP_EC(20));                  // -1/7
  p = CCP_FMA(z, p, CCP_EC(21));                  // 1/5
  p = CCP_FMA(z, p, CCP_EC(22));                  // -1/3
  double a = off + CCP_FMA(tr * z, p, tr);
  return inv ? (CCP_EC(11) - a) : a;
}

// ------------------------------------------------------------------------------------------
// Small vector / quaternion helpers (operation order is part of the definition)
// ------------------------------------------------------------------------------------------
// (x,y) <- Rz(theta) (x,y)
CCP_HD void ccp_rot2(double c, double s, double& x, double& y) {
  double nx = CCP_FMA(c, x, -(s * y));
  double ny = CCP_FMA(s, x, c * y);
  x = nx;
  y = ny;
}
// (x,y) <- Rz(theta)^T (x,y)
CCP_HD void ccp_rot2t(double c, double s, double& x, double& y) {
  double nx = CCP_FMA(c, x, s * y);
  double ny = CCP_FMA(c, y, -(s * x));
  x = nx;
  y = ny;
}

// q <- q (x) (a, b, 0, 0)   rotation about x
CCP_HD void ccp_qmul_rx(double* q, double a, double b) {
  double w = CCP_FMA(q[0], a, -(q[1] * b));
  double x = CCP_FMA(q[1], a, q[0] * b);
  double y = CCP_FMA(q[2], a, q[3] * b);
  double z = CCP_FMA(q[3], a, -(q[2] * b));
  q[0] = w; q[1] = x; q[2] = y; q[3] = z;
}
// q <- q (x) (c, 0, 0, s)   rotation about z
CCP_HD void ccp_qmul_rz(double* q, double c, double s) {
  double w = CCP_FMA(q[0], c, -(q[3] * s));
  double x = CCP_FMA(q[1], c, q[2] * s);
  double y = CCP_FMA(q[2], c, -(q[1] * s));
  double z = CCP_FMA(q[3], c, q[0] * s);
  q[0] = w; q[1] = x; q[2] = y; q[3] = z;
}
// r = conj(a) (x) b
CCP_HD void ccp_qmul_conj_left(const double* a, const double* b, double* r) {
  const double w1 = a[0], x1 = -a[1], y1 = -a[2], z1 = -a[3];
  r[0] = CCP_FMA(w1, b[0], -CCP_FMA(x1, b[1], CCP_FMA(y1, b[2], z1 * b[3])));
  r[1] = CCP_FMA(w1, b[1], CCP_FMA(x1, b[0], CCP_FMA(y1, b[3], -(z1 * b[2]))));
  r[2] = CCP_FMA(w1, b[2], CCP_FMA(y1, b[0], CCP_FMA(z1, b[1], -(x1 * b[3]))));
  r[3] = CCP_FMA(w1, b[3], CCP_FMA(z1, b[0], CCP_FMA(x1, b[2], -(y1 * b[1]))));
}
// r = a (x) conj(b)
CCP_HD void ccp_qmul_conj_right(const double* a, const double* b, double* r) {
  const double w2 = b[0], x2 = -b[1], y2 = -b[2], z2 = -b[3];
  r[0] = CCP_FMA(a[0], w2, -CCP_FMA(a[1], x2, CCP_FMA(a[2], y2, a[3] * z2)));
  r[1] = CCP_FMA(a[0], x2, CCP_FMA(a[1], w2, CCP_FMA(a[2], z2, -(a[3] * y2))));
  r[2] = CCP_FMA(a[0], y2, CCP_FMA(a[2], w2, CCP_FMA(a[3], x2, -(a[1] * z2))));
  r[3] = CCP_FMA(a[0], z2, CCP_FMA(a[3], w2, CCP_FMA(a[1], y2, -(a[2] * x2))));
}
// v <- R(q)^T v  (rotate by the conjugate of unit quaternion q)
CCP_HD void ccp_qrot_inv(const double* q, double* v) {
  // a = -vec(q);  t = a x v;  v' = v + 2 (w t + a x t)
  const double ax = -q[1], ay = -q[2], az = -q[3];
  const double tx = CCP_FMA(ay, v[2], -(az * v[1]));
  const double ty = CCP_FMA(az, v[0], -(ax * v[2]));
  const double tz = CCP_FMA(ax, v[1], -(ay * v[0]));
  const double ex = CCP_FMA(q[0], tx, CCP_FMA(ay, tz, -(az * ty)));
  const double ey = CCP_FMA(q[0], ty, CCP_FMA(az, tx, -(ax * tz)));
  const double ez = CCP_FMA(q[0], tz, CCP_FMA(ax, ty, -(ay * tx)));
  v[0] = CCP_FMA(2.0, ex, v[0]);
  v[1] = CCP_FMA(2.0, ey, v[1]);
  v[2] = CCP_FMA(2.0, ez, v[2]);
}

// ------------------------------------------------------------------------------------------
// Link code.  PANDA = true compiles the stock Panda alpha pattern in: alpha_i in {0, +-pi/2} exactly, so
// Rx(alpha) is a signed permutation (no arithmetic), the structurally zero translation component is
// skipped, and the link quaternion (1, +-1, 0, 0) is applied UNSCALED (4 adds); the accumulated factor
// sqrt(2)^12 = 64 is removed from the chain quaternion with one exact scaling.
// PANDA = false is the generic path (calibrated alpha offsets, panda_rbdl.cpp:92-95).
// ------------------------------------------------------------------------------------------
// PANDA = 2 ("stock"): on top of the structured alpha pattern the link table is the stock one where that one is ZERO —
// a = (0, 0, 0, .0825, -.0825, 0, .088), d = (.333, 0, .316, 0, .384, 0, 0) (panda_rbdl.cpp:98-99), no theta calibration —
// and the base-to-base rotation is a diagonal of +-1.  Terms that are exactly zero are not computed and their constants not
// loaded (FP64 instructions take no constant-bank operand: every constant is a load).  x + 0 and x - 0 are x and
// fma(a, b, 0) is a b rounded once, so the results equal the PANDA = 1 code's except for the sign of an exact zero.
template <int I>
struct ccp_stock_zero {
  static constexpr bool a = (I == 0 || I == 1 || I == 2 || I == 5);  // a_I == 0: no x translation
  static constexpr bool d = (I == 1 || I == 3 || I == 5 || I == 6);  // d_I == 0: no y / z translation
};
template <int I>
struct ccp_panda_sgn {  // sign of alpha_I / (pi/2): link 0 has alpha = 0
  static constexpr int value = (I == 0) ? 0 : ((I == 1 || I == 4) ? -1 : 1);
};
#define CCP_PANDA_QSCALE 0.015625  // 1/64 = (1/sqrt 2)^12: six quarter-turn links per arm, two arms

// One link going DOWN the chain (frame i -> frame i-1): v <- Rx(alpha) Rz(theta) v (+ t)
template <int PANDA, int I>
CCP_HD void ccp_down_vec(const ccp_link& L, double s, double c, double* v) {
  ccp_rot2(c, s, v[0], v[1]);
  if (!PANDA) {
    ccp_rot2(L.ca, L.sa, v[1], v[2]);
  } else if (ccp_panda_sgn<I>::value > 0) {
    const double y = v[1];
    v[1] = -v[2];
    v[2] = y;
  } else if (ccp_panda_sgn<I>::value < 0) {
    const double y = v[1];
    v[1] = v[2];
    v[2] = -y;
  }
}
template <int PANDA, int I>
CCP_HD void ccp_down_pt(const ccp_link& L, double s, double c, double* r) {
  // as ccp_down_vec, with the x translation folded into the rotation's FMA chain
  const double nx = (PANDA == 2 && ccp_stock_zero<I>::a) ? CCP_FMA(c, r[0], -(s * r[1])) : CCP_FMA(c, r[0], CCP_FMA(-s, r[1], L.tx));
  const double ny = CCP_FMA(s, r[0], c * r[1]);
  r[0] = nx;
  r[1] = ny;
  if (!PANDA) {
    ccp_rot2(L.ca, L.sa, r[1], r[2]);
  } else if (ccp_panda_sgn<I>::value > 0) {
    const double y = r[1];
    r[1] = -r[2];
    r[2] = y;
  } else if (ccp_panda_sgn<I>::value < 0) {
    const double y = r[1];
    r[1] = r[2];
    r[2] = -y;
  }
  if (PANDA == 2 && ccp_stock_zero<I>::d) return;
  if (!PANDA || ccp_panda_sgn<I>::value != 0) r[1] += L.ty;
  if (!PANDA || ccp_panda_sgn<I>::value == 0) r[2] += L.tz;
}
// One link going UP the chain (frame i-1 -> frame i): r <- Rz(theta)^T Rx(alpha)^T (r - t)
template <int PANDA, int I>
CCP_HD void ccp_up_pt(const ccp_link& L, double s, double c, double* r) {
  if (!(PANDA == 2 && ccp_stock_zero<I>::a)) r[0] -= L.tx;
  if (!(PANDA == 2 && ccp_stock_zero<I>::d)) {
    if (!PANDA || ccp_panda_sgn<I>::value != 0) r[1] -= L.ty;
    if (!PANDA || ccp_panda_sgn<I>::value == 0) r[2] -= L.tz;
  }
  if (!PANDA) {
    ccp_rot2t(L.ca, L.sa, r[1], r[2]);
  } else if (ccp_panda_sgn<I>::value > 0) {
    const double y = r[1];
    r[1] = r[2];
    r[2] = -y;
  } else if (ccp_panda_sgn<I>::value < 0) {
    const double y = r[1];
    r[1] = -r[2];
    r[2] = y;
  }
  ccp_rot2t(c, s, r[0], r[1]);
}
// q <- q (x) q_Rx(alpha_I)   (unscaled in PANDA mode)
template <int PANDA, int I>
CCP_HD void ccp_qmul_link_rx(const ccp_link& L, double* q) {
  if (!PANDA) {
    ccp_qmul_rx(q, L.cha, L.sha);
  } else if (ccp_panda_sgn<I>::value > 0) {
    const double w = q[0] - q[1], x = q[1] + q[0], y = q[2] + q[3], z = q[3] - q[2];
    q[0] = w; q[1] = x; q[2] = y; q[3] = z;
  } else if (ccp_panda_sgn<I>::value < 0) {
    const double w = q[0] + q[1], x = q[1] - q[0], y = q[2] - q[3], z = q[3] + q[2];
    q[0] = w; q[1] = x; q[2] = y; q[3] = z;
  }
}

// Per-sample storage of the 7K (sin, cos) pairs between the passes.  The host build and the simple
// kernels keep them in a local array; the projection kernel keeps them in shared memory
// ([slot][thread], conflict-free) to free ~56 registers.
// The forward pass also leaves, per joint, the (x, y) components of the lever arm to the EE-0 origin
// expressed in that joint's frame (rx, ry): the Jacobian pass reads them instead of recomputing them.
template <int K>
struct ccp_sc_local {
  double v[K][CCPC_DOF][4];
  CCP_HD double& s(int a, int i) { return v[a][i][0]; }
  CCP_HD double& c(int a, int i) { return v[a][i][1]; }
  CCP_HD double& rx(int a, int i) { return v[a][i][2]; }
  CCP_HD double& ry(int a, int i) { return v[a][i][3]; }
  CCP_HD double s(int a, int i) const { return v[a][i][0]; }
  CCP_HD double c(int a, int i) const { return v[a][i][1]; }
  CCP_HD double rx(int a, int i) const { return v[a][i][2]; }
  CCP_HD double ry(int a, int i) const { return v[a][i][3]; }
};

// ------------------------------------------------------------------------------------------
// Forward evaluation: residual f(x) and everything the Jacobian pass needs.
// ------------------------------------------------------------------------------------------
template <int K>
struct ccp_fwd {
  double tc[K - 1][3];        // translation of chain a:  R_a^T (p_0 - p_a)
  double qc[K - 1][4];        // rotation of chain a:     conj(q_a) (x) q_0   (unit)
  double d[K - 1][4];         // qc (x) conj(q_ref)
  double e[K - 1][3];         // tc - t_ref
  double e2[K - 1];           // |e|^2        (f0 = sqrt(e2))
  double sv2[K - 1];          // |vec d|^2    (f1 = 2 atan2(sqrt(sv2), |d_w|))
};

// the residual values function() reports (ConstraintFunction.h:84-102), and |vec d|
template <int K>
CCP_HD void ccp_residual(const ccp_fwd<K>& F, double* f, double* sv_out) {
#pragma unroll
  for (int p = 0; p < K - 1; ++p) {
    f[2 * p] = sqrt(F.e2[p]);
    const double sv = sqrt(F.sv2[p]);
    const double atn = ccp_atan2_pos(sv, fabs(F.d[p][0]));
    f[2 * p + 1] = atn + atn;
    if (sv_out) sv_out[p] = sv;
  }
}

// FROM_IDENTITY: the chain starts here from the identity quaternion (link 0 of an arm whose base rotation was folded
// into the other arm's start): q = q_Rx(alpha_0) (x) q_Rz(theta/2) without a multiplication when alpha_0 = 0
// `a` only indexes x and S (the cooperative kernel passes a lane's own 7 joints with a = 0); A is the arm's constants.
template <int PANDA, int I, bool FROM_IDENTITY = false, class SC, class XT>
CCP_HD void ccp_fwd_link_quat(const ccp_arm& A, int a, const XT& x, double* q, SC& S) {
  const ccp_link& L = A.link[I];
  // joint 7 carries the EE yaw; stock: no theta calibration on the others
  double h = (PANDA == 2 && I != 6) ? 0.5 * x[a * CCPC_DOF + I] : CCP_FMA(0.5, x[a * CCPC_DOF + I], (I == 6) ? A.hq7 : L.hqoff);
  double sh, ch;
  ccp_sincos(h, &sh, &ch);
  if (FROM_IDENTITY && PANDA && I == 0) {
    q[0] = ch; q[1] = 0.0; q[2] = 0.0; q[3] = sh;
  } else {
    if (FROM_IDENTITY) {
      q[0] = 1.0; q[1] = 0.0; q[2] = 0.0; q[3] = 0.0;
    }
    ccp_qmul_link_rx<PANDA, I>(L, q);
    ccp_qmul_rz(q, ch, sh);
  }
  double sh2 = sh + sh;
  S.s(a, I) = sh2 * ch;                // sin(theta)
  S.c(a, I) = CCP_FMA(-sh2, sh, 1.0);  // cos(theta)
}

// ---- the pieces of the forward evaluation, one arm at a time (ccp_forward composes them; the cooperative kernel runs
// them on the lane that owns the arm) ----
// links 1..6 of an arm's quaternion chain (link 0 differs between the arms: see ccp_forward)
template <int PANDA, class SC, class XT>
CCP_HD void ccp_fwd_quat_links_1_6(const ccp_arm& A, int a, const XT& x, double* q, SC& S) {
  ccp_fwd_link_quat<PANDA, 1>(A, a, x, q, S);
  ccp_fwd_link_quat<PANDA, 2>(A, a, x, q, S);
  ccp_fwd_link_quat<PANDA, 3>(A, a, x, q, S);
  ccp_fwd_link_quat<PANDA, 4>(A, a, x, q, S);
  ccp_fwd_link_quat<PANDA, 5>(A, a, x, q, S);
  ccp_fwd_link_quat<PANDA, 6>(A, a, x, q, S);
}
// EE-0 origin: (0, 0, fl) in frame 7' of arm 0 (on joint 7's axis: no lever arm there, and its image in frame 6
// is the constant r6) -> base 0.  Leaves the lever arms (rx, ry) of joints 5..0.  a = index of arm 0 in S.
template <int PANDA, class SC>
CCP_HD void ccp_fwd_down_arm0(const ccp_arm& A, int a, SC& S, double* r) {
  r[0] = A.r6[0]; r[1] = A.r6[1]; r[2] = A.r6[2];
  S.rx(a, 5) = r[0]; S.ry(a, 5) = r[1];
  ccp_down_pt<PANDA, 5>(A.link[5], S.s(a, 5), S.c(a, 5), r);
  S.rx(a, 4) = r[0]; S.ry(a, 4) = r[1];
  ccp_down_pt<PANDA, 4>(A.link[4], S.s(a, 4), S.c(a, 4), r);
  S.rx(a, 3) = r[0]; S.ry(a, 3) = r[1];
  ccp_down_pt<PANDA, 3>(A.link[3], S.s(a, 3), S.c(a, 3), r);
  S.rx(a, 2) = r[0]; S.ry(a, 2) = r[1];
  ccp_down_pt<PANDA, 2>(A.link[2], S.s(a, 2), S.c(a, 2), r);
  S.rx(a, 1) = r[0]; S.ry(a, 1) = r[1];
  ccp_down_pt<PANDA, 1>(A.link[1], S.s(a, 1), S.c(a, 1), r);
  S.rx(a, 0) = r[0]; S.ry(a, 0) = r[1];
  ccp_down_pt<PANDA, 0>(A.link[0], S.s(a, 0), S.c(a, 0), r);
}
// arm a >= 1: base 0 -> base a in one constant transform (t_wb_a^-1 t_wb_0, packed on the host), then up the arm to
// frame 7' (= the EE frame up to the flange shift along z).  r: the EE-0 origin in base 0; v: tc = R_a^T (p_0 - p_a).
template <int PANDA, class SC>
CCP_HD void ccp_fwd_up_arm(const ccp_arm& A, int a, SC& S, const double* r, double* v) {
  if (PANDA == 2) {  // a diagonal of +-1 (the reference's base frames: translations, and the top arm turned about z)
    v[0] = CCP_FMA(A.Rrel[0], r[0], A.prel[0]);
    v[1] = CCP_FMA(A.Rrel[4], r[1], A.prel[1]);
    v[2] = CCP_FMA(A.Rrel[8], r[2], A.prel[2]);
  } else {
    v[0] = CCP_FMA(A.Rrel[0], r[0], CCP_FMA(A.Rrel[1], r[1], CCP_FMA(A.Rrel[2], r[2], A.prel[0])));
    v[1] = CCP_FMA(A.Rrel[3], r[0], CCP_FMA(A.Rrel[4], r[1], CCP_FMA(A.Rrel[5], r[2], A.prel[1])));
    v[2] = CCP_FMA(A.Rrel[6], r[0], CCP_FMA(A.Rrel[7], r[1], CCP_FMA(A.Rrel[8], r[2], A.prel[2])));
  }
  ccp_up_pt<PANDA, 0>(A.link[0], S.s(a, 0), S.c(a, 0), v);
  S.rx(a, 0) = v[0]; S.ry(a, 0) = v[1];
  ccp_up_pt<PANDA, 1>(A.link[1], S.s(a, 1), S.c(a, 1), v);
  S.rx(a, 1) = v[0]; S.ry(a, 1) = v[1];
  ccp_up_pt<PANDA, 2>(A.link[2], S.s(a, 2), S.c(a, 2), v);
  S.rx(a, 2) = v[0]; S.ry(a, 2) = v[1];
  ccp_up_pt<PANDA, 3>(A.link[3], S.s(a, 3), S.c(a, 3), v);
  S.rx(a, 3) = v[0]; S.ry(a, 3) = v[1];
  ccp_up_pt<PANDA, 4>(A.link[4], S.s(a, 4), S.c(a, 4), v);
  S.rx(a, 4) = v[0]; S.ry(a, 4) = v[1];
  ccp_up_pt<PANDA, 5>(A.link[5], S.s(a, 5), S.c(a, 5), v);
  S.rx(a, 5) = v[0]; S.ry(a, 5) = v[1];
  ccp_up_pt<PANDA, 6>(A.link[6], S.s(a, 6), S.c(a, 6), v);
  S.rx(a, 6) = v[0]; S.ry(a, 6) = v[1];
  v[2] -= A.fl;  // frame 7' is the EE frame up to this shift along z
}
// residual of pair p (arm p + 1 against arm 0) from tc = v and the two chain quaternions
template <int K, int PANDA>
CCP_HD void ccp_fwd_pair(const ccp_pair_ref& ref, const double* v, const double* qa, const double* q0, double* tc, double* qc,
                         double* d, double* e, double& e2, double& sv2) {
  tc[0] = v[0]; tc[1] = v[1]; tc[2] = v[2];
  ccp_qmul_conj_left(qa, q0, qc);
  if (PANDA && K != 2) {  // K == 2: the exact factor 1/64 is already in arm 0's start quaternion (qrel_scaled)
    qc[0] *= CCP_PANDA_QSCALE; qc[1] *= CCP_PANDA_QSCALE; qc[2] *= CCP_PANDA_QSCALE; qc[3] *= CCP_PANDA_QSCALE;
  }
  ccp_qmul_conj_right(qc, ref.q0, d);
  const double* t0 = ref.t0;
  const double ex = tc[0] - t0[0], ey = tc[1] - t0[1], ez = tc[2] - t0[2];
  e[0] = ex; e[1] = ey; e[2] = ez;
  e2 = CCP_FMA(ex, ex, CCP_FMA(ey, ey, ez * ez));
  sv2 = CCP_FMA(d[1], d[1], CCP_FMA(d[2], d[2], d[3] * d[3]));
}

template <int K, int PANDA, class SC, class XT>
CCP_HD void ccp_forward(const ccp_model& M, const XT& x, SC& S, ccp_fwd<K>& F) {
  double q[K][4];
#pragma unroll
  for (int a = 0; a < K; ++a) {
    const ccp_arm& A = M.arm[a];
    if (K == 2 && a == 1) {
      // only the RELATIVE rotation conj(q_1) q_0 is used: arm 0's chain starts from conj(qwb_1) qwb_0 (below) and
      // arm 1's from the identity, whose product with the first link quaternion is that quaternion itself
      ccp_fwd_link_quat<PANDA, 0, true>(A, a, x, q[a], S);
    } else {
      const double* s0 = (K == 2) ? (PANDA ? M.arm[1].qrel_scaled : M.arm[1].qrel) : A.qwb;
      q[a][0] = s0[0]; q[a][1] = s0[1]; q[a][2] = s0[2]; q[a][3] = s0[3];
      ccp_fwd_link_quat<PANDA, 0>(A, a, x, q[a], S);
    }
    ccp_fwd_quat_links_1_6<PANDA>(A, a, x, q[a], S);
  }
  double r[3];
  ccp_fwd_down_arm0<PANDA>(M.arm[0], 0, S, r);
#pragma unroll
  for (int a = 1; a < K; ++a) {
    double v[3];
    ccp_fwd_up_arm<PANDA>(M.arm[a], a, S, r, v);
    ccp_fwd_pair<K, PANDA>(M.ref[a - 1], v, q[a], q[0], F.tc[a - 1], F.qc[a - 1], F.d[a - 1], F.e[a - 1], F.e2[a - 1],
                           F.sv2[a - 1]);
  }
}

// Loop test of project(): `(f0 > tol1) || (f1 > tol2)` for any pair (ConstraintFunction.h:68), on the squares.
template <int K>
CCP_HD bool ccp_needs_step(const ccp_model& M, const ccp_fwd<K>& F) {
  bool any = false;
#pragma unroll
  for (int a = 0; a < K - 1; ++a)
    any = any || (F.e2[a] > M.tol_p2) || (F.sv2[a] > M.tan2_r * (F.d[a][0] * F.d[a][0]));
  return any;
}
// Success test of project() (ConstraintFunction.h:75): norm1 is the 0/1 flag `f0 > tol1`, so
// `norm1 < tol1` means f0 <= tol1; norm2 = f1 must be STRICTLY below tol2.  NaNs fail.
template <int K>
CCP_HD bool ccp_converged(const ccp_model& M, const ccp_fwd<K>& F) {
  bool all = true;
#pragma unroll
  for (int a = 0; a < K - 1; ++a)
    all = all && (F.e2[a] <= M.tol_p2) && (F.sv2[a] < M.tan2_r * (F.d[a][0] * F.d[a][0]));
  return all;
}
// isSatisfied (ConstraintFunction.h:114-120): finite, f0 <= tol1, f1 <= tol2.
template <int K>
CCP_HD bool ccp_is_satisfied(const ccp_model& M, const double* f) {
  bool all = true;
#pragma unroll
  for (int a = 0; a < K - 1; ++a) {
    double f0 = f[2 * a], f1 = f[2 * a + 1];
    all = all && (f0 - f0 == 0.0) && (f1 - f1 == 0.0) && (f0 <= M.tol_p) && (f1 <= M.tol_r);
  }
  return all;
}
// jointValid (ConstraintFunction.h:43-55).  Branch-free (no short-circuit: the compiler would otherwise predicate a
// chain of conditional loads), two independent accumulation chains per arm: the epilogue runs with a few lanes and
// the rest of the warp waiting behind it, so its instruction count and latency both matter.
template <int K, class XT>
CCP_HD bool ccp_joint_valid(const ccp_model& M, const XT& x) {
  unsigned bad = 0u;
#pragma unroll
  for (int a = 0; a < K; ++a) {
    unsigned lo = 0u, hi = 0u;
#pragma unroll
    for (int i = 0; i < CCPC_DOF; ++i) {
      const double v = x[a * CCPC_DOF + i];
      lo |= (unsigned)(v < M.lbm[i]);
      hi |= (unsigned)(v > M.ubm[i]);
    }
    bad |= lo | hi;
  }
  return bad == 0u;
}

// ------------------------------------------------------------------------------------------
// Jacobian pass.  J0[p][row][i]:  d f_{2p+row} / d q(arm 0,   joint i);
//                 Ja[p][row][i]: -d f_{2p+row} / d q(arm p+1, joint i)  — stored with the opposite sign: the
//                 Gram matrix only sees products of two Ja entries and the update adds instead of subtracting,
//                 so no negation is ever executed (bit-identical to storing the signed value).
// ------------------------------------------------------------------------------------------
template <int K>
struct ccp_jac {
  double Ja[K - 1][2][CCPC_DOF];
  double J0[K - 1][2][CCPC_DOF];
  CCP_HD double& a(int p, int r, int i) { return Ja[p][r][i]; }
  CCP_HD double& z(int p, int r, int i) { return J0[p][r][i]; }
  CCP_HD double a(int p, int r, int i) const { return Ja[p][r][i]; }
  CCP_HD double z(int p, int r, int i) const { return J0[p][r][i]; }
};

template <int PANDA, int I, bool ARM0, class SC, class JT>
CCP_HD void ccp_jac_link(const ccp_arm& A, int a, int p, const SC& S, double* w, double* m, JT& J) {
  // joint I sees (r, w, m) in frame I:  d f0 / d q = +-(r x w)_z,  d f1 / d q = +-m_z  (+ on arm 0);
  // r = lever arm to the EE-0 origin, left behind by the forward pass
  if (ARM0 && I == 6) {  // the EE-0 origin lies on joint 7's own axis
    J.z(p, 0, I) = 0.0;
    J.z(p, 1, I) = m[2];
  } else if (ARM0) {
    const double cz = CCP_FMA(S.rx(a, I), w[1], -(S.ry(a, I) * w[0]));
    J.z(p, 0, I) = cz;
    J.z(p, 1, I) = m[2];
  } else {
    const double cz = CCP_FMA(S.rx(a, I), w[1], -(S.ry(a, I) * w[0]));
    J.a(p, 0, I) = cz;
    J.a(p, 1, I) = m[2];
  }
  if (I > 0) {
    const double s = S.s(a, I), c = S.c(a, I);
    ccp_down_vec<PANDA, I>(A.link[I], s, c, w);
    ccp_down_vec<PANDA, I>(A.link[I], s, c, m);
  }
}
template <int PANDA, bool ARM0, class SC, class JT>
CCP_HD void ccp_jac_arm(const ccp_arm& A, int a, int p, const SC& S, double* w, double* m, JT& J) {
  ccp_jac_link<PANDA, 6, ARM0>(A, a, p, S, w, m, J);
  ccp_jac_link<PANDA, 5, ARM0>(A, a, p, S, w, m, J);
  ccp_jac_link<PANDA, 4, ARM0>(A, a, p, S, w, m, J);
  ccp_jac_link<PANDA, 3, ARM0>(A, a, p, S, w, m, J);
  ccp_jac_link<PANDA, 2, ARM0>(A, a, p, S, w, m, J);
  ccp_jac_link<PANDA, 1, ARM0>(A, a, p, S, w, m, J);
  ccp_jac_link<PANDA, 0, ARM0>(A, a, p, S, w, m, J);
}

template <int K, int PANDA, class SC, class JT>
CCP_HD void ccp_jacobian(const ccp_model& M, const SC& S, const ccp_fwd<K>& F, JT& J) {
#pragma unroll
  for (int p = 0; p < K - 1; ++p) {
    const int a = p + 1;
    const double* e = F.e[p];
    const double* d = F.d[p];
    // arm a: frame EE_a (= frame 7', the yaw rides on joint 7) -> ... -> frame 1
    {
      const ccp_arm& A = M.arm[a];
      double w[3] = {e[0], e[1], e[2]};
      double m[3] = {d[1], d[2], d[3]};
      ccp_jac_arm<PANDA, false>(A, a, p, S, w, m, J);
    }
    // arm 0: e rotated into EE_0's frame by R_c^T; vec d rotated the same way is the vector part of
    // conj(q_c) d q_c = conj(q_ref) q_c (d = q_c conj(q_ref)): one product with a constant instead of a rotation
    {
      const ccp_arm& A = M.arm[0];
      double w[3] = {e[0], e[1], e[2]};
      ccp_qrot_inv(F.qc[p], w);
      double dq[4];
      ccp_qmul_conj_left(M.ref[p].q0, F.qc[p], dq);
      double m[3] = {dq[1], dq[2], dq[3]};
      ccp_jac_arm<PANDA, true>(A, 0, p, S, w, m, J);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Newton step: x <- x - step * J^T (J J^T)^-1 f         (ConstraintFunction.h:71)
// ------------------------------------------------------------------------------------------
// right-hand side D^-1 f = (f0^2, s |vec d| f1) of one pair: the only place sqrt and atan2 are needed
CCP_HD void ccp_step_rhs(double e2, double sv2, double d0, double* rhs) {
  const double sv = sqrt(sv2);
  const double atn = ccp_atan2_pos(sv, fabs(d0));
  const double h = sv * (atn + atn);
  rhs[0] = e2;
  rhs[1] = (d0 < 0.0) ? -h : h;
}
// 2x2: Cramer's rule, one division.  A vanished row (g_kk = 0) or dependent rows (det <= 0) drop row 1 / the
// vanished row, exactly what the L D L^T path of ccp_newton_step does.
CCP_HD void ccp_solve_2x2(double g00, double g01, double g11, const double* rhs, double* y) {
  const double det = CCP_FMA(g00, g11, -(g01 * g01));
  const bool k0 = g00 > 0.0, k1 = g11 > 0.0;
  if (k0 && k1 && det > 0.0) {
    const double inv = 1.0 / det;
    y[0] = CCP_FMA(g11, rhs[0], -(g01 * rhs[1])) * inv;
    y[1] = CCP_FMA(g00, rhs[1], -(g01 * rhs[0])) * inv;
  } else if (k0) {
    y[0] = rhs[0] / g00;
    y[1] = 0.0;
  } else if (k1) {
    y[0] = 0.0;
    y[1] = rhs[1] / g11;
  } else {
    y[0] = 0.0;
    y[1] = 0.0;
  }
}
// one arm's share of the Gram entry of two gradient rows (7 columns), the accumulation order of ccp_newton_step
CCP_HD double ccp_row_dot7(const double* a, const double* b) {
  double acc = a[0] * b[0];
#pragma unroll
  for (int i = 1; i < CCPC_DOF; ++i) acc = CCP_FMA(a[i], b[i], acc);
  return acc;
}

// m x m (lower triangle of G): L D L^T with row dropping — a vanished or dependent row (pivot <= 0) is dropped, its
// multiplier is 0.  Lm is unit lower triangular, D the pivots.
template <int m>
CCP_HD void ccp_solve_ldlt(const double (&G)[m][m], const double* rhs, double* y) {
  double D[m], invD[m];
  double Lm[m][m];
#pragma unroll
  for (int k = 0; k < m; ++k) {
    double dk = G[k][k];
#pragma unroll
    for (int j = 0; j < k; ++j) dk = CCP_FMA(-(Lm[k][j] * D[j]), Lm[k][j], dk);
    const bool keep = dk > 0.0;
    D[k] = keep ? dk : 0.0;
    invD[k] = keep ? 1.0 / dk : 0.0;
#pragma unroll
    for (int i = k + 1; i < m; ++i) {
      double v = G[i][k];
#pragma unroll
      for (int j = 0; j < k; ++j) v = CCP_FMA(-(Lm[i][j] * D[j]), Lm[k][j], v);
      Lm[i][k] = v * invD[k];
    }
  }
  // forward substitution L z = f, scale by D^-1, back substitution L^T y = z
#pragma unroll
  for (int k = 0; k < m; ++k) {
    double v = rhs[k];
#pragma unroll
    for (int j = 0; j < k; ++j) v = CCP_FMA(-Lm[k][j], y[j], v);
    y[k] = v;
  }
#pragma unroll
  for (int k = 0; k < m; ++k) y[k] = y[k] * invD[k];
#pragma unroll
  for (int k = m - 1; k >= 0; --k) {
    double v = y[k];
#pragma unroll
    for (int j = k + 1; j < m; ++j) v = CCP_FMA(-Lm[j][k], y[j], v);
    y[k] = (invD[k] != 0.0) ? v : 0.0;
  }
}

template <int K, class JT, class XT>
CCP_HD void ccp_newton_step(const ccp_model& M, const ccp_fwd<K>& F, const JT& J, XT& x) {
  constexpr int m = 2 * (K - 1);
  double rhs[m];
#pragma unroll
  for (int p = 0; p < K - 1; ++p) ccp_step_rhs(F.e2[p], F.sv2[p], F.d[p][0], rhs + 2 * p);
  double G[m][m];
  // Gram matrix of the rows, lower triangle.  Rows of the same pair share both arms' columns; rows of
  // different pairs only share arm 0's columns.  The two arms' partial sums are independent chains.
#pragma unroll
  for (int p = 0; p < K - 1; ++p)
#pragma unroll
    for (int ri = 0; ri < 2; ++ri)
#pragma unroll
      for (int pp = 0; pp <= p; ++pp)
#pragma unroll
        for (int rj = 0; rj < 2; ++rj) {
          const int I = 2 * p + ri, Jx = 2 * pp + rj;
          if (Jx > I) continue;
          double acc = J.z(p, ri, 0) * J.z(pp, rj, 0);
#pragma unroll
          for (int i = 1; i < CCPC_DOF; ++i) acc = CCP_FMA(J.z(p, ri, i), J.z(pp, rj, i), acc);
          if (pp == p) {
            double acca = J.a(p, ri, 0) * J.a(p, rj, 0);
#pragma unroll
            for (int i = 1; i < CCPC_DOF; ++i) acca = CCP_FMA(J.a(p, ri, i), J.a(p, rj, i), acca);
            acc = acc + acca;
          }
          // damped least squares J J^T + lambda^2 I, written for the unnormalised rows (J = D g):
          // g g^T + lambda^2 D^-2 with D^-2 = diag(f0^2, |vec d|^2).  lambda^2 = 0 (the reference) leaves acc as is.
          if (I == Jx) acc = CCP_FMA(M.damping, (ri == 0) ? F.e2[p] : F.sv2[p], acc);
          G[I][Jx] = acc;
        }
  double y[m];
  if (m == 2) {
    ccp_solve_2x2(G[0][0], G[1][0], G[1][1], rhs, y);
  } else {
    ccp_solve_ldlt<m>(G, rhs, y);
  }
  // x -= step * J^T y
#pragma unroll
  for (int i = 0; i < CCPC_DOF; ++i) {
    double dx = 0.0;
#pragma unroll
    for (int p = 0; p < K - 1; ++p) {
      dx = CCP_FMA(J.z(p, 0, i), y[2 * p], dx);
      dx = CCP_FMA(J.z(p, 1, i), y[2 * p + 1], dx);
    }
    x[i] = CCP_FMA(-M.step, dx, x[i]);
  }
#pragma unroll
  for (int p = 0; p < K - 1; ++p)
#pragma unroll
    for (int i = 0; i < CCPC_DOF; ++i) {
      double dx = CCP_FMA(J.a(p, 1, i), y[2 * p + 1], J.a(p, 0, i) * y[2 * p]);
      x[(p + 1) * CCPC_DOF + i] = CCP_FMA(M.step, dx, x[(p + 1) * CCPC_DOF + i]);  // Ja holds -J
    }
}

// opt-in joint-limit clamping of an iterate (off in the reference: ConstraintFunction.h:57-82 never clamps)
template <int K, class XT>
CCP_HD void ccp_clamp_to_limits(const ccp_model& M, XT& x) {
#pragma unroll
  for (int a = 0; a < K; ++a)
#pragma unroll
    for (int i = 0; i < CCPC_DOF; ++i) {
      double v = x[a * CCPC_DOF + i];
      v = (v < M.lb[i]) ? M.lb[i] : v;
      v = (v > M.ub[i]) ? M.ub[i] : v;
      x[a * CCPC_DOF + i] = v;
    }
}

// Dense m x n Jacobian (row-major), the layout jacobian() returns: J = D g with D = diag(1/f0, s/|vec d|)
// (a row whose residual vanishes is 0), arm-a entries with their sign restored.
template <int K>
CCP_HD void ccp_jac_dense(const ccp_fwd<K>& F, const ccp_jac<K>& J, double* out) {
  constexpr int m = 2 * (K - 1), n = CCPC_DOF * K;
#pragma unroll
  for (int i = 0; i < m * n; ++i) out[i] = 0.0;
#pragma unroll
  for (int p = 0; p < K - 1; ++p) {
    const double f0 = sqrt(F.e2[p]), sv = sqrt(F.sv2[p]);
    double sc[2];
    sc[0] = (f0 > 0.0) ? 1.0 / f0 : 0.0;
    sc[1] = (sv > 0.0) ? 1.0 / sv : 0.0;
    sc[1] = (F.d[p][0] < 0.0) ? -sc[1] : sc[1];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int i = 0; i < CCPC_DOF; ++i) {
        out[(2 * p + r) * n + i] = J.J0[p][r][i] * sc[r];
        out[(2 * p + r) * n + (p + 1) * CCPC_DOF + i] = -(J.Ja[p][r][i] * sc[r]);
      }
  }
}

// ------------------------------------------------------------------------------------------
// project(): the whole Newton loop for ONE sample (ConstraintFunction.h:57-82).
// Used as-is by the host build; the CUDA kernel runs the same three calls inside its
// lane-refill loop.
// ------------------------------------------------------------------------------------------
template <int K, int PANDA>
CCP_HD void ccp_project_one(const ccp_model& M, double* x, double* f_out, int32_t* iters, bool* converged,
                            bool* ok) {
  ccp_fwd<K> F;
  ccp_jac<K> J;
  ccp_sc_local<K> S;
  int32_t it = 0;
  ccp_forward<K, PANDA>(M, x, S, F);
  while (ccp_needs_step<K>(M, F) && it < M.max_iter) {
    ++it;
    ccp_jacobian<K, PANDA>(M, S, F, J);
    ccp_newton_step<K>(M, F, J, x);
    if (M.clamp) ccp_clamp_to_limits<K>(M, x);
    ccp_forward<K, PANDA>(M, x, S, F);
  }
  const bool conv = ccp_converged<K>(M, F);
  ccp_residual<K>(F, f_out, nullptr);
  *iters = it;
  *converged = conv;
  *ok = conv && ccp_joint_valid<K>(M, x);
}

// setInitialPosition (ConstraintFunction.h:31-40): reference chain = chain at q_start with an
// identity reference.
template <int K, int PANDA>
CCP_HD void ccp_reference_chain(ccp_model& M, const double* q_start) {
#pragma unroll
  for (int p = 0; p < K - 1; ++p) {
    M.ref[p].t0[0] = M.ref[p].t0[1] = M.ref[p].t0[2] = 0.0;
    M.ref[p].q0[0] = 1.0;
    M.ref[p].q0[1] = M.ref[p].q0[2] = M.ref[p].q0[3] = 0.0;
  }
  ccp_fwd<K> F;
  ccp_sc_local<K> S;
  ccp_forward<K, PANDA>(M, q_start, S, F);
#pragma unroll
  for (int p = 0; p < K - 1; ++p) {
#pragma unroll
    for (int k = 0; k < 3; ++k) M.ref[p].t0[k] = F.tc[p][k];
#pragma unroll
    for (int k = 0; k < 4; ++k) M.ref[p].q0[k] = F.qc[p][k];
  }
}

// ------------------------------------------------------------------------------------------
// Single-arm kinematics in the arm's BASE frame (RobotModel API, panda_rbdl.cpp:9-42).
// T: row-major 3x4 [R|p].  Jac: 6x7 row-major, rows [linear; angular] (panda_rbdl.cpp:16-19).
// ------------------------------------------------------------------------------------------
// One link of the matrix recursion: o += R t ; R <- R Rx(alpha) Rz(theta).  PANDA >= 1 compiles the stock alpha pattern
// in (Rx is a signed permutation of columns, the structurally zero translation component is skipped), PANDA = 2 also the
// stock table's zero a / d entries — the batched IK iteration (ccp_ik.h) is 8 % shorter with it.  PANDA = 0 is the
// generic code, which the public FK / Jacobian entry points always use.
template <int PANDA, int I>
CCP_HD void ccp_arm_fk_link(const ccp_link& L, double s, double c, double* R, double* o) {
  constexpr int sg = ccp_panda_sgn<I>::value;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double acc = o[r];
    if (!PANDA) {
      acc = CCP_FMA(R[3 * r], L.tx, CCP_FMA(R[3 * r + 1], L.ty, CCP_FMA(R[3 * r + 2], L.tz, acc)));
    } else {
      if (!(PANDA == 2 && ccp_stock_zero<I>::d)) {
        if (sg == 0) acc = CCP_FMA(R[3 * r + 2], L.tz, acc);  // alpha = 0: t = (a, 0, d)
        else acc = CCP_FMA(R[3 * r + 1], L.ty, acc);         // alpha = +-pi/2: t = (a, -+d, 0)
      }
      if (!(PANDA == 2 && ccp_stock_zero<I>::a)) acc = CCP_FMA(R[3 * r], L.tx, acc);
    }
    o[r] = acc;
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double c0 = R[3 * r], c1 = R[3 * r + 1], c2 = R[3 * r + 2];
    double n1, n2;  // columns 1, 2 of R Rx(alpha)
    if (!PANDA) {
      n1 = CCP_FMA(c1, L.ca, c2 * L.sa);
      n2 = CCP_FMA(c2, L.ca, -(c1 * L.sa));
    } else if (sg > 0) {
      n1 = c2; n2 = -c1;
    } else if (sg < 0) {
      n1 = -c2; n2 = c1;
    } else {
      n1 = c1; n2 = c2;
    }
    R[3 * r] = CCP_FMA(c0, c, n1 * s);
    R[3 * r + 1] = CCP_FMA(n1, c, -(c0 * s));
    R[3 * r + 2] = n2;
  }
}
template <int PANDA, int I>
CCP_HD void ccp_arm_fk_links(const ccp_arm& A, const double* q, double* R, double* o, double (*zs)[3], double (*os)[3]) {
  const ccp_link& L = A.link[I];
  double s, c;
  ccp_sincos((PANDA == 2) ? q[I] : q[I] + L.qoff, &s, &c);  // stock: no theta calibration
  ccp_arm_fk_link<PANDA, I>(L, s, c, R, o);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    zs[I][r] = R[3 * r + 2];
    os[I][r] = o[r];
  }
  if (I + 1 < CCPC_DOF) ccp_arm_fk_links<PANDA, (I + 1 < CCPC_DOF ? I + 1 : I)>(A, q, R, o, zs, os);
}
template <int PANDA>
CCP_HD void ccp_arm_fk_t(const ccp_arm& A, const double* q, double* T, double* Jac) {
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double o[3] = {0, 0, 0};
  double zs[CCPC_DOF][3], os[CCPC_DOF][3];
  ccp_arm_fk_links<PANDA, 0>(A, q, R, o, zs, os);
  double p[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) p[r] = CCP_FMA(R[3 * r + 2], A.fl, o[r]);
  if (T) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      double c0 = R[3 * r], c1 = R[3 * r + 1];
      T[4 * r] = CCP_FMA(c0, A.cphi, c1 * A.sphi);
      T[4 * r + 1] = CCP_FMA(c1, A.cphi, -(c0 * A.sphi));
      T[4 * r + 2] = R[3 * r + 2];
      T[4 * r + 3] = p[r];
    }
  }
  if (Jac) {
#pragma unroll
    for (int i = 0; i < CCPC_DOF; ++i) {
      double lx = p[0] - os[i][0], ly = p[1] - os[i][1], lz = p[2] - os[i][2];
      Jac[0 * 7 + i] = CCP_FMA(zs[i][1], lz, -(zs[i][2] * ly));
      Jac[1 * 7 + i] = CCP_FMA(zs[i][2], lx, -(zs[i][0] * lz));
      Jac[2 * 7 + i] = CCP_FMA(zs[i][0], ly, -(zs[i][1] * lx));
      Jac[3 * 7 + i] = zs[i][0];
      Jac[4 * 7 + i] = zs[i][1];
      Jac[5 * 7 + i] = zs[i][2];
    }
  }
}
CCP_HD void ccp_arm_fk(const ccp_arm& A, const double* q, double* T, double* Jac) { ccp_arm_fk_t<0>(A, q, T, Jac); }

// KinematicChainSpace::enforceBounds (KinematicChain.h:118-130): fmod wrap into [-pi, pi).
// fmod is exact in IEEE arithmetic, so host and device agree bit for bit.
CCP_HD double ccp_wrap_pi(double v) {
  const double PI = 3.14159265358979323846, TWO_PI = 2.0 * 3.14159265358979323846;
  double w = fmod(v, TWO_PI);
  if (w < -PI) w += TWO_PI;
  else if (w >= PI) w -= TWO_PI;
  return w;
}

// ------------------------------------------------------------------------------------------
// Counter-based seed stream: 53-bit uniforms from splitmix64 of (seed, sample, joint).
// ------------------------------------------------------------------------------------------
CCP_HD uint64_t ccp_mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
CCP_HD double ccp_uniform01(uint64_t seed, uint64_t sample, uint32_t lane) {
  uint64_t h = ccp_mix64(seed ^ 0xD1B54A32D192ED03ull);
  h = ccp_mix64(h + sample * 0x9E3779B97F4A7C15ull);
  h = ccp_mix64(h + (uint64_t)lane);
  return (double)(h >> 11) * 0x1.0p-53;
}
// sampleUniform: x_j = lb_j + u (ub_j - lb_j)   (OMPL RealVectorStateSampler over
// KinematicChain.h:77-99 bounds)
CCP_HD double ccp_seed_uniform(const ccp_model& M, uint64_t seed, uint64_t sample, int j) {
  const int i = j % CCPC_DOF;
  return CCP_FMA(ccp_uniform01(seed, sample, (uint32_t)j), M.ub[i] - M.lb[i], M.lb[i]);
}

// ------------------------------------------------------------------------------------------
// discreteGeodesic building blocks (jy_ProjectedStateSpace.cpp:32-96)
// ------------------------------------------------------------------------------------------
// KinematicChainSpace::interpolate, one joint (KinematicChain.h:145-171): the short way round when
// |to - from| > pi, wrapped back into [-pi, pi].
CCP_HD double ccp_interpolate_joint(double from, double to, double t) {
  const double PI = 3.14159265358979323846;
  double diff = to - from;
  if (fabs(diff) <= PI) return CCP_FMA(diff, t, from);
  if (diff > 0.0) diff = 2.0 * PI - diff;
  else diff = -2.0 * PI - diff;
  double v = CCP_FMA(-diff, t, from);
  if (v > PI) v -= 2.0 * PI;
  else if (v < -PI) v += 2.0 * PI;
  return v;
}
// RealVectorStateSpace::distance (Euclidean), sequential accumulation
template <int N, class XA, class XB>
CCP_HD double ccp_distance(const XA& a, const XB& b) {
  double acc = 0.0;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const double d = a[j] - b[j];
    acc = CCP_FMA(d, d, acc);
  }
  return sqrt(acc);
}

// Bookkeeping of one geodesic after the projection of `scratch` finished (jy_ProjectedStateSpace.cpp:65-90).
// Returns 0 = keep walking (scratch accepted, `dist` updated), 1 = stop: reached (dist < delta after accepting),
// 2 = stop: break (projection failed / deviated / wandered / no progress).  On 0 and 1 the caller stores scratch.
struct ccp_geo_state {
  double dist;   // distance(previous, to)
  double total;  // accumulated arc length
  double max;    // dist0 * lambda
};
template <int N, class XP, class XS, class XT>
CCP_HD int ccp_geodesic_advance(ccp_geo_state& g, bool projected_ok, const XP& previous, const XS& scratch, const XT& to,
                                double delta, double lambda) {
  if (!projected_ok) return 2;                                  // not on manifold
  const double step = ccp_distance<N>(previous, scratch);
  if (step > lambda * delta) return 2;                          // deviated
  g.total += step;
  if (g.total > g.max) return 2;                                // wandered too far
  const double newDist = ccp_distance<N>(scratch, to);
  if (newDist >= g.dist) return 2;                              // no closer than before
  g.dist = newDist;
  return (g.dist >= delta) ? 0 : 1;
}
