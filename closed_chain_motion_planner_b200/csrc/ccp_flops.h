// ccp_flops.h — FROZEN algorithmic FLOP counts of the closed-chain projection.
//
// FMA = 2 FLOP, add/mul = 1; sincos / atan2 / sqrt / div are NOT expanded (they are counted as
// calls and reported separately).  The counts describe the ANALYTIC algorithm with general 3x3
// products — not the reference's finite-difference Jacobian (x85 more work, not counted) and not
// the (cheaper) vector/quaternion recursions the kernel actually executes; the executed FP64
// instruction mix is reported separately from ncu (profiles/).
//
// Per Newton iteration, K = 2 arms:
//   FK 2 x [7 x (4 + 45 + 18) + 6 + 18] = 986 ; Jacobian columns 14 x 12 = 168 ; relative pose 63 ;
//   residual ~100 ; gradient rows ~170 ; J J^T 81 ; 2x2 factor+solve ~15 ; J^T y + update 70
//   => 1650 FLOP + 14 sincos + 1 atan2 + 4 sqrt + ~6 div
// K = 3 arms: 3 x 493 + 336 + 2 x (63 + 100 + 170) + 410 + 60 + 210 => 3160 FLOP + 21 sincos + 2 atan2
// Tail (the final function() evaluation of every projection): one FK set + relative pose + residual.
#pragma once

#define CCP_FLOPS_ITER_K2 1650.0
#define CCP_FLOPS_TAIL_K2 1150.0
#define CCP_FLOPS_ITER_K3 3160.0
#define CCP_FLOPS_TAIL_K3 1805.0
#define CCP_SINCOS_PER_EVAL_K2 14
#define CCP_SINCOS_PER_EVAL_K3 21
#define CCP_ATAN2_PER_EVAL_K2 1
#define CCP_ATAN2_PER_EVAL_K3 2

// Batched pose IK (ccp_ik.h), per damped-Newton iteration of ONE arm, same conventions:
//   FK with general 3x3 products 7 x (18 + 36) + 6 + 18 = 402 ; geometric Jacobian 7 x 12 = 84 ; rotation error 55 ;
//   J J^T + lambda^2 I (21 entries x 7 FMA) 294 ; 6x6 factorisation (counted as the Cholesky it replaced; the LDL^T does 9 more multiplies and no square roots) 85 ; two triangular solves 60 ; J^T y + update 91
//   => 1070 FLOP + 7 sincos + 7 sqrt + 19 div.   Tail (the final FK + error evaluation of every solve): 540.
#define CCP_FLOPS_IK_ITER 1070.0
#define CCP_FLOPS_IK_TAIL 540.0
