/*
 * ccp.h — C ABI of the B200-native batched closed-chain constraint-projection engine.
 *
 * This is the drop-in boundary for ONE hot path of jkw0701/closed_chain_motion_planner:
 * KinematicChainConstraint::{function, jacobian, project, isSatisfied, jointValid,
 * setInitialPosition, setTolerance, setArmModels} and the PandaModel kinematics it calls.
 * Every entry point cites the reference interface it replaces (paths relative to the
 * reference root).  Plain pointers and sizes only; no C++/torch types; no exceptions.
 *
 * Conventions
 *   - K = number of arms (2 or 3), n = 7*K joints per state, m = 2*(K-1) residual rows.
 *   - "dev" pointers are CUDA device pointers on the handle's device; "host" pointers are
 *     ordinary host memory (pinned or pageable).  `stream` is a cudaStream_t passed as void*
 *     (NULL = the legacy default stream).  Device-pointer calls are asynchronous on `stream`;
 *     host-pointer calls (`*_host`) return after the results are in the caller's buffers.
 *   - layout: CCP_LAYOUT_AOS = states stored state-major, double[count][n] (what the reference's
 *     OMPL RealVectorStateSpace::StateType::values look like when gathered);
 *     CCP_LAYOUT_SOA = joint-major, double[n][count] (coalesced, the engine's native layout).
 *   - All functions return CCP_OK (0) or a negative ccp_status; ccp_last_error() gives text.
 *   - There is no CPU fallback: without a CUDA device ccp_create fails with CCP_ERR_CUDA.
 *   - Threads: host-pointer calls on one handle may come from several threads; they are serialised inside (the
 *     reference serialises its callers with graphMutex_, stefanBiPRM.cpp:280,383,449).  Device-pointer projection
 *     calls on one handle (ccp_project_batch*, ccp_sample_project_batch*, ccp_project_flush) share its launch pipeline
 *     and must come from one thread at a time; use one handle per planner thread or stream otherwise.  Handles are
 *     independent of one another.
 */
#ifndef CCP_H_
#define CCP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CCP_MAX_ARMS 3
#define CCP_DOF 7

typedef enum ccp_status {
  CCP_OK = 0,
  CCP_ERR_INVALID = -1,   /* bad argument (null pointer, n_arms not 2/3, tolerance <= 0 ...) */
  CCP_ERR_CUDA = -2,      /* CUDA runtime error, or no CUDA device */
  CCP_ERR_STATE = -3,     /* call order (e.g. project before set_reference) */
  CCP_ERR_NCCL = -4
} ccp_status;

#define CCP_MODEL_NO_STOCK 1

typedef enum ccp_layout { CCP_LAYOUT_AOS = 0, CCP_LAYOUT_SOA = 1 } ccp_layout;

/* One arm.  Replaces ArmModel{rbdl_model, t_wb} (kinematics/panda_model.h:7-23) and
 * PandaModel::initModel(dh) (src/kinematics/panda_rbdl.cpp:73-148): modified (Craig) DH rows
 * plus the optional 7x4 calibration offsets already ADDED in (a, d, theta, alpha).            */
typedef struct ccp_arm_desc {
  double dh_a[CCP_DOF];            /* panda_rbdl.cpp:98  (+ dh.col(0)) */
  double dh_d[CCP_DOF];            /* panda_rbdl.cpp:99  (+ dh.col(1)) */
  double dh_alpha[CCP_DOF];        /* panda_rbdl.cpp:97  (+ dh.col(3)) */
  double dh_theta_offset[CCP_DOF]; /* dh.col(2), panda_rbdl.cpp:94,119 */
  double t_wb[12];                 /* base frame in world, row-major 3x4 [R|p]; grasping_point.cpp:11-16 */
  double flange;                   /* 0.107 m, panda_rbdl.cpp:125 */
  double ee_yaw;                   /* -pi/4,   panda_rbdl.cpp:31  */
} ccp_arm_desc;

/* The whole closed chain: arm[0] is the constraint's first arm (arm_models_[0]),
 * arm[a>=1] closes the chain against arm[0] (ConstraintFunction.h:89-92).          */
typedef struct ccp_model_desc {
  int32_t n_arms;                  /* 2 (reference) or 3 (21-DoF extension) */
  int32_t flags;                   /* 0, or CCP_MODEL_NO_STOCK: do not use the kernels specialised to the reference's stock
                                      Panda table and base frames even when the model matches them (tests, tuning)       */
  ccp_arm_desc arm[CCP_MAX_ARMS];
  double lb[CCP_DOF];              /* ConstraintFunction.h:27 */
  double ub[CCP_DOF];              /* ConstraintFunction.h:28 */
} ccp_model_desc;

typedef struct ccp_options {
  double step;          /* 0.30  ConstraintFunction.h:71 */
  int32_t max_iter;     /* 250   ConstraintFunction.h:26,68 */
  int32_t clamp;        /* 0     1 = clamp every iterate to [lb, ub]; the reference never clamps (parity: keep 0) */
  double joint_margin;  /* 1e-3  ConstraintFunction.h:45 */
  double damping;       /* 0     lambda^2 of a damped-least-squares step; the reference's step is the undamped
                                 minimum-norm one, JacobiSVD::solve at ConstraintFunction.h:71 (parity: keep 0) */
} ccp_options;

typedef struct ccp_handle ccp_handle;

/* Fill `d` with the stock Panda constants of the reference (panda_rbdl.cpp:97-99,125,31;
 * ConstraintFunction.h:27-28) for `n_arms` arms whose base frames are
 * grasping_point::t_wb[arm_index[i]] (grasping_point.cpp:11-20: 0=left, 1=right, 2=top). */
int ccp_default_model(int32_t n_arms, const int32_t* arm_index, ccp_model_desc* d);

/* ≙ constructing KinematicChainConstraint(links) + setArmModels (ConstraintFunction.h:24-29,122-126)
 * with PandaModel::initModel for every arm.  Tolerances start at the reference's 1e-3 / 5e-3
 * (ConstrainedPlanningCommon.cpp:120-121), options at step 0.30 / 250 / 1e-3.             */
int ccp_create(const ccp_model_desc* model, int32_t device, ccp_handle** out);
void ccp_destroy(ccp_handle* h);
const char* ccp_last_error(const ccp_handle* h); /* h may be NULL: error of the last failed ccp_create */
/* CUDA devices visible to the process (0 without a driver): for a host that does not link CUDA itself. */
int ccp_device_count(void);
int ccp_n_arms(const ccp_handle* h);
int ccp_device(const ccp_handle* h);

/* ≙ setInitialPosition (ConstraintFunction.h:31-40): q_start is a HOST array of 7*K doubles.
 * Computes, on the device, the grasp-closure reference pose(s) init_chain_.                */
int ccp_set_reference(ccp_handle* h, const double* q_start_host);
/* Reads back init_chain_ for pair a (arm a+1 against arm 0): t0[3] translation and
 * q0[4] = (w,x,y,z) unit quaternion of the rotation.                                        */
int ccp_get_reference(const ccp_handle* h, int32_t pair, double* t0_host, double* q0_host);
/* ≙ setTolerance (ConstraintFunction.h:104-112); CCP_ERR_INVALID when either is <= 0
 * (the reference throws ompl::Exception).                                                   */
int ccp_set_tolerance(ccp_handle* h, double tol_position, double tol_rotation);
int ccp_set_options(ccp_handle* h, const ccp_options* opt);
int ccp_get_options(const ccp_handle* h, ccp_options* opt, double* tol_position, double* tol_rotation);
/* Tuning knob, not semantics: complete (non-pipelined) launches of at most max_count samples run on the cooperative
 * kernel (two lanes per sample, one arm each; three arms: four lanes per sample — a Newton iteration takes 1.2x / 1.7x
 * fewer cycles at 1.4x - 2x the issue slots), larger ones on the thread-per-sample kernel.  Results are bit-identical
 * either way.  0 = never, negative = the built-in default (128 / 72 samples per SM for two / three arms, the measured
 * crossovers on B200; also settable through the environment variable CCP_COOP_MAX).                                                                                           */
int ccp_set_coop_threshold(ccp_handle* h, int64_t max_count);

/* ---- batched constraint API, device pointers ------------------------------------------- */

/* ≙ function(x, out) (ConstraintFunction.h:84-102) for `count` states.
 * f_dev: double[m][count] when layout is SOA, double[count][m] when AOS.                    */
int ccp_function_batch(ccp_handle* h, const double* x_dev, int64_t count, int32_t layout,
                       double* f_dev, void* stream);

/* ≙ jacobian(x, out) (inherited OMPL default, called at ConstraintFunction.h:70), computed
 * analytically.  J_dev: AOS double[count][m][n] (row-major m x n per state);
 * SOA double[m][n][count].                                                                  */
int ccp_jacobian_batch(ccp_handle* h, const double* x_dev, int64_t count, int32_t layout,
                       double* J_dev, void* stream);

/* ≙ project(x) (ConstraintFunction.h:57-82) for `count` seeds: the Newton loop with the
 * reference's loop/exit semantics.  x_out may alias seeds (in-place, like the reference).
 * Any of ok/converged/iters/resid may be NULL.
 *   ok[i]        = project()'s return value (converged AND jointValid)
 *   converged[i] = residual under tolerance at exit (f0 <= tol1 and f1 < tol2 for every pair)
 *   iters[i]     = Newton steps taken (0..max_iter)
 *   resid        = final function(x): SOA double[m][count] / AOS double[count][m]
 *   compact_dev  = AOS double[<=count][n]: the states with ok==1 densely packed by the kernel's
 *                  epilogue (order unspecified), ready for the multi-GPU all-gather;
 *   n_ok_dev     = int64 device counter the kernel ADDS the number of ok states to (caller zeroes
 *                  it); required when compact_dev != NULL, optional otherwise.
 * The last iterate is written to x_out even on failure (reference semantics).               */
int ccp_project_batch(ccp_handle* h, const double* seeds_dev, int64_t count, int32_t layout,
                      double* x_out_dev, uint8_t* ok_dev, uint8_t* converged_dev,
                      int32_t* iters_dev, double* resid_dev, double* compact_dev,
                      int64_t* n_ok_dev, void* stream);

/* Pipelined form of ccp_project_batch for a caller that projects batch after batch on ONE stream (the planner's
 * batched sampler refilling its pool, jy_ProjectedStateSpace.cpp:10-29).  Iteration counts spread 0..max_iter, so
 * when a batch's seed list runs dry a few samples are still iterating and the rest of the GPU would idle on them;
 * this call instead PARKS them (state, iteration count, index) and returns, and the next ccp_project_batch[_pipelined]
 * / ccp_sample_project_batch[_pipelined] / ccp_project_flush on the same handle adopts them as its first work items.
 * Results are bit-identical to ccp_project_batch.  Contract:
 *   - per-seed outputs (x_out, ok, converged, iters, resid) of a pipelined call are complete only after a later
 *     non-pipelined projection call or ccp_project_flush on the same stream; the arrays must stay valid until then;
 *   - compact_dev / n_ok_dev form a STREAM: a converged state is appended to the buffer of the launch it FINISHED
 *     in (so a launch may append a few states of its predecessor), ccp_project_flush appends the last ones;
 *   - all launches of an open pipeline use one stream; setters (set_reference/tolerance/options) return
 *     CCP_ERR_STATE while the pipeline is open.                                                            */
int ccp_project_batch_pipelined(ccp_handle* h, const double* seeds_dev, int64_t count, int32_t layout,
                                double* x_out_dev, uint8_t* ok_dev, uint8_t* converged_dev,
                                int32_t* iters_dev, double* resid_dev, double* compact_dev,
                                int64_t* n_ok_dev, void* stream);
/* Completes every parked sample (compact_dev / n_ok_dev may be NULL).  No-op when nothing is parked.        */
int ccp_project_flush(ccp_handle* h, double* compact_dev, int64_t* n_ok_dev, void* stream);
/* 1 when parked samples exist, 0 otherwise.                                                                */
int ccp_project_pipeline_open(const ccp_handle* h);

/* ---- fused all-gather of the converged states (multi-GPU sampler) --------------------------------
 * pool_dev_ptrs[p] is the address, valid ON THIS DEVICE, of rank p's pool: double[world][capacity][n] in
 * peer-mapped memory (CUDA P2P / symmetric memory; pool_dev_ptrs[rank] is this rank's own pool).  While peers are
 * set, every projection call that passes n_ok_dev also stores each ok state into row rank*capacity + slot of EVERY
 * rank's pool from the kernel's epilogue (slot = the value n_ok had): the all-gather of the states is done by the
 * projection kernel itself, store by store over NVLink, overlapped with the arithmetic, and only the 8-byte counts
 * remain to be exchanged.  States beyond `capacity` are counted but not stored (overflow is visible in n_ok).
 * A peer may read this rank's rows once the projection kernel has completed (e.g. after the count exchange that
 * follows it on the same stream).  world = 0 switches the mode off.                                           */
int ccp_set_gather_peers(ccp_handle* h, int32_t world, int32_t rank, const uint64_t* pool_dev_ptrs, int64_t capacity);
/* Optional, after ccp_set_gather_peers: the address on this device of a MULTICAST mapping of the same pool (CUDA
 * multicast object over all ranks' pools: NVLS; torch symmetric memory's multicast_ptr).  The epilogue then issues one
 * multimem store per 16 bytes and the NVSwitch replicates it into every rank's pool — n/2 stores per converged state
 * whatever the world size, instead of n stores per peer.  0 switches back to per-peer stores; ccp_set_gather_peers
 * resets it.                                                                                                      */
int ccp_set_gather_multicast(ccp_handle* h, uint64_t pool_multicast_dev_ptr);
/* Stores *n_ok_dev into slot `rank` of every rank's int64[world] count array (counts_dev_ptrs[p] = rank p's array,
 * peer-mapped), stream-ordered after the projection launches that counted into n_ok_dev: with a barrier of the
 * symmetric-memory group afterwards no collective library call is needed to exchange the counts.                */
int ccp_publish_count(ccp_handle* h, const int64_t* n_ok_dev, int32_t world, int32_t rank,
                      const uint64_t* counts_dev_ptrs, void* stream);

/* ≙ isSatisfied (ConstraintFunction.h:114-120): finite and f0 <= tol1 and f1 <= tol2.       */
int ccp_is_satisfied_batch(ccp_handle* h, const double* x_dev, int64_t count, int32_t layout,
                           uint8_t* out_dev, void* stream);
/* ≙ jointValid (ConstraintFunction.h:43-55).                                                */
int ccp_joint_valid_batch(ccp_handle* h, const double* x_dev, int64_t count, int32_t layout,
                          uint8_t* out_dev, void* stream);

/* ---- batched kinematics API (RobotModel virtuals, kinematics/panda_rbdl.h:13-23) -------- */

/* ≙ PandaModel::getTransform (panda_rbdl.cpp:35-42) of arm `arm` in its BASE frame for `count`
 * 7-vectors q (AOS double[count][7] or SOA double[7][count]).
 * T_dev: AOS double[count][12] row-major 3x4 [R|p]; SOA double[12][count].                  */
int ccp_fk_batch(ccp_handle* h, int32_t arm, const double* q_dev, int64_t count, int32_t layout,
                 double* T_dev, void* stream);
/* ≙ PandaModel::getJacobianMatrix (panda_rbdl.cpp:9-22): 6x7, rows [linear(3); angular(3)],
 * base frame.  J_dev: AOS double[count][6][7]; SOA double[6][7][count].                     */
int ccp_arm_jacobian_batch(ccp_handle* h, int32_t arm, const double* q_dev, int64_t count,
                           int32_t layout, double* J_dev, void* stream);

/* ---- batched projected-state sampler (jy_ProjectedStateSampler, jy_ProjectedStateSpace.cpp:10-29)
 * Seeds come from a counter-based generator: seed i depends only on (rng_seed, first_index+i),
 * so any rank/GPU can generate any slice.  mode 0 = uniform in [lb,ub] (sampleUniform),
 * mode 1 = uniform in the box near +- distance, clipped to [lb,ub] (sampleUniformNear),
 * mode 2 = gaussian(mean = near, stddev = distance), clipped to [lb,ub] (sampleGaussian).
 * Each seed is projected; if wrap_bounds != 0 the result is wrapped to [-pi,pi)
 * (KinematicChainSpace::enforceBounds, KinematicChain.h:118-130).
 * Outputs (any may be NULL except n_ok_dev when compact_dev != NULL):
 *   x_out_dev/ok_dev/iters_dev : per-seed results as in ccp_project_batch (AOS/SOA by layout)
 *   compact_dev : double[<=count][n] AOS, the states with ok==1 densely packed (order unspecified)
 *   n_ok_dev    : int64 counter (device), ADDED to (caller zeroes it)                        */
typedef struct ccp_sampler_args {
  uint64_t rng_seed;
  int64_t first_index;
  int32_t mode;
  int32_t wrap_bounds;
  double distance;
  const double* near_host; /* 7*K doubles, host; modes 1,2 */
} ccp_sampler_args;

int ccp_generate_seeds(ccp_handle* h, const ccp_sampler_args* a, int64_t count, int32_t layout,
                       double* seeds_dev, void* stream);
int ccp_sample_project_batch(ccp_handle* h, const ccp_sampler_args* a, int64_t count, int32_t layout,
                             double* x_out_dev, uint8_t* ok_dev, int32_t* iters_dev,
                             double* compact_dev, int64_t* n_ok_dev, void* stream);
/* Pipelined form (see ccp_project_batch_pipelined): the pool refill never waits on a batch's stragglers.     */
int ccp_sample_project_batch_pipelined(ccp_handle* h, const ccp_sampler_args* a, int64_t count, int32_t layout,
                                       double* x_out_dev, uint8_t* ok_dev, int32_t* iters_dev,
                                       double* compact_dev, int64_t* n_ok_dev, void* stream);
/* ≙ KinematicChainSpace::enforceBounds (KinematicChain.h:118-130), in place.                */
int ccp_enforce_bounds_batch(ccp_handle* h, double* x_dev, int64_t count, int32_t layout, void* stream);

/* ---- multi-GPU for a C++ host (the reference planner is one C++ process, src/main.cpp:27-63) ----------------------
 * (a) ONE PROCESS, one handle per GPU: a peer group gives every device a pool double[world][capacity][n] + int64
 *     counts[world] that its peers can write (cudaDeviceEnablePeerAccess between all pairs).
 *     ccp_peer_group_sample_project shards `total` counter-stream seeds over the handles (rank r projects a contiguous
 *     slice starting at a->first_index + its offset), every projection kernel stores its converged states into EVERY
 *     device's pool from its epilogue and publishes its count the same way; the call returns when all devices are done,
 *     and then pool[r][q][0 .. counts[q]) on every device r holds rank q's converged states — the all-gather of SURVEY
 *     §8e without a collective library.  counts_host (int64[world], may be NULL) receives the counts; more converged
 *     states than `capacity` on a rank is an error (CCP_ERR_INVALID).  ccp_peer_group_gather_host copies device `rank`'s
 *     gathered pool to the host, padding dropped, rank-major, and reports the rows.  The handles stay usable on their
 *     own between calls; they must outlive the group.                                                              */
typedef struct ccp_peer_group ccp_peer_group;
int ccp_peer_group_create(ccp_handle* const* handles, int32_t world, int64_t capacity, ccp_peer_group** out);
void ccp_peer_group_destroy(ccp_peer_group* g);
int32_t ccp_peer_group_world(const ccp_peer_group* g);
const char* ccp_peer_group_last_error(const ccp_peer_group* g); /* g may be NULL: error of the last failed create */
int ccp_peer_group_sample_project(ccp_peer_group* g, const ccp_sampler_args* a, int64_t total, int64_t* counts_host);
int ccp_peer_group_pool(const ccp_peer_group* g, int32_t rank, const double** pool_dev, const int64_t** counts_dev,
                        int64_t* capacity);
int ccp_peer_group_gather_host(ccp_peer_group* g, int32_t rank, double* states_host, int64_t max_rows,
                               int64_t* counts_host, int64_t* rows_out);
/* (b) ONE PROCESS PER GPU with an NCCL communicator the host owns (nccl_comm = its ncclComm_t): all-gather of the
 *     converged counts (counts_dev int64[world]) and of the compacted states padded to `capacity` rows (compact_dev
 *     double[>=capacity][n] from ccp_project_batch / ccp_sample_project_batch, pool_dev double[world][capacity][n]),
 *     asynchronous on `stream`.  NCCL is looked up at run time — first in the process image, so that the communicator
 *     and the calls belong to the same library, then libnccl.so.2 — libccp.so does not link it; CCP_ERR_NCCL when it
 *     cannot be found or a call fails.  Inside one thread driving several devices, bracket the per-device calls with the
 *     host's own ncclGroupStart/ncclGroupEnd as NCCL requires.                                                     */
int ccp_allgather_converged(ccp_handle* h, void* nccl_comm, int32_t world, const double* compact_dev,
                            const int64_t* n_ok_dev, int64_t capacity, double* pool_dev, int64_t* counts_dev, void* stream);

/* ---- batched manifold traversal --------------------------------------------------------- */
/* ≙ jy_ProjectedStateSpace::discreteGeodesic(from, to, interpolate = true, &geodesic)
 * (jy_ProjectedStateSpace.cpp:32-96) for `edges` independent edges: walk from `from` toward `to` in steps of
 * `delta` (KinematicChainSpace::interpolate, KinematicChain.h:145-171), projecting every step; stop when the
 * projection fails, the step deviates by more than lambda*delta, the arc length exceeds lambda*dist, or no progress
 * is made.  The state-validity check of the reference (MoveIt collision, :66) is NOT run: validate the returned states
 * on the host (lazily).  Reference constants: delta 0.25, lambda 2.0 (ConstrainedPlanningCommon.cpp:118-119).
 *   from_dev, to_dev : AOS double[edges][n]
 *   states_dev       : AOS double[edges][max_states][n]; states[e][0] = from[e], then the accepted projections
 *   n_states_dev     : int32[edges]  number of valid states of edge e (>= 1)
 *   reached_dev      : uint8[edges]  discreteGeodesic's return value
 *   iters_dev        : int32[edges]  Newton iterations spent on the edge, or NULL
 * An edge that needs more than max_states states is reported as not reached.                              */
int ccp_geodesic_batch(ccp_handle* h, const double* from_dev, const double* to_dev, int64_t edges, double delta,
                       double lambda, int32_t max_states, double* states_dev, int32_t* n_states_dev,
                       uint8_t* reached_dev, int32_t* iters_dev, void* stream);

/* ---- batched single-arm pose IK (goal sampling) ----------------------------------------------------------
 * ≙ IKTask::solve / random_solve (src/base/constraints/ik_task.cpp:16-49) -> panda_ik::solve /
 * TrackIKAdaptor::randomSolve (src/kinematics/panda_tracik.cpp:62-88,140-158) for a whole batch.  TRAC-IK is third
 * party and not reproduced; the solver here is a damped Newton iteration on the 6-D pose error with joint-limit
 * clamping (KDL ChainIkSolverPos_NR_JL's idea).  An answer is correct iff FK(q) hits the target within the tolerance
 * inside the limits — that is the parity criterion (tests/test_ik_gpu.py, checked with the reference-faithful FK).
 * Poses are those PandaModel::getTransform returns: the EE frame in the arm's BASE frame, row-major 3x4 [R|p]; the
 * reference's t_b7 = t_wb^-1 * T_obj * t_o7 (ik_task.cpp:24) converts to it with the constant flange/hand offsets. */
typedef struct ccp_ik_options {
  int32_t max_iter;     /* 200 Newton steps per solve */
  int32_t reserved;
  double eps_pos;       /* 1e-5 m   on every position component (TRAC-IK default eps) */
  double eps_rot;       /* 1e-5 rad on every rotation-error component */
  double damping;       /* 1e-4: lambda^2 added to the diagonal of J J^T */
  double joint_margin;  /* 1e-3: a solution must stay this far inside the limits (panda_tracik.cpp:99-108) */
} ccp_ik_options;
void ccp_ik_default_options(ccp_ik_options* o);
/* One solve per (target, seed) pair.  T_target_dev double[count][12], q_seed_dev / q_out_dev double[count][7]
 * (q_out = last iterate, also on failure), ok_dev uint8[count] (converged inside the limits), iters_dev int32[count]
 * or NULL, err_dev double[count][2] = final (|p error|_inf, |rotation error|_inf) or NULL.  opt NULL = defaults.   */
int ccp_ik_batch(ccp_handle* h, int32_t arm, const double* T_target_dev, const double* q_seed_dev, int64_t count,
                 const ccp_ik_options* opt, double* q_out_dev, uint8_t* ok_dev, int32_t* iters_dev, double* err_dev,
                 void* stream);
/* ≙ the goal sampler's per-arm loop (jy_ConstrainedValidStateSampler.h:63-189): for each of n_targets targets run
 * `restarts` (1..32) solves side by side — restart 0 from q_ref (if given), the others from N(mid-range, sigma) draws
 * clipped to the limits (TrackIKAdaptor::getRandomConfig, sigma = 0.3 there) — and keep the seeded solution if it
 * succeeded, else the successful one nearest to q_ref (without q_ref: the lowest-numbered successful restart).
 * q_best_dev double[n_targets][7] (untouched where ok == 0), ok_dev uint8[n_targets], n_success_dev int32[n_targets]
 * (how many restarts converged) or NULL.  Restarts that can no longer win are abandoned — with q_ref once the seeded one
 * has succeeded (the reference only draws the others after a failed seeded solve), without q_ref once a lower-numbered
 * one has succeeded (the draws are i.i.d.: any success is distributed alike) — so n_success counts the successful
 * restarts that got to finish; ok and q_best do not depend on it.                                                   */
int ccp_ik_sample_batch(ccp_handle* h, int32_t arm, const double* T_target_dev, int64_t n_targets, int32_t restarts,
                        uint64_t rng_seed, double sigma, const double* q_ref_dev, const ccp_ik_options* opt,
                        double* q_best_dev, uint8_t* ok_dev, int32_t* n_success_dev, void* stream);

/* ≙ jy_ValidStateSampler::sampleCalibGoal / sampleRandomGoal (base/jy_ConstrainedValidStateSampler.h:63-189) for a BATCH of
 * object poses: for every arm a of the handle the IK target is t_b7 = t_wb_a^-1 * T_obj * t_o7_a (IKTask::solve /
 * random_solve, src/base/constraints/ik_task.cpp:16-49; t_wb_a is the handle's, t_o7_a the grasp frame the reference
 * derives from the start configuration, ConstrainedPlanningCommon.cpp:105-111) and ccp_ik_sample_batch's restarts run on it
 * — restart 0 from the arm's seven joints of q_ref (sampleCalibGoal's q0_ / start_state segment) when q_ref_dev is given,
 * the others from N(mid-range, sigma); the seeded solution wins, else the successful restart nearest to the reference.
 * A pose is ok when EVERY arm found a solution: q_out then holds the 7K-vector of a closed-chain goal configuration (its
 * closure error is the IK tolerance, so project() accepts it as it is); arms that failed leave their columns untouched.
 *   T_obj_dev double[n][12] row-major 3x4 object poses in the world; t_o7_host double[K][12] (host); q_ref_dev double[n][7K]
 *   or NULL (sampleRandomGoal); q_out_dev double[n][7K]; ok_dev uint8[n].  The reference's IKValid collision check
 *   (jy_ConstrainedValidStateSampler.h:190-197) is a host concern: validate the ok rows.                               */
int ccp_goal_sample_batch(ccp_handle* h, const double* T_obj_dev, int64_t n, const double* t_o7_host,
                          const double* q_ref_dev, int32_t restarts, uint64_t rng_seed, double sigma,
                          const ccp_ik_options* opt, double* q_out_dev, uint8_t* ok_dev, void* stream);

/* ---- host-buffer entry points (what a planner that owns host states calls) -------------- */
/* Same as ccp_project_batch but all pointers are HOST memory, AOS double[count][n]
 * (the gathered OMPL states).  Copies in, projects, copies out; synchronous.                */
int ccp_project_batch_host(ccp_handle* h, const double* seeds_host, int64_t count,
                           double* x_out_host, uint8_t* ok_host, uint8_t* converged_host,
                           int32_t* iters_host, double* resid_host);
/* Streaming form of ccp_project_batch_host for a caller with batch after batch of host states: submit enqueues the
 * whole batch (copies and launches) and returns a ticket, wait blocks until that batch's results are in its buffers
 * (which must stay valid until then).  Up to two batches are in flight (a third submit first waits for the oldest).
 * When batch k + 1 is submitted before batch k is waited for, k + 1's launches finish k's stragglers — no launch tail —
 * and the copies of one batch run behind the kernels of the other.  Results are bit-identical to
 * ccp_project_batch_host.  With PAGE-LOCKED output buffers each chunk is copied out the moment its last sample has
 * finished (a stream wait on the chunk's completion count); with pageable outputs a fixed number of launches later.
 * Whatever is pending on the device after submit returns completes without a further call, so device-wide
 * synchronisation between submit and wait is safe.  Do not issue other projection calls on the handle while a ticket
 * is pending (the blocking ccp_project_batch_host may be called: it first completes the pending tickets).          */
int ccp_project_batch_host_submit(ccp_handle* h, const double* seeds_host, int64_t count, double* x_out_host,
                                  uint8_t* ok_host, uint8_t* converged_host, int32_t* iters_host,
                                  double* resid_host, int64_t* ticket_out);
int ccp_project_batch_host_wait(ccp_handle* h, int64_t ticket);
/* General form of the streaming host path: one batch = either host states (seeds_host) or counter-based sampler
 * arguments (sampler: the seeds are generated on the device, nothing is copied in — jy_ProjectedStateSampler::
 * sampleUniform/Near/Gaussian, jy_ProjectedStateSpace.cpp:10-29, for a whole pool refill), and any subset of outputs:
 *   per-seed   x_out_host [count][n], ok_host, converged_host, iters_host, resid_host [count][m]   (NULL = not wanted:
 *              neither computed-and-copied nor, for x_out, written at all)
 *   compact    compact_host [compact_capacity][n]: the states with ok == 1 of THIS batch densely packed (order
 *              unspecified; whichever launch finishes a straggler, it joins its own batch's rows);
 *              compact_index_host [compact_capacity]: the seed index (0 .. count-1) of each packed row.
 *              The reference's sampler hands the planner one projected state per call and the planner drops the failed
 *              ones; a pool refill only needs the ok rows: 112 B x ~21 % instead of 112 B per seed back over PCIe.
 * ccp_host_batch_wait returns in *n_ok_out the number of ok states of the batch (-1 when no compact output was asked
 * for); min(n_ok, compact_capacity) rows are valid.  The packed rows are copied by the wait itself (sized by the count,
 * while the device already works on the next batch).  sampler->wrap_bounds applies enforceBounds to every result.
 * Ticket and ordering rules are those of ccp_project_batch_host_submit / _wait, with which tickets are shared.        */
typedef struct ccp_host_batch {
  const double* seeds_host;
  const ccp_sampler_args* sampler;
  int64_t count;
  double* x_out_host;
  uint8_t* ok_host;
  uint8_t* converged_host;
  int32_t* iters_host;
  double* resid_host;
  double* compact_host;
  int32_t* compact_index_host;
  int64_t compact_capacity;
} ccp_host_batch;
int ccp_host_batch_submit(ccp_handle* h, const ccp_host_batch* batch, int64_t* ticket_out);
int ccp_host_batch_wait(ccp_handle* h, int64_t ticket, int64_t* n_ok_out);
int ccp_function_batch_host(ccp_handle* h, const double* x_host, int64_t count, double* f_host);
/* Host-buffer forms (AOS; synchronous) of ccp_sample_project_batch, ccp_geodesic_batch and ccp_ik_sample_batch, for a
 * C++ planner that never touches CUDA (include/closed_chain_motion_planner_b200/ProjectedStateSpace.hpp).  Any output
 * the device form allows to be NULL may be NULL; compact_host receives *n_ok_host rows.                            */
int ccp_sample_project_batch_host(ccp_handle* h, const ccp_sampler_args* a, int64_t count, double* x_out_host,
                                  uint8_t* ok_host, int32_t* iters_host, double* compact_host, int64_t* n_ok_host);
int ccp_geodesic_batch_host(ccp_handle* h, const double* from_host, const double* to_host, int64_t edges, double delta,
                            double lambda, int32_t max_states, double* states_host, int32_t* n_states_host,
                            uint8_t* reached_host, int32_t* iters_host);
int ccp_ik_sample_batch_host(ccp_handle* h, int32_t arm, const double* T_target_host, int64_t n_targets, int32_t restarts,
                             uint64_t rng_seed, double sigma, const double* q_ref_host, const ccp_ik_options* opt,
                             double* q_best_host, uint8_t* ok_host, int32_t* n_success_host);
int ccp_goal_sample_batch_host(ccp_handle* h, const double* T_obj_host, int64_t n, const double* t_o7_host,
                               const double* q_ref_host, int32_t restarts, uint64_t rng_seed, double sigma,
                               const ccp_ik_options* opt, double* q_out_host, uint8_t* ok_host);
int ccp_jacobian_batch_host(ccp_handle* h, const double* x_host, int64_t count, double* J_host);
/* ≙ PandaModel::getTransform / getJacobianMatrix for host 7-vectors (AOS): T_host double[count][12],
 * J_host double[count][42]; either output may be NULL.                                          */
int ccp_arm_fk_batch_host(ccp_handle* h, int32_t arm, const double* q_host, int64_t count,
                          double* T_host, double* J_host);

/* Page-locked host memory for the host-buffer entry points, for a caller that does not link CUDA itself: with
 * page-locked buffers the copies of the chunked / streaming host path overlap the kernels (1 M states per call: 4.8 ms
 * against 18.7 ms with pageable buffers).  Allocation is slow (it pins pages): allocate once, reuse across calls.  */
int ccp_host_alloc(void** out, size_t bytes);
void ccp_host_free(void* p);
/* Page-lock memory the caller already owns (a long-lived std::vector's storage, a numpy array) and release it again
 * before that memory is freed.  Same effect on the host path as ccp_host_alloc; also slow, do it once per buffer.  */
int ccp_host_register(void* p, size_t bytes);
int ccp_host_unregister(void* p);

/* ---- measurement helpers ---------------------------------------------------------------- */
/* Register-only DFMA chains on every SM: returns achieved FP64 FLOP/s (FMA = 2) and the
 * kernel time in ms.  Used as the MEASURED FP64 peak of the box (MEASURED_PEAKS.json has none). */
int ccp_fp64_peak_probe(ccp_handle* h, int32_t repeats, double* flops_per_s, double* ms);
/* Number of engine kernels launched through this handle since creation. */
int64_t ccp_launch_count(const ccp_handle* h);
/* ccp_project_batch bracketed by CUDA events on `stream`; synchronises and returns the kernel time. */
int ccp_project_batch_timed(ccp_handle* h, const double* seeds_dev, int64_t count, int32_t layout,
                            double* x_out_dev, uint8_t* ok_dev, uint8_t* converged_dev,
                            int32_t* iters_dev, double* resid_dev, double* compact_dev,
                            int64_t* n_ok_dev, void* stream, float* kernel_ms);

/* Algorithmic FLOPs per Newton iteration / per final evaluation for this handle's K
 * (frozen in csrc/ccp_flops.h).                                                 */
int ccp_algorithmic_flops(const ccp_handle* h, double* per_iteration, double* per_tail);

/* Same for one damped-Newton iteration / the final evaluation of the batched pose IK (csrc/ccp_flops.h). */
int ccp_algorithmic_flops_ik(double* per_iteration, double* per_tail);

const char* ccp_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CCP_H_ */
