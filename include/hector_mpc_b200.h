/*
 * hector_mpc_b200.h — C-ABI of libhector_mpc_b200.so
 *
 * A B200 (sm_100a) batched force-and-moment MPC solver that is a drop-in for the
 * C boundary of DRCL-USC/Hector_Simulation's convex MPC:
 *
 *   reference boundary : hector_control/ConvexMPC/convexMPC_interface.h:11-43
 *   reference caller   : hector_control/ConvexMPC/ConvexMPCLocomotion.cpp:410,415,428-429
 *
 * Part 1 re-exports the four reference symbols with identical signatures and semantics
 * (one robot, blocking solve, result held by the library).  Part 2 is the additive batched
 * interface (thousands of independent robots per launch).  Plain pointers and sizes only:
 * no C++/torch types cross this boundary.
 *
 * There is NO CPU fallback behind these entry points: if no CUDA device / kernel image is
 * usable, every entry point reports HMPC_ERR_CUDA (batched API) or aborts with a message
 * (reference API, which has no error channel — convexMPC_interface.h:39-43).
 */
#ifndef HECTOR_MPC_B200_H
#define HECTOR_MPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#define HMPC_EXTERNC extern "C"
#else
#define HMPC_EXTERNC
#endif

/* ------------------------------------------------------------------------------------------
 * Part 0 — POD records, byte-compatible with the reference
 * ---------------------------------------------------------------------------------------- */

#define K_MAX_GAIT_SEGMENTS 36 /* convexMPC_interface.h:3 */

/* convexMPC_interface.h:11-17.  `mu` is carried but ignored by the solver, exactly like the
 * reference (SolverMPC.cpp:488 shadows it with a local 2.0). */
struct problem_setup
{
  float dt;
  float mu;
  float f_max;
  int horizon;
};

/* convexMPC_interface.h:19-37 — same member order, types and padding (sizeof == 3016). */
struct update_data_t
{
  float p[3];                          /* CoM position, world                               */
  float v[3];                          /* CoM velocity, world                               */
  float q[4];                          /* orientation quaternion (w,x,y,z)                  */
  float w[3];                          /* angular velocity, world                           */
  float r[6];                          /* foot - CoM, layout [x0,x1,y0,y1,z0,z1]            */
  float joint_angles[10];              /* leg0 q0..q4, leg1 q0..q4 (before the solver's offset) */
  float yaw;
  float weights[12];                   /* state tracking weights (rpy, p, w, v)             */
  float traj[12 * K_MAX_GAIT_SEGMENTS];/* reference trajectory, 12 per horizon step         */
  float Alpha_K[12];                   /* input regularisation [F0 F1 M0 M1]                */
  unsigned char gait[K_MAX_GAIT_SEGMENTS]; /* contact table, [step][leg], 0/1             */
  unsigned char hack_pad[1000];
  int max_iterations;
  double rho, sigma, solver_alpha, terminate;
};

/* ------------------------------------------------------------------------------------------
 * Part 1 — the reference's own entry points (convexMPC_interface.h:39-43)
 *
 *   setup_problem        replaces convexMPC_interface.cpp:42-66  (+ resize_qp_mats, SolverMPC.cpp:196-299)
 *   update_problem_data  replaces convexMPC_interface.cpp:83-103 (+ solve_mpc, SolverMPC.cpp:371-738,
 *                         + qpOASES QProblem::init, SolverMPC.cpp:702-712)
 *   get_solution         replaces convexMPC_interface.cpp:105-110
 *   update_solver_settings replaces convexMPC_interface.cpp:112-118 (stored, otherwise unused)
 *
 * Semantics kept: update_problem_data blocks until the wrench is available; get_solution(i),
 * i in [0, 12*horizon), is step-major [F0(3) F1(3) M0(3) M1(3)] in the world frame and returns
 * 0.0 before the first solve; swing-leg entries are exactly 0.0.  Single caller thread.
 * Difference (documented in INTEGRATION.md): horizon > 19 aborts with a message instead of
 * letting a C++ exception escape (SolverMPC.cpp:140-143).
 * ---------------------------------------------------------------------------------------- */
HMPC_EXTERNC void setup_problem(double dt, int horizon, double mu, double f_max);
HMPC_EXTERNC double get_solution(int index);
HMPC_EXTERNC void update_solver_settings(int max_iter, double rho, double sigma, double solver_alpha,
                                         double terminate, double use_jcqp);
HMPC_EXTERNC void update_problem_data(double* p, double* v, double* q, double* w, double* r,
                                      double* joint_angles, double yaw, double* weights,
                                      double* state_trajectory, double* Alpha_K, int* gait);

/* The boundary above has no error channel (its functions return void / a double).  Two additive calls give a 1 kHz
 * controller one without changing the reference's call sites:
 *   hmpc_reference_last_status  the status word of the last update_problem_data (HMPC_STATUS_* macros below): code,
 *                               working-set changes, active rows — what "failed to solve!" (SolverMPC.cpp:714-715) hides.
 *   hmpc_reference_last_rc      HMPC_OK, HMPC_ERR_NOT_CONVERGED, or the error of the last tick.  A tick that fails at run
 *                               time (a CUDA error after a successful setup_problem) does NOT end the process: the error is
 *                               printed once per episode, get_solution keeps returning the last tick's wrench, and this call
 *                               (with hmpc_last_error()) tells the controller, which can fall back to its stand-still
 *                               behaviour.  HMPC_REFERENCE_ABORT=1 in the environment restores "abort on any failure".
 * Misuse and start-up failures stay fatal with a message: horizon > 19 (the reference throws), update_problem_data before
 * setup_problem, no usable GPU at setup_problem. */
HMPC_EXTERNC int hmpc_reference_last_status(void);
HMPC_EXTERNC int hmpc_reference_last_rc(void);

/* ------------------------------------------------------------------------------------------
 * Part 2 — batched interface (additive; SURVEY.md §8b "batched extension")
 * ---------------------------------------------------------------------------------------- */

typedef struct hmpc_ctx hmpc_ctx;

/* return codes */
#define HMPC_OK 0
#define HMPC_ERR_ARG 1      /* bad argument (null pointer, batch > capacity, horizon out of range) */
#define HMPC_ERR_CUDA 2     /* CUDA runtime error / no device / no sm_100a image; hmpc_last_error() has text */
#define HMPC_ERR_NOT_CONVERGED 3 /* at least one instance did not reach a KKT point; see status[] */

/* per-instance status word written with every result (never silently stale — contrast
 * SolverMPC.cpp:714-715, which prints and carries on with stale data):
 *   bits  0..7  : termination code  (0 = optimal, 1 = iteration cap, 2 = working-set capacity,
 *                                    3 = infeasible/degenerate step, 4 = Hessian not positive definite ENOUGH: a
 *                                    non-positive pivot, or max_i H_ii (H^-1)_ii above the conditioning limit of the
 *                                    fp64 sweep inversion (1.5e5; 2e2 ... 1e4 on the workloads of BASELINE.json, 3e5
 *                                    for a robot lying on its side) — the wrench of such an instance is not trusted)
 *   bits  8..19 : working-set changes performed (comparable to qpOASES nWSR)
 *   bits 20..27 : number of active constraints at the solution
 */
#define HMPC_STATUS_CODE(s) ((s) & 0xff)
#define HMPC_STATUS_ITERS(s) (((s) >> 8) & 0xfff)
#define HMPC_STATUS_NACTIVE(s) (((s) >> 20) & 0xff)

#define HMPC_MAX_HORIZON 16 /* dense fp64 working set of one QP must fit one SM's shared memory */

/* Packed device record (HBM layout, one per robot).  Only the bytes that change per tick:
 *   float state[30]  = p3 v3 q4 w3 r6 joint10 yaw1
 *   float weights[12], float alpha[12]
 *   float traj[12*N]
 *   u8    gait[2*N]   (+ zero padding to a multiple of 16 bytes)
 * hmpc_record_bytes(N) gives the stride.  N=10: 216 + 98*N = 1196 algorithmic bytes per QP
 * in+out, 720-byte input stride. */
HMPC_EXTERNC size_t hmpc_record_bytes(int horizon);
/* pack n reference records into the device layout (host side helper, pure byte shuffling) */
HMPC_EXTERNC int hmpc_pack_records(const struct update_data_t* in, int n, int horizon, void* out);

/* create a context on `device` able to hold `max_batch` robots of `horizon` steps */
HMPC_EXTERNC hmpc_ctx* hmpc_create(int max_batch, int horizon, int device);
HMPC_EXTERNC void hmpc_destroy(hmpc_ctx* ctx);
HMPC_EXTERNC const char* hmpc_last_error(void);

/* dt / f_max of problem_setup (mu is ignored like the reference).  Defaults 0.04 / 500. */
HMPC_EXTERNC int hmpc_set_problem(hmpc_ctx* ctx, const struct problem_setup* setup);

/* Host-buffer path (the reference-facing call: H2D + solve + D2H inside).
 *   in         : B reference records (host)
 *   wrench_out : [B][12*horizon] doubles, same layout as get_solution (host)
 *   status     : [B] status words (host), may be NULL
 * Blocks until results are in host memory. */
HMPC_EXTERNC int hmpc_solve_batch(hmpc_ctx* ctx, const struct update_data_t* in, int B,
                                  double* wrench_out, int* status);

/* Device-resident path: `d_records` = B packed records already in HBM, `d_wrench`
 * [B][12*horizon] float, `d_status` [B] int, all device pointers; enqueued on `stream`
 * (a cudaStream_t passed as void*), returns without synchronising. */
/* d_records must be 16-byte aligned (the kernels stage records with a bulk copy); B <= the context's capacity. */
HMPC_EXTERNC int hmpc_solve_device(hmpc_ctx* ctx, const void* d_records, int B, float* d_wrench,
                                   int* d_status, void* stream);

/* Row f-2: the same solves with the leg-controller epilogue fused in — joint torques of the
 * first-step wrench, tau[leg*5+j] = (J_force_moment^T * f_ff)[j], f_ff = -rBody * [F; M]
 * (replaces LegController.cpp:57-63 + :108-166 and ConvexMPCLocomotion.cpp:419-440 for stance legs; swing legs 0).
 * tau_out [B][10] doubles (host) / d_tau [B][10] floats (device); pass NULL to skip. */
HMPC_EXTERNC int hmpc_solve_batch_ex(hmpc_ctx* ctx, const struct update_data_t* in, int B, double* wrench_out,
                                     double* tau_out, int* status);
HMPC_EXTERNC int hmpc_solve_device_ex(hmpc_ctx* ctx, const void* d_records, int B, float* d_wrench, int* d_status,
                                      float* d_tau, void* stream);

/* In-place mode of hmpc_solve_batch(_ex).  A control loop reuses the same record / result arrays every tick; register
 * them once (cudaHostRegister underneath) and hmpc_solve_batch lets the GPU read the update_data_t records where they
 * lie and write the double wrenches (and status) where the caller wants them: no packing, no staging copies, no
 * float->double pass on the host.  It is used whenever `in`, `wrench_out` and `status` (if given) of a call all lie
 * inside pinned ranges; results are identical to the staged path.  The caller keeps the memory allocated until
 * hmpc_unpin_host_buffer / hmpc_destroy.  (The reference-style entry points pin their own globals.) */
/* Registration is page-granular: the pages that hold [ptr, ptr + bytes) are pinned whole.  Give the arrays their own pages
 * (posix_memalign to the page size, size rounded up): an unrelated host buffer that shares such a page only partly cannot be
 * the source / destination of a later cudaMemcpy (cudaErrorInvalidValue). */
HMPC_EXTERNC int hmpc_pin_host_buffer(hmpc_ctx* ctx, void* ptr, size_t bytes);
HMPC_EXTERNC int hmpc_unpin_host_buffer(hmpc_ctx* ctx, void* ptr);

/* Row f-1: the caller's data preparation on the device.  `hmpc_state_t` is what
 * ConvexMPCLocomotion::updateMPCIfNeeded reads before it builds the MPC inputs (ConvexMPCLocomotion.cpp:279-346),
 * in double precision as the reference holds it; hmpc_prepare_device turns B of them into packed records (joint
 * offsets + fmod, foot positions r, weights, the 12 x horizon reference trajectory, double -> float narrowing:
 * ConvexMPCLocomotion.cpp:283-406 + convexMPC_interface.cpp:87-99) with one GPU thread per robot, so a tick moves
 * 352 bytes per robot to the device instead of a 720-byte record.  hmpc_solve_batch_states = H2D of the states +
 * hmpc_prepare_device + the solve of hmpc_solve_batch_ex. */
struct hmpc_state_t
{
  double position[3];              /* seResult.position */
  double vWorld[3];
  double orientation[4];           /* (w,x,y,z) */
  double omegaWorld[3];
  double rpy[3];                   /* seResult.rpy */
  double leg_q[10];                /* _legController->data[leg].q (LegController's own offset already applied) */
  double leg_p[6];                 /* _legController->data[leg].p, [leg][xyz] */
  double state_des[5];             /* stateDes[3], [4] (roll, pitch), [6], [7] (body-frame vx, vy), [11] (yaw rate) */
  double world_position_desired[2];
  unsigned char gait[K_MAX_GAIT_SEGMENTS]; /* mpcTable, [step][leg] */
  unsigned char pad[4];
};
/* dtMPC is the caller's double `dt * iterationsBetweenMPC` (ConvexMPCLocomotion.cpp:20) used for the trajectory;
 * the QP itself keeps the float dt of hmpc_set_problem, exactly as the reference splits the two. */
HMPC_EXTERNC int hmpc_prepare_device(hmpc_ctx* ctx, const struct hmpc_state_t* d_states, int B, double dtMPC,
                                     void* d_records, void* stream);
HMPC_EXTERNC int hmpc_solve_batch_states(hmpc_ctx* ctx, const struct hmpc_state_t* in, int B, double dtMPC,
                                         double* wrench_out, double* tau_out, int* status);

/* Row f-3: the closed loop on the device — BASELINE config 5 (consecutive ticks of the same robots).
 * One tick = hmpc_prepare_device -> the solve -> hmpc_advance, all enqueued on one stream, no host in the loop:
 *   - the next tick's contact table is Gait::mpc_gait of the advanced iteration counter (GaitGenerator.cpp:85-103);
 *   - the body is integrated one forward-Euler step of the single rigid body the MPC predicts with
 *     (SolverMPC.cpp:145-146, 312-331: x+ = x + dt f(x,u)), feet pinned in the world, first-step wrench fed back;
 *   - a leg that touches down is placed by the heuristic of ConvexMPCLocomotion.cpp:119-160 (hip projection +
 *     0.5 v T_stance + 0.02 (v - v_des), clamped, z = 0) — the swing trajectory itself is out of scope (row f-4);
 *   - world_position_desired integrates the commanded velocity with the clamp write-back of :338-346.
 * hmpc_rollout_t is the per-robot loop state that lives next to hmpc_state_t. */
struct hmpc_rollout_t
{
  double feet_world[6];   /* [leg][xyz]: where each foot is pinned (stance) or was last pinned (swing) */
  int gait_offset[2];     /* Gait(nIterations = horizon, offsets, durations)  (ConvexMPCLocomotion.cpp:16-17) */
  int gait_duration[2];
  int iteration;          /* MPC tick counter; Gait::_iteration = iteration % horizon */
  int failures;           /* ticks whose solver status code was not 0 (accumulated) */
  int iters_total;        /* working-set changes, accumulated */
  int ticks;              /* ticks advanced so far */
};
/* Warm start: every tick after the first proposes the previous tick's optimal working set, moved one step with the
 * horizon, to the active-set stage (the reference cold-starts every tick, SolverMPC.cpp:702-709; BASELINE.json
 * configs[4] names the warm start).  The optimum is the same point either way — the working set only decides how many
 * changes the solver makes to reach it; the status word then counts the changes relative to the proposal.  The sets live in
 * the context; hmpc_reset_warm_start forgets them (a new loop on the same context).  HMPC_WARM_START=0 in the environment
 * at hmpc_create keeps every tick a cold start. */
HMPC_EXTERNC int hmpc_reset_warm_start(hmpc_ctx* ctx, void* stream);
/* ticks >= 1.  d_wrench_log: NULL or float [ticks][B][12] (first-step wrench of every tick); d_record_log: NULL or
 * [ticks][B][hmpc_record_bytes] (the packed records the solver saw, for after-the-fact parity checks). */
HMPC_EXTERNC int hmpc_rollout_device(hmpc_ctx* ctx, struct hmpc_state_t* d_states, struct hmpc_rollout_t* d_loop, int B,
                                     int ticks, double dtMPC, float* d_wrench_log, void* d_record_log, void* stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * Multi-GPU (the reference has nothing here: one robot per process).  Robots are independent, so a batch
 * larger than one device is cut into contiguous slices, one process per GPU, identical kernels, no data-path collective.
 * Every rank gets ITS slice back on the host exactly like hmpc_solve_batch (in place when the arrays are pinned); when a
 * consumer needs the whole batch on every device, `d_all` (device, float [world * B_local][12 * horizon], rank-major)
 * receives ONE ncclAllGather of the float wrenches, enqueued behind the solve on a side stream so that it overlaps the
 * next tick (results are double-buffered); hmpc_shard_wait blocks until the last gather has landed.  NCCL is loaded at run
 * time (libnccl.so.2), only by these calls.
 *   rank 0: hmpc_shard_unique_id(id)  ->  ship the 128 bytes to every rank  ->  all: hmpc_shard_init(ctx, rank, world, id)
 * B_local must be the same on every rank (pad the last slice).  The gather is a collective of the context's own NCCL
 * communicator: call hmpc_shard_wait before another communicator (e.g. torch.distributed's) runs a collective on the same
 * device, as with any two NCCL communicators. */
#define HMPC_SHARD_ID_BYTES 128
HMPC_EXTERNC int hmpc_shard_unique_id(void* id128);
HMPC_EXTERNC int hmpc_shard_init(hmpc_ctx* ctx, int rank, int world, const void* id128);
HMPC_EXTERNC int hmpc_solve_batch_sharded(hmpc_ctx* ctx, const update_data_t* in_local, int B_local, double* wrench_local,
                                          int* status_local, float* d_all);
HMPC_EXTERNC int hmpc_shard_wait(hmpc_ctx* ctx);

/* Row f-4: the swing-leg controller, batched — swingLegController::updateSwingLeg
 * (src/common/SwingLegController.cpp:46-219): foot position, swing sub-phase (Gait::getSwingSubPhase,
 * GaitGenerator.cpp:54-80), swing-time countdown, touch-down placement (:98-128), Bezier swing trajectory
 * (FootSwingTrajectory.cpp:17-36, Interpolation.h:53-74) and the approximate 5-DoF inverse kinematics (:160-193), one GPU
 * thread per robot, fp64.  hmpc_swing_t is the controller's per-robot memory (swingLegController's members);
 * hmpc_swing_cmd_t is what setDesiredJointState (:198-219) hands to the leg controller for the legs in swing
 * (zeros and swing[leg] = 0 for stance legs).  `d_phase[i]` is Gait::_phase of robot i (GaitGenerator.cpp:112); the
 * gait's offsets/durations come from hmpc_rollout_t (nIterations = the context's horizon). */
struct hmpc_swing_t
{
  double p0[6];          /* footSwingTrajectory[leg]._p0, world */
  double swing_time[2];  /* swingTimes[leg] */
  int first_swing[2];    /* firstSwing[leg] (initially 1, 1) */
};
struct hmpc_swing_cmd_t
{
  double pf[6];          /* planned touch-down position of each leg, world (footSwingTrajectory[leg]._pf) */
  double p_des[6];       /* commands[leg].pDes = pFoot_b, body frame */
  double v_des[6];       /* commands[leg].vDes = vFoot_b */
  double q_des[10];      /* commands[leg].qDes from computeIK */
  int swing[2];          /* swingStates[leg] > 0 */
};
/* dt = the controller's own period (0.001, SwingLegController.h:79), dtSwing = dtMPC (ConvexMPCLocomotion.cpp:70).
 * One call = one updateSwingLeg.  Note that the reference's ConvexMPCLocomotion::run calls updateSwingLeg inside its
 * per-foot loop (ConvexMPCLocomotion.cpp:218), i.e. TWICE per control tick, so its swing-time countdown runs at 2*dt per
 * tick; a caller reproducing that controller calls this function twice per tick (tests/test_reference_tick.py does). */
HMPC_EXTERNC int hmpc_swing_device(hmpc_ctx* ctx, const struct hmpc_state_t* d_states, const struct hmpc_rollout_t* d_loop,
                                   const double* d_phase, struct hmpc_swing_t* d_swing, int B, double dt, double dtSwing,
                                   struct hmpc_swing_cmd_t* d_cmd, void* stream);

/* number of kernel launches hmpc_solve_device enqueues per call (classification pre-pass + one per size class) */
HMPC_EXTERNC int hmpc_launches_per_solve(const hmpc_ctx* ctx);
/* launch configuration of size class `cls` (0 or 1): out[0..5] = threads per CTA, dynamic shared memory bytes,
 * working-set capacity, resident-grid cap (CTAs), max blocks of 6 variables, sweep strip width */
HMPC_EXTERNC int hmpc_class_config(const hmpc_ctx* ctx, int cls, int* out);

/* Debug/parity hook: run only the assembly stage for B packed device records and write the
 * full (un-reduced) fp32 QP data per instance: H [n*n] row-major (upper triangle valid,
 * mirrored), g [n], the 16x12 constraint block [192], lb/ub [16N].  n = 12*horizon. */
HMPC_EXTERNC int hmpc_assemble_device(hmpc_ctx* ctx, const void* d_records, int B, float* d_H,
                                      float* d_g, float* d_Fblk, float* d_lb, float* d_ub,
                                      void* stream);

#endif /* HECTOR_MPC_B200_H */
