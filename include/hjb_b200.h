/*
 * hjb_b200.h — C ABI of libhjb_b200.so: the B200 (sm_100a) implementation of the data-parallel hot path of
 * HaoxiangYou/Q_Learning_with_HJB.
 *
 * The reference has no FFI layer: its boundary is the Python class interface (Dynamics / Controller /
 * VHJBController).  Each entry point below cites the reference method(s) whose per-environment /
 * per-sample Python loop it replaces (paths under the reference root).  The Python classes in
 * q_learning_with_hjb_b200/{dynamics,controller} bind these symbols with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; every array pointer is a DEVICE pointer (fp32 unless stated) owned by
 *     the caller; the library allocates nothing persistent;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, no hidden synchronisation;
 *   - parameter structs are HOST pointers, read during the call (they are copied into the kernel's
 *     parameter space), so they may be freed right after the call returns;
 *   - return value: 0 on success, a negative hjb_status on a usage error, a positive cudaError_t on a CUDA
 *     failure; nothing throws across the ABI.  hjb_status_string() renders either;
 *   - trajectories are TIME-MAJOR: xs[t][env][i].  One thread integrates one environment, so a warp's 32
 *     states of one time step are contiguous and every store is coalesced; the reference's per-environment
 *     view xs_env[t, i] is the strided slice xs[:, env, :] (the Python layer returns it as a zero-copy
 *     permute).
 */
#ifndef HJB_B200_H
#define HJB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HJB_ABI_VERSION 2
#define HJB_MAX_N 10 /* largest state dimension (NearHoverQuadcopter) */
#define HJB_MAX_M 3  /* largest control dimension */

typedef enum hjb_status {
  HJB_OK = 0,
  HJB_ERR_BAD_ARG = -1,      /* null pointer / negative size / inconsistent dims */
  HJB_ERR_UNSUPPORTED = -2,  /* (system, controller, integrator, n, m) combination has no kernel */
  HJB_ERR_NO_DEVICE = -3     /* no sm_100 device visible */
} hjb_status;

/* ---- systems: x' = f(x) + g(x) u ----------------------------------------------------------------- */
typedef enum hjb_system_kind {
  HJB_SYS_LINEAR = 0,   /* dynamics/linear.py:7-22          f = A x, g = B; n in {2,4}, m in {1,2}        */
  HJB_SYS_CARTPOLE = 1, /* dynamics/cartpole.py:19-64       x = [p, th, dp, dth], wraps th               */
  HJB_SYS_ACROBOT = 2,  /* dynamics/acrobot.py:39-81        x = [q1, q2, dq1, dq2], wraps q1, q2          */
  HJB_SYS_QUAD2D = 3,   /* dynamics/quadrotors.py:17-70     x = [x, y, th, dx, dy, dth], wraps th        */
  HJB_SYS_QUAD10D = 4   /* dynamics/quadrotors.py:118-170   10-D near-hover quadcopter, wraps x[3], x[4]  */
} hjb_system_kind;

/* Physical parameters, in the order the reference's config dataclasses / module dict name them:
 *   CARTPOLE par = {mc, mp, l, g}                    (configs/dynamics/dynamics_config.py:35-43)
 *   ACROBOT  par = {l1, l2, m1, m2, I1, I2, g}       (dynamics/acrobot.py:8-16)
 *   QUAD2D   par = {g, m, r, I}                      (configs/dynamics/dynamics_config.py:45-51)
 *   QUAD10D  par = {g, m, kT, n0}                    (configs/dynamics/dynamics_config.py:53-59)
 *   LINEAR   A (n x n, row-major), B (n x m, row-major); par unused                                  */
typedef struct hjb_system {
  int32_t kind; /* hjb_system_kind */
  int32_t n, m;
  float dt;
  float umin[HJB_MAX_M], umax[HJB_MAX_M]; /* Dynamics.simulate clips u to these (dynamics_basic.py:118) */
  float par[8];
  float A[16], B[8];
} hjb_system;

/* ---- controllers ---------------------------------------------------------------------------------- */
typedef enum hjb_control_kind {
  /* u = -K wrap(x - xf) + uf, clipped to [umin, umax] iff `clip`:
   *   controller/lqr.py:29-30 (xf = 0, uf = 0, clip), examples/cartpole_balancing.ipynb cell 4:24-25 (no clip),
   *   controller/quadrotors_model_based_controller.py:36-38 and :73-75 (clip)                         */
  HJB_CTL_FEEDBACK = 0,
  /* controller/cartpole_energy_shaping.py:65-110; K = LQR gain about xf = [0, pi, 0, 0],
   * aux = {Ke0, Ke1, Ke2, eps_energy, eps_state}                                                       */
  HJB_CTL_CARTPOLE_ES = 1,
  /* controller/acrobot_energy_shaping.py:74-121; K, P = LQR gain / cost-to-go about xf = [pi, 0, 0, 0],
   * aux = {Ks0, Ks1, Ks2, eps}                                                                          */
  HJB_CTL_ACROBOT_ES = 2,
  /* Tracking of a TIME-VARYING reference: u_t = clip(u_ref[t] - K wrap(x - x_ref[t]), umin, umax) — the feedback law of
   * controller/quadrotors_model_based_controller.py:36-38 about the state / feed-forward input that
   * Quadrotors2DWaypointsPlanner.update(t) returns (:77-233; minimum snap + differential flatness).  The reference is a
   * device table ref[ref_steps][n + m] (row t: x_ref, then u_ref, at time t * dt; rows past the end repeat the last:
   * hover at the final way-point), the same for every environment — planned once per horizon, not once per step.      */
  HJB_CTL_TRACK = 3,
  /* The double integrator's time-optimal bang-bang law, examples/double_integrator_optimal_time.ipynb cell 18
   * (get_analytical_control): x = [pos, vel]; u = 0 inside x^T x <= aux[0]; u = +aux[1] when (vel < 0 and
   * pos <= vel^2 / 2) or (vel >= 0 and pos < -vel^2 / 2); else u = -aux[1].  n = 2, m = 1.                            */
  HJB_CTL_SWITCH_CURVE = 4,
  /* Bang-bang policy read from a value function on a regular grid (the notebook's comparison with a level-set solver,
   * cell 18: get_level_set_control): u = -aux[4] sign(T[iv][ip]) with T = dV/dvel as a device table ref[ref_steps = nv]
   * [ref_offset = np] and (iv, ip) the NEAREST grid node of (vel, pos) — node = ceil(t - 1/2) of the fractional index
   * t = (x - min) / step, clamped to the grid: scipy's RegularGridInterpolator(method="nearest", bounds_error=False,
   * fill_value=None).  aux = {pos_min, 1 / pos_step, vel_min, 1 / vel_step, amplitude}.  n = 2, m = 1.               */
  HJB_CTL_GRID_SIGN = 5
} hjb_control_kind;

typedef struct hjb_control {
  int32_t kind; /* hjb_control_kind */
  int32_t clip; /* FEEDBACK only */
  float K[HJB_MAX_M * HJB_MAX_N]; /* m x n, row-major */
  float P[16];                    /* ACROBOT_ES: 4 x 4 row-major */
  float xf[HJB_MAX_N], uf[HJB_MAX_M];
  float aux[8];
  const float* ref;   /* TRACK: device pointer, [ref_steps][n + m] row-major; GRID_SIGN: the table; else null */
  int32_t ref_steps;  /* TRACK: rows of ref; GRID_SIGN: rows (vel nodes)                                  */
  int32_t ref_offset; /* TRACK: step 0 of the call is row ref_offset (per-step calls at a given time);
                         GRID_SIGN: columns (pos nodes)                                                    */
} hjb_control;

/* ---- running cost l(x,u) = dx^T Q dx + (u-uf)^T R (u-uf), dx = wrap(x - xf) -----------------------
 * controller/vhjb.py:162-165 (same form in every notebook).  `u` is the controller output as returned,
 * i.e. BEFORE Dynamics.simulate clips it (examples/cartpole_balancing.ipynb cell 15:33-35).            */
typedef struct hjb_cost {
  float Q[HJB_MAX_N * HJB_MAX_N]; /* n x n row-major */
  float R[HJB_MAX_M * HJB_MAX_M]; /* m x m row-major */
  float xf[HJB_MAX_N], uf[HJB_MAX_M];
} hjb_cost;

typedef enum hjb_integrator {
  HJB_INT_EULER = 0,   /* the reference's integrator (dynamics_basic.py:120)                              */
  HJB_INT_RK4 = 1,     /* classical RK4 of the same xdot, u held over the step, one wrap after the step  */
  HJB_INT_DISCRETE = 2 /* LINEAR only: x <- A x + B u with A, B already discretised (exact ZOH,
                          examples/double_integrator_optimal_time.ipynb cell 4)                           */
} hjb_integrator;

typedef struct hjb_rollout_opts {
  int32_t integrator;    /* hjb_integrator */
  int32_t record_stride; /* 0: no trajectory output; s > 0: record x after every s-th step (and x0)      */
  int32_t fast_trig;     /* 1: in-line sin/cos (quadrant reduction + minimax polynomials, 7e-8 abs) and
                            MUFU.RCP, round-to-nearest wrap; 0: libdevice sincosf / tanf, IEEE division    */
  int32_t box_enabled;   /* freeze an environment once wrap(x - box_xf) leaves [box_lo, box_hi]
                            (controller/vhjb.py:176-181; examples/drone_hovering.ipynb cell 15:27)        */
  float box_xf[HJB_MAX_N], box_lo[HJB_MAX_N], box_hi[HJB_MAX_N];
} hjb_rollout_opts;

/*
 * Batched closed-loop rollout — replaces the per-environment loop
 *     for t in range(T): u = controller.get_control_efforts(x); x = dynamics.simulate(x, u)
 * (scripts/test_vhjb_policy.py:146-151, controller/cartpole_energy_shaping.py:123-125,
 *  controller/acrobot_energy_shaping.py:133-135, controller/quadrotors_model_based_controller.py:310-312,
 *  examples notebooks' test_learned_policy) for N environments in ONE launch.
 *
 *   x0       [N, n]                 initial states
 *   xs       [T/s + 1, N, n] | null recorded states, time-major (s = record_stride); xs[0] = x0
 *   us       [T/s, N, m]     | null controller outputs at steps 0, s, 2s, ...
 *   x_final  [N, n]          | null state after the last step
 *   cost     [N]             | null sum_t l(x_t, u_t) * dt  (requires `cost_spec`)
 *   steps    [N] int32       | null number of steps actually integrated (< T only with box_enabled)
 */
int hjb_rollout(const hjb_system* sys, const hjb_control* ctl, const hjb_cost* cost_spec,
                const hjb_rollout_opts* opts, const float* x0, int64_t N, int32_t T, float* xs, float* us,
                float* x_final, float* cost, int32_t* steps, void* stream);

/*
 * Which kernel instantiation hjb_rollout dispatches to for these arguments (no device work; used by the parity tests
 * and bench.py to state WHICH instantiation a number belongs to).  `recorded` = xs or us requested.
 *   out[0] integrator  out[1] recorded (0/1)  out[2] cost mode (0 none, 1 diagonal, 2 dense, 3 unit: Q = I, R = I)
 *   out[3] box (0/1)   out[4] fast_trig (0/1) out[5] controller clips (0/1)
 */
int hjb_rollout_variant(const hjb_system* sys, const hjb_control* ctl, const hjb_cost* cost_spec,
                        const hjb_rollout_opts* opts, int32_t recorded, int32_t out[6]);

/*
 * Batched single-step pieces of the same path, for per-step use and per-step parity checks:
 *   f [B, n], g [B, n, m]  Dynamics.get_control_affine_matrix (dynamics_basic.py:64-94 and overrides)
 *   xdot [B, n]            Dynamics.dynamics_step             (dynamics_basic.py:96-105), needs u
 *   x_next [B, n]          Dynamics.simulate                  (dynamics_basic.py:107-122), needs u
 * Any output may be null; u [B, m] may be null if only f / g are requested.
 */
int hjb_dynamics(const hjb_system* sys, int32_t integrator, int32_t fast_trig, const float* x, const float* u,
                 int64_t B, float* f, float* g, float* xdot, float* x_next, void* stream);

/* Batched Controller.get_control_efforts (controller/controller_basic.py:4-5 and subclasses): x [B, n] -> u [B, m]. */
int hjb_control_efforts(const hjb_system* sys, const hjb_control* ctl, int32_t fast_trig, const float* x, int64_t B,
                        float* u, void* stream);

/* Batched Dynamics.states_wrap (cartpole.py:52-64, acrobot.py:72-81, quadrotors.py:48-70,151-170), in place. */
int hjb_states_wrap(const hjb_system* sys, float* x, int64_t B, void* stream);

/*
 * Time to the goal ball of recorded trajectories — the bookkeeping of examples/double_integrator_optimal_time.ipynb
 * cell 20 (`if x_{k+1}^T x_{k+1} <= metric: optimal_t = min(k dt, optimal_t)`, optimal_t starting at T) for every
 * environment of a rollout: xs = hjb_rollout's time-major record with record_stride = 1 ([rows][N][n], row 0 = x0, so
 * rows = steps + 1); t_hit [N] (device) = (first row r >= 1 inside the ball - 1) * dt, or t_max if there is none.
 */
int hjb_time_to_goal(const float* xs, int64_t N, int32_t n, int32_t rows, float metric, float dt, float t_max, float* t_hit,
                     void* stream);

/*
 * Counter-based sampling of states on the device: x[i] = wrap(U(-std, std) + mean), the distribution of
 * Dynamics.get_initial_state (dynamics/dynamics_basic.py:28-29) and of the seed data set of controller/vhjb.py:136-151.
 * The reference's stream (the global NumPy RNG, one state per call) is serial; here sample i of the stream keyed by `seed`
 * is Philox4x32-10(key = seed, counter = (first + i, component / 4)) whichever GPU or launch produces it, so a batch split
 * over N GPUs (first = the shard's offset) is the same batch.  sys_kind selects the angle components that are wrapped.
 *   x [count, n] device.  oracle/x0_stream.py is the bit-exact NumPy twin.
 */
int hjb_sample_states(int32_t sys_kind, int32_t n, const float* mean, const float* std, uint64_t seed, int64_t first,
                      int64_t count, float* x, void* stream);

/* ==== vhjb: HJB-residual pass over sampled states (controller/vhjb.py:201-288) ====================== */
typedef enum hjb_activation {
  HJB_ACT_RELU = 0, /* controller/vhjb.py:55 */
  HJB_ACT_TANH = 1, /* examples/cartpole_balancing.ipynb cell 6 */
  HJB_ACT_SIN = 2   /* examples/double_integrator_optimal_time.ipynb cell 5 */
} hjb_activation;

typedef enum hjb_control_form {
  HJB_U_CLIPPED = 0, /* u = clip(-1/2 R^-1 g^T dV/dx + uf, umin, umax)   controller/vhjb.py:220           */
  HJB_U_BANGBANG = 1 /* u = -sign(g^T dV/dx)   examples/double_integrator_optimal_time.ipynb cell 11:17  */
} hjb_control_form;

typedef enum hjb_residual_form {
  /* loss = sum |vdot / (l(x,u) + eps) + 1| (1 - done) / (sum (1 - done) + eps)
   *        + reg * sum |V / (cost + eps) - 1| done / (sum done + eps)          controller/vhjb.py:227-253, 282-285 */
  HJB_RES_NORMALIZED = 0,
  /* loss = mean |vdot + l_i| with the running cost l_i given per sample in `costs`
   *        examples/double_integrator_optimal_time.ipynb cell 11:20 (l_i = 1[|x|^2 > 1e-4], cell 7:4)             */
  HJB_RES_MIN_TIME = 1
} hjb_residual_form;

/* Value network V(x) = |y|^2 + eps_s |z|^2, z = wrap(x - xf), h0 = (z - mean) / std,
 * y = act(act(h0 W1) W2) W3, no biases (ValueFunctionApproximator, controller/vhjb.py:17-60).
 * Kernels are stored Flax-style (in, out), row-major, back to back in ONE flat device buffer
 * params = [W1 (n x 128) | W2 (128 x 128) | W3 (128 x 64)]; features must be {128, 128, 64}.              */
typedef struct hjb_vnet {
  int32_t n;
  int32_t act; /* hjb_activation */
  int32_t features[3];
  int32_t impl; /* 0: tensor-core kernels (fp16 x 3, vhjb_tc.cuh); 1: fp32 CUDA-core kernels (vhjb_simt.cuh) — exact for
                   on-device cross-check of the tensor path (which sends what leaves its fp16 range management through the
                   same fp32 kernel by itself: hjb_vhjb_deferred) */
  const float* params; /* DEVICE pointer, 128 n + 24576 floats */
  float mean[HJB_MAX_N], std[HJB_MAX_N], xf[HJB_MAX_N];
  float eps_s; /* config.epsilon_scalar */
} hjb_vnet;

typedef struct hjb_task {
  float Q[HJB_MAX_N * HJB_MAX_N]; /* n x n row-major */
  float R[HJB_MAX_M * HJB_MAX_M]; /* m x m row-major */
  float Rinv[HJB_MAX_M * HJB_MAX_M];
  float uf[HJB_MAX_M];
  float eps;             /* config.epsilon */
  int32_t control_form;  /* hjb_control_form */
  int32_t residual_form; /* hjb_residual_form */
} hjb_task;

/* Number of floats in the flat parameter / gradient buffer of a value net on an n-dimensional state. */
int64_t hjb_vhjb_param_count(int32_t n);
/* Bytes of device scratch the vhjb entry points need (per-CTA partial sums; caller-owned). */
int64_t hjb_vhjb_workspace_bytes(int32_t n);

/* norm[0] = sum(1 - done) + eps, norm[1] = sum(done) + eps  (the denominators of controller/vhjb.py:241, 253);
 * MIN_TIME callers pass norm[0] = batch size instead.  With several GPUs all-reduce(sum) the two counts (before
 * eps is added: use eps = 0 here and add it after) so that every rank normalises by the GLOBAL batch. */
int hjb_vhjb_count(const float* dones, int64_t B, float eps, float* norm, void* workspace, void* stream);

/*
 * Residual only (rows V1-V5 of SURVEY.md 8a): for each sampled state the value V, its input gradient p = dV/dx
 * (get_v_gradient, controller/vhjb.py:201-202), the optimal control u (:204-221), and the residual argument r
 * (:228-233: vdot / (l + eps) + 1, or vdot + l_i).  Any of V [B], p [B, n], u [B, m], r [B] may be null.
 * sums (device, 2 floats, nullable): {sum |r| (1 - done), sum |V / (cost + eps) - 1| done} — un-normalised, so
 * shards can be added; MIN_TIME: {sum |r|, 0}.
 */
int hjb_vhjb_residual(const hjb_system* sys, const hjb_vnet* net, const hjb_task* task, const float* xs,
                      const float* dones, const float* costs, int64_t B, float* V, float* p, float* u, float* r,
                      float* sums, void* workspace, void* stream);

/*
 * Loss and parameter gradient of one batch (params_update's value_and_grad, controller/vhjb.py:282-285):
 *   grad (device, hjb_vhjb_param_count floats) = d/dparams [ L_hjb + reg * L_term ] restricted to this shard's
 *   samples, ALREADY divided by norm[] — summing the grads of all shards gives the full-batch gradient.
 *   sums (device, 2 floats) as in hjb_vhjb_residual.  norm (device, 2 floats) from hjb_vhjb_count.
 * Deterministic: per-CTA partial sums are reduced in a fixed order.
 */
int hjb_vhjb_loss_grad(const hjb_system* sys, const hjb_vnet* net, const hjb_task* task, const float* xs,
                       const float* dones, const float* costs, int64_t B, const float* norm, float reg, float* grad,
                       float* sums, void* workspace, void* stream);

/*
 * Same, for a batch that arrives in pieces (host batches streamed over PCIe chunk by chunk, VhjbKernels.train_step_host):
 * grad, sums and the saturation count are ADDED to what the buffers hold.  The first piece goes through
 * hjb_vhjb_loss_grad, the following ones through this entry point, all with the norm of the WHOLE batch; the result is
 * the full-batch gradient of controller/vhjb.py:282-285, summed piece by piece in launch order (deterministic).
 */
int hjb_vhjb_loss_grad_accumulate(const hjb_system* sys, const hjb_vnet* net, const hjb_task* task, const float* xs,
                                  const float* dones, const float* costs, int64_t B, const float* norm, float reg,
                                  float* grad, float* sums, void* workspace, void* stream);

/*
 * hjb_vhjb_loss_grad for a batch that is still ARRIVING from the host: one launch; the kernel reads the states / costs of
 * piece k (piece_states states each, a multiple of 64; the last piece may be shorter) only after ready[k] != 0.  The caller
 * queues, on a copy stream, H2D(piece k of xs and costs) followed by a 4-byte H2D write of a non-zero value to ready[k],
 * for k = 0, 1, ... (pieces become ready in order); dones and norm must be complete before the launch (the normalisers
 * are sums over the whole batch).  ready must be zeroed before the copies start.  The kernel reads the streamed buffers
 * with coherent loads (ld.global.cg) after an acquire of the piece's flag.  The poll is bounded (seconds) so that a flag
 * that is never set cannot hang the device; a poll that gives up is REPORTED, never silent: the loss sums of the step
 * become NaN, the workspace's stream-failure word is raised (hjb_vhjb_stream_failures) and hjb_vhjb_adam_guarded skips
 * the update.  Tensor-core kernels only (HJB_ERR_UNSUPPORTED otherwise: use the piecewise entry points above).
 */
/* The copy side of a streamed batch, queued in one call (a dozen pieces x 3 cudaMemcpyAsync): piece k of xs_host / costs_host
 * (PINNED host memory) -> xs / costs, then ones_host[k] (pinned, non-zero) -> ready[k], all on copy_stream.  Call it BEFORE
 * hjb_vhjb_loss_grad_streamed: a polling kernel must never wait for work that is queued behind it (two streams may share a
 * hardware queue). */
int hjb_vhjb_stream_batch(const float* xs_host, const float* costs_host, float* xs, float* costs, int64_t B, int32_t n,
                          int64_t piece_states, int32_t* ready, const int32_t* ones_host, void* copy_stream);
int hjb_vhjb_loss_grad_streamed(const hjb_system* sys, const hjb_vnet* net, const hjb_task* task, const float* xs,
                                const float* dones, const float* costs, int64_t B, const float* norm, float reg, float* grad,
                                float* sums, void* workspace, const int32_t* ready, int64_t piece_states, void* stream);

/* Number of warps whose wait for a piece of a streamed batch gave up, summed over the streamed launches on this workspace
 * since the last reset (`count`: device or pinned host float, nullable; reset != 0 zeroes the word after reading it). */
int hjb_vhjb_stream_failures(void* workspace, int32_t n, float* count, int32_t reset, void* stream);
/* hjb_adam on the value net's flat parameter buffer (n = state dimension), skipped — params, m, v untouched — while the
 * workspace's stream-failure word is raised: a gradient computed from data that never arrived is never applied. */
int hjb_vhjb_adam_guarded(float* params, float* m, float* v, const float* grad, int32_t n, float lr, float b1, float b2,
                          float eps, int32_t step, const void* workspace, void* stream);

/*
 * One whole single-GPU training step (VHJBController.params_update, controller/vhjb.py:255-288) in one call and three
 * launches: the normalisers (hjb_vhjb_count with this task's eps; MIN_TIME: {B, 1}), the fused loss + gradient kernel, and
 * one kernel that reduces the per-CTA partials in the same fixed order as hjb_vhjb_loss_grad and applies the optax.adam
 * update to net->params in place (step = 1-based update index).  Bit-identical to hjb_vhjb_count + hjb_vhjb_loss_grad +
 * hjb_adam; exists because the reference trains with minibatches of 256, where seven launches and as many host
 * round trips cost more than the arithmetic.  Multi-GPU callers keep the separate entry points (the all-reduce sits between
 * the gradient and Adam).  grad, sums, norm, the saturation count: as for hjb_vhjb_loss_grad.  loss_acc (device, 3 floats,
 * nullable): the step's losses {hjb + reg term, hjb, term} (the values params_update returns) are ADDED to it, so that a
 * training loop reads its epoch averages once instead of doing tensor arithmetic after every update.
 */
int hjb_vhjb_train_step(const hjb_system* sys, const hjb_vnet* net, const hjb_task* task, const float* xs, const float* dones,
                        const float* costs, int64_t B, float reg, float lr, float b1, float b2, float adam_eps, int32_t step,
                        float* m, float* v, float* norm, float* grad, float* sums, float* loss_acc, void* workspace,
                        void* stream);

/*
 * The same step on SEVERAL GPUs (one process per GPU, this rank's shard of the batch): the tail — reduction of the per-CTA
 * partials, exchange of [grad | loss sums | done-counts of the next batch] between the ranks, Adam — is ONE kernel that
 * writes its elements into every rank's exchange buffer over NVLink peer memory, publishes a flag with release semantics,
 * waits (bounded) for the peers' flags, sums the ranks' slots in rank order (identical bits on every rank) and updates the
 * weights: no NCCL call, no separate reduction / Adam launches.  `norm` = the GLOBAL normalisers of this batch (device, 2
 * floats: from counts_out of the previous call, or an all-reduce of hjb_vhjb_count); next_counts (nullable) = this rank's
 * [sum(1 - done), sum(done)] of the NEXT batch, whose global normalisers (sums + eps; {B, 1} for MIN_TIME) are returned in
 * counts_out — pass that buffer as `norm` of the next call (alternate two buffers).  peer_bufs / peer_flags:
 * device arrays [world] of pointers to every rank's exchange buffer (hjb_vhjb_peer_exchange_floats floats) and flag array
 * (hjb_vhjb_peer_exchange_flags uint32, zeroed once before the first step) in peer-accessible memory (e.g. torch symmetric
 * memory); step = 1, 2, ... identical on all ranks; world = 1 is accepted (the rank exchanges with itself: bit-identical
 * to hjb_vhjb_train_step).  A wait that gives up (~2 s) raises hjb_vhjb_stream_failures' word,
 * turns the loss sums into NaN and skips the update.
 */
int64_t hjb_vhjb_peer_exchange_floats(int32_t n, int32_t world);
int64_t hjb_vhjb_peer_exchange_flags(int32_t n, int32_t world);
int hjb_vhjb_train_step_peer(const hjb_system* sys, const hjb_vnet* net, const hjb_task* task, const float* xs, const float* dones,
                             const float* costs, int64_t B, float reg, float lr, float b1, float b2, float adam_eps, int32_t step,
                             float* m, float* v, const float* norm, float* grad, float* sums, float* loss_acc,
                             const float* next_counts, float* counts_out, void* const* peer_bufs, void* const* peer_flags,
                             int32_t rank, int32_t world, void* workspace, void* stream);

/*
 * Range management of the tensor-core gradient pass, and what is left of it for the caller to check.  The pass carries
 * per-state adjoints in fp16 with per-state power-of-two scaling (vhjb_tc.cuh).  A state whose adjoint seeds exceed 2^6
 * times the batch-typical weight (|x - xf|, |u - uf| of order 0.1 and below with the reference's eps = 1e-10; a terminal
 * sample stored with cost ~ 0) does NOT go through that chain: the kernel records its index and the fp32 CUDA-core kernel,
 * launched behind it by the same entry point, computes exactly those states; their weight gradients are summed after the
 * tensor kernel's partials in a fixed order (hjb_vhjb_deferred: how many states of the last launch took that pass).
 * Two things can still escape that management: a deferred list that is full (2048 states per epilogue warp of the tensor
 * kernel), or an adjoint chain that grows past fp16's 65504 between the seeds and the first layer (a backward gain above
 * ~1000: conversions saturate, nothing becomes inf / NaN).  The tensor kernel counts both, the fp32 pass behind it reads the
 * count and, when it is non-zero, runs the WHOLE launch in fp32; the reductions then leave the tensor launch's partials out
 * (hjb_vhjb_deferred == B for such a launch).  So every launch is exact, decided on the device, and the next launch is
 * back on the tensor cores.  hjb_vhjb_saturation returns the count of what was neither computed in range, nor deferred,
 * nor redone for the last hjb_vhjb_loss_grad on this workspace (device float `count`, stream-ordered): 0 by construction
 * for the gradient entry points of this library; kept in the ABI for callers that poll it (a non-zero value would mean:
 * set hjb_vnet.impl = 1, the fp32 CUDA-core kernels).  The workspace must be zero-filled before its first use.
 */
int hjb_vhjb_saturation(const void* workspace, int32_t n, float* count, void* stream);
/* The same count summed over every gradient launch on this workspace since the last reset (count nullable; reset != 0
 * zeroes the total after reading it; zero it once — or zero-fill the workspace — before the first use): what a training
 * loop polls once per epoch instead of once per update. */
int hjb_vhjb_saturation_total(void* workspace, int32_t n, float* count, int32_t reset, void* stream);
/* States of the last gradient launch that were computed by the fp32 pass instead of the tensor chain (see above);
 * `count`: device or pinned host float. */
int hjb_vhjb_deferred(const void* workspace, int32_t n, float* count, void* stream);

/* ==== the notebooks' soft-PD baseline ===========================================
 * SoftPDValueApproximator (examples/cartpole_balancing.ipynb cell 6, examples/drone_hovering.ipynb cell 6): an unconstrained
 * value net  z = wrap(x - xf) -> Dense(128) -> s -> Dense(128) -> s -> Dense(64) -> s -> Dense(1), every Dense WITH bias,
 * s = tanh (cart-pole notebook) or relu (drone notebook); no input normalisation.  params: one flat fp32 device buffer
 * [W1 (n x 128) | b1 | W2 (128 x 128) | b2 | W3 (128 x 64) | b3 | w4 (64) | b4], Flax (in, out) kernels.               */
typedef struct hjb_softpd {
  int32_t n;
  int32_t act;                 /* HJB_ACT_TANH or HJB_ACT_RELU */
  int32_t normalized_residual; /* 0: |vdot + l| (cart-pole nb cell 11); 1: |vdot / (l + eps) + 1| (drone nb cell 11) */
  const float* params;
  float xf[HJB_MAX_N], uf[HJB_MAX_M];
  float Q[HJB_MAX_N * HJB_MAX_N], R[HJB_MAX_M * HJB_MAX_M], Rinv[HJB_MAX_M * HJB_MAX_M];
  float K[HJB_MAX_M * HJB_MAX_N]; /* LQR gain, loss form 2 */
  float P[HJB_MAX_N * HJB_MAX_N]; /* Riccati solution, loss form 1 */
  float eps;
} hjb_softpd;
int64_t hjb_softpd_param_count(int32_t n);     /* 128 n + 128 + 128*128 + 128 + 128*64 + 64 + 64 + 1 */
int64_t hjb_softpd_workspace_bytes(int32_t n); /* caller-owned scratch for the two calls below */
/* V [B], p = dV/dx [B, n], u = clip(-R^-1 g^T p / 2 + uf) [B, m] (each nullable) — get_soft_pd_control_with_additional_term */
int hjb_softpd_policy(const hjb_system* sys, const hjb_softpd* net, const float* xs, int64_t B, float* V, float* p, float* u,
                      void* workspace, void* stream);
/* Loss and its parameter gradient (mean over the batch, as jax.value_and_grad of the notebooks' losses returns them):
 *   loss_form 0  soft_pd_hjb_loss:         mean_i [ res_i + reg max(0, V(xf) - V(x_i)) ],  u from the net
 *   loss_form 1  soft_pd_warmup_hjb_loss of the cart-pole notebook:  mean_i |V(x_i) - z_i^T P z_i|
 *   loss_form 2  soft_pd_warmup_hjb_loss of the drone notebook: form 0 with u = clip(-K z + uf)
 * grad [param_count]; sums [3] = {sum_i res_i (form 1: sum_i |V - z^T P z|), sum_i hinge_i, #{i: V(x_i) < V(xf)}}.      */
int hjb_softpd_loss_grad(const hjb_system* sys, const hjb_softpd* net, const float* xs, int64_t B, int32_t loss_form, float reg,
                         float* grad, float* sums, void* workspace, void* stream);

/*
 * One step of the learned-policy rollout for N trajectories at once (VHJBController.rollout_trajectory,
 * controller/vhjb.py:171-193; the first widening row of SURVEY.md 8f).  `u` [N, m] is the value-net policy's control at
 * `x` [N, n] (hjb_vhjb_residual's u output).  For every trajectory with alive != 0: if wrap(x - xf) left the box
 * [obs_lo, obs_hi] (:176-177) or `terminal` is set (:188-191) it ends with the sample (x, dx^T P dx, done = 1)
 * (P row-major n x n: the Riccati terminal cost, :156-160, :167-169); otherwise the sample is (x, l(x, u) dt, 0)
 * (:162-165, :184) and x <- Dynamics.simulate(x, u).  total_cost accumulates the sample costs; rec_* (nullable, one
 * time slice) receive the sample, rec_done = -1 where the trajectory had already ended.  xf, obs_lo, obs_hi, P: host.
 */
int hjb_policy_step(const hjb_system* sys, const hjb_task* task, const float* xf, const float* obs_lo, const float* obs_hi,
                    const float* P, int32_t terminal, float* x, const float* u, float* alive, float* total_cost, float* rec_x,
                    float* rec_cost, float* rec_done, int64_t N, void* stream);

/*
 * The whole learned-policy rollout of N trajectories (VHJBController.rollout_trajectory, controller/vhjb.py:171-193, for all
 * of an epoch's trajectories at once) in ONE call: T + 1 rounds of hjb_vhjb_residual (-> u, skipped in the closing round)
 * and hjb_policy_step, 2 T + 1 launches queued on the stream without returning to the caller.  x [N, n] in / out,
 * u [N, m] scratch, zeros / ones [N] (the dones / costs the residual entry point wants), alive [N] = 1 and total_cost [N] = 0
 * on entry; rec_x [T + 1, N, n], rec_cost / rec_done [T + 1, N] receive every time slice (rec_done = -1: no sample).
 */
int hjb_policy_rollout(const hjb_system* sys, const hjb_vnet* net, const hjb_task* task, const float* xf, const float* obs_lo,
                       const float* obs_hi, const float* P, int32_t T, float* x, float* u, const float* zeros, const float* ones,
                       float* alive, float* total_cost, float* rec_x, float* rec_cost, float* rec_done, int64_t N,
                       void* workspace, void* stream);

/*
 * Device-resident replay buffer (second half of SURVEY.md 8f row 1).  The reference keeps (state, cost, done) samples in
 * a deque(maxlen) behind a torch Dataset / shuffling DataLoader (controller/vhjb.py:62-73, :153-154) and extends it
 * trajectory by trajectory (:299-305).  Here the samples are a ring of `capacity` rows in HBM (buf_x [capacity, n],
 * buf_cost, buf_done [capacity]; caller-owned, like the ring's head / size bookkeeping).
 *
 * hjb_replay_append: the records of a batched rollout (hjb_policy_step's rec_* slices stacked over time: rec_x
 * [T1, N, n], rec_cost / rec_done [T1, N], rec_done < 0 = no sample) go into the ring in deque order — trajectory by
 * trajectory, time order within a trajectory: sample (t, e) has sequence number offsets[e] + t (offsets: device,
 * exclusive prefix sum of the per-trajectory sample counts) and lands in row (tail + seq) mod capacity; samples with
 * seq < skip (= max(0, total - capacity): they would be pushed out again by this same extend) are dropped.
 * hjb_replay_gather: minibatch rows xs [B, n], costs [B], dones [B] <- ring rows index[0..B) (device int64: the
 * caller's shuffled permutation, DataLoader(shuffle=True, drop_last=True)).
 */
int hjb_replay_append(const float* rec_x, const float* rec_cost, const float* rec_done, const int64_t* offsets, int64_t T1,
                      int64_t N, int32_t n, int64_t skip, int64_t tail, int64_t capacity, float* buf_x, float* buf_cost,
                      float* buf_done, void* stream);
int hjb_replay_gather(const float* buf_x, const float* buf_cost, const float* buf_done, const int64_t* index, int64_t B,
                      int32_t n, float* xs, float* costs, float* dones, void* stream);

/* optax.adam update (controller/vhjb.py:120, 286-287; defaults b1 = 0.9, b2 = 0.999, eps = 1e-8), in place on the
 * flat buffers; `step` is the 1-based index of this update. */
int hjb_adam(float* params, float* m, float* v, const float* grad, int64_t len, float lr, float b1, float b2,
             float eps, int32_t step, void* stream);

/* ---- misc ------------------------------------------------------------------------------------------ */
int hjb_abi_version(void);
const char* hjb_status_string(int status);
/* FP32 FMA micro-benchmark used as the CUDA-core roofline denominator: runs `iters` dependent FMAs on 8
 * independent chains per thread over a full-chip grid and writes one float per thread to `sink` (>= grid*block
 * floats).  Returns the number of FLOPs executed through *flops. */
int hjb_fma_peak_probe(float* sink, int64_t sink_len, int32_t iters, double* flops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HJB_B200_H */
