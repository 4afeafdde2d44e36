/*
 * visfly_b200.h — C-ABI of the B200-native quadrotor dynamics engine.
 *
 * This is the drop-in boundary for ONE hot path of SJTU-ViSYS-team/VisFly: the batched rigid-body
 * control step `Dynamics.step` (reference envs/base/dynamics.py:319-372) with its integrator
 * (reference utils/maths.py:300-389) and the reverse-mode gradient that PyTorch autograd would
 * otherwise build op by op (consumers: reference utils/algorithms/BPTT.py:107-134).
 *
 * Rules of the boundary
 *   - plain C linkage, plain pointers and sizes, no torch / C++ types in any signature;
 *   - every buffer is allocated and owned by the caller; nothing is retained after return;
 *   - all launches are asynchronous on the `stream` argument (a cudaStream_t passed as void*;
 *     NULL = legacy default stream) and are CUDA-graph capturable; only the *_host entry points
 *     synchronise (they have to: they hand results back in host memory);
 *   - return value 0 = success, non-zero = error; `vf_last_error()` describes the last failure of
 *     the calling thread;
 *   - the four step kernels (vf_step_fwd/bwd, vf_env_step_fwd/bwd) are launched with programmatic stream
 *     serialization (sm_90+): their CTAs may become resident while the preceding kernel on `stream` is still
 *     running, but they touch no memory before that kernel has completed and its writes are visible
 *     (griddepcontrol.wait), so stream order is what a caller observes.  Kernels of other libraries
 *     before or after them need nothing special.  VF_NO_PDL=1 in the environment turns the attribute off.
 *
 * HBM layout of the agent state ("packed state", float32):
 *
 *     state[plane][agent][lane]      plane in 0..4, agent in 0..n-1, lane in 0..3   (5*n*4 floats)
 *
 *     plane 0 : p.x   p.y   p.z   alpha.x        position            | angular acceleration x
 *     plane 1 : q.w   q.x   q.y   q.z            orientation quaternion (w first)
 *     plane 2 : v.x   v.y   v.z   alpha.y        ground velocity     | angular acceleration y
 *     plane 3 : w.x   w.y   w.z   alpha.z        body rates          | angular acceleration z
 *     plane 4 : W0    W1    W2    W3             motor speeds (rad/s)
 *
 * i.e. structure-of-arrays of 16-byte groups: a warp reads one plane as 512 contiguous bytes with one
 * 128-bit load per thread.  These 20 floats are exactly the quantities that carry value (and gradient)
 * across control steps in the reference: p, q, v, w, motor speed, and the angular
 * acceleration that feeds the body-rate controller's D term (reference dynamics.py:407).
 *
 * Actions are row-major (n,4) float32 in [-1,1] — the reference's own layout (dynamics.py:704-710) —
 * so one agent's action is one 128-bit load as well.
 */
#ifndef VISFLY_B200_H
#define VISFLY_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define VF_ABI_VERSION 8

/* integrator: reference `integrator=` kwarg, utils/maths.py:331 (euler) and :353 (rk4, repaired R1-R3) */
#define VF_INTEGRATOR_EULER 0
#define VF_INTEGRATOR_RK4   1

/* action type: values of reference utils/type.py:14-18 ACTION_TYPE */
#define VF_ACTION_THRUST   0
#define VF_ACTION_BODYRATE 1
#define VF_ACTION_VELOCITY 2   /* [-, vx, vy, vz] set-point, yaw follows the velocity   dynamics.py:414-454 */
#define VF_ACTION_POSITION 3   /* [yaw, x, y, z] set-point                              dynamics.py:455-496 */

/* flags */
#define VF_FLAG_CTRL_DELAY 1u  /* first-order motor lag, reference dynamics.py:510-516; off => T = T_des (:518) */

#define VF_STATE_PLANES 5
#define VF_STATE_FLOATS 20     /* floats per agent in the packed state  */
#define VF_OBS_FLOATS   13     /* [p(3) q(4) v+wind(3) w(3)], reference dynamics.py:778-786 */
#define VF_EXT_FLOATS   8      /* [acc(3) 0 thrust(4)], reference dynamics.py:347,516/518 */
#define VF_MAX_SUBSTEPS_BWD 64 /* reverse sweep keeps per-substep inputs in thread-local memory */

/* Everything the control step needs besides per-agent data.  All matrices row-major.
 * Filled on the host from the drone JSON exactly as the reference does (dynamics.py:562-689, :94-114). */
typedef struct VfParams {
    float dt;              /* integration sub-step                               dynamics.py:69      */
    float mass;            /*                                                    :565                */
    float inv_mass;
    float J[3];            /* diagonal inertia                                   :108                */
    float J_inv[3];        /*                                                    :110                */
    float B[16];           /* allocation  [F,tx,ty,tz] = B @ thrusts             :111-113            */
    float B_inv[16];       /*                                                    :114                */
    float thrust_map[3];   /* T = a W^2 + b W + c                                :579, :530-534      */
    float motor_c;         /* exp(-dt / motor_tau)                               :581                */
    float thrust_min;      /* 0                                                  :592                */
    float thrust_max;      /* thrust at motor_omega_max                          :587-590            */
    float k_lin[3];        /* linear body drag                                   :568                */
    float k_quad[3];       /* 0.5 * 1.225 * Cd * A                               :567                */
    float JKp[9];          /* J @ Kp of the body-rate PID                        :405                */
    float Kd[9];           /* Kd of the body-rate PID                            :407                */
    float act_half[4];     /* de-normalisation: cmd = a * half + mean            :704-713            */
    float act_mean[4];
    float gravity[3];      /* (0,0,-9.81)                                        :15                 */
    float wind[3];         /* constant wind, added to p-dot and to reported v    :135, maths.py:310  */
    float pos_lo[3];       /* post-step clamps ("_ugly_fix")                     :374-382            */
    float pos_hi[3];
    float vel_lim;
    float rate_lim;
    /* outer loops of the velocity / position action types (geometric attitude controller) */
    float Kp[9];           /* Kp of the body-rate PID (un-multiplied by J)       :452, :486-491      */
    float vel_kp;          /* VELOCITY_PID.p                                      :416                */
    float vel_kd;          /* VELOCITY_PID.d                                      :433, :458          */
    float pos_kd;          /* POSITION_PID.d                                      :457, :469          */
} VfParams;

int         vf_abi_version(void);
const char* vf_last_error(void);
int         vf_params_size(void);   /* sizeof(VfParams) as compiled into the library (binding self-check) */

/* Number of streaming multiprocessors of the current device (grid sizing / bench reporting). */
int vf_device_sm_count(void);

/*
 * One control step for n agents = reference Dynamics.step (dynamics.py:319-372) minus the host-side
 * comm-delay FIFO (the caller hands in the already-delayed action, dynamics.py:323-328).
 *
 *   state_in   [5][n][4]  packed state at the start of the step           (device, 16B aligned)
 *   action     [n][4]     normalised action in [-1,1]                     (device, 16B aligned)
 *   state_out  [5][n][4]  packed state after ctrl_dt = substeps*dt        (device, must not alias state_in)
 *   obs_out    [n][13]    reference `state` property, or NULL             (device, 16B aligned)
 *   ext_out    [n][8]     [acc, 0, thrusts] diagnostics, or NULL          (device, 16B aligned)
 *   wind       [n][4]     per-agent wind [wx, wy, wz, -] of this control step, or NULL = the constant
 *                         params->wind.  Carries the reference's time-varying wind functions
 *                         (dynamics.py:136-165, update_wind :384-388: evaluated by the caller once per control
 *                         step from t and the previous wind, then frozen over the sub-steps) (device, 16B aligned)
 *   fifo_push  [n][4]     the action the caller handed to step() THIS control step, or NULL     (device, 16B aligned)
 *   fifo_copy  [n][4]     receives a copy of fifo_push: the engine-owned comm-delay FIFO entry that will be consumed
 *                         comm_delay/ctrl_dt steps later — the reference's `action.T.clone()` (dynamics.py:324) done by
 *                         the same launch (+32 B per agent), so the caller may overwrite its action buffer right away.
 *                         NULL together with fifo_push.                                         (device, 16B aligned)
 */
int vf_step_fwd(const VfParams* params, int n, int substeps, int integrator, int action_type,
                unsigned flags, const float* state_in, const float* action,
                float* state_out, float* obs_out, float* ext_out, const float* wind,
                const float* fifo_push, float* fifo_copy, void* stream);

/* The comm-delay FIFO as a device-resident ring: `depth` engine-owned rows of [n][4] floats (16-byte aligned), row[0]
 * the oldest.  Used where addresses must not change from step to step (a step recorded as a CUDA graph):
 *   vf_step_fwd_ring consumes row[0] as the delayed action, moves row[r+1] -> row[r] and stores fifo_push into
 *                    row[depth-1] — per agent, in place, in the step's own launch (dynamics.py:323-328);
 *   vf_env_finish    zeroes the rows of the agents it re-initialises (dynamics.py:262-263). */
#define VF_FIFO_MAX_ROWS 8
typedef struct VfFifoRows {
    float* row[VF_FIFO_MAX_ROWS];
    int    depth;              /* 1..VF_FIFO_MAX_ROWS */
} VfFifoRows;

/* vf_step_fwd with the FIFO shifted inside the launch: arguments as vf_step_fwd, `action` replaced by the ring. */
int vf_step_fwd_ring(const VfParams* params, int n, int substeps, int integrator, int action_type, unsigned flags,
                     const float* state_in, const VfFifoRows* fifo_rows, const float* fifo_push, float* state_out,
                     float* obs_out, float* ext_out, const float* wind, void* stream);

/*
 * Reverse-mode gradient of vf_step_fwd: re-runs the substeps from (state_in, action) in registers /
 * thread-local memory, then sweeps them in reverse.  No forward intermediates are ever stored in HBM.
 *
 *   grad_state_out [5][n][4]  dL/d state_out                    (device; NULL = zeros)
 *   grad_obs       [n][13]    dL/d obs_out                      (device; NULL = zeros)
 *   grad_state_in  [5][n][4]  dL/d state_in      (written)
 *   grad_action    [n][4]     dL/d action        (written)
 *   wind           [n][4]     the wind the forward step was given, or NULL (it only decides the gradient gate
 *                             of the post-step position clamp; wind itself does not depend on the state)
 *
 * torch.autograd conventions are reproduced exactly: clamp passes gradient on the
 * closed interval, d(v|v|)/dv = 2|v|.  substeps must be <= VF_MAX_SUBSTEPS_BWD.
 * Only VF_ACTION_THRUST / VF_ACTION_BODYRATE: the reference's own autograd graph is broken for the velocity and
 * position action types (in-place row writes in a per-agent loop, dynamics.py:446-450 — backward raises), so
 * there is no gradient to be faithful to; the call returns an error for them.
 */
int vf_step_bwd(const VfParams* params, int n, int substeps, int integrator, int action_type,
                unsigned flags, const float* state_in, const float* action,
                const float* grad_state_out, const float* grad_obs,
                float* grad_state_in, float* grad_action, const float* wind, void* stream);

/*
 * Host-buffer variant of vf_step_fwd (what a caller that lives in host memory binds, e.g. an SB3/numpy
 * training loop: reference envs/base/droneGymEnv.py:218 returns numpy).  Copies `action_host` to the
 * device, runs the step on the caller's device-resident state, copies obs back and synchronises the stream.
 * `action_dev` / `obs_dev` are caller-owned device staging buffers ([n][4] and [n][13]); host buffers
 * should be page-locked for the copies to be asynchronous.
 */
int vf_step_fwd_host(const VfParams* params, int n, int substeps, int integrator, int action_type,
                     unsigned flags, const float* state_in, const float* action_host, float* action_dev,
                     float* state_out, float* obs_dev, float* obs_host, void* stream);

/*
 * Layout conversion between the reference's row-major per-field tensors and the packed state
 * (reference Dynamics.reset takes (n,k) row-major inputs, dynamics.py:229-236; properties return them,
 * :735-776).  Any input pointer may be NULL (pack: field left zero, identity quaternion; unpack: skipped).
 * `index` is NULL (all agents, dense) or int64[m] agent ids to scatter into (pack) — fields are then [m][k].
 */
int vf_pack_state(int n, int m, const long long* index,
                  const float* pos, const float* quat, const float* vel, const float* rate,
                  const float* motor, const float* alpha, float* state, void* stream);
int vf_unpack_state(int n, const float* state, float* pos, float* quat, float* vel, float* rate,
                    float* motor, float* alpha, void* stream);

/*
 * Renderer hand-off: poses of all agents in Habitat-sim's frame, straight from the packed
 * state.  Replaces the reference's per-step `std_to_habitat` on host copies (utils/common.py:131-179, called from
 * SceneManager.set_pose, utils/SceneManager.py:347-348, with Dynamics.position / orientation / velocity,
 * envs/base/droneEnv.py:376):
 *     hab_pos = (-y,  z, -x)            hab_ori (w first) = (w, -y, z, -x)            hab_vel = like hab_pos
 *   pose_out [n][7]  = [hab_pos, hab_ori] per agent — the row the reference appends to its trajectory log
 *                      (SceneManager.py:355-357)
 *   vel_out  [n][3]  = velocity incl. the constant wind (Dynamics.velocity, dynamics.py:750-752), or NULL
 * Both outputs may be device memory or page-locked host memory (written zero-copy, see VfEnvMirror); the host-side
 * consumer synchronises the stream before reading.
 */
int vf_export_pose_habitat(const VfParams* params, int n, const float* state, float* pose_out, float* vel_out,
                           void* stream);

/* =====================================================================================================
 * Fused env step: the control step above PLUS the wrapper tail the reference runs
 * around it with ~100 aten ops and per-agent Python loops — analytic bounding-box collision
 * (envs/base/droneEnv.py:345-369), the task's success test and reward (envs/HoverEnv.py:79-94,
 * envs/NavigationEnv.py:81-99, envs/RacingEnv.py:142-148,203-215), reward accumulation and termination
 * (envs/base/droneGymEnv.py:163-193), the per-episode record the reference puts into `info`
 * (droneGymEnv.py:238-275) and the auto-reset of finished agents with freshly sampled initial states
 * (droneGymEnv.py:207-208,339-349; droneEnv.py:260-288; utils/randomization.py:153-169,278-296) — in one launch.
 * ===================================================================================================== */

#define VF_TASK_HOVER      0
#define VF_TASK_NAVIGATION 1
#define VF_TASK_RACING     2
#define VF_TASK_CUSTOM     3   /* task code lives with the caller: only vf_env_finish accepts it */

#define VF_OBS_STATE13  0   /* [p q v w]                                    reference `state`, dynamics.py:778-786 */
#define VF_OBS_RACING16 1   /* [gate-p, gate2-p]/10, q, v/10, w/10          reference RacingEnv.py:254-262          */

#define VF_GEN_UNIFORM 0    /* (2u-1)*half + mean per field, utils/randomization.py:153-169; a Union of several    */
#define VF_GEN_NORMAL  1    /* (2z-1)*std + mean, :201-204                  boxes picks one uniformly (:278-296)     */
#define VF_GEN_TABLE   2    /* agent i restarts from row i of a caller-supplied [n][13] table                       */
#define VF_GEN_MAX_BOXES 4

#define VF_ENV_FLAG_NO_RESET 1u   /* is_test=True: report done but do not re-initialise (droneGymEnv.py:207)        */

/* per-agent env status bits (low byte of VfEnvStatus word 2) */
#define VF_EBIT_EPISODE_DONE  1u
#define VF_EBIT_ONCE_COLLIDED 2u
/* bits of the per-step episode record */
#define VF_RBIT_DONE          1u
#define VF_RBIT_EPISODE_DONE  2u
#define VF_RBIT_SUCCESS       4u
#define VF_RBIT_TRUNCATED     8u
#define VF_RBIT_COLLIDED     16u

/* Per-agent env status: ONE 16-byte record per agent (one 128-bit load and store), int32 status[n][4]:
 *   word 0  step_count   steps since the agent's last reset                         (droneGymEnv.py:163)
 *   word 1  returns      accumulated episode reward, float32 bit pattern            (droneGymEnv.py:185)
 *   word 2  VF_EBIT_* in bits 0..7, next gate index in bits 8..15 (racing)          (RacingEnv.py:142-148)
 *   word 3  gates passed this episode (racing)
 * The step reads status_in and writes status_out; the two may be the same buffer (in place) or different ones
 * (the start-of-step record then stays intact: it is what vf_env_step_bwd needs, and what a caller that runs
 * steps ahead of its consumer rewinds to). */
#define VF_STATUS_WORDS 4

typedef struct VfEnvSpec {
    int   task;               /* VF_TASK_*                                                                        */
    int   obs_kind;           /* VF_OBS_*                                                                         */
    int   max_episode_steps;  /* droneGymEnv.py:193                                                               */
    int   collision_reset;    /* is_collision_reset, droneGymEnv.py:189                                           */
    int   fifo_depth;         /* comm-delay steps: a delayed action older than the agent's episode reads as zero  */
                              /* (= the reference zeroing the FIFO rows of reset agents, dynamics.py:262-263)     */
    float uav_radius;         /* droneEnv.py:31,367                                                               */
    float bbox_lo[3];         /* droneEnv.py:129                                                                  */
    float bbox_hi[3];
    float target[3];          /* hover / navigation target                                                        */
    float success_radius;
    int   n_gates;            /* racing                                                                           */
    float gates[4][3];
    int   gen_kind;           /* VF_GEN_*                                                                         */
    int   gen_boxes;          /* 1, or the number of boxes of a Union generator                                   */
    float gen_mean[VF_GEN_MAX_BOXES][4][3];   /* [box][position, euler orientation, velocity, body rate][xyz]     */
    float gen_half[VF_GEN_MAX_BOXES][4][3];   /* half width (uniform) or std (normal)                             */
    int   gen_heading[VF_GEN_MAX_BOXES];      /* Uniform(heading=True): yaw points from the drawn position to the  */
                                              /* centre of the position box, randomization.py:27-29,162-165       */
    float init_motor_omega;   /* hover rotor speed after reset, dynamics.py:86                                    */
    unsigned long long seed;  /* Philox key of the reset sampler                                                  */
    unsigned agent_offset;    /* global index of this batch's agent 0: the sampler is keyed by (seed, agent_offset + i, */
                              /* step), so a shard of a larger batch draws exactly the restarts the whole batch would  */
} VfEnvSpec;

int vf_env_spec_size(void);

/* Optional second destination of what env.step hands back to a caller that lives in HOST memory (the reference's
 * numpy output mode, envs/base/droneGymEnv.py:218).  The pointers are page-locked host buffers (cudaHostAlloc /
 * cudaHostRegister; the library resolves their device alias with cudaHostGetDevicePointer): the kernel stores
 * observation, reward and done flag there directly (zero-copy over PCIe, overlapped with the arithmetic of the other
 * warps) in addition to the device outputs, so the step needs no device->host copy afterwards.  Any of obs / reward /
 * done may be NULL.
 * Completion word: if `flag` is set, the LAST thread block of the launch to finish stores `flag_value` into the
 * page-locked word `flag` after every block's host stores are visible system-wide (each block: fence.sys, then one
 * atomic on the device word `counter`, which must be zero before the launch and is zero again after it).  A host
 * thread that spins on `*flag == flag_value` (vf_wait_flag) may then read the mirror without a stream
 * synchronisation — and without waiting for whatever else was queued on the stream behind this launch. */
typedef struct VfEnvMirror {
    float*    obs;         /* [n][13|16]                                                          */
    float*    reward;      /* [n]                                                                 */
    int*      done;        /* [n]  0/1 as int32 (the reference returns done.astype(np.int32))     */
    unsigned* flag;        /* page-locked host word, or NULL                                      */
    unsigned* counter;     /* device word (zero), required with flag                              */
    unsigned  flag_value;  /* what the last block stores into *flag                               */
} VfEnvMirror;

/* Fused compute + collective: the one collective of the path is an all-gather of per-agent episode returns once per
 * rollout.  Instead of a collective launched after the rollout's last step, that step's own launch
 * stores every agent's accumulated return straight into the gather buffer of EVERY rank through peer-mapped device
 * memory (NVLink / NVSwitch; `dst[r]` = rank r's buffer as mapped into this process, e.g. from
 * torch.distributed._symmetric_memory): agent i of this shard lands at dst[r][offset + i].  What remains of the
 * collective is a cross-rank barrier on the stream (the stores are complete when the kernel is).  world = 0 disables. */
#define VF_MAX_PEERS 8
typedef struct VfPeerScatter {
    float*    dst[VF_MAX_PEERS];   /* gather buffer of each rank (this rank's own included), float[n_total]      */
    long long offset;              /* global index of this shard's agent 0                                       */
    int       world;               /* number of ranks (<= VF_MAX_PEERS); 0 = nothing to do                       */
} VfPeerScatter;

/* Spin until the page-locked word *flag equals value (see VfEnvMirror).  Returns 0, or 1 after timeout_us
 * microseconds (timeout_us <= 0: wait forever). */
int vf_wait_flag(const volatile unsigned* flag, unsigned value, long long timeout_us);

/*
 * One env step for n agents.
 *   state_in/action/state_out, wind, fifo_push/fifo_copy   as vf_step_fwd (action = output of the comm-delay FIFO);
 *                 the reported velocity, the rewards and the observation use v + wind of this step like the reference
 *   status_in   int32[n][4]   VfEnvStatus records at the start of the step (see above)
 *   status_out  int32[n][4]   records after the step (and after the auto-reset); may alias status_in
 *   step_index              counter mixed into the reset sampler (the env's global step number) ...
 *   step_base   device uint64*, or NULL: ... plus *step_base, read by the kernel — a caller that replays a captured
 *                 CUDA graph bumps this word between replays so that every replay draws fresh restarts
 *   reset_table float[n][13]  rows [p q v w] for VF_GEN_TABLE, else NULL
 *   obs_out     float[n][13|16]  observation AFTER the auto-reset (what env.step returns)
 *   reward_out  float[n], done_out uint8[n]
 *   record_out  float[n][4]   [episode return, episode length, VF_RBIT_* as float, gates passed] of this step
 *   term_obs_out float[n][13|16] or NULL: observation BEFORE the reset, written only for finished agents
 *   gate_out    int64[n] or NULL: next gate index after the step (racing) as the tensor the observation dict carries
 *                 (reference RacingEnv.py:267 `"gate": self._next_target_i`), so that no per-step unpacking is needed
 *   host_mirror NULL, or page-locked host destinations written in addition to obs_out / reward_out / done_out
 *   peer_returns NULL, or where to scatter the accumulated episode returns of this step (see VfPeerScatter)
 */
int vf_env_step_fwd(const VfParams* params, const VfEnvSpec* spec, int n, int substeps, int integrator,
                    int action_type, unsigned flags, unsigned env_flags, unsigned long long step_index,
                    const unsigned long long* step_base,
                    const float* state_in, const float* action, const float* wind, const float* fifo_push,
                    const float* reset_table, const int* status_in,
                    float* state_out, int* status_out, float* fifo_copy, float* obs_out, float* reward_out,
                    unsigned char* done_out, float* record_out, float* term_obs_out, long long* gate_out,
                    const VfEnvMirror* host_mirror, const VfPeerScatter* peer_returns, void* stream);

/*
 * Reverse mode of vf_env_step_fwd with respect to (state_in, action), given the gradients of its differentiable
 * outputs: the packed state, the returned observation and the reward.  Re-runs the step from its inputs (nothing
 * else was stored: status_in is the forward's own input buffer, kept intact by an out-of-place status_out), folds in
 * the adjoint of the task reward (torch.autograd conventions, see csrc/vf_env.cuh) and of the observation layout, then
 * the adjoint of the control step.  An agent that finished in this step was re-initialised inside it: its returned
 * state/observation are constants, only its reward carries gradient — the same graph the reference builds by
 * overwriting rows in place (dynamics.py:249-263).
 *   status_in    int32[n][4]  the forward's status_in          grad_state_out [5][n][4] or NULL
 *   grad_obs     float[n][13|16] or NULL                        grad_reward    float[n] or NULL
 *   wind         [n][4] the forward's per-agent wind, or NULL
 */
int vf_env_step_bwd(const VfParams* params, const VfEnvSpec* spec, int n, int substeps, int integrator,
                    int action_type, unsigned flags, unsigned env_flags, const float* state_in,
                    const float* action, const float* wind, const int* status_in, const float* grad_state_out,
                    const float* grad_obs, const float* grad_reward, float* grad_state_in, float* grad_action,
                    void* stream);

/*
 * The wrapper tail WITHOUT the task: for envs whose success / failure / reward are the caller's own tensor code
 * (subclasses of the reference's DroneGymEnvsBase overriding get_success / get_failure / get_reward).  The caller runs
 * vf_step_fwd, evaluates its task on the state reached, and hands the three per-agent results to this launch, which
 * does the rest of `DroneGymEnvsBase.step` (envs/base/droneGymEnv.py:163-208): step count, bounding-box collision and
 * out-of-bounds flags (droneEnv.py:345-369), return accumulation, episode_done / done, the episode record, and the
 * auto-reset of finished agents from the state generator (same sampler, same spec fields as vf_env_step_fwd; spec.task
 * may be VF_TASK_CUSTOM, target / gates are not read).  Two launches + the caller's task ops per env step instead of
 * ~100 small tensor kernels.
 *   state_in   [5][n][4]  state AFTER the control step            reward float[n]   success / failure uint8[n]
 *   state_out  [5][n][4]  = state_in, finished agents re-initialised (must not alias state_in)
 *   obs_out    [n][13] or NULL: the reference `state` after the reset      done_out uint8[n]     record_out float[n][4]
 *   fifo_rows  NULL, or the comm-delay FIFO rows to zero for re-initialised agents (VfFifoRows above)
 */
int vf_env_finish(const VfParams* params, const VfEnvSpec* spec, int n, unsigned env_flags,
                  unsigned long long step_index, const unsigned long long* step_base, const float* state_in,
                  const float* wind, const float* reset_table, const int* status_in, const float* reward,
                  const unsigned char* success, const unsigned char* failure, float* state_out, int* status_out,
                  float* obs_out, unsigned char* done_out, float* record_out, const VfFifoRows* fifo_rows,
                  void* stream);

/*
 * Renderer hand-off, ingestion side: what the reference does on the host, per agent and sensor,
 * to the images a renderer returns (envs/base/droneEnv.py:296-331: np.stack, expand_dims / transpose,
 * np.where(depth == 0, 20, depth)) — as one streaming launch per sensor over a batched image buffer.  `src` is device
 * memory or the page-locked host buffer the renderer wrote (read zero-copy); `dst` is device memory.
 *   vf_ingest_depth : float (n, H, W) -> (n, 1, H, W), zeros (no return) replaced by `background` (the reference uses 20)
 *   vf_ingest_color : uint8 (n, H, W, 4) RGBA -> (n, 3, H, W), alpha dropped; H * W must be a multiple of 4
 */
int vf_ingest_depth(long long n_images, int height, int width, const float* src, float* dst, float background,
                    void* stream);
int vf_ingest_color(long long n_images, int height, int width, const unsigned char* src, unsigned char* dst,
                    void* stream);
const char* vf_sensor_last_error(void);

/* =====================================================================================================
 * Deterministic actor of the analytic-gradient trainers (SURVEY.md §8 row n3 / BASELINE configs[2]): the policy the
 * reference's BPTT / SHAC loops evaluate once per env step (utils/algorithms/BPTT.py:107-115, shac.py:213-222)
 *     action = clip(tanh(W3 tanh(W2 tanh(W1 x + b1) + b2) + b3), lo, hi)      x (n, d <= 32), hidden h in {32, 64}
 * forward and backward as one launch each instead of ~40 library launches per env step — on the 5th-generation tensor
 * cores (tcgen05, accumulators in tensor memory, 3xTF32: every operand split into tf32 hi + lo, every product issued as
 * hi*hi + lo*hi + hi*lo, results at fp32 accuracy; csrc/vf_policy_tc.cuh) for observation widths <= 32 forward and <= 16
 * backward, on the CUDA cores otherwise (csrc/vf_policy.cu; VF_POLICY_NO_TC=1 forces them).
 * The observation may be handed over in two row-major pieces x = [xa (n, da) | xb (n, db)], d = da + db (the obs dict
 * of NavigationEnv is {state (n,13), target (n,3)}; no concatenated copy is made); db = 0, xb = NULL for one piece.
 * Weights: vf_policy_pack converts the six torch.nn.Linear tensors (row-major W1 (h, d), b1, W2 (h, h), b2,
 * W3 (4, h), b3; fp32) into one block of vf_policy_packed_floats(h) floats holding them in the tile layouts of the
 * kernels (k-major and natural, columns interleaved over threads); call it once per weight update, the forward and
 * backward of every env step of the horizon read the block.
 *   vf_policy_bwd: grad_xa (n, da) / grad_xb (n, db) may each be NULL; `partial` is scratch of
 *   vf_policy_partial_floats(n, d, h) floats; grad_params receives
 *   [dW1 (h,d) | db1 (h) | dW2 (h,h) | db2 (h) | dW3 (4,h) | db3 (4)] (overwritten, summed over agents in a fixed
 *   order: bit-reproducible).  Nothing is saved between the two calls: the backward recomputes the activations.
 * ===================================================================================================== */
int vf_policy_packed_floats(int h);
int vf_policy_pack(int d, int h, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                   const float* b3, float* packed, void* stream);
int vf_policy_fwd(int n, int da, int db, int h, const float* xa, const float* xb, const float* packed, float lo,
                  float hi, float* action, void* stream);
int vf_policy_bwd(int n, int da, int db, int h, const float* xa, const float* xb, const float* packed, float lo,
                  float hi, const float* grad_action, float* grad_xa, float* grad_xb, float* partial,
                  float* grad_params, void* stream);
int vf_policy_partial_floats(int n, int d, int h);
const char* vf_policy_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* VISFLY_B200_H */
