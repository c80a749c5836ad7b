/* b2h.h — C-ABI of the B200 batched humanoid rollout library (libb2h.so).
 *
 * The reference (redradman/MujocoPoseLearning) has no FFI of its own: its hot path is six Python calls
 * into the third-party `mujoco` wheel plus the SB3 rollout buffer.  Each entry point below names the
 * reference call site it replaces (paths relative to the reference root):
 *
 *   b2h_create            mujoco.MjModel.from_xml_path + mujoco.MjData      custom_env.py:53-54
 *                         (once per SubprocVecEnv worker, train_sb3.py:203; here once per GPU)
 *   b2h_reset             HumanoidEnv.reset: mj_resetData, noise, 1 mj_step  custom_env.py:97-150
 *   b2h_step              HumanoidEnv.step (ctrl write, mj_step x frame_skip, _get_state,
 *                         _compute_reward, truncation/termination) + the SubprocVecEnv worker's
 *                         auto-reset / terminal_observation                  custom_env.py:152-261,
 *                                                                           reward_functions.py:66-211
 *   b2h_step_vecenv       the same through host buffers in the form SubprocVecEnv.step_wait returns them
 *                         (float64 obs / reward, flags, terminal observations), train_sb3.py:203; page-locked
 *                         buffers are written by the kernel itself.  b2h_step_host: results in the arithmetic dtype
 *   b2h_get_state/set     MjData.qpos/qvel/qacc_warmstart/time field access  custom_env.py:105-117,242-246
 *   b2h_gae               RolloutBuffer.compute_returns_and_advantage (SB3 2.3.2), driven by
 *                         model.learn()                                      train_sb3.py:228
 *   b2h_rollout_collect   collect_rollouts + GAE as one device-resident loop (SB3 2.3.2)  train_sb3.py:228
 *   b2h_mlp_forward / b2h_policy_forward / b2h_policy_sample
 *                         MlpPolicy forward + DiagGaussian sampling during collect_rollouts (SB3 2.3.2)
 *                                                                           train_sb3.py:208-214
 *   b2h_choose_launch_shape, b2h_launch_info, b2h_measure_fp32_peak, b2h_get_counters, b2h_debug_forward
 *                         no reference counterpart: launch policy, measurement and parity hooks
 *
 * Conventions: every function returns 0 on success or a negative B2H_E* code and never throws; the
 * message for the last failure on the calling thread is b2h_last_error().  Pointers named *_dev are
 * caller-owned device pointers (e.g. torch tensor data_ptr()); *_host are host pointers.  All device
 * work is ordered on the cudaStream_t passed as `void* stream` (NULL = legacy default stream); only the
 * *_host / b2h_step_vecenv entry points synchronise that stream.  One handle per GPU, one host thread per handle.
 */
#ifndef B2H_H_
#define B2H_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2H_ABI_VERSION 2

#define B2H_MAX_BODY 32
#define B2H_MAX_JNT 32
#define B2H_MAX_DOF 32
#define B2H_MAX_QPOS 40
#define B2H_MAX_GEOM 32
#define B2H_MAX_PAIR 512
#define B2H_MAX_TENDON 4
#define B2H_MAX_ACT 32

enum { B2H_OK = 0, B2H_EINVAL = -1, B2H_ECUDA = -2, B2H_EUNSUPPORTED = -3, B2H_ENOMEM = -4 };
enum { B2H_F32 = 0, B2H_F64 = 1 };                        /* arithmetic type of the physics state    */
enum { B2H_REWARD_STAND = 0, B2H_REWARD_KNEELING = 1, B2H_REWARD_WALK = 2 };
enum { B2H_OBS_FULL352 = 0, B2H_OBS_QPOS_QVEL = 1 };      /* custom_env.py:242-256 / its 53-col prefix */
enum { B2H_GEOM_PLANE = 0, B2H_GEOM_SPHERE = 2, B2H_GEOM_CAPSULE = 3 };
enum { B2H_JNT_FREE = 0, B2H_JNT_HINGE = 3 };

/* Compiled model constants (the mjModel fields the path reads).  Filled by mjcf.py on the host. */
typedef struct B2HModel {
  int32_t nq, nv, nu, nbody, njnt, ngeom, ntendon, npair;
  double timestep;
  double gravity[3];
  double meaninertia;
  /* bodies */
  int32_t body_parentid[B2H_MAX_BODY];
  int32_t body_jntadr[B2H_MAX_BODY];
  int32_t body_jntnum[B2H_MAX_BODY];
  int32_t body_dofadr[B2H_MAX_BODY];
  int32_t body_dofnum[B2H_MAX_BODY];
  int32_t body_lastdof[B2H_MAX_BODY];      /* last dof on the chain root->body, -1 for the world  */
  double body_pos[B2H_MAX_BODY][3];
  double body_quat[B2H_MAX_BODY][4];
  double body_ipos[B2H_MAX_BODY][3];
  double body_iquat[B2H_MAX_BODY][4];
  double body_inertia[B2H_MAX_BODY][3];      /* principal moments (mjModel.body_inertia)           */
  double body_inertia_full[B2H_MAX_BODY][6]; /* xx yy zz xy xz yz about the COM in body axes       */
  double body_mass[B2H_MAX_BODY];
  double body_subtreemass[B2H_MAX_BODY];
  double body_invweight0[B2H_MAX_BODY][2];
  /* joints */
  int32_t jnt_type[B2H_MAX_JNT];
  int32_t jnt_bodyid[B2H_MAX_JNT];
  int32_t jnt_qposadr[B2H_MAX_JNT];
  int32_t jnt_dofadr[B2H_MAX_JNT];
  int32_t jnt_limited[B2H_MAX_JNT];
  double jnt_pos[B2H_MAX_JNT][3];
  double jnt_axis[B2H_MAX_JNT][3];
  double jnt_range[B2H_MAX_JNT][2];
  double jnt_stiffness[B2H_MAX_JNT];
  double jnt_margin[B2H_MAX_JNT];
  double jnt_solref[B2H_MAX_JNT][2];
  double jnt_solimp[B2H_MAX_JNT][5];
  /* dofs */
  int32_t dof_bodyid[B2H_MAX_DOF];
  int32_t dof_jntid[B2H_MAX_DOF];
  int32_t dof_parentid[B2H_MAX_DOF];
  double dof_armature[B2H_MAX_DOF];
  double dof_damping[B2H_MAX_DOF];
  double dof_invweight0[B2H_MAX_DOF];
  double qpos0[B2H_MAX_QPOS];
  double qpos_spring[B2H_MAX_QPOS];
  /* geoms */
  int32_t geom_type[B2H_MAX_GEOM];
  int32_t geom_bodyid[B2H_MAX_GEOM];
  double geom_size[B2H_MAX_GEOM][3];
  double geom_pos[B2H_MAX_GEOM][3];
  double geom_quat[B2H_MAX_GEOM][4];
  /* static collision candidates with mixed contact parameters */
  int32_t pair_geom1[B2H_MAX_PAIR];
  int32_t pair_geom2[B2H_MAX_PAIR];
  int32_t pair_condim[B2H_MAX_PAIR];
  double pair_friction[B2H_MAX_PAIR][3];
  double pair_solref[B2H_MAX_PAIR][2];
  double pair_solimp[B2H_MAX_PAIR][5];
  double pair_margin[B2H_MAX_PAIR];
  double pair_gap[B2H_MAX_PAIR];
  /* fixed tendons */
  int32_t ten_limited[B2H_MAX_TENDON];
  double ten_J[B2H_MAX_TENDON][B2H_MAX_DOF];
  double ten_qcoef[B2H_MAX_TENDON][B2H_MAX_QPOS];
  double ten_range[B2H_MAX_TENDON][2];
  double ten_solref[B2H_MAX_TENDON][2];
  double ten_solimp[B2H_MAX_TENDON][5];
  double ten_margin[B2H_MAX_TENDON];
  double ten_invweight0[B2H_MAX_TENDON];
  /* motors */
  int32_t actuator_dofid[B2H_MAX_ACT];
  int32_t actuator_ctrllimited[B2H_MAX_ACT];
  double actuator_gear[B2H_MAX_ACT];
  double actuator_ctrlrange[B2H_MAX_ACT][2];
} B2HModel;

/* Environment configuration: the env_config dict of custom_env.py:21-32 plus batch parameters. */
typedef struct B2HConfig {
  int32_t n_envs;          /* environments on this GPU                                             */
  int32_t env_id_offset;   /* global id of local env 0 (keys the reset-noise stream; multi-GPU)     */
  int32_t frame_skip;      /* custom_env.py:28 (default 5; train_sb3.py:199 passes 3)               */
  int32_t reward_type;     /* B2H_REWARD_*                                                         */
  int32_t obs_mode;        /* B2H_OBS_*                                                            */
  int32_t dtype;           /* B2H_F32 / B2H_F64                                                    */
  int32_t max_steps;       /* truncation threshold, custom_env.py:203 (750)                        */
  int32_t device;          /* CUDA device ordinal                                                  */
  double duration;         /* custom_env.py:22 (terminated = time >= duration, :213)               */
  uint64_t seed;           /* reset-noise Philox key                                               */
  double kneeling_params[9]; /* target_height,min_height,max_roll_pitch,com_radius,energy_w,posture_w,com_w,foot_w,alive_w */
  int32_t sensor_terms;    /* 0 (default) = the reference: data.cfrc_ext / data.subtree_linvel are identically zero (the model has
                              no sensor, so MuJoCo never computes them: reward_functions.py:109,121-122,176-177 read zeros).
                              1 = compute what mj_rnePostConstraint (contact part) / mj_subtreeVel would give and feed the rewards. */
  int32_t no_auto_reset;   /* 0 (default) = SubprocVecEnv worker semantics: a finished env is reset inside the step and the step returns
                              the first observation of the new episode.  1 = gymnasium Env.step semantics (custom_env.py:152-230):
                              the env stays in its terminal state (state, observation) until b2h_reset is called for it. */
} B2HConfig;

typedef struct B2HHandle B2HHandle;

int b2h_abi_version(void);
size_t b2h_sizeof_model(void);
size_t b2h_sizeof_config(void);
const char* b2h_last_error(void);

int b2h_create(const B2HModel* model, const B2HConfig* cfg, B2HHandle** out);
void b2h_destroy(B2HHandle* h);
int b2h_obs_dim(const B2HHandle* h);
/* Launch shape the step kernel uses on this device: CTAs, warps (= envs in flight) per CTA, dynamic smem. */
int b2h_launch_info(const B2HHandle* h, int* grid, int* warps_per_cta, size_t* smem_bytes);
/* The shape rule itself (pure host logic, no device needed): env-warps per CTA and dense constraint rows kept in shared
 * memory for n_envs on a device with n_sm SMs and max_smem_bytes of opt-in shared memory per CTA. */
int b2h_choose_launch_shape(int n_envs, int n_sm, size_t max_smem_bytes, int dtype, int* warps_per_cta, int* shared_rows);

/* Re-key the reset-noise stream (VecEnv.seed / Env.reset(seed=...), custom_env.py:99-100) and restart the
 * per-env episode counters, so the following resets replay the same noise for the same seed. */
int b2h_set_seed(B2HHandle* h, uint64_t seed);

/* Reset every env whose mask_dev[i] != 0 (mask_dev == NULL: all).  Writes the first observation of the
 * new episode into obs_dev ([n_envs, obs_dim], float for B2H_F32, double for B2H_F64). */
int b2h_reset(B2HHandle* h, const uint8_t* mask_dev, void* obs_dev, void* stream);

/* Parity hook: explicit reset noise [n_envs, nq+nv] (double, device) consumed by the next reset of each
 * env instead of the Philox stream (custom_env.py:109-117 draws it from the global numpy RNG). */
int b2h_set_reset_noise(B2HHandle* h, const double* noise_dev, void* stream);
/* The noise vector the most recent reset of each env used, [n_envs, nq+nv] double (device -> device). */
int b2h_get_last_reset_noise(B2HHandle* h, double* noise_dev, void* stream);

/* One control step for all envs (custom_env.py:152-230) followed by SubprocVecEnv auto-reset semantics:
 * for done envs obs_dev holds the first observation of the next episode and terminal_obs_dev (may be
 * NULL) the last observation of the finished one.  actions_dev: float [n_envs, nu] (already clipped by
 * the caller as SB3 does; the motor ctrlrange clamp of mj_fwdActuation is applied regardless).
 * reward_dev: float/double [n_envs]; terminated_dev / truncated_dev: uint8 [n_envs]. */
int b2h_step(B2HHandle* h, const float* actions_dev, void* obs_dev, void* reward_dev,
             uint8_t* terminated_dev, uint8_t* truncated_dev, void* terminal_obs_dev, void* stream);

/* Host-buffer form of b2h_step: copies actions H2D, steps, copies results D2H, synchronises.  obs_host is
 * float [n_envs, obs_dim] for B2H_F32 (double for B2H_F64); reward_host likewise; flags uint8.  Page-locked result
 * buffers are written by the kernel directly (as in b2h_step_vecenv); terminal_obs_host then only receives the rows of
 * envs whose episode ended. */
int b2h_step_host(B2HHandle* h, const float* actions_host, void* obs_host, void* reward_host,
                  uint8_t* terminated_host, uint8_t* truncated_host, void* terminal_obs_host, void* stream);
int b2h_reset_host(B2HHandle* h, const uint8_t* mask_host, void* obs_host, void* stream);

/* VecEnv form of the step: what SubprocVecEnv.step_wait hands to SB3 (train_sb3.py:203) - observations and rewards
 * as float64 (observation_space dtype, custom_env.py:80-85) whatever the arithmetic dtype, flags as uint8, terminal
 * observations only for the rows of envs that finished (other rows untouched).  When the host buffers are page-locked
 * (cudaHostAlloc / cudaHostRegister / torch pin_memory) the kernel writes them directly, so results cross PCIe
 * while the remaining envs are still being stepped; pageable buffers go through float64 staging in HBM.  Returns
 * after the stream is synchronised; *n_done (may be NULL) receives the number of envs whose episode ended. */
int b2h_step_vecenv(B2HHandle* h, const float* actions_host, double* obs_host, double* reward_host,
                    uint8_t* terminated_host, uint8_t* truncated_host, double* terminal_obs_host, int* n_done, void* stream);

/* Physics state access, all double on the host side regardless of dtype (tests, checkpoints):
 * qpos [n_envs,nq], qvel [n_envs,nv], warmstart [n_envs,nv], nstep int32 [n_envs] (physics steps since
 * mj_resetData: time = nstep*timestep), step_count int32 [n_envs], total_reward [n_envs].  NULL = skip. */
int b2h_get_state(B2HHandle* h, double* qpos_host, double* qvel_host, double* warmstart_host,
                  int32_t* nstep_host, int32_t* step_count_host, double* total_reward_host);
int b2h_set_state(B2HHandle* h, const double* qpos_host, const double* qvel_host, const double* warmstart_host,
                  const int32_t* nstep_host, const int32_t* step_count_host, const double* total_reward_host);

/* Debug / parity: run mj_forward-equivalent stages on the current state with ctrl = actions_dev (may be
 * NULL) and dump intermediates of env `env` as doubles into out_host; `what` selects the array by name
 * ("xpos","xquat","cinert","cvel","cdof","qM","qfrc_bias","qfrc_passive","qfrc_actuator",
 * "qacc_smooth","qacc","contact_dist","contact_pos","contact_frame","efc_aref","efc_D","efc_force",...).
 * Returns the number of doubles written (>= 0) or a negative error. */
int b2h_debug_forward(B2HHandle* h, const float* actions_dev, int env, const char* what, double* out_host, int max_out);

/* Counters since creation: [0] physics steps, [1] contact overflows (more than the kernel's contact
 * capacity were active; extras dropped), [2] solver iteration cap hits, [3] bad-state resets
 * (mj_checkPos/Vel/Acc equivalents), [4] total Newton iterations, [5] kernels launched, [6] line-search
 * cost evaluations. */
int b2h_get_counters(B2HHandle* h, uint64_t counters_host[8]);

/* Measurement aid: FP32 FMA throughput of `device` in TFLOP/s, measured with a register-resident FFMA kernel
 * (the roofline denominator of this FP32 / issue-bound path; SURVEY 8d asks for "of measured", not nominal). */
int b2h_measure_fp32_peak(int device, double* tflops);

/* GAE reverse scan (SB3 2.3.2 RolloutBuffer.compute_returns_and_advantage).  All [T, E] float arrays,
 * E fastest; last_values [E]; last_dones uint8 [E].  gamma / gae_lambda are Python floats in SB3: they are
 * rounded to float32 the way numpy does (gamma, and the double product gamma*gae_lambda, once each). */
int b2h_gae(const float* rewards_dev, const float* values_dev, const float* episode_starts_dev,
            const float* last_values_dev, const uint8_t* last_dones_dev, double gamma, double gae_lambda,
            int T, int E, float* advantages_dev, float* returns_dev, void* stream);

/* Policy / value MLP forward of collect_rollouts (SB3 2.3.2 MlpPolicy selected at train_sb3.py:209; separate
 * pi / vf trunks of two hidden layers + a linear head, ReLU: main.py:99-105, README.md:45-50) on the tcgen05 tensor
 * cores: y = W3 relu(W2 relu(W1 x + b1) + b2) + b3 for one trunk+head.  x [n_rows, in_dim]; W in nn.Linear layout
 * [out_features, in_features]; in_dim, hidden multiples of 16, hidden <= 256, out_dim <= 32.  precise = 1: tf32 hi/lo
 * split (fp32-faithful), 0: single tf32 pass.  error_flag_dev (int, device) is set to 1 if the tensor pipeline timed out. */
int b2h_mlp_forward(const float* x_dev, const float* w1_dev, const float* b1_dev, const float* w2_dev, const float* b2_dev,
                    const float* w3_dev, const float* b3_dev, float* y_dev, int n_rows, int in_dim, int hidden, int out_dim,
                    int precise, int* error_flag_dev, void* stream);
/* Both trunks of MlpPolicy in one launch (what `policy(obs_tensor)` computes in collect_rollouts): pi_dev / vf_dev
 * are {W1, b1, W2, b2, W3, b3} device pointers; mean_dev [n_rows, act_dim], value_dev [n_rows]. */
int b2h_policy_forward(const float* x_dev, const float* const pi_dev[6], const float* const vf_dev[6], float* mean_dev,
                       float* value_dev, int n_rows, int in_dim, int hidden, int act_dim, int precise, int* error_flag_dev,
                       void* stream);
const char* b2h_mlp_last_error(void);

/* DiagGaussian sampling of collect_rollouts (SB3 distributions.py): actions = mean + exp(log_std) * eps (stored
 * unclipped), clipped = clip(actions, -1, 1) for the env (custom_env.py:88-93 bounds), log_prob summed over dims.
 * eps is Philox4x32-10 keyed by (seed, row_offset + row, step); deterministic != 0 gives actions = mean. */
int b2h_policy_sample(const float* mean_dev, const float* log_std_dev, int n_rows, int act_dim, uint64_t seed, uint64_t step,
                      int row_offset, int deterministic, float* actions_dev, float* clipped_dev, float* log_prob_dev, void* stream);

/* The same forward with the weights prepared once per policy update (the round-2 kernel, mlp_forward_v2_kernel):
 * b2h_policy_pack splits both networks' weights into tf32 hi / lo parts laid out in the tensor cores' core-matrix order
 * (0.6 MB per network); b2h_policy_forward_packed then streams them with the TMA engine (cp.async.bulk.tensor) while the
 * activations flow TMEM -> shared-memory operand ring between the layers.  Call b2h_policy_pack again whenever the
 * weights changed; the biases are read from pi_dev / vf_dev ({W1, b1, W2, b2, W3, b3}) on every forward.  mean_dev or
 * value_dev may be null (only the other network runs: SB3 predict_values).  Any in_dim; hidden a multiple of 32 up to 256. */
typedef struct B2HPolicyPacked B2HPolicyPacked;
int b2h_policy_packed_create(int in_dim, int hidden, int act_dim, B2HPolicyPacked** out);
void b2h_policy_packed_destroy(B2HPolicyPacked* p);
int b2h_policy_pack(B2HPolicyPacked* p, const float* const pi_dev[6], const float* const vf_dev[6], void* stream);
int b2h_policy_forward_packed(B2HPolicyPacked* p, const float* x_dev, const float* const pi_dev[6], const float* const vf_dev[6],
                              float* mean_dev, float* value_dev, int n_rows, int precise, int* error_flag_dev, void* stream);

/* As b2h_policy_sample, with the step counter read from device memory (step = *step_dev + step_offset): a captured
 * CUDA graph of the rollout loop then draws fresh noise on every replay. */
int b2h_policy_sample_dev(const float* mean_dev, const float* log_std_dev, int n_rows, int act_dim, uint64_t seed,
                          const uint64_t* step_dev, uint64_t step_offset, int row_offset, int deterministic, float* actions_dev,
                          float* clipped_dev, float* log_prob_dev, void* stream);

/* OnPolicyAlgorithm.collect_rollouts + RolloutBuffer.compute_returns_and_advantage (SB3 2.3.2, driven by
 * train_sb3.py:228) for n_steps control steps, entirely on the device and without a host round trip per step:
 *   per step t: policy / value forward on obs[t] (tcgen05) -> DiagGaussian sample -> clip -> env step -> record.
 * The kernels write the SB3 buffer layout in place: the step kernel puts the next observation into obs[t + 1] and
 * the reward into rewards[t]; the sampler puts the raw action into actions[t] and its log-probability into
 * log_probs[t]; the value head writes values[t]; the record kernel (fused with the effort sort that follows every
 * step launch) writes episode_starts[t + 1] = done, adds gamma * V(terminal_obs) to the reward where the episode was
 * cut by the step limit only (TimeLimit.truncated), and keeps the episode statistics (raw env rewards, as SB3's
 * Monitor would).  Four kernel launches per control step (five with the time-limit bootstrap).  Then
 * last_values = V(obs[n_steps]) and the GAE scan.  Slot 0 of obs / episode_starts is the caller's carry-over from the
 * previous rollout (copy slot n_steps there, or fill it after b2h_reset).  float32 build (dtype B2H_F32) only.
 * All pointers are device pointers owned by the caller. */
typedef struct B2HRollout {
  int32_t n_steps;
  int32_t hidden, precise, deterministic, row_offset, bootstrap_timeouts;
  uint64_t seed;
  double gamma, gae_lambda;
  float* obs;             /* [n_steps + 1, E, obs_dim]                                    */
  float* actions;         /* [n_steps, E, nu]  raw (unclipped) actions                     */
  float* rewards;         /* [n_steps, E]                                                  */
  float* values;          /* [n_steps, E]                                                  */
  float* log_probs;       /* [n_steps, E]                                                  */
  float* episode_starts;  /* [n_steps + 1, E]                                              */
  float* advantages;      /* [n_steps, E]                                                  */
  float* returns;         /* [n_steps, E]                                                  */
  float* last_values;     /* [E]                                                           */
  float* mean;            /* [E, nu] scratch                                               */
  float* clipped;         /* [E, nu] scratch                                               */
  float* v_term;          /* [E] scratch (bootstrap_timeouts != 0)                         */
  float* ep_return;       /* [E] running return of the episode in flight                   */
  float* ep_len;          /* [E] running length                                            */
  double* stats;          /* [4] sum of finished returns, sum of lengths, episodes, unused */
  uint64_t* step_counter; /* [1] control steps taken so far (sampling noise counter)       */
  int32_t* mlp_error;     /* [1] set by the MLP kernel if its tensor pipeline timed out    */
  const float* pi[6];     /* W1 b1 W2 b2 W3 b3 of the policy network                       */
  const float* vf[6];     /* ... of the value network                                      */
  const float* log_std;   /* [nu]                                                          */
  B2HPolicyPacked* packed; /* b2h_policy_packed_create(obs_dim, hidden, nu): re-packed at the start of every rollout; null: the round-1 kernel */
} B2HRollout;
size_t b2h_sizeof_rollout(void);
int b2h_rollout_collect(B2HHandle* h, const B2HRollout* r, void* stream);

/* ---- PPO update: SB3 2.3.2 PPO.train as model.learn() runs it after every rollout
 * (train_sb3.py:208-231, kwargs config.py:17-32) on hand-written kernels (csrc/b2h_ppo.cu): tcgen05 GEMMs for the forward
 * and backward of both MlpPolicy trunks (fp32-faithful tf32 hi / lo split), the clipped-surrogate / value / entropy loss with
 * per-minibatch advantage normalisation, grad-norm clipping and Adam.
 *
 * The parameters live in ONE flat float vector (so do the gradient and Adam's two moment vectors):
 *   pi W1 [hidden, obs] b1 W2 [hidden, hidden] b2 W3 [act, hidden] b3 | vf W1 b1 W2 b2 W3 [1, hidden] b3 | log_std [act]
 * in nn.Linear layout, every tensor starting at a multiple of 4 floats (b2h_ppo_param_layout returns the 13 offsets and the
 * total length; the padding holds zeros).  The rollout kernels read the same memory, so an update needs no weight copy. */
typedef struct B2HPpoConfig {
  int32_t obs_dim, hidden, act_dim;
  int32_t max_batch;            /* largest minibatch (rows) the workspace is sized for                                  */
  int32_t precise;              /* 1: tf32 hi / lo split, three MMA passes (fp32-faithful, the reference trains in fp32) */
  int32_t normalize_advantage;  /* SB3 default True: (adv - mean) / (std + 1e-8) per minibatch                          */
  float clip_range, ent_coef, vf_coef, max_grad_norm;   /* SB3 defaults 0.2, 0.0, 0.5, 0.5 (config.py:23-24)            */
  float lr, beta1, beta2, adam_eps;                     /* 3e-4 (config.py:18), 0.9, 0.999, 1e-5 (SB3's Adam eps)       */
  int32_t staged_operands;      /* 0 (default): operands kept pre-split in core-matrix order and streamed by TMA; 1: the
                                   first version of the GEMM, producer warps stage plain row-major operands (kept for A/B
                                   measurements and for hidden widths that are not a multiple of 32)                     */
} B2HPpoConfig;
typedef struct B2HPpo B2HPpo;
size_t b2h_sizeof_ppo_config(void);
int64_t b2h_ppo_param_layout(int obs_dim, int hidden, int act_dim, int64_t offsets[13]);
int b2h_ppo_create(const B2HPpoConfig* cfg, B2HPpo** out);
void b2h_ppo_destroy(B2HPpo* h);
const char* b2h_ppo_last_error(void);
/* Gradient of the PPO loss on one minibatch into grad_dev (flat, overwritten).  The rollout buffer is given flattened:
 * obs [n, obs_dim], actions [n, act_dim] (raw, unclipped), old_log_probs / advantages / returns [n]; the minibatch is rows
 * idx_dev[0 .. n_rows) (int64, e.g. a slice of torch.randperm) or, with idx_dev == NULL, rows row_start .. row_start + n_rows.
 * Statistics of the minibatch (b2h_ppo_stats): [0] policy loss, [1] value loss, [2] clip fraction, [3] approx. KL. */
int b2h_ppo_minibatch_grad(B2HPpo* h, const float* obs_dev, const float* actions_dev, const float* old_log_probs_dev,
                           const float* advantages_dev, const float* returns_dev, const int64_t* idx_dev, int64_t row_start, int n_rows,
                           const float* params_dev, float* grad_dev, void* stream);
/* clip_grad_norm_(max_grad_norm) + one Adam step (step = 1, 2, ...: the bias corrections) on the flat vectors.  grad_scale
 * multiplies the gradient first: 1 / world_size after a sum all-reduce over ranks (between the two calls).
 * Statistics [5] receives the gradient norm before clipping. */
int b2h_ppo_apply(B2HPpo* h, float* params_dev, float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t step, float grad_scale,
                  void* stream);
/* PPO.train for one rank: n_epochs passes over perm_dev [n_epochs, n_samples] (each row a permutation of the buffer rows) in
 * minibatches of batch_size, gradient + apply per minibatch, nothing returns to the host in between.  After
 * b2h_ppo_p2p_attach (several ranks) every rank makes this same call and the gradients are summed over the ranks by peer
 * loads inside it (all ranks must use the same n_samples / batch_size / n_epochs). */
int b2h_ppo_train(B2HPpo* h, const float* obs_dev, const float* actions_dev, const float* old_log_probs_dev, const float* advantages_dev,
                  const float* returns_dev, const int64_t* perm_dev, int64_t n_samples, int n_epochs, int batch_size, float* params_dev,
                  float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t* step_inout, void* stream);
/* Gradient all-reduce over NVLink peer memory instead of a library collective (one node, one process per GPU).  Each rank
 * exports the allocation that holds its flat gradient (b2h_ppo_p2p_export fills a 64-byte cudaIpcMemHandle_t), the ranks
 * exchange the handles (e.g. torch.distributed.all_gather_object) and attach the peers' buffers.  Per minibatch the gradient
 * is then computed INTO b2h_ppo_p2p_grad(h) (the copy alternates) and b2h_ppo_apply_p2p runs: flag barrier across the ranks,
 * sum of all ranks' gradients by peer loads in rank order (bit-identical on every rank) fused with the sum of squares,
 * clip + Adam on the mean.  Waits are bounded: a rank that never arrives raises the error flag (b2h_ppo_stats) instead of
 * hanging the GPU. */
int b2h_ppo_p2p_export(B2HPpo* h, void* ipc_handle_out64);
int b2h_ppo_p2p_attach(B2HPpo* h, int rank, int world, const void* ipc_handles);
float* b2h_ppo_p2p_grad(B2HPpo* h);
int b2h_ppo_apply_p2p(B2HPpo* h, float* params_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t step, void* stream);
/* Statistics of the last minibatch (8 doubles) and the error flag (0 = fine; 1 = a tensor / TMA pipeline wait exceeded its bound,
 * 2 = a peer rank never signalled its gradient, 3 = the apply kernel's grid barrier timed out: all waits in the kernels are bounded,
 * a failure raises this flag instead of hanging the GPU); synchronises the stream. */
int b2h_ppo_stats(B2HPpo* h, double stats_host[8], int* error_host, void* stream);
const double* b2h_ppo_stats_dev(const B2HPpo* h);
const int* b2h_ppo_error_dev(const B2HPpo* h);
/* The GEMM of the update on its own (tests, measurement): C[m, n] (+)= A . B^T (+ bias, ReLU, . (mask > 0)) with fp32 operands
 * on the tcgen05 tensor cores.  a_kstrided = 0: A is [m, k] row-major (lda >= k); 1: A is [k, m] row-major (lda >= m); B
 * likewise with n.  transpose_c: the result is written as C[n, m] (ldc >= m).  split_k = 1: one CTA per tile; > 1 (or 0 =
 * fill the SMs once): the contraction is split over CTAs and partial tiles are ADDED to C with red.global.add, as they are
 * when accumulate != 0 — the caller zeroes C.  bias_dev [n], mask_dev [m, ldmask] may be NULL. */
int b2h_gemm(const float* a_dev, int lda, int a_kstrided, const float* b_dev, int ldb, int b_kstrided, float* c_dev, int ldc, int transpose_c,
             const float* bias_dev, const float* mask_dev, int ldmask, int m, int n, int k, int relu, int precise, int split_k, int accumulate,
             int* error_flag_dev, void* stream);

/* The TMA-fed GEMM of the update on its own (tests, measurement): the plain row-major operands are first packed into the
 * pre-split core-matrix format the update keeps its activations in (temporary device buffers; the call synchronises).
 * a_mn = 0: A is [m, k] (contraction over its columns, K-major use); 1: A is [k, m] (contraction over its rows, MN-major
 * use); B likewise with n.  split_k = 1 stores C; 0 (fill the SMs) or > 1 adds partial tiles into C (the caller zeroes it). */
int b2h_gemm_tma(const float* a_dev, int a_mn, const float* b_dev, int b_mn, float* c_dev, int ldc, int transpose_c, const float* bias_dev,
                 int m, int n, int k, int precise, int split_k, int* error_flag_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2H_H_ */
