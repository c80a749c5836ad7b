// b2h_api.cu — kernels + C-ABI (include/b2h.h) of the batched humanoid rollout library, sm_100a.
//
// Launch shape: persistent grid, one CTA per SM, one warp per environment in flight; every warp owns a
// Scratch<T> slice of dynamic shared memory and pulls environment ids from a global work counter (per-env cost
// varies with contact count and Newton iterations).  Per-env state is [n_envs, dim] row-major so a warp reads
// and writes one contiguous record per field.
#include <cuda_runtime.h>
#include <stdio.h>
#include <math.h>
#include <stdlib.h>

#include <mutex>
#include <string>
#include <vector>

#include "b2h_debug.h"

using namespace b2h;

// ------------------------------------------------------------------------------------------------ kernels
// Up to 16 env-warps of fp32 scratch (8 of fp64) share the 227 KB of shared memory of one SM; 512 threads leave
// 128 registers per thread (fp64: 256 threads, 255 registers)
extern __shared__ __align__(16) unsigned char b2h_smem[];
#ifndef B2H_MAX_THREADS
#define B2H_MAX_THREADS 512
#endif
template <typename T> constexpr int max_threads() { return sizeof(T) == 8 ? B2H_MAX_THREADS / 2 : B2H_MAX_THREADS; }
// The model tables (12 KB in fp32) lead the CTA's shared memory, one copy for all env-warps; the per-warp scratch
// slices follow.  Read through the L1 instead, they compete with the row spill and the state traffic for the ~24 KB
// of L1 this launch leaves (measured: +2.1 % at 4096 envs, +2.2 % at 16384 with the tables in shared memory, although
// 5-7 constraint rows per env move from shared memory to the spill area to make room).
template <typename T> __host__ __device__ constexpr size_t model_smem_bytes() { return (sizeof(DevModel<T>) + 15) / 16 * 16; }
template <typename T> __device__ __forceinline__ const DevModel<T>* stage_model(const DevModel<T>* g) {
  const uint4* src = reinterpret_cast<const uint4*>(g);
  uint4* dst = reinterpret_cast<uint4*>(b2h_smem);
  for (int i = threadIdx.x; i < (int)(model_smem_bytes<T>() / 16); i += blockDim.x) dst[i] = src[i];
  __syncthreads();
  return reinterpret_cast<const DevModel<T>*>(b2h_smem);
}
template <typename T> __device__ __forceinline__ Scratch<T>& my_scratch(const DevModel<T>* model) {
  return *reinterpret_cast<Scratch<T>*>(b2h_smem + model_smem_bytes<T>() +
                                        (threadIdx.x >> 5) * (sizeof(Scratch<T>) - (size_t)(NROW_S - model->nrow_s) * LD * sizeof(T)));
}

#ifdef B2H_STAGE_CLOCKS
__device__ unsigned long long g_stage_clk[48];
__device__ unsigned long long g_cta_exit[1024];  // globaltimer at which each CTA of the last step launch left the claim loop
#endif
__device__ __forceinline__ void flush_counters(const Counters& c, unsigned long long* g) {
  if (lane_id() == 0) {
#ifdef B2H_STAGE_CLOCKS
    for (int i = 0; i < 48; i++) atomicAdd(&g_stage_clk[i], (unsigned long long)c.clk[i]);
#endif
    if (c.physics_steps) atomicAdd(g + 0, (unsigned long long)c.physics_steps);
    if (c.contact_overflow) atomicAdd(g + 1, (unsigned long long)c.contact_overflow);
    if (c.iter_cap) atomicAdd(g + 2, (unsigned long long)c.iter_cap);
    if (c.bad_state) atomicAdd(g + 3, (unsigned long long)c.bad_state);
    if (c.newton_iter) atomicAdd(g + 4, (unsigned long long)c.newton_iter);
    if (c.ls_eval) atomicAdd(g + 6, (unsigned long long)c.ls_eval);
  }
}

template <typename T, int OUT>
__global__ void __launch_bounds__(max_threads<T>(), 1)
step_kernel(const DevModel<T>* __restrict__ gmodel, EnvParams P, EnvIO<T> io, int n_envs, unsigned long long* counters,
            int* work, T* spill, const int* __restrict__ perm) {
  const DevModel<T>* model = stage_model<T>(gmodel);
  Scratch<T>& S = my_scratch<T>(model);
  T* Jspill = spill + (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * ((NROW - model->nrow_s) * LD);
  Counters cnt = {};
  __shared__ int s_base;
  const int nwarps = blockDim.x >> 5;
  for (;;) {  // the CTA claims one env per warp at a time and steps them in lockstep (see env_step)
    B2H_CLK(tk);
    __syncthreads();
    if (threadIdx.x == 0) s_base = atomicAdd(work, nwarps);
    __syncthreads();
    B2H_CLK_ADD(10, tk);
    const int base = s_base;
    if (base >= n_envs) break;
    // lockstep groups are taken from the effort-sorted order (order_kernel): warps that wait for each other at the
    // sub-step barriers then carry similar solver work, and the costliest groups start first
    const int slot = base + (threadIdx.x >> 5);
    const bool active = slot < n_envs;
    const int env = active ? (perm ? perm[slot] : slot) : 0;
    // row slots of the group: what the neediest of its envs used in its previous control step (plus a margin); an env
    // that outgrows them falls back on its own (mj_step)
    const int rows = active && io.work ? (int)((unsigned)io.work[env] >> B2H_EFFORT_BITS) : 0;
    const int ns = __syncthreads_or(rows > 60) ? 3 : (__syncthreads_or(rows > 28) ? 2 : 1);
#ifdef B2H_EXP_NEWTON_BARRIER
    const int group_threads = 32 * min(nwarps, n_envs - base);
#else
    const int group_threads = 0;
#endif
    env_step<T, OUT>(*model, S, Jspill, cnt, P, io, env, active, ns, group_threads);
  }
#ifdef B2H_STAGE_CLOCKS
  if (threadIdx.x == 0) {  // when this CTA ran out of work (tail imbalance of the launch)
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_cta_exit[blockIdx.x] = t;
  }
#endif
  flush_counters(cnt, counters);
}

// Counting sort of the envs by the solver effort of their previous control step, costliest first (one CTA).
// Which warp steps which env never changes a result; the order inside a bucket is left to the atomics.
// It also re-arms the claim counter of the next step launch (no memset node on the stream).
constexpr int ORDER_BUCKETS = 256;
__device__ __forceinline__ int order_bucket(int packed, int mode = 0) {   // EnvIO::work: effort | joint limits | dense rows
  const int effort = (int)((unsigned)packed & 16383u), nlim = (int)(((unsigned)packed >> 14) & 63u), rows = (int)((unsigned)packed >> B2H_EFFORT_BITS);
  // Sort key of the lockstep schedule.  Round 1 sorted by the solver effort (Newton iterations x row weight) of the env's last control
  // step; its iteration count is nearly unpredictable (rho = 0.3 from step to step) and scrambles what IS persistent: the contact
  // configuration.  Sorting by the dense row count (ties: active joint limits) makes the groups homogeneous in per-iteration and
  // pre-solver cost: +2.2 % at 4096 envs, +1.7 % at 16384 against the effort key (profiles/r02_schedule_key_ab.txt; no sort: -2.7 %).
  // B2H_ORDER_KEY: 0 effort; 1 rows; 2 effort + rows; 5 rows + limits; 6 (default) rows, ties by limits; 7 rows, ties by effort
  int key;
  switch (mode) {
    case 0: key = effort >> 2; break;
    case 2: key = (effort >> 2) + rows; break;
    case 5: key = 2 * (rows + nlim); break;
    case 6: key = 4 * rows + min(nlim, 3); break;
    case 7: key = 3 * rows + (effort >> 5); break;
    default: key = 2 * rows; break;
  }
  return min(key, ORDER_BUCKETS - 1);
}
// What collect_rollouts does with the result of one env.step (SB3 2.3.2 on_policy_algorithm.py), per env, after the step
// kernel: episode_starts of the next slot, the time-limit bootstrap of the stored reward, the episode statistics.
struct RecordArgs {
  float* rewards_t;             // [E] slot t of the reward buffer (the step kernel wrote the env reward there)
  float* episode_starts_next;   // [E] slot t + 1
  const uint8_t *terminated, *truncated;
  const float* v_term;          // [E] V(terminal_obs), or null (no step-limit truncation possible)
  float gamma;
  float *ep_return, *ep_len;    // [E]
  double* stats;                // sum of finished returns, sum of finished lengths, finished episodes
  unsigned long long* step_counter;
};
__global__ void __launch_bounds__(1024, 1) post_step_kernel(const int* __restrict__ effort, int n, int* __restrict__ perm, int* work,
                                                            int do_sort, int do_record, RecordArgs rec, int key_mode) {
  __shared__ int hist[ORDER_BUCKETS], start[ORDER_BUCKETS];
  __shared__ double red[3][32];
  if (do_sort) {
    if (threadIdx.x == 0) *work = 0;
    for (int i = threadIdx.x; i < ORDER_BUCKETS; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&hist[order_bucket(effort[i], key_mode)], 1);
    __syncthreads();
    if (threadIdx.x < 32) {  // exclusive prefix over the buckets in descending order: 8 buckets per lane + one warp scan
      constexpr int PER = ORDER_BUCKETS / 32;
      const int lane = threadIdx.x, top = ORDER_BUCKETS - 1 - lane * PER;
      int loc[PER], sum = 0;
#pragma unroll
      for (int k = 0; k < PER; k++) { loc[k] = sum; sum += hist[top - k]; }
      int incl = sum;
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
#pragma unroll
      for (int k = 0; k < PER; k++) start[top - k] = incl - sum + loc[k];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) perm[atomicAdd(&start[order_bucket(effort[i], key_mode)], 1)] = i;
  }
  if (do_record) {
    double sr = 0, sl = 0, sn = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float raw = rec.rewards_t[i];
      const bool te = rec.terminated[i] != 0, tr = rec.truncated[i] != 0, done = te || tr;
      if (rec.v_term && tr && !te) rec.rewards_t[i] = raw + rec.gamma * rec.v_term[i];   // TimeLimit.truncated: bootstrap
      rec.episode_starts_next[i] = done ? 1.0f : 0.0f;
      const float er = rec.ep_return[i] + raw, el = rec.ep_len[i] + 1.0f;   // statistics on the raw env reward
      if (done) { sr += er; sl += el; sn += 1; }
      rec.ep_return[i] = done ? 0.0f : er;
      rec.ep_len[i] = done ? 0.0f : el;
    }
    for (int o = 16; o > 0; o >>= 1) {
      sr += __shfl_xor_sync(0xffffffffu, sr, o); sl += __shfl_xor_sync(0xffffffffu, sl, o); sn += __shfl_xor_sync(0xffffffffu, sn, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sr; red[1][threadIdx.x >> 5] = sl; red[2][threadIdx.x >> 5] = sn; }
    __syncthreads();
    if (threadIdx.x < 3) {
      double t = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[threadIdx.x][w];
      rec.stats[threadIdx.x] += t;
    }
    if (threadIdx.x == 0) *rec.step_counter += 1ull;
  }
}

template <typename T>
__global__ void __launch_bounds__(max_threads<T>(), 1)
reset_kernel(const DevModel<T>* __restrict__ gmodel, EnvParams P, EnvIO<T> io, int n_envs, const uint8_t* mask,
             unsigned long long* counters, int* work, T* spill) {
  const DevModel<T>* model = stage_model<T>(gmodel);
  Scratch<T>& S = my_scratch<T>(model);
  T* Jspill = spill + (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * ((NROW - model->nrow_s) * LD);
  Counters cnt = {};
  for (;;) {
    int env = 0;
    if (lane_id() == 0) env = atomicAdd(work, 1);
    env = __shfl_sync(0xffffffffu, env, 0);
    if (env >= n_envs) break;
    if (mask && !mask[env]) continue;
    env_reset_only<T>(*model, S, Jspill, cnt, P, io, env);
  }
  flush_counters(cnt, counters);
}

template <typename T>
__global__ void __launch_bounds__(32, 1)
debug_kernel(const DevModel<T>* model, EnvIO<T> io, int env, DebugDump<T>* out, T* Jspill, int ext) {
  model = stage_model<T>(model);
  Scratch<T>& S = *reinterpret_cast<Scratch<T>*>(b2h_smem + model_smem_bytes<T>());
  const int lane = lane_id(), nq = model->nq, nv = model->nv, nu = model->nu;
  Counters cnt = {};
  EnvState<T> st;
  st.qp = lane < nq ? io.qpos[(size_t)env * nq + lane] : T(0);
  st.qv = lane < nv ? io.qvel[(size_t)env * nv + lane] : T(0);
  st.warm = lane < nv ? io.warm[(size_t)env * nv + lane] : T(0);
  st.nstep = io.nstep[env];
  st.qfrc_act = 0;
  int a = lane < nv ? model->dof_act[lane] : -1;
  st.ctrl = (a >= 0 && io.actions) ? T(io.actions[(size_t)env * nu + a]) : T(0);
  T qacc;
  StepStats stats;
  if (ext) physics_step<T, true, NSLOT, true>(*model, S, Jspill, st, cnt, false, &stats, &qacc, out);
  else physics_step<T, true>(*model, S, Jspill, st, cnt, false, &stats, &qacc, out);
  __syncwarp();
  if (lane == 0) out->stats = stats;
}

// SB3 RolloutBuffer.compute_returns_and_advantage: one thread per env, reverse scan over T; [T, E] arrays, E fastest.
// Explicit _rn intrinsics keep the rounding sequence of the numpy float32 expression (no FMA contraction).
__global__ void gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                           const float* __restrict__ episode_starts, const float* __restrict__ last_values,
                           const uint8_t* __restrict__ dones, const float* __restrict__ dones_f, float gamma, float gl, int T, int E,
                           float* __restrict__ adv, float* __restrict__ ret) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  float last_gae = 0.0f;
  float nnt = 1.0f - (dones ? (float)dones[e] : dones_f[e]), nv = last_values[e];   // last dones as bytes or as 0 / 1 floats
  for (int t = T - 1; t >= 0; t--) {
    size_t i = (size_t)t * E + e;
    float v = values[i];
    float delta = __fsub_rn(__fadd_rn(rewards[i], __fmul_rn(__fmul_rn(gamma, nv), nnt)), v);
    last_gae = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, nnt), last_gae));
    adv[i] = last_gae;
    ret[i] = __fadd_rn(last_gae, v);
    nnt = 1.0f - episode_starts[i];
    nv = v;
  }
}

// FP32 FMA peak of the device, measured: the roofline denominator bench.py quotes for this issue/FP32-bound path
// (MEASURED_PEAKS.json only carries HBM and tensor figures).  16 independent FFMA chains per thread, 8 CTAs of 256 threads per SM.
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float x, float y) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], x, y);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ------------------------------------------------------------------------------------------------ host side
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(B2H_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

struct B2HHandle {
  B2HConfig cfg;
  B2HModel model;
  int nq, nv, nu, nbody, obs_dim, esz;  // esz = sizeof(T)
  void* dmodel = nullptr;               // DevModel<T> on the device
  void *qpos = nullptr, *qvel = nullptr, *warm = nullptr, *total_reward = nullptr;
  int *nstep = nullptr, *step_count = nullptr, *episode = nullptr;
  double* reset_noise = nullptr;
  uint8_t* noise_injected = nullptr;
  unsigned long long* counters = nullptr;
  int* work = nullptr;
  int *effort = nullptr, *perm = nullptr;  // per-env solver effort of the last control step, effort-sorted env order
  int schedule = 1;                        // 1: lockstep groups follow the effort-sorted order
  int order_key = 6;                       // sort key of that order: dense rows, ties by joint limits (tuning knob B2H_ORDER_KEY, see order_bucket)
  bool perm_valid = false;
  bool work_armed = false;                 // the claim counter was zeroed by the sort that followed the previous step launch
  cudaStream_t armed_stream = nullptr;     // ... on this stream: a step on another stream zeroes the counter itself
  cudaEvent_t step_done = nullptr;         // recorded after the step kernel (b2h_step_vecenv waits on it, not on the sort)
  void* dump = nullptr;
  void* spill = nullptr;  // per-warp dense rows beyond NROW_S
  // staging for the *_host entry points
  float* actions_stage = nullptr;
  void *obs_stage = nullptr, *tobs_stage = nullptr, *rew_stage = nullptr;
  double *obs64_stage = nullptr, *tobs64_stage = nullptr, *rew64_stage = nullptr;  // b2h_step_vecenv with pageable buffers
  uint8_t *term_stage = nullptr, *trunc_stage = nullptr, *mask_stage = nullptr;
  int grid = 0, warps = 0;
  size_t smem = 0;
  unsigned long long launches = 0;
  EnvParams P;
};

struct Out64 { double *obs = nullptr, *reward = nullptr, *tobs = nullptr; };

template <typename T>
static EnvIO<T> make_io(B2HHandle* h, const float* actions, void* obs, void* reward, uint8_t* term, uint8_t* trunc, void* tobs,
                        Out64 o64 = Out64()) {
  EnvIO<T> io;
  io.qpos = (T*)h->qpos; io.qvel = (T*)h->qvel; io.warm = (T*)h->warm; io.nstep = h->nstep; io.step_count = h->step_count;
  io.episode = h->episode; io.total_reward = (T*)h->total_reward; io.reset_noise = h->reset_noise;
  io.noise_injected = h->noise_injected; io.actions = actions; io.obs = (T*)obs; io.reward = (T*)reward;
  io.terminal_obs = (T*)tobs; io.terminated = term; io.truncated = trunc; io.obs_dim = h->obs_dim;
  io.work = h->effort;
  io.obs64 = o64.obs; io.reward64 = o64.reward; io.terminal_obs64 = o64.tobs;
  return io;
}

// Launch shape: (env-warps per CTA, dense rows kept in shared memory).  More warps hide more latency but leave
// fewer shared rows (the rest spill to L1/L2-backed global memory) and make the lockstep group larger; what pays
// depends on how many rounds of groups each SM runs.  Measured on B200 (fp32): 16 warps x 32 rows steps a group
// 4.8 % slower than 14 warps x 48 rows, i.e. 8.8 % more envs per second once the SMs stay full.  The row counts of
// the two shapes are upper bounds: each shape keeps as many rows as fit beside the model tables (fp32 on B200:
// 14 x 44 and 16 x 26).
template <typename T>
static void choose_shape(int n_envs, int nsm, size_t max_smem, int* warps_out, int* nrow_s_out) {
  const int maxw = max_threads<T>() / 32;
  struct Shape { int warps, nrow_s; double group_time; } shapes[2] = {{maxw, 32, 1.048}, {maxw - maxw / 8, NROW_S, 1.0}};
  int warps = 0, nrow_s = NROW_S;
  double best = 0;
  if (max_smem > model_smem_bytes<T>()) max_smem -= model_smem_bytes<T>(); else max_smem = 0;
  for (Shape& sh : shapes) {
    // as many shared rows as the shape's warp count leaves room for beside the model tables
    while (sh.nrow_s > NROW_S_MIN && (size_t)sh.warps * scratch_bytes<T>(sh.nrow_s) > max_smem) sh.nrow_s--;
    if (sh.nrow_s > NROW_S || sh.nrow_s < NROW_S_MIN || (size_t)sh.warps * scratch_bytes<T>(sh.nrow_s) > max_smem) continue;
    double groups = ((double)n_envs + sh.warps - 1) / sh.warps / nsm;       // per SM
    double rounds = groups <= 3.0 ? ceil(groups - 1e-9) : groups;           // few groups: whole rounds count
    double t = rounds * sh.group_time;
    if (!warps || t < best) { warps = sh.warps; nrow_s = sh.nrow_s; best = t; }
  }
  if (warps && n_envs < nsm * warps) {  // fewer envs than one full round: spread them over all SMs in smaller groups
    warps = (n_envs + nsm - 1) / nsm;
    nrow_s = NROW_S;
    while (nrow_s > NROW_S_MIN && (size_t)warps * scratch_bytes<T>(nrow_s) > max_smem) nrow_s--;
  }
  *warps_out = warps; *nrow_s_out = nrow_s;
}

template <typename T>
static int create_typed(B2HHandle* h) {
  std::vector<unsigned char> buf(sizeof(DevModel<T>));
  DevModel<T>* dm = reinterpret_cast<DevModel<T>*>(buf.data());
  std::string err = build_dev_model<T>(h->model, *dm);
  if (!err.empty()) return fail(B2H_EUNSUPPORTED, err);
  CU(cudaMalloc(&h->dmodel, sizeof(DevModel<T>)));
  CU(cudaMemcpy(h->dmodel, dm, sizeof(DevModel<T>), cudaMemcpyHostToDevice));
  CU(cudaMalloc(&h->dump, sizeof(DebugDump<T>)));
  int dev = h->cfg.device, nsm = 0, max_smem = 0;
  CU(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  CU(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  int warps = 0, nrow_s = NROW_S;
  choose_shape<T>(h->cfg.n_envs, nsm, (size_t)max_smem, &warps, &nrow_s);
  if (!warps) return fail(B2H_EUNSUPPORTED, "per-env scratch does not fit in shared memory");
  const int maxw = max_threads<T>() / 32;
  int ctas_per_sm = 1;
  if (const char* w = getenv("B2H_WARPS_PER_CTA")) {  // tuning knobs: lockstep group size (several CTAs per SM), shared rows
    int req = atoi(w);
    if (req >= 1 && req <= maxw) { warps = req; }
  }
  if (const char* r = getenv("B2H_NROW_SHARED")) { int req = atoi(r); if (req >= NROW_S_MIN && req <= NROW_S) nrow_s = req; }
  const size_t msm = model_smem_bytes<T>();
  while (warps > 1 && (size_t)warps * scratch_bytes<T>(nrow_s) + msm > (size_t)max_smem) warps--;
  if (getenv("B2H_WARPS_PER_CTA")) {
    ctas_per_sm = (int)((size_t)max_smem / ((size_t)warps * scratch_bytes<T>(nrow_s) + msm));
    if (ctas_per_sm * warps > maxw) ctas_per_sm = maxw / warps;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (const char* c = getenv("B2H_CTAS_PER_SM")) { int req = atoi(c); if (req >= 1 && req < ctas_per_sm) ctas_per_sm = req; }
  }
  dm->nrow_s = nrow_s;
  CU(cudaMemcpy(h->dmodel, dm, sizeof(DevModel<T>), cudaMemcpyHostToDevice));
  h->warps = warps;
  h->smem = (size_t)warps * scratch_bytes<T>(nrow_s) + msm;
  h->grid = nsm * ctas_per_sm;
  CU(cudaMalloc(&h->spill, (size_t)(h->grid * warps + 1) * (NROW - nrow_s) * LD * sizeof(T)));
  CU(cudaFuncSetAttribute(step_kernel<T, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
  CU(cudaFuncSetAttribute(step_kernel<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
  CU(cudaFuncSetAttribute(reset_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
  CU(cudaFuncSetAttribute(debug_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(Scratch<T>) + msm)));
  return B2H_OK;
}

extern "C" {

int b2h_abi_version(void) { return B2H_ABI_VERSION; }
size_t b2h_sizeof_model(void) { return sizeof(B2HModel); }
size_t b2h_sizeof_config(void) { return sizeof(B2HConfig); }
const char* b2h_last_error(void) { return g_err.c_str(); }

void b2h_destroy(B2HHandle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  void* ptrs[] = {h->dmodel, h->qpos, h->qvel, h->warm, h->total_reward, h->nstep, h->step_count, h->episode,
                  h->reset_noise, h->noise_injected, h->counters, h->work, h->effort, h->perm, h->dump, h->spill, h->actions_stage, h->obs_stage,
                  h->tobs_stage, h->rew_stage, h->term_stage, h->trunc_stage, h->mask_stage, h->obs64_stage, h->tobs64_stage,
                  h->rew64_stage};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (h->step_done) cudaEventDestroy(h->step_done);
  delete h;
}

int b2h_create(const B2HModel* model, const B2HConfig* cfg, B2HHandle** out) {
  if (!model || !cfg || !out) return fail(B2H_EINVAL, "null argument");
  if (cfg->n_envs <= 0 || cfg->frame_skip <= 0) return fail(B2H_EINVAL, "n_envs and frame_skip must be positive");
  if (cfg->reward_type < 0 || cfg->reward_type > B2H_REWARD_WALK) return fail(B2H_EINVAL, "Unknown reward type");
  if (cfg->obs_mode != B2H_OBS_FULL352 && cfg->obs_mode != B2H_OBS_QPOS_QVEL) return fail(B2H_EINVAL, "unknown obs_mode");
  if (cfg->dtype != B2H_F32 && cfg->dtype != B2H_F64) return fail(B2H_EINVAL, "unknown dtype");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(B2H_ECUDA, "no CUDA device: this library has no CPU path");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(B2H_EINVAL, "bad device ordinal");
  CU(cudaSetDevice(cfg->device));
  B2HHandle* h = new B2HHandle();
  h->cfg = *cfg; h->model = *model;
  h->nq = model->nq; h->nv = model->nv; h->nu = model->nu; h->nbody = model->nbody;
  h->obs_dim = cfg->obs_mode == B2H_OBS_QPOS_QVEL ? model->nq - 2 + model->nv : model->nq - 2 + model->nv + 16 * model->nbody + model->nv;
  h->esz = cfg->dtype == B2H_F64 ? 8 : 4;
  h->P.frame_skip = cfg->frame_skip; h->P.reward_type = cfg->reward_type; h->P.obs_mode = cfg->obs_mode;
  h->P.max_steps = cfg->max_steps; h->P.duration = cfg->duration; h->P.timestep = model->timestep;
  for (int k = 0; k < 9; k++) h->P.kneel[k] = cfg->kneeling_params[k];
  h->P.seed = cfg->seed; h->P.env_id_offset = cfg->env_id_offset;
  h->P.sync_mode = 2;
  h->P.sensor_terms = cfg->sensor_terms != 0;
  h->P.auto_reset = cfg->no_auto_reset == 0;
  if (const char* sm = getenv("B2H_SYNC_MODE")) h->P.sync_mode = atoi(sm);  // tuning knob, see env_step
  int rc = cfg->dtype == B2H_F64 ? create_typed<double>(h) : create_typed<float>(h);
  if (rc != B2H_OK) { b2h_destroy(h); return rc; }
  size_t E = (size_t)cfg->n_envs, esz = (size_t)h->esz;
#define ALLOC(ptr, bytes) do { cudaError_t e_ = cudaMalloc((void**)&(ptr), (bytes)); if (e_ != cudaSuccess) { b2h_destroy(h); return fail(B2H_ENOMEM, cudaGetErrorString(e_)); } cudaMemset((ptr), 0, (bytes)); } while (0)
  ALLOC(h->qpos, E * h->nq * esz); ALLOC(h->qvel, E * h->nv * esz); ALLOC(h->warm, E * h->nv * esz);
  ALLOC(h->total_reward, E * esz); ALLOC(h->nstep, E * 4); ALLOC(h->step_count, E * 4); ALLOC(h->episode, E * 4);
  ALLOC(h->reset_noise, E * (h->nq + h->nv) * 8); ALLOC(h->noise_injected, E);
  ALLOC(h->counters, 8 * 8); ALLOC(h->work, 8); /* [0] step claim counter, [1] reset claim counter */ ALLOC(h->effort, E * 4); ALLOC(h->perm, E * 4);
  if (const char* sc = getenv("B2H_SCHEDULE")) h->schedule = atoi(sc);  // tuning knob: 0 = env-id order
  if (const char* ok = getenv("B2H_ORDER_KEY")) h->order_key = atoi(ok);
  ALLOC(h->actions_stage, E * h->nu * 4); ALLOC(h->obs_stage, E * h->obs_dim * esz); ALLOC(h->tobs_stage, E * h->obs_dim * esz);
  ALLOC(h->rew_stage, E * esz); ALLOC(h->term_stage, E); ALLOC(h->trunc_stage, E); ALLOC(h->mask_stage, E);
#undef ALLOC
  CU(cudaEventCreateWithFlags(&h->step_done, cudaEventDisableTiming));
  CU(cudaDeviceSynchronize());
  *out = h;
  return B2H_OK;
}

int b2h_obs_dim(const B2HHandle* h) { return h ? h->obs_dim : B2H_EINVAL; }

int b2h_set_seed(B2HHandle* h, uint64_t seed) {
  if (!h) return fail(B2H_EINVAL, "null handle");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaDeviceSynchronize());
  h->cfg.seed = seed; h->P.seed = seed;
  CU(cudaMemset(h->episode, 0, (size_t)h->cfg.n_envs * 4));
  return B2H_OK;
}

int b2h_choose_launch_shape(int n_envs, int n_sm, size_t max_smem_bytes, int dtype, int* warps_per_cta, int* shared_rows) {
  if (n_envs <= 0 || n_sm <= 0 || !warps_per_cta || !shared_rows) return fail(B2H_EINVAL, "bad argument");
  if (dtype == B2H_F64) choose_shape<double>(n_envs, n_sm, max_smem_bytes, warps_per_cta, shared_rows);
  else choose_shape<float>(n_envs, n_sm, max_smem_bytes, warps_per_cta, shared_rows);
  return *warps_per_cta > 0 ? B2H_OK : fail(B2H_EUNSUPPORTED, "per-env scratch does not fit in shared memory");
}

int b2h_launch_info(const B2HHandle* h, int* grid, int* warps_per_cta, size_t* smem_bytes) {
  if (!h) return fail(B2H_EINVAL, "null handle");
  if (grid) *grid = h->grid;
  if (warps_per_cta) *warps_per_cta = h->warps;
  if (smem_bytes) *smem_bytes = h->smem;
  return B2H_OK;
}

int b2h_reset(B2HHandle* h, const uint8_t* mask_dev, void* obs_dev, void* stream) {
  if (!h) return fail(B2H_EINVAL, "null handle");
  CU(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  CU(cudaMemsetAsync(h->work + 1, 0, 4, s));   // the reset kernel's own claim counter: a captured step graph stays valid across resets
  if (h->cfg.dtype == B2H_F64)
    reset_kernel<double><<<h->grid, h->warps * 32, h->smem, s>>>((const DevModel<double>*)h->dmodel, h->P,
        make_io<double>(h, nullptr, obs_dev, nullptr, nullptr, nullptr, nullptr), h->cfg.n_envs, mask_dev, h->counters, h->work + 1, (double*)h->spill);
  else
    reset_kernel<float><<<h->grid, h->warps * 32, h->smem, s>>>((const DevModel<float>*)h->dmodel, h->P,
        make_io<float>(h, nullptr, obs_dev, nullptr, nullptr, nullptr, nullptr), h->cfg.n_envs, mask_dev, h->counters, h->work + 1, (float*)h->spill);
  CU(cudaGetLastError());
  h->launches++;
  return B2H_OK;
}

// A control step = the step kernel, then (on the same stream) whatever has to see its results before the next step:
// the effort sort that orders the next launch's lockstep groups and re-arms its claim counter, and -- for the
// device-resident rollout -- the record of collect_rollouts.  One small single-CTA kernel does both.
static int launch_step_kernel(B2HHandle* h, const float* actions_dev, void* obs_dev, void* reward_dev, uint8_t* terminated_dev,
                              uint8_t* truncated_dev, void* terminal_obs_dev, Out64 o64, cudaStream_t s) {
  const bool sched = h->schedule && h->P.sync_mode == 2 && h->cfg.n_envs > h->warps;
  // the last sort re-armed the claim counter -- in stream order, so only a launch on that same stream may rely on it
  if (!sched || !h->work_armed || h->armed_stream != s) CU(cudaMemsetAsync(h->work, 0, 4, s));
  const int* perm = sched && h->perm_valid ? h->perm : nullptr;   // order of the previous step's efforts (first step: env-id order)
  // one instantiation per result kind: the arithmetic-type arrays (device rollouts, b2h_step_host) or the float64 VecEnv ones
#define B2H_LAUNCH_STEP(T, OUT) step_kernel<T, OUT><<<h->grid, h->warps * 32, h->smem, s>>>((const DevModel<T>*)h->dmodel, h->P, \
      make_io<T>(h, actions_dev, obs_dev, reward_dev, terminated_dev, truncated_dev, terminal_obs_dev, o64), h->cfg.n_envs, \
      h->counters, h->work, (T*)h->spill, perm)
  const bool out64 = o64.obs != nullptr;
  if (h->cfg.dtype == B2H_F64) { if (out64) B2H_LAUNCH_STEP(double, 1); else B2H_LAUNCH_STEP(double, 0); }
  else { if (out64) B2H_LAUNCH_STEP(float, 1); else B2H_LAUNCH_STEP(float, 0); }
#undef B2H_LAUNCH_STEP
  CU(cudaGetLastError());
  h->launches++;
  if (h->step_done) CU(cudaEventRecord(h->step_done, s));   // results are complete here; the sort below is for the next step
  return B2H_OK;
}
static int launch_post_step(B2HHandle* h, const RecordArgs* rec, cudaStream_t s) {
  const bool sched = h->schedule && h->P.sync_mode == 2 && h->cfg.n_envs > h->warps;
  if (!sched && !rec) return B2H_OK;
  // sorted after the step instead of before the next one: it then overlaps the caller's host work
  post_step_kernel<<<1, 1024, 0, s>>>(h->effort, h->cfg.n_envs, h->perm, h->work, sched ? 1 : 0, rec ? 1 : 0, rec ? *rec : RecordArgs(), h->order_key);
  CU(cudaGetLastError());
  if (sched) { h->perm_valid = true; h->work_armed = true; h->armed_stream = s; }
  h->launches++;
  return B2H_OK;
}
static int launch_step(B2HHandle* h, const float* actions_dev, void* obs_dev, void* reward_dev, uint8_t* terminated_dev,
                       uint8_t* truncated_dev, void* terminal_obs_dev, Out64 o64, cudaStream_t s) {
  int rc = launch_step_kernel(h, actions_dev, obs_dev, reward_dev, terminated_dev, truncated_dev, terminal_obs_dev, o64, s);
  return rc != B2H_OK ? rc : launch_post_step(h, nullptr, s);
}

int b2h_step(B2HHandle* h, const float* actions_dev, void* obs_dev, void* reward_dev, uint8_t* terminated_dev,
             uint8_t* truncated_dev, void* terminal_obs_dev, void* stream) {
  if (!h || !actions_dev || !obs_dev || !reward_dev || !terminated_dev || !truncated_dev) return fail(B2H_EINVAL, "null argument");
  CU(cudaSetDevice(h->cfg.device));
  return launch_step(h, actions_dev, obs_dev, reward_dev, terminated_dev, truncated_dev, terminal_obs_dev, Out64(), (cudaStream_t)stream);
}

// Device-visible alias of a page-locked host buffer (cudaHostAlloc / cudaHostRegister / torch pin_memory), or null.
static void* mapped_alias(const void* host) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

int b2h_step_vecenv(B2HHandle* h, const float* actions_host, double* obs_host, double* reward_host, uint8_t* terminated_host,
                    uint8_t* truncated_host, double* terminal_obs_host, int* n_done, void* stream) {
  if (!h || !actions_host || !obs_host || !reward_host || !terminated_host || !truncated_host) return fail(B2H_EINVAL, "null argument");
  CU(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t E = (size_t)h->cfg.n_envs, obs_bytes = E * h->obs_dim * 8;
  CU(cudaMemcpyAsync(h->actions_stage, actions_host, E * h->nu * 4, cudaMemcpyHostToDevice, s));
  Out64 o64;
  o64.obs = (double*)mapped_alias(obs_host);
  o64.reward = (double*)mapped_alias(reward_host);
  uint8_t* term = (uint8_t*)mapped_alias(terminated_host);
  uint8_t* trunc = (uint8_t*)mapped_alias(truncated_host);
  o64.tobs = terminal_obs_host ? (double*)mapped_alias(terminal_obs_host) : nullptr;
  const bool direct = o64.obs && o64.reward && term && trunc && (!terminal_obs_host || o64.tobs);
  if (!direct) {  // pageable host buffers: float64 staging in HBM, copied out after the kernel
    if (!h->obs64_stage) {
      CU(cudaMalloc(&h->obs64_stage, obs_bytes)); CU(cudaMalloc(&h->tobs64_stage, obs_bytes)); CU(cudaMalloc(&h->rew64_stage, E * 8));
    }
    o64.obs = h->obs64_stage; o64.reward = h->rew64_stage; o64.tobs = terminal_obs_host ? h->tobs64_stage : nullptr;
    term = h->term_stage; trunc = h->trunc_stage;
  }
  int rc = launch_step(h, h->actions_stage, nullptr, nullptr, term, trunc, nullptr, o64, s);
  if (rc != B2H_OK) return rc;
  if (!direct) {
    CU(cudaMemcpyAsync(obs_host, o64.obs, obs_bytes, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(reward_host, o64.reward, E * 8, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(terminated_host, term, E, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(truncated_host, trunc, E, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
  } else {
    CU(cudaEventSynchronize(h->step_done));   // the kernel's host writes are complete; the sort for the next step still runs
  }
  int nd = 0;
  for (size_t i = 0; i < E; i++) nd += (terminated_host[i] | truncated_host[i]) != 0;
  if (!direct && nd && terminal_obs_host) {
    CU(cudaMemcpyAsync(terminal_obs_host, o64.tobs, obs_bytes, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
  }
  if (n_done) *n_done = nd;
  return B2H_OK;
}

int b2h_step_host(B2HHandle* h, const float* actions_host, void* obs_host, void* reward_host, uint8_t* terminated_host,
                  uint8_t* truncated_host, void* terminal_obs_host, void* stream) {
  if (!h || !actions_host || !obs_host || !reward_host || !terminated_host || !truncated_host) return fail(B2H_EINVAL, "null argument");
  CU(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  size_t E = (size_t)h->cfg.n_envs, esz = (size_t)h->esz;
  CU(cudaMemcpyAsync(h->actions_stage, actions_host, E * h->nu * 4, cudaMemcpyHostToDevice, s));
  {  // page-locked result buffers: the kernel writes them itself (see b2h_step_vecenv)
    void *obs_a = mapped_alias(obs_host), *rew_a = mapped_alias(reward_host), *tobs_a = terminal_obs_host ? mapped_alias(terminal_obs_host) : nullptr;
    uint8_t *term_a = (uint8_t*)mapped_alias(terminated_host), *trunc_a = (uint8_t*)mapped_alias(truncated_host);
    if (obs_a && rew_a && term_a && trunc_a && (!terminal_obs_host || tobs_a)) {
      int rc = launch_step(h, h->actions_stage, obs_a, rew_a, term_a, trunc_a, tobs_a, Out64(), s);
      if (rc != B2H_OK) return rc;
      CU(cudaEventSynchronize(h->step_done));
      return B2H_OK;
    }
  }
  int rc = b2h_step(h, h->actions_stage, h->obs_stage, h->rew_stage, h->term_stage, h->trunc_stage,
                    terminal_obs_host ? h->tobs_stage : nullptr, stream);
  if (rc != B2H_OK) return rc;
  CU(cudaMemcpyAsync(obs_host, h->obs_stage, E * h->obs_dim * esz, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(reward_host, h->rew_stage, E * esz, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(terminated_host, h->term_stage, E, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(truncated_host, h->trunc_stage, E, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (terminal_obs_host) {  // terminal observations only travel on the steps where an episode ended
    bool any = false;
    for (size_t i = 0; i < E && !any; i++) any = terminated_host[i] || truncated_host[i];
    if (any) {
      CU(cudaMemcpyAsync(terminal_obs_host, h->tobs_stage, E * h->obs_dim * esz, cudaMemcpyDeviceToHost, s));
      CU(cudaStreamSynchronize(s));
    }
  }
  return B2H_OK;
}

int b2h_reset_host(B2HHandle* h, const uint8_t* mask_host, void* obs_host, void* stream) {
  if (!h) return fail(B2H_EINVAL, "null handle");
  CU(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  size_t E = (size_t)h->cfg.n_envs, esz = (size_t)h->esz;
  if (mask_host) CU(cudaMemcpyAsync(h->mask_stage, mask_host, E, cudaMemcpyHostToDevice, s));
  int rc = b2h_reset(h, mask_host ? h->mask_stage : nullptr, h->obs_stage, stream);
  if (rc != B2H_OK) return rc;
  if (obs_host) CU(cudaMemcpyAsync(obs_host, h->obs_stage, E * h->obs_dim * esz, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return B2H_OK;
}

int b2h_set_reset_noise(B2HHandle* h, const double* noise_dev, void* stream) {
  if (!h || !noise_dev) return fail(B2H_EINVAL, "null argument");
  CU(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  size_t E = (size_t)h->cfg.n_envs;
  CU(cudaMemcpyAsync(h->reset_noise, noise_dev, E * (h->nq + h->nv) * 8, cudaMemcpyDeviceToDevice, s));
  CU(cudaMemsetAsync(h->noise_injected, 1, E, s));
  return B2H_OK;
}

int b2h_get_last_reset_noise(B2HHandle* h, double* noise_dev, void* stream) {
  if (!h || !noise_dev) return fail(B2H_EINVAL, "null argument");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaMemcpyAsync(noise_dev, h->reset_noise, (size_t)h->cfg.n_envs * (h->nq + h->nv) * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return B2H_OK;
}

}  // extern "C"
// double host arrays <-> T device arrays
template <typename T>
static int copy_state(B2HHandle* h, double* host, void* dev, size_t n, bool to_host) {
  std::vector<T> tmp(n);
  if (to_host) {
    CU(cudaMemcpy(tmp.data(), dev, n * sizeof(T), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; i++) host[i] = (double)tmp[i];
  } else {
    for (size_t i = 0; i < n; i++) tmp[i] = (T)host[i];
    CU(cudaMemcpy(dev, tmp.data(), n * sizeof(T), cudaMemcpyHostToDevice));
  }
  return B2H_OK;
}
static int copy_any(B2HHandle* h, double* host, void* dev, size_t n, bool to_host) {
  return h->cfg.dtype == B2H_F64 ? copy_state<double>(h, host, dev, n, to_host) : copy_state<float>(h, host, dev, n, to_host);
}
extern "C" {

int b2h_get_state(B2HHandle* h, double* qpos, double* qvel, double* warm, int32_t* nstep, int32_t* step_count, double* total_reward) {
  if (!h) return fail(B2H_EINVAL, "null handle");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaDeviceSynchronize());
  size_t E = (size_t)h->cfg.n_envs;
  int rc = B2H_OK;
  if (qpos && (rc = copy_any(h, qpos, h->qpos, E * h->nq, true))) return rc;
  if (qvel && (rc = copy_any(h, qvel, h->qvel, E * h->nv, true))) return rc;
  if (warm && (rc = copy_any(h, warm, h->warm, E * h->nv, true))) return rc;
  if (total_reward && (rc = copy_any(h, total_reward, h->total_reward, E, true))) return rc;
  if (nstep) CU(cudaMemcpy(nstep, h->nstep, E * 4, cudaMemcpyDeviceToHost));
  if (step_count) CU(cudaMemcpy(step_count, h->step_count, E * 4, cudaMemcpyDeviceToHost));
  return B2H_OK;
}

int b2h_set_state(B2HHandle* h, const double* qpos, const double* qvel, const double* warm, const int32_t* nstep,
                  const int32_t* step_count, const double* total_reward) {
  if (!h) return fail(B2H_EINVAL, "null handle");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaDeviceSynchronize());
  size_t E = (size_t)h->cfg.n_envs;
  int rc = B2H_OK;
  if (qpos && (rc = copy_any(h, (double*)qpos, h->qpos, E * h->nq, false))) return rc;
  if (qvel && (rc = copy_any(h, (double*)qvel, h->qvel, E * h->nv, false))) return rc;
  if (warm && (rc = copy_any(h, (double*)warm, h->warm, E * h->nv, false))) return rc;
  if (total_reward && (rc = copy_any(h, (double*)total_reward, h->total_reward, E, false))) return rc;
  if (nstep) CU(cudaMemcpy(h->nstep, nstep, E * 4, cudaMemcpyHostToDevice));
  if (step_count) CU(cudaMemcpy(h->step_count, step_count, E * 4, cudaMemcpyHostToDevice));
  return B2H_OK;
}

}  // extern "C"
template <typename T>
static int debug_typed(B2HHandle* h, const float* actions_dev, int env, const char* what, double* out, int max_out) {
  debug_kernel<T><<<1, 32, sizeof(Scratch<T>) + model_smem_bytes<T>()>>>((const DevModel<T>*)h->dmodel,
      make_io<T>(h, actions_dev, nullptr, nullptr, nullptr, nullptr, nullptr), env, (DebugDump<T>*)h->dump, (T*)h->spill, h->P.sensor_terms);
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  h->launches++;
  std::vector<unsigned char> hb(sizeof(DebugDump<T>)), mb(sizeof(DevModel<T>));
  CU(cudaMemcpy(hb.data(), h->dump, sizeof(DebugDump<T>), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(mb.data(), h->dmodel, sizeof(DevModel<T>), cudaMemcpyDeviceToHost));
  int n = extract_named<T>(*reinterpret_cast<DevModel<T>*>(mb.data()), *reinterpret_cast<DebugDump<T>*>(hb.data()), what, out, max_out);
  if (n == -2) return fail(B2H_EINVAL, std::string("unknown array name: ") + what);
  if (n < 0) return fail(B2H_EINVAL, "output buffer too small");
  return n;
}

extern "C" {
int b2h_debug_forward(B2HHandle* h, const float* actions_dev, int env, const char* what, double* out_host, int max_out) {
  if (!h || !what || !out_host) return fail(B2H_EINVAL, "null argument");
  if (env < 0 || env >= h->cfg.n_envs) return fail(B2H_EINVAL, "env out of range");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaDeviceSynchronize());
  return h->cfg.dtype == B2H_F64 ? debug_typed<double>(h, actions_dev, env, what, out_host, max_out)
                                 : debug_typed<float>(h, actions_dev, env, what, out_host, max_out);
}

int b2h_measure_fp32_peak(int device, double* tflops) {
  if (!tflops) return fail(B2H_EINVAL, "null argument");
  int ndev = 0, nsm = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(B2H_ECUDA, "no such CUDA device");
  CU(cudaSetDevice(device));
  CU(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
  const int grid = nsm * 8, block = 256, iters = 1 << 15;
  float* out = nullptr;
  CU(cudaMalloc(&out, (size_t)grid * block * sizeof(float)));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {  // first repetition warms the clocks up
    CU(cudaEventRecord(e0));
    ffma_peak_kernel<<<grid, block>>>(out, iters, 0.999f, 0.001f);
    CU(cudaEventRecord(e1));
    CU(cudaEventSynchronize(e1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    double tf = 2.0 * 16 * iters * (double)grid * block / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  *tflops = best;
  return B2H_OK;
}

#ifdef B2H_STAGE_CLOCKS
int b2h_debug_stage_clocks(uint64_t out[48], int reset) {   // tuning build only (not part of include/b2h.h)
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpyFromSymbol(out, g_stage_clk, 48 * 8));
  {  // out[22], out[23]: mean and max CTA exit time of the last launch, in ns after the earliest exit... relative to the min
    static unsigned long long ex[1024];
    CU(cudaMemcpyFromSymbol(ex, g_cta_exit, sizeof(ex)));
    unsigned long long mn = ~0ull, mx = 0, sum = 0; int n = 0;
    for (int i = 0; i < 1024; i++) if (ex[i]) { if (ex[i] < mn) mn = ex[i]; if (ex[i] > mx) mx = ex[i]; n++; }
    for (int i = 0; i < 1024; i++) if (ex[i]) sum += ex[i] - mn;
    out[22] = n ? sum / n : 0; out[23] = n ? mx - mn : 0;
  }
  if (reset) { unsigned long long z[48] = {0}; CU(cudaMemcpyToSymbol(g_stage_clk, z, 24 * 8)); }
  return B2H_OK;
}
#endif

int b2h_get_counters(B2HHandle* h, uint64_t counters_host[8]) {
  if (!h || !counters_host) return fail(B2H_EINVAL, "null argument");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(counters_host, h->counters, 64, cudaMemcpyDeviceToHost));
  counters_host[5] = h->launches;
  return B2H_OK;
}

int b2h_gae(const float* rewards_dev, const float* values_dev, const float* episode_starts_dev, const float* last_values_dev,
            const uint8_t* last_dones_dev, double gamma, double gae_lambda, int T, int E, float* advantages_dev,
            float* returns_dev, void* stream) {
  if (!rewards_dev || !values_dev || !episode_starts_dev || !last_values_dev || !last_dones_dev || !advantages_dev || !returns_dev)
    return fail(B2H_EINVAL, "null argument");
  if (T <= 0 || E <= 0) return fail(B2H_EINVAL, "T and E must be positive");
  // numpy float32 semantics: python-float scalars round to float32 once (gamma, and the double product gamma*lambda)
  gae_kernel<<<(E + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rewards_dev, values_dev, episode_starts_dev, last_values_dev,
      last_dones_dev, nullptr, (float)gamma, (float)(gamma * gae_lambda), T, E, advantages_dev, returns_dev);
  CU(cudaGetLastError());
  return B2H_OK;
}

size_t b2h_sizeof_rollout(void) { return sizeof(B2HRollout); }

int b2h_rollout_collect(B2HHandle* h, const B2HRollout* r, void* stream) {
  if (!h || !r) return fail(B2H_EINVAL, "null argument");
  if (h->cfg.dtype != B2H_F32) return fail(B2H_EUNSUPPORTED, "the device-resident rollout runs on the float32 build");
  if (r->n_steps <= 0 || !r->obs || !r->actions || !r->rewards || !r->values || !r->log_probs || !r->episode_starts || !r->advantages ||
      !r->returns || !r->last_values || !r->mean || !r->clipped || !r->ep_return || !r->ep_len || !r->stats || !r->step_counter ||
      !r->mlp_error || !r->log_std || (r->bootstrap_timeouts && !r->v_term))
    return fail(B2H_EINVAL, "null buffer in B2HRollout");
  for (int k = 0; k < 6; k++) if (!r->pi[k] || !r->vf[k]) return fail(B2H_EINVAL, "null network parameter in B2HRollout");
  CU(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t E = (size_t)h->cfg.n_envs, od = (size_t)h->obs_dim, nu = (size_t)h->nu;
  const int T = r->n_steps;
  if (r->packed) {   // weights are constant within a rollout: split / lay them out for the tensor cores once
    int rc = b2h_policy_pack(r->packed, r->pi, r->vf, stream);
    if (rc != B2H_OK) return fail(rc, std::string("b2h_policy_pack: ") + b2h_mlp_last_error());
  }
  for (int t = 0; t < T; t++) {
    const float* obs_t = r->obs + (size_t)t * E * od;
    int rc = r->packed ? b2h_policy_forward_packed(r->packed, obs_t, r->pi, r->vf, r->mean, r->values + (size_t)t * E, (int)E, r->precise,
                                                   r->mlp_error, stream)
                       : b2h_policy_forward(obs_t, r->pi, r->vf, r->mean, r->values + (size_t)t * E, (int)E, (int)od, r->hidden, (int)nu,
                                            r->precise, r->mlp_error, stream);
    if (rc != B2H_OK) return fail(rc, std::string("b2h_policy_forward: ") + b2h_mlp_last_error());
    rc = b2h_policy_sample_dev(r->mean, r->log_std, (int)E, (int)nu, r->seed, r->step_counter, 0, r->row_offset, r->deterministic,
                               r->actions + (size_t)t * E * nu, r->clipped, r->log_probs + (size_t)t * E, stream);
    if (rc != B2H_OK) return fail(rc, std::string("b2h_policy_sample: ") + b2h_mlp_last_error());
    // the env step writes the next observation and the reward straight into the rollout buffers
    rc = launch_step_kernel(h, r->clipped, r->obs + (size_t)(t + 1) * E * od, r->rewards + (size_t)t * E, h->term_stage, h->trunc_stage,
                            h->tobs_stage, Out64(), s);
    if (rc != B2H_OK) return rc;
    if (r->bootstrap_timeouts) {   // V(terminal_obs): rows of envs that did not finish are stale and unused
      rc = r->packed ? b2h_policy_forward_packed(r->packed, (const float*)h->tobs_stage, r->pi, r->vf, nullptr, r->v_term, (int)E, r->precise,
                                                 r->mlp_error, stream)
                     : b2h_mlp_forward((const float*)h->tobs_stage, r->vf[0], r->vf[1], r->vf[2], r->vf[3], r->vf[4], r->vf[5], r->v_term, (int)E,
                                       (int)od, r->hidden, 1, r->precise, r->mlp_error, stream);
      if (rc != B2H_OK) return fail(rc, std::string("b2h_mlp_forward: ") + b2h_mlp_last_error());
    }
    RecordArgs rec;
    rec.rewards_t = r->rewards + (size_t)t * E; rec.episode_starts_next = r->episode_starts + (size_t)(t + 1) * E;
    rec.terminated = h->term_stage; rec.truncated = h->trunc_stage; rec.v_term = r->bootstrap_timeouts ? r->v_term : nullptr;
    rec.gamma = (float)r->gamma; rec.ep_return = r->ep_return; rec.ep_len = r->ep_len; rec.stats = r->stats;
    rec.step_counter = reinterpret_cast<unsigned long long*>(r->step_counter);
    rc = launch_post_step(h, &rec, s);
    if (rc != B2H_OK) return rc;
  }
  int rc = r->packed ? b2h_policy_forward_packed(r->packed, r->obs + (size_t)T * E * od, r->pi, r->vf, nullptr, r->last_values, (int)E, r->precise,
                                                 r->mlp_error, stream)
                     : b2h_mlp_forward(r->obs + (size_t)T * E * od, r->vf[0], r->vf[1], r->vf[2], r->vf[3], r->vf[4], r->vf[5], r->last_values,
                                       (int)E, (int)od, r->hidden, 1, r->precise, r->mlp_error, stream);
  if (rc != B2H_OK) return fail(rc, std::string("b2h_mlp_forward: ") + b2h_mlp_last_error());
  // last dones = episode_starts[T] (the dones of the final step), as collect_rollouts passes them
  gae_kernel<<<((int)E + 255) / 256, 256, 0, s>>>(r->rewards, r->values, r->episode_starts, r->last_values, nullptr, r->episode_starts + (size_t)T * E,
                                                  (float)r->gamma, (float)(r->gamma * r->gae_lambda), T, (int)E, r->advantages, r->returns);
  CU(cudaGetLastError());
  return B2H_OK;
}

}  // extern "C"
