// b2h_emu.cpp — TEST-ONLY lane emulation of the warp-per-env kernels (never part of the shipped library).
//
// Compiles mujocoposelearning_b200/csrc/b2h_physics.cuh with -DB2H_HOST_EMU: the 32 lanes of one warp become
// 32 host threads and every shuffle / ballot / __syncwarp becomes a pthread barrier exchange.  It lets the
// CPU test-suite (no GPU in CI) check the kernel *source* against the fp64 oracle stage by stage.  It is
// orders of magnitude slower than anything useful and is not reachable from the package.
#define B2H_HOST_EMU 1
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../mujocoposelearning_b200/csrc/b2h_debug.h"

namespace b2h { namespace emu {
static pthread_barrier_t g_bar;
static thread_local int tl_lane = 0;
static uint64_t g_x[32];
static int g_p[32];
int lane() { return tl_lane; }
void sync() { pthread_barrier_wait(&g_bar); }
uint64_t xchg(uint64_t v, int src) {
  g_x[tl_lane] = v;
  pthread_barrier_wait(&g_bar);
  uint64_t r = g_x[src];
  pthread_barrier_wait(&g_bar);
  return r;
}
unsigned ballot(int p) {
  g_p[tl_lane] = p;
  pthread_barrier_wait(&g_bar);
  unsigned m = 0;
  for (int i = 0; i < 32; i++) if (g_p[i]) m |= 1u << i;
  pthread_barrier_wait(&g_bar);
  return m;
}
} }

using namespace b2h;

template <typename T>
struct Job {
  int mode;  // 0 step, 1 reset, 2 forward-dump
  const DevModel<T>* m;
  Scratch<T>* S;
  T* Jspill;
  EnvParams P;
  EnvIO<T> io;
  int n_envs;
  int dump_env;
  DebugDump<T>* dump;
  Counters cnt[32];
};

template <typename T>
static void* lane_main(void* arg) {
  auto* pr = (std::pair<Job<T>*, int>*)arg;
  Job<T>* j = pr->first;
  emu::tl_lane = pr->second;
  Counters& cnt = j->cnt[pr->second];
  for (int e = 0; e < j->n_envs; e++) {
    if (j->mode == 0) env_step<T>(*j->m, *j->S, j->Jspill, cnt, j->P, j->io, e, true);
    else if (j->mode == 1) env_reset_only<T>(*j->m, *j->S, j->Jspill, cnt, j->P, j->io, e);
    else if (e == j->dump_env) {
      const int lane = emu::tl_lane, nq = j->m->nq, nv = j->m->nv, nu = j->m->nu;
      EnvState<T> st;
      st.qp = lane < nq ? j->io.qpos[(size_t)e * nq + lane] : T(0);
      st.qv = lane < nv ? j->io.qvel[(size_t)e * nv + lane] : T(0);
      st.warm = lane < nv ? j->io.warm[(size_t)e * nv + lane] : T(0);
      st.nstep = j->io.nstep[e];
      st.qfrc_act = 0;
      int a = lane < nv ? j->m->dof_act[lane] : -1;
      st.ctrl = (a >= 0 && j->io.actions) ? T(j->io.actions[(size_t)e * nu + a]) : T(0);
      T qacc;
      if (j->P.sensor_terms) physics_step<T, true, NSLOT, true>(*j->m, *j->S, j->Jspill, st, cnt, false, &j->dump->stats, &qacc, j->dump);
      else physics_step<T, true>(*j->m, *j->S, j->Jspill, st, cnt, false, &j->dump->stats, &qacc, j->dump);
      emu::sync();
    }
  }
  return nullptr;
}

template <typename T>
static void run_lanes(Job<T>& job) {
  pthread_barrier_init(&emu::g_bar, nullptr, 32);
  pthread_t th[32];
  std::pair<Job<T>*, int> args[32];
  memset(job.cnt, 0, sizeof job.cnt);
  for (int i = 0; i < 32; i++) { args[i] = {&job, i}; pthread_create(&th[i], nullptr, lane_main<T>, &args[i]); }
  for (int i = 0; i < 32; i++) pthread_join(th[i], nullptr);
  pthread_barrier_destroy(&emu::g_bar);
}

// All host-facing arrays are double (state) / float (actions); converted to T inside.
template <typename T>
static int emu_run(int mode, const B2HModel* model, const B2HConfig* cfg, double* qpos, double* qvel, double* warm,
                   int* nstep, int* step_count, int* episode, double* total_reward, double* reset_noise,
                   uint8_t* noise_injected, const float* actions, double* obs, double* reward, double* terminal_obs,
                   uint8_t* terminated, uint8_t* truncated, uint64_t* counters, int dump_env, const char* what,
                   double* dump_out, int dump_max) {
  static DevModel<T> dm;
  std::string err = build_dev_model<T>(*model, dm);
  if (!err.empty()) { fprintf(stderr, "emu: %s\n", err.c_str()); return -3; }
  const int E = cfg->n_envs, nq = dm.nq, nv = dm.nv;
  const int obs_dim = cfg->obs_mode == B2H_OBS_QPOS_QVEL ? nq - 2 + nv : nq - 2 + nv + 16 * dm.nbody + nv;
  std::vector<T> q(E * nq), v(E * nv), w(E * nv), tr(E), ob((size_t)E * obs_dim), rw(E), tob((size_t)E * obs_dim);
  for (int i = 0; i < E * nq; i++) q[i] = (T)qpos[i];
  for (int i = 0; i < E * nv; i++) { v[i] = (T)qvel[i]; w[i] = (T)warm[i]; }
  for (int i = 0; i < E; i++) tr[i] = (T)total_reward[i];
  static Scratch<T> S;
  static DebugDump<T> dump;
  static T spill[(NROW - NROW_S) * LD];
  Job<T> job;
  job.Jspill = spill;
  job.mode = mode; job.m = &dm; job.S = &S; job.n_envs = E; job.dump_env = dump_env; job.dump = &dump;
  job.P.frame_skip = cfg->frame_skip; job.P.reward_type = cfg->reward_type; job.P.obs_mode = cfg->obs_mode;
  job.P.max_steps = cfg->max_steps; job.P.duration = cfg->duration; job.P.timestep = model->timestep;
  for (int k = 0; k < 9; k++) job.P.kneel[k] = cfg->kneeling_params[k];
  job.P.seed = cfg->seed; job.P.env_id_offset = cfg->env_id_offset; job.P.sync_mode = 0; job.P.sensor_terms = cfg->sensor_terms != 0; job.P.auto_reset = cfg->no_auto_reset == 0;
  job.io.qpos = q.data(); job.io.qvel = v.data(); job.io.warm = w.data(); job.io.nstep = nstep; job.io.step_count = step_count;
  job.io.episode = episode; job.io.total_reward = tr.data(); job.io.reset_noise = reset_noise; job.io.noise_injected = noise_injected;
  job.io.actions = actions; job.io.obs = ob.data(); job.io.reward = rw.data(); job.io.terminal_obs = tob.data();
  job.io.terminated = terminated; job.io.truncated = truncated; job.io.obs_dim = obs_dim; job.io.work = nullptr; job.io.obs64 = nullptr; job.io.reward64 = nullptr; job.io.terminal_obs64 = nullptr;
  run_lanes<T>(job);
  if (mode == 2) return extract_named<T>(dm, dump, what, dump_out, dump_max);
  for (int i = 0; i < E * nq; i++) qpos[i] = (double)q[i];
  for (int i = 0; i < E * nv; i++) { qvel[i] = (double)v[i]; warm[i] = (double)w[i]; }
  for (int i = 0; i < E; i++) { total_reward[i] = (double)tr[i]; if (reward) reward[i] = (double)rw[i]; }
  if (obs) for (size_t i = 0; i < (size_t)E * obs_dim; i++) obs[i] = (double)ob[i];
  if (terminal_obs) for (size_t i = 0; i < (size_t)E * obs_dim; i++) terminal_obs[i] = (double)tob[i];
  if (counters) {
    const Counters& c = job.cnt[0];
    counters[0] = c.physics_steps; counters[1] = c.contact_overflow; counters[2] = c.iter_cap; counters[3] = c.bad_state;
    counters[4] = c.newton_iter; counters[5] = c.ls_eval;
  }
  return 0;
}

extern "C" int emu_run_any(int use_f64, int mode, const B2HModel* model, const B2HConfig* cfg, double* qpos, double* qvel,
                           double* warm, int* nstep, int* step_count, int* episode, double* total_reward,
                           double* reset_noise, uint8_t* noise_injected, const float* actions, double* obs, double* reward,
                           double* terminal_obs, uint8_t* terminated, uint8_t* truncated, uint64_t* counters, int dump_env,
                           const char* what, double* dump_out, int dump_max) {
  if (use_f64)
    return emu_run<double>(mode, model, cfg, qpos, qvel, warm, nstep, step_count, episode, total_reward, reset_noise,
                           noise_injected, actions, obs, reward, terminal_obs, terminated, truncated, counters, dump_env,
                           what, dump_out, dump_max);
  return emu_run<float>(mode, model, cfg, qpos, qvel, warm, nstep, step_count, episode, total_reward, reset_noise,
                        noise_injected, actions, obs, reward, terminal_obs, terminated, truncated, counters, dump_env, what,
                        dump_out, dump_max);
}
