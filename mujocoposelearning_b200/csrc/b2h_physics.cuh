// b2h_physics.cuh — one warp simulates one humanoid environment (warp-cooperative mj_step + env epilogue).
//
// Replaces, for the MJCF subset of the reference model, what `mujoco.mj_step` does at custom_env.py:121,160
// (MuJoCo 3.2.5 pipeline: kinematics, comPos, crb, collision, makeConstraint, comVel, passive, rne, actuation,
// Newton constraint solver, implicit-damping Euler) and what HumanoidEnv.step/reset/_get_state and the
// reward functions do around it (custom_env.py:97-261, reward_functions.py:66-261).  This is not a port of the
// C engine: the recursion over the kinematic tree is re-expressed as lane-parallel sums over ancestor chains
// and depth-first subtree ranges, lanes are bodies / joints / dofs / geom pairs / constraint rows depending on
// the stage, per-env matrices live in shared memory with 128-bit row access, and cross-lane traffic is warp
// shuffles.  Joint-limit rows are never materialised (their Jacobian is +-e_dof, kept lane-resident).
//
// The same source compiles for sm_100a and, with -DB2H_HOST_EMU, for a 32-host-thread lane emulation used only
// by the CPU test-suite (tests/emu): every shuffle / ballot / sync below is executed by all 32 lanes.
#pragma once
#include "b2h_model.h"

#ifdef B2H_HOST_EMU
#define B2H_DEV inline
#define B2H_DEV_NOINLINE
#define B2H_LDG(x) (x)
namespace b2h { namespace emu {
int lane();
uint64_t xchg(uint64_t v, int src);
unsigned ballot(int p);
void sync();
} }
#else
#define B2H_DEV __device__ __forceinline__
#define B2H_DEV_NOINLINE __device__ __noinline__
#define B2H_LDG(x) (x)   // model tables: every kernel stages them at the start of the CTA's shared memory (stage_model)
extern __shared__ __align__(16) unsigned char b2h_model_smem[];  // the CTA's dynamic shared memory (model tables first)
#endif

namespace b2h {

// ------------------------------------------------------------------------------------------------ warp layer
#ifdef B2H_HOST_EMU
B2H_DEV int lane_id() { return emu::lane(); }
B2H_DEV void wsync() { emu::sync(); }
B2H_DEV void cta_sync() {}
B2H_DEV unsigned ballot(bool p) { return emu::ballot(p); }
template <typename V> B2H_DEV V shfl(V v, int src) {
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(V));
  bits = emu::xchg(bits, src & 31);
  V r;
  memcpy(&r, &bits, sizeof(V));
  return r;
}
#else
B2H_DEV int lane_id() { return threadIdx.x & 31; }
B2H_DEV void wsync() { __syncwarp(); }
B2H_DEV void cta_sync() { __syncthreads(); }
B2H_DEV unsigned ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
template <typename V> B2H_DEV V shfl(V v, int src) { return __shfl_sync(0xffffffffu, v, src); }
#endif
template <typename V> B2H_DEV V shfl_xor(V v, int m) { return shfl(v, lane_id() ^ m); }
template <typename V> B2H_DEV V wsum(V v) {  // butterfly all-reduce: every lane ends with the same bits
  v += shfl_xor(v, 16); v += shfl_xor(v, 8); v += shfl_xor(v, 4); v += shfl_xor(v, 2); v += shfl_xor(v, 1);
  return v;
}
B2H_DEV int wscan_excl(int v, int lane) {  // exclusive prefix sum over lanes
  int x = v;
  for (int o = 1; o < 32; o <<= 1) { int y = shfl(x, lane - o); if (lane >= o) x += y; }
  return x - v;
}
B2H_DEV int popc(unsigned x) {
#ifdef B2H_HOST_EMU
  return __builtin_popcount(x);
#else
  return __popc(x);
#endif
}
B2H_DEV int ffs32(unsigned x) {  // 1-based index of the lowest set bit
#ifdef B2H_HOST_EMU
  return __builtin_ffs((int)x);
#else
  return __ffs((int)x);
#endif
}

// ------------------------------------------------------------------------------------------------ scalar math
B2H_DEV float m_sqrt(float x) { return sqrtf(x); }
B2H_DEV double m_sqrt(double x) { return sqrt(x); }
B2H_DEV float m_abs(float x) { return fabsf(x); }
B2H_DEV double m_abs(double x) { return fabs(x); }
B2H_DEV float m_min(float a, float b) { return fminf(a, b); }
B2H_DEV double m_min(double a, double b) { return fmin(a, b); }
B2H_DEV float m_max(float a, float b) { return fmaxf(a, b); }
B2H_DEV double m_max(double a, double b) { return fmax(a, b); }
#ifdef B2H_HOST_EMU
B2H_DEV float m_exp(float x) { return expf(x); }
#else
B2H_DEV float m_exp(float x) { return __expf(x); }   // reward terms only (2 ulp + range reduction in the MUFU path)
#endif
B2H_DEV double m_exp(double x) { return exp(x); }
B2H_DEV float m_pow(float a, float b) { return powf(a, b); }
B2H_DEV double m_pow(double a, double b) { return pow(a, b); }
B2H_DEV float m_atan2(float a, float b) { return atan2f(a, b); }
B2H_DEV double m_atan2(double a, double b) { return atan2(a, b); }
B2H_DEV float m_asin(float a) { return asinf(a); }
B2H_DEV double m_asin(double a) { return asin(a); }
#ifdef B2H_HOST_EMU
B2H_DEV void m_sincos(float x, float* s, float* c) { *s = sinf(x); *c = cosf(x); }
#else
// half joint angles and h * |omega| / 2: always inside [-pi, pi], where the MUFU sine / cosine are good to 2^-21.4
// absolute -- below fp32 resolution of a unit quaternion component; sinf / cosf carry a 350-instruction slow path
B2H_DEV void m_sincos(float x, float* s, float* c) { __sincosf(x, s, c); }
#endif
B2H_DEV void m_sincos(double x, double* s, double* c) { *s = sin(x); *c = cos(x); }
template <typename T> B2H_DEV T clampT(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }
template <typename T> B2H_DEV bool is_bad(T x) { return !(x == x) || x > T(1e10) || x < T(-1e10); }

template <typename T> struct Tol;  // arithmetic-type dependent solver slack (0 in double = MuJoCo's tests verbatim)
template <> struct Tol<double> { static constexpr double ls_rel = 0.0, cost_rel = 0.0, step_rel = 0.0; static constexpr bool exact_stop = false; static constexpr int maxiter = 100; };
// float: MuJoCo's absolute 1e-8 tests sit below fp32 round-off.  The cost is piecewise quadratic, so a full Newton
// step that leaves the active set unchanged lands on the minimiser exactly (exact_stop); otherwise stop when the
// step is below the resolution of qacc (step_rel) or the cost stops resolving (cost_rel).
template <> struct Tol<float> { static constexpr float ls_rel = 2e-5f, cost_rel = 1e-6f, step_rel = 2e-7f; static constexpr bool exact_stop = true; static constexpr int maxiter = 30; };

#define B2H_MINVAL T(1e-15)
#ifndef B2H_LS_FUSED
#define B2H_LS_FUSED 0   // 1: bracketing phase of the line search as passes of three independent evaluations (measured)
#endif
// the (at most NSLOT) dense-row slots a lane owns, fully unrolled so per-slot registers stay registers
#define B2H_SLOTS(s) _Pragma("unroll") for (int s = 0; s < NSLOT; s++) if (s < nslot)

template <typename T> struct V4 { T x, y, z, w; };
#ifdef B2H_HOST_EMU
template <typename T> B2H_DEV V4<T> ld4(const T* p) { return V4<T>{p[0], p[1], p[2], p[3]}; }
#else
B2H_DEV V4<float> ld4(const float* p) { float4 v = *reinterpret_cast<const float4*>(p); return V4<float>{v.x, v.y, v.z, v.w}; }
B2H_DEV V4<double> ld4(const double* p) {
  double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
  return V4<double>{a.x, a.y, b.x, b.y};
}
#endif

template <typename T> B2H_DEV T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <typename T> B2H_DEV void cross3(T* r, const T* a, const T* b) {
  T x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> B2H_DEV T normalize3(T* v) {  // mju_normalize3
  T n = m_sqrt(dot3(v, v));
  if (n < B2H_MINVAL) { v[0] = 1; v[1] = 0; v[2] = 0; }
  else { T s = T(1) / n; v[0] *= s; v[1] *= s; v[2] *= s; }
  return n;
}
template <typename T> B2H_DEV void normalize4(T* q) {  // mju_normalize4
  T n = m_sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < B2H_MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
  else { T s = T(1) / n; q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s; }
}
template <typename T> B2H_DEV void mul_quat(T* r, const T* a, const T* b) {
  T t0 = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  T t1 = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  T t2 = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  T t3 = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = t0; r[1] = t1; r[2] = t2; r[3] = t3;
}
template <typename T> B2H_DEV void quat2mat(T* R, const T* q) {
  T q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  T q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3], q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  R[0] = q00 + q11 - q22 - q33; R[4] = q00 - q11 + q22 - q33; R[8] = q00 - q11 - q22 + q33;
  R[1] = 2 * (q12 - q03); R[2] = 2 * (q13 + q02); R[3] = 2 * (q12 + q03);
  R[5] = 2 * (q23 - q01); R[6] = 2 * (q13 - q02); R[7] = 2 * (q23 + q01);
}
template <typename T> B2H_DEV void rot_quat(T* r, const T* v, const T* q) {  // r = R(q) v via q v q*
  T t[3], u[3] = {q[1], q[2], q[3]};
  cross3(t, u, v);
  t[0] *= 2; t[1] *= 2; t[2] *= 2;
  T c[3];
  cross3(c, u, t);
  r[0] = v[0] + q[0] * t[0] + c[0]; r[1] = v[1] + q[0] * t[1] + c[1]; r[2] = v[2] + q[0] * t[2] + c[2];
}
template <typename T> B2H_DEV void mat_vec3(T* r, const T* R, const T* v) {
  T x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2], y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2],
    z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
// spatial 6-vectors are [angular; linear]; inertia is the 10-number cinert format
template <typename T> B2H_DEV void mul_inert_vec(T* r, const T* i, const T* v) {
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
template <typename T> B2H_DEV void cross_motion(T* r, const T* vel, const T* v) {
  r[0] = -vel[2] * v[1] + vel[1] * v[2];
  r[1] = vel[2] * v[0] - vel[0] * v[2];
  r[2] = -vel[1] * v[0] + vel[0] * v[1];
  r[3] = -vel[2] * v[4] + vel[1] * v[5] - vel[5] * v[1] + vel[4] * v[2];
  r[4] = vel[2] * v[3] - vel[0] * v[5] + vel[5] * v[0] - vel[3] * v[2];
  r[5] = -vel[1] * v[3] + vel[0] * v[4] - vel[4] * v[0] + vel[3] * v[1];
}
template <typename T> B2H_DEV void cross_force(T* r, const T* vel, const T* f) {
  r[0] = -vel[2] * f[1] + vel[1] * f[2] - vel[5] * f[4] + vel[4] * f[5];
  r[1] = vel[2] * f[0] - vel[0] * f[2] + vel[5] * f[3] - vel[3] * f[5];
  r[2] = -vel[1] * f[0] + vel[0] * f[1] - vel[4] * f[3] + vel[3] * f[4];
  r[3] = -vel[2] * f[4] + vel[1] * f[5];
  r[4] = vel[2] * f[3] - vel[0] * f[5];
  r[5] = -vel[1] * f[3] + vel[0] * f[4];
}

// ------------------------------------------------------------------------------------------------ per-warp scratch
template <typename T>
struct alignas(16) Scratch {
  T A[LD * LD];      // factor transposition buffer of chol_solve_fused; stage-local scratch TMP_* otherwise
  T M[LD * LD];      // joint-space inertia, dense symmetric
  T vec[3][32];      // lane vectors that other lanes index (qpos, qvel, matvec operand)
  T cinert[KB * 10], cvel[KB * 6];   // live until the observation is written (custom_env.py:242-256)
  T cdof[KV * 6];
  T con_dist[NCON], con_pos[NCON * 3], con_frame[NCON * 9];
  uint32_t con_info[NCON];  // body1 | body2 << 8 | class << 16
  int con_row[NCON];        // first dense row of the contact, -1 if dropped
  int row_con[NROW];        // dense row -> contact id (or -1 - tendon id)
  T com[4];
  // Last member: the first m.nrow_s (<= NROW_S) dense constraint rows (tendon limits, contact rows); column 27 is a
  // zero pad.  Until the rows are written (after collision) it holds the kinematics scratch POS_*.  The launch sizes
  // each warp's slice as scratch_bytes(nrow_s), so fewer shared rows buy more env-warps per SM.
  alignas(16) T J[NROW_S * LD];
};
template <typename T> constexpr size_t scratch_bytes(int nrow_s) { return sizeof(Scratch<T>) - (size_t)(NROW_S - nrow_s) * LD * sizeof(T); }
constexpr int NROW_S_MIN = 21;
// kinematics / collision scratch inside Scratch::J (dead before the first constraint row is written)
constexpr int POS_XPOS = 0, POS_XMAT = POS_XPOS + KB * 3, POS_XIPOS = POS_XMAT + KB * 9, POS_XANCHOR = POS_XIPOS + KB * 3,
              POS_XAXIS = POS_XANCHOR + KJ * 3, POS_GPOS = POS_XAXIS + KJ * 3, POS_GAXIS = POS_GPOS + KG * 3,
              POS_END = POS_GAXIS + KG * 3;
static_assert(POS_END <= NROW_S_MIN * LD, "kinematics scratch must fit in the row storage");
// stage-local aliases inside Scratch::A (all dead before the factorisations start)
constexpr int TMP_QLOC = 0, TMP_ANCL = TMP_QLOC + KJ * 4, TMP_AXL = TMP_ANCL + KJ * 3, TMP_XQUAT = TMP_AXL + KJ * 3,
              TMP_CRB = TMP_XQUAT + KB * 4;                       // position stage
constexpr int TMP_DOFW = 0, TMP_DOFA = TMP_DOFW + KV * 6, TMP_CFRC = TMP_DOFA + KV * 6, TMP_CDD = TMP_CFRC + KB * 6;  // velocity stage
static_assert(TMP_CRB + KB * 10 <= LD * LD, "position-stage scratch must fit in A");
static_assert(TMP_CDD + KV * 6 <= LD * LD, "velocity-stage scratch must fit in A");

constexpr int B2H_EFFORT_BITS = 20;   // EnvIO::work packs the effort (low bits) and the rows of the last control step
constexpr unsigned B2H_EFFORT_MASK = (1u << B2H_EFFORT_BITS) - 1u;
struct Counters {  // per-warp tallies of one launch (32-bit: a warp sees a few thousand events), flushed with atomics
  unsigned physics_steps, contact_overflow, iter_cap, bad_state, newton_iter, ls_eval;
  unsigned work;  // solver effort of the env in flight (Newton iterations weighted by row slots): next launch's schedule key
  unsigned rows;  // most dense constraint rows of a physics step of the env in flight: next launch's slot-count hint
  unsigned nlim;  // most active joint limits of a physics step of the env in flight (schedule key experiments)
  int sync_threads;  // tuning experiment (B2H_EXP_NEWTON_BARRIER): threads of the lockstep group that meet again before the solver, 0 = none
#ifdef B2H_STAGE_CLOCKS
  long long clk[48];  // tuning build: cycles per stage (tools/stage_clocks.py)
#endif
};
#ifdef B2H_STAGE_CLOCKS
#define B2H_CLK(var) long long var = clock64()
#define B2H_CLK_FROM(var, src) long long var = (src)
#define B2H_CLK_ADD(i, t0) do { long long t1_ = clock64(); cnt.clk[i] += t1_ - (t0); (t0) = t1_; } while (0)
#define B2H_TALLY(i) (cnt.clk[i] += 1)
#else
#define B2H_TALLY(i)
#define B2H_CLK(var)
#define B2H_CLK_FROM(var, src)
#define B2H_CLK_ADD(i, t0)
#endif

// impedance curve for a general power (solimp[4] other than 1 or 2): kept out of line, the model on the path uses 2
template <typename T>
B2H_DEV_NOINLINE T imp_power_curve(T x, T p, T mid) {
  return x <= mid ? m_pow(x, p) / m_pow(mid, p - 1) : T(1) - m_pow(T(1) - x, p) / m_pow(T(1) - mid, p - 1);
}
// per-row soft-constraint parameters (mj_makeImpedance + mj_referenceConstraint)
template <typename T> struct RowParam { T D, aref; };
template <typename T>
B2H_DEV_NOINLINE RowParam<T> row_params(T solref0, T solref1, T solimp0, T solimp1, T solimp2, T solimp3, T solimp4, T pos,
                                        T margin, T diag_approx, T vel, T rscale) {
  const T solref[2] = {solref0, solref1}, solimp[5] = {solimp0, solimp1, solimp2, solimp3, solimp4};
  T imp;
  if (solimp[0] == solimp[1] || solimp[2] <= B2H_MINVAL) imp = T(0.5) * (solimp[0] + solimp[1]);
  else {
    T x = m_abs((pos - margin) / solimp[2]);
    if (x >= T(1)) imp = solimp[1];
    else if (x <= T(0)) imp = solimp[0];
    else {
      T y, p = solimp[4], mid = solimp[3];
      if (p == T(1)) y = x;
      else if (p == T(2)) y = x <= mid ? x * x / mid : T(1) - (T(1) - x) * (T(1) - x) / (T(1) - mid);
      else y = imp_power_curve(x, p, mid);
      imp = solimp[0] + y * (solimp[1] - solimp[0]);
    }
  }
  T dmax = solimp[1], K, B;
  if (solref[0] <= T(0)) { K = -solref[0] / m_max(B2H_MINVAL, dmax * dmax); B = -solref[1] / m_max(B2H_MINVAL, dmax); }
  else {
    K = T(1) / m_max(B2H_MINVAL, dmax * dmax * solref[0] * solref[0] * solref[1] * solref[1]);
    B = T(2) / m_max(B2H_MINVAL, dmax * solref[0]);
  }
  T R = m_max(B2H_MINVAL, (T(1) - imp) * diag_approx / imp) * rscale;
  RowParam<T> out;
  out.D = T(1) / R;
  out.aref = -B * vel - K * imp * (pos - margin);
  return out;
}

// ------------------------------------------------------------------------------------------------ dense 27x27 algebra
B2H_DEV float m_rsqrt(float x) {  // callers keep x >= mjMINVAL, so the flush-to-zero approximation needs no range fix-up
#ifdef B2H_HOST_EMU
  return 1.0f / sqrtf(x);
#else
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#endif
}
B2H_DEV double m_rsqrt(double x) { return 1.0 / sqrt(x); }
B2H_DEV float m_rcp(float x) {
#ifdef B2H_HOST_EMU
  return 1.0f / x;
#else
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#endif
}
B2H_DEV double m_rcp(double x) { return 1.0 / x; }
template <typename T> B2H_DEV void st4(T* p, T x, T y, T z, T w) {
#ifdef B2H_HOST_EMU
  p[0] = x; p[1] = y; p[2] = z; p[3] = w;
#else
  if constexpr (sizeof(T) == 4) *reinterpret_cast<float4*>(p) = make_float4(x, y, z, w);
  else { *reinterpret_cast<double2*>(p) = make_double2(x, y); *reinterpret_cast<double2*>(p + 2) = make_double2(z, w); }
#endif
}

// x = A^-1 b for the SPD matrix in shared memory A (stride LD, lower triangle read, contents destroyed).
// Lane i keeps row i of the factor in registers; the right-hand side rides along as row 27 of the augmented
// matrix (lane 27), so the forward substitution is the same 27 column steps as the factorisation.
//   step j: every lane stores its (not yet scaled) entry of column j to shared memory; the pivot and the entries
//   below it come back as 128-bit broadcast loads (one shared-memory wavefront per four entries, where a shuffle
//   per entry would cost one each), then a[k] -= (a[j] / pivot) * column[k].
// The factor (and y = L^-1 b in row 27) then goes back to shared memory by rows and is read by columns for the
// backward substitution.  mju_cholFactor's pivot floor (mjMINVAL) is kept.
template <typename T>
B2H_DEV_NOINLINE T chol_solve_fused(T* A, int n, int lane, T b) {
  constexpr int N = KV - 1;
#if defined(B2H_HOST_EMU) && defined(B2H_EXP_CHOL_F64)
  // EXPERIMENT (CPU lane emulation only): the same factor + solve carried in double, to separate
  // the round-off of the fp32 solve from the round-off already present in its fp32 inputs (M, J, D, right-hand side)
  if constexpr (sizeof(T) == 4) {
    static double D[LD * LD];
    if (lane < LD) A[N * LD + lane] = lane < n ? b : T(0);
    if (lane >= n && lane < N) A[lane * LD + lane] = 1;
    wsync();
    double a[LD];
    for (int c = 0; c < LD; c++) a[c] = (double)A[(lane < LD ? lane : 0) * LD + c];
    wsync();
    for (int j = 0; j < N; j++) {
      if (lane < LD) D[j * LD + lane] = a[j];
      wsync();
      double piv = D[j * LD + j] > 1e-15 ? D[j * LD + j] : 1e-15, r = 1.0 / sqrt(piv), lr = a[j] * r * r;
      a[j] = a[j] * r;
      for (int k = j + 1; k < N; k++) a[k] -= lr * D[j * LD + k];
    }
    wsync();
    if (lane < LD) for (int c = 0; c < LD; c++) D[lane * LD + c] = a[c];
    wsync();
    const int me = lane < N ? lane : 0;
    double dinv = 1.0 / D[me * LD + me], y = D[N * LD + me], acc = 0;
    for (int k = N - 1; k > 0; k--) {
      double xk = shfl((y - acc) * dinv, k);
      if (lane < k) acc += D[k * LD + lane] * xk;
    }
    double x = (y - acc) * dinv;
    wsync();
    return lane < n ? (T)x : T(0);
  }
#endif
  if (lane < LD) A[N * LD + lane] = lane < n ? b : T(0);
  if (lane >= n && lane < N) A[lane * LD + lane] = 1;  // unused dof slots factor as identity
  wsync();
  T a[LD];
  T* const row = A + (lane < LD ? lane : 0) * LD;
#pragma unroll
  for (int c = 0; c < LD; c += 4) {
    V4<T> v = ld4(row + c);
    a[c] = v.x; a[c + 1] = v.y; a[c + 2] = v.z; a[c + 3] = v.w;
  }
  wsync();  // rows are in registers: A is free to stage the columns
#pragma unroll
  for (int j = 0; j < N; j++) {
    T* const col = A + j * LD;
    if (lane < LD) col[lane] = a[j];
    wsync();
    T c[LD];
#pragma unroll
    for (int k4 = (j / 4) * 4; k4 < LD; k4 += 4) {
      V4<T> v = ld4(col + k4);
      c[k4] = v.x; c[k4 + 1] = v.y; c[k4 + 2] = v.z; c[k4 + 3] = v.w;
    }
    T piv = m_max(c[j], B2H_MINVAL);
    T r = m_rsqrt(piv);
    T lr = a[j] * r * r;  // L[lane][j] / L[j][j]
    a[j] = a[j] * r;
#pragma unroll
    for (int k = j + 1; k < N; k++) a[k] -= lr * c[k];
  }
  wsync();
  if (lane < LD) {  // rows of L (lane 27: y = L^-1 b) to shared memory, read back by columns below
#pragma unroll
    for (int c = 0; c < LD; c += 4) st4(row + c, a[c], a[c + 1], a[c + 2], a[c + 3]);
  }
  wsync();
  const int me = lane < N ? lane : 0;
  const T dinv = m_rcp(A[me * LD + me]);  // 1 / L[lane][lane]
  const T y = A[N * LD + me];
  T acc = 0;
#pragma unroll
  for (int k = N - 1; k > 0; k--) {  // x[k] is final in lane k once the columns above it are folded into acc
    T xk = shfl((y - acc) * dinv, k);
    if (lane < k) acc += A[k * LD + lane] * xk;
  }
  T x = (y - acc) * dinv;
  wsync();
  return lane < n ? x : T(0);
}
// r[lane] = sum_k Mat[lane][k] * v[k], v taken from a 32-entry shared vector (entries >= n are zero)
template <typename T>
B2H_DEV T mat_vec(const T* Mat, const T* v, int n, int lane) {
  T s = 0;
  const T* r = Mat + (lane < n ? lane : 0) * LD;
#pragma unroll
  for (int c = 0; c < LD; c += 4) {
    V4<T> a = ld4(r + c), b = ld4(v + c);
    s += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
  }
  return lane < n ? s : T(0);
}

// ------------------------------------------------------------------------------------------------ env state
template <typename T>
struct EnvState {      // registers of one lane
  T qp, qv, warm;      // qpos[lane], qvel[lane], qacc_warmstart[lane]
  T ctrl;              // control of the motor on dof `lane` (as written to data.ctrl), 0 without motor
  T qfrc_act;          // qfrc_actuator[lane] of the last forward pass
  int nstep;           // physics steps since mj_resetData (time = nstep * timestep)
};

struct StepStats { int ncon, nrow, nlimit, niter; };

template <typename T>
struct DebugDump {  // named views of one mj_forward (parity / debug only)
  T pos[POS_END];      // xpos xmat xipos xanchor xaxis geom_xpos geom_zaxis
  T M[LD * LD], J[NROW * LD];
  T cinert[KB * 10], cvel[KB * 6], cdof[KV * 6], cdofdot[KV * 6];
  T con_dist[NCON], con_pos[NCON * 3], con_frame[NCON * 9], com[4];
  T lane[32][8];       // per-lane: qfrc_bias, qfrc_smooth, qacc_smooth, qacc, qfrc_constraint, qfrc_actuator, limit D, limit aref
  T cfrc_ext[KB * 6], sub_linvel[4];   // EXT instantiation only (section 8 f4)
  StepStats stats;
};

// One mj_step.  On return st holds the integrated state; S.cinert / S.cvel / S.com keep the pre-integration
// values of this step, which is what the reference's observation and rewards read.
// If `integrate` is false this is mj_forward: state untouched, qacc returned in *qacc_out.
// Returns true when mj_checkAcc tripped: the state was reset and the caller must run the step once more.
// DBG = true is the instantiation of the debug / parity kernel (mj_forward with named dumps); the step and reset
// kernels use DBG = false, whose instruction stream carries none of the dump code (cold code inside the lockstep
// stream costs fetch slots even when it is branched over).
// NS = dense-row slots per lane compiled in (capacity 32 * NS rows).  The bench workload needs more than 32 dense rows
// in one step out of a thousand (oracle: 99th percentile of nefc, joint limits included, is 24), so mj_step runs the
// NS = 1 instantiation -- no slot 1 / slot 2 copies of the row loops in the instruction stream -- and falls back to
// the full-capacity one (a separate, out-of-line function) when the rows do not fit: return value 2, state untouched.
enum { B2H_STEP_OK = 0, B2H_STEP_BAD_ACC = 1, B2H_STEP_MORE_ROWS = 2 };
// EXT = true (B2HConfig::sensor_terms, section 8 f4; never on the bench path) also computes what the rewards read from
// data.cfrc_ext[-2], [-1] and data.subtree_linvel[0] (reward_functions.py:109,121-122,176-177) if MuJoCo computed them:
// the contact part of mj_rnePostConstraint and the whole-model row of mj_subtreeVel; results in S.vec[0][0..4].
template <typename T, bool DBG = false, int NS = NSLOT, bool EXT = false>
B2H_DEV_NOINLINE int physics_step(const DevModel<T>& m_arg, Scratch<T>& S, T* Jspill, EnvState<T>& st, Counters& cnt, bool integrate_arg,
                                   StepStats* stats_arg, T* qacc_out_arg, DebugDump<T>* dbg_arg /* optional named dump */) {
  constexpr int NSLOT = NS;   // shadows the capacity constant for everything below (arrays, B2H_SLOTS)
  const bool integrate = DBG ? integrate_arg : true;
  StepStats* const stats = DBG ? stats_arg : nullptr;
  T* const qacc_out = DBG ? qacc_out_arg : nullptr;
  DebugDump<T>* const dbg = DBG ? dbg_arg : nullptr;
#ifndef B2H_HOST_EMU
  // the kernels stage the model tables at the start of the CTA's dynamic shared memory: address them as such
  // (LDS with immediate offsets) instead of through the generic reference this non-inlined function receives
  const DevModel<T>& m = *reinterpret_cast<const DevModel<T>*>(b2h_model_smem);
  (void)m_arg;
#else
  const DevModel<T>& m = m_arg;
#endif
  const int lane = lane_id();
  const int nv = B2H_LDG(m.nv), nq = B2H_LDG(m.nq), nbody = B2H_LDG(m.nbody), njnt = B2H_LDG(m.njnt);
  const T h = B2H_LDG(m.timestep);
  T* tmp = S.A;
  T* const xpos = S.J + POS_XPOS; T* const xmat = S.J + POS_XMAT; T* const xipos = S.J + POS_XIPOS;
  T* const xanchor = S.J + POS_XANCHOR; T* const xaxis = S.J + POS_XAXIS;
  T* const gpos = S.J + POS_GPOS; T* const gaxis = S.J + POS_GAXIS;
  // dense row r lives in shared memory below nrow_s and in this warp's global spill area above it
  const int nrow_s = B2H_LDG(m.nrow_s);
  // (the one-slot instantiation only runs when every row is in shared memory: no select, no generic pointer)
  auto jrow = [&](int r) -> T* {
    if constexpr (NS == 1) return S.J + r * LD;
    else return r < nrow_s ? S.J + r * LD : Jspill + (size_t)(r - nrow_s) * LD;
  };
  B2H_CLK(tc);

  // ---- mj_checkPos / mj_checkVel: NaN or |x| > 1e10 resets mjData (qpos0, zero velocity, time 0)
  {
    bool bad = (lane < nq && is_bad(st.qp)) || (lane < nv && is_bad(st.qv));
    if (ballot(bad)) {
      st.qp = lane < nq ? B2H_LDG(m.qpos0[lane]) : T(0);
      st.qv = 0; st.warm = 0; st.ctrl = 0; st.nstep = 0;
      cnt.bad_state++;
    }
  }

  // =============================================================== position stage
  // ---- mj_kinematics: joint-local rotations (lane = joint), per-body local chain (lane = body), tree levels
  S.vec[0][lane] = st.qp;
  if (lane == 0) {
    xpos[0] = xpos[1] = xpos[2] = 0;
    for (int k = 0; k < 9; k++) xmat[k] = (k % 4 == 0) ? T(1) : T(0);
    tmp[TMP_XQUAT] = 1; tmp[TMP_XQUAT + 1] = tmp[TMP_XQUAT + 2] = tmp[TMP_XQUAT + 3] = 0;
    xipos[0] = xipos[1] = xipos[2] = 0;
    for (int k = 0; k < 10; k++) S.cinert[k] = 0;
    for (int k = 0; k < 6; k++) S.cvel[k] = 0;
  }
  wsync();
  if (lane < njnt && B2H_LDG(m.jnt_type[lane]) == B2H_JNT_HINGE) {
    T ang = S.vec[0][B2H_LDG(m.jnt_qadr[lane])] - B2H_LDG(m.jnt_q0[lane]);
    T s, c;
    m_sincos(ang * T(0.5), &s, &c);
    if (ang == T(0)) { s = 0; c = 1; }
    tmp[TMP_QLOC + 4 * lane] = c;
    for (int k = 0; k < 3; k++) tmp[TMP_QLOC + 4 * lane + 1 + k] = B2H_LDG(m.jnt_axis[lane][k]) * s;
  }
  wsync();
  T rq[4] = {1, 0, 0, 0}, rp[3] = {0, 0, 0};  // body frame relative to its parent body
  const int my_level = lane < nbody ? B2H_LDG(m.body_level[lane]) : -1;
  const int my_parent = lane < nbody ? B2H_LDG(m.body_parent[lane]) : 0;
  if (lane > 0 && lane < nbody) {
    if (B2H_LDG(m.body_isfree[lane])) {
      for (int k = 0; k < 3; k++) rp[k] = S.vec[0][k];
      for (int k = 0; k < 4; k++) rq[k] = S.vec[0][3 + k];
      normalize4(rq);
      int j = B2H_LDG(m.body_jntadr[lane]);
      for (int k = 0; k < 3; k++) { tmp[TMP_ANCL + 3 * j + k] = rp[k]; tmp[TMP_AXL + 3 * j + k] = B2H_LDG(m.jnt_axis[j][k]); }
    } else {
      for (int k = 0; k < 3; k++) rp[k] = B2H_LDG(m.body_pos[lane][k]);
      for (int k = 0; k < 4; k++) rq[k] = B2H_LDG(m.body_quat[lane][k]);
      int ja = B2H_LDG(m.body_jntadr[lane]), jn = B2H_LDG(m.body_jntnum[lane]);
      for (int j = ja; j < ja + jn; j++) {
        T jp[3], jax[3], anc[3], ax[3], v[3];
        for (int k = 0; k < 3; k++) { jp[k] = B2H_LDG(m.jnt_pos[j][k]); jax[k] = B2H_LDG(m.jnt_axis[j][k]); }
        rot_quat(ax, jax, rq);
        rot_quat(anc, jp, rq);
        for (int k = 0; k < 3; k++) anc[k] += rp[k];
        mul_quat(rq, rq, tmp + TMP_QLOC + 4 * j);
        rot_quat(v, jp, rq);
        for (int k = 0; k < 3; k++) { rp[k] = anc[k] - v[k]; tmp[TMP_ANCL + 3 * j + k] = anc[k]; tmp[TMP_AXL + 3 * j + k] = ax[k]; }
      }
    }
  }
  const int nlevel = B2H_LDG(m.nlevel);
  for (int L = 1; L <= nlevel; L++) {
    if (my_level == L) {
      T pq[4], xq[4], v[3];
      for (int k = 0; k < 4; k++) pq[k] = tmp[TMP_XQUAT + 4 * my_parent + k];
      mul_quat(xq, pq, rq);
      normalize4(xq);
      rot_quat(v, rp, pq);
      for (int k = 0; k < 3; k++) xpos[3 * lane + k] = xpos[3 * my_parent + k] + v[k];
      for (int k = 0; k < 4; k++) tmp[TMP_XQUAT + 4 * lane + k] = xq[k];
      T R[9];
      quat2mat(R, xq);
      for (int k = 0; k < 9; k++) xmat[9 * lane + k] = R[k];
    }
    wsync();
  }
  B2H_CLK_FROM(tp, tc);
  B2H_CLK_ADD(12, tp);
  // inertial frames (lane = body), joint anchors/axes in the world (lane = joint), geoms (lane = geom)
  T my_mass = 0;
  if (lane > 0 && lane < nbody) {
    T ip[3], v[3];
    for (int k = 0; k < 3; k++) ip[k] = B2H_LDG(m.body_ipos[lane][k]);
    mat_vec3(v, xmat + 9 * lane, ip);
    for (int k = 0; k < 3; k++) xipos[3 * lane + k] = xpos[3 * lane + k] + v[k];
    my_mass = B2H_LDG(m.body_mass[lane]);
  }
  if (lane < njnt) {
    int p = B2H_LDG(m.body_parent[B2H_LDG(m.jnt_body[lane])]);
    T v[3];
    mat_vec3(v, xmat + 9 * p, tmp + TMP_ANCL + 3 * lane);
    for (int k = 0; k < 3; k++) xanchor[3 * lane + k] = xpos[3 * p + k] + v[k];
    mat_vec3(v, xmat + 9 * p, tmp + TMP_AXL + 3 * lane);
    for (int k = 0; k < 3; k++) xaxis[3 * lane + k] = v[k];
  }
  if (lane < B2H_LDG(m.ngeom)) {
    int b = B2H_LDG(m.geom_body[lane]);
    T gp[3], gz[3], v[3];
    for (int k = 0; k < 3; k++) { gp[k] = B2H_LDG(m.geom_pos[lane][k]); gz[k] = B2H_LDG(m.geom_zaxis[lane][k]); }
    mat_vec3(v, xmat + 9 * b, gp);
    for (int k = 0; k < 3; k++) gpos[3 * lane + k] = xpos[3 * b + k] + v[k];
    mat_vec3(v, xmat + 9 * b, gz);
    for (int k = 0; k < 3; k++) gaxis[3 * lane + k] = v[k];
  }
  wsync();
  // ---- mj_comPos: centre of mass (single tree), cinert, cdof
  T com[3];
  {
    T inv = B2H_LDG(m.inv_total_mass);
    for (int k = 0; k < 3; k++) com[k] = wsum(lane > 0 && lane < nbody ? my_mass * xipos[3 * lane + k] : T(0)) * inv;
    if (lane < 3) S.com[lane] = com[lane];
  }
  if (lane > 0 && lane < nbody) {  // mju_inertCom with the full body-frame tensor: R I R^T + m (|d|^2 1 - d d^T)
    const T* R = xmat + 9 * lane;
    T I6[6], dif[3];
    for (int k = 0; k < 6; k++) I6[k] = B2H_LDG(m.body_inertia[lane][k]);
    for (int k = 0; k < 3; k++) dif[k] = xipos[3 * lane + k] - com[k];
    T Ib[9] = {I6[0], I6[3], I6[4], I6[3], I6[1], I6[5], I6[4], I6[5], I6[2]}, RI[9];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) RI[3 * r + c] = R[3 * r] * Ib[c] + R[3 * r + 1] * Ib[3 + c] + R[3 * r + 2] * Ib[6 + c];
    T W[6];  // xx yy zz xy xz yz
    W[0] = RI[0] * R[0] + RI[1] * R[1] + RI[2] * R[2];
    W[1] = RI[3] * R[3] + RI[4] * R[4] + RI[5] * R[5];
    W[2] = RI[6] * R[6] + RI[7] * R[7] + RI[8] * R[8];
    W[3] = RI[0] * R[3] + RI[1] * R[4] + RI[2] * R[5];
    W[4] = RI[0] * R[6] + RI[1] * R[7] + RI[2] * R[8];
    W[5] = RI[3] * R[6] + RI[4] * R[7] + RI[5] * R[8];
    T* ci = S.cinert + 10 * lane;
    ci[0] = W[0] + my_mass * (dif[1] * dif[1] + dif[2] * dif[2]);
    ci[1] = W[1] + my_mass * (dif[0] * dif[0] + dif[2] * dif[2]);
    ci[2] = W[2] + my_mass * (dif[0] * dif[0] + dif[1] * dif[1]);
    ci[3] = W[3] - my_mass * dif[0] * dif[1];
    ci[4] = W[4] - my_mass * dif[0] * dif[2];
    ci[5] = W[5] - my_mass * dif[1] * dif[2];
    ci[6] = my_mass * dif[0]; ci[7] = my_mass * dif[1]; ci[8] = my_mass * dif[2]; ci[9] = my_mass;
  }
  T cd[6] = {0, 0, 0, 0, 0, 0};  // cdof of dof `lane`
  const int my_dbody = lane < nv ? B2H_LDG(m.dof_body[lane]) : 0;
  const int my_dparent = lane < nv ? B2H_LDG(m.dof_parent[lane]) : -1;
  if (lane < nv) {
    int j = B2H_LDG(m.dof_jnt[lane]);
    T off[3], ax[3];
    for (int k = 0; k < 3; k++) off[k] = com[k] - xanchor[3 * j + k];
    if (B2H_LDG(m.jnt_type[j]) == B2H_JNT_FREE) {
      int k = lane - B2H_LDG(m.jnt_dadr[j]);
      if (k < 3) { cd[3] = k == 0 ? T(1) : T(0); cd[4] = k == 1 ? T(1) : T(0); cd[5] = k == 2 ? T(1) : T(0); }   // (static indices: cd stays in registers)
      else {
        const T* R = xmat + 9 * my_dbody;
        ax[0] = R[k - 3]; ax[1] = R[k]; ax[2] = R[k + 3];
        cd[0] = ax[0]; cd[1] = ax[1]; cd[2] = ax[2];
        cross3(cd + 3, ax, off);
      }
    } else {
      for (int k = 0; k < 3; k++) ax[k] = xaxis[3 * j + k];
      cd[0] = ax[0]; cd[1] = ax[1]; cd[2] = ax[2];
      cross3(cd + 3, ax, off);
    }
  }
  if (lane < KV) for (int k = 0; k < 6; k++) S.cdof[6 * lane + k] = cd[k];
  wsync();
  B2H_CLK_ADD(13, tp);
  // ---- mj_crb: composite inertia = sum over the depth-first subtree range; M over ancestor chains
  if (lane > 0 && lane < nbody) {
    T c[10];
    for (int k = 0; k < 10; k++) c[k] = S.cinert[10 * lane + k];
    int end = B2H_LDG(m.body_subend[lane]);
    for (int b = lane + 1; b < end; b++)
      for (int k = 0; k < 10; k++) c[k] += S.cinert[10 * b + k];
    for (int k = 0; k < 10; k++) tmp[TMP_CRB + 10 * lane + k] = c[k];
  }
  for (int i = lane; i < LD * LD; i += 32) S.M[i] = 0;
  wsync();
  if (lane < nv) {
    T buf[6];
    mul_inert_vec(buf, tmp + TMP_CRB + 10 * my_dbody, cd);
    for (int j = lane; j >= 0; j = B2H_LDG(m.dof_parent[j])) {
      const T* cj = S.cdof + 6 * j;
      T s = cj[0] * buf[0] + cj[1] * buf[1] + cj[2] * buf[2] + cj[3] * buf[3] + cj[4] * buf[4] + cj[5] * buf[5];
      if (j == lane) s += B2H_LDG(m.dof_armature[lane]);
      S.M[lane * LD + j] = s;
      S.M[j * LD + lane] = s;
    }
  }
  wsync();

  B2H_CLK_ADD(14, tp);
  // =============================================================== collision (lane = candidate pair)
  int ncon = 0;
  {
    const int npair = B2H_LDG(m.npair);
    // Broad phase: a pair can only touch if its bounding spheres (capsule: radius + half length) come within the
    // margin (plane: the sphere reaches down to it).  A superset of what the narrow phase accepts, kept in pair
    // order so that the contact list is the same; typically 15-35 of the 159 candidates survive.
    int* const cand = reinterpret_cast<int*>(S.A);   // A is free between mj_crb and the velocity stage
    int ncand = 0;
    for (int base = 0; base < npair; base += 32) {
      int p = base + lane;
      bool keep = false;
      if (p < npair) {
        uint32_t pw = B2H_LDG(m.pair[p]);
        int g1 = pw & 255, g2 = (pw >> 8) & 255;
        T lim = B2H_LDG(m.cls_margin[pw >> 16]) + B2H_LDG(m.geom_size[g2][0]) + B2H_LDG(m.geom_size[g2][1]);
        T d[3] = {gpos[3 * g2] - gpos[3 * g1], gpos[3 * g2 + 1] - gpos[3 * g1 + 1], gpos[3 * g2 + 2] - gpos[3 * g1 + 2]};
        if (B2H_LDG(m.geom_type[g1]) == B2H_GEOM_PLANE) keep = dot3(d, gaxis + 3 * g1) <= lim * T(1.0001) + T(1e-6);
        else {
          lim += B2H_LDG(m.geom_size[g1][0]) + B2H_LDG(m.geom_size[g1][1]);
          keep = dot3(d, d) <= lim * lim * T(1.0001) + T(1e-9);
        }
      }
      unsigned km = ballot(keep);
      if (keep) cand[ncand + popc(km & ((1u << lane) - 1u))] = p;
      ncand += popc(km);
    }
    wsync();
    B2H_CLK_ADD(15, tp);
    for (int base = 0; base < ncand; base += 32) {
      int p = base + lane < ncand ? cand[base + lane] : npair;
      int n = 0;
      T cdist[2] = {}, cpos[2][3] = {}, cnrm[2][3] = {}, chint[3] = {0, 0, 0};
      uint32_t info = 0;
      if (p < npair) {
        uint32_t pw = B2H_LDG(m.pair[p]);
        int g1 = pw & 255, g2 = (pw >> 8) & 255, cls = pw >> 16;
        T margin = B2H_LDG(m.cls_margin[cls]);
        info = (uint32_t)B2H_LDG(m.geom_body[g1]) | ((uint32_t)B2H_LDG(m.geom_body[g2]) << 8) | ((uint32_t)cls << 16);
        int t1 = B2H_LDG(m.geom_type[g1]);
        T r2 = B2H_LDG(m.geom_size[g2][0]), h2 = B2H_LDG(m.geom_size[g2][1]);
        const T* p2 = gpos + 3 * g2;
        const T* a2 = gaxis + 3 * g2;
        if (t1 == B2H_GEOM_PLANE) {  // mjc_PlaneSphere / mjc_PlaneCapsule (two end spheres, frame aligned with the axis)
          const T* pn = gaxis + 3 * g1;
          const T* pp = gpos + 3 * g1;
          int nend = h2 > T(0) ? 2 : 1;
          for (int e = 0; e < nend; e++) {
            T sgn = e == 0 ? T(1) : T(-1), c[3], d[3];
            for (int k = 0; k < 3; k++) { c[k] = p2[k] + a2[k] * h2 * sgn; d[k] = c[k] - pp[k]; }
            T cd_ = dot3(d, pn);
            if (cd_ <= margin + r2) {
              T dist = cd_ - r2;
              if (dist < margin) {
                if (n == 0) { cdist[0] = dist; for (int k = 0; k < 3; k++) { cnrm[0][k] = pn[k]; cpos[0][k] = c[k] - pn[k] * (dist * T(0.5) + r2); } }
                else { cdist[1] = dist; for (int k = 0; k < 3; k++) { cnrm[1][k] = pn[k]; cpos[1][k] = c[k] - pn[k] * (dist * T(0.5) + r2); } }
                n++;
              }
            }
          }
          if (nend == 2) for (int k = 0; k < 3; k++) chint[k] = a2[k];
        } else {  // sphere / capsule pairs: closest points of two segments, then sphere-sphere
          T r1 = B2H_LDG(m.geom_size[g1][0]), h1 = B2H_LDG(m.geom_size[g1][1]);
          const T* p1 = gpos + 3 * g1;
          const T* a1 = gaxis + 3 * g1;
          T v1[2][3] = {}, v2[2][3] = {};
          int ncand = 1;
          if (h1 == T(0)) {  // mjc_SphereSphere / mjc_SphereCapsule
            T d[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
            T x = clampT(dot3(a2, d), -h2, h2);
            for (int k = 0; k < 3; k++) { v1[0][k] = p1[k]; v2[0][k] = p2[k] + a2[k] * x; }
          } else {           // mjc_CapsuleCapsule
            T dif[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]}, cr[3];
            T ma = dot3(a1, a1), mb = -dot3(a1, a2), mc = dot3(a2, a2), u = -dot3(a1, dif), v = dot3(a2, dif);
            cross3(cr, a1, a2);
            T det = dot3(cr, cr);  // = ma*mc - mb*mb without the cancellation
            if (__builtin_expect(det >= B2H_MINVAL, 1)) {
              T x1 = (mc * u - mb * v) / det, x2 = (ma * v - mb * u) / det;
              if (x1 > h1) { x1 = h1; x2 = (v - mb * h1) / mc; }
              else if (x1 < -h1) { x1 = -h1; x2 = (v + mb * h1) / mc; }
              if (x2 > h2) { x2 = h2; x1 = clampT((u - mb * h2) / ma, -h1, h1); }
              else if (x2 < -h2) { x2 = -h2; x1 = clampT((u + mb * h2) / ma, -h1, h1); }
              for (int k = 0; k < 3; k++) { v1[0][k] = p1[k] + a1[k] * x1; v2[0][k] = p2[k] + a2[k] * x2; }
            } else {  // parallel axes: end points of 1 projected on 2, then of 2 on 1; at most two contacts
              ncand = 0;
              T mind = margin + r1 + r2;
              for (int e = 0; e < 4 && ncand < 2; e++) {
                T sgn = (e & 1) ? T(-1) : T(1), a[3], b[3], t[3];
                if (e < 2) {
                  for (int k = 0; k < 3; k++) { a[k] = p1[k] + a1[k] * h1 * sgn; t[k] = a[k] - p2[k]; }
                  T x2 = clampT(dot3(t, a2), -h2, h2);
                  for (int k = 0; k < 3; k++) b[k] = p2[k] + a2[k] * x2;
                } else {
                  for (int k = 0; k < 3; k++) { b[k] = p2[k] + a2[k] * h2 * sgn; t[k] = b[k] - p1[k]; }
                  T x1 = clampT(dot3(t, a1), -h1, h1);
                  for (int k = 0; k < 3; k++) a[k] = p1[k] + a1[k] * x1;
                }
                T d[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
                if (dot3(d, d) <= mind * mind) {
                  if (ncand == 0) { for (int k = 0; k < 3; k++) { v1[0][k] = a[k]; v2[0][k] = b[k]; } }
                  else { for (int k = 0; k < 3; k++) { v1[1][k] = a[k]; v2[1][k] = b[k]; } }
                  ncand++;
                }
              }
            }
          }
          T mind = margin + r1 + r2;
          // one copy of the body (a second candidate only exists for parallel axes): candidate 1 moves into slot 0
#pragma unroll 1
          for (int e = 0; e < ncand; e++) {
            T d[3] = {v2[0][0] - v1[0][0], v2[0][1] - v1[0][1], v2[0][2] - v1[0][2]};
            if (dot3(d, d) <= mind * mind) {
              T dist = normalize3(d) - r1 - r2;
              if (dist < margin) {
                if (n == 0) { cdist[0] = dist; for (int k = 0; k < 3; k++) { cnrm[0][k] = d[k]; cpos[0][k] = v1[0][k] + d[k] * (r1 + dist * T(0.5)); } }
                else { cdist[1] = dist; for (int k = 0; k < 3; k++) { cnrm[1][k] = d[k]; cpos[1][k] = v1[0][k] + d[k] * (r1 + dist * T(0.5)); } }
                n++;
              }
            }
            for (int k = 0; k < 3; k++) { v1[0][k] = v1[1][k]; v2[0][k] = v2[1][k]; }
          }
        }
      }
      int slot = ncon + wscan_excl(n, lane);
      int total = shfl(slot + n, 31) - ncon;
      // one copy of the body: the second contact of a pair (both ends of a capsule on the plane) moves into slot 0
#pragma unroll 1
      for (int e = 0; e < n; e++) {
        int c = slot + e;
        if (c < NCON) {
          T f[9];  // mju_makeFrame
          for (int k = 0; k < 3; k++) { f[k] = cnrm[0][k]; f[3 + k] = chint[k]; }
          normalize3(f);
          if (m_sqrt(dot3(f + 3, f + 3)) < T(0.5)) {
            f[3] = f[4] = f[5] = 0;
            if (f[1] < T(0.5) && f[1] > T(-0.5)) f[4] = 1; else f[5] = 1;
          }
          T dp = dot3(f, f + 3);
          for (int k = 0; k < 3; k++) f[3 + k] -= f[k] * dp;
          normalize3(f + 3);
          cross3(f + 6, f, f + 3);
          S.con_dist[c] = cdist[0];
          for (int k = 0; k < 3; k++) S.con_pos[3 * c + k] = cpos[0][k];
          for (int k = 0; k < 9; k++) S.con_frame[9 * c + k] = f[k];
          S.con_info[c] = info;
        }
        cdist[0] = cdist[1];
        for (int k = 0; k < 3; k++) { cnrm[0][k] = cnrm[1][k]; cpos[0][k] = cpos[1][k]; }
      }
      ncon += total;
    }
    if (ncon > NCON) { cnt.contact_overflow += ncon - NCON; ncon = NCON; }
  }
  wsync();

  if (dbg) {  // kinematics scratch is about to be overwritten by the constraint rows
    for (int i = lane; i < POS_END; i += 32) dbg->pos[i] = S.J[i];
    wsync();
  }
  B2H_CLK_ADD(16, tp);
  // =============================================================== constraint rows
  // dense rows: tendon limits first, then contacts (1 row frictionless, 4 rows pyramidal condim 3)
  int nrow = 0;
  T ten_pos[KT];   // signed distance of an active tendon limit (uniform)
  int ten_row[KT], ten_side[KT];
#pragma unroll
  for (int t = 0; t < KT; t++) {
    ten_row[t] = -1; ten_side[t] = 0; ten_pos[t] = 0;
    if (t >= B2H_LDG(m.ntendon)) continue;
    T len = wsum(lane < nq ? B2H_LDG(m.ten_qcoef[t][lane]) * st.qp : T(0));
    if (B2H_LDG(m.ten_limited[t])) {
      T lo = B2H_LDG(m.ten_range[t][0]), hi = B2H_LDG(m.ten_range[t][1]), mg = B2H_LDG(m.ten_margin[t]);
      if (len - lo < mg) { ten_side[t] = 1; ten_pos[t] = len - lo; }
      else if (hi - len < mg) { ten_side[t] = -1; ten_pos[t] = hi - len; }
      if (ten_side[t]) {
        ten_row[t] = nrow;
        if (lane < LD) jrow(nrow)[lane] = lane < nv ? T(ten_side[t]) * B2H_LDG(m.ten_J[t][lane]) : T(0);
        if (lane == 0) S.row_con[nrow] = -1 - t;
        nrow++;
      }
    }
  }
  {
    int nr = 0;
    if (lane < ncon) nr = B2H_LDG(m.cls_condim[S.con_info[lane] >> 16]) == 1 ? 1 : 4;
    int r0 = nrow + wscan_excl(nr, lane);
    bool drop = lane < ncon && r0 + nr > NROW;
    if (lane < ncon) S.con_row[lane] = drop ? -1 : r0;
    unsigned dm = ballot(drop);
    if (dm) cnt.contact_overflow += popc(dm);
    int keep_rows = drop ? 0 : nr;
    nrow += wsum(keep_rows);
  }
  if (NS * 32 < NROW && nrow > (NS == 1 && nrow_s < 32 ? nrow_s : NS * 32)) return B2H_STEP_MORE_ROWS;   // warp-uniform
  cnt.rows = (unsigned)nrow > cnt.rows ? (unsigned)nrow : cnt.rows;
  wsync();
  // contact Jacobians (lane = dof): J_k[d] = frame_k . (jacp_body2[d] - jacp_body1[d]), mj_jac about the com
  for (int c = 0; c < ncon; c++) {
    int r0 = S.con_row[c];
    if (r0 < 0) continue;
    uint32_t info = S.con_info[c];
    int b1 = info & 255, b2 = (info >> 8) & 255, cls = info >> 16;
    int sg = (int)((B2H_LDG(m.body_dofmask[b2]) >> lane) & 1u) - (int)((B2H_LDG(m.body_dofmask[b1]) >> lane) & 1u);
    T off[3], jp[3];
    for (int k = 0; k < 3; k++) off[k] = S.con_pos[3 * c + k] - com[k];
    cross3(jp, cd, off);
    for (int k = 0; k < 3; k++) jp[k] = (jp[k] + cd[3 + k]) * T(sg);
    const T* f = S.con_frame + 9 * c;
    T jn = dot3(f, jp);
    if (B2H_LDG(m.cls_condim[cls]) == 1) {
      if (lane < LD) jrow(r0)[lane] = jn;
      if (lane == 0) S.row_con[r0] = c;
    } else {
      T mu = B2H_LDG(m.cls_mu[cls]);
      T j1 = dot3(f + 3, jp) * mu, j2 = dot3(f + 6, jp) * mu;
      if (lane < LD) {
        jrow(r0 + 0)[lane] = jn + j1;
        jrow(r0 + 1)[lane] = jn - j1;
        jrow(r0 + 2)[lane] = jn + j2;
        jrow(r0 + 3)[lane] = jn - j2;
      }
      if (lane < 4) S.row_con[r0 + lane] = c;
    }
  }
  S.vec[1][lane] = lane < nv ? st.qv : T(0);
  wsync();

  B2H_CLK_ADD(17, tp);
  // per-row parameters; lane owns dense rows lane, lane+32, lane+64 (slots) and the joint limit of dof `lane`
  const int nslot = (nrow + 31) >> 5;
  T rD[NSLOT] = {}, raref[NSLOT] = {};
  B2H_SLOTS(s) {
    int r = lane + 32 * s;
    if (r < nrow) {
      T vel = 0;
      const T* jr = jrow(r);
#pragma unroll
      for (int c4 = 0; c4 < LD; c4 += 4) {
        V4<T> a = ld4(jr + c4), b = ld4(S.vec[1] + c4);
        vel += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
      }
      int rc = S.row_con[r];
      T sr0, sr1, si[5], pos, mg, dA, rs = 1;
      if (rc < 0) {
        int t = -1 - rc;
        sr0 = B2H_LDG(m.ten_solref[t][0]); sr1 = B2H_LDG(m.ten_solref[t][1]);
        for (int k = 0; k < 5; k++) si[k] = B2H_LDG(m.ten_solimp[t][k]);
        pos = 0;
#pragma unroll
        for (int k = 0; k < KT; k++) if (k == t) pos = ten_pos[k];
        mg = B2H_LDG(m.ten_margin[t]); dA = B2H_LDG(m.ten_invw[t]);
      } else {
        uint32_t info = S.con_info[rc];
        int cls = info >> 16;
        sr0 = B2H_LDG(m.cls_solref[cls][0]); sr1 = B2H_LDG(m.cls_solref[cls][1]);
        for (int k = 0; k < 5; k++) si[k] = B2H_LDG(m.cls_solimp[cls][k]);
        T tran = B2H_LDG(m.body_invw[info & 255]) + B2H_LDG(m.body_invw[(info >> 8) & 255]);
        T mu = B2H_LDG(m.cls_mu[cls]);
        bool pyr = B2H_LDG(m.cls_condim[cls]) != 1;
        // mj_diagApprox: tran (frictionless) or tran + mu^2 tran; pyramidal rows share Rpy = 2 mu^2 R
        pos = S.con_dist[rc]; mg = B2H_LDG(m.cls_incmargin[cls]);
        dA = pyr ? tran + mu * mu * tran : tran;
        rs = pyr ? T(2) * mu * mu : T(1);
      }
      RowParam<T> rp_ = row_params<T>(sr0, sr1, si[0], si[1], si[2], si[3], si[4], pos, mg, dA, vel, rs);
      rD[s] = rp_.D; raref[s] = rp_.aref;
    }
  }
  T lsign = 0, lD = 0, laref = 0;  // joint limit of this lane's dof: row = lsign * e_dof
  {
    int qa = lane < nv ? B2H_LDG(m.dof_qadr[lane]) : 0;
    T q = shfl(st.qp, qa);
    if (lane < nv && B2H_LDG(m.dof_limited[lane])) {
      T lo = B2H_LDG(m.dof_lo[lane]), hi = B2H_LDG(m.dof_hi[lane]), mg = B2H_LDG(m.dof_margin[lane]), pos = 0;
      if (q - lo < mg) { lsign = 1; pos = q - lo; }
      else if (hi - q < mg) { lsign = -1; pos = hi - q; }
      if (lsign != T(0)) {
        T si[5];
        for (int k = 0; k < 5; k++) si[k] = B2H_LDG(m.dof_solimp[lane][k]);
        RowParam<T> rp_ = row_params<T>(B2H_LDG(m.dof_solref[lane][0]), B2H_LDG(m.dof_solref[lane][1]), si[0], si[1], si[2], si[3],
                                        si[4], pos, mg, B2H_LDG(m.dof_invw[lane]), lsign * st.qv, T(1));
        lD = rp_.D; laref = rp_.aref;
      }
    }
  }
  const unsigned limit_mask = ballot(lsign != T(0));
  const int nefc = nrow + popc(limit_mask);
  cnt.nlim = (unsigned)popc(limit_mask) > cnt.nlim ? (unsigned)popc(limit_mask) : cnt.nlim;

  B2H_CLK_ADD(18, tp);
  // =============================================================== velocity stage
  // ---- mj_comVel as sums over ancestor chains: vprev[d] = sum of cdof*qvel over the dofs before d
  if (lane < KV) for (int k = 0; k < 6; k++) tmp[TMP_DOFW + 6 * lane + k] = cd[k] * (lane < nv ? st.qv : T(0));
  wsync();
  T cdd[6] = {0, 0, 0, 0, 0, 0};
  T vprev[6] = {0, 0, 0, 0, 0, 0};
  if (lane < nv) {
    for (int j = B2H_LDG(m.dof_vparent[lane]); j >= 0; j = B2H_LDG(m.dof_parent[j]))
      for (int k = 0; k < 6; k++) vprev[k] += tmp[TMP_DOFW + 6 * j + k];
    if (!B2H_LDG(m.dof_cdotzero[lane])) cross_motion(cdd, vprev, cd);
  }
  if (lane < KV) for (int k = 0; k < 6; k++) { tmp[TMP_CDD + 6 * lane + k] = cdd[k]; tmp[TMP_DOFA + 6 * lane + k] = cdd[k] * (lane < nv ? st.qv : T(0)); }
  wsync();
  // ---- cvel, cacc (lane = body) as chain sums; cfrc_body = I cacc + cvel x* (I cvel)   (mj_rne, flg_acc = 0)
  T bmom[3] = {0, 0, 0};
  if (lane > 0 && lane < nbody) {
    T cv[6] = {0, 0, 0, 0, 0, 0}, ca[6] = {0, 0, 0, 0, 0, 0};
    for (int j = B2H_LDG(m.body_lastdof[lane]); j >= 0; j = B2H_LDG(m.dof_parent[j]))
      for (int k = 0; k < 6; k++) { cv[k] += tmp[TMP_DOFW + 6 * j + k]; ca[k] += tmp[TMP_DOFA + 6 * j + k]; }
    for (int k = 0; k < 3; k++) ca[3 + k] -= B2H_LDG(m.gravity[k]);
    for (int k = 0; k < 6; k++) S.cvel[6 * lane + k] = cv[k];
    T f[6], t1[6], t2[6];
    mul_inert_vec(f, S.cinert + 10 * lane, ca);
    mul_inert_vec(t1, S.cinert + 10 * lane, cv);
    if constexpr (EXT) { bmom[0] = t1[3]; bmom[1] = t1[4]; bmom[2] = t1[5]; }   // the body's linear momentum
    cross_force(t2, cv, t1);
    for (int k = 0; k < 6; k++) tmp[TMP_CFRC + 6 * lane + k] = f[k] + t2[k];
  }
  wsync();
  // ---- qfrc_bias[d] = cdof[d] . sum of cfrc over the subtree of the dof's body
  T qfrc_bias = 0;
  if (lane < nv) {
    int end = B2H_LDG(m.body_subend[my_dbody]);
    T f[6] = {0, 0, 0, 0, 0, 0};
    for (int b = my_dbody; b < end; b++)
      for (int k = 0; k < 6; k++) f[k] += tmp[TMP_CFRC + 6 * b + k];
    qfrc_bias = cd[0] * f[0] + cd[1] * f[1] + cd[2] * f[2] + cd[3] * f[3] + cd[4] * f[4] + cd[5] * f[5];
  }
  // ---- mj_passive (joint springs, dampers), mj_fwdActuation (motors: gain 1, ctrl clamp, gear)
  T qfrc_smooth = 0;
  {
    int qa = lane < nv ? B2H_LDG(m.dof_qadr[lane]) : 0;
    T q = shfl(st.qp, qa);
    if (lane < nv) {
      T passive = -B2H_LDG(m.dof_damping[lane]) * st.qv - B2H_LDG(m.dof_stiff[lane]) * (q - B2H_LDG(m.dof_qspring[lane]));
      T act = 0;
      if (B2H_LDG(m.dof_act[lane]) >= 0) {
        T c = st.ctrl;
        if (B2H_LDG(m.act_limited[lane])) c = clampT(c, B2H_LDG(m.act_lo[lane]), B2H_LDG(m.act_hi[lane]));
        act = B2H_LDG(m.act_gear[lane]) * c;
      }
      st.qfrc_act = act;
      qfrc_smooth = passive - qfrc_bias + act;
    } else st.qfrc_act = 0;
  }
  if (dbg) {
    for (int i = lane; i < KV * 6; i += 32) dbg->cdofdot[i] = tmp[TMP_CDD + i];
    for (int i = lane; i < NCON; i += 32) dbg->con_dist[i] = S.con_dist[i];
    for (int i = lane; i < NCON * 3; i += 32) dbg->con_pos[i] = S.con_pos[i];
    for (int i = lane; i < NCON * 9; i += 32) dbg->con_frame[i] = S.con_frame[i];
  }
  wsync();  // stage scratch in A is dead from here
  B2H_CLK_ADD(19, tp);

  // =============================================================== acceleration: qacc_smooth = M^-1 qfrc_smooth
  B2H_CLK_ADD(0, tc);
  for (int i = lane; i < LD * LD; i += 32) S.A[i] = S.M[i];
  wsync();
  T qacc_smooth = chol_solve_fused(S.A, nv, lane, qfrc_smooth);
  B2H_CLK_ADD(1, tc);

#if defined(B2H_EXP_NEWTON_BARRIER) && !defined(B2H_HOST_EMU)
  // experiment: the lockstep group re-aligns before the Newton loop (the pre-solver stages drift with the contact count)
  if (cnt.sync_threads) asm volatile("bar.sync 1, %0;" ::"r"(cnt.sync_threads) : "memory");
#endif
  // =============================================================== mj_fwdConstraint: Newton solver (primal)
  T qacc = qacc_smooth, qfrc_con = 0;
  int niter = 0;
  T fin[NSLOT] = {};   // EXT: the dense rows' forces at the solution
  if (nefc > 0) {
    // J*x - aref for the dense rows of this lane and its limit row; x is read from S.vec[2]
    auto jar_of = [&](T x_lane, T* jar, T* ljar) {
      wsync();
      S.vec[2][lane] = lane < nv ? x_lane : T(0);
      wsync();
      B2H_SLOTS(s) {
        int r = lane + 32 * s;
        T a = 0;
        if (r < nrow) {
          const T* jr = jrow(r);
#pragma unroll
          for (int c4 = 0; c4 < LD; c4 += 4) {
            V4<T> u = ld4(jr + c4), w = ld4(S.vec[2] + c4);
            a += u.x * w.x + u.y * w.y + u.z * w.z + u.w * w.w;
          }
        }
        jar[s] = a;
      }
      *ljar = lsign * x_lane;
    };
    auto row_cost = [&](const T* jar, T ljar) {  // sum of 0.5 D r^2 over rows with r < 0
      T c = 0;
      B2H_SLOTS(s) { T r = jar[s] - raref[s]; if (lane + 32 * s < nrow && r < T(0)) c += T(0.5) * rD[s] * r * r; }
      T r = ljar - laref;
      if (lsign != T(0) && r < T(0)) c += T(0.5) * lD * r * r;
      return wsum(c);
    };
    // ---- warmstart(): the cheaper of qacc_warmstart and qacc_smooth; J*x and M*x of the chosen point are kept
    T jar[NSLOT] = {}, ljar, jar_s[NSLOT] = {}, ljar_s;
    jar_of(st.warm, jar, &ljar);
    T cost_warm = row_cost(jar, ljar);
    T Ma = mat_vec(S.M, S.vec[2], nv, lane);
    cost_warm += wsum(lane < nv ? T(0.5) * (Ma - qfrc_smooth) * (st.warm - qacc_smooth) : T(0));
    jar_of(qacc_smooth, jar_s, &ljar_s);
    T cost_smooth = row_cost(jar_s, ljar_s);
    if (cost_warm > cost_smooth) {  // warp-uniform
      qacc = qacc_smooth;
#pragma unroll
      for (int s = 0; s < NSLOT; s++) jar[s] = jar_s[s];
      ljar = ljar_s;
      Ma = mat_vec(S.M, S.vec[2], nv, lane);  // S.vec[2] still holds qacc_smooth
    } else { qacc = st.warm; }
    // ---- mj_solPrimal (Newton): state at the starting point
    T Jaref[NSLOT], lJaref;
#pragma unroll
    for (int s = 0; s < NSLOT; s++) Jaref[s] = jar[s] - raref[s];
    lJaref = ljar - laref;
    const T scale = T(1) / (B2H_LDG(m.meaninertia) * T(nv > 1 ? nv : 1));
    const T tolerance = T(1e-8), ls_tolerance = T(0.01);
    const int ls_iterations = 50;
    T cost = 0, gauss = 0, grad = 0, search = 0;
    bool first = true, full_step = false, tiny_step = false;
    unsigned pact[NSLOT] = {}, pactl = 0;  // active sets the current Hessian was built from
    for (;;) {
      // -- PrimalUpdateConstraint: active rows, forces, cost
      unsigned act[NSLOT] = {}, actl;
      T f[NSLOT] = {}, lf = 0, c = 0;
      B2H_SLOTS(s) {
        bool on = lane + 32 * s < nrow && Jaref[s] < T(0);
        if (on) { f[s] = -rD[s] * Jaref[s]; c += T(0.5) * rD[s] * Jaref[s] * Jaref[s]; }
        act[s] = ballot(on);
      }
      {
        bool on = lsign != T(0) && lJaref < T(0);
        if (on) { lf = -lD * lJaref; c += T(0.5) * lD * lJaref * lJaref; }
        actl = ballot(on);
      }
      if constexpr (EXT) {
#pragma unroll
        for (int s = 0; s < NSLOT; s++) fin[s] = f[s];
      }
      T oldcost = cost;
      gauss = wsum(lane < nv ? T(0.5) * (Ma - qfrc_smooth) * (qacc - qacc_smooth) : T(0));
      cost = wsum(c) + gauss;
      // qfrc_constraint = J^T f (dense active rows) + limit force
      qfrc_con = lsign * lf;
      B2H_SLOTS(s) {
        unsigned am = act[s];
        while (am) {
          int b = ffs32(am) - 1;
          am &= am - 1;
          T fr = shfl(f[s], b);
          if (lane < nv) qfrc_con += jrow(b + 32 * s)[lane] * fr;
        }
      }
      if (!first) {
        T improvement = scale * (oldcost - cost);
        T gn = m_sqrt(wsum(lane < nv ? (Ma - qfrc_smooth - qfrc_con) * (Ma - qfrc_smooth - qfrc_con) : T(0)));
        niter++;
        bool same_set = actl == pactl;
#pragma unroll
        for (int s = 0; s < NSLOT; s++) same_set = same_set && act[s] == pact[s];
#ifdef B2H_STAGE_CLOCKS
        if (improvement < m_max(tolerance, Tol<T>::cost_rel * scale * m_abs(cost))) { B2H_TALLY(24); B2H_TALLY(32 + (niter < 15 ? niter : 15)); if ((full_step || tiny_step) && same_set) B2H_TALLY(29); }
        else if (scale * gn < tolerance) { B2H_TALLY(25); B2H_TALLY(32 + (niter < 15 ? niter : 15)); }
        else if (Tol<T>::exact_stop && (full_step || tiny_step) && same_set) { B2H_TALLY(26); B2H_TALLY(32 + (niter < 15 ? niter : 15)); }
        if (same_set) B2H_TALLY(31);
#endif
        if (improvement < m_max(tolerance, Tol<T>::cost_rel * scale * m_abs(cost)) || scale * gn < tolerance) break;
        if (Tol<T>::exact_stop && (full_step || tiny_step) && same_set) break;
        if (niter >= Tol<T>::maxiter) { cnt.iter_cap++; break; }
      }
      first = false;
#pragma unroll
      for (int s = 0; s < NSLOT; s++) pact[s] = act[s];
      pactl = actl;
      B2H_CLK_ADD(5, tc);
      // -- Hessian H = M + J^T diag(D active) J, built per column (lane j owns column j), then factored
      {
        T acc[LD];
#pragma unroll
        for (int i = 0; i < LD; i++) acc[i] = S.M[i * LD + (lane < LD ? lane : 0)];
        B2H_SLOTS(s) {
          unsigned am = act[s];
          while (am) {
            int b = ffs32(am) - 1;
            am &= am - 1;
            const T* jr = jrow(b + 32 * s);
            T t = shfl(rD[s], b) * jr[lane < LD ? lane : 0];
#pragma unroll
            for (int c4 = 0; c4 < LD; c4 += 4) {
              V4<T> u = ld4(jr + c4);
              acc[c4] += u.x * t; acc[c4 + 1] += u.y * t; acc[c4 + 2] += u.z * t; acc[c4 + 3] += u.w * t;
            }
          }
        }
        wsync();
        if (lane < LD) {  // H is symmetric: the lane's column is its row (the factorisation reads the lower triangle)
#pragma unroll
          for (int c4 = 0; c4 < LD; c4 += 4) st4(S.A + lane * LD + c4, acc[c4], acc[c4 + 1], acc[c4 + 2], acc[c4 + 3]);
        }
        wsync();
        if ((actl >> lane) & 1u) S.A[lane * LD + lane] += lD;
        wsync();
      }
      // -- PrimalUpdateGradient + Newton direction: search = -H^-1 grad
      grad = lane < nv ? Ma - qfrc_smooth - qfrc_con : T(0);
      B2H_CLK_ADD(6, tc);
      search = -chol_solve_fused(S.A, nv, lane, grad);
      B2H_CLK_ADD(7, tc);
      // -- PrimalSearch: exact line search on the piecewise-quadratic cost along `search`
      T snorm = m_sqrt(wsum(lane < nv ? search * search : T(0)));
      T alpha = 0;
      T Jv[NSLOT] = {}, lJv = 0, Mv = 0;
      if (snorm >= B2H_MINVAL) {
        T gtol = tolerance * ls_tolerance * snorm / scale;
        jar_of(search, Jv, &lJv);
        Mv = mat_vec(S.M, S.vec[2], nv, lane);
        T qg1 = wsum(lane < nv ? search * (Ma - qfrc_smooth) : T(0));
        T qg2 = wsum(lane < nv ? T(0.5) * search * Mv : T(0));
        T q0r[NSLOT + 1] = {}, q1r[NSLOT + 1] = {}, q2r[NSLOT + 1] = {};  // per-row quadratics (slots + limit)
        B2H_SLOTS(s) {
          if (lane + 32 * s < nrow) {
            q0r[s] = T(0.5) * rD[s] * Jaref[s] * Jaref[s]; q1r[s] = rD[s] * Jaref[s] * Jv[s]; q2r[s] = T(0.5) * rD[s] * Jv[s] * Jv[s];
          }
        }
        if (lsign != T(0)) { q0r[NSLOT] = T(0.5) * lD * lJaref * lJaref; q1r[NSLOT] = lD * lJaref * lJv; q2r[NSLOT] = T(0.5) * lD * lJv * lJv; }
        int lsiter = 0;
        struct Pnt { T alpha, cost, d0, d1; };
        auto evalp = [&](T a) {
          T s0 = 0, s1 = 0, s2 = 0;
          B2H_SLOTS(s)
            if (lane + 32 * s < nrow && Jaref[s] + a * Jv[s] < T(0)) { s0 += q0r[s]; s1 += q1r[s]; s2 += q2r[s]; }
          if (lsign != T(0) && lJaref + a * lJv < T(0)) { s0 += q0r[NSLOT]; s1 += q1r[NSLOT]; s2 += q2r[NSLOT]; }
          s0 = wsum(s0) + gauss; s1 = wsum(s1) + qg1; s2 = wsum(s2) + qg2;
          Pnt p;
          p.alpha = a; p.cost = a * a * s2 + a * s1 + s0; p.d0 = T(2) * a * s2 + s1; p.d1 = T(2) * s2;
          if (p.d1 <= T(0)) p.d1 = B2H_MINVAL;
          return p;
        };
        auto eval = [&](T a) { lsiter++; return evalp(a); };
        Pnt p0 = eval(T(0));
        gtol = m_max(gtol, Tol<T>::ls_rel * m_abs(p0.d0));
        Pnt p1 = eval(p0.alpha - p0.d0 / p0.d1);
        full_step = !(p0.cost < p1.cost);
        if (p0.cost < p1.cost) p1 = p0;
        bool done = false;
        if (m_abs(p1.d0) < gtol) { alpha = p1.alpha; done = true; }
        else full_step = false;
        B2H_TALLY(27);                     // line searches
        if (full_step) B2H_TALLY(28);      // ... that took the exact Newton step at once
        if (done) B2H_TALLY(30);           // ... that ended after two evaluations
        if (!done) {
          int dir = p1.d0 < T(0) ? 1 : -1;
          bool p2update = false;
          Pnt p2 = p1;
          while (p1.d0 * T(dir) <= -gtol && lsiter < ls_iterations) {
            p2 = p1; p2update = true;
            p1 = eval(p1.alpha - p1.d0 / p1.d1);
            if (m_abs(p1.d0) < gtol) { alpha = p1.alpha; done = true; break; }
          }
          if (!done && (lsiter >= ls_iterations || !p2update)) { alpha = p1.alpha; done = true; }
          if (!done) {
            Pnt p2next = p1;
#if B2H_LS_FUSED
            // The bracketing phase as passes of independent evaluations: the midpoint of the NEXT round depends only on the
            // brackets, not on the two Newton points evaluated after the bracket update, so the three run side by side (their
            // butterfly sums interleave); a Newton point whose bracket did not move is evaluated again and dropped.  Same
            // evaluations, same order of the counted ones, same bits.
            Pnt p1next = evalp(p1.alpha - p1.d0 / p1.d1);
            Pnt pmid = evalp(T(0.5) * (p1.alpha + p2.alpha));
            lsiter++;
            while (lsiter < ls_iterations && !done) {
              lsiter++;                            // the midpoint is consumed
              Pnt cand[3] = {p1next, p2next, pmid};
              T bestcost = 0; int best = -1;
              for (int i = 0; i < 3; i++)
                if (m_abs(cand[i].d0) < gtol && (best == -1 || cand[i].cost < bestcost)) { bestcost = cand[i].cost; best = i; }
              if (best >= 0) { alpha = best == 0 ? cand[0].alpha : best == 1 ? cand[1].alpha : cand[2].alpha; done = true; break; }
              int b1 = 0, b2 = 0;
              for (int i = 0; i < 3; i++) {
                if (p1.d0 < T(0) && cand[i].d0 < T(0) && p1.d0 < cand[i].d0) { p1 = cand[i]; b1 = 1; }
                else if (p1.d0 > T(0) && cand[i].d0 > T(0) && p1.d0 > cand[i].d0) { p1 = cand[i]; b1 = 1; }
              }
              for (int i = 0; i < 3; i++) {
                if (p2.d0 < T(0) && cand[i].d0 < T(0) && p2.d0 < cand[i].d0) { p2 = cand[i]; b2 = 1; }
                else if (p2.d0 > T(0) && cand[i].d0 > T(0) && p2.d0 > cand[i].d0) { p2 = cand[i]; b2 = 1; }
              }
              if (!b1 && !b2) { alpha = pmid.cost < p0.cost ? pmid.alpha : T(0); done = true; break; }
              Pnt e1 = evalp(p1.alpha - p1.d0 / p1.d1);
              Pnt e2 = evalp(p2.alpha - p2.d0 / p2.d1);
              pmid = evalp(T(0.5) * (p1.alpha + p2.alpha));
              if (b1) { p1next = e1; lsiter++; }
              if (b2) { p2next = e2; lsiter++; }
            }
#else
            Pnt p1next = eval(p1.alpha - p1.d0 / p1.d1);
            while (lsiter < ls_iterations && !done) {
              Pnt pmid = eval(T(0.5) * (p1.alpha + p2.alpha));
              Pnt cand[3] = {p1next, p2next, pmid};
              T bestcost = 0; int best = -1;
              for (int i = 0; i < 3; i++)
                if (m_abs(cand[i].d0) < gtol && (best == -1 || cand[i].cost < bestcost)) { bestcost = cand[i].cost; best = i; }
              if (best >= 0) { alpha = best == 0 ? cand[0].alpha : best == 1 ? cand[1].alpha : cand[2].alpha; done = true; break; }
              int b1 = 0, b2 = 0;
              for (int i = 0; i < 3; i++) {
                if (p1.d0 < T(0) && cand[i].d0 < T(0) && p1.d0 < cand[i].d0) { p1 = cand[i]; b1 = 1; }
                else if (p1.d0 > T(0) && cand[i].d0 > T(0) && p1.d0 > cand[i].d0) { p1 = cand[i]; b1 = 1; }
              }
              if (b1) p1next = eval(p1.alpha - p1.d0 / p1.d1);
              for (int i = 0; i < 3; i++) {
                if (p2.d0 < T(0) && cand[i].d0 < T(0) && p2.d0 < cand[i].d0) { p2 = cand[i]; b2 = 1; }
                else if (p2.d0 > T(0) && cand[i].d0 > T(0) && p2.d0 > cand[i].d0) { p2 = cand[i]; b2 = 1; }
              }
              if (b2) p2next = eval(p2.alpha - p2.d0 / p2.d1);
              if (!b1 && !b2) { alpha = pmid.cost < p0.cost ? pmid.alpha : T(0); done = true; }
            }
#endif
            if (!done) {
              if (p1.cost <= p2.cost && p1.cost < p0.cost) alpha = p1.alpha;
              else if (p2.cost <= p1.cost && p2.cost < p0.cost) alpha = p2.alpha;
              else alpha = 0;
            }
          }
        }
        cnt.ls_eval += lsiter;
      }
      B2H_CLK_ADD(8, tc);
      if (alpha == T(0)) break;
      if (Tol<T>::step_rel > T(0)) {  // step below the resolution of qacc
        T amax = m_abs(qacc), smax = m_abs(alpha * search);
        for (int o = 16; o > 0; o >>= 1) { amax = m_max(amax, shfl_xor(amax, o)); smax = m_max(smax, shfl_xor(smax, o)); }
        tiny_step = smax < Tol<T>::step_rel * m_max(amax, T(1));
      }
      qacc += alpha * search; Ma += alpha * Mv;
#pragma unroll
      for (int s = 0; s < NSLOT; s++) Jaref[s] += alpha * Jv[s];
      lJaref += alpha * lJv;
    }
    B2H_CLK_ADD(5, tc);
    cnt.newton_iter += niter;
    cnt.work += niter * (8 + (nrow >> 2));
    st.warm = qacc;
  } else {
    st.warm = qacc_smooth;
  }
  if constexpr (EXT) {
    // ---- mj_subtreeVel, whole-model row: total linear momentum / total mass
    T lv[3];
    for (int k = 0; k < 3; k++) lv[k] = wsum(bmom[k]) * B2H_LDG(m.inv_total_mass);
    // ---- mj_rnePostConstraint, contact part (lane = body): the contact force in the world as a spatial force about the
    // tree's centre of mass, subtracted from body 1 (unless it is the world), added to body 2
    T ce[6] = {0, 0, 0, 0, 0, 0};
    auto row_force = [&](int r) {
      T v = 0;
#pragma unroll
      for (int s = 0; s < NSLOT; s++) { T t = shfl(fin[s], r & 31); if ((r >> 5) == s) v = t; }
      return v;
    };
    for (int c = 0; c < ncon; c++) {
      int r0 = S.con_row[c];
      if (r0 < 0) continue;
      uint32_t info = S.con_info[c];
      int b1 = info & 255, b2 = (info >> 8) & 255, cls = info >> 16;
      T lf[3] = {0, 0, 0};
      if (B2H_LDG(m.cls_condim[cls]) == 1) lf[0] = row_force(r0);
      else {   // mj_contactForce, pyramidal: normal = sum of the edge forces, tangents mu * (f0 - f1), mu * (f2 - f3)
        T f0 = row_force(r0), f1 = row_force(r0 + 1), f2 = row_force(r0 + 2), f3 = row_force(r0 + 3), mu = B2H_LDG(m.cls_mu[cls]);
        lf[0] = f0 + f1 + f2 + f3; lf[1] = mu * (f0 - f1); lf[2] = mu * (f2 - f3);
      }
      const T* fr = S.con_frame + 9 * c;
      T fw[3], off[3], tq[3];
      for (int k = 0; k < 3; k++) { fw[k] = fr[k] * lf[0] + fr[3 + k] * lf[1] + fr[6 + k] * lf[2]; off[k] = S.con_pos[3 * c + k] - com[k]; }
      cross3(tq, off, fw);
      T sg = T(lane == b2 && b2 != 0 ? 1 : 0) - T(lane == b1 && b1 != 0 ? 1 : 0);
      for (int k = 0; k < 3; k++) { ce[k] += sg * tq[k]; ce[3 + k] += sg * fw[k]; }
    }
    T l1 = m_abs(ce[0]) + m_abs(ce[1]) + m_abs(ce[2]) + m_abs(ce[3]) + m_abs(ce[4]) + m_abs(ce[5]);
    T lf_ = shfl(l1, nbody - 2), rf_ = shfl(l1, nbody - 1);
    wsync();
    if (lane == 0) { S.vec[0][0] = lf_; S.vec[0][1] = rf_; S.vec[0][2] = lv[0]; S.vec[0][3] = lv[1]; S.vec[0][4] = lv[2]; }
    if (dbg) {
      if (lane < KB) for (int k = 0; k < 6; k++) dbg->cfrc_ext[6 * lane + k] = ce[k];
      if (lane < 3) dbg->sub_linvel[lane] = lv[lane];
    }
    wsync();
  }
  if (stats) { stats->ncon = ncon; stats->nrow = nrow; stats->nlimit = popc(limit_mask); stats->niter = niter; }
  if (dbg) {
    T* dl = dbg->lane[lane];
    dl[0] = qfrc_bias; dl[1] = qfrc_smooth; dl[2] = qacc_smooth; dl[3] = qacc;
    dl[4] = qfrc_con; dl[5] = st.qfrc_act; dl[6] = lsign * lD; dl[7] = laref;
    for (int i = lane; i < LD * LD; i += 32) dbg->M[i] = S.M[i];
    for (int r = 0; r < nrow; r++) if (lane < LD) dbg->J[r * LD + lane] = jrow(r)[lane];
    for (int i = lane; i < KB * 10; i += 32) dbg->cinert[i] = S.cinert[i];
    for (int i = lane; i < KB * 6; i += 32) dbg->cvel[i] = S.cvel[i];
    for (int i = lane; i < KV * 6; i += 32) dbg->cdof[i] = S.cdof[i];
    if (lane < 3) dbg->com[lane] = S.com[lane];
  }
  if (qacc_out) *qacc_out = qacc;
  if (!integrate) return B2H_STEP_OK;

  // ---- mj_checkAcc: MuJoCo resets mjData and re-runs mj_forward before integrating
  if (ballot(lane < nv && is_bad(qacc))) {
    st.qp = lane < nq ? B2H_LDG(m.qpos0[lane]) : T(0);
    st.qv = 0; st.warm = 0; st.ctrl = 0; st.nstep = 0;
    cnt.bad_state++;
    return B2H_STEP_BAD_ACC;
  }

  // =============================================================== mj_Euler (implicit joint damping) + mj_advance
  B2H_CLK_ADD(2, tc);
  wsync();
  for (int i = lane; i < LD * LD; i += 32) S.A[i] = S.M[i];
  wsync();
  if (lane < nv) S.A[lane * LD + lane] += h * B2H_LDG(m.dof_damping[lane]);
  wsync();
  T qacc_e = chol_solve_fused(S.A, nv, lane, qfrc_smooth + qfrc_con);
  if (lane < nv) st.qv += h * qacc_e;
  {  // mj_integratePos: hinges and root position by lanes, root quaternion (lanes 3..6) via mju_quatIntegrate
    int dsrc = lane < nq ? B2H_LDG(m.qpos_dof[lane]) : -1;
    bool has_free = B2H_LDG(m.jnt_type[0]) == B2H_JNT_FREE;
    T v = shfl(st.qv, dsrc >= 0 ? dsrc : (lane < 3 ? lane : 0));
    T w[3], q[4];
    for (int k = 0; k < 3; k++) w[k] = shfl(st.qv, 3 + k);
    for (int k = 0; k < 4; k++) q[k] = shfl(st.qp, 3 + k);
    if (dsrc >= 0) st.qp += h * v;
    else if (has_free && lane < 3) st.qp += h * v;
    else if (has_free && lane < 7) {
      T ang = h * normalize3(w), qr[4], s, c;
      m_sincos(ang * T(0.5), &s, &c);
      if (ang == T(0)) { s = 0; c = 1; }
      qr[0] = c; qr[1] = w[0] * s; qr[2] = w[1] * s; qr[3] = w[2] * s;
      normalize4(q);
      T out[4];
      mul_quat(out, q, qr);
      st.qp = lane == 3 ? out[0] : lane == 4 ? out[1] : lane == 5 ? out[2] : out[3];
    }
  }
  st.nstep++;
  cnt.physics_steps++;
  wsync();
  B2H_CLK_ADD(3, tc);
  return B2H_STEP_OK;
}
// ns: slots to start with (warp-uniform; the step kernel makes it CTA-uniform from the rows the group's envs needed in
// their previous control step, so that a lockstep group walks ONE instantiation); too few slots fall back to all.
template <typename T>
B2H_DEV void mj_step(const DevModel<T>& m, Scratch<T>& S, T* Jspill, EnvState<T>& st, Counters& cnt, int ns = 1, bool ext = false) {
  for (int tries = 0; tries < 2; tries++) {
    int rc;
    if (tries) cnt.sync_threads = 0;   // a retry after mj_checkAcc already met the group once
    if (ext) rc = physics_step<T, false, NSLOT, true>(m, S, Jspill, st, cnt, true, nullptr, nullptr, nullptr);   // section 8 f4, not the bench path
    else {
      if (ns <= 1) rc = physics_step<T, false, 1>(m, S, Jspill, st, cnt, true, nullptr, nullptr, nullptr);
      else if (ns == 2) rc = physics_step<T, false, 2>(m, S, Jspill, st, cnt, true, nullptr, nullptr, nullptr);
      else rc = B2H_STEP_MORE_ROWS;
      if (rc == B2H_STEP_MORE_ROWS) rc = physics_step<T, false, NSLOT>(m, S, Jspill, st, cnt, true, nullptr, nullptr, nullptr);
    }
    if (rc == B2H_STEP_OK) break;
  }
}

// ------------------------------------------------------------------------------------------------ env layer
// counter-based reset noise: Philox4x32-10 keyed by the seed, counter = (global env id, episode, lane)
B2H_DEV void philox4x32(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
B2H_DEV double u53(uint32_t hi, uint32_t lo) {  // uniform double in [0,1) from 53 random bits
  return (double)((((uint64_t)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}

struct EnvParams {
  int frame_skip, reward_type, obs_mode, max_steps;
  double duration, timestep;
  double kneel[9];
  uint64_t seed;
  int env_id_offset;
  int sync_mode;  // CTA lockstep: 0 none, 1 once per control step, 2 before every physics sub-step
  int sensor_terms;  // B2HConfig::sensor_terms (section 8 f4): 0 = the reference (cfrc_ext / subtree_linvel read as zeros)
  int auto_reset;    // 1: SubprocVecEnv worker semantics (reset inside the step); 0: gymnasium Env.step (stay terminal)
};

template <typename T> B2H_DEV void quat_to_euler(const T* q, T* roll, T* pitch) {  // utils.py:3-20 (pitch unclamped)
  T w = q[0], x = q[1], y = q[2], z = q[3];
  *roll = m_atan2(T(2) * (w * x + y * z), T(1) - T(2) * (x * x + y * y));
  *pitch = m_asin(T(2) * (w * y - z * x));
}

// reward_functions.py: stand (:156-211), kneeling (:66-154), walk (:213-261); cfrc_ext and subtree_linvel are
// identically zero in the reference (no sensors), so the foot / com-velocity terms are the constants below.
template <typename T>
B2H_DEV_NOINLINE T compute_reward(const DevModel<T>& m, const Scratch<T>& S, const EnvState<T>& st, const EnvParams& P, int lane) {
  const int nv = B2H_LDG(m.nv);
  T hgt = shfl(st.qp, 2), vx = shfl(st.qv, 0), q[4], roll, pitch;
  for (int k = 0; k < 4; k++) q[k] = shfl(st.qp, 3 + k);
  T ctrl2 = wsum(lane < nv ? st.ctrl * st.ctrl : T(0));
  T power = wsum(lane >= 6 && lane < nv ? (st.qfrc_act * st.qv) * (st.qfrc_act * st.qv) : T(0));
  quat_to_euler(q, &roll, &pitch);
  // sum |cfrc_ext[-2]|, sum |cfrc_ext[-1]|, subtree_linvel[0]: zeros in the reference (no sensors), see physics_step<EXT>
  T lff = 0, rff = 0, lv2 = 0;
  if (P.sensor_terms) { lff = S.vec[0][0]; rff = S.vec[0][1]; lv2 = S.vec[0][2] * S.vec[0][2] + S.vec[0][3] * S.vec[0][3] + S.vec[0][4] * S.vec[0][4]; }
  if (P.reward_type == B2H_REWARD_STAND) {
    if (hgt < T(0.8)) return T(0);
    T vr = m_exp(T(-2) * (vx - T(1)) * (vx - T(1)));
    T hr = m_exp(T(-2) * (hgt - T(1.282)) * (hgt - T(1.282)));
    T orr = m_exp(T(-3) * (roll * roll + pitch * pitch));
    T foot = T(1) - m_min(lff, rff) / (lff + rff + T(1e-8));
    return T(0.4) * vr + T(0.3) * (T(0.5) * hr + T(0.5) * orr) + T(0.2) * foot + T(0.1) * m_exp(T(-0.05) * ctrl2);
  }
  if (P.reward_type == B2H_REWARD_WALK) {
    if (hgt < T(0.8)) return T(0.1) * hgt / T(0.8);
    T vr = m_exp(T(-0.5) * (vx - T(10)) * (vx - T(10)));
    T hr = m_exp(T(-2) * (hgt - T(1.282)) * (hgt - T(1.282)));
    T orr = m_exp(T(-3) * (roll * roll + pitch * pitch));
    return vr + (T(0.5) * hr + T(0.5) * orr) * m_exp(T(-0.05) * ctrl2);
  }
  const T th = T(P.kneel[0]), minh = T(P.kneel[1]), mrp = T(P.kneel[2]), crad = T(P.kneel[3]);
  if (hgt < minh) return hgt * hgt;
  T oerr = (roll * roll + pitch * pitch) / (mrp * mrp);
  T posture = T(0.7) * m_exp(T(-5) * oerr) + T(0.3) * m_exp(T(-5) * (hgt - th) * (hgt - th));
  T dist = m_sqrt(S.com[0] * S.com[0] + S.com[1] * S.com[1]);
  T com_score = T(0.7) * m_exp(T(-10) * (dist / crad)) + T(0.3) * m_exp(T(-0.1) * lv2);
  T foot_balance = m_min(lff, rff) / (lff + rff + T(1e-8));
  T energy = m_exp(T(-0.01) * power);
  T alive = T(1) - m_exp(T(-0.5) * T((double)st.nstep * P.timestep));
  return T(P.kneel[5]) * posture + T(P.kneel[6]) * com_score + T(P.kneel[7]) * foot_balance + T(P.kneel[4]) * energy +
         T(P.kneel[8]) * alive;
}

// custom_env.py:242-256: qpos[2:] | qvel | cinert | cvel | qfrc_actuator  (B2H_OBS_QPOS_QVEL = first two blocks)
template <typename T, typename O>  // O = T (device-resident rollouts) or double (the VecEnv's float64 observation)
B2H_DEV void write_obs(const DevModel<T>& m, const Scratch<T>& S, const EnvState<T>& st, int obs_mode, O* obs, int lane) {
  const int nq = B2H_LDG(m.nq), nv = B2H_LDG(m.nv), nbody = B2H_LDG(m.nbody);
  if (lane >= 2 && lane < nq) obs[lane - 2] = O(st.qp);
  if (lane < nv) obs[nq - 2 + lane] = O(st.qv);
  if (obs_mode == B2H_OBS_QPOS_QVEL) return;
  int o = nq - 2 + nv;
  for (int i = lane; i < 10 * nbody; i += 32) obs[o + i] = O(S.cinert[i]);
  o += 10 * nbody;
  for (int i = lane; i < 6 * nbody; i += 32) obs[o + i] = O(S.cvel[i]);
  o += 6 * nbody;
  if (lane < nv) obs[o + lane] = O(st.qfrc_act);
}

template <typename T>
struct EnvIO {  // device arrays, all [n_envs, dim] row-major
  T *qpos, *qvel, *warm;
  int *nstep, *step_count, *episode;
  T* total_reward;
  double* reset_noise;      // [n_envs, nq+nv] noise of the most recent reset (and injected noise for the next one)
  uint8_t* noise_injected;  // 1: take reset_noise[e] as is for the next reset (parity hook)
  const float* actions;     // [n_envs, nu]
  T *obs, *reward, *terminal_obs;
  uint8_t *terminated, *truncated;
  int obs_dim;
  // float64 outputs of the VecEnv boundary (may be page-locked host memory written straight over PCIe), or null
  double *obs64, *reward64, *terminal_obs64;
  int* work;                // [n_envs] solver effort of the last control step (schedule key of the next launch), or null
};

// HumanoidEnv.reset (custom_env.py:97-150): qpos0 + masked U(-0.01,0.01) noise, one settle step with ctrl = 0
template <typename T>
B2H_DEV_NOINLINE void env_reset(const DevModel<T>& m, Scratch<T>& S, T* Jspill, EnvState<T>& st, Counters& cnt, const EnvParams& P,
                       const EnvIO<T>& io, int env, int lane) {
  const int nq = B2H_LDG(m.nq), nv = B2H_LDG(m.nv);
  const int nqv = nq + nv;
  double npos = 0, nvel = 0;
  if (io.noise_injected[env]) {
    if (lane < nq) npos = io.reset_noise[(size_t)env * nqv + lane];
    if (lane < nv) nvel = io.reset_noise[(size_t)env * nqv + nq + lane];
  } else {
    uint32_t c[4] = {(uint32_t)(env + P.env_id_offset), (uint32_t)io.episode[env], (uint32_t)lane, 0x6232683Fu};
    philox4x32(c, (uint32_t)P.seed, (uint32_t)(P.seed >> 32));
    npos = -0.01 + 0.02 * u53(c[0], c[1]);
    nvel = -0.01 + 0.02 * u53(c[2], c[3]);
    if (lane < nq) io.reset_noise[(size_t)env * nqv + lane] = npos;
    if (lane < nv) io.reset_noise[(size_t)env * nqv + nq + lane] = nvel;
  }
  wsync();
  if (lane == 0) { io.noise_injected[env] = 0; io.episode[env] += 1; }
  if (lane == 2) npos *= 0.1;
  if (lane >= 3 && lane < 7) npos = 0;
  double q0 = lane < nq ? (double)B2H_LDG(m.qpos0[lane]) : 0.0;
  if (lane == 2) q0 = 1.282;                       // init_qpos, custom_env.py:58-61
  if (lane >= 3 && lane < 7) q0 = lane == 3 ? 1.0 : 0.0;
  st.qp = lane < nq ? T(q0 + npos) : T(0);
  st.qv = lane < nv ? T(nvel) : T(0);
  st.warm = 0; st.ctrl = 0; st.qfrc_act = 0; st.nstep = 0;
  cnt.sync_threads = 0;   // only the envs that finished run this step
  mj_step<T>(m, S, Jspill, st, cnt, 1, P.sensor_terms != 0);
}

// HumanoidEnv.step + SubprocVecEnv auto-reset for one env (custom_env.py:152-230; SB3 subproc_vec_env._worker)
// OUT: which result arrays this instantiation writes -- 0 the arithmetic-type ones (io.obs, ...), 1 the float64
// VecEnv ones (io.obs64, ...), -1 whichever are non-null (the unused variants then sit in the claim loop's stream).
template <typename T, int OUT = -1>
B2H_DEV void env_step(const DevModel<T>& m, Scratch<T>& S, T* Jspill, Counters& cnt, const EnvParams& P, const EnvIO<T>& io, int env,
                      bool active, int ns = 1 /* row slots the lockstep group starts with, see mj_step */, int group_threads = 0) {
  constexpr bool kOutT = OUT != 1, kOut64 = OUT != 0;
  // All warps of a CTA enter every sub-step together (cta_sync): they then walk the same instructions at about
  // the same time, which keeps the (large, mostly straight-line) step code resident in the instruction cache.
  // `active` is warp-uniform; inactive warps (tail of the env list) only take part in the barriers.
  const int lane = lane_id();
  const int nq = B2H_LDG(m.nq), nv = B2H_LDG(m.nv), nu = B2H_LDG(m.nu);
  EnvState<T> st;
  st.qp = 0; st.qv = 0; st.warm = 0; st.ctrl = 0; st.qfrc_act = 0; st.nstep = 0;
  int a = -1, step_count = 0;
  T action = 0;   // this lane's motor command: read once per control step (the array may be page-locked host memory)
  if (active) {
    st.qp = lane < nq ? io.qpos[(size_t)env * nq + lane] : T(0);
    st.qv = lane < nv ? io.qvel[(size_t)env * nv + lane] : T(0);
    st.warm = lane < nv ? io.warm[(size_t)env * nv + lane] : T(0);
    st.nstep = io.nstep[env];
    a = lane < nv ? B2H_LDG(m.dof_act[lane]) : -1;
    if (a >= 0) action = T(io.actions[(size_t)env * nu + a]);
    step_count = io.step_count[env] + 1;
  }
  cnt.work = 0; cnt.rows = 0; cnt.nlim = 0;
  B2H_CLK(te);
  if (P.sync_mode == 1) cta_sync();
  for (int s = 0; s < P.frame_skip; s++) {
    B2H_CLK_ADD(9, te);
    if (P.sync_mode == 2) cta_sync();
    B2H_CLK_ADD(4, te);
    if (active) {
      // data.ctrl[:] = action before every mj_step (a bad-state reset inside the previous sub-step zeroed it)
      st.ctrl = action;
      cnt.sync_threads = group_threads;
      mj_step<T>(m, S, Jspill, st, cnt, ns, P.sensor_terms != 0);
    }
  }
  bool done = false;
  T total = 0;
  if (active) {
    bool truncated = step_count >= P.max_steps;
    T reward = truncated ? T(0) : compute_reward<T>(m, S, st, P, lane);
    // terminated = data.time >= duration, with time = nstep * timestep evaluated in double
    bool terminated = (double)st.nstep * P.timestep >= P.duration;
    total = io.total_reward[env] + reward;
    if (lane == 0) {
      if (kOutT && io.reward) io.reward[env] = reward;
      if (kOut64 && io.reward64) io.reward64[env] = (double)reward;
      io.terminated[env] = terminated; io.truncated[env] = truncated;
    }
    done = terminated || truncated;
    if (kOutT && done && io.terminal_obs) write_obs(m, S, st, P.obs_mode, io.terminal_obs + (size_t)env * io.obs_dim, lane);
    if (kOut64 && done && io.terminal_obs64) write_obs(m, S, st, P.obs_mode, io.terminal_obs64 + (size_t)env * io.obs_dim, lane);
  }
  B2H_CLK_ADD(9, te);
  if (P.sync_mode == 2) cta_sync();
  B2H_CLK_ADD(4, te);
  if (active && done && P.auto_reset) {
    env_reset<T>(m, S, Jspill, st, cnt, P, io, env, lane);
    step_count = 0; total = 0;
  }
  if (active) {
    if (kOutT && io.obs) write_obs(m, S, st, P.obs_mode, io.obs + (size_t)env * io.obs_dim, lane);
    if (kOut64 && io.obs64) write_obs(m, S, st, P.obs_mode, io.obs64 + (size_t)env * io.obs_dim, lane);
    if (lane < nq) io.qpos[(size_t)env * nq + lane] = st.qp;
    if (lane < nv) { io.qvel[(size_t)env * nv + lane] = st.qv; io.warm[(size_t)env * nv + lane] = st.warm; }
    if (lane == 0) {
      io.nstep[env] = st.nstep; io.step_count[env] = step_count; io.total_reward[env] = total;
      // effort (14 bits) | joint limits (6 bits) | dense rows (12 bits)
      if (io.work) io.work[env] = (int)((cnt.work < 16383u ? cnt.work : 16383u) | ((cnt.nlim & 63u) << 14) | (cnt.rows << B2H_EFFORT_BITS));
    }
  }
  B2H_CLK_ADD(9, te);
}

// reset path on its own (b2h_reset): reset env, write the first observation
template <typename T>
B2H_DEV void env_reset_only(const DevModel<T>& m, Scratch<T>& S, T* Jspill, Counters& cnt, const EnvParams& P, const EnvIO<T>& io, int env) {
  const int lane = lane_id();
  const int nq = B2H_LDG(m.nq), nv = B2H_LDG(m.nv);
  EnvState<T> st;
  env_reset<T>(m, S, Jspill, st, cnt, P, io, env, lane);
  if (io.obs) write_obs(m, S, st, P.obs_mode, io.obs + (size_t)env * io.obs_dim, lane);
  if (io.obs64) write_obs(m, S, st, P.obs_mode, io.obs64 + (size_t)env * io.obs_dim, lane);
  if (lane < nq) io.qpos[(size_t)env * nq + lane] = st.qp;
  if (lane < nv) { io.qvel[(size_t)env * nv + lane] = st.qv; io.warm[(size_t)env * nv + lane] = st.warm; }
  if (lane == 0) { io.nstep[env] = st.nstep; io.step_count[env] = 0; io.total_reward[env] = 0; }
}

}  // namespace b2h
