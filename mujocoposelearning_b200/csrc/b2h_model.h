// b2h_model.h — device-side model tables (DevModel<T>) and their host-side construction from B2HModel.
//
// B2HModel (include/b2h.h) is the reference-shaped, double-precision compiled model (mjModel fields read on
// the path: custom_env.py:53).  DevModel<T> is the kernel's view: arithmetic type T, per-lane lookup tables
// (body level / subtree ranges / dof ancestor masks), contact-parameter classes, one packed word per collision
// candidate.  Pure host C++ (no CUDA) so the CPU lane-emulation test build can share it.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <string>

#include "../../include/b2h.h"

namespace b2h {

constexpr int KB = 20;    // bodies
constexpr int KJ = 24;    // joints
constexpr int KV = 28;    // dofs (nv <= 27; index 27 is the zero pad column of LD-strided matrices)
constexpr int KQ = 32;    // qpos entries (one lane each)
constexpr int KG = 24;    // geoms
constexpr int KT = 2;     // fixed tendons
constexpr int KCLS = 8;   // contact parameter classes
constexpr int KPAIR = 512;
constexpr int LD = 28;    // leading dimension of J / M / A rows (multiple of 4 for 128-bit shared loads)
constexpr int NSLOT = 3;          // dense rows per lane
constexpr int NROW = 32 * NSLOT;  // dense constraint rows per env (tendon limits + contact rows)
#ifndef B2H_NROW_S
#define B2H_NROW_S 48
#endif
constexpr int NROW_S = B2H_NROW_S;       // most rows a warp keeps in shared memory (DevModel::nrow_s is the launch's choice); rows beyond live in a per-warp global spill area
constexpr int NCON = 32;          // contacts per env

template <typename T>
struct DevModel {
  int nq, nv, nu, nbody, njnt, ngeom, ntendon, npair, nlevel, ncls, maxsub;
  int nrow_s;  // dense constraint rows kept in shared memory by this launch configuration (<= NROW_S)
  T timestep, gravity[3], meaninertia, inv_total_mass;
  // bodies
  int body_parent[KB], body_level[KB], body_jntadr[KB], body_jntnum[KB], body_subend[KB], body_isfree[KB], body_lastdof[KB];
  uint32_t body_dofmask[KB];  // bit d set iff dof d moves body b
  T body_pos[KB][3], body_quat[KB][4], body_ipos[KB][3], body_inertia[KB][6], body_mass[KB], body_invw[KB];
  // joints
  int jnt_type[KJ], jnt_body[KJ], jnt_qadr[KJ], jnt_dadr[KJ];
  T jnt_pos[KJ][3], jnt_axis[KJ][3], jnt_q0[KJ];
  // dofs (hinge-only fields are zero / -1 for free-joint dofs)
  int dof_body[KV], dof_jnt[KV], dof_parent[KV], dof_vparent[KV], dof_cdotzero[KV], dof_qadr[KV], dof_act[KV], dof_limited[KV];
  T dof_armature[KV], dof_damping[KV], dof_invw[KV], dof_stiff[KV], dof_qspring[KV], dof_lo[KV], dof_hi[KV], dof_margin[KV];
  T dof_solref[KV][2], dof_solimp[KV][5];
  int qpos_dof[KQ];  // hinge qpos index -> dof, -1 for free-joint coordinates
  T qpos0[KQ];
  // geoms
  int geom_type[KG], geom_body[KG];
  T geom_size[KG][2], geom_pos[KG][3], geom_zaxis[KG][3];
  // collision candidates: g1 | g2 << 8 | class << 16
  uint32_t pair[KPAIR];
  int cls_condim[KCLS];
  T cls_mu[KCLS], cls_solref[KCLS][2], cls_solimp[KCLS][5], cls_margin[KCLS], cls_incmargin[KCLS];
  // fixed tendons
  int ten_limited[KT];
  T ten_J[KT][KV], ten_qcoef[KT][KQ], ten_range[KT][2], ten_solref[KT][2], ten_solimp[KT][5], ten_margin[KT], ten_invw[KT];
  // motors (at most one per dof)
  T act_gear[KV], act_lo[KV], act_hi[KV];
  int act_limited[KV];
};

inline void quat_to_mat_d(const double* q, double* R) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = w * w - x * x - y * y + z * z;
}

// getsolparam clamps (engine_core_constraint.c): refsafe, impedance range, width >= 0, midpoint range, power >= 1
template <typename T>
inline void clamp_sol(const double* solref, const double* solimp, double timestep, T* oref, T* oimp) {
  double r0 = solref[0], r1 = solref[1];
  if (r0 > 0 && r0 < 2 * timestep) r0 = 2 * timestep;
  oref[0] = (T)r0; oref[1] = (T)r1;
  auto clip = [](double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); };
  oimp[0] = (T)clip(solimp[0], 0.0001, 0.9999);
  oimp[1] = (T)clip(solimp[1], 0.0001, 0.9999);
  oimp[2] = (T)(solimp[2] < 0 ? 0 : solimp[2]);
  oimp[3] = (T)clip(solimp[3], 0.0001, 0.9999);
  oimp[4] = (T)(solimp[4] < 1 ? 1 : solimp[4]);
}

// Returns "" on success, else why this model is outside what the kernels support.
template <typename T>
std::string build_dev_model(const B2HModel& m, DevModel<T>& d) {
  memset(&d, 0, sizeof(d));
  if (m.nbody > KB || m.njnt > KJ || m.nv >= KV || m.nq > KQ || m.ngeom > KG || m.ntendon > KT || m.npair > KPAIR ||
      m.nu > KV || m.nbody < 2)
    return "model exceeds kernel capacities (bodies/joints/dofs/geoms/tendons/pairs)";
  d.nrow_s = NROW_S;
  d.nq = m.nq; d.nv = m.nv; d.nu = m.nu; d.nbody = m.nbody; d.njnt = m.njnt; d.ngeom = m.ngeom;
  d.ntendon = m.ntendon; d.npair = m.npair;
  d.timestep = (T)m.timestep; d.meaninertia = (T)m.meaninertia;
  for (int k = 0; k < 3; k++) d.gravity[k] = (T)m.gravity[k];
  double tm = 0;
  int nroot = 0;
  for (int b = 1; b < m.nbody; b++) {
    tm += m.body_mass[b];
    if (m.body_parentid[b] == 0) nroot++;
    if (m.body_parentid[b] >= b) return "bodies must be numbered parent-before-child";
  }
  if (nroot != 1) return "exactly one kinematic tree is supported";
  d.inv_total_mass = (T)(1.0 / tm);
  for (int b = 0; b < m.nbody; b++) {
    d.body_parent[b] = m.body_parentid[b];
    d.body_level[b] = b == 0 ? 0 : d.body_level[m.body_parentid[b]] + 1;
    if (d.body_level[b] > d.nlevel) d.nlevel = d.body_level[b];
    d.body_jntadr[b] = m.body_jntadr[b]; d.body_jntnum[b] = m.body_jntnum[b]; d.body_lastdof[b] = m.body_lastdof[b];
    d.body_isfree[b] = m.body_jntnum[b] == 1 && m.jnt_type[m.body_jntadr[b]] == B2H_JNT_FREE;
    if (d.body_isfree[b] && (m.body_parentid[b] != 0 || m.jnt_qposadr[m.body_jntadr[b]] != 0))
      return "a free joint must belong to the root body and come first in qpos";
    // depth-first numbering: subtree of b is the contiguous id range [b, subend)
    int e = b + 1;
    while (e < m.nbody) {
      int a = e;
      while (a > b) a = m.body_parentid[a];
      if (a != b) break;
      e++;
    }
    d.body_subend[b] = e;
    if (b > 0 && e - b > d.maxsub) d.maxsub = e - b;
    for (int a = e; a < m.nbody; a++) {  // later ids must not be descendants
      int p = a;
      while (p > b) p = m.body_parentid[p];
      if (p == b && b != 0) return "bodies must be numbered depth-first";
    }
    uint32_t mask = 0;
    for (int dd = m.body_lastdof[b]; dd >= 0; dd = m.dof_parentid[dd]) mask |= 1u << dd;
    d.body_dofmask[b] = mask;
    for (int k = 0; k < 3; k++) { d.body_pos[b][k] = (T)m.body_pos[b][k]; d.body_ipos[b][k] = (T)m.body_ipos[b][k]; }
    for (int k = 0; k < 4; k++) d.body_quat[b][k] = (T)m.body_quat[b][k];
    for (int k = 0; k < 6; k++) d.body_inertia[b][k] = (T)m.body_inertia_full[b][k];
    d.body_mass[b] = (T)m.body_mass[b];
    // mj_diagApprox (engine_core_constraint.c, contact cases) indexes body_invweight0 with geom_bodyid itself, not with
    // the weld parent: a jointless child (head, hands) carries its own value  [UNVERIFIED-vs-3.2.5, see oracle]
    d.body_invw[b] = (T)m.body_invweight0[b][0];
  }
  for (int k = 0; k < KQ; k++) d.qpos_dof[k] = -1;
  for (int i = 0; i < KV; i++) { d.dof_parent[i] = -1; d.dof_vparent[i] = -1; d.dof_act[i] = -1; d.dof_qadr[i] = 0; d.dof_cdotzero[i] = 1; }
  for (int j = 0; j < m.njnt; j++) {
    d.jnt_type[j] = m.jnt_type[j]; d.jnt_body[j] = m.jnt_bodyid[j]; d.jnt_qadr[j] = m.jnt_qposadr[j]; d.jnt_dadr[j] = m.jnt_dofadr[j];
    for (int k = 0; k < 3; k++) { d.jnt_pos[j][k] = (T)m.jnt_pos[j][k]; d.jnt_axis[j][k] = (T)m.jnt_axis[j][k]; }
    d.jnt_q0[j] = (T)m.qpos0[m.jnt_qposadr[j]];
    if (m.jnt_type[j] == B2H_JNT_FREE) {
      if (j != 0) return "only the first joint may be free";
    } else if (m.jnt_type[j] == B2H_JNT_HINGE) {
      int dd = m.jnt_dofadr[j];
      d.qpos_dof[m.jnt_qposadr[j]] = dd;
      d.dof_qadr[dd] = m.jnt_qposadr[j];
      d.dof_stiff[dd] = (T)m.jnt_stiffness[j];
      d.dof_qspring[dd] = (T)m.qpos_spring[m.jnt_qposadr[j]];
      d.dof_limited[dd] = m.jnt_limited[j];
      d.dof_lo[dd] = (T)m.jnt_range[j][0]; d.dof_hi[dd] = (T)m.jnt_range[j][1];
      d.dof_margin[dd] = (T)m.jnt_margin[j];
      if (m.jnt_limited[j] && !(m.jnt_range[j][1] - m.jnt_range[j][0] > 2 * m.jnt_margin[j]))
        return "joint range must exceed twice its margin (one active limit side per joint)";
      clamp_sol<T>(m.jnt_solref[j], m.jnt_solimp[j], m.timestep, d.dof_solref[dd], d.dof_solimp[dd]);
    } else {
      return "only free and hinge joints are supported";
    }
  }
  for (int i = 0; i < m.nv; i++) {
    d.dof_body[i] = m.dof_bodyid[i]; d.dof_jnt[i] = m.dof_jntid[i]; d.dof_parent[i] = m.dof_parentid[i];
    d.dof_armature[i] = (T)m.dof_armature[i]; d.dof_damping[i] = (T)m.dof_damping[i]; d.dof_invw[i] = (T)m.dof_invweight0[i];
    int j = m.dof_jntid[i];
    if (m.jnt_type[j] == B2H_JNT_FREE) {
      int k = i - m.jnt_dofadr[j];
      // mj_comVel: translational cdof_dot = 0; the three rotational ones all see cvel after translation only
      d.dof_cdotzero[i] = k < 3;
      d.dof_vparent[i] = k < 3 ? -1 : m.jnt_dofadr[j] + 2;
    } else {
      d.dof_cdotzero[i] = 0;
      d.dof_vparent[i] = m.dof_parentid[i];
    }
  }
  for (int k = 0; k < m.nq; k++) d.qpos0[k] = (T)m.qpos0[k];
  for (int g = 0; g < m.ngeom; g++) {
    d.geom_type[g] = m.geom_type[g]; d.geom_body[g] = m.geom_bodyid[g];
    d.geom_size[g][0] = (T)m.geom_size[g][0]; d.geom_size[g][1] = (T)m.geom_size[g][1];
    double R[9];
    quat_to_mat_d(m.geom_quat[g], R);
    for (int k = 0; k < 3; k++) { d.geom_pos[g][k] = (T)m.geom_pos[g][k]; d.geom_zaxis[g][k] = (T)R[3 * k + 2]; }
    if (m.geom_type[g] == B2H_GEOM_SPHERE) d.geom_size[g][1] = 0;
    if (m.geom_type[g] != B2H_GEOM_PLANE && m.geom_type[g] != B2H_GEOM_SPHERE && m.geom_type[g] != B2H_GEOM_CAPSULE)
      return "only plane / sphere / capsule geoms are supported";
  }
  // contact parameter classes: dedupe (condim, mu, solref, solimp, margin, gap)
  double cls[KCLS][12];
  for (int p = 0; p < m.npair; p++) {
    double key[12] = {(double)m.pair_condim[p], m.pair_friction[p][0], m.pair_solref[p][0], m.pair_solref[p][1],
                      m.pair_solimp[p][0], m.pair_solimp[p][1], m.pair_solimp[p][2], m.pair_solimp[p][3], m.pair_solimp[p][4],
                      m.pair_margin[p], m.pair_gap[p], 0};
    int c = -1;
    for (int k = 0; k < d.ncls; k++) if (!memcmp(cls[k], key, sizeof key)) c = k;
    if (c < 0) {
      if (d.ncls == KCLS) return "too many distinct contact parameter classes";
      c = d.ncls++;
      memcpy(cls[c], key, sizeof key);
      if (m.pair_condim[p] != 1 && m.pair_condim[p] != 3) return "only condim 1 and 3 are supported";
      d.cls_condim[c] = m.pair_condim[p];
      d.cls_mu[c] = (T)m.pair_friction[p][0];
      clamp_sol<T>(m.pair_solref[p], m.pair_solimp[p], m.timestep, d.cls_solref[c], d.cls_solimp[c]);
      d.cls_margin[c] = (T)m.pair_margin[p];
      d.cls_incmargin[c] = (T)(m.pair_margin[p] - m.pair_gap[p]);
    }
    int g1 = m.pair_geom1[p], g2 = m.pair_geom2[p];
    if (m.geom_type[g1] > m.geom_type[g2]) return "pair geoms must be ordered by type";
    if (m.geom_type[g2] == B2H_GEOM_PLANE) return "plane-plane pairs are not supported";
    d.pair[p] = (uint32_t)g1 | ((uint32_t)g2 << 8) | ((uint32_t)c << 16);
  }
  for (int t = 0; t < m.ntendon; t++) {
    d.ten_limited[t] = m.ten_limited[t];
    for (int i = 0; i < m.nv; i++) d.ten_J[t][i] = (T)m.ten_J[t][i];
    for (int k = 0; k < m.nq; k++) d.ten_qcoef[t][k] = (T)m.ten_qcoef[t][k];
    d.ten_range[t][0] = (T)m.ten_range[t][0]; d.ten_range[t][1] = (T)m.ten_range[t][1];
    clamp_sol<T>(m.ten_solref[t], m.ten_solimp[t], m.timestep, d.ten_solref[t], d.ten_solimp[t]);
    d.ten_margin[t] = (T)m.ten_margin[t]; d.ten_invw[t] = (T)m.ten_invweight0[t];
    if (m.ten_limited[t] && !(m.ten_range[t][1] - m.ten_range[t][0] > 2 * m.ten_margin[t]))
      return "tendon range must exceed twice its margin";
  }
  for (int a = 0; a < m.nu; a++) {
    int dd = m.actuator_dofid[a];
    if (dd < 0 || dd >= m.nv || d.dof_act[dd] >= 0) return "motors must drive distinct hinge dofs";
    d.dof_act[dd] = a;
    d.act_gear[dd] = (T)m.actuator_gear[a];
    d.act_limited[dd] = m.actuator_ctrllimited[a];
    d.act_lo[dd] = (T)m.actuator_ctrlrange[a][0]; d.act_hi[dd] = (T)m.actuator_ctrlrange[a][1];
  }
  return "";
}

}  // namespace b2h
