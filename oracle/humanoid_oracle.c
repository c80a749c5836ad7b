/* humanoid_oracle.c — TEST INFRASTRUCTURE ONLY (CPU fp64 oracle of the rollout hot path).
 *
 * PARITY UNPINNED: the arithmetic of the reference's hot path lives in third-party, un-vendored
 * dependencies (mujoco==3.2.5, stable-baselines3==2.3.2; reference environment.yml:152,204-205,289) that are
 * not installable in this image, and the reference ships no tests or golden vectors.  This file restates
 * the published MuJoCo 3.2.5 algorithms (engine_forward.c, engine_core_smooth.c, engine_core_constraint.c,
 * engine_collision_primitive.c, engine_solver.c, engine_passive.c, engine_util_*.c) for the MJCF subset the
 * reference model uses, anchored on the reference's own call sites:
 *     mujoco.mj_step            custom_env.py:121,160     -> orc_mj_step
 *     mujoco.mj_resetData       custom_env.py:102         -> orc_reset_data
 *     HumanoidEnv.reset         custom_env.py:97-150      -> orc_env_reset
 *     HumanoidEnv.step          custom_env.py:152-230     -> orc_env_step
 *     HumanoidEnv._get_state    custom_env.py:232-261     -> orc_obs
 *     stand_reward              reward_functions.py:156-211, robust_kneeling_reward :66-154,
 *     walk_reward :213-261, quaternion_to_euler utils.py:3-20
 *     SubprocVecEnv worker auto-reset (SB3 2.3.2 subproc_vec_env.py _worker) -> orc_vec_step
 *     RolloutBuffer.compute_returns_and_advantage (SB3 2.3.2 buffers.py)     -> orc_gae
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * It is written for clarity (serial, dense, double precision), one environment per OrcEnv.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/b2h.h"

#define NB B2H_MAX_BODY
#define NJ B2H_MAX_JNT
#define NV B2H_MAX_DOF
#define NQ B2H_MAX_QPOS
#define NG B2H_MAX_GEOM
#define NT B2H_MAX_TENDON
#define NU B2H_MAX_ACT
#define MAXCON 256
#define MAXEFC 768
#define MINVAL 1e-15
#define MAXVAL 1e10
#define MINIMP 0.0001
#define MAXIMP 0.9999

enum { EFC_LIMIT_JOINT = 0, EFC_LIMIT_TENDON = 1, EFC_CONTACT_FRICTIONLESS = 2, EFC_CONTACT_PYRAMIDAL = 3 };

typedef struct {
  double dist, pos[3], frame[9], friction[5], solref[2], solimp[5], includemargin;
  int geom1, geom2, pair, dim, efc_address;
} OrcContact;

typedef struct OrcEnv {
  B2HModel m;
  /* state */
  double qpos[NQ], qvel[NV], qacc_warmstart[NV], ctrl[NU];
  double time;
  int nstep; /* physics steps since mj_resetData */
  /* env bookkeeping (custom_env.py) */
  int step_count;
  double total_reward;
  int n_bad; /* mj_check* resets */
  /* section 8(f4): what MuJoCo would compute if the model had the sensors the rewards assume (off = reference behaviour:
     both arrays stay zero because XML/humanoid.xml has no sensor) */
  int sensor_terms;
  double cfrc_ext[NB][6], subtree_linvel[NB][3];
  /* position stage */
  double xpos[NB][3], xquat[NB][4], xmat[NB][9], xipos[NB][3], ximat[NB][9];
  double xanchor[NJ][3], xaxis[NJ][3];
  double geom_xpos[NG][3], geom_xmat[NG][9];
  double subtree_com[NB][3], cinert[NB][10], cdof[NV][6];
  double ten_length[NT];
  double qM[NV][NV], qL[NV][NV]; /* dense M and its Cholesky factor */
  int ncon;
  OrcContact con[MAXCON];
  int nefc;
  double efc_J[MAXEFC][NV];
  double efc_pos[MAXEFC], efc_margin[MAXEFC], efc_diagApprox[MAXEFC], efc_R[MAXEFC], efc_D[MAXEFC];
  double efc_KBIP[MAXEFC][4], efc_vel[MAXEFC], efc_aref[MAXEFC], efc_force[MAXEFC];
  int efc_type[MAXEFC], efc_id[MAXEFC], efc_state[MAXEFC];
  /* velocity / acceleration stage */
  double cvel[NB][6], cdof_dot[NV][6];
  double qfrc_passive[NV], qfrc_bias[NV], qfrc_actuator[NV], qfrc_smooth[NV], qacc_smooth[NV];
  double qacc[NV], qfrc_constraint[NV];
  int solver_niter;
  int total_newton_iter;
  double solver_cost;
} OrcEnv;

/* ------------------------------------------------------------------------------------------ small math */
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static double normalize3(double* v) { /* mju_normalize3 */
  double n = sqrt(dot3(v, v));
  if (n < MINVAL) { v[0] = 1; v[1] = 0; v[2] = 0; }
  else { double s = 1 / n; v[0] *= s; v[1] *= s; v[2] *= s; }
  return n;
}
static double normalize4(double* q) { /* mju_normalize4 */
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
  else if (fabs(n - 1) > MINVAL) { double s = 1 / n; q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s; }
  return n;
}
static void mul_quat(double* r, const double* a, const double* b) {
  double t[4] = {a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                 a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                 a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                 a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]};
  memcpy(r, t, sizeof t);
}
static void quat2mat(double* R, const double* q) { /* mju_quat2Mat */
  double q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  double q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3];
  double q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  R[0] = q00 + q11 - q22 - q33; R[4] = q00 - q11 + q22 - q33; R[8] = q00 - q11 - q22 + q33;
  R[1] = 2 * (q12 - q03); R[2] = 2 * (q13 + q02);
  R[3] = 2 * (q12 + q03); R[5] = 2 * (q23 - q01);
  R[6] = 2 * (q13 - q02); R[7] = 2 * (q23 + q01);
}
static void rot_vec_quat(double* r, const double* v, const double* q) {
  double R[9]; quat2mat(R, q);
  double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2], y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2],
         z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void mat_vec3(double* r, const double* R, const double* v) {
  double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2], y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2],
         z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void axis_angle2quat(double* q, const double* axis, double angle) { /* mju_axisAngle2Quat */
  if (angle == 0) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  double s = sin(angle * 0.5);
  q[0] = cos(angle * 0.5); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
/* spatial algebra, 6-vectors are [angular; linear] (engine_util_spatial.c) */
static void mul_inert_vec(double* r, const double* i, const double* v) {
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
static void cross_motion(double* r, const double* vel, const double* v) {
  r[0] = -vel[2] * v[1] + vel[1] * v[2];
  r[1] = vel[2] * v[0] - vel[0] * v[2];
  r[2] = -vel[1] * v[0] + vel[0] * v[1];
  r[3] = -vel[2] * v[4] + vel[1] * v[5] - vel[5] * v[1] + vel[4] * v[2];
  r[4] = vel[2] * v[3] - vel[0] * v[5] + vel[5] * v[0] - vel[3] * v[2];
  r[5] = -vel[1] * v[3] + vel[0] * v[4] - vel[4] * v[0] + vel[3] * v[1];
}
static void cross_force(double* r, const double* vel, const double* f) {
  r[0] = -vel[2] * f[1] + vel[1] * f[2] - vel[5] * f[4] + vel[4] * f[5];
  r[1] = vel[2] * f[0] - vel[0] * f[2] + vel[5] * f[3] - vel[3] * f[5];
  r[2] = -vel[1] * f[0] + vel[0] * f[1] - vel[4] * f[3] + vel[3] * f[4];
  r[3] = -vel[2] * f[4] + vel[1] * f[5];
  r[4] = vel[2] * f[3] - vel[0] * f[5];
  r[5] = -vel[1] * f[3] + vel[0] * f[4];
}

/* dense Cholesky A = L L^T (lower), returns 0 on success; mju_cholFactor floors pivots at MINVAL */
static int chol_factor(double L[NV][NV], double A[NV][NV], int n) {
  int rank = n;
  for (int j = 0; j < n; j++) {
    double s = A[j][j];
    for (int k = 0; k < j; k++) s -= L[j][k] * L[j][k];
    if (s < MINVAL) { s = MINVAL; rank--; }
    L[j][j] = sqrt(s);
    double inv = 1 / L[j][j];
    for (int i = j + 1; i < n; i++) {
      double t = A[i][j];
      for (int k = 0; k < j; k++) t -= L[i][k] * L[j][k];
      L[i][j] = t * inv;
    }
  }
  return rank == n ? 0 : 1;
}
static void chol_solve(double* x, double L[NV][NV], const double* b, int n) {
  double y[NV];
  for (int i = 0; i < n; i++) {
    double t = b[i];
    for (int k = 0; k < i; k++) t -= L[i][k] * y[k];
    y[i] = t / L[i][i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double t = y[i];
    for (int k = i + 1; k < n; k++) t -= L[k][i] * x[k];
    x[i] = t / L[i][i];
  }
}

/* ------------------------------------------------------------------------------------------ position */
/* mj_kinematics (engine_core_smooth.c) */
/* [UNVERIFIED-vs-3.2.5] mj_kinematics: hinge anchor bookkeeping (xpos corrected so that the anchor stays fixed), quaternion renormalisation points */
static void kinematics(OrcEnv* d) {
  const B2HModel* m = &d->m;
  memset(d->xpos[0], 0, sizeof d->xpos[0]);
  d->xquat[0][0] = 1; d->xquat[0][1] = d->xquat[0][2] = d->xquat[0][3] = 0;
  quat2mat(d->xmat[0], d->xquat[0]);
  memset(d->xipos[0], 0, sizeof d->xipos[0]);
  quat2mat(d->ximat[0], d->xquat[0]);
  for (int i = 1; i < m->nbody; i++) {
    double xpos[3], xquat[4];
    int ja = m->body_jntadr[i], jn = m->body_jntnum[i];
    if (jn == 1 && m->jnt_type[ja] == B2H_JNT_FREE) {
      int qa = m->jnt_qposadr[ja];
      memcpy(xpos, d->qpos + qa, 3 * sizeof(double));
      memcpy(xquat, d->qpos + qa + 3, 4 * sizeof(double));
      normalize4(xquat);
      memcpy(d->xanchor[ja], xpos, sizeof xpos);
      memcpy(d->xaxis[ja], m->jnt_axis[ja], 3 * sizeof(double));
    } else {
      int pid = m->body_parentid[i];
      mat_vec3(xpos, d->xmat[pid], m->body_pos[i]);
      for (int k = 0; k < 3; k++) xpos[k] += d->xpos[pid][k];
      mul_quat(xquat, d->xquat[pid], m->body_quat[i]);
      for (int j = ja; j < ja + jn; j++) {
        double xanchor[3], xaxis[3];
        rot_vec_quat(xaxis, m->jnt_axis[j], xquat);
        rot_vec_quat(xanchor, m->jnt_pos[j], xquat);
        for (int k = 0; k < 3; k++) xanchor[k] += xpos[k];
        if (m->jnt_type[j] == B2H_JNT_HINGE) {
          int qa = m->jnt_qposadr[j];
          double qloc[4], vec[3];
          axis_angle2quat(qloc, m->jnt_axis[j], d->qpos[qa] - m->qpos0[qa]);
          mul_quat(xquat, xquat, qloc);
          rot_vec_quat(vec, m->jnt_pos[j], xquat);
          for (int k = 0; k < 3; k++) xpos[k] = xanchor[k] - vec[k];
        }
        memcpy(d->xanchor[j], xanchor, sizeof xanchor);
        memcpy(d->xaxis[j], xaxis, sizeof xaxis);
      }
    }
    normalize4(xquat);
    memcpy(d->xquat[i], xquat, sizeof xquat);
    memcpy(d->xpos[i], xpos, sizeof xpos);
    quat2mat(d->xmat[i], xquat);
  }
  for (int i = 1; i < m->nbody; i++) { /* inertial frames: mj_local2Global */
    double v[3], q[4];
    mat_vec3(v, d->xmat[i], m->body_ipos[i]);
    for (int k = 0; k < 3; k++) d->xipos[i][k] = d->xpos[i][k] + v[k];
    mul_quat(q, d->xquat[i], m->body_iquat[i]);
    quat2mat(d->ximat[i], q);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g];
    double v[3], q[4];
    mat_vec3(v, d->xmat[b], m->geom_pos[g]);
    for (int k = 0; k < 3; k++) d->geom_xpos[g][k] = d->xpos[b][k] + v[k];
    mul_quat(q, d->xquat[b], m->geom_quat[g]);
    quat2mat(d->geom_xmat[g], q);
  }
}

static int body_rootid(const B2HModel* m, int b) {
  while (b > 0 && m->body_parentid[b] != 0) b = m->body_parentid[b];
  return b;
}

/* mj_comPos: subtree_com, cinert (mju_inertCom), cdof (mju_dofCom) */
/* [UNVERIFIED-vs-3.2.5] mj_comPos: cinert about the subtree com via mju_inertCom, cdof via mju_dofCom (free joint: world-axis translations, body-axis rotations) */
static void com_pos(OrcEnv* d) {
  const B2HModel* m = &d->m;
  memset(d->subtree_com, 0, sizeof d->subtree_com);
  for (int i = m->nbody - 1; i > 0; i--) {
    for (int k = 0; k < 3; k++) d->subtree_com[i][k] += d->xipos[i][k] * m->body_mass[i];
    int p = m->body_parentid[i];
    for (int k = 0; k < 3; k++) d->subtree_com[p][k] += d->subtree_com[i][k];
  }
  for (int i = 0; i < m->nbody; i++) {
    if (m->body_subtreemass[i] < MINVAL) memcpy(d->subtree_com[i], d->xipos[i], 3 * sizeof(double));
    else for (int k = 0; k < 3; k++) d->subtree_com[i][k] /= m->body_subtreemass[i];
  }
  memset(d->cinert[0], 0, sizeof d->cinert[0]);
  for (int i = 1; i < m->nbody; i++) {
    double dif[3], tmp[9], res[9];
    const double* com = d->subtree_com[body_rootid(m, i)];
    for (int k = 0; k < 3; k++) dif[k] = d->xipos[i][k] - com[k];
    const double* R = d->ximat[i];
    const double* in = m->body_inertia[i];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) tmp[3 * r + c] = R[3 * r + c] * in[c];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++)
      res[3 * r + c] = tmp[3 * r] * R[3 * c] + tmp[3 * r + 1] * R[3 * c + 1] + tmp[3 * r + 2] * R[3 * c + 2];
    double mass = m->body_mass[i];
    double* ci = d->cinert[i];
    ci[0] = res[0] + mass * (dif[1] * dif[1] + dif[2] * dif[2]);
    ci[1] = res[4] + mass * (dif[0] * dif[0] + dif[2] * dif[2]);
    ci[2] = res[8] + mass * (dif[0] * dif[0] + dif[1] * dif[1]);
    ci[3] = res[1] - mass * dif[0] * dif[1];
    ci[4] = res[2] - mass * dif[0] * dif[2];
    ci[5] = res[5] - mass * dif[1] * dif[2];
    ci[6] = mass * dif[0]; ci[7] = mass * dif[1]; ci[8] = mass * dif[2];
    ci[9] = mass;
  }
  for (int j = 0; j < m->njnt; j++) {
    int da = m->jnt_dofadr[j], bi = m->jnt_bodyid[j];
    double off[3];
    const double* com = d->subtree_com[body_rootid(m, bi)];
    for (int k = 0; k < 3; k++) off[k] = com[k] - d->xanchor[j][k];
    if (m->jnt_type[j] == B2H_JNT_FREE) {
      memset(d->cdof[da], 0, 18 * sizeof(double));
      for (int i = 0; i < 3; i++) d->cdof[da + i][3 + i] = 1;
      for (int i = 0; i < 3; i++) {
        double axis[3] = {d->xmat[bi][i], d->xmat[bi][i + 3], d->xmat[bi][i + 6]};
        memcpy(d->cdof[da + 3 + i], axis, sizeof axis);
        cross3(d->cdof[da + 3 + i] + 3, axis, off);
      }
    } else {
      memcpy(d->cdof[da], d->xaxis[j], 3 * sizeof(double));
      cross3(d->cdof[da] + 3, d->xaxis[j], off);
    }
  }
}

/* mj_tendon (fixed tendons only) */
static void tendon(OrcEnv* d) {
  const B2HModel* m = &d->m;
  for (int t = 0; t < m->ntendon; t++) {
    double L = 0;
    for (int q = 0; q < m->nq; q++) L += m->ten_qcoef[t][q] * d->qpos[q];
    d->ten_length[t] = L;
  }
}

/* mj_crb + dense mj_factorM */
static void crb(OrcEnv* d) {
  const B2HModel* m = &d->m;
  double crb[NB][10];
  memcpy(crb, d->cinert, sizeof crb);
  for (int i = m->nbody - 1; i > 0; i--) {
    int p = m->body_parentid[i];
    if (p > 0) for (int k = 0; k < 10; k++) crb[p][k] += crb[i][k];
  }
  memset(d->qM, 0, sizeof d->qM);
  for (int i = 0; i < m->nv; i++) {
    double buf[6];
    mul_inert_vec(buf, crb[m->dof_bodyid[i]], d->cdof[i]);
    d->qM[i][i] = m->dof_armature[i];
    for (int j = i; j >= 0; j = m->dof_parentid[j]) {
      double s = 0;
      for (int k = 0; k < 6; k++) s += d->cdof[j][k] * buf[k];
      d->qM[i][j] += s;
      d->qM[j][i] = d->qM[i][j];
    }
  }
  chol_factor(d->qL, d->qM, m->nv);
}

/* ------------------------------------------------------------------------------------------ collision */
/* engine_collision_primitive.c: raw sphere-sphere on (pos, radius) pairs */
static int sphere_sphere(OrcContact* c, double margin, const double* p1, double r1, const double* p2, double r2) {
  double dif[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  double cd2 = dot3(dif, dif), mind = margin + r1 + r2;
  if (cd2 > mind * mind) return 0;
  memcpy(c->frame, dif, sizeof dif);
  c->dist = normalize3(c->frame) - r1 - r2;
  for (int k = 0; k < 3; k++) c->pos[k] = p1[k] + c->frame[k] * (r1 + c->dist / 2);
  c->frame[3] = c->frame[4] = c->frame[5] = 0;
  return 1;
}
static int plane_sphere(OrcContact* c, double margin, const double* ppos, const double* pmat, const double* spos, double r) {
  double n[3] = {pmat[2], pmat[5], pmat[8]};
  double tmp[3] = {spos[0] - ppos[0], spos[1] - ppos[1], spos[2] - ppos[2]};
  double cdist = dot3(tmp, n);
  if (cdist > margin + r) return 0;
  c->dist = cdist - r;
  memcpy(c->frame, n, sizeof n);
  for (int k = 0; k < 3; k++) c->pos[k] = spos[k] + n[k] * (-c->dist / 2 - r);
  c->frame[3] = c->frame[4] = c->frame[5] = 0;
  return 1;
}
/* [UNVERIFIED-vs-3.2.5] mjc_PlaneCapsule: two end-sphere tests in the order +axis, -axis; contact frame seeded with the capsule axis as second axis */
static int plane_capsule(OrcContact* c, double margin, const double* ppos, const double* pmat, const double* cpos,
                         const double* cmat, const double* size) {
  double axis[3] = {cmat[2], cmat[5], cmat[8]}, pos[3];
  for (int k = 0; k < 3; k++) pos[k] = cpos[k] + axis[k] * size[1];
  int n1 = plane_sphere(c, margin, ppos, pmat, pos, size[0]);
  if (n1) memcpy(c->frame + 3, axis, sizeof axis); /* align contact frame with the capsule axis */
  for (int k = 0; k < 3; k++) pos[k] = cpos[k] - axis[k] * size[1];
  int n2 = plane_sphere(c + n1, margin, ppos, pmat, pos, size[0]);
  if (n2) memcpy(c[n1].frame + 3, axis, sizeof axis);
  return n1 + n2;
}
static double clipd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
static int sphere_capsule(OrcContact* c, double margin, const double* spos, double sr, const double* cpos, const double* cmat,
                          const double* size) {
  double axis[3] = {cmat[2], cmat[5], cmat[8]};
  double vec[3] = {spos[0] - cpos[0], spos[1] - cpos[1], spos[2] - cpos[2]};
  double x = clipd(dot3(axis, vec), -size[1], size[1]);
  double pos[3] = {cpos[0] + axis[0] * x, cpos[1] + axis[1] * x, cpos[2] + axis[2] * x};
  return sphere_sphere(c, margin, spos, sr, pos, size[0]);
}
/* [UNVERIFIED-vs-3.2.5] mjc_CapsuleCapsule: closest-point solve (ma, mb, mc, u, v, det), clamping order, parallel-axis branch (|det| < mjMINVAL: up to two contacts) */
static int capsule_capsule(OrcContact* c, double margin, const double* pos1, const double* mat1, const double* size1,
                           const double* pos2, const double* mat2, const double* size2) {
  double axis1[3] = {mat1[2], mat1[5], mat1[8]}, axis2[3] = {mat2[2], mat2[5], mat2[8]};
  double dif[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
  double ma = dot3(axis1, axis1), mb = -dot3(axis1, axis2), mc = dot3(axis2, axis2);
  double u = -dot3(axis1, dif), v = dot3(axis2, dif), det = ma * mc - mb * mb;
  double vec1[3], vec2[3];
  if (fabs(det) >= MINVAL) {
    double x1 = (mc * u - mb * v) / det, x2 = (ma * v - mb * u) / det;
    if (x1 > size1[1]) { x1 = size1[1]; x2 = (v - mb * size1[1]) / mc; }
    else if (x1 < -size1[1]) { x1 = -size1[1]; x2 = (v + mb * size1[1]) / mc; }
    if (x2 > size2[1]) { x2 = size2[1]; x1 = clipd((u - mb * size2[1]) / ma, -size1[1], size1[1]); }
    else if (x2 < -size2[1]) { x2 = -size2[1]; x1 = clipd((u + mb * size2[1]) / ma, -size1[1], size1[1]); }
    for (int k = 0; k < 3; k++) { vec1[k] = pos1[k] + axis1[k] * x1; vec2[k] = pos2[k] + axis2[k] * x2; }
    return sphere_sphere(c, margin, vec1, size1[0], vec2, size2[0]);
  }
  /* parallel axes: test the four end-point projections, keep at most two contacts */
  int n = 0;
  for (int s = 1; s >= -1 && n < 2; s -= 2) {
    double t[3];
    for (int k = 0; k < 3; k++) { vec1[k] = pos1[k] + axis1[k] * size1[1] * s; t[k] = vec1[k] - pos2[k]; }
    double x2 = clipd(dot3(t, axis2), -size2[1], size2[1]);
    for (int k = 0; k < 3; k++) vec2[k] = pos2[k] + axis2[k] * x2;
    n += sphere_sphere(c + n, margin, vec1, size1[0], vec2, size2[0]);
  }
  for (int s = 1; s >= -1 && n < 2; s -= 2) {
    double t[3];
    for (int k = 0; k < 3; k++) { vec2[k] = pos2[k] + axis2[k] * size2[1] * s; t[k] = vec2[k] - pos1[k]; }
    double x1 = clipd(dot3(t, axis1), -size1[1], size1[1]);
    for (int k = 0; k < 3; k++) vec1[k] = pos1[k] + axis1[k] * x1;
    n += sphere_sphere(c + n, margin, vec1, size1[0], vec2, size2[0]);
  }
  return n;
}
/* [UNVERIFIED-vs-3.2.5] mju_makeFrame: fallback second axis (y unless |normal.y| >= 0.5, then z), Gram-Schmidt, third = cross */
static void make_frame(double* f) { /* mju_makeFrame */
  normalize3(f);
  if (sqrt(dot3(f + 3, f + 3)) < 0.5) {
    f[3] = f[4] = f[5] = 0;
    if (f[1] < 0.5 && f[1] > -0.5) f[4] = 1; else f[5] = 1;
  }
  double dp = dot3(f, f + 3);
  for (int k = 0; k < 3; k++) f[3 + k] -= f[k] * dp;
  normalize3(f + 3);
  cross3(f + 6, f, f + 3);
}

/* mj_collision over the static candidate list (engine_collision_driver.c; broadphase only culls) */
/* [UNVERIFIED-vs-3.2.5] mj_collision: contact ORDER follows this repo's pair list (mjcf.py), which may differ from MuJoCo's body-pair sweep; order changes no result, only summation order */
static void collision(OrcEnv* d) {
  const B2HModel* m = &d->m;
  d->ncon = 0;
  for (int p = 0; p < m->npair; p++) {
    int g1 = m->pair_geom1[p], g2 = m->pair_geom2[p];
    int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
    double margin = m->pair_margin[p];
    OrcContact tmp[4];
    int n = 0;
    if (t1 == B2H_GEOM_PLANE && t2 == B2H_GEOM_SPHERE)
      n = plane_sphere(tmp, margin, d->geom_xpos[g1], d->geom_xmat[g1], d->geom_xpos[g2], m->geom_size[g2][0]);
    else if (t1 == B2H_GEOM_PLANE && t2 == B2H_GEOM_CAPSULE)
      n = plane_capsule(tmp, margin, d->geom_xpos[g1], d->geom_xmat[g1], d->geom_xpos[g2], d->geom_xmat[g2], m->geom_size[g2]);
    else if (t1 == B2H_GEOM_SPHERE && t2 == B2H_GEOM_SPHERE)
      n = sphere_sphere(tmp, margin, d->geom_xpos[g1], m->geom_size[g1][0], d->geom_xpos[g2], m->geom_size[g2][0]);
    else if (t1 == B2H_GEOM_SPHERE && t2 == B2H_GEOM_CAPSULE)
      n = sphere_capsule(tmp, margin, d->geom_xpos[g1], m->geom_size[g1][0], d->geom_xpos[g2], d->geom_xmat[g2], m->geom_size[g2]);
    else if (t1 == B2H_GEOM_CAPSULE && t2 == B2H_GEOM_CAPSULE)
      n = capsule_capsule(tmp, margin, d->geom_xpos[g1], d->geom_xmat[g1], m->geom_size[g1], d->geom_xpos[g2],
                          d->geom_xmat[g2], m->geom_size[g2]);
    for (int i = 0; i < n; i++) {
      if (!(tmp[i].dist < margin)) continue;
      if (d->ncon >= MAXCON) { fprintf(stderr, "oracle: contact buffer overflow\n"); abort(); }
      OrcContact* c = &d->con[d->ncon++];
      *c = tmp[i];
      make_frame(c->frame);
      c->geom1 = g1; c->geom2 = g2; c->pair = p; c->dim = m->pair_condim[p];
      c->friction[0] = c->friction[1] = m->pair_friction[p][0];
      c->friction[2] = m->pair_friction[p][1];
      c->friction[3] = c->friction[4] = m->pair_friction[p][2];
      memcpy(c->solref, m->pair_solref[p], sizeof c->solref);
      memcpy(c->solimp, m->pair_solimp[p], sizeof c->solimp);
      c->includemargin = m->pair_margin[p] - m->pair_gap[p];
    }
  }
}

/* ------------------------------------------------------------------------------------------ constraints */
/* translational Jacobian of `point` on `body` (mj_jac) */
static void jac_point(const OrcEnv* d, double jacp[3][NV], const double* point, int body) {
  const B2HModel* m = &d->m;
  for (int r = 0; r < 3; r++) memset(jacp[r], 0, sizeof(double) * NV);
  double off[3];
  const double* com = d->subtree_com[body_rootid(m, body)];
  for (int k = 0; k < 3; k++) off[k] = point[k] - com[k];
  for (int i = m->body_lastdof[body]; i >= 0; i = m->dof_parentid[i]) {
    double tmp[3];
    cross3(tmp, d->cdof[i], off);
    for (int k = 0; k < 3; k++) jacp[k][i] = d->cdof[i][3 + k] + tmp[k];
  }
}
static int add_row(OrcEnv* d, const double* J, double pos, double margin, int type, int id, double diagApprox) {
  if (d->nefc >= MAXEFC) { fprintf(stderr, "oracle: efc buffer overflow\n"); abort(); }
  int r = d->nefc++;
  memcpy(d->efc_J[r], J, sizeof(double) * NV);
  d->efc_pos[r] = pos; d->efc_margin[r] = margin; d->efc_type[r] = type; d->efc_id[r] = id;
  d->efc_diagApprox[r] = diagApprox;
  return r;
}
/* [UNVERIFIED-vs-3.2.5] getimpedance: x = |pos - margin| / width, power curve with midpoint, clamps at 0 / 1 */
static void get_impedance(const double* solimp, double pos, double margin, double* imp) {
  if (solimp[0] == solimp[1] || solimp[2] <= MINVAL) { *imp = 0.5 * (solimp[0] + solimp[1]); return; }
  double x = fabs((pos - margin) / solimp[2]);
  if (x >= 1 || x <= 0) { *imp = x >= 1 ? solimp[1] : solimp[0]; return; }
  double y;
  if (solimp[4] == 1) y = x;
  else if (x <= solimp[3]) y = pow(x, solimp[4]) / pow(solimp[3], solimp[4] - 1);
  else y = 1 - pow(1 - x, solimp[4]) / pow(1 - solimp[3], solimp[4] - 1);
  *imp = solimp[0] + y * (solimp[1] - solimp[0]);
}
/* mj_makeConstraint: limits (joint, tendon) then contacts; then mj_makeImpedance */
/* [UNVERIFIED-vs-3.2.5] mj_makeConstraint / mj_makeImpedance / mj_diagApprox: limit distance and Jacobian sign, pyramidal rows (n + mu t, n - mu t per tangent), diagApprox = tran (+ mu^2 tran), refsafe clamp, K / B from solref with dmax, R = max(mjMINVAL, (1 - imp) / imp * diagApprox), pyramidal Rpy = 2 mu^2 R[first row] */
static void make_constraint(OrcEnv* d) {
  const B2HModel* m = &d->m;
  int nv = m->nv;
  d->nefc = 0;
  double J[NV];
  for (int j = 0; j < m->njnt; j++) { /* mj_instantiateLimit, joints */
    if (!m->jnt_limited[j] || m->jnt_type[j] != B2H_JNT_HINGE) continue;
    double value = d->qpos[m->jnt_qposadr[j]], margin = m->jnt_margin[j];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (m->jnt_range[j][(side + 1) / 2] - value);
      if (dist < margin) {
        memset(J, 0, sizeof J);
        J[m->jnt_dofadr[j]] = -side;
        add_row(d, J, dist, margin, EFC_LIMIT_JOINT, j, m->dof_invweight0[m->jnt_dofadr[j]]);
      }
    }
  }
  for (int t = 0; t < m->ntendon; t++) { /* tendon limits */
    if (!m->ten_limited[t]) continue;
    double value = d->ten_length[t], margin = m->ten_margin[t];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (m->ten_range[t][(side + 1) / 2] - value);
      if (dist < margin) {
        for (int k = 0; k < NV; k++) J[k] = -side * m->ten_J[t][k];
        add_row(d, J, dist, margin, EFC_LIMIT_TENDON, t, m->ten_invweight0[t]);
      }
    }
  }
  for (int ci = 0; ci < d->ncon; ci++) { /* mj_instantiateContact */
    OrcContact* c = &d->con[ci];
    int b1 = m->geom_bodyid[c->geom1], b2 = m->geom_bodyid[c->geom2];
    double j1[3][NV], j2[3][NV], jc[3][NV];
    jac_point(d, j1, c->pos, b1);
    jac_point(d, j2, c->pos, b2);
    for (int r = 0; r < 3; r++)
      for (int k = 0; k < nv; k++)
        jc[r][k] = c->frame[3 * r] * (j2[0][k] - j1[0][k]) + c->frame[3 * r + 1] * (j2[1][k] - j1[1][k]) +
                   c->frame[3 * r + 2] * (j2[2][k] - j1[2][k]);
    /* [UNVERIFIED-vs-3.2.5] mj_diagApprox (engine_core_constraint.c, contact cases): the body weights are indexed
       with the geoms' own bodies, `bid = m->geom_bodyid[con->geom[side]]; tran += m->body_invweight0[2*bid]` --
       NOT the weld parent.  mj_setConst computes a distinct body_invweight0 for a jointless child (head, hands:
       mj_jacBodyCom at its own inertial frame), so the two readings differ for head / hand contacts.  Round 1 used
       the weld body; round 2 follows the upstream indexing as recalled from the 2.x and 3.x sources. */
    double tran = m->body_invweight0[b1][0] + m->body_invweight0[b2][0];
    c->efc_address = d->nefc;
    if (c->dim == 1) {
      memset(J, 0, sizeof J);
      memcpy(J, jc[0], sizeof(double) * nv);
      add_row(d, J, c->dist, c->includemargin, EFC_CONTACT_FRICTIONLESS, ci, tran);
    } else {
      for (int k = 1; k < c->dim; k++) {
        double fri = c->friction[k - 1];
        for (int s = 1; s >= -1; s -= 2) {
          memset(J, 0, sizeof J);
          for (int q = 0; q < nv; q++) J[q] = jc[0][q] + s * fri * jc[k][q];
          add_row(d, J, c->dist, c->includemargin, EFC_CONTACT_PYRAMIDAL, ci, tran + fri * fri * tran);
        }
      }
    }
  }
  /* mj_makeImpedance */
  for (int i = 0; i < d->nefc; i++) {
    double solref[2], solimp[5];
    int id = d->efc_id[i];
    switch (d->efc_type[i]) {
      case EFC_LIMIT_JOINT: memcpy(solref, m->jnt_solref[id], sizeof solref); memcpy(solimp, m->jnt_solimp[id], sizeof solimp); break;
      case EFC_LIMIT_TENDON: memcpy(solref, m->ten_solref[id], sizeof solref); memcpy(solimp, m->ten_solimp[id], sizeof solimp); break;
      default: memcpy(solref, d->con[id].solref, sizeof solref); memcpy(solimp, d->con[id].solimp, sizeof solimp);
    }
    /* getsolparam clamps */
    if (solref[0] > 0 && solref[0] < 2 * m->timestep) solref[0] = 2 * m->timestep; /* refsafe */
    solimp[0] = clipd(solimp[0], MINIMP, MAXIMP);
    solimp[1] = clipd(solimp[1], MINIMP, MAXIMP);
    solimp[2] = solimp[2] < 0 ? 0 : solimp[2];
    solimp[3] = clipd(solimp[3], MINIMP, MAXIMP);
    solimp[4] = solimp[4] < 1 ? 1 : solimp[4];
    double imp;
    get_impedance(solimp, d->efc_pos[i], d->efc_margin[i], &imp);
    double dmax = solimp[1];
    if (solref[0] <= 0) { /* direct stiffness / damping */
      d->efc_KBIP[i][0] = -solref[0] / fmax(MINVAL, dmax * dmax);
      d->efc_KBIP[i][1] = -solref[1] / fmax(MINVAL, dmax);
    } else {
      d->efc_KBIP[i][0] = 1 / fmax(MINVAL, dmax * dmax * solref[0] * solref[0] * solref[1] * solref[1]);
      d->efc_KBIP[i][1] = 2 / fmax(MINVAL, dmax * solref[0]);
    }
    d->efc_KBIP[i][2] = imp; d->efc_KBIP[i][3] = 0;
    d->efc_R[i] = fmax(MINVAL, (1 - imp) * d->efc_diagApprox[i] / imp);
  }
  for (int i = 0; i < d->nefc; i++) { /* pyramidal rows share one regulariser: Rpy = 2 mu^2 R (impratio 1) */
    if (d->efc_type[i] != EFC_CONTACT_PYRAMIDAL) continue;
    OrcContact* c = &d->con[d->efc_id[i]];
    double mu = c->friction[0];
    double Rpy = 2 * mu * mu * d->efc_R[i];
    int nrow = 2 * (c->dim - 1);
    for (int j = 0; j < nrow; j++) d->efc_R[i + j] = Rpy;
    i += nrow - 1;
  }
  for (int i = 0; i < d->nefc; i++) d->efc_D[i] = 1 / d->efc_R[i];
}

/* ------------------------------------------------------------------------------------------ velocity */
/* [UNVERIFIED-vs-3.2.5] mj_comVel: free-joint special case (translational cdof_dot = 0, the three rotational ones see cvel after the translations only) */
static void com_vel(OrcEnv* d) { /* mj_comVel */
  const B2HModel* m = &d->m;
  memset(d->cvel[0], 0, sizeof d->cvel[0]);
  for (int i = 1; i < m->nbody; i++) {
    double cvel[6];
    memcpy(cvel, d->cvel[m->body_parentid[i]], sizeof cvel);
    int bda = m->body_dofadr[i], nd = m->body_dofnum[i];
    for (int j = 0; j < nd; j++) {
      int dj = bda + j;
      if (m->jnt_type[m->dof_jntid[dj]] == B2H_JNT_FREE) {
        for (int k = 0; k < 3; k++) {
          memset(d->cdof_dot[dj + k], 0, sizeof d->cdof_dot[0]);
          for (int q = 0; q < 6; q++) cvel[q] += d->cdof[dj + k][q] * d->qvel[dj + k];
        }
        for (int k = 3; k < 6; k++) cross_motion(d->cdof_dot[dj + k], cvel, d->cdof[dj + k]);
        for (int k = 3; k < 6; k++) for (int q = 0; q < 6; q++) cvel[q] += d->cdof[dj + k][q] * d->qvel[dj + k];
        j += 5;
      } else {
        cross_motion(d->cdof_dot[dj], cvel, d->cdof[dj]);
        for (int q = 0; q < 6; q++) cvel[q] += d->cdof[dj][q] * d->qvel[dj];
      }
    }
    memcpy(d->cvel[i], cvel, sizeof cvel);
  }
}
static void passive(OrcEnv* d) { /* mj_passive: joint springs and dampers only */
  const B2HModel* m = &d->m;
  for (int i = 0; i < m->nv; i++) d->qfrc_passive[i] = -m->dof_damping[i] * d->qvel[i];
  for (int j = 0; j < m->njnt; j++) {
    if (m->jnt_type[j] != B2H_JNT_HINGE) continue;
    int qa = m->jnt_qposadr[j];
    d->qfrc_passive[m->jnt_dofadr[j]] += -m->jnt_stiffness[j] * (d->qpos[qa] - m->qpos_spring[qa]);
  }
}
/* [UNVERIFIED-vs-3.2.5] aref = -B vel - K imp (pos - margin) */
static void reference_constraint(OrcEnv* d) { /* mj_referenceConstraint */
  int nv = d->m.nv;
  for (int i = 0; i < d->nefc; i++) {
    double v = 0;
    for (int k = 0; k < nv; k++) v += d->efc_J[i][k] * d->qvel[k];
    d->efc_vel[i] = v;
    d->efc_aref[i] = -d->efc_KBIP[i][1] * v - d->efc_KBIP[i][0] * d->efc_KBIP[i][2] * (d->efc_pos[i] - d->efc_margin[i]);
  }
}
static void rne_bias(OrcEnv* d) { /* mj_rne(flg_acc=0) */
  const B2HModel* m = &d->m;
  double cacc[NB][6], cfrc[NB][6];
  memset(cacc[0], 0, sizeof cacc[0]);
  for (int k = 0; k < 3; k++) cacc[0][3 + k] = -m->gravity[k];
  for (int i = 1; i < m->nbody; i++) {
    int bda = m->body_dofadr[i];
    memcpy(cacc[i], cacc[m->body_parentid[i]], sizeof cacc[0]);
    for (int j = 0; j < m->body_dofnum[i]; j++)
      for (int q = 0; q < 6; q++) cacc[i][q] += d->cdof_dot[bda + j][q] * d->qvel[bda + j];
    double tmp[6], tmp1[6];
    mul_inert_vec(cfrc[i], d->cinert[i], cacc[i]);
    mul_inert_vec(tmp, d->cinert[i], d->cvel[i]);
    cross_force(tmp1, d->cvel[i], tmp);
    for (int q = 0; q < 6; q++) cfrc[i][q] += tmp1[q];
  }
  memset(cfrc[0], 0, sizeof cfrc[0]);
  for (int i = m->nbody - 1; i > 0; i--) {
    int p = m->body_parentid[i];
    if (p) for (int q = 0; q < 6; q++) cfrc[p][q] += cfrc[i][q];
  }
  for (int i = 0; i < m->nv; i++) {
    double s = 0;
    for (int q = 0; q < 6; q++) s += d->cdof[i][q] * cfrc[m->dof_bodyid[i]][q];
    d->qfrc_bias[i] = s;
  }
}
static void actuation(OrcEnv* d) { /* mj_fwdActuation: motors, gain 1, ctrl clamp */
  const B2HModel* m = &d->m;
  memset(d->qfrc_actuator, 0, sizeof d->qfrc_actuator);
  for (int a = 0; a < m->nu; a++) {
    double c = d->ctrl[a];
    if (m->actuator_ctrllimited[a]) c = clipd(c, m->actuator_ctrlrange[a][0], m->actuator_ctrlrange[a][1]);
    d->qfrc_actuator[m->actuator_dofid[a]] += m->actuator_gear[a] * c;
  }
}
static void acceleration(OrcEnv* d) { /* mj_fwdAcceleration */
  int nv = d->m.nv;
  for (int i = 0; i < nv; i++) d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i];
  chol_solve(d->qacc_smooth, d->qL, d->qfrc_smooth, nv);
}

/* ------------------------------------------------------------------------------------------ Newton solver */
typedef struct {
  OrcEnv* d;
  int nv, nefc;
  double Jaref[MAXEFC], Jv[MAXEFC], Ma[NV], Mv[NV], grad[NV], Mgrad[NV], search[NV];
  double quad[MAXEFC][3], quadGauss[3];
  double cost, gauss;
  double H[NV][NV], HL[NV][NV];
  int LSiter;
} Primal;
typedef struct { double alpha, cost, deriv[2]; } PrimalPnt;

static void mul_M(const OrcEnv* d, double* r, const double* v) {
  int nv = d->m.nv;
  for (int i = 0; i < nv; i++) { double s = 0; for (int k = 0; k < nv; k++) s += d->qM[i][k] * v[k]; r[i] = s; }
}
/* mj_constraintUpdate for unilateral quadratic rows: force, state, cost */
static double constraint_update(OrcEnv* d, const double* jar, int set_force) {
  double cost = 0;
  for (int i = 0; i < d->nefc; i++) {
    if (jar[i] < 0) {
      cost += 0.5 * d->efc_D[i] * jar[i] * jar[i];
      if (set_force) { d->efc_force[i] = -d->efc_D[i] * jar[i]; d->efc_state[i] = 1; }
    } else if (set_force) { d->efc_force[i] = 0; d->efc_state[i] = 0; }
  }
  return cost;
}
static void primal_update_constraint(Primal* c) {
  OrcEnv* d = c->d;
  c->cost = constraint_update(d, c->Jaref, 1);
  for (int k = 0; k < c->nv; k++) {
    double s = 0;
    for (int i = 0; i < c->nefc; i++) s += d->efc_J[i][k] * d->efc_force[i];
    d->qfrc_constraint[k] = s;
  }
  double g = 0;
  for (int k = 0; k < c->nv; k++) g += 0.5 * (c->Ma[k] - d->qfrc_smooth[k]) * (d->qacc[k] - d->qacc_smooth[k]);
  c->gauss = g;
  c->cost += g;
}
static void make_hessian(Primal* c) {
  OrcEnv* d = c->d;
  int nv = c->nv;
  for (int i = 0; i < nv; i++) for (int j = 0; j < nv; j++) c->H[i][j] = d->qM[i][j];
  for (int r = 0; r < c->nefc; r++) {
    if (!d->efc_state[r]) continue;
    double D = d->efc_D[r];
    for (int i = 0; i < nv; i++) {
      double t = D * d->efc_J[r][i];
      if (t == 0) continue;
      for (int j = 0; j < nv; j++) c->H[i][j] += t * d->efc_J[r][j];
    }
  }
  chol_factor(c->HL, c->H, nv);
}
static void primal_update_gradient(Primal* c) {
  OrcEnv* d = c->d;
  for (int k = 0; k < c->nv; k++) c->grad[k] = c->Ma[k] - d->qfrc_smooth[k] - d->qfrc_constraint[k];
  chol_solve(c->Mgrad, c->HL, c->grad, c->nv);
}
static void primal_eval(Primal* c, PrimalPnt* p) {
  double q0 = c->quadGauss[0], q1 = c->quadGauss[1], q2 = c->quadGauss[2], a = p->alpha;
  for (int i = 0; i < c->nefc; i++)
    if (c->Jaref[i] + a * c->Jv[i] < 0) { q0 += c->quad[i][0]; q1 += c->quad[i][1]; q2 += c->quad[i][2]; }
  p->cost = a * a * q2 + a * q1 + q0;
  p->deriv[0] = 2 * a * q2 + q1;
  p->deriv[1] = 2 * q2;
  if (p->deriv[1] <= 0) p->deriv[1] = MINVAL;
  c->LSiter++;
}
static int update_bracket(Primal* c, PrimalPnt* p, const PrimalPnt cand[3], PrimalPnt* pnext) {
  int flag = 0;
  for (int i = 0; i < 3; i++) {
    if (p->deriv[0] < 0 && cand[i].deriv[0] < 0 && p->deriv[0] < cand[i].deriv[0]) { *p = cand[i]; flag = 1; }
    else if (p->deriv[0] > 0 && cand[i].deriv[0] > 0 && p->deriv[0] > cand[i].deriv[0]) { *p = cand[i]; flag = 1; }
  }
  if (flag) { pnext->alpha = p->alpha - p->deriv[0] / p->deriv[1]; primal_eval(c, pnext); }
  return flag;
}
/* PrimalSearch (engine_solver.c): exact line search on the piecewise-quadratic cost along `search` */
/* [UNVERIFIED-vs-3.2.5] PrimalSearch (engine_solver.c): gtol = tolerance * ls_tolerance * snorm / scale, bracketing with p1 / p2 / midpoint candidates, the exit rules */
static double primal_search(Primal* c, double tolerance, double ls_tolerance, int ls_iterations) {
  OrcEnv* d = c->d;
  int nv = c->nv;
  c->LSiter = 0;
  double snorm = 0;
  for (int k = 0; k < nv; k++) snorm += c->search[k] * c->search[k];
  snorm = sqrt(snorm);
  if (snorm < MINVAL) return 0;
  double scale = 1 / (d->m.meaninertia * (nv > 1 ? nv : 1));
  double gtol = tolerance * ls_tolerance * snorm / scale;
  mul_M(d, c->Mv, c->search);
  for (int i = 0; i < c->nefc; i++) {
    double s = 0;
    for (int k = 0; k < nv; k++) s += d->efc_J[i][k] * c->search[k];
    c->Jv[i] = s;
  }
  c->quadGauss[0] = c->gauss; c->quadGauss[1] = 0; c->quadGauss[2] = 0;
  for (int k = 0; k < nv; k++) {
    c->quadGauss[1] += c->search[k] * (c->Ma[k] - d->qfrc_smooth[k]);
    c->quadGauss[2] += 0.5 * c->search[k] * c->Mv[k];
  }
  for (int i = 0; i < c->nefc; i++) {
    double D = d->efc_D[i];
    c->quad[i][0] = 0.5 * D * c->Jaref[i] * c->Jaref[i];
    c->quad[i][1] = D * c->Jaref[i] * c->Jv[i];
    c->quad[i][2] = 0.5 * D * c->Jv[i] * c->Jv[i];
  }
  PrimalPnt p0, p1, p2, pmid, p1next, p2next;
  p0.alpha = 0; primal_eval(c, &p0);
  p1.alpha = p0.alpha - p0.deriv[0] / p0.deriv[1]; primal_eval(c, &p1);
  if (p0.cost < p1.cost) p1 = p0;
  if (fabs(p1.deriv[0]) < gtol) return p1.alpha;
  int dir = p1.deriv[0] < 0 ? 1 : -1;
  int p2update = 0;
  p2 = p1;
  while (p1.deriv[0] * dir <= -gtol && c->LSiter < ls_iterations) {
    p2 = p1; p2update = 1;
    p1.alpha = p1.alpha - p1.deriv[0] / p1.deriv[1]; primal_eval(c, &p1);
    if (fabs(p1.deriv[0]) < gtol) return p1.alpha;
  }
  if (c->LSiter >= ls_iterations) return p1.alpha;
  if (!p2update) return p1.alpha;
  p2next = p1;
  p1next.alpha = p1.alpha - p1.deriv[0] / p1.deriv[1]; primal_eval(c, &p1next);
  while (c->LSiter < ls_iterations) {
    pmid.alpha = 0.5 * (p1.alpha + p2.alpha); primal_eval(c, &pmid);
    PrimalPnt cand[3] = {p1next, p2next, pmid};
    double bestcost = 0; int best = -1;
    for (int i = 0; i < 3; i++)
      if (fabs(cand[i].deriv[0]) < gtol && (best == -1 || cand[i].cost < bestcost)) { bestcost = cand[i].cost; best = i; }
    if (best >= 0) return cand[best].alpha;
    int b1 = update_bracket(c, &p1, cand, &p1next);
    int b2 = update_bracket(c, &p2, cand, &p2next);
    if (!b1 && !b2) return pmid.cost < p0.cost ? pmid.alpha : 0;
  }
  if (p1.cost <= p2.cost && p1.cost < p0.cost) return p1.alpha;
  if (p2.cost <= p1.cost && p2.cost < p0.cost) return p2.alpha;
  return 0;
}

/* mj_fwdConstraint: warmstart selection + mj_solNewton (solver Newton, 100 iterations, tolerance 1e-8) */
/* [UNVERIFIED-vs-3.2.5] mj_fwdConstraint + mj_solNewton: warm start = the cheaper of qacc_warmstart and qacc_smooth, scale = 1 / (meaninertia * max(1, nv)), stop on improvement or gradient < tolerance */
static void fwd_constraint(OrcEnv* d) {
  const B2HModel* m = &d->m;
  int nv = m->nv, nefc = d->nefc;
  d->solver_niter = 0;
  if (!nefc) {
    memcpy(d->qacc, d->qacc_smooth, sizeof(double) * nv);
    memcpy(d->qacc_warmstart, d->qacc_smooth, sizeof(double) * nv);
    memset(d->qfrc_constraint, 0, sizeof d->qfrc_constraint);
    return;
  }
  static __thread Primal ctx; /* large; one per thread */
  Primal* c = &ctx;
  c->d = d; c->nv = nv; c->nefc = nefc;
  /* warmstart(): the better of qacc_warmstart and qacc_smooth */
  double jar[MAXEFC], Ma[NV];
  for (int i = 0; i < nefc; i++) {
    double s = 0;
    for (int k = 0; k < nv; k++) s += d->efc_J[i][k] * d->qacc_warmstart[k];
    jar[i] = s - d->efc_aref[i];
  }
  double cost_warm = constraint_update(d, jar, 0);
  mul_M(d, Ma, d->qacc_warmstart);
  for (int k = 0; k < nv; k++) cost_warm += 0.5 * (Ma[k] - d->qfrc_smooth[k]) * (d->qacc_warmstart[k] - d->qacc_smooth[k]);
  for (int i = 0; i < nefc; i++) {
    double s = 0;
    for (int k = 0; k < nv; k++) s += d->efc_J[i][k] * d->qacc_smooth[k];
    jar[i] = s - d->efc_aref[i];
  }
  double cost_smooth = constraint_update(d, jar, 0);
  memcpy(d->qacc, cost_warm > cost_smooth ? d->qacc_smooth : d->qacc_warmstart, sizeof(double) * nv);

  /* mj_solPrimal, Newton */
  for (int i = 0; i < nefc; i++) {
    double s = 0;
    for (int k = 0; k < nv; k++) s += d->efc_J[i][k] * d->qacc[k];
    c->Jaref[i] = s - d->efc_aref[i];
  }
  mul_M(d, c->Ma, d->qacc);
  primal_update_constraint(c);
  make_hessian(c);
  primal_update_gradient(c);
  for (int k = 0; k < nv; k++) c->search[k] = -c->Mgrad[k];
  double scale = 1 / (m->meaninertia * (nv > 1 ? nv : 1));
  const double tolerance = 1e-8, ls_tolerance = 0.01;
  const int maxiter = 100, ls_iterations = 50;
  int iter = 0;
  while (iter < maxiter) {
    double alpha = primal_search(c, tolerance, ls_tolerance, ls_iterations);
    if (alpha == 0) break;
    for (int k = 0; k < nv; k++) { d->qacc[k] += alpha * c->search[k]; c->Ma[k] += alpha * c->Mv[k]; }
    for (int i = 0; i < nefc; i++) c->Jaref[i] += alpha * c->Jv[i];
    double oldcost = c->cost;
    primal_update_constraint(c);
    make_hessian(c);
    primal_update_gradient(c);
    double improvement = scale * (oldcost - c->cost);
    double gn = 0;
    for (int k = 0; k < nv; k++) gn += c->grad[k] * c->grad[k];
    double gradient = scale * sqrt(gn);
    iter++;
    if (improvement < tolerance || gradient < tolerance) break;
    for (int k = 0; k < nv; k++) c->search[k] = -c->Mgrad[k];
  }
  d->solver_niter = iter;
  d->total_newton_iter += iter;
  d->solver_cost = c->cost;
  memcpy(d->qacc_warmstart, d->qacc, sizeof(double) * nv);
}

/* ------------------------------------------------------------------------------------------ pipeline */
static int bad_vec(const double* v, int n) {
  for (int i = 0; i < n; i++) if (isnan(v[i]) || v[i] > MAXVAL || v[i] < -MAXVAL) return 1;
  return 0;
}
void orc_reset_data(OrcEnv* d) { /* mj_resetData */
  const B2HModel* m = &d->m;
  memcpy(d->qpos, m->qpos0, sizeof(double) * m->nq);
  memset(d->qvel, 0, sizeof d->qvel);
  memset(d->qacc_warmstart, 0, sizeof d->qacc_warmstart);
  memset(d->ctrl, 0, sizeof d->ctrl);
  memset(d->qacc, 0, sizeof d->qacc);
  memset(d->cinert, 0, sizeof d->cinert);
  memset(d->cvel, 0, sizeof d->cvel);
  memset(d->subtree_com, 0, sizeof d->subtree_com);
  memset(d->qfrc_actuator, 0, sizeof d->qfrc_actuator);
  memset(d->cfrc_ext, 0, sizeof d->cfrc_ext);
  memset(d->subtree_linvel, 0, sizeof d->subtree_linvel);
  d->time = 0; d->nstep = 0; d->ncon = 0; d->nefc = 0;
}
/* [UNVERIFIED-vs-3.2.5] mj_subtreeVel (linear part): body momentum m * (cvel_lin + omega x (xipos - subtree_com[root])) summed over
   each subtree, divided by the subtree mass */
static void subtree_vel(OrcEnv* d) {
  const B2HModel* m = &d->m;
  double mom[NB][3], mass[NB];
  for (int b = 0; b < m->nbody; b++) {
    double dif[3], v[3];
    const double* com = d->subtree_com[b ? body_rootid(m, b) : 0];
    for (int k = 0; k < 3; k++) dif[k] = d->xipos[b][k] - com[k];
    cross3(v, d->cvel[b], dif);
    for (int k = 0; k < 3; k++) mom[b][k] = m->body_mass[b] * (d->cvel[b][3 + k] + v[k]);
    mass[b] = m->body_mass[b];
  }
  for (int b = m->nbody - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    for (int k = 0; k < 3; k++) mom[p][k] += mom[b][k];
    mass[p] += mass[b];
  }
  for (int b = 0; b < m->nbody; b++)
    for (int k = 0; k < 3; k++) d->subtree_linvel[b][k] = mass[b] > MINVAL ? mom[b][k] / mass[b] : 0.0;
}
/* [UNVERIFIED-vs-3.2.5] mj_rnePostConstraint, contact part of cfrc_ext: the contact force (mj_contactForce: frictionless = the
   row's force; pyramidal = normal sum of the four edge forces, tangents mu * (f0 - f1), mu * (f2 - f3)) rotated to the
   world, as a spatial force [torque; force] about the subtree com of the body's tree, subtracted from body 1, added to body 2 */
static void rne_post_constraint(OrcEnv* d) {
  const B2HModel* m = &d->m;
  memset(d->cfrc_ext, 0, sizeof d->cfrc_ext);
  for (int ci = 0; ci < d->ncon; ci++) {
    const OrcContact* c = &d->con[ci];
    int r = c->efc_address;
    double lf[3] = {0, 0, 0};
    if (c->dim == 1) lf[0] = d->efc_force[r];
    else {
      lf[0] = d->efc_force[r] + d->efc_force[r + 1] + d->efc_force[r + 2] + d->efc_force[r + 3];
      lf[1] = c->friction[0] * (d->efc_force[r] - d->efc_force[r + 1]);
      lf[2] = c->friction[1] * (d->efc_force[r + 2] - d->efc_force[r + 3]);
    }
    double f[3];
    for (int k = 0; k < 3; k++) f[k] = c->frame[k] * lf[0] + c->frame[3 + k] * lf[1] + c->frame[6 + k] * lf[2];
    int bb[2] = {m->geom_bodyid[c->geom1], m->geom_bodyid[c->geom2]};
    for (int side = 0; side < 2; side++) {
      int b = bb[side];
      if (b == 0) continue;
      const double* com = d->subtree_com[body_rootid(m, b)];
      double off[3], tq[3], sg = side ? 1.0 : -1.0;
      for (int k = 0; k < 3; k++) off[k] = c->pos[k] - com[k];
      cross3(tq, off, f);
      for (int k = 0; k < 3; k++) { d->cfrc_ext[b][k] += sg * tq[k]; d->cfrc_ext[b][3 + k] += sg * f[k]; }
    }
  }
}
void orc_set_sensor_terms(OrcEnv* d, int on) {
  d->sensor_terms = on;
  memset(d->cfrc_ext, 0, sizeof d->cfrc_ext);
  memset(d->subtree_linvel, 0, sizeof d->subtree_linvel);
}
void orc_forward(OrcEnv* d) { /* mj_forward */
  kinematics(d); com_pos(d); tendon(d); crb(d); collision(d); make_constraint(d);
  com_vel(d); passive(d); reference_constraint(d); rne_bias(d);
  actuation(d); acceleration(d); fwd_constraint(d);
  if (d->sensor_terms) { subtree_vel(d); rne_post_constraint(d); }
}
/* [UNVERIFIED-vs-3.2.5] mj_Euler: implicit damping (M + h B) qacc = qfrc_smooth + qfrc_constraint whenever a dof has damping and eulerdamp is on; mj_advance: qvel first, then qpos with the NEW qvel; qacc_warmstart = qacc of the unmodified forward pass */
static void euler(OrcEnv* d) { /* mj_Euler with implicit joint damping + mj_advance */
  const B2HModel* m = &d->m;
  int nv = m->nv;
  double h = m->timestep, qacc[NV];
  int damped = 0;
  for (int i = 0; i < nv; i++) if (m->dof_damping[i] > 0) damped = 1;
  if (!damped) memcpy(qacc, d->qacc, sizeof(double) * nv);
  else {
    static __thread double H[NV][NV], L[NV][NV];
    double rhs[NV];
    for (int i = 0; i < nv; i++) for (int j = 0; j < nv; j++) H[i][j] = d->qM[i][j];
    for (int i = 0; i < nv; i++) { H[i][i] += h * m->dof_damping[i]; rhs[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i]; }
    chol_factor(L, H, nv);
    chol_solve(qacc, L, rhs, nv);
  }
  for (int i = 0; i < nv; i++) d->qvel[i] += h * qacc[i];
  for (int j = 0; j < m->njnt; j++) { /* mj_integratePos */
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == B2H_JNT_FREE) {
      for (int k = 0; k < 3; k++) d->qpos[qa + k] += h * d->qvel[da + k];
      double tmp[3] = {d->qvel[da + 3], d->qvel[da + 4], d->qvel[da + 5]}, qrot[4];
      double angle = h * normalize3(tmp); /* mju_quatIntegrate */
      axis_angle2quat(qrot, tmp, angle);
      normalize4(d->qpos + qa + 3);
      mul_quat(d->qpos + qa + 3, d->qpos + qa + 3, qrot);
    } else d->qpos[qa] += h * d->qvel[da];
  }
  d->time += h;
  d->nstep++;
}
void orc_mj_step(OrcEnv* d) { /* mj_step */
  const B2HModel* m = &d->m;
  if (bad_vec(d->qpos, m->nq) || bad_vec(d->qvel, m->nv)) { orc_reset_data(d); d->n_bad++; } /* mj_checkPos/Vel */
  orc_forward(d);
  if (bad_vec(d->qacc, m->nv)) { orc_reset_data(d); d->n_bad++; orc_forward(d); } /* mj_checkAcc */
  euler(d);
}

/* ------------------------------------------------------------------------------------------ env layer */
int orc_obs_dim(const OrcEnv* d) { return (d->m.nq - 2) + d->m.nv + 10 * d->m.nbody + 6 * d->m.nbody + d->m.nv; }
void orc_obs(const OrcEnv* d, double* obs) { /* custom_env.py:242-256 */
  const B2HModel* m = &d->m;
  int o = 0;
  for (int i = 2; i < m->nq; i++) obs[o++] = d->qpos[i];
  for (int i = 0; i < m->nv; i++) obs[o++] = d->qvel[i];
  for (int b = 0; b < m->nbody; b++) for (int k = 0; k < 10; k++) obs[o++] = d->cinert[b][k];
  for (int b = 0; b < m->nbody; b++) for (int k = 0; k < 6; k++) obs[o++] = d->cvel[b][k];
  for (int i = 0; i < m->nv; i++) obs[o++] = d->qfrc_actuator[i];
}
static void quat_to_euler(const double* q, double* roll, double* pitch) { /* utils.py:3-20 */
  double w = q[0], x = q[1], y = q[2], z = q[3];
  *roll = atan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y));
  *pitch = asin(2 * (w * y - z * x));
}
/* cfrc_ext and subtree_linvel are identically zero in the reference (no sensors) */
static double reward_stand(const OrcEnv* d) { /* reward_functions.py:156-211 */
  double h = d->qpos[2], vx = d->qvel[0], roll, pitch;
  quat_to_euler(d->qpos + 3, &roll, &pitch);
  const int nb = d->m.nbody;
  double lf = 0, rf = 0;   /* np.sum(np.abs(cfrc_ext[-2])), [-1]: identically zero unless sensor_terms */
  for (int k = 0; k < 6; k++) { lf += fabs(d->cfrc_ext[nb - 2][k]); rf += fabs(d->cfrc_ext[nb - 1][k]); }
  if (h < 0.8) return 0.0;
  double vr = exp(-2.0 * (vx - 1.0) * (vx - 1.0));
  double hr = exp(-2.0 * (h - 1.282) * (h - 1.282));
  double orr = exp(-3.0 * (roll * roll + pitch * pitch));
  double posture = 0.5 * hr + 0.5 * orr;
  double s = 0;
  for (int a = 0; a < d->m.nu; a++) s += d->ctrl[a] * d->ctrl[a];
  double torque = exp(-0.05 * s);
  double total = lf + rf + 1e-8;
  double foot = 1.0 - fmin(lf, rf) / total;
  return 0.4 * vr + 0.3 * posture + 0.2 * foot + 0.1 * torque;
}
static double reward_walk(const OrcEnv* d) { /* reward_functions.py:213-261 */
  double h = d->qpos[2], vx = d->qvel[0], roll, pitch;
  quat_to_euler(d->qpos + 3, &roll, &pitch);
  if (h < 0.8) return 0.1 * h / 0.8;
  double vr = exp(-0.5 * (vx - 10.0) * (vx - 10.0));
  double hr = exp(-2.0 * (h - 1.282) * (h - 1.282));
  double orr = exp(-3.0 * (roll * roll + pitch * pitch));
  double posture = 0.5 * hr + 0.5 * orr;
  double s = 0;
  for (int a = 0; a < d->m.nu; a++) s += d->ctrl[a] * d->ctrl[a];
  return vr + posture * exp(-0.05 * s);
}
/* params: target_height,min_height,max_roll_pitch,com_radius,energy_w,posture_w,com_w,foot_w,alive_w */
static double reward_kneeling(const OrcEnv* d, const double* p) { /* reward_functions.py:66-154 */
  double h = d->qpos[2];
  if (h < p[1]) return h * h;
  double roll, pitch;
  quat_to_euler(d->qpos + 3, &roll, &pitch);
  double oerr = (roll * roll + pitch * pitch) / (p[2] * p[2]);
  double posture = 0.7 * exp(-5.0 * oerr) + 0.3 * exp(-5.0 * (h - p[0]) * (h - p[0]));
  const double* com = d->subtree_com[0];
  double dist = sqrt(com[0] * com[0] + com[1] * com[1]);
  const double* lv = d->subtree_linvel[0];
  double com_score = 0.7 * exp(-10.0 * (dist / p[3])) + 0.3 * exp(-0.1 * (lv[0] * lv[0] + lv[1] * lv[1] + lv[2] * lv[2]));
  const int nb = d->m.nbody;
  double lf = 0, rf = 0;
  for (int k = 0; k < 6; k++) { lf += fabs(d->cfrc_ext[nb - 2][k]); rf += fabs(d->cfrc_ext[nb - 1][k]); }
  double foot_balance = fmin(lf, rf) / (lf + rf + 1e-8);
  double power = 0;
  for (int i = 6; i < d->m.nv; i++) { double t = d->qfrc_actuator[i] * d->qvel[i]; power += t * t; }
  double energy = exp(-0.01 * power);
  double alive = 1.0 - exp(-0.5 * d->time);
  return p[5] * posture + p[6] * com_score + p[7] * foot_balance + p[4] * energy + p[8] * alive;
}

/* reward of `reward_type` on an explicit mjData-like record (golden-vector checks against reward_functions.py) */
double orc_reward_eval(OrcEnv* d, int reward_type, const double* kneel, const double* qpos, const double* qvel,
                       const double* ctrl, const double* qfrc_actuator, const double* com0, double time) {
  memcpy(d->qpos, qpos, sizeof(double) * d->m.nq);
  memcpy(d->qvel, qvel, sizeof(double) * d->m.nv);
  memcpy(d->ctrl, ctrl, sizeof(double) * d->m.nu);
  memcpy(d->qfrc_actuator, qfrc_actuator, sizeof(double) * d->m.nv);
  memcpy(d->subtree_com[0], com0, sizeof(double) * 3);
  d->time = time;
  if (reward_type == B2H_REWARD_STAND) return reward_stand(d);
  if (reward_type == B2H_REWARD_KNEELING) return reward_kneeling(d, kneel);
  return reward_walk(d);
}

/* HumanoidEnv.reset (custom_env.py:97-150) with explicit noise [nq+nv] (the reference draws it from the
 * global numpy RNG: U(-0.01,0.01), pos first then vel); the z/quaternion masking is applied here. */
void orc_env_reset(OrcEnv* d, const double* noise, double* obs) {
  const B2HModel* m = &d->m;
  orc_reset_data(d);
  double pn[NQ];
  memcpy(pn, noise, sizeof(double) * m->nq);
  pn[2] *= 0.1;
  pn[3] = pn[4] = pn[5] = pn[6] = 0;
  for (int i = 0; i < m->nq; i++) d->qpos[i] = m->qpos0[i] + pn[i];
  d->qpos[2] = 1.282 + pn[2]; d->qpos[3] = 1; d->qpos[4] = d->qpos[5] = d->qpos[6] = 0; /* init_qpos, custom_env.py:58-61 */
  for (int i = 0; i < m->nv; i++) d->qvel[i] = noise[m->nq + i];
  orc_mj_step(d);
  if (obs) orc_obs(d, obs);
  d->step_count = 0;
  d->total_reward = 0;
}
/* HumanoidEnv.step (custom_env.py:152-230) */
void orc_env_step(OrcEnv* d, const float* action, int frame_skip, double duration, int reward_type, int max_steps,
                  const double* kneel_params, double* obs, double* reward, uint8_t* terminated, uint8_t* truncated) {
  const B2HModel* m = &d->m;
  d->step_count += 1;
  for (int s = 0; s < frame_skip; s++) {
    for (int a = 0; a < m->nu; a++) d->ctrl[a] = (double)action[a];
    orc_mj_step(d);
  }
  if (obs) orc_obs(d, obs);
  double r;
  int trunc = 0;
  if (d->step_count >= max_steps) { trunc = 1; r = 0.0; }
  else if (reward_type == B2H_REWARD_STAND) r = reward_stand(d);
  else if (reward_type == B2H_REWARD_KNEELING) r = reward_kneeling(d, kneel_params);
  else r = reward_walk(d);
  d->total_reward += r;
  *reward = r;
  *terminated = d->time >= duration;
  *truncated = (uint8_t)trunc;
}

/* ------------------------------------------------------------------------------------------ API */
OrcEnv* orc_create(const B2HModel* m) {
  OrcEnv* d = (OrcEnv*)calloc(1, sizeof(OrcEnv));
  if (!d) return NULL;
  d->m = *m;
  orc_reset_data(d);
  return d;
}
void orc_destroy(OrcEnv* d) { free(d); }
size_t orc_sizeof_model(void) { return sizeof(B2HModel); }
void orc_set_state(OrcEnv* d, const double* qpos, const double* qvel, const double* warm, int nstep, int step_count) {
  if (qpos) memcpy(d->qpos, qpos, sizeof(double) * d->m.nq);
  if (qvel) memcpy(d->qvel, qvel, sizeof(double) * d->m.nv);
  if (warm) memcpy(d->qacc_warmstart, warm, sizeof(double) * d->m.nv);
  if (nstep >= 0) { d->nstep = nstep; d->time = nstep * d->m.timestep; }
  if (step_count >= 0) d->step_count = step_count;
}
void orc_get_state(const OrcEnv* d, double* qpos, double* qvel, double* warm, int* nstep, int* step_count, double* total_reward) {
  if (qpos) memcpy(qpos, d->qpos, sizeof(double) * d->m.nq);
  if (qvel) memcpy(qvel, d->qvel, sizeof(double) * d->m.nv);
  if (warm) memcpy(warm, d->qacc_warmstart, sizeof(double) * d->m.nv);
  if (nstep) *nstep = d->nstep;
  if (step_count) *step_count = d->step_count;
  if (total_reward) *total_reward = d->total_reward;
}
void orc_set_ctrl(OrcEnv* d, const double* ctrl) { memcpy(d->ctrl, ctrl, sizeof(double) * d->m.nu); }

#define OUT(ptr, count) do { int n_ = (count); if (n_ > max_out) return -1; memcpy(out, (ptr), sizeof(double) * n_); return n_; } while (0)
/* copy a named intermediate (row-major, doubles) */
int orc_get(const OrcEnv* d, const char* what, double* out, int max_out) {
  const B2HModel* m = &d->m;
  int nb = m->nbody, nv = m->nv;
  double tmp[MAXEFC * NV > MAXCON * 9 ? MAXEFC * NV : MAXCON * 9];
  int n = 0;
  if (!strcmp(what, "xpos")) { for (int b = 0; b < nb; b++) for (int k = 0; k < 3; k++) tmp[n++] = d->xpos[b][k]; OUT(tmp, n); }
  if (!strcmp(what, "xquat")) { for (int b = 0; b < nb; b++) for (int k = 0; k < 4; k++) tmp[n++] = d->xquat[b][k]; OUT(tmp, n); }
  if (!strcmp(what, "xmat")) { for (int b = 0; b < nb; b++) for (int k = 0; k < 9; k++) tmp[n++] = d->xmat[b][k]; OUT(tmp, n); }
  if (!strcmp(what, "xipos")) { for (int b = 0; b < nb; b++) for (int k = 0; k < 3; k++) tmp[n++] = d->xipos[b][k]; OUT(tmp, n); }
  if (!strcmp(what, "geom_xpos")) { for (int g = 0; g < m->ngeom; g++) for (int k = 0; k < 3; k++) tmp[n++] = d->geom_xpos[g][k]; OUT(tmp, n); }
  if (!strcmp(what, "geom_xmat")) { for (int g = 0; g < m->ngeom; g++) for (int k = 0; k < 9; k++) tmp[n++] = d->geom_xmat[g][k]; OUT(tmp, n); }
  if (!strcmp(what, "subtree_com")) { for (int b = 0; b < nb; b++) for (int k = 0; k < 3; k++) tmp[n++] = d->subtree_com[b][k]; OUT(tmp, n); }
  if (!strcmp(what, "cinert")) { for (int b = 0; b < nb; b++) for (int k = 0; k < 10; k++) tmp[n++] = d->cinert[b][k]; OUT(tmp, n); }
  if (!strcmp(what, "cvel")) { for (int b = 0; b < nb; b++) for (int k = 0; k < 6; k++) tmp[n++] = d->cvel[b][k]; OUT(tmp, n); }
  if (!strcmp(what, "cdof")) { for (int i = 0; i < nv; i++) for (int k = 0; k < 6; k++) tmp[n++] = d->cdof[i][k]; OUT(tmp, n); }
  if (!strcmp(what, "cdof_dot")) { for (int i = 0; i < nv; i++) for (int k = 0; k < 6; k++) tmp[n++] = d->cdof_dot[i][k]; OUT(tmp, n); }
  if (!strcmp(what, "qM")) { for (int i = 0; i < nv; i++) for (int k = 0; k < nv; k++) tmp[n++] = d->qM[i][k]; OUT(tmp, n); }
  if (!strcmp(what, "ten_length")) OUT(d->ten_length, m->ntendon);
  if (!strcmp(what, "cfrc_ext")) OUT(d->cfrc_ext[0], 6 * m->nbody);
  if (!strcmp(what, "subtree_linvel")) OUT(d->subtree_linvel[0], 3 * m->nbody);
  if (!strcmp(what, "qfrc_passive")) OUT(d->qfrc_passive, nv);
  if (!strcmp(what, "qfrc_bias")) OUT(d->qfrc_bias, nv);
  if (!strcmp(what, "qfrc_actuator")) OUT(d->qfrc_actuator, nv);
  if (!strcmp(what, "qfrc_smooth")) OUT(d->qfrc_smooth, nv);
  if (!strcmp(what, "qacc_smooth")) OUT(d->qacc_smooth, nv);
  if (!strcmp(what, "qacc")) OUT(d->qacc, nv);
  if (!strcmp(what, "qfrc_constraint")) OUT(d->qfrc_constraint, nv);
  if (!strcmp(what, "qacc_warmstart")) OUT(d->qacc_warmstart, nv);
  if (!strcmp(what, "ncon")) { tmp[0] = d->ncon; OUT(tmp, 1); }
  if (!strcmp(what, "nefc")) { tmp[0] = d->nefc; OUT(tmp, 1); }
  if (!strcmp(what, "solver_niter")) { tmp[0] = d->solver_niter; OUT(tmp, 1); }
  if (!strcmp(what, "solver_cost")) { tmp[0] = d->solver_cost; OUT(tmp, 1); }
  if (!strcmp(what, "n_bad")) { tmp[0] = d->n_bad; OUT(tmp, 1); }
  if (!strcmp(what, "time")) { tmp[0] = d->time; OUT(tmp, 1); }
  if (!strcmp(what, "contact_dist")) { for (int c = 0; c < d->ncon; c++) tmp[n++] = d->con[c].dist; OUT(tmp, n); }
  if (!strcmp(what, "contact_pos")) { for (int c = 0; c < d->ncon; c++) for (int k = 0; k < 3; k++) tmp[n++] = d->con[c].pos[k]; OUT(tmp, n); }
  if (!strcmp(what, "contact_frame")) { for (int c = 0; c < d->ncon; c++) for (int k = 0; k < 9; k++) tmp[n++] = d->con[c].frame[k]; OUT(tmp, n); }
  if (!strcmp(what, "contact_pair")) { for (int c = 0; c < d->ncon; c++) tmp[n++] = d->con[c].pair; OUT(tmp, n); }
  if (!strcmp(what, "efc_J")) { for (int i = 0; i < d->nefc; i++) for (int k = 0; k < nv; k++) tmp[n++] = d->efc_J[i][k]; OUT(tmp, n); }
  if (!strcmp(what, "efc_pos")) OUT(d->efc_pos, d->nefc);
  if (!strcmp(what, "efc_R")) OUT(d->efc_R, d->nefc);
  if (!strcmp(what, "efc_D")) OUT(d->efc_D, d->nefc);
  if (!strcmp(what, "efc_vel")) OUT(d->efc_vel, d->nefc);
  if (!strcmp(what, "efc_aref")) OUT(d->efc_aref, d->nefc);
  if (!strcmp(what, "efc_force")) OUT(d->efc_force, d->nefc);
  if (!strcmp(what, "efc_imp")) { for (int i = 0; i < d->nefc; i++) tmp[n++] = d->efc_KBIP[i][2]; OUT(tmp, n); }
  if (!strcmp(what, "efc_type")) { for (int i = 0; i < d->nefc; i++) tmp[n++] = d->efc_type[i]; OUT(tmp, n); }
  if (!strcmp(what, "efc_id")) { for (int i = 0; i < d->nefc; i++) tmp[n++] = d->efc_id[i]; OUT(tmp, n); }
  return -2;
}

/* SubprocVecEnv semantics over n independent envs, optionally on several host threads:
 * worker step + auto-reset with terminal_observation (SB3 2.3.2 subproc_vec_env.py _worker). */
typedef struct {
  OrcEnv** envs; int lo, hi;
  const float* actions; const double* reset_noise;
  int frame_skip, reward_type, max_steps; double duration; const double* kneel;
  double *obs, *reward, *terminal_obs; uint8_t *done, *terminated, *truncated;
  int obs_dim, nu, nqv, nsteps;
} VecJob;
static void* vec_worker(void* arg) {
  VecJob* j = (VecJob*)arg;
  for (int s = 0; s < j->nsteps; s++)
    for (int e = j->lo; e < j->hi; e++) {
      double* obs = j->obs + (size_t)e * j->obs_dim;
      orc_env_step(j->envs[e], j->actions + (size_t)e * j->nu, j->frame_skip, j->duration, j->reward_type, j->max_steps,
                   j->kneel, obs, j->reward + e, j->terminated + e, j->truncated + e);
      j->done[e] = j->terminated[e] || j->truncated[e];
      if (j->done[e]) {
        if (j->terminal_obs) memcpy(j->terminal_obs + (size_t)e * j->obs_dim, obs, sizeof(double) * j->obs_dim);
        orc_env_reset(j->envs[e], j->reset_noise + (size_t)e * j->nqv, obs);
      }
    }
  return NULL;
}
/* nsteps > 1 repeats the same actions (CPU-baseline timing only) */
void orc_vec_step(OrcEnv** envs, int n, const float* actions, const double* reset_noise, int frame_skip, double duration,
                  int reward_type, int max_steps, const double* kneel, double* obs, double* reward, uint8_t* done,
                  uint8_t* terminated, uint8_t* truncated, double* terminal_obs, int nthreads, int nsteps) {
  if (n <= 0) return;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > n) nthreads = n;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  VecJob jobs[256];
  for (int t = 0; t < nthreads; t++) {
    VecJob* j = &jobs[t];
    j->envs = envs; j->lo = (int)((long)n * t / nthreads); j->hi = (int)((long)n * (t + 1) / nthreads);
    j->actions = actions; j->reset_noise = reset_noise; j->frame_skip = frame_skip; j->reward_type = reward_type;
    j->max_steps = max_steps; j->duration = duration; j->kneel = kneel; j->obs = obs; j->reward = reward;
    j->terminal_obs = terminal_obs; j->done = done; j->terminated = terminated; j->truncated = truncated;
    j->obs_dim = orc_obs_dim(envs[0]); j->nu = envs[0]->m.nu; j->nqv = envs[0]->m.nq + envs[0]->m.nv; j->nsteps = nsteps;
    if (nthreads == 1) vec_worker(j);
    else pthread_create(&th[t], NULL, vec_worker, j);
  }
  if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}

/* RolloutBuffer.compute_returns_and_advantage (SB3 2.3.2 common/buffers.py), float32 arithmetic, [T,E] */
void orc_gae(const float* rewards, const float* values, const float* episode_starts, const float* last_values,
             const uint8_t* dones, double gamma_d, double gae_lambda_d, int T, int E, float* advantages, float* returns) {
  /* numpy: python-float scalars are weak, so gamma and gamma*gae_lambda (a double product) round to float32 once */
  const float gamma = (float)gamma_d, gl = (float)(gamma_d * gae_lambda_d);
  for (int e = 0; e < E; e++) {
    float last_gae = 0.0f;
    for (int t = T - 1; t >= 0; t--) {
      float nnt, nv;
      if (t == T - 1) { nnt = 1.0f - (float)dones[e]; nv = last_values[e]; }
      else { nnt = 1.0f - episode_starts[(size_t)(t + 1) * E + e]; nv = values[(size_t)(t + 1) * E + e]; }
      float delta = rewards[(size_t)t * E + e] + gamma * nv * nnt - values[(size_t)t * E + e];
      last_gae = delta + gl * nnt * last_gae;
      advantages[(size_t)t * E + e] = last_gae;
    }
    for (int t = 0; t < T; t++) returns[(size_t)t * E + e] = advantages[(size_t)t * E + e] + values[(size_t)t * E + e];
  }
}
