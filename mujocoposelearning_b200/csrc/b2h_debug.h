// b2h_debug.h — named views into a finished mj_forward (parity/debug only; b2h_debug_forward in include/b2h.h).
#pragma once
#include <string.h>

#include "b2h_physics.cuh"

namespace b2h {

template <typename T>
struct DebugDump {
  Scratch<T> S;
  T lane[32][8];   // per-lane: qfrc_bias, qfrc_smooth, qacc_smooth, qacc, qfrc_constraint, qfrc_actuator, limit D, limit aref
  StepStats stats;
};

// Copies the array called `what` (mjData naming) as doubles; returns the count or <0.
template <typename T>
int extract_named(const DevModel<T>& m, const DebugDump<T>& d, const char* what, double* out, int max_out) {
  const Scratch<T>& S = d.S;
  int n = 0;
  auto put = [&](T v) { if (n < max_out) out[n] = (double)v; n++; };
  auto lanes = [&](int k, int cnt) { for (int i = 0; i < cnt; i++) put(d.lane[i][k]); };
  if (!strcmp(what, "xpos")) for (int i = 0; i < 3 * m.nbody; i++) put(S.xpos[i]);
  else if (!strcmp(what, "xmat")) for (int i = 0; i < 9 * m.nbody; i++) put(S.xmat[i]);
  else if (!strcmp(what, "xipos")) for (int i = 0; i < 3 * m.nbody; i++) put(S.xipos[i]);
  else if (!strcmp(what, "com")) for (int i = 0; i < 3; i++) put(S.com[i]);
  else if (!strcmp(what, "cinert")) for (int i = 0; i < 10 * m.nbody; i++) put(S.cinert[i]);
  else if (!strcmp(what, "cvel")) for (int i = 0; i < 6 * m.nbody; i++) put(S.cvel[i]);
  else if (!strcmp(what, "cdof")) for (int i = 0; i < 6 * m.nv; i++) put(S.cdof[i]);
  else if (!strcmp(what, "cdof_dot")) for (int i = 0; i < 6 * m.nv; i++) put(S.cdofdot[i]);
  else if (!strcmp(what, "xanchor")) for (int i = 0; i < 3 * m.njnt; i++) put(S.xanchor[i]);
  else if (!strcmp(what, "xaxis")) for (int i = 0; i < 3 * m.njnt; i++) put(S.xaxis[i]);
  else if (!strcmp(what, "geom_xpos")) for (int i = 0; i < 3 * m.ngeom; i++) put(S.gpos[i]);
  else if (!strcmp(what, "geom_zaxis")) for (int i = 0; i < 3 * m.ngeom; i++) put(S.gaxis[i]);
  else if (!strcmp(what, "qM")) { for (int i = 0; i < m.nv; i++) for (int j = 0; j < m.nv; j++) put(S.M[i * LD + j]); }
  else if (!strcmp(what, "qfrc_bias")) lanes(0, m.nv);
  else if (!strcmp(what, "qfrc_smooth")) lanes(1, m.nv);
  else if (!strcmp(what, "qacc_smooth")) lanes(2, m.nv);
  else if (!strcmp(what, "qacc")) lanes(3, m.nv);
  else if (!strcmp(what, "qfrc_constraint")) lanes(4, m.nv);
  else if (!strcmp(what, "qfrc_actuator")) lanes(5, m.nv);
  else if (!strcmp(what, "limit_D")) lanes(6, m.nv);
  else if (!strcmp(what, "limit_aref")) lanes(7, m.nv);
  else if (!strcmp(what, "ncon")) put((T)d.stats.ncon);
  else if (!strcmp(what, "nrow")) put((T)d.stats.nrow);
  else if (!strcmp(what, "nefc")) put((T)(d.stats.nrow + d.stats.nlimit));
  else if (!strcmp(what, "solver_niter")) put((T)d.stats.niter);
  else if (!strcmp(what, "contact_dist")) for (int i = 0; i < d.stats.ncon; i++) put(S.con_dist[i]);
  else if (!strcmp(what, "contact_pos")) for (int i = 0; i < 3 * d.stats.ncon; i++) put(S.con_pos[i]);
  else if (!strcmp(what, "contact_frame")) for (int i = 0; i < 9 * d.stats.ncon; i++) put(S.con_frame[i]);
  else if (!strcmp(what, "efc_J_dense")) { for (int r = 0; r < d.stats.nrow; r++) for (int j = 0; j < m.nv; j++) put(S.J[r * LD + j]); }
  else return -2;
  return n <= max_out ? n : -1;
}

}  // namespace b2h
