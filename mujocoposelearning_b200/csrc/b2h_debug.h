// b2h_debug.h — named views into a finished mj_forward (parity/debug only; b2h_debug_forward in include/b2h.h).
#pragma once
#include <string.h>

#include "b2h_physics.cuh"

namespace b2h {

// Copies the array called `what` (mjData naming) as doubles; returns the count or <0.
template <typename T>
int extract_named(const DevModel<T>& m, const DebugDump<T>& d, const char* what, double* out, int max_out) {
  int n = 0;
  auto put = [&](T v) { if (n < max_out) out[n] = (double)v; n++; };
  auto arr = [&](const T* p, int cnt) { for (int i = 0; i < cnt; i++) put(p[i]); };
  auto lanes = [&](int k, int cnt) { for (int i = 0; i < cnt; i++) put(d.lane[i][k]); };
  if (!strcmp(what, "xpos")) arr(d.pos + POS_XPOS, 3 * m.nbody);
  else if (!strcmp(what, "xmat")) arr(d.pos + POS_XMAT, 9 * m.nbody);
  else if (!strcmp(what, "xipos")) arr(d.pos + POS_XIPOS, 3 * m.nbody);
  else if (!strcmp(what, "xanchor")) arr(d.pos + POS_XANCHOR, 3 * m.njnt);
  else if (!strcmp(what, "xaxis")) arr(d.pos + POS_XAXIS, 3 * m.njnt);
  else if (!strcmp(what, "geom_xpos")) arr(d.pos + POS_GPOS, 3 * m.ngeom);
  else if (!strcmp(what, "geom_zaxis")) arr(d.pos + POS_GAXIS, 3 * m.ngeom);
  else if (!strcmp(what, "com")) arr(d.com, 3);
  else if (!strcmp(what, "cinert")) arr(d.cinert, 10 * m.nbody);
  else if (!strcmp(what, "cvel")) arr(d.cvel, 6 * m.nbody);
  else if (!strcmp(what, "cdof")) arr(d.cdof, 6 * m.nv);
  else if (!strcmp(what, "cdof_dot")) arr(d.cdofdot, 6 * m.nv);
  else if (!strcmp(what, "qM")) { for (int i = 0; i < m.nv; i++) for (int j = 0; j < m.nv; j++) put(d.M[i * LD + j]); }
  else if (!strcmp(what, "qfrc_bias")) lanes(0, m.nv);
  else if (!strcmp(what, "qfrc_smooth")) lanes(1, m.nv);
  else if (!strcmp(what, "qacc_smooth")) lanes(2, m.nv);
  else if (!strcmp(what, "qacc")) lanes(3, m.nv);
  else if (!strcmp(what, "qfrc_constraint")) lanes(4, m.nv);
  else if (!strcmp(what, "qfrc_actuator")) lanes(5, m.nv);
  else if (!strcmp(what, "limit_D")) lanes(6, m.nv);
  else if (!strcmp(what, "limit_aref")) lanes(7, m.nv);
  else if (!strcmp(what, "cfrc_ext")) arr(d.cfrc_ext, 6 * m.nbody);
  else if (!strcmp(what, "subtree_linvel0")) arr(d.sub_linvel, 3);
  else if (!strcmp(what, "ncon")) put((T)d.stats.ncon);
  else if (!strcmp(what, "nrow")) put((T)d.stats.nrow);
  else if (!strcmp(what, "nefc")) put((T)(d.stats.nrow + d.stats.nlimit));
  else if (!strcmp(what, "solver_niter")) put((T)d.stats.niter);
  else if (!strcmp(what, "contact_dist")) arr(d.con_dist, d.stats.ncon);
  else if (!strcmp(what, "contact_pos")) arr(d.con_pos, 3 * d.stats.ncon);
  else if (!strcmp(what, "contact_frame")) arr(d.con_frame, 9 * d.stats.ncon);
  else if (!strcmp(what, "efc_J_dense")) { for (int r = 0; r < d.stats.nrow; r++) for (int j = 0; j < m.nv; j++) put(d.J[r * LD + j]); }
  else return -2;
  return n <= max_out ? n : -1;
}

}  // namespace b2h
