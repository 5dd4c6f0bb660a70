// Class-independent device routines shared by every specialisation of a query kind: out of line (one copy per
// translation unit = per query kind), so that they stay resident in the instruction caches while the persistent
// kernel moves from one class pair to the next.
#pragma once
#include "dcol_deriv.cuh"

namespace dcol {

// r~ = r + Qn*r_offset, Q = Qn*Q_offset  (dispatch wrappers, e.g. src/problem_matrices.jl:40-53)
static __device__ __noinline__ void build_frame(const DevShape* S, const double* r, const double* att, int att_kind, Frame& F) {
  double Qn[3][3];
  if (att_kind == 0) { double q[4] = {att[0], att[1], att[2], att[3]}; dcm_from_q<double>(q, Qn); }
  else { double p[3] = {att[0], att[1], att[2]}; dcm_from_mrp<double>(p, Qn); }
  if (S->has_offset) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
      F.r[i] = r[i] + Qn[i][0] * S->r_off[0] + Qn[i][1] * S->r_off[1] + Qn[i][2] * S->r_off[2];
#pragma unroll
      for (int j = 0; j < 3; j++) F.Q[i][j] = Qn[i][0] * S->Q_off[j] + Qn[i][1] * S->Q_off[3 + j] + Qn[i][2] * S->Q_off[6 + j];
    }
  } else {
#pragma unroll
    for (int i = 0; i < 3; i++) { F.r[i] = r[i];
#pragma unroll
      for (int j = 0; j < 3; j++) F.Q[i][j] = Qn[i][j]; }
  }
}

// gradient columns of one primitive: out[0:3] = d alpha/d r_k, out[3:3+na] = d alpha/d att_k
// (src/proximity_gradient.jl:1-78: -z' du/dtheta)
static __device__ __noinline__ void grad_side(const DevShape* S, const Frame& F, const double* att, int att_kind,
                                          const SideVecs& V, double* out) {
  const bool off = S->has_offset != 0;
  double gwr[3], gzb[3], v[3], a[3], ob[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    gwr[i] = F.Q[i][0] * V.wr[0] + F.Q[i][1] * V.wr[1] + F.Q[i][2] * V.wr[2];
    gzb[i] = F.Q[i][0] * V.zb[0] + F.Q[i][1] * V.zb[1] + F.Q[i][2] * V.zb[2];
    v[i] = F.Q[i][0] * V.xt[0] + F.Q[i][1] * V.xt[1] + F.Q[i][2] * V.xt[2];
    a[i] = off ? (S->Q_off[3 * i] * V.wr[0] + S->Q_off[3 * i + 1] * V.wr[1] + S->Q_off[3 * i + 2] * V.wr[2]) : V.wr[i];
    ob[i] = off ? (S->r_off[i] + S->Q_off[3 * i] * V.Be[0] + S->Q_off[3 * i + 1] * V.Be[1] + S->Q_off[3 * i + 2] * V.Be[2]) : V.Be[i];
  }
#pragma unroll
  for (int i = 0; i < 3; i++) out[i] = gwr[i] + gzb[i];
  const int na = (att_kind == 0) ? 4 : 3;
  for (int j = 0; j < na; j++) {
    double dQ[3][3];
    dcm_partial(att, att_kind, j, dQ);
    double acc = 0.0;
#pragma unroll
    for (int p = 0; p < 3; p++) {
      acc -= v[p] * (dQ[p][0] * a[0] + dQ[p][1] * a[1] + dQ[p][2] * a[2]);
      acc += gzb[p] * (dQ[p][0] * ob[0] + dQ[p][1] * ob[1] + dQ[p][2] * ob[2]);
      if (off) acc += gwr[p] * (dQ[p][0] * S->r_off[0] + dQ[p][1] * S->r_off[1] + dQ[p][2] * S->r_off[2]);
    }
    out[3 + j] = acc;
  }
}

}  // namespace dcol
