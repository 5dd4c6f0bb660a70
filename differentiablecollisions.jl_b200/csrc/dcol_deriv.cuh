// Derivative epilogues of the proximity kernels (FP64, sm_100a), fused behind the solver loop.
//
// Reference: proximity_gradient = d/dtheta [ z'(G(theta) x - h(theta)) ]  (src/proximity_gradient.jl:1-78)
//            proximity_jacobian = implicit-function theorem on the KKT residual
//                                 dx = (G' S^-1 Z G)^-1 (r1 - G' S^-1 r2)      (src/proximity.jl:27-134)
// with theta = [r1; att1; r2; att2], differentiated THROUGH normalize(q)
// (src/problem_matrices.jl:10) as ForwardDiff does.
//
// In the body-frame formulation of dcol_common.cuh only x~_k = Q_k'(x - r~_k) depends on the pose, so
//   u_k(theta) = C_k [x~_k; alpha; e_k],   d u_k = C_k[:,0:3] d x~_k,   G'z = -sum_k T_k' (C_k' z_k)
// and every pose derivative reduces to contractions of dQn/d(att_j) (3x3, obtained with a
// one-partial dual number pushed through dcm_from_q / dcm_from_mrp) with small 3x3 matrices:
//   grad[r_k]   = Q_k wx_k                       wx_k = (C_k' z_k)[0:3]
//   grad[att_j] = < dQn_j , -v_k a_k' + (Q_k wx_k) r_off' >      v_k = x - r~_k,  a_k = Q_off wx_k
//   J[:, r_k]   = K_k Q_k'                       K_k  = P4 (A'^-1 T_k'' N_k[:,0:3])[0:4]
//   J[i, att_j] = < dQn_j , -v_k (K_k Q_off')[i,:] + (K_k Q_k')[i,:]' r_off' + Hw[i,:]' a_k' >
// where N_k = C_k' (S_k^-1 Z_k) C_k, A' = sum_k T_k'' N_k T_k' is the reduced KKT matrix in
// primitive-1's frame, P4 = blkdiag(Q_1, 1), Hw = P4 (A'^-1)[0:4,0:3] Q_1'.  Only the first four
// rows of A'^-1 are needed, so the transposed system A'' Y = [e1..e4] is solved once (4 right-hand
// sides instead of 14).
#pragma once
#include "dcol_common.cuh"

namespace dcol {

// ---------------------------------------------------------------------------
// branch-free FP64 reciprocal / reciprocal square root: hardware seed (MUFU.RCP64H / MUFU.RSQ64H,
// ~20 good bits) + one cubically convergent step.  No slow-path calls: operands here are positive, normal
// numbers (failures are flagged by explicit `ok` tests before the result is used), and the
// straight-line form keeps the solver passes in large basic blocks the scheduler can interleave.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);            // seed error e ~ 2^-20
  return fma(y, fma(e, e, e), y);              // y (1 + e + e^2): cubic step, error e^3 ~ 2^-60
}
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x * y, y, 1.0);        // 1 - x y^2
  const double p = fma(e, 0.375, 0.5) * e;     // e/2 + 3 e^2 / 8
  return fma(y, p, y);                         // cubic step
}
// ---- one-partial forward-mode dual number -------------------------------------------------
struct Dual1 { double v, d; };
__device__ __forceinline__ Dual1 operator+(const Dual1& a, const Dual1& b) { return {a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ Dual1 operator-(const Dual1& a, const Dual1& b) { return {a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ Dual1 operator*(const Dual1& a, const Dual1& b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
__device__ __forceinline__ Dual1 operator/(const Dual1& a, const Dual1& b) { double inv = fast_rcp(b.v); double q = a.v * inv; return {q, (a.d - q * b.d) * inv}; }
__device__ __forceinline__ Dual1 operator+(const Dual1& a, double b) { return {a.v + b, a.d}; }
__device__ __forceinline__ Dual1 operator-(const Dual1& a, double b) { return {a.v - b, a.d}; }
__device__ __forceinline__ Dual1 operator-(double b, const Dual1& a) { return {b - a.v, -a.d}; }
__device__ __forceinline__ Dual1 operator*(double b, const Dual1& a) { return {a.v * b, a.d * b}; }
__device__ __forceinline__ Dual1 operator*(const Dual1& a, double b) { return {a.v * b, a.d * b}; }
__device__ __forceinline__ Dual1 operator/(double b, const Dual1& a) { double inv = fast_rcp(a.v); double q = b * inv; return {q, -q * a.d * inv}; }
// 1/sqrt(a): the attitude maps only ever divide by the norm, so they take the reciprocal square root (hardware seed + one
// cubic step, ~2^-60 relative) and multiply; the IEEE sqrt + four divisions were ~150 instructions per frame
template <class T> __device__ __forceinline__ T t_rsqrt(const T& a);
template <> __device__ __forceinline__ double t_rsqrt<double>(const double& a) { return fast_rsqrt(a); }
template <> __device__ __forceinline__ Dual1 t_rsqrt<Dual1>(const Dual1& a) { double r = fast_rsqrt(a.v); return {r, -0.5 * a.d * r * r * r}; }
template <class T> __device__ __forceinline__ T t_rcp(const T& a);
template <> __device__ __forceinline__ double t_rcp<double>(const double& a) { return fast_rcp(a); }
template <> __device__ __forceinline__ Dual1 t_rcp<Dual1>(const Dual1& a) { double r = fast_rcp(a.v); return {r, -a.d * r * r}; }

// ---- attitude maps (src/problem_matrices.jl:9-27), templated on the scalar ------------------
template <class T>
__device__ __forceinline__ void dcm_from_q(const T q[4], T Q[3][3]) {
  T inrm = t_rsqrt<T>(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  T q4 = q[0] * inrm, q1 = q[1] * inrm, q2 = q[2] * inrm, q3 = q[3] * inrm;
  Q[0][0] = 2.0 * q1 * q1 + 2.0 * q4 * q4 - 1.0; Q[0][1] = 2.0 * (q1 * q2 - q3 * q4);         Q[0][2] = 2.0 * (q1 * q3 + q2 * q4);
  Q[1][0] = 2.0 * (q1 * q2 + q3 * q4);         Q[1][1] = 2.0 * q2 * q2 + 2.0 * q4 * q4 - 1.0; Q[1][2] = 2.0 * (q2 * q3 - q1 * q4);
  Q[2][0] = 2.0 * (q1 * q3 - q2 * q4);         Q[2][1] = 2.0 * (q2 * q3 + q1 * q4);         Q[2][2] = 2.0 * q3 * q3 + 2.0 * q4 * q4 - 1.0;
}
template <class T>
__device__ __forceinline__ void dcm_from_mrp(const T p[3], T Q[3][3]) {
  T p1 = p[0], p2 = p[1], p3 = p[2];
  T den1 = p1 * p1 + p2 * p2 + p3 * p3 + 1.0;
  T den = den1 * den1;
  T a = 4.0 * p1 * p1 + 4.0 * p2 * p2 + 4.0 * p3 * p3 - 4.0;
  T id = t_rcp<T>(den);
  Q[0][0] = 1.0 - (8.0 * p2 * p2 + 8.0 * p3 * p3) * id; Q[0][1] = (8.0 * p1 * p2 + p3 * a) * id; Q[0][2] = (8.0 * p1 * p3 - p2 * a) * id;
  Q[1][0] = (8.0 * p1 * p2 - p3 * a) * id; Q[1][1] = 1.0 - (8.0 * p1 * p1 + 8.0 * p3 * p3) * id; Q[1][2] = (8.0 * p2 * p3 + p1 * a) * id;
  Q[2][0] = (8.0 * p1 * p3 + p2 * a) * id; Q[2][1] = (8.0 * p2 * p3 - p1 * a) * id; Q[2][2] = 1.0 - (8.0 * p1 * p1 + 8.0 * p2 * p2) * id;
}

// dQn / d att_j  (3x3)
__device__ __forceinline__ void dcm_partial(const double* att, int att_kind, int j, double (&dQ)[3][3]) {
  Dual1 Q[3][3];
  if (att_kind == 0) {
    Dual1 q[4];
#pragma unroll
    for (int i = 0; i < 4; i++) q[i] = Dual1{att[i], (i == j) ? 1.0 : 0.0};
    dcm_from_q<Dual1>(q, Q);
  } else {
    Dual1 q[3];
#pragma unroll
    for (int i = 0; i < 3; i++) q[i] = Dual1{att[i], (i == j) ? 1.0 : 0.0};
    dcm_from_mrp<Dual1>(q, Q);
  }
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) dQ[a][b] = Q[a][b].d;
}

struct Frame { double Q[3][3]; double r[3]; };

// lambda \ w in the Jordan algebra of the second-order cone (soc_utils.jl:52-67)
template <int D>
__device__ __forceinline__ void soc_jdiv(const double (&lam)[D], const double (&w)[D], double (&out)[D]) {
  double rho = lam[0] * lam[0], nu = 0.0;
#pragma unroll
  for (int i = 1; i < D; i++) { rho -= lam[i] * lam[i]; nu += lam[i] * w[i]; }
  double irho = fast_rcp(rho), il0 = fast_rcp(lam[0]);
  out[0] = irho * (lam[0] * w[0] - nu);
#pragma unroll
  for (int i = 1; i < D; i++) out[i] = irho * ((nu * il0 - w[0]) * lam[i] + (rho * il0) * w[i]);
}

// Gaussian elimination with partial (row) pivoting, NR right-hand sides, fully unrolled (registers).
// The reference solves the same reduced KKT system G'S^-1 Z G with StaticArrays' pivoted LU (src/proximity.jl:100);
// its condition number reaches 1e7 on pairs close to degeneracy, where elimination in natural order loses digits.
// The pivot search is a chain of predicated row exchanges (no dynamic indexing, so the arrays stay in registers):
// after comparing row k with every row below it, row k holds the entry of largest magnitude in column k.
template <int N, int NR>
__device__ __forceinline__ bool ge_solve(double (&A)[N][N], double (&B)[N][NR]) {
  bool ok = true;
#pragma unroll
  for (int k = 0; k < N; k++) {
#pragma unroll
    for (int i = k + 1; i < N; i++) {
#ifdef DCOL_NO_PIVOT
      const bool sw = false;                       // measurement build: elimination in natural order (round 1)
#else
      const bool sw = fabs(A[i][k]) > fabs(A[k][k]);
#endif
#pragma unroll
      for (int j = k; j < N; j++) { const double a = A[k][j], b = A[i][j]; A[k][j] = sw ? b : a; A[i][j] = sw ? a : b; }
#pragma unroll
      for (int j = 0; j < NR; j++) { const double a = B[k][j], b = B[i][j]; B[k][j] = sw ? b : a; B[i][j] = sw ? a : b; }
    }
    double piv = A[k][k];
    ok = ok && (fabs(piv) > 0.0) && (piv == piv);
    double ip = fast_rcp(piv);
#pragma unroll
    for (int i = k + 1; i < N; i++) {
      double f = A[i][k] * ip;
#pragma unroll
      for (int j = k + 1; j < N; j++) A[i][j] -= f * A[k][j];
#pragma unroll
      for (int j = 0; j < NR; j++) B[i][j] -= f * B[k][j];
    }
    A[k][k] = ip;
  }
#pragma unroll
  for (int i = N - 1; i >= 0; i--) {
#pragma unroll
    for (int j = 0; j < NR; j++) {
      double a = B[i][j];
#pragma unroll
      for (int k = i + 1; k < N; k++) a -= A[i][k] * B[k][j];
      B[i][j] = a * A[i][i];
    }
  }
  return ok;
}

// Pose derivatives of one primitive.  The reference states the "ball" SOC blocks (capsule, cylinder,
// sphere, polygon: ||x - r~ - Q B~ e|| <= R alpha) in WORLD coordinates and everything else (orthant
// rows, cone and ellipsoid SOC blocks) in BODY coordinates; the solver here keeps all blocks in body
// coordinates (a rotation of the ball block, which leaves the iterates unchanged), so the two parts
// enter the pose derivatives differently:
//   rest (body):  u = C [Q'(x - r~); ..]     d_j     =  Q_off' dQn_j' v - Q' dQn_j r_off
//   ball (world): u = [R alpha; v - Q B~ e]   delta_j = -Q' dQn_j (r_off + Q_off B~ e)   (in body coords)
// with v = x - r~ (world).  Inputs: wr = (C' z)[0:3] of the rest part, zb = vector part of the ball
// block's z (body coords), Be = B~ e (body coords), xt = x~ (body coords).
struct SideVecs { double wr[3], zb[3], Be[3], xt[3]; };

// build_frame and grad_side are class-independent and are compiled once per query kind (dcol_shared.cuh, out of
// line): every specialisation calls the same code, which therefore stays resident in the instruction caches while
// the persistent kernel moves from one class pair to the next.  jac_side is inlined: its 4x3 operands would
// otherwise travel through local memory (measured: -11 % Jacobian time on mixed batches).

// r~ = r + Qn*r_offset, Q = Qn*Q_offset  (dispatch wrappers, e.g. src/problem_matrices.jl:40-53)
static __device__ __noinline__ void build_frame(const DevShape* S, const double* r, const double* att, int att_kind, Frame& F);

// gradient columns of one primitive: out[0:3] = d alpha/d r_k, out[3:3+na] = d alpha/d att_k
// (src/proximity_gradient.jl:1-78: -z' du/dtheta)
static __device__ __noinline__ void grad_side(const DevShape* S, const Frame& F, const double* att, int att_kind, const SideVecs& V, double* out);

// Jacobian columns of one primitive, written column-major (4 rows) at out[col*4 + row]:
// columns 0..2 = d/d r_k, 3..3+na = d/d att_k.
//   K  = P4 (A'^-1 T_k'' N_k[:,0:3])[0:4]        (4x3, N_k = rest + ball)
//   Kb = same with the ball part of N_k only      (4x3)
//   Ye = P4 (A'^-1)[0:4, extras of k]             (4xNE; dR1/dtheta has extras rows for the ball block)
//   Hw = P4 (A'^-1)[0:4,0:3] Q_1'                 (4x3)
template <int NE>
__device__ __forceinline__ void jac_side(const DevShape* S, const Frame& F, const double* att, int att_kind,
                                         const double (&K)[4][3], const double (&Kb)[4][3], const double (&Ye)[4][2],
                                         const double (&Hw)[4][3], const SideVecs& V, double* out) {
  const bool off = S->has_offset != 0;
  double KW[4][3], KrQ[4][3], KrW[4][3], KbW[4][3], v[3], a[3], ob[3], gzb[3];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double Kr[3] = {K[i][0] - Kb[i][0], K[i][1] - Kb[i][1], K[i][2] - Kb[i][2]};
#pragma unroll
    for (int p = 0; p < 3; p++) {
      KrW[i][p] = Kr[0] * F.Q[p][0] + Kr[1] * F.Q[p][1] + Kr[2] * F.Q[p][2];
      KbW[i][p] = Kb[i][0] * F.Q[p][0] + Kb[i][1] * F.Q[p][1] + Kb[i][2] * F.Q[p][2];
      KW[i][p] = KrW[i][p] + KbW[i][p];
      KrQ[i][p] = off ? (Kr[0] * S->Q_off[3 * p] + Kr[1] * S->Q_off[3 * p + 1] + Kr[2] * S->Q_off[3 * p + 2]) : Kr[p];
    }
  }
#pragma unroll
  for (int i = 0; i < 3; i++) {
    v[i] = F.Q[i][0] * V.xt[0] + F.Q[i][1] * V.xt[1] + F.Q[i][2] * V.xt[2];
    gzb[i] = F.Q[i][0] * V.zb[0] + F.Q[i][1] * V.zb[1] + F.Q[i][2] * V.zb[2];
    a[i] = off ? (S->Q_off[3 * i] * V.wr[0] + S->Q_off[3 * i + 1] * V.wr[1] + S->Q_off[3 * i + 2] * V.wr[2]) : V.wr[i];
    ob[i] = off ? (S->r_off[i] + S->Q_off[3 * i] * V.Be[0] + S->Q_off[3 * i + 1] * V.Be[1] + S->Q_off[3 * i + 2] * V.Be[2]) : V.Be[i];
  }
  double2* o2 = reinterpret_cast<double2*>(out);
#pragma unroll
  for (int p = 0; p < 3; p++) { o2[2 * p] = make_double2(KW[0][p], KW[1][p]); o2[2 * p + 1] = make_double2(KW[2][p], KW[3][p]); }
  const int na = (att_kind == 0) ? 4 : 3;
  for (int j = 0; j < na; j++) {
    double dQ[3][3];
    dcm_partial(att, att_kind, j, dQ);
    double tv[3], tg[3], da[3], dr[3], dob[3];
#pragma unroll
    for (int q = 0; q < 3; q++) {
      tv[q] = dQ[0][q] * v[0] + dQ[1][q] * v[1] + dQ[2][q] * v[2];        // dQ' v
      tg[q] = dQ[0][q] * gzb[0] + dQ[1][q] * gzb[1] + dQ[2][q] * gzb[2];  // dQ' (Q zb)
    }
#pragma unroll
    for (int p = 0; p < 3; p++) {
      da[p] = dQ[p][0] * a[0] + dQ[p][1] * a[1] + dQ[p][2] * a[2];
      dob[p] = dQ[p][0] * ob[0] + dQ[p][1] * ob[1] + dQ[p][2] * ob[2];      // dQ (r_off + Q_off B~ e)
      dr[p] = off ? (dQ[p][0] * S->r_off[0] + dQ[p][1] * S->r_off[1] + dQ[p][2] * S->r_off[2]) : 0.0;
    }
    double em[2] = {0.0, 0.0};                                              // Q_off[:,m] . tg
#pragma unroll
    for (int m = 0; m < NE; m++) em[m] = off ? (S->Q_off[m] * tg[0] + S->Q_off[3 + m] * tg[1] + S->Q_off[6 + m] * tg[2]) : tg[m];
    double c[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      double acc = Hw[i][0] * da[0] + Hw[i][1] * da[1] + Hw[i][2] * da[2];
      acc -= KrQ[i][0] * tv[0] + KrQ[i][1] * tv[1] + KrQ[i][2] * tv[2];
      acc += KbW[i][0] * dob[0] + KbW[i][1] * dob[1] + KbW[i][2] * dob[2];
      if (off) acc += KrW[i][0] * dr[0] + KrW[i][1] * dr[1] + KrW[i][2] * dr[2];
#pragma unroll
      for (int m = 0; m < NE; m++) acc -= Ye[i][m] * em[m];
      c[i] = acc;
    }
    o2[2 * (3 + j)] = make_double2(c[0], c[1]); o2[2 * (3 + j) + 1] = make_double2(c[2], c[3]);
  }
}


}  // namespace dcol
