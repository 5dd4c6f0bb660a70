// Thread-per-pair primal-dual interior-point solver (FP64, sm_100a).
//
// Restates the algorithm of the reference's solve_socp
// (src/solvers/coneqp/static_solver2.jl:106-191: Mehrotra predictor-corrector,
// Nesterov-Todd scaling NT_scaling_chol_2.jl:95-127, line search to the cone
// boundary :2-51, initial point :52-103, mu-only termination :143-146, 0.99 step
// damping :172, 50 iteration cap :132) in a form designed for the GPU:
//   * the pair's problem is never materialised as dense (G,h): each primitive is a
//     constant block C_k acting on body-frame coordinates (dcol_common.cuh);
//   * orthant rows use the unscaled Newton equations (d_i = z_i/s_i), which are
//     algebraically identical to the NT-scaled ones for the nonnegative orthant;
//   * SOC blocks use the closed form W^-1 = (1/eta) J Wbar J instead of a dense
//     Cholesky of W (NT_scaling_chol_2.jl:126) — same operator, fewer flops;
//   * per-row state (s, z, 1/(s z), rz|ds, p|dz) and the pair's relative pose live in a contiguous,
//     odd-pitch slice of shared memory per thread (conflict-free, field offsets as immediates);
//     everything else (y, KKT matrix, SOC blocks) is in registers;
//   * SOC search directions are accumulated in the NT-scaled space and all reciprocals / rsqrts are
//     branch-free (hardware seed + one cubic step).
// The warps of a CTA advance through the interior-point iterations in lockstep (dcol_persistent.cuh pulls CTA-sized
// tiles from a queue), so that they share instruction-cache lines.
#pragma once
#include "dcol_common.cuh"
#include "dcol_deriv.cuh"
#include <type_traits>

// one block-wide vote per interior-point iteration (the hand-over capacity of a bin assumes CTA-wide votes)
#define DCOL_COUNT(pred) __syncthreads_count(pred)

namespace dcol {

// max / min as one compare and one select.  fmax / fmin on doubles compile to DSETP.MAX + two selects + a NaN-quieting
// LOP3 + moves for their operands (5-7 instructions; 10 % of the interior-point loop went there); NaNs need no
// special care here: a NaN iterate is caught by the mu test of the next iteration.
// (written in PTX: with a constant operand the compiler turns the C++ ternary back into the fmax sequence)
__device__ __forceinline__ double dmax(double a, double b) {
  double r; asm("{ .reg .pred p; setp.gt.f64 p, %1, %2; selp.f64 %0, %1, %2, p; }" : "=d"(r) : "d"(a), "d"(b)); return r;
}
__device__ __forceinline__ double dmin(double a, double b) {
  double r; asm("{ .reg .pred p; setp.lt.f64 p, %1, %2; selp.f64 %0, %1, %2, p; }" : "=d"(r) : "d"(a), "d"(b)); return r;
}
#ifndef DCOL_LEAN_SQRT
#define DCOL_LEAN_SQRT 1
#endif
__device__ __forceinline__ double fast_sqrt(double x) {   // x >= 0 ; sqrt(0) = 0
#if DCOL_LEAN_SQRT
  // hardware seed y ~ 1/sqrt(x), g = x y ~ sqrt(x), e = 1 - x y^2 (|e| ~ 2^-19): sqrt(x) = g (1 - e)^(-1/2) = g (1 + e/2 + 3e^2/8 + O(e^3)),
  // 7 instructions, ~1 ulp (the step lengths it feeds are damped by 0.99; 12 instructions bought the last half ulp before)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x + 0x1p-1000));   // keeps the seed finite at 0; x + 2^-1000 == x for every x >= 2^-947
  const double g = x * y;
  const double e = fma(-g, y, 1.0);
  return fma(g, fma(e, 0.375, 0.5) * e, g);
#else
  double r = fast_rsqrt(x + 0x1p-1000);                    // keeps rsqrt(0) finite; x + 2^-1000 == x for every x >= 2^-947
  double sq = x * r;
  return fma(fma(-sq, sq, x), 0.5 * r, sq);               // one correction step on the product
#endif
}

// ---------------------------------------------------------------------------
// shared-memory row store: field f of row i of this thread
// ---------------------------------------------------------------------------
// The dynamic shared memory of the persistent kernel, declared here (not passed as a pointer through the
// out-of-line entry) so that accesses compile to LDS/STS instead of generic loads and stores.
extern __shared__ double dcol_dyn_smem[];
// Each thread owns a contiguous run of `pitch` doubles (pitch odd, >= rows_cap*NFIELD): field f of row i is at
// [threadIdx.x*pitch + i*NFIELD + f], so the field offset folds into the LDS/STS immediate and only the row index
// costs an address instruction; an odd pitch spreads the 16 lanes of a 64-bit wavefront over all banks.
struct RowStore {
  double* base;
  __device__ __forceinline__ double& operator()(int f, int i) const { return base[POSE_SLOTS + i * NFIELD + f]; }
};
// The relative pose is read 3-4 times per iteration but never written after set-up: it lives in shared memory
// (LDS with immediate offsets) instead of 24 registers that the allocator would otherwise spill to local memory.
struct RelPoseR { const double* p; __device__ __forceinline__ const double* operator[](int a) const { return p + 3 * a; } };
struct RelPoseT { const double* p; __device__ __forceinline__ double operator[](int a) const { return p[9 + a]; } };
// Slots 12..23 keep primitive 1's world frame (Q row-major, r~) for the output / derivative epilogue, which would
// otherwise rebuild it from the pose in global memory after the loop (a dependent L2 round trip per tile).

// ---------------------------------------------------------------------------
// small dense helpers (upper-triangular storage in full arrays; fully unrolled)
// ---------------------------------------------------------------------------
// The reference factors the KKT matrix by Cholesky (static_solver2.jl:92,153).  Here M = U'DU with U UNIT upper
// triangular (right-looking, in place; id = 1/D): the same solution to round-off, but the pivots need a reciprocal
// instead of a reciprocal square root and the triangular sweeps are one dependent FMA per step instead of an FMA and a
// multiply -- the factor-and-solve chain is the longest serial stretch of an iteration, and with two warps per scheduler
// its latency is exposed.  Positive definiteness is still the test "every pivot > 0".
// (nx = 8, polygon-polygon pairs only, keeps the Cholesky form: the right-looking update of a full 8x8 factor next to two
// polygon cone blocks made ptxas spill inside the loop: 138 -> 381 local-memory instructions, +19 % on that class pair)
#ifndef DCOL_LDL
#define DCOL_LDL 1
#endif
#ifndef DCOL_LDL_MAXN
#define DCOL_LDL_MAXN 7
#endif
template <int N>
__device__ __forceinline__ bool chol_upper(double (&M)[N][N], double (&id)[N]) {
  if constexpr (DCOL_LDL && N <= DCOL_LDL_MAXN) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < N; j++) {
    const double d = M[j][j];
    ok = ok && (d > 0.0);
    const double inv = fast_rcp(d);
    id[j] = inv;
#pragma unroll
    for (int i = j + 1; i < N; i++) {
      const double u = M[j][i] * inv;                      // U[j][i]; row j still holds D_j U[j][l] for l >= i
#pragma unroll
      for (int l = i; l < N; l++) M[i][l] = fma(-u, M[j][l], M[i][l]);
      M[j][i] = u;
    }
  }
  return ok;
  } else {
  bool ok = true;                                          // M = U'U in place, id = 1/diag(U)
#pragma unroll
  for (int j = 0; j < N; j++) {
    double d = M[j][j];
#pragma unroll
    for (int k = 0; k < j; k++) d -= M[k][j] * M[k][j];
    ok = ok && (d > 0.0);
    double inv = fast_rsqrt(d);
    id[j] = inv;
    M[j][j] = d * inv;
#pragma unroll
    for (int i = j + 1; i < N; i++) {
      double a = M[j][i];
#pragma unroll
      for (int k = 0; k < j; k++) a -= M[k][j] * M[k][i];
      M[j][i] = a * inv;
    }
  }
  return ok;
  }
}
template <int N>
__device__ __forceinline__ void chol_solve(const double (&U)[N][N], const double (&id)[N], const double (&b)[N], double (&x)[N]) {
  double y[N];
  if constexpr (DCOL_LDL && N <= DCOL_LDL_MAXN) {
#pragma unroll
  for (int i = 0; i < N; i++) { double a = b[i];
#pragma unroll
    for (int k = 0; k < i; k++) a = fma(-U[k][i], y[k], a);
    y[i] = a; }
#pragma unroll
  for (int i = N - 1; i >= 0; i--) { double a = y[i] * id[i];
#pragma unroll
    for (int k = i + 1; k < N; k++) a = fma(-U[i][k], x[k], a);
    x[i] = a; }
  } else {
#pragma unroll
  for (int i = 0; i < N; i++) { double a = b[i];
#pragma unroll
    for (int k = 0; k < i; k++) a -= U[k][i] * y[k];
    y[i] = a * id[i]; }
#pragma unroll
  for (int i = N - 1; i >= 0; i--) { double a = y[i];
#pragma unroll
    for (int k = i + 1; k < N; k++) a -= U[i][k] * x[k];
    x[i] = a * id[i]; }
  }
}

// ---------------------------------------------------------------------------
// class-specific SOC block structure  u_s = Cs * yh   (yh = [x~; alpha; e])
//   BALL<NE>: [R*alpha ; x~ - B~ e]            (capsule :30-39, cylinder :146-160, sphere :99-113, polygon :124-134)
//   CONE    : [tb*x~1 + k*alpha ; x~2 ; x~3]   (cone :56-66, k = tb*3H/4)
//   ELL     : [alpha ; U x~]                   (ellipsoid :172-182)
// ---------------------------------------------------------------------------
template <int CLS, int NV, int D>
__device__ __forceinline__ void soc_apply(const double* sc, const double (&yh)[NV], double (&u)[D]) {
  constexpr int SOC = Cls<CLS>::SOC, NE = Cls<CLS>::NE;
  if constexpr (SOC == SOC_BALL) {
    u[0] = sc[0] * yh[3];
    u[1] = yh[0]; u[2] = yh[1]; u[3] = yh[2];
    if constexpr (NE >= 1) u[1] -= yh[4];
    if constexpr (NE == 2) u[2] -= yh[5];
  } else if constexpr (SOC == SOC_CONE) {
    u[0] = sc[0] * yh[0] + sc[1] * yh[3]; u[1] = yh[1]; u[2] = yh[2];
  } else if constexpr (SOC == SOC_ELL) {
    u[0] = yh[3];
#pragma unroll
    for (int i = 0; i < 3; i++) u[1 + i] = sc[3 * i] * yh[0] + sc[3 * i + 1] * yh[1] + sc[3 * i + 2] * yh[2];
  }
}
// g += Cs' v
template <int CLS, int NV, int D>
__device__ __forceinline__ void soc_applyT(const double* sc, const double (&v)[D], double (&g)[NV]) {
  constexpr int SOC = Cls<CLS>::SOC, NE = Cls<CLS>::NE;
  if constexpr (SOC == SOC_BALL) {
    g[0] += v[1]; g[1] += v[2]; g[2] += v[3]; g[3] += sc[0] * v[0];
    if constexpr (NE >= 1) g[4] -= v[1];
    if constexpr (NE == 2) g[5] -= v[2];
  } else if constexpr (SOC == SOC_CONE) {
    g[0] += sc[0] * v[0]; g[1] += v[1]; g[2] += v[2]; g[3] += sc[1] * v[0];
  } else if constexpr (SOC == SOC_ELL) {
#pragma unroll
    for (int a = 0; a < 3; a++) g[a] += sc[a] * v[1] + sc[3 + a] * v[2] + sc[6 + a] * v[3];
    g[3] += v[0];
  }
}
// N(upper) += scale * Cs' * B * Cs for a symmetric DxD matrix B (full storage)
template <int CLS, int NV, int D>
__device__ __forceinline__ void soc_gram(const double* sc, const double (&B)[D][D], double scale, double (&N)[NV][NV]) {
  constexpr int SOC = Cls<CLS>::SOC, NE = Cls<CLS>::NE;
  if constexpr (SOC == SOC_BALL) {
    const double R = sc[0], sR = scale * R;
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
      for (int j = i; j < 3; j++) N[i][j] += scale * B[i + 1][j + 1];
      N[i][3] += sR * B[i + 1][0];
      if constexpr (NE >= 1) N[i][4] -= scale * B[i + 1][1];
      if constexpr (NE == 2) N[i][5] -= scale * B[i + 1][2];
    }
    N[3][3] += sR * R * B[0][0];
    if constexpr (NE >= 1) { N[3][4] -= sR * B[0][1]; N[4][4] += scale * B[1][1]; }
    if constexpr (NE == 2) { N[3][5] -= sR * B[0][2]; N[4][5] += scale * B[1][2]; N[5][5] += scale * B[2][2]; }
  } else if constexpr (SOC == SOC_CONE) {
    const double E[3] = {sc[0], 1.0, 1.0};
    const double k = sc[1];
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
      for (int j = i; j < 3; j++) N[i][j] += scale * E[i] * E[j] * B[i][j];
      N[i][3] += scale * k * E[i] * B[i][0];
    }
    N[3][3] += scale * k * k * B[0][0];
  } else if constexpr (SOC == SOC_ELL) {
    double tmp[3][3];   // Bvv * U
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int b = 0; b < 3; b++) tmp[i][b] = B[1 + i][1] * sc[b] + B[1 + i][2] * sc[3 + b] + B[1 + i][3] * sc[6 + b];
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
      for (int b = a; b < 3; b++) N[a][b] += scale * (sc[a] * tmp[0][b] + sc[3 + a] * tmp[1][b] + sc[6 + a] * tmp[2][b]);
      N[a][3] += scale * (sc[a] * B[1][0] + sc[3 + a] * B[2][0] + sc[6 + a] * B[3][0]);
    }
    N[3][3] += scale * B[0][0];
  }
}

// reference SOC line search (static_solver2.jl:12-26) in reciprocal form: returns m such that the
// step to the cone boundary is 1/m (m <= 0: no bound).  i = 1/sqrt(J(y)), den = 1/(y0*i + 1).
// DCOL_LS_QUAD (default): the same boundary step from the quadratic J(y + a dl) = 0 directly.  With nu = J(y) > 0,
// zeta = y'J dl and a = J(dl):  rho0 = zeta/nu and |rho_v|^2 = rho0^2 - a/nu (the scaled point has J(rho) = J(dl)/nu),
// so m = |rho_v| - rho0 = (sqrt(zeta^2 - a nu) - zeta)/nu: the identical number with the identical cancellation
// structure, in 20 instructions instead of 26 and without the two reciprocals 1/(y0/sqrt(nu) + 1) per cone and iteration.
// zeta^2 >= a nu for every y inside the cone (reversed Cauchy-Schwarz); |.| keeps a round-off negative out of the root.
#ifndef DCOL_LS_QUAD
#define DCOL_LS_QUAD 1
#endif
// DCOL_LS_SCALED (default): both line searches of a cone block on the NT-scaled point.  W maps the cone onto itself
// and lambda = W z = W^-1 s, so  s + a ds in K  <=>  lambda + a W^-1 ds in K  and  z + a dz in K  <=>  lambda + a W dz in K:
// the same two boundary steps from the scaled directions the passes hold anyway, with one base point whose
// J(lambda) = sqrt(J(s) J(z)) is known from the scaling.  For the affine direction W^-1 ds = -v and W dz = v - lambda
// share their discriminant (zeta_v^2 - J(v) J(lambda)), so the two searches cost one square root and, like the
// orthant rows, one maximum:  max(m_s, m_z) = sq/nu + 1/2 + |zeta_v/nu - 1/2|; the unscaled affine dz is never formed.
// The first-order term of rho needs no sum either: s'dz_a + z'ds_a = -s'z for every block by construction of dz_a.
#ifndef DCOL_LS_SCALED
#define DCOL_LS_SCALED 1
#endif
// DCOL_ETA_FREE (default, needs DCOL_LS_SCALED): the scaled quantities are kept multiplied by eta = (J(s)/J(z))^(1/4), so
// that eta itself (a reciprocal square root behind two others in pass A's dependency chain) is never formed:
//   lam~ = eta lambda = Wbar^-1 s,  v~ = eta W^-1 t = Wbar^-1 t,  J(lam~) = J(s),  kappa~ = lam~ \ (v~ o w~ + eta^2 sigma mu e),
//   W^-1 kappa = eta^-2 Wbar^-1 kappa~,  dz = eta^-2 Wbar^-1 (p~ - lam~ + kappa~);  the line searches are ratios, unchanged.
// Only eta^2 = sqrt(J(s)/J(z)) and eta^-2 enter, both products of numbers pass A has anyway.
#ifndef DCOL_ETA_FREE
#define DCOL_ETA_FREE DCOL_LS_SCALED
#endif
template <int D>
__device__ __forceinline__ double soc_linesearch_m(const double (&y)[D], const double (&dl)[D], double i, double den) {
  double zeta = y[0] * dl[0];
#pragma unroll
  for (int k = 1; k < D; k++) zeta -= y[k] * dl[k];
  const double i2 = i * i;
#if DCOL_LS_QUAD
  const double nu = den;                       // J(y), kept from pass A in place of 1/(y0 i + 1)
  double a = dl[0] * dl[0];
#pragma unroll
  for (int k = 1; k < D; k++) a -= dl[k] * dl[k];
  const double disc = fma(zeta, zeta, -a * nu);
  return (fast_sqrt(fabs(disc)) - zeta) * i2;
#else
  const double rho0 = zeta * i2;
  const double f = (zeta * i + dl[0]) * den * i2;
  double nr = 0.0;
#pragma unroll
  for (int k = 1; k < D; k++) { double r = dl[k] * i - f * y[k]; nr += r * r; }
  return fast_sqrt(nr) - rho0;
#endif
}

// zeta = y' J d and a = J(d) = d0^2 - |d_v|^2 of a direction d at the point y
template <int D>
__device__ __forceinline__ void soc_zeta_a(const double (&y)[D], const double (&d)[D], double& zeta, double& a) {
  zeta = y[0] * d[0]; a = d[0] * d[0];
#pragma unroll
  for (int k = 1; k < D; k++) { zeta -= y[k] * d[k]; a -= d[k] * d[k]; }
}

// ---------------------------------------------------------------------------
// per-primitive solver state
// ---------------------------------------------------------------------------
template <int CLS>
struct Prim {
  static constexpr int NE = Cls<CLS>::NE, NV = 4 + Cls<CLS>::NE, D = Cls<CLS>::D, DD = (Cls<CLS>::D > 0 ? Cls<CLS>::D : 1);
  // orthant-row loops: unrolled twice only where rows are the bulk of the work (polytopes, polygons: -5 % on
  // polygon pairs, neutral on mixed batches); elsewhere code size wins.
  // The once-per-pair row loops (initial point, hand-over records, derivative epilogue) are never unrolled: the
  // compiler's default (x4) made the code ahead of the main loop 40 KB, which the other CTA of the SM pays for in
  // instruction-cache misses on its loop body.
#ifndef DCOL_ROW_UNROLL_SMALL
#define DCOL_ROW_UNROLL_SMALL 1
#endif
  static constexpr int ROW_UNROLL = (CLS == CLS_POLYTOPE || CLS == CLS_POLYGON) ? 2 : DCOL_ROW_UNROLL_SMALL;
  __device__ static constexpr bool nz(int a) { return Cls<CLS>::nz(a); }   // structurally non-zero row coefficient
  // constants of row i (zeros where the class has none).  Every shape's rows start on a 16-byte boundary
  // (dcol_set_shapes), so polytope rows (4 doubles) and the slack part of polygon rows (6 doubles) are read as double2.
  __device__ __forceinline__ void load_row(int i, double (&c)[NV]) const {
    if constexpr (CLS == CLS_POLYTOPE) {
      const double2* q = reinterpret_cast<const double2*>(rows + i * 4);
      const double2 u = __ldg(q), v = __ldg(q + 1);
      c[0] = u.x; c[1] = u.y; c[2] = v.x; c[3] = v.y;
    } else if constexpr (CLS == CLS_POLYGON) {
      c[0] = 0.0; c[1] = 0.0; c[2] = 0.0;
      c[3] = __ldg(rows + i * 6 + 3);
      const double2 v = __ldg(reinterpret_cast<const double2*>(rows + i * 6 + 4));
      c[4] = v.x; c[5] = v.y;
    } else if constexpr (CLS == CLS_CAPCYL) {
      // rows are [0,0,0,L/2,-1], [0,0,0,L/2,+1] (capsule, cylinder) and [+1,0,0,L/2,0], [-1,0,0,L/2,0] (cylinder only):
      // synthesised from L/2 (kept in a register) and the row index, no loads in the solver loop
      c[1] = 0.0; c[2] = 0.0; c[3] = rc0;
      c[4] = (i == 0) ? -1.0 : ((i == 1) ? 1.0 : 0.0);
      c[0] = (i == 2) ? 1.0 : ((i == 3) ? -1.0 : 0.0);
    } else if constexpr (CLS == CLS_CONE) {
      c[0] = -1.0; c[1] = 0.0; c[2] = 0.0; c[3] = rc0;      // the base plane: alpha*H/4 - x~_1
    } else {
#pragma unroll
      for (int a = 0; a < NV; a++) c[a] = nz(a) ? __ldg(rows + i * NV + a) : 0.0;
    }
  }
  // Row iteration of the per-iteration passes.  Capsules and cylinders have 2 or 4 rows of fixed structure
  // ([0,0,0,L/2,-1], [0,0,0,L/2,+1], and for cylinders [+1,0,0,L/2,0], [-1,0,0,L/2,0]): their passes run straight-line, one
  // compile-time row index each, so the row constants fold into the arithmetic (no row synthesis, no loop control, shared-
  // memory offsets as immediates; rows took a third of the instructions of a capsule pair's iteration before).  All
  // other classes run the run-time loop.  `f` receives the row index as std::integral_constant or as int.
  template <int I> using RowC = std::integral_constant<int, I>;
  template <class F>
  __device__ __forceinline__ void for_rows(F&& f) const {
    if constexpr (CLS == CLS_CAPCYL) {
      f(RowC<0>{}); f(RowC<1>{});
      if (n == 4) { f(RowC<2>{}); f(RowC<3>{}); }
    } else {
#pragma unroll ROW_UNROLL
      for (int i = 0; i < n; i++) f(i);
    }
  }
  template <class IT>
  __device__ __forceinline__ void row_coef(IT I, double (&c)[NV]) const {
    if constexpr (std::is_same<IT, int>::value) load_row(I, c);
    else {
      constexpr int i = IT::value;
#pragma unroll
      for (int a = 0; a < NV; a++) c[a] = 0.0;
      c[3] = rc0;
      if constexpr (i == 0) c[4] = -1.0;
      if constexpr (i == 1) c[4] = 1.0;
      if constexpr (i == 2) c[0] = 1.0;
      if constexpr (i == 3) c[0] = -1.0;
    }
  }
  template <class IT>
  __device__ static constexpr bool row_nz(IT, int a) {          // structurally non-zero coefficient a of this row
    if constexpr (std::is_same<IT, int>::value) return Cls<CLS>::nz(a);
    else return a == 3 || (IT::value < 2 ? a == 4 : a == 0);
  }
  const double* rows; int n; int row0;
  double rc0;                          // capsule / cylinder: L/2, cone: H/4 (the only non-trivial row constant)
  double sc[9];
  double s[DD], z[DD];                 // SOC block iterates
  double wb[DD], bb;                   // NT scaling point wbar and 1/(1+wbar0):  Wbar^-1 = J Wbar J (rank-one form)
  double ie;                           // 1/eta  (DCOL_ETA_FREE: 1/eta^2)
  double eta2;                         // DCOL_ETA_FREE: eta^2
  double irs, irz, dens, denz;         // 1/sqrt(J(s)), 1/sqrt(J(z)); line searches: dens = J(lambda), denz = 1/J(lambda)
                                       // [DCOL_LS_SCALED=0: J(s), J(z);  DCOL_LS_QUAD=0 as well: 1/(s0*irs+1), 1/(z0*irz+1)]
  double lam[DD], rz[DD], v[DD], wc[DD];

  // out = Wbar^-1 t  with Wbar^-1 = [wb0, -wbv'; -wbv, I + bb wbv wbv']
  __device__ __forceinline__ void applyV(const double (&t)[DD], double (&out)[DD]) const {
    double c = 0.0;
#pragma unroll
    for (int i = 1; i < D; i++) c += wb[i] * t[i];
    out[0] = wb[0] * t[0] - c;
    const double f = bb * c - t[0];
#pragma unroll
    for (int i = 1; i < D; i++) out[i] = t[i] + wb[i] * f;
  }

  __device__ __forceinline__ void load(const DevShape* S, const double* rows_base, int row0_) {
    n = (CLS == CLS_CONE) ? 1 : S->n_ort;      // a cone has exactly one orthant row (its base plane)
    rows = rows_base + S->rows_off; row0 = row0_;
    if constexpr (CLS == CLS_CAPCYL || CLS == CLS_CONE) rc0 = __ldg(rows + 3);
    constexpr int SOC = Cls<CLS>::SOC;
    if constexpr (SOC == SOC_BALL) sc[0] = S->soc[0];
    if constexpr (SOC == SOC_CONE) { sc[0] = S->soc[0]; sc[1] = S->soc[1]; }
    if constexpr (SOC == SOC_ELL) {
#pragma unroll
      for (int i = 0; i < 9; i++) sc[i] = S->soc[i]; }
  }

  // ---- initial point (static_solver2.jl:86-103): unweighted Gram N = C'C and g = C' u(yh0)
  __device__ __forceinline__ void init_gram(const double (&yh0)[NV], bool has_off, double (&N)[NV][NV], double (&g)[NV]) const {
#pragma unroll
    for (int a = 0; a < NV; a++) { g[a] = 0.0;
#pragma unroll
      for (int b = a; b < NV; b++) N[a][b] = 0.0; }
#pragma unroll 1
    for (int i = 0; i < n; i++) {
      double c[NV];
      load_row(i, c);
      double u = 0.0;
      if (has_off) {
#pragma unroll
        for (int a = 0; a < NV; a++) if (nz(a)) u += c[a] * yh0[a]; }
#pragma unroll
      for (int a = 0; a < NV; a++) if (nz(a)) { g[a] += u * c[a];
#pragma unroll
        for (int b = a; b < NV; b++) if (nz(b)) N[a][b] += c[a] * c[b]; }
    }
    if constexpr (D > 0) {
      double I[D][D];
#pragma unroll
      for (int i = 0; i < D; i++)
#pragma unroll
        for (int j = 0; j < D; j++) I[i][j] = (i == j) ? 1.0 : 0.0;
      soc_gram<CLS, NV, D>(sc, I, 1.0, N);
      if (has_off) { double u[D]; soc_apply<CLS, NV, D>(sc, yh0, u); soc_applyT<CLS, NV, D>(sc, u, g); }
    }
  }
  // s~ = -u(yh), z~ = C*dh ; returns the bring2cone violations (static_solver2.jl:52-85)
  __device__ __forceinline__ void init_rows(const RowStore& rs, const double (&yh)[NV], const double (&dh)[NV],
                                            double& min_s, double& min_z, double& a_s, double& a_z) {
#pragma unroll 1
    for (int i = 0; i < n; i++) {
      double us = 0.0, uz = 0.0;
      double c[NV];
      load_row(i, c);
#pragma unroll
      for (int a = 0; a < NV; a++) if (nz(a)) { us += c[a] * yh[a]; uz += c[a] * dh[a]; }
      rs(F_S, row0 + i) = -us; rs(F_Z, row0 + i) = uz;
      min_s = dmin(min_s, -us); min_z = dmin(min_z, uz);
    }
    if constexpr (D > 0) {
      double u[D];
      soc_apply<CLS, NV, D>(sc, yh, u);
      double ns = 0.0, nz = 0.0;
#pragma unroll
      for (int i = 0; i < D; i++) s[i] = -u[i];
      soc_apply<CLS, NV, D>(sc, dh, z);
#pragma unroll
      for (int i = 1; i < D; i++) { ns += s[i] * s[i]; nz += z[i] * z[i]; }
      double res_s = s[0] - fast_sqrt(ns), res_z = z[0] - fast_sqrt(nz);
      if (res_s <= 0.0) a_s = dmax(a_s, -res_s);
      if (res_z <= 0.0) a_z = dmax(a_z, -res_z);
    }
  }
  __device__ __forceinline__ double init_shift(const RowStore& rs, double sh_s, double sh_z) {   // + (1+a)*e ; returns s.z
    double sz = 0.0;
#pragma unroll 1
    for (int i = 0; i < n; i++) {
      double a = rs(F_S, row0 + i) + sh_s, b = rs(F_Z, row0 + i) + sh_z;
      rs(F_S, row0 + i) = a; rs(F_Z, row0 + i) = b; sz += a * b;
    }
    if constexpr (D > 0) { s[0] += sh_s; z[0] += sh_z;
#pragma unroll
      for (int i = 0; i < D; i++) sz += s[i] * z[i]; }
    return sz;
  }

  // Every pass is split into its orthant-row loop and its SOC part so that solve_pair can run the row
  // loops of both primitives first and then the two (independent) straight-line SOC parts back to back,
  // which lets the scheduler interleave them.

  // ---- pass A: residual rz = s - u(yh), scalings, N = C' W^-2 C, g = C' W^-2 rz
  __device__ __forceinline__ void passA_rows(const RowStore& rs, const double (&yh)[NV], double (&N)[NV][NV], double (&g)[NV]) {
#pragma unroll
    for (int a = 0; a < NV; a++) { g[a] = 0.0;
#pragma unroll
      for (int b = a; b < NV; b++) N[a][b] = 0.0; }
    for_rows([&](auto I) {
      const int i = I;
      double c[NV];
      row_coef(I, c);
      const double si = rs(F_S, row0 + i), zi = rs(F_Z, row0 + i);
      double u = 0.0;
#pragma unroll
      for (int a = 0; a < NV; a++) if (row_nz(I, a)) u += c[a] * yh[a];
      const double r = si - u;
      const double rr = fast_rcp(si * zi);     // 1/(s z): 1/s = z*rr, 1/z = s*rr
      const double d = zi * zi * rr;           // z/s
      const double dr = d * r;
#pragma unroll
      for (int a = 0; a < NV; a++) if (row_nz(I, a)) {
        const double dc = d * c[a];
        g[a] += dr * c[a];
#pragma unroll
        for (int b = a; b < NV; b++) if (row_nz(I, b)) N[a][b] += dc * c[b];
      }
      rs(F_R, row0 + i) = rr; rs(F_RZ, row0 + i) = r;
    });
  }
  __device__ __forceinline__ bool passA_soc(const double (&yh)[NV], double (&N)[NV][NV], double (&g)[NV]) {
    if constexpr (D > 0) {
      double u[D];
      soc_apply<CLS, NV, D>(sc, yh, u);
      double Js = s[0] * s[0], Jz = z[0] * z[0];
      double dot = s[0] * z[0];
#pragma unroll
      for (int i = 0; i < D; i++) rz[i] = s[i] - u[i];
#pragma unroll
      for (int i = 1; i < D; i++) { Js -= s[i] * s[i]; Jz -= z[i] * z[i]; dot += s[i] * z[i]; }
      const bool ok = (Js > 0.0) && (Jz > 0.0);
      // Nesterov-Todd scaling (NT_scaling_chol_2.jl:95-112)
      irs = fast_rsqrt(Js); irz = fast_rsqrt(Jz);
#if DCOL_ETA_FREE
      // (set with the scaling below)
#elif DCOL_LS_SCALED
      dens = (Js * irs) * (Jz * irz);          // J(lambda) = sqrt(J(s) J(z))
      denz = irs * irz;                        // 1 / J(lambda)
#elif DCOL_LS_QUAD
      dens = Js; denz = Jz;
#else
      dens = fast_rcp(s[0] * irs + 1.0); denz = fast_rcp(z[0] * irz + 1.0);
#endif
      const double kk = 0.5 * fast_rsqrt(0.5 * (1.0 + dot * irs * irz));     // 1/(2 gamma)
      wb[0] = kk * (s[0] * irs + z[0] * irz);
#pragma unroll
      for (int i = 1; i < D; i++) wb[i] = kk * (s[i] * irs - z[i] * irz);
      bb = fast_rcp(wb[0] + 1.0);
      const double x = irs * (Jz * irz);          // sqrt(J(z)/J(s)) = 1/eta^2
      const double ie2 = x;
#if DCOL_ETA_FREE
      ie = x;                                     // eta^-2
      eta2 = (Js * irs) * irz;                    // eta^2
      dens = Js; denz = irs * irs;                // J(lam~) and its reciprocal
      applyV(s, lam);                             // lam~ = eta lambda = Wbar^-1 s
#else
      const double eta = fast_rsqrt(x);           // (J(s)/J(z))^(1/4)
      ie = x * eta;
      // lambda = W z = eta * Wbar z
      double wz = 0.0;
#pragma unroll
      for (int i = 1; i < D; i++) wz += wb[i] * z[i];
      lam[0] = eta * (wb[0] * z[0] + wz);
      const double fz = z[0] + bb * wz;
#pragma unroll
      for (int i = 1; i < D; i++) lam[i] = eta * (z[i] + wb[i] * fz);
#endif
      // Vsq = V'V with V = Wbar^-1, expanded entrywise (a Gram product: PSD to round-off, no use of
      // the identity wb0^2 - |wbv|^2 = 1, which loses accuracy as the iterates approach the boundary)
      double nw = 0.0;
#pragma unroll
      for (int i = 1; i < D; i++) nw += wb[i] * wb[i];
      const double c0 = wb[0] + 1.0 + bb * nw;
      const double c1 = 1.0 + bb * (2.0 + bb * nw);
      double Vsq[D][D];
      Vsq[0][0] = wb[0] * wb[0] + nw;
#pragma unroll
      for (int i = 1; i < D; i++) { Vsq[0][i] = -wb[i] * c0; Vsq[i][0] = Vsq[0][i];
#pragma unroll
        for (int j = i; j < D; j++) { double e = c1 * wb[i] * wb[j] + ((i == j) ? 1.0 : 0.0); Vsq[i][j] = e; Vsq[j][i] = e; } }
      // t2 = W^-2 rz as two applications of W^-1 (scaled-space accuracy)
      double t1[D], t2[D];
      applyV(rz, t1); applyV(t1, t2);
#pragma unroll
      for (int i = 0; i < D; i++) t2[i] *= ie2;
      soc_applyT<CLS, NV, D>(sc, t2, g);
      soc_gram<CLS, NV, D>(sc, Vsq, ie2, N);
      return ok;
    } else {
      return true;
    }
  }

  // ---- pass B: affine direction; step bounds (reciprocal form: m = max -d/y); sums for rho; Mehrotra correction
  // (orthant rows: the two ratio tests max q and max (1 - q) are one, 1/2 + max |q - 1/2|; mh carries max |q - 1/2|)
  __device__ __forceinline__ void passB_rows(const RowStore& rs, const double (&dyh)[NV], double& mh, double& dot1, double& dot2) {
    for_rows([&](auto I) {
      const int i = I;
      double cd = 0.0;
      double c[NV];
      row_coef(I, c);
#pragma unroll
      for (int a = 0; a < NV; a++) if (row_nz(I, a)) cd += c[a] * dyh[a];
      const double si = rs(F_S, row0 + i), zi = rs(F_Z, row0 + i), rr = rs(F_R, row0 + i), r = rs(F_RZ, row0 + i);
      const double t = r - cd;                  // -ds
      const double is = zi * rr;
      const double q = is * t;                  // -ds/s  and  1 + dz/z
      const double dza = zi * q - zi;
      mh = dmax(mh, fabs(q - 0.5));
      const double pp = -t * dza;
      rs(F_P, row0 + i) = pp;
#if DCOL_LS_SCALED
      dot2 += pp;                               // (s'dza + z'dsa = -s'z exactly: no sum needed, see DCOL_LS_SCALED)
#else
      dot1 += si * dza - zi * t; dot2 += pp;
#endif
    });
  }
  __device__ __forceinline__ void passB_soc(const double (&dyh)[NV], double& m, double& dot1, double& dot2) {
    if constexpr (D > 0) {
      double cd[D], t[D], w[D];
      soc_apply<CLS, NV, D>(sc, dyh, cd);
#pragma unroll
      for (int i = 0; i < D; i++) t[i] = rz[i] - cd[i];
      // v = W^-1 t = -W^-1 dsa ; w = v - lambda = W dza   [dza = W^-2 t - z, subtraction done in the scaled space]
      applyV(t, v);
#if DCOL_ETA_FREE
#pragma unroll
      for (int i = 0; i < D; i++) w[i] = v[i] - lam[i];                    // v~, w~ (both carry a factor eta)
#else
#pragma unroll
      for (int i = 0; i < D; i++) { v[i] *= ie; w[i] = v[i] - lam[i]; }
#endif
#if DCOL_LS_SCALED
      {
        double zv, av;
        soc_zeta_a<D>(lam, v, zv, av);
        const double sq = fast_sqrt(fabs(fma(zv, zv, -av * dens)));
        m = dmax(m, fma(sq, denz, 0.5) + fabs(fma(zv, denz, -0.5)));      // max(m_s, m_z), see DCOL_LS_SCALED
        double vw = 0.0;                                                   // dsa'dza = -(W^-1 dsa)'(W dza) = -v'w
#pragma unroll
        for (int i = 0; i < D; i++) vw += v[i] * w[i];
#if DCOL_ETA_FREE
        dot2 -= ie * vw;
#else
        dot2 -= vw;
#endif
      }
#else
      double dsa[D], dza[D];
      applyV(w, dza);
#pragma unroll
      for (int i = 0; i < D; i++) { dsa[i] = -t[i]; dza[i] *= ie; }
      m = dmax(m, soc_linesearch_m<D>(s, dsa, irs, dens));
      m = dmax(m, soc_linesearch_m<D>(z, dza, irz, denz));
#pragma unroll
      for (int i = 0; i < D; i++) { dot1 += s[i] * dza[i] + z[i] * dsa[i]; dot2 += dsa[i] * dza[i]; }
#endif
      // -(W^-1 dsa) o (W dza) = v o (v - lambda)   (static_solver2.jl:164)
      double c0 = 0.0;
#pragma unroll
      for (int i = 0; i < D; i++) c0 += v[i] * w[i];
      wc[0] = c0;
#pragma unroll
      for (int i = 1; i < D; i++) wc[i] = v[0] * w[i] + w[0] * v[i];
    }
  }

  // ---- pass C: corrector right-hand side  g2 = C' [ (sigma*mu - p)/s | W^-1 (lambda \ corr) ]
  __device__ __forceinline__ void passC_rows(const RowStore& rs, double sigmu, double (&g2)[NV]) {
#pragma unroll
    for (int a = 0; a < NV; a++) g2[a] = 0.0;
    for_rows([&](auto I) {
      const int i = I;
      const double q = (sigmu - rs(F_P, row0 + i)) * rs(F_Z, row0 + i) * rs(F_R, row0 + i);
      double c[NV];
      row_coef(I, c);
#pragma unroll
      for (int a = 0; a < NV; a++) if (row_nz(I, a)) g2[a] += q * c[a];
    });
  }
  __device__ __forceinline__ void passC_soc(double sigmu, double (&g2)[NV]) {
    if constexpr (D > 0) {
#if DCOL_ETA_FREE
      wc[0] += eta2 * sigmu;
#else
      wc[0] += sigmu;
#endif
      double kap[D], wk[D];
      soc_jdiv_fast(lam, wc, kap);              // kappa = lambda \ corr  (soc_utils.jl:52-67)
      applyV(kap, wk);
#pragma unroll
      for (int i = 0; i < D; i++) { wk[i] *= ie; wc[i] = kap[i]; }
      soc_applyT<CLS, NV, D>(sc, wk, g2);
    }
  }
  __device__ __forceinline__ void soc_jdiv_fast(const double (&l)[DD], const double (&w)[DD], double (&out)[DD]) const {
#if DCOL_LS_SCALED
    // J(lambda) = sqrt(J(s) J(z)) and its reciprocal are known from the scaling (denz), no sum and no reciprocal here
    double nu = 0.0;
#pragma unroll
    for (int i = 1; i < D; i++) nu += l[i] * w[i];
    const double il0 = fast_rcp(l[0]);
    const double irho = denz;
#else
    double rho = l[0] * l[0], nu = 0.0;
#pragma unroll
    for (int i = 1; i < D; i++) { rho -= l[i] * l[i]; nu += l[i] * w[i]; }
    const double il0 = fast_rcp(l[0]);
    const double irho = fast_rcp(rho);
#endif
    out[0] = irho * (l[0] * w[0] - nu);
    const double f1 = irho * (nu * il0 - w[0]);
#pragma unroll
    for (int i = 1; i < D; i++) out[i] = f1 * l[i] + il0 * w[i];
  }

  // ---- pass D: final direction and step bounds; directions parked in F_RZ/F_P (rows) and rz/v (SOC)
  __device__ __forceinline__ void passD_rows(const RowStore& rs, const double (&dyh)[NV], double sigmu, double& m) {
    for_rows([&](auto I) {
      const int i = I;
      double cd = 0.0;
      double c[NV];
      row_coef(I, c);
#pragma unroll
      for (int a = 0; a < NV; a++) if (row_nz(I, a)) cd += c[a] * dyh[a];
      const double si = rs(F_S, row0 + i), zi = rs(F_Z, row0 + i), rr = rs(F_R, row0 + i), r = rs(F_RZ, row0 + i), pp = rs(F_P, row0 + i);
      const double t = r - cd;
      const double is = zi * rr, iz = si * rr;
      const double dz = (zi * t + sigmu - pp) * is - zi;
      m = dmax(m, t * is); m = dmax(m, -dz * iz);
      rs(F_RZ, row0 + i) = -t; rs(F_P, row0 + i) = dz;
    });
  }
  __device__ __forceinline__ void passD_soc(const double (&dyh)[NV], double& m) {
    if constexpr (D > 0) {
      double cd[D], t[D], ds[D], dz[D], v2[D];
      soc_apply<CLS, NV, D>(sc, dyh, cd);
#pragma unroll
      for (int i = 0; i < D; i++) { t[i] = rz[i] - cd[i]; ds[i] = -t[i]; }
      // dz = W^-1 (W^-1 t - lambda + kappa)   (scaled-space accumulation, static_solver2.jl:165-168)
      applyV(t, v2);
#if DCOL_LS_SCALED
      {
        double p[D], zp, ap, z2, a2;
#pragma unroll
#if DCOL_ETA_FREE
        for (int i = 0; i < D; i++) { p[i] = v2[i]; v2[i] = p[i] - lam[i] + wc[i]; }          // p~ = -eta W^-1 ds, v2~ = eta W dz
#else
        for (int i = 0; i < D; i++) { p[i] = ie * v2[i]; v2[i] = p[i] - lam[i] + wc[i]; }     // p = -W^-1 ds, v2 = W dz
#endif
        soc_zeta_a<D>(lam, p, zp, ap);
        soc_zeta_a<D>(lam, v2, z2, a2);
        m = dmax(m, (fast_sqrt(fabs(fma(zp, zp, -ap * dens))) + zp) * denz);                 // direction -p
        m = dmax(m, (fast_sqrt(fabs(fma(z2, z2, -a2 * dens))) - z2) * denz);
      }
      applyV(v2, dz);
#pragma unroll
      for (int i = 0; i < D; i++) dz[i] *= ie;
#else
#pragma unroll
      for (int i = 0; i < D; i++) v2[i] = ie * v2[i] - lam[i] + wc[i];
      applyV(v2, dz);
#pragma unroll
      for (int i = 0; i < D; i++) dz[i] *= ie;
      m = dmax(m, soc_linesearch_m<D>(s, ds, irs, dens));
      m = dmax(m, soc_linesearch_m<D>(z, dz, irz, denz));
#endif
#pragma unroll
      for (int i = 0; i < D; i++) { rz[i] = ds[i]; v[i] = dz[i]; }
    }
  }

  // ---- pass E: take the step; returns the new s.z
  __device__ __forceinline__ double passE_rows(const RowStore& rs, double al) {
    double sz = 0.0;
    for_rows([&](auto I) {
      const int i = I;
      double a = rs(F_S, row0 + i) + al * rs(F_RZ, row0 + i), b = rs(F_Z, row0 + i) + al * rs(F_P, row0 + i);
      rs(F_S, row0 + i) = a; rs(F_Z, row0 + i) = b; sz += a * b;
    });
    return sz;
  }
  __device__ __forceinline__ double passE_soc(double al) {
    double sz = 0.0;
    if constexpr (D > 0) {
#pragma unroll
      for (int i = 0; i < D; i++) { s[i] += al * rz[i]; z[i] += al * v[i]; sz += s[i] * z[i]; }
    }
    return sz;
  }

  // ---- derivative epilogue: the vectors of dcol_deriv.cuh's SideVecs for this primitive
  //      wr = (C' z)[0:3] over the body-frame part (orthant rows, cone / ellipsoid SOC block),
  //      zb = vector part of z of a ball SOC block, Be = B~ e
  __device__ __forceinline__ void side_vecs(const RowStore& rs, const double (&yh)[NV], SideVecs& V) const {
    double w[NV];
#pragma unroll
    for (int a = 0; a < NV; a++) w[a] = 0.0;
#pragma unroll 1
    for (int i = 0; i < n; i++) {
      double zi = rs(F_Z, row0 + i);
      double c[NV];
      load_row(i, c);
#pragma unroll
      for (int a = 0; a < 3; a++) if (nz(a)) w[a] += zi * c[a];
    }
    constexpr int SOC = Cls<CLS>::SOC;
    if constexpr (D > 0 && SOC != SOC_BALL) soc_applyT<CLS, NV, D>(sc, z, w);
#pragma unroll
    for (int a = 0; a < 3; a++) { V.wr[a] = w[a]; V.xt[a] = yh[a]; V.zb[a] = 0.0; V.Be[a] = 0.0; }
    if constexpr (SOC == SOC_BALL) {
#pragma unroll
      for (int a = 0; a < 3; a++) V.zb[a] = z[1 + a];
      if constexpr (NE >= 1) V.Be[0] = yh[4];
      if constexpr (NE == 2) V.Be[1] = yh[5];
    }
  }
  // ---- derivative epilogue: Nf = C' (S^-1 Z) C (full, slightly non-symmetric on SOC blocks), with the
  // solver's slack s (src/proximity.jl:98-100: Z = arrow(z), S = arrow(s))
  // Nbx = first three columns of the ball SOC block's share of Nf (zero for other classes)
  __device__ __forceinline__ void ift_gram(const RowStore& rs, double (&Nf)[NV][NV], double (&Nbx)[NV][3]) const {
#pragma unroll
    for (int a = 0; a < NV; a++)
#pragma unroll
      for (int b = 0; b < NV; b++) Nf[a][b] = 0.0;
#pragma unroll 1
    for (int i = 0; i < n; i++) {
      double c[NV];
      load_row(i, c);
      double d = rs(F_Z, row0 + i) * fast_rcp(rs(F_S, row0 + i));
#pragma unroll
      for (int a = 0; a < NV; a++) if (nz(a)) { double dc = d * c[a];
#pragma unroll
        for (int b = a; b < NV; b++) if (nz(b)) Nf[a][b] += dc * c[b]; }
    }
#pragma unroll
    for (int a = 0; a < NV; a++) {
#pragma unroll
      for (int b = 0; b < a; b++) Nf[a][b] = Nf[b][a];
#pragma unroll
      for (int c = 0; c < 3; c++) Nbx[a][c] = 0.0;
    }
    if constexpr (D > 0) {
      // Bm = arrow(s)^-1 arrow(z), column by column through the closed-form Jordan division
      double Bm[D][D];
#pragma unroll
      for (int c = 0; c < D; c++) {
        double col[D], out[D];
#pragma unroll
        for (int i = 0; i < D; i++) col[i] = (c == 0) ? z[i] : ((i == 0) ? z[c] : ((i == c) ? z[0] : 0.0));
        soc_jdiv<D>(s, col, out);
#pragma unroll
        for (int i = 0; i < D; i++) Bm[i][c] = out[i];
      }
#pragma unroll
      for (int b = 0; b < NV; b++) {
        double e[NV], u[D], v[D], g[NV];
#pragma unroll
        for (int a = 0; a < NV; a++) { e[a] = (a == b) ? 1.0 : 0.0; g[a] = 0.0; }
        soc_apply<CLS, NV, D>(sc, e, u);
#pragma unroll
        for (int i = 0; i < D; i++) { double acc = 0.0;
#pragma unroll
          for (int k = 0; k < D; k++) acc += Bm[i][k] * u[k];
          v[i] = acc; }
        soc_applyT<CLS, NV, D>(sc, v, g);
#pragma unroll
        for (int a = 0; a < NV; a++) Nf[a][b] += g[a];
        if constexpr (Cls<CLS>::SOC == SOC_BALL) {
          if (b < 3) {
#pragma unroll
            for (int a = 0; a < NV; a++) Nbx[a][b] = g[a];
          }
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------
// pair-level assembly in primitive-1's body frame.  Pair variables
// y = [xi(3), alpha, e1(NE1), e2(NE2)]  (column order of combine_problem_matrices.jl:57-84).
// ---------------------------------------------------------------------------
// DCOL_SPHERE_FRAME (default): a sphere has no orientation the solve could see (||x - r~|| <= R alpha), so in a pair with
// a sphere the sphere's frame is taken equal to the OTHER primitive's for the duration of the interior-point loop: the
// relative rotation is the identity (tag IdentityR: no rotation of Gram blocks, right-hand sides or local vectors), and
// the iterates are a rotation of the ones in the sphere's own body frame (a rotation of a ball block's vector part is a
// cone automorphism: same mu, alpha, x, iteration counts).  The derivative epilogues, which look at the sphere's true
// frame (its offset r_offset turns with it), get the iterate converted back at the loop exit (solve_pair).
#ifndef DCOL_SPHERE_FRAME
#define DCOL_SPHERE_FRAME 1
#endif
struct IdentityR { __device__ IdentityR() {} template <class T> __device__ explicit IdentityR(const T&) {} };
template <class RM> struct is_identity_r : std::false_type {};
template <> struct is_identity_r<IdentityR> : std::true_type {};

template <int NE1, int NE2>
struct PairMap {
  static constexpr int NX = 4 + NE1 + NE2, NV1 = 4 + NE1, NV2 = 4 + NE2;
  __device__ static __forceinline__ constexpr int g2(int l) { return l < 4 ? l : 4 + NE1 + (l - 4); }   // prim-2 local -> pair index (l>=3)

  // local vectors from the pair vector; with_t adds the frame offset t to prim 2
  // R and t are anything indexable as R[a][b], t[a]: register arrays or the shared-memory view RelPoseR / RelPoseT
  template <class RM, class TV>
  __device__ static __forceinline__ void to_local(const double (&y)[NX], const RM& R, const TV& t, bool with_t,
                                                  double (&y1)[NV1], double (&y2)[NV2]) {
#pragma unroll
    for (int a = 0; a < NV1; a++) y1[a] = y[a];
    if constexpr (is_identity_r<RM>::value) {
#pragma unroll
      for (int a = 0; a < 3; a++) y2[a] = with_t ? y[a] + t[a] : y[a];
    } else {
#pragma unroll
      for (int a = 0; a < 3; a++) y2[a] = R[a][0] * y[0] + R[a][1] * y[1] + R[a][2] * y[2] + (with_t ? t[a] : 0.0);
    }
    y2[3] = y[3];
#pragma unroll
    for (int a = 4; a < NV2; a++) y2[a] = y[g2(a)];
  }
  // M(upper) = embed(N1) + T' N2 T ; b = embed(g1) + T' g2
  template <class RM>
  __device__ static __forceinline__ void assemble(const double (&N1)[NV1][NV1], const double (&N2)[NV2][NV2], const RM& R,
                                                  double (&M)[NX][NX]) {
#pragma unroll
    for (int a = 0; a < NX; a++)
#pragma unroll
      for (int b = a; b < NX; b++) M[a][b] = (a < NV1 && b < NV1) ? N1[a][b] : 0.0;
    if constexpr (is_identity_r<RM>::value) {
#pragma unroll
      for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int b = a; b < 3; b++) M[a][b] += N2[a][b];
#pragma unroll
        for (int l = 3; l < NV2; l++) M[a][g2(l)] += N2[a][l];
      }
#pragma unroll
      for (int l = 3; l < NV2; l++)
#pragma unroll
        for (int m = l; m < NV2; m++) M[g2(l)][g2(m)] += N2[l][m];
    } else {
    double tmp[3][3];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int b = 0; b < 3; b++) {
        double a0 = 0.0;
#pragma unroll
        for (int d = 0; d < 3; d++) a0 += ((c <= d) ? N2[c][d] : N2[d][c]) * R[d][b];
        tmp[c][b] = a0;
      }
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
      for (int b = a; b < 3; b++) M[a][b] += R[0][a] * tmp[0][b] + R[1][a] * tmp[1][b] + R[2][a] * tmp[2][b];
#pragma unroll
      for (int l = 3; l < NV2; l++) M[a][g2(l)] += R[0][a] * N2[0][l] + R[1][a] * N2[1][l] + R[2][a] * N2[2][l];
    }
#pragma unroll
    for (int l = 3; l < NV2; l++)
#pragma unroll
      for (int m = l; m < NV2; m++) M[g2(l)][g2(m)] += N2[l][m];
    }
  }
  template <class RM>
  __device__ static __forceinline__ void assemble_vec(const double (&g1)[NV1], const double (&g2v)[NV2], const RM& R, double (&b)[NX]) {
#pragma unroll
    for (int a = 0; a < NX; a++) b[a] = (a < NV1) ? g1[a] : 0.0;
    if constexpr (is_identity_r<RM>::value) {
#pragma unroll
      for (int a = 0; a < 3; a++) b[a] += g2v[a];
    } else {
#pragma unroll
      for (int a = 0; a < 3; a++) b[a] += R[0][a] * g2v[0] + R[1][a] * g2v[1] + R[2][a] * g2v[2];
    }
#pragma unroll
    for (int l = 3; l < NV2; l++) b[g2(l)] += g2v[l];
  }
};

// primitive 2's world frame from primitive 1's and the relative pose: R = Q2'Q1, t = Q2'(r~1 - r~2)  =>  Q2 = Q1 R', r~2 = r~1 - Q2 t
template <class RM, class TV>
__device__ __forceinline__ void frame2_from_relative(const Frame& F1, const RM& R, const TV& t, Frame& F2) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) F2.Q[i][j] = F1.Q[i][0] * R[j][0] + F1.Q[i][1] * R[j][1] + F1.Q[i][2] * R[j][2];
#pragma unroll
  for (int i = 0; i < 3; i++) F2.r[i] = F1.r[i] - (F2.Q[i][0] * t[0] + F2.Q[i][1] * t[1] + F2.Q[i][2] * t[2]);
}

// One pair: relative pose -> initial point -> PDIP loop -> outputs (+ derivative epilogue).
// VAR (= the query kind: Q_PROX, Q_GRAD, Q_JAC) selects the kernel variant: the proximity variant carries no epilogue,
// the gradient variant the small envelope-theorem epilogue, the Jacobian variant the implicit-function-theorem one
// (whose register pressure would otherwise slow the other two queries' solver loops).
template <int C1, int C2, int VAR>
__device__ __forceinline__ void solve_pair(const QueryParams& p, const TileCtx& tc, int idx) {
  using P1 = Prim<C1>; using P2 = Prim<C2>;
  using PM = PairMap<P1::NE, P2::NE>;
  constexpr int NX = PM::NX, NV1 = PM::NV1, NV2 = PM::NV2;
  // sphere frames (see IdentityR): SPH2 = primitive 2 is a sphere; SPH1 = primitive 1 is and primitive 2 is not
  constexpr bool SPH2 = DCOL_SPHERE_FRAME && (C2 == CLS_SPHERE);
  constexpr bool SPH1 = DCOL_SPHERE_FRAME && (C1 == CLS_SPHERE) && !SPH2;

  // first pass: idx is a position of the sorted, packed pair list; resume pass: idx is a slot of the bin's overflow
  // list, whose first field is the pair's sorted position.
  // Only class pairs (C1 <= C2) are instantiated: a pair whose classes arrive in the other order was packed with its two
  // primitives exchanged (alpha and the world point x do not depend on the order; the derivative blocks are written to
  // the columns of the primitive they belong to).
  const int pos = (tc.mode == 0) ? idx : (int)tc.in[idx];
  const PairRec* const rec = p.recs + pos;
  const int na = (p.att_kind == 0) ? 4 : 3;

  RowStore rs{dcol_dyn_smem + (size_t)threadIdx.x * p.row_pitch};
  // The whole record is fetched up front (eight independent 16-byte loads of one cache line in flight at once): all
  // warps of the CTA start a tile together, so a chain of dependent loads would leave the SM idle for several DRAM round
  // trips per tile.
  double pose[14];
  int s1i, s2i, pair, rflags;
  {
    const double2* q = reinterpret_cast<const double2*>(rec);
    double2 v[8];
#pragma unroll
    for (int a = 0; a < 8; a++) v[a] = __ldg(q + a);
#pragma unroll
    for (int a = 0; a < 7; a++) { pose[2 * a] = v[a].x; pose[2 * a + 1] = v[a].y; }
    s1i = __double2loint(v[7].x); s2i = __double2hiint(v[7].x);
    pair = __double2loint(v[7].y); rflags = __double2hiint(v[7].y);
  }
  const bool swp = (rflags & DCOL_REC_SWAPPED) != 0;
  const DevShape* S1 = p.shapes + s1i;
  const DevShape* S2 = p.shapes + s2i;
  const int64_t oi = p.direct ? (int64_t)pair : (int64_t)pos;      // index of the scalar outputs
  P1 A; P2 B;
  A.load(S1, p.rows, 0);
  B.load(S2, p.rows, A.n);

  // relative pose (computed in registers, parked in shared memory for the rest of the solve)
  {
    double R[3][3], t[3];
    Frame F1, F2;
    build_frame(S1, pose, pose + 3, p.att_kind, F1);
    build_frame(S2, pose + 7, pose + 10, p.att_kind, F2);
    double dr[3] = {F1.r[0] - F2.r[0], F1.r[1] - F2.r[1], F1.r[2] - F2.r[2]};
    // loop frames (DCOL_SPHERE_FRAME): sphere as primitive 2 -> its frame := primitive 1's (t in primitive 1's frame);
    // sphere as primitive 1 only -> the pair frame := primitive 2's (t as always, x = r~1 + Q2 xi)
    const double (&Qt)[3][3] = SPH2 ? F1.Q : F2.Q;          // frame the offset t is expressed in
    const double (&Qx)[3][3] = SPH1 ? F2.Q : F1.Q;          // frame of the pair variable xi
#pragma unroll
    for (int a = 0; a < 3; a++) {
      t[a] = Qt[0][a] * dr[0] + Qt[1][a] * dr[1] + Qt[2][a] * dr[2];
#pragma unroll
      for (int b = 0; b < 3; b++) R[a][b] = F2.Q[0][a] * F1.Q[0][b] + F2.Q[1][a] * F1.Q[1][b] + F2.Q[2][a] * F1.Q[2][b];
    }
#pragma unroll
    for (int a = 0; a < 3; a++) { rs.base[9 + a] = t[a]; rs.base[21 + a] = F1.r[a];
#pragma unroll
      for (int b = 0; b < 3; b++) { rs.base[3 * a + b] = R[a][b]; rs.base[12 + 3 * a + b] = Qx[a][b]; } }
  }
  const RelPoseR Rtrue{rs.base};                 // the true relative rotation Q2'Q1 (derivative epilogues)
  const RelPoseT t{rs.base};
  using RLoop = std::conditional_t<SPH1 || SPH2, IdentityR, RelPoseR>;
  const RLoop R{Rtrue};                          // the rotation the loop sees

  double y[NX];
  double M[NX][NX], idg[NX];
  double sz;
  int status = ST_OK;
  int cone_degree = A.n + B.n + (P1::D > 0 ? 1 : 0) + (P2::D > 0 ? 1 : 0);   // static_solver2.jl:124-130
  const double inv_deg = fast_rcp((double)cone_degree);

  // The complete iterate of a pair as a record of doubles with stride st: tag, k, s.z, y[NX], rows s, rows z, SOC s|z of
  // prim 1, SOC s|z of prim 2.  Used by the straggler hand-over (field-major records of a bin, tag = sorted position)
  // and by the opt-in warm start (one contiguous record per pair, tag = the pair's shape signature).
  auto store_iterate = [&](double* q, int64_t st, double tag, int k) {
    q[0] = tag; q += st;
    q[0] = (double)k; q += st;
    q[0] = sz; q += st;
#pragma unroll
    for (int a = 0; a < NX; a++) { q[0] = y[a]; q += st; }
    const int nr = A.n + B.n;
#pragma unroll 1
    for (int i = 0; i < nr; i++) { q[0] = rs(F_S, i); q += st; }
#pragma unroll 1
    for (int i = 0; i < nr; i++) { q[0] = rs(F_Z, i); q += st; }
    if constexpr (P1::D > 0) {
#pragma unroll
      for (int i = 0; i < P1::D; i++) { q[0] = A.s[i]; q += st; }
#pragma unroll
      for (int i = 0; i < P1::D; i++) { q[0] = A.z[i]; q += st; }
    }
    if constexpr (P2::D > 0) {
#pragma unroll
      for (int i = 0; i < P2::D; i++) { q[0] = B.s[i]; q += st; }
#pragma unroll
      for (int i = 0; i < P2::D; i++) { q[0] = B.z[i]; q += st; }
    }
  };
  auto load_iterate = [&](const double* q, int64_t st) {      // q points at the k field; returns k
    const int kk = (int)q[0]; q += st;
    sz = q[0]; q += st;
#pragma unroll
    for (int a = 0; a < NX; a++) { y[a] = q[0]; q += st; }
    const int nr = A.n + B.n;
#pragma unroll 1
    for (int i = 0; i < nr; i++) { rs(F_S, i) = q[0]; q += st; }
#pragma unroll 1
    for (int i = 0; i < nr; i++) { rs(F_Z, i) = q[0]; q += st; }
    if constexpr (P1::D > 0) {
#pragma unroll
      for (int i = 0; i < P1::D; i++) { A.s[i] = q[0]; q += st; }
#pragma unroll
      for (int i = 0; i < P1::D; i++) { A.z[i] = q[0]; q += st; }
    }
    if constexpr (P2::D > 0) {
#pragma unroll
      for (int i = 0; i < P2::D; i++) { B.s[i] = q[0]; q += st; }
#pragma unroll
      for (int i = 0; i < P2::D; i++) { B.z[i] = q[0]; q += st; }
    }
    return kk;
  };
  // Opt-in warm start (dcol_set_warm_start; not in the reference, excluded from parity): the pair's record holds the
  // iterate of its previous solve at the first iteration whose mu fell below warm_mu.  A record whose signature matches
  // the pair's shapes replaces the initial point; this solve leaves its own iterate behind for the next one.
  double* const ws = (p.warm && pair < p.warm_cap) ? p.warm + (int64_t)pair * p.warm_stride : nullptr;
  const double wsig = (double)(s1i * 65536 + s2i + 1);
  bool warm_saved = false;
  // A warm-started iterate does not satisfy this pair's linear constraints (the poses moved): its residual shrinks by
  // (1 - step) per iteration, and since the reference's only exit test is mu < pdip_tol, a warm-started pair must also
  // have taken steps that shrank it by 1e-5 before it may stop.  0 for cold starts: the default path is unaffected.
  double warm_res = 0.0;

  int k0 = 0;
  // ---------------- initialize (static_solver2.jl:86-103) ----------------
  if (tc.mode == 0 && ws && ws[0] == wsig) {
    (void)load_iterate(ws + 1, 1);
    ws[0] = 0.0;                       // consumed; rewritten below once this solve passes warm_mu
    warm_res = 1.0;
  } else if (tc.mode == 0) {
    double N1[NV1][NV1], g1[NV1], N2[NV2][NV2], g2[NV2];
    double y1[NV1], y2[NV2];
#pragma unroll
    for (int a = 0; a < NX; a++) y[a] = 0.0;
    PM::to_local(y, R, t, true, y1, y2);
    A.init_gram(y1, false, N1, g1);
    B.init_gram(y2, true, N2, g2);
    PM::assemble(N1, N2, R, M);
    double b0[NX], e4[NX], v2[NX];
    PM::assemble_vec(g1, g2, R, b0);
    bool ok = chol_upper<NX>(M, idg);
#pragma unroll
    for (int a = 0; a < NX; a++) { b0[a] = -b0[a]; e4[a] = (a == 3) ? 1.0 : 0.0; }
    chol_solve<NX>(M, idg, b0, y);      // x^ = (G'G)\(G'h)
    chol_solve<NX>(M, idg, e4, v2);     // z~ = G*((G'G)\(-c)) = C T v2
    double d1[NV1], d2[NV2];
    PM::to_local(y, R, t, true, y1, y2);
    PM::to_local(v2, R, t, false, d1, d2);
    double min_s = 1e300, min_z = 1e300, a_s = -1.0, a_z = -1.0;
    double ms = 1e300, mz = 1e300;
    A.init_rows(rs, y1, d1, ms, mz, a_s, a_z);
    B.init_rows(rs, y2, d2, ms, mz, a_s, a_z);
    min_s = ms; min_z = mz;
    if (A.n + B.n > 0) {   // bring2cone :62-64 — orthant part first, then max with the SOC residuals
      if (min_s <= 0.0) a_s = dmax(a_s, -min_s);
      if (min_z <= 0.0) a_z = dmax(a_z, -min_z);
    }
    double sh_s = (a_s < 0.0) ? 0.0 : 1.0 + a_s, sh_z = (a_z < 0.0) ? 0.0 : 1.0 + a_z;
    sz = A.init_shift(rs, sh_s, sh_z) + B.init_shift(rs, sh_s, sh_z);
    if (!ok) status = ST_NUMERICAL;
  } else {
    // ---------------- resume: reload the iterate parked by the first pass ----------------
    // record layout (field-major, stride = the bin's record capacity): sorted position, k, s.z, y[NX], rows s, rows z, SOC s|z of
    // prim 1, SOC s|z of prim 2; read with one walking pointer (no per-field 64-bit index arithmetic)
    k0 = load_iterate(tc.in + idx + tc.in_cap, tc.in_cap);
    warm_saved = sz * inv_deg < p.warm_mu;           // the first pass saved it when mu first fell below warm_mu
  }

  // ---------------- main loop (static_solver2.jl:132-187) ----------------
  // All warps of the CTA step through the iterations together (one block-wide vote per iteration):
  // the unrolled loop body is far larger than the SM's instruction cache, and warps in lockstep
  // share each fetched instruction line instead of streaming the loop once per warp.
  int it = 0;
  // loop invariants fetched once (a load behind the block-wide vote of every iteration sat on the critical path)
  const double tol = p.pdip_tol;
  const int max_it = p.max_iters, spill_t = tc.spill_t;
  bool active = (status == ST_OK) && !tc.dup;
  bool parked = tc.dup != 0;                // duplicate lanes write nothing
  if (active) status = ST_PDIP_FAILED;
  for (int k = k0;; k++) {
    const double mu = sz * inv_deg;
    // (a warm-started pair saves after its first step on the new poses, so that the record follows the trajectory)
    if (ws && active && !warm_saved && mu < p.warm_mu && (k > 0 || warm_res == 0.0)) { store_iterate(ws, 1, wsig, 0); warm_saved = true; }
    if (active) {
      // the reference tests mu only at the START of iterations 1..50 (static_solver2.jl:132,143-146) and raises
      // "pdip failed" after the 50th step without another test (:188): a pair that would pass on its 50th step fails
      if (k >= max_it) { active = false; it = k; }
      else if (!(mu == mu)) { status = ST_NUMERICAL; active = false; it = k; }
      else if (mu < tol && warm_res < 1e-5) { status = ST_OK; active = false; it = k; }
    }
    const int n_act = DCOL_COUNT(active);
    if (n_act == 0) break;
    if (n_act <= spill_t && k > k0) {
      // few stragglers left in this tile: park their iterates for the resume pass instead of keeping the whole CTA
      // (and its idle lanes) in the loop until the slowest pair converges
      if (active) {
        const int slot = atomicAdd(tc.ov_count, 1);
        store_iterate(tc.ov + slot, tc.ov_cap, (double)pos, k);
        parked = true;
      }
      break;
    }
    if (active) {
      double rhs[NX], dy[NX];
      double y1[NV1], y2[NV2];
      bool ok;
      {
        double N1[NV1][NV1], g1[NV1], N2[NV2][NV2], g2[NV2];
        PM::to_local(y, R, t, true, y1, y2);
        A.passA_rows(rs, y1, N1, g1);
        B.passA_rows(rs, y2, N2, g2);
        ok = A.passA_soc(y1, N1, g1);
        ok = B.passA_soc(y2, N2, g2) && ok;
        PM::assemble(N1, N2, R, M);
        PM::assemble_vec(g1, g2, R, rhs);
      }
      rhs[3] -= 1.0;                              // -c
      ok = chol_upper<NX>(M, idg) && ok;
      if (!ok) { status = ST_NUMERICAL; active = false; it = k; }
      else {
        chol_solve<NX>(M, idg, rhs, dy);            // affine direction
        double mh = 0.0, m = 0.0, dot1 = 0.0, dot2 = 0.0;
        PM::to_local(dy, R, t, false, y1, y2);
        A.passB_rows(rs, y1, mh, dot1, dot2);
        B.passB_rows(rs, y2, mh, dot1, dot2);
        A.passB_soc(y1, m, dot1, dot2);
        B.passB_soc(y2, m, dot1, dot2);
        const double aa = fast_rcp(dmax(1.0, dmax(mh + 0.5, m)));   // min(1, step to the boundary)  :159
#if DCOL_LS_SCALED
        const double rho = fma(aa * aa * dot2, fast_rcp(sz), 1.0 - aa);      // (s + a ds)'(z + a dz) / s'z with s'dz + z'ds = -s'z
#else
        const double rho = (sz + aa * (dot1 + aa * dot2)) * fast_rcp(sz);
#endif
        double sig = dmax(0.0, dmin(1.0, rho)); sig = sig * sig * sig;   // :161
        const double sigmu = sig * mu;
        {
          double g1[NV1], g2[NV2], rc[NX];
          A.passC_rows(rs, sigmu, g1);
          B.passC_rows(rs, sigmu, g2);
          A.passC_soc(sigmu, g1);
          B.passC_soc(sigmu, g2);
          PM::assemble_vec(g1, g2, R, rc);
#pragma unroll
          for (int a = 0; a < NX; a++) rhs[a] += rc[a];
        }
        chol_solve<NX>(M, idg, rhs, dy);            // combined direction
        m = 0.0;
        PM::to_local(dy, R, t, false, y1, y2);
        A.passD_rows(rs, y1, sigmu, m);
        B.passD_rows(rs, y2, sigmu, m);
        A.passD_soc(y1, m);
        B.passD_soc(y2, m);
        const double al = 0.99 * fast_rcp(dmax(1.0, m));   // min(1, 0.99 * step to the boundary)  :172
#pragma unroll
        for (int a = 0; a < NX; a++) y[a] += al * dy[a];
        warm_res *= 1.0 - al;
        sz = A.passE_rows(rs, al) + B.passE_rows(rs, al);
        sz += A.passE_soc(al) + B.passE_soc(al);
      }
    }
  }

  if (parked) return;      // finished (and written) by the resume pass

  // ---------------- outputs ----------------
  Frame F1, F2;
#pragma unroll
  for (int a = 0; a < 3; a++) { F1.r[a] = rs.base[21 + a];
#pragma unroll
    for (int b = 0; b < 3; b++) F1.Q[a][b] = rs.base[12 + 3 * a + b]; }
  p.alpha[oi] = y[3];
  if (p.x) {
#pragma unroll
    for (int a = 0; a < 3; a++) p.x[3 * oi + a] = F1.r[a] + F1.Q[a][0] * y[0] + F1.Q[a][1] * y[1] + F1.Q[a][2] * y[2];
  }
  if (p.iters) p.iters[oi] = it;
  if (p.mu) p.mu[oi] = sz * inv_deg;
  const int nt = 6 + 2 * na;
  if constexpr (VAR != Q_PROX && (SPH1 || SPH2)) {
    // The loop ran with the sphere's frame set equal to its partner's (IdentityR).  The derivative epilogues work with
    // the true frames, the true relative pose and body coordinates: convert the iterate (rotations of the pair point
    // and of the sphere's ball block by R = Q2'Q1; nothing else depends on the sphere's orientation).
    auto rot = [&](double (&v)[4], bool transpose) {
      const double a0 = v[1], a1 = v[2], a2 = v[3];
#pragma unroll
      for (int i = 0; i < 3; i++)
        v[1 + i] = transpose ? (Rtrue[0][i] * a0 + Rtrue[1][i] * a1 + Rtrue[2][i] * a2) : (Rtrue[i][0] * a0 + Rtrue[i][1] * a1 + Rtrue[i][2] * a2);
    };
    if constexpr (SPH2) {
      rot(B.s, false); rot(B.z, false);                                   // frame 1 -> the sphere's body frame
      const double t0 = t[0], t1 = t[1], t2 = t[2];                         // t was Q1'(r~1 - r~2); the epilogues want Q2'(r~1 - r~2) = R t
#pragma unroll
      for (int i = 0; i < 3; i++) rs.base[9 + i] = Rtrue[i][0] * t0 + Rtrue[i][1] * t1 + Rtrue[i][2] * t2;
    } else {
      rot(A.s, true); rot(A.z, true);                                     // frame 2 -> the sphere's body frame
      const double y0 = y[0], y1 = y[1], y2 = y[2];                         // xi: frame 2 -> frame 1 (R')
#pragma unroll
      for (int i = 0; i < 3; i++) y[i] = Rtrue[0][i] * y0 + Rtrue[1][i] * y1 + Rtrue[2][i] * y2;
      double Q1[3][3];                                                      // slots 12..20 held Q2; Q1 = Q2 R
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) Q1[i][j] = F1.Q[i][0] * Rtrue[0][j] + F1.Q[i][1] * Rtrue[1][j] + F1.Q[i][2] * Rtrue[2][j];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) F1.Q[i][j] = Q1[i][j];
    }
  }

  if constexpr (VAR == Q_GRAD) {
    // ---- proximity_gradient (src/proximity_gradient.jl:41-78): -z' du/dtheta
    {
      double* g = p.deriv + (int64_t)nt * pair;
      if (status == ST_OK) {
        frame2_from_relative(F1, Rtrue, t, F2);
        double y1[NV1], y2[NV2];
        PM::to_local(y, Rtrue, t, true, y1, y2);
        SideVecs V1, V2;
        A.side_vecs(rs, y1, V1); B.side_vecs(rs, y2, V2);
        grad_side(S1, F1, rec->a1, p.att_kind, V1, g + (swp ? 3 + na : 0));
        grad_side(S2, F2, rec->a2, p.att_kind, V2, g + (swp ? 0 : 3 + na));
      } else {
        for (int j = 0; j < nt; j++) g[j] = __longlong_as_double(0x7ff8000000000000LL);
      }
    }
  }
  if constexpr (VAR == Q_JAC) {
    // ---- proximity_jacobian (src/proximity.jl:71-134): reduced KKT solve in primitive-1's frame
    double* Jo = p.deriv + (int64_t)4 * nt * pair;
    bool ok = (status == ST_OK);
    if (ok) {
      frame2_from_relative(F1, Rtrue, t, F2);
      double K1[4][3], K2[4][3], Kb1[4][3], Kb2[4][3], Ye1[4][2], Ye2[4][2], Hw[4][3];
      SideVecs V1, V2;
      {
        double y1[NV1], y2[NV2];
        PM::to_local(y, Rtrue, t, true, y1, y2);
        A.side_vecs(rs, y1, V1); B.side_vecs(rs, y2, V2);
      }
      {
        double N1[NV1][NV1], N2[NV2][NV2], Nb1[NV1][3], Nb2[NV2][3];
        A.ift_gram(rs, N1, Nb1); B.ift_gram(rs, N2, Nb2);
        double Mt[NX][NX], Y[NX][4];        // Mt = A'^T : row b of Mt = column b of A'
#pragma unroll
        for (int b = 0; b < NX; b++) {
          double e[NX], y1[NV1], y2[NV2], n1[NV1], n2[NV2], col[NX];
#pragma unroll
          for (int a = 0; a < NX; a++) e[a] = (a == b) ? 1.0 : 0.0;
          PM::to_local(e, Rtrue, t, false, y1, y2);
#pragma unroll
          for (int a = 0; a < NV1; a++) { double acc = 0.0;
#pragma unroll
            for (int c = 0; c < NV1; c++) acc += N1[a][c] * y1[c];
            n1[a] = acc; }
#pragma unroll
          for (int a = 0; a < NV2; a++) { double acc = 0.0;
#pragma unroll
            for (int c = 0; c < NV2; c++) acc += N2[a][c] * y2[c];
            n2[a] = acc; }
          PM::assemble_vec(n1, n2, Rtrue, col);
#pragma unroll
          for (int a = 0; a < NX; a++) Mt[b][a] = col[a];
        }
#pragma unroll
        for (int a = 0; a < NX; a++)
#pragma unroll
          for (int i = 0; i < 4; i++) Y[a][i] = (a == i) ? 1.0 : 0.0;
        ok = ge_solve<NX, 4>(Mt, Y);
        // K'_k[i][c] = sum_l (T_k' Y)[l][i] N_k[l][c] ; H4[i][c] = Y[c][i] ; Ye'_k[i][m] = Y[extras_k m][i]
        double Kp1[4][3], Kp2[4][3], Kq1[4][3], Kq2[4][3], Yp1[4][2], Yp2[4][2];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          double yc[NX], y1[NV1], y2[NV2];
#pragma unroll
          for (int a = 0; a < NX; a++) yc[a] = Y[a][i];
          PM::to_local(yc, Rtrue, t, false, y1, y2);
#pragma unroll
          for (int c = 0; c < 3; c++) {
            double a1 = 0.0, a2 = 0.0, b1 = 0.0, b2 = 0.0;
#pragma unroll
            for (int l = 0; l < NV1; l++) { a1 += y1[l] * N1[l][c]; b1 += y1[l] * Nb1[l][c]; }
#pragma unroll
            for (int l = 0; l < NV2; l++) { a2 += y2[l] * N2[l][c]; b2 += y2[l] * Nb2[l][c]; }
            Kp1[i][c] = a1; Kp2[i][c] = a2; Kq1[i][c] = b1; Kq2[i][c] = b2;
          }
#pragma unroll
          for (int m = 0; m < 2; m++) {
            Yp1[i][m] = (m < P1::NE) ? y1[(4 + m < NV1) ? 4 + m : 0] : 0.0;
            Yp2[i][m] = (m < P2::NE) ? y2[(4 + m < NV2) ? 4 + m : 0] : 0.0;
          }
        }
        // rotate rows 0..2 into the world frame (P4 = blkdiag(Q1, 1)); Hw = P4 H4 Q1'
        double Hq[4][3];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int c = 0; c < 3; c++) Hq[i][c] = Y[0][i] * F1.Q[c][0] + Y[1][i] * F1.Q[c][1] + Y[2][i] * F1.Q[c][2];
#pragma unroll
        for (int c = 0; c < 3; c++) {
#pragma unroll
          for (int i = 0; i < 3; i++) {
            K1[i][c] = F1.Q[i][0] * Kp1[0][c] + F1.Q[i][1] * Kp1[1][c] + F1.Q[i][2] * Kp1[2][c];
            K2[i][c] = F1.Q[i][0] * Kp2[0][c] + F1.Q[i][1] * Kp2[1][c] + F1.Q[i][2] * Kp2[2][c];
            Kb1[i][c] = F1.Q[i][0] * Kq1[0][c] + F1.Q[i][1] * Kq1[1][c] + F1.Q[i][2] * Kq1[2][c];
            Kb2[i][c] = F1.Q[i][0] * Kq2[0][c] + F1.Q[i][1] * Kq2[1][c] + F1.Q[i][2] * Kq2[2][c];
            Hw[i][c] = F1.Q[i][0] * Hq[0][c] + F1.Q[i][1] * Hq[1][c] + F1.Q[i][2] * Hq[2][c];
          }
          K1[3][c] = Kp1[3][c]; K2[3][c] = Kp2[3][c]; Kb1[3][c] = Kq1[3][c]; Kb2[3][c] = Kq2[3][c]; Hw[3][c] = Hq[3][c];
        }
#pragma unroll
        for (int m = 0; m < 2; m++) {
#pragma unroll
          for (int i = 0; i < 3; i++) {
            Ye1[i][m] = F1.Q[i][0] * Yp1[0][m] + F1.Q[i][1] * Yp1[1][m] + F1.Q[i][2] * Yp1[2][m];
            Ye2[i][m] = F1.Q[i][0] * Yp2[0][m] + F1.Q[i][1] * Yp2[1][m] + F1.Q[i][2] * Yp2[2][m];
          }
          Ye1[3][m] = Yp1[3][m]; Ye2[3][m] = Yp2[3][m];
        }
      }
      if (ok) {
        jac_side<P1::NE>(S1, F1, rec->a1, p.att_kind, K1, Kb1, Ye1, Hw, V1, Jo + (swp ? 4 * (3 + na) : 0));
        jac_side<P2::NE>(S2, F2, rec->a2, p.att_kind, K2, Kb2, Ye2, Hw, V2, Jo + (swp ? 0 : 4 * (3 + na)));
      } else {
        status = ST_NUMERICAL;
      }
    }
    if (!ok) for (int j = 0; j < 4 * nt; j++) Jo[j] = __longlong_as_double(0x7ff8000000000000LL);
  }
  if (p.status) p.status[oi] = status;
}

// Out-of-line entry of one specialisation (same translation unit as the persistent kernel that calls it,
// dcol_persistent.cuh: whole-program compilation, so ptxas allocates registers across the call).
template <int C1, int C2, int VAR>
__device__ __noinline__ void solve_entry(const QueryParams* p, const TileCtx* tc, int idx) {
  solve_pair<C1, C2, VAR>(*p, *tc, idx);
}

}  // namespace dcol
