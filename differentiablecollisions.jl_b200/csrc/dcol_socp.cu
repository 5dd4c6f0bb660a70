// Generic batched conic solve on the device: the reference's
//     solve_socp(c, G, h, idx_ort, idx_soc1, idx_soc2; pdip_tol) -> (x, s, z)      (src/solvers/coneqp/static_solver2.jl:106-191)
// for n independent problems  min c'x  s.t.  h - G x in R+^n_ort x Q^d1 x Q^d2  of common dimensions, one thread per
// problem, dense arithmetic straight from (c, G, h).  This is NOT the proximity hot path (dcol_solver.cuh never forms
// G): it exists so that the solver subsystem can be driven on its own, e.g. on the reference's stored test problem
// (test/solver_tests.jl:38-64, example_socp_2.jld2), and makes no attempt at speed (local-memory arrays, runtime loops).
// Same algorithm: initial point :86-103 (note s~ = G x^ - h), Mehrotra predictor-corrector with Nesterov-Todd scaling
// (NT_scaling_chol_2.jl:95-127), line search to the cone boundary :2-51, 0.99 damping :172, mu-only termination tested at
// the start of iterations 1..50 :143-146, "pdip failed" after the 50th step :188.
#include "../../include/dcol_b200.h"
#include "dcol_common.cuh"
#include <cuda_runtime.h>

namespace dcol {

constexpr int G_MAXX = 16, G_MAXS = 96, G_MAXD = 8;

struct ConeDims { int nx, n_ort, d[2], ns; };

__device__ inline double soc_J(const double* u, int d) { double a = u[0] * u[0]; for (int i = 1; i < d; i++) a -= u[i] * u[i]; return a; }

// Nesterov-Todd scaling of one SOC block: W = eta * [w0 wv'; wv I + wv wv'/(1+w0)]
// W^-1 is applied through a Cholesky factorisation of the dense block, as the reference does (NT_scaling_chol_2.jl:16-29,
// 126): the closed form (1/eta) J Wbar J cancels catastrophically near the cone boundary and, at pdip_tol ~ 1e-12, lets
// the slack drift off h - G x at the 1e-5 level (the proximity kernels avoid this differently, by accumulating their SOC
// directions in the scaled space).
struct SocW { double w[G_MAXD]; double eta; int d; bool ok; double U[G_MAXD][G_MAXD]; };
__device__ inline void soc_nt(const double* s, const double* z, int d, SocW& W) {
  const double Js = soc_J(s, d), Jz = soc_J(z, d);
  W.d = d; W.ok = Js > 0.0 && Jz > 0.0;
  const double is = 1.0 / sqrt(Js), iz = 1.0 / sqrt(Jz);
  double dot = 0.0;
  for (int i = 0; i < d; i++) dot += (s[i] * is) * (z[i] * iz);
  const double gam = sqrt(0.5 * (1.0 + dot));
  W.w[0] = (s[0] * is + z[0] * iz) / (2.0 * gam);
  for (int i = 1; i < d; i++) W.w[i] = (s[i] * is - z[i] * iz) / (2.0 * gam);
  W.eta = sqrt(sqrt(Js / Jz));
  // dense W = eta * [w0 wv'; wv I + wv wv'/(1+w0)] and its upper Cholesky factor
  const double b = 1.0 / (1.0 + W.w[0]);
  for (int i = 0; i < d; i++) for (int j = i; j < d; j++)
    W.U[i][j] = W.eta * (i == 0 ? W.w[j] : ((i == j ? 1.0 : 0.0) + W.w[i] * W.w[j] * b));
  for (int j = 0; j < d && W.ok; j++) {
    double dd = W.U[j][j];
    for (int k = 0; k < j; k++) dd -= W.U[k][j] * W.U[k][j];
    if (!(dd > 0.0)) { W.ok = false; break; }
    dd = sqrt(dd); W.U[j][j] = dd;
    for (int i = j + 1; i < d; i++) { double a = W.U[j][i]; for (int k = 0; k < j; k++) a -= W.U[k][j] * W.U[k][i]; W.U[j][i] = a / dd; }
  }
}
__device__ inline void soc_mul(const SocW& W, const double* v, double* out, bool inverse) {      // out = W v  or  W^-1 v
  const int d = W.d;
  if (inverse) {          // U'U x = v
    double y[G_MAXD];
    for (int i = 0; i < d; i++) { double a = v[i]; for (int k = 0; k < i; k++) a -= W.U[k][i] * y[k]; y[i] = a / W.U[i][i]; }
    for (int i = d - 1; i >= 0; i--) { double a = y[i]; for (int k = i + 1; k < d; k++) a -= W.U[i][k] * out[k]; out[i] = a / W.U[i][i]; }
    return;
  }
  double c = 0.0;
  for (int i = 1; i < d; i++) c += W.w[i] * v[i];
  const double v0 = v[0];
  out[0] = W.eta * (W.w[0] * v0 + c);
  const double f = c / (1.0 + W.w[0]) + v0;
  for (int i = 1; i < d; i++) out[i] = W.eta * (v[i] + W.w[i] * f);
}

struct Scaling { double w_ort[G_MAXS]; SocW S[2]; };

__device__ inline void W_apply(const ConeDims& K, const Scaling& W, const double* v, double* out, bool inverse) {
  for (int i = 0; i < K.n_ort; i++) out[i] = inverse ? v[i] / W.w_ort[i] : v[i] * W.w_ort[i];
  int o = K.n_ort;
  for (int b = 0; b < 2; b++) if (K.d[b] > 0) { soc_mul(W.S[b], v + o, out + o, inverse); o += K.d[b]; }
}
// u o v  and  lam \ v  (soc_utils.jl:12-24, 52-67)
__device__ inline void cone_prod(const ConeDims& K, const double* u, const double* v, double* out) {
  for (int i = 0; i < K.n_ort; i++) out[i] = u[i] * v[i];
  int o = K.n_ort;
  for (int b = 0; b < 2; b++) if (K.d[b] > 0) {
    const int d = K.d[b]; double dot = 0.0;
    for (int i = 0; i < d; i++) dot += u[o + i] * v[o + i];
    const double u0 = u[o], v0 = v[o];
    for (int i = 1; i < d; i++) out[o + i] = u0 * v[o + i] + v0 * u[o + i];
    out[o] = dot; o += d;
  }
}
__device__ inline void cone_div(const ConeDims& K, const double* lam, const double* v, double* out) {
  for (int i = 0; i < K.n_ort; i++) out[i] = v[i] / lam[i];
  int o = K.n_ort;
  for (int b = 0; b < 2; b++) if (K.d[b] > 0) {
    const int d = K.d[b];
    double rho = lam[o] * lam[o], nu = 0.0;
    for (int i = 1; i < d; i++) { rho -= lam[o + i] * lam[o + i]; nu += lam[o + i] * v[o + i]; }
    const double l0 = lam[o], v0 = v[o];
    for (int i = 1; i < d; i++) out[o + i] = ((nu / l0 - v0) * lam[o + i] + (rho / l0) * v[o + i]) / rho;
    out[o] = (l0 * v0 - nu) / rho; o += d;
  }
}
// largest step <= 1 that keeps y + a*dl in the cone (static_solver2.jl:2-51)
__device__ inline double linesearch(const ConeDims& K, const double* y, const double* dl) {
  double a = 1.0;
  for (int i = 0; i < K.n_ort; i++) if (dl[i] < 0.0) a = fmin(a, -y[i] / dl[i]);
  int o = K.n_ort;
  for (int b = 0; b < 2; b++) if (K.d[b] > 0) {
    const int d = K.d[b];
    const double nu = soc_J(y + o, d);
    double zeta = y[o] * dl[o];
    for (int i = 1; i < d; i++) zeta -= y[o + i] * dl[o + i];
    const double sq = sqrt(nu), rho0 = zeta / nu;
    const double f = (zeta / sq + dl[o]) / (y[o] / sq + 1.0);
    double nr = 0.0;
    for (int i = 1; i < d; i++) { const double r = dl[o + i] / sq - f * y[o + i] / nu; nr += r * r; }
    nr = sqrt(nr);
    if (nr > rho0) a = fmin(a, 1.0 / (nr - rho0));
    o += d;
  }
  return a;
}
// r + (1 + a) e when r is outside or on the boundary of the cone (static_solver2.jl:52-85)
__device__ inline void bring2cone(const ConeDims& K, const double* r, double* out) {
  double a = -1.0;
  if (K.n_ort > 0) { double m = r[0]; for (int i = 1; i < K.n_ort; i++) m = fmin(m, r[i]); if (m <= 0.0) a = -m; }
  int o = K.n_ort;
  for (int b = 0; b < 2; b++) if (K.d[b] > 0) {
    double nv = 0.0; for (int i = 1; i < K.d[b]; i++) nv += r[o + i] * r[o + i];
    const double res = r[o] - sqrt(nv);
    if (res <= 0.0) a = fmax(a, -res);
    o += K.d[b];
  }
  for (int i = 0; i < K.ns; i++) out[i] = r[i];
  if (a >= 0.0) {
    for (int i = 0; i < K.n_ort; i++) out[i] += 1.0 + a;
    o = K.n_ort;
    for (int b = 0; b < 2; b++) if (K.d[b] > 0) { out[o] += 1.0 + a; o += K.d[b]; }
  }
}
// M = U'U (upper), in place; solve
__device__ inline bool chol(double (*M)[G_MAXX], int n) {
  for (int j = 0; j < n; j++) {
    double d = M[j][j];
    for (int k = 0; k < j; k++) d -= M[k][j] * M[k][j];
    if (!(d > 0.0)) return false;
    d = sqrt(d); M[j][j] = d;
    for (int i = j + 1; i < n; i++) { double a = M[j][i]; for (int k = 0; k < j; k++) a -= M[k][j] * M[k][i]; M[j][i] = a / d; }
  }
  return true;
}
__device__ inline void chol_solve(double (*U)[G_MAXX], int n, const double* b, double* x) {
  double y[G_MAXX];
  for (int i = 0; i < n; i++) { double a = b[i]; for (int k = 0; k < i; k++) a -= U[k][i] * y[k]; y[i] = a / U[i][i]; }
  for (int i = n - 1; i >= 0; i--) { double a = y[i]; for (int k = i + 1; k < n; k++) a -= U[i][k] * x[k]; x[i] = a / U[i][i]; }
}

__global__ void k_solve_socp(int64_t n, ConeDims K, const double* __restrict__ cs, const double* __restrict__ Gs, const double* __restrict__ hs,
                             double tol, int max_iters, double* __restrict__ xs, double* __restrict__ ss, double* __restrict__ zs,
                             int32_t* __restrict__ status, int32_t* __restrict__ iters, double* __restrict__ trace) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int nx = K.nx, ns = K.ns;
  const double* c = cs + p * nx; const double* G = Gs + p * (int64_t)ns * nx; const double* h = hs + p * ns;
  double x[G_MAXX], s[G_MAXS], z[G_MAXS], M[G_MAXX][G_MAXX], t1[G_MAXS], t2[G_MAXS];
  auto Gx = [&](const double* v, double* out) { for (int k = 0; k < ns; k++) { double a = 0.0; for (int j = 0; j < nx; j++) a += G[k * nx + j] * v[j]; out[k] = a; } };
  auto Gt = [&](const double* v, double* out) { for (int j = 0; j < nx; j++) { double a = 0.0; for (int k = 0; k < ns; k++) a += G[k * nx + j] * v[k]; out[j] = a; } };
  int st = DCOL_STATUS_OK, it = 0;
  // ---- initialize (:86-103)
  for (int i = 0; i < nx; i++) for (int j = i; j < nx; j++) { double a = 0.0; for (int k = 0; k < ns; k++) a += G[k * nx + i] * G[k * nx + j]; M[i][j] = a; }
  if (!chol(M, nx)) st = DCOL_STATUS_NUMERICAL;
  if (st == DCOL_STATUS_OK) {
    double b[G_MAXX], v[G_MAXX];
    Gt(h, b); chol_solve(M, nx, b, x);
    Gx(x, t1); for (int k = 0; k < ns; k++) t1[k] -= h[k];          // s~ = G x^ - h  (:94)
    bring2cone(K, t1, s);
    for (int j = 0; j < nx; j++) b[j] = -c[j];
    chol_solve(M, nx, b, v); Gx(v, t1);
    bring2cone(K, t1, z);
  }
  const int deg = K.n_ort + (K.d[0] > 0) + (K.d[1] > 0);
  Scaling W;
  double lam[G_MAXS], rz[G_MAXS], lds[G_MAXS], bz[G_MAXS], dsv[G_MAXS], dz[G_MAXS], rx[G_MAXX], rhs[G_MAXX], dx[G_MAXX];
  for (int iter = 1; st == DCOL_STATUS_OK; iter++) {
    if (iter > max_iters) { st = DCOL_STATUS_PDIP_FAILED; break; }
    // NT scaling, lambda = W z
    bool ok = true;
    for (int i = 0; i < K.n_ort; i++) { ok = ok && s[i] > 0.0 && z[i] > 0.0; W.w_ort[i] = sqrt(s[i] / z[i]); }
    { int o = K.n_ort; for (int b = 0; b < 2; b++) if (K.d[b] > 0) { soc_nt(s + o, z + o, K.d[b], W.S[b]); ok = ok && W.S[b].ok; o += K.d[b]; } }
    if (!ok) { st = DCOL_STATUS_NUMERICAL; break; }
    W_apply(K, W, z, lam, false);
    Gt(z, rx); for (int j = 0; j < nx; j++) rx[j] += c[j];
    Gx(x, rz); double sz = 0.0;
    for (int k = 0; k < ns; k++) { rz[k] += s[k] - h[k]; sz += s[k] * z[k]; }
    const double mu = sz / deg;
    if (!(mu == mu)) { st = DCOL_STATUS_NUMERICAL; break; }
    if (mu < tol) break;                                             // the only exit criterion (:144-146)
    // M = G' W^-2 G
    for (int i = 0; i < nx; i++) for (int j = i; j < nx; j++) M[i][j] = 0.0;
    for (int k = 0; k < K.n_ort; k++) { const double d = 1.0 / (W.w_ort[k] * W.w_ort[k]);
      for (int i = 0; i < nx; i++) { const double gi = d * G[k * nx + i]; for (int j = i; j < nx; j++) M[i][j] += gi * G[k * nx + j]; } }
    { int o = K.n_ort;
      for (int b = 0; b < 2; b++) if (K.d[b] > 0) {
        const int d = K.d[b];
        double col[G_MAXD], blk[G_MAXD][G_MAXX];
        for (int j = 0; j < nx; j++) { double in[G_MAXD]; for (int i = 0; i < d; i++) in[i] = G[(o + i) * nx + j]; soc_mul(W.S[b], in, col, true); for (int i = 0; i < d; i++) blk[i][j] = col[i]; }
        for (int i = 0; i < nx; i++) for (int j = i; j < nx; j++) { double a = 0.0; for (int r = 0; r < d; r++) a += blk[r][i] * blk[r][j]; M[i][j] += a; }
        o += d;
      } }
    if (!chol(M, nx)) { st = DCOL_STATUS_NUMERICAL; break; }
    // direction for a given lambda-scaled slack target lds:  bz = W^-1(-rz - W lds), dx = M^-1(-rx + G~' bz), dz = W^-1(G~ dx - bz), ds = W(lds - W dz)
    auto direction = [&](double* ds_out) {
      W_apply(K, W, lds, t1, false);
      for (int k = 0; k < ns; k++) t1[k] = -rz[k] - t1[k];
      W_apply(K, W, t1, bz, true);
      W_apply(K, W, bz, t1, true); Gt(t1, rhs);                                    // G~' bz = G' W^-1 bz
      for (int j = 0; j < nx; j++) rhs[j] -= rx[j];
      chol_solve(M, nx, rhs, dx);
      Gx(dx, t1); W_apply(K, W, t1, t2, true);                                     // G~ dx
      for (int k = 0; k < ns; k++) t2[k] -= bz[k];
      W_apply(K, W, t2, dz, true);
      W_apply(K, W, dz, t1, false);
      for (int k = 0; k < ns; k++) t1[k] = lds[k] - t1[k];
      W_apply(K, W, t1, ds_out, false);
    };
    // affine step (:149-156): lds = lam \ (-lam o lam) = -lam
    for (int k = 0; k < ns; k++) lds[k] = -lam[k];
    double ds[G_MAXS];
    direction(ds);
    const double aa = fmin(linesearch(K, s, ds), linesearch(K, z, dz));
    double num = 0.0;
    for (int k = 0; k < ns; k++) num += (s[k] + aa * ds[k]) * (z[k] + aa * dz[k]);
    double sig = fmax(0.0, fmin(1.0, num / sz)); sig = sig * sig * sig;
    // centering + correction (:164-169)
    W_apply(K, W, ds, t1, true); W_apply(K, W, dz, t2, false);
    cone_prod(K, t1, t2, dsv);
    cone_prod(K, lam, lam, t1);
    for (int k = 0; k < ns; k++) dsv[k] = -t1[k] - dsv[k];
    for (int i = 0; i < K.n_ort; i++) dsv[i] += sig * mu;
    { int o = K.n_ort; for (int b = 0; b < 2; b++) if (K.d[b] > 0) { dsv[o] += sig * mu; o += K.d[b]; } }
    cone_div(K, lam, dsv, lds);
    direction(ds);
    const double al = fmin(1.0, 0.99 * fmin(linesearch(K, s, ds), linesearch(K, z, dz)));
    for (int j = 0; j < nx; j++) x[j] += al * dx[j];
    for (int k = 0; k < ns; k++) { s[k] += al * ds[k]; z[k] += al * dz[k]; }
    it = iter;
    if (trace) {
      // the row the reference prints per iteration with verbose = true (static_solver2.jl:179-183):
      // c'x, s'z/(n_ort + 1), |G x + s - h|, |s o z + sigma mu e|, step
      double* tr = trace + (p * DCOL_TRACE_ROWS + (iter - 1)) * DCOL_TRACE_COLS;
      double obj = 0.0, gap = 0.0, pres = 0.0, cent = 0.0;
      for (int j = 0; j < nx; j++) obj += c[j] * x[j];
      Gx(x, t1);
      for (int k = 0; k < ns; k++) { const double r = t1[k] + s[k] - h[k]; pres += r * r; gap += s[k] * z[k]; }
      cone_prod(K, s, z, t2);
      for (int i = 0; i < K.n_ort; i++) t2[i] += sig * mu;
      { int o = K.n_ort; for (int b = 0; b < 2; b++) if (K.d[b] > 0) { t2[o] += sig * mu; o += K.d[b]; } }
      for (int k = 0; k < ns; k++) cent += t2[k] * t2[k];
      if (iter <= DCOL_TRACE_ROWS) { tr[0] = obj; tr[1] = gap / (K.n_ort + 1); tr[2] = sqrt(pres); tr[3] = sqrt(cent); tr[4] = al; }
    }
  }
  for (int j = 0; j < nx; j++) xs[p * nx + j] = x[j];
  for (int k = 0; k < ns; k++) { ss[p * ns + k] = s[k]; zs[p * ns + k] = z[k]; }
  if (status) status[p] = st;
  if (iters) iters[p] = it;
}

cudaError_t launch_solve_socp(int64_t n, int nx, int n_ort, int d1, int d2, const double* c, const double* G, const double* h, double tol, int max_iters,
                              double* x, double* s, double* z, int32_t* status, int32_t* iters, double* trace, cudaStream_t st) {
  ConeDims K; K.nx = nx; K.n_ort = n_ort; K.d[0] = d1; K.d[1] = d2; K.ns = n_ort + d1 + d2;
  k_solve_socp<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(n, K, c, G, h, tol, max_iters, x, s, z, status, iters, trace);
  return cudaGetLastError();
}
int socp_limit(int which) { return which == 0 ? G_MAXX : (which == 1 ? G_MAXS : G_MAXD); }
}  // namespace dcol
