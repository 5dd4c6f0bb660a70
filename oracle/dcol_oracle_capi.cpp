// ORACLE — TEST INFRASTRUCTURE ONLY (see dcol_oracle.hpp header).
// C ABI over the CPU restatement so pytest / bench.py's cpu_baseline leg can
// drive it through ctypes.  Batched entries mirror include/dcol_b200.h
// (same shape-table packing, same array layouts) minus the context handle.
#include "dcol_oracle.hpp"
#include <atomic>
#include <thread>
#include <vector>

using namespace dcol_oracle;

namespace {
struct Table {
  std::vector<Shape> shapes;
  bool ok = true;
  Table(int n_shapes, const int32_t* type, const int64_t* off, const double* params) {
    shapes.resize(n_shapes);
    for (int i = 0; i < n_shapes; i++) ok = ok && decode_shape(type[i], params + off[i], shapes[i]);
  }
};
int hw_threads() { unsigned h = std::thread::hardware_concurrency(); return h ? (int)h : 1; }
// dynamic-schedule parallel for over [0,n) in chunks (std::thread: no OpenMP runtime dependency)
template <class F>
void parallel_for(int64_t n, int nthreads, F&& body) {
  if (nthreads <= 0) nthreads = hw_threads();
  const int64_t chunk = 64;
  if (nthreads == 1 || n <= chunk) { for (int64_t i = 0; i < n; i++) body(i); return; }
  std::atomic<int64_t> next(0);
  auto worker = [&]() {
    for (;;) {
      int64_t b = next.fetch_add(chunk);
      if (b >= n) break;
      int64_t e = b + chunk < n ? b + chunk : n;
      for (int64_t i = b; i < e; i++) body(i);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; t++) th.emplace_back(worker);
  worker();
  for (auto& t : th) t.join();
}
}  // namespace

extern "C" {

int dcol_oracle_max_threads() { return hw_threads(); }
int dcol_oracle_max_faces() { return MAXF; }
void dcol_oracle_set_max_iters(int k) { max_iters_setting() = (k >= 1 && k <= 50) ? k : 50; }

// kind: 0 proximity, 1 gradient, 2 jacobian, 3 jacobian_slow (quaternion only in the reference, src/proximity.jl:154,198)
int dcol_oracle_query(int kind, int n_shapes, const int32_t* type, const int64_t* param_off, const double* params,
                      int64_t n, const int32_t* shape1, const int32_t* shape2,
                      const double* r1, const double* att1, const double* r2, const double* att2,
                      int att_kind, double pdip_tol, int nthreads,
                      double* alpha, double* x, double* deriv, int32_t* status, int32_t* iters) {
  Table T(n_shapes, type, param_off, params);
  if (!T.ok) return -1;
  int na = (att_kind == ATT_QUAT) ? 4 : 3;
  int nt = 6 + 2 * na;
  parallel_for(n, nthreads, [&](int64_t i) {
    const Shape& s1 = T.shapes[shape1[i]];
    const Shape& s2 = T.shapes[shape2[i]];
    Pose p1{r1 + 3 * i, att1 + na * i}, p2{r2 + 3 * i, att2 + na * i};
    double a = 0, xx[3] = {0, 0, 0}; int it = 0, st = 0;
    if (kind == 0) {
      st = proximity(s1, s2, p1, p2, att_kind, pdip_tol, &a, xx, &it);
    } else if (kind == 1) {
      double* g = deriv + (int64_t)nt * i;
      st = (nt == 14) ? proximity_gradient_t<14>(s1, s2, p1, p2, att_kind, pdip_tol, &a, g, &it)
                      : proximity_gradient_t<12>(s1, s2, p1, p2, att_kind, pdip_tol, &a, g, &it);
    } else {
      double* J = deriv + (int64_t)4 * nt * i;
      bool slow = (kind == 3);
      st = (nt == 14) ? proximity_jacobian_t<14>(s1, s2, p1, p2, att_kind, pdip_tol, &a, xx, J, &it, slow)
                      : proximity_jacobian_t<12>(s1, s2, p1, p2, att_kind, pdip_tol, &a, xx, J, &it, slow);
    }
    alpha[i] = a;
    if (x && kind != 1) { x[3 * i] = xx[0]; x[3 * i + 1] = xx[1]; x[3 * i + 2] = xx[2]; }
    if (status) status[i] = st;
    if (iters) iters[i] = it;
  });
  return 0;
}

// solve_socp on an explicit problem (row-major G ns x nx) — static_solver2.jl:106-191
int dcol_oracle_solve_socp(int nx, int n_ort, int d1, int d2, const double* c, const double* G, const double* h,
                           double pdip_tol, double* x, double* s, double* z, int32_t* iters) {
  static thread_local Problem<double> P; static thread_local Solution S;
  int ns = n_ort + d1 + d2;
  if (nx > MAXX || ns > MAXS || d1 > 4 || d2 > 4) return -1;
  P.nx = nx; P.ns = ns; P.n_ort = n_ort; P.d1 = d1; P.d2 = d2;
  for (int j = 0; j < nx; j++) P.c[j] = c[j];
  for (int k = 0; k < ns; k++) { P.h[k] = h[k]; for (int j = 0; j < nx; j++) P.G[k][j] = G[k * nx + j]; }
  solve_socp(P, pdip_tol, S);
  for (int j = 0; j < nx; j++) x[j] = S.x[j];
  for (int k = 0; k < ns; k++) { s[k] = S.s[k]; z[k] = S.z[k]; }
  if (iters) *iters = S.iters;
  return S.status;
}

// problem_matrices + combine for one pair: returns dims and row-major G (ns x nx), h
int dcol_oracle_pair_problem(int type1, const double* params1, int type2, const double* params2,
                             const double* r1, const double* att1, const double* r2, const double* att2, int att_kind,
                             int32_t* dims /*nx,ns,n_ort,d1,d2*/, double* G, double* h, double* c) {
  Shape s1, s2;
  if (!decode_shape(type1, params1, s1) || !decode_shape(type2, params2, s2)) return -1;
  static thread_local Problem<double> P;
  build_problem(s1, s2, Pose{r1, att1}, Pose{r2, att2}, att_kind, P);
  dims[0] = P.nx; dims[1] = P.ns; dims[2] = P.n_ort; dims[3] = P.d1; dims[4] = P.d2;
  for (int k = 0; k < P.ns; k++) { h[k] = P.h[k]; for (int j = 0; j < P.nx; j++) G[k * P.nx + j] = P.G[k][j]; }
  for (int j = 0; j < P.nx; j++) c[j] = P.c[j];
  return 0;
}

// combine_problem_matrices on explicit blocks (row-major) — combine_problem_matrices.jl:15-86
int dcol_oracle_combine(int L1, int d1, int v1, const double* Go1, const double* ho1, const double* Gs1, const double* hs1,
                        int L2, int d2, int v2, const double* Go2, const double* ho2, const double* Gs2, const double* hs2,
                        int32_t* dims, double* G, double* h, double* c) {
  static thread_local Block<double> B1, B2; static thread_local Problem<double> P;
  if (L1 > MAXF || L2 > MAXF || v1 > MAXV + 1 || v2 > MAXV + 1) return -1;
  if (v1 > MAXV || v2 > MAXV) return -2;
  B1.L = L1; B1.d = d1; B1.v = v1; B2.L = L2; B2.d = d2; B2.v = v2;
  for (int i = 0; i < L1; i++) { B1.h_ort[i] = ho1[i]; for (int j = 0; j < v1; j++) B1.G_ort[i][j] = Go1[i * v1 + j]; }
  for (int i = 0; i < d1; i++) { B1.h_soc[i] = hs1[i]; for (int j = 0; j < v1; j++) B1.G_soc[i][j] = Gs1[i * v1 + j]; }
  for (int i = 0; i < L2; i++) { B2.h_ort[i] = ho2[i]; for (int j = 0; j < v2; j++) B2.G_ort[i][j] = Go2[i * v2 + j]; }
  for (int i = 0; i < d2; i++) { B2.h_soc[i] = hs2[i]; for (int j = 0; j < v2; j++) B2.G_soc[i][j] = Gs2[i * v2 + j]; }
  combine_problem_matrices(B1, B2, P);
  dims[0] = P.nx; dims[1] = P.ns; dims[2] = P.n_ort; dims[3] = P.d1; dims[4] = P.d2;
  for (int k = 0; k < P.ns; k++) { h[k] = P.h[k]; for (int j = 0; j < P.nx; j++) G[k * P.nx + j] = P.G[k][j]; }
  for (int j = 0; j < P.nx; j++) c[j] = P.c[j];
  return 0;
}

// NT scaling ops for the nt_scaling_chol_2 tests: op 0: W\g, 1: W*g ; also returns dense W (ns x ns row-major) if Wd != NULL
int dcol_oracle_nt(int n_ort, int d1, int d2, const double* s, const double* z, int op, const double* g, double* out, double* Wd) {
  Cone K{n_ort, d1, d2};
  static thread_local NTScaling W;
  calc_NT_scalings(s, z, K, W);
  if (!W.ok) return ST_NUMERICAL;
  if (op == 0) nt_ldiv(W, g, out); else nt_mul(W, g, out);
  if (Wd) {
    int ns = K.ns();
    for (int i = 0; i < ns * ns; i++) Wd[i] = 0.0;
    for (int i = 0; i < n_ort; i++) Wd[i * ns + i] = W.ort[i];
    for (int i = 0; i < d1; i++) for (int j = 0; j < d1; j++) Wd[(n_ort + i) * ns + n_ort + j] = W.soc1[i][j];
    for (int i = 0; i < d2; i++) for (int j = 0; j < d2; j++) Wd[(n_ort + d1 + i) * ns + n_ort + d1 + j] = W.soc2[i][j];
  }
  return 0;
}

// cone algebra: op 0 cone_product(u,v), 1 inverse_cone_product(u,w), 2 gen_e, 3 linesearch(u, v) -> out[0], 4 bring2cone(u)
int dcol_oracle_cone_op(int n_ort, int d1, int d2, int op, const double* u, const double* v, double* out) {
  Cone K{n_ort, d1, d2};
  switch (op) {
    case 0: cone_product<double>(u, v, K, out); break;
    case 1: inverse_cone_product(u, v, K, out); break;
    case 2: gen_e(K, out); break;
    case 3: out[0] = linesearch(u, v, K); break;
    case 4: bring2cone(u, K, out); break;
    default: return -1;
  }
  return 0;
}

// attitude maps: op 0 dcm_from_q, 1 dcm_from_mrp, 2 mrp_from_q, 3 q_from_mrp
int dcol_oracle_att(int op, const double* in, double* out) {
  double Q[3][3];
  switch (op) {
    case 0: dcm_from_q<double>(in, Q); for (int i = 0; i < 9; i++) out[i] = Q[i / 3][i % 3]; break;
    case 1: dcm_from_mrp<double>(in, Q); for (int i = 0; i < 9; i++) out[i] = Q[i / 3][i % 3]; break;
    case 2: mrp_from_q(in, out); break;
    case 3: q_from_mrp(in, out); break;
    default: return -1;
  }
  return 0;
}

}  // extern "C"
