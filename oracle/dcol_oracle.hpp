// ORACLE — TEST INFRASTRUCTURE ONLY.  NOT part of the shipped library.
//
// CPU (C++17, FP64) restatement of the proximity hot path of
// kevin-tracy/DifferentiableCollisions.jl v0.1.5, written by following the
// reference files function by function (citations are relative to the
// reference checkout, e.g. src/proximity.jl:3-14).  It keeps every quirk of
// the reference: s~ = G*x - h in `initialize`,
// mu-only termination, 0.99 step damping, 50 iteration cap, W\ applied through
// a Cholesky factor of the dense SOC scaling block, IFT with the solver slack.
//
// PARITY UNPINNED: Julia is not available in this image and the reference
// ships no golden output vectors, so this oracle is pinned only to (i) the
// reference's own self-consistency tests restated in tests/, (ii) the decoded
// JLD2 input fixtures, (iii) the known-answer tables of SURVEY.md Appendix C
// (an independent NumPy restatement), (iv) analytic cases.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may call this code.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace dcol_oracle {

constexpr int MAXF = 128;           // max faces of a polytope / sides of a polygon
constexpr int MAXV = 6;             // max per-primitive variables (polygon: x,alpha,y1,y2)
constexpr int MAXX = 8;             // max pair variables (polygon-polygon)
constexpr int MAXS = 2 * MAXF + 8;  // max cone rows of a pair
constexpr int MAXT = 14;            // pose parameters of a pair (quaternion)

enum ShapeType { POLYTOPE = 0, CAPSULE = 1, CYLINDER = 2, CONE = 3, SPHERE = 4, ELLIPSOID = 5, POLYGON = 6 };
enum AttKind { ATT_QUAT = 0, ATT_MRP = 1 };
enum Status { ST_OK = 0, ST_PDIP_FAILED = 1, ST_NUMERICAL = 2 };

// ---------------------------------------------------------------------------
// forward-mode dual number (stands in for ForwardDiff.Dual; ForwardDiff is an
// un-vendored dependency, compat 0.10 — only exact forward-mode semantics are
// needed: src/proximity.jl:88,120,152 and src/proximity_gradient.jl:58,77)
// ---------------------------------------------------------------------------
template <int N>
struct Dual {
  double v;
  double d[N];
  Dual() : v(0.0) { for (int i = 0; i < N; i++) d[i] = 0.0; }
  Dual(double x) : v(x) { for (int i = 0; i < N; i++) d[i] = 0.0; }
};
template <int N> inline Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v + b.v; for (int i = 0; i < N; i++) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v - b.v; for (int i = 0; i < N; i++) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a) { Dual<N> r; r.v = -a.v; for (int i = 0; i < N; i++) r.d[i] = -a.d[i]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v * b.v; for (int i = 0; i < N; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; double inv = 1.0 / b.v; r.v = a.v * inv; for (int i = 0; i < N; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv; return r; }
template <int N> inline Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> r = a; r.v += b; return r; }
template <int N> inline Dual<N> operator+(double b, const Dual<N>& a) { Dual<N> r = a; r.v += b; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> inline Dual<N> operator-(double b, const Dual<N>& a) { Dual<N> r = -a; r.v += b; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, double b) { Dual<N> r; r.v = a.v * b; for (int i = 0; i < N; i++) r.d[i] = a.d[i] * b; return r; }
template <int N> inline Dual<N> operator*(double b, const Dual<N>& a) { return a * b; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, double b) { return a * (1.0 / b); }
template <int N> inline Dual<N> operator/(double a, const Dual<N>& b) { return Dual<N>(a) / b; }
template <int N> inline Dual<N> sqrt(const Dual<N>& a) { Dual<N> r; r.v = std::sqrt(a.v); double k = 0.5 / r.v; for (int i = 0; i < N; i++) r.d[i] = a.d[i] * k; return r; }
inline double value(double a) { return a; }
template <int N> inline double value(const Dual<N>& a) { return a.v; }
using std::sqrt;

// ---------------------------------------------------------------------------
// shape table (same packing as include/dcol_b200.h)
// ---------------------------------------------------------------------------
struct Shape {
  int type;
  int n;               // faces (polytope) / sides (polygon); 0 otherwise
  const double* A;     // polytope n x 3 row-major; polygon n x 2 row-major
  const double* b;     // n
  double p0, p1;       // capsule/cylinder {R,L}; cone {H,beta}; sphere {R}; polygon {R}
  const double* U;     // ellipsoid 3x3 row-major (= cholesky(P).U, src/primitives.jl:90)
  const double* r_off; // 3
  const double* Q_off; // 3x3 row-major
};

inline int shape_param_count(int type, const double* p) {
  switch (type) {
    case POLYTOPE: return 1 + 4 * (int)p[0] + 12;
    case CAPSULE: case CYLINDER: case CONE: return 2 + 12;
    case SPHERE: return 1 + 12;
    case ELLIPSOID: return 9 + 12;
    case POLYGON: return 1 + 3 * (int)p[0] + 1 + 12;
  }
  return -1;
}

inline bool decode_shape(int type, const double* p, Shape& s) {
  s.type = type; s.n = 0; s.A = s.b = s.U = nullptr; s.p0 = s.p1 = 0.0;
  const double* tail = nullptr;
  switch (type) {
    case POLYTOPE: s.n = (int)p[0]; s.A = p + 1; s.b = p + 1 + 3 * s.n; tail = p + 1 + 4 * s.n; break;
    case CAPSULE: case CYLINDER: case CONE: s.p0 = p[0]; s.p1 = p[1]; tail = p + 2; break;
    case SPHERE: s.p0 = p[0]; tail = p + 1; break;
    case ELLIPSOID: s.U = p; tail = p + 9; break;
    case POLYGON: s.n = (int)p[0]; s.A = p + 1; s.b = p + 1 + 2 * s.n; s.p0 = p[1 + 3 * s.n]; tail = p + 2 + 3 * s.n; break;
    default: return false;
  }
  if (s.n < 0 || s.n > MAXF) return false;
  s.r_off = tail; s.Q_off = tail + 3;
  return true;
}

// ---------------------------------------------------------------------------
// attitude maps — src/problem_matrices.jl:1-27
// ---------------------------------------------------------------------------
inline void mrp_from_q(const double q[4], double p[3]) {  // :1-3
  for (int i = 0; i < 3; i++) p[i] = q[i + 1] / (1 + q[0]);
}
inline void q_from_mrp(const double p[3], double q[4]) {  // :5-7
  double pp = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
  double k = 1 / (1 + pp);
  q[0] = k * (1 - pp);
  for (int i = 0; i < 3; i++) q[i + 1] = k * 2 * p[i];
}
template <class T>
inline void dcm_from_q(const T q[4], T Q[3][3]) {  // :9-16 (normalises q first)
  T nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  T q4 = q[0] / nrm, q1 = q[1] / nrm, q2 = q[2] / nrm, q3 = q[3] / nrm;
  Q[0][0] = 2 * q1 * q1 + 2 * q4 * q4 - 1; Q[0][1] = 2 * (q1 * q2 - q3 * q4);       Q[0][2] = 2 * (q1 * q3 + q2 * q4);
  Q[1][0] = 2 * (q1 * q2 + q3 * q4);       Q[1][1] = 2 * q2 * q2 + 2 * q4 * q4 - 1; Q[1][2] = 2 * (q2 * q3 - q1 * q4);
  Q[2][0] = 2 * (q1 * q3 - q2 * q4);       Q[2][1] = 2 * (q2 * q3 + q1 * q4);       Q[2][2] = 2 * q3 * q3 + 2 * q4 * q4 - 1;
}
template <class T>
inline void dcm_from_mrp(const T p[3], T Q[3][3]) {  // :18-27
  T p1 = p[0], p2 = p[1], p3 = p[2];
  T den1 = p1 * p1 + p2 * p2 + p3 * p3 + 1;
  T den = den1 * den1;
  T a = 4 * p1 * p1 + 4 * p2 * p2 + 4 * p3 * p3 - 4;
  Q[0][0] = -((8 * p2 * p2 + 8 * p3 * p3) / den - 1) * den; Q[0][1] = 8 * p1 * p2 + p3 * a; Q[0][2] = 8 * p1 * p3 - p2 * a;
  Q[1][0] = 8 * p1 * p2 - p3 * a; Q[1][1] = -((8 * p1 * p1 + 8 * p3 * p3) / den - 1) * den; Q[1][2] = 8 * p2 * p3 + p1 * a;
  Q[2][0] = 8 * p1 * p3 + p2 * a; Q[2][1] = 8 * p2 * p3 - p1 * a; Q[2][2] = -((8 * p1 * p1 + 8 * p2 * p2) / den - 1) * den;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Q[i][j] = Q[i][j] / den;
}

// ---------------------------------------------------------------------------
// per-primitive cone blocks — src/problem_matrices.jl:30-191
// ---------------------------------------------------------------------------
template <class T>
struct Block {
  int L, d, v;                    // orthant rows, SOC dimension, variables
  T G_ort[MAXF][MAXV]; T h_ort[MAXF];
  T G_soc[4][MAXV];    T h_soc[4];
};

template <class T>
inline void problem_matrices(const Shape& sh, const T r_in[3], const T* att, int att_kind, Block<T>& B) {
  // dispatch wrappers (e.g. :40-53): n_Q_b from attitude, r + n_Q_b*r_offset, n_Q_b*Q_offset
  T Qn[3][3];
  if (att_kind == ATT_QUAT) dcm_from_q(att, Qn); else dcm_from_mrp(att, Qn);
  T r[3], Q[3][3];
  for (int i = 0; i < 3; i++) {
    r[i] = r_in[i] + (Qn[i][0] * sh.r_off[0] + Qn[i][1] * sh.r_off[1] + Qn[i][2] * sh.r_off[2]);
    for (int j = 0; j < 3; j++)
      Q[i][j] = Qn[i][0] * sh.Q_off[0 * 3 + j] + Qn[i][1] * sh.Q_off[1 * 3 + j] + Qn[i][2] * sh.Q_off[2 * 3 + j];
  }
  auto zero_rows = [&](int L, int d, int v) {
    B.L = L; B.d = d; B.v = v;
    for (int i = 0; i < L; i++) { B.h_ort[i] = T(0.0); for (int j = 0; j < v; j++) B.G_ort[i][j] = T(0.0); }
    for (int i = 0; i < d; i++) { B.h_soc[i] = T(0.0); for (int j = 0; j < v; j++) B.G_soc[i][j] = T(0.0); }
  };
  switch (sh.type) {
    case CAPSULE: {  // :30-39
      double R = sh.p0, Lh = sh.p1;
      zero_rows(2, 4, 5);
      T bx[3] = {Q[0][0], Q[1][0], Q[2][0]};
      B.G_soc[0][3] = T(-R);
      for (int i = 0; i < 3; i++) { B.G_soc[1 + i][i] = T(-1.0); B.G_soc[1 + i][4] = bx[i]; B.h_soc[1 + i] = -r[i]; }
      B.G_ort[0][3] = T(-Lh / 2); B.G_ort[0][4] = T(1.0);
      B.G_ort[1][3] = T(-Lh / 2); B.G_ort[1][4] = T(-1.0);
    } break;
    case CYLINDER: {  // :146-160
      double R = sh.p0, Lh = sh.p1;
      zero_rows(4, 4, 5);
      T bx[3] = {Q[0][0], Q[1][0], Q[2][0]};
      B.G_soc[0][3] = T(-R);
      for (int i = 0; i < 3; i++) { B.G_soc[1 + i][i] = T(-1.0); B.G_soc[1 + i][4] = bx[i]; B.h_soc[1 + i] = -r[i]; }
      B.G_ort[0][3] = T(-Lh / 2); B.G_ort[0][4] = T(1.0);
      B.G_ort[1][3] = T(-Lh / 2); B.G_ort[1][4] = T(-1.0);
      T bxr = bx[0] * r[0] + bx[1] * r[1] + bx[2] * r[2];
      for (int j = 0; j < 3; j++) { B.G_ort[2][j] = -bx[j]; B.G_ort[3][j] = bx[j]; }
      B.G_ort[2][3] = T(-Lh / 2); B.G_ort[3][3] = T(-Lh / 2);
      B.h_ort[2] = -bxr; B.h_ort[3] = bxr;
    } break;
    case CONE: {  // :56-66
      double H = sh.p0, beta = sh.p1;
      double tb = std::tan(beta);
      zero_rows(1, 3, 4);
      T bx[3] = {Q[0][0], Q[1][0], Q[2][0]};
      double E[3] = {tb, 1.0, 1.0};
      for (int i = 0; i < 3; i++) {       // EQt = E*Q' ; h_soc = -EQt*r ; G_soc = [-EQt, -[tb*3H/4,0,0]]
        T acc = T(0.0);
        for (int j = 0; j < 3; j++) { T eqt = E[i] * Q[j][i]; B.G_soc[i][j] = -eqt; acc = acc + eqt * r[j]; }
        B.h_soc[i] = -acc;
      }
      B.G_soc[0][3] = T(-(tb * 3 * H / 4));
      for (int j = 0; j < 3; j++) B.G_ort[0][j] = bx[j];
      B.G_ort[0][3] = T(-H / 4);
      B.h_ort[0] = bx[0] * r[0] + bx[1] * r[1] + bx[2] * r[2];
    } break;
    case POLYTOPE: {  // :79-88
      int n = sh.n;
      zero_rows(n, 0, 4);
      for (int i = 0; i < n; i++) {        // AQt = A*Q' ; G_ort = [AQt -b] ; h_ort = AQt*r
        T acc = T(0.0);
        for (int j = 0; j < 3; j++) {
          T aq = sh.A[i * 3 + 0] * Q[j][0] + sh.A[i * 3 + 1] * Q[j][1] + sh.A[i * 3 + 2] * Q[j][2];
          B.G_ort[i][j] = aq; acc = acc + aq * r[j];
        }
        B.G_ort[i][3] = T(-sh.b[i]);
        B.h_ort[i] = acc;
      }
    } break;
    case SPHERE: {  // :99-113
      zero_rows(0, 4, 4);
      B.G_soc[0][3] = T(-sh.p0);
      for (int i = 0; i < 3; i++) { B.G_soc[1 + i][i] = T(-1.0); B.h_soc[1 + i] = -r[i]; }
    } break;
    case POLYGON: {  // :124-134
      int n = sh.n; double R = sh.p0;
      zero_rows(n, 4, 6);
      for (int i = 0; i < n; i++) { B.G_ort[i][3] = T(-sh.b[i]); B.G_ort[i][4] = T(sh.A[i * 2 + 0]); B.G_ort[i][5] = T(sh.A[i * 2 + 1]); }
      B.G_soc[0][3] = T(-R);
      for (int i = 0; i < 3; i++) { B.G_soc[1 + i][i] = T(-1.0); B.G_soc[1 + i][4] = Q[i][0]; B.G_soc[1 + i][5] = Q[i][1]; B.h_soc[1 + i] = -r[i]; }
    } break;
    case ELLIPSOID: {  // :172-182
      zero_rows(0, 4, 4);
      B.G_soc[0][3] = T(-1.0);
      for (int i = 0; i < 3; i++) {        // -U*Q' and -U*Q'*r
        T acc = T(0.0);
        for (int j = 0; j < 3; j++) {
          T uq = sh.U[i * 3 + 0] * Q[j][0] + sh.U[i * 3 + 1] * Q[j][1] + sh.U[i * 3 + 2] * Q[j][2];
          B.G_soc[1 + i][j] = -uq; acc = acc + uq * r[j];
        }
        B.h_soc[1 + i] = -acc;
      }
    } break;
  }
}

// ---------------------------------------------------------------------------
// pair assembly — src/combine_problem_matrices.jl:15-86
// rows [ort1; ort2; soc1; soc2]; primitive-1 extras at columns 5.., primitive-2
// extras after them (:73-77); c = e4 (:25)
// ---------------------------------------------------------------------------
template <class T>
struct Problem {
  int nx, ns, n_ort, d1, d2;
  T G[MAXS][MAXX]; T h[MAXS]; double c[MAXX];
};

template <class T>
inline void combine_problem_matrices(const Block<T>& B1, const Block<T>& B2, Problem<T>& P) {
  int e1 = B1.v - 4, e2 = B2.v - 4;
  P.nx = 4 + e1 + e2; P.n_ort = B1.L + B2.L; P.d1 = B1.d; P.d2 = B2.d; P.ns = P.n_ort + P.d1 + P.d2;
  for (int j = 0; j < P.nx; j++) P.c[j] = (j == 3) ? 1.0 : 0.0;
  int row = 0;
  auto put1 = [&](const T* g, const T& hv) {
    for (int j = 0; j < P.nx; j++) P.G[row][j] = T(0.0);
    for (int j = 0; j < B1.v; j++) P.G[row][j] = g[j];
    P.h[row] = hv; row++;
  };
  auto put2 = [&](const T* g, const T& hv) {
    for (int j = 0; j < P.nx; j++) P.G[row][j] = T(0.0);
    for (int j = 0; j < 4; j++) P.G[row][j] = g[j];
    for (int j = 0; j < e2; j++) P.G[row][4 + e1 + j] = g[4 + j];
    P.h[row] = hv; row++;
  };
  for (int i = 0; i < B1.L; i++) put1(B1.G_ort[i], B1.h_ort[i]);
  for (int i = 0; i < B2.L; i++) put2(B2.G_ort[i], B2.h_ort[i]);
  for (int i = 0; i < B1.d; i++) put1(B1.G_soc[i], B1.h_soc[i]);
  for (int i = 0; i < B2.d; i++) put2(B2.G_soc[i], B2.h_soc[i]);
}

// ---------------------------------------------------------------------------
// cone algebra — src/solvers/coneqp/soc_utils.jl
// a "Cone" describes the row layout [ort | soc1 | soc2] (Appendix D-5)
// ---------------------------------------------------------------------------
struct Cone { int n_ort, d1, d2; int ns() const { return n_ort + d1 + d2; } };

template <class T>
inline void soc_cone_product(const T* u, const T* v, int n, T* out) {  // :12-24
  if (n <= 0) return;
  T dot = T(0.0);
  for (int i = 0; i < n; i++) dot = dot + u[i] * v[i];
  out[0] = dot;
  for (int i = 1; i < n; i++) out[i] = u[0] * v[i] + v[0] * u[i];
}
template <class T>
inline void cone_product(const T* s, const T* z, const Cone& K, T* out) {  // :38-51
  for (int i = 0; i < K.n_ort; i++) out[i] = s[i] * z[i];
  soc_cone_product(s + K.n_ort, z + K.n_ort, K.d1, out + K.n_ort);
  soc_cone_product(s + K.n_ort + K.d1, z + K.n_ort + K.d1, K.d2, out + K.n_ort + K.d1);
}
inline void inverse_soc_cone_product(const double* u, const double* w, int n, double* out) {  // :52-67
  if (n <= 0) return;
  double u0 = u[0], w0 = w[0];
  double rho = u0 * u0, nu = 0.0;
  double uu = 0.0;
  for (int i = 1; i < n; i++) { uu += u[i] * u[i]; nu += u[i] * w[i]; }
  rho -= uu;
  double ir = 1 / rho;
  out[0] = ir * (u0 * w0 - nu);
  for (int i = 1; i < n; i++) out[i] = ir * ((nu / u0 - w0) * u[i] + (rho / u0) * w[i]);
}
inline void inverse_cone_product(const double* l, const double* v, const Cone& K, double* out) {  // :83-99
  for (int i = 0; i < K.n_ort; i++) out[i] = v[i] / l[i];
  inverse_soc_cone_product(l + K.n_ort, v + K.n_ort, K.d1, out + K.n_ort);
  inverse_soc_cone_product(l + K.n_ort + K.d1, v + K.n_ort + K.d1, K.d2, out + K.n_ort + K.d1);
}
inline void gen_e(const Cone& K, double* e) {  // :108-127
  for (int i = 0; i < K.n_ort; i++) e[i] = 1.0;
  for (int i = 0; i < K.d1; i++) e[K.n_ort + i] = (i == 0) ? 1.0 : 0.0;
  for (int i = 0; i < K.d2; i++) e[K.n_ort + K.d1 + i] = (i == 0) ? 1.0 : 0.0;
}
inline double soc_quad_J(const double* x, int n) {  // :132-136
  double a = 0.0;
  for (int i = 1; i < n; i++) a += x[i] * x[i];
  return x[0] * x[0] - a;
}
inline void arrow(const double* x, int n, double M[4][4]) {  // :2-10
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) M[i][j] = 0.0;
  for (int j = 0; j < n; j++) M[0][j] = x[j];
  for (int i = 1; i < n; i++) { M[i][0] = x[i]; M[i][i] = x[0]; }
}

// ---------------------------------------------------------------------------
// small dense factorizations (StaticArrays `cholesky`, `Cholesky \`, LU `\`;
// StaticArrays is un-vendored, compat 1.5 — standard algorithms restated)
// ---------------------------------------------------------------------------
// upper Cholesky A = U'U in place in U (n<=MAXX); returns false if not PD
template <int LD>
inline bool chol_upper(const double A[][LD], int n, double U[][LD]) {
  for (int j = 0; j < n; j++) {
    double d = A[j][j];
    for (int k = 0; k < j; k++) d -= U[k][j] * U[k][j];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    U[j][j] = d;
    for (int i = j + 1; i < n; i++) {
      double a = A[j][i];
      for (int k = 0; k < j; k++) a -= U[k][j] * U[k][i];
      U[j][i] = a / d;
    }
    for (int i = 0; i < j; i++) U[j][i] = 0.0;
  }
  return true;
}
template <int LD>
inline void chol_solve(const double U[][LD], int n, const double* b, double* x) {  // x = (U'U)\b
  double y[MAXX];
  for (int i = 0; i < n; i++) { double a = b[i]; for (int k = 0; k < i; k++) a -= U[k][i] * y[k]; y[i] = a / U[i][i]; }
  for (int i = n - 1; i >= 0; i--) { double a = y[i]; for (int k = i + 1; k < n; k++) a -= U[i][k] * x[k]; x[i] = a / U[i][i]; }
}
// LU with partial pivoting, solves A X = B for nrhs right-hand sides (in place in B)
inline bool lu_solve(double A[MAXX][MAXX], int n, double B[MAXX][MAXT], int nrhs) {
  for (int k = 0; k < n; k++) {
    int p = k; double best = std::fabs(A[k][k]);
    for (int i = k + 1; i < n; i++) if (std::fabs(A[i][k]) > best) { best = std::fabs(A[i][k]); p = i; }
    if (best == 0.0 || best != best) return false;
    if (p != k) {
      for (int j = 0; j < n; j++) { double t = A[k][j]; A[k][j] = A[p][j]; A[p][j] = t; }
      for (int j = 0; j < nrhs; j++) { double t = B[k][j]; B[k][j] = B[p][j]; B[p][j] = t; }
    }
    for (int i = k + 1; i < n; i++) {
      double f = A[i][k] / A[k][k];
      A[i][k] = f;
      for (int j = k + 1; j < n; j++) A[i][j] -= f * A[k][j];
      for (int j = 0; j < nrhs; j++) B[i][j] -= f * B[k][j];
    }
  }
  for (int i = n - 1; i >= 0; i--)
    for (int j = 0; j < nrhs; j++) {
      double a = B[i][j];
      for (int k = i + 1; k < n; k++) a -= A[i][k] * B[k][j];
      B[i][j] = a / A[i][i];
    }
  return true;
}

// ---------------------------------------------------------------------------
// Nesterov-Todd scaling — src/solvers/coneqp/NT_scaling_chol_2.jl
// ---------------------------------------------------------------------------
struct NTScaling {        // NT_scaling_2 :2-8
  Cone K;
  double ort[2 * MAXF];
  double soc1[4][4], soc1_fact[4][4];
  double soc2[4][4], soc2_fact[4][4];
  bool ok;
};
inline void soc_NT_scaling(const double* s, const double* z, int n, double W[4][4]) {  // :95-112
  if (n <= 0) return;
  double Jz = std::sqrt(soc_quad_J(z, n)), Js = std::sqrt(soc_quad_J(s, n));
  double zb[4], sb[4], wb[4];
  for (int i = 0; i < n; i++) { zb[i] = z[i] / Jz; sb[i] = s[i] / Js; }
  double dot = 0.0;
  for (int i = 0; i < n; i++) dot += zb[i] * sb[i];
  double gamma = std::sqrt((1 + dot) / 2);
  double k = 1 / (2 * gamma);
  wb[0] = k * (sb[0] + zb[0]);
  for (int i = 1; i < n; i++) wb[i] = k * (sb[i] - zb[i]);
  double b = 1 / (wb[0] + 1);
  double eta = std::pow(soc_quad_J(s, n) / soc_quad_J(z, n), 0.25);
  for (int j = 0; j < n; j++) W[0][j] = eta * wb[j];
  for (int i = 1; i < n; i++) {
    W[i][0] = eta * wb[i];
    for (int j = 1; j < n; j++) W[i][j] = eta * (((i == j) ? 1.0 : 0.0) + b * (wb[i] * wb[j]));
  }
}
inline void calc_NT_scalings(const double* s, const double* z, const Cone& K, NTScaling& W) {  // :113-127
  W.K = K; W.ok = true;
  for (int i = 0; i < K.n_ort; i++) W.ort[i] = std::sqrt(s[i] / z[i]);
  soc_NT_scaling(s + K.n_ort, z + K.n_ort, K.d1, W.soc1);
  soc_NT_scaling(s + K.n_ort + K.d1, z + K.n_ort + K.d1, K.d2, W.soc2);
  if (K.d1 > 0) W.ok = W.ok && chol_upper<4>(W.soc1, K.d1, W.soc1_fact);
  if (K.d2 > 0) W.ok = W.ok && chol_upper<4>(W.soc2, K.d2, W.soc2_fact);
}
inline void chol4_solve(const double U[4][4], int n, const double* b, double* x) {
  double y[4];
  for (int i = 0; i < n; i++) { double a = b[i]; for (int k = 0; k < i; k++) a -= U[k][i] * y[k]; y[i] = a / U[i][i]; }
  for (int i = n - 1; i >= 0; i--) { double a = y[i]; for (int k = i + 1; k < n; k++) a -= U[i][k] * x[k]; x[i] = a / U[i][i]; }
}
inline void nt_ldiv(const NTScaling& W, const double* g, double* out) {  // `\` vec :16-29
  const Cone& K = W.K;
  for (int i = 0; i < K.n_ort; i++) out[i] = g[i] / W.ort[i];
  if (K.d1 > 0) chol4_solve(W.soc1_fact, K.d1, g + K.n_ort, out + K.n_ort);
  if (K.d2 > 0) chol4_solve(W.soc2_fact, K.d2, g + K.n_ort + K.d1, out + K.n_ort + K.d1);
}
inline void nt_mul(const NTScaling& W, const double* g, double* out) {  // `*` vec :51-66
  const Cone& K = W.K;
  for (int i = 0; i < K.n_ort; i++) out[i] = g[i] * W.ort[i];
  for (int i = 0; i < K.d1; i++) { double a = 0.0; for (int j = 0; j < K.d1; j++) a += W.soc1[i][j] * g[K.n_ort + j]; out[K.n_ort + i] = a; }
  for (int i = 0; i < K.d2; i++) { double a = 0.0; for (int j = 0; j < K.d2; j++) a += W.soc2[i][j] * g[K.n_ort + K.d1 + j]; out[K.n_ort + K.d1 + i] = a; }
}

// ---------------------------------------------------------------------------
// PDIP solver — src/solvers/coneqp/static_solver2.jl
// ---------------------------------------------------------------------------
inline double ort_linesearch(const double* x, const double* dx, int n) {  // :2-11
  double a = 1.0;
  for (int i = 0; i < n; i++) if (dx[i] < 0) a = std::fmin(a, -x[i] / dx[i]);
  return a;
}
inline double soc_linesearch(const double* y, const double* D, int n) {  // :12-26
  double nu = y[0] * y[0], zeta = y[0] * D[0];
  double yy = 0.0, yd = 0.0;
  for (int i = 1; i < n; i++) { yy += y[i] * y[i]; yd += y[i] * D[i]; }
  nu -= yy; zeta -= yd;
  double sn = std::sqrt(nu);
  double rho0 = zeta / nu;
  double f = ((zeta / sn) + D[0]) / (y[0] / sn + 1);
  double nr = 0.0;
  for (int i = 1; i < n; i++) { double r = D[i] / sn - f * (y[i] / nu); nr += r * r; }
  nr = std::sqrt(nr);
  if (nr > rho0) return std::fmin(1.0, 1 / (nr - rho0));
  return 1.0;
}
inline double linesearch(const double* x, const double* dx, const Cone& K) {  // :27-51
  double a = 1.0;
  if (K.n_ort > 0) a = std::fmin(a, ort_linesearch(x, dx, K.n_ort));
  if (K.d1 > 0) a = std::fmin(a, soc_linesearch(x + K.n_ort, dx + K.n_ort, K.d1));
  if (K.d2 > 0) a = std::fmin(a, soc_linesearch(x + K.n_ort + K.d1, dx + K.n_ort + K.d1, K.d2));
  return a;
}
inline void bring2cone(const double* r, const Cone& K, double* out) {  // :52-85
  double alpha = -1;
  bool any = false; double mn = 0.0;
  for (int i = 0; i < K.n_ort; i++) { if (r[i] <= 0) any = true; if (i == 0 || r[i] < mn) mn = r[i]; }
  if (any) alpha = -mn;
  auto soc = [&](const double* rs, int n) {
    if (n <= 0) return;
    double nr = 0.0;
    for (int i = 1; i < n; i++) nr += rs[i] * rs[i];
    double res = rs[0] - std::sqrt(nr);
    if (res <= 0) alpha = std::fmax(alpha, -res);
  };
  soc(r + K.n_ort, K.d1);
  soc(r + K.n_ort + K.d1, K.d2);
  int ns = K.ns();
  if (alpha < 0) { for (int i = 0; i < ns; i++) out[i] = r[i]; return; }
  double e[MAXS]; gen_e(K, e);
  for (int i = 0; i < ns; i++) out[i] = r[i] + (1 + alpha) * e[i];
}

struct Solution { int nx, ns; double x[MAXX], s[MAXS], z[MAXS]; int iters; int status; };

// G'*W-scaled Gram and products use plain dense loops in the reference's order.
inline bool initialize(const Problem<double>& P, const Cone& K, Solution& S) {  // :86-103
  int nx = P.nx, ns = P.ns;
  double GtG[MAXX][MAXX], U[MAXX][MAXX];
  for (int i = 0; i < nx; i++) for (int j = 0; j < nx; j++) { double a = 0.0; for (int k = 0; k < ns; k++) a += P.G[k][i] * P.G[k][j]; GtG[i][j] = a; }
  if (!chol_upper<MAXX>(GtG, nx, U)) return false;
  double Gth[MAXX], mc[MAXX], x2[MAXX], st[MAXS], zt[MAXS];
  for (int i = 0; i < nx; i++) { double a = 0.0; for (int k = 0; k < ns; k++) a += P.G[k][i] * P.h[k]; Gth[i] = a; mc[i] = -P.c[i]; }
  chol_solve<MAXX>(U, nx, Gth, S.x);
  for (int k = 0; k < ns; k++) { double a = 0.0; for (int j = 0; j < nx; j++) a += P.G[k][j] * S.x[j]; st[k] = a - P.h[k]; }  // s~ = G*x^ - h (:94)
  bring2cone(st, K, S.s);
  chol_solve<MAXX>(U, nx, mc, x2);
  for (int k = 0; k < ns; k++) { double a = 0.0; for (int j = 0; j < nx; j++) a += P.G[k][j] * x2[j]; zt[k] = a; }
  bring2cone(zt, K, S.z);
  return true;
}

// iteration cap: 50 in the reference (:132).  Tests lower it (together with dcol_set_max_iters on the CUDA side) to land a
// pair's convergence exactly on the last allowed step: the reference raises "pdip failed" (:188) without testing mu again.
inline int& max_iters_setting() { static int v = 50; return v; }
inline void solve_socp(const Problem<double>& P, double pdip_tol, Solution& S, int max_iters = max_iters_setting()) {  // :106-191
  Cone K{P.n_ort, P.d1, P.d2};
  int nx = P.nx, ns = P.ns;
  S.nx = nx; S.ns = ns; S.iters = 0; S.status = ST_OK;
  if (!initialize(P, K, S)) { S.status = ST_NUMERICAL; return; }
  double e[MAXS]; gen_e(K, e);
  int cone_degree = K.n_ort + (K.d1 > 0 ? 1 : 0) + (K.d2 > 0 ? 1 : 0);
  NTScaling W;
  double lam[MAXS], ll[MAXS], rx[MAXX], rz[MAXS], bx[MAXX], t1[MAXS], t2[MAXS], lds[MAXS], bz[MAXS];
  double Gt[MAXS][MAXX], M[MAXX][MAXX], U[MAXX][MAXX], rhs[MAXX];
  double dxa[MAXX], dza[MAXS], dsa[MAXS], dx[MAXX], dz[MAXS], ds[MAXS], dsv[MAXS];
  for (int main_iter = 1; main_iter <= max_iters; main_iter++) {
    calc_NT_scalings(S.s, S.z, K, W);                                   // :134
    if (!W.ok) { S.status = ST_NUMERICAL; return; }
    nt_mul(W, S.z, lam);                                                // :137
    cone_product(lam, lam, K, ll);                                      // :138
    for (int i = 0; i < nx; i++) { double a = 0.0; for (int k = 0; k < ns; k++) a += P.G[k][i] * S.z[k]; rx[i] = a + P.c[i]; }   // :141
    double sz = 0.0;
    for (int k = 0; k < ns; k++) { double a = 0.0; for (int j = 0; j < nx; j++) a += P.G[k][j] * S.x[j]; rz[k] = S.s[k] + a - P.h[k]; sz += S.s[k] * S.z[k]; }  // :142
    double mu = sz / cone_degree;                                       // :143
    if (!(mu == mu)) { S.status = ST_NUMERICAL; return; }
    if (mu < pdip_tol) { S.iters = main_iter - 1; return; }             // :144-146
    // affine step :149-156
    for (int i = 0; i < nx; i++) bx[i] = -rx[i];
    for (int k = 0; k < ns; k++) t1[k] = -ll[k];
    inverse_cone_product(lam, t1, K, lds);
    nt_mul(W, lds, t1);
    for (int k = 0; k < ns; k++) t2[k] = -rz[k] - t1[k];
    nt_ldiv(W, t2, bz);
    {  // G~ = W\G column by column (:43-48)
      double col[MAXS], out[MAXS];
      for (int j = 0; j < nx; j++) { for (int k = 0; k < ns; k++) col[k] = P.G[k][j]; nt_ldiv(W, col, out); for (int k = 0; k < ns; k++) Gt[k][j] = out[k]; }
    }
    for (int i = 0; i < nx; i++) for (int j = 0; j < nx; j++) { double a = 0.0; for (int k = 0; k < ns; k++) a += Gt[k][i] * Gt[k][j]; M[i][j] = a; }
    if (!chol_upper<MAXX>(M, nx, U)) { S.status = ST_NUMERICAL; return; }
    for (int i = 0; i < nx; i++) { double a = 0.0; for (int k = 0; k < ns; k++) a += Gt[k][i] * bz[k]; rhs[i] = bx[i] + a; }
    chol_solve<MAXX>(U, nx, rhs, dxa);
    for (int k = 0; k < ns; k++) { double a = 0.0; for (int j = 0; j < nx; j++) a += Gt[k][j] * dxa[j]; t1[k] = a - bz[k]; }
    nt_ldiv(W, t1, dza);
    nt_mul(W, dza, t1);
    for (int k = 0; k < ns; k++) t2[k] = lds[k] - t1[k];
    nt_mul(W, t2, dsa);
    // linesearch on affine step :159-161
    double aa = std::fmin(linesearch(S.s, dsa, K), linesearch(S.z, dza, K));
    double num = 0.0;
    for (int k = 0; k < ns; k++) num += (S.s[k] + aa * dsa[k]) * (S.z[k] + aa * dza[k]);
    double rho = num / sz;
    double sig = std::fmax(0.0, std::fmin(1.0, rho)); sig = sig * sig * sig;
    // centering + correcting step :164-169
    nt_ldiv(W, dsa, t1);
    nt_mul(W, dza, t2);
    cone_product(t1, t2, K, dsv);
    for (int k = 0; k < ns; k++) dsv[k] = -ll[k] - dsv[k] + sig * mu * e[k];
    inverse_cone_product(lam, dsv, K, lds);
    nt_mul(W, lds, t1);
    for (int k = 0; k < ns; k++) t2[k] = -rz[k] - t1[k];
    nt_ldiv(W, t2, bz);
    for (int i = 0; i < nx; i++) { double a = 0.0; for (int k = 0; k < ns; k++) a += Gt[k][i] * bz[k]; rhs[i] = bx[i] + a; }
    chol_solve<MAXX>(U, nx, rhs, dx);
    for (int k = 0; k < ns; k++) { double a = 0.0; for (int j = 0; j < nx; j++) a += Gt[k][j] * dx[j]; t1[k] = a - bz[k]; }
    nt_ldiv(W, t1, dz);
    nt_mul(W, dz, t1);
    for (int k = 0; k < ns; k++) t2[k] = lds[k] - t1[k];
    nt_mul(W, t2, ds);
    // final line search :172, step :175-177
    double al = std::fmin(1.0, 0.99 * std::fmin(linesearch(S.s, ds, K), linesearch(S.z, dz, K)));
    for (int i = 0; i < nx; i++) S.x[i] += al * dx[i];
    for (int k = 0; k < ns; k++) { S.s[k] += al * ds[k]; S.z[k] += al * dz[k]; }
    S.iters = main_iter;
  }
  S.status = ST_PDIP_FAILED;  // error("pdip failed") :188
}

// ---------------------------------------------------------------------------
// query API — src/proximity.jl, src/proximity_gradient.jl
// ---------------------------------------------------------------------------
struct Pose { const double* r; const double* att; };

inline void build_problem(const Shape& s1, const Shape& s2, const Pose& p1, const Pose& p2, int att_kind, Problem<double>& P) {
  static thread_local Block<double> B1, B2;
  problem_matrices<double>(s1, p1.r, p1.att, att_kind, B1);
  problem_matrices<double>(s2, p2.r, p2.att, att_kind, B2);
  combine_problem_matrices(B1, B2, P);
}

// proximity — src/proximity.jl:3-26
inline int proximity(const Shape& s1, const Shape& s2, const Pose& p1, const Pose& p2, int att_kind, double tol,
                     double* alpha, double* x, int* iters, Solution* keep = nullptr, Problem<double>* keepP = nullptr) {
  static thread_local Problem<double> Pl; static thread_local Solution Sl;
  Problem<double>& P = keepP ? *keepP : Pl; Solution& S = keep ? *keep : Sl;
  build_problem(s1, s2, p1, p2, att_kind, P);
  solve_socp(P, tol, S);
  *alpha = S.x[3]; x[0] = S.x[0]; x[1] = S.x[1]; x[2] = S.x[2];
  if (iters) *iters = S.iters;
  return S.status;
}

// dual evaluation of the pair's (G(theta), h(theta)) with NT partials seeded on
// theta = [r1; att1; r2; att2] (src/proximity.jl:88, src/proximity_gradient.jl:58)
template <int NT>
inline void build_problem_dual(const Shape& s1, const Shape& s2, const Pose& p1, const Pose& p2, int att_kind, Problem<Dual<NT>>& P) {
  typedef Dual<NT> D;
  int na = (att_kind == ATT_QUAT) ? 4 : 3;
  D r1[3], a1[4], r2[3], a2[4];
  int k = 0;
  for (int i = 0; i < 3; i++) { r1[i] = D(p1.r[i]); r1[i].d[k++] = 1.0; }
  for (int i = 0; i < na; i++) { a1[i] = D(p1.att[i]); a1[i].d[k++] = 1.0; }
  for (int i = 0; i < 3; i++) { r2[i] = D(p2.r[i]); r2[i].d[k++] = 1.0; }
  for (int i = 0; i < na; i++) { a2[i] = D(p2.att[i]); a2[i].d[k++] = 1.0; }
  static thread_local Block<D> B1, B2;
  problem_matrices<D>(s1, r1, a1, att_kind, B1);
  problem_matrices<D>(s2, r2, a2, att_kind, B2);
  combine_problem_matrices(B1, B2, P);
}

// proximity_gradient — src/proximity_gradient.jl:80-111 ; obj_val_grad :41-78 ; lag_con_part :1-40
template <int NT>
inline int proximity_gradient_t(const Shape& s1, const Shape& s2, const Pose& p1, const Pose& p2, int att_kind, double tol,
                                double* alpha, double* grad, int* iters) {
  static thread_local Problem<double> P; static thread_local Solution S; static thread_local Problem<Dual<NT>> PD;
  double xw[3];
  int st = proximity(s1, s2, p1, p2, att_kind, tol, alpha, xw, iters, &S, &P);
  if (st != ST_OK) { for (int j = 0; j < NT; j++) grad[j] = NAN; return st; }
  build_problem_dual<NT>(s1, s2, p1, p2, att_kind, PD);
  Dual<NT> acc(0.0);                                          // z'*(G*x - h)  (:19)
  for (int k = 0; k < PD.ns; k++) {
    Dual<NT> gx(0.0);
    for (int j = 0; j < PD.nx; j++) gx = gx + PD.G[k][j] * S.x[j];
    acc = acc + S.z[k] * (gx - PD.h[k]);
  }
  for (int j = 0; j < NT; j++) grad[j] = acc.d[j];
  return st;
}

// proximity_jacobian — src/proximity.jl:168-195 ; diff_socp :71-134 ; kkt_R :27-70
// J is 4 x NT column-major (Julia SMatrix layout), rows (x1,x2,x3,alpha).
template <int NT>
inline int proximity_jacobian_t(const Shape& s1, const Shape& s2, const Pose& p1, const Pose& p2, int att_kind, double tol,
                                double* alpha, double* x, double* J, int* iters, bool slow = false) {
  typedef Dual<NT> D;
  static thread_local Problem<double> P; static thread_local Solution S; static thread_local Problem<D> PD;
  int st = proximity(s1, s2, p1, p2, att_kind, tol, alpha, x, iters, &S, &P);
  if (st != ST_OK) { for (int j = 0; j < 4 * NT; j++) J[j] = NAN; return st; }
  int nx = P.nx, ns = P.ns;
  Cone K{P.n_ort, P.d1, P.d2};
  build_problem_dual<NT>(s1, s2, p1, p2, att_kind, PD);
  // kkt_R = [c + G'z ; cone_product(h - G x, z)]
  static thread_local D Rx[MAXX], u[MAXS], zd[MAXS], Rz[MAXS];
  for (int i = 0; i < nx; i++) { D a(P.c[i]); for (int k = 0; k < ns; k++) a = a + PD.G[k][i] * S.z[k]; Rx[i] = a; }
  for (int k = 0; k < ns; k++) { D a = PD.h[k]; for (int j = 0; j < nx; j++) a = a - PD.G[k][j] * S.x[j]; u[k] = a; zd[k] = D(S.z[k]); }
  cone_product<D>(u, zd, K, Rz);
  if (!slow) {
    // Z = scaling_2(z_ort, arrow(z_soc1), arrow(z_soc2)) ; S = NT_scaling_2(s_ort, arrow(s_soc), chol(arrow(s_soc))) :98-99
    // dx = (G'*((S\Z)*G)) \ (r1 - G'*(S\r2))  :100
    double SZG[MAXS][MAXX], Sr2[MAXS][MAXT];
    for (int k = 0; k < K.n_ort; k++) {
      for (int j = 0; j < nx; j++) SZG[k][j] = (S.z[k] / S.s[k]) * P.G[k][j];
      for (int t = 0; t < NT; t++) Sr2[k][t] = (-Rz[k].d[t]) / S.s[k];
    }
    auto soc_part = [&](int off, int d) -> bool {
      if (d <= 0) return true;
      double As[4][4], Az[4][4], Us[4][4], SZ[4][4];
      arrow(S.s + off, d, As); arrow(S.z + off, d, Az);
      if (!chol_upper<4>(As, d, Us)) return false;
      for (int j = 0; j < d; j++) { double col[4], out[4]; for (int i = 0; i < d; i++) col[i] = Az[i][j]; chol4_solve(Us, d, col, out); for (int i = 0; i < d; i++) SZ[i][j] = out[i]; }
      for (int i = 0; i < d; i++) for (int j = 0; j < nx; j++) { double a = 0.0; for (int k = 0; k < d; k++) a += SZ[i][k] * P.G[off + k][j]; SZG[off + i][j] = a; }
      for (int t = 0; t < NT; t++) { double col[4], out[4]; for (int i = 0; i < d; i++) col[i] = -Rz[off + i].d[t]; chol4_solve(Us, d, col, out); for (int i = 0; i < d; i++) Sr2[off + i][t] = out[i]; }
      return true;
    };
    if (!soc_part(K.n_ort, K.d1) || !soc_part(K.n_ort + K.d1, K.d2)) { for (int j = 0; j < 4 * NT; j++) J[j] = NAN; return ST_NUMERICAL; }
    double A[MAXX][MAXX], B[MAXX][MAXT];
    for (int i = 0; i < nx; i++) {
      for (int j = 0; j < nx; j++) { double a = 0.0; for (int k = 0; k < ns; k++) a += P.G[k][i] * SZG[k][j]; A[i][j] = a; }
      for (int t = 0; t < NT; t++) { double a = 0.0; for (int k = 0; k < ns; k++) a += P.G[k][i] * Sr2[k][t]; B[i][t] = (-Rx[i].d[t]) - a; }
    }
    if (!lu_solve(A, nx, B, NT)) { for (int j = 0; j < 4 * NT; j++) J[j] = NAN; return ST_NUMERICAL; }
    for (int t = 0; t < NT; t++) for (int i = 0; i < 4; i++) J[t * 4 + i] = B[i][t];
    return st;
  }
  // diff_socp_slow :135-167 — dense (nx+ns) KKT with s~ = h - G x ; dw = -dR_dw \ dR_dtheta
  {
    constexpr int MK = MAXX + MAXS;
    static thread_local double Kk[MK][MK], Rhs[MK][MAXT];
    int n = nx + ns;
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Kk[i][j] = 0.0;
    for (int i = 0; i < nx; i++) for (int k = 0; k < ns; k++) Kk[i][nx + k] = P.G[k][i];
    double st_[MAXS];
    for (int k = 0; k < ns; k++) { double a = P.h[k]; for (int j = 0; j < nx; j++) a -= P.G[k][j] * S.x[j]; st_[k] = a; }
    for (int k = 0; k < K.n_ort; k++) { for (int j = 0; j < nx; j++) Kk[nx + k][j] = -S.z[k] * P.G[k][j]; Kk[nx + k][nx + k] = st_[k]; }
    auto soc_part = [&](int off, int d) {
      if (d <= 0) return;
      double As[4][4], Az[4][4];
      arrow(st_ + off, d, As); arrow(S.z + off, d, Az);
      for (int i = 0; i < d; i++) {
        for (int j = 0; j < nx; j++) { double a = 0.0; for (int k = 0; k < d; k++) a += Az[i][k] * P.G[off + k][j]; Kk[nx + off + i][j] = -a; }
        for (int j = 0; j < d; j++) Kk[nx + off + i][nx + off + j] = As[i][j];
      }
    };
    soc_part(K.n_ort, K.d1); soc_part(K.n_ort + K.d1, K.d2);
    for (int i = 0; i < nx; i++) for (int t = 0; t < NT; t++) Rhs[i][t] = -Rx[i].d[t];
    for (int k = 0; k < ns; k++) for (int t = 0; t < NT; t++) Rhs[nx + k][t] = -Rz[k].d[t];
    // LU with partial pivoting on the dense KKT
    for (int k = 0; k < n; k++) {
      int p = k; double best = std::fabs(Kk[k][k]);
      for (int i = k + 1; i < n; i++) if (std::fabs(Kk[i][k]) > best) { best = std::fabs(Kk[i][k]); p = i; }
      if (best == 0.0) { for (int j = 0; j < 4 * NT; j++) J[j] = NAN; return ST_NUMERICAL; }
      if (p != k) { for (int j = 0; j < n; j++) { double t = Kk[k][j]; Kk[k][j] = Kk[p][j]; Kk[p][j] = t; } for (int j = 0; j < NT; j++) { double t = Rhs[k][j]; Rhs[k][j] = Rhs[p][j]; Rhs[p][j] = t; } }
      for (int i = k + 1; i < n; i++) { double f = Kk[i][k] / Kk[k][k]; if (f == 0.0) continue; for (int j = k + 1; j < n; j++) Kk[i][j] -= f * Kk[k][j]; for (int j = 0; j < NT; j++) Rhs[i][j] -= f * Rhs[k][j]; }
    }
    for (int i = n - 1; i >= 0; i--) for (int t = 0; t < NT; t++) { double a = Rhs[i][t]; for (int k = i + 1; k < n; k++) a -= Kk[i][k] * Rhs[k][t]; Rhs[i][t] = a / Kk[i][i]; }
    for (int t = 0; t < NT; t++) for (int i = 0; i < 4; i++) J[t * 4 + i] = Rhs[i][t];
    return st;
  }
}

}  // namespace dcol_oracle
