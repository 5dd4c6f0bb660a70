// Shared definitions for the dcol_b200 CUDA library (sm_100a, FP64).
//
// Formulation: every primitive's constraint block is
//     u_k(y) = C_k * [x~_k ; alpha ; e_k],   x~_k = Q_k^T (x - r~_k)   (body frame)
// with C_k a pose-independent constant matrix (rows c_i, small SOC block), so the
// reference's  h - G*y  (src/problem_matrices.jl:30-191, src/combine_problem_matrices.jl:15-86)
// equals u(y) with  G = -C*T, h = C*t^.  The solver works in primitive-1's body
// frame (xi = x~_1), where only the relative pose R = Q_2^T Q_1, t = Q_2^T (r~_1 - r~_2)
// enters:  x~_2 = R*xi + t.  The PDIP iterates (s, z) are invariant under this
// affine change of variables, so results match the reference up to round-off.
#pragma once
#include <cstdint>

namespace dcol {

// reference shape types (include/dcol_b200.h)
enum { T_POLYTOPE = 0, T_CAPSULE = 1, T_CYLINDER = 2, T_CONE = 3, T_SPHERE = 4, T_ELLIPSOID = 5, T_POLYGON = 6 };
// structural classes the kernels are specialised on
enum { CLS_POLYTOPE = 0, CLS_SPHERE = 1, CLS_CAPCYL = 2, CLS_POLYGON = 3, CLS_CONE = 4, CLS_ELL = 5, NCLS = 6 };
enum { SOC_NONE = 0, SOC_BALL = 1, SOC_CONE = 2, SOC_ELL = 3 };
enum { ST_OK = 0, ST_PDIP_FAILED = 1, ST_NUMERICAL = 2 };
enum { Q_PROX = 0, Q_GRAD = 1, Q_JAC = 2 };

// nz(a): can coefficient a of an orthant row c = [c_x(3), c_alpha, c_e(NE)] be non-zero for this class?  The rows of
// capsules / cylinders (axis = body x), cones (base plane normal = body x) and polygons (planar constraints on the
// slack e only) are structurally sparse; the row passes skip the zero columns at compile time.
template <int CLS> struct Cls;
template <> struct Cls<CLS_POLYTOPE> { static constexpr int NE = 0, SOC = SOC_NONE, D = 0;
  __host__ __device__ static constexpr bool nz(int) { return true; } };
template <> struct Cls<CLS_SPHERE>   { static constexpr int NE = 0, SOC = SOC_BALL, D = 4;
  __host__ __device__ static constexpr bool nz(int) { return true; } };
template <> struct Cls<CLS_CAPCYL>   { static constexpr int NE = 1, SOC = SOC_BALL, D = 4;
  __host__ __device__ static constexpr bool nz(int a) { return a == 0 || a >= 3; } };
template <> struct Cls<CLS_POLYGON>  { static constexpr int NE = 2, SOC = SOC_BALL, D = 4;
  __host__ __device__ static constexpr bool nz(int a) { return a >= 3; } };
template <> struct Cls<CLS_CONE>     { static constexpr int NE = 0, SOC = SOC_CONE, D = 3;
  __host__ __device__ static constexpr bool nz(int a) { return a == 0 || a == 3; } };
template <> struct Cls<CLS_ELL>      { static constexpr int NE = 0, SOC = SOC_ELL, D = 4;
  __host__ __device__ static constexpr bool nz(int) { return true; } };

__host__ __device__ inline int cls_ne(int c) { return c == CLS_CAPCYL ? 1 : (c == CLS_POLYGON ? 2 : 0); }
__host__ __device__ inline int cls_d(int c) { return c == CLS_POLYTOPE ? 0 : (c == CLS_CONE ? 3 : 4); }

// Device shape record.  rows: n_ort x (4+NE) doubles, signs baked so that
// u_i = c_i . [x~ ; alpha ; e].  soc[]: BALL {R}; CONE {tan(beta), tan(beta)*3H/4}; ELL {U row-major 3x3}.
struct DevShape {
  int32_t cls, n_ort, type, has_offset;
  int64_t rows_off;
  double soc[9];
  double r_off[3];
  double Q_off[9];
};

constexpr int MAX_ITERS = 50;   // static_solver2.jl:132
constexpr int NFIELD = 5;       // per-row shared-memory fields
constexpr int POSE_SLOTS = 24;  // per-thread shared-memory slots ahead of the rows: relative pose R (9) | t (3) | primitive 1's frame Q (9) | r~ (3)
enum { F_S = 0, F_Z = 1, F_R = 2, F_RZ = 3, F_P = 4 };   // s, z, 1/(s z), rz | ds, p | dz

// One pair, packed by the binning pass at its SORTED position (dcol_capi.cu, k_scatter_pack): 128 bytes = one cache
// line per pair, so the solver's loads touch every fetched byte and the binning pass writes whole lines.  A pair whose
// classes arrive as (hi, lo) is stored with its two primitives exchanged (DCOL_REC_SWAPPED): the solver always sees
// class(prim 1) <= class(prim 2) and only the derivative epilogue looks at the flag (column placement).
struct alignas(16) PairRec {
  double r1[3], a1[4];             // pose of the record's first primitive (a1[3] unused for MRPs)
  double r2[3], a2[4];
  int32_t s1, s2;                  // shape-table rows
  int32_t pair;                    // the caller's pair index
  int32_t flags;
};
static_assert(sizeof(PairRec) == 128, "one cache line per pair");
enum { DCOL_REC_SWAPPED = 1 };

struct QueryParams {
  const DevShape* shapes;
  const double* rows;            // packed row constants
  const PairRec* recs;           // packed pairs in sorted order (first pass: tile position; resume: position kept in the hand-over record)
  int att_kind;                  // 0 quaternion (stride 4), 1 MRP (stride 3)
  int row_pitch;                 // doubles of shared-memory row store per thread (odd)
  int max_iters;                 // 50 (static_solver2.jl:132); smaller only through the testing knob dcol_set_max_iters
  int direct;                    // 1: alpha/x/status/iters/mu are indexed by the caller's pair index (small batches: no un-permute pass),
                                 // 0: by sorted position (coalesced stores; k_unpermute brings them into the caller's order)
  double pdip_tol;
  double* alpha; double* x; int32_t* status; int32_t* iters;
  double* mu;                    // optional: duality measure s'z/deg at the iteration that ended the solve
  double* deriv;                 // gradient n x ntheta, or J n x (4 x ntheta) column-major; always at the caller's pair index
  double* warm;                  // optional warm-start records, warm_stride doubles per pair (dcol_set_warm_start); NULL = cold starts
  int warm_stride, warm_cap;     // record size in doubles; number of pairs the buffer holds
  double warm_mu;                // an iterate is saved at the first iteration with mu below this
  int query;
};


// Work description of one persistent launch (lives in global memory, written by the host per call).
// Tiles are CTA-sized (consecutive sorted positions of one class-pair bin); tiles of a bin are
// contiguous in the queue and bins are queued heaviest first, so at any moment the CTAs in flight
// execute (nearly) the same specialisation — instruction-cache friendly — and only the end of the
// queue has a tail.
constexpr int MAX_BINS = NCLS * NCLS;
struct LaunchTable {
  QueryParams P;
  int n_bins, n_tiles, rows_cap, mode;     // mode 0: first pass (may hand stragglers over), 1: resume pass
  int tile;                        // CTA size (= pairs per tile) of the tier
  int tile_prefix[MAX_BINS + 1];   // first tile of bin b
  int begin[MAX_BINS];             // sorted-position range of bin b (resume pass: 0 .. ov_count)
  int end[MAX_BINS];
  int cp[MAX_BINS];                // class pair c1*NCLS + c2, c1 <= c2
  // straggler hand-over: when at most spill_t pairs of a tile are still iterating, they park their iterate in the
  // bin's overflow list and the tile ends; a second (resume) launch finishes them in dense tiles
  int spill_t;
  double* ov_base;                 // records this pass WRITES (hand-over target)
  long long ov_off[MAX_BINS];      // start of bin b's overflow records (doubles), SoA with stride ov_cap[b]
  int ov_cap[MAX_BINS];            // record capacity of bin b
  int ov_count[MAX_BINS];          // records written (atomics)
  int ov_stride[MAX_BINS];         // doubles per record: 3 + nx + 2*rows + 2*(d1+d2)
  double* in_base;                 // records this pass READS (resume passes)
  long long in_off[MAX_BINS];
  int in_cap[MAX_BINS];
};

// A handful of pairs whose records the HOST packed (host-pointer calls of up to DCOL_PACKED_MAX pairs): one block per
// class pair present, records of a class pair contiguous.  Passed to the kernel by value.
constexpr int DCOL_PACKED_MAX = 128;
struct PackedGroups { int n; int cp[MAX_BINS]; int first[MAX_BINS]; int count[MAX_BINS]; };

// what one tile needs to know beyond the query parameters
struct TileCtx {
  int mode;                        // 0 first pass, 1 resume
  int spill_t;                     // 0: never hand over
  double* ov;                      // this bin's hand-over records (written)
  int ov_cap;
  int* ov_count;
  const double* in;                // this bin's records to resume from (mode 1)
  int in_cap;
  int dup;                         // lane past the end of the bin: redoes the bin's last pair for the barriers only
};

}  // namespace dcol
