/* dcol_b200 — C ABI of the B200-native batched DCOL proximity library.
 *
 * The reference (kevin-tracy/DifferentiableCollisions.jl v0.1.5) is pure Julia and has
 * no FFI boundary; this header is the boundary a Julia `ccall` (or Python ctypes)
 * wrapper binds to replace the reference's call surface:
 *
 *   dcol_proximity           <-> dc.proximity            src/proximity.jl:3-14   (quat), :15-26   (MRP)
 *   dcol_proximity_gradient  <-> dc.proximity_gradient   src/proximity_gradient.jl:80-95 (quat), :96-111 (MRP)
 *   dcol_proximity_jacobian  <-> dc.proximity_jacobian   src/proximity.jl:168-181 (quat), :182-195 (MRP)
 *   dcol_set_shapes          <-> the geometry fields of the 14 primitive structs
 *                                src/primitives.jl:4-96, src/primitives_mrp.jl:2-96
 *
 * Each call is the batched form: arrays of n pairs.  The reference's single-pair call
 * is a batch of 1.  All arithmetic is FP64 on the GPU; there is no CPU fallback.
 *
 * Layouts (all caller-allocated, contiguous):
 *   shape1, shape2 : int32[n]   indices into the shape table
 *   r1, r2         : f64[n*3]   positions  (prim.r)
 *   att1, att2     : f64[n*4]   quaternions, scalar first, need not be normalised (att_kind 0; prim.q)
 *                    f64[n*3]   modified Rodrigues parameters                       (att_kind 1; prim.p)
 *   alpha          : f64[n]     minimum uniform scaling; collision <=> alpha <= 1
 *   x              : f64[n*3]   intersection / witness point in the world frame
 *   grad           : f64[n*nt]  d alpha / d[r1; att1; r2; att2],  nt = 14 (quat) or 12 (MRP)
 *   J              : f64[n*4*nt] per pair a 4 x nt matrix, COLUMN-major (Julia SMatrix layout),
 *                    rows (x1, x2, x3, alpha), columns [r1; att1; r2; att2]  (src/proximity.jl:83-86,101)
 *   status         : int32[n]   0 ok, 1 "pdip failed" (no convergence in 50 iterations,
 *                    static_solver2.jl:188), 2 numerical breakdown (non-PD / NaN)      [may be NULL]
 *   iters          : int32[n]   interior-point steps taken                               [may be NULL]
 *
 * mem: DCOL_MEM_HOST   — all pointers are host memory; the library stages H2D/D2H itself.
 *      DCOL_MEM_DEVICE — all pointers are device memory on the context's GPU; nothing is copied.
 *
 * Return value: 0 on success, <0 on API/CUDA error (see dcol_last_error).  Per-pair solver
 * failures are reported through `status`, never through the return value.  No exceptions cross
 * the ABI.  A context is NOT thread-safe; use one context per host thread / per GPU.
 */
#ifndef DCOL_B200_H
#define DCOL_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCOL_ATT_QUAT 0
#define DCOL_ATT_MRP  1
#define DCOL_MEM_HOST   0
#define DCOL_MEM_DEVICE 1

/* shape types (reference struct names) */
#define DCOL_POLYTOPE  0
#define DCOL_CAPSULE   1
#define DCOL_CYLINDER  2
#define DCOL_CONE      3
#define DCOL_SPHERE    4
#define DCOL_ELLIPSOID 5
#define DCOL_POLYGON   6

#define DCOL_STATUS_OK          0
#define DCOL_STATUS_PDIP_FAILED 1
#define DCOL_STATUS_NUMERICAL   2
#define DCOL_STATUS_CULLED      3   /* scene operators with dcol_scene_set_cull only: far pair, alpha is a lower bound */

typedef struct dcol_ctx dcol_ctx;

/* Create / destroy a context bound to one CUDA device.  Owns streams, the device copy of the
 * shape table and scratch buffers. */
int  dcol_create(dcol_ctx** out, int device_id);
void dcol_destroy(dcol_ctx* ctx);
/* Context-owned, NUL-terminated description of the last error ("" if none). Valid for ctx == NULL
 * for errors raised by dcol_create. */
const char* dcol_last_error(const dcol_ctx* ctx);
const char* dcol_version(void);

/* Optional: run on a caller-provided CUDA stream (cudaStream_t passed as void*); NULL = the
 * context's own stream.  Device-pointer calls are then ordered after work already enqueued on
 * that stream and complete before the call returns. */
int dcol_set_stream(dcol_ctx* ctx, void* cuda_stream);

/* Immutable shape table, copied to the device.  params (f64) packed at param_off[i]:
 *   polytope  {n, A row-major n x 3, b[n]}     A x~ <= alpha b          src/primitives.jl:4-14
 *   capsule   {R, L}                                                    :16-26
 *   cylinder  {R, L}                                                    :28-38
 *   cone      {H, beta}                                                 :40-50
 *   sphere    {R}                                                       :52-61
 *   ellipsoid {U row-major 3 x 3} with U = cholesky(P).U                :76-96
 *   polygon   {nh, A row-major nh x 2, b[nh], R}                        :63-74
 * followed, for every shape, by r_offset[3] and Q_offset[9] (row-major). */
int dcol_set_shapes(dcol_ctx* ctx, int n_shapes, const int32_t* type, const int64_t* param_off,
                    const double* params, int64_t n_params);

int dcol_proximity(dcol_ctx* ctx, int64_t n, const int32_t* shape1, const int32_t* shape2,
                   const double* r1, const double* att1, const double* r2, const double* att2,
                   int att_kind, double pdip_tol, int mem,
                   double* alpha, double* x, int32_t* status, int32_t* iters);

int dcol_proximity_gradient(dcol_ctx* ctx, int64_t n, const int32_t* shape1, const int32_t* shape2,
                            const double* r1, const double* att1, const double* r2, const double* att2,
                            int att_kind, double pdip_tol, int mem,
                            double* alpha, double* grad, int32_t* status, int32_t* iters);

int dcol_proximity_jacobian(dcol_ctx* ctx, int64_t n, const int32_t* shape1, const int32_t* shape2,
                            const double* r1, const double* att1, const double* r2, const double* att2,
                            int att_kind, double pdip_tol, int mem,
                            double* alpha, double* x, double* J, int32_t* status, int32_t* iters);

/* ---- next-tier operators: the callers' batch shapes, same kernels -------------------------
 *
 * Indexed query: r1/att1 and r2/att2 are pose TABLES of n_pose1 / n_pose2 rows; pair k uses rows pose1[k] and
 * pose2[k].  query: 0 proximity (deriv unused), 1 gradient (x unused, deriv = grad), 2 jacobian (deriv = J).
 * This is the call pattern of the trajectory-optimisation examples — one vehicle pose per (rollout, knot)
 * against K static obstacles (examples/trajectory_optimization/piano_mover.jl:39-53, simple_altro.jl:79-83) —
 * without replicating poses per pair. */
int dcol_query_indexed(dcol_ctx* ctx, int query, int64_t n, const int32_t* shape1, const int32_t* shape2,
                       const int32_t* pose1, const int32_t* pose2,
                       int64_t n_pose1, const double* r1, const double* att1,
                       int64_t n_pose2, const double* r2, const double* att2,
                       int att_kind, double pdip_tol, int mem,
                       double* alpha, double* x, double* deriv, int32_t* status, int32_t* iters);

/* All-pairs contact scene: every pair (i, j), i < j, of n_bodies bodies, in combinations(1:N, 2) order
 * (examples/contact_physics/ncp_contact_simulation.jl:48, :86-89, :133-139).  The pair list is generated on the
 * device; outputs have dcol_scene_pair_count(n_bodies) = N(N-1)/2 rows, derivative columns are
 * [r_i; att_i; r_j; att_j].  shape[n_bodies], r[n_bodies*3], att[n_bodies*(4|3)]. */
int64_t dcol_scene_pair_count(int n_bodies);
int dcol_scene_allpairs(dcol_ctx* ctx, int query, int n_bodies, const int32_t* shape, const double* r, const double* att,
                        int att_kind, double pdip_tol, int mem,
                        double* alpha, double* x, double* deriv, int32_t* status, int32_t* iters);

/* A slice [k_begin, k_end) of the same enumeration (the all-pairs scene shards by giving every GPU
 * the body arrays and one contiguous slice of the pair indices).  Outputs have k_end - k_begin rows; row k belongs to
 * pair k_begin + k.  Results are bit-identical to the corresponding rows of the whole call. */
int dcol_scene_allpairs_range(dcol_ctx* ctx, int query, int n_bodies, int64_t k_begin, int64_t k_end,
                              const int32_t* shape, const double* r, const double* att,
                              int att_kind, double pdip_tol, int mem,
                              double* alpha, double* x, double* deriv, int32_t* status, int32_t* iters);

/* Contact-scene accumulation: the loop of examples/contact_physics/ncp_contact_simulation.jl:133-139 as one call.  For
 * all pairs k = (i, j), i < j (order as above) with weights lambda[k] (the contact impulses of the caller's NCP):
 *     alpha[k]  = proximity_gradient(P_i, P_j)[1]
 *     wrench[b] = sum over pairs k containing body b of lambda[k] * [ d alpha_k / d r_b ; P(att_b)' d alpha_k / d att_b ]
 * with P = G(q) = L(q) H, the 4 x 3 attitude Jacobian of examples/contact_physics/variational_utils.jl:24-26, for
 * quaternions (att_kind 0) and the identity for MRPs (att_kind 1): 6 numbers per body, the generalised force the
 * example adds to each body's discrete Euler-Lagrange residual (before its constant factors h and tau_mod).  The
 * n_pairs x 14 gradient rows never leave the device (config 4: 0.94 GB that the per-pair form returns to the host).
 * lambda f64[n_pairs]; wrench f64[n_bodies*6]; alpha, status, iters [n_pairs], each may be NULL.  Pairs that fail
 * (status != 0) contribute NaN, as the reference would throw. */
int dcol_scene_wrench(dcol_ctx* ctx, int n_bodies, const int32_t* shape, const double* r, const double* att, const double* lambda,
                      int att_kind, double pdip_tol, int mem,
                      double* alpha, double* wrench, int32_t* status, int32_t* iters);
/* The part of that sum that comes from the pairs k_begin <= k < k_end: lambda, alpha, status, iters have k_end - k_begin
 * rows, wrench is the full n_bodies*6 partial sum.  The partial sums of a partition of the pairs add up to the whole
 * (the one exchange step of a sharded scene: dcol_group_scene_wrench adds them on the host, ranks under torchrun
 * all-reduce them over NCCL -- dcol_b200.sharding.scene_wrench_sharded). */
int dcol_scene_wrench_range(dcol_ctx* ctx, int n_bodies, int64_t k_begin, int64_t k_end,
                            const int32_t* shape, const double* r, const double* att, const double* lambda,
                            int att_kind, double pdip_tol, int mem,
                            double* alpha, double* wrench, int32_t* status, int32_t* iters);

/* Broad phase of the two scene operators above (opt-in: the reference's contact example,
 * examples/contact_physics/ncp_contact_simulation.jl:86-89,133-139, solves every pair).  Every primitive scales about its
 * body origin r~ = r + Qn r_offset and lies inside the ball of radius rho around it, so alpha* >= |r~_i - r~_j| / (rho_i + rho_j).
 * With alpha_cut > 0 the pairs whose bound exceeds alpha_cut are answered without the conic solve: alpha = that lower
 * bound (> alpha_cut), status = DCOL_STATUS_CULLED, iters = 0, zero derivative rows, NaN witness point; in
 * dcol_scene_wrench they contribute nothing.  All other pairs are solved exactly as without the cull (bit-identical).
 * alpha_cut <= 0 switches the cull off (the default). */
int dcol_scene_set_cull(dcol_ctx* ctx, double alpha_cut);

/* Planar vehicle against static obstacles: the constraint functions of the trajectory-optimisation examples
 * (examples/trajectory_optimization/piano_mover.jl:39-53, ineq_con_x / ineq_con_x_jac) for a batch of states.  Every state
 * s = (x, y, theta) puts the vehicle shape at r = (x, y, 0) with MRP p = (0, 0, tan(theta/4)); for every obstacle o
 * (shape ids obs_shape[n_obs], poses r_obs / p_obs [n_obs*3], MRP):
 *     c [s*n_obs + o]        = 1 - alpha(vehicle(s), obstacle o)
 *     Jc[(s*n_obs + o)*3 + m] = d c / d (x, y, theta)_m
 *                             = (-d alpha/d r1_x, -d alpha/d r1_y, -(d alpha/d p1_z) / (4 cos^2(theta/4)))
 * i.e. the non-zero columns 1, 2 and 5 of the example's constraint Jacobian rows.  The state -> pose map, the pair list
 * and the chain rule run on the device around one proximity_gradient query.  xytheta f64[n_states*3]; c f64[n_states*n_obs];
 * Jc f64[n_states*n_obs*3] (may be NULL); status int32[n_states*n_obs] (may be NULL; failed pairs give NaN in Jc). */
int dcol_se2_constraints(dcol_ctx* ctx, int64_t n_states, const double* xytheta, int32_t vehicle_shape,
                         int n_obs, const int32_t* obs_shape, const double* r_obs, const double* p_obs,
                         double pdip_tol, int mem, double* c, double* Jc, int32_t* status);

/* Rigid vehicle against static obstacles: the same constraint functions in the 3-D trajectory examples
 * (examples/trajectory_optimization/cone_through_wall.jl:56-68, cluttered_hallway_quadrotor.jl:83-104, ineq_con_x /
 * ineq_con_x_jac).  Row s of `states` (state_stride doubles per row) holds the vehicle's position in columns
 * r_col .. r_col+2 and its MRP in columns p_col .. p_col+2 (the examples' state [r; v; p; omega]: stride 12, r_col 0,
 * p_col 6); for every obstacle o (shape ids, poses r_obs / p_obs, MRP):
 *     c [s*n_obs + o]          = 1 - alpha(vehicle(s), obstacle o)
 *     Jc[(s*n_obs + o)*6 + m]  = -(d alpha / d r1)_m  for m < 3,   -(d alpha / d p1)_{m-3}  for m >= 3
 * i.e. the non-zero blocks (columns r_col.., p_col..) of the examples' constraint Jacobian rows
 * [-J[1:3]' 0 -J[4:6]' 0].  One proximity_gradient query on the device between a gather and a row kernel.
 * states f64[n_states*state_stride]; c f64[n_states*n_obs]; Jc f64[n_states*n_obs*6] (may be NULL); status as above. */
int dcol_se3_constraints(dcol_ctx* ctx, int64_t n_states, const double* states, int state_stride, int r_col, int p_col,
                         int32_t vehicle_shape, int n_obs, const int32_t* obs_shape, const double* r_obs, const double* p_obs,
                         double pdip_tol, int mem, double* c, double* Jc, int32_t* status);

/* ---- asynchronous submissions (host buffers) -------------------------------------------------------------
 * dcol_submit enqueues one query (query: 0 proximity, 1 gradient, 2 jacobian; buffers as in the synchronous
 * calls, all HOST memory, pinned for the copies to overlap) and returns a ticket >= 0 without waiting; dcol_wait
 * blocks until that submission's results are in the caller's buffers and returns its status (0 / <0).  Three
 * submissions may be in flight: the inputs of submission k+1 cross PCIe while submission k computes and the
 * results of submission k-1 go back.  Submitting a fourth waits for the oldest.  The caller's input and output
 * buffers must stay valid and untouched until the matching dcol_wait returns. */
int dcol_submit(dcol_ctx* ctx, int query, int64_t n, const int32_t* shape1, const int32_t* shape2,
                const double* r1, const double* att1, const double* r2, const double* att2,
                int att_kind, double pdip_tol,
                double* alpha, double* x, double* deriv, int32_t* status, int32_t* iters);
int dcol_wait(dcol_ctx* ctx, int ticket);

/* ---- pinned host memory -----------------------------------------------------------------------------------
 * Host-pointer calls accept any host memory, but pageable memory travels through the driver's staging copies
 * (measured on B200: 4 GB/s device-to-host instead of 55 GB/s, and the call cannot overlap its copies with its
 * kernels).  dcol_host_register page-locks a caller-owned buffer in place (cudaHostRegister) so that every later call
 * using it moves it by DMA; register the input and output arrays once and reuse them across calls (registration
 * itself costs milliseconds).  dcol_host_unregister must be called before the memory is freed.  Callers with their
 * own CUDA bindings may pin memory any other way. */
int dcol_host_register(dcol_ctx* ctx, void* ptr, int64_t bytes);
int dcol_host_unregister(dcol_ctx* ctx, void* ptr);

/* ---- the conic solver on its own --------------------------------------------------------------------------
 * Replaces solve_socp(c, G, h, idx_ort, idx_soc1, idx_soc2; pdip_tol) -> (x, s, z) of
 * src/solvers/coneqp/static_solver2.jl:106-191 for n independent problems of common dimensions:
 *     min c'x   s.t.   h - G x  in  R+^n_ort x Q^d1 x Q^d2,     rows ordered [orthant; soc1; soc2] (idx_* of the reference).
 * c: n x nx, G: n x ns x nx (row-major per problem, ns = n_ort + d1 + d2), h: n x ns; outputs x: n x nx, s, z: n x ns.
 * Limits: nx <= 16, ns <= 96, d1, d2 in {0, 2..8}.  status as for the queries (1 = "pdip failed" after 50 iterations).
 * One thread per problem, dense arithmetic: a correctness path for driving the solver subsystem directly (the reference's
 * test/solver_tests.jl fixture runs through it); the proximity entry points above never form G and are the fast path. */
int dcol_solve_socp(dcol_ctx* ctx, int64_t n, int nx, int n_ort, int d1, int d2, const double* c, const double* G, const double* h,
                    double pdip_tol, int mem, double* x, double* s, double* z, int32_t* status, int32_t* iters);
/* The same solve with the per-iteration record the reference prints under verbose = true (static_solver2.jl:117-120,
 * 179-183): trace[(p * DCOL_TRACE_ROWS + it) * DCOL_TRACE_COLS + k], it = 0 .. iters[p] - 1, k = {c'x, s'z/(n_ort + 1),
 * |G x + s - h|, |s o z + sigma mu e|, step length}; rows past iters[p] are left untouched.  `trace` lives where the
 * other arrays live (mem).  The single-pair front ends build the pair's (c, G, h) on the host and print this table. */
#define DCOL_TRACE_ROWS 50
#define DCOL_TRACE_COLS 5
int dcol_solve_socp_trace(dcol_ctx* ctx, int64_t n, int nx, int n_ort, int d1, int d2, const double* c, const double* G, const double* h,
                          double pdip_tol, int mem, double* x, double* s, double* z, int32_t* status, int32_t* iters, double* trace);

/* ---- warm start -------------------------------------------------------------------
 * Opt-in, NOT in the reference (every reference solve starts from `initialize`, static_solver2.jl:86-103) and excluded
 * from parity: iteration counts change; alpha, x and the derivatives agree with a cold solve to the solver tolerance.
 * `state` is DEVICE memory of capacity_pairs x dcol_warm_stride(ctx) doubles, zero-initialised by the caller and
 * addressed by the pair index of the queries that follow (for dcol_se2_constraints: state * n_obs + obstacle; for the
 * scene operators: the `combinations` index).  Every solve leaves behind its iterate (y, s, z) at the first iteration
 * whose duality measure fell below mu_save (<= 0 selects 1e-2); a later solve of the same pair index with the same two
 * shapes starts from that iterate.  Meant for callers that re-evaluate the same pairs at slowly changing poses (the
 * line-search trials and successive iterations of examples/trajectory_optimization/simple_altro.jl:112-163).
 * state == NULL switches it off.  Call dcol_set_shapes before dcol_warm_stride. */
int64_t dcol_warm_stride(const dcol_ctx* ctx);
int     dcol_set_warm_start(dcol_ctx* ctx, double* state, int64_t capacity_pairs, double mu_save);

/* ---- solver diagnostics ------------------------------------------------------------------------------------
 * dcol_set_mu_output: optional per-pair output of the duality measure mu = s'z/deg at the iteration that ended the
 * solve (what the reference's verbose mode prints per iteration, src/solvers/coneqp/static_solver2.jl:117-120,
 * 179-183; verbose printing itself is not reproduced).  `mu` must hold one double per pair of the following queries and
 * live where their outputs live (host memory for DCOL_MEM_HOST calls, device memory otherwise); NULL switches it off.
 * dcol_set_max_iters: TESTING knob for the iteration cap, 1..50.  The reference hard-codes 50 (static_solver2.jl:132)
 * and raises "pdip failed" after the 50th step WITHOUT testing mu again (:188); a smaller cap lets a test land a pair's
 * convergence exactly on the last allowed step.  Every new context starts at 50. */
int dcol_set_mu_output(dcol_ctx* ctx, double* mu);
int dcol_set_max_iters(dcol_ctx* ctx, int max_iters);

/* Instrumentation for bench.py / tests. */
/* kernels launched by this context since creation (solver + derivative + binning kernels) */
int64_t dcol_launch_count(const dcol_ctx* ctx);
/* device time (ms) of the solver launches (the persistent kernels with their fused derivative epilogues: first pass and
 * resume pass of every row tier) of the last query call, measured with CUDA events on the launching streams; the
 * binning / packing / un-permute kernels around them are not included -- except for batches of up to 4 096 pairs, whose
 * single preparation kernel sits inside the bracket (nothing may stand between it and the solver launch, which becomes
 * resident while it still runs); <0 if unavailable */
double  dcol_last_kernel_ms(const dcol_ctx* ctx);
/* measured FP64 FMA throughput of the device in TFLOP/s: DFMA-only kernel launched back to back for `seconds`
 * (0.6 s for the short form), median launch; the roofline denominator of bench.py (MEASURED_PEAKS.json has no FP64 entry) */
double  dcol_measure_fp64_tflops(dcol_ctx* ctx);
double  dcol_measure_fp64_tflops_for(dcol_ctx* ctx, double seconds);
/* three_source = 0: x = fma(x, a, x), the FP64 pipe's peak (what the two calls above return); 1: x = fma(x, a, b), three
 * distinct 64-bit register sources per DFMA -- 92 % of the former on B200 (profiles/r02z_k_dfma_ncu.txt) */
double  dcol_measure_fp64_tflops_mode(dcol_ctx* ctx, double seconds, int three_source);
/* measurement knob: host-pointer calls issue their host<->device copies (same staging, streams, chunking) but no kernels;
 * gives the copy-only bound of the end-to-end path */
int     dcol_set_copy_only(dcol_ctx* ctx, int on);

/* Self-test of the copy threads behind host-pointer calls on pageable memory (no GPU needed): `rounds` batches of ragged
 * slices between two buffers of `bytes` bytes on `threads` threads; 0 if every byte arrived. */
int dcol_host_copy_selftest(int threads, int64_t bytes, int rounds);

/* ---- device groups: one process, several GPUs ------------------------------------------------
 * The reference's callers are single-process Julia programs; a group lets such a caller use every GPU of the box
 * through one call.  A group owns one context per entry of device_ids (a device may appear more than once).  Pairs are
 * independent, so dcol_group_query cuts the pair range into contiguous slices of equal size (sizes differ by at most
 * one, member i takes the i-th slice), runs the same query on every member concurrently (one host thread per member)
 * and returns when all are done.  No data-path collective: every member reads its slice of the caller's HOST arrays
 * and writes its slice of the results; results are bit-identical to the single-context call.
 * query: 0 proximity (deriv unused), 1 gradient (x unused), 2 jacobian.  All pointers are host memory.
 * A group is NOT thread-safe.  dcol_group_ctx returns member i (for dcol_host_register, dcol_set_max_iters, ...);
 * it stays owned by the group.  Errors: the first failing member's return value; dcol_group_last_error names it. */
typedef struct dcol_group dcol_group;
int  dcol_group_create(dcol_group** out, const int* device_ids, int n_devices);
void dcol_group_destroy(dcol_group* g);
const char* dcol_group_last_error(const dcol_group* g);
int  dcol_group_size(const dcol_group* g);
dcol_ctx* dcol_group_ctx(dcol_group* g, int member);
int  dcol_group_set_shapes(dcol_group* g, int n_shapes, const int32_t* type, const int64_t* param_off,
                           const double* params, int64_t n_params);
int  dcol_group_query(dcol_group* g, int query, int64_t n, const int32_t* shape1, const int32_t* shape2,
                      const double* r1, const double* att1, const double* r2, const double* att2,
                      int att_kind, double pdip_tol,
                      double* alpha, double* x, double* deriv, int32_t* status, int32_t* iters);
/* The scene operators over the group (BASELINE config 4): the body arrays go to every member, member
 * i takes the i-th contiguous slice of the i < j enumeration.  Per-pair rows land in the caller's (host) arrays and are
 * bit-identical to the one-device call; the per-body wrench is the sum, in member order, of the members' partial sums
 * (n_bodies x 6 each), so it equals the one-device result up to the order of summation.  A broad phase is set per
 * member: dcol_scene_set_cull(dcol_group_ctx(g, i), alpha_cut). */
int  dcol_group_scene_allpairs(dcol_group* g, int query, int n_bodies, const int32_t* shape, const double* r, const double* att,
                               int att_kind, double pdip_tol,
                               double* alpha, double* x, double* deriv, int32_t* status, int32_t* iters);
int  dcol_group_scene_wrench(dcol_group* g, int n_bodies, const int32_t* shape, const double* r, const double* att, const double* lambda,
                             int att_kind, double pdip_tol,
                             double* alpha, double* wrench, int32_t* status, int32_t* iters);
/* device time (ms) of the solver launches of the last group call: the maximum over the members */
double dcol_group_last_kernel_ms(const dcol_group* g);

#ifdef __cplusplus
}
#endif
#endif /* DCOL_B200_H */
