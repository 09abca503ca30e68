/*
 * sgm_b200.h -- C ABI of libsgm_b200.so, the B200 (sm_100a) evaluation engine for
 * combination-technique sparse-grid interpolants.
 *
 * This is the drop-in boundary for ONE hot path of andreascaglioni/SGMethods:
 *     SGInterpolant.__init__ tables  ->  SGInterpolant.interpolate(x_new, f_on_SG)
 * (reference: sgmethods/sparse_grid_interpolant.py:65-139 and :252-279) together with the
 * tensor-product operators it calls per combination term
 * (sgmethods/tp_interpolant_wrapper.py:41-64, sgmethods/tp_inteprolants.py:38-53, :83-160,
 * :189-277, :322-360) and the Levy-Ciesielski expansion
 * (sgmethods/parametric_expansions_brownian_motion.py:17-71).
 *
 * Conventions
 *   - plain C types only; no torch / numpy types cross this boundary;
 *   - every function returns 0 on success or a negative sgm_status; the message of the
 *     last failure on the calling thread is available from sgm_last_error();
 *   - "host" functions touch no GPU state and work on machines without a GPU;
 *   - device pointers are fp64, row-major, with explicit leading dimensions; buffers are
 *     owned by the caller, the library allocates nothing on the device except the
 *     immutable plan tables;
 *   - kernels are enqueued on the given stream and are asynchronous w.r.t. the host; a plan
 *     is read-only after creation and may be used from several threads / streams.
 */
#ifndef SGM_B200_H
#define SGM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum sgm_status {
    SGM_OK = 0,
    SGM_ERR_INVALID = -1,     /* bad argument (shape, null pointer, unsorted set, ...)      */
    SGM_ERR_CUDA = -2,        /* a CUDA runtime call failed                                 */
    SGM_ERR_OVERFLOW = -3,    /* a table does not fit the packed 32-bit layout              */
    SGM_ERR_NODE_TOL = -4,    /* 1-D knots not separable at the reference's 1e-10 tolerance */
    SGM_ERR_WORKSPACE = -5    /* caller-provided workspace too small                        */
} sgm_status;

/* Operator kinds = the reference's four tensor-product operator classes. */
typedef enum sgm_kind {
    SGM_PW_LINEAR = 0,        /* TPPwLinearInterpolator    tp_inteprolants.py:21-53   */
    SGM_PW_QUADRATIC = 1,     /* TPPwQuadraticInterpolator tp_inteprolants.py:55-160  */
    SGM_PW_CUBIC = 2,         /* TPPwCubicInterpolator     tp_inteprolants.py:162-277 */
    SGM_LAGRANGE = 3,         /* TPLagrangeInterpolator    tp_inteprolants.py:279-360 */
    /* 4 is reserved (internal flavour).  Kind 5 has no reference class: it is the
     * work-reduced HIERARCHICAL-SURPLUS form of an interpolant on nested knots (csrc and
     * sgmethods_b200/hierarchical.py): every (direction, level) entry has exactly ONE basis
     * function that is non-zero at a point, so a multi-index costs one table read and at most
     * k multiplies.  Which function that is, is a property of the knot family (sgm_hier_basis,
     * desc.fam_basis): the hat of the bracket's new knot (pw-linear), the middle cardinal
     * function of the 3-knot patch (pw-quadratic), or the Lagrange cardinal polynomial of the
     * family's one new knot (Lagrange; also the two end knots of the first 3-knot
     * pw-quadratic level).  Mathematically equal to the combination formula, NOT
     * bit-identical to it (it avoids the alternating sum and is the more accurate of the two). */
    SGM_HIER_PW_LINEAR = 5,
    SGM_HIER = 5
} sgm_kind;

/* One-basis-per-entry functions of kind SGM_HIER, per knot family (desc.fam_basis). */
typedef enum sgm_hier_basis {
    SGM_HB_HAT = 0,           /* pw-linear: every bracket has one new knot; its hat at x (linear
                                 extrapolation outside the knots like scipy's RegularGridInterpolator) */
    SGM_HB_QUAD_MID = 1,      /* pw-quadratic: new knots = odd positions = middle knots of the
                                 patches [g_2j, g_2j+2]; value of the middle cardinal parabola of
                                 the patch of x (patch search of tp_inteprolants.py:111-127) */
    SGM_HB_LAG_ONE = 2,       /* Lagrange: the family has ONE new knot p; value of
                                 prod_{i != p} (x - g_i) / (g_p - g_i); NaN when x is a knot of the
                                 family (the reference's barycentric formula gives NaN there) */
    SGM_HB_QUAD_END = 3       /* like LAG_ONE on a 3-knot family, without the NaN rule and with the
                                 pw-quadratic "right of the last knot" flag */
} sgm_hier_basis;

/* Bits reported by sgm_plan_take_flags(): conditions on which the reference raises. */
#define SGM_FLAG_QUADRATIC_RIGHT 1u  /* pw-quadratic point right of the last knot or NaN:
                                        IndexError in the reference (tp_inteprolants.py:131) */

const char* sgm_last_error(void);
int sgm_version(void);

/* ------------------------------------------------------------------------------------------
 * Host table construction (integer work, bit-exact with the reference)
 * ---------------------------------------------------------------------------------------- */

/* Combination-technique coefficients  c_nu = sum_{e in {0,1}^N, nu+e in set} (-1)^{|e|}
 * for every row of a lexicographically sorted, downward-closed multi-index set.
 * Replaces the window search of SGInterpolant.setup_interpolant
 * (sparse_grid_interpolant.py:77-92).  mid_set: card x N, row-major.  coeffs: card. */
int sgm_host_combination_coeffs(const int64_t* mid_set, int64_t card, int32_t N,
                                int64_t* coeffs);

/* Sparse-grid node numbering.  Replaces SGInterpolant.setup_SG
 * (sparse_grid_interpolant.py:109-139): terms are walked in order, the nodes of a term in C
 * order over its active directions (num_nodes > 1, ascending), each node is embedded in R^N
 * with 0 in the inactive directions, nodes closer than `tol` in the 1-norm are merged, rows
 * are numbered by first appearance and keep the coordinates of their first appearance.
 * O(total nodes) via hashing instead of the reference's O(M^2) search.
 *   num_nodes  T x N   nodes per direction of every active term (lev2knots(nu).astype(int))
 *   fam_m      n_fam   the distinct values > 1 in num_nodes (plus 1 if present is ignored)
 *   fam_off    n_fam+1 offsets into fam_knots
 *   fam_knots          knots(m) for each family member, ascending order not required here  */
typedef struct sgm_grid sgm_grid;
int sgm_grid_build(int64_t T, int32_t N, const int32_t* num_nodes, int32_t n_fam,
                   const int32_t* fam_m, const int64_t* fam_off, const double* fam_knots,
                   double tol, sgm_grid** out);
/* Same numbering, with the knot vector of every (term, direction) named by a family index
 * (-1 = single node at 0) instead of by its node count; families may have one knot. */
int sgm_grid_build_families(int64_t T, int32_t N, const int32_t* term_fam, int32_t n_fam,
                            const int64_t* fam_off, const double* fam_knots, double tol,
                            sgm_grid** out);
int64_t sgm_grid_num_nodes(const sgm_grid* g);     /* M                                   */
int64_t sgm_grid_map_size(const sgm_grid* g);      /* sum_t S_t                           */
int64_t sgm_grid_nnz(const sgm_grid* g);           /* non-zero coordinates over all nodes */
/* map_off: T+1 offsets; maps: concatenated C-order maps term -> row of SG (int32). */
int sgm_grid_copy_maps(const sgm_grid* g, int64_t* map_off, int32_t* maps);
/* dense M x N coordinate matrix (the reference's SG attribute), float-bit-exact. */
int sgm_grid_copy_nodes_dense(const sgm_grid* g, double* SG);
/* CSR form of the same matrix: row_off M+1, cols/vals nnz (entries stored even when the
 * first-seen coordinate is a signed zero are NOT listed; they read back as +0.0). */
int sgm_grid_copy_nodes_csr(const sgm_grid* g, int64_t* row_off, int32_t* cols, double* vals);
void sgm_grid_destroy(sgm_grid* g);

/* Tables of the grouped hierarchical-surplus kernel (kind SGM_HIER_PW_LINEAR; layout in
 * csrc/sgm_hier_host.cpp), exposed so that CPU tests can check them without a GPU.
 *   term_fam T x N (family per direction, -1 inactive), fam_extent n_fam (new knots per
 *   family), val_off T (offset of every term's surplus block, first direction fastest).
 * Call with groups = parents = steps = hmap = NULL to get the sizes: groups 12 and parents 8 ints
 * per group, steps 2 ints each, hmap n_slots ints (value slot of the kernel's layout -> index
 * in the caller's layout, -1 = padding); root_slot = slot of the zero multi-index; split
 * n_lists + 1 ints: the groups are dealt to n_lists lists of near-equal cost, list w = groups
 * [split[w], split[w+1]) (one list per warp of the kernel: a fixed summation order).
 * SGM_ERR_INVALID if the set is not downward closed. */
int sgm_host_hier_groups(int64_t T, int32_t N, const int32_t* term_fam, int32_t n_fam,
                         const int32_t* fam_extent, const int64_t* val_off,
                         int32_t n_lists, int64_t* n_groups, int64_t* n_steps, int64_t* n_slots,
                         int32_t* root_slot, int32_t* groups, int32_t* parents, int32_t* steps,
                         int32_t* hmap, int32_t* split);

/* Row matching used by sample_on_SG's recycling (sparse_grid_interpolant.py:214-231): for
 * every row of `xx` (n x N) the index of the unique row of `old_xx` (n_old x N) with
 * 1-norm distance < tol, or -1.  Returns SGM_ERR_INVALID if a row has several matches
 * (the reference asserts).  Hash-accelerated, exact same acceptance test. */
int sgm_host_match_rows(const double* xx, int64_t n, const double* old_xx, int64_t n_old,
                        int32_t N, double tol, int64_t* match);

/* ------------------------------------------------------------------------------------------
 * Device plan = packed tables of one interpolant (immutable after creation)
 * ---------------------------------------------------------------------------------------- */
typedef struct sgm_plan_desc {
    int32_t kind;              /* sgm_kind                                                  */
    int32_t N;                 /* parametric dimension                                      */
    int64_t T;                 /* active terms (c != 0), reference order                    */
    int64_t M;                 /* rows of f_on_SG this plan addresses                       */
    const double* coeffs;      /* T    combination coefficients (exact integers in fp64)    */
    const int32_t* term_fam;   /* T*N  knot family of each direction, -1 = inactive (1 node) */
    const int64_t* map_off;    /* T+1                                                       */
    const int32_t* maps;       /* map_off[T] rows of f_on_SG, C order per term              */
    int32_t n_fam;             /* knot families (SGInterpolant: one per distinct node count) */
    const int64_t* fam_off;    /* n_fam+1 offsets into fam_knots / fam_lambda               */
    const double* fam_knots;   /* knots of each family, ascending, >= 2 per family          */
    const double* fam_lambda;  /* barycentric lambda per knot (SGM_LAGRANGE only, else NULL) */
    const int32_t* fam_newrank;/* SGM_HIER only: per knot its rank among the knots of the family
                                  that are NEW on its level (and belong to this family), -1 for
                                  the others; term tables then have one entry per new-knot tuple */
    const int32_t* fam_basis;  /* SGM_HIER only: n_fam sgm_hier_basis values, NULL = all SGM_HB_HAT */
    int32_t n_seg;             /* 0: ordinary plan.  > 0: SEGMENTED plan -- the terms form n_seg
                                  consecutive groups, each an independent small interpolant
                                  (e.g. the hierarchical surplus of one refinement candidate);
                                  evaluated by sgm_eval_segments / sgm_surplus_norms only     */
    const int64_t* seg_off;    /* n_seg+1 term offsets, seg_off[0] = 0, seg_off[n_seg] = T    */
} sgm_plan_desc;

typedef struct sgm_plan sgm_plan;
int sgm_plan_create(const sgm_plan_desc* desc, int device, sgm_plan** out);
void sgm_plan_destroy(sgm_plan* plan);
/* Re-reads the SGM_* development switches from the environment; they are
 * otherwise read once, when the plan is created.  For tests and experiment scripts. */
int sgm_plan_reload_knobs(sgm_plan* plan);
/* Bytes of packed tables resident on the device / walked once per launch. */
int64_t sgm_plan_table_bytes(const sgm_plan* plan);
/* sum_t C_t = corner (or node) contributions per point and output: 2^k, 3^k, 4^k or S_t. */
int64_t sgm_plan_corners_per_point(const sgm_plan* plan);
/* Plan-specialised kernel for few-term pw-linear plans (csrc/sgm_jit.cpp): when such a plan has
 * been asked for SGM_JIT_MIN_POINTS (default 4e6) points in total, sgm_eval generates CUDA source
 * with the plan's tables as literals, compiles it with NVRTC for the device and uses it from
 * then on (same arithmetic, bit-identical; example/1's shape at 10^7 points: 45 % -> 70 % of the
 * HBM peak).  Returns 1 = in use, 0 = not built (yet), -1 = unavailable (NVRTC / driver library
 * missing, compilation failed: the generic kernel stays), 2 = being built; `why` receives a short
 * description. */
int sgm_plan_jit_state(sgm_plan* plan, char* why, size_t why_bytes);
/* The build runs on a background thread (evaluations use the generic kernel meanwhile; 2 = being
 * built); this waits for a build in progress and returns the state.  SGM_JIT_SYNC=1 builds inside
 * the triggering sgm_eval instead, SGM_NO_JIT=1 never builds. */
int sgm_plan_jit_wait(sgm_plan* plan);
/* Host only (no GPU): the generated source for `desc` with `threads` per block and
 * `points_per_thread`; SGM_ERR_INVALID when the plan does not qualify.  `*needed` = bytes of the
 * source incl. the terminator.  compile_cc > 0 (e.g. 100) additionally compiles it with NVRTC
 * for sm_<cc>a and reports a failure as SGM_ERR_CUDA. */
int sgm_host_small_kernel_source(const sgm_plan_desc* desc, int32_t threads, int32_t points_per_thread,
                                 int32_t compile_cc, char* buf, size_t buf_bytes, size_t* needed);
/* Reads and clears the condition flags set by kernels launched with this plan
 * (synchronises `stream`). */
int sgm_plan_take_flags(sgm_plan* plan, void* stream, uint32_t* flags);

/* Scratch the caller must provide to sgm_eval for P points and NF outputs (device bytes). */
size_t sgm_eval_workspace_bytes(const sgm_plan* plan, int64_t P, int64_t NF);

/* out[p, 0:NF] = sum_t c_t * TP_t(x[p, active dims of t])   -- replaces the Python term loop
 * of SGInterpolant.interpolate (sparse_grid_interpolant.py:268-279), the wrapper's column
 * purge / constant path (tp_interpolant_wrapper.py:54-64) and the four operator classes.
 *   x    P x ldx (ldx >= N; only the first N columns are read)           device, fp64
 *   f    M x ldf (ldf >= NF) values on the sparse grid, row-major         device, fp64
 *   out  P x ldo (ldo >= NF)                                              device, fp64
 * Arithmetic order per (point, output) follows the reference exactly.  */
int sgm_eval(sgm_plan* plan, const double* x, int64_t P, int64_t ldx,
             const double* f, int64_t M, int64_t NF, int64_t ldf,
             double* out, int64_t ldo, void* workspace, size_t workspace_bytes,
             void* stream);

/* Batched evaluation of a SEGMENTED plan: for every segment s (a consecutive group of terms)
 *   out[s * seg_stride + p * ldo + c] = sum_{t in segment s} c_t * TP_t(x[p])[c]
 * in ONE launch: the bracket tables of a point are built once for all segments.  Replaces the
 * candidate loop of compute_GG_indicators_RM (sgmethods/compute_error_indicators.py:109-120:
 * per reduced-margin index a new SGInterpolant, its sampling and a full interpolate) by the
 * hierarchical surpluses Delta_nu u = sum_{e <= supp nu} (-1)^{|e|} TP_{nu-e}[u] of all
 * candidates, 2^{|supp nu|} terms each.  seg_stride >= (P-1)*ldo + NF. */
int sgm_eval_segments(sgm_plan* plan, const double* x, int64_t P, int64_t ldx,
                      const double* f, int64_t M, int64_t NF, int64_t ldf,
                      double* out, int64_t ldo, int64_t seg_stride,
                      void* workspace, size_t workspace_bytes, void* stream);
/* sumsq[s] = sum_{p < P, c < NF} vals[s * seg_stride + p * ld + c]^2, fixed summation order. */
int sgm_segment_sumsq(const double* vals, int32_t n_seg, int64_t P, int64_t NF, int64_t ld,
                      int64_t seg_stride, double* sumsq, void* stream);
/* Both steps with the surpluses kept in the caller's workspace (never copied to the host):
 * sumsq[s] = sum_p |Delta_s u(x_p)|^2, i.e. P times the squared Monte-Carlo L2 norm the
 * reference's indicator callback computes (compute_error_indicators.py:119). */
size_t sgm_surplus_norms_workspace_bytes(const sgm_plan* plan, int64_t P, int64_t NF);
int sgm_surplus_norms(sgm_plan* plan, const double* x, int64_t P, int64_t ldx,
                      const double* f, int64_t M, int64_t NF, int64_t ldf, double* sumsq,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Multilevel combination (reference: MLInterpolant.interpolate / get_ml_terms,
 * sgmethods/multilevel_interpolant.py:89-150): for n_parts plans sharing the points x,
 *   out = sum_i sign_i * I_{plan_i}[f_i](x)
 * evaluated part after part on one stream with a fused signed accumulation into `out`
 * (first part overwrites).  f[i] is M_i x ldf[i]. */
int sgm_eval_signed_sum(sgm_plan* const* plans, const double* signs, int32_t n_parts,
                        const double* x, int64_t P, int64_t ldx,
                        const double* const* f, const int64_t* M, const int64_t* ldf,
                        int64_t NF, double* out, int64_t ldo,
                        void* workspace, size_t workspace_bytes, void* stream);

/* One batch of the dimension-wise hierarchisation used by SGM_HIER_PW_LINEAR plans:
 *   alpha[child[i], :] -= wa[i] * alpha[pa[i], :] + wb[i] * alpha[pb[i], :]   (pb[i] < 0: one parent)
 * for i < n; children of a batch are distinct and never parents of the same batch.
 * alpha is M x ld (fp64, device), NF columns are updated. */
int sgm_hier_apply(double* alpha, int64_t NF, int64_t ld, const int32_t* child, const int32_t* pa,
                   const int32_t* pb, const double* wa, const double* wb, int64_t n, void* stream);

/* All batches of the hierarchisation in one call: batch b covers entries
 * [batch_off[b], batch_off[b+1]) of the concatenated child / pa / pb / wa / wb arrays (device);
 * batch_off is given both on the host and on the device (n_batches + 1 entries).  Small
 * tables run as ONE launch of one thread-block cluster, large ones as one launch per batch. */
int sgm_hier_apply_batches(double* alpha, int64_t NF, int64_t ld, const int32_t* child,
                           const int32_t* pa, const int32_t* pb, const double* wa, const double* wb,
                           const int64_t* batch_off_host, const int64_t* batch_off_dev,
                           int32_t n_batches, void* stream);

/* Validation of a work-reduced evaluation against the exact one (method="auto",
 * sparse_grid_interpolant.py of this repository; no reference counterpart): writes
 *   out[0] = max_i |a_i - b_i| / max(max_i |b_i|, tiny)      over the n entries,
 * where an entry that is NaN in BOTH arrays counts as agreement (the reference's Lagrange
 * operator returns NaN on a knot) and an entry that is NaN in one of them makes the result NaN.
 * max |b| = 0 divides by 1.  a, b, out: device, fp64; one small launch, no host sync. */
int sgm_max_rel_deviation(const double* a, const double* b, int64_t n, double* out, void* stream);

/* W[p, :] = Levy-Ciesielski expansion of Brownian motion on [0, T] at times tt for the
 * parameter vector yy[p, 0:d]  (reference: param_LC_Brownian_motion,
 * sgmethods/parametric_expansions_brownian_motion.py:17-71, one call per row there).
 *   tt nt, yy P x ldy (ldy >= d), W P x ldw (ldw >= nt), level_scale[l-1] = 2^((l-1)/2) for
 *   l = 1..L with L = ceil(log2 d) and sqrtT = sqrt(T) (host-computed so that the factors
 *   are bit-identical to the reference's python/numpy values). */
int sgm_lc_brownian(const double* tt, int64_t nt, const double* yy, int64_t P, int32_t d,
                    int64_t ldy, double T, double sqrtT, const double* level_scale, int32_t L,
                    double* W, int64_t ldw, void* stream);

/* W[p, :] = sum_n factor[n] * sin(freq[n] * tt) * yy[p, n]  (param_KL_Brownian_motion as
 * shipped, parametric_expansions_brownian_motion.py:73-91; factor/freq host-computed). */
int sgm_kl_brownian(const double* tt, int64_t nt, const double* yy, int64_t P, int32_t d,
                    int64_t ldy, const double* factor, const double* freq,
                    double* W, int64_t ldw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SGM_B200_H */
