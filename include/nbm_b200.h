/*
 * nbm_b200.h — C ABI of the B200-native neural-bootstrapping (NBM) training step for the
 * interfacial Poisson problem (drop-in for the hot path behind JAX-DIPS
 * `trainer.setup -> init_fn -> solve_fn`, jax_dips/solvers/poisson/trainer.py:980-1137).
 *
 * The reference has NO native interface for this path: the seam is the pure function
 * `Trainer.loss(params, points, dx, dy, dz)` differentiated by `jax.value_and_grad`
 * (trainer.py:786, 826, 893).  These entry points are what an XLA-FFI / ctypes binding of that
 * seam binds to.  Conventions:
 *
 *   - extern "C", plain pointers and sizes only.  Every pointer is a DEVICE pointer unless its
 *     name ends in `_host`.  The caller owns every buffer; the library allocates nothing that
 *     outlives a call except the per-device `__constant__` copy of the network parameters.
 *   - every call only enqueues work on `stream` (no host synchronisation), so a sequence of
 *     calls is CUDA-graph capturable.  Exception: functions documented as "synchronous".
 *   - return value: 0 = ok, otherwise an nbm_status; nbm_last_error() gives a message
 *     (thread-local).
 *   - all arithmetic is fp32 (the reference runs with jax_enable_x64 = False).
 *   - grids are z-fastest: index = (i*Ny + j)*Nz + k   (jax_dips/domain/mesh.py:121-153).
 *
 * "Lattice": a tensor-product set of sites given by three 1-D coordinate arrays.  The training
 * grid, its +-d shifted copies (stencil sites) and the rank-local node lattice with halo are all
 * lattices.  "Site": a position at which u^-/u^+ are needed (discretization.py:426).  "Crossed":
 * the cell of size d centred on the site has corner level-set values of mixed sign
 * (geometric_integrations_per_point.py:203-263).
 */
#ifndef NBM_B200_H
#define NBM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* nbm_stream_t; /* cudaStream_t */

enum nbm_status {
    NBM_OK = 0,
    NBM_ERR_BAD_ARG = 1,      /* null pointer / non-positive size / inconsistent dims */
    NBM_ERR_UNSUPPORTED = 2,  /* network shape or option outside the kernel contract */
    NBM_ERR_CUDA = 3,         /* a CUDA runtime call failed; see nbm_last_error() */
    NBM_ERR_WORKSPACE = 4     /* workspace too small */
};

const char* nbm_last_error(void);
int nbm_version(void);

/* ------------------------------------------------------------------------------------------
 * Level-set grid (a14).  Replaces interpolate.add_ghost_layer_3d (domain/interpolate.py:762-816)
 * + multilinear_interpolation (:906-1021) / nonoscillatory_quadratic_interpolation_per_point
 * (:388-569) + level_set.perturb_level_set_fn (geometry/level_set.py:34-48).
 * ---------------------------------------------------------------------------------------- */
enum nbm_interp { NBM_INTERP_TRILINEAR = 0, NBM_INTERP_QUADRATIC = 1 };

typedef struct {
    const float* phi_g;      /* ghosted level set, (nx+2)*(ny+2)*(nz+2), z fastest */
    const float* xg;         /* ghosted node coordinates, nx+2 */
    const float* yg;         /* ny+2 */
    const float* zg;         /* nz+2 */
    int gx, gy, gz;          /* ghosted dims = n+2 */
    int interp;              /* nbm_interp */
    float perturb_eps;       /* 1e-10 when the level set is wrapped in perturb_level_set_fn, else 0 */
    /* SAMPLED level set (an analytic lvl_set_fn, as tests/test_poisson.py passes one: the reference calls the user's
     * callable wherever it needs phi, discretization.py:90).  A callable cannot run inside a kernel, but every level-set
     * evaluation of the path happens at positions known per level: the host evaluates the callable there (same fp32
     * position arithmetic as the kernels) and hands the values in; phi_g/xg/yg/zg may then be NULL:
     *   corner_phi[c*8 + v]  phi at the 8 cell corners of crossed site c (corner order of
     *                        geometric_integrations_per_point.py:212-224)        -> nbm_cutcell_f32
     *   cube_phi[c*27 + q]   phi at the 27 cube vertices s + X_q of crossed site c  -> nbm_regression_f32
     *   eval_phi[i*7 + k]    phi at evaluation point i and at i +- dx, +- dy, +- dz (k = 0, x-, x+, y-, y+, z-, z+)
     *                                                                             -> nbm_evaluate_f32
     * (classification is then done by the host too: nbm_classify_f32 needs the grid form.) */
    const float* corner_phi;
    const float* cube_phi;
    const float* eval_phi;
} nbm_lvl_t;

/* phi (nx,ny,nz) -> phi_g (nx+2,ny+2,nz+2) and x -> xg etc. (linear extrapolation, x then y then z). */
int nbm_ghost_layer_f32(const float* phi, const float* x, const float* y, const float* z,
                        int nx, int ny, int nz,
                        float* phi_g, float* xg, float* yg, float* zg, nbm_stream_t stream);

/* phi at arbitrary points; pts is (n,3) row-major.  out[n]. */
int nbm_phi_interp_f32(const nbm_lvl_t* lvl, const float* pts, int64_t n, float* out, nbm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Lattices and site classification (a12: is_point_cell_crossed_by_interface,
 * geometry/geometric_integrations_per_point.py:203-263)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const float* xs;   /* nx site coordinates */
    const float* ys;
    const float* zs;
    int nx, ny, nz;
    /* only sites with index inside [lo, hi) per axis are classified; others get flag = 2 */
    int lo[3], hi[3];
    /* the lattice is replicated n_shift times (1..7), copy k displaced by shift[k] (fp32 add, as the
     * reference forms stencil points `point[0] - dx`, discretization.py:348-353).  Site id of
     * (copy k, node e) = k*nx*ny*nz + e. */
    int n_shift;
    float shift[7][3];
} nbm_lattice_t;

/* flag[e] in {-1, 0 (crossed), +1, 2 (not a site)};  side[e]: bit0 = phi>=0 (MLP.py:98 head
 * select), bit1 = phi>0 (discretization.py:484 branch select).  side is written for EVERY node. */
int nbm_classify_f32(const nbm_lvl_t* lvl, const nbm_lattice_t* lat, float dx, float dy, float dz,
                     int8_t* flag, uint8_t* side, nbm_stream_t stream);

/* Stream compaction of crossed sites (flag == 0): idx_out[0..*count_dev) = ascending site ids,
 * cidx[site] = position in the list or -1.  idx_out must hold `capacity` entries; count_dev is a
 * device int64.  Workspace: call with workspace == NULL to get *ws_bytes. */
int nbm_compact_crossed(const int8_t* flag, int64_t n, int64_t* idx_out, int64_t capacity,
                        int32_t* cidx, int64_t* count_dev,
                        void* workspace, size_t* ws_bytes, nbm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K1: cut-cell geometry of crossed cells (a10-a13: compute_cell_faces_areas_values :464-1009,
 * integrate_over_interface_at_point :383-421, get_vertices_of_cell_intersection_with_interface
 * :201-367, vol_fn/area_fn :370-380)
 * ---------------------------------------------------------------------------------------- */
/* For each crossed site c (lattice `lat`, site index idx[c]):
 *   frac[c*14 + 0..11] = area^- , area^+ for faces (x-,x+,y-,y+,z-,z+) interleaved (m,p)
 *   frac[c*14 + 12,13] = V^-, V^+
 *   tri[c*90 + t*9 + v*3 + a] = Gamma triangle t (<=10), vertex v, axis a (unused slots zero)
 *   tri_area[c*10 + t]
 */
int nbm_cutcell_f32(const nbm_lvl_t* lvl, const nbm_lattice_t* lat, float dx, float dy, float dz,
                    const int64_t* idx, int64_t n_crossed,
                    float* frac, float* tri, float* tri_area, nbm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K2a: regression extrapolation geometry at crossed sites (a9: get_regression_coeffs_at_point,
 * discretization.py:238-296; normal_point_fn :199-218)
 *   pos[c*3..]   site position s
 *   proj[c*3..]  projected point s - phi(s) n          (discretization.py:467)
 *   delta[c]     phi(s)
 *   Cm[c*27+q], Cp[c*27+q] = n . D^-/+ column q  (pinv(X^T W X)(W X)^T, jnp.linalg.pinv cutoff)
 *   cube_side[c] bit q = phi(s + X_q) >= 0
 * ---------------------------------------------------------------------------------------- */
int nbm_regression_f32(const nbm_lvl_t* lvl, const nbm_lattice_t* lat, float dx, float dy, float dz,
                       const int64_t* idx, int64_t n_crossed,
                       float* pos, float* proj, float* delta, float* Cm, float* Cp,
                       uint32_t* cube_side, nbm_stream_t stream);

/* K2b: jump-condition weights at crossed sites (a8, discretization.py:464-513).  The value on the
 * far side of the interface is  E = sum_q B[c*28+q] * u(s + X_q) + B[c*28+27]  (B[13] includes
 * the centre term).  mu_*_s sampled at s, alpha/beta/mu_*_proj sampled at proj. */
int nbm_site_weights_f32(int64_t n_crossed, const float* delta, const float* Cm, const float* Cp,
                         const float* mu_m_s, const float* mu_p_s,
                         const float* alpha_proj, const float* beta_proj,
                         const float* mu_m_proj, const float* mu_p_proj,
                         float* B, nbm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K2c: row assembly (a7: compute_Ax_and_b_preconditioned_fn, discretization.py:299-423)
 * ---------------------------------------------------------------------------------------- */
enum nbm_nonlinear { NBM_NL_NONE = 0, NBM_NL_SINH = 1 /* N(u) = coef * sinh(u) */ };

typedef struct {
    /* points: lattice 0; n_points = nx*ny*nz of `pts` */
    nbm_lattice_t pts;
    float dx, dy, dz;
    float bounds[6];                 /* xmin,xmax,ymin,ymax,zmin,zmax of the LVL grid (discretization.py:60-65) */
    /* site access.  shared = 1: sites are nodes of one lattice `site_dims` and slot k of point
     * (i,j,k) is node (i+pt_off[0]+e_k ...);  shared = 0: 7 lattices of the points' dims, site id
     * = k*n_points + p. */
    int shared;
    int site_dims[3];
    int pt_off[3];
    const int8_t* flag;              /* per site */
    const uint8_t* side;             /* per site */
    const int32_t* cidx;             /* per site: crossed index or -1 */
    const float* frac;               /* per crossed site (only lattice-0 / node entries are read) */
    const float* beta_gamma;         /* per crossed site: integral_Gamma beta  (host: sum_t area_t*mean beta) */
    /* coefficient samples, per point */
    const float* mu_m_faces;         /* [6][n_points]  (x-,x+,y-,y+,z-,z+) at face centres */
    const float* mu_p_faces;
    const float* k_m; const float* k_p; const float* f_m; const float* f_p; const float* g_dir;
    /* outputs */
    float* w;                        /* [7][n_out]: weight on u(site k) after division by diag */
    float* rhs;                      /* [n_out] */
    float* nl;                       /* [2][n_out] V^-/diag, V^+/diag (may be NULL when nonlinear = none) */
    int32_t* irr;                    /* [n_out] index into the irregular list or -1 */
    int64_t n_out; int64_t out_stride[3]; int64_t out_off; /* output index = out_off + i*s0 + j*s1 + k*s2 */
    /* irregular rows (a point with >=1 crossed stencil site) */
    int64_t irr_capacity;
    int64_t* irr_count;              /* device counter (zeroed by the caller) */
    int64_t* irr_point;              /* [cap] output index of the point */
    float* irr_wE;                   /* [cap][7] weight on E(site k) */
    int32_t* irr_c;                  /* [cap][7] crossed index of site k or -1 */
    uint8_t* irr_nl;                 /* [cap] 0: none, 1: N^-(E(centre)), 2: N^+(E(centre)) enters the row */
    float* irr_nlw;                  /* [cap] its weight V/diag */
    /* faces = 1 (shared layout only, needs k_m = k_p = 0): instead of the 7 row weights `w`, store per node the
     * UN-normalised coefficient mu A / d of its +x, +y, +z faces (the FV matrix is symmetric between regular
     * neighbours: one value per face serves the rows on both sides), 1/diag, and move every irregular row
     * (a crossed stencil site, or a neighbour on the other side of the interface) completely into the
     * list (irr_wU, irr_rhs).  12 + 4 bytes per node instead of 28.
     *   cface[3][n_out];  dinv[n_out]: > 0 regular row, 0 no dense row, -1 Dirichlet row (r = u - rhs);
     *   kv[n_out] = k^- V^- + k^+ V^+ (un-normalised), may be NULL when k_m = k_p = 0 everywhere */
    int faces;
    float* cface;
    float* dinv;
    float* irr_wU;                   /* [cap][7] (faces mode) weight on u(site k) */
    float* irr_rhs;                  /* [cap]    (faces mode) */
    float* kv;                       /* [n_out]  (faces mode) or NULL */
    /* learned preconditioner input (nn/preconditioner.py; discretization.py:337-339): the 26-vector coeffs_ of
     * every point, [mu A/d (imh-,imh+,iph-,iph+,jmh-,...,kph+), V-, V+, A (same order)]
     * (geometric_integrations_per_point.py:873-903, :964-996), SoA [26][n_out]; NULL = not wanted */
    float* coef26;
} nbm_assemble_t;

int nbm_assemble_f32(const nbm_assemble_t* a, nbm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K3/K4: the training step (a6, a15, a16, a17)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int layers_p, hidden_p, layers_m, hidden_m;  /* model_dict["mlp"], tanh */
} nbm_net_t;

int nbm_net_num_params(const nbm_net_t* net);

/* copy the flat parameter vector (device) into the per-device __constant__ bank */
int nbm_upload_params(const nbm_net_t* net, const float* params, nbm_stream_t stream);

typedef struct {
    nbm_net_t net;
    int nonlinear_m, nonlinear_p;    /* nbm_nonlinear */
    float nl_coef_m, nl_coef_p;
    /* node lattice with halo */
    const float* xe; const float* ye; const float* ze;
    int ex, ey, ez;
    const uint8_t* side;             /* [ex*ey*ez] */
    /* row tables in lattice layout (zero rows where no training point) */
    const float* w;                  /* [7][ne] */
    const float* rhs;                /* [ne] */
    const float* nl;                 /* [2][ne] or NULL */
    /* crossed sites */
    int64_t n_crossed; const int64_t* c_node; const float* B;  /* [nc][28] */
    /* irregular rows */
    int64_t n_irr; const int64_t* irr_point; const float* irr_wE; const int32_t* irr_c; const uint8_t* irr_nl; const float* irr_nlw;
    float inv_n_points;              /* 1 / (number of points the mean runs over) */
    /* work buffers */
    float* U; float* R; float* G;    /* [ne] each */
    float* E; float* gE;             /* [nc] each */
    float* partials;                 /* [n_partial_rows][P+1] */
    int n_partial_rows;              /* >= nbm_step_partial_rows() */
    float* loss_grad;                /* [P+1]: grad[0..P), loss at [P] */
    int stages;                      /* 0 = the whole step; else a bit mask of nbm_stage (profiling / timing) */
    /* faces mode (see nbm_assemble_t): w is unused, rows come from cface/dinv, irregular rows from the list */
    int faces;
    const float* cface; const float* dinv; const float* irr_wU; const float* irr_rhs;
    const float* kv;                 /* may be NULL */
    /* fused adjoint (faces mode, no nonlinear operator): when S != NULL the residual pass also writes
     * S = dinv * R of the regular rows, the dense adjoint stencil is NOT run as a kernel but evaluated inside
     * the gradient kernel from row tables staged plane by plane into shared memory with bulk async copies
     * (TMA, mbarrier completion).  G then only collects the list contributions (irregular rows, extrapolation):
     * it must be zero on entry (the gradient kernel re-zeroes what it consumed), nodes that can receive such a
     * contribution carry bit 2 (value 4) in `side`, and `side` must be readable up to the next multiple of 16
     * bytes past ex*ey*ez. */
    float* S;                        /* [ne] or NULL */
    /* learned preconditioner P = 0.5 + pc_scale * sigmoid(MLP(coeffs_)) (nn/preconditioner.py:10-35), tanh MLP
     * 26 -> pc_d1 -> pc_d2 -> 1, multiplying lhs/diag and rhs/diag of every row (discretization.py:418-419).
     * pc_params (device): Dense_0.kernel (26 x d1 row-major), Dense_0.bias, Dense_1.kernel (d1 x d2),
     * Dense_1.bias, Dense_2.kernel (d2), Dense_2.bias; normally the tail of the flat parameter vector.
     * With coef26 != NULL the gradient has net + pc entries: every partial row is `row_stride` floats
     * (>= n_net + n_pc + 1, loss in the last used column n_net + n_pc), rows [0, n_sm) are written by the
     * network-gradient kernel, the following n_pc_rows by the preconditioner kernel; the buffer must be zero
     * on first use.  loss_grad has n_net + n_pc + 1 entries.  Compiled: (d1, d2) = (8, 4). */
    const float* coef26;             /* [26][ne] or NULL */
    const float* pc_params;
    int pc_d1, pc_d2;
    float pc_scale;
    int n_pc_rows;                   /* partial rows reserved for the preconditioner kernel (>= 1) */
    /* optional fast path: pc_nodes_m / pc_nodes_p = the lattice nodes that hold a row and whose cell is NOT crossed, per
     * side (ascending), pc_d = the cell size.  They take per-side kernels whose first layer sees only the 6 face
     * coefficients of the node's side (the other 20 inputs are zero or grid constants); crossed cells (c_node) take
     * the generic kernel.  Needs n_pc_rows >= 3 (one third of the rows per kernel; 3 x #SM is a good value). */
    const int32_t* pc_nodes_m; int64_t n_pc_m; const int32_t* pc_nodes_p; int64_t n_pc_p; float pc_d[3];
    /* Deterministic adjoint of the lists (optional; all NULL = the atomic kernels).  The adjoint of the irregular rows
     * and of the extrapolation scatters into gE and G; given the transposed incidence in CSR form it is GATHERED
     * instead - no atomics, fixed summation order, bitwise reproducible step:
     *   ge_ptr[nc+1], ge_ent[]: per crossed site c the (row q, slot k) pairs with irr_c[q][k] == c, packed q*8+k
     *   list_nodes[n_list]    : the lattice nodes that receive a list contribution, ascending
     *   g_ptr[n_list+1], g_ent[]: per such node its contributions: >= 0: q*8+k -> irr_wU[q][k] * R[irr_point[q]]
     *                            (faces table only);  < 0: -(c*32+v)-1 -> B[c][v] * gE[c] */
    const int32_t* ge_ptr; const int32_t* ge_ent;
    const int64_t* list_nodes; int64_t n_list;
    const int32_t* g_ptr; const int32_t* g_ent;
    /* Dense stencil stage with the faces table: 0 = automatic (residual rows + adjoint stencil as ONE kernel fed by 3-D
     * TMA boxes, `stencil_tma_kernel`, whenever lattice rows are 16-byte multiples, no preconditioner sits between the
     * two stages and the deterministic list gathers are not combined with a nonlinear operator), -1 = always the two
     * separate kernels. */
    int stencil_tma;
    /* Activation stash (optional, NULL = recompute): ceil(hidden_p / 4) planes of ex*ey*ez 16-byte chunks
     * (48 bytes per node for hidden_p = 10), 16-byte aligned.  The forward kernel writes the last hidden layer of
     * every plus-side node there; the gradient kernel reads it back (cp.async ring in shared memory) instead of
     * recomputing that layer (for the default 3-10-10-1 head: 100 of the 140 forward FMAs and half of the tanh
     * evaluations of its forward part). */
    float* Hst;
    /* List chain beside the dense stencil (optional; NULL = the list kernels run in line after the dense stencil).
     * With G2 [ex*ey*ez, zero on first use; the step re-zeroes what it touched], Rq [n_irr] and list_nodes / n_list
     * (every lattice node that can receive a contribution of the lists: the 27-cube of every crossed site and the 7
     * stencil sites of every irregular row, ascending; g_ptr must be NULL) the whole step with the TMA stencil runs as
     *   stream:       fwd_nodes -> stencil_tma ------------------------------> merge_lists -> node_grad
     *   side stream:           \-> extrap -> irregular rows fwd + bwd -> extrap adjoint -/
     * The side chain needs only U: it keeps the residuals of the irregular rows in Rq and scatters its part of
     * d loss/d U into G2 with fp32 atomics; merge_lists then adds G2 into G and stores Rq into R.  The side stream
     * and its two events are library-owned (one set per device); the fork/join is CUDA-graph capturable, and all
     * work is ordered after what `stream` held on entry and before what is enqueued on `stream` afterwards. */
    float* G2; float* Rq;
    /* Optional transposed copies of the list tables for the chain kernels (thread per site / per row: with [slot][item]
     * a warp reads 32 consecutive floats per slot instead of 32 lines): B_soa [28][n_crossed], irr_c_soa / irr_wE_soa /
     * irr_wU_soa [7][n_irr].  NULL = the item-major tables above. */
    const float* B_soa; const int32_t* irr_c_soa; const float* irr_wE_soa; const float* irr_wU_soa;
} nbm_shared_step_t;

/* number of preconditioner parameters for hidden widths (d1, d2) */
int nbm_precond_num_params(int d1, int d2);

/* stages of the shared-evaluation step, in launch order */
enum nbm_stage {
    NBM_STAGE_FWD = 1,        /* U = u(node) for every lattice node */
    NBM_STAGE_EXTRAP = 2,     /* far-side values E at crossed nodes */
    NBM_STAGE_RESIDUAL = 4,   /* rows: 7-point stencil on U (+ irregular rows) */
    NBM_STAGE_ADJOINT = 8,    /* G = d loss / d U (adjoint stencil + adjoint of irregular rows / extrapolation) */
    NBM_STAGE_GRAD = 16,      /* forward recompute + backward per node, per-CTA partial sums */
    NBM_STAGE_REDUCE = 32,    /* partial rows -> [grad, loss] */
    /* modifiers for timing a selection: leave out the list kernels (crossed sites, irregular rows) / the dense kernels */
    NBM_STAGE_NO_LISTS = 64,
    NBM_STAGE_NO_DENSE = 128
};

int nbm_step_partial_rows(void);

/* FP32 FMA-pipe micro-benchmark (the roofline denominator MEASURED_PEAKS.json lacks): every thread
 * runs `iters` x 16 independent FFMAs; returns the number of FLOPs issued through *flops_host (host
 * pointer); the caller times it with events on `stream`.  out[1] receives a checksum. */
int nbm_ffma_probe_f32(int iters, float* out, double* flops_host, nbm_stream_t stream);

/* loss and d loss/d params for the rows of the plan, network parameters taken from the
 * __constant__ bank (call nbm_upload_params first).  Native-spacing "shared evaluation" path:
 * the network is evaluated once per lattice node. */
int nbm_loss_grad_shared_f32(const nbm_shared_step_t* s, nbm_stream_t stream);

typedef struct {
    nbm_net_t net;
    int nonlinear_m, nonlinear_p;
    float nl_coef_m, nl_coef_p;
    const float* xs; const float* ys; const float* zs;   /* training-grid coordinates */
    int nx, ny, nz;
    int64_t p0, p1;                  /* the batch: flattened point range [p0, p1) */
    float dx, dy, dz;                /* cell size of this level */
    const uint8_t* side;             /* [7][n_points] */
    const float* w;                  /* [7][n_points] */
    const float* rhs;                /* [n_points] */
    const float* nl;                 /* [2][n_points] or NULL */
    const int32_t* irr;              /* [n_points] */
    int64_t n_crossed; const int64_t* c_site; const float* c_pos; const uint32_t* c_cube_side; const float* B;
    int64_t n_irr; const float* irr_wE; const int32_t* irr_c; const uint8_t* irr_nl; const float* irr_nlw;
    float inv_n_points;
    float* E; float* gE;
    float* partials; int n_partial_rows;
    float* loss_grad;
    float* rows;                     /* optional [n_points]: lhs/diag - rhs/diag of every row of the batch (may be NULL) */
    /* the 7 displaced lattices: coordinates [7][nx], [7][ny], [7][nz] (slot k = grid + shift_k, formed in fp32 like
     * `point[0] - dx`, discretization.py:348-353) and work arrays [7][n_points] */
    const float* xs7; const float* ys7; const float* zs7;
    float* U7; float* G7;
    /* learned preconditioner (see nbm_shared_step_t): coef26 [26][nx*ny*nz], Pc work array [nx*ny*nz]; `rows`
     * must then be non-NULL (it carries the un-preconditioned residuals between the kernels); every partial row is
     * n_net + n_pc + 1 floats and n_pc_rows extra rows are needed after the 7*grid + rows + extrap rows */
    const float* coef26; const float* pc_params; float* Pc;
    int pc_d1, pc_d2; float pc_scale; int n_pc_rows;
    /* Zoom level 1 with whole-plane batches (cell size = half the grid spacing, data_management.py:320-326): the stencil
     * site p + d e_a of a point is the site p' - d e_a of its neighbour p' = p + 2 d e_a, so FOUR lattices - the nodes and the
     * x-, y-, z-half-offset lattices - replace the 7 displaced ones (4 network evaluations per point instead of 7).  All four
     * are stored with the padded dims (nx+1, ny+1, nz+1): xs4 [4][nx+1], ys4 [4][ny+1], zs4 [4][nz+1] (lattice l = 1, 2, 3
     * carries x - dx, y - dy, z - dz in its own direction, its last entry the + side of the last point), side4 / U4 / G4
     * [4][(nx+1)(ny+1)(nz+1)].  NULL = the 7-lattice path.  Needs p0, p1 multiples of ny*nz. */
    const float* xs4; const float* ys4; const float* zs4; const uint8_t* side4; float* U4; float* G4;
    /* Optional: the crossed sites that belong to this batch (indices into the level's crossed-site list, ascending).  The
     * cube kernels then walk these n_live sites instead of scanning all n_crossed and filtering (a 131072-point batch of a
     * 128^3 grid owns 1/16 of them).  NULL = scan. */
    const int32_t* c_live; int64_t n_live;
} nbm_points_step_t;

/* General path (any cell size, any contiguous batch): 7 network evaluations per point (the reference
 * does 197), no neighbour sharing: the node kernels of the shared path run over the 7 displaced lattices. */
int nbm_loss_grad_points_f32(const nbm_points_step_t* s, nbm_stream_t stream);

/* optax chain of solvers/optimizers.py:33-54 on device: clip_by_global_norm(max_norm) ->
 * scale_by_adam(b1,b2,eps) -> scale_by_schedule(lr*decay^(count/transition)) -> scale(-1) ->
 * apply_updates.  state = [m(P), v(P)], count = device int32 step counter.  grad_scale multiplies
 * the gradient first (1 normally).  loss_hist may be NULL, else loss_hist[count] = loss. */
typedef struct {
    int n_params;
    float lr, decay_rate, transition_steps, max_norm, b1, b2, eps;
    int optimizer;                   /* 0 = "custom" chain, 1 = optax.adam (no clip, constant lr),
                                        2 = optax.rmsprop (scale_by_rms(0.9, eps) -> -lr; b2 is the decay) */
    int scheduler;                   /* "custom" only: 0 = exponential_decay, 1 = polynomial(power 1, end 0) */
} nbm_optimizer_t;

int nbm_apply_update_f32(const nbm_optimizer_t* opt, const float* loss_grad, float* params,
                         float* state, int32_t* count, float* loss_hist, nbm_stream_t stream);

/* One kernel for the tail of an optimizer step (what the reference does in `optimizer.update` + `apply_updates` right
 * after `value_and_grad`, trainer.py:786-788): [sum the partial rows of nbm_loss_grad_shared_f32 run with
 * stages = everything but NBM_STAGE_REDUCE -> loss_grad] -> optax chain -> params, AND the refresh of the library's
 * staged copies of the network parameters (plain, pre-scaled, transposed), so that the next step starts with
 * nbm_upload_staged_params() (one device-to-device copy into the __constant__ bank) instead of nbm_upload_params().
 * partials == NULL: loss_grad is taken as given (after an all-reduce, or from nbm_loss_grad_points_f32).
 * opt->n_params may exceed the network's parameter count (learned preconditioner at the tail). */
int nbm_finalize_step_f32(const nbm_optimizer_t* opt, const nbm_net_t* net, const float* partials, int rows,
                          int row_stride, float* loss_grad, float* params, float* state, int32_t* count,
                          float* loss_hist, nbm_stream_t stream);
/* __constant__ bank <- the staged copies written by the last nbm_upload_params / nbm_finalize_step_f32 on this device */
int nbm_upload_staged_params(nbm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K4 (multi-GPU): partial-row reduction FUSED with the gradient all-reduce over NVLink peer memory.
 * Replaces jax.lax.psum(grads) / psum(loss) (trainer.py:829-830) = SUM over devices.
 *
 * Every rank owns one small communication block (allocated by the library with cudaMalloc because it
 * must be exportable through CUDA IPC); the blocks of all ranks of the node are mapped into every
 * process.  One single-CTA kernel per rank: sum the per-CTA partial rows -> PUSH [grad, loss] into this rank's
 * slot of EVERY rank's block over NVLink (slot = step parity; remote stores do not wait) -> fence -> raise this
 * rank's step flag in every block -> poll the LOCAL flags of all peers -> add the local slots in rank order
 * (bitwise identical result on every rank) -> out.
 * A wait longer than the timeout (nbm_comm_set_timeout, default 30 s, 0 = unbounded like NCCL) sets the block's error
 * word and fills THAT rank's output with NaN; the host must poll nbm_comm_error() (per checkpoint / at the end of
 * training) and raise: a timed-out rank never continues with a plausible-looking partial sum.
 * ---------------------------------------------------------------------------------------- */
#define NBM_IPC_HANDLE_BYTES 64
#define NBM_COMM_MAX_RANKS 8
/* process-wide wait bound of the exchange kernel in seconds (0 = wait for ever) */
int nbm_comm_set_timeout(double seconds);
/* single-process multi-device (the reference's pmap model, trainer.py:727-743): a block on the current device without
 * an IPC handle, and peer access `device` -> `peer` so that kernels on `device` can store into `peer`'s block */
int nbm_comm_alloc_local(void** local_block);
int nbm_enable_peer_access(int device, int peer);
/* allocate + zero the local block on the current device, export its IPC handle */
int nbm_comm_alloc(void** local_block, unsigned char handle[NBM_IPC_HANDLE_BYTES]);
/* map a peer's block (handle obtained from that peer's nbm_comm_alloc) */
int nbm_comm_open_peer(const unsigned char handle[NBM_IPC_HANDLE_BYTES], void** peer_block);
int nbm_comm_close_peer(void* peer_block);
int nbm_comm_free(void* local_block);
/* blocks_host[world]: device pointers of all ranks' blocks as seen from THIS process (own block at [rank]).
 * partials[rows][np1]; step_dev: device int32 step counter of this rank (starts at 0, incremented here);
 * out[np1] receives the sum over ranks.  np1 <= 1024. */
int nbm_reduce_allreduce_f32(const float* partials, int rows, int np1, int rank, int world,
                             void* const* blocks_host, int32_t* step_dev, float* out, nbm_stream_t stream);
/* The whole tail of a multi-GPU optimizer step as ONE kernel (update_multi_gpu, trainer.py:824-834: psum of grads
 * and loss, optimizer.update, apply_updates): nbm_reduce_allreduce_f32 followed by the tail of nbm_finalize_step_f32
 * (optax chain, staged parameter copies for the next step) without a kernel boundary between them.  loss_grad[np1]
 * receives the sum over ranks; np1 = opt->n_params + 1. */
int nbm_reduce_allreduce_finalize_f32(const nbm_optimizer_t* opt, const nbm_net_t* net, const float* partials, int rows,
                                      int np1, int rank, int world, void* const* blocks_host, int32_t* step_dev,
                                      float* loss_grad, float* params, float* state, int32_t* count, float* loss_hist,
                                      nbm_stream_t stream);
/* nonzero once a peer wait timed out on this device's block (host read of the block's error word) */
int nbm_comm_error(void* local_block);

/* ------------------------------------------------------------------------------------------
 * K5: post-training evaluation (trainer.py:960-977): u, grad u (analytic Jacobian of the
 * selected head), d u / d n with the central-difference normal (discretization.py:199-218).
 * pts (n,3); outputs u[n], grad_u[n*3], grad_n[n]  (grad_u / grad_n may be NULL)
 * ---------------------------------------------------------------------------------------- */
int nbm_evaluate_f32(const nbm_net_t* net, const nbm_lvl_t* lvl, const float* pts, int64_t n,
                     float dx, float dy, float dz, float* u, float* grad_u, float* grad_n,
                     nbm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NBM_B200_H */
