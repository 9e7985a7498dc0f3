/*
 * asg_cuda.h -- C ABI of libasg_cuda.so, the B200 (sm_100a) engine for AdaptiveSG.jl's hot path:
 * batched evaluation of the hierarchical hat-basis interpolant
 *
 *      G(x) = sum_n c_n * prod_d phi(x01_d; l_nd, i_nd)
 *
 * The reference (pure Julia) has no FFI for this path; its boundary is the Julia method table.
 * Each entry point below names the reference code (file:line in the AdaptiveSG.jl checkout) whose
 * job it takes over.  A Julia package reaches these with plain `ccall` (see INTEGRATION.md and
 * adaptivesg.jl_b200/julia/AdaptiveSGCUDA.jl); the Python mirror uses ctypes.
 *
 * Conventions
 *  - every function returns an int32 status: 0 = ASG_OK, negative = ASG_E*; nothing throws across
 *    the boundary; asg_last_error() returns a thread-local message for the last failure;
 *  - no host pointer is retained after a call returns;
 *  - matrices are addressed with explicit ELEMENT strides so that Julia's column-major
 *    (row stride 1, column stride N) and C/numpy row-major layouts are both accepted without a copy;
 *  - node arrays are Int64 like the reference's Matrix{Int} (src/class.jl:288-290);
 *  - there is NO CPU fallback: without an sm_100 device every compute entry point fails with
 *    ASG_ENODEVICE.
 *  - calls on one handle are serialised by a per-handle mutex; different handles may be used
 *    concurrently from different threads.
 */
#ifndef ASG_CUDA_H
#define ASG_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASG_OK 0
#define ASG_EINVAL (-1)        /* bad argument (null pointer, negative size, dimension mismatch)   */
#define ASG_EINVALID_NODE (-2) /* (l,i) pair the reference rejects with ArgumentError, math.jl:355 */
#define ASG_ENODEVICE (-3)     /* no usable sm_100 CUDA device; there is no CPU fallback           */
#define ASG_ECUDA (-4)         /* CUDA runtime error, see asg_last_error()                         */
#define ASG_ENOMEM (-5)        /* host or device allocation failed                                 */
#define ASG_EUNSUPPORTED (-6)  /* valid request outside the engine's limits (e.g. D > 64)          */
#define ASG_ESTATE (-7)        /* call order error (e.g. evaluate before upload)                   */

#define ASG_MAX_DIM 64        /* hard limit of every path                                         */
#define ASG_MAX_DIM_FAST 12   /* subspace-walk kernels are instantiated for D = 1..12              */

/* evaluation engines; ASG_PATH_AUTO picks SUBSPACE when the grid is canonical, else SWEEP */
#define ASG_PATH_AUTO 0
#define ASG_PATH_SWEEP 1     /* tiled all-nodes sweep, exact reference semantics for ANY node set  */
#define ASG_PATH_SUBSPACE 2  /* subspace walk, one supporting node per subspace (record-free on     */
                             /* regular sparse grids, level-vector trie otherwise)                 */
#define ASG_PATH_TRIE 3      /* subspace walk forced through the generic trie (tests: cross-check) */
#define ASG_PATH_STREAM 4    /* subspace walk forced through the flat group stream (tests)         */
#define ASG_PATH_BLOCK 5     /* subspace walk forced through the BLOCK layout; ASG_EUNSUPPORTED if */
                             /* the grid has none (asg_grid_info().block == 0)                     */

typedef struct asg_grid asg_grid;   /* opaque: device mirror of SparseGridInterpolant{D}, class.jl:281-319 */
typedef struct asg_event asg_event; /* opaque: completion handle of an *_async call                        */

typedef struct asg_grid_info_t {
  int32_t D;
  int32_t canonical;       /* 1: valid, in-range, duplicate-free nodes with finite coefs            */
  int32_t path;            /* engine the next asg_eval will use (ASG_PATH_SWEEP / _SUBSPACE)        */
  int32_t max_level;       /* max_n,d l_nd                                                          */
  int64_t N;               /* nodes                                                                 */
  int64_t max_depth;       /* max_n depth_n, depth = sum(l) - D + 1  (math.jl:103-108)              */
  int64_t subspaces;       /* distinct level vectors (= non-zero terms per point on a regular grid) */
  int64_t dense_subspaces; /* stored as full blocks                                                 */
  int64_t trie_nodes;      /* nodes of the level-vector trie over all D levels                      */
  int64_t coef_slots;      /* device coefficient slots incl. zero padding                           */
  int64_t hash_slots;      /* open-addressing slots of sparsely filled subspaces                    */
  int64_t device_bytes;    /* bytes resident in HBM for this grid                                   */
  int64_t generation;      /* bumped by every upload / append / set_coefs                           */
  int64_t regular;         /* 1: node set is exactly a regular sparse grid (record-free fast walk)  */
  int64_t block;           /* 1: BLOCK layout present (compiled walk of the last three dimensions)  */
  int64_t block_slots;     /* its coefficient slots incl. zero padding                              */
  int64_t block_records;   /* its prefix records                                                    */
  int64_t block_tiles;     /* shared-memory tiles the walk streams per batch of points              */
  int64_t block_stack;     /* saved prefix products per point the walk keeps in shared memory       */
  int64_t block_pair;      /* 1: large ordered batches walk in PAIR mode (two neighbours of the cell-key order per   */
                           /* thread share their coefficient loads)                                                  */
  int64_t packs;           /* full host packs (validate + sort + pack + upload) this handle has done so far           */
  int64_t delta_nodes;     /* nodes appended since the last pack: evaluated on top of the packed layouts by the sweep */
                           /* kernel until the next pack folds them in (asg_grid_append)                              */
} asg_grid_info_t;

/* ---- context ------------------------------------------------------------------------------- */
/* Select the CUDA device this process uses (one process per GPU). Fails with ASG_ENODEVICE when
 * the device is absent or is not compute capability 10.x. Idempotent. */
int32_t asg_init(int32_t device);
/* Select SEVERAL devices of one box for this process (the reference's only parallel entry is
 * adapt!(...; parallel = true), src/train.jl:332-340): every grid handle then keeps one replica of its packed
 * layouts per device (packed once on the host, uploaded to each), and the host-buffer entry points -- asg_eval,
 * asg_surplus, asg_surplus_nodes -- split their points into contiguous ranges over the devices, one host thread and
 * one copy/compute pipeline per device, results written straight into the caller's buffer. No collective, no NCCL,
 * no torch. devices[0] is the primary device (basis, children, node coordinates run there). The device-pointer
 * entry points run on whichever device of the list owns the pointers. asg_init(d) == asg_init_devices(&d, 1). */
int32_t asg_init_devices(const int32_t *devices, int32_t ndev);
/* the device list of the context (ndev = 0 before asg_init*) */
int32_t asg_context_devices(int32_t *devices, int32_t capacity, int32_t *ndev);
int32_t asg_shutdown(void);
int32_t asg_device_count(int32_t *count);
const char *asg_last_error(void);
const char *asg_version(void);
/* number of CUDA kernels this library has launched since asg_init (bench.py: gpu_launches) */
int32_t asg_launch_count(int64_t *count);
/* pin / unpin a caller-owned host buffer so that H2D/D2H copies of asg_eval overlap with compute */
int32_t asg_host_register(void *ptr, size_t bytes);
int32_t asg_host_unregister(void *ptr);

/* ---- node store: replaces the fields of SparseGridInterpolant{D} (src/class.jl:281-319) ----- */
/* constructor asserts of src/class.jl:296-306: D > 0, finite bounds, ub > lb */
int32_t asg_grid_create(int32_t D, const double *lb, const double *ub, asg_grid **out);
int32_t asg_grid_destroy(asg_grid *g); /* safe from any thread (Julia finalizer) */

/* Full replace of the node set (what rsg! / deserialize do, src/train.jl:68-75, src/io.jl:64-69).
 * Element (n,d) of levels/indices is at [n*sn + d*sd]. coefs may be NULL (zeros, as rsg! leaves
 * them, src/train.jl:74-75). Every (l,i) is checked with the rule of phi (src/math.jl:345-357):
 * l < 1 or (l == 2 and i not in {0,2}) -> ASG_EINVALID_NODE. Depths are derived (math.jl:103-108). */
int32_t asg_grid_upload(asg_grid *g, int64_t N, const int64_t *levels, const int64_t *indices, int64_t sn,
                        int64_t sd, const double *coefs);
/* vcat of new rows (adapt!, src/train.jl:434-438). Rows that are deeper than every node the grid holds (what adapt!
 * appends: children of the deepest nodes), canonical and finite are NOT packed at once: they join the sweep arrays on
 * the device and every evaluation adds their sum to the packed layouts' result (same nodes, same products; only the
 * order of the final additions differs). The next full pack -- when they exceed half of the packed nodes (8192 at
 * least), before a batch large enough for the extra sweep to cost more than packing, before basis / export -- folds
 * them in. An adapt! loop therefore packs once, not once per level. Any other rows trigger a full pack at once. */
int32_t asg_grid_append(asg_grid *g, int64_t n, const int64_t *levels, const int64_t *indices, int64_t sn,
                        int64_t sd, const double *coefs);
/* G.coefs[first .. first+n-1] = coefs  (0-based node numbers; train!, src/train.jl:183) */
int32_t asg_grid_set_coefs(asg_grid *g, int64_t first, int64_t n, const double *coefs);
/* G.coefs[idx[k]] = coefs[k] */
int32_t asg_grid_scatter_coefs(asg_grid *g, int64_t n, const int64_t *idx, const double *coefs);
/* read back G.coefs (coefficients(G), src/class.jl:414-416) */
int32_t asg_grid_get_coefs(asg_grid *g, int64_t first, int64_t n, double *coefs);
int32_t asg_grid_info(asg_grid *g, asg_grid_info_t *info);
int32_t asg_grid_set_path(asg_grid *g, int32_t path); /* force an engine (tests); AUTO restores */

/* ---- the basis function itself: phi(x), phi(x, l, i), phi(x, levels, indices) (src/math.jl:332-334, 345-357,
 *      376-391; exported by the reference, math.jl:13) ------------------------------------------------------------
 * out[k] = prod_{d < D} phi(x01[k*D + d]; levels[k*D + d], indices[k*D + d]) for k < n, a left fold like Julia's
 * prod; x01 in UNIT-CUBE coordinates as in the reference (no scale()). D = 1 is the scalar form. levels == indices
 * == NULL: the plain hat phi(x) = 1 - |x| on [-1, 1]. Same operations in the same order as the reference (mul, sub,
 * 1 - |t|, never fused): bit-identical values. ASG_EINVALID_NODE where the reference throws (math.jl:355). */
int32_t asg_phi(int64_t n, int32_t D, const double *x01, const int64_t *levels, const int64_t *indices, double *out);

/* ---- evaluation: _interpolate / _interpolate_if / _interpolate_until_level / G(x)
 *      (src/class.jl:629-643, 645-664, 666-680, 747-761) ----------------------------------- */
/* out[m] = G(X[m,:]) for m < M, X in DOMAIN units (scale(), src/math.jl:301-309, is applied on the
 * device); element (m,d) of X at [m*sp + d*sdim]. max_depth < 0: all nodes; otherwise only nodes
 * with depth <= max_depth (the mask of _interpolate_until_level). Host buffers. */
int32_t asg_eval(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, int64_t max_depth,
                 double *out);
/* alpha[m] = fvals[m] - G_{depth<=max_depth}(X[m,:]): the residual step of train! / adapt!
 * (src/train.jl:176-183, 362-363, 381-382) fused into the evaluation kernel's epilogue */
int32_t asg_surplus(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, const double *fvals,
                    int64_t max_depth, double *alpha);
/* Same, but the query points are the coordinates of the grid's own nodes idx[0..n-1], generated on
 * the device exactly like get_x -> scale -> scale (src/train.jl:170-180). store != 0 additionally
 * writes alpha into the coefficients of those nodes (G.coefs[i] = alpha, src/train.jl:183). */
int32_t asg_surplus_nodes(asg_grid *g, int64_t n, const int64_t *idx, const double *fvals, int64_t max_depth,
                          int32_t store, double *alpha);
/* same two entry points with DEVICE pointers (points already resident in HBM), enqueued on
 * `stream` (a cudaStream_t; NULL = the CUDA default stream) without host synchronisation */
int32_t asg_eval_device(asg_grid *g, const double *dX, int64_t M, int64_t sp, int64_t sdim, int64_t max_depth,
                        double *dout, void *stream);
int32_t asg_surplus_device(asg_grid *g, const double *dX, int64_t M, int64_t sp, int64_t sdim,
                           const double *dfvals, int64_t max_depth, double *dalpha, void *stream);
/* non-blocking host-buffer evaluation so a Julia thread is not pinned in a long ccall; the caller
 * must keep X/out alive (GC.@preserve) until asg_wait returns; asg_wait frees the event */
int32_t asg_eval_async(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, int64_t max_depth,
                       double *out, asg_event **ev);
int32_t asg_wait(asg_event *ev);

/* ---- node coordinates: get_x / stack(G) (src/math.jl:464-497, src/class.jl:375-387) --------- */
/* Xout[(k), d] = lb_d + (ub_d - lb_d) * x01(l,i), two roundings, no FMA; nodes first..first+n-1 */
int32_t asg_nodes_x(asg_grid *g, int64_t first, int64_t n, double *Xout, int64_t sp, int64_t sdim);
/* coordinates of arbitrary (l,i) rows in g's domain (adapt! children, src/train.jl:361, 380);
 * rows get_x would reject (src/math.jl:474) -> ASG_EINVALID_NODE */
int32_t asg_nodes_x_li(asg_grid *g, int64_t n, const int64_t *levels, const int64_t *indices, int64_t sn,
                       int64_t sd, double *Xout, int64_t sp, int64_t sdim);

/* ---- device-format persistence (interchange format stays the Dict of src/io.jl:30-38) ------------
 * asg_grid_export writes the node arrays and every packed layout of the uploaded grid (current coefficients) as one
 * blob: call with buf = NULL for the size, then with a buffer. asg_grid_import makes a blob the state of the handle
 * like asg_grid_upload does, without validating, sorting or packing again; the handle must have the blob's D and
 * domain. ASG_EUNSUPPORTED: written by a build with other layout constants (fall back to asg_grid_upload);
 * ASG_EINVAL: not a blob / checksum mismatch / other dimension or domain. */
int32_t asg_grid_export(asg_grid *g, void *buf, int64_t capacity, int64_t *bytes);
int32_t asg_grid_import(asg_grid *g, const void *buf, int64_t bytes);
/* deserialize (src/io.jl:49-72) with a blob next to the Dict: install the blob only if it describes exactly the node
 * arrays of the Dict (levels, indices bit for bit; coefs bit for bit when given) -- serialize aliases the arrays, so an
 * in-place update after the blob was written leaves it stale. ASG_ESTATE on mismatch: fall back to asg_grid_upload. */
int32_t asg_grid_import_checked(asg_grid *g, const void *buf, int64_t bytes, int64_t N, const int64_t *levels,
                                const int64_t *indices, int64_t sn, int64_t sd, const double *coefs);

/* ---- rsg! node enumeration (src/train.jl:40-78), host code, no device needed: the regular sparse grid node set
 * {l : l_d <= maxlevels_d, sum(l_d - 1) <= maxdepth - 1} in the reference's order (level vectors and index tuples in
 * Iterators.product order, dimension 1 fastest). Call _count, allocate N x D, call _fill (strides in elements). */
int32_t asg_rsg_count(int32_t D, int64_t maxdepth, const int64_t *maxlevels, int64_t *N);
int32_t asg_rsg_fill(int32_t D, int64_t maxdepth, const int64_t *maxlevels, int64_t N, int64_t *levels, int64_t *indices,
                     int64_t sn, int64_t sd);

/* ---- adapt! candidate generation (src/train.jl:337-398; child: src/math.jl:907-959; isvalid: 418-448) --------
 * asg_children builds, on the device, the valid children of every node of depth `depth` in the reference's visiting
 * order (for j in 1:D, for parent in node order, left then right) and removes the repeats (a child is reached from
 * up to D parents; the first one generated stays - the reference's `!haskey` rule). The reference evaluates f and G
 * at every repeat before its Dict drops them; generating the unique set first gives the same nodes with up to D
 * times fewer calls of f. n_candidates = 2 * D * n_parents. asg_children_fetch copies the n_unique rows (strides in
 * elements, as everywhere) and, if X != NULL, their coordinates (get_x, domain units). */
int32_t asg_children(asg_grid *g, int64_t depth, int64_t *n_parents, int64_t *n_candidates, int64_t *n_unique);
int32_t asg_children_fetch(asg_grid *g, int64_t n, int64_t *levels, int64_t *indices, int64_t sn, int64_t sd, double *X,
                           int64_t sp, int64_t sdim);

/* ---- basis rows: basis(x,G) / basis(Xs,G) / basis(G) / dehierarchization_matrix
 *      (src/class.jl:434-454, 483-503, 521-538, 555-567; documented intent) ----- */
/* pass 1: row_nnz[m] = number of stored entries of row m. Entry (m,n) = prod_d phi(...) is stored
 * when it is != 0 and (dropzeros == 0 or value >= atol)  (src/class.jl:450-453). */
int32_t asg_basis_count(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, double atol,
                        int32_t dropzeros, int64_t *row_nnz);
/* pass 2: CSR fill. row_ptr has M+1 entries (exclusive scan of row_nnz); colidx are 0-based node
 * numbers in the caller's node order, ascending inside a row. */
int32_t asg_basis_fill(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, double atol,
                       int32_t dropzeros, const int64_t *row_ptr, int64_t *colidx, double *vals);
/* dense variant: out[m*ldm + n*ldn], every element written (zeros included) */
/* the same entries as asg_basis_fill, transposed on the device into compressed sparse COLUMNS - the SparseMatrixCSC the
 * reference returns (src/class.jl:483-503): colptr[N+1], rowidx / vals [row_ptr[M]], 0-based, rows ascending in a column */
int32_t asg_basis_fill_csc(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, double atol,
                           int32_t dropzeros, const int64_t *row_ptr, int64_t *colptr, int64_t *rowidx, double *vals);
/* the same with 32-bit colptr / rowidx (sparse types that hold Int32 indices, e.g. scipy's default: half the index
 * bytes come home and the caller converts nothing); ASG_EUNSUPPORTED when N or the entry count needs 64 bits */
int32_t asg_basis_fill_csc32(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, double atol,
                             int32_t dropzeros, const int64_t *row_ptr, int32_t *colptr, int32_t *rowidx, double *vals);
int32_t asg_basis_dense(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, double atol,
                        int32_t dropzeros, double *out, int64_t ldm, int64_t ldn);
/* dense variant into a DEVICE buffer (bench: HBM-write roofline without the D2H copy) */
int32_t asg_basis_dense_device(asg_grid *g, const double *dX, int64_t M, int64_t sp, int64_t sdim, double atol,
                               int32_t dropzeros, double *dout, int64_t ldm, int64_t ldn, void *stream);

/* ---- host-only diagnostics -------------------------------------------------------------------- */
/* Run the packer (validation, subspace grouping, trie build) on a node set WITHOUT a device and
 * report the sizes it would upload. No evaluation happens here. When the grid is not canonical,
 * asg_last_error() says why. */
int32_t asg_pack_info(int32_t D, int64_t N, const int64_t *levels, const int64_t *indices, int64_t sn, int64_t sd,
                      asg_grid_info_t *info);

/* ---- measurement helpers (bench.py) ---------------------------------------------------------- */
/* dependent-free DFMA loop on every SM: returns the measured FP64 rate in FMA lane-ops per second
 * (multiply by 2 for FLOP/s); the denominator of the FP64 roofline (MEASURED_PEAKS.json has none) */
int32_t asg_measure_fp64_peak(double *fma_per_second);
/* time (ms, CUDA events on the library stream) of the last asg_eval_device / asg_surplus_device /
 * asg_basis_dense_device kernel sequence */
int32_t asg_last_kernel_ms(double *ms);
/* the same for the evaluation (walk) kernel alone: equal to asg_last_kernel_ms unless the call ordered its
 * points by cell key first (BLOCK walk on large batches, csrc/asg_sort.cu) */
int32_t asg_last_walk_ms(double *ms);

#ifdef __cplusplus
}
#endif
#endif /* ASG_CUDA_H */
