/*
 * zfista_b200 -- C ABI of the B200-native proximal-gradient hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / numpy types.
 * Every entry point names the reference interface it replaces (file:line in
 * zalgo3/zfista).  The reference is pure Python, so "the FFI a maintainer would
 * bind" is ctypes; INTEGRATION.md shows that stub.
 *
 * Conventions
 *   - all floating point data is IEEE fp64, row-major, contiguous;
 *   - "_device" entry points take device pointers and a CUDA stream and are
 *     asynchronous; "_host" entry points take host pointers, do their own
 *     H2D / D2H copies and return after the stream is synchronised;
 *   - every function returns 0 on success, a negative zf_status otherwise;
 *     zf_last_error() gives the message of the last failure on this thread;
 *   - there is NO CPU fallback: without a CUDA device the calls fail with
 *     ZF_ERR_CUDA.
 */
#ifndef ZFISTA_B200_H
#define ZFISTA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZF_ABI_VERSION 1
#define ZF_MAX_OBJECTIVES 4

typedef enum {
  ZF_OK = 0,
  ZF_ERR_INVALID = -1,     /* bad argument (shape, option, unsupported size)   */
  ZF_ERR_CUDA = -2,        /* CUDA runtime error / no device                   */
  ZF_ERR_UNSUPPORTED = -3  /* valid in the reference, not built on device yet  */
} zf_status;

/* Built-in problem classes: zfista/problems.py:153 (JOS1), :208 (SD), :267 (FDS),
 * :331 (ZDT1), :389 (TOI4), :451 (TRIDIA), :517 (LinearFunctionRank1).
 * ZF_LSQ_L1 is  scale*||A x - b||^2 + l1*||x||_1 , the single-objective closures
 * of tests/test_proximal_gradient.py:49-63 and examples/cameraman.ipynb. */
typedef enum {
  ZF_JOS1 = 0,
  ZF_SD = 1,
  ZF_FDS = 2,
  ZF_ZDT1 = 3,
  ZF_TOI4 = 4,
  ZF_TRIDIA = 5,
  ZF_LFR1 = 6,
  ZF_LSQ_L1 = 7
} zf_problem_kind;

/* Problem descriptor == constructor arguments of zfista.problems.Problem
 * (problems.py:63-79): n_features, n_objectives, l1_ratios, l1_shifts, bounds. */
typedef struct {
  int32_t kind;            /* zf_problem_kind */
  int32_t n_features;
  int32_t n_objectives;    /* 1..ZF_MAX_OBJECTIVES */
  int32_t has_l1;          /* l1_ratios is not None */
  double l1_ratios[ZF_MAX_OBJECTIVES];
  double l1_shifts[ZF_MAX_OBJECTIVES];
  int32_t has_bounds;      /* bounds is not None */
  int32_t bounds_are_arrays; /* 0: scalars lower/upper; 1: lower_v/upper_v (n_features) */
  double lower, upper;
  const double* lower_v;   /* device pointers for *_device calls, host for *_host */
  const double* upper_v;
  /* ZF_LSQ_L1 only (small dense instance solved inside the batched kernel) */
  const double* A;         /* n_rows x n_features */
  const double* b;         /* n_rows */
  int32_t n_rows;
  int32_t reserved;
  double scale;
  double l1;
} zf_problem;

/* Keyword arguments of minimize_proximal_gradient (proximal_gradient.py:311-331). */
typedef struct {
  double lr;                 /* initial step                       (default 1)        */
  double tol;                /* stop when max|x^k - y^k| < tol     (default 1e-5)     */
  double tol_internal;       /* inner solver / line-search tol     (default 1e-12)    */
  int64_t max_iter;          /*                                     (default 1000000) */
  int32_t max_iter_internal; /*                                     (default 100000)  */
  int32_t max_backtrack_iter;/*                                     (default 100)     */
  int32_t warm_start;        /* reuse dual weights as next initial guess              */
  int32_t nesterov;          /* FISTA if nonzero, ISTA otherwise                      */
  double decay_rate;         /* backtracking factor; 1 disables the line search       */
  double nesterov_a;         /* t_{k+1} = sqrt(t_k^2 - a t_k + b) + 1/2               */
  double nesterov_b;
  int32_t deprecated;        /* subproblem without f_i(y^k) - F_i(x^{k-1})            */
  int32_t dual_solver;       /* 0: reference-faithful (bounded Brent for m = 2,
                                   simplex Newton for m >= 3); 1: simplex Newton for
                                   every m >= 2 (faster, exact optimum)               */
  int32_t trace_capacity;    /* return_all: iterations recorded per start (0 = off)   */
  int32_t reserved;
} zf_options;

/* Per-start outputs == OptimizeResult fields (proximal_gradient.py:544-555). */
typedef struct {
  double* x;        /* n_starts x n_features                                   */
  double* fun;      /* n_starts x n_objectives  F(x) = f(x) + g(x)             */
  int64_t* nit;     /* n_starts                                                */
  int32_t* status;  /* n_starts: 1 converged, 0 max_iter reached, -1 backtracking failed */
  double* lr;       /* n_starts: final step size (may be NULL)                 */
  int64_t* nfev;    /* n_starts: evaluations of f on device (may be NULL)      */
  int64_t* n_dual;  /* n_starts: dual-function evaluations (may be NULL)       */
  double* err;      /* n_starts: last max|x^k - y^k| (may be NULL)             */
  /* return_all traces, each may be NULL; capacity = options.trace_capacity    */
  double* allerrs;  /* n_starts x cap                                          */
  double* allfuns;  /* n_starts x (cap + 1) x n_objectives, entry 0 = F(x0)    */
  double* allvecs;  /* n_starts x (cap + 1) x n_features,   entry 0 = x0       */
  /* RAGGED traces (NULL: the dense layout above).  n_starts + 1 offsets, off[0] = 0: start s
   * records at most off[s+1] - off[s] iterations; its errors start at allerrs[off[s]], its
   * F values at allfuns[(off[s] + s) * n_objectives], its iterates at
   * allvecs[(off[s] + s) * n_features] (one more entry than errors: entry 0 = x0).  With the
   * iteration counts of a first, trace-less solve as capacities nothing is allocated that is
   * not written (benchmarks/benchmark.py:320-372 asks return_all for every start).           */
  const int64_t* trace_offsets;
} zf_result;

int zf_abi_version(void);
const char* zf_last_error(void);
/* number of CUDA devices visible; <= 0 means the product path cannot run */
int zf_device_count(void);
void zf_default_options(zf_options* opt);

/* ---- (a) batched FISTA / ISTA: whole loop on device, one warp per start ----
 * replaces  Problem.minimize_proximal_gradient(x0, **kw)  called once per start
 * under joblib in benchmarks/benchmark.py:320-372 and in
 * examples/PGM_experiment_with_various_a_b.ipynb (run()).
 * x0: n_starts x n_features.  ab: optional n_starts x 2 per-start (a, b) momentum
 * pairs (NULL = options.nesterov_a/b for every start).                          */
int zf_solve_batched_device(const zf_problem* problem, const zf_options* opt,
                            int64_t n_starts, const double* d_x0, const double* d_ab,
                            const zf_result* d_out, void* cuda_stream);
int zf_solve_batched_host(const zf_problem* problem, const zf_options* opt,
                          int64_t n_starts, const double* h_x0, const double* h_ab,
                          const zf_result* h_out);

/* ---- (b) one proximal subproblem per row (the multi-objective dual) --------
 * replaces _solve_subproblem (proximal_gradient.py:35-209).
 * y, x_old: n x n_features; lr: n; deprecated: n (0/1) or NULL.
 * outputs: x (n x n_features), fun (n), weight (n x n_objectives).              */
int zf_solve_subproblem_host(const zf_problem* problem, const zf_options* opt, int64_t n,
                             const double* h_y, const double* h_x_old, const double* h_lr,
                             const int32_t* h_deprecated, double* h_x, double* h_fun,
                             double* h_weight);

/* ---- device functors of zfista.problems evaluated at a batch of points -----
 * replaces Problem.f / .g / .jac_f / .prox_wsum_g (problems.py:93-138).
 * Any output pointer may be NULL.  X: n x n_features, W: n x n_objectives
 * (weights for prox_wsum_g).  f, g: n x m; jac: n x m x n_features; prox: n x nf */
int zf_problem_eval_host(const zf_problem* problem, int64_t n, const double* h_X,
                         const double* h_W, double* h_f, double* h_g, double* h_jac,
                         double* h_prox);

/* ---- (c) large-n single-objective LASSO:  scale*||Ax-b||^2 + l1*||x||_1 ----
 * replaces minimize_proximal_gradient(f, g, jac_f, prox_wsum_g, x0, ...) with
 * the dense closures of tests/test_proximal_gradient.py:49-63 at sizes where one
 * pass over A is HBM-bound.  A handle owns the device workspace; A itself stays
 * where the caller put it (device pointer, row-major, ld = n_cols).
 * Multi-GPU: each rank passes its row shard and an `exchange` flag; the gradient
 * and residual-norm partials are then left in `zf_lasso_partial()` for the
 * caller's NCCL all-reduce between zf_lasso_grad() and zf_lasso_step().       */
typedef struct zf_lasso zf_lasso;

int zf_lasso_create(zf_lasso** out, const double* d_A, const double* d_b, int64_t n_rows,
                    int64_t n_cols, double scale, double l1, void* cuda_stream);
void zf_lasso_destroy(zf_lasso* h);
/* whole solve on one GPU; x (n_cols) in/out on device.  Returns nit etc. in *res
 * (host pointers, n_starts = 1; x/fun fields are host pointers too).            */
int zf_lasso_solve(zf_lasso* h, const zf_options* opt, const double* d_x0, double* d_x,
                   double* h_fun, int64_t* h_nit, int32_t* h_status, double* h_allerrs,
                   double* h_allfuns);
/* split form for row-sharded multi-GPU runs */
int zf_lasso_begin(zf_lasso* h, const zf_options* opt, const double* d_x0);
int zf_lasso_grad(zf_lasso* h, int which /*0: at y (gradient + f), 1: f at x_new*/);
double* zf_lasso_partial(zf_lasso* h, int64_t* n_values); /* device buffer to all-reduce */
int zf_lasso_step(zf_lasso* h, int32_t* h_done);           /* prox, line search, momentum */
int zf_lasso_finish(zf_lasso* h, double* d_x, double* h_fun, int64_t* h_nit,
                    int32_t* h_status);
/* Device-decided form of the same loop (proximal_gradient.py:474-538): lr, t_k, the F values and
 * the line-search / stop decisions live in device memory and are taken by the kernels, so
 * nothing below waits for the GPU except zf_lasso_dev_poll(wait = 1) and zf_lasso_dev_finish().
 * zf_lasso_solve() is built on it (chunks of 32 trials as one CUDA graph launch, the host polling
 * one chunk behind).  A row-sharded run enqueues the stages of a trial one by one and all-reduces
 * zf_lasso_partial() between them, on the same stream, without reading anything back:
 *   dev_begin(sharded = 1); AR(partial[n_cols:]); stage(0);
 *   repeat { stage(1); AR(partial); stage(2);
 *            if dev_needs_feval: stage(3); AR(partial[n_cols:]); stage(4) }   -- poll now and then
 *   if !dev_needs_feval: stage(6); AR(partial[n_cols:]);
 *   stage(5); dev_finish()
 * Trials enqueued after the solve has ended do nothing.                                        */
int zf_lasso_set_stream(zf_lasso* h, void* cuda_stream);
int zf_lasso_dev_begin(zf_lasso* h, const zf_options* opt, const double* d_x0, int32_t sharded,
                       int32_t want_trace);
int zf_lasso_dev_stage(zf_lasso* h, int32_t stage);
int zf_lasso_dev_needs_feval(zf_lasso* h);
int zf_lasso_dev_slots(zf_lasso* h, int32_t n_slots);       /* single GPU: whole trials */
int zf_lasso_dev_poll(zf_lasso* h, int32_t slot /*0|1*/, int32_t wait, int32_t* h_done,
                      int64_t* h_nit);
int zf_lasso_dev_finish(zf_lasso* h, double* d_x, double* h_fun, int64_t* h_nit,
                        int32_t* h_status, double* h_lr, double* h_allerrs, double* h_allfuns);
/* return_all's allvecs for the NEXT solve with traces: rows x n_cols doubles on the host,
 * row k = x^k; rows 0 .. min(nit, rows - 1) are written (proximal_gradient.py:471, 522)      */
int zf_lasso_set_allvecs(zf_lasso* h, double* h_allvecs, int64_t rows);
/* Row-sharded runs on ONE node: the exchange folded into the kernels over NVLink peer memory
 * instead of an NCCL all-reduce between the stages.  Each rank exports a 64-byte
 * cudaIpcMemHandle_t of its exchange buffer, the caller all-gathers the handles over its control
 * plane and attaches them (<= 8 ranks).  From then on stage 1 publishes this rank's
 * [A^T r | sum r^2] into its own buffer and stage 2 (the prox kernel) sums every rank's partials
 * straight out of peer memory, in rank order -- bit-identical on every rank; stages 3 / 6
 * exchange the residual norm the same way.  The caller skips its all-reduces when
 * zf_lasso_p2p_active() returns 1.  A peer that never publishes ends the solve with status -3
 * after ~10 s instead of hanging the GPU.                                                  */
int zf_lasso_p2p_export(zf_lasso* h, void* handle_out /* 64 bytes */);
int zf_lasso_p2p_attach(zf_lasso* h, int32_t rank, int32_t world, const void* handles /* world x 64 */);
int zf_lasso_p2p_active(zf_lasso* h);
/* how many times one gradient evaluation reads A from HBM with this handle's kernel choice:
 * 1 (fused A^T(Av-b) kernels) or 2 (residual pass + A^T pass); for the roofline accounting */
int zf_lasso_passes(zf_lasso* h);
/* one gradient pass only (bench / roofline): grad = 2*scale*A^T(Ax-b), returns f */
int zf_lasso_gradient_device(zf_lasso* h, const double* d_x, double* d_grad, double* d_f);

/* ---- (c'') many LASSO runs sharing one A: FP64 tensor-core DGEMM passes ----------------
 * replaces the joblib fan-out of minimize_proximal_gradient(f, g, jac_f, prox_wsum_g, x0_k,
 * nesterov_ratio=(a_k, b_k)) over runs k = 0..n_runs-1 that share the dense A of
 * tests/test_proximal_gradient.py:49-63 (examples/PGM_experiment_with_various_a_b.ipynb run(),
 * examples/cameraman.ipynb): run k minimises scale*||A x - b_k||^2 + l1*||x||_1.
 * n_runs <= 32; b is one vector (b_is_batched = 0) or n_runs x n_rows; x0 one vector or
 * n_runs x n_cols; ab n_runs x 2 on the host or NULL (options.nesterov_a/b for every run).
 * One gradient of ALL runs = two passes over A (R = A V - B, G = A^T R) on the FP64 tensor
 * cores, i.e. 2/n_runs HBM passes per run.  n_cols must be even and A 16-byte aligned
 * (ZF_ERR_UNSUPPORTED otherwise).  Outputs: x (n_runs x n_cols, device), fun / nit / status /
 * lr / err (n_runs, host), allerrs (n_runs x cap), allfuns (n_runs x (cap+1)) on the host.
 * The split form mirrors zf_lasso_begin/grad/step/finish; zf_lasso_multi_partial() is
 * [A^T R | sum r^2] of every run, the buffer a row-sharded run all-reduces.            */
typedef struct zf_lasso_multi zf_lasso_multi;

int zf_lasso_multi_create(zf_lasso_multi** out, const double* d_A, int64_t n_rows,
                          int64_t n_cols, const double* d_b, int32_t b_is_batched,
                          int32_t n_runs, double scale, double l1, void* cuda_stream);
void zf_lasso_multi_destroy(zf_lasso_multi* h);
int zf_lasso_multi_set_stream(zf_lasso_multi* h, void* cuda_stream);
int zf_lasso_multi_solve(zf_lasso_multi* h, const zf_options* opt, const double* d_x0,
                         int32_t x0_is_batched, const double* h_ab, double* d_x, double* h_fun,
                         int64_t* h_nit, int32_t* h_status, double* h_lr, double* h_err,
                         double* h_allerrs, double* h_allfuns);
int zf_lasso_multi_begin(zf_lasso_multi* h, const zf_options* opt, const double* d_x0,
                         int32_t x0_is_batched, const double* h_ab);
int zf_lasso_multi_grad(zf_lasso_multi* h, int which /*0: gradients at y, 1: f at the candidates*/);
double* zf_lasso_multi_partial(zf_lasso_multi* h, int64_t* n_values);
int zf_lasso_multi_step(zf_lasso_multi* h, int32_t* h_next);
int zf_lasso_multi_finish(zf_lasso_multi* h, double* d_x, double* h_fun, int64_t* h_nit,
                          int32_t* h_status, double* h_lr, double* h_err);
/* The device-decided rounds in pieces, for a caller that owns an exchange between the stages
 * (row-sharded runs; the protocol and the stage numbers of zf_lasso_dev_* above): begin, all-reduce
 * the residual norms (the last kp values of partial, zf_lasso_multi_layout), stage 0; per trial
 * stage 1, all-reduce partial, stage 2 and -- if zf_lasso_multi_dev_needs_feval() -- stage 3,
 * all-reduce the norms, stage 4; poll(slot, wait=0) enqueues a snapshot of the state,
 * poll(slot, wait=1, &done) waits for it; after done: (stage 6, all-reduce the norms, if no
 * feval), stage 5, finish.  Replaces the per-trial host decisions of proximal_gradient.py:474-538
 * for every run of the lockstep batch.                                                    */
int zf_lasso_multi_dev_begin(zf_lasso_multi* h, const zf_options* opt, const double* d_x0,
                             int32_t x0_is_batched, const double* h_ab);
int zf_lasso_multi_dev_stage(zf_lasso_multi* h, int32_t stage);
int zf_lasso_multi_dev_needs_feval(zf_lasso_multi* h);
int zf_lasso_multi_dev_poll(zf_lasso_multi* h, int32_t slot, int32_t wait, int32_t* done);
int zf_lasso_multi_dev_finish(zf_lasso_multi* h, double* d_x, double* h_fun, int64_t* h_nit,
                              int32_t* h_status, double* h_lr, double* h_err);
int zf_lasso_multi_layout(zf_lasso_multi* h, int64_t* kp, int64_t* pitch_c);
/* the closures at n_runs points: grad (n_runs x n_cols) = 2*scale*A^T(A x_k - b_k), f (n_runs) */
int zf_lasso_multi_gradient_device(zf_lasso_multi* h, const double* d_X, double* d_grad,
                                   double* d_f);
/* one DGEMM pass alone (bench / roofline): which = 0  R = A X - B,  1  A^T R of the last R */
int zf_lasso_multi_pass_device(zf_lasso_multi* h, const double* d_X, int which);

/* ---- (c') cameraman-style deblurring:  ||R W x - b||^2 + l1*||x||_1 ----------------
 * replaces minimize_proximal_gradient(f, g, jac_f, prox_wsum_g, x0, ...) with the closures of
 * examples/cameraman.ipynb ("Objective function" cell): R = correlate2d(., kernel,
 * mode="same", boundary="symm"), W = inverse single-level 2-D Haar transform of the
 * coefficient vector [cA, cH, cV, cD] (each height/2 x width/2).  The notebook's joblib
 * fan-out over (a, b) momentum pairs becomes n_runs runs of one call: x0 is one vector
 * (x0_is_batched = 0) or n_runs x (height*width); ab is n_runs x 2 (host) or NULL.
 * Results: x (n_runs x n), fun / nit / status / lr / err (n_runs), allerrs (n_runs x cap),
 * allfuns (n_runs x (cap+1)); allvecs is not recorded.                                   */
typedef struct zf_deblur zf_deblur;

int zf_deblur_create(zf_deblur** out, int32_t height, int32_t width, const double* h_kernel,
                     int32_t ksize, const double* h_observed, double l1, int32_t max_runs,
                     void* cuda_stream);
void zf_deblur_destroy(zf_deblur* h);
int zf_deblur_solve_host(zf_deblur* h, const zf_options* opt, int64_t n_runs,
                         const double* h_x0, int32_t x0_is_batched, const double* h_ab,
                         const zf_result* h_out);
int zf_deblur_solve_device(zf_deblur* h, const zf_options* opt, int64_t n_runs,
                           const double* d_x0, int32_t x0_is_batched, const double* h_ab,
                           const zf_result* d_out);
/* the closures themselves at n_points <= max_runs points: f, g (n_points), jac (n_points x n) */
int zf_deblur_eval_host(zf_deblur* h, int64_t n_points, const double* h_X, double* h_f,
                        double* h_g, double* h_jac);

#ifdef __cplusplus
}
#endif
#endif /* ZFISTA_B200_H */
