/* libsympgpr_b200 -- C ABI of the B200-native SympGPR hot path.
 *
 * Drop-in boundary: this library takes the place of the reference's f2py extension modules
 * (built by python/05_tokamak/SympGPR/make_sympgpr.mk and python/<example>/Makefile):
 *     sympgpr.sympgpr / fortran.sympgpr.sympgpr   python/05_tokamak/SympGPR/sympgpr.f90
 *     kernels / kernels_sq / kernels_sum          python/<example>/kernels*.f90
 *     fieldlines.fieldlines (compute_r, ath)      python/05_tokamak/SympGPR/fieldlines.f90
 * It is loaded with ctypes by sympgpr_b200/_lib.py; INTEGRATION.md shows the reference-side
 * binding.  Plain pointers and sizes only; all matrices are column-major (Fortran order) as the
 * f2py signatures require; all reals are IEEE double.
 *
 * Return value of every int function: 0 ok; > 0 numerical failure (Cholesky: 1-based index of
 * the first non-positive pivot -- the Python shim raises numpy.linalg.LinAlgError so the
 * scripts' own try/except fallbacks fire, python/02_pert_pendulum/func.py:194-204);
 * < 0 SGP_E_* argument / runtime error, text via sgp_last_error().
 * There is no CPU fallback: without a CUDA device every compute entry point returns SGP_E_NODEV.
 *
 * "host" entry points take host pointers and are synchronous on return.  "_dev" entry points
 * take device pointers, enqueue on the context's stream and return without synchronising.
 */
#ifndef SYMPGPR_B200_H
#define SYMPGPR_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define SGP_E_BADARG (-1)
#define SGP_E_CUDA   (-2)
#define SGP_E_NOMEM  (-3)
#define SGP_E_NODEV  (-4)

/* kernel families (argument `fam`); `per` is the period parameter p of the periodic factor
 * exp(-sin(p dx)^2 / (2 lx^2)): 0.5 for the committed product kernel (kernels.f90), free for
 * python/01_pendulum/implicit_period_unknown/kernels.f90; ignored by SGP_FAM_SQ. */
#define SGP_FAM_PRODUCT 0   /* periodic(q) * SE(P)   kernels.f90                      */
#define SGP_FAM_SQ      1   /* SE(q) * SE(P)         kernels_sq.f90                   */
#define SGP_FAM_SUM     2   /* periodic(q) + SE(P)   kernels_expl_per_q_sq_p.f90      */

/* post-step variants of the map loop (argument `kind`) */
#define SGP_MAP_PENDULUM 0  /* python/functions/func.py:216-237                       */
#define SGP_MAP_HENON    1  /* python/functions/func.py:239-260                       */
#define SGP_MAP_STANDARD 2  /* python/04_standard_map/func.py:218-254                 */
#define SGP_MAP_TOKAMAK  3  /* python/05_tokamak/SympGPR/func.py:182-211              */
#define SGP_MAP_TOKAMAK_SPLIT 5  /* python/05_tokamak/Split_SympGPR/func.py:184-219 (sgp_applymap_split) */
#define SGP_MAP_STANDARD_EXPL 4 /* applymap_expl python/04_standard_map/func.py:256-285: p wrapped, pdiff, q NOT wrapped */

/* root solver of the implicit equation (argument `solver`) */
#define SGP_SOLVER_HYBRD  0 /* MINPACK hybrd1, n = 1, tol 1e-13 (sympgpr.f90:107)     */
#define SGP_SOLVER_NEWTON 1 /* Newton with the analytic derivative, same tolerance   */
#define SGP_SOLVER_NEWTON_DELTA 3 /* Newton started at p + guess: for guess GPs trained on P - p (ztrainp = P - p,
                               * python/04_standard_map/main.py:89-90, 03, 05), where the reference starts hybrd1 at the
                               * bare difference.  Same root wherever the residual has one root; about half the sweeps */
#define SGP_SOLVER_EXPLICIT 2 /* no root solve: P = p - F_q(q, p); the explicit maps calcP_expl
                               * python/04_standard_map/func.py:174-179, python/01_pendulum/explicit/func_expl.py:107-127
                               * (the guess-GP arguments are ignored and may be empty) */

/* layout of the 16-double result block of the NLL entry points */
#define SGP_RES_NLL   0     /* 0.5 y'alpha + sum log diag L                           */
#define SGP_RES_DLX   1     /* dNLL/dlx                                               */
#define SGP_RES_DLY   2     /* dNLL/dly                                               */
#define SGP_RES_DSIG  3     /* dNLL/dsig (mathematically consistent form)             */
#define SGP_RES_INFO  4     /* 0, or 1-based index of the first non-positive pivot    */
#define SGP_RES_QUAD  5     /* 0.5 y'alpha                                            */
#define SGP_RES_LOGD  6     /* sum log diag L                                         */
#define SGP_RES_A     8     /* [8..10]  alpha' dK_theta alpha, theta = lx, ly, sig    */
#define SGP_RES_B     11    /* [11..13] trace(Kyinv dK_theta)                         */
#define SGP_RES_LEN   16

typedef struct sgp_ctx sgp_ctx;
typedef struct sgp_model sgp_model;

/* ---- context ------------------------------------------------------------------------------ */
int sgp_version(void);
unsigned long long sgp_launch_count(void);        /* kernels launched by this library so far */
const char* sgp_last_error(void);
int sgp_device_count(void);                       /* 0 when no usable CUDA device            */
int sgp_create(int device, sgp_ctx** out);        /* owns one stream + growable workspaces   */
int sgp_destroy(sgp_ctx* ctx);
int sgp_set_stream(sgp_ctx* ctx, void* cuda_stream);  /* borrow a cudaStream_t (NULL: own)   */
int sgp_synchronize(sgp_ctx* ctx);
int sgp_release_workspace(sgp_ctx* ctx);          /* free cached device buffers              */
/* Stage timers (bench.py): with profiling on, every NLL evaluation records CUDA events on the
 * context's stream between its stages; sgp_stage_times waits for the last evaluation and returns
 * the 7 durations in ms: fill, potrf, potrs, trtri, lauum, gradient contraction, finalize. */
int sgp_set_profiling(sgp_ctx* ctx, int on);
int sgp_stage_times(sgp_ctx* ctx, double* ms7);

/* ---- kernels / kernels_sq / kernels_sum modules: the 19 scalar functions ------------------
 * replaces REAL*8 function <name>_num(x_a, y_a, x_b, y_b, lx, ly[, p]), kernels.f90:1-231.
 * `which` indexes the generation order of python/04_standard_map/init_func.py:58-76:
 *  0 kern 1 dkdx 2 dkdy 3 dkdx0 4 dkdy0 5 d2kdxdx0 6 d2kdydy0 7 d2kdxdy0 8 d3kdxdx0dy0
 *  9 d3kdydy0dy0 10 d3kdxdy0dy0 11 dkdlx 12 dkdly 13 d3kdxdx0dlx 14 d3kdydy0dlx 15 d3kdxdy0dlx
 *  16 d3kdxdx0dly 17 d3kdydy0dly 18 d3kdxdy0dly.  Host arithmetic (a scalar call cannot pay a
 * kernel launch); the same closed forms (csrc/forms.cuh) run on the device everywhere else. */
double sgp_kernel_scalar(int fam, int which, double x_a, double y_a, double x_b, double y_b,
                         double lx, double ly, double per);

/* ---- sympgpr module, host buffers --------------------------------------------------------- */
/* sympgpr.f90:12-38  build_k(x,y,x0,y0,hyp,k): fills the (2N x 2N0) block of K (leading
 * dimension ldk >= 2N); N = len(x), N0 = len(x0). */
int sgp_build_k(sgp_ctx* ctx, int fam, double per, const double* x, const double* y, long N,
                const double* x0, const double* y0, long N0, const double* hyp3, double* K, long ldk);
/* sympgpr.f90:40-60  buildkreg(x,y,x0,y0,hyp,k): (N x N0). */
int sgp_buildkreg(sgp_ctx* ctx, int fam, double per, const double* x, const double* y, long N,
                  const double* x0, const double* y0, long N0, const double* hyp3, double* K, long ldk);
/* 2-DOF generalisation of build_k (NOT in the reference; BASELINE config 3): x = [q1; q2; P1; P2]
 * (4 arrays of N), x0 likewise (N0), hyp3 = [lq, lP, sig]; fills the (4N x 4N0) matrix of the mixed second
 * derivatives of the SE kernel in (q1, q2, P1, P2), blocks ordered like x.  csrc/dof2.cu. */
int sgp_build_k4(sgp_ctx* ctx, const double* x, long N, const double* x0, long N0, const double* hyp3,
                 double* K, long ldk);
/* sympgpr.f90:62-73  guessp(x,y,hypp,xtrainp,ytrainp,ztrainp,kyinvp) */
int sgp_guessp(sgp_ctx* ctx, int fam, double per, double x, double y, const double* hypp3,
               const double* xtrainp, const double* ytrainp, const double* ztrainp,
               const double* kyinvp, long np, double* out);
/* sympgpr.f90:75-86  calcq(x,y,xtrain,ytrain,hyp,kyinv,ztrain) */
int sgp_calcq(sgp_ctx* ctx, int fam, double per, double x, double y, const double* xtrain,
              const double* ytrain, const double* hyp3, const double* kyinv, const double* ztrain,
              long nt, double* out);
/* sympgpr.f90:88-125 calcp(x,y,hyp,hypp,xtrainp,ytrainp,ztrainp,kyinvp,xtrain,ytrain,ztrain,kyinv) */
int sgp_calcp(sgp_ctx* ctx, int fam, double per, int solver, double x, double y, const double* hyp3,
              const double* hypp3, const double* xtrainp, const double* ytrainp, const double* ztrainp,
              const double* kyinvp, long np, const double* xtrain, const double* ytrain,
              const double* ztrain, const double* kyinv, long nt, double* out);
/* sympgpr.f90:128-177 applymap_tok(...): qmap/pmap are (nm, ntest, 1) Fortran order, filled in
 * place; semantics follow the authoritative Python loop (NaN for lost orbits, numpy.mod),
 * python/05_tokamak/SympGPR/func.py:182-211; `kind` selects the other scripts' variants. */
int sgp_applymap_tok(sgp_ctx* ctx, int fam, double per, int solver, int kind, long nm, long ntest,
                     const double* hyp3, const double* hypp3, const double* q0map, const double* p0map,
                     const double* xtrainp, const double* ytrainp, const double* ztrainp,
                     const double* kyinvp, long np, const double* xtrain, const double* ytrain,
                     const double* ztrain, const double* kyinv, long nt, double* qmap, double* pmap);

/* ---- batched entry points (additive; what func.py's nll_* / main.py's inv() collapse into) -- */
/* nll_chol / nll_chol_reg / nll_grad / nll_grad_reg: python/05_tokamak/SympGPR/func.py:134-168,
 * python/02_pert_pendulum/func.py:132-162.  xin = [x(0:N); y(0:N)], z = observations (n),
 * hyp4 = [lx, ly, sig, sig2n]; reg = 0: derivative kernel, n = 2N; reg = 1: plain kernel, n = N;
 * reg = 2 / 3: the (q,q) / (P,P) Hessian block alone, n = N, value only -- nll_expl
 * python/04_standard_map/func.py:126-141 (the sum kernel's matrix is block diagonal and is fitted per block);
 * reg = 4: the 2-DOF 4x4-block kernel of sgp_build_k4, n = 4N, xin = [q1; q2; P1; P2], hyp4 = [lq, lP, sig, sig2n],
 * gradient w.r.t. (lq, lP[, sig]) as for reg = 0.
 * ngrad = 0 value only, 2 or 3 also the gradient.  res: SGP_RES_LEN doubles. */
int sgp_nll(sgp_ctx* ctx, int fam, double per, int reg, const double* hyp4, const double* xin,
            const double* z, long n, int ngrad, double* res);
int sgp_nll_dev(sgp_ctx* ctx, int fam, double per, int reg, const double* hyp4, const double* d_xin,
                const double* d_z, long n, int ngrad, double* d_res);
/* model finalisation: alpha = (K + |sig2n| I)^-1 z, optionally the full inverse (the scripts'
 * Kyinv = scipy.linalg.inv(...), python/01_pendulum/implicit/main.py:138-140,159-161) and the
 * Cholesky factor L (n x n, column-major, zeros above the diagonal).  NULL skips an output. */
int sgp_fit(sgp_ctx* ctx, int fam, double per, int reg, const double* hyp4, const double* xin,
            const double* z, long n, double* alpha, double* kyinv, double* L, double* res);

/* Ensemble map application with alpha hoisted (python/functions/func.py:216-237 and variants).
 * qmap/pmap/pdiff: (rows, E) C order, rows = 1 + (nm-1)/out_every, row r = state after
 * r*out_every steps (out_every = 1: the full history the scripts return; 0: no history).
 * qfinal/pfinal (E) always receive the last state.  stats[0] residual evaluations (all orbits),
 * stats[1] solver exits without convergence.  Any output pointer may be NULL. */
int sgp_applymap(sgp_ctx* ctx, int kind, int fam, double per, int solver, long nm, long E,
                 const double* q0, const double* p0, const double* hyp3, const double* hypp3,
                 const double* xtrainp, const double* ytrainp, const double* alphap, long np,
                 const double* xtrain, const double* ytrain, const double* alpha, long nt,
                 double* qmap, double* pmap, double* pdiff, long out_every,
                 double* qfinal, double* pfinal, unsigned long long* stats);

/* energy function of the fused quality metric (argument `ekind`) */
#define SGP_ENERGY_PENDULUM 1 /* H = p^2/2 + U0 (1 - cos(q + pi)); epar4 = {U0}: energy() python/01_pendulum/implicit/func.py:116-117 */
#define SGP_ENERGY_TOKAMAK  2 /* H = -Aph(r, q, 0), r = compute_r([p 1e-2, q, 0], 0.3); epar4 = {eps, m, phase}:
                               * energy() python/05_tokamak/Split_SympGPR/func.py:234-246, Aph fieldlines.f90:58-64 */
/* Ensemble map application with the quality metrics of `quality` (python/functions/func.py:262-272,
 * python/05_tokamak/Split_SympGPR/func.py:221-232) accumulated inside the map kernel, so that no history is
 * written: eosc[k] = std(H[:, k]) / mean(H[:, k]) and hmean[k] = mean(H[:, k]) over the rows 0, e_every,
 * 2 e_every, ... of the (nm, E) history the reference would hold (numpy.std: population standard deviation);
 * q1/p1 = row e_every (the first mapped state that `quality` compares with the reference orbit, `gd`).
 * Other arguments as sgp_applymap.  Any output pointer may be NULL. */
int sgp_applymap_quality(sgp_ctx* ctx, int kind, int fam, double per, int solver, long nm, long E,
                         const double* q0, const double* p0, const double* hyp3, const double* hypp3,
                         const double* xtrainp, const double* ytrainp, const double* alphap, long np,
                         const double* xtrain, const double* ytrain, const double* alpha, long nt,
                         int ekind, const double* epar4, long e_every, double* q1, double* p1,
                         double* qfinal, double* pfinal, double* eosc, double* hmean,
                         unsigned long long* stats);
/* 2-DOF map prediction with the 4 x 4-block kernel of sgp_build_k4 (BASELINE config 3; NOT in the reference, which
 * reduces Henon-Heiles to a section map -- python/03_henon_heiles/main.py:91-106; parity unpinned, oracle twin
 * oracle.applymap4).  Model: xtrain = [q1; q2; P1; P2] (4 N), alpha = (K4 + |sig2n| I)^-1 [p1-P1; p2-P2; Q1-q1; Q2-q2]
 * (4 N, sgp_fit with reg = 4), hyp3 = [lq, lP, sig].  One step: Newton (2 x 2, analytic Jacobian, start P = p) on
 * P = p - grad_q F(q, P), then Q = q + grad_P F(q, P).  q0, p0, qfinal, pfinal: (2, E); qmap, pmap: (rows, 2, E),
 * rows = 1 + (nm-1)/out_every, or NULL.  stats[0] Newton + dQ evaluations, stats[1] exits without convergence. */
int sgp_applymap4(sgp_ctx* ctx, long nm, long E, const double* q0, const double* p0, const double* hyp3,
                  const double* xtrain, const double* alpha, long N, double* qmap, double* pmap, long out_every,
                  double* qfinal, double* pfinal, unsigned long long* stats);
/* StandardMapIterate(k, nm, N, X0) python/04_standard_map/main.py:27-39: X0 (2, N) -> f (2, N, nm), C order,
 * J' = J + k sin(th), th' = th + J' (no wrap); the generator of the standard-map training and reference orbits. */
int sgp_standard_map_iterate(sgp_ctx* ctx, double k, long nm, long N, const double* X0, double* f);

/* Split map: nmodels learned maps (one per toroidal section) applied in turn, step s with map (s-1) mod nmodels;
 * loss test at the new angle, lost orbits NaN in q and p -- applymap_tok
 * python/05_tokamak/Split_SympGPR/func.py:184-219.  All sub-maps have np / nt training pairs; every model
 * array holds them one after the other (hyp3, hypp3: 3 each; xtrainp, ytrainp, alphap: np each; xtrain,
 * ytrain: nt each; alpha: 2 nt each).  qmap/pmap: (nsteps + 1, E) C order, row 0 = initial conditions. */
int sgp_applymap_split(sgp_ctx* ctx, int fam, double per, int solver, int nmodels, long nsteps, long E,
                       const double* q0, const double* p0, const double* hyp3, const double* hypp3,
                       const double* xtrainp, const double* ytrainp, const double* alphap, long np,
                       const double* xtrain, const double* ytrain, const double* alpha, long nt,
                       double* qmap, double* pmap, unsigned long long* stats);

/* Device-resident model + ensemble for repeated launches (bench, multi-GPU shards). */
int sgp_model_create(sgp_ctx* ctx, int fam, double per, const double* hyp3, const double* hypp3,
                     const double* xtrainp, const double* ytrainp, const double* alphap, long np,
                     const double* xtrain, const double* ytrain, const double* alpha, long nt,
                     sgp_model** out);                               /* host pointers in      */
int sgp_model_destroy(sgp_model* m);
int sgp_model_applymap_dev(sgp_ctx* ctx, const sgp_model* m, int kind, int solver, long nsteps, long E,
                           const double* d_q0, const double* d_p0, double* d_qfinal, double* d_pfinal,
                           double* d_qhist, double* d_phist, long out_every,
                           unsigned long long* d_stats);             /* async on ctx stream   */

/* sgp_model_applymap_dev with the fused quality metrics of sgp_applymap_quality; d_work3: 3 E doubles of scratch
 * (carried between work items); d_q1/d_p1 may be NULL. */
int sgp_model_applymap_quality_dev(sgp_ctx* ctx, const sgp_model* m, int kind, int solver, long nsteps, long E,
                                   const double* d_q0, const double* d_p0, double* d_qfinal, double* d_pfinal,
                                   int ekind, const double* epar4, long e_every, double* d_work3,
                                   double* d_q1, double* d_p1, double* d_eosc, double* d_hmean,
                                   unsigned long long* d_stats);     /* async on ctx stream   */

/* Statistics of the last map launch of this context (synchronises): passes over a training set summed over all
 * warps.  32 * passes against the lane-level evaluation count (stats[0] + one guess per orbit-step) tells how
 * well the independently advancing orbits of a warp share their passes. */
int sgp_map_last_passes(sgp_ctx* ctx, unsigned long long* passes);

/* sgp_guessp / sgp_calcq / sgp_calcp / sgp_applymap_tok keep alpha = Kyinv ztrain of the (Kyinv, ztrain) pairs they were
 * last handed on the device (the reference recomputes that matvec in every call, sympgpr.f90:72,85,121; an unchanged
 * Python map loop makes 2 E S calls with the same arrays).  Counters of that cache, for tests and profiling. */
int sgp_alpha_cache_stats(sgp_ctx* ctx, unsigned long long* hits, unsigned long long* misses);

/* ---- fieldlines module (tokamak loss test) ------------------------------------------------ */
/* fieldlines.f90:94-107 compute_r(z(3), rstart): z = (pth, th, ph); host arithmetic. */
double sgp_compute_r(double pth, double th, double ph, double rstart);
/* fieldlines.f90:34-39 Ath(r, th, ph) */
double sgp_ath(double r, double th, double ph);

/* ---- dense FP64 building blocks on a general SPD matrix (tests, profiling) ---------------- */
/* A: n x n column-major host (lower triangle read).  L / Ainv: n x n outputs or NULL.
 * logdet_half = sum log diag L. */
int sgp_spd_factor(sgp_ctx* ctx, const double* A, long n, double* L, double* Ainv, double* logdet_half);
/* DMMA GEMM self test: random operands, compares against a plain FP64 kernel; see csrc/dmma_gemm.cuh
 * for al/bl/mode.  Returns the largest absolute difference in *max_err. */
int sgp_selftest_gemm(sgp_ctx* ctx, int al, int bl, int mode, int Mt, int Nt, int K, double* max_err);
/* The DMMA GEMM on HOST operands (tests compare it with NumPy): C = beta C + alpha A Bt over the tile set / k-ranges of
 * `mode` (csrc/dmma_gemm.cuh); al / bl: 0 = element (m,k) at ptr[m + k ld], 1 = at ptr[k + m ld]; Mt, Nt in tiles of 128. */
int sgp_gemm_host(sgp_ctx* ctx, int al, int bl, int mode, int Mt, int Nt, int K, double alpha, double beta, const double* A,
                  long lda, const double* B, long ldb, double* C, long ldc);
/* ---- opt-in: FP64 products from the INT8 tensor pipe (tcgen05.mma kind::i8, Ozaki splitting; csrc/ozaki.cu) ---------
 * Not on any default path: north_star asks for DMMA in the Cholesky.  Self test of the tcgen05 plumbing: one 128 x 64 x K
 * INT8 product (K a multiple of 128) on the tensor pipe against a plain integer kernel; *mismatches = differing outputs. */
int sgp_i8mma_selftest(sgp_ctx* ctx, int K, int* mismatches, int* probe_ref, int* probe_got);
/* OPT-IN switch of a context: nslices = 4..8 routes the lauum stage (W = X^T X, a third of an NLL+gradient evaluation) of
 * sgp_nll / sgp_nll_dev / sgp_fit / sgp_spd_factor's inverse through the INT8 tensor pipe; 0 (default) = DMMA everywhere. */
int sgp_set_ozaki(sgp_ctx* ctx, int nslices);
/* The same switch with the stages named: stages = 1 lauum (what sgp_set_ozaki selects), 2 = Cholesky factor + triangular
 * inverse as ONE recursion whose products are sliced INT8 GEMMs (csrc/ozaki_chol.cu; blocks of <= leaf_n rows stay on the DMMA
 * kernels; leaf_n <= 0: 4096), 3 = both: then all three n^3/3 stages of an NLL+gradient evaluation leave the DMMA pipe.
 * Applies to evaluations that form the inverse and do not return the factor L (the recursion never holds L). */
int sgp_set_ozaki_ex(sgp_ctx* ctx, int nslices, int stages, long leaf_n);
/* Sliced product with every option the recursion uses, on HOST operands (test hook): la / lb = storage order of A (M x K) / B
 * (N x K): 0 element (r, k) at ptr[r + k ld], 1 at ptr[k + r ld]; ta / tb: 0 all entries valid, 1 only k <= r, 2 only k >= r (the
 * rest is never read); kmode: 1 k >= 128 tm, 2 k >= 64 tn, 4 k < 128 (tm + 1), 8 k < 64 (tn + 1) for the output tile (tm, tn) of
 * 128 x 64; lower != 0: only tiles with 64 tn <= 128 tm + 127 are written. */
int sgp_ozaki_gemm_host_ex(sgp_ctx* ctx, int ns, long M, long N, long K, double alpha, const double* A, long lda, int la, int ta,
                           const double* B, long ldb, int lb, int tb, double beta, double* C, long ldc, int kmode, int lower);
/* C (M x N, column-major) = alpha A B^T + beta C with every product formed on the INT8 tensor pipe from ns = 4..8 signed 7-bit
 * slices per operand (Ozaki splitting; exact integer slice products in TMEM, one FP64 combination per element).  A (M x K) and
 * B (N x K): element (r, k) at ptr[r + k ld] (host pointers).  ns = 7 reproduces FP64 GEMM to ~1e-14 of sum |a||b|. */
int sgp_ozaki_gemm_host(sgp_ctx* ctx, int ns, long M, long N, long K, double alpha, const double* A, long lda, const double* B,
                        long ldb, double beta, double* C, long ldc);
/* timing on random device operands: ms2[0] = slicing + GEMM per call, ms2[1] = the tensor-pipe GEMM alone (operands sliced) */
int sgp_ozaki_bench(sgp_ctx* ctx, int ns, long M, long N, long K, int reps, double* ms2);
/* vector FP64 ceiling: a register-resident DFMA loop on every SM; *dp_instr_per_s = thread-level DFMA instructions per second
 * (what the map kernels' FP64-pipe roofline is measured against) */
int sgp_bench_dfma(sgp_ctx* ctx, int reps, double* dp_instr_per_s);
/* DMMA GEMM timing: `reps` launches on random operands (alpha = -1, beta = 1), average ms per launch */
int sgp_bench_gemm(sgp_ctx* ctx, int al, int bl, int mode, int Mt, int Nt, int K, int reps, double* ms_avg);
/* timing hooks for bench.py (device pointers, async): the individual stages of one evaluation */
int sgp_fill_sym_dev(sgp_ctx* ctx, int fam, double per, int reg, const double* hyp4, const double* d_xin,
                     long n, double* d_K, long ld);
int sgp_potrf_dev(sgp_ctx* ctx, double* d_A, long n_pad, long ld, double* d_res);
/* full (2N x 2N0) build_k on device buffers: the HBM-write roofline case (8 n^2 bytes) */
int sgp_build_k_dev(sgp_ctx* ctx, int fam, double per, const double* d_x, const double* d_y, long N,
                    const double* d_x0, const double* d_y0, long N0, const double* hyp3, double* d_K, long ld);

#ifdef __cplusplus
}
#endif
#endif
