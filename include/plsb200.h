/*
 * plsb200.h -- C ABI of libplsb200.so, the B200 (sm_100a) PLS resampling engine.
 *
 * The reference (netneurolab/pypyls) has no FFI: its seam for this path is the
 * Python methods BasePLS.permutation() / BasePLS.bootstrap() and the numeric
 * primitives they call.  Each entry point below names the reference interface
 * it replaces (file:line into the reference tree).  The Python binding a
 * maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, a negative plsb_status otherwise;
 *     plsb_last_error() returns the message of the last failure on the
 *     calling thread;
 *   - all `d_` pointers are DEVICE pointers owned by the caller (e.g. torch
 *     tensors); the library never frees or keeps them beyond the call except
 *     where stated (plsb_set_data keeps no reference: it copies);
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); all
 *     work is enqueued on it and the call does not synchronise unless stated;
 *   - matrices are row-major fp64; index tables are int32, RESAMPLE-MAJOR:
 *     idx[r * S + s] = source row of destination row s in resample r (the
 *     reference stores the transpose, (S, n) int64, pyls/base.py:35,107);
 *   - a handle is used by one host thread at a time.
 *
 * Sizes:  S rows, B features, T behaviours, J = n_groups * n_cond cells,
 *         K = J*T (behavioural) or J (mean-centred), L = K (requires K <= B).
 */
#ifndef PLSB200_H
#define PLSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct plsb_ctx *plsb_handle_t;

enum plsb_status {
  PLSB_OK = 0,
  PLSB_ERR_ARG = -1,      /* bad argument / unsupported shape            */
  PLSB_ERR_CUDA = -2,     /* a CUDA runtime call or kernel failed         */
  PLSB_ERR_STATE = -3,    /* call order violated (e.g. no data set yet)   */
  PLSB_ERR_NOMEM = -4     /* workspace does not fit the configured limit  */
};

enum plsb_mode {
  PLSB_BEHAVIORAL_CORR = 0, /* behavioral_pls(covariance=False)            */
  PLSB_BEHAVIORAL_COV = 1,  /* behavioral_pls(covariance=True)             */
  PLSB_MEANCENTERED = 2,    /* meancentered_pls (mean_centering 0/1/2)     */
  PLSB_SIMPLS = 3           /* pls_regression                              */
};

/* Library version (major*10000 + minor*100 + patch). */
int plsb_version(void);
const char *plsb_last_error(void);

/* Handle life cycle.  `device` is the CUDA ordinal. */
int plsb_create(plsb_handle_t *out, int device);
int plsb_destroy(plsb_handle_t h);
/* Upper bound (bytes) for the per-chunk resample workspace: the stored
 * cross-covariances of one pass.  The library's own default is 8 GiB; the Python
 * front-end (pypyls_b200/engine.py) sets min(32 GiB, device memory / 5) -- fewer,
 * larger passes keep every grid full on a 180 GB part. */
int plsb_set_workspace_limit(plsb_handle_t h, uint64_t bytes);
/* Kernel behind the cross-covariance contraction (compute.xcorr's `Yn.T @ Xn`,
 * pyls/compute.py:92).  PLSB_GEMM_AUTO (default): the int8 slice GEMM on the
 * tcgen05 tensor cores (csrc/gemm_i8.cu: FP64 operands cut into `n_slices` signed
 * base-256 digit planes, exact int32 products in tensor memory, FP64 recombination)
 * wherever it applies -- dense operands, contraction over at most 224 rows --
 * and the FP64 DMMA kernel elsewhere; PLSB_GEMM_DMMA: always the DMMA kernel.
 * n_slices 5 / 6 / 7 (0 keeps the current value; default 6: entries within
 * ~1e-13 |a||x| of the FP64 product, 7: ~1e-15). */
enum { PLSB_GEMM_AUTO = 0, PLSB_GEMM_DMMA = 1 };
int plsb_set_gemm_backend(plsb_handle_t h, int backend, int n_slices);
/* Work the contraction kernels have executed since the last reset: int8
 * multiply-accumulates of the slice GEMM (padded tiles x digit-plane products)
 * and FP64 flop of the DMMA kernel.  For roofline accounting (bench.py). */
int plsb_gemm_work(plsb_handle_t h, double *i8_macs, double *dmma_flops, int reset);

/*
 * Analysis layout.  Replaces BasePLS.__init__ validation + utils.dummy_code
 * (pyls/base.py:254-283, pyls/utils.py:155-197): rows are ordered
 * group -> condition -> subject.  `T` is ignored for PLSB_MEANCENTERED.
 * `n_components` is used by PLSB_SIMPLS only.
 */
int plsb_configure(plsb_handle_t h, int mode, int S, int B, int T,
                   int n_groups, const int *groups, int n_cond,
                   int mean_centering, int n_components);

/*
 * Copies X (S,B) and Y (S,T) into padded internal buffers and precomputes the
 * per-cell statistics every resample shares (per-cell centred / z-scored X;
 * pyls/compute.py:84-87).  d_Y may be NULL for PLSB_MEANCENTERED.
 */
int plsb_set_data(plsb_handle_t h, const double *d_X, const double *d_Y,
                  void *stream);

/*
 * Original decomposition on the device.  Replaces BasePLS.svd on the
 * un-resampled data (pyls/base.py:362-363, 401-437; pyls/compute.py:10-52)
 * including sklearn's svd_flip sign convention on the B-side factor.
 *   d_U (B,L)  d_d (L)  d_V (K,L).  The result is also installed as the
 * "original" used by plsb_run_perms / plsb_run_boots.
 */
int plsb_decompose(plsb_handle_t h, double *d_U, double *d_d, double *d_V,
                   void *stream);
/* Install a caller-provided original decomposition instead. */
int plsb_set_original(plsb_handle_t h, const double *d_U, const double *d_d,
                      const double *d_V, void *stream);
/* x_scores = X @ U for the raw X (pyls/base.py:364).  d_out (S,L). */
int plsb_project_scores(plsb_handle_t h, const double *d_U, int L,
                        double *d_out, void *stream);

/*
 * On-device resample tables (counter-based RNG keyed by (seed, kind, resample
 * id, attempt) so any rank can generate any id).  Replace gen_permsamp /
 * gen_bootsamp (pyls/base.py:10-79, 82-159) with the same validity rules:
 * conditions stay with their subject, subjects must mix across groups, each
 * group keeps >= ceil(min_group/2) distinct subjects, duplicate columns are
 * re-drawn at most 500 times.  `first` is the global id of the first column;
 * duplicates are checked over [0, first+count) so the table is the same for
 * any sharding.  *h_n_exhausted (host) = columns that hit the 500-try cap.
 * d_idx is (count, S) int32.  These two calls synchronise the stream.
 */
int plsb_gen_perm_indices(plsb_handle_t h, uint64_t seed, int64_t first,
                          int count, int32_t *d_idx, int *h_n_exhausted,
                          void *stream);
int plsb_gen_boot_indices(plsb_handle_t h, uint64_t seed, int64_t first,
                          int count, int32_t *d_idx, int *h_n_exhausted,
                          void *stream);

/*
 * Permutation loop.  Replaces BasePLS.permutation + _single_perm with
 * use_permind=True, n_split=None (pyls/base.py:601-712) and
 * MeanCenteredPLS.make_permutation (pyls/types/meancentered.py:104-125).
 *   d_idx (count,S) int32;  d_dperm (count,L): permuted singular values
 *   (rotate != 0: Procrustes-rotated, base.py:696-700; else diag(d), :702).
 */
int plsb_run_perms(plsb_handle_t h, const int32_t *d_idx, int count,
                   int rotate, double *d_dperm, void *stream);
/* Rotated permutation singular values in SAMPLE SPACE (an algorithmic fast path,
 * not the cross-covariance GEMM): |R^T v_j|^2 = a_j^T (Xp Xp^T) a_j with the
 * S x S Gram matrix of the fixed data matrix computed once per plsb_set_data
 * and a_j the operand rows plsb_run_perms would contract with the data.  Same
 * arguments and results as plsb_run_perms(rotate = 1). */
int plsb_run_perms_gram(plsb_handle_t h, const int32_t *d_idx, int count,
                        double *d_dperm, void *stream);
/* The same for pre-permuted behaviour matrices (`permsamples` of shape (P,S,T)
 * with permindices=False, pyls/base.py:636-639, 689-692: spatial-null models
 * hand in Y already permuted): d_Yperm (count,S,T), X stays in place.
 * Behavioural modes only. */
int plsb_run_perms_prepermuted(plsb_handle_t h, const double *d_Yperm, int count,
                               int rotate, double *d_dperm, void *stream);

/*
 * Bootstrap loop.  Replaces BasePLS.bootstrap + _single_boot
 * (pyls/base.py:439-576), gen_distrib (pyls/types/behavioral.py:54-80,
 * pyls/types/meancentered.py:75-102) and compute.procrustes
 * (pyls/compute.py:240-264).
 *   d_distrib (count,K,L), may be NULL (see plsb_boot_distrib);  d_usum /
 *   d_usquare (B,L) are ACCUMULATED into (zero them first; base.py:475-476,
 *   510-511).
 */
int plsb_run_boots(plsb_handle_t h, const int32_t *d_idx, int count,
                   double *d_distrib, double *d_usum, double *d_usquare,
                   void *stream);
/* Only the bootstrap distribution of `count` resamples, d_distrib (count,K,L)
 * (gen_distrib, pyls/types/behavioral.py:54-80, pyls/types/meancentered.py:
 * 75-102): it needs the original weights alone, so a caller can have it -- and
 * start moving it to the host -- before the cross-covariance work of
 * plsb_run_boots (called with d_distrib = NULL) begins. */
int plsb_boot_distrib(plsb_handle_t h, const int32_t *d_idx, int count,
                      double *d_distrib, void *stream);
/* Resamples plsb_run_boots handles per internal pass for the current workspace
 * limit (<= count).  A caller that wants each pass's slice of `distrib` as soon as
 * it is final (to overlap its device->host copy with the next pass) can call
 * plsb_run_boots on blocks of this size; u_sum / u_square accumulate across calls. */
int plsb_boot_chunk(plsb_handle_t h, int count);

/* compute.perm_sig (pyls/compute.py:154-181): strict '>' count.
 * d_dperm (count,L), d_dorig (L) -> d_pvals (L). */
int plsb_perm_pvals(plsb_handle_t h, const double *d_dperm, int count, int L,
                    const double *d_dorig, double *d_pvals, void *stream);

/* compute.boot_ci (pyls/compute.py:184-209): numpy 'linear' percentiles over
 * the resample axis.  d_distrib (count, n_series) -> d_lo, d_hi (n_series). */
int plsb_percentile(plsb_handle_t h, const double *d_distrib, int count,
                    int n_series, double q_lo, double q_hi, double *d_lo,
                    double *d_hi, void *stream);

/* The same percentiles for series-major input: series j is the `count` values
 * at d_series + j * ld (what a rank holds after the series of the bootstrap
 * distribution have been dealt over the ranks: each rank selects the order
 * statistics of its own series over ALL resamples). */
int plsb_percentile_series(plsb_handle_t h, const double *d_series,
                           int n_series, int count, int64_t ld, double q_lo,
                           double q_hi, double *d_lo, double *d_hi,
                           void *stream);

/* d_out (cols, rows) = d_in (rows, cols)^T, both dense row-major: turns the
 * resample-major bootstrap distribution (count, K*L) into the series-major
 * blocks the ranks exchange before plsb_percentile_series. */
int plsb_transpose(plsb_handle_t h, const double *d_in, int rows, int cols,
                   double *d_out, void *stream);

/* compute.boot_rel (pyls/compute.py:212-237) incl. the "add the original
 * sample" step of behavioral.py:202-207 when add_orig != 0.
 * d_bs (B,L) = U @ diag(d) flattened to n_elem values; outputs d_bsr, d_se. */
int plsb_boot_ratio(plsb_handle_t h, const double *d_bs, const double *d_usum,
                    const double *d_usquare, int64_t n_elem, int n_boot,
                    int add_orig, double *d_bsr, double *d_se, void *stream);

/* Profiling aid: mean milliseconds per launch of one variant of the
 * cross-covariance GEMM (compute.xcorr's `Yn.T @ Xn`, pyls/compute.py:92, for a
 * stack of M operand rows) on synthetic operands.  variant 0 = store,
 * 1 = store with column scales, 2 = row sums of squares (rotated permutations). */
int plsb_gemm_probe(plsb_handle_t h, int variant, int M, int N, int Kd,
                    int k_valid, int scale_div, int iters, double *ms_out,
                    void *stream);

/* ---- primitives (exposed for unit tests and for callers that want the
 *      compute.py-level operations; pyls/compute.py:55-94, 10-52, 240-264) -- */

/* C (M,N) = A (M,Kd) @ X (Kd,N), fp64 DMMA; all three dense row-major without
 * padding (the call pads internally).  The contraction of compute.xcorr
 * (pyls/compute.py:92) as the resampling drivers use it: A = stacked
 * per-resample left operands, X = shared data matrix. */
int plsb_dgemm(plsb_handle_t h, const double *d_A, const double *d_X, int M,
               int N, int Kd, double *d_C, void *stream);
/* Cross-covariance / cross-correlation matrices of `count` resamples.
 * Replaces BasePLS.gen_covcorr on resampled data (pyls/base.py:401-437 ->
 * pyls/types/behavioral.py:27-52, pyls/types/meancentered.py:50-73).
 * bootstrap != 0: both X and Y follow d_idx (base.py:569-570); otherwise the
 * permutation rule of the analysis type applies (Y rows for behavioural,
 * base.py:599; X rows for mean-centred, meancentered.py:125).  d_idx may be
 * NULL with count == 1 for the un-resampled matrix.  d_R is (count, K, B). */
int plsb_crosscov(plsb_handle_t h, const int32_t *d_idx, int count,
                  int bootstrap, double *d_R, void *stream);
/* Batched small decomposition (replaces the per-resample compute.svd +
 * compute.procrustes, pyls/compute.py:36-49, 240-264): for each of `count`
 * matrices G (K,K) = R R^T and H (K,L) = R U_orig, M = V Q with
 * G = V diag(lam) V^T and Q the transposed polar factor of H^T V lam^-1/2, so
 * that R^T M = U_boot d Q.  d_dorig (L) are the original singular values
 * (NULL = all non-null): null original latent variables are left out of the
 * rotation.  d_lam (count,K), optional, receives lam sorted descending. */
int plsb_small_decomp(plsb_handle_t h, const double *d_G, const double *d_H,
                      int count, int K, int L, const double *d_dorig,
                      double *d_M, double *d_lam, void *stream);
/* The two streaming contractions around it, on caller-supplied dense matrices
 * (the engine runs them on its padded workspace; these entries serve tests):
 * G[r] (K,K) = R[r] R[r]^T and, if d_Uo is given, H[r] (K,L) = R[r] U_orig for
 * `count` matrices R[r] (K,B) -- the inputs of plsb_small_decomp
 * (pyls/compute.py:36-49, 260). */
int plsb_gram_proj(plsb_handle_t h, const double *d_R, int count, int K, int B,
                   const double *d_Uo, int L, double *d_G, double *d_H,
                   void *stream);
/* u_sum (B,L) += sum_r R[r]^T M[r], u_square (B,L) += sum_r (R[r]^T M[r])^2
 * (pyls/base.py:510-511 with the Procrustes rotation folded into M). */
int plsb_accum_u(plsb_handle_t h, const double *d_R, int count, int K, int B,
                 const double *d_M, int L, double *d_usum, double *d_usquare,
                 void *stream);

/*
 * Cross-validation of a behavioural analysis.  Replaces BehavioralPLS.crossval +
 * _single_crossval (pyls/types/behavioral.py:82-170) and compute.rescale_test
 * (pyls/compute.py:129-151) for `count` train / test splits at once:
 *   d_train (count,S) int32, non-zero = training row (a column of gen_splits,
 *   pyls/base.py:162-229, with test_size = the held-out fraction);
 *   max_test = largest number of test rows of any split (K + max_test rounded
 *   up to a multiple of T must not exceed 80);
 *   d_r, d_r2 (count,T): Pearson r and R^2 of every behaviour, predicted from
 *   the training decomposition, against the held-out rows.
 */
int plsb_crossval(plsb_handle_t h, const int32_t *d_train, int count, int max_test,
                  double *d_r, double *d_r2, void *stream);

/*
 * Split-half reliability of the singular vectors.  Replaces BasePLS.split_half
 * (pyls/base.py:714-770) for `count` data sets at once -- the call inside every
 * permutation (base.py:704-708: d_idx (count,S) permutation table, or d_yperm
 * (count,S,T) pre-permuted behaviour matrices; use_original = 0: every data set
 * is scored against its own decomposition) and the one on the original data
 * (base.py:373-380: d_idx = d_yperm = NULL, count = 1, use_original = 1: the
 * decomposition installed by plsb_decompose / plsb_set_original).
 *   d_masks (count, n_split, S) int32: half / half masks, non-zero = first half
 *   (columns of gen_splits(..., test_size=0.5), pyls/base.py:162-229; the
 *   reference draws a fresh set per permutation, seeded by its number);
 *   d_ucorr, d_vcorr (count, L): correlations between the two halves of the
 *   projected left / right singular vectors, averaged over the masks.
 * Needs 2 K <= 80.  Numerically null latent variables report 0.
 */
/* Device counterpart of gen_splits (pyls/base.py:162-229): `count` sets (ids
 * first .. first+count-1; one per permutation, base.py:704-708) of `n_split`
 * masks, d_masks (count, n_split, S) int32 0/1.  Per group a coin flip between
 * ceil and floor of n_g * train_fraction subjects, drawn without replacement;
 * conditions follow their subject; a mask equal to an earlier one of its set is
 * re-drawn (<= 500 draws, *h_n_exhausted counts the masks that hit the cap).
 * Counter-based: a set depends on (seed, id) only. */
int plsb_gen_split_masks(plsb_handle_t h, uint64_t seed, int64_t first, int count,
                         int n_split, double train_fraction, int32_t *d_masks,
                         int *h_n_exhausted, void *stream);
int plsb_split_half(plsb_handle_t h, const int32_t *d_idx, const double *d_yperm,
                    int count, const int32_t *d_masks, int n_split,
                    int use_original, double *d_ucorr, double *d_vcorr,
                    void *stream);

/* ---- SIMPLS (pls_regression; mode PLSB_SIMPLS, n_groups = 1, n_cond = 1) -------
 * plsb_set_data takes X (S,B) and Y (S,T) already column-centred
 * (pyls/types/regression.py:395-396).  `d_omega` tables hold the Gaussian test
 * matrices sklearn's randomized_svd would draw, (T, 11) row-major each
 * (compute.svd(Cov, n_components=1, seed) at pyls/types/regression.py:103); they
 * may be NULL when T <= 11 (the probes then span the whole space).
 *
 * plsb_simpls_decompose: original decomposition; replaces PLSRegression.svd ->
 *   simpls on the un-resampled data (pyls/types/regression.py:248-277, 56-186).
 *   d_omega (L,T,11): one table per component (the reference draws them in turn
 *   from the analysis' RandomState).  Outputs x_weights (B,L) with sklearn's
 *   svd_flip signs and pctvar in Y (L); installs the weights as the original.
 * plsb_simpls_run_perms: replaces PLSRegression._single_perm over a batch
 *   (regression.py:329-373): Y rows follow d_idx (count,S), X fixed;
 *   d_omega (count,T,11) (resample i uses RandomState(i)); d_pctvar (count,L).
 * plsb_simpls_run_boots: replaces PLSRegression._single_boot over a batch
 *   (regression.py:279-327): X and Y rows follow d_idx; x_weights are sign
 *   aligned with the original (efficient_corr, :317-320) and ACCUMULATED into
 *   d_usum / d_usquare (B,L); d_distrib (count,T,L) = Yi^T (Xi x_weights)
 *   (:323-325); d_pctvar (count,L).
 * plsb_simpls_run_boots_yres: the same for a three-dimensional Y (S,T,C): the
 *   reference aggregates a bootstrap sample of the third axis into a fresh (S,T)
 *   behaviour matrix for every bootstrap (`aggfunc(Y[..., cboot], axis=-1)`,
 *   regression.py:207-235, 308-310); d_yres (count,S,T) holds those matrices
 *   (uncentred, as the reference uses them) and resample r draws its rows from
 *   d_yres[r] through d_idx[r]. */
int plsb_simpls_set_original(plsb_handle_t h, const double *d_xweights,
                             void *stream);
/* The (T, 11) Gaussian test matrices of resamples [first, first + count) on the
 * device: table i is RandomState(i).normal(size=(T, 11)) of NumPy's legacy
 * generator (what compute.svd(..., seed=i) makes sklearn's randomized_svd draw,
 * pyls/types/regression.py:103, pyls/base.py:646-648, 502-507), replayed by an
 * MT19937 + polar-method kernel -- so that only the seed crosses PCIe.
 * d_omega (count, T, 11). */
int plsb_gen_gaussian_tables(plsb_handle_t h, int64_t first, int count, int T,
                             double *d_omega, void *stream);
/* Rows the reference's get_mask drops (pyls/types/regression.py:48-53: a row of X or
 * of Y that is all NaN): d_valid_x, d_valid_y (S) int32, 0 = that row of X / of Y is
 * missing; both NULL clears the masks.  Upload X / Y with those rows zero-filled.  Every
 * decomposition then runs on the rows of the RESAMPLED matrices whose X source and Y
 * source are both present (regression.py:271-272, 322-323; permutations move the rows of
 * Y only).  Call after plsb_set_data and before plsb_simpls_decompose. */
int plsb_simpls_set_row_mask(plsb_handle_t h, const int32_t *d_valid_x,
                             const int32_t *d_valid_y, void *stream);
int plsb_simpls_decompose(plsb_handle_t h, const double *d_omega,
                          double *d_xweights, double *d_pctvar, void *stream);
int plsb_simpls_run_perms(plsb_handle_t h, const int32_t *d_idx, int count,
                          const double *d_omega, double *d_pctvar,
                          void *stream);
int plsb_simpls_run_boots(plsb_handle_t h, const int32_t *d_idx, int count,
                          const double *d_omega, double *d_pctvar,
                          double *d_distrib, double *d_usum, double *d_usquare,
                          void *stream);
int plsb_simpls_run_boots_yres(plsb_handle_t h, const int32_t *d_idx, int count,
                               const double *d_omega, const double *d_yres,
                               double *d_pctvar, double *d_distrib,
                               double *d_usum, double *d_usquare, void *stream);

/* Counters for bench / tests: kernels launched by this handle so far. */
int64_t plsb_launch_count(plsb_handle_t h);

/* Optional per-launch CUDA-event timing, grouped by kernel class (the bench's
 * roofline uses it).  plsb_timing_read synchronises the recorded events, writes
 * the milliseconds and launch counts accumulated since the previous read into
 * arrays of plsb_timing_classes() entries and resets them. */
int plsb_timing_enable(plsb_handle_t h, int on);
int plsb_timing_classes(void);
const char *plsb_timing_class_name(int cls);
int plsb_timing_read(plsb_handle_t h, double *ms, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* PLSB200_H */
