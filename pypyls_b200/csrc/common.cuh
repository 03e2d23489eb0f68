// Internal declarations shared by the translation units of libplsb200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/plsb200.h"

namespace plsb {

// ---- error plumbing -------------------------------------------------------
void set_error(const char *fmt, ...);

#define PLSB_CUDA(expr)                                                      \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) {                                                 \
      plsb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,          \
                      cudaGetErrorString(_e));                               \
      return PLSB_ERR_CUDA;                                                  \
    }                                                                        \
  } while (0)

#define PLSB_CHECK(cond, code, ...)                                          \
  do {                                                                       \
    if (!(cond)) {                                                           \
      plsb::set_error(__VA_ARGS__);                                          \
      return (code);                                                         \
    }                                                                        \
  } while (0)

#define PLSB_TRY(expr)                                                       \
  do {                                                                       \
    int _s = (expr);                                                         \
    if (_s != PLSB_OK) return _s;                                            \
  } while (0)

// kernel launch check (no sync); every wrapper bumps the handle's counter
#define PLSB_LAUNCHED(h)                                                     \
  do {                                                                       \
    PLSB_CUDA(cudaGetLastError());                                           \
    (h)->launches++;                                                         \
  } while (0)

// ---- tiny device buffer that grows on demand -------------------------------
struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return PLSB_OK;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&p, need);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu bytes) failed: %s", need, cudaGetErrorString(e));
      (void)cudaGetLastError();
      return PLSB_ERR_NOMEM;
    }
    bytes = need;
    return PLSB_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// integer tuning knob: environment variable `name` if set, else `dflt`
inline int tune_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
inline long long round_up_ll(long long x, long long m) { return (x + m - 1) / m * m; }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }
// row pitch (doubles) of the rotation matrices M (count, K, pitch) that accum_u streams:
// the bank-conflict-free shared-memory pitch, so one bulk copy moves a whole matrix
inline int accum_ldm(int L) { return round_up(L, 8) + 4; }

// ---- GEMM tile geometry (gemm_dmma.cu) --------------------------------------
constexpr int GEMM_BM = 128;
constexpr int GEMM_BN = 128;
constexpr int GEMM_BK = 16;

// kernel classes for the optional per-launch event timing (bench / profiling)
enum KernelClass {
  KC_GEMM = 0,   // xcov_gemm_kernel (DMMA cross-covariance contraction)
  KC_BUILD,      // operand builders + distrib
  KC_GRAM,       // gram_proj
  KC_SMALL,      // small_decomp (Jacobi / Procrustes)
  KC_ACCUM,      // accum_u + partial reduction
  KC_STATS,      // colscale, finish_rowsq, pvals, percentile, boot_ratio
  KC_INDEX,      // index generation
  KC_PREP,       // pad / prep / projections
  KC_COUNT
};

constexpr int MAX_K = 80;     // largest K the fragment-table kernels (gram_proj, accum_u) are instantiated for
constexpr int SMALL_K_MAX = 12;   // up to here the FMA streaming kernels of small_k.cu take the two passes
constexpr int MAX_COND = 32;  // index generator: conditions per subject

// ---- analysis layout --------------------------------------------------------
struct Layout {
  int mode = -1;
  int S = 0, B = 0, T = 0, J = 0, K = 0, L = 0;
  int n_groups = 0, n_cond = 1, mean_centering = 0, n_components = 0;
  int n_subj = 0;
  std::vector<int> groups;      // subjects per group
  std::vector<int> cell_start;  // first row of each cell (J+1 entries)
  int S_pad = 0;                // rows of the padded data matrices == lda of operands
  int ldx = 0;                  // leading dimension (elements) of padded (., B) matrices
  int kr_max = 0;               // longest per-cell contraction range (multiple of GEMM_BK)
  bool behavioral() const { return mode == PLSB_BEHAVIORAL_CORR || mode == PLSB_BEHAVIORAL_COV; }
  bool corr() const { return mode == PLSB_BEHAVIORAL_CORR; }
  bool simpls() const { return mode == PLSB_SIMPLS; }
  // more latent rows than features: the small problem is the B x B feature-side Gram matrix
  bool tall() const { return !simpls() && K > B; }
};

}  // namespace plsb

// The opaque handle of the C ABI.
struct plsb_ctx {
  int device = 0;
  plsb::Layout lay;
  bool configured = false, has_data = false, has_original = false;
  uint64_t ws_limit = 8ull << 30;
  int64_t launches = 0;
  int sm_count = 148;
  // optional event timing: (class, start, stop) per launch since the last read
  bool timing = false;
  struct TimedLaunch { int cls; cudaEvent_t a, b; };
  std::vector<TimedLaunch> timed;

  // layout tables on the device: int cell_start[J+1], cell_of_row[S], cell_n[J],
  // group_start[n_groups+1] (subject units)
  plsb::DevBuf tables;
  const int *d_cell_start = nullptr, *d_cell_of_row = nullptr, *d_cell_n = nullptr,
            *d_group_start = nullptr;
  // per cell: staged contraction range [x, y) (whole GEMM_BK chunks) and the k steps of 4
  // inside it that touch the cell's rows [z, w)
  const int4 *d_cell_kr = nullptr;

  // data-dependent state (set_data); all (S_pad, ldx) zero padded
  plsb::DevBuf Xraw;   // raw X
  plsb::DevBuf Xcell;  // behavioural: X z-scored (corr) / centred (cov) within each cell
  plsb::DevBuf Xglob;  // X centred (and for corr scaled) with whole-column statistics
  plsb::DevBuf Y;      // Y (S, T)
  plsb::DevBuf Cmat;   // mean-centred: operator C (J, S)
  plsb::DevBuf rowmask;   // SIMPLS: optional (2, S) int32 (rows of X, rows of Y), 0 = missing
  bool has_rowmask = false;
  // original decomposition
  plsb::DevBuf Uo;     // (B, L)
  plsb::DevBuf Kx;     // Gram matrix of the permutation data matrix, (S_pad, round_up(S_pad,128))
  bool has_kx = false;
  plsb::DevBuf UoT;    // Uo transposed, (L, ldx) zero padded: gram_proj's TMA row copies
  plsb::DevBuf VoT;    // tall analyses: Vo transposed, (L, ldk) zero padded, ldk = round_up(K, 128)
  plsb::DevBuf Vo;     // (K, L)
  plsb::DevBuf dorig;  // (L)
  plsb::DevBuf Sx;     // Xraw @ normalize(Uo)   (S, L)
  plsb::DevBuf norms;  // (L)
  // int8 digit planes of the slice GEMM (gemm_i8.cu): planes of the fixed right operands
  // (the data matrices) are cached until the data changes, planes of A are per launch
  struct PlaneCache {
    const double *src = nullptr;
    uint64_t epoch = 0, stamp = 0;
    int S = 0, ldx = 0, N_pad = 0, k_valid = 0;
    bool square = false;
    plsb::DevBuf img, scale;
  };
  PlaneCache xplanes[3], xplanes_tmp;
  uint64_t data_epoch = 1, plane_stamp = 0;
  plsb::DevBuf aplanes, ascale;
  int gemm_backend = PLSB_GEMM_AUTO;
  int gemm_slices = 6;
  // work counters of the contraction kernels since the last plsb_gemm_work(reset)
  double i8_macs = 0.0, dmma_flops = 0.0;
  // per-chunk workspaces
  plsb::DevBuf A, Ac, R, S1, S2, G, H, M, lam, rowsq, part, misc, idxall, flags, maps, pctl, big;
};

namespace plsb {

// Brackets the launches of one wrapper with events when timing is enabled.
struct KernelTimer {
  plsb_ctx *h;
  cudaStream_t st;
  cudaEvent_t a = nullptr, b = nullptr;
  int cls;
  KernelTimer(plsb_ctx *h_, int cls_, cudaStream_t st_) : h(h_), st(st_), cls(cls_) {
    if (!h->timing) return;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) {
      a = b = nullptr;
      return;
    }
    cudaEventRecord(a, st);
  }
  ~KernelTimer() {
    if (!a) return;
    cudaEventRecord(b, st);
    h->timed.push_back({cls, a, b});
  }
};

// ---- kernel launch wrappers (each returns plsb_status) ----------------------

// C = A @ X (+ epilogues).  See gemm_dmma.cu.
struct GemmArgs {
  const double *A = nullptr;   // (M_pad, lda), rows beyond M zero
  int lda = 0;
  const double *X = nullptr;   // (Kd, ldx)
  int ldx = 0;
  int M_pad = 0;               // multiple of GEMM_BM
  int N_pad = 0;               // multiple of GEMM_BN (<= ldx)
  int Kd = 0;                  // contraction length, multiple of GEMM_BK (padding zero)
  int k_valid = 0;             // rows of X beyond this are zero padding (0: Kd): their k steps are skipped
  const int4 *kranges = nullptr;  // optional per-M-tile [kbeg,kend) (kbeg even) + non-zero k steps [z,w)
  int k_len = 0;               // longest contraction range in kranges (0: Kd); picks the tile
  bool x_persistent = false;   // X is a data matrix of the handle (its digit planes are cached)
  bool a_counts = false;       // A holds small integers (multiplicities): they fill one digit plane
                               // (the slice GEMM skips the empty ones, as found on the device);
                               // only the work accounting uses this hint
  bool square_b = false;       // use X*X elementwise as the right operand
  // STORE epilogue
  double *C = nullptr;
  long long ldc = 0;
  const int *row_map = nullptr;   // optional: output row of operand row m, <0 = skip
  const double *scale = nullptr;  // optional: C[m,n] *= scale[(m / scale_div) * lds + n]
  int scale_div = 1;
  int scale_rows = 0;             // rows of `scale` (0: ceil(M_pad / scale_div))
  long long lds = 0;
  // ROWSUMSQ epilogue: rowsq[split * M_pad + m] = sum_n C[m,n]^2 over the split
  double *rowsq = nullptr;
  int n_splits = 1;
};
int launch_gemm(plsb_ctx *h, const GemmArgs &a, cudaStream_t st);
bool gemm_small_tile(int klen);
int gemm_pick_splits(const plsb_ctx *h, int M_pad, int n_ntiles, bool small_tile);
// splits of the ROWSUMSQ epilogue for the kernel launch_gemm will pick for `a`
int gemm_rowsq_splits(const plsb_ctx *h, const GemmArgs &a);
// int8 slice GEMM on the tcgen05 tensor cores (gemm_i8.cu)
bool gemm_i8_applies(const plsb_ctx *h, const GemmArgs &a);
int gemm_i8_pick_splits(const plsb_ctx *h, int M_pad, int n_ntiles);
int gemm_i8_slices(const plsb_ctx *h);
int launch_gemm_i8(plsb_ctx *h, const GemmArgs &a, cudaStream_t st);

// data preparation (prep.cu)
int launch_pad_copy(plsb_ctx *h, const double *X, int S, int B, double *out, int S_pad, int ldx,
                    cudaStream_t st);
int launch_unpad_copy(plsb_ctx *h, const double *in, long long ld_in, int rows, int cols,
                      double *out, cudaStream_t st);
int launch_prep_cells(plsb_ctx *h, const double *Xraw, double *Xcell, double *Xglob,
                      cudaStream_t st);
int launch_colnorm(plsb_ctx *h, const double *U, int B, int L, double *norms, cudaStream_t st);
int launch_xproj(plsb_ctx *h, const double *Xmat, int ldx_, int S, int B, const double *U, int L,
                 const double *norms, double *out, cudaStream_t st);
int launch_normalize_flip(plsb_ctx *h, const double *Uraw, int B, int L, const double *lam,
                          double *U, double *V, int K, double *d, cudaStream_t st);
int launch_colgram(plsb_ctx *h, const double *W, int B, int K, double *G, cudaStream_t st);
int launch_matmul_small(plsb_ctx *h, const double *A, const double *Bm, int n, double *C,
                        cudaStream_t st);

// operand builders / distrib (operands.cu)
enum BuildKind { BUILD_ROT = 0, BUILD_PLAIN = 1, BUILD_BOOT = 2, BUILD_TRAIN = 3, BUILD_HALF = 4 };
// BUILD_HALF: `idx` holds split-half masks (n_perm, n_split, S); operand block r is half
// (r & 1) of mask (r >> 1) = (permutation (r >> 1) / ns, split s0 + (r >> 1) % ns)
struct HalfSpec {
  const int32_t *yidx = nullptr;   // optional (n_perm, S) permutation table of the data being split
  int ns = 1, s0 = 0, n_split = 1;
};
int launch_build(plsb_ctx *h, int kind, const int32_t *idx, const double *yperm, int count,
                 double *A, double *Ac, double *distrib, long long cellpad_w, long long cellpad_c,
                 cudaStream_t st, int *ntrain = nullptr, double *ytrain = nullptr,
                 const HalfSpec *half = nullptr);
int launch_build_maps(plsb_ctx *h, int n, int rows_pc, int stride_r, long long cellpad,
                      int *row_map, int4 *kranges, cudaStream_t st);

// streaming kernels over stored R (stream_kernels.cu)
int launch_finish_rowsq(plsb_ctx *h, const double *rowsq, int n_splits, int M_pad, int n_rows,
                        double *out, cudaStream_t st);
int launch_rowdot_sqrt(plsb_ctx *h, const double *T, long long ldt, const double *A, int lda,
                       int n_cols, long long n_rows, double *out, cudaStream_t st);
int launch_colscale(plsb_ctx *h, double *S1, const double *S2, int n_rows, long long ld,
                    cudaStream_t st, int rows_per_resample = 0, const int *nrow = nullptr);
int launch_colstats(plsb_ctx *h, const int32_t *idx, int n_res, double *scale, bool *done,
                    cudaStream_t st);
int launch_gram_proj(plsb_ctx *h, const double *R, long long ldr, int count, int K,
                     const double *UoT, int L, double *G, double *H, cudaStream_t st,
                     long long uot_stride = 0, int uot_div = 1);
int launch_accum_u(plsb_ctx *h, const double *R, long long ldr, int count, int K, int B,
                   const double *M, int L, double *usum, double *usq, cudaStream_t st);

// generic (any K, L) versions of the two passes above (large_k.cu)
int launch_gram_proj_generic(plsb_ctx *h, const double *R, long long ldr, int count, int K, int B,
                             const double *UoT, int L, double *G, double *H, cudaStream_t st,
                             long long uot_stride, int uot_div);
int launch_accum_u_generic(plsb_ctx *h, const double *R, long long ldr, int count, int K, int B,
                           const double *M, int ldm, int L, double *usum, double *usq,
                           cudaStream_t st);

// FMA + TMA streaming versions for few latent variables (small_k.cu)
bool small_k_applies(int K, int L, bool proj);
int launch_gram_proj_small(plsb_ctx *h, const double *R, long long ldr, int count, int K,
                           const double *UoT, int L, double *G, double *H, cudaStream_t st,
                           long long uot_stride, int uot_div);
int launch_accum_u_small(plsb_ctx *h, const double *R, long long ldr, int count, int K, int B,
                         const double *M, int L, double *usum, double *usq, cudaStream_t st);

// small matrices (small_matrix.cu)
// M is written as (count, K, ldm): ldm == L dense, ldm == accum_ldm(L) for launch_accum_u
int launch_small_decomp(plsb_ctx *h, const double *G, const double *H, int count, int K, int L,
                        const double *dorig, double *M, int ldm, double *lam, cudaStream_t st);
// G[r] is read at G + r * g_stride with row pitch ldg (0, 0: dense K x K blocks)
int launch_sym_eig(plsb_ctx *h, const double *G, int count, int K, double *V, double *lam,
                   int sqrt_lam, cudaStream_t st, int ldg = 0, long long g_stride = 0);
int launch_rotation_tall(plsb_ctx *h, const double *Uorig, const double *V, const double *lam,
                         int count, int K, const double *dorig, double *M, cudaStream_t st);
// tall analyses (tall.cu): batched transpose, quadratic forms, accumulation of small matrices
int launch_transpose_batch(plsb_ctx *h, const double *in, int rows, int cols, long long ld_in,
                           long long in_stride, double *out, long long ld_out,
                           long long out_stride, int count, cudaStream_t st);
int launch_quadform_sqrt(plsb_ctx *h, const double *G, const double *M, int count, int n, int L,
                         double *out, cudaStream_t st);
int launch_accum_small(plsb_ctx *h, const double *M, int count, long long n_elem, double *usum,
                       double *usq, cudaStream_t st);

// SIMPLS (simpls.cu)
int launch_simpls(plsb_ctx *h, const int32_t *idx, int count, int boot, int emit_ops,
                  const double *omega, long long om_stride_r, long long om_stride_c, double *pct,
                  double *distrib, cudaStream_t st, const double *yres = nullptr);
int launch_transpose(plsb_ctx *h, const double *in, int rows, int cols, int ld_in, double *out,
                     cudaStream_t st);
// out (cols, ld_out) = in (rows, cols)^T, columns >= rows of out zeroed
int launch_transpose_pad(plsb_ctx *h, const double *in, int rows, int cols, double *out,
                         long long ld_out, cudaStream_t st);
int launch_identity_blocks(plsb_ctx *h, double *M, int n, int L, int ldm, cudaStream_t st);
int launch_xweights_flip(plsb_ctx *h, const double *Rw, long long ldr, int B, int L, double *xw,
                         cudaStream_t st);
int launch_colcenter(plsb_ctx *h, const double *U, int B, int L, double *out, cudaStream_t st);

// cross-validation (crossval.cu)
int launch_rescale_test(plsb_ctx *h, const int32_t *mask, int count, int max_test,
                        const double *S1, const double *S2, int rows_per_resample,
                        const int *ntrain, double *Z, int stride, cudaStream_t st);
int launch_cv_score(plsb_ctx *h, const int32_t *mask, int count, int max_test, const double *Gz,
                    int stride, const double *V, const double *lam, const double *ytrain,
                    double *r_out, double *r2_out, cudaStream_t st);

// split-half resampling (splithalf.cu)
int launch_projblock(plsb_ctx *h, const double *R, int n, double *PB, cudaStream_t st);
int launch_splithalf_score(plsb_ctx *h, const double *G, const double *H, int n_perm, int ns,
                           const double *V, long long v_stride, const double *d,
                           long long d_stride, int n_split, double *ucorr, double *vcorr,
                           cudaStream_t st);

// index generation (indexgen.cu), statistics (stats.cu)
int gen_indices(plsb_ctx *h, bool boot, uint64_t seed, int64_t first, int count, int32_t *d_idx,
                int *h_n_exhausted, cudaStream_t st);
int gen_gaussian_tables(plsb_ctx *h, int64_t first, int count, int per, double *d_out,
                        cudaStream_t st);
int gen_split_masks(plsb_ctx *h, uint64_t seed, int64_t first, int count, int n_split, double frac,
                    int32_t *d_masks, int *h_n_exhausted, cudaStream_t st);
int launch_pvals(plsb_ctx *h, const double *dperm, int count, int L, const double *dorig,
                 double *pvals, cudaStream_t st);
int launch_percentile(plsb_ctx *h, const double *distrib, int count, int n_series, double qlo,
                      double qhi, double *lo, double *hi, cudaStream_t st);
int launch_percentile_series(plsb_ctx *h, const double *series, long long ld, int count,
                             int n_series, double qlo, double qhi, double *lo, double *hi,
                             cudaStream_t st);
int launch_boot_ratio(plsb_ctx *h, const double *bs, const double *usum, const double *usq,
                      long long n, int n_boot, int add_orig, double *bsr, double *se,
                      cudaStream_t st);

}  // namespace plsb
