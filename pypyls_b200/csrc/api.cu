// C ABI of libplsb200.so (include/plsb200.h): handle life cycle, the one-time
// data preparation, the original decomposition and the two resampling drivers
// that replace BasePLS.permutation / BasePLS.bootstrap (pyls/base.py:439-528,
// 601-652).  Everything here only sequences kernels on the caller's stream.
#include <stdarg.h>

#include "common.cuh"

namespace plsb {

static thread_local std::string g_err;

void set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

namespace {

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

#define PLSB_HANDLE(h)                                                                 \
  PLSB_CHECK((h) != nullptr, PLSB_ERR_ARG, "null handle");                            \
  PLSB_CUDA(cudaSetDevice((h)->device))

// data matrix the contraction runs against
const double *perm_data(const plsb_ctx *h) {
  return h->lay.behavioral() ? h->Xcell.as<double>() : h->Xglob.as<double>();
}

// zero the operand rows [rows, rows_pad)
int zero_tail(double *A, long long rows, long long rows_pad, int lda, cudaStream_t st) {
  if (rows_pad > rows)
    PLSB_CUDA(cudaMemsetAsync(A + rows * lda, 0, sizeof(double) * (size_t)(rows_pad - rows) * lda,
                              st));
  return PLSB_OK;
}

// R (n*K rows, ldx) for resamples idx[0..n): operands + GEMM (+ column scaling).
// Behavioural operands are block structured -- the rows of cell g only touch
// that cell's rows of the data matrix -- so they are laid out grouped by cell
// (every 128-row GEMM tile inside one cell), the GEMM contracts each tile over
// its cell's rows only and a row map puts the result back in resample order.
int crosscov_chunk(plsb_ctx *h, const int32_t *idx, const double *yperm, int n, bool boot,
                   double *distrib, cudaStream_t st) {
  const Layout &l = h->lay;
  const bool grouped = l.behavioral() && l.J > 1;
  const bool scaled = boot && l.corr();
  const long long Mw = (long long)n * l.K, Mc = (long long)n * l.J;
  const long long cellpad_w = grouped ? round_up_ll((long long)n * l.T, GEMM_BM) : 0;
  const long long cellpad_c = grouped ? round_up_ll(n, GEMM_BM) : 0;
  const long long Mw_op = grouped ? cellpad_w * l.J : round_up_ll(Mw, GEMM_BM);
  const long long Mc_op = grouped ? cellpad_c * l.J : round_up_ll(Mc, GEMM_BM);
  const long long Mw_pad = round_up_ll(Mw, GEMM_BM), Mc_pad = round_up_ll(Mc, GEMM_BM);
  PLSB_CHECK(Mw_op < (1ll << 31) - 1, PLSB_ERR_ARG, "chunk of %d resamples is too large", n);
  PLSB_TRY(h->A.ensure(sizeof(double) * (size_t)Mw_op * l.S_pad));
  PLSB_TRY(h->R.ensure(sizeof(double) * (size_t)Mw_pad * l.ldx));
  int *map_w = nullptr, *map_c = nullptr;
  int4 *kr_w = nullptr, *kr_c = nullptr;
  if (grouped) {
    const size_t n_int = (size_t)Mw_op + (size_t)Mc_op;
    const size_t n_kr = (size_t)(Mw_op + Mc_op) / GEMM_BM;
    PLSB_TRY(h->maps.ensure(sizeof(int4) * n_kr + sizeof(int) * n_int));
    kr_w = h->maps.as<int4>();
    kr_c = kr_w + Mw_op / GEMM_BM;
    map_w = reinterpret_cast<int *>(kr_c + Mc_op / GEMM_BM);
    map_c = map_w + Mw_op;
    PLSB_CUDA(cudaMemsetAsync(h->A.p, 0, sizeof(double) * (size_t)Mw_op * l.S_pad, st));
    PLSB_TRY(launch_build_maps(h, n, l.T, l.K, cellpad_w, map_w, kr_w, st));
  } else {
    PLSB_TRY(zero_tail(h->A.as<double>(), Mw, Mw_op, l.S_pad, st));
  }
  // column scales of the resampled X: one dedicated pass (colstats.cu) or, when a cell
  // does not fit its shared-memory tile, two count-operand GEMMs + colscale
  bool fused_scale = false;
  if (scaled) {
    PLSB_TRY(h->S1.ensure(sizeof(double) * (size_t)Mc_pad * l.ldx));
    PLSB_TRY(launch_colstats(h, idx, n, h->S1.as<double>(), &fused_scale, st));
  }
  if (scaled && !fused_scale) {
    PLSB_TRY(h->Ac.ensure(sizeof(double) * (size_t)Mc_op * l.S_pad));
    PLSB_TRY(h->S2.ensure(sizeof(double) * (size_t)Mc_pad * l.ldx));
    if (grouped) {
      PLSB_CUDA(cudaMemsetAsync(h->Ac.p, 0, sizeof(double) * (size_t)Mc_op * l.S_pad, st));
      PLSB_TRY(launch_build_maps(h, n, 1, l.J, cellpad_c, map_c, kr_c, st));
    } else {
      PLSB_TRY(zero_tail(h->Ac.as<double>(), Mc, Mc_op, l.S_pad, st));
    }
  }
  PLSB_TRY(launch_build(h, boot ? BUILD_BOOT : BUILD_PLAIN, idx, yperm, n, h->A.as<double>(),
                        scaled && !fused_scale ? h->Ac.as<double>() : nullptr, distrib, cellpad_w,
                        cellpad_c, st));
  GemmArgs g;
  g.lda = l.S_pad;
  g.ldx = l.ldx;
  g.N_pad = l.ldx;
  g.Kd = l.S_pad;
  g.k_valid = l.S;
  g.ldc = l.ldx;
  g.x_persistent = true;
  if (grouped) g.k_len = l.kr_max;
  if (scaled && !fused_scale) {
    g.A = h->Ac.as<double>();
    g.X = h->Xglob.as<double>();
    g.M_pad = (int)Mc_op;
    g.row_map = map_c;
    g.kranges = kr_c;
    g.C = h->S1.as<double>();
    g.a_counts = true;           // multiplicity rows
    PLSB_TRY(launch_gemm(h, g, st));
    g.C = h->S2.as<double>();
    g.square_b = true;
    PLSB_TRY(launch_gemm(h, g, st));
    g.square_b = false;
    g.a_counts = false;
    PLSB_TRY(launch_colscale(h, h->S1.as<double>(), h->S2.as<double>(), (int)Mc, l.ldx, st));
  }
  if (scaled) {
    g.scale = h->S1.as<double>();
    g.scale_div = l.T;
    g.scale_rows = (int)Mc_pad;
    g.lds = l.ldx;
  }
  g.A = h->A.as<double>();
  g.X = boot ? h->Xglob.as<double>() : perm_data(h);
  g.x_persistent = true;
  g.M_pad = (int)Mw_op;
  g.row_map = map_w;
  g.kranges = kr_w;
  g.C = h->R.as<double>();
  PLSB_TRY(launch_gemm(h, g, st));
  return PLSB_OK;
}

// C (M,N) = A (M,Kd) @ X (Kd,N), all dense row-major without padding
int dense_gemm(plsb_ctx *h, const double *d_A, const double *d_X, int M, int N, int Kd,
               double *d_C, cudaStream_t st) {
  const int M_pad = round_up(M, GEMM_BM), N_pad = round_up(N, GEMM_BN), K_pad = round_up(Kd, GEMM_BK);
  PLSB_TRY(h->A.ensure(sizeof(double) * (size_t)M_pad * K_pad));
  PLSB_TRY(h->misc.ensure(sizeof(double) * (size_t)K_pad * N_pad));
  PLSB_TRY(h->R.ensure(sizeof(double) * (size_t)M_pad * N_pad));
  PLSB_TRY(launch_pad_copy(h, d_A, M, Kd, h->A.as<double>(), M_pad, K_pad, st));
  PLSB_TRY(launch_pad_copy(h, d_X, Kd, N, h->misc.as<double>(), K_pad, N_pad, st));
  GemmArgs g;
  g.A = h->A.as<double>();
  g.lda = K_pad;
  g.X = h->misc.as<double>();
  g.ldx = N_pad;
  g.M_pad = M_pad;
  g.N_pad = N_pad;
  g.Kd = K_pad;
  g.C = h->R.as<double>();
  g.ldc = N_pad;
  PLSB_TRY(launch_gemm(h, g, st));
  return launch_unpad_copy(h, h->R.as<double>(), N_pad, M, N, d_C, st);
}

// resamples per chunk so that the stored matrices fit the workspace limit
int chunk_size(const plsb_ctx *h, bool boot, int count) {
  const Layout &l = h->lay;
  size_t per = sizeof(double) * ((size_t)l.K * l.ldx + (size_t)l.K * l.S_pad);
  if (boot && l.corr()) per += sizeof(double) * (2 * (size_t)l.J * l.ldx + (size_t)l.J * l.S_pad);
  per += sizeof(double) * 4 * (size_t)l.K * l.K;
  long long n = (long long)(h->ws_limit / per);
  // operand row counts are ints
  const long long cap = (long long)((1u << 31) - 1024) / std::max(l.K, 1);
  n = std::min(n, cap);
  if (n < 1) n = 1;
  return (int)std::min<long long>(n, count);
}

}  // namespace
}  // namespace plsb

using namespace plsb;

extern "C" {

int plsb_version(void) { return 100; }

const char *plsb_last_error(void) { return g_err.c_str(); }

int plsb_create(plsb_handle_t *out, int device) {
  PLSB_CHECK(out != nullptr, PLSB_ERR_ARG, "plsb_create: null output");
  int n_dev = 0;
  PLSB_CUDA(cudaGetDeviceCount(&n_dev));
  PLSB_CHECK(device >= 0 && device < n_dev, PLSB_ERR_ARG, "plsb_create: device %d of %d", device,
             n_dev);
  PLSB_CUDA(cudaSetDevice(device));
  plsb_ctx *h = new plsb_ctx();
  h->device = device;
  cudaDeviceProp prop;
  PLSB_CUDA(cudaGetDeviceProperties(&prop, device));
  h->sm_count = prop.multiProcessorCount;
  *out = h;
  return PLSB_OK;
}

int plsb_destroy(plsb_handle_t h) {
  if (!h) return PLSB_OK;
  cudaSetDevice(h->device);
  DevBuf *bufs[] = {&h->tables, &h->Xraw, &h->Xcell, &h->Xglob, &h->Y,    &h->Cmat,  &h->Uo,
                    &h->Vo,     &h->dorig, &h->Sx,   &h->norms, &h->A,    &h->Ac,    &h->R,
                    &h->S1,     &h->S2,   &h->G,     &h->H,     &h->M,    &h->lam,   &h->rowsq,
                    &h->part,   &h->misc, &h->idxall, &h->flags, &h->maps,  &h->UoT, &h->Kx, &h->rowmask, &h->pctl, &h->big, &h->VoT,
                    &h->aplanes, &h->ascale, &h->xplanes[0].img, &h->xplanes[0].scale,
                    &h->xplanes[1].img, &h->xplanes[1].scale, &h->xplanes[2].img,
                    &h->xplanes[2].scale, &h->xplanes_tmp.img,
                    &h->xplanes_tmp.scale};
  for (DevBuf *b : bufs) b->release();
  delete h;
  return PLSB_OK;
}

int plsb_set_workspace_limit(plsb_handle_t h, uint64_t bytes) {
  PLSB_CHECK(h != nullptr, PLSB_ERR_ARG, "null handle");
  PLSB_CHECK(bytes >= (1ull << 20), PLSB_ERR_ARG, "workspace limit below 1 MiB");
  h->ws_limit = bytes;
  return PLSB_OK;
}

int plsb_set_gemm_backend(plsb_handle_t h, int backend, int n_slices) {
  PLSB_CHECK(h != nullptr, PLSB_ERR_ARG, "null handle");
  PLSB_CHECK(backend == PLSB_GEMM_AUTO || backend == PLSB_GEMM_DMMA, PLSB_ERR_ARG,
             "plsb_set_gemm_backend: unknown backend %d", backend);
  PLSB_CHECK(n_slices == 0 || (n_slices >= 5 && n_slices <= 7), PLSB_ERR_ARG,
             "plsb_set_gemm_backend: n_slices must be 0, 5, 6 or 7");
  h->gemm_backend = backend;
  if (n_slices) h->gemm_slices = n_slices;
  return PLSB_OK;
}

int plsb_gemm_work(plsb_handle_t h, double *i8_macs, double *dmma_flops, int reset) {
  PLSB_CHECK(h != nullptr, PLSB_ERR_ARG, "null handle");
  if (i8_macs) *i8_macs = h->i8_macs;
  if (dmma_flops) *dmma_flops = h->dmma_flops;
  if (reset) h->i8_macs = h->dmma_flops = 0.0;
  return PLSB_OK;
}

int64_t plsb_launch_count(plsb_handle_t h) { return h ? h->launches : 0; }

int plsb_timing_enable(plsb_handle_t h, int on) {
  PLSB_CHECK(h != nullptr, PLSB_ERR_ARG, "null handle");
  h->timing = on != 0;
  return PLSB_OK;
}

int plsb_timing_classes(void) { return KC_COUNT; }

const char *plsb_timing_class_name(int cls) {
  static const char *names[KC_COUNT] = {"xcov_gemm", "build_operands", "gram_proj", "small_decomp",
                                        "accum_u",   "stats",          "indexgen",  "prep"};
  return (cls >= 0 && cls < KC_COUNT) ? names[cls] : "";
}

int plsb_timing_read(plsb_handle_t h, double *ms, int64_t *launches) {
  PLSB_HANDLE(h);
  PLSB_CHECK(ms && launches, PLSB_ERR_ARG, "plsb_timing_read: null output");
  for (int c = 0; c < KC_COUNT; ++c) {
    ms[c] = 0.0;
    launches[c] = 0;
  }
  for (auto &t : h->timed) {
    PLSB_CUDA(cudaEventSynchronize(t.b));
    float v = 0.f;
    PLSB_CUDA(cudaEventElapsedTime(&v, t.a, t.b));
    ms[t.cls] += v;
    launches[t.cls] += 1;
    cudaEventDestroy(t.a);
    cudaEventDestroy(t.b);
  }
  h->timed.clear();
  return PLSB_OK;
}

int plsb_configure(plsb_handle_t h, int mode, int S, int B, int T, int n_groups, const int *groups,
                   int n_cond, int mean_centering, int n_components) {
  PLSB_HANDLE(h);
  PLSB_CHECK(mode >= PLSB_BEHAVIORAL_CORR && mode <= PLSB_SIMPLS, PLSB_ERR_ARG,
             "plsb_configure: unknown mode %d", mode);
  PLSB_CHECK(S >= 2 && B >= 1 && n_groups >= 1 && n_cond >= 1 && groups != nullptr, PLSB_ERR_ARG,
             "plsb_configure: bad sizes S=%d B=%d n_groups=%d n_cond=%d", S, B, n_groups, n_cond);
  Layout l;
  l.mode = mode;
  l.S = S;
  l.B = B;
  l.n_groups = n_groups;
  l.n_cond = n_cond;
  l.mean_centering = mean_centering;
  l.n_components = n_components;
  int n_subj = 0;
  for (int g = 0; g < n_groups; ++g) {
    PLSB_CHECK(groups[g] >= 1, PLSB_ERR_ARG, "plsb_configure: empty group %d", g);
    l.groups.push_back(groups[g]);
    n_subj += groups[g];
  }
  l.n_subj = n_subj;
  PLSB_CHECK(n_subj * n_cond == S, PLSB_ERR_ARG,
             "plsb_configure: sum(groups) * n_cond = %d does not match the %d rows of X",
             n_subj * n_cond, S);
  l.J = n_groups * n_cond;
  if (l.behavioral()) {
    PLSB_CHECK(T >= 1, PLSB_ERR_ARG, "plsb_configure: T=%d", T);
    l.T = T;
    l.K = l.J * T;
    for (int g = 0; g < n_groups; ++g)
      PLSB_CHECK(groups[g] >= 2, PLSB_ERR_ARG,
                 "plsb_configure: group %d has a single subject (no variance)", g);
  } else if (l.simpls()) {
    PLSB_CHECK(T >= 1 && n_groups == 1 && n_cond == 1, PLSB_ERR_ARG,
               "plsb_configure: SIMPLS takes one group, one condition and T >= 1");
    PLSB_CHECK(n_components >= 1 && n_components <= std::min(S - 1, B), PLSB_ERR_ARG,
               "Provided `n_components` cannot be greater than %d", std::min(S - 1, B));
    l.T = T;
    l.K = n_components;
  } else {
    PLSB_CHECK(mean_centering >= 0 && mean_centering <= 2, PLSB_ERR_ARG,
               "Mean centering type must be in [0, 1, 2].");
    l.T = 1;
    l.K = l.J;
  }
  // compute.svd keeps min(K, B) latent variables (pyls/compute.py:36-50)
  l.L = std::min(l.K, B);
  l.S_pad = round_up(S, GEMM_BK);
  l.ldx = round_up(B, GEMM_BN);

  // layout tables
  std::vector<int> tab;
  int row = 0;
  for (int g = 0; g < n_groups; ++g)
    for (int c = 0; c < n_cond; ++c) {
      l.cell_start.push_back(row);
      row += groups[g];
    }
  l.cell_start.push_back(row);
  tab.insert(tab.end(), l.cell_start.begin(), l.cell_start.end());   // J+1
  for (int j = 0; j < l.J; ++j)
    for (int s = l.cell_start[j]; s < l.cell_start[j + 1]; ++s) tab.push_back(j);   // S
  for (int j = 0; j < l.J; ++j) tab.push_back(l.cell_start[j + 1] - l.cell_start[j]);   // J
  int acc = 0;
  for (int g = 0; g < n_groups; ++g) {
    tab.push_back(acc);
    acc += groups[g];
  }
  tab.push_back(acc);   // n_groups + 1
  while (tab.size() % 4) tab.push_back(0);   // keep the int4 table 16-byte aligned
  const size_t kr_off = tab.size();
  for (int j = 0; j < l.J; ++j) {
    // contraction range of cell j: even start, whole GEMM_BK chunks, inside [0, S_pad)
    int kb = l.cell_start[j] & ~1;
    const int nkc = cdiv(l.cell_start[j + 1] - kb, GEMM_BK);
    if (kb + nkc * GEMM_BK > l.S_pad) kb = l.S_pad - nkc * GEMM_BK;
    tab.push_back(kb);
    tab.push_back(kb + nkc * GEMM_BK);
    // k steps of 4 that touch the cell's rows: operand columns outside are structural zeros
    tab.push_back(kb + ((l.cell_start[j] - kb) & ~3));
    tab.push_back(kb + round_up(l.cell_start[j + 1] - kb, 4));
    l.kr_max = std::max(l.kr_max, nkc * GEMM_BK);
  }
  PLSB_TRY(h->tables.ensure(sizeof(int) * tab.size()));
  PLSB_CUDA(cudaMemcpy(h->tables.p, tab.data(), sizeof(int) * tab.size(), cudaMemcpyHostToDevice));
  h->d_cell_start = h->tables.as<int>();
  h->d_cell_of_row = h->d_cell_start + (l.J + 1);
  h->d_cell_n = h->d_cell_of_row + S;
  h->d_group_start = h->d_cell_n + l.J;
  h->d_cell_kr = reinterpret_cast<const int4 *>(h->tables.as<int>() + kr_off);

  if (l.mode == PLSB_MEANCENTERED) {
    // C (J,S): "cell mean minus centring mean" as a row operator
    // (pyls/compute.py:267-357)
    std::vector<double> C((size_t)l.J * S, 0.0);
    for (int j = 0; j < l.J; ++j) {
      const int g = j / n_cond, c = j % n_cond;
      const int nj = l.cell_start[j + 1] - l.cell_start[j];
      for (int s = l.cell_start[j]; s < l.cell_start[j + 1]; ++s) C[(size_t)j * S + s] += 1.0 / nj;
      if (mean_centering == 0) {
        const int r0 = l.cell_start[g * n_cond], r1 = l.cell_start[(g + 1) * n_cond];
        for (int s = r0; s < r1; ++s) C[(size_t)j * S + s] -= 1.0 / (r1 - r0);
      } else if (mean_centering == 1) {
        for (int g2 = 0; g2 < n_groups; ++g2) {
          const int j2 = g2 * n_cond + c;
          const int n2 = l.cell_start[j2 + 1] - l.cell_start[j2];
          for (int s = l.cell_start[j2]; s < l.cell_start[j2 + 1]; ++s)
            C[(size_t)j * S + s] -= 1.0 / ((double)n_groups * n2);
        }
      } else {
        for (int s = 0; s < S; ++s) C[(size_t)j * S + s] -= 1.0 / S;
      }
    }
    PLSB_TRY(h->Cmat.ensure(sizeof(double) * C.size()));
    PLSB_CUDA(cudaMemcpy(h->Cmat.p, C.data(), sizeof(double) * C.size(), cudaMemcpyHostToDevice));
  }
  h->lay = l;
  h->configured = true;
  h->has_data = h->has_original = false;
  return PLSB_OK;
}

int plsb_set_data(plsb_handle_t h, const double *d_X, const double *d_Y, void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(h->configured, PLSB_ERR_STATE, "plsb_set_data before plsb_configure");
  const Layout &l = h->lay;
  cudaStream_t st = as_stream(stream);
  PLSB_CHECK(d_X != nullptr, PLSB_ERR_ARG, "plsb_set_data: null X");
  PLSB_CHECK(!(l.behavioral() || l.simpls()) || d_Y != nullptr, PLSB_ERR_ARG,
             "plsb_set_data: null Y");
  const size_t bytes = sizeof(double) * (size_t)l.S_pad * l.ldx;
  h->has_rowmask = false;
  ++h->data_epoch;   // cached digit planes of the data matrices are stale
  if (l.simpls()) {
    // X and Y arrive column-centred (pyls/types/regression.py:395-396).  Kraw = X X^T (S,S)
    // is the only B-sized object the component loops need.
    PLSB_TRY(h->Xraw.ensure(bytes));
    PLSB_TRY(launch_pad_copy(h, d_X, l.S, l.B, h->Xraw.as<double>(), l.S_pad, l.ldx, st));
    PLSB_TRY(h->Y.ensure(sizeof(double) * (size_t)l.S * l.T));
    PLSB_CUDA(cudaMemcpyAsync(h->Y.p, d_Y, sizeof(double) * (size_t)l.S * l.T,
                              cudaMemcpyDeviceToDevice, st));
    PLSB_TRY(h->Xglob.ensure(sizeof(double) * (size_t)l.S * l.B));
    PLSB_TRY(launch_transpose(h, d_X, l.S, l.B, l.B, h->Xglob.as<double>(), st));
    PLSB_TRY(h->Cmat.ensure(sizeof(double) * (size_t)l.S * l.S));
    PLSB_TRY(dense_gemm(h, d_X, h->Xglob.as<double>(), l.S, l.S, l.B, h->Cmat.as<double>(), st));
    h->has_data = true;
    h->has_original = false;
    return PLSB_OK;
  }
  PLSB_TRY(h->Xraw.ensure(bytes));
  PLSB_TRY(h->Xglob.ensure(bytes));
  PLSB_TRY(launch_pad_copy(h, d_X, l.S, l.B, h->Xraw.as<double>(), l.S_pad, l.ldx, st));
  PLSB_CUDA(cudaMemsetAsync(h->Xglob.p, 0, bytes, st));
  double *xcell = nullptr;
  if (l.behavioral()) {
    PLSB_TRY(h->Xcell.ensure(bytes));
    PLSB_CUDA(cudaMemsetAsync(h->Xcell.p, 0, bytes, st));
    xcell = h->Xcell.as<double>();
    PLSB_TRY(h->Y.ensure(sizeof(double) * (size_t)l.S * l.T));
    PLSB_CUDA(cudaMemcpyAsync(h->Y.p, d_Y, sizeof(double) * (size_t)l.S * l.T,
                              cudaMemcpyDeviceToDevice, st));
  }
  PLSB_TRY(launch_prep_cells(h, h->Xraw.as<double>(), xcell, h->Xglob.as<double>(), st));
  h->has_data = true;
  h->has_original = false;
  h->has_kx = false;
  return PLSB_OK;
}

int plsb_set_original(plsb_handle_t h, const double *d_U, const double *d_d, const double *d_V,
                      void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(h->has_data, PLSB_ERR_STATE, "plsb_set_original before plsb_set_data");
  const Layout &l = h->lay;
  cudaStream_t st = as_stream(stream);
  PLSB_CHECK(d_U && d_d && d_V, PLSB_ERR_ARG, "plsb_set_original: null argument");
  PLSB_TRY(h->Uo.ensure(sizeof(double) * (size_t)l.B * l.L));
  PLSB_TRY(h->Vo.ensure(sizeof(double) * (size_t)l.K * l.L));
  PLSB_TRY(h->dorig.ensure(sizeof(double) * l.L));
  PLSB_TRY(h->Sx.ensure(sizeof(double) * (size_t)l.S * l.L));
  PLSB_TRY(h->norms.ensure(sizeof(double) * l.L));
  if (d_U != h->Uo.p)
    PLSB_CUDA(cudaMemcpyAsync(h->Uo.p, d_U, sizeof(double) * (size_t)l.B * l.L,
                              cudaMemcpyDeviceToDevice, st));
  if (d_V != h->Vo.p)
    PLSB_CUDA(cudaMemcpyAsync(h->Vo.p, d_V, sizeof(double) * (size_t)l.K * l.L,
                              cudaMemcpyDeviceToDevice, st));
  if (d_d != h->dorig.p)
    PLSB_CUDA(cudaMemcpyAsync(h->dorig.p, d_d, sizeof(double) * l.L, cudaMemcpyDeviceToDevice, st));
  // transposed zero-padded copy (L, ldx) for gram_proj's TMA row copies
  PLSB_TRY(h->UoT.ensure(sizeof(double) * (size_t)l.L * l.ldx));
  PLSB_TRY(launch_transpose_pad(h, h->Uo.as<double>(), l.B, l.L, h->UoT.as<double>(), l.ldx, st));
  if (l.tall()) {
    // V_orig^T (L, ldk) zero padded: projection block of the feature-side passes
    const long long ldk = round_up(l.K, GEMM_BN);
    PLSB_TRY(h->VoT.ensure(sizeof(double) * (size_t)l.L * ldk));
    PLSB_TRY(launch_transpose_pad(h, h->Vo.as<double>(), l.K, l.L, h->VoT.as<double>(), ldk, st));
  }
  // Sx = X @ normalize(U_orig)   (pyls/types/behavioral.py:78, meancentered.py:98)
  PLSB_TRY(launch_colnorm(h, h->Uo.as<double>(), l.B, l.L, h->norms.as<double>(), st));
  PLSB_TRY(launch_xproj(h, h->Xraw.as<double>(), l.ldx, l.S, l.B, h->Uo.as<double>(), l.L,
                        h->norms.as<double>(), h->Sx.as<double>(), st));
  h->has_original = true;
  return PLSB_OK;
}

// ---- tall analyses (K > B): see tall.cu ---------------------------------------------

// R' = R^T of the n resamples whose cross-covariances sit in h->R: (n * B rows, ldk) in h->S2
static int tall_transpose(plsb_ctx *h, int n, long long *ldk_out, cudaStream_t st) {
  const Layout &l = h->lay;
  const long long ldk = round_up(l.K, GEMM_BN);
  PLSB_TRY(h->S2.ensure(sizeof(double) * (size_t)n * l.B * ldk));
  *ldk_out = ldk;
  return launch_transpose_batch(h, h->R.as<double>(), l.K, l.B, l.ldx, (long long)l.K * l.ldx,
                                h->S2.as<double>(), ldk, (long long)l.B * ldk, n, st);
}

static int tall_chunk(const plsb_ctx *h, bool boot, int count) {
  const Layout &l = h->lay;
  const size_t ldk = round_up(l.K, GEMM_BN);
  size_t per = sizeof(double) * ((size_t)l.K * l.ldx + (size_t)l.K * l.S_pad + (size_t)l.B * ldk +
                                 4 * (size_t)l.B * l.B);
  if (boot && l.corr()) per += sizeof(double) * (2 * (size_t)l.J * l.ldx + (size_t)l.J * l.S_pad);
  long long n = (long long)(h->ws_limit / per);
  n = std::min<long long>(n, ((1ll << 31) - 1024) / std::max(l.K, 1));
  return (int)std::max<long long>(1, std::min<long long>(n, count));
}

static int tall_decompose(plsb_ctx *h, double *d_U, double *d_d, double *d_V, cudaStream_t st) {
  const Layout &l = h->lay;
  const int Bf = l.B, L = l.L;   // L == Bf
  PLSB_TRY(crosscov_chunk(h, nullptr, nullptr, 1, false, nullptr, st));
  long long ldk = 0;
  PLSB_TRY(tall_transpose(h, 1, &ldk, st));
  const double *Rt = h->S2.as<double>();
  const size_t bb = (size_t)Bf * Bf, kl = (size_t)l.K * L;
  PLSB_TRY(h->G.ensure(sizeof(double) * 2 * bb));
  PLSB_TRY(h->lam.ensure(sizeof(double) * Bf));
  PLSB_TRY(h->misc.ensure(sizeof(double) * 2 * kl));
  double *Gp = h->G.as<double>(), *W = Gp + bb, *lam = h->lam.as<double>();
  double *vraw = h->misc.as<double>(), *vsq = vraw + kl;
  PLSB_TRY(launch_gram_proj(h, Rt, ldk, 1, Bf, nullptr, 0, Gp, nullptr, st));
  PLSB_TRY(launch_sym_eig(h, Gp, 1, Bf, W, lam, 0, st));
  // V d = R U = R'^T W  (K x L), then column norms / signs: sklearn's svd_flip decides on
  // the FIRST factor randomized_svd returns, which is the K-side one in this orientation
  PLSB_CUDA(cudaMemsetAsync(vraw, 0, sizeof(double) * 2 * kl, st));
  const int ldm = accum_ldm(L);
  PLSB_TRY(h->M.ensure(sizeof(double) * (size_t)Bf * ldm));
  PLSB_TRY(launch_pad_copy(h, W, Bf, L, h->M.as<double>(), Bf, ldm, st));
  PLSB_TRY(launch_accum_u(h, Rt, ldk, 1, Bf, l.K, h->M.as<double>(), L, vraw, vsq, st));
  PLSB_CUDA(cudaMemcpyAsync(d_U, W, sizeof(double) * bb, cudaMemcpyDeviceToDevice, st));
  PLSB_TRY(launch_normalize_flip(h, vraw, l.K, L, lam, d_V, d_U, Bf, d_d, st));
  return PLSB_OK;
}

// permutations of a tall analysis: d_dperm (count, L)
static int tall_run_perms(plsb_ctx *h, const int32_t *d_idx, const double *d_yperm, int count,
                          int rotate, double *d_dperm, cudaStream_t st) {
  const Layout &l = h->lay;
  const int Bf = l.B, L = l.L;
  const size_t ystride = (size_t)l.S * l.T, bb = (size_t)Bf * Bf;
  const int chunk = tall_chunk(h, false, count);
  for (int off = 0; off < count; off += chunk) {
    const int n = std::min(chunk, count - off);
    PLSB_TRY(crosscov_chunk(h, d_idx ? d_idx + (size_t)off * l.S : nullptr,
                            d_yperm ? d_yperm + off * ystride : nullptr, n, false, nullptr, st));
    long long ldk = 0;
    PLSB_TRY(tall_transpose(h, n, &ldk, st));
    PLSB_TRY(h->G.ensure(sizeof(double) * (size_t)n * bb));
    if (!rotate) {
      PLSB_TRY(launch_gram_proj(h, h->S2.as<double>(), ldk, n, Bf, nullptr, 0, h->G.as<double>(),
                                nullptr, st));
      PLSB_TRY(launch_sym_eig(h, h->G.as<double>(), n, Bf, nullptr, d_dperm + (size_t)off * L, 1,
                              st));
      continue;
    }
    PLSB_TRY(h->H.ensure(sizeof(double) * (size_t)n * Bf * L));
    PLSB_TRY(h->M.ensure(sizeof(double) * (size_t)n * Bf * L));
    PLSB_TRY(launch_gram_proj(h, h->S2.as<double>(), ldk, n, Bf, h->VoT.as<double>(), L,
                              h->G.as<double>(), h->H.as<double>(), st));
    PLSB_TRY(launch_small_decomp(h, h->G.as<double>(), h->H.as<double>(), n, Bf, L,
                                 h->dorig.as<double>(), h->M.as<double>(), L, nullptr, st));
    PLSB_TRY(launch_quadform_sqrt(h, h->G.as<double>(), h->M.as<double>(), n, Bf, L,
                                  d_dperm + (size_t)off * L, st));
  }
  return PLSB_OK;
}

static int tall_run_boots(plsb_ctx *h, const int32_t *d_idx, int count, double *d_distrib,
                          double *d_usum, double *d_usquare, cudaStream_t st) {
  const Layout &l = h->lay;
  const int Bf = l.B, L = l.L;
  const size_t bb = (size_t)Bf * Bf;
  const int chunk = tall_chunk(h, true, count);
  for (int off = 0; off < count; off += chunk) {
    const int n = std::min(chunk, count - off);
    PLSB_TRY(crosscov_chunk(h, d_idx + (size_t)off * l.S, nullptr, n, true,
                            d_distrib ? d_distrib + (size_t)off * l.K * L : nullptr, st));
    long long ldk = 0;
    PLSB_TRY(tall_transpose(h, n, &ldk, st));
    PLSB_TRY(h->G.ensure(sizeof(double) * (size_t)n * bb));
    PLSB_TRY(h->H.ensure(sizeof(double) * (size_t)n * bb));
    PLSB_TRY(h->M.ensure(sizeof(double) * (size_t)n * bb));
    PLSB_TRY(h->lam.ensure(sizeof(double) * (size_t)n * Bf));
    PLSB_TRY(launch_gram_proj(h, h->S2.as<double>(), ldk, n, Bf, nullptr, 0, h->G.as<double>(),
                              nullptr, st));
    // U_boot (eigenvectors), d^2; then compute.procrustes(U_orig, U_boot, d) = U_boot d Q
    PLSB_TRY(launch_sym_eig(h, h->G.as<double>(), n, Bf, h->H.as<double>(), h->lam.as<double>(),
                            0, st));
    PLSB_TRY(launch_rotation_tall(h, h->Uo.as<double>(), h->H.as<double>(), h->lam.as<double>(),
                                  n, Bf, h->dorig.as<double>(), h->M.as<double>(), st));
    PLSB_TRY(launch_accum_small(h, h->M.as<double>(), n, (long long)bb, d_usum, d_usquare, st));
  }
  return PLSB_OK;
}

int plsb_decompose(plsb_handle_t h, double *d_U, double *d_d, double *d_V, void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(h->has_data, PLSB_ERR_STATE, "plsb_decompose before plsb_set_data");
  const Layout &l = h->lay;
  cudaStream_t st = as_stream(stream);
  PLSB_CHECK(d_U && d_d && d_V, PLSB_ERR_ARG, "plsb_decompose: null output");
  if (l.tall()) {
    PLSB_TRY(tall_decompose(h, d_U, d_d, d_V, st));
    return plsb_set_original(h, d_U, d_d, d_V, stream);
  }
  PLSB_TRY(crosscov_chunk(h, nullptr, nullptr, 1, false, nullptr, st));
  const size_t kk = (size_t)l.K * l.K, bl = (size_t)l.B * l.L;
  PLSB_TRY(h->G.ensure(sizeof(double) * 4 * kk));
  PLSB_TRY(h->lam.ensure(sizeof(double) * 2 * l.K));
  PLSB_TRY(h->misc.ensure(sizeof(double) * 2 * bl));
  double *G1 = h->G.as<double>(), *V1 = G1 + kk, *G2 = V1 + kk, *V2 = G2 + kk;
  double *lam1 = h->lam.as<double>(), *lam2 = lam1 + l.K;
  double *uraw = h->misc.as<double>(), *usq = uraw + bl;
  // pass 1: eigenvectors of the Gram matrix R R^T (accurate to eps * cond(R)^2)
  PLSB_TRY(launch_gram_proj(h, h->R.as<double>(), l.ldx, 1, l.K, nullptr, 0, G1, nullptr, st));
  PLSB_TRY(launch_sym_eig(h, G1, 1, l.K, V1, lam1, 0, st));
  // pass 2 (refinement): W = R^T V1 is computed from R itself, its Gram matrix
  // W^T W is diagonal up to pass 1's error and graded, so a second Jacobi
  // recovers the small singular directions to working precision
  PLSB_CUDA(cudaMemsetAsync(uraw, 0, sizeof(double) * 2 * bl, st));
  const int ldm = accum_ldm(l.L);
  PLSB_TRY(h->M.ensure(sizeof(double) * (size_t)l.K * ldm));
  double *Mp = h->M.as<double>();   // rotation in accum_u's padded layout
  PLSB_TRY(launch_pad_copy(h, V1, l.K, l.L, Mp, l.K, ldm, st));
  PLSB_TRY(launch_accum_u(h, h->R.as<double>(), l.ldx, 1, l.K, l.B, Mp, l.L, uraw, usq, st));
  PLSB_TRY(launch_colgram(h, uraw, l.B, l.K, G2, st));
  PLSB_TRY(launch_sym_eig(h, G2, 1, l.K, V2, lam2, 0, st));
  PLSB_TRY(launch_matmul_small(h, V1, V2, l.K, d_V, st));
  // U d = R^T V, then column norms / signs (sklearn svd_flip on the B-side factor)
  PLSB_CUDA(cudaMemsetAsync(uraw, 0, sizeof(double) * 2 * bl, st));
  PLSB_TRY(launch_pad_copy(h, d_V, l.K, l.L, Mp, l.K, ldm, st));
  PLSB_TRY(launch_accum_u(h, h->R.as<double>(), l.ldx, 1, l.K, l.B, Mp, l.L, uraw, usq, st));
  PLSB_TRY(launch_normalize_flip(h, uraw, l.B, l.L, lam2, d_U, d_V, l.K, d_d, st));
  return plsb_set_original(h, d_U, d_d, d_V, stream);
}

int plsb_project_scores(plsb_handle_t h, const double *d_U, int L, double *d_out, void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(h->has_data, PLSB_ERR_STATE, "plsb_project_scores before plsb_set_data");
  PLSB_CHECK(d_U && d_out && L >= 1, PLSB_ERR_ARG, "plsb_project_scores: bad argument");
  const Layout &l = h->lay;
  return launch_xproj(h, h->Xraw.as<double>(), l.ldx, l.S, l.B, d_U, L, nullptr, d_out,
                      as_stream(stream));
}

int plsb_gen_perm_indices(plsb_handle_t h, uint64_t seed, int64_t first, int count, int32_t *d_idx,
                          int *h_n_exhausted, void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(h->configured, PLSB_ERR_STATE, "index generation before plsb_configure");
  return gen_indices(h, false, seed, first, count, d_idx, h_n_exhausted, as_stream(stream));
}

int plsb_gen_boot_indices(plsb_handle_t h, uint64_t seed, int64_t first, int count, int32_t *d_idx,
                          int *h_n_exhausted, void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(h->configured, PLSB_ERR_STATE, "index generation before plsb_configure");
  return gen_indices(h, true, seed, first, count, d_idx, h_n_exhausted, as_stream(stream));
}

int plsb_gen_split_masks(plsb_handle_t h, uint64_t seed, int64_t first, int count, int n_split,
                         double train_fraction, int32_t *d_masks, int *h_n_exhausted,
                         void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(h->configured, PLSB_ERR_STATE, "index generation before plsb_configure");
  PLSB_CHECK(d_masks != nullptr, PLSB_ERR_ARG, "plsb_gen_split_masks: null output");
  return gen_split_masks(h, seed, first, count, n_split, train_fraction, d_masks, h_n_exhausted,
                         as_stream(stream));
}

int plsb_gen_gaussian_tables(plsb_handle_t h, int64_t first, int count, int T, double *d_omega,
                             void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(d_omega && T >= 1, PLSB_ERR_ARG, "plsb_gen_gaussian_tables: bad argument");
  return gen_gaussian_tables(h, first, count, T * 11, d_omega, as_stream(stream));
}

int plsb_crosscov(plsb_handle_t h, const int32_t *d_idx, int count, int bootstrap, double *d_R,
                  void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(h->has_data, PLSB_ERR_STATE, "plsb_crosscov before plsb_set_data");
  PLSB_CHECK(d_R != nullptr && count >= 0, PLSB_ERR_ARG, "plsb_crosscov: bad argument");
  PLSB_CHECK(d_idx != nullptr || count == 1, PLSB_ERR_ARG, "plsb_crosscov: null index table");
  const Layout &l = h->lay;
  cudaStream_t st = as_stream(stream);
  const int chunk = chunk_size(h, bootstrap != 0, count);
  for (int off = 0; off < count; off += chunk) {
    const int n = std::min(chunk, count - off);
    PLSB_TRY(crosscov_chunk(h, d_idx ? d_idx + (size_t)off * l.S : nullptr, nullptr, n,
                            bootstrap != 0,
                            nullptr, st));
    PLSB_TRY(launch_unpad_copy(h, h->R.as<double>(), l.ldx, n * l.K, l.B,
                               d_R + (size_t)off * l.K * l.B, st));
  }
  return PLSB_OK;
}

// permutations from index vectors (d_idx) or from pre-permuted Y matrices (d_yperm)
static int run_perms_impl(plsb_ctx *h, const int32_t *d_idx, const double *d_yperm, int count,
                          int rotate, double *d_dperm, cudaStream_t st) {
  PLSB_CHECK(h->has_data, PLSB_ERR_STATE, "plsb_run_perms before plsb_set_data");
  PLSB_CHECK(!rotate || h->has_original, PLSB_ERR_STATE,
             "plsb_run_perms(rotate) before the original decomposition is set");
  const Layout &l = h->lay;
  const size_t ystride = (size_t)l.S * l.T;
  if (l.tall()) return tall_run_perms(h, d_idx, d_yperm, count, rotate, d_dperm, st);
  if (rotate) {
    // |R^T v_j| for every original y-weight v_j: A = V^T-weighted operand, row sums of squares
    const size_t per = sizeof(double) * (size_t)l.L * l.S_pad;
    int chunk = (int)std::min<long long>(
        count, std::max<long long>(1, (long long)(h->ws_limit / per)));
    chunk = (int)std::min<long long>(chunk, ((1ll << 31) - 1024) / l.L);
    for (int off = 0; off < count; off += chunk) {
      const int n = std::min(chunk, count - off);
      const long long M = (long long)n * l.L, M_pad = round_up_ll(M, GEMM_BM);
      PLSB_TRY(h->A.ensure(sizeof(double) * (size_t)M_pad * l.S_pad));
      PLSB_TRY(zero_tail(h->A.as<double>(), M, M_pad, l.S_pad, st));
      PLSB_TRY(launch_build(h, BUILD_ROT, d_idx ? d_idx + (size_t)off * l.S : nullptr,
                            d_yperm ? d_yperm + off * ystride : nullptr, n, h->A.as<double>(),
                            nullptr, nullptr, 0, 0, st));
      GemmArgs g;
      g.A = h->A.as<double>();
      g.lda = l.S_pad;
      g.X = perm_data(h);
      g.x_persistent = true;
      g.ldx = l.ldx;
      g.M_pad = (int)M_pad;
      g.N_pad = l.ldx;
      g.Kd = l.S_pad;
      g.k_valid = l.S;
      const int n_splits = gemm_rowsq_splits(h, g);
      PLSB_TRY(h->rowsq.ensure(sizeof(double) * (size_t)n_splits * M_pad));
      g.rowsq = h->rowsq.as<double>();
      g.n_splits = n_splits;
      PLSB_TRY(launch_gemm(h, g, st));
      PLSB_TRY(launch_finish_rowsq(h, h->rowsq.as<double>(), n_splits, (int)M_pad, (int)M,
                                   d_dperm + (size_t)off * l.L, st));
    }
    return PLSB_OK;
  }
  // un-rotated: singular values of every R from the eigenvalues of R R^T
  const int chunk = chunk_size(h, false, count);
  for (int off = 0; off < count; off += chunk) {
    const int n = std::min(chunk, count - off);
    PLSB_TRY(crosscov_chunk(h, d_idx ? d_idx + (size_t)off * l.S : nullptr,
                            d_yperm ? d_yperm + off * ystride : nullptr, n, false, nullptr, st));
    PLSB_TRY(h->G.ensure(sizeof(double) * (size_t)n * l.K * l.K));
    PLSB_TRY(launch_gram_proj(h, h->R.as<double>(), l.ldx, n, l.K, nullptr, 0,
                              h->G.as<double>(), nullptr, st));
    PLSB_TRY(launch_sym_eig(h, h->G.as<double>(), n, l.K, nullptr, d_dperm + (size_t)off * l.L, 1,
                            st));
  }
  return PLSB_OK;
}

int plsb_run_perms(plsb_handle_t h, const int32_t *d_idx, int count, int rotate, double *d_dperm,
                   void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(d_idx && d_dperm && count >= 0, PLSB_ERR_ARG, "plsb_run_perms: bad argument");
  return run_perms_impl(h, d_idx, nullptr, count, rotate, d_dperm, as_stream(stream));
}

// Kx = Xp Xp^T (S_pad x S_pad, pitch round_up(S_pad, 128)) of the permutation data matrix
static int ensure_gram(plsb_ctx *h, cudaStream_t st) {
  if (h->has_kx) return PLSB_OK;
  const Layout &l = h->lay;
  const int N_pad = round_up(l.S_pad, GEMM_BN);
  PLSB_TRY(h->Kx.ensure(sizeof(double) * (size_t)l.S_pad * N_pad));
  PLSB_TRY(h->Ac.ensure(sizeof(double) * (size_t)l.ldx * l.S_pad));
  PLSB_TRY(h->S1.ensure(sizeof(double) * (size_t)l.S_pad * l.S_pad));
  // Xp^T (ldx, S_pad), then the dense product (padding columns of Xp are zero)
  PLSB_TRY(launch_transpose(h, perm_data(h), l.S_pad, l.ldx, l.ldx, h->Ac.as<double>(), st));
  PLSB_TRY(dense_gemm(h, perm_data(h), h->Ac.as<double>(), l.S_pad, l.S_pad, l.ldx,
                      h->S1.as<double>(), st));
  PLSB_TRY(launch_pad_copy(h, h->S1.as<double>(), l.S_pad, l.S_pad, h->Kx.as<double>(), l.S_pad,
                           N_pad, st));
  h->has_kx = true;
  return PLSB_OK;
}

int plsb_run_perms_gram(plsb_handle_t h, const int32_t *d_idx, int count, double *d_dperm,
                        void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(h->has_data && h->has_original && !h->lay.simpls(), PLSB_ERR_STATE,
             "plsb_run_perms_gram needs data and the original decomposition");
  PLSB_CHECK(!h->lay.tall(), PLSB_ERR_ARG,
             "plsb_run_perms_gram: the sample-space identity needs K <= B");
  PLSB_CHECK(d_idx && d_dperm && count >= 0, PLSB_ERR_ARG, "plsb_run_perms_gram: bad argument");
  const Layout &l = h->lay;
  cudaStream_t st = as_stream(stream);
  PLSB_TRY(ensure_gram(h, st));
  const int N_pad = round_up(l.S_pad, GEMM_BN);
  const size_t per = sizeof(double) * (size_t)l.L * (l.S_pad + N_pad);
  int chunk = (int)std::min<long long>(count,
                                       std::max<long long>(1, (long long)(h->ws_limit / per)));
  chunk = (int)std::min<long long>(chunk, ((1ll << 31) - 1024) / l.L);
  for (int off = 0; off < count; off += chunk) {
    const int n = std::min(chunk, count - off);
    const long long M = (long long)n * l.L, M_pad = round_up_ll(M, GEMM_BM);
    PLSB_TRY(h->A.ensure(sizeof(double) * (size_t)M_pad * l.S_pad));
    PLSB_TRY(h->R.ensure(sizeof(double) * (size_t)M_pad * N_pad));
    PLSB_TRY(zero_tail(h->A.as<double>(), M, M_pad, l.S_pad, st));
    PLSB_TRY(launch_build(h, BUILD_ROT, d_idx + (size_t)off * l.S, nullptr, n, h->A.as<double>(),
                          nullptr, nullptr, 0, 0, st));
    GemmArgs g;
    g.A = h->A.as<double>();
    g.lda = l.S_pad;
    g.X = h->Kx.as<double>();
    g.ldx = N_pad;
    g.M_pad = (int)M_pad;
    g.N_pad = N_pad;
    g.Kd = l.S_pad;
    g.k_valid = l.S;
    g.k_len = l.S_pad;
    g.C = h->R.as<double>();
    g.ldc = N_pad;
    PLSB_TRY(launch_gemm(h, g, st));
    PLSB_TRY(launch_rowdot_sqrt(h, h->R.as<double>(), N_pad, h->A.as<double>(), l.S_pad, l.S_pad,
                                M, d_dperm + (size_t)off * l.L, st));
  }
  return PLSB_OK;
}

int plsb_run_perms_prepermuted(plsb_handle_t h, const double *d_Yperm, int count, int rotate,
                               double *d_dperm, void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(d_Yperm && d_dperm && count >= 0, PLSB_ERR_ARG,
             "plsb_run_perms_prepermuted: bad argument");
  PLSB_CHECK(h->lay.behavioral(), PLSB_ERR_ARG,
             "plsb_run_perms_prepermuted: only behavioural analyses have a Y matrix to permute");
  return run_perms_impl(h, nullptr, d_Yperm, count, rotate, d_dperm, as_stream(stream));
}

int plsb_run_boots(plsb_handle_t h, const int32_t *d_idx, int count, double *d_distrib,
                   double *d_usum, double *d_usquare, void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(h->has_data && h->has_original, PLSB_ERR_STATE,
             "plsb_run_boots before data and original decomposition are set");
  PLSB_CHECK(d_idx && d_usum && d_usquare && count >= 0, PLSB_ERR_ARG,
             "plsb_run_boots: bad argument");
  const Layout &l = h->lay;
  cudaStream_t st = as_stream(stream);
  if (l.tall()) return tall_run_boots(h, d_idx, count, d_distrib, d_usum, d_usquare, st);
  const int chunk = chunk_size(h, true, count);
  for (int off = 0; off < count; off += chunk) {
    const int n = std::min(chunk, count - off);
    PLSB_TRY(crosscov_chunk(h, d_idx + (size_t)off * l.S, nullptr, n, true,
                            d_distrib ? d_distrib + (size_t)off * l.K * l.L : nullptr, st));
    PLSB_TRY(h->G.ensure(sizeof(double) * (size_t)n * l.K * l.K));
    PLSB_TRY(h->H.ensure(sizeof(double) * (size_t)n * l.K * l.L));
    const int ldm = accum_ldm(l.L);
    PLSB_TRY(h->M.ensure(sizeof(double) * (size_t)n * l.K * ldm));
    PLSB_TRY(launch_gram_proj(h, h->R.as<double>(), l.ldx, n, l.K, h->UoT.as<double>(), l.L,
                              h->G.as<double>(), h->H.as<double>(), st));
    PLSB_TRY(launch_small_decomp(h, h->G.as<double>(), h->H.as<double>(), n, l.K, l.L,
                                 h->dorig.as<double>(), h->M.as<double>(), ldm, nullptr, st));
    PLSB_TRY(launch_accum_u(h, h->R.as<double>(), l.ldx, n, l.K, l.B, h->M.as<double>(), l.L,
                            d_usum, d_usquare, st));
  }
  return PLSB_OK;
}

int plsb_boot_distrib(plsb_handle_t h, const int32_t *d_idx, int count, double *d_distrib,
                      void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(h->has_data && h->has_original && !h->lay.simpls(), PLSB_ERR_STATE,
             "plsb_boot_distrib before data and original decomposition are set");
  PLSB_CHECK(d_idx && d_distrib && count >= 0, PLSB_ERR_ARG, "plsb_boot_distrib: bad argument");
  return launch_build(h, BUILD_BOOT, d_idx, nullptr, count, nullptr, nullptr, d_distrib, 0, 0,
                      as_stream(stream));
}

int plsb_boot_chunk(plsb_handle_t h, int count) {
  if (!h || !h->configured || count <= 0) return 0;
  return chunk_size(h, true, count);
}

int plsb_perm_pvals(plsb_handle_t h, const double *d_dperm, int count, int L, const double *d_dorig,
                    double *d_pvals, void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(d_dperm && d_dorig && d_pvals && count >= 0 && L >= 1, PLSB_ERR_ARG,
             "plsb_perm_pvals: bad argument");
  return launch_pvals(h, d_dperm, count, L, d_dorig, d_pvals, as_stream(stream));
}

int plsb_percentile(plsb_handle_t h, const double *d_distrib, int count, int n_series, double q_lo,
                    double q_hi, double *d_lo, double *d_hi, void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(d_distrib && d_lo && d_hi, PLSB_ERR_ARG, "plsb_percentile: null argument");
  return launch_percentile(h, d_distrib, count, n_series, q_lo, q_hi, d_lo, d_hi,
                           as_stream(stream));
}

int plsb_percentile_series(plsb_handle_t h, const double *d_series, int n_series, int count,
                           int64_t ld, double q_lo, double q_hi, double *d_lo, double *d_hi,
                           void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(d_series && d_lo && d_hi, PLSB_ERR_ARG, "plsb_percentile_series: null argument");
  return launch_percentile_series(h, d_series, ld, count, n_series, q_lo, q_hi, d_lo, d_hi,
                                  as_stream(stream));
}

int plsb_transpose(plsb_handle_t h, const double *d_in, int rows, int cols, double *d_out,
                   void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(d_in && d_out && rows >= 0 && cols >= 0, PLSB_ERR_ARG, "plsb_transpose: bad argument");
  if (rows == 0 || cols == 0) return PLSB_OK;
  return launch_transpose(h, d_in, rows, cols, cols, d_out, as_stream(stream));
}

int plsb_boot_ratio(plsb_handle_t h, const double *d_bs, const double *d_usum,
                    const double *d_usquare, int64_t n_elem, int n_boot, int add_orig,
                    double *d_bsr, double *d_se, void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(d_bs && d_usum && d_usquare && d_bsr && d_se, PLSB_ERR_ARG,
             "plsb_boot_ratio: null argument");
  return launch_boot_ratio(h, d_bs, d_usum, d_usquare, n_elem, n_boot, add_orig, d_bsr, d_se,
                           as_stream(stream));
}

int plsb_dgemm(plsb_handle_t h, const double *d_A, const double *d_X, int M, int N, int Kd,
               double *d_C, void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(d_A && d_X && d_C && M >= 1 && N >= 1 && Kd >= 1, PLSB_ERR_ARG,
             "plsb_dgemm: bad argument");
  return dense_gemm(h, d_A, d_X, M, N, Kd, d_C, as_stream(stream));
}

// Times one variant of the cross-covariance GEMM on synthetic operands (profiling aid):
// variant 0 = STORE, 1 = STORE with column scales (scale_div rows share a scale row),
// 2 = ROWSUMSQ.  Returns the mean milliseconds per launch over `iters` launches.
int plsb_gemm_probe(plsb_handle_t h, int variant, int M, int N, int Kd, int k_valid, int scale_div,
                    int iters, double *ms_out, void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(ms_out && M >= 1 && N >= 1 && Kd >= 1 && iters >= 1 && variant >= 0 && variant <= 2,
             PLSB_ERR_ARG, "plsb_gemm_probe: bad argument");
  cudaStream_t st = as_stream(stream);
  const int M_pad = round_up(M, GEMM_BM), N_pad = round_up(N, GEMM_BN), K_pad = round_up(Kd, GEMM_BK);
  scale_div = std::max(scale_div, 1);
  PLSB_TRY(h->A.ensure(sizeof(double) * (size_t)M_pad * K_pad));
  PLSB_TRY(h->misc.ensure(sizeof(double) * (size_t)K_pad * N_pad));
  PLSB_CUDA(cudaMemsetAsync(h->A.p, 0x3F, sizeof(double) * (size_t)M_pad * K_pad, st));
  PLSB_CUDA(cudaMemsetAsync(h->misc.p, 0x3F, sizeof(double) * (size_t)K_pad * N_pad, st));
  GemmArgs g;
  g.A = h->A.as<double>();
  g.lda = K_pad;
  g.X = h->misc.as<double>();
  g.ldx = N_pad;
  g.M_pad = M_pad;
  g.N_pad = N_pad;
  g.Kd = K_pad;
  g.k_valid = k_valid > 0 ? k_valid : Kd;
  g.x_persistent = true;   // synthetic right operand, constant while probing
  if (variant == 2) {
    g.n_splits = gemm_rowsq_splits(h, g);
    PLSB_TRY(h->rowsq.ensure(sizeof(double) * (size_t)g.n_splits * M_pad));
    g.rowsq = h->rowsq.as<double>();
  } else {
    g.ldc = tune_int("PLSB_PROBE_LDC", N_pad);   // small pitch: same stores into few pages
    PLSB_TRY(h->R.ensure(sizeof(double) * ((size_t)M_pad * g.ldc + N_pad)));
    g.C = h->R.as<double>();
    if (variant == 1) {
      const size_t rows = (size_t)cdiv(M_pad, scale_div);
      PLSB_TRY(h->S1.ensure(sizeof(double) * rows * N_pad));
      PLSB_CUDA(cudaMemsetAsync(h->S1.p, 0x3F, sizeof(double) * rows * N_pad, st));
      g.scale = h->S1.as<double>();
      g.scale_div = scale_div;
      g.lds = tune_int("PLSB_PROBE_LDS", N_pad);   // 0: every scale row aliases the first (L2 resident)
    }
  }
  PLSB_TRY(launch_gemm(h, g, st));   // warm-up
  cudaEvent_t e0, e1;
  PLSB_CUDA(cudaEventCreate(&e0));
  PLSB_CUDA(cudaEventCreate(&e1));
  PLSB_CUDA(cudaEventRecord(e0, st));
  for (int i = 0; i < iters; ++i) PLSB_TRY(launch_gemm(h, g, st));
  PLSB_CUDA(cudaEventRecord(e1, st));
  PLSB_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  PLSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_out = (double)ms / iters;
  return PLSB_OK;
}

// ---- SIMPLS (pls_regression) ---------------------------------------------------

// x_weights operand rows D (n*L, S_pad) in h->A  ->  R (n*L, ldx) = D @ X
static int simpls_weights_gemm(plsb_ctx *h, int n, cudaStream_t st) {
  const Layout &l = h->lay;
  const long long M = (long long)n * l.L, M_pad = round_up_ll(M, GEMM_BM);
  PLSB_TRY(h->R.ensure(sizeof(double) * (size_t)M_pad * l.ldx));
  PLSB_TRY(zero_tail(h->A.as<double>(), M, M_pad, l.S_pad, st));
  GemmArgs g;
  g.A = h->A.as<double>();
  g.lda = l.S_pad;
  g.X = h->Xraw.as<double>();
  g.x_persistent = true;
  g.ldx = l.ldx;
  g.M_pad = (int)M_pad;
  g.N_pad = l.ldx;
  g.Kd = l.S_pad;
  g.k_valid = l.S;
  g.C = h->R.as<double>();
  g.ldc = l.ldx;
  return launch_gemm(h, g, st);
}

static int simpls_chunk(const plsb_ctx *h, int count, bool boot) {
  const Layout &l = h->lay;
  size_t per = sizeof(double) * (4 * (size_t)l.S * l.L + 2 * (size_t)l.S * l.T);
  if (boot) per += sizeof(double) * ((size_t)l.L * l.S_pad + (size_t)l.L * l.ldx + (size_t)l.L * l.L);
  long long n = std::max<long long>(1, (long long)(h->ws_limit / per));
  n = std::min<long long>(n, ((1ll << 31) - 1024) / std::max(l.L, 1));
  return (int)std::min<long long>(n, count);
}

int plsb_simpls_set_original(plsb_handle_t h, const double *d_xweights, void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(h->has_data && h->lay.simpls(), PLSB_ERR_STATE,
             "plsb_simpls_set_original needs a SIMPLS handle with data");
  PLSB_CHECK(d_xweights != nullptr, PLSB_ERR_ARG, "plsb_simpls_set_original: null argument");
  const Layout &l = h->lay;
  cudaStream_t st = as_stream(stream);
  PLSB_TRY(h->Uo.ensure(sizeof(double) * (size_t)l.B * l.L));
  PLSB_TRY(h->Sx.ensure(sizeof(double) * (size_t)l.S * l.L));
  // So = X (orig - colmean(orig)): sign(efficient_corr(x_weights_boot, orig)) in sample space
  PLSB_TRY(launch_colcenter(h, d_xweights, l.B, l.L, h->Uo.as<double>(), st));
  PLSB_TRY(launch_xproj(h, h->Xraw.as<double>(), l.ldx, l.S, l.B, h->Uo.as<double>(), l.L, nullptr,
                        h->Sx.as<double>(), st));
  h->has_original = true;
  return PLSB_OK;
}

int plsb_simpls_set_row_mask(plsb_handle_t h, const int32_t *d_valid_x, const int32_t *d_valid_y,
                             void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(h->has_data && h->lay.simpls(), PLSB_ERR_STATE,
             "plsb_simpls_set_row_mask needs a SIMPLS handle with data");
  h->has_rowmask = false;
  h->has_original = false;
  if (!d_valid_x && !d_valid_y) return PLSB_OK;
  PLSB_CHECK(d_valid_x && d_valid_y, PLSB_ERR_ARG,
             "plsb_simpls_set_row_mask: give both masks or none");
  const size_t bytes = sizeof(int32_t) * (size_t)h->lay.S;
  PLSB_TRY(h->rowmask.ensure(2 * bytes));
  PLSB_CUDA(cudaMemcpyAsync(h->rowmask.p, d_valid_x, bytes, cudaMemcpyDeviceToDevice,
                            as_stream(stream)));
  PLSB_CUDA(cudaMemcpyAsync(h->rowmask.as<int32_t>() + h->lay.S, d_valid_y, bytes,
                            cudaMemcpyDeviceToDevice, as_stream(stream)));
  h->has_rowmask = true;
  return PLSB_OK;
}

int plsb_simpls_decompose(plsb_handle_t h, const double *d_omega, double *d_xweights,
                          double *d_pctvar, void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(h->has_data && h->lay.simpls(), PLSB_ERR_STATE,
             "plsb_simpls_decompose needs a SIMPLS handle with data");
  PLSB_CHECK(d_xweights && d_pctvar, PLSB_ERR_ARG, "plsb_simpls_decompose: null output");
  const Layout &l = h->lay;
  cudaStream_t st = as_stream(stream);
  h->has_original = false;
  PLSB_TRY(h->A.ensure(sizeof(double) * (size_t)GEMM_BM * l.S_pad));
  PLSB_TRY(launch_simpls(h, nullptr, 1, 0, 1, d_omega, 0, (long long)l.T * 11, d_pctvar, nullptr,
                         st));
  PLSB_TRY(simpls_weights_gemm(h, 1, st));
  PLSB_TRY(launch_xweights_flip(h, h->R.as<double>(), l.ldx, l.B, l.L, d_xweights, st));
  return plsb_simpls_set_original(h, d_xweights, stream);
}

int plsb_simpls_run_perms(plsb_handle_t h, const int32_t *d_idx, int count, const double *d_omega,
                          double *d_pctvar, void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(h->has_data && h->lay.simpls(), PLSB_ERR_STATE,
             "plsb_simpls_run_perms needs a SIMPLS handle with data");
  PLSB_CHECK(d_idx && d_pctvar && count >= 0, PLSB_ERR_ARG, "plsb_simpls_run_perms: bad argument");
  const Layout &l = h->lay;
  cudaStream_t st = as_stream(stream);
  const int chunk = simpls_chunk(h, count, false);
  const long long om = (long long)l.T * 11;
  for (int off = 0; off < count; off += chunk) {
    const int n = std::min(chunk, count - off);
    PLSB_TRY(launch_simpls(h, d_idx + (size_t)off * l.S, n, 0, 0,
                           d_omega ? d_omega + (size_t)off * om : nullptr, om, 0,
                           d_pctvar + (size_t)off * l.L, nullptr, st));
  }
  return PLSB_OK;
}

static int simpls_boots_impl(plsb_ctx *h, const int32_t *d_idx, int count, const double *d_omega,
                             const double *d_yres, double *d_pctvar, double *d_distrib,
                             double *d_usum, double *d_usquare, cudaStream_t st) {
  const Layout &l = h->lay;
  const int chunk = simpls_chunk(h, count, true);
  const long long om = (long long)l.T * 11;
  const size_t ystride = (size_t)l.S * l.T;
  for (int off = 0; off < count; off += chunk) {
    const int n = std::min(chunk, count - off);
    const long long M_pad = round_up_ll((long long)n * l.L, GEMM_BM);
    PLSB_TRY(h->A.ensure(sizeof(double) * (size_t)M_pad * l.S_pad));
    PLSB_TRY(launch_simpls(h, d_idx + (size_t)off * l.S, n, 1, 1,
                           d_omega ? d_omega + (size_t)off * om : nullptr, om, 0,
                           d_pctvar + (size_t)off * l.L, d_distrib + (size_t)off * l.T * l.L, st,
                           d_yres ? d_yres + (size_t)off * ystride : nullptr));
    PLSB_TRY(simpls_weights_gemm(h, n, st));
    // u_sum += x_weights_r, u_square += x_weights_r^2 (pyls/base.py:510-511): the shared
    // accumulation kernel with identity rotations
    PLSB_TRY(h->misc.ensure(sizeof(double) * (size_t)n * l.L * accum_ldm(l.L)));
    PLSB_TRY(launch_identity_blocks(h, h->misc.as<double>(), n, l.L, accum_ldm(l.L), st));
    PLSB_TRY(launch_accum_u(h, h->R.as<double>(), l.ldx, n, l.L, l.B, h->misc.as<double>(), l.L,
                            d_usum, d_usquare, st));
  }
  return PLSB_OK;
}

int plsb_simpls_run_boots(plsb_handle_t h, const int32_t *d_idx, int count, const double *d_omega,
                          double *d_pctvar, double *d_distrib, double *d_usum, double *d_usquare,
                          void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(h->has_data && h->has_original && h->lay.simpls(), PLSB_ERR_STATE,
             "plsb_simpls_run_boots needs a SIMPLS handle with data and the original weights");
  PLSB_CHECK(d_idx && d_pctvar && d_distrib && d_usum && d_usquare && count >= 0, PLSB_ERR_ARG,
             "plsb_simpls_run_boots: bad argument");
  return simpls_boots_impl(h, d_idx, count, d_omega, nullptr, d_pctvar, d_distrib, d_usum,
                           d_usquare, as_stream(stream));
}

int plsb_simpls_run_boots_yres(plsb_handle_t h, const int32_t *d_idx, int count,
                               const double *d_omega, const double *d_yres, double *d_pctvar,
                               double *d_distrib, double *d_usum, double *d_usquare,
                               void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(h->has_data && h->has_original && h->lay.simpls(), PLSB_ERR_STATE,
             "plsb_simpls_run_boots_yres needs a SIMPLS handle with data and the original "
             "weights");
  PLSB_CHECK(d_idx && d_yres && d_pctvar && d_distrib && d_usum && d_usquare && count >= 0,
             PLSB_ERR_ARG, "plsb_simpls_run_boots_yres: bad argument");
  PLSB_CHECK(!h->has_rowmask, PLSB_ERR_ARG,
             "plsb_simpls_run_boots_yres: missing rows are not supported with per-resample Y");
  return simpls_boots_impl(h, d_idx, count, d_omega, d_yres, d_pctvar, d_distrib, d_usum,
                           d_usquare, as_stream(stream));
}

int plsb_small_decomp(plsb_handle_t h, const double *d_G, const double *d_H, int count, int K,
                      int L, const double *d_dorig, double *d_M, double *d_lam, void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(d_G && d_H && d_M && count >= 0, PLSB_ERR_ARG, "plsb_small_decomp: bad argument");
  return launch_small_decomp(h, d_G, d_H, count, K, L, d_dorig, d_M, L, d_lam, as_stream(stream));
}

// one chunk of train / test splits: see crossval.cu
static int crossval_chunk(plsb_ctx *h, const int32_t *mask, int n, int max_test, double *d_r,
                          double *d_r2, cudaStream_t st) {
  const Layout &l = h->lay;
  const int stride = l.K + round_up(max_test, l.T);   // rows of a split's stacked matrix
  const int rpr = stride / l.T;                        // rows of a split in S1 / S2
  const long long cellpad_w = round_up_ll((long long)n * l.T, GEMM_BM);
  const long long cellpad_c = round_up_ll(n, GEMM_BM);
  const long long Mw_op = cellpad_w * l.J, Mc_op = cellpad_c * l.J;
  const long long Z_rows = round_up_ll((long long)n * stride, GEMM_BM);
  const long long S_rows = round_up_ll((long long)n * rpr, GEMM_BM);
  PLSB_TRY(h->A.ensure(sizeof(double) * (size_t)Mw_op * l.S_pad));
  PLSB_TRY(h->Ac.ensure(sizeof(double) * (size_t)Mc_op * l.S_pad));
  PLSB_TRY(h->R.ensure(sizeof(double) * (size_t)Z_rows * l.ldx));
  PLSB_TRY(h->S1.ensure(sizeof(double) * (size_t)S_rows * l.ldx));
  PLSB_TRY(h->S2.ensure(sizeof(double) * (size_t)S_rows * l.ldx));
  const size_t n_kr = (size_t)(Mw_op + Mc_op) / GEMM_BM;
  PLSB_TRY(h->maps.ensure(sizeof(int4) * n_kr + sizeof(int) * (size_t)(Mw_op + Mc_op)));
  int4 *kr_w = h->maps.as<int4>(), *kr_c = kr_w + Mw_op / GEMM_BM;
  int *map_w = reinterpret_cast<int *>(kr_c + Mc_op / GEMM_BM), *map_c = map_w + Mw_op;
  PLSB_TRY(h->part.ensure(sizeof(double) * (size_t)n * l.J * l.T + sizeof(int) * (size_t)n * l.J));
  double *ytrain = h->part.as<double>();
  int *ntrain = reinterpret_cast<int *>(ytrain + (size_t)n * l.J * l.T);
  double *Z = h->R.as<double>();

  PLSB_CUDA(cudaMemsetAsync(h->A.p, 0, sizeof(double) * (size_t)Mw_op * l.S_pad, st));
  PLSB_CUDA(cudaMemsetAsync(h->Ac.p, 0, sizeof(double) * (size_t)Mc_op * l.S_pad, st));
  PLSB_CUDA(cudaMemsetAsync(Z, 0, sizeof(double) * (size_t)Z_rows * l.ldx, st));
  PLSB_TRY(launch_build_maps(h, n, l.T, stride, cellpad_w, map_w, kr_w, st));
  PLSB_TRY(launch_build_maps(h, n, 1, rpr, cellpad_c, map_c, kr_c, st));
  PLSB_TRY(launch_build(h, BUILD_TRAIN, mask, nullptr, n, h->A.as<double>(), h->Ac.as<double>(),
                        nullptr, cellpad_w, cellpad_c, st, ntrain, ytrain));
  // sums and sums of squares of the (globally standardised) training rows of every cell
  GemmArgs g;
  g.lda = l.S_pad;
  g.ldx = l.ldx;
  g.N_pad = l.ldx;
  g.Kd = l.S_pad;
  g.k_valid = l.S;
  g.ldc = l.ldx;
  g.k_len = l.kr_max;
  g.X = h->Xglob.as<double>();
  g.A = h->Ac.as<double>();
  g.M_pad = (int)Mc_op;
  g.row_map = map_c;
  g.kranges = kr_c;
  g.C = h->S1.as<double>();
  PLSB_TRY(launch_gemm(h, g, st));
  g.C = h->S2.as<double>();
  g.square_b = true;
  PLSB_TRY(launch_gemm(h, g, st));
  g.square_b = false;
  PLSB_TRY(launch_rescale_test(h, mask, n, max_test, h->S1.as<double>(), h->S2.as<double>(), rpr,
                               ntrain, Z, stride, st));
  if (l.corr()) {
    PLSB_TRY(launch_colscale(h, h->S1.as<double>(), h->S2.as<double>(), n * rpr, l.ldx, st, rpr,
                             ntrain));
    g.scale = h->S1.as<double>();
    g.scale_div = l.T;
    g.scale_rows = (int)S_rows;
    g.lds = l.ldx;
  }
  g.A = h->A.as<double>();
  g.M_pad = (int)Mw_op;
  g.row_map = map_w;
  g.kranges = kr_w;
  g.C = Z;
  PLSB_TRY(launch_gemm(h, g, st));
  // one Gram pass over [R; X_resc]: G = R R^T and P = X_resc R^T
  const size_t gz = (size_t)stride * stride, kk = (size_t)l.K * l.K;
  PLSB_TRY(h->G.ensure(sizeof(double) * (size_t)n * gz));
  PLSB_TRY(h->misc.ensure(sizeof(double) * (size_t)n * (kk + l.K)));
  double *V = h->misc.as<double>(), *lam = V + (size_t)n * kk;
  PLSB_TRY(launch_gram_proj(h, Z, l.ldx, n, stride, nullptr, 0, h->G.as<double>(), nullptr, st));
  PLSB_TRY(launch_sym_eig(h, h->G.as<double>(), n, l.K, V, lam, 0, st, stride, (long long)gz));
  return launch_cv_score(h, mask, n, max_test, h->G.as<double>(), stride, V, lam, ytrain, d_r,
                         d_r2, st);
}

int plsb_crossval(plsb_handle_t h, const int32_t *d_train, int count, int max_test, double *d_r,
                  double *d_r2, void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(h->has_data && h->lay.behavioral(), PLSB_ERR_STATE,
             "plsb_crossval needs a behavioural handle with data");
  PLSB_CHECK(d_train && d_r && d_r2 && count >= 0 && max_test >= 1, PLSB_ERR_ARG,
             "plsb_crossval: bad argument");
  PLSB_CHECK(!h->lay.tall(), PLSB_ERR_ARG,
             "plsb_crossval: not available when the latent rows (K=%d) exceed the features (B=%d)",
             h->lay.K, h->lay.B);
  const Layout &l = h->lay;
  const int stride = l.K + round_up(max_test, l.T);
  cudaStream_t st = as_stream(stream);
  const size_t per = sizeof(double) * ((size_t)stride * l.ldx + (size_t)(l.K + l.J) * l.S_pad +
                                       2 * (size_t)(stride / l.T) * l.ldx);
  const int chunk = (int)std::max<long long>(1, std::min<long long>(count, h->ws_limit / per));
  for (int off = 0; off < count; off += chunk) {
    const int n = std::min(chunk, count - off);
    PLSB_TRY(crossval_chunk(h, d_train + (size_t)off * l.S, n, max_test,
                            d_r + (size_t)off * l.T, d_r2 + (size_t)off * l.T, st));
  }
  return PLSB_OK;
}

// ---- split-half resampling (splithalf.cu) ----------------------------------------

// Z (nh*K rows, ldx) in h->R: cross-covariances of nh halves.  Half r is side (r & 1) of
// mask (r >> 1), see HalfSpec; the column statistics of a behavioural correlation run
// over the rows of the half only (like a training split of the cross-validation).
static int halves_chunk(plsb_ctx *h, const int32_t *masks, const double *yperm,
                        const HalfSpec &hs, int nh, cudaStream_t st) {
  const Layout &l = h->lay;
  const bool grouped = l.behavioral() && l.J > 1;
  const bool scaled = l.corr();
  const long long Mw = (long long)nh * l.K, Mc = (long long)nh * l.J;
  const long long cellpad_w = grouped ? round_up_ll((long long)nh * l.T, GEMM_BM) : 0;
  const long long cellpad_c = grouped ? round_up_ll(nh, GEMM_BM) : 0;
  const long long Mw_op = grouped ? cellpad_w * l.J : round_up_ll(Mw, GEMM_BM);
  const long long Mc_op = grouped ? cellpad_c * l.J : round_up_ll(Mc, GEMM_BM);
  const long long Mw_pad = round_up_ll(Mw, GEMM_BM), Mc_pad = round_up_ll(Mc, GEMM_BM);
  PLSB_CHECK(Mw_op < (1ll << 31) - 1, PLSB_ERR_ARG, "chunk of %d halves is too large", nh);
  PLSB_TRY(h->A.ensure(sizeof(double) * (size_t)Mw_op * l.S_pad));
  PLSB_TRY(h->R.ensure(sizeof(double) * (size_t)Mw_pad * l.ldx));
  int *map_w = nullptr, *map_c = nullptr;
  int4 *kr_w = nullptr, *kr_c = nullptr;
  if (grouped) {
    const size_t n_kr = (size_t)(Mw_op + Mc_op) / GEMM_BM;
    PLSB_TRY(h->maps.ensure(sizeof(int4) * n_kr + sizeof(int) * (size_t)(Mw_op + Mc_op)));
    kr_w = h->maps.as<int4>();
    kr_c = kr_w + Mw_op / GEMM_BM;
    map_w = reinterpret_cast<int *>(kr_c + Mc_op / GEMM_BM);
    map_c = map_w + Mw_op;
    PLSB_CUDA(cudaMemsetAsync(h->A.p, 0, sizeof(double) * (size_t)Mw_op * l.S_pad, st));
    PLSB_TRY(launch_build_maps(h, nh, l.T, l.K, cellpad_w, map_w, kr_w, st));
  } else {
    PLSB_TRY(zero_tail(h->A.as<double>(), Mw, Mw_op, l.S_pad, st));
  }
  int *ncell = nullptr;
  if (l.behavioral()) {
    PLSB_TRY(h->part.ensure(sizeof(int) * (size_t)nh * l.J));
    ncell = h->part.as<int>();
  }
  if (scaled) {
    PLSB_TRY(h->Ac.ensure(sizeof(double) * (size_t)Mc_op * l.S_pad));
    PLSB_TRY(h->S1.ensure(sizeof(double) * (size_t)Mc_pad * l.ldx));
    PLSB_TRY(h->S2.ensure(sizeof(double) * (size_t)Mc_pad * l.ldx));
    if (grouped) {
      PLSB_CUDA(cudaMemsetAsync(h->Ac.p, 0, sizeof(double) * (size_t)Mc_op * l.S_pad, st));
      PLSB_TRY(launch_build_maps(h, nh, 1, l.J, cellpad_c, map_c, kr_c, st));
    } else {
      PLSB_TRY(zero_tail(h->Ac.as<double>(), Mc, Mc_op, l.S_pad, st));
    }
  }
  PLSB_TRY(launch_build(h, BUILD_HALF, masks, yperm, nh, h->A.as<double>(),
                        scaled ? h->Ac.as<double>() : nullptr, nullptr, cellpad_w, cellpad_c, st,
                        ncell, nullptr, &hs));
  GemmArgs g;
  g.lda = l.S_pad;
  g.ldx = l.ldx;
  g.N_pad = l.ldx;
  g.Kd = l.S_pad;
  g.k_valid = l.S;
  g.ldc = l.ldx;
  g.x_persistent = true;
  if (grouped) g.k_len = l.kr_max;
  if (scaled) {
    g.A = h->Ac.as<double>();
    g.X = h->Xglob.as<double>();
    g.M_pad = (int)Mc_op;
    g.row_map = map_c;
    g.kranges = kr_c;
    g.C = h->S1.as<double>();
    g.a_counts = true;           // multiplicity rows
    PLSB_TRY(launch_gemm(h, g, st));
    g.C = h->S2.as<double>();
    g.square_b = true;
    PLSB_TRY(launch_gemm(h, g, st));
    g.square_b = false;
    g.a_counts = false;
    PLSB_TRY(launch_colscale(h, h->S1.as<double>(), h->S2.as<double>(), (int)Mc, l.ldx, st, l.J,
                             ncell));
    g.scale = h->S1.as<double>();
    g.scale_div = l.T;
    g.scale_rows = (int)Mc_pad;
    g.lds = l.ldx;
  }
  g.A = h->A.as<double>();
  g.X = h->Xglob.as<double>();
  g.M_pad = (int)Mw_op;
  g.row_map = map_w;
  g.kranges = kr_w;
  g.C = h->R.as<double>();
  PLSB_TRY(launch_gemm(h, g, st));
  return PLSB_OK;
}

int plsb_split_half(plsb_handle_t h, const int32_t *d_idx, const double *d_yperm, int count,
                    const int32_t *d_masks, int n_split, int use_original, double *d_ucorr,
                    double *d_vcorr, void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(h->has_data && !h->lay.simpls(), PLSB_ERR_STATE,
             "plsb_split_half needs a behavioural or mean-centred handle with data");
  PLSB_CHECK(!use_original || h->has_original, PLSB_ERR_STATE,
             "plsb_split_half(use_original) before the original decomposition is set");
  PLSB_CHECK(d_masks && d_ucorr && d_vcorr && count >= 0 && n_split >= 1, PLSB_ERR_ARG,
             "plsb_split_half: bad argument");
  PLSB_CHECK(!(d_idx && d_yperm), PLSB_ERR_ARG,
             "plsb_split_half: give a permutation table or pre-permuted Y matrices, not both");
  PLSB_CHECK(!h->lay.tall(), PLSB_ERR_ARG,
             "plsb_split_half: not available when the latent rows (K=%d) exceed the features "
             "(B=%d)", h->lay.K, h->lay.B);
  PLSB_CHECK(!d_yperm || h->lay.behavioral(), PLSB_ERR_ARG,
             "plsb_split_half: only behavioural analyses have a Y matrix to permute");
  const Layout &l = h->lay;
  const int K = l.K, K2 = 2 * K;
  cudaStream_t st = as_stream(stream);
  PLSB_CUDA(cudaMemsetAsync(d_ucorr, 0, sizeof(double) * (size_t)count * K, st));
  PLSB_CUDA(cudaMemsetAsync(d_vcorr, 0, sizeof(double) * (size_t)count * K, st));
  // halves per pass from the workspace limit; whole permutations when they fit
  const int halves_max = std::max(2, chunk_size(h, true, INT32_MAX / 2));
  int np = 1, ns = n_split;
  if (halves_max >= 2 * n_split)
    np = (int)std::min<long long>(count, halves_max / (2 * n_split));
  else
    ns = std::max(1, halves_max / 2);
  const size_t ystride = (size_t)l.S * l.T, kk = (size_t)K * K;
  for (int off = 0; off < count; off += np) {
    const int n = std::min(np, count - off);
    const int32_t *idx = d_idx ? d_idx + (size_t)off * l.S : nullptr;
    const double *yp = d_yperm ? d_yperm + (size_t)off * ystride : nullptr;
    // the data sets being split: R_p, their decomposition and the projection blocks
    PLSB_CHECK(idx || yp || n == 1, PLSB_ERR_ARG,
               "plsb_split_half: several data sets need a permutation table");
    PLSB_TRY(crosscov_chunk(h, idx, yp, n, false, nullptr, st));
    PLSB_TRY(h->misc.ensure(sizeof(double) * ((size_t)n * K2 * l.ldx + (size_t)n * (kk + K))));
    double *PB = h->misc.as<double>();
    double *V = PB + (size_t)n * K2 * l.ldx, *dv = V + (size_t)n * kk;
    PLSB_TRY(launch_projblock(h, h->R.as<double>(), n, PB, st));
    long long v_stride = (long long)kk, d_stride = K;
    if (use_original) {
      V = h->Vo.as<double>();
      dv = h->dorig.as<double>();
      v_stride = d_stride = 0;
    } else {
      PLSB_TRY(h->G.ensure(sizeof(double) * (size_t)n * kk));
      PLSB_TRY(launch_gram_proj(h, h->R.as<double>(), l.ldx, n, K, nullptr, 0, h->G.as<double>(),
                                nullptr, st));
      PLSB_TRY(launch_sym_eig(h, h->G.as<double>(), n, K, V, dv, 1, st));
    }
    for (int s0 = 0; s0 < n_split; s0 += ns) {
      const int nsc = std::min(ns, n_split - s0);
      const int pairs = n * nsc;
      HalfSpec hs;
      hs.yidx = idx;
      hs.ns = nsc;
      hs.s0 = s0;
      hs.n_split = n_split;
      PLSB_TRY(halves_chunk(h, d_masks + (size_t)off * n_split * l.S, yp, hs, 2 * pairs, st));
      PLSB_TRY(h->G.ensure(sizeof(double) * (size_t)pairs * K2 * K2));
      PLSB_TRY(h->H.ensure(sizeof(double) * (size_t)pairs * K2 * K2));
      PLSB_TRY(launch_gram_proj(h, h->R.as<double>(), l.ldx, pairs, K2, PB, K2, h->G.as<double>(),
                                h->H.as<double>(), st, (long long)K2 * l.ldx, nsc));
      PLSB_TRY(launch_splithalf_score(h, h->G.as<double>(), h->H.as<double>(), n, nsc, V, v_stride,
                                      dv, d_stride, n_split, d_ucorr + (size_t)off * K,
                                      d_vcorr + (size_t)off * K, st));
    }
  }
  return PLSB_OK;
}

// pads count matrices (K,B) into the R workspace with a row pitch that is a multiple of 128
static int pad_R(plsb_ctx *h, const double *d_R, int count, int K, int B, long long *ldr,
                 cudaStream_t st) {
  *ldr = round_up(B, GEMM_BN);
  PLSB_TRY(h->R.ensure(sizeof(double) * (size_t)count * K * (size_t)*ldr));
  return launch_pad_copy(h, d_R, count * K, B, h->R.as<double>(), count * K, (int)*ldr, st);
}

int plsb_gram_proj(plsb_handle_t h, const double *d_R, int count, int K, int B, const double *d_Uo,
                   int L, double *d_G, double *d_H, void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(d_R && d_G && count >= 0 && K >= 1 && B >= 1 && (!d_Uo || (d_H && L >= 1)),
             PLSB_ERR_ARG, "plsb_gram_proj: bad argument");
  cudaStream_t st = as_stream(stream);
  long long ldr = 0;
  PLSB_TRY(pad_R(h, d_R, count, K, B, &ldr, st));
  const double *uot = nullptr;
  if (d_Uo) {
    PLSB_TRY(h->misc.ensure(sizeof(double) * (size_t)L * ldr));
    PLSB_TRY(launch_transpose_pad(h, d_Uo, B, L, h->misc.as<double>(), ldr, st));
    uot = h->misc.as<double>();
  }
  return launch_gram_proj(h, h->R.as<double>(), ldr, count, K, uot, L, d_G, d_H, st);
}

int plsb_accum_u(plsb_handle_t h, const double *d_R, int count, int K, int B, const double *d_M,
                 int L, double *d_usum, double *d_usquare, void *stream) {
  PLSB_HANDLE(h);
  PLSB_CHECK(d_R && d_M && d_usum && d_usquare && count >= 0 && K >= 1 && B >= 1 && L >= 1,
             PLSB_ERR_ARG, "plsb_accum_u: bad argument");
  cudaStream_t st = as_stream(stream);
  long long ldr = 0;
  PLSB_TRY(pad_R(h, d_R, count, K, B, &ldr, st));
  const int ldm = accum_ldm(L);
  PLSB_TRY(h->M.ensure(sizeof(double) * (size_t)count * K * ldm));
  PLSB_TRY(launch_pad_copy(h, d_M, count * K, L, h->M.as<double>(), count * K, ldm, st));
  return launch_accum_u(h, h->R.as<double>(), ldr, count, K, B, h->M.as<double>(), L, d_usum,
                        d_usquare, st);
}

}  // extern "C"
