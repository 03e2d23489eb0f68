// SIMPLS resampling (pls_regression) in sample space, one CTA per resample.
//
// The reference (pyls/types/regression.py:56-186, 279-373) recomputes for every
// resample Cov = X0^T Y0 (B x T) and runs, per component, a randomized top-1
// SVD of Cov (sklearn randomized_svd: Omega = RandomState(seed).normal((T,11)),
// 7 or 4 power iterations), the score t = X0 r, loadings and a deflation of Cov.
// Every B-vector that algorithm touches is X0^T (S-vector), so with the S x S
// matrix Kx = X0 X0^T (for a resample: a gather of the fixed Kraw = X X^T,
// centred on both sides) the whole component loop runs on S x T, T x T and
// 11 x 11 objects:
//
//   A = Y0 (deflated in place);   KA = Kx A;   C = A^T KA            (= Cov^T Cov)
//   W = orth(C^n_iter Omega)  (T x p, p = min(T, 11); the randomized range finder)
//   Gz = W^T C W = Rc^T Rc (Cholesky);   Bm = Rc^-T W^T C;   uh = top left singular vector
//   wt = W Rc^-1 uh;   a = A wt;   t = KA wt;   x_weights column = X0^T a / |t|
//   pctvar_y = |Yp^T t|^2 / |Y0|^2;   b = MGS(t) in the Kx inner product;   A -= b (Kx b)^T A ...
//
// (the Cholesky form of the 11-dimensional step is the reference's final QR of the
// B x 11 range matrix, restated on its Gram matrix; dependent columns are dropped,
// which selects the same vector).  Only the
// final x_weights = X0^T Wcoef of a bootstrap needs a B-sized product; it goes
// through the shared DMMA GEMM as L operand rows per resample.
// The tests check this against the direct B-space restatement of the reference.
#include "common.cuh"
#include "jacobi.cuh"

namespace plsb {
namespace {

constexpr int SP_THREADS = 128;
constexpr int SP_MIN_CTAS = 4;   // resamples per SM: their single-warp sections overlap
constexpr int SP_WARPS = SP_THREADS / 32;
constexpr int SP_PROBES = 11;   // n_components (1) + n_oversamples (10) of randomized_svd

struct SimplsParams {
  int S, T, L, p, n_iter, boot, emit_ops, lda;
  const double *Kraw, *Yc, *omega, *So;
  const double *Yres;     // optional (n, S, T): resample r draws its rows from Yres[r] instead of Yc
  const int32_t *idx;
  const int32_t *valid;   // optional (2, S): row 0 = rows of X, row 1 = rows of Y; 0 = missing (all NaN)
  long long om_stride_r, om_stride_c;
  double *Wcoef, *Bs, *Gs, *Tm;   // (n, S, L) each
  double *Abuf, *KAbuf;           // (n, S, T) each: A and Kx A of every resample (L2-resident scratch)
  double *pct, *D, *distrib;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum over the CTA; every thread gets the result.  red: >= SP_WARPS doubles
__device__ double block_sum(double v, double *red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < SP_WARPS; ++w) t += red[w];
  return t;
}

// Upper Cholesky factor of the symmetric P x P matrix G (upper triangle used), in
// place, by ONE warp: G = Rc^T Rc.  A pivot at or below rel * max(diag) is dropped
// (its row of Rc is zero): a rank-deficient Gram matrix keeps the span of its
// independent columns.  The trailing update is spread over all lanes.
__device__ void warp_chol(double *G, int P, int ld, double rel) {
  const int lane = threadIdx.x & 31;
  double dmax = 0.0;
  for (int k = 0; k < P; ++k) dmax = fmax(dmax, G[k * ld + k]);
  for (int k = 0; k < P; ++k) {
    const double d = G[k * ld + k];
    const bool ok = d > rel * dmax && d > 0.0;
    const double rkk = ok ? sqrt(d) : 0.0;
    __syncwarp();
    if (lane > k && lane < P) G[k * ld + lane] = ok ? G[k * ld + lane] / rkk : 0.0;
    if (lane == k) G[k * ld + k] = rkk;
    __syncwarp();
    const int m = P - k - 1;
    for (int e = lane; e < m * m; e += 32) {
      const int i = k + 1 + e / m, j = k + 1 + e % m;
      if (j >= i) G[i * ld + j] -= G[k * ld + i] * G[k * ld + j];
    }
    __syncwarp();
  }
}

__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// KA = H (Kpi A), Kpi[i][j] = Kraw[pix[i]][pix[j]], H = column centring.  Once per
// resample (the component loop keeps KA current with the rank-one terms of the
// deflation).  DMMA m8n8k4: a warp owns 4 row fragments x all TT/8 column
// fragments; the Kraw fragments come straight from global memory (L2 resident;
// 8 rows x 4 consecutive columns = whole 32-byte sectors for permutations, nearly so
// for the sorted bootstrap tables), the A fragments from shared memory.
template <int TT>
__device__ void gram_apply(const SimplsParams &p, int S, const int *pix, const double *A,
                           double *KA, double *csum) {
  constexpr int NT = TT / 8, MT = 4;
  const int T = p.T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int n_mt = (S + 7) / 8;
  for (int m0 = warp * MT; m0 < n_mt; m0 += SP_WARPS * MT) {
    double acc[MT][NT][2];
    const double *krow[MT];
#pragma unroll
    for (int mi = 0; mi < MT; ++mi) {
      krow[mi] = p.Kraw + (size_t)pix[min((m0 + mi) * 8 + g, S - 1)] * p.S;
#pragma unroll
      for (int ni = 0; ni < NT; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    }
    int bcol[NT];   // padded columns re-read column T-1: finite, results discarded
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) bcol[ni] = min(ni * 8 + g, T - 1);
#pragma unroll 2
    for (int j0 = 0; j0 < S; j0 += 4) {
      const int j = j0 + q;
      const bool ok = j < S;
      const int jc = ok ? j : S - 1;
      const int col = pix[jc];
      double a[MT], bf[NT];
#pragma unroll
      for (int mi = 0; mi < MT; ++mi) a[mi] = ok ? __ldg(krow[mi] + col) : 0.0;
#pragma unroll
      for (int ni = 0; ni < NT; ++ni) bf[ni] = ok ? A[jc * T + bcol[ni]] : 0.0;
#pragma unroll
      for (int mi = 0; mi < MT; ++mi)
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) dmma_8x8x4(acc[mi][ni][0], acc[mi][ni][1], a[mi], bf[ni]);
    }
#pragma unroll
    for (int mi = 0; mi < MT; ++mi) {
      const int row = (m0 + mi) * 8 + g;
      if (row >= S) continue;
#pragma unroll
      for (int ni = 0; ni < NT; ++ni) {
        const int c = ni * 8 + 2 * q;
        if (c < T) KA[row * T + c] = acc[mi][ni][0];
        if (c + 1 < T) KA[row * T + c + 1] = acc[mi][ni][1];
      }
    }
  }
  __syncthreads();
  const int wl = threadIdx.x >> 5, ll = threadIdx.x & 31;
  for (int t = wl; t < T; t += SP_WARPS) {
    double v = 0.0;
    for (int i = ll; i < S; i += 32) v += KA[i * T + t];
    v = warp_sum(v);
    if (ll == 0) csum[t] = v / S;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < S * T; e += SP_THREADS) KA[e] -= csum[e % T];
  __syncthreads();
}

// out[i] = sum_j Kraw[pix[i]][pix[j]] x[j]: a warp per group of four rows, lanes along the
// rows (coalesced for permutations, nearly so for sorted bootstrap tables; sixteen
// independent loads in flight per lane hide the L2 latency); barrier
__device__ void kx_matvec(const SimplsParams &p, int S, const int *pix, const double *x,
                          double *out) {
  constexpr int RB = 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp * RB; i < S; i += SP_WARPS * RB) {
    const double *row[RB];
    double v[RB];
#pragma unroll
    for (int k = 0; k < RB; ++k) {
      row[k] = p.Kraw + (size_t)pix[min(i + k, S - 1)] * p.S;
      v[k] = 0.0;
    }
#pragma unroll 4
    for (int j = lane; j < S; j += 32) {
      const int c = pix[j];
      const double xj = x ? x[j] : 1.0;
#pragma unroll
      for (int k = 0; k < RB; ++k) v[k] += __ldg(row[k] + c) * xj;
    }
#pragma unroll
    for (int k = 0; k < RB; ++k) {
      v[k] = warp_sum(v[k]);
      if (lane == 0 && i + k < S) out[i + k] = v[k];
    }
  }
  __syncthreads();
}

template <int TT>
__global__ void __launch_bounds__(SP_THREADS, SP_MIN_CTAS) simpls_kernel(SimplsParams p) {
  extern __shared__ __align__(16) double sm[];
  const int Sf = p.S, T = p.T, L = p.L, P = p.p;   // Sf: rows of the data; S below: rows in use
  const int pe = P + (P & 1), ldz = pe | 1, halfz = pe / 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = blockIdx.x;
  // ---- shared memory carve-up ----
  // A and KA = Kx A live in global memory (streamed, L2 resident): with only the
  // small matrices in shared memory several resamples share an SM, so one's
  // single-warp sections (Cholesky, Jacobi) overlap the others' work
  double *A = p.Abuf + (size_t)blockIdx.x * Sf * T;
  double *KA = p.KAbuf + (size_t)blockIdx.x * Sf * T;
  double *tv = sm;                 // S
  double *bv = tv + Sf;            // S
  double *gv = bv + Sf;            // S
  double *C = gv + Sf;             // T*T
  double *W = C + T * T;           // T*P
  double *CW = W + T * P;          // T*P
  double *F = CW + T * P;          // P*T
  double *Bm = F + P * T;          // P*T
  double *Gz = Bm + P * T;         // pe*ldz
  double *Ez = Gz + pe * ldz;      // pe*ldz
  double *E2 = Ez + pe * ldz;      // pe*ldz
  double *Z = E2 + pe * ldz;       // max(L*T, 2L): deflation coefficients, later signs
  double *lamz = Z + max(L * T, 2 * L);   // pe
  double *inv = lamz + pe;         // pe
  double *yv = inv + pe;           // pe
  double *wt = yv + pe;            // T
  double *qv = wt + T;             // T
  double *ysum = qv + T;           // T
  double *csum = ysum + T;         // T
  double *red = csum + T;          // SP_WARPS + 2
  double *cg = red + SP_WARPS + 2; // L: G_prev^T b
  double *part = cg + L;           // 8 * SP_THREADS: partial sums of the deflation sweep
  JacobiScratch sc;
  sc.cst = part + 8 * SP_THREADS;                               // 3*halfz
  sc.pq = reinterpret_cast<int *>(sc.cst + 3 * halfz);          // 2*halfz
  int *pix = sc.pq + 2 * halfz;                                 // S
  int *piy = pix + Sf;                                          // S
  sc.blk = reinterpret_cast<short2 *>(piy + Sf);                // halfz*(halfz+1)/2
  __shared__ int s_rows;

  // behaviour matrix this resample draws its rows from (3-D Y: aggregated per bootstrap)
  const double *Yr = p.Yres ? p.Yres + (size_t)r * Sf * T : p.Yc;
  double *Wc = p.Wcoef + (size_t)r * Sf * L;
  double *Bs = p.Bs + (size_t)r * Sf * L;
  double *Gs = p.Gs + (size_t)r * Sf * L;
  double *Tm = p.Tm + (size_t)r * Sf * L;

  for (int a = tid; a < halfz; a += SP_THREADS) {
    int bi = a * halfz - a * (a - 1) / 2;
    for (int b = a; b < halfz; ++b) sc.blk[bi++] = make_short2((short)a, (short)b);
  }
  // source rows of X (pix) and Y (piy) of the rows in use.  The reference masks the rows
  // of the RESAMPLED matrices that are missing (get_mask, pyls/types/regression.py:48-53,
  // 271-272): a row is in use when both of its sources are valid; the lists keep the order
  if (warp == 0) {
    int count = 0;
    for (int base = 0; base < Sf; base += 32) {
      const int i = base + lane;
      int xs = 0, ys = 0, ok = 0;
      if (i < Sf) {
        const int src = p.idx ? p.idx[(size_t)r * Sf + i] : i;
        ys = src;
        xs = p.boot ? src : i;
        ok = !p.valid || (p.valid[xs] != 0 && p.valid[Sf + ys] != 0);
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const int pos = count + __popc(m & ((1u << lane) - 1u));
        pix[pos] = xs;
        piy[pos] = ys;
      }
      count += __popc(m);
    }
    if (lane == 0) s_rows = count;
  }
  __syncthreads();
  const int S = s_rows;
  // A = Y0 = Yp - column means;  ysum = column sums of Yp
  for (int e = tid; e < S * T; e += SP_THREADS) {
    const int i = e / T, t = e - i * T;
    A[e] = Yr[(size_t)piy[i] * T + t];
  }
  __syncthreads();
  for (int t = warp; t < T; t += SP_WARPS) {
    double v = 0.0;
    for (int i = lane; i < S; i += 32) v += A[i * T + t];
    v = warp_sum(v);
    if (lane == 0) ysum[t] = v;
  }
  __syncthreads();
  double ssy = 0.0;
  for (int e = tid; e < S * T; e += SP_THREADS) {
    const double v = A[e] - ysum[e % T] / S;
    A[e] = v;
    ssy += v * v;
  }
  ssy = block_sum(ssy, red);

  gram_apply<TT>(p, S, pix, A, KA, csum);
  // C = A^T KA (upper triangle, mirrored); the deflation keeps it current
  for (int e = tid; e < T * T; e += SP_THREADS) {
    const int t1 = e / T, t2 = e - t1 * T;
    if (t2 < t1) continue;
    double v = 0.0;
    for (int s = 0; s < S; ++s) v += A[s * T + t1] * KA[s * T + t2];
    C[t1 * T + t2] = v;
    C[t2 * T + t1] = v;
  }
  __syncthreads();
  for (int comp = 0; comp < L; ++comp) {
    // ---- randomized range finder: W = orth(C^n_iter Omega) ----
    if (T <= SP_PROBES) {
      for (int e = tid; e < T * P; e += SP_THREADS) W[e] = (e / P == e % P) ? 1.0 : 0.0;
      __syncthreads();
    } else {
      const double *om = p.omega + (size_t)r * p.om_stride_r + (size_t)comp * p.om_stride_c;
      for (int e = tid; e < T * P; e += SP_THREADS) W[e] = om[e];
      __syncthreads();
      for (int it = 0; it < p.n_iter; ++it) {
        for (int e = tid; e < T * P; e += SP_THREADS) {
          const int t = e / P, a = e - t * P;
          double v = 0.0;
          for (int u = 0; u < T; ++u) v += C[t * T + u] * W[u * P + a];
          CW[e] = v;
        }
        __syncthreads();
        // normalise the columns of CW (same span): Cholesky-QR -- Gram matrix ->
        // upper Cholesky factor Rc (warp 0) -> CW <- CW Rc^-1 (thread = row).
        // One pass keeps the basis well conditioned between power iterations
        // (orthogonal to eps * cond^2, the span to eps * cond); the last iteration
        // runs it twice so that W is orthonormal to working precision.  A
        // vanishing pivot (rank-deficient CW, e.g. after many deflations) drops
        // that column.
        const int n_pass = (it + 1 == p.n_iter) ? 2 : 1;
        for (int pass = 0; pass < n_pass; ++pass) {
          for (int e = tid; e < P * P; e += SP_THREADS) {
            const int a = e / P, b = e - a * P;
            const int lo = min(a, b), hi = max(a, b);
            double v = 0.0;
            for (int t = 0; t < T; ++t) v += CW[t * P + lo] * CW[t * P + hi];
            Gz[a * ldz + b] = v;
          }
          __syncthreads();
          if (warp == 0) warp_chol(Gz, P, ldz, 1e-26);
          __syncthreads();
          for (int t = tid; t < T; t += SP_THREADS) {
            for (int j = 0; j < P; ++j) {
              double v = CW[t * P + j];
              for (int i = 0; i < j; ++i) v -= CW[t * P + i] * Gz[i * ldz + j];
              const double rjj = Gz[j * ldz + j];
              CW[t * P + j] = rjj > 0.0 ? v / rjj : 0.0;
            }
          }
          __syncthreads();
        }
        for (int e = tid; e < T * P; e += SP_THREADS) W[e] = CW[e];
        __syncthreads();
      }
    }
    // F = W^T C (P x T);  Gz = F W (P x P, padded to pe)
    for (int e = tid; e < P * T; e += SP_THREADS) {
      const int a = e / T, t = e - a * T;
      double v = 0.0;
      for (int u = 0; u < T; ++u) v += W[u * P + a] * C[u * T + t];
      F[e] = v;
    }
    __syncthreads();
    // Gz = W^T C W = (M W)^T (M W) = Rc^T Rc: Q = M W Rc^-1 is the orthonormal basis of
    // the range (the reference's final QR); Bm = Q^T M = Rc^-T F
    for (int e = tid; e < P * P; e += SP_THREADS) {
      const int a = e / P, b = e - a * P;
      const int lo = min(a, b), hi = max(a, b);   // one summation order for both triangles
      double v = 0.0;
      for (int t = 0; t < T; ++t) v += F[lo * T + t] * W[t * P + hi];
      Ez[a * ldz + b] = v;
    }
    __syncthreads();
    if (warp == 0) warp_chol(Ez, P, ldz, 1e-12);
    __syncthreads();
    for (int t = tid; t < T; t += SP_THREADS) {
      for (int j = 0; j < P; ++j) {
        double v = F[j * T + t];
        for (int i = 0; i < j; ++i) v -= Ez[i * ldz + j] * Bm[i * T + t];
        const double rjj = Ez[j * ldz + j];
        Bm[j * T + t] = rjj > 0.0 ? v / rjj : 0.0;
      }
    }
    __syncthreads();
    // top left singular vector uh of Bm: eigenvector of Bm Bm^T (P x P)
    for (int e = tid; e < pe * pe; e += SP_THREADS) {
      const int a = e / pe, b = e - a * pe;
      double v = 0.0;
      if (a < P && b < P) {
        const int lo = min(a, b), hi = max(a, b);
        for (int t = 0; t < T; ++t) v += Bm[lo * T + t] * Bm[hi * T + t];
      }
      Gz[a * ldz + b] = v;
      E2[a * ldz + b] = (a == b && a < P) ? 1.0 : 0.0;
    }
    __syncthreads();
    if (warp == 0) jacobi_sym_t<true>(Gz, E2, P, ldz, sc);
    __syncthreads();
    if (tid == 0) {
      int kmax = 0;
      for (int k = 1; k < P; ++k)
        if (Gz[k * ldz + k] > Gz[kmax * ldz + kmax]) kmax = k;
      // yv = Rc^-1 uh  (back substitution; dropped pivots stay out)
      for (int k = P - 1; k >= 0; --k) {
        double v = E2[k * ldz + kmax];
        for (int j = k + 1; j < P; ++j) v -= Ez[k * ldz + j] * yv[j];
        const double rkk = Ez[k * ldz + k];
        yv[k] = rkk > 0.0 ? v / rkk : 0.0;
      }
    }
    __syncthreads();
    for (int t = tid; t < T; t += SP_THREADS) {
      double v = 0.0;
      for (int a = 0; a < P; ++a) v += W[t * P + a] * yv[a];
      wt[t] = v;
    }
    __syncthreads();
    // a = A wt, t = KA wt, normalise by |t|
    double nt2 = 0.0;
    for (int i = tid; i < S; i += SP_THREADS) {
      double av = 0.0, tvv = 0.0;
      for (int t = 0; t < T; ++t) {
        av += A[i * T + t] * wt[t];
        tvv += KA[i * T + t] * wt[t];
      }
      bv[i] = av;      // a, for the moment
      tv[i] = tvv;
      nt2 += tvv * tvv;
    }
    nt2 = block_sum(nt2, red);
    const double int_ = nt2 > 0.0 ? rsqrt(nt2) : 0.0;
    for (int i = tid; i < S; i += SP_THREADS) {
      Wc[(size_t)i * L + comp] = bv[i] * int_;
      tv[i] *= int_;
      Tm[(size_t)i * L + comp] = tv[i];
    }
    __syncthreads();
    // pctvar in Y:  q = Yp^T t  (t is centred, so Y0^T t = Yp^T t)
    for (int t = warp; t < T; t += SP_WARPS) {
      double v = 0.0;
      for (int i = lane; i < S; i += 32) v += Yr[(size_t)piy[i] * T + t] * tv[i];
      v = warp_sum(v);
      if (lane == 0) qv[t] = v;
    }
    __syncthreads();
    if (tid == 0) {
      double v = 0.0;
      for (int t = 0; t < T; ++t) v += qv[t] * qv[t];
      p.pct[(size_t)r * L + comp] = v / ssy;
    }
    // b = t;  g = Kx b
    for (int i = tid; i < S; i += SP_THREADS) bv[i] = tv[i];
    kx_matvec(p, S, pix, tv, gv);
    {
      double v = 0.0;
      for (int i = tid; i < S; i += SP_THREADS) v += gv[i];
      v = block_sum(v, red) / S;
      for (int i = tid; i < S; i += SP_THREADS) gv[i] -= v;
      __syncthreads();
    }
    // double modified Gram-Schmidt against the previous basis (Kx inner product)
    for (int pass = 0; pass < 2; ++pass)
      for (int j = 0; j < comp; ++j) {
        double cf = 0.0;
        for (int i = tid; i < S; i += SP_THREADS) cf += Gs[(size_t)i * L + j] * bv[i];
        cf = block_sum(cf, red);
        for (int i = tid; i < S; i += SP_THREADS) {
          bv[i] -= cf * Bs[(size_t)i * L + j];
          gv[i] -= cf * Gs[(size_t)i * L + j];
        }
        __syncthreads();
      }
    {
      double v = 0.0;
      for (int i = tid; i < S; i += SP_THREADS) v += bv[i] * gv[i];
      v = block_sum(v, red);
      const double s_ = v > 0.0 ? rsqrt(v) : 0.0;
      for (int i = tid; i < S; i += SP_THREADS) {
        bv[i] *= s_;
        gv[i] *= s_;
        Bs[(size_t)i * L + comp] = bv[i];
        Gs[(size_t)i * L + comp] = gv[i];
      }
      __syncthreads();
    }
    if (comp + 1 == L) break;
    // deflation (pyls/types/regression.py:143-147):  A' = A - b q^T with q = A^T g, then
    // A'' = A' - B_prev Z with Z = G_prev^T A' = G_prev^T A - (G_prev^T b) q^T.  One sweep
    // over A gives q and G_prev^T A, one read-modify-write sweep applies both terms; the
    // same terms keep KA = Kx A (Kx b = g, Kx B_prev = G_prev) and C = A^T Kx A
    // (b^T g = 1, B_prev^T Kx B_prev = I:  C'' = C - q q^T - Z^T Z) current.
    {
      const int ngrp = SP_THREADS / T, t = tid % T, grp = tid / T;
      for (int c0 = -1; c0 < comp; c0 += 8) {    // vectors c0 .. c0+7; -1 is the current g
        double acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.0;
        if (grp < ngrp)
          for (int i = grp; i < S; i += ngrp) {
            const double a = A[i * T + t];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int j = c0 + k;
              if (j < comp) acc[k] += (j < 0 ? gv[i] : Gs[(size_t)i * L + j]) * a;
            }
          }
        if (grp < ngrp) {
#pragma unroll
          for (int k = 0; k < 8; ++k) part[(grp * 8 + k) * T + t] = acc[k];
        }
        __syncthreads();
        for (int o = tid; o < 8 * T; o += SP_THREADS) {
          const int k = o / T, tt = o - k * T, j = c0 + k;
          if (j >= comp) continue;
          double v = 0.0;
          for (int g2 = 0; g2 < ngrp; ++g2) v += part[(g2 * 8 + k) * T + tt];
          if (j < 0) qv[tt] = v; else Z[j * T + tt] = v;
        }
        __syncthreads();
      }
    }
    if (comp > 0) {
      for (int j = warp; j < comp; j += SP_WARPS) {
        double v = 0.0;
        for (int i = lane; i < S; i += 32) v += Gs[(size_t)i * L + j] * bv[i];
        v = warp_sum(v);
        if (lane == 0) cg[j] = v;
      }
      __syncthreads();
      for (int o = tid; o < comp * T; o += SP_THREADS) Z[o] -= cg[o / T] * qv[o % T];
      __syncthreads();
    }
    for (int e = tid; e < S * T; e += SP_THREADS) {
      const int i = e / T, t = e - i * T;
      const double q_ = qv[t];
      double v = bv[i] * q_, w = gv[i] * q_;
      for (int j = 0; j < comp; ++j) {
        const double z = Z[j * T + t];
        v += Bs[(size_t)i * L + j] * z;
        w += Gs[(size_t)i * L + j] * z;
      }
      A[e] -= v;
      KA[e] -= w;
    }
    for (int e = tid; e < T * T; e += SP_THREADS) {
      const int t1 = e / T, t2 = e - t1 * T;
      double v = qv[t1] * qv[t2];
      for (int j = 0; j < comp; ++j) v += Z[j * T + t1] * Z[j * T + t2];
      C[e] -= v;
    }
    __syncthreads();
  }
  __syncthreads();
  if (!p.emit_ops) return;

  // ---- operands of x_weights = X0^T Wcoef (and, for bootstraps, signs + distrib) ----
  double *flip = Z;      // [0, L): signs, [L, 2L): column means of Kpi Wcoef
  for (int k = warp; k < L; k += SP_WARPS) {
    double v = 0.0;
    if (p.boot && p.So)
      for (int i = lane; i < S; i += 32) v += Wc[(size_t)i * L + k] * p.So[(size_t)pix[i] * L + k];
    v = warp_sum(v);
    if (lane == 0) flip[k] = (p.boot && p.So) ? (v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : 0.0)) : 1.0;
  }
  __syncthreads();
  // D[r*L + k][u] = flip_k * sum_{i: pix[i] = u} Wcoef[i][k]
  for (int u = tid; u < p.lda; u += SP_THREADS) {
    for (int k0 = 0; k0 < L; k0 += 8) {
      double v[8];
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) v[kk] = 0.0;
      if (u < Sf)
        for (int i = 0; i < S; ++i)
          if (pix[i] == u) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              if (k0 + kk < L) v[kk] += Wc[(size_t)i * L + k0 + kk];
          }
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
        if (k0 + kk < L) p.D[((size_t)r * L + k0 + kk) * p.lda + u] = v[kk] * flip[k0 + kk];
    }
  }
  if (!p.boot || !p.distrib) return;
  // distrib[t][k] = flip_k (Yp^T Tm + ysum (kappa^T Wcoef) / S),  kappa = Kpi 1
  kx_matvec(p, S, pix, nullptr, gv);
  for (int k = warp; k < L; k += SP_WARPS) {
    double v = 0.0;
    for (int i = lane; i < S; i += 32) v += gv[i] * Wc[(size_t)i * L + k];
    v = warp_sum(v);
    if (lane == 0) flip[L + k] = v / S;
  }
  __syncthreads();
  for (int o = warp; o < T * L; o += SP_WARPS) {
    const int t = o / L, k = o - t * L;
    double v = 0.0;
    for (int i = lane; i < S; i += 32)
      v += Yr[(size_t)piy[i] * T + t] * Tm[(size_t)i * L + k];
    v = warp_sum(v);
    if (lane == 0)
      p.distrib[(size_t)r * T * L + o] = flip[k] * (v + ysum[t] * flip[L + k]);
  }
}

size_t simpls_smem(int S, int T, int L) {
  const int P = std::min(T, SP_PROBES), pe = P + (P & 1), ldz = pe | 1, halfz = pe / 2;
  size_t d = 3 * (size_t)S + (size_t)T * T + 4 * (size_t)T * P +
             3 * (size_t)pe * ldz + (size_t)std::max(L * T, 2 * L) + 3 * pe + 4 * (size_t)T +
             SP_WARPS + 2 + (size_t)L + 8 * SP_THREADS + 3 * halfz;
  size_t b = d * sizeof(double) + sizeof(int) * (2 * halfz + 2 * (size_t)S) +
             sizeof(short2) * (halfz * (halfz + 1) / 2) + 16;
  return b;
}

// x_weights (B, L) from the GEMM output rows (L, ldr) with sklearn's svd_flip
// sign rule (largest-magnitude entry of every column positive): block per column
__global__ void xweights_flip_kernel(const double *__restrict__ Rw, long long ldr, int B, int L,
                                     double *__restrict__ xw) {
  __shared__ double s_best[32];
  __shared__ int s_idx[32];
  const int k = blockIdx.x;
  const double *row = Rw + (size_t)k * ldr;
  double best = -1.0;
  int bidx = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const double a = fabs(row[b]);
    if (a > best) {
      best = a;
      bidx = b;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (ob > best || (ob == best && oi < bidx)) {
      best = ob;
      bidx = oi;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) {
    s_best[warp] = best;
    s_idx[warp] = bidx;
  }
  __syncthreads();
  best = s_best[0];
  bidx = s_idx[0];
  for (int w = 1; w < nw; ++w)
    if (s_best[w] > best || (s_best[w] == best && s_idx[w] < bidx)) {
      best = s_best[w];
      bidx = s_idx[w];
    }
  const double sgn = row[bidx] < 0.0 ? -1.0 : 1.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) xw[(size_t)b * L + k] = sgn * row[b];
}

__global__ void transpose_kernel(const double *__restrict__ in, int rows, int cols, int ld_in,
                                 double *__restrict__ out) {   // out (cols, rows) contiguous
  __shared__ double tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)r * ld_in + c] : 0.0;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[(size_t)c * rows + r] = tile[threadIdx.x][i];
  }
}

// out = U with every column centred (block per column)
__global__ void colcenter_kernel(const double *__restrict__ U, int B, int L,
                                 double *__restrict__ out) {
  __shared__ double red[32];
  const int k = blockIdx.x;
  double v = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) v += U[(size_t)b * L + k];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double m = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m += red[w];
  m /= B;
  for (int b = threadIdx.x; b < B; b += blockDim.x)
    out[(size_t)b * L + k] = U[(size_t)b * L + k] - m;
}

__global__ void identity_blocks_kernel(double *M, int n, int L, int ldm) {   // (n, L, ldm)
  const size_t total = (size_t)n * L * ldm;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)((e / ldm) % L), j = (int)(e % ldm);
    M[e] = i == j ? 1.0 : 0.0;
  }
}

}  // namespace

int launch_transpose(plsb_ctx *h, const double *in, int rows, int cols, int ld_in, double *out,
                     cudaStream_t st) {
  KernelTimer kt(h, KC_PREP, st);
  dim3 grid(cdiv(cols, 32), cdiv(rows, 32)), block(32, 8);
  transpose_kernel<<<grid, block, 0, st>>>(in, rows, cols, ld_in, out);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_colcenter(plsb_ctx *h, const double *U, int B, int L, double *out, cudaStream_t st) {
  KernelTimer kt(h, KC_PREP, st);
  colcenter_kernel<<<L, 256, 0, st>>>(U, B, L, out);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_identity_blocks(plsb_ctx *h, double *M, int n, int L, int ldm, cudaStream_t st) {
  KernelTimer kt(h, KC_PREP, st);
  const size_t total = (size_t)n * L * ldm;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)h->sm_count * 8);
  identity_blocks_kernel<<<blocks, 256, 0, st>>>(M, n, L, ldm);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_xweights_flip(plsb_ctx *h, const double *Rw, long long ldr, int B, int L, double *xw,
                         cudaStream_t st) {
  KernelTimer kt(h, KC_PREP, st);
  xweights_flip_kernel<<<L, 256, 0, st>>>(Rw, ldr, B, L, xw);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

// Runs the SIMPLS component loop for `count` resamples.  Scratch (Wcoef, Bs, Gs,
// Tm: count*S*L doubles each) lives in h->G/H/M/lam; D (count*L rows x S_pad)
// in h->A when emit_ops.
int launch_simpls(plsb_ctx *h, const int32_t *idx, int count, int boot, int emit_ops,
                  const double *omega, long long om_stride_r, long long om_stride_c, double *pct,
                  double *distrib, cudaStream_t st, const double *yres) {
  KernelTimer kt(h, KC_SMALL, st);
  const Layout &l = h->lay;
  if (count <= 0) return PLSB_OK;
  PLSB_CHECK(l.T <= 32, PLSB_ERR_ARG, "SIMPLS engine supports at most 32 behaviours (T=%d)", l.T);
  const size_t smem = simpls_smem(l.S, l.T, l.L);
  PLSB_CHECK(smem <= 227 * 1024, PLSB_ERR_ARG,
             "SIMPLS engine needs %zu bytes of shared memory per resample (S*T too large)", smem);
  const size_t per = (size_t)l.S * l.L;
  PLSB_TRY(h->G.ensure(sizeof(double) * per * count));
  PLSB_TRY(h->H.ensure(sizeof(double) * per * count));
  PLSB_TRY(h->M.ensure(sizeof(double) * per * count));
  PLSB_TRY(h->lam.ensure(sizeof(double) * per * count));
  PLSB_TRY(h->S1.ensure(sizeof(double) * (size_t)l.S * l.T * count));
  PLSB_TRY(h->S2.ensure(sizeof(double) * (size_t)l.S * l.T * count));
  SimplsParams p;
  p.S = l.S; p.T = l.T; p.L = l.L;
  p.p = std::min(l.T, SP_PROBES);
  p.n_iter = (1 < 0.1 * std::min(l.B, l.T)) ? 7 : 4;   // sklearn: n_iter='auto', n_components=1
  p.boot = boot; p.emit_ops = emit_ops; p.lda = l.S_pad;
  p.Kraw = h->Cmat.as<double>();
  p.Yc = h->Y.as<double>();
  p.Yres = yres;
  p.omega = omega; p.om_stride_r = om_stride_r; p.om_stride_c = om_stride_c;
  p.So = h->has_original ? h->Sx.as<double>() : nullptr;
  p.idx = idx;
  p.valid = h->has_rowmask ? h->rowmask.as<int32_t>() : nullptr;
  p.Wcoef = h->G.as<double>(); p.Bs = h->H.as<double>(); p.Gs = h->M.as<double>();
  p.Tm = h->lam.as<double>();
  p.Abuf = h->S1.as<double>();
  p.KAbuf = h->S2.as<double>();
  p.pct = pct; p.D = h->A.as<double>(); p.distrib = distrib;
  PLSB_CHECK(l.T <= SP_PROBES || omega != nullptr, PLSB_ERR_ARG, "SIMPLS: missing Omega table");
#define PLSB_SP(TT)                                                                            \
  do {                                                                                         \
    PLSB_CUDA(cudaFuncSetAttribute(simpls_kernel<TT>,                                          \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    simpls_kernel<TT><<<count, SP_THREADS, smem, st>>>(p);                                     \
  } while (0)
  if (l.T <= 8) PLSB_SP(8);
  else if (l.T <= 16) PLSB_SP(16);
  else if (l.T <= 24) PLSB_SP(24);
  else PLSB_SP(32);
#undef PLSB_SP
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

}  // namespace plsb
