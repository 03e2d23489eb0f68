// Cyclic two-sided Jacobi eigen-solver for small symmetric matrices held in
// shared memory, run cooperatively by all threads of a CTA (see small_matrix.cu
// for the scheme).  Shared by the bootstrap decomposition and the SIMPLS kernel.
#pragma once

#include <cuda_runtime.h>

namespace plsb {

constexpr int MAX_SWEEPS = 30;
constexpr double JACOBI_TOL = 1e-14;
constexpr double JACOBI_EPS = 1e-15;

struct JacobiScratch {
  short2 *blk;   // (a, b) of every upper-triangular block, a <= b   [half*(half+1)/2]
  int *pq;       // (p, q) of every pair of the current round         [2*half]
  double *cst;   // (c, s, t) of every pair of the current round      [3*half]
};

// Diagonalises the symmetric matrix A in place: A <- J^T A J, V <- V J over
// all rotations.  A and V are row-major with leading dimension ld; n is the
// order, ne = n rounded up to even.  For odd n the extra index n is a dummy
// player: row / column n of A and column n of V must be zero on entry (they
// stay zero).  WARP = false: called by every thread of the CTA (CTA barriers);
// WARP = true: called by the 32 lanes of ONE warp (warp barriers only), for
// matrices small enough that CTA-wide barriers would dominate.
template <bool WARP>
static __device__ void jacobi_sym_t(double *A, double *V, int n, int ld, const JacobiScratch &sc) {
  const int tid = WARP ? (threadIdx.x & 31) : threadIdx.x;
  const int nthreads = WARP ? 32 : blockDim.x;
  const int ne = n + (n & 1);
  const int half = ne / 2;
  const int nb = half * (half + 1) / 2;
  if (WARP) __syncwarp(); else __syncthreads();
  if (n < 2) return;
  for (int sweep = 0; sweep < MAX_SWEEPS; ++sweep) {
    int any = 0;
    for (int round = 0; round < ne - 1; ++round) {
      // phase 0: rotation of every pair from its diagonal block
      int active = 0;
      for (int pi = tid; pi < half; pi += nthreads) {
        int a, b;
        if (pi == 0) {
          a = ne - 1;
          b = round;
        } else {
          a = (round + pi) % (ne - 1);
          b = (round - pi + (ne - 1)) % (ne - 1);
        }
        const int p = min(a, b), q = max(a, b);
        const double app = A[p * ld + p], aqq = A[q * ld + q], apq = A[p * ld + q];
        double c = 1.0, s = 0.0, t = 0.0;
        // rotate while the off-diagonal entry is above both the relative
        // (graded-matrix) threshold and the rounding level of the larger
        // diagonal entry -- below that a rotation only shuffles noise
        const double thr = fmax(JACOBI_TOL * sqrt(fabs(app * aqq)),
                                JACOBI_EPS * fmax(fabs(app), fabs(aqq)));
        if (fabs(apq) > thr) {
          const double zeta = (aqq - app) / (2.0 * apq);
          t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          c = rsqrt(1.0 + t * t);
          s = c * t;
          active = 1;
        }
        sc.pq[2 * pi] = p;
        sc.pq[2 * pi + 1] = q;
        sc.cst[3 * pi] = c;
        sc.cst[3 * pi + 1] = s;
        sc.cst[3 * pi + 2] = t;
      }
      if (WARP) { __syncwarp(); any |= __any_sync(0xffffffffu, active); }
      else any |= __syncthreads_or(active);
      // phase 1a: the 2 x 2 blocks of A
      for (int bi = tid; bi < nb; bi += nthreads) {
        const int a = sc.blk[bi].x, b = sc.blk[bi].y;
        const int pa = sc.pq[2 * a], qa = sc.pq[2 * a + 1];
        const double ca = sc.cst[3 * a], sa = sc.cst[3 * a + 1];
        if (a == b) {
          if (sa != 0.0) {
            const double t = sc.cst[3 * a + 2], apq = A[pa * ld + qa];
            A[pa * ld + pa] -= t * apq;
            A[qa * ld + qa] += t * apq;
            A[pa * ld + qa] = 0.0;
            A[qa * ld + pa] = 0.0;
          }
          continue;
        }
        const int pb = sc.pq[2 * b], qb = sc.pq[2 * b + 1];
        const double cb = sc.cst[3 * b], sb = sc.cst[3 * b + 1];
        if (sa == 0.0 && sb == 0.0) continue;
        const double x00 = A[pa * ld + pb], x01 = A[pa * ld + qb];
        const double x10 = A[qa * ld + pb], x11 = A[qa * ld + qb];
        const double y00 = cb * x00 - sb * x01, y01 = sb * x00 + cb * x01;
        const double y10 = cb * x10 - sb * x11, y11 = sb * x10 + cb * x11;
        const double z00 = ca * y00 - sa * y10, z01 = ca * y01 - sa * y11;
        const double z10 = sa * y00 + ca * y10, z11 = sa * y01 + ca * y11;
        A[pa * ld + pb] = z00;
        A[pa * ld + qb] = z01;
        A[qa * ld + pb] = z10;
        A[qa * ld + qb] = z11;
        A[pb * ld + pa] = z00;
        A[qb * ld + pa] = z01;
        A[pb * ld + qa] = z10;
        A[qb * ld + qa] = z11;
      }
      // phase 1b: V <- V J.  CTA: a warp per pair, lanes over the rows (conflict-free
      // columns); single warp: one (pair, row) item per lane
      if (WARP) {
        for (int e = tid; e < half * n; e += 32) {
          const int pi = e / n, i = e - pi * n;
          const double s = sc.cst[3 * pi + 1];
          if (s == 0.0) continue;
          const double c = sc.cst[3 * pi];
          const int p = sc.pq[2 * pi], q = sc.pq[2 * pi + 1];
          const double vx = V[i * ld + p], vy = V[i * ld + q];
          V[i * ld + p] = c * vx - s * vy;
          V[i * ld + q] = s * vx + c * vy;
        }
      } else {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = nthreads >> 5;
        for (int pi = warp; pi < half; pi += nwarps) {
          const double s = sc.cst[3 * pi + 1];
          if (s == 0.0) continue;
          const double c = sc.cst[3 * pi];
          const int p = sc.pq[2 * pi], q = sc.pq[2 * pi + 1];
          for (int i = lane; i < n; i += 32) {
            const double vx = V[i * ld + p], vy = V[i * ld + q];
            V[i * ld + p] = c * vx - s * vy;
            V[i * ld + q] = s * vx + c * vy;
          }
        }
      }
      if (WARP) __syncwarp(); else __syncthreads();
    }
    if (!any) break;
  }
}

static __device__ __forceinline__ void jacobi_sym(double *A, double *V, int n, int ld,
                                                  const JacobiScratch &sc) {
  jacobi_sym_t<false>(A, V, n, ld, sc);
}

}  // namespace plsb
