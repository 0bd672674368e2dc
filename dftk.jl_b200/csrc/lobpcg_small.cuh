// Fused small-matrix path of the LOBPCG block algebra (M <= 32 bands, i.e. the C1/C2/C4/C5-sized problems where every
// step is launch- and sync-latency bound).  The tensor-core GEMM + cuSOLVER sequence of the large path
// (gram -> hermitize -> copy -> potrf -> zero_lower -> copy -> trtri -> 3x stats, five host syncs per Cholesky pass)
// becomes gram -> k_small_chol -> k_small_rmul with ONE host sync; block-list products are one launch each.
// Same algorithm as lobpcg.cu (reference: src/eigen/lobpcg_hyper_impl.jl:90-137 LazyHcat products, :190-210
// safe_cholesky, :212 normest, :216-261 ortho!).  Bodies are __host__ __device__ (host emulation in tests/hostemu);
// on the host TLOOP runs the "threads" of a CTA one after the other, which is valid because no phase between two
// TSYNC() has dependencies between its iterations.
#pragma once
#include <float.h>
#include <math.h>
#include "fft_core.cuh"

namespace dftk {

#define SMALL_MAX_N 32      // columns of the block being orthogonalised / produced (= max bands of the small path)
#define SMALL_MAX_COLS 96   // total columns of a block list [X R P]
#define SMALL_TR 8          // rows per shared-memory tile of the Gram kernel (192 columns x 8 rows x 16 B = 24 KB)
#define SMALL_RED 256       // logical reduction width (= CTA size of k_small_chol)

struct SmallMatList {
  const cplx* p[3];
  long long ld[3];
  int cols[3];
  int start[4];   // column offsets, start[n] = total
  int n;
};
HD int small_block_of(const SmallMatList& L, int col) { return col >= L.start[2] && L.n > 2 ? 2 : (col >= L.start[1] && L.n > 1 ? 1 : 0); }
HD const cplx* small_col_ptr(const SmallMatList& L, int col) {
  const int b = small_block_of(L, col);
  return L.p[b] + L.ld[b] * (long long)(col - L.start[b]);
}

// ---- Gram: partial[cta][i + nA j] = sum over the CTA's rows of conj(A[r,i]) B[r,j]; (i,j) pairs with block(j) < block(i)
//      are skipped when upper_only (Hermitian result, only the block upper triangle is consumed)
HD void small_gram_cta(int cta, long long rows_per_cta, long long n_rows, const SmallMatList& A, const SmallMatList& B,
                       int upper_only, cplx* __restrict__ ws, cplx* sm) {
  const int nA = A.start[A.n], nB = B.start[B.n], total = nA * nB;
  const long long r_begin = (long long)cta * rows_per_cta;
  const long long r_end = r_begin + rows_per_cta < n_rows ? r_begin + rows_per_cta : n_rows;
  cplx* As = sm;
  cplx* Bs = sm + SMALL_TR * nA;
  cplx* out = ws + (size_t)cta * total;
  TLOOP(o, total) out[o] = make_double2(0.0, 0.0);
  for (long long r0 = r_begin; r0 < r_end; r0 += SMALL_TR) {
    const int nr = (int)(r_end - r0 < SMALL_TR ? r_end - r0 : SMALL_TR);
    TSYNC();
    TLOOP(e, SMALL_TR * nA) {
      const int r = e % SMALL_TR, i = e / SMALL_TR;
      As[r * nA + i] = r < nr ? small_col_ptr(A, i)[r0 + r] : make_double2(0.0, 0.0);
    }
    TLOOP(e, SMALL_TR * nB) {
      const int r = e % SMALL_TR, j = e / SMALL_TR;
      Bs[r * nB + j] = r < nr ? small_col_ptr(B, j)[r0 + r] : make_double2(0.0, 0.0);
    }
    TSYNC();
    // register tile of 1 x 4 outputs per thread: one shared-memory read of a (consecutive threads, consecutive i) serves four
    // columns of B (the same address for the whole warp: broadcast) -- the plain one-output form was bound by shared-memory
    // reads at two 16-byte loads per four FMAs.  Each output still accumulates its rows in ascending order.
    const int nJt = (nB + 3) / 4;
    TLOOP(t, nA * nJt) {
      const int i = t % nA, j0 = (t / nA) * 4;
      const int bi = small_block_of(A, i);
      int jq[4];
      bool act[4];
      bool any = false;
      for (int q = 0; q < 4; ++q) {
        const int j = j0 + q;
        act[q] = j < nB && !(upper_only && small_block_of(B, j) < bi);
        jq[q] = j < nB ? j : nB - 1;          // inactive lanes read a valid element and drop the result
        any = any || act[q];
      }
      if (!any) continue;
      double ax[4] = {0.0, 0.0, 0.0, 0.0}, ay[4] = {0.0, 0.0, 0.0, 0.0};
      for (int r = 0; r < nr; ++r) {
        const cplx a = As[r * nA + i];
        const cplx* brow = Bs + r * nB;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int q = 0; q < 4; ++q) {
          const cplx b = brow[jq[q]];
          ax[q] += a.x * b.x + a.y * b.y;     // conj(a) * b
          ay[q] += a.x * b.y - a.y * b.x;
        }
      }
      for (int q = 0; q < 4; ++q)
        if (act[q]) {
          const int o = i + nA * (j0 + q);
          out[o].x += ax[q];
          out[o].y += ay[q];
        }
    }
  }
}
// fixed-order sum of the CTA partials into C (run by the last CTA to finish / by the emulator)
HD void small_gram_reduce(int n_ctas, const SmallMatList& A, const SmallMatList& B, int upper_only,
                          const cplx* __restrict__ ws, cplx* __restrict__ C, long long ldc) {
  const int nA = A.start[A.n], nB = B.start[B.n], total = nA * nB;
  TLOOP(o, total) {
    const int i = o % nA, j = o / nA;
    if (upper_only && small_block_of(B, j) < small_block_of(A, i)) continue;
    double sx = 0.0, sy = 0.0;
    for (int c = 0; c < n_ctas; ++c) {
#ifdef __CUDA_ARCH__
      const double2 v = __ldcg(ws + (size_t)c * total + o);   // written by other CTAs: bypass L1
#else
      const cplx v = ws[(size_t)c * total + o];
#endif
      sx += v.x;
      sy += v.y;
    }
    C[i + ldc * j] = make_double2(sx, sy);
  }
}

// ---- out[r, c] = alpha * sum_l Y[r, l] cm[l, c] + beta * out[r, c]  for one row r (cm: ny x ncols, leading dim ldcm)
// The accumulators of CT columns live in registers (fully unrolled, compile-time indices): a runtime-indexed array of
// SMALL_MAX_N accumulators sits in local memory and makes every FMA a load + store.  Columns beyond the block's count
// compute on a clamped column and are dropped; each output still sums l = 0 .. ny-1 in ascending order.
template <int CT>
HD void small_blocks_times_chunk(long long r, const SmallMatList& Y, const cplx* __restrict__ cm, int ldcm, int c0, int nc,
                                 cplx* __restrict__ out, long long ldo, double alpha, double beta) {
  double ax[CT], ay[CT];
  int off[CT];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int q = 0; q < CT; ++q) {
    ax[q] = ay[q] = 0.0;
    off[q] = ldcm * (c0 + (q < nc ? q : nc - 1));
  }
  const int ny = Y.start[Y.n];
  for (int l = 0; l < ny; ++l) {
    const cplx y = small_col_ptr(Y, l)[r];
    const cplx* mrow = cm + l;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int q = 0; q < CT; ++q) {
      const cplx m = mrow[off[q]];
      ax[q] += y.x * m.x - y.y * m.y;
      ay[q] += y.x * m.y + y.y * m.x;
    }
  }
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int q = 0; q < CT; ++q) {
    if (q >= nc) continue;
    const int c = c0 + q;
    cplx o = make_double2(alpha * ax[q], alpha * ay[q]);
    if (beta != 0.0) {
      const cplx p = out[r + ldo * c];
      o.x += beta * p.x;
      o.y += beta * p.y;
    }
    out[r + ldo * c] = o;
  }
}
HD void small_blocks_times_row(long long r, const SmallMatList& Y, const cplx* __restrict__ cm, int ldcm, int ncols,
                               cplx* __restrict__ out, long long ldo, double alpha, double beta) {
  int c0 = 0;
  while (ncols - c0 > 8) {
    const int nc = ncols - c0 < 16 ? ncols - c0 : 16;
    small_blocks_times_chunk<16>(r, Y, cm, ldcm, c0, nc, out, ldo, alpha, beta);
    c0 += nc;
  }
  if (ncols - c0 > 0) small_blocks_times_chunk<8>(r, Y, cm, ldcm, c0, ncols - c0, out, ldo, alpha, beta);
}

// ---- X[r, :] <- X[r, :] * invR (upper triangular n x n, column-major with leading dimension ldr), in place.
// Output columns are produced in chunks of 8 from the highest down, the accumulators of a chunk in registers: column j needs
// the ORIGINAL x[l] for l <= j only, and a chunk is written after all its sums are complete, so lower chunks still read
// originals.  Each output sums l = 0 .. j in ascending order (as the plain row-cached form did).
HD void small_rmul_row(long long r, cplx* __restrict__ X, long long ld, int n, const cplx* __restrict__ Rinv, int ldr) {
  for (int j0 = ((n - 1) / 8) * 8; j0 >= 0; j0 -= 8) {
    double sx[8], sy[8];
    int jq[8];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int q = 0; q < 8; ++q) {
      sx[q] = sy[q] = 0.0;
      jq[q] = j0 + q < n ? j0 + q : n - 1;       // columns beyond n compute on a valid one and are dropped
    }
    const int lmax = j0 + 7 < n - 1 ? j0 + 7 : n - 1;
    for (int l = 0; l <= lmax; ++l) {
      const cplx x = X[r + ld * l];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
      for (int q = 0; q < 8; ++q) {
        if (l <= j0 + q) {                         // upper triangular: row l enters the columns j >= l
          const cplx m = Rinv[l + ldr * jq[q]];
          sx[q] += x.x * m.x - x.y * m.y;
          sy[q] += x.x * m.y + x.y * m.x;
        }
      }
    }
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int q = 0; q < 8; ++q)
      if (j0 + q < n) X[r + ld * (j0 + q)] = make_double2(sx[q], sy[q]);
  }
}

// ---- safe_cholesky + inverse + normest in one CTA (n <= SMALL_MAX_N).
// O: Hermitian, only the upper triangle (i <= j) is read.  Writes invR (n x n, zeros below the diagonal) and
// stats[0] = number of Cholesky attempts (0: all five failed), stats[1] = normest(invR), stats[2] = normest(R),
// stats[3] = ||O||_F.  Shared memory: As, Bs (n*n cplx each), red (SMALL_RED doubles), flag (2 ints).
HD void small_chol_cta(const cplx* __restrict__ O, long long ldo, int n, cplx* __restrict__ invR, long long ldi,
                       double* __restrict__ stats, cplx* As, cplx* Bs, double* red, int* flag) {
  // Frobenius norm of the full Hermitian matrix
  TLOOP(t, SMALL_RED) {
    double s = 0.0;
    for (int e = t; e < n * n; e += SMALL_RED) {
      const int i = e % n, j = e / n;
      const cplx v = i <= j ? O[i + ldo * j] : O[j + ldo * i];
      s += (i == j) ? v.x * v.x : v.x * v.x + v.y * v.y;
    }
    red[t] = s;
  }
  TSYNC();
  double onorm = 0.0;
  for (int t = 0; t < SMALL_RED; ++t) onorm += red[t];
  onorm = sqrt(onorm);
  TSYNC();
  double shift = 0.0, alpha = 100.0;
  int nchol = 0, ok = 0;
  while (nchol < 5 && !ok) {
    nchol++;
    // A = Hermitian(O) + shift I   (upper triangle only is needed)
    TLOOP(e, n * n) {
      const int i = e % n, j = e / n;
      cplx v = make_double2(0.0, 0.0);
      if (i <= j) v = O[i + ldo * j];
      if (i == j) v = make_double2(v.x + shift, 0.0);
      As[e] = v;
    }
    TLOOP(t, 1) flag[0] = 0;
    TSYNC();
    for (int k = 0; k < n; ++k) {
      TLOOP(t, 1) {
        const double d = As[k + n * k].x;
        if (!(d > 0.0) || !isfinite(d)) flag[0] = 1;
        else As[k + n * k] = make_double2(sqrt(d), 0.0);
      }
      TSYNC();
      if (flag[0]) break;
      const double rinv = 1.0 / As[k + n * k].x;
      const int m = n - k - 1;
      TLOOP(j, m) {
        cplx v = As[k + n * (k + 1 + j)];
        As[k + n * (k + 1 + j)] = make_double2(v.x * rinv, v.y * rinv);
      }
      TSYNC();
      TLOOP(e, m * m) {
        const int i = k + 1 + e % m, j = k + 1 + e / m;
        if (i <= j) {
          const cplx a = As[k + n * i], b = As[k + n * j];   // conj(R[k,i]) * R[k,j]
          cplx v = As[i + n * j];
          v.x -= a.x * b.x + a.y * b.y;
          v.y -= a.x * b.y - a.y * b.x;
          As[i + n * j] = v;
        }
      }
      TSYNC();
    }
    int failed = flag[0];
    TSYNC();
    if (!failed) {
      // invR by back substitution, one column per thread:  R x = e_j
      TLOOP(j, n) {
        for (int i = n - 1; i > j; --i) Bs[i + n * j] = make_double2(0.0, 0.0);
        Bs[j + n * j] = make_double2(1.0 / As[j + n * j].x, 0.0);
        for (int i = j - 1; i >= 0; --i) {
          double sx = 0.0, sy = 0.0;
          for (int l = i + 1; l <= j; ++l) {
            const cplx a = As[i + n * l], b = Bs[l + n * j];
            sx += a.x * b.x - a.y * b.y;
            sy += a.x * b.y + a.y * b.x;
          }
          const double d = -1.0 / As[i + n * i].x;
          Bs[i + n * j] = make_double2(sx * d, sy * d);
        }
      }
      TSYNC();
      // any non-finite entry of invR counts as a failed factorisation (@assert !any(isnan, invR), :198)
      TLOOP(t, SMALL_RED) {
        double bad = 0.0;
        for (int e = t; e < n * n; e += SMALL_RED)
          if (!isfinite(Bs[e].x) || !isfinite(Bs[e].y)) bad += 1.0;
        red[t] = bad;
      }
      TSYNC();
      double bad = 0.0;
      for (int t = 0; t < SMALL_RED; ++t) bad += red[t];
      TSYNC();
      failed = bad != 0.0;
    }
    if (!failed) {
      ok = 1;
    } else {
      // O += alpha eps ||O|| I, alpha *= 10   (:203-205)
      shift += alpha * DBL_EPSILON * onorm;
      alpha *= 10.0;
    }
  }
  // normest(M) = max |diag| + ||M - Diag||_F  for invR (Bs) and R (As, upper triangle)
  double out1 = 0.0, out2 = 0.0;
  if (ok) {
    for (int which = 0; which < 2; ++which) {
      const cplx* Mx = which == 0 ? Bs : As;
      TLOOP(t, SMALL_RED) {
        double s = 0.0;
        for (int e = t; e < n * n; e += SMALL_RED) {
          const int i = e % n, j = e / n;
          if (i < j) s += Mx[e].x * Mx[e].x + Mx[e].y * Mx[e].y;
        }
        red[t] = s;
      }
      TSYNC();
      double so = 0.0, md = 0.0;
      for (int t = 0; t < SMALL_RED; ++t) so += red[t];
      for (int i = 0; i < n; ++i) md = fmax(md, fabs(Mx[i + n * i].x));
      TSYNC();
      if (which == 0) out1 = md + sqrt(so);
      else out2 = md + sqrt(so);
    }
    TLOOP(e, n * n) invR[e % n + ldi * (e / n)] = Bs[e];
  }
  TLOOP(t, 1) {
    stats[0] = ok ? (double)nchol : 0.0;
    stats[1] = out1;
    stats[2] = out2;
    stats[3] = onorm;
  }
}

// ---- Hermitian eigensolver of the Rayleigh-Ritz step (eigen(Hermitian(XAX)), lobpcg_hyper_impl.jl:141-171) for n <=
// SMALL_MAX_COLS in one CTA: cyclic two-sided Jacobi with a round-robin ("chess tournament") ordering, n/2 disjoint
// rotations per step.  A rotation of the pair (p, q) with A[p,q] = |b| e^{i phi} is J = diag(1, e^{-i phi}) R(c, s);
// one step applies A <- J^H (A J) for all pairs at once (column phase, then row phase) and V <- V J.
// G: n x n Hermitian, only the upper triangle (i <= j) is read; on return G holds the eigenvectors (columns, ascending
// eigenvalues) and w the eigenvalues.  As: n*n cplx (shared), V: n*n cplx scratch (global, leading dimension n),
// rot: 2*(n/2+1) cplx (shared; {c, s} and the phase of each pair), red: SMALL_RED doubles, iwork: 2*(n/2+1) ints (shared).
// stats[0] = sweeps used (0: no convergence within the cap), stats[1] = final off-diagonal Frobenius norm.
// lam_out (optional): the lowest n_keep eigenvalues once more, where the caller keeps the Ritz values of its active block.
HD void small_heev_cta(cplx* __restrict__ G, long long ldg, int n, double* __restrict__ w, cplx* As, cplx* __restrict__ V,
                       cplx* rot, double* red, int* iwork, double* __restrict__ stats, double* __restrict__ lam_out, int n_keep) {
  const int m = (n + 1) & ~1, half = m / 2;
  int* pp = iwork;
  int* qq = iwork + half;
  TLOOP(e, n * n) {
    const int i = e % n, j = e / n;
    cplx v = i <= j ? G[i + ldg * j] : G[j + ldg * i];
    if (i > j) v.y = -v.y;
    if (i == j) v.y = 0.0;
    As[e] = v;
    V[e] = make_double2(i == j ? 1.0 : 0.0, 0.0);
  }
  TSYNC();
  TLOOP(t, SMALL_RED) {
    double s = 0.0;
    for (int e = t; e < n * n; e += SMALL_RED) s += As[e].x * As[e].x + As[e].y * As[e].y;
    red[t] = s;
  }
  TSYNC();
  double normA = 0.0;
  for (int t = 0; t < SMALL_RED; ++t) normA += red[t];
  normA = sqrt(normA);
  TSYNC();
  const double tiny = 1.1102230246251565e-16 * 0.0078125 * normA;   // rotations below eps/128 ||A|| are skipped
  // converged: off-diagonal norm at the rounding floor of the updates (or stagnating just above it)
  const double off_tol = 2.0 * sqrt((double)n) * DBL_EPSILON * normA;
  int sweeps = 0;
  double off = 0.0, off_prev = -1.0;
  bool done = (n <= 1) || !(normA > 0.0);
  const int max_sweeps = 60;
  while (!done && sweeps < max_sweeps) {
    sweeps++;
    for (int r = 0; r < m - 1; ++r) {
      // pairs of this step: (m-1, r) and ((r+k) mod (m-1), (r-k) mod (m-1)), k = 1 .. m/2-1; indices >= n are byes
      TLOOP(k, half) {
        int a = k == 0 ? m - 1 : (r + k) % (m - 1);
        int b = k == 0 ? r : (r - k + (m - 1)) % (m - 1);
        int p = a < b ? a : b, q = a < b ? b : a;
        cplx cs = make_double2(1.0, 0.0), ph = make_double2(1.0, 0.0);
        if (q >= n) {
          p = q = -1;
        } else {
          const cplx bq = As[p + n * q];
          const double ab = sqrt(bq.x * bq.x + bq.y * bq.y);
          if (ab > tiny) {
            const double tau = (As[q + n * q].x - As[p + n * p].x) / (2.0 * ab);
            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            const double c = 1.0 / sqrt(1.0 + t * t);
            cs = make_double2(c, t * c);
            ph = make_double2(bq.x / ab, -bq.y / ab);     // e^{-i phi}
          } else {
            p = q = -1;
          }
        }
        pp[k] = p;
        qq[k] = q;
        rot[k] = cs;
        rot[half + k] = ph;
      }
      TSYNC();
      // columns: [a_p a_q] <- [c a_p - s e^{-i phi} a_q,  s a_p + c e^{-i phi} a_q], same for V
      TLOOP(e, half * n) {
        const int k = e / n, i = e % n;
        const int p = pp[k], q = qq[k];
        if (p >= 0) {
          const double c = rot[k].x, s = rot[k].y;
          const cplx ph = rot[half + k];
          const cplx ap = As[i + n * p], aq = cmul(ph, As[i + n * q]);
          As[i + n * p] = make_double2(c * ap.x - s * aq.x, c * ap.y - s * aq.y);
          As[i + n * q] = make_double2(s * ap.x + c * aq.x, s * ap.y + c * aq.y);
          const cplx vp = V[i + n * p], vq = cmul(ph, V[i + n * q]);
          V[i + n * p] = make_double2(c * vp.x - s * vq.x, c * vp.y - s * vq.y);
          V[i + n * q] = make_double2(s * vp.x + c * vq.x, s * vp.y + c * vq.y);
        }
      }
      TSYNC();
      // rows: [a_p; a_q] <- [c a_p - s e^{+i phi} a_q;  s a_p + c e^{+i phi} a_q]
      TLOOP(e, half * n) {
        const int k = e / n, j = e % n;
        const int p = pp[k], q = qq[k];
        if (p >= 0) {
          const double c = rot[k].x, s = rot[k].y;
          const cplx ph = make_double2(rot[half + k].x, -rot[half + k].y);
          const cplx ap = As[p + n * j], aq = cmul(ph, As[q + n * j]);
          As[p + n * j] = make_double2(c * ap.x - s * aq.x, c * ap.y - s * aq.y);
          As[q + n * j] = make_double2(s * ap.x + c * aq.x, s * ap.y + c * aq.y);
        }
      }
      TSYNC();
    }
    TLOOP(t, SMALL_RED) {
      double s = 0.0;
      for (int e = t; e < n * n; e += SMALL_RED)
        if (e % n != e / n) s += As[e].x * As[e].x + As[e].y * As[e].y;
      red[t] = s;
    }
    TSYNC();
    off = 0.0;
    for (int t = 0; t < SMALL_RED; ++t) off += red[t];
    off = sqrt(off);
    TSYNC();
    done = off <= off_tol || (off_prev >= 0.0 && off <= 1e-12 * normA && off >= 0.25 * off_prev);
    off_prev = off;
  }
  // ascending order by rank counting (ties keep their index order); eigenvectors into G
  TLOOP(i, n) {
    const double li = As[i + n * i].x;
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const double lj = As[j + n * j].x;
      rank += (lj < li) || (lj == li && j < i);
    }
    iwork[i] = rank;     // pp/qq are dead by now (2*half >= n)
    w[rank] = li;
    if (lam_out && rank < n_keep) lam_out[rank] = li;
  }
  TSYNC();
  TLOOP(e, n * n) {
    const int i = e % n, j = e / n;
    G[i + ldg * iwork[j]] = V[i + n * j];
  }
  TLOOP(t, 1) {
    stats[0] = done ? (double)(sweeps > 0 ? sweeps : 1) : 0.0;
    stats[1] = off;
  }
}

}  // namespace dftk
