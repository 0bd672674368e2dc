// LOBPCG block eigensolver on the device.  Restates the algorithm of the reference's
// src/eigen/lobpcg_hyper_impl.jl:354-582 (LOBPCG with B = I), :141-171 (rayleigh_ritz), :216-261
// (ortho!), :271-323 (ortho!(X,Y,BY)), :190-210 (safe_cholesky), :264-268 (drop_small!) and the TPA
// preconditioner of src/eigen/preconditioners.jl:27-78, with Julia's active-block views expressed as
// column offsets.  All N_pw-sized work is GEMMs (blas.cu) or fused elementwise kernels below; the small
// dense factorisations (<= 3M x 3M) use cuSOLVER (heevd / potrf / trtri), as SURVEY.md §7 allows.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <numeric>
#include "structs.cuh"
#include "lobpcg_small.cuh"

namespace dftk {

static const double EPS = DBL_EPSILON;

struct Mat {
  cplx* p;
  int64_t ld, rows, cols;
  Mat cols_from(int64_t c0) const { return Mat{p + ld * c0, ld, rows, cols - c0}; }
  Mat cols_range(int64_t c0, int64_t nc) const { return Mat{p + ld * c0, ld, rows, nc}; }
};

// ------------------------------------------------------------------ small kernels
__global__ void k_residual(const cplx* __restrict__ AX, const cplx* __restrict__ X,
                           const double* __restrict__ lam, cplx* __restrict__ R, int64_t ld,
                           int64_t n_rows, const double* __restrict__ kin, double* __restrict__ norms,
                           double* __restrict__ meankin) {
  // one CTA per column: R = AX - X*lam; norms = ||R||; meankin = <X|kin|X>   (:443-445, precondprep!)
  const int64_t col = blockIdx.x;
  const cplx* ax = AX + ld * col;
  const cplx* x = X + ld * col;
  cplx* r = R + ld * col;
  const double l = lam[col];
  double s = 0.0, mk = 0.0;
  for (int64_t i = threadIdx.x; i < n_rows; i += blockDim.x) {
    cplx a = ax[i], b = x[i];
    cplx v = make_double2(a.x - l * b.x, a.y - l * b.y);
    r[i] = v;
    s += v.x * v.x + v.y * v.y;
    if (kin) mk += kin[i] * (b.x * b.x + b.y * b.y);
  }
  __shared__ double rs[32], rm[32];
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_down_sync(0xffffffffu, s, o);
    mk += __shfl_down_sync(0xffffffffu, mk, o);
  }
  if ((threadIdx.x & 31) == 0) {
    rs[threadIdx.x >> 5] = s;
    rm[threadIdx.x >> 5] = mk;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    int nw = blockDim.x >> 5;
    s = threadIdx.x < nw ? rs[threadIdx.x] : 0.0;
    mk = threadIdx.x < nw ? rm[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_down_sync(0xffffffffu, s, o);
      mk += __shfl_down_sync(0xffffffffu, mk, o);
    }
    if (threadIdx.x == 0) {
      norms[col] = sqrt(s);
      meankin[col] = mk;
    }
  }
}

// R[:,n] *= mk_n / (mk_n + kin)    (ldiv!(::PreconditionerTPA), src/gpu/linalg.jl:29-36)
__global__ void k_precondition(cplx* __restrict__ R, int64_t ld, int64_t n_rows, int64_t n_cols,
                               const double* __restrict__ kin, const double* __restrict__ meankin) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rows * n_cols) return;
  int64_t i = idx % n_rows, c = idx / n_rows;
  double mk = meankin[c];
  double f = mk / (mk + kin[i]);
  cplx v = R[i + ld * c];
  R[i + ld * c] = make_double2(v.x * f, v.y * f);
}

__global__ void k_col_norms(const cplx* __restrict__ X, int64_t ld, int64_t n_rows,
                            double* __restrict__ norms) {
  const int64_t col = blockIdx.x;
  const cplx* x = X + ld * col;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n_rows; i += blockDim.x) {
    cplx v = x[i];
    s += v.x * v.x + v.y * v.y;
  }
  __shared__ double rs[32];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) rs[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    int nw = blockDim.x >> 5;
    s = threadIdx.x < nw ? rs[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) norms[col] = sqrt(s);
  }
}

__global__ void k_scale_cols_inv(cplx* __restrict__ X, int64_t ld, int64_t n_rows, int64_t n_cols,
                                 const double* __restrict__ norms) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rows * n_cols) return;
  int64_t i = idx % n_rows, c = idx / n_rows;
  double f = 1.0 / norms[c];
  cplx v = X[i + ld * c];
  X[i + ld * c] = make_double2(v.x * f, v.y * f);
}

// strided 2D copy (dst and src column-major with different leading dimensions)
__global__ void k_copy2d(cplx* __restrict__ dst, int64_t ldd, const cplx* __restrict__ src, int64_t lds,
                         int64_t n_rows, int64_t n_cols) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rows * n_cols) return;
  int64_t i = idx % n_rows, c = idx / n_rows;
  dst[i + ldd * c] = src[i + lds * c];
}

// Hermitian(upper): mirror the strictly upper triangle into the lower one, make the diagonal real
__global__ void k_hermitize_upper(cplx* __restrict__ A, int64_t ld, int64_t n) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * n) return;
  int64_t i = idx % n, j = idx / n;
  if (i > j) {
    cplx v = A[j + ld * i];
    A[i + ld * j] = make_double2(v.x, -v.y);
  } else if (i == j) {
    A[i + ld * j].y = 0.0;
  }
}
__global__ void k_zero_lower(cplx* __restrict__ A, int64_t ld, int64_t n) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * n) return;
  int64_t i = idx % n, j = idx / n;
  if (i > j) A[i + ld * j] = make_double2(0.0, 0.0);
}
__global__ void k_add_diag(cplx* __restrict__ A, int64_t ld, int64_t n, double shift) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) A[i + ld * i].x += shift;
}
// cP = cX[:, c0:] - e,  e[newly_locked + c, c] = 1 for c < lenXn   (lobpcg_hyper_impl.jl:495-503)
__global__ void k_make_cP(cplx* __restrict__ cP, const cplx* __restrict__ cX, int64_t ld, int64_t n_rows,
                          int64_t n_cols, int64_t c0, int64_t lenXn) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rows * n_cols) return;
  int64_t i = idx % n_rows, c = idx / n_rows;
  int64_t cc = c + c0;  // column in cX / e
  cplx v = cX[i + ld * cc];
  if (cc < lenXn && i == c0 + cc) v.x -= 1.0;
  cP[i + ld * c] = v;
}

// stats[0] = max |diag|, stats[1] = sum |offdiag|^2, stats[2] = #nan/inf, stats[3] = sum |all|^2
__global__ void k_matrix_stats(const cplx* __restrict__ A, int64_t ld, int64_t n_rows, int64_t n_cols,
                               double* __restrict__ stats) {
  // single CTA (matrices are small): deterministic
  double md = 0.0, so = 0.0, bad = 0.0, sa = 0.0;
  for (int64_t idx = threadIdx.x; idx < n_rows * n_cols; idx += blockDim.x) {
    int64_t i = idx % n_rows, j = idx / n_rows;
    cplx v = A[i + ld * j];
    double a2 = v.x * v.x + v.y * v.y;
    if (!isfinite(a2)) bad += 1.0;
    sa += a2;
    if (i == j) md = fmax(md, sqrt(a2));
    else so += a2;
  }
  __shared__ double r0[32], r1[32], r2[32], r3[32];
  for (int o = 16; o > 0; o >>= 1) {
    md = fmax(md, __shfl_down_sync(0xffffffffu, md, o));
    so += __shfl_down_sync(0xffffffffu, so, o);
    bad += __shfl_down_sync(0xffffffffu, bad, o);
    sa += __shfl_down_sync(0xffffffffu, sa, o);
  }
  if ((threadIdx.x & 31) == 0) {
    r0[threadIdx.x >> 5] = md;
    r1[threadIdx.x >> 5] = so;
    r2[threadIdx.x >> 5] = bad;
    r3[threadIdx.x >> 5] = sa;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int nw = blockDim.x >> 5;
    for (int w = 1; w < nw; ++w) {
      md = fmax(md, r0[w]);
      so += r1[w];
      bad += r2[w];
      sa += r3[w];
    }
    stats[0] = md;
    stats[1] = so;
    stats[2] = bad;
    stats[3] = sa;
  }
}

// counter-based normal random numbers (for drop_small!'s re-randomisation; statistically plain)
__device__ __forceinline__ uint64_t splitmix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__global__ void k_randn_col(cplx* __restrict__ x, int64_t n_rows, uint64_t seed) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  uint64_t a = splitmix(seed + 2 * (uint64_t)i), b = splitmix(seed + 2 * (uint64_t)i + 1);
  double u1 = ((a >> 11) + 1.0) * (1.0 / 9007199254740993.0);
  double u2 = (b >> 11) * (1.0 / 9007199254740992.0);
  double r = sqrt(-2.0 * log(u1));
  x[i] = make_double2(r * cospi(2.0 * u2) * 0.70710678118654752, r * sinpi(2.0 * u2) * 0.70710678118654752);
}
__global__ void k_compute_lambda(const cplx* __restrict__ num, const cplx* __restrict__ den,
                                 double* __restrict__ lam, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // real((x'Ax)/(x'x)) with complex division, compute_λ (:341-344)
  cplx a = num[i], b = den[i];
  double d = b.x * b.x + b.y * b.y;
  lam[i] = (a.x * b.x + a.y * b.y) / d;
}

// ------------------------------------------------------------------ fused small-matrix path (lobpcg_small.cuh)
extern __shared__ __align__(16) unsigned char small_dyn_smem[];

// one launch per block-list Gram: CTA partials, the last CTA to finish sums them in a fixed order (deterministic)
__global__ void __launch_bounds__(256)
k_small_gram(SmallMatList A, SmallMatList B, long long rows_per_cta, long long n_rows, int upper_only, cplx* ws, cplx* C,
             long long ldc, unsigned* counter) {
  small_gram_cta((int)blockIdx.x, rows_per_cta, n_rows, A, B, upper_only, ws, (cplx*)small_dyn_smem);
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (is_last) {
    __threadfence();
    small_gram_reduce((int)gridDim.x, A, B, upper_only, ws, C, ldc);
    if (threadIdx.x == 0) *counter = 0;
  }
}

__global__ void __launch_bounds__(128)
k_small_blocks_times(SmallMatList Y, const cplx* cm, long long ldcm, int ncols, cplx* out, long long ldo, long long n_rows,
                     double alpha, double beta) {
  cplx* cs = (cplx*)small_dyn_smem;
  const int ny = Y.start[Y.n];
  for (int e = threadIdx.x; e < ny * ncols; e += blockDim.x) cs[e] = cm[e % ny + ldcm * (e / ny)];
  __syncthreads();
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rows) small_blocks_times_row(r, Y, cs, ny, ncols, out, ldo, alpha, beta);
}

__global__ void __launch_bounds__(128)
k_small_rmul(cplx* X, long long ld, long long n_rows, int n, const cplx* invR, long long ldr) {
  cplx* rs = (cplx*)small_dyn_smem;
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) rs[e] = invR[e % n + ldr * (e / n)];
  __syncthreads();
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rows) small_rmul_row(r, X, ld, n, rs, n);
}

__global__ void __launch_bounds__(SMALL_RED)
k_small_chol(const cplx* O, long long ldo, int n, cplx* invR, long long ldi, double* stats) {
  __shared__ cplx As[SMALL_MAX_N * SMALL_MAX_N], Bs[SMALL_MAX_N * SMALL_MAX_N];
  __shared__ double red[SMALL_RED];
  __shared__ int flag[2];
  small_chol_cta(O, ldo, n, invR, ldi, stats, As, Bs, red, flag);
}

static inline unsigned nblk(int64_t n) { return (unsigned)((n + 255) / 256); }

// Optional section timing (env DFTK_B200_PROFILE=1): stream-synchronising wall clock per LOBPCG section.
struct SectionProf {
  bool on;
  cudaStream_t st;
  std::vector<std::pair<std::string, double>> acc;
  double t0 = 0;
  std::string cur;
  static double now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
  }
  void begin(const char* name) {
    if (!on) return;
    end();
    cudaStreamSynchronize(st);
    cur = name;
    t0 = now();
  }
  void end() {
    if (!on || cur.empty()) return;
    cudaStreamSynchronize(st);
    double dt = now() - t0;
    for (auto& a : acc)
      if (a.first == cur) {
        a.second += dt;
        cur.clear();
        return;
      }
    acc.push_back({cur, dt});
    cur.clear();
  }
  // DFTK_B200_PROFILE=2: accumulate over all solves of the process, print once at exit (small systems: hundreds of solves)
  struct Global {
    std::vector<std::pair<std::string, double>> acc;
    long iters = 0, solves = 0;
    ~Global() {
      if (acc.empty()) return;
      double tot = 0;
      for (auto& a : acc) tot += a.second;
      fprintf(stderr, "[dftk_b200 lobpcg profile, all solves] %ld solves, %ld iterations, %.3f s in sections\n", solves, iters, tot);
      for (auto& a : acc) fprintf(stderr, "  %-22s %9.3f s  %5.1f %%\n", a.first.c_str(), a.second, 100 * a.second / tot);
    }
  };
  static Global& global() {
    static Global g;
    return g;
  }
  bool aggregate = false;
  void report(int niter) {
    if (!on) return;
    end();
    if (aggregate) {
      Global& g = global();
      g.iters += niter;
      g.solves++;
      for (auto& a : acc) {
        bool found = false;
        for (auto& b : g.acc)
          if (b.first == a.first) {
            b.second += a.second;
            found = true;
          }
        if (!found) g.acc.push_back(a);
      }
      return;
    }
    double tot = 0;
    for (auto& a : acc) tot += a.second;
    fprintf(stderr, "[dftk_b200 lobpcg profile] %d iterations, %.3f s in sections\n", niter, tot);
    for (auto& a : acc) fprintf(stderr, "  %-22s %9.3f s  %5.1f %%\n", a.first.c_str(), a.second, 100 * a.second / tot);
  }
};

// ------------------------------------------------------------------ solver object
struct Lobpcg {
  dftk_b200_kblock* kb;
  dftk_b200_ctx* ctx;
  int64_t N, M;
  bool use_prec;
  uint64_t rng_counter = 0x5EEDull;
  // device scalars / small vectors
  double *d_lam, *d_norms, *d_meankin, *d_stats, *d_w;
  cplx* d_cdots;
  // small dense scratch (leading dimension S3 = 3M)
  int64_t S3;
  cplx *G, *cX, *cP, *Ochol, *invR, *BYX, *tmpS;
  // big scratch
  cplx* tmpN;  // N x M
  // fused small-matrix path (M <= SMALL_MAX_N): see lobpcg_small.cuh
  bool small = false;
  unsigned* d_counter = nullptr;

  static SmallMatList mklist(const std::vector<Mat>& v) {
    SmallMatList L{};
    REQUIRE(v.size() >= 1 && v.size() <= 3, "small path: block list too long");
    L.n = (int)v.size();
    int off = 0;
    for (int i = 0; i < 3; ++i) {
      L.start[i] = off;
      if (i < L.n) {
        L.p[i] = v[i].p;
        L.ld[i] = v[i].ld;
        L.cols[i] = (int)v[i].cols;
        off += (int)v[i].cols;
      } else {
        L.p[i] = nullptr;
        L.ld[i] = 0;
        L.cols[i] = 0;
      }
    }
    L.start[3] = off;
    for (int i = L.n; i < 4; ++i) L.start[i] = off;
    REQUIRE(off <= SMALL_MAX_COLS, "small path: too many columns");
    return L;
  }
  void small_gram(const std::vector<Mat>& A, const std::vector<Mat>& B, cplx* C, int64_t ldc, bool upper_only) {
    SmallMatList LA = mklist(A), LB = mklist(B);
    const int nA = LA.start[LA.n], nB = LB.start[LB.n];
    if (nA == 0 || nB == 0) return;
    const int64_t rows = A[0].rows;
    int64_t n_ctas = std::max<int64_t>(1, std::min<int64_t>((rows + 127) / 128, 2 * (int64_t)ctx->sm_count));
    int64_t rpc = (rows + n_ctas - 1) / n_ctas;
    rpc = (rpc + SMALL_TR - 1) / SMALL_TR * SMALL_TR;
    n_ctas = std::max<int64_t>(1, (rows + rpc - 1) / rpc);
    cplx* ws = (cplx*)ctx->gemm_ws.ensure((size_t)n_ctas * nA * nB * sizeof(cplx));
    const size_t smem = (size_t)SMALL_TR * (nA + nB) * sizeof(cplx);
    LAUNCH(ctx, k_small_gram, (unsigned)n_ctas, 256, smem, LA, LB, (long long)rpc, (long long)rows, upper_only ? 1 : 0, ws,
           C, (long long)ldc, d_counter);
  }
  void small_blocks_times(const std::vector<Mat>& Y, const cplx* c, int64_t ldc, int64_t ncols, Mat out, double alpha,
                          double beta) {
    if (ncols == 0 || out.rows == 0) return;
    REQUIRE(ncols <= SMALL_MAX_N, "small path: too many output columns");
    SmallMatList LY = mklist(Y);
    const int ny = LY.start[LY.n];
    LAUNCH(ctx, k_small_blocks_times, (unsigned)((out.rows + 127) / 128), 128, (size_t)ny * ncols * sizeof(cplx), LY, c,
           (long long)ldc, (int)ncols, out.p, (long long)out.ld, (long long)out.rows, alpha, beta);
  }

  void copy2d(Mat dst, Mat src) {
    if (src.rows == 0 || src.cols == 0) return;
    LAUNCH(ctx, k_copy2d, nblk(src.rows * src.cols), 256, 0, dst.p, dst.ld, (const cplx*)src.p, src.ld,
           src.rows, src.cols);
  }
  void get(void* host, const void* dev, size_t bytes) {
    CUDA_CHECK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  }
  void stats(const cplx* A, int64_t ld, int64_t r, int64_t c, double* out4) {
    LAUNCH(ctx, k_matrix_stats, 1, 1024, 0, A, ld, r, c, d_stats);
    get(out4, d_stats, 4 * sizeof(double));
  }
  double normest(const cplx* A, int64_t ld, int64_t n) {
    double s[4];
    stats(A, ld, n, n, s);
    return s[0] + std::sqrt(s[1]);
  }

  // C = op(A)' * B accumulated over block lists (LazyHcat products, :90-137)
  void gram(const std::vector<Mat>& A, const std::vector<Mat>& B, cplx* C, int64_t ldc, bool upper_only) {
    if (small) return small_gram(A, B, C, ldc, upper_only);
    const cplx one = make_double2(1, 0), zero = make_double2(0, 0);
    int64_t oc = 0;
    for (size_t ib = 0; ib < B.size(); ++ib) {
      int64_t orow = 0;
      for (size_t ia = 0; ia < A.size(); ++ia) {
        if (!(upper_only && ib < ia))
          zgemm(ctx, 2, A[ia].cols, B[ib].cols, A[ia].rows, one, A[ia].p, A[ia].ld, B[ib].p, B[ib].ld,
                zero, C + orow + ldc * oc, ldc, /*upper tiles only on diagonal blocks*/ upper_only && ia == ib);
        orow += A[ia].cols;
      }
      oc += B[ib].cols;
    }
  }
  // out (=|+=) alpha * [Y blocks] * c     (mul!(res, ::LazyHcat, B, α, β), :124-132)
  void blocks_times(const std::vector<Mat>& Y, const cplx* c, int64_t ldc, int64_t ncols, Mat out,
                    double alpha, double beta) {
    if (small) return small_blocks_times(Y, c, ldc, ncols, out, alpha, beta);
    int64_t off = 0;
    for (size_t i = 0; i < Y.size(); ++i) {
      zgemm(ctx, 0, Y[i].rows, ncols, Y[i].cols, make_double2(alpha, 0), Y[i].p, Y[i].ld, c + off, ldc,
            make_double2(i == 0 ? beta : 1.0, 0), out.p, out.ld);
      off += Y[i].cols;
    }
  }

  // ortho!(X) :216-261.  X: rows x n (in place).  `tmp` must hold rows x n.
  // returns growth factor; throws on the (never observed) SVD-fallback condition
  double ortho(Mat X, cplx* tmp, int64_t ldtmp) {
    const int64_t n = X.cols;
    if (n == 0) return 1.0;
    double growth = 1.0;
    if (small) {
      // gram -> fused safe_cholesky/inverse/normest (one CTA) -> one host sync -> X *= invR
      for (int outer = 0; outer < 50; ++outer) {
        gram({X}, {X}, Ochol, S3, true);
        LAUNCH(ctx, k_small_chol, 1, SMALL_RED, 0, (const cplx*)Ochol, (long long)S3, (int)n, invR, (long long)S3, d_stats);
        double s[4];
        get(s, d_stats, 4 * sizeof(double));
        const int nchol = (int)s[0];
        if (nchol == 0) throw Error(DFTK_B200_ENUM, "ortho!: Cholesky failing badly (SVD fallback not implemented)");
        LAUNCH(ctx, k_small_rmul, (unsigned)((X.rows + 127) / 128), 128, (size_t)n * n * sizeof(cplx), X.p, (long long)X.ld,
               (long long)X.rows, (int)n, (const cplx*)invR, (long long)S3);
        const double norminvR = s[1];
        growth *= norminvR;
        const double condR = s[2] * norminvR;
        const double est = EPS * condR * condR;
        n_chol_total += nchol;
        if (nchol == 1 && est < 2 * EPS) break;
        if (outer == 49) throw Error(DFTK_B200_ENUM, "ortho!: did not converge");
      }
      return growth;
    }
    for (int outer = 0; outer < 50; ++outer) {
      gram({X}, {X}, Ochol, S3, true);
      LAUNCH(ctx, k_hermitize_upper, nblk(n * n), 256, 0, Ochol, S3, n);
      // safe_cholesky :190-210
      int nchol = 0;
      double alpha = 100.0;
      bool ok = false;
      double onorm = -1.0;
      while (nchol < 5) {
        nchol++;
        copy2d(Mat{invR, S3, n, n}, Mat{Ochol, S3, n, n});  // factor a copy (invR doubles as R storage)
        int info = potrf_upper(invR, n);
        if (info == 0) {
          // R = upper factor; keep R in tmpS, invert in invR
          LAUNCH(ctx, k_zero_lower, nblk(n * n), 256, 0, invR, S3, n);
          copy2d(Mat{tmpS, S3, n, n}, Mat{invR, S3, n, n});
          int info2 = trtri_upper(invR, n);
          double s[4];
          stats(invR, S3, n, n, s);
          if (info2 == 0 && s[2] == 0.0) {
            ok = true;
            break;
          }
        }
        if (onorm < 0) {
          double s[4];
          stats(Ochol, S3, n, n, s);
          onorm = std::sqrt(s[3]);
        }
        LAUNCH(ctx, k_add_diag, nblk(n), 256, 0, Ochol, S3, n, alpha * EPS * onorm);
        // note: the reference recomputes norm(O) of the shifted matrix; the difference is O(eps)
        alpha *= 10;
      }
      if (!ok) throw Error(DFTK_B200_ENUM, "ortho!: Cholesky failing badly (SVD fallback not implemented)");
      // X <- X * invR   (rmul!(X, invR))
      zgemm(ctx, 0, X.rows, n, n, make_double2(1, 0), X.p, X.ld, invR, S3, make_double2(0, 0), tmp, ldtmp,
            /*invR is upper triangular*/ true);
      copy2d(X, Mat{tmp, ldtmp, X.rows, n});
      double norminvR = normest(invR, S3, n);
      growth *= norminvR;
      double condR = normest(tmpS, S3, n) * norminvR;
      double est = EPS * condR * condR;
      n_chol_total += nchol;
      if (nchol == 1 && est < 2 * EPS) break;
      if (outer == 49) throw Error(DFTK_B200_ENUM, "ortho!: did not converge");
    }
    return growth;
  }
  int64_t n_chol_total = 0;

  int potrf_upper(cplx* A, int64_t n) {
    int lwork = 0;
    CUSOLVER_CHECK(cusolverDnZpotrf_bufferSize(ctx->cusolver, CUBLAS_FILL_MODE_UPPER, (int)n,
                                               (cuDoubleComplex*)A, (int)S3, &lwork));
    char* w = ctx->solver_work.ensure((size_t)lwork * sizeof(cuDoubleComplex) + 16);
    int* dinfo = ctx->dev_info.ensure(4);
    CUSOLVER_CHECK(cusolverDnZpotrf(ctx->cusolver, CUBLAS_FILL_MODE_UPPER, (int)n, (cuDoubleComplex*)A,
                                    (int)S3, (cuDoubleComplex*)w, lwork, dinfo));
    ctx->launches++;
    int info = 0;
    get(&info, dinfo, sizeof(int));
    return info;
  }
  int trtri_upper(cplx* A, int64_t n) {
    size_t wd = 0, wh = 0;
    CUSOLVER_CHECK(cusolverDnXtrtri_bufferSize(ctx->cusolver, CUBLAS_FILL_MODE_UPPER, CUBLAS_DIAG_NON_UNIT,
                                               n, CUDA_C_64F, A, S3, &wd, &wh));
    char* w = ctx->solver_work.ensure(wd + 16);
    std::vector<char> hw(wh + 16);
    int* dinfo = ctx->dev_info.ensure(4);
    CUSOLVER_CHECK(cusolverDnXtrtri(ctx->cusolver, CUBLAS_FILL_MODE_UPPER, CUBLAS_DIAG_NON_UNIT, n,
                                    CUDA_C_64F, A, S3, w, wd, hw.data(), wh, dinfo));
    ctx->launches++;
    int info = 0;
    get(&info, dinfo, sizeof(int));
    return info;
  }
  // eigen(Hermitian(G)) upper triangle; eigenvectors overwrite G, eigenvalues -> d_w (ascending)
  void heevd(cplx* A, int64_t n) {
    // 64-bit generic API: unlike the legacy cusolverDnZheevd it has no OpenMP host stage whose speed depends on the
    // process' OMP_* environment (measured: 26 ms for n = 1509 under every setting vs 30-700 ms for Zheevd)
    if (!ctx->solver_params) CUSOLVER_CHECK(cusolverDnCreateParams(&ctx->solver_params));
    size_t wd = 0, wh = 0;
    CUSOLVER_CHECK(cusolverDnXsyevd_bufferSize(ctx->cusolver, ctx->solver_params, CUSOLVER_EIG_MODE_VECTOR,
                                               CUBLAS_FILL_MODE_UPPER, n, CUDA_C_64F, A, S3, CUDA_R_64F, d_w, CUDA_C_64F,
                                               &wd, &wh));
    char* w = ctx->solver_work.ensure(wd + 16);
    if (ctx->solver_host_work.size() < wh + 16) ctx->solver_host_work.resize(wh + 16);
    int* dinfo = ctx->dev_info.ensure(4);
    CUSOLVER_CHECK(cusolverDnXsyevd(ctx->cusolver, ctx->solver_params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n,
                                    CUDA_C_64F, A, S3, CUDA_R_64F, d_w, CUDA_C_64F, w, wd, ctx->solver_host_work.data(), wh,
                                    dinfo));
    ctx->launches++;
    int info = 0;
    get(&info, dinfo, sizeof(int));
    if (info != 0) throw Error(DFTK_B200_ENUM, "rayleigh_ritz: heevd failed, info=" + std::to_string(info));
  }

  // ortho!(X, Y, BY=Y) :271-323.  X: rows x n in place; Y: block list with the same row count.
  void ortho_against(Mat X, const std::vector<Mat>& Y, cplx* tmp, int64_t ldtmp) {
    const int64_t n = X.cols;
    if (n == 0) return;
    const double tol = 2 * EPS;
    int64_t ny = 0;
    for (auto& y : Y) ny += y.cols;
    REQUIRE(ny <= ldBYX, "ortho_against: workspace too small");
    LAUNCH(ctx, k_col_norms, (unsigned)n, 256, 0, (const cplx*)X.p, X.ld, X.rows, d_norms);
    LAUNCH(ctx, k_scale_cols_inv, nblk(X.rows * n), 256, 0, X.p, X.ld, X.rows, n, (const double*)d_norms);
    std::vector<double> norms(n);
    for (int niter = 1;; ++niter) {
      gram(Y, {X}, BYX, ldBYX, false);
      blocks_times(Y, BYX, ldBYX, n, X, -1.0, 1.0);  // X -= Y * BY'X
      // drop_small! :264-268
      LAUNCH(ctx, k_col_norms, (unsigned)n, 256, 0, (const cplx*)X.p, X.ld, X.rows, d_norms);
      // ||BY'X|| is needed below; it does not depend on the re-randomisation, so both results share one host sync
      LAUNCH(ctx, k_matrix_stats, 1, 1024, 0, (const cplx*)BYX, ldBYX, ny, n, d_stats);
      const size_t span = (size_t)(d_stats - d_norms) + 4;
      std::vector<double> both(span);
      get(both.data(), d_norms, span * sizeof(double));
      std::copy(both.begin(), both.begin() + n, norms.begin());
      const double* s = both.data() + (d_stats - d_norms);
      for (int64_t c = 0; c < n; ++c) {
        if (norms[c] <= tol) {
          Mat xc = X.cols_range(c, 1);
          rng_counter += 0x100000000ull;
          LAUNCH(ctx, k_randn_col, nblk(X.rows), 256, 0, xc.p, X.rows, rng_counter);
          // X[:,c] -= Y (BY' X[:,c])
          gram(Y, {xc}, tmpS, S3, false);
          blocks_times(Y, tmpS, S3, 1, xc, -1.0, 1.0);
        }
      }
      if (std::sqrt(s[3]) < tol && niter > 1) break;
      double growth = ortho(X, tmp, ldtmp);
      if (growth * EPS < tol) break;
      if (niter > 10) throw Error(DFTK_B200_ENUM, "ortho!(X,Y): failing badly (SVD fallback not implemented)");
    }
  }
  int64_t ldBYX = 0;
};

int lobpcg_run(dftk_b200_kblock* kb, cplx* Xio, int64_t M, double tol, int miniter, int maxiter,
               int64_t n_conv_check, bool use_prec, double* lambda_host, double* resid_host, int* n_iter_out,
               int64_t* n_matvec_out, int* converged_out) {
  dftk_b200_ctx* ctx = kb->grid->ctx;
  const int64_t N = kb->n_pw;
  REQUIRE(M >= 1, "lobpcg: n_bands must be >= 1");
  REQUIRE(N > 3 * M, "The eigenproblem is too small, and the iterative eigensolver will fail; increase "
                     "the number of degrees of freedom, or use a dense eigensolver.");
  if (n_conv_check <= 0 || n_conv_check > M) n_conv_check = M;
  use_prec = use_prec && kb->has_kin;

  SectionProf prof;
  prof.on = getenv("DFTK_B200_PROFILE") != nullptr;
  prof.aggregate = prof.on && atoi(getenv("DFTK_B200_PROFILE")) == 2;
  prof.st = ctx->stream;
  Lobpcg L;
  L.kb = kb;
  L.ctx = ctx;
  L.N = N;
  L.M = M;
  L.use_prec = use_prec;
  const int64_t S3 = 3 * M;
  L.S3 = S3;
  L.ldBYX = 2 * M > S3 ? 2 * M : S3;
  // workspaces
  cplx* big = kb->lobpcg_ws.ensure((size_t)11 * N * M);
  cplx *AX = big, *R = big + N * M, *AR = big + 2 * N * M, *P = big + 3 * N * M, *AP = big + 4 * N * M,
       *nX = big + 5 * N * M, *nAX = big + 6 * N * M, *nR = big + 7 * N * M, *nP = big + 8 * N * M,
       *nAP = big + 9 * N * M;
  L.tmpN = big + 10 * N * M;
  size_t small_elems = (size_t)S3 * S3 * 4 + (size_t)S3 * M * 2 + (size_t)L.ldBYX * M + 4 * M + 64;
  cplx* sm = kb->small_ws.ensure(small_elems);
  L.G = sm;
  L.Ochol = sm + S3 * S3;
  L.invR = sm + 2 * S3 * S3;
  L.tmpS = sm + 3 * S3 * S3;
  L.cX = sm + 4 * S3 * S3;
  L.cP = L.cX + S3 * M;
  L.BYX = L.cP + S3 * M;
  L.d_cdots = L.BYX + L.ldBYX * M;
  double* dsc = ctx->scal.ensure(4 * M + 3 * S3 + 64);
  L.d_lam = dsc;
  L.d_norms = dsc + M;
  L.d_meankin = dsc + 2 * M;
  L.d_w = dsc + 3 * M;
  L.d_stats = dsc + 3 * M + 3 * S3;
  L.small = M <= SMALL_MAX_N && ctx->small_dense != 0;
  if (L.small) {
    unsigned* c = (unsigned*)ctx->small_counter.ensure(4);
    CUDA_CHECK(cudaMemsetAsync(c, 0, 4 * sizeof(int), ctx->stream));   // also recovers from an aborted earlier solve
    L.d_counter = c;
  }

  Mat X{Xio, N, N, M};
  auto mat = [&](cplx* p) { return Mat{p, N, N, M}; };
  auto applyH = [&](Mat in, Mat out) {
    // A*X: full H apply on a block of columns (mul!(AX, A, X), :379,416)
    kb_apply_local_kinetic(kb, in.p, out.p, in.cols, kb->has_V, kb->has_kin, false);
    kb_apply_nonlocal(kb, in.p, out.p, in.cols);
  };
  const size_t colbytes = (size_t)N * sizeof(cplx);
  auto copycols = [&](cplx* dst, const cplx* src, int64_t c0, int64_t nc) {
    if (nc > 0)
      CUDA_CHECK(cudaMemcpyAsync(dst + N * c0, src + N * c0, colbytes * nc, cudaMemcpyDeviceToDevice, ctx->stream));
  };

  std::vector<double> resid_hist((size_t)M * (maxiter + 1), 0.0);
  auto RH = [&](int64_t i, int it) -> double& { return resid_hist[(size_t)it * M + i]; };

  // X = ortho!(copy(X)) :370
  prof.begin("ortho(X0)");
  L.ortho(X, L.tmpN, N);
  prof.begin("H*X");
  int64_t n_matvec = M;
  applyH(X, mat(AX));
  prof.begin("misc");
  CUDA_CHECK(cudaMemsetAsync(big + N * M, 0, (size_t)4 * N * M * sizeof(cplx), ctx->stream));   // R, AR, P, AP
  CUDA_CHECK(cudaMemsetAsync(nR, 0, (size_t)3 * N * M * sizeof(cplx), ctx->stream));           // nR, nP, nAP
  copycols(nX, Xio, 0, M);
  copycols(nAX, AX, 0, M);
  // λ = compute_λ(X, AX, X)
  columnwise_dots(ctx, Xio, N, AX, N, N, M, L.d_cdots);
  columnwise_dots(ctx, Xio, N, Xio, N, N, M, L.d_cdots + M);
  LAUNCH(ctx, k_compute_lambda, nblk(M), 256, 0, (const cplx*)L.d_cdots, (const cplx*)(L.d_cdots + M),
         L.d_lam, M);

  int64_t nlocked = 0, a0 = 0;
  int niter = 0;
  std::vector<double> norms(M), lam_h(M);
  int64_t ncolsY = 0;
  bool done = false;
  int final_iter = maxiter;
  while (true) {
    const int64_t Ma = M - a0;
    std::vector<Mat> Y, AY;
    if (niter > 0) {
      prof.begin("H*R");
      applyH(mat(R).cols_from(a0), mat(AR).cols_from(a0));
      n_matvec += Ma;
      Y = {X.cols_from(a0), mat(R).cols_from(a0)};
      AY = {mat(AX).cols_from(a0), mat(AR).cols_from(a0)};
      if (niter > 1) {
        Y.push_back(mat(P).cols_from(a0));
        AY.push_back(mat(AP).cols_from(a0));
      }
      ncolsY = (int64_t)Y.size() * Ma;
      // rayleigh_ritz :141-171
      prof.begin("RR gram Y'AY");
      L.gram(Y, AY, L.G, S3, true);
      prof.begin("RR heevd");
      // only the block upper triangle of G is written; heevd reads the upper triangle only
      L.heevd(L.G, ncolsY);
      prof.begin("X,AX = Y cX");
      // cX = vectors[:, 1:Ma], λ = values[1:Ma]
      L.copy2d(Mat{L.cX, S3, ncolsY, Ma}, Mat{L.G, S3, ncolsY, Ma});
      CUDA_CHECK(cudaMemcpyAsync(L.d_lam + a0, L.d_w, Ma * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
      L.blocks_times(Y, L.cX, S3, Ma, mat(nX).cols_from(a0), 1.0, 0.0);
      L.blocks_times(AY, L.cX, S3, Ma, mat(nAX).cols_from(a0), 1.0, 0.0);
    }
    prof.begin("residual+precond");
    // residuals :443-445 (+ precondprep! :452-457 fused)
    LAUNCH(ctx, k_residual, (unsigned)Ma, 256, 0, (const cplx*)(nAX + N * a0), (const cplx*)(nX + N * a0),
           (const double*)(L.d_lam + a0), nR + N * a0, N, N, use_prec ? (const double*)kb->kin.p : nullptr,
           L.d_norms, L.d_meankin);
    L.get(norms.data(), L.d_norms, Ma * sizeof(double));
    for (int64_t i = 0; i < Ma; ++i) RH(a0 + i, niter) = norms[i];
    if (use_prec)
      LAUNCH(ctx, k_precondition, nblk(N * Ma), 256, 0, nR + N * a0, N, N, Ma, (const double*)kb->kin.p,
             (const double*)L.d_meankin);

    const int64_t prev_nlocked = nlocked;
    if (niter >= miniter) {
      for (int64_t i = nlocked; i < M; ++i) {
        if (RH(i, niter) < tol) nlocked++;
        else break;
      }
    }
    if (nlocked >= n_conv_check) {
      copycols(Xio, nX, a0, Ma);
      copycols(AX, nAX, a0, Ma);
      final_iter = niter;
      done = true;
      break;
    }
    const int64_t newly = nlocked - prev_nlocked;

    if (niter > 0) {
      const int64_t lenXn = Ma - newly;
      prof.begin("cP ortho");
      // cP = (cX - e)[:, newly:Ma]; ortho!(cP, cX, cX)
      LAUNCH(ctx, k_make_cP, nblk(ncolsY * lenXn), 256, 0, L.cP, (const cplx*)L.cX, S3, ncolsY, lenXn, newly,
             lenXn);
      L.ortho_against(Mat{L.cP, S3, ncolsY, lenXn}, {Mat{L.cX, S3, ncolsY, Ma}}, L.G, S3);
      prof.begin("P,AP = Y cP");
      L.blocks_times(Y, L.cP, S3, lenXn, mat(nP).cols_from(a0 + newly), 1.0, 0.0);
      L.blocks_times(AY, L.cP, S3, lenXn, mat(nAP).cols_from(a0 + newly), 1.0, 0.0);
    }
    prof.begin("copies+check");
    copycols(Xio, nX, a0, Ma);
    copycols(AX, nAX, a0, Ma);
    copycols(R, nR, a0, Ma);
    // sanity check :531-535
    {
      LAUNCH(ctx, k_col_norms, (unsigned)Ma, 256, 0, (const cplx*)(Xio + N * a0), N, N, L.d_norms);
      L.get(norms.data(), L.d_norms, Ma * sizeof(double));
      for (int64_t i = 0; i < Ma; ++i)
        if (!(std::fabs(norms[i] * norms[i] - 1.0) < std::sqrt(EPS)))
          throw Error(DFTK_B200_ENUM, "LOBPCG is badly failing to keep the vectors normalized; this should never happen");
    }
    a0 += newly;
    std::vector<Mat> Z = {X};
    if (niter > 0) {
      copycols(P, nP, a0, M - a0);
      copycols(AP, nAP, a0, M - a0);
      Z.push_back(mat(P).cols_from(a0));
    }
    prof.begin("ortho R vs (X,P)");
    L.ortho_against(mat(R).cols_from(a0), Z, L.tmpN, N);
    prof.end();

    if (niter >= maxiter) break;
    niter++;
  }
  (void)done;
  prof.report(niter);
  // final_retval :325-338
  L.get(lam_h.data(), L.d_lam, M * sizeof(double));
  std::vector<int64_t> perm(M);
  std::iota(perm.begin(), perm.end(), 0);
  bool sorted = std::is_sorted(lam_h.begin(), lam_h.end());
  if (!sorted) {
    std::stable_sort(perm.begin(), perm.end(), [&](int64_t a, int64_t b) { return lam_h[a] < lam_h[b]; });
    // permute X columns through tmpN
    for (int64_t c = 0; c < M; ++c)
      CUDA_CHECK(cudaMemcpyAsync(L.tmpN + N * c, Xio + N * perm[c], colbytes, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(Xio, L.tmpN, colbytes * M, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  double maxres = 0.0;
  for (int64_t c = 0; c < M; ++c) {
    lambda_host[c] = lam_h[perm[c]];
    resid_host[c] = RH(perm[c], final_iter);
    if (c < n_conv_check) maxres = std::max(maxres, resid_host[c]);
  }
  *n_iter_out = final_iter;
  *n_matvec_out = n_matvec;
  *converged_out = maxres < tol ? 1 : 0;
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // namespace dftk
