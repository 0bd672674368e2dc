// Reference device pipeline of the INT8-emulated FP64 complex GEMM (i8emu_core.cuh): column scales -> int8 residue planes
// -> per-modulus integer dot products -> CRT.  Reached only through option gemm_backend = 2 (C = A^H B, alpha = 1,
// beta = 0).  The integer products here are plain CUDA-core loops: this file is the checker and the plumbing into which a
// `tcgen05.mma.kind::i8` kernel drops (scripts/tcgen05_i8_probe.cu is the hardware bring-up probe); it is groundwork, not
// a measured path, and has not run on hardware yet.
#include "structs.cuh"
#include "i8emu_core.cuh"

namespace dftk {

// e[col] = scale exponent of column col (largest |re|, |im| over the rows); one CTA per column
__global__ void __launch_bounds__(256)
k_i8_col_exponent(const cplx* __restrict__ X, int64_t ld, int64_t n_rows, int bits, int* __restrict__ e) {
  const int64_t col = blockIdx.x;
  double mx = 0.0;
  for (int64_t r = threadIdx.x; r < n_rows; r += blockDim.x) {
    const cplx v = X[r + ld * col];
    mx = fmax(mx, fmax(fabs(v.x), fabs(v.y)));
  }
  __shared__ double red[8];
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) mx = fmax(mx, red[w]);
    e[col] = i8_scale_exponent(mx, bits);
  }
}

// planes[(2 t + part)][col][row] (row fastest: K-major per column), plane stride = n_cols * n_rows
__global__ void __launch_bounds__(256)
k_i8_residues(const cplx* __restrict__ X, int64_t ld, int64_t n_rows, int64_t n_cols, const int* __restrict__ e, int n_mod,
              signed char* __restrict__ planes) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t col = blockIdx.y;
  if (r >= n_rows) return;
  i8_residues_entry(X[r + ld * col], e[col], n_mod, planes + (r + n_rows * col), (long long)n_rows * n_cols);
}

// same with a padded leading dimension (tensor-core path: rows 16-byte aligned, K padded with zeros to the stage size)
__global__ void __launch_bounds__(256)
k_i8_residues_ld(const cplx* __restrict__ X, int64_t ld, int64_t n_rows, int64_t n_cols, int64_t ldk, const int* __restrict__ e,
                 int n_mod, signed char* __restrict__ planes) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t col = blockIdx.y;
  if (r >= n_rows) return;
  i8_residues_entry(X[r + ld * col], e[col], n_mod, planes + (r + ldk * col), (long long)ldk * n_cols);
}

// one warp per (i, j, t): residues of conj(a_i) . b_j modulo p_t;  res[(2 t + part)][j][i]
__global__ void __launch_bounds__(256)
k_i8_dot_ref(const signed char* __restrict__ ra, const signed char* __restrict__ rb, int64_t m, int64_t n, int64_t k,
             int n_mod, int* __restrict__ res) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= m * n * n_mod) return;
  const int64_t i = w % m, j = (w / m) % n;
  const int t = (int)(w / (m * n));
  const int p = i8_modulus(t);
  const signed char* ar = ra + (size_t)(2 * t) * m * k + k * i;
  const signed char* ai = ra + (size_t)(2 * t + 1) * m * k + k * i;
  const signed char* br = rb + (size_t)(2 * t) * n * k + k * j;
  const signed char* bi = rb + (size_t)(2 * t + 1) * n * k + k * j;
  int sre = 0, sim = 0;
  for (int64_t k0 = 0; k0 < k; k0 += 32 * (int64_t)I8_K_CHUNK) {       // per lane at most I8_K_CHUNK terms per int32 sum
    const int64_t k1 = k0 + 32 * (int64_t)I8_K_CHUNK < k ? k0 + 32 * (int64_t)I8_K_CHUNK : k;
    int x1 = 0, x2 = 0, x3 = 0, x4 = 0;
    for (int64_t q = k0 + lane; q < k1; q += 32) {
      const int a0 = ar[q], a1 = ai[q], b0 = br[q], b1 = bi[q];
      x1 += a0 * b0;
      x2 += a1 * b1;
      x3 += a0 * b1;
      x4 += a1 * b0;
    }
    sre = (sre + x1 % p + x2 % p) % p;
    sim = (sim + x3 % p - x4 % p) % p;
  }
  for (int o = 16; o > 0; o >>= 1) {                                     // |partial| < p: the warp sum stays below 32 p
    sre += __shfl_down_sync(0xffffffffu, sre, o);
    sim += __shfl_down_sync(0xffffffffu, sim, o);
  }
  if (lane == 0) {
    res[((size_t)(2 * t) * n + j) * m + i] = i8_sym(sre, p);
    res[((size_t)(2 * t + 1) * n + j) * m + i] = i8_sym(sim, p);
  }
}

__global__ void __launch_bounds__(128)
k_i8_crt(const int* __restrict__ res, int64_t m, int64_t n, I8Tables T, const int* __restrict__ ea, const int* __restrict__ eb,
         cplx* __restrict__ C, int64_t ldc, int accumulate) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m * n) return;
  const int64_t i = idx % m, j = idx / m;
  int rre[I8_MAX_MODULI], rim[I8_MAX_MODULI];
  for (int t = 0; t < T.n_mod; ++t) {
    rre[t] = res[((size_t)(2 * t) * n + j) * m + i];
    rim[t] = res[((size_t)(2 * t + 1) * n + j) * m + i];
  }
  const int sh = -(ea[i] + eb[j]);
  cplx v = make_double2(ldexp(i8_crt(rre, T), sh), ldexp(i8_crt(rim, T), sh));
  if (accumulate) {
    const cplx o = C[i + ldc * j];
    v.x += o.x;
    v.y += o.y;
  }
  C[i + ldc * j] = v;
}

// ---- update-type products C (m x n) (+)= A B, A: m x k (rows scaled), B: k x n (columns scaled) ----
// e[row] = scale exponent of row `row` of A (largest |re|, |im| over the k columns)
__global__ void __launch_bounds__(256)
k_i8_row_exponent(const cplx* __restrict__ A, int64_t lda, int64_t m, int64_t k, int bits, int* __restrict__ e) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  double mx = 0.0;
  for (int64_t c = 0; c < k; ++c) {
    const cplx v = A[i + lda * c];
    mx = fmax(mx, fmax(fabs(v.x), fabs(v.y)));
  }
  e[i] = i8_scale_exponent(mx, bits);
}
// planes[(2 t + part)][c][i] (i fastest, like A itself), plane stride = m * k
__global__ void __launch_bounds__(256)
k_i8_residues_rows(const cplx* __restrict__ A, int64_t lda, int64_t m, int64_t k, const int* __restrict__ e, int n_mod,
                   signed char* __restrict__ planes) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t c = blockIdx.y;
  if (i >= m) return;
  i8_residues_entry(A[i + lda * c], e[i], n_mod, planes + (i + m * c), (long long)m * k);
}
// one thread per (i, j, t);  res[(2 t + part)][j][i]
__global__ void __launch_bounds__(256)
k_i8_dot_plain_ref(const signed char* __restrict__ ra, const signed char* __restrict__ rb, int64_t m, int64_t n, int64_t k,
                   int n_mod, int* __restrict__ res) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= m * n * n_mod) return;
  const int64_t i = w % m, j = (w / m) % n;
  const int t = (int)(w / (m * n));
  int re, im;
  i8_dot_plain(ra + (size_t)(2 * t) * m * k + i, ra + (size_t)(2 * t + 1) * m * k + i, m,
               rb + (size_t)(2 * t) * n * k + k * j, rb + (size_t)(2 * t + 1) * n * k + k * j, 1, k, i8_modulus(t), &re, &im);
  res[((size_t)(2 * t) * n + j) * m + i] = re;
  res[((size_t)(2 * t + 1) * n + j) * m + i] = im;
}

// number of moduli for FP64-level accuracy at contraction length K (55 bits per operand, scripts/ozaki_study.py)
static I8Tables tables_for(int64_t K) {
  for (int n = 8; n <= I8_MAX_MODULI; ++n) {
    I8Tables T = i8_make_tables(n, K);
    if (T.bits >= 55) return T;
  }
  return i8_make_tables(I8_MAX_MODULI, K);
}

// C (m x n) = A^H B,  A: k x m, B: k x n (column-major, complex).  tensor_cores: integer products by k_i8_gemm_tc (i8tc.cu)
void zgemm_i8_cn(dftk_b200_ctx* ctx, int64_t m, int64_t n, int64_t k, const cplx* A, int64_t lda, const cplx* B, int64_t ldb,
                 cplx* C, int64_t ldc, bool tensor_cores) {
  if (m == 0 || n == 0) return;
  REQUIRE(k >= 1 && m <= 65535 && n <= 65535, "zgemm_i8: unsupported shape");
  const I8Tables T = tables_for(2 * k);
  const int64_t ldk = tensor_cores ? (k + 127) / 128 * 128 : k;
  const int n_chunks = (int)((ldk + I8_K_CHUNK - 1) / I8_K_CHUNK);
  const size_t plane_a = (size_t)m * ldk, plane_b = (size_t)n * ldk;
  const size_t n_res = 2 * (size_t)T.n_mod * m * n;
  const size_t bytes = (size_t)(m + n) * sizeof(int) + 64 + n_res * sizeof(int) + 2 * (size_t)T.n_mod * (plane_a + plane_b) + 256 +
                       (tensor_cores ? n_res * n_chunks * sizeof(short) : 0);
  char* ws = (char*)ctx->gemm_ws.ensure(bytes);
  int* ea = (int*)ws;
  int* eb = ea + m;
  int* res = eb + n;
  signed char* ra = (signed char*)(res + n_res);
  ra += (16 - ((uintptr_t)ra & 15)) & 15;                       // cp.async needs 16-byte aligned rows
  signed char* rb = ra + 2 * (size_t)T.n_mod * plane_a;
  short* part = (short*)(rb + 2 * (size_t)T.n_mod * plane_b);
  LAUNCH(ctx, k_i8_col_exponent, (unsigned)m, 256, 0, A, lda, k, T.bits, ea);
  LAUNCH(ctx, k_i8_col_exponent, (unsigned)n, 256, 0, B, ldb, k, T.bits, eb);
  if (tensor_cores) {
    CUDA_CHECK(cudaMemsetAsync(ra, 0, 2 * (size_t)T.n_mod * (plane_a + plane_b), ctx->stream));      // zero K padding
    LAUNCH(ctx, k_i8_residues_ld, dim3((unsigned)((k + 255) / 256), (unsigned)m), 256, 0, A, lda, k, m, ldk, (const int*)ea,
           T.n_mod, ra);
    LAUNCH(ctx, k_i8_residues_ld, dim3((unsigned)((k + 255) / 256), (unsigned)n), 256, 0, B, ldb, k, n, ldk, (const int*)eb,
           T.n_mod, rb);
    i8tc_products(ctx, ra, rb, m, n, ldk, T.n_mod, part, res);
  } else {
    LAUNCH(ctx, k_i8_residues, dim3((unsigned)((k + 255) / 256), (unsigned)m), 256, 0, A, lda, k, m, (const int*)ea, T.n_mod, ra);
    LAUNCH(ctx, k_i8_residues, dim3((unsigned)((k + 255) / 256), (unsigned)n), 256, 0, B, ldb, k, n, (const int*)eb, T.n_mod, rb);
    const int64_t warps = m * n * T.n_mod;
    LAUNCH(ctx, k_i8_dot_ref, (unsigned)((warps * 32 + 255) / 256), 256, 0, (const signed char*)ra, (const signed char*)rb, m, n,
           k, T.n_mod, res);
  }
  LAUNCH(ctx, k_i8_crt, (unsigned)((m * n + 127) / 128), 128, 0, (const int*)res, m, n, T, (const int*)ea, (const int*)eb, C, ldc, 0);
}

// C (m x n) = A B (+ C if accumulate): reference pipeline only (integer products on CUDA cores); returns false when the
// residue planes would not fit the workspace budget (the caller then uses the DMMA kernel)
bool zgemm_i8_nn(dftk_b200_ctx* ctx, int64_t m, int64_t n, int64_t k, const cplx* A, int64_t lda, const cplx* B, int64_t ldb,
                 cplx* C, int64_t ldc, bool accumulate) {
  if (m == 0 || n == 0 || k == 0) return true;
  const I8Tables T = tables_for(2 * k);
  const size_t n_res = 2 * (size_t)T.n_mod * m * n;
  const size_t bytes = (size_t)(m + n) * sizeof(int) + 64 + n_res * sizeof(int) + 2 * (size_t)T.n_mod * (size_t)k * (m + n) + 64;
  if (bytes > ((size_t)8 << 30) || k > 65535 || m * n * T.n_mod > (int64_t)1 << 40) return false;
  char* ws = (char*)ctx->gemm_ws.ensure(bytes);
  int* ea = (int*)ws;
  int* eb = ea + m;
  int* res = eb + n;
  signed char* ra = (signed char*)(res + n_res);
  signed char* rb = ra + 2 * (size_t)T.n_mod * m * k;
  LAUNCH(ctx, k_i8_row_exponent, (unsigned)((m + 255) / 256), 256, 0, A, lda, m, k, T.bits, ea);
  LAUNCH(ctx, k_i8_col_exponent, (unsigned)n, 256, 0, B, ldb, k, T.bits, eb);
  LAUNCH(ctx, k_i8_residues_rows, dim3((unsigned)((m + 255) / 256), (unsigned)k), 256, 0, A, lda, m, k, (const int*)ea, T.n_mod, ra);
  LAUNCH(ctx, k_i8_residues, dim3((unsigned)((k + 255) / 256), (unsigned)n), 256, 0, B, ldb, k, n, (const int*)eb, T.n_mod, rb);
  const int64_t threads = m * n * T.n_mod;
  LAUNCH(ctx, k_i8_dot_plain_ref, (unsigned)((threads + 255) / 256), 256, 0, (const signed char*)ra, (const signed char*)rb, m, n,
         k, T.n_mod, res);
  LAUNCH(ctx, k_i8_crt, (unsigned)((m * n + 127) / 128), 128, 0, (const int*)res, m, n, T, (const int*)ea, (const int*)eb, C, ldc,
         accumulate ? 1 : 0);
  return true;
}

}  // namespace dftk
