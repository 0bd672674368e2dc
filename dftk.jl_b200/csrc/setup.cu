// Setup kernels next to the hot path: the O(n_atoms x N) structure-factor work that dominates the
// time to the first SCF step of large cells.
//   * structure factors on the FFT cube (build_local_potential, src/terms/local.jl:108-138; guess_density,
//     src/density_methods.jl:103-181): out[G] = sum_a c_a exp(-2 pi i G.r_a), G from the cube index
//   * projector table of a k-block (build_projection_vectors, src/terms/nonlocal.jl:166-199):
//     P[(a, p), G] = exp(-2 pi i (G+k).r_a) ff[p, G]
#include "structs.cuh"

namespace dftk {

__device__ __forceinline__ int wrapped_index(int i, int n) { return i <= (n - 1) / 2 ? i : i - n; }   // src/fft.jl:24-31

// one thread per cube point, atoms staged through shared memory in tiles of 256
__global__ void __launch_bounds__(256)
k_structure_factor(int nx, int ny, int nz, int n_atoms, const double* __restrict__ pos, const double* __restrict__ coeff,
                   cplx* __restrict__ out) {
  __shared__ double sp[256 * 4];
  const int64_t N = (int64_t)nx * ny * nz;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = idx < N;
  const int ix = (int)(idx % nx), iy = (int)((idx / nx) % ny), iz = (int)(idx / ((int64_t)nx * ny));
  const double gx = wrapped_index(ix, nx), gy = wrapped_index(iy, ny), gz = wrapped_index(live ? iz : 0, nz);
  double re = 0.0, im = 0.0;
  for (int a0 = 0; a0 < n_atoms; a0 += 256) {
    const int na = min(256, n_atoms - a0);
    __syncthreads();
    if ((int)threadIdx.x < na) {
      sp[4 * threadIdx.x] = pos[3 * (a0 + threadIdx.x)];
      sp[4 * threadIdx.x + 1] = pos[3 * (a0 + threadIdx.x) + 1];
      sp[4 * threadIdx.x + 2] = pos[3 * (a0 + threadIdx.x) + 2];
      sp[4 * threadIdx.x + 3] = coeff ? coeff[a0 + threadIdx.x] : 1.0;
    }
    __syncthreads();
    if (live)
      for (int a = 0; a < na; ++a) {
        double s, c;
        sincospi(-2.0 * (gx * sp[4 * a] + gy * sp[4 * a + 1] + gz * sp[4 * a + 2]), &s, &c);
        re += sp[4 * a + 3] * c;
        im += sp[4 * a + 3] * s;
      }
  }
  if (live) out[idx] = make_double2(re, im);
}

// grid (ceil(n_pw/256), n_atoms): all n_rows projectors of one atom for 256 plane waves
__global__ void __launch_bounds__(256)
k_build_projectors(int64_t n_pw, const double* __restrict__ gpk, const double* __restrict__ pos, int n_rows,
                   const cplx* __restrict__ ff, cplx* __restrict__ P) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pw) return;
  const int a = blockIdx.y;
  double s, c;
  sincospi(-2.0 * (gpk[i] * pos[3 * a] + gpk[n_pw + i] * pos[3 * a + 1] + gpk[2 * n_pw + i] * pos[3 * a + 2]), &s, &c);
  for (int p = 0; p < n_rows; ++p) {
    const cplx f = ff[(int64_t)p * n_pw + i];
    P[((int64_t)a * n_rows + p) * n_pw + i] = make_double2(c * f.x - s * f.y, c * f.y + s * f.x);
  }
}

void structure_factor(dftk_b200_grid* g, int n_atoms, const double* pos_host, const double* coeff_host, cplx* out) {
  dftk_b200_ctx* ctx = g->ctx;
  double* d = ctx->sym_d.ensure((size_t)4 * n_atoms + 8);
  CUDA_CHECK(cudaMemcpyAsync(d, pos_host, (size_t)3 * n_atoms * sizeof(double), cudaMemcpyDefault, ctx->stream));
  if (coeff_host) CUDA_CHECK(cudaMemcpyAsync(d + 3 * n_atoms, coeff_host, (size_t)n_atoms * sizeof(double), cudaMemcpyDefault, ctx->stream));
  LAUNCH(ctx, k_structure_factor, (unsigned)((g->N + 255) / 256), 256, 0, g->nx, g->ny, g->nz, n_atoms, (const double*)d,
         coeff_host ? (const double*)(d + 3 * n_atoms) : (const double*)nullptr, out);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));     // the host arrays may go away
}

void build_projectors(dftk_b200_ctx* ctx, int64_t n_pw, const double* gpk, int n_atoms, const double* pos_host, int n_rows,
                      const cplx* ff, cplx* P) {
  if (n_atoms == 0 || n_rows == 0) return;
  REQUIRE(n_atoms <= 65535, "build_projectors: too many atoms in one group");
  double* d = ctx->sym_d.ensure((size_t)3 * n_atoms + 8);
  CUDA_CHECK(cudaMemcpyAsync(d, pos_host, (size_t)3 * n_atoms * sizeof(double), cudaMemcpyDefault, ctx->stream));
  LAUNCH(ctx, k_build_projectors, dim3((unsigned)((n_pw + 255) / 256), (unsigned)n_atoms), 256, 0, n_pw, gpk, (const double*)d,
         n_rows, ff, P);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

}  // namespace dftk
