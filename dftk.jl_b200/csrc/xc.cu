// Kernels for the SCF plumbing adjacent to the hot path: pointwise XC (replaces the Libxc dispatch of
// ext/DFTKCUDAExt.jl:17-25 / src/terms/xc.jl:104-113) and the symmetrisation gather of
// accumulate_over_symmetries! (src/symmetry.jl:282-327).  Bodies live in xc_core.cuh.
#include "structs.cuh"
#include "xc_core.cuh"

namespace dftk {

template <int NSPIN, bool GGA>
__global__ void __launch_bounds__(128)
k_xc(int mask, int64_t N, const double* rho, const double* sigma, double* e, double* vrho, double* vsigma) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) xc_eval_range<NSPIN, GGA>(mask, i, N, rho, sigma, e, vrho, vsigma);
}

__global__ void __launch_bounds__(256)
k_symmetrize(int nx, int ny, int nz, const cplx* in, cplx* out, int n_sym, const int* invS, const double* tau) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (int64_t)nx * ny * nz) symmetrize_point(i, nx, ny, nz, in, out, n_sym, invS, tau);
}

void xc_evaluate(dftk_b200_ctx* ctx, int mask, int n_spin, bool gga, int64_t N, const double* rho,
                 const double* sigma, double* e, double* vrho, double* vsigma) {
  if (N == 0) return;
  unsigned grid = (unsigned)((N + 127) / 128);
  if (n_spin == 1 && !gga) LAUNCH(ctx, (k_xc<1, false>), grid, 128, 0, mask, N, rho, sigma, e, vrho, vsigma);
  else if (n_spin == 2 && !gga) LAUNCH(ctx, (k_xc<2, false>), grid, 128, 0, mask, N, rho, sigma, e, vrho, vsigma);
  else if (n_spin == 1 && gga) LAUNCH(ctx, (k_xc<1, true>), grid, 128, 0, mask, N, rho, sigma, e, vrho, vsigma);
  else if (n_spin == 2 && gga) LAUNCH(ctx, (k_xc<2, true>), grid, 128, 0, mask, N, rho, sigma, e, vrho, vsigma);
  else throw Error(DFTK_B200_EINVAL, "xc_evaluate: n_spin must be 1 or 2");
}

void symmetrize_fourier(dftk_b200_grid* g, const cplx* in, cplx* out, int n_sym, const int* invS_host,
                        const double* tau_host) {
  dftk_b200_ctx* ctx = g->ctx;
  int* dS = (int*)ctx->sym_i.ensure((size_t)9 * n_sym);
  double* dT = ctx->sym_d.ensure((size_t)3 * n_sym);
  CUDA_CHECK(cudaMemcpyAsync(dS, invS_host, (size_t)9 * n_sym * sizeof(int), cudaMemcpyDefault, ctx->stream));
  CUDA_CHECK(cudaMemcpyAsync(dT, tau_host, (size_t)3 * n_sym * sizeof(double), cudaMemcpyDefault, ctx->stream));
  LAUNCH(ctx, k_symmetrize, (unsigned)((g->N + 255) / 256), 256, 0, g->nx, g->ny, g->nz, in, out, n_sym,
         (const int*)dS, (const double*)dT);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));   // host tables may be released by the caller
}

}  // namespace dftk
