// Force kernels next to the hot path: local-potential forces (src/terms/local.jl:152-181) and
// nonlocal-projector forces (src/terms/nonlocal.jl:49-100).  Bodies live in forces_core.cuh.
#include "structs.cuh"
#include "forces_core.cuh"

namespace dftk {

// grid (blocks_per_atom, n_atoms); partial[(atom * gridDim.x + blockIdx.x) * 3 + α]
__global__ void __launch_bounds__(256)
k_local_forces(int nx, int ny, int nz, const cplx* w, const double* pos, double* partial) {
  const int atom = blockIdx.y;
  const double rx = pos[3 * atom], ry = pos[3 * atom + 1], rz = pos[3 * atom + 2];
  const int64_t N = (int64_t)nx * ny * nz;
  double acc[3] = {0.0, 0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
    local_force_point(i, nx, ny, nz, w, rx, ry, rz, acc);
  __shared__ double red[3][8];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    double v = acc[a];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[a][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double v = 0.0;
    for (int wv = 0; wv < 8; ++wv) v += red[threadIdx.x][wv];
    partial[((int64_t)atom * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = v;
  }
}

// one thread per (atom, α): fixed-order sum of the block partials, F = -2π Σ
__global__ void k_local_forces_reduce(const double* partial, int n_atoms, int n_blocks, double* out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * n_atoms) return;
  const int atom = t / 3, a = t % 3;
  double s = 0.0;
  for (int b = 0; b < n_blocks; ++b) s += partial[((int64_t)atom * n_blocks + b) * 3 + a];
  out[t] = -2.0 * FORCES_PI * s;
}

__global__ void __launch_bounds__(256)
k_scale_by_momentum(int64_t n_rows, int64_t nb, const double* gpk, const cplx* psi, int64_t ld, cplx* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  scale_by_momentum_point(i, blockIdx.y, blockIdx.z, n_rows, nb, gpk, psi, ld, out);
}

__global__ void k_nonlocal_force_rows(int64_t np, int64_t nb, const cplx* dproj, const cplx* pa, const double* w,
                                      double* f) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= np) return;
  nonlocal_force_row(j, blockIdx.y, np, nb, dproj, pa, w, f);
}

void local_forces(dftk_b200_grid* g, const cplx* w, int n_atoms, const double* pos_host, double* out_host) {
  dftk_b200_ctx* ctx = g->ctx;
  if (n_atoms == 0) return;
  int n_blocks = std::max(1, (4 * ctx->sm_count + n_atoms - 1) / n_atoms);
  n_blocks = (int)std::min<int64_t>(n_blocks, (g->N + 255) / 256);
  double* d = ctx->sym_d.ensure((size_t)3 * n_atoms * (2 + n_blocks));
  double* d_pos = d;
  double* d_out = d + 3 * n_atoms;
  double* d_part = d + 6 * n_atoms;
  CUDA_CHECK(cudaMemcpyAsync(d_pos, pos_host, (size_t)3 * n_atoms * sizeof(double), cudaMemcpyDefault, ctx->stream));
  for (int a0 = 0; a0 < n_atoms; a0 += 65535) {     // gridDim.y limit
    const int na = std::min(65535, n_atoms - a0);
    LAUNCH(ctx, k_local_forces, dim3((unsigned)n_blocks, (unsigned)na), 256, 0, g->nx, g->ny, g->nz, w,
           (const double*)(d_pos + 3 * a0), d_part + (size_t)3 * a0 * n_blocks);
  }
  LAUNCH(ctx, k_local_forces_reduce, (unsigned)((3 * n_atoms + 127) / 128), 128, 0, (const double*)d_part, n_atoms,
         n_blocks, d_out);
  CUDA_CHECK(cudaMemcpyAsync(out_host, d_out, (size_t)3 * n_atoms * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

void kb_nonlocal_force_rows(dftk_b200_kblock* kb, const cplx* psi, const double* occ_w_host, int64_t n_bands,
                            const double* gpk, double* out_host) {
  dftk_b200_ctx* ctx = kb->grid->ctx;
  const int64_t np = kb->n_proj, n_pw = kb->n_pw;
  for (int64_t i = 0; i < 3 * np; ++i) out_host[i] = 0.0;
  if (np == 0 || n_bands == 0) return;
  REQUIRE(n_pw <= (int64_t)2147483647 * 256, "nonlocal forces: n_pw too large");
  // bands in chunks so that the three scaled copies of psi stay within 1 GiB of scratch
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(n_bands, 65535),
                                                               ((int64_t)1 << 30) / (3 * n_pw * (int64_t)sizeof(cplx))));
  cplx* scaled = kb->W1.ensure((size_t)3 * n_pw * chunk);
  cplx* proj = kb->proj.ensure((size_t)5 * np * chunk);     // P'psi | D P'psi | P'(p_α psi) α = 0..2
  cplx* dproj = proj + (size_t)np * chunk;
  cplx* pa = proj + (size_t)2 * np * chunk;
  kb->wts.upload(occ_w_host, n_bands, ctx->stream);
  double* f = ctx->scal.ensure((size_t)3 * np + 8);
  CUDA_CHECK(cudaMemsetAsync(f, 0, (size_t)3 * np * sizeof(double), ctx->stream));
  const cplx one = make_double2(1.0, 0.0), zero = make_double2(0.0, 0.0);
  for (int64_t b0 = 0; b0 < n_bands; b0 += chunk) {
    const int64_t nb = std::min(chunk, n_bands - b0);
    const cplx* p = psi + (size_t)b0 * n_pw;
    LAUNCH(ctx, k_scale_by_momentum, dim3((unsigned)((n_pw + 255) / 256), (unsigned)nb, 3), 256, 0, n_pw, nb, gpk, p,
           n_pw, scaled);
    zgemm(ctx, 2, np, nb, n_pw, one, kb->P.p, n_pw, p, n_pw, zero, proj, np);
    zgemm(ctx, 0, np, nb, np, one, kb->Dc.p, np, proj, np, zero, dproj, np);
    zgemm(ctx, 2, np, 3 * nb, n_pw, one, kb->P.p, n_pw, scaled, n_pw, zero, pa, np);
    LAUNCH(ctx, k_nonlocal_force_rows, dim3((unsigned)((np + 127) / 128), 3), 128, 0, np, nb, (const cplx*)dproj,
           (const cplx*)pa, (const double*)(kb->wts.p + b0), f);
  }
  CUDA_CHECK(cudaMemcpyAsync(out_host, f, (size_t)3 * np * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

}  // namespace dftk
