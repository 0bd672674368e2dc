// Hellmann-Feynman force bodies (reference: src/terms/local.jl:152-181 forces_local,
// src/terms/nonlocal.jl:49-100).  __host__ __device__ so tests/hostemu can run them on the CPU.
//
// Local:    F_a,α = -Re( Σ_G -2πi G_α e^{-2πi G·r_a} w_G ) = -2π Im( Σ_G G_α e^{-2πi G·r_a} w_G ),
//           w_G = conj(ρ_G) v_loc(|G|) / sqrt(Ω) prepared by the caller per atom group.
// Nonlocal: the reference forms δHψ = P D (dP/dR_α)†ψ per atom and direction (3·n_atoms GEMM pairs of full
//           height).  With dP/dR_α = -2πi (G+k)_α P this is (dP/dR_α)†ψ = 2πi P†(p_α ψ), so four projections
//           P†[ψ, p_x ψ, p_y ψ, p_z ψ] (DMMA GEMMs) give every atom at once; per projector row j
//           f_j,α = Σ_n w_n 4π Im( conj((D P†ψ)_jn) (P† p_α ψ)_jn )   and F_a,α = Σ_{j in atom a} f_j,α.
#pragma once
#include <math.h>
#include "fft_core.cuh"

namespace dftk {

#define FORCES_PI 3.14159265358979323846

HD int force_freq_of_index(int i, int n) { return i <= (n - 1) / 2 ? i : i - n; }

// contribution of cube point idx to z_α = Σ_G G_α Im(e^{-2πi G·r} w_G), α = 0..2 (added to acc)
HD void local_force_point(int64_t idx, int nx, int ny, int nz, const cplx* __restrict__ w, double rx, double ry,
                          double rz, double* acc) {
  const int ix = (int)(idx % nx), iy = (int)((idx / nx) % ny), iz = (int)(idx / ((int64_t)nx * ny));
  const int gx = force_freq_of_index(ix, nx), gy = force_freq_of_index(iy, ny), gz = force_freq_of_index(iz, nz);
  double t = gx * rx + gy * ry + gz * rz;
  t -= rint(t);                                   // e^{-2πi t} is 1-periodic; keeps the argument small
  double s, c;
#ifdef __CUDA_ARCH__
  sincospi(-2.0 * t, &s, &c);
#else
  s = sin(-2.0 * FORCES_PI * t);
  c = cos(-2.0 * FORCES_PI * t);
#endif
  const cplx v = w[idx];
  const double im = c * v.y + s * v.x;            // Im((c + i s)(v.x + i v.y))
  acc[0] += gx * im;
  acc[1] += gy * im;
  acc[2] += gz * im;
}

// out[i + n_rows * (b + nb * a)] = gpk[a * n_rows + i] * psi[i + ld * b]   (a = 0..2)
HD void scale_by_momentum_point(int64_t i, int64_t b, int a, int64_t n_rows, int64_t nb, const double* __restrict__ gpk,
                                const cplx* __restrict__ psi, int64_t ld, cplx* __restrict__ out) {
  const double p = gpk[(int64_t)a * n_rows + i];
  const cplx v = psi[i + ld * b];
  out[i + n_rows * (b + nb * a)] = make_double2(p * v.x, p * v.y);
}

// f[a * np + j] += Σ_n w[n] 4π Im(conj(dproj[j + np n]) * pa[j + np (n + nb a)])
HD void nonlocal_force_row(int64_t j, int a, int64_t np, int64_t nb, const cplx* __restrict__ dproj,
                           const cplx* __restrict__ pa, const double* __restrict__ w, double* __restrict__ f) {
  double s = 0.0;
  for (int64_t n = 0; n < nb; ++n) {
    const cplx d = dproj[j + np * n], q = pa[j + np * (n + nb * a)];
    s += w[n] * (d.x * q.y - d.y * q.x);
  }
  f[(int64_t)a * np + j] += 4.0 * FORCES_PI * s;
}

}  // namespace dftk
