// Pointwise exchange-correlation functionals and the density symmetrisation gather (device side of the SCF
// plumbing next to the hot path; K16 of SURVEY §2.5).
//
// XC: the reference evaluates libxc through Libxc.jl (src/DispatchFunctional.jl:55-56,108-128; call site
// src/terms/xc.jl:104-113).  libxc is third-party code that is not under /root/reference; the closed forms are
// restated here (Dirac exchange, VWN5, PW92 / PW92-mod, PBE) with libxc's constants.  Energies per volume `e`
// and the derivatives vrho / vsigma come from ONE expression evaluated on forward-mode dual numbers, so they are
// mutually consistent by construction.
//
// Bodies are __host__ __device__ (host emulation in tests/hostemu).
#pragma once
#include <math.h>
#include "fft_core.cuh"

namespace dftk {

template <int NV>
struct Dual {
  double v;
  double d[NV];
};
template <int NV>
HD Dual<NV> dconst(double c) {
  Dual<NV> r;
  r.v = c;
#pragma unroll
  for (int i = 0; i < NV; ++i) r.d[i] = 0.0;
  return r;
}
template <int NV>
HD Dual<NV> dvar(double x, int idx) {
  Dual<NV> r = dconst<NV>(x);
  r.d[idx] = 1.0;
  return r;
}
#define DUAL_BIN(op, VEXPR, DEXPR)                                      \
  template <int NV>                                                     \
  HD Dual<NV> operator op(const Dual<NV>& a, const Dual<NV>& b) {       \
    Dual<NV> r;                                                         \
    r.v = VEXPR;                                                        \
    _Pragma("unroll") for (int i = 0; i < NV; ++i) r.d[i] = DEXPR;      \
    return r;                                                           \
  }
DUAL_BIN(+, a.v + b.v, a.d[i] + b.d[i])
DUAL_BIN(-, a.v - b.v, a.d[i] - b.d[i])
DUAL_BIN(*, a.v * b.v, a.d[i] * b.v + b.d[i] * a.v)
DUAL_BIN(/, a.v / b.v, (a.d[i] - b.d[i] * (a.v / b.v)) / b.v)
#undef DUAL_BIN
template <int NV> HD Dual<NV> operator+(const Dual<NV>& a, double c) { Dual<NV> r = a; r.v += c; return r; }
template <int NV> HD Dual<NV> operator+(double c, const Dual<NV>& a) { return a + c; }
template <int NV> HD Dual<NV> operator-(const Dual<NV>& a, double c) { return a + (-c); }
template <int NV>
HD Dual<NV> operator-(const Dual<NV>& a) {
  Dual<NV> r;
  r.v = -a.v;
#pragma unroll
  for (int i = 0; i < NV; ++i) r.d[i] = -a.d[i];
  return r;
}
template <int NV> HD Dual<NV> operator-(double c, const Dual<NV>& a) { return (-a) + c; }
template <int NV>
HD Dual<NV> operator*(const Dual<NV>& a, double c) {
  Dual<NV> r;
  r.v = a.v * c;
#pragma unroll
  for (int i = 0; i < NV; ++i) r.d[i] = a.d[i] * c;
  return r;
}
template <int NV> HD Dual<NV> operator*(double c, const Dual<NV>& a) { return a * c; }
template <int NV> HD Dual<NV> operator/(const Dual<NV>& a, double c) { return a * (1.0 / c); }
template <int NV> HD Dual<NV> operator/(double c, const Dual<NV>& a) { return dconst<NV>(c) / a; }
template <int NV>
HD Dual<NV> dchain(const Dual<NV>& a, double fv, double g) {   // f(a) with f(a.v) = fv, f'(a.v) = g
  Dual<NV> r;
  r.v = fv;
#pragma unroll
  for (int i = 0; i < NV; ++i) r.d[i] = a.d[i] * g;
  return r;
}
template <int NV> HD Dual<NV> dlog(const Dual<NV>& a) { return dchain(a, log(a.v), 1.0 / a.v); }
template <int NV> HD Dual<NV> dexp(const Dual<NV>& a) { double e = exp(a.v); return dchain(a, e, e); }
template <int NV> HD Dual<NV> dsqrt(const Dual<NV>& a) { double q = sqrt(a.v); return dchain(a, q, 0.5 / q); }
template <int NV> HD Dual<NV> datan(const Dual<NV>& a) { return dchain(a, atan(a.v), 1.0 / (1.0 + a.v * a.v)); }
template <int NV> HD Dual<NV> dcbrt(const Dual<NV>& a) { double c = cbrt(a.v); return dchain(a, c, c / (3.0 * a.v)); }
template <int NV> HD Dual<NV> dpow(const Dual<NV>& a, double p) { return dchain(a, pow(a.v, p), p * pow(a.v, p - 1.0)); }

#define XC_LDA_X 1
#define XC_LDA_C_VWN 2
#define XC_LDA_C_PW 4
#define XC_GGA_X_PBE 8
#define XC_GGA_C_PBE 16
#define XC_DENS_THRESHOLD 1e-15

#define XC_PI 3.14159265358979323846
template <class T> HD T xc_fzeta(const T& z) {
  return (dpow(1.0 + z, 4.0 / 3.0) + dpow(1.0 - z, 4.0 / 3.0) - 2.0) / (2.5198420997897464 - 2.0);   // 2^(4/3) - 2
}
template <class T> HD T xc_ex_unif(const T& n) { return (-0.75 * 0.98474502184269641) * n * dcbrt(n); }   // (3/pi)^(1/3)

template <class T>
HD T xc_vwn_piece(const T& x, double A, double b, double c, double x0) {
  const double Q = sqrt(4.0 * c - b * b);
  T X = x * x + b * x + c;
  const double X0 = x0 * x0 + b * x0 + c;
  T at = datan(Q / (2.0 * x + b));
  return A * (dlog(x * x / X) + (2.0 * b / Q) * at -
              (b * x0 / X0) * (dlog((x - x0) * (x - x0) / X) + (2.0 * (b + 2.0 * x0) / Q) * at));
}
template <class T>
HD T xc_ec_vwn(const T& rs, const T* zeta) {
  T x = dsqrt(rs);
  T p0 = xc_vwn_piece(x, 0.0310907, 3.72744, 12.9352, -0.10498);
  if (!zeta) return p0;
  T p1 = xc_vwn_piece(x, 0.01554535, 7.06042, 18.0578, -0.32500);
  T p2 = xc_vwn_piece(x, -1.0 / (6.0 * XC_PI * XC_PI), 1.13107, 13.0045, -0.0047584);
  T fz = xc_fzeta(*zeta);
  T z2 = (*zeta) * (*zeta);
  T z4 = z2 * z2;
  const double fpp0 = 4.0 / (9.0 * (1.2599210498948732 - 1.0));   // 4 / (9 (2^(1/3) - 1))
  return p0 + p2 * fz * (1.0 - z4) / fpp0 + (p1 - p0) * fz * z4;
}
template <class T>
HD T xc_pw_G(const T& rs, double a, double a1, double b1, double b2, double b3, double b4) {
  T s = dsqrt(rs);
  T den = (2.0 * a) * (b1 * s + b2 * rs + b3 * rs * s + b4 * rs * rs);
  return (-2.0 * a) * (1.0 + a1 * rs) * dlog(1.0 + 1.0 / den);
}
template <class T>
HD T xc_ec_pw(const T& rs, const T* zeta, bool mod) {
  const double a0 = mod ? 0.0310906908696548950 : 0.0310907, a1 = mod ? 0.01554534543482744750 : 0.01554535,
               a2 = mod ? 0.0168868639404617 : 0.0168869, fz20 = mod ? 1.709920934161365617563962776245 : 1.709921;
  T g0 = xc_pw_G(rs, a0, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294);
  if (!zeta) return g0;
  T g1 = xc_pw_G(rs, a1, 0.20548, 14.1189, 6.1977, 3.3662, 0.62517);
  T mac = xc_pw_G(rs, a2, 0.11125, 10.357, 3.6231, 0.88026, 0.49671);
  T fz = xc_fzeta(*zeta);
  T z2 = (*zeta) * (*zeta);
  T z4 = z2 * z2;
  return g0 - mac * fz * (1.0 - z4) / fz20 + (g1 - g0) * fz * z4;
}
#define XC_KAPPA 0.8040
#define XC_BETA 0.06672455060314922
template <class T>
HD T xc_ex_pbe(const T& n, const T& sigma) {
  const double mu = XC_BETA * (XC_PI * XC_PI / 3.0);
  T kF = dcbrt((3.0 * XC_PI * XC_PI) * n);
  T s2 = sigma / (4.0 * kF * kF * n * n);
  return xc_ex_unif(n) * ((1.0 + XC_KAPPA) - XC_KAPPA / (1.0 + (mu / XC_KAPPA) * s2));
}
template <class T>
HD T xc_ec_pbe(const T& n, const T& rs, const T* zeta, const T& sigma) {
  const double gamma = (1.0 - 0.69314718055994531) / (XC_PI * XC_PI);
  T ec = xc_ec_pw(rs, zeta, true);
  T phi2 = ec * 0.0 + 1.0, phi3 = ec * 0.0 + 1.0;
  if (zeta) {
    T phi = (dpow(1.0 + *zeta, 2.0 / 3.0) + dpow(1.0 - *zeta, 2.0 / 3.0)) * 0.5;
    phi2 = phi * phi;
    phi3 = phi2 * phi;
  }
  T kF = dcbrt((3.0 * XC_PI * XC_PI) * n);
  T t2 = sigma / (4.0 * phi2 * ((4.0 / XC_PI) * kF) * n * n);
  T Aa = (XC_BETA / gamma) / (dexp(-ec / (gamma * phi3)) - 1.0);
  T At2 = Aa * t2;
  return ec + gamma * phi3 * dlog(1.0 + (XC_BETA / gamma) * t2 * (1.0 + At2) / (1.0 + At2 + At2 * At2));
}

// One grid point.  rho: n_spin values; sigma: 1 (unpolarised) or 3 (uu, ud, dd) values, ignored for LDA.
// NV = n_spin (LDA) or n_spin + n_sigma (GGA).  Outputs: e, vrho[n_spin], vsigma[n_sigma].
template <int NSPIN, bool GGA>
HD void xc_point(int mask, const double* rho, const double* sigma, double* e, double* vrho, double* vsigma) {
  constexpr int NSIG = GGA ? (NSPIN == 1 ? 1 : 3) : 0;
  constexpr int NV = NSPIN + NSIG;
  typedef Dual<NV> T;
  double tot = 0.0;
  for (int s = 0; s < NSPIN; ++s) tot += rho[s];
  if (!(tot > XC_DENS_THRESHOLD)) {
    *e = 0.0;
    for (int s = 0; s < NSPIN; ++s) vrho[s] = 0.0;
    for (int s = 0; s < NSIG; ++s) vsigma[s] = 0.0;
    return;
  }
  T r[NSPIN];
  for (int s = 0; s < NSPIN; ++s) r[s] = dvar<NV>(rho[s], s);
  T sg[NSIG > 0 ? NSIG : 1];
  for (int s = 0; s < NSIG; ++s) sg[s] = dvar<NV>(sigma[s], NSPIN + s);
  T n = r[0];
  T zeta = dconst<NV>(0.0);
  if (NSPIN == 2) {
    n = r[0] + r[1];
    zeta = (r[0] - r[1]) / n;
    if (zeta.v > 1.0 - 1e-14) zeta.v = 1.0 - 1e-14;
    if (zeta.v < -1.0 + 1e-14) zeta.v = -1.0 + 1e-14;
  }
  const T* zp = NSPIN == 2 ? &zeta : nullptr;
  T rs = 0.62035049089940009 / dcbrt(n);   // (3/(4 pi))^(1/3)
  T acc = dconst<NV>(0.0);
  if (mask & XC_LDA_X) {
    if (NSPIN == 1) acc = acc + xc_ex_unif(n);
    else
      for (int s = 0; s < NSPIN; ++s) {
        T rr = r[s];
        if (rr.v < 1e-30) rr.v = 1e-30;
        acc = acc + 0.5 * xc_ex_unif(2.0 * rr);
      }
  }
  if (mask & XC_LDA_C_VWN) acc = acc + n * xc_ec_vwn(rs, zp);
  if (mask & XC_LDA_C_PW) acc = acc + n * xc_ec_pw(rs, zp, false);
  if (GGA && (mask & XC_GGA_X_PBE)) {
    if (NSPIN == 1) acc = acc + xc_ex_pbe(n, sg[0]);
    else
      for (int s = 0; s < NSPIN; ++s) {
        T rr = r[s];
        if (rr.v < 1e-30) rr.v = 1e-30;
        acc = acc + 0.5 * xc_ex_pbe(2.0 * rr, 4.0 * sg[s == 0 ? 0 : (NSIG - 1)]);
      }
  }
  if (GGA && (mask & XC_GGA_C_PBE)) {
    T st = sg[0];
    if (NSPIN == 2) st = sg[0] + 2.0 * sg[NSIG > 1 ? 1 : 0] + sg[NSIG > 2 ? 2 : 0];
    acc = acc + n * xc_ec_pbe(n, rs, zp, st);
  }
  *e = acc.v;
  for (int s = 0; s < NSPIN; ++s) vrho[s] = acc.d[s];
  for (int s = 0; s < NSIG; ++s) vsigma[s] = acc.d[NSPIN + s];
}

// rho, sigma, vrho, vsigma are stored component-major: x[component * N + i]
template <int NSPIN, bool GGA>
HD void xc_eval_range(int mask, int64_t i, int64_t N, const double* rho, const double* sigma, double* e,
                      double* vrho, double* vsigma) {
  constexpr int NSIG = GGA ? (NSPIN == 1 ? 1 : 3) : 0;
  double r[NSPIN], s[NSIG > 0 ? NSIG : 1], vr[NSPIN], vs[NSIG > 0 ? NSIG : 1], ee;
  for (int c = 0; c < NSPIN; ++c) r[c] = rho[c * N + i];
  for (int c = 0; c < NSIG; ++c) s[c] = sigma[c * N + i];
  xc_point<NSPIN, GGA>(mask, r, s, &ee, vr, vs);
  e[i] = ee;
  for (int c = 0; c < NSPIN; ++c) vrho[c * N + i] = vr[c];
  for (int c = 0; c < NSIG; ++c) vsigma[c * N + i] = vs[c];
}

// ---- accumulate_over_symmetries! (src/symmetry.jl:282-327): out[G] = (1/n_sym) sum_s e^{-2 pi i G.tau_s} in[S_s^-1 G]
HD int wrap_index(int g, int n) {   // integer frequency -> array index, or -1 if outside the FFT box
  const int start = -(n / 2), stop = (n - 1) / 2;
  if (g < start || g > stop) return -1;
  return g < 0 ? g + n : g;
}
HD int freq_of_index(int i, int n) { return i <= (n - 1) / 2 ? i : i - n; }
HD void symmetrize_point(int64_t idx, int nx, int ny, int nz, const cplx* __restrict__ in, cplx* __restrict__ out,
                         int n_sym, const int* __restrict__ invS, const double* __restrict__ tau) {
  const int ix = (int)(idx % nx), iy = (int)((idx / nx) % ny), iz = (int)(idx / ((int64_t)nx * ny));
  const int G[3] = {freq_of_index(ix, nx), freq_of_index(iy, ny), freq_of_index(iz, nz)};
  double ax = 0.0, ay = 0.0;
  for (int s = 0; s < n_sym; ++s) {
    const int* M = invS + 9 * s;
    const int g0 = M[0] * G[0] + M[1] * G[1] + M[2] * G[2];
    const int g1 = M[3] * G[0] + M[4] * G[1] + M[5] * G[2];
    const int g2 = M[6] * G[0] + M[7] * G[1] + M[8] * G[2];
    const int j0 = wrap_index(g0, nx), j1 = wrap_index(g1, ny), j2 = wrap_index(g2, nz);
    if (j0 < 0 || j1 < 0 || j2 < 0) continue;
    cplx v = in[j0 + (int64_t)nx * (j1 + (int64_t)ny * j2)];
    const double* t = tau + 3 * s;
    if (t[0] != 0.0 || t[1] != 0.0 || t[2] != 0.0) {
      const double ph = -2.0 * XC_PI * (G[0] * t[0] + G[1] * t[1] + G[2] * t[2]);
      v = cmul(v, make_double2(cos(ph), sin(ph)));
    }
    ax += v.x;
    ay += v.y;
  }
  out[idx] = make_double2(ax / n_sym, ay / n_sym);
}

}  // namespace dftk
