// Host-side plan / table construction for the FFT engine (pure host C++; shared with tests/hostemu).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <stdexcept>
#include <vector>
#include "fft_core.cuh"

namespace dftk {

inline FftPlan make_plan(int n) {
  FftPlan p;
  p.n = n;
  p.npass = 0;
  int m = n;
  auto push = [&](int r) {
    if (p.npass >= DFTK_MAX_PASSES) throw std::runtime_error("fft size has too many factors");
    p.radix[p.npass++] = r;
  };
  // odd factors first (their twiddles are then trivial: k == 0 in the first passes), radix 4 last
  for (int r : {5, 3}) {
    while (m % r == 0) {
      push(r);
      m /= r;
    }
  }
  for (int r = 7; m > 1 && r <= m; r += 2) {
    while (m % r == 0) {
      push(r);
      m /= r;
    }
  }
  // remaining power of two
  std::vector<int> two;
  while (m % 4 == 0) {
    two.push_back(4);
    m /= 4;
  }
  if (m % 2 == 0) {
    two.push_back(2);
    m /= 2;
  }
  if (m != 1) push(m);  // large prime
  for (int r : two) push(r);
  if (p.npass == 0) push(1 == n ? 1 : n);
  return p;
}

inline std::vector<double> make_twiddles(int n) {
  std::vector<double> tw(2 * (size_t)n);
  for (int m = 0; m < n; ++m) {
    long double ang = -2.0L * 3.141592653589793238462643383279502884L * (long double)m / (long double)n;
    tw[2 * m] = (double)cosl(ang);
    tw[2 * m + 1] = (double)sinl(ang);
  }
  return tw;
}

// lines per CTA tile: as many as fit comfortably in shared memory (two ping-pong buffers)
inline int choose_lines(int n, size_t smem_budget = 100 * 1024) {
  for (int L : {16, 8, 4, 2, 1}) {
    int Lp = L | 1;
    if (2 * (size_t)n * Lp * sizeof(cplx) <= smem_budget) return L;
  }
  throw std::runtime_error("fft axis too long for the shared-memory engine");
}

// Factor pair (A, B), n = A*B, for the register two-pass engine (fft_reg.cuh); {0,0} = use the generic
// Stockham engine.  Only pairs instantiated in fft_reg.cu may be listed here.
#define DFTK_REG_PAIRS(X) \
  X(3, 5) X(4, 4) X(3, 6) X(4, 5) X(4, 6) X(5, 5) X(3, 9) X(5, 6) X(4, 8) X(6, 6) X(5, 8) X(5, 9) X(6, 8) \
  X(6, 9) X(6, 10) X(8, 8) X(8, 9) X(5, 15) X(8, 10) X(9, 10) X(8, 12) X(10, 10) X(9, 12) X(10, 12) X(5, 25) \
  X(8, 16) X(9, 15) X(12, 12) X(10, 15) X(10, 16) X(12, 15) X(12, 16) X(10, 20) X(12, 18) X(15, 15) X(15, 16) \
  X(16, 16)
inline void reg_pair_for(int n, int* A, int* B) {
  *A = 0;
  *B = 0;
#define DFTK_X(a, b) if (n == (a) * (b) && *A == 0) { *A = (a); *B = (b); }
  DFTK_REG_PAIRS(DFTK_X)
#undef DFTK_X
}

struct SphereTablesHost {
  int nx, ny, nz;
  int64_t n_pw;
  int n_cols, cnt_max, n_zc;
  std::vector<int> col_start, col_cnt, slot_ix, slot_src, zlist, colmap, col_y, col_z, zc_of;
  // Range form of the pruning maps: a convex sphere occupies at most two contiguous index ranges per axis
  // (one if it does not wrap).  Planes: z in [z_s0, z_s0+z_n0) U [z_s1, z_s1+z_n1), numbered in that order.
  // In plane izc the columns are y in [pl_s0, +pl_n0) U [pl_s1, +pl_n1), numbered from pl_col0 in that order.
  // ranges_ok == 0 means the structure does not hold and the kernels use the lookup tables.
  int ranges_ok = 0, z_s0 = 0, z_n0 = 0, z_s1 = 0, z_n1 = 0;
  std::vector<int> pl_s0, pl_n0, pl_s1, pl_n1, pl_col0;
  // per column: sphere points are x in [cx_s0, +cx_n0) U [cx_s1, +cx_n1) in slot order (requires an ascending mapping)
  std::vector<int> cx_s0, cx_n0, cx_s1, cx_n1;
};

// mapping: 0-based linear cube indices (x fastest) of the sphere coefficients, any order.
inline SphereTablesHost build_sphere_tables(int nx, int ny, int nz, int64_t n_pw,
                                            const int64_t* mapping) {
  SphereTablesHost T;
  T.nx = nx;
  T.ny = ny;
  T.nz = nz;
  T.n_pw = n_pw;
  const int64_t N = (int64_t)nx * ny * nz;
  std::vector<int> order(n_pw);
  std::iota(order.begin(), order.end(), 0);
  bool sorted = true;
  for (int64_t i = 0; i < n_pw; ++i) {
    if (mapping[i] < 0 || mapping[i] >= N) throw std::runtime_error("mapping index out of range");
    if (i && mapping[i] <= mapping[i - 1]) sorted = false;
  }
  if (!sorted)
    std::sort(order.begin(), order.end(), [&](int a, int b) { return mapping[a] < mapping[b]; });
  T.slot_ix.resize(n_pw);
  T.slot_src.resize(n_pw);
  T.colmap.assign(0, -1);
  int64_t prev_col = -1;
  std::vector<int> zplane_of_col;
  for (int64_t s = 0; s < n_pw; ++s) {
    int64_t lin = mapping[order[s]];
    if (s && lin == mapping[order[s - 1]]) throw std::runtime_error("duplicate mapping index");
    int ix = (int)(lin % nx);
    int64_t col = lin / nx;  // iy + ny*iz
    if (col != prev_col) {
      T.col_start.push_back((int)s);
      T.col_cnt.push_back(0);
      T.col_y.push_back((int)(col % ny));
      T.col_z.push_back((int)(col / ny));
      prev_col = col;
    }
    T.col_cnt.back()++;
    T.slot_ix[s] = ix;
    T.slot_src[s] = order[s];
  }
  T.n_cols = (int)T.col_start.size();
  T.cnt_max = 0;
  for (int c : T.col_cnt) T.cnt_max = std::max(T.cnt_max, c);
  // z planes (columns are sorted by (iz, iy))
  for (int c = 0; c < T.n_cols; ++c)
    if (T.zlist.empty() || T.zlist.back() != T.col_z[c]) T.zlist.push_back(T.col_z[c]);
  T.n_zc = (int)T.zlist.size();
  T.zc_of.assign(nz, -1);
  for (int i = 0; i < T.n_zc; ++i) T.zc_of[T.zlist[i]] = i;
  T.colmap.assign((size_t)T.n_zc * ny, -1);
  int izc = -1, lastz = -1;
  for (int c = 0; c < T.n_cols; ++c) {
    if (T.col_z[c] != lastz) {
      ++izc;
      lastz = T.col_z[c];
    }
    T.colmap[(size_t)izc * ny + T.col_y[c]] = c;
  }
  // range descriptors (validated against the tables)
  {
    auto two_ranges = [](const std::vector<int>& present, int& s0, int& n0, int& s1, int& n1) {
      // present[i] != 0 marks occupied indices; returns false if more than two runs
      const int n = (int)present.size();
      s0 = n0 = s1 = n1 = 0;
      int i = 0, runs = 0;
      while (i < n) {
        if (!present[i]) { ++i; continue; }
        int st = i;
        while (i < n && present[i]) ++i;
        if (runs == 0) { s0 = st; n0 = i - st; }
        else if (runs == 1) { s1 = st; n1 = i - st; }
        else return false;
        ++runs;
      }
      return true;
    };
    std::vector<int> zp(nz, 0);
    for (int z : T.zlist) zp[z] = 1;
    bool ok = two_ranges(zp, T.z_s0, T.z_n0, T.z_s1, T.z_n1);
    T.pl_s0.assign(T.n_zc, 0); T.pl_n0.assign(T.n_zc, 0); T.pl_s1.assign(T.n_zc, 0); T.pl_n1.assign(T.n_zc, 0);
    T.pl_col0.assign(T.n_zc, 0);
    std::vector<int> yp(ny);
    for (int p = 0; p < T.n_zc && ok; ++p) {
      const int* cm = &T.colmap[(size_t)p * ny];
      for (int iy = 0; iy < ny; ++iy) yp[iy] = cm[iy] >= 0;
      ok = two_ranges(yp, T.pl_s0[p], T.pl_n0[p], T.pl_s1[p], T.pl_n1[p]);
      if (ok && T.pl_n0[p] > 0) T.pl_col0[p] = cm[T.pl_s0[p]];
    }
    // x ranges of every column; the register engine also assumes slot_src is the identity (ascending mapping)
    T.cx_s0.assign(T.n_cols, 0); T.cx_n0.assign(T.n_cols, 0); T.cx_s1.assign(T.n_cols, 0); T.cx_n1.assign(T.n_cols, 0);
    ok = ok && sorted;
    std::vector<int> xp(nx);
    for (int c = 0; c < T.n_cols && ok; ++c) {
      std::fill(xp.begin(), xp.end(), 0);
      for (int i = 0; i < T.col_cnt[c]; ++i) xp[T.slot_ix[T.col_start[c] + i]] = 1;
      ok = two_ranges(xp, T.cx_s0[c], T.cx_n0[c], T.cx_s1[c], T.cx_n1[c]);
      ok = ok && (T.cx_n0[c] + T.cx_n1[c] == T.col_cnt[c]);
    }
    T.ranges_ok = ok ? 1 : 0;
  }
  return T;
}

}  // namespace dftk
