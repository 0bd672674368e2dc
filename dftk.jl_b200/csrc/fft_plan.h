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

struct SphereTablesHost {
  int nx, ny, nz;
  int64_t n_pw;
  int n_cols, cnt_max, n_zc;
  std::vector<int> col_start, col_cnt, slot_ix, slot_src, zlist, colmap, col_y, col_z;
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
  T.colmap.assign((size_t)T.n_zc * ny, -1);
  int izc = -1, lastz = -1;
  for (int c = 0; c < T.n_cols; ++c) {
    if (T.col_z[c] != lastz) {
      ++izc;
      lastz = T.col_z[c];
    }
    T.colmap[(size_t)izc * ny + T.col_y[c]] = c;
  }
  return T;
}

}  // namespace dftk
