// SphereTablesX: device view of the pruning tables incl. the inverse z-plane map (register engine).
#pragma once
#include "fft_core.cuh"
namespace dftk {
struct SphereTablesX : SphereTables {
  const int* zc_of;  // [nz] plane index of wrapped z, or -1
  // range form (see SphereTablesHost); ranges_ok == 0 => fall back to zc_of / colmap lookups
  int ranges_ok, z_s0, z_n0, z_s1, z_n1;
  const int *pl_s0, *pl_n0, *pl_s1, *pl_n1, *pl_col0;
};
// plane index of wrapped z (or -1): pure arithmetic when the range form holds (no dependent global load)
HD int zc_index(const SphereTablesX& T, int iz) {
  if (T.ranges_ok) {
    unsigned a = (unsigned)(iz - T.z_s0), b = (unsigned)(iz - T.z_s1);
    return a < (unsigned)T.z_n0 ? (int)a : (b < (unsigned)T.z_n1 ? T.z_n0 + (int)b : -1);
  }
  return T.zc_of[iz];
}
struct PlaneCols {
  int s0, n0, s1, n1, col0;
  const int* cm;  // table fallback
  HD int col(int iy) const {
    if (cm) return cm[iy];
    unsigned a = (unsigned)(iy - s0), b = (unsigned)(iy - s1);
    return a < (unsigned)n0 ? col0 + (int)a : (b < (unsigned)n1 ? col0 + n0 + (int)b : -1);
  }
};
HD PlaneCols plane_cols(const SphereTablesX& T, int izc) {
  PlaneCols p;
  if (T.ranges_ok) {
    p.s0 = T.pl_s0[izc]; p.n0 = T.pl_n0[izc]; p.s1 = T.pl_s1[izc]; p.n1 = T.pl_n1[izc]; p.col0 = T.pl_col0[izc];
    p.cm = nullptr;
  } else {
    p.s0 = p.n0 = p.s1 = p.n1 = p.col0 = 0;
    p.cm = T.colmap + (size_t)izc * T.ny;
  }
  return p;
}
}  // namespace dftk
