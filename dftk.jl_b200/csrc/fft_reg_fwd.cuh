// SphereTablesX: device view of the pruning tables incl. the inverse z-plane map (register engine).
#pragma once
#include "fft_core.cuh"
namespace dftk {
struct SphereTablesX : SphereTables {
  const int* zc_of;  // [nz] plane index of wrapped z, or -1
};
}  // namespace dftk
