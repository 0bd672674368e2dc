// Tensor-core kernel of the INT8-residue GEMM emulation (see i8emu_core.cuh): for one modulus, one K chunk (<= 2^16 terms)
// and one 128 x 128 output tile it forms the four integer products
//     X1 = Ar^T Br,  X2 = Ai^T Bi,  X3 = Ar^T Bi,  X4 = Ai^T Br          (s8 x s8 -> s32)
// with `tcgen05.mma.cta_group::1.kind::i8` into four TMEM accumulators (4 x 128 columns = all of TMEM) and writes the
// residues  (X1 + X2) mod p,  (X3 - X4) mod p  of  conj(a) . b.
//
//   warps 0-3: producers (cp.async 16 B pieces of the K-major residue planes into the canonical no-swizzle UMMA layout:
//              8-row x 16-byte core matrices, K chunk stride = LBO, 8-row stride = SBO), later the epilogue
//              (tcgen05.ld: one TMEM lane = one output row per thread);
//   warp 4:    one elected lane issues the MMAs; tcgen05.commit releases shared-memory stages / signals the epilogue.
//
// STATUS: compiled for sm_100a only (UTCIMMA / LDTM / UTCBAR in the SASS); it has NOT run on hardware -- no GPU minutes were
// left when it was written.  It is reachable only through option gemm_backend = 3 and is checked against the CUDA-core
// reference pipeline (gemm_backend = 2) by a GPU test that is committed but skipped.  scripts/tcgen05_i8_probe.cu settles
// the descriptor convention (I8TC_SWAP_LBO_SBO) on the first GPU call of the next round.
#include <algorithm>
#include <string>
#include <vector>
#include "structs.cuh"
#include "i8emu_core.cuh"

namespace dftk {

#ifndef I8TC_SWAP_LBO_SBO
#define I8TC_SWAP_LBO_SBO 0
#endif

constexpr int TC_M = 128;        // output rows per tile (TMEM lanes)
constexpr int TC_N = 128;        // output columns per tile (4 accumulators x 128 columns = 512 TMEM columns)
constexpr int TC_BK = 128;       // K bytes per stage (8 core-matrix columns)
constexpr int TC_STAGES = 3;
constexpr int TC_TILE_BYTES = TC_M * TC_BK;                      // one operand tile (128 rows x 128 B)
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;                // Ar, Ai, Br, Bi
constexpr int TC_SMEM = TC_STAGES * TC_STAGE_BYTES + 1024;       // + alignment slack
constexpr int TC_THREADS = 160;

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(tc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" :: "r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
  // bounded wait (bring-up): a protocol error traps after ~2 s instead of hanging the GPU
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(ok) : "r"(tc_smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // cute::UMMA::SmemDescriptor: start address, leading / stride byte offsets in 16-byte units, version 1, no swizzle
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
      :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

// planes: [(2 t + part)][row][ldk] int8, ldk a multiple of TC_BK, zero padded.  grid (m tiles, n tiles, n_mod * n_chunks).
// part[(chunk)][(2 t + part)][j][i] int16: residues of this chunk (summed modulo p by k_i8_sum_chunks)
__global__ void __launch_bounds__(TC_THREADS, 1)
k_i8_gemm_tc(const signed char* __restrict__ ra, const signed char* __restrict__ rb, int64_t m, int64_t n, int64_t ldk,
             int n_mod, int n_chunks, int64_t chunk_len, short* __restrict__ part, int swap_lbo_sbo, int simple,
             int* __restrict__ dbg) {
  extern __shared__ unsigned char tc_raw[];
  unsigned char* sm = (unsigned char*)(((uintptr_t)tc_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[TC_STAGES], empty_bar[TC_STAGES], accum_bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t = blockIdx.z % n_mod, chunk = blockIdx.z / n_mod;
  const int p = i8_modulus(t);
  const int64_t i0 = (int64_t)blockIdx.x * TC_M, j0 = (int64_t)blockIdx.y * TC_N;
  const int64_t k_begin = (int64_t)chunk * chunk_len;
  const int64_t k_end = k_begin + chunk_len < ldk ? k_begin + chunk_len : ldk;
  const int n_iters = (int)((k_end - k_begin) / TC_BK);           // chunk_len and ldk are multiples of TC_BK

  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      tc_mbar_init(&full_bar[s], 128);       // every producer thread arrives once per stage
      tc_mbar_init(&empty_bar[s], 1);        // one tcgen05.commit
    }
    tc_mbar_init(&accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tc_smem_u32(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;

  if (warp < 4) {
    // ---------------- producers: stage s holds [Ar | Ai | Br | Bi], each TC_M rows x TC_BK bytes, canonical layout
    //                  byte offset of (row r, k byte c):  (c / 16) * (rows * 16) + (r / 8) * 128 + (r % 8) * 16 + c % 16
    const signed char* src[4];
    src[0] = ra + ((size_t)(2 * t) * m) * ldk;
    src[1] = ra + ((size_t)(2 * t + 1) * m) * ldk;
    src[2] = rb + ((size_t)(2 * t) * n) * ldk;
    src[3] = rb + ((size_t)(2 * t + 1) * n) * ldk;
    if (simple) {
      // bring-up mode: one stage, plain 16-byte loads/stores, a full CTA barrier per K block (no cp.async, no overlap)
      for (int it = 0; it < n_iters; ++it) {
        if (it > 0) tc_mbar_wait(&empty_bar[0], (uint32_t)((it - 1) & 1));
        const int64_t kb = k_begin + (int64_t)it * TC_BK;
        for (int q = tid; q < 4096; q += 128) {
          const int tile = q >> 10, r = (q >> 3) & 127, c16 = q & 7;
          const int64_t rows = tile < 2 ? m : n;
          int64_t row = (tile < 2 ? i0 : j0) + r;
          if (row >= rows) row = rows - 1;
          const uint4 v = *reinterpret_cast<const uint4*>(src[tile] + (size_t)row * ldk + kb + c16 * 16);
          *reinterpret_cast<uint4*>(sm + tile * TC_TILE_BYTES + c16 * (TC_M * 16) + (r >> 3) * 128 + (r & 7) * 16) = v;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_mbar_arrive(&full_bar[0]);
      }
    } else
    for (int it = 0; it < n_iters + TC_STAGES - 1; ++it) {
      if (it < n_iters) {
        const int s = it % TC_STAGES;
        if (it >= TC_STAGES) tc_mbar_wait(&empty_bar[s], (uint32_t)((it / TC_STAGES - 1) & 1));
        const int64_t kb = k_begin + (int64_t)it * TC_BK;
        unsigned char* stage = sm + (size_t)s * TC_STAGE_BYTES;
        // 4 tiles x 128 rows x 8 pieces of 16 B = 4096 pieces, 32 per thread; consecutive threads take consecutive pieces
        // of one row (128 B contiguous in global memory)
#pragma unroll 4
        for (int q = tid; q < 4096; q += 128) {
          const int tile = q >> 10, r = (q >> 3) & 127, c16 = q & 7;
          const int64_t rows = tile < 2 ? m : n;
          int64_t row = (tile < 2 ? i0 : j0) + r;
          if (row >= rows) row = rows - 1;                          // overhang rows are computed but never stored
          const signed char* g = src[tile] + (size_t)row * ldk + kb + c16 * 16;
          const uint32_t dst = tc_smem_u32(stage + tile * TC_TILE_BYTES + c16 * (TC_M * 16) + (r >> 3) * 128 + (r & 7) * 16);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(g) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (it >= TC_STAGES - 1) {
        // the group committed TC_STAGES - 1 iterations ago has landed: publish it to the async proxy and to the MMA warp
        asm volatile("cp.async.wait_group %0;" :: "n"(TC_STAGES - 1) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_mbar_arrive(&full_bar[(it - (TC_STAGES - 1)) % TC_STAGES]);
      }
    }
    // ---------------- epilogue: thread = TMEM lane = output row
    tc_mbar_wait(&accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int64_t i = i0 + tid;
    short* out_re = part + (((size_t)chunk * 2 * n_mod + 2 * t) * n) * m;
    short* out_im = part + (((size_t)chunk * 2 * n_mod + 2 * t + 1) * n) * m;
    for (int c0 = 0; c0 < TC_N; c0 += 32) {
      uint32_t x1[32], x2[32], x3[32], x4[32];
      const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      tc_ld32(lane_base + 0 * TC_N, x1);
      tc_ld32(lane_base + 1 * TC_N, x2);
      tc_ld32(lane_base + 2 * TC_N, x3);
      tc_ld32(lane_base + 3 * TC_N, x4);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
        // bring-up dump: dbg[x][row][col] raw s32 accumulators of the first tile (x = 0..3 for X1..X4)
        for (int c = 0; c < 32; ++c) {
          dbg[(0 * TC_M + tid) * TC_N + c0 + c] = (int)x1[c];
          dbg[(1 * TC_M + tid) * TC_N + c0 + c] = (int)x2[c];
          dbg[(2 * TC_M + tid) * TC_N + c0 + c] = (int)x3[c];
          dbg[(3 * TC_M + tid) * TC_N + c0 + c] = (int)x4[c];
        }
      }
      if (i < m) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int64_t j = j0 + c0 + c;
          if (j < n) {
            const int re = ((int)x1[c] % p + (int)x2[c] % p) % p;
            const int im = ((int)x3[c] % p - (int)x4[c] % p) % p;
            out_re[(size_t)j * m + i] = (short)re;
            out_im[(size_t)j * m + i] = (short)im;
          }
        }
      }
    }
    (void)lane;
  } else {
    // ---------------- MMA issuer
    // instruction descriptor: D = s32, A = B = s8, both K-major, N = 128, M = 128  (cute::UMMA::InstrDescriptor)
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
    const bool swp = (I8TC_SWAP_LBO_SBO != 0) != (swap_lbo_sbo != 0);
    const uint32_t lbo = swp ? 128u : (uint32_t)(TC_M * 16);
    const uint32_t sbo = swp ? (uint32_t)(TC_M * 16) : 128u;
    for (int it = 0; it < n_iters; ++it) {
      const int s = simple ? 0 : it % TC_STAGES;
      tc_mbar_wait(&full_bar[s], (uint32_t)((simple ? it : it / TC_STAGES) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t base = tc_smem_u32(sm + (size_t)s * TC_STAGE_BYTES);
#pragma unroll
        for (int kk = 0; kk < TC_BK / 32; ++kk) {
          const uint32_t koff = (uint32_t)kk * 2u * (uint32_t)(TC_M * 16);       // two 16-byte K chunks per MMA
          const uint64_t dAr = tc_desc(base + 0 * TC_TILE_BYTES + koff, lbo, sbo);
          const uint64_t dAi = tc_desc(base + 1 * TC_TILE_BYTES + koff, lbo, sbo);
          const uint64_t dBr = tc_desc(base + 2 * TC_TILE_BYTES + koff, lbo, sbo);
          const uint64_t dBi = tc_desc(base + 3 * TC_TILE_BYTES + koff, lbo, sbo);
          const uint32_t acc = (uint32_t)(it > 0 || kk > 0);
          tc_mma_i8(tmem + 0 * TC_N, dAr, dBr, idesc, acc);
          tc_mma_i8(tmem + 1 * TC_N, dAi, dBi, idesc, acc);
          tc_mma_i8(tmem + 2 * TC_N, dAr, dBi, idesc, acc);
          tc_mma_i8(tmem + 3 * TC_N, dAi, dBr, idesc, acc);
        }
        tc_commit(&empty_bar[s]);                    // frees the stage once these MMAs have read it
        if (it == n_iters - 1) tc_commit(&accum_bar);
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

// res[(2 t + part)][j][i] = sum over chunks (mod p), symmetric representative
__global__ void k_i8_sum_chunks(const short* __restrict__ part, int n_chunks, int n_mod, int64_t mn, int* __restrict__ res) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * (int64_t)n_mod * mn) return;
  const int t = (int)(idx / (2 * mn));
  const int p = i8_modulus(t);
  int s = 0;
  for (int c = 0; c < n_chunks; ++c) s = (s + part[(size_t)c * 2 * n_mod * mn + idx]) % p;
  res[idx] = i8_sym(s, p);
}

void i8tc_set_attributes() {
  // experimental kernel: never let a failure here leak into the error state of the product path
  if (cudaFuncSetAttribute(k_i8_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM) != cudaSuccess) cudaGetLastError();
}

// integer stage of C = A^H B on the tensor cores; ra / rb: padded residue planes, res: int residues [(2 t + part)][j][i]
void i8tc_products(dftk_b200_ctx* ctx, const signed char* ra, const signed char* rb, int64_t m, int64_t n, int64_t ldk,
                   int n_mod, short* part, int* res) {
  const int64_t chunk_len = I8_K_CHUNK;                         // multiple of TC_BK
  const int n_chunks = (int)((ldk + chunk_len - 1) / chunk_len);
  dim3 grid((unsigned)((m + TC_M - 1) / TC_M), (unsigned)((n + TC_N - 1) / TC_N), (unsigned)(n_mod * n_chunks));
  // bring-up switches (see scripts/next_round_first_call.sh): descriptor convention and the unpipelined variant
  const char* e1 = getenv("DFTK_B200_I8TC_SWAP");
  const char* e2 = getenv("DFTK_B200_I8TC_SIMPLE");
  // DFTK_B200_I8TC_DUMP=<prefix>: write the raw accumulators of the first tile (modulus 0, chunk 0) and the operand rows
  // that produced them to <prefix>.x / .a / .b for scripts/i8tc_debug.py
  const char* e3 = getenv("DFTK_B200_I8TC_DUMP");
  int* dbg = nullptr;
  if (e3 && e3[0]) {
    CUDA_CHECK(cudaMalloc((void**)&dbg, (size_t)4 * TC_M * TC_N * sizeof(int)));
    CUDA_CHECK(cudaMemsetAsync(dbg, 0x7f, (size_t)4 * TC_M * TC_N * sizeof(int), ctx->stream));
  }
  LAUNCH(ctx, k_i8_gemm_tc, grid, TC_THREADS, TC_SMEM, ra, rb, m, n, ldk, n_mod, n_chunks, chunk_len, part,
         e1 && e1[0] == '1' ? 1 : 0, e2 && e2[0] == '1' ? 1 : 0, dbg);
  if (dbg) {
    cudaError_t err = cudaStreamSynchronize(ctx->stream);
    const int64_t kc = std::min<int64_t>(chunk_len, ldk);
    std::vector<int> hx((size_t)4 * TC_M * TC_N);
    auto dump = [&](const char* ext, const void* host, size_t bytes) {
      std::string path = std::string(e3) + ext;
      if (FILE* f = fopen(path.c_str(), "wb")) {
        fwrite(host, 1, bytes, f);
        fclose(f);
      }
    };
    if (err == cudaSuccess) {
      cudaMemcpy(hx.data(), dbg, hx.size() * sizeof(int), cudaMemcpyDeviceToHost);
      dump(".x", hx.data(), hx.size() * sizeof(int));
      // operand rows of the first tile: [part][row][k] for modulus 0, first chunk (rows beyond m / n clamp like the kernel)
      for (int which = 0; which < 2; ++which) {
        const int64_t rows = which == 0 ? m : n;
        const signed char* src = which == 0 ? ra : rb;
        std::vector<signed char> h((size_t)2 * TC_M * kc);
        for (int part_i = 0; part_i < 2; ++part_i)
          for (int r = 0; r < TC_M; ++r) {
            const int64_t row = std::min<int64_t>(r, rows - 1);
            cudaMemcpy(h.data() + ((size_t)part_i * TC_M + r) * kc, src + ((size_t)part_i * rows + row) * ldk, (size_t)kc,
                       cudaMemcpyDeviceToHost);
          }
        dump(which == 0 ? ".a" : ".b", h.data(), h.size());
      }
      fprintf(stderr, "[i8tc] dumped first-tile accumulators and operands to %s.{x,a,b} (k = %lld)\n", e3, (long long)kc);
    } else {
      fprintf(stderr, "[i8tc] kernel failed before the dump: %s\n", cudaGetErrorString(err));
    }
    cudaFree(dbg);
    CUDA_CHECK(err);
  }
  const int64_t tot = 2 * (int64_t)n_mod * m * n;
  LAUNCH(ctx, k_i8_sum_chunks, (unsigned)((tot + 255) / 256), 256, 0, (const short*)part, n_chunks, n_mod, m * n, res);
}

}  // namespace dftk
