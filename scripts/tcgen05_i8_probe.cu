// Bring-up probe for the INT8 tensor-core path (`tcgen05.mma.kind::i8`, s8 x s8 -> s32 in TMEM) that the FP64-by-integer
// GEMM emulation of dftk.jl_b200/csrc/i8emu_core.cuh targets.  NOT part of the library: a standalone program that
//   (1) checks one 128 x N x 32 MMA, operands in the canonical K-major no-swizzle shared-memory layout, against a CPU
//       reference for a handful of descriptor conventions (so a wrong guess about LBO/SBO shows up in the first GPU call),
//   (2) measures the issue rate of back-to-back MMAs on resident operands (the per-SM INT8 ceiling).
// Build / run on the B200 box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/i8probe scripts/tcgen05_i8_probe.cu && /tmp/i8probe
// No GPU was available when this was written; it has only been compiled (see DESIGN.md "what comes next").
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, LBO, SBO in 16-byte units, version 1,
// layout type 0 = no swizzle ("interleave")
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, int version) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)(version & 3) << 46;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor) for kind::i8: c = s32 (2), a = b = signed 8 bit (1), K-major
__host__ __device__ inline uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 2u << 4;                       // c_format = S32
  d |= 1u << 7;                       // a_format = INT8
  d |= 1u << 10;                      // b_format = INT8
  d |= (uint32_t)(a_mn_major & 1) << 15;
  d |= (uint32_t)(b_mn_major & 1) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
      :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();                 // bounded: a wrong protocol traps after ~2 s instead of hanging the GPU
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// byte offset of element (row r, k) of an R x 32-byte K-major tile made of 8 x 16 B core matrices:
//   core matrix (r / 8, k / 16) at  (k / 16) * lbo + (r / 8) * sbo,  row r % 8 at +16 B each
__host__ __device__ inline uint32_t tile_off(int r, int k, uint32_t lbo, uint32_t sbo) {
  return (uint32_t)(k / 16) * lbo + (uint32_t)(r / 8) * sbo + (uint32_t)(r % 8) * 16 + (uint32_t)(k % 16);
}

constexpr int M_ = 128, K_ = 32;

// variant bits: 1 = swap the LBO / SBO fields of the descriptors, 2 = descriptor version 0
template <int N_>
__global__ void __launch_bounds__(128) probe_mma(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ D,
                                                  int variant) {
  __shared__ __align__(1024) int8_t sA[M_ * K_];
  __shared__ __align__(1024) int8_t sB[N_ * K_];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t sboA = 128, lboA = (M_ / 8) * 128, sboB = 128, lboB = (N_ / 8) * 128;
  for (int e = tid; e < M_ * K_; e += 128) sA[tile_off(e / K_, e % K_, lboA, sboA)] = A[e];      // A: row-major M x K
  for (int e = tid; e < N_ * K_; e += 128) sB[tile_off(e / K_, e % K_, lboB, sboB)] = B[e];      // B: row-major N x K
  if (tid == 0) mbar_init(&bar, 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(N_ < 32 ? 32 : N_) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> async proxy (UMMA reads)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const bool swap = variant & 1;
    const int ver = (variant & 2) ? 0 : 1;
    const uint64_t da = make_desc(smem_u32(sA), swap ? sboA : lboA, swap ? lboA : sboA, ver);
    const uint64_t db = make_desc(smem_u32(sB), swap ? sboB : lboB, swap ? lboB : sboB, ver);
    mma_i8(tmem, da, db, make_idesc(M_, N_, 0, 0), 0u);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // warp w reads TMEM lanes 32 w .. 32 w + 31 (= accumulator rows), 32 columns at a time
  for (int c0 = 0; c0 < N_; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) D[(size_t)tid * N_ + c0 + j] = (int32_t)v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(N_ < 32 ? 32 : N_) : "memory");
}

// A in MN-major form (the update-type products: output rows contiguous in the stored residue planes).  Canonical no-swizzle
// MN-major layout (cute: ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO)), T = 16 bytes): core matrix = 8 K-rows of 16 M-contiguous bytes;
// element (m, k) at (m / 16) * sbo + (k / 8) * lbo + (k % 8) * 16 + m % 16.   variant bit 1 swaps the two descriptor fields.
__global__ void __launch_bounds__(128) probe_mma_mn(const int8_t* __restrict__ A_km, const int8_t* __restrict__ B, int32_t* __restrict__ D,
                                                   int variant) {
  constexpr int N_ = 64;
  __shared__ __align__(1024) int8_t sA[M_ * K_];
  __shared__ __align__(1024) int8_t sB[N_ * K_];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t lboA = 128, sboA = (K_ / 8) * 128;                 // core matrices ordered [m / 16][k / 8]
  const uint32_t sboB = 128, lboB = (N_ / 8) * 128;
  for (int e = tid; e < M_ * K_; e += 128) {                        // A_km: K x M, M contiguous
    const int k = e / M_, m = e % M_;
    sA[(m / 16) * sboA + (k / 8) * lboA + (k % 8) * 16 + m % 16] = A_km[e];
  }
  for (int e = tid; e < N_ * K_; e += 128) sB[tile_off(e / K_, e % K_, lboB, sboB)] = B[e];
  if (tid == 0) mbar_init(&bar, 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(N_) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const bool swap = variant & 1;
    const uint64_t da = make_desc(smem_u32(sA), swap ? sboA : lboA, swap ? lboA : sboA, 1);
    const uint64_t db = make_desc(smem_u32(sB), lboB, sboB, 1);      // use the K-major convention that passed (edit if variant 1 won)
    mma_i8(tmem, da, db, make_idesc(M_, N_, 1, 0), 0u);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < N_; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) D[(size_t)tid * N_ + c0 + j] = (int32_t)v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(N_) : "memory");
}

static int run_mn(int variant) {
  constexpr int N_ = 64;
  std::vector<int8_t> A(M_ * K_), B(N_ * K_);
  srand(99);
  for (auto& x : A) x = (int8_t)(rand() % 256 - 128);              // A_km[k * M + m]
  for (auto& x : B) x = (int8_t)(rand() % 256 - 128);
  std::vector<int32_t> ref(M_ * N_), got(M_ * N_);
  for (int i = 0; i < M_; ++i)
    for (int j = 0; j < N_; ++j) {
      int s = 0;
      for (int k = 0; k < K_; ++k) s += (int)A[k * M_ + i] * (int)B[j * K_ + k];
      ref[i * N_ + j] = s;
    }
  int8_t *dA, *dB;
  int32_t* dD;
  CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dD, got.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0x7f, got.size() * 4));
  probe_mma_mn<<<1, 128>>>(dA, dB, dD, variant);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  MN-major A, variant %d: kernel failed: %s\n", variant, cudaGetErrorString(e)); return -1; }
  CK(cudaMemcpy(got.data(), dD, got.size() * 4, cudaMemcpyDeviceToHost));
  long bad = 0;
  for (size_t i = 0; i < ref.size(); ++i) bad += ref[i] != got[i];
  printf("  MN-major A, variant %d (%s): %ld of %zu entries differ%s\n", variant,
         (variant & 1) ? "fields swapped" : "LBO = 8-K-row stride, SBO = 16-M-byte block stride", bad, ref.size(), bad ? "" : "   <-- MATCH");
  return bad == 0;
}

// issue-rate probe: `iters` x 4 MMAs of 128 x 256 x 32 per CTA on resident (arbitrary) operands, one CTA per SM
__global__ void __launch_bounds__(128) probe_rate(int iters, long long* cycles) {
  extern __shared__ __align__(1024) int8_t sm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int N = 256;
  int8_t* sA = sm;
  int8_t* sB = sm + 4 * M_ * K_;
  for (int e = tid; e < 4 * (M_ + N) * K_; e += 128) sm[e] = (int8_t)(e * 7 + 3);
  if (tid == 0) mbar_init(&bar, 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  long long t0 = clock64();
  if (tid == 0) {
    const uint32_t idesc = make_idesc(M_, N, 0, 0);
    for (int it = 0; it < iters; ++it)
      for (int s = 0; s < 4; ++s) {
        const uint64_t da = make_desc(smem_u32(sA + s * M_ * K_), (M_ / 8) * 128, 128, 1);
        const uint64_t db = make_desc(smem_u32(sB + s * N * K_), (N / 8) * 128, 128, 1);
        mma_i8(tmem + (uint32_t)((it & 1) * N), da, db, idesc, (uint32_t)(it > 1 || s > 0));
      }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

template <int N_>
static int run_variant(int variant) {
  std::vector<int8_t> A(M_ * K_), B(N_ * K_);
  srand(1234);
  for (auto& x : A) x = (int8_t)(rand() % 256 - 128);
  for (auto& x : B) x = (int8_t)(rand() % 256 - 128);
  std::vector<int32_t> ref(M_ * N_), got(M_ * N_, 0x7fffffff);
  for (int i = 0; i < M_; ++i)
    for (int j = 0; j < N_; ++j) {
      int s = 0;
      for (int k = 0; k < K_; ++k) s += (int)A[i * K_ + k] * (int)B[j * K_ + k];
      ref[i * N_ + j] = s;
    }
  int8_t *dA, *dB;
  int32_t* dD;
  CK(cudaMalloc(&dA, A.size()));
  CK(cudaMalloc(&dB, B.size()));
  CK(cudaMalloc(&dD, got.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0x7f, got.size() * 4));
  probe_mma<N_><<<1, 128>>>(dA, dB, dD, variant);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("  N=%3d variant %d: kernel failed: %s\n", N_, variant, cudaGetErrorString(e));
    return -1;
  }
  CK(cudaMemcpy(got.data(), dD, got.size() * 4, cudaMemcpyDeviceToHost));
  long bad = 0;
  for (size_t i = 0; i < ref.size(); ++i) bad += ref[i] != got[i];
  printf("  N=%3d variant %d (%s LBO/SBO, descriptor version %d): %ld of %zu entries differ%s\n", N_, variant,
         (variant & 1) ? "swapped" : "K-major: LBO = K-chunk stride, SBO = 8-row stride", (variant & 2) ? 0 : 1, bad, ref.size(),
         bad ? "" : "   <-- MATCH");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return bad == 0;
}

int main(int argc, char** argv) {
  // usage: i8probe <variant 0..3>   one descriptor convention (separate processes: a faulting kernel poisons the context)
  //        i8probe rate             issue-rate measurement
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  if (prop.major < 10) { printf("needs sm_100a\n"); return 1; }
  if (argc < 2) { printf("usage: %s <0|1|2|3|mn0|mn1|rate>\n", argv[0]); return 1; }
  if (argv[1][0] == 'm') return run_mn(argv[1][2] - '0') > 0 ? 0 : 3;
  if (argv[1][0] != 'r') {
    const int v = atoi(argv[1]);
    printf("(1) single tcgen05.mma.kind::i8 128 x N x 32 against the CPU, descriptor variant %d:\n", v);
    const int r = run_variant<64>(v);
    if (r > 0) run_variant<256>(v);
    return r > 0 ? 0 : 3;
  }
  printf("(2) issue rate, 128 x 256 x 32 MMAs on resident operands, one CTA per SM:\n");
  const int iters = 4096;
  long long* dc;
  CK(cudaMalloc(&dc, prop.multiProcessorCount * sizeof(long long)));
  const size_t smem = 4 * (M_ + 256) * K_;
  CK(cudaFuncSetAttribute(probe_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  probe_rate<<<prop.multiProcessorCount, 128, smem>>>(16, dc);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  probe_rate<<<prop.multiProcessorCount, 128, smem>>>(iters, dc);
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  const double ops = 2.0 * M_ * 256 * K_ * 4.0 * iters * prop.multiProcessorCount;
  printf("  %d x 4 MMAs per SM in %.3f ms -> %.2f POPS (dense INT8 nominal 4.5)\n", iters, ms, ops / (ms * 1e-3) / 1e15);
  return 0;
}
