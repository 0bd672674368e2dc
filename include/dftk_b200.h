/*
 * libdftk_b200 -- C ABI of the B200-native plane-wave Kohn-Sham hot path.
 *
 * Drop-in boundary for the seam where DFTK.jl's ext/DFTKCUDAExt.jl + src/architecture.jl plug in
 * today (reference paths relative to the DFTK.jl tree).  Each entry point names the reference
 * function it replaces.  Conventions:
 *   - all functions return 0 on success, a negative DFTK_B200_E* code on failure; the message of the
 *     last failure on a context is available from dftk_b200_last_error(); nothing throws.
 *   - column-major arrays, complex = interleaved double[2], indices 0-based (Julia glue subtracts 1
 *     once when it passes `kpt.mapping`).
 *   - "dev/host" pointers may be device or host memory (resolved through UVA); hot-path buffers
 *     (psi, hpsi, X, rho) are expected on the device -- host buffers are staged through H2D/D2H copies
 *     inside the call (this is what the end-to-end benchmark measures).
 *   - handles are opaque and not thread-safe; one context per GPU / rank; the caller owns psi/rho/V
 *     buffers, the library owns plans, scratch and its copies of mapping/kin/P/D.
 */
#ifndef DFTK_B200_H
#define DFTK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dftk_b200_ctx dftk_b200_ctx;
typedef struct dftk_b200_grid dftk_b200_grid;
typedef struct dftk_b200_kblock dftk_b200_kblock;

#define DFTK_B200_OK 0
#define DFTK_B200_EINVAL (-1)   /* bad argument */
#define DFTK_B200_ECUDA (-2)    /* CUDA runtime / library failure */
#define DFTK_B200_ENUM (-3)     /* numerical failure (e.g. LOBPCG cannot keep vectors normalised) */
#define DFTK_B200_ENCCL (-4)    /* NCCL failure / communicator missing */

#define DFTK_B200_F64 0
#define DFTK_B200_I64 1

/* ---- context (replaces the `architecture=GPU(CuArray)` selection, src/architecture.jl:4-53;
 *      synchronize_device / memory_usage, ext/DFTKCUDAExt.jl:11-15) ---- */
int dftk_b200_ctx_create(int device, dftk_b200_ctx** out);
/* distributed variant: `nccl_unique_id` = 128 bytes from dftk_b200_nccl_unique_id on rank 0 */
int dftk_b200_ctx_create_dist(int device, const void* nccl_unique_id, int rank, int nranks,
                              dftk_b200_ctx** out);
int dftk_b200_nccl_unique_id(void* out128);
int dftk_b200_ctx_destroy(dftk_b200_ctx* ctx);
const char* dftk_b200_last_error(dftk_b200_ctx* ctx); /* ctx may be NULL: last global error */
/* Every kernel, copy and library call of this context is enqueued on ONE CUDA stream: the legacy default stream after
 * creation (ordered with the caller's default-stream work, which is what DFTK's GPU path uses), or the `cudaStream_t`
 * given here (e.g. CUDA.jl's task-local stream, `CUDA.stream().handle`).  Handles created from the context follow. */
int dftk_b200_ctx_set_stream(dftk_b200_ctx* ctx, void* cuda_stream);
int dftk_b200_sync(dftk_b200_ctx* ctx);
int dftk_b200_mem_info(dftk_b200_ctx* ctx, int64_t* free_bytes, int64_t* total_bytes);
/* number of kernel launches issued by this library on the context since creation / last reset */
int64_t dftk_b200_launch_count(dftk_b200_ctx* ctx, int reset);
/* number of scheduler rounds (= host synchronisations) of the batched LOBPCG solves since creation / last reset */
int64_t dftk_b200_sync_count(dftk_b200_ctx* ctx, int reset);
/* FP64-equivalent GEMM flops the large-path LOBPCG solves of this context have executed since the last reset: Gram-type products
 * (8 rows·m·n, half of it for the upper-triangle-only diagonal blocks), update-type products, the X R^-1 updates of ortho! and the
 * two projector products of every H apply -- the "sum actually executed" of SURVEY §8d.  Batched small-block solves are not counted. */
double dftk_b200_lobpcg_flops(dftk_b200_ctx* ctx, int reset);
/* tuning knobs: "gemm_backend" (4 = default: the Gram-type and update-type products of contractions with at least
 * "i8_min_rows" (32768) rows run on the INT8 tensor cores -- tcgen05.mma.kind::i8 fed by TMA, FP64-equivalent results through
 * INT8 residues + CRT -- and everything smaller on the own FP64 DMMA kernels; 0 = DMMA kernels only; 1 = cuBLAS, for A/B
 * comparison and peak calibration only; 2 / 3 = checkers of the INT8 scheme: integer products on CUDA cores / cp.async-fed
 * tensor-core kernel), "gemm_stages" (cp.async ring depth 2|3 of the DMMA kernels), "band_chunk" (bands per batched-FFT
 * launch, 0 = auto), "fft_engine" (0 = register two-pass engine where a factor pair exists, 1 = generic Stockham; applies to
 * grids created afterwards), "small_dense" (1 = batched small-matrix path for LOBPCG solves with <= 32 bands, 0 = the GEMM +
 * cuSOLVER sequence of the large path), "z_pipeline" (1 = persistent cp.async-pipelined fused z stage; default 0),
 * "batch_pipeline" (1 = batched solves of >= 8 k-blocks run as two groups on two streams so that one group's host work hides
 * behind the other's kernels; 0 = one group, one stream synchronisation per round; default 0: measured no faster),
 * "force_svd_fallback" (test hook) */
int dftk_b200_set_option(dftk_b200_ctx* ctx, const char* name, int64_t value);

/* ---- FFT grid (FFTGrid + build_fft_plans!, src/fft.jl:57-98,343-362) ---- */
int dftk_b200_grid_create(dftk_b200_ctx* ctx, int nx, int ny, int nz, double unit_cell_volume,
                          dftk_b200_grid** out);
int dftk_b200_grid_destroy(dftk_b200_grid* grid);
/* in-place unnormalised 3D C2C transform of `batch` cubes; direction -1 = forward (e^{-iGr}),
 * +1 = backward (ipFFT / ipBFFT of src/fft.jl:107,119,159,166) */
int dftk_b200_fft_cube(dftk_b200_grid* grid, void* data /*dev: complex[N*batch]*/, int direction,
                       int64_t batch);

/* ---- k-block = Kpoint + DftHamiltonianBlock data (src/Kpoint.jl:6-41,
 *      src/terms/Hamiltonian.jl:22-57, kinetic.jl:24-35, nonlocal.jl:9-28) ----
 * mapping: n_pw int64, 0-based linear cube index of each sphere coefficient (kpt.mapping - 1)
 * kin:     n_pw doubles, ½|k+G|² (FourierMultiplication multiplier); may be NULL (no kinetic term)
 * P:       n_pw × n_proj complex, column-major (NonlocalOperator.P); n_proj may be 0
 * D:       n_proj × n_proj real, column-major (NonlocalOperator.D, block diagonal) */
int dftk_b200_kblock_create(dftk_b200_grid* grid, int64_t n_pw, const int64_t* mapping,
                            const double* kin, int64_t n_proj, const void* P, const double* D,
                            int spin, double kweight, dftk_b200_kblock** out);
int dftk_b200_kblock_destroy(dftk_b200_kblock* kb);
/* total local potential on the real-space grid for this block's spin (sum of all
 * RealSpaceMultiplication operators, src/terms/operators.jl:213-222); N_fft doubles.  NULL = none */
int dftk_b200_kblock_set_potential(dftk_b200_kblock* kb, const double* V);

/* The same potential for all k-blocks of one spin channel: ONE pre-scaled copy per (grid, spin) instead of one per block
 * (all blocks of a spin share the term potentials, src/terms/Hamiltonian.jl:200-227).  A block opts in with
 * dftk_b200_kblock_use_grid_potential(kb, spin) (spin = -1: back to its own copy). */
int dftk_b200_grid_set_potential(dftk_b200_grid* grid, int spin, const double* V);
int dftk_b200_kblock_use_grid_potential(dftk_b200_kblock* kb, int spin);

/* Frees the scratch a k-block has grown (LOBPCG workspaces, INT8 residue-plane pools of the solver, FFT intermediates); the
 * operator data (kinetic energies, projectors and their prepared planes, potential) stays, and every later call re-grows what
 * it needs.  A 503-band solve at n_pw = 264 859 holds ≈ 60 GB of such scratch (no reference counterpart: Julia's GC does this). */
int dftk_b200_kblock_trim(dftk_b200_kblock* kb);

/* ---- sphere <-> real-space transforms (ifft!/fft! with Gvec_mapping, src/fft.jl:110-122,162-172) */
int dftk_b200_fft_sphere_to_real(dftk_b200_kblock* kb, const void* psi /*n_pw×n_bands*/,
                                 void* out_real /*N_fft×n_bands complex*/, int64_t n_bands,
                                 int normalize);
int dftk_b200_fft_real_to_sphere(dftk_b200_kblock* kb, const void* in_real /*N_fft×n_bands*/,
                                 void* out /*n_pw×n_bands*/, int64_t n_bands, int normalize);

/* ---- Hψ (LinearAlgebra.mul!(Hψ, ::DftHamiltonianBlock, ψ), src/terms/Hamiltonian.jl:137-192) ----
 * hpsi = FFT[V·IFFT[psi]] + kin·psi + P (D (P' psi)), all bands of the block in one batched pass. */
int dftk_b200_apply_h(dftk_b200_kblock* kb, const void* psi, void* hpsi, int64_t n_bands);
/* individual operators, `apply!(out, op, in)` semantics of src/terms/operators.jl (ACCUMULATE into hpsi
 * when accumulate != 0): parts bitmask 1 = local (RealSpaceMultiplication :71-78),
 * 2 = kinetic (FourierMultiplication :104-112), 4 = nonlocal (NonlocalOperator :119-129) */
int dftk_b200_apply_terms(dftk_b200_kblock* kb, const void* psi, void* hpsi, int64_t n_bands,
                          int parts, int accumulate);
/* per-band <psi|kin|psi> and <psi|P D P'|psi> (ene_ops, kinetic.jl:40-57, nonlocal.jl:31-47); host out */
int dftk_b200_band_energies(dftk_b200_kblock* kb, const void* psi, int64_t n_bands,
                            double* ekin_host, double* enl_host);

/* the same for all k-blocks of a rank in four launches and ONE synchronisation (blocks of <= 32 bands and <= 96 projectors);
 * ekin_host / enl_host: n_blocks × ld_out, either may be NULL */
int dftk_b200_band_energies_multi(int64_t n_blocks, dftk_b200_kblock* const* kblocks, const void* const* psi,
                                  const int32_t* n_bands, int64_t ld_out, double* ekin_host, double* enl_host);

/* ---- LOBPCG (lobpcg_hyper, src/eigen/diag_lobpcg_hyper.jl:5-18 -> LOBPCG,
 *      src/eigen/lobpcg_hyper_impl.jl:354-582, PreconditionerTPA src/eigen/preconditioners.jl:27-78) ----
 * X: n_pw × n_bands, in: guess, out: eigenvectors (device).  lambda/resid: host, n_bands each. */
int dftk_b200_lobpcg(dftk_b200_kblock* kb, void* X, int64_t n_bands, double tol, int miniter,
                     int maxiter, int64_t n_conv_check, int use_tpa_preconditioner,
                     double* lambda_host, double* resid_host, int* n_iter, int64_t* n_matvec,
                     int* converged);

/* All (k, spin) blocks of a rank at once (diagonalize_all_kblocks, src/eigen/diag.jl:9-65: the per-k eigenproblems are
 * independent).  Same algorithm and same results per block as dftk_b200_lobpcg; for blocks of <= 32 bands the solves advance
 * in lockstep and every operation of all blocks is ONE kernel launch (one host synchronisation per round instead of per
 * block) -- the launch-latency-bound regime of small cells with many k-points.  X[i]: n_pw_i × n_bands (device);
 * lambda_host / resid_host: n_blocks × n_bands (block-major); n_iter / n_matvec / converged: n_blocks each. */
int dftk_b200_lobpcg_multi(int64_t n_blocks, dftk_b200_kblock* const* kblocks, void* const* X, int64_t n_bands,
                           double tol, int miniter, int maxiter, int64_t n_conv_check, int use_tpa_preconditioner,
                           double* lambda_host, double* resid_host, int* n_iter, int64_t* n_matvec, int* converged);

/* ONE k-block solved by all ranks of a distributed context together (single-k multi-GPU; the reference cannot use more
 * than one process for one k-point, docs/src/tricks/parallelization.md:83-84).  Collective: every rank of the context calls
 * it with its own k-block handle of the SAME k-point and the same X (n_pw × n_bands, device, identical on all ranks).
 * The plane-wave rows of every tall block of lobpcg_hyper_impl.jl are cut into one slab per rank: Gram products, norms and
 * Rayleigh quotients are local products completed by ncclAllReduce, the small dense algebra (Cholesky, Rayleigh-Ritz) runs
 * replicated, and H is applied band-wise after a rows <-> bands exchange over NCCL point-to-point.  On return X holds the
 * eigenvectors on every rank; the other outputs are those of dftk_b200_lobpcg.  exchange_bytes (may be NULL): bytes this
 * rank sent in the rows <-> bands exchanges. */
int dftk_b200_lobpcg_slab(dftk_b200_kblock* kb, void* X, int64_t n_bands, double tol, int miniter, int maxiter,
                          int64_t n_conv_check, int use_tpa_preconditioner, double* lambda_host, double* resid_host,
                          int* n_iter, int64_t* n_matvec, int* converged, double* exchange_bytes);

/* Start vectors (random_orbitals, src/common/orbitals.jl:82-87: orthonormalised complex normal numbers) for several
 * k-blocks at once: X[i] (n_pw_i × n_bands, device) is filled and orthonormalised on the device. */
int dftk_b200_random_orbitals(int64_t n_blocks, dftk_b200_kblock* const* kblocks, void* const* X, int64_t n_bands,
                              uint64_t seed);

/* ---- density (compute_density inner loop, src/densities.jl:32-44):
 *      rho[:,:,:] += sum_n occ_w[n] |IFFT psi_n|² / Ω   with occ_w[n] = occupation·kweight (host) ---- */
int dftk_b200_density_accumulate(dftk_b200_kblock* kb, const void* psi, const double* occ_w_host,
                                 int64_t n_bands, double* rho /*dev: N_fft doubles of this spin*/);

/* compute_density's loop over the k-blocks of a rank in one call: rho (n_spin × N_fft, device) += contributions of all
 * blocks (each into the channel of its spin); occ_w_host: n_blocks × ld_w.  Blocks that share the grid's register FFT engine
 * are transformed together (two launches for all of them) and accumulated by one launch per spin channel. */
int dftk_b200_density_accumulate_multi(int64_t n_blocks, dftk_b200_kblock* const* kblocks, const void* const* psi,
                                       const double* occ_w_host, int64_t ld_w, const int32_t* n_bands, double* rho);

/* ---- collectives (mpi_sum!/mpi_min/mpi_max over basis.comm_kpts, src/common/mpi.jl:19-31) ---- */
int dftk_b200_allreduce(dftk_b200_ctx* ctx, void* buf /*dev*/, int64_t count, int dtype,
                        int op /*0 sum, 1 min, 2 max*/);
int dftk_b200_allgather(dftk_b200_ctx* ctx, const void* send /*dev*/, void* recv /*dev*/,
                        int64_t count_per_rank, int dtype);

/* ---- SCF plumbing next to the hot path ----
 * Pointwise exchange-correlation (replaces the Libxc dispatch, ext/DFTKCUDAExt.jl:17-25, src/terms/xc.jl:104-113).
 * functional_mask: 1 lda_x | 2 lda_c_vwn | 4 lda_c_pw | 8 gga_x_pbe | 16 gga_c_pbe.  Arrays are component-major
 * device doubles: rho[n_spin][n], sigma[1|3][n] (uu, ud, dd; GGA only), e[n] (energy per volume),
 * vrho[n_spin][n], vsigma[1|3][n] -- the quantities libxc returns as zk*rho, vrho, vsigma. */
int dftk_b200_xc_evaluate(dftk_b200_ctx* ctx, int functional_mask, int n_spin, int64_t n_points,
                          const double* rho, const double* sigma, double* e, double* vrho, double* vsigma);
/* accumulate_over_symmetries! + normalisation (src/symmetry.jl:282-327,340-357) on Fourier coefficients of the cube:
 * out[G] = 1/n_sym * sum_s exp(-2 pi i G.tau_s) in[S_s^-1 G]  (terms outside the FFT box dropped).
 * invS: n_sym row-major 3x3 integer matrices (host), tau: n_sym fractional translations (host). */
int dftk_b200_symmetrize_fourier(dftk_b200_grid* grid, const void* rho_fourier_in, void* rho_fourier_out,
                                 int n_sym, const int32_t* invS, const double* tau);

/* ---- Hellmann-Feynman forces (compute_forces, src/postprocess/forces.jl:24-30) ----
 * Local term (forces_local, src/terms/local.jl:152-181): for every atom of one species,
 *   F_a,α = -Re( Σ_G -2πi G_α e^{-2πi G·r_a} w_G ),  w_G = conj(ρ_G) v_loc(|G|) / sqrt(Ω)  (device, N_fft complex),
 * positions: 3·n_atoms fractional coordinates (host), forces_host: 3·n_atoms reduced-coordinate forces. */
int dftk_b200_local_forces(dftk_b200_grid* grid, const void* w, int n_atoms, const double* positions,
                           double* forces_host);
/* Nonlocal term (compute_forces(::TermAtomicNonlocal), src/terms/nonlocal.jl:49-100) for one k-block: per projector
 * row j and direction α the contribution  -Σ_n occ_w[n] 2 Re <ψ_n| P D (dP_j/dR_α)† |ψ_n>  with
 * dP/dR_α = -2πi (G+k)_α P, obtained from four projections P†[ψ, p_x ψ, p_y ψ, p_z ψ] instead of the reference's
 * 3·n_atoms full-height GEMM pairs.  gpk: 3 × n_pw reduced G+k components, component-major (device);
 * rows_host: 3 × n_proj (α-major).  The caller sums the rows of each atom, allreduces over k and symmetrises. */
int dftk_b200_nonlocal_force_rows(dftk_b200_kblock* kb, const void* psi, const double* occ_w_host, int64_t n_bands,
                                  const double* gpk, double* rows_host);

/* Ewald energy and forces of the ionic point charges (energy_forces_ewald, src/terms/ewald.jl:64-168, q = 0): the real-space
 * and the reciprocal-space lattice sums as one kernel each.  lattice: 3×3 column-major (columns = lattice vectors), charges:
 * n_atoms, positions: 3·n_atoms fractional (all host); eta and the summation limits (|G_i| <= glims[i], |R_i| <= rlims[i]) are
 * the caller's (ewald.jl:86-104).  energy_host: 1 double (Hartree); forces_host: 3·n_atoms, reduced coordinates; either NULL. */
int dftk_b200_ewald(dftk_b200_ctx* ctx, const double* lattice, int n_atoms, const double* charges, const double* positions,
                    double eta, const int32_t* glims, const int32_t* rlims, double* energy_host, double* forces_host);

/* ---- setup kernels: the O(n_atoms × N) structure-factor work before the first SCF step ----
 * out[G] = Σ_a c_a exp(-2πi G·r_a) on the whole FFT cube (G from the cube index, src/fft.jl:24-31): the atomic sums of
 * build_local_potential (src/terms/local.jl:108-138) and guess_density (src/density_methods.jl:103-181).
 * positions: 3·n_atoms fractional (host), coefficients: n_atoms (host) or NULL = 1, out: N_fft complex (device). */
int dftk_b200_structure_factor(dftk_b200_grid* grid, int n_atoms, const double* positions, const double* coefficients, void* out);
/* P[(a, p), G] = exp(-2πi (G+k)·r_a) · ff[p, G]: the projector table of one species for one k-block
 * (build_projection_vectors, src/terms/nonlocal.jl:166-199).  gpk: 3 × n_pw reduced G+k, component-major (device);
 * positions: 3·n_atoms (host); form_factors: n_rows × n_pw complex, row = projector (l, m, i) already divided by sqrt(Ω)
 * (device); P: (n_atoms·n_rows) × n_pw complex, i.e. the column-major n_pw × n_proj block of these atoms (device). */
int dftk_b200_build_projectors(dftk_b200_ctx* ctx, int64_t n_pw, const double* gpk, int n_atoms, const double* positions,
                               int n_rows, const void* form_factors, void* P);

/* ---- small dense helpers used by the host driver (columnwise_dots, src/common/linalg.jl:2-15) ---- */
int dftk_b200_columnwise_dots(dftk_b200_ctx* ctx, const void* A, const void* B, int64_t n_rows,
                              int64_t n_cols, void* out_host /*complex[n_cols]*/);
/* out (n_cols_a × n_cols_b, HOST, column-major complex) = A' B for tall column-major device blocks with at most 96 columns
 * each: one fused launch.  With real vectors viewed as complex pairs the real part is the real Gram matrix -- the history
 * dot products of Anderson mixing (src/scf/anderson.jl:81-130) and of GMRES (LdosMixing, src/scf/mixing.jl:283). */
int dftk_b200_tall_gram(dftk_b200_ctx* ctx, const void* A, int64_t lda, int64_t n_cols_a, const void* B, int64_t ldb,
                        int64_t n_cols_b, int64_t n_rows, void* out_host);
/* C = alpha * op(A) * B + beta * C on complex128 column-major device arrays (own DMMA kernels);
 * transA: 0 = N, 2 = C (conjugate transpose) */
int dftk_b200_zgemm(dftk_b200_ctx* ctx, int transA, int64_t m, int64_t n, int64_t k,
                    const double* alpha2, const void* A, int64_t lda, const void* B, int64_t ldb,
                    const double* beta2, void* C, int64_t ldc);

#ifdef __cplusplus
}
#endif
#endif
