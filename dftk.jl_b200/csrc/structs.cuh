// Grid / k-block handle definitions shared by the translation units of libdftk_b200.
#pragma once
#include "common.cuh"
#include "fft_reg_fwd.cuh"
#include "fft_plan.h"

namespace dftk {
struct SphereTablesX;
// kernel entry points of the register two-pass engine for one factor pair (fft_reg.cu)
struct RegKernels {
  int A, B, T;
  const void *sphere_to_x, *y_backward, *z_apply, *z_to_cube, *z_from_cube, *z_density, *y_forward, *x_to_sphere;
  const void* z_apply_pipe;   // persistent, software-pipelined form of z_apply (cp.async staged input tiles)
  const void *m_sphere_to_x, *m_y_backward, *m_z_apply, *m_y_forward, *m_x_to_sphere, *m_z_density;   // many k-blocks per launch
};
// one k-block's share of a batched H-apply launch (local + kinetic part)
struct FftMultiItem {
  SphereTablesX T;
  const cplx* psi;
  long long ldpsi;
  cplx *W1, *W2;
  const double* V;
  cplx* out;
  long long ldout;
  const double* kin;
  const double* wts;     // density accumulation: band weights of this block (device) and their number
  int nb;
};
const RegKernels* reg_kernels_for(int n);   // nullptr: use the generic Stockham engine
void reg_set_attributes();
}  // namespace dftk

namespace dftk {
struct I8Operand {                  // an operand of C = A^H B converted to INT8 residue planes (i8emu.cu)
  const signed char* planes = nullptr;
  const int* exps = nullptr;
  int64_t cols = 0, k = 0, ldk = 0;
  int n_mod = 0;
};
}  // namespace dftk

struct dftk_b200_grid {
  dftk_b200_ctx* ctx;
  int nx, ny, nz;
  int64_t N;
  double omega;
  double ifft_norm, fft_norm;  // src/fft.jl:87-88
  dftk::FftPlan px, py, pz;
  dftk::DevBuf<double> twx, twy, twz;
  int Lx, Ly, Lz;
  const dftk::RegKernels *rx = nullptr, *ry = nullptr, *rz = nullptr;  // register engine per axis
  dftk::DevBuf<double> Vs[2];   // total local potential per spin (pre-scaled by 1/N), shared by the k-blocks that opt in
  bool has_Vs[2] = {false, false};
};

struct dftk_b200_kblock {
  dftk_b200_grid* grid;
  int64_t n_pw, n_proj;
  int spin;
  double kweight;
  dftk::SphereTablesHost Th;
  dftk::SphereTablesX T;  // device view
  dftk::DevBuf<int> d_col_start, d_col_cnt, d_slot_ix, d_slot_src, d_zlist, d_colmap, d_zc_of, d_pl_s0, d_pl_n0, d_pl_s1, d_pl_n1, d_pl_col0, d_cx_s0, d_cx_n0, d_cx_s1,
      d_cx_n1;
  dftk::DevBuf<double> kin;       // n_pw (may be empty)
  bool has_kin = false;
  dftk::DevBuf<dftk::cplx> P;     // n_pw x n_proj
  std::vector<double> D_host;     // n_proj x n_proj
  dftk::DevBuf<dftk::cplx> Dc;    // complex copy of D on the device (n_proj x n_proj)
  dftk::DevBuf<signed char> i8_pool[8];  // gemm_backend 4: residue planes of the LOBPCG blocks (cache slots of a solve)
  dftk::DevBuf<int> i8_epool[8];
  dftk::I8Operand i8_Pop;                // prepared projector table (kept for the lifetime of the block)
  dftk::DevBuf<signed char> i8_psi_planes;   // planes of the orbitals entering P'psi (own buffer: the pool slots belong to LOBPCG's cache)
  dftk::DevBuf<int> i8_psi_exps;
  dftk::DevBuf<signed char> i8_planes;   // gemm_backend 4: cached INT8 residue planes of P (built at first use)
  dftk::DevBuf<int> i8_exps;
  dftk::DevBuf<dftk::cplx> PD;    // P D (n_pw x n_proj), kept when n_proj is small: Hψ += (P D)(P'ψ) as two batched small products
  dftk::DevBuf<double> V;         // N, pre-scaled by 1/N (fft_norm*ifft_norm)
  bool has_V = false;
  int grid_V = -1;                // >= 0: use grid->Vs[grid_V] instead of the block's own copy
  const double* Vp() const { return grid_V >= 0 ? grid->Vs[grid_V].p : V.p; }
  // scratch
  dftk::DevBuf<dftk::cplx> W1, W2;    // pruned intermediates for a chunk of bands
  dftk::DevBuf<dftk::cplx> proj;      // n_proj x n_bands (+ D*proj)
  dftk::DevBuf<dftk::cplx> lobpcg_ws; // big LOBPCG workspace
  dftk::DevBuf<dftk::cplx> small_ws;  // small dense LOBPCG workspace
  dftk::DevBuf<dftk::cplx> slab_x, slab_stage;   // slab-distributed solve (lobpcg_run_slab): this rank's rows of X, reassembly staging
  dftk::DevBuf<double> wts;
  dftk::DevBuf<double> scal;          // per-block scalars of a LOBPCG solve (several blocks are solved side by side)
};

namespace dftk {
// fft.cu
int band_chunk_for(dftk_b200_kblock* kb, int64_t n_bands);
void fft_cube_inplace(dftk_b200_grid* g, cplx* data, int sign, int64_t batch);
void kb_sphere_to_planes(dftk_b200_kblock* kb, const cplx* psi, int64_t ldpsi, int nb);
void kb_planes_to_sphere(dftk_b200_kblock* kb, cplx* out, int64_t ldout, int nb, double scale,
                         const double* kin, const cplx* psi, int64_t ldpsi, int accumulate);
void kb_apply_local_kinetic(dftk_b200_kblock* kb, const cplx* psi, cplx* hpsi, int64_t n_bands,
                            bool with_local, bool with_kin, bool accumulate);
// the same for several k-blocks of ONE grid in five launches in total; returns false (nothing done) when the blocks do
// not qualify (different grids, generic FFT engine, missing potential / kinetic term): the caller then loops
bool kb_apply_local_kinetic_multi(int n, dftk_b200_kblock* const* kbs, const cplx* const* psi, cplx* const* hpsi,
                                  const int* n_bands, const void* (*upload)(void* self, const void* host, size_t bytes), void* self);
void kb_sphere_to_real(dftk_b200_kblock* kb, const cplx* psi, cplx* cube, int64_t n_bands, double scale);
void kb_real_to_sphere(dftk_b200_kblock* kb, const cplx* cube, cplx* out, int64_t n_bands, double scale);
void kb_density_accumulate(dftk_b200_kblock* kb, const cplx* psi, const double* occ_w_host,
                           int64_t n_bands, double* rho);
// all k-blocks of a rank: stages A, B batched over the blocks, one accumulation launch per spin channel;
// rho: n_spin x N (device), occ_w_host: n x ld_w.  Returns false (nothing done) when the blocks do not qualify.
bool kb_density_accumulate_multi(int n, dftk_b200_kblock* const* kbs, const cplx* const* psi, const double* occ_w_host,
                                 int64_t ld_w, const int* n_bands, double* rho);
void fft_set_attributes();
// blas.cu
void zgemm(dftk_b200_ctx* ctx, int transA, int64_t m, int64_t n, int64_t k, cplx alpha, const cplx* A,
           int64_t lda, const cplx* B, int64_t ldb, cplx beta, cplx* C, int64_t ldc, bool upper_only = false);
// upper_only: transA == 2 -> only tiles on/above the diagonal are computed (Hermitian result);
//             transA == 0 -> B is upper triangular (trmm-like: half the flops)
void blas_set_attributes();
void kb_apply_nonlocal(dftk_b200_kblock* kb, const cplx* psi, cplx* hpsi, int64_t n_bands);
void columnwise_dots(dftk_b200_ctx* ctx, const cplx* A, int64_t lda, const cplx* B, int64_t ldb,
                     int64_t n_rows, int64_t n_cols, cplx* out_dev);
void kin_dots(dftk_b200_ctx* ctx, const cplx* X, int64_t ldx, const double* kin, int64_t n_rows,
              int64_t n_cols, double* out_dev);
void scale_kin_add(dftk_b200_ctx* ctx, const cplx* psi, cplx* hpsi, const double* kin, int64_t n_rows,
                   int64_t n_cols, int accumulate);
// xc.cu
void xc_evaluate(dftk_b200_ctx* ctx, int mask, int n_spin, bool gga, int64_t N, const double* rho,
                 const double* sigma, double* e, double* vrho, double* vsigma);
void symmetrize_fourier(dftk_b200_grid* g, const cplx* in, cplx* out, int n_sym, const int* invS_host,
                        const double* tau_host);
// i8emu.cu (experimental, option gemm_backend = 2)
void zgemm_i8_cn(dftk_b200_ctx* ctx, int64_t m, int64_t n, int64_t k, const cplx* A, int64_t lda, const cplx* B, int64_t ldb,
                 cplx* C, int64_t ldc, int tc_mode, const signed char* ra_cached = nullptr, const int* ea_cached = nullptr);
void i8tc2_products(dftk_b200_ctx* ctx, const signed char* ra, const signed char* rb, int64_t m, int64_t n, int64_t ldk,
                    int n_mod, short* part, int* res, bool upper_only);
I8Operand i8_prepare(dftk_b200_ctx* ctx, const cplx* X, int64_t ld, int64_t cols, int64_t k, DevBuf<signed char>& store,
                     DevBuf<int>& estore);
void i8_gram(dftk_b200_ctx* ctx, const I8Operand& A, const I8Operand& B, cplx* C, int64_t ldc, bool upper_only);
// C (rows x n) = alpha sum_b A_b B[rows of b, :] + beta C from prepared tall operands (update-type products)
void i8_update(dftk_b200_ctx* ctx, int n_blocks, const I8Operand* A, const cplx* B, int64_t ldb, int64_t n, cplx* C, int64_t ldc,
               double alpha, double beta);
void i8tc2_products_nn(dftk_b200_ctx* ctx, int n_blocks, const signed char* const* ra, const int* kcols, const int* k_off,
                       int64_t ldm, const signed char* rb, int64_t ldkb, int64_t m, int64_t n, int n_mod, signed char* resid);
void i8tc2_set_attributes();
bool zgemm_i8_nn(dftk_b200_ctx* ctx, int64_t m, int64_t n, int64_t k, const cplx* A, int64_t lda, const cplx* B, int64_t ldb,
                 cplx* C, int64_t ldc, bool accumulate);
// i8tc.cu (experimental, option gemm_backend = 3): tcgen05.mma.kind::i8 kernel of the integer products
void i8tc_products(dftk_b200_ctx* ctx, const signed char* ra, const signed char* rb, int64_t m, int64_t n, int64_t ldk,
                   int n_mod, short* part, int* res);
void i8tc_set_attributes();
// forces.cu
void local_forces(dftk_b200_grid* g, const cplx* w, int n_atoms, const double* pos_host, double* out_host);
void kb_nonlocal_force_rows(dftk_b200_kblock* kb, const cplx* psi, const double* occ_w_host, int64_t n_bands,
                            const double* gpk, double* out_host);
void ewald(dftk_b200_ctx* ctx, const double* lattice_colmajor, int n_atoms, const double* charges, const double* positions,
           double eta, const int* glims, const int* rlims, double* energy_host, double* forces_host);
// setup.cu
void structure_factor(dftk_b200_grid* g, int n_atoms, const double* pos_host, const double* coeff_host, cplx* out);
void build_projectors(dftk_b200_ctx* ctx, int64_t n_pw, const double* gpk, int n_atoms, const double* pos_host, int n_rows,
                      const cplx* ff, cplx* P);
// lobpcg.cu
int lobpcg_run(dftk_b200_kblock* kb, cplx* X, int64_t M, double tol, int miniter, int maxiter,
               int64_t n_conv_check, bool use_prec, double* lambda_host, double* resid_host,
               int* n_iter, int64_t* n_matvec, int* converged);
int lobpcg_run_multi(int64_t n_blocks, dftk_b200_kblock* const* kbs, cplx* const* Xs, int64_t M, double tol, int miniter,
                     int maxiter, int64_t n_conv_check, bool use_prec, double* lambda_host, double* resid_host, int* n_iter,
                     int64_t* n_matvec, int* converged);
int lobpcg_run_slab(dftk_b200_kblock* kb, cplx* Xfull, int64_t M, double tol, int miniter, int maxiter, int64_t n_conv_check,
                    bool use_prec, double* lambda_host, double* resid_host, int* n_iter, int64_t* n_matvec, int* converged,
                    double* exchange_bytes);
void random_orbitals_multi(int64_t n_blocks, dftk_b200_kblock* const* kbs, cplx* const* Xs, int64_t M, uint64_t seed);
void band_energies_multi(int64_t n, dftk_b200_kblock* const* kbs, const cplx* const* psi, const int* n_bands, int64_t ld_out,
                         double* ekin_host, double* enl_host);
void tall_gram(dftk_b200_ctx* ctx, const cplx* A, int64_t lda, int nA, const cplx* B, int64_t ldb, int nB, int64_t n_rows,
               cplx* out_host);
void lobpcg_set_attributes();
}  // namespace dftk
