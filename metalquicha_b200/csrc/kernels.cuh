// kernels.cuh -- host-callable launchers of every CUDA kernel of the engine.
// Implementations: prep_kernels.cu (layout conversion, assembly),
// j_kernels.cu (Coulomb, HBM-bound), k_kernels.cu (exchange, FP64 DMMA).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace mqcb200 {

// ---- layout / preparation -------------------------------------------------
// Full-square slabs (q_count of them, each n*n column-major) -> packed rows.
void launch_pack_tensor(const double *d_full, int n, int q_count, double *d_packed, cudaStream_t s);
// Same from LOWER-packed slabs: slab q holds its columns nu = 0..n-1 one after the other, column
// nu being rows nu..n-1 (n(n+1)/2 doubles per slab) -- what set_tensor ships over PCIe.
void launch_pack_tensor_lower(const double *d_lower, int n, int q_count, double *d_packed, cudaStream_t s);
// Counter-based synthetic tensor written directly in packed form.
void launch_synth_tensor(double *d_packed, int n, int q_global_begin, int q_count, uint64_t seed,
                         double scale, cudaStream_t s);
// w[L]: the density in packed order with the off-diagonal weight folded in.
// (batch > 1: blockIdx.y walks the fragments of a batch, every operand moving on by its stride -- doubles)
void launch_pack_density(const double *d_density, int n, double *d_w, const int *d_skip_flag, cudaStream_t s, int batch = 1,
                         size_t d_stride = 0, size_t w_stride = 0);
// C(n x n_occ, ld) -> fragment-ordered, zero-padded [nt][nib][4][32] operand.
// d_cep (nullable): the same coefficients in accumulator order, for the Coulomb-vector epilogue.
void launch_pack_coeff(const double *d_coeff, int ldc, int n, int n_occ, int nib, double *d_ctf, double *d_cep,
                       cudaStream_t s, int batch = 1, size_t coeff_stride = 0, size_t ctf_stride = 0);
// Sum the J / K partial buffers in fixed order and unpack to full n x n matrices.
// Either output may be null.  k_factor multiplies K (2 for RHF, 1 per spin).
// With d_fock non-null the Fock matrix F = H + jf*J - kf*K is assembled in the same pass.
// n_ksplits_diag: partials held by the DIAGONAL K tiles (<= n_ksplits; -1 when equal).
// n_ksplits_edge: partials held by the off-diagonal K tiles of the last panel row (-1 when equal to n_ksplits).
void launch_finalize_jk(const double *d_jpart, int n_jslices, const double *d_kpart, int n_ksplits,
                        int ktile, int n, double k_factor, double *d_j, double *d_k, cudaStream_t s,
                        const double *d_h = nullptr, double jf = 0.0, double kf = 0.0, double *d_fock = nullptr,
                        int n_ksplits_diag = -1, int batch = 1, size_t jpart_stride = 0, size_t kpart_stride = 0,
                        size_t out_stride = 0, int n_ksplits_edge = -1);
// F = H + jf*J - kf*K  (any of J/K may be null == zero).
void launch_assemble_fock(const double *d_h, const double *d_j, const double *d_k, double jf, double kf,
                          int n, double *d_fock, cudaStream_t s);
// G = J + ka*Ka + kb*Kb (null operands are skipped).
void launch_combine_g(const double *d_j, const double *d_ka, const double *d_kb, double ka, double kb, int n,
                      double *d_g, cudaStream_t s);
// e = 1/2 sum D (H + F) into d_out[0]; fixed-order two-level reduction.  d_scratch: 129 doubles,
// zero-initialised once (the kernel leaves its counter at zero).
void launch_energy(const double *d_density, const double *d_h, const double *d_fock, int n,
                   double *d_scratch, double *d_out, cudaStream_t s, int batch = 1, size_t mat_stride = 0,
                   size_t out_stride = 0);   // batch > 1: d_scratch holds 130 doubles per fragment

// ---- whitening: Bp = half . Tp on the packed tensor (build_df_tensor's GEMM), slab by slab ----
constexpr int WHITEN_MAX_RANKS = 8;
struct WhitenDst {                      // where the rows of the whitened slab go: rank r owns rows [q_begin[r], q_begin[r+1])
  double *base[WHITEN_MAX_RANKS];       // packed tensor of every rank (own entry included), device-addressable
  int q_begin[WHITEN_MAX_RANKS + 1];
  int n_ranks;
  long long ld;                         // packed row length L
  long long col_off;                    // first packed column of the slab
};
size_t whiten_half_elems(int m_rows, int naux);   // doubles of the fragment-ordered copy of m_rows rows of `half`
void launch_pack_half(const double *d_half, int naux, int row0, int m_rows, double *d_af, cudaStream_t s);
void launch_pack_slab(const double *d_slab, int n, int naux, int nu_begin, int nu_count, double *d_tp, cudaStream_t s);
void launch_whiten_slab(const double *d_af, int m_rows, int naux, const double *d_tp, long long ld_tp, long long n_cols,
                        const WhitenDst &dst, cudaStream_t s);
// metric^(-1/2) pieces: one round of one-sided Jacobi on the columns of g (and v); identity; scaling by lambda^-1/2
void launch_hestenes_round(double *d_g, double *d_v, int n, int round, int *d_rotated, cudaStream_t s);
// The whole iteration (all rounds, all sweeps, convergence test) in one cooperative launch; d_state: 2 ints
// ([1] receives the sweep count).  False when the device cannot launch cooperatively: use the per-round kernels.
bool launch_hestenes_solve(double *d_g, double *d_v, int n, int max_sweeps, int *d_state, cudaStream_t s);
void launch_set_identity(double *d_v, int n, cudaStream_t s);
void launch_metric_scale(const double *d_g, const double *d_v, int n, double threshold, double *d_scaled, double *d_lambda,
                         int *d_n_kept, cudaStream_t s);

// ---- cross-GPU sum of [J|K] over NVLink peer memory (sharded builds) ----------
constexpr int XGPU_MAX_RANKS = 8;
struct XgpuPeers {                         // device-addressable pointers of every rank (own entry included)
  const double *in[XGPU_MAX_RANKS];        // partial [J|K_a|K_b]
  double *out[XGPU_MAX_RANKS];             // reduced [J|K_a|K_b]
  unsigned long long *flags[XGPU_MAX_RANKS];   // [2*XGPU_MAX_RANKS] arrival epochs per rank
};
// out[k][first..first+count) = sum_r in[r][first..first+count) on every rank k, in rank order.
// d_error: device-addressable HOST-mapped int, set to 1 when a peer did not arrive within timeout_ns.
void launch_xgpu_allreduce(const XgpuPeers &peers, int n_ranks, int rank, unsigned long long epoch, size_t first,
                           size_t count, unsigned int *d_counter, int *d_error, unsigned long long timeout_ns,
                           cudaStream_t s);

// ---- J: two passes over the packed tensor -----------------------------------
struct JPlan {
  int n_seg;      // pass 1: segments of a packed row
  int n_slices;   // pass 2: slices of the auxiliary range
  size_t gamma_partial_elems;  // n_seg * q_count
  size_t j_partial_elems;      // n_slices * L
};
JPlan plan_j(int n, int q_count);
// d_skip_flag (nullable device int): when it reads non-zero the pass is skipped on the
// device (gamma comes from the half-transform instead, see launch_gamma_from_x).
void launch_j_gamma(const double *d_packed, long long L, int q_count, const double *d_w,
                    const JPlan &plan, double *d_gamma_partial, double *d_gamma, const int *d_skip_flag,
                    cudaStream_t s);
// gamma[q] = factor * (sum of part_a[q][:] + sum of part_b[q][:]) if *d_flag, else untouched.
void launch_gamma_from_x(const double *d_part_a, int stride_a, const double *d_part_b, int stride_b,
                         double factor, int q_count, const int *d_flag, double *d_gamma, cudaStream_t s);
// w[L] like launch_pack_density AND, in the same pass, *d_flag = 1 iff the density is the
// orbitals' own: max|D - f*(Ca Ca^T + Cb Cb^T)| <= 4*2^-53*max(n_a+n_b,16) * max(1, max|D|) over
// BOTH triangles, every element finite.  d_scratch: 24 bytes, zero before the first call (the
// kernel leaves it zero).  Maxima are order-independent, so the decision is deterministic.
void launch_density_prep(const double *d_density, int n, double *d_w, const double *d_ca, int lda, int na,
                         const double *d_cb, int ldb, int nb, double f, unsigned long long *d_scratch,
                         int *d_flag, cudaStream_t s);
void launch_j_accumulate(const double *d_packed, long long L, int q_count, const double *d_gamma,
                         const JPlan &plan, double *d_jpart, cudaStream_t s);
// Pass 2 shaped to sit BESIDE k_accumulate_kernel on every SM (64 threads x 64 registers, 24 KiB ring fed by
// bulk TMA): used when the Coulomb kernels run on their own stream next to the exchange kernels.
// Partials: jpart[n_slices][L], the same slices and summation order as launch_j_accumulate (bit-identical).
void launch_j_accumulate_tma(const double *d_packed, long long L, int q_count, const double *d_gamma, int n_slices,
                             int sm_count, double *d_jpart, cudaStream_t s);

// ---- fragment-sized problems: J and K in one pass over the tensor ---------------
struct FragPlan {
  int nt, nib, L, grid, n_ktiles;
  size_t smem_bytes, jpart_elems, kpart_elems;
};
bool fragment_path_applies(int n, int n_occ);       // n <= 80 and n_occ <= 64
FragPlan plan_fragment(int n, int n_occ, int q_count, int sm_count);
// d_ctf: coefficients packed with nib = ceil(n_occ/8).  Partials: jpart [grid][L],
// kpart [grid][n_ktiles][64*64] -- the layouts launch_finalize_jk sums (ktile = 64).
void launch_fragment_jk(const double *d_packed, int q_count, const double *d_w, const double *d_ctf,
                        const FragPlan &p, bool want_j, bool want_k, double *d_jpart, double *d_kpart,
                        cudaStream_t s, int batch = 1);   // batch > 1: fragment f owns slabs [f*q_count, (f+1)*q_count), w/ctf/partials contiguous per fragment

// ---- general strided-batched FP64 GEMM on DMMA + helpers of the gradient densities (gemm_kernels.cu)
// C[b] = alpha * op(A[b]) * op(B[b]) + beta * C[b]; column-major, op(A) is M x K, op(B) is K x N.
void launch_dgemm_batched(int M, int N, int K, double alpha, const double *a, long long lda, long long stride_a, bool ta,
                          const double *b, long long ldb, long long stride_b, bool tb, double beta, double *c,
                          long long ldc, long long stride_c, int batch, cudaStream_t s);
// Packed slabs -> full symmetric n x n slabs (column-major), q_count of them.
void launch_unpack_tensor(const double *d_packed, int n, int q_count, double *d_full, cudaStream_t s);
// gamma(:,:,p) (+)= rho[p] * D - w * z(:,:,p) for p_count auxiliary functions (rho, D, z may be null == 0).
void launch_gradient_gamma(const double *d_rho, const double *d_density, const double *d_z, double w, int n, int p_count,
                           bool accumulate, double *d_gamma, cudaStream_t s);

// ---- pieces of the general-size device-resident SCF step (gemm_kernels.cu) ----------------
void launch_scf_guess(const double *d_h, const double *d_s, int n, bool gwh, double *d_f, cudaStream_t s);
void launch_antisym(const double *d_a, int n, double *d_out, cudaStream_t s);                       // out = a - a^T
void launch_symmetrize_shift(double *d_g, int m, const double *d_shift, cudaStream_t s);           // g = (g+g^T)/2 + shift*1
// mode 0: sum a*b, 1: sum (a-b)^2, 2: sqrt(sum a*a); fixed order; d_scratch: 130 doubles, zeroed once
void launch_reduce(const double *d_a, const double *d_b, size_t count, int mode, double *d_scratch, double *d_out, cudaStream_t s);
void launch_eig_lambda(const double *d_g, const double *d_v, int m, const double *d_shift, double *d_lambda, cudaStream_t s);
void launch_rank_sort(const double *d_lambda, int m, double threshold, int *d_order, int *d_n_dropped, cudaStream_t s);
void launch_gather_columns(const double *d_src, int rows, const int *d_order, int k0, int cols, const double *d_lambda,
                           bool inv_sqrt, double *d_dst, cudaStream_t s);
void launch_lincomb(const double *d_vecs, size_t count, const double *d_coef, const int *d_slots, int n_terms, double *d_out,
                    cudaStream_t s);

// ---- device-resident SCF step for fragment-sized problems (scf_kernels.cu) ---------------
struct ScfStepLaunch {
  int batch = 1;                 // fragments: one CTA each, every pointer below moving on by block_stride doubles
  size_t block_stride = 0;
  const int *n_mo_dev = nullptr; // batch: orbitals surviving the overlap threshold, per fragment (stride block_stride ints*2)
  int n, n_mo, n_occ, diis_max, mode /*0 guess, 1 iteration*/, guess /*0 core, 1 GWH*/;
  const double *h, *s, *x;
  double *fock, *density, *coeff, *eps, *work, *diis_f, *diis_e, *diis_b;
  int *state;        // [0] n_stored [1] newest [2] iterations [3] converged
  double *scalars;   // [0] e_elec (in) [1] e_old [2] |dE| [3] rms(dD) [4] extrapolated
  double energy_tol, density_tol;
};
bool scf_path_applies(int n);                       // n <= 80
void configure_scf_kernels();
// x (n x n_mo) = U s^-1/2 over the eigenvalues of S above 1e-7; *d_n_mo = surviving orbitals.
void launch_scf_orthogonalizer(const double *d_s, int n, double *d_x, int *d_n_mo, cudaStream_t s, int batch = 1,
                               size_t block_stride = 0);
void launch_scf_step(const ScfStepLaunch &a, cudaStream_t s);

// ---- K: half-transform + symmetric accumulation on the FP64 tensor pipe -----
struct KPlan {
  int nb;          // n8-blocks per warp in the half-transform (BN = 16*nb)
  int n_ntiles;    // N tiles of the half-transform
  int nib;         // padded occupied count / 8   (= n_ntiles * 2 * nb)
  int nkc;         // padded occupied count / 16
  int nmb;         // 8-row blocks of X per (q, kc): 2*nt
  int ks_last;     // valid 4-wide k-subs in the last 16-wide chunk of the occupied range
  int ktile;       // edge of the square K tiles of the accumulation (64 or 128)
  int n_panels;    // ktile-row panels of K
  int n_ktiles;    // lower-triangular ktile x ktile tiles of K
  int n_splits;    // split of the auxiliary range in the accumulation (off-diagonal tiles)
  int n_splits_diag;  // ... of the diagonal tiles (fewer: they do 9/16 of the work per auxiliary function)
  int n_splits_edge;  // ... of the off-diagonal tiles of a partly filled last panel row (LB/8 of the work); == n_splits when there is no such class
  int q_chunk;     // auxiliary functions per half-transform launch
  int sm_count;    // persistent grid size of the half-transform
  int gamma_stride;  // Coulomb-vector partials per auxiliary function written by the half-transform
  int pair_off;      // rank-2 form: chunk k of the X half pairs with chunk k + pair_off of the C half (0 = SYRK)
  int npairs;        // rank-2 form: 16-wide chunks per half
  int trim_last;     // half-transform: the last n8-block of the padded occupied range is all padding and is skipped
  size_t x_elems_per_q;     // doubles of X per auxiliary function
  size_t kpart_elems;       // n_splits * n_ktiles * ktile*ktile
};
// rank2_occ > 0: plan the rank-2 (SYR2K) form for two n x rank2_occ factors; the half-transform then
// takes the stacked operand of launch_stack_factors (2*ceil16(rank2_occ) columns), n_occ is ignored.
KPlan plan_k(int n, int n_occ, int q_count, size_t workspace_limit_bytes, int sm_count, int rank2_occ = 0);
// S (n x 2*o16, ld n) = [X | 0 | C | 0]: each n x n_occ factor starts on a 16-column boundary.
void launch_stack_factors(const double *d_x, int ldx, const double *d_c, int ldc, int n, int n_occ, double *d_s,
                          cudaStream_t s);
// d_gamma_part (nullable, needs d_cep): [q_count][plan.gamma_stride] partials of sum_{mu,i} X_Q[mu,i] C[mu,i].
struct HalfTail { long long n_full; int split; };
HalfTail plan_half_tail(long long n_units, int ctas);
void launch_k_half_transform(const double *d_packed, long long L, int n, int q_count,
                             const double *d_ctf, const double *d_cep, const KPlan &plan, double *d_x,
                             double *d_gamma_part, cudaStream_t s);
void launch_k_accumulate(const double *d_x, int q_count, const KPlan &plan, double *d_kpart,
                         int accumulate, cudaStream_t s);
// Host side of set_tensor: the lower triangles (column nu: rows nu..n-1) of q_count full-square column-major
// slabs, src_slab_stride doubles apart, gathered back to back into dst (n(n+1)/2 doubles per slab).  Reads
// nothing outside the q_count slabs and writes nothing outside dst's q_count triangles (tests/native).
void gather_lower(const double *src, int n, size_t q_count, size_t src_slab_stride, double *dst);
// Host side of the any-size device SCF: diis_coefficients + solve_diis of src/methods/mqc_diis.f90:146-273 on
// the ring's cached overlaps.  overlap[8][8] is indexed by ring slot, `newest` is the 1-based slot of the newest
// vector (as in diis_state_t), n_stored <= dmax <= 8.  On success coef[i] / slots[i] (i = 0 oldest .. n_stored-1
// newest) are the extrapolation weights and the 0-based slots they belong to.  false: fewer than two vectors,
// or a pivot below 1e-14 (the caller then uses the plain Fock matrix).
bool diis_solve_host(const double *overlap, int newest, int n_stored, int dmax, double *coef, int *slots);
// One-time (per process and device) opt-in to large dynamic shared memory.
void configure_kernels();            // K kernels; calls the two below
void configure_fragment_kernels();
void configure_whiten_kernels();

}  // namespace mqcb200
