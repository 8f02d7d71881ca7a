/* mqcb200.h -- C ABI of the B200-native density-fitted J/K Fock-build engine.
 *
 * This is the drop-in boundary for ONE hot path of JorgeG94/metalquicha: the
 * CPU routine `build_fock_df` and the tensor it consumes.  Every entry point
 * below names the reference interface it replaces (paths relative to the
 * reference tree).  The Fortran side binds these with ISO_C_BINDING
 * (metalquicha_b200/fortran/mqc_b200_iface.f90); INTEGRATION.md shows the
 * patch a maintainer applies.
 *
 * Conventions (reference: backends/cuest/bindings/cublas.f90:20-44,
 * src/interface/mqc_capi_status.f90:21-23):
 *   - every function returns int: MQCB200_OK=0, MQCB200_FAIL=1, MQCB200_BAD_HANDLE=2;
 *     after a non-zero return mqcb200_last_error() gives the message (per thread);
 *   - handles are opaque `void*`; one handle per host thread per GPU; calls block
 *     (results are valid on return);
 *   - all matrices are float64, COLUMN-major (Fortran order), contiguous;
 *   - the caller owns every host array; the engine owns its device copies.
 * There is no CPU fallback: without a CUDA device mqcb200_create fails.
 */
#ifndef MQCB200_H
#define MQCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MQCB200_OK 0
#define MQCB200_FAIL 1
#define MQCB200_BAD_HANDLE 2

#define MQCB200_SLOT_FULL_RANGE 0 /* bmat    (mqc_libcint_rhf.f90:420,486-491) */
#define MQCB200_SLOT_ATTENUATED 1 /* bmat_lr (mqc_libcint_rhf.f90:421,493-497) */
#define MQCB200_NUM_SLOTS 2

/* Number of per-phase timers returned by mqcb200_last_timings. */
#define MQCB200_NUM_TIMERS 8
/* indices: 0 upload(H2D+prep) 1 J pass 1 (gamma) 2 J pass 2 3 K half-transform
 *          4 K accumulation 5 finalize/assemble 6 all-reduce 7 download(D2H) */

/* ---- lifetime ------------------------------------------------------------
 * Replaces the lazily created per-rank device context of the reference's GPU
 * backend (backends/cuest/backend/mqc_cuest_context.f90:166-233, :285-299):
 * binds device `mod(device_rank, device_count)` and owns grow-only pools. */
int mqcb200_create(int device_rank, void **handle);
int mqcb200_destroy(void *handle);
/* Message of the last failure on this thread (cf. mqc_run_last_error,
 * src/interface/mqc_capi_run.f90:147). Always NUL-terminated. */
void mqcb200_last_error(int buffer_len, char *buffer);
int mqcb200_version(void);
/* CUDA stream the engine launches on (a cudaStream_t), for callers that want to
 * bracket calls with their own events. */
int mqcb200_get_stream(void *handle, void **stream);
/* Upper bound (bytes) of the K half-transform scratch (cf. DF_EXCHANGE_BUFFER_BYTES,
 * backends/cuest/backend/mqc_cuest_integrals.f90:122). Default 4 GiB. */
int mqcb200_set_workspace_limit(void *handle, size_t bytes);
/* Packed-tensor size (bytes per handle) from which a build whose density equals f*C*C^T
 * takes its Coulomb vector from the half-transform instead of a pass over the tensor.
 * Default 32 MiB (below that the pass over B is cheaper than the extra launch); 0 = always try,
 * SIZE_MAX = never.  The decision is made on the device, per build: max|D - f*C*C^T| (both
 * triangles) <= 4*2^-53*max(n_occ,16)*max(1,max|D|) and every element finite.  The fused
 * Coulomb vector then differs from the reference's sum(B_P*D) by at most that tolerance times
 * sum|B_P| -- the size of the rounding error of the reference's own n^2-term sum. */
int mqcb200_set_fuse_threshold(void *handle, size_t bytes);
/* 1 (default): the HBM-bound Coulomb kernels run on a second stream, concurrently with the
 * tensor-bound exchange kernels; 0: one stream.  Results are bit-identical either way. */
int mqcb200_set_overlap(void *handle, int on);

/* ---- the fitted tensor ---------------------------------------------------
 * Replaces the host array `bmat(nao*nao, naux)` that run_libcint_rhf receives
 * from build_df_tensor (mqc_libcint_integrals.F90:913-990; layout :985-986,
 * :1425-1437).  B is packed on the device into lower-triangular 16x16 tiles; only
 * the lower triangle mu >= nu of each slab is read (slabs are symmetric by
 * construction).  The tensor stays resident until the next set on that slot.
 *
 * The *_shard form holds only auxiliary functions [q_begin, q_begin+q_count) of a
 * tensor with naux_total of them (whole-molecule builds sharded over GPUs);
 * `b` then points at the first slab of the shard. */
int mqcb200_set_tensor(void *handle, int slot, int n, int naux, const double *b);
int mqcb200_set_tensor_shard(void *handle, int slot, int n, int naux_total,
                             int q_begin, int q_count, const double *b_shard);
/* b = three . metric^(-1/2) on the device (build_df_tensor's last stage,
 * mqc_libcint_integrals.F90:985-986): `half` is metric^(-1/2) (naux x naux, symmetric).  The GEMM
 * runs nu-slab by nu-slab straight into the packed resident layout: no second device copy. */
int mqcb200_set_tensor_from_3c(void *handle, int slot, int n, int naux,
                               const double *three, const double *half);
/* metric_inverse_sqrt(metric, half, error) (mqc_libcint_integrals.F90:992-1038) on the device:
 * half = U s^(-1/2) U^T over the eigenvalues above null_threshold (the reference's 1e-10), the rest
 * zeroed; *n_kept = surviving modes (may be NULL).  The eigendecomposition is a one-sided Jacobi
 * iteration on the GPU (the reference calls LAPACK dsyev).  Fails with the reference's message
 * "density fitting: the auxiliary metric is singular" when no mode survives. */
int mqcb200_metric_inverse_sqrt(void *handle, int naux, const double *metric,
                                double null_threshold, double *half, int *n_kept);
/* The last two stages of build_df_tensor in one call: metric^(-1/2) on the device, then
 * b = three . half slab by slab into `slot`.  half_out (naux x naux, host; may be NULL) receives
 * metric^(-1/2) -- the gradient (mqcb200_df_gradient_densities) needs it again. */
int mqcb200_build_df_tensor(void *handle, int slot, int n, int naux, const double *three,
                            const double *metric, double null_threshold, double *half_out);
/* The same GEMM fed by the caller slab by slab, so that three(nao*nao, naux) -- 115 GB for the
 * 200-atom case -- never has to exist: the integral code produces (mu nu|P) for a block of nu
 * (all mu, all P) and hands it over.
 *   begin: allocates the packed slab [q_begin, q_begin+q_count) of `slot` and takes half (host).
 *   push : three_cols points at element (mu=0, nu=nu_begin, P=0); consecutive auxiliary functions
 *          are ld_aux doubles apart (nao*nao inside a full three, nao*nu_count for a compact
 *          block).  nu_begin is a multiple of 16 and nu_count too, unless the block ends at nao.
 *          Every nu must be pushed exactly once (in any order, by any rank).  Blocking.
 *   end  : the tensor becomes usable.
 * With a communicator active (whole-molecule builds sharded by auxiliary index) the three calls are
 * collective in the sense that every rank makes them, but each rank pushes only ITS share of the
 * nu blocks: a rank whitens its blocks for ALL auxiliary rows and the GEMM's epilogue stores every
 * row directly into the packed tensor of the rank that owns it, over NVLink peer memory -- the
 * all-to-all from mu-nu slabs to Q slabs costs no separate pass.  A rank whose push failed must still
 * call mqcb200_whiten_end: the call fails on EVERY rank then, and no rank's tensor becomes usable. */
int mqcb200_whiten_begin(void *handle, int slot, int n, int naux_total, int q_begin, int q_count,
                         const double *half);
int mqcb200_whiten_push(void *handle, int slot, int nu_begin, int nu_count,
                        const double *three_cols, long long ld_aux);
int mqcb200_whiten_end(void *handle, int slot);
/* Synthetic tensor generated on the device from a counter-based generator that
 * metalquicha_b200/synth.py reproduces bit-for-bit on the host (benchmarks at
 * sizes whose full-square host tensor does not fit in RAM). */
int mqcb200_synth_tensor(void *handle, int slot, int n, int naux_total,
                         int q_begin, int q_count, uint64_t seed, double scale);
int mqcb200_clear_tensor(void *handle, int slot);
/* Shape of the tensor resident on a slot (all zero when empty): lets a binding check that the
 * operands it is about to pass (n x n, no size travels with the build calls) belong to the
 * tensor that is resident -- a stale tensor from another fragment would otherwise be read out
 * of bounds.  Any output pointer may be NULL. */
int mqcb200_tensor_shape(void *handle, int slot, int *n, int *naux_total, int *q_begin, int *q_count);
/* Packed device bytes held by a slot (0 when empty). */
int mqcb200_tensor_bytes(void *handle, int slot, size_t *bytes);

/* ---- the Fock build ------------------------------------------------------
 * mqcb200_build_fock == build_fock_df(h, b, density, coeff, n_occ, fock, k_scale, j_scale)
 * (mqc_libcint_rhf.f90:1576-1646):  F = H + j_scale*J - (0.5*k_scale)*K with
 * J = sum_P (B_P.D) B_P and K = 2 sum_P (B_P C_occ)(B_P C_occ)^T.  `coeff` has
 * leading dimension ldc >= n and at least n_occ columns; only (:,1:n_occ) is
 * read (:1618).  The optional Fortran arguments become plain doubles: pass 1.0
 * for an absent k_scale / j_scale.  n_occ may equal n (pseudo-orbital guess
 * build, mqc_libcint_rhf.f90:1402-1405).  With n_occ == 0 K is zero. */
int mqcb200_build_fock(void *handle, int slot, const double *h, const double *density,
                       const double *coeff, int ldc, int n_occ,
                       double k_scale, double j_scale, double *fock);
/* The two contractions alone (either output may be NULL to skip it). K carries
 * the restricted factor 2 of mqc_libcint_rhf.f90:1637. */
int mqcb200_build_jk(void *handle, int slot, const double *density,
                     const double *coeff, int ldc, int n_occ, double *j, double *k);
/* Two-spin build (conventions of
 * backends/cuest/backend/mqc_cuest_scf.f90:48-57, :826-838): J from the total
 * density, K_sigma = sum_P (B_P C_sigma)(B_P C_sigma)^T with no factor 2.  An
 * empty BETA channel (n_beta == 0) is skipped and k_beta left untouched
 * (mqc_cuest_integrals.f90:1694-1701); an empty alpha channel gives k_alpha = 0. */
int mqcb200_build_jk_uhf(void *handle, int slot, const double *density_total,
                         const double *coeff_a, int lda, int n_alpha,
                         const double *coeff_b, int ldb, int n_beta,
                         double *j, double *k_alpha, double *k_beta);
/* F_sigma = H + J[Da+Db] - k_scale*K[C_sigma]  (shape of build_fock_uhf,
 * mqc_libcint_rhf.f90:1648-1681, with the fitted K). */
int mqcb200_build_fock_uhf(void *handle, int slot, const double *h,
                           const double *density_total,
                           const double *coeff_a, int lda, int n_alpha,
                           const double *coeff_b, int ldb, int n_beta,
                           double k_scale, double *fock_a, double *fock_b);
/* G = J[density] + ka*K[Ca] + kb*K[Cb] with K[C] = sum_P (B_P C)(B_P C)^T (no factor 2):
 * a Coulomb term plus two weighted symmetric exchange terms in one call (round 1 assembled the
 * operators below from it; they now have their own direct kernels).  A channel with n == 0 is
 * skipped. */
int mqcb200_build_g_two_factor(void *handle, int slot, const double *density,
                               const double *coeff_a, int lda, int n_a,
                               const double *coeff_b, int ldb, int n_b,
                               double ka, double kb, double *g);

/* mqcb200_response_operator == response_operator_df(b, x, c_occ, dtilde, g, k_scale)
 * (backends/libcint/mqc_libcint_cphf.F90:499-566) with b resident:
 *   g = J[dtilde] - (k_scale/2) * sum_P [(B_P X)(B_P C)^T + (B_P C)(B_P X)^T]
 * computed in that direct rank-2 form -- ONE half-transform of the stacked [X | C] and a
 * SYR2K-form accumulation -- so a small trial vector X keeps full relative accuracy (no
 * difference of squares).  x and c_occ are n x n_occ with leading dimensions ldx, ldc.
 * mqcb200_fitted_potential_general == fitted_potential_general(b, dens, g, k_scale)
 * (mqc_libcint_cphf.F90:568-616):  g = J[D] - (k_scale/2) * sum_P B_P D B_P  for the symmetric,
 * otherwise arbitrary D the reference requires; same kernels with X = D, C = 1 (n^3 naux). */
int mqcb200_response_operator(void *handle, int slot, const double *x, int ldx,
                              const double *c_occ, int ldc, int n_occ,
                              const double *dtilde, double k_scale, double *g);
int mqcb200_fitted_potential_general(void *handle, int slot, const double *dens,
                                     double k_scale, double *g);

/* ---- device-resident SCF of a fragment ----------------------
 * mqcb200_scf_fragment == run_libcint_rhf(mol, nelec, max_iter, energy_tol, density_tol, ...,
 * aux=aux, diis_vectors, guess) (mqc_libcint_rhf.f90:321-680) minus the integral generation: the
 * caller passes the core Hamiltonian and the overlap (n x n, host) and the fitted tensor is the
 * one resident on `slot`.  Everything between two Fock builds -- commutator X^T(FDS-SDF)X
 * (:1326-1352), DIIS (src/methods/mqc_diis.f90), F' = X^T F X + eigendecomposition + C = X C'
 * (:1464-1489), D = 2 C_occ C_occ^T (src/scf/mqc_scf_common.f90:84-96), the convergence test
 * (:626-636) -- runs on the GPU; per iteration only a few scalars cross PCIe (cf.
 * backends/cuest/backend/mqc_cuest_scf.f90:444-553).  Fragment-sized problems only (n <= 80,
 * n_occ <= 64); anything larger is refused, not run on the host.
 *   guess: 0 = core Hamiltonian, 1 = generalised Wolfsberg-Helmholz (the reference's default);
 *   diis_vectors: 0..8 (reference default 8), 0 = plain iteration;
 *   outputs: electronic energy of the FINAL rebuild (:646-649), iterations, converged flag,
 *   n_mo (orbitals surviving the 1e-7 overlap threshold, mqc_scf_common.f90:27), and -- each
 *   may be NULL -- coeff (n x n_mo), orbital_energies (n_mo), density (n x n),
 *   e_history (max_iter: electronic energy of every iteration's build).
 * Errors carry the reference's messages ("SCF: overlap matrix is singular", ...). */
int mqcb200_scf_fragment(void *handle, int slot, const double *hcore, const double *overlap,
                         int n_electrons, int guess, int max_iter,
                         double energy_tol, double density_tol, int diis_vectors, double k_scale,
                         double *e_electronic, int *iterations, int *converged, int *n_mo,
                         double *coeff, double *orbital_energies, double *density,
                         double *e_history);
/* The same loop for ANY size (and for a tensor sharded over GPUs: every rank then runs the same
 * bit-identical step around the exchanged Fock matrix): mqcb200_scf takes the one-CTA fragment route
 * when the problem fits it and otherwise keeps every matrix on the GPU with the general Fock-build
 * kernels, the batched DMMA GEMM for the dense algebra, and a one-sided Jacobi iteration on the shifted
 * F' = Y^T F Y (Y = the previous orbitals) in place of dsyev.  Per iteration the host sees the DIIS
 * overlaps of the newest error vector (<= 8 doubles), the eigensolver's "rotated anything?" flags and
 * the energy / rms(dD) scalars.  Same arguments and outputs as mqcb200_scf_fragment. */
int mqcb200_scf(void *handle, int slot, const double *hcore, const double *overlap,
                int n_electrons, int guess, int max_iter,
                double energy_tol, double density_tol, int diis_vectors, double k_scale,
                double *e_electronic, int *iterations, int *converged, int *n_mo,
                double *coeff, double *orbital_energies, double *density, double *e_history);
/* The same for a BATCH of fragments of one kind, driven in lock-step: the slot holds n_fragments
 * tensors of equal size back to back (set it with naux = n_fragments * naux_per_fragment: fragment f
 * owns auxiliary slabs [f*naux_per_fragment, (f+1)*naux_per_fragment)), hcore_all / overlap_all are
 * (n, n, n_fragments), and every kernel of an iteration runs once with the fragment index on a grid
 * axis -- six launches per iteration whatever the batch size, one SCF-step CTA per fragment.  This is
 * how an MBE run feeds the GPU: the fragments of one n-mer level share (n, naux, n_occ)
 * (src/fragmentation/mbe: the work items the reference hands to its workers one at a time).
 * Outputs are per fragment: e_electronic, iterations, converged (1 yes, 0 not within max_iter,
 * -1 the overlap threshold left fewer orbitals than occupied), n_mo; coeff_all (n, n, n_fragments),
 * orbital_energies_all (n, n_fragments) and density_all (n, n, n_fragments) may be NULL. */
int mqcb200_scf_fragment_batch(void *handle, int slot, int n_fragments,
                               const double *hcore_all, const double *overlap_all,
                               int n_electrons, int guess, int max_iter,
                               double energy_tol, double density_tol, int diis_vectors,
                               double k_scale, double *e_electronic, int *iterations,
                               int *converged, int *n_mo, double *coeff_all,
                               double *orbital_energies_all, double *density_all);
/* Iterations queued on the GPU between two looks at the convergence flag (default 1; queued
 * iterations after convergence are no-ops on the device). */
int mqcb200_set_scf_check_every(void *handle, int iterations);

/* ---- DF two-electron gradient: the two densities ---------------
 * The contraction half of df_two_electron_gradient + add_exchange_channel
 * (backends/libcint/mqc_libcint_gradient.f90:1545-1812): everything between the energy-side
 * integrals and the derivative integrals, i.e.
 *     rho = J^-1 g,  f^P = (J^-1 e)^P with e^P_ij = C^T (mu nu|P) C,
 *     gamma(:, :, P) = rho_P D - sum_spin w_s C f^{s,P} C^T           (nao, nao, naux)
 *     omega          = -1/2 rho rho^T + sum_spin (w_s/2) sum_ij f^{s,P}_ij f^{s,Q}_ij   (naux, naux)
 * with w = 2*exx_fraction for the single closed-shell channel and exx_fraction for each of the
 * two unrestricted ones (unrestricted != 0; a channel with n == 0 is skipped), exx_fraction == 0
 * skipping the exchange assembly, with_coulomb == 0 dropping both Coulomb shapes.  Formed from the
 * tensor RESIDENT on `slot` (the whitened B = (mu nu|P) . J^-1/2) and `half` = J^-1/2 (naux x naux,
 * the matrix metric_inverse_sqrt returned when the tensor was built), so neither (mu nu|P) nor the
 * metric is regenerated.  The derivative integrals (three_centre_deriv, two_centre_deriv) and the
 * final per-atom sums (:1738-1770) stay with the caller.  gamma and omega are host arrays. */
int mqcb200_df_gradient_densities(void *handle, int slot, const double *half,
                                  const double *total_density,
                                  const double *orbitals, int lda, int n_occupied,
                                  const double *orbitals_beta, int ldb, int n_occupied_beta,
                                  int unrestricted, double exx_fraction, int with_coulomb,
                                  double *gamma, double *omega);

/* E = 1/2 sum D (H + F)  (electronic_energy, mqc_libcint_rhf.f90:1691-1697),
 * evaluated on the device from the operands of the last mqcb200_build_fock. */
int mqcb200_last_energy(void *handle, double *e_elec);

/* Device-resident variant: all pointers are DEVICE pointers on the engine's
 * GPU (n x n, ld == n; coeff n x n_occ with ld == n).  No host<->device copies;
 * the call is asynchronous on the engine's stream unless `sync` is non-zero. */
int mqcb200_build_fock_device(void *handle, int slot, const double *d_h,
                              const double *d_density, const double *d_coeff, int n_occ,
                              double k_scale, double j_scale, double *d_fock, int sync);

/* Two-spin device-resident variant (same conventions as mqcb200_build_fock_uhf). */
int mqcb200_build_fock_uhf_device(void *handle, int slot, const double *d_h,
                                  const double *d_density_total,
                                  const double *d_coeff_a, int n_alpha,
                                  const double *d_coeff_b, int n_beta, double k_scale,
                                  double *d_fock_a, double *d_fock_b, int sync);

/* ---- multi-GPU (whole-molecule builds sharded by auxiliary index) ---------
 * New behaviour relative to the reference, which never splits one Fock build
 * (SURVEY 2.2).  One process per GPU; each holds a shard set with
 * mqcb200_set_tensor_shard; every build ends in one sum exchange of [J;K]: a single
 * kernel over NVLink peer memory (buffers mapped with CUDA IPC at comm_init -- or addressed
 * directly with peer access when several ranks are threads of one process), with the
 * NCCL all-reduce as fallback (MQCB200_P2P_ALLREDUCE=0 forces it).  A rank that waits longer than
 * MQCB200_XGPU_TIMEOUT_S (default 30) for a peer fails that build (an asynchronous build's failure
 * is returned by the next call on the handle); the next sharded build resynchronises the ranks.  The 128-byte id is created on rank 0 and carried to the other
 * ranks by the host program's own transport (MPI bcast in the reference,
 * src/parallel/mqc_bcast.f90). */
int mqcb200_comm_unique_id(char id[128]);
int mqcb200_comm_init(void *handle, int n_ranks, int rank, const char id[128]);
int mqcb200_comm_destroy(void *handle);

/* ---- fragment FIFO (fragmented MBE/GMBE runs, one fragment per GPU) -------
 * Same semantics as queue_t (src/fragmentation/common/mqc_work_queue.f90:10-57):
 * a FIFO of int64 ids, pop returns has_item=0 and id=-1 when drained.  Thread
 * safe, so one host thread per GPU can pull from it. */
int mqcb200_queue_create(const int64_t *ids, int64_t count, void **queue);
int mqcb200_queue_pop(void *queue, int64_t *id, int *has_item);
int mqcb200_queue_is_empty(void *queue, int *is_empty);
int mqcb200_queue_destroy(void *queue);

/* ---- instrumentation -------------------------------------------------------
 * With profiling on, every phase of every build is bracketed by CUDA events on the
 * engine's stream.  mqcb200_last_timings synchronises the stream, returns the per-
 * phase sums (ms) over all builds since the previous call, and clears them -- so a
 * loop of asynchronous builds can be timed without a host sync per build.
 * mqcb200_last_launches: kernels launched by the last build. */
int mqcb200_set_profiling(void *handle, int on);
int mqcb200_last_timings(void *handle, double ms[MQCB200_NUM_TIMERS]);
int mqcb200_last_launches(void *handle, int *n_kernels);
/* 1 when the last build took its Coulomb vector from the half-transform (the density was
 * verified on the device to be f*C*C^T), 0 when it made the general pass over the tensor. */
int mqcb200_last_gamma_fused(void *handle, int *fused);
/* Stream time (ms, host gathers included) and PCIe bytes of the last mqcb200_set_tensor[_shard]:
 * only the lower triangles cross the bus, double-buffered ("timed separately"). */
int mqcb200_last_set_tensor(void *handle, double *ms, double *h2d_bytes);
/* Device time (ms) of the whitening GEMM kernels of the last whitening (from_3c, build_df_tensor or
 * begin/push/end on this rank) and their flop count 2*rows*naux*L ("timed separately", not part
 * of builds/sec); device time and Jacobi sweeps of the last metric^(-1/2). */
int mqcb200_last_whiten(void *handle, double *ms, double *flops);
int mqcb200_last_metric(void *handle, double *ms, int *sweeps);

#ifdef __cplusplus
}
#endif
#endif /* MQCB200_H */
