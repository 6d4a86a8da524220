/*
 * smoqyelph_b200.h -- C ABI of libsmoqyelph_b200.so
 *
 * B200 (sm_100a) implementation of the linear-scaling electron-phonon hot path of SmoQyElPhQMC.jl.
 * The reference has no FFI layer: its seam is Julia multiple dispatch.  Each entry
 * point below names the reference method (file:line under /root/reference) whose body it replaces;
 * julia/SmoQyElPhB200.jl and INTEGRATION.md show the `ccall` binding a maintainer would add.
 *
 * Conventions
 *  - Every function returns 0 on success, non-zero on failure; sq_last_error() gives the message.
 *    Status 3 = numerical instability (NaN / non-finite residual, Lanczos coefficient or action), the
 *    condition the reference's callers turn into a rejected update; 1 = anything else (bad argument,
 *    CUDA / NCCL error, watchdog time-out of a resident kernel), which callers must NOT swallow.
 *    The Julia shim raises on non-zero so the reference's try/catch "numerical instability =>
 *    reject the update" semantics are preserved (src/EFAPFFHMCUpdater.jl:168-187).  CG
 *    non-convergence is NOT an error: iters == maxiter is returned (ConjugateGradient.jl:248).
 *  - All array arguments are dense, column-major HOST arrays owned by the caller, exactly as Julia
 *    holds them: space-time vectors are (Ltau x N) Complex{Float64} with tau fastest; V is
 *    (N x Ltau), t is (Nh x Ltau), x / p / dSdx are (Nph x Ltau) Float64.  Index tables are Int64
 *    and 1-BASED (Julia's), converted inside the library.  sq_complex = interleaved (re, im).
 *  - Entry points with the suffix _dev take DEVICE pointers in the library's internal layout
 *    (site fastest: element (l, i) at i + l*N) and never synchronise the host; they exist for
 *    pipelines that keep their vectors resident in HBM (bench.py `value`, multi-GPU drivers).
 *  - Calls are blocking (stream-synchronised on return) unless suffixed _dev.  One handle is used
 *    from one host thread at a time.  The library owns all device memory behind its handles.
 *  - There is no CPU fallback: every entry point fails with an error if no CUDA device is usable.
 */
#ifndef SMOQYELPH_B200_H
#define SMOQYELPH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } sq_complex;
typedef struct sq_fdm sq_fdm;       /* FermionDetMatrix{T,E}      src/FermionDetMatrix.jl:19-55,137-148 */
typedef struct sq_kpm sq_kpm;       /* KPMPreconditioner          src/KPMPreconditioner.jl:28-190        */
typedef struct sq_elph sq_elph;     /* ElectronPhononParameters + FermionPathIntegral fields the path reads */
typedef struct sq_pff sq_pff;       /* PFFCalculator              src/PFFCalculator.jl:9-53              */
typedef struct sq_hmc sq_hmc;       /* EFAPFFHMCUpdater           src/EFAPFFHMCUpdater.jl:9-99           */
typedef struct sq_greens sq_greens; /* GreensEstimator (solves + scalar measurements) src/Measurements/GreensEstimator.jl:10-175 */

enum { SQ_OP_M = 0, SQ_OP_MT = 1, SQ_OP_MTM = 2, SQ_OP_MMT = 3 };

/* ---- library ------------------------------------------------------------------------------- */
const char *sq_last_error(void);
int sq_version(void);
int sq_device_count(int *count);

/* ---- FermionDetMatrix ------------------------------------------------------------------------ */
/* SymFermionDetMatrix(fpi; maxiter, tol) / AsymFermionDetMatrix(...): src/FermionDetMatrix.jl:66-111,159-204.
 * nt (2 x Nh) is the PERMUTED neighbour table, perm (Nh) maps checkerboard index -> original hopping
 * index, color_lo/hi (ncolors) the inclusive 1-based bond ranges of each colour -- the outputs of
 * Checkerboard.checkerboard_decomposition! (:95-97); the library never recomputes the colouring. */
int sq_fdm_create(sq_fdm **out, int sym, int64_t Ltau, int64_t N, int64_t Nh, const int64_t *nt,
                  const int64_t *perm, int64_t ncolors, const int64_t *color_lo, const int64_t *color_hi,
                  double tol, int64_t maxiter, int device);
int sq_fdm_destroy(sq_fdm *f);
/* update!(fdm, fpi): src/FermionDetMatrix.jl:208-236.  V (N x Ltau), t (Nh x Ltau, original order). */
int sq_fdm_update(sq_fdm *f, const double *V, const double *t, double dtau);
/* mul_M! :385/:430, mul_Mt! :484/:528, mul_MtM! :329, mul_MMt! :357 (in == out allowed). */
int sq_fdm_mul(sq_fdm *f, int op, sq_complex *out, const sq_complex *in);
/* ldiv!(x, fdm, b; preconditioner, maxiter, tol): src/FermionDetMatrix.jl:248-267 + cg_solve!
 * src/IterativeSolvers/ConjugateGradient.jl:93-249.  zero_start != 0 <=> `x === b` (x is output only).
 * kpm may be NULL (preconditioner = I).  If kpm != NULL and lanczos_start != NULL the preconditioner is
 * refreshed first (update_preconditioner!, KPMPreconditioner.jl:554) with that N-vector as the randn!
 * start; lanczos_start == NULL draws it from the library's Philox stream. */
int sq_fdm_cg(sq_fdm *f, sq_complex *x, const sq_complex *b, int zero_start, sq_kpm *kpm, int refresh_kpm,
              const double *lanczos_start, double tol, int64_t maxiter, int64_t *iters, double *eps);
/* nrhs systems M^T M x_j = b_j at once (X, B: (Ltau N) x nrhs column-major, like the GR / Rt arrays of the GreensEstimator): the loop
 * over random vectors of update_greens_estimator! (src/Measurements/GreensEstimator.jl:152-169) as ONE lock-step solve.  Every system keeps
 * its own scalars, convergence test and iteration count (iters[j], eps[j]) -- the recurrence per system is ldiv!'s.  With an active
 * preconditioner on one GPU all kernels are batched over the systems; otherwise the systems are solved one by one. */
int sq_fdm_cg_batch(sq_fdm *f, sq_complex *X, const sq_complex *B, int64_t nrhs, int zero_start, sq_kpm *kpm, int refresh_kpm,
                    const double *lanczos_start, double tol, int64_t maxiter, int64_t *iters, double *eps);
/* read back expnΔτV (Ltau x N), coshΔτt, sinhΔτt (Ltau x Nh) in the reference's layout (tests) */
int sq_fdm_get_coefficients(sq_fdm *f, double *expV, double *cosh_t, double *sinh_t);
/* device-resident variants (internal layout, no host sync) */
int sq_fdm_mul_dev(sq_fdm *f, int op, void *d_out, const void *d_in);
int sq_fdm_cg_dev(sq_fdm *f, void *d_x, const void *d_b, int zero_start, sq_kpm *kpm, double tol,
                  int64_t maxiter, int64_t *iters, double *eps);
/* kernel configuration of the fused matvec (slices per CTA, threads per CTA; 0 = autotune) */
int sq_fdm_set_tuning(sq_fdm *f, int slab, int threads);
int sq_fdm_get_tuning(sq_fdm *f, int *slab, int *threads, int *path);   /* path: 0 generic fused, 1 global passes, 2 fast fused, 3 register path */
int sq_fdm_set_fast_path(sq_fdm *f, int enable);                       /* 0 generic, 1 fast fused kernel (Sym, <= 8 colours), 2 (+ 256 S) register path (rectangular lattices) */
int sq_fdm_stream(sq_fdm *f, void **cuda_stream);
/* measurement aid (bench.py, tools/): `reps` back-to-back launches of op on device vectors, timed with CUDA events on the
 * library stream; *us_per_launch = elapsed / reps (L2-hot: the working set of the named configs stays in L2).  If
 * d_flush != NULL a write of flush_bytes to d_flush precedes every launch (L2-cold); the time of the same flushes alone is
 * subtracted.  op 102 = M^T M of the register path on vectors in its native order (the kernel the CG solver runs). */
int sq_fdm_time_mul(sq_fdm *f, int op, void *d_out, const void *d_in, int reps, void *d_flush, int64_t flush_bytes,
                    double *us_per_launch);
/* tau-slab partitioning over the GPUs of one node (the reference is single-process).  Rank g of `world`
 * produces the contiguous slices [lo, hi) of every vector; arrays stay full length and sq_fdm_mul_dev / sq_fdm_cg_dev
 * exchange the one-slice halos with the ring neighbours over NCCL (libnccl is dlopen'ed).  id128 comes from
 * sq_nccl_unique_id on rank 0 and is broadcast by the host (MPI / torch.distributed).  Only the slices [lo, hi) of
 * the outputs are defined; unpreconditioned CG only. */
int sq_nccl_unique_id(char *out128);
int sq_fdm_init_slab(sq_fdm *f, int rank, int world, const char *id128);
/* Optional, after sq_fdm_init_slab: peer-mapped mailboxes (CUDA IPC) for the resident multi-GPU CG -- the whole solve is one
 * launch per rank, dot products and boundary slices travel as device-initiated NVLink stores instead of NCCL calls.
 * create returns this rank's 64-byte IPC handle; the host gathers the handles of all ranks (rank-major, world x 64 bytes)
 * and passes them to open.  Without mailboxes (or for lattices without the register path) the NCCL loop runs. */
int sq_fdm_mailbox_create(sq_fdm *f, char *handle64);
int sq_fdm_mailbox_open(sq_fdm *f, const char *handles64);
/* One Markov chain over several GPUs ("sharded solve"), after sq_fdm_init_slab (+ mailboxes): every rank keeps the FULL state and runs
 * everything but the CG solves redundantly and deterministically (same seeds => same chain on every rank); each unpreconditioned solve is
 * partitioned into tau-slabs and its solution broadcast back, so PFFCalculator, EFAPFFHMCUpdater and GreensEstimator work unchanged on
 * full-length vectors.  With a KPM preconditioner the solve stays local (replicated). */
int sq_fdm_set_sharded_solve(sq_fdm *f, int enable);
int sq_fdm_set_slab_range(sq_fdm *f, int64_t lo, int64_t hi);      /* single-process testing of the range logic */
int sq_fdm_get_slab(sq_fdm *f, int64_t *lo, int64_t *hi, int *rank, int *world);
int64_t sq_fdm_launch_count(sq_fdm *f);
/* Diagnostic counters since creation (no reference counterpart; the reference logs through @warn): out[0..n) =
 *  0 CG solves, 1 by the whole-solve resident register kernel, 2 by the cooperative shared-memory kernel, 3 by the unpreconditioned
 *  launch loop, 4 preconditioned, 5 tau-slab NCCL loop, 6 tau-slab resident kernels, 7 spin-wait watchdogs fired, 8 numerical
 *  instabilities turned into rejections, 9 CG iterations, 10 / 11 preconditioner applies (register / shared-memory Chebyshev kernel),
 *  12 right-hand sides solved by the batched solver, 13 tau-slab preconditioned solves. */
int sq_fdm_stats(sq_fdm *f, int64_t *out, int n);

/* ---- KPMPreconditioner ----------------------------------------------------------------------- */
/* KPMPreconditioner(fdm; rng, rbuf, n, a1, a2): src/KPMPreconditioner.jl:198-284 (does NOT run the
 * first update; call sq_kpm_update). */
int sq_kpm_create(sq_kpm **out, sq_fdm *f, double rbuf, int64_t n, double a1, double a2);
int sq_kpm_destroy(sq_kpm *k);
/* seed of the library-drawn Lanczos start vectors (`rng` of KPMPreconditioner(fdm; rng, ...), :198; the shim passes rand(rng, UInt64)) */
int sq_kpm_set_seed(sq_kpm *k, uint64_t seed);
/* update_preconditioner!(P, fdm, rng): :554-597.  lanczos_start: N normals (NULL = library RNG). */
int sq_kpm_update(sq_kpm *k, const double *lanczos_start, int *active, double *bounds);
/* test hook: refresh B-bar and force the eigenvalue bounds (identical coefficients on both sides) */
int sq_kpm_set_bounds(sq_kpm *k, double emin, double emax);
int sq_kpm_get_orders(sq_kpm *k, int64_t *ncoef, int64_t *orders /* may be NULL to query ncoef */);
int sq_kpm_get_coefs(sq_kpm *k, int64_t l /* 0-based */, sq_complex *coefs);
/* ldiv!(u', P, u) for complex vectors: Sym :355-414, Asym :488-550 */
int sq_kpm_ldiv(sq_kpm *k, sq_complex *out, const sq_complex *in);
int sq_kpm_ldiv_dev(sq_kpm *k, void *d_out, const void *d_in);
/* U v / U^-1 v of FourierTransformer: src/FourierTransformer.jl:39-64 */
int sq_kpm_fourier(sq_kpm *k, sq_complex *v, int forward);

/* ---- electron-phonon model ------------------------------------------------------------------- */
/* The arrays the path reads from SmoQyDQMC's ElectronPhononParameters / TightBindingParameters:
 * PhononParameters Omega, Omega4, M (Nph); HolsteinParameters coupling_to_phonon, coupling_to_site,
 * alpha..alpha4 (Nhol) and ph_sym_form expanded per coupling; SSHParameters coupling_to_phonon
 * (2 x Nssh), the hopping each coupling modulates (inverse of hopping_to_couplings), alpha..alpha4;
 * bare on-site energy minus mu (N) and bare hopping (Nh, original order). */
int sq_elph_create(sq_elph **out, sq_fdm *f, double dtau, int64_t Nph, const double *Omega, const double *Omega4,
                   const double *M, int64_t Nhol, const int64_t *hol_phonon, const int64_t *hol_site,
                   const double *hol_a, const double *hol_a2, const double *hol_a3, const double *hol_a4,
                   const int32_t *hol_phsym, int64_t Nssh, const int64_t *ssh_phonon, const int64_t *ssh_hopping,
                   const double *ssh_a, const double *ssh_a2, const double *ssh_a3, const double *ssh_a4,
                   const double *V0, const double *t0);
/* V0 / t0 may be NULL at creation -- PFFCalculator(elph, fdm) (src/PFFCalculator.jl:30-33) sees neither TightBindingParameters nor the
 * FermionPathIntegral -- and be supplied here by the first call that does (hmc_update!, the global moves, update_chemical_potential!).
 * Only sq_elph_refresh_fdm / sq_elph_get_Vt need them; sample / action / force work on the operator state set by sq_fdm_update. */
int sq_elph_set_bare(sq_elph *e, const double *V0, const double *t0);
/* DispersionParameters of ElectronPhononParameters: disp_phonon (2 x Ndisp, 1-based) = dispersion_to_phonon, Omega / Omega4 per coupling.
 * Enters bosonic_action (hmc_update! :136,238 and the global moves) and the kick through eval_derivative_dispersive_action!
 * (src/EFAPFFHMCUpdater.jl:193).  Arithmetic un-vendored (SmoQyDQMC): restated from the published Hamiltonian, DESIGN.md section 2. */
int sq_elph_set_dispersion(sq_elph *e, int64_t Ndisp, const int64_t *disp_phonon, const double *Omega, const double *Omega4);
/* F (Nph x Ltau) = anharmonic + dispersive derivative of the bosonic action (the terms :190-193 add to the fermionic force; tests) */
int sq_elph_potential_derivative(sq_elph *e, double *F);
int sq_elph_destroy(sq_elph *e);
int sq_elph_set_x(sq_elph *e, const double *x);          /* (Nph x Ltau) */
int sq_elph_get_x(sq_elph *e, double *x);
/* x-mutations of the global moves, on the device (phonon indices 1-based, ranges inclusive):
 *   reflection_update!  src/reflection_update.jl:94   `@. x_i = -x_i`            -> sq_elph_scale_x(e, p, p, -1)
 *   swap_update!        src/swap_update.jl:95         `SmoQyDQMC.swap!(x_i, x_j)` -> sq_elph_swap_x(e, p_i, p_j)
 *   radial_update!      src/radial_update.jl:114      `@. x' = expγ * x'`         -> sq_elph_scale_x(e, first, last, exp(γ))
 * backup / restore implement the rejection branch (the reference undoes the mutation arithmetically, :152-166). */
int sq_elph_scale_x(sq_elph *e, int64_t p_first, int64_t p_last, double factor);
int sq_elph_swap_x(sq_elph *e, int64_t p_i, int64_t p_j);
int sq_elph_backup_x(sq_elph *e);
int sq_elph_restore_x(sq_elph *e);
int sq_elph_shift_mu(sq_elph *e, double dmu);             /* V += -mu' + mu: update_chemical_potential.jl:66-67 */
/* SmoQyDQMC.update!(fpi, elph, x, +1) followed by update!(fdm, fpi) (EFAPFFHMCUpdater.jl:152-153), on device */
int sq_elph_refresh_fdm(sq_elph *e);
int sq_elph_get_Vt(sq_elph *e, double *V, double *t);
int sq_elph_bosonic_action(sq_elph *e, double *Sb);

/* ---- PFFCalculator --------------------------------------------------------------------------- */
int sq_pff_create(sq_pff **out, sq_elph *e);                                   /* src/PFFCalculator.jl:30-53 */
int sq_pff_destroy(sq_pff *p);
/* seed of the library-drawn pseudofermion noise (the `rng` argument of sample_pseudofermion_fields!, :56); without it
 * sq_hmc_create derives one from the HMC seed */
int sq_pff_set_seed(sq_pff *p, uint64_t seed);
int sq_pff_set_exact_holstein(sq_pff *p, int flag);    /* 0 = reference behaviour, 1 = exact derivative */
/* sample_pseudofermion_fields!: :56-76.  R = the randn!(rng, Phi) draw (Ltau x N), NULL = library RNG. */
int sq_pff_sample(sq_pff *p, const sq_complex *R, double *Sf);
/* calculate_fermionic_action!: :79-116 */
int sq_pff_action(sq_pff *p, sq_kpm *kpm, const double *lanczos_start, double tol, int64_t maxiter,
                  double *Sf, int64_t *iters, double *eps);
/* calculate_derivative_fermionic_action!: :119-158.  dSdx (Nph x Ltau) is accumulated into (+=). */
int sq_pff_force(sq_pff *p, double *dSdx, sq_kpm *kpm, const double *lanczos_start, double tol, int64_t maxiter,
                 double *Sf, int64_t *iters, double *eps);
int sq_pff_get_fields(sq_pff *p, sq_complex *Phi, sq_complex *Psi, double *Lambda);   /* any may be NULL */
int sq_pff_set_Phi(sq_pff *p, const sq_complex *Phi);
/* holstein_shift_matrix.jl:47-153 on caller vectors; which: 0 mul_Λ!, 1 ldiv_Λ!, 2 mul_Λᵀ!, 3 ldiv_Λᵀ! */
int sq_pff_lambda_op(sq_pff *p, int which, sq_complex *out, const sq_complex *in);
/* mul_νRe∂M∂x! (fermion_det_matrix_dervative.jl:2/117) and mul_νRe∂Λ∂x! (holstein_shift_matrix.jl:156) */
int sq_pff_dM_dx(sq_pff *p, double *F, double nu, const sq_complex *u, const sq_complex *v);
int sq_pff_dLambda_dx(sq_pff *p, double *F, double nu, const sq_complex *up, const sq_complex *u);

/* ---- EFAPFFHMCUpdater ------------------------------------------------------------------------ */
/* EFAPFFHMCUpdater(; electron_phonon_parameters, Nt, dt, eta, delta): src/EFAPFFHMCUpdater.jl:40-72 */
int sq_hmc_create(sq_hmc **out, sq_pff *p, int64_t Nt, double dt, double eta, double delta, uint64_t seed);
int sq_hmc_destroy(sq_hmc *h);
/* hmc_update!: :102-279.  randoms == NULL draws from the library's Philox stream; otherwise the
 * stream documented in DESIGN.md (same order as oracle/ref_c.c ref_hmc_update) is consumed.
 * info[8] = iters_avg, dH, Sf0, Sf1, Sb0, Sb1, K0, K1. */
int sq_hmc_update(sq_hmc *h, sq_kpm *kpm, double tol_action, double tol_force, int64_t maxiter,
                  const double *randoms, int64_t nrandoms, int *accepted, double *info);
/* re-key the Philox stream of the next trajectories (the Julia shim passes rand(rng, UInt64) before every hmc_update!, so the
 * caller's `rng` (:107) determines the trajectory call after call) */
int sq_hmc_set_seed(sq_hmc *h, uint64_t seed);
/* reason of the last forced rejection (the reference's `@warn "... rejecting update"`, :181,226), "" if the last update was stable */
const char *sq_hmc_last_reject(sq_hmc *h);
/* EFA pieces for parity tests: SmoQyDQMC initialize_momentum!, kinetic_energy, evolve_eom! */
int sq_hmc_init_momentum(sq_hmc *h, const double *R, double *p, double *K);
int sq_hmc_kinetic(sq_hmc *h, const double *p, double *K);
int sq_hmc_evolve(sq_hmc *h, double *x, double *p, double dt);

/* ---- GreensEstimator ------------------------------------------------------------------------- */
int sq_greens_create(sq_greens **out, sq_fdm *f, int64_t Nrv, uint64_t seed);   /* GreensEstimator.jl:63-118 */
int sq_greens_destroy(sq_greens *g);
int sq_greens_set_seed(sq_greens *g, uint64_t seed);      /* re-key the stream of the library-drawn random vectors (`rng` of update_greens_estimator!, :128) */
/* update_greens_estimator!: :125-175.  R (V x Nrv) unit-modulus vectors or NULL (library RNG);
 * warm start from the previous GR as in the reference. */
int sq_greens_update(sq_greens *g, sq_kpm *kpm, const sq_complex *R, double tol, int64_t maxiter, double *avg_iters);
int sq_greens_get(sq_greens *g, sq_complex *R, sq_complex *GR);
int sq_greens_set_GR(sq_greens *g, const sq_complex *GR);
/* measure_n :15, measure_double_occ :112, measure_Nsqrd :31 of src/Measurements/scalar_measurements.jl */
int sq_greens_measure(sq_greens *g, sq_complex *n, sq_complex *double_occ, sq_complex *Nsqrd);
/* measure_GΔ0!(correlation, greens_estimator, (a, b))  src/Measurements/GreensEstimator.jl:177-233: the translation-averaged
 * time-displaced Green's function G_ab(Δτ, Δr) from the current R, G R.  norb orbitals per unit cell (site = orbital + norb x
 * cell), dims[0..ndim) unit cells per direction (first fastest), a, b 1-based.  out: (Ltau + 1) x dims... complex, tau fastest
 * (the reference's CΔ0; the caller adds it to `correlation` with the tau axis moved last, :718-729). */
int sq_greens_measure_GD0(sq_greens *g, int norb, int ndim, const int64_t *dims, int a, int b, sq_complex *out);
/* The four-point contractions of src/Measurements/GreensEstimator.jl (no hopping weights tΔ / t0):
 *   kind 0  measure_GΔ0_GΔ0!  :236-388      kind 1  measure_GΔΔ_G00!  :391-467      kind 2  measure_G0Δ_GΔ0!  :470-606
 * orbitals[4] = (a, b, c, d) 1-based, r = the static displacements r1..r4 in unit cells, 4 x ndim (r1 first).  out as above;
 * the caller multiplies by `coef` and adds it to `correlation` (add_contraction_to_correlation!, :718-729). */
int sq_greens_measure_contraction(sq_greens *g, int kind, int norb, int ndim, const int64_t *dims, const int *orbitals, const int64_t *r,
                                  sq_complex *out);
/* The same contractions with the hopping weights tΔ, t0 of _measure_CΔ0! (:610-652) -- the building block of
 * measure_current_correlation! (src/Measurements/Correlations/current.jl:2-151).  tD, t0: real, Ltau x cells, tau fastest (the
 * PermutedDimsArray of fermion_path_integral.t built in make_measurements.jl:316-320), either may be NULL (= nothing); the
 * delta-function terms need both.  Real hoppings only, so conj_tΔ / conj_t0 have no effect and are not arguments. */
int sq_greens_measure_contraction_weighted(sq_greens *g, int kind, int norb, int ndim, const int64_t *dims, const int *orbitals, const int64_t *r,
                                           const double *tD, const double *t0, sq_complex *out);
/* Building blocks of the local measurements (src/Measurements/tight_binding_measurements.jl:43-133,
 * electron_phonon_measurements.jl: measure_holstein_energy, measure_ssh_energy):
 *   weighted_density: sum_{i,l} w[i,l] n(l,i),  n(l,i) = mean_rv (1 - GR[l,i,rv] Rt[l,i,rv]);  w (N x Ltau) real, site fastest
 *   weighted_bonds:   sum_{m,l} w[m,l] h(l; i_m->f_m) + conj(w[m,l]) h(l; f_m->i_m),  h(l; i->f) = mean_rv GR[l,i,rv] Rt[l,f,rv];
 *                     bonds 2 x nbonds 1-based sites, w (nbonds x Ltau) complex, bond fastest
 * The normalisation (1/N, 1/Ltau, ...) is part of the weights. */
int sq_greens_weighted_density(sq_greens *g, const double *w, sq_complex *out);
int sq_greens_weighted_bonds(sq_greens *g, int64_t nbonds, const int64_t *bonds, const sq_complex *w, sq_complex *out);
/* measure_n(greens_estimator, orbital)  src/Measurements/scalar_measurements.jl:2-12 */
int sq_greens_measure_n_orbital(sq_greens *g, int norb, int a, sq_complex *n);
/* measure_double_occ(greens_estimator, orbital)  src/Measurements/scalar_measurements.jl:98-109 (normalised by the total V, as there) */
int sq_greens_measure_double_occ_orbital(sq_greens *g, int norb, int a, sq_complex *d);
/* update_chemical_potential! minus the MuTuner scalar logic (stays in Julia): returns n, N^2 then
 * applies the new mu via sq_elph_shift_mu + sq_elph_refresh_fdm.  src/update_chemical_potential.jl:21-73 */

#ifdef __cplusplus
}
#endif
#endif
