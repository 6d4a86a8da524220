// sq_internal.h -- handle definitions behind the opaque types of include/smoqyelph_b200.h
#pragma once
#include "common.cuh"

#include <map>
#include <memory>
#include <string>
#include "smoqyelph_b200.h"

#define SQ_MAXC 32          // maximum number of checkerboard colours
#define SQ_MAXPART 4096     // maximum number of per-CTA partial sums of one reduction

// Device layout everywhere: site fastest, element (l, i) at i + l*N  ("[l][i]").
// A time slice is a contiguous N-vector, tau-slabs are contiguous ranges.

struct KParams {            // by-value kernel parameter block of the operator kernels
    int L, N, Nh, C, sym;
    int S;                  // slices owned by one CTA (fused kernels)
    int lb, le;             // slices [lb, le) this launch produces (whole axis, or this rank's tau-slab)
    int TX, TXshift;        // threads along the bond / site index (power of two)
    int clo[SQ_MAXC], chi[SQ_MAXC];
    int nunc0;              // sites not touched by colour 0 (fused middle step)
    const int2 *nt;         // Nh bonds (i, j), checkerboard order, 0-based
    const double2 *cs;      // [l][h] (cosh, sinh)
    const double *expV;     // [l][i]
    const int *unc0;
};

struct CgState {            // device-resident CG scalars, ping-ponged between iterations
    double rz_re, rz_im;    // r.z (or r.r)
    double normb;           // |b|
    double eps;             // |r|/|b| after the last completed iteration
    double tol;
    int iters, done;
};

// diagnostic counters behind sq_fdm_stats (which solver ran, how often a watchdog or a NaN check fired)
enum {
    SQ_STAT_CG_SOLVES = 0,      // CG solves entered
    SQ_STAT_CG_RESIDENT = 1,    // ... run by the whole-solve resident register kernel (k_cg_v3_resident1)
    SQ_STAT_CG_PERSIST_V2 = 2,  // ... by the cooperative shared-memory kernel (k_cg_persistent)
    SQ_STAT_CG_LOOP = 3,        // ... by the launch loop, unpreconditioned
    SQ_STAT_CG_PREC = 4,        // ... preconditioned (launch loop)
    SQ_STAT_CG_SLAB_NCCL = 5,   // tau-slab solves through the NCCL loop
    SQ_STAT_CG_SLAB_RESIDENT = 6, // tau-slab solves through the resident kernels + peer mailboxes
    SQ_STAT_WATCHDOG = 7,       // spin-wait watchdogs that fired (a stalled resident kernel)
    SQ_STAT_INSTABILITY = 8,    // numerical instabilities converted into rejected updates
    SQ_STAT_CG_ITERS = 9,       // CG iterations, all solves
    SQ_STAT_KPM_REG = 10,       // preconditioner applies with the register Chebyshev kernel
    SQ_STAT_KPM_SMEM = 11,      // ... with the shared-memory Chebyshev kernels
    SQ_STAT_CG_BATCHED = 12,    // right-hand sides solved by the batched (multi-RHS) solver
    SQ_STAT_CG_SLAB_PREC = 13,  // tau-slab preconditioned solves (frequency-sharded KPM, all-to-all)
    SQ_NSTATS = 16
};

struct sq_fdm {
    int device = 0;
    i64 stats[SQ_NSTATS] = {0};
    cudaStream_t stream = nullptr;
    int sym = 1;
    i64 L = 0, N = 0, Nh = 0, C = 0;
    std::vector<int> clo, chi;
    std::vector<int> h_perm;                 // checkerboard index -> original hopping (0-based)
    std::vector<int2> h_nt;
    DevBuf<int2> nt;
    DevBuf<int> perm, unc0;
    int nunc0 = 0;
    // fast path (fdm_v2.cu): shared-memory slot of every site, bonds as slot pairs
    DevBuf<int2> nts;
    DevBuf<int> slot;
    std::vector<int> h_slot;
    std::vector<int> h_abi_chk;              // internal bond index -> checkerboard index of the ABI tables
    int use_v2 = 0;                          // chosen by the autotuner / sq_fdm_set_fast_path
    int cs_uniform = 0;                      // (cosh, sinh) do not depend on tau (no SSH coupling): register-resident path
    int cs_coluni = 0;                       // ... and are equal for all bonds of one colour (uniform hopping): fdm_v3.cu
    // register path (fdm_v3.cu): rectangular lattice, one warp per slice
    int v3_ok = 0, v3_lxl = 0, v3_ry = 0, v3_cls[4] = {0, 0, 0, 0};
    int v3_kind = 0;                         // 0: square (v3_lxl = Lx / 4, v3_ry = rows per lane), 1: honeycomb (v3_lxl = L1, v3_ry = L2),
                                             // 2: chain (v3_lxl = site pairs per lane: N = 64 v3_lxl)
    int v3_pb_ok = 0;                        // a per-bond engine exists: (cosh, sinh) of every bond and slice in registers (SSH couplings)
    int v3_ncs = 0;                          // ... coefficient slots per lane and slice
    DevBuf<int> v3_csmap;                    // ... bond index of slot (q, lane)
    DevBuf<double2> v3_csn;                  // ... coefficients in slot order [l][q][lane], rebuilt with v3_expVn
    int v3_native_pb = -1;                   // engine family the native-order copies were prepared for
    // graph engine of the resident CG (fdm_v3.cu: V3Graph): ANY lattice with N <= 64 sites and <= 4 colours, padded to 64 sites
    int v3g_ok = 0;
    DevBuf<int> v3g_part, v3g_csmap;         // partner lane / value of (colour, value, lane); bond of coefficient slot (q, lane) or -1
    DevBuf<double> v3g_expVn;                // [l][64]
    DevBuf<double2> v3g_csn;                 // [l][8][32]
    DevBuf<double> v3g_x, v3g_r;             // [l][part][64]
    i64 v3g_version = -1;
    int v3_S = 3;                            // slices per CTA of the register path
    int use_v3 = 0;                          // stand-alone products (library-order vectors): chosen by timing
    int v3_cg = 0;                           // CG solves: register path whenever it applies (native order + resident kernel)
    DevBuf<double2> v3_ctn;                  // (cosh, tanh) per colour for the scaled rotations
    DevBuf<double> v3_expVn;                 // exp(-dtau V) in the native order of the register path
    DevBuf<double2> v3_x, v3_r;              // CG vectors in native order
    i64 v3_expv_version = -1;
    DevBuf<double> v3_halo;                  // boundary slices exchanged by the one-sum resident CG kernel
    DevBuf<char> v3_slots;                   // grid-sum slots of the resident CG kernel
    DevBuf<int> flag;                        // device scratch flags (4 ints: uniformity probe / grid barrier / abort)
    DevBuf<double> expV;                     // [l][i]
    DevBuf<double2> cs;                      // [l][h]
    DevBuf<double2> tmp1, tmp2, r, p, z;     // [l][i]
    DevBuf<double2> io1, io2;                // staging in the host (tau-fastest) layout
    DevBuf<double> iod1, iod2;               // real staging
    DevBuf<double> part;                     // 8 * SQ_MAXPART partial sums
    DevBuf<CgState> cg;                      // 2 states
    // multi-RHS solver workspace (cg_batch.cu), nrhs vectors each, allocated on first use
    DevBuf<double2> bt_r, bt_p, bt_z, bt_q, bt_zt;
    DevBuf<CgState> bt_st;
    DevBuf<double> bt_part;
    DevBuf<unsigned> bt_ticket;
    int bt_cap = 0;
    DevBuf<double2> prec_q;                  // M^T M p of the preconditioned solver when the p update is fused into the matvec (lazy)
    DevBuf<unsigned> cg_ticket;              // arrival counter of the last-block convergence test (zero between launches)
    CgState *h_cg = nullptr;                 // pinned
    double tol = 1e-6;
    i64 maxiter = 0;
    // fused-kernel configuration
    int path = 0;                            // 0 = slices staged in shared memory, 1 = global-memory passes
    int slab = 0, threads = 0;
    int tuned[3][6] = {{0}, {0}, {0}};         // per coefficient mode (general / tau-uniform / colour-uniform): valid, slab, threads, v2, v3, v3_S
    int manual_tuning = 0;
    int prec_iters_hint[2] = {0, 0};         // iterations of the last preconditioned solve (tight / loose tolerance): places the first read-back
    int num_sms = 148;
    size_t smem_optin = 0;
    i64 launches = 0;
    i64 coef_version = 0;                    // bumped by every operator refresh (KPM B-bar cache)
    // tau-slab partitioning (multi-GPU): this rank produces slices [slab_lo, slab_hi); arrays stay full length and the
    // one-slice halos are exchanged in place at their global index (slab.cu)
    int slab_lo = 0, slab_hi = 0;
    int rank = 0, world = 1;
    void *comm = nullptr;                    // ncclComm_t
    DevBuf<double> scal;                     // packed scalars for all-reduces
    // mailboxes of the multi-GPU resident CG (peer-mapped through CUDA IPC): own buffer + the peers' mappings
    DevBuf<char> mail;
    void *mail_ptr[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int mail_ready = 0;
    unsigned long long v3_it_base = 0;       // iteration count carried across multi-GPU solves (slot validity tags)
    // "sharded solve" mode (one Markov chain over several GPUs): every rank keeps the full state and runs everything but the CG
    // solves redundantly; a solve is partitioned into tau-slabs [shard_lo, shard_hi) and its solution all-gathered (slab.cu)
    int sharded = 0, shard_lo = 0, shard_hi = 0;
    int force_local = 0;                     // sharded chain: solve on this rank alone (right-hand sides distributed over the ranks, greens.cu)
    int tuned_shard[3][6] = {{0}, {0}, {0}};   // tuning cache of the slab range while the full range is active (and vice versa)

    KParams kparams(int S, int T) const;
    size_t vec_bytes() const { return (size_t)L * N * sizeof(double2); }
};

struct sq_kpm {
    sq_fdm *f = nullptr;
    int active = 0;
    double rbuf = 0.1, a1 = 1, a2 = 1;
    i64 nlanczos = 20;
    double bounds[2] = {0, 0};
    i64 ncoef = 0;
    std::vector<i64> order;
    std::vector<std::vector<double2>> coefs;
    // device
    DevBuf<double> Dbar;                     // N
    DevBuf<double2> csbar;                   // Nh
    DevBuf<double2> tw;                      // Ltau forward twiddles exp(-2 pi i k / L)
    DevBuf<double2> theta;                   // Ltau twist
    DevBuf<int> d_order, d_coef_off, d_freq_sched;   // per frequency (Ltau): order, offset into d_coefs; schedule
    DevBuf<double2> d_coefs;
    DevBuf<double> d_scale1;                 // per frequency: scalar applied by the FFT store when order == 1
    int nsched = 0;                          // frequencies with order > 1
    std::vector<int> h_sched;                // host copy of the schedule (longest recurrence first)
    i64 sched_version = 0;                   // bumped whenever the expansions (and with them the schedule) change
    // tau-slab mode (slab.cu): this rank's share of the schedule and the staging buffers of the two all-to-all exchanges
    DevBuf<int> d_sched_slab;
    int nsched_slab = 0;
    i64 sched_slab_version = -1;
    int sched_slab_lo = -1, sched_slab_hi = -1;
    DevBuf<double2> slab_a, slab_b, slab_recv;
    DevBuf<int> site_bond;                   // per-bond register Chebyshev kernel: (colour, site) -> internal bond index
    DevBuf<double2> ztmp;                    // [n][i] frequency-space scratch
    DevBuf<double> lan;                      // Lanczos alpha/beta read-back
    DevBuf<double> lan_start;                // N
    DevBuf<double2> cheb_ws;                 // workspace for the global-memory Chebyshev fallback
    std::vector<int> radices;
    i64 bbar_version = -1;
    int max_order = 0;
    uint64_t seed = 0x5eed, rng_counter = 0;  // Lanczos start vectors (sq_kpm_set_seed)
};

struct sq_elph {
    sq_fdm *f = nullptr;
    double dtau = 0;
    i64 Nph = 0, Nhol = 0, Nssh = 0;
    std::vector<double> h_M;
    DevBuf<double> x;                        // [l][p]
    DevBuf<double> x_backup;                 // copy taken by a global move, restored on rejection
    DevBuf<double> Om, Om4, M;
    DevBuf<int> fin;                         // isfinite(M[p])
    // Holstein couplings
    DevBuf<int> hol_ph, hol_site, hol_sym;
    DevBuf<double> ha;                       // a[c], a2[Nhol+c], a3[2Nhol+c], a4[3Nhol+c]
    DevBuf<int> site_ptr, site_cpl;          // CSR: site -> Holstein couplings (ascending coupling index)
    // SSH couplings
    DevBuf<int> ssh_p, ssh_pp;               // the phonon pair (p, p') of each coupling
    DevBuf<int> ssh_bond;                    // checkerboard bond index the coupling modulates
    DevBuf<double> sa;                       // 4 * Nssh
    DevBuf<int> bond_ptr, bond_cpl;          // CSR: checkerboard bond -> SSH couplings
    // force gather lists: phonon -> Holstein couplings, phonon -> signed SSH couplings (+-(c+1))
    DevBuf<int> ph_hol_ptr, ph_hol_cpl, ph_ssh_ptr, ph_ssh_cpl;
    // dispersive phonon couplings (DispersionParameters): pair (p, p'), reduced mass x Omega^2 and x Omega4^2 per coupling, and the gather
    // list phonon -> signed couplings (+-(c+1): + for the second phonon of the pair) of the derivative
    i64 Ndisp = 0;
    DevBuf<int> disp_p, disp_pp, ph_disp_ptr, ph_disp_cpl;
    DevBuf<double> disp_k2, disp_k4;
    DevBuf<double> V0, t0;                   // bare on-site energy (N), bare hopping (Nh, ORIGINAL order)
    DevBuf<double> V, t;                     // materialised only by sq_elph_get_Vt: [l][i], [l][h] original order
    bool any_phsym = false;
    bool t0_coluni = false;                  // bare hopping uniform inside every colour
    bool bare_set = false;                   // V0 / t0 known (sq_elph_create with non-NULL arrays, or sq_elph_set_bare)
};

struct sq_pff {
    sq_elph *e = nullptr;
    sq_fdm *owner = nullptr;                 // = e->f, kept for the destroy path (the parents may already be gone)
    DevBuf<double2> Phi, u, up, upp, w1, w2; // [l][i]
    DevBuf<double> Lam;                      // [l][i]
    DevBuf<double> F;                        // [l][p] force accumulator on the device
    DevBuf<double> HV, HL, SV;               // per-coupling force contributions [l][c]
    DevBuf<double> part;                     // reduction partials
    int exact_holstein = 0;
    uint64_t seed = 0x0ff1ce, rng_counter = 0;   // pseudofermion noise (sq_pff_set_seed; sq_hmc_create derives it from the HMC seed)
    bool seed_set = false;
};

struct sq_hmc {
    sq_pff *p = nullptr;
    sq_fdm *owner = nullptr;                 // = p->e->f
    i64 Nt = 0;
    double dt = 0, eta = 0, delta = 0;
    uint64_t seed = 0, counter = 0;
    DevBuf<double> x0, pm, dS;               // [l][p]
    DevBuf<double> Mt, wd;                   // [w][p]
    DevBuf<double2> fz, fw;                  // Fourier work arrays [w][p]
    DevBuf<double2> tw;                      // twiddles
    DevBuf<double> rnd;                      // random stream of one trajectory
    DevBuf<double> part;
    std::vector<int> radices;
    std::string last_reject;                 // reason of the last forced rejection (numerical instability), "" if none
};

struct sq_greens {
    sq_fdm *f = nullptr;
    i64 Nrv = 0;
    uint64_t seed = 0, counter = 0;
    DevBuf<double2> R, GR, MtR;              // Nrv vectors, [l][i] each
    DevBuf<double2> MtRb, Xb;                // M^T R (and, in a sharded chain, the start vectors) of the systems solved as a batch
    DevBuf<double2> wa, wb, wc, wt;          // work arrays of the correlation measurements (2 Ltau x cells)
    DevBuf<double> wreal;                    // weights of the local measurements
    DevBuf<double2> wcplx;
    DevBuf<int2> wbond;
    std::map<int, std::unique_ptr<DevBuf<double2>>> fft_tw;   // twiddles per transform length
    DevBuf<double> part;
};

// Philox (seed, stream) namespaces: the component tag sits in the top byte of the stream index, so that two components created
// with the same seed (the Python / Julia defaults) never draw the same normals (round-1 ADVICE: HMC, Greens and PFF all
// counted their streams from 0).
enum { SQ_RNG_HMC = 1, SQ_RNG_GREENS = 2, SQ_RNG_PFF = 3, SQ_RNG_KPM = 4 };
static inline uint64_t sq_rng_stream(int component, uint64_t counter) { return ((uint64_t)component << 56) | (counter & 0x00ffffffffffffffULL); }
// splitmix64 finaliser: derives a child seed from (seed, tag) -- chain seeds, component seeds
static inline uint64_t sq_mix_seed(uint64_t seed, uint64_t tag) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (tag + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// CG x / r update fused into the load phase of the forward tau-FFT of the preconditioner (fft.cu): the transform reads r anyway, so
//   alpha = (r.z) / (p.Ap);  x += alpha p;  r -= alpha q;  |r|^2 partials;  last CTA: eps, convergence test -> nxt
// ride along and one launch (and one pass over x, r, p, q) per iteration disappears.  All arrays are indexed by the batch (blockIdx.y).
struct FftCgUpdate {
    double2 *x, *r;             // updated in place (r is the transform's input)
    const double2 *p, *q;
    const CgState *cur;         // [batch]
    CgState *nxt;               // [batch]
    const double *pAp_part;     // [batch][pap_stride]
    int npart, pap_stride;
    double *rr_part;            // [batch][SQ_MAXPART]
    unsigned *ticket;           // [batch]
    int iter;
};

// ---- functions shared between translation units (all enqueue on f->stream) ------------------------
void fdm_mul_dev(sq_fdm *f, int op, double2 *out, const double2 *in, double *pAp_partials = nullptr,
                 int *npart = nullptr, const CgState *skip_if_done = nullptr);
void fdm_host_to_dev(sq_fdm *f, double2 *d_dst, const void *h_src);      // (Ltau x N) tau-fastest -> [l][i]
void fdm_dev_to_host(sq_fdm *f, void *h_dst, const double2 *d_src);
void fdm_transpose_real(sq_fdm *f, double *dst, const double *src, int rows_fast_src, int cols, bool to_host_layout);
void fdm_sweep_global(sq_fdm *f, double2 *u, int lo, int hi, bool inverse);
void fdm_scale_global(sq_fdm *f, double2 *u, bool inverse);
void fdm_cg_dev(sq_fdm *f, double2 *x, const double2 *b, bool zero_start, sq_kpm *kpm, double tol, i64 maxiter,
                i64 *iters, double *eps);
void kpm_ldiv_dev(sq_kpm *k, double2 *out, const double2 *in, const CgState *skip_if_done = nullptr);
int kpm_ldiv_dev_dot(sq_kpm *k, double2 *out, const double2 *in, const CgState *skip, const double2 *dot_with, double *dot_part);
struct FftCgUpdate;
int kpm_ldiv_dev_fused(sq_kpm *k, double2 *z, const FftCgUpdate &upd, double *dot_part);
void kpm_fft_cheb_batch(sq_kpm *k, double2 *out, const double2 *in, double2 *zt, int nrhs, size_t stride, const CgState *skip,
                        const double2 *dot_with, double *dot_part, int *npart, const FftCgUpdate *upd = nullptr);
void kpm_fourier_dev(sq_kpm *k, double2 *v, bool forward);
void kpm_cheb_apply(sq_kpm *k, double2 *z, const int *d_sched, int nsched, int nrhs, size_t rhs_stride, const CgState *skip);
void kpm_lanczos(sq_kpm *k, const double *h_start, const double *d_start, double *emin, double *emax);
void kpm_update(sq_kpm *k, const double *h_lanczos_start, const double *d_lanczos_start);
void kpm_set_bounds(sq_kpm *k, double emin, double emax);
void elph_refresh_fdm(sq_elph *e);
void elph_set_bare(sq_elph *e, const double *V0, const double *t0);
void elph_set_dispersion(sq_elph *e, i64 Ndisp, const i64 *disp_ph, const double *Om, const double *Om4);
void elph_add_potential_derivative(sq_elph *e, double *pm, const double *dS, double dt);   // p -= dt (dS + anharmonic + dispersive)
double elph_bosonic_action(sq_elph *e);
void elph_update_lambda(sq_elph *e, double *Lam);
void elph_lambda_op(sq_elph *e, int which, double2 *out, const double2 *in, const double *Lam);
void fdm_update_dev(sq_fdm *f, const double *dV, const double *dt_, double dtau);
double reduce_partials_host(sq_fdm *f, const double *d_part, int n);
int tau_fft_launch(cudaStream_t stream, const std::vector<int> &radices, int L, int N, double2 *out, const double2 *in, bool inverse,
                   bool twist, const double2 *tw, const double2 *theta, const double *scale1, const double2 *dot_with,
                   double *dot_part, const CgState *skip, size_t smem_limit);
void rng_fill_normal(double *d_out, size_t n, uint64_t seed, uint64_t stream, cudaStream_t s);
void rng_fill_uniform(double *d_out, size_t n, uint64_t seed, uint64_t stream, cudaStream_t s);
void fdm_select_tuning(sq_fdm *f);
// register-path fused kernels (fdm_v3.cu); nbatch > 1: a batch of vectors bstride elements apart (p.Ap partials bpart doubles apart,
// one CgState per vector in skip[])
int fdm_v3_launch(sq_fdm *f, int mode, int S, double2 *out, const double2 *in, double *part, const CgState *skip, bool native = false,
                  int nbatch = 1, size_t bstride = 0, int bpart = 0);
bool fdm_v3_supported(const sq_fdm *f, int S);
// small lattices (graph engine): whole unpreconditioned solve in the resident kernel; x (in / out) and r (in) in library order.
// Returns false when the problem does not qualify (the caller continues with its other solvers; x and r are untouched).
bool fdm_v3g_cg(sq_fdm *f, double2 *x, const double2 *r, bool zero_start, CgState *state, i64 maxiter);
// engine family of the register path: uniform engines (one cosh / tanh per colour) when the hoppings are colour-uniform and the lattice
// has one, otherwise the per-bond engines
inline bool fdm_v3_perbond(const sq_fdm *f) { return !(f->v3_ok && f->cs_coluni); }
int tau_fft_launch_batch(cudaStream_t stream, const std::vector<int> &radices, int L, int N, double2 *out, const double2 *in, bool inverse,
                         bool twist, const double2 *tw, const double2 *theta, const double *scale1, const double2 *dot_with,
                         double *dot_part, const CgState *skip, size_t smem_limit, int nbatch, size_t bstride, const FftCgUpdate *upd = nullptr);
// multi-RHS preconditioned CG (cg_batch.cu): nrhs systems M^T M x_j = b_j, vectors V elements apart
void fdm_cg_batch_dev(sq_fdm *f, double2 *X, const double2 *B, int nrhs, bool zero_start, sq_kpm *kpm, double tol, i64 maxiter, i64 *iters,
                      double *eps);
bool fdm_cg_batch_applicable(const sq_fdm *f, const sq_kpm *kpm, int nrhs);
void fdm_sync_if_alive(sq_fdm *f);       // stream-synchronise f if it has not been destroyed yet, else the device
void fdm_halo_exchange(sq_fdm *f, double2 *v);
void fdm_allreduce_sum(sq_fdm *f, double *d_buf, int count);
void fdm_cg_slab(sq_fdm *f, double2 *x, const double2 *b, bool zero_start, sq_kpm *kpm, double tol, i64 maxiter, i64 *iters, double *eps);
void fdm_cg_sharded(sq_fdm *f, double2 *x, const double2 *b, bool zero_start, sq_kpm *kpm, double tol, i64 maxiter, i64 *iters, double *eps);
void slab_set_sharded(sq_fdm *f, int enable);
void slab_broadcast_columns(sq_fdm *f, double2 *cols, size_t V, int ncols);    // column j from rank j % world (grouped ncclBroadcast)
void fft_radices(i64 n, std::vector<int> &rad);
void fft_make_twiddles(i64 n, std::vector<double2> &tw);
