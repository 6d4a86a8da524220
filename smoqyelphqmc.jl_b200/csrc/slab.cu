// slab.cu -- tau-slab partitioning of the space-time vector over the GPUs of one node.
//
// Rank g owns the contiguous slices [slab_lo, slab_hi).  M couples slice l to l-1 and M^T to l+1, so the fused
// M^T M kernel of a slab needs exactly one boundary slice from each ring neighbour; the wrap-around link carries the
// antiperiodic + sign, which the kernels already derive from the GLOBAL slice index.  Every rank keeps full-length
// arrays (6.5 MB per vector at the named size -- memory is not the constraint) and only produces its own slices, so a
// halo is simply the neighbour's boundary slice written at its global position:
//     send v[slab_lo]   -> previous rank (it is their v[hi]),      send v[slab_hi-1] -> next rank (their v[lo-1]).
// Per CG iteration: one halo exchange of p (2 x 16 N bytes each way) and two all-reduces of one double.
// The plumbing is NCCL (ncclSend/ncclRecv grouped, ncclAllReduce) on the library stream; libnccl is resolved at run
// time with dlopen so that the library has no link-time dependency and shares the NCCL already loaded by the host
// process (torch's), if any.  The reference has no counterpart: it is single-process (its MPI mode = independent
// chains, which needs no code here).
#include "sq_internal.h"

#include <dlfcn.h>

#include <cstring>

// minimal NCCL surface (types as in nccl.h)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSum = 0 };
enum { ncclChar = 0, ncclDouble = 8 };

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static void nccl_load() {
    if (g_nccl.lib) return;
    const char *names[] = {getenv("SQ_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n) continue;
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) throw SqError("tau-slab mode needs NCCL: could not dlopen libnccl.so.2 (set SQ_NCCL_LIB)");
#define SQ_SYM(field, name)                                                   \
    *(void **)(&g_nccl.field) = dlsym(g_nccl.lib, name);                      \
    if (!g_nccl.field) throw SqError(std::string("libnccl lacks symbol ") + name)
    SQ_SYM(GetUniqueId, "ncclGetUniqueId");
    SQ_SYM(CommInitRank, "ncclCommInitRank");
    SQ_SYM(CommDestroy, "ncclCommDestroy");
    SQ_SYM(Send, "ncclSend");
    SQ_SYM(Recv, "ncclRecv");
    SQ_SYM(AllReduce, "ncclAllReduce");
    SQ_SYM(GroupStart, "ncclGroupStart");
    SQ_SYM(GroupEnd, "ncclGroupEnd");
    SQ_SYM(Broadcast, "ncclBroadcast");
    SQ_SYM(GetErrorString, "ncclGetErrorString");
#undef SQ_SYM
}
#define SQ_NCCL(expr)                                                                                              \
    do {                                                                                                           \
        ncclResult_t _r = (expr);                                                                                  \
        if (_r != 0) throw SqError(std::string("NCCL error '") + g_nccl.GetErrorString(_r) + "' in " #expr);       \
    } while (0)

void slab_unique_id(char *out128) {
    nccl_load();
    ncclUniqueId id;
    SQ_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out128, id.internal, 128);
}

// balanced contiguous partition: the first L % world ranks get one extra slice
static void slab_range(int L, int world, int rank, int *lo, int *hi) {
    int base = L / world, extra = L % world;
    *lo = rank * base + std::min(rank, extra);
    *hi = *lo + base + (rank < extra ? 1 : 0);
}

void slab_set_range(sq_fdm *f, int lo, int hi) {
    SQ_REQUIRE(lo >= 0 && hi > lo && hi <= f->L, "slab range out of bounds");
    f->slab_lo = lo;
    f->slab_hi = hi;
    f->tuned[0][0] = f->tuned[1][0] = f->tuned[2][0] = 0;         // the best (slab, threads) depends on the number of local slices
    f->manual_tuning = 0;
}

void slab_init(sq_fdm *f, int rank, int world, const char *id128) {
    SQ_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank / world size");
    SQ_REQUIRE(world <= f->L, "more ranks than time slices");
    SQ_CUDA(cudaSetDevice(f->device));
    f->rank = rank;
    f->world = world;
    int lo, hi;
    slab_range((int)f->L, world, rank, &lo, &hi);
    slab_set_range(f, lo, hi);
    if (!f->scal.p) f->scal.alloc(16);
    if (world > 1) {
        nccl_load();
        ncclUniqueId id;
        memcpy(id.internal, id128, 128);
        ncclComm_t c;
        SQ_NCCL(g_nccl.CommInitRank(&c, world, id, rank));
        f->comm = (void *)c;
    }
}

// ---- sharded solve: full state on every rank, only the CG solves are partitioned --------------------------------------------
void slab_set_sharded(sq_fdm *f, int enable) {
    SQ_REQUIRE(f->world >= 1 && (f->world == 1 || f->comm), "call sq_fdm_init_slab first");
    if (enable && !f->sharded) {
        f->shard_lo = f->slab_lo; f->shard_hi = f->slab_hi;
        memcpy(f->tuned_shard, f->tuned, sizeof(f->tuned));
        memset(f->tuned, 0, sizeof(f->tuned));
        f->slab_lo = 0; f->slab_hi = (int)f->L;
        f->sharded = 1;
    } else if (!enable && f->sharded) {
        memcpy(f->tuned, f->tuned_shard, sizeof(f->tuned));
        f->slab_lo = f->shard_lo; f->slab_hi = f->shard_hi;
        f->sharded = 0;
    }
    f->manual_tuning = 0;
}
static void shard_swap(sq_fdm *f, bool enter) {
    int tmp[3][6];
    memcpy(tmp, f->tuned, sizeof(tmp));
    memcpy(f->tuned, f->tuned_shard, sizeof(tmp));
    memcpy(f->tuned_shard, tmp, sizeof(tmp));
    if (enter) { f->slab_lo = f->shard_lo; f->slab_hi = f->shard_hi; }
    else { f->slab_lo = 0; f->slab_hi = (int)f->L; }
}
// one solve of the sharded mode: slab solve, then every rank broadcasts its slab of the solution to the others
void fdm_cg_sharded(sq_fdm *f, double2 *x, const double2 *b, bool zero_start, sq_kpm *kpm, double tol, i64 maxiter, i64 *iters, double *eps) {
    shard_swap(f, true);
    try {
        fdm_cg_slab(f, x, b, zero_start, kpm, tol, maxiter, iters, eps);
    } catch (...) {
        shard_swap(f, false);
        throw;
    }
    shard_swap(f, false);
    if (f->world > 1) {
        SQ_NCCL(g_nccl.GroupStart());
        for (int q = 0; q < f->world; q++) {
            int lo, hi;
            slab_range((int)f->L, f->world, q, &lo, &hi);
            double2 *seg = x + (size_t)lo * f->N;
            SQ_NCCL(g_nccl.Broadcast(seg, seg, (size_t)(hi - lo) * f->N * 2, /*ncclDouble*/ 8, q, (ncclComm_t)f->comm, f->stream));
        }
        SQ_NCCL(g_nccl.GroupEnd());
    }
}

// column j of `cols` (V elements each) is valid on rank j % world: make all columns valid everywhere
void slab_broadcast_columns(sq_fdm *f, double2 *cols, size_t V, int ncols) {
    if (f->world <= 1) return;
    SQ_NCCL(g_nccl.GroupStart());
    for (int j = 0; j < ncols; j++) {
        double2 *c = cols + (size_t)j * V;
        SQ_NCCL(g_nccl.Broadcast(c, c, V * 2, ncclDouble, j % f->world, (ncclComm_t)f->comm, f->stream));
    }
    SQ_NCCL(g_nccl.GroupEnd());
}

// ---- mailboxes of the multi-GPU resident CG: one buffer per rank, mapped into every peer through CUDA IPC -------------------
size_t fdm_v3_mailbox_bytes(const sq_fdm *f);
void slab_mailbox_create(sq_fdm *f, char *out64) {
    SQ_REQUIRE(f->world >= 1 && f->world <= 8, "mailboxes support up to 8 ranks");
    SQ_CUDA(cudaSetDevice(f->device));
    const size_t bytes = fdm_v3_mailbox_bytes(f);
    if (!f->mail.p) f->mail.alloc(bytes);            // zero-filled: validity tag 0 / epoch 0
    f->mail_ptr[f->rank] = f->mail.p;
    cudaIpcMemHandle_t h;
    SQ_CUDA(cudaIpcGetMemHandle(&h, f->mail.p));
    static_assert(sizeof(h) == 64, "CUDA IPC handle size");
    memcpy(out64, &h, 64);
}
// handles64: world x 64 bytes, rank-major (every rank passes the same gathered array)
void slab_mailbox_open(sq_fdm *f, const char *handles64) {
    SQ_REQUIRE(f->mail.p != nullptr, "create the own mailbox first");
    SQ_CUDA(cudaSetDevice(f->device));
    for (int q = 0; q < f->world; q++) {
        if (q == f->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles64 + 64 * q, 64);
        void *p = nullptr;
        SQ_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        f->mail_ptr[q] = p;
    }
    f->mail_ready = 1;
}

void slab_destroy(sq_fdm *f) {
    for (int q = 0; q < 8; q++)
        if (f->mail_ready && q != f->rank && f->mail_ptr[q]) { cudaIpcCloseMemHandle(f->mail_ptr[q]); f->mail_ptr[q] = nullptr; }
    f->mail_ready = 0;
    if (f->comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)f->comm);
    f->comm = nullptr;
}

// exchange the boundary slices of v with the ring neighbours (in place, at their global index)
void fdm_halo_exchange(sq_fdm *f, double2 *v) {
    if (f->world <= 1) return;
    const int L = (int)f->L;
    const size_t nb = (size_t)f->N * sizeof(double2);
    const int prev = (f->rank + f->world - 1) % f->world, next = (f->rank + 1) % f->world;
    const int lo = f->slab_lo, hi = f->slab_hi;
    const size_t N = (size_t)f->N;
    ncclComm_t c = (ncclComm_t)f->comm;
    SQ_NCCL(g_nccl.GroupStart());
    SQ_NCCL(g_nccl.Send(v + (size_t)lo * N, nb, ncclChar, prev, c, f->stream));                    // their v[hi]
    SQ_NCCL(g_nccl.Send(v + (size_t)(hi - 1) * N, nb, ncclChar, next, c, f->stream));              // their v[lo-1]
    SQ_NCCL(g_nccl.Recv(v + (size_t)(hi % L) * N, nb, ncclChar, next, c, f->stream));              // next's first slice
    SQ_NCCL(g_nccl.Recv(v + (size_t)((lo - 1 + L) % L) * N, nb, ncclChar, prev, c, f->stream));    // prev's last slice
    SQ_NCCL(g_nccl.GroupEnd());
}

void fdm_allreduce_sum(sq_fdm *f, double *d_buf, int count) {
    if (f->world <= 1) return;
    SQ_NCCL(g_nccl.AllReduce(d_buf, d_buf, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)f->comm, f->stream));
}

// ---------------------------------------------------------------------------------------------------
// CG on a tau-slab: the reference recurrence (ConjugateGradient.jl:93-167) with local vector updates, one halo
// exchange and two scalar all-reduces per iteration.  scal[]: 0 |b|^2, 1 rr_old, 2 pAp, 3 rr_new, 4 done flag.
// ---------------------------------------------------------------------------------------------------
__global__ void k_pack_sum(const double *__restrict__ part, int n, double *__restrict__ dst) {
    double s = warp_sum_partials(part, n);
    if (threadIdx.x == 0) *dst = s;
}
__global__ void k_norm2_part(const double2 *__restrict__ a, size_t n, double *__restrict__ part) {
    __shared__ double red[32];
    double acc = 0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        double2 v = a[k];
        acc += v.x * v.x + v.y * v.y;
    }
    double v[1] = {acc};
    block_sum<1>(v, red);
    if (threadIdx.x == 0) part[blockIdx.x] = v[0];
}
__global__ void k_sub(double2 *__restrict__ r, const double2 *__restrict__ b, size_t n) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        r[k] = make_double2(b[k].x - r[k].x, b[k].y - r[k].y);
}
// x += alpha p ; r -= alpha z ; partial |r|^2 ; alpha = rr_old / pAp (both already all-reduced)
__global__ void k_slab_update_xr(const double *__restrict__ scal, double2 *__restrict__ x, double2 *__restrict__ r,
                                 const double2 *__restrict__ p, const double2 *__restrict__ z, size_t n, double *__restrict__ part) {
    __shared__ double red[32];
    if (scal[4] != 0.0) return;
    double alpha = scal[1] / scal[2];
    double acc = 0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        double2 pk = p[k], zk = z[k], xk = x[k], rk = r[k];
        xk = make_double2(fma(alpha, pk.x, xk.x), fma(alpha, pk.y, xk.y));
        rk = make_double2(fma(-alpha, zk.x, rk.x), fma(-alpha, zk.y, rk.y));
        x[k] = xk;
        r[k] = rk;
        acc += rk.x * rk.x + rk.y * rk.y;
    }
    double v[1] = {acc};
    block_sum<1>(v, red);
    if (threadIdx.x == 0) part[blockIdx.x] = v[0];
}
// eps test, beta = rr_new / rr_old, p = r + beta p ; block 0 rolls the scalars afterwards (separate tiny kernel)
__global__ void k_slab_update_p(const double *__restrict__ scal, double2 *__restrict__ p, const double2 *__restrict__ r, size_t n, double tol) {
    if (scal[4] != 0.0) return;
    double eps = sqrt(scal[3] / scal[0]);
    if (eps < tol || !(eps == eps)) return;
    double beta = scal[3] / scal[1];
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        double2 pk = p[k], rk = r[k];
        p[k] = make_double2(fma(beta, pk.x, rk.x), fma(beta, pk.y, rk.y));
    }
}
__global__ void k_slab_roll(double *scal, double tol, int iter) {
    if (scal[4] != 0.0) return;
    double eps = sqrt(scal[3] / scal[0]);
    scal[5] = eps;
    scal[6] = (double)iter;
    if (eps < tol) scal[4] = 1.0;
    else if (!(eps == eps)) scal[4] = 2.0;
    else scal[1] = scal[3];
}

bool fdm_v3_multi_possible(const sq_fdm *f);
bool fdm_v3_cg_resident1_multi(sq_fdm *f, double2 *x, double2 *r, CgState *state, i64 maxiter);
void fdm_v3_multi_reset_boundaries(sq_fdm *f);
void fdm_v3_prepare_native(sq_fdm *f);
void fdm_v3_to_native(sq_fdm *f, double2 *dst, const double2 *src);
void fdm_v3_from_native(sq_fdm *f, double2 *dst, const double2 *src);

// Resident multi-GPU solve (fdm_v3.cu, MULTI kernels): the whole solve is one cooperative launch per rank; sums and halos travel
// through the peer-mapped mailboxes.  Returns false if the configuration does not qualify (the NCCL loop below runs instead).
static bool cg_slab_resident(sq_fdm *f, double2 *x, const double2 *b, bool zero_start, double tol, i64 maxiter, i64 *iters, double *eps) {
    if (getenv("SQ_NO_RESIDENT_CG") || maxiter <= 0) return false;
    // (no fdm_select_tuning here: the autotuner's trial launches write the staging buffers the caller's b may still live in)
    if (!fdm_v3_multi_possible(f)) return false;
    const size_t N = (size_t)f->N, V = (size_t)f->L * N;
    const size_t off = (size_t)f->slab_lo * N, n = (size_t)(f->slab_hi - f->slab_lo) * N;
    const int TB = 256;
    const int G = (int)std::max<size_t>(1, std::min<size_t>((n + TB - 1) / TB, (size_t)f->num_sms * 4));
    cudaStream_t s = f->stream;
    double *part = f->part.p, *scal = f->scal.p;
    double2 *r = f->r.p;
    SQ_CUDA(cudaMemsetAsync(scal, 0, 16 * sizeof(double), s));
    k_norm2_part<<<G, TB, 0, s>>>(b + off, n, part);
    k_pack_sum<<<1, 32, 0, s>>>(part, G, scal + 0);
    if (zero_start) {
        SQ_CUDA(cudaMemcpyAsync(r + off, b + off, n * sizeof(double2), cudaMemcpyDeviceToDevice, s));
        SQ_CUDA(cudaMemsetAsync(x + off, 0, n * sizeof(double2), s));
    } else {
        fdm_halo_exchange(f, x);
        fdm_mul_dev(f, SQ_OP_MTM, r, x);
        k_sub<<<G, TB, 0, s>>>(r + off, b + off, n);
    }
    k_norm2_part<<<G, TB, 0, s>>>(r + off, n, part);
    k_pack_sum<<<1, 32, 0, s>>>(part, G, scal + 1);
    fdm_allreduce_sum(f, scal, 2);
    double h[2];
    SQ_CUDA(cudaMemcpyAsync(h, scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    SQ_CUDA(cudaStreamSynchronize(s));
    const double normb = std::sqrt(h[0]), eps0 = std::sqrt(h[1]) / normb;
    if (!(eps0 == eps0)) throw SqNumericalInstability("conjugate gradient (tau-slab): NaN encountered in the residual");
    if (eps0 < tol) { *iters = 0; *eps = eps0; return true; }
    fdm_v3_multi_reset_boundaries(f);                          // (before the exchange below: see there)
    fdm_halo_exchange(f, r);                                   // p0 = r0 of the neighbours' boundary slices
    fdm_v3_prepare_native(f);
    fdm_v3_to_native(f, f->v3_r.p, r);
    if (zero_start) SQ_CUDA(cudaMemsetAsync(f->v3_x.p, 0, V * sizeof(double2), s));
    else fdm_v3_to_native(f, f->v3_x.p, x);
    CgState st;
    memset(&st, 0, sizeof(st));
    st.normb = normb; st.tol = tol; st.eps = eps0; st.rz_re = h[1];
    *f->h_cg = st;
    SQ_CUDA(cudaMemcpyAsync(f->cg.p, f->h_cg, sizeof(CgState), cudaMemcpyHostToDevice, s));
    if (!fdm_v3_cg_resident1_multi(f, f->v3_x.p, f->v3_r.p, f->cg.p, maxiter)) return false;
    fdm_v3_from_native(f, f->tmp2.p, f->v3_x.p);
    SQ_CUDA(cudaMemcpyAsync(x + off, f->tmp2.p + off, n * sizeof(double2), cudaMemcpyDeviceToDevice, s));
    SQ_CUDA(cudaMemcpyAsync(f->h_cg, f->cg.p, sizeof(CgState), cudaMemcpyDeviceToHost, s));
    SQ_CUDA(cudaStreamSynchronize(s));
    f->v3_it_base += (unsigned long long)f->h_cg->iters + 4;      // identical on every rank: the tags of the next solve continue
    f->stats[SQ_STAT_CG_SOLVES]++; f->stats[SQ_STAT_CG_SLAB_RESIDENT]++;
    if (f->h_cg->done == 3) { f->stats[SQ_STAT_WATCHDOG]++; throw SqError("conjugate gradient (tau-slab): a grid-wide sum over the GPUs timed out (watchdog)"); }
    if (f->h_cg->done == 2) throw SqNumericalInstability("conjugate gradient (tau-slab): NaN encountered in the residual");
    *iters = f->h_cg->done ? f->h_cg->iters : maxiter;
    *eps = f->h_cg->eps;
    f->stats[SQ_STAT_CG_ITERS] += *iters;
    return true;
}


// ---------------------------------------------------------------------------------------------------
// KPM / tau-Fourier preconditioner in tau-slab mode: the all-to-all of SURVEY.md 8e.
//
// P^-1 = U^-1 f(B-bar) U mixes all time slices (U is the twisted DFT along tau), so a tau-slab vector has to change its
// partition on the way in and on the way out.  Rank g holds slices [lo_g, hi_g); it also OWNS the Matsubara frequencies with the
// same indices.  The DFT is linear in its input slices:
//     forward:  every rank transforms its own slices (all others zero-padded) into partial sums for ALL frequencies, then the
//               partial sums of the frequencies of rank q travel to rank q (grouped ncclSend / ncclRecv: the all-to-all) and are
//               added there in fixed rank order (deterministic);
//     middle:   rank g runs the Chebyshev recurrences of ITS frequencies (the order-1 frequencies are scalars folded into the
//               forward transform, so only frequencies with order > 1 cost anything);
//     inverse:  every rank transforms its frequencies (all others zero) into partial sums for ALL slices, second all-to-all, each
//               rank adds up the pieces of its own slices.
// Two exchanges of V 16 (P-1)/P bytes per GPU and apply, against src/KPMPreconditioner.jl:375-406 (FFT, transpose, per-frequency
// kpm_lmul!, transpose, inverse FFT) on one process.  B-bar (tau-means of the coefficient arrays) is computed locally: every
// rank keeps full-length coefficient arrays.
// ---------------------------------------------------------------------------------------------------
__global__ void k_slab_pad(double2 *__restrict__ dst, const double2 *__restrict__ src, size_t own_lo, size_t own_hi, size_t n) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x)
        dst[e] = (e >= own_lo && e < own_hi) ? src[e] : make_double2(0.0, 0.0);
}
// dst[e] = sum over ranks q = 0 .. world-1 (in this order) of q's partial sum: the own one from `own`, the others from recv[q]
__global__ void k_slab_sum_pieces(double2 *__restrict__ dst, const double2 *__restrict__ own, const double2 *__restrict__ recv, size_t piece,
                                  int world, int rank, size_t n) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        double2 acc = make_double2(0.0, 0.0);
        for (int q = 0; q < world; q++) {
            const double2 v = (q == rank) ? own[e] : recv[(size_t)q * piece + e];
            acc.x += v.x; acc.y += v.y;
        }
        dst[e] = acc;
    }
}
// rows [lo_q, hi_q) of src (partial sums for everybody) go to rank q; the pieces of the own rows are added up into dst's own rows
static void slab_exchange_sum(sq_fdm *f, sq_kpm *k, double2 *dst, const double2 *src) {
    const int W = f->world, me = f->rank, L = (int)f->L;
    const size_t N = (size_t)f->N;
    int lo, hi;
    slab_range(L, W, me, &lo, &hi);
    const size_t piece = (size_t)(L / W + 1) * N, n_own = (size_t)(hi - lo) * N;
    if (k->slab_recv.n < piece * W) k->slab_recv.alloc(piece * W, false);
    if (W > 1) {
        ncclComm_t c = (ncclComm_t)f->comm;
        SQ_NCCL(g_nccl.GroupStart());
        for (int q = 0; q < W; q++) {
            if (q == me) continue;
            int qlo, qhi;
            slab_range(L, W, q, &qlo, &qhi);
            SQ_NCCL(g_nccl.Send(src + (size_t)qlo * N, (size_t)(qhi - qlo) * N * 2, ncclDouble, q, c, f->stream));
            SQ_NCCL(g_nccl.Recv(k->slab_recv.p + (size_t)q * piece, n_own * 2, ncclDouble, q, c, f->stream));
        }
        SQ_NCCL(g_nccl.GroupEnd());
    }
    const int G = (int)std::max<size_t>(1, std::min<size_t>((n_own + 255) / 256, (size_t)f->num_sms * 4));
    k_slab_sum_pieces<<<G, 256, 0, f->stream>>>(dst + (size_t)lo * N, src + (size_t)lo * N, k->slab_recv.p, piece, W, me, n_own);
    SQ_LAUNCH_CHECK();
    f->launches++;
}

int tau_fft_launch(cudaStream_t stream, const std::vector<int> &radices, int L, int N, double2 *out, const double2 *in, bool inverse,
                   bool twist, const double2 *tw, const double2 *theta, const double *scale1, const double2 *dot_with,
                   double *dot_part, const CgState *skip, size_t smem_limit);

// out[own slices] = (P^-1 in)[own slices]; `in` is valid on the own slices.  The slab must be the rank's balanced share (slab_init).
void kpm_ldiv_slab(sq_kpm *k, double2 *out, const double2 *in) {
    sq_fdm *f = k->f;
    const int L = (int)f->L, W = f->world, me = f->rank;
    const size_t N = (size_t)f->N, V = (size_t)L * N;
    int lo, hi;
    slab_range(L, W, me, &lo, &hi);
    SQ_REQUIRE(lo == f->slab_lo && hi == f->slab_hi, "the preconditioner in tau-slab mode needs the balanced partition of sq_fdm_init_slab");
    if (!k->slab_a.p) { k->slab_a.alloc(V); k->slab_b.alloc(V); }
    // this rank's share of the Chebyshev schedule: the frequencies [lo, hi) with order > 1, longest first
    if (k->sched_slab_version != k->sched_version || k->sched_slab_lo != lo || k->sched_slab_hi != hi) {
        std::vector<int> mine;
        for (int n : k->h_sched) if (n >= lo && n < hi) mine.push_back(n);
        k->nsched_slab = (int)mine.size();
        k->d_sched_slab.alloc(mine.size() + 1, false);
        k->d_sched_slab.upload(mine.data(), mine.size(), f->stream);
        SQ_CUDA(cudaStreamSynchronize(f->stream));
        k->sched_slab_version = k->sched_version; k->sched_slab_lo = lo; k->sched_slab_hi = hi;
    }
    const int G = (int)std::min<size_t>((V + 255) / 256, (size_t)f->num_sms * 8);
    k_slab_pad<<<G, 256, 0, f->stream>>>(k->slab_a.p, in, (size_t)lo * N, (size_t)hi * N, V);
    double2 *zt = k->ztmp.p;
    tau_fft_launch(f->stream, k->radices, L, (int)N, zt, k->slab_a.p, false, true, k->tw.p, k->theta.p, k->d_scale1.p, nullptr, nullptr, nullptr,
                   f->smem_optin);
    SQ_CUDA(cudaMemsetAsync(k->slab_b.p, 0, V * sizeof(double2), f->stream));
    slab_exchange_sum(f, k, k->slab_b.p, zt);                                   // all-to-all #1: tau-slab -> frequency-slab
    kpm_cheb_apply(k, k->slab_b.p, k->d_sched_slab.p, k->nsched_slab, 1, 0, nullptr);
    tau_fft_launch(f->stream, k->radices, L, (int)N, zt, k->slab_b.p, true, true, k->tw.p, k->theta.p, nullptr, nullptr, nullptr, nullptr,
                   f->smem_optin);
    slab_exchange_sum(f, k, out, zt);                                           // all-to-all #2: frequency-slab -> tau-slab
    f->launches += 3;
}

// partials of conj(a).b (re, im) on a slab
__global__ void k_slab_dotc_part(const double2 *__restrict__ a, const double2 *__restrict__ b, size_t n, double *__restrict__ part) {
    __shared__ double red[2 * 32];
    double v[2] = {0.0, 0.0};
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const double2 x = a[e], y = b[e];
        v[0] += x.x * y.x + x.y * y.y;
        v[1] += x.x * y.y - x.y * y.x;
    }
    block_sum<2>(v, red);
    if (threadIdx.x == 0) { part[blockIdx.x] = v[0]; part[SQ_MAXPART + blockIdx.x] = v[1]; }
}
// scal: 0 |b|^2, 1-2 r.z (old), 3 p.Ap, 4 done, 5 eps, 6 iterations, 7 |r|^2, 8-9 r.z (new)
__global__ void k_slab_prec_update_xr(const double *__restrict__ scal, double2 *__restrict__ x, double2 *__restrict__ r,
                                      const double2 *__restrict__ p, const double2 *__restrict__ q, size_t n, double *__restrict__ part) {
    __shared__ double red[32];
    if (scal[4] != 0.0) return;
    const double2 alpha = make_double2(scal[1] / scal[3], scal[2] / scal[3]);       // (r.z) / (p.Ap), p.Ap = |M p|^2 real
    double acc = 0;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const double2 pk = p[e], qk = q[e];
        const double2 xk = cadd(x[e], cmul(alpha, pk)), rk = csub(r[e], cmul(alpha, qk));
        x[e] = xk;
        r[e] = rk;
        acc += rk.x * rk.x + rk.y * rk.y;
    }
    double v[1] = {acc};
    block_sum<1>(v, red);
    if (threadIdx.x == 0) part[blockIdx.x] = v[0];
}
__global__ void k_slab_prec_check(double *scal, double tol, int iter) {
    if (scal[4] != 0.0) return;
    const double eps = sqrt(scal[7] / scal[0]);
    scal[5] = eps;
    scal[6] = (double)iter;
    if (eps < tol) scal[4] = 1.0;
    else if (!(eps == eps)) scal[4] = 2.0;
}
__global__ void k_slab_prec_update_p(const double *__restrict__ scal, double2 *__restrict__ p, const double2 *__restrict__ z, size_t n) {
    if (scal[4] != 0.0) return;
    const double2 beta = cdiv(make_double2(scal[8], scal[9]), make_double2(scal[1], scal[2]));
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x)
        p[e] = cadd(z[e], cmul(beta, p[e]));
}
__global__ void k_slab_prec_roll(double *scal) {
    if (scal[4] != 0.0) return;
    scal[1] = scal[8];
    scal[2] = scal[9];
}
__global__ void k_pack_sum2(const double *__restrict__ part, int n, double *__restrict__ dst) {
    const double a = warp_sum_partials(part, n), b = warp_sum_partials(part + SQ_MAXPART, n);
    if (threadIdx.x == 0) { dst[0] = a; dst[1] = b; }
}

// Preconditioned CG on a tau-slab: the reference recurrence (ConjugateGradient.jl:169-249) with the distributed P^-1 above, one halo
// exchange and three small all-reduces (p.Ap, |r|^2, r.z) per iteration.
static void cg_slab_prec(sq_fdm *f, double2 *x, const double2 *b, bool zero_start, sq_kpm *kpm, double tol, i64 maxiter, i64 *iters, double *eps) {
    SQ_REQUIRE(f->world > 1 || (f->slab_lo == 0 && f->slab_hi == (int)f->L), "a preconditioned solve on a partial slice range needs the other ranks");
    const size_t N = (size_t)f->N;
    const size_t off = (size_t)f->slab_lo * N, n = (size_t)(f->slab_hi - f->slab_lo) * N;
    const int TB = 256;
    const int G = (int)std::max<size_t>(1, std::min<size_t>((n + TB - 1) / TB, (size_t)f->num_sms * 4));
    cudaStream_t s = f->stream;
    double *part = f->part.p, *scal = f->scal.p;
    double2 *r = f->r.p, *p = f->p.p, *z = f->z.p, *q = nullptr;
    if (f->prec_q.n < (size_t)f->L * N) f->prec_q.alloc((size_t)f->L * N, false);      // (tmp1 / tmp2 are scratch of the matvec itself)
    q = f->prec_q.p;
    SQ_CUDA(cudaMemsetAsync(scal, 0, 16 * sizeof(double), s));
    k_norm2_part<<<G, TB, 0, s>>>(b + off, n, part);
    k_pack_sum<<<1, 32, 0, s>>>(part, G, scal + 0);
    if (zero_start) {
        SQ_CUDA(cudaMemcpyAsync(r + off, b + off, n * sizeof(double2), cudaMemcpyDeviceToDevice, s));
        SQ_CUDA(cudaMemsetAsync(x + off, 0, n * sizeof(double2), s));
    } else {
        fdm_halo_exchange(f, x);
        fdm_mul_dev(f, SQ_OP_MTM, r, x);
        k_sub<<<G, TB, 0, s>>>(r + off, b + off, n);
    }
    k_norm2_part<<<G, TB, 0, s>>>(r + off, n, part);
    k_pack_sum<<<1, 32, 0, s>>>(part, G, scal + 7);
    fdm_allreduce_sum(f, scal, 1);
    fdm_allreduce_sum(f, scal + 7, 1);
    k_slab_prec_check<<<1, 1, 0, s>>>(scal, tol, 0);                              // eps0 test (:206-214)
    kpm_ldiv_slab(kpm, z, r);
    SQ_CUDA(cudaMemcpyAsync(p + off, z + off, n * sizeof(double2), cudaMemcpyDeviceToDevice, s));
    k_slab_dotc_part<<<G, TB, 0, s>>>(r + off, z + off, n, part);
    k_pack_sum2<<<1, 32, 0, s>>>(part, G, scal + 1);
    fdm_allreduce_sum(f, scal + 1, 2);
    f->launches += 8;
    double h[10];
    SQ_CUDA(cudaMemcpyAsync(h, scal, 10 * sizeof(double), cudaMemcpyDeviceToHost, s));
    SQ_CUDA(cudaStreamSynchronize(s));
    i64 it = 0;
    const int batch = 4;
    bool finished = (h[4] != 0.0) || maxiter <= 0;
    while (!finished) {
        const i64 upto = std::min<i64>(maxiter, it + batch);
        for (; it < upto;) {
            it++;
            fdm_halo_exchange(f, p);
            int npart = 0;
            fdm_mul_dev(f, SQ_OP_MTM, q, p, part, &npart, nullptr);               // q = M^T M p on the slab, |M p|^2 partials
            k_pack_sum<<<1, 32, 0, s>>>(part, npart, scal + 3);
            fdm_allreduce_sum(f, scal + 3, 1);
            k_slab_prec_update_xr<<<G, TB, 0, s>>>(scal, x + off, r + off, p + off, q + off, n, part);
            k_pack_sum<<<1, 32, 0, s>>>(part, G, scal + 7);
            fdm_allreduce_sum(f, scal + 7, 1);
            k_slab_prec_check<<<1, 1, 0, s>>>(scal, tol, (int)it);
            kpm_ldiv_slab(kpm, z, r);
            k_slab_dotc_part<<<G, TB, 0, s>>>(r + off, z + off, n, part);
            k_pack_sum2<<<1, 32, 0, s>>>(part, G, scal + 8);
            fdm_allreduce_sum(f, scal + 8, 2);
            k_slab_prec_update_p<<<G, TB, 0, s>>>(scal, p + off, z + off, n);
            k_slab_prec_roll<<<1, 1, 0, s>>>(scal);
            f->launches += 8;
        }
        SQ_LAUNCH_CHECK();
        SQ_CUDA(cudaMemcpyAsync(h, scal, 10 * sizeof(double), cudaMemcpyDeviceToHost, s));
        SQ_CUDA(cudaStreamSynchronize(s));
        if (h[4] != 0.0 || it >= maxiter) finished = true;
    }
    f->stats[SQ_STAT_CG_SOLVES]++; f->stats[SQ_STAT_CG_SLAB_PREC]++;
    if (h[4] == 2.0) throw SqNumericalInstability("conjugate gradient (tau-slab, preconditioned): NaN encountered in the residual");
    *iters = h[4] != 0.0 ? (i64)h[6] : maxiter;
    *eps = h[5];
    f->stats[SQ_STAT_CG_ITERS] += *iters;
}

void fdm_cg_slab(sq_fdm *f, double2 *x, const double2 *b, bool zero_start, sq_kpm *kpm, double tol, i64 maxiter, i64 *iters, double *eps) {
    if (kpm && kpm->active) { cg_slab_prec(f, x, b, zero_start, kpm, tol, maxiter, iters, eps); return; }
    if (cg_slab_resident(f, x, b, zero_start, tol, maxiter, iters, eps)) return;
    const size_t N = (size_t)f->N;
    const size_t off = (size_t)f->slab_lo * N, n = (size_t)(f->slab_hi - f->slab_lo) * N;
    const int TB = 256;
    const int G = (int)std::max<size_t>(1, std::min<size_t>((n + TB - 1) / TB, (size_t)f->num_sms * 4));
    cudaStream_t s = f->stream;
    double *part = f->part.p, *scal = f->scal.p;
    double2 *r = f->r.p, *p = f->p.p, *z = f->z.p;
    SQ_CUDA(cudaMemsetAsync(scal, 0, 16 * sizeof(double), s));
    k_norm2_part<<<G, TB, 0, s>>>(b + off, n, part);
    k_pack_sum<<<1, 32, 0, s>>>(part, G, scal + 0);
    if (zero_start) {
        SQ_CUDA(cudaMemcpyAsync(r + off, b + off, n * sizeof(double2), cudaMemcpyDeviceToDevice, s));
        SQ_CUDA(cudaMemsetAsync(x + off, 0, n * sizeof(double2), s));
    } else {
        fdm_halo_exchange(f, x);
        fdm_mul_dev(f, SQ_OP_MTM, r, x);
        k_sub<<<G, TB, 0, s>>>(r + off, b + off, n);
    }
    SQ_CUDA(cudaMemcpyAsync(p + off, r + off, n * sizeof(double2), cudaMemcpyDeviceToDevice, s));
    k_norm2_part<<<G, TB, 0, s>>>(r + off, n, part);
    k_pack_sum<<<1, 32, 0, s>>>(part, G, scal + 1);
    fdm_allreduce_sum(f, scal, 2);
    SQ_CUDA(cudaMemcpyAsync(scal + 3, scal + 1, sizeof(double), cudaMemcpyDeviceToDevice, s));
    k_slab_roll<<<1, 1, 0, s>>>(scal, tol, 0);                                   // initial eps test (iter 0)
    f->launches += 5;
    i64 it = 0;
    const int batch = 16;
    double h[8];
    bool finished = false;
    while (!finished) {
        i64 upto = std::min<i64>(maxiter, it + batch);
        for (; it < upto;) {
            it++;
            fdm_halo_exchange(f, p);
            int npart = 0;
            fdm_mul_dev(f, SQ_OP_MTM, z, p, part, &npart, nullptr);               // z = M^T M p on the slab, |Mp|^2 partials
            k_pack_sum<<<1, 32, 0, s>>>(part, npart, scal + 2);
            fdm_allreduce_sum(f, scal + 2, 1);
            k_slab_update_xr<<<G, TB, 0, s>>>(scal, x + off, r + off, p + off, z + off, n, part);
            k_pack_sum<<<1, 32, 0, s>>>(part, G, scal + 3);
            fdm_allreduce_sum(f, scal + 3, 1);
            k_slab_update_p<<<G, TB, 0, s>>>(scal, p + off, r + off, n, tol);
            k_slab_roll<<<1, 1, 0, s>>>(scal, tol, (int)it);
            f->launches += 5;
        }
        SQ_LAUNCH_CHECK();
        SQ_CUDA(cudaMemcpyAsync(h, scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, s));
        SQ_CUDA(cudaStreamSynchronize(s));
        if (h[4] != 0.0 || it >= maxiter) finished = true;
    }
    if (h[4] == 2.0) throw SqNumericalInstability("conjugate gradient (tau-slab): NaN encountered in the residual");
    *iters = h[4] != 0.0 ? (i64)h[6] : maxiter;
    *eps = h[5];
    f->stats[SQ_STAT_CG_SOLVES]++; f->stats[SQ_STAT_CG_SLAB_NCCL]++;
    f->stats[SQ_STAT_CG_ITERS] += *iters;
}
