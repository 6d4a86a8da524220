// fdm_v2.cu -- K1 fast path: the fused M / M^T / M^T M kernels for the symmetric propagator.
//
// Same algorithm and slab structure as k_fdm_fused (fdm.cu) -- one CTA stages a tau-slab in shared memory and
// produces S output slices from one pass over HBM -- re-engineered around what ncu showed on the first version
// (profiles/r1_*): the kernel was bound by exposed L2 latency of per-step index / coefficient loads, ~100
// instructions per bond item and 26 % shared-memory bank conflicts, not by HBM.
//   * every thread owns ONE bond per colour for the whole kernel: its two shared-memory slots per sweep step
//     live in registers (no index loads in the inner loop);
//   * (cosh, sinh) of the next sweep step and exp(-dtau V) of the middle step are prefetched into registers
//     before the barrier of the current step (software pipeline over the 2C-1 steps);
//   * sites are renumbered inside shared memory ("slots") and bonds re-ordered inside each colour at create
//     time so that consecutive lanes touch consecutive 16-byte slots (conflict-free LDS.128 / STS.128);
//   * 32-bit index arithmetic throughout.
// The arithmetic (fma pattern, order of colour factors) is identical to fdm.cu, so both paths give bit-identical
// results; tests/test_gpu_fdm.py checks one against the other and both against the oracle.
#include "sq_internal.h"

#include <cstring>
#include <type_traits>

struct K2Params {
    int L, N, Nh, C, S, TX;
    int lb, le;             // slices [lb, le) produced by this launch
    int clo[SQ_MAXC], chi[SQ_MAXC];
    int nunc0;
    const int2 *nts;        // bond -> (slot_i, slot_j)
    const int2 *nt;         // bond -> (site_i, site_j)   (natural indices, for exp(-dtau V))
    const int *slot;        // site -> slot
    const int *unc0;        // sites not touched by colour 0
    const double2 *cs;      // [l][h]
    const double *expV;     // [l][i]
    long long *dbg;         // optional per-phase clock stamps of CTA 0 (profiling aid, NULL in production)
    // CG fusion (MODE 2): the search direction is formed on load, p_new = d + beta * in  (d = r, or z = P^-1 r),
    // beta = (sum of beta_part) / (cur->rz); the convergence test |r|/|b| < tol of the PREVIOUS iteration is
    // evaluated in the prologue from rr_part and recorded in *nxt by CTA 0.  cg_d == NULL: plain matvec.
    const double2 *cg_d;
    double2 *cg_pnew;
    const CgState *cg_cur;
    CgState *cg_nxt;
    const double *cg_rr_part, *cg_beta_part;
    int cg_nrr, cg_nbeta, cg_beta_complex, cg_iter, cg_check;
};

__device__ __forceinline__ void rot2(double2 &a, double2 &b, double c, double s) {
    double2 na = make_double2(fma(s, b.x, c * a.x), fma(s, b.y, c * a.y));
    double2 nb = make_double2(fma(s, a.x, c * b.x), fma(s, a.y, c * b.y));
    a = na;
    b = nb;
}

// During a sweep the slices are independent: threads with the same ty (one slice set) only have to synchronise with
// each other.  Named barriers per ty-group are cheaper than a CTA-wide barrier (fewer warps to drain) and let the
// groups slip against each other, overlapping one group's shared-memory latency with the other's arithmetic.
__device__ __forceinline__ void group_sync(int ty, int TX, int TY) {
    // measured on B200 (cfg4, S=3, 1024 threads): per-group named barriers were 15 % SLOWER than the CTA-wide barrier
    // (phase 1: 10.6k vs 9.1k cycles), so the CTA-wide barrier stays; kept as a switch for experiments.
#ifdef SQ_GROUP_BARRIERS
    if (TY > 1 && TY <= 15) { asm volatile("bar.sync %0, %1;" ::"r"(ty + 1), "r"(TX) : "memory"); return; }
#endif
    __syncthreads();
}

// sweep-step -> colour for B = Gamma D Gamma^T: C-1, ..., 1, 0 (fused with D), 1, ..., C-1
__device__ __forceinline__ int step_color(int st, int C) { return st < C - 1 ? C - 1 - st : st - (C - 1); }

// pick element c of a small register array without dynamic register indexing (selects, no local memory)
template <int CMAX>
__device__ __forceinline__ int pick_reg(const int (&a)[CMAX], int c) {
    int r = a[0];
#pragma unroll
    for (int q = 1; q < CMAX; q++) r = (c == q) ? a[q] : r;
    return r;
}

template <int CMAX, int KMAX>
struct Engine {
    int c_i[CMAX], c_j[CMAX];         // shared-memory slots of this thread's bond in every colour (-1: idle in that colour)
    int ni0, nj0;                     // natural site indices of the colour-0 bond
    int tx, ty, TY;

    __device__ __forceinline__ void init(const K2Params &P) {
        tx = threadIdx.x % P.TX;
        ty = threadIdx.x / P.TX;
        TY = blockDim.x / P.TX;
#pragma unroll
        for (int c = 0; c < CMAX; c++) {
            c_i[c] = -1;
            c_j[c] = 0;
            if (c < P.C) {
                int lo = P.clo[c];
                if (tx < P.chi[c] - lo) {
                    int2 ij = __ldg(P.nts + lo + tx);
                    c_i[c] = ij.x;
                    c_j[c] = ij.y;
                }
            }
        }
        ni0 = -1;
        nj0 = 0;
        if (tx < P.chi[0] - P.clo[0]) {
            int2 q = __ldg(P.nt + P.clo[0] + tx);
            ni0 = q.x;
            nj0 = q.y;
        }
    }

    // buf[k] <- B_{lfirst + k} buf[k] for k < nsl (slot-ordered slices in shared memory).  Rolled loop over the
    // 2C-1 sweep steps; the coefficients of step st+1 are in flight while step st computes.
    __device__ __forceinline__ void apply_B(double2 *buf, int nsl, int lfirst, const K2Params &P) {
        const int N = P.N, Nh = P.Nh, L = P.L, C = P.C;
        const int nsteps = 2 * C - 1;
        int lj[KMAX];
#pragma unroll
        for (int j = 0; j < KMAX; j++) {
            int k = ty + j * TY;
            int l = lfirst + k;
            l = l >= L ? l - L : l;
            lj[j] = (k < nsl) ? l : -1;
        }
        double2 nxt[KMAX];
        int c = step_color(0, C);
        int oi = pick_reg<CMAX>(c_i, c), oj = pick_reg<CMAX>(c_j, c);
#pragma unroll
        for (int j = 0; j < KMAX; j++) {
            nxt[j] = make_double2(1.0, 0.0);
            if (lj[j] >= 0 && oi >= 0) nxt[j] = __ldg(P.cs + lj[j] * Nh + P.clo[c] + tx);
        }
        for (int st = 0; st < nsteps; st++) {
            double2 cur[KMAX];
#pragma unroll
            for (int j = 0; j < KMAX; j++) cur[j] = nxt[j];
            const int oi_c = oi, oj_c = oj;
            const bool mid = (st == C - 1);
            double di[KMAX], dj[KMAX];
            if (mid) {
#pragma unroll
                for (int j = 0; j < KMAX; j++) {
                    di[j] = dj[j] = 1.0;
                    if (lj[j] >= 0 && ni0 >= 0) { di[j] = __ldg(P.expV + lj[j] * N + ni0); dj[j] = __ldg(P.expV + lj[j] * N + nj0); }
                }
            }
            if (st + 1 < nsteps) {                                  // prefetch the next step's (cosh, sinh)
                c = step_color(st + 1, C);
                oi = pick_reg<CMAX>(c_i, c);
                oj = pick_reg<CMAX>(c_j, c);
#pragma unroll
                for (int j = 0; j < KMAX; j++)
                    if (lj[j] >= 0 && oi >= 0) nxt[j] = __ldg(P.cs + lj[j] * Nh + P.clo[c] + tx);
            }
            if (oi_c >= 0) {
#pragma unroll
                for (int j = 0; j < KMAX; j++) {
                    if (lj[j] >= 0) {
                        double2 *u = buf + (ty + j * TY) * N;
                        double2 a = u[oi_c], b = u[oj_c];
                        rot2(a, b, cur[j].x, cur[j].y);
                        if (mid) {
                            a = make_double2(di[j] * a.x, di[j] * a.y);
                            b = make_double2(dj[j] * b.x, dj[j] * b.y);
                            rot2(a, b, cur[j].x, cur[j].y);
                        }
                        u[oi_c] = a;
                        u[oj_c] = b;
                    }
                }
            }
            if (mid && P.nunc0 > 0) {                               // sites colour 0 does not touch still get D
                for (int k = ty; k < nsl; k += TY) {
                    int l = lfirst + k;
                    l = l >= L ? l - L : l;
                    for (int q = tx; q < P.nunc0; q += P.TX) {
                        int i = __ldg(P.unc0 + q);
                        double d = __ldg(P.expV + l * N + i);
                        int s = __ldg(P.slot + i);
                        double2 a = buf[k * N + s];
                        buf[k * N + s] = make_double2(d * a.x, d * a.y);
                    }
                }
            }
            __syncthreads();
        }
    }
};

// tau-independent hoppings (every Holstein-type model: t does not depend on the phonon field): the (cosh, sinh) of
// the thread's bond in each colour are kernel-lifetime constants held in registers, so the inner loop has no global
// loads at all except exp(-dtau V) of the middle step, which is fetched at the start of the sweep.
template <int CMAX, int KMAX>
struct EngineU {
    int c_i[CMAX], c_j[CMAX];
    double cc[CMAX], ss[CMAX];
    int ni0, nj0;
    int tx, ty, TY;

    __device__ __forceinline__ void init(const K2Params &P) {
        tx = threadIdx.x % P.TX;
        ty = threadIdx.x / P.TX;
        TY = blockDim.x / P.TX;
#pragma unroll
        for (int c = 0; c < CMAX; c++) {
            c_i[c] = -1;
            c_j[c] = 0;
            cc[c] = 1.0;
            ss[c] = 0.0;
            if (c < P.C) {
                int lo = P.clo[c];
                if (tx < P.chi[c] - lo) {
                    int2 ij = __ldg(P.nts + lo + tx);
                    double2 v = __ldg(P.cs + lo + tx);            // slice 0 == every slice
                    c_i[c] = ij.x;
                    c_j[c] = ij.y;
                    cc[c] = v.x;
                    ss[c] = v.y;
                }
            }
        }
        ni0 = -1;
        nj0 = 0;
        if (tx < P.chi[0] - P.clo[0]) {
            int2 q = __ldg(P.nt + P.clo[0] + tx);
            ni0 = q.x;
            nj0 = q.y;
        }
    }

    template <int Q>
    __device__ __forceinline__ void step(double2 *buf, const int (&kj)[KMAX]) {
        if (c_i[Q] >= 0) {
#pragma unroll
            for (int j = 0; j < KMAX; j++) {
                if (kj[j] >= 0) {
                    double2 *u = buf + kj[j];
                    double2 a = u[c_i[Q]], b = u[c_j[Q]];
                    rot2(a, b, cc[Q], ss[Q]);
                    u[c_i[Q]] = a;
                    u[c_j[Q]] = b;
                }
            }
        }
        group_sync(ty, blockDim.x / TY, TY);
    }

    template <int Q>
    __device__ __forceinline__ void down(double2 *buf, const int (&kj)[KMAX], int C) {   // colours CMAX-1 ... 1
        if constexpr (Q >= 1) {
            if (Q < C) step<Q>(buf, kj);
            down<Q - 1>(buf, kj, C);
        }
    }
    template <int Q>
    __device__ __forceinline__ void up(double2 *buf, const int (&kj)[KMAX], int C) {     // colours 1 ... CMAX-1
        if constexpr (Q < CMAX) {
            if (Q < C) step<Q>(buf, kj);
            up<Q + 1>(buf, kj, C);
        }
    }

    __device__ __forceinline__ void apply_B(double2 *buf, int nsl, int lfirst, const K2Params &P) {
        const int N = P.N, L = P.L, C = P.C;
        int kj[KMAX];                       // element offset of the thread's slices inside buf (-1: none)
        double di[KMAX], dj[KMAX];
#pragma unroll
        for (int j = 0; j < KMAX; j++) {
            int k = ty + j * TY;
            int l = lfirst + k;
            l = l >= L ? l - L : l;
            kj[j] = (k < nsl) ? k * N : -1;
            di[j] = dj[j] = 1.0;
            if (k < nsl && ni0 >= 0) { di[j] = __ldg(P.expV + l * N + ni0); dj[j] = __ldg(P.expV + l * N + nj0); }
        }
        down<CMAX - 1>(buf, kj, C);
        if (c_i[0] >= 0) {
#pragma unroll
            for (int j = 0; j < KMAX; j++) {
                if (kj[j] >= 0) {
                    double2 *u = buf + kj[j];
                    double2 a = u[c_i[0]], b = u[c_j[0]];
                    rot2(a, b, cc[0], ss[0]);
                    a = make_double2(di[j] * a.x, di[j] * a.y);
                    b = make_double2(dj[j] * b.x, dj[j] * b.y);
                    rot2(a, b, cc[0], ss[0]);
                    u[c_i[0]] = a;
                    u[c_j[0]] = b;
                }
            }
        }
        if (P.nunc0 > 0) {
            for (int k = ty; k < nsl; k += TY) {
                int l = lfirst + k;
                l = l >= L ? l - L : l;
                for (int q = tx; q < P.nunc0; q += P.TX) {
                    int i = __ldg(P.unc0 + q);
                    double d = __ldg(P.expV + l * N + i);
                    int sl = __ldg(P.slot + i);
                    double2 a = buf[k * N + sl];
                    buf[k * N + sl] = make_double2(d * a.x, d * a.y);
                }
            }
        }
        group_sync(ty, blockDim.x / TY, TY);
        up<1>(buf, kj, C);
    }
};

template <int MODE, int CMAX, int KMAX, int UNI>
__global__ void __launch_bounds__(1024, 1)
k_fdm_fused_v2(const __grid_constant__ K2Params P, double2 *__restrict__ out, const double2 *__restrict__ in,
               double *__restrict__ pAp_part, const CgState *__restrict__ skip) {
    extern __shared__ double2 smem[];
    __shared__ double red[32];
    sq_pdl_prologue();
    if (skip && skip->done) {
        // converged earlier in this batch: keep the ping-ponged solver state consistent, do nothing else
        if (MODE == 2 && P.cg_d != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *P.cg_nxt = *skip;
        return;
    }
#define SQ_STAMP(k) do { if (P.dbg && blockIdx.x == 0 && threadIdx.x == 0) P.dbg[k] = clock64(); } while (0)
    SQ_STAMP(0);
    typename std::conditional<UNI != 0, EngineU<CMAX, KMAX>, Engine<CMAX, KMAX>>::type E;
    E.init(P);
    SQ_STAMP(1);
    const int L = P.L, N = P.N, T = blockDim.x;
    const int l0 = P.lb + blockIdx.x * P.S;
    const int ns = min(P.S, P.le - l0);
    // A: input slices; W: work slices.  MODE 2: A[k] = v[l0-1+k] (k <= ns+1), W[k] = A[k] (k <= ns).
    // MODE 0: A[k] = v[l0-1+k] (k <= ns), W[k] = A[k] (k < ns).  MODE 1: A[k] = v[l0+k] (k <= ns), W[k] = A[k+1] (k < ns).
    const int nA = (MODE == 2) ? ns + 2 : ns + 1;
    double2 *A = smem, *W = smem + ((MODE == 2) ? P.S + 2 : P.S + 1) * N;
    const int lA0 = (MODE == 1) ? l0 : l0 - 1;
    const bool fuse = (MODE == 2) && (P.cg_d != nullptr);
    double2 beta = make_double2(0.0, 0.0);
    if (fuse) {
        __shared__ double sh[4];
        const CgState st = *P.cg_cur;
        if (threadIdx.x < 32) {
            double rr = warp_sum_partials(P.cg_rr_part, P.cg_nrr);
            double br = rr, bi = 0.0;
            if (P.cg_beta_part != P.cg_rr_part) {
                br = warp_sum_partials(P.cg_beta_part, P.cg_nbeta);
                bi = P.cg_beta_complex ? warp_sum_partials(P.cg_beta_part + SQ_MAXPART, P.cg_nbeta) : 0.0;
            }
            if (threadIdx.x == 0) {
                double eps = sqrt(rr) / st.normb;
                int stop = 0;
                if (P.cg_check) stop = (eps < st.tol) ? 1 : ((eps == eps) ? 0 : 2);
                double2 b = cdiv(make_double2(br, bi), make_double2(st.rz_re, st.rz_im));
                sh[0] = b.x; sh[1] = b.y; sh[2] = (double)stop;
                if (blockIdx.x == 0) {
                    CgState nx = st;
                    if (P.cg_check) { nx.eps = eps; nx.iters = P.cg_iter - 1; nx.done = stop; }
                    if (!stop) { nx.rz_re = br; nx.rz_im = bi; }
                    *P.cg_nxt = nx;
                }
            }
        }
        __syncthreads();
        if (sh[2] != 0.0) return;
        beta = make_double2(sh[0], sh[1]);
    }
    for (int i = threadIdx.x; i < N; i += T) {
        const int s = __ldg(P.slot + i);
        for (int k = 0; k < nA; k++) {
            int l = lA0 + k;
            l = l < 0 ? l + L : (l >= L ? l - L : l);
            double2 v = in[l * N + i];
            if (fuse) {
                v = cadd(P.cg_d[l * N + i], cmul(beta, v));
                if (k >= 1 && k <= ns) P.cg_pnew[l * N + i] = v;
            }
            A[k * N + s] = v;
            if (MODE == 2) { if (k <= ns) W[k * N + s] = v; }
            else if (MODE == 0) { if (k < ns) W[k * N + s] = v; }
            else { if (k >= 1) W[(k - 1) * N + s] = v; }
        }
    }
    __syncthreads();
    SQ_STAMP(2);
    double acc = 0.0;
    const int nphase = (MODE == 2) ? 2 : 1;
    for (int phase = 0; phase < nphase; phase++) {
        if (phase == 0) {
            int lf = (MODE == 1) ? l0 + 1 : l0;
            E.apply_B(W, (MODE == 2) ? ns + 1 : ns, lf >= L ? lf - L : lf, P);
        } else {
            int lf = l0 + 1;
            E.apply_B(A + 2 * N, ns, lf >= L ? lf - L : lf, P);
        }
        if (UNI) __syncthreads();          // the uniform engine synchronises per slice group inside the sweep
        SQ_STAMP(3 + 2 * phase);
        if (MODE == 2 && phase == 0) {
            // w[l0+k] = v[l0+k] -/+ W[k] (+ on the antiperiodic slice); T[k] := A[k+1] for k >= 1.  Slot space, and
            // the same thread walks all k so A[k+1] is read before it is overwritten.
            for (int i = threadIdx.x; i < N; i += T) {
                for (int k = 0; k <= ns; k++) {
                    int l = l0 + k;
                    l = l >= L ? l - L : l;
                    double sg = (l == 0) ? 1.0 : -1.0;
                    double2 a = A[(k + 1) * N + i], b = W[k * N + i];
                    double2 w = make_double2(fma(sg, b.x, a.x), fma(sg, b.y, a.y));
                    W[k * N + i] = w;
                    if (k >= 1) A[(k + 1) * N + i] = w;
                    if (k < ns) acc += w.x * w.x + w.y * w.y;
                }
            }
            __syncthreads();
            SQ_STAMP(4);
        }
    }
    for (int i = threadIdx.x; i < N; i += T) {
        const int s = __ldg(P.slot + i);
        for (int k = 0; k < ns; k++) {
            double2 a, b;
            int lb;
            if (MODE == 2) { a = W[k * N + s]; b = A[(k + 2) * N + s]; lb = l0 + k + 1; }      // out[l0+k] = w[l0+k] -/+ B^T_{l0+k+1} w[l0+k+1]
            else if (MODE == 0) { a = A[(k + 1) * N + s]; b = W[k * N + s]; lb = l0 + k; }     // out[l0+k] = v[l0+k] -/+ B_{l0+k} v[l0+k-1]
            else { a = A[k * N + s]; b = W[k * N + s]; lb = l0 + k + 1; }                       // out[l0+k] = v[l0+k] -/+ B^T_{l0+k+1} v[l0+k+1]
            lb = lb >= L ? lb - L : lb;
            double sg = (lb == 0) ? 1.0 : -1.0;
            out[(l0 + k) * N + i] = make_double2(fma(sg, b.x, a.x), fma(sg, b.y, a.y));
        }
    }
    SQ_STAMP(6);
    if (MODE == 2 && pAp_part) {
        double v[1] = {acc};
        block_sum<1>(v, red);
        if (threadIdx.x == 0) pAp_part[blockIdx.x] = v[0];
    }
    SQ_STAMP(7);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef void (*v2_kernel_t)(const K2Params, double2 *, const double2 *, double *, const CgState *);

template <int MODE, int CMAX, int UNI>
static v2_kernel_t pick_k(int KMAX) {
    switch (KMAX) {
        case 1: return k_fdm_fused_v2<MODE, CMAX, 1, UNI>;
        case 2: return k_fdm_fused_v2<MODE, CMAX, 2, UNI>;
        default: return k_fdm_fused_v2<MODE, CMAX, 4, UNI>;
    }
}
template <int UNI>
static v2_kernel_t pick_u(int mode, int ncol, int KMAX) {
    if (ncol <= 4) {
        if (mode == 0) return pick_k<0, 4, UNI>(KMAX);
        if (mode == 1) return pick_k<1, 4, UNI>(KMAX);
        return pick_k<2, 4, UNI>(KMAX);
    }
    if (mode == 0) return pick_k<0, 8, UNI>(KMAX);
    if (mode == 1) return pick_k<1, 8, UNI>(KMAX);
    return pick_k<2, 8, UNI>(KMAX);
}
static v2_kernel_t pick(int mode, int ncol, int KMAX, int uni = 0) {
    return uni ? pick_u<1>(mode, ncol, KMAX) : pick_u<0>(mode, ncol, KMAX);
}

int fdm_v2_tx(const sq_fdm *f) {
    int nbmax = 1;
    for (int c = 0; c < f->C; c++) nbmax = std::max(nbmax, f->chi[c] - f->clo[c]);
    return ((nbmax + 31) / 32) * 32;
}

// can the fast path run this (mode, S, T)?
bool fdm_v2_supported(const sq_fdm *f, int mode, int S, int T) {
    if (!f->sym || f->C < 1 || f->C > 8 || !f->slot.p) return false;
    int TX = fdm_v2_tx(f);
    if (TX > 1024 || T < TX || T % TX != 0 || T > 1024) return false;
    int TY = T / TX;
    int nsl = (mode == 2) ? S + 1 : S;
    int KMAX = (nsl + TY - 1) / TY;
    return KMAX <= 4;
}

void fdm_v2_set_attributes(sq_fdm *f) {
    for (int mode = 0; mode < 3; mode++)
        for (int ncol : {4, 8})
            for (int K : {1, 2, 4})
                for (int uni : {0, 1})
                    SQ_CUDA(cudaFuncSetAttribute(pick(mode, ncol, K, uni), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem_optin));
}

struct CgFuseArgs {
    const double2 *d; double2 *pnew; const CgState *cur; CgState *nxt;
    const double *rr_part, *beta_part; int nrr, nbeta, beta_complex, iter, check;
};
static thread_local const CgFuseArgs *g_fuse = nullptr;     // set by fdm_v2_launch_cg around the launch (per host thread: handles may be driven from different threads)

void fdm_v2_launch(sq_fdm *f, int mode, int S, int T, double2 *out, const double2 *in, double *part, const CgState *skip) {
    K2Params P;
    P.cg_d = nullptr; P.cg_pnew = nullptr; P.cg_cur = nullptr; P.cg_nxt = nullptr; P.cg_rr_part = nullptr; P.cg_beta_part = nullptr;
    P.cg_nrr = P.cg_nbeta = P.cg_beta_complex = P.cg_iter = P.cg_check = 0;
    if (g_fuse) {
        P.cg_d = g_fuse->d; P.cg_pnew = g_fuse->pnew; P.cg_cur = g_fuse->cur; P.cg_nxt = g_fuse->nxt;
        P.cg_rr_part = g_fuse->rr_part; P.cg_beta_part = g_fuse->beta_part; P.cg_nrr = g_fuse->nrr; P.cg_nbeta = g_fuse->nbeta;
        P.cg_beta_complex = g_fuse->beta_complex; P.cg_iter = g_fuse->iter; P.cg_check = g_fuse->check;
    }
    P.L = (int)f->L; P.N = (int)f->N; P.Nh = (int)f->Nh; P.C = (int)f->C; P.S = S; P.TX = fdm_v2_tx(f);
    P.lb = f->slab_lo; P.le = f->slab_hi;
    for (int c = 0; c < SQ_MAXC; c++) { P.clo[c] = c < f->C ? f->clo[c] : 0; P.chi[c] = c < f->C ? f->chi[c] : 0; }
    P.nunc0 = f->nunc0;
    static long long *dbg = nullptr;
    if (!dbg && getenv("SQ_DEBUG_STAMPS")) {
        SQ_CUDA(cudaMallocManaged((void **)&dbg, 16 * sizeof(long long)));
        for (int q = 0; q < 16; q++) dbg[q] = 0;
    }
    P.dbg = dbg;
    if (dbg && getenv("SQ_DEBUG_PRINT")) {
        cudaStreamSynchronize(f->stream);
        fprintf(stderr, "stamps(cycles): init %lld load %lld B1 %lld comb %lld B2 %lld store %lld red %lld total %lld\n", dbg[1] - dbg[0],
                dbg[2] - dbg[1], dbg[3] - dbg[2], dbg[4] - dbg[3], dbg[5] - dbg[4], dbg[6] - dbg[5], dbg[7] - dbg[6], dbg[7] - dbg[0]);
    }
    P.nts = f->nts.p; P.nt = f->nt.p; P.slot = f->slot.p; P.unc0 = f->unc0.p; P.cs = f->cs.p; P.expV = f->expV.p;
    int TY = T / P.TX;
    int nsl = (mode == 2) ? S + 1 : S;
    int KMAX = (nsl + TY - 1) / TY;
    KMAX = KMAX <= 1 ? 1 : (KMAX <= 2 ? 2 : 4);
    size_t slices = (mode == 2) ? (size_t)(2 * S + 3) : (size_t)(2 * S + 1);
    size_t smem = slices * f->N * sizeof(double2);
    int grid = (f->slab_hi - f->slab_lo + S - 1) / S;
    v2_kernel_t k = pick(mode, (int)f->C, KMAX, f->cs_uniform);
    SQ_CUDA(sq_launch(k, dim3(grid), dim3(T), smem, f->stream, P, out, in, part, skip));
    SQ_LAUNCH_CHECK();
    f->launches++;
}

// z = M^T M p_new with p_new = d + beta p_old formed on load (see K2Params).  Returns the number of pAp partials.
int fdm_v2_launch_cg(sq_fdm *f, double2 *z, const double2 *p_old, double2 *p_new, const double2 *d, const CgState *cur, CgState *nxt,
                     const double *rr_part, int nrr, const double *beta_part, int nbeta, int beta_complex, int iter, int check,
                     double *pAp_part) {
    CgFuseArgs a = {d, p_new, cur, nxt, rr_part, beta_part, nrr, nbeta, beta_complex, iter, check};
    g_fuse = &a;
    try {
        fdm_v2_launch(f, 2, f->slab, f->threads, z, p_old, pAp_part, cur);
    } catch (...) { g_fuse = nullptr; throw; }
    g_fuse = nullptr;
    return (f->slab_hi - f->slab_lo + f->slab - 1) / f->slab;
}

// ---------------------------------------------------------------------------------------------------
// Persistent cooperative CG: the whole unpreconditioned solve in ONE launch.
//
// The named problem is latency bound (hard part 1): one fused M^T M v is ~14 us of work per CTA while
// every extra launch, host hand-off and separate BLAS-1 pass costs several us.  Here every CTA keeps its tau-slab for
// the whole solve and iterates
//     load p = r + beta p_old (own + halo slices, p written to the other ping-pong buffer)  ->  w = M p, |w|^2 partial
//     -> z = M^T w kept in shared memory  -> [grid barrier] -> alpha -> x += alpha p, r -= alpha z (own slices), |r|^2
//     partial -> [grid barrier] -> eps test, beta
// with two grid-wide barriers per iteration (an atomic counter in L2) instead of kernel boundaries: no launches, no z
// round trip through memory, reductions still fixed-order (bit reproducible).  Launched with
// cudaLaunchCooperativeKernel so that all CTAs are co-resident; vectors other CTAs write are read with ld.global.cg.
// A watchdog turns a barrier that does not complete within ~2 s into an error instead of a hang.
// ---------------------------------------------------------------------------------------------------
struct CgPersist {
    double2 *x, *r, *p0, *p1;
    CgState *state;             // in: rz (= |r0|^2), normb, tol ; out: iters, eps, done
    double *part_a, *part_b;    // per-CTA partials (pAp, rr)
    unsigned int *barrier;      // monotonically increasing arrival counter (zeroed by the host)
    int *abort_flag;
    int maxiter;
};

__device__ __forceinline__ bool grid_barrier(unsigned int *counter, unsigned int &epoch, int *abort_flag) {
    __syncthreads();
    epoch++;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        const unsigned int target = epoch * gridDim.x;
        long long t0 = clock64();
        while (true) {
            unsigned int v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (v >= target) break;
            if (clock64() - t0 > 4000000000LL) { atomicExch(abort_flag, 1); break; }
            __nanosleep(20);
        }
        __threadfence();
    }
    __syncthreads();
    return *((volatile int *)abort_flag) == 0;
}

template <int CMAX, int KMAX, int UNI>
__global__ void __launch_bounds__(1024, 1)
k_cg_persistent(const __grid_constant__ K2Params P, const CgPersist C) {
    extern __shared__ double2 smem[];
    __shared__ double red[32];
    __shared__ double sh[2];
    typename std::conditional<UNI != 0, EngineU<CMAX, KMAX>, Engine<CMAX, KMAX>>::type E;
    E.init(P);
    const int L = P.L, N = P.N, T = blockDim.x;
    const int l0 = P.lb + blockIdx.x * P.S;
    const int ns = min(P.S, P.le - l0);
    double2 *A = smem, *W = smem + (P.S + 2) * N;
    const double normb = C.state->normb, tol = C.state->tol;
    double rz = C.state->rz_re;                       // |r_j|^2 (real for P = I)
    double beta = 0.0, eps = C.state->eps;
    unsigned int epoch = 0;
    int it = 0, done = 0, pc = 0;
    while (it < C.maxiter) {
        it++;
        const double2 *p_old = pc ? C.p1 : C.p0;
        double2 *p_new = pc ? C.p0 : C.p1;
        // ---- load p = r + beta p_old for own + halo slices
        for (int i = threadIdx.x; i < N; i += T) {
            const int s = __ldg(P.slot + i);
            for (int k = 0; k <= ns + 1; k++) {
                int l = l0 - 1 + k;
                l = l < 0 ? l + L : (l >= L ? l - L : l);
                double2 rv = __ldcg(C.r + l * N + i), pv = __ldcg(p_old + l * N + i);
                double2 v = make_double2(fma(beta, pv.x, rv.x), fma(beta, pv.y, rv.y));
                if (k >= 1 && k <= ns) p_new[l * N + i] = v;
                A[k * N + s] = v;
                if (k <= ns) W[k * N + s] = v;
            }
        }
        __syncthreads();
        E.apply_B(W, ns + 1, l0 >= L ? l0 - L : l0, P);
        if (UNI) __syncthreads();
        double acc = 0.0;
        for (int i = threadIdx.x; i < N; i += T) {
            for (int k = 0; k <= ns; k++) {
                int l = l0 + k;
                l = l >= L ? l - L : l;
                double sg = (l == 0) ? 1.0 : -1.0;
                double2 a = A[(k + 1) * N + i], b = W[k * N + i];
                double2 w = make_double2(fma(sg, b.x, a.x), fma(sg, b.y, a.y));
                W[k * N + i] = w;
                if (k >= 1) A[(k + 1) * N + i] = w;
                if (k < ns) acc += w.x * w.x + w.y * w.y;
            }
        }
        __syncthreads();
        {
            int lf = l0 + 1;
            E.apply_B(A + 2 * N, ns, lf >= L ? lf - L : lf, P);
        }
        if (UNI) __syncthreads();
        // z[l0+k] = w[l0+k] -/+ B^T w[l0+k+1], kept in shared memory (W[k], slot order)
        for (int i = threadIdx.x; i < N; i += T) {
            for (int k = 0; k < ns; k++) {
                int lb = l0 + k + 1;
                lb = lb >= L ? lb - L : lb;
                double sg = (lb == 0) ? 1.0 : -1.0;
                double2 a = W[k * N + i], b = A[(k + 2) * N + i];
                W[k * N + i] = make_double2(fma(sg, b.x, a.x), fma(sg, b.y, a.y));
            }
        }
        {
            double v[1] = {acc};
            block_sum<1>(v, red);
            if (threadIdx.x == 0) C.part_a[blockIdx.x] = v[0];
        }
        if (!grid_barrier(C.barrier, epoch, C.abort_flag)) { done = 3; break; }
        if (threadIdx.x < 32) {
            double s = 0.0;
            for (int q = threadIdx.x; q < (int)gridDim.x; q += 32) s += __ldcg(C.part_a + q);
            s = warp_sum(s);
            if (threadIdx.x == 0) sh[0] = s;
        }
        __syncthreads();
        const double alpha = rz / sh[0];
        // ---- x += alpha p ; r -= alpha z (own slices) ; |r|^2 partial
        acc = 0.0;
        for (int i = threadIdx.x; i < N; i += T) {
            const int s = __ldg(P.slot + i);
            for (int k = 0; k < ns; k++) {
                const int g = (l0 + k) * N + i;
                double2 pv = p_new[g], zv = W[k * N + s], xv = C.x[g], rv = __ldcg(C.r + g);
                xv = make_double2(fma(alpha, pv.x, xv.x), fma(alpha, pv.y, xv.y));
                rv = make_double2(fma(-alpha, zv.x, rv.x), fma(-alpha, zv.y, rv.y));
                C.x[g] = xv;
                C.r[g] = rv;
                acc += rv.x * rv.x + rv.y * rv.y;
            }
        }
        {
            double v[1] = {acc};
            block_sum<1>(v, red);
            if (threadIdx.x == 0) C.part_b[blockIdx.x] = v[0];
        }
        if (!grid_barrier(C.barrier, epoch, C.abort_flag)) { done = 3; break; }
        if (threadIdx.x < 32) {
            double s = 0.0;
            for (int q = threadIdx.x; q < (int)gridDim.x; q += 32) s += __ldcg(C.part_b + q);
            s = warp_sum(s);
            if (threadIdx.x == 0) sh[1] = s;
        }
        __syncthreads();
        const double rr = sh[1];
        eps = sqrt(rr) / normb;
        if (eps < tol) { done = 1; break; }
        if (!(eps == eps)) { done = 2; break; }
        beta = rr / rz;
        rz = rr;
        pc ^= 1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        CgState s = *C.state;
        s.iters = it;
        s.eps = eps;
        s.done = done;
        s.rz_re = rz;
        *C.state = s;
    }
}

typedef void (*persist_kernel_t)(const K2Params, const CgPersist);
template <int CMAX, int UNI>
static persist_kernel_t pick_pk(int KMAX) {
    switch (KMAX) {
        case 1: return k_cg_persistent<CMAX, 1, UNI>;
        case 2: return k_cg_persistent<CMAX, 2, UNI>;
        default: return k_cg_persistent<CMAX, 4, UNI>;
    }
}
static persist_kernel_t pick_persist(int ncol, int KMAX, int uni) {
    if (ncol <= 4) return uni ? pick_pk<4, 1>(KMAX) : pick_pk<4, 0>(KMAX);
    return uni ? pick_pk<8, 1>(KMAX) : pick_pk<8, 0>(KMAX);
}

// Runs the persistent solve if all CTAs of the current tuning can be co-resident.  x, r, state are prepared by the
// caller (cg.cu); p0/p1 are the two p buffers (zeroed here).  Returns false if the cooperative launch is not possible.
bool fdm_v2_cg_persistent(sq_fdm *f, double2 *x, double2 *r, double2 *p0, double2 *p1, CgState *state, double *part_a, double *part_b,
                          i64 maxiter) {
    const int S = f->slab, T = f->threads;
    if (!fdm_v2_supported(f, 2, S, T)) return false;
    K2Params P;
    memset(&P, 0, sizeof(P));
    P.L = (int)f->L; P.N = (int)f->N; P.Nh = (int)f->Nh; P.C = (int)f->C; P.S = S; P.TX = fdm_v2_tx(f);
    P.lb = f->slab_lo; P.le = f->slab_hi;
    for (int c = 0; c < SQ_MAXC; c++) { P.clo[c] = c < f->C ? f->clo[c] : 0; P.chi[c] = c < f->C ? f->chi[c] : 0; }
    P.nunc0 = f->nunc0;
    P.nts = f->nts.p; P.nt = f->nt.p; P.slot = f->slot.p; P.unc0 = f->unc0.p; P.cs = f->cs.p; P.expV = f->expV.p;
    int TY = T / P.TX, nsl = S + 1;
    int KMAX = (nsl + TY - 1) / TY;
    KMAX = KMAX <= 1 ? 1 : (KMAX <= 2 ? 2 : 4);
    size_t smem = (size_t)(2 * S + 3) * f->N * sizeof(double2);
    int grid = (f->slab_hi - f->slab_lo + S - 1) / S;
    persist_kernel_t k = pick_persist((int)f->C, KMAX, f->cs_uniform);
    SQ_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem_optin));
    int per_sm = 0;
    SQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, T, smem));
    if (per_sm * f->num_sms < grid) return false;
    if (!f->flag.p) f->flag.alloc(4);
    SQ_CUDA(cudaMemsetAsync(f->flag.p, 0, 4 * sizeof(int), f->stream));
    SQ_CUDA(cudaMemsetAsync(p0, 0, f->vec_bytes(), f->stream));
    SQ_CUDA(cudaMemsetAsync(p1, 0, f->vec_bytes(), f->stream));
    CgPersist C;
    C.x = x; C.r = r; C.p0 = p0; C.p1 = p1; C.state = state; C.part_a = part_a; C.part_b = part_b;
    C.barrier = (unsigned int *)f->flag.p; C.abort_flag = f->flag.p + 1; C.maxiter = (int)std::min<i64>(maxiter, 2000000000);
    void *args[] = {(void *)&P, (void *)&C};
    SQ_CUDA(cudaLaunchCooperativeKernel((const void *)k, dim3(grid), dim3(T), args, smem, f->stream));
    f->launches++;
    return true;
}
