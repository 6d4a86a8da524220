// kpm_reg.cu -- K4 on the register engine: the Chebyshev recurrences of the KPM preconditioner for the lattices of the
// register path (rectangular Lx x Ly with the canonical 4 colours, L x L honeycomb with its 3 bond types; colour-uniform,
// tau-independent hoppings -- every Holstein-type model on these lattices).
//
// Replaces the `kpm_lmul!` calls of src/KPMPreconditioner.jl:394 (Sym, complex vectors) for those lattices.  What bounds the
// preconditioner is not throughput but the LATENCY of its longest recurrence: at cfg4 the lowest Matsubara frequency needs
// order ~ 160 sequential applications of B-bar to one N-vector while all frequencies together are only ~ 2500 applications
// (1.6 us of FP64 issue on the whole chip).  The shared-memory kernel (kpm.cu, k_kpm_cheb_fast) spends ~ 0.9 us per
// application: one site pair per thread, 2C - 2 CTA-wide barriers and shared-memory round trips.  Here one chain (frequency,
// real or imaginary part -- B-bar is real, so the two parts are independent recurrences) is ONE CTA of W warps that holds its
// N-vector in registers for the whole recurrence:
//
//   * square lattice: lane = (group of 4 consecutive x, block of 2 rows), warp w owns 2 (32 / LXL) consecutive rows; x-even and
//     y-even bonds are lane-local, x-odd bonds one shuffle per two sites, y-odd bonds shuffle the lane's two rows -- except across
//     the W warps of the chain, where the boundary rows cross through a small double-buffered shared-memory window (one CTA barrier);
//   * honeycomb: lane = R1 x R2 block of cells, the intra-cell bond is lane-local, the two inter-cell bond types shuffle one
//     boundary column / row (rows across warps through the same window);
//   * scaled rotations a' = a + tanh b with prod_c cosh_c^2, the Chebyshev rescaling 2 / mag and the tau-mean diagonal folded
//     into ONE per-site factor kept in registers;
//   * the recurrence runs in the frame rotated by the OUTER colour step K (B-bar = K A K  =>  K B-bar K^-1 = K^2 A, and K^2 is one
//     step with the doubled angle): one outer step -- the expensive one, it crosses lanes and warps -- per application instead of
//     two, i.e. 2C - 1 colour steps + the diagonal (one DFMA per site each) + 3 DFMA per site for T_{q+1} = 2 B' T_q - T_{q-1} and
//     the accumulation; the result is rotated back once per chain;
//   * W is chosen so that a lane holds 4 - 12 sites: T_{q-1}, T_q, the working copy, the accumulator and the diagonal fit in
//     registers, and the four FP64 pipes of the SM all work on the one chain that matters.
// Chains are scheduled longest first, one CTA per SM (CTA b runs chains b, b + G, ...: the long chains get an SM to themselves).
// The same kernel serves a batch of right-hand sides (multi-RHS solves: chain = (rhs, frequency, part)).
#include "sq_internal.h"

#include <algorithm>

struct ChebRegParams {
    int N, L;
    int nchain;                     // nsched * 2 * nrhs
    int nsched, nrhs;
    size_t rhs_stride;              // double2 elements between the frequency arrays of consecutive right-hand sides
    double2 *z;                     // [rhs][n][i]
    const int *sched, *order, *coef_off;
    const double2 *coefs;
    const double2 *csbar;           // (mean cosh, mean sinh) per bond; colour c = csbar[clo[c]]
    int clo[4];
    const double *Dbar;             // [i]
    const int *site_bond;           // per-bond engines: internal bond index of site i in colour c at [c * N + i]
    double avg, imag_;              // (emax + emin) / 2, 2 / (emax - emin)
    const CgState *skip;
};

// ---- square lattice -------------------------------------------------------------------------------------------------
template <int LXL_, int W_>
struct RegSquare {
    static constexpr int LXL = LXL_, W = W_, RY = 2, YH = 32 / LXL, LX = 4 * LXL, LY = RY * YH * W, N = LX * LY, NV = 4 * RY, NCOL = 4;
    static constexpr int XCH = 2 * W * 2 * LX;          // doubles of the exchange window: [buffer][warp][bottom / top row][x]
    int xl, yh, w, lane_r, lane_l, lane_u, lane_d;
    double t[4], t2o;           // tanh per colour; doubled-angle tanh of the outer colour

    __device__ __forceinline__ void init(const ChebRegParams &P, double &g) {
        const int lane = threadIdx.x & 31;
        w = threadIdx.x >> 5;
        xl = lane % LXL;
        yh = lane / LXL;
        lane_r = lane - xl + (xl + 1) % LXL;
        lane_l = lane - xl + (xl + LXL - 1) % LXL;
        lane_u = xl + LXL * ((yh + 1) % YH);
        lane_d = xl + LXL * ((yh + YH - 1) % YH);
        g = 1.0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const double2 q = __ldg(P.csbar + P.clo[c]);
            t[c] = q.y / q.x;
            g *= q.x * q.x;
        }
        t2o = 2.0 * t[3] / (1.0 + t[3] * t[3]);
        g *= 1.0 + t[3] * t[3];                                   // K^2 = (1 + t^2)(1 + t' sigma)
    }
    // site of value k = 4 r + j
    __device__ __forceinline__ int site(int k) const { return 4 * xl + (k & 3) + LX * ((w * YH + yh) * RY + (k >> 2)); }

    template <int CL>
    __device__ __forceinline__ void step(double (&v)[NV], double *xb, double s_outer = 0.0) const {
        const double s = (CL == NCOL - 1) ? s_outer : t[CL];
        if (CL == 0) {                                            // x-even: (0,1), (2,3)
#pragma unroll
            for (int r = 0; r < RY; r++) {
                const double a = v[4 * r], b = v[4 * r + 1], c = v[4 * r + 2], d = v[4 * r + 3];
                v[4 * r] = fma(s, b, a); v[4 * r + 1] = fma(s, a, b); v[4 * r + 2] = fma(s, d, c); v[4 * r + 3] = fma(s, c, d);
            }
        } else if (CL == 1) {                                     // x-odd: (1,2) inside, (3 | 0 of the right neighbour)
#pragma unroll
            for (int r = 0; r < RY; r++) {
                const double fromR = __shfl_sync(0xffffffffu, v[4 * r], lane_r);
                const double fromL = __shfl_sync(0xffffffffu, v[4 * r + 3], lane_l);
                const double b = v[4 * r + 1], c = v[4 * r + 2];
                v[4 * r + 1] = fma(s, c, b); v[4 * r + 2] = fma(s, b, c);
                v[4 * r + 3] = fma(s, fromR, v[4 * r + 3]);
                v[4 * r] = fma(s, fromL, v[4 * r]);
            }
        } else if (CL == 2) {                                     // y-even: rows (0, 1) of the lane
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const double a = v[j], b = v[4 + j];
                v[j] = fma(s, b, a); v[4 + j] = fma(s, a, b);
            }
        } else {                                                  // y-odd: row 1 with row 0 of the block above
            double up[4], dn[4];
            if (W > 1) {                                          // rows that cross a warp boundary go through shared memory
                if (yh == 0) *reinterpret_cast<double4 *>(xb + (w * 2 + 0) * LX + 4 * xl) = make_double4(v[0], v[1], v[2], v[3]);
                if (yh == YH - 1) *reinterpret_cast<double4 *>(xb + (w * 2 + 1) * LX + 4 * xl) = make_double4(v[4], v[5], v[6], v[7]);
                __syncthreads();
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                up[j] = __shfl_sync(0xffffffffu, v[j], lane_u);
                dn[j] = __shfl_sync(0xffffffffu, v[4 + j], lane_d);
            }
            if (W > 1) {
                if (yh == YH - 1) {
                    const double4 q = *reinterpret_cast<const double4 *>(xb + (((w + 1) % W) * 2 + 0) * LX + 4 * xl);
                    up[0] = q.x; up[1] = q.y; up[2] = q.z; up[3] = q.w;
                }
                if (yh == 0) {
                    const double4 q = *reinterpret_cast<const double4 *>(xb + (((w + W - 1) % W) * 2 + 1) * LX + 4 * xl);
                    dn[0] = q.x; dn[1] = q.y; dn[2] = q.z; dn[3] = q.w;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                v[4 + j] = fma(s, up[j], v[4 + j]);
                v[j] = fma(s, dn[j], v[j]);
            }
        }
    }
    // One application in the ROTATED frame u = K t, K = (1 + t3 sigma3) the outer (y-odd) colour step: B-bar = K A K with
    // A = colours 2, 1, 0, diagonal, 0, 1, 2, so K B-bar K^-1 = K^2 A and K^2 = (1 + t3^2)(1 + t3' sigma3), t3' = 2 t3 / (1 + t3^2):
    // ONE outer step with the doubled angle per application instead of two -- the outer step is the expensive one (it crosses lanes
    // and warps).  xb: this application's half of the exchange window.
    __device__ __forceinline__ void apply(double (&v)[NV], const double (&dg)[NV], double *xb) const {
        step<2>(v, xb); step<1>(v, xb); step<0>(v, xb);
#pragma unroll
        for (int k = 0; k < NV; k++) v[k] *= dg[k];
        step<0>(v, xb); step<1>(v, xb); step<2>(v, xb);
        step<3>(v, xb, t2o);
    }
    __device__ __forceinline__ void outer_in(double (&v)[NV], double *xb) const { step<3>(v, xb, t[3]); }          // u = K v
    __device__ __forceinline__ void outer_out(double (&v)[NV], double *xb) const {                                   // v = K^-1 u
        step<3>(v, xb, -t[3]);
        const double back = 1.0 / (1.0 - t[3] * t[3]);
#pragma unroll
        for (int k = 0; k < NV; k++) v[k] *= back;
    }
    static constexpr int XHALF = W * 2 * LX;
};

// ---- honeycomb --------------------------------------------------------------------------------------------------------
// site = orb + 2 (c1 + L1 c2); colour 0: A(c1,c2)-B(c1,c2), colour 1: A(c1,c2)-B(c1-1,c2), colour 2: A(c1,c2)-B(c1,c2-1)
// lane = g1 + G1 g2 of warp w holds the cells c1 = R1 g1 + a1, c2 = R2 (G2 w + g2) + a2; value k = 2 (a2 R1 + a1) + orb
template <int G1_, int R1_, int R2_, int W_>
struct RegHoney {
    static constexpr int G1 = G1_, R1 = R1_, R2 = R2_, W = W_, G2 = 32 / G1, L1 = G1 * R1, L2 = G2 * R2 * W;
    static constexpr int NP = R1 * R2, NV = 2 * NP, N = 2 * L1 * L2, NCOL = 3;
    static constexpr int XCH = 2 * W * 2 * L1;
    int g1, g2, w, lane_r, lane_l, lane_u, lane_d;
    double t[3], t2o;

    __device__ __forceinline__ void init(const ChebRegParams &P, double &g) {
        const int lane = threadIdx.x & 31;
        w = threadIdx.x >> 5;
        g1 = lane % G1;
        g2 = lane / G1;
        lane_r = (g1 + 1) % G1 + G1 * g2;
        lane_l = (g1 + G1 - 1) % G1 + G1 * g2;
        lane_u = g1 + G1 * ((g2 + 1) % G2);
        lane_d = g1 + G1 * ((g2 + G2 - 1) % G2);
        g = 1.0;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double2 q = __ldg(P.csbar + P.clo[c]);
            t[c] = q.y / q.x;
            g *= q.x * q.x;
        }
        t2o = 2.0 * t[2] / (1.0 + t[2] * t[2]);
        g *= 1.0 + t[2] * t[2];
    }
    __device__ __forceinline__ int site(int k) const {
        const int u = k >> 1, a1 = u % R1, a2 = u / R1;
        return (k & 1) + 2 * ((R1 * g1 + a1) + L1 * (R2 * (G2 * w + g2) + a2));
    }
    static __device__ __forceinline__ void rot(double &a, double &b, double s) {
        const double na = fma(s, b, a), nb = fma(s, a, b);
        a = na;
        b = nb;
    }
    template <int CL>
    __device__ __forceinline__ void step(double (&v)[NV], double *xb, double s_outer = 0.0) const {
        const double s = (CL == NCOL - 1) ? s_outer : t[CL];
        if (CL == 0) {
#pragma unroll
            for (int u = 0; u < NP; u++) rot(v[2 * u], v[2 * u + 1], s);
        } else if (CL == 1) {                             // A(a1, a2) - B(a1 - 1, a2): first column with the left lane's last column
#pragma unroll
            for (int a2 = 0; a2 < R2; a2++) {
                const int u0 = a2 * R1, u1 = a2 * R1 + R1 - 1;
                const double fromL = __shfl_sync(0xffffffffu, v[2 * u1 + 1], lane_l);     // B of the left lane's last column
                const double fromR = __shfl_sync(0xffffffffu, v[2 * u0], lane_r);         // A of the right lane's first column
#pragma unroll
                for (int a1 = R1 - 1; a1 >= 1; a1--) rot(v[2 * (u0 + a1)], v[2 * (u0 + a1 - 1) + 1], s);
                v[2 * u0] = fma(s, fromL, v[2 * u0]);
                v[2 * u1 + 1] = fma(s, fromR, v[2 * u1 + 1]);
            }
        } else {                                          // A(a1, a2) - B(a1, a2 - 1): first row with the row below
            double fromD[R1], fromU[R1];
            if (W > 1) {
                if (g2 == 0) {
#pragma unroll
                    for (int a1 = 0; a1 < R1; a1++) xb[(w * 2 + 0) * L1 + R1 * g1 + a1] = v[2 * a1];                              // A of the first row
                }
                if (g2 == G2 - 1) {
#pragma unroll
                    for (int a1 = 0; a1 < R1; a1++) xb[(w * 2 + 1) * L1 + R1 * g1 + a1] = v[2 * ((R2 - 1) * R1 + a1) + 1];        // B of the last row
                }
                __syncthreads();
            }
#pragma unroll
            for (int a1 = 0; a1 < R1; a1++) {
                fromD[a1] = __shfl_sync(0xffffffffu, v[2 * ((R2 - 1) * R1 + a1) + 1], lane_d);
                fromU[a1] = __shfl_sync(0xffffffffu, v[2 * a1], lane_u);
            }
            if (W > 1) {
                if (g2 == 0) {
#pragma unroll
                    for (int a1 = 0; a1 < R1; a1++) fromD[a1] = xb[(((w + W - 1) % W) * 2 + 1) * L1 + R1 * g1 + a1];
                }
                if (g2 == G2 - 1) {
#pragma unroll
                    for (int a1 = 0; a1 < R1; a1++) fromU[a1] = xb[(((w + 1) % W) * 2 + 0) * L1 + R1 * g1 + a1];
                }
            }
#pragma unroll
            for (int a1 = 0; a1 < R1; a1++) {
#pragma unroll
                for (int a2 = R2 - 1; a2 >= 1; a2--) rot(v[2 * (a2 * R1 + a1)], v[2 * ((a2 - 1) * R1 + a1) + 1], s);
                v[2 * a1] = fma(s, fromD[a1], v[2 * a1]);
                v[2 * ((R2 - 1) * R1 + a1) + 1] = fma(s, fromU[a1], v[2 * ((R2 - 1) * R1 + a1) + 1]);
            }
        }
    }
    // rotated frame, as RegSquare::apply: colours 1, 0, diagonal, 0, 1, then the outer colour 2 once with the doubled angle
    __device__ __forceinline__ void apply(double (&v)[NV], const double (&dg)[NV], double *xb) const {
        step<1>(v, xb); step<0>(v, xb);
#pragma unroll
        for (int k = 0; k < NV; k++) v[k] *= dg[k];
        step<0>(v, xb); step<1>(v, xb);
        step<2>(v, xb, t2o);
    }
    __device__ __forceinline__ void outer_in(double (&v)[NV], double *xb) const { step<2>(v, xb, t[2]); }
    __device__ __forceinline__ void outer_out(double (&v)[NV], double *xb) const {
        step<2>(v, xb, -t[2]);
        const double back = 1.0 / (1.0 - t[2] * t[2]);
#pragma unroll
        for (int k = 0; k < NV; k++) v[k] *= back;
    }
    static constexpr int XHALF = W * 2 * L1;
};

// ---- per-bond coefficients (SSH models, disordered hoppings): square lattice and chain ----------------------------------------
// B-bar's bonds carry their own tau-mean (cosh, sinh) -- no common factor to fold, so a colour step is c a + s b (DMUL + DFMA per
// site) with the coefficients of the lane's 8 sites x 4 colours kept in registers for the whole recurrence.  The frame rotation by
// the outer colour works bond by bond: K = c + s sigma, K^2 = (c^2 + s^2) + 2 c s sigma, K^-1 = (c - s sigma) / (c^2 - s^2) (the
// tau-means do not satisfy c^2 - s^2 = 1).
template <int LXL_, int W_>
struct RegSquarePB {
    static constexpr int LXL = LXL_, W = W_, RY = 2, YH = 32 / LXL, LX = 4 * LXL, LY = RY * YH * W, N = LX * LY, NV = 4 * RY, NCOL = 4;
    static constexpr int XCH = 2 * W * 2 * LX, XHALF = W * 2 * LX;
    int xl, yh, w, lane_r, lane_l, lane_u, lane_d;
    double2 cs[3][NV];          // (c, s) of the bond that touches value k in the inner colours 0, 1, 2
    double2 k2[NV];             // outer colour squared: (c^2 + s^2, 2 c s); K itself is re-read for the two frame changes of a chain
    const double2 *csbar;
    const int *sb3;             // site -> bond of the outer colour

    __device__ __forceinline__ int site(int k) const { return 4 * xl + (k & 3) + LX * ((w * YH + yh) * RY + (k >> 2)); }
    __device__ __forceinline__ void init(const ChebRegParams &P, double &g) {
        const int lane = threadIdx.x & 31;
        w = threadIdx.x >> 5;
        xl = lane % LXL;
        yh = lane / LXL;
        lane_r = lane - xl + (xl + 1) % LXL;
        lane_l = lane - xl + (xl + LXL - 1) % LXL;
        lane_u = xl + LXL * ((yh + 1) % YH);
        lane_d = xl + LXL * ((yh + YH - 1) % YH);
        g = 1.0;
        csbar = P.csbar;
        sb3 = P.site_bond + 3 * N;
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int k = 0; k < NV; k++) cs[c][k] = __ldg(P.csbar + __ldg(P.site_bond + c * N + site(k)));
#pragma unroll
        for (int k = 0; k < NV; k++) {
            const double2 q = __ldg(csbar + __ldg(sb3 + site(k)));
            k2[k] = make_double2(q.x * q.x + q.y * q.y, 2.0 * q.x * q.y);
        }
    }
    // one colour step; the outer colour takes its coefficients from `co` (K, K^2 or K^-1)
    template <int CL>
    __device__ __forceinline__ void step(double (&v)[NV], double *xb, const double2 (&co)[NV]) const {
        if (CL == 0) {
#pragma unroll
            for (int r = 0; r < RY; r++) {
                const double a = v[4 * r], b = v[4 * r + 1], c = v[4 * r + 2], d = v[4 * r + 3];
                v[4 * r] = fma(co[4 * r].y, b, co[4 * r].x * a); v[4 * r + 1] = fma(co[4 * r + 1].y, a, co[4 * r + 1].x * b);
                v[4 * r + 2] = fma(co[4 * r + 2].y, d, co[4 * r + 2].x * c); v[4 * r + 3] = fma(co[4 * r + 3].y, c, co[4 * r + 3].x * d);
            }
        } else if (CL == 1) {
#pragma unroll
            for (int r = 0; r < RY; r++) {
                const double fromR = __shfl_sync(0xffffffffu, v[4 * r], lane_r);
                const double fromL = __shfl_sync(0xffffffffu, v[4 * r + 3], lane_l);
                const double b = v[4 * r + 1], c = v[4 * r + 2];
                v[4 * r + 1] = fma(co[4 * r + 1].y, c, co[4 * r + 1].x * b); v[4 * r + 2] = fma(co[4 * r + 2].y, b, co[4 * r + 2].x * c);
                v[4 * r + 3] = fma(co[4 * r + 3].y, fromR, co[4 * r + 3].x * v[4 * r + 3]);
                v[4 * r] = fma(co[4 * r].y, fromL, co[4 * r].x * v[4 * r]);
            }
        } else if (CL == 2) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const double a = v[j], b = v[4 + j];
                v[j] = fma(co[j].y, b, co[j].x * a); v[4 + j] = fma(co[4 + j].y, a, co[4 + j].x * b);
            }
        } else {
            double up[4], dn[4];
            if (W > 1) {
                if (yh == 0) *reinterpret_cast<double4 *>(xb + (w * 2 + 0) * LX + 4 * xl) = make_double4(v[0], v[1], v[2], v[3]);
                if (yh == YH - 1) *reinterpret_cast<double4 *>(xb + (w * 2 + 1) * LX + 4 * xl) = make_double4(v[4], v[5], v[6], v[7]);
                __syncthreads();
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                up[j] = __shfl_sync(0xffffffffu, v[j], lane_u);
                dn[j] = __shfl_sync(0xffffffffu, v[4 + j], lane_d);
            }
            if (W > 1) {
                if (yh == YH - 1) {
                    const double4 q = *reinterpret_cast<const double4 *>(xb + (((w + 1) % W) * 2 + 0) * LX + 4 * xl);
                    up[0] = q.x; up[1] = q.y; up[2] = q.z; up[3] = q.w;
                }
                if (yh == 0) {
                    const double4 q = *reinterpret_cast<const double4 *>(xb + (((w + W - 1) % W) * 2 + 1) * LX + 4 * xl);
                    dn[0] = q.x; dn[1] = q.y; dn[2] = q.z; dn[3] = q.w;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                v[4 + j] = fma(co[4 + j].y, up[j], co[4 + j].x * v[4 + j]);
                v[j] = fma(co[j].y, dn[j], co[j].x * v[j]);
            }
        }
    }
    __device__ __forceinline__ void apply(double (&v)[NV], const double (&dg)[NV], double *xb) const {
        step<2>(v, xb, cs[2]); step<1>(v, xb, cs[1]); step<0>(v, xb, cs[0]);
#pragma unroll
        for (int k = 0; k < NV; k++) v[k] *= dg[k];
        step<0>(v, xb, cs[0]); step<1>(v, xb, cs[1]); step<2>(v, xb, cs[2]);
        step<3>(v, xb, k2);
    }
    __device__ __forceinline__ void outer_in(double (&v)[NV], double *xb) const {                                    // u = K v
        double2 co[NV];
#pragma unroll
        for (int k = 0; k < NV; k++) co[k] = __ldg(csbar + __ldg(sb3 + site(k)));
        step<3>(v, xb, co);
    }
    __device__ __forceinline__ void outer_out(double (&v)[NV], double *xb) const {                                   // v = K^-1 u
        double2 co[NV];
#pragma unroll
        for (int k = 0; k < NV; k++) {
            const double2 q = __ldg(csbar + __ldg(sb3 + site(k)));
            const double inv = 1.0 / (q.x * q.x - q.y * q.y);
            co[k] = make_double2(q.x * inv, -q.y * inv);
        }
        step<3>(v, xb, co);
    }
};

// periodic chain of N = 64 W sites, colour 0 = even bonds (2m, 2m+1), colour 1 = odd bonds (2m+1, 2m+2): lane holds one even bond
template <int W_>
struct RegChainPB {
    static constexpr int W = W_, N = 64 * W, NV = 2, NCOL = 2;
    static constexpr int XCH = 2 * W * 2, XHALF = W * 2;
    int lane, w;
    double2 cs[2][NV], k2[NV];
    double kinv[NV];
    __device__ __forceinline__ int site(int k) const { return 2 * (32 * w + lane) + k; }
    __device__ __forceinline__ void init(const ChebRegParams &P, double &g) {
        lane = threadIdx.x & 31;
        w = threadIdx.x >> 5;
        g = 1.0;
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
            for (int k = 0; k < NV; k++) cs[c][k] = __ldg(P.csbar + __ldg(P.site_bond + c * N + site(k)));
#pragma unroll
        for (int k = 0; k < NV; k++) {
            const double c = cs[1][k].x, s_ = cs[1][k].y;
            k2[k] = make_double2(c * c + s_ * s_, 2.0 * c * s_);
            kinv[k] = 1.0 / (c * c - s_ * s_);
        }
    }
    // odd bonds: value 1 pairs with value 0 of the next lane (next warp through the window), value 0 with value 1 of the previous
    __device__ __forceinline__ void odd(double (&v)[NV], double *xb, const double2 (&co)[NV]) const {
        if (W > 1) {
            if (lane == 0) xb[w * 2 + 0] = v[0];
            if (lane == 31) xb[w * 2 + 1] = v[1];
            __syncthreads();
        }
        double nx = __shfl_sync(0xffffffffu, v[0], (lane + 1) & 31), pv = __shfl_sync(0xffffffffu, v[1], (lane + 31) & 31);
        if (W > 1) {
            if (lane == 31) nx = xb[((w + 1) % W) * 2 + 0];
            if (lane == 0) pv = xb[((w + W - 1) % W) * 2 + 1];
        }
        const double a = v[0], b = v[1];
        v[1] = fma(co[1].y, nx, co[1].x * b);
        v[0] = fma(co[0].y, pv, co[0].x * a);
    }
    __device__ __forceinline__ void even(double (&v)[NV]) const {
        const double a = v[0], b = v[1];
        v[0] = fma(cs[0][0].y, b, cs[0][0].x * a);
        v[1] = fma(cs[0][1].y, a, cs[0][1].x * b);
    }
    __device__ __forceinline__ void apply(double (&v)[NV], const double (&dg)[NV], double *xb) const {
        even(v);
        v[0] *= dg[0]; v[1] *= dg[1];
        even(v);
        odd(v, xb, k2);
    }
    __device__ __forceinline__ void outer_in(double (&v)[NV], double *xb) const { odd(v, xb, cs[1]); }
    __device__ __forceinline__ void outer_out(double (&v)[NV], double *xb) const {
        double2 ci[NV];
#pragma unroll
        for (int k = 0; k < NV; k++) ci[k] = make_double2(cs[1][k].x * kinv[k], -cs[1][k].y * kinv[k]);
        odd(v, xb, ci);
    }
};

// One CTA = one chain at a time (W warps).  The recurrence runs in the rotated frame u_q = K T_q (see apply()):
//     y'' = (2 / mag) K B-bar K^-1 u_q  comes out of apply() through the folded diagonal,
//     u_1 = (y'' - kappa u_0) / 2,   u_{q+1} = (y'' - kappa u_q) - u_{q-1},   kappa = 2 avg / mag,
// and the result is rotated back once: sum_q c_q T_q v = K^-1 sum_q c_q u_q,  K^-1 = (1 - t sigma) / (1 - t^2).
template <class G>
__global__ void __launch_bounds__(32 * G::W, 1)
k_kpm_cheb_reg(const __grid_constant__ ChebRegParams P) {
    constexpr int NV = G::NV;
    __shared__ __align__(32) double xch[G::XCH > 0 ? G::XCH : 1];
    sq_pdl_prologue();
    if (P.skip && P.nrhs == 1 && P.skip->done) return;
    G E;
    double g;
    E.init(P, g);
    const double kappa = 2.0 * P.avg * P.imag_, dscale = 2.0 * P.imag_ * g;       // g: whatever the engine folds into the diagonal
    double dg[NV];
    int off[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) {
        off[k] = E.site(k);
        dg[k] = dscale * __ldg(P.Dbar + off[k]);
    }
    int flip = 0;                                               // which half of the exchange window the next outer step uses
#define XB() (xch + ((flip ^= 1) ? G::XHALF : 0))
#pragma unroll 1
    for (int ch = blockIdx.x; ch < P.nchain; ch += gridDim.x) {
        // chains are ordered longest first: (schedule index, part, rhs) with the schedule index slowest
        const int si = ch / (2 * P.nrhs), rem = ch - si * 2 * P.nrhs, part = rem & 1, rhs = rem >> 1;
        if (P.skip && P.skip[rhs].done) continue;                 // CTA-uniform: all warps skip the chain together
        const int n = __ldg(P.sched + si);
        const int np = (n + 1 > (P.L + 1) / 2) ? P.L - 1 - n : n;                     // KPMPreconditioner.jl:387
        const int ord = __ldg(P.order + np);
        const double2 *c = P.coefs + __ldg(P.coef_off + np);                           // real coefficients (Sym): .x
        double *zn = reinterpret_cast<double *>(P.z + (size_t)rhs * P.rhs_stride + (size_t)n * P.N) + part;
        double ta[NV], tb[NV], y[NV], acc[NV];
#pragma unroll
        for (int k = 0; k < NV; k++) ta[k] = zn[2 * off[k]];
        const double c0 = __ldg(&c[0].x), c1 = __ldg(&c[1].x);
        E.outer_in(ta, XB());                                   // u_0 = K T_0
#pragma unroll
        for (int k = 0; k < NV; k++) y[k] = ta[k];
        E.apply(y, dg, XB());
#pragma unroll
        for (int k = 0; k < NV; k++) {
            tb[k] = 0.5 * fma(-kappa, ta[k], y[k]);
            acc[k] = fma(c1, tb[k], c0 * ta[k]);
        }
        // two orders per trip, so that (u_{q-1}, u_q) swap roles without register moves
        // the coefficients of a trip are loaded one trip ahead (an L2 hit costs as much as two colour steps)
        int q = 2;
        double na = (q < ord) ? __ldg(&c[q].x) : 0.0, nb = (q + 1 < ord) ? __ldg(&c[q + 1].x) : 0.0;
#pragma unroll 1
        for (; q + 1 < ord; q += 2) {
            const double ca = na, cb = nb;
            if (q + 2 < ord) na = __ldg(&c[q + 2].x);
            if (q + 3 < ord) nb = __ldg(&c[q + 3].x);
#pragma unroll
            for (int k = 0; k < NV; k++) y[k] = tb[k];
            E.apply(y, dg, XB());
#pragma unroll
            for (int k = 0; k < NV; k++) {
                ta[k] = fma(-kappa, tb[k], y[k]) - ta[k];
                acc[k] = fma(ca, ta[k], acc[k]);
                y[k] = ta[k];
            }
            E.apply(y, dg, XB());
#pragma unroll
            for (int k = 0; k < NV; k++) {
                tb[k] = fma(-kappa, ta[k], y[k]) - tb[k];
                acc[k] = fma(cb, tb[k], acc[k]);
            }
        }
        if (q < ord) {
            const double ca = na;
#pragma unroll
            for (int k = 0; k < NV; k++) y[k] = tb[k];
            E.apply(y, dg, XB());
#pragma unroll
            for (int k = 0; k < NV; k++) acc[k] = fma(ca, fma(-kappa, tb[k], y[k]) - ta[k], acc[k]);
        }
        E.outer_out(acc, XB());                                 // back to the original frame
#pragma unroll
        for (int k = 0; k < NV; k++) zn[2 * off[k]] = acc[k];
    }
#undef XB
}

typedef void (*cheb_reg_t)(const ChebRegParams);
struct ChebRegPick { cheb_reg_t k; int threads; };

// periodic chain in natural order with colour 0 = even bonds, colour 1 = odd bonds, N a multiple of 64 (one lane per even bond)
static bool chain_geometry(const sq_fdm *f) {
    if (!f->sym || f->C != 2 || f->Nh != f->N || f->N % 64 || f->N / 64 > 8) return false;
    for (int c = 0; c < 2; c++) {
        if (f->chi[c] - f->clo[c] != f->N / 2) return false;
        for (int h = f->clo[c]; h < f->chi[c]; h++) {
            int i = f->h_nt[h].x, j = f->h_nt[h].y;
            if ((j + 1) % f->N == i) std::swap(i, j);
            if ((i + 1) % f->N != j || (i & 1) != c) return false;
        }
    }
    return true;
}

// perbond = false: colour-uniform coefficients (scaled rotations); true: every bond its own (cosh, sinh) means
static ChebRegPick pick_cheb_reg(const sq_fdm *f, bool perbond) {
    if (!f->sym) return {nullptr, 0};
    if (perbond && chain_geometry(f)) {
        switch ((int)(f->N / 64)) {
            case 1: return {k_kpm_cheb_reg<RegChainPB<1>>, 32};
            case 2: return {k_kpm_cheb_reg<RegChainPB<2>>, 64};
            case 4: return {k_kpm_cheb_reg<RegChainPB<4>>, 128};
            case 8: return {k_kpm_cheb_reg<RegChainPB<8>>, 256};
        }
        return {nullptr, 0};
    }
    if (!f->v3_ok) return {nullptr, 0};
    if (f->v3_kind == 0) {                                     // square: v3_lxl = Lx / 4, v3_ry = rows per lane of the matvec engine
        const int lxl = f->v3_lxl, w = f->v3_ry / 2;           // Ly = v3_ry (32 / lxl) = 2 (32 / lxl) W
        if (perbond) {
            if (lxl == 8 && w == 2) return {k_kpm_cheb_reg<RegSquarePB<8, 2>>, 64};
            if (lxl == 8 && w == 4) return {k_kpm_cheb_reg<RegSquarePB<8, 4>>, 128};
            if (lxl == 4 && w == 1) return {k_kpm_cheb_reg<RegSquarePB<4, 1>>, 32};  // 16 x 16 (cfg3)
            if (lxl == 4 && w == 2) return {k_kpm_cheb_reg<RegSquarePB<4, 2>>, 64};
            return {nullptr, 0};
        }
        if (lxl == 8 && w == 2) return {k_kpm_cheb_reg<RegSquare<8, 2>>, 64};       // 32 x 16
        if (lxl == 8 && w == 4) return {k_kpm_cheb_reg<RegSquare<8, 4>>, 128};      // 32 x 32
        if (lxl == 8 && w == 8) return {k_kpm_cheb_reg<RegSquare<8, 8>>, 256};      // 32 x 64
        if (lxl == 4 && w == 1) return {k_kpm_cheb_reg<RegSquare<4, 1>>, 32};       // 16 x 16
        if (lxl == 4 && w == 2) return {k_kpm_cheb_reg<RegSquare<4, 2>>, 64};       // 16 x 32
        if (lxl == 4 && w == 4) return {k_kpm_cheb_reg<RegSquare<4, 4>>, 128};      // 16 x 64
        return {nullptr, 0};
    }
    if (perbond) return {nullptr, 0};
    if (f->v3_lxl == 24 && f->v3_ry == 24) return {k_kpm_cheb_reg<RegHoney<8, 3, 2, 3>>, 96};
    if (f->v3_lxl == 16 && f->v3_ry == 16) return {k_kpm_cheb_reg<RegHoney<4, 4, 1, 2>>, 64};
    if (f->v3_lxl == 8 && f->v3_ry == 8) return {k_kpm_cheb_reg<RegHoney<4, 2, 1, 1>>, 32};
    return {nullptr, 0};
}

// which register kernel serves this preconditioner: 0 none (shared-memory kernels), 1 colour-uniform, 2 per-bond
static int kpm_reg_mode(const sq_kpm *k) {
    const sq_fdm *f = k->f;
    if (const char *e = getenv("SQ_KPM_REG")) if (atoi(e) == 0) return 0;
    if (!f->sym) return 0;
    if (f->v3_ok && f->cs_coluni && pick_cheb_reg(f, false).k) return 1;
    if (pick_cheb_reg(f, true).k) return 2;
    return 0;
}
// Is a register Chebyshev kernel available for this preconditioner?  (symmetric propagator; a register-path lattice with colour-uniform
// tau-independent hoppings, or a square lattice / chain with arbitrary real hoppings; SQ_KPM_REG=0 forces the shared-memory kernels)
bool kpm_reg_ok(const sq_kpm *k) { return kpm_reg_mode(k) != 0; }

// z: nrhs frequency-major arrays [n][i] (rhs_stride elements apart); applies sum_q c_q T_q(B') to every scheduled frequency in place
void kpm_cheb_reg_launch(sq_kpm *k, double2 *z, const int *d_sched, int nsched, int nrhs, size_t rhs_stride, const CgState *skip) {
    sq_fdm *f = k->f;
    if (nsched <= 0) return;
    const int mode = kpm_reg_mode(k);
    const ChebRegPick pk = pick_cheb_reg(f, mode == 2);
    if (mode == 2 && !k->site_bond.p) {                          // (colour, site) -> internal bond index, built once
        std::vector<int> sb((size_t)f->C * f->N, 0);
        for (int c = 0; c < f->C; c++)
            for (int h = f->clo[c]; h < f->chi[c]; h++) { sb[(size_t)c * f->N + f->h_nt[h].x] = h; sb[(size_t)c * f->N + f->h_nt[h].y] = h; }
        k->site_bond.from_vector(sb, f->stream);
    }
    ChebRegParams P;
    P.N = (int)f->N; P.L = (int)f->L; P.nsched = nsched; P.nrhs = nrhs; P.nchain = nsched * 2 * nrhs; P.rhs_stride = rhs_stride;
    P.z = z; P.sched = d_sched; P.order = k->d_order.p; P.coef_off = k->d_coef_off.p; P.coefs = k->d_coefs.p;
    P.csbar = k->csbar.p; P.Dbar = k->Dbar.p; P.site_bond = k->site_bond.p;
    for (int c = 0; c < 4; c++) P.clo[c] = c < f->C ? f->clo[c] : 0;
    P.avg = 0.5 * (k->bounds[1] + k->bounds[0]);
    P.imag_ = 2.0 / (k->bounds[1] - k->bounds[0]);
    P.skip = skip;
    // one CTA per SM while the chains are few (the long ones get an SM to themselves: latency is what counts); a batch of right-hand
    // sides has ~10x more chains than SMs, there the FP64 pipes are the limit and several chains share an SM
    int grid = std::min(P.nchain, f->num_sms);
    if (P.nchain > 2 * f->num_sms) {
        int occ = 1;
        SQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)pk.k, pk.threads, 0));
        if (const char *e = getenv("SQ_KPM_REG_OCC")) occ = std::max(1, atoi(e));
        grid = std::min(P.nchain, f->num_sms * std::max(1, std::min(occ, 4)));
    }
    SQ_CUDA(sq_launch(pk.k, dim3(grid), dim3(pk.threads), 0, f->stream, P));
    SQ_LAUNCH_CHECK();
    f->launches++;
    f->stats[SQ_STAT_KPM_REG]++;
}
