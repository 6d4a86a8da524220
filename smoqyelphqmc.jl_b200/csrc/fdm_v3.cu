// fdm_v3.cu -- K1 register path: fused M / M^T / M^T M for rectangular lattices.  One warp holds one part (real or
// imaginary) of one time slice in registers; no shared-memory round trips and no barriers inside the checkerboard sweeps.
//
// ncu on the shared-memory kernels (fdm.cu, fdm_v2.cu; profiles/r1_ncu_k_fdm_fused_v2_uniform_cfg4.txt) showed the fused
// matvec bound by the shared-memory crossbar and by one CTA-wide barrier per colour step: every colour step moves every
// site through shared memory (2C-1 = 7 round trips per propagator, ~450 B of crossbar traffic per site for 72 B of HBM
// traffic).  On a rectangular Lx x Ly lattice with the (x-even, x-odd, y-even, y-odd) checkerboard that traffic is
// unnecessary.  The hopping coefficients are real, so B never mixes the real and the imaginary part of a vector: the two
// parts are independent problems (blockIdx.y).  A warp holds its part of one slice for the whole kernel:
//
//     lane = xl + LXL * yq                     xl : which group of 4 consecutive x      (LXL = Lx / 4 groups)
//                                              yq : which block of RY consecutive rows  (32 / LXL blocks)
//     v[r][j]  (RY x 4 doubles per lane)       site x = 4 xl + j,  y = RY yq + r
//
// so that x-even and y-even bonds are lane-local, x-odd bonds need one shuffle per 2 sites and y-odd bonds shuffle only
// the two boundary rows of the lane's row block.  A propagator B = Gamma D Gamma^T is 8 colour steps + the diagonal, all
// in registers; what remains is FP64 issue (2 instructions per site and part per colour step; measured on B200:
// 64 FP64 lanes/clk/SM, latency 8.6 clk, tools/ubench/fp64_rate.cu) plus ~100 warp shuffles per slice and part.
//
// CTA = S+1 warps (x) one part.  Warp k loads v[l0-1+k], applies B_{l0+k}, forms w[l0+k] = v[l0+k] -/+ (.) (|w|^2 partial
// = p.Ap), publishes w to shared memory (the only shared-memory traffic: one 8-byte store and one load per site and
// part), applies B_{l0+k} again (B is symmetric for the symmetric propagator) and writes
// out[l0+k-1] = w[l0+k-1] -/+ B_{l0+k} w[l0+k].  One HBM pass, as in the other fused kernels, and the same
// floating-point expression per site (fma(s, other, c * self)), so the result is bit-identical to fdm.cu / fdm_v2.cu.
// ~110 registers per thread: several CTAs per SM, so one warp's load latency hides behind another's arithmetic.
//
// Requirements (checked by fdm_v3_detect / fdm_v3_supported, otherwise the shared-memory kernels run): symmetric
// propagator, natural site order i = x + Lx y, colour c is bond class c (x-even, x-odd, y-even, y-odd) -- or the honeycomb
// lattice of the V3Honey engine below, or a chain (V3ChainPB).  Two engine families: the UNIFORM engines (V3Lane, V3Honey) need the
// (cosh, sinh) of all bonds of one colour equal and tau-independent (any Holstein-type model with uniform hopping) and keep one
// tanh per colour; the PER-BOND engines (V3LanePB: 16 x 16, V3ChainPB: 64 / 128 / 256 sites) keep the (cosh, sinh) of every bond
// of the warp's slice in registers (SSH couplings, disordered hoppings): fdm_v3_perbond() picks the family per operator update.
#include "sq_internal.h"

#include <algorithm>
#include <cstring>
#include <map>

void fdm_v3_prepare_native(sq_fdm *f);

struct V3Params {
    int L, lb, le, S, C;
    int pre;                // 2: stage the late operands of a task in shared memory with bulk async copies (NAT = 2 kernels)
    int nphase;             // 2 for M^T M, 1 otherwise (a kernel parameter so that the phase loop stays a loop: one copy of B in the code)
    int cls[4];             // bond class of colour c: 0 x-even, 1 x-odd, 2 y-even, 3 y-odd
    int clo[4];
    const double2 *cs;      // (cosh, sinh) of colour c = cs[clo[c]]
    const double *expV;     // [l][i]
    const double *expVn;    // native order (NAT kernels)
    const double2 *ctn;     // (cosh, tanh) per colour, prepared with expVn (the in-kernel divisions were 18 % of the stall samples)
    long long *dbg;         // optional clock stamps of warp 1 of CTA 0 (profiling aid, NULL in production)
    const int *gpart;       // graph engine: partner of value a of lane `lane` in colour c = gpart[(c * 2 + a) * 32 + lane]: lane | value << 5
    const double2 *csn;     // per-bond engines: (cosh, sinh) of slot q of lane `lane` of slice l = csn[(l * NCS + q) * 32 + lane]
    size_t bstride;         // batch of vectors (blockIdx.z): elements between consecutive vectors
    int bpart;              // ... doubles between their p.Ap partials
    // CG fusion, same meaning as K2Params (fdm_v2.cu)
    const double2 *cg_d;
    double2 *cg_pnew;
    const CgState *cg_cur;
    CgState *cg_nxt;
    const double *cg_rr_part, *cg_beta_part;
    int cg_nrr, cg_nbeta, cg_beta_complex, cg_iter, cg_check;
};

// SC = 0: the reference expression a' = c a + s b (bit-identical to fdm.cu / fdm_v2.cu).
// SC = 1: a' = a + t b with t = s / c; the factor c is the same for every bond of the colour and every site is touched by
// every colour (perfect matchings), so the product of all c^2 is a global scalar folded into the native copy of
// exp(-dtau V) -- one dependent-free DFMA per site and colour step instead of DMUL + DFMA (results differ at 1e-16).
template <int SC>
__device__ __forceinline__ void rot1(double &a, double &b, double c, double s) {
    if (SC) {
        const double na = fma(s, b, a), nb = fma(s, a, b);
        a = na;
        b = nb;
    } else {
        const double na = fma(s, b, c * a), nb = fma(s, a, c * b);
        a = na;
        b = nb;
    }
}
template <int SC>
__device__ __forceinline__ double rot_half(double self, double other, double c, double s) {
    return SC ? fma(s, other, self) : fma(s, other, c * self);
}

template <int LXL, int RY>
struct V3Lane {
    static constexpr int YH = 32 / LXL, LX = 4 * LXL, LY = RY * YH, N = LX * LY;
    // engine interface of the kernels: NV values per lane, stored as NP pairs; pair u of lane `lane` is the double2
    // u * 32 + lane of a slice-part in native order, and the two consecutive sites pair_site(u), pair_site(u) + 1 in library order
    static constexpr int NV = 4 * RY, NP = 2 * RY, NCOL = 4;
    static constexpr int REGS_LIGHT = RY <= 8;          // two CTAs per SM fit
    int xl, part, yh;
    int lane_r, lane_l, lane_u, lane_d;
    int site0;              // site of v[0][0]; v[r][j] is site0 + LX r + j
    double cc[4], ss[4];    // per colour: cosh, sinh (scaled rotations: ss = tanh)

    template <int SC>
    __device__ __forceinline__ void init(const V3Params &P, int part_) {
        const int lane = threadIdx.x & 31;
        xl = lane % LXL;
        part = part_;
        yh = lane / LXL;
        lane_r = lane - xl + (xl + 1) % LXL;
        lane_l = lane - xl + (xl + LXL - 1) % LXL;
        lane_u = xl + LXL * ((yh + 1) % YH);
        lane_d = xl + LXL * ((yh + YH - 1) % YH);
        site0 = LX * RY * yh + 4 * xl;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const double2 q = SC ? __ldg(P.ctn + c) : __ldg(P.cs + P.clo[c]);
            cc[c] = q.x; ss[c] = q.y;
        }
    }

    __device__ __forceinline__ void load_cs(const V3Params &, int) {}      // uniform coefficients: nothing per slice

    template <int CL, int SC>
    __device__ __forceinline__ void step(double (&v)[RY][4]) const {
        const double c = cc[CL], s = ss[CL];
        if (CL == 0) {                                            // x-even: (0,1), (2,3) inside the lane
#pragma unroll
            for (int r = 0; r < RY; r++) { rot1<SC>(v[r][0], v[r][1], c, s); rot1<SC>(v[r][2], v[r][3], c, s); }
        } else if (CL == 1) {                                     // x-odd: (1,2) inside, (3 | 0 of the right neighbour)
#pragma unroll
            for (int r = 0; r < RY; r++) {
                const double fromR = __shfl_sync(0xffffffffu, v[r][0], lane_r);
                const double fromL = __shfl_sync(0xffffffffu, v[r][3], lane_l);
                rot1<SC>(v[r][1], v[r][2], c, s);
                v[r][3] = rot_half<SC>(v[r][3], fromR, c, s);
                v[r][0] = rot_half<SC>(v[r][0], fromL, c, s);
            }
        } else if (CL == 2) {                                     // y-even: rows (2k, 2k+1)
#pragma unroll
            for (int r = 0; r < RY; r += 2)
#pragma unroll
                for (int j = 0; j < 4; j++) rot1<SC>(v[r][j], v[r + 1][j], c, s);
        } else {                                                  // y-odd: rows (2k+1, 2k+2), block boundary by shuffle
            if (YH > 1) {
                double up[4], dn[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    up[j] = __shfl_sync(0xffffffffu, v[0][j], lane_u);
                    dn[j] = __shfl_sync(0xffffffffu, v[RY - 1][j], lane_d);
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    v[RY - 1][j] = rot_half<SC>(v[RY - 1][j], up[j], c, s);
                    v[0][j] = rot_half<SC>(v[0][j], dn[j], c, s);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) rot1<SC>(v[RY - 1][j], v[0][j], c, s);
            }
#pragma unroll
            for (int r = 1; r + 1 < RY; r += 2)
#pragma unroll
                for (int j = 0; j < 4; j++) rot1<SC>(v[r][j], v[r + 1][j], c, s);
        }
    }

    // v <- B_l v,  B = Gamma D Gamma^T: colours 3, 2, 1, 0, D, 0, 1, 2, 3 (colour c == bond class c, checked at create)
    __device__ __forceinline__ int pair_site(int u) const { return site0 + LX * (u >> 1) + 2 * (u & 1); }
    // flat-array overload: ev = slice base of exp(-dtau V) (+ 2 lane in native order)
    template <int NAT, int SM>
    __device__ __forceinline__ void apply_B_ev(double (&v)[NV], const double *ev) const {
        apply_B_ev<NAT, SM>(reinterpret_cast<double (&)[RY][4]>(v), NAT ? ev : ev + site0);
    }
    template <int NAT>
    __device__ __forceinline__ void apply_B(double (&v)[RY][4], int l, const V3Params &P) const {
        // NAT: exp(-dtau V) x prod_c cosh_c^2 in the native order [l][r][j/2][lane][j%2] (one coalesced 16-byte load per two
        // sites), and the scaled rotations (rot1<1>)
        apply_B_ev<NAT, 0>(v, NAT ? P.expVn + (size_t)l * N + 2 * (threadIdx.x & 31) : P.expV + (size_t)l * N + site0);
    }
    // ev: this lane's first diagonal element of the slice; SM = 1: the slice (native order) is in shared memory
    template <int NAT, int SM>
    __device__ __forceinline__ void apply_B_ev(double (&v)[RY][4], const double *ev) const {
        step<3, NAT>(v); step<2, NAT>(v); step<1, NAT>(v); step<0, NAT>(v);
#pragma unroll
        for (int r = 0; r < RY; r++) {
            const double2 *q01 = (const double2 *)(ev + (NAT ? 128 * r : LX * r)), *q23 = (const double2 *)(ev + (NAT ? 128 * r + 64 : LX * r + 2));
            const double2 e01 = SM ? *q01 : __ldg(q01), e23 = SM ? *q23 : __ldg(q23);
            v[r][0] *= e01.x; v[r][1] *= e01.y; v[r][2] *= e23.x; v[r][3] *= e23.y;
        }
        step<0, NAT>(v); step<1, NAT>(v); step<2, NAT>(v); step<3, NAT>(v);
    }
};

// ---------------------------------------------------------------------------------------------------
// Honeycomb engine: two orbitals per cell, site = orb + 2 (c1 + L1 c2), three bond types = three colours
//     colour 0: A(c1, c2) - B(c1, c2)        colour 1: A(c1, c2) - B(c1 - 1, c2)        colour 2: A(c1, c2) - B(c1, c2 - 1)
// (the lattice of the reference's tutorials, tutorials/holstein_honeycomb.jl).  lane = g1 + G1 g2 holds an R1 x R2 block
// of cells, L1 = G1 R1, L2 = (32 / G1) R2; pair u = a2 R1 + a1 is the cell (a1, a2) of the block: v[2u] = A, v[2u + 1] = B.
// Colour 0 is lane-local; colours 1 / 2 are local except for the first column / row of the block, which pairs with the B
// orbitals of the last column / row of the neighbouring lane (one shuffle per boundary cell).
// ---------------------------------------------------------------------------------------------------
template <int G1, int R1, int R2>
struct V3Honey {
    static constexpr int G2 = 32 / G1, L1 = G1 * R1, L2 = G2 * R2;
    static constexpr int NP = R1 * R2, NV = 2 * NP, N = 2 * L1 * L2, NCOL = 3;
    static constexpr int REGS_LIGHT = NV <= 36;
    int g1, g2, part;
    int lane_r, lane_l, lane_u, lane_d;
    double cc[3], ss[3];

    template <int SC>
    __device__ __forceinline__ void init(const V3Params &P, int part_) {
        const int lane = threadIdx.x & 31;
        g1 = lane % G1;
        g2 = lane / G1;
        part = part_;
        lane_r = (g1 + 1) % G1 + G1 * g2;
        lane_l = (g1 + G1 - 1) % G1 + G1 * g2;
        lane_u = g1 + G1 * ((g2 + 1) % G2);
        lane_d = g1 + G1 * ((g2 + G2 - 1) % G2);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double2 q = SC ? __ldg(P.ctn + c) : __ldg(P.cs + P.clo[c]);
            cc[c] = q.x; ss[c] = q.y;
        }
    }
    __device__ __forceinline__ int pair_site(int u) const { return 2 * ((R1 * g1 + u % R1) + L1 * (R2 * g2 + u / R1)); }
    __device__ __forceinline__ void load_cs(const V3Params &, int) {}

    template <int CL, int SC>
    __device__ __forceinline__ void step(double (&v)[NV]) const {
        const double c = cc[CL], s = ss[CL];
        if (CL == 0) {
#pragma unroll
            for (int u = 0; u < NP; u++) rot1<SC>(v[2 * u], v[2 * u + 1], c, s);
        } else if (CL == 1) {                             // A(a1, a2) - B(a1 - 1, a2)
#pragma unroll
            for (int a2 = 0; a2 < R2; a2++) {
                const int u0 = a2 * R1, u1 = a2 * R1 + R1 - 1;
                if (G1 > 1) {
                    const double fromL = __shfl_sync(0xffffffffu, v[2 * u1 + 1], lane_l);     // B of the left lane's last column
                    const double fromR = __shfl_sync(0xffffffffu, v[2 * u0], lane_r);         // A of the right lane's first column
#pragma unroll
                    for (int a1 = R1 - 1; a1 >= 1; a1--) rot1<SC>(v[2 * (u0 + a1)], v[2 * (u0 + a1 - 1) + 1], c, s);
                    // the two boundary values: A(0, a2) with fromL; B(R1 - 1, a2) with fromR -- after the inner bonds when R1 > 1
                    // touches neither (A(0) and B(R1-1) are not part of an inner bond)
                    v[2 * u0] = rot_half<SC>(v[2 * u0], fromL, c, s);
                    v[2 * u1 + 1] = rot_half<SC>(v[2 * u1 + 1], fromR, c, s);
                } else {
#pragma unroll
                    for (int a1 = 0; a1 < R1; a1++) rot1<SC>(v[2 * (u0 + a1)], v[2 * (u0 + (a1 + R1 - 1) % R1) + 1], c, s);
                }
            }
        } else {                                          // A(a1, a2) - B(a1, a2 - 1)
#pragma unroll
            for (int a1 = 0; a1 < R1; a1++) {
                const int u0 = a1, u1 = (R2 - 1) * R1 + a1;
                if (G2 > 1) {
                    const double fromD = __shfl_sync(0xffffffffu, v[2 * u1 + 1], lane_d);
                    const double fromU = __shfl_sync(0xffffffffu, v[2 * u0], lane_u);
#pragma unroll
                    for (int a2 = R2 - 1; a2 >= 1; a2--) rot1<SC>(v[2 * (a2 * R1 + a1)], v[2 * ((a2 - 1) * R1 + a1) + 1], c, s);
                    v[2 * u0] = rot_half<SC>(v[2 * u0], fromD, c, s);
                    v[2 * u1 + 1] = rot_half<SC>(v[2 * u1 + 1], fromU, c, s);
                } else {
#pragma unroll
                    for (int a2 = 0; a2 < R2; a2++) rot1<SC>(v[2 * (a2 * R1 + a1)], v[2 * (((a2 + R2 - 1) % R2) * R1 + a1) + 1], c, s);
                }
            }
        }
    }

    // v <- B v,  B = Gamma D Gamma^T: colours 2, 1, 0, D, 0, 1, 2.  ev: slice base of exp(-dtau V) (+ 2 lane in native order)
    template <int NAT, int SM>
    __device__ __forceinline__ void apply_B_ev(double (&v)[NV], const double *ev) const {
        step<2, NAT>(v); step<1, NAT>(v); step<0, NAT>(v);
#pragma unroll
        for (int u = 0; u < NP; u++) {
            const double2 *q = (const double2 *)(ev + (NAT ? 64 * u : pair_site(u)));
            const double2 e = SM ? *q : __ldg(q);
            v[2 * u] *= e.x; v[2 * u + 1] *= e.y;
        }
        step<0, NAT>(v); step<1, NAT>(v); step<2, NAT>(v);
    }
};

// ---------------------------------------------------------------------------------------------------
// Per-bond engines: every bond of every slice has its own (cosh, sinh) -- SSH couplings, disordered hoppings.  A warp applies
// the propagator of ONE slice for the whole kernel (both passes of M^T M, every iteration of the resident solver), so the
// coefficients of that slice live in REGISTERS: slot q of a lane is one bond the lane takes part in; bonds that cross to a
// neighbouring lane are held by both lanes (no coefficient shuffles).  The slots are filled from a slot-ordered copy of the
// coefficient table (csn, built with the native-order copy of exp(-dtau V)): one coalesced 16-byte load per slot.  The update
// is the reference expression a' = c a + s b (rot1<0>): bit-identical to fdm.cu / fdm_v2.cu in every order.
// The slot enumeration below is mirrored on the host by v3pb_slot_bonds().
// ---------------------------------------------------------------------------------------------------
template <int LXL, int RY>
struct V3LanePB {
    static constexpr int YH = 32 / LXL, LX = 4 * LXL, LY = RY * YH, N = LX * LY;
    static constexpr int NV = 4 * RY, NP = 2 * RY, NCOL = 4;
    static constexpr int REGS_LIGHT = 0;
    // slots: x-even 2 RY | x-odd 3 RY (inner, to the right lane, from the left lane) | y-even 2 RY | y-odd 4 up + 4 down + inner
    static constexpr int Q0 = 0, Q1 = 2 * RY, Q2 = 5 * RY, Q3 = 7 * RY, NCS = 9 * RY + 4;
    static_assert(YH > 1, "row blocks of at least two lanes");
    int xl, part, yh;
    int lane_r, lane_l, lane_u, lane_d;
    int site0;
    double2 cf[NCS];

    template <int SC>
    __device__ __forceinline__ void init(const V3Params &, int part_) {
        const int lane = threadIdx.x & 31;
        xl = lane % LXL;
        part = part_;
        yh = lane / LXL;
        lane_r = lane - xl + (xl + 1) % LXL;
        lane_l = lane - xl + (xl + LXL - 1) % LXL;
        lane_u = xl + LXL * ((yh + 1) % YH);
        lane_d = xl + LXL * ((yh + YH - 1) % YH);
        site0 = LX * RY * yh + 4 * xl;
    }
    __device__ __forceinline__ void load_cs(const V3Params &P, int l) {
        const double2 *g = P.csn + (size_t)l * NCS * 32 + (threadIdx.x & 31);
#pragma unroll
        for (int q = 0; q < NCS; q++) cf[q] = __ldg(g + 32 * q);
    }
    __device__ __forceinline__ int pair_site(int u) const { return site0 + LX * (u >> 1) + 2 * (u & 1); }

    template <int CL>
    __device__ __forceinline__ void step(double (&v)[RY][4]) const {
        if (CL == 0) {
#pragma unroll
            for (int r = 0; r < RY; r++) {
                rot1<0>(v[r][0], v[r][1], cf[Q0 + 2 * r].x, cf[Q0 + 2 * r].y);
                rot1<0>(v[r][2], v[r][3], cf[Q0 + 2 * r + 1].x, cf[Q0 + 2 * r + 1].y);
            }
        } else if (CL == 1) {
#pragma unroll
            for (int r = 0; r < RY; r++) {
                const double fromR = __shfl_sync(0xffffffffu, v[r][0], lane_r);
                const double fromL = __shfl_sync(0xffffffffu, v[r][3], lane_l);
                rot1<0>(v[r][1], v[r][2], cf[Q1 + 3 * r].x, cf[Q1 + 3 * r].y);
                v[r][3] = rot_half<0>(v[r][3], fromR, cf[Q1 + 3 * r + 1].x, cf[Q1 + 3 * r + 1].y);
                v[r][0] = rot_half<0>(v[r][0], fromL, cf[Q1 + 3 * r + 2].x, cf[Q1 + 3 * r + 2].y);
            }
        } else if (CL == 2) {
#pragma unroll
            for (int r = 0; r < RY; r += 2)
#pragma unroll
                for (int j = 0; j < 4; j++) rot1<0>(v[r][j], v[r + 1][j], cf[Q2 + 2 * r + j].x, cf[Q2 + 2 * r + j].y);
        } else {
            double up[4], dn[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                up[j] = __shfl_sync(0xffffffffu, v[0][j], lane_u);
                dn[j] = __shfl_sync(0xffffffffu, v[RY - 1][j], lane_d);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                v[RY - 1][j] = rot_half<0>(v[RY - 1][j], up[j], cf[Q3 + j].x, cf[Q3 + j].y);
                v[0][j] = rot_half<0>(v[0][j], dn[j], cf[Q3 + 4 + j].x, cf[Q3 + 4 + j].y);
            }
#pragma unroll
            for (int r = 1; r + 1 < RY; r += 2)
#pragma unroll
                for (int j = 0; j < 4; j++) rot1<0>(v[r][j], v[r + 1][j], cf[Q3 + 8 + 2 * (r - 1) + j].x, cf[Q3 + 8 + 2 * (r - 1) + j].y);
        }
    }
    template <int NAT, int SM>
    __device__ __forceinline__ void apply_B_ev(double (&w)[NV], const double *ev0) const {
        double (&v)[RY][4] = reinterpret_cast<double (&)[RY][4]>(w);
        const double *ev = NAT ? ev0 : ev0 + site0;
        step<3>(v); step<2>(v); step<1>(v); step<0>(v);
#pragma unroll
        for (int r = 0; r < RY; r++) {
            const double2 *q01 = (const double2 *)(ev + (NAT ? 128 * r : LX * r)), *q23 = (const double2 *)(ev + (NAT ? 128 * r + 64 : LX * r + 2));
            const double2 e01 = SM ? *q01 : __ldg(q01), e23 = SM ? *q23 : __ldg(q23);
            v[r][0] *= e01.x; v[r][1] *= e01.y; v[r][2] *= e23.x; v[r][3] *= e23.y;
        }
        step<0>(v); step<1>(v); step<2>(v); step<3>(v);
    }
};

// Chain, N = 64 R sites in natural order, colour 0 = bonds (2k, 2k+1), colour 1 = bonds (2k+1, 2k+2): a lane holds 2 R consecutive
// sites; pair u = sites 2 R lane + 2u, + 1.  Slots: R even bonds | R - 1 inner odd bonds | to the right lane | from the left lane.
template <int R>
struct V3ChainPB {
    static constexpr int N = 64 * R, NV = 2 * R, NP = R, NCOL = 2, NCS = 2 * R + 1;
    static constexpr int REGS_LIGHT = 1;
    int part, lane_r, lane_l, site0;
    double2 cf[NCS];

    template <int SC>
    __device__ __forceinline__ void init(const V3Params &, int part_) {
        const int lane = threadIdx.x & 31;
        part = part_;
        lane_r = (lane + 1) & 31;
        lane_l = (lane + 31) & 31;
        site0 = 2 * R * lane;
    }
    __device__ __forceinline__ void load_cs(const V3Params &P, int l) {
        const double2 *g = P.csn + (size_t)l * NCS * 32 + (threadIdx.x & 31);
#pragma unroll
        for (int q = 0; q < NCS; q++) cf[q] = __ldg(g + 32 * q);
    }
    __device__ __forceinline__ int pair_site(int u) const { return site0 + 2 * u; }

    template <int CL>
    __device__ __forceinline__ void step(double (&v)[NV]) const {
        if (CL == 0) {
#pragma unroll
            for (int u = 0; u < R; u++) rot1<0>(v[2 * u], v[2 * u + 1], cf[u].x, cf[u].y);
        } else {
            const double fromR = __shfl_sync(0xffffffffu, v[0], lane_r);
            const double fromL = __shfl_sync(0xffffffffu, v[NV - 1], lane_l);
#pragma unroll
            for (int u = 0; u + 1 < R; u++) rot1<0>(v[2 * u + 1], v[2 * u + 2], cf[R + u].x, cf[R + u].y);
            v[NV - 1] = rot_half<0>(v[NV - 1], fromR, cf[2 * R - 1].x, cf[2 * R - 1].y);
            v[0] = rot_half<0>(v[0], fromL, cf[2 * R].x, cf[2 * R].y);
        }
    }
    template <int NAT, int SM>
    __device__ __forceinline__ void apply_B_ev(double (&v)[NV], const double *ev0) const {
        const double *ev = NAT ? ev0 : ev0 + site0;
        step<1>(v); step<0>(v);
#pragma unroll
        for (int u = 0; u < NP; u++) {
            const double2 *q = (const double2 *)(ev + (NAT ? 64 * u : 2 * u));
            const double2 e = SM ? *q : __ldg(q);
            v[2 * u] *= e.x; v[2 * u + 1] *= e.y;
        }
        step<0>(v); step<1>(v);
    }
};

// Graph engine (resident CG only): ANY lattice with N <= 64 sites and C <= 4 colours -- the 18-site honeycomb of BASELINE config 1,
// short chains, 4 x 4 ... 8 x 8 squares.  The slice is padded to 64 sites, lane = sites 2 lane, 2 lane + 1 (native order = natural site
// order, real and imaginary planes); every bond goes through shuffles: per colour a lane fetches both values of the partner lane of each
// of its two sites and picks one.  Per-bond, per-slice (cosh, sinh) in registers as in the per-bond engines (slot q = 2 c + a); a site
// without a bond in a colour -- and every padding site -- carries (1, 0) and is its own partner, so that it passes through unchanged.
struct V3Graph {
    static constexpr int N = 64, NV = 2, NP = 1, NCOL = 4, NCS = 8, REGS_LIGHT = 1;
    int part, C;
    int pl[4][2];
    double2 cf[NCS];

    template <int SC>
    __device__ __forceinline__ void init(const V3Params &P, int part_) {
        const int lane = threadIdx.x & 31;
        part = part_;
        C = P.C;
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int a = 0; a < 2; a++) pl[c][a] = __ldg(P.gpart + (c * 2 + a) * 32 + lane);
    }
    __device__ __forceinline__ void load_cs(const V3Params &P, int l) {
        const double2 *g = P.csn + (size_t)l * NCS * 32 + (threadIdx.x & 31);
#pragma unroll
        for (int q = 0; q < NCS; q++) cf[q] = __ldg(g + 32 * q);
    }
    template <int CL>
    __device__ __forceinline__ void step(double (&v)[NV]) const {
        if (CL >= C) return;                               // (warp-uniform)
        const int w0 = pl[CL][0], w1 = pl[CL][1];
        const double a0 = __shfl_sync(0xffffffffu, v[0], w0 & 31), a1 = __shfl_sync(0xffffffffu, v[1], w0 & 31);
        const double b0 = __shfl_sync(0xffffffffu, v[0], w1 & 31), b1 = __shfl_sync(0xffffffffu, v[1], w1 & 31);
        v[0] = rot_half<0>(v[0], (w0 & 32) ? a1 : a0, cf[2 * CL].x, cf[2 * CL].y);
        v[1] = rot_half<0>(v[1], (w1 & 32) ? b1 : b0, cf[2 * CL + 1].x, cf[2 * CL + 1].y);
    }
    // v <- B v, B = Gamma D Gamma^T: colours C-1 ... 0, D, 0 ... C-1.  ev: the slice's diagonal factors (+ 2 lane)
    template <int NAT, int SM>
    __device__ __forceinline__ void apply_B_ev(double (&v)[NV], const double *ev) const {
        step<3>(v); step<2>(v); step<1>(v); step<0>(v);
        const double2 e = *reinterpret_cast<const double2 *>(ev);
        v[0] *= e.x; v[1] *= e.y;
        step<0>(v); step<1>(v); step<2>(v); step<3>(v);
    }
};

// NAT = 1: all vectors (in, out, cg_d, cg_pnew) are in the NATIVE order of this kernel -- slice l, part q, then
// [r][j/2][lane][j%2] doubles -- so that every load / store is one fully coalesced 16-byte access per lane.  The CG solver
// keeps its vectors in this order for the whole solve (cg.cu); the order is converted once on entry and once on exit.
template <int MODE, int FUSE, int NAT, class G>
__global__ void __launch_bounds__(256, G::REGS_LIGHT ? 2 : 1)
k_fdm_v3(const __grid_constant__ V3Params P, double2 *__restrict__ out, const double2 *__restrict__ in,
         double *__restrict__ pAp_part, const CgState *__restrict__ skip) {
    // NAT = 2: native order with the two late operands of a task -- the slice combined after B and the diagonal factors used inside
    // B -- staged in shared memory by bulk async copies (TMA, one 8 KB copy each) issued before the first load, so that a task waits
    // for memory once instead of three times
    constexpr int N = G::N, NP = G::NP;
    constexpr bool TMA = (NAT == 2);
    constexpr int NATB = NAT ? 1 : 0;
    extern __shared__ __align__(128) double wsm[];      // [S][NP][32] double2 (this CTA's part); TMA: [S + 1] of those + [S + 1][N] diagonal factors
    __shared__ __align__(8) unsigned long long mbar[8];
    __shared__ double red[32];
    sq_pdl_prologue();
    if (!FUSE && gridDim.z > 1) {                       // multi-RHS batch: one vector, one set of partials and one solver state per blockIdx.z
        in += (size_t)blockIdx.z * P.bstride;
        out += (size_t)blockIdx.z * P.bstride;
        if (pAp_part) pAp_part += (size_t)blockIdx.z * P.bpart;
        if (skip) skip += blockIdx.z;
    }
    if (skip && skip->done) {
        if (FUSE && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *P.cg_nxt = *skip;
        return;
    }
#define V3_STAMP(q) do { if (P.dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 32) P.dbg[q] = clock64(); } while (0)
    V3_STAMP(0);
    G E;
    E.template init<NATB>(P, blockIdx.y);
    const int L = P.L;
    const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;
    const int nper = (MODE == 2) ? P.S : P.S + 1;       // output slices per CTA
    const int l0 = P.lb + blockIdx.x * nper;
    const int ns = min(nper, P.le - l0);
    double2 beta = make_double2(0.0, 0.0);
    if (FUSE) {
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
                if (blockIdx.x == 0 && blockIdx.y == 0) {
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
    // slices of this warp: propagate in[lin] with B_{lB}, combine with in[lself]
    bool active;
    int lin, lself, lB;
    if (MODE == 2) { active = k <= ns; lin = l0 - 1 + k; lself = l0 + k; lB = l0 + k; }
    else if (MODE == 0) { active = k < ns; lin = l0 + k - 1; lself = l0 + k; lB = l0 + k; }
    else { active = k < ns; lin = l0 + k + 1; lself = l0 + k; lB = l0 + k + 1; }
    lin = lin < 0 ? lin + L : (lin >= L ? lin - L : lin);
    lB = lB >= L ? lB - L : lB;
    lself = lself >= L ? lself - L : lself;
    const double sg = (lB == 0) ? 1.0 : -1.0;
    const int part = E.part;
    if (active) E.load_cs(P, lB);                       // per-bond engines: this warp's slice of (cosh, sinh) into registers
    const bool publish = (MODE == 2) && (k < ns);
    double *evsm = wsm + (size_t)(P.S + 1) * N + (size_t)k * N;
    double2 *selfsm = reinterpret_cast<double2 *>(wsm) + (size_t)k * NP * 32;
    const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar[k]);
    if (TMA && active) {
        if (lane == 0) {
            const unsigned bytes = (unsigned)(N * sizeof(double));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(2 * bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"((unsigned)__cvta_generic_to_shared(selfsm)),
                           "l"(reinterpret_cast<const double *>(in) + ((size_t)lself * 2 + E.part) * N), "r"(bytes), "r"(mb) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"((unsigned)__cvta_generic_to_shared(evsm)), "l"(P.expVn + (size_t)lB * N), "r"(bytes), "r"(mb) : "memory");
        }
        __syncwarp();
    }
    double v[G::NV];
    double acc = 0.0;
    // two elements (r, 2 jp), (r, 2 jp + 1) of slice l, this CTA's part; with CG fusion the vector is p = d + beta * in
    // formed on the fly (beta is real in the native-order solver: unpreconditioned CG)
    auto off = [&](int l, int u) -> size_t {             // offset in doubles of the first of the two elements of pair u
        if (NAT) return ((size_t)l * 2 + part) * N + (u * 32 + lane) * 2;
        return 2 * ((size_t)l * N + E.pair_site(u)) + part;
    };
    auto ld2 = [&](const double2 *src, size_t o) -> double2 {
        const double *q = reinterpret_cast<const double *>(src) + o;
        if (NAT) return *reinterpret_cast<const double2 *>(q);
        return make_double2(q[0], q[2]);
    };
    auto st2 = [&](double2 *dst, size_t o, double a, double b) {
        double *q = reinterpret_cast<double *>(dst) + o;
        if (NAT) *reinterpret_cast<double2 *>(q) = make_double2(a, b);
        else { q[0] = a; q[2] = b; }
    };
    auto load2 = [&](int l, int u) -> double2 {
        const size_t o = off(l, u);
        if (!FUSE) return ld2(in, o);
        if (NAT) {
            const double2 a = ld2(in, o), d = ld2(P.cg_d, o);
            return make_double2(__dadd_rn(d.x, __dmul_rn(beta.x, a.x)), __dadd_rn(d.y, __dmul_rn(beta.x, a.y)));
        }
        const size_t g = (size_t)l * N + E.pair_site(u);
        const double2 p0 = cadd(P.cg_d[g], cmul(beta, in[g])), p1 = cadd(P.cg_d[g + 1], cmul(beta, in[g + 1]));
        return part ? make_double2(p0.y, p1.y) : make_double2(p0.x, p1.x);
    };
    const int nphase = (MODE == 2) ? P.nphase : 1;
#pragma unroll 1
    for (int ph = 0; ph < nphase; ph++) {
        int phase = ph;
        asm volatile("" : "+r"(phase));                  // opaque: keeps the compiler from peeling the loop into two copies of B
        const bool work = active && (phase == 0 || k >= 1);
        if (work) {
            if (phase == 0) {
#pragma unroll
                for (int u = 0; u < NP; u++) {
                    const double2 q = load2(lin, u);
                    v[2 * u] = q.x; v[2 * u + 1] = q.y;
                }
                if (P.dbg) { double t = 0; for (int u = 0; u < NP; u++) t += v[2 * u]; if (t == 1.2345e300) P.dbg[15] = 1; }   // wait for the loads
                V3_STAMP(1);
            }
            if (TMA && phase == 0) {                     // both staged operands have landed
                unsigned ok;
                do {
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                                 : "=r"(ok) : "r"(mb) : "memory");
                } while (!ok);
            }
            E.template apply_B_ev<NATB, TMA ? 1 : 0>(v, TMA ? evsm + 2 * lane : (NAT ? P.expVn + (size_t)lB * N + 2 * lane : P.expV + (size_t)lB * N));
            if (P.dbg) { double t = 0; for (int u = 0; u < NP; u++) t += v[2 * u]; if (t == 1.2345e300) P.dbg[15] = 1; }
            V3_STAMP(2 + 3 * phase);
        }
        if (phase == 0) {
            if (work) {
#pragma unroll
                for (int u = 0; u < NP; u++) {
                    const double2 self = TMA ? selfsm[u * 32 + lane] : load2(lself, u);
                    const double w0 = fma(sg, v[2 * u], self.x), w1 = fma(sg, v[2 * u + 1], self.y);
                    if (MODE == 2) {
                        v[2 * u] = w0; v[2 * u + 1] = w1;
                        if (publish) {
                            acc += w0 * w0;
                            acc += w1 * w1;
                            reinterpret_cast<double2 *>(wsm)[(k * NP + u) * 32 + lane] = make_double2(w0, w1);
                            if (FUSE) st2(P.cg_pnew, off(lself, u), self.x, self.y);
                        }
                    } else {
                        st2(out, off(lself, u), w0, w1);
                    }
                }
            }
            V3_STAMP(3);
        } else {
            __syncthreads();
            V3_STAMP(6);
            if (work) {
                const int lo = l0 + k - 1;                  // < le <= L: no wrap
#pragma unroll
                for (int u = 0; u < NP; u++) {
                    const double2 w = reinterpret_cast<const double2 *>(wsm)[((k - 1) * NP + u) * 32 + lane];
                    st2(out, off(lo, u), fma(sg, v[2 * u], w.x), fma(sg, v[2 * u + 1], w.y));
                }
            }
        }
    }
    V3_STAMP(7);
    if (MODE == 2 && pAp_part) {
        double a[1] = {acc};
        block_sum<1>(a, red);
        if (threadIdx.x == 0) pAp_part[blockIdx.x + gridDim.x * blockIdx.y] = a[0];
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef void (*v3_kernel_t)(const V3Params, double2 *, const double2 *, double *, const CgState *);

template <int LXL, int RY>
static v3_kernel_t pick_mode(int mode) {       // mode 3: M^T M with the CG p update fused into the load; +4: native order
    switch (mode) {
        case 0: return k_fdm_v3<0, 0, 0, V3Lane<LXL, RY>>;
        case 1: return k_fdm_v3<1, 0, 0, V3Lane<LXL, RY>>;
        case 2: return k_fdm_v3<2, 0, 0, V3Lane<LXL, RY>>;
        case 3: return k_fdm_v3<2, 1, 0, V3Lane<LXL, RY>>;
        case 6: return k_fdm_v3<2, 0, 1, V3Lane<LXL, RY>>;
        case 7: return k_fdm_v3<2, 1, 1, V3Lane<LXL, RY>>;
        case 10: return k_fdm_v3<2, 0, 2, V3Lane<LXL, RY>>;
    }
    return nullptr;
}
template <int G1, int R1, int R2>
static v3_kernel_t pick_mode_h(int mode) {
    typedef V3Honey<G1, R1, R2> H;
    switch (mode) {
        case 0: return k_fdm_v3<0, 0, 0, H>;
        case 1: return k_fdm_v3<1, 0, 0, H>;
        case 2: return k_fdm_v3<2, 0, 0, H>;
        case 3: return k_fdm_v3<2, 1, 0, H>;
        case 6: return k_fdm_v3<2, 0, 1, H>;
        case 7: return k_fdm_v3<2, 1, 1, H>;
        case 10: return k_fdm_v3<2, 0, 2, H>;
    }
    return nullptr;
}
template <class G>
static v3_kernel_t pick_mode_pb(int mode) {    // per-bond engines (no bulk-copy staged variant: these lattices are L2-resident)
    switch (mode) {
        case 0: return k_fdm_v3<0, 0, 0, G>;
        case 1: return k_fdm_v3<1, 0, 0, G>;
        case 2: return k_fdm_v3<2, 0, 0, G>;
        case 3: return k_fdm_v3<2, 1, 0, G>;
        case 6: return k_fdm_v3<2, 0, 1, G>;
        case 7: return k_fdm_v3<2, 1, 1, G>;
    }
    return nullptr;
}
// per-bond geometries: kind 0 square (lxl, ry), kind 2 chain (lxl = site pairs per lane)
static v3_kernel_t pick3pb(int kind, int a, int b, int mode) {
    if (kind == 0 && a == 4 && b == 2) return pick_mode_pb<V3LanePB<4, 2>>(mode);      // 16 x 16
    if (kind == 2 && a == 1) return pick_mode_pb<V3ChainPB<1>>(mode);                  // N = 64
    if (kind == 2 && a == 2) return pick_mode_pb<V3ChainPB<2>>(mode);                  // N = 128
    if (kind == 2 && a == 4) return pick_mode_pb<V3ChainPB<4>>(mode);                  // N = 256
    return nullptr;
}
// honeycomb geometries: (lxl, ry) hold (L1, L2) with kind = 1
static v3_kernel_t pick3h(int L1, int L2, int mode) {
    if (L1 == 24 && L2 == 24) return pick_mode_h<8, 3, 6>(mode);
    if (L1 == 16 && L2 == 16) return pick_mode_h<4, 4, 2>(mode);
    if (L1 == 8 && L2 == 8) return pick_mode_h<4, 2, 1>(mode);
    return nullptr;
}
static v3_kernel_t pick3(int lxl, int ry, int mode) {
    if (lxl == 8 && ry == 4) return pick_mode<8, 4>(mode);        // 32 x 16
    if (lxl == 8 && ry == 8) return pick_mode<8, 8>(mode);        // 32 x 32
    if (lxl == 8 && ry == 16) return pick_mode<8, 16>(mode);      // 32 x 64
    if (lxl == 4 && ry == 2) return pick_mode<4, 2>(mode);        // 16 x 16
    if (lxl == 4 && ry == 4) return pick_mode<4, 4>(mode);        // 16 x 32
    if (lxl == 4 && ry == 8) return pick_mode<4, 8>(mode);        // 16 x 64
    return nullptr;
}

// Is the lattice an Lx x Ly periodic rectangle in natural site order whose 4 colours are the 4 bond classes?
// honeycomb: site = orb + 2 (c1 + L1 c2), colour c = bond type c: A(c1, c2) - B(c1 + d1, c2 + d2), d = (0,0), (-1,0), (0,-1)
static bool fdm_v3_detect_honeycomb(sq_fdm *f) {
    if (!f->sym || f->C != 3 || 2 * f->Nh != 3 * f->N) return false;
    for (int L : {24, 16, 8}) {
        if (f->N != 2 * L * L || !pick3h(L, L, 2)) continue;
        static const int d1[3] = {0, -1, 0}, d2[3] = {0, 0, -1};
        bool ok = true;
        for (int c = 0; c < 3 && ok; c++) {
            if (f->chi[c] - f->clo[c] != f->N / 2) { ok = false; break; }
            for (int h = f->clo[c]; h < f->chi[c] && ok; h++) {
                int a = f->h_nt[h].x, b = f->h_nt[h].y;
                if (a & 1) std::swap(a, b);                   // a: A orbital, b: B orbital
                if ((a & 1) || !(b & 1)) { ok = false; break; }
                const int ca = a / 2, cb = b / 2, a1 = ca % L, a2 = ca / L, b1 = cb % L, b2 = cb / L;
                ok = (b1 == (a1 + d1[c] + L) % L) && (b2 == (a2 + d2[c] + L) % L);
            }
        }
        if (!ok) continue;
        f->v3_ok = 1; f->v3_kind = 1; f->v3_lxl = L; f->v3_ry = L;
        for (int c = 0; c < 4; c++) f->v3_cls[c] = c;
        for (int mode : {0, 1, 2, 3, 6, 7, 10})
            SQ_CUDA(cudaFuncSetAttribute(pick3h(L, L, mode), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem_optin));
        return true;
    }
    return false;
}

// chain of N = 64 R sites in natural order: colour 0 = bonds (2k, 2k+1), colour 1 = bonds (2k+1, 2k+2 mod N)
static bool fdm_v3_detect_chain(sq_fdm *f) {
    if (!f->sym || f->C != 2 || f->Nh != f->N || f->N % 64) return false;
    const int N = (int)f->N, R = N / 64;
    if (!pick3pb(2, R, 0, 2)) return false;
    for (int c = 0; c < 2; c++) {
        if (f->chi[c] - f->clo[c] != N / 2) return false;
        for (int h = f->clo[c]; h < f->chi[c]; h++) {
            int a = f->h_nt[h].x, b = f->h_nt[h].y;
            if ((a + 1) % N != b) std::swap(a, b);
            if ((a + 1) % N != b || (a & 1) != c) return false;
        }
    }
    f->v3_kind = 2; f->v3_lxl = R; f->v3_ry = 0; f->v3_pb_ok = 1;
    for (int c = 0; c < 4; c++) f->v3_cls[c] = c;
    for (int mode : {0, 1, 2, 3, 6, 7})
        SQ_CUDA(cudaFuncSetAttribute(pick3pb(2, R, 0, mode), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem_optin));
    return true;
}

void fdm_v3_detect(sq_fdm *f) {
    f->v3_ok = 0;
    f->v3_pb_ok = 0;
    f->v3_kind = 0;
    f->v3g_ok = (f->sym && f->N <= 64 && f->C >= 1 && f->C <= 4 && f->Nh >= 1) ? 1 : 0;      // graph engine of the resident CG
    if (fdm_v3_detect_honeycomb(f)) return;
    if (fdm_v3_detect_chain(f)) return;
    if (!f->sym || f->C != 4 || f->Nh != 2 * f->N) return;
    for (int LX : {32, 16}) {
        if (f->N % LX) continue;
        const int LY = (int)(f->N / LX), LXL = LX / 4, YH = 32 / LXL;
        if (LY % YH) continue;
        const int RY = LY / YH;
        if (RY % 2 || !pick3(LXL, RY, 2)) continue;
        int cls[4];
        bool ok = true, seen[4] = {false, false, false, false};
        for (int c = 0; c < 4 && ok; c++) {
            if (f->chi[c] - f->clo[c] != f->N / 2) { ok = false; break; }
            int cl = -1;
            for (int h = f->clo[c]; h < f->chi[c] && ok; h++) {
                const int i = f->h_nt[h].x, j = f->h_nt[h].y;
                const int xi = i % LX, yi = i / LX, xj = j % LX, yj = j / LX;
                int b = -1;
                if (yi == yj) {
                    if ((xi + 1) % LX == xj) b = xi % 2;
                    else if ((xj + 1) % LX == xi) b = xj % 2;
                } else if (xi == xj) {
                    if ((yi + 1) % LY == yj) b = 2 + yi % 2;
                    else if ((yj + 1) % LY == yi) b = 2 + yj % 2;
                }
                if (b < 0 || (cl >= 0 && b != cl)) ok = false;
                cl = b;
            }
            if (ok && (cl < 0 || seen[cl])) ok = false;
            if (ok) { seen[cl] = true; cls[c] = cl; }
        }
        for (int c = 0; c < 4 && ok; c++) ok = (cls[c] == c);     // the kernels are written for the canonical colour order
        if (!ok) continue;
        f->v3_ok = 1; f->v3_lxl = LXL; f->v3_ry = RY;
        for (int c = 0; c < 4; c++) f->v3_cls[c] = cls[c];
        for (int mode : {0, 1, 2, 3, 6, 7, 10})
            SQ_CUDA(cudaFuncSetAttribute(pick3(LXL, RY, mode), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem_optin));
        if (pick3pb(0, LXL, RY, 2)) {
            f->v3_pb_ok = 1;
            for (int mode : {0, 1, 2, 3, 6, 7})
                SQ_CUDA(cudaFuncSetAttribute(pick3pb(0, LXL, RY, mode), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem_optin));
        }
        return;
    }
}

bool fdm_v3_supported(const sq_fdm *f, int S) {
    if (!(f->v3_ok && f->cs_coluni) && !f->v3_pb_ok) return false;
    if (getenv("SQ_V3_NO_PERBOND") && fdm_v3_perbond(f)) return false;
    if (S < 1 || S > 7) return false;
    return (size_t)S * f->N * sizeof(double) <= f->smem_optin;
}

struct CgFuse3 {
    const double2 *d; double2 *pnew; const CgState *cur; CgState *nxt;
    const double *rr_part, *beta_part; int nrr, nbeta, beta_complex, iter, check;
};
static thread_local const CgFuse3 *g_fuse3 = nullptr;

// returns the number of CTAs (= p.Ap partials for mode 2)
int fdm_v3_launch(sq_fdm *f, int mode, int S, double2 *out, const double2 *in, double *part, const CgState *skip, bool native, int nbatch,
                  size_t bstride, int bpart) {
    V3Params P;
    memset(&P, 0, sizeof(P));
    P.bstride = bstride; P.bpart = bpart;
    if (g_fuse3) {
        P.cg_d = g_fuse3->d; P.cg_pnew = g_fuse3->pnew; P.cg_cur = g_fuse3->cur; P.cg_nxt = g_fuse3->nxt;
        P.cg_rr_part = g_fuse3->rr_part; P.cg_beta_part = g_fuse3->beta_part; P.cg_nrr = g_fuse3->nrr; P.cg_nbeta = g_fuse3->nbeta;
        P.cg_beta_complex = g_fuse3->beta_complex; P.cg_iter = g_fuse3->iter; P.cg_check = g_fuse3->check;
    }
    P.L = (int)f->L; P.lb = f->slab_lo; P.le = f->slab_hi; P.S = S; P.C = (int)f->C; P.nphase = (mode == 2) ? 2 : 1;
    {
        // operand staging by bulk async copies pays once the vectors stream from HBM (measured: 0.52 -> 0.61 of the HBM peak at
        // 32 x 32 x 6400); while they fit L2 the plain loads are as fast and leave more CTAs per SM.  Tried and dropped: L1 prefetch
        // hints (slower below 6400 slices), staging only the combined slice (slower than staging both)
        const char *e = getenv("SQ_V3_PRE");
        P.pre = e ? atoi(e) : ((size_t)40 * f->N * f->L > ((size_t)100 << 20) ? 2 : 0);
    }
    for (int c = 0; c < 4; c++) { P.cls[c] = f->v3_cls[c]; P.clo[c] = c < f->C ? f->clo[c] : 0; }
    const bool pb = fdm_v3_perbond(f);
    if (pb) fdm_v3_prepare_native(f);                   // slot-ordered coefficients (version-checked: a no-op between operator updates)
    P.cs = f->cs.p; P.expV = f->expV.p; P.expVn = f->v3_expVn.p; P.ctn = f->v3_ctn.p; P.csn = f->v3_csn.p;
    static long long *dbg = nullptr;
    if (!dbg && getenv("SQ_DEBUG_STAMPS")) {
        SQ_CUDA(cudaMallocManaged((void **)&dbg, 16 * sizeof(long long)));
        for (int q = 0; q < 16; q++) dbg[q] = 0;
    }
    P.dbg = dbg;
    if (dbg && getenv("SQ_DEBUG_PRINT")) {
        cudaStreamSynchronize(f->stream);
        fprintf(stderr, "v3 stamps(cycles): init+load %lld B1 %lld combine %lld B2 %lld sync %lld store %lld total %lld\n", dbg[1] - dbg[0],
                dbg[2] - dbg[1], dbg[3] - dbg[2], dbg[5] - dbg[3], dbg[6] - dbg[5], dbg[7] - dbg[6], dbg[7] - dbg[0]);
    }
    const int nper = (mode == 2) ? S : S + 1;
    const int grid = (f->slab_hi - f->slab_lo + nper - 1) / nper;
    size_t smem = (mode == 2) ? (size_t)S * f->N * sizeof(double) : 0;
    if (native) SQ_REQUIRE(mode == 2 && f->v3_expVn.p && f->v3_expv_version == f->coef_version, "native-order operator not prepared");
    int kmode = ((mode == 2 && g_fuse3) ? 3 : mode) + (native ? 4 : 0);
    if (kmode == 6 && !pb && P.pre == 2 && (size_t)2 * (S + 1) * f->N * sizeof(double) <= f->smem_optin) {       // TMA-staged operands
        kmode = 10;
        smem = (size_t)2 * (S + 1) * f->N * sizeof(double);
    }
    v3_kernel_t k = pb ? pick3pb(f->v3_kind, f->v3_lxl, f->v3_ry, kmode)
                       : (f->v3_kind == 1 ? pick3h(f->v3_lxl, f->v3_ry, kmode) : pick3(f->v3_lxl, f->v3_ry, kmode));
    SQ_REQUIRE(k != nullptr, "register path: no kernel for this lattice / mode");
    SQ_CUDA(sq_launch(k, dim3(grid, 2, nbatch), dim3(32 * (S + 1)), smem, f->stream, P, out, (const double2 *)in, part, skip));
    SQ_LAUNCH_CHECK();
    f->launches++;
    return 2 * grid;
}

int fdm_v3_launch_cg(sq_fdm *f, double2 *z, const double2 *p_old, double2 *p_new, const double2 *d, const CgState *cur, CgState *nxt,
                     const double *rr_part, int nrr, const double *beta_part, int nbeta, int beta_complex, int iter, int check,
                     double *pAp_part, bool native) {
    CgFuse3 a = {d, p_new, cur, nxt, rr_part, beta_part, nrr, nbeta, beta_complex, iter, check};
    g_fuse3 = &a;
    int n = 0;
    try {
        n = fdm_v3_launch(f, 2, f->v3_S, z, p_old, pAp_part, cur, native);
    } catch (...) { g_fuse3 = nullptr; throw; }
    g_fuse3 = nullptr;
    return n;
}

// ---------------------------------------------------------------------------------------------------
// Resident CG: the whole unpreconditioned solve in ONE cooperative launch.  One CTA per SM owns S slices (both parts:
// 2 (S+1) warps) for the whole solve and nothing except two boundary slices per CTA touches global memory inside an iteration:
//   * x and r of the warp's slice live in REGISTERS, p (own + 2 halo slices) and the w hand-over in SHARED memory;
//   * the neighbours' halo slices are rebuilt locally (see k_cg_v3_resident1 below) -- no extra synchronisation.
// (Earlier generations -- a persistent kernel that streamed r / p / x through L2, 21 us per iteration at cfg4, and a resident
// kernel with two grid-wide sums, 9.4 us -- were removed in round 2; the launch loop in cg.cu is the only fallback.)
// ---------------------------------------------------------------------------------------------------
// Grid-wide deterministic sum without atomics or a separate barrier: every CTA publishes (value, epoch) in its own 16-byte
// slot (fence + relaxed store of the epoch), warp 0 of every CTA polls all slots for the epoch with relaxed loads that
// bypass L1 and adds the values in a fixed order -- the same order in every CTA, so all CTAs get the same bits.  The
// fence orders every earlier write of the CTA (cumulative through the preceding __syncthreads).  This (value, epoch) form serves the
// exact check at the end of a solve; the per-iteration sums use the fence-free tagged form below (v3_slot_sum2).  No L1 invalidation:
// shared data is only ever read with L2 loads.
struct V3Slot { double v; unsigned long long e; };
// warp-level: `t` is this CTA's partial (same value in all lanes); publishes it and returns the grid total (all lanes).
// One warp polls with up to 8 slots per lane in flight per round.  (Measured alternatives: every thread of the CTA polling
// one slot each, or two warps sharing the slots, were 0.5 - 1 us SLOWER per sum -- more pollers, more L2 contention.)
__device__ __forceinline__ double v3_slot_sum(double t, V3Slot *slots, unsigned int stride, unsigned long long epoch, unsigned int nblk,
                                              unsigned int bid, bool &bad) {
    const int lane = threadIdx.x & 31;
    if (lane == 0) {
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(&slots[(size_t)bid * stride]), "l"(__double_as_longlong(t)), "l"(epoch) : "memory");
    }
    double s = 0.0;
    const long long t0 = clock64();
    for (unsigned int base = 0; base < nblk; base += 256) {
        long long val[8];
        unsigned long long ep[8];
        while (true) {
#pragma unroll
            for (int u = 0; u < 8; u++) {                         // all loads first, then the checks
                const unsigned int q = base + lane + 32 * u;
                ep[u] = epoch;
                val[u] = 0;
                if (q < nblk) asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(val[u]), "=l"(ep[u]) : "l"(&slots[(size_t)q * stride]) : "memory");
            }
            bool ready = true;
#pragma unroll
            for (int u = 0; u < 8; u++) ready = ready && (ep[u] >= epoch);
            if (ready) break;
            if (clock64() - t0 > 4000000000LL) { bad = true; break; }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) s += __longlong_as_double(val[u]);
    }
    return warp_sum(s);
}
__device__ __forceinline__ double v3_grid_sum(double acc, double *red, double *sh_out, V3Slot *slots, unsigned int stride,
                                              unsigned long long epoch, unsigned int nblk, unsigned int bid, bool &aborted) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    int bad = 0;
    if (warp == 0) {
        double t = lane < nw ? red[lane] : 0.0;
        t = warp_sum(t);
        bool b = false;
        const double s = v3_slot_sum(t, slots, stride, epoch, nblk, bid, b);
        bad = b ? 1 : 0;
        if (lane == 0) *sh_out = s;
    }
    aborted = __syncthreads_or(bad) != 0;
    return *sh_out;
}

// ---------------------------------------------------------------------------------------------------
// Resident CG with ONE grid-wide sum per iteration.
//
// A grid-wide sum costs >= 2 us on this two-die part and the textbook recurrence needs two (p.Ap, then |r_new|^2).  Both
// follow from four sums taken BEFORE the update:
//     a = |M p|^2 = p.Ap,   b = r.z,   c = |z|^2,   d = |r|^2            (z = A p, r the current residual)
//     alpha = d / a,   |r - alpha z|^2 = d - 2 alpha b + alpha^2 c,   beta = |r_new|^2 / d
// d is the exact norm of the stored residual, recomputed every iteration, so the estimate never feeds back into itself (no
// drift); its rounding error is a few ulp of d, the order of the error of the updated residual itself.  The estimate drives
// beta and the stopping test; a positive test (and the maxiter exit) is confirmed with the exact |r_new|^2 -- one extra
// sum, once per solve -- so the returned eps is exact and the solver never stops early.
// Halo: with alpha and beta known to everybody, a CTA can update its copies of the neighbours' boundary r and p itself
// if it knows their boundary z: r_h -= alpha z_h, p_h = r_h + beta p_h (the same fma's the owner executes: same bits).
// The boundary slices of z are stored before the sum and read after it (one GPU, slices above 256 sites: published by a device-scope
// fence that an otherwise idle warp executes in parallel with the sum, then a per-CTA flag) or while it is in flight (several GPUs and
// small slices: every word of the boundary carries a validity tag, no fence at all) -- see FLAGS / TAGH in the kernel.
// The four partials travel as two 16-byte stores per CTA; the two low mantissa bits of every double carry the iteration
// count mod 4 as validity tag (slot arrays alternate with the parity of the iteration), so one polling load returns two
// values and their tag together.
// ---------------------------------------------------------------------------------------------------
struct CgResident1 {
    double *x;                  // native order, in: start vector, out: solution
    const double *r;            // native order, initial residual
    CgState *state;             // in: normb, tol, eps ; out: iters, eps, done
    char *slots;                // 2 arrays (iteration parity) x nblk x slot_stride bytes: (a, b), (c, d) of every CTA
    char *slots_check;          // nblk x slot_stride bytes: (value, epoch) slots of the exact check
    unsigned int slot_stride;   // bytes
    size_t slot_array_bytes;
    double *halo;               // [cta][side][part][N] boundary slices of z
    unsigned long long *flags;  // [cta] (128 bytes apart): last iteration whose boundary z of that CTA is visible device-wide
    int maxiter;
    // tau-slab mode over several GPUs (MULTI kernels): every rank runs this kernel on its slab; the grid-wide sums run over the
    // CTAs of ALL ranks and the boundary z of the first / last CTA travels to the neighbour rank.  Each rank owns a mailbox
    // (peer-mapped through CUDA IPC): slots of all global CTAs, then the check slots, then the halo inbox
    // [parity][side: 0 from the left rank, 1 from the right rank][part][N].  Producers PUSH into the consumers' mailboxes
    // (posted NVLink stores), consumers poll their own memory.  it_base continues the iteration count across solves so that
    // the validity tags never repeat within a slot (no mailbox reset between solves, which would race with early peers).
    int world, rank;
    unsigned int gid0, gtot;    // first global CTA index of this rank, number of CTAs of all ranks
    unsigned long long it_base;
    char *mail[8];
    size_t off_check, off_inbox;
};

__device__ __forceinline__ double v3_tag(double x, long long tag) { return __longlong_as_double((__double_as_longlong(x) & ~3LL) | tag); }

// Two warps share the work: warp `half` (0 or 1) publishes and polls the 16-byte half `half` of every slot -- (a, b) or (c, d).
// All lanes hold the CTA's two partials of that half in t[]; returns the two grid totals (identical bits in every CTA).
#ifdef SQ_V3_STAMPS
__device__ long long g_sumdbg[8];       // profiling build: fence cycles, store -> all slots valid, poll rounds, sums (CTA nblk / 2, warp 0)
#endif
template <bool FENCE>
__device__ __forceinline__ void v3_slot_sum2(double (&t)[2], int half, char *slots, unsigned int stride_bytes, long long tag, unsigned int nblk,
                                             unsigned int bid, bool &bad) {
    const int lane = threadIdx.x & 31;
#ifdef SQ_V3_STAMPS
    const bool sdbg = bid == nblk / 2 && half == 0 && lane == 0;
    const long long ts0 = clock64();
    long long ts1 = ts0, rounds = 0;
#endif
    if (lane == 0) {
        char *mine = slots + (size_t)bid * stride_bytes + 16 * half;
        if (FENCE) asm volatile("fence.acq_rel.gpu;" ::: "memory");
#ifdef SQ_V3_STAMPS
        ts1 = clock64();
#endif
        asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1, %2};" ::"l"(mine), "d"(v3_tag(t[0], tag)), "d"(v3_tag(t[1], tag)) : "memory");
    }
    double s[2] = {0.0, 0.0};
    const long long t0 = clock64();
    for (unsigned int base = 0; base < nblk; base += 256) {
        long long val[8][2];
        while (true) {
#ifdef SQ_V3_STAMPS
            rounds++;
#endif
#pragma unroll
            for (int u = 0; u < 8; u++) {                         // all loads first, then the checks
                const unsigned int q = base + lane + 32 * u;
                val[u][0] = val[u][1] = tag;
                if (q < nblk) asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(val[u][0]), "=l"(val[u][1]) : "l"(slots + (size_t)q * stride_bytes + 16 * half) : "memory");
            }
            bool ready = true;
#pragma unroll
            for (int u = 0; u < 8; u++) ready = ready && ((val[u][0] & 3LL) == tag) && ((val[u][1] & 3LL) == tag);
#ifdef SQ_V3_STAMPS
            if (sdbg && rounds == 1) g_sumdbg[4] += clock64() - ts1;      // one poll round trip
            if (sdbg && !ready) { int miss = 0; for (int u = 0; u < 8; u++) miss += ((val[u][0] & 3LL) != tag); g_sumdbg[5] += miss; }
#endif
            if (ready) break;
            if (clock64() - t0 > 4000000000LL) { bad = true; break; }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const unsigned int q = base + lane + 32 * u;
            if (q < nblk) { s[0] += __longlong_as_double(val[u][0]); s[1] += __longlong_as_double(val[u][1]); }
        }
    }
    t[0] = warp_sum(s[0]);
    t[1] = warp_sum(s[1]);
#ifdef SQ_V3_STAMPS
    if (sdbg) { g_sumdbg[0] += ts1 - ts0; g_sumdbg[1] += clock64() - ts1; g_sumdbg[2] += rounds; g_sumdbg[3] += 1; }
#endif
}

// Multi-GPU versions: publish into every rank's mailbox (system scope), poll the own one.
template <bool FENCE>
__device__ __forceinline__ void v3_slot_sum2_multi(double (&t)[2], int half, char *const *mail, int world, int rank, size_t slot_off,
                                                   unsigned int stride_bytes, long long tag, unsigned int gtot, unsigned int gid, bool &bad) {
    const int lane = threadIdx.x & 31;
    if (lane == 0) {
        if (FENCE) asm volatile("fence.acq_rel.sys;" ::: "memory");
        for (int q = 0; q < world; q++) {
            char *dst = mail[q] + slot_off + (size_t)gid * stride_bytes + 16 * half;
            asm volatile("st.relaxed.sys.global.v2.f64 [%0], {%1, %2};" ::"l"(dst), "d"(v3_tag(t[0], tag)), "d"(v3_tag(t[1], tag)) : "memory");
        }
    }
    const char *slots = mail[rank] + slot_off;
    double s[2] = {0.0, 0.0};
    const long long t0 = clock64();
    for (unsigned int base = 0; base < gtot; base += 256) {
        long long val[8][2];
        while (true) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const unsigned int q = base + lane + 32 * u;
                val[u][0] = val[u][1] = tag;
                if (q < gtot) asm volatile("ld.relaxed.sys.global.v2.b64 {%0, %1}, [%2];" : "=l"(val[u][0]), "=l"(val[u][1]) : "l"(slots + (size_t)q * stride_bytes + 16 * half) : "memory");
            }
            bool ready = true;
#pragma unroll
            for (int u = 0; u < 8; u++) ready = ready && ((val[u][0] & 3LL) == tag) && ((val[u][1] & 3LL) == tag);
            if (ready) break;
            if (clock64() - t0 > 8000000000LL) { bad = true; break; }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const unsigned int q = base + lane + 32 * u;
            if (q < gtot) { s[0] += __longlong_as_double(val[u][0]); s[1] += __longlong_as_double(val[u][1]); }
        }
    }
    t[0] = warp_sum(s[0]);
    t[1] = warp_sum(s[1]);
}
// exact check sum over all ranks: (value, epoch) slots at mail[.] + off
__device__ __forceinline__ double v3_grid_sum_multi(double acc, double *red, char *const *mail, int world, int rank, size_t off,
                                                    unsigned int stride_bytes, unsigned long long epoch, unsigned int gtot, unsigned int gid,
                                                    bool &aborted) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    int bad = 0;
    if (warp == 0) {
        double t = lane < nw ? red[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) {
            asm volatile("fence.acq_rel.sys;" ::: "memory");
            for (int q = 0; q < world; q++)
                asm volatile("st.relaxed.sys.global.v2.b64 [%0], {%1, %2};" ::"l"(mail[q] + off + (size_t)gid * stride_bytes), "l"(__double_as_longlong(t)), "l"(epoch) : "memory");
        }
        const char *slots = mail[rank] + off;
        double s = 0.0;
        const long long t0 = clock64();
        for (unsigned int q = lane; q < gtot; q += 32) {
            long long val;
            unsigned long long ep;
            while (true) {
                asm volatile("ld.relaxed.sys.global.v2.b64 {%0, %1}, [%2];" : "=l"(val), "=l"(ep) : "l"(slots + (size_t)q * stride_bytes) : "memory");
                if (ep >= epoch) break;
                if (clock64() - t0 > 8000000000LL) { bad = 1; break; }
            }
            s += __longlong_as_double(val);
        }
        s = warp_sum(s);
        if (lane == 0) red[32] = s;
    }
    aborted = __syncthreads_or(bad) != 0;
    return red[32];
}

template <class G, int MULTI>
__global__ void __launch_bounds__(256, 1)
k_cg_v3_resident1(const __grid_constant__ V3Params P, const CgResident1 C) {
    constexpr int N = G::N, NP = G::NP;
    extern __shared__ double smem[];
    __shared__ double red[4 * 8 + 1];
    __shared__ double sh[6];
    const int S = P.S, L = P.L;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int part = wid / (S + 1), k = wid - part * (S + 1);
    G E;
    E.template init<1>(P, part);
    const int l0 = P.lb + blockIdx.x * S;
    const int ns = min(S, P.le - l0);
    const unsigned int nblk = gridDim.x, bid = blockIdx.x;
    const unsigned int gtot = MULTI ? C.gtot : nblk, gid = MULTI ? C.gid0 + bid : bid;
    const bool active = k <= ns, owner = active && k >= 1, publish = k < ns;
    // Single GPU: the device-scope fence that publishes the boundary z (>= 1000 cycles on this part) is executed by a warp that is
    // idle at that point -- the slice-less warp of part 1 -- in parallel with the (fence-free) grid-wide sum; it then raises this CTA's
    // flag to the iteration number, and the neighbours wait for the flag before they fetch the boundary z.
    // (One GPU, slices above 256 sites only: see TAGH below for everything else.)
    constexpr bool FLAGS = true;
    // Small slices (N <= 256, one GPU): no fence and no flag at all.  The boundary z carries the iteration tag in the two low mantissa bits of
    // every double (as the partial sums do) -- the owner rounds its boundary z to the tagged value BEFORE using it, so both sides compute
    // with the same bits -- and the slice-less warps poll the neighbours' boundary slices WHILE the grid-wide sum is in flight (the sum is run
    // by two owner warps instead): on these lattices the iteration is a chain of L2 round trips, and this takes two of them off the chain.
    // (Measured: 64-site chain 2.77 -> 2.40 us, 8 x 8 honeycomb 3.53 -> 3.08 us, 16 x 16 square 4.08 -> 3.60 us per iteration with all polling
    // loads in flight at once -- one load at a time made 16 x 16 SLOWER, 4.7 us; at 32 x 32 / 24 x 24 honeycomb it changes nothing, 7.12 / 7.47 us:
    // there the slice-less warps have slack and the fence-on-idle-warp protocol stays.)
    // Several GPUs: always -- a system-scope fence waits for the NVLink stores of the boundary to be acknowledged and the flag then crosses
    // the link once more: with tags the boundary is in the neighbour rank's inbox one link latency after the z phase, before the sum completes.
    constexpr bool TAGH = MULTI || G::N <= 256;
    auto flag_of = [&](unsigned int cta) -> unsigned long long * { return C.flags + (size_t)cta * 16; };
    const bool fwarp = FLAGS && !TAGH && part == 1 && k == 0, bwarp = FLAGS && owner && (k == 1 || k == ns);
    const bool sumw = TAGH ? (k == 1) : (wid < 2);       // the two warps that run the grid-wide sum: half 0 = (a, b), half 1 = (c, d)
    const int sumh = TAGH ? part : wid;
    const int fcnt = 32 * (1 + (ns > 1 ? 4 : 2)), tcnt = (int)blockDim.x;
    int lself = l0 + k;
    lself = lself >= L ? lself - L : lself;
    const int lB = lself, lo = l0 + k - 1;               // owner: slice lo
    const double sg = (lB == 0) ? 1.0 : -1.0;
    if (active) E.load_cs(P, lB);
    // shared memory (doubles): p [part][S+2][N] | w [part][S][N] | exp(-dtau V) [S+1][N] | r of the halo slices [part][2][N]
    double2 *Pb = reinterpret_cast<double2 *>(smem) + (size_t)part * (S + 2) * (N / 2);
    double2 *W = reinterpret_cast<double2 *>(smem) + (size_t)2 * (S + 2) * (N / 2) + (size_t)part * S * (N / 2);
    double *EV = smem + (size_t)2 * (2 * S + 2) * N;
    double2 *Rh = reinterpret_cast<double2 *>(EV + (size_t)(S + 1) * N) + (size_t)part * 2 * (N / 2);
    const double *evk = EV + (size_t)k * N + 2 * lane;
    auto el = [&](int u) -> int { return u * 32 + lane; };
    auto gslice = [&](const double *base, int l) -> const double2 * { return reinterpret_cast<const double2 *>(base + ((size_t)l * 2 + part) * N); };
    // boundary z: [parity of the iteration][cta][side][part][N]; side 0 = first own slice, 1 = last own slice
    auto hslice = [&](int par, unsigned int cta, int side) -> double2 * { return reinterpret_cast<double2 *>(C.halo + ((((size_t)par * nblk + cta) * 2 + side) * 2 + part) * N); };
    const unsigned int left = (bid + nblk - 1) % nblk, right = (bid + 1) % nblk;
    // MULTI: halo inbox of rank q, [parity][side][part][N] doubles
    auto inbox = [&](int q, int par, int side) -> double2 * { return reinterpret_cast<double2 *>(C.mail[q] + C.off_inbox) + ((size_t)(par * 2 + side) * 2 + part) * (N / 2); };
    const int rank_l = MULTI ? (C.rank + C.world - 1) % C.world : 0, rank_r = MULTI ? (C.rank + 1) % C.world : 0;
    const double normb = C.state->normb, tol = C.state->tol;
    double eps = C.state->eps;
    int it = 0, done = 0;
    double v[G::NV], xr[G::NV], rr_[G::NV];
    double rr_lane = 0.0;        // this lane's share of |r|^2 of the stored residual: accumulated by the update, summed with the next iteration's dots
    if (threadIdx.x == 0) sh[5] = 0.0;
#ifdef SQ_V3_STAMPS
    // per-phase cycle counts of one owner warp (k = 1) and the halo warp (k = 0) of a CTA in the middle of the grid (profiling build only)
    const bool stamp = P.dbg && !MULTI && bid == nblk / 2 && part == 0 && lane == 0 && (k == 0 || k == 1);
    long long tph[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tprev = stamp ? clock64() : 0;
#define R1_STAMP(q) do { if (stamp) { const long long tn = clock64(); tph[q] += tn - tprev; tprev = tn; } } while (0)
#else
#define R1_STAMP(q) do { } while (0)
#endif
    if (part == 0 && active) {
        const double2 *g = reinterpret_cast<const double2 *>(P.expVn + (size_t)lB * N);
        for (int e = lane; e < N / 2; e += 32) reinterpret_cast<double2 *>(EV + (size_t)k * N)[e] = __ldg(g + e);
    }
    if (owner) {
        const double2 *gx = gslice(C.x, lo), *gr = gslice(C.r, lo);
#pragma unroll
        for (int u = 0; u < NP; u++) {
                const double2 a = gx[el(u)], b = gr[el(u)];
                xr[2 * u] = a.x; xr[2 * u + 1] = a.y;
                rr_[2 * u] = b.x; rr_[2 * u + 1] = b.y;
                rr_lane += b.x * b.x; rr_lane += b.y * b.y;
                Pb[(size_t)k * (N / 2) + el(u)] = b;
            }
    }
    if (k == 0) {                                         // r0 = p0 of the two halo slices
#pragma unroll 1
        for (int side = 0; side < 2; side++) {
            const int q = side ? ns + 1 : 0;
            int l = side ? l0 + ns : l0 - 1;
            l = l < 0 ? l + L : (l >= L ? l - L : l);
            const double2 *gr = gslice(C.r, l);
            for (int e = lane; e < N / 2; e += 32) { const double2 t = gr[e]; Pb[(size_t)q * (N / 2) + e] = t; Rh[(size_t)side * (N / 2) + e] = t; }
        }
    }
    __syncthreads();
#pragma unroll 1
    while (it < C.maxiter) {
        it++;
        const unsigned long long itg = MULTI ? C.it_base + (unsigned long long)it : (unsigned long long)it;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        // tag of the boundary slices of this iteration: 1, 2 or 3 -- never 0, the state of a cleared buffer -- and different for the iterations
        // it - 2 and it - 4 that used the same buffer before
        const long long htag = 1 + (long long)(itg % 3ULL);
        if (active) {
#pragma unroll
            for (int u = 0; u < NP; u++) {
                    const double2 a = Pb[(size_t)k * (N / 2) + el(u)];
                    v[2 * u] = a.x; v[2 * u + 1] = a.y;
                }
            E.template apply_B_ev<1, 1>(v, evk);
            R1_STAMP(0);
            if (k == ns && it > 1) asm volatile("bar.sync %0, 64;" ::"r"(1 + part) : "memory");       // upper halo slice rebuilt by warp 0
#pragma unroll
            for (int u = 0; u < NP; u++) {
                    const double2 self = Pb[(size_t)(k + 1) * (N / 2) + el(u)];
                    const double w0 = fma(sg, v[2 * u], self.x), w1 = fma(sg, v[2 * u + 1], self.y);
                    v[2 * u] = w0; v[2 * u + 1] = w1;
                    if (publish) {
                        acc[0] += w0 * w0;
                        acc[0] += w1 * w1;
                        W[(size_t)k * (N / 2) + el(u)] = make_double2(w0, w1);
                    }
                }
            R1_STAMP(1);
        }
        if (owner) E.template apply_B_ev<1, 1>(v, evk);
        R1_STAMP(2);
        __syncthreads();                                  // w of all slices is in W
        R1_STAMP(3);
        if (owner) {                                      // z[lo] in registers; r.z, |z|^2, |r|^2; boundary z for the neighbours
            double2 *h0 = (k == 1) ? hslice((int)(itg & 1), bid, 0) : nullptr, *h1 = (k == ns) ? hslice((int)(itg & 1), bid, 1) : nullptr;
            if (MULTI) {                                  // the slab's outer boundaries go to the neighbour ranks' inboxes
                if (k == 1 && bid == 0) h0 = inbox(rank_l, (int)(itg & 1), 1);            // I am their right neighbour
                if (k == ns && bid == nblk - 1) h1 = inbox(rank_r, (int)(itg & 1), 0);    // I am their left neighbour
            }
#pragma unroll
            for (int u = 0; u < NP; u++) {
                    const double2 w = W[(size_t)(k - 1) * (N / 2) + el(u)];
                    double z0 = fma(sg, v[2 * u], w.x), z1 = fma(sg, v[2 * u + 1], w.y);
                    if (TAGH && bwarp) { z0 = v3_tag(z0, htag); z1 = v3_tag(z1, htag); }
                    v[2 * u] = z0; v[2 * u + 1] = z1;
                    const double r0 = rr_[2 * u], r1 = rr_[2 * u + 1];
                    acc[1] += r0 * z0; acc[1] += r1 * z1;
                    acc[2] += z0 * z0; acc[2] += z1 * z1;
                    if (TAGH && MULTI) {                  // self-validating words: strong stores, no fence
                        if (h0) asm volatile("st.relaxed.sys.global.v2.f64 [%0], {%1, %2};" ::"l"(h0 + el(u)), "d"(z0), "d"(z1) : "memory");
                        if (h1) asm volatile("st.relaxed.sys.global.v2.f64 [%0], {%1, %2};" ::"l"(h1 + el(u)), "d"(z0), "d"(z1) : "memory");
                    } else if (TAGH) {
                        if (h0) asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1, %2};" ::"l"(h0 + el(u)), "d"(z0), "d"(z1) : "memory");
                        if (h1) asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1, %2};" ::"l"(h1 + el(u)), "d"(z0), "d"(z1) : "memory");
                    } else {
                        if (h0) h0[el(u)] = make_double2(z0, z1);
                        if (h1) h1[el(u)] = make_double2(z0, z1);
                    }
                }
            acc[3] = rr_lane;                             // same elements, same order as a fresh accumulation: same bits
            if (bwarp && !TAGH) {                         // boundary z stored: hand it to the fencing warp
                __threadfence_block();
                asm volatile("bar.arrive 3, %0;" ::"r"(fcnt) : "memory");
            }
        }
        R1_STAMP(4);
        // ---- the grid-wide sum of (a, b, c, d)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const double t = warp_sum(acc[c]);
            if (lane == 0) red[c * 8 + wid] = t;
        }
        if (fwarp) {                                      // arrives without waiting, then: boundary z of both parts -> fence -> flag
            __threadfence_block();
            asm volatile("bar.arrive 4, %0;" ::"r"(tcnt) : "memory");       // (its own barrier id: this warp runs ahead of the others)
            asm volatile("bar.sync 3, %0;" ::"r"(fcnt) : "memory");
            if (lane == 0) {
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
                asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(flag_of(bid)), "l"(itg) : "memory");
            }
            __syncwarp();
        } else if (FLAGS) {
            asm volatile("bar.sync 4, %0;" ::"r"(tcnt) : "memory");
        } else {
            __syncthreads();
        }
        R1_STAMP(5);
        if (sumw) {                                       // half 0: (a, b), half 1: (c, d)
            const int nw = blockDim.x >> 5;
            double t[2];
#pragma unroll
            for (int c = 0; c < 2; c++) { t[c] = 0.0; for (int w = 0; w < nw; w++) t[c] += red[(2 * sumh + c) * 8 + w]; }
            bool bad = false;
            if (MULTI) v3_slot_sum2_multi<false>(t, sumh, C.mail, C.world, C.rank, (size_t)(itg & 1) * C.slot_array_bytes, C.slot_stride, (long long)(itg & 3), gtot, gid, bad);
            else v3_slot_sum2<false>(t, sumh, C.slots + (size_t)(it & 1) * C.slot_array_bytes, C.slot_stride, (long long)(it & 3), nblk, bid, bad);
            if (lane == 0) { sh[2 * sumh] = t[0]; sh[2 * sumh + 1] = t[1]; if (bad) sh[5] = 1.0; }
        } else if (TAGH && k == 0) {
            // meanwhile: the neighbours' boundary z of this iteration, valid word by word once its tag matches
            const double2 *gu = hslice((int)(itg & 1), right, 0), *gl = hslice((int)(itg & 1), left, 1);
            if (MULTI) {                                  // the slab's outer boundaries arrive in this rank's inbox
                if (bid == nblk - 1) gu = inbox(C.rank, (int)(itg & 1), 1);
                if (bid == 0) gl = inbox(C.rank, (int)(itg & 1), 0);
            }
            const long long tg = htag, t0 = clock64();
            constexpr int NH = N / 64;                    // double2 per lane and boundary slice
            long long q[2 * NH][2];
            while (true) {                                // all loads in flight, then the checks: one L2 round trip per attempt
#pragma unroll
                for (int u = 0; u < NH; u++) {
                    if (MULTI) {
                        asm volatile("ld.relaxed.sys.global.v2.b64 {%0, %1}, [%2];" : "=l"(q[u][0]), "=l"(q[u][1]) : "l"(gl + lane + 32 * u) : "memory");
                        asm volatile("ld.relaxed.sys.global.v2.b64 {%0, %1}, [%2];" : "=l"(q[NH + u][0]), "=l"(q[NH + u][1]) : "l"(gu + lane + 32 * u) : "memory");
                    } else {
                        asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(q[u][0]), "=l"(q[u][1]) : "l"(gl + lane + 32 * u) : "memory");
                        asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(q[NH + u][0]), "=l"(q[NH + u][1]) : "l"(gu + lane + 32 * u) : "memory");
                    }
                }
                bool ready = true;
#pragma unroll
                for (int u = 0; u < 2 * NH; u++) ready = ready && (q[u][0] & 3LL) == tg && (q[u][1] & 3LL) == tg;
                if (ready) break;
                if (clock64() - t0 > (MULTI ? 8000000000LL : 4000000000LL)) { sh[5] = 1.0; break; }
            }
#pragma unroll
            for (int u = 0; u < NH; u++) {
                xr[2 * u] = __longlong_as_double(q[u][0]); xr[2 * u + 1] = __longlong_as_double(q[u][1]);
                rr_[2 * u] = __longlong_as_double(q[NH + u][0]); rr_[2 * u + 1] = __longlong_as_double(q[NH + u][1]);
            }
        }
        R1_STAMP(6);
        __syncthreads();
        R1_STAMP(7);
        if (sh[5] != 0.0) { done = 3; break; }
        const double pAp = sh[0], rz = sh[1], zz = sh[2], rr_old = sh[3];
        const double alpha = rr_old / pAp;
        double rr_new = fma(alpha, fma(alpha, zz, -2.0 * rz), rr_old);       // |r - alpha z|^2
        rr_new = rr_new > 0.0 ? rr_new : (rr_new == rr_new ? 0.0 : rr_new);
        eps = sqrt(rr_new) / normb;
        const bool stop_est = (eps < tol) || !(eps == eps) || it == C.maxiter;
        const double beta = rr_new / rr_old;
        double chk = 0.0;
        if (owner) {                                      // x += alpha p ; r -= alpha z ; p = r + beta p
#pragma unroll
            for (int u = 0; u < NP; u++) {
                    const double2 pv = Pb[(size_t)k * (N / 2) + el(u)];
                    xr[2 * u] = fma(alpha, pv.x, xr[2 * u]);
                    xr[2 * u + 1] = fma(alpha, pv.y, xr[2 * u + 1]);
                    const double r0 = fma(-alpha, v[2 * u], rr_[2 * u]), r1 = fma(-alpha, v[2 * u + 1], rr_[2 * u + 1]);
                    rr_[2 * u] = r0; rr_[2 * u + 1] = r1;
                    chk += r0 * r0; chk += r1 * r1;
                    Pb[(size_t)k * (N / 2) + el(u)] = make_double2(fma(beta, pv.x, r0), fma(beta, pv.y, r1));
                }
            rr_lane = chk;
        }
        // Warp 0 of each part owns no slice.  Right after the sum it fetches the neighbours' boundary z (published before the sum, so
        // visible now) into registers -- its xr / rr_ registers, which only owners use: N / 64 double2 per lane = NV doubles -- so that
        // the L2 round trip of the two 8 KB slices overlaps with the owners' update; the halo copies of r and p are rebuilt from them
        // after the barrier below, overlapped with the owners' first B.  (Round 1 issued the loads after the barrier: the chain L2 load ->
        // rebuild -> this warp's own B was then the longest of the CTA, on the critical path of every iteration.  Measured alternatives:
        // the whole rebuild before the barrier 8.11 us, staging through shared memory with bulk async copies 7.73 us, this 7.45 us,
        // round 1 7.95 us per iteration at cfg4.)
        static_assert(G::NV == N / 32, "one warp holds one slice-part");
        if (k == 0 && !TAGH) {
            const double2 *gu = hslice((int)(itg & 1), right, 0), *gl = hslice((int)(itg & 1), left, 1);
            if (FLAGS) {                                  // the neighbours' boundary z of this iteration is visible once their flag says so
                if (lane < 2) {
                    const unsigned long long *fl = flag_of(lane ? right : left);
                    unsigned long long got;
                    const long long t0 = clock64();
                    while (true) {
                        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(got) : "l"(fl) : "memory");
                        if (got >= itg) break;
                        if (clock64() - t0 > 4000000000LL) { sh[5] = 1.0; break; }
                    }
                }
                __syncwarp();
            }
#pragma unroll
            for (int u = 0; u < N / 64; u++) { const double2 q = __ldcg(gl + lane + 32 * u); xr[2 * u] = q.x; xr[2 * u + 1] = q.y; }
#pragma unroll
            for (int u = 0; u < N / 64; u++) { const double2 q = __ldcg(gu + lane + 32 * u); rr_[2 * u] = q.x; rr_[2 * u + 1] = q.y; }
        }
        if (stop_est) {                                   // confirm with the exact |r_new|^2 (every CTA takes this branch together)
            bool aborted;
            const double rr_exact = MULTI ? v3_grid_sum_multi(chk, red, C.mail, C.world, C.rank, C.off_check, C.slot_stride, itg, gtot, gid, aborted)
                                          : v3_grid_sum(chk, red, &sh[4], reinterpret_cast<V3Slot *>(C.slots_check), C.slot_stride / 16, (unsigned long long)it, nblk, bid, aborted);
            if (aborted) { done = 3; break; }
            eps = sqrt(rr_exact) / normb;
            if (eps < tol) { done = 1; break; }
            if (!(eps == eps)) { done = 2; break; }
            if (it == C.maxiter) break;
        }
        R1_STAMP(8);
        __syncthreads();                                  // own p slices are in Pb: the owners start the next B right away
        R1_STAMP(9);
        if (k == 0) {
            // r_h -= alpha z_h, p_h = r_h + beta p_h (the same fma's the owner executes): lower slice first (this warp's own B needs it),
            // then the upper one, handed to warp ns through a named barrier
#pragma unroll
            for (int u = 0; u < N / 64; u++) {
                const size_t e = lane + 32 * u;
                const double2 rv = Rh[e], pv = Pb[e];
                const double r0 = fma(-alpha, xr[2 * u], rv.x), r1 = fma(-alpha, xr[2 * u + 1], rv.y);
                Rh[e] = make_double2(r0, r1);
                Pb[e] = make_double2(fma(beta, pv.x, r0), fma(beta, pv.y, r1));
            }
#pragma unroll
            for (int u = 0; u < N / 64; u++) {
                const size_t e = lane + 32 * u;
                const double2 rv = Rh[(size_t)1 * (N / 2) + e], pv = Pb[(size_t)(ns + 1) * (N / 2) + e];
                const double r0 = fma(-alpha, rr_[2 * u], rv.x), r1 = fma(-alpha, rr_[2 * u + 1], rv.y);
                Rh[(size_t)1 * (N / 2) + e] = make_double2(r0, r1);
                Pb[(size_t)(ns + 1) * (N / 2) + e] = make_double2(fma(beta, pv.x, r0), fma(beta, pv.y, r1));
            }
            __threadfence_block();
            asm volatile("bar.arrive %0, 64;" ::"r"(1 + part) : "memory");
            __syncwarp();
        }
    }
    if (owner) {                                          // the solution
        double2 *gx = reinterpret_cast<double2 *>(C.x + ((size_t)lo * 2 + part) * N);
#pragma unroll
        for (int u = 0; u < NP; u++) gx[el(u)] = make_double2(xr[2 * u], xr[2 * u + 1]);
    }
#ifdef SQ_V3_STAMPS
    if (stamp) { for (int q = 0; q < 10; q++) P.dbg[16 * k + q] = tph[q]; P.dbg[16 * k + 10] = it; }
#endif
    if (bid == 0 && threadIdx.x == 0) {
        CgState st = *C.state;
        st.iters = it;
        st.eps = eps;
        st.done = done;
        *C.state = st;
    }
}

typedef void (*v3_resident1_t)(const V3Params, const CgResident1);
static v3_resident1_t pick3_resident1(int lxl, int ry) {
    if (lxl == 8 && ry == 4) return k_cg_v3_resident1<V3Lane<8, 4>, 0>;
    if (lxl == 8 && ry == 8) return k_cg_v3_resident1<V3Lane<8, 8>, 0>;
    if (lxl == 4 && ry == 2) return k_cg_v3_resident1<V3Lane<4, 2>, 0>;
    if (lxl == 4 && ry == 4) return k_cg_v3_resident1<V3Lane<4, 4>, 0>;
    if (lxl == 4 && ry == 8) return k_cg_v3_resident1<V3Lane<4, 8>, 0>;
    return nullptr;
}
static v3_resident1_t pick3h_resident1(int L1, int L2) {
    if (L1 == 24 && L2 == 24) return k_cg_v3_resident1<V3Honey<8, 3, 6>, 0>;
    if (L1 == 16 && L2 == 16) return k_cg_v3_resident1<V3Honey<4, 4, 2>, 0>;
    if (L1 == 8 && L2 == 8) return k_cg_v3_resident1<V3Honey<4, 2, 1>, 0>;
    return nullptr;
}
static v3_resident1_t pick3pb_resident1(int kind, int a, int b) {
    if (kind == 0 && a == 4 && b == 2) return k_cg_v3_resident1<V3LanePB<4, 2>, 0>;
    if (kind == 2 && a == 1) return k_cg_v3_resident1<V3ChainPB<1>, 0>;
    if (kind == 2 && a == 2) return k_cg_v3_resident1<V3ChainPB<2>, 0>;
    if (kind == 2 && a == 4) return k_cg_v3_resident1<V3ChainPB<4>, 0>;
    return nullptr;
}
// tau-slab over several GPUs: the geometries of the named multi-GPU configurations
static v3_resident1_t pick3_resident1_multi(int kind, int a, int b) {
    if (kind == 0 && a == 8 && b == 8) return k_cg_v3_resident1<V3Lane<8, 8>, 1>;          // 32 x 32 square
    if (kind == 0 && a == 4 && b == 2) return k_cg_v3_resident1<V3Lane<4, 2>, 1>;          // 16 x 16 square
    if (kind == 1 && a == 24 && b == 24) return k_cg_v3_resident1<V3Honey<8, 3, 6>, 1>;    // 24 x 24 honeycomb
    if (kind == 1 && a == 8 && b == 8) return k_cg_v3_resident1<V3Honey<4, 2, 1>, 1>;      // 8 x 8 honeycomb
    return nullptr;
}

// One-sum resident kernel (k_cg_v3_resident1).  Returns false if it cannot run (the caller falls back to the launch loop in cg.cu).
// graph: the V3Graph kernel on the padded 64-site slices prepared by fdm_v3g_cg (its own native copies of the operator)
static bool fdm_v3_cg_resident1(sq_fdm *f, double2 *x, double2 *r, CgState *state, i64 maxiter, bool graph = false) {
    const bool pb = fdm_v3_perbond(f);
    v3_resident1_t k;
    if (graph) {
        k = k_cg_v3_resident1<V3Graph, 0>;
    } else {
        if (!fdm_v3_supported(f, 1)) return false;
        k = pb ? pick3pb_resident1(f->v3_kind, f->v3_lxl, f->v3_ry)
               : (f->v3_kind == 1 ? pick3h_resident1(f->v3_lxl, f->v3_ry) : pick3_resident1(f->v3_lxl, f->v3_ry));
    }
    if (!k) return false;
    const size_t Nslice = graph ? (size_t)V3Graph::N : (size_t)f->N;      // sites of a slice as the kernel sees it
    const int nsl = f->slab_hi - f->slab_lo;
    int S = (nsl + f->num_sms - 1) / f->num_sms;
    if (const char *e = getenv("SQ_V3_RESIDENT_SLAB")) S = atoi(e);
    S = std::max(S, 2);
    if (S > 3 || nsl < S) return false;
    const int grid = (nsl + S - 1) / S, T = 64 * (S + 1);
    const size_t smem = (size_t)(5 * S + 9) * Nslice * sizeof(double);
    if (smem > f->smem_optin || grid > f->num_sms || grid < 2) return false;
    V3Params P;
    memset(&P, 0, sizeof(P));
    P.L = (int)f->L; P.lb = f->slab_lo; P.le = f->slab_hi; P.S = S; P.C = (int)f->C; P.nphase = 2;
    for (int c = 0; c < 4; c++) { P.cls[c] = f->v3_cls[c]; P.clo[c] = c < f->C ? f->clo[c] : 0; }
    P.cs = f->cs.p; P.expV = f->expV.p; P.expVn = f->v3_expVn.p; P.ctn = f->v3_ctn.p; P.csn = f->v3_csn.p;
    if (graph) { P.expVn = f->v3g_expVn.p; P.csn = f->v3g_csn.p; P.gpart = f->v3g_part.p; }
    SQ_CUDA(cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)k, T, smem));
    if (per_sm * f->num_sms < grid) return false;
    unsigned int stride_bytes = 1024;
    if (const char *e = getenv("SQ_V3_SLOT_STRIDE")) stride_bytes = (unsigned)atoi(e);
    stride_bytes = std::max(32u, stride_bytes / 32 * 32);
    const size_t arr = (size_t)grid * stride_bytes;
    const size_t slot_bytes = 3 * arr + (size_t)grid * 128;
    if (f->v3_slots.n < slot_bytes) f->v3_slots.alloc(slot_bytes);
    SQ_CUDA(cudaMemsetAsync(f->v3_slots.p, 0, slot_bytes, f->stream));
    const size_t nh = (size_t)8 * grid * Nslice;
    if (f->v3_halo.n < nh) f->v3_halo.alloc(nh);
    if (Nslice <= 256) SQ_CUDA(cudaMemsetAsync(f->v3_halo.p, 0, nh * sizeof(double), f->stream));      // tagged boundary slices: no stale tags of an earlier solve
    CgResident1 C;
    memset(&C, 0, sizeof(C));
    C.x = (double *)x; C.r = (const double *)r; C.state = state;
    C.slots = f->v3_slots.p; C.slot_array_bytes = arr; C.slots_check = f->v3_slots.p + 2 * arr; C.slot_stride = stride_bytes;
    C.flags = reinterpret_cast<unsigned long long *>(f->v3_slots.p + 3 * arr);
    C.halo = f->v3_halo.p; C.maxiter = (int)std::min<i64>(maxiter, 2000000000);
#ifdef SQ_V3_STAMPS
    static long long *dbg = nullptr;
    if (!dbg && getenv("SQ_DEBUG_STAMPS")) {
        SQ_CUDA(cudaMallocManaged((void **)&dbg, 64 * sizeof(long long)));
        for (int q = 0; q < 64; q++) dbg[q] = 0;
    }
    P.dbg = dbg;
#endif
    void *args[] = {(void *)&P, (void *)&C};
    SQ_CUDA(cudaLaunchCooperativeKernel((const void *)k, dim3(grid), dim3(T), args, smem, f->stream));
    f->launches++;
#ifdef SQ_V3_STAMPS
    if (dbg) {
        cudaStreamSynchronize(f->stream);
        static const char *names[10] = {"load p + B1", "halo wait + combine", "B2", "sync (w ready)", "z + dots + halo store", "warp sums + sync",
                                        "grid-wide sum (warps 0/1) / idle", "sync (sum known)", "update x r p", "sync (p ready)"};
        for (int w = 0; w < 2; w++) {
            const double n = (double)std::max<long long>(dbg[16 * w + 10], 1);
            double tot = 0;
            for (int q = 0; q < 10; q++) tot += dbg[16 * w + q] / n;
            fprintf(stderr, "resident1 stamps, CTA %d/%d warp k=%d (%s), %lld iterations, %.0f cycles/iteration:\n", grid / 2, grid, w,
                    w ? "first owner" : "halo warp, sums", dbg[16 * w + 10], tot);
            for (int q = 0; q < 10; q++) fprintf(stderr, "    %-36s %8.0f\n", names[q], dbg[16 * w + q] / n);
        }
        long long sd[8];
        cudaMemcpyFromSymbol(sd, g_sumdbg, sizeof(sd));
        const double ns_ = (double)std::max<long long>(sd[3], 1);
        fprintf(stderr, "grid-wide sum of that CTA (lane 0 of warp 0, per sum): fence %.0f, store -> total %.0f cycles, %.2f poll rounds, first round trip %.0f "
                        "cycles, %.1f slots of this lane missing per failed round\n", sd[0] / ns_, sd[1] / ns_, sd[2] / ns_, sd[4] / ns_,
                sd[2] > sd[3] ? (double)sd[5] / (double)(sd[2] - sd[3]) : 0.0);
    }
#endif
    return true;
}

// ---- graph engine: small lattices ------------------------------------------------------------------------
__global__ void k_v3g_to_native(double *__restrict__ dst, const double2 *__restrict__ src, int L, int N) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)L * 64) return;
    const int l = (int)(idx / 64), i = (int)(idx % 64);
    const double2 v = i < N ? src[(size_t)l * N + i] : make_double2(0.0, 0.0);      // padding sites stay zero for the whole solve
    dst[((size_t)l * 2) * 64 + i] = v.x;
    dst[((size_t)l * 2 + 1) * 64 + i] = v.y;
}
__global__ void k_v3g_from_native(double2 *__restrict__ dst, const double *__restrict__ src, int L, int N) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)L * N) return;
    const int l = (int)(idx / N), i = (int)(idx % N);
    dst[idx] = make_double2(src[((size_t)l * 2) * 64 + i], src[((size_t)l * 2 + 1) * 64 + i]);
}
// operator in the engine's layout: exp(-dtau V) padded with zeros, (cosh, sinh) per slot with (1, 0) where a site has no bond
__global__ void k_v3g_operator(double *__restrict__ ev, double2 *__restrict__ csn, const double *__restrict__ expV, const double2 *__restrict__ cs,
                               const int *__restrict__ map, int L, int N, int Nh) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)L * 256) return;
    const int l = (int)(idx / 256), e = (int)(idx % 256);
    if (e < 64) ev[(size_t)l * 64 + e] = e < N ? expV[(size_t)l * N + e] : 0.0;
    const int h = map[e];
    csn[idx] = h >= 0 ? cs[(size_t)l * Nh + h] : make_double2(1.0, 0.0);
}

static void v3g_build_tables(sq_fdm *f) {
    std::vector<int> part(256), map(256, -1);
    for (int c = 0; c < 4; c++)
        for (int a = 0; a < 2; a++)
            for (int lane = 0; lane < 32; lane++) part[(c * 2 + a) * 32 + lane] = lane | (a << 5);      // its own partner: passes through
    for (int c = 0; c < (int)f->C; c++)
        for (int h = f->clo[c]; h < f->chi[c]; h++) {
            const int i = f->h_nt[h].x, j = f->h_nt[h].y;
            part[(c * 2 + (i & 1)) * 32 + i / 2] = (j / 2) | ((j & 1) << 5);
            part[(c * 2 + (j & 1)) * 32 + j / 2] = (i / 2) | ((i & 1) << 5);
            map[(c * 2 + (i & 1)) * 32 + i / 2] = h;
            map[(c * 2 + (j & 1)) * 32 + j / 2] = h;
        }
    f->v3g_part.alloc(256);
    f->v3g_csmap.alloc(256);
    f->v3g_part.upload(part.data(), 256, f->stream);
    f->v3g_csmap.upload(map.data(), 256, f->stream);
    SQ_CUDA(cudaStreamSynchronize(f->stream));
}

bool fdm_v3g_cg(sq_fdm *f, double2 *x, const double2 *r, bool zero_start, CgState *state, i64 maxiter) {
    if (!f->v3g_ok || f->slab_lo != 0 || f->slab_hi != (int)f->L) return false;
    const int L = (int)f->L;
    const int S = std::max(2, (L + f->num_sms - 1) / f->num_sms);
    if (S > 3 || L < S || (L + S - 1) / S < 2 || (L + S - 1) / S > f->num_sms) return false;      // (the launch below would refuse: nothing converted yet)
    const size_t nn = (size_t)L * 2 * 64;
    if (!f->v3g_part.p) v3g_build_tables(f);
    if (!f->v3g_x.p) { f->v3g_x.alloc(nn); f->v3g_r.alloc(nn); f->v3g_expVn.alloc((size_t)L * 64); f->v3g_csn.alloc((size_t)L * 256); f->v3g_version = -1; }
    cudaStream_t s = f->stream;
    if (f->v3g_version != f->coef_version) {
        k_v3g_operator<<<(unsigned)(((size_t)L * 256 + 255) / 256), 256, 0, s>>>(f->v3g_expVn.p, f->v3g_csn.p, f->expV.p, f->cs.p, f->v3g_csmap.p, L, (int)f->N,
                                                                                 (int)f->Nh);
        SQ_LAUNCH_CHECK();
        f->launches++;
        f->v3g_version = f->coef_version;
    }
    const unsigned gb = (unsigned)(((size_t)L * 64 + 255) / 256);
    k_v3g_to_native<<<gb, 256, 0, s>>>(f->v3g_r.p, r, L, (int)f->N);
    if (zero_start) SQ_CUDA(cudaMemsetAsync(f->v3g_x.p, 0, nn * sizeof(double), s));
    else k_v3g_to_native<<<gb, 256, 0, s>>>(f->v3g_x.p, x, L, (int)f->N);
    SQ_LAUNCH_CHECK();
    f->launches += zero_start ? 1 : 2;
    if (!fdm_v3_cg_resident1(f, (double2 *)f->v3g_x.p, (double2 *)f->v3g_r.p, state, maxiter, true)) return false;
    k_v3g_from_native<<<(unsigned)(((size_t)L * f->N + 255) / 256), 256, 0, s>>>(x, f->v3g_x.p, L, (int)f->N);
    SQ_LAUNCH_CHECK();
    f->launches++;
    return true;
}

// ---- tau-slab over several GPUs ---------------------------------------------------------------------
// Mailbox layout (bytes): [2 parity][gtot_max] main slots | [gtot_max] check slots | halo inbox [2][2][2][N] doubles
static const size_t V3_MAIL_STRIDE = 64, V3_MAIL_MAXCTA = 148 * 8;
static size_t v3_mail_off_check() { return 2 * V3_MAIL_MAXCTA * V3_MAIL_STRIDE; }
static size_t v3_mail_off_inbox() { return 3 * V3_MAIL_MAXCTA * V3_MAIL_STRIDE; }
size_t fdm_v3_mailbox_bytes(const sq_fdm *f) { return v3_mail_off_inbox() + (size_t)8 * f->N * sizeof(double); }

// Can this slab configuration run the multi-GPU resident kernel?  Fills S and the CTA counts of all ranks.
static bool v3_multi_plan(const sq_fdm *f, int *S_out, std::vector<int> *ctas) {
    if (f->world < 2 || f->world > 8 || !f->v3_ok || !f->cs_coluni || !f->mail_ready) return false;
    if (!pick3_resident1_multi(f->v3_kind, f->v3_lxl, f->v3_ry)) return false;
    const int L = (int)f->L, W = f->world, base = L / W, extra = L % W;
    int nmax = base + (extra ? 1 : 0);
    int S = std::max(2, (nmax + f->num_sms - 1) / f->num_sms);
    if (S > 3 || base < S) return false;
    ctas->clear();
    size_t tot = 0;
    for (int q = 0; q < W; q++) {
        const int n = base + (q < extra ? 1 : 0);
        ctas->push_back((n + S - 1) / S);
        tot += ctas->back();
    }
    if (tot > V3_MAIL_MAXCTA) return false;
    *S_out = S;
    return true;
}
bool fdm_v3_multi_possible(const sq_fdm *f) {
    int S;
    std::vector<int> c;
    return v3_multi_plan(f, &S, &c);
}

// Tagged boundary slices: clear this rank's boundary buffer and halo inbox so that no tag of an earlier solve validates.  Must be
// enqueued BEFORE the collective that precedes the solve on this rank's stream (the peers start their kernels -- and write into this
// inbox -- only after that collective, which needs this rank's part of it).
void fdm_v3_multi_reset_boundaries(sq_fdm *f) {
    int S;
    std::vector<int> ctas;
    if (!v3_multi_plan(f, &S, &ctas)) return;
    const size_t nh = (size_t)8 * ctas[f->rank] * f->N;
    if (f->v3_halo.n < nh) f->v3_halo.alloc(nh);
    SQ_CUDA(cudaMemsetAsync(f->v3_halo.p, 0, f->v3_halo.n * sizeof(double), f->stream));
    SQ_CUDA(cudaMemsetAsync((char *)f->mail_ptr[f->rank] + v3_mail_off_inbox(), 0, (size_t)8 * f->N * sizeof(double), f->stream));
}

// x (in/out), r (in): native order, own slab + the two halo slices of r valid.  state: normb / tol / eps0 set by the caller (global
// values).  Every rank calls this collectively.
bool fdm_v3_cg_resident1_multi(sq_fdm *f, double2 *x, double2 *r, CgState *state, i64 maxiter) {
    int S;
    std::vector<int> ctas;
    if (!v3_multi_plan(f, &S, &ctas)) return false;
    v3_resident1_t k = pick3_resident1_multi(f->v3_kind, f->v3_lxl, f->v3_ry);
    const int grid = ctas[f->rank], T = 64 * (S + 1);
    const size_t smem = (size_t)(5 * S + 9) * f->N * sizeof(double);
    if (smem > f->smem_optin || grid > f->num_sms) return false;
    V3Params P;
    memset(&P, 0, sizeof(P));
    P.L = (int)f->L; P.lb = f->slab_lo; P.le = f->slab_hi; P.S = S; P.C = (int)f->C; P.nphase = 2;
    for (int c = 0; c < 4; c++) { P.cls[c] = f->v3_cls[c]; P.clo[c] = c < f->C ? f->clo[c] : 0; }
    P.cs = f->cs.p; P.expV = f->expV.p; P.expVn = f->v3_expVn.p; P.ctn = f->v3_ctn.p;
    SQ_CUDA(cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)k, T, smem));
    if (per_sm * f->num_sms < grid) return false;
    const size_t nh = (size_t)8 * grid * f->N;
    if (f->v3_halo.n < nh) f->v3_halo.alloc(nh);
    CgResident1 C;
    memset(&C, 0, sizeof(C));
    C.x = (double *)x; C.r = (const double *)r; C.state = state;
    C.slot_stride = (unsigned)V3_MAIL_STRIDE; C.slot_array_bytes = V3_MAIL_MAXCTA * V3_MAIL_STRIDE;
    C.halo = f->v3_halo.p; C.maxiter = (int)std::min<i64>(maxiter, 2000000000);
    C.world = f->world; C.rank = f->rank; C.gtot = 0; C.gid0 = 0;
    for (int q = 0; q < f->world; q++) { if (q < f->rank) C.gid0 += ctas[q]; C.gtot += ctas[q]; C.mail[q] = (char *)f->mail_ptr[q]; }
    C.it_base = f->v3_it_base;
    C.off_check = v3_mail_off_check(); C.off_inbox = v3_mail_off_inbox();
    void *args[] = {(void *)&P, (void *)&C};
    SQ_CUDA(cudaLaunchCooperativeKernel((const void *)k, dim3(grid), dim3(T), args, smem, f->stream));
    f->launches++;
    return true;
}

// Whole-solve resident kernel (one CTA per SM, one grid-wide sum per iteration).  x (in/out) and r (in) are native-order
// vectors.  Returns false when the problem does not fit (more than 3 slices per SM, lattice without a register engine): the
// caller then runs the two-launches-per-iteration loop.
bool fdm_v3_cg_resident(sq_fdm *f, double2 *x, double2 *r, CgState *state, i64 maxiter) {
    return fdm_v3_cg_resident1(f, x, r, state, maxiter);
}

// ---------------------------------------------------------------------------------------------------
// native order  <->  library order [l][i] (complex interleaved)
// ---------------------------------------------------------------------------------------------------
// offset inside one slice-part (doubles).  kind 0: square (lxl, ry); kind 1: honeycomb (lxl = L1, ry = L2, block sizes from the
// engine table: G1 = 8 for L1 = 24, else 4)
__device__ __forceinline__ size_t v3_native_index(int i, int lxl, int ry, int kind) {
    if (kind == 1) {
        const int L1 = lxl, L2 = ry, G1 = (L1 == 24) ? 8 : 4, G2 = 32 / G1, R1 = L1 / G1, R2 = L2 / G2;
        const int orb = i & 1, c = i >> 1, c1 = c % L1, c2 = c / L1;
        const int lane = c1 / R1 + G1 * (c2 / R2), u = (c2 % R2) * R1 + c1 % R1;
        return (size_t)(u * 32 + lane) * 2 + orb;
    }
    if (kind == 2) {                                           // chain: lane holds 2 lxl consecutive sites
        const int lane = i / (2 * lxl), u = (i % (2 * lxl)) / 2;
        return (size_t)(u * 32 + lane) * 2 + (i & 1);
    }
    const int LX = 4 * lxl, x = i % LX, y = i / LX;
    const int lane = x / 4 + lxl * (y / ry), r = y % ry, j = x % 4;
    return (size_t)((r * 2 + j / 2) * 32 + lane) * 2 + (j & 1);
}
__global__ void k_v3_to_native(double *__restrict__ dst, const double2 *__restrict__ src, int L, int N, int lxl, int ry, int kind) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)L * N) return;
    const int l = (int)(idx / N), i = (int)(idx % N);
    const double2 v = src[idx];
    const size_t o = v3_native_index(i, lxl, ry, kind);
    dst[((size_t)l * 2) * N + o] = v.x;
    dst[((size_t)l * 2 + 1) * N + o] = v.y;
}
__global__ void k_v3_from_native(double2 *__restrict__ dst, const double *__restrict__ src, int L, int N, int lxl, int ry, int kind) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)L * N) return;
    const int l = (int)(idx / N), i = (int)(idx % N);
    const size_t o = v3_native_index(i, lxl, ry, kind);
    dst[idx] = make_double2(src[((size_t)l * 2) * N + o], src[((size_t)l * 2 + 1) * N + o]);
}
struct V3Clo { int lo[4]; };
__global__ void k_v3_expV_native(double *__restrict__ dst, const double *__restrict__ src, int L, int N, int lxl, int ry,
                                 const double2 *__restrict__ cs, const V3Clo clo, double2 *__restrict__ ctn, int kind, int ncol, int scaled) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (scaled && idx < (size_t)ncol) { const double2 q = __ldg(cs + clo.lo[idx]); ctn[idx] = make_double2(q.x, q.y / q.x); }
    if (idx >= (size_t)L * N) return;
    const int l = (int)(idx / N), i = (int)(idx % N);
    double g = 1.0;                                            // prod_c cosh_c^2 (each colour is applied twice in B)
    if (scaled) {
#pragma unroll
        for (int c = 0; c < ncol; c++) { const double ch = __ldg(cs + clo.lo[c]).x; g *= ch * ch; }
    }
    dst[(size_t)l * N + v3_native_index(i, lxl, ry, kind)] = scaled ? g * src[idx] : src[idx];
}
// per-bond engines: csn[l][e] = cs[l][map[e]], e = slot * 32 + lane
__global__ void k_v3_cs_native(double2 *__restrict__ dst, const double2 *__restrict__ cs, const int *__restrict__ map, int L, int Nh, int ne) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)L * ne) return;
    const int l = (int)(idx / ne), e = (int)(idx % ne);
    dst[idx] = __ldg(cs + (size_t)l * Nh + __ldg(map + e));
}

// Host mirror of the slot enumeration of the per-bond engines: the two sites of slot q of every lane (index q * 32 + lane).
static std::vector<int2> v3pb_slot_sites(int kind, int a, int b) {
    std::vector<int2> out;
    if (kind == 2) {
        const int R = a, N = 64 * R, NCS = 2 * R + 1;
        out.resize((size_t)NCS * 32);
        for (int lane = 0; lane < 32; lane++) {
            const int s0 = 2 * R * lane;
            int q = 0;
            for (int u = 0; u < R; u++) out[(q++) * 32 + lane] = make_int2(s0 + 2 * u, s0 + 2 * u + 1);
            for (int u = 0; u + 1 < R; u++) out[(q++) * 32 + lane] = make_int2(s0 + 2 * u + 1, s0 + 2 * u + 2);
            out[(q++) * 32 + lane] = make_int2(s0 + 2 * R - 1, (s0 + 2 * R) % N);
            out[(q++) * 32 + lane] = make_int2((s0 + N - 1) % N, s0);
        }
        return out;
    }
    const int LXL = a, RY = b, YH = 32 / LXL, LX = 4 * LXL, NCS = 9 * RY + 4;
    out.resize((size_t)NCS * 32);
    auto site = [&](int xl, int yh, int r, int j) { return LX * (RY * ((yh + YH) % YH) + r) + 4 * ((xl + LXL) % LXL) + j; };
    for (int lane = 0; lane < 32; lane++) {
        const int xl = lane % LXL, yh = lane / LXL;
        int q = 0;
        for (int r = 0; r < RY; r++) {
            out[(q++) * 32 + lane] = make_int2(site(xl, yh, r, 0), site(xl, yh, r, 1));
            out[(q++) * 32 + lane] = make_int2(site(xl, yh, r, 2), site(xl, yh, r, 3));
        }
        for (int r = 0; r < RY; r++) {
            out[(q++) * 32 + lane] = make_int2(site(xl, yh, r, 1), site(xl, yh, r, 2));
            out[(q++) * 32 + lane] = make_int2(site(xl, yh, r, 3), site(xl + 1, yh, r, 0));
            out[(q++) * 32 + lane] = make_int2(site(xl - 1, yh, r, 3), site(xl, yh, r, 0));
        }
        for (int r = 0; r < RY; r += 2)
            for (int j = 0; j < 4; j++) out[(q++) * 32 + lane] = make_int2(site(xl, yh, r, j), site(xl, yh, r + 1, j));
        for (int j = 0; j < 4; j++) out[(q++) * 32 + lane] = make_int2(site(xl, yh, RY - 1, j), site(xl, yh + 1, 0, j));
        for (int j = 0; j < 4; j++) out[(q++) * 32 + lane] = make_int2(site(xl, yh - 1, RY - 1, j), site(xl, yh, 0, j));
        for (int r = 1; r + 1 < RY; r += 2)
            for (int j = 0; j < 4; j++) out[(q++) * 32 + lane] = make_int2(site(xl, yh, r, j), site(xl, yh, r + 1, j));
    }
    return out;
}
static void v3pb_build_map(sq_fdm *f) {
    const std::vector<int2> ss = v3pb_slot_sites(f->v3_kind, f->v3_lxl, f->v3_ry);
    std::map<std::pair<int, int>, int> bond;
    for (int h = 0; h < (int)f->Nh; h++) bond[std::make_pair(std::min(f->h_nt[h].x, f->h_nt[h].y), std::max(f->h_nt[h].x, f->h_nt[h].y))] = h;
    std::vector<int> map(ss.size());
    for (size_t e = 0; e < ss.size(); e++) {
        auto it = bond.find(std::make_pair(std::min(ss[e].x, ss[e].y), std::max(ss[e].x, ss[e].y)));
        SQ_REQUIRE(it != bond.end(), "register path: a slot of the per-bond engine has no bond in the neighbour table");
        map[e] = it->second;
    }
    f->v3_ncs = (int)(ss.size() / 32);
    f->v3_csmap.alloc(map.size());
    f->v3_csmap.upload(map.data(), map.size(), f->stream);
    SQ_CUDA(cudaStreamSynchronize(f->stream));
}


void fdm_v3_to_native(sq_fdm *f, double2 *dst, const double2 *src) {
    const size_t n = (size_t)f->L * f->N;
    k_v3_to_native<<<(unsigned)((n + 255) / 256), 256, 0, f->stream>>>((double *)dst, src, (int)f->L, (int)f->N, f->v3_lxl, f->v3_ry, f->v3_kind);
    SQ_LAUNCH_CHECK();
    f->launches++;
}
void fdm_v3_from_native(sq_fdm *f, double2 *dst, const double2 *src) {
    const size_t n = (size_t)f->L * f->N;
    k_v3_from_native<<<(unsigned)((n + 255) / 256), 256, 0, f->stream>>>(dst, (const double *)src, (int)f->L, (int)f->N, f->v3_lxl, f->v3_ry, f->v3_kind);
    SQ_LAUNCH_CHECK();
    f->launches++;
}
// native-order copy of exp(-dtau V), refreshed when the operator changed
void fdm_v3_prepare_native(sq_fdm *f) {
    const size_t n = (size_t)f->L * f->N;
    if (!f->v3_expVn.p) { f->v3_expVn.alloc(n); f->v3_ctn.alloc(4); f->v3_x.alloc(n); f->v3_r.alloc(n); f->v3_expv_version = -1; }
    const int pb = fdm_v3_perbond(f) ? 1 : 0;
    if (f->v3_expv_version == f->coef_version && f->v3_native_pb == pb) return;
    V3Clo clo;
    for (int c = 0; c < 4; c++) clo.lo[c] = c < f->C ? f->clo[c] : 0;
    k_v3_expV_native<<<(unsigned)((n + 255) / 256), 256, 0, f->stream>>>(f->v3_expVn.p, f->expV.p, (int)f->L, (int)f->N, f->v3_lxl, f->v3_ry,
                                                                        f->cs.p, clo, f->v3_ctn.p, f->v3_kind, (int)f->C, pb ? 0 : 1);
    SQ_LAUNCH_CHECK();
    f->launches++;
    if (pb) {
        if (!f->v3_csmap.p) v3pb_build_map(f);
        const size_t ne = (size_t)f->v3_ncs * 32, tot = ne * f->L;
        if (f->v3_csn.n < tot) f->v3_csn.alloc(tot);
        k_v3_cs_native<<<(unsigned)((tot + 255) / 256), 256, 0, f->stream>>>(f->v3_csn.p, f->cs.p, f->v3_csmap.p, (int)f->L, (int)f->Nh, (int)ne);
        SQ_LAUNCH_CHECK();
        f->launches++;
    }
    f->v3_expv_version = f->coef_version;
    f->v3_native_pb = pb;
}
