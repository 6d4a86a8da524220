// chain.h -- "chain-in-warp" layouts for pairs of consecutive checkerboard colours.
//
// The union of two colours (a, b) of a checkerboard decomposition is a set of disjoint paths / rings that alternate
// a- and b-bonds (x-even/x-odd bonds of a square lattice form the 32-site rings along x; the honeycomb zig-zags;
// a chain's two colours its single ring).  If every ring has at most 32 a-bonds it fits in one warp: lane k keeps the
// two sites (u, v) of its a-bond in registers, the a-rotation is thread local, and the b-bond (v_k, u_next) is applied
// with two warp shuffles -- no shared-memory round trip and no block barrier for either colour.  Shared memory is only
// needed to move from one colour pair to the next (e.g. x-rings -> y-rings).
#pragma once
#include "sq_internal.h"

struct ChainLayout {
    bool ok = false;
    int a = -1, b = -1;               // colours (b = -1: single colour)
    int T = 0;                        // threads (multiple of 32)
    std::vector<int> site_u, site_v;  // per thread: the two sites of its a-bond (-1: idle lane)
    std::vector<int> bond_a, bond_b;  // internal bond indices (bond_b: the b-bond (v, u_next), -1 if none)
    std::vector<int> next, prev;      // lane of the next / previous a-bond along the ring (self if none)
    std::vector<int> has_prev;        // 1 if the previous lane applies a b-bond that ends on this lane's u
};

ChainLayout chain_build(const sq_fdm *f, int a, int b);
