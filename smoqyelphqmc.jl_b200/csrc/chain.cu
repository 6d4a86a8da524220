// chain.cu -- host-side construction of the chain-in-warp layouts (see chain.h)
#include "chain.h"

#include <algorithm>

ChainLayout chain_build(const sq_fdm *f, int a, int b) {
    ChainLayout out;
    out.a = a; out.b = b;
    const int N = (int)f->N;
    if (a < 0 || a >= f->C) return out;
    const int alo = f->clo[a], ahi = f->chi[a];
    if (2 * (ahi - alo) != N) return out;                          // colour a must touch every site
    std::vector<int> abond(N, -1), aother(N, -1), bbond(N, -1), bother(N, -1);
    for (int h = alo; h < ahi; h++) {
        int2 ij = f->h_nt[h];
        abond[ij.x] = abond[ij.y] = h;
        aother[ij.x] = ij.y; aother[ij.y] = ij.x;
    }
    if (b >= 0 && b < f->C) {
        for (int h = f->clo[b]; h < f->chi[b]; h++) {
            int2 ij = f->h_nt[h];
            bbond[ij.x] = bbond[ij.y] = h;
            bother[ij.x] = ij.y; bother[ij.y] = ij.x;
        }
    }
    struct Elem { int u, v, ba, bb; };
    std::vector<std::vector<Elem>> chains;
    std::vector<char> cyc;
    std::vector<char> seen(f->Nh, 0);
    auto walk = [&](int u0, int v0, bool is_cycle) {
        std::vector<Elem> ch;
        int u = u0, v = v0;
        while (true) {
            int e = abond[u];
            seen[e] = 1;
            int w = bother[v];
            bool closes = (w >= 0 && abond[w] == abond[u0] && w == u0);
            ch.push_back({u, v, e, (w >= 0) ? bbond[v] : -1});
            if (w < 0 || closes) break;
            if (seen[abond[w]]) { ch.back().bb = -1; break; }     // defensive: malformed input
            u = w;
            v = aother[w];
        }
        chains.push_back(ch);
        cyc.push_back(is_cycle ? 1 : 0);
    };
    // open paths: start from an a-bond with an endpoint that has no b-bond
    for (int h = alo; h < ahi; h++) {
        if (seen[h]) continue;
        int2 ij = f->h_nt[h];
        if (bbond[ij.x] < 0) walk(ij.x, ij.y, false);
        else if (bbond[ij.y] < 0) walk(ij.y, ij.x, false);
    }
    for (int h = alo; h < ahi; h++) {
        if (seen[h]) continue;
        int2 ij = f->h_nt[h];
        walk(ij.x, ij.y, true);
    }
    for (auto &ch : chains) if ((int)ch.size() > 32) return out;   // a ring does not fit in a warp
    // first-fit decreasing packing of chains into warps
    std::vector<int> order(chains.size());
    for (size_t k = 0; k < order.size(); k++) order[k] = (int)k;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return chains[x].size() > chains[y].size(); });
    std::vector<int> fill;                                          // lanes used per warp
    std::vector<std::pair<int, int>> place(chains.size());          // (warp, first lane)
    for (int c : order) {
        int len = (int)chains[c].size(), wsel = -1;
        for (size_t w = 0; w < fill.size(); w++) if (fill[w] + len <= 32) { wsel = (int)w; break; }
        if (wsel < 0) { fill.push_back(0); wsel = (int)fill.size() - 1; }
        place[c] = {wsel, fill[wsel]};
        fill[wsel] += len;
    }
    out.T = (int)fill.size() * 32;
    if (out.T > 1024 || out.T == 0) return out;
    out.site_u.assign(out.T, -1); out.site_v.assign(out.T, -1);
    out.bond_a.assign(out.T, -1); out.bond_b.assign(out.T, -1);
    out.next.resize(out.T); out.prev.resize(out.T); out.has_prev.assign(out.T, 0);
    for (int t = 0; t < out.T; t++) out.next[t] = out.prev[t] = t & 31;
    for (size_t c = 0; c < chains.size(); c++) {
        auto &ch = chains[c];
        int w = place[c].first, l0 = place[c].second, len = (int)ch.size();
        for (int k = 0; k < len; k++) {
            int t = w * 32 + l0 + k;
            out.site_u[t] = ch[k].u; out.site_v[t] = ch[k].v; out.bond_a[t] = ch[k].ba; out.bond_b[t] = ch[k].bb;
            bool has_next = (ch[k].bb >= 0);
            int kn = (k + 1 < len) ? k + 1 : (cyc[c] ? 0 : k);
            if (has_next && kn != k) {
                out.next[t] = l0 + kn;
                int tn = w * 32 + l0 + kn;
                out.prev[tn] = l0 + k;
                out.has_prev[tn] = 1;
            } else if (has_next && kn == k) {
                // ring of one a-bond closed by a b-bond on the same two sites: the b-bond couples (v, u) of this lane
                out.next[t] = l0 + k;
                out.prev[t] = l0 + k;
                out.has_prev[t] = 1;
            }
        }
    }
    out.ok = true;
    return out;
}
