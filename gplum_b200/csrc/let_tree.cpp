// let_tree.cpp -- host-side builder of the interaction lists the force pass consumes.
//
// This is the caller side of the hot path (CPU version): what FDPS does between
// setParticleLocalTree and calcForce -- Morton sort, octree, monopole/quadrupole moments,
// i-group construction and the per-group tree walk -- re-implemented from the published
// semantics, single rank, open boundary, SEARCH_MODE_LONG_SYMMETRY:
//   * a cell is a leaf when it holds <= n_leaf_limit particles            (FDPS/src/tree.hpp isLeaf)
//   * i-groups are the shallowest cells with <= n_group_limit particles   (tree_for_force_utils.hpp:619-650)
//   * cell boxes: vertex_in = bbox(pos), vertex_out = bbox(pos +- 1.1*r_search)
//                                                                        (tree.hpp:1186-1205, particle.h:119-125)
//   * moments: mass, centre of mass, raw second moment about it           (tree.hpp:576-641)
//   * walk: a child cell is opened if group.in overlaps cell.out, or group.out overlaps
//     cell.in, or dist^2(group.in, cell.com) <= (cell.size/theta)^2; an unopened non-empty
//     cell becomes one superparticle; leaves contribute all their particles
//                                                                        (tree_walk.hpp:545-583,706-785)
// Hence every j within 1.1*max(r_search_i, r_search_j) of any i of a group is in that group's
// EP list -- the property the neighbour detection relies on.
// It is used by bench.py / tests to produce workloads without the reference, and as the
// list builder of the stand-alone force call.  No arithmetic of the force kernels lives here.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../include/gplum_b200.h"

namespace {

struct Box {
    double lo[3], hi[3];
    void init() { for (int k = 0; k < 3; k++) { lo[k] = 1e300; hi[k] = -1e300; } }
    void merge(const double *p, double r) { for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], p[k] - r); hi[k] = std::max(hi[k], p[k] + r); } }
    void merge(const Box &b) { for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); } }
    bool overlaps(const Box &b) const {
        for (int k = 0; k < 3; k++) if (hi[k] < b.lo[k] || b.hi[k] < lo[k]) return false;
        return true;
    }
    double dist2(const double *p) const {
        double d2 = 0;
        for (int k = 0; k < 3; k++) { const double d = std::max(0.0, std::max(lo[k] - p[k], p[k] - hi[k])); d2 += d * d; }
        return d2;
    }
};

struct Cell {
    int first, n, child, level;     // child = index of the first of 8 children, -1 for a leaf
    double mass, com[3], quad[6];   // quad order xx,yy,zz,xy,xz,yz
    double size;
    Box in, out;
};

struct Tree {
    int n = 0, n_leaf = 8, n_group = 64;
    double theta = 0.5;
    std::vector<uint64_t> key;
    std::vector<int> order;               // sorted -> original index
    std::vector<double> pos, mass, rsrch; // sorted order; rsrch = 1.1 * r_search
    std::vector<Cell> cell;
    std::vector<int> group;               // cell indices
    // outputs
    std::vector<int> adr_epj, adr_spj, n_epj, n_spj, epi_off, ni;
    std::vector<long long> epj_disp, spj_disp;
    std::vector<double> r_out_s, r_search_s;
} T;

inline uint64_t spread3(uint64_t x)
{
    x &= 0x1fffff;
    x = (x | x << 32) & 0x1f00000000ffffULL;
    x = (x | x << 16) & 0x1f0000ff0000ffULL;
    x = (x | x << 8) & 0x100f00f00f00f00fULL;
    x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
    x = (x | x << 2) & 0x1249249249249249ULL;
    return x;
}

constexpr int MAX_LEVEL = 21;

void build_cells(int ci)
{
    // iterative, breadth-first: children of cell ci are appended as a block of 8
    std::vector<int> todo{ci};
    size_t head = 0;
    while (head < todo.size()) {
        const int c = todo[head++];
        Cell cur = T.cell[c];
        if (cur.n <= T.n_leaf || cur.level >= MAX_LEVEL) continue;
        const int base = (int)T.cell.size();
        T.cell[c].child = base;
        const int shift = 3 * (MAX_LEVEL - 1 - cur.level);
        int p = cur.first;
        const int end = cur.first + cur.n;
        for (int o = 0; o < 8; o++) {
            Cell ch;
            ch.first = p; ch.level = cur.level + 1; ch.child = -1; ch.size = cur.size * 0.5;
            // particles of octant o are contiguous: find the end by binary search on the key digit
            int lo = p, hi = end;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if ((int)((T.key[mid] >> shift) & 7) <= o) lo = mid + 1; else hi = mid;
            }
            ch.n = lo - p;
            p = lo;
            T.cell.push_back(ch);
        }
        for (int o = 0; o < 8; o++) if (T.cell[base + o].n > 0) todo.push_back(base + o);
    }
}

void moments()
{
    // children always have larger indices than their parent: sweep backwards
    for (int c = (int)T.cell.size() - 1; c >= 0; c--) {
        Cell &x = T.cell[c];
        x.mass = 0; x.com[0] = x.com[1] = x.com[2] = 0;
        for (int k = 0; k < 6; k++) x.quad[k] = 0;
        x.in.init(); x.out.init();
        if (x.n == 0) continue;
        if (x.child < 0) {
            for (int i = x.first; i < x.first + x.n; i++) {
                const double *p = &T.pos[3 * i];
                x.mass += T.mass[i];
                for (int k = 0; k < 3; k++) x.com[k] += T.mass[i] * p[k];
                x.in.merge(p, 0.0); x.out.merge(p, T.rsrch[i]);
            }
            for (int k = 0; k < 3; k++) x.com[k] = (x.mass != 0.0) ? x.com[k] / x.mass : 0.0;
            for (int i = x.first; i < x.first + x.n; i++) {
                const double *p = &T.pos[3 * i];
                const double d[3] = {p[0] - x.com[0], p[1] - x.com[1], p[2] - x.com[2]}, m = T.mass[i];
                x.quad[0] += m * d[0] * d[0]; x.quad[1] += m * d[1] * d[1]; x.quad[2] += m * d[2] * d[2];
                x.quad[3] += m * d[0] * d[1]; x.quad[4] += m * d[0] * d[2]; x.quad[5] += m * d[1] * d[2];
            }
        } else {
            for (int o = 0; o < 8; o++) {
                const Cell &ch = T.cell[x.child + o];
                if (ch.n == 0) continue;
                x.mass += ch.mass;
                for (int k = 0; k < 3; k++) x.com[k] += ch.mass * ch.com[k];
                x.in.merge(ch.in); x.out.merge(ch.out);
            }
            for (int k = 0; k < 3; k++) x.com[k] = (x.mass != 0.0) ? x.com[k] / x.mass : 0.0;
            for (int o = 0; o < 8; o++) {
                const Cell &ch = T.cell[x.child + o];
                if (ch.n == 0) continue;
                const double d[3] = {ch.com[0] - x.com[0], ch.com[1] - x.com[1], ch.com[2] - x.com[2]}, m = ch.mass;
                x.quad[0] += m * d[0] * d[0] + ch.quad[0]; x.quad[1] += m * d[1] * d[1] + ch.quad[1];
                x.quad[2] += m * d[2] * d[2] + ch.quad[2]; x.quad[3] += m * d[0] * d[1] + ch.quad[3];
                x.quad[4] += m * d[0] * d[2] + ch.quad[4]; x.quad[5] += m * d[1] * d[2] + ch.quad[5];
            }
        }
    }
}

void make_groups(int c)
{
    std::vector<int> st{c};
    while (!st.empty()) {
        const int x = st.back(); st.pop_back();
        const Cell &cl = T.cell[x];
        if (cl.n == 0) continue;
        if (cl.n <= T.n_group || cl.child < 0) { T.group.push_back(x); continue; }
        for (int o = 7; o >= 0; o--) st.push_back(cl.child + o);
    }
}

void walk_group(const Cell &g, std::vector<int> &ep, std::vector<int> &sp)
{
    const double inv_theta2 = 1.0 / (T.theta * T.theta);
    int st[512];
    int top = 0;
    st[top++] = 0;
    while (top > 0) {
        const Cell &c = T.cell[st[--top]];
        if (c.child < 0) {            // leaf: every particle
            for (int i = c.first; i < c.first + c.n; i++) ep.push_back(i);
            continue;
        }
        for (int o = 7; o >= 0; o--) {
            const int ci = c.child + o;
            const Cell &ch = T.cell[ci];
            if (ch.n == 0) continue;
            const bool open = g.in.overlaps(ch.out) || g.out.overlaps(ch.in) ||
                              g.in.dist2(ch.com) <= ch.size * ch.size * inv_theta2;
            if (open) st[top++] = ci; else sp.push_back(ci);
        }
    }
}

}  // namespace

extern "C" {

// Build tree + groups + lists for n particles.  sizes[0..7] = n_walk, n_epi(=n), n_adr_epj,
// n_adr_spj, n_epj_all(=n), n_spj_all(=n_cells), n_interaction_epep, n_interaction_epsp.
int gplum_b200_tree_build(int n, const double *pos, const double *mass, const double *r_out,
                          const double *r_search, double theta, int n_leaf_limit, int n_group_limit,
                          long long *sizes)
{
    if (n <= 0 || !pos || !mass || !r_out || !r_search || theta <= 0.0) return GPLUM_B200_ERR_ARG;
    T = Tree();
    T.n = n; T.theta = theta; T.n_leaf = n_leaf_limit; T.n_group = n_group_limit;
    // root cube: bbox of pos +- 1.1*r_search; the cube is centred on it except that a dimension
    // much thinner than the cube (a disk's z) is pushed wholly into one half, so the first levels
    // do not cut through the mid-plane (FDPS/src/tree_for_force_impl.hpp:846-866)
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) {
            const double r = 1.1 * r_search[i] * 1.000001;
            lo[k] = std::min(lo[k], pos[3 * i + k] - r); hi[k] = std::max(hi[k], pos[3 * i + k] + r);
        }
    double full = 0, cen[3];
    for (int k = 0; k < 3; k++) { cen[k] = 0.5 * (lo[k] + hi[k]); full = std::max(full, hi[k] - lo[k]); }
    for (int k = 0; k < 3; k++) if (hi[k] - lo[k] < 0.1 * full) cen[k] -= (hi[k] - lo[k]) * 0.51;
    double half = 0.5 * full * 1.000001;
    if (half <= 0) half = 1.0;
    const double len = 2.0 * half, inv = (double)(1u << MAX_LEVEL) / len;
    std::vector<std::pair<uint64_t, int>> ko(n);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        uint64_t c[3];
        for (int k = 0; k < 3; k++) {
            double f = (pos[3 * i + k] - (cen[k] - half)) * inv;
            f = std::min(std::max(f, 0.0), (double)((1u << MAX_LEVEL) - 1));
            c[k] = (uint64_t)f;
        }
        ko[i] = {spread3(c[0]) << 2 | spread3(c[1]) << 1 | spread3(c[2]), i};
    }
    std::sort(ko.begin(), ko.end());
    T.key.resize(n); T.order.resize(n); T.pos.resize(3 * (size_t)n); T.mass.resize(n); T.rsrch.resize(n);
    T.r_out_s.resize(n); T.r_search_s.resize(n);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        const int o = ko[i].second;
        T.key[i] = ko[i].first; T.order[i] = o;
        for (int k = 0; k < 3; k++) T.pos[3 * i + k] = pos[3 * o + k];
        T.mass[i] = mass[o]; T.rsrch[i] = 1.1 * r_search[o];
        T.r_out_s[i] = r_out[o]; T.r_search_s[i] = r_search[o];
    }
    Cell root;
    root.first = 0; root.n = n; root.child = -1; root.level = 0; root.size = len;
    T.cell.reserve((size_t)n);
    T.cell.push_back(root);
    build_cells(0);
    moments();
    make_groups(0);
    const int ng = (int)T.group.size();
    // groups in Morton order of their first particle (FDPS's ipg_ order)
    std::sort(T.group.begin(), T.group.end(), [](int a, int b) { return T.cell[a].first < T.cell[b].first; });
    std::vector<std::vector<int>> ep(ng), sp(ng);
#pragma omp parallel for schedule(dynamic, 4)
    for (int g = 0; g < ng; g++) walk_group(T.cell[T.group[g]], ep[g], sp[g]);
    T.epi_off.resize(ng); T.ni.resize(ng); T.n_epj.resize(ng); T.n_spj.resize(ng);
    T.epj_disp.resize(ng); T.spj_disp.resize(ng);
    long long ne = 0, ns = 0, iee = 0, ies = 0;
    for (int g = 0; g < ng; g++) {
        const Cell &c = T.cell[T.group[g]];
        T.epi_off[g] = c.first; T.ni[g] = c.n;
        T.epj_disp[g] = ne; T.spj_disp[g] = ns;
        T.n_epj[g] = (int)ep[g].size(); T.n_spj[g] = (int)sp[g].size();
        ne += T.n_epj[g]; ns += T.n_spj[g];
        iee += (long long)c.n * T.n_epj[g]; ies += (long long)c.n * T.n_spj[g];
    }
    T.adr_epj.resize(ne); T.adr_spj.resize(ns);
#pragma omp parallel for schedule(static)
    for (int g = 0; g < ng; g++) {
        std::copy(ep[g].begin(), ep[g].end(), T.adr_epj.begin() + T.epj_disp[g]);
        std::copy(sp[g].begin(), sp[g].end(), T.adr_spj.begin() + T.spj_disp[g]);
    }
    if (sizes) {
        sizes[0] = ng; sizes[1] = n; sizes[2] = ne; sizes[3] = ns; sizes[4] = n; sizes[5] = (long long)T.cell.size();
        sizes[6] = iee; sizes[7] = ies;
    }
    return 0;
}

// Copy out the result of the last build.  epi / epj_all are written in the reference's AoS
// layouts (EPIGrav 48 B, EPJGrav 112 B: id_local = original index, myrank = rank, id = original
// index, vel/acc_d = 0), spj_all as MySPJQuadrupole (80 B) or MySPJMonopole (32 B).
int gplum_b200_tree_copy(void *epi_, int *epi_off, int *ni, int *adr_epj, long long *epj_disp, int *n_epj,
                         int *adr_spj, long long *spj_disp, int *n_spj, void *epj_all_, void *spj_all_,
                         int quad, int rank, int *sorted_to_original)
{
    if (T.n == 0) return GPLUM_B200_ERR_STATE;
    struct Epi { int id_local, myrank; double pos[3]; double r_out, r_search; };
    struct Epj { int id_local, myrank; double pos[3]; double r_out, r_search; long long id; double mass; double vel[3]; double acc_d[3]; };
    struct SpjQ { double mass; double pos[3]; double quad[6]; };
    struct SpjM { double mass; double pos[3]; };
    static_assert(sizeof(Epi) == 48 && sizeof(Epj) == 112 && sizeof(SpjQ) == 80 && sizeof(SpjM) == 32, "layout");
    Epi *epi = (Epi *)epi_; Epj *epj = (Epj *)epj_all_;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < T.n; i++) {
        Epj j;
        memset(&j, 0, sizeof(j));
        j.id_local = T.order[i]; j.myrank = rank; j.id = T.order[i];
        for (int k = 0; k < 3; k++) j.pos[k] = T.pos[3 * i + k];
        j.r_out = T.r_out_s[i]; j.r_search = T.r_search_s[i]; j.mass = T.mass[i];
        if (epj) epj[i] = j;
        if (epi) { Epi e; e.id_local = j.id_local; e.myrank = rank; for (int k = 0; k < 3; k++) e.pos[k] = j.pos[k]; e.r_out = j.r_out; e.r_search = j.r_search; epi[i] = e; }
        if (sorted_to_original) sorted_to_original[i] = T.order[i];
    }
    const int nc = (int)T.cell.size();
    if (spj_all_) {
        for (int c = 0; c < nc; c++) {
            const Cell &x = T.cell[c];
            if (quad) { SpjQ s; s.mass = x.mass; for (int k = 0; k < 3; k++) s.pos[k] = x.com[k]; for (int k = 0; k < 6; k++) s.quad[k] = x.quad[k]; ((SpjQ *)spj_all_)[c] = s; }
            else { SpjM s; s.mass = x.mass; for (int k = 0; k < 3; k++) s.pos[k] = x.com[k]; ((SpjM *)spj_all_)[c] = s; }
        }
    }
    const size_t ng = T.group.size();
    if (epi_off) memcpy(epi_off, T.epi_off.data(), ng * 4);
    if (ni) memcpy(ni, T.ni.data(), ng * 4);
    if (n_epj) memcpy(n_epj, T.n_epj.data(), ng * 4);
    if (n_spj) memcpy(n_spj, T.n_spj.data(), ng * 4);
    if (epj_disp) memcpy(epj_disp, T.epj_disp.data(), ng * 8);
    if (spj_disp) memcpy(spj_disp, T.spj_disp.data(), ng * 8);
    if (adr_epj) memcpy(adr_epj, T.adr_epj.data(), T.adr_epj.size() * 4);
    if (adr_spj) memcpy(adr_spj, T.adr_spj.data(), T.adr_spj.size() * 4);
    return 0;
}

void gplum_b200_tree_free(void) { T = Tree(); }

}  // extern "C"
