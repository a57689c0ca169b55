// let_tree.cpp -- host-side builder of the interaction lists the force pass consumes (libgplum_lists.so).
//
// The caller side of the hot path (CPU version): what FDPS does between setParticleLocalTree and
// calcForce on one rank with an open boundary and SEARCH_MODE_LONG_SYMMETRY, written from FDPS's semantics and
// pinned against the compiled reference LIST FOR LIST (tests/test_tree.py: same particle order, same groups, same
// EP and SP index lists in the same order, same SPJ bits as the tree of the unmodified FDPS compiled from the reference):
//   * root cell: bounding box of the POSITIONS (FDPS's GetMyRSearch trait does not find EPJGrav::getRSearch, so no
//     search radius enters), centre = box centre, a dimension thinner than 0.1 of the longest is pushed into one half
//     (centre -= 0.51 * extent), edge = longest extent * 1.000001     (FDPS/src/tree_for_force_impl.hpp:770-868)
//   * Morton key: 42 bits per coordinate, n = (U64)((pos - centre + half) * ((1 / edge) * 2^42)), as two 63-bit
//     words hi (levels 1..21) and lo (levels 22..42); particles sorted by (hi, lo)     (FDPS/src/key.hpp:118-226)
//   * cells: root = 0, cells 1..7 unused, children of the splitting cells of a level appended as blocks of 8 in
//     cell order; a cell splits while it holds more than n_leaf_limit particles and its level is below 42
//                                                                      (tree_for_force_utils.hpp:289-420 LinkCell)
//   * moments bottom-up: mass, centre of mass, quadrupole about it; in-box = bbox(pos), out-box = bbox(pos +- 1.1
//     r_search); size = edge * 2^-level               (utils_moment.hpp:6-60, tree.hpp:576-641,1164-1212, particle.h:119)
//   * i-groups: depth-first, the first cell on a path with <= n_group_limit particles or a leaf (utils.hpp:619-650)
//   * walk, depth-first over children 0..7: a non-empty child is opened if group.in overlaps child.out, or group.out
//     overlaps child.in, or dist^2(group.in, child.com) <= size^2 / theta^2; an unopened child becomes one
//     superparticle (its cell index), a leaf contributes all its particles        (tree_walk.hpp:545-583,706-785)
// Hence every j within 1.1*max(r_search_i, r_search_j) of any i of a group is in that group's EP list (SURVEY
// Appendix C) -- the property the neighbour detection relies on.
// Workload tooling for bench.py / tests and the host mirror of the GPU builder (dev_tree.cu); it is NOT part of
// libgplum_b200.so and no arithmetic of the force kernels lives here.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../include/gplum_b200_lists.h"

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
        double d[3];
        for (int k = 0; k < 3; k++) d[k] = (p[k] > hi[k]) ? (p[k] - hi[k]) : ((p[k] < lo[k]) ? (lo[k] - p[k]) : 0.0);
        return d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    }
};

struct Cell {
    int first, n, child, level;     // child = index of the first of 8 children, -1 for a leaf
    double mass, com[3], quad[6];   // quad order xx,yy,zz,xy,xz,yz
    double size;
    Box in, out;
};

constexpr int MAX_LEVEL = 42, LEVEL_HI = 21;       // TREE_LEVEL_LIMIT, KEY_LEVEL_MAX_HI (FDPS/src/ps_defs.hpp:206-208)

struct Key { uint64_t hi, lo; };

struct Tree {
    int n = 0, n_leaf = 8, n_group = 64;
    double theta = 0.5;
    std::vector<Key> key;
    std::vector<int> order;               // sorted -> original index
    std::vector<double> pos, mass, rsrch; // sorted order; rsrch = 1.1 * r_search
    std::vector<Cell> cell;
    std::vector<int> group;               // cell indices
    // outputs
    std::vector<int> adr_epj, adr_spj, n_epj, n_spj, epi_off, ni;
    std::vector<long long> epj_disp, spj_disp;
    std::vector<double> r_out_s, r_search_s;
} T;

inline uint64_t spread3(uint64_t x)       // FDPS/src/key.hpp:137-144 (21 bits -> every third bit)
{
    x &= 0x1fffff;
    x = (x | x << 32) & 0xffff00000000ffffULL;
    x = (x | x << 16) & 0x00ff0000ff0000ffULL;
    x = (x | x << 8) & 0xf00f00f00f00f00fULL;
    x = (x | x << 4) & 0x30c30c30c30c30c3ULL;
    x = (x | x << 2) & 0x9249249249249249ULL;
    return x;
}

// octant of a key on level `lev` (1 .. 42): FDPS/src/key.hpp:213-226 getCellID
inline int key_digit(const Key &k, int lev)
{
    return lev <= LEVEL_HI ? (int)((k.hi >> ((LEVEL_HI - lev) * 3)) & 7) : (int)((k.lo >> ((LEVEL_HI - (lev - LEVEL_HI)) * 3)) & 7);
}

void build_cells()
{
    // level by level; the cells of a level that split get their blocks of 8 children in cell order
    size_t lvl_begin = 0, lvl_end = 1;
    T.cell.resize(8);                                   // root + 7 unused cells (LinkCell: tc_array[1..7])
    for (int k = 1; k < 8; k++) { Cell d; memset(&d, 0, sizeof(d)); d.child = -1; T.cell[k] = d; }
    for (int lev = 0; lev < MAX_LEVEL; lev++) {
        const size_t next_begin = T.cell.size();
        for (size_t c = lvl_begin; c < lvl_end; c++) {
            const Cell cur = T.cell[c];
            if (cur.n <= T.n_leaf) continue;
            const int base = (int)T.cell.size();
            T.cell[c].child = base;
            int p = cur.first;
            const int end = cur.first + cur.n;
            for (int o = 0; o < 8; o++) {
                Cell ch;
                memset(&ch, 0, sizeof(ch));
                ch.first = p; ch.level = lev + 1; ch.child = -1; ch.size = cur.size * 0.5;
                // particles of octant o are contiguous: find the end by binary search on the key digit
                int lo = p, hi = end;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (key_digit(T.key[mid], lev + 1) <= o) lo = mid + 1; else hi = mid;
                }
                ch.n = lo - p;
                p = lo;
                T.cell.push_back(ch);
            }
        }
        if (T.cell.size() == next_begin) break;
        lvl_begin = next_begin;
        lvl_end = T.cell.size();
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
            { const double inv_m = 1.0 / x.mass;   // PS::F64vec / F64 multiplies by the reciprocal (FDPS/src/vector3.hpp:217-220)
              for (int k = 0; k < 3; k++) x.com[k] = (x.mass != 0.0) ? x.com[k] * inv_m : 0.0; }
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
            { const double inv_m = 1.0 / x.mass;   // PS::F64vec / F64 multiplies by the reciprocal (FDPS/src/vector3.hpp:217-220)
              for (int k = 0; k < 3; k++) x.com[k] = (x.mass != 0.0) ? x.com[k] * inv_m : 0.0; }
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

// depth-first over children 0..7, the order of FDPS's recursion: EP indices come out ascending, SP in encounter order
void walk_cell(const Cell &g, int ci, double inv_theta2, std::vector<int> &ep, std::vector<int> &sp)
{
    const Cell &c = T.cell[ci];
    if (c.child < 0) {            // leaf: every particle
        for (int i = c.first; i < c.first + c.n; i++) ep.push_back(i);
        return;
    }
    for (int o = 0; o < 8; o++) {
        const int cj = c.child + o;
        const Cell &ch = T.cell[cj];
        if (ch.n == 0) continue;
        const bool open = g.in.overlaps(ch.out) || g.out.overlaps(ch.in) ||
                          g.in.dist2(ch.com) <= ch.size * ch.size * inv_theta2;
        if (open) walk_cell(g, cj, inv_theta2, ep, sp); else sp.push_back(cj);
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
    if (n <= 0 || !pos || !mass || !r_out || !r_search || theta <= 0.0) return 2;   /* GPLUM_B200_ERR_ARG */
    T = Tree();
    T.n = n; T.theta = theta; T.n_leaf = n_leaf_limit; T.n_group = n_group_limit;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], pos[3 * i + k]); hi[k] = std::max(hi[k], pos[3 * i + k]); }
    double length = 0, cen[3], len_dim[3];
    for (int k = 0; k < 3; k++) { cen[k] = (hi[k] + lo[k]) * 0.5; len_dim[k] = hi[k] - lo[k]; length = std::max(length, len_dim[k]); }
    for (int k = 0; k < 3; k++) if (len_dim[k] < 0.1 * length) cen[k] -= len_dim[k] * 0.51;
    length *= 1.000001;
    if (!(length > 0)) length = 1.0;                           // a single point: any cube will do
    const double hlen = length * 0.5;
    const double nfactor = (1.0 / (hlen * 2.0)) * (double)(1ULL << MAX_LEVEL);
    const uint64_t nmax = (1ULL << MAX_LEVEL) - 1;
    struct KO { Key k; int i; };
    std::vector<KO> ko(n);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        uint64_t c[3];
        for (int k = 0; k < 3; k++) {
            const double f = (pos[3 * i + k] - cen[k] + hlen) * nfactor;
            c[k] = f < 0.0 ? 0 : (f >= 9.2e18 ? nmax : (uint64_t)f);
            if (c[k] > nmax) c[k] = nmax;
        }
        Key key;
        key.hi = spread3(c[0] >> LEVEL_HI) << 2 | spread3(c[1] >> LEVEL_HI) << 1 | spread3(c[2] >> LEVEL_HI);
        key.lo = spread3(c[0] & 0x1fffff) << 2 | spread3(c[1] & 0x1fffff) << 1 | spread3(c[2] & 0x1fffff);
        ko[i] = {key, i};
    }
    std::sort(ko.begin(), ko.end(), [](const KO &a, const KO &b) {
        return a.k.hi != b.k.hi ? a.k.hi < b.k.hi : (a.k.lo != b.k.lo ? a.k.lo < b.k.lo : a.i < b.i);
    });
    T.key.resize(n); T.order.resize(n); T.pos.resize(3 * (size_t)n); T.mass.resize(n); T.rsrch.resize(n);
    T.r_out_s.resize(n); T.r_search_s.resize(n);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        const int o = ko[i].i;
        T.key[i] = ko[i].k; T.order[i] = o;
        for (int k = 0; k < 3; k++) T.pos[3 * i + k] = pos[3 * o + k];
        T.mass[i] = mass[o]; T.rsrch[i] = 1.1 * r_search[o];
        T.r_out_s[i] = r_out[o]; T.r_search_s[i] = r_search[o];
    }
    Cell root;
    memset(&root, 0, sizeof(root));
    root.first = 0; root.n = n; root.child = -1; root.level = 0; root.size = hlen * 2.0;
    T.cell.reserve((size_t)n + 64);
    T.cell.push_back(root);
    build_cells();
    moments();
    make_groups(0);
    const int ng = (int)T.group.size();
    std::vector<std::vector<int>> ep(ng), sp(ng);
    const double inv_theta2 = 1.0 / (theta * theta);
#pragma omp parallel for schedule(dynamic, 4)
    for (int g = 0; g < ng; g++) walk_cell(T.cell[T.group[g]], 0, inv_theta2, ep[g], sp[g]);
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
    if (T.n == 0) return 3;                                   /* GPLUM_B200_ERR_STATE */
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
