// items.h -- how one walk (i-group) is cut into i-tiles ("work items") of the force pass.
// Shared by the host work-list builder (gplum_b200.cu: build_items) and the device one
// (dev_tree.cu: emit_items_kernel) so both produce the same items and the same cost keys.
#pragma once

#if defined(__CUDACC__)
#define GB_HD __host__ __device__ __forceinline__
#else
#define GB_HD inline
#endif

namespace gb {

// kernel configuration code of a tile shape (kernels.cuh: force_pass_kernel decodes it)
GB_HD int tile_cfg_of(int c) { return c == 64 ? 1 : c == 32 ? 0 : c == 16 ? 9 : c == 8 ? 10 : 11; }

// Next tile of a walk with `rem` i-particles left (rmax <= 2 builds): n = i-particles it takes,
// shape = lanes x registers capacity of the tile (64 / 32 / 16 / 8 / 4).
// Finer decompositions of a remainder (e.g. 20 -> 16 + 4) were measured slower at
// n_group_limit = 64: every extra item pays its own staging.
GB_HD void tile_next(int rem, int cap, bool split, int &n, int &shape)
{
    if (rem >= cap) { n = cap; shape = cap; return; }
    if (!split) { n = rem; shape = rem > 32 ? 64 : 32; return; }
    if (rem > 32 && rem <= 48) { n = 32; shape = 32; return; }      // 32 + a j-split tail beats a half-empty 64
    n = rem;
    shape = rem > 32 ? 64 : rem > 16 ? 32 : rem > 8 ? 16 : rem > 4 ? 8 : 4;
}

// cost model in issue slots per lane (EP-EP 18.5, EP-SP 37 per pair) + per-tile staging overhead
GB_HD double tile_cost(int n_epj, int n_spj, int shape)
{
    const double cost_j = 18.5 * n_epj + 37.0 * n_spj;
    const double cost_tiles = 90.0 * ((n_epj + 63) / 64 + (n_spj + 63) / 64) + 200.0;
    return cost_j * shape / 32.0 + cost_tiles;
}

// EP/SP split of full-width tiles (kernels.cuh: warp_force, `part`): cfg bits
constexpr int TILE_EP_ONLY = 16, TILE_SP_ONLY = 32;
GB_HD double tile_cost_ep(int n_epj, int shape) { return 18.5 * n_epj * shape / 32.0 + 90.0 * ((n_epj + 63) / 64) + 200.0; }
GB_HD double tile_cost_sp(int n_spj, int shape) { return 37.0 * n_spj * shape / 32.0 + 90.0 * ((n_spj + 63) / 64) + 200.0; }

// Tile capacity of a pass: 64 i-particles per warp is the most efficient shape (staging is amortised
// over the most pairs), and measured on 1/4- and 1/8-size shards it stays the fastest even at 0.6 waves.
// Only a pass that cannot give every fourth warp slot an item (per-call functor form, a small boundary
// set) is latency-bound on one item's serial chain: there use 32, or j-split tiles.
// n_items_at[k] = sum over walks of ceil(ni / (64 >> k)), k = 0..4.
GB_HD int tile_cap_choose(const long long n_items_at[5], long long warp_slots, int tile_cap, bool split, int rmax)
{
    int cap = rmax >= 2 ? 64 : 32;
    if (tile_cap > 0) return cap < tile_cap ? cap : tile_cap;
    if (split && rmax <= 2) {
        const long long target = warp_slots / 4;
        for (; cap > 4; cap >>= 1) {
            int k = 0;
            for (int c = 64; c > cap; c >>= 1) k++;
            if (n_items_at[k] >= target) break;
        }
    }
    return cap;
}

}  // namespace gb
