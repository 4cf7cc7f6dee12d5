// ORACLE (test infrastructure): mwatershed.agglom restated -- the mutex watershed of Wolf et al. 2018 as the Rust package
// `mwatershed` (PyPI, unpinned in the reference's pyproject.toml:35) exposes it to post/mws.py:52-57.
// PARITY UNPINNED: the package is neither vendored in /root/reference nor installed here; restated from the published
// algorithm.  Recalled details behind switches:
//   * every (offset c, voxel p) with p + offset_c inside the volume is an edge of weight affs[c][p]; with `strides`, only
//     voxels p whose coordinates are multiples of strides[c] contribute offset c (randomized_strides draws an unseeded
//     random subset instead: not reproducible, not restated);
//   * w > 0 attractive, w < 0 repulsive (zero_is_repulsive: where w == 0 goes), NaN edges are skipped;
//   * edges are visited by descending |w|.  The Rust sort is unstable, so the order among equal |w| is not defined
//     upstream: DECLARED DEVIATION D4 -- ties are visited by ascending (channel, raveled voxel index), here and in the
//     CUDA path;
//   * attractive: union unless a mutex separates the two clusters; repulsive: mutex unless already one cluster;
//   * every voxel ends up labelled.  Label = 1 + the smallest raveled voxel index of its cluster (the partition is what
//     is compared; any labelling is a permutation of this one).
#include <stdint.h>
#include <math.h>

#include <algorithm>
#include <numeric>
#include <vector>

namespace {

struct Uf {
    std::vector<uint32_t> parent, rank_;
    explicit Uf(size_t n) : parent(n), rank_(n, 0) { std::iota(parent.begin(), parent.end(), 0u); }
    uint32_t find(uint32_t x) {
        while (parent[x] != x) {
            parent[x] = parent[parent[x]];
            x = parent[x];
        }
        return x;
    }
};

// clusters separated by a mutex share the id of the repulsive edge that created it (sorted id lists per root)
bool share(const std::vector<uint64_t> &a, const std::vector<uint64_t> &b) {
    size_t i = 0, j = 0;
    while (i < a.size() && j < b.size()) {
        if (a[i] == b[j]) return true;
        if (a[i] < b[j])
            i++;
        else
            j++;
    }
    return false;
}

}  // namespace

extern "C" int64_t mws_agglom(const double *affs, int C, const int64_t *shape, const int64_t *offsets, const int64_t *strides,
                              int zero_is_repulsive, uint64_t *labels_out, int64_t *counters_out) {
    const int64_t Z = shape[0], Y = shape[1], X = shape[2], V = Z * Y * X;
    if (V <= 0 || V >= (1LL << 32)) return -1;
    std::vector<uint64_t> order;   // edge id = c * V + p
    for (int c = 0; c < C; c++) {
        const int64_t oz = offsets[3 * c], oy = offsets[3 * c + 1], ox = offsets[3 * c + 2];
        const int64_t sz = strides ? strides[3 * c] : 1, sy = strides ? strides[3 * c + 1] : 1, sx = strides ? strides[3 * c + 2] : 1;
        for (int64_t z = 0; z < Z; z++) {
            if (z + oz < 0 || z + oz >= Z || z % sz) continue;
            for (int64_t y = 0; y < Y; y++) {
                if (y + oy < 0 || y + oy >= Y || y % sy) continue;
                for (int64_t x = 0; x < X; x++) {
                    if (x + ox < 0 || x + ox >= X || x % sx) continue;
                    const int64_t p = (z * Y + y) * X + x;
                    if (isnan(affs[(int64_t)c * V + p])) continue;
                    order.push_back((uint64_t)c * (uint64_t)V + (uint64_t)p);
                }
            }
        }
    }
    std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) {
        const double wa = fabs(affs[a]), wb = fabs(affs[b]);
        if (wa != wb) return wa > wb;
        return a < b;   // D4: (channel, voxel) ascending among equal |w|
    });
    Uf uf((size_t)V);
    std::vector<std::vector<uint64_t>> mutexes((size_t)V);
    int64_t n_merge = 0, n_mutex = 0, n_blocked = 0;
    for (uint64_t e : order) {
        const int c = (int)(e / (uint64_t)V);
        const int64_t p = (int64_t)(e % (uint64_t)V);
        const int64_t q = p + (offsets[3 * c] * Y + offsets[3 * c + 1]) * X + offsets[3 * c + 2];
        const double w = affs[e];
        uint32_t ra = uf.find((uint32_t)p), rb = uf.find((uint32_t)q);
        if (ra == rb) continue;
        const bool attractive = w > 0.0 || (w == 0.0 && !zero_is_repulsive);
        if (attractive) {
            if (share(mutexes[ra], mutexes[rb])) {
                n_blocked++;
                continue;
            }
            if (uf.rank_[ra] < uf.rank_[rb]) std::swap(ra, rb);
            uf.parent[rb] = ra;
            if (uf.rank_[ra] == uf.rank_[rb]) uf.rank_[ra]++;
            if (!mutexes[rb].empty()) {
                std::vector<uint64_t> merged;
                merged.reserve(mutexes[ra].size() + mutexes[rb].size());
                std::set_union(mutexes[ra].begin(), mutexes[ra].end(), mutexes[rb].begin(), mutexes[rb].end(), std::back_inserter(merged));
                mutexes[ra].swap(merged);
                std::vector<uint64_t>().swap(mutexes[rb]);
            }
            n_merge++;
        } else {
            // ids arrive in no particular numeric order: keep the lists sorted
            mutexes[ra].insert(std::upper_bound(mutexes[ra].begin(), mutexes[ra].end(), e), e);
            mutexes[rb].insert(std::upper_bound(mutexes[rb].begin(), mutexes[rb].end(), e), e);
            n_mutex++;
        }
    }
    std::vector<uint32_t> first((size_t)V, 0xFFFFFFFFu);
    for (int64_t p = 0; p < V; p++) {
        const uint32_t r = uf.find((uint32_t)p);
        if (first[r] == 0xFFFFFFFFu) first[r] = (uint32_t)p;
    }
    for (int64_t p = 0; p < V; p++) labels_out[p] = (uint64_t)first[uf.find((uint32_t)p)] + 1u;
    if (counters_out) {
        counters_out[0] = (int64_t)order.size();
        counters_out[1] = n_merge;
        counters_out[2] = n_mutex;
        counters_out[3] = n_blocked;
    }
    return 0;
}
