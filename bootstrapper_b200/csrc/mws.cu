// Mutex watershed on the device: the `bs segment --mws` fragmenter (sm_100a).
//
// Replaces mwatershed.agglom as post/mws.py:52-57 calls it (mwatershed_from_affinities: shift = noise + bias, weights =
// affs + shift in float64; every (offset c, voxel p) with p + offset_c inside the volume -- and p on the stride lattice of c
// -- is an edge; w > 0 attractive, w <= 0 repulsive; edges visited by descending |w|; attractive: union unless a mutex
// separates the clusters, repulsive: mutex unless already one cluster).  Declared tie rule D4: equal |w| are
// visited by ascending (channel, raveled voxel) -- the upstream Rust sort is unstable, so no order is defined there.
//
// The sequential pass is reproduced EXACTLY by rounds of locally safe edges.  All edges are ranked once (stable LSD radix
// sort of the float64 |w| bits, descending); rank = position in the sequential order.  In a round, with bestA[c] = the
// smallest rank of a live attractive edge at cluster c:
//   * a repulsive edge (A, B) of rank r is executed when r < bestA[A] and r < bestA[B]: no pending attractive edge of higher
//     priority touches either cluster, so A and B are the clusters the sequential pass would see -- insert mutex (A, B);
//   * an attractive edge (A, B) is executed when it IS bestA[A] and bestA[B]: every higher-priority edge at A or B is done
//     (the repulsive ones of this round first), so the mutex test and the union happen in the state the sequential pass has.
//   * an attractive edge that is bestA[A] only is executed as well when A is FREE: A carries no mutex and no live repulsive
//     edge.  Nothing can then block this edge before its sequential turn (a block needs a mutex at A; A's other attractive
//     edges come later than its top edge; A gets no mutex without a repulsive edge of its own), and A brings neither
//     mutexes nor repulsive edges into B, so B's own decisions are what they would have been.  This is what lets the
//     interior of an object assemble in a few rounds instead of one voxel per round around its growing core.
// Executed edges at different clusters commute; a cluster that is not free takes part in at most one union per round.
//
// A round costs O(window): roots come from finds on the forest (path halving), and the mutex set is keyed by EPOCH roots --
// the roots as of the last rebuild.  Every cluster keeps the list of the non-free epoch clusters it is made of (a union of
// two non-free clusters concatenates two lists; free clusters carry no mutex and are never listed), a mutex goes in under
// the epoch roots of its two voxels, and the test "is there a mutex between A and B" probes the pairs of the two lists.
// Every BS_MWS_EPOCH rounds (or when the probes of a round exceed a budget) the set is rebuilt: all voxels get their roots,
// the lists collapse to one entry per cluster and the mutex list is re-keyed and deduplicated -- the only O(V) + O(mutexes)
// step, which used to run every round.
// Distinct weights (the reference adds noise by default for this reason, post/mws.py:28-31) give O(log V)-ish rounds;
// long runs of equal weights zip up one voxel per round along the tie order.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include <cooperative_groups.h>

#include "common.cuh"
#include "../../include/bsnative.h"

namespace cg = cooperative_groups;

namespace bs {

static constexpr uint32_t NONE32 = 0xFFFFFFFFu;
static constexpr unsigned long long EMPTY64 = 0xFFFFFFFFFFFFFFFFull;
static constexpr uint32_t ATTR_BIT = 0x80000000u;
static constexpr unsigned FULL32 = 0xFFFFFFFFu;

struct MwsGeom {
    int C, Z, Y, X;
    int off[32][3], st[32][3];
    // stride lattice of channel c: z in [z0, z0 + nz * sz) step sz, ... (voxels whose partner lies inside the volume)
    int z0[32], y0[32], x0[32], nz[32], ny[32], nx[32];
    unsigned long long ebase[33];   // first edge slot of channel c (slots enumerate the lattice in raster order)
    double bias[32];
    int has_noise;
    double noise_eps, noise_k;
    unsigned long long seed;
    // blockwise use (volara ExtractFrags): the volume is nseg read ROIs of depth seg_z stacked along z; the lattices above
    // describe ONE segment, edges never cross a segment boundary, noise is keyed by (seg_seeds[segment], channel, voxel
    // within the segment).  Single volume: nseg = 1, seg_z = Z, seg_seeds = nullptr.
    int nseg, seg_z;
    const unsigned long long *seg_seeds;
};

__device__ __forceinline__ uint64_t mws_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t mws_mixw(uint64_t x, long long w) { return mws_splitmix64(x ^ ((uint64_t)w * 0x9E3779B97F4A7C15ull)); }

// slot -> (channel, voxel p, partner q)
__device__ __forceinline__ void mws_slot(const MwsGeom &G, unsigned long long e, int &c, uint32_t &p, uint32_t &q) {
    c = 0;
    while (c + 1 < G.C && e >= G.ebase[c + 1]) c++;
    unsigned long long k = e - G.ebase[c];
    const int ix = (int)(k % (unsigned)G.nx[c]);
    k /= (unsigned)G.nx[c];
    const int iy = (int)(k % (unsigned)G.ny[c]);
    k /= (unsigned)G.ny[c];
    const int iz = (int)(k % (unsigned)max(G.nz[c], 1));
    const int seg = (int)(k / (unsigned)max(G.nz[c], 1));
    const int z = seg * G.seg_z + G.z0[c] + iz * G.st[c][0], y = G.y0[c] + iy * G.st[c][1], x = G.x0[c] + ix * G.st[c][2];
    p = (uint32_t)(((long long)z * G.Y + y) * G.X + x);
    q = (uint32_t)(((long long)(z + G.off[c][0]) * G.Y + (y + G.off[c][1])) * G.X + (x + G.off[c][2]));
}

// weight of edge (c, p) exactly as numpy computes it (post/mws.py:33-57): affs_data (float64; uint8 input / 255.0; * mask),
// shift = zeros; shift += noise * eps; shift += bias; w = affs_data + shift
template <typename T>
__device__ __forceinline__ double mws_weight(const MwsGeom &G, const T *affs, const uint8_t *mask, int c, uint32_t p) {
    const size_t V = (size_t)G.Z * G.Y * G.X;
    double a;
    if constexpr (sizeof(T) == 1)
        a = __ddiv_rn((double)affs[(size_t)c * V + p], 255.0);
    else
        a = (double)affs[(size_t)c * V + p];
    if (mask && mask[p] == 0) a = __dmul_rn(a, 0.0);
    double sh = 0.0;
    if (G.has_noise) {
        // seeded stand-in for numpy's unseeded randn: sum of four 16-bit uniforms, centred and scaled to unit variance
        uint64_t seed = G.seed;
        long long pl = (long long)p;
        if (G.seg_seeds) {
            const long long segv = (long long)G.seg_z * G.Y * G.X;
            seed = G.seg_seeds[pl / segv];
            pl = pl % segv;
        }
        const uint64_t h = mws_mixw(mws_mixw(mws_mixw(seed, c), pl), 11);
        const long long sum = (long long)(h & 0xFFFF) + (long long)((h >> 16) & 0xFFFF) + (long long)((h >> 32) & 0xFFFF) + (long long)(h >> 48);
        const double n = __dmul_rn((double)(sum - 131070), G.noise_k);
        sh = __dadd_rn(sh, __dmul_rn(n, G.noise_eps));
    }
    sh = __dadd_rn(sh, G.bias[c]);
    return __dadd_rn(a, sh);
}

template <typename T>
__global__ void __launch_bounds__(256) k_mws_keys(MwsGeom G, const T *__restrict__ affs, const uint8_t *__restrict__ mask,
                                                  unsigned long long E, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (unsigned long long)gridDim.x * blockDim.x) {
        int c;
        uint32_t p, q;
        mws_slot(G, e, c, p, q);
        const double w = mws_weight<T>(G, affs, mask, c, p);
        // descending |w|: invert the bits of the non-negative double; NaN edges sort last and are dropped later
        uint64_t k = w != w ? EMPTY64 : ~(uint64_t)__double_as_longlong(fabs(w));
        keys[e] = k;
        vals[e] = (uint32_t)e;
    }
}

// sorted position -> endpoints (+ attractive bit); NaN edges get NONE32 / NONE32
template <typename T>
__global__ void __launch_bounds__(256) k_mws_endpoints(MwsGeom G, const T *__restrict__ affs, const uint8_t *__restrict__ mask,
                                                       unsigned long long E, const uint32_t *__restrict__ vals, int zero_is_repulsive,
                                                       uint32_t *__restrict__ eu, uint32_t *__restrict__ ev,
                                                       unsigned long long *__restrict__ counts) {
    unsigned long long nrep = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (unsigned long long)gridDim.x * blockDim.x) {
        int c;
        uint32_t p, q;
        mws_slot(G, vals[i], c, p, q);
        const double w = mws_weight<T>(G, affs, mask, c, p);
        if (w != w) {
            eu[i] = NONE32, ev[i] = NONE32;
        } else {
            const bool attractive = w > 0.0 || (w == 0.0 && !zero_is_repulsive);
            eu[i] = p | (attractive ? ATTR_BIT : 0u);
            ev[i] = q;
            nrep += attractive ? 0 : 1;
        }
    }
    nrep = __reduce_add_sync(0xFFFFFFFFu, (unsigned)nrep);
    if ((threadIdx.x & 31) == 0 && nrep) atomicAdd(&counts[0], nrep);
}

// root of x with path halving; safe while no union runs (phase A) and benign beside concurrent finds
__device__ __forceinline__ uint32_t mws_find(uint32_t *parent, uint32_t x) {
    for (;;) {
        const uint32_t p = __ldcg(&parent[x]);
        if (p == x) return x;
        const uint32_t gp = __ldcg(&parent[p]);
        if (gp == p) return p;
        parent[x] = gp;
        x = gp;
    }
}

__global__ void k_mws_init(uint32_t *__restrict__ parent, size_t V) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (size_t)gridDim.x * blockDim.x) parent[i] = (uint32_t)i;
}

// The rounds work on a WINDOW of the live edges: the lowest-ranked edges not executed yet (survivors of earlier rounds plus a
// refill from the sorted order).  A cluster with an edge in the window has its top edge in the window, so the tests above are
// exact on the window alone; edges beyond it simply wait.
// phase A: roots of the window's edges (stored for the later phases: parent[] is rewritten by the unions); dead edges (one
// cluster, NaN) are flagged; bestA over the attractive ones
__global__ void __launch_bounds__(256) k_mws_best(const uint32_t *__restrict__ win, size_t nwin, const uint32_t *__restrict__ eu,
                                                  const uint32_t *__restrict__ ev, uint32_t *parent,
                                                  uint32_t *__restrict__ bestA, uint8_t *__restrict__ keep, uint8_t *__restrict__ did,
                                                  uint2 *__restrict__ wroots) {
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < nwin; j += (size_t)gridDim.x * blockDim.x) {
        const uint32_t i = win[j];
        const uint32_t u = eu[i], v = ev[i];
        uint32_t ru = NONE32, rv = NONE32;
        if (v != NONE32) ru = mws_find(parent, u & ~ATTR_BIT), rv = mws_find(parent, v);
        wroots[j] = make_uint2(ru, rv);
        did[j] = 0;
        if (ru == rv) {          // also the NaN edges (NONE32, NONE32)
            keep[j] = 0;
            continue;
        }
        keep[j] = 1;
        if (u & ATTR_BIT) {
            atomicMin(&bestA[ru], i);
            atomicMin(&bestA[rv], i);
        }
    }
}

__device__ __forceinline__ uint64_t mws_hash(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return k;
}
// returns true if the key was new
__device__ __forceinline__ bool mws_set_insert(unsigned long long *tab, uint64_t mask, unsigned long long key) {
    uint64_t s = mws_hash(key) & mask;
    for (uint64_t n = 0; n <= mask; n++) {      // the table is kept at most half full; the bound only rules out spinning
        unsigned long long k = tab[s];
        if (k == key) return false;
        if (k == EMPTY64) {
            k = atomicCAS(&tab[s], EMPTY64, key);
            if (k == EMPTY64) return true;
            if (k == key) return false;
        }
        s = (s + 1) & mask;
    }
    return false;
}
__device__ __forceinline__ bool mws_set_has(const unsigned long long *tab, uint64_t mask, unsigned long long key) {
    uint64_t s = mws_hash(key) & mask;
    for (uint64_t n = 0; n <= mask; n++) {
        const unsigned long long k = tab[s];
        if (k == key) return true;
        if (k == EMPTY64) return false;
        s = (s + 1) & mask;
    }
    return false;
}

// phase B: repulsive edges with no pending higher-priority attractive edge at either cluster become mutexes
__global__ void __launch_bounds__(256) k_mws_repulsive(const uint32_t *__restrict__ win, size_t nwin, const uint32_t *__restrict__ eu,
                                                       const uint32_t *__restrict__ ev, const uint2 *__restrict__ wroots,
                                                       const uint32_t *__restrict__ bestA, const uint32_t *__restrict__ eroot,
                                                       uint8_t *__restrict__ keep, unsigned long long *__restrict__ tab, uint64_t tmask,
                                                       uint2 *__restrict__ mlist, unsigned long long *__restrict__ counts) {
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < nwin; j += (size_t)gridDim.x * blockDim.x) {
        if (!keep[j]) continue;
        const uint32_t i = win[j];
        const uint32_t u = eu[i], v = ev[i];
        if (u & ATTR_BIT) continue;
        const uint2 r = wroots[j];
        if (i < bestA[r.x] && i < bestA[r.y]) {
            keep[j] = 0;
            const uint32_t ea = eroot[u], eb = eroot[v];     // epoch clusters of the two voxels (different: r.x != r.y)
            const unsigned long long key = ((unsigned long long)min(ea, eb) << 32) | max(ea, eb);
            if (mws_set_insert(tab, tmask, key)) {
                const unsigned long long slot = atomicAdd(&counts[1], 1ull);    // mutex list length
                mlist[slot] = make_uint2(u, v);
            }
            atomicAdd(&counts[3], 1ull);
        }
    }
}

// phase C: attractive edges that are the top edge of both their clusters (union unless a mutex separates them), or the top
// edge of a FREE cluster (no mutex, no live repulsive edge: nothing can block it)
__global__ void __launch_bounds__(256) k_mws_attractive(const uint32_t *__restrict__ win, size_t nwin, const uint32_t *__restrict__ eu,
                                                        const uint2 *__restrict__ wroots, uint32_t *__restrict__ parent,
                                                        const uint32_t *__restrict__ bestA, const uint8_t *__restrict__ nonfree,
                                                        const uint32_t *__restrict__ ehead, const uint32_t *__restrict__ enext,
                                                        uint32_t *__restrict__ pairmark, uint32_t round, uint8_t *__restrict__ keep,
                                                        uint8_t *__restrict__ did, const unsigned long long *__restrict__ tab,
                                                        uint64_t tmask, unsigned long long *__restrict__ counts) {
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < nwin; j += (size_t)gridDim.x * blockDim.x) {
        if (!keep[j]) continue;
        const uint32_t i = win[j];
        if (!(eu[i] & ATTR_BIT)) continue;
        const uint32_t ru = wroots[j].x, rv = wroots[j].y;      // the roots as of the start of the round
        const bool top_u = bestA[ru] == i, top_v = bestA[rv] == i;
        if (!top_u && !top_v) continue;
        const bool nf_u = nonfree[ru], nf_v = nonfree[rv];
        const bool free_u = top_u && !nf_u, free_v = top_v && !nf_v;
        if (!(top_u && top_v) && !free_u && !free_v) continue;
        keep[j] = 0;
        bool blocked = false;
        if (nf_u && nf_v) {
            // a mutex between any epoch cluster of one side and any of the other
            unsigned long long probes = 0;
            for (uint32_t ea = ehead[ru]; ea != NONE32 && !blocked; ea = enext[ea])
                for (uint32_t eb = ehead[rv]; eb != NONE32; eb = enext[eb]) {
                    probes++;
                    if (mws_set_has(tab, tmask, ((unsigned long long)min(ea, eb) << 32) | max(ea, eb))) {
                        blocked = true;
                        break;
                    }
                }
            atomicAdd(&counts[8], probes);
        }
        if (blocked) {
            atomicAdd(&counts[4], 1ull);        // blocked by a mutex
        } else {
            if (nf_u && nf_v) pairmark[ru] = round, pairmark[rv] = round;   // the one union of two non-free clusters of their component
            did[j] = 1;
            uf_union(parent, ru, rv);           // lock-free, the smaller index becomes the root
            atomicAdd(&counts[2], 1ull);        // merges (all rounds) ...
            atomicAdd(&counts[5], 1ull);        // ... and of this round
        }
    }
}

// after the unions of a round: the epoch-cluster list of every new cluster.  A component of this round's unions holds at most
// one union of two non-free clusters (its thread concatenates the two lists) plus free clusters that joined (they carry no
// list); where the root moved to a free cluster's voxel the list moves with it.
__global__ void __launch_bounds__(256) k_mws_post_lists(const uint2 *__restrict__ wroots, const uint8_t *__restrict__ did, size_t nwin,
                                                        uint32_t *parent, const uint8_t *__restrict__ nonfree,
                                                        const uint32_t *__restrict__ pairmark, uint32_t round, uint32_t *ehead,
                                                        uint32_t *etail, uint32_t *enext) {
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < nwin; j += (size_t)gridDim.x * blockDim.x) {
        if (!did[j]) continue;
        const uint32_t x = wroots[j].x, y = wroots[j].y;
        const bool nfx = nonfree[x], nfy = nonfree[y];
        if (!nfx && !nfy) continue;
        const uint32_t R = mws_find(parent, x);
        if (nfx && nfy) {
            const uint32_t hx = ehead[x], tx = etail[x], hy = ehead[y], ty = etail[y];
            enext[tx] = hy;
            ehead[R] = hx;
            etail[R] = ty;
        } else {
            const uint32_t n = nfx ? x : y;
            if (pairmark[n] == round || R == n) continue;
            const uint32_t h = ehead[n], t = etail[n];
            ehead[R] = h;
            etail[R] = t;
        }
    }
}
// ... and its flag: a cluster that absorbed a non-free one is not free
__global__ void __launch_bounds__(256) k_mws_post_flags(const uint2 *__restrict__ wroots, const uint8_t *__restrict__ did, size_t nwin,
                                                        uint32_t *parent, uint8_t *nonfree) {
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < nwin; j += (size_t)gridDim.x * blockDim.x) {
        if (!did[j]) continue;
        const uint32_t x = wroots[j].x, y = wroots[j].y;
        if (nonfree[x] || nonfree[y]) nonfree[mws_find(parent, x)] = 1;
    }
}

// bestA back to "none" for the clusters the window touched (instead of clearing the whole array every round)
__global__ void __launch_bounds__(256) k_mws_reset_best(const uint2 *__restrict__ wroots, size_t nwin, uint32_t *__restrict__ bestA) {
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < nwin; j += (size_t)gridDim.x * blockDim.x) {
        const uint2 r = wroots[j];
        if (r.x != NONE32) bestA[r.x] = NONE32, bestA[r.y] = NONE32;
    }
}

// survivors keep their order at the front of the next window, the refill continues the sorted order behind them
__global__ void k_mws_next_window(const uint32_t *__restrict__ win, const uint8_t *__restrict__ keep, const uint32_t *__restrict__ pos,
                                  size_t nwin, const uint32_t *__restrict__ nkeep_dev, uint32_t cursor, uint32_t wcap, uint32_t E,
                                  uint32_t *__restrict__ out) {
    const uint32_t nkeep = *nkeep_dev;
    const uint32_t nfill = min(wcap - nkeep, E - cursor);
    const size_t n = max(nwin, (size_t)nfill);
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x) {
        if (j < nwin && keep[j]) out[pos[j]] = win[j];
        if (j < nfill) out[nkeep + j] = cursor + (uint32_t)j;
    }
}

// rebuild, part 1: roots of all voxels (= their epoch roots until the next rebuild), into `out` and -- path compression -- into
// the forest itself; every cluster's epoch list collapses to the cluster itself
__global__ void k_mws_flatten(uint32_t *parent, uint32_t *__restrict__ out, uint32_t *__restrict__ ehead,
                              uint32_t *__restrict__ etail, uint32_t *__restrict__ enext, size_t V) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t x = (uint32_t)i, p = parent[x];
        while (p != x) {
            x = p;
            p = parent[x];
        }
        out[i] = x;
        parent[i] = x;              // in place as well: a walker that meets the new value continues from an ancestor
        if (x == (uint32_t)i) {     // lists hang off roots only
            ehead[i] = (uint32_t)i;
            etail[i] = (uint32_t)i;
            enext[i] = NONE32;
        }
    }
}

// every voxel with a repulsive edge starts non-free
__global__ void __launch_bounds__(256) k_mws_mark_repulsive(const uint32_t *__restrict__ eu, const uint32_t *__restrict__ ev, size_t E,
                                                            uint8_t *__restrict__ nonfree) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t u = eu[i], v = ev[i];
        if (v != NONE32 && !(u & ATTR_BIT)) nonfree[u] = 1, nonfree[v] = 1;
    }
}

// re-key the mutexes by the clusters' new roots; duplicates leave the list
__global__ void __launch_bounds__(256) k_mws_rekey(const uint2 *__restrict__ mlist, size_t nm, const uint32_t *__restrict__ root,
                                                   unsigned long long *__restrict__ tab, uint64_t tmask, uint2 *__restrict__ mlist_out,
                                                   unsigned long long *__restrict__ n_out) {
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < nm; j += (size_t)gridDim.x * blockDim.x) {
        const uint2 m = mlist[j];
        const uint32_t ru = root[m.x], rv = root[m.y];
        const unsigned long long key = ((unsigned long long)min(ru, rv) << 32) | max(ru, rv);
        if (mws_set_insert(tab, tmask, key)) mlist_out[atomicAdd(n_out, 1ull)] = m;
    }
}

__global__ void k_mws_isroot(const uint32_t *__restrict__ parent, size_t V, uint8_t *__restrict__ flag) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (size_t)gridDim.x * blockDim.x) flag[i] = parent[i] == (uint32_t)i;
}
__global__ void k_mws_dense_labels(const uint32_t *__restrict__ parent, const uint32_t *__restrict__ rank, size_t V, uint32_t *__restrict__ labels) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (size_t)gridDim.x * blockDim.x)
        labels[i] = rank[uf_find(parent, (uint32_t)i)] + 1u;
}
template <typename L>
__global__ void k_mws_labels(const uint32_t *__restrict__ parent, size_t V, L *__restrict__ labels) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (size_t)gridDim.x * blockDim.x)
        labels[i] = (L)uf_find(parent, (uint32_t)i) + 1u;
}

// remove_small_objects(labels, min_size) of simple_mutex (post/watershed_mutex.py:272-277): labels with fewer voxels -> 0
__global__ void k_mws_count(const uint64_t *__restrict__ labels, size_t V, uint32_t *__restrict__ cnt) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (size_t)gridDim.x * blockDim.x)
        if (labels[i]) atomicAdd(&cnt[labels[i] - 1], 1u);
}
__global__ void k_mws_debris(const uint64_t *__restrict__ labels, size_t V, const uint32_t *__restrict__ cnt, uint32_t min_size,
                             uint64_t *__restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (size_t)gridDim.x * blockDim.x) {
        const uint64_t l = labels[i];
        out[i] = (l && cnt[l - 1] >= min_size) ? l : 0;
    }
}

// ------------------------------------------------------------------ the rounds of one epoch in ONE cooperative launch
// With a window of 2^17 edges the nine launches + host sync of a round cost more than its work; this kernel keeps the grid
// resident and runs round after round (phases separated by grid-wide barriers) until the window is empty, a rebuild of the
// mutex set is due, or max_rounds is reached.  Phase bodies = the kernels above; the survivors are compacted in order by a
// block-sum / ordered-scatter pass.  Arrays that change inside the kernel are accessed through plain (coherent) pointers.
static constexpr int COOP_NT = 512;
struct MwsCoop {
    const uint32_t *eu, *ev;
    uint32_t E, wcap;
    uint32_t *parent;
    const uint32_t *eroot;
    uint32_t *bestA, *ehead, *etail, *enext, *pairmark;
    uint8_t *nonfree;
    uint32_t *win[2];
    uint2 *wroots;
    uint8_t *keep, *did;
    unsigned long long *tab;
    unsigned long long tmask;
    uint2 *mlist;
    unsigned long long *cnt;   // the counters of mws_rounds
    uint32_t *ctl;             // [0] nwin [1] cursor [2] window buffer in use [3] rounds so far [4] rounds since the rebuild
                               // [5] unions since the rebuild [6] stop: 1 = rebuild due, 2 = a round executed nothing
    uint32_t *blocksum;        // gridDim.x
    uint32_t epoch, max_rounds, cap_rounds;
    unsigned long long probe_budget;
};

__global__ void __launch_bounds__(COOP_NT) k_mws_coop(MwsCoop a) {
    cg::grid_group grid = cg::this_grid();
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gn = (size_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ uint32_t s_w[COOP_NT / 32], s_w2[COOP_NT / 32];
    __shared__ uint32_t s_base, s_total;
    // every thread keeps the round's bookkeeping itself (same inputs, same result): no barrier for it
    uint32_t nwin = a.ctl[0], cursor = a.ctl[1], par = a.ctl[2], round = a.ctl[3], since = a.ctl[4], unions = a.ctl[5], stop = a.ctl[6];
    volatile unsigned long long *cntv = a.cnt;
    for (uint32_t it = 0; it < a.max_rounds && nwin != 0 && stop == 0; it++) {
        round++;
        const uint32_t *win = a.win[par];
        uint32_t *wout = a.win[par ^ 1];
        // per-round counters (merges, probes) rotate through three slots: the slot of the NEXT round is cleared now -- its last
        // readers (the bookkeeping of round - 2) are two barriers behind, its next writers one round ahead
        unsigned long long *rc = a.cnt + 16 + 4 * (round % 3);
        if (gtid == 0) a.cnt[16 + 4 * ((round + 1) % 3)] = 0, a.cnt[16 + 4 * ((round + 1) % 3) + 1] = 0;
        // ---- A: roots, dead edges, bestA
        for (size_t j = gtid; j < nwin; j += gn) {
            const uint32_t i = win[j];
            const uint32_t u = a.eu[i], v = a.ev[i];
            uint32_t ru = NONE32, rv = NONE32;
            if (v != NONE32) ru = mws_find(a.parent, u & ~ATTR_BIT), rv = mws_find(a.parent, v);
            a.wroots[j] = make_uint2(ru, rv);
            a.did[j] = 0;
            if (ru == rv) {
                a.keep[j] = 0;
                continue;
            }
            a.keep[j] = 1;
            if (u & ATTR_BIT) {
                atomicMin(&a.bestA[ru], i);
                atomicMin(&a.bestA[rv], i);
            }
        }
        grid.sync();
        // ---- B: repulsive edges
        for (size_t j = gtid; j < nwin; j += gn) {
            if (!a.keep[j]) continue;
            const uint32_t i = win[j];
            const uint32_t u = a.eu[i], v = a.ev[i];
            if (u & ATTR_BIT) continue;
            const uint2 r = a.wroots[j];
            if (i < a.bestA[r.x] && i < a.bestA[r.y]) {
                a.keep[j] = 0;
                const uint32_t ea = a.eroot[u], eb = a.eroot[v];
                const unsigned long long key = ((unsigned long long)min(ea, eb) << 32) | max(ea, eb);
                if (mws_set_insert(a.tab, a.tmask, key)) {
                    const unsigned long long slot = atomicAdd(&a.cnt[1], 1ull);
                    a.mlist[slot] = make_uint2(u, v);
                }
                atomicAdd(&a.cnt[3], 1ull);
            }
        }
        grid.sync();
        // ---- C: attractive edges; did = 1 | non-free(u) << 1 | non-free(v) << 2 for an executed union (the flags as of now:
        // the next phase must not read them while it also sets them)
        for (size_t j = gtid; j < nwin; j += gn) {
            if (!a.keep[j]) continue;
            const uint32_t i = win[j];
            if (!(a.eu[i] & ATTR_BIT)) continue;
            const uint32_t ru = a.wroots[j].x, rv = a.wroots[j].y;
            const bool top_u = a.bestA[ru] == i, top_v = a.bestA[rv] == i;
            if (!top_u && !top_v) continue;
            const bool nf_u = a.nonfree[ru], nf_v = a.nonfree[rv];
            const bool free_u = top_u && !nf_u, free_v = top_v && !nf_v;
            if (!(top_u && top_v) && !free_u && !free_v) continue;
            a.keep[j] = 0;
            bool blocked = false;
            if (nf_u && nf_v) {
                unsigned long long probes = 0;
                for (uint32_t ea = a.ehead[ru]; ea != NONE32 && !blocked; ea = a.enext[ea])
                    for (uint32_t eb = a.ehead[rv]; eb != NONE32; eb = a.enext[eb]) {
                        probes++;
                        if (mws_set_has(a.tab, a.tmask, ((unsigned long long)min(ea, eb) << 32) | max(ea, eb))) {
                            blocked = true;
                            break;
                        }
                    }
                atomicAdd(&rc[1], probes);
            }
            if (blocked) {
                atomicAdd(&a.cnt[4], 1ull);
            } else {
                if (nf_u && nf_v) a.pairmark[ru] = round, a.pairmark[rv] = round;
                a.did[j] = (uint8_t)(1u | (nf_u ? 2u : 0u) | (nf_v ? 4u : 0u));
                uf_union(a.parent, ru, rv);
                atomicAdd(&a.cnt[2], 1ull);
                atomicAdd(&rc[0], 1ull);
            }
        }
        grid.sync();
        // ---- D: epoch-cluster lists and non-free flags of the new clusters, bestA reset, survivors per CTA chunk
        for (size_t j = gtid; j < nwin; j += gn) {
            const uint2 r = a.wroots[j];
            if (r.x != NONE32) a.bestA[r.x] = NONE32, a.bestA[r.y] = NONE32;
            const uint32_t d = a.did[j];
            if (!d) continue;
            const uint32_t x = r.x, y = r.y;
            const bool nfx = (d & 2u) != 0, nfy = (d & 4u) != 0;
            if (!nfx && !nfy) continue;
            const uint32_t R = mws_find(a.parent, x);
            a.nonfree[R] = 1;
            if (nfx && nfy) {
                const uint32_t hx = a.ehead[x], tx = a.etail[x], hy = a.ehead[y], ty = a.etail[y];
                a.enext[tx] = hy;
                a.ehead[R] = hx;
                a.etail[R] = ty;
            } else {
                const uint32_t n = nfx ? x : y;
                if (a.pairmark[n] == round || R == n) continue;
                const uint32_t h = a.ehead[n], t = a.etail[n];
                a.ehead[R] = h;
                a.etail[R] = t;
            }
        }
        const uint32_t chunk = (nwin + gridDim.x - 1) / gridDim.x;
        const uint32_t lo = min(nwin, blockIdx.x * chunk), hi = min(nwin, lo + chunk);
        {
            uint32_t c = 0;
            for (uint32_t j = lo + threadIdx.x; j < hi; j += COOP_NT) c += a.keep[j];
            c = __reduce_add_sync(FULL32, c);
            if (lane == 0) s_w[warp] = c;
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t t = 0;
                for (int w = 0; w < COOP_NT / 32; w++) t += s_w[w];
                a.blocksum[blockIdx.x] = t;
            }
        }
        grid.sync();
        // ---- E: ordered compaction + refill
        {
            // gridDim.x <= COOP_NT: one block sum per thread, reduced by the block
            const uint32_t v = threadIdx.x < gridDim.x ? __ldcg(&a.blocksum[threadIdx.x]) : 0u;
            const uint32_t tb = __reduce_add_sync(FULL32, threadIdx.x < blockIdx.x ? v : 0u), tt = __reduce_add_sync(FULL32, v);
            __syncthreads();                       // s_w of phase D is consumed
            if (lane == 0) s_w[warp] = tb, s_w2[warp] = tt;
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t b = 0, t = 0;
                for (int w = 0; w < COOP_NT / 32; w++) b += s_w[w], t += s_w2[w];
                s_base = b;
                s_total = t;
            }
            __syncthreads();
        }
        const uint32_t nkeep = s_total;
        uint32_t run = s_base;
        for (uint32_t j0 = lo; j0 < hi; j0 += COOP_NT) {
            const uint32_t j = j0 + threadIdx.x;
            const bool f = j < hi && a.keep[j];
            const unsigned bal = __ballot_sync(FULL32, f);
            __syncthreads();                       // s_w of the previous tile is consumed
            if (lane == 0) s_w[warp] = __popc(bal);
            __syncthreads();
            uint32_t wpre = 0, tot = 0;
            for (int w = 0; w < COOP_NT / 32; w++) {
                const uint32_t v = s_w[w];
                if (w < warp) wpre += v;
                tot += v;
            }
            if (f) wout[run + wpre + __popc(bal & ((1u << lane) - 1u))] = win[j];
            run += tot;
        }
        const uint32_t nfill = min(a.wcap - nkeep, a.E - cursor);
        for (size_t j = gtid; j < nfill; j += gn) wout[nkeep + j] = cursor + (uint32_t)j;
        grid.sync();
        // ---- bookkeeping, by every thread for itself (the counters of this round are complete and stay until round + 2)
        const unsigned long long merges = cntv[16 + 4 * (round % 3)], probes = cntv[16 + 4 * (round % 3) + 1];
        const uint32_t nnew = nkeep + nfill;
        since++;
        if (merges > 0) unions = 1;
        if (nkeep == nwin && nfill == 0)
            stop = 2;
        else if (nnew > 0 && ((unions && (since >= a.epoch || probes > a.probe_budget)) || since >= a.cap_rounds))
            stop = 1;
        cursor += nfill;
        nwin = nnew;
        par ^= 1;
    }
    if (gtid == 0) {
        a.ctl[0] = nwin, a.ctl[1] = cursor, a.ctl[2] = par, a.ctl[3] = round, a.ctl[4] = since, a.ctl[5] = unions, a.ctl[6] = stop;
    }
}

// edges per round (BS_MWS_WINDOW overrides; the result does not depend on it)
static unsigned long long g_mws_window = getenv("BS_MWS_WINDOW") ? strtoull(getenv("BS_MWS_WINDOW"), nullptr, 10) : (1ull << 17);

// rounds between two rebuilds of the mutex set (BS_MWS_EPOCH), and the number of set probes in one round that forces one early
static unsigned long long g_mws_epoch = getenv("BS_MWS_EPOCH") ? strtoull(getenv("BS_MWS_EPOCH"), nullptr, 10) : 64;
static unsigned long long g_mws_probe_budget = getenv("BS_MWS_PROBES") ? strtoull(getenv("BS_MWS_PROBES"), nullptr, 10) : (1ull << 23);

// BS_MWS_COOP=0: one set of launches per round instead of the cooperative kernel (same results)
static double g_mws_rebuild_s = 0.0;   // BS_MWS_VERBOSE: seconds spent in rebuilds (cooperative path)
static int g_mws_coop = getenv("BS_MWS_COOP") ? atoi(getenv("BS_MWS_COOP")) : 1;

static unsigned grid_for(size_t n) { return (unsigned)std::min<size_t>(std::max<size_t>((n + 255) / 256, 1), 148 * 16); }

// The rounds: eu (voxel / node | ATTR_BIT, NONE32 = dropped), ev, E edges in sequential order (rank = index), V nodes, nrep =
// number of repulsive edges.  `parent` (4 V bytes, identity on entry) holds the forest on return (root = smallest index).
// h_cnt: [2] merges [3] mutex edges [4] blocked; rounds / rebuilds counted.
static int mws_rounds(const uint32_t *eu, const uint32_t *ev, unsigned long long E, size_t V, unsigned long long nrep, uint32_t *parent,
                      unsigned long long *d_cnt, unsigned long long *h_cnt, int *rounds_out, int *rebuilds_out, cudaStream_t s) {
    DevBuf root, nonfree, win, win2, wroots, keep, did, pos, bestA, tab, mlist, mlist2, ehead, etail, enext, pairmark;
    int rounds = 0, rebuilds = 0;
    {
        // mutex set: at most one entry per executed repulsive edge
        uint64_t tcap = 1024;
        while (tcap < 2 * nrep + 16) tcap <<= 1;
        BS_TRY(tab.alloc_fill(8 * tcap, 0xFF, s));
        // ... of which an epoch uses a prefix sized for the mutexes it starts with plus what its rounds can add (a window of
        // repulsive edges per round): clearing and probing a table of the live size instead of the worst-case one
        const uint64_t win_edges = std::min<unsigned long long>(E, g_mws_window);
        auto table_for = [&](uint64_t live) {
            uint64_t t = 1024;
            while (t < 2 * (live + win_edges * (g_mws_epoch + 1)) + 16 && t < tcap) t <<= 1;
            return std::min(t, tcap);
        };
        uint64_t teff = table_for(0);
        // rounds an epoch may last before its inserts could fill half the table (reached only by epochs without unions)
        // (the whole table holds every mutex the run can make: no limit then)
        auto rounds_for = [&](uint64_t live) {
            if (teff == tcap) return (uint32_t)(1u << 30);
            return (uint32_t)std::min<uint64_t>((teff / 2 > live ? (teff / 2 - live) / win_edges : 1), 1u << 30);
        };
        uint32_t cap_rounds = std::max<uint32_t>(rounds_for(0), 1);
        BS_TRY(mlist.alloc(8 * (size_t)(nrep + 1), s));
        BS_TRY(mlist2.alloc(8 * (size_t)(nrep + 1), s));
        BS_TRY(bestA.alloc_fill(4 * V, 0xFF, s));
        BS_TRY(root.alloc(4 * V, s));
        BS_TRY(ehead.alloc(4 * V, s));
        BS_TRY(etail.alloc(4 * V, s));
        BS_TRY(enext.alloc(4 * V, s));
        BS_TRY(pairmark.alloc_zero(4 * V, s));
        BS_TRY(nonfree.alloc_zero(V, s));
        // epoch 0: every voxel is its own cluster
        BS_LAUNCH(k_mws_flatten, grid_for(V), 256, 0, s, parent, root.as<uint32_t>(), ehead.as<uint32_t>(), etail.as<uint32_t>(),
                  enext.as<uint32_t>(), V);
        BS_LAUNCH(k_mws_mark_repulsive, grid_for(E), 256, 0, s, eu, ev, (size_t)E, nonfree.as<uint8_t>());
        const uint32_t wcap = (uint32_t)std::min<unsigned long long>(E, g_mws_window);
        BS_TRY(win.alloc(4 * (size_t)wcap, s));
        BS_TRY(win2.alloc(4 * (size_t)wcap, s));
        BS_TRY(wroots.alloc(8 * (size_t)wcap, s));
        BS_TRY(keep.alloc((size_t)wcap, s));
        BS_TRY(did.alloc((size_t)wcap, s));
        BS_TRY(pos.alloc(4 * (size_t)wcap, s));
        // first window: the wcap best edges
        uint32_t cursor = 0;
        size_t nwin = 0;
        BS_CUDA(cudaMemsetAsync(d_cnt + 6, 0, 8, s));
        BS_LAUNCH(k_mws_next_window, grid_for(wcap), 256, 0, s, win.as<uint32_t>(), keep.as<uint8_t>(), pos.as<uint32_t>(), (size_t)0,
                  (const uint32_t *)(d_cnt + 6), cursor, wcap, (uint32_t)E, win.as<uint32_t>());
        nwin = wcap;
        cursor = wcap;
        int since_rebuild = 0;
        bool unions_since_rebuild = false;
        // ---- cooperative path: all rounds of an epoch in one launch
        int coop_grid = 0;
        if (g_mws_coop) {
            int dev = 0, ok = 0, sms = 0, nb = 0;
            if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&ok, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess && ok &&
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_mws_coop, COOP_NT, 0) == cudaSuccess && nb >= 1)
                coop_grid = std::min(getenv("BS_MWS_COOP_GRID") ? atoi(getenv("BS_MWS_COOP_GRID")) : 2 * sms, std::min(sms * nb, COOP_NT));
            else
                cudaGetLastError();
        }
        if (coop_grid > 0) {
            DevBuf ctl, blocksum;
            BS_TRY(ctl.alloc_zero(32, s));
            BS_TRY(blocksum.alloc_zero(4 * (size_t)coop_grid, s));
            uint32_t h_ctl[8] = {(uint32_t)nwin, cursor, 0, 0, 0, 0, 0, 0};
            BS_CUDA(cudaMemcpyAsync(ctl.p, h_ctl, 32, cudaMemcpyHostToDevice, s));
            BS_CUDA(cudaMemsetAsync(d_cnt + 5, 0, 8, s));
            BS_CUDA(cudaMemsetAsync(d_cnt + 8, 0, 8, s));
            MwsCoop A;
            A.eu = eu, A.ev = ev, A.E = (uint32_t)E, A.wcap = wcap;
            A.parent = parent, A.eroot = root.as<uint32_t>();
            A.bestA = bestA.as<uint32_t>(), A.ehead = ehead.as<uint32_t>(), A.etail = etail.as<uint32_t>(), A.enext = enext.as<uint32_t>();
            A.pairmark = pairmark.as<uint32_t>(), A.nonfree = nonfree.as<uint8_t>();
            A.win[0] = win.as<uint32_t>(), A.win[1] = win2.as<uint32_t>();
            A.wroots = wroots.as<uint2>(), A.keep = keep.as<uint8_t>(), A.did = did.as<uint8_t>();
            A.tab = tab.as<unsigned long long>();
            A.cnt = d_cnt, A.ctl = ctl.as<uint32_t>(), A.blocksum = blocksum.as<uint32_t>();
            A.epoch = (uint32_t)g_mws_epoch, A.max_rounds = 1u << 14, A.probe_budget = g_mws_probe_budget;
            for (;;) {
                A.mlist = mlist.as<uint2>();          // swapped by a rebuild
                A.tmask = teff - 1, A.cap_rounds = cap_rounds;
                void *args[] = {&A};
                BS_CUDA(cudaLaunchCooperativeKernel((void *)k_mws_coop, dim3((unsigned)coop_grid), dim3(COOP_NT), args, 0, s));
                g_launches++;
                BS_CUDA(cudaMemcpyAsync(h_ctl, ctl.p, 32, cudaMemcpyDeviceToHost, s));
                BS_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, 128, cudaMemcpyDeviceToHost, s));
                BS_CUDA(cudaStreamSynchronize(s));
                rounds = (int)h_ctl[3];
                BS_ARG(h_ctl[6] != 2, "bs_mws_agglom: a round executed no edge (internal error)");
                if (h_ctl[0] == 0) break;
                if (h_ctl[6] == 1) {
                    rebuilds++;
                    const auto r0 = std::chrono::steady_clock::now();
                    BS_LAUNCH(k_mws_flatten, grid_for(V), 256, 0, s, parent, root.as<uint32_t>(), ehead.as<uint32_t>(), etail.as<uint32_t>(),
                              enext.as<uint32_t>(), V);
                    const size_t nm = (size_t)h_cnt[1];
                    teff = table_for(nm);
                    cap_rounds = std::max<uint32_t>(rounds_for(nm), 1);
                    BS_CUDA(cudaMemsetAsync(tab.p, 0xFF, 8 * teff, s));
                    if (nm) {
                        BS_CUDA(cudaMemsetAsync(d_cnt + 1, 0, 8, s));
                        BS_LAUNCH(k_mws_rekey, grid_for(nm), 256, 0, s, mlist.as<uint2>(), nm, root.as<uint32_t>(), tab.as<unsigned long long>(),
                                  teff - 1, mlist2.as<uint2>(), d_cnt + 1);
                        mlist.swap(mlist2);
                    }
                    BS_CUDA(cudaMemsetAsync(ctl.as<uint32_t>() + 4, 0, 12, s));
                    if (getenv("BS_MWS_VERBOSE")) {
                        cudaStreamSynchronize(s);
                        g_mws_rebuild_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - r0).count();
                    }
                }
            }
            nwin = 0;
        }
        while (nwin > 0) {
            rounds++;
            since_rebuild++;
            BS_CUDA(cudaMemsetAsync(d_cnt + 5, 0, 8, s));
            BS_CUDA(cudaMemsetAsync(d_cnt + 8, 0, 8, s));
            BS_LAUNCH(k_mws_best, grid_for(nwin), 256, 0, s, win.as<uint32_t>(), nwin, eu, ev, parent,
                      bestA.as<uint32_t>(), keep.as<uint8_t>(), did.as<uint8_t>(), wroots.as<uint2>());
            BS_LAUNCH(k_mws_repulsive, grid_for(nwin), 256, 0, s, win.as<uint32_t>(), nwin, eu, ev,
                      wroots.as<uint2>(), bestA.as<uint32_t>(), root.as<uint32_t>(), keep.as<uint8_t>(), tab.as<unsigned long long>(), teff - 1,
                      mlist.as<uint2>(), d_cnt);
            BS_LAUNCH(k_mws_attractive, grid_for(nwin), 256, 0, s, win.as<uint32_t>(), nwin, eu, wroots.as<uint2>(),
                      parent, bestA.as<uint32_t>(), nonfree.as<uint8_t>(), ehead.as<uint32_t>(), enext.as<uint32_t>(),
                      pairmark.as<uint32_t>(), (uint32_t)rounds, keep.as<uint8_t>(), did.as<uint8_t>(), tab.as<unsigned long long>(), teff - 1,
                      d_cnt);
            BS_LAUNCH(k_mws_post_lists, grid_for(nwin), 256, 0, s, wroots.as<uint2>(), did.as<uint8_t>(), nwin, parent,
                      nonfree.as<uint8_t>(), pairmark.as<uint32_t>(), (uint32_t)rounds, ehead.as<uint32_t>(), etail.as<uint32_t>(),
                      enext.as<uint32_t>());
            BS_LAUNCH(k_mws_post_flags, grid_for(nwin), 256, 0, s, wroots.as<uint2>(), did.as<uint8_t>(), nwin, parent,
                      nonfree.as<uint8_t>());
            BS_LAUNCH(k_mws_reset_best, grid_for(nwin), 256, 0, s, wroots.as<uint2>(), nwin, bestA.as<uint32_t>());
            // next window: survivors (order preserved) + refill
            BS_TRY(scan_exclusive_u8(keep.as<uint8_t>(), pos.as<uint32_t>(), nwin, (uint32_t *)(d_cnt + 6), s));
            BS_LAUNCH(k_mws_next_window, grid_for(wcap), 256, 0, s, win.as<uint32_t>(), keep.as<uint8_t>(), pos.as<uint32_t>(), nwin,
                      (const uint32_t *)(d_cnt + 6), cursor, wcap, (uint32_t)E, win2.as<uint32_t>());
            win.swap(win2);
            BS_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, 128, cudaMemcpyDeviceToHost, s));
            BS_CUDA(cudaStreamSynchronize(s));
            const uint32_t nkeep = (uint32_t)h_cnt[6];
            const uint32_t nfill = std::min<uint32_t>(wcap - nkeep, (uint32_t)E - cursor);
            BS_ARG(nkeep < nwin || nfill > 0, "bs_mws_agglom: a round executed no edge (internal error)");
            cursor += nfill;
            nwin = (size_t)nkeep + nfill;
            if (h_cnt[5] > 0) unions_since_rebuild = true;
            if (nwin > 0 && ((unions_since_rebuild && ((unsigned long long)since_rebuild >= g_mws_epoch || h_cnt[8] > g_mws_probe_budget)) ||
                             (uint32_t)since_rebuild >= cap_rounds)) {
                // rebuild: roots of all voxels become the epoch roots, the lists collapse, the mutex set is re-keyed
                rebuilds++;
                since_rebuild = 0;
                unions_since_rebuild = false;
                BS_LAUNCH(k_mws_flatten, grid_for(V), 256, 0, s, parent, root.as<uint32_t>(), ehead.as<uint32_t>(),
                          etail.as<uint32_t>(), enext.as<uint32_t>(), V);
                const size_t nm = (size_t)h_cnt[1];
                teff = table_for(nm);
                cap_rounds = std::max<uint32_t>(rounds_for(nm), 1);
                BS_CUDA(cudaMemsetAsync(tab.p, 0xFF, 8 * teff, s));
                if (nm) {
                    BS_CUDA(cudaMemsetAsync(d_cnt + 1, 0, 8, s));
                    BS_LAUNCH(k_mws_rekey, grid_for(nm), 256, 0, s, mlist.as<uint2>(), nm, root.as<uint32_t>(), tab.as<unsigned long long>(),
                              teff - 1, mlist2.as<uint2>(), d_cnt + 1);
                    mlist.swap(mlist2);
                }
            }
        }
    }
    *rounds_out = rounds;
    *rebuilds_out = rebuilds;
    return BS_OK;
}

template <typename T>
static int mws_run(const T *affs, const uint8_t *mask, MwsGeom &G, int zero_is_repulsive, int remove_debris, uint64_t *labels_out,
                   uint32_t *labels32_out, uint64_t *seg_out, int64_t *counters_out, cudaStream_t s) {
    const size_t V = (size_t)G.Z * G.Y * G.X;
    const unsigned long long E = G.ebase[G.C];
    BS_ARG(V < (1ull << 31), "bs_mws_agglom: volume too large for 31-bit voxel indices");
    BS_ARG(E < (1ull << 32) - 1, "bs_mws_agglom: more than 2^32 edges");
    DevBuf parent, keys, keys2, vals, vals2, eu, ev, counts;
    BS_TRY(parent.alloc(4 * V, s));
    BS_LAUNCH(k_mws_init, grid_for(V), 256, 0, s, parent.as<uint32_t>(), V);
    BS_TRY(counts.alloc_zero(256, s));
    unsigned long long *d_cnt = counts.as<unsigned long long>();   // [0] repulsive edges [1] mutex list length [2] merges
                                                                   // [3] repulsive executed [4] blocked [5] merges this round [6] survivors
                                                                   // [8] mutex-set probes of this round
    unsigned long long h_cnt[16];
    memset(h_cnt, 0, sizeof(h_cnt));
    int rounds = 0, rebuilds = 0;
    const bool verbose = getenv("BS_MWS_VERBOSE") != nullptr;
    auto now = [&]() {
        if (verbose) cudaStreamSynchronize(s);
        return std::chrono::steady_clock::now();
    };
    const auto t0 = now();
    auto t1 = t0, t2 = t0;
    if (E) {
        BS_TRY(keys.alloc(8 * (size_t)E, s));
        BS_TRY(keys2.alloc(8 * (size_t)E, s));
        BS_TRY(vals.alloc(4 * (size_t)E, s));
        BS_TRY(vals2.alloc(4 * (size_t)E, s));
        BS_LAUNCH((k_mws_keys<T>), grid_for(E), 256, 0, s, G, affs, mask, E, keys.as<uint64_t>(), vals.as<uint32_t>());
        BS_TRY(radix_sort_pairs(keys.as<uint64_t>(), vals.as<uint32_t>(), keys2.as<uint64_t>(), vals2.as<uint32_t>(), (size_t)E, 0, 64, s));
        keys.release();
        keys2.release();
        vals2.release();
        BS_TRY(eu.alloc(4 * (size_t)E, s));
        BS_TRY(ev.alloc(4 * (size_t)E, s));
        BS_LAUNCH((k_mws_endpoints<T>), grid_for(E), 256, 0, s, G, affs, mask, E, vals.as<uint32_t>(), zero_is_repulsive, eu.as<uint32_t>(),
                  ev.as<uint32_t>(), d_cnt);
        vals.release();
        BS_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, 128, cudaMemcpyDeviceToHost, s));
        BS_CUDA(cudaStreamSynchronize(s));
        t1 = now();
        BS_TRY(mws_rounds(eu.as<uint32_t>(), ev.as<uint32_t>(), E, V, h_cnt[0], parent.as<uint32_t>(), d_cnt, h_cnt, &rounds, &rebuilds, s));
        t2 = now();
    }
    h_cnt[7] = (unsigned long long)rounds;
    if (verbose)
        fprintf(stderr, "[bs mws] edges %llu rounds %d rebuilds %d; rank edges %.3f s, rounds + rebuilds %.3f s (rebuilds %.3f s)\n",
                (unsigned long long)E, rounds, rebuilds, std::chrono::duration<double>(t1 - t0).count(),
                std::chrono::duration<double>(t2 - t1).count(), g_mws_rebuild_s), g_mws_rebuild_s = 0.0;
    unsigned long long n_labels = 0;
    if (labels32_out) {
        // dense labels 1..n in the order of every cluster's first voxel
        DevBuf flag, rank, tot;
        BS_TRY(flag.alloc(V, s));
        BS_TRY(rank.alloc(4 * V, s));
        BS_TRY(tot.alloc_zero(16, s));
        BS_LAUNCH(k_mws_isroot, grid_for(V), 256, 0, s, parent.as<uint32_t>(), V, flag.as<uint8_t>());
        BS_TRY(scan_exclusive_u8(flag.as<uint8_t>(), rank.as<uint32_t>(), V, tot.as<uint32_t>(), s));
        BS_LAUNCH(k_mws_dense_labels, grid_for(V), 256, 0, s, parent.as<uint32_t>(), rank.as<uint32_t>(), V, labels32_out);
        uint32_t h_tot = 0;
        BS_CUDA(cudaMemcpyAsync(&h_tot, tot.p, 4, cudaMemcpyDeviceToHost, s));
        BS_CUDA(cudaStreamSynchronize(s));
        n_labels = h_tot;
    }
    if (labels_out) BS_LAUNCH((k_mws_labels<uint64_t>), grid_for(V), 256, 0, s, parent.as<uint32_t>(), V, labels_out);
    if (seg_out && labels_out) {
        if (remove_debris > 0) {
            DevBuf cnt;
            BS_TRY(cnt.alloc_zero(4 * V, s));
            BS_LAUNCH(k_mws_count, grid_for(V), 256, 0, s, labels_out, V, cnt.as<uint32_t>());
            BS_LAUNCH(k_mws_debris, grid_for(V), 256, 0, s, labels_out, V, cnt.as<uint32_t>(), (uint32_t)remove_debris, seg_out);
            BS_CUDA(cudaStreamSynchronize(s));
        } else {
            BS_CUDA(cudaMemcpyAsync(seg_out, labels_out, 8 * V, cudaMemcpyDeviceToDevice, s));
        }
    }
    BS_CUDA(cudaStreamSynchronize(s));
    BS_CUDA(cudaGetLastError());
    if (counters_out) {
        counters_out[0] = (int64_t)E;
        counters_out[1] = (int64_t)h_cnt[2];
        counters_out[2] = (int64_t)h_cnt[3];
        counters_out[3] = (int64_t)h_cnt[4];
        counters_out[4] = (int64_t)h_cnt[7];
        if (labels32_out) counters_out[5] = (int64_t)n_labels;
    }
    return BS_OK;
}

static int mws_agglom_impl(const void *affs, int aff_dtype, const uint8_t *mask, int C, int nseg, int Z, int Y, int X, const int32_t *offsets,
                           const int32_t *strides, const double *bias, double noise_eps, unsigned long long noise_seed,
                           const unsigned long long *seg_seeds_host, int zero_is_repulsive, int remove_debris, uint64_t *labels_out,
                           uint32_t *labels32_out, uint64_t *seg_out, int64_t *counters_out, cudaStream_t s) {
    BS_ARG(C >= 1 && C <= 32, "bs_mws_agglom: 1..32 affinity channels");
    BS_ARG(Z > 0 && Y > 0 && X > 0 && nseg >= 1, "bs_mws_agglom: empty volume");
    BS_ARG((long long)nseg * Z < (1LL << 31), "bs_mws_agglom: volume too large");
    MwsGeom G;
    memset(&G, 0, sizeof(G));
    G.C = C, G.Z = nseg * Z, G.Y = Y, G.X = X;
    G.nseg = nseg, G.seg_z = Z;
    const int dims[3] = {Z, Y, X};
    unsigned long long e = 0;
    for (int c = 0; c < C; c++) {
        int first[3], count[3];
        for (int d = 0; d < 3; d++) {
            G.off[c][d] = offsets[3 * c + d];
            G.st[c][d] = strides ? strides[3 * c + d] : 1;
            BS_ARG(G.st[c][d] >= 1, "bs_mws_agglom: strides must be >= 1");
            // voxels v with 0 <= v + off < n and v % stride == 0
            const int lo = std::max(0, -G.off[c][d]), hi = std::min(dims[d], dims[d] - G.off[c][d]);   // [lo, hi)
            const int f = ((lo + G.st[c][d] - 1) / G.st[c][d]) * G.st[c][d];
            first[d] = f;
            count[d] = hi > f ? (hi - f + G.st[c][d] - 1) / G.st[c][d] : 0;
        }
        G.z0[c] = first[0], G.y0[c] = first[1], G.x0[c] = first[2];
        G.nz[c] = count[0], G.ny[c] = count[1], G.nx[c] = count[2];
        G.ebase[c] = e;
        e += (unsigned long long)nseg * count[0] * count[1] * count[2];
        if (count[0] == 0 || count[1] == 0 || count[2] == 0) G.nz[c] = 0, G.ny[c] = 1, G.nx[c] = 1;   // empty lattice, no division by zero
        G.bias[c] = bias ? bias[c] : 0.0;
    }
    G.ebase[C] = e;
    G.has_noise = noise_eps != 0.0;
    G.noise_eps = noise_eps;
    G.noise_k = 1.7320508075688772 / 65536.0;    // sqrt(3) / 2^16: unit variance for the sum of four 16-bit uniforms
    G.seed = noise_seed;
    DevBuf seeds;
    if (seg_seeds_host && G.has_noise) {
        BS_TRY(seeds.alloc(8 * (size_t)nseg, s));
        BS_CUDA(cudaMemcpyAsync(seeds.p, seg_seeds_host, 8 * (size_t)nseg, cudaMemcpyHostToDevice, s));
        BS_CUDA(cudaStreamSynchronize(s));
        G.seg_seeds = seeds.as<unsigned long long>();
    }
    if (aff_dtype == BS_DTYPE_U8)
        return mws_run<uint8_t>((const uint8_t *)affs, mask, G, zero_is_repulsive, remove_debris, labels_out, labels32_out, seg_out,
                                counters_out, s);
    return mws_run<float>((const float *)affs, mask, G, zero_is_repulsive, remove_debris, labels_out, labels32_out, seg_out, counters_out, s);
}

int mws_agglom(const void *affs, int aff_dtype, const uint8_t *mask, int C, int Z, int Y, int X, const int32_t *offsets,
               const int32_t *strides, const double *bias, double noise_eps, unsigned long long noise_seed, int zero_is_repulsive,
               int remove_debris, uint64_t *labels_out, uint64_t *seg_out, int64_t *counters_out, cudaStream_t s) {
    return mws_agglom_impl(affs, aff_dtype, mask, C, 1, Z, Y, X, offsets, strides, bias, noise_eps, noise_seed, nullptr, zero_is_repulsive,
                           remove_debris, labels_out, nullptr, seg_out, counters_out, s);
}

// volara ExtractFrags' fragmenter for many blocks at once: n_blocks read ROIs (Z, Y, X) stacked along z, one independent mutex
// watershed each (edges never leave a block, noise keyed per block).  labels: 1 + smallest raveled index of the cluster in the
// stacked volume.
int mws_agglom_blocks(const void *affs, int aff_dtype, const uint8_t *mask, int C, int n_blocks, int Z, int Y, int X, const int32_t *offsets,
                      const int32_t *strides, const double *bias, double noise_eps, const unsigned long long *block_seeds_host,
                      int zero_is_repulsive, uint32_t *labels_out, int64_t *counters_out, cudaStream_t s) {
    return mws_agglom_impl(affs, aff_dtype, mask, C, n_blocks, Z, Y, X, offsets, strides, bias, noise_eps, 0, block_seeds_host,
                           zero_is_repulsive, 0, nullptr, labels_out, nullptr, counters_out, s);
}

// ------------------------------------------------------------------ mutex watershed on a graph (volara GraphMWS)
// nodes ascending ids; edge weight w = weight * score + bias in float64 (NaN scores dropped); visited by descending |w|, equal
// |w| in input order (the caller passes edges sorted by (u, v): declared tie rule, the upstream order is the database's);
// w > 0 attractive, w <= 0 repulsive.  out[i] = smallest node id of node i's cluster.
__global__ void __launch_bounds__(256) k_gmws_keys(const float *__restrict__ scores, size_t m, double weight, double bias,
                                                   uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += (size_t)gridDim.x * blockDim.x) {
        const double w = __dadd_rn(__dmul_rn(weight, (double)scores[e]), bias);
        keys[e] = w != w ? EMPTY64 : ~(uint64_t)__double_as_longlong(fabs(w));
        vals[e] = (uint32_t)e;
    }
}
__device__ __forceinline__ uint32_t gmws_rank(const uint64_t *__restrict__ nodes, uint32_t n, uint64_t id) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (nodes[mid] < id)
            lo = mid + 1;
        else
            hi = mid;
    }
    return (lo < n && nodes[lo] == id) ? lo : NONE32;
}
__global__ void __launch_bounds__(256) k_gmws_endpoints(const uint64_t *__restrict__ nodes, uint32_t n, const uint64_t *__restrict__ u,
                                                        const uint64_t *__restrict__ v, const float *__restrict__ scores, size_t m,
                                                        double weight, double bias, const uint32_t *__restrict__ vals,
                                                        uint32_t *__restrict__ eu, uint32_t *__restrict__ ev,
                                                        unsigned long long *__restrict__ counts) {
    unsigned long long nrep = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t e = vals[i];
        const double w = __dadd_rn(__dmul_rn(weight, (double)scores[e]), bias);
        const uint32_t a = gmws_rank(nodes, n, u[e]), b = gmws_rank(nodes, n, v[e]);
        if (w != w || a == NONE32 || b == NONE32 || a == b) {
            eu[i] = NONE32, ev[i] = NONE32;
        } else {
            const bool attractive = w > 0.0;
            eu[i] = a | (attractive ? ATTR_BIT : 0u);
            ev[i] = b;
            nrep += attractive ? 0 : 1;
        }
    }
    nrep = __reduce_add_sync(0xFFFFFFFFu, (unsigned)nrep);
    if ((threadIdx.x & 31) == 0 && nrep) atomicAdd(&counts[0], nrep);
}
__global__ void k_gmws_out(const uint32_t *__restrict__ parent, const uint64_t *__restrict__ nodes, size_t n, uint64_t *__restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = nodes[uf_find(parent, (uint32_t)i)];
}

int graph_mws(const uint64_t *nodes, int64_t n, const uint64_t *u, const uint64_t *v, const float *scores, int64_t m, double weight,
              double bias, uint64_t *out, int64_t *counters_out, cudaStream_t s) {
    BS_ARG(n >= 0 && n < (1LL << 31) && m >= 0 && m < (1LL << 32) - 1, "bs_graph_mws: graph too large");
    if (n == 0) return BS_OK;
    DevBuf parent, keys, keys2, vals, vals2, eu, ev, counts;
    BS_TRY(parent.alloc(4 * (size_t)n, s));
    BS_LAUNCH(k_mws_init, grid_for((size_t)n), 256, 0, s, parent.as<uint32_t>(), (size_t)n);
    BS_TRY(counts.alloc_zero(256, s));
    unsigned long long *d_cnt = counts.as<unsigned long long>();
    unsigned long long h_cnt[16];
    memset(h_cnt, 0, sizeof(h_cnt));
    int rounds = 0, rebuilds = 0;
    if (m) {
        BS_TRY(keys.alloc(8 * (size_t)m, s));
        BS_TRY(keys2.alloc(8 * (size_t)m, s));
        BS_TRY(vals.alloc(4 * (size_t)m, s));
        BS_TRY(vals2.alloc(4 * (size_t)m, s));
        BS_LAUNCH(k_gmws_keys, grid_for((size_t)m), 256, 0, s, scores, (size_t)m, weight, bias, keys.as<uint64_t>(), vals.as<uint32_t>());
        BS_TRY(radix_sort_pairs(keys.as<uint64_t>(), vals.as<uint32_t>(), keys2.as<uint64_t>(), vals2.as<uint32_t>(), (size_t)m, 0, 64, s));
        keys.release();
        keys2.release();
        vals2.release();
        BS_TRY(eu.alloc(4 * (size_t)m, s));
        BS_TRY(ev.alloc(4 * (size_t)m, s));
        BS_LAUNCH(k_gmws_endpoints, grid_for((size_t)m), 256, 0, s, nodes, (uint32_t)n, u, v, scores, (size_t)m, weight, bias, vals.as<uint32_t>(),
                  eu.as<uint32_t>(), ev.as<uint32_t>(), d_cnt);
        vals.release();
        BS_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, 128, cudaMemcpyDeviceToHost, s));
        BS_CUDA(cudaStreamSynchronize(s));
        BS_TRY(mws_rounds(eu.as<uint32_t>(), ev.as<uint32_t>(), (unsigned long long)m, (size_t)n, h_cnt[0], parent.as<uint32_t>(), d_cnt, h_cnt,
                          &rounds, &rebuilds, s));
    }
    BS_LAUNCH(k_gmws_out, grid_for((size_t)n), 256, 0, s, parent.as<uint32_t>(), nodes, (size_t)n, out);
    BS_CUDA(cudaStreamSynchronize(s));
    BS_CUDA(cudaGetLastError());
    if (counters_out) {
        counters_out[0] = m;
        counters_out[1] = (int64_t)h_cnt[2];
        counters_out[2] = (int64_t)h_cnt[3];
        counters_out[3] = (int64_t)h_cnt[4];
        counters_out[4] = rounds;
    }
    return BS_OK;
}

}  // namespace bs
