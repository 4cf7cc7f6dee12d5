// waterz BinQueue<256> agglomeration of one daisy block with the whole working set in shared memory.
//
// Same semantics as k_agglomerate (stage2.cu): exact emulation of waterz mergeUntil with the 256-bin FIFO
// queue and lazy stale re-scoring (post/blockwise/waterz_agglom.py:131-139), one warp per block.
// What differs is the data layout, chosen so that every dependent access of the sequential emulation is a
// shared-memory access (~30 cycles) instead of an L2 round trip (~700 cycles):
//   * nodes are numbered compactly per block (only fragments that carry an edge), so ids are 16 bit;
//   * per edge: endpoints u16, score f32, (time | dead) u16, affinity sum + count;
//   * incidence lists are singly linked lists threaded through the 2E half-edges (anext); merging two
//     clusters splices the lists in O(1) and dead entries are unlinked lazily while a list is walked;
//   * the "does (a, x) already exist" test of waterz' mergeRegions uses node marks: the absorbed cluster's
//     neighbours are marked with their edge, then the survivor's list is walked -- no pair hash;
//   * the 256 FIFO bins are chains of 8-entry chunks (u16 edge numbers) from a recycling pool, a 256-bit
//     occupancy bitmap finds the lowest non-empty bin.
// Blocks whose graph does not fit (agglom_smem_bytes > 227 KB) take the global-memory kernel.
#include <type_traits>

#include "agglom.cuh"

namespace bs {

#ifdef BS_TRACE
#define AGG_CHK(cond, tag)                                                                      \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            printf("[agg %d] check failed: %s (line %d, lane %d)\n", bi, tag, __LINE__, lane);  \
            fail = true;                                                                        \
        }                                                                                       \
    } while (0)
#else
#define AGG_CHK(cond, tag) \
    do {                   \
    } while (0)
#endif

static constexpr uint32_t NONE32 = 0xFFFFFFFFu;
static constexpr unsigned FULL = 0xFFFFFFFFu;
static constexpr int QCH = 8;        // entries per queue chunk = largest batch of pops per step
static constexpr int NBINS = 256;

__host__ __device__ static inline uint32_t smem_qc(uint32_t Ecap) { return Ecap / QCH + 2 * NBINS + 32; }

// bytes of one block's working set; idx = bytes per node / edge index (2 in shared memory, 4 in global memory)
size_t agglom_work_bytes(uint32_t Ecap, uint32_t Ncap, bool sum64, int idx) {
    size_t b = 0;
    b += (size_t)(sum64 ? 8 : 4) * Ecap;   // esum
    b += 4 * (size_t)Ecap * 2;             // escore, ecnt
    b += 32;                               // occupancy bitmap
    b += (size_t)idx * Ecap * 5;           // eu, ev, etd, anext[2E]
    const size_t QC = smem_qc(Ecap);
    b += (size_t)idx * (QC * QCH + QC);    // queue chunks + chunk links
    b += (size_t)idx * Ncap * 7;           // ufp, stamp, ahead, tnode, clevel, mark, markgen
    b += (size_t)idx * NBINS * 4;          // bin head chunk / tail chunk / head offset / tail fill
    return (b + 255) & ~(size_t)255;
}
size_t agglom_smem_bytes(uint32_t Ecap, uint32_t Ncap, bool sum64) { return agglom_work_bytes(Ecap, Ncap, sum64, 2); }

template <typename IdxT>
__device__ __forceinline__ uint32_t find16(IdxT *ufp, uint32_t x) {
    // path halving; concurrent lanes only ever write ancestors
    for (;;) {
        uint32_t p = ufp[x];
        if (p == x) return x;
        uint32_t gp = ufp[p];
        if (gp == p) return p;
        ufp[x] = (IdxT)gp;
        x = gp;
    }
}

// IdxT = uint16_t with the working set in shared memory (SMEM), uint32_t with it in a global-memory slab
// (blocks too large for shared memory; same algorithm, L1 / L2 latencies instead of shared-memory ones).
template <bool U8, typename SumT, typename IdxT, bool SMEM>
__global__ void __launch_bounds__(32) k_agglomerate_lists(const AggBlk *__restrict__ blks, const int *__restrict__ list,
                                                          AggArrays A, float threshold, int keep_cheaper, uint32_t Ecap_,
                                                          uint32_t Ncap_, unsigned char *__restrict__ gwork,
                                                          const unsigned long long *__restrict__ gwoff) {
    extern __shared__ __align__(16) unsigned char smraw[];
    constexpr uint32_t N16 = (uint32_t)(IdxT)~(IdxT)0;                 // "none" in the index type
    constexpr uint32_t DEADBIT = 1u << (8 * sizeof(IdxT) - 1);        // etd = time | dead flag
    const int bi = list[blockIdx.x];
    const AggBlk B = blks[bi];
    const int lane = threadIdx.x;
    const uint32_t E = B.E, nc = B.nv;
    const uint32_t Ecap = SMEM ? Ecap_ : ((max(E, 8u) + 7u) & ~7u), Ncap = SMEM ? Ncap_ : ((max(nc, 8u) + 7u) & ~7u);
    const uint32_t QC = smem_qc(Ecap);
    unsigned char *work = SMEM ? smraw : gwork + gwoff[blockIdx.x];

    SumT *esum = (SumT *)work;
    float *escore = (float *)(esum + Ecap);
    uint32_t *ecnt = (uint32_t *)(escore + Ecap);
    uint32_t *occ = ecnt + Ecap;
    IdxT *eu = (IdxT *)(occ + 8);
    IdxT *ev = eu + Ecap, *etd = ev + Ecap, *anext = etd + Ecap;
    IdxT *qent = anext + 2 * Ecap;
    IdxT *qcnext = qent + QC * QCH;
    IdxT *ufp = qcnext + QC, *stamp = ufp + Ncap, *ahead = stamp + Ncap, *tnode = ahead + Ncap, *clevel = tnode + Ncap,
         *mark = clevel + Ncap, *markgen = mark + Ncap;
    IdxT *bhc = markgen + Ncap, *btc = bhc + NBINS, *bho = btc + NBINS, *btf = bho + NBINS;

    const uint32_t *geu = A.eu + B.ebase, *gev = A.ev + B.ebase, *gcnt = A.ecnt + B.ebase;
    const unsigned long long *gsum = A.esum + B.ebase;
    uint32_t *tparent = A.tparent + 2 * (size_t)B.vbase, *tlevel = A.tlevel + 2 * (size_t)B.vbase;
    float *tscore = A.tscore + 2 * (size_t)B.vbase;
    uint32_t *ha = A.ha + B.vbase, *hb = A.hb + B.vbase;
    float *hs = A.hs + B.vbase;

    // ---- load the block's graph
    for (uint32_t i = lane; i < nc; i += 32) {
        ufp[i] = (IdxT)i;
        stamp[i] = 0;
        ahead[i] = N16;
        tnode[i] = (IdxT)i;
        clevel[i] = 0;
        markgen[i] = 0;
        tparent[i] = NONE32;
        tlevel[i] = 0;
        tscore[i] = 0.f;
    }
    for (int i = lane; i < NBINS; i += 32) {
        bhc[i] = N16;
        btc[i] = N16;
        bho[i] = 0;
        btf[i] = 0;
    }
    if (lane < 8) occ[lane] = 0;
    __syncwarp();
    for (uint32_t e = lane; e < E; e += 32) {
        uint32_t u = geu[e], v = gev[e];
        SumT s = (SumT)gsum[e];
        uint32_t c = gcnt[e];
        eu[e] = (IdxT)u;
        ev[e] = (IdxT)v;
        esum[e] = s;
        ecnt[e] = c;
        escore[e] = edge_score<U8>((unsigned long long)s, c);
        etd[e] = 0;
    }
    __syncwarp();
    // incidence lists (order is irrelevant): lanes that share a node chain their half-edges, the group
    // leader hooks the chain in front of the node's list
    for (uint32_t e0 = 0; e0 < E; e0 += 32) {
        const uint32_t e = e0 + lane;
        const bool v = e < E;
        const unsigned act = __ballot_sync(FULL, v);
#pragma unroll
        for (int side = 0; side < 2; side++) {
            uint32_t node = 0;
            unsigned peers = 0;
            if (v) {
                node = side ? ev[e] : eu[e];
                peers = __match_any_sync(act, node);
                const unsigned higher = peers & ~((2u << lane) - 1u);
                const uint32_t nxt = higher ? 2 * (e0 + (__ffs(higher) - 1)) + side : (uint32_t)ahead[node];
                anext[2 * e + side] = (IdxT)nxt;
            }
            __syncwarp();   // the old heads are read before any leader replaces them
            if (v && lane == __ffs(peers) - 1) ahead[node] = (IdxT)(2 * e + side);
            __syncwarp();
        }
    }
#ifdef BS_TRACE
    if (lane == 0) printf("[agg %d] loaded E=%u nc=%u\n", bi, E, nc);
#endif

    // ---- initial queue: edges in creation order (waterz mergeUntil first call), placed by a counting pass
    // pass 1: entries per bin (in btf)
    for (uint32_t e0 = 0; e0 < E; e0 += 32) {
        uint32_t e = e0 + lane;
        bool v = e < E;
        int bin = v ? score_bin(escore[e], NBINS) : -1;
        unsigned act = __ballot_sync(FULL, v);
        if (v) {
            unsigned peers = __match_any_sync(act, bin);
            if (lane == __ffs(peers) - 1) btf[bin] = (IdxT)(btf[bin] + __popc(peers));
        }
        __syncwarp();
    }
    uint32_t q_bump = 0, q_free = N16;
    {
        // chunk ranges: lane l owns bins [8l, 8l + 8)
        uint32_t cnt[8], nch[8], loc = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            cnt[j] = btf[lane * 8 + j];
            nch[j] = (cnt[j] + QCH - 1) / QCH;
            loc += nch[j];
        }
        uint32_t incl = loc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        q_bump = __shfl_sync(FULL, incl, 31);
        uint32_t cs = incl - loc;
        uint32_t bits = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            int bin = lane * 8 + j;
            if (cnt[j]) {
                bhc[bin] = (IdxT)cs;
                btc[bin] = (IdxT)(cs + nch[j] - 1);
                btf[bin] = (IdxT)(cnt[j] - QCH * (nch[j] - 1));
                for (uint32_t c = 0; c < nch[j]; c++) qcnext[cs + c] = (IdxT)(c + 1 < nch[j] ? cs + c + 1 : N16);
                bits |= 1u << j;
            }
            cs += nch[j];
        }
        // occupancy word w collects lanes 4w .. 4w+3
        uint32_t b0 = __shfl_sync(FULL, bits, (lane & 7) * 4 + 0), b1 = __shfl_sync(FULL, bits, (lane & 7) * 4 + 1),
                 b2 = __shfl_sync(FULL, bits, (lane & 7) * 4 + 2), b3 = __shfl_sync(FULL, bits, (lane & 7) * 4 + 3);
        if (lane < 8) occ[lane] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    }
    bool fail = q_bump > QC;
    __syncwarp();
    // pass 2: placement (running fill per bin in bho)
    if (!fail)
        for (uint32_t e0 = 0; e0 < E; e0 += 32) {
            uint32_t e = e0 + lane;
            bool v = e < E;
            int bin = v ? score_bin(escore[e], NBINS) : -1;
            unsigned act = __ballot_sync(FULL, v);
            unsigned peers = 0;
            uint32_t base = 0;
            if (v) {
                peers = __match_any_sync(act, bin);
                base = bho[bin];
            }
            __syncwarp();
            if (v) {
                uint32_t p = base + __popc(peers & lanemask_lt());
                qent[((uint32_t)bhc[bin] + p / QCH) * QCH + (p % QCH)] = (IdxT)e;
                if (lane == __ffs(peers) - 1) bho[bin] = (IdxT)(base + __popc(peers));
            }
            __syncwarp();
        }
    for (int i = lane; i < NBINS; i += 32) bho[i] = 0;
    __syncwarp();

    uint32_t n_pops = 0, n_stale = 0, n_dead = 0, n_iter = 0, n_chunk = 0, n_append = 0;
#ifdef BS_TRACE
    if (lane == 0) printf("[agg %d] queue ready chunks=%u fail=%d\n", bi, q_bump, (int)fail);
#endif

    auto alloc_chunk = [&]() -> uint32_t {
        uint32_t c;
        if (q_free != N16) {
            c = q_free;
            q_free = qcnext[c];
        } else {
            c = q_bump++;
            if (c >= QC) {
                fail = true;
                c = 0;
            }
        }
        return c;
    };
    // order-preserving (lane order) append of edge `e` to bin `bin` for lanes with `valid` (at most QCH lanes)
    auto bin_append = [&](bool valid, int bin, uint32_t e) {
        for (;;) {
            int mine = valid ? bin : 0x7fffffff;
            int Bn = __reduce_min_sync(FULL, mine);
            if (Bn == 0x7fffffff) break;
            n_append++;
            bool c = valid && bin == Bn;
            unsigned m = __ballot_sync(FULL, c);
            uint32_t total = __popc(m), off = __popc(m & lanemask_lt());
            uint32_t tc = btc[Bn];
            uint32_t tf = tc == N16 ? (uint32_t)QCH : (uint32_t)btf[Bn];
            bool need_new = tf + total > (uint32_t)QCH;
            uint32_t newc = N16;
            if (need_new) newc = alloc_chunk();
            AGG_CHK(!need_new || newc < QC, "new chunk");
            AGG_CHK(tc == N16 || tc < QC, "tail chunk");
            __syncwarp();
            if (c) {
                uint32_t pos = tf + off;
                if (pos < (uint32_t)QCH)
                    qent[tc * QCH + pos] = (IdxT)e;
                else
                    qent[newc * QCH + (pos - QCH)] = (IdxT)e;
                valid = false;
            }
            if (lane == 0) {
                if (need_new) {
                    qcnext[newc] = N16;
                    if (tc != N16)
                        qcnext[tc] = (IdxT)newc;
                    else {
                        bhc[Bn] = (IdxT)newc;
                        bho[Bn] = 0;
                        occ[Bn >> 5] |= 1u << (Bn & 31);
                    }
                    btc[Bn] = (IdxT)newc;
                    btf[Bn] = (IdxT)(tf + total - QCH);
                } else {
                    btf[Bn] = (IdxT)(tf + total);
                }
            }
            __syncwarp();
        }
    };
    // walk an incidence list: dead entries are unlinked, live ones handed to proc() 32 at a time
    // (`h` = half-edge, valid for the first `cnt` lanes).  head / tail are updated in place.
    auto walk = [&](uint32_t &head, uint32_t &tail, auto proc) {
        uint32_t h = head, prev = N16, guard = 0;
        for (;;) {
            int cnt = 0;
            uint32_t mineh = N16;
            while (h != N16 && cnt < 32) {
                if (++guard > 2 * Ecap) {   // a list can never hold more than the 2E half-edges
                    fail = true;
                    h = N16;
                    break;
                }
                AGG_CHK(h < 2 * E, "half-edge");
                if (h >= 2 * E) {
                    h = N16;
                    break;
                }
                uint32_t nx = anext[h];
                bool dead = (etd[h >> 1] & DEADBIT) != 0;
                if (dead) {
                    if (prev == N16)
                        head = nx;
                    else if (lane == 0)
                        anext[prev] = (IdxT)nx;
                } else {
                    if (lane == cnt) mineh = h;
                    cnt++;
                    prev = h;
                }
                h = nx;
            }
            if (cnt == 0) break;
            n_chunk++;
            __syncwarp();
            proc(mineh);
            __syncwarp();
            if (h == N16) break;
        }
        tail = prev;
    };

    uint32_t clock = 0, nmerge = 0;
    int minbin = 0;
    while (!fail) {
        // ---- lowest non-empty bin >= minbin
        uint32_t w = lane < 8 ? occ[lane] : 0u;
        if (lane == (minbin >> 5))
            w &= ~((1u << (minbin & 31)) - 1u);
        else if (lane < (minbin >> 5))
            w = 0;
        unsigned nzb = __ballot_sync(FULL, w != 0);
        if (!nzb) break;
        const int wl = __ffs(nzb) - 1;
        const uint32_t ww = __shfl_sync(FULL, w, wl);
        const int cb = wl * 32 + __ffs(ww) - 1;
        minbin = cb;
        if (++n_iter > 64u * E + 4096u) {   // every step pops an entry; entries are re-queued at most once per merge
            fail = true;
            break;
        }
        const uint32_t hc = bhc[cb], ho = bho[cb], tc = btc[cb], tf = btf[cb];
        const uint32_t k = (hc == tc ? tf : (uint32_t)QCH) - ho;
        const bool act = (uint32_t)lane < k;
        uint32_t e = 0, ru = 0, rv = 0;
        int cls = 1;   // 0 stop, 1 dead/inactive, 2 stale, 3 merge
        float newsc = 0.f;
        int nbin = 0;
        if (act) {
            AGG_CHK(hc < QC && ho + lane < (uint32_t)QCH, "head chunk");
            e = qent[hc * QCH + ho + lane];
            AGG_CHK(e < E, "queue entry");
            if (e >= E) e = 0;
            float sc = escore[e];
            uint32_t td = etd[e];
            if (sc >= threshold)
                cls = 0;
            else if (td & DEADBIT)
                cls = 1;
            else {
                ru = find16<IdxT>(ufp, eu[e]);
                rv = find16<IdxT>(ufp, ev[e]);
                AGG_CHK(ru < nc && rv < nc && ru != rv, "roots");
                if (stamp[ru] > td || stamp[rv] > td) {
                    cls = 2;
                    newsc = edge_score<U8>((unsigned long long)esum[e], ecnt[e]);
                    nbin = score_bin(newsc, NBINS);
                } else
                    cls = 3;
            }
        }
        const bool trig = act && (cls == 0 || cls == 3 || (cls == 2 && nbin < cb));
        const unsigned tb = __ballot_sync(FULL, trig);
        const int rstar = tb ? __ffs(tb) - 1 : (int)k;
        int tcls = __shfl_sync(FULL, cls, rstar & 31);
        if (!tb) tcls = -1;
        // stale entries before the trigger (and a stale trigger itself) are re-scored and re-queued
        const bool redo = act && cls == 2 && (lane < rstar || (lane == rstar && tcls == 2));
        if (redo) {
            escore[e] = newsc;
            etd[e] = (IdxT)clock;
        }
        const uint32_t consumed = (uint32_t)rstar + ((tcls == 3 || tcls == 2) ? 1u : 0u);
        n_pops += consumed;
        n_stale += __popc(__ballot_sync(FULL, redo));
        n_dead += __popc(__ballot_sync(FULL, act && cls == 1 && lane < rstar));
        __syncwarp();
        bin_append(redo, nbin, e);
        // ---- advance the head of bin cb (the append may have moved its tail)
        {
            const uint32_t ho2 = ho + consumed;
            const uint32_t tc2 = btc[cb], tf2 = btf[cb];
            bool freed = false;
            if (hc == tc2) {
                if (ho2 == tf2) {
                    freed = true;
                    if (lane == 0) {
                        bhc[cb] = N16;
                        btc[cb] = N16;
                        bho[cb] = 0;
                        btf[cb] = 0;
                        occ[cb >> 5] &= ~(1u << (cb & 31));
                    }
                } else if (lane == 0)
                    bho[cb] = (IdxT)ho2;
            } else if (ho2 == (uint32_t)QCH) {
                freed = true;
                if (lane == 0) {
                    bhc[cb] = qcnext[hc];
                    bho[cb] = 0;
                }
            } else if (lane == 0)
                bho[cb] = (IdxT)ho2;
            if (freed) {
                if (lane == 0) qcnext[hc] = (IdxT)q_free;
                q_free = hc;
            }
            __syncwarp();
        }
        if (tcls == 0) break;
        if (tcls == 2) {
            minbin = __shfl_sync(FULL, nbin, rstar);
            continue;
        }
        if (tcls != 3) continue;

        // ---- merge: edge me joins clusters a < b, a survives (waterz mergeRegions)
        const uint32_t me = __shfl_sync(FULL, e, rstar);
        const uint32_t r1 = __shfl_sync(FULL, ru, rstar), r2 = __shfl_sync(FULL, rv, rstar);
        const uint32_t a = min(r1, r2), b = max(r1, r2);
        clock++;
        const uint32_t gen = clock;
        const float msc = escore[me];
        if (lane == 0) etd[me] = (IdxT)(etd[me] | DEADBIT);
        __syncwarp();
        // pass 1: b's neighbours are marked with their edge
        uint32_t head_b = ahead[b], tail_b = N16;
        walk(head_b, tail_b, [&](uint32_t h) {
            if (h != N16) {
                uint32_t ne = h >> 1;
                uint32_t x1 = find16<IdxT>(ufp, eu[ne]), x2 = find16<IdxT>(ufp, ev[ne]);
                uint32_t x = x1 == b ? x2 : x1;
                AGG_CHK((x1 == b || x2 == b) && x < nc && x != b && x != a, "b neighbour");
                mark[x] = (IdxT)ne;
                markgen[x] = (IdxT)gen;
            }
        });
        // pass 2: a's edges to a common neighbour absorb (or are absorbed by) b's edge
        uint32_t head_a = ahead[a], tail_a = N16;
        walk(head_a, tail_a, [&](uint32_t h) {
            if (h != N16) {
                uint32_t ae = h >> 1;
                uint32_t x1 = find16<IdxT>(ufp, eu[ae]), x2 = find16<IdxT>(ufp, ev[ae]);
                uint32_t x = x1 == a ? x2 : x1;
                AGG_CHK((x1 == a || x2 == a) && x < nc && x != b && x != a, "a neighbour");
                if (markgen[x] == gen) {
                    uint32_t ne = mark[x];
                    if (!keep_cheaper || escore[ne] > escore[ae]) {
                        esum[ae] += esum[ne];
                        ecnt[ae] += ecnt[ne];
                        etd[ne] = (IdxT)(etd[ne] | DEADBIT);
                    } else {
                        esum[ne] += esum[ae];
                        ecnt[ne] += ecnt[ae];
                        etd[ae] = (IdxT)(etd[ae] | DEADBIT);
                    }
                }
            }
        });
        // splice b's list behind a's, union, merge tree
        if (lane == 0) {
            if (head_b != N16) {
                if (head_a == N16)
                    head_a = head_b;
                else
                    anext[tail_a] = (IdxT)head_b;
            }
            ahead[a] = (IdxT)head_a;
            ufp[b] = (IdxT)a;
            stamp[a] = (IdxT)clock;
            const uint32_t t = nc + nmerge, ta = tnode[a], tbn = tnode[b];
            const uint32_t lvl = max((uint32_t)clevel[a], (uint32_t)clevel[b]) + 1;
            tparent[ta] = t;
            tparent[tbn] = t;
            tparent[t] = NONE32;
            tlevel[t] = lvl;
            tscore[t] = msc;
            tnode[a] = (IdxT)t;
            clevel[a] = (IdxT)lvl;
            ha[nmerge] = a;
            hb[nmerge] = b;
            hs[nmerge] = msc;
        }
        nmerge++;
        __syncwarp();
    }
    fail = __any_sync(FULL, fail);
#ifdef BS_TRACE
    if (lane == 0) printf("[agg %d] done merges=%u iters=%u fail=%d\n", bi, nmerge, n_iter, (int)fail);
#endif
    if (lane == 0) {
        A.nmerges[bi] = nmerge;
        A.counters[6 * bi + 0] = n_pops;
        A.counters[6 * bi + 1] = n_stale;
        A.counters[6 * bi + 2] = n_dead;
        A.counters[6 * bi + 3] = n_iter;
        A.counters[6 * bi + 4] = n_chunk;
        A.counters[6 * bi + 5] = n_append;
        if (fail) atomicExch(A.error, 1u);
    }
}

int agglom_smem_launch(const AggBlk *blks, const int *list, int nlist, const AggArrays &A, float threshold, int keep_cheaper,
                       bool u8, bool sum64, uint32_t Ecap, uint32_t Ncap, cudaStream_t s) {
    if (nlist == 0) return BS_OK;
    const size_t smem = agglom_smem_bytes(Ecap, Ncap, sum64);
    BS_ARG(smem <= 227 * 1024 && Ecap <= 32760 && Ncap <= 32760 && (Ecap % 8) == 0 && (Ncap % 8) == 0,
           "agglom_smem_launch: block graph does not fit in shared memory");
#define BS_AGG_SMEM(U8_, SumT_)                                                                                              \
    do {                                                                                                                     \
        BS_CUDA(cudaFuncSetAttribute(k_agglomerate_lists<U8_, SumT_, uint16_t, true>,                                        \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));                              \
        BS_LAUNCH((k_agglomerate_lists<U8_, SumT_, uint16_t, true>), nlist, 32, smem, s, blks, list, A, threshold,           \
                  keep_cheaper, Ecap, Ncap, nullptr, nullptr);                                                               \
    } while (0)
    if (u8 && !sum64)
        BS_AGG_SMEM(true, uint32_t);
    else if (u8)
        BS_AGG_SMEM(true, unsigned long long);
    else
        BS_AGG_SMEM(false, unsigned long long);
#undef BS_AGG_SMEM
    return BS_OK;
}

// blocks too large for shared memory: the same kernel on a global-memory slab per block (32-bit indices)
int agglom_global_launch(const AggBlk *blks, const int *list, int nlist, const AggArrays &A, float threshold, int keep_cheaper,
                         bool u8, unsigned char *work, const unsigned long long *woff, cudaStream_t s) {
    if (nlist == 0) return BS_OK;
    if (u8)
        BS_LAUNCH((k_agglomerate_lists<true, unsigned long long, uint32_t, false>), nlist, 32, 0, s, blks, list, A, threshold,
                  keep_cheaper, 0u, 0u, work, woff);
    else
        BS_LAUNCH((k_agglomerate_lists<false, unsigned long long, uint32_t, false>), nlist, 32, 0, s, blks, list, A, threshold,
                  keep_cheaper, 0u, 0u, work, woff);
    return BS_OK;
}

}  // namespace bs
