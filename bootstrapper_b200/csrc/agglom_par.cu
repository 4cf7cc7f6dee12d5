// waterz BinQueue<256> agglomeration of one daisy block by a CTA of PNW warps that commits several merges per round.
//
// Same result as a single warp popping one entry at a time (the kernel of round 1, since removed), to the bit: the pop order of the FIFO bin queue is
// emulated exactly, but a whole batch of queue entries is resolved per round and the merges found in it run
// concurrently, one warp each.  A prefix of the batch may commit together when processing it entry by entry would
// give the same state:
//   * an entry (stale or merge) that shares a cluster with an earlier merge of the batch would see a changed edge
//     (re-keyed, re-scored or deleted)                                          -> the batch is cut before it;
//   * two merges whose clusters are joined by an edge would both rewrite that edge -> the later one waits;
//   * everything else commutes: merges touch only edges incident to their own two clusters, stale entries only
//     their own edge, and clocks / merge numbers are assigned in queue order.
// Round = A (warp 0: pick the lowest bin, classify up to PQCH entries, cut at the first stop / re-queue into a lower
// bin / cluster conflict)  ->  B (one warp per candidate merge: read-only walk of both incidence lists, neighbours of
// the absorbed cluster into a warp-private hash, adjacency conflicts with the other candidates)  ->  C (warp 0:
// commit the stale / dead entries of the final prefix, advance the queue)  ||  D (candidate warps: resolve common
// neighbours against the private hash, splice lists, union, merge-tree node).
// A merge whose absorbed cluster has more than PBIG neighbours runs alone on the node-mark arrays.
#include <type_traits>

#include "agglom.cuh"

namespace bs {

static constexpr uint32_t NONE32 = 0xFFFFFFFFu;
static constexpr unsigned FULL = 0xFFFFFFFFu;
static constexpr int PQCH = 16;      // entries per queue chunk = batch size of a round
static constexpr int PNW = 16;       // warps per CTA
static constexpr int PCAND = PNW / 2; // merges per round: warp c walks the absorbed cluster's list, warp PCAND + c the survivor's
static constexpr int PHASH = 128;    // slots of a warp's private neighbour hash
static constexpr int PFILT = 64;     // slots of the round's cluster -> candidate table (at most 2 * PCAND entries)
static constexpr int PBIG = 80;      // neighbours of the absorbed cluster the private hash takes
static constexpr int NBINS = 256;
static_assert(PQCH == 16, "the conflict match packs the two endpoint slots of a 16-entry batch into one warp");

__host__ __device__ static inline uint32_t par_qc(uint32_t Ecap) { return Ecap / PQCH + 2 * NBINS + 32; }

struct ParCtl {
    uint32_t ncand, cut1, cut2, clock0, nmerge0;
    int exit_, fail;
    uint32_t ce[PCAND], ca[PCAND], cb[PCAND], clane[PCAND], headb[PCAND], big[PCAND];
    uint32_t hkey[PCAND][PHASH], hval[PCAND][PHASH];
    uint32_t fkey[PFILT], fval[PFILT];   // cluster -> candidate index of this round's merges (open addressing)
};

size_t agglom_par_bytes(uint32_t Ecap, uint32_t Ncap, bool sum64, int idx) {
    size_t b = 0;
    b += (size_t)(sum64 ? 8 : 4) * Ecap;   // esum
    b += 4 * (size_t)Ecap * 2;             // escore, ecnt
    b += 32;                               // occupancy bitmap
    b += (size_t)idx * Ecap * 5;           // eu, ev, etd, anext[2E]
    const size_t QC = par_qc(Ecap);
    b += (size_t)idx * (QC * PQCH + QC);   // queue chunks + chunk links
    b += (size_t)idx * Ncap * 7;           // ufp, stamp, ahead, tnode, clevel, mark, markgen
    b += (size_t)idx * NBINS * 4;          // bin head chunk / tail chunk / head offset / tail fill
    return (b + 255) & ~(size_t)255;
}
size_t agglom_par_static_smem() { return sizeof(ParCtl) + 64; }
// hybrid layout of the global-slab kernel: queue bins, occupancy bits, union-find parents and cluster stamps (16 bit)
// stay in shared memory -- they sit on every dependent-load chain of a round -- and only the edge / list arrays live in
// the slab
static size_t agglom_par_hyb_bytes(uint32_t Ncap) { return 32 + 4 * (size_t)NBINS * 4 + 2 * (size_t)Ncap * 2; }

template <typename IdxT>
__device__ __forceinline__ uint32_t pfind(IdxT *ufp, uint32_t x) {
    // path halving; concurrent lanes / warps only ever write ancestors
    for (;;) {
        uint32_t p = ufp[x];
        if (p == x) return x;
        uint32_t gp = ufp[p];
        if (gp == p) return p;
        ufp[x] = (IdxT)gp;
        x = gp;
    }
}

template <bool U8, typename SumT, typename IdxT, bool SMEM, bool HYB>
__global__ void __launch_bounds__(32 * PNW) k_agglomerate_par(const AggBlk *__restrict__ blks, const int *__restrict__ list,
                                                              AggArrays A, float threshold, int keep_cheaper, uint32_t Ecap_,
                                                              uint32_t Ncap_, unsigned char *__restrict__ gwork,
                                                              const unsigned long long *__restrict__ gwoff) {
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ ParCtl C;
    static_assert(!(SMEM && HYB), "the hybrid layout belongs to the global-slab kernel");
    using NodT = typename std::conditional<HYB, uint16_t, IdxT>::type;   // union-find parents, cluster stamps
    constexpr uint32_t N16 = (uint32_t)(IdxT)~(IdxT)0;                 // "none" in the index type
    constexpr uint32_t DEADBIT = 1u << (8 * sizeof(IdxT) - 1);        // etd = time | dead flag
    const int bi = list[blockIdx.x];
    const AggBlk B = blks[bi];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t E = B.E, nc = B.nv;
    const uint32_t Ecap = SMEM ? Ecap_ : ((max(E, 8u) + 7u) & ~7u), Ncap = SMEM ? Ncap_ : ((max(nc, 8u) + 7u) & ~7u);
    const uint32_t QC = par_qc(Ecap);
    unsigned char *work = SMEM ? smraw : gwork + gwoff[blockIdx.x];

    SumT *esum = (SumT *)work;
    float *escore = (float *)(esum + Ecap);
    uint32_t *ecnt = (uint32_t *)(escore + Ecap);
    uint32_t *occ = HYB ? (uint32_t *)smraw : ecnt + Ecap;
    IdxT *eu = (IdxT *)(ecnt + Ecap + 8);
    IdxT *ev = eu + Ecap, *etd = ev + Ecap, *anext = etd + Ecap;
    IdxT *qent = anext + 2 * Ecap;
    IdxT *qcnext = qent + QC * PQCH;
    IdxT *nodes0 = qcnext + QC;
    IdxT *ahead = nodes0 + 2 * Ncap, *tnode = ahead + Ncap, *clevel = tnode + Ncap, *mark = clevel + Ncap, *markgen = mark + Ncap;
    IdxT *bhc = HYB ? (IdxT *)(smraw + 32) : markgen + Ncap;
    IdxT *btc = bhc + NBINS, *bho = btc + NBINS, *btf = bho + NBINS;
    NodT *ufp = HYB ? (NodT *)(smraw + 32 + 4 * NBINS * sizeof(IdxT)) : (NodT *)nodes0;
    NodT *stamp = ufp + Ncap;

    const uint32_t *geu = A.eu + B.ebase, *gev = A.ev + B.ebase, *gcnt = A.ecnt + B.ebase;
    const unsigned long long *gsum = A.esum + B.ebase;
    uint32_t *tparent = A.tparent + 2 * (size_t)B.vbase, *tlevel = A.tlevel + 2 * (size_t)B.vbase;
    float *tscore = A.tscore + 2 * (size_t)B.vbase;
    uint32_t *ha = A.ha + B.vbase, *hb = A.hb + B.vbase;
    float *hs = A.hs + B.vbase;

    // ---- load the block's graph (all warps)
    for (uint32_t i = threadIdx.x; i < nc; i += blockDim.x) {
        ufp[i] = (NodT)i;
        stamp[i] = 0;
        ahead[i] = N16;
        tnode[i] = (IdxT)i;
        clevel[i] = 0;
        markgen[i] = 0;
        tparent[i] = NONE32;
        tlevel[i] = 0;
        tscore[i] = 0.f;
    }
    for (int i = threadIdx.x; i < NBINS; i += blockDim.x) {
        bhc[i] = N16;
        btc[i] = N16;
        bho[i] = 0;
        btf[i] = 0;
    }
    if (threadIdx.x < 8) occ[threadIdx.x] = 0;
    for (uint32_t e = threadIdx.x; e < E; e += blockDim.x) {
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
    if (threadIdx.x == 0) {
        C.exit_ = 0;
        C.fail = 0;
        C.ncand = 0;
    }
    __syncthreads();

    // warp 0 owns the queue; its replicated (uniform) state lives in registers
    uint32_t q_bump = 0, q_free = N16;
    bool fail = false;
    if (warp == 0) {
        // incidence lists (order is irrelevant): lanes that share a node chain their half-edges
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
        // initial queue: edges in creation order, placed by a counting pass
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
        {
            uint32_t cnt[8], nch[8], loc = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                cnt[j] = btf[lane * 8 + j];
                nch[j] = (cnt[j] + PQCH - 1) / PQCH;
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
                    btf[bin] = (IdxT)(cnt[j] - PQCH * (nch[j] - 1));
                    for (uint32_t c = 0; c < nch[j]; c++) qcnext[cs + c] = (IdxT)(c + 1 < nch[j] ? cs + c + 1 : N16);
                    bits |= 1u << j;
                }
                cs += nch[j];
            }
            uint32_t b0 = __shfl_sync(FULL, bits, (lane & 7) * 4 + 0), b1 = __shfl_sync(FULL, bits, (lane & 7) * 4 + 1),
                     b2 = __shfl_sync(FULL, bits, (lane & 7) * 4 + 2), b3 = __shfl_sync(FULL, bits, (lane & 7) * 4 + 3);
            if (lane < 8) occ[lane] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
        }
        fail = q_bump > QC;
        __syncwarp();
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
                    qent[((uint32_t)bhc[bin] + p / PQCH) * PQCH + (p % PQCH)] = (IdxT)e;
                    if (lane == __ffs(peers) - 1) bho[bin] = (IdxT)(base + __popc(peers));
                }
                __syncwarp();
            }
        for (int i = lane; i < NBINS; i += 32) bho[i] = 0;
        __syncwarp();
    }
    __syncthreads();

    uint32_t n_pops = 0, n_stale = 0, n_dead = 0, n_iter = 0, n_chunk = 0, n_append = 0;
#if defined(BS_PROBE) && BS_PROBE == 2
    uint32_t n_c2 = 0, n_full = 0, n_short = 0, n_low = 0, n_conf = 0;
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
    // order-preserving (lane order) append of edge `e` to bin `bin` for lanes with `valid` (at most PQCH lanes); warp 0
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
            uint32_t tf = tc == N16 ? (uint32_t)PQCH : (uint32_t)btf[Bn];
            bool need_new = tf + total > (uint32_t)PQCH;
            uint32_t newc = N16;
            if (need_new) newc = alloc_chunk();
            __syncwarp();
            if (c) {
                uint32_t pos = tf + off;
                if (pos < (uint32_t)PQCH)
                    qent[tc * PQCH + pos] = (IdxT)e;
                else
                    qent[newc * PQCH + (pos - PQCH)] = (IdxT)e;
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
                    btf[Bn] = (IdxT)(tf + total - PQCH);
                } else {
                    btf[Bn] = (IdxT)(tf + total);
                }
            }
            __syncwarp();
        }
    };
    // walk an incidence list (one warp): dead entries are unlinked, live ones handed to proc() 32 at a time
    auto walk = [&](uint32_t &head, uint32_t &tail, auto proc) {
        uint32_t h = head, prev = N16, guard = 0;
        for (;;) {
            int cnt = 0;
            uint32_t mineh = N16;
            while (h != N16 && cnt < 32) {
                if (++guard > 2 * Ecap || h >= 2 * E) {   // a list can never hold more than the 2E half-edges
                    fail = true;
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
    // warp 0: the next round's batch -- lowest occupied bin from `minbin` on and up to PQCH entry ids (the rest of its head
    // chunk and, if the bin goes on, the start of the next chunk).  Runs at the end of phase C, so the loads of the edge
    // ids overlap the merges of phase D, which never touch the queue.
    uint32_t e = 0, k = 0, hc = 0, ho = 0, nxc = N16;
    int cb = 0;
    bool act = false, have = false;
    auto select = [&]() {
        uint32_t w = lane < 8 ? occ[lane] : 0u;
        if (lane == (minbin >> 5))
            w &= ~((1u << (minbin & 31)) - 1u);
        else if (lane < (minbin >> 5))
            w = 0;
        const unsigned nzb = __ballot_sync(FULL, w != 0);
        have = nzb != 0;
        act = false;
        if (!have) return;
        const int wl = __ffs(nzb) - 1;
        const uint32_t ww = __shfl_sync(FULL, w, wl);
        cb = wl * 32 + __ffs(ww) - 1;
        minbin = cb;
        hc = bhc[cb];
        ho = bho[cb];
        const uint32_t tc = btc[cb], tf = btf[cb];
        const uint32_t k1 = (hc == tc ? tf : (uint32_t)PQCH) - ho;
        nxc = hc == tc ? N16 : (uint32_t)qcnext[hc];
        const uint32_t k2 = hc == tc ? 0u : min((uint32_t)PQCH - k1, nxc == tc ? tf : (uint32_t)PQCH);
        k = k1 + k2;
        act = (uint32_t)lane < k;
        e = act ? ((uint32_t)lane < k1 ? (uint32_t)qent[hc * PQCH + ho + lane] : (uint32_t)qent[nxc * PQCH + (lane - k1)]) : 0u;
    };
    if (warp == 0) select();
#ifdef BS_PROBE
    long long tA = 0, tB = 0, tC = 0, tD = 0, t0 = clock64();
#endif
    for (;;) {
        // ================= phase A (warp 0): batch, classification, first cut, candidate merges
        uint32_t ru = 0, rv = 0;
        int cls = 1;   // 0 stop, 1 dead/inactive, 2 stale, 3 merge
        float newsc = 0.f;
        int nbin = 0;
        unsigned mbits = 0;
        if (warp == 0) {
            if (!have || fail || n_iter > 64u * E + 4096u) {
                if (have) fail = true;
                if (lane == 0) {
                    C.exit_ = 1;
                    C.ncand = 0;
                    C.cut1 = 0;
                    C.cut2 = NONE32;
                }
            } else {
                n_iter++;
                if (act) {
                    // every field of the entry in one round trip (the slab is L2 / HBM resident)
                    const float sc = escore[e];
                    const uint32_t td = etd[e], u0 = eu[e], v0 = ev[e], c0 = ecnt[e];
                    const SumT s0 = esum[e];
                    if (sc >= threshold)
                        cls = 0;
                    else if (td & DEADBIT)
                        cls = 1;
                    else {
                        ru = pfind<NodT>(ufp, u0);
                        rv = pfind<NodT>(ufp, v0);
                        if (stamp[ru] > td || stamp[rv] > td) {
                            cls = 2;
                            newsc = edge_score<U8>((unsigned long long)s0, c0);
                            nbin = score_bin(newsc, NBINS);
                        } else
                            cls = 3;
                    }
                }
                // an entry that shares a cluster with an earlier merge of the batch must wait for it.  One match over 32
                // slots: slot i holds ru of entry i, slot 16 + i its rv (a batch has at most PQCH = 16 entries).
                bool conflict = false;
                {
                    const uint32_t rv_up = __shfl_sync(FULL, rv, lane & 15);
                    const int cls_up = __shfl_sync(FULL, cls, lane & 15);
                    const bool act_up = __shfl_sync(FULL, (int)act, lane & 15) != 0;
                    const bool slot_ok = lane < 16 ? (act && (cls == 2 || cls == 3)) : (act_up && (cls_up == 2 || cls_up == 3));
                    const uint32_t key = slot_ok ? (lane < 16 ? ru : rv_up) : (0xFFFF0000u | (uint32_t)lane);
                    const unsigned mm = __match_any_sync(FULL, key);
                    const unsigned mv = __shfl_sync(FULL, mm, (lane + 16) & 31);
                    const unsigned m16 = __ballot_sync(FULL, act && cls == 3) & 0xFFFFu;
                    const unsigned lower = (1u << (lane & 15)) - 1u;
                    conflict = lane < 16 && slot_ok && ((mm | mv) & (m16 | (m16 << 16)) & (lower | (lower << 16))) != 0;
                }
                mbits = __ballot_sync(FULL, act && cls == 3);
                const bool over = act && cls == 3 && __popc(mbits & lanemask_lt()) >= PCAND;
                const bool cutf = act && (cls == 0 || (cls == 2 && nbin < cb) || conflict || over);
                const unsigned cbits = __ballot_sync(FULL, cutf);
                const uint32_t cut1 = cbits ? (uint32_t)(__ffs(cbits) - 1) : k;
                const bool cand = act && cls == 3 && (uint32_t)lane < cut1;
                const unsigned candbits = __ballot_sync(FULL, cand);
                C.fkey[lane] = NONE32;
                C.fkey[lane + 32] = NONE32;
                __syncwarp();
                if (cand) {
                    const int r = __popc(candbits & lanemask_lt());
                    C.ce[r] = e;
                    C.ca[r] = min(ru, rv);
                    C.cb[r] = max(ru, rv);
                    C.clane[r] = lane;
#pragma unroll
                    for (int side = 0; side < 2; side++) {
                        const uint32_t node = side ? rv : ru;
                        uint32_t sl = (node * 2654435761u) >> 26;   // 6 bits
                        for (;;) {
                            if (atomicCAS(&C.fkey[sl], NONE32, node) == NONE32) {
                                C.fval[sl] = (uint32_t)r;
                                break;
                            }
                            sl = (sl + 1) & (PFILT - 1);
                        }
                    }
                }
                if (lane == 0) {
                    C.ncand = __popc(candbits);
                    C.cut1 = cut1;
                    C.cut2 = NONE32;
                    C.clock0 = clock;
                    C.nmerge0 = nmerge;
                }
                // what stopped the batch (meaningful when the final cut stays at cut1)
                const bool conf_at = __shfl_sync(FULL, (int)(conflict || over), cut1 & 31) != 0;
                // keep for phase C: class at cut1 -- 0 stop, 2 re-queue into a lower bin, 4 wait, -1 batch exhausted
                const int c1 = __shfl_sync(FULL, cls, cut1 & 31);
                cls = (cls & 0xff) | ((cut1 < k ? (conf_at ? 4 : c1) : 0xff) << 8) | ((int)cut1 << 16);
            }
        }
        __syncthreads();
#ifdef BS_PROBE
        { long long t1 = clock64(); tA += t1 - t0; t0 = t1; }
#endif
        // ================= phase B: read-only walks (warp c: absorbed cluster b -> private neighbour hash; warp PCAND + c:
        // surviving cluster a, first 32 entries kept in registers), adjacency conflicts with the other candidates
        const uint32_t ncand = C.ncand;
        const int cidx = warp < PCAND ? warp : warp - PCAND;
        const bool bwalker = warp < PCAND && (uint32_t)cidx < ncand, awalker = warp >= PCAND && (uint32_t)cidx < ncand;
        uint32_t ma = 0, mb = 0, me = 0, head_a = N16, tail_a = N16, sa_h = N16, sa_x = 0, nbat = 0;
        uint32_t *hkey = C.hkey[cidx], *hval = C.hval[cidx];
        auto other_cand = [&](uint32_t x) {
            // is x a cluster of another merge of this round?  (the clusters of a round's merges are pairwise distinct)
            uint32_t sl = (x * 2654435761u) >> 26;
            for (;;) {
                const uint32_t kx = C.fkey[sl];
                if (kx == NONE32) return;
                if (kx == x) {
                    const uint32_t c = C.fval[sl];
                    if (c != (uint32_t)cidx) atomicMin(&C.cut2, C.clane[max(c, (uint32_t)cidx)]);
                    return;
                }
                sl = (sl + 1) & (PFILT - 1);
            }
        };
        if (bwalker || awalker) ma = C.ca[cidx], mb = C.cb[cidx], me = C.ce[cidx];
        if (bwalker) {
            for (int j = lane; j < PHASH; j += 32) hkey[j] = NONE32;
            __syncwarp();
            uint32_t deg = 0, head_b = ahead[mb], tail_b = N16;
            bool big = false;
            walk(head_b, tail_b, [&](uint32_t h) {
                bool ins = false;
                uint32_t x = 0, ne = 0;
                if (h != N16) {
                    ne = h >> 1;
                    if (ne != me) {
                        uint32_t x1 = pfind<NodT>(ufp, eu[ne]), x2 = pfind<NodT>(ufp, ev[ne]);
                        x = x1 == mb ? x2 : x1;
                        other_cand(x);
                        ins = true;
                    }
                }
                const unsigned ib = __ballot_sync(FULL, ins);
                deg += __popc(ib);
                if (deg > (uint32_t)PBIG) big = true;
                if (ins && !big) {
                    uint32_t s = (x * 2654435761u) >> 25;   // 7 bits
                    for (;;) {
                        uint32_t old = atomicCAS(&hkey[s], NONE32, x);
                        if (old == NONE32) {
                            hval[s] = ne;
                            break;
                        }
                        s = (s + 1) & (PHASH - 1);
                    }
                }
            });
            if (lane == 0) {
                ahead[mb] = (IdxT)head_b;   // dead entries at the front may have been unlinked
                C.headb[cidx] = head_b;
                C.big[cidx] = big ? 1u : 0u;
                if (big) {
                    // runs alone: first candidate -> everybody else waits, otherwise it waits itself
                    if (cidx == 0) {
                        if (ncand > 1) atomicMin(&C.cut2, C.clane[1]);
                    } else
                        atomicMin(&C.cut2, C.clane[cidx]);
                }
                if (fail) C.fail = 1;
            }
        }
        if (awalker) {
            head_a = ahead[ma];
            walk(head_a, tail_a, [&](uint32_t h) {
                uint32_t x = 0;
                if (h != N16) {
                    uint32_t ae = h >> 1;
                    if (ae != me) {
                        uint32_t x1 = pfind<NodT>(ufp, eu[ae]), x2 = pfind<NodT>(ufp, ev[ae]);
                        x = x1 == ma ? x2 : x1;
                        other_cand(x);
                    } else
                        h = N16;
                }
                if (nbat == 0) sa_h = h, sa_x = x;
                nbat++;
            });
            if (lane == 0) {
                ahead[ma] = (IdxT)head_a;
                if (fail) C.fail = 1;
            }
        }
        __syncthreads();
#ifdef BS_PROBE
        { long long t1 = clock64(); tB += t1 - t0; t0 = t1; }
#endif
        // ================= phase C (warp 0): commit the non-merge entries of the final prefix, advance the queue
        const uint32_t cut = min(C.cut1, C.cut2);
        if (warp == 0 && !C.exit_) {
            const int mycls = cls & 0xff;
            const uint32_t cut1 = (uint32_t)(cls >> 16);
            int tcls = (cls >> 8) & 0xff;   // 0 stop, 2 lower, 4 wait, 0xff exhausted
#if defined(BS_PROBE) && BS_PROBE == 2
            if (cut < cut1) n_c2++;
            else if (tcls == 0xff) { if (k == PQCH) n_full++; else n_short++; }
            else if (tcls == 2) n_low++;
            else if (tcls == 4) n_conf++;
#endif
            if (cut < cut1) tcls = 4;
            const bool lower_trig = tcls == 2;
            const bool redo = act && mycls == 2 && ((uint32_t)lane < cut || ((uint32_t)lane == cut && lower_trig));
            if (redo) {
                escore[e] = newsc;
                etd[e] = (IdxT)(clock + __popc(mbits & lanemask_lt()));   // merges committed before this pop
            }
            const uint32_t consumed = cut + (lower_trig ? 1u : 0u);
            const unsigned mcommit = mbits & ((cut >= 32 ? 0u : (1u << cut)) - 1u);
            n_pops += consumed;
            n_stale += __popc(__ballot_sync(FULL, redo));
            n_dead += __popc(__ballot_sync(FULL, act && mycls == 1 && (uint32_t)lane < cut));
            __syncwarp();
            bin_append(redo, nbin, e);
            {
                // the head moves by `consumed` entries, possibly into the next chunk (the append may have grown the tail)
                uint32_t hcur = hc, ho2 = ho + consumed;
                if (ho2 >= (uint32_t)PQCH && hcur != (uint32_t)btc[cb]) {
                    const uint32_t nx2 = nxc != N16 ? nxc : (uint32_t)qcnext[hcur];   // the append may have linked a new chunk
                    __syncwarp();
                    if (lane == 0) qcnext[hcur] = (IdxT)q_free;
                    q_free = hcur;
                    hcur = nx2;
                    ho2 -= PQCH;
                }
                __syncwarp();
                const uint32_t tc2 = btc[cb], tf2 = btf[cb];
                if (hcur == tc2 && ho2 == tf2) {
                    // bin exhausted
                    if (lane == 0) {
                        qcnext[hcur] = (IdxT)q_free;
                        bhc[cb] = N16;
                        btc[cb] = N16;
                        bho[cb] = 0;
                        btf[cb] = 0;
                        occ[cb >> 5] &= ~(1u << (cb & 31));
                    }
                    q_free = hcur;
                } else if (lane == 0) {
                    bhc[cb] = (IdxT)hcur;
                    bho[cb] = (IdxT)ho2;
                }
                __syncwarp();
            }
            if (lower_trig) minbin = __shfl_sync(FULL, nbin, cut & 31);
            const uint32_t nm = __popc(mcommit);
            clock += nm;
            nmerge += nm;
            if (tcls == 0 || fail) {
                if (lane == 0) C.exit_ = 1;   // stop after this round's merges
            } else
                select();
        }
#ifdef BS_PROBE
        { long long t1 = clock64(); tC += t1 - t0; t0 = t1; }
#endif
        // ================= phase D (the a-walkers of candidates before the final cut): the merges
        if (awalker && C.clane[cidx] < cut) {
            const uint32_t myclock = C.clock0 + cidx + 1, mynum = C.nmerge0 + cidx;
            const float msc = escore[me];
            const bool big = C.big[cidx] != 0;
            uint32_t head_b = C.headb[cidx];
            if (lane == 0) etd[me] = (IdxT)(etd[me] | DEADBIT);
            __syncwarp();
            auto resolve = [&](uint32_t ae, uint32_t x) {
                uint32_t ne = NONE32;
                if (big) {
                    if (markgen[x] == (IdxT)myclock) ne = mark[x];
                } else {
                    uint32_t s = (x * 2654435761u) >> 25;
                    for (;;) {
                        uint32_t kx = hkey[s];
                        if (kx == x) {
                            ne = hval[s];
                            break;
                        }
                        if (kx == NONE32) break;
                        s = (s + 1) & (PHASH - 1);
                    }
                }
                if (ne != NONE32) {
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
            };
            if (big) {
                // alone in this round: neighbours of the absorbed cluster go through the node marks
                uint32_t tail_b = N16;
                head_b = ahead[mb];
                walk(head_b, tail_b, [&](uint32_t h) {
                    if (h != N16) {
                        uint32_t ne = h >> 1;
                        uint32_t x1 = pfind<NodT>(ufp, eu[ne]), x2 = pfind<NodT>(ufp, ev[ne]);
                        uint32_t x = x1 == mb ? x2 : x1;
                        mark[x] = (IdxT)ne;
                        markgen[x] = (IdxT)myclock;
                    }
                });
            }
            if (nbat <= 1) {
                // the survivor's whole list sits in registers since phase B (its clusters did not change since)
                if (sa_h != N16) resolve(sa_h >> 1, sa_x);
                __syncwarp();
            } else {
                head_a = ahead[ma];
                walk(head_a, tail_a, [&](uint32_t h) {
                    if (h != N16) {
                        uint32_t ae = h >> 1;
                        uint32_t x1 = pfind<NodT>(ufp, eu[ae]), x2 = pfind<NodT>(ufp, ev[ae]);
                        resolve(ae, x1 == ma ? x2 : x1);
                    }
                });
            }
            if (lane == 0) {
                // dead entries (the merged edge among them) stay linked until a later walk meets them
                if (head_b != N16) {
                    if (head_a == N16)
                        head_a = head_b;
                    else
                        anext[tail_a] = (IdxT)head_b;
                }
                ahead[ma] = (IdxT)head_a;
                ufp[mb] = (NodT)ma;
                stamp[ma] = (NodT)myclock;
                const uint32_t t = nc + mynum, ta = tnode[ma], tbn = tnode[mb];
                const uint32_t lvl = max((uint32_t)clevel[ma], (uint32_t)clevel[mb]) + 1;
                tparent[ta] = t;
                tparent[tbn] = t;
                tparent[t] = NONE32;
                tlevel[t] = lvl;
                tscore[t] = msc;
                tnode[ma] = (IdxT)t;
                clevel[ma] = (IdxT)lvl;
                ha[mynum] = ma;
                hb[mynum] = mb;
                hs[mynum] = msc;
                if (fail) C.fail = 1;
            }
        }
        __syncthreads();
#ifdef BS_PROBE
        { long long t1 = clock64(); tD += t1 - t0; t0 = t1; }
#endif
        if (C.exit_ || C.fail) break;
    }
    if (threadIdx.x == 0) {
        A.nmerges[bi] = nmerge;
        A.counters[6 * bi + 0] = n_pops;
        A.counters[6 * bi + 1] = n_stale;
        A.counters[6 * bi + 2] = n_dead;
        A.counters[6 * bi + 3] = n_iter;
        A.counters[6 * bi + 4] = n_chunk;
        A.counters[6 * bi + 5] = n_append;
#if defined(BS_PROBE) && BS_PROBE == 2
        A.counters[6 * bi + 0] = n_c2;
        A.counters[6 * bi + 1] = n_full;
        A.counters[6 * bi + 2] = n_short;
        A.counters[6 * bi + 4] = n_low;
        A.counters[6 * bi + 5] = n_conf;
#elif defined(BS_PROBE)
        A.counters[6 * bi + 0] = (uint32_t)(tA / 16);
        A.counters[6 * bi + 1] = (uint32_t)(tB / 16);
        A.counters[6 * bi + 2] = (uint32_t)(tC / 16);
        A.counters[6 * bi + 4] = (uint32_t)(tD / 16);
#endif
        if (fail || C.fail) atomicExch(A.error, 1u);
    }
}

int agglom_par_launch(const AggBlk *blks, const int *list, int nlist, const AggArrays &A, float threshold, int keep_cheaper,
                      bool u8, bool sum64, uint32_t Ecap, uint32_t Ncap, cudaStream_t s) {
    if (nlist == 0) return BS_OK;
    const size_t smem = agglom_par_bytes(Ecap, Ncap, sum64, 2);
    BS_ARG(smem + agglom_par_static_smem() <= 227 * 1024 && Ecap <= 32760 && Ncap <= 32760 && (Ecap % 8) == 0 && (Ncap % 8) == 0,
           "agglom_par_launch: block graph does not fit in shared memory");
#define BS_AGG_PAR(U8_, SumT_)                                                                                             \
    do {                                                                                                                   \
        BS_CUDA(cudaFuncSetAttribute(k_agglomerate_par<U8_, SumT_, uint16_t, true, false>,                                 \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024 - agglom_par_static_smem()))); \
        BS_LAUNCH((k_agglomerate_par<U8_, SumT_, uint16_t, true, false>), nlist, 32 * PNW, smem, s, blks, list, A, threshold, \
                  keep_cheaper, Ecap, Ncap, nullptr, nullptr);                                                             \
    } while (0)
    if (u8 && !sum64)
        BS_AGG_PAR(true, uint32_t);
    else if (u8)
        BS_AGG_PAR(true, unsigned long long);
    else
        BS_AGG_PAR(false, unsigned long long);
#undef BS_AGG_PAR
    return BS_OK;
}

int agglom_par_global_launch(const AggBlk *blks, const int *list, int nlist, const AggArrays &A, float threshold, int keep_cheaper,
                             bool u8, unsigned char *work, const unsigned long long *woff, uint32_t Ncap_max, bool hybrid,
                             cudaStream_t s) {
    if (nlist == 0) return BS_OK;
    // Ncap_max: the largest (padded) node count of the listed blocks.  16-bit parents / stamps hold node numbers and
    // merge clocks below it.
    const size_t hyb = agglom_par_hyb_bytes(Ncap_max);
    const bool use_hyb = hybrid && Ncap_max <= 65528 && hyb + agglom_par_static_smem() <= 227 * 1024;
#define BS_AGG_GLOB(U8_, HYB_, SM_)                                                                                         \
    do {                                                                                                                    \
        if (HYB_)                                                                                                           \
            BS_CUDA(cudaFuncSetAttribute(k_agglomerate_par<U8_, unsigned long long, uint32_t, false, HYB_>,                 \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,                                       \
                                         (int)(227 * 1024 - agglom_par_static_smem())));                                    \
        BS_LAUNCH((k_agglomerate_par<U8_, unsigned long long, uint32_t, false, HYB_>), nlist, 32 * PNW, SM_, s, blks, list, \
                  A, threshold, keep_cheaper, 0u, 0u, work, woff);                                                          \
    } while (0)
    if (u8 && use_hyb)
        BS_AGG_GLOB(true, true, hyb);
    else if (u8)
        BS_AGG_GLOB(true, false, 0);
    else if (use_hyb)
        BS_AGG_GLOB(false, true, hyb);
    else
        BS_AGG_GLOB(false, false, 0);
#undef BS_AGG_GLOB
    return BS_OK;
}

}  // namespace bs
