"""Stage-by-stage parity diagnostics of libbsnative against the oracle (run on a GPU box).

    python tests/gpu_diag.py [--shape Z Y X] [--block Z Y X] [--context Z Y X] [--seed S]

Not a pytest: prints mismatch counts per stage so that one GPU call localises a fault.
"""
import argparse
import ctypes as C
import sys
import time
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bootstrapper_b200 import native  # noqa: E402
from bootstrapper_b200.synth import synth_affs  # noqa: E402
from oracle import blockwise as ob  # noqa: E402
from oracle.native import Waterz  # noqa: E402
from oracle.ws import watershed_from_boundary_distance  # noqa: E402
from scipy.ndimage import distance_transform_edt  # noqa: E402


def same_partition(a, b):
    """label-permutation-invariant equality"""
    a = a.ravel().astype(np.int64)
    b = b.ravel().astype(np.int64)
    if ((a == 0) != (b == 0)).any():
        return False, int(((a == 0) != (b == 0)).sum())
    pairs = np.unique(np.stack([a, b], 1), axis=0)
    ok = len(np.unique(pairs[:, 0])) == len(pairs) and len(np.unique(pairs[:, 1])) == len(pairs)
    if ok:
        return True, 0
    # count voxels whose mapping is not the majority mapping
    return False, -1


def prims(dev):
    lib = native.lib()
    rng = np.random.default_rng(0)
    ok = True
    for n in [1, 31, 4096, 4097, 100000, 3_000_001]:
        x = rng.integers(0, 5, n).astype(np.uint32)
        t = torch.from_numpy(x.view(np.int32)).to(dev)
        o = torch.empty_like(t)
        tot = torch.zeros(1, dtype=torch.int32, device=dev)
        native._check(lib.bs_dbg_scan_u32(C.c_void_p(t.data_ptr()), C.c_void_p(o.data_ptr()), C.c_int64(n),
                                          C.c_void_p(tot.data_ptr()), None))
        torch.cuda.synchronize()
        ref = np.concatenate([[0], np.cumsum(x)[:-1]]).astype(np.uint32)
        good = (o.cpu().numpy().view(np.uint32) == ref).all() and int(tot.item()) == int(x.sum())
        ok &= bool(good)
        x8 = (x & 1).astype(np.uint8)
        t8 = torch.from_numpy(x8).to(dev)
        native._check(lib.bs_dbg_scan_u8(C.c_void_p(t8.data_ptr()), C.c_void_p(o.data_ptr()), C.c_int64(n),
                                         C.c_void_p(tot.data_ptr()), None))
        torch.cuda.synchronize()
        ref = np.concatenate([[0], np.cumsum(x8.astype(np.uint32))[:-1]]).astype(np.uint32)
        good = (o.cpu().numpy().view(np.uint32) == ref).all() and int(tot.item()) == int(x8.sum())
        ok &= bool(good)
    print("prims.scan", "OK" if ok else "FAIL")
    ok = True
    for n in [1, 2, 33, 5000, 1_000_003]:
        k = rng.integers(0, 1 << 40, n).astype(np.uint64)
        k[: n // 2] &= np.uint64(0xFFFF)   # many duplicates -> stability matters
        v = np.arange(n, dtype=np.uint32)
        tk = torch.from_numpy(k.view(np.int64)).to(dev)
        tv = torch.from_numpy(v.view(np.int32)).to(dev)
        tk2, tv2 = torch.empty_like(tk), torch.empty_like(tv)
        native._check(lib.bs_dbg_sort_pairs(C.c_void_p(tk.data_ptr()), C.c_void_p(tv.data_ptr()), C.c_void_p(tk2.data_ptr()),
                                            C.c_void_p(tv2.data_ptr()), C.c_int64(n), 0, 40, None))
        torch.cuda.synchronize()
        order = np.argsort(k, kind="stable")
        good = (tk.cpu().numpy().view(np.uint64) == k[order]).all() and (tv.cpu().numpy().view(np.uint32) == v[order]).all()
        ok &= bool(good)
    print("prims.sort", "OK" if ok else "FAIL")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", type=int, nargs=3, default=[20, 160, 160])
    ap.add_argument("--block", type=int, nargs=3, default=[10, 80, 80])
    ap.add_argument("--context", type=int, nargs=3, default=[2, 10, 10])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--xy", type=int, default=1)
    ap.add_argument("--dtype", default="u8")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    shape, block, ctx = tuple(a.shape), tuple(a.block), tuple(a.context)
    prims(dev)

    np_dtype = np.uint8 if a.dtype == "u8" else np.float32
    t_dtype = torch.uint8 if a.dtype == "u8" else torch.float32
    affs = synth_affs(shape, seed=a.seed, dtype=np_dtype)
    g = native.synth_affs(shape, seed=a.seed, dtype=t_dtype, device=dev)
    torch.cuda.synchronize()
    neq = int((g.cpu().numpy() != affs).sum())
    print("synth: device vs numpy mismatches:", neq, "of", affs.size)
    daffs = torch.from_numpy(affs).to(dev)

    params = dict(ob.WS_DEFAULTS)
    params["fragments_in_xy"] = bool(a.xy)
    t0 = time.time()
    ref = ob.waterz_pipeline(affs, params, block_size=block, context=ctx, seed_tie="index", stats_mode="canonical")
    print("oracle pipeline %.2fs; fragments %d edges %d" % (time.time() - t0, len(ref["rag"].node_pos), len(ref["rag"].edges)))

    native.set_debug(True)
    native.set_profiling(True)
    plan = native.Plan(shape, block, ctx, native._aff_dtype(daffs), fragments_in_xy=bool(a.xy))
    frags = plan.fragments(daffs)
    torch.cuda.synchronize()
    print("stage1 profile (ms):", {k: round(v, 3) for k, v in native.get_profile().items()})
    fst = plan.debug_fetch("flood_stats", np.uint32)
    print("flood stats: steps_total %d interrupts %d max_steps_per_tile %d" % (fst[0], fst[1], fst[2]))

    # ---- per tile intermediates
    ids, wo, ws = plan.block_info()
    d2 = plan.debug_fetch("d2", np.uint32)
    seeds = plan.debug_fetch("seeds", np.uint32)
    flood = plan.debug_fetch("flood", np.uint32)
    base = 0
    bad = dict(d2=0, seeds=0, flood=0, tiles=0, flood_tiles=0)
    blocks = {b.block_id: b for b in ref["blocks"]}
    for bi in range(len(ids)):
        b = blocks[int(ids[bi])]
        assert tuple(b.write_offset) == tuple(wo[bi])
        ad = ob.to_ndarray(affs, b.read_offset, b.read_shape, 0)
        ad = ad.astype(np.float64) / 255.0 if affs.dtype == np.uint8 else ad
        if a.xy:
            H, W = b.read_shape[1], b.read_shape[2]
            for z in range(b.write_shape[0]):
                zz = z + ctx[0]
                mean = 0.5 * (ad[2][zz] + ad[1][zz])
                mask = mean > 0.5
                dist = distance_transform_edt(mask)
                r_d2 = np.rint(dist * dist).astype(np.uint32)
                n = H * W
                t_d2 = d2[base:base + n].reshape(H, W)
                t_seed = seeds[base:base + n].reshape(H, W)
                t_flood = flood[base:base + n].reshape(H, W)
                base += (n + 31) & ~31          # tile bases are 32-aligned in xy mode
                bad["tiles"] += 1
                bad["d2"] += int((t_d2 != r_d2).sum())
                fr, _, sd = watershed_from_boundary_distance(dist, mask, return_seeds=True, seed_tie="index")
                sd = sd * mask
                ts = np.where((t_seed == 0xFFFFFFFF), 0, t_seed)
                ok, _ = same_partition(ts, sd)
                bad["seeds"] += 0 if ok else 1
                tf = np.where(t_flood >= 0x80000000, 0, t_flood)
                ok, _ = same_partition(tf, fr)
                if not ok:
                    bad["flood_tiles"] += 1
                    # voxel-level difference under the best label matching: map via seeds
                    bad["flood"] += int((np.unique(np.stack([tf.ravel(), fr.ravel().astype(np.int64)], 1), axis=0).shape[0]))
        else:
            D, H, W = b.read_shape
            mask = np.mean(ad[:3], axis=0) > 0.5
            dist = distance_transform_edt(mask)
            r_d2 = np.rint(dist * dist).astype(np.uint32)
            n = D * H * W
            t_d2 = d2[base:base + n].reshape(D, H, W)
            t_seed = seeds[base:base + n].reshape(D, H, W)
            t_flood = flood[base:base + n].reshape(D, H, W)
            base += n
            bad["tiles"] += 1
            bad["d2"] += int((t_d2 != r_d2).sum())
            fr, _, sd = watershed_from_boundary_distance(dist, mask, return_seeds=True, seed_tie="index")
            sd = sd * mask
            ts = np.where((t_seed == 0xFFFFFFFF), 0, t_seed)
            ok, _ = same_partition(ts, sd)
            bad["seeds"] += 0 if ok else 1
            tf = np.where(t_flood >= 0x80000000, 0, t_flood)
            ok, _ = same_partition(tf, fr)
            if not ok:
                bad["flood_tiles"] += 1
    print("stage1 intermediates:", bad)

    f = frags.cpu().numpy().view(np.uint64)
    rf = ref["fragments"]
    print("fragments: exact id mismatches %d of %d; partition equal: %s" % (int((f != rf).sum()), f.size, same_partition(f, rf)[0]))
    nid, npos, nsz = [t.cpu().numpy() for t in plan.nodes(dev)]
    rn = np.array(sorted(ref["rag"].node_pos.keys()), dtype=np.uint64)
    same_n = len(nid) == len(rn) and (nid.view(np.uint64) == rn).all()
    print("nodes: count %d vs %d ids equal %s" % (len(nid), len(rn), same_n))
    if same_n:
        rp = np.array([ref["rag"].node_pos[int(i)] for i in rn])
        rs = np.array([ref["rag"].node_size[int(i)] for i in rn])
        print("nodes: pos mismatches %d size mismatches %d" % (int((npos != rp).any(1).sum()), int((nsz != rs).sum())))

    # ---- stage 2 (on the oracle's fragments if ours differ, to keep testing downstream)
    fin = frags if (f == rf).all() else torch.from_numpy(rf.view(np.int64)).to(dev)
    if not (f == rf).all():
        print("!! stage 2 runs on the ORACLE fragments because stage 1 differs")
    plan.agglomerate(daffs, fin)
    torch.cuda.synchronize()
    print("stage2 profile (ms):", {k: round(v, 3) for k, v in native.get_profile().items()})
    eu, ev, es = [t.cpu().numpy() for t in plan.edges(dev)]
    got = {(int(u), int(v)): s for u, v, s in zip(eu.view(np.uint64), ev.view(np.uint64), es)}
    want = ref["rag"].edges
    print("edges: %d vs oracle %d; same key set: %s" % (len(got), len(want), set(got) == set(want)))
    nbad = 0
    worst = 0.0
    for k, s in want.items():
        if k not in got:
            continue
        gsc = got[k]
        if s is None:
            nbad += 0 if np.isnan(gsc) else 1
        elif np.isnan(gsc):
            nbad += 1
        else:
            rel = abs(float(gsc) - s) / max(abs(s), 1e-12)
            worst = max(worst, rel)
            nbad += rel > 1e-6
    print("edge scores: mismatches(>1e-6 rel) %d, worst rel err %.3g" % (nbad, worst))
    cnt = plan.debug_fetch("s2_counters", np.uint32).reshape(-1, 6)[:, :3]
    print("agglomeration counters (pops, stale, dead) summed:", cnt.sum(0), "merges:", plan.debug_fetch("s2_nmerges", np.uint32).sum())

    # per-block deep check of block 0: initial edge statistics + history
    b0 = blocks[int(ids[0])]
    dbg = ob.agglomerate_in_block(b0, affs, rf, ref["rag"], (0, 0, 0), "canonical", True, return_debug=True)
    ebase = plan.debug_fetch("s2_ebase", np.uint32)
    e0, e1 = int(ebase[0]), int(ebase[1])
    du = plan.debug_fetch("s2_eu", np.uint64)[e0:e1]
    dv = plan.debug_fetch("s2_ev", np.uint64)[e0:e1]
    ds = plan.debug_fetch("s2_escore", np.float32)[e0:e1]
    ou, ov, os0 = dbg["initial"]
    same_order = len(du) == len(ou) and (np.minimum(du, dv) == np.minimum(ou, ov)).all() and (np.maximum(du, dv) == np.maximum(ou, ov)).all()
    print("block0: edges %d vs %d, creation order equal: %s" % (len(du), len(ou), same_order))
    if same_order:
        l = dbg["lca"]
        both_nan = np.isnan(l) & np.isnan(ds)
        print("block0: lca score mismatches %d" % int((~both_nan & (l.astype(np.float32) != ds)).sum()))
    hn = int(plan.debug_fetch("s2_nmerges", np.uint32)[0])
    print("block0: merges %d vs oracle %d; counters %s vs %s" % (hn, len(dbg["history"][0]), cnt[0], dbg["counters"]))

    # ---- stage 3
    nodes_t = torch.from_numpy(rn.view(np.int64)).to(dev)
    wk = [(k, s) for k, s in want.items() if s is not None]
    wu = torch.tensor([k[0] for k, _ in wk], dtype=torch.int64, device=dev)
    wv = torch.tensor([k[1] for k, _ in wk], dtype=torch.int64, device=dev)
    wsc = torch.tensor([s for _, s in wk], dtype=torch.float32, device=dev)
    rft = torch.from_numpy(rf.view(np.int64)).to(dev)
    for thr in ref["params"]["thresholds"]:
        comp = native.connected_components(nodes_t, wu, wv, wsc, thr)
        seg = native.relabel(rft, nodes_t, comp)
        torch.cuda.synchronize()
        rs_ = ref["segs"][thr]["seg"]
        print("thr %.2f: lut equal %s; seg exact mismatches %d" % (
            thr, bool((comp.cpu().numpy().view(np.uint64) == ref["segs"][thr]["lut"][1]).all()),
            int((seg.cpu().numpy().view(np.uint64) != rs_).sum())))


if __name__ == "__main__":
    main()
