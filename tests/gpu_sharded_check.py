"""Multi-rank parity (run under torchrun, one rank per GPU): the z-slab sharded path gives exactly the oracle's
single-process result for any rank count.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/gpu_sharded_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bootstrapper_b200.sharded import ShardedSegmenter  # noqa: E402
from bootstrapper_b200.synth import synth_affs  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok_all = True
    for shape, block, ctx, params in [((36, 120, 120), (6, 60, 60), (2, 8, 8), {}),
                                      ((30, 100, 100), (8, 50, 50), (1, 6, 6), {"fragments_in_xy": False}),
                                      ((32, 96, 96), (8, 48, 48), (2, 6, 6), {"fragments_in_xy": False, "seed_eps": 0.01})]:
        from oracle.blockwise import waterz_pipeline
        affs = synth_affs(shape, seed=2)
        ref = waterz_pipeline(affs, params, block_size=block, context=ctx, seed_tie="index", stats_mode="canonical")
        seg = ShardedSegmenter(shape, block, ctx, params, rank=rank, world=world, device=dev)
        g = seg.geo
        win = torch.from_numpy(np.ascontiguousarray(affs[:, g["w0"]:g["w1"]])).to(dev)
        dev_synth = seg.synth_local_affs(seed=2)
        ok = bool(torch.equal(win, dev_synth))
        r = seg.run(win)
        torch.cuda.synchronize()
        own = r["own_fragments"].cpu().numpy().view(np.uint64)
        ok &= bool(np.array_equal(own, ref["fragments"][g["z0"]:g["z1"]]))
        halo = r["fragments"].cpu().numpy().view(np.uint64)
        ok &= bool(np.array_equal(halo, ref["fragments"][g["w0"]:g["w1"]]))
        eu, ev, es = [t.cpu().numpy() for t in r["edges"]]
        got = dict(zip(zip(eu.view(np.uint64).tolist(), ev.view(np.uint64).tolist()), es.tolist()))
        want = ref["rag"].edges
        ok &= set(got) == set(want)
        for k, s in want.items():
            if k in got:
                ok &= bool(np.isnan(got[k])) if s is None else bool(abs(got[k] - s) <= 1e-6 * abs(s))
        ok &= bool(np.array_equal(r["nodes"].cpu().numpy().view(np.uint64), np.array(sorted(ref["rag"].node_pos), np.uint64)))
        for thr, sg in r["segs"].items():
            ok &= bool(np.array_equal(sg.cpu().numpy().view(np.uint64), ref["segs"][thr]["seg"][g["z0"]:g["z1"]]))
        print(f"rank {rank}/{world} shape {shape} xy={params.get('fragments_in_xy', True)}: {'OK' if ok else 'MISMATCH'} "
              f"(own planes {g['z0']}:{g['z1']}, window {g['w0']}:{g['w1']}, edges {len(got)})", flush=True)
        ok_all &= ok
    t = torch.tensor([1 if ok_all else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
