"""N > 1 host logic on CPU: slab partition, fragment-halo exchange, edge all-gather and the global node list,
over torch.distributed with the gloo backend (world_size 2 and 3).  The per-rank compute results are taken
from the oracle so that only the exchange layer is under test."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bootstrapper_b200 import sharded


def test_slab_partition_covers_all_layers():
    for n_layers, world in [(5, 2), (5, 8), (8, 8), (1, 2), (7, 3)]:
        parts = sharded.slab_layers(n_layers, world)
        assert parts[0][0] == 0 and parts[-1][1] == n_layers and len(parts) == world
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
    g = sharded.slab_geometry((125, 10, 10), (25, 5, 5), (3, 1, 1), 1, 2)
    assert (g["z0"], g["z1"], g["w0"], g["w1"]) == (75, 125, 72, 125)
    g = sharded.slab_geometry((120, 10, 10), (25, 5, 5), (3, 1, 1), 0, 2)
    assert (g["z0"], g["z1"], g["w0"], g["w1"]) == (0, 75, 0, 78)


def test_global_node_ids_from_counts():
    ids = sharded.global_node_ids([0, 2, 5], [2, 0, 3], 1000, "cpu")
    assert ids.tolist() == [1, 2, 5001, 5002, 5003]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shape, block, ctx, frags_ref, edges_ref, owner_rank_of_edge, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        geos = [sharded.slab_geometry(shape, block, ctx, r, world) for r in range(world)]
        g = geos[rank]
        # this rank's stage-1 output: own planes of the reference fragments inside a zeroed window
        win = torch.zeros((g["w1"] - g["w0"],) + tuple(shape[1:]), dtype=torch.int64)
        win[g["z0"] - g["w0"]:g["z1"] - g["w0"]] = torch.from_numpy(frags_ref[g["z0"]:g["z1"]].astype(np.int64))
        sharded.exchange_halos(win, g, geos, rank, world)
        ok_halo = bool((win.numpy() == frags_ref[g["w0"]:g["w1"]].astype(np.int64)).all())
        mine = [i for i, r in enumerate(owner_rank_of_edge) if r == rank]
        u = torch.tensor([edges_ref[i][0] for i in mine], dtype=torch.int64)
        v = torch.tensor([edges_ref[i][1] for i in mine], dtype=torch.int64)
        s = torch.tensor([edges_ref[i][2] for i in mine], dtype=torch.float32)
        state = {}
        U, V, S = sharded.allgather_edges(u, v, s, world, state=state)
        got = sorted(zip(U.tolist(), V.tolist(), [x if x == x else None for x in S.tolist()]), key=lambda t: t[:2])
        want = sorted([(a, b, None if c != c else float(np.float32(c))) for a, b, c in edges_ref], key=lambda t: t[:2])
        # the remembered capacity serves the next call without a size exchange; a rank whose edges outgrow it (here: every
        # edge three times over) makes all ranks grow the capacity and repeat
        cap0 = state["cap"]
        U2, V2, S2 = sharded.allgather_edges(u, v, s, world, state=state)
        ok_again = U2.tolist() == U.tolist() and state["cap"] == cap0
        state["cap"] = max(1, u.numel() // 2) if rank == 0 else state["cap"]
        caps = torch.tensor([state["cap"]], dtype=torch.int64)
        dist.all_reduce(caps, op=dist.ReduceOp.MIN)
        state["cap"] = int(caps.item())
        U3, V3, S3 = sharded.allgather_edges(u, v, s, world, state=state)
        ok_grow = U3.tolist() == U.tolist() and V3.tolist() == V.tolist() and state["cap"] > int(caps.item())
        assert ok_again and ok_grow
        # the form the GPU path ships: dense int32 node numbers, 12 bytes per edge, NaN scores intact
        ids = sorted({e[0] for e in edges_ref} | {e[1] for e in edges_ref})
        rank_of = {n: k + 1 for k, n in enumerate(ids)}
        du = torch.tensor([rank_of[int(x)] for x in u.tolist()], dtype=torch.int32)
        dv = torch.tensor([rank_of[int(x)] for x in v.tolist()], dtype=torch.int32)
        D, E2, S4 = sharded.allgather_edges(du, dv, s, world, state={})
        assert D.dtype == torch.int32 and S4.dtype == torch.float32
        back = sorted(zip([ids[k - 1] for k in D.tolist()], [ids[k - 1] for k in E2.tolist()], [x if x == x else None for x in S4.tolist()]),
                      key=lambda t: t[:2])
        assert back == want
        counts = np.zeros(4, np.int64)
        counts[rank] = rank + 1
        tot = sharded.allgather_counts(counts, world, "cpu")
        out_q.put((rank, ok_halo, got == want, tot.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_layer_over_gloo(world):
    from bootstrapper_b200.synth import synth_affs
    from oracle.blockwise import waterz_pipeline
    shape, block, ctx = (24, 60, 60), (6, 30, 30), (2, 4, 4)
    affs = synth_affs(shape, seed=3)
    ref = waterz_pipeline(affs, {}, block_size=block, context=ctx, seed_tie="index", stats_mode="canonical")
    frags = ref["fragments"]
    nvox = int(np.prod(block))
    layer_of_block = {b.block_id: b.index[0] for b in ref["blocks"]}
    parts = sharded.slab_layers(-(-shape[0] // block[0]), world)
    edges, owners = [], []
    for (u, v), s in ref["rag"].edges.items():
        layer = layer_of_block[u // nvox]                   # the edge is owned by the block holding node min(u, v)
        owners.append([r for r, (a, b) in enumerate(parts) if a <= layer < b][0])
        edges.append((u, v, np.nan if s is None else s))
    assert len(set(owners)) == world
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_worker, args=(r, world, port, shape, block, ctx, frags, edges, owners, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_halo, ok_edges, tot in res:
        assert ok_halo, f"rank {rank}: halo planes differ from the global fragments"
        assert ok_edges, f"rank {rank}: gathered edge set differs"
        assert tot[:world] == list(range(1, world + 1))
