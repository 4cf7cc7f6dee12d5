"""One-process-per-GPU driver of the blockwise ws path: contiguous z-slabs of whole daisy blocks per rank.

The reference's blocks are independent inside a task and tasks are separated by global barriers
(post/watershed.py:137,141,192; data crosses blocks through zarr + SQL).  Here the same three barriers are
three small exchanges over torch.distributed (NCCL on GPUs, gloo in the CPU tests):
  after stage 1   all-gather of per-block fragment counts (dense node numbering) and a z-halo exchange of
                  `context[0]` fragment planes with both slab neighbours (stage 2 reads neighbours' fragments
                  inside its read ROI, waterz_agglom.py:111)
  after stage 2   variable-size all-gather of the owned (u, v, merge_score) edges — every cross-slab edge is
                  owned by exactly one block, so this *is* the cross-slab RAG merge
  stage 3         every rank runs the (tiny) thresholded CC on the full graph and relabels its own slab
Block geometry never depends on the rank count, so results are identical for any number of GPUs.
The exchange helpers are device-agnostic so that the N > 1 host logic is covered by gloo tests on CPU.
"""
import queue
import threading

import numpy as np
import torch
import torch.distributed as dist

from . import native
from .post.pipeline import resolve_ws_params


def slab_layers(n_layers, world):
    """contiguous, balanced split of the z-layers of blocks: [(l0, l1)] per rank"""
    base, extra = divmod(n_layers, world)
    out, l = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((l, l + n))
        l += n
    return out


def slab_geometry(vol_shape, block_size, context, rank, world):
    """own planes [z0, z1) and window [w0, w1) (own + context halo, clipped) of a rank"""
    Z, bz, cz = vol_shape[0], block_size[0], context[0]
    n_layers = -(-Z // bz)
    l0, l1 = slab_layers(n_layers, world)[rank]
    z0, z1 = min(l0 * bz, Z), min(l1 * bz, Z)
    w0, w1 = max(0, z0 - cz), min(Z, z1 + cz)
    if l0 == l1:
        w0 = w1 = z0
    return dict(l0=l0, l1=l1, z0=z0, z1=z1, w0=w0, w1=w1)


def exchange_halos(frags_win, geo, geos, rank, world, group=None):
    """fill the halo planes of this rank's fragment window with the neighbours' own planes.
    frags_win: (w1 - w0, Y, X) int64; planes [z0, z1) are this rank's own."""
    if world == 1:
        return
    ops, bufs = [], []
    for nb in (rank - 1, rank + 1):
        if nb < 0 or nb >= world:
            continue
        g, h = geo, geos[nb]
        if g["z0"] == g["z1"] or h["z0"] == h["z1"]:
            continue
        # planes of mine the neighbour needs: my own planes inside its window
        s0, s1 = max(g["z0"], h["w0"]), min(g["z1"], h["w1"])
        if s1 > s0:
            send = frags_win[s0 - g["w0"]:s1 - g["w0"]].contiguous()
            bufs.append(send)
            ops.append(dist.P2POp(dist.isend, send, nb, group=group))
        # planes of the neighbour inside my window
        r0, r1 = max(h["z0"], g["w0"]), min(h["z1"], g["w1"])
        if r1 > r0:
            recv = torch.empty((r1 - r0,) + tuple(frags_win.shape[1:]), dtype=frags_win.dtype, device=frags_win.device)
            bufs.append((recv, r0 - g["w0"], r1 - g["w0"]))
            ops.append(dist.P2POp(dist.irecv, recv, nb, group=group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for b in bufs:
        if isinstance(b, tuple):
            frags_win[b[1]:b[2]] = b[0]


def allgather_counts(counts, world, device, group=None):
    """per-block fragment counts: every block is owned by exactly one rank -> element-wise sum"""
    if world == 1:
        return np.asarray(counts, dtype=np.int64)
    t = torch.as_tensor(counts, dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def allgather_edges(u, v, s, world, group=None, state=None):
    """variable-size all-gather of the owned edges in ONE collective: every rank contributes a fixed-capacity record
    [n, u[0..n), v[0..n), score bits[0..n)] to all_gather_into_tensor; the edge count travels in the record's header, so no
    separate size exchange is needed.  u, v: int64 ids, or int32 dense node numbers (12 instead of 24 bytes per edge: the
    dense numbering is global once the block counts are exchanged); the record's word type follows them.  The capacity is
    remembered in `state` (a dict the caller keeps between calls) and grown -- with one extra round -- when some rank's
    edges do not fit."""
    if world == 1:
        return u, v, s
    state = state if state is not None else {}
    n = u.numel()
    dev = u.device
    wt = u.dtype
    sbits = s.view(torch.int32) if wt == torch.int32 else s.view(torch.int32).to(torch.int64)
    while True:
        cap = int(state.get("cap", 0))
        if cap == 0:
            # first call: agree on a capacity
            t = torch.tensor([n], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            cap = int(t.item()) * 5 // 4 + 1024
            state["cap"] = cap
        rec = torch.empty(1 + 3 * cap, dtype=wt, device=dev)
        k = min(n, cap)
        rec[0] = n
        rec[1:1 + k] = u[:k]
        rec[1 + cap:1 + cap + k] = v[:k]
        rec[1 + 2 * cap:1 + 2 * cap + k] = sbits[:k]
        out = torch.empty(world * (1 + 3 * cap), dtype=wt, device=dev)
        dist.all_gather_into_tensor(out, rec, group=group)
        out = out.view(world, 1 + 3 * cap)
        sizes = out[:, 0].cpu().tolist()
        if max(sizes) > cap:                 # somebody overflowed: every rank sees it, grow and repeat
            state["cap"] = max(sizes) * 5 // 4 + 1024
            continue
        U = torch.cat([out[r, 1:1 + sizes[r]] for r in range(world)])
        V = torch.cat([out[r, 1 + cap:1 + cap + sizes[r]] for r in range(world)])
        S = torch.cat([out[r, 1 + 2 * cap:1 + 2 * cap + sizes[r]] for r in range(world)]).to(torch.int32).view(torch.float32)
        return U, V, S


class HostExpander:
    """Decoder thread of the expanded host path: the results cross the bus in the compact form (run_host_compact, 4 bytes per
    voxel) and the uint64 fragments + segmentations the reference writes are rebuilt in HOST memory by bs_expand_compact while
    the device works on the next volume.  Two compact buffer sets alternate: acquire() before a set is overwritten,
    submit() after run_host_compact(wait=False) queued its copies, flush() before the arrays are read."""

    def __init__(self, threads):
        import queue
        import threading
        self.threads = int(threads)
        self.q = queue.Queue()
        self.free = threading.Semaphore(2)
        self.error = None
        self.thread = threading.Thread(target=self._work, daemon=True)
        self.thread.start()

    def _work(self):
        while True:
            item = self.q.get()
            try:
                if item is None:
                    return
                done, hs, n, outs = item
                done.synchronize()
                native.expand_compact(hs["dense"], hs["nodes"][:n], [l[:n] for l in hs["luts"]], outs[0], outs[1:], threads=self.threads)
            except Exception as e:  # noqa: BLE001
                self.error = e
            finally:
                if item is not None:
                    self.free.release()
                self.q.task_done()

    def acquire(self):
        self.free.acquire()
        if self.error is not None:
            raise native.BsError(f"host expander failed: {self.error!r}")

    def submit(self, done, host_set, n_nodes, outs):
        self.q.put((done, host_set, int(n_nodes), outs))

    def flush(self):
        self.q.join()
        if self.error is not None:
            raise native.BsError(f"host expander failed: {self.error!r}")

    def close(self):
        self.q.put(None)
        self.thread.join()


def global_node_ids(block_ids, counts, nvox_block, device):
    """fragment ids are 1..n per block + block_id * prod(block_size) (watershed_frags.py:224): the sorted
    global node list follows from the per-block counts alone (blocks ascending by id)."""
    parts = [torch.arange(1, int(c) + 1, dtype=torch.int64, device=device) + int(b) * int(nvox_block)
             for b, c in zip(block_ids, counts) if c > 0]
    return torch.cat(parts) if parts else torch.zeros(0, dtype=torch.int64, device=device)


class HostRing:
    """Device -> host streaming of (Z, Y, X) int64 result arrays through a bounded ring of page-locked z-chunk buffers.

    A downloader thread owns a copy stream: every submitted device array is cut into chunks of whole z planes, each chunk
    is copied into the next ring slot as soon as the slot's previous content has landed and been handed to `sink(name,
    z0, z1, pinned_view)` (the consumer: a zarr writer, a checksum, nothing).  Page-locked memory is n_slots * chunk_bytes
    per rank whatever the volume size, so streaming stays on with 8 ranks on one host."""

    def __init__(self, plane_shape, device, chunk_bytes=64 << 20, n_slots=16, sink=None):
        Y, X = int(plane_shape[-2]), int(plane_shape[-1])
        self.planes = max(1, int(chunk_bytes) // (Y * X * 8))
        self.n_slots = int(n_slots)
        self.slots = [torch.empty((self.planes, Y, X), dtype=torch.int64, pin_memory=True) for _ in range(self.n_slots)]
        self.pinned_bytes = self.n_slots * self.planes * Y * X * 8
        self.pending = [None] * self.n_slots       # (cuda event of the copy into the slot, (name, z0, z1))
        device = torch.device(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.sink = sink
        self.next = 0
        self.chunks_done = 0
        self.error = None
        self.q = queue.Queue()
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def _consume(self, k):
        if self.pending[k] is not None:
            ev, meta = self.pending[k]
            ev.synchronize()
            if self.sink is not None:
                self.sink(meta[0], meta[1], meta[2], self.slots[k][:meta[2] - meta[1]])
            self.pending[k] = None
            self.chunks_done += 1

    def wait(self, issued):
        """block until the downloader has issued every chunk of a submitted array; a dead downloader raises instead of
        hanging the caller"""
        while not issued["flag"].wait(0.5):
            if self.error is not None or not self.thread.is_alive():
                raise RuntimeError(f"HostRing downloader stopped: {self.error!r}")
        if self.error is not None:
            raise RuntimeError(f"HostRing downloader failed: {self.error!r}")

    def _loop(self):
        try:
            torch.cuda.set_device(self.device)
        except Exception as e:  # noqa: BLE001
            self.error = e
            return
        while True:
            job = self.q.get()
            if job is None:
                return
            kind, name, tensor, ready, issued = job
            try:
                if kind == "flush":
                    for i in range(self.n_slots):
                        self._consume((self.next + i) % self.n_slots)
                else:
                    with torch.cuda.stream(self.stream):
                        self.stream.wait_event(ready)
                        Z = tensor.shape[0]
                        ev = None
                        for z0 in range(0, Z, self.planes):
                            z1 = min(Z, z0 + self.planes)
                            k = self.next % self.n_slots
                            self.next += 1
                            self._consume(k)
                            self.slots[k][:z1 - z0].copy_(tensor[z0:z1], non_blocking=True)
                            ev = torch.cuda.Event()
                            ev.record(self.stream)
                            self.pending[k] = (ev, (name, z0, z1))
                        issued["event"] = ev          # the device array is free again once its last chunk has left
            except Exception as e:  # noqa: BLE001
                self.error = e
            finally:
                issued["flag"].set()

    def submit(self, name, tensor, ready):
        """queue the download of `tensor` (Z, Y, X) int64 once the cuda event `ready` has fired; returns a handle whose
        `flag` is set when every chunk has been issued and whose `event` then marks the end of the last copy"""
        issued = {"flag": threading.Event(), "event": None}
        self.q.put(("copy", name, tensor, ready, issued))
        return issued

    def flush(self):
        """wait until every queued array has landed and been handed to the sink"""
        issued = {"flag": threading.Event(), "event": None}
        self.q.put(("flush", None, None, None, issued))
        self.wait(issued)

    def close(self):
        self.flush()
        self.q.put(None)
        self.thread.join()
        self.slots = []


class ShardedSegmenter:
    def __init__(self, vol_shape, block_size, context, params=None, rank=0, world=1, device=None, group=None):
        self.vol_shape = tuple(int(v) for v in vol_shape)
        self.block_size, self.context = tuple(block_size), tuple(context)
        self.p = resolve_ws_params(params)
        self.rank, self.world, self.group = rank, world, group
        self.device = device if device is not None else torch.device("cuda")
        self.geos = [slab_geometry(self.vol_shape, self.block_size, self.context, r, world) for r in range(world)]
        self.geo = self.geos[rank]
        g = self.geo
        self.win_shape = (g["w1"] - g["w0"],) + self.vol_shape[1:]
        self.own_shape = (g["z1"] - g["z0"],) + self.vol_shape[1:]
        self.nvox_block = int(np.prod(self.block_size))
        self.plan = None
        self._copy_stream = None
        self._edge_state = {}
        self._inflight = []
        self.last_profile = {}

    def _plan(self, dtype_code):
        if self.plan is None:
            g = self.geo
            win = dict(win_z0=g["w0"], win_z=g["w1"] - g["w0"]) if self.world > 1 else {}
            self.plan = native.Plan(self.vol_shape, self.block_size, self.context, dtype_code,
                                    fragments_in_xy=self.p["fragments_in_xy"], min_seed_distance=self.p["min_seed_distance"],
                                    filter_fragments=self.p["filter_fragments"], remove_debris=self.p["remove_debris"],
                                    bias=self.p["bias"], seed_eps=self.p["seed_eps"], sigma=self.p["sigma"],
                                    noise_eps=self.p["noise_eps"], noise_seed=self.p.get("noise_seed", 0) or 0,
                                    block_begin=g["l0"] if self.world > 1 else -1, block_end=g["l1"] if self.world > 1 else -1,
                                    **win)
            self.block_ids, _, _ = self.plan.block_info()
        return self.plan

    def synth_local_affs(self, seed=0, dtype=torch.uint8):
        """this rank's window of the synthetic volume, generated on the device (block-addressable generator)"""
        g = self.geo
        return native.synth_affs(self.win_shape, seed=seed, dtype=dtype, offset=(g["w0"], 0, 0), device=self.device)

    def run(self, affs_win, out=None, frag_sink=None, out_ready=None, relabel=True):
        """affs_win: (C, w1-w0, Y, X) on this rank's device.  Returns dict with the fragment window, the
        segmentations of the own planes per threshold and the global graph."""
        plan = self._plan(native._aff_dtype(affs_win))
        g = self.geo
        prof = {}
        # every voxel of the task ROI lies in some block's write ROI; halo planes of a slab window are filled by
        # the neighbour exchange, so only multi-rank windows need the zero fill
        alloc = torch.zeros if self.world > 1 else torch.empty
        frags = alloc(self.win_shape, dtype=torch.int64, device=affs_win.device)
        plan.fragments(affs_win, frags_out=frags)
        prof.update(native.get_profile())
        if frag_sink is not None:
            # the own planes are final after stage 1: their device->host copy overlaps stages 2 and 3
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=affs_win.device)
            ev = torch.cuda.Event()
            ev.record()
            if isinstance(frag_sink, HostRing):
                frag_sink.submit(("fragments", None), frags[g["z0"] - g["w0"]:g["z1"] - g["w0"]], ev)
            else:
                with torch.cuda.stream(self._copy_stream):
                    self._copy_stream.wait_event(ev)
                    frag_sink.copy_(frags[g["z0"] - g["w0"]:g["z1"] - g["w0"]], non_blocking=True)
        counts = allgather_counts(plan.block_counts(), self.world, affs_win.device, self.group)
        if self.world > 1:
            plan.set_block_counts(counts)
            exchange_halos(frags, g, self.geos, self.rank, self.world, self.group)
        plan.agglomerate(affs_win, frags)
        prof.update(native.get_profile())
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        sub = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        evs[0].record()
        eu, ev, es = plan.edges(affs_win.device)
        nodes = plan.node_ids(affs_win.device)
        sub[0].record()
        if self.world > 1 and nodes.numel() < (1 << 31) - 1:
            # ship dense node numbers (int32) instead of ids: 12 bytes per edge on the wire
            du, dv, es = allgather_edges(plan.dense_fragments(eu), plan.dense_fragments(ev), es, self.world, self.group, self._edge_state)
            eu, ev = nodes[(du - 1).long()], nodes[(dv - 1).long()]
        else:
            eu, ev, es = allgather_edges(eu, ev, es, self.world, self.group, self._edge_state)
        sub[1].record()
        own = frags[g["z0"] - g["w0"]:g["z1"] - g["w0"]]
        thrs = list(self.p["thresholds"])
        evs[1].record()
        cmap = plan.components(nodes, eu, ev, es, thrs)
        evs[2].record()
        comps = [cmap[thr] for thr in thrs]
        luts = dict(zip(thrs, comps))
        segs = {}
        own = own.contiguous()
        if out_ready is not None:
            # `out` is still being read by an earlier volume's device->host copies: only the relabel waits for them
            torch.cuda.current_stream().wait_event(out_ready)
        for i in range(0, len(thrs) if relabel else 0, 8):
            for thr, sg in zip(thrs[i:i + 8], plan.relabel(own, comps[i:i + 8], None if out is None else out[i:i + 8])):
                segs[thr] = sg
        evs[3].record()
        evs[3].synchronize()
        prof["s3.graph"] = evs[0].elapsed_time(evs[1])          # edge / node tables (+ the all-gather when sharded)
        if self.world > 1:
            prof["s3.graph.gather"] = sub[0].elapsed_time(sub[1])   # part of s3.graph: the edge all-gather and its repacking
        prof["s3.components"] = evs[1].elapsed_time(evs[2])
        prof["s3.relabel"] = evs[2].elapsed_time(evs[3])
        self.last_profile = prof
        return dict(fragments=frags, own_fragments=own, segs=segs, luts=luts, nodes=nodes, edges=(eu, ev, es))

    def run_host(self, host_affs, host_out, out=None, wait=True, ring=None):
        """end-to-end with HOST buffers: pinned affinities in, fragments + segmentations out.
        host_out: pinned (own_shape) int64 arrays [fragments, seg per threshold] the results are copied into, or None with
        ring: a HostRing -- the results stream through its bounded set of page-locked z-chunk buffers to its sink instead.
        out: optional device staging buffers for the segmentations (one per threshold).
        wait=False: streaming use over many volumes -- the call returns once the copies are queued; the device->host
        copies of this volume then overlap the next calls' upload and compute (the copy stream is FIFO, and a relabel that
        writes into `out` buffers an earlier volume is still being copied from waits for those copies).  Alternate
        between two sets of `out` (and `host_out`) buffers and call `drain()` before reading the last results."""
        affs = host_affs.to(self.device, non_blocking=True)
        key = out[0].data_ptr() if out else None
        busy = [f[0] for f in self._inflight if key is not None and f[3] == key]
        out_ready = busy[-1] if busy else None
        if isinstance(out_ready, dict):          # ring handle: the downloader thread records the event
            ring.wait(out_ready)
            out_ready = out_ready["event"]
        self._ring = ring
        r = self.run(affs, out=out, frag_sink=host_out[0] if ring is None else ring, out_ready=out_ready)
        ready = torch.cuda.Event()
        ready.record()
        if ring is not None:
            done = None
            for thr in self.p["thresholds"]:
                done = ring.submit(("seg", thr), r["segs"][thr], ready)
        else:
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(ready)
                for i, thr in enumerate(self.p["thresholds"]):
                    host_out[1 + i].copy_(r["segs"][thr], non_blocking=True)
                done = torch.cuda.Event()
                done.record(self._copy_stream)
        self._inflight.append((done, r, affs, key))     # the device arrays stay referenced until their copies are through
        while len(self._inflight) > (0 if wait else 2):
            self._wait_done(self._inflight.pop(0)[0])
        return r

    def _wait_done(self, done):
        if isinstance(done, dict):
            self._ring.wait(done)
            if done["event"] is not None:
                done["event"].synchronize()
        elif done is not None:
            done.synchronize()

    def run_host_compact(self, host_affs, host_out, wait=True):
        """end-to-end with HOST buffers in the compact result form (include/bsnative.h): pinned affinities in; out come ONE
        int32 plane of dense fragment numbers for the own planes, the node-id table and a LUT row per threshold -- 4 bytes
        per voxel across the bus instead of 8 (T + 1).  native.expand_compact rebuilds the uint64 arrays on the host.
        host_out: dict(dense=pinned int32 (own_shape), nodes=pinned int64 (capacity,), luts=[pinned int64 (capacity,)] * T),
        filled in place; returns dict(n_nodes=...) (the tables hold n_nodes valid entries).
        wait=False: as run_host -- alternate between two sets of host_out and call drain() before reading the last one."""
        affs = host_affs.to(self.device, non_blocking=True)
        r = self.run(affs, relabel=False)
        g = self.geo
        dense = self.plan.dense_fragments(r["own_fragments"])
        n = r["nodes"].numel()
        if n > host_out["nodes"].numel():
            raise native.BsError(f"compact host tables hold {host_out['nodes'].numel()} nodes, the volume has {n}")
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ready)
            host_out["dense"].copy_(dense, non_blocking=True)
            host_out["nodes"][:n].copy_(r["nodes"], non_blocking=True)
            for i, thr in enumerate(self.p["thresholds"]):
                host_out["luts"][i][:n].copy_(r["luts"][thr], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        self._inflight.append((done, (r, dense), affs, None))
        while len(self._inflight) > (0 if wait else 2):
            self._wait_done(self._inflight.pop(0)[0])
        return dict(n_nodes=n, done=done)

    def drain(self):
        """wait for the device->host copies of every volume queued by run_host(wait=False)"""
        while self._inflight:
            self._wait_done(self._inflight.pop(0)[0])
        if getattr(self, "_ring", None) is not None:
            if self._ring.thread.is_alive():
                self._ring.flush()
            else:
                self._ring = None          # closed by its owner
