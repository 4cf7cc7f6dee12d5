"""Step-count model of the flood kernels (CPU, numpy): how many 32-wide steps does one tile take, and how many does
its busiest warp take when the tile's mask components are dealt to G warps?

    python tools/flood_step_model.py [G ...]

The model follows k_flood2's step rule: a step takes up to 32 entries of the current level's FIFO; every free mask
neighbour is claimed by the first entry that reaches it; the step is cut after the first entry that queues a pixel of a
higher level, and the flood continues there; otherwise the level runs dry and the next lower occupied level follows.
"""
import sys

import numpy as np
from scipy.ndimage import distance_transform_edt, label, maximum_filter

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from bootstrapper_b200.synth import synth_affs  # noqa: E402


def steps_of(level, free, seed_px, W, width=32):
    """level: int per pixel (higher pops first); free: bool per pixel (in mask, not labelled); seed_px: seed pixels in
    ascending index.  Returns (steps, entries)."""
    H = level.shape[0] // W
    fifo = {}
    for p in seed_px:
        fifo.setdefault(level[p], []).append(p)
    heads = {l: 0 for l in fifo}
    steps = entries = 0
    cur = max(fifo) if fifo else None
    while cur is not None:
        q, h = fifo[cur], heads[cur]
        if h == len(q):
            lower = [l for l in fifo if l < cur and heads[l] < len(fifo[l])]
            higher = [l for l in fifo if l > cur and heads[l] < len(fifo[l])]
            assert not higher
            cur = max(lower) if lower else None
            continue
        steps += 1
        batch = q[h:h + width]
        used, jump = 0, None
        for p in batch:
            used += 1
            y, x = divmod(p, W)
            for nb, ok in ((p - W, y > 0), (p - 1, x > 0), (p + 1, x + 1 < W), (p + W, y + 1 < H)):
                if ok and free[nb]:
                    free[nb] = False
                    l = level[nb]
                    fifo.setdefault(l, []).append(nb)
                    heads.setdefault(l, 0)
                    if l > cur:
                        jump = l if jump is None else max(jump, l)
            if jump is not None:
                break
        heads[cur] = h + used
        entries += used
        if jump is not None:
            cur = jump
    return steps, entries


def main():
    groups = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]
    affs = synth_affs((3, 312, 312), seed=0)
    for z in range(3):
        mask = (affs[1, z].astype(np.int32) + affs[2, z]) > 255
        d2 = np.rint(distance_transform_edt(mask) ** 2).astype(np.int64)
        seeds = (maximum_filter(d2, 10) == d2) & mask
        comps, nc = label(mask)
        H, W = mask.shape
        lv = d2.ravel()
        sizes = np.bincount(comps.ravel(), minlength=nc + 1)
        line = [f"slice {z}: {nc} components, {int(mask.sum())} mask pixels"]
        for G in groups:
            # greedy balance: largest component first into the lightest warp
            load = np.zeros(G, dtype=np.int64)
            owner = np.zeros(nc + 1, dtype=np.int64)
            for c in np.argsort(-sizes[1:]) + 1:
                g = int(np.argmin(load))
                owner[c] = g
                load[g] += sizes[c]
            worst = total_steps = total_entries = 0
            for g in range(G):
                mine = (owner[comps] == g) & mask
                free = (mine & ~seeds).ravel().copy()
                seed_px = np.flatnonzero((mine & seeds).ravel())
                st, en = steps_of(lv, free, seed_px, W)
                worst = max(worst, st)
                total_steps += st
                total_entries += en
            line.append(f"G={G}: busiest warp {worst} steps, all warps {total_steps} steps, fill {total_entries / max(total_steps, 1):.1f}/32")
        print("; ".join(line), flush=True)


if __name__ == "__main__":
    main()
