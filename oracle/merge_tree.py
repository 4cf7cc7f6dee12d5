"""ORACLE: merge-history tree and LCA score query — restates post/merge_tree.py:5-113.

PINNED: tests/test_oracle_pinned.py replays tests/golden/merge_tree_*.npz, which were
produced by executing the reference's own post/merge_tree.py (tests/golden/make_golden.py).
"""
import numpy as np


class MergeTree:
    def __init__(self, leaf_nodes=None):
        self._idx = {}
        self._level = []
        self._next = []
        self._score = []
        self.id_to_node = {}
        self.next_id = 0
        if leaf_nodes is not None:
            leaves = [int(n) for n in leaf_nodes]
            for n in leaves:
                if n not in self._idx:
                    self._add(n, 0, 0.0)
                    self.id_to_node[n] = n
            self.next_id = max(leaves) + 1                     # merge_tree.py:59

    def _add(self, node_id, level, score):
        idx = len(self._level)
        self._idx[node_id] = idx
        self._level.append(level)
        self._next.append(-1)
        self._score.append(score)
        return idx

    def merge(self, u, v, target, score):                      # merge_tree.py:70-83
        u, v, target = int(u), int(v), int(target)
        t = self.next_id
        self.next_id += 1
        iu = self._idx[self.id_to_node[u]]
        iv = self._idx[self.id_to_node[v]]
        level = max(self._level[iu], self._level[iv]) + 1
        it = self._add(t, level, float(score))
        self._next[iu] = it
        self._next[iv] = it
        self.id_to_node[target] = t

    def find_merges(self, us, vs):                             # merge_tree.py:5-27, 94-109
        level, nxt, score = self._level, self._next, self._score
        out = np.empty(len(us), dtype=np.float64)
        for k in range(len(us)):
            u = self._idx.get(int(us[k]), -1)
            v = self._idx.get(int(vs[k]), -1)
            if u < 0 or v < 0:
                out[k] = np.nan
                continue
            while True:
                if u == v:
                    out[k] = score[u]
                    break
                if level[u] > level[v]:
                    u, v = v, u
                if nxt[u] < 0:
                    out[k] = np.nan
                    break
                u = nxt[u]
        return out
