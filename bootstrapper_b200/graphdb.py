"""Fragment RAG store: SQLite file with the node / edge attributes the reference writes through
volara's SQLite wrapper (post/watershed.py:104-113: edge_attrs={"merge_score": "float"}; nodes carry
position and size, post/blockwise/watershed_frags.py:238-246).  Table layout follows
funlib.persistence's SQLite graph provider as far as it can be recalled without its source
(nodes: id, position_0..2, size; edges: u, v, merge_score) — SURVEY N1, unverified.
PostgreSQL configs are rejected (no server / driver in this image).
"""
import os
import sqlite3

import numpy as np


class SQLiteRag:
    db_type = "sqlite"

    def __init__(self, path, edge_attrs=None):
        self.path = path
        self.edge_attrs = edge_attrs or {"merge_score": "float"}
        self.attr = list(self.edge_attrs)[0]      # one float attribute per edge: merge_score (ws), zyx_aff (mws)

    @property
    def id(self):
        return os.path.splitext(os.path.basename(self.path))[0]

    def _con(self):
        os.makedirs(os.path.dirname(os.path.abspath(self.path)), exist_ok=True)
        return sqlite3.connect(self.path)

    def init(self):
        with self._con() as con:
            con.execute("CREATE TABLE IF NOT EXISTS nodes (id INTEGER PRIMARY KEY, position_0 INTEGER, "
                        "position_1 INTEGER, position_2 INTEGER, size INTEGER)")
            con.execute(f"CREATE TABLE IF NOT EXISTS edges (u INTEGER, v INTEGER, {self.attr} REAL, PRIMARY KEY (u, v))")

    def drop(self):
        if os.path.exists(self.path):
            os.remove(self.path)

    def drop_edges(self):
        if os.path.exists(self.path):
            with self._con() as con:
                con.execute("DROP TABLE IF EXISTS edges")

    def write_nodes(self, ids, positions, sizes):
        rows = zip(ids.astype(np.int64).tolist(), *[positions[:, d].tolist() for d in range(3)], sizes.tolist())
        with self._con() as con:
            con.executemany("INSERT OR REPLACE INTO nodes VALUES (?, ?, ?, ?, ?)", rows)

    def write_edges(self, u, v, scores):
        sc = [None if np.isnan(s) else float(s) for s in scores]
        with self._con() as con:
            con.executemany("INSERT OR REPLACE INTO edges VALUES (?, ?, ?)",
                            zip(u.astype(np.int64).tolist(), v.astype(np.int64).tolist(), sc))

    def read_graph(self, roi=None):
        """-> nodes (N,) uint64 ascending, edges (E,2) uint64, scores (E,) float32 (NaN = NULL).
        roi = (offset, shape) in world units: only the nodes positioned inside it and the edges whose node u lies inside it
        (funlib.persistence read_graph(roi) as post/watershed.py:156 calls it with total_roi)."""
        with self._con() as con:
            if roi is None:
                nodes = np.array([r[0] for r in con.execute("SELECT id FROM nodes ORDER BY id")], dtype=np.int64)
                rows = list(con.execute(f"SELECT u, v, {self.attr} FROM edges"))
            else:
                lo = [int(v) for v in roi[0]]
                hi = [int(o) + int(n) for o, n in zip(roi[0], roi[1])]
                cond = " AND ".join(f"position_{d} >= ? AND position_{d} < ?" for d in range(3))
                args = [v for d in range(3) for v in (lo[d], hi[d])]
                nodes = np.array([r[0] for r in con.execute(f"SELECT id FROM nodes WHERE {cond} ORDER BY id", args)], dtype=np.int64)
                rows = list(con.execute(f"SELECT u, v, {self.attr} FROM edges WHERE u IN (SELECT id FROM nodes WHERE {cond})", args))
        edges = np.array([(r[0], r[1]) for r in rows], dtype=np.int64).reshape(-1, 2)
        scores = np.array([np.nan if r[2] is None else r[2] for r in rows], dtype=np.float32)
        return nodes.view(np.uint64), edges.view(np.uint64), scores


def open_db(db_config, edge_attrs=None):
    """post/watershed.py:104-113 (edge_attrs merge_score), post/watershed_mutex.py:108-117 (edge_attrs zyx_aff)"""
    if "db_file" in db_config:
        return SQLiteRag(db_config["db_file"], edge_attrs=edge_attrs or {"merge_score": "float"})
    raise NotImplementedError("PostgreSQL RAG stores are not available in this environment; use db_file (SQLite)")


class LUT:
    """volara.lut.LUT: `<path>.npz` with `fragment_segment_lut` = (2, N) uint64 and the `edges` key volara's
    save(lut, edges=None) always writes (None for the ws pipeline, post/watershed.py:187-188)."""

    def __init__(self, path):
        self.path = path

    def save(self, lut, edges=None):
        os.makedirs(os.path.dirname(os.path.abspath(self.path)), exist_ok=True)
        np.savez_compressed(self.path + ".npz", fragment_segment_lut=np.asarray(lut, dtype=np.uint64), edges=edges)

    def load(self):
        return np.load(self.path + ".npz")["fragment_segment_lut"]
