"""KGraph hand-off: the flat CSR view of /root/reference/src/fromhnsw/kgraph.rs:108-120 and its file format.

The reference's KGraph has no serialisation; `write_csr`/`read_csr` define the interchange file
that a Rust-side writer fills from `KGraph::get_neighbours()` (kgraph.rs:157) and `get_data_id_from_idx` (:335).

File layout (little endian):
    magic  8 bytes  b"ANNKGCSR"
    u32 version (=1), u32 flags (=0)
    u64 n, u64 E, u64 max_nbng
    u64 row_ptr[n+1] ; u32 col[E] ; f32 dist[E] ; u64 data_id[n]
"""
from __future__ import annotations

import struct

import numpy as np

MAGIC = b"ANNKGCSR"


class KGraph:
    """Neighbour lists of every node, sorted by increasing distance (kgraph.rs:508-509), as CSR arrays.

    node index i (0..n) is the rank in the graph; `data_id[i]` is the caller's DataId (kgraph.rs:116-119 node_set).
    """

    def __init__(self, row_ptr, col, dist, data_id=None, max_nbng=None):
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        self.col = np.ascontiguousarray(col, dtype=np.uint32)
        self.dist = np.ascontiguousarray(dist, dtype=np.float32)
        n = len(self.row_ptr) - 1
        if n < 0 or len(self.col) != len(self.dist) or (n >= 0 and int(self.row_ptr[-1]) != len(self.col)):
            raise ValueError("inconsistent CSR arrays")
        self.data_id = np.arange(n, dtype=np.uint64) if data_id is None else np.ascontiguousarray(data_id, np.uint64)
        if len(self.data_id) != n:
            raise ValueError("data_id must have one entry per node")
        deg = np.diff(self.row_ptr.astype(np.int64))
        self.max_nbng = int(max_nbng) if max_nbng is not None else (int(deg.max()) if n else 0)
        self._idx_of = None

    # accessors named after kgraph.rs:147-164,335-348
    def get_nb_nodes(self) -> int:
        return len(self.row_ptr) - 1

    def get_max_nbng(self) -> int:
        return self.max_nbng

    def get_nb_edges(self) -> int:
        return len(self.col)

    def get_neighbours(self):
        """(row_ptr, col, dist): the flat equivalent of &Vec<Vec<OutEdge<F>>>."""
        return self.row_ptr, self.col, self.dist

    def get_out_edges_by_idx(self, node: int):
        lo, hi = int(self.row_ptr[node]), int(self.row_ptr[node + 1])
        return self.col[lo:hi], self.dist[lo:hi]

    def get_data_id_from_idx(self, idx: int) -> int:
        return int(self.data_id[idx])

    def get_idx_from_dataid(self, data_id: int) -> int:
        if self._idx_of is None:
            self._idx_of = {int(d): i for i, d in enumerate(self.data_id)}
        return self._idx_of[int(data_id)]

    @classmethod
    def from_knn(cls, idx: np.ndarray, dist: np.ndarray, data_id=None) -> "KGraph":
        """Regular graph from (n,k) neighbour indices / distances (rows ascending, no self)."""
        n, k = idx.shape
        row_ptr = np.arange(0, (n + 1) * k, k, dtype=np.uint64)
        return cls(row_ptr, idx.reshape(-1), dist.reshape(-1), data_id, max_nbng=k)


def write_csr(path: str, g: KGraph) -> None:
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<II", 1, 0))
        f.write(struct.pack("<QQQ", g.get_nb_nodes(), g.get_nb_edges(), g.max_nbng))
        f.write(g.row_ptr.astype("<u8").tobytes())
        f.write(g.col.astype("<u4").tobytes())
        f.write(g.dist.astype("<f4").tobytes())
        f.write(g.data_id.astype("<u8").tobytes())


def read_csr(path: str) -> KGraph:
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError("not an ANNKGCSR file")
        version, _flags = struct.unpack("<II", f.read(8))
        if version != 1:
            raise ValueError(f"unsupported ANNKGCSR version {version}")
        n, e, max_nbng = struct.unpack("<QQQ", f.read(24))
        row_ptr = np.frombuffer(f.read(8 * (n + 1)), dtype="<u8")
        col = np.frombuffer(f.read(4 * e), dtype="<u4")
        dist = np.frombuffer(f.read(4 * e), dtype="<f4")
        data_id = np.frombuffer(f.read(8 * n), dtype="<u8")
        if len(row_ptr) != n + 1 or len(col) != e or len(dist) != e or len(data_id) != n:
            raise ValueError("truncated ANNKGCSR file")
    return KGraph(row_ptr, col, dist, data_id, max_nbng)
