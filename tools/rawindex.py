"""Test / benchmark infrastructure: write an index file in the reference's cereal layout (SURVEY.md §8c:
60-byte header + [vector | M x u32 links | i32 label] nodes, Index.h:134-141, 555-573) straight from arrays,
without running the reference's construction.  With `links=None` every link slot is a self-loop (the value
allocateNode writes for unused slots, Index.h:270), which is all the exact-scan tests need."""
from __future__ import annotations

import numpy as np

DT_CODE = {np.dtype(np.float32): 9, np.dtype(np.uint8): 0, np.dtype(np.int8): 4}  # util/Datatype.h:11-24


def index_bytes(vectors: np.ndarray, M: int = 8, links: np.ndarray | None = None, labels: np.ndarray | None = None,
                max_nodes: int | None = None) -> bytes:
    v = np.ascontiguousarray(vectors)
    n, dim = v.shape
    max_nodes = max_nodes or n
    data_size = dim * v.dtype.itemsize
    node_size = data_size + 4 * M + 4
    head = np.array([DT_CODE[v.dtype]], dtype="<i4").tobytes() + np.array(
        [M, data_size, node_size, max_nodes, n, dim, data_size], dtype="<u8").tobytes()
    blob = np.zeros((max_nodes, node_size), dtype=np.uint8)
    blob[:n, :data_size] = v.view(np.uint8).reshape(n, data_size)
    if links is None:
        links = np.repeat(np.arange(n, dtype=np.uint32)[:, None], M, axis=1)
    blob[:n, data_size:data_size + 4 * M] = np.ascontiguousarray(links, dtype="<u4").view(np.uint8).reshape(n, 4 * M)
    if labels is None:
        labels = np.arange(n, dtype=np.int32)
    blob[:n, data_size + 4 * M:] = np.ascontiguousarray(labels, dtype="<i4").view(np.uint8).reshape(n, 4)
    return head + blob.tobytes()
