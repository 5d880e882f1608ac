"""flatnav_b200 — B200-native batched query engine for FlatNav's search hot path.

Drop-in for the search side of the reference's `flatnav` Python package
(python-bindings/src/flatnav/__init__.py): `flatnav_b200.index.IndexL2Float.load_index(path)
.search(queries, K, ef_search)`, `flatnav_b200.data_type.DataType`, `flatnav_b200.MetricType`.
Everything runs in hand-written CUDA for sm_100a behind the C ABI of include/flatnav_b200.h; there is
no CPU fallback.
"""
import enum

from . import data_type, index  # noqa: F401

__version__ = "0.1.0"


class MetricType(enum.IntEnum):
    """flatnav::distances::MetricType (bindings.cpp:517-521)."""
    L2 = 0
    IP = 1


__all__ = ["MetricType", "data_type", "index", "__version__"]
