"""`flatnav` — the reference's Python package name (python-bindings/src/flatnav/__init__.py:1-34) on the B200 engine.

    import flatnav
    from flatnav.index import IndexL2Float, create
    from flatnav.data_type import DataType

`_core` is a pybind11 extension over the C ABI of flatnav_b200/libflatnav_b200.so (hand-written CUDA for sm_100a);
there is no CPU path.  Build both with `python -c "import __graft_entry__ as g; g.build()"` (or `make -C flatnav`).
"""
import sys

from ._core import MetricType, __doc__, __version__, data_type  # noqa: F401


class _DataTypeModule:
    from ._core.data_type import DataType


class _IndexModule:
    from ._core.index import (IndexIPFloat, IndexIPInt8, IndexIPUint8, IndexL2Float, IndexL2Int8, IndexL2Uint8,  # noqa: F401
                              create)


index = _IndexModule
sys.modules["flatnav.index"] = _IndexModule
sys.modules["flatnav.data_type"] = _DataTypeModule

__all__ = ["MetricType", "data_type", "index", "__version__", "__doc__"]
