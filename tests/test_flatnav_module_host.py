"""CPU: the `flatnav`-named package (flatnav/__init__.py + the pybind11 `_core` over the C ABI) has the reference's module
layout and names (python-bindings/src/flatnav/__init__.py:1-34, bindings.cpp:426-538).  No compute."""
import inspect
import sys

import pytest

import flatnav
from flatnav.data_type import DataType
from flatnav.index import IndexIPFloat, IndexL2Float, create  # the reference's own unit tests import exactly these


def test_layout_matches_the_reference_package():
    assert flatnav.__all__ == ["MetricType", "data_type", "index", "__version__", "__doc__"]
    assert sys.modules["flatnav.index"] is flatnav.index and "flatnav.data_type" in sys.modules
    for name in ("IndexL2Float", "IndexIPFloat", "IndexL2Uint8", "IndexIPUint8", "IndexL2Int8", "IndexIPInt8", "create"):
        assert hasattr(flatnav.index, name)
    assert int(DataType.float32) == 9 and int(DataType.int8) == 4 and int(DataType.uint8) == 0  # util/Datatype.h:11-24
    assert flatnav.data_type.float32 == DataType.float32  # export_values()
    assert int(flatnav.MetricType.L2) == 0 and int(flatnav.MetricType.IP) == 1
    assert isinstance(flatnav.__version__, str)


def test_methods_and_keyword_names_match_the_binding():
    want = {
        "add": ["data", "ef_construction", "num_initializations", "labels"],
        "allocate_nodes": ["data"],
        "search_single": ["query", "K", "ef_search", "num_initializations"],
        "search": ["queries", "K", "ef_search", "num_initializations"],
        "save": ["filename"], "build_graph_links": ["mtx_filename"], "reorder": ["strategies"],
        "set_num_threads": ["num_threads"], "load_index": ["filename"],
        "get_query_distance_computations": [], "get_graph_outdegree_table": [],
    }
    for cls in (IndexL2Float, IndexIPFloat, flatnav.index.IndexL2Uint8, flatnav.index.IndexIPInt8):
        for meth, args in want.items():
            doc = getattr(cls, meth).__doc__.splitlines()[0]  # pybind11 puts the signature on the first docstring line
            for a in args:
                assert f"{a}:" in doc, (cls.__name__, meth, a, doc)
        assert "num_initializations: typing.SupportsInt | typing.SupportsIndex = 100" in cls.search.__doc__ or \
            "num_initializations: int = 100" in cls.search.__doc__
        assert isinstance(inspect.getattr_static(cls, "max_edges_per_node"), property)
        assert isinstance(inspect.getattr_static(cls, "num_threads"), property)
    doc = create.__doc__.splitlines()[0]
    for a in ("distance_type", "dim", "dataset_size", "max_edges_per_node", "index_data_type", "verbose", "collect_stats"):
        assert f"{a}:" in doc


def test_errors_without_compute():
    with pytest.raises(ValueError, match="Invalid distance type"):  # bindings.cpp:397-407
        create(distance_type="cosine", dim=8, dataset_size=10, max_edges_per_node=4)
    with pytest.raises(RuntimeError, match="Unable to open file for reading"):  # Index.h:445-447
        IndexL2Float.load_index("/nonexistent/x.idx")
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        with pytest.raises(RuntimeError, match="no CPU path"):  # loud failure, not a fallback
            create(distance_type="l2", dim=8, dataset_size=10, max_edges_per_node=4)
