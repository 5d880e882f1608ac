"""Multi-GPU execution of the search path: one process per GPU, `torch.distributed` for the plumbing.

The reference's only parallelism is a thread fan-out over queries inside one process
(`executeInParallel`, include/flatnav/util/Multithreading.h:18-48, used by bindings.cpp:196-212); it has no
multi-process or multi-device mode.  Two modes here:

* query sharding   — every rank holds a full replica; the batch is split into contiguous slices, one per rank.
                     No collective on the data path; `gather_results` is only for callers that want the whole
                     result on every rank.
* dataset sharding — every rank holds ONE sub-graph built over a contiguous id range whose labels are global ids;
                     every rank searches ALL queries on its shard and the per-shard `(dist, label)[Q, K]` lists are
                     combined into the global top-K (ties -> lower label).  Two exchange paths:
                       "peer"  (default on GPUs) `fnb_search_sharded`: after the traversal ONE kernel pushes the lists
                               into every peer's gather buffer through CUDA-IPC peer pointers over NVLink, signals /
                               awaits per-rank flags and merges — no collective call on the data path;
                       "nccl"  the baseline: `all_gather_into_tensor` + `fnb_merge_topk`.

The search and merge steps are injectable so the host logic (partitioning, gather layout, merge order) is covered
by world_size-2 `gloo` tests on CPU with the oracle standing in for the CUDA kernels (tests only).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def partition(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [start, start+count) of rank `rank`: ceil-sized slices, like fnb_search's replica split."""
    per = (n_items + world - 1) // world
    start = min(n_items, per * rank)
    stop = min(n_items, per * (rank + 1))
    return start, stop - start


def _dist():
    import torch.distributed as dist
    return dist


class QueryShardedSearcher:
    """Replicated index, queries split across ranks."""

    def __init__(self, index, group=None):
        self.index = index
        self.group = group

    def search_local(self, queries: np.ndarray, K: int, ef_search: int, num_initializations: int = 100):
        dist = _dist()
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        start, count = partition(queries.shape[0], world, rank)
        d, l = self.index.search(queries[start:start + count], K, ef_search, num_initializations)
        return start, d, l

    def search(self, queries: np.ndarray, K: int, ef_search: int, num_initializations: int = 100):
        """Whole-batch result on every rank (slices gathered in rank order)."""
        import torch
        dist = _dist()
        world = dist.get_world_size(self.group)
        start, d, l = self.search_local(queries, K, ef_search, num_initializations)
        per = (queries.shape[0] + world - 1) // world
        pad_d = np.full((per, K), np.inf, dtype=np.float32)
        pad_l = np.full((per, K), -1, dtype=np.int32)
        pad_d[: d.shape[0]] = d
        pad_l[: l.shape[0]] = l
        td, tl = torch.from_numpy(pad_d), torch.from_numpy(pad_l)
        if dist.get_backend(self.group) == "nccl":
            td, tl = td.cuda(), tl.cuda()
        gd = [torch.empty_like(td) for _ in range(world)]
        gl = [torch.empty_like(tl) for _ in range(world)]
        dist.all_gather(gd, td, group=self.group)
        dist.all_gather(gl, tl, group=self.group)
        Q = queries.shape[0]
        return (torch.cat(gd).cpu().numpy()[:Q], torch.cat(gl).cpu().numpy()[:Q])


def merge_topk_cuda(gathered_dist, gathered_label, K: int):
    """[S, Q, K] device tensors -> ([Q, K], [Q, K]) device tensors, via fnb_merge_topk on the current stream."""
    import torch

    from . import _capi
    S, Q, _ = gathered_dist.shape
    od = torch.empty((Q, K), dtype=torch.float32, device=gathered_dist.device)
    ol = torch.empty((Q, K), dtype=torch.int32, device=gathered_dist.device)
    _capi.check(_capi.lib().fnb_merge_topk(gathered_dist.data_ptr(), gathered_label.data_ptr(), S, Q, K, od.data_ptr(),
                                           ol.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return od, ol


class DatasetShardedSearcher:
    """One sub-graph per rank (labels are global ids); all ranks answer all queries; all-gather + k-way merge."""

    def __init__(self, shard_index, group=None,
                 search_fn: Optional[Callable] = None, merge_fn: Optional[Callable] = None,
                 exchange: str = "peer", max_queries: int = 1 << 17, max_k: int = 128):
        self.index = shard_index
        self.group = group
        self._search_fn = search_fn
        self._merge_fn = merge_fn or merge_topk_cuda
        self.exchange = "nccl" if search_fn is not None else exchange
        self._ex = None
        self._cap = (int(max_queries), int(max_k))

    # ---- "peer" path: gather buffers in NVLink peer memory, set up once -----------------------------------
    def _peer_exchange(self):
        if self._ex is not None:
            return self._ex
        import ctypes as C

        import torch

        from . import _capi
        dist = _dist()
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        ex = C.c_void_p()
        _capi.check(_capi.lib().fnb_exchange_create(torch.cuda.current_device(), rank, world, self._cap[0], self._cap[1],
                                                    C.byref(ex)))
        mine = C.create_string_buffer(_capi.FNB_IPC_HANDLE_BYTES)
        _capi.check(_capi.lib().fnb_exchange_handle(ex, mine))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(mine.raw), group=self.group)  # control plane only
        blob = C.create_string_buffer(b"".join(handles), world * _capi.FNB_IPC_HANDLE_BYTES)
        _capi.check(_capi.lib().fnb_exchange_attach(ex, blob))
        dist.barrier(group=self.group)
        self._ex = ex
        return ex

    def close(self):
        if self._ex is not None:
            from . import _capi
            _dist().barrier(group=self.group)  # nobody may still be pushing into a buffer that is about to go
            _capi.lib().fnb_exchange_free(self._ex)
            self._ex = None

    def check_status(self) -> None:
        """After a synchronise: raises if a peer never delivered its lists to this rank (peer exchange only)."""
        if self._ex is not None:
            import torch

            from . import _capi
            torch.cuda.synchronize()
            _capi.check(_capi.lib().fnb_exchange_status(self._ex))

    def search_device(self, d_queries: int, Q: int, K: int, ef_search: int, num_initializations: int, d_out_dist: int,
                      d_out_label: int, stream: int = 0) -> None:
        """Collective, asynchronous: raw device pointers in, global top-K [Q, K] out (peer-memory path)."""
        from . import _capi
        _capi.check(_capi.lib().fnb_search_sharded(self.index._h, self._peer_exchange(), d_queries, Q, K, ef_search,
                                                   num_initializations, d_out_dist, d_out_label, stream or None))

    def search_tensors(self, dq, K: int, ef_search: int, num_initializations: int = 100):
        """Collective and asynchronous on the current stream: `dq` is a CUDA tensor [Q, dim] of the index dtype;
        returns CUDA tensors (float32 [Q, K], int32 [Q, K]) with the global top-K.  Used by the benchmarks."""
        import torch
        Q = dq.shape[0]
        stream = torch.cuda.current_stream().cuda_stream
        od = torch.empty((Q, K), dtype=torch.float32, device=dq.device)
        ol = torch.empty((Q, K), dtype=torch.int32, device=dq.device)
        if self.exchange == "peer":
            self.search_device(dq.data_ptr(), Q, K, ef_search, num_initializations, od.data_ptr(), ol.data_ptr(), stream)
            return od, ol
        dist = _dist()
        world = dist.get_world_size(self.group)
        d = torch.empty((Q, K), dtype=torch.float32, device=dq.device)
        l = torch.empty((Q, K), dtype=torch.int32, device=dq.device)
        self.index.search_device(dq.data_ptr(), Q, K, ef_search, num_initializations, d.data_ptr(), l.data_ptr(), stream)
        gd = torch.empty((world * Q, K), dtype=d.dtype, device=d.device)
        gl = torch.empty((world * Q, K), dtype=l.dtype, device=l.device)
        dist.all_gather_into_tensor(gd, d, group=self.group)
        dist.all_gather_into_tensor(gl, l, group=self.group)
        return merge_topk_cuda(gd.view(world, Q, K), gl.view(world, Q, K), K)

    def _search_peer(self, queries, K, ef, ninit):
        import torch

        from . import _capi
        Q = queries.shape[0]
        dq = torch.from_numpy(np.ascontiguousarray(queries)).cuda()
        od = torch.empty((Q, K), dtype=torch.float32, device="cuda")
        ol = torch.empty((Q, K), dtype=torch.int32, device="cuda")
        self.search_device(dq.data_ptr(), Q, K, ef, ninit, od.data_ptr(), ol.data_ptr(),
                           torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        _capi.check(_capi.lib().fnb_exchange_status(self._ex))
        return od.cpu().numpy(), ol.cpu().numpy()

    def _local(self, queries, K, ef, ninit):
        import torch
        if self._search_fn is not None:  # tests: oracle stand-in on CPU
            d, l = self._search_fn(queries, K, ef, ninit)
            return torch.from_numpy(np.ascontiguousarray(d)), torch.from_numpy(np.ascontiguousarray(l))
        Q = queries.shape[0]
        dq = torch.from_numpy(np.ascontiguousarray(queries)).cuda()
        d = torch.empty((Q, K), dtype=torch.float32, device="cuda")
        l = torch.empty((Q, K), dtype=torch.int32, device="cuda")
        self.index.search_device(dq.data_ptr(), Q, K, ef, ninit, d.data_ptr(), l.data_ptr(),
                                 torch.cuda.current_stream().cuda_stream)
        return d, l

    def search(self, queries: np.ndarray, K: int, ef_search: int, num_initializations: int = 100):
        import torch
        dist = _dist()
        world = dist.get_world_size(self.group)
        if self.exchange == "peer":
            return self._search_peer(queries, K, ef_search, num_initializations)
        d, l = self._local(queries, K, ef_search, num_initializations)
        Q = d.shape[0]
        gd = torch.empty((world * Q, K), dtype=d.dtype, device=d.device)  # rank-major concatenation
        gl = torch.empty((world * Q, K), dtype=l.dtype, device=l.device)
        dist.all_gather_into_tensor(gd, d, group=self.group)
        dist.all_gather_into_tensor(gl, l, group=self.group)
        od, ol = self._merge_fn(gd.view(world, Q, K), gl.view(world, Q, K), K)
        return od.cpu().numpy(), ol.cpu().numpy()
