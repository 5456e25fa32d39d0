"""Multi-GPU layout of the hot path: one process per GPU,
`torch.distributed` for the plumbing.

  * search: the corpus is row-sharded contiguously (global id = id_offset + local row),
    queries are replicated, every rank scans its shard.  One exchange step, no other collective:
      - nq <= 64 (streaming scan): the exchange is fused into the scan kernel
        (css_index_search_exchange_device): the CTA that finishes the local top-k stores it into
        every peer's memory over NVLink (CUDA IPC mappings, set up once with one all-gather of the
        64-byte handles), waits for the peers' lists and merges -- no NCCL call, no extra launch;
      - larger batches (tensor-core path): one all-gather of the packed lists (k x 12 B per query and
        rank) + css_topk_merge_device on every rank.
  * encode: sequences are split contiguously across ranks; no collective (each rank's rows
    can be appended to its own corpus shard).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np

from . import _native


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of n items over `world` ranks (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def merge_topk_host(D_all: np.ndarray, I_all: np.ndarray, k: int, metric: int = _native.METRIC_INNER_PRODUCT):
    """[W, nq, k] lists -> [nq, k]; best first, ties by ascending id, -1 ids are holes.
    Host restatement of css_topk_merge_device (used on the gloo/CPU path and in tests)."""
    W, nq, _ = D_all.shape
    D = np.full((nq, k), -np.finfo(np.float32).max if metric == 0 else np.finfo(np.float32).max, np.float32)
    I = np.full((nq, k), -1, np.int64)
    for q in range(nq):
        d = D_all[:, q, :].reshape(-1)
        i = I_all[:, q, :].reshape(-1)
        ok = i >= 0
        d, i = d[ok], i[ok]
        key = -d.astype(np.float64) if metric == 0 else d.astype(np.float64)
        o = np.lexsort((i, key))[:k]
        D[q, :len(o)] = d[o]
        I[q, :len(o)] = i[o]
    return D, I


class ShardedSearch:
    """Exact top-k over a row-sharded corpus; call collectively on every rank."""

    def __init__(self, index: Optional[_native.Index], id_offset: int, group=None,
                 local_search: Optional[Callable] = None):
        import torch.distributed as dist
        self.dist = dist
        self.index = index
        self.id_offset = int(id_offset)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        # CSS_EXCHANGE=nccl keeps the all-gather + merge-kernel exchange for every nq
        import os
        self.use_exchange = os.environ.get("CSS_EXCHANGE", "p2p") != "nccl"
        self.local_search = local_search  # CPU/gloo tests inject a host search; None = the device index
        self._bufs = {}
        self._ex = None
        self._pin = {}

    def _exchange(self):
        """The in-kernel exchange of this rank, connected to all peers on first use (NCCL groups only)."""
        if self._ex is None:
            ex = _native.Exchange(self.index.device, self.world, self.dist.get_rank(self.group))
            handles = [None] * self.world
            self.dist.all_gather_object(handles, ex.handle, group=self.group)   # 64 bytes per rank, any backend
            ex.connect(handles)
            self.dist.barrier(group=self.group)
            self._ex = ex
        return self._ex

    def close(self):
        if self._ex is not None:
            import torch
            torch.cuda.synchronize(self.index.device)
            if self.dist.is_initialized():
                self.dist.barrier(group=self.group)   # nobody unmaps memory a peer's kernel may still write
            self._ex.close()
            self._ex = None

    def _device_bufs(self, torch, dev, nq, k):
        key = (nq, k)
        if key not in self._bufs:
            # one packed buffer per rank -- scores, then (8-byte aligned) ids -- so that a search needs ONE all-gather
            d_bytes = (nq * k * 4 + 7) // 8 * 8
            stride = d_bytes + nq * k * 8
            loc = torch.empty(stride, device=dev, dtype=torch.uint8)
            allb = torch.empty(self.world * stride, device=dev, dtype=torch.uint8)
            self._bufs[key] = dict(
                loc=loc, all=allb, stride=stride, d_bytes=d_bytes,
                D_loc=loc[:nq * k * 4].view(torch.float32).view(nq, k),
                I_loc=loc[d_bytes:].view(torch.int64).view(nq, k),
                outp=(outp := torch.empty(stride, device=dev, dtype=torch.uint8)),   # merged result, same packing
                D_out=outp[:nq * k * 4].view(torch.float32).view(nq, k),
                I_out=outp[d_bytes:].view(torch.int64).view(nq, k))
        return self._bufs[key]

    def search_device(self, q, k: int, mask_ptr: int = 0):
        """q: CUDA float32 tensor [nq, d] (replicated).  Returns (D, I) CUDA tensors holding the
        merged global top-k on every rank.  Runs on torch's current stream."""
        import torch
        dev = q.device
        nq = q.shape[0]
        b = self._device_bufs(torch, dev, nq, k)
        sp = torch.cuda.current_stream(dev).cuda_stream
        if self.world > 1 and self.use_exchange and nq <= _native.EXCHANGE_MAX_NQ:
            self.index.search_exchange_device(self._exchange(), q.data_ptr(), nq, k, b["D_out"].data_ptr(),
                                              b["I_out"].data_ptr(), mask_ptr, self.id_offset, sp)
            return b["D_out"], b["I_out"]
        self.index.search_device(q.data_ptr(), nq, k, b["D_loc"].data_ptr(), b["I_loc"].data_ptr(), mask_ptr,
                                 self.id_offset, sp)
        if self.world == 1:
            return b["D_loc"], b["I_loc"]
        self.dist.all_gather_into_tensor(b["all"], b["loc"], group=self.group)
        base = b["all"].data_ptr()
        _native.topk_merge_strided_device(base, b["stride"] // 4, base + b["d_bytes"], b["stride"] // 8, self.world, nq, k,
                                          self.index.metric, b["D_out"].data_ptr(), b["I_out"].data_ptr(), sp)
        return b["D_out"], b["I_out"]

    def search_host(self, q: np.ndarray, k: int):
        """Host buffers in and out.  NCCL group: the query goes to the device once, the merged
        list comes back once.  gloo group (CPU tests): host gather + host merge."""
        import torch
        q = np.ascontiguousarray(q, np.float32).reshape(-1, q.shape[-1])
        if self.local_search is None and self.world > 1 and self.dist.get_backend(self.group) == "nccl":
            # one pinned block per (nq, k): query in, [scores | ids] out -- one H2D, one D2H, one sync
            dev = torch.device("cuda", self.index.device)
            nq = q.shape[0]
            if nq == 1 and self.use_exchange:
                # single query: staged, searched, exchanged and returned through mapped host memory inside the library
                return self.index.search_exchange_host(self._exchange(), q, k, self.id_offset)
            key = (nq, k, q.shape[1])
            b = self._device_bufs(torch, dev, nq, k)
            if key not in self._pin:
                self._pin[key] = dict(
                    q=torch.empty((nq, q.shape[1]), dtype=torch.float32).pin_memory(),
                    qd=torch.empty((nq, q.shape[1]), dtype=torch.float32, device=dev),
                    out=torch.empty(b["stride"], dtype=torch.uint8).pin_memory())
            pb = self._pin[key]
            pb["q"].numpy()[...] = q
            pb["qd"].copy_(pb["q"], non_blocking=True)
            self.search_device(pb["qd"], k)          # merged lists land in b["outp"] (scores | ids)
            pb["out"].copy_(b["outp"], non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            o = pb["out"].numpy()
            return (o[:nq * k * 4].view(np.float32).reshape(nq, k).copy(),
                    o[b["d_bytes"]:].view(np.int64).reshape(nq, k).copy())
        if self.local_search is not None:
            D, I = self.local_search(q, k)
        else:
            D, I = self.index.search(q, k)
        I = np.where(I >= 0, I + self.id_offset, -1)
        if self.world == 1:
            return D, I
        Dl = torch.from_numpy(np.ascontiguousarray(D))
        Il = torch.from_numpy(np.ascontiguousarray(I))
        Dg = [torch.empty_like(Dl) for _ in range(self.world)]
        Ig = [torch.empty_like(Il) for _ in range(self.world)]
        self.dist.all_gather(Dg, Dl, group=self.group)
        self.dist.all_gather(Ig, Il, group=self.group)
        return merge_topk_host(torch.stack(Dg).numpy(), torch.stack(Ig).numpy(), k)
