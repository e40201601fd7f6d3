"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

The permutation null shards naturally: rank r runs permutations
[P*r/R, P*(r+1)/R) against replicated inputs and only the uint32 exceedance counters are
combined - one integer all-reduce, so results are bit-identical for any number of GPUs.
`by_sample` shards whole samples: no data-path collective; the per-rank result columns (numbers and
integer codes, never strings or pickled frames) are concatenated with one packed all-gather.
Row-sharded staging of a replicated host matrix: `row_block` + `all_gather_csr`.
"""
from __future__ import annotations

import numpy as np


def bind_to_gpu_numa(device_index: int) -> bool:
    """Restrict this process (and the threads it starts later: the engine's upload threads) to the CPU cores NVML
    reports as local to the GPU.  With one process per GPU on a two-socket host, the pageable -> pinned staging copies
    of an upload then stay on the socket the GPU hangs off instead of crossing the inter-socket link.  Call it before
    the big host arrays are allocated (first touch puts them on the local node).  Returns False when NVML or
    `sched_setaffinity` is not available; nothing else changes then."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1 and 64 * w + b < n_cpu}
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return False
        os.sched_setaffinity(0, allowed)
        return True
    except Exception:
        return False


_D2H_STAGE = {}
D2H_CHUNKED_MIN_BYTES = 8 << 20          # smaller columns take torch's plain `.cpu()`


def d2h_chunked(tensors, chunk_bytes: int = 32 << 20, n_threads: int = 4):
    """CUDA tensors (1-D, contiguous) -> new NumPy arrays in pageable memory, through two small pinned staging buffers:
    the copy engine fills one (full PCIe rate) while host threads drain the other into the destination.  torch's
    `.cpu()` of a pageable destination goes through the driver's own bounce buffer one piece at a time and is several
    times slower for gigabyte-sized results; page-locking the whole destination would cost more than the copy."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    tensors = [t.contiguous().view(torch.uint8) if t.dtype != torch.uint8 else t.contiguous() for t in tensors]
    if not tensors:
        return []
    dev = tensors[0].device
    key = (dev.index, chunk_bytes)
    if key not in _D2H_STAGE:
        _D2H_STAGE[key] = ([torch.empty(chunk_bytes, dtype=torch.uint8, pin_memory=True) for _ in range(2)],
                           [torch.cuda.Event() for _ in range(2)], ThreadPoolExecutor(n_threads, thread_name_prefix="lrperm-d2h"))
    stage, events, pool = _D2H_STAGE[key]
    outs = [np.empty(int(t.numel()), dtype=np.uint8) for t in tensors]
    pieces = [(i, o, min(chunk_bytes, int(t.numel()) - o)) for i, t in enumerate(tensors) for o in range(0, int(t.numel()), chunk_bytes)]
    stream = torch.cuda.current_stream(dev)

    def issue(k):
        i, o, n = pieces[k]
        stage[k & 1][:n].copy_(tensors[i][o:o + n], non_blocking=True)
        events[k & 1].record(stream)

    def drain(k):
        i, o, n = pieces[k]
        src = stage[k & 1].numpy()
        step = (n + n_threads - 1) // n_threads
        list(pool.map(lambda a: np.copyto(outs[i][o + a:o + min(n, a + step)], src[a:min(n, a + step)]), range(0, n, step)))

    if pieces:
        issue(0)
    for k in range(len(pieces)):
        events[k & 1].synchronize()
        if k + 1 < len(pieces):
            issue(k + 1)              # the other buffer: its previous contents were drained in the last iteration
        drain(k)
    return outs


class DistContext:
    def __init__(self, rank: int = 0, world_size: int = 1, device=None, group=None):
        self.rank, self.world_size, self.device, self.group = rank, world_size, device, group

    @classmethod
    def from_torch(cls, group=None):
        """Context of the already-initialised default process group (or a single-rank context)."""
        try:
            import torch.distributed as dist
        except Exception:     # pragma: no cover
            return cls()
        if not (dist.is_available() and dist.is_initialized()):
            return cls()
        import torch
        backend = dist.get_backend(group)
        device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        return cls(dist.get_rank(group), dist.get_world_size(group), device, group)

    def perm_range(self, n_perms: int):
        """Contiguous block of global permutation ids owned by this rank."""
        return (n_perms * self.rank) // self.world_size, (n_perms * (self.rank + 1)) // self.world_size

    def all_reduce_sum_u32(self, counts: np.ndarray) -> np.ndarray:
        """Sum a host uint32 vector over ranks (counts <= n_perms, so int32 transport is exact)."""
        if self.world_size == 1:
            return counts
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(counts.astype(np.int32, copy=True)).to(self.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy().astype(np.uint32)

    def all_reduce_sum_(self, tensor):
        """In-place sum of a device int32 tensor (the bench's HBM-resident path)."""
        if self.world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group)
        return tensor

    def sample_shard(self, sizes):
        """CONTIGUOUS blocks of samples per rank, cut where the running number of cells crosses k/world of the total:
        every rank's frames are then already in sample order and the gathered per-rank frames concatenate in rank
        order without re-sorting tens of millions of rows.  Returns the indices owned by this rank."""
        sizes = np.asarray(sizes, dtype=np.float64)
        if self.world_size == 1 or len(sizes) == 0:
            return list(range(len(sizes)))
        mid = np.cumsum(sizes) - sizes / 2.0                      # centre of mass of every sample on the cell axis
        owner = np.minimum((mid / max(sizes.sum(), 1.0) * self.world_size).astype(np.int64), self.world_size - 1)
        return [int(i) for i in np.flatnonzero(owner == self.rank)]

    def upload_csr_sharded(self, indptr, indices, data):
        """Replicated host CSR -> replicated DEVICE CSR without every rank pulling the whole matrix over its own PCIe
        link: rank r uploads the r-th 1/world slice of the flat `indices` / `data` arrays (pinned host tensors or
        NumPy arrays) and the slices are all-gathered over NVLink (NCCL).  Returns torch CUDA tensors
        (indptr, indices, data) to pass to `Engine.set_matrix_csr`.  With one rank it is a plain upload."""
        import torch
        import torch.distributed as dist

        def as_tensor(a):
            return a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))

        indptr, indices, data = as_tensor(indptr), as_tensor(indices), as_tensor(data)
        dev = self.device if self.device is not None else torch.device("cuda", torch.cuda.current_device())
        d_indptr = indptr.to(dev, non_blocking=True)
        if self.world_size == 1:
            out = (d_indptr, indices.to(dev, non_blocking=True), data.to(dev, non_blocking=True))
            if dev.type == "cuda":
                torch.cuda.current_stream(dev).synchronize()
            return out
        nnz = int(indices.numel())
        chunk = (nnz + self.world_size - 1) // self.world_size
        lo, hi = min(nnz, self.rank * chunk), min(nnz, (self.rank + 1) * chunk)
        outs = []
        for arr in (indices, data):
            local = torch.zeros(chunk, dtype=arr.dtype, device=dev)
            local[:hi - lo].copy_(arr[lo:hi], non_blocking=True)
            full = torch.empty(chunk * self.world_size, dtype=arr.dtype, device=dev)
            if dist.get_backend(self.group) == "nccl":
                dist.all_gather_into_tensor(full, local, group=self.group)
            else:
                parts = [torch.empty_like(local) for _ in range(self.world_size)]
                dist.all_gather(parts, local, group=self.group)
                full = torch.cat(parts)
            outs.append(full[:nnz])
        if dev.type == "cuda":
            # the engine works on its own non-blocking stream: hand the tensors over complete
            torch.cuda.current_stream(dev).synchronize()
        return d_indptr, outs[0], outs[1]

    def all_reduce_np(self, arr: np.ndarray, op: str = "sum") -> np.ndarray:
        """Element-wise sum / max of a small host array (float64 or int64) over ranks."""
        if self.world_size == 1:
            return arr
        import torch
        import torch.distributed as dist
        a = np.ascontiguousarray(arr)
        t = torch.from_numpy(a.copy()).to(self.device if self.device is not None else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX, group=self.group)
        return t.cpu().numpy().reshape(a.shape)

    def row_block(self, n_rows: int, indptr=None):
        """Contiguous block of rows [r0, r1) of a replicated host matrix this rank stages: cut where the running number
        of stored entries crosses k/world of the total (CSR `indptr` given), else by rows."""
        w, r = self.world_size, self.rank
        if indptr is None or n_rows == 0:
            return (n_rows * r) // w, (n_rows * (r + 1)) // w
        total = int(indptr[n_rows])
        cuts = [int(np.searchsorted(indptr, total * k // w, side="left")) for k in range(w + 1)]
        cuts[0], cuts[-1] = 0, n_rows
        cuts = [min(max(c, 0), n_rows) for c in cuts]
        for k in range(1, w + 1):
            cuts[k] = max(cuts[k], cuts[k - 1])
        return cuts[r], cuts[r + 1]

    def all_gather_csr(self, engine):
        """Every rank's engine holds the committed matrix of ITS block of cells (same genes): rebuild the whole
        committed matrix on every rank - blocks in rank order - from one padded all-gather per array over NVLink.
        Returns torch tensors (indptr int32 [N+1], indices int32, data float32) on this rank's device, to hand to
        `Engine.set_matrix_csr`."""
        import torch
        import torch.distributed as dist
        dev = self.device if self.device is not None else torch.device("cpu")
        ip, ix, dt = engine.export_matrix_csr(dev)
        n_local, nnz_local = int(ip.numel()) - 1, int(ix.numel())
        meta = self.all_reduce_np(np.array([[n_local, nnz_local] if r == self.rank else [0, 0] for r in range(self.world_size)],
                                           dtype=np.int64))
        rows, nnzs = [int(v) for v in meta[:, 0]], [int(v) for v in meta[:, 1]]

        def gather(local, sizes):
            n_max = max(max(sizes), 1)
            pad = torch.zeros(n_max, dtype=local.dtype, device=dev)
            pad[:local.numel()].copy_(local)
            if dist.get_backend(self.group) == "nccl":
                full = torch.empty(n_max * self.world_size, dtype=local.dtype, device=dev)
                dist.all_gather_into_tensor(full, pad, group=self.group)
                parts = [full[r * n_max:r * n_max + sizes[r]] for r in range(self.world_size)]
            else:
                bufs = [torch.empty_like(pad) for _ in range(self.world_size)]
                dist.all_gather(bufs, pad, group=self.group)
                parts = [bufs[r][:sizes[r]] for r in range(self.world_size)]
            return parts

        ix_all = torch.cat(gather(ix, nnzs))
        dt_all = torch.cat(gather(dt, nnzs))
        counts = torch.cat(gather(torch.diff(ip.to(torch.int64)).to(torch.int32) if n_local else ip[:0], rows))
        indptr = torch.zeros(sum(rows) + 1, dtype=torch.int64, device=dev)
        indptr[1:] = torch.cumsum(counts.to(torch.int64), 0)
        if dev.type == "cuda":
            torch.cuda.current_stream(dev).synchronize()     # the engine works on its own stream
        return indptr.to(torch.int32), ix_all, dt_all

    def all_gather_rows(self, columns: dict, dst=None):
        """Concatenate, in rank order, the per-rank NumPy columns of a table (same column names and dtypes on every
        rank, any number of rows per rank - also zero).  The columns of a rank travel as ONE byte buffer: one size
        exchange, one H2D copy, one padded collective over NVLink (gloo in the CPU tests), one D2H copy.  No pickling,
        no strings.  dst=None: every rank gets the table (`all_gather_into_tensor`); dst=r: only rank r does (`gather`,
        the others return None) - the device -> host copy and everything the caller builds from the table then run on
        one rank with the whole host to itself instead of `world` times side by side."""
        if self.world_size == 1:
            return columns
        import torch
        import torch.distributed as dist
        names = list(columns)
        if not names:
            return columns
        n_local = int(len(columns[names[0]]))
        dev = self.device if self.device is not None else torch.device("cpu")
        nccl = dist.get_backend(self.group) == "nccl"
        sizes = self.all_reduce_np(np.array([n_local if r == self.rank else 0 for r in range(self.world_size)], dtype=np.int64))
        sizes = [int(v) for v in sizes]
        dtypes = [np.dtype(columns[c].dtype) for c in names]
        row_bytes = [d.itemsize for d in dtypes]
        n_max = max(sizes)
        if n_max == 0:
            return {c: columns[c][:0] for c in names} if dst is None or dst == self.rank else None
        # local buffer: column after column, every column padded to n_max rows (8-byte aligned starts)
        col_off, off = [], 0
        for rb in row_bytes:
            col_off.append(off)
            off += (n_max * rb + 7) & ~7
        stride = off
        send = torch.zeros(stride, dtype=torch.uint8, device=dev)
        for c, o, rb in zip(names, col_off, row_bytes):
            if n_local:
                a = np.ascontiguousarray(columns[c]).view(np.uint8).reshape(-1)
                send[o:o + n_local * rb].copy_(torch.from_numpy(a), non_blocking=True)
        mine = dst is None or dst == self.rank
        if nccl and dst is None:
            full = torch.empty(stride * self.world_size, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(full, send, group=self.group)
        elif dst is None:
            bufs = [torch.empty_like(send) for _ in range(self.world_size)]
            dist.all_gather(bufs, send, group=self.group)
            full = torch.cat(bufs)
        else:
            bufs = [torch.empty_like(send) for _ in range(self.world_size)] if mine else None
            dist.gather(send, bufs, dst=dst, group=self.group)
            full = torch.cat(bufs) if mine else None
        if not mine:
            return None
        # compact on the device (rank after rank inside every column), then one device -> host copy per column into the
        # arrays the caller keeps.  They are pageable (page-locking gigabytes for a one-off copy costs more than it saves),
        # and torch's `.cpu()` into pageable memory runs at 2 GB/s: the copy goes through two small pinned buffers instead
        # (`d2h_chunked`: 9-14 GB/s, `profiles/r02_d2h_probe.log`)
        out = {}
        for c, o, d in zip(names, col_off, dtypes):
            parts = [full[r * stride + o:r * stride + o + sizes[r] * d.itemsize] for r in range(self.world_size)]
            col = torch.cat(parts)
            out[c] = (d2h_chunked([col])[0] if col.is_cuda and col.numel() >= D2H_CHUNKED_MIN_BYTES else col.cpu().numpy()).view(d)
        return out

    def all_gather_objects(self, obj):
        if self.world_size == 1:
            return [obj]
        import torch.distributed as dist
        out = [None] * self.world_size
        dist.all_gather_object(out, obj, group=self.group)
        return out
