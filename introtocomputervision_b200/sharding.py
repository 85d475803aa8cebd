"""Multi-GPU sharding of the stereo matcher: one process per GPU, ``torch.distributed`` for the plumbing.

The path has no exchange step inside the computation — every output pixel depends on a
(2R+1)-row neighbourhood of the two inputs — so it shards two ways, neither of which the reference has
(it is single-GPU, single-stream: ProblemSets/ps2_cpp/lib/DisparitySSD.cu:143-207):

* **by stereo pair** for batches (BASELINE config 5): rank r owns a contiguous block of pairs;
* **by row band with a halo** for one large image (BASELINE config 4): rank r owns output rows
  [r0, r1) and reads input rows [r0-R-1, r1+R+1) clamped to the image — R rows of window plus one more
  for the neighbouring padded row the reference's SSD reads through its flat index.

The only exchange is the gather of the per-rank results: ``all_gather_into_tensor`` (NCCL over NVLink on
GPUs, gloo in the CPU tests) or, on the GPUs of one box, copy-engine pushes into every rank's buffer
(``PeerGather``), which overlap the persistent hot kernel where a collective cannot.  The compute itself is a call into libstereo_b200.so per rank; a
``compute`` object can be injected so that the partition/gather logic is testable without a GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _capi


def pair_shard(n_pairs: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [begin, end) of pair indices owned by `rank`; block sizes differ by at most one."""
    if n_pairs < 0 or world <= 0 or not (0 <= rank < world):
        raise ValueError("bad pair_shard arguments")
    base, extra = divmod(n_pairs, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def band_shard(rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Output rows [r0, r1) of `rank`: equal bands of ceil(rows / world) rows, the last one(s) shorter or empty."""
    if rows <= 0 or world <= 0 or not (0 <= rank < world):
        raise ValueError("bad band_shard arguments")
    band = -(-rows // world)
    r0 = min(rows, rank * band)
    return r0, min(rows, r0 + band)


def band_halo(rows: int, r0: int, r1: int, window_rad: int) -> Tuple[int, int]:
    """Input rows [h0, h1) the band [r0, r1) needs (``stereo_band_halo_rows``: host arithmetic, no device)."""
    h0, h1 = C.c_int(0), C.c_int(0)
    st = _capi.lib().stereo_band_halo_rows(int(rows), int(r0), int(r1), int(window_rad), C.byref(h0), C.byref(h1))
    if st != _capi.STEREO_OK:
        raise ValueError(_capi.last_error())
    return int(h0.value), int(h1.value)


CUDA_STREAM_LEGACY = 1   # cudaStreamLegacy: torch reports its default stream as handle 0, which the C ABI reads as "context stream"


class GpuCompute:
    """Per-rank compute on this rank's B200 through the C ABI (device buffers are torch tensors).  Work is
    enqueued on torch's current stream, so it is ordered with the surrounding torch ops and collectives."""

    def __init__(self, ctx, device):
        import torch
        self.torch, self.ctx, self.device = torch, ctx, device
        self._elem = {torch.int8: 1, torch.int16: 2, torch.int32: 4}

    def band(self, cost, left_slab, right_slab, rows, cols, r0, r1, h0, h1, window_rad, min_disp, max_disp, dtype):
        torch = self.torch
        dl = torch.as_tensor(np.ascontiguousarray(left_slab)).to(self.device, non_blocking=True)
        dr = torch.as_tensor(np.ascontiguousarray(right_slab)).to(self.device, non_blocking=True)
        out = torch.empty((r1 - r0, cols), dtype=dtype, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream or CUDA_STREAM_LEGACY
        st = _capi.lib().stereo_disparity_band_halo_u8_device(
            self.ctx.handle, int(cost), dl.data_ptr(), cols, dr.data_ptr(), cols, rows, cols, r0, r1, h0, h1,
            int(window_rad), int(min_disp), int(max_disp), out.data_ptr(), cols * self._elem[dtype], self._elem[dtype],
            C.c_void_p(stream))
        if st != _capi.STEREO_OK:
            raise RuntimeError(_capi.last_error())
        return out

    def pair_band(self, cost, left_slab, right_slab, rows, cols, r0, r1, h0, h1, window_rad, disparity_range, dtype):
        """Both maps of the band [r0, r1) in one launch sequence: (2, r1 - r0, cols)."""
        torch = self.torch
        dl = torch.as_tensor(np.ascontiguousarray(left_slab)).to(self.device, non_blocking=True)
        dr = torch.as_tensor(np.ascontiguousarray(right_slab)).to(self.device, non_blocking=True)
        out = torch.empty((2, r1 - r0, cols), dtype=dtype, device=self.device)
        e = self._elem[dtype]
        stream = torch.cuda.current_stream(self.device).cuda_stream or CUDA_STREAM_LEGACY
        st = _capi.lib().stereo_disparity_pair_band_halo_u8_device(
            self.ctx.handle, int(cost), dl.data_ptr(), cols, dr.data_ptr(), cols, rows, cols, r0, r1, h0, h1,
            int(window_rad), int(disparity_range), out[0].data_ptr(), out[1].data_ptr(), cols * e, e, C.c_void_p(stream))
        if st != _capi.STEREO_OK:
            raise RuntimeError(_capi.last_error())
        return out

    def pair_batch(self, cost, lefts, rights, window_rad, disparity_range, dtype):
        torch = self.torch
        n, rows, cols = lefts.shape
        dl = torch.as_tensor(np.ascontiguousarray(lefts)).to(self.device, non_blocking=True)
        dr = torch.as_tensor(np.ascontiguousarray(rights)).to(self.device, non_blocking=True)
        out = torch.empty((2, n, rows, cols), dtype=dtype, device=self.device)
        e = self._elem[dtype]
        stream = torch.cuda.current_stream(self.device).cuda_stream or CUDA_STREAM_LEGACY
        if n:
            st = _capi.lib().stereo_disparity_pair_batch_u8_device(
                self.ctx.handle, int(cost), n, dl.data_ptr(), dr.data_ptr(), cols, rows * cols, rows, cols,
                int(window_rad), int(disparity_range), out[0].data_ptr(), out[1].data_ptr(), cols * e, rows * cols * e, e,
                C.c_void_p(stream))
            if st != _capi.STEREO_OK:
                raise RuntimeError(_capi.last_error())
        return out


class _RawDeviceBytes:
    """Lets torch view a raw device pointer (a buffer owned by libstereo_b200.so) as a uint8 tensor."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerGatherUnavailable(RuntimeError):
    """Raised on every rank of the group when any rank could not create or map a gather buffer."""


class PeerGather:
    """Gather buffer filled by copy-engine pushes over NVLink (``stereo_peer_*`` in include/stereo_b200.h).

    Every rank owns a buffer of ``nbytes``; ``push(offset, src_ptr, n, stream)`` copies n bytes of this rank's
    device memory into EVERY rank's buffer at ``offset`` (its own included), stream-ordered after the work
    already enqueued on ``stream`` and without using an SM — the persistent hot kernel occupies every SM, so
    an NCCL collective could only run between its launches, a copy-engine push runs underneath them.
    ``torch.distributed`` (NCCL on GPUs) is used for the plumbing: the exchange of the 64-byte buffer handles
    and the barrier that makes a gather complete."""

    def __init__(self, ctx, nbytes: int, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.ctx, self.group = torch, dist, ctx, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.nbytes = int(nbytes)
        lib = _capi.lib()
        local = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        # Set-up is collective and fails on EVERY rank or on none: a rank that cannot create or map a buffer
        # (no peer access between two GPUs, IPC disabled in a container) still takes part in both exchanges.
        err = None
        self.local_ptr = 0
        if lib.stereo_peer_buffer_create(ctx.handle, self.nbytes, C.byref(local), handle) == _capi.STEREO_OK:
            self.local_ptr = int(local.value)
        else:
            err = _capi.last_error()
        handles = [bytes(handle) if err is None else None]
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle) if err is None else None, group=group)
        self.ptrs = (C.c_void_p * self.world)()
        self._opened = []
        if all(h is not None for h in handles):
            for r in range(self.world):
                if r == self.rank:
                    self.ptrs[r] = self.local_ptr
                    continue
                peer = C.c_void_p()
                if lib.stereo_peer_buffer_open(ctx.handle, (C.c_ubyte * 64).from_buffer_copy(handles[r]), C.byref(peer)) != _capi.STEREO_OK:
                    err = err or f"rank {self.rank} cannot map rank {r}'s buffer: {_capi.last_error()}"
                    break
                self.ptrs[r] = peer.value
                self._opened.append(int(peer.value))
        else:
            err = err or "another rank could not create its gather buffer"
        errs = [err]
        if self.world > 1:
            errs = [None] * self.world
            dist.all_gather_object(errs, err, group=group)
        bad = [e for e in errs if e]
        if bad:
            for p_ in self._opened:
                lib.stereo_peer_buffer_close(ctx.handle, C.c_void_p(p_))
            self._opened = []
            if self.local_ptr:
                lib.stereo_peer_buffer_destroy(ctx.handle, C.c_void_p(self.local_ptr))
                self.local_ptr = 0
            raise PeerGatherUnavailable(bad[0])

    @staticmethod
    def _check(st):
        if st != _capi.STEREO_OK:
            raise RuntimeError(_capi.last_error())

    def push(self, offset: int, src_ptr: int, nbytes: int, stream: int, to=None) -> None:
        """`to`: ranks whose buffers receive the bytes (default: every rank's - a gather to all; ``to=[0]`` gathers to rank 0)."""
        if offset < 0 or offset + nbytes > self.nbytes:
            raise ValueError("push outside the gather buffer")
        ptrs = self.ptrs
        if to is not None:
            key = tuple(sorted(set(int(r) for r in to)))
            ptrs = self._subsets.get(key) if hasattr(self, "_subsets") else None
            if ptrs is None:
                if not hasattr(self, "_subsets"):
                    self._subsets = {}
                ptrs = (C.c_void_p * self.world)()
                for r in key:
                    ptrs[r] = self.ptrs[r]
                self._subsets[key] = ptrs          # (entries left null are skipped by stereo_peer_push)
        self._check(_capi.lib().stereo_peer_push(self.ctx.handle, ptrs, self.world, int(offset), C.c_void_p(int(src_ptr)),
                                                 int(nbytes), C.c_void_p(int(stream) or CUDA_STREAM_LEGACY)))

    def mark(self) -> int:
        t = C.c_int(-1)
        self._check(_capi.lib().stereo_peer_mark(self.ctx.handle, C.byref(t)))
        return int(t.value)

    def wait(self, ticket: int, stream: int) -> None:
        self._check(_capi.lib().stereo_peer_wait(self.ctx.handle, int(ticket), C.c_void_p(int(stream) or CUDA_STREAM_LEGACY)))

    def local_bytes(self, device):
        """This rank's gather buffer as a uint8 tensor (no copy)."""
        return self.torch.as_tensor(_RawDeviceBytes(self.local_ptr, self.nbytes), device=device)

    def complete(self, stream: int) -> None:
        """Blocks until every rank's pushes so far have landed everywhere (local join + barrier)."""
        self.wait(self.mark(), stream)
        self._check(_capi.lib().stereo_ctx_synchronize(self.ctx.handle, C.c_void_p(int(stream) or CUDA_STREAM_LEGACY)))
        if self.world > 1:
            self.dist.barrier(group=self.group)

    def close(self) -> None:
        lib = _capi.lib()
        if self.world > 1:
            self.dist.barrier(group=self.group)            # nobody may still be pushing into a buffer that goes away
        for p in self._opened:
            lib.stereo_peer_buffer_close(self.ctx.handle, C.c_void_p(p))
        self._opened = []
        if self.local_ptr:
            lib.stereo_peer_buffer_destroy(self.ctx.handle, C.c_void_p(self.local_ptr))
            self.local_ptr = 0


class ShardedStereo:
    """Row-band and pair sharding over the ranks of a ``torch.distributed`` process group."""

    def __init__(self, compute, group=None, gather: str = "collective"):
        """gather = "collective": ``all_gather_into_tensor`` (NCCL / gloo);  "peer": copy-engine pushes into every
        rank's buffer (PeerGather; GPUs of one box only)."""
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.compute, self.group = compute, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if gather not in ("collective", "peer"):
            raise ValueError("gather must be 'collective' or 'peer'")
        self.gather_mode = gather
        self._pg = None

    def _gather_peer(self, mine):
        mine = mine.contiguous()
        n = mine.numel() * mine.element_size()
        if self._pg is None or self._pg.nbytes != self.world * n:
            if self._pg is not None:
                self._pg.close()
            self._pg = PeerGather(self.compute.ctx, self.world * n, group=self.group)
        stream = self.torch.cuda.current_stream(mine.device).cuda_stream
        self._pg.push(self.rank * n, mine.data_ptr(), n, stream)
        self._pg.complete(stream)
        out = self._pg.local_bytes(mine.device).clone()      # the buffer is reused by the next call
        if self.world > 1:
            self.dist.barrier(group=self.group)               # everyone has read before anyone pushes again
        return out.view(mine.dtype).view((self.world,) + tuple(mine.shape))

    def close(self):
        if self._pg is not None:
            self._pg.close()
            self._pg = None

    def _gather(self, mine):
        if self.world == 1:
            return mine.unsqueeze(0)
        if self.gather_mode == "peer":
            return self._gather_peer(mine)
        # output = the ranks' tensors concatenated along dim 0 (the layout both NCCL and gloo accept)
        # gathered as raw bytes: the maps are int8/int16/int32 and gloo has no int16 collectives
        mine = mine.contiguous()
        flat = mine.view(self.torch.uint8).reshape(-1)
        out = self.torch.empty(self.world * flat.numel(), dtype=self.torch.uint8, device=mine.device)
        self.dist.all_gather_into_tensor(out, flat, group=self.group)
        return out.view(mine.dtype).view((self.world,) + tuple(mine.shape))

    # -- one large image, row bands with halo (BASELINE config 4) -----------------------------------
    def disparity_bands(self, cost: int, left: np.ndarray, right: np.ndarray, window_rad: int, min_disp: int,
                        max_disp: int, dtype=None):
        """Full rows x cols disparity map on every rank.  `left`/`right` are uint8 host images; each rank
        touches only its slab of them."""
        torch = self.torch
        dtype = dtype or torch.int16
        rows, cols = left.shape
        band = -(-rows // self.world)
        r0, r1 = band_shard(rows, self.world, self.rank)
        mine = torch.zeros((band, cols), dtype=dtype, device=getattr(self.compute, "device", "cpu"))
        if r1 > r0:
            h0, h1 = band_halo(rows, r0, r1, window_rad)
            mine[:r1 - r0] = self.compute.band(cost, left[h0:h1], right[h0:h1], rows, cols, r0, r1, h0, h1,
                                               window_rad, min_disp, max_disp, dtype)
        return self._gather(mine).reshape(self.world * band, cols)[:rows]

    def disparity_pair_bands(self, cost: int, left: np.ndarray, right: np.ndarray, window_rad: int, disparity_range: int,
                             dtype=None):
        """(left-referenced map, right-referenced map), each rows x cols, on every rank; every rank computes both
        maps of its row band in one launch sequence (BASELINE config 4 for a pair)."""
        torch = self.torch
        dtype = dtype or torch.int16
        rows, cols = left.shape
        band = -(-rows // self.world)
        r0, r1 = band_shard(rows, self.world, self.rank)
        mine = torch.zeros((2, band, cols), dtype=dtype, device=getattr(self.compute, "device", "cpu"))
        if r1 > r0:
            h0, h1 = band_halo(rows, r0, r1, window_rad)
            mine[:, :r1 - r0] = self.compute.pair_band(cost, left[h0:h1], right[h0:h1], rows, cols, r0, r1, h0, h1,
                                                       window_rad, disparity_range, dtype)
        g = self._gather(mine)                                        # world x 2 x band x cols
        return g[:, 0].reshape(self.world * band, cols)[:rows], g[:, 1].reshape(self.world * band, cols)[:rows]

    # -- batches, sharded by pair (BASELINE config 5) ---------------------------------------------------
    def disparity_pair_batch(self, cost: int, lefts: np.ndarray, rights: np.ndarray, window_rad: int,
                             disparity_range: int, dtype=None):
        """(left maps, right maps), each n x rows x cols, on every rank; rank r computes pairs pair_shard(n, world, r)."""
        torch = self.torch
        dtype = dtype or torch.int8
        n, rows, cols = lefts.shape
        per = -(-n // self.world)
        b, e = pair_shard(n, self.world, self.rank)
        mine = torch.zeros((2, per, rows, cols), dtype=dtype, device=getattr(self.compute, "device", "cpu"))
        if e > b:
            mine[:, :e - b] = self.compute.pair_batch(cost, lefts[b:e], rights[b:e], window_rad, disparity_range, dtype)
        g = self._gather(mine)                                        # world x 2 x per x rows x cols
        counts = [pair_shard(n, self.world, r) for r in range(self.world)]
        dl = torch.cat([g[r, 0, :c1 - c0] for r, (c0, c1) in enumerate(counts)])
        dr = torch.cat([g[r, 1, :c1 - c0] for r, (c0, c1) in enumerate(counts)])
        return dl, dr
