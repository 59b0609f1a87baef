"""Multi-GPU image sharding: one process per GPU, the scene replicated, image rows
cut into bands of `band_rows` rows dealt round-robin to the ranks (band b -> rank b mod N), and ONE
exchange per frame that lands the finished RGBA8 rows on rank 0 over NVLink.  This replaces the
reference's thread/TCP tile distribution (hypothesised only — /root/reference/README:1-3 holds no
code).  No exchange happens while rays are traced.

Two exchange modes:
  "gather"    every rank renders its bands into a compact device buffer; `torch.distributed.gather`
              (NCCL send/recv over NVLink) brings them to rank 0, where nt_deinterleave_device
              scatters them into the full frame.
  "p2p_store" rank 0's frame buffer is opened on every rank through CUDA IPC and each rank's render
              kernel stores its pixels straight into it (peer stores over NVLink overlap with the
              tracing); one tiny all-reduce per frame orders "all shards written" before rank 0 reads.

The backend object does the device work, so the bookkeeping (band arithmetic, buffer strides,
collective call order) can be exercised by world_size-2 gloo tests on CPU with a test-only backend;
the product backend is `CudaBackend` and nothing else ships.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import abi
from .scene import make_params, shard_rows


class CudaBackend:
    """Device work for one rank through the C ABI (the only backend the package provides)."""

    def __init__(self, scene, device_index: int):
        from .lib import check, load
        from .renderer import Renderer
        self._check, self._lib = check, load()
        self.device_index = int(device_index)
        self.device = torch.device("cuda", self.device_index)
        torch.cuda.set_device(self.device)
        self.renderer = Renderer(scene, self.device_index)

    def empty(self, *shape):
        return torch.empty(*shape, dtype=torch.uint8, device=self.device)

    def stream_ptr(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def render_shard(self, params: abi.nt_render_params, out_ptr: int, row_stride: int):
        self.renderer.render_device(params, out_ptr, row_stride, self.stream_ptr())

    def deinterleave(self, compact_all, shard_stride: int, full, width, height, band_rows, world):
        self._check(self._lib.nt_deinterleave_device(C.c_void_p(compact_all.data_ptr()), shard_stride,
                                                     C.c_void_p(full.data_ptr()), width * 4, width, height,
                                                     band_rows, world, self.device_index,
                                                     C.c_void_p(self.stream_ptr())))

    # ---- peer frame buffer (p2p_store) ----
    def peer_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self._lib.nt_device_malloc(self.device_index, nbytes, C.byref(p)))
        return p.value

    def peer_free(self, ptr: int):
        self._lib.nt_device_free(self.device_index, C.c_void_p(ptr))

    def ipc_export(self, ptr: int) -> bytes:
        h = (C.c_uint8 * 64)()
        self._check(self._lib.nt_ipc_export(C.c_void_p(ptr), self.device_index, h))
        return bytes(h)

    def ipc_open(self, handle: bytes) -> int:
        h = (C.c_uint8 * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        self._check(self._lib.nt_ipc_open(h, self.device_index, C.byref(p)))
        return p.value

    def ipc_close(self, ptr: int):
        self._lib.nt_ipc_close(C.c_void_p(ptr), self.device_index)

    def wrap(self, ptr: int, height: int, width: int):
        class _Raw:
            __cuda_array_interface__ = {"shape": (height, width, 4), "typestr": "|u1", "data": (ptr, False),
                                        "version": 3, "strides": None}
        return torch.as_tensor(_Raw(), device=self.device)

    def stats(self) -> dict:
        return self.renderer.device_stats(self.stream_ptr())

    def close(self):
        self.renderer.close()


class ShardedRenderer:
    def __init__(self, backend, rank: int, world: int, band_rows: int = 16, mode: str = "gather", group=None):
        assert mode in ("gather", "p2p_store")
        self.b, self.rank, self.world, self.band_rows, self.group = backend, rank, world, band_rows, group
        self.mode = mode if world > 1 else "single"
        self._shape = None
        self._peer_base = None   # rank 0: owned allocation; other ranks: opened IPC mapping
        self._flag = None
        self.full = self.shard = self.slots = None

    # -- buffers sized per (width, height) --
    def _prepare(self, width, height):
        if self._shape == (width, height):
            return
        self._release_peer()
        self._shape = (width, height)
        self.max_rows = max(shard_rows(height, self.band_rows, r, self.world) for r in range(self.world))
        self.my_rows = shard_rows(height, self.band_rows, self.rank, self.world)
        if self.mode == "single":
            self.full = self.b.empty(height, width, 4)
        elif self.mode == "gather":
            # equal-sized slots so one gather moves everything; the tail rows of short shards are unused
            self.shard = self.b.empty(self.max_rows, width, 4)
            if self.rank == 0:
                self.slots = self.b.empty(self.world, self.max_rows, width, 4)
                self.full = self.b.empty(height, width, 4)
        else:
            handle = torch.zeros(64, dtype=torch.uint8)
            if self.rank == 0:
                self._peer_base = self.b.peer_alloc(height * width * 4)
                handle = torch.tensor(list(self.b.ipc_export(self._peer_base)), dtype=torch.uint8)
                self.full = self.b.wrap(self._peer_base, height, width)
            handle = self._bcast_bytes(handle)
            if self.rank != 0:
                self._peer_base = self.b.ipc_open(bytes(handle.tolist()))
            self._flag = self.b.empty(4).zero_()

    def _bcast_bytes(self, t):
        if dist.get_backend(self.group) == "nccl":
            d = t.to(self.b.device)
            dist.broadcast(d, src=0, group=self.group)
            return d.cpu()
        dist.broadcast(t, src=0, group=self.group)
        return t

    def _release_peer(self):
        if self._peer_base is not None:
            if self.rank == 0:
                self.full = None
                self.b.peer_free(self._peer_base)
            else:
                self.b.ipc_close(self._peer_base)
            self._peer_base = None

    def shard_params(self, width, height, spp, max_depth, camera, precision=abi.NT_F64_STRICT, ray_epsilon=0.0):
        layout = abi.NT_LAYOUT_COMPACT if self.mode == "gather" else abi.NT_LAYOUT_FULL
        return make_params(width, height, spp, max_depth, camera, precision, ray_epsilon, self.rank, self.world,
                           self.band_rows, layout)

    def render(self, params: abi.nt_render_params, kernel_done=None):
        """Enqueue one frame on the current stream.  Returns the full device frame on rank 0, None
        elsewhere.  Asynchronous: synchronise the stream before reading.  `kernel_done` (an event
        with .record()) is recorded between this rank's render kernel and the exchange."""
        w, h = params.width, params.height
        self._prepare(w, h)
        if self.mode == "single":
            self.b.render_shard(params, self.full.data_ptr(), w * 4)
            if kernel_done is not None:
                kernel_done.record()
            return self.full
        if self.mode == "gather":
            if self.my_rows:
                self.b.render_shard(params, self.shard.data_ptr(), w * 4)
            if kernel_done is not None:
                kernel_done.record()
            if self.rank == 0:
                dist.gather(self.shard, list(self.slots.unbind(0)), dst=0, group=self.group)
                self.b.deinterleave(self.slots, self.max_rows * w * 4, self.full, w, h, self.band_rows, self.world)
                return self.full
            dist.gather(self.shard, None, dst=0, group=self.group)
            return None
        if self.my_rows:
            self.b.render_shard(params, self._peer_base, w * 4)
        if kernel_done is not None:
            kernel_done.record()
        dist.all_reduce(self._flag, group=self.group)  # "every shard is written", stream-ordered
        return self.full if self.rank == 0 else None

    def close(self):
        self._release_peer()
        self.b.close()
