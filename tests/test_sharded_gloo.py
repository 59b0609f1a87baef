"""world_size-2 gloo tests (CPU) of the multi-GPU bookkeeping in nettracer_b200/sharded.py: band
arithmetic, buffer strides, collective call order, rank-0 reassembly.  The device work is replaced
by a TEST-ONLY backend that fills shards from the CPU oracle; the product ships CudaBackend only."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nettracer_b200 import abi, scenes
from nettracer_b200.renderer import deinterleave_host
from nettracer_b200.scene import make_params, shard_rows
from nettracer_b200.sharded import ShardedRenderer


class OracleHostBackend:
    """Test-only stand-in for CudaBackend: CPU memory, shards rendered by the oracle.  "Device pointers" are integers
    that name numpy buffers; peer allocations live in POSIX shared memory so that the second process can map them, and
    the frame-synchronisation flags are emulated by spinning on such a buffer - the protocol order of
    ShardedRenderer (sequence numbers, double buffering, acknowledgement) is exercised for real across two processes."""
    device = torch.device("cpu")

    def __init__(self, scene):
        from multiprocessing import shared_memory
        self._shm_mod = shared_memory
        self.scene = scene
        self.bufs = {}      # base pointer -> uint8 numpy array
        self.shms = {}      # base pointer -> SharedMemory
        self.log = []

    def _register(self, arr, shm=None):
        base = arr.ctypes.data  # the real address: torch tensors made from these buffers answer data_ptr() with it
        self.bufs[base] = arr
        if shm is not None:
            self.shms[base] = shm
        return base

    def _resolve(self, ptr, n):
        for base, arr in self.bufs.items():
            if base <= ptr and ptr + n <= base + arr.size:
                return arr[ptr - base: ptr - base + n]
        raise KeyError(ptr)

    def empty(self, *shape):
        t = torch.zeros(*shape, dtype=torch.uint8)
        self._register(t.numpy().reshape(-1))
        return t

    # -- flags --
    def _word(self, ptr):
        return self._resolve(ptr, 4).view(np.uint32)

    def _spin(self, ptr, value):
        import time
        t0 = time.time()
        while np.int32(np.uint32(int(self._word(ptr)[0]) - int(value) & 0xffffffff)) < 0:
            assert time.time() - t0 < 30, "flag wait timed out"
            time.sleep(0.0005)

    def render_shard(self, params, out_ptr, row_stride, sync=None):
        from oracle import oracle
        if sync is not None and sync.post_at_start:
            self._word(sync.post_at_start)[0] = sync.post_at_start_value
        if sync is not None and sync.wait_before_store:
            self._spin(sync.wait_before_store, sync.wait_value)
        rows = shard_rows(params.height, params.band_rows, params.shard_index, params.shard_count)
        assert row_stride == params.width * 4
        if rows:
            img, _ = oracle.render(self.scene, params, compact_rows=rows if params.layout == abi.NT_LAYOUT_COMPACT else 0)
            if params.layout == abi.NT_LAYOUT_COMPACT:
                self._resolve(out_ptr, rows * row_stride)[:] = img.reshape(-1)
            else:
                from nettracer_b200.scene import owned_rows
                out = self._resolve(out_ptr, params.height * row_stride).reshape(params.height, params.width, 4)
                ys = owned_rows(params.height, params.band_rows, params.shard_index, params.shard_count)
                out[ys] = img[ys]
        self.log.append(("render", out_ptr))
        if sync is not None and sync.post_when_done:
            self._word(sync.post_when_done)[0] = sync.post_when_done_value

    def wait_flags(self, flags_ptr, n, value):
        for i in range(n):
            self._spin(flags_ptr + 4 * i, value)

    def deinterleave(self, compact_all, shard_stride, full, width, height, band_rows, world):
        assert shard_stride == compact_all.shape[1] * width * 4
        parts = [compact_all[i, :shard_rows(height, band_rows, i, world)].numpy() for i in range(world)]
        full.copy_(torch.from_numpy(deinterleave_host(parts, height, width, band_rows)))

    # -- peer memory = named shared memory --
    def peer_alloc(self, nbytes):
        shm = self._shm_mod.SharedMemory(create=True, size=nbytes)
        arr = np.frombuffer(shm.buf, dtype=np.uint8, count=nbytes)
        arr[:] = 0
        return self._register(arr, shm)

    @staticmethod
    def _drop(shm, unlink):
        if unlink:
            shm.unlink()
        try:
            shm.close()
        except BufferError:  # numpy views of the mapping are still alive: the mapping goes with the process
            shm.close = lambda: None

    def peer_free(self, ptr):
        shm = self.shms.pop(ptr)
        del self.bufs[ptr]
        self._drop(shm, True)

    def ipc_export(self, ptr):
        arr = self.bufs[ptr]
        return (self.shms[ptr].name + ":" + str(arr.size)).encode().ljust(64, b"\0")

    def ipc_open(self, handle):
        name, size = handle.rstrip(b"\0").decode().split(":")
        shm = self._shm_mod.SharedMemory(name=name)
        return self._register(np.frombuffer(shm.buf, dtype=np.uint8, count=int(size)), shm)

    def ipc_close(self, ptr):
        shm = self.shms.pop(ptr)
        del self.bufs[ptr]
        self._drop(shm, False)

    def wrap(self, ptr, height, width):
        return torch.from_numpy(self._resolve(ptr, height * width * 4).reshape(height, width, 4))

    # -- shared host frame --
    def host_frame_open(self, name, nbytes, world, create):
        shm = self._shm_mod.SharedMemory(name=name.strip("/"), create=create, size=nbytes + 8192)
        arr = np.frombuffer(shm.buf, dtype=np.uint8, count=nbytes + 8192)
        if create:
            arr[:8192] = 0
        base = self._register(arr, shm)
        return base, base + 8192

    def host_frame_close(self, h, unlink):
        shm = self.shms.pop(h)
        del self.bufs[h]
        self._drop(shm, unlink)

    def host_post(self, h, rank, seq):
        self._word(h + 64 * (2 + rank))[0] = seq

    def host_wait_all(self, h, seq, timeout_ms=0):
        for r in range(2):
            self._spin(h + 64 * (2 + r), seq)

    def host_ack(self, h, seq):
        self._word(h + 64)[0] = seq & 0xffffffff

    def host_wait_ack(self, h, seq, timeout_ms=0):
        self._spin(h + 64, seq & 0xffffffff)

    def render_to_host_frame(self, params, h, pixels_ptr, row_stride, rank, seq):
        self.render_shard(params, pixels_ptr, row_stride)
        self.host_post(h, rank, seq)

    def stats(self):
        return {}

    def host_view(self, px, h, w):
        return self._resolve(px, h * w * 4).reshape(h, w, 4)

    def synchronize(self):
        pass

    def close(self):
        pass


def _worker(rank, world, port, w, h, band, out_path, mode="gather"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scene, cam = scenes.cornell_box()
        if mode == "host":
            sr = ShardedRenderer(OracleHostBackend(scene), rank, world, band_rows=band, mode="p2p_store")
            p = sr.shard_params(w, h, 1, 2, cam.resolve(w, h))
            for _ in range(3):
                full, _ = sr.render_host(p)
            if rank == 0:
                np.save(out_path, np.array(full))
            else:
                assert full is None
            sr.close()
            return
        sr = ShardedRenderer(OracleHostBackend(scene), rank, world, band_rows=band, mode=mode)
        p = sr.shard_params(w, h, 1, 2, cam.resolve(w, h))
        assert p.shard_index == rank and p.shard_count == world
        assert p.layout == (abi.NT_LAYOUT_COMPACT if mode == "gather" else abi.NT_LAYOUT_FULL)
        frames = []
        for _ in range(5 if mode == "p2p_store" else 2):  # buffers are reused (p2p_store: two alternate)
            full = sr.render(p)
            frames.append(None if full is None else full.data_ptr())
        if rank == 0:
            np.save(out_path, full.numpy())
            if mode == "p2p_store":
                assert frames[0] == frames[2] == frames[4] != frames[1] == frames[3]  # double buffering
        else:
            assert full is None
        sr.close()
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("h,band,mode", [(45, 8, "gather"), (32, 16, "gather"), (9, 4, "gather"),
                                         (45, 8, "p2p_store"), (9, 16, "p2p_store"), (45, 8, "host")])
def test_sharded_modes_world2_gloo(tmp_path, h, band, mode):
    """(9, 16): rank 1 owns no row and must still run the flag protocol."""
    from oracle import oracle
    w = 40
    out = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(2, _free_port(), w, h, band, out, mode), nprocs=2, join=True)
    scene, cam = scenes.cornell_box()
    ref, _ = oracle.render(scene, make_params(w, h, 1, 2, cam.resolve(w, h)))
    assert np.array_equal(np.load(out), ref)
