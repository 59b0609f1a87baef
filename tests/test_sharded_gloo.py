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
    """Test-only stand-in for CudaBackend: CPU tensors, shards rendered by the oracle."""
    device = torch.device("cpu")

    def __init__(self, scene):
        self.scene = scene
        self.bufs = {}

    def empty(self, *shape):
        t = torch.zeros(*shape, dtype=torch.uint8)
        self.bufs[t.data_ptr()] = t
        return t

    def render_shard(self, params, out_ptr, row_stride):
        from oracle import oracle
        out = self.bufs[out_ptr]
        rows = shard_rows(params.height, params.band_rows, params.shard_index, params.shard_count)
        img, _ = oracle.render(self.scene, params, compact_rows=rows)
        assert params.layout == abi.NT_LAYOUT_COMPACT and row_stride == params.width * 4
        out[:rows] = torch.from_numpy(img)

    def deinterleave(self, compact_all, shard_stride, full, width, height, band_rows, world):
        assert shard_stride == compact_all.shape[1] * width * 4
        parts = [compact_all[i, :shard_rows(height, band_rows, i, world)].numpy() for i in range(world)]
        full.copy_(torch.from_numpy(deinterleave_host(parts, height, width, band_rows)))

    def close(self):
        pass


def _worker(rank, world, port, w, h, band, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scene, cam = scenes.cornell_box()
        sr = ShardedRenderer(OracleHostBackend(scene), rank, world, band_rows=band, mode="gather")
        p = sr.shard_params(w, h, 1, 2, cam.resolve(w, h))
        assert p.shard_index == rank and p.shard_count == world and p.layout == abi.NT_LAYOUT_COMPACT
        for _ in range(2):  # twice: buffers are reused
            full = sr.render(p)
        if rank == 0:
            np.save(out_path, full.numpy())
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


@pytest.mark.parametrize("h,band", [(45, 8), (32, 16), (9, 4)])
def test_gather_mode_world2_gloo(tmp_path, h, band):
    from oracle import oracle
    w = 40
    out = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(2, _free_port(), w, h, band, out), nprocs=2, join=True)
    scene, cam = scenes.cornell_box()
    ref, _ = oracle.render(scene, make_params(w, h, 1, 2, cam.resolve(w, h)))
    assert np.array_equal(np.load(out), ref)
