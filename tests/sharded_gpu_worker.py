"""torchrun worker for test_two_gpu_sharded_renderer_modes: renders one frame sharded over the ranks
and checks rank 0's frame against the CPU oracle (test infrastructure)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from nettracer_b200 import scenes  # noqa: E402
from nettracer_b200.scene import make_params  # noqa: E402
from nettracer_b200.sharded import CudaBackend, ShardedRenderer  # noqa: E402

mode = sys.argv[1]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
scene, cam = scenes.cornell_box()
w, h = 322, 181
sr = ShardedRenderer(CudaBackend(scene, local), rank, world, band_rows=8, mode=mode)
p = sr.shard_params(w, h, 4, 4, cam.resolve(w, h))
for _ in range(3):
    full = sr.render(p)
torch.cuda.synchronize()
if rank == 0:
    from oracle import oracle
    ref, _ = oracle.render(scene, make_params(w, h, 4, 4, cam.resolve(w, h)))
    got = full.cpu().numpy()
    nbad = int((got != ref).any(axis=-1).sum())
    assert nbad <= 2, nbad
    print("SHARDED_OK", mode, nbad)
dist.barrier()
sr.close()
dist.destroy_process_group()
