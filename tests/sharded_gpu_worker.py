"""torchrun worker for the sharded-render tests (test infrastructure): renders a short SEQUENCE of different frames
sharded over the ranks - so that a frame landing in the wrong buffer, or being overwritten before rank 0 has read it,
shows - and checks every frame rank 0 receives against the CPU oracle.

  sharded_gpu_worker.py <gather|p2p_store|host> [same_gpu]

same_gpu: every rank uses cuda:0 and the process group is gloo (NCCL refuses two ranks on one device).  CUDA IPC between
two processes on one GPU, the flag protocol and the shared host frame are exactly the multi-GPU code path, so a 1-GPU box
exercises it."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from nettracer_b200 import scenes  # noqa: E402
from nettracer_b200.scene import Camera, make_params  # noqa: E402
from nettracer_b200.sharded import CudaBackend, ShardedRenderer  # noqa: E402

mode = sys.argv[1]
same_gpu = len(sys.argv) > 2 and sys.argv[2] == "same_gpu"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
if same_gpu:
    local = 0
torch.cuda.set_device(local)
if same_gpu:
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
scene, cam0 = scenes.cornell_box()
w, h = 322, 181
cams = [Camera(eye=(0.0 + 0.7 * i, 5.0, 15.0 - 0.5 * i), at=(0.0, 3.2, 0.0), vfov_deg=42.0) for i in range(5)]
sr = ShardedRenderer(CudaBackend(scene, local), rank, world, band_rows=8, mode="p2p_store" if mode == "host" else mode)
got = []
if mode == "host":
    for cam in cams:
        full, st = sr.render_host(sr.shard_params(w, h, 4, 4, cam.resolve(w, h)))
        if rank == 0:
            got.append(np.array(full))      # copy: the shared frame is reused by the next call
else:
    # enqueue the whole sequence without a host synchronisation in between: rank 0 copies each frame out on its
    # stream, the peers run ahead as far as the acknowledgement protocol lets them
    outs = []
    for cam in cams:
        full = sr.render(sr.shard_params(w, h, 4, 4, cam.resolve(w, h)))
        if rank == 0:
            outs.append(full.clone())
    torch.cuda.synchronize()
    st = sr.b.stats()
    got = [o.cpu().numpy() for o in outs]
if rank == 0:
    from oracle import oracle
    worst = 0
    for cam, img in zip(cams, got):
        ref, _ = oracle.render(scene, make_params(w, h, 4, 4, cam.resolve(w, h)))
        nbad = int((img != ref).any(axis=-1).sum())
        worst = max(worst, nbad)
        assert nbad <= 2, nbad
    print("SHARDED_OK", mode, worst)
dist.barrier()
sr.close()
dist.destroy_process_group()
