"""Offline model of the flat kernel's tile scheduling (analysis aid, CPU only; uses the oracle's per-sample ray counts).

The lanes of a warp walk their ray trees in lockstep, one tree node per loop iteration (nearest hit, then a shadow query
per light, then the children), so a warp tile costs  o + tau * (1 + lights) * max over its 32 lanes of (tree nodes of the
lane's sample)  - an iteration takes as long as its busiest lane, and with 32 lanes some lane is lit by every light.
(With the cost proportional to the lane's RAYS instead, glass tiles come out 1.4 x too cheap: inside glass few shadow
rays are cast per node.)  W persistent warps take tiles from one sequence (greedy list scheduling, each warp at its own
speed: the profile shows the schedulers are not saturated, so contention is ignored).
Prints the modelled makespan of a whole frame and of an 8-row-band shard for several tile orders, against the two bounds
no tile order can beat: total work / W, and the most expensive single tile.

  python scripts/sim_tile_schedule.py [shards]"""
import heapq
import sys

import numpy as np

sys.path.insert(0, ".")
from nettracer_b200 import scenes  # noqa: E402
from nettracer_b200.renderer import primary_rects  # noqa: E402
from nettracer_b200.scene import make_params, owned_rows  # noqa: E402
from oracle import oracle  # noqa: E402

W_WARPS = 148 * 4 * 8
TAU, OVERHEAD = 1.0, 3.0  # per lane-ray and per tile, in "ray units" (prologue + epilogue ~ 17 % of the instructions)


def tile_costs(cost, rows, twx=4, twy=2):
    """cost [h, w, spp] rays per sample; rows = image rows owned by the shard (in order) -> [tiles_y, tiles_x] of max-lane rays."""
    c = cost[rows]                                   # [vrows, w, spp]
    vr, w, spp = c.shape
    ty, tx = -(-vr // twy), -(-w // twx)
    pad = np.zeros((ty * twy, tx * twx, spp), dtype=c.dtype)
    pad[:vr, :w] = c
    t = pad.reshape(ty, twy, tx, twx, spp).max(axis=(1, 3, 4))
    s = pad.reshape(ty, twy, tx, twx, spp).sum(axis=(1, 3, 4))
    return t.astype(np.float64), s.astype(np.float64)


def makespan(costs_in_order, n_warps=W_WARPS):
    t = TAU * costs_in_order + OVERHEAD
    n = len(t)
    if n <= n_warps:
        return t.max()
    heap = list(t[:n_warps])
    heapq.heapify(heap)
    for c in t[n_warps:]:
        heapq.heappush(heap, heapq.heappop(heap) + c)
    return max(heap)


def main():
    shards = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    scene, cam = scenes.cornell_box()
    w, h, spp, depth = 1920, 1080, 4, 5
    p = make_params(w, h, spp, depth, cam.resolve(w, h))
    cost, st = oracle.sample_costs(scene, p)
    rays = cost[..., 0].astype(np.int64)
    nodes = cost[..., 1].astype(np.int64) * (1 + len(scene.lights))
    rects = primary_rects(scene, p)
    glass = [1, 4, 7]
    y_first = min(int(rects[j][2]) for j in glass)
    print(f"rays per sample: mean {rays.mean():.2f}  p50 {np.percentile(rays, 50):.0f}  p99 {np.percentile(rays, 99):.0f}  max {rays.max()};  "
          f"tree nodes per sample: mean {cost[..., 1].mean():.2f}  p99 {np.percentile(cost[..., 1], 99):.0f}  max {cost[..., 1].max()}")
    results = {}
    for label, n in (("whole frame", 1), (f"1/{shards} frame (8-row bands, shard 0)", shards)):
        rows = owned_rows(h, 8, 0, n)
        tmax, tsum = tile_costs(nodes, rows)
        ty, tx = tmax.shape
        flat = tmax.reshape(-1)
        total = (TAU * flat + OVERHEAD).sum()
        lower = max(total / W_WARPS, TAU * flat.max() + OVERHEAD)
        rot_row = min(int(y_first * len(rows) / h) // 2, ty - 1)
        orders = {
            "row-major": np.arange(len(flat)),
            "rotated to the first glass row (shipped)": np.roll(np.arange(len(flat)), -rot_row * tx),
            "reversed": np.arange(len(flat))[::-1],
            "longest tile first (LPT, needs the costs)": np.argsort(-flat, kind="stable"),
        }
        print(f"\n{label}: {len(flat)} tiles, lane utilisation {tsum.sum() / (32 * flat.sum()):.2f}, "
              f"bounds: work / warps {total / W_WARPS:.1f}, largest tile {TAU * flat.max() + OVERHEAD:.1f} -> {lower:.1f}")
        for name, order in orders.items():
            m = makespan(flat[order])
            results[(label, name)] = m
            print(f"   {name:45s} makespan {m:8.1f}   ({m / lower:.2f} x the bound)")
    a = results[("whole frame", "rotated to the first glass row (shipped)")]
    b = results[(f"1/{shards} frame (8-row bands, shard 0)", "rotated to the first glass row (shipped)")]
    print(f"\nmodelled strong scaling at {shards} shards (shipped order, exchange not included): {a / b:.2f} x")


if __name__ == "__main__":
    main()
