#!/usr/bin/env python
"""Scene-level end to end: yt.create_scene(load_uniform_grid(N^3)).render() through the public API.

Times, on one GPU: dataset construction (host), the FIRST render (kd-tree build, cell upload, vertex centring on the
device, table upload, cast, image download, rot180) and the steady-state render of the same scene (bricks resident), plus
a render after a camera move.  Prints one JSON line.

    python tools/bench_scene.py [cells=1024] [image=2048] [frames=5]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import yt_b200 as yt
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    res = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    frames = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    t0 = time.perf_counter()
    arr = np.random.default_rng(0).random((n, n, n), dtype=np.float32) + 0.25
    t1 = time.perf_counter()
    ds = yt.load_uniform_grid({"density": arr}, arr.shape)
    sc = yt.create_scene(ds)
    sc.camera.resolution = (res, res)
    t2 = time.perf_counter()
    im = sc.render()
    t3 = time.perf_counter()
    steady = []
    for _ in range(frames):
        a = time.perf_counter()
        im = sc.render()
        steady.append(time.perf_counter() - a)
    sc.camera.yaw(0.1)
    a = time.perf_counter()
    im2 = sc.render()
    moved = time.perf_counter() - a
    src = sc.get_source()
    out = {"what": "create_scene(load_uniform_grid(%d^3)).render() at %d^2, default transfer function (log density), one B200" % (n, res),
           "host_random_field_s": t1 - t0, "load_and_create_scene_s": t2 - t1, "first_render_s": t3 - t2,
           "steady_render_ms": [1e3 * s for s in steady], "steady_render_ms_median": 1e3 * float(np.median(steady)),
           "render_after_camera_move_ms": 1e3 * moved, "bricks": len(src.volume.tree.traverse()),
           "image_max": float(np.asarray(im)[:, :, :3].max()), "moved_differs": bool(np.abs(np.asarray(im2) - np.asarray(im)).max() > 0)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
