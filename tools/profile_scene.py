#!/usr/bin/env python
"""Where a steady-state Scene.render() spends its host time: cProfile of a few frames after the first (bricks resident),
plus the kernel time of the same cast.  python tools/profile_scene.py [cells=1024] [image=2048]"""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import yt_b200 as yt
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    res = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    arr = np.random.default_rng(0).random((n, n, n), dtype=np.float32) + 0.25
    ds = yt.load_uniform_grid({"density": arr}, arr.shape)
    sc = yt.create_scene(ds)
    sc.camera.resolution = (res, res)
    sc.render()
    sc.render()
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    for _ in range(4):
        sc.render()
    pr.disable()
    print("4 frames: %.1f ms per frame" % (1e3 * (time.perf_counter() - t0) / 4))
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
    print(s.getvalue())


if __name__ == "__main__":
    main()
