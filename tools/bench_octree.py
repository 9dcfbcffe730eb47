#!/usr/bin/env python
"""Throughput of the octree cast (ytb200_frame_cast_octree) next to the reference's cast_through_octree on the host.

    python tools/bench_octree.py [--res 1024] [--steps 5] [--base-level 6]

One JSON line: an adaptively refined octree over the unit cube (base level cells, two more levels inside nested
spheres), perspective camera outside the domain, ColorTransferFunction, n_samples = 10.  GPU: CUDA-event time of the
cast kernel and the end-to-end sampler call (upload of image + cells, cast, download).  CPU: the compiled reference
(oracle/_ref, kind "reference") or the C oracle (kind "port") on a bounded 256^2 image with all host threads, in the
same unit (rays per second; every ray of both images crosses the domain centre region the same way)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def refined_octree(base_level, depth):
    """(ipos int32 (n,3), lvl int32 (n), centre (n,3), dx (n)): level `base_level` cells, split once inside r < 0.3 and
    twice inside r < 0.15 of the domain centre."""
    n = 1 << base_level
    g = np.stack(np.meshgrid(*[np.arange(n)] * 3, indexing="ij"), -1).reshape(-1, 3).astype("int64")
    lvl = np.full(len(g), base_level, "int64")
    kids = np.stack(np.meshgrid(*[np.arange(2)] * 3, indexing="ij"), -1).reshape(-1, 3)
    for radius in (0.3, 0.15):
        centre = (g + 0.5) / (1 << lvl)[:, None]
        split = (np.linalg.norm(centre - 0.5, axis=1) < radius) & (lvl == lvl.max())
        child = (2 * g[split][:, None, :] + kids[None, :, :]).reshape(-1, 3)
        g = np.concatenate([g[~split], child])
        lvl = np.concatenate([lvl[~split], np.full(len(child), lvl.max() + 1, "int64")])
    dx = 1.0 / (1 << lvl).astype("float64")
    centre = (g + 0.5) * dx[:, None]
    ipos = np.floor(centre * (1 << depth)).astype("int64").astype("int32")
    return ipos, lvl.astype("int32"), centre, dx


def cells14(centre, dx):
    out = np.empty((len(dx), 14))
    c = 0
    for a in range(2):
        for b in range(2):
            for k in range(2):
                v = centre + (np.array([a, b, k]) - 0.5)[None, :] * dx[:, None]
                out[:, c] = 0.5 + 0.3 * np.sin(9.0 * v[:, 0]) * np.cos(7.0 * v[:, 1]) + 0.15 * np.sin(11.0 * v[:, 2])
                c += 1
    out[:, 8:11] = centre - dx[:, None] / 2
    out[:, 11:14] = centre + dx[:, None] / 2
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--base-level", type=int, default=6)
    ap.add_argument("--cpu-res", type=int, default=256)
    args = ap.parse_args()
    from oracle import oracle as O
    from tests import common as T
    from yt_b200 import _lib as L
    from yt_b200 import image_samplers as S
    from yt_b200 import partitioned_grid as G
    from yt_b200.octree_raytracing import _OctreeRayTracing
    depth = args.base_level + 2
    ipos, lvl, centre, dx = refined_octree(args.base_level, depth)
    cells = cells14(centre, dx)
    keys = np.arange(len(lvl), dtype=np.int32)
    tf = T.colour_tf(n_gauss=4)

    def make(res, mod, grid_mod, tree_cls):
        cam = T.make_camera("perspective", (res, res), position=[1.7, 1.5, 1.3], width=[0.8, 0.8, 1.0])
        a, kw = T.sampler_args(cam)
        a[8] = "Octree"
        s = mod.VolumeRenderSampler(*a, tf, 10, **kw)
        tree = tree_cls(np.zeros(3), np.ones(3), depth)
        tree.add_nodes(ipos, lvl, keys)
        pg = grid_mod.PartitionedGrid(0, cells.reshape(1, -1, 14, 1), np.ones((len(lvl), 14, 1), "uint8"),
                                      np.ascontiguousarray(cells[:, 8:11]).ravel(), np.ascontiguousarray(cells[:, 11:14]).ravel(),
                                      np.ones(3, "int64"), n_fields=1)
        return s, tree, pg

    s, tree, pg = make(args.res, S, G, _OctreeRayTracing)
    e2e, kern = [], []
    for _ in range(args.steps + 3):
        t0 = time.perf_counter()
        s(pg, oct=tree)
        e2e.append((time.perf_counter() - t0) * 1e3)
        st = L.Stats()
        L.check(L.load().ytb200_frame_cast_octree(s._frame.handle, int(s.kind), tree.handle, s._tables.handle, 10, int(s._precision), C.byref(st)))
        kern.append(st.kernel_ms)
    rays = args.res * args.res
    kernel_ms, e2e_ms = float(np.mean(kern[3:])), float(np.mean(e2e[3:]))
    # host baseline on a bounded image
    if O.ref_available():
        ref_s, ref_g = O.load_ref()
        from yt.utilities.lib._octree_raytracing import _OctreeRayTracing as RefTree
        kind, (cs, ctree, cpg) = "reference", make(args.cpu_res, ref_s, ref_g, RefTree)
    else:
        kind, (cs, ctree, cpg) = "port", make(args.cpu_res, O, O, O._OctreeRayTracing)
    best = 1e30
    for _ in range(2):
        t0 = time.perf_counter()
        cs(cpg, oct=ctree, num_threads=0)
        best = min(best, time.perf_counter() - t0)
    cpu_rays = args.cpu_res * args.cpu_res / best
    print(json.dumps({
        "metric": "octree cast Mrays/s", "value": rays / (kernel_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_frame": kernel_ms,
        "config": {"workload": "octree_vr", "cells": int(len(lvl)), "levels": [int(lvl.min()), int(lvl.max())], "nodes": tree.node_count,
                   "image": [args.res, args.res], "n_samples": 10, "lens": "perspective"},
        "e2e": {"value": rays / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_frame": e2e_ms,
                "note": "sampler(pg, oct=tree): upload of image, z-buffer and the cell array, cast, download"},
        "cpu_baseline": {"value": cpu_rays / 1e6, "unit": "Mrays/s", "cores": O.max_threads(), "kind": kind,
                         "sample": "%d^2 image, same tree / camera / tables" % args.cpu_res}}))


if __name__ == "__main__":
    main()
