#!/usr/bin/env python
"""Throughput of the on-device vertex centring (ytb200_brick_from_cells) against the HBM roofline.

    python tools/bench_vertex_centre.py [--n 1024] [--steps 5]

One JSON line: algorithmic bytes = n^3 * sizeof(cell) read + (n+1)^3 * sizeof(vertex) written, divided by the
CUDA-event time of the kernel; peak = MEASURED_PEAKS.json.hbm_gbs.  The host baseline beside it is the numpy
mirror (loaders.vertex_centre_no_ghost) on a 256^3 sample."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--log", action="store_true")
    args = ap.parse_args()
    import torch
    from yt_b200 import _lib as L
    from yt_b200 import device as D
    from yt_b200 import loaders, synthetic
    n = args.n
    dev = torch.device("cuda", 0)
    ctx = D.Context(0)
    out = {}
    for label, f32 in (("f32_cells", True), ("f64_cells", False)):
        cells = synthetic.cell_planes(n, 0, n, 0, dev) + 0.05
        if not f32:
            cells = cells.double()
        torch.cuda.synchronize()
        times = []
        for _ in range(args.steps + 2):
            bs = D.BrickSet(ctx, 1, 1, L.VERTEX_CENTRED, L.STORE_F32)
            # (the call's CUDA-event bracket includes the cudaMalloc of the output; `ncu` gives the pure kernel time)
            st = bs.from_cells(0, [cells.data_ptr()], [n] * 3, [0] * 3, [n] * 3, [0.0] * 3, [1.0] * 3,
                               log_fields=[args.log], cells_are_f32=f32, keep=cells, timed=True)
            times.append(st.kernel_ms)
            del bs
        ms = float(np.mean(times[2:]))
        nbytes = n ** 3 * (4 if f32 else 8) + (n + 1) ** 3 * 4
        out[label] = {"kernel_ms": ms, "GB/s": nbytes / (ms * 1e-3) / 1e9, "algorithmic_bytes": nbytes}
        del cells
    peak = 6650.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    m = 256
    host = np.random.default_rng(0).random((m, m, m))
    t0 = time.perf_counter()
    v = loaders.vertex_centre_no_ghost(host).astype("float32")
    dt = time.perf_counter() - t0
    line = {"metric": "vertex centring GB/s", "n": n, "log10": bool(args.log), "results": out,
            "roofline": {"bound": "hbm", "achieved": out["f32_cells"]["GB/s"], "peak": peak, "unit": "GB/s",
                         "frac": out["f32_cells"]["GB/s"] / peak},
            "cpu_baseline": {"kind": "port (numpy mirror, 1 thread)", "sample": "%d^3 cells" % m,
                             "GB/s": (m ** 3 * 8 + (m + 1) ** 3 * 4) / dt / 1e9}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
