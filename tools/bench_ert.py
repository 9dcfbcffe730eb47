#!/usr/bin/env python
"""Early ray termination on the north-star volume: plain cast against the terminated cast (eps = 1e-6) for the tables of
tools/ert_histogram.py -- the benchmark's table with grey opacity, and the same 30x more opaque.  One JSON line per
table: kernel time of both casts, the cells the terminated cast sampled of the algorithmic count, and the normalised
difference of the two images.

    python tools/bench_ert.py [cells=1024] [image=2048]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from bench import N_SAMPLES, SEED, bench_camera, bench_tf
    from yt_b200 import _lib as L
    from yt_b200 import device as D
    from yt_b200 import image_samplers as S
    from yt_b200 import synthetic
    from yt_b200.metrics import normalised_diff
    from yt_b200.volume_rendering.transfer_functions import ColorTransferFunction
    from yt_b200.volume_rendering.utils import sampler_arguments
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    res = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    eps = 1e-6
    dev = torch.device("cuda", 0)
    cam = bench_camera(res)
    vol = synthetic.vertex_slab(n, [0, 0, 0], [n, n, n], SEED, dev)
    torch.cuda.synchronize()

    def scaled(tf, factor):
        out = ColorTransferFunction(tf.x_bounds, tf.nbins, grey_opacity=True)
        for c in range(4):
            out.funcs[c].y[:] = tf.funcs[c].y * (factor if c == 3 else 1.0)
        return out

    for name, tf in (("grey", scaled(bench_tf(), 1.0)), ("grey_x30", scaled(bench_tf(), 30.0))):
        res_ms, images, cells, total = {}, {}, {}, None
        for ert in (0.0, eps):
            args, kw = sampler_arguments(cam)
            kw.pop("camera_data")
            s = S.VolumeRenderSampler(*args, tf, N_SAMPLES, early_termination=ert, **kw)
            bs = D.BrickSet(s._ctx, 1, 1, L.VERTEX_CENTRED, L.STORE_F32)
            bs.attach(0, [vol.data_ptr()], 0, [0.0] * 3, [1.0] * 3, [n] * 3, keep=vol)
            bs.commit()
            order = np.arange(1, dtype=np.int64)
            fr = s._frame
            ms = []
            for it in range(6):
                fr.clear()
                st = fr.cast(s.kind, bs, order, tf=s._tables, n_samples=N_SAMPLES, timed=True)
                if it:
                    ms.append(st.kernel_ms)
            res_ms[ert], cells[ert] = float(np.mean(ms)), int(st.n_cells)
            images[ert] = fr.download(np.empty((res, res, 4)))
            if total is None:
                total = int(fr.count(s.kind, bs, order, N_SAMPLES).n_cells)
        print(json.dumps({
            "what": "early ray termination, %d^3 volume @ %d^2, n_samples %d, table %s (grey opacity)" % (n, res, N_SAMPLES, name),
            "eps": eps, "plain_kernel_ms": res_ms[0.0], "terminated_kernel_ms": res_ms[eps],
            "cells_algorithmic": total, "cells_sampled": cells[eps], "cells_skipped_frac": 1.0 - cells[eps] / total,
            "gsamples_per_s_algorithmic": total * N_SAMPLES / res_ms[eps] / 1e6,
            "normalised_diff_terminated_vs_plain": normalised_diff(images[eps], images[0.0], True)}))


if __name__ == "__main__":
    main()
