#!/usr/bin/env python
"""Would early ray termination pay?  

The reference composites back to front.  A front-to-back walk of the same samples, (A, B) <- (A + B a_k, B b_k), may stop
once max_c B_c < eps: whatever lies behind contributes less than eps * (bounded colour).  This tool MEASURES the final
transmittance B of every ray of the benchmark volume (the partial cast already computes it: C -> A + B C) for

  * bench      bench.py's ColorTransferFunction (4 gaussians, alpha up to 100), non-grey
  * default    the transfer function yt.create_scene builds (TransferFunctionHelper.setup_default: 10 layers of the
               colormap over the field's log range), non-grey
  * grey       the bench table with grey_opacity=True (alpha attenuates all channels: the opaque case)
  * grey_x30   the same with alpha 30x larger (an 'opaque surfaces' render)

and prints, per table, the distribution of max_c B_c over the rays that cross the volume and the fraction of SAMPLES a
perfect ERT at eps = 1e-6 / 1e-4 could skip, assuming (optimistically) that a ray's opacity accumulates uniformly
along its length:  skipped fraction of a ray = max(0, 1 - ln(eps) / ln(B_final)).

    python tools/ert_histogram.py [cells=512] [image=1024]
"""
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
    from yt_b200.volume_rendering.transfer_function_helper import TransferFunctionHelper
    from yt_b200.volume_rendering.transfer_functions import ColorTransferFunction
    from yt_b200.volume_rendering.utils import sampler_arguments
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    res = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    dev = torch.device("cuda", 0)
    cam = bench_camera(res)
    vol = synthetic.vertex_slab(n, [0, 0, 0], [n, n, n], SEED, dev)
    torch.cuda.synchronize()

    def scaled(tf, factor, grey):
        out = ColorTransferFunction(tf.x_bounds, tf.nbins, grey_opacity=grey)
        for c in range(4):
            out.funcs[c].y[:] = tf.funcs[c].y * (factor if c == 3 else 1.0)
        return out

    # the default table of create_scene over this field's range (the helper works on log10 of the data)
    lo, hi = float(vol.min()), float(vol.max())
    tfh = TransferFunctionHelper(None)
    tfh.bounds, tfh.log = (lo, hi), False
    tfh.build_transfer_function()
    tfh.setup_default()
    tables = {"bench": bench_tf(), "default": tfh.tf, "grey": scaled(bench_tf(), 1.0, True), "grey_x30": scaled(bench_tf(), 30.0, True)}
    print("volume %d^3 (benchmark field), image %d^2, n_samples %d" % (n, res, N_SAMPLES))
    for name, tf in tables.items():
        args, kw = sampler_arguments(cam)
        kw.pop("camera_data")
        s = S.VolumeRenderSampler(*args, tf, N_SAMPLES, **kw)
        bs = D.BrickSet(s._ctx, 1, 1, L.VERTEX_CENTRED, L.STORE_F32)
        bs.attach(0, [vol.data_ptr()], 0, [0.0] * 3, [1.0] * 3, [n] * 3, keep=vol)
        bs.commit()
        part = torch.zeros((res * res, 8), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        s._frame.clear()
        s._frame.cast_partial(s.kind, bs, np.arange(1, dtype=np.int64), part.data_ptr(), tf=s._tables, n_samples=N_SAMPLES, f32=True)
        s._ctx.synchronize()
        p = part.cpu().numpy()
        A, B = p[:, :4], p[:, 4:]
        nb = 4 if tf.grey_opacity else 3
        hit = np.abs(A[:, :3]).sum(axis=1) > 0
        bmax = B[hit, :nb].max(axis=1).astype("float64")
        q = np.quantile(bmax, [0.0, 0.01, 0.1, 0.5, 0.9, 1.0])
        line = "%-9s rays %8d  max_c B: min %.3g  p1 %.3g  p10 %.3g  median %.3g  p90 %.3g  max %.3g" % ((name, hit.sum()) + tuple(q))
        for eps in (1e-6, 1e-4):
            with np.errstate(divide="ignore", invalid="ignore"):
                frac = np.where(bmax < eps, 1.0 - np.log(eps) / np.log(np.maximum(bmax, 1e-300)), 0.0)
            line += "  | eps %g: rays below %.4f, samples skippable %.4f" % (eps, (bmax < eps).mean(), frac.mean())
        print(line)


if __name__ == "__main__":
    main()
