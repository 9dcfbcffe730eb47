"""Diagnostic: where does the cfg2 (512^3 @ 1024^2) FAST frame differ from the reference?"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests import common as T
from bench import N_SAMPLES, bench_camera, bench_tf
from yt_b200 import _lib as L, device as D, image_samplers as S, synthetic
from oracle import oracle as O
N, R = int(sys.argv[1]) if len(sys.argv) > 1 else 512, int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device("cuda", 0)
cam, tf = bench_camera(R), bench_tf()
vol = synthetic.vertex_slab(N, [0, 0, 0], [N, N, N], 0, dev)
def gpu(storage, precision):
    args, kw = T.sampler_args(cam); kw.pop("camera_data")
    s = S.VolumeRenderSampler(*args, tf, N_SAMPLES, storage=storage, precision=precision, **kw)
    st = L.STORE_F32 if storage == "f32" else L.STORE_F64
    v = vol if storage == "f32" else vol.double()
    bs = D.BrickSet(s._ctx, 1, 1, L.VERTEX_CENTRED, st)
    bs.attach(0, [v.data_ptr()], 0, [0.0] * 3, [1.0] * 3, [N] * 3, keep=v)
    bs.commit()
    s.cast_bricks(bs, np.arange(1, dtype=np.int64))
    return np.array(s.aimage)
fast = gpu("f32", "fast")
exact = gpu("f64", "exact")
fast64 = gpu("f64", "fast")
Sr, Gr = O.load_ref() if O.ref_available() else (O, O)
img = np.zeros((R, R, 4))
a2, kw2 = T.sampler_args(cam, image=img); kw2.pop("camera_data")
ref = Sr.VolumeRenderSampler(*a2, tf, N_SAMPLES, **kw2)
data = vol.cpu().numpy().astype("float64")
ref(Gr.PartitionedGrid(0, [data], np.ones((N, N, N), "uint8"), np.zeros(3), np.ones(3), np.array([N] * 3, "int64")), num_threads=0)
want = np.array(ref.aimage)
cmax = want[:, :, :3].sum(axis=2).max()
for name, got in (("fast", fast), ("exact", exact), ("fast/f64 storage", fast64)):
    d = np.abs(got[:, :, :3] - want[:, :, :3]).max(axis=2) / cmax
    bad = np.argwhere(d > 1e-4)
    print(name, "max", d.max(), "n_bad", len(bad), "of", R * R, "cmax", cmax)
    if len(bad):
        print("  bbox", bad.min(axis=0), bad.max(axis=0), "i%16 hist", np.bincount(bad[:, 0] % 16, minlength=16), "j%16 hist", np.bincount(bad[:, 1] % 16, minlength=16))
        order = np.argsort(-d[bad[:, 0], bad[:, 1]])[:12]
        for k in order:
            i, j = bad[k]
            print("  px", i, j, "err", d[i, j], "got", got[i, j, :3], "want", want[i, j, :3])
        np.save(os.path.join(ROOT, "gpurun_out", "diag_bad_%s.npy" % name.split("/")[0]), bad)
