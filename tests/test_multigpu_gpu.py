"""Multi-GPU parity: the scene API with the kd-tree sharded over 2 (and 4, when available) GPUs must give
the SERIAL oracle image (SURVEY.md finding 8: never compare with the reference's MPI result).
Skipped on single-GPU boxes; run with `gpurun --gpus 2 -- python -m pytest tests/test_multigpu_gpu.py -m gpu`."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["pp_colour", "persp_grey", "amr_pp_grey"]


def build_scene(name):
    import yt_b200 as yt
    from yt_b200 import synthetic
    if name.startswith("amr"):
        g, dims = synthetic.nested_amr_grids(base_cells=16, tile=8, levels=3, seed=4)
        ds = yt.load_amr_grids(g, dims)
    else:
        arr = np.random.default_rng(11).random((32, 32, 32)) + 0.3
        ds = yt.load_uniform_grid({"density": arr}, arr.shape, nprocs=8)
    lens = "perspective" if "persp" in name else "plane-parallel"
    sc = yt.create_scene(ds, lens_type=lens)
    sc.camera.resolution = (96, 80)
    if lens == "perspective":
        sc.camera.position = [1.7, 1.5, 1.3]
        sc.camera.width = [0.7, 0.7, 1.0]
    src = sc.get_source()
    src.tfh.grey_opacity = "grey" in name
    return sc


def _device_count():
    try:
        from yt_b200 import _lib
        return _lib.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("nranks", [2, 4])
def test_sharded_scene_matches_serial_oracle(nranks, tmp_path, gpu):
    if _device_count() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    from tests import common as T
    from tests.test_scene_gpu import oracle_render
    out = str(tmp_path / "mgpu")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks),
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + nranks), os.path.join(ROOT, "tests", "mgpu_worker.py"), out]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    got = [np.load(out + ".rank%d.npz" % k) for k in range(nranks)]
    for name in CASES:
        sc = build_scene(name)           # serial: comm_size == 1 in this process
        want, nb = oracle_render(sc)
        grey = "grey" in name
        assert sum(int(g[name + "/nbricks"][0]) for g in got) == nb   # the ranks partition the leaves
        assert all(int(g[name + "/nbricks"][0]) > 0 for g in got)
        for k in range(nranks):
            err = T.normalised_diff(got[k][name], want, grey)
            assert err <= 1e-4, (name, k, err)
