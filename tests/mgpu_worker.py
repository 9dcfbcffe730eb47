"""Worker of tests/test_multigpu_gpu.py: one process per GPU (torchrun).  Renders the test scenes through
the public scene API with the kd-tree sharded over the ranks; rank 0 saves the images."""
import sys

import numpy as np

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import yt_b200 as yt  # noqa: E402
from tests.test_multigpu_gpu import build_scene, CASES  # noqa: E402


def main(out):
    rank, size = yt.enable_parallelism()
    images = {}
    for name in CASES:
        sc = build_scene(name)
        im = np.array(sc.render())
        images[name] = im
        src = sc.get_source()
        images[name + "/nbricks"] = np.array([len(src.volume.tree.traverse())])
    # every rank must hold the full image (scatter_image semantics)
    np.savez(out + ".rank%d.npz" % rank, **images)
    if size > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
