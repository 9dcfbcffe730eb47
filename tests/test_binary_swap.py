"""Host-side logic of the multi-GPU path, on CPU: kd partition, binary-swap schedule, ordered
composition.  The world_size-2/4 runs use torch.distributed with the gloo backend."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from yt_b200.parallel import bit_reverse, plan_binary_swap, strip_bounds  # noqa: E402


def _paths(world, n=64):
    from bench import kd_slab
    return [kd_slab(n, r, world) for r in range(world)]


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_kd_slabs_tile_the_volume(world):
    n = 64
    cover = np.zeros((n, n, n), "int32")
    for lo, hi, path in _paths(world, n):
        cover[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] += 1
        assert len(path) == int(np.log2(world))
    assert (cover == 1).all()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_schedule_is_consistent(world):
    npx = 1000
    vp = np.array([3.0, -2.0, 0.4])
    plans = [plan_binary_swap(r, world, npx, p[2], vp) for r, p in enumerate(_paths(world))]
    strips = strip_bounds(world, npx)
    seen = np.zeros(npx, "int32")
    for r, (rounds, strip) in enumerate(plans):
        assert strip == strips[r]
        seen[strip[0]:strip[1]] += 1
        for j, rd in enumerate(rounds):
            other = plans[rd["partner"]][0][j]
            assert other["partner"] == r
            assert other["keep"] == rd["send"] and other["send"] == rd["keep"]
            assert other["earlier"] != rd["earlier"]
            assert (other["axis"], other["pos"]) == (rd["axis"], rd["pos"])
    assert (seen == 1).all()
    assert bit_reverse(1, 3) == 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, npx, seed, out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from bench import kd_slab
    from yt_b200.parallel import BinarySwap
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(seed + rank)
    partial = torch.from_numpy(np.concatenate([rng.uniform(0, 1, (npx, 4)), rng.uniform(0.2, 1, (npx, 4))], axis=1))
    image = torch.from_numpy(np.random.RandomState(seed - 1).uniform(0, 1, (npx, 4)))
    _, _, path = kd_slab(64, rank, world)
    vp = np.array([3.0, -2.0, 0.4])

    def compose(e, l, o):
        A = l[:, :4] + l[:, 4:] * e[:, :4]
        B = l[:, 4:] * e[:, 4:]
        o[:, :4] = A
        o[:, 4:] = B

    def apply(p, im):
        im.copy_(p[:, :4] + p[:, 4:] * im)

    sw = BinarySwap(None, npx, 1, path, vp, rank, world, compose=compose, apply=apply)
    sw.composite_tensors(partial.clone(), image)
    sw.gather_strips(image, 0)
    if rank == 0:
        np.save(out, image.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_binary_swap_over_gloo_matches_serial_composition(world, tmp_path):
    import torch.multiprocessing as mp
    from bench import kd_slab
    npx, seed = 1237, 11
    out = str(tmp_path / "img.npy")
    mp.spawn(_worker, args=(world, _free_port(), npx, seed, out), nprocs=world, join=True)
    got = np.load(out)
    # serial reference: apply the ranks' maps in the kd traversal order for this viewpoint
    vp = np.array([3.0, -2.0, 0.4])
    parts = []
    for r in range(world):
        rng = np.random.RandomState(seed + r)
        parts.append(np.concatenate([rng.uniform(0, 1, (npx, 4)), rng.uniform(0.2, 1, (npx, 4))], axis=1))

    def order(ranks, depth):
        if len(ranks) == 1:
            return ranks
        _, _, path = kd_slab(64, ranks[0], world)
        ax, pos, _ = path[depth]
        half = len(ranks) // 2
        lower, upper = ranks[:half], ranks[half:]
        first, second = (upper, lower) if vp[ax] <= pos else (lower, upper)
        return order(first, depth + 1) + order(second, depth + 1)

    img = np.random.RandomState(seed - 1).uniform(0, 1, (npx, 4))
    for r in order(list(range(world)), 0):
        img = parts[r][:, :4] + parts[r][:, 4:] * img
    assert np.allclose(got, img, rtol=1e-13, atol=1e-13)
