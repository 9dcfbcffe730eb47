"""The native kd-tree (csrc/kdtree.cpp) against the reference's own Node class
(yt/utilities/lib/amr_kdtools.pyx compiled unmodified into oracle/_ref) and against golden
dumps generated from it (tests/golden/kdtree_golden.npz, made by tests/golden/make_kdtree_golden.py)."""
import os
import sys

import numpy as np
import pytest

from oracle import oracle as O
from yt_b200 import loaders, synthetic
from yt_b200.amr_kdtree import AMRKDTree, Tree

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kdtree_golden.npz")


def cases():
    rng = np.random.default_rng(3)
    out = {}
    g, dims = synthetic.nested_amr_grids(base_cells=16, tile=8, levels=3, seed=1)
    out["nested3"] = loaders.load_amr_grids(g, dims)
    g, dims = synthetic.nested_amr_grids(base_cells=32, tile=16, levels=2, seed=2)
    out["nested2"] = loaders.load_amr_grids(g, dims)
    out["uniform1"] = loaders.load_uniform_grid({"density": rng.random((8, 8, 8))}, (8, 8, 8))
    out["uniform12"] = loaders.load_uniform_grid({"density": rng.random((12, 8, 10))}, (12, 8, 10), nprocs=12)
    # an off-centre refined patch on a 2-grid base
    base = [{"left_edge": [0, 0, 0], "right_edge": [0.5, 1, 1], "level": 0, "dimensions": [8, 16, 16], "density": rng.random((8, 16, 16))},
            {"left_edge": [0.5, 0, 0], "right_edge": [1, 1, 1], "level": 0, "dimensions": [8, 16, 16], "density": rng.random((8, 16, 16))},
            {"left_edge": [0.25, 0.125, 0.5], "right_edge": [0.5, 0.5, 0.75], "level": 1, "dimensions": [8, 12, 8], "density": rng.random((8, 12, 8))},
            {"left_edge": [0.5, 0.125, 0.5], "right_edge": [0.625, 0.375, 1.0], "level": 1, "dimensions": [4, 8, 16], "density": rng.random((4, 8, 16))}]
    out["patches"] = loaders.load_amr_grids(base, [16, 16, 16])
    return out


VIEWPOINTS = [None, (2.0, 2.0, 2.0), (-1.0, 0.4, 3.0), (0.5, 0.5, 0.5), (0.3, -5.0, 0.45), (0.26, 0.74, 0.51)]


def reference_dump(ds, rank=0, size=1):
    """Build the same tree with the reference's Node and dump it in depth-first order."""
    ref_dir = os.path.join(os.path.dirname(os.path.abspath(O.__file__)), "_ref")
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    from yt.utilities.lib.amr_kdtools import Node
    trunk = Node(None, None, None, np.array([-np.inf] * 3), np.array([np.inf] * 3), -1, 1)
    for lvl in range(ds.index.max_level + 1):
        grids = [g for g in ds.index.grids if g.Level == lvl]
        if not grids:
            continue
        trunk.add_grids(len(grids), np.array([g.LeftEdge for g in grids]), np.array([g.RightEdge for g in grids]),
                        np.array([g.id for g in grids], dtype="int64"), rank, size)
    nodes = {n.node_id: (n.grid, n.get_split_dim(), n.get_split_pos(), n.get_left_edge(), n.get_right_edge())
             for n in trunk.depth_traverse()}
    orders = []
    for vp in VIEWPOINTS:
        orders.append(np.array([n.node_id for n in trunk.kd_traverse(viewpoint=None if vp is None else np.array(vp))], dtype="int64"))
    return nodes, orders


def native_dump(ds, rank=0, size=1):
    t = Tree(ds, comm_rank=rank, comm_size=size)
    nd = t.nodes
    nodes = {int(nd["node_id"][i]): (int(nd["grid"][i]), int(nd["split_dim"][i]), float(nd["split_pos"][i]),
                                     nd["left_edge"][i], nd["right_edge"][i]) for i in range(len(nd["grid"]))}
    orders = [nd["node_id"][t.traverse(vp)] for vp in VIEWPOINTS]
    return nodes, orders


def _same(a, b):
    assert set(a) == set(b)
    for k in a:
        ga, da, pa, la, ra = a[k]
        gb, db, pb, lb, rb = b[k]
        assert ga == gb and da == db, (k, a[k], b[k])
        assert (np.isnan(pa) and np.isnan(pb)) or pa == pb
        assert np.array_equal(la, lb) and np.array_equal(ra, rb)


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name", sorted(cases()))
@pytest.mark.parametrize("rank,size", [(0, 1), (0, 2), (1, 2), (3, 4)])
def test_native_tree_equals_reference_node(name, rank, size):
    ds = cases()[name]
    rn, ro = reference_dump(ds, rank, size)
    nn, no = native_dump(ds, rank, size)
    _same(nn, rn)
    for a, b in zip(no, ro):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("name", sorted(cases()))
def test_native_tree_matches_golden(name):
    gold = np.load(GOLD)
    ds = cases()[name]
    nn, no = native_dump(ds)
    ids = gold[name + "/node_id"]
    assert sorted(nn) == sorted(ids.tolist())
    for i, k in enumerate(ids):
        assert nn[int(k)][0] == gold[name + "/grid"][i]
        assert nn[int(k)][1] == gold[name + "/split_dim"][i]
    for j, o in enumerate(no):
        assert np.array_equal(o, gold[name + "/order%d" % j])


def test_leaves_tile_the_domain_and_bricks_slice_the_grids():
    ds = cases()["nested3"]
    t = AMRKDTree(ds)
    t.set_fields([("stream", "density")], [False], True)
    vol = sum(float(np.prod(b.RightEdge - b.LeftEdge)) for b in t.bricks)
    assert abs(vol - 1.0) < 1e-12
    for b in t.bricks:
        g = ds.index.grids[b.parent_grid_id]
        assert np.allclose((b.RightEdge - b.LeftEdge) / b.dims, g.dds)
        assert b.my_data[0].shape == tuple(b.dims + 1)
    # every leaf of a coarse grid avoids the region covered by its children
    for b in t.bricks:
        g = ds.index.grids[b.parent_grid_id]
        for c in g.Children:
            overlap = np.minimum(b.RightEdge, c.RightEdge) - np.maximum(b.LeftEdge, c.LeftEdge)
            assert not np.all(overlap > 1e-12)


@pytest.mark.parametrize("size", [2, 4, 8])
def test_rank_subtrees_partition_the_leaves_and_paths_follow_rank_bits(size):
    """Each rank builds only its own subtree (should_i_build); together they hold every serial leaf, and the
    kd path handed to the binary swap is the chain of splits above heap node size + rank."""
    from yt_b200 import parallel
    ds = cases()["nested3"]
    serial = Tree(ds)
    sn = serial.nodes
    serial_leaves = sorted((tuple(sn["left_edge"][i]), tuple(sn["right_edge"][i]), int(sn["grid"][i])) for i in serial.traverse())
    mine = []
    k = int(np.log2(size))
    for rank in range(size):
        t = Tree(ds, comm_rank=rank, comm_size=size)
        nd = t.nodes
        mine += [(tuple(nd["left_edge"][i]), tuple(nd["right_edge"][i]), int(nd["grid"][i])) for i in t.traverse()]
        path = parallel.kd_path_of_rank(nd, rank, size)
        assert len(path) == k
        for level, (ax, pos, upper) in enumerate(path):
            assert upper == bool((rank >> (k - 1 - level)) & 1)
        # the subtree root's box is the intersection of the half-spaces along the path
        by_id = {int(i): n for n, i in enumerate(nd["node_id"])}
        root = by_id[size + rank]
        for ax, pos, upper in path:
            if upper:
                assert nd["left_edge"][root][ax] >= pos
            else:
                assert nd["right_edge"][root][ax] <= pos
        rounds, strip = parallel.plan_binary_swap(rank, size, 1000, path, np.array([2.0, -1.0, 0.3]))
        assert len(rounds) == k
    assert sorted(mine) == serial_leaves
