"""AMRKDTree -- brick decomposition of a grid dataset and its view-ordered traversal.

Mirror of yt/utilities/amr_kdtree/amr_kdtree.py (Tree 63-160, AMRKDTree 163-374).  The tree itself
(node splitting, ownership, traversal order) lives in libytb200.so (csrc/kdtree.cpp, a C++
restatement of yt/utilities/lib/amr_kdtools.pyx); this module slices grid data into
``PartitionedGrid`` bricks (get_brick_data, amr_kdtree.py:325-374) and -- what the reference does
not have -- keeps ALL leaves resident on the device as one brick table, so that a frame is
"ask the native tree for the leaf order, launch once" instead of a Python generator over bricks.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from .partitioned_grid import PartitionedGrid, build_brickset


#: the kernels follow a ray with 32-bit running offsets: a leaf with more stored elements than this is cut into
#: cell-aligned pieces (which reproduce the single-brick image when composited in increasing t -- SURVEY.md finding 6)
MAX_BRICK_ELEMS = 2 ** 31


def split_box(lo, hi, max_elems):
    """kd bisection (longest axis, cell boundary) of the vertex box lo..hi until every piece stores <= max_elems
    vertices.  Returns a tree: ("leaf", lo, hi) | ("split", axis, cut_index, lower, upper)."""
    lo, hi = [int(v) for v in lo], [int(v) for v in hi]
    ext = [hi[a] - lo[a] for a in range(3)]
    if float(np.prod([e + 1 for e in ext])) <= max_elems or max(ext) <= 1:
        return ("leaf", lo, hi)
    ax = int(np.argmax(ext))
    mid = lo[ax] + ext[ax] // 2
    lower_hi, upper_lo = list(hi), list(lo)
    lower_hi[ax] = mid
    upper_lo[ax] = mid
    return ("split", ax, mid, split_box(lo, lower_hi, max_elems), split_box(upper_lo, hi, max_elems))


def _tree_leaves(tree):
    if tree[0] == "leaf":
        return [tree]
    return _tree_leaves(tree[3]) + _tree_leaves(tree[4])


class Tree:
    """Native kd-tree handle; ``build()`` adds the grids level by level (amr_kdtree.py:110-118)."""

    def __init__(self, ds, comm_rank=0, comm_size=1, left=None, right=None, min_level=None, max_level=None,
                 data_source=None):
        self.ds = ds
        self._id_offset = getattr(ds.index.grids[0], "_id_offset", 0)
        self.data_source = ds.all_data() if data_source is None else data_source
        left = np.array([-np.inf] * 3) if left is None else np.asarray(left, dtype="float64")
        right = np.array([np.inf] * 3) if right is None else np.asarray(right, dtype="float64")
        self.min_level = 0 if min_level is None else min_level
        self.max_level = ds.index.max_level if max_level is None else max_level
        self.comm_rank, self.comm_size = comm_rank, comm_size
        lib = L.load()
        h = C.c_void_p()
        rc = lib.ytb200_kdtree_create(left.ctypes.data_as(L.dp), right.ctypes.data_as(L.dp), int(comm_rank), int(comm_size), C.byref(h))
        if rc:
            raise L.YTB200Error(lib.ytb200_kdtree_last_error().decode())
        self.handle = h
        self._nodes = None
        self.build()

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            L.load().ytb200_kdtree_destroy(h)

    def add_grids(self, grids):
        gles = np.ascontiguousarray([g.LeftEdge for g in grids], dtype="float64")
        gres = np.ascontiguousarray([g.RightEdge for g in grids], dtype="float64")
        gids = np.ascontiguousarray([g.id for g in grids], dtype="int64")
        rc = L.load().ytb200_kdtree_add_grids(self.handle, len(grids), gles.ctypes.data_as(L.dp), gres.ctypes.data_as(L.dp),
                                              gids.ctypes.data_as(L.i64p))
        if rc:
            raise L.YTB200Error(L.load().ytb200_kdtree_last_error().decode())
        self._nodes = None

    def build(self):
        for lvl in range(self.min_level, self.max_level + 1):
            grids = [b for b, mask in self.data_source.blocks if b.Level == lvl]
            if len(grids) == 0:
                continue
            self.add_grids(grids)

    @property
    def nodes(self):
        """Structured dump of every node (index == native node index)."""
        if self._nodes is None:
            lib = L.load()
            n = lib.ytb200_kdtree_node_count(self.handle)
            d = {"node_id": np.empty(n, "int64"), "grid": np.empty(n, "int64"), "left": np.empty(n, "int32"),
                 "right": np.empty(n, "int32"), "split_dim": np.empty(n, "int32"), "split_pos": np.empty(n, "float64"),
                 "left_edge": np.empty((n, 3), "float64"), "right_edge": np.empty((n, 3), "float64")}
            lib.ytb200_kdtree_nodes(self.handle, d["node_id"].ctypes.data_as(L.i64p), d["grid"].ctypes.data_as(L.i64p),
                                    d["left"].ctypes.data_as(L.i32p), d["right"].ctypes.data_as(L.i32p),
                                    d["split_dim"].ctypes.data_as(L.i32p), d["split_pos"].ctypes.data_as(L.dp),
                                    d["left_edge"].ctypes.data_as(L.dp), d["right_edge"].ctypes.data_as(L.dp))
            self._nodes = d
        return self._nodes

    def traverse(self, viewpoint=None):
        """Native node indices of the leaves carrying a grid, in Node.kd_traverse(viewpoint) order."""
        lib = L.load()
        n = lib.ytb200_kdtree_node_count(self.handle)
        out = np.empty(n, dtype="int32")
        vp = None
        if viewpoint is not None:
            vp = np.ascontiguousarray(viewpoint, dtype="float64")
        k = lib.ytb200_kdtree_traverse(self.handle, vp.ctypes.data_as(L.dp) if vp is not None else None,
                                       out.ctypes.data_as(L.i32p), n)
        if k < 0:
            raise L.YTB200Error(lib.ytb200_kdtree_last_error().decode())
        return out[:k]

    def sum_cells(self):
        nd = self.nodes
        cells = 0
        for i in self.traverse():
            grid = self.ds.index.grids[nd["grid"][i] - self._id_offset]
            li = np.rint((nd["left_edge"][i] - grid.LeftEdge) / grid.dds).astype("int64")
            ri = np.rint((nd["right_edge"][i] - grid.LeftEdge) / grid.dds).astype("int64")
            cells += int(np.prod(ri - li))
        return cells


class AMRKDTree:
    fields = None
    log_fields = None
    no_ghost = True

    def __init__(self, ds, min_level=None, max_level=None, data_source=None, comm_rank=None, comm_size=None):
        from . import parallel
        if comm_rank is None:
            comm_rank, comm_size = parallel.comm_rank(), parallel.comm_size()
        self.comm_rank, self.comm_size = comm_rank, comm_size
        if not hasattr(ds.index, "grids"):
            raise RuntimeError("AMRKDTree does not support particle or octree-based data.")
        self.ds = ds
        self.current_vcds = []
        self.current_saved_grids = []
        self._bricks = None          # host PartitionedGrids of every leaf, built on first use
        self.sdx = ds.index.get_smallest_dx()
        self._initialized = False
        self._id_offset = getattr(ds.index.grids[0], "_id_offset", 0)
        self.data_source = ds.all_data() if data_source is None else data_source
        self.tree = Tree(ds, comm_rank, comm_size, min_level=min_level, max_level=max_level, data_source=self.data_source)
        self._brick_of_node = {}     # native node index -> PartitionedGrid
        self._leaf_slot = None       # native node index -> position in the device brick table
        self._device_sets = {}       # (device, storage, centred on device) -> BrickSet over all leaves
        self._set_meta = {}          # id(BrickSet) -> (leaf slot table, piece plan) of that set
        self._piece_plan = {}

    # -- data ------------------------------------------------------------------------------------------
    def set_fields(self, fields, log_fields, no_ghost, force=False):
        new_fields = self.data_source._determine_fields(fields)
        regenerate = self.fields is None or len(self.fields) != len(new_fields) or self.fields != new_fields or force
        if not isinstance(log_fields, (list, tuple)):
            log_fields = [log_fields]
        new_log = list(log_fields)
        if not regenerate and (self.log_fields != new_log or self.no_ghost != no_ghost):
            regenerate = True  # the reference flips log in place (_apply_log); re-slicing gives the same bricks
        self.fields, self.log_fields, self.no_ghost = new_fields, new_log, no_ghost
        if regenerate:
            self._brick_of_node, self._device_sets, self._leaf_slot = {}, {}, None
            self._set_meta, self._piece_plan = {}, {}
            self.current_vcds, self.current_saved_grids = [], []
        # the reference builds every brick here (amr_kdtree.py:219-227).  The batched device path never reads the host
        # bricks (each grid's cells are vertex-centred on the device), so they are built on first use: a 1024^3
        # grid spent over a minute in the host's vertex centring before its first frame (profiles/r02h_bench_scene_before.json)
        self._bricks = None
        self._initialized = True

    @property
    def bricks(self):
        if self._bricks is None:
            self._bricks = np.array(list(self.traverse()), dtype=object)
        return self._bricks

    @property
    def brick_dimensions(self):
        return np.array([b.dims for b in self.bricks])

    def initialize_source(self, fields, log_fields, no_ghost):
        if fields == self.fields and log_fields == self.log_fields and no_ghost == self.no_ghost:
            return
        self.set_fields(fields, log_fields, no_ghost)

    def traverse(self, viewpoint=None):
        """Yield the bricks back to front w.r.t. ``viewpoint`` (amr_kdtree.py:248-250)."""
        for i in self.tree.traverse(viewpoint):
            yield self.get_brick_data(int(i))

    def get_brick_data(self, node):
        """amr_kdtree.py:325-374 for the leaf with native index ``node``."""
        if node in self._brick_of_node:
            return self._brick_of_node[node]
        nd = self.tree.nodes
        grid = self.ds.index.grids[nd["grid"][node] - self._id_offset]
        dds, gle = grid.dds, grid.LeftEdge
        nle, nre = nd["left_edge"][node].copy(), nd["right_edge"][node].copy()
        li = np.rint((nle - gle) / dds).astype("int32")
        ri = np.rint((nre - gle) / dds).astype("int32")
        dims = ri - li
        assert np.all(grid.LeftEdge <= nle) and np.all(grid.RightEdge >= nre)
        if grid.id in self.current_saved_grids:
            vcd = self.current_vcds[self.current_saved_grids.index(grid.id)]
        else:
            raw = grid.get_vertex_centered_data(self.fields, smoothed=True, no_ghost=self.no_ghost)
            vcd = []
            for i, field in enumerate(self.fields):
                v = raw[field].astype("float64")
                if self.log_fields[i]:
                    v[v <= 0] = np.nan
                    v = np.log10(v)
                vcd.append(v)
            self.current_saved_grids.append(grid.id)
            self.current_vcds.append(vcd)
        if self.data_source.selector is None:
            mask = np.ones(dims, dtype="uint8")
        else:
            mask, _ = self.data_source.selector.fill_mask_regular_grid(grid)
            mask = mask[li[0]:ri[0], li[1]:ri[1], li[2]:ri[2]].astype("uint8")
        data = [d[li[0]:ri[0] + 1, li[1]:ri[1] + 1, li[2]:ri[2] + 1].copy() for d in vcd]
        brick = PartitionedGrid(grid.id, data, mask, nle, nre, dims.astype("int64"))
        self._brick_of_node[node] = brick
        return brick

    # -- device residency (ours) ---------------------------------------------------------------------------
    def device_bricks(self, ctx, storage, centre_on_device=None):
        """One device brick table over every leaf (depth-first order); built once, reused every frame.

        With ``centre_on_device`` (default: whenever ``no_ghost``) the host never builds a vertex array:
        each grid's cell-centred fields go to the device once and every leaf is vertex-centred, log10'd,
        sliced and converted by ``ytb200_brick_from_cells`` (get_brick_data's arithmetic, amr_kdtree.py:343-365)."""
        if centre_on_device is None:
            centre_on_device = True
        key = (ctx.device, storage, bool(centre_on_device))
        if key not in self._device_sets:
            leaves = [int(i) for i in self.tree.traverse()]
            # slot table and piece plan belong to THIS set (another set of the same tree may cut its leaves differently)
            self._leaf_slot = np.full(len(self.tree.nodes["grid"]), -1, dtype="int64")
            self._leaf_slot[leaves] = np.arange(len(leaves))
            self._piece_plan = {}
            if centre_on_device and leaves:
                bs = self._device_bricks_from_cells(leaves, ctx, storage)
            else:
                bs = build_brickset([self.get_brick_data(i) for i in leaves], ctx, storage)
            self._device_sets[key] = bs
            self._set_meta[id(bs)] = (self._leaf_slot, self._piece_plan)
        bs = self._device_sets[key]
        self._leaf_slot, self._piece_plan = self._set_meta[id(bs)]   # view_order() without `bricks` means this set
        return bs

    def _device_bricks_from_cells(self, leaves, ctx, storage):
        from . import _lib as L
        from . import device as D
        ghost = not self.no_ghost

        def grid_cells(grid):
            """The grid's fields on the device: its cells (no ghost), or the cells padded with one layer of ghost zones
            (the smoothed covering grid of grid_patch.py:298, assembled on the host: it gathers from other grids)."""
            if ghost:
                cg = grid.retrieve_ghost_zones(1, self.fields, smoothed=True)
                return [D.DeviceCells(ctx, cg[f]) for f in self.fields]
            if np.any(np.asarray(grid.ActiveDimensions) < 2):
                raise NotImplementedError("no-ghost vertex centring needs >= 2 cells per grid axis")
            return [D.DeviceCells(ctx, grid[f]) for f in self.fields]
        nd = self.tree.nodes
        # leaves too large for 32-bit offsets become several device bricks (pieces of ONE reference brick)
        plans, n_slots = [], 0
        for node in leaves:
            grid = self.ds.index.grids[nd["grid"][node] - self._id_offset]
            li = np.rint((nd["left_edge"][node] - grid.LeftEdge) / grid.dds).astype("int64")
            ri = np.rint((nd["right_edge"][node] - grid.LeftEdge) / grid.dds).astype("int64")
            tree = split_box(li, ri, MAX_BRICK_ELEMS)
            pieces = _tree_leaves(tree)
            plans.append((node, grid, tree, pieces, n_slots))
            n_slots += len(pieces)
        self._piece_plan.clear()
        if n_slots != len(leaves):
            slot_of = {}
            for node, grid, tree, pieces, first in plans:
                for k, pc in enumerate(pieces):
                    slot_of[id(pc)] = first + k
                if len(pieces) > 1:
                    self._piece_plan[node] = (tree, grid, slot_of)
                self._leaf_slot[node] = first
        out = D.BrickSet(ctx, n_slots, len(self.fields), L.VERTEX_CENTRED, storage)
        cells_of_grid = {}
        for node, grid, tree, pieces, first in plans:
            if len(pieces) > 1:
                if grid.id not in cells_of_grid:
                    cells_of_grid[grid.id] = grid_cells(grid)
                for k, (_, plo, phi) in enumerate(pieces):
                    ple = grid.LeftEdge + np.array(plo) * grid.dds
                    pre = grid.LeftEdge + np.array(phi) * grid.dds
                    mask = None
                    if self.data_source.selector is not None:
                        mask, _ = self.data_source.selector.fill_mask_regular_grid(grid)
                        mask = mask[plo[0]:phi[0], plo[1]:phi[1], plo[2]:phi[2]].astype("uint8")
                    out.from_cells(first + k, [c.ptr for c in cells_of_grid[grid.id]], grid.ActiveDimensions, plo, phi, ple, pre,
                                   log_fields=self.log_fields, mask=mask, ghost=ghost)
        for node, grid, tree, pieces, slot in plans:
            if len(pieces) > 1:
                continue
            if grid.id not in cells_of_grid:
                cells_of_grid[grid.id] = grid_cells(grid)
            cells = cells_of_grid[grid.id]
            dds, gle = grid.dds, grid.LeftEdge
            nle, nre = nd["left_edge"][node].copy(), nd["right_edge"][node].copy()
            li = np.rint((nle - gle) / dds).astype("int64")   # amr_kdtree.py:333-335
            ri = np.rint((nre - gle) / dds).astype("int64")
            mask = None
            if self.data_source.selector is not None:
                mask, _ = self.data_source.selector.fill_mask_regular_grid(grid)
                mask = mask[li[0]:ri[0], li[1]:ri[1], li[2]:ri[2]].astype("uint8")
            out.from_cells(slot, [c.ptr for c in cells], grid.ActiveDimensions, li, ri, nle, nre,
                           log_fields=self.log_fields, mask=mask, ghost=ghost)
        ctx.synchronize()   # the cell arrays may be released once every leaf has been centred
        return out.commit()

    def view_order(self, viewpoint, camera=None, bricks=None):
        """Positions in the device brick table `bricks` (default: the set ``device_bricks`` returned last), in the
        order ``traverse(viewpoint)`` yields the leaves.  A leaf that
        was cut into pieces expands to its pieces in increasing ray parameter t (what walk_volume does inside ONE
        brick): along the ray direction for a plane-parallel lens, away from the common ray origin
        (the camera position) otherwise."""
        if self._leaf_slot is None:
            raise RuntimeError("device_bricks() must be called first")
        leaf_slot, plan = (self._leaf_slot, self._piece_plan) if bricks is None else self._set_meta[id(bricks)]
        leaves = self.tree.traverse(viewpoint)
        if not plan:
            return leaf_slot[leaves]
        if camera is not None and camera.lens.lens_type == "plane-parallel":
            direction = np.asarray(camera.unit_vectors[2], dtype="float64") * float(camera.width[2])
            lower_first = lambda grid, ax, cut: direction[ax] >= 0          # noqa: E731
        else:
            origin = np.asarray(camera.position if camera is not None else viewpoint, dtype="float64")
            if camera is None:   # only the kd viewpoint is known: its far side first, like the tree itself
                lower_first = lambda grid, ax, cut: not (origin[ax] <= grid.LeftEdge[ax] + cut * grid.dds[ax])   # noqa: E731
            else:
                lower_first = lambda grid, ax, cut: origin[ax] < grid.LeftEdge[ax] + cut * grid.dds[ax]          # noqa: E731
        out = []
        for node in leaves:
            node = int(node)
            if node not in plan:
                out.append(int(leaf_slot[node]))
                continue
            tree, grid, slot_of = plan[node]

            def walk(t):
                if t[0] == "leaf":
                    out.append(slot_of[id(t)])
                    return
                first, second = (t[3], t[4]) if lower_first(grid, t[1], t[2]) else (t[4], t[3])
                walk(first)
                walk(second)
            walk(tree)
        return np.asarray(out, dtype="int64")

    # -- parallel reduction -------------------------------------------------------------------------------
    def reduce_tree_images(self, image, viewpoint, *, use_opacity=True):
        """Single process: nothing to do.  With several ranks the compositing happens on the devices, inside
        ``ImageSampler.flush_parallel`` (binary swap of affine partial images), before the image reaches the
        host -- so by the time the reference would call this, `image` is already the reduced one."""
        return image

    def kd_path(self):
        """Splits from the root down to this rank's subtree (see parallel.kd_path_of_rank)."""
        from . import parallel
        return parallel.kd_path_of_rank(self.tree.nodes, self.comm_rank, self.comm_size)

    def locate_node(self, pos):
        nd = self.tree.nodes
        pos = np.asarray(pos, dtype="float64")
        n = 0
        assert np.all(nd["left_edge"][0] <= pos) and np.all(nd["right_edge"][0] > pos)
        while nd["left"][n] >= 0:
            n = nd["left"][n] if pos[nd["split_dim"][n]] < nd["split_pos"][n] else nd["right"][n]
        return n

    def count_cells(self):
        return self.tree.sum_cells()
