"""In-memory ("stream") grid datasets: ``load_uniform_grid`` / ``load_amr_grids``.

Only what feeds the ray caster is restated from the reference: the grid hierarchy
(yt/loaders.py:204-448 uniform, 450-688 AMR; yt/frontends/stream), the vertex centring that
turns cell data into brick data -- no-ghost (grid_patch.py:270-296 -> interpolators.pyx:177-227)
and with ghost zones (grid_patch.py:239-268, 298-310 on top of the smoothed covering grid,
construction_data_containers.py:1398-1653 + misc_utilities.pyx:707-789) -- and the child mask
used by the off-axis projection loop
(off_axis_projection.py:506-515).  Everything is in ``code_length``; there is no unit system.
"""
import numpy as np

#: stream fields the reference renders in log space by default (their FieldInfo.take_log)
LOG_FIELDS = {"density", "dark_matter_density", "temperature", "pressure", "number_density",
              "thermal_energy", "metal_density", "entropy"}


class FieldInfo:
    def __init__(self, name, take_log, units=""):
        self.name, self.take_log, self.units = name, bool(take_log), units

    def get_label(self):
        return self.name[1]


def ghost_zone_interpolate(rf, input_field, input_left, out_shape, output_left):
    """interpolators.pyx:177-227 on the host, term for term: every output point takes the 8-term weighted sum of the
    input cells (ii, ij, ik)..(ii+1, ij+1, ik+1), ii = clip(int(pos), 0, n - 2), in the reference's term order and
    with its weight products ((xm*ym)*zm ...), so the result is bit-identical to the Cython loop and to the device
    kernel (``vertex_centre_kernel``).  Output positions accumulate ``opos += ods`` exactly as the reference does."""
    f = np.ascontiguousarray(input_field, dtype="float64")
    input_left = np.asarray(input_left, dtype="float64")
    output_left = np.asarray(output_left, dtype="float64")
    idx, wp = [], []
    for a in range(3):
        n_in, n_out = f.shape[a], int(out_shape[a])
        if n_in < 2:
            raise NotImplementedError("vertex centring needs >= 2 cells per axis (the reference divides by shape - 1)")
        temp = input_left[a] + (rf * (n_in - 1))
        ids = (temp - input_left[a]) / (n_in - 1)
        iids = 1.0 / ids
        if n_out > 1:
            temp = output_left[a] + n_out - 1
            ods = (temp - output_left[a]) / (n_out - 1)
        else:
            ods = 0.0
        opos = np.empty(n_out, dtype="float64")
        acc = np.float64(output_left[a])
        for o in range(n_out):        # sequential accumulation, like the reference's `opos += ods`
            opos[o] = acc
            acc = acc + ods
        ropos = (opos - input_left[a]) * iids
        ii = np.clip(np.trunc(ropos).astype("int64"), 0, n_in - 2)   # C <int> truncates toward zero
        idx.append(ii)
        wp.append(ropos - ii)
    (ii, ij, ik), (xp, yp, zp) = idx, wp
    xm, ym, zm = 1.0 - xp, 1.0 - yp, 1.0 - zp
    X = lambda v: v[:, None, None]   # noqa: E731
    Y = lambda v: v[None, :, None]   # noqa: E731
    Z = lambda v: v[None, None, :]   # noqa: E731
    g = lambda a, b, c: f[np.ix_(ii + a, ij + b, ik + c)]   # noqa: E731
    out = g(0, 0, 0) * ((X(xm) * Y(ym)) * Z(zm))
    out = out + g(1, 0, 0) * ((X(xp) * Y(ym)) * Z(zm))
    out = out + g(0, 1, 0) * ((X(xm) * Y(yp)) * Z(zm))
    out = out + g(0, 0, 1) * ((X(xm) * Y(ym)) * Z(zp))
    out = out + g(1, 0, 1) * ((X(xp) * Y(ym)) * Z(zp))
    out = out + g(0, 1, 1) * ((X(xm) * Y(yp)) * Z(zp))
    out = out + g(1, 1, 0) * ((X(xp) * Y(yp)) * Z(zm))
    out = out + g(1, 1, 1) * ((X(xp) * Y(yp)) * Z(zp))
    return np.ascontiguousarray(out)


def vertex_centre_no_ghost(cells):
    """grid_patch.py:284-296: ghost_zone_interpolate(1, cells, [0.5]*3, out, [0]*3) -- trilinear interpolation of
    the cell centres onto the vertices, linearly extrapolated at the faces.  Needs >= 2 cells per axis (a one-cell
    axis is 0/0 in the reference), like the device path."""
    c = np.asarray(cells, dtype="float64")
    return ghost_zone_interpolate(1, c, [0.5] * 3, [n + 1 for n in c.shape], [0.0] * 3)


def vertex_centre_from_ghost(src):
    """grid_patch.py:298-310: with `src` the cells padded by one ghost layer, each vertex is 0.125 times the sum of its
    8 surrounding cells, added in this order (fp64, so the order is part of the result)."""
    src = np.asarray(src, dtype="float64")
    dest = np.zeros(tuple(n - 1 for n in src.shape), dtype="float64")
    np.add(dest, src[1:, 1:, 1:], dest)
    np.add(dest, src[:-1, 1:, 1:], dest)
    np.add(dest, src[1:, :-1, 1:], dest)
    np.add(dest, src[1:, 1:, :-1], dest)
    np.add(dest, src[:-1, 1:, :-1], dest)
    np.add(dest, src[1:, :-1, :-1], dest)
    np.add(dest, src[:-1, :-1, 1:], dest)
    np.add(dest, src[:-1, :-1, :-1], dest)
    np.multiply(dest, 0.125, dest)
    return dest


def fill_region(ds, key, level, left_index, dims):
    """What misc_utilities.pyx:710-789 leaves in a zeroed (dims) box whose first cell has index `left_index` at `level`,
    fed every cell of the level-`level` grids (a data source with min_level = max_level = level selects a grid's
    cells whether or not finer grids cover them): each cell lands at its own index, and at its periodic images when
    the box pokes out of the domain (the reference wraps whenever the box does; a non-periodic axis simply has no
    cells selected beyond the domain, so its ghost cells stay zero)."""
    out = np.zeros(tuple(int(d) for d in dims), dtype="float64")
    rf = int(ds.refine_by) ** int(level)
    level_dims = ds.domain_dimensions.astype("int64") * rf
    left_index = np.asarray(left_index, dtype="int64")
    dims = np.asarray(dims, dtype="int64")
    shifts = []
    for a in range(3):
        sh = [0]
        if ds.periodicity[a]:
            if left_index[a] < 0:
                sh.append(-1)          # cells near the upper domain face also appear below the lower one
            if left_index[a] + dims[a] >= level_dims[a]:
                sh.append(1)
        shifts.append(sh)
    for g in ds.index.grids:
        if g.Level != level:
            continue
        gs = g.get_global_startindex()
        ge = gs + g.ActiveDimensions
        v = g.fields[key]
        for sx in shifts[0]:
            for sy in shifts[1]:
                for sz in shifts[2]:
                    off = np.array([sx, sy, sz], dtype="int64") * level_dims
                    lo = np.maximum(gs + off, left_index)
                    hi = np.minimum(ge + off, left_index + dims)
                    if np.any(hi <= lo):
                        continue
                    a0, a1 = lo - left_index, hi - left_index
                    b0, b1 = lo - off - gs, hi - off - gs
                    out[a0[0]:a1[0], a0[1]:a1[1], a0[2]:a1[2]] = v[b0[0]:b1[0], b0[1]:b1[1], b0[2]:b1[2]]
    return out


def _finest_uncovered_level(ds, level, lo, hi):
    """Coarsest level among the cells a region of level-`level` cells [lo, hi) selects when levels 0..level are allowed:
    cells of coarser grids count only where no finer grid covers them (construction_data_containers.py:1460-1492)."""
    out = level
    for g in ds.index.grids:
        if g.Level >= out:
            continue
        r = int(ds.refine_by) ** (level - g.Level)
        gs = g.get_global_startindex()
        ldims = ds.domain_dimensions.astype("int64") * int(ds.refine_by) ** level
        for sx in (-1, 0, 1):
            for sy in (-1, 0, 1):
                for sz in (-1, 0, 1):
                    sh = np.array([sx, sy, sz], dtype="int64")
                    if any(sh[a] != 0 and not ds.periodicity[a] for a in range(3)):
                        continue
                    off = sh * ldims
                    a0 = np.maximum(np.floor_divide(lo - off, r), gs) - gs
                    a1 = np.minimum(-np.floor_divide(-(hi - off), r), gs + g.ActiveDimensions) - gs
                    if np.any(a1 <= a0):
                        continue
                    if g.child_mask[a0[0]:a1[0], a0[1]:a1[1], a0[2]:a1[2]].any():
                        out = min(out, g.Level)
    return out


def smoothed_covering_grid(ds, level, start_index, dims, key):
    """YTSmoothedCoveringGrid (construction_data_containers.py:1398-1653) for one field: the box of `dims` cells at
    `level` starting at global index `start_index`, filled level by level from the coarsest one that matters -- the
    previous level's box is interpolated onto this level's (``ghost_zone_interpolate``), then every cell a grid of
    this level covers with a NON-ZERO value replaces the interpolated one (the reference tests ``output != 0``).
    At level > 0 every box carries one buffer cell per side, stripped at the end."""
    start_index = np.asarray(start_index, dtype="int64")
    dims = np.asarray(dims, dtype="int64")
    rb = int(ds.refine_by)
    base_dx = ds.domain_width / ds.domain_dimensions.astype("float64")
    dle = ds.domain_left_edge
    my_dx = base_dx / rb ** int(level)
    left_edge = start_index * my_dx + dle
    right_edge = left_edge + dims * my_dx
    eps = np.finfo(np.float64).eps

    def minimal_box(dx):                     # :1600-1618
        cell_start = ((left_edge - dle) + eps) / dx
        cell_end = ((right_edge - dle) + eps) / dx
        if level == 0:
            start = np.floor(cell_start).astype("int64")
            d = np.rint((dims * my_dx) / dx).astype("int64")
        else:
            start = np.rint(cell_start).astype("int64") - 1
            d = np.rint(cell_end).astype("int64") - start + 1
        return start, d

    # the coarsest level whose cells can survive in the box (:1460-1492)
    min_level = 0
    for lvl in range(int(level), 0, -1):
        dx = base_dx / rb ** lvl
        st, d = minimal_box(dx)
        min_level = _finest_uncovered_level(ds, lvl, st - 1, st + d + 1)
        if min_level >= lvl:
            break
    cur_start, cur_dims = minimal_box(base_dx)
    field = np.zeros(tuple(int(x) for x in cur_dims), dtype="float64")
    for lvl in range(int(level) + 1):
        if lvl >= min_level:
            out = fill_region(ds, key, lvl, cur_start, cur_dims)
            m = out != 0
            field[m] = out[m]
        if lvl < level:                      # _update_level_state, :1620-1653
            rf = float(rb)
            old_start = cur_start
            cur_start, cur_dims = minimal_box(base_dx / rb ** (lvl + 1))
            field = ghost_zone_interpolate(rf, field, old_start * rf + 1, cur_dims, cur_start + 0.5)
    if level > 0:
        field = field[1:-1, 1:-1, 1:-1]
    assert field.shape == tuple(int(x) for x in dims), (field.shape, dims)
    return np.ascontiguousarray(field)


class Grid:
    """One grid patch (AMRGridPatch): cell-centred fields on ``ActiveDimensions`` cells."""
    _id_offset = 0

    def __init__(self, gid, level, left_edge, right_edge, dims, fields, ds):
        self.id = int(gid)
        self.Level = int(level)
        self.LeftEdge = np.array(left_edge, dtype="float64")
        self.RightEdge = np.array(right_edge, dtype="float64")
        self.ActiveDimensions = np.array(dims, dtype="int32")
        self.dds = (self.RightEdge - self.LeftEdge) / self.ActiveDimensions
        self.fields = fields
        self.ds = ds
        self.Children = []
        self.Parent = None
        self._child_mask = None

    def __getitem__(self, field):
        return self.fields[self.ds._resolve(field)]

    def clear_data(self):
        pass

    def get_global_startindex(self):
        return np.rint((self.LeftEdge - self.ds.domain_left_edge) / self.dds).astype("int64")

    @property
    def child_mask(self):
        """1 where a cell is not covered by a finer grid (grid_patch.py child_mask)."""
        if self._child_mask is None:
            m = np.ones(tuple(self.ActiveDimensions), dtype="uint8")
            for c in self.Children:
                lo = np.clip(np.rint((c.LeftEdge - self.LeftEdge) / self.dds).astype("int64"), 0, self.ActiveDimensions)
                hi = np.clip(np.rint((c.RightEdge - self.LeftEdge) / self.dds).astype("int64"), 0, self.ActiveDimensions)
                m[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = 0
            self._child_mask = m
        return self._child_mask

    def retrieve_ghost_zones(self, n_zones, fields, all_levels=False, smoothed=False):
        """grid_patch.py:239-268: this grid's cells plus ``n_zones`` layers around it, at the grid's own level, as
        {field: array of shape ActiveDimensions + 2 n_zones}.  Only the smoothed form (the one vertex centring asks
        for) is restated: ``smoothed_covering_grid``."""
        if all_levels or not smoothed:
            raise NotImplementedError("only retrieve_ghost_zones(n, fields, smoothed=True) is on the ray-casting path")
        start = self.get_global_startindex() - int(n_zones)
        dims = self.ActiveDimensions.astype("int64") + 2 * int(n_zones)
        out = {}
        for f in set(fields):
            key = self.ds._resolve(f)
            if key not in out:
                out[key] = smoothed_covering_grid(self.ds, self.Level, start, dims, key)
            out[f] = out[key]          # the cube answers to the name it was asked for and to the resolved one
        return out

    def get_vertex_centered_data(self, fields, smoothed=True, no_ghost=False):
        """grid_patch.py:270-312.  ``no_ghost``: trilinear extrapolation from the grid's own cells.  Otherwise every
        vertex is the mean of the 8 cells around it, taken from the grid padded with one layer of ghost zones
        (neighbouring grids of the same level, periodic images, or values interpolated from the coarser levels),
        summed in the reference's order."""
        out = {}
        cg = None if no_ghost else self.retrieve_ghost_zones(1, fields, smoothed=smoothed)
        for f in set(fields):
            key = self.ds._resolve(f)
            if key not in out:
                out[key] = vertex_centre_no_ghost(self[f]) if no_ghost else vertex_centre_from_ghost(cg[key])
            out[f] = out[key]
        return out

    def __repr__(self):
        return "Grid_%04d (%s)" % (self.id, tuple(self.ActiveDimensions))


class GridIndex:
    def __init__(self, grids):
        self.grids = np.empty(len(grids), dtype=object)
        for i, g in enumerate(grids):
            self.grids[i] = g
        self.max_level = max(g.Level for g in grids)
        self.num_grids = len(grids)

    def get_smallest_dx(self):
        return min(g.dds.min() for g in self.grids)

    def select_grids(self, level):
        return [g for g in self.grids if g.Level == level]


class _Quantities(dict):
    def __init__(self, src):
        super().__init__()
        self["Extrema"] = src._extrema


class AllData:
    """``ds.all_data()``: the whole domain as a data source (no selector: every cell is in)."""
    selector = None

    def __init__(self, ds):
        self.ds = self.pf = ds
        self.quantities = _Quantities(self)

    @property
    def blocks(self):
        for g in self.ds.index.grids:
            yield g, g.child_mask.astype(bool)

    def _determine_fields(self, fields):
        if isinstance(fields, (str, tuple)):
            fields = [fields]
        return [self.ds._resolve(f) for f in fields]

    def _extrema(self, field, non_zero=False):
        key = self.ds._resolve(field)
        lo, hi = np.inf, -np.inf
        for g, mask in self.blocks:
            v = g.fields[key]
            if not mask.all():
                v = v[mask]
            if non_zero and v.size and not (v.min() > 0):   # (one pass when everything is positive, the usual case)
                v = v[v > 0]
            if v.size:
                lo, hi = min(lo, v.min()), max(hi, v.max())
        return np.float64(lo), np.float64(hi)

    def all_data(self):
        return self

    def __repr__(self):
        return "YTRegion (%s)" % self.ds


class StreamDataset:
    """The slice of ``yt.frontends.stream.StreamDataset`` the volume renderer touches."""

    def __init__(self, grids_spec, domain_dimensions, bbox, refine_by=2, periodicity=(True, True, True), name="stream"):
        self.domain_dimensions = np.array(domain_dimensions, dtype="int64")
        bbox = np.array(bbox, dtype="float64")
        self.domain_left_edge = bbox[:, 0].copy()
        self.domain_right_edge = bbox[:, 1].copy()
        self.domain_width = self.domain_right_edge - self.domain_left_edge
        self.domain_center = 0.5 * (self.domain_left_edge + self.domain_right_edge)
        self.refine_by = refine_by
        self.periodicity = tuple(periodicity)
        self.basename = name
        self.field_info = {}
        self.field_list = []
        grids = []
        for gid, spec in enumerate(grids_spec):
            fields = {}
            for k, v in spec["fields"].items():
                key = k if isinstance(k, tuple) else ("stream", k)
                arr = np.ascontiguousarray(v, dtype="float64")
                if arr.shape != tuple(int(d) for d in spec["dimensions"]):
                    raise RuntimeError("field %s has shape %s, grid dimensions are %s" % (k, arr.shape, tuple(spec["dimensions"])))
                fields[key] = arr
                if key not in self.field_info:
                    self.field_info[key] = FieldInfo(key, key[1] in LOG_FIELDS)
                    self.field_list.append(key)
            grids.append(Grid(gid, spec["level"], spec["left_edge"], spec["right_edge"], spec["dimensions"], fields, self))
        self.index = GridIndex(grids)
        self._link_children()
        self.derived_field_list = list(self.field_list)
        self.default_field = self._pick_default()
        self.unit_registry = None

    # -- field names: "density", ("gas", "density") and ("stream", "density") all name the same array -------
    def _resolve(self, field):
        if isinstance(field, tuple) and field in self.field_info:
            return field
        name = field[1] if isinstance(field, tuple) else field
        hits = [k for k in self.field_list if k[1] == name]
        if len(hits) != 1:
            raise KeyError("Could not find field %r in %s." % (field, self))
        return hits[0]

    def _get_field_info(self, field):
        return self.field_info[self._resolve(field)]

    def _pick_default(self):
        for k in self.field_list:
            if k[1] == "density":
                return k
        return self.field_list[0]

    def _link_children(self):
        """Parent/child links: a grid one level finer whose box lies inside the coarser one."""
        by_level = {}
        for g in self.index.grids:
            by_level.setdefault(g.Level, []).append(g)
        eps = 1e-12 * self.domain_width
        for lvl, gs in by_level.items():
            kids = by_level.get(lvl + 1, [])
            if not kids:
                continue
            # one vectorised containment test per parent (a 16 k-grid hierarchy links in well under a second)
            kle = np.array([c.LeftEdge for c in kids])
            kre = np.array([c.RightEdge for c in kids])
            for g in gs:
                inside = np.all(kle >= g.LeftEdge - eps, axis=1) & np.all(kre <= g.RightEdge + eps, axis=1)
                for k in np.nonzero(inside)[0]:
                    g.Children.append(kids[k])
                    kids[k].Parent = g

    def all_data(self):
        return AllData(self)

    def __repr__(self):
        return self.basename


def _prime_powers(n):
    """[(prime, power)] of n, primes ascending."""
    out, p = [], 2
    while n > 1:
        m = 0
        while n % p == 0:
            n //= p
            m += 1
        if m:
            out.append((p, m))
        p += 1 if p == 2 else 2
    return out


def _decomposition_quality(n_d, pieces, ldom):
    """Score of the processor grid `ldom` for an array of shape `n_d` (decompose.py:27-50: load balance over a
    surface penalty, with the reference's constants; 0 when an axis would get more pieces than cells)."""
    eff = int((n_d > 1).sum())
    ideal = eff * pieces ** (1.0 / eff) * np.prod(n_d) ** (float(eff - 1) / float(eff))
    m = np.where(n_d > 1)
    nd = np.array(n_d, dtype=np.float64)[m]
    bsize = int(np.sum(ldom[m] / nd * np.prod(nd)))
    balance = float(np.prod(n_d)) / (float(pieces) * np.prod((n_d - 1) // ldom + 1))
    q = balance / (1 + 0.25 * (bsize / ideal - 1.0))
    q *= 1.0 - (0.001 * ldom[0] + 0.0001 * ldom[1]) / pieces
    return 0 if np.any(ldom > n_d) else q


def get_psize(n_d, pieces):
    """Processor grid (px, py, pz) with px*py*pz == pieces that the reference picks for an array of shape `n_d`
    (decompose.py:69-105).  Every way of dealing each prime power p^m over the three axes is scored and the FIRST
    best one wins, so the candidates must come in the reference's order: per prime the exponent triples
    (m - t, t - k, k) for t = m..0, k = t..0, the smallest prime's triple changing fastest.  The cut positions
    decide the no-ghost vertex values on internal grid faces, hence the image (each grid extrapolates at its own
    faces), so this has to be the reference's choice, not just any cell-aligned one."""
    n_d = np.asarray(n_d, dtype=np.int64)
    pieces = int(pieces)
    if pieces == 1:
        return np.ones(3, dtype=np.int64)
    per_prime = []
    for p, m in _prime_powers(pieces):
        per_prime.append([(p, (m - t, t - k, k)) for t in range(m, -1, -1) for k in range(t, -1, -1)])
    best, psize = 0.0, np.ones(3, dtype=np.int64)
    pos = [0] * len(per_prime)
    while True:
        ldom = np.ones(3, dtype=np.int64)
        for lst, i in zip(per_prime, pos):
            p, e = lst[i]
            ldom *= p ** np.array(e)
        q = _decomposition_quality(n_d, pieces, ldom)
        if q > best:
            best, psize = q, ldom
        # odometer, first prime fastest; the reference stops when the LAST prime runs out of triples
        j = 0
        while j < len(pos):
            if pos[j] + 1 < len(per_prime[j]):
                pos[j] += 1
                break
            pos[j] = 0
            j += 1
        if j == len(pos):
            return psize


def piece_edges(n_d, psize):
    """Cell ranges [(lo, hi)] of the px*py*pz pieces in the reference's order (decompose.py:108-148, x slowest):
    lo = n * i // p, hi = n * (i + 1) // p."""
    n_d = np.asarray(n_d, dtype=np.int64)
    psize = np.asarray(psize, dtype=np.int64)
    out = []
    for i in range(int(psize[0])):
        for j in range(int(psize[1])):
            for k in range(int(psize[2])):
                pc = np.array((i, j, k), dtype=np.int64)
                out.append((n_d * pc // psize, n_d * (pc + 1) // psize))
    return out


def load_uniform_grid(data, domain_dimensions, length_unit=None, bbox=None, nprocs=1, sim_time=0.0,
                      periodicity=(True, True, True), geometry="cartesian", **kwargs):
    """One level-0 grid (or ``nprocs`` cell-aligned pieces of it) holding ``data`` (loaders.py:204-448)."""
    if geometry != "cartesian":
        raise NotImplementedError("only cartesian geometry reaches the ray caster")
    dims = np.array(domain_dimensions, dtype="int64")
    if bbox is None:
        bbox = np.array([[0.0, 1.0], [0.0, 1.0], [0.0, 1.0]])
    bbox = np.array(bbox, dtype="float64")
    le, re = bbox[:, 0], bbox[:, 1]
    fields = {}
    for k, v in data.items():
        if isinstance(v, tuple):  # (array, units)
            v = v[0]
        fields[k] = np.asarray(v)
    specs = []
    if nprocs == 1:
        specs.append({"level": 0, "left_edge": le, "right_edge": re, "dimensions": dims, "fields": fields})
    else:
        # loaders.py:355-372: the processor grid comes from the FIRST field's array shape
        psize = get_psize(dims, nprocs)
        dds = (re - le) / dims
        for lo, hi in piece_edges(dims, psize):
            sub = {f: v[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] for f, v in fields.items()}
            specs.append({"level": 0, "left_edge": le + lo * dds, "right_edge": le + hi * dds,
                          "dimensions": hi - lo, "fields": sub})
    return StreamDataset(specs, dims, bbox, periodicity=periodicity, name="UniformGridData")


def load_amr_grids(grid_data, domain_dimensions, bbox=None, sim_time=0.0, length_unit=None, periodicity=(True, True, True),
                   geometry="cartesian", refine_by=2, **kwargs):
    """A hierarchy of grid patches given as dicts with ``left_edge``, ``right_edge``, ``level``,
    ``dimensions`` and one array per field (loaders.py:450-688)."""
    if geometry != "cartesian":
        raise NotImplementedError("only cartesian geometry reaches the ray caster")
    dims = np.array(domain_dimensions, dtype="int64")
    if bbox is None:
        bbox = np.array([[0.0, 1.0], [0.0, 1.0], [0.0, 1.0]])
    specs = []
    for g in grid_data:
        g = dict(g)
        spec = {"left_edge": g.pop("left_edge"), "right_edge": g.pop("right_edge"), "level": g.pop("level"),
                "dimensions": g.pop("dimensions")}
        g.pop("number_of_particles", None)
        spec["fields"] = {k: (v[0] if isinstance(v, tuple) else v) for k, v in g.items()}
        specs.append(spec)
    return StreamDataset(specs, dims, bbox, refine_by=refine_by, periodicity=periodicity, name="AMRGridData")


# ---------------------------------------------------------------------------------------------------
# octree cells (OctreeVolumeSource, render_source.py:619-693)
# ---------------------------------------------------------------------------------------------------
class OctreeIndex:
    """Marker for ``create_volume_source`` (render_source.py:184-185): datasets whose index is an octree."""


class OctreeAllData(AllData):
    """The cells of an octree dataset as a data source: ``src["x"|"y"|"z"|"dx"|"grid_level"]`` in unitary
    coordinates, the way OctreeRayTracing / OctreeVolumeSource.render read them (octree_raytracing.py:33-34,
    render_source.py:669-672)."""

    @property
    def blocks(self):
        return iter(())

    def __getitem__(self, key):
        return self.ds.cell_arrays[key[1] if isinstance(key, tuple) else key]

    def _extrema(self, field, non_zero=False):
        v = self.ds.vertex_data[self.ds._resolve(field)]
        if non_zero:
            v = v[v > 0]
        return np.float64(v.min()), np.float64(v.max())

    def vertex_centered_data(self, field):
        """(2, 2, 2, n_cells) vertex values in the orientation OctreeRayTracing.vertex_centered_data returns them
        (octree_raytracing.py:43-69): first index z, last index x -- ``set_fields`` swaps axes 0 and 2 back
        (octree_raytracing.py:93-98) so that the sampler finds corner (x, y, z) at x*4 + y*2 + z."""
        return self.ds.vertex_data[self.ds._resolve(field)].swapaxes(0, 2)


class OctreeCellDataset(StreamDataset):
    """Terminal cells of an octree over the unit cube with caller-supplied vertex-centred data.  (The reference
    derives the vertex values from ghost zones of the oct handler -- yt machinery outside the ray-casting path.)"""

    def __init__(self, centres, levels, vertex_data, name="octree_cells"):
        centres = np.ascontiguousarray(centres, dtype="float64")
        levels = np.ascontiguousarray(levels, dtype="int64")
        n = len(levels)
        if centres.shape != (n, 3):
            raise RuntimeError("centres must have shape (n_cells, 3)")
        self.domain_left_edge, self.domain_right_edge = np.zeros(3), np.ones(3)
        self.domain_width, self.domain_center = np.ones(3), np.full(3, 0.5)
        self.domain_dimensions = np.array([2, 2, 2], dtype="int64")
        self.refine_by, self.periodicity, self.basename = 2, (True, True, True), name
        self.min_level, self.max_level = 0, int(levels.max())
        self.index = OctreeIndex()
        dx = 1.0 / (2 << levels).astype("float64")    # level 0 cells are the 8 children of the root oct
        self.cell_arrays = {"x": centres[:, 0], "y": centres[:, 1], "z": centres[:, 2], "dx": dx, "grid_level": levels}
        self.field_info, self.field_list, self.vertex_data = {}, [], {}
        for k, v in vertex_data.items():
            key = k if isinstance(k, tuple) else ("stream", k)
            arr = np.ascontiguousarray(v, dtype="float64")
            if arr.shape != (2, 2, 2, n):
                raise RuntimeError("vertex data of %s must have shape (2, 2, 2, n_cells)" % (k,))
            self.vertex_data[key] = arr
            self.field_info[key] = FieldInfo(key, key[1] in LOG_FIELDS)
            self.field_list.append(key)
        self.derived_field_list = list(self.field_list)
        self.default_field = self._pick_default()
        self.unit_registry = None

    def all_data(self):
        return OctreeAllData(self)


def load_octree_cells(centres, levels, vertex_data):
    """An in-memory octree dataset for ``create_scene`` / ``OctreeVolumeSource``: cell centres (n, 3) in the unit
    cube, their levels (level L cells have dx = 2**-(L+1)) and, per field, the (2, 2, 2, n) vertex values indexed
    (x, y, z) within a cell."""
    return OctreeCellDataset(centres, levels, vertex_data)
