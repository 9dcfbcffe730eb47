"""Object-space multi-GPU rendering: kd partition + ordered binary-swap compositing.

Replaces ``AMRKDTree.reduce_tree_images`` (yt/utilities/amr_kdtree/amr_kdtree.py:297-323,
amr_kdtools.py:6-50), which tree-reduces *full* fp64 images to one rank over MPI.  Here every
rank renders its convex kd subtree into a per-pixel affine partial ``C -> A + B*C`` (SURVEY.md
A.4), and log2(P) rounds of pairwise half-image exchanges (``torch.distributed`` P2P over
NCCL/NVLink) with our compose kernel in between leave each rank with a fully composited 1/P
strip; strips are then gathered on rank 0.  Per-GPU traffic is (1 - 1/P) of one partial image.

Order: the partner pair of round j are the two children of the kd split at depth k-1-j.  Which
child is applied first follows the reference's traversal rule (amr_kdtools.pyx:708-755):
``viewpoint[dim] <= pos`` -> the upper child is farther and therefore composited first.
"""
import os

import ctypes as C
import weakref

import numpy as np

# ---- communicator: one process per GPU (torchrun), torch.distributed for the plumbing -------------------
_comm = {"rank": 0, "size": 1, "group": None, "backend": None}


def enable_parallelism(backend=None):
    """``yt.enable_parallelism()`` for the ray caster: join the torchrun process group (NCCL when CUDA is
    visible, gloo otherwise) and bind this process to GPU ``LOCAL_RANK``.  Returns (rank, size)."""
    size = int(os.environ.get("WORLD_SIZE", "1"))
    if size == 1:
        return 0, 1
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    _comm.update(rank=dist.get_rank(), size=dist.get_world_size(), backend=dist.get_backend())
    if _comm["size"] & (_comm["size"] - 1):
        raise RuntimeError("the kd-tree partition needs a power-of-two number of ranks (got %d)" % _comm["size"])
    return _comm["rank"], _comm["size"]


def comm_rank():
    return _comm["rank"]


def comm_size():
    return _comm["size"]


def kd_path_of_rank(nodes, rank, size):
    """[(axis, split_pos, is_upper)] from the root split down to the subtree owned by `rank`: heap node
    ``size + rank`` (amr_kdtools.pyx:670-676, amr_kdtree.py:283-295).  `nodes` is Tree.nodes.  Levels that
    do not exist (an empty region was never split) get a dummy split: both sides are empty there."""
    k = int(round(np.log2(size))) if size > 1 else 0
    by_id = {int(i): n for n, i in enumerate(nodes["node_id"])}
    path = []
    target = size + rank
    for level in range(k):
        anc = target >> (k - level)          # ancestor at this depth
        upper = (target >> (k - level - 1)) & 1
        n = by_id.get(anc)
        if n is None or nodes["split_dim"][n] < 0:
            path.append((0, 0.0, bool(upper)))
        else:
            path.append((int(nodes["split_dim"][n]), float(nodes["split_pos"][n]), bool(upper)))
    return path


def bit_reverse(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def plan_binary_swap(rank, world, n_pixels, path, viewpoint):
    """Pure schedule: list of rounds and the final strip of `rank`.

    path: [(axis, split_pos, is_upper)] from the root split to this rank's leaf (len = log2(world)).
    Returns (rounds, (p0, p1)); each round is a dict with the partner, the pixel range kept
    ("keep"), the range sent ("send") and whether this rank's content is applied first ("earlier").
    """
    k = int(round(np.log2(world))) if world > 1 else 0
    if (1 << k) != world:
        raise ValueError("binary swap needs a power-of-two number of ranks (got %d)" % world)
    if len(path) != k:
        raise ValueError("kd path length %d does not match log2(world)=%d" % (len(path), k))
    p0, p1 = 0, int(n_pixels)
    rounds = []
    for j in range(k):
        ax, pos, upper = path[k - 1 - j]
        assert bool(upper) == bool((rank >> j) & 1), "rank bits must follow the kd path"
        mid = p0 + (p1 - p0) // 2
        keep_lower = ((rank >> j) & 1) == 0
        keep = (p0, mid) if keep_lower else (mid, p1)
        send = (mid, p1) if keep_lower else (p0, mid)
        upper_first = bool(viewpoint[ax] <= pos)
        rounds.append({"partner": rank ^ (1 << j), "keep": keep, "send": send,
                       "earlier": bool(upper) == upper_first, "axis": ax, "pos": pos})
        p0, p1 = keep
    return rounds, (p0, p1)


def strip_bounds(world, n_pixels):
    """Final strip [p0, p1) of every rank (depends only on the rank bits)."""
    k = int(round(np.log2(world))) if world > 1 else 0
    out = []
    for r in range(world):
        p0, p1 = 0, int(n_pixels)
        for j in range(k):
            mid = p0 + (p1 - p0) // 2
            p0, p1 = (p0, mid) if ((r >> j) & 1) == 0 else (mid, p1)
        out.append((p0, p1))
    return out


class _DevArray:
    """Zero-copy view of raw device memory for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class BinarySwap:
    """Executes the schedule.  `compose(earlier, later, out)` and `apply(partial, image)` default to the
    CUDA kernels behind the C ABI; tests inject torch-CPU versions to exercise the schedule over gloo."""

    def __init__(self, ctx, nx, ny, path, viewpoint, rank, world, compose=None, apply=None, group=None):
        import torch
        self.torch = torch
        self.ctx, self.nx, self.ny, self.rank, self.world, self.group = ctx, nx, ny, rank, world, group
        self.npx = nx * ny
        self.rounds, self.strip = plan_binary_swap(rank, world, self.npx, path, viewpoint)
        self.strips = strip_bounds(world, self.npx)
        self._compose = compose or self._compose_cuda
        self._apply = apply or self._apply_cuda
        self._recv = None
        self.launches_per_composite = len(self.rounds) + 1

    # -- CUDA kernels through the C ABI ---------------------------------------------------------------
    def _compose_cuda(self, earlier, later, out):
        from . import _lib as L
        L.check(L.load().ytb200_partial_compose(self.ctx.handle, earlier.data_ptr(), later.data_ptr(), out.data_ptr(), out.shape[0]))

    def _apply_cuda(self, partial, image):
        from . import _lib as L
        L.check(L.load().ytb200_partial_apply(self.ctx.handle, partial.data_ptr(), image.data_ptr(), image.shape[0]))

    def _exchange(self, send, recv, partner):
        dist = self.torch.distributed
        ops = [dist.P2POp(dist.isend, send, partner, group=self.group), dist.P2POp(dist.irecv, recv, partner, group=self.group)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    def composite_tensors(self, partial, image):
        """partial: (npx, 8) tensor of this rank; image: (npx, 4) tensor holding the initial colours C_0
        (only this rank's final strip is read).  Returns the strip (p0, p1) now final in `image`."""
        torch = self.torch
        cur = partial.view(self.npx, 8)
        if self._recv is None or self._recv.device != cur.device:
            self._recv = torch.empty((max(self.npx // 2 + 1, 1), 8), dtype=cur.dtype, device=cur.device)
        for rd in self.rounds:
            k0, k1 = rd["keep"]
            s0, s1 = rd["send"]
            recv = self._recv[: k1 - k0]
            self._exchange(cur[s0:s1], recv, rd["partner"])
            mine = cur[k0:k1]
            if rd["earlier"]:
                self._compose(mine, recv, mine)
            else:
                self._compose(recv, mine, mine)
        p0, p1 = self.strip
        self._apply(cur[p0:p1], image[p0:p1])
        return p0, p1

    def gather_strips(self, image, dst=0):
        """Collect every rank's final strip of `image` (npx, 4) on rank `dst`."""
        dist = self.torch.distributed
        if self.world == 1:
            return
        ops = []
        if self.rank == dst:
            for r, (p0, p1) in enumerate(self.strips):
                if r != dst and p1 > p0:
                    ops.append(dist.P2POp(dist.irecv, image[p0:p1], r, group=self.group))
        else:
            p0, p1 = self.strip
            if p1 > p0:
                ops.append(dist.P2POp(dist.isend, image[p0:p1], dst, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    # -- frame-level convenience used by the bench and the scene mirror -----------------------------------
    def composite(self, partial, frame, broadcast=False):
        """Compose this rank's partial with all others and leave the final image in rank 0's frame (in
        every rank's frame with ``broadcast``, like scatter_image of amr_kdtools.py:43-50).
        The frame's device image must already hold the initial colours C_0 (``frame.clear()`` or
        ``frame.upload()``), exactly what it held when the partial was cast."""
        image = self.torch.as_tensor(_DevArray(frame.device_image, (self.npx, 4)), device=partial.device)
        self.composite_tensors(partial, image)
        if broadcast and self.world > 1:
            self.allgather_strips(image)
        else:
            self.gather_strips(image, 0)

    def allgather_strips(self, image):
        """Every rank ends up with every strip (strips have equal length up to the halving remainders)."""
        dist = self.torch.distributed
        for r, (p0, p1) in enumerate(self.strips):
            if p1 > p0:
                dist.broadcast(image[p0:p1], src=r, group=self.group)


class PeerSwap:
    """Binary swap over NVLink peer memory (``ytb200_swap_*``): the same schedule as :class:`BinarySwap`, but each
    round is one CUDA kernel that waits for the partner's flag, reads the partner's half directly from its HBM and
    signals the next reader, and the last round stores the composed strips straight into the destination ranks'
    slabs.  Partials travel in the smallest exact format of the affine map (``partial_format``: 16 B per pixel for the
    ColorTransferFunction render, 8 B per summed field for the projections).  The object is independent of any one
    frame: create it once per (image size, world, format) and ``set_view`` it when the camera moves.
    torch.distributed is used once, at construction, to pass the CUDA IPC handles around; a frame issues no collective.

        swap = PeerSwap(ctx, frame, path, viewpoint, rank, world, fmt=partial_format(kind, tf, n_fields, precision))
        swap.cast(frame, kind, bricks, order, tf=tf)
        swap.composite(frame)            # result in rank 0's frame image (every rank's with broadcast=True)
        swap.composite_host(frame, img)  # or: every rank stores its strip into the shared page-locked `img`
    """

    def __init__(self, ctx, frame, path, viewpoint, rank, world, f32=True, group=None, fmt=None):
        import torch.distributed as dist
        from . import _lib as L
        lib = L.load()
        if fmt is None:
            fmt = L.PARTIAL_AB32 if f32 else L.PARTIAL_AB64
        self.ctx, self.rank, self.world, self.fmt = ctx, rank, world, int(fmt)
        self.f32 = self.fmt in (L.PARTIAL_AB32, L.PARTIAL_A32, L.PARTIAL_A32G)
        self.bytes_per_pixel = int(lib.ytb200_partial_bytes(self.fmt))
        self.npx = frame.nx * frame.ny
        h = C.c_void_p()
        L.check(lib.ytb200_swap_create(ctx.handle, self.npx, rank, world, self.fmt, C.byref(h)))
        self.handle = h
        self._fin = weakref.finalize(self, lib.ytb200_swap_destroy, h)
        if world > 1:
            mine_slab = C.create_string_buffer(64)
            L.check(lib.ytb200_swap_handle(h, mine_slab))
            everyone = [None] * world
            dist.all_gather_object(everyone, (rank, mine_slab.raw), group=group)
            for r, slab in everyone:
                if r != rank:
                    L.check(lib.ytb200_swap_open(h, int(r), slab))
        self.set_view(path, viewpoint)
        # launches of one frame's composite on rank 0: one per round + the apply kernel
        self.launches_per_composite = len(self.rounds) + 1

    def set_view(self, path, viewpoint):
        """(Re)plan the rounds for a viewpoint: which side of every kd split is composited first."""
        from . import _lib as L
        self.rounds, self.strip = plan_binary_swap(self.rank, self.world, self.npx, path, viewpoint)
        n = len(self.rounds)
        partner = np.ascontiguousarray([int(r["partner"]) for r in self.rounds], dtype=np.int32)
        earlier = np.ascontiguousarray([1 if r["earlier"] else 0 for r in self.rounds], dtype=np.int32)
        p0 = np.zeros(n, dtype=np.int64)
        p1 = np.zeros(n, dtype=np.int64)
        k0 = np.ascontiguousarray([r["keep"][0] for r in self.rounds], dtype=np.int64)
        k1 = np.ascontiguousarray([r["keep"][1] for r in self.rounds], dtype=np.int64)
        for j, r in enumerate(self.rounds):
            p0[j] = min(r["keep"][0], r["send"][0])
            p1[j] = max(r["keep"][1], r["send"][1])
        L.check(L.load().ytb200_swap_plan(self.handle, n, partner.ctypes.data_as(L.i32p), p0.ctypes.data_as(L.i64p),
                                          p1.ctypes.data_as(L.i64p), k0.ctypes.data_as(L.i64p), k1.ctypes.data_as(L.i64p),
                                          earlier.ctypes.data_as(L.i32p)))

    def cast(self, frame, kind, bricks, order, tf=None, n_samples=10, precision=None, timed=False):
        """This rank's partial cast of the next frame, straight into the slab (``ytb200_swap_cast``)."""
        from . import _lib as L
        o = np.ascontiguousarray(order, dtype=np.int64)
        st = L.Stats() if timed else None
        L.check(L.load().ytb200_swap_cast(
            self.handle, frame.handle, int(kind), bricks.handle, o.ctypes.data_as(L.i64p), int(o.shape[0]),
            tf.handle if tf is not None else None, int(n_samples), int(L.PREC_FAST if precision is None else precision),
            C.byref(st) if timed else None))
        return st

    def composite(self, frame, broadcast=False):
        from . import _lib as L
        L.check(L.load().ytb200_swap_composite(self.handle, frame.handle, 1 if broadcast else 0))

    def composite_host(self, frame, image):
        """Every rank stores the finished pixels of its strip into `image`: the SAME page-locked array mapped into
        every rank's process (:func:`shared_pinned_empty`), holding the initial colours on entry."""
        from . import _lib as L
        L.check(L.load().ytb200_swap_composite_host(self.handle, frame.handle, image.ctypes.data))

    def check(self):
        """Synchronise and raise if a peer failed to arrive (the device-side waits give up after 10 s)."""
        from . import _lib as L
        e = C.c_int(0)
        L.check(L.load().ytb200_swap_error(self.handle, C.byref(e)))
        if e.value:
            raise RuntimeError("binary swap: a peer did not signal within 10 s")


def partial_format(kind, tf, n_fields, precision=None):
    """The partial-image format the cast of (sampler kind, transfer tables, fields, precision) produces."""
    from . import _lib as L
    f = L.load().ytb200_partial_format(int(kind), tf.handle if tf is not None else None, int(n_fields),
                                       int(L.PREC_FAST if precision is None else precision))
    if f < 0:
        L.check(f)
    return int(f)


def shared_pinned_empty(shape, dtype=np.float64, rank=0, world=1, group=None, tag="img"):
    """One host array shared by every rank of the job (POSIX shared memory created by rank 0), page-locked and mapped
    into each rank's CUDA context (``ytb200_host_register``): the image every rank's last swap kernel stores its strip
    into.  Collective.  The mapping lives as long as the returned array."""
    import mmap
    from . import _lib as L
    lib = L.load()
    dt = np.dtype(dtype)
    nbytes = max(int(np.prod(shape)) * dt.itemsize, 1)
    size = (nbytes + mmap.PAGESIZE - 1) // mmap.PAGESIZE * mmap.PAGESIZE
    name = None
    if world > 1:
        import torch.distributed as dist
        box = ["/dev/shm/ytb200_%s_%d_%d" % (tag, os.getpid(), id(shape) & 0xffff)] if rank == 0 else [None]
        if rank == 0:
            with open(box[0], "wb") as fh:
                fh.truncate(size)
        dist.broadcast_object_list(box, src=0, group=group)
        name = box[0]
        fh = open(name, "r+b")
        mm = mmap.mmap(fh.fileno(), size)
        fh.close()
        dist.barrier(group=group)
        if rank == 0:
            os.unlink(name)      # the mappings keep the memory alive
    else:
        mm = mmap.mmap(-1, size)
    arr = np.frombuffer(mm, dtype=dt, count=int(np.prod(shape))).reshape(shape)
    addr = arr.ctypes.data
    L.check(lib.ytb200_host_register(C.c_void_p(addr), size))
    weakref.finalize(mm, lib.ytb200_host_unregister, C.c_void_p(addr))
    return arr


def make_compositor(ctx, frame, path, viewpoint, rank, world, f32=True, group=None, prefer=None, fmt=None):
    """The compositor every rank of the job agrees on: :class:`PeerSwap` when ALL ranks can map each other's slabs
    (one host, at most 8 GPUs, peer access between the devices, CUDA IPC working), :class:`BinarySwap` over NCCL
    send/recv otherwise -- for more than 8 ranks, several nodes, or devices without peer access.  ``prefer``
    ("peer" / "nccl"; default: ``YTB200_SWAP`` or "peer") only expresses a wish: a rank that cannot open a peer makes
    the whole job fall back, it never raises.  Collective: every rank must call it.  Returns (swap, is_peer)."""
    import socket
    import torch
    import torch.distributed as dist
    prefer = prefer or os.environ.get("YTB200_SWAP", "peer")
    swap, ok = None, 0
    if world > 1 and prefer != "nccl" and world <= 8:
        hosts = [None] * world
        dist.all_gather_object(hosts, (socket.gethostname(), int(ctx.device)), group=group)
        same_host = len({h for h, _ in hosts}) == 1
        can = same_host
        if same_host:
            for _, d in hosts:
                if d != ctx.device and not torch.cuda.can_device_access_peer(ctx.device, d):
                    can = False
        # the IPC exchange inside PeerSwap is itself collective: enter it only if every rank will
        flags = [None] * world
        dist.all_gather_object(flags, bool(can), group=group)
        if all(flags):
            try:
                swap = PeerSwap(ctx, frame, path, viewpoint, rank, world, f32=f32, group=group, fmt=fmt)
                ok = 1
            except Exception as exc:  # noqa: BLE001  (cudaIpcOpenMemHandle refused, out of memory, ...)
                import sys
                sys.stderr.write("rank %d: peer-memory swap unavailable (%s); falling back to the NCCL compositor\n" % (rank, exc))
                swap, ok = None, 0
            flags = [None] * world
            dist.all_gather_object(flags, bool(ok), group=group)
            ok = int(all(flags))
    if world > 1 and not ok:
        swap = BinarySwap(ctx, frame.nx, frame.ny, path, viewpoint, rank, world, group=group)
    return swap, bool(ok)
