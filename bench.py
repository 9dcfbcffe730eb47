#!/usr/bin/env python
"""bench.py -- the north-star frame: 1024^3 uniform-grid VolumeSource rendered at 2048^2.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A *step* is one frame: cast every ray of the image through the whole volume (VolumeRenderSampler,
n_samples=10, ColorTransferFunction) -- the reference's `for brick in traverse(): sampler(brick)`.

  value     Gsamples/s with bricks, tables and the (zeroed) image already in HBM; timed with CUDA
            events on the launching stream, max over ranks.
  e2e       the same frame through the C-ABI call ytb200_render with HOST (pinned) image and z-buffer:
            H2D + cast + D2H inside the timed region.  Bricks stay device-resident across frames, as
            the reference keeps PartitionedGrids cached on the kd-tree nodes (amr_kdtree.py:326-327).
  roofline  algorithmic HBM bytes of the cast kernel / its measured duration vs the measured copy
            bandwidth (MEASURED_PEAKS.json).  The volume-render kernel is issue-bound, not HBM-bound
            (SURVEY.md 8d): `issue` carries the instruction-issue view next to it.
  cpu_baseline / --impl reference
            the reference's own Cython samplers (compiled unmodified into oracle/_ref) on the box's
            host cores, on a bounded sample (256^3 brick @ 512^2, same view / tables / n_samples).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (cells per side, image side, sampler kind)
    "northstar_vr_1024c_2048px": (1024, 2048, "vr"),
    "cfg2_vr_512c_1024px": (512, 1024, "vr"),
    "cfg3_proj_1024c_2048px": (1024, 2048, "proj"),
    "iproj_1024c_2048px": (1024, 2048, "iproj"),               # InterpolatedProjectionSampler on the north-star volume
    "cfg4_amr_vr_15kbricks_persp_2048px": (128, 2048, "vr"),   # nested-box AMR, PerspectiveLens (BASELINE configs[3])
    "cfg4_amr_vr_24kbricks_persp_2048px": (144, 2048, "vr"),   # the same on 144^3 root cells: 23 328 grids, 23 952 kd leaves ("~20k bricks")
    "cfg5_vr_2048c_4096px": (2048, 4096, "vr"),                # BASELINE configs[4]; 34.4 GB of vertex data
    "tiny_vr_64c_256px": (64, 256, "vr"),
    "tiny_amr_vr_persp_256px": (16, 256, "vr"),
}
AMR_WORKLOADS = {"cfg4_amr_vr_15kbricks_persp_2048px": (128, 8, 4), "cfg4_amr_vr_24kbricks_persp_2048px": (144, 8, 4),
                 "tiny_amr_vr_persp_256px": (16, 8, 3)}
MAX_BRICK_ELEMS = 2 ** 31    # the kernels follow a ray with 32-bit running offsets
N_SAMPLES = 10
SEED = 0


def bench_tf():
    from yt_b200.volume_rendering.transfer_functions import ColorTransferFunction
    tf = ColorTransferFunction((0.0, 1.0), 256)
    cols = [(0.9, 0.2, 0.1), (0.2, 0.8, 0.2), (0.1, 0.3, 0.9), (0.9, 0.9, 0.3)]
    for k, c in enumerate(cols):
        tf.add_gaussian(0.35 + 0.1 * k, 0.02 ** 2, list(c) + [10.0 ** (k - 1)])
    return tf


def bench_camera(res):
    """Default camera of yt.create_scene on a unit-cube dataset (camera.py:127-137): corner view."""
    from yt_b200.synthetic import make_camera
    return make_camera("plane-parallel", resolution=(res, res))


def sampler_call_args(cam, image, zbuffer):
    from yt_b200.volume_rendering.utils import sampler_arguments
    args, kwargs = sampler_arguments(cam, zbuffer=zbuffer, image=image)
    kwargs.pop("camera_data")
    return args, kwargs


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        busy = [s for s in sm if s > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def kd_slab(n, rank, world):
    """Cell range of rank `rank`'s convex sub-volume: kd split of the single grid at depth log2(world),
    always halving the longest remaining axis on a cell boundary."""
    lo, hi = [0, 0, 0], [n, n, n]
    path = []
    size, r = world, rank
    while size > 1:
        ext = [hi[a] - lo[a] for a in range(3)]
        ax = int(np.argmax(ext))
        mid = lo[ax] + ext[ax] // 2
        half = size // 2
        upper = r >= half
        path.append((ax, mid, upper))
        if upper:
            lo[ax] = mid; r -= half
        else:
            hi[ax] = mid
        size = half
    return lo, hi, path


def split_slab(lo, hi, vd, centred):
    """Cut the cell box [lo, hi) into pieces of at most MAX_BRICK_ELEMS stored elements (kd bisection of the
    longest axis on a cell boundary) and return them in the order a ray with direction `vd` meets them: the
    pieces of ONE reference brick must be composited in increasing t (SURVEY.md H14 / finding 6)."""
    ext = [hi[a] - lo[a] for a in range(3)]
    if float(np.prod([e + centred for e in ext])) <= MAX_BRICK_ELEMS:
        return [(list(lo), list(hi))]
    ax = int(np.argmax(ext))
    mid = lo[ax] + ext[ax] // 2
    lower_hi, upper_lo = list(hi), list(lo)
    lower_hi[ax] = mid
    upper_lo[ax] = mid
    lower = split_slab(lo, lower_hi, vd, centred)
    upper = split_slab(upper_lo, hi, vd, centred)
    return lower + upper if vd[ax] >= 0 else upper + lower


def amr_camera(res):
    """PerspectiveLens outside the domain corner looking at the centre (BASELINE configs[3])."""
    from yt_b200.synthetic import make_camera
    return make_camera("perspective", resolution=(res, res), position=[1.6, 1.4, 1.2], width=[0.9, 0.9, 1.0])


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def lsu_pipe_estimate(kind, samples, kernel_ms, sm_mhz, sm_count):
    """Busy fraction of the L1/shared LSU data pipe (one wavefront per cycle per SM).  Shared-memory wavefronts are
    exact by construction (two LDS.128 per sample, 4 wavefronts each, zero bank conflicts); wavefronts per global
    load are the ncu averages (VR, profiles/r02zz_vr_northstar_cast_kernel.txt: 4 corner loads per cell x 6.0 -- the walk
    reuses the face two consecutive cells share; slab-walk projection, profiles/r02zz_proj_cfg3_proj_kernel.txt: 3.3
    wavefronts per warp of cell visits)."""
    if kind == "vr":
        cells = samples / N_SAMPLES
        wf = samples / 32.0 * 8.0 + cells / 32.0 * 4 * 6.0
    else:
        wf = samples / 32.0 * 3.32
    cycles = kernel_ms * 1e-3 * sm_mhz * 1e6
    return {"wavefronts": wf, "frac": wf / sm_count / cycles,
            "source": "LDS wavefronts counted, LDG wavefronts per request from ncu (profiles/r02zz_vr_northstar_cast_kernel.txt, "
                      "r02zz_proj_cfg3_proj_kernel.txt); ncu measured 0.895 (VR) / 0.595 (PROJ) on the same kernels"}


def recorded_traffic(workload):
    """DRAM bytes per launch of the cast kernel from the committed ncu --set full capture of this workload at N=1."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(workload)
    return None


def roofline_block(workload, kind, world, my_samples, kernel_ms, alg_bytes, clk, sm_count):
    """The dominant kernel against the roof that bounds it.  The volume-render and interpolated-projection casts do
    7.7 samples per HBM byte: their roof is the L1/shared LSU data pipe (one wavefront per cycle per SM; ncu
    `l1tex__data_pipe_lsu_wavefronts`), so `bound`/`achieved`/`peak`/`frac` describe THAT pipe and the HBM figure
    sits beside it as `hbm_frac`.  The ProjectionSampler cast reads every cell once and does ~10 instructions per
    cell: its roof is HBM.  `traffic` is the recorded N=1 ncu figure of this workload (null at N > 1: not measured)."""
    peak, peak_src, sm_max = measured_peak()
    sm_mhz = (clk or {}).get("sm_mhz") or sm_max
    hbm = alg_bytes / (kernel_ms * 1e-3) / 1e9
    lsu = lsu_pipe_estimate(kind, my_samples, kernel_ms, sm_mhz, sm_count)
    issue = {"samples_per_s_per_sm": my_samples / (kernel_ms * 1e-3) / sm_count,
             "lane_slots_per_sample_at_sm_clock": (sm_count * 128 * sm_mhz * 1e6) / (my_samples / (kernel_ms * 1e-3))}
    traffic = recorded_traffic(workload) if world == 1 else None
    common = {"kernel": "proj_kernel" if kind == "proj" else "cast_kernel", "kernel_ms": kernel_ms, "algorithmic_bytes": alg_bytes,
              "traffic": traffic, "traffic_source": "recorded: ncu --set full of this workload at N=1 (profiles/traffic.json), not this run" if traffic else None,
              "hbm_achieved_gbs": hbm, "hbm_peak_gbs": peak, "hbm_frac": hbm / peak, "peak_source": peak_src,
              "lsu_pipe": lsu, "issue": issue}
    if kind == "proj":
        common.update({"bound": "hbm", "achieved": hbm, "peak": peak, "unit": "GB/s", "frac": hbm / peak})
    elif kind == "iproj":
        # no table lookups: the cast is the cell walk itself, bound by instruction issue (ncu r02y: 123 warp-instructions per
        # warp-cell, issue 61 % at two CTAs per SM); reported against the issue rate of the chip
        cells = my_samples / N_SAMPLES
        wi = 123.0 * cells / 32.0
        peak_wi = sm_count * 4 * sm_mhz * 1e-3          # G warp-instructions per second
        common["lsu_pipe"] = None
        common.update({"bound": "issue", "achieved": wi / (kernel_ms * 1e-3) / 1e9, "peak": peak_wi, "unit": "Gwarp-instructions/s",
                       "frac": wi / (kernel_ms * 1e-3) / 1e9 / peak_wi,
                       "note": "123 warp-instructions per warp-cell from ncu (profiles/r02y_iproj_cast_kernel.txt), counted cells from this run"})
    else:
        gw = lsu["wavefronts"] / (kernel_ms * 1e-3) / 1e9
        common.update({"bound": "l1_lsu_pipe", "achieved": gw, "peak": sm_count * sm_mhz * 1e-3, "unit": "Gwavefronts/s",
                       "frac": lsu["frac"],
                       "note": "7.7 samples per HBM byte: the HBM fraction is structurally 1-2 % (hbm_frac); the unit that is full is the "
                               "L1/shared LSU data pipe (two LDS.128 table rows per sample + 4 corner loads per cell), then instruction issue"})
    return common


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own samplers (oracle/_ref) or, if absent, the C oracle port
# ------------------------------------------------------------------------------------------------
def cpu_render_sample(kind, n=256, res=512, threads=0, repeats=1):
    from oracle import oracle as O
    from yt_b200.loaders import vertex_centre_no_ghost as vertex_centre
    from yt_b200.synthetic import host_cells
    if O.ref_available():
        S, G = O.load_ref()
        G = G  # partitioned_grid module
        label = "reference"
    else:
        S, G, label = O, O, "port"
    cells = host_cells(n, SEED)
    data = cells if kind == "proj" else vertex_centre(cells)
    cam = bench_camera(res)
    threads = threads or os.cpu_count()
    pg = G.PartitionedGrid(0, [np.ascontiguousarray(data)], np.ones((n, n, n), "uint8"), np.zeros(3), np.ones(3),
                           np.array([n] * 3, "int64"))
    best = None
    for _ in range(repeats):
        img = np.zeros((res, res, 4))
        args, kwargs = sampler_call_args(cam, img, None)
        if kind == "vr":
            s = S.VolumeRenderSampler(*args, bench_tf(), N_SAMPLES, **kwargs)
        elif kind == "iproj":
            s = S.InterpolatedProjectionSampler(*args, N_SAMPLES, **kwargs)
        else:
            s = S.ProjectionSampler(*args, N_SAMPLES, **kwargs)
        t0 = time.perf_counter()
        s(pg, num_threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    # exact sample count of the sample frame from the C oracle's counter
    img = np.zeros((res, res, 4))
    args, kwargs = sampler_call_args(cam, img, None)
    oc = (O.VolumeRenderSampler(*args, bench_tf(), N_SAMPLES, **kwargs) if kind == "vr" else
          (O.InterpolatedProjectionSampler(*args, N_SAMPLES, **kwargs) if kind == "iproj" else O.ProjectionSampler(*args, N_SAMPLES, **kwargs)))
    oc(O.PartitionedGrid(0, [np.ascontiguousarray(data)], np.ones((n, n, n), "uint8"), np.zeros(3), np.ones(3), np.array([n] * 3, "int64")),
       num_threads=0)
    samples = oc.cells * (1 if kind == "proj" else N_SAMPLES)
    return {"seconds": best, "samples": samples, "kind": label, "cores": threads,
            "sample": "%d^3 brick @ %d^2, same view/tables/n_samples=%d, best of %d" % (n, res, N_SAMPLES, repeats)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, res, kind = WORKLOADS[args.workload]
    times = []
    info = None
    for i in range(args.warmup + args.steps):
        info = cpu_render_sample(kind, args.cpu_size, args.cpu_res, threads=0, repeats=1)
        if i >= args.warmup:
            times.append(info["seconds"])
    ms = 1e3 * float(np.mean(times))
    gs = info["samples"] / (ms * 1e-3) / 1e9
    line = {"impl": "reference", "metric": "Gsamples/s", "value": gs, "unit": "Gsamples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "step": "bounded CPU sample of the workload: " + info["sample"]},
            "cpu_baseline": {"value": gs, "unit": "Gsamples/s", "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
            "e2e": {"value": gs, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def single_rank_image(workload, ctx, frame, cam, sargs, tf, kcode, kind, dev):
    """The workload's frame cast by THIS rank alone over the whole volume (world == 1 code path: one launch, no
    compositing), on the caller's frame; returns the host image.  Reference for `parity_vs_n1`."""
    from yt_b200 import _lib as L
    from yt_b200 import device as D
    from yt_b200 import synthetic
    n, res, _ = WORKLOADS[workload]
    if workload in AMR_WORKLOADS:
        from yt_b200 import loaders
        from yt_b200.amr_kdtree import AMRKDTree
        base, tile, levels = AMR_WORKLOADS[workload]
        grids, dd = synthetic.nested_amr_grids(base_cells=base, tile=tile, levels=levels, seed=SEED)
        ds = loaders.load_amr_grids(grids, dd)
        kd = AMRKDTree(ds, comm_rank=0, comm_size=1)
        kd.set_fields([ds.default_field], [False], True)
        bricks = kd.device_bricks(ctx, L.STORE_F32)
        order = np.ascontiguousarray(kd.view_order(cam.lens.viewpoint), dtype=np.int64)
    else:
        centred = 0 if kind == "proj" else 1
        vd = np.asarray(sargs[1], dtype=np.float64).ravel()[:3]
        pieces = split_slab([0, 0, 0], [n, n, n], vd, centred)
        bricks = D.BrickSet(ctx, len(pieces), 1, L.CELL_CENTRED if kind == "proj" else L.VERTEX_CENTRED, L.STORE_F32)
        for i, (plo, phi) in enumerate(pieces):
            if kind != "proj":
                vol, pitch = synthetic.vertex_slab(n, plo, phi, SEED, dev), None
            else:
                vol, pitch = synthetic.cell_slab(n, plo, phi, SEED, dev)
            bricks.attach(i, [vol.data_ptr()], 0, [plo[a] / n for a in range(3)], [phi[a] / n for a in range(3)],
                          [phi[a] - plo[a] for a in range(3)], keep=vol, row_pitch=pitch)
        bricks.commit()
        order = np.arange(len(pieces), dtype=np.int64)
    frame.clear()
    frame.cast(kcode, bricks, order, tf=tf, n_samples=N_SAMPLES)
    out = frame.download(np.empty((res, res, 4)))
    del bricks
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from yt_b200 import _lib as L
    from yt_b200 import device as D
    from yt_b200 import synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun with --nproc-per-node %d" % (args.gpus, args.gpus))
    local = local % max(torch.cuda.device_count(), 1)   # a launcher may expose one device per rank
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, res, kind = WORKLOADS[args.workload]
    stream = torch.cuda.Stream(device=dev)
    ctx = D.Context(local, stream=stream.cuda_stream)
    kcode = {"vr": L.VOLUME_RENDER, "proj": L.PROJECTION, "iproj": L.INTERPOLATED_PROJECTION}[kind]

    amr = args.workload in AMR_WORKLOADS
    with torch.cuda.stream(stream):
        tf = D.TransferTables(ctx, bench_tf()) if kind == "vr" else None
        cam = amr_camera(res) if amr else bench_camera(res)
        img_host = D.pinned_empty((res, res, 4))
        zb_host = D.pinned_empty((res, res))
        img_host[:] = 0.0
        zb_host[:] = 1.0
        from yt_b200 import image_samplers as S
        from yt_b200.volume_rendering.utils import sampler_arguments
        sargs, skw = sampler_arguments(cam, zbuffer=zb_host, image=img_host)
        if not amr:
            skw.pop("camera_data")
        skw["context"] = ctx
        sampler = (S.VolumeRenderSampler(*sargs, bench_tf(), N_SAMPLES, **skw) if kind == "vr"
                   else (S.InterpolatedProjectionSampler(*sargs, N_SAMPLES, **skw) if kind == "iproj"
                         else S.ProjectionSampler(*sargs, N_SAMPLES, **skw)))
        frame = sampler._frame
        centred = 0 if kind == "proj" else 1
        if amr:
            # ---- nested-box AMR hierarchy -> kd-tree leaves of this rank's subtree, resident in HBM ----
            from yt_b200 import loaders
            from yt_b200.amr_kdtree import AMRKDTree
            base, tile, levels = AMR_WORKLOADS[args.workload]
            grids, dd = synthetic.nested_amr_grids(base_cells=base, tile=tile, levels=levels, seed=SEED)
            ds = loaders.load_amr_grids(grids, dd)
            kd = AMRKDTree(ds, comm_rank=rank, comm_size=world)
            kd.set_fields([ds.default_field], [False], True)
            bricks = kd.device_bricks(ctx, L.STORE_F32)
            order = np.ascontiguousarray(kd.view_order(cam.lens.viewpoint), dtype=np.int64)
            path = kd.kd_path() if world > 1 else []
            my_elems = float(sum(np.prod(np.asarray(kd.get_brick_data(int(i)).my_data[0].shape)) for i in kd.tree.traverse()))
            n_bricks = len(order)
            cells_desc = "nested-box AMR, %d levels over %d^3 root cells, %d-cell tiles: %d grids" % (levels, base, tile, len(grids))
        else:
            # ---- this rank's convex sub-volume of the single grid, resident in HBM -------------------
            lo, hi, path = kd_slab(n, rank, world)
            if args.shard:     # one GPU casting the sub-volume rank R of P would own (kernel studies; not a bench line)
                r_, p_ = (int(x) for x in args.shard.split("/"))
                lo, hi, _ = kd_slab(n, r_, p_)
            vd = np.asarray(sargs[1], dtype=np.float64).ravel()[:3]
            pieces = split_slab(lo, hi, vd, centred)
            bricks = D.BrickSet(ctx, len(pieces), 1, L.CELL_CENTRED if kind == "proj" else L.VERTEX_CENTRED, L.STORE_F32)
            my_elems = 0.0
            for i, (plo, phi) in enumerate(pieces):
                if kind != "proj":
                    vol, pitch = synthetic.vertex_slab(n, plo, phi, SEED, dev), None
                else:
                    vol, pitch = synthetic.cell_slab(n, plo, phi, SEED, dev)   # rows padded off the 4096-byte stride
                my_elems += float(vol.numel())
                bricks.attach(i, [vol.data_ptr()], 0, [plo[a] / n for a in range(3)], [phi[a] / n for a in range(3)],
                              [phi[a] - plo[a] for a in range(3)], keep=vol, row_pitch=pitch)
            bricks.commit()
            order = np.arange(len(pieces), dtype=np.int64)
            n_bricks = len(pieces)
            cells_desc = [n, n, n]
        # multi-GPU compositing: binary swap over NVLink peer memory when every rank can map its peers, else the NCCL
        # send/recv compositor (still entirely on the GPUs); YTB200_SWAP=nccl forces the baseline
        swap = partial = None
        peer = False
        if world > 1:
            from yt_b200.parallel import make_compositor, partial_format
            fmt = partial_format(kcode, tf, bricks.n_fields)
            swap, peer = make_compositor(ctx, frame, path, cam.lens.viewpoint, rank, world, fmt=fmt)
            if not peer:
                partial = torch.empty((res, res, 8), dtype=torch.float64, device=dev)

        def cast_partial(timed=False):
            if peer:
                return swap.cast(frame, kcode, bricks, order, tf=tf, n_samples=N_SAMPLES, timed=timed)
            return frame.cast_partial(kcode, bricks, order, partial.data_ptr(), tf=tf, n_samples=N_SAMPLES, timed=timed)

        def composite():
            if peer:
                swap.composite(frame)
            else:
                swap.composite(partial, frame)
        stream.synchronize()

        # ---- exact work per frame ----------------------------------------------------------------
        frame.clear()
        st = frame.count(kcode, bricks, order, N_SAMPLES)
        my_samples = int(st.n_samples)
        if world > 1:
            t = torch.tensor([my_samples], dtype=torch.int64, device=dev)
            dist.all_reduce(t)
            total_samples = int(t.item())
        else:
            total_samples = my_samples

        def frame_device():
            if world == 1:
                frame.clear()
                frame.cast(kcode, bricks, order, tf=tf, n_samples=N_SAMPLES)
            else:
                frame.clear()
                cast_partial()
                composite()

        def frame_e2e():
            if world == 1:
                frame.render(kcode, bricks, order, img_host, zbuffer=zb_host, tf=tf, n_samples=N_SAMPLES)
            elif peer:
                # ONE page-locked image / z-buffer mapped into every rank's process: each rank reads max_t and its strip of
                # the initial colours in place and its last swap kernel stores the finished strip back over its own PCIe
                # link; rank 0's stream ends with a wait for every rank's strip
                cast_partial()
                swap.composite_host(frame, img_shared)
                ctx.synchronize()
            else:
                frame.upload(img_host, zb_host)
                cast_partial()
                composite()
                if rank == 0:
                    frame.download(img_host)
                else:
                    ctx.synchronize()

        def timed(fn, steps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms

        # the clock sampler (nvidia-smi -lms 100) starts BEFORE the warm-up: its NVML start-up stalls the GPU for a few
        # milliseconds once, which must not land in the timed region (r02zz: 41.69 ms per step over 5 steps against 41.13
        # over 10); it keeps sampling through the timed region
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
            time.sleep(0.3)
        for _ in range(max(args.warmup, 3)):
            frame_device()
        ctx.synchronize()
        ms_total = timed(frame_device, args.steps)
        clk = clocks.stop() if rank == 0 else None
        ms_step = ms_total / args.steps

        # kernel-only duration of the dominant (cast) kernel, CUDA events on its stream
        kt = []
        for _ in range(min(args.steps, 5)):
            if world == 1:
                frame.clear()
                s = frame.cast(kcode, bricks, order, tf=tf, n_samples=N_SAMPLES, timed=True)
            else:
                s = cast_partial(timed=True)
                composite()   # keeps the ranks' frame counters in step
            kt.append(s.kernel_ms)
        kernel_ms = float(np.mean(kt))
        rank_kernel_ms = [kernel_ms]
        if world > 1:
            t = torch.zeros(world, dtype=torch.float64, device=dev)
            t[rank] = kernel_ms
            dist.all_reduce(t)
            rank_kernel_ms = [float(x) for x in t.tolist()]

        # ---- the N-rank frame against the SAME frame cast by rank 0 alone (SCALE lines carry multi-GPU parity) ----
        parity_vs_n1 = None
        dev_img = None
        if world > 1:
            frame_device()
            ctx.synchronize()
            if rank == 0:
                dev_img = frame.download(np.empty((res, res, 4)))
            if not args.no_parity and (amr or n <= 1024):
                if rank == 0:
                    solo = single_rank_image(args.workload, ctx, frame, cam, sargs, tf, kcode, kind, dev)
                    from yt_b200.metrics import normalised_diff, projection_diff
                    parity_vs_n1 = normalised_diff(dev_img, solo, False) if kind == "vr" else projection_diff(dev_img[:, :, 0], solo[:, :, 0])
                    del solo
                    torch.cuda.empty_cache()
                dist.barrier()
        img_shared = zb_shared = None
        if peer:
            from yt_b200.parallel import shared_pinned_empty
            img_shared = shared_pinned_empty((res, res, 4), rank=rank, world=world, tag="img")
            zb_shared = shared_pinned_empty((res, res), rank=rank, world=world, tag="zb")
            if rank == 0:
                img_shared[:] = 0.0
                zb_shared[:] = 1.0
            dist.barrier()
            frame.bind_host(img_shared, zb_shared)
        for _ in range(2):
            frame_e2e()
        ms_e2e = timed(frame_e2e, args.steps) / args.steps
        if peer:
            swap.check()   # raises if a device-side wait of the swap timed out
            # one more frame from zeroed initial colours: the host-route image must equal the device-route image
            dist.barrier()
            if rank == 0:
                img_shared[:] = 0.0
            dist.barrier()
            frame_e2e()
            swap.check()
            dist.barrier()
            frame.bind_host(None)
            if rank == 0 and not np.array_equal(img_shared, dev_img):
                raise SystemExit("e2e image differs from the device-route image (max |diff| %g)" % np.abs(img_shared - dev_img).max())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    elem = 4
    alg_bytes = my_elems * elem + 2.0 * res * res * 32 + res * res * 8 + (2.0 * res * res * 24 if amr else 0.0)
    gs = total_samples / (ms_step * 1e-3) / 1e9
    gs_e2e = total_samples / (ms_e2e * 1e-3) / 1e9
    launches_per_step = 2 if world == 1 else 2 + swap.launches_per_composite   # clear + cast (+ rounds + apply)
    line = {
        "metric": "Gsamples/s", "value": gs, "unit": "Gsamples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32 samples / f64 traversal", "data": "synthetic",
        "config": {"workload": args.workload, "cells": cells_desc, "image": [res, res], "sampler": kind,
                   "n_samples": 1 if kind == "proj" else N_SAMPLES, "lens": "perspective" if amr else "plane-parallel",
                   "view": "camera at (1.6,1.4,1.2) looking at the domain centre" if amr else "default corner camera, width 1.5",
                   "bricks_per_rank": n_bricks, "samples_per_frame": total_samples, "parallelism": "kd%d" % world,
                   "composite": ("none" if world == 1 else ("binary swap over NVLink peer memory, %d B per pixel (partial format %d)" % (swap.bytes_per_pixel, swap.fmt)
                                 if peer else "binary swap, NCCL send/recv, fp64 (A,B)")),
                   "l2": "brick (%.2f GB per rank) larger than L2; no flush needed" % (my_elems * elem / 1e9)},
        "ms_per_frame": ms_step,
        "e2e": {"value": gs_e2e, "unit": "Gsamples/s", "ms_per_frame": ms_e2e,
                "h2d_bytes_per_step": res * res * 40 if not peer else res * res * 8 * world + res * res * 32,
                "d2h_bytes_per_step": res * res * 32,
                "note": (("host (page-locked) image+zbuffer in, host image out through ytb200_render: staged in 8 row bands (copy-in, cast, copy-out overlapped on three streams); bricks device-resident across frames" if kind == "proj" else
                          "host (page-locked) image+zbuffer in, host image out through ytb200_render: the cast kernel reads/writes the host arrays in place over PCIe; bricks device-resident across frames") if world == 1 else "one page-locked image + z-buffer mapped into every rank: each rank reads max_t in place, fetches its strip of the initial colours with the copy engine during the cast and sends its finished strip out with one bulk copy over its own PCIe link (ytb200_swap_composite_host); checked equal to the device-route image")},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clk,
        "roofline": roofline_block(args.workload, kind, world, my_samples, kernel_ms, alg_bytes, clk, ctx.sm_count),
    }
    if world > 1:
        line["rank_kernel_ms"] = rank_kernel_ms
    if args.shard:
        line["config"]["shard"] = args.shard
    if parity_vs_n1 is not None:
        line["parity_vs_n1"] = {"value": parity_vs_n1, "bar": 1e-4 if kind == "vr" else 1e-6,
                                "what": ("normalised max |RGBA difference| (SURVEY.md A.5)" if kind == "vr" else "max |difference| / max |reference|")
                                + " between this N-rank composited frame and the same frame cast by rank 0 alone in one launch"}
    if world == 1 and not args.no_cpu and not amr and n <= 1024:
        info = cpu_render_sample(kind, args.cpu_size, args.cpu_res, threads=0, repeats=2)
        v = info["samples"] / info["seconds"] / 1e9
        line["cpu_baseline"] = {"value": v, "unit": "Gsamples/s", "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="northstar_vr_1024c_2048px", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-size", type=int, default=256)
    ap.add_argument("--cpu-res", type=int, default=512)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--shard", default="", help="R/P at --gpus 1: cast only the sub-volume rank R of P would own (kernel studies)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip rank 0's single-GPU render of the same frame (parity_vs_n1)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
