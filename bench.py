#!/usr/bin/env python3
"""bench.py — rasterisation throughput of the B200 backend on BASELINE.json's workload.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU backend on the host cores

Default workload = the configuration BASELINE.json's metric is quoted on ("at 1/2/4/8 B200"): config 4a,
1M random quad/cubic paths on ONE 16384x16384 canvas (seed 4), which fits one GPU.  At N > 1 the canvas is
partitioned into N contiguous bands of tile rows (STRONG scaling: the same frame, N times the hardware): every
rank holds the part of the display list that can reach its band (skb_display_list_cull_rows, made once per scene on
the host like the encode itself; the device culls again before flattening), renders its band, and its fine pass
stores the finished pixels straight into rank 0's canvas over NVLink peer memory (skb_surface_set_remote_canvas) —
the gather of the north star fused into the producing kernel.  The same gather done by NCCL send/recv after the
frame is timed beside it (`gather`).

A "step" is one frame: clear, flatten, setup, walk, coverage, bin, fine — all paths, all pixels.
  value  Mpix/s of canvas filled, whole job, display list resident in HBM, CUDA events on the surface's stream
         over exactly K steps, max over ranks.
  e2e    the same metric through the C ABI with HOST buffers, all inside the timed region, three frames in flight:
         N = 1: every step uploads the display list from pinned host memory (H2D), renders, reads the canvas back
         into pinned host memory (D2H).  N > 1: every rank uploads its band's list, renders its band and copies it
         into its rows of ONE page-locked host image in shared memory (multigpu.SharedHostImage) — N PCIe links at
         once, the frame whole in rank 0's address space; the route through rank 0's canvas is `e2e.via_rank0_canvas`.
Other workloads (--workload): c1 (10k paths 4096^2; at N > 1 one canvas per rank, weak scaling), c4b (batch
of 1920x1080 canvases split by canvas, 64 per display list), c2, c3.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures (profiles/); None where
# no capture of that (workload, kernel) exists.  Filled in from profiles/r02_*.
NCU_TRAFFIC = {}
try:
    NCU_TRAFFIC = {tuple(k.split("/")): v for k, v in json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).items()}
except Exception:
    pass

METRIC = "canvas_mpix_per_s"
UNIT = "Mpix/s"
C4B_BATCH = 64          # canvases per display list (one launch sequence renders all of them)
C4B_TOTAL = 1024


def workload(name, rank, world):
    """-> (scene or list of scenes, description, partition, scaling)"""
    from skity_b200 import scene
    if name == "c4a":
        return scene.scene_c4a(), "c4a: 1M random quad/cubic paths (128 px), solid fill, ONE 16384x16384 canvas", "bands", "strong"
    if name == "c4a-small":
        return scene.scene_random_fills_fast(20000, 4096, 4, 128.0), "c4a-small: 20k paths 4096x4096 (smoke size)", "bands", "strong"
    if name == "c1":
        return scene.scene_c1(10000, 4096, 1 + rank), "c1: 10k random quad/cubic paths, mixed nonzero/even-odd, solid fill, 4096x4096", "canvas", "weak"
    if name == "c1-small":
        return scene.scene_c1(1000, 1024, 1 + rank), "c1-small: 1k paths 1024x1024 (smoke size)", "canvas", "weak"
    if name == "c3":
        return scene.scene_c3(2000, 8192, 3 + rank), "c3: 2k blurred paths (sigma 4-64) 8192x8192", "canvas", "weak"
    if name == "c2":
        return scene.scene_c2(20000, 4096, 2 + rank), "c2: 20k stroked+filled gradient paths with the nested clip stack 4096x4096", "canvas", "weak"
    if name == "c4b":
        per_rank = C4B_TOTAL // world
        mine = [scene.scene_c4b(i) for i in range(rank * per_rank, rank * per_rank + min(per_rank, C4B_BATCH))]
        return mine, (f"c4b: batch of {C4B_TOTAL} independent 1920x1080 canvases x 1000 paths split by canvas; each step renders "
                      f"{len(mine)} canvases per rank as one display list"), "batch", "weak"
    raise SystemExit(f"unknown workload {name}")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for k, n in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def bind_near_gpu(local_rank):
    """Keeps this process (and the host buffers it touches first) on the NUMA node its GPU hangs off.  -> node or None."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        addr = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{addr}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:  # noqa: BLE001
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from skity_b200 import device, hostlib, multigpu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    numa_node = bind_near_gpu(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    sc, desc, partition, scaling = workload(args.workload, rank, world)
    t_enc = time.perf_counter()
    if partition == "batch":
        blobs = [s.encode() for s in sc]
        dl, canvas_ids = hostlib.encode_scene_batch(blobs)
        W, H, n_canvases = 1920, 1080, len(sc)
        n_paths = sum(s.n_draws for s in sc)
        surf_w = surf_h = 16
        blob = blobs[0]
    else:
        blob = sc.encode()
        dl = hostlib.encode_scene(blob)             # host-side encode (CudaCanvas), outside every timed region
        W, H, n_canvases = sc.width, sc.height, 1
        n_paths = sc.n_draws
        surf_w, surf_h = W, H
    host_encode_s = time.perf_counter() - t_enc
    dl_pinned = torch.empty(len(dl), dtype=torch.uint8).pin_memory()
    dl_pinned.copy_(torch.frombuffer(bytearray(dl), dtype=torch.uint8))
    del dl
    n_dl = dl_pinned.numel()
    bands = multigpu.band_ranges(H, world) if partition == "bands" else None
    banded = partition == "bands" and world > 1
    cull_ms = None
    n_dl_full = n_dl
    if banded:
        # the band's own display list (skb_display_list_cull_rows, host side, once per scene like the encode): a rank
        # uploads and processes what can reach its rows, not N copies of everything
        t_c = time.perf_counter()
        y0b, y1b = bands[rank]
        dl_band = torch.empty(device.cull_display_list_rows((dl_pinned.data_ptr(), n_dl), y0b, y1b, size_only=True),
                              dtype=torch.uint8).pin_memory()
        device.cull_display_list_rows((dl_pinned.data_ptr(), n_dl), y0b, y1b, out=(dl_band.data_ptr(), dl_band.numel()))
        cull_ms = (time.perf_counter() - t_c) * 1e3
        dl_pinned = dl_band
        n_dl = dl_pinned.numel()

    dev = device.Device(local_rank)
    if args.coverage_mode == "area":
        _create = dev.create_surface

        def create_area_surface(w, h):
            sf = _create(w, h)
            sf.set_coverage_mode(1)
            return sf
        dev.create_surface = create_area_surface
    surf = dev.create_surface(surf_w, surf_h)
    stream = torch.cuda.ExternalStream(surf.stream(), device=torch.device("cuda", local_rank))
    if banded:
        surf.set_band(*bands[rank])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # what rank 0 reads back per step
    if partition == "batch":
        out_pinned = torch.empty((n_canvases, H, W, 4), dtype=torch.uint8).pin_memory()
    elif rank == 0 or not banded:
        out_pinned = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
    else:
        out_pinned = None
    out_np = out_pinned.numpy() if out_pinned is not None else None

    def read_back():
        if partition == "batch":
            for i, sid in enumerate(canvas_ids):
                surf.read_batch_canvas(sid, W, H, out=out_np[i])
        elif out_np is not None:
            surf.read_pixels(out=out_np)

    # ---- upload once, warm up
    surf.begin(True)
    surf.encode((dl_pinned.data_ptr(), n_dl))
    surf.flush()
    surf.sync()
    single_check = None
    if banded:
        # first the separate NCCL gather (bands rendered into each rank's own canvas), timed on its own ...
        def band_tensor(a, b):
            return multigpu.surface_band_tensor(surf, a, b)[0]
        gather_ms = []
        for _ in range(3):
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            multigpu.gather_bands(band_tensor, bands, rank, world, dist)
            g1.record()
            torch.cuda.synchronize()
            gather_ms.append(g0.elapsed_time(g1))
        tg = torch.tensor([min(gather_ms)], device="cuda")
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        nccl_gather_ms = float(tg[0])
        if rank == 0:
            single_check = int(surf.read_pixels(0, bands[-1][0], W, min(64, bands[-1][1] - bands[-1][0])).astype(np.uint64).sum())
        # ... then the gather fused into the fine pass: from here on every rank stores its band into rank 0's canvas
        barrier()
        multigpu.fuse_gather_into_fine_pass(surf, rank, dist)
        barrier()

    def step_resident():
        surf.begin(True)
        surf.flush()

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    if banded and rank == 0:
        fused_check = int(surf.read_pixels(0, bands[-1][0], W, min(64, bands[-1][1] - bands[-1][0])).astype(np.uint64).sum())
        if fused_check != single_check:
            raise SystemExit("band stored over NVLink differs from the band gathered by NCCL")
    barrier()

    # ---- device-resident throughput
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = np.zeros(8)
    launches = 0
    st = None
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
        st = surf.stats()      # waits for the frame; per-stage CUDA-event timings of this step
        stage_ms += np.array(st["ms_stage"])
        launches += st["n_launches"]
    e1.record(stream)
    barrier()
    ms_resident = e0.elapsed_time(e1) / args.steps

    # ---- end to end through the C ABI with host buffers: H2D of the display list on every rank, render, the bands
    # land in rank 0's canvas (peer stores), rank 0 reads the canvas back; one frame at a time
    def step_e2e():
        surf.begin(True)
        surf.encode((dl_pinned.data_ptr(), n_dl))
        surf.flush()
        if banded:
            surf.sync()
            dist.barrier()          # every band is in rank 0's canvas
        read_back()
        if banded:
            dist.barrier()          # rank 0 has its copy: the canvas may be overwritten

    step_e2e()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    f0.record(stream)
    for _ in range(args.steps):
        step_e2e()
    f1.record(stream)
    barrier()
    ms_e2e_serial = (time.perf_counter() - t0) * 1e3 / args.steps
    ms_e2e = max(f0.elapsed_time(f1) / args.steps, 0.0)
    ms_e2e_via_rank0 = None

    # ---- banded (N > 1): the host does not need the frame on rank 0's GPU first.  Every rank copies the band it rendered
    # from its own canvas into its rows of ONE host image in shared memory (multigpu.SharedHostImage, page-locked in every
    # process): the canvas reaches the host over N PCIe links at once and sits, whole, in rank 0's address space.  The
    # figures above (bands into rank 0's canvas over NVLink, rank 0 reads everything back) stay as `via_rank0_canvas`.
    n_flight = 3
    shared = None
    shm_ok = [bool(banded and multigpu.SharedHostImage.room_for(n_flight * H * W * 4))]
    if banded:
        dist.broadcast_object_list(shm_ok, src=0)
    band_e2e = banded and shm_ok[0]
    if band_e2e:
        ms_e2e_via_rank0 = ms_e2e_serial
        surf.sync()
        barrier()
        surf.set_remote_canvas(None)          # from here on the band stays in this rank's own canvas
        tag = os.environ.get("MASTER_PORT", "0")
        shared = [multigpu.SharedHostImage(f"skb_bench_{tag}_{j}", (H, W, 4), rank, dist, register=True,
                                           my_rows=None if os.environ.get("SKB_BENCH_SHM_TOUCH") == "owner" else bands[rank]) for j in range(n_flight)]
        y0b, y1b = bands[rank]
        host_image_locked = all(sh.registered for sh in shared)

        def step_e2e_band(sf, img, group=None):
            sf.begin(True)
            sf.encode((dl_pinned.data_ptr(), n_dl))
            sf.flush()
            if y1b > y0b:
                sf.read_pixels_async(img.rows(y0b, y1b), 0, y0b)
            sf.sync()
            dist.barrier(group=group)         # every band of the frame is in the host image

        step_e2e_band(surf, shared[0])
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e_band(surf, shared[0])
        barrier()
        ms_e2e_serial = (time.perf_counter() - t0) * 1e3 / args.steps

    # ---- the same with several frames in flight: every step still uploads its display list and reads its canvas back
    # to pinned host memory, but frames alternate between surfaces (each with its own stream, arenas and canvas), so
    # the upload, the host-side validation and the read-back of one frame overlap the rendering of the others — what an
    # application streaming frames does.  Three surfaces, one host thread each (measured on C4a at N = 1: 1 / 2 / 3
    # frames in flight = 94 / 76 / 66 ms per frame against 60.6 ms of device time).  Banded (N > 1): a frame ends with a
    # host barrier across the ranks (all its bands are in the host image), each thread on a gloo group of its own.
    ms_e2e_wall = ms_e2e_serial
    pipelined = partition != "batch" and (not banded or band_e2e)
    if not pipelined:
        n_flight = 1
    if pipelined:
        extra = [dev.create_surface(surf_w, surf_h) for _ in range(n_flight - 1)]
        flight_groups = [dist.new_group(backend="gloo") for _ in range(n_flight)] if band_e2e else None
        for sb in extra:
            if banded:
                sb.set_band(*bands[rank])
        pair = [surf] + extra
        if band_e2e:
            outs = [sh.array if rank == 0 else None for sh in shared]
        else:
            outs = [out_np] + [torch.empty((H, W, 4), dtype=torch.uint8).pin_memory().numpy() if out_np is not None else None
                               for _ in extra]

        def run_pipelined(n_steps):
            # one host thread per surface (the C ABI is thread-safe per surface): the threads take frames in turn, so
            # frame k + 1 is uploaded, validated and rendered while frame k is still being read back
            errors = []

            def worker(j):
                try:
                    torch.cuda.set_device(local_rank)
                    sf = pair[j]
                    for k in range(j, n_steps, n_flight):
                        if band_e2e:
                            step_e2e_band(sf, shared[j], flight_groups[j])
                            continue
                        sf.begin(True)
                        sf.encode((dl_pinned.data_ptr(), n_dl))
                        sf.flush()
                        if outs[j] is not None:
                            sf.read_pixels_async(outs[j])
                        sf.sync()
                except Exception as e:  # noqa: BLE001
                    errors.append(e)
            ths = [threading.Thread(target=worker, args=(j,)) for j in range(n_flight)]
            for th in ths:
                th.start()
            for th in ths:
                th.join()
            if errors:
                raise errors[0]

        run_pipelined(n_flight)
        barrier()
        n_e2e = max(args.steps, 2 * n_flight)       # every surface renders at least two timed frames
        t0 = time.perf_counter()
        run_pipelined(n_e2e)
        barrier()
        ms_e2e_wall = (time.perf_counter() - t0) * 1e3 / n_e2e
        if outs[0] is not None:
            for o in outs[1:]:      # first rows (rank 0's band) and last rows (the last rank's band)
                if not (np.array_equal(outs[0][:64], o[:64]) and np.array_equal(outs[0][-64:], o[-64:])):
                    raise SystemExit("frames rendered on different surfaces differ")
            if band_e2e and int(outs[0][bands[-1][0]:bands[-1][0] + min(64, bands[-1][1] - bands[-1][0])].astype(np.uint64).sum()) != single_check:
                raise SystemExit("the last band read back end to end differs from the band gathered by NCCL")
        for sb in extra:
            sb.close()
        if shared:
            barrier()
            outs = None
            for sh in shared:
                sh.close()
    clocks = sampler.stop() if rank == 0 else None

    # ---- the complete plug-in path (what a skity::Canvas user pays): CudaContextCreate'd surface -> LockCanvas ->
    # Canvas calls (host encode) -> Flush -> ReadPixels, per step; N = 1 only
    ms_canvas = None
    if world == 1 and partition != "batch" and not args.no_canvas_e2e and args.coverage_mode == "exact":
        try:
            surf.close()      # its arenas: the plug-in's own surface needs the room
            surf = None
            ms_canvas, _ = hostlib.render_scene_cuda_frames(blob, max(2, min(args.steps, 4)), local_rank)
        except Exception as e:  # noqa: BLE001
            ms_canvas = f"failed: {e}"

    h2d_total = int(n_dl) * world
    if world > 1:
        t = torch.tensor([ms_resident, ms_e2e, ms_e2e_wall, ms_e2e_serial, ms_e2e_via_rank0 or 0.0, cull_ms or 0.0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_resident, ms_e2e, ms_e2e_wall, ms_e2e_serial = float(t[0]), float(t[1]), float(t[2]), float(t[3])
        ms_e2e_via_rank0, cull_ms = (float(t[4]) or None), (float(t[5]) or None)
        tb = torch.tensor([int(n_dl)], device="cuda", dtype=torch.int64)
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        h2d_total = int(tb[0])
        ts = torch.tensor(stage_ms, device="cuda")
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        stage_ms = ts.cpu().numpy()
        tl = torch.tensor([launches], device="cuda")
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
        launches = int(tl[0])

    if rank == 0:
        canvases = n_canvases * (world if partition in ("canvas", "batch") else 1)
        mpix = canvases * W * H / 1e6
        paths = n_paths * (world if partition in ("canvas", "batch") else 1)
        value = mpix / (ms_resident / 1e3)
        ms_e2e_used = ms_e2e_wall   # host copies are part of it: wall clock around the K steps, max over ranks
        stage_ms = stage_ms / args.steps
        peak, peak_kind = measured_peak()
        names = device.STAGE_NAMES
        cands = [("k_walk", "sweep: edges -> trapezoid rows", float(stage_ms[2]), st["bytes_walk"]),
                 ("k_cover", "coverage: trapezoid rows -> A8 tile masks", float(stage_ms[3]), st["bytes_cover"])
                 if args.coverage_mode == "exact" else
                 ("k_area_bin+k_area_cover", "AREA coverage: binned lines -> A8 tile masks", float(stage_ms[3]), st["bytes_area"]),
                 ("k_fine", "fine: paint + blend, RGBA8 tiles", float(stage_ms[5]), st["bytes_fine"])]
        if stage_ms[6] > 0 and st.get("bytes_blur"):
            cands.append(("k_blur", "blur: separable StackBlur, 2 passes", float(stage_ms[6]), st["bytes_blur"]))
        dom = max(cands, key=lambda c: c[2])
        achieved = dom[3] / (dom[2] / 1e3) / 1e9 if dom[2] > 0 else 0.0
        wkey = args.workload
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_resident, 4), "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "i32 16.16 fixed point + u8 (fp32 for curve lowering)",
            "data": "synthetic",
            "config": {"workload": desc, "canvases_per_step": canvases, "paths_per_step": paths,
                       "l2": "working set per step (display list, edges, records, A8 masks, canvas) is several GB at N=1, far above the 126 MB L2; no explicit flush",
                       "partition": {"bands": f"tile-row bands of one canvas over {world} GPU(s); every rank holds the part of the display list that can reach "
                                              "its band (skb_display_list_cull_rows on the host, once per scene; the device culls again before flattening); "
                                              "bands stored into rank 0's canvas by the fine pass (NVLink peer memory)" if world > 1 else "single GPU, whole canvas",
                                     "canvas": "one canvas per rank", "batch": "canvases dealt to ranks, one display list per rank"}[partition],
                       "coord_mode": "wide (canvas > 8192 px: the reference's 16.16 conversion without its int32 wrap)" if max(W, H) > 8192 else "reference",
                       "frames_in_flight": 1, "coverage_mode": args.coverage_mode},
            "paths_per_s": round(paths / (ms_resident / 1e3), 1),
            "e2e": {"value": round(mpix / (ms_e2e_used / 1e3), 2), "unit": UNIT, "h2d_bytes_per_step": h2d_total,
                    "d2h_bytes_per_step": int(canvases * W * H * 4), "ms_per_step": round(ms_e2e_used, 4),
                    "what": "C ABI with host buffers: "
                            + ("every rank uploads ITS band's display list from pinned memory (H2D), renders its band and copies it (D2H) into its rows of "
                               "one page-locked host image in shared memory that rank 0 owns — N PCIe links at once, no gather; a host barrier ends the frame; "
                               if band_e2e else ("display list H2D from pinned memory, frame, " + ("bands into rank 0's canvas over NVLink, barrier, " if banded else "")
                                                 + "canvas D2H into pinned memory; "))
                            + ((f"{n_flight} frames in flight on {n_flight} surfaces, one host thread each (upload, validation and read-back of one frame overlap the rendering of the others)") if pipelined else "one frame at a time"),
                    "frames_in_flight": n_flight, "frames_timed": n_e2e if pipelined else args.steps,
                    "one_frame_at_a_time": {"value": round(mpix / (ms_e2e_serial / 1e3), 2), "ms_per_step": round(ms_e2e_serial, 4)}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": f"{dom[0]} ({dom[1]})", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 5), "traffic": NCU_TRAFFIC.get((wkey, dom[0])),
                         "peak_kind": peak_kind, "algorithmic_bytes_per_launch": int(dom[3]), "ms_per_launch": round(dom[2], 4),
                         "per_rank": world > 1},
            "roofline_stages": {c[0]: {"ms": round(c[2], 4), "algorithmic_bytes": int(c[3]),
                                       "achieved_gbs": round(c[3] / (c[2] / 1e3) / 1e9, 2) if c[2] > 0 else None,
                                       "frac": round(c[3] / (c[2] / 1e3) / 1e9 / peak, 5) if c[2] > 0 else None,
                                       "traffic": NCU_TRAFFIC.get((wkey, c[0]))} for c in cands},
            "stages_ms": {names[i]: round(float(stage_ms[i]), 4) for i in range(8)},
            "frame_counters": {k: int(st[k]) for k in ("n_ops", "n_prims", "n_rows", "n_records", "n_items", "n_cmds", "n_retries")},
            "host_encode_ms_outside_timed_region": round(host_encode_s * 1e3, 1),
            "clocks": clocks,
        }
        if banded:
            line["gather"] = {"fused_in_fine_pass": True, "nccl_send_recv_ms": round(nccl_gather_ms, 4),
                              "bytes_into_rank0": int((H - bands[0][1]) * W * 4)}
            if ms_e2e_via_rank0:
              line["e2e"]["via_rank0_canvas"] = {
                "value": round(mpix / (ms_e2e_via_rank0 / 1e3), 2), "ms_per_step": round(ms_e2e_via_rank0, 4),
                "what": "one frame at a time: bands stored into rank 0's canvas by the fine pass (NVLink), barrier, rank 0 reads the whole canvas back"}
            if band_e2e:
                line["e2e"]["host_image_page_locked_on_rank0"] = bool(host_image_locked)
                line["e2e"]["rank0_bound_to_numa_node"] = numa_node
            line["band_display_lists"] = {"bytes_all_ranks": h2d_total, "bytes_whole_list": int(n_dl_full),
                                          "host_cull_ms_outside_timed_region": round(cull_ms or 0.0, 1)}
        if ms_canvas is not None:
            line["e2e_canvas"] = ({"value": round(mpix / (ms_canvas["total"] / 1e3), 2), "unit": UNIT, "ms_per_step": round(ms_canvas["total"], 3),
                                   "ms_canvas_calls_host_encode": round(ms_canvas["canvas_calls"], 3), "ms_flush": round(ms_canvas["flush"], 3),
                                   "ms_read_pixels": round(ms_canvas["read_pixels"], 3),
                                   "what": "skity plug-in API, one GPUContext + GPUSurface, per step: LockCanvas(true) -> the scene's Canvas::DrawPath "
                                           "calls (CudaCanvas encodes on one host thread) -> Flush -> ReadPixels into a Pixmap; wall clock"}
                                  if not isinstance(ms_canvas, str) else {"error": ms_canvas})
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload, blob, W, H)
        print(json.dumps(line), flush=True)
    if surf is not None:
        surf.close()
    dev.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _c4a_window_blobs(blob, size):
    from oracle import windows
    rec = windows.fills_records(blob)
    n = windows.n_windows(size)
    return [windows.window_scene(blob, rec, ix, iy, size) for iy in range(n) for ix in range(n)]


def cpu_baseline(name, blob, W, H):
    """The reference's software backend (oracle/_ref) — or the oracle port where it is not built — on ONE host core
    (SWCanvas is single-threaded), on a bounded sample of the workload."""
    from oracle import refsw, port
    from skity_b200 import hostlib
    use_ref = refsw.available()
    kind = "reference" if use_ref else "port"

    def render_seconds(b):
        if use_ref:
            return refsw.render_scene(b, return_seconds=True)[1]
        d = hostlib.encode_scene(b)
        t0 = time.perf_counter()
        port.render(d)
        return time.perf_counter() - t0

    if W > 8192 or H > 8192:
        # beyond the reference's numeric range (coordinates wrap at 8192 px): one 4096^2 window of the canvas, rendered
        # under a translation with the paths that touch it (oracle/windows.py) = 1/16 of the frame
        sub, _, cnt = _c4a_window_blobs(blob, W)[5]
        sec = render_seconds(sub)
        px = 4096 * 4096
        sample = f"one 4096x4096 window of the canvas (window (1,1) of 4x4, {cnt} paths), draw loop only"
    else:
        sec = render_seconds(blob)
        px = W * H
        sample = "the full scene once (all paths), draw loop only"
    return {"value": round(px / 1e6 / sec, 3), "unit": UNIT, "cores": 1, "kind": kind, "sample": sample, "seconds": round(sec, 3)}


def _ref_worker(q_in, q_out, blobs, fast=False):
    from oracle import refsw
    if fast:
        refsw.use_fast_build()
    refsw.lib()
    while True:
        msg = q_in.get()
        if msg is None:
            break
        sec = 0.0
        for b in blobs:
            sec += refsw.render_scene(b, return_seconds=True)[1]
        q_out.put(sec)


def _band_blob(blob, n_bands, band):
    """The scene under a whole-pixel ClipRect band (the SW fast path, sw_canvas.cc:305-312)."""
    import struct
    from skity_b200 import scene as sc
    magic, ver, w, h, n_ops, _ = struct.unpack_from("<6I", blob, 0)
    y0, y1 = h * band // n_bands, h * (band + 1) // n_bands
    clip = struct.pack("<2I", sc.OP_CLIP_RECT, 20) + struct.pack("<4fI", 0.0, float(y0), float(w), float(y1), 1)
    return struct.pack("<6I", magic, ver, w, h, n_ops + 1, 0) + clip + blob[24:]


def run_reference(args):
    """The reference's own software backend on all host cores.  Canvases within its numeric range: one whole-pixel
    ClipRect band per core.  The 16384^2 canvas: the 16 translated 4096^2 windows (oracle/windows.py) dealt to the cores;
    with fewer than 16 cores a step renders the first `cores` windows and says so (a bounded sample)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import refsw
    sc, desc, partition, scaling = workload(args.workload, 0, 1)
    if partition == "batch":
        sc_list = sc
        W, H = 1920, 1080
    else:
        sc_list = None
        W, H = sc.width, sc.height
    cores = os.cpu_count() or 1
    if not refsw.available():
        from oracle import port
        from skity_b200 import hostlib
        blob = (sc_list[0] if sc_list else sc).encode()
        if W > 8192:
            blob = _c4a_window_blobs(blob, W)[5][0]
            px, sample = 4096 * 4096, "oracle port, one 4096x4096 window per step (compiled reference absent)"
        else:
            px, sample = W * H, "oracle port, whole scene per step (compiled reference absent)"
        dl = hostlib.encode_scene(blob)
        times = []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            port.render(dl)
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
        sec, used, kind = sum(times) / len(times), 1, "port"
        fast_ms = None
    else:
        kind = "reference"
        if sc_list is not None:
            per = [[] for _ in range(cores)]
            for i, s in enumerate(sc_list):
                per[i % cores].append(s.encode())
            per = [p for p in per if p]
            px = len(sc_list) * W * H
            sample = f"each step renders {len(sc_list)} canvases, dealt to {len(per)} processes"
        elif W > 8192 or H > 8192:
            wins = [w[0] for w in _c4a_window_blobs(sc.encode(), W)]
            n_used = len(wins) if cores >= len(wins) else cores
            per = [[] for _ in range(min(cores, n_used))]
            for i in range(n_used):
                per[i % len(per)].append(wins[i])
            px = n_used * 4096 * 4096
            sample = (f"each step renders {n_used} of the canvas's 16 translated 4096x4096 windows (the reference wraps at 8192 px), "
                      f"one process per window, {len(per)} processes")
        else:
            blob = sc.encode()
            per = [[_band_blob(blob, cores, b)] for b in range(cores)]
            px = W * H
            sample = "each step renders the whole scene once, split into one whole-pixel ClipRect band per host core (every process walks the full display list)"
        used = len(per)
        ctx = mp.get_context("fork")

        def timed_steps(fast, warm, steps, budget_s):
            q_out = ctx.Queue()
            qs, procs = [], []
            for blobs in per:
                q = ctx.Queue()
                p = ctx.Process(target=_ref_worker, args=(q, q_out, blobs, fast), daemon=True)
                p.start()
                qs.append(q)
                procs.append(p)
            ts = []
            budget_end = time.perf_counter() + budget_s     # the whole run ends within a few minutes
            for i in range(warm + steps):
                if i >= warm + 1 and time.perf_counter() > budget_end:
                    break
                t0 = time.perf_counter()
                for q in qs:
                    q.put(1)
                for _ in qs:
                    q_out.get()
                if i >= warm:
                    ts.append(time.perf_counter() - t0)
            for q in qs:
                q.put(None)
            for p in procs:
                p.join(timeout=5)
            return ts

        times = timed_steps(False, min(args.warmup, 1), args.steps, 200.0)
        n_timed = len(times)
        # the same steps with the reference compiled -O3 -march=x86-64-v3, one timed step: reported beside
        # the -O2 figure (it renders the same bytes, tests/test_oracle_pinning.py)
        fast_ms = None
        if refsw.fast_available():
            tf = timed_steps(True, 1 if sum(times) < 20.0 else 0, 1, 60.0)
            fast_ms = round(tf[0] * 1e3, 3) if tf else None
        sec = sum(times) / len(times)
        if n_timed < args.steps:
            sample += f"; {n_timed} timed steps fit the time budget"
    value = px / 1e6 / sec
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 3), "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "i32 16.16 fixed point + u8", "data": "synthetic",
            "config": {"workload": desc},
            "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": used, "kind": kind, "sample": sample,
                             "build": "g++ -O2 (x86-64 baseline)" if kind == "reference" else "gcc -O2 port"},
            "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if kind == "reference" and fast_ms:
        line["cpu_baseline"]["O3_march_x86_64_v3"] = {"value": round(px / 1e6 / (fast_ms / 1e3), 3), "unit": UNIT, "ms_per_step": fast_ms,
                                                      "steps": 1}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4a")
    ap.add_argument("--coverage-mode", default="exact", choices=["exact", "area"],
                    help="exact (default): the software backend's coverage, bit for bit — the parity path and the headline; "
                         "area: the coverage-AA tiler's algorithm (north star stages 2-3), exact against its own oracle, NOT the "
                         "software backend's coverage (DESIGN.md 4c)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-canvas-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
