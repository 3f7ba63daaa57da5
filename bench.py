#!/usr/bin/env python3
"""bench.py — rasterisation throughput of the B200 backend on BASELINE.json's workload.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU backend on the host cores

A "step" is one frame: clear the canvas and rasterise the whole synthetic scene (flatten, walk,
coverage, bin, fine [, blur]).  Workload at N=1: BASELINE.json configs[1] — 10k random quad/cubic
paths, mixed nonzero/even-odd, solid fill, 4096x4096 (seed 1).  At N>1 every rank renders its own
canvas of that workload (seed 1+rank): the "batch of independent canvases split by canvas"
partition, no data-path collective, weak scaling.

  value  Mpix/s of canvas filled, whole job, display list already resident in HBM, timed with CUDA
         events on the surface's stream over exactly K steps, max over ranks.
  e2e    same metric through the C ABI with HOST buffers: every step copies the display list from
         pinned host memory (H2D) and reads the finished canvas back (D2H) inside the timed region.
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

# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures of this
# workload (profiles/r01_ncu_top3_c1.txt); None where no capture exists
NCU_TRAFFIC = {("c1", "k_walk"): 29028352 + 112788480, ("c1", "k_cover"): 146949632 + 124314624,
               ("c1", "k_fine"): 235782912 + 49222656}

E2E_LANES = int(os.environ.get("SKB_BENCH_LANES", "4"))      # host threads the end-to-end loop drives frames with
E2E_SURFACES = int(os.environ.get("SKB_BENCH_SURFACES_PER_LANE", "1"))  # surfaces a thread alternates between: the
                                                             # read-back of its last frame overlaps its next frame
METRIC = "canvas_mpix_per_s"
UNIT = "Mpix/s"


def workload(name, rank):
    from skity_b200 import scene
    if name == "c1":
        return scene.scene_c1(10000, 4096, 1 + rank), "c1: 10k random quad/cubic paths, mixed nonzero/even-odd, solid fill, 4096x4096"
    if name == "c1-small":
        return scene.scene_c1(1000, 1024, 1 + rank), "c1-small: 1k paths 1024x1024 (smoke size)"
    if name == "c3":
        return scene.scene_c3(2000, 8192, 3 + rank), "c3: 2k blurred paths (sigma 4-64) 8192x8192"
    if name == "c2":
        return scene.scene_c2(20000, 4096, 2 + rank, clip_every=0), "c2 (no clip stack): 20k stroked+filled gradient paths 4096x4096"
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


def run_ours(args):
    import torch
    import torch.distributed as dist
    from skity_b200 import device, hostlib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    sc, desc = workload(args.workload, rank)
    W, H = sc.width, sc.height
    n_paths = sc.n_draws
    blob = sc.encode()
    dl = hostlib.encode_scene(blob)                 # host-side encode (CudaCanvas), outside every timed region
    dl_pinned = torch.empty(len(dl), dtype=torch.uint8).pin_memory()
    dl_pinned.copy_(torch.frombuffer(bytearray(dl), dtype=torch.uint8))
    out_pinned = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
    out_np = out_pinned.numpy()

    dev = device.Device(local_rank)
    surf = dev.create_surface(W, H)
    stream = torch.cuda.ExternalStream(surf.stream(), device=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        surf.begin(True)
        surf.flush()

    def step_e2e():
        surf.begin(True)
        surf.encode((dl_pinned.data_ptr(), len(dl)))
        surf.flush()
        surf.read_pixels(out=out_np)

    # ---- device-resident throughput
    surf.begin(True)
    surf.encode((dl_pinned.data_ptr(), len(dl)))
    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = np.zeros(8)
    launches = 0
    bytes_fine = bytes_cover = bytes_walk = 0
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
        st = surf.stats()      # waits for the frame; per-stage CUDA-event timings of this step
        stage_ms += np.array(st["ms_stage"])
        launches += st["n_launches"]
        bytes_fine, bytes_cover, bytes_walk = st["bytes_fine"], st["bytes_cover"], st["bytes_walk"]
    e1.record(stream)
    barrier()
    ms_resident = e0.elapsed_time(e1) / args.steps

    # ---- end to end through the C ABI with host buffers, one frame at a time
    for _ in range(2):
        step_e2e()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for _ in range(args.steps):
        step_e2e()
    f1.record(stream)
    barrier()
    ms_e2e_serial = f0.elapsed_time(f1) / args.steps

    # ---- end to end with several frames in flight: every step still uploads its display list from pinned
    # memory, renders, and reads its canvas back to pinned memory, but steps are dealt round-robin to E2E_LANES
    # surfaces, each driven by its own host thread on its own stream, so that the read-back of one frame
    # overlaps the rendering of others and the latency-bound sweep of one frame shares the SMs with the
    # coverage / fine passes of another — what an application streaming frames through the backend does
    import threading
    lanes = [(surf, out_np)]
    for _ in range(E2E_LANES * E2E_SURFACES - 1):
        sf = dev.create_surface(W, H)
        lanes.append((sf, torch.empty((H, W, 4), dtype=torch.uint8).pin_memory().numpy()))
    lane_streams = [torch.cuda.ExternalStream(sf.stream(), device=torch.device("cuda", local_rank)) for sf, _ in lanes]

    def run_e2e_lanes(n_steps, host_buffers=True):
        def work(t):
            mine = lanes[t * E2E_SURFACES:(t + 1) * E2E_SURFACES]
            for k, _ in enumerate(range(t, n_steps, E2E_LANES)):
                sf, out = mine[k % E2E_SURFACES]
                sf.begin(True)
                if host_buffers:
                    sf.encode((dl_pinned.data_ptr(), len(dl)))
                sf.flush()
                if host_buffers:
                    sf.read_pixels_async(out)   # stream-ordered: the next begin() on this surface waits for it
        threads = [threading.Thread(target=work, args=(t,)) for t in range(E2E_LANES)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()

    run_e2e_lanes(2 * E2E_LANES * E2E_SURFACES)
    for sf, _ in lanes:
        sf.sync()
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)                               # every stream is idle here
    run_e2e_lanes(args.steps)
    for st_ in lane_streams[1:]:                    # p1 after the last frame of every lane
        ev = torch.cuda.Event()
        ev.record(st_)
        stream.wait_event(ev)
    p1.record(stream)
    barrier()
    ms_e2e = p0.elapsed_time(p1) / args.steps
    # the same lanes with the display list resident and no read-back (for comparison with `value`, which is
    # measured one frame at a time so that its per-stage timings mean something)
    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for sf, _ in lanes:
        sf.sync()
    q0.record(stream)
    run_e2e_lanes(args.steps, host_buffers=False)
    for st_ in lane_streams[1:]:
        ev = torch.cuda.Event()
        ev.record(st_)
        stream.wait_event(ev)
    q1.record(stream)
    barrier()
    ms_resident_lanes = q0.elapsed_time(q1) / args.steps
    for sf, out in lanes[1:]:
        if not np.array_equal(out, lanes[0][1]):
            raise SystemExit("frames rendered on different surfaces differ")
        sf.close()
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        t = torch.tensor([ms_resident, ms_e2e, ms_e2e_serial, ms_resident_lanes], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_resident, ms_e2e, ms_e2e_serial, ms_resident_lanes = float(t[0]), float(t[1]), float(t[2]), float(t[3])

    if rank == 0:
        mpix = W * H / 1e6
        value = world * mpix / (ms_resident / 1e3)
        e2e_value = world * mpix / (ms_e2e / 1e3)
        stage_ms /= args.steps
        peak, peak_kind = measured_peak()
        names = device.STAGE_NAMES
        # dominant kernel = the longest of the three data-facing stages, each a single kernel:
        # k_walk (sweep), k_cover (coverage), k_fine (paint + blend)
        cands = [("k_walk (edge sweep)", float(stage_ms[2]), bytes_walk),
                 ("k_cover (coverage)", float(stage_ms[3]), bytes_cover),
                 ("k_fine (paint+blend)", float(stage_ms[5]), bytes_fine)]
        dom, dom_ms, dom_bytes = max(cands, key=lambda c: c[1])
        achieved = dom_bytes / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_resident, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "i32 16.16 fixed point + u8 (fp32 for curve lowering)",
            "data": "synthetic",
            "config": {"workload": desc, "canvases_per_step": world, "paths_per_canvas": n_paths,
                       "l2": "working set per step (records + A8 masks + canvas, ~0.4 GB) exceeds the 126 MB L2; no explicit flush",
                       "partition": "by canvas" if world > 1 else "single",
                       "frames_in_flight": "value: 1 (so that the per-stage timings are those of a frame); e2e: %d host threads x %d surfaces" % (E2E_LANES, E2E_SURFACES)},
            "resident_frames_in_flight": {"frames_in_flight": E2E_LANES, "ms_per_step": round(ms_resident_lanes, 4),
                                          "value": round(world * W * H / 1e6 / (ms_resident_lanes / 1e3), 2)},
            "paths_per_s": round(world * n_paths / (ms_resident / 1e3), 1),
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": len(dl),
                    "d2h_bytes_per_step": W * H * 4, "ms_per_step": round(ms_e2e, 4), "frames_in_flight": E2E_LANES,
                    "host_threads": E2E_LANES, "surfaces_per_thread": E2E_SURFACES,
                    "one_frame_at_a_time": {"value": round(world * mpix / (ms_e2e_serial / 1e3), 2),
                                            "ms_per_step": round(ms_e2e_serial, 4)}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 5), "traffic": NCU_TRAFFIC.get((args.workload, dom.split()[0])),
                         "peak_kind": peak_kind,
                         "algorithmic_bytes_per_launch": int(dom_bytes), "ms_per_launch": round(dom_ms, 4)},
            "stages_ms": {names[i]: round(float(stage_ms[i]), 4) for i in range(8)},
            "stage_frac_of_hbm_roofline": {n: round(b / (m / 1e3) / 1e9 / peak, 5) if m > 0 else None for n, m, b in cands},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(blob, dl, W, H)
        print(json.dumps(line), flush=True)
    surf.close()
    dev.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(blob, dl, W, H):
    """The reference's software backend (oracle/_ref) — or the oracle port where it is not built —
    timed on one host core on the same scene (SWCanvas is single-threaded)."""
    from oracle import refsw, port
    if refsw.available():
        _, sec = refsw.render_scene(blob, return_seconds=True)
        kind = "reference"
    else:
        t0 = time.perf_counter()
        port.render(dl)
        sec = time.perf_counter() - t0
        kind = "port"
    return {"value": round(W * H / 1e6 / sec, 3), "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "the full scene once (all paths), draw loop only", "seconds": round(sec, 3)}


def _ref_band_worker(q_in, q_out, blob, n_bands, band):
    """Renders the scene under a whole-pixel ClipRect band (the SW fast path, sw_canvas.cc:305-312)."""
    import struct
    from oracle import refsw
    from skity_b200 import scene as sc
    magic, ver, w, h, n_ops, _ = struct.unpack_from("<6I", blob, 0)
    y0, y1 = h * band // n_bands, h * (band + 1) // n_bands
    clip = struct.pack("<2I", sc.OP_CLIP_RECT, 20) + struct.pack("<4fI", 0.0, float(y0), float(w), float(y1), 1)
    banded = struct.pack("<6I", magic, ver, w, h, n_ops + 1, 0) + clip + blob[24:]
    refsw.lib()
    while True:
        msg = q_in.get()
        if msg is None:
            break
        _, sec = refsw.render_scene(banded, return_seconds=True)
        q_out.put(sec)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import refsw
    sc, desc = workload(args.workload, 0)
    W, H = sc.width, sc.height
    blob = sc.encode()
    if not refsw.available():
        # the compiled reference is absent: fall back to the oracle port, single core
        from oracle import port
        from skity_b200 import hostlib
        dl = hostlib.encode_scene(blob)
        times = []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            port.render(dl)
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
        sec, cores, kind = sum(times) / len(times), 1, "port"
    else:
        cores = os.cpu_count() or 1
        ctx = mp.get_context("fork")
        q_out = ctx.Queue()
        qs, procs = [], []
        for b in range(cores):
            q = ctx.Queue()
            p = ctx.Process(target=_ref_band_worker, args=(q, q_out, blob, cores, b), daemon=True)
            p.start()
            qs.append(q)
            procs.append(p)
        times = []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            for q in qs:
                q.put(1)
            for _ in qs:
                q_out.get()
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
        for q in qs:
            q.put(None)
        for p in procs:
            p.join(timeout=5)
        sec, kind = sum(times) / len(times), "reference"
    value = W * H / 1e6 / sec
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "i32 16.16 fixed point + u8", "data": "synthetic",
            "config": {"workload": desc, "canvases_per_step": 1},
            "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": "each step renders the whole scene once, split into one whole-pixel ClipRect band per host core "
                                       "(every process walks the full display list)"},
            "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c1")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
