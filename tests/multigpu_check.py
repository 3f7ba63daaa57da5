"""Manual multi-GPU check (run under torchrun on N GPUs of one box):
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py [c4a]
Band-split rendering of one canvas + NCCL gather to rank 0, compared with a single-GPU render."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from skity_b200 import device, hostlib, multigpu, scene

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
big = len(sys.argv) > 1 and sys.argv[1] == "c4a"
s = scene.scene_c4a() if big else scene.scene_random_fills_fast(20000, 4096, 4, box=160.0)
W, H = s.width, s.height
dl = hostlib.encode_scene(s.encode())
dev = device.Device(local)
surf = dev.create_surface(W, H)
bands = multigpu.band_ranges(H, world)
y0, y1 = bands[rank]
surf.set_band(y0, y1)
stream = torch.cuda.ExternalStream(surf.stream(), device=torch.device("cuda", local))
times = []
for it in range(4):
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    surf.begin(True)
    surf.encode(dl) if it == 0 else None
    surf.flush()
    e1.record(stream)
    surf.sync()
    times.append(e0.elapsed_time(e1))
t = torch.tensor([min(times[1:])], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
render_ms = float(t[0])


def band_tensor(a, b):
    return multigpu.surface_band_tensor(surf, a, b)[0]


gather = []
for it in range(3):
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    multigpu.gather_bands(band_tensor, bands, rank, world, dist)
    torch.cuda.synchronize()
    dist.barrier()
    gather.append((time.perf_counter() - t0) * 1e3)
if rank == 0:
    got = surf.read_pixels()
    full = dev.create_surface(W, H)
    want = full.render(dl)
    ok = bool(np.array_equal(got, want))
    print(f"world={world} {W}x{H} paths={s.n_draws} band-render max {render_ms:.2f} ms "
          f"({W * H / 1e3 / render_ms:.0f} Mpix/s, {s.n_draws / render_ms * 1e3:.0f} paths/s) "
          f"gather {min(gather):.2f} ms ({(H - bands[0][1]) * surf.device_ptr()[1] / 1e6 / min(gather):.1f} GB/s into rank 0) "
          f"bands==single-GPU: {ok}", flush=True)
    assert ok
# ---- the same split with the gather fused into the fine pass (peer-memory stores, no copy)
dist.barrier()
surf.begin(True)            # rank 0's canvas is cleared by rank 0; the others' clears are local
surf.sync()
dist.barrier()
multigpu.fuse_gather_into_fine_pass(surf, rank, dist)
fused = []
for it in range(4):
    surf.begin(True)
    surf.sync()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    surf.flush()
    surf.sync()
    dist.barrier()          # all bands are in rank 0's canvas
    fused.append((time.perf_counter() - t0) * 1e3)
if rank == 0:
    got2 = surf.read_pixels()
    ok2 = bool(np.array_equal(got2, want))
    print(f"fused gather: frame + barrier {min(fused[1:]):.2f} ms wall (separate: render {render_ms:.2f} + gather {min(gather):.2f}), "
          f"canvas==single-GPU: {ok2}", flush=True)
    assert ok2
if rank != 0:
    surf.set_remote_canvas(None)
dist.barrier()
dist.destroy_process_group()
