"""Manual probe (not a test): where the end-to-end frame time of C4a goes — two host threads, one surface each."""
import sys, time, threading
sys.path.insert(0, '.')
import numpy as np, torch
from skity_b200 import scene, hostlib, device
s = scene.scene_c4a()
dl = hostlib.encode_scene(s.encode())
dlp = torch.empty(len(dl), dtype=torch.uint8).pin_memory(); dlp.copy_(torch.frombuffer(bytearray(dl), dtype=torch.uint8)); n_dl = dlp.numel(); del dl
dev = device.Device(0)
NS = int(sys.argv[1]) if len(sys.argv) > 1 else 3
surfs = [dev.create_surface(s.width, s.height) for _ in range(NS)]
outs = [torch.empty((s.height, s.width, 4), dtype=torch.uint8).pin_memory().numpy() for _ in range(NS)]
for sf in surfs:
    sf.begin(True); sf.encode((dlp.data_ptr(), n_dl)); sf.flush(); sf.sync()
def run(n_threads, upload, readback, steps=12):
    def worker(j):
        sf = surfs[j]
        for k in range(j, steps, n_threads):
            sf.begin(True)
            if upload: sf.encode((dlp.data_ptr(), n_dl))
            sf.flush()
            if readback: sf.read_pixels_async(outs[j])
            sf.sync()
    t0 = time.perf_counter()
    ths = [threading.Thread(target=worker, args=(j,)) for j in range(n_threads)]
    [t.start() for t in ths]; [t.join() for t in ths]
    return (time.perf_counter() - t0) * 1e3 / steps
for nt in range(1, NS + 1):
    for up, rb in ((0, 0), (1, 0), (1, 1)):
        run(nt, up, rb, nt)
        print(f"threads {nt} upload {up} readback {rb}: {run(nt, up, rb):.1f} ms/frame", flush=True)
# the pieces alone
sf = surfs[0]
t0 = time.perf_counter(); sf.encode((dlp.data_ptr(), n_dl)); t1 = time.perf_counter(); sf.sync(); t2 = time.perf_counter()
print(f"encode call {1e3*(t1-t0):.1f} ms host, upload done after {1e3*(t2-t0):.1f} ms")
t0 = time.perf_counter(); sf.read_pixels_async(outs[0]); sf.sync(); print(f"read-back alone {1e3*(time.perf_counter()-t0):.1f} ms")
