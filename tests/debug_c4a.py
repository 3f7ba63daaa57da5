"""Debug aid (gpurun): find single paths of C4a near the canvas edge where device != wide-mode port."""
import os, sys, struct
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from oracle import port, windows
from skity_b200 import device, hostlib, scene

sc = scene.scene_c4a()
blob = sc.encode()
off, ln, bbox = windows.fills_records(blob)
sel = np.nonzero((bbox[:, 2] > 16256) & (bbox[:, 1] < 128))[0]
print("candidates", len(sel))
dev = device.Device(0)
surf = dev.create_surface(16384, 16384)
found = 0
for i in sel:
    sub = struct.pack("<6I", scene.MAGIC, 1, 16384, 16384, 1, 0) + blob[off[i]:off[i] + ln[i]]
    dl = hostlib.encode_scene(sub)
    surf.begin(True); surf.encode(dl); surf.flush()
    got = surf.read_pixels(16384 - 512, 0, 512, 256)
    port.set_row_band(0, 256)
    want = port.render(dl)[0:256, 16384 - 512:]
    port.set_row_band(0, 0)
    if not np.array_equal(got, want):
        d = (got != want).any(axis=2)
        py, px = np.nonzero(d)
        print("path", int(i), "bbox", bbox[i], "differs at", len(py), "px; first", px[0] + 16384 - 512, py[0], got[py[0], px[0]], want[py[0], px[0]])
        print("record hex", blob[off[i]:off[i] + ln[i]].hex())
        found += 1
        if found >= 3:
            break
print("found", found)
