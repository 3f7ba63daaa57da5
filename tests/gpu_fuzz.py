"""Manual fuzz (not collected by pytest): random scenes of every feature class rendered through the C ABI
on the GPU and compared with the oracle port.  Usage: python tests/gpu_fuzz.py [n_rounds] [seed0]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from skity_b200 import scene, hostlib, device
from skity_b200.scene import Scene, Paint, PathData
from oracle import port

n_rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
dev = device.Device(0)
bad = 0
t0 = time.time()


rand_scene = scene.scene_fuzz


for r in range(n_rounds):
    seed = seed0 + r
    s, exact = rand_scene(seed)
    dl = hostlib.encode_scene(s.encode())
    want = port.render(dl)
    surf = dev.create_surface(s.width, s.height)
    try:
        got = surf.render(dl)
    except device.SkbError as e:
        print('seed', seed, 'kind', seed % 10, 'ERROR', e)
        bad += 1
        surf.close()
        continue
    surf.close()
    d = np.abs(got.astype(int) - want.astype(int)).max(axis=2)
    ok = d.max() == 0 if exact else (d.max() <= 2 and (d <= 1).mean() >= 0.999)
    if not ok:
        bad += 1
        ys, xs = np.nonzero(d)
        print('seed', seed, 'kind', seed % 10, s.width, s.height, 'MISMATCH max', d.max(), 'n', len(ys), list(zip(xs[:4], ys[:4])))
print(f'fuzz: {n_rounds} scenes, {bad} bad, {time.time() - t0:.1f}s')
