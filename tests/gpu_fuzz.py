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


def rand_scene(seed):
    rng = np.random.RandomState(seed)
    kind = seed % 10
    w, h = int(rng.randint(40, 700)), int(rng.randint(40, 700))
    if kind == 0:
        return scene.scene_random_fills(int(rng.randint(5, 150)), 0, seed, box=float(rng.uniform(20, 500)), width=w, height=h), True
    if kind == 1:
        return scene.scene_c2(int(rng.randint(5, 80)), max(w, h), seed, clip_every=0), False
    if kind == 2:
        return scene.scene_c2(int(rng.randint(10, 80)), max(w, h), seed, clip_every=int(rng.randint(4, 20)),
                              clip_box=float(rng.uniform(60, 400))), False
    if kind == 3:
        return scene.scene_c3(int(rng.randint(1, 8)), max(w, h), seed, box=float(rng.uniform(30, 200))), True
    if kind == 4:  # transforms + conics + tiny/huge shapes
        s = Scene(w, h)
        for i in range(int(rng.randint(3, 40))):
            s.save()
            s.translate(float(rng.uniform(0, w)), float(rng.uniform(0, h)))
            s.rotate(float(rng.uniform(0, 360)))
            s.scale(float(rng.uniform(0.2, 3)), float(rng.uniform(0.2, 3)))
            p = PathData(int(rng.randint(0, 2)))
            r = float(rng.uniform(0.3, 120))
            p.move_to(r, 0).conic_to(r, r, 0, r, 0.7071).conic_to(-r, r, -r, 0, 0.7071)
            p.conic_to(-r, -r, 0, -r, float(rng.uniform(0.1, 3))).conic_to(r, -r, r, 0, 0.7071).close()
            if rng.uniform() < 0.5:
                p.move_to(-r / 3, -r / 3).line_to(r / 3, -r / 3).line_to(0, r / 2).close()
            col = tuple(np.float32(v) for v in rng.uniform(0, 1, 4))
            style = int(rng.randint(0, 4))
            s.draw_path(p, Paint(style=style, fill=col, stroke=col[::-1], stroke_width=float(rng.uniform(0.2, 12)),
                                 cap=int(rng.randint(0, 3)), join=int(rng.randint(0, 3))))
            s.restore()
        return s, True
    if kind == 6:
        return scene.scene_blend_modes(seed, size=int(rng.randint(160, 600))), False
    if kind == 7:
        return scene.scene_filters(seed, size=int(rng.randint(200, 600))), True
    if kind == 8:
        return scene.scene_layers(seed, size=int(rng.randint(300, 640))), True
    if kind == 9:
        return scene.scene_conical(seed, size=int(rng.randint(200, 600))), False
    s = Scene(w, h)  # solid draws under nested clips and rect clips
    depth = 0
    for i in range(int(rng.randint(5, 60))):
        if rng.uniform() < 0.2 and depth < 3:
            s.save(); depth += 1
            if rng.uniform() < 0.5:
                s.clip_path(scene._random_closed_path(rng, rng.uniform(0, w), rng.uniform(0, h), float(rng.uniform(50, 400)), int(rng.randint(0, 6))), True)
            else:
                x, y = rng.uniform(0, w), rng.uniform(0, h)
                s.clip_rect(float(x), float(y), float(x + rng.uniform(10, 300)), float(y + rng.uniform(10, 300)), True)
        elif rng.uniform() < 0.1 and depth > 0:
            s.restore(); depth -= 1
        col = tuple(np.float32(v) for v in rng.uniform(0, 1, 4))
        s.draw_path(scene._random_closed_path(rng, rng.uniform(0, w), rng.uniform(0, h), float(rng.uniform(20, 400)), i),
                    Paint(style=int(rng.randint(0, 3)), fill=col, stroke=col, stroke_width=float(rng.uniform(0.5, 9))))
    while depth:
        s.restore(); depth -= 1
    return s, True


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
