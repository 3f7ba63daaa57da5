"""Manual perf probe (not a test): per-stage device times for a few scene sizes."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from skity_b200 import scene, hostlib, device

dev = device.Device(0)
which = sys.argv[1:] or ['c1', 'p100k', 'c4a']
for name in which:
    if name == 'c1': s = scene.scene_c1()
    elif name == 'p100k': s = scene.scene_random_fills_fast(100000, 8192, 9, box=192.0)
    elif name == 'c4a': s = scene.scene_c4a()
    elif name == 'c4b': s = scene.scene_c4b(0)
    elif name.startswith('c4bbatch'):
        n = int(name[8:] or 64)
        blobs = [scene.scene_c4b(i).encode() for i in range(n)]
        t = time.time(); dl, _ids = hostlib.encode_scene_batch(blobs); te = time.time() - t
        surf = dev.create_surface(16, 16)
        for it in range(3):
            surf.begin(True); surf.encode(dl); surf.flush(); surf.sync(); st = surf.stats()
        print(name, 'encode %.2fs' % te, 'dev %.2f ms' % st['ms_total'], dict(zip(device.STAGE_NAMES, [round(x, 3) for x in st['ms_stage']])), 'canvases/s %.0f' % (n / st['ms_total'] * 1e3), 'Mpix/s %.0f' % (n * 1920 * 1080 / 1e3 / st['ms_total']), 'paths/s %.0f' % (n * 1000 / st['ms_total'] * 1e3), flush=True)
        surf.close(); continue
    elif name == 'c3': s = scene.scene_c3()
    elif name == 'c3s': s = scene.scene_c3(200, 4096, 3)
    elif name == 'c2': s = scene.scene_c2(20000, 4096, 2, clip_every=0)
    elif name == 'c2clip': s = scene.scene_c2(20000, 4096, 2)
    elif name == 'c2clip2k': s = scene.scene_c2(2000, 4096, 2)
    t = time.time(); dl = hostlib.encode_scene(s.encode()); te = time.time() - t
    surf = dev.create_surface(s.width, s.height)
    for it in range(3):
        t = time.time(); surf.begin(True); surf.encode(dl); surf.flush(); surf.sync(); tw = time.time() - t
        st = surf.stats()
    print(name, 'encode %.2fs' % te, 'wall %.1f ms' % (tw * 1e3), 'dev %.2f ms' % st['ms_total'],
          dict(zip(device.STAGE_NAMES, [round(x, 3) for x in st['ms_stage']])),
          {k: st[k] for k in ('n_prims', 'n_rows', 'n_records', 'n_items', 'n_cmds', 'n_launches', 'n_retries', 'n_rw_retried', 'n_rw_sequential')},
          'Mpix/s %.0f' % (s.width * s.height / 1e3 / st['ms_total']), 'paths/s %.0f' % (s.n_draws / st['ms_total'] * 1e3), flush=True)
    surf.close()
