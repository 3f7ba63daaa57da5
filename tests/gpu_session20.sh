#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/s20_tests.log 2>&1
tail -n 3 gpurun_out/s20_tests.log
timeout 600 python - <<'PY'
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from skity_b200 import scene, hostlib, device
from oracle import port
dev = device.Device(0)
for mode in ("carved", "mixed", "refined", "flat"):
    ok = bad = refused_enc = refused_run = 0
    for seed in range(7000, 7150):
        s = scene.scene_difference_clips(seed, mode)
        dl = hostlib.encode_scene(s.encode())
        surf = dev.create_surface(s.width, s.height)
        try:
            surf.begin(True)
            try:
                surf.encode(dl)
            except device.SkbError:
                refused_enc += 1; continue
            try:
                surf.flush(); got = surf.read_pixels()
            except device.SkbError as e:
                refused_run += 1; continue
            if np.array_equal(got, port.render(dl)): ok += 1
            else: bad += 1
        finally:
            surf.close()
    print(mode, "exact", ok, "wrong", bad, "refused at encode", refused_enc, "refused at run time", refused_run, flush=True)
PY
