import sys, time, json; sys.path.insert(0,'.')
import numpy as np
from skity_b200 import scene, hostlib, device
from oracle import port, refsw
dev = device.Device(0)
print('sm', dev.sm_count, device.lib().skb_version_string())
def run(name, s, oracle='port'):
    blob = s.encode()
    dl = hostlib.encode_scene(blob)
    t=time.time(); ref = port.render(dl) if oracle=='port' else refsw.render_scene(blob); tc=time.time()-t
    surf = dev.create_surface(s.width, s.height)
    got = surf.render(dl)
    st = surf.stats()
    got2 = surf.render(dl); st = surf.stats()
    d = np.abs(ref.astype(int)-got.astype(int)).max(axis=2)
    print(name, 'maxdiff', d.max(), 'ndiff', int((d>0).sum()), 'of', d.size, 'cpu %.3fs'%tc, 'gpu ms %.3f'%st['ms_total'], [round(x,3) for x in st['ms_stage']], {k:st[k] for k in ('n_prims','n_rows','n_records','n_items','n_cmds','n_launches','n_retries')}, flush=True)
    if d.max()>0:
        ys,xs=np.nonzero(d); print('   first diffs', list(zip(xs[:8],ys[:8])), ref[ys[0],xs[0]], got[ys[0],xs[0]])
    assert (got==got2).all(), 'non-deterministic'
    surf.close()
run('c0-star', scene.scene_c0(blur=False))
run('c1-200', scene.scene_c1(200, 1024, 1))
run('c0-blur', scene.scene_c0(blur=True))
run('c2-noclip-300', scene.scene_c2(300, 1024, 2, clip_every=0))
run('c3-30', scene.scene_c3(30, 2048, 3))
run('c1-10k', scene.scene_c1(10000, 4096, 1), oracle='port')
