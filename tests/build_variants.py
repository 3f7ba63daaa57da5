"""Manual tool: builds variant CUDA libraries (one nvcc call each, -D flags per variant) into the untracked
gpurun_variants/<name>.so, to be compared on the GPU with SKB_LIB=... python tests/perf_probe.py.
Usage: python tests/build_variants.py base="" f16="-DFINE_MINB=16" ..."""
import subprocess, sys, os, concurrent.futures as cf
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skity_b200 import build as b
os.makedirs(os.path.join(b.REPO, 'gpurun_variants'), exist_ok=True)
V = dict(a.split('=', 1) for a in sys.argv[1:])   # name="-DX=1 -DY=2"
def one(kv):
    name, flags = kv
    out = os.path.join(b.REPO, 'gpurun_variants', name + '.so')
    cmd = ["nvcc", *b.NVCC_FLAGS, *flags.split(), f"-I{b.REPO}", f"-I{b.REPO}/include", "-shared", "-o", out,
           *b.cuda_sources(), "-cudart", "static", "-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode: print(name, 'FAILED', r.stderr[-2000:]); return
    lines = r.stderr.splitlines()
    for i, l in enumerate(lines):
        if 'Compiling entry function' in l and any(k in l for k in ('k_walkE', 'k_coverE', 'k_fineE')):
            print(name, l.split("'")[1][:24], ' '.join(lines[i+1:i+3])[:200])
with cf.ThreadPoolExecutor(8) as ex: list(ex.map(one, V.items()))
