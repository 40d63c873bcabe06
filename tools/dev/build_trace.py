"""Dev build: libfesr_tr.so = libfesr.so compiled with -DFL_TRACE (per-role clock64 accounting in the fused layer kernels)."""
import os, subprocess, sys, concurrent.futures as cf
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fesr_b200 import build as B
OBJ = os.path.join(ROOT, "fesr_b200", "build_tr")
os.makedirs(OBJ, exist_ok=True)
def comp(src):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    srcp = os.path.join(B.CSRC, src)
    if os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(os.path.join(B.CSRC, f)) for f in os.listdir(B.CSRC)):
        return obj
    r = subprocess.run([B.NVCC, *B.FLAGS, "-DFL_TRACE", *os.environ.get("FESR_DEV_FLAGS", "").split(), "-c", srcp, "-o", obj], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return obj
with cf.ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(comp, B._sources()))
lib = os.path.join(B.LIBDIR, "libfesr_tr.so")
r = subprocess.run([B.NVCC, "-shared", "-cudart", "shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"], capture_output=True, text=True)
assert r.returncode == 0, r.stderr
print(lib)
