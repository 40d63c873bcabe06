"""Dev build: python tools/dev/build_variant.py <tag> [nvcc flags...] -> fesr_b200/lib/libfesr_<tag>.so (objects in
fesr_b200/build_tr/<tag>/); select it with FESR_LIB_PATH.  Used for same-box A/B timing of kernel variants.
VARIANT_FILES=<prefix>[,<prefix>...] names the source files that are recompiled with the flags (default: layer_fused*);
e.g. VARIANT_FILES=edge_grad python tools/dev/build_variant.py minb6 -DEGL_MINB=6."""
import os, subprocess, sys, concurrent.futures as cf
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fesr_b200 import build as B
tag, flags = sys.argv[1], sys.argv[2:]
OBJ = os.path.join(ROOT, "fesr_b200", "build_tr", tag)
os.makedirs(OBJ, exist_ok=True)
PREFIX = tuple(os.environ.get("VARIANT_FILES", "layer_fused").split(","))      # the variants differ in these files only
only = [f for f in B._sources() if f.startswith(PREFIX)]
def comp(src):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    r = subprocess.run([B.NVCC, *B.FLAGS, *flags, "-c", os.path.join(B.CSRC, src), "-o", obj], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    open(obj + ".log", "w").write(r.stderr)
    return obj
with cf.ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(comp, only))
rest = [os.path.join(B.OBJ, f[:-3] + ".o") for f in B._sources() if f not in only]
lib = os.path.join(B.LIBDIR, f"libfesr_{tag}.so")
r = subprocess.run([B.NVCC, "-shared", "-cudart", "shared", "-o", lib, *objs, *rest, "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"], capture_output=True, text=True)
assert r.returncode == 0, r.stderr
for o in objs:
    for l in open(o + ".log"):
        if "spill" in l and not l.strip().startswith("0 bytes stack"): print(os.path.basename(o), l.strip())
print(lib)
