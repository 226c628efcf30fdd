"""Build libhmmc_head variants with different -D switches into hmmc_b200/_variants/<name>.so (measurement aid).
    python tools/build_variants.py name1:"-DX=1 -DY=2" name2:"..."
A tool script selects one by setting hmmc_b200._lib.LIB_PATH before the first load."""
import os
import subprocess
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmmc_b200 import build as B

out_dir = os.path.join(B.HERE, "_variants")
os.makedirs(out_dir, exist_ok=True)
for spec in sys.argv[1:]:
    name, flags = spec.split(":", 1)
    objs = []
    procs = []
    for src in B.SOURCES:
        obj = os.path.join(out_dir, "%s_%s" % (name, src.replace(".cu", ".o")))
        objs.append(obj)
        cmd = [B.NVCC] + [f for f in B.FLAGS if f not in ("-Xptxas", "-v")] + flags.split() + ["-c", os.path.join(B.CSRC, src), "-o", obj]
        procs.append(subprocess.Popen(cmd))
    for p in procs:
        assert p.wait() == 0
    lib = os.path.join(out_dir, name + ".so")
    subprocess.check_call([B.NVCC, "-shared", "-o", lib] + objs + ["-lcudart"])
    for o in objs:
        os.remove(o)
    print("built", lib)
