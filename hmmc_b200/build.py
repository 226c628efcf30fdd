"""Build libhmmc_head.so in-tree with nvcc for sm_100a (no GPU needed: cross-compiles).

    python -m hmmc_b200.build [--force]

The shared library is written next to this file so it travels to the GPU box with the
repo snapshot (it is git-ignored, not gpurun-ignored).
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libhmmc_head.so")
SOURCES = ["core.cu", "pretrain.cu", "finetune.cu", "eval.cu", "eval_fused.cu", "optim.cu", "mlp.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _digest():
    h = hashlib.sha256()
    for root, _, files in sorted(os.walk(CSRC)):
        for f in sorted(files):
            with open(os.path.join(root, f), "rb") as fh:
                h.update(f.encode())
                h.update(fh.read())
    with open(os.path.join(os.path.dirname(HERE), "include", "hmmc_head.h"), "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _run(cmd, log):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), p.stdout[-8000:]))
    return p.stdout


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "digest")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    objs = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        futs = []
        for src in SOURCES:
            obj = os.path.join(OBJ, src.replace(".cu", ".o"))
            objs.append(obj)
            cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
            futs.append(ex.submit(_run, cmd, obj + ".log"))
        for f in futs:
            out = f.result()
            if verbose:
                print(out)
    _run([NVCC, "-shared", "-o", LIB] + objs + ["-lcudart"], os.path.join(OBJ, "link.log"))
    with open(stamp, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
