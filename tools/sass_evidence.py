"""cuobjdump -sass excerpts that show which hardware paths the shipped kernels use (run on the CPU box):

    python tools/sass_evidence.py > profiles/r2_sass_evidence.md

Per contraction kernel of hmmc_b200/libhmmc_head.so: counts of the tcgen05 / TMA / TMEM instructions
(UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA load / store, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
UTCATOMSWS / UTCALLOC = TMEM allocation) and the first occurrence of each with its neighbours."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hmmc_b200", "libhmmc_head.so")
KEYS = ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "UTMACCTL", "UTCATOMSWS", "ACQBULK", "MUFU.EX2", "F2FP",
        "SYNCS", "UCGABAR", "UTMACMDFLUSH", "ELECT", "ACQSHMINIT")
WANT = ("umma_gemm_pair_kernel", "umma_gemm_kernel", "eval_rank_pair_kernel", "eval_rank_kernel")

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs, name = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        funcs[name] = []
        continue
    if name and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
        funcs[name].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).rstrip())
dem = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
sha = subprocess.run(["sha256sum", LIB], capture_output=True, text=True).stdout.split()[0]
print("# SASS evidence (round 2)\n")
print("`cuobjdump -sass hmmc_b200/libhmmc_head.so` (sha256 `%s`), built by `python -m hmmc_b200.build` with "
      "`-gencode arch=compute_100a,code=sm_100a`.\n" % sha[:16])
print("| kernel | instructions | " + " | ".join(KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
rows = []
for (mangled, body), d in zip(funcs.items(), dem):
    if not any(w in d for w in WANT):
        continue
    short = re.sub(r"\(.*$", "", d.replace("hmmc::", "").replace("void ", ""))
    cnt = [sum(1 for l in body if k in l) for k in KEYS]
    print("| `%s` | %d | %s |" % (short, len(body), " | ".join(str(c) for c in cnt)))
    rows.append((short, body))
print("\n## First occurrences (two lines of context)\n")
for short, body in rows:
    if "pair_kernel<hmmc::EpiInfoNCE<1" not in short.replace("(int)", "") and "eval_rank_pair_kernel<3" not in short.replace("(int)", ""):
        continue
    print("### `%s`\n\n```" % short)
    for k in ("UTMALDG", "UTCHMMA", "UTCBAR", "LDTM", "MUFU.EX2", "UTMASTG"):
        for i, l in enumerate(body):
            if k in l:
                for x in body[max(0, i - 1):i + 2]:
                    print(x.strip()[:120])
                print("...")
                break
    print("```\n")
