"""Extract per-launch numbers of the shipped kernels from `ncu --set full` report files:

    python tools/ncu_traffic.py gpurun_out/r2x/*.ncu-rep --json profiles/r2_ncu_traffic.json --md profiles/r2_ncu_kernels.md

Writes (a) the JSON bench.py reads its `roofline.traffic` from (dram__bytes_read.sum + dram__bytes_write.sum per
launch, last captured launch of each kernel) and (b) a markdown table of the counters the judge asked for:
duration, DRAM bytes, L2 bytes, tensor-pipe activity, achieved bandwidth / FLOP rate.  Needs `ncu` on PATH (CPU box)."""
import argparse
import csv
import io
import json
import re
import subprocess

COLS = {
    "time_us": "gpu__time_duration.sum",
    "dram_bytes_read": "dram__bytes_read.sum",
    "dram_bytes_write": "dram__bytes_write.sum",
    "lts_sectors": "lts__t_sectors.sum",
    "tensor_pct_elapsed": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "tensor_pct_active": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm_active_cycles": "sm__cycles_active.avg",
    "elapsed_cycles": "sm__cycles_elapsed.max",
    "regs": "launch__registers_per_thread",
    "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "sm_ghz": "gpc__cycles_elapsed.avg.per_second",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6,
         "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3, "second": 1e6, "Ghz": 1e9, "Mhz": 1e6, "hz": 1.0}


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("hmmc::", "")


def load(path):
    if path.endswith(".csv"):        # `ncu -i report --page raw --csv` exported on the GPU box (reports can exceed the 64 MiB that travel back)
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        d = {"kernel": short(r[idx["Kernel Name"]]), "grid": r[idx["Grid Size"]], "block": r[idx["Block Size"]]}
        for k, col in COLS.items():
            if col not in idx or r[idx[col]] in ("", "n/a"):
                d[k] = None
                continue
            v = float(r[idx[col]].replace(",", ""))
            d[k] = v * SCALE.get(units[idx[col]], 1.0)
        d["lts_bytes"] = d["lts_sectors"] * 32.0 if d.get("lts_sectors") is not None else None
        res.append(d)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+")
    ap.add_argument("--json")
    ap.add_argument("--md")
    a = ap.parse_args()
    kernels, lines = {}, []
    for path in a.reports:
        for d in load(path):
            d["report"] = path
            kernels[d["kernel"]] = d            # the last captured launch of each kernel wins
    for name, d in kernels.items():
        t = d["time_us"] or 0.0
        dram = (d["dram_bytes_read"] or 0) + (d["dram_bytes_write"] or 0)
        lines.append("| `%s` | %s x %s | %.1f | %.1f / %.1f | %.1f | %.0f | %s | %s | %s | %s |" % (
            name[:70], d["grid"], d["block"], t, (d["dram_bytes_read"] or 0) / 1e6, (d["dram_bytes_write"] or 0) / 1e6,
            (d["lts_bytes"] or 0) / 1e6, dram / t / 1e3 if t else 0.0,
            "%.1f" % d["tensor_pct_elapsed"] if d["tensor_pct_elapsed"] is not None else "-",
            "%.0f" % d["regs"] if d["regs"] else "-", "%.1f" % d["issue_pct"] if d["issue_pct"] else "-",
            "%.2f" % (d["sm_ghz"] / 1e9) if d["sm_ghz"] else "-"))
    if a.json:
        json.dump({"source": "ncu --set full --clock-control none, one launch per kernel (cold-ish caches, serialised)",
                   "reports": a.reports,
                   "kernels": {k: {kk: v[kk] for kk in ("time_us", "dram_bytes_read", "dram_bytes_write", "lts_bytes",
                                                        "tensor_pct_elapsed", "grid", "report")} for k, v in kernels.items()}},
                  open(a.json, "w"), indent=1)
    md = ("# ncu --set full, one launch per shipped kernel\n\n"
          "Captured with `tools/r2_profile.sh` (`ncu --set full --import-source on --clock-control none "
          "--kernel-name-base demangled -k regex:hmmc:: python tools/profile_all.py`, after the same program exited 0 "
          "without ncu), exported on the GPU box with `ncu -i ... --page raw --csv` (the report itself exceeds what "
          "travels back; the raw page is committed beside this file) and tabulated by `tools/ncu_traffic.py`. Times are "
          "ncu's: serialised launches, cold-ish caches, no clock control - shares of a step, not bench values. "
          "Pre-train kernels: the last captured launch is the b = 256 loss; eval: config 2 (`eval_rank_kernel<2, 1>`) "
          "and a 100k x 10k slice of config 5; optimizer and EMA: the full 172 M parameters.\n\n"
          "| kernel | grid x block | time us | DRAM read / write MB | L2 MB | DRAM GB/s | tensor pipe % of elapsed | regs | issue % | SM GHz |\n"
          "|---|---|---|---|---|---|---|---|---|---|\n" + "\n".join(lines) + "\n")
    if a.md:
        open(a.md, "w").write(md)
    else:
        print(md)


if __name__ == "__main__":
    main()
