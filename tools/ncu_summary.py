"""Summarise ncu outputs for profiles/:
    python tools/ncu_summary.py launches <launches.csv>          -> per-kernel share table (markdown)
    python tools/ncu_summary.py report <file.ncu-rep> [...]      -> key metrics per captured launch (markdown)
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max",
]


def launches(path, json_out=None):
    """Per-kernel launch count, time share and (when the capture has them) DRAM bytes."""
    rows = list(csv.reader(open(path)))
    hdr, agg, n = None, collections.OrderedDict(), 0
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        metric = d.get("Metric Name")
        if metric not in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"):
            continue
        name = re.sub(r"<.*", "", d["Kernel Name"]).split("(")[0].replace("void ", "")
        v, unit = float(d["Metric Value"].replace(",", "")), d["Metric Unit"].lower()
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])      # launches, us, bytes read, bytes written
        if metric == "gpu__time_duration.sum":
            v = v / 1e3 if unit.startswith("n") else (v * 1e3 if unit.startswith("m") else v)
            a[0] += 1
            a[1] += v
            n += 1
        else:
            mult = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
            a[2 if metric.endswith("read.sum") else 3] += v * mult
    tot = sum(a[1] for a in agg.values())
    have_bytes = any(a[2] + a[3] > 0 for a in agg.values())
    print("launches captured: %d, sum of durations: %.1f us (cold-cache, serialised: compare SHARES)\n" % (n, tot))
    if have_bytes:
        print("| kernel | launches | total us | share | DRAM read GB | DRAM write GB | GB/s |\n|---|---:|---:|---:|---:|---:|---:|")
    else:
        print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if have_bytes:
            print("| %s | %d | %.1f | %.1f%% | %.2f | %.2f | %.0f |" % (k, a[0], a[1], 100 * a[1] / tot, a[2] / 1e9,
                                                                  a[3] / 1e9, (a[2] + a[3]) / 1e3 / max(a[1], 1e-9)))
        else:
            print("| %s | %d | %.1f | %.1f%% |" % (k, a[0], a[1], 100 * a[1] / tot))
    if json_out:
        import json
        with open(json_out, "w") as f:
            json.dump({k: {"launches": a[0], "us": a[1], "dram_read_bytes": a[2], "dram_write_bytes": a[3]}
                       for k, a in agg.items()}, f, indent=1)


def report(paths):
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        print("### %s\n" % p)
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            print("kernel: `%s`\n" % d.get("Kernel Name", "?")[:150])
            print("| metric | value | unit |\n|---|---:|---|")
            for k in KEYS:
                if k in d:
                    print("| %s | %s | %s |" % (k, d[k], u.get(k, "")))
            print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        report(sys.argv[2:])
