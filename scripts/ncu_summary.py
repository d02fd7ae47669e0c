#!/usr/bin/env python
"""Summarise ncu output into the small text files committed under profiles/.

  python scripts/ncu_summary.py launches gpurun_out/X_launches.csv [--skip-steps N --kernels-per-step K] > profiles/...
  python scripts/ncu_summary.py full gpurun_out/X_full.ncu-rep > profiles/...

`launches`: per-kernel share of the summed device time of every captured launch (cold-cache, serialised: shares, not
absolutes).  `full`: one line per captured launch with duration, DRAM bytes, DRAM / L2 / tensor-pipe utilisation.
"""
import collections
import csv
import subprocess
import sys


def short(name):
    name = name.replace("void ", "").replace("ae::", "")
    cut = name.find("(")
    return (name[:cut] if cut > 0 else name)[:70]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    ours = 0
    for r in data:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        k = short(r[ki])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        if "ae::" in r[ki]:
            ours += 1
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {len(data)} launches ({ours} of this library), summed gpu__time_duration {tot / 1e3:.1f} us")
    print(f"{'share':>7} {'count':>6} {'avg us':>9}  kernel")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{a[1] / tot * 100:6.2f}% {a[0]:6d} {a[1] / a[0] / 1e3:9.2f}  {k}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, name, scale_to=None):
        i = col.get(name)
        if i is None or r[i] == "":
            return float("nan")
        v = float(r[i].replace(",", ""))
        u = units[i]
        if scale_to == "MB":
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
        if scale_to == "us":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        return v

    print(f"# {path}: ncu --set full, {len(data)} launches")
    print(f"{'us':>8} {'dramRdMB':>9} {'dramWrMB':>9} {'dram%':>6} {'L2%':>6} {'tensor%':>8} {'sm%':>6} {'occ%':>6} {'regs':>5} {'grid':>7}  kernel")
    for r in data:
        print(f"{get(r, 'gpu__time_duration.sum', 'us'):8.2f} {get(r, 'dram__bytes_read.sum', 'MB'):9.2f} "
              f"{get(r, 'dram__bytes_write.sum', 'MB'):9.2f} {get(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
              f"{get(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
              f"{get(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):8.2f} "
              f"{get(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
              f"{get(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} "
              f"{get(r, 'launch__registers_per_thread'):5.0f} {get(r, 'launch__grid_size'):7.0f}  {short(r[col['Kernel Name']])}")


def traffic(path, out_json, precision="fp32"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the captured launches) -> profiles/r2_roofline_traffic.json"""
    import json
    import os
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = []
    for r in data:
        b = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = col[name]
            b += float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
        tot.append(b)
    d = json.load(open(out_json)) if os.path.exists(out_json) else {}
    d[precision] = sum(tot) / len(tot)
    d[precision + "_source"] = f"{os.path.basename(path)}: ncu --set full, {len(tot)} launches of {short(data[0][col['Kernel Name']])}"
    json.dump(d, open(out_json, "w"), indent=1)
    print(d)


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "fp32")
        sys.exit(0)
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
