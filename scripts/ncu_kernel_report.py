#!/usr/bin/env python
"""Per-kernel digest of an `ncu --set full --import-source on` report: pipe utilisation, top stall reasons, the
instructions that collect the most stall samples and the opcode mix.   python scripts/ncu_kernel_report.py X.ncu-rep [N]"""
import csv
import subprocess
import sys


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main(path, top=12):
    rows = list(csv.reader(run([path, "--page", "raw", "--csv"]).splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size"]
    seen = {}
    for idx, r in enumerate(rows[2:]):
        name = r[col["Kernel Name"]]
        if name in seen:
            continue
        seen[name] = idx
        print("==", name[:100])
        for w in want:
            if w in col and r[col[w]] != "":
                print(f"   {w:85s} {units[col[w]]:14s} {r[col[w]]}")
        st = [(h, float(r[i].replace(",", ""))) for h, i in col.items()
              if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
        print("   stalls per issue:", ", ".join(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}"
                                                for h, v in sorted(st, key=lambda x: -x[1])[:7]))
        src = list(csv.reader(run([path, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"]).splitlines()))
        if len(src) < 3:
            continue
        sh = src[1]
        ci = {h: i for i, h in enumerate(sh)}
        data = []
        for x in src[2:]:
            if x and x[0] == "Kernel Name":                 # the next launch's section
                break
            if len(x) >= len(sh) - 2 and x[0] != "Address":
                data.append(x)
        tot = sum(int(x[ci["# Samples"]]) for x in data) or 1
        print(f"   {len(data)} instructions, {tot} samples; most-sampled:")
        for x in sorted(data, key=lambda x: -int(x[ci["# Samples"]]))[:top]:
            print(f"      {int(x[ci['# Samples']]) / tot * 100:5.1f}%  exec {x[ci['Instructions Executed']]:>9s}  {x[ci['Source']].strip()[:90]}")
        ops = {}
        for x in data:
            t = x[ci["Source"]].split()
            op = (t[1] if t and t[0].startswith("@") else t[0]) if t else "?"
            ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + int(x[ci["Instructions Executed"]])
        n = sum(ops.values()) or 1
        print("   opcode mix:", ", ".join(f"{k} {v / n * 100:.1f}%" for k, v in sorted(ops.items(), key=lambda x: -x[1])[:12]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 12)
