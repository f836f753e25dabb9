#!/usr/bin/env python
"""Turns ncu output into the small summaries committed under profiles/.

  python profiles/summarize.py launches <launch.csv> "<command>" > profiles/rN_launches_x.md
  python profiles/summarize.py report <file.ncu-rep> [...] > profiles/rN_ncu_full_x.md
  python profiles/summarize.py traffic <file.ncu-rep> "<source note>" > profiles/rN_ncu_traffic.json

`launches` reads the CSV of `ncu --metrics gpu__time_duration.sum --clock-control none --csv`;
`report` reads `ncu --set full` reports through `ncu -i ... --page raw --csv`.
"""
import csv
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_static",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__inst_executed.sum",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("gzb::", "")


def launches(path, command):
    agg = OrderedDict()
    with open(path, newline="") as f:
        rows = [r for r in csv.reader(f) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
    for r in rows:
        k = short(r[4])
        ns = float(r[14].replace(",", ""))
        if r[13] == "us":
            ns *= 1e3
        elif r[13] == "ms":
            ns *= 1e6
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(v[1] for v in agg.values())
    print("# ncu launch list: `%s`\n" % command)
    print("`ncu --metrics gpu__time_duration.sum --clock-control none --csv` on one B200, after the same command had exited 0")
    print("without ncu. Per-launch times are cold-cache and serialised: compare SHARES with bench.py's live CUDA-event numbers.\n")
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.1f | %.1f | %.1f%% |" % (k, n, ns / 1e3, ns / 1e3 / n, 100 * ns / tot))
    print("\nTotal GPU time %.1f ms over %d launches." % (tot / 1e6, sum(v[0] for v in agg.values())))


def report(paths):
    print("# ncu --set full captures (one launch per kernel; `--clock-control none --import-source on`)\n")
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        for row in rows[2:]:
            d = {h: (row[i], units[i]) for i, h in enumerate(hdr)}
            print("## %s  (%s)\n" % (short(d["Kernel Name"][0]), p.split("/")[-1]))
            print("grid %s block %s\n" % (d.get("Grid Size", ("?",))[0], d.get("Block Size", ("?",))[0]))
            print("| metric | value | unit |\n|---|---:|---|")
            for k in KEYS:
                if k in d and d[k][0] != "":
                    print("| %s | %s | %s |" % (k, d[k][0], d[k][1]))
            print()


def traffic(path, note):
    """dram bytes per launch, FP64 / XU pipe utilisation and duration per kernel (bench.py reads
    roofline.traffic and fp64_pipe_pct_of_peak from this file)."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]

    def num(d, key, scale=None):
        if key not in d or d[key][0] == "":
            return None
        v = float(d[key][0].replace(",", ""))
        u = d[key][1]
        if scale == "bytes":
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        if scale == "us":
            v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
        return v
    kernels = OrderedDict()
    for row in rows[2:]:
        d = {h: (row[i], units[i]) for i, h in enumerate(hdr)}
        k = short(d["Kernel Name"][0]).replace("void ", "").split("<")[0]
        e = kernels.setdefault(k, {"launches": 0, "dram_bytes": 0.0, "time_us": 0.0})
        e["launches"] += 1
        e["dram_bytes"] += (num(d, "dram__bytes_read.sum", "bytes") or 0) + (num(d, "dram__bytes_write.sum", "bytes") or 0)
        e["time_us"] += num(d, "gpu__time_duration.sum", "us") or 0
        e["fp64_pipe_pct"] = num(d, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")
        e["xu_pipe_pct"] = num(d, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active")
        e["issue_active_pct"] = num(d, "sm__issue_active.avg.pct_of_peak_sustained_elapsed")
        # double-precision flops of the launch: one per DADD / DMUL, two per DFMA (thread-level, predicated-on)
        dadd = num(d, "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum")
        dmul = num(d, "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum")
        dfma = num(d, "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum")
        if dadd is not None and dmul is not None and dfma is not None:
            e["fp64_flop"] = e.get("fp64_flop", 0.0) + dadd + dmul + 2 * dfma
    for e in kernels.values():
        e["dram_bytes_per_launch"] = e["dram_bytes"] / e["launches"]
        if "fp64_flop" in e:
            e["fp64_flop_per_launch"] = e["fp64_flop"] / e["launches"]
    return {"source": note, "kernels": kernels}


def traffic_sizes(args):
    """traffic2 <out.json> <WxH> <file.ncu-rep> <note> [<WxH> <file.ncu-rep> <note> ...]: one entry per image size
    (bench.py looks its workload's size up)."""
    import json
    out = {"sizes": {}}
    for i in range(1, len(args), 3):
        out["sizes"][args[i]] = traffic(args[i + 1], args[i + 2])
    with open(args[0], "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        import json
        print(json.dumps(traffic(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""), indent=1))
    elif sys.argv[1] == "traffic2":
        traffic_sizes(sys.argv[2:])
    elif sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
    else:
        report(sys.argv[2:])
