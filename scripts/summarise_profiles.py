#!/usr/bin/env python
"""Summarise the ncu reports of scripts/capture_profiles.sh (`ncu -i`, no GPU needed; the capture script runs it on the GPU
box into gpurun_out/prof/ because the reports themselves exceed what gpurun copies back): writes r02_traffic.json (DRAM bytes
per launch, read by bench.py for roofline.traffic) and one text summary per report.
    python scripts/summarise_profiles.py [output directory, default profiles/]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles")   # summaries + r02_traffic.json go here
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__sass_inst_executed_op_utcmma.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "msecond": 1e-3, "usecond": 1e-6, "second": 1.0, "nsecond": 1e-9}


def raw(rep):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    out = []
    for row in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, row):
            if h in ("Kernel Name", "ID"):
                d[h] = v
            elif h in KEYS:
                try:
                    d[h] = float(v.replace(",", "")) * UNIT.get(u, 1.0)
                except ValueError:
                    d[h] = v
        out.append(d)
    return out


def main():
    os.makedirs(PROF, exist_ok=True)
    traffic = {}
    for d in (PROF, os.path.join(ROOT, "profiles")):   # entries of reports that are not re-captured are kept
        try:
            traffic = json.load(open(os.path.join(d, "r02_traffic.json")))
            break
        except Exception:
            pass
    for rep in sorted(f for f in os.listdir(OUT) if f.startswith("r02_") and f.endswith(".ncu-rep")):
        ks = raw(os.path.join(OUT, rep))
        if not ks:
            continue
        tag = rep[:-len(".ncu-rep")]
        lines = ["ncu --set full --clock-control none --import-source on  (%s; scripts/capture_profiles.sh)" % rep]
        tot = 0.0
        for k in ks:
            name = k.get("Kernel Name", "?")
            lines.append("\n## " + name[:150])
            for key in KEYS:
                if key in k:
                    lines.append("%s = %s" % (key, k[key]))
            tot += k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0)
        open(os.path.join(PROF, tag + "_ncu_summary.txt"), "w").write("\n".join(lines) + "\n")
        if tag == "r02_k_episode":
            traffic["k_episode"] = {"games": 4096, "sims": 200, "dram_bytes_per_launch": tot,
                                    "kernel_ms_under_ncu": 1e3 * ks[0].get("gpu__time_duration.sum", 0.0)}
        elif tag.startswith("r02_net_"):
            w, h, b, mode = tag[len("r02_net_"):].split("_", 3)
            if b == "8192":
                traffic["net_%sx%s_%s" % (w, h, mode)] = {"batch": int(b), "dram_bytes_per_launch": tot,
                                                          "kernels": [k.get("Kernel Name", "?")[:60] for k in ks]}
    json.dump(traffic, open(os.path.join(PROF, "r02_traffic.json"), "w"), indent=1)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
