#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call left in gpurun_out/ into the tracked summaries under
profiles/ (CPU only; needs the `ncu` CLI to read .ncu-rep files).

    python profiles/summarize.py r01

writes profiles/<round>_launches_<env>.csv        (per-kernel totals of the launch list)
       profiles/<round>_ncu_<env>.json            (key metrics of each --set full capture)
       profiles/<round>_steady_dram_<env>.json    (single-pass dram bytes, --cache-control none)
       profiles/<round>_ncu_<env>_wide.json       (the same for the high-occupancy build, tools/profile_wide.sh)
       profiles/roofline_traffic.json             (bytes per launch that bench.py reports as `traffic`)
       profiles/kernel_isolated.json              (avg duration, us, of ONE isolated step-kernel launch per env,
                                                   from the launch list: what bench.py reports as `kernel_us_isolated`)
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
ENVS = ("cartpole", "mountain_car", "pendulum")

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
]

TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def read_csv_lines(text):
    return list(csv.reader(io.StringIO("".join(l for l in text.splitlines(True) if l.startswith('"')))))


def launches(path):
    rows = read_csv_lines(open(path).read())
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) * (1e-3 if r[ui] == "ns" else 1.0)  # -> us
    return agg


def raw_page(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = read_csv_lines(txt)
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                v = r[i].replace(",", "")
                try:
                    v = float(v)
                except ValueError:
                    continue
                if units[i] in TO_BYTES:
                    v, u = v * TO_BYTES[units[i]], "byte"
                else:
                    u = units[i]
                d[k] = {"value": v, "unit": u}
        out.append(d)
    return out


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
    traffic = {}
    tp = os.path.join(PROF, "roofline_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp))
    isolated = {}
    ip = os.path.join(PROF, "kernel_isolated.json")
    if os.path.exists(ip):
        isolated = json.load(open(ip))
    for env in ENVS:
        lp = os.path.join(OUT, f"launches_{env}.csv")
        if os.path.exists(lp):
            agg = launches(lp)
            total = sum(t for _, t in agg.values())
            steps = [(c, t) for k, (c, t) in agg.items() if "step_kernel" in k]
            if steps:
                isolated[env] = sum(t for _, t in steps) / sum(c for c, _ in steps)
            with open(os.path.join(PROF, f"{rnd}_launches_{env}.csv"), "w") as f:
                f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches)\n")
                f.write("launches,total_us,avg_us,share,kernel\n")
                for k, (c, t) in agg.items():
                    f.write(f"{c},{t:.2f},{t / c:.2f},{t / total:.3f},\"{k}\"\n")
        rep = os.path.join(OUT, f"prof_{env}.ncu-rep")
        if os.path.exists(rep):
            caps = raw_page(rep)
            json.dump(caps, open(os.path.join(PROF, f"{rnd}_ncu_{env}.json"), "w"), indent=1)
        sp = os.path.join(OUT, f"steady_dram_{env}.csv")
        if os.path.exists(sp):
            rows = read_csv_lines(open(sp).read())
            hdr = rows[0]
            ki, mi, vi, ui = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
            per = collections.defaultdict(list)
            for r in rows[1:]:
                if "step_kernel" in r[ki] or "rollout_kernel" in r[ki]:
                    scale = TO_BYTES.get(r[ui], 1e-3 if r[ui] == "ns" else 1.0)
                    per[r[mi]].append(float(r[vi].replace(",", "")) * scale)
            if per:
                n = len(per["dram__bytes_read.sum"])
                summ = {k: sum(v) / len(v) for k, v in per.items()}
                summ["launches"] = n
                summ["note"] = ("single-pass metrics, --cache-control none, consecutive launches over the ring: "
                                "write-backs of earlier launches drain during later ones, so the per-launch average "
                                "is the steady-state DRAM traffic")
                json.dump(summ, open(os.path.join(PROF, f"{rnd}_steady_dram_{env}.json"), "w"), indent=1)
                traffic[env] = summ["dram__bytes_read.sum"] + summ["dram__bytes_write.sum"]
    for env in ENVS:  # tools/profile_wide.sh: the high-occupancy build of the step kernel
        rep = os.path.join(OUT, f"prof_wide_{env}.ncu-rep")
        if os.path.exists(rep):
            json.dump(raw_page(rep), open(os.path.join(PROF, f"{rnd}_ncu_{env}_wide.json"), "w"), indent=1)
    rep = os.path.join(OUT, "prof_rollout_cartpole.ncu-rep")
    if os.path.exists(rep):
        json.dump(raw_page(rep), open(os.path.join(PROF, f"{rnd}_ncu_rollout_cartpole.json"), "w"), indent=1)
    json.dump(traffic, open(tp, "w"), indent=1)
    json.dump(isolated, open(ip, "w"), indent=1)
    print("roofline traffic (bytes per launch):", traffic)
    print("isolated step-kernel launch (us):", isolated)


if __name__ == "__main__":
    main()
