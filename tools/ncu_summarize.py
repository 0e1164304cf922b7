"""Turn an ncu report (gpurun_out/*.ncu-rep) into the small, tracked summaries under profiles/:
   python tools/ncu_summarize.py gpurun_out/prof_x.ncu-rep profiles/r01_x  [workload-key]
writes <out>.json (key metrics per captured launch), <out>_opcodes.txt (executed-instruction mix of the first launch)
and, when a workload key is given, updates profiles/ncu_summary.json (read by bench.py for roofline.traffic)."""
import csv
import io
import json
import os
import subprocess
import sys
from collections import Counter

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sectors.sum": "l2_sectors",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slot_utilisation_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_instruction",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__occupancy_limit_registers": "occupancy_limit_registers_blocks",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "smsp__inst_executed.sum": "warp_instructions",
}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, out = sys.argv[1], sys.argv[2]
    workload = sys.argv[3] if len(sys.argv) > 3 else None
    rows = ncu_csv(rep, "raw")
    head, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[head.index("Kernel Name")]}
        for k, name in KEYS.items():
            if k not in head:
                continue
            i = head.index(k)
            try:
                v = float(r[i])
            except ValueError:
                continue
            u = units[i]
            if name == "duration":
                d["duration_ms"] = v * UNIT_SCALE.get(u, 1.0)
            elif name in ("dram_read", "dram_write"):
                d[name + "_bytes"] = v * UNIT_SCALE.get(u, 1.0)
            else:
                d[name] = v
        d["dram_bytes"] = d.get("dram_read_bytes", 0.0) + d.get("dram_write_bytes", 0.0)
        d["warp_execution_efficiency_pct"] = 100.0 * d.get("threads_per_instruction", 0.0) / 32.0
        launches.append(d)
    json.dump({"report": os.path.basename(rep), "launches": launches}, open(out + ".json", "w"), indent=1)
    # opcode mix of the first captured launch
    src = ncu_csv(rep, "source", ("--print-source", "sass"))
    hi = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    h = src[hi]
    ia, ie = h.index("Source"), h.index("Instructions Executed")
    mix = Counter()
    total = 0
    for r in src[hi + 1:]:
        if len(r) <= ie or (r and r[0] == "Kernel Name"):
            break
        s = r[ia].strip().split()
        if not s:
            continue
        op = s[1] if s[0].startswith("@") and len(s) > 1 else s[0]
        n = int(r[ie] or 0)
        mix[op.split(".")[0]] += n
        total += n
    with open(out + "_opcodes.txt", "w") as f:
        f.write(f"# executed warp-instructions by opcode, first captured launch of {os.path.basename(rep)}; total {total}\n")
        for op, n in mix.most_common(40):
            f.write(f"{op:10s} {n:14d} {100.0 * n / max(total, 1):6.2f}%\n")
    if workload and launches:
        path = os.path.join(os.path.dirname(os.path.abspath(out)), "ncu_summary.json")
        allw = json.load(open(path)) if os.path.exists(path) else {}
        first = launches[0]
        allw[workload] = {"dram_bytes_per_launch": first["dram_bytes"], "l2_hit_pct": first.get("l2_hit_pct"), "l1_hit_pct": first.get("l1_hit_pct"),
                          "issue_slot_utilisation_pct": first.get("issue_slot_utilisation_pct"), "warp_execution_efficiency_pct": first["warp_execution_efficiency_pct"],
                          "duration_ms_under_ncu": first.get("duration_ms"), "warp_instructions_per_launch": first.get("warp_instructions"),
                          "threads_per_instruction": first.get("threads_per_instruction"), "kernel": first.get("kernel"), "source": os.path.basename(out) + ".json"}
        json.dump(allw, open(path, "w"), indent=1)
    print(json.dumps(launches[0], indent=1))


if __name__ == "__main__":
    main()
