"""Condense an ncu report into the few per-kernel numbers the roofline discussion needs.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_full_summary.md [traffic.json]

Reads the report with `ncu -i ... --page raw --csv` (works without a GPU) and writes one markdown table
row per captured launch; optionally writes {kernel: dram bytes per launch} for bench.py's `traffic`.
"""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    ("time_ms", "gpu__time_duration.sum"),
    ("dram_rd_GB", "dram__bytes_read.sum"),
    ("dram_wr_GB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("fp64_pipe_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tensor_dmma_pct", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("smem_dyn_KB", "launch__shared_mem_per_block_dynamic"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("waves", "launch__waves_per_multiprocessor"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
]

_SCALE = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
          "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}


def main():
    rep, out_md = sys.argv[1], sys.argv[2]
    traffic_path = sys.argv[3] if len(sys.argv) > 3 else None
    if rep.endswith(".csv"):                     # the raw page exported on the GPU box (`ncu -i ... --page raw --csv`)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = ["| kernel | " + " | ".join(k for k, _ in METRICS) + " |", "|---|" + "---|" * len(METRICS)]
    traffic = {}
    for r in data:
        name = r[idx["Kernel Name"]].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
        name = name.split("(")[0].replace("void ", "").replace("cmf::", "")
        cells = []
        vals = {}
        for key, m in METRICS:
            if m not in idx or r[idx[m]] == "":
                cells.append("-")
                continue
            v = float(r[idx[m]].replace(",", ""))
            u = units[idx[m]]
            if key == "time_ms" or key.endswith("_GB"):
                v *= _SCALE.get(u, 1.0)
            if key == "smem_dyn_KB":
                v *= {"byte": 1 / 1024.0, "Kbyte": 1.0}.get(u, 1.0)
            vals[key] = v
            cells.append("%.3f" % v if abs(v) < 100 else "%.0f" % v)
        lines.append("| %s | %s |" % (name, " | ".join(cells)))
        base = name.split("<")[0]
        if "dram_rd_GB" in vals and base not in traffic:
            traffic[base] = (vals["dram_rd_GB"] + vals.get("dram_wr_GB", 0.0)) * 1e9
    with open(out_md, "w") as fh:
        fh.write("ncu --set full --clock-control none capture (%s); one row per captured launch.\n" % rep)
        fh.write("Times are cold-cache and serialised under the profiler: compare shares, not absolutes.\n\n")
        fh.write("\n".join(lines) + "\n")
    if traffic_path:
        json.dump(traffic, open(traffic_path, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
