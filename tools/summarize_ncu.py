"""Turns the ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches.md [--skip N]
    python tools/summarize_ncu.py full gpurun_out/prof_sweep.ncu-rep profiles/r01_plane_sweep_ncu.md
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict, defaultdict


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^void ", "", name)
    return name[:90]


def launches(path, out, skip=0):
    rows = []
    with open(path, newline="") as fh:
        text = fh.read()
    start = text.find('"ID"')
    reader = csv.DictReader(io.StringIO(text[start:]))
    for r in reader:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            val = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
            rows.append((short(r["Kernel Name"]), val * scale))
    rows = rows[skip:]
    tot = sum(t for _, t in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for k, t in rows:
        agg[k][0] += 1
        agg[k][1] += t
    with open(out, "w") as fh:
        fh.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; serialised + cold-cache: compare SHARES)\n\n")
        fh.write("source: `%s`, %d launches, total %.1f us\n\n| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n" % (path, len(rows), tot))
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write("| `%s` | %d | %.1f | %.1f%% | %.2f |\n" % (k, n, t, 100 * t / tot, t / n))
    print("wrote", out)


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fma.sum"]


def full(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    reader = csv.reader(io.StringIO(txt))
    header = next(reader)
    units = next(reader)
    rows = list(reader)
    idx = {h: i for i, h in enumerate(header)}
    with open(out, "w") as fh:
        fh.write("# ncu --set full summary of `%s`\n\n" % path)
        for r in rows:
            fh.write("## %s (launch id %s)\n\n| metric | value | unit |\n|---|---:|---|\n" % (short(r[idx["Kernel Name"]]), r[idx["ID"]]))
            for k in header:
                if any(k == kk or k.startswith(kk) for kk in KEYS) or "stall" in k and "pct" in k:
                    fh.write("| %s | %s | %s |\n" % (k, r[idx[k]], units[idx[k]]))
            fh.write("\n")
    print("wrote", out)


def table(path, out):
    """One row per profiled launch: duration, tensor-pipe activity, issue slots, DRAM traffic -- the compact evidence table."""
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    reader = csv.reader(io.StringIO(txt))
    header = next(reader)
    next(reader)
    rows = list(reader)
    idx = {h: i for i, h in enumerate(header)}
    cols = [("us", "gpu__time_duration.sum"), ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            ("tensor (elapsed) %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            ("SM busy %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"), ("grid", "launch__grid_size"),
            ("DRAM rd MB", "dram__bytes_read.sum"), ("DRAM wr MB", "dram__bytes_write.sum")]
    cols = [(n, k) for n, k in cols if k in idx]
    with open(out, "w") as fh:
        fh.write("# per-launch table from `%s` (ncu --set full, --clock-control none)\n\n" % path)
        fh.write("| # | kernel | " + " | ".join(n for n, _ in cols) + " |\n|---|---|" + "---:|" * len(cols) + "\n")
        for r in rows:
            fh.write("| %s | `%s` | %s |\n" % (r[idx["ID"]], short(r[idx["Kernel Name"]]), " | ".join(r[idx[k]] for _, k in cols)))
    print("wrote", out)


if __name__ == "__main__":
    if sys.argv[1] == "table":
        table(sys.argv[2], sys.argv[3])
        sys.exit(0)
    if sys.argv[1] == "launches":
        skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
        launches(sys.argv[2], sys.argv[3], skip)
    else:
        full(sys.argv[2], sys.argv[3])
