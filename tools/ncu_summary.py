"""Selected metrics of one or more .ncu-rep captures as one CSV (columns = captures).
usage: python tools/ncu_summary.py out.csv label=file.ncu-rep [label=file.ncu-rep ...]   (runs ncu -i; no GPU needed)"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active"]
STALL = "smsp__average_warps_issue_stalled_"


SCALE = {"ns": ("ms", 1e-6), "us": ("ms", 1e-3), "ms": ("ms", 1.0), "s": ("ms", 1e3), "nsecond": ("ms", 1e-6), "usecond": ("ms", 1e-3),
         "msecond": ("ms", 1.0), "second": ("ms", 1e3), "byte": ("Mbyte", 1e-6), "Kbyte": ("Mbyte", 1e-3), "Mbyte": ("Mbyte", 1.0),
         "Gbyte": ("Mbyte", 1e3)}


def load(path):
    """One capture as {metric: (unit, value)}; durations are normalised to ms and byte counts to Mbyte (ncu picks the
    unit per capture)."""
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    d = {}
    for k, u, v in zip(rows[0], rows[1], rows[2]):
        if (k.startswith("gpu__time_duration") or k.startswith("dram__bytes")) and u in SCALE:
            try:
                u, v = SCALE[u][0], "%.6f" % (float(v.replace(",", "")) * SCALE[u][1])
            except ValueError:
                pass
        d[k] = (u, v)
    return d


def main():
    dst, caps = sys.argv[1], [a.rsplit("=", 1) for a in sys.argv[2:]]
    data = [(label, load(path)) for label, path in caps]
    keys = list(WANT)
    for k in data[0][1]:
        if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
            keys.append(k)
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [label for label, _ in data])
        for k in keys:
            if any(k in d for _, d in data):
                unit = next(d[k][0] for _, d in data if k in d)
                w.writerow([k, unit] + [d.get(k, ("", ""))[1] for _, d in data])


if __name__ == "__main__":
    main()
