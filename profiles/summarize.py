"""Summarise ncu reports into the text files committed under profiles/.

    python profiles/summarize.py launches gpurun_out/r01_launches.csv > profiles/r01_launches_summary.txt
    python profiles/summarize.py full gpurun_out/prof.ncu-rep          > profiles/r01_full_summary.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
    "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        agg.setdefault(row["Kernel Name"][:72], []).append(float(row["Metric Value"].replace(",", "")))
    def aux(k):   # synthetic-frame generators and the bench's spin gate in front of the timed region
        return "k_background" in k or "k_birds" in k or "spin_kernel" in k
    tot = sum(sum(v) for k, v in agg.items() if not aux(k))
    print("# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none`")
    print("# (cold cache, serialised: compare shares, not absolutes; generator kernels and the bench's spin gate excluded from shares)")
    for k, v in agg.items():
        gen = aux(k)
        share = "   gen" if gen else "%5.1f%%" % (100 * sum(v) / tot)
        print("%-72s n=%3d avg=%9.1f us  share=%s" % (k, len(v), sum(v) / len(v) / 1e3, share))


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    for row in rows[2:]:
        print("=" * 100)
        print(row[hdr.index("Kernel Name")][:98])
        for k in KEYS:
            if k in hdr:
                print("  %-70s %-16s %s" % (k, units[hdr.index(k)], row[hdr.index(k)]))
        tops = sorted(((float(row[hdr.index(h)] or 0), h) for h in stall), reverse=True)[:6]
        for v, h in tops:
            print("  stall %-64s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
