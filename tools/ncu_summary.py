"""Summaries of ncu output for profiles/: (1) launch list (--metrics gpu__time_duration.sum CSV) grouped by kernel,
(2) key metrics of --set full reports, (3) the DRAM traffic json bench.py quotes as roofline.traffic.
Usage: python tools/ncu_summary.py launches <launches.csv>
       python tools/ncu_summary.py full <report.ncu-rep> [...]
       python tools/ncu_summary.py traffic <forward report.ncu-rep> <frames per launch> > profiles/r2_traffic.json"""
import collections
import csv
import io
import json
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"), ("launch__occupancy_limit_shared_mem", "occ limit smem (CTAs)"),
    ("launch__occupancy_limit_registers", "occ limit regs (CTAs)"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__inst_executed.sum", "warp instructions"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]
STALLS = "smsp__pcsamp_warps_issue_stalled_"


def raw_rows(rep):
    if rep.endswith(".csv"):  # already exported (ncu -i report --page raw --csv) on the GPU box
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    head, units = r[0], r[1]
    return head, units, r[2:]


def full(reps):
    for rep in reps:
        head, units, rows = raw_rows(rep)
        ci = {k: i for i, k in enumerate(head)}
        for row in rows:
            print(f"== {row[ci['Kernel Name']]}   [{rep.split('/')[-1]}]")
            for k, label in KEYS:
                if k in ci:
                    print(f"   {label:28s} {row[ci[k]]} {units[ci[k]]}")
            st = {k[len(STALLS):]: int(float(row[i] or 0)) for k, i in ci.items() if k.startswith(STALLS) and "not_issued" not in k}
            tot = sum(st.values()) or 1
            top = sorted(st.items(), key=lambda kv: -kv[1])[:6]
            print("   stall samples               " + ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in top))


def launches(path):
    text = open(path).read()
    start = text.index('"ID"')
    r = list(csv.DictReader(io.StringIO(text[start:])))
    agg = collections.OrderedDict()
    for row in r:
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        a = agg.setdefault(row["Kernel Name"].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print(f"# {path.split('/')[-1]}: {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time (serialised, cold caches, under ncu)")
    print(f"{'kernel':70s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {n:8d} {us:12.1f} {us / n:10.1f} {100 * us / tot:6.1f}%")


def traffic(rep, frames):
    head, units, rows = raw_rows(rep)
    ci = {k: i for i, k in enumerate(head)}

    def bytes_of(row, k):
        v, u = float(row[ci[k]]), units[ci[k]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]

    ks, total = [], 0.0
    for row in rows:
        rd, wr = bytes_of(row, "dram__bytes_read.sum"), bytes_of(row, "dram__bytes_write.sum")
        total += rd + wr
        ks.append({"kernel": row[ci["Kernel Name"]].split("(")[0], "dram_read_bytes": rd, "dram_write_bytes": wr,
                   "duration_us": float(row[ci["gpu__time_duration.sum"]]) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(units[ci["gpu__time_duration.sum"]], 1)})
    git = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    print(json.dumps({"batch": frames, "fft_log2": 20, "is_real": False, "dram_bytes_per_launch_group": total,
                      "dram_bytes_per_frame": total / frames, "kernels": ks,
                      "source": f"ncu --set full --clock-control none, one launch of each forward kernel at {frames} frames per launch "
                                f"({rep.split('/')[-1]}), build {git}"}, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "full":
        full(sys.argv[2:])
    else:
        traffic(sys.argv[2], int(sys.argv[3]))
