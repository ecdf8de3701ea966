"""Summarise ncu output into small text files for profiles/.

  python tools/summarize_ncu.py launches <launches.csv> <out.txt>      (gpu__time_duration list)
  python tools/summarize_ncu.py full <report.ncu-rep> <out.txt>        (--set full capture)
"""
import collections
import csv
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct",
    "sm__cycles_elapsed.max",
]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        name = r[kn].split("(")[0][:90]
        agg[name][0] += 1
        agg[name][1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# source: {src}; {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.2f} ms total\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1] / 1e6:10.3f} ms {100 * v[1] / tot:6.2f}%  n={v[0]:5d}  {k}\n")


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none; source: {src}\n")
        for r in rows[2:]:
            f.write(f"--- {r[idx['Kernel Name']][:110]}\n")
            for m in FULL_METRICS:
                if m in idx:
                    f.write(f"    {m} = {r[idx[m]]} {units[idx[m]]}\n")
            top = sorted(((float(r[idx[h]] or 0), h) for h in stall), reverse=True)[:6]
            f.write("    top stalls (warps per issue): " + ", ".join(
                f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}={v:.2f}" for v, h in top) + "\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
