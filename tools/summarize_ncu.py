"""Summarise ncu outputs into small text files for profiles/.
  python tools/summarize_ncu.py launches <launches.csv> <out.md>
  python tools/summarize_ncu.py full <report.ncu-rep> <out.csv>"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEY_METRICS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
               "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
               "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
               "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
               "launch__occupancy_limit_registers", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
               "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
               "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct" ]


def launches(path, out):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ni, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        name = r[ni].split("(")[0][:110]
        tot[name] += float(r[vi].replace(",", "")) / 1e3   # ns -> us
        cnt[name] += 1
    total = sum(tot.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({sum(cnt.values())} launches, {total/1e3:.2f} ms of kernel time; "
                "cold-cache, serialised: compare SHARES)\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k in sorted(tot, key=tot.get, reverse=True)[:40]:
            f.write(f"| `{k}` | {cnt[k]} | {tot[k]:.1f} | {100*tot[k]/total:.2f}% |\n")


def full(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[0]
    idx = [hdr.index(m) for m in KEY_METRICS if m in hdr]
    with open(out, "w") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([rows[1][i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
