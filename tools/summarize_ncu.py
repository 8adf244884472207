"""Summaries of ncu outputs for profiles/ (run here, on the CPU box).
  python tools/summarize_ncu.py launches <launches.csv>            -> per-kernel share table (markdown)
  python tools/summarize_ncu.py full <report.ncu-rep> [max_rows]   -> key counters per captured launch (markdown)
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").strip()
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"| `{k[:80]}` | {cnt[k]} | {v / 1e3:.2f} | {100 * v / T:.1f}% | {v / cnt[k]:.1f} |")
    print(f"\ntotal {T / 1e3:.1f} ms over {sum(cnt.values())} launches (cold-cache, serialised: compare shares, not absolutes)")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]


def full(path, max_rows=6):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    for r in rows[2:2 + max_rows]:
        print(f"### `{r[idx['Kernel Name']]}` grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n")
        print("| counter | value |\n|---|---|")
        for w in WANT:
            if w in idx:
                print(f"| {w} | {r[idx[w]]} {units[idx[w]]} |")
        st = sorted([(float(r[idx[h]]), h) for h in stall if r[idx[h]] not in ("", "n/a")], reverse=True)[:6]
        print("| top stalls (warps per issue) | " + ", ".join(
            f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, h in st) + " |\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 6)
