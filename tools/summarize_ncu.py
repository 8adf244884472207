"""Summaries of ncu outputs for profiles/ (run here, on the CPU box).
  python tools/summarize_ncu.py launches <launches.csv>            -> per-kernel share table (markdown)
  python tools/summarize_ncu.py full <report.ncu-rep> [max_rows]   -> key counters per captured launch (markdown)
  python tools/summarize_ncu.py eval <launches.csv> [N] [json_out] -> ONE evaluation (first kbuild .. first grad_finalize) of a
        launch list taken with gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum: per-kernel time share and
        DRAM traffic; json_out receives the per-evaluation totals bench.py reports as roofline.traffic
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


def one_eval(path, N=32768, json_out=None):
    import json
    lines = [l for l in open(path) if not l.startswith("==")]
    by = {}
    for r in csv.DictReader(lines):
        d = by.setdefault(int(r["ID"]), {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").strip()})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    ids = sorted(by)
    k0 = next(i for i in ids if "kbuild" in by[i]["name"])
    k1 = next(i for i in ids if "grad_finalize" in by[i]["name"] and i > k0)
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for i in ids:
        if k0 <= i <= k1:
            a = agg[by[i]["name"]]
            a[0] += 1
            a[1] += by[i].get("gpu__time_duration.sum", 0.0)
            a[2] += by[i].get("dram__bytes_read.sum", 0.0)
            a[3] += by[i].get("dram__bytes_write.sum", 0.0)
    T = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | share | DRAM read GB | DRAM write GB | avg GB/s |\n|---|---:|---:|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k[:70]}` | {a[0]} | {a[1] / 1e6:.2f} | {100 * a[1] / T:.1f}% | {a[2] / 1e9:.2f} | {a[3] / 1e9:.2f} | {(a[2] + a[3]) / a[1]:.0f} |")
    tot_r, tot_w = sum(a[2] for a in agg.values()), sum(a[3] for a in agg.values())
    print(f"\none evaluation: {k1 - k0 + 1} launches, {T / 1e6:.1f} ms under ncu (cold-cache, serialised), DRAM read {tot_r / 1e9:.1f} GB + write {tot_w / 1e9:.1f} GB; "
          f"algorithmic minimum 24 N^2 = {24 * N * N / 1e9:.1f} GB")
    if json_out:
        g = [a for k, a in agg.items() if "dgemm128" in k]
        json.dump({"N": N, "source": path.split("/")[-1], "launches_per_eval": k1 - k0 + 1,
                   "dgemm128_launches": sum(a[0] for a in g), "dgemm128_dram_gb_per_eval": round(sum(a[2] + a[3] for a in g) / 1e9, 2),
                   "all_kernels_dram_gb_per_eval": round((tot_r + tot_w) / 1e9, 2), "algorithmic_gb_per_eval": round(24 * N * N / 1e9, 2),
                   "dgemm128_time_share_under_ncu": round(sum(a[1] for a in g) / T, 4)}, open(json_out, "w"), indent=1)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        # tcgen05 kernels: the (whole) tensor pipe, the integer sub-pipe, L2 -> SM traffic, TMEM / tensor-core instruction pipes
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum",
        "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


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
    if sys.argv[1] == "eval":
        one_eval(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 32768, sys.argv[4] if len(sys.argv) > 4 else None)
    elif sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 6)
