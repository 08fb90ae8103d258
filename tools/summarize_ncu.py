"""Turn the round's ncu outputs in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_ncu.py <tag-in-gpurun_out> <tag-in-profiles> [full-only]

launches_<tag>.csv (gpu__time_duration pass) -> ncu_launches_<ptag>.csv + ncu_launches_<ptag>_summary.txt (per-kernel shares)
prof_<tag>.ncu-rep (--set full)              -> ncu_full_<ptag>_kernels.csv (one row per captured launch, key metrics)
With "full-only" (e.g. the training-kernel capture prof_train_<tag>.ncu-rep) only the second file is produced.
"""
import collections
import csv
import re
import shutil
import subprocess
import sys

tag, ptag = sys.argv[1], sys.argv[2]
full_only = len(sys.argv) > 3 and sys.argv[3] == "full-only"

if not full_only:
    rows = list(csv.reader(open(f"gpurun_out/launches_{tag}.csv", errors="replace")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = re.sub(r"\(.*", "", r[kn])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", "")) / 1e6  # ns -> ms
    tot = sum(v[1] for v in agg.values())
    shutil.copy(f"gpurun_out/launches_{tag}.csv", f"profiles/ncu_launches_{ptag}.csv")
    with open(f"profiles/ncu_launches_{ptag}_summary.txt", "w") as fh:
        fh.write("ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 420  python bench.py --steps 1 --warmup 0 --k 2 --no-cpu-baseline\n")
        fh.write("(cold-cache, serialised launches: compare SHARES)\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"{k[:64]:64s} n={v[0]:4d} total={v[1]:8.3f} ms share={100 * v[1] / tot:5.1f}%  avg={1e3 * v[1] / v[0]:8.1f} us\n")
        fh.write(f"total {tot:.3f} ms\n")
    print(open(f"profiles/ncu_launches_{ptag}_summary.txt").read())

want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "sm__cycles_elapsed.avg.per_second", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", f"gpurun_out/prof_{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
cols = [h.index("ID"), h.index("Kernel Name")] + [h.index(w) for w in want if w in h]
with open(f"profiles/ncu_full_{ptag}_kernels.csv", "w", newline="") as fh:
    w = csv.writer(fh)
    for r in rr:
        w.writerow([r[c] for c in cols])
print("wrote", f"profiles/ncu_full_{ptag}_kernels.csv", len(rr) - 2, "launches")
