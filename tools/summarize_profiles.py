"""Turn gpurun_out ncu artefacts into the tracked summaries under profiles/.

  python tools/summarize_profiles.py <tag> <launches.csv> <gemm.ncu-rep> <attn_ln.ncu-rep>
"""
import collections, csv, io, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
out = open(os.path.join(ROOT, "profiles", f"{tag}_summary.md"), "w")

def w(*a): print(*a, file=out)

w(f"# ncu summary `{tag}` — DiT-B, 64 beatmaps x 2048 datapoints (128 model rows), CFG, band W=128")
w("\nLaunch list: `ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 400 python bench.py --steps 1")
w("--warmup 1` (launches 2000-2399 of the bench's own sampling loop, ~4 denoising steps). `--set full` captures: the")
w("same model/shape through `python tools/profile_step.py 64 2` (two denoising steps) so that the replayed kernels are")
w("reached quickly. One B200, `--clock-control none`. ncu times are cold-cache and serialised: compare SHARES with")
w("bench.py's live `share_of_step`, not absolutes.\n")

lines = [l for l in open(launches) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v
    name = row["Kernel Name"]
    key = re.sub(r"\(.*", "", name)
    key = re.sub(r"^void ", "", key).replace("osudit::", "")
    agg[key][0] += 1; agg[key][1] += v
tot = sum(v[1] for v in agg.values())
w("## Launch list (gpu__time_duration.sum), aggregated by kernel\n")
w("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if v[1] / tot < 0.0005: continue
    w(f"| `{k[:70]}` | {v[0]} | {v[1]:.3f} | {v[1] / tot:.3f} |")
w(f"| total | | {tot:.3f} | 1.000 |")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    if len(r) < 3: continue
    hdr, units = r[0], r[1]
    w(f"\n## `--set full` capture: {os.path.basename(rep)}\n")
    cols = [hdr.index(c) for c in WANT if c in hdr]
    w("| kernel | " + " | ".join(f"{hdr[c].split('.')[0].replace('__', ' ')} [{units[c]}]" for c in cols) + " |")
    w("|---|" + "---:|" * len(cols))
    ki = hdr.index("Kernel Name")
    for row in r[2:]:
        w(f"| `{re.sub(r'osudit::|void ', '', row[ki])[:60]}` | " + " | ".join(row[c][:10] for c in cols) + " |")
out.close()
print(open(out.name).read())
