"""Round-2 (second half: attn_stream.cu, converged MMA issue, weight re-pack kernel) ncu artefacts -> profiles/r02b_summary.md (+ the launch lists copied beside it).

  python tools/summarize_profiles_r02.py gpurun_out

Expects r02b_launches.csv (bench.py's sampling loop), r02b_train_launches.csv (tools/train_step_for_ncu.py),
r02b_attn.ncu-rep (tools/attn_for_ncu.py, --set full) and r02_train_gemm.ncu-rep (--set full over the training GEMMs)."""
import collections, csv, io, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out")
out = open(os.path.join(ROOT, "profiles", "r02b_summary.md"), "w")
def w(*a): print(*a, file=out)

def launch_table(path, title, note):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        v = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v
        key = re.sub(r"\(.*", "", row["Kernel Name"])
        key = re.sub(r"^void ", "", key).replace("osudit::", "")
        agg[key][0] += 1; agg[key][1] += v
    tot = sum(v[1] for v in agg.values())
    w(f"\n## {title}\n\n{note}\n")
    w("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 0.002: continue
        w(f"| `{k[:78]}` | {v[0]} | {v[1]:.3f} | {v[1] / tot:.3f} |")
    w(f"| total | | {tot:.3f} | 1.000 |")

WANT = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]
SHORT = ["grid", "block", "regs", "time", "dram rd", "dram wr", "tensor pipe %", "XU pipe %", "dram %", "issue %", "L2 hit %"]

def full_table(path, title, note, dedup=True):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    hdr, units = r[0], r[1]
    cols = [(hdr.index(c), s) for c, s in zip(WANT, SHORT) if c in hdr]
    w(f"\n## {title}\n\n{note}\n")
    w("| kernel | " + " | ".join(f"{s} [{units[c]}]" if units[c] else s for c, s in cols) + " |")
    w("|---|" + "---:|" * len(cols))
    ki = hdr.index("Kernel Name")
    seen = set()
    for row in r[2:]:
        name = re.sub(r"osudit::|void ", "", row[ki])[:64]
        key = (name, row[cols[0][0]], round(float(row[hdr.index("gpu__time_duration.sum")].replace(",", "")) / 20))
        if dedup and key in seen: continue
        seen.add(key)
        w(f"| `{name}` | " + " | ".join(row[c][:9] for c, _ in cols) + " |")

w("# ncu summary `r02b` (round 2, after attn_stream.cu / converged MMA issue / re-pack kernel) — one B200, `--clock-control none`")
w("\nncu times are cold-cache and serialised: compare SHARES with the live numbers (`bench.py`'s `share_of_step`, "
  "`tools/train_profile.py`), not absolutes.")
launch_table(os.path.join(src, "r02b_launches.csv"), "Sampling: launch list of bench.py (BASELINE config 2)",
             "`ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 400 python bench.py --steps 1 --warmup 1 "
             "--no-train`: launches 2000-2399 of the sampling loop (~4 denoising steps of DiT-B, 128 rows x 2048, CFG, band W=128).")
launch_table(os.path.join(src, "r02b_train_launches.csv"), "Training: launch list of one DiT-B step (BASELINE config 3, batch 256 x 128)",
             "`ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 500 python tools/train_step_for_ncu.py` "
             "(CUDA graphs off so that ncu sees the launches; fused optimizer): the third step.")
full_table(os.path.join(src, "r02b_attn.ncu-rep"), "`--set full`: attention kernels (tools/attn_for_ncu.py)",
           "In launch order: streaming forward `attn_stream_kernel` with log-sum-exp at config 3 (256 x 128 x 12 heads) and the "
           "tcgen05 backward `attn_bwd_tc_kernel` at the same shape; `attn_stream_kernel` at config 5's per-GPU shape (128 x 512 x 16 "
           "heads, full attention); `attn_stream_kernel` and the round-1 window kernel at config 2 (128 x 2048 x 12 heads, band W=128). "
           "Repeated launches of the same shape are shown once.")
out.close()
for f in ("r02b_launches.csv", "r02b_train_launches.csv"):
    shutil.copyfile(os.path.join(src, f), os.path.join(ROOT, "profiles", f))
print(open(out.name).read())
