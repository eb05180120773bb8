"""profiles/<tag>_sass_mnemonics.md: per kernel of libosudit.so, the counts of the SASS mnemonics that show which
hardware path it uses (tcgen05 / TMEM / TMA vs mma.sync).   python tools/sass_table.py r02"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
so = os.path.join(ROOT, "osu-diffusion_b200", "libosudit.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTCBAR[.\w]*|UTCATOMSWS[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|"
                 r"UTMAREDG[.\w]*|UBLKCP[.\w]*|SYNCS\.PHASECHK[.\w]*|HMMA[.\w]*|LDSM[.\w]*|MUFU\.EX2|MUFU\.TANH[.\w]*)")
out, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name).replace("osudit::", "").replace("void ", "")
        out[cur] = collections.Counter()
        continue
    if cur:
        m = pat.search(line)
        if m:
            out[cur][m.group(1)] += 1
path = os.path.join(ROOT, "profiles", f"{tag}_sass_mnemonics.md")
with open(path, "w") as f:
    f.write(f"# SASS evidence (`cuobjdump -sass osu-diffusion_b200/libosudit.so`, sm_100a), round tag `{tag}`\n\n"
            "`UTCHMMA[.2CTA]` = `tcgen05.mma` (cta_group::1 / ::2), `LDTM` / `STTM` = `tcgen05.ld` / `tcgen05.st`, `UTCBAR` = "
            "`tcgen05.commit`, `UTCATOMSWS` = TMEM allocation, `UTMALDG` / `UTMASTG` / `UTMAREDG.ADD` = TMA tensor load / store / "
            "reduce-add, `SYNCS.PHASECHK…TRYWAIT` = `mbarrier.try_wait`; `HMMA.16816` + `LDSM` = the `mma.sync` / `ldmatrix` "
            "flash kernels (general fallback: head_dim 72, generic masks, backward beyond 128 datapoints).\n\n"
            "| kernel | mnemonic counts |\n|---|---|\n")
    for k, c in out.items():
        if c and any(x.startswith(("UTC", "LDTM", "HMMA", "UTMA")) for x in c):
            f.write(f"| `{k}` | " + ", ".join(f"{n} ×{v}" for n, v in sorted(c.items())) + " |\n")
print(open(path).read()[:3000])
