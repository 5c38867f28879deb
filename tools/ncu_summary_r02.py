"""Turn the captures of tools/profile_r02.sh (gpurun_out/*_r02.ncu-rep, launches_r02.csv) into
profiles/ncu_r02_stream_summary.md, profiles/launches_r02_largev3_b1.csv and profiles/ncu_r02_stream.json
(the measured DRAM bytes per greedy step that bench.py reports as roofline.traffic)."""
import collections, csv, json, shutil, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
sys.path.insert(0, str(ROOT / "tools"))
from ncu_summary import KEYS, raw

STEPS = 8
out = ["# ncu summaries, round 2 (B200, whisper-large-v3 bf16, 8 s clips; commands in `tools/profile_r02.sh`)\n",
       "Per-launch times under ncu are cold-cache and serialised: the kernel's SHARE of the step is what carries over.\n"]
rows = [r for r in csv.reader(open(OUT / "launches_r02.csv")) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[1:]]
tot = sum(v for _, v in seq)
agg = collections.OrderedDict()
for k, v in seq:
    a = agg.setdefault(k.split("(")[0][:70], [0, 0.0]); a[0] += 1; a[1] += v
out.append(f"## Launch list of one step (encode + 4-token prefill + 32 greedy steps in ONE decoder launch, batch 1): {tot / 1e3:.0f} us over {len(seq)} launches\n")
out.append("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {n} | {v / 1e3:.1f} | {100 * v / tot:.1f}% | {v / n / 1e3:.1f} |")
traffic = {}
for rep, b in (("prof_stream_b1_r02.ncu-rep", 1), ("prof_stream_b4_r02.ncu-rep", 4), ("prof_stream_b1_fp8_r02.ncu-rep", "1_fp8")):
    if not (OUT / rep).exists():
        continue
    recs, units = raw(OUT / rep)
    out.append(f"\n## `--set full`: decoder_stream_kernel, batch {b} ({STEPS} greedy steps in one launch)\n")
    for r in recs:
        out.append("| metric | value |\n|---|---|")
        for k in KEYS + ["smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "lts__t_sector_hit_rate.pct"]:
            if k in r and r[k] != "":
                out.append(f"| {k} | {r[k]} {units.get(k, '')} |")
        def val(k):
            v, u = float(r[k].replace(",", "")), units.get(k, "")
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
        traffic[str(b)] = (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) / STEPS
        out.append(f"\nDRAM bytes per greedy step (read + write) = {traffic[str(b)] / 1e9:.4f} GB\n")
for rep, title in (("prof_attn_b4_r02.ncu-rep", "fused encoder attention, 4 clips x 8 s (first launch of the step)"),
                   ("prof_gemm_b4_r02.ncu-rep", "encoder GEMMs, 4 clips x 8 s (first three launches: conv1, conv2, layer-0 QKV)")):
    if not (OUT / rep).exists():
        continue
    recs, units = raw(OUT / rep)
    out.append(f"\n## `--set full`: {title}\n")
    for r in recs:
        out.append("| metric | value |\n|---|---|")
        for k in KEYS:
            if k in r and r[k] != "":
                out.append(f"| {k} | {r[k]} {units.get(k, '')} |")
        out.append("")
(ROOT / "profiles" / "ncu_r02_stream_summary.md").write_text("\n".join(out) + "\n" + (sys.argv[1] if len(sys.argv) > 1 else ""))
(ROOT / "profiles" / "ncu_r02_stream.json").write_text(json.dumps(
    {"preset": "whisper-large-v3", "precision": "bf16", "steps_per_launch": STEPS, "dram_bytes_per_step": traffic,
     "source": "ncu --set full --clock-control none, tools/profile_r02.sh"}, indent=1))
shutil.copy(OUT / "launches_r02.csv", ROOT / "profiles" / "launches_r02_largev3_b1.csv")
print("\n".join(out)[:5000])
