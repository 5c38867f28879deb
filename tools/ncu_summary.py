"""Turn the captures of tools/profile_all.sh (gpurun_out/*.ncu-rep, launches_r1.csv) into profiles/ncu_r01_summary.md."""
import collections
import csv
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "sm__icc_request_hit_rate.pct"]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def table(rep, title, out):
    recs, units = raw(rep)
    out.append(f"\n## `--set full`: {title}\n")
    for r in recs:
        out.append("| metric | value |\n|---|---|")
        for k in KEYS:
            if k in r and r[k] != "":
                out.append(f"| {k} | {r[k]} {units.get(k, '')} |")
        out.append("")


def main():
    out = ["# ncu summaries, round 1 (B200, whisper-large-v3 bf16, 8 s clips; commands in `tools/profile_all.sh`)\n",
           "Per-launch times under ncu are cold-cache and serialised: the kernel's SHARE of the step is what carries over.\n"]
    rows = [r for r in csv.reader(open(OUT / "launches_r1.csv")) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    seq = [(r[ki], r[gi], float(r[vi].replace(",", ""))) for r in rows[1:]]
    tot = sum(v for _, _, v in seq)
    agg = collections.OrderedDict()
    for k, g, v in seq:
        a = agg.setdefault(k.split("(")[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    out.append(f"## Launch list of one step (encode + 4-token prefill + 32 greedy steps, batch 1): {tot / 1e3:.0f} us over {len(seq)} launches\n")
    out.append("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {v / 1e3:.1f} | {100 * v / tot:.1f}% | {v / n / 1e3:.1f} |")
    for rep, title in (("prof_ring_r1.ncu-rep", "decoder_ring_kernel<1> (8 greedy steps in one launch, batch 1)"),
                       ("prof_attn_r1.ncu-rep", "attention_tc_kernel, batch 1 (grid 4 x 20 x 1)"),
                       ("prof_attn_b4_r1.ncu-rep", "attention_tc_kernel, batch 4 (grid 4 x 20 x 4)"),
                       ("prof_gemm_r1.ncu-rep", "gemm_tc_kernel (encoder linears, batch 1: M = 400)")):
        if (OUT / rep).exists():
            table(OUT / rep, title, out)
    (ROOT / "profiles" / "ncu_r01_summary.md").write_text("\n".join(out) + "\n" + (sys.argv[1] if len(sys.argv) > 1 else ""))
    print("\n".join(out)[:6000])


if __name__ == "__main__":
    main()
