"""profiles/ncu_r01_qwen_summary.md from the Qwen3-ASR captures (gpurun_out/qwen_launches_final.csv: launch list of one decode
step; prof_qwen_gemv.ncu-rep / prof_qwen_attn.ncu-rep: `ncu --set full` of the four GEMVs of a layer and of the split attention)."""
import collections
import csv
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
from ncu_summary import OUT, ROOT, table  # noqa: E402


def main():
    out = ["# ncu summaries, round 1: Qwen3-ASR-0.6B bf16, batch 1, 30 s clip (`bench.py --preset qwen3-asr-0.6b`)\n",
           "Per-launch times under ncu are cold-cache and serialised (no programmatic-dependent-launch overlap): shares carry over, absolutes do not.\n"]
    lines = [l for l in open(OUT / "qwen_launches_final.csv") if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    idx = [i for i, n in enumerate(names) if "qwen_embed_kernel" in n]
    step = rows[idx[0]:idx[1]]
    agg = collections.OrderedDict()
    for r in step:
        a = agg.setdefault((r["Kernel Name"].split("(")[0][-60:], r["Grid Size"]), [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"].replace(",", ""))
    tot = sum(v for _, v in agg.values())
    out.append(f"## Launch list of one decode step: {tot / 1e3:.0f} us over {len(step)} launches (replayed as one CUDA graph: {sys.argv[1] if len(sys.argv) > 1 else '?'} ms measured)\n")
    out.append("| kernel | grid | launches | total us | share | avg us |\n|---|---|---:|---:|---:|---:|")
    for (k, g), (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {g} | {n} | {v / 1e3:.1f} | {100 * v / tot:.1f}% | {v / n / 1e3:.1f} |")
    for rep, title in (("prof_qwen_gemv.ncu-rep", "qwen_gemv_kernel: the four GEMVs of one decoder layer (qkv 8.4 MB, o 4.2 MB, gate_up 12.6 MB, down 6.3 MB of bf16 weights)"),
                       ("prof_qwen_attn.ncu-rep", "qwen_attn_split_kernel (16 heads x 8 key ranges, kv_len ~ 410)")):
        if (OUT / rep).exists():
            table(OUT / rep, title, out)
    (ROOT / "profiles" / "ncu_r01_qwen_summary.md").write_text("\n".join(out) + "\n")
    print("\n".join(out)[:5000])


if __name__ == "__main__":
    main()
