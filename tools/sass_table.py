"""profiles/sass_r02_hot_kernels.md from `cuobjdump -sass libb200asr.so`: per hot kernel the instruction count and the
Blackwell-native mnemonics (run on the dev box, no GPU needed)."""
import collections, re, subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
so = ROOT / "automatic-speech-recognition-asr-onnx_b200" / "libb200asr.so"
out = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
name = None
cnt = collections.defaultdict(collections.Counter)
ninstr = collections.Counter()
MN = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTMASTG", "SYNCS", "REDG", "HMMA")
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); continue
    if name and re.match(r"\s+/\*[0-9a-f]+\*/", line):
        ninstr[name] += 1
        mm = re.search(r"\b(" + "|".join(MN) + r")\b", line)
        if mm:
            cnt[name][mm.group(1)] += 1
dem = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
rows = []
for n in sorted(ninstr, key=dem):
    d = dem(n)
    if not re.search(r"gemm_tc_kernel|attention_tc|decoder_stream_kernel|decoder_ring_kernel<\(int\)1", d):
        continue
    d = re.sub(r"\((int|bool)\)", "", d.replace("b200asr::", "").replace("void ", ""))
    d = re.sub(r"\(CUtensorMap_st.*", "", d)
    c = cnt[n]
    rows.append(f"| `{d}` | {ninstr[n]} | {ninstr[n] * 16 // 1024} | {c['UTCHMMA']} | {c['UTCQMMA']} | {c['UTCBAR']} | {c['LDTM']} | {c['UTMALDG']} | {c['UBLKCP']} | {c['SYNCS']} | {c['REDG']} | {c['HMMA']} |")
text = ("# SASS evidence, round 2 (final state): `cuobjdump -sass libb200asr.so`, per hot kernel: instruction count and the Blackwell-native mnemonics\n\n"
        "(`python tools/sass_table.py`.  UTCHMMA = `tcgen05.mma.kind::f16`, UTCQMMA = `tcgen05.mma.kind::f8f6f4` (FP8 weight path), UTCBAR = "
        "`tcgen05.commit`, LDTM / STTM = `tcgen05.ld / st`, UTMALDG / UBLKCP = TMA tensor / bulk copies, SYNCS = mbarrier, REDG = `red.global` "
        "(the 64-bit fixed-point exchange), HMMA = legacy `mma.sync`.  `decoder_stream_kernel<rows, instrumented, fp8, lean>`; "
        "`decoder_ring_kernel` = round 1's CUDA-core fallback.)\n\n"
        "| kernel | SASS instr | KB | UTCHMMA | UTCQMMA | UTCBAR | LDTM | UTMALDG | UBLKCP | SYNCS | REDG | HMMA |\n"
        "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n" + "\n".join(rows) + "\n")
(ROOT / "profiles" / "sass_r02_hot_kernels.md").write_text(text)
print(text)
