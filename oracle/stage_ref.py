"""Stage the reference's own graph-defining modules into oracle/_ref/ as compiled code (DEV CONTAINER ONLY).

    python oracle/stage_ref.py          (needs /root/reference; run by __graft_entry__.build())

TEST / BASELINE INFRASTRUCTURE.  The reference is Python, its hot path executes inside the onnxruntime wheel (not installed,
not vendored), and Whisper/Export_Whisper.py cannot be imported (module-level code loads a checkpoint and exports).  What can
run is the reference's own nn.Module wrappers -- the code the ONNX graphs are traced from.  This recipe AST-extracts exactly the
definitions oracle/ref_loader.py uses (same list), compiles them and Whisper/STFT_Process.py where they lie under
/root/reference, and writes only the marshalled code objects to oracle/_ref/whisper_ref.bin -- a build output like a compiled
.so: git-ignored (no reference source enters the history), not gpurun-ignored (it travels to the GPU box, which has no
/root/reference).  `bench.py --impl reference` and `cpu_baseline` then time the reference's OWN modules under torch eager on
the box's host cores (kind "reference") instead of the oracle port; tests/test_ref_staged.py checks staged == live == oracle.
"""
from __future__ import annotations

import ast
import hashlib
import json
import marshal
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_loader   # noqa: E402

OUT = ROOT / "oracle" / "_ref" / "whisper_ref.bin"


def stage() -> Path | None:
    if not ref_loader.reference_available():
        return None
    exp = ref_loader.REF_ROOT / "Whisper" / "Export_Whisper.py"
    stft = ref_loader.REF_ROOT / "Whisper" / "STFT_Process.py"
    src = ref_loader.patched_export_source(exp.read_text())
    body = [n for n in ast.parse(src).body
            if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and n.name in ref_loader._WANT]
    code_defs = compile(ast.Module(body=body, type_ignores=[]), "Whisper/Export_Whisper.py", "exec")
    code_stft = compile(stft.read_text(), "Whisper/STFT_Process.py", "exec")
    meta = dict(python=list(sys.version_info[:3]), names=sorted(n.name for n in body),
                sha256={p.name: hashlib.sha256(p.read_bytes()).hexdigest() for p in (exp, stft)})
    OUT.parent.mkdir(parents=True, exist_ok=True)
    OUT.write_bytes(marshal.dumps((json.dumps(meta), code_defs, code_stft)))
    return OUT


if __name__ == "__main__":
    p = stage()
    print(p if p else "/root/reference not present: nothing staged")
