"""Build libb200asr.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python automatic-speech-recognition-asr-onnx_b200/csrc/build.py [--force]

The shared library lands next to the Python package (git-ignored, but it does
travel to the GPU box with the gpurun snapshot).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

CSRC = Path(__file__).resolve().parent
PKG = CSRC.parent
ROOT = PKG.parent
OUT = PKG / "libb200asr.so"
OBJ = ROOT / "build" / "b200asr"
SOURCES = ["frontend.cu", "gemm_simt.cu", "gemm_tc.cu", "layers.cu", "decoder.cu", "decoder_mega.cu", "decoder_ring.cu", "decoder_stream.cu", "attention_tc.cu", "sanm.cu", "qwen.cu", "engine.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]
    deps = srcs + list(CSRC.glob("*.cuh")) + [ROOT / "include" / "b200asr.h"]
    stamp = OBJ / "stamp"
    digest = _digest(deps)
    if not force and OUT.exists() and stamp.exists() and stamp.read_text() == digest:
        return OUT
    OBJ.mkdir(parents=True, exist_ok=True)

    def compile_one(src: Path):
        obj = OBJ / (src.stem + ".o")
        cmd = [NVCC, *FLAGS, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (OBJ / (src.stem + ".ptxas.txt")).write_text(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [NVCC, "-shared", "-o", str(OUT), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(digest)
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
