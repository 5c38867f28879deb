"""oracle/_ref (the reference's own modules, compiled by oracle/stage_ref.py) is what `bench.py --impl reference` times when it
is present.  Hold it to the oracle port and -- in the dev container -- to the live reference: same greedy stream, and the staged
code is the code under /root/reference (sha256 recorded at staging time)."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import ref_loader, whisper_oracle as wo
from b200asr.synth import synth_pcm

pytestmark = pytest.mark.skipif(not ref_loader.staged_available(), reason="oracle/_ref not staged (python oracle/stage_ref.py)")

PROMPT = [3, 10, 11, 12]
SUP, BEG = [1, 5, 7, 13], [220, 2]


def test_staged_reference_equals_oracle_port_stream():
    dims = wo.TINY_TEST
    raw = wo.make_raw_weights(dims, 5, pos_scale=100.0)      # non-degenerate greedy stream, as in bench.py
    fw = wo.fold_weights(raw, dims, SUP, BEG)
    mods = ref_loader.build_reference_whisper(raw, dims, SUP, BEG, staged=True)
    pcm = synth_pcm(3, 24160)
    ref = ref_loader.reference_greedy(mods, dims, pcm, PROMPT, 9)
    with torch.no_grad():
        port = wo.greedy_transcribe(pcm, fw, dims, PROMPT, stop_tokens=[], max_new=9, return_logits=False)["tokens"]
    assert ref == port
    assert len(set(ref)) >= 4


@pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference only exists in the dev container")
def test_staged_code_is_the_reference_that_is_on_disk():
    ns = ref_loader.load_staged_namespace()
    meta = ns["__staged_meta__"]
    for name in ("Export_Whisper.py", "STFT_Process.py"):
        live = hashlib.sha256((ref_loader.REF_ROOT / "Whisper" / name).read_bytes()).hexdigest()
        assert meta["sha256"][name] == live, f"{name} changed since staging: re-run oracle/stage_ref.py"
    assert set(meta["names"]) == set(ref_loader._WANT)
    live_ns = ref_loader.load_whisper_namespace()
    for n in ref_loader._WANT:
        a, b = ns[n], live_ns[n]
        ca = a.__code__ if hasattr(a, "__code__") else a.forward.__code__
        cb = b.__code__ if hasattr(b, "__code__") else b.forward.__code__
        assert ca.co_code == cb.co_code, n
