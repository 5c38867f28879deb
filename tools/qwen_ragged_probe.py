import sys; sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
import test_gpu_qwen_ragged as t
qw = t.qw
clips = t._clips(); pcm, lens = qw.QwenEngine.pad_ragged(clips)
eng = t._engine(3, "f32")
for mx in (6, 12, -1):
    tb = eng.transcribe(pcm, t.Q, t.L, max_new=mx, lens=lens)
    ts = [eng.transcribe(c, t.Q, t.L, max_new=mx)[0] for c in clips]
    for b in range(4):
        print(mx, b, len(tb[b]), len(ts[b]), tb[b] == ts[b], tb[b][:10], ts[b][:10])
eng.set_option("graph", 0)
tb = eng.transcribe(pcm, t.Q, t.L, max_new=12, lens=lens)
print("nograph", [x[:12] for x in tb])
