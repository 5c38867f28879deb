"""Layer-by-layer check of the streaming decode kernel's accumulator words against a numpy restatement of the first
prompt token's pass (tiny golden model): finds the first phase whose output is wrong."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np
from gpu_common import GOLD, load_case, make_engine

g, raw, tensors = load_case(GOLD[0])
eng = make_engine(tensors, "bf16")
dims = eng.dims
d, f, H, L, V = dims.d_model, dims.ffn, dims.n_heads, dims.dec_layers, dims.vocab
eng.encode(g["pcm"])
T = eng.T_enc
eng.set_decode_options(stop_ids=[])
eng.set_option("stream_multi", 0)        # this tool reads the one-row accumulator layout
NTOK = int(sys.argv[1]) if len(sys.argv) > 1 else 2
toks = [int(t) for t in g["prompt"].reshape(-1)[:NTOK]]
logits, first = eng.prefill(np.array([toks], np.int32))
ck = eng.get_stage("cross_k", 1 * L * H * T * 64).reshape(L, H, T, 64)
cv = eng.get_stage("cross_v", 1 * L * H * T * 64).reshape(L, H, T, 64)
NRT = 1
lw = (NRT * (6 * d + f) + 6 * NRT + 15) // 16 * 16
xr = (NRT * d + 2 * NRT + 15) // 16 * 16
setw = xr + L * lw
allv = eng.get_stage("stream_acc_val", 2 * setw)
allc = eng.get_stage("stream_acc_cnt", 2 * setw)
si = (NTOK - 1) & 1
val = allv[si * setw:(si + 1) * setw]
cnt = allc[si * setw:(si + 1) * setw]
import math, torch

def bf(x):
    return torch.tensor(np.asarray(x, np.float32)).to(torch.bfloat16).to(torch.float32).numpy()

def W(n): return bf(np.asarray(tensors[n], np.float32))
def Fv(n): return np.asarray(tensors[n], np.float32).reshape(-1)

def ln(x):
    mu = x.mean(); var = ((x - mu) ** 2).mean()
    return (x - mu) / np.sqrt(var + 1e-5)

def report(name, got, ref, c=None):
    err = np.abs(got - ref).max()
    print(f"{name:16s} max|d| {err:10.3e}  ref absmax {np.abs(ref).max():9.3f}" + (f"  cnt {c.min():.0f}..{c.max():.0f}" if c is not None else ""))

emb = W("dec.embed"); pos = Fv("dec.pos").reshape(-1, d)
kc = [[] for _ in range(L)]; vc = [[] for _ in range(L)]
for it, tok in enumerate(toks):
    last = it == len(toks) - 1
    x = emb[tok] + pos[it]
    for l in range(L):
        p = f"dec.L{l}."
        base = xr + l * lw
        if last:
            report(f"L{l} qkv raw", val[base:base + 3 * d], W(p + "qkv.w").reshape(3 * d, d) @ x, cnt[base:base + 3 * d])
        qkv = W(p + "qkv.w").reshape(3 * d, d) @ ln(x) + Fv(p + "qkv.b")
        kc[l].append(bf(qkv[d:2 * d])); vc[l].append(bf(qkv[2 * d:]))
        K = np.stack(kc[l]).reshape(-1, H, 64); Vv = np.stack(vc[l]).reshape(-1, H, 64)
        q = qkv[:d].reshape(H, 64)
        ctx1 = np.zeros((H, 64), np.float32)
        for h in range(H):
            sc = K[:, h] @ q[h]
            pr = np.exp(sc - sc.max()); pr /= pr.sum()
            ctx1[h] = pr @ Vv[:, h]
        ctx1 = ctx1.reshape(-1)
        if last: report(f"L{l} ctx1", val[base + 3 * d:base + 4 * d], ctx1, cnt[base + 3 * d:base + 4 * d])
        x = x + W(p + "out.w").reshape(d, d) @ ctx1 + Fv(p + "out.b")
        if last: report(f"L{l} cq raw", val[base + 4 * d:base + 5 * d], W(p + "cq.w").reshape(d, d) @ x, cnt[base + 4 * d:base + 5 * d])
        q = (W(p + "cq.w").reshape(d, d) @ ln(x) + Fv(p + "cq.b")).reshape(H, 64)
        ctx2 = np.zeros((H, 64), np.float32)
        for h in range(H):
            sc = ck[l, h] @ q[h]
            pr = np.exp(sc - sc.max()); pr /= pr.sum()
            ctx2[h] = pr @ cv[l, h]
        if last: report(f"L{l} ctx2", val[base + 5 * d:base + 6 * d], ctx2.reshape(-1), cnt[base + 5 * d:base + 6 * d])
        x = x + W(p + "cout.w").reshape(d, d) @ ctx2.reshape(-1) + Fv(p + "cout.b")
        if last: report(f"L{l} fc1 raw", val[base + 6 * d:base + 6 * d + f], W(p + "fc1.w").reshape(f, d) @ x, cnt[base + 6 * d:base + 6 * d + f])
        hmid = W(p + "fc1.w").reshape(f, d) @ ln(x) + Fv(p + "fc1.b")
        hmid = 0.5 * hmid * (1 + np.vectorize(math.erf)(hmid / math.sqrt(2)))
        x = x + W(p + "fc2.w").reshape(d, f) @ hmid.astype(np.float32) + Fv(p + "fc2.b")
report("x final", val[:d], x, cnt[:d])
xn = ln(x) * Fv("dec.ln.g") + Fv("dec.ln.b")
lg = emb @ xn + Fv("dec.suppress_bias")
report("logits", logits[0], lg)
