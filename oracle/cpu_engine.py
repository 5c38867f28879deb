"""TEST INFRASTRUCTURE ONLY (never imported by the product package): a CPU stand-in with the surface of
`b200asr.engine.WhisperEngine` that `b200asr.session.WhisperSessions` uses, backed by the oracle
(oracle/whisper_oracle.py).  It exists so that the reference driver's OWN functions -- AST-extracted from
/root/reference/Whisper/Inference_Whisper_ONNX.py by oracle/gen_script_golden.py -- can be run, unmodified, against
the session facade in this container (no GPU here), minting the token streams the GPU engine is then held to behind
the very same facade (tests/test_gpu_script_goldens.py)."""
from __future__ import annotations

import numpy as np
import torch

from oracle import whisper_oracle as wo


class OracleWhisperEngine:
    def __init__(self, dims, fw, suppress_tokens):
        self.dims = dims                          # b200asr.config dims (n_heads, head_dim, dec_layers, vocab, max_target, hop, ...)
        self.od = wo.WhisperDims(**dims.to_dict())
        self.fw = fw
        self.suppress = list(suppress_tokens)
        self.batch = 1
        self.T_enc = 0
        self.stop, self.limit_cfg, self.pen, self.pen_range = set(), 0, 1.0, 20
        self._reset()

    def _reset(self):
        self.sk, self.sv = wo.empty_self_kv(self.od)
        self.kv_len = 0
        self.save_id = torch.zeros(1, 0, dtype=torch.int32)
        self.generated = 0
        self.logits = None

    def set_decode_options(self, stop_ids=(), generate_limit=0, repeat_penalty=1.0, penalty_range=20):
        self.stop, self.limit_cfg, self.pen, self.pen_range = set(int(s) for s in stop_ids), int(generate_limit), float(repeat_penalty), int(penalty_range)

    def encode(self, pcm):
        pcm = np.asarray(pcm).reshape(-1)
        audio = wo.prepare_audio(pcm) if pcm.dtype == np.int16 else torch.from_numpy(np.ascontiguousarray(pcm, np.float32)).reshape(1, 1, -1)
        with torch.no_grad():
            self.ck, self.cv, _ = wo.encoder(audio, self.fw, self.od)
        self.T_enc = (pcm.shape[0] // self.od.hop + 1) // 2

    def prefill(self, ids, want_logits=True):
        self._reset()
        ids = torch.as_tensor(np.asarray(ids, np.int32).reshape(1, -1))
        with torch.no_grad():
            self.sk, self.sv, logits = wo.decoder(ids, 0, self.sk, self.sv, self.ck, self.cv, self.fw, self.od)
        self.kv_len = ids.shape[-1]
        self.logits = logits
        sel = int(wo.argmax_head(wo.begin_suppress(logits, self.fw))[0, 0])
        self.save_id = torch.cat([self.save_id, torch.tensor([[sel]], dtype=torch.int32)], dim=-1)
        if sel not in self.stop:
            self.generated = 1
        return logits.numpy().copy(), np.asarray([sel], np.int32)

    def decode_step(self, token_in=None, want_logits=True):
        feed = int(self.save_id[0, -1]) if token_in is None else int(np.asarray(token_in).reshape(-1)[0])
        with torch.no_grad():
            self.sk, self.sv, logits = wo.decoder(torch.tensor([[feed]], dtype=torch.int32), self.kv_len, self.sk, self.sv,
                                                  self.ck, self.cv, self.fw, self.od)
        self.kv_len += 1
        self.logits = logits
        head = logits
        if self.pen != 1.0:
            head = wo.apply_penalty(logits, self.save_id, self.pen if self.generated >= self.pen_range else 1.0, self.pen_range)
        sel = int(wo.argmax_head(head)[0, 0])
        self.save_id = torch.cat([self.save_id, torch.tensor([[sel]], dtype=torch.int32)], dim=-1)
        if sel not in self.stop:
            self.generated += 1
        return (logits.numpy().copy() if want_logits else None), np.asarray([sel], np.int32)

    def no_speech_prob(self, token):
        return wo.no_speech_prob(self.logits, self.suppress, int(token)).numpy().reshape(-1)

    def get_stage(self, name, capacity):
        if name == "selected":
            return self.save_id.numpy().reshape(-1).astype(np.float32)
        raise NotImplementedError(name)
