"""ORT-shaped facade over the engine for the Whisper driver script.

The reference script talks to three merged graphs through ``onnxruntime.InferenceSession`` +
``io_binding()`` (/root/reference/Whisper/Inference_Whisper_ONNX.py:247-248, call sites :489, :549, :640, :697).
This module provides objects with the same surface and the same tensor names
(/root/reference/Whisper/Shared_Merged.py:864-921: inputs ``in_de_key_layer_i`` / ``in_de_value_layer_i``,
``[audio]``, ``[en_key_layer_i, en_value_layer_i]``, ``embed_input_ids``, ``prefill_ids_len``,
``prefill_history_len`` / ``decode_kv_seq_len``; outputs ``out_de_*``, ``[encoder_en_*]``,
``argmax_max_logits_idx`` or ``greedy_max_logits_idx`` + ``greedy_save_id_out``, ``[logits]``,
``prefill_kv_seq_len`` / ``decode_kv_seq_len_next``) so `_plan_merged_io` (:323-392) and the probe / prefill /
decode loops run unchanged against a B200 engine instead of ORT.

State tensors (self-KV, cross-KV) stay resident in HBM inside the engine: the OrtValues handed back for them are
handles that are only materialised by ``.numpy()``.  Feeding a handle of another engine or of an older encode is a
``ValueError`` (ORT would compute on stale data silently; this shim refuses).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from .engine import WhisperEngine

_TYPE_OF = {np.dtype(np.float32): "tensor(float)", np.dtype(np.int16): "tensor(int16)", np.dtype(np.int32): "tensor(int32)",
            np.dtype(np.int64): "tensor(int64)", np.dtype(np.float16): "tensor(float16)"}


@dataclass
class NodeArg:
    name: str
    shape: list
    type: str


class ModelMeta:
    def __init__(self, custom: Dict[str, str]):
        self.custom_metadata_map = dict(custom)


class OrtDevice:
    """Stand-in for onnxruntime.capi._pybind_state.OrtDevice (only identity matters here)."""
    def __init__(self, *a, **k):
        pass

    @staticmethod
    def cpu():
        return 0

    @staticmethod
    def cuda():
        return 1

    @staticmethod
    def default_memory():
        return 0


class OrtValue:
    """Host array or a handle to an engine-resident tensor."""

    def __init__(self, array: Optional[np.ndarray] = None, *, engine: Optional[WhisperEngine] = None, stage: str = "",
                 index: int = 0, epoch: int = -1, meta: Optional[NodeArg] = None):
        self._array = array
        self._engine = engine
        self._stage = stage
        self._index = index
        self._epoch = epoch
        self._meta = meta

    @staticmethod
    def ortvalue_from_numpy(array, device_type="cpu", device_id=0):
        return OrtValue(np.ascontiguousarray(array))

    def is_resident(self) -> bool:
        return self._engine is not None

    def numpy(self) -> np.ndarray:
        if self._array is not None:
            return self._array
        eng, d = self._engine, self._engine.dims
        H, dh = d.n_heads, d.head_dim
        if self._stage in ("self_k", "self_v"):
            L = d.dec_layers
            flat = eng.get_stage(self._stage, L * eng.batch * H * d.max_target * dh)
            kv = flat.reshape(L, eng.batch, H, -1, dh)[self._index]
            return np.ascontiguousarray(kv.transpose(0, 1, 3, 2)) if self._stage == "self_k" else kv    # K is (B,H,dh,kv)
        T = eng.T_enc
        flat = eng.get_stage(self._stage, eng.batch * d.dec_layers * H * T * dh).reshape(eng.batch, d.dec_layers, H, T, dh)
        kv = flat[0, self._index]
        return np.ascontiguousarray(kv.transpose(0, 2, 1)) if self._stage == "cross_k" else kv                # K is (H,dh,T)

    def update_inplace(self, array):
        if self._array is None:
            raise ValueError("update_inplace on an engine-resident value")
        np.copyto(self._array, np.asarray(array, dtype=self._array.dtype).reshape(self._array.shape))

    def shape(self):
        return list(self.numpy().shape)


class IOBinding:
    def __init__(self, session: "InferenceSession"):
        self._session = session
        self._inputs: Dict[str, OrtValue] = {}
        self._bound_outputs: List[str] = []
        self._outputs: List[OrtValue] = []
        self._iobinding = self          # the script reaches through binding._iobinding.bind_output(name, device)

    def bind_ortvalue_input(self, name: str, value: OrtValue):
        self._session._check_input(name)
        self._inputs[name] = value

    def bind_cpu_input(self, name: str, array):
        self.bind_ortvalue_input(name, OrtValue(np.ascontiguousarray(array)))

    def bind_output(self, name: str, device=None):
        self._session._check_output(name)
        if name not in self._bound_outputs:
            self._bound_outputs.append(name)

    def bind_ortvalue_output(self, name: str, value: OrtValue):
        self.bind_output(name)

    def clear_binding_outputs(self):
        self._bound_outputs = []
        self._outputs = []

    def clear_binding_inputs(self):
        self._inputs = {}

    def get_outputs(self) -> List[OrtValue]:
        return list(self._outputs)


class InferenceSession:
    """One of the reference's merged graphs: kind in {"probe", "prefill", "decode", "no_speech"}."""

    def __init__(self, kind: str, owner: "WhisperSessions"):
        self.kind = kind
        self._o = owner
        self._inputs, self._outputs = owner._signature(kind)

    # -- ORT surface ---------------------------------------------------------
    def get_inputs(self) -> List[NodeArg]:
        return list(self._inputs)

    def get_outputs(self) -> List[NodeArg]:
        return list(self._outputs)

    def get_modelmeta(self) -> ModelMeta:
        return ModelMeta(self._o.metadata)

    def get_providers(self) -> List[str]:
        return ["B200ExecutionProvider"]

    def io_binding(self) -> IOBinding:
        return IOBinding(self)

    def _check_input(self, name):
        if name not in {m.name for m in self._inputs}:
            raise ValueError(f"{self.kind}: no input named {name!r}")

    def _check_output(self, name):
        if name not in {m.name for m in self._outputs}:
            raise ValueError(f"{self.kind}: no output named {name!r}")

    def run_with_iobinding(self, binding: IOBinding, run_options=None):
        missing = [m.name for m in self._inputs if m.name not in binding._inputs]
        if missing:
            raise ValueError(f"{self.kind}: unbound inputs {missing[:4]}{'...' if len(missing) > 4 else ''}")
        produced = self._o._run(self.kind, binding._inputs)
        names = binding._bound_outputs or [m.name for m in self._outputs]
        order = [m.name for m in self._outputs if m.name in names]       # graph output order, as ORT returns them
        binding._outputs = [produced[n] for n in order]

    def run(self, output_names: Optional[Sequence[str]], feeds: Dict[str, Any]):
        b = self.io_binding()
        for k, v in feeds.items():
            b.bind_cpu_input(k, v)
        self.run_with_iobinding(b)
        by_name = dict(zip([m.name for m in self._outputs], b.get_outputs()))
        names = output_names or [m.name for m in self._outputs]
        return [by_name[n].numpy() for n in names]


class WhisperSessions:
    """The PROBE / PREFILL / DECODE / NO_SPEECH session set of the driver script on one engine (batch 1, like the
    reference graphs: Whisper/Export_Whisper.py:432)."""

    def __init__(self, engine: WhisperEngine, metadata: Dict[str, str], *, strategy: str = "greedy",
                 no_speech_token: Optional[int] = None, audio_dtype=np.int16, repeat_penalty: float = 1.0,
                 penalty_range: int = 20, stop_ids: Sequence[int] = (), generate_limit: int = 0):
        if strategy not in ("greedy", "penalty_greedy"):
            raise ValueError("strategy must be 'greedy' or 'penalty_greedy'")
        if strategy == "greedy" and repeat_penalty != 1.0:
            raise ValueError("strategy 'greedy' has no penalty head; use strategy='penalty_greedy'")
        # the penalty head of the merged graph takes its value / range as bound inputs every step
        # (Inference_Whisper_ONNX.py:629-633); the engine applies them inside the decode kernel, so they are configured here
        # and the bound tensors are checked against them in _run
        self.repeat_penalty = float(repeat_penalty) if strategy == "penalty_greedy" else 1.0
        self.penalty_range = int(penalty_range)
        engine.set_decode_options(stop_ids=list(stop_ids), generate_limit=generate_limit, repeat_penalty=self.repeat_penalty,
                                  penalty_range=self.penalty_range)
        self.engine = engine
        self.metadata = dict(metadata)
        self.strategy = strategy
        self.no_speech_token = no_speech_token
        self.audio_dtype = np.dtype(audio_dtype)
        self.epoch = 0                      # bumps at every encode: older cross-KV handles become stale
        self.kv_epoch = 0                   # bumps at every decoder launch: older self-KV handles become stale
        self.kv_len = 0
        self.probe = InferenceSession("probe", self)
        self.prefill = InferenceSession("prefill", self)
        self.decode = InferenceSession("decode", self)
        self.no_speech = InferenceSession("no_speech", self)

    # -- graph signatures (names and order of Shared_Merged.py:864-921) ----------
    def _signature(self, kind):
        d = self.engine.dims
        L, H, dh, V = d.dec_layers, d.n_heads, d.head_dim, d.vocab
        f = "tensor(float)"
        if kind == "no_speech":
            return [NodeArg("logits", [1, V], f)], [NodeArg("no_speech_probability", [1], f)]
        ins = [NodeArg(f"in_de_key_layer_{i}", ["batch", H, dh, "history_len"], f) for i in range(L)]
        ins += [NodeArg(f"in_de_value_layer_{i}", ["batch", H, "history_len", dh], f) for i in range(L)]
        outs = [NodeArg(f"out_de_key_layer_{i}", ["batch", H, dh, "kv_seq_len"], f) for i in range(L)]
        outs += [NodeArg(f"out_de_value_layer_{i}", ["batch", H, "kv_seq_len", dh], f) for i in range(L)]
        if kind == "probe":
            ins.append(NodeArg("audio", [1, 1, "audio_len"], _TYPE_OF[self.audio_dtype]))
            outs += [NodeArg(f"encoder_en_key_layer_{i}", [H, dh, "signal_len"], f) for i in range(L)]
            outs += [NodeArg(f"encoder_en_value_layer_{i}", [H, "signal_len", dh], f) for i in range(L)]
        else:
            ins += [NodeArg(f"en_key_layer_{i}", [H, dh, "signal_len"], f) for i in range(L)]
            ins += [NodeArg(f"en_value_layer_{i}", [H, "signal_len", dh], f) for i in range(L)]
        ins.append(NodeArg("embed_input_ids", ["batch", "ids_len"], "tensor(int32)"))
        pen = self.strategy == "penalty_greedy"
        head = "greedy_max_logits_idx" if pen else "argmax_max_logits_idx"
        if kind == "decode":
            ins.append(NodeArg("decode_kv_seq_len", [1], "tensor(int64)"))
            if pen:
                ins += [NodeArg("penalty_save_id_in", ["batch", "history"], "tensor(int32)"),
                        NodeArg("greedy_save_id_in", ["batch", "history"], "tensor(int32)"),
                        NodeArg("penalty_penalty_value", [1], f), NodeArg("penalty_penalty_range", [1], "tensor(int64)")]
            outs.append(NodeArg(head, ["batch", 1], "tensor(int32)"))
            if pen:
                outs.append(NodeArg("greedy_save_id_out", ["batch", "history_next"], "tensor(int32)"))
            outs.append(NodeArg("decode_kv_seq_len_next", [1], "tensor(int64)"))
        else:
            ins += [NodeArg("prefill_ids_len", [1], "tensor(int64)"), NodeArg("prefill_history_len", [1], "tensor(int64)")]
            if pen:
                ins.append(NodeArg("greedy_save_id_in", ["batch", "history"], "tensor(int32)"))
            outs.append(NodeArg(head, ["batch", 1], "tensor(int32)"))
            if pen:
                outs.append(NodeArg("greedy_save_id_out", ["batch", "history_next"], "tensor(int32)"))
            outs += [NodeArg("logits", ["batch", V], f), NodeArg("prefill_kv_seq_len", [1], "tensor(int64)")]
        return ins, outs

    # -- execution -----------------------------------------------------------------
    def _state_handles(self):
        L = self.engine.dims.dec_layers
        out = {}
        for i in range(L):
            out[f"out_de_key_layer_{i}"] = OrtValue(engine=self.engine, stage="self_k", index=i, epoch=self.kv_epoch)
            out[f"out_de_value_layer_{i}"] = OrtValue(engine=self.engine, stage="self_v", index=i, epoch=self.kv_epoch)
        return out

    def _check_cross(self, feeds):
        for i in range(self.engine.dims.dec_layers):
            for n in (f"en_key_layer_{i}", f"en_value_layer_{i}"):
                v = feeds[n]
                if not v.is_resident() or v._engine is not self.engine or v._epoch != self.epoch:
                    raise ValueError(f"{n}: not the cross-KV of this engine's latest encode")

    def _head_outputs(self, first_or_tok, kind):
        pen = self.strategy == "penalty_greedy"
        out = {("greedy_max_logits_idx" if pen else "argmax_max_logits_idx"):
               OrtValue(np.asarray(first_or_tok, np.int32).reshape(1, 1))}
        if pen:
            sel = self.engine.get_stage("selected", self.engine.dims.max_target).astype(np.int32)
            out["greedy_save_id_out"] = OrtValue(sel.reshape(1, -1))
        return out

    def _run(self, kind, feeds: Dict[str, OrtValue]) -> Dict[str, OrtValue]:
        eng = self.engine
        L = eng.dims.dec_layers
        if kind == "no_speech":
            if self.no_speech_token is None:
                raise ValueError("no_speech_token not configured")
            return {"no_speech_probability": OrtValue(eng.no_speech_prob(self.no_speech_token).astype(np.float32))}
        ids = feeds["embed_input_ids"].numpy().astype(np.int32)
        if kind in ("probe", "prefill"):
            for i in range(L):      # the merged prefill graphs start from empty self-KV (Inference_Whisper_ONNX.py:441-451)
                v = feeds[f"in_de_key_layer_{i}"]
                if v.is_resident() or v.numpy().shape[-1] != 0:
                    raise ValueError("prefill expects empty self-KV inputs (history_len = 0)")
            if int(feeds["prefill_history_len"].numpy().reshape(-1)[0]) != 0:
                raise ValueError("prefill_history_len must be 0")
            if int(feeds["prefill_ids_len"].numpy().reshape(-1)[0]) != ids.shape[-1]:
                raise ValueError("prefill_ids_len does not match embed_input_ids")
            out = {}
            if kind == "probe":
                audio = feeds["audio"].numpy()
                eng.encode(audio.reshape(1, -1) if audio.dtype == np.int16 else audio.reshape(1, -1).astype(np.float32))
                self.epoch += 1
            else:
                self._check_cross(feeds)
            logits, first = eng.prefill(ids.reshape(1, -1))
            self.kv_epoch += 1
            self.kv_len = ids.shape[-1]
            out.update(self._state_handles())
            if kind == "probe":
                for i in range(L):
                    out[f"encoder_en_key_layer_{i}"] = OrtValue(engine=eng, stage="cross_k", index=i, epoch=self.epoch)
                    out[f"encoder_en_value_layer_{i}"] = OrtValue(engine=eng, stage="cross_v", index=i, epoch=self.epoch)
            out.update(self._head_outputs(first, kind))
            out["logits"] = OrtValue(logits.reshape(1, -1))
            out["prefill_kv_seq_len"] = OrtValue(np.asarray([self.kv_len], np.int64))
            return out
        # decode
        self._check_cross(feeds)
        for i in range(L):
            for n in (f"in_de_key_layer_{i}", f"in_de_value_layer_{i}"):
                v = feeds[n]
                if not v.is_resident() or v._engine is not eng or v._epoch != self.kv_epoch:
                    raise ValueError(f"{n}: not the self-KV produced by this engine's previous launch")
        if int(feeds["decode_kv_seq_len"].numpy().reshape(-1)[0]) != self.kv_len:
            raise ValueError("decode_kv_seq_len does not match the resident cache length")
        if self.strategy == "penalty_greedy":
            # the driver binds penalty_on (value, range) once generated >= range and penalty_off (1.0) before that; the engine
            # switches at the same step, so either the configured value or the neutral 1.0 is acceptable -- anything else
            # would silently not be applied
            pv = float(feeds["penalty_penalty_value"].numpy().reshape(-1)[0])
            pr = int(feeds["penalty_penalty_range"].numpy().reshape(-1)[0])
            if not (abs(pv - 1.0) < 1e-6 or abs(pv - self.repeat_penalty) < 1e-6):
                raise ValueError(f"penalty_penalty_value {pv} differs from the configured repeat_penalty {self.repeat_penalty}")
            if abs(pv - 1.0) >= 1e-6 and pr != self.penalty_range:
                raise ValueError(f"penalty_penalty_range {pr} differs from the configured penalty_range {self.penalty_range}")
        _, tok = eng.decode_step(token_in=ids.reshape(-1), want_logits=False)
        self.kv_epoch += 1
        self.kv_len += 1
        out = self._state_handles()
        out.update(self._head_outputs(tok, kind))
        out["decode_kv_seq_len_next"] = OrtValue(np.asarray([self.kv_len], np.int64))
        return out


class NarSessions:
    """The single graph of the SenseVoice / Paraformer drivers as an ORT-shaped session (`.session`): inputs `audio`
    [1, 1, audio_len] (+ `language_idx` [1] for SenseVoice), outputs `token_ids` [num_token] and `num_id` [1]
    (SenseVoice/Export_SenseVoice.py:375-376, Paraformer/Non-Streaming/Export_Paraformer.py:600-601), driven the way the
    scripts drive it (Inference_SenseVoice_ONNX.py:284-305: bind_cpu_input / bind_ortvalue_input, `_iobinding.bind_output`,
    `run_with_iobinding`, `get_outputs()[0].numpy()`).  `engine` is a SenseVoiceEngine or ParaformerEngine."""

    def __init__(self, engine, metadata: Optional[Dict[str, str]] = None, *, audio_dtype=np.int16):
        from .paraformer import ParaformerEngine
        self.engine = engine
        self.metadata = dict(metadata or {})
        self.audio_dtype = np.dtype(audio_dtype)
        self.has_language = not isinstance(engine, ParaformerEngine)
        self.session = InferenceSession("nar", self)

    def _signature(self, kind):
        ins = [NodeArg("audio", [1, 1, "audio_len"], _TYPE_OF[self.audio_dtype])]
        if self.has_language:
            ins.append(NodeArg("language_idx", [1], "tensor(int32)"))
        outs = [NodeArg("token_ids", ["num_token"], "tensor(int32)"), NodeArg("num_id", [1], "tensor(int32)")]
        return ins, outs

    def _run(self, kind, feeds: Dict[str, OrtValue]) -> Dict[str, OrtValue]:
        audio = feeds["audio"].numpy()
        if audio.ndim != 3 or audio.shape[0] != 1 or audio.shape[1] != 1:
            raise ValueError(f"audio must be [1, 1, audio_len], got {list(audio.shape)}")
        if audio.dtype != self.audio_dtype:
            raise ValueError(f"audio must be {_TYPE_OF[self.audio_dtype]}, got {audio.dtype}")
        lang = 0
        if self.has_language:
            li = feeds["language_idx"].numpy()
            if li.shape != (1,) or li.dtype != np.int32:
                raise ValueError("language_idx must be int32 [1]")
            lang = int(li[0])
        toks = self.engine.run(np.ascontiguousarray(audio.reshape(1, -1)), lang)[0]
        return {"token_ids": OrtValue(np.asarray(toks, dtype=np.int32)), "num_id": OrtValue(np.asarray([len(toks)], dtype=np.int32))}

