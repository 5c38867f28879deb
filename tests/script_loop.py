"""Test helper: the per-clip protocol of the reference driver (probe -> language arg-max -> no-speech -> prefill -> decode,
Inference_Whisper_ONNX.py:437-663,766-823) restated over the ORT-shaped facade (b200asr.session), for boxes without the
reference checkout.  tests/test_script_goldens_cpu.py holds this restatement to the reference's OWN functions (run by
oracle/gen_script_golden.py against the same facade); tests/test_gpu_script_goldens.py then drives the CUDA engine with it."""
import numpy as np

from b200asr.ort_io import array_for, filled_for, metadata_by_name, scalar_for
from b200asr.session import OrtValue


def _val(a):
    return OrtValue.ortvalue_from_numpy(np.ascontiguousarray(a))


def _bind_prefill_like(sess, binding, ids, strategy):
    meta = metadata_by_name(sess.get_inputs())
    for name in meta:
        if name.startswith("in_de_"):
            axis = 3 if "key" in name else 2
            binding.bind_ortvalue_input(name, _val(filled_for(meta[name], axes={0: 1, axis: 0})))
    binding.bind_ortvalue_input("embed_input_ids", _val(array_for(meta["embed_input_ids"], ids, axes={0: 1, 1: len(ids[0])})))
    binding.bind_ortvalue_input("prefill_ids_len", _val(scalar_for(meta["prefill_ids_len"], len(ids[0]))))
    binding.bind_ortvalue_input("prefill_history_len", _val(scalar_for(meta["prefill_history_len"], 0)))
    if strategy == "penalty_greedy":
        binding.bind_ortvalue_input("greedy_save_id_in", _val(filled_for(meta["greedy_save_id_in"], axes={0: 1, 1: 0})))
    for m in sess.get_outputs():
        binding._iobinding.bind_output(m.name, None)


def drive(S, pcm_i16, prompt, *, strategy, repeat_penalty, penalty_range, detect_language, no_speech, language_token_ids,
          stop_tokens, max_seq_len, threshold=2.0):
    head = "greedy_max_logits_idx" if strategy == "penalty_greedy" else "argmax_max_logits_idx"
    start, lang_id, task, nots = prompt
    out = {"detected_language_token": None, "no_speech_probability": None}
    needs_probe = detect_language or no_speech
    # ---- probe (encoder + first prefill) ----
    b = S.probe.io_binding()
    pmeta = metadata_by_name(S.probe.get_inputs())
    audio = _val(filled_for(pmeta["audio"], axes={0: 1, 1: 1, 2: len(pcm_i16)}))
    audio.update_inplace(array_for(pmeta["audio"], np.asarray(pcm_i16).reshape(1, 1, -1), axes={0: 1, 1: 1, 2: len(pcm_i16)}))
    b.bind_ortvalue_input("audio", audio)
    _bind_prefill_like(S.probe, b, [[start]] if needs_probe else [[start, lang_id, task, nots]], strategy)
    S.probe.run_with_iobinding(b)
    po = dict(zip([m.name for m in S.probe.get_outputs()], b.get_outputs()))
    cross = {n.replace("encoder_", ""): v for n, v in po.items() if n.startswith("encoder_en_")}
    pre = po
    if needs_probe:
        if detect_language:
            lg = po["logits"].numpy().reshape(-1)
            ids = np.asarray(language_token_ids, np.int64)
            lang_id = int(ids[np.argmax(lg[ids])])
            out["detected_language_token"] = lang_id
        if no_speech:
            nb = S.no_speech.io_binding()
            nb.bind_ortvalue_input("logits", po["logits"])
            S.no_speech.run_with_iobinding(nb)
            p = float(nb.get_outputs()[0].numpy().reshape(-1)[0])
            out["no_speech_probability"] = p
            if p >= threshold:
                out["tokens"] = []
                return out
        b = S.prefill.io_binding()
        for n, v in cross.items():
            b.bind_ortvalue_input(n, v)
        _bind_prefill_like(S.prefill, b, [[start, lang_id, task, nots]], strategy)
        S.prefill.run_with_iobinding(b)
        pre = dict(zip([m.name for m in S.prefill.get_outputs()], b.get_outputs()))
    # ---- decode loop (:584-663) ----
    limit = max(0, max_seq_len - 4)
    stop = set(stop_tokens)
    out_pre = [m.name for m in S.prefill.get_outputs()]
    state = [pre[n] for n in out_pre if n.startswith("out_de_")]
    next_token, kv_len = pre[head], pre["prefill_kv_seq_len"]
    selected = int(pre[head].numpy().reshape(-1)[0])
    saved = pre.get("greedy_save_id_out") if strategy == "penalty_greedy" else None
    host, generated = [], 0
    if selected not in stop and limit > 0:
        generated = 1
        if saved is None:
            host.append(selected)
    dmeta = metadata_by_name(S.decode.get_inputs())
    out_dec = [m.name for m in S.decode.get_outputs()]
    bindings = [S.decode.io_binding(), S.decode.io_binding()]
    steps = 0
    while generated < limit and selected not in stop:
        b = bindings[steps & 1]
        for n, v in cross.items():
            b.bind_ortvalue_input(n, v)
        b.bind_ortvalue_input("embed_input_ids", next_token)
        b.bind_ortvalue_input("decode_kv_seq_len", kv_len)
        for n, v in zip([m for m in dmeta if m.startswith("in_de_")], state):
            b.bind_ortvalue_input(n, v)
        if strategy == "penalty_greedy":
            b.bind_ortvalue_input("penalty_save_id_in", saved)
            b.bind_ortvalue_input("greedy_save_id_in", saved)
            b.bind_ortvalue_input("penalty_penalty_range", _val(scalar_for(dmeta["penalty_penalty_range"], penalty_range)))
            b.bind_ortvalue_input("penalty_penalty_value",
                                  _val(scalar_for(dmeta["penalty_penalty_value"], repeat_penalty if generated >= penalty_range else 1.0)))
        b.clear_binding_outputs()
        for n in out_dec:
            b._iobinding.bind_output(n, None)
        S.decode.run_with_iobinding(b)
        o = dict(zip(out_dec, b.get_outputs()))
        state = [o[n] for n in out_dec if n.startswith("out_de_")]
        next_token, kv_len = o[head], o["decode_kv_seq_len_next"]
        selected = int(o[head].numpy().reshape(-1)[0])
        if strategy == "penalty_greedy":
            saved = o["greedy_save_id_out"]
        if selected not in stop:
            generated += 1
            if saved is None:
                host.append(selected)
        steps += 1
    if saved is not None:
        host = []
        for t in saved.numpy()[0]:
            t = int(t)
            if t in stop or len(host) >= limit:
                break
            host.append(t)
    out["tokens"] = host
    out["decode_steps"] = steps
    return out


def drive_case(S, g, cfg, meta):
    prompt = [int(t) for t in g["prompt"].reshape(-1)]
    kw = dict(strategy=cfg["strategy"], repeat_penalty=cfg["repeat_penalty"], penalty_range=cfg["penalty_range"],
              detect_language=cfg["detect_language"], no_speech=cfg["no_speech"], language_token_ids=meta["language_token_ids"],
              max_seq_len=meta["max_seq_len"])
    res = drive(S, g["pcm"], prompt, stop_tokens=[], **kw)
    if cfg["stop_at"] is not None:
        stop = res["tokens"][cfg["stop_at"]]
        res2 = drive(S, g["pcm"], prompt, stop_tokens=[stop], **kw)
        res = {**res2, "stop_token": int(stop), "free_tokens": res["tokens"]}
    return res
