"""Audio-in -> token-ids-out driver for Whisper on the B200 engine: the host loop of the reference script
(/root/reference/Whisper/Inference_Whisper_ONNX.py:721-842) with the same knobs, the same protocol and the same
report line, minus ONNX Runtime.

Protocol per clip (reference defaults, :78-86): probe([SOT]) on the freshly encoded window -> language = argmax of
the raw SOT logits over the language token ids (:789-797) -> no-speech probability = softmax(logits with the -128
suppress bias undone)[nospeech] against NO_SPEECH_THRESHOLD (:798-805) -> prefill([SOT, lang, task, notimestamps])
-> greedy / penalty-greedy decode until a stop token or MAX_SEQ_LEN - prompt (:821-827).  Long audio is cut into
windows of the encoder's input length with stride SLIDING_WINDOW (:752-768); only window 0 is probed.

The decode loop itself runs on the device (b200asr_decode): the reference's per-token `.numpy()` round trip
(:645) does not exist here.
"""
from __future__ import annotations

import json
import time
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np

from .engine import WhisperEngine
from .ort_io import load_special_token_ids, load_supported_languages, resolve_supported_language


@dataclass
class InferenceOptions:
    """Module-level constants of the reference script (:71-100), same names, same defaults."""
    USE_SAMPLING: bool = False
    TEMPERATURE: float = 0.8
    TOP_K: int = 10
    TOP_P: float = 0.95
    SAMPLING_REPETITION_PENALTY: float = 1.0
    REPEAT_PENALTY: float = 0.8          # 1.0 selects greedy; another value selects penalty-greedy
    PENALTY_RANGE: int = 20
    REMOVE_REPEATED_PARTS: bool = False
    TARGET_LANGUAGE: str = "en"
    TASK: str = "transcribe"
    DETECT_LANGUAGE: bool = True
    NO_SPEECH_DETECTION: bool = True
    NO_SPEECH_THRESHOLD: float = 0.6
    SLIDING_WINDOW: int = 0
    USE_NORMALISE_AUDIO: bool = False
    INPUT_AUDIO_LENGTH: int = 0          # 0 = dynamic audio axis (Export_Whisper.py:743): one window = the whole clip
    SAMPLING_SEED: int = 0
    BATCH_WINDOWS: bool = True           # not a reference constant: windows after the first go through the engine as batches (same ids)


def prepare_audio_input(audio_int16: np.ndarray, target_dtype, *, audio_pcm_scale: int, target_rms: float = 4096.0,
                        use_normalise_audio: bool = False) -> np.ndarray:
    """int16 PCM -> the model's audio dtype (reference :103-126).  int16 models take the samples as they are (the
    1/32768 is folded into the STFT basis / applied on device); float models take x / audio_pcm_scale; optional RMS
    normalisation to 4096 with clipping to the int16 range."""
    target_dtype = np.dtype(target_dtype)
    if not use_normalise_audio and target_dtype == np.dtype(np.int16):
        return np.ascontiguousarray(audio_int16, dtype=target_dtype)
    audio = audio_int16.astype(np.float32)
    if use_normalise_audio:
        rms = np.sqrt(np.mean(audio * audio, dtype=np.float32), dtype=np.float32)
        if rms > 0:
            audio *= target_rms / (rms + 1e-7)
            np.clip(audio, -float(audio_pcm_scale), float(audio_pcm_scale) - 1.0, out=audio)
    if target_dtype == np.dtype(np.int16):
        return np.ascontiguousarray(audio, dtype=target_dtype)
    audio *= np.float32(1.0 / audio_pcm_scale)
    return np.ascontiguousarray(audio, dtype=target_dtype)


def remove_repeated_parts(ids, repeat_words_threshold: int, ids_len: int):
    """Cut the token list where a window of `repeat_words_threshold` ids repeats later on (reference :129-139)."""
    if ids_len <= repeat_words_threshold:
        return ids
    left = repeat_words_threshold // 2
    right = left + 1
    boundary = ids_len - left
    for i in range(left, boundary):
        for j in range(i + repeat_words_threshold, boundary):
            if all(ids[j + k] == ids[i + k] for k in range(-left, right)):
                return ids[:j - left]
    return ids


def plan_windows(audio_len: int, input_audio_length: int, sliding_window: int):
    """(windows, stride, aligned_length) exactly as the reference computes them (:752-760)."""
    stride = input_audio_length if sliding_window <= 0 else sliding_window
    if audio_len <= input_audio_length:
        windows = 1
    else:
        windows = int(np.ceil((audio_len - input_audio_length) / stride)) + 1
    return windows, stride, (windows - 1) * stride + input_audio_length


@dataclass
class ClipResult:
    tokens: List[int]
    language: str
    language_token: int
    no_speech_probability: Optional[float]
    no_speech: bool
    elapsed_s: float
    audio_s: float
    decode_steps: int
    windows: int = 1

    @property
    def rtf(self) -> float:
        return self.elapsed_s / self.audio_s if self.audio_s > 0 else float("nan")


class WhisperPipeline:
    """One engine + the metadata map of ASR_Metadata.onnx (keys audio_pcm_scale, max_seq_len, sample_rate,
    special_token_ids, supported_languages; reference :281-285)."""

    def __init__(self, engine: WhisperEngine, metadata: Dict[str, str], options: Optional[InferenceOptions] = None):
        self.engine = engine
        self.opt = options or InferenceOptions()
        self.audio_pcm_scale = int(metadata["audio_pcm_scale"])
        self.max_seq_len = int(metadata["max_seq_len"])
        self.sample_rate = int(metadata["sample_rate"])
        self.special = load_special_token_ids(metadata)
        self.languages = load_supported_languages(metadata)
        self.lang_token_to_code = {int(e["token_id"]): code for code, e in self.languages.items()}
        self.lang_token_ids = np.asarray(list(self.lang_token_to_code), dtype=np.int64)
        self.start_token = int(self.special["decoder_start"])
        self.task_token = int(self.special["tasks"][self.opt.TASK])
        stop = self.special["stop"]
        self.stop_tokens = set(stop if isinstance(stop, list) else [stop])
        self.no_timestamps = int(self.special["no_timestamps"])
        self.no_speech_token = self.special.get("no_speech")
        if self.max_seq_len != engine.dims.max_target:
            raise ValueError("metadata max_seq_len does not match the engine's max_target")
        if self.opt.USE_SAMPLING:
            self.strategy = "sampling"
        else:
            self.strategy = "penalty_greedy" if self.opt.REPEAT_PENALTY != 1.0 else "greedy"

    def _configure(self, generate_limit: int):
        o = self.opt
        if self.strategy == "sampling":
            self.engine.set_decode_options(stop_ids=sorted(self.stop_tokens), generate_limit=generate_limit)
            self.engine.set_sampling(temperature=o.TEMPERATURE, top_k=o.TOP_K, top_p=o.TOP_P,
                                     repetition_penalty=o.SAMPLING_REPETITION_PENALTY, seed=o.SAMPLING_SEED)
        else:
            self.engine.set_decode_options(stop_ids=sorted(self.stop_tokens), generate_limit=generate_limit,
                                           repeat_penalty=o.REPEAT_PENALTY if self.strategy == "penalty_greedy" else 1.0,
                                           penalty_range=o.PENALTY_RANGE)
            if hasattr(self.engine, "set_sampling"):
                self.engine.set_sampling(temperature=0.0)

    def transcribe_pcm(self, raw_audio: np.ndarray, language: Optional[str] = None, verbose: bool = False) -> ClipResult:
        o = self.opt
        raw_audio = np.asarray(raw_audio, dtype=np.int16).reshape(-1)
        audio_len = raw_audio.size
        language = language or o.TARGET_LANGUAGE
        language, entry = resolve_supported_language(self.languages, language)
        language_id = int(entry["token_id"])
        audio = prepare_audio_input(raw_audio.reshape(1, 1, -1), np.int16, audio_pcm_scale=self.audio_pcm_scale,
                                    use_normalise_audio=o.USE_NORMALISE_AUDIO)
        input_len = audio_len if o.INPUT_AUDIO_LENGTH <= 0 else o.INPUT_AUDIO_LENGTH
        windows, stride, aligned = plan_windows(audio_len, input_len, o.SLIDING_WINDOW)
        if audio.shape[-1] < aligned:
            padded = np.zeros((1, 1, aligned), dtype=audio.dtype)
            padded[..., :audio.shape[-1]] = audio
            audio = padded
        all_tokens: List[int] = []
        steps = 0
        no_speech = False
        prob = None
        t0 = time.time()
        max_b = int(getattr(self.engine, "max_batch", 1)) if o.BATCH_WINDOWS else 1
        w = 0
        while w < windows:
            window = audio[:, :, w * stride:w * stride + input_len]
            needs_probe = w == 0 and (o.DETECT_LANGUAGE or o.NO_SPEECH_DETECTION)
            prompt = [self.start_token, language_id, self.task_token, self.no_timestamps]
            generate_limit = max(0, self.max_seq_len - len(prompt))
            self._configure(generate_limit)
            if not needs_probe and max_b > 1 and windows - w > 1:
                # windows share no state (the reference re-runs encoder + prefill per window with the same prompt, :766-827), so the
                # remaining ones go through the engine as batches: same ids, window order kept
                nb = min(max_b, windows - w)
                batch = np.stack([audio[0, 0, (w + j) * stride:(w + j) * stride + input_len] for j in range(nb)])
                self.engine.encode(batch)
                self.engine.prefill(prompt, want_logits=False)
                for toks in self.engine.decode():
                    steps += max(0, len(toks) - 1)
                    all_tokens.extend(toks)
                w += nb
                continue
            self.engine.encode(window.reshape(1, -1))
            if needs_probe:
                logits, _ = self.engine.prefill([self.start_token])
                if o.DETECT_LANGUAGE:
                    row = logits.reshape(-1)
                    detected = int(self.lang_token_ids[np.argmax(row[self.lang_token_ids])])
                    language = self.lang_token_to_code.get(detected, language)
                    language_id = detected
                    if verbose:
                        print(f"Detected Language: {language}")
                if o.NO_SPEECH_DETECTION and self.no_speech_token is not None:
                    prob = float(self.engine.no_speech_prob(int(self.no_speech_token))[0])
                    if verbose:
                        print(f"No-Speech Probability: {prob:.3f}")
                    if prob >= o.NO_SPEECH_THRESHOLD:
                        no_speech = True
                        if verbose:
                            print("Audio classified as silence / non-speech; skipping transcription.")
                        break
                prompt = [self.start_token, language_id, self.task_token, self.no_timestamps]
            self.engine.prefill(prompt, want_logits=False)
            toks = self.engine.decode()[0]
            steps += max(0, len(toks) - 1)
            all_tokens.extend(toks)
            w += 1
        elapsed = time.time() - t0
        if o.REMOVE_REPEATED_PARTS and all_tokens:
            all_tokens = list(remove_repeated_parts(all_tokens, 3, len(all_tokens)))
        return ClipResult(tokens=all_tokens, language=language, language_token=language_id, no_speech_probability=prob,
                          no_speech=no_speech, elapsed_s=elapsed, audio_s=audio_len / self.sample_rate,
                          decode_steps=steps, windows=windows)

    def transcribe_batch(self, clips: Sequence[np.ndarray], language: Optional[str] = None) -> List[ClipResult]:
        """The same protocol for several clips of DIFFERENT lengths in one ragged batch (`lens=`: every clip keeps its
        single-clip semantics, so the ids are the ones `transcribe_pcm` gives clip by clip): one encoder launch, one
        probe prefill, per-clip language / no-speech decisions on the host, one prefill with per-clip prompts, one
        device-resident decode.  Single-window clips only (INPUT_AUDIO_LENGTH = 0, the dynamic audio axis); a clip
        classified as silence keeps an empty token list.  The reference has no batched driver (its graphs are batch 1,
        :722-768): this is the throughput path the ragged C-ABI entry points exist for."""
        o = self.opt
        if o.INPUT_AUDIO_LENGTH > 0:
            raise ValueError("transcribe_batch handles single-window clips (INPUT_AUDIO_LENGTH = 0); use transcribe_pcm for windows")
        raws = [np.asarray(c, dtype=np.int16).reshape(-1) for c in clips]
        B = len(raws)
        language = language or o.TARGET_LANGUAGE
        language, entry = resolve_supported_language(self.languages, language)
        prepared = [prepare_audio_input(r.reshape(1, 1, -1), np.int16, audio_pcm_scale=self.audio_pcm_scale,
                                        use_normalise_audio=o.USE_NORMALISE_AUDIO).reshape(-1) for r in raws]
        pcm, lens = WhisperEngine.pad_ragged(prepared)
        langs = [language] * B
        lang_ids = [int(entry["token_id"])] * B
        probs: List[Optional[float]] = [None] * B
        silent = [False] * B
        t0 = time.time()
        self._configure(max(0, self.max_seq_len - 4))
        self.engine.encode(pcm, lens=lens)
        if o.DETECT_LANGUAGE or o.NO_SPEECH_DETECTION:
            logits, _ = self.engine.prefill([[self.start_token]] * B)
            if o.DETECT_LANGUAGE:
                for b in range(B):
                    detected = int(self.lang_token_ids[np.argmax(logits[b][self.lang_token_ids])])
                    langs[b] = self.lang_token_to_code.get(detected, langs[b])
                    lang_ids[b] = detected
            if o.NO_SPEECH_DETECTION and self.no_speech_token is not None:
                pr = self.engine.no_speech_prob(int(self.no_speech_token))
                for b in range(B):
                    probs[b] = float(pr[b])
                    silent[b] = probs[b] >= o.NO_SPEECH_THRESHOLD
        prompts = [[self.start_token, lang_ids[b], self.task_token, self.no_timestamps] for b in range(B)]
        self.engine.prefill(prompts, want_logits=False)
        toks = self.engine.decode()
        elapsed = time.time() - t0
        out = []
        for b in range(B):
            t = [] if silent[b] else list(toks[b])
            if o.REMOVE_REPEATED_PARTS and t:
                t = list(remove_repeated_parts(t, 3, len(t)))
            out.append(ClipResult(tokens=t, language=langs[b], language_token=lang_ids[b], no_speech_probability=probs[b],
                                  no_speech=silent[b], elapsed_s=elapsed / B, audio_s=raws[b].size / self.sample_rate,
                                  decode_steps=max(0, len(t) - 1), windows=1))
        return out

    def report(self, res: ClipResult, text: str) -> str:
        """The reference's result block (:836-841)."""
        body = "[no speech detected]" if res.no_speech else text
        return (f"\nASR Result:\n{body}\n\nRTF: {res.rtf:.3f}   ({res.elapsed_s:.3f}s for {res.audio_s:.2f}s audio, "
                f"{len(res.tokens)} tokens; device-resident decode; 1 launch/clip)")


def load_metadata(folder: Path) -> Dict[str, str]:
    """`ASR_Metadata.json` in the model folder: the custom_metadata_map of the reference's ASR_Metadata.onnx."""
    return {k: (v if isinstance(v, str) else json.dumps(v)) for k, v in json.loads((Path(folder) / "ASR_Metadata.json").read_text()).items()}
