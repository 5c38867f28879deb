"""ctypes binding of include/b200asr.h.  There is no fallback: if the shared
library is missing the import of the engine fails loudly."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("B200ASR_LIB", PKG / "libb200asr.so"))

OK, E_INVALID, E_CUDA, E_MISSING, E_NOGPU = 0, -1, -2, -3, -4
PRECISION_F32, PRECISION_BF16 = 0, 1
PCM_I16, PCM_F32 = 0, 1


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_mels", "d_model", "n_heads", "ffn", "enc_layers", "dec_layers", "vocab", "max_source", "max_target",
        "n_fft", "hop", "max_batch", "max_samples", "precision", "device", "use_tensor_cores")]


class NarConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "kind", "n_mels", "nfft", "win", "hop", "lfr_m", "lfr_n", "d_model", "n_heads", "ffn", "n_blocks0", "n_blocks",
        "n_tp_blocks", "vocab", "blank_id", "n_prompt", "n_lang", "fsmn_kernel", "max_batch", "max_samples", "precision",
        "device", "use_tensor_cores")] + [("ln_eps", C.c_float)] + [(n, C.c_int32) for n in (
        "dec_att_blocks", "dec_ffn_blocks", "dec_ffn", "cif_kernel")] + [("tail_threshold", C.c_float), ("dec_ln_eps", C.c_float)]


class QwenConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "device", "max_batch", "max_samples", "precision", "use_tensor_cores", "n_mels", "n_fft", "hop", "enc_layers", "enc_d",
        "enc_heads", "enc_ffn", "conv_ch", "out_dim", "chunks_per_window")] + [("enc_ln_eps", C.c_float)] + [(n, C.c_int32) for n in (
        "vocab", "hidden", "inter", "dec_layers", "heads", "kv_heads", "head_dim", "max_seq_len")] + [("rms_eps", C.c_float)]


# every symbol include/b200asr.h declares: (restype, argtypes)
_P = C.c_void_p
_I32P = C.POINTER(C.c_int32)
_F32P = C.POINTER(C.c_float)
SYMBOLS = {
    "b200asr_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "b200asr_destroy": (None, [_P]),
    "b200asr_last_error": (C.c_char_p, [_P]),
    "b200asr_set_tensor": (C.c_int, [_P, C.c_char_p, _F32P, C.c_int64]),
    "b200asr_finalize_weights": (C.c_int, [_P]),
    "b200asr_encode": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32]),
    "b200asr_upload_pcm": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32]),
    "b200asr_encode_resident": (C.c_int, [_P]),
    "b200asr_encode_ragged": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _I32P]),
    "b200asr_upload_pcm_ragged": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _I32P]),
    "b200asr_transcribe_ragged": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _I32P, _I32P, C.c_int32, C.c_int32, _I32P,
                                            C.c_int32, _I32P]),
    "b200asr_set_decode_options": (C.c_int, [_P, _I32P, C.c_int32, C.c_int32, C.c_float, C.c_int32]),
    "b200asr_set_sampling": (C.c_int, [_P, C.c_float, C.c_int32, C.c_float, C.c_float, C.c_uint64, _F32P, C.c_int32]),
    "b200asr_prefill": (C.c_int, [_P, _I32P, C.c_int32, _F32P, _I32P]),
    "b200asr_decode_step": (C.c_int, [_P, _I32P, _F32P, _I32P]),
    "b200asr_decode": (C.c_int, [_P, C.c_int32, _I32P, C.c_int32, _I32P]),
    "b200asr_no_speech_prob": (C.c_int, [_P, C.c_int32, _F32P]),
    "b200asr_transcribe": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _I32P, C.c_int32, C.c_int32, _I32P,
                                     C.c_int32, _I32P]),
    "b200asr_transcribe_resident": (C.c_int, [_P, _I32P, C.c_int32, C.c_int32, _I32P, C.c_int32, _I32P]),
    "b200asr_get_stage": (C.c_int, [_P, C.c_char_p, _F32P, C.c_int64, C.POINTER(C.c_int64)]),
    "b200asr_set_option": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "b200asr_stream": (_P, [_P]),
    "b200asr_synchronize": (C.c_int, [_P]),
    "b200asr_kernel_launches": (C.c_int64, [_P]),
    "b200asr_num_sms": (C.c_int, [_P]),
    "b200asr_nar_create": (C.c_int, [C.POINTER(NarConfig), C.POINTER(C.c_void_p)]),
    "b200asr_nar_destroy": (None, [C.c_void_p]),
    "b200asr_nar_last_error": (C.c_char_p, [C.c_void_p]),
    "b200asr_nar_set_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_float), C.c_int64]),
    "b200asr_nar_finalize_weights": (C.c_int, [C.c_void_p]),
    "b200asr_nar_run": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                                  C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32)]),
    "b200asr_nar_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "b200asr_nar_run_resident": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32)]),
    "b200asr_nar_run_ragged": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                                         C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32)]),
    "b200asr_nar_upload_ragged": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                                            C.POINTER(C.c_int32)]),
    "b200asr_nar_get_stage": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_float), C.c_int64, C.POINTER(C.c_int64)]),
    "b200asr_nar_kernel_launches": (C.c_int64, [C.c_void_p]),
    "b200asr_nar_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "b200asr_nar_stream": (C.c_void_p, [C.c_void_p]),
    "b200asr_qwen_create": (C.c_int, [C.POINTER(QwenConfig), C.POINTER(C.c_void_p)]),
    "b200asr_qwen_create_error": (C.c_char_p, []),
    "b200asr_qwen_destroy": (None, [C.c_void_p]),
    "b200asr_qwen_last_error": (C.c_char_p, [C.c_void_p]),
    "b200asr_qwen_set_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_float), C.c_int64]),
    "b200asr_qwen_finalize_weights": (C.c_int, [C.c_void_p]),
    "b200asr_qwen_set_prompt": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32), C.c_int32,
                                          C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32), C.c_int32]),
    "b200asr_qwen_set_decode_options": (C.c_int, [C.c_void_p, C.c_float, C.c_int32]),
    "b200asr_qwen_set_sampling": (C.c_int, [C.c_void_p, C.c_float, C.c_int32, C.c_float, C.c_float, C.c_uint64, C.POINTER(C.c_float), C.c_int32]),
    "b200asr_qwen_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32,
                                      C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32)]),
    "b200asr_qwen_prefill": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "b200asr_qwen_decode_step": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "b200asr_qwen_decode": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32)]),
    "b200asr_qwen_transcribe": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32,
                                          C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32,
                                          C.POINTER(C.c_int32)]),
    "b200asr_qwen_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "b200asr_qwen_transcribe_resident": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32), C.c_int32,
                                                   C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32)]),
    "b200asr_qwen_upload_ragged": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "b200asr_qwen_encode_ragged": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                                             C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32)]),
    "b200asr_qwen_transcribe_ragged": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                                                 C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_int32,
                                                 C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32)]),
    "b200asr_qwen_get_stage": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_float), C.c_int64, C.POINTER(C.c_int64)]),
    "b200asr_qwen_kernel_launches": (C.c_int64, [C.c_void_p]),
    "b200asr_qwen_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "b200asr_qwen_stream": (C.c_void_p, [C.c_void_p]),
    "b200asr_test_gemm": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _F32P, _F32P, _F32P, _F32P,
                                    C.c_int32, _F32P, C.c_char_p, C.c_int32]),
}

_lib = None


def load():
    """dlopen libb200asr.so and type every entry point.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
            "(nvcc, sm_100a).  b200asr has no CPU or PyTorch fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)           # AttributeError = header/library drift
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
