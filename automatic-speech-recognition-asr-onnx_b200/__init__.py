"""b200asr: B200-native (sm_100a) replacement for the ONNX Runtime hot path of
DakeQQ/Automatic-Speech-Recognition-ASR-ONNX (Whisper family first).

Only the pieces the hot path needs live here: ``csrc/`` (CUDA kernels + the
C ABI declared in ``include/b200asr.h``), the ctypes binding, the weight
folder/packer, and the host-side mirror of the reference's ORT-facing API.
"""
__version__ = "0.1.0"
