"""Reader for the ONNX files of the reference's `--onnx-folder` contract, without `onnx` / `onnxruntime` (neither is installed).

What the reference's runtime side does with these files (and what is restated here):

* `ASR_Metadata.onnx` -- a graph-less model whose `metadata_props` carry the run-time constants
  (`audio_pcm_scale, max_seq_len, sample_rate, special_token_ids, supported_languages`); the scripts read them through
  `InferenceSession.get_modelmeta().custom_metadata_map` (Whisper/Inference_Whisper_ONNX.py:270-289,
  written by Whisper/Shared_Merged.py:115-126).  -> `read_metadata`.
* `<Model>_SharedInitializers.onnx` + `.onnx.data` -- one ModelProto whose `graph.initializer` entries are external
  references `{location, offset, length}` into a single raw blob; byte-identical tensors alias one `(offset, length)`
  (Whisper/Shared_Merged.py:152-224).  The scripts mmap the blob and hand every tensor to
  `SessionOptions.add_initializer` (`attach_shared_initializers`, :1713-1743).  -> `read_shared_initializers`
  (same checks: external location present, `length == prod(dims) * itemsize`, string / undefined tensors skipped).

The parser is a plain protobuf wire-format walk over the handful of fields involved (field numbers from onnx.proto3:
ModelProto.graph = 7, .metadata_props = 14; GraphProto.node = 1, .initializer = 5; TensorProto.dims = 1, .data_type = 2,
.name = 8, .raw_data = 9, .external_data = 13, .data_location = 14; StringStringEntryProto.key = 1, .value = 2;
NodeProto.input = 1, .output = 2, .name = 3, .op_type = 4).

Scope note: the exported graphs name their MatMul weights `onnx::MatMul_<n>` (TorchScript export with
constant folding, Whisper/Export_Whisper.py:753-763), so binding blob tensors to engine tensors needs the donor graphs'
node structure.  `bind_by_name` covers the tensors that keep their module-path names and reports the rest; checkpoints are
ingested from the HF / FunASR folders (`ingest.py`), which is what the exporter itself starts from.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

# TensorProto.DataType -> numpy (onnx.helper.tensor_dtype_to_np_dtype)
_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 4: np.uint16, 5: np.int16, 6: np.int32, 7: np.int64, 9: np.bool_,
           10: np.float16, 11: np.float64, 12: np.uint32, 13: np.uint64}
_BFLOAT16 = 16
_UNSHAREABLE = (0, 8)        # UNDEFINED, STRING (Shared_Merged.py `_UNSHAREABLE_INIT_TYPES` / `_is_shareable_initializer`)


class OnnxFormatError(ValueError):
    pass


def _varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    result = shift = 0
    while True:
        if pos >= len(buf):
            raise OnnxFormatError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise OnnxFormatError("varint longer than 64 bits")


def _fields(buf: memoryview) -> Iterator[Tuple[int, int, object]]:
    """(field number, wire type, value) for every field of one message; length-delimited values are memoryviews."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v, pos = bytes(buf[pos:pos + 8]), pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            if pos + ln > n:
                raise OnnxFormatError("length-delimited field runs past the end of its message")
            v, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            v, pos = bytes(buf[pos:pos + 4]), pos + 4
        else:
            raise OnnxFormatError(f"unsupported wire type {wt}")
        yield fno, wt, v


def _sint64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


@dataclass
class Initializer:
    name: str = ""
    dims: Tuple[int, ...] = ()
    data_type: int = 0
    data_location: int = 0                       # 1 = EXTERNAL
    external: Dict[str, str] = field(default_factory=dict)
    raw: Optional[bytes] = None


@dataclass
class Node:
    op_type: str
    inputs: Tuple[str, ...]
    outputs: Tuple[str, ...]
    name: str = ""


@dataclass
class Model:
    metadata: Dict[str, str]
    initializers: List[Initializer]
    nodes: List[Node]
    ir_version: int = 0
    producer_name: str = ""


def _kv(buf: memoryview) -> Tuple[str, str]:
    k = v = ""
    for fno, wt, val in _fields(buf):
        if fno == 1 and wt == 2:
            k = bytes(val).decode("utf-8")
        elif fno == 2 and wt == 2:
            v = bytes(val).decode("utf-8")
    return k, v


def _tensor(buf: memoryview) -> Initializer:
    t = Initializer()
    dims: List[int] = []
    for fno, wt, val in _fields(buf):
        if fno == 1:                                   # dims: packed or one varint per entry
            if wt == 2:
                p = 0
                while p < len(val):
                    d, p = _varint(val, p)
                    dims.append(_sint64(d))
            else:
                dims.append(_sint64(val))
        elif fno == 2 and wt == 0:
            t.data_type = int(val)
        elif fno == 8 and wt == 2:
            t.name = bytes(val).decode("utf-8")
        elif fno == 9 and wt == 2:
            t.raw = bytes(val)
        elif fno == 13 and wt == 2:
            k, v = _kv(val)
            t.external[k] = v
        elif fno == 14 and wt == 0:
            t.data_location = int(val)
    t.dims = tuple(dims)
    return t


def _node(buf: memoryview) -> Node:
    ins: List[str] = []
    outs: List[str] = []
    name = op = ""
    for fno, wt, val in _fields(buf):
        if wt != 2:
            continue
        if fno == 1:
            ins.append(bytes(val).decode("utf-8"))
        elif fno == 2:
            outs.append(bytes(val).decode("utf-8"))
        elif fno == 3:
            name = bytes(val).decode("utf-8")
        elif fno == 4:
            op = bytes(val).decode("utf-8")
    return Node(op, tuple(ins), tuple(outs), name)


def parse_model(path) -> Model:
    """ModelProto -> metadata_props, graph.initializer (without touching external data), graph.node."""
    data = memoryview(Path(path).read_bytes())
    m = Model({}, [], [])
    try:
        for fno, wt, val in _fields(data):
            if fno == 1 and wt == 0:
                m.ir_version = int(val)
            elif fno == 2 and wt == 2:
                m.producer_name = bytes(val).decode("utf-8")
            elif fno == 14 and wt == 2:
                k, v = _kv(val)
                m.metadata[k] = v
            elif fno == 7 and wt == 2:
                for gno, gwt, gval in _fields(val):
                    if gno == 5 and gwt == 2:
                        m.initializers.append(_tensor(gval))
                    elif gno == 1 and gwt == 2:
                        m.nodes.append(_node(gval))
    except (IndexError, UnicodeDecodeError) as exc:
        raise OnnxFormatError(f"{path}: not an ONNX ModelProto ({exc})") from exc
    return m


def read_metadata(path) -> Dict[str, str]:
    """`InferenceSession(path).get_modelmeta().custom_metadata_map` of the scripts (Inference_Whisper_ONNX.py:270-277)."""
    return dict(parse_model(path).metadata)


def _np_dtype(data_type: int):
    if data_type == _BFLOAT16:
        return np.uint16                       # raw bf16 bits; `to_float32` widens them
    if data_type not in _DTYPES:
        raise OnnxFormatError(f"unsupported TensorProto data_type {data_type}")
    return _DTYPES[data_type]


def to_float32(array: np.ndarray, data_type: int) -> np.ndarray:
    """fp32 copy of a blob tensor (what `b200asr_set_tensor` takes)."""
    if data_type == _BFLOAT16:
        return (array.astype(np.uint32) << 16).view(np.float32)
    return np.asarray(array, dtype=np.float32)


def read_shared_initializers(path, *, with_types: bool = False):
    """`attach_shared_initializers` (Whisper/Shared_Merged.py:1713-1743) without ORT: every shareable initializer of the shared
    model as a read-only `np.memmap` over the single `.data` blob (aliased tensors map the same bytes).  Raises like the
    reference when an initializer is not external or its recorded length disagrees with its shape."""
    path = Path(path)
    model = parse_model(path)
    arrays: Dict[str, np.ndarray] = {}
    types: Dict[str, int] = {}
    for init in model.initializers:
        if init.data_type in _UNSHAREABLE:
            continue
        location = init.external.get("location")
        if not location:
            raise RuntimeError(f"Shared initializer {init.name!r} is not external.")
        offset = int(init.external.get("offset", "0"))
        length = int(init.external.get("length", "0"))
        dtype = np.dtype(_np_dtype(init.data_type))
        shape = tuple(int(d) for d in init.dims)
        expected = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        if length and length != expected:
            raise RuntimeError(f"Shared initializer {init.name!r} length mismatch: {length} != {expected}.")
        data_path = path.parent / location
        if expected == 0:
            arrays[init.name] = np.zeros(shape, dtype)
        else:
            if offset + expected > data_path.stat().st_size:
                raise RuntimeError(f"Shared initializer {init.name!r} runs past the end of {data_path.name}.")
            arrays[init.name] = np.memmap(data_path, dtype=dtype, mode="r", offset=offset, shape=shape)
        types[init.name] = init.data_type
    return (arrays, types) if with_types else arrays


def blob_summary(path) -> Dict[str, int]:
    """The counters `save_shared_initializers_from_tensors` records in metadata_props, recomputed from the references."""
    model = parse_model(path)
    spans = set()
    logical = 0
    for init in model.initializers:
        if init.data_location != 1:
            continue
        off, ln = int(init.external.get("offset", "0")), int(init.external.get("length", "0"))
        spans.add((off, ln))
        logical += ln
    return {"initializer_count": len(model.initializers), "unique_data_count": len(spans),
            "deduplicated_initializer_count": len(model.initializers) - len(spans), "logical_data_bytes": logical,
            "physical_data_bytes": sum(ln for _, ln in spans)}


def bind_by_name(arrays: Dict[str, np.ndarray], types: Dict[str, int], wanted: Dict[str, Tuple[int, ...]]):
    """Pick blob tensors by name.  `wanted` maps a blob initializer name to the shape the caller expects; returns
    ({name: fp32 array}, [names that are missing or have another shape]).  Anonymous `onnx::MatMul_<n>` weights cannot be
    requested this way (see the module docstring)."""
    got: Dict[str, np.ndarray] = {}
    unresolved: List[str] = []
    for name, shape in wanted.items():
        a = arrays.get(name)
        if a is None or tuple(a.shape) != tuple(shape):
            unresolved.append(name)
            continue
        got[name] = to_float32(a, types.get(name, 1))
    return got, unresolved
