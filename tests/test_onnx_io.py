"""onnx_io: the ONNX files of the reference's --onnx-folder contract read without onnx / onnxruntime.

The writer side of the test is independent of the parser: the messages are serialised by google.protobuf from a descriptor
built here with onnx.proto3's field numbers, laid out the way Whisper/Shared_Merged.py:152-224
(`save_shared_initializers_from_tensors`: sorted names, one raw blob, byte-identical tensors aliased to one (offset, length))
and :115-126 (`ASR_Metadata.onnx`: metadata_props only) write them."""
import hashlib
import json

import numpy as np
import pytest
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

from b200asr import onnx_io

T = descriptor_pb2.FieldDescriptorProto


def _messages():
    fd = descriptor_pb2.FileDescriptorProto(name="onnx_subset.proto", package="onnx_subset", syntax="proto3")

    def msg(name, fields):
        m = fd.message_type.add(name=name)
        for fname, num, typ, label, tname in fields:
            f = m.field.add(name=fname, number=num, type=typ, label=label)
            if tname:
                f.type_name = ".onnx_subset." + tname
    O, R = T.LABEL_OPTIONAL, T.LABEL_REPEATED
    msg("StringStringEntryProto", [("key", 1, T.TYPE_STRING, O, None), ("value", 2, T.TYPE_STRING, O, None)])
    msg("TensorProto", [("dims", 1, T.TYPE_INT64, R, None), ("data_type", 2, T.TYPE_INT32, O, None), ("name", 8, T.TYPE_STRING, O, None),
                        ("raw_data", 9, T.TYPE_BYTES, O, None), ("external_data", 13, T.TYPE_MESSAGE, R, "StringStringEntryProto"),
                        ("data_location", 14, T.TYPE_INT32, O, None)])
    msg("NodeProto", [("input", 1, T.TYPE_STRING, R, None), ("output", 2, T.TYPE_STRING, R, None), ("name", 3, T.TYPE_STRING, O, None),
                      ("op_type", 4, T.TYPE_STRING, O, None)])
    msg("GraphProto", [("node", 1, T.TYPE_MESSAGE, R, "NodeProto"), ("name", 2, T.TYPE_STRING, O, None),
                       ("initializer", 5, T.TYPE_MESSAGE, R, "TensorProto")])
    msg("OperatorSetIdProto", [("domain", 1, T.TYPE_STRING, O, None), ("version", 2, T.TYPE_INT64, O, None)])
    msg("ModelProto", [("ir_version", 1, T.TYPE_INT64, O, None), ("producer_name", 2, T.TYPE_STRING, O, None),
                       ("graph", 7, T.TYPE_MESSAGE, O, "GraphProto"), ("opset_import", 8, T.TYPE_MESSAGE, R, "OperatorSetIdProto"),
                       ("metadata_props", 14, T.TYPE_MESSAGE, R, "StringStringEntryProto")])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, "GetMessageClass", None)
    if get is None:                                            # older protobuf
        factory = message_factory.MessageFactory(pool)
        get = factory.GetPrototype
    return {n: get(pool.FindMessageTypeByName("onnx_subset." + n)) for n in ("ModelProto", "TensorProto")}


ONNX_TYPE = {np.dtype(np.float32): 1, np.dtype(np.int64): 7, np.dtype(np.float16): 10, np.dtype(np.int8): 3}


def _write_shared(tmp_path, tensors, bf16_names=()):
    """save_shared_initializers_from_tensors' layout."""
    M = _messages()
    path = tmp_path / "Whisper_SharedInitializers.onnx"
    data_name = path.name + ".data"
    model = M["ModelProto"](ir_version=10, producer_name="Whisper/Shared_Merged.py")
    model.opset_import.add(domain="", version=20)
    model.graph.name = "whisper_shared_initializers"
    offset, seen = 0, {}
    with open(tmp_path / data_name, "wb") as f:
        for name, arr in sorted(tensors.items()):
            raw = arr.tobytes()
            dt = 16 if name in bf16_names else ONNX_TYPE[arr.dtype]
            fp = (dt, arr.shape, len(raw), hashlib.sha256(raw).digest())
            if fp not in seen:
                seen[fp] = offset
                f.write(raw)
                offset += len(raw)
            ref = model.graph.initializer.add(name=name, data_type=dt, data_location=1)
            ref.dims.extend(arr.shape)
            for k, v in (("location", data_name), ("offset", str(seen[fp])), ("length", str(len(raw)))):
                ref.external_data.add(key=k, value=v)
    model.metadata_props.add(key="whisper_shared_initializers", value="1")
    model.metadata_props.add(key="initializer_count", value=str(len(tensors)))
    path.write_bytes(model.SerializeToString())
    return path


def test_shared_initializer_blob_roundtrip_with_aliasing(tmp_path):
    rng = np.random.default_rng(0)
    w = rng.standard_normal((48, 32)).astype(np.float32)
    tensors = {
        "encoder.layers.0.self_attn.qkv.bias": rng.standard_normal(96).astype(np.float32),
        "onnx::MatMul_4211": w,
        "onnx::MatMul_9000": w.copy(),                                   # byte-identical: aliased in the blob
        "decoder.embed_tokens.weight": rng.standard_normal((100, 32)).astype(np.float16),
        "position_ids": np.arange(448, dtype=np.int64),
        "quant.w": rng.integers(-128, 127, (16, 16)).astype(np.int8),
        "bf16.w": (rng.standard_normal((8, 4)).astype(np.float32).view(np.uint32) >> 16).astype(np.uint16),
        "scalar": np.asarray(3.5, np.float32),
    }
    path = _write_shared(tmp_path, tensors, bf16_names=("bf16.w",))
    arrays, types = onnx_io.read_shared_initializers(path, with_types=True)
    assert set(arrays) == set(tensors)
    for k, v in tensors.items():
        assert arrays[k].shape == v.shape and np.array_equal(np.asarray(arrays[k]), v), k
    assert isinstance(arrays["onnx::MatMul_4211"], np.memmap)
    assert arrays["onnx::MatMul_4211"].offset == arrays["onnx::MatMul_9000"].offset      # one physical copy
    s = onnx_io.blob_summary(path)
    assert s["initializer_count"] == 8 and s["unique_data_count"] == 7 and s["deduplicated_initializer_count"] == 1
    assert s["physical_data_bytes"] == (tmp_path / (path.name + ".data")).stat().st_size
    assert s["logical_data_bytes"] == sum(v.nbytes for v in tensors.values())
    bf = onnx_io.to_float32(arrays["bf16.w"], types["bf16.w"])
    assert bf.dtype == np.float32 and np.array_equal(bf.view(np.uint32) >> 16, tensors["bf16.w"])
    got, unresolved = onnx_io.bind_by_name(arrays, types, {"encoder.layers.0.self_attn.qkv.bias": (96,),
                                                           "decoder.embed_tokens.weight": (100, 32),
                                                           "encoder.layers.0.self_attn.qkv.weight": (96, 32),
                                                           "position_ids": (10,)})
    assert sorted(unresolved) == ["encoder.layers.0.self_attn.qkv.weight", "position_ids"]
    assert got["decoder.embed_tokens.weight"].dtype == np.float32
    assert onnx_io.parse_model(path).metadata["initializer_count"] == "8"


def test_reference_error_behaviour(tmp_path):
    M = _messages()
    # an initializer that is not external -> the reference raises RuntimeError (Shared_Merged.py:1727-1728)
    model = M["ModelProto"](ir_version=10)
    model.graph.initializer.add(name="inline", data_type=1, raw_data=np.zeros(4, np.float32).tobytes()).dims.extend([4])
    p = tmp_path / "a.onnx"
    p.write_bytes(model.SerializeToString())
    with pytest.raises(RuntimeError, match="not external"):
        onnx_io.read_shared_initializers(p)
    # recorded length != prod(dims) * itemsize -> RuntimeError (:1733-1736)
    model = M["ModelProto"](ir_version=10)
    ref = model.graph.initializer.add(name="w", data_type=1, data_location=1)
    ref.dims.extend([4, 4])
    for k, v in (("location", "b.onnx.data"), ("offset", "0"), ("length", "60")):
        ref.external_data.add(key=k, value=v)
    (tmp_path / "b.onnx.data").write_bytes(bytes(64))
    p = tmp_path / "b.onnx"
    p.write_bytes(model.SerializeToString())
    with pytest.raises(RuntimeError, match="length mismatch"):
        onnx_io.read_shared_initializers(p)
    # string tensors are skipped, as `_UNSHAREABLE_INIT_TYPES` are
    model = M["ModelProto"](ir_version=10)
    model.graph.initializer.add(name="s", data_type=8)
    p = tmp_path / "c.onnx"
    p.write_bytes(model.SerializeToString())
    assert onnx_io.read_shared_initializers(p) == {}
    with pytest.raises(onnx_io.OnnxFormatError):
        q = tmp_path / "junk.onnx"
        q.write_bytes(b"\x0a\xff\xff\xff\xff\x0fnot a protobuf")
        onnx_io.parse_model(q)


def test_asr_metadata_model_drives_the_pipeline_constants(tmp_path):
    """ASR_Metadata.onnx (metadata_props only) -> the custom_metadata_map the scripts read (:270-289)."""
    from b200asr.cli import whisper_metadata
    from b200asr.config import WHISPER_TINY_TEST
    gen = {"lang_to_id": {"<|en|>": 20, "<|zh|>": 21}, "task_to_id": {"transcribe": 11, "translate": 12}, "no_timestamps_token_id": 14,
           "decoder_start_token_id": 3, "eos_token_id": 2, "no_speech_token_id": 13}
    md = whisper_metadata(WHISPER_TINY_TEST, gen)
    M = _messages()
    model = M["ModelProto"](ir_version=10, producer_name="Whisper/Shared_Merged.py")
    model.opset_import.add(domain="", version=20)
    model.graph.name = "metadata"
    for k, v in md.items():
        model.metadata_props.add(key=str(k), value=str(v))
    p = tmp_path / "ASR_Metadata.onnx"
    p.write_bytes(model.SerializeToString())
    back = onnx_io.read_metadata(p)
    assert back == md
    assert json.loads(back["special_token_ids"])["no_speech"] == 13 and int(back["max_seq_len"]) == WHISPER_TINY_TEST.max_target


def test_graph_nodes_are_listed(tmp_path):
    M = _messages()
    model = M["ModelProto"](ir_version=10)
    n = model.graph.node.add(name="/layers.0/qkv/MatMul", op_type="MatMul")
    n.input.extend(["x", "onnx::MatMul_1"]); n.output.extend(["y"])
    n = model.graph.node.add(name="/layers.0/qkv/Add", op_type="Add")
    n.input.extend(["encoder.layers.0.qkv.bias", "y"]); n.output.extend(["z"])
    p = tmp_path / "g.onnx"
    p.write_bytes(model.SerializeToString())
    nodes = onnx_io.parse_model(p).nodes
    assert [(x.op_type, x.inputs, x.outputs) for x in nodes] == [("MatMul", ("x", "onnx::MatMul_1"), ("y",)),
                                                                 ("Add", ("encoder.layers.0.qkv.bias", "y"), ("z",))]
