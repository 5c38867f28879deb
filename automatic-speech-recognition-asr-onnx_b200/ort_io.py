"""Host helpers that build contiguous numpy inputs from model I/O metadata and parse the ASR metadata map.

Own restatement of the reference's helper module (/root/reference/ORT_IO.py: numpy_dtype :27, resolve_shape :37,
array_for :61, filled_for :95, scalar_for :109, metadata_by_name :117, load_special_token_ids :144,
load_supported_languages :149, resolve_supported_language :163) so the drop-in scripts keep calling the same
names with the same argument meaning and the same exceptions.  A "value_meta" is anything with ``name``,
``shape`` (ints, or strings/None for dynamic axes) and ``type`` ("tensor(float)", ...), i.e. an ORT ``NodeArg``
or the ``NodeArg`` of b200asr.session.
"""
from __future__ import annotations

import json
from typing import Any, Mapping, Sequence

import numpy as np

_TYPE_TABLE = {
    "bool": np.bool_, "double": np.float64, "float": np.float32, "float16": np.float16,
    "int8": np.int8, "int16": np.int16, "int32": np.int32, "int64": np.int64,
    "uint8": np.uint8, "uint16": np.uint16, "uint32": np.uint32, "uint64": np.uint64,
}


def numpy_dtype(value_or_type: Any) -> np.dtype:
    """dtype declared by a NodeArg or by a "tensor(<elem>)" string; KeyError for unknown types."""
    name = value_or_type if isinstance(value_or_type, str) else value_or_type.type
    if not (name.startswith("tensor(") and name.endswith(")")):
        raise KeyError(name)
    return np.dtype(_TYPE_TABLE[name[len("tensor("):-1]])


def is_dynamic_dim(dim: Any) -> bool:
    return not isinstance(dim, (int, np.integer))


def resolve_shape(value_meta: Any, *, symbols: Mapping[str, int] | None = None,
                  axes: Mapping[int, int] | None = None) -> tuple[int, ...]:
    """Static dims win; then an explicit per-axis override; then a symbol table; else int(dim) (raises)."""
    symbols = symbols or {}
    axes = axes or {}
    out = []
    for axis, dim in enumerate(value_meta.shape):
        if not is_dynamic_dim(dim):
            out.append(int(dim))
        elif axes.get(axis) is not None:
            out.append(int(axes[axis]))
        elif isinstance(dim, str) and dim in symbols:
            out.append(int(symbols[dim]))
        else:
            out.append(int(dim))
    return tuple(out)


def array_for(value_meta: Any, value: Any, *, symbols: Mapping[str, int] | None = None,
              axes: Mapping[int, int] | None = None) -> np.ndarray:
    """Contiguous array of the declared dtype, reshaped to the declared shape; dynamic axes take the value's own
    extent unless overridden; a dynamic axis beyond the value's rank with no override is a ValueError."""
    arr = np.asarray(value, dtype=numpy_dtype(value_meta))
    given = dict(axes or {})
    runtime = {}
    missing = []
    for axis, dim in enumerate(value_meta.shape):
        if not is_dynamic_dim(dim):
            continue
        if axis < arr.ndim:
            runtime[axis] = int(arr.shape[axis])
        elif axis not in given:
            missing.append(axis)
    if missing:
        raise ValueError(f"Value for {value_meta.name!r} has rank {arr.ndim}; provide axes for dynamic "
                         f"dimensions {missing!r}.")
    runtime.update(given)
    return np.ascontiguousarray(arr.reshape(resolve_shape(value_meta, symbols=symbols, axes=runtime)))


def filled_for(value_meta: Any, fill_value: Any = 0, *, symbols: Mapping[str, int] | None = None,
               axes: Mapping[int, int] | None = None) -> np.ndarray:
    return np.full(resolve_shape(value_meta, symbols=symbols, axes=axes), fill_value, dtype=numpy_dtype(value_meta))


def scalar_for(value_meta: Any, value: Any) -> np.ndarray:
    """A rank-0 array for a ``[]`` input, a one-element vector otherwise."""
    dt = numpy_dtype(value_meta)
    if tuple(value_meta.shape) == ():
        return np.asarray(value, dtype=dt).reshape(())
    return np.asarray([value], dtype=dt)


def metadata_by_name(values: Sequence[Any]) -> dict[str, Any]:
    return {v.name: v for v in values}


def metadata_int(metadata: Mapping[str, str], key: str, *, minimum: int | None = None) -> int:
    return int(metadata[key])


def metadata_int_list(metadata: Mapping[str, str], key: str) -> list[int]:
    return [int(tok) for tok in metadata[key].split(",") if tok]


def metadata_json_object(metadata: Mapping[str, str], key: str) -> dict[str, Any]:
    return json.loads(metadata[key])


def load_special_token_ids(metadata: Mapping[str, str]) -> dict[str, Any]:
    return metadata_json_object(metadata, "special_token_ids")


def load_supported_languages(metadata: Mapping[str, str]) -> dict[str, dict[str, Any]]:
    """The ``supported_languages`` catalog: code -> {name, aliases, prompt_token_ids, ...}, whitespace-trimmed."""
    catalog: dict[str, dict[str, Any]] = {}
    for code, raw in metadata_json_object(metadata, "supported_languages").items():
        key = code.strip()
        entry = dict(raw)
        entry["name"] = entry.get("name", key).strip()
        entry["aliases"] = [a.strip() for a in entry.get("aliases", [])]
        entry["prompt_token_ids"] = entry.get("prompt_token_ids", [])
        catalog[key] = entry
    return catalog


def resolve_supported_language(catalog: Mapping[str, Mapping[str, Any]], language: str):
    """Canonical code first (case-insensitive), then a unique alias; ValueError otherwise."""
    want = language.strip().casefold()
    for code, entry in catalog.items():
        if code.casefold() == want:
            return code, entry
    hits = [(code, entry) for code, entry in catalog.items()
            if any(str(a).casefold() == want for a in entry.get("aliases", ()))]
    if len(hits) == 1:
        return hits[0]
    raise ValueError(f"Unsupported language {language!r}; choose one of {sorted(catalog)}.")
