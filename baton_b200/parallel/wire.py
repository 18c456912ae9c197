"""Message (de)serialisation for the control plane.

The reference ships ``pickle.dumps(dict)`` in HTTP bodies in both directions
(manager.py:77-85, worker.py:109-117) and ``pickle.loads`` whatever arrives
(manager.py:98, worker.py:92) -- remote code execution for any peer that can
reach the port (quirk 9).  The schemas are kept:

  round_start : {"state_dict": OrderedDict[str, Tensor], "update_name": str, "n_epoch": int}
  update      : {"state_dict": OrderedDict[str, Tensor], "n_samples": int,
                 "update_name": str, "loss_history": list[float]}

but decoding goes through an allow-listing unpickler that only rebuilds
tensors, containers and scalars, so a stock reference peer's payload still
loads while ``os.system`` gadgets do not.  ``torch.storage._load_from_bytes``
(what a pickled tensor's storage reduces to) is itself ``torch.load(...,
weights_only=False)``, i.e. a second, unrestricted pickle nested inside the
payload; it is therefore NOT resolved to the torch function but to
:func:`_safe_load_from_bytes`, which decodes the nested stream with
``weights_only=True``.  Messages that carry no tensors (the
fused NVLink data plane moves tensors GPU-to-GPU) are plain JSON.
"""
from __future__ import annotations

import io
import json
import pickle
from typing import Any

_ALLOWED = {
    ("collections", "OrderedDict"),
    ("builtins", "dict"), ("builtins", "list"), ("builtins", "tuple"), ("builtins", "set"),
    ("builtins", "frozenset"), ("builtins", "int"), ("builtins", "float"), ("builtins", "str"),
    ("builtins", "bool"), ("builtins", "bytes"), ("builtins", "complex"), ("builtins", "slice"),
    ("builtins", "bytearray"),
    ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_tensor"),
    ("torch._utils", "_rebuild_parameter"), ("torch._utils", "_rebuild_parameter_with_state"),
    ("torch._utils", "_rebuild_qtensor"), ("torch._tensor", "_rebuild_from_type_v2"),
    ("torch.storage", "_load_from_bytes"), ("torch", "Size"), ("torch", "device"),
    ("torch", "Tensor"), ("torch.nn.parameter", "Parameter"),
    ("torch.serialization", "_get_layout"),
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("numpy", "ndarray"), ("numpy", "dtype"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
}
_ALLOWED_PREFIX_ATTRS = {
    "torch": {"float32", "float64", "float16", "bfloat16", "int64", "int32", "int16", "int8", "uint8",
              "bool", "float8_e4m3fn", "float8_e5m2", "FloatStorage", "DoubleStorage", "HalfStorage",
              "BFloat16Storage", "LongStorage", "IntStorage", "ShortStorage", "CharStorage",
              "ByteStorage", "BoolStorage", "UntypedStorage", "strided"},
    "torch.storage": {"UntypedStorage", "TypedStorage"},
}


class UnsafePayload(pickle.UnpicklingError):
    pass


def _safe_load_from_bytes(b):
    """Replacement for ``torch.storage._load_from_bytes``: the nested storage stream is itself a pickle, so it
    goes through torch's restricted ``weights_only`` unpickler instead of the stock unrestricted one."""
    import torch
    try:
        return torch.load(io.BytesIO(b), weights_only=True)
    except pickle.UnpicklingError as exc:
        raise UnsafePayload("nested storage payload rejected: {}".format(exc)) from exc


class _TensorUnpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str):
        if (module, name) == ("torch.storage", "_load_from_bytes"):
            return _safe_load_from_bytes
        if (module, name) in _ALLOWED or name in _ALLOWED_PREFIX_ATTRS.get(module, ()):
            return super().find_class(module, name)
        raise UnsafePayload("refusing to unpickle {}.{}".format(module, name))


def dumps(obj: Any, *, prefer_json: bool = False) -> bytes:
    """Serialise a control-plane message.  ``prefer_json`` is used by tensor-free
    messages; everything else is reference-compatible pickle."""
    if prefer_json:
        return json.dumps(obj).encode("utf-8")
    return pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)


def loads(body: bytes, *, trusted: bool = False) -> Any:
    """Inverse of :func:`dumps`; sniffs JSON vs pickle."""
    if not body:
        raise ValueError("empty message body")
    head = body.lstrip()[:1]
    if head in (b"{", b"["):
        return json.loads(body.decode("utf-8"))
    if trusted:
        return pickle.loads(body)
    return _TensorUnpickler(io.BytesIO(body)).load()
