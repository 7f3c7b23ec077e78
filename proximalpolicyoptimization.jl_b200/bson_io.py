"""The subset of BSON.jl's wire format the reference's disk rollouts use (writer side + a small parser).

BSON.jl lowers a bits-type array to ``{tag:"array", type:{tag:"datatype", params:[], name:[module.., T]}, size:[..],
data:<raw little-endian bytes, column-major>}`` and a struct to ``{tag:"struct", type:{..}, data:[fields..]}``.
The array encoding is pinned byte-for-byte against the reference's own ``output/states/sample_1.bson``
(tests/test_disk_replay.py); the struct encoding follows BSON.jl's ``lower`` for ``StateData``
(test/quad_game_utilities.jl:17-20) and is an ASSUMPTION where the reference holds no fixture.
"""
from __future__ import annotations

import struct

import numpy as np

_JL = {np.dtype(np.int64): ("Core", "Int64"), np.dtype(np.float32): ("Core", "Float32"),
       np.dtype(np.float64): ("Core", "Float64"), np.dtype(np.int32): ("Core", "Int32"),
       np.dtype(np.uint8): ("Core", "UInt8"), np.dtype(np.bool_): ("Core", "Bool")}


def _cstr(s):
    return s.encode() + b"\x00"


def _doc(items):
    body = b"".join(items)
    return struct.pack("<i", len(body) + 5) + body + b"\x00"


def _e_str(k, v):
    b = v.encode() + b"\x00"
    return b"\x02" + _cstr(k) + struct.pack("<i", len(b)) + b


def _e_doc(k, d, array=False):
    return (b"\x04" if array else b"\x03") + _cstr(k) + d


def _e_i64(k, v):
    return b"\x12" + _cstr(k) + struct.pack("<q", v)


def _e_bin(k, b):
    return b"\x05" + _cstr(k) + struct.pack("<i", len(b)) + b"\x00" + b


def _datatype(name):
    return _doc([_e_str("tag", "datatype"), _e_doc("params", _doc([]), array=True),
                 _e_doc("name", _doc([_e_str(str(i), n) for i, n in enumerate(name)]), array=True)])


def lower_array(a_colmajor_shape, arr):
    """``arr``: numpy array whose C-order bytes are the Julia column-major bytes; ``a_colmajor_shape``: Julia size."""
    arr = np.ascontiguousarray(arr)
    return _doc([_e_str("tag", "array"), _e_doc("type", _datatype(_JL[arr.dtype])),
                 _e_doc("size", _doc([_e_i64(str(i), int(s)) for i, s in enumerate(a_colmajor_shape)]), array=True),
                 _e_bin("data", arr.tobytes())])


def lower_state_data(vertex_score, action_mask, type_name=("Main", "StateData")):
    """StateData(vertex_score [nhe, nf] C-order == Julia [nf, nhe], action_mask [A])."""
    vs = np.ascontiguousarray(vertex_score)
    am = np.ascontiguousarray(action_mask)
    fields = _doc([_e_doc("0", lower_array((vs.shape[1], vs.shape[0]), vs)), _e_doc("1", lower_array((am.shape[0],), am))])
    return _doc([_e_str("tag", "struct"), _e_doc("type", _datatype(type_name)), _e_doc("data", fields, array=True)])


def save_state(path, state):
    """``BSON.@save path state`` (src/rollouts_to_disk.jl:47-51): top-level document {state: lowered}."""
    if hasattr(state, "vertex_score"):
        low = lower_state_data(state.vertex_score, state.action_mask)
    else:
        a = np.asarray(state)
        low = lower_array(a.shape[::-1] if a.ndim > 1 else a.shape, a)
    with open(path, "wb") as f:
        f.write(_doc([_e_doc("state", low)]))


# ---- a small pure-Python parser (host-side load_sample; the bulk path is the C++ loader) ------------------------
def _parse_doc(b, pos=0, as_list=False):
    size = struct.unpack_from("<i", b, pos)[0]
    end, p = pos + size, pos + 4
    out = {}
    while b[p] != 0:
        t = b[p]; p += 1
        e = b.index(b"\x00", p); key = b[p:e].decode(); p = e + 1
        if t in (3, 4):
            out[key], p = _parse_doc(b, p, as_list=(t == 4))
        elif t == 2:
            n = struct.unpack_from("<i", b, p)[0]; out[key] = b[p + 4:p + 4 + n - 1].decode(); p += 4 + n
        elif t == 5:
            n = struct.unpack_from("<i", b, p)[0]; out[key] = b[p + 5:p + 5 + n]; p += 5 + n
        elif t == 18:
            out[key] = struct.unpack_from("<q", b, p)[0]; p += 8
        elif t == 16:
            out[key] = struct.unpack_from("<i", b, p)[0]; p += 4
        elif t == 1:
            out[key] = struct.unpack_from("<d", b, p)[0]; p += 8
        elif t == 8:
            out[key] = bool(b[p]); p += 1
        elif t == 10:
            out[key] = None
        else:
            raise ValueError(f"unsupported BSON element type {t}")
    if as_list:
        out = [out[k] for k in sorted(out, key=int)]
    return out, end


_NP = {"Int64": np.int64, "Float32": np.float32, "Float64": np.float64, "Int32": np.int32, "UInt8": np.uint8, "Bool": np.bool_}


def raise_value(v):
    """Inverse of the lowering: arrays -> numpy (C-order = Julia's trailing dimension first), structs -> field list."""
    if isinstance(v, dict) and v.get("tag") == "array":
        dt = _NP[v["type"]["name"][-1]]
        return np.frombuffer(v["data"], dtype=dt).reshape(tuple(v["size"])[::-1]).copy()
    if isinstance(v, dict) and v.get("tag") == "struct":
        return [raise_value(f) for f in v["data"]]
    return v


def load_state(path):
    """``BSON.load(path)[:state]`` (src/dataset.jl:40)."""
    doc, _ = _parse_doc(open(path, "rb").read())
    return raise_value(doc["state"])
