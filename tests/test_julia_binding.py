"""The Julia binding (proximalpolicyoptimization.jl_b200/julia/PPOB200.jl) cannot be executed here (no Julia in the
image), so it is guarded statically: every `ccall((:sym, lib), Ret, (ArgTypes...), args...)` in it is parsed and checked
against the prototype of `sym` in include/ppo_b200.h — the symbol exists, the return type, the NUMBER of arguments and
every argument's C type agree, and the call passes as many values as it declares types.  The same check runs over
baseline/julia_ref.jl's ccalls (none expected) and over the stub shown in INTEGRATION.md."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ppo_b200.h")
JULIA = os.path.join(ROOT, "proximalpolicyoptimization.jl_b200", "julia", "PPOB200.jl")

# Julia ccall type -> canonical C class
JL = {
    "Cint": "i32", "Int64": "i64", "UInt64": "u64", "Cdouble": "f64", "Cfloat": "f32", "Cstring": "cstr",
    "Ptr{Cvoid}": "ptr", "Ref{Ptr{Cvoid}}": "pptr", "Ptr{Float32}": "p_f32", "Ptr{Int64}": "p_i64", "Ptr{UInt8}": "p_u8",
    "Ptr{Int8}": "p_i8", "Ptr{Int16}": "p_i16", "Ptr{UInt64}": "p_u64", "Ptr{Cint}": "p_i32", "Ref{Cint}": "p_i32", "Ref{Int64}": "p_i64",
    "Ref{Cdouble}": "p_f64", "Ptr{Cdouble}": "p_f64", "Ptr{Ptr{Float32}}": "pp_f32",
}


def c_class(t):
    """canonical class of a C parameter / return type"""
    t = re.sub(r"\bconst\b", "", t).strip()
    t = re.sub(r"\s+", " ", t).replace(" *", "*")
    if t in ("int", "ppo_status"):
        return "i32"
    if t in ("int64_t", "uint64_t", "double", "float"):
        return {"int64_t": "i64", "uint64_t": "u64", "double": "f64", "float": "f32"}[t]
    if t == "char*":
        return "cstr"
    if t == "void*" or re.fullmatch(r"ppo_\w+\*", t):
        return "ptr"
    if re.fullmatch(r"ppo_\w+\*\*", t):
        return "pptr"
    if t in ("float**", "float* *"):
        return "pp_f32"
    m = re.fullmatch(r"(float|double|int|int8_t|int16_t|int64_t|uint8_t|uint64_t)\*", t)
    if m:
        return "p_" + {"float": "f32", "double": "f64", "int": "i32", "int8_t": "i8", "int16_t": "i16", "int64_t": "i64",
                       "uint8_t": "u8", "uint64_t": "u64"}[m.group(1)]
    raise ValueError(f"unclassified C type {t!r}")


def compatible(jl, c):
    if jl == c:
        return True
    # an opaque byte buffer (IPC handle, NCCL id) may be passed as Ptr{UInt8}; a C string buffer as Ptr{UInt8}
    return (jl, c) in {("p_u8", "ptr"), ("p_u8", "cstr"), ("ptr", "cstr")}


def header_prototypes():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"^\s*([A-Za-z_][\w \*]*?)\s*\b(ppo_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.M | re.S):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith("typedef"):
            continue
        plist = []
        if params and params != "void":
            for prm in params.split(","):
                prm = prm.strip()
                prm = re.sub(r"\s*\[\s*\]$", "*", prm)
                mm = re.match(r"^(.*?)(\b[A-Za-z_]\w*)$", prm)          # strip the parameter name
                typ = mm.group(1).strip() if mm and mm.group(1).strip() else prm
                plist.append(c_class(typ))
        protos[name] = (c_class(ret), plist)
    return protos


def split_top(s):
    """split on commas at nesting depth 0"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def julia_ccalls(path):
    src = open(path).read()
    src = re.sub(r"#[^\n]*", "", src)
    calls = []
    for m in re.finditer(r"ccall\(\(\s*:(\w+)\s*,\s*lib\s*\)\s*,", src):
        # take the balanced argument list of this ccall
        i = src.index("(", m.start())
        depth, j = 0, i
        while True:
            depth += src[j] == "("
            depth -= src[j] == ")"
            if depth == 0:
                break
            j += 1
        parts = split_top(src[i + 1:j])
        ret, types = parts[1], parts[2]
        assert types.startswith("(") and types.endswith(")"), (m.group(1), types)
        tlist = split_top(types[1:-1])
        calls.append((m.group(1), ret, tlist, parts[3:]))
    return calls


def test_header_parses_every_exported_function():
    protos = header_prototypes()
    from ppo_b200 import _lib
    assert set(_lib.SIGNATURES) <= set(protos), sorted(set(_lib.SIGNATURES) - set(protos))
    assert len(protos) >= 55


@pytest.mark.parametrize("path", [JULIA])
def test_every_ccall_matches_the_header(path):
    protos = header_prototypes()
    calls = julia_ccalls(path)
    assert len(calls) >= 25
    for sym, ret, types, values in calls:
        assert sym in protos, f"ccall binds :{sym}, which include/ppo_b200.h does not declare"
        c_ret, c_args = protos[sym]
        assert ret in JL, (sym, ret)
        assert JL[ret] == c_ret, f":{sym} returns {c_ret} in the header, {ret} in the ccall"
        assert len(types) == len(c_args), f":{sym} takes {len(c_args)} arguments in the header, the ccall declares {len(types)}"
        assert len(values) == len(types), f":{sym}: {len(types)} argument types but {len(values)} values"
        for k, (jt, ct) in enumerate(zip(types, c_args)):
            assert jt in JL, (sym, k, jt)
            assert compatible(JL[jt], ct), f":{sym} argument {k + 1}: header {ct}, ccall {jt}"


def test_binding_covers_the_reference_api():
    """methods of the package's own generics a drop-in needs, specialised on the device containers"""
    src = open(JULIA).read()
    for method in ["PPO.update!(b::DeviceRollouts", "Base.length(b::DeviceRollouts", "PPO.compute_state_value!(b::DeviceRollouts",
                   "PPO.collect_rollouts!(b::DeviceRollouts", "PPO.permute!(b::DeviceRollouts", "PPO.shuffle!(b::DeviceRollouts",
                   "PPO.construct_dataset(b::DeviceRollouts", "Base.getindex(d::DeviceDataset", "PPO.step_batch!(p::DevicePolicy",
                   "PPO.step_epoch!(p::DevicePolicy", "PPO.ppo_train!(p::DevicePolicy", "PPO.action_probabilities(p::DevicePolicy",
                   "PPO.batch_action_probabilities(p::DevicePolicy"]:
        assert method in src, method
    assert src.count("function PPO.ppo_iterate!(policy::DevicePolicy") == 2      # buffer and disk variants


def test_python_and_julia_bind_the_same_hot_path_symbols():
    from ppo_b200 import _lib
    jl = {c[0] for c in julia_ccalls(JULIA)}
    hot = {"ppo_buffer_create", "ppo_buffer_append", "ppo_buffer_append_i8", "ppo_compute_returns", "ppo_permutation_generate",
           "ppo_permutation_set", "ppo_gather_indices", "ppo_policy_create", "ppo_policy_read", "ppo_adam_create",
           "ppo_step_batch_host", "ppo_step_epoch", "ppo_batch_action_probabilities", "ppo_disk_dataset_load"}
    assert hot <= jl and hot <= set(_lib.SIGNATURES)
