"""The C-ABI library loads and exports every symbol include/ppo_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "ppo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ppo_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    import ppo_b200
    from ppo_b200 import _lib
    names = _declared()
    assert len(names) >= 40
    lib = ctypes.CDLL(ppo_b200.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in ppo_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_loud_failure_without_gpu_or_library():
    import ppo_b200
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ppo_b200.PPOError) as e:
        ppo_b200.Context(0)
    assert e.value.code == -2   # PPO_ERR_CUDA: no silent CPU path


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "proximalpolicyoptimization.jl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".jl")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src and "ppo_oracle" not in src.replace(
                    "oracle/ppo_oracle", ""), f


def test_sass_has_bulk_copy_kernels():
    # the gather's TMA bulk path compiled for sm_100a (UBLKCP in SASS)
    import shutil
    import subprocess
    import ppo_b200
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-sass", "-fun", "gather_rows_bulk_kernel", ppo_b200.LIB_PATH],
                         capture_output=True, text=True).stdout
    if "UBLKCP" not in out:
        out = subprocess.run(["cuobjdump", "-sass", ppo_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in out
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", ppo_b200.LIB_PATH], capture_output=True,
                                       text=True).stdout
