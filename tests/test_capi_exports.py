"""CPU-only: the C-ABI library builds, loads, and exports exactly what include/wkv6_b200.h declares.
No compute calls (no GPU here)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "wkv6_b200.h")).read()
    return set(re.findall(r"WKV6_API\s+[\w \*]+?\b(\w+)\s*\(", src))


def test_header_and_library_agree():
    from rwkv_lm_ext_b200 import _lib
    lib = _lib.load()
    declared = _declared()
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert declared <= exported, declared - exported
    # ... and nothing else: no test kernels, no profiling hooks in the product library
    assert exported <= declared, exported - declared
    assert lib.wkv6b200_abi_version() == 4


def test_argument_validation_without_gpu():
    from rwkv_lm_ext_b200 import _lib
    lib = _lib.load()
    # C != H*64 is rejected before anything touches the device
    rc = lib.wkv6_forward(1, 4, 100, 2, None, None, None, None, None, None, None)
    assert rc == -1 and b"C == H*64" in lib.wkv6b200_last_error()
    rc = lib.wkv6_forward(1, 4, 128, 2, None, None, None, None, None, None, None)
    assert rc == -1 and b"null pointer" in lib.wkv6b200_last_error()
    # empty problems are a successful no-op (reference: grid of 0 blocks)
    assert lib.wkv6_forward(0, 4, 128, 2, None, None, None, None, None, None, None) == 0
    assert lib.wkv6_backward_workspace_bytes(2, 8, 128, 2) >= 2 * 8 * 128 * 4


def test_cpu_tensors_fail_loudly():
    import torch
    import rwkv_lm_ext_b200 as M
    x = torch.zeros(1, 2, 64, dtype=torch.bfloat16)
    u = torch.zeros(1, 64, dtype=torch.bfloat16)
    with pytest.raises(M.Wkv6B200Error):
        M.RUN_CUDA_RWKV6(1, 2, 64, 1, x, x, x, x, u)
    with pytest.raises(AssertionError):          # reference-style dtype assert (src/model.py:195)
        M.RUN_CUDA_RWKV6(1, 2, 64, 1, x.float(), x, x, x, u)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "rwkv_lm_ext_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle/", "").lower() or f == "_lib.py", f
