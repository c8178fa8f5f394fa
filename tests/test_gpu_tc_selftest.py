"""tcgen05 / TMA / TMEM building blocks: one 64xNx64 MMA per operand layout against torch.matmul.
`pytest -m gpu`."""
import ctypes
import os
import subprocess

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libwkv6_b200_selftest.so")


def build_selftest(force=False):
    """tests/csrc/tc_selftest.cu -> tests/libwkv6_b200_selftest.so (git-ignored, travels to the GPU box).  Test code only:
    the product library does not contain it."""
    src = os.path.join(HERE, "csrc", "tc_selftest.cu")
    hdr = os.path.join(HERE, "..", "rwkv_lm_ext_b200", "csrc", "tc_common.cuh")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "-gencode", "arch=compute_100a,code=sm_100a", "-O3",
                               "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-o", SO, src, "-lcuda"])
    return SO


@pytest.fixture(scope="module")
def lib():
    return ctypes.CDLL(build_selftest())


def _check(lib, rc, what):
    fn = lib.wkv6b200_selftest_last_error
    fn.restype = ctypes.c_char_p
    assert rc == 0, f"{what}: {fn().decode()}"


@pytest.mark.parametrize("manual", [0, 4])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("N,boff", [(64, 0), (16, 0), (16, 8)])
def test_mma_layouts(lib, a_mn, b_mn, manual, N, boff):
    if boff and b_mn:
        pytest.skip("row offset is a K-major B case")
    fn = lib.wkv6b200_tc_selftest
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    g = torch.Generator().manual_seed(a_mn * 2 + b_mn + N)
    A = torch.randn(64, 64, generator=g).bfloat16().cuda()
    B = torch.randn(64, 64, generator=g).bfloat16().cuda()
    D = torch.full((64, 64), float("nan"), device="cuda")
    flags = a_mn | (b_mn << 1) | manual | boff
    _check(lib, fn(flags, N, A.data_ptr(), B.data_ptr(), D.data_ptr(), torch.cuda.current_stream().cuda_stream), "selftest")
    torch.cuda.synchronize()
    Am = (A.t() if a_mn else A).float()                  # [M, K]
    Bm = (B.t() if b_mn else B).float()                  # [N, K]
    if boff:
        Bm = Bm[16:]
    want = Am @ Bm[:N].t()
    got = D[:, :N]
    err = (got - want).abs().max().item()
    assert err < 1e-3 * max(1.0, want.abs().max().item()), f"flags={flags} N={N}: max err {err}\n{got[:4,:4]}\n{want[:4,:4]}"


@pytest.mark.parametrize("co", [0, 1, 2, 3])
def test_fragment_building_blocks(lib, co):
    """ldmatrix/stmatrix (+trans) on swizzled tiles, tcgen05.ld.16x256b fragment layout, TMA store,
    and an MN-major N=16 B operand at a column offset inside the swizzle row."""
    fn = lib.wkv6b200_tc_selftest2
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 6
    g = torch.Generator().manual_seed(100 + co)
    A = torch.randn(64, 64, generator=g).bfloat16().cuda()
    B = torch.randn(64, 64, generator=g).bfloat16().cuda()
    O = torch.zeros(64, 64, dtype=torch.bfloat16, device="cuda")
    F = torch.full((64, 64), float("nan"), device="cuda")
    D = torch.full((64 * 16 + 1,), float("nan"), device="cuda")
    D[-1] = 0.0
    _check(lib, fn(co, A.data_ptr(), B.data_ptr(), O.data_ptr(), F.data_ptr(), D.data_ptr(),
                   torch.cuda.current_stream().cuda_stream), "selftest2")
    torch.cuda.synchronize()
    assert torch.equal(F, A.float().t()), "ldmatrix.trans fragment layout"
    assert torch.equal(O, A.t()), "stmatrix + TMA store"
    assert D[-1].item() == 0.0, "an M=64 MMA disturbed TMEM lanes 16-31 (used as per-thread scratch)"
    D = D[:-1].view(64, 16)
    want = A.float() @ B.float()[:, 16 * co:16 * co + 16]
    err = (D - want).abs().max().item()
    assert err < 1e-3 * max(1.0, want.abs().max().item()), f"co={co}: max err {err}\n{D[:4,:4]}\n{want[:4,:4]}"
