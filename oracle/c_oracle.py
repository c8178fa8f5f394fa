"""ctypes binding of oracle/wkv6_oracle.c (the C port of the reference recurrence).
TEST INFRASTRUCTURE ONLY -- see the header of wkv6_oracle.c."""
import ctypes
import os
import subprocess

import torch

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "libwkv6_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_DIR, "wkv6_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _DIR, "libwkv6_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.wkv6_oracle_num_threads.restype = ctypes.c_int
    return _lib


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _f32(t):
    return None if t is None else t.detach().to(torch.float32).contiguous()


def num_threads() -> int:
    return int(lib().wkv6_oracle_num_threads())


def use_all_cores() -> int:
    """Use every host core the process may run on (torchrun exports OMP_NUM_THREADS=1)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().wkv6_oracle_set_threads(int(n))
    return num_threads()


def forward(r, k, v, w, u, s0=None, w_kind=0, want_state=False):
    """r,k,v,w [B,T,C]; u [H,64]; s0 None | [H,64,64] | [B,H,64,64] in [value,key] layout.
    Returns y fp32 [B,T,C] (and the final state [B,H,64,64] when want_state)."""
    B, T, C = r.shape
    H = u.shape[0]
    r, k, v, w, u, s0 = map(_f32, (r, k, v, w, u, s0))
    y = torch.empty(B, T, C, dtype=torch.float32)
    sT = torch.empty(B, H, 64, 64, dtype=torch.float32) if want_state else None
    lib().wkv6_oracle_forward(B, T, H, _p(r), _p(k), _p(v), _p(w), _p(u), _p(s0),
                              int(s0 is not None and s0.dim() == 4), _p(y), _p(sT), w_kind)
    return (y, sT) if want_state else y


def backward(r, k, v, w, u, gy, s0=None, w_kind=0, zero_gw0=None):
    """Returns dict gr,gk,gv,gw [B,T,C], gu [B,C] (per-batch partials), gs [B,H,64,64] or None."""
    B, T, C = r.shape
    H = u.shape[0]
    r, k, v, w, u, gy, s0 = map(_f32, (r, k, v, w, u, gy, s0))
    out = {n: torch.empty(B, T, C, dtype=torch.float32) for n in ("gr", "gk", "gv", "gw")}
    out["gu"] = torch.empty(B, C, dtype=torch.float32)
    out["gs"] = torch.empty(B, H, 64, 64, dtype=torch.float32) if s0 is not None else None
    if zero_gw0 is None:
        zero_gw0 = s0 is None
    lib().wkv6_oracle_backward(B, T, H, _p(r), _p(k), _p(v), _p(w), _p(u), _p(s0),
                               int(s0 is not None and s0.dim() == 4), _p(gy), _p(out["gr"]),
                               _p(out["gk"]), _p(out["gv"]), _p(out["gw"]), _p(out["gu"]),
                               _p(out["gs"]), w_kind, int(zero_gw0))
    return out
