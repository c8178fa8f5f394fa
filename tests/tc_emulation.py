"""CPU emulation of the chunked tensor-core algorithm (csrc/wkv6_tc4_fwd.cu / wkv6_tc4_bwd.cu) --
TEST INFRASTRUCTURE ONLY.

Same algebra, same reference points, bf16 rounding exactly where the kernels round (MMA operands,
checkpoints, the parked E / F factors), fp32 accumulation.  Used on the CPU (no GPU needed) to
validate the chunked identities against the fp64 oracle and to size the numerical error of a
design choice before spending GPU time.  Run as a script for a small report.

The scheme (all decay quantities in log2 units, per channel i, chunk of 64 tokens):
  l_t   = max(-exp(w_t) log2(e), -LCLAMP)         built-in floor: 2^-13 per token is below bf16 resolution
  cum   = inclusive prefix sum of l,  exc = cum - l,  Lam = cum_63
  rho_b = rint(exc at token 16b + 8), b = 0..3     INTEGER reference of every 16-token block
  E_t = 2^(exc_t - rho_b(t)),  F_s = 2^(rho_b(s) - cum_s)          within 8 decay steps of 1: no overflow
  Rt = bf16(r E),  Kt_own = bf16(k F)
  Kt_q[s] = Kt_own[s] 2^(rho_q - rho_b(s))   (q >= b(s));   Rp_p[t] = Rt[t] 2^(rho_b(t) - rho_p)   (p <= b(t))
     exact power-of-two multiples (<= 1) of the own versions: every pair product r k 2^(exc_t - cum_s)
     has ONE bf16 value whichever reference it was built from.
"""
from __future__ import annotations

import torch

L = 64
LOG2E = 1.4426950408889634
LN2 = 0.6931471805599453
LCLAMP = 13.0

NOROUND = set()   # experiment switch: operand names whose rounding is skipped
REAL_RHO = False  # experiment switch: real-valued references (shows why the integer grid matters)
BF16_EF = False   # experiment switch: E and F reach the output stages as bf16 (costs the max-abs bound of gk at full shape)


def rb(x, on=True, name=None):
    """round to bf16 and come back (what an MMA operand / bf16 store sees)"""
    if name is not None and name in NOROUND:
        return x
    return x.bfloat16().to(x.dtype) if on else x


def _streams(x, H):
    B, T, C = x.shape
    return x.reshape(B, T, H, 64).permute(0, 2, 1, 3).reshape(B * H, T, 64)


def _unstreams(x, B, H):
    S, T, N = x.shape
    return x.reshape(B, H, T, N).permute(0, 2, 1, 3).reshape(B, T, H * N)


def _pow2(x):
    return torch.exp2(x)


def chunk_quantities(w, valid, dtype, lclamp):
    """w [S,64,64] raw logits of one chunk -> l, cum, exc (log2 units), rho_b [S,4,64], rho_row, Lam"""
    l = -torch.exp(w.to(dtype)) * LOG2E
    l = torch.clamp(l, min=-lclamp)
    l = torch.where(valid[None, :, None], l, torch.zeros_like(l))     # no decay on the padded rows
    cum = torch.cumsum(l, dim=1)
    exc = cum - l
    rho_b = exc[:, 8::16, :]
    if not REAL_RHO:
        rho_b = torch.round(rho_b)
    rho_row = rho_b.repeat_interleave(16, dim=1)
    return l, cum, exc, rho_b, rho_row, cum[:, -1, :]


def _versions_k(Kt_own, rho_b, bf):
    """Kt_q for q = 0..3 (rows of blocks <= q valid): cascade of exact power-of-two scalings"""
    S = Kt_own.shape[0]
    out = []
    cur = torch.zeros_like(Kt_own)
    for q in range(4):
        if q > 0:
            cur = rb((cur.double() * _pow2((rho_b[:, q] - rho_b[:, q - 1]).double())[:, None, :]).to(cur.dtype), bf)   # two exact factors in the kernel
        cur = cur.clone()
        cur[:, 16 * q:16 * q + 16] = Kt_own[:, 16 * q:16 * q + 16]
        out.append(cur)
    return out


def _versions_r(Rt, rho_b, bf):
    """Rp_p for p = 0..3 (rows of blocks >= p valid)"""
    out = [None] * 4
    cur = torch.zeros_like(Rt)
    for p in range(3, -1, -1):
        if p < 3:
            cur = rb((cur.double() * _pow2((rho_b[:, p + 1] - rho_b[:, p]).double())[:, None, :]).to(cur.dtype), bf)
        cur = cur.clone()
        cur[:, 16 * p:16 * p + 16] = Rt[:, 16 * p:16 * p + 16]
        out[p] = cur
    return out


def forward_backward(r, k, v, w, u, gy, s0=None, *, dtype=torch.float32, bf=True, lclamp=LCLAMP):
    """r,k,v,w,gy [B,T,C] bf16-valued tensors, u [H,64], s0 None or [B,H,64(key),64(value)].
    Returns dict(y, gr, gk, gv, gw, gu [H,64], gs [B,H,key,value], sT).  T is padded to a multiple of 64
    with zero rows (l = 0 there), exactly like the TMA zero fill."""
    B, T, C = r.shape
    H = u.shape[0]
    NC = (T + L - 1) // L
    Tp = NC * L
    pad = lambda x: torch.cat([x, torch.zeros(B, Tp - T, C, dtype=x.dtype)], 1) if Tp > T else x
    R, K, V, W, GY = (_streams(pad(x.to(dtype)), H).reshape(B * H, NC, L, 64) for x in (r, k, v, w, gy))
    valid = (torch.arange(Tp) < T).reshape(NC, L)
    U = u.to(dtype).reshape(1, H, 64).expand(B, H, 64).reshape(B * H, 64)
    S = torch.zeros(B * H, 64, 64, dtype=dtype) if s0 is None else s0.to(dtype).reshape(B * H, 64, 64).clone()
    tril = torch.tril(torch.ones(L, L, dtype=torch.bool), -1)        # [t,s], s < t
    blk = torch.arange(L) // 16

    ckpt = []
    Y = torch.zeros(B * H, NC, L, 64, dtype=dtype)
    per_chunk = []
    for c in range(NC):
        rc, kc, vc, wc = R[:, c], K[:, c], V[:, c], W[:, c]
        l, cum, exc, rho_b, rho_row, lam = chunk_quantities(wc, valid[c], dtype, lclamp)
        E = _pow2(exc - rho_row)
        F = _pow2(rho_row - cum)
        Rt = rb(rc * E, bf, "Rt")
        Kt_own = rb(kc * F, bf, "Kt")
        Rh = rb(Rt * _pow2(rho_row), bf)                       # exact unless it leaves the bf16 range
        Kh = kc * F * _pow2(lam[:, None, :] - rho_row)          # fp32
        Kh_hi = rb(Kh, bf)
        Kh_lo = rb(Kh - Kh_hi, bf)
        Ktq = _versions_k(Kt_own, rho_b, bf)
        A = torch.zeros(B * H, L, L, dtype=dtype)                # [t,s]
        for q in range(4):
            hi = 16 * q + 16
            A[:, 16 * q:hi, :hi] = torch.einsum("sti,sui->stu", Rt[:, 16 * q:hi], Ktq[q][:, :hi])
        diagu = (rc * U[:, None, :] * kc).sum(-1)
        P = rb(torch.where(tril, A, torch.zeros_like(A)) + torch.diag_embed(diagu), bf)
        Sb = rb(S, bf, "ckpt")
        ckpt.append(Sb)                                          # bf16 state at the chunk start (key,value)
        Y[:, c] = torch.einsum("sti,sij->stj", Rh, Sb) + torch.einsum("stu,suj->stj", P, vc)
        S = _pow2(lam)[:, :, None] * S + torch.einsum("sti,stj->sij", Kh_hi, vc) + torch.einsum("sti,stj->sij", Kh_lo, vc)
        per_chunk.append((l, cum, exc, rho_b, rho_row, lam, E, F, Rt, Kt_own, Ktq, P))
    y = _unstreams(Y.reshape(B * H, Tp, 64), B, H)[:, :T]
    sT = S.reshape(B, H, 64, 64)

    # ------------------------------------------------------------------ backward (reverse sweep)
    # G is kept in TMEM as G' = G_true 2^(-sigma) per key row, sigma = rho_0 of the chunk processed last
    Gs = torch.zeros(B * H, 64, 64, dtype=dtype)
    sigma = torch.zeros(B * H, 64, dtype=dtype)
    GR, GK, GV, GW = (torch.zeros(B * H, NC, L, 64, dtype=dtype) for _ in range(4))
    GU = torch.zeros(B * H, 64, dtype=dtype)
    for c in range(NC - 1, -1, -1):
        rc, kc, vc, gyc = R[:, c], K[:, c], V[:, c], GY[:, c]
        l, cum, exc, rho_b, rho_row, lam, E, F, Rt, Kt_own, Ktq, P = per_chunk[c]
        Rpp = _versions_r(Rt, rho_b, bf)
        Sin = ckpt[c]
        Epk, Fpk = (rb(E, bf), rb(F, bf)) if BF16_EF else (E, F)
        # ---- M1
        Bm = torch.einsum("stj,suj->stu", gyc, vc)                  # [t,s]
        bd = torch.diagonal(Bm, dim1=1, dim2=2)
        dA = rb(torch.where(tril, Bm, torch.zeros_like(Bm)), bf, "dA")
        Drs = torch.einsum("sij,stj->sti", Sin, gyc)                # [t,i]
        # ---- T1: <S_in, G_old>, Gb' = bf16(G_old 2^(Lam - rho_3)), G' <- G_old 2^(Lam - rho_0)
        q0 = _pow2(sigma) * (Sin * Gs).sum(-1)                      # <S_in, G_old>
        Gold = Gs * _pow2(sigma)[:, :, None]                        # two factors: their product may leave the fp32 range
        Gbp = rb(Gold * _pow2(lam - rho_b[:, 3])[:, :, None], bf, "Gb")
        Gs = Gold * _pow2(lam - rho_b[:, 0])[:, :, None]
        # ---- M2
        gv = torch.einsum("stu,stj->suj", P, gyc) + torch.einsum("sui,sij->suj", Ktq[3], Gbp)
        Dr = torch.zeros(B * H, L, 64, dtype=dtype)
        Dk = torch.zeros(B * H, L, 64, dtype=dtype)
        for q in range(4):
            hi = 16 * q + 16
            Dr[:, 16 * q:hi] = torch.einsum("stu,sui->sti", dA[:, 16 * q:hi, :hi], Ktq[q][:, :hi])
            Dk[:, 16 * q:hi] = torch.einsum("stu,sti->sui", dA[:, 16 * q:, 16 * q:hi], Rpp[q][:, 16 * q:])
        Gs = Gs + torch.einsum("sti,stj->sij", Rpp[0], gyc)         # G' += Rp_0^T GY
        sigma = rho_b[:, 0]
        # ---- T2
        Z = Dr + _pow2(rho_row) * Drs
        GR[:, c] = Epk * Z + U[:, None, :] * kc * bd[:, :, None]
        XA = Rt * Z                                                 # Ai + Ae
        # ---- M3 / T3
        Dks = torch.einsum("sij,stj->sti", Gbp, vc)                 # [s,i] = 2^(Lam - rho_3) (G_old v_s)
        Zk = Dk + _pow2(rho_b[:, 3][:, None, :] - rho_row) * Dks
        GK[:, c] = Fpk * Zk + U[:, None, :] * rc * bd[:, :, None]
        GV[:, c] = gv
        GU += (rc * kc * bd[:, :, None]).sum(1)
        X = XA - Kt_own * Dk                                        # XA - Bi
        Be = Kt_own * (_pow2(rho_b[:, 3][:, None, :] - rho_row) * Dks)
        D = Be - X
        preD = torch.cumsum(D, 1) - D
        gl = (_pow2(lam) * q0)[:, None, :] + X.sum(1, keepdim=True) + preD - XA
        GW[:, c] = l * LN2 * gl
    G = Gs * _pow2(sigma)[:, :, None]
    un = lambda X_: _unstreams(X_.reshape(B * H, Tp, 64), B, H)[:, :T]
    gw = un(GW)
    if T > 0:
        if s0 is None:
            gw[:, 0] = 0
        gw[:, T - 1] = 0
    return dict(y=y, gr=un(GR), gk=un(GK), gv=un(GV), gw=gw, gu=GU.reshape(B, H, 64).sum(0),
                gs=G.reshape(B, H, 64, 64), sT=sT)


def report(B=1, T=130, H=2, decay="model", seed=0, w_shift=0.0, w_scale=1.0, **kw):
    from oracle import wkv6_oracle as O
    from rwkv_lm_ext_b200.synthetic import make_inputs
    from tests.util import BF16_MAXABS_ABS, BF16_MAXABS_REL, relrms
    r, k, v, w, u, gy = make_inputs(B, T, H, seed=seed, decay=decay)
    if w_shift or w_scale != 1.0:
        w = (w.float() * w_scale + w_shift).bfloat16()
    ref = O.wkv6_backward(r, k, v, w, u, gy)          # the UNCLAMPED fp64 recurrence
    out = forward_backward(r, k, v, w, u, gy, **kw)
    res = {}
    for key in ("y", "gr", "gk", "gv", "gw", "gu"):
        a = rb(out[key].double()) if key != "gu" else out[key].double()
        b = ref[key].double()
        err = (a - b).abs().max().item()
        bound = BF16_MAXABS_REL * b.abs().max().item() + BF16_MAXABS_ABS
        res[key] = (relrms(a, b), err / bound)
    return res


if __name__ == "__main__":
    import sys
    for decay, kw in (("model", {}), ("randn", {}), ("randn", dict(w_shift=1.0, w_scale=1.5)), ("randn", dict(w_shift=2.5))):
        for bf in (False, True):
            for (B, T, H) in ((1, 17, 1), (2, 64, 2), (1, 257, 1)):
                res = report(B, T, H, decay, seed=B * 1000 + T, bf=bf, **kw)
                print(decay, kw, "bf" if bf else "fp32", (B, T, H), " ".join(f"{k}:{a:.2e}/{b:.2f}" for k, (a, b) in res.items()))
        sys.stdout.flush()
