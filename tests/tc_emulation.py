"""CPU emulation of the chunked tensor-core algorithm (csrc/wkv6_tc_fwd.cu / wkv6_tc_bwd.cu) --
TEST INFRASTRUCTURE ONLY.

Same algebra, same reference points (rho_q at block middles), bf16 rounding exactly where the
kernels round (MMA operands, checkpoints), fp32 accumulation.  Used on the CPU (no GPU needed) to
validate the chunked backward identities against the fp64 oracle and to size the numerical error
of a design choice before spending GPU time.  Run as a script for a small report.
"""
from __future__ import annotations

import math

import torch

L = 64
LOG2E = 1.4426950408889634


INT_RHO = False
AB_ROUNDED = False
NOROUND = set()   # experiment switch: operand names whose rounding is skipped


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


def chunk_quantities(w, dtype):
    """per chunk [S,64,64]: l (natural log decay), cum, exc, rho per row (own block middle), Lam"""
    l = -torch.exp(w.to(dtype))
    cum = torch.cumsum(l, dim=1)
    exc = cum - l
    # rho_q = exc at row 16q+8
    rho_blocks = exc[:, 8::16, :]                      # [S,4,64]
    if INT_RHO:                                        # references on the integer log2 grid
        rho_blocks = torch.round(rho_blocks * LOG2E) / LOG2E
    rho_row = rho_blocks.repeat_interleave(16, dim=1)  # [S,64,64]
    lam = cum[:, -1, :]
    return l, cum, exc, rho_blocks, rho_row, lam


def forward_backward(r, k, v, w, u, gy, s0=None, *, dtype=torch.float32, bf=True, stage_bf16=False,
                     gl_mode="ab"):
    """r,k,v,w,gy [B,T,C] bf16-valued tensors, u [H,64], s0 None or [B,H,64(key),64(value)].
    Returns dict(y, gr, gk, gv, gw, gu [H,64], gs [B,H,key,value]).  T is padded to a multiple of 64
    with zero rows (l = 0 there), exactly like the TMA zero fill + `valid` predicate."""
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

    ckpt = []
    Y = torch.zeros(B * H, NC, L, 64, dtype=dtype)
    per_chunk = []
    for c in range(NC):
        rc, kc, vc, wc = R[:, c], K[:, c], V[:, c], W[:, c]
        wc = torch.where(valid[c][None, :, None], wc, torch.full_like(wc, -1e30))   # l = 0 on padded rows
        l, cum, exc, rho_b, rho_row, lam = chunk_quantities(wc, dtype)
        E = torch.exp(exc - rho_row)            # Rt = r * E
        F = torch.exp(rho_row - cum)            # Kt(own block) = k * F
        Rt = rb(rc * E, bf, "Rt")
        Rh = rb(rc * E * torch.exp(rho_row), bf, "Rh")
        Kh = kc * F * torch.exp(lam[:, None, :] - rho_row)
        Kh_hi = rb(Kh, bf)
        Kh_lo = rb(Kh - Kh_hi, bf)
        # A[t,s] for s<t, computed per target block q with reference rho_q
        A = torch.zeros(B * H, L, L, dtype=dtype)
        Ktq = []
        for q in range(4):
            Kt_q = rb(kc * torch.exp(rho_b[:, q][:, None, :] - cum), bf, "Kt")          # all s rows (only s <= 16q+15 used)
            Ktq.append(Kt_q)
            A[:, 16 * q:16 * q + 16, :] = torch.einsum("sti,sui->stu", Rt[:, 16 * q:16 * q + 16], Kt_q)
        diagu = (rc * U[:, None, :] * kc).sum(-1)
        P = torch.where(tril, A, torch.zeros_like(A))
        P = rb(P + torch.diag_embed(diagu), bf)
        ckpt.append(rb(S, bf, "ckpt"))                                     # bf16 state at chunk start (key,value)
        Y[:, c] = torch.einsum("sti,sij->stj", Rh, rb(S, bf)) + torch.einsum("stu,suj->stj", P, vc)
        S = torch.exp(lam)[:, :, None] * S + torch.einsum("sti,stj->sij", Kh_hi, vc) + torch.einsum("sti,stj->sij", Kh_lo, vc)
        per_chunk.append((l, cum, exc, rho_b, rho_row, lam, E, F, Rt, Rh, Kh_hi, Ktq, P, diagu))
    y = _unstreams(Y.reshape(B * H, Tp, 64), B, H)[:, :T]

    # ------------------------------------------------------------------ backward (reverse sweep)
    G = torch.zeros(B * H, 64, 64, dtype=dtype)          # dL/dS at the chunk end [key,value]
    GR, GK, GV, GW = (torch.zeros(B * H, NC, L, 64, dtype=dtype) for _ in range(4))
    GU = torch.zeros(B * H, 64, dtype=dtype)
    triu = tril.transpose(0, 1)
    for c in range(NC - 1, -1, -1):
        rc, kc, vc, gyc = R[:, c], K[:, c], V[:, c], GY[:, c]
        l, cum, exc, rho_b, rho_row, lam, E, F, Rt, Rh, Kh_hi, Ktq, P, diagu = per_chunk[c]
        Sin = ckpt[c]
        Gb = rb(G, bf, "Gb")
        Bm = torch.einsum("stj,suj->stu", gyc, vc)                  # [t,s]
        bd = torch.diagonal(Bm, dim1=1, dim2=2)
        dA = rb(torch.where(tril, Bm, torch.zeros_like(Bm)), bf, "dA")    # [t,s], s<t
        # gv
        gv = torch.einsum("stu,stj->suj", P, gyc) + torch.einsum("sui,sij->suj", Kh_hi, Gb)
        # Xr[i,t] = sum_{s<t} Kt_q[s,i] dA[t,s] + e^{rho_q} sum_j Sin[i,j] gy_t[j]
        Drs = torch.einsum("sij,stj->sti", Sin, gyc)                # [t,i]
        Dks = torch.einsum("sij,stj->sti", Gb, vc)                  # [s,i]
        Dr = torch.zeros(B * H, L, 64, dtype=dtype)
        Dk = torch.zeros(B * H, L, 64, dtype=dtype)
        for q in range(4):
            hi = 16 * q + 16
            Dr[:, 16 * q:hi] = torch.einsum("stu,sui->sti", dA[:, 16 * q:hi, :hi], Ktq[q][:, :hi])
            # Rp_p[t,i] = r_t exp(exc_t - rho_p), t >= 16p
            Rp = rb(rc[:, 16 * q:] * torch.exp(exc[:, 16 * q:] - rho_b[:, q][:, None, :]), bf, "Rp")
            Dk[:, 16 * q:hi] = torch.einsum("stu,sti->sui", dA[:, 16 * q:, 16 * q:hi], Rp)
        Xr = Dr + torch.exp(rho_row) * Drs
        Xk = Dk + torch.exp(lam[:, None, :] - rho_row) * Dks
        if stage_bf16:
            Xr, Xk = rb(Xr), rb(Xk)
        grs = E * Xr
        gks = F * Xk
        GR[:, c] = grs + U[:, None, :] * kc * bd[:, :, None]
        GK[:, c] = gks + U[:, None, :] * rc * bd[:, :, None]
        GV[:, c] = gv
        GU += (rc * kc * bd[:, :, None]).sum(1)
        if AB_ROUNDED:
            # A / B from the SAME rounded factors the MMAs multiplied (own-block versions)
            Kt_own = torch.cat([Ktq[q][:, 16 * q:16 * q + 16] for q in range(4)], 1)
            Aterm = Rt * Dr + rc * E * torch.exp(rho_row) * Drs
            Bterm = Kt_own * Dk + kc * F * torch.exp(lam[:, None, :] - rho_row) * Dks
        else:
            Aterm = rc * grs
            Bterm = kc * gks
        # Q at the chunk end = <S_end, G>, S_end = checkpoint of the next chunk
        if c == NC - 1:
            qend = torch.zeros(B * H, 64, dtype=dtype)
        else:
            qend = (ckpt[c + 1] * G).sum(-1)
        if gl_mode == "ab":
            dd = Aterm - Bterm
            # Q_{t+1} = qend + sum_{s>t} dd_s
            suffix = torch.flip(torch.cumsum(torch.flip(dd, [1]), 1), [1]) - dd
            gl = qend[:, None, :] + suffix - Bterm
        elif gl_mode == "direct":
            # no cancellation between chunk-level quantities: inter terms as direct prefix / suffix sums
            Ai, Bi = Rt * Dr, Kt_own * Dk                                  # intra (bit-identical pair products)
            Ae = rc * E * torch.exp(rho_row) * Drs                        # inter: r e^{exc} (S_in gy)
            Be = kc * F * torch.exp(lam[:, None, :] - rho_row) * Dks      # inter: k e^{Lam-cum} (G v)
            q0 = torch.exp(lam) * (Sin * G).sum(-1)
            suf = lambda x: torch.flip(torch.cumsum(torch.flip(x, [1]), 1), [1]) - x
            pre = lambda x: torch.cumsum(x, 1) - x
            gl = q0[:, None, :] + pre(Be) + suf(Ae) + suf(Ai - Bi) - Bi
        else:
            raise ValueError(gl_mode)
        GW[:, c] = l * gl
        G = torch.exp(lam)[:, :, None] * G + torch.einsum("sti,stj->sij", Rh, gyc)
    un = lambda X: _unstreams(X.reshape(B * H, Tp, 64), B, H)[:, :T]
    gw = un(GW)
    if s0 is None and T > 0:
        gw[:, 0] = 0
    return dict(y=y, gr=un(GR), gk=un(GK), gv=un(GV), gw=gw, gu=GU.reshape(B, H, 64).sum(0),
                gs=G.reshape(B, H, 64, 64))


def report(B=1, T=130, H=2, decay="model", seed=0, **kw):
    from oracle import wkv6_oracle as O
    from rwkv_lm_ext_b200.synthetic import make_inputs
    from tests.util import BF16_MAXABS_ABS, BF16_MAXABS_REL, relrms
    r, k, v, w, u, gy = make_inputs(B, T, H, seed=seed, decay=decay)
    ref = O.wkv6_backward(r, k, v, w, u, gy)
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
    for decay in ("model", "randn"):
        for kw in (dict(bf=False), dict(bf=True), dict(bf=True, stage_bf16=True)):
            for (B, T, H) in ((1, 17, 1), (2, 64, 2), (1, 257, 1)):
                res = report(B, T, H, decay, seed=B * 1000 + T, **kw)
                print(decay, kw, (B, T, H), " ".join(f"{k}:{a:.2e}/{b:.2f}" for k, (a, b) in res.items()))
        sys.stdout.flush()
