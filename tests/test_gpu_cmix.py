"""Channel-mix kernels against the eager bf16 chain of RWKV_CMix_x060.forward (src/model.py:635-644) --
bit-identical forward -- and against fp64 autograd for the gradients.  `pytest -m gpu`."""
import pytest
import torch

from tests.util import assert_bf16_close, relrms

pytestmark = pytest.mark.gpu
DEV = "cuda"


class CMixRef(torch.nn.Module):
    def __init__(self, D, F_):
        super().__init__()
        self.time_maa_k = torch.nn.Parameter(torch.rand(1, 1, D))
        self.time_maa_r = torch.nn.Parameter(torch.rand(1, 1, D))
        self.key = torch.nn.Linear(D, F_, bias=False)
        self.receptance = torch.nn.Linear(D, D, bias=False)
        self.value = torch.nn.Linear(F_, D, bias=False)

    def forward(self, x, shift=None):
        prev = torch.zeros_like(x[:, :1]) if shift is None else shift.unsqueeze(1)
        xx = torch.cat([prev, x[:, :-1]], 1) - x
        xk = x + xx * self.time_maa_k
        xr = x + xx * self.time_maa_r
        k = torch.relu(self.key(xk)) ** 2
        return torch.sigmoid(self.receptance(xr)) * self.value(k)


@pytest.mark.parametrize("B,T,D,with_state", [(2, 37, 128, False), (3, 130, 320, True), (1, 1, 64, True)])
def test_cmix_forward_bit_exact_and_gradients(B, T, D, with_state):
    import rwkv_lm_ext_b200 as M
    torch.manual_seed(5)
    ref = CMixRef(D, 3 * D + 64)
    x = torch.randn(B, T, D)
    st = torch.randn(B, D) if with_state else None
    gout = torch.randn(B, T, D)
    layer = CMixRef(D, 3 * D + 64)
    layer.load_state_dict(ref.state_dict())
    layer = layer.bfloat16().to(DEV)
    xb = x.bfloat16().to(DEV).requires_grad_(True)
    stb = st.bfloat16().to(DEV).requires_grad_(True) if with_state else None
    if with_state:
        out, new = M.cmix_x060_forward(layer, xb, stb)
        assert torch.equal(new, xb[:, -1])
    else:
        out = M.cmix_x060_forward(layer, xb)
    with torch.no_grad():
        eager = layer(xb.detach(), None if stb is None else stb.detach())
    assert torch.equal(out, eager)                      # same op-by-op bf16 rounding as the eager chain
    out.backward(gout.bfloat16().to(DEV))
    # fp64 reference on the bf16-rounded parameters
    ref64 = CMixRef(D, 3 * D + 64).double()
    ref64.load_state_dict({k: v.detach().cpu().double() for k, v in layer.state_dict().items()})
    x64 = xb.detach().cpu().double().requires_grad_(True)
    st64 = stb.detach().cpu().double().requires_grad_(True) if with_state else None
    (ref64(x64, st64) * gout.bfloat16().double()).sum().backward()
    assert relrms(xb.grad, x64.grad) < 2e-2
    if with_state:
        assert relrms(stb.grad, st64.grad) < 2e-2
    for (n, p), (_, q) in zip(layer.named_parameters(), ref64.named_parameters()):
        assert relrms(p.grad, q.grad) < 3e-2, n


def test_relu_sq_and_sigmoid_mul_pieces():
    import rwkv_lm_ext_b200 as M
    g = torch.Generator().manual_seed(9)
    a = (torch.randn(3, 50, 448, generator=g) * 2).bfloat16().to(DEV).requires_grad_(True)
    b = torch.randn(3, 50, 448, generator=g).bfloat16().to(DEV).requires_grad_(True)
    go = torch.randn(3, 50, 448, generator=g).bfloat16().to(DEV)
    y = M.relu_sq(a)
    assert torch.equal(y, torch.relu(a.detach()) ** 2)
    y.backward(go)
    a64 = a.detach().double().requires_grad_(True)
    (torch.relu(a64) ** 2 * go.double()).sum().backward()
    assert_bf16_close(a.grad, a64.grad, "relu_sq grad")
    a.grad = None
    z = M.sigmoid_mul(a, b)
    assert torch.equal(z, torch.sigmoid(a.detach()) * b.detach())
    z.backward(go)
    a64 = a.detach().double().requires_grad_(True)
    b64 = b.detach().double().requires_grad_(True)
    (torch.sigmoid(a64) * b64 * go.double()).sum().backward()
    assert_bf16_close(a.grad, a64.grad, "sigmoid_mul gr")
    assert_bf16_close(b.grad, b64.grad, "sigmoid_mul gkv")
