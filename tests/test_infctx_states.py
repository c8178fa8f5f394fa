"""Host logic of the carried-state container (rwkv_lm_ext_b200/infctx.py) -- same behaviour as
src/infctx_module.py:3-50.  CPU only."""
import torch

from rwkv_lm_ext_b200.infctx import BlockState, BlockStateList, ChannelMixState, TimeMixState


def test_create_shapes_dtypes_and_zero():
    s = BlockStateList.create(3, 2, 128, 2, "cpu", torch.float32)
    assert s.wkv_states.shape == (3, 2, 2, 64, 64) and s.wkv_states.dtype == torch.bfloat16   # bf16 whatever dtype says
    assert s.shift_states.shape == (3, 2, 2, 128) and s.shift_states.dtype == torch.float32
    assert s.wkv_states.abs().sum() == 0 and s.shift_states.abs().sum() == 0 and len(s) == 3
    s32 = BlockStateList.create(1, 1, 64, 1, "cpu", torch.bfloat16, wkv_dtype=torch.float32)
    assert s32.wkv_states.dtype == torch.float32


def test_getitem_returns_views_and_setitem_copies():
    s = BlockStateList.create(2, 1, 64, 1, "cpu", torch.float32)
    blk = s[1]
    blk.time_mix_state.wkv_state.fill_(2.0)                     # a view: writes through (the in-place infctx op relies on it)
    assert s.wkv_states[1].float().mean() == 2.0 and s.wkv_states[0].abs().sum() == 0
    new = BlockState(TimeMixState(torch.ones(1, 64), torch.full((1, 1, 64, 64), 3.0)), ChannelMixState(torch.full((1, 64), 5.0)))
    s[0] = new
    assert s.shift_states[0, 0].mean() == 1 and s.shift_states[0, 1].mean() == 5 and s.wkv_states[0].float().mean() == 3
    new.time_mix_state.shift_state.zero_()                      # a copy: later changes do not leak in
    assert s.shift_states[0, 0].mean() == 1
    d = s.detach()
    assert not d.wkv_states.requires_grad and d.wkv_states.data_ptr() == s.wkv_states.data_ptr()
