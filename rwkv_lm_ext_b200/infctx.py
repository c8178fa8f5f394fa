"""Carried states of infinite-context training: the same five names, constructor
arguments and indexing behaviour as src/infctx_module.py:3-50, so `src/model.py`'s infctx blocks
(`Block.forward(x, last_state: BlockState)`, :904-933) and `tmix_x060_forward(layer, x, last_state)`
take either implementation.

One extension: `BlockStateList.create/empty(..., wkv_dtype=torch.float32)` carries the WKV states in
fp32.  The reference stores them in bf16 whatever `dtype` says (src/infctx_module.py:36-38), i.e. it
rounds the state at every chunk boundary; `RUN_CUDA_RWKV6_STATE` accepts both.
"""
from dataclasses import dataclass

import torch


@dataclass
class TimeMixState:
    shift_state: torch.Tensor        # [B, C]            last token of the previous chunk (time-mix input)
    wkv_state: torch.Tensor          # [B, H, 64, 64]    WKV state, layout [value][key]


@dataclass
class ChannelMixState:
    shift_state: torch.Tensor        # [B, C]            last token of the previous chunk (channel-mix input)


@dataclass
class BlockState:
    time_mix_state: TimeMixState
    channel_mix_state: ChannelMixState


class BlockStateList:
    """All layers' states in two tensors: wkv_states [L, B, H, 64, 64], shift_states [L, 2, B, C]
    (index 0 = time mix, 1 = channel mix).  `states[layer]` returns views; `states[layer] = s` copies in."""

    def __init__(self, shift_states: torch.Tensor, wkv_states: torch.Tensor):
        self.wkv_states = wkv_states
        self.shift_states = shift_states

    @staticmethod
    def empty(N, B, C, H, device, dtype, wkv_dtype=torch.bfloat16):
        n = C // H
        return BlockStateList(torch.empty((N, 2, B, C), device=device, dtype=dtype),
                              torch.empty((N, B, H, n, n), device=device, dtype=wkv_dtype))

    @staticmethod
    def create(N, B, C, H, device, dtype, wkv_dtype=torch.bfloat16):
        out = BlockStateList.empty(N, B, C, H, device, dtype, wkv_dtype)
        out.wkv_states.zero_()
        out.shift_states.zero_()
        return out

    def __len__(self):
        return self.wkv_states.shape[0]

    def __getitem__(self, layer: int) -> BlockState:
        return BlockState(TimeMixState(self.shift_states[layer, 0], self.wkv_states[layer]),
                          ChannelMixState(self.shift_states[layer, 1]))

    def __setitem__(self, layer: int, state: BlockState):
        self.shift_states[layer, 0].copy_(state.time_mix_state.shift_state)
        self.wkv_states[layer].copy_(state.time_mix_state.wkv_state)
        self.shift_states[layer, 1].copy_(state.channel_mix_state.shift_state)

    def detach(self):
        """Truncated BPTT between chunks (the reference re-wraps `.detach()`ed tensors by hand)."""
        return BlockStateList(self.shift_states.detach(), self.wkv_states.detach())
