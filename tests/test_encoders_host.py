"""Host logic of the batched encoders (no GPU): length bucketing and cross-encoder row layout."""
import torch

from rwkv_lm_ext_b200.encoders import cross_encoder_rows, length_buckets


def test_length_buckets_cover_everything_within_the_token_budget():
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(1, 700, (200,), generator=g).tolist()
    batches = length_buckets(lens, max_tokens=4096)
    seen = sorted(i for members, _ in batches for i in members)
    assert seen == list(range(200))
    for members, width in batches:
        assert width % 64 == 0 and width >= max(lens[i] for i in members)
        assert width * len(members) <= 4096 or len(members) == 1
    pad = sum(width * len(m) for m, width in batches) / sum(lens)
    assert pad < 1.25                                           # sorted buckets waste little on padding


def test_cross_encoder_rows():
    rows = cross_encoder_rows([[5, 6]], [[8, 9, 10]], max_len=10, sep_id=2, class_id=1, pad_id=0)
    assert rows.tolist() == [[5, 6, 2, 8, 9, 10, 1, 0, 0, 0]]
    rows = cross_encoder_rows([[5] * 20], [[8] * 20], max_len=8)
    assert rows.tolist() == [[5] * 7 + [1]]
