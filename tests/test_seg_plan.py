"""Host logic of the time-axis segmentation (csrc/seg_scan.cu) through the C ABI: no GPU needed."""
import ctypes
import itertools

from rwkv_lm_ext_b200 import _lib


def plan(B, T, H, training):
    lib = _lib.load()
    n, sc = ctypes.c_int(0), ctypes.c_int(0)
    lib.wkv6b200_seg_plan(B, T, H, int(training), ctypes.byref(n), ctypes.byref(sc))
    return n.value, sc.value


def test_plan_invariants():
    for B, T, H, training in itertools.product((1, 2, 3, 8), (64, 1000, 2048, 4096, 4100, 16384, 65536), (1, 2, 12, 32, 40, 64),
                                               (False, True)):
        nseg, sc = plan(B, T, H, training)
        NC = (T + 63) // 64
        assert nseg >= 1 and sc >= 1
        if nseg == 1:
            assert sc == NC
            continue
        assert B * H * nseg <= 296                       # every row is resident at once (148 SMs x 2 CTAs)
        assert nseg * sc >= NC and (nseg - 1) * sc < NC  # the segments cover the sequence, none is empty
        assert sc >= (4 if training else 2)
        assert T >= (2048 if training else 4096)
        assert B * H <= (74 if training else 147)


def test_benchmark_shape_is_not_segmented_and_3b_infctx_is():
    assert plan(8, 4096, 32, True) == (1, 64)
    assert plan(8, 4096, 32, False) == (1, 64)
    assert plan(1, 4096, 40, True) == (7, 10)            # BASELINE config 5: 7 x 40 = 280 rows of <= 10 chunks
    assert plan(1, 2048, 32, True)[0] == 8


def test_saved_buffer_covers_the_segmented_checkpoints():
    lib = _lib.load()
    for B, T, H in ((1, 4096, 40), (1, 2048, 32), (2, 8192, 3), (8, 4096, 32)):
        nseg, sc = plan(B, T, H, True)
        slots = nseg * sc if nseg > 1 else (T + 63) // 64
        need = 4 * (B * H + B * H * nseg) + B * H * slots * 8192
        assert lib.wkv6_saved_bytes(B, T, H * 64, H) >= need
