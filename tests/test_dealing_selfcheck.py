"""Host-only check of the search's work decomposition (rma_chamfer_dealing_selfcheck, csrc/chamfer_filter.cu: FDeal): no GPU
needed, so the geometry that decides which virtual CTA computes which (row block x column run) - and where the resolve
looks for the records - is covered on CPU for many shapes, SM counts and forced run lengths."""
import ctypes
import random

import pytest

from rma_net_b200 import _lib

SHAPES = [(16, 4096, 4096), (16, 2048, 2048), (64, 1024, 1024), (28, 2048, 2048), (512, 2048, 2048), (4, 2048, 2048),
          (16, 8192, 8192), (4, 16384, 16384), (1, 65536, 65536), (128, 8192, 8192), (1, 33, 2050), (3, 700, 1300),
          (2, 2100, 95), (1, 1, 1), (5, 513, 64), (7, 1025, 4097), (200, 300, 5000), (37, 5000, 129), (1, 1 << 18, 4096)]


def _check(B, N, M, sms, unit_run=0):
    lib = _lib.load()
    mode = ctypes.c_int(-1)
    load = (ctypes.c_int64 * 2)()
    o = ctypes.byref(_lib.ChamferOpts(_lib.PATH_FILTER, unit_run, 0))
    bad = lib.rma_chamfer_dealing_selfcheck(B, N, M, sms, o, ctypes.byref(mode), ctypes.cast(load, ctypes.c_void_p))
    return bad, mode.value, load[0], load[1]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("sms", [148, 132, 8])
def test_every_unit_is_covered_once(shape, sms):
    bad, mode, lo, hi = _check(*shape, sms)
    assert bad == 0, (shape, sms, mode)
    assert mode in (0, 1) and 0 <= lo <= hi


def test_forced_run_lengths_and_random_shapes():
    rng = random.Random(7)
    for _ in range(150):
        B, N, M = rng.randint(1, 40), rng.randint(1, 6000), rng.randint(1, 6000)
        sms = rng.choice([148, 132, 64, 3])
        run = rng.choice([0, 0, 1, 2, 3, 5, 17])
        bad, mode, lo, hi = _check(B, N, M, sms, run)
        assert bad == 0, (B, N, M, sms, run, mode)
        if run:
            assert mode == 0


def test_balance_of_the_benchmark_shapes():
    """The dealing's point: the most loaded SM carries at most one unit more than the average asks for (rounded up)."""
    for (B, N, M), want_mode in (((16, 4096, 4096), 1), ((16, 2048, 2048), 1), ((28, 2048, 2048), 1),
                                 ((512, 2048, 2048), 0)):
        bad, mode, lo, hi = _check(B, N, M, 148)
        rows_per_cta = 2048 if N > 1024 else (1024 if N > 512 else 512)
        units = B * -(-N // rows_per_cta) * (-(-M // 64) * 2)
        assert bad == 0 and mode == want_mode
        assert hi <= -(-units // 148) + 1, (B, N, M, lo, hi, units)
