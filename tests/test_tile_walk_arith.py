"""Host-side check of the arithmetic the tile walks rely on (csrc/gsr_common.cuh rect_rcp / rect_row): the row of
instance i inside a rectangle of w tiles per row is floor((i + 0.5) / w) evaluated in fp32 as (float(i) + 0.5f) * fl(1/w),
claimed exact below 2^22 tiles per rectangle.  numpy float32 uses the same IEEE round-to-nearest operations as the
__fadd_rn / __fmul_rn / __frcp_rn intrinsics of the kernel, so this is the kernel's arithmetic, not a model of it."""
import numpy as np


def _rows(i, w):
    rcp = np.float32(1.0) / np.float32(w)
    return ((i.astype(np.float32) + np.float32(0.5)) * rcp).astype(np.uint32)      # the cast truncates like F2I.TRUNC


def test_row_index_is_exact_for_every_width_up_to_256():
    for w in range(1, 257):
        n = min(w * 4096, 1 << 20)
        i = np.arange(n, dtype=np.uint32)
        assert np.array_equal(_rows(i, w), i // np.uint32(w)), w


def test_row_index_is_exact_at_row_boundaries_up_to_2_pow_22_tiles():
    rng = np.random.default_rng(0)
    widths = np.unique(np.concatenate([np.arange(1, 2049), rng.integers(2049, 65536, size=500)]))
    for w in widths:
        w = int(w)
        rows = max(1, min((1 << 22) // w, 65535))
        r = np.arange(rows, dtype=np.uint64)
        first = (r * w).astype(np.uint32)                 # first instance of every row
        last = (r * w + (w - 1)).astype(np.uint32)        # last instance of every row
        assert np.array_equal(_rows(first, w), r.astype(np.uint32)), w
        assert np.array_equal(_rows(last, w), r.astype(np.uint32)), w
