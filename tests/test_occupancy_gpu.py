"""The occupancy tile rows (what the candidate kernel's ticket tests and scans read) against their definition, for both occupancy
kernels: bit (x, y) of sector s = some pixel of sector s at (x + dx, y + dy) for one of the search's shift offsets (dx, dy)."""
import numpy as np
import pytest

from colormipsearch_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(n_dev=1)
    yield c
    c.close()


def reference_occupancy(valid, W, H, shift):
    n, _, S, vp = valid.shape
    tp = ((W + 7) // 8 + 3) // 4 * 4
    nzw = ((6 * tp + 31) // 32 + 3) // 4 * 4
    HT = (H + 3) // 4
    bits = np.unpackbits(valid.view(np.uint8).reshape(n, H, S, vp, 4), axis=-1, bitorder="little").reshape(n, H, S, vp * 32).astype(bool)
    wide = vp * 32
    dil = np.zeros((n, HT * 4, S, wide), bool)
    # xyShift 2: the 3 x 3 offsets of ring 2; xyShift 4: ring 2 and ring 4 (17 distinct offsets, cds_cand.cu offset_of)
    offsets = {(0, 0)}
    for ring in range(2, shift + 1, 2):
        offsets |= {(dy, dx) for dy in (-ring, 0, ring) for dx in (-ring, 0, ring)}
    for dy, dx in sorted(offsets):
        if True:
            ys0, ys1 = max(0, -dy), min(H, H - dy)            # source rows y + dy inside the image
            xs0, xs1 = max(0, -dx), min(wide, wide - dx)
            dil[:, ys0:ys1, :, xs0:xs1] |= bits[:, ys0 + dy:ys1 + dy, :, xs0 + dx:xs1 + dx]
    dil[:, H:] = False
    out = np.zeros((n, HT, 7 * tp + nzw), np.uint32)
    # tile (ty, tx): bit r * 8 + c = pixel (8 tx + c, 4 ty + r)
    t = dil[:, :, :, :tp * 8].reshape(n, HT, 4, S, tp, 8).transpose(0, 1, 3, 4, 2, 5).reshape(n, HT, S, tp, 32)
    words = np.packbits(t, axis=-1, bitorder="little").view(np.uint32).reshape(n, HT, S, tp)
    out[:, :, :6 * tp] = words.reshape(n, HT, 6 * tp)
    out[:, :, 6 * tp:7 * tp] = np.bitwise_or.reduce(words, axis=2)
    nz = (words.reshape(n, HT, 6 * tp) != 0)
    nzbits = np.zeros((n, HT, nzw * 32), bool)
    nzbits[:, :, :6 * tp] = nz
    out[:, :, 7 * tp:] = np.packbits(nzbits, axis=-1, bitorder="little").view(np.uint32).reshape(n, HT, nzw)
    return out


@pytest.mark.parametrize("W,H", [(1210, 566), (97, 50), (640, 33), (2048, 8), (31, 5)])
@pytest.mark.parametrize("shift", [0, 2, 4])
def test_occupancy_rows_equal_definition(ctx, W, H, shift):
    rng = np.random.default_rng(W * 7 + H + shift)
    vp = ((W + 31) // 32 + 3) // 4 * 4
    n = 3
    sector = rng.integers(0, 6, (n, H, W))
    lit = rng.random((n, H, W)) < np.array([0.002, 0.05, 0.5])[:, None, None]
    lit[0, 0, 0] = lit[0, H - 1, W - 1] = lit[0, 0, W - 1] = lit[0, H - 1, 0] = True          # the corners
    bits = np.zeros((n, H, 6, vp * 32), bool)
    for s in range(6):
        bits[:, :, s, :W] = lit & (sector == s)
    valid = np.packbits(bits, axis=-1, bitorder="little").view(np.uint32).reshape(n, H, 6, vp)
    want = reference_occupancy(valid, W, H, shift)
    for version in (1, 0):
        ctx.set_option("occupancy_kernel", version)
        got = ctx.debug_occupancy(valid, W, H, shift)
        assert np.array_equal(got, want), (version, W, H, shift, np.argwhere(got != want)[:5])
    ctx.set_option("occupancy_kernel", 1)
