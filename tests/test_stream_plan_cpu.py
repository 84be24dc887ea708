"""The chunk plan of the streaming searches (cds_stream.cu: stream_chunk_plan) -- host logic, checked without a device: every target in
exactly one chunk and in order, chunks round-robin over the devices, the ramp of a search over files, equal chunks after it, the byte cap."""
import numpy as np
import pytest

from colormipsearch_b200 import capi


def check_cover(dev, first, cnt, n, D):
    assert first.tolist() == np.concatenate([[0], np.cumsum(cnt)[:-1]]).tolist() if len(cnt) else n == 0
    assert int(cnt.sum()) == n and (cnt > 0).all()
    assert dev.tolist() == [i % D for i in range(len(dev))]


@pytest.mark.parametrize("D", [1, 2, 3, 8])
@pytest.mark.parametrize("n", [0, 1, 255, 256, 257, 1000, 12500, 100003])
def test_pixels_plan_is_equal_chunks(D, n):
    dev, first, cnt = capi.stream_plan(D, n, 256)
    check_cover(dev, first, cnt, n, D)
    assert (cnt[:-1] == 256).all()


@pytest.mark.parametrize("D", [1, 2, 4])
@pytest.mark.parametrize("n", [1, 100, 256, 300, 4000, 12500, 12544, 50000])
@pytest.mark.parametrize("chunk", [7, 256, 1024, 4096])
def test_files_plan_ramp_then_equal_chunks(D, n, chunk):
    off = np.arange(n + 1, dtype=np.int64) * 130000            # files of equal size
    dev, first, cnt = capi.stream_plan(D, n, chunk, off)
    check_cover(dev, first, cnt, n, D)
    assert cnt.max() <= chunk
    ramp_end = 0
    for i, c in enumerate(cnt[:-1]):                           # (the last chunk takes what is left)
        want = 256 << min(i // D, 8)
        if want < chunk:
            assert c == want, (i, c)                           # the ramp: 256, 512, ... per device
            ramp_end = i + 1
    rest = cnt[ramp_end:]
    if len(rest) > 1:
        assert rest[:-1].max() - rest[:-1].min() == 0          # equal chunks after the ramp ...
        assert rest[-1] <= rest[0] and rest[-1] > rest[0] - len(rest) * D - 1      # ... and the last one is not a stump


def test_files_plan_respects_the_byte_cap():
    rng = np.random.default_rng(5)
    n = 5000
    sizes = rng.integers(1000, 400000, n)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    cap = 50_000_000
    dev, first, cnt = capi.stream_plan(2, n, 4096, off, cap)
    check_cover(dev, first, cnt, n, 2)
    bytes_per_chunk = off[first + cnt] - off[first]
    assert ((bytes_per_chunk <= cap) | (cnt == 1)).all()
    # a cap below every file: one file per chunk
    dev, first, cnt = capi.stream_plan(1, 100, 4096, off[:101], 1)
    assert (cnt == 1).all() and len(cnt) == 100


def test_plan_arguments():
    with pytest.raises(capi.CdsIllegalArgument):
        capi.stream_plan(0, 10, 256)
    with pytest.raises(capi.CdsIllegalArgument):
        capi.stream_plan(1, 10, 0)
