"""f1: the reference's best-lines / best-samples / best-matches selection, C ABI vs the oracle's restatement, pinned on
the scenarios of the reference's own ItemsHandlingTest
(colormipsearch-api/src/test/java/org/janelia/colormipsearch/results/ItemsHandlingTest.java:17-163)."""
import numpy as np
import pytest

from colormipsearch_b200 import capi
from oracle import oracle as O


def reference_test_data():
    """createTestData(), ItemsHandlingTest.java:123-163: 3 lines x 4 samples x 3 matches."""
    data = []
    for li, top in ((1, 45), (2, 44), (3, 43)):
        for si in range(4):
            for k in range(3):
                data.append(("l%d" % li, "s%d.%d" % (li, si + 1), top - 10 * si - k))
    return data


def test_java_string_hash():
    for s in ("", "l1", "UNKNOWN", "GMR_31G04_AE_01", "s3.4"):
        assert capi.java_string_hash(s) == (O.java_string_hash(s) ^ 0x80000000) - 0x80000000
    assert capi.java_string_hash("l1") == 3397 and capi.java_string_hash("UNKNOWN") == 433141802      # values a JVM prints


@pytest.mark.parametrize("top,sub", [(-1, -1), (-1, 1), (-1, 2), (-1, 3), (1, -1), (2, -1), (3, -1), (1, 1), (2, 3), (3, 2)])
def test_reference_scenarios_select_top_ranked(top, sub):
    """The assertions of selectAllElements*, selectTopRankedElementsWith*SubResults, on the oracle."""
    data = reference_test_data()
    by_line = {}
    for m in data:
        by_line.setdefault(m[0], []).append(m)
    ranked = O.select_top_ranked(list(data), lambda m: m[0], lambda m: m[2], top, sub)
    assert len(ranked) == (len(by_line) if top <= 0 else top)
    for name, best, items in ranked:
        assert len(items) == (len(by_line[name]) if sub <= 0 else sub)
        assert all(m[2] <= best for m in by_line[name]) and all(m[0] == name for m in items)
    assert [e[1] for e in ranked] == sorted((e[1] for e in ranked), reverse=True)


def _run_both(matches, a, b, c):
    exp = O.select_best_matches(list(matches), a, b, c)
    got = capi.select_best_matches([m[0] for m in matches], [m[1] for m in matches], [m[2] for m in matches], a, b, c)
    assert [matches[i] for i in got.tolist()] == exp
    return exp


@pytest.mark.parametrize("limits", [(0, 0, 0), (1, 1, 1), (2, 3, 2), (3, 2, 0), (2, 0, 1), (300, 1, 1)])
def test_select_best_matches_reference_data(limits):
    data = [m + (i,) for i, m in enumerate(reference_test_data())]
    exp = _run_both(data, *limits)
    a, b, c = limits
    lines = []
    for m in exp:
        if m[0] not in lines:
            lines.append(m[0])
    assert len(lines) == (3 if a <= 0 else min(a, 3))
    assert lines == ["l1", "l2", "l3"][: len(lines)]                 # best scores 45 > 44 > 43


def test_select_best_matches_random_with_ties():
    """Many ties, blank names, tables that grow past 16 and 32 buckets: the order of equal-score groups is the HashMap's."""
    rng = np.random.default_rng(12)
    for trial in range(30):
        n_lines, n_samples = int(rng.integers(1, 60)), int(rng.integers(1, 90))
        n = int(rng.integers(1, 400))
        lines = ["" if rng.random() < 0.03 else "R%dG%02d" % (rng.integers(10, 99), j) for j in rng.integers(0, n_lines, n)]
        samples = ["%d" % (2000000 + 7919 * int(j)) for j in rng.integers(0, n_samples, n)]
        scores = rng.integers(1, 12, n).tolist()                    # few distinct scores -> ties everywhere
        matches = [(l, s, sc, i) for i, (l, s, sc) in enumerate(zip(lines, samples, scores))]
        for limits in ((0, 0, 0), (5, 2, 1), (int(rng.integers(1, 20)), int(rng.integers(0, 4)), int(rng.integers(0, 3)))):
            _run_both(matches, *limits)


def test_select_bad_arguments():
    with pytest.raises(capi.CdsIllegalArgument):
        import ctypes as C
        cnt = C.c_int64(0)
        one = np.zeros(1, np.int32)
        capi._check(capi.lib().cds_select_best_matches(one.ctypes.data_as(capi._i32p), one.ctypes.data_as(capi._i32p), one.ctypes.data_as(capi._i32p), 1,
                                                       one.ctypes.data_as(capi._i32p), 0, one.ctypes.data_as(capi._i32p), 1, 0, 0, 0,
                                                       np.zeros(1, np.int64).ctypes.data_as(capi._i64p), C.byref(cnt)))
