"""Result wire format: the JSON the reference's writer produces, restated in colormipsearch_b200/wire.py.
Pinned on the reference's own sample file (tests/golden/ref_cdsmatches_sample.json = colormipsearch-persist/src/test/resources/
cdsmatches/testcdsmatches.json): Jackson's pretty-printer layout, CDMatchEntity's field names and order, Java's Float.toString."""
import json
import os
from collections import OrderedDict

import numpy as np
import pytest

from colormipsearch_b200 import wire

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SAMPLE = os.path.join(ROOT, "tests", "golden", "ref_cdsmatches_sample.json")


@pytest.fixture(scope="module")
def sample():
    text = open(SAMPLE, encoding="utf-8").read()
    return text, wire.parse_reference_json(text)


def test_serialiser_reproduces_the_reference_file_byte_for_byte(sample):
    text, matches = sample
    assert wire.jackson_pretty(matches) == text.rstrip("\n")


def test_java_float_to_string():
    # values Float.toString is documented / known to print this way
    cases = {0.03888351: "0.03888351", 1e-4: "1.0E-4", 1.0: "1.0", 12345678.0: "1.2345678E7", 0.001: "0.001", 9999999.0: "9999999.0",
             1e7: "1.0E7", 3.4028235e38: "3.4028235E38", 0.0: "0.0", 100.0: "100.0", 0.024040014: "0.024040014"}
    for v, s in cases.items():
        assert wire.java_float_str(v) == s, (v, s)
    # every ratio of the sample survives a float32 round trip through the printer
    for m in wire.parse_reference_json(open(SAMPLE, encoding="utf-8").read()):
        assert np.float32(float(wire.java_float_str(m["matchingPixelsRatio"]))) == np.float32(m["matchingPixelsRatio"])


def test_grouping_by_mask_like_the_reference_writer(sample):
    _, matches = sample
    groups = wire.group_matches_by_mask(matches)
    ids = [m["maskImage"]["mipId"] for m in matches]
    assert list(groups) == list(OrderedDict.fromkeys(ids))                    # one group per mask mipId
    assert sum(len(g["results"]) for g in groups.values()) == len(matches)
    for key, g in groups.items():
        mine = [m for m in matches if m["maskImage"]["mipId"] == key]
        # the mask keeps its identity but not the three compute files that move into the matches
        assert g["inputImage"]["mipId"] == key
        for gone in ("InputColorDepthImage", "GradientImage", "ZGapImage"):
            assert gone not in g["inputImage"].get("computeFiles", {})
        assert "SourceColorDepthImage" in g["inputImage"]["computeFiles"]
        px = [r["matchingPixels"] for r in g["results"]]
        assert px == sorted(px, reverse=True)
        for r in g["results"]:
            assert "maskImage" not in r
            # the files of THIS match's own mask (masks that share a mipId may differ in their variants)
            own = [m for m in mine if m["image"]["mipId"] == r["image"]["mipId"] and m["matchingPixels"] == r["matchingPixels"]]
            src = own[0]["maskImage"]["computeFiles"]
            assert r["matchComputeFiles"]["MaskColorDepthImage"] == src["InputColorDepthImage"]
            assert r["matchComputeFiles"]["MaskGradientImage"] == src["GradientImage"]
            assert r["matchComputeFiles"]["MaskZGapImage"] == src["ZGapImage"]
            assert list(r)[-1] == "class" and r["class"] == wire.CD_MATCH_CLASS


def test_sort_is_stable_and_unmatched_entries_are_dropped():
    mask = OrderedDict((("class", "org.janelia.colormipsearch.model.EMNeuronEntity"), ("mipId", "m1"),
                        ("computeFiles", OrderedDict((("InputColorDepthImage", "/m1.tif"),)))))
    def tgt(i):
        return OrderedDict((("class", "org.janelia.colormipsearch.model.LMNeuronEntity"), ("mipId", "t%d" % i)))
    ms = [wire.make_match(mask, tgt(0), 10, 100, False), wire.make_match(mask, tgt(1), 30, 100, True),
          wire.make_match(mask, tgt(2), 10, 100, False), wire.make_match(mask, None, 99, 100, False),
          wire.make_match(mask, tgt(3), 30, 100, False)]
    g = wire.group_matches_by_mask(ms)["m1"]
    assert [r["image"]["mipId"] for r in g["results"]] == ["t1", "t3", "t0", "t2"]      # ties keep their input order
    assert "computeFiles" not in g["inputImage"]                                        # nothing left -> NON_EMPTY drops the field
    assert g["results"][0]["matchComputeFiles"] == {"MaskColorDepthImage": "/m1.tif"}
    assert wire.java_float_str(g["results"][0]["matchingPixelsRatio"]) == "0.3"


def test_files_written_per_mask(tmp_path, sample):
    _, matches = sample
    n = wire.write_matches_by_mask(matches, str(tmp_path))
    names = sorted(os.listdir(tmp_path))
    assert n == len(names) == len({m["maskImage"]["mipId"] for m in matches})
    for name in names:
        text = open(tmp_path / name, encoding="utf-8").read()
        doc = wire.parse_reference_json(text)
        assert list(doc) == ["inputImage", "results"] and doc["inputImage"]["mipId"] + ".json" == name
        assert wire.jackson_pretty(doc) == text                                  # the printer is idempotent on its own output
        json.loads(text)


def test_matches_from_search_arrays():
    masks = [OrderedDict((("mipId", "m%d" % i),)) for i in range(2)]
    targets = [OrderedDict((("mipId", "t%d" % i),)) for i in range(4)]
    # the order cds_search_stream_matches_* returns: by mask, descending score, ascending target
    mi = np.array([0, 0, 0, 1]); ti = np.array([2, 0, 3, 1]); sc = np.array([50, 40, 40, 7]); mr = np.array([0, 1, 0, 0])
    ms = wire.matches_from_search(masks, targets, [200, 70], mi, ti, sc, mr)
    g = wire.group_matches_by_mask(ms)
    assert [r["image"]["mipId"] for r in g["m0"]["results"]] == ["t2", "t0", "t3"]
    assert [r["mirrored"] for r in g["m0"]["results"]] == [False, True, False]
    assert wire.java_float_str(g["m1"]["results"][0]["matchingPixelsRatio"]) == "0.1"
    assert g["m0"]["results"][0]["matchingPixelsRatio"] == np.float32(50 / 200)


def _plain(v):
    """order-insensitive view of a JSON value"""
    if isinstance(v, dict):
        return {k: _plain(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_plain(x) for x in v]
    return float(v) if isinstance(v, float) else v


def test_expand_is_the_inverse_of_grouping(sample):
    """MatchEntitiesGrouping.expandResultsByMask undoes groupByMaskFields: the mask gets its three compute files back, the match
    compute files disappear -- on the reference's sample every match comes back as it was."""
    _, matches = sample
    groups = wire.group_matches_by_mask(matches)
    back = [m for g in groups.values() for m in wire.expand_results_by_mask(g)]
    assert len(back) == len(matches)
    def key(m):
        return (m["maskImage"]["mipId"], m["image"]["mipId"], m["matchingPixels"], m["maskImage"]["computeFiles"]["InputColorDepthImage"])
    want = {key(m): m for m in matches}
    for m in back:
        o = want[key(m)]
        # everything but the mask is the match as it was; the mask is the GROUP's key (the first match's mask, ItemsHandling.java:57)
        # with this match's own three compute files restored
        assert _plain({k: v for k, v in m.items() if k != "maskImage"}) == _plain({k: v for k, v in o.items() if k not in ("maskImage", "matchComputeFiles")})
        assert m["maskImage"]["mipId"] == o["maskImage"]["mipId"]
        for f in ("InputColorDepthImage", "GradientImage", "ZGapImage"):
            assert m["maskImage"]["computeFiles"][f] == o["maskImage"]["computeFiles"][f]


def test_grouping_by_target(sample, tmp_path):
    _, matches = sample
    groups = wire.group_matches_by_target(matches)
    assert set(groups) == {m["image"]["mipId"] for m in matches}
    for key, g in groups.items():
        assert g["inputImage"]["mipId"] == key
        for gone in ("InputColorDepthImage", "GradientImage", "ZGapImage"):
            assert gone not in g["inputImage"].get("computeFiles", {})
        px = [r["matchingPixels"] for r in g["results"]]
        assert px == sorted(px, reverse=True)
        for r in g["results"]:
            assert "maskImage" not in r
            assert r["image"]["class"].endswith("EMNeuronEntity")           # the original mask is now the result's image
            src = [m for m in matches if m["image"]["mipId"] == key][0]["image"]["computeFiles"]
            assert r["matchComputeFiles"]["MaskColorDepthImage"] == src["InputColorDepthImage"]
    assert wire.write_matches_by_target(matches, str(tmp_path)) == len(groups) == len(os.listdir(tmp_path))
