"""Result wire format: colour-depth matches as the JSON files the reference writes.

Host logic only.  Restates, for matches produced by this library,
  * the per-mask grouping and ordering of the reference's file writer
    (colormipsearch-persist/.../dataio/fs/JSONNeuronMatchesWriter.java:59-71 -> colormipsearch-api/.../results/
    MatchEntitiesGrouping.java:56-103 (groupByMaskFields) -> results/ItemsHandling.java:38-71 (groupItems)):
    one file <mask mipId>.json = {"inputImage": mask, "results": [matches]}; the mask loses its InputColorDepthImage /
    GradientImage / ZGapImage compute files, which move into every match's "matchComputeFiles" as MaskColorDepthImage /
    MaskGradientImage / MaskZGapImage; matches without a matched image are dropped; matches are STABLY sorted by descending
    matchingPixels (colormipsearch-tools/.../cmd/ColorDepthSearchCmd.java:403-409) and lose their "maskImage";
  * the field names and order of CDMatchEntity as Jackson emits them (the reference's sample
    colormipsearch-persist/src/test/resources/cdsmatches/testcdsmatches.json, kept as tests/golden/ref_cdsmatches_sample.json),
    Jackson's DefaultPrettyPrinter layout ('"key" : value', '[ {' ... '}, {' ... '} ]') and Java's Float.toString.
Pinned: serialising the parsed sample reproduces the sample file byte for byte (tests/test_wire_cpu.py).  The layout of the
GROUPED files has no sample in the reference tree; it follows the annotations of GroupedMatchedEntities ("inputImage", "results").
"""
import json
import os
from collections import OrderedDict

import numpy as np

MASK_COMPUTE_TO_MATCH = (("ZGapImage", "MaskZGapImage"), ("InputColorDepthImage", "MaskColorDepthImage"), ("GradientImage", "MaskGradientImage"))
MATCH_FIELD_ORDER = ("maskImage", "mirrored", "matchComputeFiles", "normalizedScore", "matchingPixels", "matchingPixelsRatio",
                     "bidirectionalAreaGap", "gradientAreaGap", "highExpressionArea", "image", "files", "class")
CD_MATCH_CLASS = "org.janelia.colormipsearch.model.CDMatchEntity"


class JavaFloat(float):
    """A value that came from (or goes to) a Java `Float`: printed like Float.toString."""


def java_float_str(x):
    """Float.toString: shortest decimal that identifies the float32, at least one fraction digit, computerised scientific
    notation outside [1e-3, 1e7)."""
    f = np.float32(x)
    if np.isnan(f):
        return "NaN"
    if np.isinf(f):
        return "Infinity" if f > 0 else "-Infinity"
    if f == 0:
        return "-0.0" if np.signbit(f) else "0.0"
    a = abs(float(f))
    if 1e-3 <= a < 1e7:
        s = np.format_float_positional(f, unique=True, trim="0")
        return s
    s = np.format_float_scientific(f, unique=True, trim="0", exp_digits=1)        # d.dddde-05
    mant, exp = s.split("e")
    if "." not in mant:
        mant += ".0"
    return "%sE%d" % (mant, int(exp))


def _scalar(v):
    if isinstance(v, bool):
        return "true" if v else "false"
    if v is None:
        return "null"
    if isinstance(v, JavaFloat) or isinstance(v, np.float32):
        return java_float_str(v)
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    if isinstance(v, float):
        return repr(v)                      # Java Double.toString agrees with repr for the values that occur here
    return json.dumps(v, ensure_ascii=False)


def jackson_pretty(v, indent=0):
    """Jackson's DefaultPrettyPrinter: objects one field per line, two spaces per level, ' : ' between key and value; arrays on
    one line with a space after '[' and before ']'."""
    pad = "  " * indent
    if isinstance(v, dict):
        if not v:
            return "{ }"
        items = [pad + "  " + json.dumps(k, ensure_ascii=False) + " : " + jackson_pretty(x, indent + 1) for k, x in v.items()]
        return "{\n" + ",\n".join(items) + "\n" + pad + "}"
    if isinstance(v, (list, tuple)):
        if not v:
            return "[ ]"
        return "[ " + ", ".join(jackson_pretty(x, indent) for x in v) + " ]"
    return _scalar(v)


def parse_reference_json(text):
    """json.loads that keeps field order and marks the Float fields of CDMatchEntity, so that jackson_pretty can reproduce them."""
    def hook(pairs):
        d = OrderedDict(pairs)
        for k in ("normalizedScore", "matchingPixelsRatio"):
            if k in d and isinstance(d[k], float):
                d[k] = JavaFloat(d[k])
        return d
    return json.loads(text, object_pairs_hook=hook)


def make_match(mask_entity, target_entity, matching_pixels, mask_size, mirrored, normalized_score=None, files=None):
    """One CDMatchEntity as AbstractColorMIPSearchProcessor.findPixelMatch fills it (cdsprocess/AbstractColorMIPSearchProcessor.java:60-84):
    matchingPixelsRatio = (float) ((double) matchingPixels / maskSize) (PixelMatchScore.getNormalizedScore)."""
    m = OrderedDict()
    m["maskImage"] = mask_entity
    m["mirrored"] = bool(mirrored)
    if normalized_score is not None:
        m["normalizedScore"] = JavaFloat(np.float32(normalized_score))
    m["matchingPixels"] = int(matching_pixels)
    m["matchingPixelsRatio"] = JavaFloat(np.float32(float(matching_pixels) / float(mask_size))) if mask_size else JavaFloat(0.0)
    m["image"] = target_entity
    if files:
        m["files"] = files
    m["class"] = CD_MATCH_CLASS
    return m


def _ordered(match):
    out = OrderedDict()
    for k in MATCH_FIELD_ORDER:
        if k in match and match[k] not in (None, {}, []):
            out[k] = match[k]
    for k, v in match.items():
        if k not in out and v not in (None, {}, []):
            out[k] = v
    return out


def group_matches_by_mask(matches):
    """MatchEntitiesGrouping.groupByMaskFields with the writer's arguments: key = mask mipId, filter = has a matched image,
    ranking = descending matchingPixels (stable).  Returns {mipId: {"inputImage": ..., "results": [...]}} in first-seen order."""
    groups = OrderedDict()
    for m in matches:
        mask = m.get("maskImage")
        if mask is None:
            continue
        key = mask.get("mipId")
        g = groups.get(key)
        if g is None:
            input_image = OrderedDict(mask)                      # the key comes from the group's first item (ItemsHandling.java:57)
            cf = OrderedDict(mask.get("computeFiles", {}))
            for src, _ in MASK_COMPUTE_TO_MATCH:
                cf.pop(src, None)
            if cf:
                input_image["computeFiles"] = cf
            else:
                input_image.pop("computeFiles", None)            # @JsonInclude(NON_EMPTY), AbstractBaseEntity.java:19
            g = groups[key] = {"inputImage": input_image, "results": []}
        if m.get("image") is None:
            continue
        r = OrderedDict((k, v) for k, v in m.items() if k != "maskImage")
        mcf = OrderedDict(r.get("matchComputeFiles", {}))
        for src, dst in MASK_COMPUTE_TO_MATCH:
            v = mask.get("computeFiles", {}).get(src)
            if v is not None:
                mcf[dst] = v
        if mcf:
            r["matchComputeFiles"] = mcf
        g["results"].append(_ordered(r))
    for g in groups.values():
        g["results"].sort(key=lambda r: -float(r.get("matchingPixels") or 0))     # list.sort is stable, like List.sort in Java
    return groups


def group_matches_by_target(matches):
    """MatchEntitiesGrouping.groupByTargetFields (MatchEntitiesGrouping.java:118-150) with the writer's arguments: the roles are
    swapped -- the matched image becomes the group's key ("inputImage"), the original mask becomes every result's "image", and
    the match compute files are taken from the MATCHED image's compute files."""
    swapped = []
    for m in matches:
        tgt = m.get("image")
        if tgt is None:
            continue
        r = OrderedDict((k, v) for k, v in m.items() if k not in ("maskImage", "image", "matchComputeFiles"))
        r["maskImage"] = tgt
        r["image"] = m.get("maskImage")
        swapped.append(_ordered(r))
    return group_matches_by_mask(swapped)


def expand_results_by_mask(group):
    """MatchEntitiesGrouping.expandResultsByMask (:152-175), the reader's inverse of the grouping: every result gets the group's
    mask back, with the three compute files restored from its matchComputeFiles, which are then dropped."""
    out = []
    for r in group["results"]:
        mask = OrderedDict(group["inputImage"])
        cf = OrderedDict(mask.get("computeFiles", {}))
        for src, dst in MASK_COMPUTE_TO_MATCH:
            v = r.get("matchComputeFiles", {}).get(dst)
            if v is not None:
                cf[src] = v
        if cf:
            mask["computeFiles"] = cf
        m = OrderedDict((k, v) for k, v in r.items() if k != "matchComputeFiles")
        m["maskImage"] = mask
        out.append(_ordered(m))
    return out


def write_matches_by_target(matches, out_dir):
    """JSONNeuronMatchesWriter.writeMatchesByTarget: one <matched mipId>.json per matched image."""
    os.makedirs(out_dir, exist_ok=True)
    groups = group_matches_by_target(matches)
    for key, g in groups.items():
        if not key or not str(key).strip():
            continue
        doc = OrderedDict((("inputImage", g["inputImage"]), ("results", g["results"])))
        with open(os.path.join(out_dir, "%s.json" % key), "w", encoding="utf-8") as f:
            f.write(jackson_pretty(doc))
    return len(groups)


def write_matches_by_mask(matches, out_dir):
    """JSONNeuronMatchesWriter.writeMatchesByMask: one <mipId>.json per mask.  Returns the number of files written."""
    os.makedirs(out_dir, exist_ok=True)
    groups = group_matches_by_mask(matches)
    for key, g in groups.items():
        if not key or not str(key).strip():
            continue                                             # ItemsWriterToJSONFile.getJsonFile: blank name -> no file
        doc = OrderedDict((("inputImage", g["inputImage"]), ("results", g["results"])))
        with open(os.path.join(out_dir, "%s.json" % key), "w", encoding="utf-8") as f:
            f.write(jackson_pretty(doc))
    return len(groups)


def matches_from_search(mask_entities, target_entities, mask_sizes, mask_idx, target_idx, scores, mirrored):
    """The arrays of cds_search_stream_matches_* / cds_search_matches -> CDMatchEntity records, in the order given."""
    return [make_match(mask_entities[int(mi)], target_entities[int(ti)], int(sc), int(mask_sizes[int(mi)]), bool(mr))
            for mi, ti, sc, mr in zip(mask_idx, target_idx, scores, mirrored)]
