// cds_select.cpp -- the reference's selection of the matches that go on to shape scoring:
// ColorMIPProcessUtils.selectBestMatches (colormipsearch-tools/src/main/java/org/janelia/colormipsearch/cmd/cdsprocess/
// ColorMIPProcessUtils.java:12-34) = two nested ItemsHandling.selectTopRankedElements
// (colormipsearch-api/src/main/java/org/janelia/colormipsearch/results/ItemsHandling.java:80-109): group a mask's matches by the
// target's published name ("line"), keep the top lines by their best matchingPixels, inside each line group by neuron id
// ("sample"), keep the top samples, inside each sample keep the top matches.
//
// Host-side integer logic over the output of the searches; no device work.  What needs care is ORDER, because it decides which
// groups survive a cut when scores tie: groups are collected in a java.util.HashMap (Collectors.groupingBy) and then
// stable-sorted by score, so ties keep the HashMap's iteration order -- bucket index (spread(hash) & (capacity - 1)) ascending,
// insertion order inside a bucket, capacity = the table size after inserting that many keys.  The callers pass the Java
// hashCode() of every group key and this file reproduces that order.  (Bins that turn into trees -- >= 8 colliding keys in a
// table of >= 64 buckets -- are not modelled; String hashes do not produce them in practice.)
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/cdsgpu.h"
#include "cds_runtime.h"

namespace cds { void set_tls_error(const std::string &msg); }

namespace {

// keys (ids) in first-appearance order -> the same ids in java.util.HashMap iteration order
std::vector<int32_t> hashmap_order(const std::vector<int32_t> &ids_in_insertion_order, const int32_t *hash)
{
    const size_t n = ids_in_insertion_order.size();
    uint32_t cap = 16;
    while ((double) n > 0.75 * cap) cap <<= 1;          // resize happens when ++size > threshold = 0.75 * capacity
    std::vector<std::pair<uint32_t, uint32_t>> key(n);  // (bucket, insertion position)
    for (size_t i = 0; i < n; i++) {
        const uint32_t h = (uint32_t) hash[ids_in_insertion_order[i]];
        key[i] = {(h ^ (h >> 16)) & (cap - 1), (uint32_t) i};
    }
    std::sort(key.begin(), key.end());
    std::vector<int32_t> out(n);
    for (size_t i = 0; i < n; i++) out[i] = ids_in_insertion_order[key[i].second];
    return out;
}

struct Group { int32_t id; int32_t best; std::vector<int64_t> items; };

// ItemsHandling.selectTopRankedElements over `items` (indices into score[] / key[])
std::vector<Group> select_top_ranked(const std::vector<int64_t> &items, const int32_t *key, const int32_t *key_hash, int32_t n_keys,
                                     const int32_t *score, int32_t top_results, int32_t limit_sub_results, std::vector<int32_t> &slot_of_key)
{
    // groupingBy: lists in encounter order, keys remembered in first-appearance order
    std::vector<int32_t> first_seen;
    std::vector<Group> groups;
    for (int64_t it : items) {
        const int32_t k = key[it];
        if (slot_of_key[k] < 0) { slot_of_key[k] = (int32_t) groups.size(); first_seen.push_back(k); groups.push_back({k, 0, {}}); }
        groups[slot_of_key[k]].items.push_back(it);
    }
    for (Group &g : groups) {
        // r.sort(csrComparison.reversed()): stable, descending score
        std::stable_sort(g.items.begin(), g.items.end(), [&](int64_t a, int64_t b) { return score[a] > score[b]; });
        g.best = score[g.items.front()];                                         // Collections.max
        if (limit_sub_results > 0 && (size_t) limit_sub_results < g.items.size()) g.items.resize(limit_sub_results);
    }
    // entrySet().stream() runs in HashMap order; sorted(...) is stable, descending group score
    std::vector<Group> ordered;
    ordered.reserve(groups.size());
    for (int32_t k : hashmap_order(first_seen, key_hash)) ordered.push_back(std::move(groups[slot_of_key[k]]));
    for (int32_t k : first_seen) slot_of_key[k] = -1;                            // leave the scratch table clean
    (void) n_keys;
    std::stable_sort(ordered.begin(), ordered.end(), [](const Group &a, const Group &b) { return a.best > b.best; });
    if (top_results > 0 && ordered.size() > (size_t) top_results) ordered.resize(top_results);
    return ordered;
}

}  // namespace

extern "C" cds_status cds_select_best_matches(const int32_t *line, const int32_t *sample, const int32_t *score, int64_t n,
                                              const int32_t *line_hash, int32_t n_lines, const int32_t *sample_hash, int32_t n_samples,
                                              int32_t top_lines, int32_t top_samples_per_line, int32_t top_matches_per_sample,
                                              int64_t *selected, int64_t *n_selected)
{
    return cds::abi_guard("cds_select_best_matches", [&]() -> cds_status {
        if (n < 0 || n_lines < 0 || n_samples < 0 || !n_selected || (n > 0 && (!line || !sample || !score || !line_hash || !sample_hash || !selected))) {
            cds::set_tls_error("cds_select_best_matches: bad arguments");
            return CDS_ERR_BAD_ARG;
        }
        for (int64_t i = 0; i < n; i++)
            if (line[i] < 0 || line[i] >= n_lines || sample[i] < 0 || sample[i] >= n_samples) {
                cds::set_tls_error("cds_select_best_matches: group id out of range");
                return CDS_ERR_BAD_ARG;
            }
        std::vector<int64_t> all(n);
        std::iota(all.begin(), all.end(), 0);
        std::vector<int32_t> line_slot(n_lines, -1), sample_slot(n_samples, -1);
        int64_t out = 0;
        for (const Group &ln : select_top_ranked(all, line, line_hash, n_lines, score, top_lines, -1, line_slot))
            for (const Group &sm : select_top_ranked(ln.items, sample, sample_hash, n_samples, score, top_samples_per_line, top_matches_per_sample, sample_slot))
                for (int64_t it : sm.items) selected[out++] = it;
        *n_selected = out;
        return CDS_OK;
    });
}

// String.hashCode() of a UTF-16 string given as UTF-8 restricted to the Basic Multilingual Plane subset the names use (ASCII):
// s[0]*31^(n-1) + ... + s[n-1] in 32-bit arithmetic.  Convenience for C / Python callers; Java callers pass hashCode() directly.
extern "C" int32_t cds_java_string_hash(const char *ascii)
{
    uint32_t h = 0;
    if (ascii) for (const unsigned char *p = (const unsigned char *) ascii; *p; p++) h = 31u * h + *p;
    return (int32_t) h;
}
