// cds_tables.cpp -- see cds_tables.h.  Compile with -ffp-contract=off: the interval end points must come from the
// same unfused IEEE double operations the Java reference performs.
#include "cds_tables.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>

namespace cds {

const RatioTable &ratio_table()
{
    static RatioTable table = [] {
        RatioTable t;
        std::vector<double> all;
        all.reserve(256 * 255 / 2);
        for (int b = 1; b <= 255; b++)
            for (int a = 0; a < b; a++) all.push_back((double) a / (double) b);
        std::sort(all.begin(), all.end());
        all.erase(std::unique(all.begin(), all.end()), all.end());
        t.ratios = all;
        t.rank.assign(256 * 256, 0);
        for (int b = 1; b <= 255; b++)
            for (int a = 0; a < b; a++) {
                double v = (double) a / (double) b;
                t.rank[a * 256 + b] = (uint16_t) (std::lower_bound(all.begin(), all.end(), v) - all.begin());
            }
        return t;
    }();
    return table;
}

uint32_t encode_color(int r, int g, int b, int data_threshold)
{
    int second, maxv;
    int sector = classify_color(r, g, b, second, maxv);
    uint32_t sr = sector < 0 ? (uint32_t) CDS_SR_NONE
                             : (uint32_t) sector * CDS_SECTOR_STRIDE + ratio_table().rank[second * 256 + maxv];
    uint32_t code = (sr << CDS_CODE_SR_SHIFT) | (uint32_t) maxv;
    if (!(maxv > data_threshold)) code |= CDS_CODE_BELOW_BIT;
    return code;
}

namespace {

// Sector-boundary constants and guards of calculatePixelGap (:183-187 and the guards at :272, :289, :295, :312, :318,
// :335, :341, :359, :365, :382).  For the ordered pair (mask sector s1, target sector s2) adjacent on the colour wheel:
//   low  kind: r1 < g(s1) && r2 < g(s2) && (r1 - c) + (r2 - c) <= tol    (true on a prefix of target ranks)
//   high kind: r1 > 0.8   && r2 > 0.8   && (c - r1) + (c - r2) <= tol    (true on a suffix of target ranks)
struct Adjacency { int s1, s2; bool high; double c, g1, g2; };
const double BrBg = 0.354862745, BgGb = 0.996078431, GbGr = 0.505882353, GrRg = 0.996078431, RgRb = 0.505882353;
const Adjacency kAdj[] = {
    {0, 1, false, BrBg, 0.44, 0.54},   // BR mask, BG data
    {1, 0, false, BrBg, 0.54, 0.44},   // BG mask, BR data
    {1, 2, true, BgGb, 0.8, 0.8},      // BG mask, GB data
    {2, 1, true, BgGb, 0.8, 0.8},      // GB mask, BG data
    {2, 3, false, GbGr, 0.7, 0.7},     // GB mask, GR data
    {3, 2, false, GbGr, 0.7, 0.7},     // GR mask, GB data
    {3, 4, true, GrRg, 0.8, 0.8},      // GR mask, RG data
    {4, 3, true, GrRg, 0.8, 0.8},      // RG mask, GR data
    {4, 5, false, RgRb, 0.7, 0.7},     // RG mask, RB data
    {5, 4, false, RgRb, 0.7, 0.7},     // RB mask, RG data
};

inline bool same_sector_match(double r1, double r2, double tol)
{
    if (!(r1 > 0 && r2 > 0)) return false;
    double gap = (r1 != r2) ? std::fabs(r2 - r1) : 0;
    return gap <= tol;
}

inline bool adjacent_match(const Adjacency &a, double r1, double r2, double tol)
{
    if (a.high) {
        if (!(r1 > a.g1 && r2 > a.g2)) return false;
        double gap1 = a.c - r1;
        double gap2 = a.c - r2;
        return gap1 + gap2 <= tol;
    }
    if (!(r1 < a.g1 && r2 < a.g2)) return false;
    double gap1 = r1 - a.c;
    double gap2 = r2 - a.c;
    return gap1 + gap2 <= tol;
}

cds_class_interval compute_interval(double tol, int s1, int k1)
{
    const std::vector<double> &R = ratio_table().ratios;
    const int NR = (int) R.size();
    const double r1 = R[k1];
    cds_class_interval out = {CDS_IV_EMPTY, 0, CDS_IV_EMPTY, 0};

    // own sector: matches form a contiguous run of ranks around k1 (fl(r2 - r1) is monotone in r2)
    if (same_sector_match(r1, r1, tol)) {
        int lo = k1, hi = k1;
        {   // first rank in [1, k1] that matches
            int a = 1, b = k1;       // invariant: b matches
            while (a < b) { int m = (a + b) / 2; if (same_sector_match(r1, R[m], tol)) b = m; else a = m + 1; }
            lo = b;
        }
        {   // last rank in [k1, NR-1] that matches
            int a = k1, b = NR - 1;  // invariant: a matches
            while (a < b) { int m = (a + b + 1) / 2; if (same_sector_match(r1, R[m], tol)) a = m; else b = m - 1; }
            hi = a;
        }
        out.lo1 = (uint32_t) (s1 * CDS_SECTOR_STRIDE + lo);
        out.len1 = (uint32_t) (hi - lo);
    }

    // neighbouring sector: the two guards of a sector are mutually exclusive, so at most one neighbour can match
    for (const Adjacency &a : kAdj) {
        if (a.s1 != s1) continue;
        int lo = -1, hi = -1;
        if (a.high) {
            if (adjacent_match(a, r1, R[NR - 1], tol)) {
                int x = 0, y = NR - 1;   // invariant: y matches; find first match
                while (x < y) { int m = (x + y) / 2; if (adjacent_match(a, r1, R[m], tol)) y = m; else x = m + 1; }
                lo = y; hi = NR - 1;
            }
        } else {
            if (adjacent_match(a, r1, R[0], tol)) {
                int x = 0, y = NR - 1;   // invariant: x matches; find last match
                while (x < y) { int m = (x + y + 1) / 2; if (adjacent_match(a, r1, R[m], tol)) x = m; else y = m - 1; }
                lo = 0; hi = x;
            }
        }
        if (lo >= 0) {
            // (cannot happen twice: guards exclusive) keep the first, the table self-test flags a second one
            if (out.lo2 == CDS_IV_EMPTY) {
                out.lo2 = (uint32_t) (a.s2 * CDS_SECTOR_STRIDE + lo);
                out.len2 = (uint32_t) (hi - lo);
            } else {
                out.lo2 = CDS_IV_EMPTY - 1;  // poison, detected by class_table()
            }
        }
    }
    return out;
}

std::mutex g_mu;
std::map<uint64_t, std::shared_ptr<const ClassTable>> g_tables;

}  // namespace

cds_class_interval class_interval(double z_tolerance, int sector, int rank)
{
    return compute_interval(z_tolerance, sector, rank);
}

std::shared_ptr<const ClassTable> class_table(double z_tolerance)
{
    uint64_t key;
    std::memcpy(&key, &z_tolerance, sizeof key);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_tables.find(key);
        if (it != g_tables.end()) return it->second;
    }
    auto t = std::make_shared<ClassTable>();
    t->z_tolerance = z_tolerance;
    t->iv.resize((size_t) CDS_NUM_CLASSES);
    t->max_len = 0;
    const int NR = (int) ratio_table().ratios.size();
    for (int s = 0; s < CDS_NUM_SECTORS; s++)
        for (int k = 0; k < NR; k++) {
            cds_class_interval iv = compute_interval(z_tolerance, s, k);
            if (iv.lo2 == CDS_IV_EMPTY - 1) return nullptr;
            if (iv.lo1 != CDS_IV_EMPTY) t->max_len = std::max(t->max_len, iv.len1);
            if (iv.lo2 != CDS_IV_EMPTY) t->max_len = std::max(t->max_len, iv.len2);
            t->iv[(size_t) s * CDS_NUM_RANKS + k] = iv;
        }
    std::lock_guard<std::mutex> lk(g_mu);
    g_tables[key] = t;
    return t;
}

}  // namespace cds
