// cds_tables.h -- host-side construction of the integer match tables (IEEE double arithmetic, done once).
#ifndef CDS_TABLES_H
#define CDS_TABLES_H

#include <memory>
#include <vector>
#include "cds_common.h"

namespace cds {

struct RatioTable {
    std::vector<double> ratios;       // ascending distinct values of (double)a/(double)b, 0 <= a < b <= 255
    std::vector<uint16_t> rank;       // [a * 256 + b] -> index into ratios (valid for a < b)
};
const RatioTable &ratio_table();

// Sector of a colour exactly as the strict-inequality ladder of
// AbstractColorDepthSearchAlgorithm.calculatePixelGap decides it (:195-257): 0..5 = BR,BG,GB,GR,RG,RB, -1 = none.
// second/maxv are the numerator / denominator of the ratio.
inline int classify_color(int r, int g, int b, int &second, int &maxv)
{
    if (b > r && b > g) { maxv = b; if (r > g) { second = r; return 0; } second = g; return 1; }
    if (g > b && g > r) { maxv = g; if (b > r) { second = b; return 2; } second = r; return 3; }
    if (r > b && r > g) { maxv = r; if (g > b) { second = g; return 4; } second = b; return 5; }
    maxv = r > g ? (r > b ? r : b) : (g > b ? g : b);
    second = 0;
    return -1;
}

// Code word of a colour for a given data threshold (same function the encode kernel implements).
uint32_t encode_color(int r, int g, int b, int data_threshold);

// Match intervals (SR units) of every mask class for one zTolerance: index = sector * CDS_NUM_RANKS + rank.
struct ClassTable {
    double z_tolerance;
    uint32_t max_len;                  // longest interval (SR units); compact palettes need it <= CDS_PAL_MAX_LEN
    std::vector<cds_class_interval> iv;
};
std::shared_ptr<const ClassTable> class_table(double z_tolerance);
cds_class_interval class_interval(double z_tolerance, int sector, int rank);

}  // namespace cds
#endif
