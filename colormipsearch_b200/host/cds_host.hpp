// cds_host.hpp -- C++ mirror of the reference's scoring API over the C ABI of libcdsgpu (include/cdsgpu.h).
//
// The reference is Java; this image has no JVM, so the host layer that a Java maintainer would write with JNI / Panama FFM
// (java/ holds that source, INTEGRATION.md explains it) is mirrored here in C++ with the SAME type names, method names,
// argument meaning and error behaviour, so that the parity tests read like the reference's own JUnit tests:
//
//   reference (API/ = colormipsearch-api/src/main/java/org/janelia/colormipsearch/)          here
//   API/imageprocessing/ImageArray.java:12-68 (+ Color/Byte/ShortImageArray)                 ImageArray
//   API/imageprocessing/ImageRegionDefinition.java                                           ImageRegionDefinition
//   API/model/ComputeFileType.java:5-17                                                      ComputeFileType
//   API/cds/ColorDepthSearchParams.java:9-86                                                 ColorDepthSearchParams
//   API/cds/PixelMatchScore.java:18-30, ShapeMatchScore.java:12-65                           PixelMatchScore, ShapeMatchScore
//   API/cds/ColorDepthSearchAlgorithm.java:17-62                                             ColorDepthSearchAlgorithm<S>
//   API/cds/ColorDepthSearchAlgorithmProvider.java:12-38                                     ColorDepthSearchAlgorithmProvider<S>
//   API/cds/ColorDepthSearchAlgorithmProviderFactory.java:30-127                             ColorDepthSearchAlgorithmProviderFactory
//   API/cds/ColorMIPSearch.java:13-47                                                        ColorMIPSearch
//   API/cds/GradientAreaGapUtils.java:199-235                                                GradientAreaGapUtils
//   colormipsearch-tools/.../cmd/cdsprocess/ColorMIPSearchProcessor.java:8-12                GpuColorMIPSearchProcessor (batched seam)
//
// IllegalArgumentException -> std::invalid_argument, IllegalStateException -> std::runtime_error.
// Every score comes from the device: there is no host-side scoring code in this file.
#ifndef CDS_HOST_HPP
#define CDS_HOST_HPP

#include <cstdint>
#include <algorithm>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/cdsgpu.h"

namespace colormipsearch {

enum class ImageType { UNKNOWN = -1, GRAY8 = 0, GRAY16 = 1, RGB = 4 };   // ImagePlus.GRAY8 / GRAY16 / COLOR_RGB

// Pixel container crossing the boundary: RGB = interleaved R,G,B bytes, gray8 bytes, gray16 shorts (host byte order).
struct ImageArray {
    ImageType type = ImageType::UNKNOWN;
    int width = 0, height = 0;
    std::vector<uint8_t> bytes;       // RGB: 3*W*H, GRAY8: W*H, GRAY16: 2*W*H
    int getWidth() const { return width; }
    int getHeight() const { return height; }
    int getPixelCount() const { return width * height; }
    int get(int pi) const {           // ColorImageArray.get :25-31 / ByteImageArray / ShortImageArray
        switch (type) {
            case ImageType::RGB: return (int) (0xFF000000u | (bytes[3 * pi] << 16) | (bytes[3 * pi + 1] << 8) | bytes[3 * pi + 2]);
            case ImageType::GRAY8: return bytes[pi];
            case ImageType::GRAY16: return reinterpret_cast<const uint16_t *>(bytes.data())[pi];
            default: return 0;
        }
    }
    int getPixel(int x, int y) const { return (x >= 0 && x < width && y >= 0 && y < height) ? get(y * width + x) : 0; }   // :51-58
};

enum class ComputeFileType { InputColorDepthImage, SourceColorDepthImage, GradientImage, ZGapImage, Vol3DSegmentation, SkeletonSWC, SkeletonOBJ };

// The reference takes an arbitrary predicate (x, y) -> bool; every definition in the tools is a union of rectangles
// (colormipsearch-tools/.../cmd/AbstractColorDepthMatchArgs.java:101-119), which is what the device consumes.
using ImageRegionDefinition = std::function<std::vector<cds_rect>(const ImageArray &)>;

inline ImageRegionDefinition textLabelRegions(bool hasColorScaleLabel = true, bool hasNameLabel = true, int colorScaleWidth = 270)
{
    return [=](const ImageArray &img) {
        std::vector<cds_rect> r;
        if (hasColorScaleLabel && img.getWidth() > colorScaleWidth) r.push_back({img.getWidth() - colorScaleWidth, 0, img.getWidth(), 90});
        if (hasNameLabel) r.push_back({0, 0, 330, 100});
        return r;
    };
}

class ColorDepthSearchParams {
    std::map<std::string, double> num_;
    std::map<std::string, bool> has_;
public:
    ColorDepthSearchParams &setParam(const std::string &name, double v) { num_[name] = v; has_[name] = true; return *this; }
    ColorDepthSearchParams &setParam(const std::string &name, bool v) { num_[name] = v ? 1 : 0; has_[name] = true; return *this; }
    ColorDepthSearchParams &setParam(const std::string &name, int v) { num_[name] = v; has_[name] = true; return *this; }
    bool hasParam(const std::string &name) const { return has_.count(name) != 0; }
    int getIntParam(const std::string &name, int dflt) const { auto it = num_.find(name); return it == num_.end() ? dflt : (int) it->second; }
    double getDoubleParam(const std::string &name, double dflt) const { auto it = num_.find(name); return it == num_.end() ? dflt : it->second; }
    bool getBoolParam(const std::string &name, bool dflt) const { auto it = num_.find(name); return it == num_.end() ? dflt : it->second != 0; }
    const std::map<std::string, double> &asMap() const { return num_; }
};

struct ColorDepthMatchScore {
    virtual ~ColorDepthMatchScore() = default;
    virtual int getScore() const = 0;
    virtual float getNormalizedScore() const = 0;
    virtual bool isMirrored() const = 0;
};

class PixelMatchScore : public ColorDepthMatchScore {          // API/cds/PixelMatchScore.java
    int matchingPixNum_; double ratio_; bool mirrored_;
public:
    PixelMatchScore(int n, double ratio, bool mirrored) : matchingPixNum_(n), ratio_(ratio), mirrored_(mirrored) {}
    int getScore() const override { return matchingPixNum_; }
    float getNormalizedScore() const override { return (float) ratio_; }
    bool isMirrored() const override { return mirrored_; }
};

struct GradientAreaGapUtils {                                  // API/cds/GradientAreaGapUtils.java:199-235 (arithmetic lives in libcdsgpu)
    static long long calculate2DShapeScore(long long gap, long long highExpr) { return cds_shape_score_2d(gap, highExpr); }
    static double calculateNormalizedScore(int pix, long long shape, long long maxPix, long long maxShape) { return cds_normalized_score(pix, shape, maxPix, maxShape); }
};

class ShapeMatchScore : public ColorDepthMatchScore {          // API/cds/ShapeMatchScore.java
    long long gap_, highExpr_, maxGap_; bool mirrored_;
public:
    ShapeMatchScore(long long gap, long long highExpr, long long maxGap, bool mirrored) : gap_(gap), highExpr_(highExpr), maxGap_(maxGap), mirrored_(mirrored) {}
    int getScore() const override { return (int) GradientAreaGapUtils::calculate2DShapeScore(gap_, highExpr_); }
    float getNormalizedScore() const override { long long s = getScore(); return maxGap_ > 0 ? s / (float) maxGap_ : (float) s; }
    bool isMirrored() const override { return mirrored_; }
    long long getGradientAreaGap() const { return gap_; }
    long long getHighExpressionArea() const { return highExpr_; }
};

using VariantSuppliers = std::map<ComputeFileType, std::function<std::shared_ptr<ImageArray>()>>;

template <class S>
class ColorDepthSearchAlgorithm {                              // API/cds/ColorDepthSearchAlgorithm.java:17-62
public:
    virtual ~ColorDepthSearchAlgorithm() = default;
    virtual const ImageArray &getQueryImage() const = 0;
    virtual int getQuerySize() const = 0;
    virtual std::set<ComputeFileType> getRequiredTargetVariantTypes() const = 0;
    virtual S calculateMatchingScore(const ImageArray &targetImageArray, const VariantSuppliers &variantImageSuppliers) = 0;
};

template <class S>
class ColorDepthSearchAlgorithmProvider {                      // API/cds/ColorDepthSearchAlgorithmProvider.java:12-38
public:
    virtual ~ColorDepthSearchAlgorithmProvider() = default;
    virtual const ColorDepthSearchParams &getDefaultCDSParams() const = 0;
    virtual std::shared_ptr<ColorDepthSearchAlgorithm<S>> createColorDepthSearchAlgorithm(const ImageArray &queryImage, int queryThreshold,
                                                                                        int queryBorderSize, const ColorDepthSearchParams &cdsParams) = 0;
    std::shared_ptr<ColorDepthSearchAlgorithm<S>> createColorDepthQuerySearchAlgorithmWithDefaultParams(const ImageArray &queryImage, int queryThreshold,
                                                                                                      int queryBorderSize)
    {
        return createColorDepthSearchAlgorithm(queryImage, queryThreshold, queryBorderSize, ColorDepthSearchParams());
    }
};

// One context per process, shared by every provider (lifetime of a command run).
class GpuContext {
    cds_ctx *ctx_ = nullptr;
public:
    explicit GpuContext(int n_dev = 1) { check(cds_ctx_create(nullptr, n_dev, &ctx_), nullptr); }
    ~GpuContext() { cds_ctx_destroy(ctx_); }
    GpuContext(const GpuContext &) = delete;
    GpuContext &operator=(const GpuContext &) = delete;
    cds_ctx *get() const { return ctx_; }
    static void check(cds_status st, cds_ctx *ctx)
    {
        if (st == CDS_OK) return;
        std::string msg = cds_last_error(ctx);
        if (st == CDS_ERR_BAD_ARG || st == CDS_ERR_SIZE_MISMATCH) throw std::invalid_argument(msg);   // IllegalArgumentException
        throw std::runtime_error(msg);                                                               // IllegalStateException
    }
};

inline void requireRGB(const ImageArray &img, const char *what)
{
    if (img.type != ImageType::RGB) throw std::invalid_argument(std::string(what) + " must be an RGB image");
}

// PixelMatchColorDepthSearchAlgorithm (API/cds/PixelMatchColorDepthSearchAlgorithm.java) backed by a one-mask device mask set.
class GpuPixelMatchColorDepthSearchAlgorithm : public ColorDepthSearchAlgorithm<PixelMatchScore> {
    std::shared_ptr<GpuContext> gpu_;
    ImageArray query_;
    cds_maskset *ms_ = nullptr;
    int querySize_ = 0;
public:
    GpuPixelMatchColorDepthSearchAlgorithm(std::shared_ptr<GpuContext> gpu, const ImageArray &queryImage, int queryThreshold, bool mirrorQuery,
                                           int targetThreshold, double zTolerance, int xyShift, const ImageRegionDefinition &excludedRegions)
        : gpu_(std::move(gpu)), query_(queryImage)
    {
        requireRGB(queryImage, "query");
        cds_pixparams p{};
        p.mask_threshold = queryThreshold; p.data_threshold = targetThreshold; p.z_tolerance = zTolerance;
        p.xy_shift = xyShift; p.mirror = mirrorQuery ? 1 : 0;
        std::vector<cds_rect> rects = excludedRegions ? excludedRegions(queryImage) : std::vector<cds_rect>();
        if (rects.size() > CDS_MAX_RECTS) throw std::invalid_argument("too many excluded regions");
        p.n_rects = (int32_t) rects.size();
        for (size_t i = 0; i < rects.size(); i++) p.rects[i] = rects[i];
        GpuContext::check(cds_maskset_create(gpu_->get(), queryImage.width, queryImage.height, &p, &ms_), gpu_->get());
        int32_t size = 0;
        cds_status st = cds_maskset_add_rgb(ms_, queryImage.bytes.data(), 1, &size);
        if (st != CDS_OK) { cds_maskset_destroy(ms_); ms_ = nullptr; GpuContext::check(st, gpu_->get()); }
        querySize_ = size;
    }
    ~GpuPixelMatchColorDepthSearchAlgorithm() override { cds_maskset_destroy(ms_); }
    const ImageArray &getQueryImage() const override { return query_; }
    int getQuerySize() const override { return querySize_; }
    std::set<ComputeFileType> getRequiredTargetVariantTypes() const override { return {}; }
    PixelMatchScore calculateMatchingScore(const ImageArray &target, const VariantSuppliers &) override
    {
        int32_t score = 0, mirrored = 0;
        double ratio = 0;
        if (querySize_ != 0) requireRGB(target, "target");
        GpuContext::check(cds_score_pair_rgb(gpu_->get(), ms_, 0, target.bytes.data(), target.width, target.height, &score, &ratio, &mirrored), gpu_->get());
        return PixelMatchScore(score, ratio, mirrored != 0);
    }
};

// Shape2DMatchColorDepthSearchAlgorithm (API/cds/Shape2DMatchColorDepthSearchAlgorithm.java) backed by a one-mask shape mask set.
class GpuShape2DMatchColorDepthSearchAlgorithm : public ColorDepthSearchAlgorithm<ShapeMatchScore> {
    std::shared_ptr<GpuContext> gpu_;
    ImageArray query_;
    cds_shape_maskset *sms_ = nullptr;
    long long qmSize_ = 0, heSize_ = 0;
public:
    GpuShape2DMatchColorDepthSearchAlgorithm(std::shared_ptr<GpuContext> gpu, const ImageArray &queryImage, int queryThreshold, int queryBorderSize,
                                             bool mirrorQuery, const ImageArray *roiMask, const ImageRegionDefinition &excludedRegions)
        : gpu_(std::move(gpu)), query_(queryImage)
    {
        requireRGB(queryImage, "query");
        std::vector<cds_rect> rects = excludedRegions ? excludedRegions(queryImage) : std::vector<cds_rect>();
        GpuContext::check(cds_shape_maskset_create(gpu_->get(), queryImage.width, queryImage.height, queryThreshold, queryBorderSize, mirrorQuery ? 1 : 0,
                                                   rects.data(), (int32_t) rects.size(), roiMask ? roiMask->bytes.data() : nullptr, &sms_), gpu_->get());
        int64_t qm = 0, he = 0;
        cds_status st = cds_shape_maskset_add_rgb(sms_, queryImage.bytes.data(), 1, &qm, &he);
        if (st != CDS_OK) { cds_shape_maskset_destroy(sms_); sms_ = nullptr; GpuContext::check(st, gpu_->get()); }
        qmSize_ = qm; heSize_ = he;
    }
    ~GpuShape2DMatchColorDepthSearchAlgorithm() override { cds_shape_maskset_destroy(sms_); }
    const ImageArray &getQueryImage() const override { return query_; }
    int getQuerySize() const override { return (int) qmSize_; }
    long long getQueryMaskSize() const { return qmSize_; }                 // sum of the gray > 2 query mask (17340 in the reference test)
    long long getHighExpressionMaskSize() const { return heSize_; }        // sum of the high-expression mask (70640)
    std::set<ComputeFileType> getRequiredTargetVariantTypes() const override { return {ComputeFileType::GradientImage, ComputeFileType::ZGapImage}; }
    ShapeMatchScore calculateMatchingScore(const ImageArray &target, const VariantSuppliers &variants) override
    {
        auto fetch = [&](ComputeFileType t) -> std::shared_ptr<ImageArray> {
            auto it = variants.find(t);
            return it == variants.end() || !it->second ? nullptr : it->second();
        };
        std::shared_ptr<ImageArray> grad = fetch(ComputeFileType::GradientImage), zgap = fetch(ComputeFileType::ZGapImage);
        if (!grad || !zgap) return ShapeMatchScore(-1, -1, -1, false);     // :155-158
        requireRGB(target, "target");
        requireRGB(*zgap, "zgap");
        std::vector<uint16_t> g16((size_t) grad->width * grad->height);
        if (grad->type == ImageType::GRAY16) std::copy((const uint16_t *) grad->bytes.data(), (const uint16_t *) grad->bytes.data() + g16.size(), g16.begin());
        else if (grad->type == ImageType::GRAY8) std::copy(grad->bytes.begin(), grad->bytes.end(), g16.begin());
        else throw std::invalid_argument("gradient must be a gray image");
        if (grad->width != target.width || grad->height != target.height || zgap->width != target.width || zgap->height != target.height ||
            target.width != query_.width || target.height != query_.height)
            throw std::invalid_argument("Invalid image size - target, gradient and zgap images must match the query's size");
        int32_t pm = 0; int64_t pt = 0, gap = 0, he = 0; uint8_t mir = 0;
        GpuContext::check(cds_shape_score_pairs(gpu_->get(), sms_, target.bytes.data(), g16.data(), zgap->bytes.data(), nullptr, 1, &pm, &pt, 1, &gap, &he, &mir), gpu_->get());
        return ShapeMatchScore(gap, he, -1, mir != 0);
    }
};

class ColorDepthSearchAlgorithmProviderFactory {                // API/cds/ColorDepthSearchAlgorithmProviderFactory.java
public:
    static std::shared_ptr<ColorDepthSearchAlgorithmProvider<PixelMatchScore>> createPixMatchCDSAlgorithmProvider(
        std::shared_ptr<GpuContext> gpu, bool mirrorMask, int targetThreshold, double pixColorFluctuation, int xyShiftParam,
        ImageRegionDefinition ignoredRegionsProvider)
    {
        struct P : ColorDepthSearchAlgorithmProvider<PixelMatchScore> {
            std::shared_ptr<GpuContext> gpu; bool mirror; int thr; double fluct; int xy; ImageRegionDefinition regions; ColorDepthSearchParams dflt;
            const ColorDepthSearchParams &getDefaultCDSParams() const override { return dflt; }
            std::shared_ptr<ColorDepthSearchAlgorithm<PixelMatchScore>> createColorDepthSearchAlgorithm(const ImageArray &query, int queryThreshold, int,
                                                                                                      const ColorDepthSearchParams &params) override
            {
                const double zTolerance = params.getDoubleParam("pixColorFluctuation", fluct) / 100;          // :55-56
                const int xyShift = params.getIntParam("xyShift", xy);
                if ((xyShift & 0x1) == 1) throw std::invalid_argument("XY shift parameter must be an even number.");   // :57-60
                return std::make_shared<GpuPixelMatchColorDepthSearchAlgorithm>(gpu, query, queryThreshold, params.getBoolParam("mirrorMask", mirror),
                                                                              params.getIntParam("dataThreshold", thr), zTolerance, xyShift, regions);
            }
        };
        auto p = std::make_shared<P>();
        p->gpu = std::move(gpu); p->mirror = mirrorMask; p->thr = targetThreshold; p->fluct = pixColorFluctuation; p->xy = xyShiftParam;
        p->regions = std::move(ignoredRegionsProvider);
        p->dflt.setParam("mirrorMask", mirrorMask).setParam("dataThreshold", targetThreshold).setParam("pixColorFluctuation", pixColorFluctuation).setParam("xyShift", xyShiftParam);
        return p;
    }

    static std::shared_ptr<ColorDepthSearchAlgorithmProvider<ShapeMatchScore>> createShapeMatchCDSAlgorithmProvider(
        std::shared_ptr<GpuContext> gpu, bool mirrorMask, std::shared_ptr<ImageArray> roiMaskImageArray, ImageRegionDefinition excludedRegions)
    {
        struct P : ColorDepthSearchAlgorithmProvider<ShapeMatchScore> {
            std::shared_ptr<GpuContext> gpu; bool mirror; std::shared_ptr<ImageArray> roi; ImageRegionDefinition regions; ColorDepthSearchParams dflt;
            const ColorDepthSearchParams &getDefaultCDSParams() const override { return dflt; }
            std::shared_ptr<ColorDepthSearchAlgorithm<ShapeMatchScore>> createColorDepthSearchAlgorithm(const ImageArray &query, int queryThreshold, int queryBorderSize,
                                                                                                      const ColorDepthSearchParams &params) override
            {
                return std::make_shared<GpuShape2DMatchColorDepthSearchAlgorithm>(gpu, query, queryThreshold, queryBorderSize,
                                                                                params.getBoolParam("mirrorMask", mirror), roi.get(), regions);
            }
        };
        auto p = std::make_shared<P>();
        p->gpu = std::move(gpu); p->mirror = mirrorMask; p->roi = std::move(roiMaskImageArray); p->regions = std::move(excludedRegions);
        p->dflt.setParam("mirrorMask", mirrorMask);
        return p;
    }
};

class ColorMIPSearch {                                          // API/cds/ColorMIPSearch.java:13-47
    std::shared_ptr<ColorDepthSearchAlgorithmProvider<PixelMatchScore>> provider_;
    int defaultQueryThreshold_; double pctPositivePixels_;
public:
    ColorMIPSearch(double pctPositivePixels, int defaultQueryThreshold, std::shared_ptr<ColorDepthSearchAlgorithmProvider<PixelMatchScore>> provider)
        : provider_(std::move(provider)), defaultQueryThreshold_(defaultQueryThreshold), pctPositivePixels_(pctPositivePixels) {}
    std::shared_ptr<ColorDepthSearchAlgorithm<PixelMatchScore>> createQueryColorDepthSearchWithDefaultThreshold(const ImageArray &query)
    { return provider_->createColorDepthQuerySearchAlgorithmWithDefaultParams(query, defaultQueryThreshold_, 0); }
    std::shared_ptr<ColorDepthSearchAlgorithm<PixelMatchScore>> createQueryColorDepthSearch(const ImageArray &query, int queryThreshold, int borderSize)
    { return provider_->createColorDepthQuerySearchAlgorithmWithDefaultParams(query, queryThreshold, borderSize); }
    bool isMatch(const PixelMatchScore &s) const
    {
        const double pixMatchRatioThreshold = pctPositivePixels_ / 100;
        return s.getScore() > 0 && s.getNormalizedScore() > pixMatchRatioThreshold;    // :42-45
    }
    double pctPositivePixels() const { return pctPositivePixels_; }
};

// One kept pair of the batched search: what AbstractColorMIPSearchProcessor.findPixelMatch turns into a CDMatchEntity
// (colormipsearch-tools/.../cmd/cdsprocess/AbstractColorMIPSearchProcessor.java:60-84).
struct CDMatch {
    int maskIndex; long long targetIndex; int matchingPixels; float matchingPixelsRatio; bool mirrored;
};

// The batched seam: ColorMIPSearchProcessor.findAllColorDepthMatches(masks, targets)
// (colormipsearch-tools/.../cmd/cdsprocess/ColorMIPSearchProcessor.java:8-12) as ONE native search over device-resident data.
class GpuColorMIPSearchProcessor {
    std::shared_ptr<GpuContext> gpu_;
    cds_pixparams params_{};
    double pctPositivePixels_;
public:
    GpuColorMIPSearchProcessor(std::shared_ptr<GpuContext> gpu, bool mirrorMask, int maskThreshold, int dataThreshold, double pixColorFluctuation, int xyShift,
                               double pctPositivePixels, const std::vector<cds_rect> &rects)
        : gpu_(std::move(gpu)), pctPositivePixels_(pctPositivePixels)
    {
        if (xyShift & 1) throw std::invalid_argument("XY shift parameter must be an even number.");
        params_.mask_threshold = maskThreshold; params_.data_threshold = dataThreshold; params_.z_tolerance = pixColorFluctuation / 100;
        params_.xy_shift = xyShift; params_.mirror = mirrorMask ? 1 : 0; params_.n_rects = (int32_t) rects.size();
        for (size_t i = 0; i < rects.size() && i < CDS_MAX_RECTS; i++) params_.rects[i] = rects[i];
    }
    // every pair that passes ColorMIPSearch.isMatch, per mask in descending matchingPixels (ties: ascending target); maxPerMask <= 0 keeps
    // all of them (the reference's behaviour), a positive value only the best maxPerMask of each mask
    std::vector<CDMatch> findAllColorDepthMatches(const std::vector<const ImageArray *> &masks, const std::vector<const ImageArray *> &targets, int maxPerMask)
    {
        std::vector<CDMatch> out;
        if (masks.empty() || targets.empty()) return out;
        const int W = masks[0]->width, H = masks[0]->height;
        cds_maskset *ms = nullptr;
        GpuContext::check(cds_maskset_create(gpu_->get(), W, H, &params_, &ms), gpu_->get());
        try {
            std::vector<int32_t> sizes(masks.size());
            for (size_t i = 0; i < masks.size(); i++) {
                requireRGB(*masks[i], "mask");
                if (masks[i]->width != W || masks[i]->height != H) throw std::invalid_argument("all masks must have the same size");
                GpuContext::check(cds_maskset_add_rgb(ms, masks[i]->bytes.data(), 1, &sizes[i]), gpu_->get());
            }
            // targets: one contiguous host buffer, streamed to the devices in chunks (upload of chunk i+1 overlaps the search of chunk i)
            const size_t imgBytes = (size_t) W * H * 3;
            std::vector<uint8_t> all(imgBytes * targets.size());
            for (size_t i = 0; i < targets.size(); i++) {
                requireRGB(*targets[i], "target");
                if (targets[i]->width != W || targets[i]->height != H)
                    throw std::invalid_argument("Invalid image size - target's image size must match query's image size");
                std::copy(targets[i]->bytes.begin(), targets[i]->bytes.end(), all.begin() + i * imgBytes);
            }
            if (maxPerMask <= 0) {
                // the reference keeps EVERY pair that passes isMatch (LocalColorMIPSearchProcessor.java:93-105)
                int64_t cap = std::max<int64_t>(1024, 4 * (int64_t) masks.size()), n = 0;
                std::vector<int32_t> mk, sc; std::vector<int64_t> tg; std::vector<uint8_t> mir;
                for (int attempt = 0; attempt < 2; attempt++) {
                    mk.resize(cap); sc.resize(cap); tg.resize(cap); mir.resize(cap);
                    cds_status st = cds_search_stream_matches_rgb(gpu_->get(), ms, all.data(), (int64_t) targets.size(), pctPositivePixels_, cap,
                                                                  mk.data(), tg.data(), sc.data(), mir.data(), &n);
                    if (st == CDS_ERR_CAPACITY && attempt == 0) { cap = n; continue; }
                    GpuContext::check(st, gpu_->get());
                    break;
                }
                for (int64_t i = 0; i < n; i++)
                    out.push_back({mk[i], tg[i], sc[i], (float) ((double) sc[i] / (double) sizes[mk[i]]), mir[i] != 0});
            } else {
                const int K = std::max(1, std::min<int>(maxPerMask, (int) targets.size()));
                std::vector<int32_t> score((size_t) masks.size() * K), count(masks.size());
                std::vector<int64_t> target((size_t) masks.size() * K);
                std::vector<uint8_t> mir((size_t) masks.size() * K);
                GpuContext::check(cds_search_stream_rgb(gpu_->get(), ms, all.data(), (int64_t) targets.size(), K, pctPositivePixels_,
                                                        score.data(), target.data(), mir.data(), count.data()), gpu_->get());
                for (size_t m = 0; m < masks.size(); m++)
                    for (int i = 0; i < count[m]; i++) {
                        const size_t o = m * K + i;
                        out.push_back({(int) m, target[o], score[o], (float) ((double) score[o] / (double) sizes[m]), mir[o] != 0});
                    }
            }
        } catch (...) { cds_maskset_destroy(ms); throw; }
        cds_maskset_destroy(ms);
        return out;
    }

    // The same search for masks and targets that are TIFF FILES (the bytes as stored: PackBits or uncompressed RGB): nothing is
    // decoded on the host, the files are uploaded as they are and decoded on the devices (cds_maskset_add_tiff,
    // cds_search_stream_matches_tiff / cds_search_stream_tiff).  Replaces NeuronMIPUtils.loadComputeFile + ImageArrayUtils.readImageArray
    // (colormipsearch-api/.../mips/NeuronMIPUtils.java:62-103, imageprocessing/ImageArrayUtils.java:98-258) in front of the search.
    std::vector<CDMatch> findAllColorDepthMatchesInFiles(int width, int height, const std::vector<std::vector<uint8_t>> &maskFiles,
                                                         const std::vector<std::vector<uint8_t>> &targetFiles, int maxPerMask)
    {
        std::vector<CDMatch> out;
        if (maskFiles.empty() || targetFiles.empty()) return out;
        auto pack = [](const std::vector<std::vector<uint8_t>> &files, std::vector<uint8_t> &blob, std::vector<int64_t> &offsets) {
            offsets.assign(1, 0);
            for (const auto &f : files) offsets.push_back(offsets.back() + (int64_t) f.size());
            blob.resize((size_t) offsets.back() + 64);
            for (size_t i = 0; i < files.size(); i++) std::copy(files[i].begin(), files[i].end(), blob.begin() + offsets[i]);
        };
        std::vector<uint8_t> mblob, tblob;
        std::vector<int64_t> moff, toff;
        pack(maskFiles, mblob, moff);
        pack(targetFiles, tblob, toff);
        cds_maskset *ms = nullptr;
        GpuContext::check(cds_maskset_create(gpu_->get(), width, height, &params_, &ms), gpu_->get());
        try {
            std::vector<int32_t> sizes(maskFiles.size());
            GpuContext::check(cds_maskset_add_tiff(ms, mblob.data(), moff.data(), (int32_t) maskFiles.size(), sizes.data()), gpu_->get());
            const int64_t T = (int64_t) targetFiles.size();
            if (maxPerMask <= 0) {
                int64_t cap = std::max<int64_t>(1024, 4 * (int64_t) maskFiles.size()), n = 0;
                std::vector<int32_t> mk, sc; std::vector<int64_t> tg; std::vector<uint8_t> mir;
                for (int attempt = 0; attempt < 2; attempt++) {
                    mk.resize(cap); sc.resize(cap); tg.resize(cap); mir.resize(cap);
                    cds_status st = cds_search_stream_matches_tiff(gpu_->get(), ms, tblob.data(), toff.data(), T, pctPositivePixels_, cap,
                                                                   mk.data(), tg.data(), sc.data(), mir.data(), &n);
                    if (st == CDS_ERR_CAPACITY && attempt == 0) { cap = n; continue; }
                    GpuContext::check(st, gpu_->get());
                    break;
                }
                for (int64_t i = 0; i < n; i++)
                    out.push_back({mk[i], tg[i], sc[i], (float) ((double) sc[i] / (double) sizes[mk[i]]), mir[i] != 0});
            } else {
                const int K = std::max(1, std::min<int>(maxPerMask, (int) T));
                std::vector<int32_t> score(maskFiles.size() * K), count(maskFiles.size());
                std::vector<int64_t> target(maskFiles.size() * K);
                std::vector<uint8_t> mir(maskFiles.size() * K);
                GpuContext::check(cds_search_stream_tiff(gpu_->get(), ms, tblob.data(), toff.data(), T, K, pctPositivePixels_,
                                                         score.data(), target.data(), mir.data(), count.data()), gpu_->get());
                for (size_t m = 0; m < maskFiles.size(); m++)
                    for (int i = 0; i < count[m]; i++) {
                        const size_t o = m * K + i;
                        out.push_back({(int) m, target[o], score[o], (float) ((double) score[o] / (double) sizes[m]), mir[o] != 0});
                    }
            }
        } catch (...) { cds_maskset_destroy(ms); throw; }
        cds_maskset_destroy(ms);
        return out;
    }
};

}  // namespace colormipsearch
#endif
