// host_mirror_main.cpp -- replays the reference's own JUnit scenarios through the C++ mirror of its API
// (colormipsearch_b200/host/cds_host.hpp over libcdsgpu.so).  Driven by tests/test_host_mirror_gpu.py, which writes the
// fixture images as raw files and checks the printed lines against the golden vectors of
//   colormipsearch-api/src/test/java/org/janelia/colormipsearch/cds/PixelMatchColorDepthSearchAlgorithmTest.java:72-103
//   colormipsearch-api/src/test/java/org/janelia/colormipsearch/cds/Shape2DMatchColorDepthSearchAlgorithmTest.java:86-132
// usage: host_mirror <dir> ; <dir> holds W H in dims.txt and <name>.rgb / <name>.g16 raw images
#include <cstdio>
#include <fstream>
#include <iostream>
#include <iterator>
#include <sstream>

#include "../colormipsearch_b200/host/cds_host.hpp"

using namespace colormipsearch;

static ImageArray load(const std::string &dir, const std::string &name, ImageType type, int W, int H)
{
    ImageArray a;
    a.type = type; a.width = W; a.height = H;
    const size_t n = (size_t) W * H * (type == ImageType::RGB ? 3 : type == ImageType::GRAY16 ? 2 : 1);
    a.bytes.resize(n);
    std::ifstream f(dir + "/" + name, std::ios::binary);
    if (!f.read(reinterpret_cast<char *>(a.bytes.data()), (std::streamsize) n)) throw std::runtime_error("cannot read " + name);
    return a;
}

int main(int argc, char **argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: host_mirror <dir>\n"); return 2; }
    const std::string dir = argv[1];
    int W = 0, H = 0;
    { std::ifstream f(dir + "/dims.txt"); f >> W >> H; }
    try {
        auto gpu = std::make_shared<GpuContext>(1);
        const ImageArray em1 = load(dir, "em_12191.rgb", ImageType::RGB, W, H), em2 = load(dir, "em_12191_FL.rgb", ImageType::RGB, W, H);
        const ImageArray lmA = load(dir, "lm_VT033614.rgb", ImageType::RGB, W, H), lmB = load(dir, "lm_BJD.rgb", ImageType::RGB, W, H),
                         lmC = load(dir, "lm_VT016795.rgb", ImageType::RGB, W, H);

        // ---- PixelMatchColorDepthSearchAlgorithmTest.pixelMatchScore: provider(mirror, thr 20, pixColorFluctuation 1, xyShift 2)
        auto provider = ColorDepthSearchAlgorithmProviderFactory::createPixMatchCDSAlgorithmProvider(gpu, true, 20, 1.0, 2, textLabelRegions());
        struct Case { const ImageArray *m, *t; const char *name; };
        const Case cases[] = {{&em1, &lmA, "12191xVT033614"}, {&em1, &lmB, "12191xBJD"}, {&em2, &lmA, "FLxVT033614"},
                              {&em2, &lmC, "FLxVT016795"}, {&em1, &lmC, "12191xVT016795"}};
        ColorMIPSearch search(1.0, 20, provider);
        for (const Case &c : cases) {
            auto alg = provider->createColorDepthQuerySearchAlgorithmWithDefaultParams(*c.m, 20, 0);
            PixelMatchScore s = alg->calculateMatchingScore(*c.t, {});
            std::printf("pixel %s %d %d querySize %d isMatch %d ratio %.9g\n", c.name, s.getScore(), s.isMirrored() ? 1 : 0, alg->getQuerySize(),
                        search.isMatch(s) ? 1 : 0, (double) s.getNormalizedScore());
        }
        // IllegalArgumentException sites
        try {
            provider->createColorDepthSearchAlgorithm(em1, 20, 0, ColorDepthSearchParams().setParam("xyShift", 3));
            std::printf("odd_xyshift no_exception\n");
        } catch (const std::invalid_argument &e) { std::printf("odd_xyshift IllegalArgumentException %s\n", e.what()); }
        try {
            auto alg = provider->createColorDepthQuerySearchAlgorithmWithDefaultParams(em1, 20, 0);
            ImageArray small; small.type = ImageType::RGB; small.width = 10; small.height = 10; small.bytes.assign(300, 0);
            alg->calculateMatchingScore(small, {});
            std::printf("size_mismatch no_exception\n");
        } catch (const std::invalid_argument &e) { std::printf("size_mismatch IllegalArgumentException\n"); }
        {   // empty mask: (0, 0.0, false) and no exception even for a wrong-size target (PixelMatch...:169-175)
            ImageArray black = em1; std::fill(black.bytes.begin(), black.bytes.end(), 0);
            auto alg = provider->createColorDepthQuerySearchAlgorithmWithDefaultParams(black, 20, 0);
            PixelMatchScore s = alg->calculateMatchingScore(lmA, {});
            std::printf("empty_mask %d %d querySize %d\n", s.getScore(), s.isMirrored() ? 1 : 0, alg->getQuerySize());
        }

        // ---- the batched seam: findAllColorDepthMatches over the same images
        {
            GpuColorMIPSearchProcessor proc(gpu, true, 20, 20, 1.0, 2, 1.0, textLabelRegions()(em1));
            auto matches = proc.findAllColorDepthMatches({&em1, &em2}, {&lmA, &lmB, &lmC}, 3);
            for (const CDMatch &m : matches)
                std::printf("batched mask %d target %lld pixels %d mirrored %d\n", m.maskIndex, m.targetIndex, m.matchingPixels, m.mirrored ? 1 : 0);
            auto every = proc.findAllColorDepthMatches({&em1, &em2}, {&lmA, &lmB, &lmC}, 0);      // no limit: every isMatch pair
            for (const CDMatch &m : every)
                std::printf("allpairs mask %d target %lld pixels %d mirrored %d\n", m.maskIndex, m.targetIndex, m.matchingPixels, m.mirrored ? 1 : 0);
        }

        // ---- the same seam fed with TIFF files (no decode on the host): <name>.tif next to the raw images
        {
            auto slurp = [&](const std::string &name) {
                std::ifstream f(dir + "/" + name, std::ios::binary);
                if (!f) throw std::runtime_error("cannot read " + name);
                return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
            };
            GpuColorMIPSearchProcessor proc(gpu, true, 20, 20, 1.0, 2, 1.0, textLabelRegions()(em1));
            auto every = proc.findAllColorDepthMatchesInFiles(W, H, {slurp("em_12191.tif"), slurp("em_12191_FL.tif")},
                                                              {slurp("lm_VT033614.tif"), slurp("lm_BJD.tif"), slurp("lm_VT016795.tif")}, 0);
            for (const CDMatch &m : every)
                std::printf("tiffpairs mask %d target %lld pixels %d mirrored %d\n", m.maskIndex, m.targetIndex, m.matchingPixels, m.mirrored ? 1 : 0);
            auto best = proc.findAllColorDepthMatchesInFiles(W, H, {slurp("em_12191.tif")}, {slurp("lm_VT033614.tif"), slurp("lm_BJD.tif"), slurp("lm_VT016795.tif")}, 1);
            for (const CDMatch &m : best)
                std::printf("tiffbest mask %d target %lld pixels %d mirrored %d\n", m.maskIndex, m.targetIndex, m.matchingPixels, m.mirrored ? 1 : 0);
        }

        // ---- Shape2DMatchColorDepthSearchAlgorithmTest: provider(mirror), thr 20; zgap = the on-disk file for BJD
        {
            auto sprov = ColorDepthSearchAlgorithmProviderFactory::createShapeMatchCDSAlgorithmProvider(gpu, true, nullptr, textLabelRegions());
            auto salg = sprov->createColorDepthQuerySearchAlgorithmWithDefaultParams(em2, 20, 0);
            auto *impl = dynamic_cast<GpuShape2DMatchColorDepthSearchAlgorithm *>(salg.get());
            std::printf("shape_masks %lld %lld required %zu\n", impl->getQueryMaskSize(), impl->getHighExpressionMaskSize(), salg->getRequiredTargetVariantTypes().size());
            auto grad = std::make_shared<ImageArray>(load(dir, "grad_BJD.g16", ImageType::GRAY16, W, H));
            auto zgap = std::make_shared<ImageArray>(load(dir, "zgap_BJD.rgb", ImageType::RGB, W, H));
            VariantSuppliers v;
            v[ComputeFileType::GradientImage] = [grad]() { return grad; };
            v[ComputeFileType::ZGapImage] = [zgap]() { return zgap; };
            auto salg1 = sprov->createColorDepthQuerySearchAlgorithmWithDefaultParams(em1, 20, 0);
            ShapeMatchScore s = salg1->calculateMatchingScore(lmB, v);
            std::printf("shape 12191xBJD_zgapfile %lld %lld %d %d\n", s.getGradientAreaGap(), s.getHighExpressionArea(), s.getScore(), s.isMirrored() ? 1 : 0);
            ShapeMatchScore none = salg1->calculateMatchingScore(lmB, {});
            std::printf("shape_missing %lld %lld %d\n", none.getGradientAreaGap(), none.getHighExpressionArea(), none.getScore());
        }
        std::printf("normalized %.6f\n", GradientAreaGapUtils::calculateNormalizedScore(636, GradientAreaGapUtils::calculate2DShapeScore(156, 1897), 679, 1114361));
    } catch (const std::exception &e) {
        std::fprintf(stderr, "host_mirror failed: %s\n", e.what());
        return 1;
    }
    return 0;
}
