package org.janelia.colormipsearch.cds;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.ValueLayout;
import java.util.ArrayList;
import java.util.Collections;
import java.util.List;
import java.util.Map;
import java.util.Set;
import java.util.function.BiPredicate;
import java.util.function.Supplier;

import javax.annotation.Nonnull;

import org.janelia.colormipsearch.cds.gpu.CdsGpu;
import org.janelia.colormipsearch.imageprocessing.ImageArray;
import org.janelia.colormipsearch.imageprocessing.ImageArrayAccess;
import org.janelia.colormipsearch.imageprocessing.ImageRegionDefinition;
import org.janelia.colormipsearch.model.ComputeFileType;

/**
 * Drop-in for the provider returned by ColorDepthSearchAlgorithmProviderFactory.createPixMatchCDSAlgorithmProvider
 * (ColorDepthSearchAlgorithmProviderFactory.java:30-74): same parameters, same per-mask overrides, same exceptions; the
 * scoring runs in libcdsgpu (no JVM-side compute).  Selected with `--cds-provider gpu` (INTEGRATION.md).
 *
 * This class serves the literal single-pair API (one native call per calculateMatchingScore); throughput comes from
 * GpuColorMIPSearchProcessor, which hands whole mask / target lists to one native search.
 *
 * UNVERIFIED (no JDK in the build image of the GPU library).
 */
public class GpuPixelMatchColorDepthSearchAlgorithmProvider implements ColorDepthSearchAlgorithmProvider<PixelMatchScore> {
    private final ColorDepthSearchParams defaults = new ColorDepthSearchParams();
    private final ImageRegionDefinition ignoredRegionsProvider;

    public GpuPixelMatchColorDepthSearchAlgorithmProvider(boolean mirrorMask, int targetThreshold, double pixColorFluctuation,
                                                          int xyShift, ImageRegionDefinition ignoredRegionsProvider) {
        if (pixColorFluctuation < 0 || pixColorFluctuation > 100) throw new IllegalArgumentException("Invalid value for pixel color fluctuation " + pixColorFluctuation);
        defaults.setParam("mirrorMask", mirrorMask).setParam("dataThreshold", targetThreshold)
                .setParam("pixColorFluctuation", pixColorFluctuation).setParam("xyShift", xyShift);
        this.ignoredRegionsProvider = ignoredRegionsProvider;
    }

    @Override
    public ColorDepthSearchParams getDefaultCDSParams() { return defaults; }

    @Override
    public ColorDepthSearchAlgorithm<PixelMatchScore> createColorDepthSearchAlgorithm(ImageArray<?> queryImage, int queryThreshold,
                                                                                      int queryBorderSize, ColorDepthSearchParams cdsParams) {
        double fluct = cdsParams.getDoubleParam("pixColorFluctuation", defaults.getDoubleParam("pixColorFluctuation", 2.0));
        int xyShift = cdsParams.getIntParam("xyShift", defaults.getIntParam("xyShift", 0));
        if ((xyShift & 0x1) == 1) throw new IllegalArgumentException("XY shift parameter must be an even number.");   // :57-60
        return new Algorithm(queryImage, queryThreshold,
                cdsParams.getBoolParam("mirrorMask", defaults.getBoolParam("mirrorMask", false)),
                cdsParams.getIntParam("dataThreshold", defaults.getIntParam("dataThreshold", 100)),
                fluct / 100, xyShift, rectanglesOf(ignoredRegionsProvider, queryImage));
    }

    /**
     * The reference describes excluded regions as a predicate over (x, y); the device takes rectangles.  Every region
     * the tools define is a union of rectangles (AbstractColorDepthMatchArgs.java:101-119); recover them by scanning the
     * predicate once per mask size and merging equal row runs.
     */
    public static int[][] rectanglesOf(ImageRegionDefinition def, ImageArray<?> img) {
        if (def == null) return new int[0][];
        BiPredicate<Integer, Integer> in = def.getRegion(img);
        List<int[]> open = new ArrayList<>(), done = new ArrayList<>();
        for (int y = 0; y <= img.getHeight(); y++) {
            List<int[]> runs = new ArrayList<>();
            if (y < img.getHeight())
                for (int x = 0; x < img.getWidth(); ) {
                    if (!in.test(x, y)) { x++; continue; }
                    int x0 = x;
                    while (x < img.getWidth() && in.test(x, y)) x++;
                    runs.add(new int[]{x0, y, x, y + 1});
                }
            List<int[]> next = new ArrayList<>();
            for (int[] r : runs) {
                int[] cont = null;
                for (int[] o : open) if (o[0] == r[0] && o[2] == r[2]) cont = o;
                if (cont != null) { open.remove(cont); cont[3] = y + 1; next.add(cont); } else next.add(r);
            }
            done.addAll(open);
            open = next;
        }
        if (done.size() > CdsGpu.CDS_MAX_RECTS) throw new IllegalArgumentException("excluded regions are not a union of <= 8 rectangles");
        return done.toArray(new int[0][]);
    }

    static final class Algorithm implements ColorDepthSearchAlgorithm<PixelMatchScore> {
        private final ImageArray<?> queryImage;
        private final transient MemorySegment maskSet;   // native handle: not serialisable, Spark mode is unsupported on this provider
        private final int querySize;

        Algorithm(ImageArray<?> queryImage, int queryThreshold, boolean mirror, int dataThreshold, double zTolerance, int xyShift, int[][] rects) {
            this.queryImage = queryImage;
            try (Arena a = Arena.ofConfined()) {
                MemorySegment out = a.allocate(ValueLayout.ADDRESS), size = a.allocate(ValueLayout.JAVA_INT);
                MemorySegment p = CdsGpu.pixParams(a, queryThreshold, dataThreshold, zTolerance, xyShift, mirror, rects);
                CdsGpu.check((int) CdsGpu.masksetCreate.invokeExact(CdsGpu.context(), queryImage.getWidth(), queryImage.getHeight(), p, out));
                maskSet = out.get(ValueLayout.ADDRESS, 0);
                CdsGpu.check((int) CdsGpu.masksetAddRgb.invokeExact(maskSet, CdsGpu.copyBytes(a, ImageArrayAccess.rgbBytes(queryImage)), 1, size));
                querySize = size.get(ValueLayout.JAVA_INT, 0);
            } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new IllegalStateException(t); }
        }

        @Override public ImageArray<?> getQueryImage() { return queryImage; }
        @Override public int getQuerySize() { return querySize; }
        @Override public int getQueryFirstPixelIndex() { return 0; }   // unused by the tools (SURVEY a2)
        @Override public int getQueryLastPixelIndex() { return queryImage.getPixelCount() - 1; }
        @Override public Set<ComputeFileType> getRequiredTargetVariantTypes() { return Collections.emptySet(); }

        /** Thread-safe: calls on one context are serialised inside the library. */
        @Override
        public PixelMatchScore calculateMatchingScore(@Nonnull ImageArray<?> target, Map<ComputeFileType, Supplier<ImageArray<?>>> variants) {
            try (Arena a = Arena.ofConfined()) {
                MemorySegment score = a.allocate(ValueLayout.JAVA_INT), ratio = a.allocate(ValueLayout.JAVA_DOUBLE), mir = a.allocate(ValueLayout.JAVA_INT);
                CdsGpu.check((int) CdsGpu.scorePairRgb.invokeExact(CdsGpu.context(), maskSet, 0, CdsGpu.copyBytes(a, ImageArrayAccess.rgbBytes(target)),
                        target.getWidth(), target.getHeight(), score, ratio, mir));
                return new PixelMatchScore(score.get(ValueLayout.JAVA_INT, 0), ratio.get(ValueLayout.JAVA_DOUBLE, 0), mir.get(ValueLayout.JAVA_INT, 0) != 0);
            } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new IllegalStateException(t); }
        }

        public void close() {
            try { CdsGpu.masksetDestroy.invokeExact(maskSet); } catch (Throwable ignored) { }
        }
    }
}
