package org.janelia.colormipsearch.cds;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.ValueLayout;
import java.util.EnumSet;
import java.util.Map;
import java.util.Set;
import java.util.function.Supplier;

import javax.annotation.Nonnull;

import org.janelia.colormipsearch.cds.gpu.CdsGpu;
import org.janelia.colormipsearch.imageprocessing.ImageArray;
import org.janelia.colormipsearch.imageprocessing.ImageArrayAccess;
import org.janelia.colormipsearch.imageprocessing.ImageRegionDefinition;
import org.janelia.colormipsearch.model.ComputeFileType;

/**
 * Drop-in for createShapeMatchCDSAlgorithmProvider (ColorDepthSearchAlgorithmProviderFactory.java:76-127).  Lives in package
 * org.janelia.colormipsearch.cds because ShapeMatchScore's constructors are package-private (ShapeMatchScore.java:12,20).
 * UNVERIFIED (no JDK in the build image of the GPU library).
 */
public class GpuShape2DMatchColorDepthSearchAlgorithmProvider implements ColorDepthSearchAlgorithmProvider<ShapeMatchScore> {
    private final ColorDepthSearchParams defaults = new ColorDepthSearchParams();
    private final ImageArray<?> roiMask;
    private final ImageRegionDefinition excludedRegions;

    public GpuShape2DMatchColorDepthSearchAlgorithmProvider(boolean mirrorMask, ImageArray<?> roiMaskImageArray, ImageRegionDefinition excludedRegions) {
        defaults.setParam("mirrorMask", mirrorMask);
        this.roiMask = roiMaskImageArray;
        this.excludedRegions = excludedRegions;
    }

    @Override public ColorDepthSearchParams getDefaultCDSParams() { return defaults; }

    @Override
    public ColorDepthSearchAlgorithm<ShapeMatchScore> createColorDepthSearchAlgorithm(ImageArray<?> queryImage, int queryThreshold,
                                                                                      int queryBorderSize, ColorDepthSearchParams cdsParams) {
        boolean mirror = cdsParams.getBoolParam("mirrorMask", defaults.getBoolParam("mirrorMask", false));
        int[][] rects = GpuPixelMatchColorDepthSearchAlgorithmProvider.rectanglesOf(excludedRegions, queryImage);
        return new Algorithm(queryImage, queryThreshold, queryBorderSize, mirror, roiMask, rects);
    }

    static final class Algorithm implements ColorDepthSearchAlgorithm<ShapeMatchScore> {
        private static final Set<ComputeFileType> REQUIRED = EnumSet.of(ComputeFileType.GradientImage, ComputeFileType.ZGapImage);
        private final ImageArray<?> queryImage;
        private final transient MemorySegment shapeMaskSet;
        private final int querySize;

        Algorithm(ImageArray<?> queryImage, int queryThreshold, int border, boolean mirror, ImageArray<?> roi, int[][] rects) {
            this.queryImage = queryImage;
            try (Arena a = Arena.ofConfined()) {
                MemorySegment out = a.allocate(ValueLayout.ADDRESS), qm = a.allocate(ValueLayout.JAVA_LONG), he = a.allocate(ValueLayout.JAVA_LONG);
                MemorySegment r = a.allocate(16L * Math.max(rects.length, 1));
                for (int i = 0; i < rects.length; i++) for (int k = 0; k < 4; k++) r.set(ValueLayout.JAVA_INT, 16L * i + 4L * k, rects[i][k]);
                MemorySegment roiSeg = roi == null ? MemorySegment.NULL : CdsGpu.copyBytes(a, ImageArrayAccess.rgbBytes(roi));
                CdsGpu.check((int) CdsGpu.shapeMasksetCreate.invokeExact(CdsGpu.context(), queryImage.getWidth(), queryImage.getHeight(), queryThreshold, border,
                        mirror ? 1 : 0, r, rects.length, roiSeg, out));
                shapeMaskSet = out.get(ValueLayout.ADDRESS, 0);
                CdsGpu.check((int) CdsGpu.shapeMasksetAddRgb.invokeExact(shapeMaskSet, CdsGpu.copyBytes(a, ImageArrayAccess.rgbBytes(queryImage)), 1, qm, he));
                querySize = (int) qm.get(ValueLayout.JAVA_LONG, 0);
            } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new IllegalStateException(t); }
        }

        @Override public ImageArray<?> getQueryImage() { return queryImage; }
        @Override public int getQuerySize() { return querySize; }
        @Override public int getQueryFirstPixelIndex() { return 0; }
        @Override public int getQueryLastPixelIndex() { return queryImage.getPixelCount() - 1; }
        @Override public Set<ComputeFileType> getRequiredTargetVariantTypes() { return REQUIRED; }

        @Override
        public ShapeMatchScore calculateMatchingScore(@Nonnull ImageArray<?> target, Map<ComputeFileType, Supplier<ImageArray<?>>> variants) {
            ImageArray<?> grad = fetch(variants, ComputeFileType.GradientImage), zgap = fetch(variants, ComputeFileType.ZGapImage);
            if (grad == null || zgap == null) return new ShapeMatchScore(-1, -1, -1, false);   // Shape2DMatch...:155-158
            try (Arena a = Arena.ofConfined()) {
                int n = target.getPixelCount();
                MemorySegment g16 = a.allocate(2L * n, 2);
                if (ImageArrayAccess.isGray16(grad)) MemorySegment.copy(ImageArrayAccess.gray16(grad), 0, g16, ValueLayout.JAVA_SHORT, 0, n);
                else { byte[] g8 = ImageArrayAccess.gray8(grad); for (int i = 0; i < n; i++) g16.set(ValueLayout.JAVA_SHORT, 2L * i, (short) (g8[i] & 0xFF)); }
                MemorySegment pm = a.allocate(ValueLayout.JAVA_INT), pt = a.allocate(ValueLayout.JAVA_LONG);
                MemorySegment gap = a.allocate(ValueLayout.JAVA_LONG), he = a.allocate(ValueLayout.JAVA_LONG), mir = a.allocate(ValueLayout.JAVA_BYTE);
                CdsGpu.check((int) CdsGpu.shapeScorePairs.invokeExact(CdsGpu.context(), shapeMaskSet,
                        CdsGpu.copyBytes(a, ImageArrayAccess.rgbBytes(target)), g16, CdsGpu.copyBytes(a, ImageArrayAccess.rgbBytes(zgap)),
                        MemorySegment.NULL, 1L, pm, pt, 1L, gap, he, mir));
                return new ShapeMatchScore(gap.get(ValueLayout.JAVA_LONG, 0), he.get(ValueLayout.JAVA_LONG, 0), -1, mir.get(ValueLayout.JAVA_BYTE, 0) != 0);
            } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new IllegalStateException(t); }
        }

        private static ImageArray<?> fetch(Map<ComputeFileType, Supplier<ImageArray<?>>> v, ComputeFileType t) {
            Supplier<ImageArray<?>> s = v == null ? null : v.get(t);
            return s == null ? null : s.get();
        }
    }
}
