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
 * This class serves the literal single-pair API: every calculateMatchingScore is one blocking cds_pairq_score call.  The calls of
 * the processor's ~40 pool threads (LocalColorMIPSearchProcessor.java:93-105) meet in the library's micro-batching queue and are
 * scored by one kernel launch per batch; a target that was scored before is found in the device-side cache by its identity (the
 * reference serves the same ImageArray object to every mask out of its Guava cache, CachedMIPsUtils.java:60-110), so it crosses
 * PCIe once.  All algorithms of one parameter set share ONE native mask set and ONE queue (masks are appended as algorithms are
 * created); raise --task-concurrency to a few hundred threads to fill the batches.  Bulk throughput still comes from
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

    /** One native mask set + pair queue per (image size, parameter set). */
    private static final class Shared {
        MemorySegment maskSet, queue;
        int nMasks;
    }
    private final Map<String, Shared> shared = new java.util.HashMap<>();

    /** A unique id per target ImageArray OBJECT (System.identityHashCode can collide; equal keys must mean equal pixels). */
    private static final Map<ImageArray<?>, Long> TARGET_IDS = Collections.synchronizedMap(new java.util.WeakHashMap<>());
    private static final java.util.concurrent.atomic.AtomicLong NEXT_ID = new java.util.concurrent.atomic.AtomicLong();
    static long targetKey(ImageArray<?> target) { return TARGET_IDS.computeIfAbsent(target, t -> NEXT_ID.incrementAndGet()); }

    private synchronized int addMask(String key, ImageArray<?> queryImage, int queryThreshold, boolean mirror, int dataThreshold, double zTolerance,
                                     int xyShift, int[][] rects, Shared[] out, int[] sizeOut) {
        Shared sh = shared.get(key);
        try (Arena a = Arena.ofConfined()) {
            MemorySegment o = a.allocate(ValueLayout.ADDRESS), size = a.allocate(ValueLayout.JAVA_INT);
            if (sh == null) {
                sh = new Shared();
                MemorySegment p = CdsGpu.pixParams(a, queryThreshold, dataThreshold, zTolerance, xyShift, mirror, rects);
                CdsGpu.check((int) CdsGpu.masksetCreate.invokeExact(CdsGpu.context(), queryImage.getWidth(), queryImage.getHeight(), p, o));
                sh.maskSet = o.get(ValueLayout.ADDRESS, 0);
                shared.put(key, sh);
            }
            CdsGpu.check((int) CdsGpu.masksetAddRgb.invokeExact(sh.maskSet, CdsGpu.copyBytes(a, ImageArrayAccess.rgbBytes(queryImage)), 1, size));
            sizeOut[0] = size.get(ValueLayout.JAVA_INT, 0);
            if (sh.queue == null) {
                // max batch 256, wait at most 60 us for company, 4096 cached targets per device (11 GB of code planes for 1210 x 566)
                CdsGpu.check((int) CdsGpu.pairqCreate.invokeExact(CdsGpu.context(), sh.maskSet, 256, 60, Integer.getInteger("cdsgpu.cachedTargets", 4096), o));
                sh.queue = o.get(ValueLayout.ADDRESS, 0);
            }
            out[0] = sh;
            return sh.nMasks++;
        } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new IllegalStateException(t); }
    }

    @Override
    public ColorDepthSearchAlgorithm<PixelMatchScore> createColorDepthSearchAlgorithm(ImageArray<?> queryImage, int queryThreshold,
                                                                                      int queryBorderSize, ColorDepthSearchParams cdsParams) {
        double fluct = cdsParams.getDoubleParam("pixColorFluctuation", defaults.getDoubleParam("pixColorFluctuation", 2.0));
        int xyShift = cdsParams.getIntParam("xyShift", defaults.getIntParam("xyShift", 0));
        if ((xyShift & 0x1) == 1) throw new IllegalArgumentException("XY shift parameter must be an even number.");   // :57-60
        boolean mirror = cdsParams.getBoolParam("mirrorMask", defaults.getBoolParam("mirrorMask", false));
        int dataThreshold = cdsParams.getIntParam("dataThreshold", defaults.getIntParam("dataThreshold", 100));
        int[][] rects = rectanglesOf(ignoredRegionsProvider, queryImage);
        String key = queryImage.getWidth() + "x" + queryImage.getHeight() + ":" + queryThreshold + ":" + mirror + ":" + dataThreshold + ":" + fluct + ":" + xyShift
                + ":" + java.util.Arrays.deepToString(rects);
        Shared[] sh = new Shared[1];
        int[] size = new int[1];
        int index = addMask(key, queryImage, queryThreshold, mirror, dataThreshold, fluct / 100, xyShift, rects, sh, size);
        return new Algorithm(queryImage, sh[0].queue, index, size[0]);
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
        private final transient MemorySegment queue;   // native handle: not serialisable, Spark mode is unsupported on this provider
        private final int maskIndex, querySize;

        Algorithm(ImageArray<?> queryImage, MemorySegment queue, int maskIndex, int querySize) {
            this.queryImage = queryImage; this.queue = queue; this.maskIndex = maskIndex; this.querySize = querySize;
        }

        @Override public ImageArray<?> getQueryImage() { return queryImage; }
        @Override public int getQuerySize() { return querySize; }
        @Override public int getQueryFirstPixelIndex() { return 0; }   // unused by the tools (SURVEY a2)
        @Override public int getQueryLastPixelIndex() { return queryImage.getPixelCount() - 1; }
        @Override public Set<ComputeFileType> getRequiredTargetVariantTypes() { return Collections.emptySet(); }

        /** Thread-safe and re-entrant: cds_pairq_score takes no context-wide lock; calls that arrive together share a kernel launch. */
        @Override
        public PixelMatchScore calculateMatchingScore(@Nonnull ImageArray<?> target, Map<ComputeFileType, Supplier<ImageArray<?>>> variants) {
            try (Arena a = Arena.ofConfined()) {
                MemorySegment score = a.allocate(ValueLayout.JAVA_INT), ratio = a.allocate(ValueLayout.JAVA_DOUBLE), mir = a.allocate(ValueLayout.JAVA_INT);
                // the pixels are only read when the key is not in the device cache; a heap array has to be copied off-heap to be passed at all
                // (FFM cannot pin byte[]), a JNI shim would use GetPrimitiveArrayCritical instead
                CdsGpu.check((int) CdsGpu.pairqScore.invokeExact(queue, maskIndex, targetKey(target), CdsGpu.copyBytes(a, ImageArrayAccess.rgbBytes(target)),
                        target.getWidth(), target.getHeight(), score, ratio, mir));
                return new PixelMatchScore(score.get(ValueLayout.JAVA_INT, 0), ratio.get(ValueLayout.JAVA_DOUBLE, 0), mir.get(ValueLayout.JAVA_INT, 0) != 0);
            } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new IllegalStateException(t); }
        }
    }
}
