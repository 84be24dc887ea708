package org.janelia.colormipsearch.cmd.cdsprocess;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.ValueLayout;
import java.util.ArrayList;
import java.util.List;
import java.util.Set;

import org.janelia.colormipsearch.cds.ColorMIPSearch;
import org.janelia.colormipsearch.cds.gpu.CdsGpu;
import org.janelia.colormipsearch.cmd.CachedMIPsUtils;
import org.janelia.colormipsearch.imageprocessing.ImageArrayAccess;
import org.janelia.colormipsearch.mips.NeuronMIP;
import org.janelia.colormipsearch.mips.NeuronMIPUtils;
import org.janelia.colormipsearch.model.AbstractNeuronEntity;
import org.janelia.colormipsearch.model.CDMatchEntity;
import org.janelia.colormipsearch.model.ComputeFileType;
import org.janelia.colormipsearch.model.ProcessingType;

/**
 * The batched seam: ColorMIPSearchProcessor.findAllColorDepthMatches (ColorMIPSearchProcessor.java:8-12) as ONE streaming
 * native search -- the same seam --use-spark already uses (ColorDepthSearchCmd.java:279-295).  Masks are prepared once on the
 * device; targets are decoded by the JVM (as today, through CachedMIPsUtils) into a pinned staging buffer and streamed to the
 * GPUs; only matches that pass ColorMIPSearch.isMatch come back (LocalColorMIPSearchProcessor.java:93-105 keeps exactly those).
 * With -Dcdsgpu.ingest=tiff and targets that are plain .tif files the JVM does not decode at all: the files' bytes go into the
 * pinned buffer as stored and cds_search_stream_*_tiff decodes them on the device (PackBits or uncompressed RGB; anything else
 * makes the native call return CDS_ERR_UNSUPPORTED and the processor falls back to the decoded path).
 * UNVERIFIED (no JDK in the build image of the GPU library).
 */
public class GpuColorMIPSearchProcessor<M extends AbstractNeuronEntity, T extends AbstractNeuronEntity> implements ColorMIPSearchProcessor<M, T> {
    private final Number cdsRunId;
    private final double pctPositivePixels;
    private final int maskThreshold, dataThreshold, xyShift, maxMatchesPerMask;
    private final double pixColorFluctuation;
    private final boolean mirrorMask;
    private final int[][] labelRects;
    private final Set<String> tags;

    public GpuColorMIPSearchProcessor(Number cdsRunId, double pctPositivePixels, int maskThreshold, int dataThreshold, double pixColorFluctuation,
                                      int xyShift, boolean mirrorMask, int[][] labelRects, int maxMatchesPerMask, Set<String> tags) {
        if ((xyShift & 0x1) == 1) throw new IllegalArgumentException("XY shift parameter must be an even number.");
        this.cdsRunId = cdsRunId; this.pctPositivePixels = pctPositivePixels; this.maskThreshold = maskThreshold; this.dataThreshold = dataThreshold;
        this.pixColorFluctuation = pixColorFluctuation; this.xyShift = xyShift; this.mirrorMask = mirrorMask; this.labelRects = labelRects;
        this.maxMatchesPerMask = maxMatchesPerMask; this.tags = tags;
    }

    @Override
    @SuppressWarnings("unchecked")
    public List<CDMatchEntity<M, T>> findAllColorDepthMatches(List<M> queryMIPs, List<T> targetMIPs) {
        List<CDMatchEntity<M, T>> results = new ArrayList<>();
        List<NeuronMIP<M>> masks = new ArrayList<>();
        for (M q : queryMIPs) {
            NeuronMIP<M> m = NeuronMIPUtils.loadComputeFile(q, ComputeFileType.InputColorDepthImage);
            if (m != null && !m.hasNoImageArray()) masks.add(m);
        }
        List<NeuronMIP<T>> targets = new ArrayList<>();
        for (T t : targetMIPs) {
            NeuronMIP<T> tm = CachedMIPsUtils.loadMIP(t, ComputeFileType.InputColorDepthImage);
            if (NeuronMIPUtils.hasImageArray(tm)) targets.add(tm);
        }
        if (masks.isEmpty() || targets.isEmpty()) return results;
        int w = masks.get(0).getImageArray().getWidth(), h = masks.get(0).getImageArray().getHeight();
        long imgBytes = 3L * w * h;
        int k = Math.min(maxMatchesPerMask > 0 ? maxMatchesPerMask : targets.size(), targets.size());
        try (Arena a = Arena.ofConfined()) {
            MemorySegment ctx = CdsGpu.context();
            MemorySegment out = a.allocate(ValueLayout.ADDRESS);
            CdsGpu.check((int) CdsGpu.masksetCreate.invokeExact(ctx, w, h,
                    CdsGpu.pixParams(a, maskThreshold, dataThreshold, pixColorFluctuation / 100, xyShift, mirrorMask, labelRects), out));
            MemorySegment ms = out.get(ValueLayout.ADDRESS, 0);
            int[] maskSizes = new int[masks.size()];
            MemorySegment size = a.allocate(ValueLayout.JAVA_INT);
            for (int i = 0; i < masks.size(); i++) {
                CdsGpu.check((int) CdsGpu.masksetAddRgb.invokeExact(ms, CdsGpu.copyBytes(a, ImageArrayAccess.rgbBytes(masks.get(i).getImageArray())), 1, size));
                maskSizes[i] = size.get(ValueLayout.JAVA_INT, 0);
            }
            if ("tiff".equals(System.getProperty("cdsgpu.ingest")) && searchTiffFiles(a, ctx, ms, masks, targetMIPs, maskSizes, k, results)) {
                CdsGpu.masksetDestroy.invokeExact(ms);
                return results;
            }
            // targets: one pinned buffer, filled by the JVM's decoders, streamed by the library (H2D of chunk i+1 overlaps the search of chunk i)
            CdsGpu.check((int) CdsGpu.hostAlloc.invokeExact(ctx, imgBytes * targets.size(), out));
            MemorySegment pinned = out.get(ValueLayout.ADDRESS, 0).reinterpret(imgBytes * targets.size());
            for (int i = 0; i < targets.size(); i++)
                MemorySegment.copy(ImageArrayAccess.rgbBytes(targets.get(i).getImageArray()), 0, pinned, ValueLayout.JAVA_BYTE, imgBytes * i, (int) imgBytes);
            if (maxMatchesPerMask <= 0) {
                // like LocalColorMIPSearchProcessor: every pair that passes ColorMIPSearch.isMatch
                long cap = Math.max(1024L, 4L * masks.size());
                MemorySegment count = a.allocate(ValueLayout.JAVA_LONG);
                for (int attempt = 0; ; attempt++) {
                    MemorySegment mk = a.allocate(4 * cap, 4), tg = a.allocate(8 * cap, 8), sc = a.allocate(4 * cap, 4), mir = a.allocate(cap);
                    int st = (int) CdsGpu.searchStreamMatches.invokeExact(ctx, ms, pinned, (long) targets.size(), pctPositivePixels, cap, mk, tg, sc, mir, count);
                    long n = count.get(ValueLayout.JAVA_LONG, 0);
                    if (st == CdsGpu.CDS_ERR_CAPACITY && attempt == 0) { cap = n; continue; }
                    CdsGpu.check(st);
                    for (long i = 0; i < n; i++)
                        results.add(newMatch(masks, targets, maskSizes, mk.get(ValueLayout.JAVA_INT, 4 * i), (int) tg.get(ValueLayout.JAVA_LONG, 8 * i),
                                sc.get(ValueLayout.JAVA_INT, 4 * i), mir.get(ValueLayout.JAVA_BYTE, i) != 0));
                    break;
                }
            } else {
                MemorySegment score = a.allocate(4L * masks.size() * k, 4), target = a.allocate(8L * masks.size() * k, 8);
                MemorySegment mirrored = a.allocate((long) masks.size() * k), count = a.allocate(4L * masks.size(), 4);
                CdsGpu.check((int) CdsGpu.searchStream.invokeExact(ctx, ms, pinned, (long) targets.size(), k, pctPositivePixels, score, target, mirrored, count));
                for (int m = 0; m < masks.size(); m++) {
                    int n = count.get(ValueLayout.JAVA_INT, 4L * m);
                    for (int i = 0; i < n; i++) {
                        long o = (long) m * k + i;
                        results.add(newMatch(masks, targets, maskSizes, m, (int) target.get(ValueLayout.JAVA_LONG, 8 * o),
                                score.get(ValueLayout.JAVA_INT, 4 * o), mirrored.get(ValueLayout.JAVA_BYTE, o) != 0));
                    }
                }
            }
            CdsGpu.check((int) CdsGpu.hostFree.invokeExact(ctx, pinned));
            CdsGpu.masksetDestroy.invokeExact(ms);
        } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new IllegalStateException(t); }
        return results;
    }

    /**
     * Targets as TIFF files: raw bytes -> pinned blob -> cds_search_stream_matches_tiff / cds_search_stream_tiff.  Returns false
     * (nothing added to `results`) when a target is not a plain .tif file or the library cannot decode one of them.
     */
    private boolean searchTiffFiles(Arena a, MemorySegment ctx, MemorySegment ms, List<NeuronMIP<M>> masks, List<T> targetMIPs, int[] maskSizes,
                                    int k, List<CDMatchEntity<M, T>> results) throws Throwable {
        List<T> kept = new ArrayList<>();
        List<java.nio.file.Path> paths = new ArrayList<>();
        long total = 0;
        for (T t : targetMIPs) {
            org.janelia.colormipsearch.model.FileData fd = t.getComputeFileData(ComputeFileType.InputColorDepthImage);
            if (fd == null) continue;
            String name = fd.getFileName() == null ? "" : fd.getFileName().toLowerCase();
            if (fd.getDataType() != org.janelia.colormipsearch.model.FileData.FileDataType.file || !(name.endsWith(".tif") || name.endsWith(".tiff"))) return false;
            java.nio.file.Path p = java.nio.file.Paths.get(fd.getFileName());
            kept.add(t); paths.add(p);
            total += java.nio.file.Files.size(p);
        }
        if (kept.isEmpty()) return true;
        MemorySegment out = a.allocate(ValueLayout.ADDRESS);
        CdsGpu.check((int) CdsGpu.hostAlloc.invokeExact(ctx, total + 64, out));
        MemorySegment blob = out.get(ValueLayout.ADDRESS, 0).reinterpret(total + 64);
        MemorySegment offsets = a.allocate(8L * (kept.size() + 1), 8);
        long at = 0;
        for (int i = 0; i < kept.size(); i++) {
            offsets.set(ValueLayout.JAVA_LONG, 8L * i, at);
            try (java.nio.channels.FileChannel ch = java.nio.channels.FileChannel.open(paths.get(i))) {
                java.nio.ByteBuffer bb = blob.asSlice(at, ch.size()).asByteBuffer();
                while (bb.hasRemaining() && ch.read(bb) >= 0) { }
                at += ch.size();
            }
        }
        offsets.set(ValueLayout.JAVA_LONG, 8L * kept.size(), at);
        List<NeuronMIP<T>> targets = new ArrayList<>();
        for (T t : kept) targets.add(new NeuronMIP<>(t, t.getComputeFileData(ComputeFileType.InputColorDepthImage), null));
        boolean done = true;
        if (maxMatchesPerMask <= 0) {
            long cap = Math.max(1024L, 4L * masks.size());
            MemorySegment count = a.allocate(ValueLayout.JAVA_LONG);
            for (int attempt = 0; ; attempt++) {
                MemorySegment mk = a.allocate(4 * cap, 4), tg = a.allocate(8 * cap, 8), sc = a.allocate(4 * cap, 4), mir = a.allocate(cap);
                int st = (int) CdsGpu.searchStreamMatchesTiff.invokeExact(ctx, ms, blob, offsets, (long) kept.size(), pctPositivePixels, cap, mk, tg, sc, mir, count);
                long n = count.get(ValueLayout.JAVA_LONG, 0);
                if (st == CdsGpu.CDS_ERR_CAPACITY && attempt == 0) { cap = n; continue; }
                if (st == CdsGpu.CDS_ERR_UNSUPPORTED) { done = false; break; }
                CdsGpu.check(st);
                for (long i = 0; i < n; i++)
                    results.add(newMatch(masks, targets, maskSizes, mk.get(ValueLayout.JAVA_INT, 4 * i), (int) tg.get(ValueLayout.JAVA_LONG, 8 * i),
                            sc.get(ValueLayout.JAVA_INT, 4 * i), mir.get(ValueLayout.JAVA_BYTE, i) != 0));
                break;
            }
        } else {
            int kk = Math.min(k, kept.size());
            MemorySegment score = a.allocate(4L * masks.size() * kk, 4), target = a.allocate(8L * masks.size() * kk, 8);
            MemorySegment mirrored = a.allocate((long) masks.size() * kk), count = a.allocate(4L * masks.size(), 4);
            int st = (int) CdsGpu.searchStreamTiff.invokeExact(ctx, ms, blob, offsets, (long) kept.size(), kk, pctPositivePixels, score, target, mirrored, count);
            if (st == CdsGpu.CDS_ERR_UNSUPPORTED) done = false;
            else {
                CdsGpu.check(st);
                for (int m = 0; m < masks.size(); m++) {
                    int n = count.get(ValueLayout.JAVA_INT, 4L * m);
                    for (int i = 0; i < n; i++) {
                        long o = (long) m * kk + i;
                        results.add(newMatch(masks, targets, maskSizes, m, (int) target.get(ValueLayout.JAVA_LONG, 8 * o),
                                score.get(ValueLayout.JAVA_INT, 4 * o), mirrored.get(ValueLayout.JAVA_BYTE, o) != 0));
                    }
                }
            }
        }
        CdsGpu.check((int) CdsGpu.hostFree.invokeExact(ctx, blob));
        return done;
    }

    @SuppressWarnings("unchecked")
    private CDMatchEntity<M, T> newMatch(List<NeuronMIP<M>> masks, List<NeuronMIP<T>> targets, int[] maskSizes, int m, int t, int pix, boolean mirrored) {
        CDMatchEntity<M, T> r = new CDMatchEntity<>();
        r.setMaskImage((M) masks.get(m).getNeuronInfo().addProcessedTags(ProcessingType.ColorDepthSearch, tags));
        r.setMatchedImage((T) targets.get(t).getNeuronInfo().addProcessedTags(ProcessingType.ColorDepthSearch, tags));
        r.setSessionRefId(cdsRunId);
        r.setMatchFound(true);                                              // only isMatch pairs are returned
        r.setMatchingPixels(pix);
        r.setMatchingPixelsRatio((float) ((double) pix / maskSizes[m]));     // PixelMatchScore.getNormalizedScore
        r.setMirrored(mirrored);
        r.addAllTags(tags);
        return r;
    }

    @Override
    public void terminate() { }
}
