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
 * GPUs in chunks of cdsgpu.targetChunk images (one pinned staging buffer, whatever the size of the library; masks are grouped by
 * image size); only matches that pass ColorMIPSearch.isMatch come back (LocalColorMIPSearchProcessor.java:93-105 keeps exactly those).
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

    /** Targets per native call on the decoded path: bounds the pinned staging buffer (2 GB for 1210 x 566 images) whatever the library's size. */
    private static final int TARGET_CHUNK = Integer.getInteger("cdsgpu.targetChunk", 1024);

    @Override
    public List<CDMatchEntity<M, T>> findAllColorDepthMatches(List<M> queryMIPs, List<T> targetMIPs) {
        List<CDMatchEntity<M, T>> results = new ArrayList<>();
        // masks grouped by image size: a mask set holds images of one size, and the reference compares a mask only with targets of
        // its own size (PixelMatchColorDepthSearchAlgorithm.java:171-175 throws otherwise)
        java.util.Map<Long, List<NeuronMIP<M>>> bySize = new java.util.LinkedHashMap<>();
        for (M q : queryMIPs) {
            NeuronMIP<M> m = NeuronMIPUtils.loadComputeFile(q, ComputeFileType.InputColorDepthImage);
            if (m == null || m.hasNoImageArray()) continue;
            long key = ((long) m.getImageArray().getWidth() << 32) | m.getImageArray().getHeight();
            bySize.computeIfAbsent(key, x -> new ArrayList<>()).add(m);
        }
        for (List<NeuronMIP<M>> masks : bySize.values()) searchOneSize(masks, targetMIPs, results);
        return results;
    }

    private void searchOneSize(List<NeuronMIP<M>> masks, List<T> targetMIPs, List<CDMatchEntity<M, T>> results) {
        int w = masks.get(0).getImageArray().getWidth(), h = masks.get(0).getImageArray().getHeight();
        long imgBytes = 3L * w * h;
        try (Arena a = Arena.ofConfined()) {
            MemorySegment ctx = CdsGpu.context();
            MemorySegment out = a.allocate(ValueLayout.ADDRESS);
            CdsGpu.check((int) CdsGpu.masksetCreate.invokeExact(ctx, w, h,
                    CdsGpu.pixParams(a, maskThreshold, dataThreshold, pixColorFluctuation / 100, xyShift, mirrorMask, labelRects), out));
            MemorySegment ms = out.get(ValueLayout.ADDRESS, 0);
            int[] maskSizes = new int[masks.size()];
            MemorySegment size = a.allocate(ValueLayout.JAVA_INT);
            for (int i = 0; i < masks.size(); i++) {
                try (Arena one = Arena.ofConfined()) {
                    CdsGpu.check((int) CdsGpu.masksetAddRgb.invokeExact(ms, CdsGpu.copyBytes(one, ImageArrayAccess.rgbBytes(masks.get(i).getImageArray())), 1, size));
                }
                maskSizes[i] = size.get(ValueLayout.JAVA_INT, 0);
            }
            if ("tiff".equals(System.getProperty("cdsgpu.ingest")) && searchTiffFiles(a, ctx, ms, masks, targetMIPs, maskSizes,
                    maxMatchesPerMask > 0 ? maxMatchesPerMask : targetMIPs.size(), results)) {
                CdsGpu.masksetDestroy.invokeExact(ms);
                return;
            }
            // Decoded targets: ONE pinned staging buffer of TARGET_CHUNK images, refilled by the JVM's decoders chunk after chunk; inside a
            // chunk the library overlaps the upload with the search.  Targets of another size are skipped for this mask group.
            CdsGpu.check((int) CdsGpu.hostAlloc.invokeExact(ctx, imgBytes * TARGET_CHUNK, out));
            MemorySegment pinned = out.get(ValueLayout.ADDRESS, 0).reinterpret(imgBytes * TARGET_CHUNK);
            List<List<CDMatchEntity<M, T>>> perMask = new ArrayList<>();
            for (int m = 0; m < masks.size(); m++) perMask.add(new ArrayList<>());
            List<NeuronMIP<T>> chunk = new ArrayList<>();
            for (int t0 = 0; t0 <= targetMIPs.size(); t0++) {
                if (t0 < targetMIPs.size()) {
                    NeuronMIP<T> tm = CachedMIPsUtils.loadMIP(targetMIPs.get(t0), ComputeFileType.InputColorDepthImage);
                    if (NeuronMIPUtils.hasImageArray(tm) && tm.getImageArray().getWidth() == w && tm.getImageArray().getHeight() == h) {
                        MemorySegment.copy(ImageArrayAccess.rgbBytes(tm.getImageArray()), 0, pinned, ValueLayout.JAVA_BYTE, imgBytes * chunk.size(), (int) imgBytes);
                        chunk.add(tm);
                    }
                    if (chunk.size() < TARGET_CHUNK) continue;
                }
                if (chunk.isEmpty()) continue;
                // every pair of the chunk that passes ColorMIPSearch.isMatch, like LocalColorMIPSearchProcessor.java:93-105
                long cap = Math.max(1024L, 4L * masks.size());
                try (Arena ca = Arena.ofConfined()) {
                    MemorySegment count = ca.allocate(ValueLayout.JAVA_LONG);
                    for (int attempt = 0; ; attempt++) {
                        MemorySegment mk = ca.allocate(4 * cap, 4), tg = ca.allocate(8 * cap, 8), sc = ca.allocate(4 * cap, 4), mir = ca.allocate(cap);
                        int st = (int) CdsGpu.searchStreamMatches.invokeExact(ctx, ms, pinned, (long) chunk.size(), pctPositivePixels, cap, mk, tg, sc, mir, count);
                        long n = count.get(ValueLayout.JAVA_LONG, 0);
                        if (st == CdsGpu.CDS_ERR_CAPACITY && attempt == 0) { cap = n; continue; }
                        CdsGpu.check(st);
                        for (long i = 0; i < n; i++) {
                            int m = mk.get(ValueLayout.JAVA_INT, 4 * i);
                            perMask.get(m).add(newMatch(masks, chunk, maskSizes, m, (int) tg.get(ValueLayout.JAVA_LONG, 8 * i),
                                    sc.get(ValueLayout.JAVA_INT, 4 * i), mir.get(ValueLayout.JAVA_BYTE, i) != 0));
                        }
                        break;
                    }
                }
                chunk = new ArrayList<>();
            }
            for (List<CDMatchEntity<M, T>> l : perMask) {
                // chunks arrive in target order and every chunk's list is sorted by descending matchingPixels: a stable sort restores the
                // per-mask order of a single search; the optional cap keeps the best maxMatchesPerMask
                l.sort((x, y) -> Integer.compare(y.getMatchingPixels(), x.getMatchingPixels()));
                results.addAll(maxMatchesPerMask > 0 && l.size() > maxMatchesPerMask ? l.subList(0, maxMatchesPerMask) : l);
            }
            CdsGpu.check((int) CdsGpu.hostFree.invokeExact(ctx, pinned));
            CdsGpu.masksetDestroy.invokeExact(ms);
        } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new IllegalStateException(t); }
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
