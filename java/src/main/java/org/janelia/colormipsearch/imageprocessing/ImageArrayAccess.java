package org.janelia.colormipsearch.imageprocessing;

/**
 * ImageArray keeps its pixel array and type package-private (ImageArray.java:14-17, 65), so the GPU providers need this
 * one accessor declared inside the package.  No copies: the arrays are handed to the native layer, which copies on upload
 * and never retains them.
 */
public final class ImageArrayAccess {
    public static boolean isRGB(ImageArray<?> a) { return a instanceof ColorImageArray; }
    public static boolean isGray16(ImageArray<?> a) { return a instanceof ShortImageArray; }
    public static boolean isGray8(ImageArray<?> a) { return a instanceof ByteImageArray; }
    /** interleaved R,G,B bytes of a ColorImageArray (ColorImageArray.java:6-31) */
    public static byte[] rgbBytes(ImageArray<?> a) { return ((ColorImageArray) a).getPixels(); }
    public static short[] gray16(ImageArray<?> a) { return ((ShortImageArray) a).getPixels(); }
    public static byte[] gray8(ImageArray<?> a) { return ((ByteImageArray) a).getPixels(); }
    private ImageArrayAccess() {}
}
