package jwave.transforms.cuda;

import java.lang.foreign.MemorySegment;

import jwave.transforms.EfficientMODWTTransform;
import jwave.transforms.MODWTTransform;
import jwave.transforms.wavelets.Wavelet;

/**
 * Drop-in for {@link MODWTTransform}: forwardMODWT / inverseMODWT run as fused CUDA kernels
 * (jwc_modwt_forward / jwc_modwt_inverse).  The flattened 1-D API (forward/reverse(double[]) and (double[], int)),
 * precomputeFilters, clearFilterCache and getMaxDecompositionLevel are inherited: they only call the two methods
 * overridden here.  Results equal the reference's DIRECT convolution (MODWTTransform.java:677-716) to ~1e-16; the
 * reference's default FFT path is itself 1.6e-12 away from that at N = 65536 (SURVEY.md section 0.3).
 *
 * Validation order and messages follow MODWTTransform.java:257-282.  The base filters g~, h~ are recomputed here
 * (the reference keeps them in private fields): normalise by the L2 norm, then divide by sqrt(2)
 * (MODWTTransform.java:462-475, 599-606).  Thread-safe: the filters are immutable once published and the native
 * context is re-entrant.
 */
public class CudaMODWTTransform extends MODWTTransform {

  private static final int MAX_LEVEL = 13;   // MODWTTransform.java:111
  private volatile double[][] base;          // {g~, h~}

  public CudaMODWTTransform(Wavelet wavelet) {
    super(wavelet);
  }

  private double[][] filters() {
    double[][] b = base;
    if (b == null) {
      double[] g = normalize(_wavelet.getScalingDeComposition());
      double[] h = normalize(_wavelet.getWaveletDeComposition());
      double s = Math.sqrt(2.0);
      for (int i = 0; i < g.length; i++) { g[i] = g[i] / s; h[i] = h[i] / s; }
      base = b = new double[][] { g, h };
    }
    return b;
  }

  private static double[] normalize(double[] f) {
    double energy = 0.0;
    for (double c : f) energy += c * c;
    double norm = Math.sqrt(energy);
    if (norm > 1e-12) for (int i = 0; i < f.length; i++) f[i] /= norm;
    return f;
  }

  private static void checkLevel(int maxLevel) {
    if (maxLevel < 1)
      throw new IllegalArgumentException("MODWTTransform#forwardMODWT - "
          + "decomposition level must be at least 1, requested: " + maxLevel);
    if (maxLevel > MAX_LEVEL)
      throw new IllegalArgumentException("MODWTTransform#forwardMODWT - "
          + "maximum supported decomposition level is " + MAX_LEVEL + ", requested: " + maxLevel);
  }

  private static void checkLimit(int maxLevel, int n) {
    int limit = n > 0 ? 31 - Integer.numberOfLeadingZeros(n) : 0;
    if (maxLevel > limit)
      throw new IllegalArgumentException("Decomposition level " + maxLevel + " exceeds theoretical limit "
          + limit + " for signal length " + n);
  }

  @Override public double[][] forwardMODWT(double[] data, int maxLevel) {
    checkLevel(maxLevel);
    if (data == null || data.length == 0) {
      double[][] empty = new double[maxLevel + 1][];
      for (int i = 0; i <= maxLevel; i++) empty[i] = new double[0];
      return empty;
    }
    int n = data.length;
    checkLimit(maxLevel, n);
    double[][] f = filters();
    double[] flat = JwcNative.run(JwcNative.MODWT_FORWARD, CudaContext.get(), data, 1, n, maxLevel,
        (maxLevel + 1) * n, f[0], f[1], 0);
    double[][] rows = new double[maxLevel + 1][n];
    for (int r = 0; r <= maxLevel; r++) System.arraycopy(flat, r * n, rows[r], 0, n);
    return rows;
  }

  @Override public double[] inverseMODWT(double[][] coefficients) {
    if (coefficients == null || coefficients.length == 0) return new double[0];
    int maxLevel = coefficients.length - 1;
    if (maxLevel <= 0) return new double[0];
    int n = coefficients[0].length;
    if (n == 0) return new double[0];
    double[] flat = new double[(maxLevel + 1) * n];
    for (int r = 0; r <= maxLevel; r++) System.arraycopy(coefficients[r], 0, flat, r * n, n);
    double[][] f = filters();
    return JwcNative.run(JwcNative.MODWT_INVERSE, CudaContext.get(), flat, 1, n, maxLevel, n, f[0], f[1], 0);
  }

  /**
   * The result of forwardMODWT in the reference's single-backing-array wire format
   * ({@link EfficientMODWTTransform.MODWTCoefficients}, EfficientMODWTTransform.java:28-86; level views are
   * {@link EfficientMODWTTransform.ArrayView}, :88-117): the flat array [W_1 | ... | W_J | V_J] the device writes IS the
   * backing array, so wrapping a GPU result copies nothing.  Rows carry forwardMODWT's meaning (W_j = h~ conv V_{j-1},
   * V_J last, MODWTTransform.java:298-303 / flat form :406-416); the reference's forwardMODWTEfficient (:151-170) labels
   * its two branches the other way round and is covered by none of its tests, so that labelling is not reproduced.
   */
  public EfficientMODWTTransform.MODWTCoefficients forwardMODWTCoefficients(double[] data, int maxLevel) {
    checkLevel(maxLevel);
    if (data == null || data.length == 0)
      return new EfficientMODWTTransform.MODWTCoefficients(new double[0], 0, maxLevel);
    int n = data.length;
    checkLimit(maxLevel, n);
    double[][] f = filters();
    double[] flat = JwcNative.run(JwcNative.MODWT_FORWARD, CudaContext.get(), data, 1, n, maxLevel,
        (maxLevel + 1) * n, f[0], f[1], 0);
    return new EfficientMODWTTransform.MODWTCoefficients(flat, n, maxLevel);
  }

  /**
   * Inverse of {@link #forwardMODWTCoefficients}.  MODWTCoefficients keeps its backing array private, so the rows are
   * read through its level views (one copy per row, as a caller of the reference class would have to do) and go to
   * the device as one flat array.
   */
  public double[] inverseMODWTCoefficients(EfficientMODWTTransform.MODWTCoefficients coeffs, int signalLength,
      int maxLevel) {
    checkLevel(maxLevel);
    if (coeffs == null || signalLength == 0) return new double[0];
    if (coeffs.getTotalSize() != (maxLevel + 1) * signalLength)
      throw new IllegalArgumentException("backing array length " + coeffs.getTotalSize()
          + " != (levels + 1) * signalLength");
    double[] flat = new double[(maxLevel + 1) * signalLength];
    for (int r = 0; r <= maxLevel; r++) {
      EfficientMODWTTransform.ArrayView v = coeffs.getView(r + 1);
      for (int i = 0; i < signalLength; i++) flat[r * signalLength + i] = v.get(i);
    }
    double[][] f = filters();
    return JwcNative.run(JwcNative.MODWT_INVERSE, CudaContext.get(), flat, 1, signalLength, maxLevel, signalLength,
        f[0], f[1], 0);
  }

  /** Flat form of the inverse: coeffs = [W_1 | ... | W_J | V_J], the layout of MODWTTransform.forward(double[], level). */
  public double[] inverseMODWTFlat(double[] flat, int signalLength, int maxLevel) {
    checkLevel(maxLevel);
    if (flat == null || signalLength == 0) return new double[0];
    if (flat.length != (maxLevel + 1) * signalLength)
      throw new IllegalArgumentException("coefficient array length " + flat.length + " != (levels + 1) * signalLength");
    double[][] f = filters();
    return JwcNative.run(JwcNative.MODWT_INVERSE, CudaContext.get(), flat, 1, signalLength, maxLevel, signalLength,
        f[0], f[1], 0);
  }

  /**
   * Batch: x [batch][n] -> coeffs [batch][maxLevel+1][n] (rows W_1..W_J, V_J per signal), pinned off-heap segments.
   * A Java double[] cannot hold the BASELINE batches (8192 x 65536 x 9 doubles), hence MemorySegment.
   */
  public void forwardMODWT(MemorySegment x, MemorySegment coeffs, long batch, int n, int maxLevel) {
    checkLevel(maxLevel);
    checkLimit(maxLevel, n);
    double[][] f = filters();
    JwcNative.run(JwcNative.MODWT_FORWARD, CudaContext.get(), x, coeffs, batch, n, maxLevel, f[0], f[1], 0);
  }

  public void inverseMODWT(MemorySegment coeffs, MemorySegment x, long batch, int n, int maxLevel) {
    checkLevel(maxLevel);
    double[][] f = filters();
    JwcNative.run(JwcNative.MODWT_INVERSE, CudaContext.get(), coeffs, x, batch, n, maxLevel, f[0], f[1], 0);
  }

  /**
   * Sliding-window analysis of one long series (the loop of MODWTSlidingWindowTest.java:20-70 without the per-window
   * arraycopy): window w = series[w*hop .. w*hop + window), coeffs [nWindows][maxLevel+1][window], with
   * nWindows = (seriesLength - window) / hop + 1.  The windows are read in place on the device.
   */
  public void forwardMODWTWindows(MemorySegment series, MemorySegment coeffs, long seriesLength, int window, long hop,
      int maxLevel) {
    checkLevel(maxLevel);
    checkLimit(maxLevel, window);
    if (window < 1 || hop < 1 || seriesLength < window)
      throw new IllegalArgumentException("need 1 <= window <= series length and hop >= 1");
    double[][] f = filters();
    JwcNative.runWindows(CudaContext.get(), series, coeffs, seriesLength, window, hop, maxLevel, f[0], f[1], 0);
  }
}
