package jwave.transforms.cuda;

import java.lang.foreign.MemorySegment;

import jwave.exceptions.JWaveException;
import jwave.exceptions.JWaveFailure;
import jwave.transforms.WaveletPacketTransform;
import jwave.transforms.wavelets.Wavelet;

/**
 * Drop-in for {@link WaveletPacketTransform} (and for ParallelWaveletPacketTransform, whose results it equals):
 * full packet tree, leaves in natural (Paley) order, validation as WaveletPacketTransform.java:76-84,144-152.
 */
public class CudaWaveletPacketTransform extends WaveletPacketTransform {

  public CudaWaveletPacketTransform(Wavelet wavelet) {
    super(wavelet);
  }

  @Override public double[] forward(double[] arrTime, int level) throws JWaveException {
    check(arrTime.length, level, "forward");
    return JwcNative.run(JwcNative.WPT_FORWARD, CudaContext.get(), arrTime, 1, arrTime.length, level, arrTime.length,
        _wavelet.getScalingDeComposition(), _wavelet.getWaveletDeComposition(), 0);
  }

  @Override public double[] reverse(double[] arrHilb, int level) throws JWaveException {
    check(arrHilb.length, level, "reverse");
    return JwcNative.run(JwcNative.WPT_INVERSE, CudaContext.get(), arrHilb, 1, arrHilb.length, level, arrHilb.length,
        _wavelet.getScalingReConstruction(), _wavelet.getWaveletReConstruction(), 0);
  }

  public void forward(MemorySegment in, MemorySegment out, long batch, int n, int level) throws JWaveException {
    check(n, level, "forward");
    JwcNative.run(JwcNative.WPT_FORWARD, CudaContext.get(), in, out, batch, n, level,
        _wavelet.getScalingDeComposition(), _wavelet.getWaveletDeComposition(), 0);
  }

  public void reverse(MemorySegment in, MemorySegment out, long batch, int n, int level) throws JWaveException {
    check(n, level, "reverse");
    JwcNative.run(JwcNative.WPT_INVERSE, CudaContext.get(), in, out, batch, n, level,
        _wavelet.getScalingReConstruction(), _wavelet.getWaveletReConstruction(), 0);
  }

  /**
   * 2-D forward, BasicTransform.java:361-399: every row through forward(row, lvlN), then every column through
   * forward(col, lvlM) -- here one row pass and an in-place column pass on the device (jwc_wpt2d_forward) instead of
   * rows + cols separate 1-D calls.  forward(double[][]) (:336-340) delegates here with full depth.
   */
  @Override public double[][] forward(double[][] matTime, int lvlM, int lvlN) throws JWaveException {
    check(matTime[0].length, lvlN, "forward");
    check(matTime.length, lvlM, "forward");
    return JwcNative.run2d(JwcNative.WPT2D_FORWARD, CudaContext.get(), matTime, lvlM, lvlN,
        _wavelet.getScalingDeComposition(), _wavelet.getWaveletDeComposition(), 0);
  }

  /** 2-D reverse, BasicTransform.java:436-474: columns (lvlM) first, then rows (lvlN). */
  @Override public double[][] reverse(double[][] matFreq, int lvlM, int lvlN) throws JWaveException {
    check(matFreq[0].length, lvlN, "reverse");
    check(matFreq.length, lvlM, "reverse");
    return JwcNative.run2d(JwcNative.WPT2D_INVERSE, CudaContext.get(), matFreq, lvlM, lvlN,
        _wavelet.getScalingReConstruction(), _wavelet.getWaveletReConstruction(), 0);
  }

  private void check(int length, int level, String dir) throws JWaveException {
    if (!isBinary(length))
      throw new JWaveFailure("given array length is not 2^p | p E N ... = 1, 2, 4, 8, 16, 32, .. "
          + "please use the Ancient Egyptian Decomposition for any other array length!");
    int noOfLevels = calcExponent(length);
    if (level < 0 || level > noOfLevels)
      throw new JWaveFailure("WaveletPacketTransform#" + dir + " - given level is out of range for given array");
  }
}
