package jwave.transforms.cuda;

import java.lang.foreign.MemorySegment;

import jwave.exceptions.JWaveException;
import jwave.exceptions.JWaveFailure;
import jwave.transforms.FastWaveletTransform;
import jwave.transforms.wavelets.Wavelet;

/**
 * Drop-in for {@link FastWaveletTransform}: same constructor, same _name ("Fast Wavelet Transform", so
 * TransformBuilder.identify keeps working), same validation and messages (FastWaveletTransform.java:74-83,122-131);
 * the level loops and Wavelet.forward/reverse run as CUDA kernels (jwc_fwt_forward / jwc_fwt_inverse).
 * Inherited decompose/recompose, Complex[], 2-D and 3-D drivers only call these 1-D methods and keep working.
 */
public class CudaFastWaveletTransform extends FastWaveletTransform {

  public CudaFastWaveletTransform(Wavelet wavelet) {
    super(wavelet);
  }

  @Override public double[] forward(double[] arrTime, int level) throws JWaveException {
    check(arrTime.length, level, "forward");
    return JwcNative.run(JwcNative.FWT_FORWARD, CudaContext.get(), arrTime, 1, arrTime.length, level, arrTime.length,
        _wavelet.getScalingDeComposition(), _wavelet.getWaveletDeComposition(), 0);
  }

  @Override public double[] reverse(double[] arrHilb, int level) throws JWaveException {
    check(arrHilb.length, level, "reverse");
    return JwcNative.run(JwcNative.FWT_INVERSE, CudaContext.get(), arrHilb, 1, arrHilb.length, level, arrHilb.length,
        _wavelet.getScalingReConstruction(), _wavelet.getWaveletReConstruction(), 0);
  }

  /** Batch of independent signals, row-major [batch][n], in pinned off-heap segments (zero-copy staging). */
  public void forward(MemorySegment in, MemorySegment out, long batch, int n, int level) throws JWaveException {
    check(n, level, "forward");
    JwcNative.run(JwcNative.FWT_FORWARD, CudaContext.get(), in, out, batch, n, level,
        _wavelet.getScalingDeComposition(), _wavelet.getWaveletDeComposition(), 0);
  }

  public void reverse(MemorySegment in, MemorySegment out, long batch, int n, int level) throws JWaveException {
    check(n, level, "reverse");
    JwcNative.run(JwcNative.FWT_INVERSE, CudaContext.get(), in, out, batch, n, level,
        _wavelet.getScalingReConstruction(), _wavelet.getWaveletReConstruction(), 0);
  }

  private void check(int length, int level, String dir) throws JWaveException {
    if (!isBinary(length))
      throw new JWaveFailure("FastWaveletTransform#" + dir + " - "
          + "given array length is not 2^p | p E N ... = 1, 2, 4, 8, 16, 32, .. "
          + "please use the Ancient Egyptian Decomposition for any other array length!");
    int noOfLevels = calcExponent(length);
    if (level < 0 || level > noOfLevels)
      throw new JWaveFailure("FastWaveletTransform#" + dir + " - given level is out of range for given array");
  }
}
