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
 * Inherited decompose/recompose, Complex[] and 3-D drivers only call these 1-D methods and keep working; the 2-D
 * matrix overloads are overridden to run as one device call.
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

  /**
   * 2-D forward, BasicTransform.java:361-399: every row through forward(row, lvlN), then every column through
   * forward(col, lvlM) -- here one row pass and an in-place column pass on the device (jwc_fwt2d_forward) instead of
   * rows + cols separate 1-D calls.  forward(double[][]) (:336-340) delegates here with full depth.
   */
  @Override public double[][] forward(double[][] matTime, int lvlM, int lvlN) throws JWaveException {
    check(matTime[0].length, lvlN, "forward");
    check(matTime.length, lvlM, "forward");
    return JwcNative.run2d(JwcNative.FWT2D_FORWARD, CudaContext.get(), matTime, lvlM, lvlN,
        _wavelet.getScalingDeComposition(), _wavelet.getWaveletDeComposition(), 0);
  }

  /** 2-D reverse, BasicTransform.java:436-474: columns (lvlM) first, then rows (lvlN). */
  @Override public double[][] reverse(double[][] matFreq, int lvlM, int lvlN) throws JWaveException {
    check(matFreq[0].length, lvlN, "reverse");
    check(matFreq.length, lvlM, "reverse");
    return JwcNative.run2d(JwcNative.FWT2D_INVERSE, CudaContext.get(), matFreq, lvlM, lvlN,
        _wavelet.getScalingReConstruction(), _wavelet.getWaveletReConstruction(), 0);
  }

  /**
   * 3-D forward, BasicTransform.java:509-565: the 2-D forward(mat, lvlP, lvlQ) of every matrix spcTime[i] -- rows of
   * length r with lvlQ, columns of length q with lvlP -- then every line along the first axis with lvlR; here the 2-D
   * passes of all matrices and one in-place first-axis pass on the device (jwc_fwt3d_forward).  forward(double[][][])
   * (:487-495) delegates here with the exponents of the three dimensions.
   */
  @Override public double[][][] forward(double[][][] spcTime, int lvlP, int lvlQ, int lvlR) throws JWaveException {
    check(spcTime[0][0].length, lvlQ, "forward");
    check(spcTime[0].length, lvlP, "forward");
    check(spcTime.length, lvlR, "forward");
    return JwcNative.run3d(JwcNative.FWT3D_FORWARD, CudaContext.get(), spcTime, lvlP, lvlQ, lvlR,
        _wavelet.getScalingDeComposition(), _wavelet.getWaveletDeComposition(), 0);
  }

  /** 3-D reverse, BasicTransform.java:602-640: the 2-D reverse of every matrix first, then the first axis (lvlR). */
  @Override public double[][][] reverse(double[][][] spcHilb, int lvlP, int lvlQ, int lvlR) throws JWaveException {
    check(spcHilb[0][0].length, lvlQ, "reverse");
    check(spcHilb[0].length, lvlP, "reverse");
    check(spcHilb.length, lvlR, "reverse");
    return JwcNative.run3d(JwcNative.FWT3D_INVERSE, CudaContext.get(), spcHilb, lvlP, lvlQ, lvlR,
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
