package jwave.transforms.cuda;

import java.lang.foreign.MemorySegment;

import jwave.exceptions.JWaveException;
import jwave.exceptions.JWaveFailure;
import jwave.transforms.AncientEgyptianDecomposition;
import jwave.transforms.WaveletPacketTransform;
import jwave.transforms.WaveletTransform;

/**
 * Drop-in for {@link AncientEgyptianDecomposition} around one of the CUDA pyramid transforms: the block loop of
 * AncientEgyptianDecomposition.java:97-181 (one forward()/reverse() per 2^p block, with a copy in and out) becomes one
 * native call that transforms every block of every signal where it lies (jwc_fwt_aed_* / jwc_wpt_aed_*).
 * The plain reference wrapper also works with the CUDA transforms (it only calls forward(double[])), block by block.
 */
public class CudaAncientEgyptianDecomposition extends AncientEgyptianDecomposition {

  private final WaveletTransform _transform;
  private final boolean _packet;

  public CudaAncientEgyptianDecomposition(WaveletTransform transform) throws JWaveException {
    super(transform);
    if (!(transform instanceof CudaFastWaveletTransform) && !(transform instanceof CudaWaveletPacketTransform))
      throw new JWaveFailure("CudaAncientEgyptianDecomposition wraps CudaFastWaveletTransform or "
          + "CudaWaveletPacketTransform");
    _transform = transform;
    _packet = transform instanceof WaveletPacketTransform;
  }

  @Override public double[] forward(double[] arrTime) throws JWaveException {
    if (arrTime.length < 1) throw new JWaveFailure("the supported number for decomposition is smaller than one");
    return JwcNative.runAed(_packet ? JwcNative.WPT_AED_FORWARD : JwcNative.FWT_AED_FORWARD, CudaContext.get(), arrTime,
        _transform.getWavelet().getScalingDeComposition(), _transform.getWavelet().getWaveletDeComposition(), 0);
  }

  @Override public double[] reverse(double[] arrHilb) throws JWaveException {
    if (arrHilb.length < 1) throw new JWaveFailure("the supported number for decomposition is smaller than one");
    return JwcNative.runAed(_packet ? JwcNative.WPT_AED_INVERSE : JwcNative.FWT_AED_INVERSE, CudaContext.get(), arrHilb,
        _transform.getWavelet().getScalingReConstruction(), _transform.getWavelet().getWaveletReConstruction(), 0);
  }

  /** Batch of arbitrary-length signals, row-major [batch][n], off-heap (pinned) segments. */
  public void forward(MemorySegment in, MemorySegment out, long batch, long n) {
    JwcNative.runAed(_packet ? JwcNative.WPT_AED_FORWARD : JwcNative.FWT_AED_FORWARD, CudaContext.get(), in, out, batch,
        n, _transform.getWavelet().getScalingDeComposition(), _transform.getWavelet().getWaveletDeComposition(), 0);
  }

  public void reverse(MemorySegment in, MemorySegment out, long batch, long n) {
    JwcNative.runAed(_packet ? JwcNative.WPT_AED_INVERSE : JwcNative.FWT_AED_INVERSE, CudaContext.get(), in, out, batch,
        n, _transform.getWavelet().getScalingReConstruction(), _transform.getWavelet().getWaveletReConstruction(), 0);
  }
}
