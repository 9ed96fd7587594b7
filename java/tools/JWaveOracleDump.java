/*
 * JWaveOracleDump -- for anyone WITH a JDK 21 and the reference on the classpath: dumps the real reference's outputs
 * for a little-endian fp64 input file so they can be compared with oracle/jwave_oracle.c and with the GPU path.
 *
 *   javac -cp jwave-pro.jar java/tools/JWaveOracleDump.java
 *   java  -cp jwave-pro.jar:java/tools JWaveOracleDump modwt Daubechies4 6 in.bin out.bin   (DIRECT convolution)
 *   java  ...                          JWaveOracleDump fwt   Daubechies8 20 in.bin out.bin
 *   java  ...                          JWaveOracleDump wpt   Symlet8 6 in.bin out.bin
 *
 * in.bin = one signal of N doubles; out.bin = flattened coefficients.  Not compiled in this repository (no JDK).
 */
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.file.Files;
import java.nio.file.Paths;

import jwave.transforms.FastWaveletTransform;
import jwave.transforms.MODWTTransform;
import jwave.transforms.WaveletPacketTransform;
import jwave.transforms.wavelets.Wavelet;

public class JWaveOracleDump {
  public static void main(String[] a) throws Throwable {
    String kind = a[0];
    String cls = a[1];
    int level = Integer.parseInt(a[2]);
    ByteBuffer bb = ByteBuffer.wrap(Files.readAllBytes(Paths.get(a[3]))).order(ByteOrder.LITTLE_ENDIAN);
    double[] x = new double[bb.remaining() / 8];
    bb.asDoubleBuffer().get(x);
    String pkg = cls.startsWith("Haar") ? "haar" : cls.startsWith("Daub") ? "daubechies"
        : cls.startsWith("Sym") ? "symlets" : "coiflet";
    Wavelet w = (Wavelet) Class.forName("jwave.transforms.wavelets." + pkg + "." + cls).getDeclaredConstructor()
        .newInstance();
    double[] out;
    if (kind.equals("modwt")) {
      MODWTTransform t = new MODWTTransform(w);
      t.setConvolutionMethod(MODWTTransform.ConvolutionMethod.DIRECT);
      double[][] c = t.forwardMODWT(x, level);
      out = new double[(level + 1) * x.length];
      for (int r = 0; r <= level; r++) System.arraycopy(c[r], 0, out, r * x.length, x.length);
    } else if (kind.equals("fwt")) {
      out = new FastWaveletTransform(w).forward(x, level);
    } else {
      out = new WaveletPacketTransform(w).forward(x, level);
    }
    ByteBuffer ob = ByteBuffer.allocate(out.length * 8).order(ByteOrder.LITTLE_ENDIAN);
    ob.asDoubleBuffer().put(out);
    Files.write(Paths.get(a[4]), ob.array());
  }
}
