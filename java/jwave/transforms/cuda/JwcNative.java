/*
 * JwcNative -- Panama FFM (java.lang.foreign) binding of libjwavecuda.so (include/jwavecuda.h).
 *
 * NOT COMPILED IN THIS REPOSITORY: the build image has no JDK (SURVEY.md section 0.2).  Java 21 needs
 * --enable-preview for java.lang.foreign (final in 22).  The same C ABI is exercised by the Python ctypes
 * mirror (jwave-pro_b200/_native.py), which is what the test-suite runs.
 *
 * Library lookup: -Djwave.cuda.lib=/path/to/libjwavecuda.so, else System.loadLibrary-style "jwavecuda".
 */
package jwave.transforms.cuda;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

public final class JwcNative {

  public static final int FLAG_EXACT = 1;          // JWC_FLAG_EXACT: unfused mul/add, bit-identical to the JVM loops
  public static final int FLAG_FORCE_GENERIC = 2;  // JWC_FLAG_FORCE_GENERIC

  private static final Linker LINKER = Linker.nativeLinker();
  private static final SymbolLookup LIB;
  private static final MethodHandle CREATE, DESTROY, LAST_ERROR, ALLOC_PINNED, FREE_PINNED;
  private static final MethodHandle[] TRANSFORMS = new MethodHandle[6];
  private static final MethodHandle[] TRANSFORMS_2D = new MethodHandle[4];
  private static final String[] NAMES_2D = {
      "jwc_fwt2d_forward", "jwc_fwt2d_inverse", "jwc_wpt2d_forward", "jwc_wpt2d_inverse" };
  public static final int FWT2D_FORWARD = 0, FWT2D_INVERSE = 1, WPT2D_FORWARD = 2, WPT2D_INVERSE = 3;
  private static final MethodHandle[] TRANSFORMS_3D = new MethodHandle[4];
  private static final String[] NAMES_3D = {
      "jwc_fwt3d_forward", "jwc_fwt3d_inverse", "jwc_wpt3d_forward", "jwc_wpt3d_inverse" };
  public static final int FWT3D_FORWARD = 0, FWT3D_INVERSE = 1, WPT3D_FORWARD = 2, WPT3D_INVERSE = 3;
  private static final MethodHandle WINDOWS;
  private static final MethodHandle[] TRANSFORMS_AED = new MethodHandle[4];
  private static final String[] NAMES_AED = {
      "jwc_fwt_aed_forward", "jwc_fwt_aed_inverse", "jwc_wpt_aed_forward", "jwc_wpt_aed_inverse" };
  public static final int FWT_AED_FORWARD = 0, FWT_AED_INVERSE = 1, WPT_AED_FORWARD = 2, WPT_AED_INVERSE = 3;
  private static final String[] NAMES = {
      "jwc_modwt_forward", "jwc_modwt_inverse", "jwc_fwt_forward", "jwc_fwt_inverse", "jwc_wpt_forward",
      "jwc_wpt_inverse" };

  public static final int MODWT_FORWARD = 0, MODWT_INVERSE = 1, FWT_FORWARD = 2, FWT_INVERSE = 3, WPT_FORWARD = 4,
      WPT_INVERSE = 5;

  static {
    String path = System.getProperty("jwave.cuda.lib");
    LIB = path != null ? SymbolLookup.libraryLookup(path, Arena.global())
                       : SymbolLookup.libraryLookup(System.mapLibraryName("jwavecuda"), Arena.global());
    CREATE = handle("jwc_create", FunctionDescriptor.of(ADDRESS, ADDRESS, JAVA_INT));
    DESTROY = handle("jwc_destroy", FunctionDescriptor.ofVoid(ADDRESS));
    LAST_ERROR = handle("jwc_last_error", FunctionDescriptor.of(ADDRESS));
    ALLOC_PINNED = handle("jwc_alloc_pinned", FunctionDescriptor.of(ADDRESS, JAVA_LONG));
    FREE_PINNED = handle("jwc_free_pinned", FunctionDescriptor.ofVoid(ADDRESS));
    // int f(jwc_ctx*, const double* in, double* out, int64 batch, int64 n, int levels,
    //       const double* f0, const double* f1, int L, unsigned flags)
    FunctionDescriptor t = FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_INT,
        ADDRESS, ADDRESS, JAVA_INT, JAVA_INT);
    for (int i = 0; i < NAMES.length; i++) TRANSFORMS[i] = handle(NAMES[i], t);
    // int f(jwc_ctx*, const double* in, double* out, int64 batch, int64 rows, int64 cols, int lvlM, int lvlN,
    //       const double* lo, const double* hi, int L, unsigned flags)
    FunctionDescriptor t2 = FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_LONG,
        JAVA_INT, JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT);
    for (int i = 0; i < NAMES_2D.length; i++) TRANSFORMS_2D[i] = handle(NAMES_2D[i], t2);
    // int f(jwc_ctx*, const double* in, double* out, int64 batch, int64 p, int64 q, int64 r, int lvlP, int lvlQ,
    //       int lvlR, const double* lo, const double* hi, int L, unsigned flags)
    FunctionDescriptor t3 = FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_LONG,
        JAVA_LONG, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT);
    for (int i = 0; i < NAMES_3D.length; i++) TRANSFORMS_3D[i] = handle(NAMES_3D[i], t3);
    // int f(jwc_ctx*, const double* in, double* out, int64 batch, int64 n, const double* lo, const double* hi, int L,
    //       unsigned flags)   -- n arbitrary (Ancient-Egyptian blocks)
    FunctionDescriptor ta = FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS,
        ADDRESS, JAVA_INT, JAVA_INT);
    for (int i = 0; i < NAMES_AED.length; i++) TRANSFORMS_AED[i] = handle(NAMES_AED[i], ta);
    // int jwc_modwt_forward_windows(jwc_ctx*, const double* series, double* coeffs, int64 series_len, int64 window,
    //                               int64 hop, int levels, const double* g, const double* h, int L, unsigned flags)
    WINDOWS = handle("jwc_modwt_forward_windows", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG,
        JAVA_LONG, JAVA_LONG, JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT));
  }

  private static MethodHandle handle(String name, FunctionDescriptor fd) {
    return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), fd);
  }

  private JwcNative() { }

  /** One context per process is enough; devices == null means "current CUDA device". No CPU fallback exists. */
  public static MemorySegment create(int[] devices) {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment dev = devices == null ? MemorySegment.NULL : a.allocateArray(JAVA_INT, devices);
      MemorySegment ctx = (MemorySegment) CREATE.invokeExact(dev, devices == null ? 0 : devices.length);
      if (ctx.equals(MemorySegment.NULL)) throw new IllegalStateException("jwc_create: " + lastError());
      return ctx;
    } catch (RuntimeException e) {
      throw e;
    } catch (Throwable t) {
      throw new IllegalStateException(t);
    }
  }

  public static void destroy(MemorySegment ctx) {
    try { DESTROY.invokeExact(ctx); } catch (Throwable t) { throw new IllegalStateException(t); }
  }

  public static String lastError() {
    try {
      MemorySegment p = (MemorySegment) LAST_ERROR.invokeExact();
      return p.reinterpret(512).getUtf8String(0);
    } catch (Throwable t) {
      return t.toString();
    }
  }

  /** Pinned, off-heap staging buffer of `doubles` fp64 values (cudaHostAlloc); free with freePinned. */
  public static MemorySegment allocPinned(long doubles) {
    try {
      MemorySegment p = (MemorySegment) ALLOC_PINNED.invokeExact(doubles * 8L);
      if (p.equals(MemorySegment.NULL)) throw new OutOfMemoryError("jwc_alloc_pinned: " + lastError());
      return p.reinterpret(doubles * 8L);
    } catch (Error e) {
      throw e;
    } catch (Throwable t) {
      throw new IllegalStateException(t);
    }
  }

  public static void freePinned(MemorySegment p) {
    try { FREE_PINNED.invokeExact(p); } catch (Throwable t) { throw new IllegalStateException(t); }
  }

  /** Run one transform on (pinned or plain off-heap) host segments; throws on a non-zero status. */
  public static void run(int which, MemorySegment ctx, MemorySegment in, MemorySegment out, long batch, long n,
      int levels, double[] f0, double[] f1, int flags) {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment s0 = a.allocateArray(JAVA_DOUBLE, f0);
      MemorySegment s1 = a.allocateArray(JAVA_DOUBLE, f1);
      int rc = (int) TRANSFORMS[which].invokeExact(ctx, in, out, batch, n, levels, s0, s1, f0.length, flags);
      if (rc != 0) throw new IllegalStateException(NAMES[which] + " failed (" + rc + "): " + lastError());
    } catch (RuntimeException e) {
      throw e;
    } catch (Throwable t) {
      throw new IllegalStateException(t);
    }
  }

  /** Convenience for the double[] API: copy in, transform, copy out (the reference never mutates its input). */
  public static double[] run(int which, MemorySegment ctx, double[] in, long batch, long n, int levels, int outLen,
      double[] f0, double[] f1, int flags) {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment si = a.allocateArray(JAVA_DOUBLE, in);
      MemorySegment so = a.allocateArray(JAVA_DOUBLE, outLen);
      run(which, ctx, si, so, batch, n, levels, f0, f1, flags);
      return so.toArray(JAVA_DOUBLE);
    }
  }

  /** 2-D transform of `batch` row-major rows x cols matrices held in off-heap segments. */
  public static void run2d(int which, MemorySegment ctx, MemorySegment in, MemorySegment out, long batch, long rows,
      long cols, int lvlM, int lvlN, double[] f0, double[] f1, int flags) {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment s0 = a.allocateArray(JAVA_DOUBLE, f0);
      MemorySegment s1 = a.allocateArray(JAVA_DOUBLE, f1);
      int rc = (int) TRANSFORMS_2D[which].invokeExact(ctx, in, out, batch, rows, cols, lvlM, lvlN, s0, s1, f0.length,
          flags);
      if (rc != 0) throw new IllegalStateException(NAMES_2D[which] + " failed (" + rc + "): " + lastError());
    } catch (RuntimeException e) {
      throw e;
    } catch (Throwable t) {
      throw new IllegalStateException(t);
    }
  }

  /** double[][] convenience: flatten, transform, un-flatten (the reference allocates a fresh matrix as well). */
  public static double[][] run2d(int which, MemorySegment ctx, double[][] mat, int lvlM, int lvlN, double[] f0,
      double[] f1, int flags) {
    int rows = mat.length, cols = mat[0].length;
    try (Arena a = Arena.ofConfined()) {
      MemorySegment si = a.allocateArray(JAVA_DOUBLE, (long) rows * cols);
      MemorySegment so = a.allocateArray(JAVA_DOUBLE, (long) rows * cols);
      for (int i = 0; i < rows; i++)
        MemorySegment.copy(mat[i], 0, si, JAVA_DOUBLE, (long) i * cols * Double.BYTES, cols);
      run2d(which, ctx, si, so, 1, rows, cols, lvlM, lvlN, f0, f1, flags);
      double[][] out = new double[rows][cols];
      for (int i = 0; i < rows; i++)
        MemorySegment.copy(so, JAVA_DOUBLE, (long) i * cols * Double.BYTES, out[i], 0, cols);
      return out;
    }
  }

  /** Spaces [batch][p][q][r] (BasicTransform.java:487-640), one native call. */
  public static void run3d(int which, MemorySegment ctx, MemorySegment in, MemorySegment out, long batch, long p,
      long q, long r, int lvlP, int lvlQ, int lvlR, double[] f0, double[] f1, int flags) {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment s0 = a.allocateArray(JAVA_DOUBLE, f0);
      MemorySegment s1 = a.allocateArray(JAVA_DOUBLE, f1);
      int rc = (int) TRANSFORMS_3D[which].invokeExact(ctx, in, out, batch, p, q, r, lvlP, lvlQ, lvlR, s0, s1,
          f0.length, flags);
      if (rc != 0) throw new IllegalStateException(NAMES_3D[which] + " failed (" + rc + "): " + lastError());
    } catch (RuntimeException e) {
      throw e;
    } catch (Throwable t) {
      throw new IllegalStateException(t);
    }
  }

  /** double[][][] convenience: flatten, transform, un-flatten (the reference allocates a fresh space as well). */
  public static double[][][] run3d(int which, MemorySegment ctx, double[][][] spc, int lvlP, int lvlQ, int lvlR,
      double[] f0, double[] f1, int flags) {
    int p = spc.length, q = spc[0].length, r = spc[0][0].length;
    try (Arena a = Arena.ofConfined()) {
      MemorySegment si = a.allocateArray(JAVA_DOUBLE, (long) p * q * r);
      MemorySegment so = a.allocateArray(JAVA_DOUBLE, (long) p * q * r);
      for (int i = 0; i < p; i++)
        for (int j = 0; j < q; j++)
          MemorySegment.copy(spc[i][j], 0, si, JAVA_DOUBLE, ((long) i * q + j) * r * Double.BYTES, r);
      run3d(which, ctx, si, so, 1, p, q, r, lvlP, lvlQ, lvlR, f0, f1, flags);
      double[][][] out = new double[p][q][r];
      for (int i = 0; i < p; i++)
        for (int j = 0; j < q; j++)
          MemorySegment.copy(so, JAVA_DOUBLE, ((long) i * q + j) * r * Double.BYTES, out[i][j], 0, r);
      return out;
    }
  }

  /** Arbitrary-length batch [batch][n] through the Ancient-Egyptian block decomposition, one native call. */
  public static void runAed(int which, MemorySegment ctx, MemorySegment in, MemorySegment out, long batch, long n,
      double[] f0, double[] f1, int flags) {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment s0 = a.allocateArray(JAVA_DOUBLE, f0);
      MemorySegment s1 = a.allocateArray(JAVA_DOUBLE, f1);
      int rc = (int) TRANSFORMS_AED[which].invokeExact(ctx, in, out, batch, n, s0, s1, f0.length, flags);
      if (rc != 0) throw new IllegalStateException(NAMES_AED[which] + " failed (" + rc + "): " + lastError());
    } catch (RuntimeException e) {
      throw e;
    } catch (Throwable t) {
      throw new IllegalStateException(t);
    }
  }

  public static double[] runAed(int which, MemorySegment ctx, double[] in, double[] f0, double[] f1, int flags) {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment si = a.allocateArray(JAVA_DOUBLE, in);
      MemorySegment so = a.allocateArray(JAVA_DOUBLE, in.length);
      runAed(which, ctx, si, so, 1, in.length, f0, f1, flags);
      return so.toArray(JAVA_DOUBLE);
    }
  }

  /** Forward MODWT of every window series[w*hop .. w*hop + window) of one series; coeffs [nwin][levels+1][window]. */
  public static void runWindows(MemorySegment ctx, MemorySegment series, MemorySegment coeffs, long seriesLength,
      long window, long hop, int levels, double[] g, double[] h, int flags) {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment s0 = a.allocateArray(JAVA_DOUBLE, g);
      MemorySegment s1 = a.allocateArray(JAVA_DOUBLE, h);
      int rc = (int) WINDOWS.invokeExact(ctx, series, coeffs, seriesLength, window, hop, levels, s0, s1, g.length, flags);
      if (rc != 0) throw new IllegalStateException("jwc_modwt_forward_windows failed (" + rc + "): " + lastError());
    } catch (RuntimeException e) {
      throw e;
    } catch (Throwable t) {
      throw new IllegalStateException(t);
    }
  }
}
