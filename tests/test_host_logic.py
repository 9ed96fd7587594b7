"""CPU-only: validation order, exception types and message substrings of the host mirror
(reference tests MODWT1DInterfaceTest.java:108-137, MODWTTheoreticalLimitTest.java:40-56,
MODWTLevelLimitTest.java:20-87, ParallelWPTTest.java:128-151).  None of these reach the native library."""
import numpy as np
import pytest


def test_modwt_forward_level_checks(jw):
    t = jw.CudaMODWTTransform(jw.wavelets.Daubechies4())
    x = np.ones(64)
    with pytest.raises(jw.IllegalArgumentException, match="at least 1"):
        t.forwardMODWT(x, 0)
    with pytest.raises(jw.IllegalArgumentException, match="maximum supported decomposition level is 13"):
        t.forwardMODWT(x, 14)
    with pytest.raises(jw.IllegalArgumentException, match="exceeds theoretical limit 6 for signal length 64"):
        t.forwardMODWT(x, 7)
    # order: level checks come before the null/empty check (MODWTTransform.java:257-273)
    with pytest.raises(jw.IllegalArgumentException, match="at least 1"):
        t.forwardMODWT(None, 0)
    rows = t.forwardMODWT(None, 3)
    assert len(rows) == 4 and all(len(r) == 0 for r in rows)
    rows = t.forwardMODWT([], 2)
    assert len(rows) == 3
    assert jw.CudaMODWTTransform.getMaxDecompositionLevel() == 13
    # floor(log2 N) for non powers of two (MODWTLog2CalculationTest.java:90-121)
    with pytest.raises(jw.IllegalArgumentException, match="theoretical limit 6 for signal length 100"):
        t.forwardMODWT(np.ones(100), 7)


def test_modwt_inverse_degenerate_inputs(jw):
    t = jw.CudaMODWTTransform(jw.wavelets.Haar1())
    assert len(t.inverseMODWT(None)) == 0
    assert len(t.inverseMODWT([])) == 0
    assert len(t.inverseMODWT([[1.0, 2.0]])) == 0   # < 2 rows (MODWTTransform.java:342-346)


def test_modwt_flat_interface_errors(jw):
    t = jw.CudaMODWTTransform(jw.wavelets.Haar1())
    assert len(t.forward([], 2)) == 0 and len(t.forward(None)) == 0 and len(t.reverse([], 1)) == 0
    with pytest.raises(jw.JWaveFailure, match=r"2\^p"):
        t.forward(np.ones(7), 2)
    with pytest.raises(jw.JWaveFailure, match="out of range"):
        t.forward(np.ones(8), -1)
    with pytest.raises(jw.JWaveFailure, match="out of range"):
        t.forward(np.ones(8), 4)
    with pytest.raises(jw.IllegalArgumentException, match="at least 1"):
        t.forward(np.ones(8), 0)          # passes the JWaveFailure checks, then forwardMODWT's
    with pytest.raises(jw.JWaveFailure, match="maximum supported"):
        t.forward(np.ones(1 << 14), 14)
    with pytest.raises(jw.JWaveFailure, match="Invalid coefficient array"):
        t.reverse(np.ones(21), 2)         # 21 / 3 = 7 is not 2^p
    with pytest.raises(jw.JWaveFailure, match="does not match"):
        t.reverse(np.ones(25), 2)         # 25 // 3 = 8 but 8*3 != 25
    with pytest.raises(jw.IllegalArgumentException):
        t.forward(np.ones(1 << 14))       # level-less forward on N > 8192: J = 14 > 13, unchecked (A.6)
    with pytest.raises(jw.JWaveFailure):
        t.forward(np.ones(12))            # calcExponent on a non power of two


def test_fwt_wpt_shape_errors(jw):
    for cls, name in ((jw.CudaFastWaveletTransform, "Fast Wavelet Transform"),
                      (jw.CudaWaveletPacketTransform, "Wavelet Packet Transform")):
        t = cls(jw.wavelets.Daubechies4())
        assert t.getName() == name
        with pytest.raises(jw.JWaveFailure, match=r"2\^p"):
            t.forward(np.ones(17), 2)
        with pytest.raises(jw.JWaveFailure, match=r"2\^p"):
            t.forward(np.ones(17))
        with pytest.raises(jw.JWaveFailure, match="given level is out of range for given array"):
            t.forward(np.ones(8), 10)
        with pytest.raises(jw.JWaveFailure, match="given level is out of range for given array"):
            t.reverse(np.ones(8), -1)
        with pytest.raises(jw.JWaveFailure, match=r"2\^p"):
            t.forward(np.ones(0))
        with pytest.raises(jw.JWaveFailure, match="out of range"):
            t.recompose(np.ones((4, 8)), 4)


def test_modwt_base_filters_match_oracle(jw, oracle):
    for cls in jw.wavelets.ALL_CLASSES:
        w = jw.wavelets.create(cls)
        t = jw.CudaMODWTTransform(w)
        t.initializeFilterCache()
        g, h = oracle.modwt_filters(w.getScalingDeComposition(), w.getWaveletDeComposition())
        assert np.array_equal(t._g, g) and np.array_equal(t._h, h)
        t.clearFilterCache()
        assert t._g is None
        t.precomputeFilters(5)
        assert np.array_equal(t._g, g)
    with pytest.raises(jw.IllegalArgumentException):
        jw.CudaMODWTTransform(jw.wavelets.Haar1()).precomputeFilters(14)


def test_getters_return_copies(jw):
    w = jw.wavelets.Symlet8()
    a = w.getScalingDeComposition()
    a[0] = 99.0
    assert w.getScalingDeComposition()[0] != 99.0
    assert w.getMotherWavelength() == 16 and w.getTransformWavelength() == 2
    assert jw.wavelets.create("Daubechies 20").getMotherWavelength() == 40


def test_modwt_coefficients_wire_format(jw):
    """EfficientMODWTTransform.java:28-117: one backing array [W_1|...|W_J|V_J], level views without copies."""
    n, J = 8, 2
    backing = np.arange((J + 1) * n, dtype=np.float64)
    c = jw.MODWTCoefficients(backing, n, J)
    assert c.getTotalSize() == 24
    assert np.array_equal(c.getDetails(1), backing[:8]) and np.array_equal(c.getDetails(2), backing[8:16])
    assert np.array_equal(c.getApproximation(), backing[16:])
    v = c.getView(3)
    assert v.length() == 8 and v.get(0) == 16.0 and np.array_equal(v.toArray(), backing[16:])
    with pytest.raises(IndexError):
        v.get(8)
    for bad in (0, 3):
        with pytest.raises(jw.IllegalArgumentException, match="Invalid level"):
            c.getDetails(bad)
    with pytest.raises(jw.IllegalArgumentException, match="Invalid level"):
        c.getView(4)
    backing[16] = -1.0          # a view reads the backing array, a copy does not
    assert v.get(0) == -1.0


def test_ancient_egyptian_multipliers(jw, oracle):
    """tools/MathToolKit.java:57-84 decompose (e.g. 42 = 2^5 + 2^3 + 2^1; the javadoc's 127 = 64|32|16|8|4|2|1)."""
    dec = jw.AncientEgyptianDecomposition.decompose
    assert dec(42) == [5, 3, 1]
    assert dec(127) == [6, 5, 4, 3, 2, 1, 0]
    assert dec(1) == [0] and dec(1 << 20) == [20]
    for n in (3, 17, 1000, 65537, 2 ** 31 - 1):
        assert sum(1 << p for p in dec(n)) == n and dec(n) == oracle.aed_blocks(n)
    with pytest.raises(jw.JWaveFailure):
        dec(0)


def test_2d_and_window_argument_checks_need_no_gpu(jw):
    """Validation happens on the host before any native call, with the reference's exception types and message
    substrings (BasicTransform.java:336-399 delegates to the 1-D checks of FastWaveletTransform.java:74-83)."""
    f = jw.CudaFastWaveletTransform(jw.wavelets.Daubechies4())
    with pytest.raises(jw.JWaveFailure, match=r"2\^p"):
        f.forward(np.zeros((6, 8)))                 # rows not a power of two
    with pytest.raises(jw.JWaveFailure, match=r"2\^p"):
        f.forward2DBatch(np.zeros((2, 8, 12)))      # cols not a power of two
    with pytest.raises(jw.JWaveFailure, match="out of range"):
        f.forward(np.zeros((8, 8)), 4, 3)
    with pytest.raises(jw.JWaveFailure, match="out of range"):
        f.reverse(np.zeros((8, 8)), 3, -1)
    # 3-D overloads (BasicTransform.java:487-640): lvlP goes to the q axis, lvlQ to the r axis, lvlR to the p axis
    with pytest.raises(jw.JWaveFailure, match=r"2\^p"):
        f.forward(np.zeros((8, 6, 8)))
    with pytest.raises(jw.JWaveFailure, match="out of range"):
        f.forward(np.zeros((32, 8, 8)))             # default levels (5, 3, 3): lvlP = 5 > log2(q) = 3, as in the reference
    with pytest.raises(jw.JWaveFailure, match="out of range"):
        f.reverse(np.zeros((8, 8, 8)), 3, 3, 4)
    m = jw.CudaMODWTTransform(jw.wavelets.Haar1())
    with pytest.raises(jw.IllegalArgumentException):
        m.forwardMODWTWindows(np.zeros(100), 128, 16, 3)      # window longer than the series
    with pytest.raises(jw.IllegalArgumentException):
        m.forwardMODWTWindows(np.zeros(100), 32, 0, 3)        # hop < 1
    with pytest.raises(jw.IllegalArgumentException, match="exceeds theoretical limit"):
        m.forwardMODWTWindows(np.zeros(100), 32, 8, 6)        # J > log2(window)
    with pytest.raises(jw.JWaveFailure):
        jw.AncientEgyptianDecomposition(m)                     # only the pyramid transforms can be wrapped


def test_modwt_constructor_variants_and_convolution_method(jw):
    """MODWTTransform.java:180-213: (wavelet) / (wavelet, fftThreshold) constructors and the ConvolutionMethod accessors;
    on the device there is one arithmetic, so the setting is stored and read back, nothing else."""
    w = jw.wavelets.Daubechies4()
    t = jw.CudaMODWTTransform(w)
    assert t.getConvolutionMethod() == jw.ConvolutionMethod.AUTO and t.ConvolutionMethod is jw.ConvolutionMethod
    t.setConvolutionMethod(jw.ConvolutionMethod.FFT)
    assert t.getConvolutionMethod() == "FFT"
    with pytest.raises(jw.IllegalArgumentException):
        t.setConvolutionMethod("fastest")
    assert jw.CudaMODWTTransform(w, 1024)._fftThreshold == 1024
    assert jw.CudaMODWTTransform(w)._fftThreshold == 4096          # MODWTTransform.java:144
    assert jw.CudaMODWTTransform.getMaxDecompositionLevel() == 13
