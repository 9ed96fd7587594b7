/* oracle/jwave_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C99, built with -ffp-contract=off so every multiply and add rounds
 * separately exactly as the JVM does) of the JWave-Pro wavelet filter-bank hot path.  Only
 * tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py may
 * load it; the product (libjwavecuda.so) never links or calls anything in this directory.
 *
 * Parity status: PINNED against the reference's own known-answer tests (tests/test_oracle_golden.py):
 *   MODWT Haar [1..8]            src/test/java/jwave/transforms/MODWTTransformTest.java:39-71
 *   all-ones 2^(p/2) ladders     src/test/java/jwave/SteppingTest.java:37-314 (44 wavelets, N=4 and 64)
 *   Haar level-1 fixtures        src/test/resources/testdata/haar_level1_{approx,detail}_manual.txt
 *                                (src/test/java/jwave/transforms/CrossValidationTest.java:187-209)
 *   adjoint == matrix transpose  src/test/java/jwave/transforms/MODWTFFTAdjointVerificationTest.java:44-101
 * The reference is Java and no JVM exists in this image, so it cannot be executed here
 * (SURVEY.md section 0.2); a second, independent numpy restatement (oracle/np_oracle.py) must agree
 * bit-for-bit with this file (tests/test_oracle_golden.py).
 *
 * Paths below are relative to /root/reference/src/main/java/jwave/.
 */
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define JWO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * Filter construction
 * ---------------------------------------------------------------------------------------- */

/* transforms/wavelets/Wavelet.java:104-122 (_buildOrthonormalSpace):
 * waveletDeCom[i] = +scalingDeCom[L-1-i] (i even) / -scalingDeCom[L-1-i] (i odd); the two
 * reconstruction filters are copies of the decomposition filters.
 * NOTE haar/Haar1.java:52-68 sets {s1, -s0} by hand, which is the same rule for L = 2. */
JWO_API void jwo_build_orthonormal(const double* scaling, int L, double* wavelet) {
  for (int i = 0; i < L; i++)
    wavelet[i] = (i % 2 == 0) ? scaling[(L - 1) - i] : -scaling[(L - 1) - i];
}

/* transforms/MODWTTransform.java:599-606 (normalize) */
static void jwo_normalize(double* f, int L) {
  double energy = 0.0;
  for (int i = 0; i < L; i++) energy += f[i] * f[i];
  double norm = sqrt(energy);
  if (norm > 1e-12)
    for (int i = 0; i < L; i++) f[i] /= norm;
}

/* transforms/MODWTTransform.java:452-484 (initializeFilterCache): g~ = normalize(scalingDeCom)/sqrt(2),
 * h~ = normalize(waveletDeCom)/sqrt(2); two separate element-wise divisions. */
JWO_API void jwo_modwt_filters(const double* scaling, const double* wavelet, int L, double* g, double* h) {
  double* gd = (double*)malloc(sizeof(double) * (size_t)L);
  double* hd = (double*)malloc(sizeof(double) * (size_t)L);
  memcpy(gd, scaling, sizeof(double) * (size_t)L);
  memcpy(hd, wavelet, sizeof(double) * (size_t)L);
  jwo_normalize(gd, L);
  jwo_normalize(hd, L);
  double scale = sqrt(2.0);
  for (int i = 0; i < L; i++) {
    g[i] = gd[i] / scale;
    h[i] = hd[i] / scale;
  }
  free(gd);
  free(hd);
}

/* transforms/MODWTTransform.java:618-630 (upsample): level j inserts 2^(j-1)-1 zeros between taps.
 * Returns the new length (L-1)*2^(j-1)+1; `out` must hold that many doubles. */
JWO_API int jwo_upsample(const double* f, int L, int level, double* out) {
  if (level <= 1) {
    memcpy(out, f, sizeof(double) * (size_t)L);
    return L;
  }
  int gap = (1 << (level - 1)) - 1;
  int M = L + (L - 1) * gap;
  for (int i = 0; i < M; i++) out[i] = 0.0;
  for (int i = 0; i < L; i++) out[i * (gap + 1)] = f[i];
  return M;
}

static inline int64_t floormod(int64_t a, int64_t n) {
  int64_t r = a % n;
  return r < 0 ? r + n : r;
}

/* ------------------------------------------------------------------------------------------
 * Direct circular convolution -- THE PARITY ORACLE for MODWT
 * ---------------------------------------------------------------------------------------- */

/* transforms/MODWTTransform.java:677-690 (circularConvolve): sums over the zero-stuffed filter too. */
JWO_API void jwo_circular_convolve(const double* x, int N, const double* f, int M, double* out) {
  for (int n = 0; n < N; n++) {
    double sum = 0.0;
    for (int m = 0; m < M; m++) sum += x[floormod((int64_t)n - m, N)] * f[m];
    out[n] = sum;
  }
}

/* transforms/MODWTTransform.java:703-716 (circularConvolveAdjoint) */
JWO_API void jwo_circular_convolve_adjoint(const double* x, int N, const double* f, int M, double* out) {
  for (int n = 0; n < N; n++) {
    double sum = 0.0;
    for (int m = 0; m < M; m++) sum += x[floormod((int64_t)n + m, N)] * f[m];
    out[n] = sum;
  }
}

/* Same sums with the structural zeros of the upsampled filter skipped: taps f[m] at stride st.
 * Adding x*0.0 = +-0 to a running sum that started at +0.0 never changes it (finite x), so this
 * is value-identical to the dense loops above; tests/test_oracle_golden.py checks that bit-for-bit. */
static void conv_sparse(const double* x, int64_t N, const double* f, int L, int64_t st, int adjoint, double* out) {
  for (int64_t n = 0; n < N; n++) {
    double sum = 0.0;
    for (int m = 0; m < L; m++) {
      int64_t idx = adjoint ? floormod(n + m * st, N) : floormod(n - m * st, N);
      sum += x[idx] * f[m];
    }
    out[n] = sum;
  }
}

/* transforms/MODWTTransform.java:256-306 (forwardMODWT), DIRECT method.
 * out = rows W_1..W_J, V_J, each N long (row-major, (J+1)*N doubles).
 * dense != 0 walks the literal zero-stuffed filters (O(N*M)); dense == 0 skips the zeros. */
JWO_API int jwo_modwt_forward(const double* x, int N, int J, const double* g, const double* h, int L,
                              double* out, int dense) {
  if (N <= 0 || J < 1) return -1;
  double* v = (double*)malloc(sizeof(double) * (size_t)N);
  double* vn = (double*)malloc(sizeof(double) * (size_t)N);
  double* gu = NULL;
  double* hu = NULL;
  memcpy(v, x, sizeof(double) * (size_t)N);
  for (int j = 1; j <= J; j++) {
    double* w = out + (size_t)(j - 1) * (size_t)N;
    if (dense) {
      int M = (L - 1) * (1 << (j - 1)) + 1;
      gu = (double*)realloc(gu, sizeof(double) * (size_t)M);
      hu = (double*)realloc(hu, sizeof(double) * (size_t)M);
      jwo_upsample(g, L, j, gu);
      jwo_upsample(h, L, j, hu);
      jwo_circular_convolve(v, N, hu, M, w);  /* :295 */
      jwo_circular_convolve(v, N, gu, M, vn); /* :296 */
    } else {
      conv_sparse(v, N, h, L, (int64_t)1 << (j - 1), 0, w);
      conv_sparse(v, N, g, L, (int64_t)1 << (j - 1), 0, vn);
    }
    double* t = v;
    v = vn;
    vn = t;
  }
  memcpy(out + (size_t)J * (size_t)N, v, sizeof(double) * (size_t)N);
  free(v);
  free(vn);
  free(gu);
  free(hu);
  return 0;
}

/* transforms/MODWTTransform.java:337-375 (inverseMODWT), DIRECT method.
 * The two adjoint sums are formed separately and then added (:363-369). */
JWO_API int jwo_modwt_inverse(const double* coeffs, int N, int J, const double* g, const double* h, int L,
                              double* x, int dense) {
  if (N <= 0 || J < 1) return -1;
  double* v = (double*)malloc(sizeof(double) * (size_t)N);
  double* a = (double*)malloc(sizeof(double) * (size_t)N);
  double* d = (double*)malloc(sizeof(double) * (size_t)N);
  double* gu = NULL;
  double* hu = NULL;
  memcpy(v, coeffs + (size_t)J * (size_t)N, sizeof(double) * (size_t)N);
  for (int j = J; j >= 1; j--) {
    const double* w = coeffs + (size_t)(j - 1) * (size_t)N;
    if (dense) {
      int M = (L - 1) * (1 << (j - 1)) + 1;
      gu = (double*)realloc(gu, sizeof(double) * (size_t)M);
      hu = (double*)realloc(hu, sizeof(double) * (size_t)M);
      jwo_upsample(g, L, j, gu);
      jwo_upsample(h, L, j, hu);
      jwo_circular_convolve_adjoint(v, N, gu, M, a);
      jwo_circular_convolve_adjoint(w, N, hu, M, d);
    } else {
      conv_sparse(v, N, g, L, (int64_t)1 << (j - 1), 1, a);
      conv_sparse(w, N, h, L, (int64_t)1 << (j - 1), 1, d);
    }
    for (int i = 0; i < N; i++) v[i] = a[i] + d[i];
  }
  memcpy(x, v, sizeof(double) * (size_t)N);
  free(v);
  free(a);
  free(d);
  free(gu);
  free(hu);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * FFT circular convolution -- the reference's DEFAULT (AUTO) MODWT path, used only as the timed
 * CPU baseline (it is 1.6e-12 away from exact at N = 65536).
 * Power-of-two lengths run Cooley-Tukey, every other length Bluestein's chirp-z (FastFourierTransform.java:153-163).
 * ---------------------------------------------------------------------------------------- */

/* transforms/FastFourierTransform.java:172-212 (fftCooleyTukey): bit reversal, radix-2 DIT,
 * twiddle advanced by repeated complex multiplication, 1/n on the inverse.  Complex arithmetic as
 * datatypes/natives/Complex.java:260-301. */
static void fft_cooley_tukey(double* re, double* im, int n, int inverse) {
  int bits = 0;
  while ((1 << bits) < n) bits++;
  for (int k = 0; k < n; k++) {
    unsigned r = 0, v = (unsigned)k;
    for (int b = 0; b < bits; b++) {
      r = (r << 1) | (v & 1u);
      v >>= 1;
    }
    int j = (int)r;
    if (j > k) {
      double t = re[j]; re[j] = re[k]; re[k] = t;
      t = im[j]; im[j] = im[k]; im[k] = t;
    }
  }
  for (int size = 2; size <= n; size *= 2) {
    double angle = 2 * M_PI / size * (inverse ? 1 : -1);
    double wr = cos(angle), wi = sin(angle);
    int half = size / 2;
    for (int start = 0; start < n; start += size) {
      double wnr = 1.0, wni = 0.0;
      for (int k = 0; k < half; k++) {
        double ur = re[start + k], ui = im[start + k];
        double xr = re[start + k + half], xi = im[start + k + half];
        double tr = wnr * xr - wni * xi;
        double ti = wnr * xi + wni * xr;
        re[start + k] = ur + tr;
        im[start + k] = ui + ti;
        re[start + k + half] = ur - tr;
        im[start + k + half] = ui - ti;
        double nr = wnr * wr - wni * wi;
        double ni = wnr * wi + wni * wr;
        wnr = nr;
        wni = ni;
      }
    }
  }
  if (inverse) {
    double s = 1.0 / n;
    for (int i = 0; i < n; i++) {
      re[i] = re[i] * s;
      im[i] = im[i] * s;
    }
  }
}

/* transforms/FastFourierTransform.java:218-244 (fftCooleyTukeyInternal): the same butterflies, never normalised */
static void fft_cooley_tukey_internal(double* re, double* im, int n, int inverse) {
  fft_cooley_tukey(re, im, n, inverse);
  if (inverse) {   // undo the 1/n of fft_cooley_tukey: multiply back (n is a power of two, so this is exact)
    for (int i = 0; i < n; i++) {
      re[i] = re[i] * (double)n;
      im[i] = im[i] * (double)n;
    }
  }
}

/* transforms/FastFourierTransform.java:259-324 (fftBluestein): chirp z-transform for arbitrary n through two forward
 * and one inverse power-of-two FFT of length m >= 2n - 1.  Complex products as datatypes/natives/Complex.java. */
static void fft_bluestein(double* xr, double* xi, int n, int inverse) {
  int m = 1;
  while (m < 2 * n - 1) m *= 2;
  double* cr = (double*)calloc(2 * (size_t)n, sizeof(double));
  double* ci = cr + n;
  double* ar = (double*)calloc(4 * (size_t)m, sizeof(double));
  double* ai = ar + m;
  double* br = ai + m;
  double* bi = br + m;
  for (int i = 0; i < n; i++) {
    double angle = M_PI * i * i / n * (inverse ? 1 : -1);   /* ((pi * i) * i) / n, all in double as in Java */
    cr[i] = cos(angle);
    ci[i] = sin(angle);
  }
  for (int i = 0; i < n; i++) {                              /* a = x * chirp */
    ar[i] = xr[i] * cr[i] - xi[i] * ci[i];
    ai[i] = xr[i] * ci[i] + xi[i] * cr[i];
  }
  br[0] = cr[0];
  bi[0] = -ci[0];
  for (int i = 1; i < n; i++) {                              /* b = conjugate chirp, mirrored */
    br[i] = cr[i];       bi[i] = -ci[i];
    br[m - i] = cr[i];   bi[m - i] = -ci[i];
  }
  fft_cooley_tukey_internal(ar, ai, m, 0);
  fft_cooley_tukey_internal(br, bi, m, 0);
  for (int i = 0; i < m; i++) {
    double pr = ar[i] * br[i] - ai[i] * bi[i];
    double pi = ar[i] * bi[i] + ai[i] * br[i];
    ar[i] = pr;
    ai[i] = pi;
  }
  fft_cooley_tukey_internal(ar, ai, m, 1);
  double sm = 1.0 / m;
  for (int i = 0; i < m; i++) {
    ar[i] = ar[i] * sm;
    ai[i] = ai[i] * sm;
  }
  for (int i = 0; i < n; i++) {
    double rr = ar[i] * cr[i] - ai[i] * ci[i];
    double ri = ar[i] * ci[i] + ai[i] * cr[i];
    if (inverse) {
      double sn = 1.0 / n;
      rr = rr * sn;
      ri = ri * sn;
    }
    xr[i] = rr;
    xi[i] = ri;
  }
  free(cr);
  free(ar);
}

/* FastFourierTransform.java:153-163 / :130-142: Cooley-Tukey for 2^p, Bluestein otherwise */
static void fft_any(double* re, double* im, int n, int inverse) {
  if (n <= 1) return;
  if ((n & (n - 1)) == 0) fft_cooley_tukey(re, im, n, inverse);
  else fft_bluestein(re, im, n, inverse);
}

/* exported for the known-answer tests: the reference's FFT fixtures (src/test/resources/testdata/fft_*.txt, checked by
 * CrossValidationTest.java:119-154 through Transform.forward: forward unscaled, inverse 1/n) */
JWO_API void jwo_fft(double* re, double* im, int n, int inverse) { fft_any(re, im, n, inverse); }

/* transforms/MODWTTransform.java:729-741 (wrapFilterToSignalLength), :752-786 (circularConvolveFFT),
 * :798-837 (circularConvolveFFTAdjoint, conjugates the filter spectrum).  ws = 4*N doubles. */
static void conv_fft(const double* x, int N, const double* f, int M, int adjoint, double* out, double* ws) {
  double* sr = ws;
  double* si = ws + N;
  double* fr = ws + 2 * (size_t)N;
  double* fi = ws + 3 * (size_t)N;
  for (int i = 0; i < N; i++) {
    sr[i] = x[i];
    si[i] = 0.0;
    fr[i] = 0.0;
    fi[i] = 0.0;
  }
  for (int i = 0; i < M; i++) fr[i % N] += f[i];
  fft_any(sr, si, N, 0);
  fft_any(fr, fi, N, 0);
  for (int i = 0; i < N; i++) {
    double br = fr[i], bi = adjoint ? -fi[i] : fi[i];
    double pr = sr[i] * br - si[i] * bi;
    double pi = sr[i] * bi + si[i] * br;
    sr[i] = pr;
    si[i] = pi;
  }
  fft_any(sr, si, N, 1);
  for (int i = 0; i < N; i++) out[i] = sr[i];
}

JWO_API int jwo_modwt_forward_fft(const double* x, int N, int J, const double* g, const double* h, int L, double* out) {
  if (N <= 0 || J < 1) return -1;
  size_t maxM = (size_t)(L - 1) * ((size_t)1 << (J - 1)) + 1;
  double* v = (double*)malloc(sizeof(double) * (size_t)N);
  double* vn = (double*)malloc(sizeof(double) * (size_t)N);
  double* ws = (double*)malloc(sizeof(double) * 4 * (size_t)N);
  double* gu = (double*)malloc(sizeof(double) * maxM);
  double* hu = (double*)malloc(sizeof(double) * maxM);
  memcpy(v, x, sizeof(double) * (size_t)N);
  for (int j = 1; j <= J; j++) {
    int M = jwo_upsample(g, L, j, gu);
    jwo_upsample(h, L, j, hu);
    conv_fft(v, N, hu, M, 0, out + (size_t)(j - 1) * (size_t)N, ws);
    conv_fft(v, N, gu, M, 0, vn, ws);
    double* t = v;
    v = vn;
    vn = t;
  }
  memcpy(out + (size_t)J * (size_t)N, v, sizeof(double) * (size_t)N);
  free(v); free(vn); free(ws); free(gu); free(hu);
  return 0;
}

JWO_API int jwo_modwt_inverse_fft(const double* coeffs, int N, int J, const double* g, const double* h, int L, double* x) {
  if (N <= 0 || J < 1) return -1;
  size_t maxM = (size_t)(L - 1) * ((size_t)1 << (J - 1)) + 1;
  double* v = (double*)malloc(sizeof(double) * (size_t)N);
  double* a = (double*)malloc(sizeof(double) * (size_t)N);
  double* d = (double*)malloc(sizeof(double) * (size_t)N);
  double* ws = (double*)malloc(sizeof(double) * 4 * (size_t)N);
  double* gu = (double*)malloc(sizeof(double) * maxM);
  double* hu = (double*)malloc(sizeof(double) * maxM);
  memcpy(v, coeffs + (size_t)J * (size_t)N, sizeof(double) * (size_t)N);
  for (int j = J; j >= 1; j--) {
    int M = jwo_upsample(g, L, j, gu);
    jwo_upsample(h, L, j, hu);
    conv_fft(v, N, gu, M, 1, a, ws);
    conv_fft(coeffs + (size_t)(j - 1) * (size_t)N, N, hu, M, 1, d, ws);
    for (int i = 0; i < N; i++) v[i] = a[i] + d[i];
  }
  memcpy(x, v, sizeof(double) * (size_t)N);
  free(v); free(a); free(d); free(ws); free(gu); free(hu);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * One decimated analysis / synthesis step (FWT and WPT inner loop)
 * ---------------------------------------------------------------------------------------- */

/* transforms/wavelets/Wavelet.java:236-260 (forward): out[0..h/2) = low pass, out[h/2..h) = high
 * pass, periodic wrap by repeated subtraction (legal for L > h). `x` is not modified. */
JWO_API void jwo_wavelet_forward(const double* x, int len, const double* s, const double* w, int L, double* out) {
  int h = len >> 1;
  for (int i = 0; i < h; i++) {
    out[i] = out[i + h] = 0.;
    for (int j = 0; j < L; j++) {
      int k = (i << 1) + j;
      while (k >= len) k -= len;
      out[i] += x[k] * s[j];
      out[i + h] += x[k] * w[j];
    }
  }
}

/* transforms/wavelets/Wavelet.java:277-303 (reverse): scatter-add, i outer, j inner. */
JWO_API void jwo_wavelet_reverse(const double* c, int len, const double* sr, const double* wr, int L, double* out) {
  for (int i = 0; i < len; i++) out[i] = 0.;
  int h = len >> 1;
  for (int i = 0; i < h; i++) {
    for (int j = 0; j < L; j++) {
      int k = (i << 1) + j;
      while (k >= len) k -= len;
      out[k] += (c[i] * sr[j]) + (c[i + h] * wr[j]);
    }
  }
}

/* transforms/FastWaveletTransform.java:71-101 (forward with level); N must be 2^p, 0<=level<=p. */
JWO_API int jwo_fwt_forward(const double* x, int N, int level, const double* s, const double* w, int L, double* out) {
  double* tmp = (double*)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
  memcpy(out, x, sizeof(double) * (size_t)N);
  int l = 0, h = N;
  while (h >= 2 && l < level) {
    jwo_wavelet_forward(out, h, s, w, L, tmp);
    memcpy(out, tmp, sizeof(double) * (size_t)h);
    h >>= 1;
    l++;
  }
  free(tmp);
  return 0;
}

/* transforms/FastWaveletTransform.java:119-153 (reverse with level) */
JWO_API int jwo_fwt_reverse(const double* c, int N, int level, const double* sr, const double* wr, int L, double* out) {
  double* tmp = (double*)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
  memcpy(out, c, sizeof(double) * (size_t)N);
  int steps = 0;
  while ((1 << steps) < N) steps++;
  int64_t h = 2;
  for (int l = level; l < steps; l++) h <<= 1;
  while (h <= N && h >= 2) {
    jwo_wavelet_reverse(out, (int)h, sr, wr, L, tmp);
    memcpy(out, tmp, sizeof(double) * (size_t)h);
    h <<= 1;
  }
  free(tmp);
  return 0;
}

/* transforms/WaveletPacketTransform.java:73-124 (forward with level): every aligned block of length h
 * is replaced by its own [lo|hi]. */
JWO_API int jwo_wpt_forward(const double* x, int N, int level, const double* s, const double* w, int L, double* out) {
  double* ib = (double*)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
  double* ob = (double*)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
  memcpy(out, x, sizeof(double) * (size_t)N);
  int l = 0, h = N;
  while (h >= 2 && l < level) {
    int g = N / h;
    for (int p = 0; p < g; p++) {
      memcpy(ib, out + (size_t)p * h, sizeof(double) * (size_t)h);
      jwo_wavelet_forward(ib, h, s, w, L, ob);
      memcpy(out + (size_t)p * h, ob, sizeof(double) * (size_t)h);
    }
    h >>= 1;
    l++;
  }
  free(ib);
  free(ob);
  return 0;
}

/* transforms/WaveletPacketTransform.java:141-191 (reverse with level) */
JWO_API int jwo_wpt_reverse(const double* c, int N, int level, const double* sr, const double* wr, int L, double* out) {
  double* ib = (double*)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
  double* ob = (double*)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
  memcpy(out, c, sizeof(double) * (size_t)N);
  int steps = 0;
  while ((1 << steps) < N) steps++;
  int64_t h = 2;
  for (int l = level; l < steps; l++) h <<= 1;
  while (h <= N && h >= 2) {
    int g = N / (int)h;
    for (int p = 0; p < g; p++) {
      memcpy(ib, out + (size_t)p * h, sizeof(double) * (size_t)h);
      jwo_wavelet_reverse(ib, (int)h, sr, wr, L, ob);
      memcpy(out + (size_t)p * h, ob, sizeof(double) * (size_t)h);
    }
    h <<= 1;
  }
  free(ib);
  free(ob);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Batch drivers: signals spread over host threads.  This is how the CPU baseline is timed
 * (bench.py cpu_baseline / --impl reference): the reference itself is single-threaded per signal
 * (ParallelWaveletPacketTransform.java:155-158,197-233 only forks inside one signal when a level has
 * more than 16 packets), so "one signal per core" is the strongest honest use of the host cores.
 * op: 0 modwt fwd direct, 1 modwt inv direct, 2 modwt fwd FFT, 3 modwt inv FFT,
 *     4 fwt fwd, 5 fwt rev, 6 wpt fwd, 7 wpt rev.
 * For MODWT f0 = g~, f1 = h~; for FWT/WPT f0 = scaling, f1 = wavelet filter of that direction.
 * in/out strides per signal: MODWT fwd N -> (J+1)N, inv (J+1)N -> N, others N -> N.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int op, N, level, L, tid, nthreads;
  int64_t batch;
  const double *in, *f0, *f1;
  double* out;
  int rc;
} jwo_job;

static void* jwo_worker(void* p) {
  jwo_job* j = (jwo_job*)p;
  size_t N = (size_t)j->N, JN = (size_t)(j->level + 1) * N;
  j->rc = 0;
  for (int64_t b = j->tid; b < j->batch; b += j->nthreads) {
    int rc = 0;
    switch (j->op) {
      case 0: rc = jwo_modwt_forward(j->in + b * N, j->N, j->level, j->f0, j->f1, j->L, j->out + b * JN, 0); break;
      case 1: rc = jwo_modwt_inverse(j->in + b * JN, j->N, j->level, j->f0, j->f1, j->L, j->out + b * N, 0); break;
      case 2: rc = jwo_modwt_forward_fft(j->in + b * N, j->N, j->level, j->f0, j->f1, j->L, j->out + b * JN); break;
      case 3: rc = jwo_modwt_inverse_fft(j->in + b * JN, j->N, j->level, j->f0, j->f1, j->L, j->out + b * N); break;
      case 4: rc = jwo_fwt_forward(j->in + b * N, j->N, j->level, j->f0, j->f1, j->L, j->out + b * N); break;
      case 5: rc = jwo_fwt_reverse(j->in + b * N, j->N, j->level, j->f0, j->f1, j->L, j->out + b * N); break;
      case 6: rc = jwo_wpt_forward(j->in + b * N, j->N, j->level, j->f0, j->f1, j->L, j->out + b * N); break;
      case 7: rc = jwo_wpt_reverse(j->in + b * N, j->N, j->level, j->f0, j->f1, j->L, j->out + b * N); break;
      default: rc = -2;
    }
    if (rc) j->rc = rc;
  }
  return NULL;
}

JWO_API int jwo_batch(int op, const double* in, double* out, int64_t batch, int N, int level,
                      const double* f0, const double* f1, int L, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 1024) nthreads = 1024;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
  jwo_job* jobs = (jwo_job*)malloc(sizeof(jwo_job) * (size_t)nthreads);
  for (int t = 0; t < nthreads; t++) {
    jwo_job j = {op, N, level, L, t, nthreads, batch, in, f0, f1, out, 0};
    jobs[t] = j;
    if (t > 0) pthread_create(&th[t], NULL, jwo_worker, &jobs[t]);
  }
  jwo_worker(&jobs[0]);
  int rc = jobs[0].rc;
  for (int t = 1; t < nthreads; t++) {
    pthread_join(th[t], NULL);
    if (jobs[t].rc) rc = jobs[t].rc;
  }
  free(th);
  free(jobs);
  return rc;
}

/* ------------------------------------------------------------------------------------------
 * transforms/ParallelWaveletPacketTransform.java -- the CPU baseline north_star names for the WPT.
 * Same arithmetic as WaveletPacketTransform (its own tests, ParallelWPTTest.java:154-178, require equality with the
 * sequential transform to 1e-10); what differs is the SCHEDULE, restated here:
 *   :79-110 / :113-147   level loop on ONE signal: h = N, N/2, ... (reverse: h = 2^(steps-level+1) ... N), packets = N / h
 *   :155-158             a level runs in parallel only when packetSize >= 64 && packets >= 8, else sequentially (:163-184)
 *   :197-233             WPTLevelTask: the packet range is halved recursively until a task holds <= 16 packets; the
 *                        leaves run on a ForkJoinPool (work stealing), invoke() joins before the next level starts
 * The ForkJoinPool is a persistent pool of `nthreads` workers; here: a persistent pthread pool, the leaves of a level in
 * a shared array handed out by an atomic counter (the stand-in for work stealing), one barrier pair per parallel level
 * (fork / join).  Signals of a batch go through one after the other, exactly as a caller looping over
 * ParallelWaveletPacketTransform.forward would run them.
 * ---------------------------------------------------------------------------------------- */
/* fork / join of a level: a sense-reversing spin barrier (ForkJoinPool workers spin before they park; a futex-based
 * pthread_barrier costs tens of microseconds per level on a 16-thread VM, several times the work of a level) */
typedef struct { int count, gen, n; } jwo_spin_barrier;
static void jwo_spin_init(jwo_spin_barrier* b, int n) { b->count = 0; b->gen = 0; b->n = n; }
static void jwo_spin_wait(jwo_spin_barrier* b) {
  const int gen = __atomic_load_n(&b->gen, __ATOMIC_ACQUIRE);
  if (__atomic_add_fetch(&b->count, 1, __ATOMIC_ACQ_REL) == b->n) {
    __atomic_store_n(&b->count, 0, __ATOMIC_RELAXED);
    __atomic_store_n(&b->gen, gen + 1, __ATOMIC_RELEASE);
  } else {
    int spins = 0;
    while (__atomic_load_n(&b->gen, __ATOMIC_ACQUIRE) == gen)
      if (++spins > 20000) { sched_yield(); spins = 0; }
  }
}

typedef struct {
  jwo_spin_barrier start, done;
  int nthreads, stop;
  /* the level in flight */
  double* data;
  int h, L, forward, nleaves;
  const double *f0, *f1;
  int (*leaves)[2];
  int next;   /* atomic cursor into leaves */
} jwo_pool;

static void jwo_pwpt_packet(double* data, int h, int p, const double* f0, const double* f1, int L, int forward,
                            double* ib, double* ob) {
  /* :244-262 processPacket: copy the packet out, one Wavelet step, copy it back */
  memcpy(ib, data + (size_t)p * h, sizeof(double) * (size_t)h);
  if (forward) jwo_wavelet_forward(ib, h, f0, f1, L, ob);
  else jwo_wavelet_reverse(ib, h, f0, f1, L, ob);
  memcpy(data + (size_t)p * h, ob, sizeof(double) * (size_t)h);
}

static void jwo_pwpt_drain(jwo_pool* pl, double* ib, double* ob) {
  for (;;) {
    int i = __atomic_fetch_add(&pl->next, 1, __ATOMIC_RELAXED);
    if (i >= pl->nleaves) break;
    for (int p = pl->leaves[i][0]; p < pl->leaves[i][1]; p++)
      jwo_pwpt_packet(pl->data, pl->h, p, pl->f0, pl->f1, pl->L, pl->forward, ib, ob);
  }
}

typedef struct { jwo_pool* pl; int N; } jwo_pwpt_arg;

static void* jwo_pwpt_worker(void* a) {
  jwo_pwpt_arg* arg = (jwo_pwpt_arg*)a;
  jwo_pool* pl = arg->pl;
  double* ib = (double*)malloc(sizeof(double) * (size_t)arg->N);
  double* ob = (double*)malloc(sizeof(double) * (size_t)arg->N);
  for (;;) {
    jwo_spin_wait(&pl->start);
    if (pl->stop) break;
    jwo_pwpt_drain(pl, ib, ob);
    jwo_spin_wait(&pl->done);
  }
  free(ib);
  free(ob);
  return NULL;
}

/* :215-232 WPTLevelTask.compute: halve [a, b) until <= 16 packets */
static void jwo_pwpt_split(int a, int b, int (*leaves)[2], int* n) {
  if (b - a <= 16) { leaves[*n][0] = a; leaves[*n][1] = b; (*n)++; return; }
  int mid = a + (b - a) / 2;
  jwo_pwpt_split(a, mid, leaves, n);
  jwo_pwpt_split(mid, b, leaves, n);
}

JWO_API int jwo_parallel_wpt(const double* in, double* out, int64_t batch, int N, int level, const double* f0,
                             const double* f1, int L, int nthreads, int reverse) {
  if (N < 1 || (N & (N - 1)) || level < 0) return -1;
  int steps = 0;
  while ((1 << steps) < N) steps++;
  if (level > steps) return -1;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 1024) nthreads = 1024;
  jwo_pool pl;
  memset(&pl, 0, sizeof(pl));
  pl.nthreads = nthreads;
  pl.f0 = f0; pl.f1 = f1; pl.L = L; pl.forward = !reverse;
  pl.leaves = (int (*)[2])malloc(sizeof(int[2]) * (size_t)(N / 2 + 1));
  jwo_spin_init(&pl.start, nthreads);
  jwo_spin_init(&pl.done, nthreads);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
  jwo_pwpt_arg arg = {&pl, N};
  for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, jwo_pwpt_worker, &arg);
  double* ib = (double*)malloc(sizeof(double) * (size_t)N);
  double* ob = (double*)malloc(sizeof(double) * (size_t)N);
  for (int64_t b = 0; b < batch; b++) {
    double* d = out + (size_t)b * N;
    memcpy(d, in + (size_t)b * N, sizeof(double) * (size_t)N);   /* Arrays.copyOf :88 / :122 */
    int64_t h;
    int l = 0;
    if (!reverse) h = N;
    else { h = 2; for (int q = level; q < steps; q++) h <<= 1; }
    while (!reverse ? (h >= 2 && l < level) : (h <= N && h >= 2)) {
      int packets = (int)(N / h);
      if (h >= 64 && packets >= 8) {          /* shouldUseParallel :155-158 */
        pl.data = d; pl.h = (int)h; pl.nleaves = 0; pl.next = 0;
        jwo_pwpt_split(0, packets, pl.leaves, &pl.nleaves);
        if (nthreads > 1) jwo_spin_wait(&pl.start);   /* fork */
        jwo_pwpt_drain(&pl, ib, ob);
        if (nthreads > 1) jwo_spin_wait(&pl.done);    /* join (invoke returns) */
      } else {
        for (int p = 0; p < packets; p++) jwo_pwpt_packet(d, (int)h, p, f0, f1, L, !reverse, ib, ob);   /* :163-184 */
      }
      if (!reverse) { h >>= 1; l++; } else h <<= 1;
    }
  }
  pl.stop = 1;
  if (nthreads > 1) jwo_spin_wait(&pl.start);
  for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th); free(ib); free(ob); free(pl.leaves);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * compressions/CompressorMagnitude.java:78-90 (compress(double[]): magnitude = left-to-right sum of |c| / length;
 * the double[][] :97-113 and double[][][] :120-139 overloads run the same sum in row-major order, which is this loop
 * on the flattened array) + compressions/Compressor.java:97-112 (keep c where |c| >= magnitude * threshold, else 0).
 * Returns the magnitude.
 * ---------------------------------------------------------------------------------------- */
JWO_API double jwo_compress_magnitude(const double* in, int64_t count, double threshold, double* out) {
  double magnitude = 0.0;
  for (int64_t i = 0; i < count; i++) magnitude += fabs(in[i]);
  magnitude /= (double)count;
  for (int64_t i = 0; i < count; i++) out[i] = (fabs(in[i]) >= magnitude * threshold) ? in[i] : 0.0;
  return magnitude;
}
