// C++ host-mirror test (reads like the reference's JUnit tests): KATs from MODWTTransformTest.java:39-71,
// SteppingTest.java:37-179, error behaviour from MODWT1DInterfaceTest.java:108-137.  Needs a GPU; run by
// tests/test_cpp_mirror.py (gpu-marked).  Exit code 0 = all checks passed.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../include/jwave_cuda.hpp"
#include "../../include/jwc_filters_generated.h"

using namespace jwave;

static int fails = 0;
#define CHECK(c) do { if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); fails++; } } while (0)

static Wavelet make(const char* cls) {
  for (int i = 0; i < jwc_num_wavelet_tables; i++)
    if (!strcmp(jwc_wavelet_tables[i].cls, cls))
      return Wavelet(jwc_wavelet_tables[i].name, std::vector<double>(jwc_wavelet_tables[i].scaling_decom, jwc_wavelet_tables[i].scaling_decom + jwc_wavelet_tables[i].L));
  throw std::runtime_error("no such wavelet");
}

int main() {
  auto ctx = std::make_shared<Context>();
  {  // MODWT Haar level 1 of [1..8]
    CudaMODWTTransform t(make("Haar1"), ctx);
    std::vector<double> x = {1, 2, 3, 4, 5, 6, 7, 8};
    auto c = t.forwardMODWT(x, 1);
    const double d1[8] = {-3.5, .5, .5, .5, .5, .5, .5, .5}, a1[8] = {4.5, 1.5, 2.5, 3.5, 4.5, 5.5, 6.5, 7.5};
    for (int i = 0; i < 8; i++) { CHECK(std::fabs(c[0][i] - d1[i]) < 1e-9); CHECK(std::fabs(c[1][i] - a1[i]) < 1e-9); }
    auto xr = t.inverseMODWT(c);
    for (int i = 0; i < 8; i++) CHECK(std::fabs(xr[i] - x[i]) < 1e-9);
    auto flat = t.forward(x, 2);
    CHECK(flat.size() == 24);
    auto back = t.reverse(flat, 2);
    for (int i = 0; i < 8; i++) CHECK(std::fabs(back[i] - x[i]) < 1e-9);
    auto back2 = t.reverse(flat);   // level search: 24 = 8 * 3
    for (int i = 0; i < 8; i++) CHECK(std::fabs(back2[i] - x[i]) < 1e-9);
    bool thrown = false;
    try { t.forwardMODWT(x, 0); } catch (const std::invalid_argument& e) { thrown = strstr(e.what(), "at least 1") != nullptr; }
    CHECK(thrown);
    thrown = false;
    try { t.forwardMODWT(x, 4); } catch (const std::invalid_argument& e) { thrown = strstr(e.what(), "exceeds theoretical limit 3") != nullptr; }
    CHECK(thrown);
    thrown = false;
    try { t.forward(std::vector<double>(7, 1.0), 2); } catch (const JWaveFailure& e) { thrown = strstr(e.what(), "2^p") != nullptr; }
    CHECK(thrown);
    CHECK(t.forwardMODWT({}, 3).size() == 4);
  }
  // all-ones ladders for every in-scope wavelet, FWT and WPT, N = 64
  for (int wi = 0; wi < jwc_num_wavelet_tables; wi++) {
    Wavelet w(jwc_wavelet_tables[wi].name, std::vector<double>(jwc_wavelet_tables[wi].scaling_decom, jwc_wavelet_tables[wi].scaling_decom + jwc_wavelet_tables[wi].L));
    CudaFastWaveletTransform fwt(w, ctx);
    CudaWaveletPacketTransform wpt(w, ctx);
    std::vector<double> ones(64, 1.0);
    for (int p = 0; p <= 6; p++) {
      auto a = fwt.forward(ones, p), b = wpt.forward(ones, p);
      for (int i = 0; i < 64; i++) {
        const double e = (i < (64 >> p)) ? std::pow(2.0, p / 2.0) : 0.0;
        CHECK(std::fabs(a[i] - e) < 1e-8);
        CHECK(std::fabs(b[i] - e) < 1e-8);
      }
      auto ra = fwt.reverse(a, p), rb = wpt.reverse(b, p);
      for (int i = 0; i < 64; i++) { CHECK(std::fabs(ra[i] - 1.0) < 1e-8); CHECK(std::fabs(rb[i] - 1.0) < 1e-8); }
    }
  }
  {
    CudaFastWaveletTransform fwt(make("Daubechies4"), ctx);
    CHECK(fwt.getName() == "Fast Wavelet Transform");
    bool thrown = false;
    try { fwt.forward(std::vector<double>(17, 1.0), 2); } catch (const JWaveFailure& e) { thrown = strstr(e.what(), "2^p") != nullptr; }
    CHECK(thrown);
    thrown = false;
    try { fwt.forward(std::vector<double>(8, 1.0), 10); } catch (const JWaveFailure& e) { thrown = strstr(e.what(), "out of range") != nullptr; }
    CHECK(thrown);
    auto m = fwt.decompose(std::vector<double>(16, 1.0));
    CHECK(m.size() == 5);
    auto r = fwt.recompose(m, 3);
    for (double v : r) CHECK(std::fabs(v - 1.0) < 1e-8);
  }
  {  // 2-D overloads (BasicTransform.java:336-474): constant 8 x 16 matrix -> sqrt(128) in the corner; round trip
    CudaFastWaveletTransform fwt(make("Daubechies4"), ctx);
    CudaWaveletPacketTransform wpt(make("Symlet8"), ctx);
    Matrix ones(8, std::vector<double>(16, 1.0));
    Matrix h = fwt.forward(ones);
    for (int i = 0; i < 8; i++)
      for (int j = 0; j < 16; j++) CHECK(std::fabs(h[i][j] - ((i == 0 && j == 0) ? std::sqrt(128.0) : 0.0)) < 1e-8);
    Matrix ramp(8, std::vector<double>(16));
    for (int i = 0; i < 8; i++)
      for (int j = 0; j < 16; j++) ramp[i][j] = std::sin(0.37 * i + 0.11 * j * j);
    Matrix back = fwt.reverse(fwt.forward(ramp, 2, 3), 2, 3);
    Matrix back2 = wpt.reverse(wpt.forward(ramp));
    for (int i = 0; i < 8; i++)
      for (int j = 0; j < 16; j++) { CHECK(std::fabs(back[i][j] - ramp[i][j]) < 1e-10); CHECK(std::fabs(back2[i][j] - ramp[i][j]) < 1e-10); }
    bool thrown = false;
    try { fwt.forward(Matrix(6, std::vector<double>(16, 1.0))); } catch (const JWaveFailure& e) { thrown = true; }
    CHECK(thrown);
    thrown = false;
    try { fwt.forward(ramp, 4, 3); } catch (const JWaveFailure& e) { thrown = strstr(e.what(), "out of range") != nullptr; }
    CHECK(thrown);
  }
  {  // 3-D overloads (BasicTransform.java:487-640): constant 8 x 8 x 8 cube -> sqrt(512) in the corner; round trips
    CudaFastWaveletTransform fwt(make("Daubechies4"), ctx);
    CudaWaveletPacketTransform wpt(make("Haar1"), ctx);
    Space ones(8, Matrix(8, std::vector<double>(8, 1.0)));
    Space h = fwt.forward(ones);
    for (int i = 0; i < 8; i++)
      for (int j = 0; j < 8; j++)
        for (int k = 0; k < 8; k++) CHECK(std::fabs(h[i][j][k] - ((i + j + k == 0) ? std::sqrt(512.0) : 0.0)) < 1e-8);
    Space wave(4, Matrix(8, std::vector<double>(16)));
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 8; j++)
        for (int k = 0; k < 16; k++) wave[i][j][k] = std::cos(0.3 * i * i + 0.21 * j - 0.07 * k * k);
    Space back = fwt.reverse(fwt.forward(wave, 3, 4, 2), 3, 4, 2);   // lvlP -> the 8-axis, lvlQ -> the 16-axis, lvlR -> the 4-axis
    Space back2 = wpt.reverse(wpt.forward(wave, 1, 2, 1), 1, 2, 1);
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 8; j++)
        for (int k = 0; k < 16; k++) { CHECK(std::fabs(back[i][j][k] - wave[i][j][k]) < 1e-10); CHECK(std::fabs(back2[i][j][k] - wave[i][j][k]) < 1e-10); }
    bool thrown = false;
    try { fwt.forward(wave); } catch (const JWaveFailure& e) { thrown = strstr(e.what(), "out of range") != nullptr; }   // lvlQ = log2(8) = 3 fits the 16-axis, lvlR = log2(16) = 4 > log2(4): as in the reference
    CHECK(thrown);
  }
  {  // Complex[] overloads (BasicTransform.java:257-320) and the MODWT configuration accessors (MODWTTransform.java:191-213)
    CudaFastWaveletTransform fwt(make("Daubechies4"), ctx);
    std::vector<std::complex<double>> z(64);
    for (size_t i = 0; i < z.size(); i++) z[i] = {std::sin(0.3 * (double)i), std::cos(0.11 * (double)(i * i))};
    std::vector<double> bulk(128);
    for (size_t i = 0; i < z.size(); i++) { bulk[2 * i] = z[i].real(); bulk[2 * i + 1] = z[i].imag(); }
    const auto hz = fwt.forward(z);
    const auto hb = fwt.forward(bulk);
    for (size_t i = 0; i < z.size(); i++) { CHECK(hz[i].real() == hb[2 * i]); CHECK(hz[i].imag() == hb[2 * i + 1]); }
    const auto back = fwt.reverse(hz);
    for (size_t i = 0; i < z.size(); i++) CHECK(std::abs(back[i] - z[i]) < 1e-10);
    CudaMODWTTransform m(make("Haar1"), 1024, ctx);
    CHECK(m.getConvolutionMethod() == CudaMODWTTransform::ConvolutionMethod::AUTO);
    m.setConvolutionMethod(CudaMODWTTransform::ConvolutionMethod::DIRECT);
    CHECK(m.getConvolutionMethod() == CudaMODWTTransform::ConvolutionMethod::DIRECT);
  }
  printf(fails ? "%d checks FAILED\n" : "host mirror ok\n", fails);
  return fails ? 1 : 0;
}
