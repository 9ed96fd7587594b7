// jwave_cuda.hpp -- header-only C++17 host mirror of the reference's transform classes for the GPU path.
//
// The reference is Java (no JDK in this image), so besides the Java sources under java/ the drop-in classes also
// exist in C++ above the same C ABI (jwavecuda.h): same class and method names, argument meaning, validation order and
// error behaviour as
//   transforms/BasicTransform.java:99-157,671-697      isBinary / calcExponent / 1-D API
//   transforms/WaveletTransform.java:77-182            full-depth defaults, decompose / recompose
//   transforms/FastWaveletTransform.java:71-153        CudaFastWaveletTransform
//   transforms/WaveletPacketTransform.java:73-191      CudaWaveletPacketTransform
//   transforms/MODWTTransform.java:256-443,854-912     CudaMODWTTransform
// (paths relative to /root/reference/src/main/java/jwave/).  JWaveFailure mirrors the checked exception of the same
// name, std::invalid_argument stands in for java.lang.IllegalArgumentException.  No CPU path: without a CUDA device the
// Context constructor throws.
#pragma once
#include <cmath>
#include <complex>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "jwavecuda.h"

namespace jwave {

struct JWaveException : std::runtime_error { using std::runtime_error::runtime_error; };
struct JWaveFailure : JWaveException { using JWaveException::JWaveException; };

// transforms/wavelets/Wavelet.java: coefficient carrier; orthonormal construction rule of :104-122
class Wavelet {
 public:
  Wavelet(std::string name, std::vector<double> scalingDeCom) : name_(std::move(name)), s_(std::move(scalingDeCom)) {
    const size_t L = s_.size();
    w_.resize(L);
    for (size_t i = 0; i < L; i++) w_[i] = (i % 2 == 0) ? s_[L - 1 - i] : -s_[L - 1 - i];
  }
  const std::string& getName() const { return name_; }
  int getMotherWavelength() const { return (int)s_.size(); }
  int getTransformWavelength() const { return 2; }
  std::vector<double> getScalingDeComposition() const { return s_; }
  std::vector<double> getWaveletDeComposition() const { return w_; }
  std::vector<double> getScalingReConstruction() const { return s_; }
  std::vector<double> getWaveletReConstruction() const { return w_; }

 private:
  std::string name_;
  std::vector<double> s_, w_;
};

class Context {
 public:
  explicit Context(const std::vector<int>& devices = {}) {
    ctx_ = jwc_create(devices.empty() ? nullptr : devices.data(), (int)devices.size());
    if (!ctx_) throw std::runtime_error(std::string("jwc_create: ") + jwc_last_error());
  }
  ~Context() { jwc_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  jwc_ctx* handle() const { return ctx_; }

 private:
  jwc_ctx* ctx_;
};

class BasicTransform {
 public:
  virtual ~BasicTransform() = default;
  static bool isBinary(int64_t n) { return n > 0 && (n & (n - 1)) == 0; }   // tools/MathToolKit.java:185
  static int calcExponent(int64_t n) {                                     // BasicTransform.java:687-697
    if (!isBinary(n)) throw JWaveFailure("BasicTransform#calcExponent - given number is not binary: 2^p | pEN .. = 1, 2, 4, 8, 16, 32, .. ");
    int p = 0;
    while (((int64_t)1 << p) < n) p++;
    return p;
  }
  const std::string& getName() const { return name_; }

 protected:
  std::string name_;
};

class WaveletTransform : public BasicTransform {
 public:
  WaveletTransform(Wavelet w, std::shared_ptr<Context> ctx) : wavelet_(std::move(w)), ctx_(std::move(ctx)) {
    if (!ctx_) ctx_ = std::make_shared<Context>();
  }
  const Wavelet& getWavelet() const { return wavelet_; }
  virtual std::vector<double> forward(const std::vector<double>& arrTime, int level) = 0;
  virtual std::vector<double> reverse(const std::vector<double>& arrHilb, int level) = 0;
  // WaveletTransform.java:77-112
  virtual std::vector<double> forward(const std::vector<double>& arrTime) {
    if (!isBinary((int64_t)arrTime.size())) throw JWaveFailure("WaveletTransform#forward - given array length is not 2^p | p E N ... = 1, 2, 4, 8, 16, 32, .. please use the Ancient Egyptian Decomposition for any other array length!");
    return forward(arrTime, calcExponent((int64_t)arrTime.size()));
  }
  virtual std::vector<double> reverse(const std::vector<double>& arrHilb) {
    if (!isBinary((int64_t)arrHilb.size())) throw JWaveFailure("WaveletTransform#reverse - given array length is not 2^p | p E N ... = 1, 2, 4, 8, 16, 32, .. please use the Ancient Egyptian Decomposition for any other array length!");
    return reverse(arrHilb, calcExponent((int64_t)arrHilb.size()));
  }
  // BasicTransform.java:257-320 forward / reverse(Complex[]): N complex numbers as ONE real array of length 2N, real and
  // imaginary parts interleaved, through the 1-D transform at full depth
  std::vector<std::complex<double>> forward(const std::vector<std::complex<double>>& arrTime) {
    std::vector<double> bulk(2 * arrTime.size());
    for (size_t i = 0; i < arrTime.size(); i++) { bulk[2 * i] = arrTime[i].real(); bulk[2 * i + 1] = arrTime[i].imag(); }
    const std::vector<double> h = forward(bulk);
    std::vector<std::complex<double>> out(arrTime.size());
    for (size_t i = 0; i < out.size(); i++) out[i] = {h[2 * i], h[2 * i + 1]};
    return out;
  }
  std::vector<std::complex<double>> reverse(const std::vector<std::complex<double>>& arrHilb) {
    std::vector<double> bulk(2 * arrHilb.size());
    for (size_t i = 0; i < arrHilb.size(); i++) { bulk[2 * i] = arrHilb[i].real(); bulk[2 * i + 1] = arrHilb[i].imag(); }
    const std::vector<double> t = reverse(bulk);
    std::vector<std::complex<double>> out(arrHilb.size());
    for (size_t i = 0; i < out.size(); i++) out[i] = {t[2 * i], t[2 * i + 1]};
    return out;
  }
  // WaveletTransform.java:136-182
  std::vector<std::vector<double>> decompose(const std::vector<double>& arrTime) {
    const int levels = calcExponent((int64_t)arrTime.size());
    std::vector<std::vector<double>> m;
    for (int p = 0; p <= levels; p++) m.push_back(forward(arrTime, p));
    return m;
  }
  std::vector<double> recompose(const std::vector<std::vector<double>>& m, int level) {
    if (level < 0 || level >= (int)m.size()) throw JWaveFailure("WaveletTransform#recompose - given level is out of range");
    return reverse(m[(size_t)level], level);
  }

 protected:
  using Fn = int (*)(jwc_ctx*, const double*, double*, int64_t, int64_t, int, const double*, const double*, int, unsigned);
  void call(Fn fn, const char* what, const double* in, double* out, int64_t batch, int64_t n, int levels,
            const std::vector<double>& f0, const std::vector<double>& f1, unsigned flags = 0) const {
    const int rc = fn(ctx_->handle(), in, out, batch, n, levels, f0.data(), f1.data(), (int)f0.size(), flags);
    if (rc != JWC_OK) throw std::runtime_error(std::string(what) + " failed: " + jwc_last_error());
  }
  using Fn2 = int (*)(jwc_ctx*, const double*, double*, int64_t, int64_t, int64_t, int, int, const double*, const double*, int, unsigned);
  void call2d(Fn2 fn, const char* what, const double* in, double* out, int64_t batch, int64_t rows, int64_t cols, int lvlM,
              int lvlN, const std::vector<double>& f0, const std::vector<double>& f1, unsigned flags = 0) const {
    const int rc = fn(ctx_->handle(), in, out, batch, rows, cols, lvlM, lvlN, f0.data(), f1.data(), (int)f0.size(), flags);
    if (rc != JWC_OK) throw std::runtime_error(std::string(what) + " failed: " + jwc_last_error());
  }
  using Fn3 = int (*)(jwc_ctx*, const double*, double*, int64_t, int64_t, int64_t, int64_t, int, int, int, const double*, const double*, int, unsigned);
  void call3d(Fn3 fn, const char* what, const double* in, double* out, int64_t batch, int64_t p, int64_t q, int64_t r, int lvlP,
              int lvlQ, int lvlR, const std::vector<double>& f0, const std::vector<double>& f1, unsigned flags = 0) const {
    const int rc = fn(ctx_->handle(), in, out, batch, p, q, r, lvlP, lvlQ, lvlR, f0.data(), f1.data(), (int)f0.size(), flags);
    if (rc != JWC_OK) throw std::runtime_error(std::string(what) + " failed: " + jwc_last_error());
  }
  Wavelet wavelet_;
  std::shared_ptr<Context> ctx_;
};

using Matrix = std::vector<std::vector<double>>;
using Space = std::vector<Matrix>;   // double[][][] of the reference's 3-D overloads

namespace detail {
inline std::vector<double> flatten(const Matrix& m) {
  std::vector<double> flat;
  if (m.empty()) return flat;
  flat.reserve(m.size() * m[0].size());
  for (const auto& r : m) {
    if (r.size() != m[0].size()) throw JWaveFailure("BasicTransform - given matrix is not rectangular");
    flat.insert(flat.end(), r.begin(), r.end());
  }
  return flat;
}
inline Matrix unflatten(const std::vector<double>& flat, size_t rows, size_t cols) {
  Matrix m(rows, std::vector<double>(cols));
  for (size_t i = 0; i < rows; i++) std::copy(flat.begin() + (long)(i * cols), flat.begin() + (long)((i + 1) * cols), m[i].begin());
  return m;
}
inline std::vector<double> flatten3(const Space& s) {
  std::vector<double> flat;
  for (const auto& m : s) {
    if (m.size() != s[0].size()) throw JWaveFailure("BasicTransform - given space is not a box");
    const std::vector<double> f = flatten(m);
    if (!m.empty() && m[0].size() != s[0][0].size()) throw JWaveFailure("BasicTransform - given space is not a box");
    flat.insert(flat.end(), f.begin(), f.end());
  }
  return flat;
}
inline Space unflatten3(const std::vector<double>& flat, size_t p, size_t q, size_t r) {
  Space s(p);
  for (size_t i = 0; i < p; i++)
    s[i] = unflatten(std::vector<double>(flat.begin() + (long)(i * q * r), flat.begin() + (long)((i + 1) * q * r)), q, r);
  return s;
}
}  // namespace detail

namespace detail {
inline void check_pyramid(const char* cls, const char* dir, int64_t len, int level) {
  if (!BasicTransform::isBinary(len))
    throw JWaveFailure(std::string(cls) + "#" + dir + " - given array length is not 2^p | p E N ... = 1, 2, 4, 8, 16, 32, .. please use the Ancient Egyptian Decomposition for any other array length!");
  if (level < 0 || level > BasicTransform::calcExponent(len))
    throw JWaveFailure(std::string(cls) + "#" + dir + " - given level is out of range for given array");
}
}  // namespace detail

class CudaFastWaveletTransform : public WaveletTransform {
 public:
  explicit CudaFastWaveletTransform(Wavelet w, std::shared_ptr<Context> ctx = nullptr) : WaveletTransform(std::move(w), std::move(ctx)) {
    name_ = "Fast Wavelet Transform";   // FastWaveletTransform.java:52
  }
  using WaveletTransform::forward;
  using WaveletTransform::reverse;
  std::vector<double> forward(const std::vector<double>& x, int level) override {
    detail::check_pyramid("FastWaveletTransform", "forward", (int64_t)x.size(), level);
    std::vector<double> out(x.size());
    call(jwc_fwt_forward, "jwc_fwt_forward", x.data(), out.data(), 1, (int64_t)x.size(), level,
         wavelet_.getScalingDeComposition(), wavelet_.getWaveletDeComposition());
    return out;
  }
  std::vector<double> reverse(const std::vector<double>& c, int level) override {
    detail::check_pyramid("FastWaveletTransform", "reverse", (int64_t)c.size(), level);
    std::vector<double> out(c.size());
    call(jwc_fwt_inverse, "jwc_fwt_inverse", c.data(), out.data(), 1, (int64_t)c.size(), level,
         wavelet_.getScalingReConstruction(), wavelet_.getWaveletReConstruction());
    return out;
  }
  // rows = independent signals, [batch][n] row-major host buffers
  void forwardBatch(const double* in, double* out, int64_t batch, int64_t n, int level) const {
    detail::check_pyramid("FastWaveletTransform", "forward", n, level);
    call(jwc_fwt_forward, "jwc_fwt_forward", in, out, batch, n, level, wavelet_.getScalingDeComposition(), wavelet_.getWaveletDeComposition());
  }
  void reverseBatch(const double* in, double* out, int64_t batch, int64_t n, int level) const {
    detail::check_pyramid("FastWaveletTransform", "reverse", n, level);
    call(jwc_fwt_inverse, "jwc_fwt_inverse", in, out, batch, n, level, wavelet_.getScalingReConstruction(), wavelet_.getWaveletReConstruction());
  }
  // 2-D: BasicTransform.java:336-399 forward(double[][][, lvlM, lvlN]) = rows (lvlN) then columns (lvlM);
  //      :412-474 reverse = columns first, then rows.  [batch][rows][cols] row-major host buffers for the batch form.
  void forward2DBatch(const double* in, double* out, int64_t batch, int64_t rows, int64_t cols, int lvlM, int lvlN) const {
    detail::check_pyramid("FastWaveletTransform", "forward", cols, lvlN);
    detail::check_pyramid("FastWaveletTransform", "forward", rows, lvlM);
    call2d(jwc_fwt2d_forward, "jwc_fwt2d_forward", in, out, batch, rows, cols, lvlM, lvlN, wavelet_.getScalingDeComposition(), wavelet_.getWaveletDeComposition());
  }
  void reverse2DBatch(const double* in, double* out, int64_t batch, int64_t rows, int64_t cols, int lvlM, int lvlN) const {
    detail::check_pyramid("FastWaveletTransform", "reverse", cols, lvlN);
    detail::check_pyramid("FastWaveletTransform", "reverse", rows, lvlM);
    call2d(jwc_fwt2d_inverse, "jwc_fwt2d_inverse", in, out, batch, rows, cols, lvlM, lvlN, wavelet_.getScalingReConstruction(), wavelet_.getWaveletReConstruction());
  }
  Matrix forward(const Matrix& matTime, int lvlM, int lvlN) const {
    const std::vector<double> flat = detail::flatten(matTime);
    std::vector<double> out(flat.size());
    forward2DBatch(flat.data(), out.data(), 1, (int64_t)matTime.size(), (int64_t)matTime.at(0).size(), lvlM, lvlN);
    return detail::unflatten(out, matTime.size(), matTime[0].size());
  }
  Matrix forward(const Matrix& matTime) const {
    return forward(matTime, calcExponent((int64_t)matTime.size()), calcExponent((int64_t)matTime.at(0).size()));
  }
  Matrix reverse(const Matrix& matHilb, int lvlM, int lvlN) const {
    const std::vector<double> flat = detail::flatten(matHilb);
    std::vector<double> out(flat.size());
    reverse2DBatch(flat.data(), out.data(), 1, (int64_t)matHilb.size(), (int64_t)matHilb.at(0).size(), lvlM, lvlN);
    return detail::unflatten(out, matHilb.size(), matHilb[0].size());
  }
  Matrix reverse(const Matrix& matHilb) const {
    return reverse(matHilb, calcExponent((int64_t)matHilb.size()), calcExponent((int64_t)matHilb.at(0).size()));
  }

  // 3-D: BasicTransform.java:487-640 forward / reverse(double[][][][, lvlP, lvlQ, lvlR]): the 2-D transform of every matrix
  //      spc[i] with (lvlP, lvlQ) -- rows of length r with lvlQ, columns of length q with lvlP --, then every line along
  //      the first axis (length p) with lvlR; the reverse keeps that order.  [batch][p][q][r] row-major for the batch form.
  void forward3DBatch(const double* in, double* out, int64_t batch, int64_t p, int64_t q, int64_t r, int lvlP, int lvlQ, int lvlR) const {
    detail::check_pyramid("FastWaveletTransform", "forward", r, lvlQ);
    detail::check_pyramid("FastWaveletTransform", "forward", q, lvlP);
    detail::check_pyramid("FastWaveletTransform", "forward", p, lvlR);
    call3d(jwc_fwt3d_forward, "jwc_fwt3d_forward", in, out, batch, p, q, r, lvlP, lvlQ, lvlR, wavelet_.getScalingDeComposition(), wavelet_.getWaveletDeComposition());
  }
  void reverse3DBatch(const double* in, double* out, int64_t batch, int64_t p, int64_t q, int64_t r, int lvlP, int lvlQ, int lvlR) const {
    detail::check_pyramid("FastWaveletTransform", "reverse", r, lvlQ);
    detail::check_pyramid("FastWaveletTransform", "reverse", q, lvlP);
    detail::check_pyramid("FastWaveletTransform", "reverse", p, lvlR);
    call3d(jwc_fwt3d_inverse, "jwc_fwt3d_inverse", in, out, batch, p, q, r, lvlP, lvlQ, lvlR, wavelet_.getScalingReConstruction(), wavelet_.getWaveletReConstruction());
  }
  Space forward(const Space& spcTime, int lvlP, int lvlQ, int lvlR) const {
    const std::vector<double> flat = detail::flatten3(spcTime);
    std::vector<double> out(flat.size());
    forward3DBatch(flat.data(), out.data(), 1, (int64_t)spcTime.size(), (int64_t)spcTime.at(0).size(), (int64_t)spcTime.at(0).at(0).size(), lvlP, lvlQ, lvlR);
    return detail::unflatten3(out, spcTime.size(), spcTime[0].size(), spcTime[0][0].size());
  }
  Space forward(const Space& spcTime) const {   // :490-493: the exponents of the three dimensions, in this order
    return forward(spcTime, calcExponent((int64_t)spcTime.size()), calcExponent((int64_t)spcTime.at(0).size()), calcExponent((int64_t)spcTime.at(0).at(0).size()));
  }
  Space reverse(const Space& spcHilb, int lvlP, int lvlQ, int lvlR) const {
    const std::vector<double> flat = detail::flatten3(spcHilb);
    std::vector<double> out(flat.size());
    reverse3DBatch(flat.data(), out.data(), 1, (int64_t)spcHilb.size(), (int64_t)spcHilb.at(0).size(), (int64_t)spcHilb.at(0).at(0).size(), lvlP, lvlQ, lvlR);
    return detail::unflatten3(out, spcHilb.size(), spcHilb[0].size(), spcHilb[0][0].size());
  }
  Space reverse(const Space& spcHilb) const {
    return reverse(spcHilb, calcExponent((int64_t)spcHilb.size()), calcExponent((int64_t)spcHilb.at(0).size()), calcExponent((int64_t)spcHilb.at(0).at(0).size()));
  }
};

class CudaWaveletPacketTransform : public WaveletTransform {
 public:
  explicit CudaWaveletPacketTransform(Wavelet w, std::shared_ptr<Context> ctx = nullptr) : WaveletTransform(std::move(w), std::move(ctx)) {
    name_ = "Wavelet Packet Transform";   // WaveletPacketTransform.java:54
  }
  using WaveletTransform::forward;
  using WaveletTransform::reverse;
  std::vector<double> forward(const std::vector<double>& x, int level) override {
    detail::check_pyramid("WaveletPacketTransform", "forward", (int64_t)x.size(), level);
    std::vector<double> out(x.size());
    call(jwc_wpt_forward, "jwc_wpt_forward", x.data(), out.data(), 1, (int64_t)x.size(), level,
         wavelet_.getScalingDeComposition(), wavelet_.getWaveletDeComposition());
    return out;
  }
  std::vector<double> reverse(const std::vector<double>& c, int level) override {
    detail::check_pyramid("WaveletPacketTransform", "reverse", (int64_t)c.size(), level);
    std::vector<double> out(c.size());
    call(jwc_wpt_inverse, "jwc_wpt_inverse", c.data(), out.data(), 1, (int64_t)c.size(), level,
         wavelet_.getScalingReConstruction(), wavelet_.getWaveletReConstruction());
    return out;
  }
  void forwardBatch(const double* in, double* out, int64_t batch, int64_t n, int level) const {
    detail::check_pyramid("WaveletPacketTransform", "forward", n, level);
    call(jwc_wpt_forward, "jwc_wpt_forward", in, out, batch, n, level, wavelet_.getScalingDeComposition(), wavelet_.getWaveletDeComposition());
  }
  void reverseBatch(const double* in, double* out, int64_t batch, int64_t n, int level) const {
    detail::check_pyramid("WaveletPacketTransform", "reverse", n, level);
    call(jwc_wpt_inverse, "jwc_wpt_inverse", in, out, batch, n, level, wavelet_.getScalingReConstruction(), wavelet_.getWaveletReConstruction());
  }
  // 2-D: BasicTransform.java:336-399 forward(double[][][, lvlM, lvlN]) = rows (lvlN) then columns (lvlM);
  //      :412-474 reverse = columns first, then rows.  [batch][rows][cols] row-major host buffers for the batch form.
  void forward2DBatch(const double* in, double* out, int64_t batch, int64_t rows, int64_t cols, int lvlM, int lvlN) const {
    detail::check_pyramid("WaveletPacketTransform", "forward", cols, lvlN);
    detail::check_pyramid("WaveletPacketTransform", "forward", rows, lvlM);
    call2d(jwc_wpt2d_forward, "jwc_wpt2d_forward", in, out, batch, rows, cols, lvlM, lvlN, wavelet_.getScalingDeComposition(), wavelet_.getWaveletDeComposition());
  }
  void reverse2DBatch(const double* in, double* out, int64_t batch, int64_t rows, int64_t cols, int lvlM, int lvlN) const {
    detail::check_pyramid("WaveletPacketTransform", "reverse", cols, lvlN);
    detail::check_pyramid("WaveletPacketTransform", "reverse", rows, lvlM);
    call2d(jwc_wpt2d_inverse, "jwc_wpt2d_inverse", in, out, batch, rows, cols, lvlM, lvlN, wavelet_.getScalingReConstruction(), wavelet_.getWaveletReConstruction());
  }
  Matrix forward(const Matrix& matTime, int lvlM, int lvlN) const {
    const std::vector<double> flat = detail::flatten(matTime);
    std::vector<double> out(flat.size());
    forward2DBatch(flat.data(), out.data(), 1, (int64_t)matTime.size(), (int64_t)matTime.at(0).size(), lvlM, lvlN);
    return detail::unflatten(out, matTime.size(), matTime[0].size());
  }
  Matrix forward(const Matrix& matTime) const {
    return forward(matTime, calcExponent((int64_t)matTime.size()), calcExponent((int64_t)matTime.at(0).size()));
  }
  Matrix reverse(const Matrix& matHilb, int lvlM, int lvlN) const {
    const std::vector<double> flat = detail::flatten(matHilb);
    std::vector<double> out(flat.size());
    reverse2DBatch(flat.data(), out.data(), 1, (int64_t)matHilb.size(), (int64_t)matHilb.at(0).size(), lvlM, lvlN);
    return detail::unflatten(out, matHilb.size(), matHilb[0].size());
  }
  Matrix reverse(const Matrix& matHilb) const {
    return reverse(matHilb, calcExponent((int64_t)matHilb.size()), calcExponent((int64_t)matHilb.at(0).size()));
  }

  // 3-D: BasicTransform.java:487-640 forward / reverse(double[][][][, lvlP, lvlQ, lvlR]): the 2-D transform of every matrix
  //      spc[i] with (lvlP, lvlQ) -- rows of length r with lvlQ, columns of length q with lvlP --, then every line along
  //      the first axis (length p) with lvlR; the reverse keeps that order.  [batch][p][q][r] row-major for the batch form.
  void forward3DBatch(const double* in, double* out, int64_t batch, int64_t p, int64_t q, int64_t r, int lvlP, int lvlQ, int lvlR) const {
    detail::check_pyramid("WaveletPacketTransform", "forward", r, lvlQ);
    detail::check_pyramid("WaveletPacketTransform", "forward", q, lvlP);
    detail::check_pyramid("WaveletPacketTransform", "forward", p, lvlR);
    call3d(jwc_wpt3d_forward, "jwc_wpt3d_forward", in, out, batch, p, q, r, lvlP, lvlQ, lvlR, wavelet_.getScalingDeComposition(), wavelet_.getWaveletDeComposition());
  }
  void reverse3DBatch(const double* in, double* out, int64_t batch, int64_t p, int64_t q, int64_t r, int lvlP, int lvlQ, int lvlR) const {
    detail::check_pyramid("WaveletPacketTransform", "reverse", r, lvlQ);
    detail::check_pyramid("WaveletPacketTransform", "reverse", q, lvlP);
    detail::check_pyramid("WaveletPacketTransform", "reverse", p, lvlR);
    call3d(jwc_wpt3d_inverse, "jwc_wpt3d_inverse", in, out, batch, p, q, r, lvlP, lvlQ, lvlR, wavelet_.getScalingReConstruction(), wavelet_.getWaveletReConstruction());
  }
  Space forward(const Space& spcTime, int lvlP, int lvlQ, int lvlR) const {
    const std::vector<double> flat = detail::flatten3(spcTime);
    std::vector<double> out(flat.size());
    forward3DBatch(flat.data(), out.data(), 1, (int64_t)spcTime.size(), (int64_t)spcTime.at(0).size(), (int64_t)spcTime.at(0).at(0).size(), lvlP, lvlQ, lvlR);
    return detail::unflatten3(out, spcTime.size(), spcTime[0].size(), spcTime[0][0].size());
  }
  Space forward(const Space& spcTime) const {   // :490-493: the exponents of the three dimensions, in this order
    return forward(spcTime, calcExponent((int64_t)spcTime.size()), calcExponent((int64_t)spcTime.at(0).size()), calcExponent((int64_t)spcTime.at(0).at(0).size()));
  }
  Space reverse(const Space& spcHilb, int lvlP, int lvlQ, int lvlR) const {
    const std::vector<double> flat = detail::flatten3(spcHilb);
    std::vector<double> out(flat.size());
    reverse3DBatch(flat.data(), out.data(), 1, (int64_t)spcHilb.size(), (int64_t)spcHilb.at(0).size(), (int64_t)spcHilb.at(0).at(0).size(), lvlP, lvlQ, lvlR);
    return detail::unflatten3(out, spcHilb.size(), spcHilb[0].size(), spcHilb[0][0].size());
  }
  Space reverse(const Space& spcHilb) const {
    return reverse(spcHilb, calcExponent((int64_t)spcHilb.size()), calcExponent((int64_t)spcHilb.at(0).size()), calcExponent((int64_t)spcHilb.at(0).at(0).size()));
  }
};

class CudaMODWTTransform : public WaveletTransform {
 public:
  static constexpr int MAX_DECOMPOSITION_LEVEL = 13;   // MODWTTransform.java:111
  explicit CudaMODWTTransform(Wavelet w, std::shared_ptr<Context> ctx = nullptr) : WaveletTransform(std::move(w), std::move(ctx)) {
    name_ = "MODWT";
    // MODWTTransform.java:462-475 + normalize :599-606
    g_ = normalize(wavelet_.getScalingDeComposition());
    h_ = normalize(wavelet_.getWaveletDeComposition());
    const double s = std::sqrt(2.0);
    for (size_t i = 0; i < g_.size(); i++) { g_[i] = g_[i] / s; h_[i] = h_[i] / s; }
  }
  static int getMaxDecompositionLevel() { return MAX_DECOMPOSITION_LEVEL; }
  // MODWTTransform.java:148-153, :191-213: the reference switches its CPU loops between the direct and the FFT circular
  // convolution; the device path has one arithmetic, so the setting (and the FFT threshold of the two-argument
  // constructor) is stored for callers that read it back and changes nothing
  enum class ConvolutionMethod { AUTO, DIRECT, FFT };
  CudaMODWTTransform(Wavelet w, int fftThreshold, std::shared_ptr<Context> ctx = nullptr) : CudaMODWTTransform(std::move(w), std::move(ctx)) {
    fftThreshold_ = fftThreshold;
  }
  void setConvolutionMethod(ConvolutionMethod m) { method_ = m; }
  ConvolutionMethod getConvolutionMethod() const { return method_; }

  // MODWTTransform.java:256-306; rows W_1..W_J, V_J
  std::vector<std::vector<double>> forwardMODWT(const std::vector<double>& data, int maxLevel) const {
    checkLevel(maxLevel);
    if (data.empty()) return std::vector<std::vector<double>>((size_t)maxLevel + 1);
    const int64_t n = (int64_t)data.size();
    checkLimit(maxLevel, n);
    std::vector<double> flat((size_t)(maxLevel + 1) * (size_t)n);
    call(jwc_modwt_forward, "jwc_modwt_forward", data.data(), flat.data(), 1, n, maxLevel, g_, h_);
    std::vector<std::vector<double>> rows;
    for (int r = 0; r <= maxLevel; r++) rows.emplace_back(flat.begin() + (size_t)r * n, flat.begin() + (size_t)(r + 1) * n);
    return rows;
  }
  // MODWTTransform.java:337-375
  std::vector<double> inverseMODWT(const std::vector<std::vector<double>>& coeffs) const {
    if (coeffs.size() < 2) return {};
    const int maxLevel = (int)coeffs.size() - 1;
    const int64_t n = (int64_t)coeffs[0].size();
    if (n == 0) return {};
    std::vector<double> flat;
    for (const auto& r : coeffs) flat.insert(flat.end(), r.begin(), r.end());
    std::vector<double> x((size_t)n);
    call(jwc_modwt_inverse, "jwc_modwt_inverse", flat.data(), x.data(), 1, n, maxLevel, g_, h_);
    return x;
  }
  // flattened 1-D interface, MODWTTransform.java:389-443, 854-912
  std::vector<double> forward(const std::vector<double>& x, int level) override {
    if (x.empty()) return {};
    if (!isBinary((int64_t)x.size())) throw JWaveFailure("MODWTTransform#forward - given array length is not 2^p | p E N ... = 1, 2, 4, 8, 16, 32, .. ");
    if (level < 0 || level > calcExponent((int64_t)x.size())) throw JWaveFailure("MODWTTransform#forward - given level is out of range for given array");
    if (level > MAX_DECOMPOSITION_LEVEL) throw JWaveFailure("MODWTTransform#forward - maximum supported decomposition level is 13, requested: " + std::to_string(level));
    return flatten(forwardMODWT(x, level));
  }
  std::vector<double> forward(const std::vector<double>& x) override {
    if (x.empty()) return {};
    return flatten(forwardMODWT(x, calcExponent((int64_t)x.size())));
  }
  std::vector<double> reverse(const std::vector<double>& c, int level) override {
    if (c.empty()) return {};
    const int64_t n = (int64_t)c.size() / (level + 1);
    if (!isBinary(n)) throw JWaveFailure("MODWTTransform#reverse - Invalid coefficient array for given level");
    if ((int64_t)c.size() != n * (level + 1)) throw JWaveFailure("MODWTTransform#reverse - Coefficient array length does not match expected size for given level");
    return inverseMODWT(unflatten(c, level, n));
  }
  std::vector<double> reverse(const std::vector<double>& c) override {
    if (c.empty()) return {};
    const int64_t total = (int64_t)c.size();
    for (int64_t testN = 1; testN <= total; testN++) {   // :888-897
      if (total % testN) continue;
      const int64_t lv = total / testN - 1;
      if (lv >= 0 && isBinary(testN) && lv <= calcExponent(testN)) return inverseMODWT(unflatten(c, (int)lv, testN));
    }
    throw JWaveFailure("MODWTTransform#reverse - Invalid flattened coefficient array length. Cannot determine original signal dimensions.");
  }
  // batch: x [batch][n] -> coeffs [batch][J+1][n]
  void forwardMODWTBatch(const double* x, double* coeffs, int64_t batch, int64_t n, int maxLevel) const {
    checkLevel(maxLevel);
    checkLimit(maxLevel, n);
    call(jwc_modwt_forward, "jwc_modwt_forward", x, coeffs, batch, n, maxLevel, g_, h_);
  }
  void inverseMODWTBatch(const double* coeffs, double* x, int64_t batch, int64_t n, int maxLevel) const {
    checkLevel(maxLevel);
    call(jwc_modwt_inverse, "jwc_modwt_inverse", coeffs, x, batch, n, maxLevel, g_, h_);
  }

 private:
  static std::vector<double> normalize(std::vector<double> f) {
    double energy = 0.0;
    for (double c : f) energy += c * c;
    const double norm = std::sqrt(energy);
    if (norm > 1e-12) for (double& c : f) c /= norm;
    return f;
  }
  static void checkLevel(int J) {   // MODWTTransform.java:257-265
    if (J < 1) throw std::invalid_argument("MODWTTransform#forwardMODWT - decomposition level must be at least 1, requested: " + std::to_string(J));
    if (J > MAX_DECOMPOSITION_LEVEL) throw std::invalid_argument("MODWTTransform#forwardMODWT - maximum supported decomposition level is 13, requested: " + std::to_string(J));
  }
  static void checkLimit(int J, int64_t n) {   // :276-282
    int lim = 0;
    while (((int64_t)1 << (lim + 1)) <= n) lim++;
    if (J > lim) throw std::invalid_argument("Decomposition level " + std::to_string(J) + " exceeds theoretical limit " + std::to_string(lim) + " for signal length " + std::to_string(n));
  }
  static std::vector<double> flatten(const std::vector<std::vector<double>>& rows) {
    std::vector<double> f;
    for (const auto& r : rows) f.insert(f.end(), r.begin(), r.end());
    return f;
  }
  static std::vector<std::vector<double>> unflatten(const std::vector<double>& c, int level, int64_t n) {
    std::vector<std::vector<double>> rows;
    for (int r = 0; r <= level; r++) rows.emplace_back(c.begin() + (size_t)r * n, c.begin() + (size_t)(r + 1) * n);
    return rows;
  }
  std::vector<double> g_, h_;
  int fftThreshold_ = 4096;   // MODWTTransform.java:144
  ConvolutionMethod method_ = ConvolutionMethod::AUTO;
};

}  // namespace jwave
