#!/usr/bin/env python3
"""Collect the reference's own known-answer vectors for the hot path into tests/golden/reference_kats.json.

Runs only where /root/reference exists (the build container); the JSON it writes is committed and is what the
tests read.  Sources (relative to /root/reference/src/test/):
  resources/testdata/haar_simple_input.txt, haar_level1_{approx,detail}_manual.txt,
  resources/testdata/filter_haar_{dec,rec}_{lo,hi}.txt           (java/jwave/transforms/CrossValidationTest.java:159-209)
  java/jwave/transforms/MODWTTransformTest.java:39-71              MODWT Haar level 1 of [1..8] (literals below)
  java/jwave/transforms/MODWTFFTAdjointVerificationTest.java:44-101  adjoint == transpose of the convolution matrix
  resources/testdata/fft_{dc,impulse}_{input,output_real,output_imag}.txt   (CrossValidationTest.java:119-154)
  resources/testdata/filter_db2_dec_lo.txt, filter_db4_dec_{lo,hi}.txt, haar_{constant,linear}_input.txt
"""
import json
import os

R = "/root/reference/src/test/resources/testdata"
HERE = os.path.dirname(os.path.abspath(__file__))


def read(name):
    vals = []
    for line in open(os.path.join(R, name)):
        line = line.strip()
        if line and not line.startswith("#"):
            vals.append(float(line))
    return vals


def main():
    g = {}
    g["haar_fwt_level1"] = {
        "source": "CrossValidationTest.java:187-209, tolerance 1e-10",
        "input": read("haar_simple_input.txt"),
        "approx": read("haar_level1_approx_manual.txt"),
        "detail": read("haar_level1_detail_manual.txt"),
    }
    g["haar_filters"] = {
        "source": "CrossValidationTest.java:159-182, tolerance 1e-10",
        "dec_lo": read("filter_haar_dec_lo.txt"), "dec_hi": read("filter_haar_dec_hi.txt"),
        "rec_lo": read("filter_haar_rec_lo.txt"), "rec_hi": read("filter_haar_rec_hi.txt"),
    }
    # MODWTTransformTest.java:39-71 (expected values are written there as products with the taps +-0.5)
    x = [1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0]
    g["modwt_haar_level1"] = {
        "source": "MODWTTransformTest.java:39-71, tolerance 1e-9",
        "input": x,
        "D1": [x[i] * 0.5 + x[i - 1] * -0.5 for i in range(8)],
        "A1": [x[i] * 0.5 + x[i - 1] * 0.5 for i in range(8)],
    }
    # MODWTFFTAdjointVerificationTest.java:44-101: H[i][j] = f[(i-j) mod N] if (i-j) mod N < M; adjoint = H^T x
    sig, f = [1.0, 2.0, 3.0, 4.0], [0.5, -0.5]
    N, M = 4, 2
    H = [[(f[(i - j) % N] if (i - j) % N < M else 0.0) for j in range(N)] for i in range(N)]
    g["adjoint_transpose"] = {
        "source": "MODWTFFTAdjointVerificationTest.java:44-101, tolerance 1e-10",
        "signal": sig, "filter": f,
        "direct": [sum(H[i][j] * sig[j] for j in range(N)) for i in range(N)],
        "adjoint": [sum(H[j][i] * sig[j] for j in range(N)) for i in range(N)],
    }
    # FFT fixtures (the FFT is only the CPU baseline's engine here, but the baseline should be the reference's FFT)
    g["fft_fixtures"] = {
        "source": "CrossValidationTest.java:119-154 (testFFTWithReferenceData), tolerance 1e-10",
        "dc": {"input": read("fft_dc_input.txt"), "real": read("fft_dc_output_real.txt"),
               "imag": read("fft_dc_output_imag.txt")},
        "impulse": {"input": read("fft_impulse_input.txt"), "real": read("fft_impulse_output_real.txt"),
                    "imag": read("fft_impulse_output_imag.txt")},
        "sine_input": read("fft_sine_simple_input.txt"),
    }
    # filter fixtures the reference ships beside the Haar ones (PyWavelets naming: db2 = 2 taps = Haar1, db4 = 4 taps
    # = JWave's Daubechies2); no reference test reads them, they pin the extracted tables to 17 digits
    g["filter_fixtures"] = {
        "source": "src/test/resources/testdata/filter_db2_dec_lo.txt, filter_db4_dec_{lo,hi}.txt",
        "Haar1": {"dec_lo": read("filter_db2_dec_lo.txt")},
        "Daubechies2": {"dec_lo": read("filter_db4_dec_lo.txt"), "dec_hi": read("filter_db4_dec_hi.txt")},
    }
    g["haar_more_inputs"] = {
        "source": "src/test/resources/testdata/haar_constant_input.txt, haar_linear_input.txt (inputs only)",
        "constant": read("haar_constant_input.txt"), "linear": read("haar_linear_input.txt"),
    }
    with open(os.path.join(HERE, "reference_kats.json"), "w") as fh:
        json.dump(g, fh, indent=1)
    print(json.dumps(g, indent=1)[:600])


if __name__ == "__main__":
    main()
