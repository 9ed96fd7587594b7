#!/usr/bin/env python3
"""Write the synthetic inputs of SURVEY.md section 8d as little-endian fp64 .bin files, so that the oracle, the GPU path
and -- where a JDK exists -- the real reference (java/tools/JWaveOracleDump.java) read identical bits.

    python tools/make_inputs.py outdir            # c1_random.bin, c1_chirp.bin, c2_random_row0.bin, ...
    python tools/make_inputs.py outdir --check ref_out.bin modwt Daubechies4 6 c2_random_row0.bin
        compares a dump produced by JWaveOracleDump with the oracle (bitwise) and prints the max deviation
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jwave_pro_b200.synth import chirp, splitmix_uniform  # noqa: E402

SHAPES = {"c1": 1024, "c2": 65536, "c3": 1 << 20, "c4": 65536, "c5": 65536}


def main():
    out = sys.argv[1]
    if "--check" in sys.argv:
        i = sys.argv.index("--check")
        dump, kind, cls, level, inp = sys.argv[i + 1:i + 6]
        import jwave_pro_b200 as jw
        from oracle import c_oracle as oracle
        x = np.fromfile(os.path.join(out, inp), dtype="<f8")
        w = jw.wavelets.create(cls)
        s, wv = w.getScalingDeComposition(), w.getWaveletDeComposition()
        if kind == "modwt":
            g, h = oracle.modwt_filters(s, wv)
            ref = oracle.modwt_forward(x, int(level), g, h).reshape(-1)
        elif kind == "fwt":
            ref = oracle.fwt_forward(x, int(level), s, wv)
        else:
            ref = oracle.wpt_forward(x, int(level), s, wv)
        got = np.fromfile(dump, dtype="<f8")
        print("bitwise equal:", np.array_equal(got, ref), " max |diff|:", float(np.max(np.abs(got - ref))))
        return
    os.makedirs(out, exist_ok=True)
    for i, (name, n) in enumerate(SHAPES.items()):
        splitmix_uniform(0x5EED0000 + i + 1, (n,)).astype("<f8").tofile(os.path.join(out, "%s_random_row0.bin" % name))
        chirp(1, n)[0].astype("<f8").tofile(os.path.join(out, "%s_chirp_row0.bin" % name))
    print("wrote", sorted(os.listdir(out)))


if __name__ == "__main__":
    main()
