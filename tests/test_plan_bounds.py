"""CPU-only: the pass planners (jwc_modwt_plan.cuh, jwc_dwt_plan.cuh) against the kernels' indexing rules.

compute-sanitizer is closed on the GPU pool, so the shared-memory bounds of the fused kernels are audited here: for a
grid of shapes the plan is dumped (tests/cpp/plan_dump.cpp, host-only C++) and every shared-memory index the kernels
can touch under that plan -- following the formulas in jwc_modwt_fast.cu / jwc_dwt_fast.cu -- must stay inside the
buffers, the buffers inside the dynamic shared-memory size, and the passes must cover the levels exactly once."""
import json
import math
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "plan_dump")


@pytest.fixture(scope="module")
def dump():
    src = os.path.join(ROOT, "tests", "cpp", "plan_dump.cpp")
    deps = [src] + [os.path.join(ROOT, "jwave-pro_b200", "csrc", f) for f in ("jwc_modwt_plan.cuh", "jwc_dwt_plan.cuh")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-x", "c++", src, "-o", EXE])

    def run(kind, n, levels, L, inverse, budget):
        out = subprocess.run([EXE, kind, str(n), str(levels), str(L), str(int(inverse)), str(budget)],
                             capture_output=True, text=True, check=True).stdout
        return json.loads(out)
    return run


@pytest.mark.parametrize("L", [2, 4, 8, 16, 20, 40])
@pytest.mark.parametrize("n,J", [(65536, 6), (65536, 8), (65536, 13), (1024, 3), (8192, 13), (100, 6), (65538, 6),
                                 (65537, 4), (1 << 20, 10), (40000, 7), (8, 3)])
@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("budget", [75776, 113000])
def test_modwt_plan_bounds(dump, L, n, J, inverse, budget):
    plan = dump("modwt", n, J, L, inverse, budget)
    R = plan["R"]
    j = 0
    for p in plan["passes"]:
        assert p["j0"] == j and p["k"] >= 1
        j += p["k"]
        P = 1 << p["logP"]
        S0 = 1 << p["j0"]
        G = math.gcd(S0, n)   # interleaved cycles of the walk t -> t + 2^j0 (mod n); = 2^j0 when that divides n
        assert p["cycles"] == G and n % G == 0
        assert P <= G and p["T2"] >= 2 and p["T2"] % 2 == 0 and p["vcap"] % 2 == 0
        Nd = n // G
        # along a cycle, i -> (i 2^j0) mod n visits every position of the cycle's residue class exactly once
        if n <= 70000 and p["j0"] > 0:
            seen = {(i * S0) % n for i in range(Nd)}
            assert len(seen) == Nd and all(t % G == 0 for t in seen) and max(seen) + G <= n
        tlen2 = min(p["T2"], Nd)
        H = (L - 1) * ((1 << p["k"]) - 1)
        assert p["Hp"] >= H
        if p["mode"] == 0:   # bulk: 16-byte pieces, single wrap
            assert P == 1 and p["j0"] == 0 and n % 2 == 0 and p["Hp"] % 2 == 0 and p["Hp"] <= n
        vcap = p["vcap"]
        assert P * (tlen2 + p["Hp"]) <= vcap, "tile + halo fits the V buffer"
        for jj in range(1, p["k"] + 1):
            s = P << (jj - 1)
            if not inverse:
                hrem = (L - 1) * ((1 << p["k"]) - (1 << jj))
                e0, ln = P * (p["Hp"] - hrem), P * (hrem + tlen2)
                rows = -(-ln // s)
                nrb = -(-rows // R)
                top = e0 + (nrb * R - 1) * s + (s - 1)          # highest index an item may read (junk rows included)
                low = e0 - (L - 1) * s                            # lowest index read
                assert low >= 0 and top < vcap, (jj, low, top, vcap)
            else:
                hout = (L - 1) * ((1 << (jj - 1)) - 1)
                ln = P * (tlen2 + hout)
                rows = -(-ln // s)
                nrb = -(-rows // R)
                top = (nrb * R - 1) * s + (s - 1) + (L - 1) * s   # rows R+L-2 beyond the item start
                assert top < vcap, (jj, top, vcap)
        staged = 0 if (P >= 4 and not inverse) else 2 * P * p["T2"]
        doubles = (4 * vcap) if inverse else (2 * vcap + staged)
        assert doubles * 8 + 1024 + 64 <= p["smem"] <= budget + 2048
    if plan["all_fused"]:
        assert j == J
    else:
        assert plan["generic_from"] == j < J


@pytest.mark.parametrize("L", [2, 8, 16, 40])
@pytest.mark.parametrize("kind,n,levels", [("fwt", 1 << 20, 20), ("fwt", 65536, 16), ("wpt", 65536, 6), ("wpt", 65536, 16),
                                           ("wpt", 1024, 10), ("fwt", 4, 2), ("wpt", 64, 6), ("fwt", 2, 1),
                                           ("wpt", 1 << 20, 8), ("fwt", 4096, 3)])
@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("budget", [45000, 113000])
def test_dwt_plan_bounds(dump, L, kind, n, levels, inverse, budget):
    plan = dump(kind, n, levels, L, inverse, budget)
    if not plan["ok"]:
        pytest.skip("shape declined by the fused path (generic kernels take it)")
    R = plan["R"]
    tree = kind == "wpt"

    def stride(ln):
        return ln + (ln & 1) + 2 * R

    def inv_halo(jj):
        h = 0
        for _ in range(jj):
            h = (h + 1) // 2 + (L // 2 - 1)
            h += h & 1
        return h

    l = 0
    for p in plan["passes"]:
        assert p["l0"] == l and p["k"] >= 1
        l += p["k"]
        h = n >> p["l0"]
        tlen = min(h, p["T"])
        k, cap = p["k"], p["cap"]
        assert tlen >= (1 << k) and tlen & (tlen - 1) == 0 and cap % 2 == 0
        for jj in range(0, k + 1):
            nodes = 1 if jj == 0 else ((1 << jj) if tree else 2)
            if not inverse:
                ln = (tlen >> jj) + (L - 2) * ((1 << (k - jj)) - 1)
            else:
                ln = (tlen >> jj) + inv_halo(jj)
            assert nodes * stride(ln) <= cap, (jj, nodes, ln, cap)
            if jj >= 1 and not inverse:
                # an item of R outputs reads pairs up to index 2*(i0 + R - 1) + L - 1 of its parent, i0 <= len_out - 1
                ln_in = (tlen >> (jj - 1)) + (L - 2) * ((1 << (k - jj + 1)) - 1)
                nb = -(-ln // R)
                top = 2 * (nb * R - 1) + L - 1
                assert top < stride(ln_in), (jj, top, stride(ln_in))   # stays within the padded parent node
                assert 2 * ln + L - 2 <= ln_in
            if jj >= 1 and inverse:
                hl_out, hl_in = inv_halo(jj - 1), inv_halo(jj)
                npairs = (hl_out >> 1) + (tlen >> jj)
                off = hl_in - (hl_out >> 1) - (L // 2 - 1)
                assert off >= 0
                nb = -(-npairs // R)
                top = off + nb * R - 1 + L // 2 - 1
                assert top < stride(ln), (jj, top, stride(ln))
                assert 2 * npairs == (tlen >> (jj - 1)) + hl_out
        if p["mode"] == 0 and inverse:
            assert (tlen >> k) % 2 == 0
        assert 2 * cap * 8 + 2 * 64 * 8 + 128 <= p["smem"] <= budget + 4096
    assert l == levels


@pytest.mark.parametrize("d", [1, 2, 3, 5, 7, 39, 40, 41, 64, 79, 512, 1000, 4095, 4096, 4097, 65535, 65536, 99999,
                               (1 << 20) + 1, (1 << 30) - 1, 1 << 30, (1 << 31) - 1])
def test_fastdiv(dump, d):
    """The MODWT tile kernels take blockIdx apart with a multiply-high by the tile count's magic number
    (jwc_modwt_plan.cuh::make_fastdiv / fastdiv): exact for every n < 2^31 the launcher allows -- small n, multiples
    of d and their neighbours, the top of the range, 2 M pseudo-random values."""
    r = dump("fastdiv", d, 0, 0, False, 0)
    assert r["ok"] == 1 and r["checked"] > 2_000_000, r


@pytest.mark.parametrize("L", [2, 4, 8, 16, 20, 30, 40])
@pytest.mark.parametrize("n,levels", [(1 << 20, 20), (65536, 16), (4096, 12), (1 << 18, 7), (1024, 10), (256, 8)])
@pytest.mark.parametrize("group", [0, 4, 5, 8, 12])
@pytest.mark.parametrize("budget", [45000, 113000])
def test_fwt_inverse_upfront_tiles_nest(dump, L, n, levels, group, budget):
    """Pyramid inverse with every detail tile of a pass requested in the prologue (jwc_dwt_fast.cu, `upfront`): whenever
    the host-side condition of fast_dwt_inverse holds, the k + 1 TMA-written tiles and the low-pass arrays the levels
    write never overlap while they are live, and every slot stays inside its buffer."""
    exe_args = ["fwt", n, levels, L, True, budget]
    import subprocess as sp
    out = sp.run([EXE] + [str(int(a)) if not isinstance(a, str) else a for a in exe_args] + ([str(group)] if group else []),
                 capture_output=True, text=True, check=True).stdout
    plan = json.loads(out)
    if not plan["ok"]:
        pytest.skip("shape declined by the fused path")
    R = plan["R"]

    def stride(ln):
        return ln + (ln & 1) + 2 * R

    def halo(jj):
        h = 0
        for _ in range(jj):
            h = (h + 1) // 2 + (L // 2 - 1)
            h += h & 1
        return h

    checked = 0
    for p in plan["passes"]:
        k, cap = p["k"], p["cap"]
        tlen = min(n >> p["l0"], p["T"])
        ln = [(tlen >> j) + halo(j) for j in range(k + 1)]
        if p["mode"] != 0 or k < 2 or k > 16:
            continue
        if not all(stride(ln[j - 2]) >= stride(ln[j]) + ln[j] for j in range(3, k + 1)):
            continue   # the kernel keeps the one-level-ahead prefetch for this pass
        checked += 1
        # (buffer, lo, hi, first step, last step) of everything that holds data; step -1 = prologue, step u = level k - u
        live = [(0, 0, ln[k], -1, 0), (0, stride(ln[k]), stride(ln[k]) + ln[k], -1, 0)]
        for j in range(k - 1, 0, -1):
            u = k - j
            live.append((u & 1, stride(ln[j]), stride(ln[j]) + ln[j], -1, u))          # D_j, TMA-written, read by step u
        for u in range(k):
            j_out = k - u - 1
            live.append(((u + 1) & 1, 0, ln[j_out], u, u + 1))                            # A_{j_out}, written by step u
        for (b, lo, hi, _, _) in live:
            assert 0 <= lo < hi <= cap, (p, b, lo, hi)
        for x in range(len(live)):
            for y in range(x + 1, len(live)):
                bx, lx, hx, fx, tx = live[x]
                by, ly, hy, fy, ty = live[y]
                if bx == by and max(fx, fy) <= min(tx, ty) and max(lx, ly) < min(hx, hy):
                    raise AssertionError(("overlap", p, live[x], live[y]))
    if L <= 16 and n >= 65536 and budget == 45000:
        assert checked > 0   # the shapes of the benchmark do take the up-front path
