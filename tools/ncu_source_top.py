"""Summarise `ncu --page source --csv` output: per kernel, stall-sample totals by reason, by opcode, and by code region.

    python tools/ncu_source_top.py gpurun_out/r3c_c5f_source.csv [--regions N] [--dump lo hi]

A region = a run of instructions with the same execution count (straight-line code executed together)."""
import csv
import sys
from collections import defaultdict


def kernels(path):
    cur, hdr, rows = None, None, []
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "Kernel Name":
            if cur is not None:
                yield cur, hdr, rows
            cur, hdr, rows = r[1], None, []
        elif r[0] == "Address":
            hdr = r
        elif hdr is not None:
            rows.append(r)
    if cur is not None:
        yield cur, hdr, rows


def main():
    path = sys.argv[1]
    nreg = int(sys.argv[sys.argv.index("--regions") + 1]) if "--regions" in sys.argv else 12
    dump = None
    if "--dump" in sys.argv:
        i = sys.argv.index("--dump")
        dump = (int(sys.argv[i + 1]), int(sys.argv[i + 2]))
    for name, hdr, rows in kernels(path):
        ix = {h: i for i, h in enumerate(hdr)}
        S, E = ix["# Samples"], ix["Instructions Executed"]
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[S]) for r in rows)
        inst = sum(int(r[E]) for r in rows)
        print("=" * 100)
        print(name[:110])
        print("instructions(SASS lines)=%d  warp-instr executed=%d  samples=%d" % (len(rows), inst, tot))
        st = {c: sum(int(r[ix[c]]) for r in rows) for c in stall_cols}
        print("stalls: " + "  ".join("%s=%.1f%%" % (c[6:], 100.0 * v / max(tot, 1)) for c, v in sorted(st.items(), key=lambda kv: -kv[1]) if v * 200 > tot))
        byop = defaultdict(lambda: [0, 0])
        for r in rows:
            op = r[ix["Source"]].split()
            op = [t for t in op if not t.startswith("@")]
            o = op[0].split(".")[0] if op else "?"
            byop[o][0] += int(r[S])
            byop[o][1] += int(r[E])
        print("by opcode (samples%, exec%): " + "  ".join("%s %.1f/%.1f" % (o, 100.0 * v[0] / max(tot, 1), 100.0 * v[1] / max(inst, 1))
                                                       for o, v in sorted(byop.items(), key=lambda kv: -kv[1][0])[:14]))
        # regions of equal execution count
        regs = []
        start = 0
        for i in range(1, len(rows) + 1):
            if i == len(rows) or rows[i][E] != rows[start][E]:
                regs.append((start, i))
                start = i
        regs = [(a, b, sum(int(rows[k][S]) for k in range(a, b)), int(rows[a][E])) for a, b in regs]
        print("top regions [first..last line) samples%  exec/instr  n_instr  dfma  lds  sts  stall mix")
        for a, b, s, e in sorted(regs, key=lambda t: -t[2])[:nreg]:
            ops = [rows[k][ix["Source"]] for k in range(a, b)]
            nd = sum("DFMA" in o or "DADD" in o or "DMUL" in o for o in ops)
            nl = sum(" LDS" in o or o.strip().startswith("LDS") for o in ops)
            ns = sum("STS" in o for o in ops)
            mix = {c: sum(int(rows[k][ix[c]]) for k in range(a, b)) for c in stall_cols}
            top = "  ".join("%s=%.0f%%" % (c[6:], 100.0 * v / max(s, 1)) for c, v in sorted(mix.items(), key=lambda kv: -kv[1])[:5])
            print("  [%5d..%5d) %5.1f%%  exec=%-9d n=%-4d dfma=%-4d lds=%-3d sts=%-3d %s" % (a, b, 100.0 * s / max(tot, 1), e, b - a, nd, nl, ns, top))
        if dump:
            for k in range(dump[0], min(dump[1], len(rows))):
                r = rows[k]
                mix = sorted(((int(r[ix[c]]), c[6:]) for c in stall_cols), reverse=True)[:3]
                print("%5d %6s %9s  %-60s %s" % (k, r[S], r[E], r[ix["Source"]].strip()[:60], " ".join("%s=%d" % (c, v) for v, c in mix if v)))


if __name__ == "__main__":
    main()
