"""Cut the SASS of one kernel (ncu -i X.ncu-rep --page source --csv, optionally gzipped) into runs of equal execution
count and print, per run, its share of the executed warp instructions and of the stall samples.

    python tools/ncu_exec_segments.py source.csv[.gz] [warps_of_the_grid] [top_n]

Used for profiles/r2_ncu_fwt_inverse.txt: where the instructions of a tile kernel go outside its item loop."""
import csv
import gzip
import sys


def main():
    path = sys.argv[1]
    warps = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
    rows = list(csv.reader(f))[2:]
    ex = [int(r[5]) for r in rows]
    sm = [int(r[4]) for r in rows]
    tot, ts = sum(ex), sum(sm)
    print("total warp instructions %d, stall samples %d" % (tot, ts))
    segs, i = [], 0
    while i < len(rows):
        j = i
        while j + 1 < len(rows) and abs(ex[j + 1] - ex[i]) <= 0.02 * max(ex[i], 1):
            j += 1
        segs.append((i, j + 1, ex[i], sum(ex[i:j + 1]), sum(sm[i:j + 1])))
        i = j + 1
    segs.sort(key=lambda s: -s[3])
    for s in segs[:top]:
        per_warp = ("  per-warp=%6.1f" % (s[3] / warps)) if warps else ""
        print("  [%4d..%4d) n=%3d exec/instr=%9d  instr=%5.1f%%  samples=%5.1f%%%s  first: %s"
              % (s[0], s[1], s[1] - s[0], s[2], 100.0 * s[3] / tot, 100.0 * s[4] / max(ts, 1), per_warp,
                 rows[s[0]][1].strip()[:60]))


if __name__ == "__main__":
    main()
