"""Where the shared loads sit between the DFMAs of the big straight-line segments of a kernel (cuobjdump -sass text).

    python tools/sass_lds_spread.py file.sass [name-substring] [min_dfma]
"""
import re
import sys

txt = open(sys.argv[1]).read()
sub = sys.argv[2] if len(sys.argv) > 2 else ""
mind = int(sys.argv[3]) if len(sys.argv) > 3 else 100
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n')[0]
    if sub not in name:
        continue
    lines = [l for l in f.split('\n') if re.search(r'/\*[0-9a-f]{4}\*/', l)]
    ops = [re.sub(r'/\*[0-9a-f]+\*/', '', l).strip().split(';')[0] for l in lines]
    print(name[-70:], 'instr', len(ops), 'DFMA', sum('DFMA' in o for o in ops))
    start = 0
    for i, o in enumerate(ops + ['BRA']):
        if re.search(r'\b(BRA|EXIT|BAR|CALL|RET)\b', o):
            seg = ops[start:i]
            nd = sum('DFMA' in x for x in seg)
            if nd >= mind:
                lds = [k for k, x in enumerate(seg) if re.search(r'\bLDS', x)]
                print('  segment [%d..%d) n=%d dfma=%d lds=%d  LDS at %s' % (start, i, len(seg), nd, len(lds), lds))
            start = i + 1
