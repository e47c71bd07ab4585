"""FP64 / spill / shared / global instruction counts per 200-instruction window of one kernel's SASS.
usage: python scripts/sass_regions.py lib.so mangled-name-substring"""
import re, subprocess, sys
txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n')[0]
    if sys.argv[2] not in name:
        continue
    ins = [l for l in f.split('\n') if re.search(r'/\*[0-9a-f]{4,5}\*/\s+\S', l)]
    print(name, len(ins), "instructions")
    for i in range(0, len(ins), 200):
        w = ins[i:i + 200]
        c = lambda pat: sum(1 for l in w if re.search(pat, l))
        print(f"{i:5d} fp64 {c(r'D(FMA|MUL|ADD|SETP)'):3d} spill {c(r'(LDL|STL)'):3d} lds/sts {c(r'(LDS|STS)'):3d} ldg/stg {c(r'(LDG|STG)'):3d} "
              f"imad/mov {c(r'(IMAD|MOV)'):3d} bra {c(r'BRA'):3d} call {c(r'CALL'):2d} syncs {c(r'SYNCS'):2d}")
