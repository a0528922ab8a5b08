#!/usr/bin/env python
"""Static footprint of the HOT part of a kernel from an ncu report (--set full --import-source on): how many SASS
instructions (16 B each) are executed at least once per `frac` of the games, and where the no_instruction (instruction
fetch) stall samples sit.  The SM's instruction caches are small (L0 ~6 KB per scheduler, L1.5 32 KB per SM), so a
per-game path longer than that refetches from L2 every game.  usage: tools/sass_hot_footprint.py <rep> <games> [frac]"""
import csv, io, subprocess, sys
rep, games = sys.argv[1], int(sys.argv[2])
frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if r and r[0] == "Address")
ie, ns = hdr.index("Instructions Executed"), hdr.index("# Samples")
noinst = [i for i, h in enumerate(hdr) if "no_instruction" in h.lower() or "No Instruction" in h]
body = [r for r in rows if r and r[0].startswith("0x")]
ex = [int(r[ie]) for r in body]
total = len(body)
for f in (0.9, frac, 0.1, 0.01):
    n = sum(1 for e in ex if e >= f * games)
    print("instructions executed >= %.2f x games: %5d (%.1f KB) of %d (%.1f KB)" % (f, n, n * 16 / 1024, total, total * 16 / 1024))
print("dynamic instructions per game: %.0f" % (sum(ex) / games))
# address span of the hot instructions: a hot path scattered over a wide range touches more cache lines
hot = [i for i, e in enumerate(ex) if e >= frac * games]
lines = {i // 8 for i in hot}
print("hot instructions span indices %d..%d, distinct 128-byte lines: %d (%.1f KB)" % (hot[0], hot[-1], len(lines), len(lines) * 128 / 1024))
if noinst:
    tot = sum(int(r[noinst[0]] or 0) for r in body)
    print("no_instruction samples: %d of %d" % (tot, sum(int(r[ns] or 0) for r in body)))
