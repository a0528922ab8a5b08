#!/usr/bin/env python
"""Per-source-line stall samples and executed instructions from an ncu report captured with
--import-source on (kernel compiled with -lineinfo).  usage: tools/ncu_source_lines.py <rep> [top_n]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
fname, hdr = None, None
samples = collections.Counter(); execd = collections.Counter(); text = {}
for row in csv.reader(io.StringIO(raw)):
    if not row:
        continue
    if row[0] == "File Path":
        fname = row[1].split("/")[-1]; continue
    if row[0] == "Function Name":
        continue
    if row[0] == "Line No":
        hdr = row; si = hdr.index("# Samples"); ei = hdr.index("Instructions Executed"); continue
    if hdr and row[0] not in ("", "-") and row[0].isdigit():
        key = (fname, int(row[0]))
        try:
            samples[key] += int(row[si]); execd[key] += int(row[ei])
        except ValueError:
            pass
        text[key] = row[1].strip()[:90]
ts, te = sum(samples.values()), sum(execd.values())
print("total samples %d, instructions executed %d" % (ts, te))
print("%-22s %8s %6s %10s %6s  %s" % ("line", "samples", "%", "instr", "%", "source"))
for key, v in samples.most_common(top):
    print("%-22s %8d %5.1f%% %10d %5.1f%%  %s" % ("%s:%d" % key, v, 100.0 * v / ts, execd[key], 100.0 * execd[key] / te, text[key]))
