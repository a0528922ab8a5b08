#!/usr/bin/env python
"""SASS instruction count / code bytes per function of the built extension (I-cache budgeting aid)."""
import re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "stratego_env_b200/csrc/libstratego_b200.so"
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn, cnt = None, {}
for l in out.splitlines():
    m = re.search(r'Function : (\S+)', l)
    if m:
        fn = m.group(1); cnt[fn] = 0
    elif fn and re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s', l):
        cnt[fn] += 1
for k, v in cnt.items():
    if flt in k:
        print("%6d instr %5.1f KB  %s" % (v, v * 16 / 1024, k[:110]))
