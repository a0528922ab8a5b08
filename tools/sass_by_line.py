#!/usr/bin/env python
"""SASS instructions per top-level source line of one kernel (which source statement costs how much code).
usage: tools/sass_by_line.py <kernel substring> [so]"""
import re, subprocess, sys, collections, tempfile, os, glob
pat = sys.argv[1]
so = sys.argv[2] if len(sys.argv) > 2 else "stratego_env_b200/csrc/libstratego_b200.so"
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
txt = subprocess.run(["nvdisasm", "--print-line-info-inline"] + glob.glob(d + "/*.cubin"), capture_output=True, text=True).stdout
fn, cur, inner = None, None, None
top = collections.Counter(); leaf = collections.Counter()
for line in txt.splitlines():
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', line)
    if m:
        fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
    if m:
        leafk = (os.path.basename(m.group(1)), int(m.group(2)))
        if m.group(3):
            # keep the outermost frame seen in a run of inline records
            cur_candidate = (os.path.basename(m.group(3)), int(m.group(4)))
            if inner is None:
                inner = leafk
            cur = cur_candidate
        else:
            if inner is None:
                inner = leafk
            cur = cur if False else leafk if inner == leafk else cur
            cur = leafk if cur is None else cur
        pending = True
        continue
    if fn and pat in fn and re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s', line):
        top[cur] += 1
        leaf[inner] += 1
        inner = None
print("total", sum(top.values()))
print("--- by outermost frame")
for k, v in top.most_common(45):
    print("%6d  %s:%d" % (v, k[0], k[1]) if k else "%6d  ?" % v)
print("--- by leaf line")
for k, v in leaf.most_common(30):
    print("%6d  %s:%d" % (v, k[0], k[1]) if k else "%6d  ?" % v)
