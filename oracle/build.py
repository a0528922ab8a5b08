"""Compile the CPU parity oracle (test infrastructure) into oracle/_build/.

The reference is Python + numba: there is no C/C++ reference source to compile into
``oracle/_ref/``, so the oracle is a port (``cpu_baseline.kind == "port"``).
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "libstratego_oracle.so")
CFLAGS = ["-O2", "-std=gnu99", "-fPIC", "-pthread", "-ffp-contract=off", "-fno-fast-math", "-Wall", "-Wextra"]


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "stratego_oracle.c")
    hdr = os.path.join(HERE, "stratego_oracle.h")
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return OUT
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else (shutil.which("gcc") or shutil.which("cc"))
    if cc is None:
        raise RuntimeError("no C compiler found for the parity oracle")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    tmp = OUT + ".tmp.%d" % os.getpid()
    subprocess.run([cc, *CFLAGS, "-shared", "-o", tmp, src], check=True)
    os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
