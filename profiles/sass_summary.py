#!/usr/bin/env python
"""Counts the Blackwell-specific SASS mnemonics of every kernel in libeqnn_b200.so -> profiles/sass_summary.txt.

UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP / UBLKPF = cp.async.bulk (TMA, non-tensor form) and its
L2 prefetch, UTCBAR = tcgen05.commit, HMMA would be the legacy mma.sync path (must stay 0).  Run by
__graft_entry__.build() whenever the library is rebuilt; the file is tracked so that the evidence is in the tree."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "..", "eqnn_jax_b200", "libeqnn_b200.so")
OUT = os.path.join(HERE, "sass_summary.txt")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UBLKPF", "UTMALDG", "SYNCS", "HMMA", "MUFU", "FFMA2", "FMUL2"]


def main():
    if os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(LIB) and "--force" not in sys.argv:
        return
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    rows, name, counts = [], None, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                rows.append((name, counts))
            name, counts = m.group(1), dict.fromkeys(KEYS, 0)
            continue
        if name:
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                op = m.group(1)
                for k in KEYS:
                    if op.startswith(k) and not (k == "HMMA" and op.startswith("HMMA") is False):
                        counts[k] += 1
                        break
    if name:
        rows.append((name, counts))
    dem = subprocess.run(["cu++filt"] + [r[0] for r in rows], capture_output=True, text=True)
    names = dem.stdout.splitlines() if dem.returncode == 0 and dem.stdout else [r[0] for r in rows]
    with open(OUT, "w") as fh:
        fh.write("# cuobjdump -sass eqnn_jax_b200/libeqnn_b200.so (sm_100a): instruction counts per kernel\n")
        fh.write("# " + " ".join(f"{k:>8s}" for k in KEYS) + "  kernel\n")
        for (raw, c), n in sorted(zip(rows, names), key=lambda t: -(t[0][1]["UTCHMMA"] + t[0][1]["LDTM"])):
            short = re.sub(r"\(anonymous namespace\)::|<unnamed>::|\((?:int|bool)\)", "", n)
            short = re.sub(r"\((?!.*>).*$", "", short)      # drop the argument list (the last parenthesis group)
            fh.write("  " + " ".join(f"{c[k]:8d}" for k in KEYS) + "  " + short + "\n")
    print(f"wrote {OUT} ({len(rows)} kernels)")


if __name__ == "__main__":
    main()
