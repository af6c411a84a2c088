#!/usr/bin/env python
"""Decode the scoreboard fields of the SASS control words of one kernel (sm_90 / sm_100 128-bit encoding: in the upper
64-bit word, bits 41-44 stall count, 45 yield, 46-48 write-barrier index, 49-51 read-barrier index, 52-57 wait mask).
Lists every long-latency instruction (LDG / LDGSTS / REDG / LDS / SHFL / MUFU) with the scoreboard it signals and every
instruction that waits, and counts the global loads per scoreboard.

usage: python tools_sass_scoreboards.py <library.so> <mangled kernel name> [> profiles/...txt]"""
import collections
import re
import subprocess
import sys


def main():
    lib, fun = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, lib], capture_output=True, text=True).stdout.split("\n")
    pat = re.compile(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/")
    pat2 = re.compile(r"^\s+/\* (0x[0-9a-f]{16}) \*/")
    ins, i = [], 0
    while i < len(txt):
        m = pat.match(txt[i])
        m2 = pat2.match(txt[i + 1]) if m and i + 1 < len(txt) else None
        if m and m2:
            hi = int(m2.group(1), 16)
            ins.append((int(m.group(1), 16), m.group(2).strip(), (hi >> 41) & 0xF, (hi >> 46) & 7, (hi >> 49) & 7, (hi >> 52) & 0x3F))
            i += 2
        else:
            i += 1
    per_sb = collections.Counter()
    print(f"# {fun}: {len(ins)} instructions")
    for addr, s, stall, wb, rb, wm in ins:
        op = s.split()[1] if s.startswith("@") else s.split()[0]
        is_gld = op.startswith("LDG") and not op.startswith("LDGDEPBAR")
        if is_gld:
            per_sb[wb] += 1
        longlat = op.split(".")[0] in ("LDG", "LDGSTS", "REDG", "LDS", "SHFL", "MUFU", "LDL", "ATOMG", "DEPBAR", "LDGDEPBAR")
        if longlat or wm:
            tag = (f" W{wb}" if wb != 7 else "") + (f" R{rb}" if rb != 7 else "")
            if wm:
                tag += " wait[" + ",".join(str(b) for b in range(6) if wm >> b & 1) + "]"
            if longlat or (wm >> 5 & 1):
                print(f"{addr:05x}  {s[:78]:78s} stall{stall}{tag}")
    print("# global loads (LDG / LDGSTS) per write scoreboard:", dict(per_sb))


if __name__ == "__main__":
    main()
