#!/usr/bin/env python
"""Attribute the executed SASS instructions (and stall samples) of one kernel in an ncu report to CUDA source lines.
usage: python tools_sass_lines.py <report.ncu-rep> <kernel regex> [library.so] [top N]
Joins `ncu --page source --csv` (per-instruction counters, in program order) with `nvdisasm -g` line markers of the
same cubin, by instruction index inside the function."""
import csv
import collections
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(lib, mangled_hint):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
    out = {}
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        func, cur, lines = None, None, []
        for ln in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                if func:
                    out[func] = lines
                func, cur, lines = m.group(1), None, []
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if func and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
                lines.append(cur)
        if func:
            out[func] = lines
    return out


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    lib = sys.argv[3] if len(sys.argv) > 3 else "annembed_b200/libannembed_cuda.so"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    kname = rows[h - 1][1]
    hdr = rows[h]
    ie, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    inst = []
    for r in rows[h + 1:]:
        if not r or r[0] in ("Address", "Kernel Name"):
            break
        inst.append((int(r[ie]), int(r[isamp]), r[isrc]))
    table = sass_lines(lib, kname)
    # pick the function whose instruction count matches
    cands = [f for f, l in table.items() if len(l) == len(inst)]
    base = re.sub(r"<.*", "", kname.split("(")[0].split()[-1])
    cands = [f for f in cands if base in f] or cands
    if not cands:
        sys.exit(f"no function with {len(inst)} instructions in {lib} (rebuilt since the capture?)")
    lines = table[cands[0]]
    by_e, by_s = collections.Counter(), collections.Counter()
    for (e, s, _), loc in zip(inst, lines):
        by_e[loc] += e
        by_s[loc] += s
    tot_e, tot_s = sum(by_e.values()), sum(by_s.values())
    print(f"{kname}\n  {cands[0]}: {len(inst)} SASS instructions, {tot_e} executed, {tot_s} samples")
    src = {}
    for loc, e in by_e.most_common(top):
        if loc and loc[0] not in src:
            for d in ("annembed_b200/csrc", "."):
                p = os.path.join(d, loc[0])
                if os.path.exists(p):
                    src[loc[0]] = open(p).read().splitlines()
        text = src.get(loc[0], [""] * 10**6)[loc[1] - 1].strip()[:90] if loc else "?"
        print(f"{100 * e / tot_e:5.1f}% inst {100 * by_s[loc] / max(1, tot_s):5.1f}% samp  {loc[0] if loc else '?'}:{loc[1] if loc else 0:<5d} {text}")


if __name__ == "__main__":
    main()
