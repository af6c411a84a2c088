#!/usr/bin/env python
"""Summarise the ncu launch list (`--metrics gpu__time_duration.sum --csv`) written by tools_profile.sh.
usage: python tools_launch_list.py gpurun_out/launches_<tag>.csv profiles/<name>.json ["capture note"]"""
import collections
import csv
import json
import sys


def main():
    src, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt, seq = collections.Counter(), collections.Counter(), collections.defaultdict(list)
    for r in rows[1:]:
        v = float(r[iv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        k = r[ik].split("(")[0][:60]
        tot[k] += v
        cnt[k] += 1
        seq[k].append(v)
    total = sum(tot.values())
    d = {"capture": note, "total_ms": round(total, 3), "kernels": []}
    for k, v in tot.most_common():
        e = {"kernel": k, "launches": cnt[k], "total_ms": round(v, 3), "share_pct": round(100 * v / total, 2)}
        if "k_epoch" in k and cnt[k] == 94:                  # graded schedule of 4 batches: 9 + 17 + 34 + 34 launches
            s = seq[k]
            e["mean_ms_by_level"] = {"M=9": round(sum(s[:9]) / 9, 4), "M=17": round(sum(s[9:26]) / 17, 4),
                                     "M=34": round(sum(s[26:]) / 68, 4)}
        d["kernels"].append(e)
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps(d["kernels"][:3], indent=1))


if __name__ == "__main__":
    main()
