#!/usr/bin/env python3
"""Read the kernel timestamps a handle dumps with CSLAM_KTRACE=<file> (one file per rank: <file>.r<rank>): per update
group 8 stamps (snapshot in/out, group-gain in/out, three phase marks of block 0, spare) in globaltimer ns.

    CSLAM_KTRACE=/tmp/kt python bench.py --no-extras ...; python tools/ktrace_report.py /tmp/kt.r0 [first] [count]
"""
import sys

import numpy as np


def main():
    t = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 8).astype(np.int64)
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    count = int(sys.argv[3]) if len(sys.argv) > 3 else len(t) - first
    t = t[first:first + count]
    snap = (t[:, 1] - t[:, 0]) / 1e3
    gap = (t[:, 2] - t[:, 1]) / 1e3
    grp = (t[:, 3] - t[:, 2]) / 1e3
    nxt = (t[1:, 0] - t[:-1, 3]) / 1e3
    period = (t[1:, 0] - t[:-1, 0]) / 1e3
    ph0 = (t[:, 4] - t[:, 2]) / 1e3
    ph1 = (t[:, 5] - t[:, 4]) / 1e3
    ph1b = (t[:, 6] - t[:, 5]) / 1e3
    ph2 = (t[:, 3] - t[:, 6]) / 1e3
    print(f"{len(t)} groups; us: snapshot kernel, gap, group-gain kernel [block 0: header+flags, marginal replay, wait rows, rows->last block out], "
          "group out -> next snapshot in (gate + launches), period")
    for i in range(len(t)):
        print(f"{first + i:5d} {snap[i]:8.1f} {gap[i]:8.1f} {grp[i]:8.1f} [{ph0[i]:6.1f} {ph1[i]:6.1f} {ph1b[i]:6.1f} {ph2[i]:6.1f}] "
              + (f"{nxt[i]:9.1f} {period[i]:9.1f}" if i + 1 < len(t) else ""))
    sel = period < 1000
    print("median (period < 1 ms):", f"snapshot {np.median(snap):.1f}  gap {np.median(gap):.1f}  group {np.median(grp):.1f}  "
          f"to-next {np.median(nxt[sel]):.1f}  period {np.median(period[sel]):.1f}")


if __name__ == "__main__":
    main()
