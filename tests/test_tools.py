"""The small analysis tools under tools/ (CPU tier): they turn raw profiler / timestamp dumps into the summaries
committed under profiles/, so they get the same treatment as the rest — a known input, a checked output."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ktrace_report_reads_the_dump_of_a_handle(tmp_path):
    # three update groups, 8 stamps each (ns): snapshot in/out, group-gain in/out, three phase marks, spare
    t0 = 1_000_000_000
    rows = []
    for k in range(3):
        b = t0 + 100_000 * k
        rows.append([b, b + 4_000, b + 9_000, b + 25_000, b + 10_000, b + 21_000, b + 15_000, 0])
    f = tmp_path / "kt.r0"
    np.asarray(rows, dtype=np.uint64).tofile(f)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ktrace_report.py"), str(f)], capture_output=True,
                         text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines()
    assert lines[0].startswith("3 groups")
    first = lines[1].split()
    # snapshot kernel 4.0 us, gap 5.0 us, group-gain kernel 16.0 us
    assert [float(first[1]), float(first[2]), float(first[3])] == [4.0, 5.0, 16.0]
    assert "period 100.0" in lines[-1]


def test_launch_summary_aggregates_an_ncu_launch_list():
    src = os.path.join(ROOT, "profiles", "launches_r02_ekf20k.csv")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), src], capture_output=True,
                         text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    body = out.stdout.splitlines()[1:]
    assert body and "k_cov_update_tma_dense" in body[0]      # the covariance pass leads the list ...
    share = sum(float(l.split()[-1].rstrip("%")) for l in body if "k_cov_update" in l)
    assert share > 90.0                                        # ... with > 90 % of the GPU time of the loop
