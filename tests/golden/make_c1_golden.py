"""Generates tests/golden/c1_trace_literal.npz: the oracle's config-1 run (reference
test/main.cpp loop, noise switches off, REF_LITERAL quirks, FP64), sub-sampled.
The reference ships no golden vectors, so this fixture pins OUR oracle against
regressions; it is produced by the oracle itself:   python tests/golden/make_c1_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_py  # noqa: E402

STRIDE = 250
tape = oracle_py.sim_tape(noise_seed=0)
o = oracle_py.OracleEKF(0)
o.reset(np.zeros(3), np.zeros((3, 3)))
rows = []
oracle_py.run_tape(o, tape, dense_heading=False, on_step=lambda s, f: rows.append(np.concatenate([f.X[:3], [f.n]])))
rows = np.asarray(rows)
np.savez_compressed(os.path.join(HERE, "c1_trace_literal.npz"), steps=tape["steps"], stride=STRIDE,
                    trace=rows[::STRIDE], X_final=o.X, P_diag_final=np.diag(o.P),
                    obs_steps=int(tape["obs_flag"].sum()), tags_seen=np.unique(tape["tags"]))
print("steps", tape["steps"], "n", o.n, "final pose", o.X[:3])
