"""Worker for the multi-process tests (launched by torch.distributed.run, one rank per GPU —
or, with --cpu, gloo ranks that only exercise the host-side rendezvous / sharding logic)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def cpu_checks():
    import torch.distributed as dist

    from conan_slam_b200 import dist as cd
    rank, world = dist.get_rank(), dist.get_world_size()
    payload = bytes(range(128)) if rank == 0 else b""
    got = cd.broadcast_bytes(payload, 128, src=0)
    assert got == bytes(range(128)), "broadcast_bytes corrupted the NCCL id blob"
    for n in (3, 127, 128, 129, 1000, 4003):
        rows = cd.shard_rows(n, rank, world)
        allrows = [None] * world
        dist.all_gather_object(allrows, rows)
        flat = sorted(r for rr in allrows for r in rr)
        assert flat == list(range(n)), "every covariance row must be stored by exactly one rank"
        loc = [cd.shard_local_row(r, world) for r in rows]
        assert loc == sorted(loc) and len(set(loc)) == len(loc)
        if rows:
            assert max(loc) < ((n + 127) // 128 + world - 1) // world * 128
    print(f"rank {rank}: cpu checks ok")


def ekf_checks():
    import torch
    import torch.distributed as dist

    import conan_slam_b200 as cs
    import helpers
    import oracle_py
    from conan_slam_b200 import dist as cd
    from helpers import QE, RE, rel_err
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    for N, flags in ((150, 0), (700, cs.FLAG_INTENDED)):
        nid = cd.nccl_unique_id(device=f"cuda:{local}")
        X, P, lm = helpers.synthetic_map(N, 900 + N)
        g = cs.EKF(capacity_landmarks=N + 8, device=local, flags=flags, rank=rank, world=world, nccl_id=nid)
        o = oracle_py.OracleEKF(flags)
        g.reset(X, P)
        o.reset(X, P)
        rng = np.random.default_rng(N)
        # observed landmarks lie inside the reference's sensor range (slam.h:79): far landmarks make S cancel by
        # many orders of magnitude and any re-association of the arithmetic visible at the 1e-9 level
        near = np.argsort(np.hypot(lm[0] - X[0], lm[1] - X[1]))[:32]

        def check(tag):
            ex = rel_err(g.X, o.X)
            Pg = g.P  # collective
            iu = np.triu_indices(o.n)
            ep = rel_err(Pg[iu], o.P[iu])
            assert g.n == o.n and ex < 1e-9 and ep < 1e-9, (tag, N, rank, ex, ep)

        check("reset")
        for f in (g, o):
            f.predict(83.33, 0.02, QE, 73.0, 0.01)
            f.observeHeading(float(X[2]) + 1e-4, True)
        check("predict+heading")
        ids = (rng.choice(near, size=5, replace=False) + 1).astype(np.int32)
        Z = helpers.observe(o.X, lm, ids, rng)
        jg = g.gate(Z, RE, 50.0, 1000.0)[0]
        jo = o.gate(Z, RE, 50.0, 1000.0)[0]
        assert np.array_equal(jg, jo) and np.array_equal(jg, ids), (jg, jo, ids)
        for f in (g, o):
            f.update(Z, RE, ids, False)
        check("sequential update")
        # the replicated diagonal-block cache FOLLOWED the grouped update (no re-pack): same decisions, same nd
        def same_gate(Zq=None):
            Zq = Z if Zq is None else Zq
            gg, go = g.gate(Zq, RE, 50.0, 1000.0), o.gate(Zq, RE, 50.0, 1000.0)
            hit = go[0] != 0
            assert np.array_equal(gg[0], go[0]) and np.array_equal(np.isinf(gg[2]), np.isinf(go[2])), (gg, go)
            assert not hit.any() or rel_err(gg[2][hit], go[2][hit]) < 1e-9, (gg[2], go[2])
            assert hit.all() or np.allclose(gg[3][~hit], go[3][~hit], rtol=1e-9, atol=0), (gg[3], go[3])

        same_gate()
        for f in (g, o):
            f.observeHeading(float(o.X[2]) + 5e-5, True)
            f.update(Z[:, :1], RE, ids[:1], False)
        check("heading + single update (cache follows rank-1 and rank-2 passes)")
        same_gate()
        # fused scan on the sharded handle: 10 observations (two groups), one of them passing no gate
        ids_s = (rng.choice(near, size=9, replace=False) + 1).astype(np.int32)
        Zs = np.concatenate([helpers.observe(o.X, lm, ids_s, rng)[:, :5], np.array([[9000.0], [1.0]]),
                             helpers.observe(o.X, lm, ids_s, rng)[:, 5:]], axis=1)
        jo = o.gate(Zs, RE, 50.0, 1000.0)[0]
        o.update(Zs[:, jo != 0], RE, jo[jo != 0], False)
        jg, _ = g.scan(Zs, RE, 50.0, 1000.0)
        assert np.array_equal(jg, jo) and jo[5] == 0, (jg, jo)
        check("fused scan (sharded)")
        same_gate(helpers.observe(o.X, lm, (rng.choice(near, size=6, replace=False) + 1).astype(np.int32), rng))
        ids2 = (rng.choice(near, size=16, replace=False) + 1).astype(np.int32)
        Z2 = helpers.observe(o.X, lm, ids2, rng)
        for f in (g, o):
            f.update(Z2, RE, ids2, True)
        check("joint update (DMMA, sharded)")
        Zn = np.array([[700.0, 1200.0, 300.0], [0.4, -0.9, 2.0]])
        phi_meas = float(o.X[2]) - 2e-4
        for f in (g, o):
            f.augment(Zn, RE)
            f.predict(83.33, -0.01, QE, 73.0, 0.01)
            f.observeHeading(phi_meas, True)
        check("augment")
        # k control steps in one call: per step predict + heading gain + eager rows 0..2 / cache update,
        # the k heading passes over the rest of the sharded covariance merged into one
        kk = 5
        sw = 0.02 * np.sin(np.arange(kk))
        ph = float(o.X[2]) + np.cumsum(83.33 * 0.01 * np.sin(sw) / 73.0) + 1e-4 * np.cos(np.arange(kk))
        tr = g.controlSteps(np.full(kk, 83.33), sw, ph, True, QE, 73.0, 0.01)
        for i in range(kk):
            o.predict(83.33, sw[i], QE, 73.0, 0.01)
            o.observeHeading(ph[i], True)
        assert rel_err(tr[-1], o.X[:3]) < 1e-9
        check("control steps (merged heading passes)")
        same_gate(helpers.observe(o.X, lm, (rng.choice(near, size=6, replace=False) + 1).astype(np.int32), rng))
        ids3 = np.array([N + 1, N + 3, 7], dtype=np.int32)
        Z3 = np.stack([np.array([700.0, 300.0, Z[0, 0]]), np.array([0.4, 2.0, Z[1, 0]])])
        for f in (g, o):
            f.update(Z3[:, :2], RE, ids3[:2], False)
        check("update of fresh landmarks")
        # numerically skipped updates (slam.h:252-255) are counted, not failures; 1 << 20 would be a peer that never
        # arrived at a column exchange
        assert g.sync() < (1 << 20)
        # every rank holds the same replicated state bit for bit
        xs = [None] * world
        dist.all_gather_object(xs, g.X.tobytes())
        assert all(x == xs[0] for x in xs), "replicated X diverged across ranks"
        # sharded checkpoint: every rank writes / reads its own rows (<path>.r<rank>of<world>)
        path = f"/tmp/cslam_sharded_ckpt_{N}"
        g.save(path)
        nid2 = cd.nccl_unique_id(device=f"cuda:{local}")
        g2 = cs.EKF(capacity_landmarks=N + 8, device=local, flags=flags, rank=rank, world=world, nccl_id=nid2)
        g2.load(path)
        assert g2.n == g.n and np.array_equal(g2.X, g.X)
        assert np.array_equal(np.triu(g2.P), np.triu(g.P)), "sharded checkpoint round trip changed the covariance"
        g2.close()
        g.close()
    print(f"rank {rank}: sharded EKF parity ok")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--what", default="ekf")
    args = ap.parse_args()
    import torch.distributed as dist
    if args.cpu:
        dist.init_process_group("gloo")
        cpu_checks()
    else:
        import torch
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        if args.what in ("ekf", "all"):
            ekf_checks()
        if args.what in ("pf", "all"):
            from mgpu_pf import pf_checks
            pf_checks()
    dist.barrier()
    dist.destroy_process_group()
