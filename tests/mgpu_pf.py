"""Multi-GPU particle-filter parity (imported by mgpu_worker.py): `world` ranks x Pl particles on the
GPUs against the CPU oracle running all world*Pl particles; resampled indices must match exactly."""
import os

import numpy as np


def pf_checks():
    import torch
    import torch.distributed as dist

    import conan_slam_b200 as cs
    import helpers
    import oracle_py
    from conan_slam_b200 import dist as cd
    from helpers import QE, rel_err
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    R2 = 2 * helpers.R_BASE
    for Pl, nfeat in ((32, 3), (4096, 6), (33 * 1024, 4)):
        P = Pl * world
        sl = slice(rank * Pl, (rank + 1) * Pl)
        nid = cd.nccl_unique_id(device=f"cuda:{local}")
        rng = np.random.default_rng(Pl)           # same stream on every rank
        g = cs.PF(num_particles=Pl, capacity_landmarks=nfeat + 2, device=local, flags=cs.FLAG_INTENDED, rank=rank,
                  world=world, nccl_id=nid)
        o = oracle_py.OraclePF(P, cs.FLAG_INTENDED)
        lm = rng.uniform(-900, 900, size=(2, nfeat))
        for f in (g, o):
            for k in range(6):
                f.predict(83.33, 0.03, QE, 73.0, 0.01)
                f.observeHeading(0.0005 * (k + 1), True)
        X = o.poses
        Z0 = np.stack([np.hypot(lm[0] - X[0, 0], lm[1] - X[0, 1]),
                       np.arctan2(lm[1] - X[0, 1], lm[0] - X[0, 0]) - X[0, 2]])
        xi0 = rng.normal(size=(P, 3))
        g.samplePose(xi0[sl])
        o.samplePose(xi0)
        for f in (g, o):
            f.addOneNewFeature(Z0, R2)
            for k in range(6):
                f.predict(83.33, -0.02, QE, 73.0, 0.01)
                f.observeHeading(0.004 + 0.0005 * k, True)
        base = np.array([[4e-4, 1e-5, 1e-7], [1e-5, 5e-4, -2e-7], [1e-7, -2e-7, 3e-8]])
        covs = np.tile(base.reshape(-1), (P, 1)) * (1.0 + 0.1 * rng.uniform(size=(P, 1)))
        poses = o.poses
        g.set_poses(poses[sl], covs[sl])
        o.set_poses(poses, covs)
        ids = np.array([2, 1], dtype=np.int32)
        Z = np.stack([np.hypot(lm[0, ids - 1] - poses[0, 0], lm[1, ids - 1] - poses[0, 1]) + 0.004,
                      np.arctan2(lm[1, ids - 1] - poses[0, 1], lm[0, ids - 1] - poses[0, 0]) - poses[0, 2] + 2e-6])
        xi = rng.normal(size=(P, 3))
        g.sampleProposal(Z, ids, R2, xi[sl])
        o.sampleProposal(Z, ids, R2, xi)
        for f in (g, o):
            f.featureUpdate(Z, ids, R2)
        wo = o.weights
        assert np.max(np.abs(g.weights - wo[sl]) / wo[sl]) < 1e-9
        # skew the weights so that survivors cross rank boundaries
        w = rng.uniform(0.0, 1.0, size=P) ** 4 + 1e-12
        w[: P // 3] *= 50.0
        g.weights = w[sl]
        o.weights = w
        u = rng.normal(size=P) * 0.3
        kg, ng, dg = g.resampleParticles(float("inf"), u[sl], True)
        ko, no, do = o.resampleParticles(float("inf"), u, True)
        assert dg and do
        assert np.array_equal(kg, ko[sl]), (Pl, rank, np.flatnonzero(kg != ko[sl])[:5])
        assert abs(ng - no) <= 1e-12 * no
        cross = int(np.sum((kg // Pl) != rank))
        assert rel_err(g.poses, o.poses[sl]) < 1e-12
        for p in (0, Pl // 2, Pl - 1):
            XFg, PFg = g.features(p)
            XFo, PFo = o.features(rank * Pl + p)
            assert rel_err(XFg, XFo) < 1e-12 and rel_err(PFg, PFo) < 1e-12
        assert np.all(g.weights == 1.0 / P)
        print(f"rank {rank}: PF {world}x{Pl} ok, {cross} of {Pl} survivors fetched from peers")
        g.close()
    print(f"rank {rank}: sharded PF parity ok")
