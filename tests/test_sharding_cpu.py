"""CPU: multi-GPU host logic.  Frame / landmark partitioning and the landmark-sharded LM driver run with
world_size 2 over gloo (one all-reduce of the reduced camera system per LM trial) against the unsharded C oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_shard_partition(pkg):
    for n, w in [(10, 3), (64, 8), (5, 8), (0, 2), (128, 1)]:
        parts = [pkg.sharding.frame_shard(n, w, r) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [e - b for b, e in parts]
        assert max(sizes) - min(sizes) <= 1


def test_landmark_shards_balanced(pkg):
    p = pkg.synth.synth_ba_problem(3, 10, 500)
    for w in (1, 2, 4, 8):
        sh = pkg.sharding.landmark_shards(p["obs_point"], 500, w)
        assert sh[0][0] == 0 and sh[-1][1] == 500 and all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
        per = [int(((p["obs_point"] >= b) & (p["obs_point"] < e)).sum()) for b, e in sh]
        assert sum(per) == len(p["obs_point"])
        assert max(per) - min(per) <= 8       # a landmark has at most 6 observations


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import vslam_b200_loader
    from cpu_ba_session import CpuBaSession
    pkg = vslam_b200_loader.pkg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = pkg.synth.synth_ba_problem(21, 4, 60, outlier_frac=0.05)
    shards = pkg.sharding.landmark_shards(p["obs_point"], len(p["points"]), world)
    n1, n2, n3 = pkg.ffi.ba_reduce_sizes(4)
    r1, r2, r3 = (torch.zeros(n, dtype=torch.float64) for n in (n1, n2, n3))
    sess = CpuBaSession(p, shards[rank], r1, r2, r3)
    res = pkg.sharding.ba_optimize_sharded(sess, r1, r2, r3, 4, len(p["obs_pose"]), num_iterations=6,
                                           group=dist.group.WORLD)
    poses, pts, chi2, inl = sess.end()
    tp, tc, ti = torch.from_numpy(pts), torch.from_numpy(chi2), torch.from_numpy(inl.astype(np.int32))
    for t in (tp, tc, ti):
        dist.all_reduce(t)                     # shards are disjoint, zeros elsewhere
    if rank == 0:
        q.put((res, poses, tp.numpy(), tc.numpy(), ti.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_lm_gloo_world2_matches_oracle(pkg):
    import torch.multiprocessing as mp
    from oracle import ba_oracle as B
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res, poses, pts, chi2, inl = q.get(timeout=240)
    for pr in procs:
        pr.join(60)
        assert pr.exitcode == 0
    p = pkg.synth.synth_ba_problem(21, 4, 60, outlier_frac=0.05)
    o = B.optimize(p["poses"], p["points"], p["obs_pose"], p["obs_point"], p["obs_uv"], p["K"], num_iterations=6)
    assert res["iterations"] == o["iterations"] and res["trials"] == o["trials"] and res["accepted"] == o["accepted"]
    assert np.isclose(res["chi2_final"], o["chi2_final"], rtol=1e-9) and np.isclose(res["lambda_final"], o["lambda_final"], rtol=1e-9)
    assert np.allclose(poses, o["poses"], rtol=1e-8, atol=1e-10) and np.allclose(pts, o["points"], rtol=1e-8, atol=1e-10)
    assert np.allclose(chi2, o["chi2_per_obs"], rtol=1e-7, atol=1e-10)
    assert res["chi2_threshold"] == o["chi2_threshold"] and np.array_equal(inl.astype(bool), o["point_inlier"])
