"""GPU: K17 landmark-sharded BA session.  world = 1 through the driver, and two shards emulated on one GPU (two
contexts, two driver threads, an in-process sum standing in for the NCCL all-reduce) against the unsharded kernel and
the C oracle.  The real 2/4/8-GPU NCCL run is tools/run_ba_sharded.py under torchrun."""
import threading

import numpy as np
import pytest

from oracle import ba_oracle as B

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4


class ThreadGroup:
    """all-reduce between driver threads of one process (sum / max over the participants' tensors)"""

    def __init__(self, n):
        self.n, self.bar, self.slots, self.lock = n, threading.Barrier(n), {}, threading.Lock()

    def size(self):
        return self.n

    def all_reduce(self, t, op):
        import torch
        me = threading.get_ident()
        with self.lock:
            self.slots[me] = t
        self.bar.wait()
        ts = list(self.slots.values())
        red = torch.stack([x.to(ts[0].device) for x in ts]).sum(0) if op == "sum" else torch.stack([x.to(ts[0].device) for x in ts]).max(0).values
        self.bar.wait()
        t.copy_(red.to(t.device))
        self.bar.wait()
        with self.lock:
            self.slots.pop(me, None)
        self.bar.wait()


def _run_rank(pkg, p, shard, group, nit, out, idx):
    import torch
    K = len(p["poses"])
    ctx = pkg.Context(device=0, max_images=0, max_width=0, max_height=0, max_keypoints=1, max_ba_poses=64,
                      max_ba_points=32768, max_ba_obs=262144)
    n1, n2, n3 = pkg.ffi.ba_reduce_sizes(K)
    r1, r2, r3 = (torch.zeros(n, dtype=torch.float64, device="cuda:0") for n in (n1, n2, n3))
    torch.cuda.synchronize()
    sess = ctx.ba_session(p, shard, r1, r2, r3, num_iterations=nit)

    class Synced:  # the context runs on its own stream: synchronise around every collective / host read
        BUILD, IMPORT_BUILD, SCHUR, SOLVE_UPDATE, RELABEL_COUNT, RELABEL_APPLY = 1, 2, 3, 4, 5, 6

        def phase(self, ph, v=0.0):
            torch.cuda.synchronize()
            sess.phase(ph, v)
            ctx.synchronize()

        def trial_done(self, a):
            sess.trial_done(a)

    res = pkg.sharding.ba_optimize_sharded(Synced(), r1, r2, r3, K, len(p["obs_pose"]), num_iterations=nit, group=group)
    out[idx] = (res,) + sess.end()
    ctx.close()


@pytest.mark.parametrize("seed,nk,nl,nit,world", [(31, 8, 600, 6, 1), (32, 10, 900, 6, 2), (33, 30, 1200, 4, 2)])
def test_sharded_session_vs_unsharded(pkg, seed, nk, nl, nit, world):
    p = pkg.synth.synth_ba_problem(seed, nk, nl, outlier_frac=0.04)
    shards = pkg.sharding.landmark_shards(p["obs_point"], nl, world)
    group = ThreadGroup(world)
    out = [None] * world
    th = [threading.Thread(target=_run_rank, args=(pkg, p, shards[r], group, nit, out, r)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(300)
    assert all(o is not None for o in out)
    res = out[0][0]
    poses = out[0][1]
    pts = sum(o[2] for o in out); chi2 = sum(o[3] for o in out); inl = sum(o[4].astype(np.int32) for o in out)
    o = B.optimize(p["poses"], p["points"], p["obs_pose"], p["obs_point"], p["obs_uv"], p["K"], num_iterations=nit)
    assert res["iterations"] == o["iterations"] and res["trials"] == o["trials"] and res["accepted"] == o["accepted"]
    assert np.isclose(res["chi2_final"], o["chi2_final"], rtol=1e-6)
    assert np.abs(poses - o["poses"]).max() / np.abs(o["poses"]).max() < REL_TOL
    assert (np.abs(pts - o["points"]).max(axis=1) / np.linalg.norm(o["points"], axis=1)).max() < REL_TOL
    assert np.allclose(chi2, o["chi2_per_obs"], rtol=1e-5, atol=1e-7)
    assert res["chi2_threshold"] == o["chi2_threshold"] and np.array_equal(inl.astype(bool), o["point_inlier"])
    for r in range(1, world):   # replicated poses are identical on every rank (same reduced system, same solve)
        assert np.array_equal(out[r][1], poses)
