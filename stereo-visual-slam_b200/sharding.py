"""Multi-GPU host logic: how the hot path is split over one-process-per-GPU ranks.

* Frontend (K1-K11): stereo frames are independent units -> contiguous blocks of frames per rank, no collective on
  the data path (`frame_shard`).
* BA (K13-K16): landmarks (with all their observations) are split into contiguous, observation-balanced ranges
  (`landmark_shards`); poses are replicated.  `ba_optimize_sharded` is the Levenberg-Marquardt driver: the same
  schedule as g2o / vslam_ba_optimize (tau = 1e-5, nu doubling, good-step clamp [1/3, 2/3], <= 10 trials), with one
  all-reduce of the reduced camera system [S | b] per LM trial.  It is written against a small session interface
  (build / schur / solve_update / relabel phases over three reduce buffers) so that the same driver runs on the GPU
  session (ffi.GpuBaSession, NCCL) and on a CPU stand-in in the gloo tests.
"""
from __future__ import annotations

import math

import numpy as np


def frame_shard(n_frames: int, world: int, rank: int):
    """Contiguous block [begin, end) of frames for `rank`; sizes differ by at most one."""
    base, rem = divmod(n_frames, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def landmark_shards(obs_point: np.ndarray, n_points: int, world: int):
    """Split landmarks [0, n_points) into `world` contiguous ranges with (nearly) equal numbers of observations.
    Returns a list of (begin, end)."""
    counts = np.bincount(np.asarray(obs_point, dtype=np.int64), minlength=n_points)
    cum = np.concatenate([[0], np.cumsum(counts)])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        c = int(np.searchsorted(cum, target, side="left"))
        cuts.append(min(max(c, cuts[-1]), n_points))
    cuts.append(n_points)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def _world(group):
    if hasattr(group, "all_reduce"):
        return group.size()
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return 1
    return dist.get_world_size(group)


def _all_reduce(t, group, op="sum"):
    """In-place all-reduce of a WHOLE tensor (never a slice view: not every backend reduces views in place).
    `group` is a torch.distributed process group (None = WORLD) or any object with .size() and .all_reduce(t, op)."""
    if hasattr(group, "all_reduce"):
        group.all_reduce(t, op)
        return
    import torch.distributed as dist
    if _world(group) == 1:
        return
    dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX, group=group)


def _reduce_r1(r1, K, group):
    """sum-reduce [Hpp | bp | chi2], max-reduce the trailing max|diag Hll| entry"""
    if _world(group) == 1:
        return
    md = r1[42 * K + 1:42 * K + 2].clone()
    _all_reduce(md, group, "max")
    _all_reduce(r1, group)
    r1[42 * K + 1:42 * K + 2] = md


def ba_optimize_sharded(session, r1, r2, r3, n_poses: int, n_obs_total: int, num_iterations: int = 10,
                        max_trials: int = 10, tau: float = 1e-5, chi2_th: float = 5.991, group=None):
    """LM control flow over a landmark-sharded session.  r1/r2/r3 are torch tensors (float64) on the session's
    device aliasing its reduce buffers.  Returns dict(iterations, trials, accepted, chi2_initial, chi2_final,
    lambda_final, chi2_threshold, n_inlier_obs, n_outlier_obs).  Collectives per LM trial: r2 (the reduced camera
    system) and two scalars of r3; per outer iteration: r1."""
    S = session
    K = n_poses
    world = _world(group)
    lam, ni = 0.0, 2.0
    trials = accepted = it = 0
    chi_first = chi_last = 0.0
    for it in range(num_iterations):
        S.phase(S.BUILD)
        _reduce_r1(r1, K, group)
        S.phase(S.IMPORT_BUILD)
        head = r1.cpu().numpy() if hasattr(r1, "cpu") else np.asarray(r1)
        current_chi = float(head[42 * K])
        if it == 0:
            chi_first = current_chi
            hpp = head[:36 * K].reshape(K, 6, 6)
            md = max(float(head[42 * K + 1]), float(np.abs(np.einsum("kii->ki", hpp)).max()))
            lam, ni = tau * md, 2.0
        rho, qmax = 0.0, 0
        while True:
            S.phase(S.SCHUR, lam)
            _all_reduce(r2, group)
            S.phase(S.SOLVE_UPDATE, lam)
            _all_reduce(r3, group)  # [chi2_trial, scale, ok] summed over ranks (the rest of r3 is scratch here)
            t3 = r3.cpu().numpy() if hasattr(r3, "cpu") else np.asarray(r3)
            ok2 = t3[2] >= world - 0.5  # every rank factorised the (identical) system
            temp_chi = float(t3[0]) if ok2 else float(np.finfo(np.float64).max)
            scale = (float(t3[1]) if ok2 else 0.0) + 1e-3
            rho = (current_chi - temp_chi) / scale
            good = rho > 0 and math.isfinite(temp_chi)
            if good:
                alpha = min(1.0 - (2 * rho - 1) ** 3, 2.0 / 3.0)
                lam *= max(1.0 / 3.0, alpha)
                ni = 2.0
                current_chi = temp_chi
                accepted += 1
            else:
                lam *= ni
                ni *= 2
            S.trial_done(good)
            qmax += 1
            trials += 1
            if not (rho < 0 and qmax < max_trials):
                break
        chi_last = current_chi
        if qmax == max_trials or rho == 0:
            it += 1
            break
    else:
        it = num_iterations
    if num_iterations <= 0:
        S.phase(S.BUILD)
        _reduce_r1(r1, K, group)
        head = r1.cpu().numpy() if hasattr(r1, "cpu") else np.asarray(r1)
        chi_first = chi_last = float(head[42 * K])
        it = 0
    S.phase(S.RELABEL_COUNT)
    _all_reduce(r3, group)
    cnt = (r3.cpu().numpy() if hasattr(r3, "cpu") else np.asarray(r3))[4:10]
    th, r = chi2_th, 0
    while r < 5:
        if cnt[r] / float(n_obs_total) > 0.5:
            break
        th *= 2
        r += 1
    S.phase(S.RELABEL_APPLY, th)
    return dict(iterations=it, trials=trials, accepted=accepted, chi2_initial=chi_first, chi2_final=chi_last,
                lambda_final=lam, chi2_threshold=th, n_inlier_obs=int(cnt[r]), n_outlier_obs=int(n_obs_total - cnt[r]))
