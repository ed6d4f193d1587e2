"""Landmark-sharded BA over NCCL, one process per GPU:
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 tools/run_ba_sharded.py [cfg3|cfg5]
Prints one JSON line from rank 0: LM iterations/s (device time, max over ranks) and parity vs the C oracle."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import vslam_b200_loader
pkg = vslam_b200_loader.pkg

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
seed, nk, nl, nobs = (42, 10, 5000, None) if cfg == "cfg3" else (43, 50, 20000, 100000)
p = pkg.synth.synth_ba_problem(seed, nk, nl, n_obs_exact=nobs)
shards = pkg.sharding.landmark_shards(p["obs_point"], nl, world)
ctx = pkg.Context(device=local, max_images=0, max_width=0, max_height=0, max_keypoints=1, max_ba_poses=64,
                  max_ba_points=32768, max_ba_obs=262144)
stream = torch.cuda.Stream(device=dev)
ctx.set_stream(stream.cuda_stream)
n1, n2, n3 = pkg.ffi.ba_reduce_sizes(nk)
nit = 10
with torch.cuda.stream(stream):
    r1, r2, r3 = (torch.zeros(n, dtype=torch.float64, device=dev) for n in (n1, n2, n3))
    times = []
    for rep in range(4):
        sess = ctx.ba_session(p, shards[rank], r1, r2, r3, num_iterations=nit)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        res = pkg.sharding.ba_optimize_sharded(sess, r1, r2, r3, nk, len(p["obs_pose"]), num_iterations=nit,
                                               group=dist.group.WORLD if world > 1 else None)
        torch.cuda.synchronize(dev)
        times.append(time.perf_counter() - t0)
        poses, pts, chi2, inl = sess.end()
tt = torch.tensor([min(times[1:])], dtype=torch.float64, device=dev)
tp = torch.from_numpy(pts).to(dev)
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dist.all_reduce(tp)
if rank == 0:
    from oracle import ba_oracle
    o = ba_oracle.optimize(p["poses"], p["points"], p["obs_pose"], p["obs_point"], p["obs_uv"], p["K"], num_iterations=nit)
    print(json.dumps({"cfg": cfg, "n_gpus": world, "lm_iterations": res["iterations"], "lm_trials": res["trials"],
                      "iters_per_s": res["iterations"] / float(tt.item()), "ms_per_iter": float(tt.item()) * 1e3 / res["iterations"],
                      "pose_rel_err_vs_oracle": float(np.abs(poses - o["poses"]).max() / np.abs(o["poses"]).max()),
                      "point_rel_err_vs_oracle": float((np.abs(tp.cpu().numpy() - o["points"]).max(1) / np.linalg.norm(o["points"], axis=1)).max()),
                      "trials_match_oracle": res["trials"] == o["trials"], "allreduce_bytes_per_trial": 8 * (n2 + n3),
                      "chi2": [res["chi2_initial"], res["chi2_final"]]}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
