#!/usr/bin/env python
"""bench.py -- stereo frames/sec of the B200-native frontend (detect + match + triangulate), BASELINE.json metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs B] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of B synthetic 1241x376 stereo pairs at 2000 ORB keypoints
(BASELINE.json configs[1], batched): ORB on every left and right image, L<->R brute-force Hamming matching with
cross-check + the reference's distance gate, per-match DLT triangulation with the depth gates.
  value  : whole-job stereo frames/s with the images already resident in HBM (CUDA events on the launching stream,
           max over ranks); two input sets are alternated so a step's inputs are not L2-resident from the last step.
  e2e    : the same metric through the host-buffer C-ABI call (pinned host images in, all results out, H2D and D2H
           copies inside the timed region).
  roofline / cpu_baseline / clocks / gpu_launches as the driver contract asks.
Under torchrun (N > 1) every rank runs its own batch (frames are independent: weak scaling, no collective on the
data path); timing is bracketed by barrier + synchronize and reduced with MAX over ranks.
`--impl reference` times the reference's own CPU path for the same workload: live cv2 4.13.0 (cv::ORB, cv::BFMatcher,
cv::triangulatePoints -- the OpenCV calls the reference makes) wrapped by the oracle, on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, NFEAT = 1241, 376, 2000
PITCH = (W + 15) // 16 * 16   # device-resident images use a 16-byte aligned row pitch (the row_pitch argument of the C-ABI)
WORKLOAD = ("configs[1] batched: 1241x376 synthetic stereo pairs, 2000 ORB kp, "
            "detect(L,R)+BF-Hamming cross-check match+DLT triangulate")
METRIC = "stereo_frames_per_sec"
UNIT = "stereo frames/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML from a thread every 5 ms (an nvidia-smi
    process needs > 100 ms per query on an 8-GPU box and saw nothing of a 0.1 s region); nvidia-smi -lms as fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.nvml = None
        self.stop_flag = False
        self.sm, self.mx, self.reasons = [], [], set()
        self.err = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown",
                                               getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown",
                                               getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4))}
        reasons_fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        try:
            self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)))
        except Exception:
            pass
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                r = int(reasons_fn(self.h))
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception as ex:
                if not self.err:
                    self.err = repr(ex)
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            try:  # one sample from the calling thread (the GPU is still busy when the timed region has just ended), in
                # case the polling thread never got the interpreter during a very short region
                self.sm.append(float(self.nvml.nvmlDeviceGetClockInfo(self.h, self.nvml.NVML_CLOCK_SM)))
            except Exception:
                pass
            self.stop_flag = True
            self.t.join(timeout=1.0)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "samples": len(self.sm), "reasons": sorted(self.reasons), "source": "nvml, 5 ms",
                    **({"error": self.err} if self.err else {})}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 20"}


def make_batch(pkg, n_pairs: int, seed0: int, texture: str = "dense"):
    """n_pairs distinct synthetic stereo pairs are expensive to synthesise on the host; 8 base pairs are generated and
    cyclically rolled by whole rows so that every image of the batch is a different array with the same statistics."""
    base = [pkg.synth.synth_pair(seed0 + s, texture=texture)[:2] for s in range(8)]
    L = np.empty((n_pairs, H, W), np.uint8)
    R = np.empty((n_pairs, H, W), np.uint8)
    for i in range(n_pairs):
        l, r = base[i % 8]
        sh = (i // 8) * 7
        L[i] = np.roll(l, sh, axis=0)
        R[i] = np.roll(r, sh, axis=0)
    return L, R


def cv2_frontend(cv2, left, right, P1, P2, nfeat=NFEAT):
    """The reference CPU path for one stereo pair: the OpenCV calls the reference's VO path makes, live cv2."""
    orb = cv2.ORB_create(nfeat)
    kl, dl = orb.detectAndCompute(left, None)
    kr, dr = orb.detectAndCompute(right, None)
    m = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(dl, dr)
    if not m:
        return 0
    dist = np.array([x.distance for x in m])
    keep = dist <= max(2.0 * dist.min(), 30.0)
    xl = np.array([kl[x.queryIdx].pt for x, k in zip(m, keep) if k], np.float64).T
    xr = np.array([kr[x.trainIdx].pt for x, k in zip(m, keep) if k], np.float64).T
    X = cv2.triangulatePoints(P1, P2, xl, xr)
    z = X[2] / X[3]
    return int(((z > 10) & (z < 400)).sum())


def run_reference(args, rank, world):
    if rank != 0:
        return
    import cv2
    import vslam_b200_loader
    pkg = vslam_b200_loader.pkg
    ncores = os.cpu_count() or 1
    cv2.setNumThreads(ncores)
    s = pkg.synth
    P1, P2 = s.stereo_projection_matrices()
    sample = 4  # stereo pairs per step: a bounded sample of the batch workload
    nfeat = 4000 if args.config == 3 else NFEAT
    L, R = make_batch(pkg, sample, 0)
    for _ in range(max(args.warmup, 1)):
        cv2_frontend(cv2, L[0], R[0], P1, P2, nfeat)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for i in range(sample):
            cv2_frontend(cv2, L[i], R[i], P1, P2, nfeat)
    dt = time.perf_counter() - t0
    fps = args.steps * sample / dt
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD if args.config != 3 else
                       "configs[3]: batched 64 synthetic 1241x376 stereo pairs, 4000 ORB kp, frame-parallel over the GPUs, "
                       "detect(L,R)+BF-Hamming cross-check match+DLT triangulate",
                       "pairs_per_step": sample, "nfeatures": nfeat},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": ncores, "kind": "port",
                             "sample": f"{sample} stereo pairs/step x {args.steps} steps; live cv2 {cv2.__version__} "
                                       f"(cv::ORB({nfeat}) detectAndCompute x2, BFMatcher crossCheck + gate, "
                                       "triangulatePoints) = the OpenCV calls of the reference's VO path, "
                                       f"cv2.setNumThreads({ncores})"},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def algorithmic_bytes(n_images: int, n_pairs: int, n_kp: float, n_match: float):
    """ALGORITHMIC bytes per launch for each kernel, for a batch of n_images / n_pairs."""
    lv = [(1241, 376), (1034, 313), (862, 261), (718, 218), (598, 181), (499, 151), (416, 126), (346, 105)]
    px = [w * h for w, h in lv]
    tot = sum(px)
    return {
        # one launch per level l>=1: read level l-1, write level l  (sum over the 7 launches / 7 = per-launch average)
        "resize_level_kernel": n_images * (sum(px[:-1]) + sum(px[1:])) / 7.0,
        # every pixel that can reach a kept keypoint, read once: the frame ORB's 31-px border filter keeps plus the
        # 4-px ring of its NMS neighbours' segment tests (+ small candidate list); 1 065 395 of 1 444 097 pyramid pixels
        "fast_kernel": n_images * sum((w - 54) * (h - 54) for w, h in lv),
        "blur_kernel": n_images * 2 * tot,                  # read + write every pyramid pixel
        "harris_select_kernel": n_images * (2 * n_kp * (81 + 8) + n_kp * 8),   # 9x9 patch per candidate + lists
        "describe_kernel": n_images * n_kp * (749 + 512 + 28 + 32),            # IC disc + 512 samples + outputs
        "hamming_argmin_kernel": n_pairs * (2 * n_kp * 32 + 2 * n_kp * 4),     # both descriptor sets + fw/bw keys
        "crosscheck_gate_compact_kernel": n_pairs * (2 * n_kp * 4 + n_match * 16),
        "triangulate_kernel": n_pairs * n_match * (16 + 2 * 28 + 12 + 1),
    }


def _to_bytes(txt):
    """'93.319168 Mbyte' -> bytes"""
    v, u = txt.split()[:2]
    return float(v) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


def load_ncu_summary(kind="frontend"):
    """Newest profiles/rNN_vM_ncu_full_<kind>_summary.json (ncu --set full, tools/ncu_summary.py): per kernel the DRAM
    bytes per UNIT of work (image or pair) of that capture and the pipe percentages -- roofline.traffic and the pipe
    figures are read from here, never from literals."""
    import glob
    import re
    best = None
    for f in glob.glob(os.path.join(ROOT, "profiles", f"r*_v*_ncu_full_{kind}_summary.json")):
        m = re.search(r"r(\d+)_v(\d+)_", os.path.basename(f))
        key = (int(m.group(1)), int(m.group(2)))
        if best is None or key > best[0]:
            best = (key, f)
    if best is None:
        return None
    d = json.load(open(best[1]))
    m = re.search(r"(\d+) stereo pairs", d.get("command", ""))
    pairs = int(m.group(1)) if m else 32
    per_pair = ("hamming_argmin_kernel", "crosscheck_gate_compact_kernel", "triangulate_matches_kernel", "triangulate_kernel")
    out = {"source": os.path.relpath(best[1], ROOT), "pairs_per_launch": pairs, "kernels": {}}
    acc = {}
    for l in d["launches"]:
        k = l["kernel"]
        a = acc.setdefault(k, {"n": 0, "bytes": 0.0, "pipes": {}})
        a["n"] += 1
        a["bytes"] += _to_bytes(l["dram__bytes_read.sum"]) + _to_bytes(l["dram__bytes_write.sum"])
        for name, key in (("alu_pipe_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                          ("lsu_pipe_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
                          ("xu_pipe_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                          ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                          ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")):
            if key in l:
                a["pipes"][name] = a["pipes"].get(name, 0.0) + float(l[key].split()[0])
    for k, a in acc.items():
        units = pairs if k in per_pair else 2 * pairs
        name = "triangulate_kernel" if k == "triangulate_matches_kernel" else k
        # resize runs once per level: bytes of all its launches belong to one pass over the image
        tot = a["bytes"] if k == "resize_level_kernel" else a["bytes"] / a["n"]
        out["kernels"][name] = {"dram_bytes_per_unit": tot / units,
                                "pipes": {p: round(v / a["n"], 1) for p, v in a["pipes"].items()}}
    return out


def pin_to_gpu_numa(index: int):
    """Pin this process (and its first-touch pinned host buffers) to the CPUs of the NUMA node the GPU hangs off:
    with 8 ranks the host staging of the e2e leg otherwise crosses sockets.  Returns a short description."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = "0000:" + bus.split(":", 1)[1]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return "numa node unknown (single node)"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"numa node {node}, {len(cpus)} cpus"
        return f"numa node {node}: no allowed cpus"
    except Exception as ex:  # best effort
        return f"not pinned ({type(ex).__name__})"


def to_pitched(torch, img_batch, dev):
    """(B, H, W) uint8 numpy -> device tensor (B, H, PITCH), rows 16-byte aligned, padding zero"""
    t = torch.zeros((img_batch.shape[0], H, PITCH), dtype=torch.uint8, device=dev)
    t[:, :, :W] = torch.from_numpy(img_batch).to(dev)
    return t


def frontend_fps(pkg, torch, ctx, dev, stream, dist, world, B, nfeat, cap, steps, warmup, seed0, texture="dense",
                 kernel_times=False):
    """Device-resident stereo frames/s of one workload (CUDA events on the launching stream, max over ranks)."""
    P1, P2 = pkg.synth.stereo_projection_matrices()
    sets = []
    for k in range(2):
        L, R = make_batch(pkg, B, seed0 + 17 * k, texture)
        sets.append((to_pitched(torch, L, dev), to_pitched(torch, R, dev)))
    d_kp = torch.zeros((2 * B, cap, 7), dtype=torch.int32, device=dev)
    d_desc = torch.zeros((2 * B, cap, 32), dtype=torch.uint8, device=dev)
    d_nkp = torch.zeros(2 * B, dtype=torch.int32, device=dev)
    d_m = torch.zeros((B, cap, 4), dtype=torch.int32, device=dev)
    d_nm = torch.zeros(B, dtype=torch.int32, device=dev)
    d_xyz = torch.zeros((B, cap, 3), dtype=torch.float32, device=dev)
    d_fl = torch.zeros((B, cap), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step(i):
        dl, dr = sets[i & 1]
        ctx.stereo_frontend_dev(dl, dr, B, W, H, PITCH, PITCH * H, P1, P2, None, d_kp, d_desc, d_nkp, d_m, d_nm, d_xyz, d_fl,
                                nfeatures=nfeat)
    with torch.cuda.stream(stream):
        for i in range(warmup):
            step(i)
    ctx.orb_last_flags(2 * B)
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        flush.fill_(1)
        e0.record(stream)
        for i in range(steps):
            step(i)
        e1.record(stream)
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    rec = {"stereo_frames_per_s": world * B * steps / (ms * 1e-3), "ms_per_step": ms / steps, "pairs_per_gpu": B,
           "pairs_total": world * B, "nfeatures": nfeat, "steps": steps,
           "mean_keypoints_per_image": float(d_nkp.float().mean()), "mean_matches_per_pair": float(d_nm.float().mean())}
    if kernel_times:  # every launch alone on the stream, as for the roofline block of the headline workload
        ctx.set_concurrency(False)
        ctx.timing_enable(True)
        with torch.cuda.stream(stream):
            for i in range(2):
                step(i)
        torch.cuda.synchronize(dev)
        rec["kernel_ms_per_launch"] = {k: round(v[0] / max(v[1], 1), 4) for k, v in ctx.timing_read().items()}
        ctx.timing_enable(False)
        ctx.set_concurrency(True)
    return rec


def run_ours(args, rank, world, local_rank):
    import torch
    import vslam_b200_loader
    pkg = vslam_b200_loader.pkg

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = pin_to_gpu_numa(local_rank)
    if args.config == 3:
        # BASELINE.json configs[3]: 64 pairs at 4000 keypoints, frame-parallel over the GPUs of the box (strong scaling)
        B, NFEAT, cap = max(1, 64 // world), 4000, 4608
        workload = ("configs[3]: batched 64 synthetic 1241x376 stereo pairs, 4000 ORB kp, frame-parallel over the GPUs, "
                    "detect(L,R)+BF-Hamming cross-check match+DLT triangulate")
    else:
        B, NFEAT, cap = args.pairs, 2000, 2304
        workload = WORKLOAD
    ctx = pkg.Context(device=local_rank, max_images=2 * B, max_keypoints=cap, max_ba_poses=64, max_ba_points=32768,
                      max_ba_obs=262144)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    s = pkg.synth
    P1, P2 = s.stereo_projection_matrices()

    # two alternating input sets (2 x 2B x 466 KB > L2 for B >= 64), pinned on the host for the e2e leg
    sets = []
    for k in range(2):
        L, R = make_batch(pkg, B, 100 * rank + 17 * k)
        hl = torch.from_numpy(L).pin_memory()
        hr = torch.from_numpy(R).pin_memory()
        sets.append((hl, hr, to_pitched(torch, L, dev), to_pitched(torch, R, dev)))
    d_kp = torch.zeros((2 * B, cap, 7), dtype=torch.int32, device=dev)
    d_desc = torch.zeros((2 * B, cap, 32), dtype=torch.uint8, device=dev)
    d_nkp = torch.zeros(2 * B, dtype=torch.int32, device=dev)
    d_m = torch.zeros((B, cap, 4), dtype=torch.int32, device=dev)
    d_nm = torch.zeros(B, dtype=torch.int32, device=dev)
    d_xyz = torch.zeros((B, cap, 3), dtype=torch.float32, device=dev)
    d_fl = torch.zeros((B, cap), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_dev(i):
        _, _, dl, dr = sets[i & 1]
        ctx.stereo_frontend_dev(dl, dr, B, W, H, PITCH, PITCH * H, P1, P2, None, d_kp, d_desc, d_nkp, d_m, d_nm, d_xyz, d_fl,
                                nfeatures=NFEAT)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing (value) -------------------------------------------------------------
    with torch.cuda.stream(stream):
        for i in range(args.warmup):
            step_dev(i)
    ctx.orb_last_flags(2 * B)  # synchronises; raises on a work-list overflow
    n_kp_mean = float(d_nkp.float().mean())
    n_m_mean = float(d_nm.float().mean())
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = ctx.launch_count
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        flush.fill_(1)  # evict the previous step's lines; enqueued before the start event
        e0.record(stream)
        for i in range(args.steps):
            step_dev(i)
        e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    clk = clocks.stop()
    # per-kernel durations for the roofline block: the same steps once more with every launch alone on the stream
    # (the timed region above runs the library default: two half-batches interleaved on two streams, where the
    # CUDA-event time of one kernel includes its neighbour's)
    ctx.set_concurrency(False)
    ctx.timing_enable(True)
    with torch.cuda.stream(stream):
        for i in range(min(args.steps, 5)):
            step_dev(i)
    torch.cuda.synchronize(dev)
    ktimes = ctx.timing_read()
    ctx.timing_enable(False)
    ctx.set_concurrency(True)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max * 1e-3)

    # ---- end-to-end through the host-buffer C-ABI call (e2e) -------------------------------------------
    out = ctx.alloc_frontend_outputs(B)
    keep_alive = []
    for k, v in list(out.items()):  # pinned result buffers (structured dtypes go through a byte view)
        tt = torch.from_numpy(v.view(np.uint8).reshape(-1)).pin_memory()
        keep_alive.append(tt)
        out[k] = tt.numpy().view(v.dtype).reshape(v.shape)
    h2d = 2 * B * W * H
    d2h = sum(v.nbytes for v in out.values())
    for i in range(min(args.warmup, 3)):
        ctx.stereo_frontend(sets[i & 1][0], sets[i & 1][1], P1, P2, nfeatures=NFEAT, out=out)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        ctx.stereo_frontend(sets[i & 1][0], sets[i & 1][1], P1, P2, nfeatures=NFEAT, out=out)  # synchronous
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t.item())
    # the same through the streaming form of the call: two contexts alternate, so two batches are in flight and the first
    # upload / last download of one batch hide behind the other's kernels (every step still uploads its 2 x B images
    # from pinned host memory and downloads all its results inside the timed region)
    e2e_stream_value = None
    try:
        ctx_b = pkg.Context(device=local_rank, max_images=2 * B, max_keypoints=cap, max_ba_poses=0, max_ba_points=0, max_ba_obs=0)
        out_b = ctx_b.alloc_frontend_outputs(B)
        for k, v in list(out_b.items()):
            tt = torch.from_numpy(v.view(np.uint8).reshape(-1)).pin_memory()
            keep_alive.append(tt)
            out_b[k] = tt.numpy().view(v.dtype).reshape(v.shape)
        cs, outs = (ctx, ctx_b), (out, out_b)
        def stream_run(n):
            cs[0].stereo_frontend_begin(sets[0][0], sets[0][1], P1, P2, outs[0], nfeatures=NFEAT)
            for i in range(1, n):
                cs[i & 1].stereo_frontend_begin(sets[i & 1][0], sets[i & 1][1], P1, P2, outs[i & 1], nfeatures=NFEAT)
                cs[(i - 1) & 1].stereo_frontend_end()
            cs[(n - 1) & 1].stereo_frontend_end()
        stream_run(2)
        barrier()
        t0 = time.perf_counter()
        stream_run(args.steps)
        torch.cuda.synchronize(dev)
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_stream_value = world * B * args.steps / float(t.item())
        ctx_b.close()
    except Exception as ex:
        e2e_stream_value = None
        sys.stderr.write(f"[bench] streaming e2e leg failed: {ex!r}\n")
    # configs[1] read literally: ONE stereo pair per call through the host-buffer C-ABI (latency, not throughput)
    out1 = ctx.alloc_frontend_outputs(1)
    l1, r1 = sets[0][0][:1], sets[0][1][:1]
    for _ in range(3):
        ctx.stereo_frontend(l1, r1, P1, P2, nfeatures=NFEAT, out=out1)
    t0 = time.perf_counter()
    for _ in range(20):
        ctx.stereo_frontend(l1, r1, P1, P2, nfeatures=NFEAT, out=out1)
    single_pair_ms = (time.perf_counter() - t0) / 20 * 1e3
    matches_e2e = int(out["n_matches"].sum())
    usable = int(sum(int((out["flags"][i, :out["n_matches"][i]] & 1).sum()) for i in range(B)))

    # ---- secondary metric on every rank: one BA window per GPU (replicas, weak scaling) and, for N > 1, the
    # landmark-sharded window with its per-trial all-reduce (strong scaling) ---------------------------------
    ba_block = None
    if args.ba:
        try:
            ba_block = bench_ba(ctx, pkg, args, torch, dev, dist, rank, world, stream)
        except Exception as ex:  # the BA numbers are secondary; never lose the headline line
            ba_block = {"error": repr(ex)}

    # ---- BASELINE.json configs[3] beside the headline: 64 pairs at 4000 keypoints over all ranks ------------------
    cfg3_block = None
    if args.config != 3 and args.cfg3:
        try:
            ctx3 = pkg.Context(device=local_rank, max_images=2 * max(1, 64 // world), max_keypoints=4608, max_ba_poses=0,
                               max_ba_points=0, max_ba_obs=0)
            ctx3.set_stream(stream.cuda_stream)
            cfg3_block = frontend_fps(pkg, torch, ctx3, dev, stream, dist, world, max(1, 64 // world), 4000, 4608, 10, 3,
                                      1000 + 100 * rank)
            cfg3_block["scaling"] = "strong (64 pairs in total, frame-parallel, no data-path collective)"
            ctx3.close()
        except Exception as ex:
            cfg3_block = {"error": repr(ex)}

    # ---- the headline workload on a second texture with the FAST-corner density of street scenes (3 % of the pixels
    # instead of 17 %): the segment-test score pass, the top kernel's main cost, only runs where the compass pre-test
    # passes, so the bench texture is the kernel's worst case ------------------------------------------------------
    street_block = None
    if args.street:
        try:
            street_block = frontend_fps(pkg, torch, ctx, dev, stream, dist, world, B, NFEAT, cap, args.steps, args.warmup,
                                        5000 + 100 * rank, texture="street", kernel_times=True)
            street_block["texture"] = ("synth_canvas_street: smooth shading + soft-edged flat objects + patches of fine "
                                       "texture + sensor noise; 3.2 % of the pixels pass FAST-9/16 at t=20 (bench texture: 17.4 %)")
        except Exception as ex:
            street_block = {"error": repr(ex)}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -----------------------------------------------------------------
    peak, peak_src = _peaks()
    alg = algorithmic_bytes(2 * B, B, n_kp_mean, n_m_mean)
    share = {k: v[0] for k, v in ktimes.items()}
    tot_k = sum(share.values()) or 1.0
    top = max(share, key=share.get)
    avg_ms = ktimes[top][0] / ktimes[top][1]
    achieved = alg.get(top, 0.0) / (avg_ms * 1e-3) / 1e9
    ncu = load_ncu_summary("frontend")
    per_pair = ("hamming_argmin_kernel", "triangulate_kernel", "crosscheck_gate_compact_kernel")
    nk = (ncu or {}).get("kernels", {}).get(top)
    roof = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": nk["dram_bytes_per_unit"] * (B if top in per_pair else 2 * B) if nk else None,
            "traffic_source": (ncu["source"] + " (dram__bytes_read+write per unit of the newest ncu --set full capture x "
                               "units per launch here)") if ncu else None,
            "peak_source": peak_src, "avg_launch_ms": avg_ms,
            "ncu_pipes": {"source": ncu["source"] if ncu else None, top: nk["pipes"] if nk else None,
                          "note": "pipe utilisation (% of peak, sustained active) of the dominant kernel in that capture"},
            "algorithmic_bytes_per_launch": alg.get(top, 0.0),
            "kernel_time_share": {k: round(v / tot_k, 4) for k, v in sorted(share.items(), key=lambda kv: -kv[1])},
            "kernel_ms_per_step": {k: round(v[0] / min(args.steps, 5), 4) for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1][0])},
            "note": "per-kernel CUDA-event times measured live in this run, right after the timed region, with every launch "
                    "alone on the stream (the timed region interleaves two half-batches on two streams); "
                    "see DESIGN.md §4 for the bytes"}
    # Hamming kernel: the north star asks for its HBM fraction; it is POPC-pipe bound by construction
    if "hamming_argmin_kernel" in ktimes:
        hm = ktimes["hamming_argmin_kernel"][0] / ktimes["hamming_argmin_kernel"][1]
        hk = (ncu or {}).get("kernels", {}).get("hamming_argmin_kernel")
        roof["hamming"] = {"hbm_GBps": alg["hamming_argmin_kernel"] / (hm * 1e-3) / 1e9,
                           "hbm_frac": alg["hamming_argmin_kernel"] / (hm * 1e-3) / 1e9 / peak,
                           # 256-bit distance = 8 XOR words, two carry-save adders -> 6 POPC (match.cu, MT_CSA = 2)
                           "popc_issued_Tops": B * n_kp_mean * n_kp_mean * 6 / (hm * 1e-3) / 1e12,
                           "popc_pipe_peak_Tops": 148 * 16 * (clk["sm_mhz"] or 1965.0) * 1e6 / 1e12,
                           "ncu_pipes": hk["pipes"] if hk else None,
                           "bound": "integer ALU (LOP3 carry-save + min/compare) with the POPC (XU) pipe second; HBM "
                                    "fraction is tiny by construction: brute force has N/4 ops per byte (SURVEY 8d)",
                           "avg_launch_ms": hm}

    # ---- CPU baseline beside it (bounded sample, all host threads) ----------------------------------------
    import cv2
    ncores = os.cpu_count() or 1
    cv2.setNumThreads(ncores)
    Ls, Rs = sets[0][0].numpy(), sets[0][1].numpy()
    cv2_frontend(cv2, Ls[0], Rs[0], P1, P2, NFEAT)
    n_cpu, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < args.cpu_seconds and n_cpu < B:
        cv2_frontend(cv2, Ls[n_cpu], Rs[n_cpu], P1, P2, NFEAT)
        n_cpu += 1
    cpu_fps = n_cpu / (time.perf_counter() - t0)
    # the same path on one thread and stage by stage: median of 5 on pair 0
    def _med(fn, reps=5):
        ts = []
        for _ in range(reps):
            t1 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t1)
        return float(np.median(ts) * 1e3)
    orb = cv2.ORB_create(NFEAT)
    kl0, dl0 = orb.detectAndCompute(Ls[0], None)
    kr0, dr0 = orb.detectAndCompute(Rs[0], None)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True)
    stages = {"threads": ncores,
              "orb_detect_ms": _med(lambda: orb.detect(Ls[0], None)),
              "orb_compute_ms": _med(lambda: orb.compute(Ls[0], kl0)),
              "bf_match_ms": _med(lambda: bf.match(dl0, dr0))}
    cv2.setNumThreads(1)
    stages_1t = {"threads": 1, "frontend_pair_ms": _med(lambda: cv2_frontend(cv2, Ls[0], Rs[0], P1, P2, NFEAT), 3),
                 "bf_match_ms": _med(lambda: bf.match(dl0, dr0), 3)}
    cv2.setNumThreads(ncores)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.config == 3 else "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload,
                       "pairs_per_step_per_gpu": B, "nfeatures": NFEAT, "anms": "off (configs[1])", "host_affinity": numa,
                       "device_row_pitch": PITCH,
                       "l2": "two alternating input sets + 256 MiB flush before the timed region; per-step working "
                             "set (inputs+pyramids+blurred) ~%.0f MB > 126 MB L2" % (2 * B * 3.7),
                       "mean_keypoints_per_image": n_kp_mean, "mean_matches_per_pair": n_m_mean,
                       "parallelism": f"frames sharded over {world} GPU(s), no data-path collective"},
            "e2e": {"value": max(e2e_value, e2e_stream_value or 0.0), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h,
                    "mode": ("vslam_stereo_frontend_batch_begin/_end, two contexts alternating = two batches in flight"
                             if (e2e_stream_value or 0.0) > e2e_value else "vslam_stereo_frontend_batch, one call at a time"),
                    "one_call_at_a_time": e2e_value, "two_batches_in_flight": e2e_stream_value,
                    "matches_per_step": matches_e2e, "usable_points_per_step": usable,
                    "single_pair_latency_ms": single_pair_ms},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof,
            "cpu_baseline": {"value": cpu_fps, "unit": UNIT, "cores": ncores, "kind": "port",
                             "sample": f"{n_cpu} stereo pairs of the same batch through the oracle's cv2 "
                                       f"{cv2.__version__} path (ORB({NFEAT}) x2, BFMatcher crossCheck + gate, "
                                       f"triangulatePoints), cv2.setNumThreads({ncores})",
                             "stages_ms": stages, "single_thread": stages_1t}}
    if ba_block is not None:
        line["ba"] = ba_block
    if cfg3_block is not None:
        line["configs3_64pairs_4000kp"] = cfg3_block
    if street_block is not None:
        line["street_texture"] = street_block
    try:
        line["pnp"] = bench_pnp(ctx, pkg)
    except Exception as ex:
        line["pnp"] = {"error": repr(ex)}
    try:
        line["vo_loop"] = bench_vo_loop(pkg)
    except Exception as ex:
        line["vo_loop"] = {"error": repr(ex)}
    if args.sgbm:
        try:
            line["sgbm"] = bench_sgbm(ctx, pkg, torch, dev, stream, sets[0][2][:8], sets[0][3][:8], peak)
        except Exception as ex:
            line["sgbm"] = {"error": repr(ex)}
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def bench_vo_loop(pkg):
    """The reference's own published number: wall-clock per keyframe / per non-keyframe of the sequential VO loop
    (README.md:90: ~0.18 s and ~0.04 s on an unstated CPU, KITTI 00).  Here: the C++ drop-in layer's run_vslam on a
    24-frame synthetic 1241x376 stereo sequence at the reference's operating point -- ORB(3000) + ANMS(500), dense
    StereoSGBM depth (the default), and per keyframe with a full window optimize_map x3 (5+5+10 LM iterations) +
    optimize_pose_only (10).  A second run with fewer features makes every frame a keyframe so that the BA part is timed."""
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.abspath(__file__))
    exe = os.path.join(root, "stereo-visual-slam_b200", "run_vslam")
    if not os.path.exists(exe):
        return {"error": "run_vslam not built"}
    n = 24
    lefts, rights, t, _ = pkg.synth.synth_sequence(3, n)
    out = {"published_reference": {"keyframe_s": 0.18, "non_keyframe_s": 0.04, "hardware": "not stated (README.md:90)"}}
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(d + "/image_0")
        os.makedirs(d + "/image_1")
        for i in range(n):
            pkg.synth.write_pgm(f"{d}/image_0/{i:06d}.pgm", lefts[i])
            pkg.synth.write_pgm(f"{d}/image_1/{i:06d}.pgm", rights[i])
        for name, extra in (("reference_defaults_dense", []),   # dense StereoSGBM depth is the drop-in layer's default
                            ("keyframe_every_frame_dense", ["--nfeatures", "1000", "--anms", "110"]),
                            # the north star's sparse-stereo depth (ORB on both images + L<->R matching + DLT) instead of SGBM
                            ("keyframe_every_frame_sparse", ["--sparse", "--nfeatures", "1000", "--anms", "110"]),
                            # BASELINE.json configs[2]: the stream at 2000 keypoints per frame (no ANMS thinning); its
                            # K=10 / L=5000 BA window is timed in the "ba" block (cfg3_K10_L5000.reference_call_sequence_ms)
                            ("configs2_stream_2000kp", ["--nfeatures", "2000", "--anms", "0"])):
            with tempfile.TemporaryDirectory() as w:
                r = subprocess.run([exe, d + "/", str(n), *extra], cwd=w, capture_output=True, text=True, timeout=300)
            rows = [l.split() for l in r.stdout.splitlines() if l.startswith("frame ")]
            if r.returncode != 0 or len(rows) < 3:
                out[name] = {"error": (r.stderr or r.stdout)[-200:]}
                continue
            ms = np.array([float(x[18]) for x in rows])
            kf = np.array([int(x[15]) for x in rows]).astype(bool)
            nkf_window = np.array([int(x[16]) for x in rows])
            pos = np.array([[float(x[5]), float(x[9]), float(x[13])] for x in rows])
            err = np.abs(pos - t[:len(rows)]).max()
            ate = float(np.sqrt(((pos - t[:len(rows)]) ** 2).sum(axis=1).mean()))  # no alignment: frame 0 is the origin
            steady = np.arange(len(rows)) >= 2  # frame 0 = initialisation, frame 1 pays one-off allocations
            full = nkf_window >= 10
            rec = {"frames": len(rows), "keyframes": int(kf.sum()), "max_abs_position_error_m": float(err), "ate_rmse_m": ate,
                   "path_length_m": float(np.linalg.norm(np.diff(t[:len(rows)], axis=0), axis=1).sum()),
                   "non_keyframe_ms": float(np.median(ms[steady & ~kf])) if (steady & ~kf).any() else None,
                   "keyframe_ms": float(np.median(ms[steady & kf])) if (steady & kf).any() else None,
                   "keyframe_with_full_window_ba_ms": float(np.median(ms[steady & kf & full])) if (steady & kf & full).any() else None}
            out[name] = rec
    return out


def bench_pnp(ctx, pkg):
    """VO::motion_estimation's solvePnPRansac call (visual_odometry.cpp:277: 100 iterations, 4 px, 0.99) on 400
    landmark-keypoint pairs with 20 % gross outliers: host-buffer C-ABI call vs live cv2 on the host."""
    import cv2
    rng = np.random.default_rng(0)
    n = 400
    K = pkg.synth.kitti_K()
    R, t = pkg.synth.se3_exp(np.array([0.4, -0.05, 0.9, 0.01, 0.03, -0.005]))
    pw = np.stack([rng.uniform(-15, 15, n), rng.uniform(-3, 3, n), rng.uniform(6, 45, n)], 1).astype(np.float32)
    pc = pw.astype(np.float64) @ R.T + t
    uv = (pc[:, :2] / pc[:, 2:3]) * [K[0, 0], K[1, 1]] + [K[0, 2], K[1, 2]] + rng.normal(0, 0.3, (n, 2))
    bad = rng.permutation(n)[:n // 5]
    uv[bad] += rng.uniform(30, 200, (len(bad), 2)) * rng.choice([-1, 1], (len(bad), 2))
    uv = uv.astype(np.float32)
    g = ctx.pnp_ransac(pw, uv, K, 100, 4.0, 0.99)
    t0 = time.perf_counter()
    for _ in range(20):
        g = ctx.pnp_ransac(pw, uv, K, 100, 4.0, 0.99)
    gpu_ms = (time.perf_counter() - t0) / 20 * 1e3
    ok, rvec, tvec, inl = cv2.solvePnPRansac(pw, uv, K, None, iterationsCount=100, reprojectionError=4.0, confidence=0.99)
    t0 = time.perf_counter()
    for _ in range(20):
        cv2.solvePnPRansac(pw, uv, K, None, iterationsCount=100, reprojectionError=4.0, confidence=0.99)
    cpu_ms = (time.perf_counter() - t0) / 20 * 1e3
    return {"points": n, "e2e_ms_per_call": gpu_ms, "cv2_ms_per_call": cpu_ms,
            "inlier_sets_differ_by": int(len(np.setxor1d(g["inliers"], inl.ravel()))),
            "t_rel_diff_vs_cv2": float(np.abs(g["tvec"] - tvec.ravel()).max() / np.linalg.norm(tvec))}


def bench_sgbm(ctx, pkg, torch, dev, stream, dl, dr, hbm_peak):
    """Dense stereo (VO::disparity_map, the reference's StereoSGBM call, visual_odometry.cpp:163-168): device-resident
    throughput on a batch of 8 pairs, single-pair latency through the host-buffer C-ABI call, per-kernel algorithmic
    HBM rates (one u16 volume = H x (W-96) x 96 x 2 B per pair) and live cv2.StereoSGBM beside it."""
    import cv2
    B = int(dl.shape[0])
    ctx.set_stream(stream.cuda_stream)  # the events below bracket this stream
    d16 = torch.empty((B, H, W), dtype=torch.int16, device=dev)
    with torch.cuda.stream(stream):
        for _ in range(2):
            ctx.sgbm_compute_dev(dl, dr, B, W, H, PITCH, PITCH * H, d16)
    torch.cuda.synchronize(dev)
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):  # throughput: chunks interleaved on two streams (the library default)
        e0.record(stream)
        for _ in range(reps):
            ctx.sgbm_compute_dev(dl, dr, B, W, H, PITCH, PITCH * H, d16)
        e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    ctx.set_concurrency(False)       # per-kernel times: every launch alone on the stream
    ctx.timing_enable(True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(reps):
            ctx.sgbm_compute_dev(dl, dr, B, W, H, PITCH, PITCH * H, d16)
        e1.record(stream)
    torch.cuda.synchronize(dev)
    ms_serial = e0.elapsed_time(e1) / reps
    kt = {k: v for k, v in ctx.timing_read().items() if k.startswith("sgbm")}
    ctx.timing_enable(False)
    ctx.set_concurrency(True)
    vol = H * (W - 96) * 96 * 2
    # algorithmic volume transfers per pair: cost writes C; vertical reads C once and writes 3 path volumes;
    # row-forward reads C + 3 paths and writes S4; row-backward reads C + S4
    alg = {"sgbm_cost_kernel": 1 * vol, "sgbm_vertical_kernel": 4 * vol, "sgbm_row_forward_kernel": 5 * vol, "sgbm_row_backward_kernel": 2 * vol}
    kern = {}
    for k, (tot, n) in kt.items():
        per = tot / max(n, 1)
        kern[k] = {"ms_per_launch": per}
        if k in alg:
            gbs = alg[k] * B / (per * 1e-3) / 1e9
            kern[k].update({"algorithmic_GBps": gbs, "hbm_frac": gbs / hbm_peak})
    hl, hr = np.ascontiguousarray(dl[0, :, :W].cpu().numpy()), np.ascontiguousarray(dr[0, :, :W].cpu().numpy())
    ctx.sgbm_compute(hl, hr)
    t0 = time.perf_counter()
    for _ in range(5):
        one = ctx.sgbm_compute(hl, hr)
    lat_ms = (time.perf_counter() - t0) / 5 * 1e3
    sg = cv2.StereoSGBM_create(0, 96, 9, 8 * 81, 32 * 81, 1, 63, 10, 100, 32)
    ref = sg.compute(hl, hr)
    t0 = time.perf_counter()
    for _ in range(3):
        sg.compute(hl, hr)
    cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
    return {"workload": f"cv::StereoSGBM(0,96,9,648,2592,1,63,10,100,32) on {B} synthetic 1241x376 pairs per launch",
            "pairs_per_s_device": B / (ms * 1e-3), "ms_per_pair_device": ms / B,
            "pairs_per_s_device_single_stream": B / (ms_serial * 1e-3),
            "single_pair_e2e_ms": lat_ms, "cv2_ms_per_pair": cpu_ms, "cv2_threads": cv2.getNumThreads(),
            "bit_exact_vs_cv2": bool(np.array_equal(one, ref) and np.array_equal(d16[0].cpu().numpy(), ref)),
            "kernels": kern}


def bench_ba(ctx, pkg, args, torch, dev, dist, rank, world, stream):
    """Secondary metric of BASELINE.json: BA LM-iterations/s on configs[2] (K=10, L=5000) and configs[4]
    (K=50, L=20000, 100k observations), through the host-buffer C-ABI call (e2e) and kernel-only (CUDA events),
    next to the single-thread C oracle (g2o's default build has no OpenMP).

    N GPUs: every rank optimises its own window at the same time (independent windows = replicas, no collective;
    whole-job rate = N x iterations / max-over-ranks time), and the landmark-sharded session solves ONE window on
    all ranks with one all-reduce of the reduced camera system per LM trial (strong scaling; at these window sizes a
    trial is shorter than its collectives)."""
    out = {}
    cpu_group = dist.new_group(backend="gloo") if dist is not None else None
    for name, (seed, nk, nl, nobs, nit) in {"cfg3_K10_L5000": (42, 10, 5000, None, 10),
                                            "cfg5_K50_L20000_obs100k": (43, 50, 20000, 100000, 10)}.items():
        p = pkg.synth.synth_ba_problem(seed, nk, nl, n_obs_exact=nobs)
        a = (p["poses"], p["points"], p["obs_pose"], p["obs_point"], p["obs_uv"], p["K"])
        for _ in range(2):
            r = ctx.ba_optimize(*a, num_iterations=nit)
        if dist is not None:
            dist.barrier()
        ctx.timing_enable(True)
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            r = ctx.ba_optimize(*a, num_iterations=nit)
        dt = (time.perf_counter() - t0) / reps
        kt = ctx.timing_read().get("ba_lm_kernel", (0.0, 1))
        ctx.timing_enable(False)
        phase_us = ctx.ba_last_phase_us()   # device-side profile of the last of those calls
        kms = kt[0] / max(kt[1], 1)
        tt = torch.tensor([dt * 1e3, kms], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms, kms = float(tt[0]), float(tt[1])
        rec = {"observations": int(len(p["obs_pose"])), "lm_iterations": r["iterations"], "lm_trials": r["trials"],
               "windows_in_parallel": world, "scaling": "weak (one window per GPU, no collective)",
               "e2e_iters_per_s": world * r["iterations"] / (e2e_ms * 1e-3),
               "kernel_iters_per_s": world * r["iterations"] / (kms * 1e-3),
               "kernel_ms": kms, "e2e_ms": e2e_ms, "chi2": [r["chi2_initial"], r["chi2_final"]]}
        if name.startswith("cfg3"):
            # the reference's per-keyframe call sequence on this window (run_vslam.cpp:61-70): optimize_map(5), (5), (10)
            # and optimize_pose_only(10), four host-buffer C-ABI calls
            t0 = time.perf_counter()
            for _ in range(reps):
                for nit2, po in ((5, False), (5, False), (10, False), (10, True)):
                    ctx.ba_optimize(*a, num_iterations=nit2, pose_only=po)
            rec["reference_call_sequence_ms"] = (time.perf_counter() - t0) / reps * 1e3
        if world > 1:
            # strong scaling: one window, landmarks sharded over the ranks
            shards = pkg.sharding.landmark_shards(p["obs_point"], nl, world)
            n1, n2, n3 = pkg.ffi.ba_reduce_sizes(nk)
            with torch.cuda.stream(stream):
                r1, r2, r3 = (torch.zeros(n, dtype=torch.float64, device=dev) for n in (n1, n2, n3))
                times = []
                for _ in range(3):
                    sess = ctx.ba_session(p, shards[rank], r1, r2, r3, num_iterations=nit)
                    torch.cuda.synchronize(dev)
                    dist.barrier()
                    t0 = time.perf_counter()
                    res = pkg.sharding.ba_optimize_sharded(sess, r1, r2, r3, nk, len(p["obs_pose"]), num_iterations=nit,
                                                           group=dist.group.WORLD)
                    torch.cuda.synchronize(dev)
                    times.append(time.perf_counter() - t0)
                    sp, _, _, _ = sess.end()
            ts = torch.tensor([min(times[1:])], dtype=torch.float64, device=dev)
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
            rec["sharded"] = {"scaling": "strong (one window, landmarks sharded, 1 all-reduce of [S|b] per LM trial)",
                              "iters_per_s": res["iterations"] / float(ts[0]), "ms_per_iter": float(ts[0]) * 1e3 / res["iterations"],
                              "allreduce_bytes_per_trial": 8 * (n2 + n3),
                              "pose_rel_diff_vs_single_gpu": float(np.abs(sp - r["poses"]).max() / np.abs(r["poses"]).max())}
        if world > 1:
            # strong scaling, device-side: ONE process (rank 0) drives all GPUs of the box through
            # vslam_ba_optimize_multi -- persistent LM kernel per GPU, partial systems exchanged by peer stores over
            # NVLink, no host or NCCL in the loop.  The other ranks idle at the barrier meanwhile.
            # (the other ranks wait on a CPU-side gloo barrier: an NCCL barrier would park a spinning kernel on their GPU,
            # which this rank is about to use)
            torch.cuda.synchronize(dev)
            dist.barrier(group=cpu_group)
            if rank == 0:
                try:
                    ctxs = [ctx] + [pkg.Context(device=dd, max_images=0, max_width=0, max_height=0, max_keypoints=1,
                                                max_ba_poses=64, max_ba_points=32768, max_ba_obs=262144)
                                    for dd in range(world) if dd != dev.index]
                    ctx.set_stream(None)
                    rm = pkg.Context.ba_optimize_multi(ctxs, *a, num_iterations=nit)
                    for _ in range(2):
                        rm = pkg.Context.ba_optimize_multi(ctxs, *a, num_iterations=nit)
                    for cc in ctxs:
                        cc.timing_enable(True)
                    t0 = time.perf_counter()
                    for _ in range(5):
                        rm = pkg.Context.ba_optimize_multi(ctxs, *a, num_iterations=nit)
                    dtm = (time.perf_counter() - t0) / 5
                    kmax = 0.0
                    for cc in ctxs:
                        ktm = cc.timing_read().get("ba_lm_kernel", (0.0, 1))
                        kmax = max(kmax, ktm[0] / max(ktm[1], 1))
                        cc.timing_enable(False)
                    rec["multi_device"] = {
                        "scaling": "strong (one window, landmarks sharded over the GPUs of one process, device-side LM, "
                                   "P2P exchange of [S|bs|bp|chi2] per trial + scalar exchanges)",
                        "gpus": world, "kernel_ms_max_over_gpus": kmax, "e2e_ms": dtm * 1e3,
                        "kernel_iters_per_s": rm["iterations"] / (kmax * 1e-3), "ms_per_iter_kernel": kmax / rm["iterations"],
                        "e2e_iters_per_s": rm["iterations"] / dtm, "exchanges": rm["exchanges"], "trials": rm["trials"],
                        "exchange_bytes_per_trial_per_gpu": 8 * ((6 * nk) ** 2 // 2 + 3 * 6 * nk) * world,
                        "device_profile_us": ctx.ba_multi_last_profile_us(),
                        "speedup_vs_one_gpu_kernel": kms / kmax,
                        "pose_rel_diff_vs_single_gpu": float(np.abs(rm["poses"] - r["poses"]).max() / np.abs(r["poses"]).max())}
                    for cc in ctxs[1:]:
                        cc.close()
                    ctx.set_stream(stream.cuda_stream)
                except Exception as ex:
                    rec["multi_device"] = {"error": repr(ex)}
            dist.barrier(group=cpu_group)
        if rank == 0:
            try:   # device-side phase profile of the last call and, for the large window, the dense DMMA Schur probe
                tr = max(r["trials"], 1)
                rec["phase_us_per_trial"] = {k: round(v / tr, 2) for k, v in phase_us.items()}
                if name.startswith("cfg5"):
                    n = 6 * nk
                    n1, n2, n3 = pkg.ffi.ba_reduce_sizes(nk)
                    with torch.cuda.stream(stream):
                        q1, q2, q3 = (torch.zeros(k, dtype=torch.float64, device=dev) for k in (n1, n2, n3))
                        Sd = torch.zeros(n * n, dtype=torch.float64, device=dev)
                    torch.cuda.synchronize(dev)
                    with torch.cuda.stream(stream):  # GpuBaSession moves the context onto torch's current stream
                        sess = ctx.ba_session(p, (0, nl), q1, q2, q3, num_iterations=1)
                    sess.phase(sess.BUILD)
                    sess.phase(sess.SCHUR, 1e-3)
                    best = None
                    for _ in range(3):
                        m = sess.schur_dense(Sd)
                        best = m if best is None or m[1] < best[1] else best
                    torch.cuda.synchronize(dev)
                    Ss = q2[:n * n].reshape(n, n)
                    iu = torch.triu(torch.ones(n, n, dtype=torch.bool, device=dev))
                    diff = float((Ss - Sd.reshape(n, n))[iu].abs().max() / Ss[iu].abs().max())
                    sess.end()
                    flops = 2.0 * n * n * 3 * nl
                    rec["dense_dmma_schur"] = {
                        "what": "the same Schur product as ONE dense fp64 SYRK on the tensor cores (mma.sync.m8n8k4.f64)",
                        "fill_ms": best[0], "syrk_ms": best[1], "dense_flops": flops,
                        "upper_tile_TFLOPs": 0.5 * flops * 1.2 / (best[1] * 1e-3) / 1e12,
                        "sparse_schur_us": rec["phase_us_per_trial"].get("schur"),
                        "sparse_flops": "77e6 (SURVEY 8d)", "max_rel_diff_vs_sparse": diff}
            except Exception as ex:
                rec["dense_dmma_schur"] = {"error": repr(ex)}
            from oracle import ba_oracle
            t0 = time.perf_counter()
            o = ba_oracle.optimize(*a, num_iterations=nit)
            dt_cpu = time.perf_counter() - t0
            rec.update({"cpu_oracle_iters_per_s": o["iterations"] / dt_cpu, "cpu_cores": 1,
                        "pose_rel_err_vs_oracle": float(np.abs(r["poses"] - o["poses"]).max() / np.abs(o["poses"]).max())})
        out[name] = rec
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=256, help="stereo pairs per step per GPU (configs[1] batched)")
    ap.add_argument("--config", type=int, default=1, choices=[1, 3],
                    help="headline workload: BASELINE.json configs[1] (batched, 2000 kp) or configs[3] (64 pairs, 4000 kp)")
    ap.add_argument("--no-cfg3", dest="cfg3", action="store_false")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-ba", dest="ba", action="store_false")
    ap.add_argument("--no-sgbm", dest="sgbm", action="store_false")
    ap.add_argument("--no-street", dest="street", action="store_false")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
