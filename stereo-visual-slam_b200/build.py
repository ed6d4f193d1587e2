"""In-tree native build of libvslam_b200.so (sm_100a only; nvcc cross-compiles without a GPU).

    python stereo-visual-slam_b200/build.py [--force] [--verbose]

Objects go to stereo-visual-slam_b200/build/, the shared library next to this file so that it travels
to the GPU box with the repo snapshot.  Static cudart (nvcc default): the library shares the primary
context -- and therefore streams and device pointers -- with torch in the same process.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libvslam_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# -fmad=false: the bit-exact float stages (Harris, fastAtan2, rBRIEF rotation, blur tail) must not be
# contracted; fused operations are written explicitly with fmaf()/fma() where the CPU path fuses.
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]
NO_FMAD = {"orb.cu", "pnp.cu"}  # (pnp.cu: OpenCV's EPnP in its exact fp64 arithmetic) everything else may contract


def _newer(src: str, dst: str, extra=()) -> bool:
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(p) > t for p in (src, *extra))


def build_native(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".inc"))]
    headers.append(os.path.join(HERE, "..", "include", "vslam_b200.h"))
    objs = []
    procs = []
    for f in sorted(os.listdir(CSRC)):
        if not f.endswith(".cu"):
            continue
        src = os.path.join(CSRC, f)
        obj = os.path.join(BUILD, f[:-3] + ".o")
        objs.append(obj)
        if force or _newer(src, obj, headers):
            cmd = [NVCC, *ARCH, *CFLAGS, *(["-fmad=false"] if f in NO_FMAD else []), "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas")
                cmd.insert(2, "-v")
                print(" ".join(cmd))
            procs.append((f, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for f, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"[build] {f} FAILED\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(f"[build] {f}\n{out}\n")
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("link failed")
    return LIB


HOST = os.path.join(HERE, "host")
HOST_LIB = os.path.join(HERE, "libvslam_b200_host.so")
RUN_VSLAM = os.path.join(HERE, "run_vslam")


def build_host(force: bool = False) -> str:
    """C++ drop-in layer (VO, Map, optimize_*) + the run_vslam main loop, linked against libvslam_b200.so."""
    cxx = os.environ.get("CXX", "g++")
    srcs = [os.path.join(HOST, f) for f in ("types_def.cpp", "map.cpp", "optimization.cpp", "visual_odometry.cpp",
                                            "png_reader.cpp")]
    deps = srcs + [os.path.join(HOST, "run_vslam.cpp"), os.path.join(HOST, "compat", "vslam_compat.hpp"),
                   os.path.join(HERE, "..", "include", "vslam_b200.h")]
    deps += [os.path.join(HOST, "stereo_visual_slam_main", f) for f in os.listdir(os.path.join(HOST, "stereo_visual_slam_main"))]
    if force or not os.path.exists(HOST_LIB) or any(os.path.getmtime(d) > os.path.getmtime(HOST_LIB) for d in deps):
        cmd = [cxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-I", HOST, "-o", HOST_LIB, *srcs,
               "-L", HERE, "-lvslam_b200", "-lz", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("host library build failed")
    if force or not os.path.exists(RUN_VSLAM) or os.path.getmtime(RUN_VSLAM) < os.path.getmtime(HOST_LIB) or \
            os.path.getmtime(RUN_VSLAM) < os.path.getmtime(os.path.join(HOST, "run_vslam.cpp")):
        cmd = [cxx, "-std=c++17", "-O2", "-Wall", "-I", HOST, "-o", RUN_VSLAM, os.path.join(HOST, "run_vslam.cpp"),
               "-L", HERE, "-lvslam_b200_host", "-lvslam_b200", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("run_vslam build failed")
    return HOST_LIB


REF_MAIN_SRC = "/root/reference/src/run_vslam.cpp"
RUN_VSLAM_REF = os.path.join(HERE, "run_vslam_ref")


def build_reference_main(force: bool = False):
    """The reference's own main loop, /root/reference/src/run_vslam.cpp, compiled UNMODIFIED from where it lies against
    the drop-in headers and linked with the drop-in host library ("run_vslam.cpp links unchanged apart from ROS
    plumbing").  Only possible where /root/reference exists (this container); the binary is git-ignored and travels to
    the GPU box with the snapshot.  Returns the path or None."""
    if not os.path.exists(REF_MAIN_SRC):
        return RUN_VSLAM_REF if os.path.exists(RUN_VSLAM_REF) else None
    cxx = os.environ.get("CXX", "g++")
    if force or not os.path.exists(RUN_VSLAM_REF) or os.path.getmtime(RUN_VSLAM_REF) < os.path.getmtime(HOST_LIB):
        cmd = [cxx, "-std=c++17", "-O2", "-I", HOST, "-o", RUN_VSLAM_REF, REF_MAIN_SRC,
               "-L", HERE, "-lvslam_b200_host", "-lvslam_b200", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("reference run_vslam.cpp did not compile against the drop-in layer")
    return RUN_VSLAM_REF


if __name__ == "__main__":
    if "--host" in sys.argv:
        build_native()
        print(build_host(force="--force" in sys.argv))
        print(build_reference_main(force="--force" in sys.argv))
        sys.exit(0)
    print(build_native(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
