"""Build csrc/libgridworld_b200.so in-tree with nvcc for sm_100a (no GPU needed to compile)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libgridworld_b200.so")
SOURCES = ["gw_kernels.cu", "gw_actor.cu", "gw_replay.cu", "gw_train_ops.cu", "gw_maddpg.cu", "gw_maddpg_cluster.cu", "gw_wide.cu"]
HEADERS = ["gw_device.cuh", "gw_internal.h", "gw_replay_dev.cuh", "gw_maddpg.cuh", os.path.join("..", "..", "include", "gridworld_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]
OBJ_DIR = os.path.join(CSRC, "_obj")          # per-source objects (git-ignored): only stale sources are recompiled


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile every stale source to an object (in parallel), then link the shared library."""
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    hdr_t = max(os.path.getmtime(os.path.join(CSRC, f)) for f in HEADERS)
    procs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        objs.append(obj)
        if (not force and os.path.exists(obj)
                and os.path.getmtime(obj) > max(hdr_t, os.path.getmtime(os.path.join(CSRC, src)))):
            continue
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get("GW_NVCC_FLAGS", "").split() + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + out)
        if verbose:
            sys.stderr.write(out)
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs, cwd=CSRC,
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
