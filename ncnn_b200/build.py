"""Build libncnn_b200.so in-tree: nvcc (sm_100a) for csrc/cuda/*.cu, g++ for csrc/host/*.cpp.

    python -m ncnn_b200.build [-j N] [--force]

The shared library lands at ncnn_b200/libncnn_b200.so (git-ignored, shipped to the GPU box by gpurun).
Objects are cached under ncnn_b200/_build/ keyed on source + header mtimes.
"""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libncnn_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("NCNN_B200_CXX", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++")

INCLUDES = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(HERE, "csrc", "host"), "-I" + os.path.join(HERE, "csrc", "cuda")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--ftz=true", "-Xcompiler", "-fPIC,-fvisibility=hidden",
              "-ccbin", CXX, "--expt-relaxed-constexpr", "-Xcudafe", "--diag_suppress=177"]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-Wall", "-Wno-unused-function", "-pthread", "-I/usr/local/cuda/include"]


def _newest(paths):
    return max((os.path.getmtime(p) for p in paths), default=0.0)


def _compile(src, obj, cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return src, r.returncode, r.stdout


def build(jobs=None, force=False, verbose=True):
    os.makedirs(BUILD, exist_ok=True)
    cu = sorted(glob.glob(os.path.join(HERE, "csrc", "cuda", "*.cu")))
    cpp = sorted(glob.glob(os.path.join(HERE, "csrc", "host", "*.cpp")) + glob.glob(os.path.join(HERE, "csrc", "host", "layer", "*.cpp")))
    headers = glob.glob(os.path.join(HERE, "csrc", "**", "*.h"), recursive=True) + glob.glob(os.path.join(HERE, "csrc", "**", "*.cuh"), recursive=True) \
        + glob.glob(os.path.join(ROOT, "include", "*.h"))
    hdr_time = _newest(headers + [os.path.abspath(__file__)])
    tasks, objs = [], []
    for s in cu + cpp:
        o = os.path.join(BUILD, os.path.relpath(s, os.path.join(HERE, "csrc")).replace(os.sep, "__") + ".o")
        objs.append(o)
        if not force and os.path.exists(o) and os.path.getmtime(o) > max(os.path.getmtime(s), hdr_time):
            continue
        if s.endswith(".cu"):
            cmd = [NVCC] + NVCC_FLAGS + INCLUDES + ["-c", s, "-o", o]
        else:
            cmd = [CXX] + CXX_FLAGS + INCLUDES + ["-c", s, "-o", o]
        tasks.append((s, o, cmd))
    jobs = jobs or min(len(tasks) or 1, os.cpu_count() or 4)
    failed = False
    with concurrent.futures.ThreadPoolExecutor(max_workers=jobs) as ex:
        for src, rc, out in ex.map(lambda t: _compile(*t), tasks):
            if verbose:
                print("[build] %s %s" % ("ok  " if rc == 0 else "FAIL", os.path.relpath(src, ROOT)))
            if out.strip() and (rc != 0 or verbose):
                print(out)
            failed = failed or rc != 0
    if failed:
        raise RuntimeError("ncnn_b200 build failed")
    if tasks or not os.path.exists(LIB) or force:
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-cudart", "static", "-Xcompiler", "-pthread", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            print(r.stdout)
            raise RuntimeError("ncnn_b200 link failed")
        if verbose:
            print("[build] linked", os.path.relpath(LIB, ROOT))
    return LIB


if __name__ == "__main__":
    j = None
    if "-j" in sys.argv:
        j = int(sys.argv[sys.argv.index("-j") + 1])
    build(jobs=j, force="--force" in sys.argv)
