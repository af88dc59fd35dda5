"""In-tree build of libfmatch.so (nvcc, sm_100a only)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
LIB = os.path.join(_HERE, "libfmatch.so")


def sources():
    return sorted(os.path.join(CSRC, s) for s in os.listdir(CSRC) if s.endswith(".cu"))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "fastmatch_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """Compile csrc/*.cu into libfmatch.so.  `defines`/`out` build instrumented variants
    (e.g. -DFM_TC_PROF -> libfmatch_prof.so, used only by tools/)."""
    if out is None and not force and not needs_build():
        return LIB
    nvcc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-I" + INCLUDE, "-o", out or LIB]
    cmd += ["-D" + d for d in defines] + sources()
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    env = dict(os.environ)
    env.pop("CC", None)   # the image's CC wrapper is not meant for nvcc's host pass
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return out or LIB


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
