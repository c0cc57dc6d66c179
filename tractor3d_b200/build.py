"""Builds libt3d_particles.so in-tree with nvcc for sm_100a (the GPU box uses the prebuilt file that travels
with the snapshot; `python -m tractor3d_b200.build` rebuilds it anywhere nvcc is present)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libt3d_particles.so")
SOURCES = ["t3d_kernels.cu", "t3d_update.cu", "t3d_runtime.cu", "t3d_host.cpp"]
HEADERS = ["t3d_internal.h", "t3d_device.cuh", "t3d_host.h", os.path.join("..", "..", "include", "t3d_particles.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",                      # mul and add stay separate roundings, as in the reference's scalar code
    "-Xcompiler", "-fPIC,-ffp-contract=off",
    "-Xptxas", "-v",
    "-shared",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libt3d_particles.so (no fallback exists)")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libt3d_particles.so")
    if verbose:
        print(log)
    return OUT


def build_facade_driver():
    """Compiles the C++ ParticleEmitter facade (tractor3d_b200/host) together with tests/facade_driver.cpp against
    the library; proves the host-side mirror of the reference interface builds and links."""
    root = os.path.dirname(HERE)
    out_dir = os.path.join(HERE, "host", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "facade_driver")
    cmd = ["g++", "-std=c++17", "-O2", "-DGP_ERRORS_AS_WARNINGS", "-o", out,
           os.path.join(root, "tests", "facade_driver.cpp"), os.path.join(HERE, "host", "ParticleEmitter.cpp"),
           "-L" + CSRC, "-lt3d_particles", "-Wl,-rpath,$ORIGIN/../../csrc"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building the facade driver")
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(OUT)
