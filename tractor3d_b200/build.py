"""Builds libt3d_particles.so in-tree with nvcc for sm_100a (the GPU box uses the prebuilt file that travels
with the snapshot; `python -m tractor3d_b200.build` rebuilds it anywhere nvcc is present)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libt3d_particles.so")
SOURCES = ["t3d_kernels.cu", "t3d_update.cu", "t3d_sprites.cu", "t3d_sort.cu", "t3d_runtime.cu", "t3d_host.cpp"]
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


def build_variant(name, defines):
    """Another build of the library with extra -D switches (experiments): csrc/_build/libt3d_particles_<name>.so, selected at
    run time with T3D_LIBRARY=<path>."""
    nvcc = find_nvcc()
    out_dir = os.path.join(CSRC, "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"libt3d_particles_{name}.so")
    cmd = [nvcc] + [f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")] + ["-D" + d for d in defines] + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed building variant {name}")
    return out


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


def build_multi_ctx_driver():
    """tests/multi_ctx_driver.cpp: one process, one thread, several contexts, through the C ABI only."""
    root = os.path.dirname(HERE)
    out_dir = os.path.join(HERE, "host", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "multi_ctx_driver")
    cmd = ["g++", "-std=c++17", "-O2", "-o", out, os.path.join(root, "tests", "multi_ctx_driver.cpp"),
           "-L" + CSRC, "-lt3d_particles", "-Wl,-rpath,$ORIGIN/../../csrc"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building the multi-context driver")
    return out


REFERENCE = os.environ.get("T3D_REFERENCE", "/root/reference")


def engine_caller_paths():
    root = os.path.dirname(HERE)
    return (os.path.join(root, "oracle", "_ref", "engine_caller_ref"), os.path.join(HERE, "host", "_build", "engine_caller_b200"))


def build_engine_callers():
    """tests/engine_caller.cpp -- application code written against the ENGINE's headers (what SceneLoader and
    ParticlesSample do with a ParticleEmitter) -- compiled twice: against the reference's own class and objects
    (oracle/_ref/engine_caller_ref), and with -DT3D_WITH_ENGINE_HEADERS against the facade
    (host/_build/engine_caller_b200: the engine's math / Ref / Properties objects + libt3d_particles.so).  Needs the
    reference tree and the objects `make -C oracle ref` leaves behind; returns None where the tree is absent (the GPU box
    runs the prebuilt programs)."""
    root = os.path.dirname(HERE)
    ref_inc = os.path.join(REFERENCE, "Tractor3Dlib", "include")
    obj_dir = os.path.join(root, "oracle", "_ref", "obj")
    if not os.path.isdir(ref_inc) or not os.path.isdir(obj_dir):
        return None
    out_ref, out_b200 = engine_caller_paths()
    tmp_ref = os.path.join(root, "oracle", "_ref", "engine")
    tmp_b200 = os.path.join(HERE, "host", "_build", "engine")
    os.makedirs(tmp_ref, exist_ok=True)
    os.makedirs(tmp_b200, exist_ok=True)
    overlay = ["-I" + os.path.join(root, "oracle", "ref_build", "overlay"), "-I" + os.path.join(root, "oracle", "ref_build", "shim"),
               "-I" + ref_inc, "-I" + os.path.join(root, "include")]
    cxx = ["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-w"]
    objs = []
    for dirpath, _, files in os.walk(obj_dir):
        # (ref_capi.o is the test harness; ui/Sprite.o needs the animation stubs that live in it and is not on this path)
        objs += [os.path.join(dirpath, f) for f in files if f.endswith(".o") and f != "ref_capi.o" and not (f == "Sprite.o" and os.path.basename(dirpath) == "ui")]
    objs.sort()

    def run(cmd):
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
            raise RuntimeError("building the engine-header callers failed")

    caller, support = os.path.join(root, "tests", "engine_caller.cpp"), os.path.join(root, "tests", "engine_support.cpp")
    run(cxx + overlay + ["-c", support, "-o", os.path.join(tmp_ref, "support.o")])
    # (1) the reference's own class
    run(cxx + overlay + ["-c", caller, "-o", os.path.join(tmp_ref, "caller.o")])
    run(["g++", "-o", out_ref, os.path.join(tmp_ref, "caller.o"), os.path.join(tmp_ref, "support.o")] + objs)
    # (2) the facade, found first on the include path as graphics/ParticleEmitter.h
    shim = ["-DT3D_WITH_ENGINE_HEADERS", "-I" + os.path.join(HERE, "host", "engine")]
    run(cxx + shim + overlay + ["-c", caller, "-o", os.path.join(tmp_b200, "caller.o")])
    run(cxx + shim + overlay + ["-c", os.path.join(HERE, "host", "ParticleEmitter.cpp"), "-o", os.path.join(tmp_b200, "facade.o")])
    engine_objs = [o for o in objs if os.path.basename(o) not in ("ParticleEmitter.o", "SpriteBatch.o", "TileSet.o", "Font.o")]
    run(["g++", "-o", out_b200, os.path.join(tmp_b200, "caller.o"), os.path.join(tmp_b200, "facade.o"), os.path.join(tmp_ref, "support.o")]
        + engine_objs + ["-L" + CSRC, "-lt3d_particles", "-Wl,-rpath,$ORIGIN/../../csrc"])
    return out_ref, out_b200


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(OUT)
