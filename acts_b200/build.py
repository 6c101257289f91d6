"""Build helpers: compile the CUDA plugin (sm_100a) and the GPU-less model.

``python -m acts_b200.build`` builds everything in-tree so that the shared
objects travel with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "acts_b200", "csrc")
PLUGIN_SO = os.path.join(ROOT, "acts_b200", "libacts_b200_seeding.so")
MODEL_SO = os.path.join(ROOT, "tests", "model", "libgpu_model.so")

COMMON_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++20",
]
# exact binary32 replay of the reference: no FMA contraction, IEEE div/sqrt
EXACT_FLAGS = ["-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
               "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall"]
# relaxedFloat fast path: contraction and approximate reciprocals allowed on the device
RELAXED_FLAGS = ["-fmad=true", "-prec-div=false", "-prec-sqrt=false", "-ftz=true",
                 "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall"]
# The library holds two engines compiled from the same sources under private symbol
# prefixes / namespaces (csrc/engine_symbols.h) plus the exported ABI (csrc/seeding_abi.cpp).
ENGINES = {
    "exact": ["-DB200SEED_ENGINE_PREFIX=b200ex_", "-DB200SEED_NS=b200seed", *EXACT_FLAGS],
    "relaxed": ["-DB200SEED_ENGINE_PREFIX=b200rx_", "-DB200SEED_NS=b200seed_rx", "-DB200SEED_RELAXED=1", *RELAXED_FLAGS],
}
NVCC_FLAGS = [*COMMON_FLAGS, *EXACT_FLAGS, "-shared"]  # single-engine builds (tools/build_variant.py)
OBJ_DIR = os.path.join(ROOT, "build", "obj")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def plugin_sources():
    names = ["seeding_plugin.cu", "host_plan.cpp", "seeding_abi.cpp", "engine_symbols.h", "seeding_kernels.cuh",
             "orthogonal_kernels.cuh", "kd_tree_host.hpp", "seed_math.h", "host_plan.hpp"]
    return [os.path.join(CSRC, n) for n in names] + [os.path.join(ROOT, "include", "acts_b200_seeding.h")]


def engine_compile_commands(verbose: bool = False, extra=()):
    """nvcc -c command lines of the engine objects + the ABI object: [(object, argv)]."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    pre = ["-include", os.path.join(CSRC, "engine_symbols.h")]
    cmds = []
    for name, flags in ENGINES.items():
        for src in ("seeding_plugin.cu", "host_plan.cpp"):
            obj = os.path.join(OBJ_DIR, "%s_%s.o" % (name, src.split(".")[0]))
            ptxas = ["-Xptxas", "-v" if verbose else "-warn-spills"] if src.endswith(".cu") else []
            cmds.append((obj, ["nvcc", *COMMON_FLAGS, *flags, *extra, *pre, *ptxas, "-c", "-o", obj, os.path.join(CSRC, src)]))
    obj = os.path.join(OBJ_DIR, "seeding_abi.o")
    cmds.append((obj, ["nvcc", *COMMON_FLAGS, *EXACT_FLAGS, "-c", "-o", obj, os.path.join(CSRC, "seeding_abi.cpp")]))
    return cmds


def build_plugin(force: bool = False, verbose: bool = False) -> str:
    srcs = plugin_sources()
    if force or _stale(PLUGIN_SO, srcs):
        cmds = engine_compile_commands(verbose)
        procs = [(obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)) for obj, cmd in cmds]
        failed = False
        for obj, proc in procs:
            out, _ = proc.communicate()
            if verbose or proc.returncode != 0:
                sys.stderr.write(out)
            failed |= proc.returncode != 0
        if failed:
            raise RuntimeError("nvcc failed building the seeding plugin")
        link = ["nvcc", *COMMON_FLAGS, "-shared", "-o", PLUGIN_SO, *[obj for obj, _ in cmds]]
        res = subprocess.run(link, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("linking the seeding plugin failed")
    return PLUGIN_SO


def build_model(force: bool = False) -> str:
    src = os.path.join(ROOT, "tests", "model", "gpu_model.cpp")
    srcs = [src, os.path.join(CSRC, "seed_math.h")]
    if force or _stale(MODEL_SO, srcs):
        cmd = [os.environ.get("CXX", "g++"), "-O2", "-g", "-std=gnu++20", "-fPIC", "-ffp-contract=off", "-Wall",
               "-shared", "-o", MODEL_SO, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("g++ failed building the GPU-less model")
    return MODEL_SO


HOST_TEST_BIN = os.path.join(ROOT, "tests", "cpp", "host_mirror_main")


def build_host_mirror_test(force: bool = False) -> str:
    """C++20 driver of the host mirror class, linked against the plugin."""
    src = os.path.join(ROOT, "tests", "cpp", "host_mirror_main.cpp")
    hdr = os.path.join(ROOT, "acts_b200", "host", "GridTripletSeedingAlgorithm.hpp")
    hdr2 = os.path.join(ROOT, "acts_b200", "host", "OrthogonalTripletSeedingAlgorithm.hpp")
    build_plugin()
    if force or _stale(HOST_TEST_BIN, [src, hdr, hdr2, PLUGIN_SO]):
        cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=gnu++20", "-Wall", "-o", HOST_TEST_BIN, src,
               "-L" + os.path.dirname(PLUGIN_SO), "-lacts_b200_seeding", "-Wl,-rpath," + os.path.dirname(PLUGIN_SO)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("g++ failed building the host mirror test")
    return HOST_TEST_BIN


if __name__ == "__main__":
    print(build_plugin(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_model(force="--force" in sys.argv))
    print(build_host_mirror_test(force="--force" in sys.argv))
