"""Build recipes.

  build_product()  nvcc -> imageworsener_b200/libimagew_b200.so   (sm_100a only; the shipped library)
  build_emulator() g++  -> tests/_emu/libimagew_emu.so            (TEST-ONLY host emulation of the kernel bodies,
                                                                   used to debug host logic on a GPU-less box;
                                                                   never imported by the package or the bench)
  build_oracle()   make -C oracle (C restatement + oracle/_ref when /root/reference is present)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
PRODUCT_SO = os.path.join(HERE, "libimagew_b200.so")
EMU_DIR = os.path.join(ROOT, "tests", "_emu")
EMU_SO = os.path.join(EMU_DIR, "libimagew_emu.so")

HOST_SOURCES = ["api.cpp", "plan.cpp", "weights.cpp", "tables.cpp", "process.cpp", "optimize.cpp", "util.cpp", "estimate.cpp"]
DEVICE_SOURCES = ["pipeline.cu", "stream2_v.cu", "stream2_v16.cu", "stream2_h1.cu", "stream2_h2.cu", "stream2_h3.cu", "stream2_h4.cu",
                  "gather4.cu", "gather5.cu", "stream3_e1.cu", "stream3_e2.cu", "stream3_e3.cu", "stream3_e4.cu"]
OBJ_DIR = os.path.join(HERE, "_obj")


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _all_deps() -> list[str]:
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(ROOT, "include", "imagew_b200.h"))
    deps.append(os.path.abspath(__file__))
    return deps


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build failed: " + cmd[0])
    if r.stderr.strip() and os.environ.get("IWB_BUILD_VERBOSE"):
        sys.stderr.write(r.stderr)


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def build_product(force: bool = False, verbose_ptxas: bool = False, out: str | None = None) -> str:
    """Each translation unit is compiled to an object in parallel (the kernel templates dominate the build
    time), then linked into one shared library.  IWB_QUICK=1 in the environment builds only the kernel variants
    of the headline configuration (tuning builds; never committed or shipped that way)."""
    target = out or PRODUCT_SO
    from concurrent.futures import ThreadPoolExecutor
    obj_dir = OBJ_DIR if not out else OBJ_DIR + "_" + os.path.splitext(os.path.basename(out))[0]
    os.makedirs(obj_dir, exist_ok=True)
    flags = ["-std=c++17", "-O3", "-lineinfo",
             "-gencode", "arch=compute_100a,code=sm_100a",
             "-fmad=false",                      # parity: the reference never fuses multiply-add
             "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden,-Wall",
             "-DIWB_BUILDING=1"]
    if os.environ.get("IWB_QUICK") == "1":
        flags.append("-DIWB_QUICK=1")
    for d in os.environ.get("IWB_DEFS", "").split():
        flags.append("-D" + d)
    # the library remembers how it was built: a tuning build (IWB_QUICK / IWB_DEFS) left in the tree must never be
    # mistaken for the product
    stamp = target + ".flags"
    flavour = " ".join(flags)
    same = os.path.exists(stamp) and open(stamp).read() == flavour
    if not force and same and _newer(target, _all_deps()):
        return target
    if verbose_ptxas:
        flags.append("-Xptxas=-v")

    def compile_one(name: str) -> str:
        obj = os.path.join(obj_dir, os.path.splitext(name)[0] + ".o")
        _run([nvcc_path()] + flags + ["-c", os.path.join(CSRC, name), "-o", obj])
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, DEVICE_SOURCES + HOST_SOURCES))
    _run([nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-shared",
          "-Xlinker", "-Bsymbolic", "-Xlinker", "--version-script=" + os.path.join(CSRC, "exports.map"),
          "-o", target] + objs)
    with open(stamp, "w") as fh:
        fh.write(flavour)
    return target


def build_emulator(force: bool = False) -> str:
    if not force and _newer(EMU_SO, _all_deps()):
        return EMU_SO
    os.makedirs(EMU_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in HOST_SOURCES]
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-fvisibility=hidden",
           "-Wl,-Bsymbolic", "-DIWB_EMU=1", "-DIWB_BUILDING=1", "-Wall", "-o", EMU_SO] + srcs
    cmd += ["-x", "c++", os.path.join(CSRC, "pipeline.cu")]   # kernel bodies as host loops (the stream2_*.cu units are device-only)
    _run(cmd)
    return EMU_SO


def build_oracle() -> None:
    _run(["make", "-C", os.path.join(ROOT, "oracle"), "all"])


if __name__ == "__main__":
    what = sys.argv[1:] or ["product", "emulator", "oracle"]
    if "product" in what:
        print(build_product(force=True, verbose_ptxas="-v" in what))
    if "emulator" in what:
        print(build_emulator(force=True))
    if "oracle" in what:
        build_oracle()
