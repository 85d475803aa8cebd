"""In-tree build of the CUDA library (sm_100a only; nvcc cross-compiles without a GPU).

The hot-kernel instantiations (csrc/fast_inst.cu, one translation unit per SB_PART) and the host/C-ABI
unit (csrc/stereo_b200.cu) compile in parallel into csrc/_build/*.o and are linked into
libstereo_b200.so next to this file."""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
# Experiment builds: STEREO_BUILD_TAG=<tag> compiles into csrc/_build_<tag>/ and links libstereo_b200_<tag>.so (loaded with
# STEREO_LIB_TAG=<tag>, _capi.py); SB_* environment variables become -DSB_*=<value> tuning macros of the kernels.
TAG = os.environ.get("STEREO_BUILD_TAG", "")
OBJ = (Path("/tmp") / f"sb_build_{TAG}") if TAG else CSRC / "_build"        # experiment objects stay out of the tree
LIB = PKG / ("libstereo_b200" + (f"_{TAG}" if TAG else "") + ".so")
FAST_PARTS = 48          # 16 cost x radius x strips-per-warp parts + 2 x 5 fused SSD parts + 3 x 4 float-operand parts + 2 x 5 fused NCC parts

NVCC_FLAGS = [
    *[f"-D{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("SB_") and v != ""],
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-diag-suppress", "177,128",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (needed to build libstereo_b200.so for sm_100a)")


def units():
    """(object name, source, extra flags) of every translation unit."""
    u = [("stereo_b200.o", CSRC / "stereo_b200.cu", []), ("host_pack.o", CSRC / "host_pack.cpp", []),
         ("mgpu.o", CSRC / "mgpu.cpp", [])]
    u += [(f"fast_inst_{i}.o", CSRC / "fast_inst.cu", [f"-DSB_PART={i}"]) for i in range(FAST_PARTS)]
    return u


def _mkdirs():
    OBJ.mkdir(exist_ok=True)


def _deps(src: Path):
    """Files a translation unit is rebuilt for: the hot-kernel parts only include fast_kernel.cuh / common.cuh."""
    api = list((PKG.parent / "include").glob("*.h"))
    if src.name == "fast_inst.cu":
        return [src, CSRC / "fast_kernel.cuh", CSRC / "common.cuh", *api]
    if src.name == "host_pack.cpp":
        return [src, CSRC / "host_pack.hpp"]
    if src.name == "mgpu.cpp":
        return [src, *api]
    return [*CSRC.glob("*.cu"), *CSRC.glob("*.cuh"), *CSRC.glob("*.hpp"), *api]


def sources():
    return sorted(CSRC.glob("*.cu"))


def needs_build() -> bool:
    if TAG:
        return True
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [*CSRC.glob("*.cu"), *CSRC.glob("*.cuh"), *CSRC.glob("*.cpp"), *CSRC.glob("*.hpp"), *(PKG.parent / "include").glob("*.h")]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/*.cu into libstereo_b200.so next to this file."""
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    OBJ.mkdir(exist_ok=True)
    # experiment builds may recompile only some hot-kernel parts (STEREO_BUILD_PARTS="18,22") and borrow the other objects
    # from the default build - such a library is only valid for the configurations those parts serve
    only = os.environ.get("STEREO_BUILD_PARTS", "")
    borrowed = set()
    if TAG and only:
        keep = {f"fast_inst_{int(i)}.o" for i in only.split(",")} | {"stereo_b200.o", "host_pack.o", "mgpu.o"}
        for name, _, _ in units():
            if name not in keep:
                src_obj = CSRC / "_build" / name
                if not src_obj.exists():
                    raise RuntimeError(f"{src_obj} missing: build the default library first")
                shutil.copy2(src_obj, OBJ / name)
                borrowed.add(name)

    def compile_one(unit):
        name, src, extra = unit
        if name in borrowed:
            return ""
        obj = OBJ / name
        if not force and obj.exists() and all(d.stat().st_mtime <= obj.stat().st_mtime for d in _deps(src)):
            return ""
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", "-o", str(OBJ / name), str(src)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name} {extra}:\n" + res.stdout + res.stderr)
        return res.stderr

    with ThreadPoolExecutor(max_workers=min(len(units()), os.cpu_count() or 1)) as ex:
        logs = list(ex.map(compile_one, units()))
    if verbose:
        print("\n".join(logs))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB), *(str(OBJ / n) for n, _, _ in units())]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
