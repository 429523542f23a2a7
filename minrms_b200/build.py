"""Build the sm_100a shared library in-tree (minrms_b200/libminrms_b200.so).

nvcc cross-compiles without a GPU.  -fmad=false keeps float arithmetic in the reference's
(non-contracted) form; div/sqrt stay IEEE (nvcc defaults)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "minrms_b200.cu")
VP_SRC = os.path.join(HERE, "csrc", "mr_dp_vp.cu")
GEN = os.path.join(HERE, "csrc", "_gen")
OUT = os.path.join(HERE, "libminrms_b200.so")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17"]
NVCC_FLAGS = ARCH_FLAGS + ["-Xcompiler", "-fPIC", "-shared"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    for root in (os.path.join(HERE, "csrc"), os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            if f != "_gen" and os.path.getmtime(os.path.join(root, f)) > t:
                return True
    return os.path.getmtime(os.path.join(HERE, "sass_patch.py")) > t


def build_vp_cubin(nvcc, extra, verbose=False):
    """csrc/mr_dp_vp.cu -> cubin -> sass_patch (VIMNMX.S16x2 -> .U32, predicate outputs kept) ->
    csrc/_gen/mr_vp_cubin.h, the byte array the library loads through the driver API."""
    from minrms_b200 import sass_patch
    os.makedirs(GEN, exist_ok=True)
    raw, out = os.path.join(GEN, "mr_dp_vp.raw.cubin"), os.path.join(GEN, "mr_dp_vp.cubin")
    subprocess.check_call([nvcc] + ARCH_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-cubin", "-o", raw, VP_SRC])
    info = sass_patch.patch(raw, out)
    sass_patch.write_header(out, os.path.join(GEN, "mr_vp_cubin.h"))
    return info


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("MR_NVCC_EXTRA", "").split()   # experiments only, e.g. -DMR_DP_MIN3
    out = os.environ.get("MR_BUILD_OUT", OUT)                # experiments only: variant builds for A/B runs
    build_vp_cubin(nvcc, extra, verbose)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, SRC]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
