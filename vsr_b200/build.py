"""Build libvsr_sm100.so (hand-written sm_100a kernels + C-ABI) in-tree with nvcc.

    python -m vsr_b200.build [--force]

nvcc cross-compiles without a GPU; the .so lands in vsr_b200/lib/ (git-ignored, shipped to the
GPU box with the working tree).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBPATH = os.path.join(LIBDIR, "libvsr_sm100.so")
SOURCES = ["core.cu", "tapgemm_simt.cu", "tma_host.cu", "tapgemm_tc2.cu", "wgrad_tc.cu", "wgrad_shared.cu", "layers.cu", "lastconv_mma.cu", "firstconv_mma.cu", "metrics.cu",
           "resample.cu", "resample_int.cu", "dense3d.cu", "downscale.cu", "flow.cu", "split.cu", "toflow.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "vsr_b200.h"))
    return deps


def up_to_date():
    if not os.path.exists(LIBPATH):
        return False
    t = os.path.getmtime(LIBPATH)
    return all(os.path.getmtime(d) <= t for d in _deps())


def build(force=False, verbose=False, attrib=False):
    """attrib=True adds -DVSR_ATTRIB: the timing-attribution build of the tensor-core kernels (tools/attrib.py)"""
    if not force and up_to_date():
        return LIBPATH
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-DVSR_ATTRIB"] if attrib else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr:
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    cmd = [nvcc, "-shared", "-o", LIBPATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIBPATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv or "--attrib" in sys.argv, verbose=True, attrib="--attrib" in sys.argv))
