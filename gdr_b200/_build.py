"""Build recipe for libgdr_b200.so (nvcc, sm_100a only, in-tree so the .so travels with gpurun)."""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIB_DIR, "libgdr_b200.so")
SOURCES = ["api.cu", "invert.cu", "score_simt.cu", "score_umma.cu", "score_umma_x2.cu", "score_tile_f32.cu", "similarity.cu", "score_fused.cu", "topk.cu", "topk_grouped.cu", "mask.cu", "tree.cu", "contrastive.cu", "xchg.cu", "partition.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]
# Developer builds only: `python -m gdr_b200._build --debug-knobs` compiles the stage-skipping timing experiments in
# (GDR_UMMA_DEBUG / GDR_TOPK_DEBUG environment masks; they invalidate results, so the product library does not contain them).
if os.environ.get("GDR_BUILD_UM_STAGES"):        # ring depth of the tcgen05 scoring CTA (default 6; the fused kernels need 6)
    FLAGS.append("-DGDR_UM_STAGES=" + str(int(os.environ["GDR_BUILD_UM_STAGES"])))
if "--debug-knobs" in sys.argv or os.environ.get("GDR_BUILD_DEBUG_KNOBS") == "1":
    FLAGS.append("-DGDR_DEBUG_KNOBS")


def _digest() -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)) + ["../../include/gdr_b200.h"]:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode() + f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a and link the shared library.  Returns its path."""
    os.makedirs(LIB_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, "build.sha256")
    digest = _digest()
    if not force and os.path.isfile(LIB) and os.path.isfile(stamp) and open(stamp).read() == digest:
        return LIB

    def compile_one(src):
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{log}")
        if verbose:
            print(log)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "shared"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
