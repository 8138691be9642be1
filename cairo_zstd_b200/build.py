"""Builds libcairo_zstd_b200.so in-tree with nvcc for sm_100a (no torch dependency)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcairo_zstd_b200.so")
SOURCES = ["czb_api.cu", "czb_handle.cu", "k_scan.cu", "k_huff.cu", "k_fse.cu", "k_exec.cu", "k_exec_flow.cu", "k_xxh.cu"]
HEADERS = ["czb_internal.cuh", "czb_parse.cuh", "czb_fse_build.cuh", "czb_host.h", "czb_exec.cuh", "k_exec_flow_impl.cuh",
           "../../include/cairo_zstd_b200.h", "../../include/czstd_status.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--use_fast_math"]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, lib_out=None, objdir="build"):
    """lib_out / objdir: build a variant (other -D flags through CZB_NVCC_FLAGS) next to the product library, for A/B runs:
    CZB_LIB=<lib_out> makes api.load_library() pick it up."""
    lib_out = lib_out or os.environ.get("CZB_LIB_OUT") or LIB
    if lib_out != LIB:
        force, objdir = True, "build/" + os.path.basename(lib_out).replace(".so", "")
    if not (force or stale()):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, objdir), exist_ok=True)
    for s in SOURCES:
        o = os.path.join(HERE, objdir, s.replace(".cu", ".o"))
        objs.append(o)
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get("CZB_NVCC_FLAGS", "").split() + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    fail = False
    for s, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {s}\n{out}")
        if p.returncode != 0:
            fail = True
    text = "\n".join(log)
    open(os.path.join(HERE, objdir, "ptxas.log"), "w").write(text)
    if fail or verbose:
        print(text, file=sys.stderr if fail else sys.stdout)
    if fail:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib_out] + objs)
    return lib_out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)
