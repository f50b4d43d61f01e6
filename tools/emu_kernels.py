"""tools/emu_kernels.py -- the masked dot's CUDA kernels run on the HOST from their own source (no GPU).

kernels.cuh (dot_kernel, dotg_kernel with its cuckoo tables, regular and hub walks) and the set-up
kernels of engine_dot.cu are compiled with g++ against tests/emu/cuda_runtime.h, a stand-in in which one
CUDA thread is one OS thread, __syncthreads / warp intrinsics are barriers and atomics are the GCC
builtins.  The harness builds matrices with hub vectors (longer than one table load), a dense vector,
short vectors and split walks, runs the whole masked dot C<M>=A'*B exactly as run_dot sequences it
(classification with the trim -> lists -> tasks -> items -> dot_kernel for small pairs -> dotg_kernel for
hub and regular items of both orientations) and compares every C(i,j) and its presence flag with a plain
intersection loop.  The only edits made to the source text: the launchers at the end of kernels.cuh are
cut off (`<<< >>>` is not C++), and the `extern __shared__` array becomes a pointer to the emulated
dynamic shared memory.  The 32-byte PTX load has a plain-C twin under GB200_HOST_EMULATION (kernels.cuh).

The vector multiplies (streamed SpMV, masked pull, push: kernels_vec.cuh as run_dotv / run_saxpyv
sequence them) are covered the same way by tests/emu/emu_vec.cpp.

    python tools/emu_kernels.py [--cases 2] [--only emu_dotg|emu_vec] [--keep]
"""
import argparse
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS = os.path.join(ROOT, "graphblas_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "emu")


def cut(text: str, start: str) -> str:
    a = text.index(start)
    m = re.compile(r"^\}", re.M).search(text, a)
    return text[a:m.end()] + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=2)
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--only", default="", choices=["", "emu_dotg", "emu_vec"])
    args = ap.parse_args()
    kern = open(os.path.join(CS, "kernels.cuh")).read()
    eng = open(os.path.join(CS, "engine_dot.cu")).read()
    engv = open(os.path.join(CS, "engine_vec.cu")).read()
    stop = kern.index("// launchers, one set per (xy type)")
    stop = kern.rindex("// ----", 0, stop)
    kern = kern[:stop] + "\n} // namespace gb200\n"
    decl = "extern __shared__ __align__ (16) unsigned char dotg_raw [] ;"
    assert decl in kern
    kern = kern.replace(decl, "unsigned char *dotg_raw = emu::dyn_smem ;")
    setup = "".join([
        cut(eng, "__global__ void expand_vec_kernel"),
        cut(eng, "__global__ void dot_cum_list_kernel"),
        cut(eng, "__device__ __forceinline__ void dotg_trim"),
        cut(eng, "__global__ void dotg_classify_kernel"),
        cut(eng, "__global__ void dotg_lists_kernel"),
        cut(eng, "__global__ void dotg_tasks_kernel"),
        cut(eng, "__device__ __forceinline__ int64_t dotg_chunk_of"),
        cut(eng, "__global__ void dotg_nchunks_kernel"),
        cut(eng, "__global__ void dotg_items_kernel"),
    ])
    setup_vec = "".join([
        cut(engv, "__global__ void vec_nseg_kernel"),
        cut(engv, "__global__ void vec_items_kernel"),
        cut(engv, "__global__ void tile_row_kernel"),
    ])
    d = tempfile.mkdtemp(prefix="emu_kernels_")
    open(os.path.join(d, "kernels_emu.cuh"), "w").write(kern)
    open(os.path.join(d, "setup_emu.cuh"), "w").write("namespace gb200 {\n" + setup + "}\n")
    open(os.path.join(d, "setup_vec_emu.cuh"), "w").write("namespace gb200 {\n" + setup_vec + "}\n")
    rc = 0
    for name, argv in (("emu_dotg", [str(args.cases)]), ("emu_vec", [])):
        if args.only and args.only != name:
            continue
        exe = os.path.join(d, name)
        subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-w", "-I", EMU, "-I", CS, "-I", d,
                               "-o", exe, os.path.join(EMU, name + ".cpp")])
        rc |= subprocess.call([exe] + argv)
    if args.keep:
        print("sources kept in", d)
    sys.exit(rc)


if __name__ == "__main__":
    main()
