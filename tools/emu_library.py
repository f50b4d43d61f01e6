"""tools/emu_library.py -- builds libgb_b200_emu.so: the WHOLE library (host orchestration + kernels)
compiled for the host from the product sources, for tests only.

Every file of graphblas_b200/csrc is copied to a scratch tree with two textual edits
  * `kernel <<<grid, block[, smem[, stream]]>>> (args)` becomes
    `emu::launch_cfg (grid, block[, smem[, stream]], [&] { kernel (args) ; })`;
  * the one `extern __shared__` array becomes a pointer to the emulated dynamic shared memory;
and compiled with g++ against tests/emu/cuda_runtime.h (one OS thread per CUDA thread, device memory =
host memory, one "SM").  The result exports the same C ABI as libgb_b200.so, so that
tests/test_emulated_library.py can drive the real engine code through the real ctypes binding on a
machine without a GPU.  TEST INFRASTRUCTURE: the product never loads this library.

    python tools/emu_library.py --out /tmp/gb200_emu [--types int64,fp64,bool,int32]
"""
import argparse
import concurrent.futures
import hashlib
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS = os.path.join(ROOT, "graphblas_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "emu")
ALL_TYPES = ["bool", "int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "uint64", "fp32", "fp64"]


SYNC_TOKENS = ("__syncthreads", "__syncwarp", "__shfl", "__ballot_sync", "__any_sync", "__all_sync",
               "__reduce_add_sync", "__reduce_or_sync", "__match_")


def function_bodies(text: str) -> dict:
    """name -> body of every __global__ / __device__ function defined in `text`"""
    out = {}
    for m in re.finditer(r"__(?:global|device)__[^;{()]*?\b(\w+)\s*\(", text):
        name = m.group(1)
        depth, i = 0, m.end() - 1
        while True:                                 # the parameter list
            if text[i] == "(":
                depth += 1
            elif text[i] == ")":
                depth -= 1
                if depth == 0:
                    break
            i += 1
        j = i + 1
        while j < len(text) and text[j] not in "{;":
            j += 1
        if j >= len(text) or text[j] == ";":
            continue
        depth, k = 0, j
        while True:
            if text[k] == "{":
                depth += 1
            elif text[k] == "}":
                depth -= 1
                if depth == 0:
                    break
            k += 1
        out[name] = out.get(name, "") + text[j:k + 1]
    return out


def sync_free_kernels(texts) -> set:
    """kernels that neither synchronise nor call, at any depth, a device function that does"""
    bodies = {}
    for t in texts:
        for k, v in function_bodies(t).items():
            bodies[k] = bodies.get(k, "") + v
    syncing = {k for k, v in bodies.items() if any(tok in v for tok in SYNC_TOKENS)}
    changed = True
    while changed:
        changed = False
        for k, v in bodies.items():
            if k not in syncing and any(re.search(r"\b" + re.escape(f) + r"\b", v) for f in syncing):
                syncing.add(k)
                changed = True
    return set(bodies) - syncing


def rewrite_launches(src: str, name: str, seq=frozenset()) -> str:
    out, i = [], 0
    while True:
        k = src.find("<<<", i)
        if k < 0:
            out.append(src[i:])
            break
        j = k - 1
        while src[j].isspace():
            j -= 1
        if src[j] == ">":                       # template arguments of the kernel
            depth = 0
            while True:
                if src[j] == ">":
                    depth += 1
                elif src[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                j -= 1
            j -= 1
        while src[j].isalnum() or src[j] in "_:":
            j -= 1
        start = j + 1
        callee = src[start:k].strip()
        e = src.index(">>>", k)
        cfg = src[k + 3:e].strip()
        a = e + 3
        while src[a].isspace():
            a += 1
        assert src[a] == "(", f"{name}: no argument list after a launch of {callee}"
        depth, b = 0, a
        while True:
            if src[b] == "(":
                depth += 1
            elif src[b] == ")":
                depth -= 1
                if depth == 0:
                    break
            b += 1
        out.append(src[i:start])
        fn = "launch_cfg_seq" if re.match(r"\w+", callee).group(0) in seq else "launch_cfg"
        out.append(f"emu::{fn} ({cfg}, [&] {{ {callee} {src[a:b + 1]} ; }})")
        i = b + 1
    return "".join(out)


def source_hash() -> str:
    h = hashlib.sha256()
    for d in (CS, EMU, os.path.join(ROOT, "include")):
        for f in sorted(os.listdir(d)):
            p = os.path.join(d, f)
            if os.path.isfile(p):
                h.update(f.encode())
                h.update(open(p, "rb").read())
    h.update(open(__file__, "rb").read())
    return h.hexdigest()[:16]


def build(out: str, types, sanitize: bool = False) -> str:
    """sanitize: AddressSanitizer + UBSan build (load it with LD_PRELOAD=$(gcc -print-file-name=libasan.so)
    and ASAN_OPTIONS=detect_leaks=0): "device" memory is heap memory, so an out-of-bounds access of a kernel
    or of the engine is reported with a stack"""
    tag = source_hash() + "_" + "_".join(types) + ("_asan" if sanitize else "")
    san = ["-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-g"] if sanitize else []
    lib = os.path.join(out, "libgb_b200_emu.so")
    stamp = os.path.join(out, "stamp")
    if os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == tag:
        return lib
    shutil.rmtree(out, ignore_errors=True)
    cs = os.path.join(out, "root", "graphblas_b200", "csrc")
    os.makedirs(cs)
    os.makedirs(os.path.join(out, "root", "include"))
    shutil.copy(os.path.join(ROOT, "include", "gb_b200.h"), os.path.join(out, "root", "include"))
    # extern __shared__ [__align__ (n)] T name [] ;  ->  T *name = (T *) emu::dyn_smem ;
    ext = re.compile(r"extern\s+__shared__\s+(?:__align__\s*\(\s*\d+\s*\)\s*)?([A-Za-z_][A-Za-z_0-9 ]*?)\s+(\w+)\s*\[\s*\]\s*;")
    units = []
    seq = sync_free_kernels([open(os.path.join(CS, f)).read() for f in sorted(os.listdir(CS))
                             if f.endswith(".cu") or f.endswith(".cuh")])
    open(os.path.join(out, "sync_free_kernels.txt"), "w").write("\n".join(sorted(seq)) + "\n")
    for f in sorted(os.listdir(CS)):
        p = os.path.join(CS, f)
        if not os.path.isfile(p) or not (f.endswith(".cu") or f.endswith(".cuh")):
            continue
        if f.startswith("inst_") and f[5:-3] not in types:
            continue
        src = rewrite_launches(open(p).read(), f, seq)
        src = ext.sub(lambda m: f"{m.group (1)} *{m.group (2)} = ({m.group (1)} *) emu::dyn_smem ;", src)
        dst = os.path.join(cs, f[:-3] + ".cpp" if f.endswith(".cu") else f)
        open(dst, "w").write(src)
        if f.endswith(".cu"):
            units.append(dst)
    # the types left out of this build: their launchers decline
    missing = [t for t in ALL_TYPES if t not in types]
    stub = os.path.join(cs, "inst_missing.cpp")
    open(stub, "w").write('#include "kernels.cuh"\nnamespace gb200 {\n' + "".join(
        f"bool launch_{t} (int, int, int, int, const void *, LaunchCfg) {{ return false ; }}\n" for t in missing) + "}\n")
    units.append(stub)

    def cc(u):
        o = u[:-4] + ".o"
        subprocess.check_call(["g++", "-O1", "-std=c++20", "-fPIC", "-pthread", "-w", *san, "-I", EMU, "-c", u, "-o", o])
        return o
    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(cc, units))
    subprocess.check_call(["g++", "-shared", "-pthread", *san, "-Wl,-Bsymbolic", "-o", lib] + objs)   # binds its own gb200_* calls locally: the real library may be loaded RTLD_GLOBAL in the same process
    open(stamp, "w").write(tag)
    return lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="/tmp/gb200_emu")
    ap.add_argument("--types", default="bool,int32,int64,fp64")
    ap.add_argument("--sanitize", action="store_true")
    args = ap.parse_args()
    print(build(args.out, [t for t in args.types.split(",") if t], args.sanitize))


if __name__ == "__main__":
    main()
