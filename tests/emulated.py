"""emulated.py -- test helper: builds libgb_b200_emu.so (tools/emu_library.py: the product sources compiled
for the host against tests/emu/cuda_runtime.h) and points the ctypes binding of graphblas_b200 at it for
the duration of a test.  The product package knows nothing about this; without the swap it only ever loads
libgb_b200.so and fails without a GPU."""
import contextlib
import ctypes as C
import importlib.util
import os
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_FUNCS = ("gb200_last_error", "gb200_version", "gb200_kernel_launches", "gb200_multiplies", "gb200_init",
          "gb200_finalize", "gb200_upload", "gb200_upload_from_device", "gb200_dmatrix_free", "gb200_AxB_device",
          "gb200_AxB_host", "gb200_result_get_info", "gb200_result_fetch", "gb200_result_free",
          "gb200_flopcount_device", "gb200_partition_by_flops", "gb200_semiring_canonical",
          "gb200_device_count", "gb200_timer_mark", "gb200_timer_elapsed_ms", "gb200_host_malloc",
          "gb200_host_free", "gb200_host_trim", "gb200_select_host", "gb200_transpose_host")
_lib = None


def _sanitize():
    """GB200_EMU_SANITIZE=1: the AddressSanitizer + UBSan build (tools/emu_library.py); run python with
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0"""
    return os.environ.get("GB200_EMU_SANITIZE", "") not in ("", "0")


def _outdir():
    return os.path.join(tempfile.gettempdir(), f"gb200_emu_{os.getuid()}" + ("_asan" if _sanitize() else ""))


def library(types="bool,int8,uint8,int16,uint16,int32,uint32,int64,uint64,fp32,fp64", global_scope=False):
    """global_scope: load with RTLD_GLOBAL -- needed (and must happen BEFORE graphblas_b200 is imported) when the
    emulated shim is used in this process, so that the shim's gb200_* references find this library first"""
    global _lib
    if _lib is None:
        spec = importlib.util.spec_from_file_location("emu_library", os.path.join(ROOT, "tools", "emu_library.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _lib = C.CDLL(mod.build(_outdir(), [t for t in types.split(",") if t], _sanitize()),
                      mode=C.RTLD_GLOBAL if global_scope else C.DEFAULT_MODE)
    return _lib


def shim(ref="/root/reference"):
    """the reference-side binding (csrc/shim/gb_axb_parallel_shim.c) linked against the emulated library;
    None where the reference's headers are absent"""
    if not os.path.isdir(os.path.join(ref, "Source")):
        return None
    library()
    out = _outdir()
    so = os.path.join(out, "libgb_b200_shim_emu.so")
    src = os.path.join(ROOT, "graphblas_b200", "csrc", "shim", "gb_axb_parallel_shim.c")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(out, "libgb_b200_emu.so"))):
        import subprocess
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-DNDEBUG", "-Wno-pragmas",
                               "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ref, "Source", "Template"),
                               "-I", os.path.join(ref, "Source"), "-I", os.path.join(ref, "Include"),
                               "-fvisibility=hidden", "-o", so, src, "-L", out, "-l:libgb_b200_emu.so",
                               "-Wl,-rpath," + out, "-ldl"])
    return so


@contextlib.contextmanager
def swapped():
    """graphblas_b200.lib -> the emulated library, inside the with block"""
    import graphblas_b200 as gb
    emu, real = library(), gb.lib
    for name in tuple(_FUNCS) + tuple(getattr(gb, "_SIGS", {})):
        if not hasattr(emu, name):
            continue
        f, g = getattr(real, name), getattr(emu, name)
        g.restype = f.restype
        if f.argtypes is not None:
            g.argtypes = f.argtypes
    gb.lib = emu
    try:
        yield gb
    finally:
        gb.lib = real
