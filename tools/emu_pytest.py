"""tools/emu_pytest.py -- run `-m gpu` tests on a machine WITHOUT a GPU against the emulated library.

    python tools/emu_pytest.py tests/test_gpu_seam.py -k "masked_dot"
    python tools/emu_pytest.py tests/test_gpu_parity.py tests/test_demo_programs.py -k "tricount or tri_demo"

Builds libgb_b200_emu.so (tools/emu_library.py: the product sources compiled for the host against
tests/emu/cuda_runtime.h) and the reference-side shim linked against it, points the ctypes binding, the
reference loader of the tests and the LD_PRELOAD of the demo-program tests at them, and hands the remaining
arguments to pytest with `-m gpu`.  A development aid and the back end of tests/test_emulated_parity.py;
a green run is not a GPU parity claim (see tests/emu/cuda_runtime.h)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import pytest
    import emulated
    emulated.library(global_scope=True)             # ahead of libgb_b200.so in the global symbol scope
    import graphblas_b200 as gb
    cm = emulated.swapped()
    cm.__enter__()                                  # for the life of this process
    so = emulated.shim()
    if so is not None:
        gb.SHIM_PATH = so
        import grbref
        grbref.SHIM_LIB = so
        import test_demo_programs
        test_demo_programs.SHIM = so
    sys.exit(pytest.main(["-m", "gpu", "-q", "-p", "no:cacheprovider"] + sys.argv[1:]))


if __name__ == "__main__":
    main()
