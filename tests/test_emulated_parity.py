"""The reference-API parity tests on the host: the UNMODIFIED compiled reference (`oracle/_ref`), its
`GB_AxB_parallel` taken over by the shim, the shim linked against the emulated library
(tools/emu_library.py) -- so `GrB_mxm` / `GrB_vxm` / `GrB_mxv` and the reference's own demo binaries run
the product's engine code and kernels on the host, compared with the reference's own result exactly as in
`tests/test_gpu_parity.py` / `tests/test_demo_programs.py`.  Runs in a subprocess (`tools/emu_pytest.py`):
the emulated library has to be first in the global symbol scope, which the test process cannot offer once
another test has loaded the real one.  A sample sized for the CPU suite; see tests/test_emulated_library.py
for what an emulated run does and does not show."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(files, expr):
    if not os.path.isdir("/root/reference/Source"):
        pytest.skip("the shim is compiled against the reference's headers, which are not here")
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libgraphblas_ref.so")):
        pytest.skip("oracle/_ref not built")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emu_pytest.py"), *files, "-x", "-k", expr],
                       capture_output=True, text=True, timeout=1500, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail
    return r.stdout


def test_grb_api_parity_on_the_host():
    out = _run(["tests/test_gpu_parity.py"],
               "test_tricount and INT64-8 or test_masked or test_hypersparse or test_bfs_levels or "
               "test_sssp_bellman_ford or test_ktruss_iterations or test_empty_and_ragged or test_aliased_C_is_A")
    assert "deselected" in out


def test_reference_demo_programs_on_the_host():
    """tri_demo, bfs_demo, mis_demo of the reference, unmodified, under LD_PRELOAD of the emulated shim"""
    _run(["tests/test_demo_programs.py"], "args0 or args1")       # the small inputs: Wathen 4x4, random 5x5


def test_neighbours_of_the_multiply_on_the_host():
    """rows f1 / f2: the interposed GB_transpose and GB_accum_mask (device transpose by radix sort, device
    accum / mask) against the reference through GrB_transpose / GrB_mxm, a sample"""
    out = _run(["tests/test_gpu_parity.py"],
               "test_transpose_seam and between or test_grb_transpose_with_mask_and_accum or "
               "test_accum_mask_typecasts or test_accum_mask_vectors_bfs_and_sssp_steps or "
               "test_assign_scalar_bfs_level_step or "
               "test_no_neighbour_call_failed_on_the_device")
    assert "deselected" in out
