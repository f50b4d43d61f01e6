"""The CUDA kernels of the hot path run on the HOST from their own source text (tools/emu_kernels.py,
tests/emu/): one OS thread per CUDA thread, a block barrier for __syncthreads, a rendezvous of the named
lanes for the warp intrinsics, GCC atomics.
  * masked dot C<M>=A'*B: dot_kernel and dotg_kernel (cuckoo tables in shared memory, regular and hub
    walks, split walks, pattern-only and valued paths, terminal exit) sequenced as run_dot sequences them,
    compared pair by pair with a plain intersection loop;
  * vector multiplies: streamed SpMV (SSSP), masked pull, push with a complemented mask (BFS) as run_dotv /
    run_saxpyv sequence them, compared with plain loops.
Not a substitute for the GPU parity tests (the memory model and the PTX load are emulated), but it checks
the kernels' index logic without a GPU."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("which,last", [("emu_dotg", "emu_kernels: ok"), ("emu_vec", "emu_vec: ok")])
def test_kernels_on_the_host(which, last):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emu_kernels.py"), "--cases", "1",
                        "--only", which], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.splitlines()[-1] == last
