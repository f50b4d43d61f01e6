"""The masked dot's CUDA kernels (dot_kernel, dotg_kernel: cuckoo tables in shared memory, regular and hub
walks, split walks, pattern-only and valued paths, terminal exit) run on the HOST from their own source
text (tools/emu_kernels.py, tests/emu/): one OS thread per CUDA thread, barriers for __syncthreads and the
warp intrinsics.  Sequenced as run_dot sequences them and compared pair by pair with a plain intersection
loop.  Not a substitute for the GPU parity tests (the memory model and the PTX load are emulated), but it
checks the kernels' index logic without a GPU."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_masked_dot_kernels_on_the_host():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emu_kernels.py"), "--cases", "1"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.splitlines()[-1] == "emu_kernels: ok"
