"""The set-up kernels of the masked dot (classification with the trim, pair lists, task records) are
per-element loops: tools/emu_setup.py compiles their SOURCE TEXT with g++ behind a few macros and checks
on random standard / hypersparse inputs that the tasks tile exactly the part of every walked list that
lies inside [first, last] of its owner, that split pairs are marked, that small pairs go to the small
list and that only pairs with an empty trimmed walk are dropped.  Runs without a GPU."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_setup_kernels_on_the_host():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emu_setup.py"), "--cases", "80"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok" in r.stdout.splitlines()[-1]
