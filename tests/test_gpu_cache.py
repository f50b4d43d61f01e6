"""Operand residency (include/gb_b200.h gb200_cache_*, SURVEY.md 8b "Residency", row f3) through the
unmodified GrB_mxm + shim: hits on repeated multiplies, never a stale copy (tests/cache_check.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_operand_cache_is_coherent():
    r = subprocess.run([sys.executable, os.path.join(HERE, "cache_check.py")], capture_output=True,
                       text=True, timeout=600, cwd=HERE, env={**os.environ, "PYTHONPATH": HERE})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "cache_check: ok" in r.stdout
