"""The C ABI is usable with no torch in the process: tools/sanitize_driver.py drives create / reset /
step_host / obs_from_packed / debug / export / import / destroy through ctypes + numpy only (both
tile sizes, both modes, the host-expand path)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_driver_runs_without_torch():
    code = ("import sys, runpy; sys.modules['torch'] = None; "
            f"runpy.run_path(r'{os.path.join(ROOT, 'tools', 'sanitize_driver.py')}', run_name='__main__')")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "sanitize driver finished" in r.stdout
