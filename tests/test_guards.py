"""Our own memcheck: compute-sanitizer is closed on the GPU pool (profiles/r02_memcheck.log), so the end-to-end case of
tools/sanitize_case.py -- multi-chunk scoring, both kernel instantiations, compat calls, mesh normals, set consumers,
down-weighting -- runs with guard words behind every device buffer of the map (Q3DR_GUARD=1) and checks them."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_guards_are_off_by_default_and_say_so():
    from quad3dr_b200 import engine as E

    lib = E.load_library()
    assert lib.q3dr_map_check_guards(None, None) == E.Q3DR_ERR_INVALID_ARG


@pytest.mark.gpu
def test_no_device_buffer_is_written_past_its_end():
    env = dict(os.environ, Q3DR_GUARD="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_case.py")], capture_output=True, text=True,
                       timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "guard words behind all device buffers intact" in r.stdout and "sanitize case OK" in r.stdout
