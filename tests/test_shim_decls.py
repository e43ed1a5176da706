"""Boundary check: the public declarations of the drop-in headers equal the reference's token for token
(tools/check_shim_decls.py against the committed digest tests/golden/ref_decls.json; re-derived from the reference tree where it exists)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_declarations_equal_the_reference():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_shim_decls.py")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ORBextractor" in r.stdout and "EvImConverter" in r.stdout and "ORBmatcher" in r.stdout
