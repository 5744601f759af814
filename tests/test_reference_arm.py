"""T0: the reference's OWN test-suite, run against the copy `bench.py --impl reference` times (oracle/_ref,
installed unmodified with its C extension by __graft_entry__.build).  It pins that the reference arm is the reference:
math types of both back ends, DynamicTree query == brute-force overlap, ray casts.  Needs the reference tree (build
container only); skipped where /root/reference is absent."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REF_TESTS = "/root/reference/tests"
REF_ARM = os.path.join(ROOT, "oracle", "_ref")


@pytest.mark.skipif(not (os.path.isdir(REF_TESTS) and os.path.exists(os.path.join(REF_ARM, "pypybox2d", "__init__.py"))),
                    reason="reference tree or oracle/_ref not present")
def test_reference_own_tests_pass_on_the_installed_reference_arm(tmp_path):
    files = sorted(f for f in os.listdir(REF_TESTS) if f.endswith("_test.py") or f.startswith("test_"))
    assert files, "no reference tests found"
    env = dict(os.environ, PYTHONPATH=REF_ARM)
    ran = 0
    for f in files:
        out = subprocess.run([sys.executable, os.path.join(REF_TESTS, f)], cwd=str(tmp_path), env=env, stdout=subprocess.PIPE,
                             stderr=subprocess.STDOUT, text=True, timeout=600)
        assert out.returncode == 0, "%s failed on oracle/_ref:\n%s" % (f, out.stdout[-3000:])
        assert "OK" in out.stdout, "%s: %s" % (f, out.stdout[-500:])
        ran += 1
    assert ran == len(files)
