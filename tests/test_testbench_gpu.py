"""GPU: the reference's own TestBench harness classes (source/test/*harness.cpp, compiled unmodified
into oracle/_ref/TestBench_b200_{8,10}) run against the EncoderPrimitives table filled by OUR
setupAssemblyPrimitives() -- the drop-in boundary of SURVEY.md 8b.  Exit code 0 = every installed slot
agreed bit for bit with the C reference on the harness's own fixtures."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("depth", [8, 10])
def test_reference_testbench_against_b200_table(depth):
    exe = os.path.join(ROOT, "oracle", "_ref", "TestBench_b200_%d" % depth)
    assert os.path.exists(exe), "oracle/_ref/TestBench_b200_* missing (built by oracle/Makefile where /root/reference exists)"
    for seed in ("0x5eed265", "0x1234"):
        r = subprocess.run([exe, "--seed", seed], capture_output=True, text=True, timeout=1500)
        tail = (r.stdout + r.stderr)[-3000:]
        assert r.returncode == 0, tail
        assert r.stdout.count("PASS") == 4, tail
