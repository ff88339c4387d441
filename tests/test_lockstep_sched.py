"""CPU check of the baton scheduler behind the lockstep solves (csrc/lockstep_sched.h): fake solves with random work,
2-4 fibers, refill from a list of points; invariants and a deadlock watchdog live in tests/lockstep_sched_test.cpp."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lockstep_scheduler_invariants():
    exe = os.path.join(tempfile.mkdtemp(prefix="bh_ls_"), "ls_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "lockstep_sched_test.cpp")])
    out = subprocess.run([exe, "60"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failed" in out.stdout
