"""The exp/ln the device oracles use (csrc/glibc_math.cuh) is plain C++ when compiled for the host: this builds
tools/check_glibc_math.c and checks 4 million arguments per function against this machine's libm, bit for bit."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_glibc_exp_log_port_matches_libm(tmp_path):
    exe = tmp_path / "check_glibc_math"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-mfma", "-x", "c++", os.path.join(ROOT, "tools", "check_glibc_math.c"),
                    "-o", str(exe), "-lm"], check=True)
    p = subprocess.run([str(exe), "4000000"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "exp mismatches 0, log mismatches 0" in p.stdout
