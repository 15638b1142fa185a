"""vx_accum_jump (csrc/vx_jump.h) must equal the serial f32 accumulation of the reference span loop
(rasterizer.rs:1458-1461) bit for bit.  Compiled for the host with gcc -ffp-contract=off; runs on CPU."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_jump_matches_serial_chain():
    exe = os.path.join(tempfile.mkdtemp(prefix="vxjump"), "jump_check")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "jump_check.c"), "-lm"])
    out = subprocess.run([exe, "700000"], capture_output=True, text=True, timeout=300)
    print(out.stdout)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "mismatches=0" in out.stdout
