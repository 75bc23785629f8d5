"""CPU emulation of the profile-based banded SWIPE kernel (tests/emu_swipe_prof.cpp) against the oracle: thousands of
random problems with real neighbouring sequences on both sides of query and target, random int8 bias, masked letters,
corner bands and every register-tile width R.  Runs without a GPU."""
import os, subprocess
from conftest import ROOT


def test_profile_kernel_emulation_matches_oracle(oracle_lib, tmp_path):
    exe = str(tmp_path / "emu")
    subprocess.run(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "emu_swipe_prof.cpp"), "-o", exe,
                    "-L" + os.path.join(ROOT, "oracle", "_build"), "-ldmnd_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle", "_build")], check=True)
    for args in (["1", "300", "200", "1500"], ["2", "1500", "1000", "150"], ["3", "60", "40", "2500"]):
        out = subprocess.run([exe] + args, check=True, capture_output=True, text=True).stdout
        assert "fails=0 " in out, out
