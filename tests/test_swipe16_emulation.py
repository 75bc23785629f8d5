"""The packed 16-bit banded SWIPE kernel's SOURCE (diamond_b200/csrc/cuda/swipe16.cuh: swipe16_kernel<R, TRACE> and walk_kernel)
compiled for the CPU behind tests/emu_cuda.h and run four problems per warp in lock step against the oracle's
dmnd_banded_swipe: scores, coordinates, identities / mismatches / gap openings and transcripts of random problems with real
neighbouring sequences, Hauser-like biases, masked and ambiguous letters, corner bands, every register tile (R = 4, 8, 12, 16)
and warps whose problems differ in size.  Runs without a GPU."""
import os, subprocess
from conftest import ROOT


def test_packed_kernel_emulation_matches_oracle(oracle_lib, tmp_path):
    exe = str(tmp_path / "emu16")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wno-unknown-pragmas", os.path.join(ROOT, "tests", "emu_swipe16.cpp"), "-o", exe,
                    "-L" + os.path.join(ROOT, "oracle", "_build"), "-ldmnd_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle", "_build")], check=True)
    # seed, longest query, widest band, problems, traceback (1) / score only (0)
    for args in (["1", "300", "128", "240"], ["2", "60", "40", "400"], ["3", "900", "128", "80"], ["4", "300", "128", "200", "0"]):
        r = subprocess.run([exe] + args, capture_output=True, text=True)
        assert r.returncode == 0 and "fails=0 " in r.stdout, r.stdout + r.stderr
