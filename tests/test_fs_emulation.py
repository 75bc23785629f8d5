"""The 3-frame banded DP with frameshifts and its traceback (diamond_b200/csrc/cuda/fs_kernels.cuh: fs_swipe_kernel<R, TRACE>, fs_walk_kernel)
compiled for the CPU behind tests/emu_cuda.h and run against the oracle's dmnd_banded_3frame_swipe (the restatement of
dp/swipe/banded_3frame_swipe.cpp that the frameshift goldens pin): scores, begin / end frames, coordinates, counts and transcripts of
random problems whose targets are stitched from pieces of the three reading frames, with corner bands and every register tile."""
import os, subprocess
from conftest import ROOT


def test_frameshift_kernels_emulation_matches_oracle(oracle_lib, tmp_path):
    exe = str(tmp_path / "emu_fs")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wno-unknown-pragmas", os.path.join(ROOT, "tests", "emu_fs.cpp"), "-o", exe,
                    "-L" + os.path.join(ROOT, "oracle", "_build"), "-ldmnd_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle", "_build")], check=True)
    # seed, longest read (nucleotides), widest band, problems, traceback (1) / score only (0)
    for args in (["1", "600", "128", "100"], ["2", "150", "40", "200"], ["3", "3000", "700", "24"], ["4", "600", "128", "100", "0"]):
        r = subprocess.run([exe] + args, capture_output=True, text=True)
        assert r.returncode == 0 and "fails=0 " in r.stdout, r.stdout + r.stderr
