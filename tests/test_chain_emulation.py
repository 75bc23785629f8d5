"""The device host-bridge code (diamond_b200/csrc/cuda/chain_kernels.cuh: hits of one (query, target) pair -> x-drop segments ->
greedy chaining -> merged DP bands) compiled for the CPU against the host restatement of the reference (csrc/host/chaining.cpp and
the band merge of pipeline.cpp) on thousands of random related sequence pairs with indels and internal repeats.  Runs without a
GPU; the pipeline-level check of the same path is every golden test of the CPU suite (the oracle library implements dmnd_hits_chain
through the host code, and the host pipeline consumes its records)."""
import os, subprocess
from conftest import ROOT


def test_chain_kernels_emulation_matches_host_chaining(tmp_path):
    exe = str(tmp_path / "emu_chain")
    host = os.path.join(ROOT, "diamond_b200", "csrc", "host")
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas", os.path.join(ROOT, "tests", "emu_chain.cpp"), os.path.join(host, "chaining.cpp"),
                    os.path.join(host, "scoring.cpp"), "-I" + os.path.join(ROOT, "include"), "-o", exe], check=True)
    for args in (["1", "4000"], ["2", "4000"], ["3", "4000"]):
        r = subprocess.run([exe] + args, capture_output=True, text=True)
        assert r.returncode == 0 and "fails=0 " in r.stdout, r.stdout + r.stderr
        assert "multi_segment=" in r.stdout and int(r.stdout.split("multi_segment=")[1].split()[0]) > 2000
