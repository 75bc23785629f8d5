"""The seed stage's kernel SOURCE (diamond_b200/csrc/cuda/seed_kernels.cuh: ref_enum, bloom/bucket build, probe, entropy
masking, stage 1 flags, stage-2 window + left-most filter, motif SEED_MASK marking) compiled for the CPU behind
tests/emu_cuda.h and driven by a mirror of build_ref_index + search_shape_impl, against the oracle's dmnd_search_shape:
hits incl. ungapped window scores, stage counters and SEED_MASK bits, for BOTH shapes of the default sensitivity on masked
blocks (bits of shape 0 stay set for shape 1).  Runs without a GPU.  (`emu_seed DIR 0 0` = --fast, `DIR 1 0` on the fam2
workload = 339 104 hits incl. the 255-capped scores, 4 minutes: run by hand.)"""
import os, subprocess
from conftest import ROOT, workload_blocks


def test_seed_stage_emulation_matches_oracle_default_sensitivity(oracle_lib, tmp_path):
    from diamond_b200 import api
    w, q_raw, q_lim, _, _ = workload_blocks("edge")
    r_raw, r_lim = api.block_image(w["db_letters"][:w["db_off"][800]], w["db_off"][:801])  # 800 of the 2 000 proteins: the emulation's cost is one coroutine per reference position
    q_raw.tofile(str(tmp_path / "q.i8")); q_lim.tofile(str(tmp_path / "q.i64")); r_raw.tofile(str(tmp_path / "r.i8")); r_lim.tofile(str(tmp_path / "r.i64"))
    exe = str(tmp_path / "emu_seed")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wno-unknown-pragmas", os.path.join(ROOT, "tests", "emu_seed.cpp"), "-o", exe,
                    "-L" + os.path.join(ROOT, "oracle", "_build"), "-ldmnd_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle", "_build")], check=True)
    r = subprocess.run([exe, str(tmp_path), "1", "1"], capture_output=True, text=True)
    assert r.returncode == 0 and "shapes=2 " in r.stdout and "fails=0 " in r.stdout, r.stdout + r.stderr


def test_seed_stage_emulation_translated_frames(oracle_lib, tmp_path):
    """blastx at the default sensitivity: frames of <= 85 letters take the whole-frame window (search/stage2.h:58-63) in the
    stage-2 kernel; the query block = six translated contexts of the first reads of the `bx` workload."""
    from diamond_b200 import api, synth
    f, kw = synth.BX_WORKLOADS["bx"]
    w = f(**kw)
    ql, qo = api.translate_reads(w["dna"][:32])
    q_raw, q_lim = api.block_image(ql, qo)
    r_raw, r_lim = api.block_image(w["db_letters"][:w["db_off"][500]], w["db_off"][:501])  # (the emulation's cost is the reference side)
    assert ((qo[1:] - qo[:-1]) <= 85).mean() > 0.5 and ((qo[1:] - qo[:-1]) > 85).any()
    q_raw.tofile(str(tmp_path / "q.i8")); q_lim.tofile(str(tmp_path / "q.i64")); r_raw.tofile(str(tmp_path / "r.i8")); r_lim.tofile(str(tmp_path / "r.i64"))
    exe = str(tmp_path / "emu_seed")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wno-unknown-pragmas", os.path.join(ROOT, "tests", "emu_seed.cpp"), "-o", exe,
                    "-L" + os.path.join(ROOT, "oracle", "_build"), "-ldmnd_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle", "_build")], check=True)
    r = subprocess.run([exe, str(tmp_path), "1", "1", "6"], capture_output=True, text=True)
    assert r.returncode == 0 and "shapes=2 " in r.stdout and "fails=0 " in r.stdout, r.stdout + r.stderr
