"""gapped_filter_kernel's SOURCE (diamond_b200/csrc/cuda/gf_kernels.cuh) compiled for the CPU behind tests/emu_cuda.h and
checked against the oracle's dmnd_hits_gapped_filter on the hits of a --sensitive seed search (both outcomes occur)."""
import os, subprocess
from conftest import ROOT, workload_blocks


def test_gapped_filter_kernel_emulation_matches_oracle(oracle_lib, tmp_path):
    w, q_raw, q_lim, r_raw, r_lim = workload_blocks("edge")
    q_raw.tofile(str(tmp_path / "q.i8")); q_lim.tofile(str(tmp_path / "q.i64")); r_raw.tofile(str(tmp_path / "r.i8")); r_lim.tofile(str(tmp_path / "r.i64"))
    exe = str(tmp_path / "emu_gf")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wno-unknown-pragmas", os.path.join(ROOT, "tests", "emu_gf.cpp"), "-o", exe,
                    "-L" + os.path.join(ROOT, "oracle", "_build"), "-ldmnd_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle", "_build")], check=True)
    r = subprocess.run([exe, str(tmp_path), "400"], capture_output=True, text=True)
    assert r.returncode == 0 and "fails=0 " in r.stdout, r.stdout + r.stderr
    f = dict(kv.split("=") for kv in r.stdout.split() if "=" in kv)
    assert int(f["hits"]) > 100 and 0 < int(f["pass"]) < int(f["hits"])


def test_gapped_filter_kernel_emulation_translated_frames(oracle_lib, tmp_path):
    """blastx: cutoffs read at the length of the query's first frame, queries of < 100 letters pass after the 64-diagonal scan
    (align/gapped_filter.cpp:44-55)."""
    from diamond_b200 import api, synth
    f, kw = synth.BX_WORKLOADS["bx"]
    w = f(**kw)
    ql, qo = api.translate_reads(w["dna"])
    q_raw, q_lim = api.block_image(ql, qo)
    r_raw, r_lim = api.block_image(w["db_letters"], w["db_off"])
    q_raw.tofile(str(tmp_path / "q.i8")); q_lim.tofile(str(tmp_path / "q.i64")); r_raw.tofile(str(tmp_path / "r.i8")); r_lim.tofile(str(tmp_path / "r.i64"))
    exe = str(tmp_path / "emu_gf")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wno-unknown-pragmas", os.path.join(ROOT, "tests", "emu_gf.cpp"), "-o", exe,
                    "-L" + os.path.join(ROOT, "oracle", "_build"), "-ldmnd_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle", "_build")], check=True)
    r = subprocess.run([exe, str(tmp_path), "4000", "6"], capture_output=True, text=True)
    assert r.returncode == 0 and "fails=0 " in r.stdout, r.stdout + r.stderr
    f = dict(kv.split("=") for kv in r.stdout.split() if "=" in kv)
    assert int(f["hits"]) > 100 and 0 < int(f["pass"]) < int(f["hits"])
