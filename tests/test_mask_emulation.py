"""The masking kernels' SOURCE (diamond_b200/csrc/cuda/mask_kernels.cuh: tantan_kernel, popc_kernel, motif_hit_kernel,
motif_apply_kernel, clear_bits_kernel, motif_seedmask_kernel) compiled for the CPU behind tests/emu_cuda.h -- coroutine
threads, real lane-to-lane shuffles with the kernels' own masks, the library's grid/block sizes -- and checked against the
oracle on the `rep` workload (tandem repeats, motifs at sequence ends, > 50 % covered and sub-motif-length sequences, X).
Runs without a GPU; the same comparison runs on the device in tests/test_zz_masking_gpu.py."""
import os, subprocess
from conftest import ROOT, workload_blocks


def test_masking_kernels_emulation_matches_oracle(oracle_lib, tmp_path):
    w, q_raw, q_lim, r_raw, r_lim = workload_blocks("rep")
    r_raw.tofile(str(tmp_path / "raw.i8")); r_lim.tofile(str(tmp_path / "lim.i64"))
    exe = str(tmp_path / "emu_mask")
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas", os.path.join(ROOT, "tests", "emu_mask.cpp"), "-o", exe,
                    "-L" + os.path.join(ROOT, "oracle", "_build"), "-ldmnd_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle", "_build")], check=True)
    r = subprocess.run([exe, str(tmp_path), "100"], capture_output=True, text=True)
    assert r.returncode == 0 and "fails=0 " in r.stdout, r.stdout + r.stderr
    f = dict(kv.split("=") for kv in r.stdout.split() if "=" in kv)
    assert int(f["tantan_masked"]) > 1000 and int(f["soft_letters"]) > 30 and int(f["seed_mask_positions"]) > int(f["soft_letters"])


def test_masking_kernels_emulation_translated_frames(oracle_lib, tmp_path):
    """The query block of a blastx run: six translated frames per read -- stop codons (letter 24) inside sequences, X-ed out
    ORFs, empty sequences (reads shorter than a codon), a trinucleotide repeat that tantan masks in a frame."""
    from diamond_b200 import api, synth
    f, kw = synth.BX_WORKLOADS["bx"]
    ql, qo = api.translate_reads(f(**kw)["dna"][:160])
    assert ((qo[1:] - qo[:-1]) == 0).any() and (ql == 24).any()
    raw, lim = api.block_image(ql, qo)
    raw.tofile(str(tmp_path / "raw.i8")); lim.tofile(str(tmp_path / "lim.i64"))
    exe = str(tmp_path / "emu_mask")
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas", os.path.join(ROOT, "tests", "emu_mask.cpp"), "-o", exe,
                    "-L" + os.path.join(ROOT, "oracle", "_build"), "-ldmnd_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle", "_build")], check=True)
    r = subprocess.run([exe, str(tmp_path), "100"], capture_output=True, text=True)
    assert r.returncode == 0 and "fails=0 " in r.stdout, r.stdout + r.stderr
    f = dict(kv.split("=") for kv in r.stdout.split() if "=" in kv)
    assert int(f["tantan_masked"]) > 0
