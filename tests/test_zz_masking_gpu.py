"""Masking on the device (dmnd_block_mask: tantan hard masking + motif soft-masking table, and their effect inside
dmnd_search_shape) against the oracle, and the GPU pipeline with the reference's DEFAULT flags against the L2 goldens
(tests/golden/*.l2.*, produced by the unmodified reference without --masking 0 --motif-masking 0).  Through the C ABI."""
import json, os, subprocess
import numpy as np
import pytest
from conftest import GOLDEN, ROOT, workload_blocks

pytestmark = pytest.mark.gpu


def sorted_hits(h):
    return np.sort(h, order=["query", "subject_score", "seed_offset"])


@pytest.mark.parametrize("name", ["rep", "edge", "c1"])
def test_block_mask_and_search_shape_match_oracle(oracle_lib, product_lib, name):
    """Same masked letters (bit-exact fp32 forward-backward pass), same positions, and -- the only way the motif table shows
    through the ABI -- identical seed-stage hits, counters and SEED_MASK bits with both blocks masked; whole-block and
    per-range masking (what the query lanes do) agree."""
    from diamond_b200 import api
    w, q_raw, q_lim, r_raw, r_lim = workload_blocks(name)
    nq, nr = len(q_lim) - 1, len(r_lim) - 1
    res = []
    for lib in (oracle_lib, product_lib):
        c = api.Context(lib, threads=8)
        qb, rb = c.upload(q_raw, q_lim), c.upload(r_raw, r_lim)
        rpos = c.mask_block(rb, 5, 0, nr)
        cuts = [0, nq // 3, nq // 3, (2 * nq) // 3 + 1, nq]  # ranges, one of them empty
        qpos = np.concatenate([c.mask_block(qb, 5, a, b) for a, b in zip(cuts[:-1], cuts[1:])])
        ql, rl = c.download_letters(qb, q_raw.size), c.download_letters(rb, r_raw.size)
        hits, cn = c.search_shape(qb, rb, 0)
        seeded = c.download_letters(qb, q_raw.size)
        c.clear_seed_mask(qb)
        cleared = c.download_letters(qb, q_raw.size)
        res.append((rpos, qpos, ql, rl, hits, cn, seeded, cleared))
        c.free_block(qb); c.free_block(rb); c.close()
    o, g = res
    assert np.array_equal(o[0], g[0]) and np.array_equal(o[1], g[1]), "hard-masked positions"
    assert np.all(np.diff(g[0].astype(np.int64)) > 0), "ABI promise: ascending offsets"
    assert np.array_equal(o[2], g[2]) and np.array_equal(o[3], g[3]), "letters after tantan"
    assert np.all(g[3][g[0].astype(np.int64)] == 23)
    assert len(o[4]) == len(g[4]) and np.array_equal(sorted_hits(o[4]), sorted_hits(g[4]))
    assert o[5] == g[5], "stage counters with motif soft masking"
    assert np.array_equal(o[6], g[6]), "SEED_MASK bits (entropy masking + MaskingTable::remove)"
    assert np.array_equal(o[7], g[7]) and np.array_equal(g[7], g[2])
    if name == "rep":
        assert len(g[0]) > 2000 and (g[6] != g[2]).sum() > 0


@pytest.mark.parametrize("name", ["c1", "fam2", "edge", "long", "rep"])
def test_blastp_default_flags_match_reference_golden(product_lib, name):
    from diamond_b200 import api
    w, q_raw, q_lim, r_raw, r_lim = workload_blocks(name)
    g = api.Context(product_lib, threads=8, comp_based_stats=1, masking=1, motif_masking=1)
    m, _, st = g.blastp(q_raw, q_lim, r_raw, r_lim)
    g.close()
    assert api.fmt6(m) == open(os.path.join(GOLDEN, f"{name}.l2.tsv")).read()
    cn = json.load(open(os.path.join(GOLDEN, f"{name}.l2.counters.json")))
    for k in ("seeds_hit", "seed_hits", "tentative_matches1", "tentative_matches2", "tentative_matches3"):
        assert st["seed"][k] == cn[k], k
    assert st["targets"] == cn["targets"] and st["dp_problems_round2"] == cn["targets_round2"]


def test_resident_masked_blocks_and_lanes(product_lib, monkeypatch):
    """dmnd_blastp_resident on blocks the caller masked (+ masked host images) equals the e2e call; three query lanes mask
    their own ranges concurrently on their own streams."""
    from diamond_b200 import api
    w, q_raw, q_lim, r_raw, r_lim = workload_blocks("rep")
    gold = open(os.path.join(GOLDEN, "rep.l2.tsv")).read()
    g = api.Context(product_lib, threads=8, comp_based_stats=1, masking=1, motif_masking=1)
    qb, rb = g.upload(q_raw, q_lim), g.upload(r_raw, r_lim)
    g.mask_block(qb, 5, 0, len(q_lim) - 1); g.mask_block(rb, 5, 0, len(r_lim) - 1)
    qm, rm = g.download_letters(qb, q_raw.size), g.download_letters(rb, r_raw.size)
    for _ in range(2):  # the motif table is a block property: a second step must not depend on leftovers of the first
        m, _, _ = g.blastp_resident(qb, rb, qm, q_lim, rm, r_lim)
        assert api.fmt6(m) == gold
    g.free_block(qb); g.free_block(rb)
    monkeypatch.setenv("DMND_LANES", "3")
    m, _, _ = g.blastp(q_raw, q_lim, r_raw, r_lim)
    assert api.fmt6(m) == gold
    g.close()


def test_cli_default_flags(product_lib, tmp_path):
    from diamond_b200 import synth
    w, *_ = workload_blocks("rep")
    q, d, o = (str(tmp_path / x) for x in ("q.faa", "d.faa", "o.tsv"))
    synth.write_fasta(q, w["q_letters"], w["q_off"], "q")
    synth.write_fasta(d, w["db_letters"], w["db_off"], "d")
    cli = os.path.join(ROOT, "diamond_b200", "bin", "dmnd-b200")
    r = subprocess.run([cli, "blastp", "--fast", "-q", q, "-d", d, "-o", o, "-p", "8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(o).read() == open(os.path.join(GOLDEN, "rep.l2.tsv")).read()


@pytest.mark.parametrize("name", ["c1", "long", "rep"])
def test_cli_transcript_fields(product_lib, name, tmp_path):
    """cigar / btop / gapped sequences from the device's traceback (and the masked letters) against the reference's columns."""
    from diamond_b200 import synth
    w, *_ = workload_blocks(name)
    q, d, o = (str(tmp_path / x) for x in ("q.faa", "d.faa", "o.tsv"))
    synth.write_fasta(q, w["q_letters"], w["q_off"], "q")
    synth.write_fasta(d, w["db_letters"], w["db_off"], "d")
    cli = os.path.join(ROOT, "diamond_b200", "bin", "dmnd-b200")
    fields = "qseqid sseqid pident length mismatch gapopen qstart qend sstart send evalue bitscore cigar btop qseq_gapped sseq_gapped".split()
    r = subprocess.run([cli, "blastp", "--fast", "-q", q, "-d", d, "-f", "6"] + fields + ["-o", o, "-p", "8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(o).read() == open(os.path.join(GOLDEN, f"{name}.t2.tsv")).read()


def test_cli_pairwise_format(product_lib, tmp_path):
    from diamond_b200 import synth
    w, *_ = workload_blocks("edge")
    q, d, o = (str(tmp_path / x) for x in ("q.faa", "d.faa", "o.txt"))
    synth.write_fasta(q, w["q_letters"], w["q_off"], "q")
    synth.write_fasta(d, w["db_letters"], w["db_off"], "d")
    cli = os.path.join(ROOT, "diamond_b200", "bin", "dmnd-b200")
    r = subprocess.run([cli, "blastp", "--fast", "-q", q, "-d", d, "-f", "0", "-o", o, "-p", "8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(o).read() == open(os.path.join(GOLDEN, "edge.f0.txt")).read()


def test_cli_makedb_matches_reference_database(product_lib, tmp_path):
    """makedb with the device's tantan: byte-identical to the reference's .dmnd (tests/golden/rep100.dmnd)."""
    from diamond_b200 import synth
    w, *_ = workload_blocks("rep")
    d100 = str(tmp_path / "d100.faa")
    synth.write_fasta(d100, w["db_letters"][: w["db_off"][100]], w["db_off"][:101], "d")
    cli = os.path.join(ROOT, "diamond_b200", "bin", "dmnd-b200")
    r = subprocess.run([cli, "makedb", "--in", d100, "-d", str(tmp_path / "ours")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(tmp_path / "ours.dmnd", "rb").read() == open(os.path.join(GOLDEN, "rep100.dmnd"), "rb").read()
