"""The reference's DEFAULT sensitivity on the device: both shapes through dmnd_search_shape (stage-2 ungapped window filter,
call-size dependent score cap, multi-shape left-most filter) against the oracle, and the GPU pipeline / CLI against the S1
goldens (the reference run with no sensitivity flag).  Through the C ABI."""
import json, os, subprocess
import numpy as np
import pytest
from conftest import GOLDEN, ROOT, workload_blocks

# History: in round 1 these failed on the B200 (2-5 hits of ~3 100 too many for shape 1, different from run to run) although the same
# kernel source passed under the CPU emulation.  Round 2 traced it to the code ptxas generated for the left-most filter when it was
# inlined into stage2_window_kernel (profiles/lm_variants_r2.txt); the filter is a real call now and the tests are ordinary tests.
pytestmark = pytest.mark.gpu

def sorted_hits(h):
    return np.sort(h, order=["query", "subject_score", "seed_offset"])


def hit_diff(ho, hg, r_lim):
    """Readable difference of two hit lists for the assertion message: hits only one side has (ignoring the score), and hits
    both have with different scores -- (query, seed offset, target, target position, score)."""
    def key(h):
        return {(int(x["query"]), int(x["seed_offset"]), int(x["subject_score"]) & ((1 << 48) - 1)): int(x["subject_score"]) >> 48 for x in h}
    a, b = key(ho), key(hg)
    def show(k, sc):
        t = int(np.searchsorted(r_lim, k[2], side="right")) - 1
        return f"(q{k[0]} off {k[1]} d{t}+{k[2] - int(r_lim[t])} score {sc})"
    only_o = [show(k, a[k]) for k in sorted(set(a) - set(b))][:8]
    only_g = [show(k, b[k]) for k in sorted(set(b) - set(a))][:8]
    other = [show(k, f"{a[k]} vs {b[k]}") for k in sorted(set(a) & set(b)) if a[k] != b[k]][:8]
    return f"hits: oracle {len(ho)} device {len(hg)}; only oracle {len(set(a) - set(b))}: {only_o}; only device {len(set(b) - set(a))}: {only_g}; other score {sum(a[k] != b[k] for k in set(a) & set(b))}: {other}"


@pytest.mark.parametrize("name,masking", [("fam2", 0), ("rep", 1), ("edge", 1)])
def test_search_shapes_match_oracle(oracle_lib, product_lib, name, masking):
    from diamond_b200 import api
    w, q_raw, q_lim, r_raw, r_lim = workload_blocks(name)
    res = []
    for lib in (oracle_lib, product_lib):
        c = api.Context(lib, threads=8, sensitivity=1)
        qb, rb = c.upload(q_raw, q_lim), c.upload(r_raw, r_lim)
        if masking:
            c.mask_block(qb, 5, 0, len(q_lim) - 1); c.mask_block(rb, 5, 0, len(r_lim) - 1)
        out = []
        for sid in (0, 1):  # SEED_MASK bits of shape 0 stay set for shape 1 (run/double_indexed.cpp:185-214)
            hits, cn = c.search_shape(qb, rb, sid)
            out.append((hits, cn, c.download_letters(qb, q_raw.size)))
        res.append(out)
        c.free_block(qb); c.free_block(rb); c.close()
    for sid in (0, 1):
        (ho, co, lo), (hg, cg, lg) = res[0][sid], res[1][sid]
        assert co == cg, f"shape {sid}: stage counters oracle {co} device {cg}\n" + hit_diff(ho, hg, r_lim)
        assert len(ho) == len(hg) and np.array_equal(sorted_hits(ho), sorted_hits(hg)), f"shape {sid}: hits incl. ungapped scores\n" + hit_diff(ho, hg, r_lim)
        assert np.array_equal(lo, lg), f"shape {sid}: SEED_MASK bits differ at {np.flatnonzero(lo != lg)[:20]}"
    if name == "fam2":
        sc = (res[1][0][0]["subject_score"] >> np.uint64(48)).astype(np.int64)
        assert (sc == 255).sum() > 100 and (sc > 255).sum() > 0


@pytest.mark.parametrize("name", ["c1", "fam2", "edge", "long", "rep"])
def test_blastp_default_sensitivity_matches_reference_golden(product_lib, name):
    from diamond_b200 import api
    w, q_raw, q_lim, r_raw, r_lim = workload_blocks(name)
    g = api.Context(product_lib, threads=8, comp_based_stats=1, masking=1, motif_masking=1, sensitivity=1)
    m, _, st = g.blastp(q_raw, q_lim, r_raw, r_lim)
    g.close()
    assert api.fmt6(m) == open(os.path.join(GOLDEN, f"{name}.s1.tsv")).read()
    cn = json.load(open(os.path.join(GOLDEN, f"{name}.s1.counters.json")))
    for k in ("seeds_hit", "seed_hits", "tentative_matches1", "tentative_matches2", "tentative_matches3"):
        assert st["seed"][k] == cn[k], k
    assert st["targets"] == cn["targets"] and st["dp_problems_round2"] == cn["targets_round2"]


@pytest.mark.parametrize("name", ["c1", "rep"])
def test_gapped_filter_flags_match_oracle(oracle_lib, product_lib, name):
    """dmnd_hits_gapped_filter (--sensitive): per-hit pass flags of the 64/128-diagonal scans, device vs oracle."""
    from diamond_b200 import api
    w, q_raw, q_lim, r_raw, r_lim = workload_blocks(name)
    res = []
    for lib in (oracle_lib, product_lib):
        c = api.Context(lib, threads=8, sensitivity=3)
        qb, rb = c.upload(q_raw, q_lim), c.upload(r_raw, r_lim)
        c.compute_bias(qb, 1)
        hits, cn, gf = c.search_shape(qb, rb, 0, gapped_filter=True)
        order = np.argsort(hits, order=["query", "subject_score", "seed_offset"])
        res.append((hits[order], gf[order]))
        c.free_block(qb); c.free_block(rb); c.close()
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert 0 < res[1][1].sum() < len(res[1][1])


@pytest.mark.parametrize("sens,level", [(2, "s2"), (3, "s3")])
@pytest.mark.parametrize("name", ["c1", "edge", "rep"])
def test_blastp_mid_sensitive_and_sensitive_match_reference_golden(product_lib, name, sens, level):
    from diamond_b200 import api
    w, q_raw, q_lim, r_raw, r_lim = workload_blocks(name)
    g = api.Context(product_lib, threads=8, comp_based_stats=1, masking=1, motif_masking=1, sensitivity=sens)
    m, _, st = g.blastp(q_raw, q_lim, r_raw, r_lim)
    g.close()
    assert api.fmt6(m) == open(os.path.join(GOLDEN, f"{name}.{level}.tsv")).read()
    cn = json.load(open(os.path.join(GOLDEN, f"{name}.{level}.counters.json")))
    for k in ("seeds_hit", "seed_hits", "tentative_matches1", "tentative_matches2", "tentative_matches3"):
        assert st["seed"][k] == cn[k], k
    assert st["targets"] == cn["targets"] and st["dp_problems_round2"] == cn["targets_round2"]


def test_cli_without_sensitivity_flag(product_lib, tmp_path):
    from diamond_b200 import synth
    w, *_ = workload_blocks("c1")
    q, d, o = (str(tmp_path / x) for x in ("q.faa", "d.faa", "o.tsv"))
    synth.write_fasta(q, w["q_letters"], w["q_off"], "q")
    synth.write_fasta(d, w["db_letters"], w["db_off"], "d")
    cli = os.path.join(ROOT, "diamond_b200", "bin", "dmnd-b200")
    r = subprocess.run([cli, "blastp", "-q", q, "-d", d, "-o", o, "-p", "8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(o).read() == open(os.path.join(GOLDEN, "c1.s1.tsv")).read()


@pytest.mark.parametrize("sens", [0, 1, 3])
def test_sliced_search_equals_unsliced(product_lib, sens, monkeypatch):
    """dmnd_search_shape cuts a query range into slices of DMND_SEED_SLICE letters (bounded entry / pair lists for the low-weight shapes
    of the sensitive modes); the concatenated hit lists must be the unsliced search's, hit for hit, and stay grouped by query."""
    from diamond_b200 import api
    w, q_raw, q_lim, r_raw, r_lim = workload_blocks("rep")
    res = []
    for slice_letters in (None, "3000"):
        if slice_letters:
            monkeypatch.setenv("DMND_SEED_SLICE", slice_letters)
        else:
            monkeypatch.setenv("DMND_SEED_SLICE", "1000000000")
        c = api.Context(product_lib, threads=8, sensitivity=sens)
        qb, rb = c.upload(q_raw, q_lim), c.upload(r_raw, r_lim)
        c.mask_block(qb, 5, 0, len(q_lim) - 1); c.mask_block(rb, 5, 0, len(r_lim) - 1)
        out = []
        for sid in range(min(c.params.n_shapes, 3)):
            hits, cn = c.search_shape(qb, rb, sid)
            assert np.all(np.diff(hits["query"].astype(np.int64)) >= 0), "hits not grouped by ascending query"
            out.append((sorted_hits(hits), {k: cn[k] for k in ("seed_hits", "tentative_matches1", "tentative_matches2", "tentative_matches3")}, c.download_letters(qb, q_raw.size)))
        res.append(out)
        c.free_block(qb); c.free_block(rb); c.close()
    for (h0, c0, l0), (h1, c1, l1) in zip(*res):
        assert c0 == c1 and np.array_equal(h0, h1) and np.array_equal(l0, l1)
    assert len(res[0][0][0]) > 0
