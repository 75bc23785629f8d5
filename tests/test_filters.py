"""Report filters --id / --query-cover / --subject-cover (dmnd_search_opts.min_id, query_cover, subject_cover).

They are not a post-filter: with any of them set the reference's extension only sorts the targets after round 1, runs round 2 in
steps of max_target_seqs targets and applies Match::apply_filters after every step, until enough targets passed
(align/extend.cpp:94-96,288,331-336; align/gapped_final.cpp:107-158; align/culling.cpp:144-184); equal query and subject
covers >= 50 in a protein search also set min_length_ratio: length-sorted blocks and the mutual-coverage seed stage
(run/config.cpp:156-159, run/double_indexed.cpp:112-115,727-731, search/hamming/kernel_mutual_cov.h:28-52); frameshift mode filters in
QueryMapper::generate_output (align/legacy/query_mapper.cpp:242,338-349).  Goldens: the unmodified reference (tests/golden/make_golden.py,
levels I1, M1, XI, XFI).  CPU: host pipeline over the oracle kernels; tests/test_zzz_blastx_gpu.py runs the same files through the device."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, workload_blocks

GOLDEN = os.path.join(ROOT, "tests", "golden")
CLI = os.path.join(ROOT, "oracle", "_build", "dmnd-oracle-cli")

PROTEIN = [("c1", "i1", ["--fast", "--id", "60", "--query-cover", "50"]),
           ("edge", "i1", ["--fast", "--id", "60", "--query-cover", "50"]),
           ("fam2", "i1", ["--fast", "--id", "60", "--query-cover", "50"]),
           ("edge", "m1", ["--query-cover", "70", "--subject-cover", "70"]),
           ("fam2", "m1", ["--query-cover", "70", "--subject-cover", "70"])]
TRANSLATED = [("xi", ["--fast", "--id", "50", "--query-cover", "60"]), ("xfi", ["--fast", "-F", "15", "--id", "50", "--subject-cover", "20"])]


def run_protein(cli, name, flags, tmp_path):
    from diamond_b200 import synth
    w, *_ = workload_blocks(name)
    q, d, o = (str(tmp_path / x) for x in ("q.faa", "d.faa", "o.tsv"))
    synth.write_fasta(q, w["q_letters"], w["q_off"], "q")
    synth.write_fasta(d, w["db_letters"], w["db_off"], "d")
    r = subprocess.run([cli, "blastp"] + flags + ["-q", q, "-d", d, "-o", o, "-p", "8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return open(o).read()


def run_translated(cli, flags, tmp_path):
    from diamond_b200 import synth
    f, kw = synth.BX_WORKLOADS["bx"]
    w = f(**kw)
    q, d, o = (str(tmp_path / x) for x in ("q.fna", "d.faa", "o.tsv"))
    synth.write_dna_fasta(q, w["dna"])
    synth.write_fasta(d, w["db_letters"], w["db_off"], "d")
    r = subprocess.run([cli, "blastx"] + flags + ["-q", q, "-d", d, "-o", o, "-p", "8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return open(o).read()


@pytest.mark.parametrize("name,lvl,flags", PROTEIN)
def test_filtered_protein_search_matches_reference_golden(oracle_lib, name, lvl, flags, tmp_path):
    gold = open(os.path.join(GOLDEN, f"{name}.{lvl}.tsv")).read()
    assert run_protein(CLI, name, flags, tmp_path) == gold
    if lvl == "i1":  # every reported alignment does satisfy the filters (the reference prints pident with one decimal, rounded)
        for l in gold.splitlines():
            assert float(l.split("\t")[2]) >= 59.95
    else:  # M1: queries come out longest first (length-sorted query block)
        w, *_ = workload_blocks(name)
        lens = np.diff(w["q_off"])
        order = [int(l.split("\t")[0][1:]) for l in gold.splitlines()]
        seen = [q for i, q in enumerate(order) if i == 0 or order[i - 1] != q]
        assert all(lens[a] >= lens[b] for a, b in zip(seen, seen[1:]))


@pytest.mark.parametrize("lvl,flags", TRANSLATED)
def test_filtered_translated_search_matches_reference_golden(oracle_lib, lvl, flags, tmp_path):
    assert run_translated(CLI, flags, tmp_path) == open(os.path.join(GOLDEN, f"bx.{lvl}.tsv")).read()


def test_filters_are_not_a_post_filter(oracle_lib):
    """The filtered schedule finds targets a post-filter of the unfiltered best-25 list would miss (fam2: 400 members per family, so the
    25 best targets of a query are all near-identical ones; asking for --id <= some bound is not expressible, but a high query cover
    keeps extending past them) -- and the library call gives the CLI golden."""
    from diamond_b200 import api
    w, q_raw, q_lim, r_raw, r_lim = workload_blocks("fam2")
    ctx = api.Context(oracle_lib, masking=1, motif_masking=1, min_id=60.0, query_cover=50.0)
    m, _, stats = ctx.blastp(q_raw, q_lim, r_raw, r_lim)
    ctx.close()
    assert api.fmt6(m) == open(os.path.join(GOLDEN, "fam2.i1.tsv")).read()
    plain = {tuple(l.split("\t")[:2]) for l in open(os.path.join(GOLDEN, "fam2.l2.tsv")).read().splitlines()}
    filt = {tuple(l.split("\t")[:2]) for l in open(os.path.join(GOLDEN, "fam2.i1.tsv")).read().splitlines()}
    assert filt - plain, "the filtered run reports (query, target) pairs the unfiltered best-25 lists do not hold"


def test_filter_options_are_checked(oracle_lib):
    from diamond_b200 import api
    w, q_raw, q_lim, r_raw, r_lim = workload_blocks("edge")
    ctx = api.Context(oracle_lib, min_id=120.0)
    with pytest.raises(api.DmndError, match="percentages"):
        ctx.blastp(q_raw, q_lim, r_raw, r_lim)
    ctx.close()


def check_no_self_hits(cli, tmp_path):
    from diamond_b200 import synth
    w, *_ = workload_blocks("fam2")
    q, d, o = (str(tmp_path / x) for x in ("qs.faa", "d.faa", "o.tsv"))
    synth.write_fasta(d, w["db_letters"], w["db_off"], "d")
    synth.write_fasta(q, w["db_letters"][:w["db_off"][150]], w["db_off"][:151], "d")  # the first 150 database sequences under their database titles
    r = subprocess.run([cli, "blastp", "--fast", "--no-self-hits", "-k", "3", "-q", q, "-d", d, "-o", o, "-p", "8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    gold = open(os.path.join(GOLDEN, "fam2.n1.tsv")).read()
    assert open(o).read() == gold
    assert not any(l.split("\t")[0] == l.split("\t")[1] for l in gold.splitlines())


def test_no_self_hits(oracle_lib, tmp_path):
    check_no_self_hits(CLI, tmp_path)


def test_library_self_targets_and_min_score(oracle_lib):
    """dmnd_search_opts.self_targets / min_bit_score through the Python mirror: the N1 golden from the library call, and a bit-score bound that
    keeps exactly the lines of the unfiltered golden at or above it (--min-score replaces the e-value bound; nothing else moves at -k 25 here)."""
    from diamond_b200 import api
    w, q_raw, q_lim, r_raw, r_lim = workload_blocks("fam2")
    n = 150
    qs_raw, qs_lim = api.block_image(w["db_letters"][:w["db_off"][n]], w["db_off"][:n + 1])
    ctx = api.Context(oracle_lib, masking=1, motif_masking=1, max_target_seqs=3, self_targets=np.arange(n, dtype=np.uint32))
    m, _, _ = ctx.blastp(qs_raw, qs_lim, r_raw, r_lim)
    ctx.close()
    assert api.fmt6(m, q_prefix="d") == open(os.path.join(GOLDEN, "fam2.n1.tsv")).read()
    w, q_raw, q_lim, r_raw, r_lim = workload_blocks("edge")
    ctx = api.Context(oracle_lib, masking=1, motif_masking=1, min_bit_score=100.0)
    m, _, _ = ctx.blastp(q_raw, q_lim, r_raw, r_lim)
    ctx.close()
    gold = [l for l in open(os.path.join(GOLDEN, "edge.l2.tsv")).read().splitlines() if float(l.split("\t")[11]) >= 100.0]
    assert api.fmt6(m).splitlines() == gold
