"""blastx on the device: the same kernels as blastp --fast over a query block of six translated contexts per
read; what changes is host logic (frames per target, frame-aware culling, nucleotide coordinates).  Against the reference's
goldens (tests/golden/bx.*); the CPU suite (tests/test_blastx.py) runs the identical host code over the oracle's K layer."""
import json, os, subprocess
import pytest
from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu


def _bx():
    from diamond_b200 import synth
    f, kw = synth.BX_WORKLOADS["bx"]
    return f(**kw)


@pytest.mark.parametrize("lanes", ["1", "2"])
def test_blastx_library_matches_reference_golden(product_lib, lanes, monkeypatch):
    from diamond_b200 import api
    monkeypatch.setenv("DMND_LANES", lanes)
    w = _bx()
    ql, qo = api.translate_reads(w["dna"])
    q_raw, q_lim = api.block_image(ql, qo)
    r_raw, r_lim = api.block_image(w["db_letters"], w["db_off"])
    g = api.Context(lib=product_lib, masking=1, motif_masking=1, query_contexts=6)
    m, _, st = g.blastp(q_raw, q_lim, r_raw, r_lim)
    g.close()
    assert api.fmt6_translated(m, [len(r) for r in w["dna"]]) == open(os.path.join(GOLDEN, "bx.x0.tsv")).read()
    gold = json.load(open(os.path.join(GOLDEN, "bx.x0.counters.json")))
    assert st["seed"]["seed_hits"] == gold["seed_hits"] and st["seed"]["tentative_matches3"] == gold["tentative_matches3"] and st["targets"] == gold["targets"]


def test_blastx_cli_transcript_fields(product_lib, tmp_path):
    from diamond_b200 import synth
    w = _bx()
    q, d, o = (str(tmp_path / x) for x in ("q.fna", "d.faa", "o.tsv"))
    synth.write_dna_fasta(q, w["dna"])
    synth.write_fasta(d, w["db_letters"], w["db_off"], "d")
    cli = os.path.join(ROOT, "diamond_b200", "bin", "dmnd-b200")
    fields = "qseqid sseqid pident length mismatch gapopen qstart qend sstart send evalue bitscore cigar btop qseq_gapped sseq_gapped score qlen slen".split()
    r = subprocess.run([cli, "blastx", "--fast", "-q", q, "-d", d, "-f", "6"] + fields + ["-o", o, "-p", "8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(o).read() == open(os.path.join(GOLDEN, "bx.xt.tsv")).read()
