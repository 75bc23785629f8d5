"""BLAST XML (-f 5, output/xml_format.cpp:31-176) and the unaligned records of a blocked run (output/join_blocks.cpp:302-308,365-372), against
files written by the unmodified reference (tests/golden/make_golden.py, levels F5, XX, B1)."""
import os
import subprocess

import pytest

from conftest import ROOT, workload_blocks
from test_filters import run_protein, run_translated

GOLDEN = os.path.join(ROOT, "tests", "golden")
CLI = os.path.join(ROOT, "oracle", "_build", "dmnd-oracle-cli")


def check_xml_protein(cli, tmp_path):
    out = run_protein(cli, "edge", ["--fast", "-f", "5"], tmp_path)
    gold = open(os.path.join(GOLDEN, "edge.f5.xml")).read()
    # the header quotes the -d argument as given: the golden was written for "<tmp>/d.faa"
    strip = lambda s: "\n".join(l for l in s.split("\n") if "<BlastOutput_db>" not in l)
    assert strip(out) == strip(gold)
    assert not gold.endswith("\n")  # the footer has no final newline


def check_xml_translated(cli, tmp_path):
    out = run_translated(cli, ["--fast", "-f", "5", "-k", "1", "-e", "1e-20"], tmp_path)
    gold = open(os.path.join(GOLDEN, "bx.xx.xml")).read()
    strip = lambda s: "\n".join(l for l in s.split("\n") if "<BlastOutput_db>" not in l)
    assert strip(out) == strip(gold)
    assert "<Hsp_query-frame>-2</Hsp_query-frame>" in gold


def check_blocked_unaligned(cli, tmp_path):
    out = run_protein(cli, "rep", ["--fast", "-b", "0.00003", "--unal", "1", "-k", "3"], tmp_path)
    gold = open(os.path.join(GOLDEN, "rep.b1.tsv")).read()
    assert out == gold
    assert sum(1 for l in gold.splitlines() if l.split("\t")[1] == "*") > 50  # every query without an alignment, seed hits or not


def test_xml_protein(oracle_lib, tmp_path):
    check_xml_protein(CLI, tmp_path)


def test_xml_translated(oracle_lib, tmp_path):
    check_xml_translated(CLI, tmp_path)


def test_blocked_run_reports_every_unaligned_query(oracle_lib, tmp_path):
    check_blocked_unaligned(CLI, tmp_path)


def test_fastq_titles_keep_their_newline(oracle_lib, tmp_path):
    """The reference's FASTQ reader appends a newline to the title (data/fasta/parser.h:247); qtitle shows it."""
    from diamond_b200 import synth
    f, kw = synth.BX_WORKLOADS["bx"]
    w = f(**kw)
    q, d, o = (str(tmp_path / x) for x in ("q.fastq", "d.faa", "o.tsv"))
    with open(q, "w") as fh:
        for i, s in enumerate(w["dna"][:40]):
            fh.write(f"@r{i} read {i}\n{s}\n+\n{'I' * len(s)}\n")
    synth.write_fasta(d, w["db_letters"], w["db_off"], "d")
    r = subprocess.run([CLI, "blastx", "--fast", "-q", q, "-d", d, "-o", o, "-f", "6", "qseqid", "qtitle", "sseqid", "-k", "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = open(o).read().split("\n")
    assert lines[0].startswith("r") and lines[0].split("\t")[1].startswith("r") and lines[1].startswith("\td")  # "id<TAB>title<NL><TAB>target"


FRAMESHIFT_FORMATS = [("xfp", "paf", ["--fast", "-F", "15", "-f", "paf"]), ("xfs", "sam", ["--fast", "-F", "15", "-f", "sam", "-k", "1", "-e", "1e-20"]), ("xlb", "tsv", ["--long-reads", "-b", "0.0002"])]


def check_frameshift_format(cli, lvl, ext, flags, tmp_path):
    """Frameshift mode in the PAF and SAM formats (read coordinates of an alignment whose end lies in another frame; frameshift operations in the
    CIGAR) and --long-reads over several reference blocks (the join's culling per query range): the reference's files."""
    out = run_translated(cli, flags, tmp_path)
    gold = open(os.path.join(GOLDEN, f"bx.{lvl}.{ext}")).read()
    drop_pg = lambda s: [l for l in s.split("\n") if not l.startswith("@PG")]  # SAM quotes the program's own command line there
    assert drop_pg(out) == drop_pg(gold)
    if lvl == "xfs":
        assert any("\\" in l.split("\t")[5] or "/" in l.split("\t")[5] for l in gold.splitlines() if not l.startswith("@"))


@pytest.mark.parametrize("lvl,ext,flags", FRAMESHIFT_FORMATS)
def test_frameshift_formats(oracle_lib, lvl, ext, flags, tmp_path):
    check_frameshift_format(CLI, lvl, ext, flags, tmp_path)


def check_daa(cli, tmp_path):
    """-f 100: the DIAMOND alignment archive, byte for byte (one thread: the dictionary order of the reference depends on its threads)."""
    import struct
    from diamond_b200 import synth
    w, *_ = workload_blocks("rep")
    q, d, o = (str(tmp_path / x) for x in ("q.faa", "d.faa", "o"))
    synth.write_fasta(q, w["q_letters"], w["q_off"], "q")
    synth.write_fasta(d, w["db_letters"], w["db_off"], "d")
    r = subprocess.run([cli, "blastp", "--fast", "-k", "2", "-f", "100", "-q", q, "-d", d, "-o", o, "-p", "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = open(o + ".daa", "rb").read()  # the extension is appended, as the reference does
    gold = open(os.path.join(GOLDEN, "rep.d1.daa"), "rb").read()
    assert got == gold
    magic, version = struct.unpack("<QQ", gold[:16])
    assert magic == 0x3c0e53476d3ee36b and version == 1


def test_daa(oracle_lib, tmp_path):
    check_daa(CLI, tmp_path)


def check_view(cli, tmp_path):
    """`view`: the reference's archive (golden D1) read back and printed -- for a protein search exactly what the search itself prints, in the
    tabular format with the transcript fields and in the pairwise format (legacy/daa/daa_record.cpp:30-86, view.cpp)."""
    fields = "qseqid sseqid pident length mismatch gapopen qstart qend sstart send evalue bitscore cigar btop qseq_gapped sseq_gapped positive score".split()
    for fmt in (["-f", "6"] + fields, ["-f", "0"]):
        direct = run_protein(cli, "rep", ["--fast", "-k", "2"] + fmt, tmp_path)
        o = str(tmp_path / "view.out")
        r = subprocess.run([cli, "view", "-a", os.path.join(GOLDEN, "rep.d1.daa"), "-o", o] + fmt, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        viewed = open(o).read()
        if fmt[1] == "0":  # (an archive holds aligned queries only: the search's "No hits found" records are not in it)
            assert "***** No hits found *****" not in viewed and viewed.count("Query= ") <= direct.count("Query= ")
            direct = "".join(b for b in __import__("re").split(r"(?=Query= )", direct) if "No hits found" not in b)
        assert viewed == direct


def test_view(oracle_lib, tmp_path):
    check_view(CLI, tmp_path)


def test_stdio_compress_dbinfo(oracle_lib, tmp_path):
    """Queries from standard input, output to standard output, --compress 1 (gzip, ".gz" appended), dbinfo."""
    import gzip
    from diamond_b200 import synth
    w, *_ = workload_blocks("edge")
    q, d, o = (str(tmp_path / x) for x in ("q.faa", "d.faa", "o.tsv"))
    synth.write_fasta(q, w["q_letters"], w["q_off"], "q")
    synth.write_fasta(d, w["db_letters"], w["db_off"], "d")
    gold = open(os.path.join(GOLDEN, "edge.l2.tsv")).read()
    r = subprocess.run([CLI, "blastp", "--fast", "-d", d, "-p", "8"], input=open(q, "rb").read(), capture_output=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.decode() == gold
    r = subprocess.run([CLI, "blastp", "--fast", "-q", q, "-d", d, "-o", o, "-p", "8", "--compress", "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert not os.path.exists(o) and gzip.open(o + ".gz", "rt").read() == gold
    db = str(tmp_path / "db")
    assert subprocess.run([CLI, "makedb", "--in", d, "-d", db], capture_output=True).returncode == 0
    r = subprocess.run([CLI, "dbinfo", "-d", db], capture_output=True, text=True)
    n_letters = int(w["db_off"][-1])
    assert r.returncode == 0 and f"Sequences  {len(w['db_off']) - 1}\n" in r.stdout and f"Letters  {n_letters}\n" in r.stdout and "Database format version  3" in r.stdout


def test_json_flat(oracle_lib, tmp_path):
    """-f 104 / json-flat: the tabular fields as an array of flat objects.  Same values as the tabular golden, and -- unlike the reference's
    file, which loses the comma between two records wherever one of its query bins ends (a new OutputWriter starts "first", output/output.h:97-110)
    -- always valid JSON."""
    import json
    out = run_protein(CLI, "edge", ["--fast", "-f", "104"], tmp_path)
    recs = json.loads(out)
    gold = [l.split("\t") for l in open(os.path.join(GOLDEN, "edge.l2.tsv")).read().splitlines()]
    assert len(recs) == len(gold) and list(recs[0].keys()) == "qseqid sseqid pident length mismatch gapopen qstart qend sstart send evalue bitscore".split()
    values = [[v.strip('"') for v in __import__("re").findall(r'^\t"[a-z_]+":(.*?),?$', blk, flags=8)] for blk in out.split("\n\t{\n")[1:]]
    assert values == gold
