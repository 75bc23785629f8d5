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


# ---- frameshift alignment mode (blastx -F 15): legacy pipeline + fs_swipe_kernel / fs_walk_kernel on the device
def test_frameshift_cli_matches_reference_golden(product_lib, tmp_path):
    from diamond_b200 import synth
    w = _bx()
    q, d, o = (str(tmp_path / x) for x in ("q.fna", "d.faa", "o.tsv"))
    synth.write_dna_fasta(q, w["dna"])
    synth.write_fasta(d, w["db_letters"], w["db_off"], "d")
    cli = os.path.join(ROOT, "diamond_b200", "bin", "dmnd-b200")
    fields = "qseqid sseqid pident length mismatch gapopen qstart qend sstart send evalue bitscore cigar btop qseq_gapped sseq_gapped score qlen slen qframe".split()
    r = subprocess.run([cli, "blastx", "--fast", "-F", "15", "-q", q, "-d", d, "-f", "6"] + fields + ["-o", o, "-p", "8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(o).read() == open(os.path.join(GOLDEN, "bx.xf.tsv")).read()
    r = subprocess.run([cli, "blastx", "--fast", "-F", "15", "-q", q, "-d", d, "-f", "0", "-o", o, "-p", "8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(o).read() == open(os.path.join(GOLDEN, "bx.xf0.txt")).read()
    r = subprocess.run([cli, "blastx", "--sensitive", "-F", "15", "-q", q, "-d", d, "-o", o, "-p", "8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(o).read() == open(os.path.join(GOLDEN, "bx.xf3.tsv")).read()
    r = subprocess.run([cli, "blastx", "--long-reads", "-q", q, "-d", d, "-o", o, "-p", "8"], capture_output=True, text=True)  # range culling: the score-only round's column feeds the read ranges
    assert r.returncode == 0, r.stderr
    assert open(o).read() == open(os.path.join(GOLDEN, "bx.xl.tsv")).read()


@pytest.mark.parametrize("seed,maxdna,maxband", [(1, 600, 130), (2, 150, 40), (3, 4000, 900)])
def test_3frame_swipe_kernels_match_oracle(product_lib, oracle_lib, seed, maxdna, maxband):
    """K layer: dmnd_banded_3frame_swipe of the product library against the oracle on random problems (targets stitched from pieces of
    the three frames, corner bands, every register tile incl. the spilling R = 32 one), score-only and traceback."""
    import numpy as np
    from diamond_b200 import api
    rng = np.random.default_rng(seed)
    n = 160 if maxdna < 1000 else 40
    qs, ts, probs = [], [], []
    for it in range(n):
        L = int(rng.integers(3, maxdna))
        fr = [rng.integers(0, 20, (L - f) // 3).astype(np.int8) for f in range(3)]
        for v in fr:
            v[rng.random(len(v)) < 0.03] = 24
        t, f, i = [], int(rng.integers(0, 3)), int(rng.integers(0, max(1, len(fr[0]) // 3)))
        pre = int(rng.integers(0, 40))
        t += rng.integers(0, 20, pre).tolist()
        while i < len(fr[f]) and len(t) < 2500:
            u = rng.integers(0, 1000)
            if u < 25:
                f = (f + 1 + int(rng.integers(0, 2))) % 3
            elif u < 40:
                i += int(rng.integers(1, 4))
            elif u < 55:
                t += rng.integers(0, 20, int(rng.integers(1, 4))).tolist()
            else:
                t.append(int(rng.integers(0, 20)) if u < 300 else int(fr[f][i]))
                i += 1
        t += rng.integers(0, 20, int(rng.integers(0, 40))).tolist()
        t = t or [0]
        qs += fr
        ts.append(np.array(t, dtype=np.int8))
        qlen, tlen = len(fr[0]), len(t)
        lo, hi = -(tlen - 1), max(qlen, -(tlen - 1) + 1)
        wdt, kind, c = int(rng.integers(1, maxband)), int(rng.integers(0, 10)), -pre + int(rng.integers(-20, 21))
        if kind < 7:
            d0 = max(lo, c - wdt // 2); d1 = min(hi, d0 + wdt)
        elif kind == 7:
            d0 = lo; d1 = min(hi, lo + wdt)
        elif kind == 8:
            d1 = hi; d0 = max(lo, hi - wdt)
        else:
            d0 = lo + int(rng.integers(0, hi - lo)); d1 = d0 + 1
        d1 = max(d1, d0 + 1)
        probs.append((3 * it, it, d0, d1))

    def flat(seqs):
        off = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum([len(s) for s in seqs], out=off[1:])
        return (np.concatenate(seqs).astype(np.int8) if seqs else np.zeros(0, np.int8)), off
    q_raw, q_lim = api.block_image(*flat(qs))
    r_raw, r_lim = api.block_image(*flat(ts))
    pr = np.array(probs, dtype=api.PROBLEM_DTYPE)
    cap = int(sum(2 * len(ts[p[1]]) + len(qs[p[0]]) + 8 for p in probs))
    out = {}
    for name, lib in (("gpu", product_lib), ("oracle", oracle_lib)):
        g = api.Context(lib=lib, query_contexts=6, frame_shift=15)
        qb, rb = g.upload(q_raw, q_lim), g.upload(r_raw, r_lim)
        so, _ = g.banded_3frame_swipe(qb, rb, pr, 15, False)
        tb, tr = g.banded_3frame_swipe(qb, rb, pr, 15, True, cap)
        out[name] = (so["score"].copy(), tb.copy(), [bytes(tr[int(x["transcript_off"]):int(x["transcript_off"]) + int(x["transcript_len"])]) for x in tb], so["t_end"].copy())
        g.close()
    assert np.array_equal(out["gpu"][0], out["oracle"][0]) and np.array_equal(out["gpu"][3], out["oracle"][3])  # score-only: score and t_end (first column of the score)
    a, b = out["gpu"][1], out["oracle"][1]
    for k in ("score", "q_begin", "q_end", "frame_begin", "frame_end", "t_begin", "t_end", "identities", "mismatches", "gap_openings", "length", "gaps", "positives", "transcript_len", "status"):
        assert np.array_equal(a[k], b[k]), k
    assert out["gpu"][2] == out["oracle"][2]
    assert (b["score"] > 0).sum() > n // 2 and any(0x41 in t or 0x42 in t for t in out["oracle"][2])
