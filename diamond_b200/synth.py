"""Seeded synthetic protein workloads for the blastp hot path.

DB: `n_db` proteins, length clip(Gamma(k=4, theta=75), 50, 2000), letters iid from the Robinson-Robinson
background over ARNDCQEGHILKMFPSTWYV.  Queries: planted homologs -- a window (<= `qlen`) of a uniformly chosen DB
sequence, each letter substituted with per-query probability r ~ U(0.1, 0.6) by a background draw, deleted with
p = 0.01, and followed by an inserted background letter with p = 0.01.  PRNG = numpy default_rng(seed).

Everything is produced as *encoded* letters (reference alphabet order ARNDCQEGHILKMFPSTWYV -> 0..19,
basic/value.h:53) in one flat int8 array + int64 offsets, so the bench can feed the library directly; FASTA
text is only written when a file is requested (for the reference CLI).
"""
from __future__ import annotations
import numpy as np

ALPHABET = "ARNDCQEGHILKMFPSTWYV"
# Robinson & Robinson 1991 amino-acid background frequencies, in ALPHABET order.
RR_FREQ = np.array([0.07805, 0.05129, 0.04487, 0.05364, 0.01925, 0.04264, 0.06295, 0.07377, 0.02199, 0.05142,
                    0.09019, 0.05744, 0.02243, 0.03856, 0.05203, 0.07120, 0.05841, 0.01330, 0.03216, 0.06441])
RR_FREQ = RR_FREQ / RR_FREQ.sum()
# 16-bit inverse-CDF table: letter = _LUT[u16] draws from RR_FREQ quantised to 1/65536 (fast enough for 3e8 letters)
_LUT = np.searchsorted(np.cumsum(RR_FREQ), (np.arange(65536) + 0.5) / 65536.0).clip(0, 19).astype(np.int8)


def draw_letters(rng: np.random.Generator, n: int) -> np.ndarray:
    return _LUT[rng.integers(0, 65536, size=n, dtype=np.uint16)]


def make_db(n_db: int, rng: np.random.Generator):
    lens = np.clip(rng.gamma(4.0, 75.0, n_db), 50, 2000).astype(np.int64)
    off = np.zeros(n_db + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    letters = draw_letters(rng, int(off[-1]))
    return letters, off


def make_queries(n_q: int, db_letters: np.ndarray, db_off: np.ndarray, rng: np.random.Generator, qlen: int = 300,
                 rate_range=(0.1, 0.6)):
    n_db = len(db_off) - 1
    src = rng.integers(0, n_db, n_q)
    slen = db_off[src + 1] - db_off[src]
    wlen = np.minimum(slen, qlen)
    start = (rng.random(n_q) * (slen - wlen + 1)).astype(np.int64)
    rate = rng.uniform(rate_range[0], rate_range[1], n_q)
    woff = np.zeros(n_q + 1, dtype=np.int64)
    np.cumsum(wlen, out=woff[1:])
    total = int(woff[-1])
    qid = np.repeat(np.arange(n_q), wlen)
    pos_in_w = np.arange(total, dtype=np.int64) - woff[qid]
    base = db_letters[db_off[src][qid] + start[qid] + pos_in_w]
    sub = rng.random(total, dtype=np.float32) < rate[qid].astype(np.float32)
    bg = draw_letters(rng, total)
    base = np.where(sub, bg, base)
    keep = rng.random(total, dtype=np.float32) >= np.float32(0.01)
    ins = rng.random(total, dtype=np.float32) < np.float32(0.01)
    ins_letter = draw_letters(rng, total)
    # emit: kept letter (0/1) followed by inserted letter (0/1)
    cnt = keep.astype(np.int64) + ins.astype(np.int64)
    out_off_flat = np.zeros(total + 1, dtype=np.int64)
    np.cumsum(cnt, out=out_off_flat[1:])
    out = np.empty(int(out_off_flat[-1]), dtype=np.int8)
    out[out_off_flat[:-1][keep]] = base[keep]
    out[(out_off_flat[:-1] + keep)[ins]] = ins_letter[ins]
    q_off = out_off_flat[woff]
    # guarantee non-empty queries
    assert np.all(np.diff(q_off) > 0)
    return out, q_off.astype(np.int64), src


def workload(n_q: int, n_db: int, seed: int, qlen: int = 300, chunk: int = 8192, threads: int | None = None,
             q_stream: int = 0):
    """DB from default_rng(seed); queries in fixed chunks of `chunk`, chunk k drawn from default_rng([seed, 1 + q_stream, k]) so the
    result does not depend on the number of worker threads (numpy releases the GIL in the heavy passes)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    rng = np.random.default_rng(seed)
    dbl, dbo = make_db(n_db, rng)
    starts = list(range(0, n_q, chunk))

    def job(k):
        n = min(chunk, n_q - starts[k])
        return make_queries(n, dbl, dbo, np.random.default_rng([seed, 1 + q_stream, k]), qlen)

    threads = threads or min(32, os.cpu_count() or 1)
    if len(starts) == 1 or threads == 1:
        parts = [job(k) for k in range(len(starts))]
    else:
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(job, range(len(starts))))
    ql = np.concatenate([p[0] for p in parts])
    lens = np.concatenate([np.diff(p[1]) for p in parts])
    qo = np.zeros(n_q + 1, dtype=np.int64)
    np.cumsum(lens, out=qo[1:])
    src = np.concatenate([p[2] for p in parts])
    return {"q_letters": ql, "q_off": qo, "db_letters": dbl, "db_off": dbo, "src": src}


def write_fasta(path: str, letters: np.ndarray, off: np.ndarray, prefix: str) -> None:
    lut = np.frombuffer(b"ARNDCQEGHILKMFPSTWYVBJZX*_", dtype=np.uint8)
    text = lut[letters]
    with open(path, "wb") as f:
        chunks = []
        for i in range(len(off) - 1):
            chunks.append(b">%s%d\n" % (prefix.encode(), i))
            chunks.append(text[off[i]:off[i + 1]].tobytes())
            chunks.append(b"\n")
            if len(chunks) >= 30000:
                f.write(b"".join(chunks)); chunks = []
        f.write(b"".join(chunks))


def edge_workload(seed: int, n_db: int = 2000, n_q: int = 300):
    """Short / ragged queries (25..150 letters, some shorter than the 16-letter seed span), masked letters (X) and stop
    codons (*) inside queries, queries without any homolog -- the edge cases the reference's own goldens cover with
    data.faa (short sequences) and nanopore reads (stops)."""
    rng = np.random.default_rng(seed)
    dbl, dbo = make_db(n_db, rng)
    qs = []
    for k in range(n_q):
        if k % 29 == 0:  # no homolog at all
            L = int(rng.integers(10, 150))
            qs.append(draw_letters(rng, L))
            continue
        s = int(rng.integers(0, n_db))
        slen = int(dbo[s + 1] - dbo[s])
        L = int(min(slen, rng.integers(12, 150)))
        st = int(rng.integers(0, slen - L + 1))
        q = dbl[dbo[s] + st: dbo[s] + st + L].copy()
        sub = rng.random(L) < rng.uniform(0.05, 0.4)
        q[sub] = draw_letters(rng, int(sub.sum()))
        if k % 10 == 3:
            q[rng.integers(0, L, size=min(3, L))] = 23  # X
        if k % 17 == 5:
            q[int(rng.integers(0, L))] = 24  # *
        qs.append(q)
    qo = np.zeros(n_q + 1, dtype=np.int64)
    np.cumsum([len(q) for q in qs], out=qo[1:])
    return {"q_letters": np.concatenate(qs).astype(np.int8), "q_off": qo, "db_letters": dbl, "db_off": dbo, "src": None}


def family_workload(n_fam: int, fam_size: int, n_q: int, seed: int, base_len: int = 350, member_div=(0.10, 0.35),
                    query_div=(0.1, 0.6)):
    """DB of `n_fam` families, each `fam_size` diverged copies (10-35 % substitutions, 1 % indels) of a random base
    protein; queries are further-mutated windows of family members.  Gives queries with hundreds of targets, which
    exercises the ranking-chunk loop (align/extend.cpp:259-336) and -k culling."""
    rng = np.random.default_rng(seed)
    seqs = []
    for f in range(n_fam):
        L = int(rng.integers(base_len // 2, base_len * 2))
        base = draw_letters(rng, L)
        for m in range(fam_size):
            r = rng.uniform(member_div[0], member_div[1])
            s = base.copy()
            sub = rng.random(L) < r
            s[sub] = draw_letters(rng, int(sub.sum()))
            keep = rng.random(L) >= 0.01
            s = s[keep]
            seqs.append(s)
    perm = rng.permutation(len(seqs))
    seqs = [seqs[i] for i in perm]
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    dbl = np.concatenate(seqs).astype(np.int8)
    ql, qo, src = make_queries(n_q, dbl, off, rng, 400, query_div)
    return {"q_letters": ql, "q_off": qo, "db_letters": dbl, "db_off": off, "src": src}


def long_workload(seed: int, n_db: int = 300, n_long: int = 5, n_short_q: int = 20):
    """A few very long proteins (7 000..12 000 letters) with diverged, indel-carrying query copies among ordinary
    sequences: their round-2 problems exceed max_swipe_dp (band x columns > 10^6), so the reference aligns them with
    the statistics passes instead of a traceback (dp/swipe/swipe_wrapper.cpp:89-96)."""
    rng = np.random.default_rng(seed)
    dbl, dbo = make_db(n_db, rng)
    seqs = [dbl[dbo[i]:dbo[i + 1]] for i in range(n_db)]
    qs = []
    for k in range(n_long):
        L = int(rng.integers(7000, 12000))
        base = draw_letters(rng, L)
        seqs.append(base)
        # query: substitutions + a handful of indels that widen the chain's diagonal range
        q = base.copy()
        sub = rng.random(L) < rng.uniform(0.45, 0.7)
        q[sub] = draw_letters(rng, int(sub.sum()))
        parts, pos = [], 0
        for cut in sorted(rng.integers(200, L - 200, size=int(rng.integers(3, 9)))):
            parts.append(q[pos:cut])
            if rng.random() < 0.5:
                parts.append(draw_letters(rng, int(rng.integers(3, 30))))  # insertion
                pos = cut
            else:
                pos = min(L, cut + int(rng.integers(3, 30)))  # deletion
        parts.append(q[pos:])
        qs.append(np.concatenate(parts))
        if k % 2 == 0:  # a long fragment as well (different length ratio, band clipped by the query end)
            st = int(rng.integers(0, L // 3))
            qs.append(qs[-1][st: st + int(rng.integers(3000, 6000))].copy())
    for k in range(n_short_q):
        s = int(rng.integers(0, n_db))
        q = seqs[s].copy()
        sub = rng.random(len(q)) < 0.2
        q[sub] = draw_letters(rng, int(sub.sum()))
        qs.append(q)
    order = rng.permutation(len(seqs))
    seqs = [seqs[i] for i in order]
    dbo2 = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(x) for x in seqs], out=dbo2[1:])
    qo = np.zeros(len(qs) + 1, dtype=np.int64)
    np.cumsum([len(x) for x in qs], out=qo[1:])
    return {"q_letters": np.concatenate(qs).astype(np.int8), "q_off": qo, "db_letters": np.concatenate(seqs).astype(np.int8), "db_off": dbo2, "src": None}


def motif_codes():
    """The reference's abundant-motif 8-mers (diamond_b200/csrc/host/motif_table.h) as letter arrays."""
    import os, re
    h = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "host", "motif_table.h")).read()
    out = []
    for c in re.findall(r"0x([0-9a-f]+)ull", h):
        v, m = int(c, 16), []
        for _ in range(8):
            m.append(v % 20); v //= 20
        out.append(np.array(m[::-1], dtype=np.int8))
    return out


def repeat_workload(seed: int, n_db: int = 1500, n_q: int = 400):
    """Sequences that the masking stages act on (masking/tantan.cpp, masking/masking.cpp:110-131): tandem repeats of
    period 1..40 with diverged copies, homopolymer / two-letter low-complexity runs, abundant-motif 8-mers (single, adjacent
    pairs that merge into one range, chains longer than max_motif_len, at the very start and end of a sequence, covering more
    than half of a short sequence), X letters next to repeats, sequences shorter than a motif.  Queries are mutated windows of
    those sequences, so seeds, x-drop extensions and alignments run across masked letters."""
    rng = np.random.default_rng(seed)
    motifs = motif_codes()
    seqs = []
    for k in range(n_db):
        L = int(np.clip(rng.gamma(4.0, 75.0), 30, 1500))
        s = draw_letters(rng, L)
        kind = k % 8
        if kind in (0, 1, 2):  # tandem repeat
            period = int(rng.integers(1, 41)) if kind != 2 else int(rng.integers(1, 4))
            copies = int(rng.integers(3, 30))
            unit = draw_letters(rng, period)
            rep = np.tile(unit, copies)
            mut = rng.random(len(rep)) < rng.uniform(0.0, 0.25)
            rep[mut] = draw_letters(rng, int(mut.sum()))
            at = int(rng.integers(0, max(1, L - 1)))
            s = np.concatenate([s[:at], rep, s[at:]])
        elif kind == 3:  # motifs
            mode = (k // 8) % 6
            pick = lambda: motifs[int(rng.integers(0, len(motifs)))]
            if mode == 0: ins = [pick()]
            elif mode == 1: ins = [pick(), pick()]                      # adjacent: merges into one 16-letter range
            elif mode == 2: ins = [pick() for _ in range(5)]            # 40 letters > max_motif_len: not masked
            elif mode == 3: ins = [pick(), draw_letters(rng, 3), pick()]
            else: ins = [pick()]
            piece = np.concatenate(ins)
            if mode == 4: s = np.concatenate([piece, s])               # at the very start (SEED_MASK range clipped at 0)
            elif mode == 5: s = np.concatenate([s, piece])             # at the very end
            else:
                at = int(rng.integers(0, L))
                s = np.concatenate([s[:at], piece, s[at:]])
        elif kind == 4 and (k // 8) % 5 == 0:  # short sequence, motifs cover more than half of it: nothing is masked
            s = np.concatenate([motifs[int(rng.integers(0, len(motifs)))], draw_letters(rng, 5), motifs[int(rng.integers(0, len(motifs)))], draw_letters(rng, 4)])
        elif kind == 4 and (k // 8) % 5 == 1:  # shorter than a motif
            s = draw_letters(rng, int(rng.integers(1, 8)))
        elif kind == 5:  # low complexity: two-letter run + X letters around it
            a, b = draw_letters(rng, 2)
            run = np.where(rng.random(int(rng.integers(10, 80))) < 0.7, a, b).astype(np.int8)
            at = int(rng.integers(0, L))
            s = np.concatenate([s[:at], run, s[at:]])
            s[rng.integers(0, len(s), size=3)] = 23
        seqs.append(s.astype(np.int8))
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(x) for x in seqs], out=off[1:])
    dbl = np.concatenate(seqs).astype(np.int8)
    qs = []
    for k in range(n_q):
        sid = int(rng.integers(0, n_db))
        src = seqs[sid]
        L = int(min(len(src), rng.integers(20, 400)))
        st = int(rng.integers(0, len(src) - L + 1))
        q = src[st:st + L].copy()
        sub = rng.random(L) < rng.uniform(0.0, 0.35)
        q[sub] = draw_letters(rng, int(sub.sum()))
        qs.append(q)
    qo = np.zeros(n_q + 1, dtype=np.int64)
    np.cumsum([len(q) for q in qs], out=qo[1:])
    return {"q_letters": np.concatenate(qs).astype(np.int8), "q_off": qo, "db_letters": dbl, "db_off": off, "src": None}


# ---- blastx: DNA reads over a protein database ------------------------------------------------------------------------------
_CODE = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"  # standard code, TCAG order
_CODONS = {}
for _i, _a in enumerate(_CODE):
    _CODONS.setdefault(_a, []).append("TCAG"[_i >> 4] + "TCAG"[(_i >> 2) & 3] + "TCAG"[_i & 3])


def _revcomp(s: str) -> str:
    return s[::-1].translate(str.maketrans("ACGTNR", "TGCANY"))


def reads_workload(seed: int, n_db: int = 1500, n_q: int = 500):
    """DNA queries for blastx: back-translated, mutated windows of database proteins on either strand with random flanks;
    plus frame-shifted reads (one target hit in two frames), chimeras of two proteins, reads with N / IUPAC codes, trinucleotide
    repeats (tantan masks a translated frame), reads without a homolog, reads shorter than a codon, and a few gene-length
    queries.  The database carries mutated paralogs of its first 200 proteins (41 of the first five) so that a read has several targets."""
    rng = np.random.default_rng(seed)
    dbl, dbo = make_db(n_db, rng)
    seqs = [dbl[dbo[i]:dbo[i + 1]] for i in range(n_db)]
    for i in list(range(200)) + [j % 5 for j in range(200)]:  # one paralog each, 40 more of the first five (culling at 25 targets)
        c = seqs[i].copy()
        sub = rng.random(len(c)) < rng.uniform(0.05, 0.3)
        c[sub] = draw_letters(rng, int(sub.sum()))
        seqs.append(c)
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(x) for x in seqs], out=off[1:])
    dbl = np.concatenate(seqs).astype(np.int8)

    def rand_dna(n):
        return "".join("ACGT"[x] for x in rng.integers(0, 4, n))

    def coding(max_aa):
        u = rng.random()
        sid = int(rng.integers(0, 5)) if u < 0.1 else int(rng.integers(0, 200)) if u < 0.6 else int(rng.integers(0, len(seqs)))
        p = seqs[sid]
        L = int(min(len(p), rng.integers(12, max_aa)))
        st = int(rng.integers(0, len(p) - L + 1))
        w = p[st:st + L].copy()
        sub = rng.random(L) < rng.uniform(0.03, 0.35)
        w[sub] = draw_letters(rng, int(sub.sum()))
        return "".join(_CODONS[ALPHABET[a]][int(rng.integers(0, len(_CODONS[ALPHABET[a]])))] for a in w)

    reads = []
    for k in range(n_q):
        if k % 29 == 0:
            r = rand_dna(int(rng.integers(30, 400)))
        elif k % 31 == 1:
            r = rand_dna(int(rng.integers(1, 6)))  # shorter than two codons: (nearly) empty frames
        else:
            r = coding(600 if k % 7 == 0 else 110)
            if k % 11 == 2 and len(r) > 60:  # frame shift in the middle
                m = len(r) // 2
                r = r[:m] + rand_dna(int(rng.integers(1, 3))) + r[m:]
            if k % 13 == 3:  # chimera, second part possibly on the other strand
                r2 = coding(110)
                r = r + rand_dna(int(rng.integers(0, 9))) + (_revcomp(r2) if rng.random() < 0.5 else r2)
            if k % 23 == 4:
                r = r + "GCA" * 30
            r = rand_dna(int(rng.integers(0, 40))) + r + rand_dna(int(rng.integers(0, 40)))
            if k % 10 == 3:
                r = list(r)
                for x in rng.integers(0, len(r), 3):
                    r[int(x)] = "N"
                r = "".join(r)
            if k % 19 == 4:
                x = int(rng.integers(0, len(r)))
                r = r[:x] + "R" + r[x + 1:]
            if rng.random() < 0.5:
                r = _revcomp(r)
        reads.append(r)
    return {"dna": reads, "db_letters": dbl, "db_off": off}


def write_dna_fasta(path: str, reads, prefix: str = "r") -> None:
    with open(path, "w") as f:
        for i, r in enumerate(reads):
            f.write(">%s%d\n%s\n" % (prefix, i, r))


def c3_workload(n_reads: int, n_db: int, seed: int, read_len: int = 150, q_stream: int = 0, indel_rate: float = 0.0):
    """BASELINE configs[2]: reads of `read_len` nt = a window of read_len / 3 residues of a database protein, substituted with a
    per-read rate U(0.1, 0.6), back-translated with uniform codon choice, on either strand.  Vectorised (10^5 reads in a second):
    returns the reads as an (n, read_len) array of nucleotide codes 0..3 = ACGT next to the database."""
    rng = np.random.default_rng(seed)
    dbl, dbo = make_db(n_db, rng)
    rq = np.random.default_rng([seed, 1 + q_stream])
    naa = read_len // 3
    lens = np.diff(dbo)
    src = rq.integers(0, n_db, n_reads)
    while True:  # proteins shorter than the window are redrawn
        bad = lens[src] < naa
        if not bad.any():
            break
        src[bad] = rq.integers(0, n_db, int(bad.sum()))
    st = (rq.random(n_reads) * (lens[src] - naa + 1)).astype(np.int64)
    aa = dbl[(dbo[src] + st)[:, None] + np.arange(naa)[None, :]].astype(np.int64)
    rate = rq.uniform(0.1, 0.6, n_reads)
    sub = rq.random((n_reads, naa)) < rate[:, None]
    aa[sub] = draw_letters(rq, int(sub.sum()))
    # codon table: for amino acid a, its codons as rows of nucleotide codes; uniform choice
    code = {c: i for i, c in enumerate("ACGT")}
    ncod = np.array([len(_CODONS[ALPHABET[a]]) for a in range(20)], dtype=np.int64)
    tab = np.zeros((20, int(ncod.max()), 3), dtype=np.uint8)
    for a in range(20):
        for k, cod in enumerate(_CODONS[ALPHABET[a]]):
            tab[a, k] = [code[x] for x in cod]
    pick = (rq.random((n_reads, naa)) * ncod[aa]).astype(np.int64)
    dna = tab[aa, pick].reshape(n_reads, naa * 3)
    if dna.shape[1] < read_len:
        dna = np.concatenate([dna, rq.integers(0, 4, (n_reads, read_len - dna.shape[1]), dtype=np.uint8)], axis=1)
    if indel_rate > 0.0:  # configs[4]: a single-nucleotide deletion or insertion inside the read (length kept) = one frameshift per affected read
        hit = np.flatnonzero(rq.random(n_reads) < indel_rate)
        pos = rq.integers(read_len // 4, 3 * read_len // 4, len(hit))
        ins = rq.random(len(hit)) < 0.5
        fill = rq.integers(0, 4, len(hit), dtype=np.uint8)
        for r, p, i, b in zip(hit.tolist(), pos.tolist(), ins.tolist(), fill.tolist()):
            if i:
                dna[r, p + 1:] = dna[r, p:-1].copy(); dna[r, p] = b
            else:
                dna[r, p:-1] = dna[r, p + 1:].copy(); dna[r, -1] = b
    rev = rq.random(n_reads) < 0.5
    dna[rev] = (3 - dna[rev])[:, ::-1]  # reverse complement: A<->T, C<->G with codes ACGT = 0123
    return {"dna_codes": np.ascontiguousarray(dna), "db_letters": dbl, "db_off": dbo, "src": src}


def write_dna_codes_fasta(path: str, codes: np.ndarray, prefix: str = "r") -> None:
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    text = lut[codes]
    with open(path, "wb") as f:
        chunks = []
        for i in range(codes.shape[0]):
            chunks.append(b">%s%d\n" % (prefix.encode(), i))
            chunks.append(text[i].tobytes())
            chunks.append(b"\n")
            if len(chunks) >= 30000:
                f.write(b"".join(chunks)); chunks = []
        f.write(b"".join(chunks))


BX_WORKLOADS = {"bx": (reads_workload, dict(seed=55))}  # goldens: tests/golden/bx.x0.tsv (blastx --fast, default flags)


WORKLOADS = {
    # name: (factory, kwargs)  -- the committed golden fixtures under tests/golden/ are keyed by these names
    "c1": (workload, dict(n_q=1000, n_db=10000, seed=1)),
    "fam2": (family_workload, dict(n_fam=4, fam_size=400, n_q=120, seed=12, member_div=(0.02, 0.12), query_div=(0.03, 0.3))),
    "edge": (edge_workload, dict(seed=21)),
    "long": (long_workload, dict(seed=33)),
    "rep": (repeat_workload, dict(seed=44)),
}


def named(name: str):
    f, kw = WORKLOADS[name]
    return f(**kw)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--nq", type=int, required=True)
    ap.add_argument("--ndb", type=int, required=True)
    ap.add_argument("--seed", type=int, required=True)
    ap.add_argument("--out", required=True, help="output prefix: writes <out>.q.faa and <out>.db.faa")
    a = ap.parse_args()
    w = workload(a.nq, a.ndb, a.seed)
    write_fasta(a.out + ".q.faa", w["q_letters"], w["q_off"], "q")
    write_fasta(a.out + ".db.faa", w["db_letters"], w["db_off"], "d")
    print(f"queries={a.nq} letters={len(w['q_letters'])}  db={a.ndb} letters={len(w['db_letters'])}")
