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


def make_db(n_db: int, rng: np.random.Generator):
    lens = np.clip(rng.gamma(4.0, 75.0, n_db), 50, 2000).astype(np.int64)
    off = np.zeros(n_db + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    letters = rng.choice(20, size=int(off[-1]), p=RR_FREQ).astype(np.int8)
    return letters, off


def make_queries(n_q: int, db_letters: np.ndarray, db_off: np.ndarray, rng: np.random.Generator, qlen: int = 300):
    n_db = len(db_off) - 1
    src = rng.integers(0, n_db, n_q)
    slen = db_off[src + 1] - db_off[src]
    wlen = np.minimum(slen, qlen)
    start = (rng.random(n_q) * (slen - wlen + 1)).astype(np.int64)
    rate = rng.uniform(0.1, 0.6, n_q)
    woff = np.zeros(n_q + 1, dtype=np.int64)
    np.cumsum(wlen, out=woff[1:])
    total = int(woff[-1])
    qid = np.repeat(np.arange(n_q), wlen)
    pos_in_w = np.arange(total, dtype=np.int64) - woff[qid]
    base = db_letters[db_off[src][qid] + start[qid] + pos_in_w]
    sub = rng.random(total) < rate[qid]
    bg = rng.choice(20, size=total, p=RR_FREQ).astype(np.int8)
    base = np.where(sub, bg, base)
    keep = rng.random(total) >= 0.01
    ins = rng.random(total) < 0.01
    ins_letter = rng.choice(20, size=total, p=RR_FREQ).astype(np.int8)
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


def workload(n_q: int, n_db: int, seed: int, qlen: int = 300):
    rng = np.random.default_rng(seed)
    dbl, dbo = make_db(n_db, rng)
    ql, qo, src = make_queries(n_q, dbl, dbo, rng, qlen)
    return {"q_letters": ql, "q_off": qo, "db_letters": dbl, "db_off": dbo, "src": src}


def write_fasta(path: str, letters: np.ndarray, off: np.ndarray, prefix: str) -> None:
    lut = np.frombuffer(ALPHABET.encode(), dtype=np.uint8)
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
