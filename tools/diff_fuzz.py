"""Differential fuzz of the CLI (oracle-linked build by default) against the unmodified reference: random small workloads x random
option combinations from the surface the path implements; every run must give byte-identical output.

    python tools/diff_fuzz.py [--runs 40] [--seed 1] [--cli oracle/_build/dmnd-oracle-cli]

Needs oracle/_ref/diamond (make ref).  Prints the failing command lines, exit code 1 if any."""
import argparse, os, random, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diamond_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "diamond")
MODES = [["--fast"], [], ["--mid-sensitive"], ["--sensitive"], ["--more-sensitive"], ["--very-sensitive"]]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=40)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cli", default=os.path.join(ROOT, "oracle", "_build", "dmnd-oracle-cli"))
    ap.add_argument("--frameshift", type=float, default=None, help="probability of -F on a blastx run (default 0.35); 1 with --translated-only = frameshift runs only")
    ap.add_argument("--translated-only", action="store_true")
    ap.add_argument("--filters", type=float, default=0.25, help="probability of --id / --query-cover / --subject-cover on a run")
    ap.add_argument("--protein-only", action="store_true")
    ap.add_argument("--titles", type=float, default=0.3, help="probability of descriptive / merged sequence titles in both files")
    ap.add_argument("--dmnd", type=float, default=0.15, help="probability that the database is a .dmnd file made by the reference")
    ap.add_argument("--all-vs-all", type=float, default=0.08, help="probability that a protein run searches the database against itself (70 %% of those with --no-self-hits)")
    ap.add_argument("--format", default=None, help="use this output format on every run (6, 6f, 6g, 6c, 0, 5, 100, sam, paf)")
    ap.add_argument("--blocks", type=float, default=0.15, help="probability of -b (several reference blocks) on a run without -F")
    ap.add_argument("--coords-only", action="store_true", help="include field lists of coordinates only (known to differ in rare ties, see the comment at fmt 6c)")
    ap.add_argument("--keep", default=None, help="directory that receives the inputs and both outputs of every differing run")
    ap.add_argument("--min-score", type=float, default=0.1, help="probability of --min-score on a run")
    a = ap.parse_args()
    rnd = random.Random(a.seed)
    bad = 0
    with tempfile.TemporaryDirectory() as td:
        for run in range(a.runs):
            seed = rnd.randrange(1 << 30)
            translated = not a.protein_only and (a.translated_only or rnd.random() < 0.4)
            d = os.path.join(td, "d.faa")
            if translated:
                w = synth.reads_workload(seed, n_db=rnd.choice([300, 800]), n_q=rnd.choice([80, 200]))
                q = os.path.join(td, "q.fna")
                synth.write_dna_fasta(q, w["dna"])
                if rnd.random() < 0.15:  # the same reads as FASTQ (titles then end in a newline, as the reference's FASTQ reader leaves them)
                    recs = open(q).read().split(">")[1:]
                    q = os.path.join(td, "q.fastq")
                    with open(q, "w") as fh:
                        for rec in recs:
                            name, _, seq = rec.partition("\n")
                            seq = seq.replace("\n", "")
                            if seq: fh.write("@%s\n%s\n+\n%s\n" % (name, seq, "".join(chr(33 + (7 * i) % 40) for i in range(len(seq)))))
            else:
                kind = rnd.choice(["edge", "rep", "fam", "plain"])
                if kind == "edge":
                    w = synth.edge_workload(seed, n_db=rnd.choice([300, 1000]), n_q=rnd.choice([60, 200]))
                elif kind == "rep":
                    w = synth.repeat_workload(seed, n_db=rnd.choice([300, 800]), n_q=rnd.choice([60, 150]))
                elif kind == "fam":
                    w = synth.family_workload(n_fam=2, fam_size=rnd.choice([40, 120]), n_q=30, seed=seed, member_div=(0.02, 0.2), query_div=(0.03, 0.4))
                else:
                    w = synth.workload(n_q=rnd.choice([50, 200]), n_db=rnd.choice([500, 2000]), seed=seed, threads=1)
                q = os.path.join(td, "q.faa")
                synth.write_fasta(q, w["q_letters"], w["q_off"], "q")
            synth.write_fasta(d, w["db_letters"], w["db_off"], "d")
            if rnd.random() < a.titles:  # descriptive titles, some of them merged records (" >" between the titles), and characters XML has to escape
                for path in (q, d):
                    out, k = [], 0
                    for l in open(path).read().split("\n"):
                        if l.startswith(">"):
                            k += 1
                            if k % 3 == 0: l += " first <desc> & 'quoted' >alt%d second \"desc\"" % k
                            elif k % 3 == 1: l += " only desc|with.pipe"
                        out.append(l)
                    open(path, "w").write("\n".join(out))
            opts = list(rnd.choice(MODES))
            if rnd.random() < 0.1: opts += [rnd.choice(["--salltitles", "--sallseqid"])]
            if rnd.random() < a.dmnd:  # a DIAMOND database file instead of FASTA (blocks are then cut by letters, titles come from the file)
                subprocess.run([REF, "makedb", "--in", d, "-d", os.path.join(td, "db"), "--quiet"], check=True, capture_output=True)
                d = os.path.join(td, "db.dmnd")
            if rnd.random() < 0.15: opts += ["--header", "simple"]
            if rnd.random() < 0.1: opts += ["--ext", rnd.choice(["banded-fast", "banded-slow"])]
            if not translated and d.endswith(".faa") and rnd.random() < a.all_vs_all:  # the database against itself, with and without --no-self-hits
                q = d
                if rnd.random() < 0.7: opts += ["--no-self-hits"]
            opts += ["-p", str(rnd.choice([1, 4, 8]))]
            if rnd.random() < 0.3: opts += ["-c", str(rnd.choice([1, 2, 4]))]
            if rnd.random() < 0.3: opts += ["--masking", rnd.choice(["0", "1"])]
            if rnd.random() < 0.2: opts += ["--motif-masking", rnd.choice(["0", "1"])]
            if rnd.random() < 0.3: opts += ["--comp-based-stats", rnd.choice(["0", "1"])]
            r = rnd.random()
            if r < 0.25: opts += ["-k", str(rnd.choice([0, 1, 3, 50]))]
            elif r < 0.45: opts += ["--top", str(rnd.choice([0, 5, 30, 100]))]
            if rnd.random() < 0.3: opts += ["-e", rnd.choice(["10", "1e-10", "1e-30"])]
            if rnd.random() < a.min_score: opts += ["--min-score", str(rnd.choice([20, 40, 60, 100, 250]))]  # overrides -e
            if rnd.random() < a.filters:  # report filters: the extension's filtered schedule (align/extend.cpp:288, gapped_final.cpp:107-158)
                u = rnd.random()
                if u < 0.5 or rnd.random() < 0.3: opts += [rnd.choice(["--id", "--id", "--approx-id"]), str(rnd.choice([30, 50, 70, 90]))]  # (--approx-id >= 50 / 90 also raises the seed stage's Hamming cutoff)
                if u >= 0.3: opts += ["--query-cover", str(rnd.choice([20, 40, 60, 80, 95]))]
                if u >= 0.6: opts += ["--subject-cover", str(rnd.choice([10, 30, 55, 70, 90]))]
                if u >= 0.6 and rnd.random() < 0.5:  # equal covers >= 50: min_length_ratio, length-sorted blocks, the mutual-coverage seed stage (protein searches)
                    opts[-1] = opts[-3] = str(rnd.choice([50, 60, 80, 90]))
            if translated and rnd.random() < 0.15: opts += ["--query-gencode", str(rnd.choice([1, 2, 4, 11, 25]))]
            if translated:
                if rnd.random() < 0.3: opts += ["--strand", rnd.choice(["plus", "minus"])]
                if rnd.random() < 0.3: opts += ["--min-orf", str(rnd.choice([1, 10, 35]))]
            fshift = translated and rnd.random() < (a.frameshift if a.frameshift is not None else 0.35)  # blastx -F: the legacy pipeline + 3-frame DP
            if fshift:
                u = rnd.random()
                if u < 0.25 and "--top" not in opts and "-k" not in opts: opts += ["--long-reads"]  # = --range-culling --top 10 -F 15
                else:
                    opts += ["-F", str(rnd.choice([15, 15, 10, 20]))]
                    if u < 0.5:
                        opts += ["--range-culling"]
                        if rnd.random() < 0.3: opts += ["--range-cover", str(rnd.choice([20, 35, 80]))]
            if rnd.random() < a.blocks: opts += ["-b", rnd.choice(["0.00005", "0.0001", "0.0003"])]  # reference blocks + join_blocks
            fmt = rnd.choice(["6", "6", "6f", "0", "5", "sam", "paf", "100"]) if fshift else rnd.choice(["6", "6", "6f", "6g", "6c", "0", "5", "paf", "sam", "100"])
            if a.format: fmt = a.format
            if fmt == "6f" and fshift:
                opts += ["-f", "6", "qseqid", "sseqid", "pident", "length", "mismatch", "gapopen", "qstart", "qend", "sstart", "send", "evalue", "bitscore", "cigar", "btop", "qlen", "slen", "score", "qframe", "qseq_gapped", "sseq_gapped", "gaps", "nident", "qseq", "sseq", "qcovhsp", "scovhsp", "positive", "ppos", "qstrand", "qtitle", "stitle", "sallseqid", "salltitles", "qnum", "snum", "qseq_translated", "full_qseq", "qqual"]
                if rnd.random() < 0.5: opts += ["--unal", "1"]
            elif fmt == "6f":
                opts += ["-f", "6", "qseqid", "sseqid", "pident", "length", "mismatch", "gapopen", "qstart", "qend", "sstart", "send", "evalue", "bitscore", "cigar", "btop", "qlen", "slen", "score"]
                if rnd.random() < 0.5: opts += ["--unal", "1"]
            elif fmt == "6g":
                opts += ["-f", "6", "qseqid", "qtitle", "sseqid", "stitle", "positive", "ppos", "qcovhsp", "scovhsp", "qframe", "qstrand", "qseq", "sseq", "gaps", "nident", "qseq_gapped", "sseq_gapped", "sallseqid", "salltitles", "full_sseq", "full_qseq", "qnum", "snum", "full_qqual"] + (["qseq_translated"] if translated else [])
                if rnd.random() < 0.5: opts += ["--unal", "1"]
            elif fmt == "6c":  # short field lists: score-only (the reference's round 2 then runs without coordinates) and statistics without a transcript.
                # Lists of coordinates ONLY (qstart .. send, qcovhsp, scovhsp and nothing that needs the traceback) are left out unless --coords-only
                # is given: the reference then takes begin coordinates from a reversed score-only pass (HspValues::COORDS, dp/swipe/swipe_wrapper.cpp:98-99,364-444),
                # which among equally scoring alignments can pick another begin than its own traceback does; this build always traces back
                lists = [["qseqid", "sseqid", "evalue", "bitscore", "score"], ["qseqid", "sseqid", "length", "pident", "evalue"], ["qseqid", "sseqid", "qlen", "slen", "nident", "mismatch", "qstart", "send"]]
                if a.coords_only: lists += [["qseqid", "sseqid", "qstart", "qend", "sstart", "send", "evalue", "bitscore"], ["qseqid", "sseqid", "qlen", "slen", "qcovhsp", "scovhsp", "evalue"]]
                opts += ["-f", "6"] + rnd.choice(lists)
            elif fmt == "100":  # DAA: one block, and one thread -- the reference numbers its target dictionary in the order its threads reach the targets
                if "-b" in opts: del opts[opts.index("-b"):opts.index("-b") + 2]
                opts[opts.index("-p") + 1] = "1"
                opts += ["-f", "100"]
            elif fmt != "6":
                opts += ["-f", fmt]
                if fmt == "5" and rnd.random() < 0.3: opts += [rnd.choice(["--xml-blord-format", "--no-parse-seqids"])]
                if fmt == "sam" and rnd.random() < 0.3: opts += ["--sam-query-len"]
            if rnd.random() < 0.1 and not q.endswith(".gz") and not q.endswith(".fastq") and q != d and "qqual" not in opts and "full_qqual" not in opts:  # (with a quality field the reference loads nothing from a .gz file either)  # gzip-compressed FASTA queries (the reference built here loads NO query from a gzip-compressed FASTQ file; this CLI reads it)
                import gzip, shutil
                with open(q, "rb") as fi, gzip.open(q + ".gz", "wb") as fo: shutil.copyfileobj(fi, fo)
                q += ".gz"
            cmd = ["blastx" if translated else "blastp", "-q", q, "-d", d] + opts
            ext = ".daa" if fmt == "100" else ".out"  # (both programs append .daa to the name of a DAA file that lacks it)
            ro, oo = os.path.join(td, "r" + ext), os.path.join(td, "o" + ext)
            r1 = subprocess.run([REF] + cmd + ["-o", ro, "--quiet"], capture_output=True, text=True)
            r2 = subprocess.run([a.cli] + cmd + ["-o", oo], capture_output=True, text=True)
            def content(path):  # a SAM file quotes the program's own command line in its @PG line
                return [l for l in open(path, "rb") if not l.startswith(b"@PG")]
            ok = r1.returncode == 0 and r2.returncode == 0 and content(ro) == content(oo)
            n = sum(1 for _ in open(ro, "rb")) if r1.returncode == 0 else -1
            print(("ok   " if ok else "DIFF ") + f"run {run} seed {seed} lines {n}: " + " ".join(cmd[:1] + opts), flush=True)
            if not ok:
                bad += 1
                if a.keep:
                    import shutil
                    kd = os.path.join(a.keep, f"run{run}")
                    os.makedirs(kd, exist_ok=True)
                    for f in (q, d, ro, oo):
                        if os.path.exists(f): shutil.copy(f, kd)
                    open(os.path.join(kd, "cmd.txt"), "w").write(" ".join(cmd) + "\n")
                print("     ref rc", r1.returncode, r1.stderr[-200:], "| ours rc", r2.returncode, r2.stderr[-200:])
    print(f"{a.runs - bad} of {a.runs} identical")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
