"""Regenerates the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref/diamond, built from
/root/reference by oracle/ref_build/Makefile) on the seeded synthetic workloads of diamond_b200/synth.py.

    python tests/golden/make_golden.py            # needs oracle/_ref/diamond (make ref)

For every workload W and parity-ladder level L:
    L0 = --masking 0 --motif-masking 0 --comp-based-stats 0      L1 = --masking 0 --motif-masking 0 (Hauser CBS on)
    L2 = the reference's default flags (tantan masking of both blocks, motif soft-masking, Hauser CBS)
    S1 = no sensitivity flag at all: the reference's DEFAULT sensitivity (2 shapes of weight 10, stage-2 ungapped window filter)
         with its default flags -- `diamond blastp -q Q -d DB` as most users run it
    S2 = --mid-sensitive (8 shapes of weight 9, same filters) with default flags
    S3 = --sensitive (16 shapes of weight 8, stage-2 window filter, gapped filter) with default flags
    S4 / S5 / S6 = --more-sensitive / --very-sensitive / --ultra-sensitive (no motif masking, BANDED_SLOW bands; 14 x 7 and 64 x 7 shapes,
         Hamming cutoff 9, one index chunk for the last two)
    F0 = L2 in the BLAST pairwise format (-f 0): .txt
    T2 = L2 (--fast, default flags) with the transcript-bearing output fields: cigar, btop, qseq_gapped, sseq_gapped -- pins the
         traceback (and the masked letters the reference prints) byte for byte
    X0 / XT = `blastx --fast` with default flags on the DNA reads of synth.BX_WORKLOADS (12 default fields / + the transcript
         fields, qlen, slen): six translated frames per read, DNA coordinates, frame-aware culling
    X1 / X3 / X5 = blastx at the default sensitivity / --sensitive / --very-sensitive (whole-frame stage-2 window and the gapped filter's
         exceptions for short translated queries; cutoff_table_short differs from cutoff_table only in X5)
    XF / XF0 / XF3 = blastx in frameshift alignment mode (-F 15: the legacy extension pipeline + the 3-frame banded DP): --fast with
         the transcript fields (cigar / btop carry the \\ and / frameshift marks) + qframe, --fast in the pairwise format (.txt, with
         the "No hits found" records of every unaligned read), --sensitive with the default fields
    F5 = L2 in the BLAST XML format (-f 5): .xml
    D1 = --fast -k 2 -p 1 in the DAA format (-f 100): .daa, the binary alignment archive (headers, packed queries and transcripts, target dictionary)
    B1 = --fast -b 0.00003 --unal 1 -k 3: several reference blocks joined per query (output/join_blocks.cpp); a blocked run reports EVERY query
         without an alignment as unaligned, not only those with seed hits
    XX = blastx --fast -k 1 -e 1e-20 in the BLAST XML format (read coordinates, query frame)
    N1 = --fast --no-self-hits -k 3, the first 150 database sequences as queries (same titles): the alignment of a sequence with its own copy is dropped after
         round 2 (filter_hsp, align/culling.cpp:166-168), so two of the three best targets remain
    I1 = --fast with the report filters --id 60 --query-cover 50: the extension's filtered schedule (targets only sorted after round 1,
         round 2 in steps with Match::apply_filters, align/extend.cpp:288, align/gapped_final.cpp:107-158)
    M1 = default sensitivity with --query-cover 70 --subject-cover 70: equal covers >= 50 set min_length_ratio = 0.65 -- length-sorted
         blocks (queries are reported longest first) and the mutual-coverage seed stage (search/hamming/kernel_mutual_cov.h)
    XI / XFI = blastx --fast --id 50 --query-cover 60 / blastx --fast -F 15 --id 50 --subject-cover 20 (the legacy pipeline's
         Target::apply_filters, align/legacy/query_mapper.cpp:338-349)
    XFP / XFS = blastx --fast -F 15 in the PAF format / in the SAM format (-k 1 -e 1e-20; frameshift operations in the CIGAR, the @PG line
         quotes the reference's own command line and is skipped by the tests)
    XLB = blastx --long-reads -b 0.0002: several reference blocks, the join culling per query range
    XL = blastx --long-reads (= --range-culling --top 10 -F 15, default sensitivity): targets ranked and culled per query range
it writes  W.L.tsv  (fmt 6, byte-exact)  and  W.L.counters.json  (the --log stage counters, basic/basic.cpp:186-211).
The reference is always run with -p 8 (seedp_bits = 8) and default -c (4 index chunks): its output depends on both.
"""
import json, os, re, subprocess, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from diamond_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "diamond")
LEVELS = {"l0": ["--masking", "0", "--motif-masking", "0", "--comp-based-stats", "0"],
          "l1": ["--masking", "0", "--motif-masking", "0"],
          "l2": [],
          "s1": [],
          "s2": [],
          "s3": [], "s4": [], "s5": [], "s6": [],
          "t2": [], "f0": [],
          "f5": [], "b1": ["-b", "0.00003", "--unal", "1", "-k", "3"], "d1": ["-k", "2"],
          "n1": ["--no-self-hits", "-k", "3"],
          "i1": ["--id", "60", "--query-cover", "50"],
          "m1": ["--query-cover", "70", "--subject-cover", "70"]}
FORMAT = {"f0": "0", "f5": "5", "d1": "100"}  # BLAST pairwise (-f 0), BLAST XML (-f 5); everything else is tabular (-f 6)
EXT = {"f0": "txt", "f5": "xml", "d1": "daa"}
THREADS = {"d1": "1"}  # the reference numbers the DAA's target dictionary in the order its threads reach the targets: one thread = one order
FIELDS = {"t2": "qseqid sseqid pident length mismatch gapopen qstart qend sstart send evalue bitscore cigar btop qseq_gapped sseq_gapped".split()}
MODE = {"m1": [], "s1": [], "s2": ["--mid-sensitive"], "s3": ["--sensitive"], "s4": ["--more-sensitive"], "s5": ["--very-sensitive"], "s6": ["--ultra-sensitive"]}
SELF = ("n1",)  # levels whose queries are copies of database sequences (same titles)
ONLY = {"n1": ("fam2",), "f5": ("edge",), "b1": ("rep",), "d1": ("rep",), "i1": ("c1", "fam2", "edge"), "m1": ("fam2", "edge"), "f0": ("edge", "long"), "t2": ("c1", "edge", "long", "rep"), "s4": ("c1", "edge", "rep"), "s5": ("c1", "edge", "rep"), "s6": ("c1", "edge", "rep")}  # the many-shape modes: small workloads only  # every other level runs --fast
COUNTERS = {"seeds_hit": r"Seeds hit\s+= (\d+)", "seed_hits": r"Hits \(filter stage 0\) = (\d+)",
            "tentative_matches1": r"Hits \(filter stage 1\) = (\d+)", "tentative_matches2": r"Hits \(filter stage 2\) = (\d+)",
            "tentative_matches3": r"Hits \(filter stage 3\) = (\d+)", "targets": r"Target hits \(stage 0\) = (\d+)",
            "targets_round2": r"Target hits \(stage 5\) = (\d+)", "seedp_bits": r"Seed partition bits = (\d+)",
            "targets_extended": r"Target hits \(stage 3\) = (\d+)", "targets_stage2": r"Target hits \(stage 2\) = (\d+)"}


def main():
    if not os.path.exists(REF):
        sys.exit("oracle/_ref/diamond missing: run `make ref` where /root/reference is available")
    for name in synth.WORKLOADS:
        w = synth.named(name)
        with tempfile.TemporaryDirectory() as td:
            q, d = os.path.join(td, "q.faa"), os.path.join(td, "d.faa")
            synth.write_fasta(q, w["q_letters"], w["q_off"], "q")
            synth.write_fasta(d, w["db_letters"], w["db_off"], "d")
            qs = os.path.join(td, "qs.faa")  # SELF levels: the first 150 database sequences, under their database titles, as the queries
            synth.write_fasta(qs, w["db_letters"][:w["db_off"][150]], w["db_off"][:151], "d")
            for lvl, flags in LEVELS.items():
                if lvl in ONLY and name not in ONLY[lvl]:
                    continue
                out = os.path.join(HERE, f"{name}.{lvl}.{EXT.get(lvl, 'tsv')}")
                if os.path.exists(out) and "--missing" in sys.argv:
                    continue
                r = subprocess.run([REF, "blastp"] + MODE.get(lvl, ["--fast"]) + ["-q", qs if lvl in SELF else q, "-d", d, "-f", FORMAT.get(lvl, "6")] + FIELDS.get(lvl, []) + ["-o", out, "-p", THREADS.get(lvl, "8"), "--log"] + flags,
                                   capture_output=True, text=True, check=True)
                log = r.stderr + r.stdout
                cn = {k: int(re.search(p, log).group(1)) for k, p in COUNTERS.items()}
                json.dump(cn, open(os.path.join(HERE, f"{name}.{lvl}.counters.json"), "w"), indent=1, sort_keys=True)
                print(name, lvl, sum(1 for _ in open(out, "rb")), cn)


def main_blastx():
    for name in synth.BX_WORKLOADS:
        f, kw = synth.BX_WORKLOADS[name]
        w = f(**kw)
        with tempfile.TemporaryDirectory() as td:
            q, d = os.path.join(td, "q.fna"), os.path.join(td, "d.faa")
            synth.write_dna_fasta(q, w["dna"])
            synth.write_fasta(d, w["db_letters"], w["db_off"], "d")
            for lvl, fields in (("x0", []), ("xt", FIELDS["t2"] + ["score", "qlen", "slen"]), ("x1", []), ("x3", []), ("x5", []),
                                ("xf", FIELDS["t2"] + ["score", "qlen", "slen", "qframe"]), ("xf0", []), ("xf3", []), ("xl", []), ("xi", []), ("xfi", []), ("xx", []), ("xfp", []), ("xfs", []), ("xlb", [])):
                out = os.path.join(HERE, f"{name}.{lvl}." + {"xf0": "txt", "xx": "xml", "xfp": "paf", "xfs": "sam"}.get(lvl, "tsv"))
                if os.path.exists(out) and "--missing" in sys.argv:
                    continue
                mode = {"x1": [], "x3": ["--sensitive"], "x5": ["--very-sensitive"], "xf3": ["--sensitive"]}.get(lvl, ["--fast"])  # x1 = no flag: the default sensitivity
                if lvl.startswith("xf"):
                    mode = mode + ["-F", "15"]
                if lvl == "xl":
                    mode = ["--long-reads"]  # default sensitivity, --range-culling --top 10 -F 15
                if lvl == "xfs":
                    mode = mode + ["-k", "1", "-e", "1e-20"]
                if lvl == "xlb":
                    mode = ["--long-reads", "-b", "0.0002"]
                if lvl == "xx":
                    mode = mode + ["-k", "1", "-e", "1e-20"]
                if lvl == "xi":
                    mode = mode + ["--id", "50", "--query-cover", "60"]
                if lvl == "xfi":
                    mode = mode + ["--id", "50", "--subject-cover", "20"]
                r = subprocess.run([REF, "blastx"] + mode + ["-q", q, "-d", d, "-f", {"xf0": "0", "xx": "5", "xfp": "paf", "xfs": "sam"}.get(lvl, "6")] + fields + ["-o", out, "-p", "8", "--log"], capture_output=True, text=True, check=True)
                log = r.stderr + r.stdout
                cn = {k: int(re.search(p, log).group(1)) for k, p in COUNTERS.items()}
                json.dump(cn, open(os.path.join(HERE, f"{name}.{lvl}.counters.json"), "w"), indent=1, sort_keys=True)
                print(name, lvl, sum(1 for _ in open(out, "rb")), cn)


if __name__ == "__main__":
    main()
    main_blastx()
