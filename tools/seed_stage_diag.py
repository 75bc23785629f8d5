"""One-call bisect of the device seed stage against the oracle (needs a GPU):

    python tools/seed_stage_diag.py [--sens 1] [--workloads edge,rep,fam2] > gpurun_out/seed_diag.txt

For every workload (masked with the reference's default masking when it has repeats / motifs) and every shape of the mode it
prints, WITHOUT stopping at the first difference:
  1. masking: hard-masked letters and the soft-masking table, device vs oracle;
  2. the reference-side seed index per shape: same locations, equal keys <=> equal seeds, locations ascending inside a key;
  3. dmnd_search_shape per shape: every stage counter on both sides, hits only one side has, hits with another score,
     SEED_MASK bits left on the query block.
The first line that says DIFF names the stage to look at.  Uses the oracle as the checker (test infrastructure)."""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from diamond_b200 import api, synth  # noqa: E402


def hit_key(h):
    return {(int(x["query"]), int(x["seed_offset"]), int(x["subject_score"]) & ((1 << 48) - 1)): int(x["subject_score"]) >> 48 for x in h}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sens", type=int, default=1)
    ap.add_argument("--workloads", default="edge,rep,fam2")
    ap.add_argument("--max-shapes", type=int, default=4)
    a = ap.parse_args()
    olib = api.load(os.path.join(ROOT, "oracle", "_build", "libdmnd_oracle.so"))
    plib = api.load()
    for name in a.workloads.split(","):
        w = synth.named(name)
        q_raw, q_lim = api.block_image(w["q_letters"], w["q_off"])
        r_raw, r_lim = api.block_image(w["db_letters"], w["db_off"])
        for masking in (0, 1):
            print(f"== {name} sensitivity {a.sens} masking {masking}")
            side = []
            for lib in (olib, plib):
                c = api.Context(lib, threads=8, sensitivity=a.sens)
                qb, rb = c.upload(q_raw, q_lim), c.upload(r_raw, r_lim)
                if masking:
                    c.mask_block(qb, 5, 0, len(q_lim) - 1); c.mask_block(rb, 5, 0, len(r_lim) - 1)
                st = {"ql": c.download_letters(qb, q_raw.size), "rl": c.download_letters(rb, r_raw.size),
                      "qs": c.debug_block_soft(qb, q_raw.size), "rs": c.debug_block_soft(rb, r_raw.size), "shapes": []}
                ns = min(c.params.n_shapes, a.max_shapes)
                for sid in range(ns):
                    keys, locs = c.debug_ref_index(rb, sid, r_raw.size)
                    hits, cn = c.search_shape(qb, rb, sid)
                    st["shapes"].append((keys, locs, hits, cn, c.download_letters(qb, q_raw.size)))
                side.append(st)
                c.free_block(qb); c.free_block(rb); c.close()
            o, g = side
            for k, what in (("ql", "query letters after masking"), ("rl", "reference letters after masking"), ("qs", "query soft table"), ("rs", "reference soft table")):
                d = np.flatnonzero(o[k] != g[k])
                print(f"  {'ok  ' if d.size == 0 else 'DIFF'} {what}: {d.size} positions differ {d[:10].tolist() if d.size else ''} (set: oracle {int((o[k] == (23 if k[1] == 'l' else 1)).sum())})")
            for sid, ((ko, lo, ho, co, mo), (kg, lg, hg, cg, mg)) in enumerate(zip(o["shapes"], g["shapes"])):
                # index: same location set; equal keys <=> equal seeds; ascending locations inside a device key
                same_locs = np.array_equal(np.sort(lo), np.sort(lg))
                asc = bool(np.all((kg[1:] != kg[:-1]) | (lg[1:] > lg[:-1]))) and bool(np.all(kg[1:] >= kg[:-1]))
                part_ok = None
                if same_locs and lo.size:
                    so = dict(zip(lo.tolist(), ko.tolist()))
                    fwd, bwd, part_ok = {}, {}, True
                    for loc, kd in zip(lg.tolist(), kg.tolist()):
                        s = so[loc]
                        if fwd.setdefault(kd, s) != s or bwd.setdefault(s, kd) != kd:
                            part_ok = False
                            break
                print(f"  {'ok  ' if same_locs and asc and part_ok is not False else 'DIFF'} shape {sid} reference index: entries oracle {lo.size} device {lg.size}; same locations {same_locs}; keys sorted + locations ascending inside a key {asc}; keys <=> seeds {part_ok}")
                print(f"  {'ok  ' if co == cg else 'DIFF'} shape {sid} counters oracle {co}")
                if co != cg:
                    print(f"       shape {sid} counters device {cg}")
                A, B = hit_key(ho), hit_key(hg)
                only_o, only_g = sorted(set(A) - set(B)), sorted(set(B) - set(A))
                other = [(k, A[k], B[k]) for k in sorted(set(A) & set(B)) if A[k] != B[k]]
                def show(k):
                    t = int(np.searchsorted(r_lim, k[2], side="right")) - 1
                    return f"(q{k[0]} off {k[1]} d{t}+{k[2] - int(r_lim[t])})"
                print(f"  {'ok  ' if not only_o and not only_g and not other else 'DIFF'} shape {sid} hits oracle {len(ho)} device {len(hg)}; only oracle {len(only_o)} {[show(k) for k in only_o[:6]]}; only device {len(only_g)} {[show(k) for k in only_g[:6]]}; "
                      f"other score {len(other)} {[(show(k), x, y) for k, x, y in other[:6]]}")
                d = np.flatnonzero(mo != mg)
                print(f"  {'ok  ' if d.size == 0 else 'DIFF'} shape {sid} SEED_MASK bits: {d.size} letters differ {d[:10].tolist() if d.size else ''}")


if __name__ == "__main__":
    main()
