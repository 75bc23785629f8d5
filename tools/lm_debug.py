"""Device-vs-oracle bisect of the left-most filter (needs a GPU): for the hits of shape `sid` only the device returns, prints the
filter's intermediates on both sides (dmnd_debug_left_most) for every index chunk, then repeats the shape's search on the final
SEED_MASK state.    python tools/lm_debug.py [--workload edge] [--sens 1] [--sid 1] > gpurun_out/lm_debug.txt"""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from diamond_b200 import api, synth  # noqa: E402


def key(h):
    return {(int(x["query"]), int(x["seed_offset"]), int(x["subject_score"]) & ((1 << 48) - 1)) for x in h}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="edge"); ap.add_argument("--sens", type=int, default=1); ap.add_argument("--sid", type=int, default=1)
    a = ap.parse_args()
    w = synth.named(a.workload)
    q_raw, q_lim = api.block_image(w["q_letters"], w["q_off"])
    r_raw, r_lim = api.block_image(w["db_letters"], w["db_off"])
    libs = (api.load(os.path.join(ROOT, "oracle", "_build", "libdmnd_oracle.so")), api.load())
    ctxs, blocks, hits = [], [], []
    for lib in libs:
        c = api.Context(lib, threads=8, sensitivity=a.sens)
        qb, rb = c.upload(q_raw, q_lim), c.upload(r_raw, r_lim)
        hh = []
        for sid in range(a.sid + 1):
            h, cn = c.search_shape(qb, rb, sid)
            hh.append((h, cn))
        ctxs.append(c); blocks.append((qb, rb)); hits.append(hh)
    ko, kg = key(hits[0][a.sid][0]), key(hits[1][a.sid][0])
    print(f"shape {a.sid}: oracle {len(ko)} device {len(kg)} only-oracle {len(ko - kg)} only-device {len(kg - ko)}")
    nchunks = ctxs[0].params.index_chunks
    for (q, off, sloc) in sorted(kg - ko)[:8] + sorted(ko - kg)[:4]:
        qloc = int(q_lim[q]) + off
        print(f"pair q{q} off {off} qloc {qloc} sloc {sloc}")
        for chunk in range(nchunks):
            o = ctxs[0].debug_left_most(*blocks[0], a.sid, chunk, qloc, sloc)
            g = ctxs[1].debug_left_most(*blocks[1], a.sid, chunk, qloc, sloc)
            same = np.array_equal(o, g)
            print(f"  chunk {chunk}: {'same' if same else 'DIFF'}")
            for name, v in (("oracle", o), ("device", g)):
                n = int(v[29])
                cands = [f"pos{int(x >> 56)} v{int(x >> 48) & 1} cur{int(x >> 40) & 1} ok{int(x >> 36) & 1} fp{int(x >> 24) & 0xfff} part{int(x) & 0xffffff}" for x in v[9:9 + n]]
                print(f"    {name}: pass {int(v[0])} mm {int(v[1]):016x} seedbits {int(v[2]):016x} left raw/masked {int(v[3]):016x} right {int(v[4]):08x} geo {int(v[5]):x} {int(v[6]):x} vl {int(v[7])} vr {int(v[8])} cands {cands}")
    # the same search once more on the final SEED_MASK state
    for name, c, (qb, rb) in (("oracle", ctxs[0], blocks[0]), ("device", ctxs[1], blocks[1])):
        for rep in range(3):
            h, cn = c.search_shape(qb, rb, a.sid)
            print(f"repeat {rep} {name}: {len(h)} hits {cn}")
    for c, (qb, rb) in zip(ctxs, blocks):
        c.free_block(qb); c.free_block(rb); c.close()


if __name__ == "__main__":
    main()
