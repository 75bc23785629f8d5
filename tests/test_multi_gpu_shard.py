"""The N>1 path on CPU: two gloo ranks, each with the oracle-backed pipeline, shard one query block by contiguous ranges,
receive the packed reference block by broadcast from rank 0 and reproduce the unsharded reference golden."""
import os, subprocess, sys, textwrap
from conftest import ROOT, GOLDEN

WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, ROOT)
    import numpy as np, torch.distributed as dist
    from diamond_b200 import api, synth, shard
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    w = synth.named("c1")
    q_raw, q_lim = api.block_image(w["q_letters"], w["q_off"])
    if rank == 0:
        r_raw, r_lim = api.block_image(w["db_letters"], w["db_off"])
    else:
        r_raw, r_lim = np.zeros(0, np.int8), np.zeros(0, np.int64)
    r_raw, r_lim = shard.broadcast_reference(r_raw, r_lim, dist)
    b, e = shard.query_ranges(q_lim, world)[rank]
    sq_raw, sq_lim = shard.sub_block(q_raw, q_lim, b, e)
    ctx = api.Context(api.load(os.path.join(ROOT, "oracle", "_build", "libdmnd_oracle.so")), threads=8, comp_based_stats=1)
    m, _, st = ctx.blastp(sq_raw, sq_lim, r_raw, r_lim)
    ctx.close()
    allm = shard.gather_matches(m, b, dist)
    if rank == 0:
        assert api.fmt6(allm) == open(os.path.join(ROOT, "tests", "golden", "c1.l1.tsv")).read()
        print("SHARD_OK", world, len(allm))
    dist.destroy_process_group()
''')


def test_two_rank_query_sharding_matches_golden(oracle_lib, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text("ROOT = %r\n" % ROOT + WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541", DMND_HOST_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARD_OK 2" in r.stdout


def test_query_ranges_cover_and_balance():
    import numpy as np
    from diamond_b200 import shard
    lim = np.cumsum(np.r_[256, np.random.default_rng(0).integers(20, 500, 1000) + 1]).astype(np.int64)
    for parts in (1, 2, 3, 8):
        r = shard.query_ranges(lim, parts)
        assert r[0][0] == 0 and r[-1][1] == 1000 and all(r[k][1] == r[k + 1][0] for k in range(parts - 1))
        sizes = [lim[e] - lim[b] for b, e in r]
        assert max(sizes) - min(sizes) < 1200
