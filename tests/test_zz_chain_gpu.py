"""dmnd_hits_chain on the device (cuda/chain.cu: sort by (query, target), one thread per pair, chain_kernels.cuh) against the oracle
library's twin (the host code of csrc/host through dmnd_host_chain_pair): per-query records and the DP problem list must be identical,
queries the device hands to the host path (capacities) must be a subset the oracle chains itself; and the chained DP call returns
the results of dmnd_banded_swipe on the same list.  Through the C ABI."""
import numpy as np
import pytest
from conftest import workload_blocks

pytestmark = pytest.mark.gpu
XDROP = 32  # only has to be the same on both sides


@pytest.mark.parametrize("name,max_targets", [("edge", 64), ("rep", 64), ("fam2", 64), ("fam2", 2000), ("c1", 3)])
def test_hits_chain_matches_oracle(oracle_lib, product_lib, name, max_targets):
    from diamond_b200 import api
    w, q_raw, q_lim, r_raw, r_lim = workload_blocks(name)
    out = []
    for lib in (oracle_lib, product_lib):
        c = api.Context(lib, threads=8, sensitivity=0, comp_based_stats=1)
        qb, rb = c.upload(q_raw, q_lim), c.upload(r_raw, r_lim)
        c.compute_bias(qb, 1)
        o = c.hits_chain(qb, rb, 0, xdrop=XDROP, max_targets=max_targets, align=True)
        if lib is product_lib and len(o["problems"]):
            res2, _ = c.banded_swipe(qb, rb, o["problems"], True)
            assert np.array_equal(res2["score"], o["results"]["score"]) and np.array_equal(res2["q_begin"], o["results"]["q_begin"])
        out.append(o)
        c.free_block(qb); c.free_block(rb); c.close()
    o, g = out
    assert np.array_equal(o["queries"]["query"], g["queries"]["query"]) and np.array_equal(o["queries"]["n_targets"], g["queries"]["n_targets"])
    assert np.array_equal(o["queries"]["n_hits"], g["queries"]["n_hits"]) and o["n_pairs"] == g["n_pairs"]
    host_o, host_g = (o["queries"]["flags"] & 1) != 0, (g["queries"]["flags"] & 1) != 0
    assert not (host_o & ~host_g).any(), "the oracle flags a query for the host path that the device chained"
    # problems of the queries both sides chained
    both = ~host_g
    po = [o["problems"][q["first"]:q["first"] + q["n_problems"]] for q in o["queries"][both]]
    pg = [g["problems"][q["first"]:q["first"] + q["n_problems"]] for q in g["queries"][both]]
    assert all(np.array_equal(a, b) for a, b in zip(po, pg)), "DP problem lists differ"
    assert sum(len(a) for a in po) > 0 or name == "c1"
    if name == "fam2" and max_targets == 64:
        assert host_g.any(), "fam2 has queries with more than 64 targets"
    # host-path queries: the same hits (as a set per query), segments and sites
    for qo, qg in zip(o["queries"][host_o], g["queries"][host_o]):
        a = np.sort(o["host_hits"][qo["first"]:qo["first"] + qo["n_hits"]], order=["subject_score", "seed_offset"])
        b = np.sort(g["host_hits"][qg["first"]:qg["first"] + qg["n_hits"]], order=["subject_score", "seed_offset"])
        assert np.array_equal(a, b)
