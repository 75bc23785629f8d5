"""Intermediate state of the device seed stage against the oracle (dmnd_debug_block_soft / dmnd_debug_ref_index): the soft-masking
table of both blocks and, per shape of the default sensitivity, the reference-side index -- same locations, equal keys <=> equal
seeds, locations ascending inside a key (what the stage-2 window filter's call sizes depend on).  These localise a difference
that tests/test_zz_sens1_gpu.py only sees in the hits; tools/seed_stage_diag.py prints the same comparison for every stage.
(Sorted last: diagnostics.)"""
import numpy as np
import pytest
from conftest import workload_blocks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["edge", "rep"])
def test_soft_masking_table_matches_oracle(oracle_lib, product_lib, name):
    from diamond_b200 import api
    w, q_raw, q_lim, r_raw, r_lim = workload_blocks(name)
    out = []
    for lib in (oracle_lib, product_lib):
        c = api.Context(lib, threads=8, sensitivity=1)
        qb, rb = c.upload(q_raw, q_lim), c.upload(r_raw, r_lim)
        c.mask_block(qb, 5, 0, len(q_lim) - 1); c.mask_block(rb, 5, 0, len(r_lim) - 1)
        out.append((c.debug_block_soft(qb, q_raw.size), c.debug_block_soft(rb, r_raw.size)))
        c.free_block(qb); c.free_block(rb); c.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    if name == "rep":
        assert out[0][1].sum() > 0


@pytest.mark.parametrize("name,masking", [("edge", 1), ("fam2", 0)])
def test_reference_index_order_and_content(oracle_lib, product_lib, name, masking):
    from diamond_b200 import api
    w, q_raw, q_lim, r_raw, r_lim = workload_blocks(name)
    idx = []
    for lib in (oracle_lib, product_lib):
        c = api.Context(lib, threads=8, sensitivity=1)
        rb = c.upload(r_raw, r_lim)
        if masking:
            c.mask_block(rb, 5, 0, len(r_lim) - 1)
        idx.append([c.debug_ref_index(rb, sid, r_raw.size) for sid in (0, 1)])
        c.free_block(rb); c.close()
    for sid in (0, 1):
        (ko, lo), (kg, lg) = idx[0][sid], idx[1][sid]
        assert np.array_equal(np.sort(lo), np.sort(lg)), f"shape {sid}: the device indexes other locations"
        assert np.all(kg[1:] >= kg[:-1]) and np.all((kg[1:] != kg[:-1]) | (lg[1:] > lg[:-1])), f"shape {sid}: locations of a key not ascending"
        seed_of = dict(zip(lo.tolist(), ko.tolist()))
        fwd, bwd = {}, {}
        for loc, kd in zip(lg.tolist(), kg.tolist()):
            s = seed_of[loc]
            assert fwd.setdefault(kd, s) == s and bwd.setdefault(s, kd) == kd, f"shape {sid}: device key {kd} / seed {s} do not correspond at location {loc}"
