"""Worker of tests/test_zzz_broadcast_gpu.py (one rank per GPU under torch.distributed.run): rank 0 makes the reference block resident
and masks it, dmnd_block_broadcast sends it device to device, every rank checks the received block byte for byte against rank 0's
and runs the same blastp step on it."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from diamond_b200 import api, shard, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = synth.named("rep")
q_raw, q_lim = api.block_image(w["q_letters"], w["q_off"])
r_raw, r_lim = api.block_image(w["db_letters"], w["db_off"])
ctx = api.Context(device=local, threads=8, masking=1, motif_masking=1, comp_based_stats=1)
rb = None
if rank == 0:
    rb = ctx.upload(r_raw, r_lim)
    ctx.mask_block(rb, 5, 0, len(r_lim) - 1)
rb, raw_len, lim = shard.broadcast_reference_block(ctx, dist, torch.device("cuda", local), rb)
assert raw_len == r_raw.size and np.array_equal(lim, r_lim)
letters = ctx.download_letters(rb, raw_len)
soft = ctx.debug_block_soft(rb, raw_len)
t = torch.from_numpy(np.concatenate([letters.view(np.uint8), soft])).cuda()
ref = t.clone()
dist.broadcast(ref, 0)
assert torch.equal(t, ref), f"rank {rank}: received block differs from rank 0's"
assert (letters != r_raw).any(), "the block arrived masked"
qb = ctx.upload(q_raw, q_lim)
ctx.mask_block(qb, 5, 0, len(q_lim) - 1)
m, _, st = ctx.blastp_resident(qb, rb, ctx.download_letters(qb, q_raw.size), q_lim, letters, lim)
gold = open(os.path.join(ROOT, "tests", "golden", "rep.l2.tsv")).read()
assert api.fmt6(m) == gold, f"rank {rank}: output on the broadcast block differs from the reference golden"
dist.barrier()
if rank == 0:
    print(f"broadcast ok on {world} ranks: {raw_len} letters, {len(m)} alignments per rank")
ctx.close()
dist.destroy_process_group()
