#!/usr/bin/env python
"""bench.py -- GCUPS of the blastp --fast hot path (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA library behind the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the UNMODIFIED reference (oracle/_ref/diamond) on host cores

One "step" = one full pass of the hot path (seed search stages 0-2, extension rounds 1+2, culling) over one batch of
synthetic queries against the resident reference block.  Workload at N=1 = BASELINE.json configs[1]: 1 M synthetic
queries (len <= 300) x 100 k-protein DB, blastp --fast; at N>1 every rank processes its own 1 M-query block against the
same DB (query sharding, no data-path collective; the packed reference block is NCCL-broadcast once) -> "weak".
GCUPS numerator = algorithmic DP cells = sum over every banded DP problem of rounds 1 and 2 of band x cols
(dp/dp.h:121-124); it is a property of the workload (the DP target list is parity-checked against the reference).
`value` times steps with both blocks already resident in HBM; `e2e` times dmnd_blastp() with pinned HOST buffers
(block upload, problem lists, hit/result downloads inside).  Inputs (242 MB + 30 MB) exceed the 126 MB L2.
Flags: the reference's DEFAULTS (tantan masking of both blocks, motif soft masking, Hauser composition bias) on both arms.
Masking belongs to loading a block (run/double_indexed.cpp:122-127, :737-741): resident blocks are masked when they are
made resident (outside the `value` region, like the upload); `e2e` and the reference arm mask inside the timed region.
`--masking 0` runs both arms on the parity-ladder rung without masking (--masking 0 --motif-masking 0).
"""
import argparse, json, os, subprocess, sys, tempfile, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "diamond")
NO_MASKING = ["--masking", "0", "--motif-masking", "0"]  # parity-ladder rung L1; [] = the reference's default flags


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--queries", type=int, default=1_000_000, help="queries per GPU")
    ap.add_argument("--db", type=int, default=100_000)
    ap.add_argument("--sample", type=int, default=100_000, help="queries of the bounded CPU-baseline sample")
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--masking", type=int, default=1, choices=[0, 1], help="1 = reference default flags (tantan + motif masking), 0 = --masking 0 --motif-masking 0 on both arms")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.p, self.index = [], None, index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: [self.rows.append(l) for l in self.p.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for l in self.rows:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v == "Active":
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def effective_cpus():
    """min(visible CPUs, cgroup CFS quota): what 'all the host threads it can use' means inside a quota-limited container."""
    n = os.cpu_count() or 1
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            n = min(n, max(1, -(-int(q) // int(per))))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read()); per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, -(-q // per)))
        except Exception:
            pass
    return n


def make_workload(args, rank):
    from diamond_b200 import api, synth
    w = synth.workload(args.queries, args.db, args.seed, q_stream=rank)
    q_raw, q_lim = api.block_image(w["q_letters"], w["q_off"])
    r_raw, r_lim = api.block_image(w["db_letters"], w["db_off"])
    return w, q_raw, q_lim, r_raw, r_lim


def write_sample_fasta(w, n, td):
    from diamond_b200 import synth
    q, d = os.path.join(td, "q.faa"), os.path.join(td, "d.faa")
    synth.write_fasta(q, w["q_letters"][: w["q_off"][n]], w["q_off"][: n + 1], "q")
    synth.write_fasta(d, w["db_letters"], w["db_off"], "d")
    return q, d


def run_reference(q, d, out, threads, masking=1):
    t0 = time.perf_counter()
    r = subprocess.run([REF_BIN, "blastp", "--fast", "-q", q, "-d", d, "-f", "6", "-o", out, "-p", str(threads), "--log"] + ([] if masking else NO_MASKING),
                       capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("reference failed: " + r.stderr[-400:])
    return dt, r.stderr + r.stdout


def sample_cells_and_tsv(args, w, threads, device):
    """Algorithmic cell count (and fmt-6 text) of the bounded sample, from our pipeline -- the numerator both arms share."""
    from diamond_b200 import api
    n = min(args.sample, args.queries)
    q_raw, q_lim = api.block_image(w["q_letters"][: w["q_off"][n]], w["q_off"][: n + 1])
    r_raw, r_lim = api.block_image(w["db_letters"], w["db_off"])
    ctx = api.Context(device=device, threads=threads, masking=args.masking, motif_masking=args.masking)
    m, _, st = ctx.blastp(q_raw, q_lim, r_raw, r_lim)
    ctx.close()
    return st["cells_round1"] + st["cells_round2"], api.fmt6(m), n


def main():
    args = parse()
    if args.gpus > 1 and "RANK" not in os.environ:
        # launched plainly with --gpus N: re-launch as one rank per GPU (what the driver does itself with torch.distributed.run)
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                   "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533"), os.path.abspath(__file__)] + sys.argv[1:])
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    ncpu = effective_cpus()
    ref_threads = ncpu  # the reference arm uses every host thread; our seedp_bits must follow the same -p (setup.cpp:306-309)
    config = {"workload": f"blastp --fast, {args.queries} synthetic queries (len<=300) per GPU x {args.db}-protein DB (BASELINE configs[1])",
              "queries_per_gpu": args.queries, "db_seqs": args.db, "parallelism": f"query-sharded x{world}", "seed": args.seed,
              "flags": "--fast (reference defaults: tantan masking, motif masking, comp-based-stats 1) -k 25 -e 0.001" if args.masking else "--fast --masking 0 --motif-masking 0 --comp-based-stats 1 -k 25 -e 0.001",
              "masking": "resident blocks are masked when made resident (block load); e2e and the reference arm mask inside the timed region" if args.masking else "off on both arms",
              "reference_threads": ref_threads, "visible_cpus": os.cpu_count(),
              "l2": "inputs (242 MB queries + 30 MB reference per GPU) larger than the 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return
        import torch
        w, *_ = make_workload(args, 0)
        if not os.path.exists(REF_BIN):
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/diamond not present in the snapshot"}))
            return
        cells, _, n = sample_cells_and_tsv(args, w, ref_threads, 0) if torch.cuda.is_available() else (None, None, min(args.sample, args.queries))
        with tempfile.TemporaryDirectory() as td:
            q, d = write_sample_fasta(w, n, td)
            times = []
            for s in range(args.warmup + args.steps):
                dt, log = run_reference(q, d, os.path.join(td, "o.tsv"), ref_threads, args.masking)
                if s >= args.warmup:
                    times.append(dt)
        T = sum(times)
        val = (cells * len(times) / T / 1e9) if cells else None
        sample = f"first {n} queries of rank 0's block x full {args.db}-protein DB, one reference process per step (FASTA in, fmt 6 out), -p {ref_threads}"
        print(json.dumps({"impl": "reference", "metric": "GCUPS blastp --fast", "value": val, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * T / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "int8/int16 saturating SIMD (AVX2)", "data": "synthetic", "config": dict(config, sample=sample),
                          "cpu_baseline": {"value": val, "unit": "GCUPS", "cores": ref_threads, "kind": "reference", "sample": sample},
                          "e2e": {"value": val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import numpy as np
    import torch
    from diamond_b200 import api
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.environ.setdefault("DMND_HOST_THREADS", str(max(1, ncpu // world)))
    w, q_raw, q_lim, r_raw, r_lim = make_workload(args, rank)
    if world > 1:
        # the packed reference block travels once over NVLink (NCCL broadcast from rank 0); every rank then adopts it
        from diamond_b200 import shard
        if rank != 0:
            r_raw, r_lim = np.zeros(0, np.int8), np.zeros(0, np.int64)
        r_raw, r_lim = shard.broadcast_reference(r_raw, r_lim, dist, device=torch.device("cuda", local))
    # pinned host copies for the e2e path
    q_pin = torch.from_numpy(q_raw).pin_memory().numpy()
    r_pin = torch.from_numpy(r_raw).pin_memory().numpy()
    ctx = api.Context(device=local, threads=ref_threads, masking=args.masking, motif_masking=args.masking)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_ms = []  # host wall time of every step of the last timed() call (each step ends with a device->host read)

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.timing(reset=True)
        t0 = time.perf_counter()
        e0.record()
        last = None
        step_ms.clear()
        for _ in range(n):
            ts = time.perf_counter()
            last = fn()
            step_ms.append(round((time.perf_counter() - ts) * 1e3, 2))
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 0.0)
        ms = max(ms, 0.0) if ms > 0.5 * wall * 1e3 else wall * 1e3  # events sit on torch's idle stream: fall back to wall if skewed
        if world > 1:
            tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, last, ctx.timing()

    qb, rb = ctx.upload(q_raw, q_lim), ctx.upload(r_raw, r_lim)
    q_res, r_res = q_raw, r_raw
    mask_info = None
    if args.masking:
        # making the blocks resident includes masking them (what the reference does when it loads a block); the host images
        # the resident call reads (chaining link scores) are the equally masked letters
        torch.cuda.synchronize()
        t0 = time.perf_counter(); nmq = len(ctx.mask_block(qb, 5, 0, len(q_lim) - 1)); t1 = time.perf_counter()
        nmr = len(ctx.mask_block(rb, 5, 0, len(r_lim) - 1)); t2 = time.perf_counter()
        mask_info = {"query_block_ms": round((t1 - t0) * 1e3, 2), "reference_block_ms": round((t2 - t1) * 1e3, 2), "letters_masked": [nmq, nmr],
                     "note": "dmnd_block_mask (tantan + motif table) of the whole block incl. the position list download, one call each, host wall"}
        q_res, r_res = ctx.download_letters(qb, q_raw.size), ctx.download_letters(rb, r_raw.size)
    step_res = lambda: ctx.blastp_resident(qb, rb, q_res, q_lim, r_res, r_lim)
    step_e2e = lambda: ctx.blastp(q_pin, q_lim, r_pin, r_lim)
    for _ in range(args.warmup):
        step_res()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, (m, _, st), tm = timed(step_res, args.steps)
    step_ms_res = list(step_ms)
    clk = clocks.stop() if rank == 0 else None
    for _ in range(min(args.warmup, 2)):  # the e2e path has its own first-call allocations (block pool, staging)
        step_e2e()
    ms_e2e, (m2, _, st2), tm2 = timed(step_e2e, args.steps)
    step_ms_e2e = list(step_ms)
    cells = st["cells_round1"] + st["cells_round2"]
    tot = torch.tensor([float(cells), float(st2["cells_round1"] + st2["cells_round2"])], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot)
    cells_all, cells_all_e2e = float(tot[0].item()), float(tot[1].item())
    value = cells_all * args.steps / (ms / 1e3) / 1e9
    e2e = cells_all_e2e * args.steps / (ms_e2e / 1e3) / 1e9
    # kernel time for the roofline: query lanes overlap on the device, so their stream times cannot be added up; the DP
    # kernels are timed in a separate pass with ONE lane (every kernel of the step serialised on one stream, same work)
    lanes_env = os.environ.get("DMND_LANES")
    os.environ["DMND_LANES"] = "1"
    rsteps = max(1, min(args.steps, 2))
    step_res()
    _, _, tm1 = timed(step_res, rsteps)
    if lanes_env is None:
        del os.environ["DMND_LANES"]
    else:
        os.environ["DMND_LANES"] = lanes_env
    ctx.free_block(qb); ctx.free_block(rb)

    out = None
    if rank == 0:
        # roofline of the dominant kernels (banded SWIPE, both rounds): integer-ALU bound, see DESIGN.md
        dp_ms = (tm1["dp_score_ms"] + tm1["dp_trace_ms"]) / rsteps
        laneops = st["cells_round1"] * 9 + st["cells_round2"] * 13  # 9 lane-ops / score cell, +4 for the trace masks
        # peak: every DPX instruction (VIADDMNMX / VIMNMX3) retires two of those lane-ops; its issue rate is measured live
        peak = 2.0 * ctx.int_peak()
        ach = laneops / (dp_ms / 1e3) / 1e12 if dp_ms > 0 else None
        roofline = {"bound": "int-alu", "kernel": "swipe_kernel<R,*> (banded SWIPE rounds 1+2)", "achieved": ach, "peak": peak, "unit": "Tlaneop/s",
                    "frac": (ach / peak) if (ach and peak) else None,
                    # dram__bytes_read + dram__bytes_write of swipe_prof_kernel<4,1>, one launch over 200 k queries (ncu --set full,
                    # profiles/ncu_summary_r1.txt); 1.70e9 of it is the algorithmic trace (16 R bytes per macro step and problem)
                    "traffic": 1.825e9, "traffic_note": "ncu capture of the dominant launch at 200 k queries; the kernel is ALU-pipe bound (89.5 % busy), not HBM bound", "kernel_ms_per_step": dp_ms,
                    "kernel_gcups": cells / (dp_ms / 1e3) / 1e9 if dp_ms > 0 else None,
                    "seed_stage_ms_per_step": tm1["seed_ms"] / rsteps, "timed": f"CUDA events on the library stream, {rsteps} single-lane step(s) after the timed region"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline and os.path.exists(REF_BIN):
            scells, tsv, n = sample_cells_and_tsv(args, w, ref_threads, local)
            with tempfile.TemporaryDirectory() as td:
                q, d = write_sample_fasta(w, n, td)
                dt, log = run_reference(q, d, os.path.join(td, "o.tsv"), ref_threads, args.masking)
                same = open(os.path.join(td, "o.tsv")).read() == tsv
            cpu = {"value": scells / dt / 1e9, "unit": "GCUPS", "cores": ref_threads, "kind": "reference",
                   "sample": f"first {n} queries x full DB, one reference run (FASTA in, fmt 6 out), wall {dt:.2f} s; fmt-6 identical to ours: {same}"}
        out = {"metric": "GCUPS blastp --fast", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32 (exact; reference int8/int16 lane semantics)",
               "data": "synthetic", "config": config, "clocks": clk,
               "e2e": {"value": e2e, "unit": "GCUPS", "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": tm2["h2d_bytes"] // args.steps,
                       "d2h_bytes_per_step": tm2["d2h_bytes"] // args.steps},
               "gpu_launches": int(tm["launches"]), "roofline": roofline, "cpu_baseline": cpu,
               "step_ms": {"resident": step_ms_res, "e2e": step_ms_e2e}, "masking_ms": mask_info,
               "breakdown_ms_per_step": {"seed_stage": st["seed_ms"], "host_bridge": st["host_bridge_ms"], "dp_round1": st["dp1_ms"], "dp_round2": st["dp2_ms"], "total": st["total_ms"]},
               "work": {"cells_per_gpu": cells, "dp_problems_round1": st["dp_problems_round1"], "dp_problems_round2": st["dp_problems_round2"], "dp_problems_fused": st["dp_problems_fused"],
                        "note": "cells = sum band x cols over the reference's round-1 and round-2 problem lists (dp/dp.h:121-124): a property of the workload, "
                                "the same for both arms; fused queries evaluate a surviving problem's matrix once (with traceback) instead of twice",
                        "alignments": int(len(m)), "hits": st["hits"]}}
        print(json.dumps(out))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
