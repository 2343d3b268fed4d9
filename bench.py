#!/usr/bin/env python
"""bench.py -- k-mer lookups/s of the batched query path on N B200s (one process per GPU).

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): SSHash rebuilt from
the yeast chr01 unitigs of tests/data/pf1/yeast_chr01_index (k=31, w=15, skew 32, hash seed 0), queried
in random-access mode with synthetic 150 bp reads: 50 % sampled from the 230,218 bp reference on a
random strand, 50 % uniform random ACGT.  One "step" = one pass of the hot path (encode -> canonical
k-mers -> minimizers -> MPHF -> bucket bounds -> positions -> verify -> unitig id/offset/orientation)
over one batch of --reads reads per GPU (default 10 M reads = 1.2e9 k-mer lookups).  Multi-GPU:
index replicated, reads sharded (each rank its own batch, weak scaling), no data-path collective.

  value      whole-job lookups/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e        same metric through the C ABI with HOST (pinned) buffers: H2D of the reads and D2H of every
             16-byte hit record inside the timed region
  roofline   dominant kernel (query_reads_kernel<0>): algorithmic bytes (273.25 B/lookup for an
             SSHash random lookup) / measured kernel time vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline   the CPU oracle (a C++ port of mazu's query path; the Rust reference cannot be built here)
             timed on this box's host cores on a bounded sample of the same workload

--impl reference times that CPU port (all host threads) on the same config and prints the same line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

YEAST = os.path.join(ROOT, "tests", "data", "pf1", "yeast_chr01_index")
K, W, SKEW, SEED = 31, 15, 32, 0
READ_LEN = 150
ALG_BYTES_PER_LOOKUP = 273.25  # 8 sectors * 32 B + 17.25 B stream (SSHash random lookup)
METRIC = "k-mer lookups/sec (pos+neg, bit-exact)"
UNIT = "lookups/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mazu_b200", choices=["mazu_b200", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per step")
    ap.add_argument("--mode", default="random", choices=["random", "streaming"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU seconds for the cpu_baseline sample")
    return ap.parse_args()


def config_dict(args, n_gpus):
    return {
        "workload": "configs[1]: SSHash(yeast chr01 unitigs, k=31, m=15, skew=32, seed 0) random-access queries of "
                    "%d synthetic 150bp reads per GPU (50%% reference-sampled random strand / 50%% uniform random)" % args.reads,
        "mode": args.mode,
        "reads_per_gpu": args.reads,
        "read_len": READ_LEN,
        "lookups_per_step_per_gpu": args.reads * (READ_LEN - K + 1),
        "parallelism": "index replicated x%d, reads sharded by batch, no collective" % n_gpus,
        "l2_policy": "inputs (%.2f GB of reads + %.1f GB of results per step) are larger than the 126 MB L2" %
                     (args.reads * READ_LEN / 1e9, args.reads * (READ_LEN - K + 1) * 16 / 1e9),
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows = []
        self.stop = threading.Event()
        self.gpu = gpu_index
        self.th = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def load_oracle_index():
    import _oracle
    o = _oracle.OracleIndex.dense_from_pf1(YEAST)
    return o, o.rebuild_k2u(1, w=W, skew=SKEW, seed=SEED)


def ref_codes_from_oracle(o):
    import _gen
    return _gen.unpack_2bit(o.refseq_words(), int(o.ref_prefix()[-1]))


def cpu_sample_reads(ref_codes, n_reads, seed, mode):
    import _gen
    return _gen.sample_reads_fast(ref_codes, n_reads, READ_LEN, seed, frac_ref=0.5 if mode == "random" else 0.7,
                                  sub_rate=0.0 if mode == "random" else 0.01)


def time_oracle(os_idx, ref_codes, target_seconds, mode, threads, seed=4242):
    """CPU port timed on a bounded sample: calibrate on 20k reads, then size the sample for ~target_seconds."""
    streaming = mode == "streaming"
    cal = cpu_sample_reads(ref_codes, 20000, seed, mode)
    offs = np.arange(20001, dtype=np.uint64) * READ_LEN
    t0 = time.time()
    _, c, _ = os_idx.query_reads(cal, offs, streaming=streaming, want_hits=False, n_threads=threads)
    rate = float(c[0]) / max(time.time() - t0, 1e-6)
    n_reads = int(min(2_000_000, max(20000, rate * target_seconds / (READ_LEN - K + 1))))
    bases = cpu_sample_reads(ref_codes, n_reads, seed + 1, mode)
    offs = np.arange(n_reads + 1, dtype=np.uint64) * READ_LEN
    t0 = time.time()
    _, c, _ = os_idx.query_reads(bases, offs, streaming=streaming, want_hits=True, n_threads=threads)
    dt = time.time() - t0
    return float(c[0]) / dt, n_reads, dt, c


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The Rust crate cannot be built
    in this image (no cargo/rustc, four un-vendored crates), so this arm runs the C++ port (oracle/)
    with all host threads; each step is a bounded sample of the workload."""
    if rank != 0:
        return
    o, os_idx = load_oracle_index()
    ref_codes = ref_codes_from_oracle(o)
    threads = os.cpu_count() or 1
    streaming = args.mode == "streaming"
    # size one step for ~cpu_seconds / 2 of work
    cal = cpu_sample_reads(ref_codes, 20000, 99, args.mode)
    offs = np.arange(20001, dtype=np.uint64) * READ_LEN
    t0 = time.time()
    _, c, _ = os_idx.query_reads(cal, offs, streaming=streaming, want_hits=False, n_threads=threads)
    rate = float(c[0]) / max(time.time() - t0, 1e-6)
    n_reads = int(min(args.reads, max(20000, rate * 4.0 / (READ_LEN - K + 1))))
    bases = cpu_sample_reads(ref_codes, n_reads, 42, args.mode)
    offs = np.arange(n_reads + 1, dtype=np.uint64) * READ_LEN
    for _ in range(args.warmup):
        os_idx.query_reads(bases[:READ_LEN * 20000], offs[:20001], streaming=streaming, want_hits=True, n_threads=threads)
    t0 = time.time()
    total = 0
    for _ in range(args.steps):
        _, c, _ = os_idx.query_reads(bases, offs, streaming=streaming, want_hits=True, n_threads=threads)
        total += int(c[0])
    dt = time.time() - t0
    v = total / dt
    sample = "%d reads x %d bp (%d lookups) per step, %d steps" % (n_reads, READ_LEN, n_reads * (READ_LEN - K + 1), args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": config_dict(args, args.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "C++ restatement of mazu's query path (oracle/); the Rust reference cannot be built in this image"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1 and "RANK" not in os.environ:
        # launched without torchrun: re-launch one process per GPU
        port = 29500 + (os.getpid() % 2000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import _gen
    import mazu_b200 as mz

    if not torch.cuda.is_available() or mz.device_count() <= 0:
        raise SystemExit("bench.py needs a CUDA device: mazu_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    mode = mz.MODE_RANDOM if args.mode == "random" else mz.MODE_STREAMING
    # ---- index: pufferfish yeast_chr01 unitigs -> SSHash::from_unitig_set(k=31, w=15, skew 32, seed 0), replicated per GPU
    dense = mz.DenseIndex.deserialize_from_cpp(YEAST, device=local_rank)
    index = dense.rebuild_k2u(mz.K2U_SSHASH, w=W, skew_param=SKEW, seed=SEED)
    assert index.k == K

    # ---- synthetic reads, generated on the device (seed 42 + rank); reference codes from refseq.bin
    import _oracle  # only the host-side fixture reader + checker; never on the timed path
    o, os_idx = load_oracle_index() if rank == 0 else (None, None)
    if rank == 0:
        ref_codes = ref_codes_from_oracle(o)
    else:
        oo = _oracle.OracleIndex.dense_from_pf1(YEAST)
        ref_codes = ref_codes_from_oracle(oo)
        del oo
    n_reads = args.reads
    nk_per_read = READ_LEN - K + 1
    n_lookups = n_reads * nk_per_read
    gen = torch.Generator(device=dev)
    gen.manual_seed(42 + rank)
    ref_t = torch.from_numpy(ref_codes.astype(np.uint8)).to(dev)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    comp = torch.tensor([3, 2, 1, 0], dtype=torch.uint8, device=dev)
    frac_ref = 0.5 if args.mode == "random" else 0.7
    sub_rate = 0.0 if args.mode == "random" else 0.01
    bases = torch.empty(n_reads * READ_LEN, dtype=torch.uint8, device=dev)
    CH = 1_000_000
    ar = torch.arange(READ_LEN, device=dev)
    for r0 in range(0, n_reads, CH):
        n = min(CH, n_reads - r0)
        starts = torch.randint(0, len(ref_codes) - READ_LEN + 1, (n,), generator=gen, device=dev)
        codes = ref_t[starts[:, None] + ar[None, :]]
        strand = torch.rand(n, generator=gen, device=dev) < 0.5
        rc = comp[codes.flip(1).long()]
        codes = torch.where(strand[:, None], rc, codes)
        is_ref = torch.rand(n, generator=gen, device=dev) < frac_ref
        rnd = torch.randint(0, 4, (n, READ_LEN), generator=gen, device=dev, dtype=torch.uint8)
        codes = torch.where(is_ref[:, None], codes, rnd)
        if sub_rate > 0:
            m = torch.rand((n, READ_LEN), generator=gen, device=dev) < sub_rate
            codes = torch.where(m, (codes + torch.randint(1, 4, (n, READ_LEN), generator=gen, device=dev, dtype=torch.uint8)) & 3, codes)
        bases[r0 * READ_LEN:(r0 + n) * READ_LEN] = acgt[codes.long()].reshape(-1)
    del ref_t
    hits = torch.empty((n_lookups, 4), dtype=torch.int32, device=dev)
    counts = torch.zeros(3, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        index.query_reads(bases, None, n_reads=n_reads, uniform_read_len=READ_LEN, mode=mode, out_hits=hits, counts=counts,
                          mem=mz.MEM_DEVICE, stream=stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    counts.zero_()
    # ---- timed region: exactly K steps, CUDA events on the launching stream, per-step events for the kernel time
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with ClockSampler(local_rank) as clk:
        barrier()
        evs[0].record(stream)
        for i in range(args.steps):
            step()
            evs[i + 1].record(stream)
        barrier()
    total_ms = evs[0].elapsed_time(evs[-1])
    step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    cnt = counts.cpu().numpy().astype(np.int64)
    assert cnt[0] == n_lookups * args.steps, "every window of the synthetic reads is a valid k-mer"
    tot_cnt = torch.tensor(cnt, dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tot_cnt, op=dist.ReduceOp.SUM)  # final gather of per-shard hit counts (off the hot path)
    tot_cnt = tot_cnt.cpu().numpy()
    value = float(n_lookups) * args.steps * world / (max_ms * 1e-3)
    kernel_ms = float(np.mean(step_ms))

    # ---- e2e: same metric through the C ABI with HOST buffers (pinned): H2D reads + D2H every hit record, per step
    e2e = None
    if not args.no_e2e:
        h_bases = torch.empty(n_reads * READ_LEN, dtype=torch.uint8, pin_memory=True)
        h_bases.copy_(bases)
        h_hits = torch.empty((n_lookups, 4), dtype=torch.int32, pin_memory=True)
        h_cnt = np.zeros(3, dtype=np.uint64)
        hb = h_bases.numpy()
        e2e_steps = args.steps

        def e2e_step():
            index.query_reads(hb, None, n_reads=n_reads, uniform_read_len=READ_LEN, mode=mode, out_hits=h_hits, counts=h_cnt)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_val = float(n_lookups) * e2e_steps * world / float(tt.item())
        e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": n_reads * READ_LEN * world,
               "d2h_bytes_per_step": (n_lookups * 16 + 24) * world, "steps": e2e_steps,
               "api": "mazu_b200_query_reads(MAZU_MEM_HOST): pinned host reads in, every 16-byte mazu_hit_t out"}
        # the e2e result must equal the device-resident result
        same = bool(torch.equal(h_hits[: 120 * 50000], hits[: 120 * 50000].cpu()))
        e2e["matches_device_path"] = same

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- parity spot check against the oracle on the first reads of this exact batch
    n_chk = min(n_reads, 5000)
    chk_bases = bases[: n_chk * READ_LEN].cpu().numpy()
    want, wcnt, _ = os_idx.query_reads(chk_bases, np.arange(n_chk + 1, dtype=np.uint64) * READ_LEN, streaming=(mode == mz.MODE_STREAMING))
    got = hits[: n_chk * nk_per_read].cpu().numpy().view(np.uint32).reshape(-1).view(mz.HIT_DTYPE)
    parity_ok = bool(np.array_equal(got, want))

    # ---- CPU baseline (port) on a bounded sample, rank 0 only
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        v, n_s, dt, c = time_oracle(os_idx, ref_codes, args.cpu_seconds, args.mode, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d reads x %d bp (%d lookups) of the same workload, %.1f s" % (n_s, READ_LEN, int(c[0]), dt),
               "note": "C++ restatement of mazu's query path (oracle/); the Rust reference cannot be built in this image"}
        v1, n1, dt1, c1 = time_oracle(os_idx, ref_codes, min(args.cpu_seconds, 6.0), args.mode, 1)
        cpu["single_thread_value"] = v1
        cpu["ns_per_kmer_single_thread"] = 1e9 / v1

    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = n_lookups * ALG_BYTES_PER_LOOKUP / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr_path):
        try:
            tj = json.load(open(tr_path))
            per_lookup = tj.get("query_reads_kernel<%d>" % (0 if args.mode == "random" else 1), {}).get("dram_bytes_per_lookup")
            if per_lookup is not None:
                traffic = per_lookup * n_lookups
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic (torch.Generator seed 42+rank; reads sampled from refseq.bin of the yeast_chr01 fixture)",
        "config": config_dict(args, world),
        "clocks": clk.summary(),
        "gpu_launches": args.steps,
        "kernel": "mazu::query_reads_kernel<%d>" % (0 if args.mode == "random" else 1),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                     "algorithmic_bytes_per_lookup": ALG_BYTES_PER_LOOKUP, "lookups_per_launch": n_lookups, "kernel_ms": kernel_ms,
                     "note": "the yeast index (~1 MB) is L2-resident: DRAM carries only the read/result stream; see DESIGN.md"},
        "counts": {"n_kmers": int(tot_cnt[0]), "n_hit": int(tot_cnt[1]), "n_miss": int(tot_cnt[2])},
        "parity_spot_check_vs_oracle": parity_ok,
        "index_device_bytes": index.device_bytes,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
