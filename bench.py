#!/usr/bin/env python
"""bench.py -- k-mer lookups/s of the batched query path on N B200s (one process per GPU).

Default workload (BASELINE.json configs[1], the configuration the metric is quoted on): SSHash rebuilt
from the yeast chr01 unitigs of tests/data/pf1/yeast_chr01_index (k=31, w=15, skew 32, hash seed 0),
queried in random-access mode with synthetic 150 bp reads: 50 % sampled from the 230,218 bp reference
on a random strand, 50 % uniform random ACGT.  One "step" = one pass of the hot path (encode ->
canonical k-mers -> minimizers -> MPHF -> bucket bounds -> positions -> verify -> unitig id / offset /
orientation) over one batch of --reads reads per GPU (default 10 M reads = 1.2e9 k-mer lookups).
Multi-GPU: index replicated, reads sharded (each rank its own batch, weak scaling), no data-path
collective; one all_reduce of three counters after the timed region.

Other workloads (--workload), same JSON line:
  config3        the same index through .as_streaming(): 70 % reference-sampled reads with 1 % substitutions / 30 % random
  config1        pufferfish yeast_chr01 DenseIndex (PFHash + C++ BooPHF): all reference + unitig k-mers, shuffled, via k2u_batch
  config4        U2Pos occurrence decode on a synthetic high-multiplicity unitig table (metric: occurrences/s)
  config5        synthetic unitig set (len 31+Geom(68)), SSHash k=31 w=19 skew 64, --scale 1.0 = 36,145,130 unitigs / ~2.5e9 k-mers;
                 index >> L2, reads generated on the device (70/30 mix); --mode random|streaming
  config5-kmers  the same index queried with a flat batch of random-order k-mers (k2u_batch): the random-access HBM regime

  value      whole-job units/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e        same metric through the C ABI with HOST (pinned) buffers: H2D of the inputs and D2H of every 16-byte result record
             (the headline); e2e.bound compares its D2H rate with the measured PCIe peak.  Reported next to it, never instead:
             compact_records (8-byte records), hit_runs (lossless run format, ~1.2 B per lookup, expansion checked),
             host_reads_in_device_records_out (records stay in HBM for the next device stage)
  roofline   dominant kernel: algorithmic bytes / measured kernel time vs the measured peak
  cpu_baseline   the CPU oracle (a C++ port of mazu's query path; the Rust reference cannot be built here)
             timed on this box's host cores on a bounded sample of the same workload

--impl reference times that CPU port (all host threads) on the same config and prints the same line.
"""
import argparse
import ctypes as C
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
READ_LEN = 150
ALG_SSHASH = 273.25   # 8 sectors * 32 B + 17.25 B stream per SSHash random lookup
ALG_PFHASH = 209.25   # 6 sectors + 17.25 B
METRIC = "k-mer lookups/sec (pos+neg, bit-exact)"
UNIT = "lookups/s"
HUMAN_UNITIGS = 36_145_130


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mazu_b200", choices=["mazu_b200", "reference"])
    ap.add_argument("--workload", default="config2", choices=["config1", "config2", "config3", "config4", "config5", "config5-kmers"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads (or k-mers / 120, or unitig queries) per GPU per step")
    ap.add_argument("--mode", default=None, choices=["random", "streaming"])
    ap.add_argument("--scale", type=float, default=0.1, help="config5: fraction of the human-scale unitig count")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU seconds for the cpu_baseline sample")
    ap.add_argument("--builder", default="gpu", choices=["host", "gpu"], help="config5: where SSHash::from_unitig_set runs")
    ap.add_argument("--validate", action="store_true", help="config5: also run k2u_validate_self on the device (every unitig k-mer)")
    a = ap.parse_args()
    if a.mode is None:
        a.mode = "streaming" if a.workload == "config3" else "random"
    return a


class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line).
    Read in-process through NVML: spawning nvidia-smi ten times a second visibly disturbed workloads whose
    whole timed region is a few tens of milliseconds (config 4).  Falls back to nvidia-smi if pynvml is missing."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows = []
        self.stop = threading.Event()
        self.gpu = gpu_index
        self.th = threading.Thread(target=self.run, daemon=True)
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                ids = [x.strip() for x in vis.split(",") if x.strip()]
                if gpu_index < len(ids) and ids[gpu_index].isdigit():
                    phys = int(ids[gpu_index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        act = lambda bit: "Active" if r & bit else "Not Active"
        # NVML bit masks: SwPowerCap 0x4, HwSlowdown 0x8, SwThermalSlowdown 0x20, HwThermalSlowdown 0x40
        return [str(sm), str(mx), "%.2f" % pw, act(0x8), act(0x40), act(0x20), act(0x4)]

    def run(self):
        while not self.stop.is_set():
            try:
                if self.nvml is not None:
                    self.rows.append(self.sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.02 if self.nvml is not None else 0.1)

    def __enter__(self):
        self.th.start()
        time.sleep(0.05 if self.nvml is not None else 0.25)
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = [num(r[0]) for r in self.rows if r and num(r[0]) is not None]
        mx = [num(r[1]) for r in self.rows if len(r) > 1 and num(r[1]) is not None]
        pw = [num(r[2]) for r in self.rows if len(r) > 2 and num(r[2]) is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------
# oracle helpers (checker + CPU baseline; never on the timed GPU path)
# ------------------------------------------------------------------------------------------------
def load_oracle_yeast(sshash=True):
    import _oracle
    o = _oracle.OracleIndex.dense_from_pf1(YEAST)
    return o, (o.rebuild_k2u(1, w=15, skew=32, seed=0) if sshash else o)


def yeast_ref_codes(o):
    import _gen
    return _gen.unpack_2bit(o.refseq_words(), int(o.ref_prefix()[-1]))


def mix_of(mode):
    return (0.5, 0.0) if mode == "random" else (0.7, 0.01)


def time_oracle_reads(os_idx, ref_codes, target_seconds, mode, threads, k, seed=4242):
    """CPU port on a bounded sample: calibrate on 20k reads, then size the sample for ~target_seconds."""
    import _gen
    frac_ref, sub = mix_of(mode)
    streaming = mode == "streaming"
    cal = _gen.sample_reads_fast(ref_codes, 20000, READ_LEN, seed, frac_ref, sub)
    offs = np.arange(20001, dtype=np.uint64) * READ_LEN
    t0 = time.time()
    _, c, _ = os_idx.query_reads(cal, offs, streaming=streaming, want_hits=False, n_threads=threads)
    rate = float(c[0]) / max(time.time() - t0, 1e-6)
    n_reads = int(min(2_000_000, max(20000, rate * target_seconds / (READ_LEN - k + 1))))
    bases = _gen.sample_reads_fast(ref_codes, n_reads, READ_LEN, seed + 1, frac_ref, sub)
    offs = np.arange(n_reads + 1, dtype=np.uint64) * READ_LEN
    t0 = time.time()
    _, c, _ = os_idx.query_reads(bases, offs, streaming=streaming, want_hits=True, n_threads=threads)
    dt = time.time() - t0
    return float(c[0]) / dt, n_reads, dt, c


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The Rust crate cannot be built in
    this image (no cargo/rustc, four un-vendored crates), so this arm runs the C++ port (oracle/) with all
    host threads on the default workload's index; each step is a bounded sample of the workload."""
    if rank != 0:
        return
    import _gen
    o, os_idx = load_oracle_yeast(sshash=args.workload != "config1")
    ref_codes = yeast_ref_codes(o)
    threads = os.cpu_count() or 1
    streaming = args.mode == "streaming"
    frac_ref, sub = mix_of(args.mode)
    k = os_idx.k
    cal = _gen.sample_reads_fast(ref_codes, 20000, READ_LEN, 99, frac_ref, sub)
    offs = np.arange(20001, dtype=np.uint64) * READ_LEN
    t0 = time.time()
    _, c, _ = os_idx.query_reads(cal, offs, streaming=streaming, want_hits=False, n_threads=threads)
    rate = float(c[0]) / max(time.time() - t0, 1e-6)
    n_reads = int(min(args.reads, max(20000, rate * 4.0 / (READ_LEN - k + 1))))
    bases = _gen.sample_reads_fast(ref_codes, n_reads, READ_LEN, 42, frac_ref, sub)
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
    sample = "%d reads x %d bp (%d lookups) per step, %d steps" % (n_reads, READ_LEN, n_reads * (READ_LEN - k + 1), args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": {"workload": workload_name(args), "mode": args.mode, "read_len": READ_LEN},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "C++ restatement of mazu's query path (oracle/); the Rust reference cannot be built in this image"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(args):
    return {
        "config1": "configs[0]: pufferfish yeast_chr01 DenseIndex (PFHash + BooPHF): all reference k-mers + unitig fw/rc k-mers, shuffled (seed 1), k2u_batch",
        "config2": "configs[1]: SSHash(yeast chr01 unitigs, k=31, m=15, skew=32, seed 0) random-access queries of %d synthetic 150bp reads per GPU "
                   "(50%% reference-sampled random strand / 50%% uniform random)" % args.reads,
        "config3": "configs[2]: same SSHash index via as_streaming(), %d synthetic 150bp reads per GPU (70%% reference-sampled with 1%% substitutions / 30%% random)" % args.reads,
        "config4": "configs[3]: u2pos occurrence decode, synthetic table (4096 refs, max ref 2^27, Zipf(1.2) multiplicity <= 65536), queries ~ multiplicity",
        "config5": "configs[4]: synthetic unitig set scale %.3g (%d unitigs, len 31+Geom(68)), SSHash k=31 w=19 skew 64, %d synthetic 150bp reads per GPU (70/30 mix)" %
                   (args.scale, int(HUMAN_UNITIGS * args.scale), args.reads),
        "config5-kmers": "configs[4] index (scale %.3g) queried with a flat random-order k-mer batch (50%% present / 50%% random), k2u_batch" % args.scale,
    }[args.workload]


def main():
    args = parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1 and "RANK" not in os.environ:
        port = 29500 + (os.getpid() % 2000)  # launched without torchrun: re-launch one process per GPU
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
    stream = torch.cuda.current_stream()
    gen = torch.Generator(device=dev)
    gen.manual_seed(42 + rank)
    mode = mz.MODE_RANDOM if args.mode == "random" else mz.MODE_STREAMING
    frac_ref, sub_rate = mix_of(args.mode)
    W = args.workload
    info = {}
    unit, metric = UNIT, METRIC
    check = None          # callable -> bool : parity spot check (rank 0)
    e2e_step = None       # callable: one end-to-end step through host buffers
    e2e_compact_step = None  # same with 8-byte records
    e2e_extra = {}        # name -> (callable, d2h bytes per step, api description): further end-to-end variants, reported next to the headline
    e2e_bytes = (0, 0)
    e2e_units = None      # units per e2e step (defaults to n_units)
    cpu_fn = None         # callable -> cpu_baseline dict
    t_build = time.time()

    # ---------------------------------------------------------------- indexes
    if W in ("config1", "config2", "config3"):
        dense = mz.DenseIndex.deserialize_from_cpp(YEAST, device=local_rank)
        index = dense if W == "config1" else dense.rebuild_k2u(mz.K2U_SSHASH, w=15, skew_param=32, seed=0)
        o, os_idx = load_oracle_yeast(sshash=W != "config1")
        ref_codes = yeast_ref_codes(o)
        alg = ALG_PFHASH if W == "config1" else ALG_SSHASH
        peak_kind = "hbm"
    elif W in ("config5", "config5-kmers"):
        n_unitigs = max(1000, int(HUMAN_UNITIGS * args.scale))
        words, n_bases, accum = _gen.synthetic_unitigs_packed(n_unitigs, 68, 31, seed=45)
        us = mz.UnitigSet(31, words, n_bases, accum)
        t_b = time.time()
        index = mz.SSHash.from_unitig_set(us, 19, 64, seed=0, device=local_rank, builder=args.builder)
        info["index_builder"] = {"where": args.builder, "seconds": round(time.time() - t_b, 2)}
        info["index"] = {"n_unitigs": index.n_unitigs, "n_kmers": index.n_kmers, "sum_unitigs_len": index.sum_unitigs_len,
                         "n_minimizers": index.n_minimizers, "n_kmers_in_skew_index": index.n_kmers_in_skew_index,
                         "mphf_levels": index.info(mz.INFO_MPHF_LEVELS)}
        useq_dev = torch.from_numpy(np.concatenate([words, np.zeros(2, dtype=np.uint64)]).view(np.int64)).to(dev)
        alg = ALG_SSHASH
        peak_kind = "prand"
    else:  # config4
        k = 31
        n_unitigs = 40_000  # Zipf(1.2) clipped to 65,536 has mean ~1,600: 40k unitigs give the ~6e7 occurrences SURVEY 8(d) sizes config 4 for
        codes, accum = _gen.synthetic_unitigs(2000, 68, k, seed=44)  # the K2U half is irrelevant for the decode; keep it tiny
        rng = np.random.default_rng(44)
        us = mz.UnitigSet(k, mz.pack_2bit(codes), len(codes), accum)
        # a unitig set with n_unitigs entries is needed for the offsets vector: build a trivial one
        lens = np.full(n_unitigs, k, dtype=np.uint64)
        accum4 = np.zeros(n_unitigs + 1, dtype=np.uint64)
        np.cumsum(lens, out=accum4[1:])
        words4 = rng.integers(0, np.iinfo(np.uint64).max, size=(2 * int(accum4[-1]) + 63) // 64, dtype=np.uint64)
        index = mz.PFHash.from_unitig_set(mz.UnitigSet(k, words4, int(accum4[-1]), accum4), device=local_rank)
        mult = np.minimum(rng.zipf(1.2, size=n_unitigs), 65536).astype(np.uint64)
        offsets = np.zeros(n_unitigs + 1, dtype=np.uint64)
        np.cumsum(mult, out=offsets[1:])
        n_occ = int(offsets[-1])
        n_refs, max_ref_len = 4096, 1 << 27
        ref_ids = rng.integers(0, n_refs, size=n_occ, dtype=np.uint64)
        poss = rng.integers(0, max_ref_len - 4096, size=n_occ, dtype=np.uint64)
        fws = rng.integers(0, 2, size=n_occ, dtype=np.uint64)
        pos_bits, ref_bits = 28, 13
        enc = (ref_ids << np.uint64(pos_bits + 1)) | (poss << np.uint64(1)) | fws
        index.attach_u2pos_piscem(mz.PackedVec.pack(enc, 1 + pos_bits + ref_bits), pos_bits + 1, (1 << pos_bits) - 1, mz.PackedVec.pack(offsets))
        info["index"] = {"n_unitigs": n_unitigs, "n_occs": n_occ, "encoding": "piscem 42-bit packed"}
        peak_kind = "hbm"
    info["index_build_s"] = round(time.time() - t_build, 2)
    info["index_device_bytes"] = index.device_bytes
    K = index.k
    nk_per_read = READ_LEN - K + 1

    # ---------------------------------------------------------------- inputs + step closures
    if W in ("config2", "config3", "config5"):
        n_reads = args.reads
        n_units = n_reads * nk_per_read
        if W == "config5":
            bases = _gen.device_reads_from_packed(torch, useq_dev, n_bases, n_reads, READ_LEN, gen, 0.7, 0.01)
        else:
            ref_t = torch.from_numpy(ref_codes.astype(np.uint8)).to(dev)
            bases = torch.empty(n_reads * READ_LEN, dtype=torch.uint8, device=dev)
            acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
            ar = torch.arange(READ_LEN, device=dev)
            for r0 in range(0, n_reads, 1_000_000):
                n = min(1_000_000, n_reads - r0)
                starts = torch.randint(0, len(ref_codes) - READ_LEN + 1, (n,), generator=gen, device=dev)
                codes = ref_t[starts[:, None] + ar[None, :]]
                strand = torch.rand(n, generator=gen, device=dev) < 0.5
                codes = torch.where(strand[:, None], 3 - codes.flip(1), codes)
                is_ref = torch.rand(n, generator=gen, device=dev) < frac_ref
                rnd = torch.randint(0, 4, (n, READ_LEN), generator=gen, device=dev, dtype=torch.uint8)
                codes = torch.where(is_ref[:, None], codes, rnd)
                if sub_rate > 0:
                    m = torch.rand((n, READ_LEN), generator=gen, device=dev) < sub_rate
                    codes = torch.where(m, (codes + torch.randint(1, 4, (n, READ_LEN), generator=gen, device=dev, dtype=torch.uint8)) & 3, codes)
                bases[r0 * READ_LEN:(r0 + n) * READ_LEN] = acgt[codes.long()].reshape(-1)
        hits = torch.empty((n_units, 4), dtype=torch.int32, device=dev)
        counts = torch.zeros(3, dtype=torch.int64, device=dev)
        kernel_name = "mazu::query_reads_kernel<%d,SSHASH,NATIVE>" % (0 if mode == mz.MODE_RANDOM else 1)
        launches_per_step = 1

        def step():
            index.query_reads(bases, None, n_reads=n_reads, uniform_read_len=READ_LEN, mode=mode, out_hits=hits, counts=counts,
                              mem=mz.MEM_DEVICE, stream=stream.cuda_stream)

        if not args.no_e2e:
            # pinned host buffers: 150 B in + 1920 B out per read.  With several ranks on one host the e2e step is a
            # bounded batch of the same reads so that N ranks never pin more than ~8 GB each.
            e2e_reads = n_reads if world == 1 else min(n_reads, 2_500_000)
            e2e_units = e2e_reads * nk_per_read
            h_bases = torch.empty(e2e_reads * READ_LEN, dtype=torch.uint8, pin_memory=True)
            h_bases.copy_(bases[: e2e_reads * READ_LEN])
            h_hits = torch.empty((e2e_units, 4), dtype=torch.int32, pin_memory=True)
            h_cnt = np.zeros(3, dtype=np.uint64)
            hb = h_bases.numpy()

            def e2e_step():
                index.query_reads(hb, None, n_reads=e2e_reads, uniform_read_len=READ_LEN, mode=mode, out_hits=h_hits, counts=h_cnt)

            e2e_bytes = (e2e_reads * READ_LEN, e2e_units * 16 + 24)
            h_hits8 = torch.empty((e2e_units, 2), dtype=torch.int32, pin_memory=True)

            def e2e_compact_step():
                index.query_reads(hb, None, n_reads=e2e_reads, uniform_read_len=READ_LEN, mode=mode, out_hits=h_hits8, counts=h_cnt, compact=True)

            def e2e_devout_step():  # reads from pinned host memory, records left in HBM for the next device stage, counters back
                index.query_reads(hb, None, n_reads=e2e_reads, uniform_read_len=READ_LEN, mode=mode, out_hits=hits, counts=h_cnt,
                                  mem=mz.MEM_HOST_IN_DEVICE_OUT)

            # hit runs: a lossless compact form of the same records (codes + run starts); the buffers are reused across steps
            h_codes = torch.empty(e2e_units, dtype=torch.uint8, pin_memory=True)
            h_runs = torch.empty((max(1 << 20, e2e_units // 8), 4), dtype=torch.int32, pin_memory=True)
            h_rro = torch.zeros(e2e_reads + 1, dtype=torch.int64, pin_memory=True)
            runs_state = {"n_runs": 0}

            def e2e_runs_step():
                n_runs = C.c_uint64(0)
                mz._check(mz.lib().mazu_b200_query_reads_runs(index._h, mz._any_ptr(hb), None, e2e_reads, READ_LEN, mode, None, mz._any_ptr(h_codes),
                                                              mz._any_ptr(h_runs), h_runs.shape[0], mz._any_ptr(h_rro), C.byref(n_runs), mz._np_ptr(h_cnt)))
                runs_state["n_runs"] = n_runs.value

            e2e_extra = {"hit_runs": (e2e_runs_step, None,
                         "mazu_b200_query_reads_runs: pinned host reads in; one code byte per k-mer slot + the 16-byte record of every run start + "
                         "per-read run offsets out (lossless: mazu_b200_expand_hit_runs rebuilds the exact records)"),
                         "host_reads_in_device_records_out": (e2e_devout_step, 24,
                         "MAZU_MEM_HOST_IN_DEVICE_OUT: pinned host reads in, 16-byte records stay in HBM (input of project_hits on the device), "
                         "the three counters of `kphf bench` (src/bin/kphf/main.rs:282-284) come back")}

        def check():
            if W == "config5":  # no oracle index at this scale: verify sampled hits directly against the packed sequence
                n = min(n_reads, 20000)
                h = hits[: n * nk_per_read].cpu().numpy().view(np.uint32).reshape(-1).view(mz.HIT_DTYPE)
                b = bases[: n * READ_LEN].cpu().numpy()
                lut = np.zeros(256, dtype=np.uint8)
                for i, ch in enumerate(b"ACGT"):
                    lut[ch] = i
                rk = _gen.kmer_words_from_codes(lut[b], K)
                slot = np.nonzero(h["match"] != 0)[0]
                fw = rk[(slot // nk_per_read) * READ_LEN + (slot % nk_per_read)]
                gpos = accum[h["unitig_id"][slot].astype(np.int64)] + h["pos"][slot].astype(np.uint64)
                wi, sh = (gpos >> np.uint64(5)).astype(np.int64), (gpos & np.uint64(31)) << np.uint64(1)
                wpad = np.concatenate([words, np.zeros(2, dtype=np.uint64)])
                lo = wpad[wi] >> sh
                hi = np.where(sh == 0, np.uint64(0), wpad[wi + 1] << ((np.uint64(64) - sh) & np.uint64(63)))
                uk = (lo | hi) & np.uint64((1 << (2 * K)) - 1)
                ident = h["match"][slot] == mz.IDENTITY_MATCH
                import _oracle
                rc = np.array([_oracle.lib().orc_revcomp(int(x), K) for x in fw[~ident][:20000]], dtype=np.uint64)
                ok = bool(np.array_equal(uk[ident], fw[ident])) and bool(np.array_equal(uk[~ident][:20000], rc))
                ok &= bool((h["unitig_len"][slot] == (accum[h["unitig_id"][slot].astype(np.int64) + 1] - accum[h["unitig_id"][slot].astype(np.int64)])).all())
                return ok and len(slot) > 0
            n_chk = min(n_reads, 5000)
            chk = bases[: n_chk * READ_LEN].cpu().numpy()
            want, _, _ = os_idx.query_reads(chk, np.arange(n_chk + 1, dtype=np.uint64) * READ_LEN, streaming=(mode == mz.MODE_STREAMING))
            got = hits[: n_chk * nk_per_read].cpu().numpy().view(np.uint32).reshape(-1).view(mz.HIT_DTYPE)
            return bool(np.array_equal(got, want))

        if W != "config5":
            def cpu_fn():
                threads = os.cpu_count() or 1
                v, n_s, dt, c = time_oracle_reads(os_idx, ref_codes, args.cpu_seconds, args.mode, threads, K)
                v1, _, _, _ = time_oracle_reads(os_idx, ref_codes, min(args.cpu_seconds, 6.0), args.mode, 1, K)
                return {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": "%d reads x %d bp (%d lookups) of the same workload, %.1f s" % (n_s, READ_LEN, int(c[0]), dt),
                        "single_thread_value": v1, "ns_per_kmer_single_thread": 1e9 / v1,
                        "note": "C++ restatement of mazu's query path (oracle/); the Rust reference cannot be built in this image"}

    elif W in ("config1", "config5-kmers"):
        if W == "config1":
            useq_codes = _gen.unpack_2bit(o.useq_words(), o.total_len)
            uk = _gen.kmer_words_from_codes(useq_codes, K)
            import _oracle
            qs = np.concatenate([_gen.kmer_words_from_codes(ref_codes, K), uk])
            rng = np.random.default_rng(1)
            reps = max(1, min(430, (args.reads * nk_per_read) // len(qs)))  # >= 2e8 queries at the default size
            q = np.tile(qs, reps)
            rng.shuffle(q)
            kmers = torch.from_numpy(q.view(np.int64)).to(dev)
            kernel_name = "mazu::k2u_batch_kernel (PFHash/BooPHF)"
        else:
            n = args.reads * nk_per_read
            kmers = _gen.device_kmers_from_packed(torch, useq_dev, n_bases, n, K, gen, 0.5)
            kernel_name = "mazu::k2u_batch_kernel (SSHash)"
        n_units = int(kmers.numel())
        hits = torch.empty((n_units, 4), dtype=torch.int32, device=dev)
        launches_per_step = 1

        def step():
            index.k2u_batch(kmers, out=hits, mem=mz.MEM_DEVICE, stream=stream.cuda_stream, n=n_units)

        if not args.no_e2e:
            h_k = torch.empty(n_units, dtype=torch.int64, pin_memory=True)
            h_k.copy_(kmers)
            h_hits = torch.empty((n_units, 4), dtype=torch.int32, pin_memory=True)
            hk = h_k.numpy().view(np.uint64)

            def e2e_step():
                index.k2u_batch(hk, out=h_hits)

            e2e_bytes = (n_units * 8, n_units * 16)

        def check():
            n_chk = min(n_units, 200000)
            got = hits[:n_chk].cpu().numpy().view(np.uint32).reshape(-1).view(mz.HIT_DTYPE)
            qq = kmers[:n_chk].cpu().numpy().view(np.uint64)
            if W == "config1":
                return bool(np.array_equal(got, os_idx.k2u_batch(qq, n_threads=os.cpu_count() or 1)))
            hit = got["match"] != 0  # verify hits against the packed sequence
            gpos = accum[got["unitig_id"][hit].astype(np.int64)] + got["pos"][hit].astype(np.uint64)
            wi, sh = (gpos >> np.uint64(5)).astype(np.int64), (gpos & np.uint64(31)) << np.uint64(1)
            wpad = np.concatenate([words, np.zeros(2, dtype=np.uint64)])
            ukk = ((wpad[wi] >> sh) | np.where(sh == 0, np.uint64(0), wpad[wi + 1] << ((np.uint64(64) - sh) & np.uint64(63)))) & np.uint64((1 << (2 * K)) - 1)
            ident = got["match"][hit] == mz.IDENTITY_MATCH
            return bool(np.array_equal(ukk[ident], qq[hit][ident])) and 0.3 < hit.mean() < 0.7

        if W == "config1":
            def cpu_fn():
                threads = os.cpu_count() or 1
                qq = kmers[: min(n_units, 40_000_000)].cpu().numpy().view(np.uint64)
                t0 = time.time()
                os_idx.k2u_batch(qq, n_threads=threads)
                dt = time.time() - t0
                return {"value": len(qq) / dt, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": "%d k-mers of the same shuffled query set, %.1f s" % (len(qq), dt),
                        "note": "C++ restatement of PFHash::k2u over the C++ BooPHF (oracle/)"}
    else:  # config4
        metric, unit = "unitig occurrences decoded+projected/sec", "occurrences/s"
        n_q = args.reads
        rng = np.random.default_rng(45 + rank)
        p = mult.astype(np.float64)
        qids = rng.choice(n_unitigs, size=n_q, p=p / p.sum()).astype(np.uint32)
        # queries drawn ~ multiplicity hit the heavy lists: bound the decoded output to ~1e9 occurrences (12 GB of records)
        cum = np.cumsum(mult[qids])
        n_q = int(max(1, min(n_q, np.searchsorted(cum, 1_000_000_000))))
        qids = qids[:n_q]
        d_q = torch.from_numpy(qids.view(np.int32)).to(dev)
        d_offs = torch.zeros(n_q + 1, dtype=torch.int64, device=dev)
        total = int(mult[qids].sum())
        d_out = torch.empty((total, 3), dtype=torch.int32, device=dev)
        n_units = total
        alg = (32.0 * n_q + (6 + 12) * total) / total  # bytes per occurrence: offsets pair per query + 42-bit word in + 12 B out
        kernel_name = "mazu::occ_fill_tma_kernel<false> (+ occ_lens_kernel + cub scan)"
        launches_per_step = 4

        def step():
            mz._check(mz.lib().mazu_b200_decode_occs(index._h, mz._any_ptr(d_q), n_q, mz._any_ptr(d_offs), mz._any_ptr(d_out), total, None,
                                                     mz.MEM_DEVICE, mz._any_ptr(stream.cuda_stream)))

        # the projected variant (GetRefPos::project_hits, src/index.rs:156-216): one hit record per query with a random position /
        # orientation on its unitig; measured after the main loop and reported as `projected`
        h_np = np.zeros(n_q, dtype=mz.HIT_DTYPE)
        h_np["unitig_id"] = qids
        h_np["unitig_len"] = 4096
        h_np["pos"] = rng.integers(0, 4096 - k + 1, size=n_q)
        h_np["match"] = rng.integers(1, 3, size=n_q)
        d_hits_q = torch.from_numpy(h_np.view(np.uint32).reshape(-1, 4).view(np.int32)).to(dev)

        def project_step():
            mz._check(mz.lib().mazu_b200_project_hits(index._h, mz._any_ptr(d_hits_q), n_q, mz._any_ptr(d_offs), mz._any_ptr(d_out), total, None,
                                                      mz.MEM_DEVICE, mz._any_ptr(stream.cuda_stream)))

        def projected_line():
            for _ in range(3):
                project_step()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(args.steps):
                project_step()
            b.record(stream)
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / args.steps
            # spot check against project_onto_u_occ on the host
            offs_h = d_offs[:2001].cpu().numpy().view(np.uint64)
            out_h = d_out[: int(offs_h[-1])].cpu().numpy().view(np.uint32)
            ok = True
            for q in range(0, 2000, 97):
                e = int(offsets[qids[q]])
                fwd = bool(fws[e])
                want_pos = int(h_np["pos"][q]) + int(poss[e]) if fwd else int(poss[e]) + (4096 - int(h_np["pos"][q])) - k
                o = 1 if h_np["match"][q] == 1 else 0
                want_o = o if fwd else o ^ 1
                r = out_h[int(offs_h[q])]
                ok &= (int(r[0]) == int(ref_ids[e])) and (int(r[1]) == want_pos % (1 << 32)) and (int(r[2]) == want_o)
            return {"value": total / (ms * 1e-3), "unit": unit, "ms_per_step": ms, "frac_of_hbm_peak": total * (alg + 16.0 * n_q / total) / (ms * 1e-3) / 1e9 / 6553.9,
                    "parity_spot_check": bool(ok), "kernel": "mazu::occ_fill_tma_kernel<true>"}

        info["projected_fn"] = projected_line

        def check():
            n_chk = min(n_q, 20000)
            offs_h = d_offs[: n_chk + 1].cpu().numpy().view(np.uint64)
            out_h = d_out[: int(offs_h[-1])].cpu().numpy().view(np.uint32)
            ok = np.array_equal(np.diff(offs_h), mult[qids[:n_chk]])
            first = offsets[qids[:n_chk]].astype(np.int64)
            ok &= np.array_equal(out_h[offs_h[:-1].astype(np.int64), 0], ref_ids[first].astype(np.uint32))
            ok &= np.array_equal(out_h[offs_h[:-1].astype(np.int64), 1], poss[first].astype(np.uint32))
            return bool(ok)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    if W in ("config2", "config3", "config5"):
        counts.zero_()
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
    cnt_line = None
    if W in ("config2", "config3", "config5"):
        tot_cnt = counts.clone()
        if world > 1:
            dist.all_reduce(tot_cnt, op=dist.ReduceOp.SUM)  # final gather of per-shard hit counts (off the hot path)
        tot_cnt = tot_cnt.cpu().numpy()
        cnt_line = {"n_kmers": int(tot_cnt[0]), "n_hit": int(tot_cnt[1]), "n_miss": int(tot_cnt[2])}
    value = float(n_units) * args.steps * world / (max_ms * 1e-3)
    kernel_ms = float(np.mean(step_ms))

    e2e = None
    if e2e_step is not None:
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        eu = e2e_units if e2e_units is not None else n_units
        e2e = {"value": float(eu) * args.steps * world / float(tt.item()), "unit": unit, "units_per_step_per_gpu": eu,
               "h2d_bytes_per_step": e2e_bytes[0] * world,
               "d2h_bytes_per_step": e2e_bytes[1] * world, "steps": args.steps,
               "api": "C ABI with MAZU_MEM_HOST: pinned host inputs in, every result record out",
               "matches_device_path": bool(torch.equal(h_hits[:1_000_000], hits[:1_000_000].cpu()))}
        try:  # what bounds it: the box's pinned-copy bandwidth with both directions busy (profiles/pcie_probe.py)
            pc = json.load(open(os.path.join(ROOT, "profiles", "r01_pcie.json")))
            d2h_gbs = e2e_bytes[1] * args.steps / float(tt.item()) / 1e9
            e2e["bound"] = {"kind": "pcie d2h (16 B per lookup out)", "achieved_d2h_gbs_per_gpu": d2h_gbs, "measured_d2h_peak_gbs": pc["d2h_gbs"],
                            "measured_d2h_with_h2d_busy_gbs": pc["duplex_each_gbs"], "frac_of_d2h_peak": d2h_gbs / pc["d2h_gbs"]}
        except Exception:
            pass
        if e2e_compact_step is not None:
            e2e_compact_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                e2e_compact_step()
            barrier()
            tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e["compact_records"] = {"value": float(eu) * args.steps * world / float(tt.item()), "unit": unit,
                                      "d2h_bytes_per_step": (eu * 8 + 24) * world,
                                      "api": "mazu_b200_query_reads_compact: 8-byte {unitig_id, pos|match<<30} records"}

        for name, (fn, d2h, api) in e2e_extra.items():
            fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                fn()
            barrier()
            tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            if name == "hit_runs":
                d2h = eu + 16 * runs_state["n_runs"] + 8 * (e2e_reads + 1) + 24
            e2e[name] = {"value": float(eu) * args.steps * world / float(tt.item()), "unit": unit, "h2d_bytes_per_step": e2e_bytes[0] * world,
                         "d2h_bytes_per_step": d2h * world, "api": api}
            if name == "hit_runs":  # the expansion on the host reproduces the full records (checked on the head of the batch)
                m = min(e2e_reads, 20000)
                exp = mz.ModIndex.expand_hit_runs(h_codes.numpy()[: m * nk_per_read], h_runs.numpy().view(np.uint32).reshape(-1).view(mz.HIT_DTYPE),
                                                  h_rro.numpy().view(np.uint64)[: m + 1], uniform_slots=nk_per_read)
                e2e[name]["expands_to_full_records"] = bool(np.array_equal(exp.view(np.uint32).reshape(-1, 4), hits[: m * nk_per_read].cpu().numpy().view(np.uint32)))
                e2e[name]["n_runs_per_step_per_gpu"] = runs_state["n_runs"]

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    parity_ok = bool(check()) if check else None
    if "projected_fn" in info:
        info["projected"] = info.pop("projected_fn")()
    if args.validate and W.startswith("config5"):
        c = index.k2u_validate_self()
        info["k2u_validate_self"] = {"n_queries": c[0], "n_identity": c[1], "n_twin": c[2], "n_fail": c[4], "n_fail_not_found": c[3],
                                     "note": "failures that are not 'not found' are duplicated canonical k-mers of the random unitig set (expected ~1.3 pairs at 2.5e9 k-mers)"}
    cpu = cpu_fn() if (cpu_fn and not args.no_cpu_baseline and world == 1) else None

    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    achieved = n_units * alg / (kernel_ms * 1e-3) / 1e9
    if W == "config4":
        note = "streaming decode: algorithmic bytes = 32 B (offset pair) per queried unitig + 6 B packed word in + 12 B record out per occurrence"
    else:
        note = ("algorithmic bytes = SURVEY 8(d) ideal-layout figure with nothing cached or shared; consecutive k-mers of a read share "
                "minimizer, bucket and window sectors and small indexes are L2-resident, so `traffic` (measured DRAM bytes) is far below it "
                "and frac can exceed 1 -- see DESIGN.md 7b for the DRAM/issue figures from ncu")
    if peak_kind == "hbm":
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_source = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"
    else:
        pr = json.load(open(os.path.join(ROOT, "profiles", "r01_prand.json")))
        peak = float(pr["32GiB"]["GBps"])
        peak_source = "P_rand: measured independent random 32-byte gathers over a 32 GiB table (profiles/r01_prand.json), in GB/s of sectors"
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr_path):
        try:
            tj = json.load(open(tr_path))
            if W in ("config2", "config3"):
                per = tj.get("query_reads_kernel<%d>" % (0 if mode == mz.MODE_RANDOM else 1), {}).get("dram_bytes_per_lookup")
            elif W == "config5" and mode == mz.MODE_RANDOM:
                per = tj.get("config5 query_reads_kernel<0>", {}).get("dram_bytes_per_lookup")
            elif W == "config5-kmers":
                per = tj.get("config5 k2u_batch_kernel", {}).get("dram_bytes_per_lookup")
            else:
                per = None
            traffic = per * n_units if per is not None else None
            if W == "config4":
                traffic = tj.get("occ_fill_kernel<false>", {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    line = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic (torch.Generator seed 42+rank on the device; fixture sequences from tests/data)",
        "config": {"workload": workload_name(args), "mode": args.mode, "units_per_step_per_gpu": n_units,
                   "parallelism": "index replicated x%d, inputs sharded by batch, no collective" % world,
                   "l2_policy": "inputs and results per step (%.2f GB) are larger than the 126 MB L2" % ((e2e_bytes[0] + e2e_bytes[1]) / 1e9 if e2e_bytes[0] else n_units * 17.25 / 1e9)},
        "clocks": clk.summary(),
        "gpu_launches": args.steps * launches_per_step,
        "kernel": kernel_name,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_source, "algorithmic_bytes_per_unit": alg, "units_per_launch": n_units, "kernel_ms": kernel_ms,
                     "step_ms": [round(x, 3) for x in step_ms], "note": note,
                     # what the kernel is actually limited by (ncu, profiles/*.summary.txt): measured DRAM bytes / time against the same peak
                     "dram_frac_measured": (traffic / (kernel_ms * 1e-3) / 1e9 / float(peaks.get("hbm_gbs", 6650.0))) if traffic else None},
        "parity_spot_check": parity_ok,
    }
    # both denominators, always -- the streaming HBM peak and the measured random-gather rate (32-byte sectors)
    try:
        pr = json.load(open(os.path.join(ROOT, "profiles", "r01_prand.json")))
        p_rand = float(pr["32GiB"]["GBps"])
        line["roofline"]["vs_hbm_stream_peak"] = achieved / float(peaks.get("hbm_gbs", 6650.0))
        line["roofline"]["vs_random_gather_peak"] = achieved / p_rand
        line["roofline"]["random_gather_peak_gbs_of_sectors"] = p_rand
    except Exception:
        pass
    line.update(info)
    if cnt_line:
        line["counts"] = cnt_line
    if e2e is not None:
        line["e2e"] = e2e
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line, default=lambda o: o.item() if hasattr(o, "item") else str(o)), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
