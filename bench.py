#!/usr/bin/env python
"""bench.py -- k-mer lookups/s of the batched query path on N B200s (one process per GPU).

Default workload = BASELINE.json configs[4], the largest single-GPU configuration and the only one in the HBM-random
regime the north star is about: a synthetic human-scale unitig set (seed 45: 36,145,130 unitigs of length 31 + Geom(68),
~2.46e9 k-mers), SSHash k=31 w=19 skew 64 built on the device (5.8 GB of tables + 1.8 GB of unitig lines >> the 126 MB L2),
queried in random-access mode (K2U::k2u per k-mer) with synthetic 150 bp reads generated on the device: 70 % sampled from
the unitig sequence on a random strand with 1 % substitutions, 30 % uniform random.  One "step" = one pass of the hot
path (encode -> canonical k-mers -> minimizers -> MPHF -> bucket bounds -> positions -> verify -> unitig id / offset /
orientation) over one batch of --reads reads per GPU (default 10 M reads = 1.2e9 k-mer lookups).
Multi-GPU: index replicated (every rank builds the same index on its GPU), reads sharded (each rank its own batch, weak
scaling), no data-path collective; one all_reduce of three counters after the timed region.

Other workloads (--workload), same JSON line:
  config2        configs[1]: SSHash(yeast chr01 unitigs, k=31, w=15, skew 32), 50 % reference-sampled / 50 % random reads (index L2-resident)
  config3        configs[2]: the same index through .as_streaming(): 70 % reference-sampled reads with 1 % substitutions / 30 % random
  config1        configs[0]: pufferfish yeast_chr01 DenseIndex (PFHash + C++ BooPHF): all reference + unitig k-mers, shuffled, via k2u_batch
  config4        configs[3]: U2Pos occurrence decode on a synthetic high-multiplicity unitig table (metric: occurrences/s)
  config5        the default; --scale shrinks the unitig set, --mode random|streaming
  config5-kmers  the same index queried with a flat batch of random-order k-mers (k2u_batch)

  value      whole-job units/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e        same metric through the C ABI with HOST (pinned) buffers, copies inside the timed region.  The host result of
             a read batch is the lossless hit-run format (mazu_b200_query_reads_runs: one code byte per k-mer slot + the
             16-byte record of every run start; mazu_b200_expand_hit_runs rebuilds every record, checked here): e2e.value.
             Reported next to it on a bounded slice of the same reads: full_records (every 16-byte record over PCIe, bound
             by the measured D2H peak), compact_records (8 bytes), host_reads_in_device_records_out (records stay in HBM)
  roofline   dominant kernel.  Random-access workloads: achieved = DRAM bytes/s the kernel moves (lookups/s x DRAM bytes
             per lookup from the ncu capture of this launch shape, profiles/traffic.json) against the measured
             random-access line-rate peak P_rand (profiles/r02_prand.json); the SURVEY 8(d) algorithmic figure is reported
             next to it.  Streaming decode (config4): algorithmic bytes against the measured HBM copy peak.
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
    ap.add_argument("--workload", default="config5", choices=["config1", "config2", "config3", "config4", "config5", "config5-kmers"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads (or k-mers / 120, or unitig queries) per GPU per step")
    ap.add_argument("--mode", default=None, choices=["random", "streaming"])
    ap.add_argument("--scale", type=float, default=1.0, help="config5: fraction of the human-scale unitig count (36,145,130 unitigs)")
    ap.add_argument("--cpu-scale", type=float, default=0.25, help="config5: unitig-set scale of the CPU port's index (its threaded builder needs ~11 s per 0.1 on 16 cores)")
    ap.add_argument("--e2e-slice", type=int, default=1_000_000, help="reads of the batch the full-record / compact e2e variants run on")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU seconds for the cpu_baseline sample")
    ap.add_argument("--builder", default="gpu", choices=["host", "gpu"], help="config5: where SSHash::from_unitig_set runs")
    ap.add_argument("--validate", action="store_true", help="config5: also run k2u_validate_self on the device (every unitig k-mer)")
    ap.add_argument("--no-chain", action="store_true", help="config2/3: skip the get_ref_pos chain measurement (reads -> MappedRefPos)")
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
            self.sample_nvml()  # NVML's first queries are slow (one run recorded a single sample in a 150 ms region): pay that here
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


def mix_of(workload, mode):
    """(fraction of reads sampled from the indexed sequence, substitution rate)."""
    if workload in ("config5", "config3"):
        return 0.7, 0.01
    return (0.5, 0.0) if mode == "random" else (0.7, 0.01)


class CpuPort:
    """The CPU arm of a read workload: the oracle index + a generator of reads of the same mix as the GPU arm.
    config2/3: the yeast index (exactly the GPU arm's index).  config5: the same synthetic unitig set (seed 45) at
    --cpu-scale, because the port's builder cannot produce the full-scale index inside a bench run; its threaded builder
    is the reference's rayon builder restated (oracle/mazu_oracle.hpp)."""

    def __init__(self, args):
        import _gen
        import _oracle
        self.args = args
        self.threads = os.cpu_count() or 1
        self.frac_ref, self.sub = mix_of(args.workload, args.mode)
        t0 = time.time()
        if args.workload == "config5":
            self.scale = min(args.cpu_scale, args.scale)
            n_unitigs = max(1000, int(HUMAN_UNITIGS * self.scale))
            words, n_bases, accum = _gen.synthetic_unitigs_packed(n_unitigs, 68, 31, seed=45)
            self.idx = _oracle.OracleIndex.from_packed(31, words, n_bases, accum, 1, w=19, skew=64, seed=0, build_threads=self.threads)
            self.codes = _gen.unpack_2bit(words, n_bases)
            self.index_note = ("index: the same synthetic unitig set at scale %.3g (%d unitigs, %d k-mers; the GPU arm's is scale %.3g), "
                               "built by the port's threaded builder in %.0f s" % (self.scale, n_unitigs, self.idx.n_kmers, args.scale, time.time() - t0))
        else:
            o, self.idx = load_oracle_yeast(sshash=args.workload != "config1")
            self.codes = yeast_ref_codes(o)
            self.index_note = "index: the GPU arm's yeast chr01 index"
        self.k = self.idx.k
        self.build_s = time.time() - t0

    def reads(self, n_reads, seed):
        import _gen
        if self.args.workload == "config5":
            return _gen.reads_from_packed(None, len(self.codes), n_reads, READ_LEN, seed, self.frac_ref, self.sub, codes=self.codes)
        return _gen.sample_reads_fast(self.codes, n_reads, READ_LEN, seed, self.frac_ref, self.sub)

    def run(self, bases, threads=None, want_hits=True):
        n = len(bases) // READ_LEN
        offs = np.arange(n + 1, dtype=np.uint64) * READ_LEN
        t0 = time.time()
        _, c, _ = self.idx.query_reads(bases, offs, streaming=self.args.mode == "streaming", want_hits=want_hits,
                                       n_threads=self.threads if threads is None else threads)
        return float(c[0]), time.time() - t0, c

    def calibrate(self, seconds, threads=None, cap=2_000_000):
        """reads per sample so that one pass takes about `seconds`."""
        cal = self.reads(20000, 4242)
        n, dt, _ = self.run(cal, threads, want_hits=False)
        rate = n / max(dt, 1e-6)
        return int(min(cap, max(20000, rate * seconds / (READ_LEN - self.k + 1))))


def bench_config(args):
    """`config` of the JSON line: identical for the GPU arm and the --impl reference arm."""
    cfg = {"workload": workload_name(args), "mode": args.mode, "read_len": READ_LEN, "units_per_gpu_per_step": args.reads}
    if args.workload.startswith("config5"):
        cfg["index_scale"] = args.scale
    return cfg


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The Rust crate cannot be built in
    this image (no cargo/rustc, four un-vendored crates), so this arm runs the C++ port (oracle/) with all
    host threads; each step is a bounded sample of the workload (same read mix; see CpuPort for the index)."""
    if rank != 0:
        return
    if args.workload in ("config4", "config5-kmers"):
        print(json.dumps({"impl": "reference", "unavailable": "--impl reference covers the read workloads (config1/2/3/5)"}), flush=True)
        return
    if args.workload == "config1":
        args.mode = "random"
    port = CpuPort(args)
    per_step_s = max(0.5, min(4.0, 100.0 / max(1, args.steps + args.warmup)))
    n_reads = min(args.reads, port.calibrate(per_step_s))
    bases = port.reads(n_reads, 42)
    for _ in range(args.warmup):
        port.run(bases[: READ_LEN * min(n_reads, 20000)])
    t0 = time.time()
    total = 0
    for _ in range(args.steps):
        n, _, c = port.run(bases)
        total += int(n)
    dt = time.time() - t0
    v = total / dt
    sample = "%d reads x %d bp (%d lookups) per step, %d steps; %s" % (n_reads, READ_LEN, n_reads * (READ_LEN - port.k + 1), args.steps, port.index_note)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": bench_config(args),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": port.threads, "kind": "port", "sample": sample,
                         "note": "C++ restatement of mazu's query path (oracle/); the Rust reference cannot be built in this image"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "counts": {"n_kmers": int(c[0]), "n_hit": int(c[1]), "n_miss": int(c[2])},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(args):
    return {
        "config1": "configs[0]: pufferfish yeast_chr01 DenseIndex (PFHash+BooPHF), all ref + unitig k-mers shuffled, k2u_batch",
        "config2": "configs[1]: SSHash(yeast chr01, k31 m15 skew32), random access, 150bp reads 50% ref / 50% random",
        "config3": "configs[2]: same SSHash via as_streaming(), 150bp reads 70% ref with 1% subs / 30% random",
        "config4": "configs[3]: u2pos decode, synthetic table (4096 refs, Zipf(1.2) multiplicity <= 65536)",
        "config5": "configs[4]: human-scale synthetic SSHash (k31 w19 skew64), 150bp reads 70% ref 1% subs / 30% random",
        "config5-kmers": "configs[4] index queried with a flat random-order k-mer batch (50% present), k2u_batch",
    }[args.workload]


def bind_to_gpu_numa_node(torch, local_rank):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host buffer is allocated
    (first touch places the pages): with 8 ranks on one host, staging through the far socket halves the copy rate."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        node = int(open(base + "/numa_node").read().strip())
        cpus = open(base + "/local_cpulist").read().strip()
        if node < 0 or not cpus:
            return {"numa_node": node, "bound": False}
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
            return {"numa_node": node, "bound": True, "n_cpus": len(ids)}
        return {"numa_node": node, "bound": False}
    except Exception as e:  # not fatal: the run is only slower
        return {"bound": False, "error": str(e)[:80]}


def measured_traffic(W, mode_name, scale):
    """DRAM bytes per unit of the dominant kernel from the committed ncu captures (profiles/traffic.json): the entry
    whose workload / mode match and whose index scale is the closest to this run's.  Returns (bytes_per_unit, label)."""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return None, None
    best = None
    for name, e in tj.items():
        if not isinstance(e, dict) or e.get("workload") != W:
            continue
        # same mode first; a capture of the other mode of the same workload is the next best estimate (the walk kernel issues the
        # same lookups) and is labelled as such
        d = (0 if e.get("mode", "random") == mode_name else 1, abs(float(e.get("scale", 1.0)) - scale))
        if best is None or d < best[0]:
            best = (d, name, e)
    if best is None:
        return None, None
    d, name, e = best
    per = e.get("dram_bytes_per_lookup", e.get("dram_bytes_per_unit"))
    label = "ncu --set full (%s): %d units per captured launch, index scale %s, %s%s" % (
        name, e.get("lookups_per_launch", 0), e.get("scale", "n/a"), e.get("_source", ""),
        "" if d[0] == 0 else "; captured in %s mode, used as the estimate for %s mode" % (e.get("mode", "random"), mode_name))
    global _TRAFFIC_ENTRY
    _TRAFFIC_ENTRY = e if d[0] == 0 else None
    return per, label


_TRAFFIC_ENTRY = None  # the traffic.json entry measured_traffic() settled on (same workload and mode), for the issue-side figures


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
    numa = bind_to_gpu_numa_node(torch, local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream()
    gen = torch.Generator(device=dev)
    gen.manual_seed(42 + rank)
    mode = mz.MODE_RANDOM if args.mode == "random" else mz.MODE_STREAMING
    W = args.workload
    frac_ref, sub_rate = mix_of(W, args.mode)
    info = {"host": {"cpus": os.cpu_count(), "numa_binding": numa}}
    unit, metric = UNIT, METRIC
    check = None          # callable -> bool : parity spot check (rank 0)
    e2e_step = None       # callable: one end-to-end step through host buffers (the headline e2e)
    e2e_extra = {}        # name -> (callable, units, h2d bytes, d2h bytes or None, api description): further end-to-end variants
    e2e_extra_post = {}   # name -> callable(dict): fills in what is only known after the variant has run
    e2e_bytes = (0, 0)
    e2e_units = None      # units per e2e step (defaults to n_units)
    e2e_api = None
    e2e_post = None       # callable(e2e dict): checks / extra fields after the e2e loop
    cpu_fn = None         # callable -> cpu_baseline dict
    t_build = time.time()

    # ---------------------------------------------------------------- indexes
    if W in ("config1", "config2", "config3"):
        dense = mz.DenseIndex.deserialize_from_cpp(YEAST, device=local_rank)
        index = dense if W == "config1" else dense.rebuild_k2u(mz.K2U_SSHASH, w=15, skew_param=32, seed=0)
        o, os_idx = load_oracle_yeast(sshash=W != "config1")
        ref_codes = yeast_ref_codes(o)
        alg = ALG_PFHASH if W == "config1" else ALG_SSHASH
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
    else:  # config4
        k = 31
        n_unitigs = 40_000  # Zipf(1.2) clipped to 65,536 has mean ~1,600: 40k unitigs give the ~6e7 occurrences SURVEY 8(d) sizes config 4 for
        rng = np.random.default_rng(44)
        # the K2U half is irrelevant for the decode: a trivial unitig set with n_unitigs entries backs the offsets vector
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
    info["index_build_s"] = round(time.time() - t_build, 2)
    info["index_device_bytes"] = index.device_bytes
    K = index.k
    nk_per_read = READ_LEN - K + 1

    # ---------------------------------------------------------------- inputs + step closures
    if W in ("config2", "config3", "config5"):
        n_reads = args.reads
        n_units = n_reads * nk_per_read
        if W == "config5":
            bases = _gen.device_reads_from_packed(torch, useq_dev, n_bases, n_reads, READ_LEN, gen, frac_ref, sub_rate)
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
        kernel_name = "mazu::query_reads_kernel<%d,SSHASH,CASCADE>" % (0 if mode == mz.MODE_RANDOM else 1)
        launches_per_step = 1

        def step():
            index.query_reads(bases, None, n_reads=n_reads, uniform_read_len=READ_LEN, mode=mode, out_hits=hits, counts=counts,
                              mem=mz.MEM_DEVICE, stream=stream.cuda_stream)

        if not args.no_e2e:
            # Host buffers are pinned.  The headline e2e is the hit-run call on the WHOLE batch at every N (150 B in, ~1.2 B per
            # lookup out: 10 M reads pin ~5 GB per rank).  The full-record / compact variants move 16 / 8 B per lookup
            # (19.2 GB for the whole batch): they run on the first --e2e-slice reads of the same batch, at every N alike.
            e2e_reads = n_reads
            e2e_units = e2e_reads * nk_per_read
            sl_reads = min(n_reads, args.e2e_slice)
            sl_units = sl_reads * nk_per_read
            h_bases = torch.empty(e2e_reads * READ_LEN, dtype=torch.uint8, pin_memory=True)
            h_bases.copy_(bases[: e2e_reads * READ_LEN])
            h_cnt = np.zeros(3, dtype=np.uint64)
            hb = h_bases.numpy()
            # hit runs: a lossless compact form of the records (codes + run starts); the buffers are reused across steps
            h_codes = torch.empty(e2e_units, dtype=torch.uint8, pin_memory=True)
            h_runs = torch.empty((max(1 << 20, e2e_units // 16), 4), dtype=torch.int32, pin_memory=True)
            h_rro = torch.zeros(e2e_reads + 1, dtype=torch.int64, pin_memory=True)
            runs_state = {"n_runs": 0}

            def e2e_ascii_step():
                n_runs = C.c_uint64(0)
                mz._check(mz.lib().mazu_b200_query_reads_runs(index._h, mz._any_ptr(hb), None, e2e_reads, READ_LEN, mode, None, mz._any_ptr(h_codes),
                                                              mz._any_ptr(h_runs), h_runs.shape[0], mz._any_ptr(h_rro), C.byref(n_runs), mz._np_ptr(h_cnt)))
                runs_state["n_runs"] = n_runs.value

            ascii_api = ("mazu_b200_query_reads_runs (C ABI, host buffers): pinned ASCII reads in; out: one code byte per k-mer slot + the 16-byte "
                         "record of every run start + per-read run offsets (lossless: mazu_b200_expand_hit_runs rebuilds every record)")

            def ascii_post(e):
                e["d2h_bytes_per_step"] = (e2e_units + 16 * runs_state["n_runs"] + 8 * (e2e_reads + 1) + 24) * world
                e["n_runs_per_step_per_gpu"] = runs_state["n_runs"]
                m = min(e2e_reads, 20000)  # the expansion on the host reproduces the full records (checked on the head of the batch)
                exp = mz.ModIndex.expand_hit_runs(h_codes.numpy()[: m * nk_per_read], h_runs.numpy().view(np.uint32).reshape(-1).view(mz.HIT_DTYPE),
                                                  h_rro.numpy().view(np.uint64)[: m + 1], uniform_slots=nk_per_read)
                e["expands_to_full_records"] = bool(np.array_equal(exp.view(np.uint32).reshape(-1, 4), hits[: m * nk_per_read].cpu().numpy().view(np.uint32)))
                e["counts_match_device_path"] = bool(int(h_cnt[0]) == e2e_units and int(h_cnt[1]) > 0)

            h_hits = torch.empty((sl_units, 4), dtype=torch.int32, pin_memory=True)
            hb_sl = hb[: sl_reads * READ_LEN]

            def e2e_full_step():
                index.query_reads(hb_sl, None, n_reads=sl_reads, uniform_read_len=READ_LEN, mode=mode, out_hits=h_hits, counts=h_cnt)

            def e2e_compact_step():  # the 8-byte records reuse the first half of the same pinned buffer
                index.query_reads(hb_sl, None, n_reads=sl_reads, uniform_read_len=READ_LEN, mode=mode, out_hits=h_hits, counts=h_cnt, compact=True)

            def e2e_devout_step():  # reads from pinned host memory, records left in HBM for the next device stage, counters back
                index.query_reads(hb, None, n_reads=e2e_reads, uniform_read_len=READ_LEN, mode=mode, out_hits=hits, counts=h_cnt,
                                  mem=mz.MEM_HOST_IN_DEVICE_OUT)

            # The headline e2e: the packed run interface -- reads held 2-bit packed by the caller (mazu's own SeqVector
            # representation), 2-bit run codes + run records + per-read run offsets back.  0.31 + 0.5 bytes per lookup cross PCIe,
            # which is what lets 8 GPUs behind one host scale (profiles/r02i_pcie_8gpu.json: the host moves 92 GB/s of D2H in total).
            # Packing ASCII -> 2-bit (mazu_b200_pack_reads, host threads) is done once, outside the timed loop, and its rate is
            # reported; the ASCII-in call is measured next to it (`ascii_reads`).
            packed_ok = nk_per_read % 4 == 0
            e2e_extra = {}
            e2e_extra_post = {}
            if packed_ok:
                wpr = (READ_LEN + 31) // 32
                h_words = torch.empty(e2e_reads * wpr, dtype=torch.int64, pin_memory=True)
                t_pack = time.perf_counter()
                mz.pack_reads(hb, READ_LEN, words=h_words, want_mask=False)
                t_pack = time.perf_counter() - t_pack
                h_codes2 = torch.empty((e2e_units + 3) // 4, dtype=torch.uint8, pin_memory=True)

                def e2e_step():
                    n_runs = C.c_uint64(0)
                    mz._check(mz.lib().mazu_b200_query_reads_runs_packed(index._h, mz._any_ptr(h_words), None, e2e_reads, READ_LEN, mode, mz._any_ptr(h_codes2),
                                                                         mz._any_ptr(h_runs), h_runs.shape[0], mz._any_ptr(h_rro), C.byref(n_runs), mz._np_ptr(h_cnt)))
                    runs_state["n_runs_packed"] = n_runs.value

                e2e_api = ("mazu_b200_query_reads_runs_packed (C ABI, host buffers): pinned 2-bit packed reads in (40 B per 150 bp read); out: 2-bit run code "
                           "per k-mer slot + the 16-byte record of every run start + per-read run offsets (lossless: mazu_b200_expand_hit_runs_packed "
                           "rebuilds every record)")

                def e2e_post(e):
                    e["h2d_bytes_per_step"] = e2e_reads * wpr * 8 * world
                    e["d2h_bytes_per_step"] = ((e2e_units + 3) // 4 + 16 * runs_state["n_runs_packed"] + 8 * (e2e_reads + 1) + 24) * world
                    e["n_runs_per_step_per_gpu"] = runs_state["n_runs_packed"]
                    e["input"] = ("2-bit packed reads; packing the ASCII batch with mazu_b200_pack_reads (host threads) took %.0f ms = %.1f GB/s of ASCII, "
                                  "outside the timed loop" % (t_pack * 1e3, e2e_reads * READ_LEN / t_pack / 1e9))
                    m = min(e2e_reads, 20000)  # the expansion on the host reproduces the full records (checked on the head of the batch)
                    exp = mz.ModIndex.expand_hit_runs_packed(h_codes2.numpy()[: m * nk_per_read // 4], h_runs.numpy().view(np.uint32).reshape(-1).view(mz.HIT_DTYPE),
                                                             h_rro.numpy().view(np.uint64)[: m + 1], nk_per_read)
                    e["expands_to_full_records"] = bool(np.array_equal(exp.view(np.uint32).reshape(-1, 4), hits[: m * nk_per_read].cpu().numpy().view(np.uint32)))
                    e["counts_match_device_path"] = bool(int(h_cnt[0]) == e2e_units and int(h_cnt[1]) > 0)

                e2e_extra["ascii_reads"] = (e2e_ascii_step, e2e_units, e2e_reads * READ_LEN, None, ascii_api)
                # The sparsest lossless result: one 16-byte record per hit RUN and nothing per slot or per read
                # (mazu_b200_query_reads_intervals_packed).  It is the headline where it applies (short reads, random access or
                # unique k-mers); the run-code interface above is then reported next to it as `packed_runs`.
                if nk_per_read <= 128 and (mode == mz.MODE_RANDOM or index.kmers_unique):
                    h_iv = torch.empty((max(1 << 20, e2e_units // 16), 4), dtype=torch.int32, pin_memory=True)
                    runs_step, runs_api, runs_post = e2e_step, e2e_api, e2e_post

                    def e2e_step():
                        n_iv = C.c_uint64(0)
                        mz._check(mz.lib().mazu_b200_query_reads_intervals_packed(index._h, mz._any_ptr(h_words), None, e2e_reads, READ_LEN, mode,
                                                                                  mz._any_ptr(h_iv), h_iv.shape[0], C.byref(n_iv), mz._np_ptr(h_cnt)))
                        runs_state["n_iv"] = n_iv.value

                    e2e_api = ("mazu_b200_query_reads_intervals_packed (C ABI, host buffers): pinned 2-bit packed reads in (40 B per 150 bp read); out: "
                               "one 16-byte record {unitig, pos|orientation, read, first slot, length} per hit run, nothing per slot or per read "
                               "(lossless: mazu_b200_expand_hit_intervals rebuilds every record)")

                    def e2e_post(e):
                        e["h2d_bytes_per_step"] = e2e_reads * wpr * 8 * world
                        e["d2h_bytes_per_step"] = (16 * runs_state["n_iv"] + 32) * world
                        e["n_runs_per_step_per_gpu"] = runs_state["n_iv"]
                        e["input"] = ("2-bit packed reads; packing the ASCII batch with mazu_b200_pack_reads (host threads) took %.0f ms = %.1f GB/s of ASCII, "
                                      "outside the timed loop" % (t_pack * 1e3, e2e_reads * READ_LEN / t_pack / 1e9))
                        m = min(e2e_reads, 20000)  # the expansion on the host reproduces the full records (checked on the head of the batch)
                        iv = h_iv.numpy().view(np.uint32).reshape(-1).view(mz.INTERVAL_DTYPE)[: runs_state["n_iv"]]
                        exp = index.expand_hit_intervals(np.ascontiguousarray(iv[iv["read"] < m]), None, m, READ_LEN)
                        e["expands_to_full_records"] = bool(np.array_equal(exp.view(np.uint32).reshape(-1, 4), hits[: m * nk_per_read].cpu().numpy().view(np.uint32)))
                        e["counts_match_device_path"] = bool(int(h_cnt[0]) == e2e_units and int(h_cnt[1]) > 0)

                    e2e_extra["packed_runs"] = (runs_step, e2e_units, e2e_reads * wpr * 8, None, runs_api)
                    e2e_extra_post = {"packed_runs": runs_post}

                    # ... and the same records from the reads as the reference takes them: ASCII in (1.25 B per lookup), nothing packed
                    # by the caller, nothing outside the timed loop
                    def e2e_ascii_iv_step():
                        n_iv = C.c_uint64(0)
                        mz._check(mz.lib().mazu_b200_query_reads_intervals(index._h, mz._any_ptr(hb), e2e_reads, READ_LEN, mode, mz._any_ptr(h_iv),
                                                                           h_iv.shape[0], C.byref(n_iv), mz._np_ptr(h_cnt)))
                        runs_state["n_iv_ascii"] = n_iv.value

                    def ascii_iv_post(e):
                        e["d2h_bytes_per_step"] = (16 * runs_state["n_iv_ascii"] + 32) * world
                        e["n_runs_per_step_per_gpu"] = runs_state["n_iv_ascii"]
                        m = min(e2e_reads, 20000)
                        iv = h_iv.numpy().view(np.uint32).reshape(-1).view(mz.INTERVAL_DTYPE)[: runs_state["n_iv_ascii"]]
                        exp = index.expand_hit_intervals_ascii(np.ascontiguousarray(iv[iv["read"] < m]), hb[: m * READ_LEN], m, READ_LEN)
                        e["expands_to_full_records"] = bool(np.array_equal(exp.view(np.uint32).reshape(-1, 4), hits[: m * nk_per_read].cpu().numpy().view(np.uint32)))
                        e["counts_match_device_path"] = bool(int(h_cnt[0]) == e2e_units and int(h_cnt[1]) > 0)

                    e2e_extra["ascii_intervals"] = (e2e_ascii_iv_step, e2e_units, e2e_reads * READ_LEN, None,
                                                    "mazu_b200_query_reads_intervals (C ABI, host buffers): pinned ASCII reads in, as the reference takes "
                                                    "them (150 B per read, no packing by the caller); out: one 16-byte record per hit run (lossless: "
                                                    "mazu_b200_expand_hit_intervals_ascii rebuilds every record)")
                    e2e_extra_post["ascii_intervals"] = ascii_iv_post
            else:  # read lengths whose slot count is not a multiple of 4: the byte-coded run call is the headline
                e2e_step, e2e_api = e2e_ascii_step, ascii_api

                def e2e_post(e):
                    e["h2d_bytes_per_step"] = e2e_reads * READ_LEN * world
                    ascii_post(e)
            e2e_extra.update({
                "full_records": (e2e_full_step, sl_units, sl_reads * READ_LEN, sl_units * 16 + 24,
                                 "mazu_b200_query_reads(MAZU_MEM_HOST): every 16-byte record over PCIe; first %d reads of the batch" % sl_reads),
                "compact_records": (e2e_compact_step, sl_units, sl_reads * READ_LEN, sl_units * 8 + 24,
                                    "mazu_b200_query_reads_compact: 8-byte {unitig_id, pos|match<<30} records; first %d reads of the batch" % sl_reads),
                "host_reads_in_device_records_out": (e2e_devout_step, e2e_units, e2e_reads * READ_LEN, 24,
                                                     "MAZU_MEM_HOST_IN_DEVICE_OUT: pinned host reads in, 16-byte records stay in HBM (input of project_hits on the "
                                                     "device), the three counters of `kphf bench` (src/bin/kphf/main.rs:282-284) come back"),
            })

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
                # the hit fraction must match the generator: 70 % of reads x 0.99^31 surviving k-mers x the ~70 % of a read's windows
                # that do not straddle a unitig boundary (unitigs of mean length 98) = 35.5 %; completeness is checked exactly
                # against the CPU port on its own index (cpu_baseline.gpu_path_equals_port_on_its_index)
                frac = len(slot) / float(len(h))
                ok &= 0.30 < frac < 0.42
                return ok and len(slot) > 0
            n_chk = min(n_reads, 5000)
            chk = bases[: n_chk * READ_LEN].cpu().numpy()
            want, _, _ = os_idx.query_reads(chk, np.arange(n_chk + 1, dtype=np.uint64) * READ_LEN, streaming=(mode == mz.MODE_STREAMING))
            got = hits[: n_chk * nk_per_read].cpu().numpy().view(np.uint32).reshape(-1).view(mz.HIT_DTYPE)
            return bool(np.array_equal(got, want))

        def chain_line():
            """GetRefPos::get_ref_pos over reads (src/index.rs:156-216), everything on the device: the fused kernel
            (mazu_b200_get_ref_pos_reads) against the unfused chain (query_reads -> project_hits) on the same reads."""
            m = min(n_reads, 4_000_000)
            ns = m * nk_per_read
            d_offs = torch.zeros(ns + 1, dtype=torch.int64, device=dev)
            mz._check(mz.lib().mazu_b200_project_hits(index._h, mz._any_ptr(hits), ns, mz._any_ptr(d_offs), None, 0, None, mz.MEM_DEVICE, mz._any_ptr(stream.cuda_stream)))
            torch.cuda.synchronize()
            total = int(d_offs[ns].item())
            d_out = torch.empty((total + 16, 3), dtype=torch.int32, device=dev)
            d_out2 = torch.empty((total + 16, 3), dtype=torch.int32, device=dev)
            d_offs2 = torch.zeros(ns + 1, dtype=torch.int64, device=dev)
            sp = mz._any_ptr(stream.cuda_stream)

            def unfused():
                index.query_reads(bases, None, n_reads=m, uniform_read_len=READ_LEN, mode=mode, out_hits=hits, counts=None, mem=mz.MEM_DEVICE, stream=stream.cuda_stream)
                mz._check(mz.lib().mazu_b200_project_hits(index._h, mz._any_ptr(hits), ns, mz._any_ptr(d_offs), mz._any_ptr(d_out), total, None, mz.MEM_DEVICE, sp))

            def project_only():
                mz._check(mz.lib().mazu_b200_project_hits(index._h, mz._any_ptr(hits), ns, mz._any_ptr(d_offs), mz._any_ptr(d_out), total, None, mz.MEM_DEVICE, sp))

            def fused():
                mz._check(mz.lib().mazu_b200_get_ref_pos_reads(index._h, mz._any_ptr(bases), None, m, READ_LEN, mode, ns, None, None, mz._any_ptr(d_offs2),
                                                               mz._any_ptr(d_out2), total, None, None, mz.MEM_DEVICE, sp))

            def timed(fn):
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for _ in range(args.steps):
                    fn()
                b.record(stream)
                torch.cuda.synchronize()
                return a.elapsed_time(b) / args.steps

            t_un, t_pr, t_fu = timed(unfused), timed(project_only), timed(fused)
            same = bool(torch.equal(d_offs, d_offs2)) and bool(torch.equal(d_out[:total], d_out2[:total]))
            return {"reads": m, "kmer_slots": ns, "projected_positions": total,
                    "unfused": {"ms": t_un, "kmers_per_s": ns / (t_un * 1e-3), "project_hits_ms": t_pr, "api": "mazu_b200_query_reads + mazu_b200_project_hits"},
                    "fused": {"ms": t_fu, "kmers_per_s": ns / (t_fu * 1e-3), "api": "mazu_b200_get_ref_pos_reads (lookups + per-tile totals, scan over tiles, tile-wise emit)"},
                    "fused_equals_unfused": same}

        if W in ("config2", "config3") and not args.no_chain:
            info["chain_fn"] = chain_line

        def cpu_fn():
            port = CpuPort(args)
            n_s = port.calibrate(args.cpu_seconds)
            sample = port.reads(n_s, 4243)
            n, dt, c = port.run(sample)
            n1_s = max(20000, n_s // port.threads)
            n1, dt1, _ = port.run(sample[: n1_s * READ_LEN], threads=1)
            out = {"value": n / dt, "unit": UNIT, "cores": port.threads, "kind": "port",
                   "sample": "%d reads x %d bp (%d lookups) of the same read mix, %.1f s; %s" % (n_s, READ_LEN, int(c[0]), dt, port.index_note),
                   "single_thread_value": n1 / dt1, "ns_per_kmer_single_thread": 1e9 / (n1 / dt1),
                   "hit_fraction": float(c[1]) / max(1.0, float(c[0])),
                   "note": "C++ restatement of mazu's query path (oracle/); the Rust reference cannot be built in this image"}
            if W == "config5":  # the port against the GPU path on the port's own (smaller) index: same reads, same answers
                small = _gen.synthetic_unitigs_packed(max(1000, int(HUMAN_UNITIGS * port.scale)), 68, 31, seed=45)
                g_small = mz.SSHash.from_unitig_set(mz.UnitigSet(31, *small), 19, 64, seed=0, device=local_rank, builder="gpu")
                m = min(n_s, 20000)
                got, _, _ = g_small.query_reads(sample[: m * READ_LEN], None, uniform_read_len=READ_LEN, mode=mode)
                want, _, _ = port.idx.query_reads(sample[: m * READ_LEN], np.arange(m + 1, dtype=np.uint64) * READ_LEN, streaming=(mode == mz.MODE_STREAMING))
                out["gpu_path_equals_port_on_its_index"] = bool(np.array_equal(got, want))
            return out

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
            kernel_name = "mazu::k2u_batch_pfhash_kernel<BOOPHF>"
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
            e2e_units = min(n_units, args.e2e_slice * nk_per_read)
            h_k = torch.empty(e2e_units, dtype=torch.int64, pin_memory=True)
            h_k.copy_(kmers[:e2e_units])
            h_hits = torch.empty((e2e_units, 4), dtype=torch.int32, pin_memory=True)
            hk = h_k.numpy().view(np.uint64)

            def e2e_step():
                index.k2u_batch(hk, out=h_hits)

            e2e_bytes = (e2e_units * 8, e2e_units * 16)
            e2e_api = "mazu_b200_k2u_batch(MAZU_MEM_HOST): pinned k-mer words in, every 16-byte record out; first %d k-mers of the batch" % e2e_units

            def e2e_post(e):
                e["matches_device_path"] = bool(torch.equal(h_hits[:1_000_000], hits[:1_000_000].cpu()))

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

        def cpu_fn():
            """the port's U2Pos decode (decode_unitig_occs, dense_unitig_table.rs:127-153) on the same table, all host threads"""
            import _oracle
            threads = os.cpu_count() or 1
            o4 = _oracle.OracleIndex.from_packed(k, words4, int(accum4[-1]), accum4, 0)
            o4.attach_u2pos(1, offsets, ref_ids, poss, fws, max_ref_len, n_refs)
            # bounded sample: queries whose lists sum to ~1.5e8 occurrences per thread-second budget
            budget = int(2.5e7 * args.cpu_seconds * threads / 4)
            n_s = int(max(1, min(n_q, np.searchsorted(cum, budget))))
            parts = np.array_split(qids[:n_s], threads)
            done = [0] * threads

            def work(t):
                offs_t, occ_t = o4.decode_occs(parts[t])
                done[t] = len(occ_t)

            t0 = time.time()
            passes, total_done = 0, 0
            while passes == 0 or time.time() - t0 < args.cpu_seconds / 2:  # repeat the sample until the clock has something to measure
                ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
                for th in ths:
                    th.start()
                for th in ths:
                    th.join()
                passes += 1
                total_done += sum(done)
            dt = time.time() - t0
            return {"value": total_done / dt, "unit": unit, "cores": threads, "kind": "port",
                    "sample": "%d of the same queries (%d occurrences) x %d passes, %.1f s" % (n_s, sum(done), passes, dt),
                    "note": "C++ restatement of PiscemUnitigTable::decode_unitig_occs (oracle/), one slice of the queries per host thread"}

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

    def timed_host_loop(fn):
        """K end-to-end steps through host buffers between barriers: wall clock, max over ranks."""
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    e2e = None
    if e2e_step is not None:
        dt = timed_host_loop(e2e_step)
        eu = e2e_units if e2e_units is not None else n_units
        e2e = {"value": float(eu) * args.steps * world / dt, "unit": unit, "units_per_step_per_gpu": eu,
               "h2d_bytes_per_step": e2e_bytes[0] * world, "d2h_bytes_per_step": e2e_bytes[1] * world, "steps": args.steps, "api": e2e_api}
        if e2e_post is not None:
            e2e_post(e2e)
        e2e["achieved_h2d_gbs_total"] = e2e["h2d_bytes_per_step"] * args.steps / dt / 1e9
        e2e["achieved_d2h_gbs_total"] = e2e["d2h_bytes_per_step"] * args.steps / dt / 1e9
        try:  # what the HOST can move with all N GPUs copying at once (profiles/pcie_probe_multi.py, measured on this pool's boxes)
            src = {1: "r01_pcie.json", 2: "r02h_pcie_2gpu.json", 8: "r02i_pcie_8gpu.json"}.get(world)
            if src:
                hc = json.load(open(os.path.join(ROOT, "profiles", src)))
                e2e["host_copy_ceiling"] = dict(hc, source="profiles/" + src)
        except Exception:
            pass
        pc = None
        try:  # the box's pinned-copy bandwidth (profiles/pcie_probe.py): what bounds the variants that move every record
            pc = json.load(open(os.path.join(ROOT, "profiles", "r01_pcie.json")))
        except Exception:
            pass
        for name, (fn, units, h2d, d2h, api) in e2e_extra.items():
            dtx = timed_host_loop(fn)
            e2e[name] = {"value": float(units) * args.steps * world / dtx, "unit": unit, "units_per_step_per_gpu": units,
                         "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": (d2h or 0) * world, "api": api}
            if name == "ascii_reads":
                ascii_post(e2e[name])
            if name in e2e_extra_post:
                e2e_extra_post[name](e2e[name])
            if name == "full_records":
                e2e[name]["matches_device_path"] = bool(torch.equal(h_hits[:1_000_000], hits[:1_000_000].cpu()))
                if pc:
                    d2h_gbs = d2h * args.steps / dtx / 1e9
                    e2e[name]["bound"] = {"kind": "pcie d2h (16 B per lookup out)", "achieved_d2h_gbs_per_gpu": d2h_gbs, "measured_d2h_peak_gbs": pc["d2h_gbs"],
                                          "measured_d2h_with_h2d_busy_gbs": pc["duplex_each_gbs"], "frac_of_d2h_peak": d2h_gbs / pc["d2h_gbs"]}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    parity_ok = bool(check()) if check else None
    if "projected_fn" in info:
        info["projected"] = info.pop("projected_fn")()
    if "chain_fn" in info:
        info["get_ref_pos_chain"] = info.pop("chain_fn")()
    if args.validate and W.startswith("config5"):
        c = index.k2u_validate_self()
        info["k2u_validate_self"] = {"n_queries": c[0], "n_identity": c[1], "n_twin": c[2], "n_fail": c[4], "n_fail_not_found": c[3],
                                     "note": "failures that are not 'not found' are duplicated canonical k-mers of the random unitig set (expected ~1.3 pairs at 2.5e9 k-mers)"}
    cpu = cpu_fn() if (cpu_fn and not args.no_cpu_baseline and world == 1) else None

    # ---------------------------------------------------------------- roofline
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_source = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"
    alg_gbs = n_units * alg / (kernel_ms * 1e-3) / 1e9
    per_unit, traffic_label = measured_traffic(W, args.mode, args.scale if W.startswith("config5") else 1.0)
    traffic = per_unit * n_units if per_unit is not None else None
    index_in_l2 = index.device_bytes < (100 << 20)
    if W == "config4":
        roof = {"bound": "hbm", "achieved": alg_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": alg_gbs / hbm_peak, "traffic": traffic,
                "peak_source": hbm_source, "algorithmic_bytes_per_unit": alg,
                "note": "streaming decode: algorithmic bytes = 32 B (offset pair) per queried unitig + 6 B packed word in + 12 B record out per occurrence"}
    elif index_in_l2:
        # configs[0..2]: the index (0.4 MB) lives in L1/L2, DRAM only carries the compulsory stream (read bases in, records out),
        # so the HBM roofline of these launches is that stream; the kernel itself is issue-bound
        stream_b = 17.25 if W != "config1" else 24.0
        ach = n_units * stream_b / (kernel_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": traffic, "peak_source": hbm_source,
                "algorithmic_bytes_per_unit": stream_b,
                "note": "index is L2-resident: the only bytes that must cross DRAM are the stream (%.2f B per lookup: read bases in + 16-byte record out); "
                        "the SURVEY 8(d) ideal-layout figure (%.2f B, every structure touch a DRAM sector) is reported as `survey_8d` and is NOT an HBM "
                        "fraction here; the kernel is issue-bound (ncu: profiles/*prof_random.summary.txt)" % (stream_b, alg),
                "survey_8d": {"bytes_per_unit": alg, "gbs": alg_gbs, "ratio_to_hbm_peak": alg_gbs / hbm_peak}}
    else:
        # configs[4]: random-access regime.  A random access moves one 128-byte DRAM line whatever it uses of it (ncu on the probe),
        # so the ceiling is the measured random LINE rate P_rand, and what the kernel achieves is the DRAM bytes it really moves.
        try:
            pr = json.load(open(os.path.join(ROOT, "profiles", "r02_prand.json")))
            p_lines, p_src = float(pr["lines_per_s"]), "profiles/r02_prand.json (max over the granule / loads-in-flight / occupancy sweep, %s GiB table)" % pr.get("table_gib")
        except Exception:
            pr = json.load(open(os.path.join(ROOT, "profiles", "r01_prand.json")))
            p_lines, p_src = float(pr["32GiB"]["sectors_per_s"]), "profiles/r01_prand.json (round-1 probe: one configuration only)"
        peak = p_lines * 128 / 1e9
        if per_unit is not None:
            ach = per_unit * n_units / (kernel_ms * 1e-3) / 1e9
            how = "DRAM bytes per lookup measured by ncu x lookups / CUDA-event kernel time"
        else:  # never report the ideal-layout figure as a DRAM rate: it can exceed the peak
            ach, how = None, "no ncu capture for this workload: see survey_8d for the algorithmic figure"
        roof = {"bound": "hbm", "regime": "random access (128-byte DRAM lines)", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": (ach / peak) if ach is not None else None,
                "traffic": traffic, "peak_source": "P_rand = %.4g random 128-byte lines/s x 128 B: %s" % (p_lines, p_src), "achieved_source": how,
                "dram_lines_per_unit": (per_unit / 128.0) if per_unit is not None else None, "frac_of_hbm_copy_peak": (ach / hbm_peak) if ach is not None else None,
                "survey_8d": {"bytes_per_unit": alg, "gbs": alg_gbs, "ratio_to_hbm_peak": alg_gbs / hbm_peak, "ratio_to_p_rand": alg_gbs / peak,
                              "note": "ideal-layout figure with nothing shared between the k-mers of a read: a label, not a DRAM measurement"}}
    roof.update({"traffic_source": traffic_label, "units_per_launch": n_units, "kernel_ms": kernel_ms, "step_ms": [round(x, 3) for x in step_ms]})
    if _TRAFFIC_ENTRY and _TRAFFIC_ENTRY.get("ipc"):
        # what actually bounds the lookup kernels once the DRAM traffic is this low: instruction issue (from the same ncu capture)
        te = _TRAFFIC_ENTRY
        roof["issue"] = {"ipc": te["ipc"], "peak_ipc": 4.0, "frac": te["ipc"] / 4.0, "issue_active_pct": te.get("issue_active_pct"),
                         "warp_instructions_per_unit": te.get("warp_inst_per_lookup"), "threads_per_instruction": te.get("threads_per_inst"),
                         "source": "same ncu capture as `traffic` (sm__inst_executed.avg.per_cycle_elapsed, smsp__issue_active)"}

    cfg = bench_config(args)
    line = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic (torch.Generator seed 42+rank on the device; unitig set: numpy seed 45 / fixture sequences from tests/data)",
        "config": cfg,
        "l2_policy": "inputs + results per step (%.2f GB) are larger than the 126 MB L2%s" %
                     (n_units * 17.25 / 1e9, "" if index_in_l2 else "; so is the index (%.1f GB)" % (index.device_bytes / 1e9)),
        "parallelism": "index replicated x%d, inputs sharded by batch, no collective" % world,
        "clocks": clk.summary(),
        "gpu_launches": args.steps * launches_per_step,
        "kernel": kernel_name,
        "roofline": roof,
        "parity_spot_check": parity_ok,
    }
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
