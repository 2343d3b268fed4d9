"""N>1 host logic on CPU (gloo, world_size 2): read sharding, per-rank seeds, max-over-ranks timing
and the final counter reduction of bench.py's multi-GPU path.  No GPU, no compute kernels."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
    import _gen, _oracle
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # every rank holds the same (replicated) index; reads are sharded by contiguous batch
    o = _oracle.OracleIndex.dense_from_pf1(os.path.join(%r, "tests/data/pf1/yeast_chr01_index"))
    ref = _gen.unpack_2bit(o.refseq_words(), int(o.ref_prefix()[-1]))
    n_total, read_len = 4000, 150
    bases = _gen.sample_reads_fast(ref, n_total, read_len, seed=42, frac_ref=0.5)
    lo, hi = n_total * rank // world, n_total * (rank + 1) // world
    shard = bases[lo * read_len: hi * read_len]
    offs = np.arange(hi - lo + 1, dtype=np.uint64) * read_len
    hits, cnt, _ = o.query_reads(shard, offs)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)      # stand-in for per-rank elapsed ms
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    c = torch.tensor(cnt.astype(np.int64))
    dist.all_reduce(c, op=dist.ReduceOp.SUM)                        # final gather of per-shard hit counts
    digest = torch.tensor([int(np.bitwise_xor.reduce(hits.view(np.uint32).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(1))], dtype=torch.int64)
    gathered = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(gathered, digest)
    if rank == 0:
        full, fcnt, _ = o.query_reads(bases, np.arange(n_total + 1, dtype=np.uint64) * read_len)
        print(json.dumps({"max_ms": t.item(), "counts": c.tolist(), "full_counts": [int(x) for x in fcnt], "world": world,
                          "n_digests": len(gathered)}))
    dist.barrier()
    dist.destroy_process_group()
""")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sharded_counts_match_unsharded(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % (ROOT, ROOT, ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    import json
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == 2 and out["max_ms"] == 2.0 and out["n_digests"] == 2
    assert out["counts"] == out["full_counts"]  # sum over shards == unsharded run


def test_bench_reference_arm_only_rank0_prints(tmp_path):
    """--impl reference under torchrun: rank 0 alone runs and prints, the other rank exits 0 silently."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "1", "--reads", "20000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    import json
    out = json.loads(lines[0])
    assert out["impl"] == "reference" and out["cpu_baseline"]["kind"] == "port" and out["value"] > 0
    assert out["e2e"]["h2d_bytes_per_step"] == 0
