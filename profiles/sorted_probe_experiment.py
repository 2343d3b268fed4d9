#!/usr/bin/env python3
"""Upper bound of what "sorted probe batches" could gain for flat k-mer batches on an index that does not fit the L2.
The batch is sorted by the level-0 MPHF block of each query's key (mazu_b200_debug_probe_key), which makes the MPHF,
bucket-bound and positions accesses of neighbouring queries share DRAM lines; the k2u_batch kernel is then timed on the
sorted and on the original batch.  The sort itself and the permutation of the inputs / results are timed separately
(torch.sort + gathers): they are what a real implementation would have to pay.
usage: sorted_probe_experiment.py [scale] -> one JSON line"""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _gen
import mazu_b200 as mz
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
dev = torch.device("cuda", 0)
n_unitigs = int(36_145_130 * scale)
words, n_bases, accum = _gen.synthetic_unitigs_packed(n_unitigs, 68, 31, seed=45)
index = mz.SSHash.from_unitig_set(mz.UnitigSet(31, words, n_bases, accum), 19, 64, seed=0, device=0, builder="gpu")
useq_dev = torch.from_numpy(np.concatenate([words, np.zeros(2, dtype=np.uint64)]).view(np.int64)).to(dev)
gen = torch.Generator(device=dev); gen.manual_seed(7)
n = 480_000_000
kmers = _gen.device_kmers_from_packed(torch, useq_dev, n_bases, n, 31, gen, 0.5)
hits = torch.empty((n, 4), dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
t_plain = timed(lambda: index.k2u_batch(kmers, out=hits, mem=mz.MEM_DEVICE, stream=st, n=n))
ref = hits.clone()
keys = torch.empty(n, dtype=torch.int32, device=dev)
t_key = timed(lambda: mz._check(mz.lib().mazu_b200_debug_probe_key(index._h, mz._any_ptr(kmers), n, mz._any_ptr(keys), mz._any_ptr(st))))
t0 = time.perf_counter(); order = torch.argsort(keys); torch.cuda.synchronize(); t_sort = (time.perf_counter() - t0) * 1e3
t_gather = timed(lambda: kmers.index_select(0, order))
sk = kmers.index_select(0, order)
t_sorted = timed(lambda: index.k2u_batch(sk, out=hits, mem=mz.MEM_DEVICE, stream=st, n=n))
out = torch.empty_like(hits)
t_scatter = timed(lambda: out.index_copy_(0, order, hits))
ok = bool(torch.equal(out, ref))
print(json.dumps({"scale": scale, "n_queries": n, "index_bytes": index.device_bytes, "kernel_ms_unsorted": t_plain, "kernel_ms_sorted_input": t_sorted,
                  "key_ms": t_key, "argsort_ms_torch": t_sort, "gather_inputs_ms": t_gather, "scatter_results_ms": t_scatter,
                  "total_ms_sorted_pipeline": t_key + t_sort + t_gather + t_sorted + t_scatter, "results_identical": ok}))
