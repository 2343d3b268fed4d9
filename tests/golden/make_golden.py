#!/usr/bin/env python3
"""Regenerates tests/golden/*.npz: answers of the CPU oracle (oracle/, itself pinned to the reference's own golden
vectors by tests/test_oracle_golden.py) on small seeded inputs.  The GPU parity tests compare the CUDA path against
these files WITHOUT running the oracle, and a CPU test checks that the oracle still reproduces them, so a drift of
either side shows up against a committed artefact.

    python tests/golden/make_golden.py          # rewrites the fixtures in place

Inputs are derived from the pufferfish-generated fixtures under tests/data (the reference's own test data) with
numpy's default_rng and fixed seeds; nothing here reads /root/reference."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _gen  # noqa: E402
from _oracle import OracleIndex  # noqa: E402

YEAST = os.path.join(os.path.dirname(HERE), "data", "pf1", "yeast_chr01_index")


def inputs():
    """the seeded query batch shared by make_golden and the tests: ragged reads with substitutions, Ns and lower case"""
    o = OracleIndex.dense_from_pf1(YEAST)
    ref_codes = _gen.unpack_2bit(o.refseq_words(), int(o.ref_prefix()[-1]))
    bases, offs = _gen.sample_reads(ref_codes, 700, 220, seed=20261018, frac_ref=0.7, sub_rate=0.01, n_rate=0.002, ragged=True)
    return o, bases, offs


def main():
    o, bases, offs = inputs()
    idx = {"pfhash": o, "sshash": o.rebuild_k2u(1, w=15, skew=32, seed=0)}
    out = {"bases": bases, "read_offsets": offs, "validate_self": np.array(o.validate_self(), dtype=np.uint64)}
    for name, ix in idx.items():
        for streaming in (False, True):
            hits, cnt, koffs = ix.query_reads(bases, offs, streaming=streaming, reset_per_read=True)
            tag = "%s_%s" % (name, "streaming" if streaming else "random")
            out[tag + "_hits"] = hits.view(np.uint32).reshape(-1, 4)
            out[tag + "_counts"] = np.asarray(cnt, dtype=np.uint64)
            out["kmer_offsets"] = np.asarray(koffs, dtype=np.uint64)
    hits = idx["sshash"].query_reads(bases, offs)[0]
    po, pr = idx["sshash"].project_hits(hits)
    out["project_offsets"] = np.asarray(po, dtype=np.uint64)
    out["project_records"] = np.asarray(pr).view(np.uint32).reshape(-1, 3)
    uids = np.arange(0, o.n_unitigs, 7, dtype=np.uint32)
    do, dr = o.decode_occs(uids)
    out["decode_unitig_ids"] = uids
    out["decode_offsets"] = np.asarray(do, dtype=np.uint64)
    out["decode_records"] = np.asarray(dr).view(np.uint32).reshape(-1, 3)
    path = os.path.join(HERE, "yeast_chr01_queries.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
