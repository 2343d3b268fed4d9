"""CPU-only checks of the product: the C-ABI library loads and exports every symbol the header
declares, the header and the ctypes binding agree, and the host builders / device layouts are
self-consistent (tests/host_check.cpp).  No compute calls are made without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

import mazu_b200 as mz

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols(paths=None):
    out = set()
    for path in paths or (mz.HEADER_PATH, mz.DEBUG_HEADER_PATH):
        text = open(path).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        out |= set(re.findall(r"\b(mazu_b200_[a-z0-9_]+)\s*\(", text))
    return sorted(out)


def test_debug_hooks_live_in_their_own_header():
    """the drop-in boundary (mazu_b200.h) declares no debug / measurement hook; those are in mazu_b200_debug.h"""
    assert not [s for s in _header_symbols([mz.HEADER_PATH]) if "debug" in s or "measure" in s]
    assert all("debug" in s for s in _header_symbols([mz.DEBUG_HEADER_PATH]))


def test_header_matches_binding():
    assert _header_symbols() == mz.exported_symbols()


def test_rust_sys_crate_declares_every_symbol():
    """bindings/rust/mazu-b200-sys mirrors include/mazu_b200.h symbol for symbol (no Rust toolchain here: a text check)."""
    text = open(os.path.join(ROOT, "bindings", "rust", "mazu-b200-sys", "src", "lib.rs")).read()
    rust = sorted(set(re.findall(r"pub fn (mazu_b200_[a-z0-9_]+)\s*\(", text)))
    assert rust == _header_symbols()


def _build_c_example(tmp_path):
    exe = str(tmp_path / "query_reads")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "query_reads.c"),
           "-L" + os.path.dirname(mz.LIB_PATH), "-lmazu_b200", "-Wl,-rpath," + os.path.dirname(mz.LIB_PATH), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_plain_c_and_example_links(tmp_path):
    """include/mazu_b200.h compiles as C99 with -Wall -Wextra -Werror and a C program links against the library;
    without a GPU the example stops at its device check (exit 3), never in a fallback."""
    exe = _build_c_example(tmp_path)
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "data", "pf1", "yeast_chr01_index")], capture_output=True, text=True)
    if mz.device_count() <= 0:
        assert r.returncode == 3 and "no CPU fallback" in r.stderr


def test_library_exports_every_declared_symbol():
    assert os.path.exists(mz.LIB_PATH), "libmazu_b200.so not built (run __graft_entry__.build())"
    L = C.CDLL(mz.LIB_PATH)
    for name in _header_symbols():
        assert hasattr(L, name), name


def test_library_has_sm100a_code():
    out = subprocess.run(["cuobjdump", "-lelf", mz.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_library_has_the_kw_instantiations_of_the_sshash_read_kernels():
    """The SSHash read kernels are compiled with k and w read from the index (KW = 0) and with (31, 19) / (31, 15) folded into
    the code (capi.cu: kw_code, kernels.cuh: KW).  The GPU parity test runs both forms; here: all of them are in the cubin."""
    out = subprocess.run(["cuobjdump", "-elf", mz.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    text = out.stdout
    kw = {"generic": "Lj0E", "k31 w19": "Lj%dE" % ((31 << 8) | 19), "k31 w15": "Lj%dE" % ((31 << 8) | 15)}
    # mangled template argument lists: <MODE, KIND = SSHASH (1), FAMILY = NATIVE (1), OCC, KW> etc.
    for label, code in kw.items():
        assert re.search(r"query_reads_kernelILi0ELi1ELj1ELi\d+E%sE" % code, text), "random-access read kernel, " + label
        assert re.search(r"query_reads_kernelILi1ELi1ELj1ELi\d+E%sE" % code, text), "streaming walk, " + label
        assert re.search(r"get_ref_pos_pass1_kernelILi1ELj1E%sE" % code, text), "get_ref_pos pass 1, " + label
        for io in range(4):
            assert re.search(r"query_reads_runs_kernelILi1ELj1ELi%dE%sE" % (io, code), text), "fused run kernel io %d, %s" % (io, label)


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    if mz.device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(mz.MazuError) as e:
        mz.DenseIndex.deserialize_from_cpp(os.path.join(ROOT, "tests", "data", "pf1", "tiny_index"))
    assert e.value.code == -5


def test_product_does_not_reference_oracle():
    """Nothing under mazu_b200/ or include/ may include, link or import oracle/."""
    for base in ("mazu_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".so", ".log", ".pyc")):
                    continue
                text = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in text.lower(), os.path.join(dp, f)
    out = subprocess.run(["ldd", mz.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_host_builders_self_consistent(tmp_path):
    exe = str(tmp_path / "host_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-Wno-unknown-pragmas", os.path.join(ROOT, "tests", "host_check.cpp"), "-o", exe])
    r = subprocess.run([exe, os.path.join(ROOT, "tests/data/pf1/yeast_chr01_index"), os.path.join(ROOT, "tests/data/cf/tiny/tiny")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "host_check: 0 failures" in r.stdout


def test_hit_run_decoder_round_trip():
    """mazu_b200_expand_hit_runs is host code (a format decoder, no device work): encode the oracle's records as runs with a
    plain numpy restatement of the format and check the library's decoder gives the records back, ragged and uniform."""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _gen
    from _oracle import OracleIndex
    o = OracleIndex.dense_from_pf1(os.path.join(ROOT, "tests", "data", "pf1", "yeast_chr01_index")).rebuild_k2u(1, w=15, skew=32, seed=0)
    ref_codes = _gen.unpack_2bit(o.refseq_words(), int(o.ref_prefix()[-1]))
    for ragged in (True, False):
        bases, offs = _gen.sample_reads(ref_codes, 800, 160, seed=77, frac_ref=0.7, sub_rate=0.01, n_rate=0.003, ragged=ragged)
        hits, _, koffs = o.query_reads(bases, offs)
        n = len(hits)
        is_hit = (hits["match"] == mz.IDENTITY_MATCH) | (hits["match"] == mz.TWIN_MATCH)
        first = np.zeros(n, dtype=bool)
        first[koffs[:-1][np.diff(koffs.astype(np.int64)) > 0].astype(np.int64)] = True  # first slot of every non-empty read
        prev = np.roll(hits, 1)
        step = np.where(hits["match"] == mz.IDENTITY_MATCH, prev["pos"] + np.uint32(1), prev["pos"] - np.uint32(1))
        cont = is_hit & ~first & (prev["match"] == hits["match"]) & (prev["unitig_id"] == hits["unitig_id"]) & (hits["pos"] == step)
        codes = np.zeros(n, dtype=np.uint8)
        codes[is_hit & cont] = 1
        codes[is_hit & ~cont] = 2
        codes[hits["match"] == mz.SKIPPED] = 3
        runs = np.ascontiguousarray(hits[codes == 2])
        starts_before = np.concatenate([[0], np.cumsum(codes == 2)])
        rro = starts_before[koffs.astype(np.int64)].astype(np.uint64)
        if ragged:
            got = mz.ModIndex.expand_hit_runs(codes, runs, rro, kmer_offsets=koffs)
        else:
            got = mz.ModIndex.expand_hit_runs(codes, runs, rro, uniform_slots=160 - o.k + 1)
        assert np.array_equal(got.view(np.uint32), hits.view(np.uint32))
        assert len(runs) < n // 8


def test_pack_reads_and_packed_run_decoder():
    """host halves of the packed run interface (no device work): mazu_b200_pack_reads against numpy bit packing, and
    mazu_b200_expand_hit_runs_packed against mazu_b200_expand_hit_runs on the same runs with 2-bit codes."""
    import numpy as np
    rng = np.random.default_rng(3)
    n_reads, read_len, k = 300, 150, 31
    bases = np.frombuffer(b"ACGTacgtNx", dtype=np.uint8)[rng.choice(10, size=n_reads * read_len, p=[.2, .2, .2, .2, .04, .04, .04, .04, .02, .02])].copy()
    words, mask, bad = mz.pack_reads(bases, read_len)
    codes = np.full(256, 4, dtype=np.uint8)
    for i, ch in enumerate(b"ACGT"):
        codes[ch] = i
        codes[ch | 0x20] = i
    c = codes[bases].reshape(n_reads, read_len)
    assert bad == int((c == 4).sum())
    wpr, mpr = (read_len + 31) // 32, (read_len + 63) // 64
    want_w = np.zeros((n_reads, wpr), dtype=np.uint64)
    want_m = np.zeros((n_reads, mpr), dtype=np.uint64)
    for j in range(read_len):
        want_w[:, j // 32] |= np.where(c[:, j] < 4, c[:, j], 0).astype(np.uint64) << np.uint64(2 * (j % 32))
        want_m[:, j // 64] |= (c[:, j] == 4).astype(np.uint64) << np.uint64(j % 64)
    assert np.array_equal(words.reshape(n_reads, wpr), want_w) and np.array_equal(mask.reshape(n_reads, mpr), want_m)
    # every byte value, read lengths around the 8-base steps of the packer and the 32- / 64-base word boundaries
    for rl in (1, 7, 8, 9, 31, 32, 33, 63, 64, 65, 150, 151):
        nr = 97
        b = rng.integers(0, 256, size=nr * rl, dtype=np.uint8)
        keep = rng.random(nr * rl) < 0.8  # mostly bases, the rest any byte
        b[keep] = np.frombuffer(b"ACGTacgt", dtype=np.uint8)[rng.integers(0, 8, size=int(keep.sum()))]
        if nr * rl >= 256:
            b[:256] = np.arange(256, dtype=np.uint8)
        w2, m2, bad2 = mz.pack_reads(b, rl)
        cc = codes[b].reshape(nr, rl)
        assert bad2 == int((cc == 4).sum()), rl
        ww = np.zeros((nr, (rl + 31) // 32), dtype=np.uint64)
        mm = np.zeros((nr, (rl + 63) // 64), dtype=np.uint64)
        for j in range(rl):
            ww[:, j // 32] |= np.where(cc[:, j] < 4, cc[:, j], 0).astype(np.uint64) << np.uint64(2 * (j % 32))
            mm[:, j // 64] |= (cc[:, j] == 4).astype(np.uint64) << np.uint64(j % 64)
        assert np.array_equal(w2.reshape(nr, -1), ww) and np.array_equal(m2.reshape(nr, -1), mm), rl
    # decoder: random codes / runs, byte codes vs 2-bit codes
    slots = read_len - k + 1
    code = rng.choice(4, size=n_reads * slots, p=[.4, .45, .05, .1]).astype(np.uint8)
    code.reshape(n_reads, slots)[:, 0] = np.where(code.reshape(n_reads, slots)[:, 0] == 1, 2, code.reshape(n_reads, slots)[:, 0])  # a read never starts with "continues"
    flat = code.reshape(-1)
    for i in range(1, len(flat)):  # "continues" must follow a hit
        if flat[i] == 1 and flat[i - 1] not in (1, 2):
            flat[i] = 2
    runs = np.zeros(int((flat == 2).sum()), dtype=mz.HIT_DTYPE)
    runs["unitig_id"] = rng.integers(0, 1000, len(runs))
    runs["unitig_len"] = 5000
    runs["pos"] = rng.integers(200, 4000, len(runs))
    runs["match"] = rng.integers(1, 3, len(runs))
    rro = np.concatenate([[0], np.cumsum((flat == 2).reshape(n_reads, slots).sum(axis=1))]).astype(np.uint64)
    want = mz.ModIndex.expand_hit_runs(flat, runs, rro, uniform_slots=slots)
    packed = np.zeros((len(flat) + 3) // 4, dtype=np.uint8)
    for q in range(4):
        part = flat[q::4]
        packed[: len(part)] |= (part << (2 * q)).astype(np.uint8)
    got = mz.ModIndex.expand_hit_runs_packed(packed, runs, rro, slots)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("extra,prefix", [(["--workload", "config2"], "configs[1]"), (["--cpu-scale", "0.002"], "configs[4]")])
def test_bench_reference_arm_contract(extra, prefix):
    """`bench.py --impl reference` needs no GPU: it must print ONE JSON line with the driver's keys (metric, value, unit, n_gpus,
    steps, warmup, ms_per_step, higher_is_better, scaling, vs_baseline, dtype, data, config.workload) plus impl / cpu_baseline / e2e.
    The default workload is configs[4] (the port builds its index at --cpu-scale); `config` is the GPU arm's, key for key."""
    import json
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-seconds", "1"] + extra,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "lookups/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith(prefix) and len(d["config"]["workload"]) <= 118  # the driver keeps 120 characters
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # same config dict as the GPU arm would print for the same flags
    sys.path.insert(0, ROOT)
    import bench
    sys.argv = ["bench.py"] + extra
    assert d["config"] == bench.bench_config(bench.parse_args())
    if prefix == "configs[4]":
        # 70 % reference reads x 0.99^31 surviving k-mers x ~70 % of the windows of a 150 bp read inside one unitig (mean length 98)
        assert 0.3 < d["counts"]["n_hit"] / d["counts"]["n_kmers"] < 0.42
