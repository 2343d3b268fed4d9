"""Pins the CPU oracle (oracle/) against every golden vector / known-answer test the reference
holds for the k-mer query path.  Each test cites the reference test it ports.
CPU only."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import _oracle as O
from _oracle import OracleIndex, encode_kmer, lib, _ptr

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
PF1 = os.path.join(DATA, "pf1")
TINY_INDEX = os.path.join(PF1, "tiny_index")
TINY_REFS_INDEX = os.path.join(PF1, "tiny-multi-refs", "tiny-multi-refs_index")
SMALL_TXOME = os.path.join(PF1, "small_txome_index")
YEAST_CHR01 = os.path.join(PF1, "yeast_chr01_index")
TINY_CF = os.path.join(DATA, "cf", "tiny", "tiny")
YEAST_CF = os.path.join(DATA, "cf", "yeast_chr7", "yeast_chr7")
NOSKEW = O.USIZE_MAX


# ---- pf1/boophf/hash.rs:152-253 -----------------------------------------------------------
def test_simple_hash_first10():
    true_hashes = [0x6E1BCCDB7AA2BC25, 0x54676A7B01425B7, 0x5C9BE323E5AD1BE1, 0x9567829F5E948F83, 0xCF71E329165C79B5,
                   0x9F1219F1BCD9D206, 0x6BD828B35DBA940E, 0xF55B08C3340017C3, 0xD178AE94742FA575, 0x5DC299D49318DC6B]
    for key, h in enumerate(true_hashes):
        assert lib().orc_simple_hash64(key, 0xAAAAAAAA55555555) == h


def test_multihash_zero_five():
    out = np.zeros(5, dtype=np.uint64)
    lib().orc_multihash_chain(0, 5, _ptr(out))
    assert [int(x) for x in out] == [7934160411570650149, 4031181471818755726, 7802733314557663513,
                                     5772550616205298107, 3882642898705877381]


# ---- pf1/boophf/mod.rs:339-424 -----------------------------------------------------------
@pytest.fixture(scope="module")
def bbhash10():
    h = lib().orc_boophf_load(os.path.join(PF1, "bbhash_n=10.bin").encode())
    assert h
    yield C.c_void_p(h)
    lib().orc_boophf_free(C.c_void_p(h))


def _lookup(h, key):
    out = C.c_uint64(0)
    ok = lib().orc_boophf_lookup(h, key, C.byref(out))
    return out.value if ok else None


def test_boophf_parity_from_cpp_serialized(bbhash10):
    assert lib().orc_boophf_info(bbhash10, 0, 0) == 10
    assert lib().orc_boophf_info(bbhash10, 2, 0) == 2
    assert lib().orc_boophf_info(bbhash10, 1, 0) == 2


def test_boophf_level0_bits(bbhash10):
    assert lib().orc_boophf_info(bbhash10, 3, 0) == 2312599096050843650


def test_boophf_levels_ranks(bbhash10):
    assert lib().orc_boophf_check_ranks(bbhash10) == 1


def test_boophf_lookups(bbhash10):
    assert _lookup(bbhash10, 0) == 2
    for item, want in enumerate([2, 0, 8, 3, 5, 4, 1, 7, 6, 9, 7]):
        assert _lookup(bbhash10, item) == want


def test_boophf_lookup_does_not_exist(bbhash10):
    assert _lookup(bbhash10, 10) == 7
    assert _lookup(bbhash10, 11) == 0
    assert _lookup(bbhash10, 12) == 0
    for i in range(13, 20):
        assert _lookup(bbhash10, i) is None


def test_boophf_final_hash(bbhash10):
    out = C.c_uint64(0)
    assert lib().orc_boophf_final_lookup(bbhash10, 9, C.byref(out)) and out.value == 9
    assert lib().orc_boophf_final_lookup(bbhash10, 2, C.byref(out)) and out.value == 8


@pytest.mark.parametrize("name", ["example_10_100", "example_100_10", "example_1e6_1e3"])
def test_boophf_example_tables(name):
    """Key->hash known-answer tables dumped by the C++ BBHash (unused by the reference's own tests)."""
    h = C.c_void_p(lib().orc_boophf_load(os.path.join(PF1, name + ".bin").encode()))
    assert h
    js = json.load(open(os.path.join(PF1, name + ".json")))
    assert lib().orc_boophf_info(h, 0, 0) == js["nelems"]
    n = 0
    for table in ("random_hashed_elems", "random_elems"):
        for key, val in js[table].items():
            got = _lookup(h, int(key))
            want = None if val == O.USIZE_MAX else val
            assert got == want, (table, key)
            n += 1
    assert n > 0
    lib().orc_boophf_free(h)


# ---- pf1/cpp.rs:244-289 ---------------------------------------------------------------------
def test_compact_vector_tiny_pos():
    w, n = C.c_uint64(0), C.c_uint64(0)
    out = np.zeros(16, dtype=np.uint64)
    assert lib().orc_compact_vector_read(os.path.join(TINY_INDEX, "pos.bin").encode(), C.byref(w), C.byref(n), _ptr(out), 16) == 0
    assert (w.value, n.value) == (3, 4)
    assert [int(x) for x in out[:4]] == [3, 2, 0, 1]


def test_compact_vector_yeast_pos_width():
    w, n = C.c_uint64(0), C.c_uint64(0)
    assert lib().orc_compact_vector_read(os.path.join(YEAST_CHR01, "pos.bin").encode(), C.byref(w), C.byref(n), None, 0) == 0
    assert w.value == 18 and n.value == 221918


# ---- elias_fano.rs:13-20,147-165 ---------------------------------------------------------------
def test_ef_vigna_fig1():
    xs = np.array([5, 8, 8, 15, 32], dtype=np.uint64)
    out = np.zeros(5, dtype=np.uint64)
    assert lib().orc_ef_roundtrip(_ptr(xs), 5, 32, 0, _ptr(out), None) == 0
    assert list(out) == list(xs)


def test_ef_not_monotone():
    xs = np.array([5, 8, 7, 15, 32], dtype=np.uint64)
    out = np.zeros(5, dtype=np.uint64)
    assert lib().orc_ef_roundtrip(_ptr(xs), 5, 32, 0, _ptr(out), None) == -1
    assert b"EFNotMonotone" in lib().orc_last_error()


def test_ef_random_roundtrip():
    rng = np.random.default_rng(7)
    for n, hi in [(1, 1), (10, 3), (1000, 10), (1000, 100000), (5000, 2)]:
        xs = np.cumsum(rng.integers(0, hi, size=n)).astype(np.uint64)
        if xs[-1] == 0:
            xs[-1] = 1
        out = np.zeros(n, dtype=np.uint64)
        assert lib().orc_ef_roundtrip(_ptr(xs), n, 0, 1, _ptr(out), None) == 0
        assert np.array_equal(out, xs)


# ---- spt_compact.rs:504-520, index.rs:320-346 ---------------------------------------------------
def test_occ_encoding_piscem():
    ref_shift = 3
    pos_mask = (1 << (ref_shift - 1)) - 1
    word = lib().orc_encode_piscem(0, 1, 0, ref_shift)
    assert word == 0b010
    out = np.zeros(3, dtype=np.uint32)
    lib().orc_decode_piscem(ref_shift, pos_mask, word, _ptr(out))
    assert list(out) == [0, 1, 0]


def test_occ_encoding_pf1_roundtrip():
    rng = np.random.default_rng(3)
    out = np.zeros(3, dtype=np.uint32)
    for _ in range(200):
        r, p, f = int(rng.integers(0, 2**32 - 1)), int(rng.integers(0, 2**31)), int(rng.integers(0, 2))
        w = lib().orc_encode_pf1(r, p, f)
        assert w == ((p | (0x80000000 if f else 0)) << 32 | r)
        lib().orc_decode_pf1(w, _ptr(out))
        assert list(out) == [r, p, f]


def test_required_num_bits():
    out = np.zeros(3, dtype=np.uint32)
    assert lib().orc_required_num_bits(1 << 27, 4096, _ptr(out)) == 0
    assert list(out) == [28, 13, 42]  # SURVEY config 4
    assert lib().orc_required_num_bits(0, 1, _ptr(out)) == -1


# ---- unitig_set.rs:353-381 --------------------------------------------------------------------
def test_unitig_set_tiny():
    idx = OracleIndex.from_cf(TINY_CF, 0)
    assert idx.k == 7
    assert idx.n_unitigs == 2
    assert idx.unitig_len(0) == 10 and idx.unitig_len(1) == 10
    for i in range(10):
        assert idx.pos_to_id(i) == 0
    for i in range(10, 20):
        assert idx.pos_to_id(i) == 1
    assert idx.total_len == 20
    words = idx.useq_words()
    s = "".join("ACGT"[(int(words[(2 * i) // 64]) >> ((2 * i) % 64)) & 3] for i in range(20))
    assert s == "CACACACCAC" + "CCTCAATACG"


# ---- pf1/dense_index.rs:109-328 ----------------------------------------------------------------
@pytest.fixture(scope="module")
def tiny_dense():
    return OracleIndex.dense_from_pf1(TINY_INDEX)


@pytest.fixture(scope="module")
def yeast_dense():
    return OracleIndex.dense_from_pf1(YEAST_CHR01)


def test_tiny_n_unitigs(tiny_dense):
    assert tiny_dense.n_unitigs == 1


def test_yeast_n_unitigs(yeast_dense):
    assert yeast_dense.n_unitigs == 577
    assert yeast_dense.n_kmers == 221918
    assert yeast_dense.total_len == 239228
    assert yeast_dense.n_total_occs == 1029


def test_tiny_query_does_not_exist(tiny_dense):
    for km in ["tat", "ata", "act", "ctg", "cct"]:
        assert tiny_dense.get_ref_pos_eager(km) is None


def test_tiny_kmer_wrong_size_panics(tiny_dense):
    with pytest.raises(O.OracleError):
        tiny_dense.get_ref_pos_eager("aaaaaa")
    with pytest.raises(O.OracleError):
        tiny_dense.get_ref_pos_eager("aa")


def _revcomp(s):
    return s.upper()[::-1].translate(str.maketrans("ACGT", "TGCA"))


def _check_tiny_positions(idx):
    # Indexed string: AAACCC
    for pos, km in enumerate(["aaa", "aac", "acc", "ccc"]):
        _, mrps = idx.get_ref_pos_eager(km)
        assert mrps == [(0, pos, 1)]
        _, mrps = idx.get_ref_pos_eager(_revcomp(km))
        assert mrps == [(0, pos, 0)]


def test_tiny_kmer_positions(tiny_dense):
    _check_tiny_positions(tiny_dense)


def test_tiny_sshash_k2u(tiny_dense):
    idx = tiny_dense.rebuild_k2u(1, w=2, skew=NOSKEW)
    _check_tiny_positions(idx)
    assert idx.validate_self()[4] == 0


@pytest.mark.parametrize("d", [TINY_INDEX, TINY_REFS_INDEX, SMALL_TXOME])
def test_validate_dense_small(d):
    idx = OracleIndex.dense_from_pf1(d)
    c = idx.validate_self()
    assert c[0] > 0 and c[4] == 0


def test_validate_yeast_dense(yeast_dense):
    # 230,188 queries -> 170,689 identity + 59,499 twin -> 262,130 projected positions
    assert yeast_dense.validate_self() == [230188, 170689, 59499, 262130, 0]
    c = yeast_dense.k2u_validate_self()
    assert c[0] == 443836 and c[4] == 0


def test_validate_yeast_sshash(yeast_dense):
    idx = yeast_dense.rebuild_k2u(1, w=15, skew=32)
    assert idx.validate_self() == [230188, 170689, 59499, 262130, 0]
    c = idx.k2u_validate_self()
    assert c[0] == 443836 and c[4] == 0


# ---- kphf/sshash.rs:633-884 ----------------------------------------------------------------------
def _tiny_w_params(w):
    idx = OracleIndex.from_cf(TINY_CF, 1, w=w, skew=NOSKEW)
    cases = [("CACACAC", 0, 0), ("ACACACC", 0, 1), ("ACACCAC", 0, 3), ("CCTCAAT", 1, 0), ("CAATACG", 1, 3)]
    for s, uid, pos in cases:
        fw = idx.k2u_fw_batch([encode_kmer(s)])[0]
        both = idx.k2u_batch([encode_kmer(s), encode_kmer(_revcomp(s))])
        assert tuple(fw) == (uid, 10, pos, O.MATCH_IDENTITY)
        assert tuple(both[0]) == (uid, 10, pos, O.MATCH_IDENTITY)
        assert tuple(both[1]) == (uid, 10, pos, O.MATCH_TWIN)
    miss = idx.k2u_fw_batch([encode_kmer("AAAAAAA")])[0]
    assert miss["match"] == O.MATCH_NONE and miss["unitig_id"] == O.MISS
    both = idx.k2u_batch([encode_kmer("AAAAAAA"), encode_kmer("TTTTTTT")])
    assert all(b["match"] == O.MATCH_NONE for b in both)


def test_sshash_tiny_w3():
    _tiny_w_params(3)


def test_sshash_tiny_w5():
    _tiny_w_params(5)


def test_sshash_tiny_vary_windows():
    for w in range(1, 8):
        _tiny_w_params(w)


def test_sshash_tiny_validate_self():
    assert OracleIndex.from_cf(TINY_CF, 1, w=3, skew=NOSKEW).k2u_validate_self()[4] == 0
    assert OracleIndex.from_cf(TINY_CF, 1, w=5, skew=0).k2u_validate_self()[4] == 0


def test_sshash_unitigs_share_mmer():
    seqs = ["ACAACTTACCCTCCATTACCCTACCTCCCCA", "CAACTTACCCTCCATTACCCTACCTCCCCAC"]
    idx = OracleIndex.from_seqs(seqs, 31, 1, w=15, skew=NOSKEW)
    k1 = encode_kmer(seqs[1])
    assert idx.k2u_fw_batch([encode_kmer(_revcomp(seqs[1]))])[0]["match"] == O.MATCH_NONE
    assert tuple(idx.k2u_fw_batch([k1])[0]) == (1, 31, 0, O.MATCH_IDENTITY)
    assert tuple(idx.k2u_fw_batch([encode_kmer(seqs[0])])[0]) == (0, 31, 0, O.MATCH_IDENTITY)
    assert idx.k2u_validate_self()[4] == 0


def test_sshash_tiny_kmer_too_small_panics():
    idx = OracleIndex.from_cf(TINY_CF, 1, w=3, skew=NOSKEW)
    with pytest.raises(O.OracleError):
        idx.k2u_fw_batch([0], qk=idx.k + 1)
    with pytest.raises(O.OracleError):
        idx.k2u_batch([0], qk=idx.k + 1)


def test_pfhash_tiny_and_wrong_k():  # kphf/pfhash.rs:292-309
    idx = OracleIndex.from_cf(TINY_CF, 0)
    assert idx.k2u_validate_self()[4] == 0
    with pytest.raises(O.OracleError):
        idx.k2u_batch([0], qk=idx.k + 1)


def test_sshash_tiny_skew_index():
    skew = OracleIndex.from_cf(TINY_CF, 1, w=3, skew=0)
    noskew = OracleIndex.from_cf(TINY_CF, 1, w=3, skew=NOSKEW)
    assert skew.n_kmers_in_skew_index == skew.n_kmers
    for s in ["CACACAC", "ACACCAC", "CCTCAAT", "CCTCAAT"]:
        q = [encode_kmer(s), encode_kmer(_revcomp(s))]
        a, b = skew.k2u_batch(q), noskew.k2u_batch(q)
        assert np.array_equal(a, b) and a[0]["match"] == O.MATCH_IDENTITY
    assert skew.k2u_batch([encode_kmer("AAAAAAA")])[0]["match"] == O.MATCH_NONE


# ---- kphf/mod.rs:147-161 -----------------------------------------------------------------------
def test_yeast_cf_sshash_validate_self():
    idx = OracleIndex.from_cf(YEAST_CF, 1, w=15, skew=NOSKEW)
    c = idx.k2u_validate_self()
    assert c[0] == 2 * 1071346 and c[4] == 0


def test_yeast_cf_pfhash_validate_self():
    idx = OracleIndex.from_cf(YEAST_CF, 0)
    c = idx.k2u_validate_self()
    assert c[0] == 2 * 1071346 and c[4] == 0


# ---- spt.rs:156-211, spt_compact.rs:415-495 ------------------------------------------------------
@pytest.mark.parametrize("kind", [0, 1])
def test_spt_tiny(kind):
    idx = OracleIndex.from_cf(TINY_CF, kind, w=3, skew=NOSKEW)
    assert idx.n_refs == 2
    assert idx.n_total_occs == 4
    assert idx.ref_len(0) == 3 + 10 + 1 + 10
    offs, occs = idx.decode_occs([0, 1])
    assert list(offs) == [0, 2, 4]
    assert [tuple(o) for o in occs] == [(0, 3, 1), (1, 11, 0), (0, 14, 0), (1, 0, 1)]


@pytest.mark.parametrize("prefix", [TINY_CF, YEAST_CF])
def test_compare_spt_impls(prefix):
    a = OracleIndex.from_cf(prefix, 0)
    b = OracleIndex.from_cf(prefix, 1, w=15 if prefix == YEAST_CF else 3, skew=NOSKEW)
    ids = np.arange(a.n_unitigs, dtype=np.uint32)
    oa, xa = a.decode_occs(ids)
    ob, xb = b.decode_occs(ids)
    assert np.array_equal(oa, ob) and np.array_equal(xa, xb)
    if prefix == YEAST_CF:
        assert a.n_total_occs == 1051


# ---- index/piscem_index.rs:63-99, index/defaults.rs:60-71, index/caching.rs:238-253 ---------------
def test_piscem_validate_tiny():
    idx = OracleIndex.from_cf(TINY_CF, 1, w=3, skew=2)
    c = idx.validate_fasta(TINY_CF + ".fa")
    assert c[0] == 16 and c[4] == 0  # 4 + 4 valid windows per record (N-skipping), 2 records


def test_piscem_validate_yeast_random_and_streaming():
    idx = OracleIndex.from_cf(YEAST_CF, 1, w=15, skew=32)
    assert idx.n_minimizers < idx.n_kmers
    assert idx.n_kmers_in_skew_index != 0
    c = idx.validate_fasta(YEAST_CF + ".fa")
    assert c[0] == 1090910 and c[4] == 0
    s = idx.validate_fasta(YEAST_CF + ".fa", streaming=True)
    assert s == c


def test_pufferfish_dense_from_cf_validate_yeast():
    idx = OracleIndex.from_cf(YEAST_CF, 0)
    c = idx.validate_fasta(YEAST_CF + ".fa")
    assert c[0] == 1090910 and c[4] == 0
    assert idx.validate_fasta(YEAST_CF + ".fa", streaming=True) == c


# ---- pf1/sparse_index.rs:145-192 (SampledPFHash, pufferfish sparse index) ------------------------
SMALL_TXOME_SPARSE = os.path.join(PF1, "small_txome_index_sparse")


def test_sparse_index_params_and_validate():
    idx = OracleIndex.sparse_from_pf1(SMALL_TXOME_SPARSE)
    assert idx.info(12) == 9 and idx.info(13) == 4  # sample_size, extension_size
    c = idx.k2u_validate_self()
    assert c[0] == 2 * 18902 and c[4] == 0
    v = idx.validate_self()
    assert v[0] == 28112 and v[4] == 0
    dense = OracleIndex.dense_from_pf1(SMALL_TXOME).validate_self()
    assert (v[0], v[4]) == (dense[0], dense[4])  # same transcriptome through the dense index


def test_sparse_index_sshash_drop_in():
    idx = OracleIndex.sparse_from_pf1(SMALL_TXOME_SPARSE)
    ss = idx.rebuild_k2u(1, w=2, skew=NOSKEW)
    assert ss.validate_self()[4] == 0


# ---- refseq.rs:280-309 test_contig_iter (ModIndex::iter_unitigs_on_ref, index.rs:363-424) ----------
def test_contig_iter_tiny_multi_refs():
    idx = OracleIndex.dense_from_pf1(TINY_REFS_INDEX)
    t0, t1 = idx.iter_unitigs_on_ref(0), idx.iter_unitigs_on_ref(1)
    assert list(t0["unitig_len"]) == [5, 8, 9, 8, 5] and list(t0["unitig_id"]) == [0, 1, 2, 3, 4]
    assert list(t1["unitig_len"]) == [5, 9, 9, 9, 5] and list(t1["unitig_id"]) == [0, 5, 2, 6, 4]
    assert idx.ref_len(1) == 21 and idx.k == 5


def test_oracle_reproduces_golden_fixture():
    """tests/golden/yeast_chr01_queries.npz (made by tests/golden/make_golden.py) is still what the oracle answers"""
    import os
    import sys
    import numpy as np
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_golden
    g = np.load(os.path.join(here, "golden", "yeast_chr01_queries.npz"))
    o, bases, offs = make_golden.inputs()
    assert np.array_equal(bases, g["bases"]) and np.array_equal(offs, g["read_offsets"])
    assert list(o.validate_self()) == list(g["validate_self"])
    ss = o.rebuild_k2u(1, w=15, skew=32, seed=0)
    for name, ix in (("pfhash", o), ("sshash", ss)):
        for streaming in (False, True):
            hits, cnt, koffs = ix.query_reads(bases, offs, streaming=streaming, reset_per_read=True)
            tag = "%s_%s" % (name, "streaming" if streaming else "random")
            assert np.array_equal(hits.view(np.uint32).reshape(-1, 4), g[tag + "_hits"]), tag
            assert list(cnt) == list(g[tag + "_counts"])
            assert np.array_equal(koffs, g["kmer_offsets"])
