"""GPU parity tests proper: the CUDA path, called through the C ABI (libmazu_b200.so), must be
bit-exact with the CPU oracle on the same inputs.  Every test is marked `gpu`."""
import os

import numpy as np
import pytest

import _gen
import _oracle as O
import mazu_b200 as mz
from _oracle import OracleIndex

pytestmark = pytest.mark.gpu

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
PF1 = os.path.join(DATA, "pf1")
TINY_INDEX = os.path.join(PF1, "tiny_index")
TINY_REFS_INDEX = os.path.join(PF1, "tiny-multi-refs", "tiny-multi-refs_index")
SMALL_TXOME = os.path.join(PF1, "small_txome_index")
YEAST_CHR01 = os.path.join(PF1, "yeast_chr01_index")
TINY_CF = os.path.join(DATA, "cf", "tiny", "tiny")
YEAST_CF = os.path.join(DATA, "cf", "yeast_chr7", "yeast_chr7")
NOSKEW = mz.SKEW_NONE


def _revcomp(s):
    return s.upper()[::-1].translate(str.maketrans("ACGT", "TGCA"))


def assert_hits_equal(got, want, what=""):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, what
    if not np.array_equal(got, want):
        bad = np.nonzero(got != want)[0]
        i = int(bad[0])
        raise AssertionError("%s: %d/%d records differ; first at %d: got %s want %s" % (what, len(bad), len(got), i, got[i], want[i]))


def read_fasta(path):
    recs, cur = [], None
    for line in open(path):
        line = line.rstrip("\r\n")
        if line.startswith(">"):
            cur = []
            recs.append(cur)
        elif cur is not None:
            cur.append(line)
    return ["".join(r) for r in recs]


def fasta_as_reads(path):
    seqs = read_fasta(path)
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(s) for s in seqs])
    return np.frombuffer("".join(seqs).encode(), dtype=np.uint8).copy(), offs


# --------------------------------------------------------------------------------------------
# fixtures: (device index, oracle index) pairs
# --------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def yeast_dense():
    return mz.DenseIndex.deserialize_from_cpp(YEAST_CHR01), OracleIndex.dense_from_pf1(YEAST_CHR01)


@pytest.fixture(scope="module")
def yeast_sshash(yeast_dense):
    g, o = yeast_dense
    return g.rebuild_k2u(mz.K2U_SSHASH, w=15, skew_param=32, seed=0), o.rebuild_k2u(1, w=15, skew=32, seed=0)


@pytest.fixture(scope="module")
def yeast_queries(yeast_dense):
    """all reference k-mers + all unitig k-mers fw and rc + random negatives (config 1's query set)."""
    _, o = yeast_dense
    k = o.k
    ref_codes = _gen.unpack_2bit(o.refseq_words(), int(o.ref_prefix()[-1]))
    ref_kmers = _gen.kmer_words_from_codes(ref_codes, k)
    useq_codes = _gen.unpack_2bit(o.useq_words(), o.total_len)
    u_kmers = _gen.kmer_words_from_codes(useq_codes, k)  # includes boundary-straddling windows: fine, they are queries too
    rng = np.random.default_rng(1)
    neg = rng.integers(0, 1 << 62, size=50000, dtype=np.uint64)
    rc = np.array([O.lib().orc_revcomp(int(x), k) for x in u_kmers[:20000]], dtype=np.uint64)
    q = np.concatenate([ref_kmers, u_kmers, rc, neg])
    rng.shuffle(q)
    return q, ref_codes


# --------------------------------------------------------------------------------------------
# pf1/dense_index.rs:109-328 through the GPU path
# --------------------------------------------------------------------------------------------
def test_tiny_dense_golden_answers():
    idx = mz.DenseIndex.deserialize_from_cpp(TINY_INDEX)
    assert idx.n_unitigs == 1 and idx.k == 3
    for pos, km in enumerate(["aaa", "aac", "acc", "ccc"]):
        assert idx.get_ref_pos_eager(km) == [(0, pos, 1)]
        assert idx.get_ref_pos_eager(_revcomp(km)) == [(0, pos, 0)]
    for km in ["tat", "ata", "act", "ctg", "cct"]:
        assert idx.get_ref_pos_eager(km) is None
    for bad in ["aaaaaa", "aa"]:  # the reference panics (dense_index.rs:148-163)
        with pytest.raises(mz.MazuError) as e:
            idx.get_ref_pos_eager(bad)
        assert e.value.code == mz.ERR_K_MISMATCH
    # same answers with SSHash w=2 behind the same U2Pos (dense_index.rs:214-265)
    ss = idx.rebuild_k2u(mz.K2U_SSHASH, w=2, skew_param=NOSKEW)
    for pos, km in enumerate(["aaa", "aac", "acc", "ccc"]):
        assert ss.get_ref_pos_eager(km) == [(0, pos, 1)]
        assert ss.get_ref_pos_eager(_revcomp(km)) == [(0, pos, 0)]
    assert ss.validate_self()[4] == 0


@pytest.mark.parametrize("d", [TINY_INDEX, TINY_REFS_INDEX, SMALL_TXOME, YEAST_CHR01])
def test_validate_self_dense(d):
    g, o = mz.DenseIndex.deserialize_from_cpp(d), OracleIndex.dense_from_pf1(d)
    assert (g.k, g.n_unitigs, g.n_kmers, g.sum_unitigs_len) == (o.k, o.n_unitigs, o.n_kmers, o.total_len)
    want = o.validate_self()
    assert g.validate_self() == want and want[4] == 0
    assert g.k2u_validate_self() == o.k2u_validate_self()
    if d == YEAST_CHR01:
        assert want == [230188, 170689, 59499, 262130, 0]


def test_validate_self_yeast_sshash(yeast_sshash):
    g, o = yeast_sshash
    assert g.validate_self() == [230188, 170689, 59499, 262130, 0]
    c = g.k2u_validate_self()
    assert c == o.k2u_validate_self() and c[0] == 443836 and c[4] == 0
    assert g.n_kmers_in_skew_index == o.n_kmers_in_skew_index > 0


def test_k2u_batch_pfhash_yeast(yeast_dense, yeast_queries):
    g, o = yeast_dense
    q, _ = yeast_queries
    assert_hits_equal(g.k2u_batch(q), o.k2u_batch(q), "PFHash/BooPHF k2u")


def test_k2u_batch_native_pfhash_yeast(yeast_dense, yeast_queries):
    g, o = yeast_dense
    q, _ = yeast_queries
    g2 = g.rebuild_k2u(mz.K2U_PFHASH)
    assert_hits_equal(g2.k2u_batch(q), o.k2u_batch(q), "PFHash/native MPHF k2u")
    assert g2.validate_self() == [230188, 170689, 59499, 262130, 0]


@pytest.mark.parametrize("w,skew", [(15, 32), (15, NOSKEW), (19, 4), (9, 64), (31, 0), (1, 8)])
def test_k2u_batch_sshash_yeast(yeast_dense, yeast_queries, w, skew):
    g, o = yeast_dense
    q, _ = yeast_queries
    q = q[:200000]
    gs = g.rebuild_k2u(mz.K2U_SSHASH, w=w, skew_param=skew, seed=7)
    os_ = o.rebuild_k2u(1, w=w, skew=skew, seed=7)
    assert_hits_equal(gs.k2u_batch(q), os_.k2u_batch(q), "SSHash k2u w=%d skew=%s" % (w, skew))
    assert gs.n_kmers_in_skew_index == os_.n_kmers_in_skew_index


def test_k2u_batch_device_mode_matches_host_mode(yeast_sshash, yeast_queries):
    import torch
    g, _ = yeast_sshash
    q, _ = yeast_queries
    q = q[:100000]
    host = g.k2u_batch(q)
    dq = torch.from_numpy(q.view(np.int64)).cuda()
    dout = torch.empty((len(q), 4), dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream()
    g.k2u_batch(dq, out=dout, mem=mz.MEM_DEVICE, stream=s.cuda_stream, n=len(q))
    s.synchronize()
    dev = dout.cpu().numpy().view(np.uint32).reshape(-1).view(mz.HIT_DTYPE)
    assert_hits_equal(dev, host, "device-mode vs host-mode")


def test_k2u_batch_edge_cases(yeast_sshash):
    g, _ = yeast_sshash
    assert len(g.k2u_batch(np.zeros(0, dtype=np.uint64))) == 0
    with pytest.raises(mz.MazuError) as e:
        g.k2u_batch(np.zeros(4, dtype=np.uint64), k=g.k + 1)
    assert e.value.code == mz.ERR_K_MISMATCH
    one = g.k2u_batch(np.array([0], dtype=np.uint64))  # a single k-mer = what a Rust `impl K2U` would send
    assert len(one) == 1


# --------------------------------------------------------------------------------------------
# kphf/sshash.rs:633-884 golden answers through the GPU path
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("w", list(range(1, 8)))
def test_sshash_tiny_golden(w):
    g = mz.PiscemIndex.from_cf_prefix(TINY_CF, w, NOSKEW)
    cases = [("CACACAC", 0, 0), ("ACACACC", 0, 1), ("ACACCAC", 0, 3), ("CCTCAAT", 1, 0), ("CAATACG", 1, 3)]
    q = np.array([mz.encode_kmer(s) for s, _, _ in cases] + [mz.encode_kmer(_revcomp(s)) for s, _, _ in cases]
                 + [mz.encode_kmer("AAAAAAA"), mz.encode_kmer("TTTTTTT")], dtype=np.uint64)
    h = g.k2u_batch(q)
    for i, (_, uid, pos) in enumerate(cases):
        assert tuple(h[i]) == (uid, 10, pos, mz.IDENTITY_MATCH)
        assert tuple(h[i + 5]) == (uid, 10, pos, mz.TWIN_MATCH)
    assert tuple(h[10]) == (mz.MISS, mz.MISS, mz.MISS, mz.NO_MATCH) and h[11]["match"] == mz.NO_MATCH
    assert g.k2u_validate_self()[4] == 0


def test_sshash_tiny_skew_equals_no_skew():
    a = mz.PiscemIndex.from_cf_prefix(TINY_CF, 3, 0)
    b = mz.PiscemIndex.from_cf_prefix(TINY_CF, 3, NOSKEW)
    assert a.n_kmers_in_skew_index == a.n_kmers
    q = np.array([mz.encode_kmer(s) for s in ["CACACAC", "ACACCAC", "CCTCAAT", "GTGTGTG", "ATTGAGG", "AAAAAAA"]], dtype=np.uint64)
    assert_hits_equal(a.k2u_batch(q), b.k2u_batch(q))
    assert a.k2u_validate_self()[4] == 0


def test_sshash_unitigs_share_mmer():
    seqs = ["ACAACTTACCCTCCATTACCCTACCTCCCCA", "CAACTTACCCTCCATTACCCTACCTCCCCAC"]
    g = mz.SSHash.from_unitig_set_no_skew_index(mz.UnitigSet.from_seqs(seqs, 31), 15)
    o = OracleIndex.from_seqs(seqs, 31, 1, w=15)
    q = np.array([mz.encode_kmer(s) for s in seqs] + [mz.encode_kmer(_revcomp(s)) for s in seqs], dtype=np.uint64)
    h = g.k2u_batch(q)
    assert tuple(h[0]) == (0, 31, 0, mz.IDENTITY_MATCH) and tuple(h[1]) == (1, 31, 0, mz.IDENTITY_MATCH)
    assert_hits_equal(h, o.k2u_batch(q))
    assert g.k2u_validate_self() == o.k2u_validate_self()


# --------------------------------------------------------------------------------------------
# K1: encode / canonical k-mers / minimizers
# --------------------------------------------------------------------------------------------
def test_encode_reads_matches_oracle(yeast_sshash, yeast_queries):
    import torch
    g, o = yeast_sshash
    _, ref_codes = yeast_queries
    bases, offs = _gen.sample_reads(ref_codes, 300, 200, seed=3, frac_ref=0.6, sub_rate=0.01, n_rate=0.004, ragged=True)
    koffs = o.kmer_offsets(offs)
    n = int(koffs[-1])
    dev = lambda a: torch.from_numpy(a).cuda()
    d_bases, d_offs, d_koffs = dev(bases), dev(offs.view(np.int64)), dev(koffs.view(np.int64))
    fw = torch.zeros(n, dtype=torch.int64, device="cuda"); rc = torch.zeros_like(fw); mm = torch.zeros_like(fw)
    off = torch.zeros(n, dtype=torch.int32, device="cuda"); valid = torch.zeros(n, dtype=torch.uint8, device="cuda")
    g.encode_reads(d_bases, d_offs, len(offs) - 1, 0, d_koffs, fw, rc, mm, off, valid, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    fw, rc, mm = (t.cpu().numpy().view(np.uint64) for t in (fw, rc, mm))
    off, valid = off.cpu().numpy().view(np.uint32), valid.cpu().numpy()
    k, w = o.k, 15
    for r in range(len(offs) - 1):
        seq = bases[int(offs[r]):int(offs[r + 1])]
        ns = int(koffs[r + 1] - koffs[r])
        efw = np.zeros(max(ns, 1), dtype=np.uint64); erc = np.zeros_like(efw); emm = np.zeros_like(efw)
        eoff = np.zeros(max(ns, 1), dtype=np.uint32); ev = np.zeros(max(ns, 1), dtype=np.uint8)
        O.lib().orc_encode_read(O._ptr(seq), len(seq), k, w, 0, O._ptr(efw), O._ptr(erc), O._ptr(emm), O._ptr(eoff), O._ptr(ev))
        s = slice(int(koffs[r]), int(koffs[r]) + ns)
        assert np.array_equal(valid[s], ev[:ns]), r
        assert np.array_equal(fw[s], efw[:ns]) and np.array_equal(rc[s], erc[:ns]), r
        assert np.array_equal(mm[s], emm[:ns]) and np.array_equal(off[s], eoff[:ns]), r


# --------------------------------------------------------------------------------------------
# query_reads: random-access and streaming
# --------------------------------------------------------------------------------------------
def _check_reads(g, o, bases, offs, mode):
    streaming = mode == mz.MODE_STREAMING
    want, wcnt, wk = o.query_reads(bases, offs, streaming=streaming, reset_per_read=True)
    got, gcnt, gk = g.query_reads(bases, offs, mode=mode)
    assert np.array_equal(gk, wk)
    assert_hits_equal(got, want, "query_reads mode=%d" % mode)
    assert list(gcnt) == list(wcnt)
    # counts-only call (the `kphf bench` shape)
    _, c2, _ = g.query_reads(bases, offs, mode=mode, want_hits=False)
    assert list(c2) == list(wcnt)
    return want, wcnt


@pytest.mark.parametrize("mode", [mz.MODE_RANDOM, mz.MODE_STREAMING])
@pytest.mark.parametrize("which", ["sshash", "pfhash"])
def test_query_reads_mixed_ragged(yeast_dense, yeast_sshash, yeast_queries, mode, which):
    g, o = yeast_sshash if which == "sshash" else yeast_dense
    _, ref_codes = yeast_queries
    bases, offs = _gen.sample_reads(ref_codes, 4000, 260, seed=11, frac_ref=0.7, sub_rate=0.01, n_rate=0.002, ragged=True)
    want, cnt = _check_reads(g, o, bases, offs, mode)
    assert cnt[1] > 0 and cnt[2] > 0 and (want["match"] == mz.SKIPPED).any()


@pytest.mark.parametrize("mode", [mz.MODE_RANDOM, mz.MODE_STREAMING])
def test_query_reads_uniform_150bp(yeast_sshash, yeast_queries, mode):
    """config 2/3 shape: uniform 150 bp reads through the uniform_read_len fast path == ragged path == oracle."""
    g, o = yeast_sshash
    _, ref_codes = yeast_queries
    bases, offs = _gen.sample_reads(ref_codes, 20000, 150, seed=42, frac_ref=0.5, sub_rate=0.01 if mode else 0.0)
    want, wcnt = _check_reads(g, o, bases, offs, mode)
    got_u, cnt_u, _ = g.query_reads(bases, None, uniform_read_len=150, mode=mode)
    assert_hits_equal(got_u, want, "uniform path")
    assert list(cnt_u) == list(wcnt)


def test_query_reads_edge_cases(yeast_sshash):
    g, o = yeast_sshash
    k = g.k
    # empty batch, empty reads, reads shorter than k, exactly k, all-N
    seqs = ["", "ACGT", "A" * (k - 1), "ACGTACGTACGTACGTACGTACGTACGTACG", "N" * 64, "acgtn" * 20, ""]
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(s) for s in seqs])
    bases = np.frombuffer("".join(seqs).encode(), dtype=np.uint8).copy()
    for mode in (mz.MODE_RANDOM, mz.MODE_STREAMING):
        _check_reads(g, o, bases, offs, mode)
    hits, cnt, ko = g.query_reads(np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.uint64))
    assert len(hits) == 0 and list(cnt) == [0, 0, 0] and list(ko) == [0]


def test_query_reads_device_mode(yeast_sshash, yeast_queries):
    import torch
    g, o = yeast_sshash
    _, ref_codes = yeast_queries
    bases, offs = _gen.sample_reads(ref_codes, 3000, 180, seed=5, frac_ref=0.6, sub_rate=0.02, n_rate=0.001, ragged=True)
    for mode in (mz.MODE_RANDOM, mz.MODE_STREAMING):
        want, wcnt, wk = o.query_reads(bases, offs, streaming=bool(mode))
        d_bases = torch.from_numpy(bases).cuda()
        d_offs = torch.from_numpy(offs.view(np.int64)).cuda()
        d_k = torch.zeros(len(offs), dtype=torch.int64, device="cuda")
        d_hits = torch.zeros((len(want), 4), dtype=torch.int32, device="cuda")
        d_cnt = torch.zeros(3, dtype=torch.int64, device="cuda")
        g.query_reads(d_bases, d_offs, n_reads=len(offs) - 1, mode=mode, out_hits=d_hits, kmer_offsets=d_k, counts=d_cnt,
                      mem=mz.MEM_DEVICE, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(d_k.cpu().numpy().view(np.uint64), wk)
        assert_hits_equal(d_hits.cpu().numpy().view(np.uint32).reshape(-1).view(mz.HIT_DTYPE), want, "device mode %d" % mode)
        assert list(d_cnt.cpu().numpy()) == list(wcnt.astype(np.int64))


@pytest.mark.parametrize("ragged", [True, False])
def test_query_reads_host_in_device_out(yeast_sshash, yeast_queries, ragged):
    """MAZU_MEM_HOST_IN_DEVICE_OUT: reads on the host, records left in HBM (then projected on the device), counters back."""
    import torch
    g, o = yeast_sshash
    _, ref_codes = yeast_queries
    bases, offs = _gen.sample_reads(ref_codes, 2500, 150, seed=9, frac_ref=0.6, sub_rate=0.02, n_rate=0.001, ragged=ragged)
    for mode in (mz.MODE_RANDOM, mz.MODE_STREAMING):
        for compact in (False, True):
            want, wcnt, wk = o.query_reads(bases, offs, streaming=bool(mode))
            d_hits = torch.zeros((len(want), 2 if compact else 4), dtype=torch.int32, device="cuda")
            kw = dict(read_offsets=offs) if ragged else dict(uniform_read_len=150)
            _, cnt, ko = g.query_reads(bases, mode=mode, out_hits=d_hits, mem=mz.MEM_HOST_IN_DEVICE_OUT, compact=compact, **kw)
            assert list(cnt) == list(wcnt)
            if ragged:
                assert np.array_equal(ko, wk)
            got = d_hits.cpu().numpy().view(np.uint32).reshape(-1)
            if compact:
                got = got.view(mz.HIT8_DTYPE)
                hit = (want["match"] == mz.IDENTITY_MATCH) | (want["match"] == mz.TWIN_MATCH)
                assert np.array_equal(got["unitig_id"][hit], want["unitig_id"][hit])
                assert np.array_equal(got["pos_match"] >> 30, want["match"])
                assert np.array_equal((got["pos_match"] & 0x3FFFFFFF)[hit], want["pos"][hit])
            else:
                assert_hits_equal(got.view(mz.HIT_DTYPE), want, "host-in/device-out mode %d" % mode)
    with pytest.raises(mz.MazuError) as e:  # the other entry points only know HOST and DEVICE
        g.k2u_batch(np.zeros(4, dtype=np.uint64), out=np.zeros(4, dtype=mz.HIT_DTYPE), mem=mz.MEM_HOST_IN_DEVICE_OUT, n=4)
    assert e.value.code == -7


@pytest.mark.parametrize("which", ["dense", "sshash"])
def test_long_reads_are_cut_into_segments(yeast_dense, yeast_sshash, yeast_queries, which):
    """Random-access lookups of one read are independent: the kernel cuts reads longer than 2048 k-mer positions into
    segments handled by different warps.  Lengths around the segment boundary, mixed with short and empty reads; the
    streaming walk (sequential per read, never segmented) must agree with its own oracle on the same batch."""
    g, o = yeast_dense if which == "dense" else yeast_sshash
    _, ref_codes = yeast_queries
    k = g.k
    rng = np.random.default_rng(77)
    lens = [0, k - 1, k, 150, 2047 + k, 2048 + k - 1, 2048 + k, 4096 + k - 1, 4097 + k - 1, 30000, 150, 7, 100000, 2048 * 3 + k - 1, 151]
    parts, offs = [], [0]
    for i, ln in enumerate(lens):
        if i % 3 == 2 or ln > len(ref_codes):
            codes = rng.integers(0, 4, size=ln, dtype=np.uint8)
        else:
            s0 = int(rng.integers(0, len(ref_codes) - ln + 1))
            codes = ref_codes[s0:s0 + ln].copy()
            if i % 2:
                codes = _gen.COMP[codes[::-1]]
        m = rng.random(ln) < 0.003
        codes[m] = (codes[m] + 1) & 3
        b = _gen.ACGT[codes].copy()
        b[rng.random(ln) < 0.0005] = ord("N")
        parts.append(b)
        offs.append(offs[-1] + ln)
    bases = np.concatenate(parts)
    offs = np.array(offs, dtype=np.uint64)
    for mode in (mz.MODE_RANDOM, mz.MODE_STREAMING):
        _check_reads(g, o, bases, offs, mode)
    # device-pointer mode goes through the same segment table
    import torch
    want, wcnt, wk = o.query_reads(bases, offs, streaming=False)
    d_hits = torch.zeros((len(want), 4), dtype=torch.int32, device="cuda")
    d_cnt = torch.zeros(3, dtype=torch.int64, device="cuda")
    g.query_reads(torch.from_numpy(bases).cuda(), torch.from_numpy(offs.view(np.int64)).cuda(), n_reads=len(offs) - 1, out_hits=d_hits,
                  counts=d_cnt, mem=mz.MEM_DEVICE, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert_hits_equal(d_hits.cpu().numpy().view(np.uint32).reshape(-1).view(mz.HIT_DTYPE), want, "segmented, device mode")
    assert list(d_cnt.cpu().numpy()) == list(wcnt.astype(np.int64))


def test_streaming_exact_with_duplicate_kmers():
    """A unitig set whose canonical k-mers are NOT unique (not a valid cdBG): random-access and
    streaming answers differ there, and the GPU walk must reproduce the reference's sequential
    warm/cold decisions (caching.rs:65-103) exactly, including first-match order inside a bucket."""
    rng = np.random.default_rng(9)
    k = 15
    core = "".join("ACGT"[i] for i in rng.integers(0, 4, size=60))
    # unitig 1 owns k-mers 0..9 of `core` alone and shares 10..25 with unitig 0, which is earlier in
    # bucket order: a walk that enters on unitig 1 stays there, a random-access lookup answers unitig 0
    seqs = [core[10:60], core[:40], "".join("ACGT"[i] for i in rng.integers(0, 4, size=50)), _revcomp(core[5:45]), core[20:50] + "ACGTTGCA"]
    us = mz.UnitigSet.from_seqs(seqs, k)
    reads = [core, _revcomp(core), core[3:33] + "T" + core[34:], seqs[2] + core[10:40], core[:25] + "N" + core[26:], core[12:58]]
    offs = np.zeros(len(reads) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    bases = np.frombuffer("".join(reads).encode(), dtype=np.uint8).copy()
    differs = False
    for w, skew in [(7, NOSKEW), (5, 1), (11, 0)]:
        g = mz.SSHash.from_unitig_set(us, w, skew)
        o = OracleIndex.from_seqs(seqs, k, 1, w=w, skew=skew)
        assert not g.kmers_unique  # detected at creation: the cursor-walk kernel serves streaming queries
        r_want, _ = _check_reads(g, o, bases, offs, mz.MODE_RANDOM)
        s_want, _ = _check_reads(g, o, bases, offs, mz.MODE_STREAMING)
        differs |= not np.array_equal(r_want, s_want)
    assert differs, "test input should make streaming and random-access answers differ"


def test_unique_kmers_make_streaming_equal_random(yeast_dense, yeast_sshash, yeast_queries):
    """A set with pairwise distinct canonical k-mers (any cdBG): a warm hit of StreamingK2U is the k-mer's only occurrence, so
    streaming answers equal K2U::k2u record for record; the library detects this at creation (MAZU_INFO_KMERS_UNIQUE) and serves
    streaming queries with the random-access kernel.  Both modes still equal the oracle's streaming / random walks."""
    _, ref_codes = yeast_queries
    bases, offs = _gen.sample_reads(ref_codes, 1500, 150, seed=21, frac_ref=0.7, sub_rate=0.01, n_rate=0.002, ragged=True)
    for g, o in (yeast_dense, yeast_sshash):
        assert g.kmers_unique
        r_want, _ = _check_reads(g, o, bases, offs, mz.MODE_RANDOM)
        s_want, _ = _check_reads(g, o, bases, offs, mz.MODE_STREAMING)
        assert np.array_equal(r_want, s_want)


# --------------------------------------------------------------------------------------------
# validate_fasta on the cuttlefish fixtures (piscem_index.rs:63-99, defaults.rs:60-71, caching.rs:238-253)
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["piscem", "pufferfish"])
def test_validate_fasta_tiny_with_poly_n(kind):
    if kind == "piscem":
        g, o = mz.PiscemIndex.from_cf_prefix(TINY_CF, 3, 2), OracleIndex.from_cf(TINY_CF, 1, w=3, skew=2)
    else:
        g, o = mz.PufferfishDenseIndex.from_cf_prefix(TINY_CF), OracleIndex.from_cf(TINY_CF, 0)
    # Validate::validate_fasta / StreamingIndex::validate_fasta through the C ABI: the LIBRARY reads the file (poly-N, lower case)
    for mode in (mz.MODE_RANDOM, mz.MODE_STREAMING):
        got = g.validate_fasta(TINY_CF + ".fa", mode)
        assert got == o.validate_fasta(TINY_CF + ".fa", streaming=bool(mode)) and got[0] == 16 and got[4] == 0
    assert g.as_streaming().validate_fasta(TINY_CF + ".fa")[4] == 0
    fa = mz.Fasta(TINY_CF + ".fa")
    bases, offs = fa.bases, fa.offsets
    b2, o2 = fasta_as_reads(TINY_CF + ".fa")
    assert np.array_equal(bases, b2) and np.array_equal(offs, o2)
    for mode in (mz.MODE_RANDOM, mz.MODE_STREAMING):
        want, cnt = _check_reads(g, o, bases, offs, mode)
        assert cnt[0] == 16 and cnt[2] == 0
    # projected positions contain (record, pos) for every valid window
    hits, _, koffs = g.query_reads(bases, offs)
    poffs, mrps = g.project_hits(hits)
    ooffs, omrps = o.project_hits(hits)
    assert np.array_equal(poffs, ooffs) and np.array_equal(mrps, omrps)
    for r in range(len(offs) - 1):
        for p in range(int(koffs[r + 1] - koffs[r])):
            i = int(koffs[r]) + p
            if hits[i]["match"] == mz.SKIPPED:
                continue
            lst = mrps[int(poffs[i]):int(poffs[i + 1])]
            assert any(m["ref_id"] == r and m["pos"] == p for m in lst)


@pytest.mark.parametrize("kind", ["piscem", "pufferfish"])
def test_validate_fasta_yeast_chr7_long_record(kind):
    """One 1.09 Mbp record: exercises the chunked walk of a long read in both modes."""
    if kind == "piscem":
        g, o = mz.PiscemIndex.from_cf_prefix(YEAST_CF, 15, 32), OracleIndex.from_cf(YEAST_CF, 1, w=15, skew=32)
        assert g.n_kmers_in_skew_index == o.n_kmers_in_skew_index != 0
        assert g.n_minimizers == o.n_minimizers < g.n_kmers
    else:
        g, o = mz.PufferfishDenseIndex.from_cf_prefix(YEAST_CF), OracleIndex.from_cf(YEAST_CF, 0)
    for mode in (mz.MODE_RANDOM, mz.MODE_STREAMING):
        got = g.validate_fasta(YEAST_CF + ".fa", mode)
        assert got == o.validate_fasta(YEAST_CF + ".fa", streaming=bool(mode)) and got[0] == 1090910 and got[4] == 0
    fa = mz.Fasta(YEAST_CF + ".fa")
    bases, offs = fa.bases, fa.offsets
    for mode in (mz.MODE_RANDOM, mz.MODE_STREAMING):
        want, cnt = _check_reads(g, o, bases, offs, mode)
        assert cnt[0] == 1090910 and cnt[2] == 0
    hits, _, _ = g.query_reads(bases, offs)
    poffs, mrps = g.project_hits(hits)
    ooffs, omrps = o.project_hits(hits)
    assert np.array_equal(poffs, ooffs) and np.array_equal(mrps, omrps)
    # every k-mer maps back to its own position on reference 0 (validate_ckmers)
    pos = np.arange(len(hits), dtype=np.uint32)
    owner = np.repeat(pos, np.diff(poffs).astype(np.int64))
    ok = np.zeros(len(hits), dtype=bool)
    ok[owner[(mrps["ref_id"] == 0) & (mrps["pos"] == owner)]] = True
    assert ok.all()
    assert g.k2u_validate_self() == o.k2u_validate_self()


@pytest.mark.parametrize("which", ["sshash", "dense"])
def test_fused_get_ref_pos_reads_equals_the_chain(yeast_dense, yeast_sshash, yeast_queries, which):
    """mazu_b200_get_ref_pos_reads (reads -> K2UPos -> occurrences -> MappedRefPos in one kernel, single-pass look-back) must equal
    query_reads followed by project_hits on the oracle, record for record: ragged reads with N / short / empty reads, uniform
    reads, long reads spanning many tiles, both modes."""
    g, o = yeast_sshash if which == "sshash" else yeast_dense
    _, ref_codes = yeast_queries
    cases = [_gen.sample_reads(ref_codes, 2500, 150, seed=31, frac_ref=0.7, sub_rate=0.01, n_rate=0.002, ragged=True),
             _gen.sample_reads(ref_codes, 40, 5000, seed=32, frac_ref=0.9, sub_rate=0.002, n_rate=0.0005, ragged=True)]
    for bases, offs in cases:
        for mode in (mz.MODE_RANDOM, mz.MODE_STREAMING):
            want_hits, wcnt, wko = o.query_reads(bases, offs, streaming=bool(mode))
            want_offs, want_mrps = o.project_hits(want_hits)
            hits, poffs, mrps, cnt, ko = g.get_ref_pos_reads(bases, read_offsets=offs, mode=mode)
            assert_hits_equal(hits, want_hits, "fused hits")
            assert np.array_equal(poffs, want_offs) and np.array_equal(mrps, want_mrps)
            assert list(cnt) == list(wcnt) and np.array_equal(ko, wko)
    bases, offs = _gen.sample_reads(ref_codes, 3000, 150, seed=33, frac_ref=0.6, sub_rate=0.01)
    want_hits, wcnt, _ = o.query_reads(bases, offs)
    want_offs, want_mrps = o.project_hits(want_hits)
    hits, poffs, mrps, cnt, _ = g.get_ref_pos_reads(bases, uniform_read_len=150, want_hits=False)
    assert hits is None and np.array_equal(poffs, want_offs) and np.array_equal(mrps, want_mrps) and list(cnt) == list(wcnt)
    # empty batch and a batch without a single k-mer
    _, poffs, mrps, cnt, _ = g.get_ref_pos_reads(np.zeros(0, dtype=np.uint8), read_offsets=np.zeros(1, dtype=np.uint64))
    assert list(poffs) == [0] and len(mrps) == 0
    _, poffs, mrps, cnt, _ = g.get_ref_pos_reads(np.frombuffer(b"ACGTACGT", dtype=np.uint8), read_offsets=np.array([0, 3, 3, 8], dtype=np.uint64))
    assert list(poffs) == [0] and len(mrps) == 0 and list(cnt) == [0, 0, 0]


def test_fused_get_ref_pos_high_multiplicity_and_small_capacity():
    """long occurrence lists are written by the whole warp; an undersized device buffer is never overrun"""
    import torch
    import ctypes as C
    k, U = 31, 400
    codes, accum = _gen.synthetic_unitigs(U, 68, k, seed=51)
    us = mz.UnitigSet(k, mz.pack_2bit(codes), len(codes), accum)
    g = mz.SSHash.from_unitig_set(us, 19, 64)
    o = OracleIndex.from_packed(k, us.useq_words, us.n_bases, accum, 1, w=19, skew=64)
    n_refs, max_ref_len = 4096, 1 << 27
    offsets, ref_ids, poss, fws = _synthetic_u2pos(U, n_refs, max_ref_len, 52, 3000)
    o.attach_u2pos(1, offsets, ref_ids, poss, fws, max_ref_len, n_refs)
    enc = (ref_ids.astype(np.uint64) << np.uint64(29)) | (poss.astype(np.uint64) << np.uint64(1)) | fws.astype(np.uint64)
    g.attach_u2pos_piscem(mz.PackedVec.pack(enc, 42), 29, (1 << 28) - 1, mz.PackedVec.pack(offsets))
    bases, offs = _gen.sample_reads(codes, 600, 150, seed=53, frac_ref=0.8, sub_rate=0.01, n_rate=0.001, ragged=True)
    want_hits, wcnt, _ = o.query_reads(bases, offs)
    want_offs, want_mrps = o.project_hits(want_hits)
    assert int(np.diff(want_offs).max()) >= 32  # lists long enough for the warp-cooperative path
    hits, poffs, mrps, cnt, _ = g.get_ref_pos_reads(bases, read_offsets=offs)
    assert_hits_equal(hits, want_hits, "fused hits")
    assert np.array_equal(poffs, want_offs) and np.array_equal(mrps, want_mrps)
    # device mode, capacity smaller than the output, no out_total: filled up to cap, guard untouched, need reported in out_offsets[n]
    n_slots, total = len(want_hits), len(want_mrps)
    d_b = torch.from_numpy(bases).cuda()
    d_ro = torch.from_numpy(offs.view(np.int64)).cuda()
    for cap in (total, total // 3, 0):
        d_offs = torch.zeros(n_slots + 1, dtype=torch.int64, device="cuda")
        d_out = torch.full((cap + 1024, 3), -1, dtype=torch.int32, device="cuda")
        mz._check(mz.lib().mazu_b200_get_ref_pos_reads(g._h, mz._any_ptr(d_b), mz._any_ptr(d_ro), len(offs) - 1, 0, 0, n_slots, None, None,
                                                       mz._any_ptr(d_offs), mz._any_ptr(d_out), cap, None, None, mz.MEM_DEVICE, None))
        torch.cuda.synchronize()
        got = d_out.cpu().numpy().view(np.uint32)
        assert np.array_equal(d_offs.cpu().numpy().view(np.uint64), want_offs)
        assert (got[cap:] == 0xFFFFFFFF).all() and np.array_equal(got[:cap].reshape(-1).view(mz.OCC_DTYPE), want_mrps[:cap])
    tot = C.c_uint64(0)
    d_out = torch.full((total + 8, 3), -1, dtype=torch.int32, device="cuda")
    rc = mz.lib().mazu_b200_get_ref_pos_reads(g._h, mz._any_ptr(d_b), mz._any_ptr(d_ro), len(offs) - 1, 0, 0, n_slots, None, None,
                                              mz._any_ptr(d_offs), mz._any_ptr(d_out), total - 1, C.byref(tot), None, mz.MEM_DEVICE, None)
    assert rc == -7 and tot.value == total


def test_validate_fasta_reports_failures(tmp_path, yeast_dense):
    """records that are NOT the index's references: every k-mer that misses, or maps elsewhere, is a failure (the reference panics)"""
    g, o = yeast_dense
    ref_codes = _gen.unpack_2bit(o.refseq_words(), int(o.ref_prefix()[-1]))
    ref = _gen.ACGT[ref_codes[:5000]].tobytes().decode()
    rng = np.random.default_rng(5)
    rnd = _gen.ACGT[rng.integers(0, 4, 600)].tobytes().decode()
    # record 0 = a true prefix of reference 0 (passes, lower-cased and wrapped); record 1 = a slice from position 1000, so its k-mers
    # project onto (ref 0, 1000 + p), not (ref 1, p): found but failing; record 2 = random sequence: misses; record 3 shorter than k
    path = tmp_path / "mix.fa"
    with open(path, "w") as f:
        f.write(">r0 a prefix\n")
        low = ref[:3000].lower()
        for i in range(0, len(low), 70):
            f.write(low[i:i + 70] + "\r\n")
        f.write(">r1\n" + ref[1000:1500] + "\n>r2\n" + rnd + "\nNNNN\n>r3\nACGT\n")
    for mode in (mz.MODE_RANDOM, mz.MODE_STREAMING):
        got = g.validate_fasta(path, mode)
        want = o.validate_fasta(str(path), streaming=bool(mode))
        assert got == want
        assert got[0] == (3000 - 30) + (500 - 30) + (600 - 30) and got[4] == (500 - 30) + (600 - 30)
    fa = mz.Fasta(path)
    assert fa.names == ["r0 a prefix", "r1", "r2", "r3"] and fa.seqs()[0] == ref[:3000].lower() and fa.seqs()[2] == rnd + "NNNN"
    assert g.validate_reads(fa.bases, fa.offsets) == g.validate_fasta(path)
    with pytest.raises(mz.MazuError) as e:
        g.validate_fasta(tmp_path / "missing.fa")
    assert e.value.code == -1


def test_fasta_reader_semantics(tmp_path):
    """FastaReader's own unit test (src/util.rs:157-185) + FASTQ; host code, but it is the ingest of the GPU validate path"""
    p = tmp_path / "a.fa"
    p.write_text(">A\nACG\nAC\n>B\nACG\nACG\nC")
    fa = mz.Fasta(p)
    assert fa.names == ["A", "B"] and fa.seqs() == ["ACGAC", "ACGACGC"] and list(fa.offsets) == [0, 5, 12]
    q = tmp_path / "a.fq"
    q.write_text("@r1 x\nACGTN\n+\nIIIII\n@r2\nGG\n+r2\nII\n")
    fq = mz.Fasta(q)
    assert fq.names == ["r1 x", "r2"] and fq.seqs() == ["ACGTN", "GG"]
    bad = tmp_path / "bad.fq"
    bad.write_text("@r1\nACGT\n+\nIII\n")
    with pytest.raises(mz.MazuError) as e:
        mz.Fasta(bad)
    assert e.value.code == -2
    empty = tmp_path / "e.fa"
    empty.write_text("")
    assert mz.Fasta(empty).names == []


def test_unitig_seq_accessor(yeast_dense):
    """K2U::unitig_seq (src/kphf/mod.rs:64) through the ABI equals the oracle's unitig sequence"""
    g, o = yeast_dense
    codes = _gen.unpack_2bit(o.useq_words(), o.total_len)
    for ui in (0, 1, 17, o.n_unitigs - 1):
        s0, n = o.unitig_start(ui), o.unitig_len(ui)
        assert g.unitig_seq(ui) == _gen.ACGT[codes[s0:s0 + n]].tobytes().decode()
    with pytest.raises(mz.MazuError):
        g.unitig_seq(o.n_unitigs)


# --------------------------------------------------------------------------------------------
# K4: occurrence decode and projection
# --------------------------------------------------------------------------------------------
def test_decode_occs_fixtures(yeast_dense):
    g, o = yeast_dense
    ids = np.arange(g.n_unitigs, dtype=np.uint32)
    go, gx = g.decode_occs(ids)
    oo, ox = o.decode_occs(ids)
    assert np.array_equal(go, oo) and np.array_equal(gx, ox) and len(gx) == 1029
    for prefix, kind in [(TINY_CF, 0), (TINY_CF, 1), (YEAST_CF, 1)]:
        gi = mz.ModIndex.from_cf_prefix(prefix, kind, w=3 if prefix == TINY_CF else 15)
        oi = OracleIndex.from_cf(prefix, kind, w=3 if prefix == TINY_CF else 15)
        ids = np.concatenate([np.arange(gi.n_unitigs, dtype=np.uint32), np.array([mz.MISS], dtype=np.uint32)])
        a, b = gi.decode_occs(ids), oi.decode_occs(ids)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    gi = mz.ModIndex.from_cf_prefix(TINY_CF, 1, w=3)
    offs, occs = gi.decode_occs(np.array([0, 1], dtype=np.uint32))
    assert list(offs) == [0, 2, 4] and [tuple(x) for x in occs] == [(0, 3, 1), (1, 11, 0), (0, 14, 0), (1, 0, 1)]  # spt.rs:156-211


def _synthetic_u2pos(n_unitigs, n_refs, max_ref_len, seed, max_mult):
    rng = np.random.default_rng(seed)
    mult = np.minimum(rng.zipf(1.2, size=n_unitigs), max_mult).astype(np.uint64)
    offsets = np.zeros(n_unitigs + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(mult)
    n = int(offsets[-1])
    ref_ids = rng.integers(0, n_refs, size=n, dtype=np.uint32)
    poss = rng.integers(0, max_ref_len - 2000, size=n, dtype=np.uint32)
    fws = rng.integers(0, 2, size=n, dtype=np.uint8)
    return offsets, ref_ids, poss, fws


@pytest.mark.parametrize("kind", ["dense", "piscem"])
def test_decode_and_project_synthetic_high_multiplicity(kind):
    """config 4 shape at test size: Zipf(1.2) multiplicities, both encodings, queries drawn ~ multiplicity."""
    k, U = 31, 3000
    codes, accum = _gen.synthetic_unitigs(U, 68, k, seed=44)
    us = mz.UnitigSet(k, mz.pack_2bit(codes), len(codes), accum)
    g = mz.SSHash.from_unitig_set(us, 19, 64)
    o = OracleIndex.from_packed(k, us.useq_words, us.n_bases, accum, 1, w=19, skew=64)
    n_refs, max_ref_len = 4096, 1 << 27
    offsets, ref_ids, poss, fws = _synthetic_u2pos(U, n_refs, max_ref_len, 44, 5000)
    o.attach_u2pos(0 if kind == "dense" else 1, offsets, ref_ids, poss, fws, max_ref_len, n_refs)
    off_vec = mz.PackedVec.pack(offsets)
    if kind == "dense":
        words = ((poss.astype(np.uint64) | (fws.astype(np.uint64) << np.uint64(31))) << np.uint64(32)) | ref_ids.astype(np.uint64)
        g.attach_u2pos_dense(words, off_vec)
    else:
        out = np.zeros(3, dtype=np.uint32)
        assert O.lib().orc_required_num_bits(max_ref_len, n_refs, O._ptr(out)) == 0
        pos_bits, ref_bits, total = (int(x) for x in out)
        assert (pos_bits, ref_bits, total) == (28, 13, 42)
        ref_shift, pos_mask = pos_bits + 1, (1 << pos_bits) - 1
        enc = (ref_ids.astype(np.uint64) << np.uint64(ref_shift)) | (poss.astype(np.uint64) << np.uint64(1)) | fws.astype(np.uint64)
        g.attach_u2pos_piscem(mz.PackedVec.pack(enc, total), ref_shift, pos_mask, off_vec)
    rng = np.random.default_rng(45)
    mult = np.diff(offsets).astype(np.float64)
    q = rng.choice(U, size=20000, p=mult / mult.sum()).astype(np.uint32)
    q[::97] = mz.MISS
    a, b = g.decode_occs(q), o.decode_occs(q)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # projection for a random (pos, orientation) per query
    hits = np.zeros(len(q), dtype=mz.HIT_DTYPE)
    ulen = np.diff(accum).astype(np.uint32)
    valid = q != mz.MISS
    hits["unitig_id"] = q
    hits["unitig_len"][valid] = ulen[q[valid]]
    hits["pos"][valid] = (rng.random(int(valid.sum())) * (ulen[q[valid]] - k + 1)).astype(np.uint32)
    hits["match"] = np.where(valid, rng.integers(1, 3, size=len(q)), 0)
    hits["unitig_id"][~valid] = mz.MISS
    a, b = g.project_hits(hits), o.project_hits(hits)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and len(a[1]) > len(q)
    with pytest.raises(mz.MazuError) as e:
        mz.SSHash.from_unitig_set(us, 19, 64).decode_occs(q)
    assert e.value.code == mz.ERR_NO_U2POS


@pytest.mark.parametrize("kind", ["dense", "piscem"])
def test_decode_device_mode_undersized_buffer_is_never_overrun(kind):
    """MAZU_MEM_DEVICE without out_total cannot report a too-small buffer (no synchronisation): the fill kernels must stop at
    `cap`.  A guard region behind the buffer stays untouched, the records before it are exact, out_offsets[n] reports the
    need; long lists take the staged (bulk-copy) kernel, short ones the plain one.  Unitig ids outside the table are empty lists."""
    import torch
    import ctypes as C
    k = 31
    for U, lo_mult in ((300, 5000), (3000, 1)):
        codes, accum = _gen.synthetic_unitigs(U, 68, k, seed=44)
        us = mz.UnitigSet(k, mz.pack_2bit(codes), len(codes), accum)
        g = mz.PFHash.from_unitig_set(us)
        o = OracleIndex.from_packed(k, us.useq_words, us.n_bases, accum, 0)
        n_refs, max_ref_len = 4096, 1 << 27
        offsets, ref_ids, poss, fws = _synthetic_u2pos(U, n_refs, max_ref_len, 44, 5000)
        if lo_mult > 1:  # make every list long so the launcher picks the staged kernel
            mult = np.maximum(np.diff(offsets), 200).astype(np.uint64)
            offsets = np.concatenate([[0], np.cumsum(mult)]).astype(np.uint64)
            n = int(offsets[-1])
            rng = np.random.default_rng(7)
            ref_ids, poss, fws = rng.integers(0, n_refs, n, dtype=np.uint64), rng.integers(0, max_ref_len - 4096, n, dtype=np.uint64), rng.integers(0, 2, n, dtype=np.uint64)
        o.attach_u2pos(0 if kind == "dense" else 1, offsets, ref_ids, poss, fws, max_ref_len, n_refs)
        off_vec = mz.PackedVec.pack(offsets)
        if kind == "dense":
            g.attach_u2pos_dense(((poss.astype(np.uint64) | (fws.astype(np.uint64) << np.uint64(31))) << np.uint64(32)) | ref_ids.astype(np.uint64), off_vec)
        else:
            enc = (ref_ids.astype(np.uint64) << np.uint64(29)) | (poss.astype(np.uint64) << np.uint64(1)) | fws.astype(np.uint64)
            g.attach_u2pos_piscem(mz.PackedVec.pack(enc, 42), 29, (1 << 28) - 1, off_vec)
        q = np.random.default_rng(3).integers(0, U, size=4000).astype(np.uint32)
        q[5] = U + 7          # outside the table: an empty list, not an out-of-bounds read
        q[9] = mz.MISS
        want_offs, want = o.decode_occs(np.where(q >= U, mz.MISS, q).astype(np.uint32))
        total = int(want_offs[-1])
        for cap in (total, total - 1, total // 2 + 3, 1, 0):
            guard = 4096
            d_q = torch.from_numpy(q.view(np.int32)).cuda()
            d_offs = torch.zeros(len(q) + 1, dtype=torch.int64, device="cuda")
            d_out = torch.full(((cap + guard), 3), -1, dtype=torch.int32, device="cuda")
            mz._check(mz.lib().mazu_b200_decode_occs(g._h, mz._any_ptr(d_q), len(q), mz._any_ptr(d_offs), mz._any_ptr(d_out), cap, None, mz.MEM_DEVICE, None))
            torch.cuda.synchronize()
            got = d_out.cpu().numpy().view(np.uint32)
            assert np.array_equal(d_offs.cpu().numpy().view(np.uint64), want_offs)
            assert (got[cap:] == 0xFFFFFFFF).all(), "records written beyond the capacity (cap %d of %d)" % (cap, total)
            assert np.array_equal(got[:cap].reshape(-1).view(mz.OCC_DTYPE), want[:cap])
        # with out_total the host learns the need and the call reports the error
        tot = C.c_uint64(0)
        rc = mz.lib().mazu_b200_decode_occs(g._h, mz._any_ptr(d_q), len(q), mz._any_ptr(d_offs), mz._any_ptr(d_out), total - 1, C.byref(tot), mz.MEM_DEVICE, None)
        assert rc == -7 and tot.value == total
        # mazu_b200_project_hits on a hit batch: short lists go by tiles of 128 records (tile totals -> scan over tiles -> emit),
        # long ones through the staged kernel.  Same contract: device buffers without out_total are clipped at cap, a sizing call
        # (out == NULL) fills the offsets, unitig ids outside the table and misses are empty lists, any batch length.
        rng = np.random.default_rng(11)
        for nh in (4000, 129, 128, 1):
            hits = np.zeros(nh, dtype=mz.HIT_DTYPE)
            hits["unitig_id"] = q[:nh]
            ulen = (accum[1:] - accum[:-1]).astype(np.uint32)
            inside = q[:nh] < U
            hits["unitig_len"] = np.where(inside, ulen[np.where(inside, q[:nh], 0)], 0)
            hits["pos"] = (rng.integers(0, 1 << 30, nh) % np.maximum(hits["unitig_len"].astype(np.int64) - k + 1, 1)).astype(np.uint32)
            hits["match"] = rng.integers(0, 4, nh).astype(np.uint32)  # NoMatch / Identity / Twin / skipped
            hits["match"][q[:nh] == mz.MISS] = 0
            ohits = hits.copy()
            ohits["match"][~inside] = 0  # the oracle is not asked about ids it would reject
            pw_offs, pw = o.project_hits(ohits)
            ptotal = int(pw_offs[-1])
            d_h = torch.from_numpy(hits.view(np.int32).reshape(-1, 4)).cuda()
            d_po = torch.zeros(nh + 1, dtype=torch.int64, device="cuda")
            tot = C.c_uint64(0)
            mz._check(mz.lib().mazu_b200_project_hits(g._h, mz._any_ptr(d_h), nh, mz._any_ptr(d_po), None, 0, C.byref(tot), mz.MEM_DEVICE, None))
            assert tot.value == ptotal and np.array_equal(d_po.cpu().numpy().view(np.uint64), pw_offs)
            for cap in (ptotal, ptotal // 2 + 1, 0):
                d_po.zero_()
                d_pout = torch.full(((cap + 4096), 3), -1, dtype=torch.int32, device="cuda")
                mz._check(mz.lib().mazu_b200_project_hits(g._h, mz._any_ptr(d_h), nh, mz._any_ptr(d_po), mz._any_ptr(d_pout), cap, None, mz.MEM_DEVICE, None))
                torch.cuda.synchronize()
                got = d_pout.cpu().numpy().view(np.uint32)
                assert np.array_equal(d_po.cpu().numpy().view(np.uint64), pw_offs)
                nw = min(cap, ptotal)
                assert (got[nw:] == 0xFFFFFFFF).all(), "projected records written beyond the capacity (cap %d of %d)" % (cap, ptotal)
                assert np.array_equal(got[:nw].reshape(-1).view(mz.OCC_DTYPE), pw[:nw])
            if ptotal > 1:
                rc = mz.lib().mazu_b200_project_hits(g._h, mz._any_ptr(d_h), nh, mz._any_ptr(d_po), mz._any_ptr(d_pout), ptotal - 1, C.byref(tot), mz.MEM_DEVICE, None)
                assert rc == -7 and tot.value == ptotal
            assert np.array_equal(g.project_hits(hits)[1], pw)  # host buffers


def test_fuzz_decode_tables():
    """Differential fuzzing of the occurrence decode / projection (K4): field widths from 12 to 60 bits (beyond 48 the staged
    kernel falls back to the generic tile path), lists from empty to thousands of entries so that tiles start and end at
    every alignment inside, at and across list boundaries, queries with repeats and misses, both encodings."""
    rng = np.random.default_rng(424242)
    k = 21
    for case in range(24):
        U = int(rng.integers(5, 400))
        codes, accum = _gen.synthetic_unitigs(U, 30, k, seed=1000 + case)
        us = mz.UnitigSet(k, mz.pack_2bit(codes), len(codes), accum)
        g = mz.SSHash.from_unitig_set(us, 11, 16)
        o = OracleIndex.from_packed(k, us.useq_words, us.n_bases, accum, 1, w=11, skew=16)
        kind = "dense" if case % 3 == 0 else "piscem"
        n_refs = int(2 ** rng.integers(1, 28)) if kind == "piscem" else int(rng.integers(1, 1 << 20))
        max_ref_len = int(2 ** rng.integers(8, 31)) if kind == "piscem" else (1 << 27)
        mult = rng.integers(0, 4, size=U).astype(np.uint64)                       # many short (and empty) lists ...
        heavy = rng.random(U) < (0.08 if case % 4 else 0.9)                          # ... and a few long ones (every 4th case: mostly
        mult[heavy] = rng.integers(200, 3000, size=int(heavy.sum()))                # long lists, which is what selects the staged kernel)
        offsets = np.zeros(U + 1, dtype=np.uint64)
        offsets[1:] = np.cumsum(mult)
        n = int(offsets[-1])
        ref_ids = rng.integers(0, n_refs, size=n, dtype=np.uint32)
        poss = rng.integers(0, max(1, max_ref_len - 2000), size=n, dtype=np.uint32)
        fws = rng.integers(0, 2, size=n, dtype=np.uint8)
        o.attach_u2pos(0 if kind == "dense" else 1, offsets, ref_ids, poss, fws, max_ref_len, n_refs)
        off_vec = mz.PackedVec.pack(offsets)
        if kind == "dense":
            words = ((poss.astype(np.uint64) | (fws.astype(np.uint64) << np.uint64(31))) << np.uint64(32)) | ref_ids.astype(np.uint64)
            g.attach_u2pos_dense(words, off_vec)
            what = "case %d dense U=%d occs=%d" % (case, U, n)
        else:
            out = np.zeros(3, dtype=np.uint32)
            assert O.lib().orc_required_num_bits(max_ref_len, n_refs, O._ptr(out)) == 0
            pos_bits, ref_bits, total = (int(x) for x in out)
            ref_shift, pos_mask = pos_bits + 1, (1 << pos_bits) - 1
            enc = (ref_ids.astype(np.uint64) << np.uint64(ref_shift)) | (poss.astype(np.uint64) << np.uint64(1)) | fws.astype(np.uint64)
            g.attach_u2pos_piscem(mz.PackedVec.pack(enc, total), ref_shift, pos_mask, off_vec)
            what = "case %d piscem width=%d U=%d occs=%d" % (case, total, U, n)
        q = rng.integers(0, U, size=int(rng.integers(1, 3000))).astype(np.uint32)
        q[rng.random(len(q)) < 0.05] = mz.MISS
        if heavy.any():  # make sure long lists are queried, several times in a row
            hv = np.nonzero(heavy)[0].astype(np.uint32)
            q = np.concatenate([q, np.repeat(hv[:4], 3), q[:50]])
        a, b = g.decode_occs(q), o.decode_occs(q)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), what
        hits = np.zeros(len(q), dtype=mz.HIT_DTYPE)
        ulen = np.diff(accum).astype(np.uint32)
        valid = q != mz.MISS
        hits["unitig_id"] = q
        hits["unitig_len"][valid] = ulen[q[valid]]
        hits["pos"][valid] = (rng.random(int(valid.sum())) * (ulen[q[valid]] - k + 1)).astype(np.uint32)
        hits["match"] = np.where(valid, rng.integers(1, 3, size=len(q)), 0)
        a, b = g.project_hits(hits), o.project_hits(hits)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), what + " (projection)"


# --------------------------------------------------------------------------------------------
# larger synthetic index: size-independent properties + sampled oracle comparison
# --------------------------------------------------------------------------------------------
def test_synthetic_index_properties():
    k = 31
    codes, accum = _gen.synthetic_unitigs(60000, 68, k, seed=45)  # ~6e6 bases
    us = mz.UnitigSet(k, mz.pack_2bit(codes), len(codes), accum)
    g = mz.SSHash.from_unitig_set(us, 19, 64)
    c = g.k2u_validate_self()  # every unitig k-mer, fw and swapped, maps back to itself
    assert c[0] == 2 * g.n_kmers and c[4] == 0 and c[1] == c[2] == g.n_kmers
    # reads sampled from the unitig concatenation: every hit must verify against the sequence
    bases = _gen.sample_reads_fast(codes, 50000, 150, seed=46, frac_ref=0.7, sub_rate=0.01)
    hr, cr, _ = g.query_reads(bases, None, uniform_read_len=150)
    hs, cs, _ = g.as_streaming().query_reads(bases, None, uniform_read_len=150)
    assert_hits_equal(hs, hr, "streaming == random on an index with unique canonical k-mers")
    assert list(cr) == list(cs) and cr[0] == 50000 * 120
    hit = hr["match"] != 0
    assert 0.2 < hit.mean() < 0.8
    starts = accum[hr["unitig_id"][hit].astype(np.int64)] + hr["pos"][hit]
    read_codes = np.frombuffer(bases, dtype=np.uint8)
    lut = np.zeros(256, dtype=np.uint8)
    for i, ch in enumerate(b"ACGT"):
        lut[ch] = i
    rk = _gen.kmer_words_from_codes(lut[read_codes], k).reshape(-1)
    # slot i of read r starts at base r*150 + p
    slot = np.nonzero(hit)[0]
    base_idx = (slot // 120) * 150 + (slot % 120)
    fw = rk[base_idx]
    uk = _gen.kmer_words_from_codes(codes, k)[starts.astype(np.int64)]
    rc = np.array([O.lib().orc_revcomp(int(x), k) for x in fw[:5000]], dtype=np.uint64)
    ident = hr["match"][hit] == mz.IDENTITY_MATCH
    assert np.array_equal(uk[ident], fw[ident])
    assert np.array_equal(uk[:5000][~ident[:5000]], rc[~ident[:5000]])
    # sampled oracle comparison
    o = OracleIndex.from_packed(k, us.useq_words, us.n_bases, accum, 1, w=19, skew=64)
    want, _, _ = o.query_reads(bases[:150 * 3000], np.arange(3001, dtype=np.uint64) * 150)
    assert_hits_equal(hr[:120 * 3000], want, "sampled oracle comparison")


def test_full_size_config2_properties(yeast_sshash, yeast_queries):
    """BASELINE.json configs[1]/[2] at FULL size (10 M reads x 150 bp = 1.2e9 lookups, far more than the oracle can
    replay in a test): size-independent properties instead.  (1) counters add up; (2) every k-mer of an error-free
    reference read hits and no k-mer of a uniform-random read does, so n_hit is known exactly from the generator;
    (3) the streaming walk equals random access on an index with unique canonical k-mers; (4) every hit of a sampled
    slice verifies against the unitig sequence; (5) the slice equals the oracle."""
    import torch
    g, o = yeast_sshash
    _, ref_codes = yeast_queries
    k, L, n = g.k, 150, 10_000_000
    nk = L - k + 1
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(4242)
    ref_t = torch.from_numpy(ref_codes.astype(np.uint8)).to(dev)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    ar = torch.arange(L, device=dev)
    bases = torch.empty(n * L, dtype=torch.uint8, device=dev)
    n_ref = 0
    for r0 in range(0, n, 1_000_000):
        m = min(1_000_000, n - r0)
        starts = torch.randint(0, len(ref_codes) - L + 1, (m,), generator=gen, device=dev)
        codes = ref_t[starts[:, None] + ar[None, :]]
        strand = torch.rand(m, generator=gen, device=dev) < 0.5
        codes = torch.where(strand[:, None], 3 - codes.flip(1), codes)
        is_ref = torch.rand(m, generator=gen, device=dev) < 0.5
        rnd = torch.randint(0, 4, (m, L), generator=gen, device=dev, dtype=torch.uint8)
        codes = torch.where(is_ref[:, None], codes, rnd)
        n_ref += int(is_ref.sum().item())
        bases[r0 * L:(r0 + m) * L] = acgt[codes.long()].reshape(-1)
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    for mode in (mz.MODE_RANDOM, mz.MODE_STREAMING):
        hits = torch.empty((n * nk, 4), dtype=torch.int32, device=dev)
        cnt = torch.zeros(3, dtype=torch.int64, device=dev)
        g.query_reads(bases, None, n_reads=n, uniform_read_len=L, mode=mode, out_hits=hits, counts=cnt, mem=mz.MEM_DEVICE, stream=st)
        torch.cuda.synchronize()
        c = cnt.cpu().numpy()
        assert c[0] == n * nk and c[1] + c[2] == c[0]                     # (1)
        assert c[1] == n_ref * nk                                          # (2)
        res[mode] = hits
    assert torch.equal(res[mz.MODE_RANDOM], res[mz.MODE_STREAMING])        # (3)
    hits = res[mz.MODE_RANDOM]
    assert int((hits[:, 3] == mz.IDENTITY_MATCH).sum().item() + (hits[:, 3] == mz.TWIN_MATCH).sum().item()) == n_ref * nk
    # (4) + (5) on the last 20,000 reads
    m = 20000
    h = hits[(n - m) * nk:].cpu().numpy().view(np.uint32).reshape(-1).view(mz.HIT_DTYPE)
    b = bases[(n - m) * L:].cpu().numpy()
    want, _, _ = o.query_reads(b, np.arange(m + 1, dtype=np.uint64) * L)
    assert_hits_equal(h, want, "tail slice of the full-size batch")
    lut = np.zeros(256, dtype=np.uint8)
    for i, ch in enumerate(b"ACGT"):
        lut[ch] = i
    rk = _gen.kmer_words_from_codes(lut[b], k)
    useq_codes = _gen.unpack_2bit(o.useq_words(), o.total_len)
    uk = _gen.kmer_words_from_codes(useq_codes, k)
    accum = np.array([0] + list(np.cumsum([g.unitig_len(u) for u in range(g.n_unitigs)])), dtype=np.uint64)
    slot = np.nonzero(h["match"] == mz.IDENTITY_MATCH)[0]
    fw = rk[(slot // nk) * L + (slot % nk)]
    assert len(slot) > 0 and np.array_equal(uk[(accum[h["unitig_id"][slot].astype(np.int64)] + h["pos"][slot]).astype(np.int64)], fw)


# --------------------------------------------------------------------------------------------
# other k / w, including the maximum k = 32 (even k: palindromic k-mers have fw == rc)
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,w,skew", [(32, 20, 8), (32, 32, NOSKEW), (21, 11, 4), (15, 7, 0), (5, 3, 2), (31, 1, 64)])
def test_other_k_w_synthetic(k, w, skew):
    codes, accum = _gen.synthetic_unitigs(400, 40, k, seed=100 + k)
    if k <= 7:  # tiny k: random unitigs repeat k-mers; still must match the oracle's first-match order
        codes, accum = _gen.synthetic_unitigs(12, 6, k, seed=100 + k)
    us = mz.UnitigSet(k, mz.pack_2bit(codes), len(codes), accum)
    g = mz.SSHash.from_unitig_set(us, w, skew, seed=3)
    o = OracleIndex.from_packed(k, us.useq_words, us.n_bases, accum, 1, w=w, skew=skew, seed=3)
    gp = mz.PFHash.from_unitig_set(us)
    assert g.k2u_validate_self() == o.k2u_validate_self()
    bases, offs = _gen.sample_reads(codes, 600, 180, seed=k, frac_ref=0.7, sub_rate=0.02, n_rate=0.003, ragged=True)
    for mode in (mz.MODE_RANDOM, mz.MODE_STREAMING):
        want, _ = _check_reads(g, o, bases, offs, mode)
    if k >= 21:  # canonical k-mers are unique at these sizes: PFHash (native MPHF) must give the same answers as SSHash
        got, _, _ = gp.query_reads(bases, offs)
        assert_hits_equal(got, o.query_reads(bases, offs)[0], "PFHash native on synthetic k=%d" % k)
    rng = np.random.default_rng(k)
    q = np.concatenate([_gen.kmer_words_from_codes(codes, k), rng.integers(0, 1 << 62, size=2000, dtype=np.uint64) & np.uint64((1 << (2 * k)) - 1 if k < 32 else 0xFFFFFFFFFFFFFFFF)])
    assert_hits_equal(g.k2u_batch(q), o.k2u_batch(q), "k2u_batch k=%d w=%d" % (k, w))


def test_fuzz_small_indexes():
    """Differential fuzzing with a fixed seed: 150 (MAZU_FUZZ_CASES) random small unitig sets (any k in [3, 32], any w <= k, skew thresholds
    from 0 to none, a handful to a few hundred unitigs, low-complexity and repeated sequence so that canonical k-mers
    collide, heavy buckets form and the skew index is used), queried with ragged reads with substitutions and Ns in both
    modes, with flat k-mer batches, through the host and the device builder.  Everything must equal the oracle."""
    rng = np.random.default_rng(20261018)
    n_cases = int(os.environ.get("MAZU_FUZZ_CASES", "150"))
    for case in range(n_cases):
        k = int(rng.integers(3, 33))
        w = int(rng.integers(1, k + 1))
        skew = [0, 1, 2, 8, 64, NOSKEW][int(rng.integers(0, 6))]
        n_unitigs = int(rng.integers(1, 200 if k > 10 else 12))
        alphabet = int(rng.integers(1, 5))  # 1..4 letters: low complexity makes duplicates and heavy buckets
        lens = k + rng.integers(0, 60, size=n_unitigs)
        accum = np.zeros(n_unitigs + 1, dtype=np.uint64)
        accum[1:] = np.cumsum(lens)
        codes = rng.integers(0, alphabet, size=int(accum[-1]), dtype=np.uint8)
        if rng.random() < 0.3 and n_unitigs > 1:  # repeat a unitig verbatim: every one of its k-mers is duplicated
            a, b = int(accum[0]), int(accum[1])
            ln = min(b - a, int(lens[-1]))
            codes[int(accum[-2]):int(accum[-2]) + ln] = codes[a:a + ln]
        us = mz.UnitigSet(k, mz.pack_2bit(codes), len(codes), accum)
        seed = int(rng.integers(0, 1 << 30))
        what = "case %d: k=%d w=%d skew=%s unitigs=%d alphabet=%d" % (case, k, w, skew, n_unitigs, alphabet)
        o = OracleIndex.from_packed(k, us.useq_words, us.n_bases, accum, 1, w=w, skew=skew, seed=seed)
        for builder in ("host", "gpu"):
            g = mz.SSHash.from_unitig_set(us, w, skew, seed=seed, builder=builder)
            assert g.k2u_validate_self() == o.k2u_validate_self(), what
            bases, offs = _gen.sample_reads(codes, 60, 3 * k + 40, seed=case, frac_ref=0.7, sub_rate=0.02, n_rate=0.004, ragged=True)
            for mode in (mz.MODE_RANDOM, mz.MODE_STREAMING):
                want, wcnt, wk = o.query_reads(bases, offs, streaming=mode == mz.MODE_STREAMING, reset_per_read=True)
                got, gcnt, gk = g.query_reads(bases, offs, mode=mode)
                assert_hits_equal(got, want, what + " mode %d builder %s" % (mode, builder))
                assert list(gcnt) == list(wcnt), what
            q = np.concatenate([_gen.kmer_words_from_codes(codes, k),
                                rng.integers(0, 1 << 62, size=300, dtype=np.uint64) & np.uint64((1 << (2 * k)) - 1 if k < 32 else 0xFFFFFFFFFFFFFFFF)])
            assert_hits_equal(g.k2u_batch(q), o.k2u_batch(q), what + " k2u_batch builder " + builder)


def test_palindromic_kmers_even_k():
    """even k: a palindromic k-mer equals its reverse complement; the reference reports IdentityMatch first."""
    k = 8
    seqs = ["ACGTACGTTTGACCA", "GGATCCGGAAGCTTCA", "TTTTAAAACCCCGGGG"]  # ACGTACGT, GGATCC.., AAGCTT.. contain palindromes
    us = mz.UnitigSet.from_seqs(seqs, k)
    for w, skew in [(4, NOSKEW), (3, 0)]:
        g = mz.SSHash.from_unitig_set(us, w, skew)
        o = OracleIndex.from_seqs(seqs, k, 1, w=w, skew=skew)
        codes = _gen.unpack_2bit(us.useq_words, us.n_bases)
        q = _gen.kmer_words_from_codes(codes, k)
        rc = np.array([O.lib().orc_revcomp(int(x), k) for x in q], dtype=np.uint64)
        assert (q == rc).any(), "test needs at least one palindromic k-mer"
        both = np.concatenate([q, rc])
        assert_hits_equal(g.k2u_batch(both), o.k2u_batch(both), "palindromes")


def test_compact_records_equal_full_records(yeast_sshash, yeast_queries):
    """mazu_b200_query_reads_compact: the 8-byte record carries the same answer (unitig_len is looked up by id)."""
    g, o = yeast_sshash
    _, ref_codes = yeast_queries
    bases, offs = _gen.sample_reads(ref_codes, 3000, 200, seed=8, frac_ref=0.6, sub_rate=0.02, n_rate=0.002, ragged=True)
    for mode in (mz.MODE_RANDOM, mz.MODE_STREAMING):
        full, cnt, _ = g.query_reads(bases, offs, mode=mode)
        comp, cnt8, _ = g.query_reads(bases, offs, mode=mode, compact=True)
        assert list(cnt) == list(cnt8)
        assert np.array_equal(comp["unitig_id"], full["unitig_id"])
        assert np.array_equal(comp["pos_match"] >> 30, full["match"])
        assert np.array_equal(comp["pos_match"] & 0x3FFFFFFF, full["pos"] & 0x3FFFFFFF)
        hit = full["match"] != mz.NO_MATCH
        hit &= full["match"] != mz.SKIPPED
        lens = np.array([g.unitig_len(int(u)) for u in np.unique(comp["unitig_id"][hit])])
        assert len(lens) > 0


# --------------------------------------------------------------------------------------------
# boundary: descriptors, error codes, concurrency
# --------------------------------------------------------------------------------------------
def _parse_boophf(path):
    """mphf.bin -> (levels words, n_bits, last_bitset_rank, n_elem, final keys, final vals); src/pf1/boophf/mod.rs:50-86,269-293"""
    import struct
    b = open(path, "rb").read()
    p = 0
    gamma, = struct.unpack_from("<d", b, p); p += 8
    nl, = struct.unpack_from("<i", b, p); p += 4
    lbr, n_elem = struct.unpack_from("<QQ", b, p); p += 16
    words, nbits = [], []
    for _ in range(nl):
        nb, nw = struct.unpack_from("<QQ", b, p); p += 16
        words.append(np.frombuffer(b, dtype="<u8", count=nw, offset=p).copy()); p += 8 * nw
        rs, = struct.unpack_from("<Q", b, p); p += 8 + 8 * rs
        nbits.append(nb)
    nf, = struct.unpack_from("<Q", b, p); p += 8
    kv = np.frombuffer(b, dtype="<u8", count=2 * nf, offset=p).reshape(-1, 2) if nf else np.zeros((0, 2), dtype=np.uint64)
    return words, np.array(nbits, dtype=np.uint64), lbr, n_elem, kv[:, 0].copy(), kv[:, 1].copy()


def test_from_parts_descriptors(yeast_dense, yeast_queries):
    """ModIndex::from_parts through plain descriptors: UnitigSet + C++ BooPHF + pos vector + U2Pos + refseq, as a
    Rust host would pass them (PFHash::from_parts, src/kphf/pfhash.rs:34-36; src/index.rs:80-87)."""
    import ctypes as C
    g, o = yeast_dense
    q, _ = yeast_queries
    us = mz.UnitigSet(o.k, o.useq_words(), o.total_len, o.unitig_starts())
    words, nbits, lbr, n_elem, fk, fv = _parse_boophf(os.path.join(YEAST_CHR01, "mphf.bin"))
    ptrs = (C.c_void_p * len(words))(*[w.ctypes.data for w in words])
    bd = mz.BooPHFDesc(len(words), ptrs, nbits.ctypes.data_as(C.c_void_p), lbr, n_elem, fk.ctypes.data_as(C.c_void_p),
                       fv.ctypes.data_as(C.c_void_p), len(fk))
    w, n = C.c_uint64(0), C.c_uint64(0)
    assert O.lib().orc_compact_vector_read(os.path.join(YEAST_CHR01, "pos.bin").encode(), C.byref(w), C.byref(n), None, 0) == 0
    vals = np.zeros(n.value, dtype=np.uint64)
    O.lib().orc_compact_vector_read(os.path.join(YEAST_CHR01, "pos.bin").encode(), C.byref(w), C.byref(n), O._ptr(vals), n.value)
    pos = mz.PackedVec.pack(vals, w.value)
    out = C.c_void_p(0)
    ud, pd = us.desc(), pos.desc()
    mz._check(mz.lib().mazu_b200_index_create_pfhash_from_parts(C.byref(ud), C.byref(bd), C.byref(pd), 0, C.byref(out)))
    idx = mz.ModIndex(out.value)
    assert_hits_equal(idx.k2u_batch(q[:100000]), o.k2u_batch(q[:100000]), "from_parts")
    # U2Pos + refseq attached from decoded parts -> validate_self as the loader-built index
    offs, occs = o.decode_occs(np.arange(o.n_unitigs, dtype=np.uint32))
    ctable = ((occs["pos"].astype(np.uint64) | (occs["fw"].astype(np.uint64) << np.uint64(31))) << np.uint64(32)) | occs["ref_id"].astype(np.uint64)
    idx.attach_u2pos_dense(ctable, mz.PackedVec.pack(offs))
    with pytest.raises(mz.MazuError) as e:
        idx.validate_self()
    assert e.value.code == mz.ERR_NO_REFSEQ
    idx.attach_refseq(o.refseq_words(), o.ref_prefix())
    assert idx.validate_self() == [230188, 170689, 59499, 262130, 0]


def test_error_codes():
    import ctypes as C
    with pytest.raises(mz.MazuError) as e:
        mz.DenseIndex.deserialize_from_cpp(os.path.join(PF1, "does_not_exist"))
    assert e.value.code == -1  # MAZU_ERR_IO
    us = mz.UnitigSet.from_seqs(["ACGTACGTAC", "TTGACCATGA"], 7)
    bad = mz.UnitigSet(7, us.useq_words, us.n_bases, np.array([0, 12, 10, 20], dtype=np.uint64))
    with pytest.raises(mz.MazuError) as e:
        mz.SSHash.from_unitig_set(bad, 3, mz.SKEW_NONE)
    assert e.value.code == -3  # MAZU_ERR_EF_NOT_MONOTONE (EFVector::from_iter, src/elias_fano.rs:89-91)
    with pytest.raises(mz.MazuError) as e:
        mz.SSHash.from_unitig_set(us, 9, mz.SKEW_NONE)  # assert!(w <= k), src/kphf/sshash.rs:92
    assert e.value.code == -7
    with pytest.raises(mz.MazuError):
        mz.UnitigSet.from_seqs(["ACGTNACGT"], 3)
    idx = mz.SSHash.from_unitig_set(us, 3, mz.SKEW_NONE)
    with pytest.raises(mz.MazuError) as e:
        idx.query_reads(np.frombuffer(b"ACGTACGTAC", dtype=np.uint8), np.array([0, 10], dtype=np.uint64), mode=7)
    assert e.value.code == -7


def test_cuda_path_reproduces_golden_fixture():
    """tests/golden/yeast_chr01_queries.npz: committed answers for a seeded batch; compared here without loading the oracle"""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "yeast_chr01_queries.npz"))
    bases, offs = g["bases"], g["read_offsets"]
    dense = mz.DenseIndex.deserialize_from_cpp(YEAST_CHR01)
    assert dense.validate_self() == [int(x) for x in g["validate_self"]]
    ss = dense.rebuild_k2u(mz.K2U_SSHASH, w=15, skew_param=32, seed=0)
    ss_gpu = None
    for name, ix in (("pfhash", dense), ("sshash", ss)):
        for mode in (mz.MODE_RANDOM, mz.MODE_STREAMING):
            hits, cnt, koffs = ix.query_reads(bases, offs, mode=mode)
            tag = "%s_%s" % (name, "streaming" if mode == mz.MODE_STREAMING else "random")
            assert np.array_equal(hits.view(np.uint32).reshape(-1, 4), g[tag + "_hits"]), tag
            assert [int(x) for x in cnt] == [int(x) for x in g[tag + "_counts"]]
            assert np.array_equal(koffs, g["kmer_offsets"])
    hits, _, _ = ss.query_reads(bases, offs)
    po, pr = ss.project_hits(hits)
    assert np.array_equal(po, g["project_offsets"]) and np.array_equal(pr.view(np.uint32).reshape(-1, 3), g["project_records"])
    do, dr = dense.decode_occs(g["decode_unitig_ids"])
    assert np.array_equal(do, g["decode_offsets"]) and np.array_equal(dr.view(np.uint32).reshape(-1, 3), g["decode_records"])


def test_c_example_runs(tmp_path):
    """examples/query_reads.c: a plain C caller of the ABI end to end (load, rebuild K2U, validate_self, query both modes, project)"""
    import subprocess
    from test_cpu_product import _build_c_example
    exe = _build_c_example(tmp_path)
    r = subprocess.run([exe, YEAST_CHR01], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "validate_self: 230188 queries, 170689 identity, 59499 twin, 262130 projected, 0 failures" in r.stdout
    # the oracle's answer for the two reads of the example: 32 + 1 valid windows (the N invalidates 31 of 32), the poly-T 31-mer hits
    assert "mode 0: 33 k-mers, 1 hits, 32 misses, 64 slots" in r.stdout
    assert "mode 1: 33 k-mers, 1 hits, 32 misses, 64 slots" in r.stdout
    assert "projected reference positions: 6" in r.stdout


@pytest.mark.parametrize("read_len", [150, 170])
@pytest.mark.parametrize("ragged", [True, False])
def test_hit_runs_expand_to_the_exact_records(yeast_dense, yeast_sshash, yeast_queries, ragged, read_len):
    """mazu_b200_query_reads_runs + mazu_b200_expand_hit_runs == mazu_b200_query_reads == oracle, record for record; the run
    format is far smaller than the records it stands for; a too-small run buffer reports the capacity it needs.
    150 bp reads fit one chunk and take the fused kernel (lookups + run encoding in one pass), 170 bp reads the four-kernel chain."""
    import ctypes as C
    _, ref_codes = yeast_queries
    bases, offs = _gen.sample_reads(ref_codes, 5000, read_len, seed=51, frac_ref=0.7, sub_rate=0.01, n_rate=0.002, ragged=ragged)
    for g, o in (yeast_sshash, yeast_dense):
        for mode in (mz.MODE_RANDOM, mz.MODE_STREAMING):
            want, wcnt, wk = o.query_reads(bases, offs, streaming=mode == mz.MODE_STREAMING, reset_per_read=True)
            kw = dict(read_offsets=offs) if ragged else dict(uniform_read_len=read_len)
            codes, runs, rro, cnt, koffs = g.query_reads_runs(bases, mode=mode, **kw)
            assert list(cnt) == list(wcnt)
            assert len(codes) == len(want) and rro[-1] == len(runs) and int(rro.max()) <= len(runs)
            if ragged:
                assert np.array_equal(koffs, wk)
                got = mz.ModIndex.expand_hit_runs(codes, runs, rro, kmer_offsets=koffs)
            else:
                got = mz.ModIndex.expand_hit_runs(codes, runs, rro, uniform_slots=read_len - g.k + 1)
            assert_hits_equal(got, want, "expanded runs, mode %d" % mode)
            n_hit = int(((want["match"] == mz.IDENTITY_MATCH) | (want["match"] == mz.TWIN_MATCH)).sum())
            assert int((codes == 2).sum()) == len(runs) and int(((codes == 1) | (codes == 2)).sum()) == n_hit
            assert len(codes) + 16 * len(runs) + 8 * len(rro) < 0.2 * 16 * len(want)  # the point of the format
            # pinned (device-addressable) run buffer: the sync-free path stores the records straight into it -- same bytes
            pr = mz.PinnedArray((len(runs) + 7,), mz.HIT_DTYPE)
            codes2, runs2, rro2, cnt2, _ = g.query_reads_runs(bases, mode=mode, runs=pr.array, **kw)
            assert np.array_equal(codes2, codes) and list(cnt2) == list(cnt) and len(runs2) == len(runs)
            # (where a read's runs sit in the array is unspecified: compare what they expand to, and the multiset of records)
            got2 = mz.ModIndex.expand_hit_runs(codes2, runs2, rro2, kmer_offsets=koffs) if ragged else mz.ModIndex.expand_hit_runs(codes2, runs2, rro2, uniform_slots=read_len - g.k + 1)
            assert_hits_equal(got2, want, "expanded runs (pinned buffers), mode %d" % mode)
            assert np.array_equal(np.sort(runs2.view(np.uint32).reshape(-1, 4), axis=0), np.sort(runs.view(np.uint32).reshape(-1, 4), axis=0))
    # capacity: one run record is not enough; the call says how many it needs
    g, _ = yeast_sshash
    n_reads = len(offs) - 1
    n_runs = C.c_uint64(0)
    codes = np.empty(g.count_kmer_slots(offs, n_reads, 0), dtype=np.uint8)
    rc = mz.lib().mazu_b200_query_reads_runs(g._h, mz._np_ptr(bases), mz._np_ptr(offs), n_reads, 0, mz.MODE_RANDOM, None, mz._any_ptr(codes),
                                             mz._any_ptr(np.empty(1, dtype=mz.HIT_DTYPE)), 1, mz._any_ptr(np.zeros(n_reads + 1, dtype=np.uint64)),
                                             C.byref(n_runs), None)
    assert rc == -7 and n_runs.value > 1


def test_packed_reads_in_packed_codes_out(yeast_sshash, yeast_dense, yeast_queries):
    """mazu_b200_query_reads_runs_packed: 2-bit packed reads (+ N mask) over PCIe, 2-bit codes back; expands to exactly the
    records of the ASCII call and of the oracle, N windows and lower case included; both modes, several pipeline chunks."""
    _, ref_codes = yeast_queries
    n_reads, read_len = 6000, 150
    bases, offs = _gen.sample_reads(ref_codes, n_reads, read_len, seed=61, frac_ref=0.7, sub_rate=0.01, n_rate=0.003)
    words, mask, bad = mz.pack_reads(bases, read_len)
    assert bad == int(((bases == ord("N")) | (bases == ord("n"))).sum()) > 0
    os.environ["MAZU_B200_CHUNK_MIB"] = "1"  # several chunks
    try:
        for g, o in (yeast_sshash, yeast_dense):
            for mode in (mz.MODE_RANDOM, mz.MODE_STREAMING):
                want, wcnt, _ = o.query_reads(bases, offs, streaming=bool(mode))
                codes2, runs, rro, cnt = g.query_reads_runs_packed(words, mask, n_reads, read_len, mode=mode)
                got = mz.ModIndex.expand_hit_runs_packed(codes2, runs, rro, read_len - g.k + 1)
                assert_hits_equal(got, want, "packed runs mode %d" % mode)
                assert list(cnt) == list(wcnt) and len(codes2) == (len(want) + 3) // 4
    finally:
        del os.environ["MAZU_B200_CHUNK_MIB"]
    # other read lengths: 1-4 packed words and 1-2 mask words per read (the fused kernel reads them in place), a last word
    # that is only partly inside the read
    g, o = yeast_sshash
    for rl in (46, 66, 102, 130):
        b2, o2 = _gen.sample_reads(ref_codes, 700, rl, seed=62 + rl, frac_ref=0.8, sub_rate=0.01, n_rate=0.004)
        w2, m2, _ = mz.pack_reads(b2, rl)
        want, wcnt, _ = o.query_reads(b2, o2)
        codes2, runs, rro, cnt = g.query_reads_runs_packed(w2, m2, 700, rl)
        assert_hits_equal(mz.ModIndex.expand_hit_runs_packed(codes2, runs, rro, rl - g.k + 1), want, "packed runs, read length %d" % rl)
        assert list(cnt) == list(wcnt)
    # without a mask the N positions read as 'A': the caller's contract, not checked here; wrong shapes are rejected
    g, _ = yeast_sshash
    with pytest.raises(mz.MazuError):
        g.query_reads_runs_packed(words, mask, n_reads, 151)  # 121 slots per read: not a multiple of 4


def test_hit_intervals_expand_to_the_exact_records(yeast_sshash, yeast_dense, yeast_queries):
    """mazu_b200_query_reads_intervals_packed: one 16-byte record per hit run and nothing else; with the caller's N mask it
    expands to exactly the oracle's records -- several read lengths, N windows, both index kinds, pageable and pinned output
    buffers, several pipeline chunks, capacity too small."""
    import ctypes as C

    _, ref_codes = yeast_queries
    os.environ["MAZU_B200_CHUNK_MIB"] = "1"
    recs = lambda a: sorted(a.tobytes()[16 * i:16 * i + 16] for i in range(len(a)))  # the order of the records is unspecified
    try:
        for rl, n_reads in ((150, 9000), (158, 3000), (100, 1500), (47, 1200), (31, 500)):
            bases, offs = _gen.sample_reads(ref_codes, n_reads, rl, seed=90 + rl, frac_ref=0.75, sub_rate=0.01, n_rate=0.003)
            bases = bases.copy()
            bases[::7] |= 0x20  # lower case is the same base (src/pf1/dense_index.rs:180-183)
            words, mask, _ = mz.pack_reads(bases, rl)
            for g, o in (yeast_sshash, yeast_dense):
                want, wcnt, _ = o.query_reads(bases, offs)
                for mode in (mz.MODE_RANDOM, mz.MODE_STREAMING):  # unique k-mers: streaming answers are k2u's
                    iv, cnt = g.query_reads_intervals_packed(words, mask, n_reads, rl, mode=mode)
                    assert_hits_equal(g.expand_hit_intervals(iv, mask, n_reads, rl), want, "intervals, read length %d mode %d" % (rl, mode))
                    assert list(cnt) == list(wcnt)
                    # the same records from the ASCII reads as the reference takes them (lower case, N and all): no packing by the caller
                    iva, cnta = g.query_reads_intervals(bases, n_reads, rl, mode=mode)
                    assert list(cnta) == list(wcnt) and recs(iva) == recs(iv)
                    assert_hits_equal(g.expand_hit_intervals_ascii(iva, bases, n_reads, rl), want, "ASCII intervals, read length %d mode %d" % (rl, mode))
                    # maximal runs: one record per run start of the run format
                    slots = rl - g.k + 1
                    hit = (want["match"] == 1) | (want["match"] == 2)
                    assert int(iv["len"].sum()) == int(hit.sum()) and int(iv["len"].max(initial=0)) <= slots
        # pinned output buffer: records are published by the device, no per-chunk synchronisation
        g, o = yeast_sshash
        pin = mz.PinnedArray((len(iv) + 8,), mz.INTERVAL_DTYPE)
        iv2, _ = g.query_reads_intervals_packed(words, mask, n_reads, rl, intervals=pin.array)
        assert len(iv2) == len(iv) and recs(iv2) == recs(iv)
        pin.array[:] = 0
        iv3, _ = g.query_reads_intervals(bases, n_reads, rl, intervals=pin.array)
        assert len(iv3) == len(iv) and recs(iv3) == recs(iv)
        # too small a buffer: the call says how many records it needs and never writes past the capacity
        small = np.zeros(4, dtype=mz.INTERVAL_DTYPE)
        guard = small.copy()
        n = C.c_uint64(0)
        rc = mz.lib().mazu_b200_query_reads_intervals_packed(g._h, mz._any_ptr(words), mz._any_ptr(mask), n_reads, rl, 0, mz._any_ptr(small), 2, C.byref(n), None)
        assert rc == -7 and n.value == len(iv) and np.array_equal(small[2:], guard[2:])
        # long reads are not for this interface
        with pytest.raises(mz.MazuError):
            g.query_reads_intervals_packed(words, mask, 10, 400)
    finally:
        del os.environ["MAZU_B200_CHUNK_MIB"]


@pytest.mark.parametrize("w,skew", [(19, 64), (15, 32), (17, 32)])
def test_read_kernels_kw_specialisations(yeast_dense, yeast_queries, w, skew):
    """The SSHash read kernels exist twice for (k, w) = (31, 19) and (31, 15): with k and w folded into the code and with
    both read from the index (capi.cu: kw_code; (31, 17) has the second form only).  Every read interface must give the
    oracle's records through either: records, counts, streaming, hit intervals, reads -> MappedRefPos."""
    g0, o0 = yeast_dense
    g, o = g0.rebuild_k2u(mz.K2U_SSHASH, w=w, skew_param=skew, seed=0), o0.rebuild_k2u(1, w=w, skew=skew, seed=0)
    _, ref_codes = yeast_queries
    rag = _gen.sample_reads(ref_codes, 2500, 260, seed=70 + w, frac_ref=0.7, sub_rate=0.01, n_rate=0.002, ragged=True)
    uni = _gen.sample_reads(ref_codes, 4000, 150, seed=170 + w, frac_ref=0.7, sub_rate=0.01, n_rate=0.002)
    want_u, wcnt_u, _ = o.query_reads(*uni)
    woffs_u, wmrps_u = o.project_hits(want_u)
    words, mask, _ = mz.pack_reads(uni[0], 150)
    try:
        for generic in ("0", "1"):
            os.environ["MAZU_B200_GENERIC_KW"] = generic
            for mode in (mz.MODE_RANDOM, mz.MODE_STREAMING):
                _check_reads(g, o, rag[0], rag[1], mode)
                got, cnt, _ = g.query_reads(uni[0], None, uniform_read_len=150, mode=mode)
                want, wcnt, _ = o.query_reads(uni[0], uni[1], streaming=bool(mode), reset_per_read=True)
                assert_hits_equal(got, want, "uniform reads, w=%d generic=%s mode=%d" % (w, generic, mode))
                assert list(cnt) == list(wcnt)
            iv, cnt = g.query_reads_intervals_packed(words, mask, 4000, 150)
            assert_hits_equal(g.expand_hit_intervals(iv, mask, 4000, 150), want_u, "intervals, w=%d generic=%s" % (w, generic))
            assert list(cnt) == list(wcnt_u)
            iva, cnta = g.query_reads_intervals(uni[0], 4000, 150)
            assert_hits_equal(g.expand_hit_intervals_ascii(iva, uni[0], 4000, 150), want_u, "ASCII intervals, w=%d generic=%s" % (w, generic))
            hits, poffs, mrps, cnt, _ = g.get_ref_pos_reads(uni[0], uniform_read_len=150)
            assert_hits_equal(hits, want_u, "get_ref_pos_reads, w=%d generic=%s" % (w, generic))
            assert np.array_equal(poffs, woffs_u) and np.array_equal(mrps, wmrps_u) and list(cnt) == list(wcnt_u)
    finally:
        del os.environ["MAZU_B200_GENERIC_KW"]


def test_pinned_host_buffers(yeast_sshash, yeast_queries):
    """mazu_b200_alloc_pinned: page-locked buffers for callers that do not link CUDA; same answers as pageable numpy arrays"""
    g, o = yeast_sshash
    _, ref_codes = yeast_queries
    bases, offs = _gen.sample_reads(ref_codes, 3000, 150, seed=31, frac_ref=0.6, sub_rate=0.01, n_rate=0.001, ragged=False)
    want, wcnt, _ = o.query_reads(bases, offs)
    pb = mz.PinnedArray(bases.shape, np.uint8)
    pb.array[:] = bases
    ph = mz.PinnedArray((len(want),), mz.HIT_DTYPE)
    got, cnt, _ = g.query_reads(pb.array, uniform_read_len=150, out_hits=ph.array)
    assert got is ph.array
    assert_hits_equal(ph.array, want, "pinned buffers")
    assert list(cnt) == list(wcnt)


def test_two_devices_in_one_process(yeast_queries):
    """index replicated per device, reads sharded by batch, one host thread per device, no collective.
    Skipped on a single-GPU box (the driver's GPU tier); run with `gpurun --gpus 2`."""
    import threading
    if mz.device_count() < 2:
        pytest.skip("needs two GPUs")
    _, ref_codes = yeast_queries
    o = OracleIndex.dense_from_pf1(YEAST_CHR01).rebuild_k2u(1, w=15, skew=32, seed=0)
    bases, offs = _gen.sample_reads(ref_codes, 6000, 150, seed=41, frac_ref=0.6, sub_rate=0.01, n_rate=0.001, ragged=False)
    want, wcnt, _ = o.query_reads(bases, offs)
    wpo, wpr = o.project_hits(want)
    half = 3000
    res = [None, None]

    def work(dev):
        g = mz.DenseIndex.deserialize_from_cpp(YEAST_CHR01, device=dev).rebuild_k2u(mz.K2U_SSHASH, w=15, skew_param=32, seed=0)
        b = bases[dev * half * 150:(dev + 1) * half * 150]
        hits, cnt, _ = g.query_reads(b, uniform_read_len=150)
        po, pr = g.project_hits(hits)  # the staged decode kernel raises its shared-memory limit per device
        res[dev] = (hits, cnt, po, pr)

    th = [threading.Thread(target=work, args=(d,)) for d in (0, 1)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert res[0] is not None and res[1] is not None
    got = np.concatenate([res[0][0], res[1][0]])
    assert_hits_equal(got, want, "two devices")
    assert [int(a) + int(b) for a, b in zip(res[0][1], res[1][1])] == [int(x) for x in wcnt]  # the final host-side gather of counters
    pr = np.concatenate([res[0][3], res[1][3]])
    assert np.array_equal(pr.view(np.uint32).reshape(-1, 3), np.asarray(wpr).view(np.uint32).reshape(-1, 3))


def test_replicate_and_sharded_queries(yeast_sshash, yeast_queries):
    """mazu_b200_index_replicate + the sharded calls: replicas are bit-identical copies made device to device,
    a sharded call over them returns exactly the single-handle answers.  With one visible GPU both replicas live on device 0
    (the copy / pointer-rebasing / sharding / run-region logic is the same); with two or more they are on different devices."""
    g, o = yeast_sshash
    _, ref_codes = yeast_queries
    n_dev = mz.device_count()
    devices = [0, 1 % n_dev, 2 % n_dev]
    reps = g.replicate(devices)
    assert [r.device for r in reps] == devices
    for which in range(8):  # every named table is byte-identical on every replica
        assert all(r.table_digest(which) == g.table_digest(which) for r in reps)
    assert all(r.k2u_validate_self() == g.k2u_validate_self() for r in reps)
    for ragged in (True, False):
        bases, offs = _gen.sample_reads(ref_codes, 3001, 150, seed=11, frac_ref=0.7, sub_rate=0.01, n_rate=0.002, ragged=ragged)
        for mode in (mz.MODE_RANDOM, mz.MODE_STREAMING):
            want, wcnt, wko = o.query_reads(bases, offs, streaming=bool(mode))
            if ragged:
                got, cnt, ko = mz.query_reads_sharded(reps, bases, read_offsets=offs, mode=mode)
            else:
                got, cnt, ko = mz.query_reads_sharded(reps, bases, uniform_read_len=150, mode=mode)
            assert_hits_equal(got, want, "sharded ragged=%s mode=%d" % (ragged, mode))
            assert list(cnt) == list(wcnt) and np.array_equal(ko, wko)
            kw = dict(read_offsets=offs) if ragged else dict(uniform_read_len=150)
            codes, runs, rro, cnt2, ko2, n_runs = mz.query_reads_runs_sharded(reps, bases, mode=mode, **kw)
            exp = mz.ModIndex.expand_hit_runs(codes, runs, rro, kmer_offsets=ko2)
            assert_hits_equal(exp, want, "sharded runs ragged=%s mode=%d" % (ragged, mode))
            assert list(cnt2) == list(wcnt) and n_runs == int((codes == 2).sum())
            # pinned run buffers take the sync-free path
            pr = mz.PinnedArray(len(runs), mz.HIT_DTYPE)
            pc = mz.PinnedArray(len(codes), np.uint8)
            po = mz.PinnedArray(len(rro), np.uint64)
            c3, r3, o3, cnt3, ko3, n3 = mz.query_reads_runs_sharded(reps, bases, mode=mode, codes=pc.array, runs=pr.array, read_run_offsets=po.array, **kw)
            assert_hits_equal(mz.ModIndex.expand_hit_runs(c3, r3, o3, kmer_offsets=ko3), want, "sharded pinned runs")
            assert n3 == n_runs
    # a region too small for one shard's runs is reported with the capacity needed
    bases, offs = _gen.sample_reads(ref_codes, 900, 150, seed=12, frac_ref=1.0, sub_rate=0.05)
    with pytest.raises(mz.MazuError) as e:
        import ctypes as C
        codes = np.empty(g.count_kmer_slots(offs), dtype=np.uint8)
        runs = np.empty(6, dtype=mz.HIT_DTYPE)
        rro = np.zeros(len(offs), dtype=np.uint64)
        n_runs = C.c_uint64(0)
        mz._check(mz.lib().mazu_b200_query_reads_runs_sharded(mz._handle_array(reps), len(reps), mz._np_ptr(bases), mz._np_ptr(offs), len(offs) - 1, 0, 0, None,
                                                               mz._np_ptr(codes), mz._np_ptr(runs), len(runs), mz._np_ptr(rro), C.byref(n_runs), None))
    assert e.value.code == -7 and n_runs.value > 6
    # the dense (PFHash + U2Pos + references) handle replicates too: validate_self needs every group of tables
    gd = mz.DenseIndex.deserialize_from_cpp(YEAST_CHR01)
    rd = gd.replicate([n_dev - 1])[0]
    assert rd.validate_self() == gd.validate_self() == [230188, 170689, 59499, 262130, 0]


def test_scratch_pool_is_reused_and_released(yeast_sshash, yeast_queries):
    """host-buffer calls stage through the handle's pool: results stay exact call after call, and release_scratch gives the memory back"""
    g, o = yeast_sshash
    _, ref_codes = yeast_queries
    bases, offs = _gen.sample_reads(ref_codes, 4000, 150, seed=21, frac_ref=0.6, sub_rate=0.01, n_rate=0.0, ragged=False)
    want, wcnt, _ = o.query_reads(bases, offs)
    for _ in range(3):
        got, cnt, _ = g.query_reads(bases, uniform_read_len=150)
        assert_hits_equal(got, want, "repeat")
        assert list(cnt) == list(wcnt)
    assert g.release_scratch() > 0
    assert g.release_scratch() == 0
    got, cnt, _ = g.query_reads(bases, uniform_read_len=150)
    assert_hits_equal(got, want, "after release")


def test_concurrent_queries_on_one_handle(yeast_sshash, yeast_queries):
    """queries take &self and are Sync in the reference (src/kphf/mod.rs:69-72): one immutable handle, many host threads."""
    import threading
    g, o = yeast_sshash
    _, ref_codes = yeast_queries
    inputs = [_gen.sample_reads(ref_codes, 2000, 150, seed=50 + t, frac_ref=0.6, sub_rate=0.01) for t in range(6)]
    want = [o.query_reads(b, offs, streaming=bool(t % 2))[0] for t, (b, offs) in enumerate(inputs)]
    got = [None] * len(inputs)

    def work(t):
        for _ in range(3):
            got[t] = g.query_reads(inputs[t][0], inputs[t][1], mode=t % 2)[0]

    ths = [threading.Thread(target=work, args=(t,)) for t in range(len(inputs))]
    [t.start() for t in ths]
    [t.join() for t in ths]
    for t in range(len(inputs)):
        assert_hits_equal(got[t], want[t], "thread %d" % t)


# --------------------------------------------------------------------------------------------
# SampledPFHash -- the pufferfish sparse index (src/pf1/sparse_index.rs:145-192)
# --------------------------------------------------------------------------------------------
SMALL_TXOME_SPARSE = os.path.join(PF1, "small_txome_index_sparse")


def test_sparse_index_validate_and_parity():
    g = mz.SparseIndex.deserialize_from_cpp(SMALL_TXOME_SPARSE)
    o = OracleIndex.sparse_from_pf1(SMALL_TXOME_SPARSE)
    assert g.info(mz.INFO_SAMPLE_SIZE) == 9 and g.info(mz.INFO_EXTENSION_SIZE) == 4
    assert g.info(mz.INFO_K2U_KIND) == mz.K2U_SAMPLED_PFHASH
    c = g.k2u_validate_self()
    assert c == o.k2u_validate_self() and c[0] == 2 * 18902 and c[4] == 0
    v = g.validate_self()
    assert v == o.validate_self() and v[0] == 28112 and v[4] == 0
    # k-mers: every unitig k-mer both strands, boundary-straddling windows, random negatives
    codes = _gen.unpack_2bit(o.useq_words(), o.total_len)
    rng = np.random.default_rng(4)
    q = np.concatenate([_gen.kmer_words_from_codes(codes, o.k), rng.integers(0, 1 << 62, size=20000, dtype=np.uint64)])
    assert_hits_equal(g.k2u_batch(q), o.k2u_batch(q), "SampledPFHash k2u")
    ref_codes = _gen.unpack_2bit(o.refseq_words(), int(o.ref_prefix()[-1]))
    bases, offs = _gen.sample_reads(ref_codes, 3000, 200, seed=6, frac_ref=0.7, sub_rate=0.01, n_rate=0.002, ragged=True)
    for mode in (mz.MODE_RANDOM, mz.MODE_STREAMING):
        _check_reads(g, o, bases, offs, mode)
    # sshash_drop_in (sparse_index.rs:176-191)
    ss = g.rebuild_k2u(mz.K2U_SSHASH, w=2, skew_param=NOSKEW)
    assert ss.validate_self()[4] == 0


# --------------------------------------------------------------------------------------------
# ModIndex::iter_unitigs_on_ref (src/index.rs:363-424; golden: src/refseq.rs:280-309)
# --------------------------------------------------------------------------------------------
def test_iter_unitigs_on_ref():
    g = mz.DenseIndex.deserialize_from_cpp(TINY_REFS_INDEX)
    t0, t1 = g.iter_unitigs_on_ref(0), g.iter_unitigs_on_ref(1)
    assert list(t0["unitig_len"]) == [5, 8, 9, 8, 5] and list(t0["unitig_id"]) == [0, 1, 2, 3, 4]
    assert list(t1["unitig_len"]) == [5, 9, 9, 9, 5] and list(t1["unitig_id"]) == [0, 5, 2, 6, 4]
    for d, loader, oloader in [(YEAST_CHR01, mz.DenseIndex.deserialize_from_cpp, OracleIndex.dense_from_pf1),
                               (SMALL_TXOME, mz.DenseIndex.deserialize_from_cpp, OracleIndex.dense_from_pf1),
                               (SMALL_TXOME_SPARSE, mz.SparseIndex.deserialize_from_cpp, OracleIndex.sparse_from_pf1)]:
        gi, oi = loader(d), oloader(d)
        for r in range(oi.n_refs):
            a, b = gi.iter_unitigs_on_ref(r), oi.iter_unitigs_on_ref(r)
            assert np.array_equal(a, b), (d, r)
            # the tiles rebuild the reference length: sum(len - k + 1) + k - 1
            assert int((a["unitig_len"].astype(np.int64) - gi.k + 1).sum()) + gi.k - 1 == oi.ref_len(r)
    ss = mz.DenseIndex.deserialize_from_cpp(YEAST_CHR01).rebuild_k2u(mz.K2U_SSHASH, w=15, skew_param=32)
    assert np.array_equal(ss.iter_unitigs_on_ref(0), OracleIndex.dense_from_pf1(YEAST_CHR01).iter_unitigs_on_ref(0))
    with pytest.raises(mz.MazuError):
        mz.PiscemIndex.from_cf_prefix(TINY_CF, 3, NOSKEW).iter_unitigs_on_ref(0)  # no reference sequence


# --------------------------------------------------------------------------------------------
# index construction on the device (SSHashBuilder::from_unitig_set, src/kphf/sshash.rs:86-329)
# --------------------------------------------------------------------------------------------
TABLES = ["mphf blocks", "bucket-bound blocks", "bucket-bound exceptions", "packed positions", "skew mphf blocks", "skew positions", "mphf fallback keys"]


def _assert_same_tables(a, b, what):
    for i, name in enumerate(TABLES):
        assert a.table_digest(i) == b.table_digest(i), "%s: table '%s' differs between the host and the GPU builder" % (what, name)


@pytest.mark.parametrize("w,skew", [(15, 32), (15, NOSKEW), (19, 4), (9, 2), (31, 0), (3, 8)])
def test_gpu_builder_bit_identical_yeast(yeast_dense, yeast_queries, w, skew):
    _, o = yeast_dense
    q, ref_codes = yeast_queries
    us = mz.UnitigSet(o.k, o.useq_words(), o.total_len, o.unitig_starts())
    host = mz.SSHash.from_unitig_set(us, w, skew, seed=5)
    gpu = mz.SSHash.from_unitig_set(us, w, skew, seed=5, builder="gpu")
    for what in (mz.INFO_N_MINIMIZERS, mz.INFO_N_MINIMIZER_OCCS, mz.INFO_N_KMERS_IN_SKEW_INDEX, mz.INFO_MPHF_LEVELS):
        assert host.info(what) == gpu.info(what), what
    _assert_same_tables(host, gpu, "yeast w=%d skew=%s" % (w, skew))
    c = gpu.k2u_validate_self()
    assert c[0] == 443836 and c[4] == 0
    os_ = o.rebuild_k2u(1, w=w, skew=skew, seed=5)
    assert_hits_equal(gpu.k2u_batch(q[:100000]), os_.k2u_batch(q[:100000]), "GPU-built index vs oracle")
    bases, offs = _gen.sample_reads(ref_codes, 1500, 200, seed=77, frac_ref=0.7, sub_rate=0.01, n_rate=0.002, ragged=True)
    for mode in (mz.MODE_RANDOM, mz.MODE_STREAMING):
        _check_reads(gpu, os_, bases, offs, mode)


def test_gpu_builder_small_and_odd_sets():
    # tiny cuttlefish unitigs, every w; duplicated k-mers; k = 32; unitigs of length exactly k
    tiny = mz.UnitigSet.from_seqs(["CACACACCAC", "CCTCAATACG"], 7)
    for w in range(1, 8):
        for skew in (NOSKEW, 0, 1):
            _assert_same_tables(mz.SSHash.from_unitig_set(tiny, w, skew), mz.SSHash.from_unitig_set(tiny, w, skew, builder="gpu"), "tiny w=%d" % w)
    for k, w, skew, n_u, extra, seed in [(32, 20, 8, 300, 50, 1), (21, 11, 4, 500, 30, 2), (15, 7, 3, 200, 40, 3), (31, 19, 64, 2000, 1, 4), (5, 3, 2, 10, 6, 5)]:
        codes, accum = _gen.synthetic_unitigs(n_u, extra, k, seed=seed)
        us = mz.UnitigSet(k, mz.pack_2bit(codes), len(codes), accum)
        host, gpu = mz.SSHash.from_unitig_set(us, w, skew), mz.SSHash.from_unitig_set(us, w, skew, builder="gpu")
        _assert_same_tables(host, gpu, "synthetic k=%d w=%d" % (k, w))
        assert gpu.k2u_validate_self() == host.k2u_validate_self()


def test_gpu_builder_larger_synthetic():
    k = 31
    codes, accum = _gen.synthetic_unitigs(200000, 68, k, seed=45)  # ~2e7 bases
    us = mz.UnitigSet(k, mz.pack_2bit(codes), len(codes), accum)
    host, gpu = mz.SSHash.from_unitig_set(us, 19, 64), mz.SSHash.from_unitig_set(us, 19, 64, builder="gpu")
    _assert_same_tables(host, gpu, "2e7-base synthetic set")
    c = gpu.k2u_validate_self()
    assert c[0] == 2 * gpu.n_kmers and c[4] == 0


def test_gpu_builder_pfhash_bit_identical(yeast_dense, yeast_queries):
    _, o = yeast_dense
    q, _ = yeast_queries
    us = mz.UnitigSet(o.k, o.useq_words(), o.total_len, o.unitig_starts())
    host, gpu = mz.PFHash.from_unitig_set(us), mz.PFHash.from_unitig_set(us, builder="gpu")
    for i in (0, 3, 6):
        assert host.table_digest(i) == gpu.table_digest(i), TABLES[i]
    assert gpu.k2u_validate_self() == host.k2u_validate_self() == o.k2u_validate_self()
    assert_hits_equal(gpu.k2u_batch(q[:100000]), o.k2u_batch(q[:100000]), "GPU-built PFHash vs oracle")
    codes, accum = _gen.synthetic_unitigs(100000, 40, 31, seed=9)
    us2 = mz.UnitigSet(31, mz.pack_2bit(codes), len(codes), accum)
    a, b = mz.PFHash.from_unitig_set(us2), mz.PFHash.from_unitig_set(us2, builder="gpu")
    for i in (0, 3, 6):
        assert a.table_digest(i) == b.table_digest(i), TABLES[i]
    c = b.k2u_validate_self()
    assert c[0] == 2 * b.n_kmers and c[4] == 0
