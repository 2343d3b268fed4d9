"""mazu_b200 -- thin Python mirror of mazu's index/query trait surface over libmazu_b200.so.

The product is the C-ABI library (include/mazu_b200.h, mazu_b200/csrc): C++ host code plus
hand-written sm_100a CUDA kernels.  This module only binds it with ctypes so the parity tests and
bench.py read like the reference's own tests:

    DenseIndex.deserialize_from_cpp(dir)          src/pf1/dense_index.rs:33-97
    PiscemIndex.from_cf_prefix(prefix, w, skew)   src/index/piscem_index.rs:17-20
    PufferfishDenseIndex.from_cf_prefix(prefix)   src/index/defaults.rs:18-21
    SSHash.from_unitig_set(unitigs, w, skew, bh)  src/kphf/sshash.rs:405-412
    PFHash.from_unitig_set(unitigs)               src/kphf/pfhash.rs:40-73
    ModIndex.from_parts(k2u, u2pos, refs)         src/index.rs:80-87
    index.as_streaming()                          src/index/caching.rs:227-231
    index.validate_self()                         src/index/validate.rs:24-52

There is NO CPU fallback: importing works anywhere (so the CPU test-suite can check the exported
symbols), but every query needs a CUDA device and raises MazuError otherwise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MAZU_B200_LIB") or os.path.join(_HERE, "libmazu_b200.so")  # override: A/B of two builds
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "mazu_b200.h")
DEBUG_HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "mazu_b200_debug.h")  # test / measurement hooks, not the boundary

HIT_DTYPE = np.dtype([("unitig_id", "<u4"), ("unitig_len", "<u4"), ("pos", "<u4"), ("match", "<u4")])
INTERVAL_DTYPE = np.dtype([("unitig_id", "<u4"), ("pos_o", "<u4"), ("read", "<u4"), ("start", "<u2"), ("len", "<u2")])  # mazu_hit_interval_t
TILE_DTYPE = np.dtype([("unitig_id", "<u4"), ("unitig_len", "<u4"), ("pos", "<u4"), ("fw", "<u4")])
HIT8_DTYPE = np.dtype([("unitig_id", "<u4"), ("pos_match", "<u4")])
OCC_DTYPE = np.dtype([("ref_id", "<u4"), ("pos", "<u4"), ("fw", "<u4")])

NO_MATCH, IDENTITY_MATCH, TWIN_MATCH, SKIPPED = 0, 1, 2, 3
MEM_HOST, MEM_DEVICE, MEM_HOST_IN_DEVICE_OUT = 0, 1, 2
MODE_RANDOM, MODE_STREAMING = 0, 1
K2U_PFHASH, K2U_SSHASH, K2U_SAMPLED_PFHASH = 0, 1, 2
U2POS_NONE, U2POS_DENSE, U2POS_PISCEM = 0, 1, 2
INDEX_PUFFERFISH_DENSE, INDEX_PISCEM = 0, 1
SKEW_NONE = 0xFFFFFFFFFFFFFFFF
MISS = 0xFFFFFFFF

ERR_K_MISMATCH = -6
ERR_NO_U2POS = -8
ERR_NO_REFSEQ = -9

(INFO_K, INFO_N_UNITIGS, INFO_N_KMERS, INFO_SUM_UNITIGS_LEN, INFO_N_MINIMIZERS, INFO_N_KMERS_IN_SKEW_INDEX, INFO_N_REFS,
 INFO_N_TOTAL_OCCS, INFO_K2U_KIND, INFO_U2POS_KIND, INFO_DEVICE_BYTES, INFO_W, INFO_N_MINIMIZER_OCCS, INFO_MPHF_LEVELS,
 INFO_DEVICE, INFO_SAMPLE_SIZE, INFO_EXTENSION_SIZE, INFO_KMERS_UNIQUE) = range(18)


class PinnedArray:
    """numpy view of page-locked host memory from mazu_b200_alloc_pinned (freed with the object)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * self.dtype.itemsize
        p = C.c_void_p(0)
        _check(lib().mazu_b200_alloc_pinned(n, C.byref(p)))
        self._p = p.value
        buf = (C.c_uint8 * max(n, 1)).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        try:
            if getattr(self, "_p", None):
                self.array = None
                lib().mazu_b200_free_pinned(self._p)
                self._p = None
        except Exception:
            pass


class MazuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("[%d] %s" % (code, msg))
        self.code = code


class UnitigSetDesc(C.Structure):
    _fields_ = [("k", C.c_uint32), ("useq_words", C.c_void_p), ("n_bases", C.c_uint64), ("accum_lens", C.c_void_p),
                ("n_unitigs", C.c_uint64)]


class PackedVecDesc(C.Structure):
    _fields_ = [("words", C.c_void_p), ("width", C.c_uint64), ("len", C.c_uint64)]


class BooPHFDesc(C.Structure):
    _fields_ = [("n_levels", C.c_uint32), ("level_words", C.POINTER(C.c_void_p)), ("level_n_bits", C.c_void_p),
                ("last_bitset_rank", C.c_uint64), ("n_elem", C.c_uint64), ("final_keys", C.c_void_p),
                ("final_vals", C.c_void_p), ("n_final", C.c_uint64)]


_lib = None

_SIGNATURES = None


def _signatures():
    u64, u32, i32, vp, cp = C.c_uint64, C.c_uint32, C.c_int32, C.c_void_p, C.c_char_p
    pp = C.POINTER(C.c_void_p)
    return {
        "mazu_b200_last_error": (cp, []),
        "mazu_b200_device_count": (i32, []),
        "mazu_b200_dense_index_deserialize_from_cpp": (i32, [cp, i32, pp]),
        "mazu_b200_sparse_index_deserialize_from_cpp": (i32, [cp, i32, pp]),
        "mazu_b200_index_from_cf_prefix": (i32, [cp, i32, u32, u64, u64, i32, pp]),
        "mazu_b200_index_create_sshash": (i32, [C.POINTER(UnitigSetDesc), u32, u64, u64, i32, pp]),
        "mazu_b200_index_create_sshash_gpu": (i32, [C.POINTER(UnitigSetDesc), u32, u64, u64, i32, pp]),
        "mazu_b200_index_create_pfhash_gpu": (i32, [C.POINTER(UnitigSetDesc), i32, pp]),
        "mazu_b200_debug_table_digest": (i32, [vp, i32, vp, vp]),
        "mazu_b200_debug_probe_key": (i32, [vp, vp, u64, vp, vp]),
        "mazu_b200_index_create_pfhash": (i32, [C.POINTER(UnitigSetDesc), i32, pp]),
        "mazu_b200_index_create_pfhash_from_parts": (i32, [C.POINTER(UnitigSetDesc), C.POINTER(BooPHFDesc), C.POINTER(PackedVecDesc), i32, pp]),
        "mazu_b200_index_rebuild_k2u": (i32, [vp, i32, u32, u64, u64, pp]),
        "mazu_b200_index_attach_u2pos_dense": (i32, [vp, vp, u64, C.POINTER(PackedVecDesc)]),
        "mazu_b200_index_attach_u2pos_piscem": (i32, [vp, C.POINTER(PackedVecDesc), u64, u64, C.POINTER(PackedVecDesc)]),
        "mazu_b200_index_attach_refseq": (i32, [vp, vp, vp, u64]),
        "mazu_b200_index_destroy": (None, [vp]),
        "mazu_b200_index_release_scratch": (i32, [vp, vp]),
        "mazu_b200_alloc_pinned": (i32, [u64, vp]),
        "mazu_b200_free_pinned": (None, [vp]),
        "mazu_b200_index_info": (u64, [vp, i32]),
        "mazu_b200_unitig_len": (i32, [vp, u64, vp, vp]),
        "mazu_b200_k2u_batch": (i32, [vp, vp, u64, u32, vp, i32, vp]),
        "mazu_b200_count_kmer_slots": (u64, [vp, vp, u64, u64]),
        "mazu_b200_query_reads": (i32, [vp, vp, vp, u64, u64, i32, vp, vp, vp, i32, vp]),
        "mazu_b200_query_reads_compact": (i32, [vp, vp, vp, u64, u64, i32, vp, vp, vp, i32, vp]),
        "mazu_b200_query_reads_runs": (i32, [vp, vp, vp, u64, u64, i32, vp, vp, vp, u64, vp, vp, vp]),
        "mazu_b200_expand_hit_runs": (i32, [vp, vp, vp, vp, u64, u64, vp]),
        "mazu_b200_query_reads_runs_packed": (i32, [vp, vp, vp, u64, u64, i32, vp, vp, u64, vp, vp, vp]),
        "mazu_b200_pack_reads": (i32, [vp, u64, u64, vp, vp, vp]),
        "mazu_b200_query_reads_intervals_packed": (i32, [vp, vp, vp, u64, u64, i32, vp, u64, vp, vp]),
        "mazu_b200_expand_hit_intervals": (i32, [vp, vp, u64, vp, u64, u64, vp]),
        "mazu_b200_query_reads_intervals": (i32, [vp, vp, u64, u64, i32, vp, u64, vp, vp]),
        "mazu_b200_expand_hit_intervals_ascii": (i32, [vp, vp, u64, vp, u64, u64, vp]),
        "mazu_b200_expand_hit_runs_packed": (i32, [vp, vp, vp, u64, u64, vp]),
        "mazu_b200_encode_reads": (i32, [vp, vp, vp, u64, u64, vp, vp, vp, vp, vp, vp, vp]),
        "mazu_b200_decode_occs": (i32, [vp, vp, u64, vp, vp, u64, vp, i32, vp]),
        "mazu_b200_project_hits": (i32, [vp, vp, u64, vp, vp, u64, vp, i32, vp]),
        "mazu_b200_get_ref_pos_reads": (i32, [vp, vp, vp, u64, u64, i32, u64, vp, vp, vp, vp, u64, vp, vp, i32, vp]),
        "mazu_b200_iter_unitigs_on_ref": (i32, [vp, u64, vp, u64, vp]),
        "mazu_b200_validate_self": (i32, [vp, vp]),
        "mazu_b200_k2u_validate_self": (i32, [vp, vp]),
        "mazu_b200_unitig_seq": (i32, [vp, u64, vp, u64, vp]),
        "mazu_b200_fasta_open": (i32, [cp, pp]),
        "mazu_b200_fasta_close": (None, [vp]),
        "mazu_b200_fasta_n_records": (u64, [vp]),
        "mazu_b200_fasta_bases": (vp, [vp]),
        "mazu_b200_fasta_offsets": (vp, [vp]),
        "mazu_b200_fasta_name": (cp, [vp, u64]),
        "mazu_b200_validate_fasta": (i32, [vp, cp, i32, vp]),
        "mazu_b200_validate_reads": (i32, [vp, vp, vp, u64, i32, vp]),
        "mazu_b200_index_replicate": (i32, [vp, vp, i32, pp]),
        "mazu_b200_query_reads_sharded": (i32, [vp, i32, vp, vp, u64, u64, i32, vp, vp, vp]),
        "mazu_b200_query_reads_runs_sharded": (i32, [vp, i32, vp, vp, u64, u64, i32, vp, vp, vp, u64, vp, vp, vp]),
        "mazu_b200_debug_gather_probe": (i32, [u64, u64, i32, i32, i32, i32, i32, C.POINTER(C.c_double)]),
    }


def exported_symbols():
    """Names every entry point include/mazu_b200.h declares (used by the CPU symbol test)."""
    return sorted(_signatures().keys())


def lib():
    """Load libmazu_b200.so.  Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MazuError(-5, "libmazu_b200.so is missing (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                                "mazu_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _signatures().items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise MazuError(rc, lib().mazu_b200_last_error().decode(errors="replace"))


def _np_ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _any_ptr(x):
    """numpy array -> host pointer; torch tensor -> data_ptr(); int/None passed through."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data_as(C.c_void_p)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(int(x))


def device_count():
    return int(lib().mazu_b200_device_count())


def encode_kmer(s):
    """kmers::Kmer::from(&str): base i at bits [2i, 2i+2), A=0 C=1 G=2 T=3."""
    w = 0
    for i, ch in enumerate(s.upper()):
        w |= "ACGT".index(ch) << (2 * i)
    return w


class UnitigSet:
    """UnitigSet (src/unitig_set.rs:31-36) as host arrays: 2-bit packed sequence + prefix lengths."""

    def __init__(self, k, useq_words, n_bases, accum_lens):
        self.k = int(k)
        self.useq_words = np.ascontiguousarray(useq_words, dtype=np.uint64)
        self.n_bases = int(n_bases)
        self.accum_lens = np.ascontiguousarray(accum_lens, dtype=np.uint64)

    @classmethod
    def from_seqs(cls, seqs, k):  # unitig_set.rs:74-106
        lens = np.array([len(s) for s in seqs], dtype=np.uint64)
        accum = np.zeros(len(seqs) + 1, dtype=np.uint64)
        accum[1:] = np.cumsum(lens)
        codes = np.frombuffer("".join(seqs).upper().encode(), dtype=np.uint8)
        lut = np.full(256, 255, dtype=np.uint8)
        for i, ch in enumerate(b"ACGT"):
            lut[ch] = i
        c = lut[codes]
        if (c == 255).any():
            raise MazuError(-2, "non-ACGT base in unitig sequence")
        return cls(k, pack_2bit(c), len(c), accum)

    def desc(self):
        return UnitigSetDesc(self.k, _np_ptr(self.useq_words), self.n_bases, _np_ptr(self.accum_lens), len(self.accum_lens) - 1)


def pack_2bit(codes):
    """codes: uint8 array of values 0..3 -> uint64 words, base i at bits [2i, 2i+2)."""
    n = len(codes)
    nw = (2 * n + 63) // 64
    padded = np.zeros(nw * 32, dtype=np.uint64)
    padded[:n] = codes
    shifts = (np.arange(32, dtype=np.uint64) * np.uint64(2))
    return np.bitwise_or.reduce(padded.reshape(nw, 32) << shifts, axis=1).astype(np.uint64) if nw else np.zeros(0, dtype=np.uint64)


class PackedVec:
    """simple-sds IntVector / pufferfish compact vector view for descriptors."""

    def __init__(self, words, width, length):
        self.words = np.ascontiguousarray(words, dtype=np.uint64)
        self.width = int(width)
        self.len = int(length)

    @classmethod
    def pack(cls, values, width=None):
        values = np.asarray(values, dtype=np.uint64)
        if width is None:
            width = max(1, int(values.max()).bit_length()) if len(values) else 1
        n = len(values)
        words = np.zeros((n * width + 63) // 64 + 1, dtype=np.uint64)
        if n:
            bit = np.arange(n, dtype=np.uint64) * np.uint64(width)
            wi = (bit >> np.uint64(6)).astype(np.int64)
            sh = bit & np.uint64(63)

            def scatter_or(idx, vals):  # idx is non-decreasing: OR-reduce each run, one store per word
                first = np.flatnonzero(np.concatenate(([True], idx[1:] != idx[:-1])))
                words[idx[first]] |= np.bitwise_or.reduceat(vals, first)

            scatter_or(wi, values << sh)
            spill = (sh + np.uint64(width)) > np.uint64(64)
            if spill.any():
                scatter_or(wi[spill] + 1, values[spill] >> (np.uint64(64) - sh[spill]))
        return cls(words, width, n)

    def desc(self):
        return PackedVecDesc(_np_ptr(self.words), self.width, self.len)


class ModIndex:
    """ModIndex<H, T> (src/index.rs:50-55): handle on a device-resident index."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)
        self._keep = []

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().mazu_b200_index_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def release_scratch(self):
        """Give the handle's idle staging memory back to the driver; returns the bytes released."""
        n = C.c_uint64(0)
        _check(lib().mazu_b200_index_release_scratch(self._h, C.byref(n)))
        return int(n.value)

    # --- construction ----------------------------------------------------------------------
    @staticmethod
    def _out():
        return C.c_void_p(0)

    @classmethod
    def deserialize_from_cpp(cls, d, device=0):
        out = cls._out()
        _check(lib().mazu_b200_dense_index_deserialize_from_cpp(os.fspath(d).encode(), device, C.byref(out)))
        return cls(out.value)

    @classmethod
    def from_cf_prefix(cls, prefix, index_kind, w=0, skew_param=SKEW_NONE, seed=0, device=0):
        out = cls._out()
        _check(lib().mazu_b200_index_from_cf_prefix(os.fspath(prefix).encode(), index_kind, w, skew_param, seed, device, C.byref(out)))
        return cls(out.value)

    @classmethod
    def sshash_from_unitig_set(cls, unitigs, w, skew_param=SKEW_NONE, seed=0, device=0):
        out = cls._out()
        d = unitigs.desc()
        _check(lib().mazu_b200_index_create_sshash(C.byref(d), w, skew_param, seed, device, C.byref(out)))
        return cls(out.value)

    @classmethod
    def sshash_from_unitig_set_gpu(cls, unitigs, w, skew_param=SKEW_NONE, seed=0, device=0):
        """SSHash::from_unitig_set with every build stage on the device (tables bit-identical to the host builder)."""
        out = cls._out()
        d = unitigs.desc()
        _check(lib().mazu_b200_index_create_sshash_gpu(C.byref(d), w, skew_param, seed, device, C.byref(out)))
        return cls(out.value)

    def table_digest(self, which):
        dg, nb = C.c_uint64(0), C.c_uint64(0)
        _check(lib().mazu_b200_debug_table_digest(self._h, which, C.byref(dg), C.byref(nb)))
        return dg.value, nb.value

    @classmethod
    def pfhash_from_unitig_set(cls, unitigs, device=0, builder="host"):
        out = cls._out()
        d = unitigs.desc()
        fn = lib().mazu_b200_index_create_pfhash_gpu if builder == "gpu" else lib().mazu_b200_index_create_pfhash
        _check(fn(C.byref(d), device, C.byref(out)))
        return cls(out.value)

    def rebuild_k2u(self, k2u_kind, w=0, skew_param=SKEW_NONE, seed=0):
        """ModIndex::from_parts(base, <new K2U over the same unitigs>, u2pos.clone(), refs.clone())."""
        out = self._out()
        _check(lib().mazu_b200_index_rebuild_k2u(self._h, k2u_kind, w, skew_param, seed, C.byref(out)))
        return ModIndex(out.value)

    def attach_u2pos_dense(self, ctable_words, contig_offsets):
        ctable_words = np.ascontiguousarray(ctable_words, dtype=np.uint64)
        d = contig_offsets.desc()
        _check(lib().mazu_b200_index_attach_u2pos_dense(self._h, _np_ptr(ctable_words), len(ctable_words), C.byref(d)))

    def attach_u2pos_piscem(self, ctable, ref_shift, pos_mask, contig_offsets):
        a, b = ctable.desc(), contig_offsets.desc()
        _check(lib().mazu_b200_index_attach_u2pos_piscem(self._h, C.byref(a), ref_shift, pos_mask, C.byref(b)))

    def attach_refseq(self, seq_words, prefix_sum):
        prefix_sum = np.ascontiguousarray(prefix_sum, dtype=np.uint64)
        sw = None if seq_words is None else np.ascontiguousarray(seq_words, dtype=np.uint64)
        _check(lib().mazu_b200_index_attach_refseq(self._h, _np_ptr(sw), _np_ptr(prefix_sum), len(prefix_sum) - 1))

    # --- K2U accessors (src/kphf/mod.rs:58-67) ------------------------------------------------
    def info(self, what):
        return int(lib().mazu_b200_index_info(self._h, what))

    k = property(lambda s: s.info(INFO_K))
    n_unitigs = property(lambda s: s.info(INFO_N_UNITIGS))
    n_kmers = property(lambda s: s.info(INFO_N_KMERS))
    sum_unitigs_len = property(lambda s: s.info(INFO_SUM_UNITIGS_LEN))
    n_minimizers = property(lambda s: s.info(INFO_N_MINIMIZERS))
    n_kmers_in_skew_index = property(lambda s: s.info(INFO_N_KMERS_IN_SKEW_INDEX))
    n_refs = property(lambda s: s.info(INFO_N_REFS))
    n_total_occs = property(lambda s: s.info(INFO_N_TOTAL_OCCS))
    device_bytes = property(lambda s: s.info(INFO_DEVICE_BYTES))
    device = property(lambda s: s.info(INFO_DEVICE))
    kmers_unique = property(lambda s: bool(s.info(INFO_KMERS_UNIQUE)))

    def unitig_len(self, ui):
        ln, st = C.c_uint64(0), C.c_uint64(0)
        _check(lib().mazu_b200_unitig_len(self._h, ui, C.byref(ln), C.byref(st)))
        return ln.value

    def unitig_start_pos(self, ui):
        ln, st = C.c_uint64(0), C.c_uint64(0)
        _check(lib().mazu_b200_unitig_len(self._h, ui, C.byref(ln), C.byref(st)))
        return st.value

    # --- queries ----------------------------------------------------------------------------
    def k2u_batch(self, fw_words, k=None, out=None, mem=MEM_HOST, stream=None, n=None):
        """K2U::k2u for a batch of forward k-mer words.  Host mode: numpy in/out.  Device mode: pass
        torch tensors (or raw pointers) for fw_words/out plus `n` and the CUDA stream handle."""
        k = self.k if k is None else k
        if mem == MEM_HOST:
            fw_words = np.ascontiguousarray(fw_words, dtype=np.uint64)
            n = len(fw_words)
            if out is None:
                out = np.empty(n, dtype=HIT_DTYPE)
        elif n is None:
            n = int(fw_words.numel())
        _check(lib().mazu_b200_k2u_batch(self._h, _any_ptr(fw_words), n, k, _any_ptr(out), mem, _any_ptr(stream)))
        return out

    def count_kmer_slots(self, read_offsets=None, n_reads=0, uniform_read_len=0):
        ro = None if read_offsets is None else np.ascontiguousarray(read_offsets, dtype=np.uint64)
        if ro is not None:
            n_reads = len(ro) - 1
        return int(lib().mazu_b200_count_kmer_slots(self._h, _np_ptr(ro), n_reads, uniform_read_len))

    def query_reads(self, bases, read_offsets=None, n_reads=None, uniform_read_len=0, mode=MODE_RANDOM, want_hits=True,
                    out_hits=None, kmer_offsets=None, counts=None, mem=MEM_HOST, stream=None, compact=False):
        """The read loop of `kphf bench` / validate_ckmers.  Host mode returns (hits, counts, kmer_offsets).
        compact=True writes 8-byte mazu_hit8_t records (HIT8_DTYPE) instead of 16-byte mazu_hit_t."""
        fn = lib().mazu_b200_query_reads_compact if compact else lib().mazu_b200_query_reads
        if mem in (MEM_HOST, MEM_HOST_IN_DEVICE_OUT):
            bases = np.ascontiguousarray(bases, dtype=np.uint8)
            if uniform_read_len:
                n_reads = len(bases) // uniform_read_len if n_reads is None else n_reads
                ro = None
            else:
                ro = np.ascontiguousarray(read_offsets, dtype=np.uint64)
                n_reads = len(ro) - 1
            if kmer_offsets is None:  # implied by arithmetic for uniform reads; materialised only for ragged batches
                koffs = None if uniform_read_len else np.zeros(n_reads + 1, dtype=np.uint64)
            else:
                koffs = kmer_offsets
            if mem == MEM_HOST_IN_DEVICE_OUT and want_hits and out_hits is None:
                raise ValueError("MEM_HOST_IN_DEVICE_OUT needs a device buffer in out_hits")
            if want_hits and out_hits is None:
                out_hits = np.empty(self.count_kmer_slots(ro, n_reads, uniform_read_len), dtype=HIT8_DTYPE if compact else HIT_DTYPE)
            cnt = np.zeros(3, dtype=np.uint64) if counts is None else counts
            _check(fn(self._h, _np_ptr(bases), _np_ptr(ro), n_reads, uniform_read_len, mode, _np_ptr(koffs),
                                               _any_ptr(out_hits) if want_hits else None, _np_ptr(cnt), mem, None))
            return out_hits, cnt, koffs
        _check(fn(self._h, _any_ptr(bases), _any_ptr(read_offsets), n_reads, uniform_read_len, mode,
                  _any_ptr(kmer_offsets), _any_ptr(out_hits), _any_ptr(counts), MEM_DEVICE, _any_ptr(stream)))
        return out_hits, counts, kmer_offsets

    def query_reads_runs(self, bases, read_offsets=None, uniform_read_len=0, mode=MODE_RANDOM, codes=None, runs=None, read_run_offsets=None):
        """query_reads with the hit records as runs (host buffers): returns (codes, runs[:n_runs], read_run_offsets, counts, kmer_offsets)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        if uniform_read_len:
            n_reads, ro = len(bases) // uniform_read_len, None
            koffs = None
        else:
            ro = np.ascontiguousarray(read_offsets, dtype=np.uint64)
            n_reads = len(ro) - 1
            koffs = np.zeros(n_reads + 1, dtype=np.uint64)
        n_slots = self.count_kmer_slots(ro, n_reads, uniform_read_len)
        codes = np.empty(n_slots, dtype=np.uint8) if codes is None else codes
        rro = np.zeros(n_reads + 1, dtype=np.uint64) if read_run_offsets is None else read_run_offsets
        cnt = np.zeros(3, dtype=np.uint64)
        n_runs = C.c_uint64(0)
        if runs is None:
            runs = np.empty(max(1024, n_slots // 16), dtype=HIT_DTYPE)
        while True:
            rc = lib().mazu_b200_query_reads_runs(self._h, _np_ptr(bases), _np_ptr(ro), n_reads, uniform_read_len, mode, _np_ptr(koffs), _any_ptr(codes),
                                                  _any_ptr(runs), len(runs), _any_ptr(rro), C.byref(n_runs), _np_ptr(cnt))
            if rc != 0 and n_runs.value > len(runs):  # capacity: the call reports what it needs
                runs = np.empty(n_runs.value, dtype=HIT_DTYPE)
                continue
            _check(rc)
            break
        return codes, runs[: n_runs.value], rro, cnt, koffs

    def query_reads_runs_packed(self, words, n_mask, n_reads, read_len, mode=MODE_RANDOM, codes2=None, runs=None, read_run_offsets=None):
        """mazu_b200_query_reads_runs_packed: 2-bit reads in, 2-bit codes out: (codes2, runs[:n_runs], read_run_offsets, counts)."""
        n_slots = n_reads * max(read_len - self.k + 1, 0)
        codes2 = np.zeros((n_slots + 3) // 4, dtype=np.uint8) if codes2 is None else codes2
        rro = np.zeros(n_reads + 1, dtype=np.uint64) if read_run_offsets is None else read_run_offsets
        cnt = np.zeros(3, dtype=np.uint64)
        n_runs = C.c_uint64(0)
        if runs is None:
            runs = np.empty(max(1024, n_slots // 16), dtype=HIT_DTYPE)
        while True:
            rc = lib().mazu_b200_query_reads_runs_packed(self._h, _any_ptr(words), _any_ptr(n_mask), n_reads, read_len, mode, _any_ptr(codes2),
                                                         _any_ptr(runs), len(runs), _any_ptr(rro), C.byref(n_runs), _np_ptr(cnt))
            if rc != 0 and n_runs.value > len(runs):
                runs = np.empty(n_runs.value, dtype=HIT_DTYPE)
                continue
            _check(rc)
            break
        return codes2, runs[: n_runs.value], rro, cnt

    def query_reads_intervals_packed(self, words, n_mask, n_reads, read_len, mode=MODE_RANDOM, intervals=None):
        """mazu_b200_query_reads_intervals_packed: 2-bit reads in, one 16-byte record per hit run out: (intervals[:n], counts)."""
        cnt = np.zeros(3, dtype=np.uint64)
        n = C.c_uint64(0)
        if intervals is None:
            intervals = np.empty(max(1024, n_reads * max(read_len - self.k + 1, 0) // 16), dtype=INTERVAL_DTYPE)
        while True:
            rc = lib().mazu_b200_query_reads_intervals_packed(self._h, _any_ptr(words), _any_ptr(n_mask), n_reads, read_len, mode, _any_ptr(intervals),
                                                              len(intervals), C.byref(n), _np_ptr(cnt))
            if rc != 0 and n.value > len(intervals):
                intervals = np.empty(n.value, dtype=INTERVAL_DTYPE)
                continue
            _check(rc)
            break
        return intervals[: n.value], cnt

    def query_reads_intervals(self, bases, n_reads, read_len, mode=MODE_RANDOM, intervals=None):
        """mazu_b200_query_reads_intervals: ASCII reads (n_reads x read_len bytes) in, one 16-byte record per hit run out."""
        cnt = np.zeros(3, dtype=np.uint64)
        n = C.c_uint64(0)
        if intervals is None:
            intervals = np.empty(max(1024, n_reads * max(read_len - self.k + 1, 0) // 16), dtype=INTERVAL_DTYPE)
        while True:
            rc = lib().mazu_b200_query_reads_intervals(self._h, _any_ptr(bases), n_reads, read_len, mode, _any_ptr(intervals), len(intervals),
                                                       C.byref(n), _np_ptr(cnt))
            if rc != 0 and n.value > len(intervals):
                intervals = np.empty(n.value, dtype=INTERVAL_DTYPE)
                continue
            _check(rc)
            break
        return intervals[: n.value], cnt

    def expand_hit_intervals_ascii(self, intervals, bases, n_reads, read_len, out=None):
        """mazu_b200_expand_hit_intervals_ascii: every mazu_hit_t of the batch from the interval records + the ASCII reads."""
        out = np.empty(n_reads * max(read_len - self.k + 1, 0), dtype=HIT_DTYPE) if out is None else out
        _check(lib().mazu_b200_expand_hit_intervals_ascii(self._h, _any_ptr(intervals), len(intervals), _any_ptr(bases), n_reads, read_len,
                                                          _any_ptr(out)))
        return out

    def expand_hit_intervals(self, intervals, n_mask, n_reads, read_len, out=None):
        """mazu_b200_expand_hit_intervals: every mazu_hit_t of the batch from the interval records (+ the N mask)."""
        out = np.empty(n_reads * max(read_len - self.k + 1, 0), dtype=HIT_DTYPE) if out is None else out
        _check(lib().mazu_b200_expand_hit_intervals(self._h, _any_ptr(intervals), len(intervals), _any_ptr(n_mask), n_reads, read_len, _any_ptr(out)))
        return out

    @staticmethod
    def expand_hit_runs_packed(codes2, runs, read_run_offsets, uniform_slots, out=None):
        n_reads = len(read_run_offsets) - 1
        out = np.empty(n_reads * uniform_slots, dtype=HIT_DTYPE) if out is None else out
        _check(lib().mazu_b200_expand_hit_runs_packed(_any_ptr(codes2), _any_ptr(runs), _any_ptr(read_run_offsets), n_reads, uniform_slots, _any_ptr(out)))
        return out

    @staticmethod
    def expand_hit_runs(codes, runs, read_run_offsets, kmer_offsets=None, uniform_slots=0, out=None):
        n_reads = len(read_run_offsets) - 1
        out = np.empty(len(codes), dtype=HIT_DTYPE) if out is None else out
        _check(lib().mazu_b200_expand_hit_runs(_any_ptr(codes), _any_ptr(runs), _any_ptr(read_run_offsets), _np_ptr(kmer_offsets), n_reads, uniform_slots,
                                               _any_ptr(out)))
        return out

    def encode_reads(self, bases, read_offsets, n_reads, uniform_read_len, kmer_offsets, out_fw, out_rc, out_mm, out_off, out_valid,
                     stream=None):
        _check(lib().mazu_b200_encode_reads(self._h, _any_ptr(bases), _any_ptr(read_offsets), n_reads, uniform_read_len,
                                            _any_ptr(kmer_offsets), _any_ptr(out_fw), _any_ptr(out_rc), _any_ptr(out_mm), _any_ptr(out_off),
                                            _any_ptr(out_valid), _any_ptr(stream)))

    def _occ_call(self, fn, inp, n):
        offs = np.zeros(n + 1, dtype=np.uint64)
        total = C.c_uint64(0)
        _check(fn(self._h, _np_ptr(inp), n, _np_ptr(offs), None, 0, C.byref(total), MEM_HOST, None))
        out = np.empty(total.value, dtype=OCC_DTYPE)
        if total.value:
            _check(fn(self._h, _np_ptr(inp), n, _np_ptr(offs), _np_ptr(out), total.value, C.byref(total), MEM_HOST, None))
        return offs, out

    def decode_occs(self, unitig_ids):
        """U2Pos::encoded_unitig_occs + decode_unitig_occs for a batch of unitig ids (host mode)."""
        unitig_ids = np.ascontiguousarray(unitig_ids, dtype=np.uint32)
        return self._occ_call(lib().mazu_b200_decode_occs, unitig_ids, len(unitig_ids))

    def project_hits(self, hits):
        """GetRefPos::project_hits for a batch of hit records (host mode)."""
        hits = np.ascontiguousarray(hits, dtype=HIT_DTYPE)
        return self._occ_call(lib().mazu_b200_project_hits, hits, len(hits))

    def get_ref_pos_reads(self, bases, read_offsets=None, uniform_read_len=0, mode=MODE_RANDOM, want_hits=True):
        """GetRefPos::get_ref_pos + project_hits over reads in one pass (host mode): (hits, offsets, mrps, counts, kmer_offsets)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        if uniform_read_len:
            n_reads, ro = len(bases) // uniform_read_len, None
        else:
            ro = np.ascontiguousarray(read_offsets, dtype=np.uint64)
            n_reads = len(ro) - 1
        n_slots = self.count_kmer_slots(ro, n_reads, uniform_read_len)
        koffs = np.zeros(n_reads + 1, dtype=np.uint64)
        hits = np.empty(n_slots, dtype=HIT_DTYPE) if want_hits else None
        offs = np.zeros(n_slots + 1, dtype=np.uint64)
        cnt = np.zeros(3, dtype=np.uint64)
        total = C.c_uint64(0)
        cap = max(1024, 2 * n_slots)
        while True:
            out = np.empty(cap, dtype=OCC_DTYPE)
            rc = lib().mazu_b200_get_ref_pos_reads(self._h, _np_ptr(bases), _np_ptr(ro), n_reads, uniform_read_len, mode, n_slots, _np_ptr(koffs),
                                                   _np_ptr(hits), _np_ptr(offs), _np_ptr(out), cap, C.byref(total), _np_ptr(cnt), MEM_HOST, None)
            if rc != 0 and total.value > cap:
                cap = total.value
                continue
            _check(rc)
            break
        return hits, offs, out[: total.value], cnt, koffs

    def get_ref_pos_eager(self, kmer):
        """GetRefPos::get_ref_pos_eager for one k-mer string: None or list of (ref_id, pos, fw)."""
        hit = self.k2u_batch(np.array([encode_kmer(kmer)], dtype=np.uint64), k=len(kmer))
        if hit[0]["match"] == NO_MATCH:
            return None
        _, mrps = self.project_hits(hit)
        return [(int(m["ref_id"]), int(m["pos"]), int(m["fw"])) for m in mrps]

    def iter_unitigs_on_ref(self, ref_id):
        """ModIndex::iter_unitigs_on_ref (src/index.rs:363-424): tiles as records {unitig_id, unitig_len, pos, fw}."""
        n = C.c_uint64(0)
        _check(lib().mazu_b200_iter_unitigs_on_ref(self._h, ref_id, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=TILE_DTYPE)
        if n.value:
            _check(lib().mazu_b200_iter_unitigs_on_ref(self._h, ref_id, _np_ptr(out), n.value, C.byref(n)))
        return out

    def validate_self(self):
        c = np.zeros(5, dtype=np.uint64)
        _check(lib().mazu_b200_validate_self(self._h, _np_ptr(c)))
        return [int(x) for x in c]

    def k2u_validate_self(self):
        c = np.zeros(5, dtype=np.uint64)
        _check(lib().mazu_b200_k2u_validate_self(self._h, _np_ptr(c)))
        return [int(x) for x in c]

    def validate_fasta(self, path, mode=MODE_RANDOM):
        """Validate::validate_fasta / StreamingIndex::validate_fasta: the library reads the file (FASTA or FASTQ)."""
        c = np.zeros(5, dtype=np.uint64)
        _check(lib().mazu_b200_validate_fasta(self._h, os.fspath(path).encode(), mode, _np_ptr(c)))
        return [int(x) for x in c]

    def validate_reads(self, bases, read_offsets, mode=MODE_RANDOM):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        ro = np.ascontiguousarray(read_offsets, dtype=np.uint64)
        c = np.zeros(5, dtype=np.uint64)
        _check(lib().mazu_b200_validate_reads(self._h, _np_ptr(bases), _np_ptr(ro), len(ro) - 1, mode, _np_ptr(c)))
        return [int(x) for x in c]

    def unitig_seq(self, ui):
        """K2U::unitig_seq(id) as a string."""
        n = C.c_uint64(0)
        _check(lib().mazu_b200_unitig_seq(self._h, ui, None, 0, C.byref(n)))
        w = np.zeros((n.value + 31) // 32, dtype=np.uint64)
        _check(lib().mazu_b200_unitig_seq(self._h, ui, _np_ptr(w), len(w), C.byref(n)))
        codes = ((w[:, None] >> (np.arange(32, dtype=np.uint64) * np.uint64(2))[None, :]) & np.uint64(3)).astype(np.uint8).reshape(-1)[: n.value]
        return np.frombuffer(b"ACGT", dtype=np.uint8)[codes].tobytes().decode()

    def replicate(self, devices):
        """mazu_b200_index_replicate: device-to-device copies of every table onto each device of `devices`."""
        dv = np.ascontiguousarray(devices, dtype=np.int32)
        out = (C.c_void_p * len(dv))()
        _check(lib().mazu_b200_index_replicate(self._h, _np_ptr(dv), len(dv), out))
        return [ModIndex(h) for h in out]

    def as_streaming(self):
        return StreamingIndex(self)


def _handle_array(indexes):
    arr = (C.c_void_p * len(indexes))()
    for i, ix in enumerate(indexes):
        arr[i] = ix._h.value
    return arr


def query_reads_sharded(indexes, bases, read_offsets=None, uniform_read_len=0, mode=MODE_RANDOM, out_hits=None):
    """mazu_b200_query_reads_sharded over replicas (host buffers): returns (hits, counts, kmer_offsets)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    if uniform_read_len:
        n_reads, ro = len(bases) // uniform_read_len, None
    else:
        ro = np.ascontiguousarray(read_offsets, dtype=np.uint64)
        n_reads = len(ro) - 1
    koffs = np.zeros(n_reads + 1, dtype=np.uint64)
    if out_hits is None:
        out_hits = np.empty(indexes[0].count_kmer_slots(ro, n_reads, uniform_read_len), dtype=HIT_DTYPE)
    cnt = np.zeros(3, dtype=np.uint64)
    _check(lib().mazu_b200_query_reads_sharded(_handle_array(indexes), len(indexes), _np_ptr(bases), _np_ptr(ro), n_reads, uniform_read_len, mode,
                                               _np_ptr(koffs), _any_ptr(out_hits), _np_ptr(cnt)))
    return out_hits, cnt, koffs


def query_reads_runs_sharded(indexes, bases, read_offsets=None, uniform_read_len=0, mode=MODE_RANDOM, codes=None, runs=None, read_run_offsets=None):
    """mazu_b200_query_reads_runs_sharded: returns (codes, runs (whole array), read_run_offsets, counts, kmer_offsets, n_runs)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    if uniform_read_len:
        n_reads, ro = len(bases) // uniform_read_len, None
    else:
        ro = np.ascontiguousarray(read_offsets, dtype=np.uint64)
        n_reads = len(ro) - 1
    koffs = np.zeros(n_reads + 1, dtype=np.uint64)
    n_slots = indexes[0].count_kmer_slots(ro, n_reads, uniform_read_len)
    codes = np.empty(n_slots, dtype=np.uint8) if codes is None else codes
    rro = np.zeros(n_reads + 1, dtype=np.uint64) if read_run_offsets is None else read_run_offsets
    if runs is None:
        runs = np.empty(max(1024 * len(indexes), n_slots // 8), dtype=HIT_DTYPE)
    cnt = np.zeros(3, dtype=np.uint64)
    n_runs = C.c_uint64(0)
    while True:
        rc = lib().mazu_b200_query_reads_runs_sharded(_handle_array(indexes), len(indexes), _np_ptr(bases), _np_ptr(ro), n_reads, uniform_read_len, mode,
                                                      _np_ptr(koffs), _any_ptr(codes), _any_ptr(runs), len(runs), _any_ptr(rro), C.byref(n_runs), _np_ptr(cnt))
        if rc != 0 and n_runs.value > len(runs):
            runs = np.empty(n_runs.value, dtype=HIT_DTYPE)
            continue
        _check(rc)
        break
    return codes, runs, rro, cnt, koffs, n_runs.value


class Fasta:
    """FastaReader (src/util.rs:93-149) through the library: records of a FASTA / FASTQ file as (bases, offsets, names)."""

    def __init__(self, path):
        h = C.c_void_p(0)
        _check(lib().mazu_b200_fasta_open(os.fspath(path).encode(), C.byref(h)))
        try:
            n = int(lib().mazu_b200_fasta_n_records(h))
            self.offsets = np.ctypeslib.as_array(C.cast(lib().mazu_b200_fasta_offsets(h), C.POINTER(C.c_uint64)), shape=(n + 1,)).copy()
            total = int(self.offsets[-1])
            self.bases = (np.ctypeslib.as_array(C.cast(lib().mazu_b200_fasta_bases(h), C.POINTER(C.c_uint8)), shape=(total,)).copy()
                          if total else np.zeros(0, dtype=np.uint8))
            self.names = [lib().mazu_b200_fasta_name(h, i).decode() for i in range(n)]
        finally:
            lib().mazu_b200_fasta_close(h)

    def seqs(self):
        return [self.bases[int(self.offsets[i]):int(self.offsets[i + 1])].tobytes().decode() for i in range(len(self.names))]


class StreamingIndex:
    """StreamingIndex / StreamingK2U (src/index/caching.rs:13-219): same queries through the
    contig-walk cache; on the device the cursor is per read (one warp per read)."""

    def __init__(self, index):
        self.index = index

    k = property(lambda s: s.index.k)

    def query_reads(self, bases, read_offsets=None, **kw):
        kw["mode"] = MODE_STREAMING
        return self.index.query_reads(bases, read_offsets, **kw)

    def validate_fasta(self, path):
        return self.index.validate_fasta(path, MODE_STREAMING)


class DenseIndex(ModIndex):
    """pf1::DenseIndex = ModIndex<PFHash<BooPHF<u64>>, DenseUnitigTable> (src/pf1/dense_index.rs:25)."""


class SparseIndex:
    """pf1::SparseIndex = ModIndex<SampledPFHash<BooPHF<u64>>, DenseUnitigTable> (src/pf1/sparse_index.rs:24)."""

    @staticmethod
    def deserialize_from_cpp(d, device=0):
        out = C.c_void_p(0)
        _check(lib().mazu_b200_sparse_index_deserialize_from_cpp(os.fspath(d).encode(), device, C.byref(out)))
        return ModIndex(out.value)


class PiscemIndex:
    @staticmethod
    def from_cf_prefix(prefix, w, skew_param, seed=0, device=0):
        return ModIndex.from_cf_prefix(prefix, INDEX_PISCEM, w, skew_param, seed, device)


class PufferfishDenseIndex:
    @staticmethod
    def from_cf_prefix(prefix, device=0):
        return ModIndex.from_cf_prefix(prefix, INDEX_PUFFERFISH_DENSE, device=device)


class SSHash:
    @staticmethod
    def from_unitig_set(unitigs, w, skew_param, seed=0, device=0, builder="host"):
        if builder == "gpu":
            return ModIndex.sshash_from_unitig_set_gpu(unitigs, w, skew_param, seed, device)
        return ModIndex.sshash_from_unitig_set(unitigs, w, skew_param, seed, device)

    @staticmethod
    def from_unitig_set_no_skew_index(unitigs, w, seed=0, device=0):
        return ModIndex.sshash_from_unitig_set(unitigs, w, SKEW_NONE, seed, device)


class PFHash:
    @staticmethod
    def from_unitig_set(unitigs, device=0, builder="host"):
        return ModIndex.pfhash_from_unitig_set(unitigs, device, builder)


def pack_reads(bases, read_len, words=None, n_mask=None, want_mask=True):
    """mazu_b200_pack_reads: ASCII uniform reads -> (2-bit words, N mask or None, number of non-ACGT bases)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8) if isinstance(bases, np.ndarray) else bases
    n_reads = (len(bases) if isinstance(bases, np.ndarray) else bases.numel()) // read_len
    wpr, mpr = (read_len + 31) // 32, (read_len + 63) // 64
    words = np.zeros(n_reads * wpr, dtype=np.uint64) if words is None else words
    if n_mask is None and want_mask:
        n_mask = np.zeros(n_reads * mpr, dtype=np.uint64)
    bad = C.c_uint64(0)
    _check(lib().mazu_b200_pack_reads(_any_ptr(bases), n_reads, read_len, _any_ptr(words), _any_ptr(n_mask), C.byref(bad)))
    return words, n_mask, int(bad.value)


def gather_probe(table_bytes, n_items, granule_bytes=32, ilp=1, blocks_per_sm=8, iters=3, device=0):
    """One point of the random-access roofline sweep (include/mazu_b200_debug.h): granules per second."""
    out = C.c_double(0.0)
    _check(lib().mazu_b200_debug_gather_probe(table_bytes, n_items, granule_bytes, ilp, blocks_per_sm, iters, device, C.byref(out)))
    return out.value
