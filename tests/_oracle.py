"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

The oracle is the checker; nothing in mazu_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")

HIT = np.dtype([("unitig_id", "<u4"), ("unitig_len", "<u4"), ("pos", "<u4"), ("match", "<u4")])
OCC = np.dtype([("ref_id", "<u4"), ("pos", "<u4"), ("fw", "<u4")])
TILE = np.dtype([("unitig_id", "<u4"), ("unitig_len", "<u4"), ("pos", "<u4"), ("fw", "<u4")])
MISS = 0xFFFFFFFF
MATCH_NONE, MATCH_IDENTITY, MATCH_TWIN, MATCH_SKIPPED = 0, 1, 2, 3
USIZE_MAX = 0xFFFFFFFFFFFFFFFF

_lib = None


def build():
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("capi.cpp", "mazu_oracle.hpp", "Makefile")]
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        u64, u32, i32, vp, cp = C.c_uint64, C.c_uint32, C.c_int, C.c_void_p, C.c_char_p
        sig = {
            "orc_last_error": (cp, []),
            "orc_simple_hash64": (u64, [u64, u64]),
            "orc_multihash_chain": (None, [u64, u64, vp]),
            "orc_boophf_load": (vp, [cp]),
            "orc_boophf_free": (None, [vp]),
            "orc_boophf_lookup": (i32, [vp, u64, vp]),
            "orc_boophf_final_lookup": (i32, [vp, u64, vp]),
            "orc_boophf_info": (u64, [vp, i32, u64]),
            "orc_boophf_check_ranks": (i32, [vp]),
            "orc_ef_roundtrip": (i32, [vp, u64, u64, i32, vp, vp]),
            "orc_compact_vector_read": (i32, [cp, vp, vp, vp, u64]),
            "orc_encode_pf1": (u64, [u32, u32, u32]),
            "orc_decode_pf1": (None, [u64, vp]),
            "orc_encode_piscem": (u64, [u32, u32, u32, u64]),
            "orc_decode_piscem": (None, [u64, u64, u64, vp]),
            "orc_required_num_bits": (i32, [u64, u64, vp]),
            "orc_revcomp": (u64, [u64, i32]),
            "orc_revcomp_loop": (u64, [u64, i32]),
            "orc_set_build_threads": (i32, [i32]),
            "orc_mm_hash32": (u64, [u64, u64]),
            "orc_canonical_minimizer": (None, [u64, i32, i32, u64, vp, vp]),
            "orc_encode_read": (u64, [vp, u64, i32, i32, u64, vp, vp, vp, vp, vp]),
            "orc_dense_index_load": (vp, [cp]),
            "orc_sparse_index_load": (vp, [cp]),
            "orc_index_from_cf": (vp, [cp, i32, i32, u64, u64]),
            "orc_index_from_packed": (vp, [i32, vp, u64, vp, u64, i32, i32, u64, u64]),
            "orc_index_from_seqs": (vp, [cp, vp, u64, i32, i32, i32, u64, u64]),
            "orc_index_rebuild_k2u": (vp, [vp, i32, i32, u64, u64]),
            "orc_attach_u2pos": (i32, [vp, i32, vp, u64, vp, vp, vp, u64, u64, u64]),
            "orc_index_free": (None, [vp]),
            "orc_index_info": (u64, [vp, i32]),
            "orc_unitig_len": (u64, [vp, u64]),
            "orc_unitig_start": (u64, [vp, u64]),
            "orc_pos_to_id": (u64, [vp, u64]),
            "orc_ref_len": (u64, [vp, u64]),
            "orc_copy_useq_words": (None, [vp, vp, u64]),
            "orc_ref_total_len": (u64, [vp]),
            "orc_copy_refseq_words": (None, [vp, vp, u64]),
            "orc_k2u_batch": (i32, [vp, vp, u64, i32, vp, i32]),
            "orc_k2u_fw_batch": (i32, [vp, vp, u64, i32, vp]),
            "orc_kmer_offsets": (u64, [vp, u64, i32, vp]),
            "orc_query_reads": (i32, [vp, vp, vp, u64, i32, i32, vp, vp, vp, i32]),
            "orc_decode_occs": (i32, [vp, vp, u64, vp, vp]),
            "orc_project_hits": (i32, [vp, vp, u64, vp, vp]),
            "orc_validate_self": (i32, [vp, vp]),
            "orc_k2u_validate_self": (i32, [vp, vp]),
            "orc_validate_fasta": (i32, [vp, cp, i32, vp]),
            "orc_iter_unitigs_on_ref": (C.c_longlong, [vp, u64, vp, u64]),
            "orc_get_ref_pos_eager_str": (i32, [vp, cp, vp, i32, vp]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _check(rc):
    if rc != 0:
        raise OracleError(lib().orc_last_error().decode())


def encode_kmer(s):
    w = 0
    for i, ch in enumerate(s.upper()):
        w |= "ACGT".index(ch) << (2 * i)
    return w


class OracleIndex:
    """Handle on an oracle ModIndex (k2u + optional u2pos + refs)."""

    def __init__(self, handle):
        if not handle:
            raise OracleError(lib().orc_last_error().decode())
        self.h = C.c_void_p(handle)

    def __del__(self):
        try:
            if self.h:
                lib().orc_index_free(self.h)
                self.h = None
        except Exception:
            pass

    # --- constructors -------------------------------------------------------------------
    @classmethod
    def dense_from_pf1(cls, d):
        return cls(lib().orc_dense_index_load(d.encode()))

    @classmethod
    def sparse_from_pf1(cls, d):
        return cls(lib().orc_sparse_index_load(d.encode()))

    @classmethod
    def from_cf(cls, prefix, kind, w=0, skew=USIZE_MAX, seed=0):
        return cls(lib().orc_index_from_cf(prefix.encode(), kind, w, skew, seed))

    @classmethod
    def from_seqs(cls, seqs, k, kind, w=0, skew=USIZE_MAX, seed=0):
        concat = "".join(seqs).encode()
        offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(s) for s in seqs])
        return cls(lib().orc_index_from_seqs(concat, _ptr(offs), len(seqs), k, kind, w, skew, seed))

    @classmethod
    def from_packed(cls, k, words, n_bases, accum, kind, w=0, skew=USIZE_MAX, seed=0, build_threads=1):
        """build_threads > 1: the one-time builder runs threaded (the reference's builder is rayon-parallel); same tables."""
        words = np.ascontiguousarray(words, dtype=np.uint64)
        accum = np.ascontiguousarray(accum, dtype=np.uint64)
        prev = lib().orc_set_build_threads(build_threads)
        try:
            return cls(lib().orc_index_from_packed(k, _ptr(words), n_bases, _ptr(accum), len(accum) - 1, kind, w, skew, seed))
        finally:
            lib().orc_set_build_threads(prev)

    def rebuild_k2u(self, kind, w=0, skew=USIZE_MAX, seed=0):
        return OracleIndex(lib().orc_index_rebuild_k2u(self.h, kind, w, skew, seed))

    def attach_u2pos(self, kind, offsets, ref_ids, poss, fws, max_ref_len, n_refs):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        ref_ids = np.ascontiguousarray(ref_ids, dtype=np.uint32)
        poss = np.ascontiguousarray(poss, dtype=np.uint32)
        fws = np.ascontiguousarray(fws, dtype=np.uint8)
        _check(lib().orc_attach_u2pos(self.h, kind, _ptr(offsets), len(offsets) - 1, _ptr(ref_ids), _ptr(poss), _ptr(fws),
                                      len(ref_ids), max_ref_len, n_refs))

    # --- info ---------------------------------------------------------------------------
    def info(self, what):
        return int(lib().orc_index_info(self.h, what))

    k = property(lambda s: s.info(0))
    n_unitigs = property(lambda s: s.info(1))
    n_kmers = property(lambda s: s.info(2))
    total_len = property(lambda s: s.info(3))
    n_minimizers = property(lambda s: s.info(4))
    n_kmers_in_skew_index = property(lambda s: s.info(5))
    n_refs = property(lambda s: s.info(6))
    n_total_occs = property(lambda s: s.info(7))

    def unitig_len(self, ui):
        return int(lib().orc_unitig_len(self.h, ui))

    def unitig_start(self, ui):
        return int(lib().orc_unitig_start(self.h, ui))

    def pos_to_id(self, pos):
        return int(lib().orc_pos_to_id(self.h, pos))

    def ref_len(self, ri):
        return int(lib().orc_ref_len(self.h, ri))

    def useq_words(self):
        n = (2 * self.total_len + 63) // 64 + 2
        out = np.zeros(n, dtype=np.uint64)
        lib().orc_copy_useq_words(self.h, _ptr(out), n)
        return out

    def refseq_words(self):
        n = (2 * int(lib().orc_ref_total_len(self.h)) + 63) // 64 + 2
        out = np.zeros(n, dtype=np.uint64)
        lib().orc_copy_refseq_words(self.h, _ptr(out), n)
        return out

    def unitig_starts(self):
        return np.array([self.unitig_start(i) for i in range(self.n_unitigs)] + [self.total_len], dtype=np.uint64)

    def ref_prefix(self):
        p = [0]
        for r in range(self.n_refs):
            p.append(p[-1] + self.ref_len(r))
        return np.array(p, dtype=np.uint64)

    # --- queries ------------------------------------------------------------------------
    def k2u_batch(self, fw_words, qk=None, n_threads=1):
        fw_words = np.ascontiguousarray(fw_words, dtype=np.uint64)
        out = np.zeros(len(fw_words), dtype=HIT)
        _check(lib().orc_k2u_batch(self.h, _ptr(fw_words), len(fw_words), self.k if qk is None else qk, _ptr(out), n_threads))
        return out

    def k2u_fw_batch(self, fw_words, qk=None):
        fw_words = np.ascontiguousarray(fw_words, dtype=np.uint64)
        out = np.zeros(len(fw_words), dtype=HIT)
        _check(lib().orc_k2u_fw_batch(self.h, _ptr(fw_words), len(fw_words), self.k if qk is None else qk, _ptr(out)))
        return out

    def kmer_offsets(self, read_offsets):
        read_offsets = np.ascontiguousarray(read_offsets, dtype=np.uint64)
        out = np.zeros(len(read_offsets), dtype=np.uint64)
        lib().orc_kmer_offsets(_ptr(read_offsets), len(read_offsets) - 1, self.k, _ptr(out))
        return out

    def query_reads(self, bases, read_offsets, streaming=False, reset_per_read=True, want_hits=True, n_threads=1):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        read_offsets = np.ascontiguousarray(read_offsets, dtype=np.uint64)
        koffs = self.kmer_offsets(read_offsets)
        out = np.zeros(int(koffs[-1]), dtype=HIT) if want_hits else None
        counts = np.zeros(3, dtype=np.uint64)
        _check(lib().orc_query_reads(self.h, _ptr(bases), _ptr(read_offsets), len(read_offsets) - 1, int(streaming),
                                     int(reset_per_read), _ptr(koffs), _ptr(out), _ptr(counts), n_threads))
        return out, counts, koffs

    def decode_occs(self, unitig_ids):
        unitig_ids = np.ascontiguousarray(unitig_ids, dtype=np.uint32)
        offs = np.zeros(len(unitig_ids) + 1, dtype=np.uint64)
        _check(lib().orc_decode_occs(self.h, _ptr(unitig_ids), len(unitig_ids), _ptr(offs), None))
        out = np.zeros(int(offs[-1]), dtype=OCC)
        _check(lib().orc_decode_occs(self.h, _ptr(unitig_ids), len(unitig_ids), _ptr(offs), _ptr(out)))
        return offs, out

    def project_hits(self, hits):
        hits = np.ascontiguousarray(hits, dtype=HIT)
        offs = np.zeros(len(hits) + 1, dtype=np.uint64)
        _check(lib().orc_project_hits(self.h, _ptr(hits), len(hits), _ptr(offs), None))
        out = np.zeros(int(offs[-1]), dtype=OCC)
        _check(lib().orc_project_hits(self.h, _ptr(hits), len(hits), _ptr(offs), _ptr(out)))
        return offs, out

    def validate_self(self):
        c = np.zeros(5, dtype=np.uint64)
        _check(lib().orc_validate_self(self.h, _ptr(c)))
        return [int(x) for x in c]

    def k2u_validate_self(self):
        c = np.zeros(5, dtype=np.uint64)
        _check(lib().orc_k2u_validate_self(self.h, _ptr(c)))
        return [int(x) for x in c]

    def validate_fasta(self, path, streaming=False):
        c = np.zeros(5, dtype=np.uint64)
        _check(lib().orc_validate_fasta(self.h, path.encode(), int(streaming), _ptr(c)))
        return [int(x) for x in c]

    def iter_unitigs_on_ref(self, ref_id):
        cap = self.ref_len(ref_id) + 1
        out = np.zeros((cap, 4), dtype=np.uint32)
        n = lib().orc_iter_unitigs_on_ref(self.h, ref_id, _ptr(out), cap)
        if n < 0:
            raise OracleError(lib().orc_last_error().decode())
        return out[:n].copy().view(TILE).reshape(-1)

    def get_ref_pos_eager(self, kmer):
        """Returns None (miss) or (hit_record, [(ref_id,pos,fw),...]); raises on contract panic."""
        out = np.zeros(64, dtype=OCC)
        hit = np.zeros(1, dtype=HIT)
        n = lib().orc_get_ref_pos_eager_str(self.h, kmer.encode(), _ptr(out), 64, _ptr(hit))
        if n == -2:
            raise OracleError(lib().orc_last_error().decode())
        if n == -1:
            return None
        return hit[0], [(int(o["ref_id"]), int(o["pos"]), int(o["fw"])) for o in out[:n]]
