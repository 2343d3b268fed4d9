/*
 * mazu_b200.h -- C ABI of libmazu_b200.so: the B200 (sm_100a) implementation of mazu's batched
 * k-mer query path.  Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * The reference (COMBINE-lab/mazu, Rust) has no FFI; its "operator API" for this path is the
 * trait surface K2U / U2Pos / GetRefPos / Validate and four constructors.  Each entry point
 * below names the reference interface it replaces (path:line under the mazu repository).
 * INTEGRATION.md shows the Rust `extern "C"` block + `impl K2U for GpuIndex` a maintainer would add.
 *
 * Conventions
 *  - every function returns a mazu_status_t (0 = ok, negative = error); the message of the last
 *    error on the calling thread is available from mazu_b200_last_error().
 *  - contract violations that PANIC in the reference (wrong k: src/index.rs:157-163,
 *    src/kphf/sshash.rs:473, src/kphf/pfhash.rs:109) return MAZU_ERR_K_MISMATCH, never UB.
 *  - a miss (`None` in the reference) is NOT an error: it is the record {~0,~0,~0,MAZU_NO_MATCH}.
 *  - `mem` says where the caller's buffers live: MAZU_MEM_HOST (the library stages them through
 *    its own device buffers, chunked and overlapped) or MAZU_MEM_DEVICE (device pointers, e.g.
 *    torch tensors' data_ptr(); work is enqueued on `stream` and the call returns without
 *    synchronising).  mazu_b200_query_reads[_compact] also take MAZU_MEM_HOST_IN_DEVICE_OUT: reads
 *    (bases, read_offsets, kmer_offsets, counts) on the host, out_hits a DEVICE pointer -- the hit
 *    records stay in HBM for the next device stage (mazu_b200_project_hits with MAZU_MEM_DEVICE)
 *    and only the three counters come back; the call returns after the last chunk has finished.
 *  - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *  - an index handle is immutable after creation and may be used concurrently from several host
 *    threads / streams (reference: queries take &self and are Sync, src/kphf/mod.rs:69-72).
 */
#ifndef MAZU_B200_H
#define MAZU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t mazu_status_t;
enum {
  MAZU_OK = 0,
  MAZU_ERR_IO = -1,               /* src/err.rs: Error::IO                        */
  MAZU_ERR_INVALID_DATA = -2,     /* src/err.rs: Error::InvalidData               */
  MAZU_ERR_EF_NOT_MONOTONE = -3,  /* src/err.rs: Error::EFNotMonotone             */
  MAZU_ERR_EF_EMPTY = -4,         /* src/err.rs: Error::EFEmpty                   */
  MAZU_ERR_CUDA = -5,             /* CUDA runtime / no device                     */
  MAZU_ERR_K_MISMATCH = -6,       /* reference panics (see above)                 */
  MAZU_ERR_INVALID_ARG = -7,
  MAZU_ERR_NO_U2POS = -8,         /* occurrence query on an index without a U2Pos */
  MAZU_ERR_NO_REFSEQ = -9,        /* validate_self needs reference sequence (src/index/validate.rs:25) */
  MAZU_ERR_OTHER = -10            /* src/err.rs: Error::Other                     */
};

/* MatchType of the `kmers` crate as used by K2UPos.o (src/kphf/mod.rs:13-19) */
enum { MAZU_NO_MATCH = 0, MAZU_IDENTITY_MATCH = 1, MAZU_TWIN_MATCH = 2,
       MAZU_SKIPPED = 3 /* record filler for windows CanonicalKmerIterator skips (non-ACGT) */ };

enum { MAZU_MEM_HOST = 0, MAZU_MEM_DEVICE = 1, MAZU_MEM_HOST_IN_DEVICE_OUT = 2 };
enum { MAZU_MODE_RANDOM = 0,    /* K2U::k2u per k-mer                 (src/bin/kphf/main.rs:311-322) */
       MAZU_MODE_STREAMING = 1  /* .as_streaming() / StreamingK2U     (src/index/caching.rs:65-103); cursor reset per read */ };
/* (A streaming answer differs from K2U::k2u only for k-mers the unitig set holds twice: see MAZU_INFO_KMERS_UNIQUE.) */
enum { MAZU_K2U_PFHASH = 0, MAZU_K2U_SSHASH = 1,
       MAZU_K2U_SAMPLED_PFHASH = 2 /* pufferfish sparse index, load-only (src/kphf/pfhash.rs:137-285) */ };
enum { MAZU_U2POS_NONE = 0, MAZU_U2POS_DENSE = 1, MAZU_U2POS_PISCEM = 2 };
enum { MAZU_INDEX_PUFFERFISH_DENSE = 0, /* ModIndex<PFHash, DenseUnitigTable>  src/index/defaults.rs:14 */
       MAZU_INDEX_PISCEM = 1            /* ModIndex<SSHash, PiscemUnitigTable> src/index/defaults.rs:15 */ };

#define MAZU_SKEW_NONE UINT64_MAX /* skew_param == usize::MAX: no skew index (src/kphf/sshash.rs:81-84,222) */

/* Option<K2UPos> (src/kphf/mod.rs:13-19); miss = {~0,~0,~0,MAZU_NO_MATCH} (src/index/caching.rs:48-53) */
typedef struct mazu_hit {
  uint32_t unitig_id;
  uint32_t unitig_len;
  uint32_t pos;
  uint32_t match;
} mazu_hit_t;

/* Compact K2UPos for bandwidth-bound callers (PCIe): unitig_len is omitted because the caller owns the UnitigSet
 * (K2U::unitig_len(unitig_id), src/kphf/mod.rs:61).  pos_match = pos | match << 30; miss = {~0, 0x3FFFFFFF | MAZU_NO_MATCH << 30};
 * skipped window = {~0, 0x3FFFFFFF | MAZU_SKIPPED << 30}.  Requires every unitig to be shorter than 2^30 bases. */
typedef struct mazu_hit8 {
  uint32_t unitig_id;
  uint32_t pos_match;
} mazu_hit8_t;

/* UnitigOcc (src/index.rs:304-309) and MappedRefPos (src/index.rs:25-31); fw: Forward=1, Backward=0 (src/lib.rs:51-56) */
typedef struct mazu_occ {
  uint32_t ref_id;
  uint32_t pos;
  uint32_t fw;
} mazu_occ_t;

typedef struct mazu_index mazu_index_t; /* opaque: owns the device-resident index */

/* UnitigSet (src/unitig_set.rs:31-36) as plain host arrays */
typedef struct mazu_unitig_set_desc {
  uint32_t k;
  const uint64_t* useq_words; /* 2-bit packed concatenated unitigs, base i at bits [2i,2i+2), ceil(2*n_bases/64) words */
  uint64_t n_bases;
  const uint64_t* accum_lens; /* n_unitigs + 1 prefix lengths, accum_lens[0] == 0 */
  uint64_t n_unitigs;
} mazu_unitig_set_desc_t;

/* simple-sds IntVector / pufferfish compact vector: `len` fields of `width` bits, LSB-first */
typedef struct mazu_packed_vec_desc {
  const uint64_t* words;
  uint64_t width;
  uint64_t len;
} mazu_packed_vec_desc_t;

/* C++ BooPHF as parsed by src/pf1/boophf/mod.rs:50-86,269-293 */
typedef struct mazu_boophf_desc {
  uint32_t n_levels;
  const uint64_t* const* level_words; /* per level: bit array words            */
  const uint64_t* level_n_bits;       /* per level: number of slots            */
  uint64_t last_bitset_rank;
  uint64_t n_elem;
  const uint64_t* final_keys;         /* final_hash map, any order             */
  const uint64_t* final_vals;
  uint64_t n_final;
} mazu_boophf_desc_t;

const char* mazu_b200_last_error(void);
/* number of visible CUDA devices; <=0 means the library cannot run (it has NO CPU fallback) */
int32_t mazu_b200_device_count(void);

/* ---------------------------------------------------------------------------------------------
 * Construction (host work + one upload).  `device` = CUDA device ordinal.
 * ------------------------------------------------------------------------------------------- */
/* DenseIndex::deserialize_from_cpp(dir)                         src/pf1/dense_index.rs:33-97 */
mazu_status_t mazu_b200_dense_index_deserialize_from_cpp(const char* dir, int32_t device, mazu_index_t** out);
/* SparseIndex::deserialize_from_cpp(dir): ModIndex<SampledPFHash<BooPHF>, DenseUnitigTable>   src/pf1/sparse_index.rs:32-110 */
mazu_status_t mazu_b200_sparse_index_deserialize_from_cpp(const char* dir, int32_t device, mazu_index_t** out);
/* PufferfishDenseIndexDefault::from_cf_prefix / PiscemIndex::from_cf_prefix
 *                                   src/index/defaults.rs:17-58, src/index/piscem_index.rs:14-58 */
mazu_status_t mazu_b200_index_from_cf_prefix(const char* prefix, int32_t index_kind, uint32_t w, uint64_t skew_param,
                                             uint64_t hash_seed, int32_t device, mazu_index_t** out);
/* SSHash::from_unitig_set(unitigs, w, skew_param, WyHashState(seed))   src/kphf/sshash.rs:405-412
 * (skew_param == MAZU_SKEW_NONE: from_unitig_set_no_skew_index, :397-403) */
mazu_status_t mazu_b200_index_create_sshash(const mazu_unitig_set_desc_t* unitigs, uint32_t w, uint64_t skew_param,
                                            uint64_t hash_seed, int32_t device, mazu_index_t** out);
/* the same constructor with every build stage on the device (minimizer collection, radix sort, MPHF cascade, scans,
 * Elias-Fano encoding, packing): produces tables bit-identical to mazu_b200_index_create_sshash    */
mazu_status_t mazu_b200_index_create_sshash_gpu(const mazu_unitig_set_desc_t* unitigs, uint32_t w, uint64_t skew_param,
                                                uint64_t hash_seed, int32_t device, mazu_index_t** out);
/* PFHash::from_unitig_set on the device (tables bit-identical to mazu_b200_index_create_pfhash) */
mazu_status_t mazu_b200_index_create_pfhash_gpu(const mazu_unitig_set_desc_t* unitigs, int32_t device, mazu_index_t** out);
/* PFHash::from_unitig_set(unitigs)                                     src/kphf/pfhash.rs:40-73 */
mazu_status_t mazu_b200_index_create_pfhash(const mazu_unitig_set_desc_t* unitigs, int32_t device, mazu_index_t** out);
/* PFHash::from_parts(unitigs, BooPHF, pos)                             src/kphf/pfhash.rs:34-36 */
mazu_status_t mazu_b200_index_create_pfhash_from_parts(const mazu_unitig_set_desc_t* unitigs, const mazu_boophf_desc_t* mphf,
                                                       const mazu_packed_vec_desc_t* pos, int32_t device, mazu_index_t** out);
/* ModIndex::from_parts(base, SSHash::from_unitig_set(idx.as_ref().clone(), ..), u2pos.clone(), refs.clone())
 * as in src/pf1/dense_index.rs:315-328: new handle with a rebuilt K2U sharing U2Pos + refs */
mazu_status_t mazu_b200_index_rebuild_k2u(const mazu_index_t* src, int32_t k2u_kind, uint32_t w, uint64_t skew_param,
                                          uint64_t hash_seed, mazu_index_t** out);
/* the U2Pos half of ModIndex::from_parts (src/index.rs:80-87):
 * DenseUnitigTable{ctable: Vec<u64>, contig_offsets}    src/index/dense_unitig_table.rs:13-18 */
mazu_status_t mazu_b200_index_attach_u2pos_dense(mazu_index_t* idx, const uint64_t* ctable, uint64_t n_occs,
                                                 const mazu_packed_vec_desc_t* contig_offsets);
/* PiscemUnitigTable{ref_shift,pos_mask,ctable: IntVector,contig_offsets}  src/index/dense_unitig_table.rs:109-118 */
mazu_status_t mazu_b200_index_attach_u2pos_piscem(mazu_index_t* idx, const mazu_packed_vec_desc_t* ctable, uint64_t ref_shift,
                                                  uint64_t pos_mask, const mazu_packed_vec_desc_t* contig_offsets);
/* RefSeqCollection::from_parts(seq, prefix_sum)                        src/refseq.rs:124-126 */
mazu_status_t mazu_b200_index_attach_refseq(mazu_index_t* idx, const uint64_t* seq_words, const uint64_t* prefix_sum,
                                            uint64_t n_refs);
void mazu_b200_index_destroy(mazu_index_t* idx);
/* Calls with host buffers stage through device scratch taken from a pool the handle owns; the pool keeps what it
 * has used (about 1 GB after a large mazu_b200_query_reads) so later calls pay no allocation.  This returns the
 * idle scratch to the driver; *released (optional) receives the bytes given back. */
mazu_status_t mazu_b200_index_release_scratch(mazu_index_t* idx, uint64_t* released);

/* ---------------------------------------------------------------------------------------------
 * K2U accessors (src/kphf/mod.rs:58-67) and index stats
 * ------------------------------------------------------------------------------------------- */
enum { MAZU_INFO_K = 0, MAZU_INFO_N_UNITIGS = 1, MAZU_INFO_N_KMERS = 2, MAZU_INFO_SUM_UNITIGS_LEN = 3,
       MAZU_INFO_N_MINIMIZERS = 4 /* SSHash::n_minimizers, sshash.rs:333-335 */,
       MAZU_INFO_N_KMERS_IN_SKEW_INDEX = 5 /* sshash.rs:337-339 */, MAZU_INFO_N_REFS = 6, MAZU_INFO_N_TOTAL_OCCS = 7,
       MAZU_INFO_K2U_KIND = 8, MAZU_INFO_U2POS_KIND = 9, MAZU_INFO_DEVICE_BYTES = 10, MAZU_INFO_W = 11,
       MAZU_INFO_N_MINIMIZER_OCCS = 12, MAZU_INFO_MPHF_LEVELS = 13, MAZU_INFO_DEVICE = 14,
       MAZU_INFO_SAMPLE_SIZE = 15, MAZU_INFO_EXTENSION_SIZE = 16 /* SampledPFHash, src/kphf/pfhash.rs:149-150 */,
       MAZU_INFO_KMERS_UNIQUE = 17 /* 1: the canonical k-mers of the unitig set are pairwise distinct (checked on the device at
                                      creation).  MAZU_MODE_STREAMING then returns the records of MAZU_MODE_RANDOM by construction and is
                                      served by the random-access kernel; the cursor-walk kernel runs for sets with duplicated k-mers */ };
uint64_t mazu_b200_index_info(const mazu_index_t* idx, int32_t what);
/* K2U::unitig_len(id)  src/kphf/mod.rs:61 ; UnitigSet::unitig_start_pos  src/unitig_set.rs:197-199 */
mazu_status_t mazu_b200_unitig_len(const mazu_index_t* idx, uint64_t unitig_id, uint64_t* len, uint64_t* start_pos);
/* K2U::unitig_seq(id) (src/kphf/mod.rs:64; UnitigSet::unitig_seq src/unitig_set.rs:189-195): the unitig's bases as 2-bit codes
 * packed LSB-first into out_words (base j at bits [2j, 2j+2) of the word array, ceil(len / 32) words; cap_words is the
 * capacity).  Served from the handle's host-side UnitigSet, so a Rust `impl K2U` needs no copy of its own. */
mazu_status_t mazu_b200_unitig_seq(const mazu_index_t* idx, uint64_t unitig_id, uint64_t* out_words, uint64_t cap_words, uint64_t* len);

/* ---------------------------------------------------------------------------------------------
 * Queries
 * ------------------------------------------------------------------------------------------- */
/* K2U::k2u for a batch of forward k-mer words (src/kphf/mod.rs:66; PFHash src/kphf/pfhash.rs:108-134,
 * SSHash src/kphf/sshash.rs:471-555).  fw_words[i] holds a k-mer with base j at bits [2j,2j+2);
 * `k` is the query k-mer length (must equal the index's k).  A Rust `impl K2U` calls this with n = 1. */
mazu_status_t mazu_b200_k2u_batch(const mazu_index_t* idx, const uint64_t* fw_words, uint64_t n, uint32_t k,
                                  mazu_hit_t* out_hits, int32_t mem, void* stream);

/* The read loop of `kphf bench` / validate_ckmers (src/bin/kphf/main.rs:299-322,
 * src/index/validate.rs:54-81, src/index/caching.rs:175-201): CanonicalKmerIterator::from_u8_slice
 * over each read + k2u (MAZU_MODE_RANDOM) or StreamingK2U::k2u_streaming (MAZU_MODE_STREAMING).
 *   bases         ASCII reads, concatenated
 *   read_offsets  n_reads+1 byte offsets; may be NULL when uniform_read_len > 0 (read r = bases[r*len, (r+1)*len))
 *   kmer_offsets  optional out, n_reads+1: slot of read r's first k-mer = sum_{r'<r} max(len-k+1,0)
 *   out_hits      optional out, one record per k-mer slot (slot = kmer_offsets[r] + position in read);
 *                 windows with a non-ACGT base get {~0,~0,~0,MAZU_SKIPPED}
 *   counts        optional out {n_kmers (valid windows), n_hit, n_miss} -- the counters of main.rs:282-284
 * In MAZU_MEM_DEVICE all four are device pointers and counts is ACCUMULATED into (zero it first). */
mazu_status_t mazu_b200_query_reads(const mazu_index_t* idx, const uint8_t* bases, const uint64_t* read_offsets,
                                    uint64_t n_reads, uint64_t uniform_read_len, int32_t mode, uint64_t* kmer_offsets,
                                    mazu_hit_t* out_hits, uint64_t* counts, int32_t mem, void* stream);
/* same call, 8-byte mazu_hit8_t records (halves the device->host traffic of the 16-byte form) */
mazu_status_t mazu_b200_query_reads_compact(const mazu_index_t* idx, const uint8_t* bases, const uint64_t* read_offsets,
                                            uint64_t n_reads, uint64_t uniform_read_len, int32_t mode, uint64_t* kmer_offsets,
                                            mazu_hit8_t* out_hits, uint64_t* counts, int32_t mem, void* stream);
/* same call for PCIe-bound callers (host buffers only): the hit records as RUNS.  Consecutive k-mers of a read that hit walk along
 * one unitig, so slot s "continues" slot s-1 when both hit the same unitig with the same match type and pos differs by +1
 * (Identity) or -1 (Twin).  Outputs:
 *   out_codes             one byte per k-mer slot: 0 miss, 1 hit continuing the previous slot's run, 2 hit starting a run, 3 skipped
 *   out_runs              the full mazu_hit_t of every run start (cap_runs records; at most one per slot); the runs of one read are
 *                         consecutive and in slot order, the order of different reads' runs is unspecified
 *   out_read_run_offsets  n_reads + 1: entry r = index in out_runs of read r's first run (not monotone in r), entry n_reads = number of runs
 *   out_n_runs            runs written; if it exceeds cap_runs the call fails with MAZU_ERR_INVALID_ARG and this is the capacity needed
 * ~1.1 bytes per lookup cross PCIe instead of 16; mazu_b200_expand_hit_runs rebuilds the exact mazu_hit_t array on the host. */
mazu_status_t mazu_b200_query_reads_runs(const mazu_index_t* idx, const uint8_t* bases, const uint64_t* read_offsets,
                                         uint64_t n_reads, uint64_t uniform_read_len, int32_t mode, uint64_t* kmer_offsets,
                                         uint8_t* out_codes, mazu_hit_t* out_runs, uint64_t cap_runs, uint64_t* out_read_run_offsets,
                                         uint64_t* out_n_runs, uint64_t* counts);
/* The same call for hosts where PCIe / host memory bandwidth is the limit (several GPUs behind one host): reads arrive 2-bit
 * packed and the run codes leave 2 bits per slot -- 0.31 + 0.25 bytes per lookup cross PCIe instead of 1.25 + 1.
 *   packed_reads  uniform reads in the kmers::SeqVector layout: read r = words [r * wpr, (r + 1) * wpr), wpr = ceil(read_len / 32),
 *                 base j at bits [2j, 2j + 2) of its words, A=0 C=1 G=2 T=3 (mazu_b200_pack_reads makes it from ASCII)
 *   n_mask        optional, 1 bit per base, ceil(read_len / 64) words per read: set = the base is not ACGT (NULL: none is)
 *   out_codes2    2 bits per k-mer slot: slot s at bits [2 (s % 4), 2 (s % 4) + 2) of byte s / 4, same code values
 * read_len - k + 1 must be a multiple of 4.  mazu_b200_expand_hit_runs_packed rebuilds every mazu_hit_t on the host. */
mazu_status_t mazu_b200_query_reads_runs_packed(const mazu_index_t* idx, const uint64_t* packed_reads, const uint64_t* n_mask,
                                                uint64_t n_reads, uint64_t read_len, int32_t mode, uint8_t* out_codes2,
                                                mazu_hit_t* out_runs, uint64_t cap_runs, uint64_t* out_read_run_offsets,
                                                uint64_t* out_n_runs, uint64_t* counts);
/* ASCII -> 2-bit words (+ N mask) on the host, multi-threaded; *n_non_acgt (optional) counts the masked bases */
mazu_status_t mazu_b200_pack_reads(const uint8_t* bases, uint64_t n_reads, uint64_t read_len, uint64_t* out_words, uint64_t* out_n_mask,
                                   uint64_t* n_non_acgt);
mazu_status_t mazu_b200_expand_hit_runs_packed(const uint8_t* codes2, const mazu_hit_t* runs, const uint64_t* read_run_offsets,
                                               uint64_t n_reads, uint64_t uniform_slots, mazu_hit_t* out_hits);
/* The sparsest lossless host result of a short-read batch: one self-contained 16-byte record per RUN and nothing per k-mer
 * slot or per read -- ~0.19 bytes per lookup on the bench's read mix (the run format above: 0.5).  For hosts that feed several
 * GPUs, where the host's D2H bandwidth, not the GPUs, bounds the run format (DESIGN.md "Multi-GPU").
 * A run is a maximal stretch of consecutive k-mer slots of one read whose hits walk along one unitig in one orientation
 * (the same relation as in mazu_b200_query_reads_runs); slots in no run are misses or -- where the window holds a non-ACGT
 * base, which the caller's n_mask says -- skipped.  The order of the records is unspecified.
 *   packed_reads, n_mask  as in mazu_b200_query_reads_runs_packed
 *   out_intervals         capacity `cap` records; *out_n receives the number of runs (if it exceeds cap the call fails with
 *                         MAZU_ERR_INVALID_ARG and *out_n is the capacity needed)
 * Restrictions (MAZU_ERR_INVALID_ARG otherwise): read_len - k + 1 <= 128 (short reads: one read per warp), n_reads < 2^32,
 * and MAZU_MODE_STREAMING only on indexes whose k-mers are unique (MAZU_INFO_KMERS_UNIQUE; its answers are then k2u's).
 * mazu_b200_expand_hit_intervals rebuilds every mazu_hit_t on the host (unitig lengths from the handle's UnitigSet). */
typedef struct mazu_hit_interval {
  uint32_t unitig_id;
  uint32_t pos_o;  /* bits 0-30: K2UPos::pos of the run's first k-mer; bit 31: 1 = MatchType::TwinMatch (pos falls by 1 per slot), 0 = IdentityMatch (rises) */
  uint32_t read;   /* index of the read in the batch */
  uint16_t start;  /* first k-mer slot of the run inside the read */
  uint16_t len;    /* number of slots */
} mazu_hit_interval_t;
mazu_status_t mazu_b200_query_reads_intervals_packed(const mazu_index_t* idx, const uint64_t* packed_reads, const uint64_t* n_mask,
                                                     uint64_t n_reads, uint64_t read_len, int32_t mode,
                                                     mazu_hit_interval_t* out_intervals, uint64_t cap, uint64_t* out_n, uint64_t* counts);
/* The same result for reads as the reference takes them -- ASCII, n_reads x read_len bytes (the sequence slices handed to
 * CanonicalKmerIterator::from_u8_slice in src/bin/kphf/main.rs:303,314 and src/index/validate.rs:57), any case, non-ACGT bases
 * skipped: 1.25 bytes per lookup in, ~0.19 out, nothing for the caller to pack.  Same restrictions as the packed call. */
mazu_status_t mazu_b200_query_reads_intervals(const mazu_index_t* idx, const uint8_t* bases, uint64_t n_reads, uint64_t read_len, int32_t mode,
                                              mazu_hit_interval_t* out_intervals, uint64_t cap, uint64_t* out_n, uint64_t* counts);
/* host-side decoder (multi-threaded): out_hits[r * (read_len - k + 1) + slot]; n_mask as passed to the query (or NULL) */
mazu_status_t mazu_b200_expand_hit_intervals(const mazu_index_t* idx, const mazu_hit_interval_t* intervals, uint64_t n_intervals,
                                             const uint64_t* n_mask, uint64_t n_reads, uint64_t read_len, mazu_hit_t* out_hits);
/* ... and for the ASCII call: the skipped windows are read off the caller's bases */
mazu_status_t mazu_b200_expand_hit_intervals_ascii(const mazu_index_t* idx, const mazu_hit_interval_t* intervals, uint64_t n_intervals,
                                                   const uint8_t* bases, uint64_t n_reads, uint64_t read_len, mazu_hit_t* out_hits);
/* host-side decoder of the run format (multi-threaded, no device work): out_hits[slot] for every slot of every read.
 * kmer_offsets may be NULL for uniform batches (uniform_slots = read_len - k + 1). */
mazu_status_t mazu_b200_expand_hit_runs(const uint8_t* codes, const mazu_hit_t* runs, const uint64_t* read_run_offsets,
                                        const uint64_t* kmer_offsets, uint64_t n_reads, uint64_t uniform_slots, mazu_hit_t* out_hits);
/* number of k-mer slots query_reads writes for this batch (host arithmetic; read_offsets on host or NULL) */
uint64_t mazu_b200_count_kmer_slots(const mazu_index_t* idx, const uint64_t* read_offsets, uint64_t n_reads,
                                    uint64_t uniform_read_len);

/* Stage 1 alone (kmers::CanonicalKmerIterator + Kmer::canonical_minimizer):
 * per k-mer slot of ONE uniform batch: fw word, rc word, minimizer word, minimizer offset, valid flag.
 * Device pointers only; any output may be NULL. For PFHash indexes (no w) the minimizer outputs are zero. */
mazu_status_t mazu_b200_encode_reads(const mazu_index_t* idx, const uint8_t* bases, const uint64_t* read_offsets,
                                     uint64_t n_reads, uint64_t uniform_read_len, const uint64_t* kmer_offsets,
                                     uint64_t* out_fw, uint64_t* out_rc, uint64_t* out_mm_word, uint32_t* out_mm_offset,
                                     uint8_t* out_valid, void* stream);

/* U2Pos::encoded_unitig_occs + decode_unitig_occs for a batch of unitig ids
 * (src/index.rs:349-361, src/index/dense_unitig_table.rs:55-76,127-153, UnitigOcc::decode_pf1 src/index.rs:335-346,
 * UnitigOcc::decode_piscem src/spt_compact.rs:99-110).  unitig id ~0 (a miss) yields an empty list.
 *   out_offsets  n+1 prefix of list lengths (always written)
 *   out_occs     capacity `cap` records; may be NULL to only size the output (then *out_total is the need)
 * Returns MAZU_ERR_INVALID_ARG if cap is too small (out_total still valid).  With MAZU_MEM_DEVICE and out_total == NULL
 * the call does not synchronise and cannot return that error: an undersized buffer is then filled up to `cap` records,
 * never beyond, and out_offsets[n] (device) holds the size the full output needs.  Unitig ids outside the table
 * (including ~0) yield empty lists. */
mazu_status_t mazu_b200_decode_occs(const mazu_index_t* idx, const uint32_t* unitig_ids, uint64_t n, uint64_t* out_offsets,
                                    mazu_occ_t* out_occs, uint64_t cap, uint64_t* out_total, int32_t mem, void* stream);
/* GetRefPos::project_hits / project_onto_u_occs for a batch of hits (src/index.rs:174-216): same output
 * convention as decode_occs but records are MappedRefPos.  Misses / skipped records yield empty lists. */
mazu_status_t mazu_b200_project_hits(const mazu_index_t* idx, const mazu_hit_t* hits, uint64_t n, uint64_t* out_offsets,
                                     mazu_occ_t* out_mrps, uint64_t cap, uint64_t* out_total, int32_t mem, void* stream);

/* GetRefPos::get_ref_pos + project_hits over a read batch in ONE pass (src/index.rs:156-216 inside the read loop of
 * validate_ckmers, src/index/validate.rs:54-81): reads -> K2UPos -> occurrence list -> MappedRefPos, without materialising and
 * re-reading the hit records.  Outputs, per k-mer slot (slot = kmer_offsets[r] + position in read r):
 *   out_offsets  n_slots + 1: prefix of the slots' list lengths (misses and skipped windows have empty lists)
 *   out_mrps     the projected positions in slot order, capacity `cap` records
 *   out_hits     optional: the K2UPos records, as mazu_b200_query_reads writes them
 *   out_total    optional: number of records (if given the call synchronises and returns MAZU_ERR_INVALID_ARG when cap is too
 *                small; without it an undersized buffer is filled up to cap, never beyond, and out_offsets[n_slots] tells)
 * n_slots must be mazu_b200_count_kmer_slots() of the batch (the size of out_offsets - 1).  Results equal
 * mazu_b200_query_reads followed by mazu_b200_project_hits record for record; MAZU_MODE_STREAMING on an index with duplicated
 * k-mers runs exactly that chain. */
mazu_status_t mazu_b200_get_ref_pos_reads(const mazu_index_t* idx, const uint8_t* bases, const uint64_t* read_offsets, uint64_t n_reads,
                                          uint64_t uniform_read_len, int32_t mode, uint64_t n_slots, uint64_t* kmer_offsets,
                                          mazu_hit_t* out_hits, uint64_t* out_offsets, mazu_occ_t* out_mrps, uint64_t cap,
                                          uint64_t* out_total, uint64_t* counts, int32_t mem, void* stream);

/* ModIndex::iter_unitigs_on_ref / RefSeqContigIterator (src/index.rs:363-424): the unitig tiling of reference `ref_id`.
 * Records are RefSeqUnitigOcc {unitig_id, unitig_len, pos (on the reference), fw} in a mazu_hit_t (match field = 1 forward,
 * 0 backward).  Every reference position is looked up on the device in one pass; the unitig-jumping walk
 * (pos += unitig_len - k + 1) then runs over those answers.  Host buffers.  A reference k-mer that is not in the index is
 * MAZU_ERR_INVALID_DATA (the reference unwraps a None there).  *n_out receives the number of tiles even if cap is too small. */
mazu_status_t mazu_b200_iter_unitigs_on_ref(const mazu_index_t* idx, uint64_t ref_id, mazu_hit_t* out, uint64_t cap, uint64_t* n_out);

/* Validate::validate_self (src/index/validate.rs:24-52): every reference k-mer must map back to its own
 * (ref_id,pos).  counts = {n_queries, n_identity, n_twin, n_projected, n_fail}; the reference panics when n_fail != 0. */
mazu_status_t mazu_b200_validate_self(const mazu_index_t* idx, uint64_t counts[5]);
/* K2U::validate_self / validate_self_parallel (src/kphf/mod.rs:69-139): every unitig k-mer, fw then swapped.
 * counts[3] = failures that were plain misses (the rest of n_fail found the k-mer at ANOTHER position, i.e. the
 * unitig set holds a duplicated canonical k-mer, which the reference's validate_self would also reject). */
mazu_status_t mazu_b200_k2u_validate_self(const mazu_index_t* idx, uint64_t counts[5]);

/* ---------------------------------------------------------------------------------------------
 * FASTA / FASTQ ingest and Validate::validate_fasta  (the caller on the other side of the path)
 * ------------------------------------------------------------------------------------------- */
/* FastaReader (src/util.rs:93-149): a record starts at a line beginning with '>' (the rest of the line is its name, verbatim),
 * every following line up to the next '>' is appended to its sequence.  A file whose first byte is '@' is read as FASTQ
 * (records of four lines: @name, sequence, +, qualities).  '\r' before a line end is dropped.  The whole file is held in memory. */
typedef struct mazu_fasta mazu_fasta_t;
mazu_status_t mazu_b200_fasta_open(const char* path, mazu_fasta_t** out);
void mazu_b200_fasta_close(mazu_fasta_t* f);
uint64_t mazu_b200_fasta_n_records(const mazu_fasta_t* f);
/* the records' sequences, concatenated (ASCII, as in the file), and n_records + 1 byte offsets: the (bases, read_offsets) pair
 * mazu_b200_query_reads takes.  The pointers stay valid until mazu_b200_fasta_close. */
const uint8_t* mazu_b200_fasta_bases(const mazu_fasta_t* f);
const uint64_t* mazu_b200_fasta_offsets(const mazu_fasta_t* f);
const char* mazu_b200_fasta_name(const mazu_fasta_t* f, uint64_t record);
/* Validate::validate_fasta (src/index/validate.rs:83-100) with mode = MAZU_MODE_RANDOM; StreamingIndex::validate_fasta
 * (src/index/caching.rs:204-218) with MAZU_MODE_STREAMING: record i of the file is reference i, and every k-mer the
 * CanonicalKmerIterator yields at position p of record i (windows with a non-ACGT base are skipped, lower case counts as upper)
 * must be found and project onto (ref i, pos p).  Lookups and the check of the projected positions run on the device.
 * counts = {n_queries, n_identity, n_twin, n_projected, n_fail}; the reference panics when n_fail != 0. */
mazu_status_t mazu_b200_validate_fasta(const mazu_index_t* idx, const char* path, int32_t mode, uint64_t counts[5]);
/* the same check for records already in memory (validate_ckmers over a batch, src/index/validate.rs:54-81): host buffers */
mazu_status_t mazu_b200_validate_reads(const mazu_index_t* idx, const uint8_t* bases, const uint64_t* read_offsets, uint64_t n_reads,
                                       int32_t mode, uint64_t counts[5]);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU: the index is immutable, so it is REPLICATED per device and reads are SHARDED in contiguous
 * blocks; no collective on the data path, the counters of the shards are summed on the host -- the shape of
 * K2U::validate_self_parallel (src/kphf/mod.rs:105-139: rayon over independent units, one reduction at the end).
 * ------------------------------------------------------------------------------------------- */
/* Copies every device table of `src` onto each device of `devices` with device-to-device copies (NVLink peer copies where
 * the devices can reach each other; no host-side rebuild, no second upload): out[i] is an independent handle on
 * devices[i] that shares src's host-side metadata.  A device equal to src's own gets a full copy too. */
mazu_status_t mazu_b200_index_replicate(const mazu_index_t* src, const int32_t* devices, int32_t n_devices, mazu_index_t** out);
/* mazu_b200_query_reads (host buffers) over n replicas: reads are cut into n contiguous blocks of about equal size in
 * bases, block i runs on handles[i] from its own host thread with its own streams, every record lands in its slot of
 * out_hits, counts is the sum over the shards.  All handles must hold the same index (same k, unitigs, K2U). */
mazu_status_t mazu_b200_query_reads_sharded(const mazu_index_t* const* handles, int32_t n_handles, const uint8_t* bases,
                                            const uint64_t* read_offsets, uint64_t n_reads, uint64_t uniform_read_len, int32_t mode,
                                            uint64_t* kmer_offsets, mazu_hit_t* out_hits, uint64_t* counts);
/* mazu_b200_query_reads_runs over n replicas.  out_runs is cut into n regions of cap_runs / n records, shard i fills region
 * i; out_read_run_offsets[r] (n_reads entries are meaningful) is the index in out_runs of read r's first run, which is all
 * mazu_b200_expand_hit_runs needs -- the runs of different shards are NOT adjacent.  *out_n_runs = total number of runs;
 * MAZU_ERR_INVALID_ARG if a shard's region is too small (then *out_n_runs = n x the largest shard's need). */
mazu_status_t mazu_b200_query_reads_runs_sharded(const mazu_index_t* const* handles, int32_t n_handles, const uint8_t* bases,
                                                 const uint64_t* read_offsets, uint64_t n_reads, uint64_t uniform_read_len, int32_t mode,
                                                 uint64_t* kmer_offsets, uint8_t* out_codes, mazu_hit_t* out_runs, uint64_t cap_runs,
                                                 uint64_t* out_read_run_offsets, uint64_t* out_n_runs, uint64_t* counts);

/* ---------------------------------------------------------------------------------------------
 * Page-locked host memory for MAZU_MEM_HOST callers that do not link the CUDA runtime themselves (a Rust
 * Vec is pageable: copies from / to it are staged by the driver, synchronously, and the H2D / kernel / D2H
 * overlap of the host path is lost -- results are identical, throughput is not).  Buffers obtained here
 * are pinned and portable across devices.
 * ------------------------------------------------------------------------------------------- */
mazu_status_t mazu_b200_alloc_pinned(uint64_t bytes, void** out);
void mazu_b200_free_pinned(void* p);

#ifdef __cplusplus
}
#endif
#endif /* MAZU_B200_H */
