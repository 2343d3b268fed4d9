//! Raw bindings: one declaration per symbol of `include/mazu_b200.h` (same order).  See that header for the
//! reference file:line each entry point replaces.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

pub type mazu_status_t = i32;
pub const MAZU_OK: i32 = 0;
pub const MAZU_ERR_IO: i32 = -1;
pub const MAZU_ERR_INVALID_DATA: i32 = -2;
pub const MAZU_ERR_EF_NOT_MONOTONE: i32 = -3;
pub const MAZU_ERR_EF_EMPTY: i32 = -4;
pub const MAZU_ERR_CUDA: i32 = -5;
pub const MAZU_ERR_K_MISMATCH: i32 = -6;
pub const MAZU_ERR_INVALID_ARG: i32 = -7;
pub const MAZU_ERR_NO_U2POS: i32 = -8;
pub const MAZU_ERR_NO_REFSEQ: i32 = -9;
pub const MAZU_ERR_OTHER: i32 = -10;

pub const MAZU_NO_MATCH: u32 = 0;
pub const MAZU_IDENTITY_MATCH: u32 = 1;
pub const MAZU_TWIN_MATCH: u32 = 2;
pub const MAZU_SKIPPED: u32 = 3;

pub const MAZU_MEM_HOST: i32 = 0;
pub const MAZU_MEM_DEVICE: i32 = 1;
pub const MAZU_MEM_HOST_IN_DEVICE_OUT: i32 = 2;
pub const MAZU_MODE_RANDOM: i32 = 0;
pub const MAZU_MODE_STREAMING: i32 = 1;
pub const MAZU_K2U_PFHASH: i32 = 0;
pub const MAZU_K2U_SSHASH: i32 = 1;
pub const MAZU_K2U_SAMPLED_PFHASH: i32 = 2;
pub const MAZU_INDEX_PUFFERFISH_DENSE: i32 = 0;
pub const MAZU_INDEX_PISCEM: i32 = 1;
pub const MAZU_SKEW_NONE: u64 = u64::MAX;

pub const MAZU_INFO_K: i32 = 0;
pub const MAZU_INFO_N_UNITIGS: i32 = 1;
pub const MAZU_INFO_N_KMERS: i32 = 2;
pub const MAZU_INFO_SUM_UNITIGS_LEN: i32 = 3;
pub const MAZU_INFO_N_MINIMIZERS: i32 = 4;
pub const MAZU_INFO_N_KMERS_IN_SKEW_INDEX: i32 = 5;
pub const MAZU_INFO_N_REFS: i32 = 6;
pub const MAZU_INFO_N_TOTAL_OCCS: i32 = 7;

#[repr(C)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub struct mazu_hit_t {
    pub unitig_id: u32,
    pub unitig_len: u32,
    pub pos: u32,
    pub r#match: u32,
}
#[repr(C)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub struct mazu_hit8_t {
    pub unitig_id: u32,
    pub pos_match: u32,
}
/// One run of consecutive k-mer slots of one read whose hits walk along one unitig (`mazu_b200_query_reads_intervals_packed`).
#[repr(C)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub struct mazu_hit_interval_t {
    pub unitig_id: u32,
    /// bits 0-30: `K2UPos::pos` of the run's first k-mer; bit 31: 1 = `MatchType::TwinMatch`
    pub pos_o: u32,
    pub read: u32,
    pub start: u16,
    pub len: u16,
}
#[repr(C)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub struct mazu_occ_t {
    pub ref_id: u32,
    pub pos: u32,
    pub fw: u32,
}
#[repr(C)]
pub struct mazu_index_t {
    _private: [u8; 0],
}
#[repr(C)]
pub struct mazu_fasta_t {
    _private: [u8; 0],
}
#[repr(C)]
pub struct mazu_unitig_set_desc_t {
    pub k: u32,
    pub useq_words: *const u64,
    pub n_bases: u64,
    pub accum_lens: *const u64,
    pub n_unitigs: u64,
}
#[repr(C)]
pub struct mazu_packed_vec_desc_t {
    pub words: *const u64,
    pub width: u64,
    pub len: u64,
}
#[repr(C)]
pub struct mazu_boophf_desc_t {
    pub n_levels: u32,
    pub level_words: *const *const u64,
    pub level_n_bits: *const u64,
    pub last_bitset_rank: u64,
    pub n_elem: u64,
    pub final_keys: *const u64,
    pub final_vals: *const u64,
    pub n_final: u64,
}

extern "C" {
    pub fn mazu_b200_last_error() -> *const c_char;
    pub fn mazu_b200_device_count() -> i32;
    pub fn mazu_b200_dense_index_deserialize_from_cpp(dir: *const c_char, device: i32, out: *mut *mut mazu_index_t) -> mazu_status_t;
    pub fn mazu_b200_sparse_index_deserialize_from_cpp(dir: *const c_char, device: i32, out: *mut *mut mazu_index_t) -> mazu_status_t;
    pub fn mazu_b200_index_from_cf_prefix(prefix: *const c_char, index_kind: i32, w: u32, skew_param: u64, hash_seed: u64, device: i32,
                                          out: *mut *mut mazu_index_t) -> mazu_status_t;
    pub fn mazu_b200_index_create_sshash(unitigs: *const mazu_unitig_set_desc_t, w: u32, skew_param: u64, hash_seed: u64, device: i32,
                                         out: *mut *mut mazu_index_t) -> mazu_status_t;
    pub fn mazu_b200_index_create_sshash_gpu(unitigs: *const mazu_unitig_set_desc_t, w: u32, skew_param: u64, hash_seed: u64, device: i32,
                                             out: *mut *mut mazu_index_t) -> mazu_status_t;
    pub fn mazu_b200_index_create_pfhash_gpu(unitigs: *const mazu_unitig_set_desc_t, device: i32, out: *mut *mut mazu_index_t) -> mazu_status_t;
    pub fn mazu_b200_debug_table_digest(idx: *const mazu_index_t, which: i32, digest: *mut u64, n_bytes: *mut u64) -> mazu_status_t;
    pub fn mazu_b200_debug_probe_key(idx: *const mazu_index_t, fw_words: *const u64, n: u64, out_block: *mut u32, stream: *mut c_void) -> mazu_status_t;
    pub fn mazu_b200_index_create_pfhash(unitigs: *const mazu_unitig_set_desc_t, device: i32, out: *mut *mut mazu_index_t) -> mazu_status_t;
    pub fn mazu_b200_index_create_pfhash_from_parts(unitigs: *const mazu_unitig_set_desc_t, mphf: *const mazu_boophf_desc_t,
                                                    pos: *const mazu_packed_vec_desc_t, device: i32, out: *mut *mut mazu_index_t) -> mazu_status_t;
    pub fn mazu_b200_index_rebuild_k2u(src: *const mazu_index_t, k2u_kind: i32, w: u32, skew_param: u64, hash_seed: u64,
                                       out: *mut *mut mazu_index_t) -> mazu_status_t;
    pub fn mazu_b200_index_attach_u2pos_dense(idx: *mut mazu_index_t, ctable: *const u64, n_occs: u64,
                                              contig_offsets: *const mazu_packed_vec_desc_t) -> mazu_status_t;
    pub fn mazu_b200_index_attach_u2pos_piscem(idx: *mut mazu_index_t, ctable: *const mazu_packed_vec_desc_t, ref_shift: u64, pos_mask: u64,
                                               contig_offsets: *const mazu_packed_vec_desc_t) -> mazu_status_t;
    pub fn mazu_b200_index_attach_refseq(idx: *mut mazu_index_t, seq_words: *const u64, prefix_sum: *const u64, n_refs: u64) -> mazu_status_t;
    pub fn mazu_b200_index_destroy(idx: *mut mazu_index_t);
    pub fn mazu_b200_index_release_scratch(idx: *mut mazu_index_t, released: *mut u64) -> mazu_status_t;
    pub fn mazu_b200_index_info(idx: *const mazu_index_t, what: i32) -> u64;
    pub fn mazu_b200_unitig_len(idx: *const mazu_index_t, unitig_id: u64, len: *mut u64, start_pos: *mut u64) -> mazu_status_t;
    pub fn mazu_b200_k2u_batch(idx: *const mazu_index_t, fw_words: *const u64, n: u64, k: u32, out_hits: *mut mazu_hit_t, mem: i32,
                               stream: *mut c_void) -> mazu_status_t;
    pub fn mazu_b200_query_reads(idx: *const mazu_index_t, bases: *const u8, read_offsets: *const u64, n_reads: u64, uniform_read_len: u64,
                                 mode: i32, kmer_offsets: *mut u64, out_hits: *mut mazu_hit_t, counts: *mut u64, mem: i32,
                                 stream: *mut c_void) -> mazu_status_t;
    pub fn mazu_b200_query_reads_compact(idx: *const mazu_index_t, bases: *const u8, read_offsets: *const u64, n_reads: u64,
                                         uniform_read_len: u64, mode: i32, kmer_offsets: *mut u64, out_hits: *mut mazu_hit8_t,
                                         counts: *mut u64, mem: i32, stream: *mut c_void) -> mazu_status_t;
    pub fn mazu_b200_query_reads_runs(idx: *const mazu_index_t, bases: *const u8, read_offsets: *const u64, n_reads: u64, uniform_read_len: u64,
                                      mode: i32, kmer_offsets: *mut u64, out_codes: *mut u8, out_runs: *mut mazu_hit_t, cap_runs: u64,
                                      out_read_run_offsets: *mut u64, out_n_runs: *mut u64, counts: *mut u64) -> mazu_status_t;
    pub fn mazu_b200_expand_hit_runs(codes: *const u8, runs: *const mazu_hit_t, read_run_offsets: *const u64, kmer_offsets: *const u64,
                                     n_reads: u64, uniform_slots: u64, out_hits: *mut mazu_hit_t) -> mazu_status_t;
    pub fn mazu_b200_count_kmer_slots(idx: *const mazu_index_t, read_offsets: *const u64, n_reads: u64, uniform_read_len: u64) -> u64;
    pub fn mazu_b200_encode_reads(idx: *const mazu_index_t, bases: *const u8, read_offsets: *const u64, n_reads: u64, uniform_read_len: u64,
                                  kmer_offsets: *const u64, out_fw: *mut u64, out_rc: *mut u64, out_mm_word: *mut u64,
                                  out_mm_offset: *mut u32, out_valid: *mut u8, stream: *mut c_void) -> mazu_status_t;
    pub fn mazu_b200_decode_occs(idx: *const mazu_index_t, unitig_ids: *const u32, n: u64, out_offsets: *mut u64, out_occs: *mut mazu_occ_t,
                                 cap: u64, out_total: *mut u64, mem: i32, stream: *mut c_void) -> mazu_status_t;
    pub fn mazu_b200_project_hits(idx: *const mazu_index_t, hits: *const mazu_hit_t, n: u64, out_offsets: *mut u64, out_mrps: *mut mazu_occ_t,
                                  cap: u64, out_total: *mut u64, mem: i32, stream: *mut c_void) -> mazu_status_t;
    pub fn mazu_b200_iter_unitigs_on_ref(idx: *const mazu_index_t, ref_id: u64, out: *mut mazu_hit_t, cap: u64, n_out: *mut u64) -> mazu_status_t;
    pub fn mazu_b200_validate_self(idx: *const mazu_index_t, counts: *mut u64) -> mazu_status_t;
    pub fn mazu_b200_k2u_validate_self(idx: *const mazu_index_t, counts: *mut u64) -> mazu_status_t;
    pub fn mazu_b200_alloc_pinned(bytes: u64, out: *mut *mut c_void) -> mazu_status_t;
    pub fn mazu_b200_free_pinned(p: *mut c_void);
    pub fn mazu_b200_get_ref_pos_reads(idx: *const mazu_index_t, bases: *const u8, read_offsets: *const u64, n_reads: u64, uniform_read_len: u64, mode: i32, n_slots: u64, kmer_offsets: *mut u64, out_hits: *mut mazu_hit_t, out_offsets: *mut u64, out_mrps: *mut mazu_occ_t, cap: u64, out_total: *mut u64, counts: *mut u64, mem: i32, stream: *mut c_void) -> mazu_status_t;
    pub fn mazu_b200_query_reads_runs_packed(idx: *const mazu_index_t, packed_reads: *const u64, n_mask: *const u64, n_reads: u64, read_len: u64, mode: i32, out_codes2: *mut u8, out_runs: *mut mazu_hit_t, cap_runs: u64, out_read_run_offsets: *mut u64, out_n_runs: *mut u64, counts: *mut u64) -> mazu_status_t;
    pub fn mazu_b200_query_reads_intervals_packed(idx: *const mazu_index_t, packed_reads: *const u64, n_mask: *const u64, n_reads: u64, read_len: u64, mode: i32, out_intervals: *mut mazu_hit_interval_t, cap: u64, out_n: *mut u64, counts: *mut u64) -> mazu_status_t;
    pub fn mazu_b200_expand_hit_intervals(idx: *const mazu_index_t, intervals: *const mazu_hit_interval_t, n_intervals: u64, n_mask: *const u64, n_reads: u64, read_len: u64, out_hits: *mut mazu_hit_t) -> mazu_status_t;
    pub fn mazu_b200_query_reads_intervals(idx: *const mazu_index_t, bases: *const u8, n_reads: u64, read_len: u64, mode: i32, out_intervals: *mut mazu_hit_interval_t, cap: u64, out_n: *mut u64, counts: *mut u64) -> mazu_status_t;
    pub fn mazu_b200_expand_hit_intervals_ascii(idx: *const mazu_index_t, intervals: *const mazu_hit_interval_t, n_intervals: u64, bases: *const u8, n_reads: u64, read_len: u64, out_hits: *mut mazu_hit_t) -> mazu_status_t;
    pub fn mazu_b200_pack_reads(bases: *const u8, n_reads: u64, read_len: u64, out_words: *mut u64, out_n_mask: *mut u64, n_non_acgt: *mut u64) -> mazu_status_t;
    pub fn mazu_b200_expand_hit_runs_packed(codes2: *const u8, runs: *const mazu_hit_t, read_run_offsets: *const u64, n_reads: u64, uniform_slots: u64, out_hits: *mut mazu_hit_t) -> mazu_status_t;
    pub fn mazu_b200_unitig_seq(idx: *const mazu_index_t, unitig_id: u64, out_words: *mut u64, cap_words: u64, len: *mut u64) -> mazu_status_t;
    pub fn mazu_b200_fasta_open(path: *const c_char, out: *mut *mut mazu_fasta_t) -> mazu_status_t;
    pub fn mazu_b200_fasta_close(f: *mut mazu_fasta_t);
    pub fn mazu_b200_fasta_n_records(f: *const mazu_fasta_t) -> u64;
    pub fn mazu_b200_fasta_bases(f: *const mazu_fasta_t) -> *const u8;
    pub fn mazu_b200_fasta_offsets(f: *const mazu_fasta_t) -> *const u64;
    pub fn mazu_b200_fasta_name(f: *const mazu_fasta_t, record: u64) -> *const c_char;
    pub fn mazu_b200_validate_fasta(idx: *const mazu_index_t, path: *const c_char, mode: i32, counts: *mut u64) -> mazu_status_t;
    pub fn mazu_b200_validate_reads(idx: *const mazu_index_t, bases: *const u8, read_offsets: *const u64, n_reads: u64, mode: i32, counts: *mut u64) -> mazu_status_t;
    pub fn mazu_b200_index_replicate(src: *const mazu_index_t, devices: *const i32, n_devices: i32, out: *mut *mut mazu_index_t) -> mazu_status_t;
    pub fn mazu_b200_query_reads_sharded(handles: *const *const mazu_index_t, n_handles: i32, bases: *const u8, read_offsets: *const u64, n_reads: u64, uniform_read_len: u64, mode: i32, kmer_offsets: *mut u64, out_hits: *mut mazu_hit_t, counts: *mut u64) -> mazu_status_t;
    pub fn mazu_b200_query_reads_runs_sharded(handles: *const *const mazu_index_t, n_handles: i32, bases: *const u8, read_offsets: *const u64, n_reads: u64, uniform_read_len: u64, mode: i32, kmer_offsets: *mut u64, out_codes: *mut u8, out_runs: *mut mazu_hit_t, cap_runs: u64, out_read_run_offsets: *mut u64, out_n_runs: *mut u64, counts: *mut u64) -> mazu_status_t;
    pub fn mazu_b200_debug_gather_probe(table_bytes: u64, n_items: u64, granule_bytes: i32, ilp: i32, blocks_per_sm: i32, iters: i32, device: i32, items_per_s: *mut f64) -> mazu_status_t;
}
