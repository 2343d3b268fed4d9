//! Safe wrapper over `libmazu_b200.so`.
//!
//! `GpuIndex` owns the opaque device index.  Its methods mirror mazu's trait surface for the query path:
//!   * `K2U::{k, unitig_len, n_unitigs, n_kmers, sum_unitigs_len, k2u}`        (mazu src/kphf/mod.rs:58-67)
//!   * `U2Pos::{decode_unitig_occs}` / `GetRefPos::project_hits`, batched         (mazu src/index.rs:133-216,349-361)
//!   * the read loop of `kphf bench` / `validate_ckmers`, random or `.as_streaming()` (mazu src/bin/kphf/main.rs:299-322)
//!   * `Validate::validate_self`, `K2U::validate_self`                             (mazu src/index/validate.rs:24-52, src/kphf/mod.rs:69-139)
//! NOT compiled in the repository that ships it (no Rust toolchain in that image); see ../README.md.
use mazu_b200_sys as sys;
use std::ffi::{CStr, CString};
use std::path::Path;
use std::ptr;

/// `kmers::MatchType` as used by `K2UPos.o`.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum MatchType {
    NoMatch,
    IdentityMatch,
    TwinMatch,
}
/// `mazu::kphf::K2UPos` (src/kphf/mod.rs:13-19).
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub struct K2UPos {
    pub unitig_id: usize,
    pub unitig_len: usize,
    pub pos: usize,
    pub o: MatchType,
}
/// `mazu::index::MappedRefPos` (src/index.rs:25-31); `fw` = Orientation::Forward.
pub type MappedRefPos = sys::mazu_occ_t;
pub type Hit = sys::mazu_hit_t;

/// `mazu::Error` (src/err.rs) + the CUDA-side failures.
#[derive(Debug)]
pub enum Error {
    IO(String),
    InvalidData(String),
    EFNotMonotone,
    EFEmpty,
    Cuda(String),
    InvalidArg(String),
    NoU2Pos,
    NoRefseq,
    Other(String),
}
pub type Result<T> = std::result::Result<T, Error>;

fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::mazu_b200_last_error()).to_string_lossy().into_owned() }
}
fn check(rc: i32) -> Result<()> {
    match rc {
        sys::MAZU_OK => Ok(()),
        sys::MAZU_ERR_IO => Err(Error::IO(last_error())),
        sys::MAZU_ERR_INVALID_DATA => Err(Error::InvalidData(last_error())),
        sys::MAZU_ERR_EF_NOT_MONOTONE => Err(Error::EFNotMonotone),
        sys::MAZU_ERR_EF_EMPTY => Err(Error::EFEmpty),
        sys::MAZU_ERR_CUDA => Err(Error::Cuda(last_error())),
        // the reference panics on a query of the wrong k (src/index.rs:157-163, src/kphf/sshash.rs:473, src/kphf/pfhash.rs:109)
        sys::MAZU_ERR_K_MISMATCH => panic!("{}", last_error()),
        sys::MAZU_ERR_INVALID_ARG => Err(Error::InvalidArg(last_error())),
        sys::MAZU_ERR_NO_U2POS => Err(Error::NoU2Pos),
        sys::MAZU_ERR_NO_REFSEQ => Err(Error::NoRefseq),
        _ => Err(Error::Other(last_error())),
    }
}
fn hit_to_k2upos(h: &Hit) -> Option<K2UPos> {
    let o = match h.r#match {
        sys::MAZU_IDENTITY_MATCH => MatchType::IdentityMatch,
        sys::MAZU_TWIN_MATCH => MatchType::TwinMatch,
        _ => return None, // NoMatch, or a window CanonicalKmerIterator skips
    };
    Some(K2UPos { unitig_id: h.unitig_id as usize, unitig_len: h.unitig_len as usize, pos: h.pos as usize, o })
}

/// Page-locked host buffer (`mazu_b200_alloc_pinned`).  Results and reads that live in one of these move over PCIe
/// asynchronously and overlapped with the kernels; a plain `Vec` is pageable and measured 3x slower through the same call.
pub struct PinnedBuf<T: Copy> {
    ptr: *mut T,
    len: usize,
}
unsafe impl<T: Copy + Send> Send for PinnedBuf<T> {}
impl<T: Copy> PinnedBuf<T> {
    pub fn new(len: usize, fill: T) -> Result<Self> {
        let mut p: *mut std::os::raw::c_void = ptr::null_mut();
        check(unsafe { sys::mazu_b200_alloc_pinned((len * std::mem::size_of::<T>()) as u64, &mut p) })?;
        let ptr = p as *mut T;
        for i in 0..len {
            unsafe { ptr.add(i).write(fill) };
        }
        Ok(Self { ptr, len })
    }
    pub fn as_slice(&self) -> &[T] {
        unsafe { std::slice::from_raw_parts(self.ptr, self.len) }
    }
    pub fn as_mut_slice(&mut self) -> &mut [T] {
        unsafe { std::slice::from_raw_parts_mut(self.ptr, self.len) }
    }
}
impl<T: Copy> Drop for PinnedBuf<T> {
    fn drop(&mut self) {
        unsafe { sys::mazu_b200_free_pinned(self.ptr as *mut std::os::raw::c_void) }
    }
}

/// A device-resident index.  Immutable after construction: `&self` queries may run concurrently from many threads
/// (the reference's queries take `&self` and are `Sync`, src/kphf/mod.rs:69-72).
pub struct GpuIndex {
    raw: *mut sys::mazu_index_t,
}
unsafe impl Send for GpuIndex {}
unsafe impl Sync for GpuIndex {}
impl Drop for GpuIndex {
    fn drop(&mut self) {
        unsafe { sys::mazu_b200_index_destroy(self.raw) }
    }
}

/// `UnitigSet` (src/unitig_set.rs:31-36) as plain arrays: 2-bit packed sequence words, base count, prefix lengths.
pub struct UnitigSetParts<'a> {
    pub k: usize,
    pub useq_words: &'a [u64],
    pub n_bases: usize,
    pub accum_lens: &'a [u64],
}
impl<'a> UnitigSetParts<'a> {
    fn desc(&self) -> sys::mazu_unitig_set_desc_t {
        sys::mazu_unitig_set_desc_t {
            k: self.k as u32,
            useq_words: self.useq_words.as_ptr(),
            n_bases: self.n_bases as u64,
            accum_lens: self.accum_lens.as_ptr(),
            n_unitigs: (self.accum_lens.len() - 1) as u64,
        }
    }
}

impl GpuIndex {
    /// `DenseIndex::deserialize_from_cpp(dir)` (src/pf1/dense_index.rs:33-97).
    pub fn dense_index_from_cpp<P: AsRef<Path>>(dir: P, device: i32) -> Result<Self> {
        let c = CString::new(dir.as_ref().to_string_lossy().as_bytes()).map_err(|e| Error::InvalidArg(e.to_string()))?;
        let mut raw = ptr::null_mut();
        check(unsafe { sys::mazu_b200_dense_index_deserialize_from_cpp(c.as_ptr(), device, &mut raw) })?;
        Ok(Self { raw })
    }
    /// `SSHash::from_unitig_set(unitigs, w, skew_param, WyHashState::with_seed(seed))` (src/kphf/sshash.rs:405-412);
    /// `skew_param = None` is `from_unitig_set_no_skew_index` (:397-403).  `on_device` builds every table on the GPU.
    pub fn sshash_from_unitig_set(u: &UnitigSetParts, w: usize, skew_param: Option<usize>, seed: u64, device: i32, on_device: bool) -> Result<Self> {
        let d = u.desc();
        let skew = skew_param.map(|s| s as u64).unwrap_or(sys::MAZU_SKEW_NONE);
        let mut raw = ptr::null_mut();
        let rc = unsafe {
            if on_device {
                sys::mazu_b200_index_create_sshash_gpu(&d, w as u32, skew, seed, device, &mut raw)
            } else {
                sys::mazu_b200_index_create_sshash(&d, w as u32, skew, seed, device, &mut raw)
            }
        };
        check(rc)?;
        Ok(Self { raw })
    }
    /// `PFHash::from_unitig_set(unitigs)` (src/kphf/pfhash.rs:40-73).
    pub fn pfhash_from_unitig_set(u: &UnitigSetParts, device: i32) -> Result<Self> {
        let d = u.desc();
        let mut raw = ptr::null_mut();
        check(unsafe { sys::mazu_b200_index_create_pfhash(&d, device, &mut raw) })?;
        Ok(Self { raw })
    }
    /// The `U2Pos` half of `ModIndex::from_parts` for a `PiscemUnitigTable` (src/index/dense_unitig_table.rs:109-118):
    /// pass the words of the two `IntVector`s.
    pub fn attach_piscem_table(&mut self, ctable_words: &[u64], ctable_width: usize, n_occs: usize, ref_shift: usize, pos_mask: u64,
                               offset_words: &[u64], offset_width: usize, n_offsets: usize) -> Result<()> {
        let ct = sys::mazu_packed_vec_desc_t { words: ctable_words.as_ptr(), width: ctable_width as u64, len: n_occs as u64 };
        let co = sys::mazu_packed_vec_desc_t { words: offset_words.as_ptr(), width: offset_width as u64, len: n_offsets as u64 };
        check(unsafe { sys::mazu_b200_index_attach_u2pos_piscem(self.raw, &ct, ref_shift as u64, pos_mask, &co) })
    }
    /// `RefSeqCollection::from_parts(seq, prefix_sum)` (src/refseq.rs:124-126) -- needed by `validate_self`.
    pub fn attach_refseq(&mut self, seq_words: &[u64], prefix_sum: &[u64]) -> Result<()> {
        check(unsafe { sys::mazu_b200_index_attach_refseq(self.raw, seq_words.as_ptr(), prefix_sum.as_ptr(), (prefix_sum.len() - 1) as u64) })
    }

    fn info(&self, what: i32) -> usize {
        unsafe { sys::mazu_b200_index_info(self.raw, what) as usize }
    }
    // ---- K2U accessors (src/kphf/mod.rs:58-67)
    pub fn k(&self) -> usize { self.info(sys::MAZU_INFO_K) }
    pub fn n_unitigs(&self) -> usize { self.info(sys::MAZU_INFO_N_UNITIGS) }
    pub fn n_kmers(&self) -> usize { self.info(sys::MAZU_INFO_N_KMERS) }
    pub fn sum_unitigs_len(&self) -> usize { self.info(sys::MAZU_INFO_SUM_UNITIGS_LEN) }
    pub fn unitig_len(&self, id: usize) -> usize {
        let (mut len, mut start) = (0u64, 0u64);
        check(unsafe { sys::mazu_b200_unitig_len(self.raw, id as u64, &mut len, &mut start) }).expect("unitig id out of range");
        len as usize
    }

    /// Batched `K2U::k2u`: `fw_words[i]` = `CanonicalKmer::get_fw_mer().into_u64()` of query i.
    pub fn k2u_batch(&self, fw_words: &[u64]) -> Result<Vec<Option<K2UPos>>> {
        let mut out = vec![Hit { unitig_id: !0, unitig_len: !0, pos: !0, r#match: 0 }; fw_words.len()];
        check(unsafe {
            sys::mazu_b200_k2u_batch(self.raw, fw_words.as_ptr(), fw_words.len() as u64, self.k() as u32, out.as_mut_ptr(), sys::MAZU_MEM_HOST,
                                     ptr::null_mut())
        })?;
        Ok(out.iter().map(hit_to_k2upos).collect())
    }
    /// `K2U::k2u(&self, km)` -- a batch of one; use `k2u_batch` / `query_reads` for throughput.
    pub fn k2u_word(&self, fw_word: u64) -> Option<K2UPos> {
        self.k2u_batch(&[fw_word]).expect("k2u").pop().unwrap()
    }

    /// The read loop of `kphf bench` / `validate_ckmers`: every k-mer of every read, random access (`streaming = false`) or
    /// `.as_streaming()` semantics with the cursor reset per read.  `read_offsets` has `n_reads + 1` byte offsets into `bases`.
    /// Returns one record per k-mer slot (`kmer_offsets[r] + position in read`), the slot offsets, and {n_kmers, n_hit, n_miss}.
    pub fn query_reads(&self, bases: &[u8], read_offsets: &[u64], streaming: bool) -> Result<(Vec<Hit>, Vec<u64>, [u64; 3])> {
        let n_reads = (read_offsets.len() - 1) as u64;
        let n_slots = unsafe { sys::mazu_b200_count_kmer_slots(self.raw, read_offsets.as_ptr(), n_reads, 0) } as usize;
        let mut hits = vec![Hit { unitig_id: !0, unitig_len: !0, pos: !0, r#match: 0 }; n_slots];
        let mut koffs = vec![0u64; read_offsets.len()];
        let mut counts = [0u64; 3];
        let mode = if streaming { sys::MAZU_MODE_STREAMING } else { sys::MAZU_MODE_RANDOM };
        check(unsafe {
            sys::mazu_b200_query_reads(self.raw, bases.as_ptr(), read_offsets.as_ptr(), n_reads, 0, mode, koffs.as_mut_ptr(), hits.as_mut_ptr(),
                                       counts.as_mut_ptr(), sys::MAZU_MEM_HOST, ptr::null_mut())
        })?;
        Ok((hits, koffs, counts))
    }

    /// `query_reads` into caller-owned pinned buffers (reuse them across batches): `hits` must hold `count_kmer_slots` records.
    pub fn query_reads_into(&self, bases: &PinnedBuf<u8>, read_offsets: &[u64], streaming: bool, hits: &mut PinnedBuf<Hit>) -> Result<[u64; 3]> {
        let n_reads = (read_offsets.len() - 1) as u64;
        let n_slots = unsafe { sys::mazu_b200_count_kmer_slots(self.raw, read_offsets.as_ptr(), n_reads, 0) } as usize;
        if hits.len < n_slots {
            return Err(Error::InvalidArg(format!("hits buffer holds {} records, need {}", hits.len, n_slots)));
        }
        let mut counts = [0u64; 3];
        let mode = if streaming { sys::MAZU_MODE_STREAMING } else { sys::MAZU_MODE_RANDOM };
        check(unsafe {
            sys::mazu_b200_query_reads(self.raw, bases.ptr, read_offsets.as_ptr(), n_reads, 0, mode, ptr::null_mut(), hits.ptr, counts.as_mut_ptr(),
                                       sys::MAZU_MEM_HOST, ptr::null_mut())
        })?;
        Ok(counts)
    }

    /// `query_reads` with the result as hit RUNS (one code byte per k-mer slot + the record of every run start + per-read run
    /// offsets; ~1.2 B per lookup over PCIe instead of 16).  Pinned buffers take the library's synchronisation-free path.
    /// Returns the number of runs written; `expand_hit_runs` rebuilds the exact per-slot records.
    pub fn query_reads_runs(&self, bases: &PinnedBuf<u8>, read_offsets: &[u64], streaming: bool, codes: &mut PinnedBuf<u8>,
                            runs: &mut PinnedBuf<Hit>, read_run_offsets: &mut [u64], kmer_offsets: &mut [u64]) -> Result<(usize, [u64; 3])> {
        let n_reads = (read_offsets.len() - 1) as u64;
        let mut counts = [0u64; 3];
        let mut n_runs = 0u64;
        let mode = if streaming { sys::MAZU_MODE_STREAMING } else { sys::MAZU_MODE_RANDOM };
        check(unsafe {
            sys::mazu_b200_query_reads_runs(self.raw, bases.ptr, read_offsets.as_ptr(), n_reads, 0, mode, kmer_offsets.as_mut_ptr(), codes.ptr, runs.ptr,
                                            runs.len as u64, read_run_offsets.as_mut_ptr(), &mut n_runs, counts.as_mut_ptr())
        })?;
        Ok((n_runs as usize, counts))
    }
    /// Host-side decoder of the run format (no device work).
    pub fn expand_hit_runs(codes: &[u8], runs: &[Hit], read_run_offsets: &[u64], kmer_offsets: &[u64], out: &mut [Hit]) -> Result<()> {
        check(unsafe {
            sys::mazu_b200_expand_hit_runs(codes.as_ptr(), runs.as_ptr(), read_run_offsets.as_ptr(), kmer_offsets.as_ptr(),
                                           (read_run_offsets.len() - 1) as u64, 0, out.as_mut_ptr())
        })
    }

    /// The sparsest lossless host result of a short-read batch: one 16-byte `HitInterval` per run of consecutive k-mer slots whose
    /// hits walk along one unitig, nothing per slot or per read (~0.19 B per lookup over PCIe).  `packed_reads`: 2-bit words in the
    /// `kmers::SeqVector` layout, `ceil(read_len / 32)` words per read; `n_mask`: one bit per base, set = not ACGT.  Returns the
    /// number of intervals written; `expand_hit_intervals` rebuilds the per-slot records.
    pub fn query_reads_intervals(&self, packed_reads: &PinnedBuf<u64>, n_mask: Option<&[u64]>, n_reads: usize, read_len: usize, streaming: bool,
                                 out: &mut PinnedBuf<sys::mazu_hit_interval_t>) -> Result<(usize, [u64; 3])> {
        let mut counts = [0u64; 3];
        let mut n = 0u64;
        let mode = if streaming { sys::MAZU_MODE_STREAMING } else { sys::MAZU_MODE_RANDOM };
        check(unsafe {
            sys::mazu_b200_query_reads_intervals_packed(self.raw, packed_reads.ptr, n_mask.map_or(std::ptr::null(), |m| m.as_ptr()), n_reads as u64,
                                                        read_len as u64, mode, out.ptr, out.len as u64, &mut n, counts.as_mut_ptr())
        })?;
        Ok((n as usize, counts))
    }
    /// The same for ASCII reads (`n_reads * read_len` bytes, any case, non-ACGT skipped) -- the slices the reference hands to
    /// `CanonicalKmerIterator::from_u8_slice` -- with nothing for the caller to pack.
    pub fn query_reads_intervals_ascii(&self, bases: &PinnedBuf<u8>, n_reads: usize, read_len: usize, streaming: bool,
                                       out: &mut PinnedBuf<sys::mazu_hit_interval_t>) -> Result<(usize, [u64; 3])> {
        let mut counts = [0u64; 3];
        let mut n = 0u64;
        let mode = if streaming { sys::MAZU_MODE_STREAMING } else { sys::MAZU_MODE_RANDOM };
        check(unsafe {
            sys::mazu_b200_query_reads_intervals(self.raw, bases.ptr, n_reads as u64, read_len as u64, mode, out.ptr, out.len as u64, &mut n,
                                                 counts.as_mut_ptr())
        })?;
        Ok((n as usize, counts))
    }
    /// Decoder for the ASCII call: skipped windows are read off the bases.
    pub fn expand_hit_intervals_ascii(&self, intervals: &[sys::mazu_hit_interval_t], bases: &[u8], n_reads: usize, read_len: usize,
                                      out: &mut [Hit]) -> Result<()> {
        check(unsafe {
            sys::mazu_b200_expand_hit_intervals_ascii(self.raw, intervals.as_ptr(), intervals.len() as u64, bases.as_ptr(), n_reads as u64,
                                                      read_len as u64, out.as_mut_ptr())
        })
    }
    /// Host-side decoder of the interval format: `out[r * (read_len - k + 1) + slot]`.
    pub fn expand_hit_intervals(&self, intervals: &[sys::mazu_hit_interval_t], n_mask: Option<&[u64]>, n_reads: usize, read_len: usize,
                                out: &mut [Hit]) -> Result<()> {
        check(unsafe {
            sys::mazu_b200_expand_hit_intervals(self.raw, intervals.as_ptr(), intervals.len() as u64, n_mask.map_or(std::ptr::null(), |m| m.as_ptr()),
                                                n_reads as u64, read_len as u64, out.as_mut_ptr())
        })
    }

    /// Batched `GetRefPos::project_hits` (src/index.rs:156-216): occurrences of hit i are `out[offsets[i]..offsets[i+1]]`.
    pub fn project_hits(&self, hits: &[Hit]) -> Result<(Vec<u64>, Vec<MappedRefPos>)> {
        self.occ_call(Some(hits), None)
    }
    /// Batched `U2Pos::decode_unitig_occs` (src/index/dense_unitig_table.rs:55-76,127-153).
    pub fn decode_occs(&self, unitig_ids: &[u32]) -> Result<(Vec<u64>, Vec<sys::mazu_occ_t>)> {
        self.occ_call(None, Some(unitig_ids))
    }
    fn occ_call(&self, hits: Option<&[Hit]>, uids: Option<&[u32]>) -> Result<(Vec<u64>, Vec<sys::mazu_occ_t>)> {
        let n = hits.map(|h| h.len()).or(uids.map(|u| u.len())).unwrap();
        let mut offs = vec![0u64; n + 1];
        let mut total = 0u64;
        let call = |out: *mut sys::mazu_occ_t, cap: u64, offs: &mut Vec<u64>, total: &mut u64| unsafe {
            match (hits, uids) {
                (Some(h), _) => sys::mazu_b200_project_hits(self.raw, h.as_ptr(), n as u64, offs.as_mut_ptr(), out, cap, total, sys::MAZU_MEM_HOST, ptr::null_mut()),
                (_, Some(u)) => sys::mazu_b200_decode_occs(self.raw, u.as_ptr(), n as u64, offs.as_mut_ptr(), out, cap, total, sys::MAZU_MEM_HOST, ptr::null_mut()),
                _ => unreachable!(),
            }
        };
        check(call(ptr::null_mut(), 0, &mut offs, &mut total))?; // sizes first
        let mut out = vec![sys::mazu_occ_t { ref_id: 0, pos: 0, fw: 0 }; total as usize];
        if total > 0 {
            check(call(out.as_mut_ptr(), total, &mut offs, &mut total))?;
        }
        Ok((offs, out))
    }

    /// `Validate::validate_self` (src/index/validate.rs:24-52) on the device: {n_queries, n_identity, n_twin, n_projected, n_fail}.
    pub fn validate_self(&self) -> Result<[u64; 5]> {
        let mut c = [0u64; 5];
        check(unsafe { sys::mazu_b200_validate_self(self.raw, c.as_mut_ptr()) })?;
        Ok(c)
    }
    /// `K2U::validate_self` (src/kphf/mod.rs:69-139): every unitig k-mer in both orientations.
    pub fn k2u_validate_self(&self) -> Result<[u64; 5]> {
        let mut c = [0u64; 5];
        check(unsafe { sys::mazu_b200_k2u_validate_self(self.raw, c.as_mut_ptr()) })?;
        Ok(c)
    }
    /// `Validate::validate_fasta` (src/index/validate.rs:83-100) / `StreamingIndex::validate_fasta` (src/index/caching.rs:204-218):
    /// the library reads the FASTA / FASTQ file; record i is reference i.  {n_queries, n_identity, n_twin, n_projected, n_fail}.
    pub fn validate_fasta<P: AsRef<Path>>(&self, path: P, streaming: bool) -> Result<[u64; 5]> {
        let c_path = CString::new(path.as_ref().to_string_lossy().as_bytes()).map_err(|e| Error::InvalidArg(e.to_string()))?;
        let mut c = [0u64; 5];
        check(unsafe { sys::mazu_b200_validate_fasta(self.raw, c_path.as_ptr(), streaming as i32, c.as_mut_ptr()) })?;
        Ok(c)
    }
    /// `GetRefPos::get_ref_pos` + `project_hits` over a read batch in one call (src/index.rs:156-216): (slot offsets, positions).
    pub fn get_ref_pos_reads(&self, bases: &[u8], read_offsets: &[u64], streaming: bool) -> Result<(Vec<u64>, Vec<MappedRefPos>)> {
        let n_reads = (read_offsets.len() - 1) as u64;
        let n_slots = unsafe { sys::mazu_b200_count_kmer_slots(self.raw, read_offsets.as_ptr(), n_reads, 0) };
        let mut offs = vec![0u64; n_slots as usize + 1];
        let mut cap = 2 * n_slots + 1024;
        loop {
            let mut out = vec![MappedRefPos { ref_id: 0, pos: 0, fw: 0 }; cap as usize];
            let mut total = 0u64;
            let rc = unsafe {
                sys::mazu_b200_get_ref_pos_reads(self.raw, bases.as_ptr(), read_offsets.as_ptr(), n_reads, 0, streaming as i32, n_slots,
                                                 std::ptr::null_mut(), std::ptr::null_mut(), offs.as_mut_ptr(), out.as_mut_ptr() as *mut sys::mazu_occ_t,
                                                 cap, &mut total, std::ptr::null_mut(), 0 /* MAZU_MEM_HOST */, std::ptr::null_mut())
            };
            if rc != 0 && total > cap {
                cap = total;
                continue;
            }
            check(rc)?;
            out.truncate(total as usize);
            return Ok((offs, out));
        }
    }
    /// `K2U::unitig_seq` (src/kphf/mod.rs:64): the unitig's 2-bit words (base j at bits [2j, 2j+2)) and its length.
    pub fn unitig_seq_words(&self, id: usize) -> Result<(Vec<u64>, usize)> {
        let mut len = 0u64;
        check(unsafe { sys::mazu_b200_unitig_seq(self.raw, id as u64, std::ptr::null_mut(), 0, &mut len) })?;
        let mut w = vec![0u64; ((len + 31) / 32) as usize];
        check(unsafe { sys::mazu_b200_unitig_seq(self.raw, id as u64, w.as_mut_ptr(), w.len() as u64, &mut len) })?;
        Ok((w, len as usize))
    }
    /// One copy of the index per device, made with device-to-device copies (the index is replicated, reads are sharded).
    pub fn replicate(&self, devices: &[i32]) -> Result<Vec<GpuIndex>> {
        let mut raw = vec![std::ptr::null_mut(); devices.len()];
        check(unsafe { sys::mazu_b200_index_replicate(self.raw, devices.as_ptr(), devices.len() as i32, raw.as_mut_ptr()) })?;
        Ok(raw.into_iter().map(|r| GpuIndex { raw: r }).collect())
    }
}

/// Replicas of one index on several devices: `query_reads` shards the reads over them in contiguous blocks
/// (one host thread + streams per device inside the library) and sums the counters -- the shape of
/// `K2U::validate_self_parallel` (src/kphf/mod.rs:105-139).
pub struct GpuIndexSet {
    pub replicas: Vec<GpuIndex>,
}
impl GpuIndexSet {
    pub fn query_reads(&self, bases: &[u8], read_offsets: &[u64], streaming: bool) -> Result<(Vec<Hit>, [u64; 3])> {
        let handles: Vec<*const sys::mazu_index_t> = self.replicas.iter().map(|r| r.raw as *const _).collect();
        let n_reads = (read_offsets.len() - 1) as u64;
        let n_slots = unsafe { sys::mazu_b200_count_kmer_slots(handles[0], read_offsets.as_ptr(), n_reads, 0) };
        let mut hits = vec![Hit { unitig_id: !0, unitig_len: !0, pos: !0, r#match: 0 }; n_slots as usize];
        let mut counts = [0u64; 3];
        check(unsafe {
            sys::mazu_b200_query_reads_sharded(handles.as_ptr(), handles.len() as i32, bases.as_ptr(), read_offsets.as_ptr(), n_reads, 0, streaming as i32,
                                               std::ptr::null_mut(), hits.as_mut_ptr() as *mut sys::mazu_hit_t, counts.as_mut_ptr())
        })?;
        Ok((hits, counts))
    }
}

// With `mazu` as a dependency the trait impl is a thin shim (UnitigSet kept next to the handle for unitig_seq / AsRef):
//
// pub struct GpuK2U { idx: GpuIndex, unitigs: mazu::unitig_set::UnitigSet }
// impl mazu::kphf::K2U for GpuK2U {
//     fn k(&self) -> usize { self.idx.k() }
//     fn unitig_len(&self, id: usize) -> usize { self.unitigs.unitig_len(id) }
//     fn n_unitigs(&self) -> usize { self.unitigs.n_unitigs() }
//     fn unitig_seq(&self, id: usize) -> SeqVectorSlice { self.unitigs.unitig_seq(id) }
//     fn n_kmers(&self) -> usize { self.unitigs.n_kmers() }
//     fn sum_unitigs_len(&self) -> usize { self.unitigs.total_len() }
//     fn k2u(&self, km: &CanonicalKmer) -> Option<mazu::kphf::K2UPos> {
//         assert_eq!(km.len(), self.k());
//         self.idx.k2u_word(km.get_fw_mer().into_u64()).map(|p| mazu::kphf::K2UPos { unitig_id: p.unitig_id, unitig_len: p.unitig_len, pos: p.pos,
//             o: if p.o == MatchType::IdentityMatch { kmers::naive_impl::MatchType::IdentityMatch } else { kmers::naive_impl::MatchType::TwinMatch } })
//     }
// }
// impl AsRef<mazu::unitig_set::UnitigSet> for GpuK2U { fn as_ref(&self) -> &mazu::unitig_set::UnitigSet { &self.unitigs } }
//
// `ModIndex::from_parts(base, GpuK2U, u2pos, refs)` then works unchanged: GetRefPos, Validate and StreamingK2U are generic over K2U.
