// CUDA kernels of the batched k-mer query path (sm_100a).
//
//  K1  encode / roll canonical k-mers / minimizers            (warp per read; ballot + redux + smem)
//  K2  random-access lookup: MPHF -> bucket bounds -> positions -> k-mer verify -> unitig bounds
//  K3  streaming walk (.as_streaming()), one warp per read, speculate-then-commit
//  K4  occurrence decode (+ projection)
//  plus the GPU validate drivers and the random-gather roofline probe.
//
// None of this is GEMM-shaped: no tensor cores.  The bound is L2/HBM sector traffic and
// dependent-load latency, so the kernels are organised around (a) one aligned 32-byte block per
// dependent step (index_layout.hpp) and (b) keeping consecutive k-mers of a read in adjacent
// lanes, which makes lanes that share a minimizer issue identical addresses (the LSU merges them).
#pragma once
#include <cuda_runtime.h>

#include <cuda/ptx>

#include "index_layout.hpp"

namespace mazu {

struct __align__(16) Hit {
  u32 unitig_id, unitig_len, pos, match;
};
struct OccRec {
  u32 ref_id, pos, fw;
};

__device__ __forceinline__ Hit hit_none(u32 match) { return Hit{~0u, ~0u, ~0u, match}; }
// fire-and-forget fetch of the line holding *p into L2: used where a later dependent load's address is already known
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void store_hit(Hit* out, const Hit& h) {
  // evict-first store: the result stream (16 B per lookup, 19 GB per step) is never read back by the kernel and should not push
  // index lines out of the L2 (+1.6 % on config 5, +0.6 % on config 2 against a plain store)
  __stcs(reinterpret_cast<uint4*>(out), make_uint4(h.unitig_id, h.unitig_len, h.pos, h.match));
}

// record store: 16-byte mazu_hit_t, or the 8-byte mazu_hit8_t {unitig_id, pos | match << 30} of the compact calls
__device__ __forceinline__ void store_rec(void* base, u64 idx, const Hit& h, bool compact) {
  if (compact) reinterpret_cast<uint2*>(base)[idx] = make_uint2(h.unitig_id, (h.pos & 0x3FFFFFFFu) | (h.match << 30));
  else store_hit(reinterpret_cast<Hit*>(base) + idx, h);
}

// Unitig lines (index_layout.hpp) from the flat arrays, one thread per line.  n_words = readable words of useq.
__global__ void __launch_bounds__(256) build_unitig_lines_kernel(const __grid_constant__ UnitigsView u, u64 n_lines, u64 n_words,
                                                                 UnitigLine* __restrict__ lines) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_lines; i += (u64)gridDim.x * blockDim.x) {
    UnitigLine L;
    const u64 base = i << ULINE_SHIFT;
#pragma unroll
    for (int j = 0; j < 9; ++j) L.seq[j] = 8 * i + j < n_words ? u.useq[8 * i + j] : 0ULL;
    L.ends[0] = L.ends[1] = L.ends[2] = L.ends[3] = 0ULL;
    L.first_id = (u32)(u.n_unitigs ? u.n_unitigs - 1 : 0);
    L.start_delta = L.end_delta = 0;
    if (base < u.total_len) {
      u64 id, start, end;
      unitig_locate(u, base, id, start, end);
      L.first_id = (u32)id;
      L.start_delta = (u32)(base - start);
      while (end <= base + 256) {  // unitigs ending inside the line: their last base is base + (end - 1 - base)
        const u32 j = (u32)(end - 1 - base);
        L.ends[j >> 6] |= 1ULL << (j & 63);
        if (++id >= u.n_unitigs) break;
        end = u.starts[id + 1];
      }
      L.end_delta = (u32)(end - base);  // only read when no unitig ends at or after the queried base inside this line
    }
    u32 cnt = 0, prev = 255u;
    L.cnt_before = L.prev_end = L.next_end = 0;
#pragma unroll
    for (u32 w = 0; w < 4; ++w) {
      L.cnt_before |= cnt << (8 * w);
      L.prev_end |= prev << (8 * w);
      if (L.ends[w]) prev = 64 * w + 63 - (u32)__clzll((long long)L.ends[w]);
      cnt += (u32)__popcll(L.ends[w]);
    }
    u32 next = 0;
#pragma unroll
    for (int w = 3; w >= 0; --w) {
      L.next_end |= next << (8 * w);
      if (L.ends[w]) next = 64 * w + (u32)__ffsll((long long)L.ends[w]) - 1;
    }
    lines[i] = L;
  }
}

// kmers::CanonicalKmer::get_word_equivalency
__device__ __forceinline__ u32 word_equivalency(u64 fw, u64 rc, u64 kw) {
  return kw == fw ? (u32)IDENTITY_MATCH : (kw == rc ? (u32)TWIN_MATCH : (u32)NO_MATCH);
}

// K2UPos from a verified useq position: pos_to_id, unitig_len, unitig_start_pos (+ the boundary
// guard of src/kphf/sshash.rs:513-514,539-541 when `guard` is set)
__device__ __forceinline__ bool finish_hit(const UnitigsView& u, u64 km_pos, u32 mt, bool guard, Hit& out, u64* ustart = nullptr, u32* dup = nullptr) {
  u64 id, start, end;
  line_locate(u, km_pos, id, start, end, dup);
  if (guard && km_pos + u.k > end) return false;
  if (ustart) *ustart = start;
  out.unitig_id = (u32)id;
  out.unitig_len = (u32)(end - start);
  out.pos = (u32)(km_pos - start);
  out.match = mt;
  return true;
}

// PFHash::k2u (src/kphf/pfhash.rs:108-134)
__device__ __forceinline__ bool pfhash_k2u(const IndexView& ix, u64 fw, u64 rc, Hit& out) {
  u64 word = fw <= rc ? fw : rc;
  u64 h;
  if (!mphf_lookup(ix.mphf, word, h)) return false;
  if (h >= ix.pos.len) return false;
  u64 km_pos = packed_get(ix.pos, h);
  u32 mt = word_equivalency(fw, rc, line_window(ix.unitigs, km_pos));
  if (mt == NO_MATCH) return false;
  return finish_hit(ix.unitigs, km_pos, mt, false, out);
}

// SSHash::k2u (src/kphf/sshash.rs:471-555), second half: from the minimizer's slot h (bucket bounds -> skew index or bucket
// entries -> candidate windows -> unitig id / bounds).  `offset` = minimizer offset in fw-mer coordinates.
__device__ __forceinline__ bool sshash_finish(const IndexView& ix, u64 fw, u64 rc, u64 h, u32 offset, Hit& out) {
  if (h + 1 >= ix.sizes.n) return false;
  u64 pos_start, pos_end;
  blocked_ef_get2(ix.sizes, h, pos_start, pos_end);  // occs_prefix_sum.get(h), get(h+1)
  if (pos_end - pos_start > ix.skew_param) {         // k2u_skew_index (sshash.rs:415-433)
    if (!ix.has_skew) return false;
    u64 word = fw <= rc ? fw : rc, hs;
    if (!mphf_lookup(ix.skew_mphf, word, hs)) return false;
    if (hs >= ix.skew_pos.len) return false;
    u64 p = packed_get(ix.skew_pos, hs);
    u32 mt = word_equivalency(fw, rc, line_window(ix.unitigs, p));
    if (mt == NO_MATCH) return false;
    return finish_hit(ix.unitigs, p, mt, false, out);
  }
  const u32 k = ix.unitigs.k;
  const u64 last_km_start_pos = ix.unitigs.total_len - k;
  const u64 rc_offset = (u64)(k - offset - ix.w);
  for (u64 pi = pos_start; pi < pos_end; ++pi) {  // (the builders drop an entry equal to its predecessor: host_build.hpp step 2b)
    u64 mm_pos = packed_get(ix.pos, pi);
    if (mm_pos >= offset && mm_pos - offset <= last_km_start_pos) {  // sshash.rs:498
      u64 km_pos = mm_pos - offset;
      u32 mt = word_equivalency(fw, rc, line_window(ix.unitigs, km_pos));
      if (mt != NO_MATCH && finish_hit(ix.unitigs, km_pos, mt, true, out)) return true;
    }
    if (mm_pos >= rc_offset && mm_pos - rc_offset <= last_km_start_pos) {  // sshash.rs:527
      u64 km_pos = mm_pos - rc_offset;
      u32 mt = word_equivalency(fw, rc, line_window(ix.unitigs, km_pos));
      if (mt != NO_MATCH && finish_hit(ix.unitigs, km_pos, mt, true, out)) return true;
    }
  }
  return false;
}
// SSHash::k2u given the canonical minimizer (word, offset in fw-mer coordinates)
__device__ __forceinline__ bool sshash_k2u(const IndexView& ix, u64 fw, u64 rc, u64 mm_word, u32 offset, Hit& out) {
  u64 h;
  if (!cascade_lookup(ix.mphf, ix.sizes, mm_word, h)) return false;  // fingerprinted cascade: slot + membership filter in one (index_layout.hpp)
  return sshash_finish(ix, fw, rc, h, offset, out);
}

template <u32 FAMILY>
__device__ __forceinline__ bool sampled_pfhash_k2u_t(const IndexView& ix, u64 fw, u64 rc, Hit& out, u64* ustart, u32* dup = nullptr);
__device__ __forceinline__ bool k2u_any(const IndexView& ix, u64 fw, u64 rc, Hit& out) {
  if (ix.k2u_kind == MAZU_K2U_PFHASH) return pfhash_k2u(ix, fw, rc, out);
  if (ix.k2u_kind == MAZU_K2U_SAMPLED_PFHASH) return sampled_pfhash_k2u_t<MPHF_FAMILY_BOOPHF>(ix, fw, rc, out, nullptr);
  MinimizerResult m = canonical_minimizer_naive(fw, rc, ix.unitigs.k, ix.w, ix.seed);
  return sshash_k2u(ix, fw, rc, m.word, m.offset, out);
}

// ---------------------------------------------------------------------------------------------
// K2 on a flat batch of k-mer words: K2U::k2u per element (random order, no read context)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k2u_batch_kernel(const __grid_constant__ IndexView ix, const u64* __restrict__ fw_words, u64 n,
                                                        Hit* __restrict__ out) {
  const u32 k = ix.unitigs.k;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    u64 fw = fw_words[i] & kmer_mask(k);
    u64 rc = revcomp(fw, k);
    Hit h;
    if (!k2u_any(ix, fw, rc, h)) h = hit_none(NO_MATCH);
    store_hit(out + i, h);
  }
}

// K2U::k2u over a flat batch on a PFHash index (src/kphf/pfhash.rs:108-134; configs[0]: the pufferfish DenseIndex).
// The MPHF level loop is the divergent part of such a lookup -- one key per lane keeps 12 of 32 lanes busy
// (profiles/r01k_prof_config1.summary.txt) -- so a lane takes KB_N keys and walks them through ONE loop, refilling itself
// (mphf_levels_multi); the hashes before the loop and rank / position / window / unitig bounds after it run with all lanes.
#ifndef MAZU_KB_N
#define MAZU_KB_N 4
#endif
static const int KB_N = MAZU_KB_N;
template <u32 FAMILY>
__global__ void __launch_bounds__(256) k2u_batch_pfhash_kernel(const __grid_constant__ IndexView ix, const u64* __restrict__ fw_words, u64 n,
                                                               Hit* __restrict__ out) {
  __shared__ u64 sa[KB_N * 256];
  __shared__ u64 sb[FAMILY == MPHF_FAMILY_NATIVE ? 1 : KB_N * 256];
  const u32 k = ix.unitigs.k, tid = threadIdx.x;
  const u64 kmask = kmer_mask(k), tile = (u64)KB_N * 256;
  for (u64 base = (u64)blockIdx.x * tile; base < n; base += (u64)gridDim.x * tile) {
    u64 fw[KB_N];
    u32 n_mine = 0;
#pragma unroll
    for (int j = 0; j < KB_N; ++j) {
      const u64 i = base + (u64)j * 256 + tid;
      fw[j] = 0;
      if (i < n) {
        fw[j] = fw_words[i] & kmask;
        const u64 rc = revcomp(fw[j], k), key = fw[j] <= rc ? fw[j] : rc;
        if (FAMILY == MPHF_FAMILY_NATIVE) {
          sa[j * 256 + tid] = fmix64(key);
        } else {
          sa[j * 256 + tid] = boophf_hash64(key, BOOPHF_SEED0);
          sb[j * 256 + tid] = boophf_hash64(key, BOOPHF_SEED1);
        }
        n_mine = j + 1;
      }
    }
    mphf_levels_multi<FAMILY>(ix.mphf, sa + tid, sb + (FAMILY == MPHF_FAMILY_NATIVE ? 0 : tid), 256, n_mine);
#pragma unroll
    for (int j = 0; j < KB_N; ++j) {
      const u64 i = base + (u64)j * 256 + tid;
      if (i < n) {
        const u64 rc = revcomp(fw[j], k);
        Hit h = hit_none(NO_MATCH);
        u64 hv;
        if (mphf_multi_rank(ix.mphf, sa[j * 256 + tid], fw[j] <= rc ? fw[j] : rc, hv) && hv < ix.pos.len) {
          const u64 km_pos = packed_get(ix.pos, hv);
          const u32 mt = word_equivalency(fw[j], rc, line_window(ix.unitigs, km_pos));
          if (mt == NO_MATCH || !finish_hit(ix.unitigs, km_pos, mt, false, h)) h = hit_none(NO_MATCH);
        }
        store_hit(out + i, h);
      }
    }
  }
}

// measurement hook: level-0 MPHF block of every query's key
__global__ void probe_key_kernel(const __grid_constant__ IndexView ix, const u64* __restrict__ fw_words, u64 n, u32* __restrict__ out_block) {
  const u32 k = ix.unitigs.k;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    const u64 fw = fw_words[i] & kmer_mask(k), rc = revcomp(fw, k);
    u64 key = fw <= rc ? fw : rc;
    if (ix.k2u_kind == MAZU_K2U_SSHASH) key = canonical_minimizer_naive(fw, rc, k, ix.w, ix.seed).word;
    u64 blk = 0;
    u32 bit = 0;
    if (ix.mphf.family == MPHF_FAMILY_CASCADE) blk = cascade_slot(fmix64(key), 0, ix.mphf.size[0]) >> 5;
    else if (ix.mphf.family == MPHF_FAMILY_NATIVE) native_slot(key, 0, ix.mphf.size[0], blk, bit);
    else blk = mulhi64(boophf_hash64(key, BOOPHF_SEED0), ix.mphf.size[0]) / MPHF_BLOCK_BITS;
    out_block[i] = (u32)blk;
  }
}

// ---------------------------------------------------------------------------------------------
// Read kernels (K1 / K2-on-reads / K3).  A warp owns one read and walks it in chunks of QR_CHUNK
// k-mer start positions; everything a chunk needs is staged in that warp's slice of shared memory:
//
//   stage E  encode: coalesced byte loads -> 2-bit words by warp OR-reduction (redux.sync) and
//            ballot; the chunk's k-mer words go to smem (fw[]), validity to four ballot masks
//   stage M  minimizers: one 32-bit hash per w-mer and strand (hf[], hr[]), then per k-mer the
//            smallest (hash | offset) key over its window of k-w+1 w-mers
//   stage B  buckets: k-mers of a super-k-mer share their minimizer, hence MPHF slot and bucket;
//            the FIRST k-mer of every run of equal minimizers is a leader.  Leaders of the whole
//            chunk (~15 of 128 k-mers) are compacted and resolved in one pass:
//            MPHF -> blocked Elias-Fano bounds -> (start, size) parked in smem at the leader's slot
//   stage V  verify, per k-mer: bucket positions -> candidate windows -> equality -> unitig id / bounds
//
// So the dependent MPHF/Elias-Fano chain runs once per super-k-mer, not once per k-mer, and the
// kernel holds ONE copy of each stage (the first version inlined the lookup four times; ncu showed
// 72 % of stall samples in `no_instruction`, i.e. instruction-cache misses -- profiles/r01_*).
// ---------------------------------------------------------------------------------------------
#ifndef MAZU_QR_WARPS
#define MAZU_QR_WARPS 8
#endif
static const int QR_WARPS = MAZU_QR_WARPS;  // warps per CTA (A/B knob)
static const int QR_CHUNK = 128;  // k-mer start positions per chunk
static const int QR_BASES = 160;  // bases staged per chunk (CHUNK + k - 1 <= 159)
static const u64 QR_SEGMENT = 2048;  // k-mer positions per work item when a long read is cut up (a multiple of QR_CHUNK)
static const u32 BN_SKEW = 0xFFFFFFFFu;
static const u32 QR_LOC = 64, QR_NO_LOC = 255;

struct WarpStage {
  u64 fw[QR_CHUNK];      // forward k-mer word per chunk position (garbage where invalid)
  u64 rc[QR_CHUNK];      // its reverse complement, written by stage M (SSHash only) so stages B and V do not recompute it
  u64 bstart[QR_CHUNK];  // leaders only: index of the first bucket entry
  union {
    struct {
      u32 hf[QR_BASES];  // w-mer hash keys (stage M) ...
      u32 hr[QR_BASES];  // ... and the same keys stored REVERSED (hr[QR_BASES-1-q]) so both strands scan upwards
    };
    u64 bfirst[QR_CHUNK];  // leaders only: the first bucket entry itself (minimizer position), fetched by the leader once the
                           // keys are dead (stage B); 2 * QR_BASES * 4 >= QR_CHUNK * 8
  };
  u32 bn[QR_CHUNK];      // leaders only: bucket size, 0 = minimizer unknown, BN_SKEW = heavy bucket
  u8 off[QR_CHUNK];      // minimizer offset in fw-mer coordinates
  u8 leader[QR_CHUNK];   // chunk position of this k-mer's leader
  u8 lead_list[QR_CHUNK];
  // the unitig around the first bucket entry, located by the leader (first QR_LOC leaders of a chunk; a 150 bp read has ~18):
  u8 lrank[QR_CHUNK];    // leaders only: slot in the three arrays below, QR_NO_LOC = none
  u32 uid[QR_LOC];       // unitig id
  u32 sdelta[QR_LOC];    // entry - unitig start
  u32 edelta[QR_LOC];    // unitig end - entry; bit 31: DUP flag of the lines a k-mer of the super-k-mer can start in
};
static_assert(sizeof(WarpStage) * QR_WARPS <= 48 * 1024, "WarpStage must fit the static shared-memory limit");

struct ChunkInfo {
  u32 vm[4];  // ballot masks: bit `lane` of vm[t] <=> k-mer 32t+lane is a valid window inside the read
  u32 n_c;    // k-mer positions of this chunk that lie inside the read
};

// stage E, second half: every lane holds the chunk's packed words w[] and invalid-base masks inv[]; k-mer words to shared
// memory, validity of the chunk's k-mer windows to four ballot masks
__device__ __forceinline__ void stage_encode_finish(const u64 (&w)[6], const u32 (&inv)[6], u32 n_c, u32 k, u32 lane, WarpStage& S, ChunkInfo& ci) {
  const u64 kmask = kmer_mask(k);
  const u64 vmask = (1ULL << k) - 1ULL;  // k <= 32
  const u32 sh = 2 * lane;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    u64 x = w[t] >> sh;
    if (sh) x |= w[t + 1] << (64 - sh);
    S.fw[32 * t + lane] = x & kmask;
    u64 m = ((u64)inv[t] | ((u64)inv[t + 1] << 32)) >> lane;
    bool valid = (32u * t + lane < n_c) && ((m & vmask) == 0ULL);
    ci.vm[t] = __ballot_sync(0xffffffffu, valid);
  }
  ci.n_c = n_c;
}

// stage E
__device__ __forceinline__ void stage_encode(const u8* __restrict__ seq, u64 len, u64 c0, u32 n_c, u32 k, u32 lane, WarpStage& S, ChunkInfo& ci) {
  u64 w[6];
  u32 inv[6];
#pragma unroll
  for (int t = 0; t < 5; ++t) {
    u64 q = c0 + 32u * t + lane;
    u32 code = 4;
    if (q < len) code = base_code(seq[q]);
    bool bad = code > 3;
    inv[t] = __ballot_sync(0xffffffffu, bad);
    u32 v = bad ? 0u : code;
    u32 lo = __reduce_or_sync(0xffffffffu, lane < 16 ? (v << (2 * lane)) : 0u);
    u32 hi = __reduce_or_sync(0xffffffffu, lane >= 16 ? (v << (2 * (lane - 16))) : 0u);
    w[t] = ((u64)hi << 32) | lo;
  }
  w[5] = 0;
  inv[5] = 0xffffffffu;
  stage_encode_finish(w, inv, n_c, k, lane, S, ci);
}

// stage E for a read that arrives 2-bit packed (kmers::SeqVector layout: base j at bits [2j, 2j+2) of word j / 32; optional
// N mask, one bit per base) and fits one chunk: its words ARE the packed words stage E builds, so the warp loads them with
// uniform (broadcast) loads and no ASCII copy of the read ever exists (mazu_b200_query_reads_runs_packed).
__device__ __forceinline__ void stage_encode_packed(const u64* __restrict__ words, const u64* __restrict__ n_mask, u32 len, u32 n_c, u32 k, u32 lane,
                                                    WarpStage& S, ChunkInfo& ci) {
  const u32 wpr = (len + 31) / 32, mpr = (len + 63) / 64;  // len <= QR_BASES - 1: at most 5 words, 3 mask words
  u64 w[6], nm[3];
  u32 inv[6];
#pragma unroll
  for (int j = 0; j < 3; ++j) nm[j] = (n_mask && (u32)j < mpr) ? __ldg(n_mask + j) : 0ULL;
#pragma unroll
  for (int t = 0; t < 5; ++t) {
    w[t] = (u32)t < wpr ? __ldg(words + t) : 0ULL;
    const int rem = (int)len - 32 * t;  // bases of the read in this word
    inv[t] = (u32)(nm[t >> 1] >> (32 * (t & 1))) | (rem >= 32 ? 0u : (rem <= 0 ? 0xffffffffu : 0xffffffffu << rem));
  }
  w[5] = 0;
  inv[5] = 0xffffffffu;
  stage_encode_finish(w, inv, n_c, k, lane, S, ci);
}

__device__ __forceinline__ bool chunk_valid(const ChunkInfo& ci, u32 q) {
  u32 word = q >> 5;
  u32 m = word == 0 ? ci.vm[0] : word == 1 ? ci.vm[1] : word == 2 ? ci.vm[2] : ci.vm[3];
  return (m >> (q & 31)) & 1u;
}

// canonical-orientation minimizer word of the k-mer at chunk position p, from its stored offset
__device__ __forceinline__ u64 mm_word_of(u64 fw, u64 rc, u32 off_fw, u32 k, u32 w) {
  const u64 wf = (fw >> (2 * off_fw)) & kmer_mask(w), wr = (rc >> (2 * (k - w - off_fw))) & kmer_mask(w);  // the w-mer on either strand
  return wf <= wr ? wf : wr;
}

// smallest (hash key | offset) over h[0..span]; unrolled for the spans of the named configs
template <int SPAN>
__device__ __forceinline__ u32 window_min_fixed(const u32* __restrict__ h) {
  u32 best = h[0];
#pragma unroll
  for (int c = 1; c <= SPAN; ++c) best = min(best, h[c] | (u32)c);
  return best;
}
__device__ __forceinline__ u32 window_min(const u32* __restrict__ h, u32 span) {
  switch (span) {
    case 16: return window_min_fixed<16>(h);  // k=31, w=15
    case 12: return window_min_fixed<12>(h);  // k=31, w=19
    default: {
      u32 best = h[0];
      for (u32 c = 1; c <= span; ++c) best = min(best, h[c] | c);
      return best;
    }
  }
}

// stage M + stage B (SSHash only)
template <u32 FAMILY, bool WALK = false, u32 KW = 0>  // WALK: the streaming walk also wants the DUP flags of the super-k-mer's lines
__device__ __forceinline__ void stage_buckets(const IndexView& ix, const ChunkInfo& ci, u32 lane, WarpStage& S) {
  const u32 k = kw_k<KW>(ix.unitigs.k), w = kw_w<KW>(ix.w), span = k - w;
  const u64 wmask = kmer_mask(w);
  __syncwarp();
  // reverse complement of every k-mer word of the chunk, once (stages M, B and V read it from here)
#pragma unroll
  for (int t = 0; t < 4; ++t) S.rc[32 * t + lane] = revcomp(S.fw[32 * t + lane], k);
  __syncwarp();
  // w-mer hash keys for chunk positions [0, 160), stored twice (hf upwards, hr REVERSED) so that the window scan of a
  // fw-canonical and of an rc-canonical k-mer both run upwards with `| c`: w-mer q is the low 2w bits of k-mer q (q < 128),
  // positions beyond come from the tail of k-mer 127.  Its reverse complement is the low 2w bits of
  // rc(k-mer q - span) (the top w bases of that k-mer), so no w-mer is reverse-complemented on its own.
  // Only w-mers q <= 127 + span belong to a k-mer of the chunk; the keys beyond are never read.
  // this loop and the window loop below are unrolled: +4 % on config 5 (random access and streaming walk alike), no spills at
  // 64 registers in the random-access kernels, the walk's 64-register build spills 40 bytes rolled or not (profiles/experiments)
#ifdef MAZU_STAGE_M_ROLLED  // A/B knob
  constexpr int UNROLL_KEYS = 1, UNROLL_MIN = 1;
#else
  constexpr int UNROLL_KEYS = 5, UNROLL_MIN = 4;
#endif
#pragma unroll UNROLL_KEYS
  for (int t = 0; t < 5; ++t) {
    u32 q = 32 * t + lane;
    u64 x;
    if (q < QR_CHUNK) x = S.fw[q];
    else {
      u32 d = q - (QR_CHUNK - 1);  // 1..32
      x = d < 32 ? (S.fw[QR_CHUNK - 1] >> (2 * d)) : 0ULL;
    }
    u64 wf = x & wmask;
    u64 wr = q >= span ? S.rc[min(q - span, (u32)QR_CHUNK - 1)] : (S.rc[0] >> (2 * (span - q)));
    const u32 key = mm_key(wf, wr & wmask, ix.seed);  // one strand-symmetric key per w-mer position (minimizer order v3)
    S.hf[q] = key;
    S.hr[QR_BASES - 1 - q] = key;
  }
  __syncwarp();
  // per k-mer minimizer, leader detection, leader compaction
  u32 n_lead = 0;
  u64 carry_mm = 0;
  u32 carry_valid = 0, carry_leader = 0;
#pragma unroll UNROLL_MIN  // static t also keeps ChunkInfo in registers (the rolled loop indexes ci.vm[t] through local memory)
  for (int t = 0; t < 4; ++t) {
    const u32 p = 32 * t + lane;
    const u32 vmask = ci.vm[t];
    const bool valid = (vmask >> lane) & 1u;
    u64 mmw = 0;
    if (valid) {
      u64 fw = S.fw[p], rc = S.rc[p];
      bool fw_canon = fw <= rc;
      // offset c in the canonical k-mer: fw strand position p+c, rc strand position p+span-c (= reversed index below + c)
      const u32* h = fw_canon ? S.hf + p : S.hr + (QR_BASES - 1 - p - span);
      u32 best = window_min(h, span);
      u32 bi = best & 31u;
      const u32 off_fw = fw_canon ? bi : span - bi;
      mmw = mm_word_of(fw, rc, off_fw, k, w);
      S.off[p] = (u8)off_fw;
    }
    // neighbour's minimizer (lane-1, or lane 31 of the previous pass)
    u64 prev_mm = __shfl_up_sync(0xffffffffu, mmw, 1);
    u32 prev_valid = (vmask << 1) & (1u << lane);
    if (lane == 0) {
      prev_mm = carry_mm;
      prev_valid = carry_valid;
    }
    const bool is_leader = valid && (!prev_valid || prev_mm != mmw);
    const u32 lmask = __ballot_sync(0xffffffffu, is_leader);
    if (valid) {
      u32 lm = lmask & ((2u << lane) - 1u);  // leaders at lanes <= lane
      S.leader[p] = (u8)(lm ? 32 * t + (31 - __clz(lm)) : carry_leader);
    }
    if (is_leader) S.lead_list[n_lead + __popc(lmask & ((1u << lane) - 1u))] = (u8)p;
    n_lead += __popc(lmask);
    carry_mm = __shfl_sync(0xffffffffu, mmw, 31);
    carry_valid = (vmask >> 31) & 1u;
    if (lmask) carry_leader = 32 * t + (31 - __clz(lmask));
  }
  __syncwarp();
  // resolve the leaders' buckets: MPHF -> Elias-Fano bounds (SSHash::k2u, sshash.rs:476-490)
#pragma unroll 1
  for (u32 i = 0; i < n_lead; i += 32) {
    if (i + lane < n_lead) {
      u32 p = S.lead_list[i + lane];
      u64 fw = S.fw[p], rc = S.rc[p];
      u64 mmw = mm_word_of(fw, rc, S.off[p], k, w);
      u64 h, a = 0, b = 0;
      u32 n = 0, loc = QR_NO_LOC;
      if (cascade_lookup(ix.mphf, ix.sizes, mmw, h) && h + 1 < ix.sizes.n) {  // the probe that finds the slot has fetched its bounds block
        blocked_ef_get2(ix.sizes, h, a, b);
        u64 cnt = b - a;
        n = cnt > ix.skew_param ? BN_SKEW : (u32)cnt;
        // the bucket's first entry now, once per super-k-mer: stage V then starts at the unitig line (one dependent DRAM access
        // per group of 32 k-mers instead of two)
        if (cnt && n != BN_SKEW) {
          const u64 mm_pos = packed_get(ix.pos, a);
          S.bfirst[p] = mm_pos;
          // ... and the unitig that holds it.  A k-mer verified against this entry starts at mm_pos - (offset <= k - w), so it
          // either lies inside this unitig or crosses its start (then the boundary guard of sshash.rs:513-514 rejects it):
          // stage V needs no unitig lookup of its own for the first entry, and the line it compares against is already on its way.
          if (i + lane < QR_LOC) {
            u64 id, st, en;
            u32 dup;
            line_locate(ix.unitigs, mm_pos, id, st, en, &dup);
            // DUP flag for the streaming walk (index_layout.hpp, ULINE_DUP): of the line of the entry, or of the line before it
            // when a k-mer of this super-k-mer can start there (a superset never hurts: flagged groups settle their cursors)
            if (WALK && (u32)(mm_pos & 255u) < k - w && mm_pos >= 256) dup |= __ldg(&ix.unitigs.lines[(mm_pos >> ULINE_SHIFT) - 1].end_delta) >> 31;
            S.uid[i + lane] = (u32)id;
            S.sdelta[i + lane] = (u32)(mm_pos - st);
            S.edelta[i + lane] = (u32)(en - mm_pos) | (dup << 31);  // unitigs are shorter than 2^31
            loc = i + lane;
          }
        }
      }
      S.bstart[p] = a;
      S.bn[p] = n;
      S.lrank[p] = (u8)loc;
    }
  }
  __syncwarp();
}

// stage V for one k-mer of an SSHash index: the loop of sshash.rs:494-552 / k2u_skew_index
template <u32 FAMILY, u32 KW = 0>
__device__ __forceinline__ bool verify_sshash(const IndexView& ix, const WarpStage& S, u32 p, u64 fw, u64 rc, Hit& out, u64* ustart = nullptr,
                                              u32* dup = nullptr) {
  const u32 lp = S.leader[p];
  const u32 n = S.bn[lp];
  if (n == 0) return false;
  const u32 k = kw_k<KW>(ix.unitigs.k);
  if (n == BN_SKEW) {
    if (!ix.has_skew) return false;
    u64 word = fw <= rc ? fw : rc, hs;
    if (!mphf_lookup_t<FAMILY>(ix.skew_mphf, word, hs)) return false;
    if (hs >= ix.skew_pos.len) return false;
    u64 pos = packed_get(ix.skew_pos, hs);
    u32 mt = word_equivalency(fw, rc, line_window<KW>(ix.unitigs, pos));
    if (mt == NO_MATCH) return false;
    return finish_hit(ix.unitigs, pos, mt, false, out, ustart, dup);
  }
  const u64 pos_start = S.bstart[lp];
  const u32 offset = S.off[p];
  const u32 rc_offset = (k - kw_w<KW>(ix.w)) - offset;
  const u64 last_km_start_pos = ix.unitigs.total_len - k;
  const u32 loc = S.lrank[lp];
#pragma unroll 1
  for (u32 e = 0; e < n; ++e) {  // (no entry equals its predecessor: the builders drop the copy the second stream pushes)
    u64 mm_pos = e == 0 ? S.bfirst[lp] : packed_get(ix.pos, pos_start + e);
    const bool located = e == 0 && loc != QR_NO_LOC;
#pragma unroll
    for (int c = 0; c < 2; ++c) {  // candidate from the fw offset (sshash.rs:498), then from the rc offset (sshash.rs:527; same window when equal)
      const u32 o = c == 0 ? offset : rc_offset;
      if (c == 1 && rc_offset == offset) break;
      if (mm_pos >= o && mm_pos - o <= last_km_start_pos) {
        const u64 km_pos = mm_pos - o;
        const u32 mt = word_equivalency(fw, rc, line_window<KW>(ix.unitigs, km_pos));
        if (mt != NO_MATCH) {
          if (located) {  // unitig of the entry, from the leader: inside it <=> sdelta >= o; boundary guard <=> k - o <= edelta
            const u32 sd = S.sdelta[loc], edf = S.edelta[loc], ed = edf & 0x7FFFFFFFu;
            if (sd >= o && k - o <= ed) {
              out.unitig_id = S.uid[loc];
              out.unitig_len = sd + ed;
              out.pos = sd - o;
              out.match = mt;
              if (ustart) *ustart = mm_pos - sd;
              if (dup) *dup = edf >> 31;
              return true;
            }
          } else if (finish_hit(ix.unitigs, km_pos, mt, true, out, ustart, dup)) {
            return true;
          }
        }
      }
    }
  }
  return false;
}

template <u32 FAMILY>
__device__ __forceinline__ bool pfhash_k2u_t(const IndexView& ix, u64 fw, u64 rc, Hit& out, u64* ustart = nullptr, u32* dup = nullptr) {
  u64 word = fw <= rc ? fw : rc;
  u64 h;
  if (!mphf_lookup_t<FAMILY>(ix.mphf, word, h)) return false;
  if (h >= ix.pos.len) return false;
  u64 km_pos = packed_get(ix.pos, h);
  u32 mt = word_equivalency(fw, rc, line_window(ix.unitigs, km_pos));
  if (mt == NO_MATCH) return false;
  return finish_hit(ix.unitigs, km_pos, mt, false, out, ustart, dup);
}

// SampledPFHash::k2u (src/kphf/pfhash.rs:190-285): sampled k-mers carry their position; the others walk
// <= extension_size stored bases to the nearest sampled k-mer, re-hash, and shift the sampled position back.
template <u32 FAMILY>
__device__ __forceinline__ bool sampled_pfhash_k2u_t(const IndexView& ix, u64 fw, u64 rc, Hit& out, u64* ustart, u32* dup) {
  const u32 k = ix.unitigs.k;
  u64 idx;
  if (!mphf_lookup_t<FAMILY>(ix.mphf, fw <= rc ? fw : rc, idx)) return false;
  if (idx >= ix.sampled.n_keys) return false;
  u64 blk = idx / MPHF_BLOCK_BITS;
  u32 bit = (u32)(idx - blk * MPHF_BLOCK_BITS);
  u64 pos;
  if (ranked_test(ix.sampled, 0, blk, bit)) {
    pos = packed_get(ix.pos, ranked_rank(ix.sampled, 0, blk, bit));
  } else {
    const u64 ext_pos = idx - ranked_rank(ix.sampled, 0, blk, bit);  // unsampled k-mers have one extension entry each
    const u64 ext_word = packed_get(ix.ext_bases, ext_pos);
    u64 f = fw, r = rc;
    const bool canon_bit = (__ldg(ix.canonical_bits + (ext_pos >> 6)) >> (ext_pos & 63)) & 1ULL;
    if ((!canon_bit) != (!(f <= r))) {  // km.swap() (pfhash.rs:213-215)
      u64 t = f;
      f = r;
      r = t;
    }
    const bool shift_fw = (__ldg(ix.direction_bits + (ext_pos >> 6)) >> (ext_pos & 63)) & 1ULL;
    const u32 n_ext = (u32)packed_get(ix.ext_sizes, ext_pos) + 1;  // extension_size - llimit
    const u64 mask = kmer_mask(k);
    long long shift = 0;
    for (u32 j = 0; j < n_ext; ++j) {
      u64 code = (ext_word >> (2 * (ix.extension_size - 1 - j))) & 3ULL;
      if (shift_fw) {  // append_base
        f = ((f >> 2) | (code << (2 * (k - 1)))) & mask;
        r = ((r << 2) | (3ULL - code)) & mask;
        --shift;
      } else {  // prepend_base
        f = ((f << 2) | code) & mask;
        r = (r >> 2) | ((3ULL - code) << (2 * (k - 1)));
        ++shift;
      }
    }
    if (!mphf_lookup_t<FAMILY>(ix.mphf, f <= r ? f : r, idx)) return false;
    if (idx >= ix.sampled.n_keys) return false;
    blk = idx / MPHF_BLOCK_BITS;
    bit = (u32)(idx - blk * MPHF_BLOCK_BITS);
    if (!ranked_test(ix.sampled, 0, blk, bit)) return false;
    pos = (u64)((long long)packed_get(ix.pos, ranked_rank(ix.sampled, 0, blk, bit)) + shift);
    // is_valid_useq_pos (unitig_set.rs:235-245)
    if (pos > ix.unitigs.total_len - k) return false;
    u64 id, s, e;
    line_locate(ix.unitigs, pos, id, s, e);
    if (pos + k > e) return false;
  }
  u32 mt = word_equivalency(fw, rc, line_window(ix.unitigs, pos));
  if (mt == NO_MATCH) return false;
  return finish_hit(ix.unitigs, pos, mt, false, out, ustart, dup);
}

struct StreamState {  // StreamingK2U { is_warm, prev_k2upos } (src/index/caching.rs:13-17)
  u32 warm, uid, ulen, pos, o;
  u64 ustart;
};

// One kernel for the read loop of `kphf bench` / validate_ckmers.
//   MODE 0: K2U::k2u per k-mer.  MODE 1: StreamingK2U::k2u_streaming, cursor reset per read.
//   KIND: MAZU_K2U_PFHASH / MAZU_K2U_SSHASH.  FAMILY: MPHF family of the index.
// OCC = resident CTAs per SM the instantiation is compiled for (register budget: 3 -> <= 80, 4 -> <= 64).  The streaming
// walk wants its 80 registers while the index is cache-resident (+3.5 % on the yeast configs) and the extra warps once
// lookups wait on DRAM (+13 % on a 1.7 GB index); the launcher picks by index size.  Random-access mode is built for 4
// (MAZU_QR_RANDOM_OCC; 3 CTAs / 80 registers measured 13 % slower on config 5).
// KW = MZ_KW(k, w) of the index folded into the code (SSHash instantiations for the named (k, w) pairs), 0 = read from the view.
template <int MODE, int KIND, u32 FAMILY, int OCC, u32 KW = 0>
__global__ void __launch_bounds__(QR_WARPS * 32, OCC) query_reads_kernel(const __grid_constant__ IndexView ix, const u8* __restrict__ bases,
                                                                    const u64* __restrict__ read_offsets, u64 n_reads, u64 uniform_len,
                                                                    const u64* __restrict__ kmer_offsets, void* __restrict__ out, u32 compact,
                                                                    unsigned long long* __restrict__ counts, const u64* __restrict__ seg_offsets) {
  __shared__ WarpStage s_stage[QR_WARPS];
  const u32 lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  WarpStage& S = s_stage[wib];
  const u32 k = kw_k<KW>(ix.unitigs.k);
  constexpr bool SS = KIND == MAZU_K2U_SSHASH;
  u32 n_valid = 0, n_hit = 0;
  const u32 lt_mask = (1u << lane) - 1u;
  // Work items.  Normally one per read.  Random-access lookups of one read are independent, so ragged batches that
  // contain LONG reads (a chromosome through validate_fasta, nanopore reads) are cut into segments of QR_SEGMENT k-mer
  // positions: seg_offsets[r] = first item of read r, seg_offsets[n_reads] = number of items; when that equals n_reads
  // no read is long and item == read.  The streaming walk is sequential per read and never segments.
  const u64 n_items = (MODE == 0 && seg_offsets) ? seg_offsets[n_reads] : n_reads;
  const bool segmented = n_items != n_reads;

  for (u64 item = (u64)blockIdx.x * QR_WARPS + wib; item < n_items; item += (u64)gridDim.x * QR_WARPS) {
    u64 r = item, c_begin = 0, c_end = ~0ULL;
    if (MODE == 0 && segmented) {
      u64 lo = 0, hi = n_reads;  // largest r with seg_offsets[r] <= item
      while (hi - lo > 1) {
        u64 mid = (lo + hi) >> 1;
        if (seg_offsets[mid] <= item) lo = mid; else hi = mid;
      }
      r = lo;
      c_begin = (item - seg_offsets[r]) * QR_SEGMENT;
      c_end = c_begin + QR_SEGMENT;
    }
    u64 beg, len, slot0;
    if (uniform_len) {
      beg = r * uniform_len;
      len = uniform_len;
      slot0 = uniform_len >= k ? r * (uniform_len - k + 1) : 0;
    } else {
      beg = read_offsets[r];
      len = read_offsets[r + 1] - beg;
      slot0 = kmer_offsets[r];
    }
    const u8* seq = bases + beg;
    const u64 nk = len >= k ? len - k + 1 : 0;
#ifdef MAZU_PREFETCH  // A/B knob: L2 prefetch of the warp's next read (profiles/experiments/README.md)
    if (uniform_len && !segmented) {  // the warp's next read: its bases are the first thing stage E waits for
      const u64 nxt = item + (u64)gridDim.x * QR_WARPS;
      if (nxt < n_items && lane < 2) prefetch_l2(bases + nxt * uniform_len + 128 * lane);
    }
#endif
    StreamState st;
    st.warm = 0;
    st.uid = st.ulen = st.pos = ~0u;
    st.o = NO_MATCH;
    st.ustart = 0;

    const u64 c_stop = min(nk, c_end);
#pragma unroll 1
    for (u64 c0 = c_begin; c0 < c_stop; c0 += QR_CHUNK) {
      ChunkInfo ci;
      const u32 n_c = (u32)min((u64)QR_CHUNK, c_stop - c0);
      __syncwarp();
      stage_encode(seq, len, c0, n_c, k, lane, S, ci);
      __syncwarp();
      char* o = out ? static_cast<char*>(out) + (slot0 + c0) * (compact ? 8ULL : 16ULL) : nullptr;
      if (MODE == 0) {
        if (SS) stage_buckets<FAMILY, false, KW>(ix, ci, lane, S);
#pragma unroll 1  // (two copies of stage V: -15 %, r03w)
        for (u32 p = lane; p < n_c; p += 32) {
          Hit h = hit_none(SKIPPED);
          if (chunk_valid(ci, p)) {
            u64 fw = S.fw[p], rc = SS ? S.rc[p] : revcomp(fw, k);
            bool ok = SS ? verify_sshash<FAMILY, KW>(ix, S, p, fw, rc, h)
                         : (KIND == MAZU_K2U_SAMPLED_PFHASH ? sampled_pfhash_k2u_t<FAMILY>(ix, fw, rc, h, nullptr) : pfhash_k2u_t<FAMILY>(ix, fw, rc, h));
            ++n_valid;
            if (ok) ++n_hit; else h = hit_none(NO_MATCH);
          }
          if (o) store_rec(o, p, h, compact);
        }
      } else {
        // StreamingK2U::k2u_streaming (caching.rs:65-103).  Every k-mer of the chunk first gets its COLD answer through the
        // same staged path as the random-access mode (stages M, B, V -- the dependent chain runs once per super-k-mer),
        // then the sequential cursor semantics are settled per group of 32 consecutive k-mers:
        //   hypothesis  each lane's answer is its cold answer, so the cursor a lane sees is the cold hit of the nearest
        //               earlier hit lane (or the cursor carried in).
        //   check       a lane whose cold answer is exactly "cursor + 1 on the same unitig" needs nothing: the warm test of
        //               the reference (k2u_warm, caching.rs:73-97) reads that very window and answers the same.  Any other
        //               lane with a warm cursor in range does the warm test; if it matches, the walk answers THERE although
        //               the cold lookup answered elsewhere or missed (only possible with duplicated canonical k-mers).
        //   commit      lanes up to and including the first such lane commit (its own cursor was right), the cursor becomes
        //               its warm answer and the remaining lanes of the group are re-checked against it -- their cold answers
        //               are kept, nothing is looked up twice.  A miss leaves the cursor untouched (`?` at caching.rs:100).
        // (An earlier version extended warm cursors first and looked up cold only on demand; on a GPU the cold path is
        // already amortised per super-k-mer: always-cold + verify measured +5 % on config 3 and, being small enough for the
        // 64-register build, +15 % on a 1.7 GB index.)
        if (SS) stage_buckets<FAMILY, true, KW>(ix, ci, lane, S);
#pragma unroll 1
        for (u32 g0 = 0; g0 < n_c; g0 += 32) {
          const u32 q = g0 + lane;
          const bool active = q < n_c;
          const bool valid = active && chunk_valid(ci, q);
          u64 fw = 0, rc = 0;
          Hit cold = hit_none(NO_MATCH);
          u64 cold_ustart = 0;
          u32 cold_dup = 0;
          bool cold_hit = false;
          if (valid) {
            fw = S.fw[q];
            rc = SS ? S.rc[q] : revcomp(fw, k);
            cold_hit = SS ? verify_sshash<FAMILY, KW>(ix, S, q, fw, rc, cold, &cold_ustart, &cold_dup)
                          : (KIND == MAZU_K2U_SAMPLED_PFHASH ? sampled_pfhash_k2u_t<FAMILY>(ix, fw, rc, cold, &cold_ustart, &cold_dup)
                                                             : pfhash_k2u_t<FAMILY>(ix, fw, rc, cold, &cold_ustart, &cold_dup));
            if (!cold_hit) cold = hit_none(NO_MATCH);
          }
          const u32 chm = __ballot_sync(0xffffffffu, cold_hit);
          // The walk can answer differently from the lookup only for a k-mer the unitig set holds twice, and every occurrence of
          // such a k-mer lies in a line flagged at creation (ULINE_DUP).  A group without a flagged cold hit therefore commits its
          // cold answers as they are; the cursor is simply its last hit.
          if (__ballot_sync(0xffffffffu, cold_hit && cold_dup) == 0) {
            if (active) {
              if (valid) {
                ++n_valid;
                if (cold_hit) ++n_hit;
              }
              if (o) store_rec(o, q, valid ? cold : hit_none(SKIPPED), compact);
            }
            if (chm) {
              const int last = 31 - __clz(chm);
              st.uid = __shfl_sync(0xffffffffu, cold.unitig_id, last);
              st.ulen = __shfl_sync(0xffffffffu, cold.unitig_len, last);
              st.pos = __shfl_sync(0xffffffffu, cold.pos, last);
              st.o = __shfl_sync(0xffffffffu, cold.match, last);
              st.ustart = __shfl_sync(0xffffffffu, cold_ustart, last);
              st.warm = 1;
            }
            continue;
          }
          u32 start = 0;  // first lane of the group that has not committed yet
#pragma unroll 1
          while (true) {
            const u32 ge_start = 0xffffffffu << start;
            const bool mine = active && lane >= start;
            if ((chm & ge_start) == 0 && !st.warm) {
              // no cursor reaches any remaining lane (cold on entry, no cold hit ahead): every answer is its cold miss
              if (mine) {
                if (valid) ++n_valid;
                if (o) store_rec(o, q, hit_none(valid ? (u32)NO_MATCH : (u32)SKIPPED), compact);
              }
              break;
            }
            const u32 before = chm & lt_mask & ge_start;
            const int prev = before ? 31 - __clz(before) : 0;
            u32 s_uid = __shfl_sync(0xffffffffu, cold.unitig_id, prev);
            u32 s_ulen = __shfl_sync(0xffffffffu, cold.unitig_len, prev);
            u32 s_pos = __shfl_sync(0xffffffffu, cold.pos, prev);
            u64 s_ustart = __shfl_sync(0xffffffffu, cold_ustart, prev);
            bool s_warm = before != 0;
            if (!before) {
              s_uid = st.uid;
              s_ulen = st.ulen;
              s_pos = st.pos;
              s_ustart = st.ustart;
              s_warm = st.warm;
            }
            Hit res = cold;
            bool res_hit = cold_hit;
            bool differs = false;
            if (mine && valid && s_warm && (u64)s_pos + 1 + k <= (u64)s_ulen &&
                !(cold_hit && cold.unitig_id == s_uid && cold.pos == s_pos + 1)) {
              u32 m = word_equivalency(fw, rc, line_window<KW>(ix.unitigs, s_ustart + s_pos + 1));
              if (m != NO_MATCH) {  // the walk answers here although the cold lookup answered elsewhere (or missed)
                res = Hit{s_uid, s_ulen, s_pos + 1, m};
                res_hit = true;
                differs = true;
              }
            }
            const u32 dm = __ballot_sync(0xffffffffu, differs);
            const u32 g = dm ? (u32)__ffs(dm) : 32u;  // lanes [start, g) commit (g includes the first differing lane)
            if (mine && lane < g) {
              Hit h = res;
              if (!valid) h = hit_none(SKIPPED);
              else {
                ++n_valid;
                if (res_hit) ++n_hit;
              }
              if (o) store_rec(o, q, h, compact);
            }
            const u32 gm = g >= 32 ? 0xffffffffu : ((1u << g) - 1u);
            const u32 hm = __ballot_sync(0xffffffffu, res_hit && mine) & gm;
            if (hm) {  // cursor = last committed hit (a miss leaves it untouched, caching.rs:100)
              const int last = 31 - __clz(hm);
              u64 us = differs ? s_ustart : cold_ustart;
              st.uid = __shfl_sync(0xffffffffu, res.unitig_id, last);
              st.ulen = __shfl_sync(0xffffffffu, res.unitig_len, last);
              st.pos = __shfl_sync(0xffffffffu, res.pos, last);
              st.o = __shfl_sync(0xffffffffu, res.match, last);
              st.ustart = __shfl_sync(0xffffffffu, us, last);
              st.warm = 1;
            }
            if (g >= 32) break;
            start = g;
          }
        }
      }
    }
  }
  // counters of src/bin/kphf/main.rs:282-284
  if (counts) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      n_valid += __shfl_xor_sync(0xffffffffu, n_valid, d);
      n_hit += __shfl_xor_sync(0xffffffffu, n_hit, d);
    }
    if (lane == 0 && n_valid) {
      atomicAdd(counts + 0, (unsigned long long)n_valid);
      atomicAdd(counts + 1, (unsigned long long)n_hit);
      atomicAdd(counts + 2, (unsigned long long)(n_valid - n_hit));
    }
  }
}

// K1 alone: per k-mer slot fw / rc / minimizer word / minimizer offset / valid
__global__ void __launch_bounds__(QR_WARPS * 32) encode_reads_kernel(const __grid_constant__ IndexView ix, const u8* __restrict__ bases,
                                                                     const u64* __restrict__ read_offsets, u64 n_reads, u64 uniform_len,
                                                                     const u64* __restrict__ kmer_offsets, u64* __restrict__ out_fw,
                                                                     u64* __restrict__ out_rc, u64* __restrict__ out_mm, u32* __restrict__ out_off,
                                                                     u8* __restrict__ out_valid) {
  __shared__ WarpStage s_stage[QR_WARPS];
  const u32 lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  WarpStage& S = s_stage[wib];
  const u32 k = ix.unitigs.k, w = ix.w;
  const bool ss = ix.k2u_kind == MAZU_K2U_SSHASH;
  for (u64 r = (u64)blockIdx.x * QR_WARPS + wib; r < n_reads; r += (u64)gridDim.x * QR_WARPS) {
    u64 beg, len, slot0;
    if (uniform_len) {
      beg = r * uniform_len;
      len = uniform_len;
      slot0 = uniform_len >= k ? r * (uniform_len - k + 1) : 0;
    } else {
      beg = read_offsets[r];
      len = read_offsets[r + 1] - beg;
      slot0 = kmer_offsets[r];
    }
    const u8* seq = bases + beg;
    const u64 nk = len >= k ? len - k + 1 : 0;
    for (u64 c0 = 0; c0 < nk; c0 += QR_CHUNK) {
      ChunkInfo ci;
      const u32 n_c = (u32)min((u64)QR_CHUNK, nk - c0);
      __syncwarp();
      stage_encode(seq, len, c0, n_c, k, lane, S, ci);
      __syncwarp();
      if (ss) stage_buckets<MPHF_FAMILY_NATIVE>(ix, ci, lane, S);
      for (u32 p = lane; p < n_c; p += 32) {
        u64 fw = 0, rc = 0, mmw = 0;
        u32 off = 0;
        bool valid = chunk_valid(ci, p);
        if (valid) {
          fw = S.fw[p];
          rc = revcomp(fw, k);
          if (ss) {
            off = S.off[p];
            mmw = mm_word_of(fw, rc, off, k, w);
          }
        }
        u64 s = slot0 + c0 + p;
        if (out_fw) out_fw[s] = fw;
        if (out_rc) out_rc[s] = rc;
        if (out_mm) out_mm[s] = mmw;
        if (out_off) out_off[s] = off;
        if (out_valid) out_valid[s] = valid ? 1 : 0;
      }
    }
  }
}

// per-read k-mer slot counts (input of the exclusive scan that yields kmer_offsets)
// per-read work items of the random-access kernel: ceil(k-mer slots / QR_SEGMENT), at least one
__global__ void segment_counts_kernel(const u64* __restrict__ read_offsets, u64 n_reads, u32 k, u64 segment, u64* __restrict__ counts) {
  for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += (u64)gridDim.x * blockDim.x) {
    u64 len = read_offsets[r + 1] - read_offsets[r];
    u64 nk = len >= k ? len - k + 1 : 0;
    counts[r] = nk <= segment ? 1 : (nk + segment - 1) / segment;
  }
}
__global__ void kmer_counts_kernel(const u64* __restrict__ read_offsets, u64 n_reads, u32 k, u64* __restrict__ counts) {
  for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += (u64)gridDim.x * blockDim.x) {
    u64 len = read_offsets[r + 1] - read_offsets[r];
    counts[r] = len >= k ? len - k + 1 : 0;
  }
}

// ---------------------------------------------------------------------------------------------
// Hit runs: a lossless, compact form of a read batch's hit records for PCIe-bound callers.
// Consecutive k-mers of a read that hit walk along one unitig, so their records are redundant: slot s continues the
// run of slot s-1 when both are hits on the same unitig with the same match type and pos differs by +1 (Identity: the
// read runs along the unitig) or -1 (Twin: against it).  Output: one code byte per k-mer slot
//   0 miss | 1 hit, continues the previous slot's run | 2 hit, starts a run (its 16-byte record is in `runs`) | 3 skipped window
// plus the run records in slot order and read_run_offsets[r] = index of read r's first run.  Typical volume: ~1.1 bytes
// per lookup instead of 16.  mazu_b200_expand_hit_runs() rebuilds the exact records.
// ---------------------------------------------------------------------------------------------
enum : u32 { RUN_MISS = 0, RUN_CONT = 1, RUN_START = 2, RUN_SKIPPED = 3 };
__device__ __forceinline__ u32 run_code_of(const Hit& h, const Hit& prev, bool has_prev) {
  if (h.match == SKIPPED) return RUN_SKIPPED;
  if (h.match != IDENTITY_MATCH && h.match != TWIN_MATCH) return RUN_MISS;
  const bool cont = has_prev && prev.match == h.match && prev.unitig_id == h.unitig_id &&
                    h.pos == (h.match == IDENTITY_MATCH ? prev.pos + 1u : prev.pos - 1u);
  return cont ? RUN_CONT : RUN_START;
}
// pass 1: codes for every slot, run count per read (a warp per read)
__global__ void __launch_bounds__(256) hit_run_codes_kernel(const Hit* __restrict__ hits, const u64* __restrict__ kmer_offsets, u64 n_reads,
                                                            u64 uniform_slots, u8* __restrict__ codes, u64* __restrict__ run_counts) {
  const u32 lane = threadIdx.x & 31;
  const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((u64)gridDim.x * blockDim.x) >> 5;
  for (u64 r = warp; r < n_reads; r += n_warps) {
    const u64 s0 = kmer_offsets ? kmer_offsets[r] : r * uniform_slots;
    const u64 ns = kmer_offsets ? kmer_offsets[r + 1] - s0 : uniform_slots;
    Hit carry = hit_none(NO_MATCH);
    u32 n_runs = 0;
    for (u64 b = 0; b < ns; b += 32) {
      const u64 i = b + lane;
      Hit h = hit_none(NO_MATCH);
      if (i < ns) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(hits + s0 + i));
        h = Hit{v.x, v.y, v.z, v.w};
      }
      Hit p;
      p.unitig_id = __shfl_up_sync(0xffffffffu, h.unitig_id, 1);
      p.unitig_len = 0;
      p.pos = __shfl_up_sync(0xffffffffu, h.pos, 1);
      p.match = __shfl_up_sync(0xffffffffu, h.match, 1);
      if (lane == 0) p = carry;
      const u32 c = run_code_of(h, p, b + lane > 0);
      if (i < ns) codes[s0 + i] = (u8)c;
      n_runs += __popc(__ballot_sync(0xffffffffu, i < ns && c == RUN_START));
      carry.unitig_id = __shfl_sync(0xffffffffu, h.unitig_id, 31);
      carry.pos = __shfl_sync(0xffffffffu, h.pos, 31);
      carry.match = __shfl_sync(0xffffffffu, h.match, 31);
    }
    if (lane == 0) run_counts[r] = n_runs;
  }
}
// pass 2: the records of the run starts, in slot order, at read_run_offsets[r] + rank
__global__ void __launch_bounds__(256) hit_run_fill_kernel(const Hit* __restrict__ hits, const u8* __restrict__ codes,
                                                           const u64* __restrict__ kmer_offsets, u64 n_reads, u64 uniform_slots,
                                                           const u64* __restrict__ read_run_offsets, u64 run_base, u64 cap, Hit* __restrict__ runs) {
  const u32 lane = threadIdx.x & 31;
  const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((u64)gridDim.x * blockDim.x) >> 5;
  for (u64 r = warp; r < n_reads; r += n_warps) {
    const u64 s0 = kmer_offsets ? kmer_offsets[r] : r * uniform_slots;
    const u64 ns = kmer_offsets ? kmer_offsets[r + 1] - s0 : uniform_slots;
    u64 o = read_run_offsets[r] - run_base;
    for (u64 b = 0; b < ns; b += 32) {
      const u64 i = b + lane;
      const bool start = i < ns && codes[s0 + i] == RUN_START;
      const u32 m = __ballot_sync(0xffffffffu, start);
      if (start) {
        const u64 dst = o + __popc(m & ((1u << lane) - 1u));
        if (dst < cap) store_hit(runs + dst, hits[s0 + i]);
      }
      o += __popc(m);
    }
  }
}

// sync-free variant of pass 2 for run buffers the device can address (pinned host memory): records go straight to
// runs[*base + offset], the chunk's offsets are globalised in place, and one thread advances *base afterwards
__global__ void __launch_bounds__(256) hit_run_fill_global_kernel(const Hit* __restrict__ hits, const u8* __restrict__ codes,
                                                                  const u64* __restrict__ kmer_offsets, u64 n_reads, u64 uniform_slots,
                                                                  u64* __restrict__ read_run_offsets, const u64* __restrict__ base_ptr, u64 cap,
                                                                  Hit* __restrict__ runs) {
  const u32 lane = threadIdx.x & 31;
  const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((u64)gridDim.x * blockDim.x) >> 5;
  const u64 base = *base_ptr;
  for (u64 r = warp; r < n_reads; r += n_warps) {
    const u64 s0 = kmer_offsets ? kmer_offsets[r] : r * uniform_slots;
    const u64 ns = kmer_offsets ? kmer_offsets[r + 1] - s0 : uniform_slots;
    u64 o = read_run_offsets[r] + base;
    __syncwarp();
    if (lane == 0) read_run_offsets[r] = o;  // becomes the global offset the host receives
    for (u64 b = 0; b < ns; b += 32) {
      const u64 i = b + lane;
      const bool start = i < ns && codes[s0 + i] == RUN_START;
      const u32 m = __ballot_sync(0xffffffffu, start);
      if (start) {
        const u64 dst = o + __popc(m & ((1u << lane) - 1u));
        if (dst < cap) store_hit(runs + dst, hits[s0 + i]);
      }
      o += __popc(m);
    }
  }
}
// sync-free publication of a chunk's runs: offsets become global, the chunk-local run records are copied into the caller's
// (device-addressable, usually pinned host) buffer with fully coalesced 16-byte stores -- writing them one record at a time
// from the fill kernel crossed PCIe as tiny transactions
__global__ void __launch_bounds__(256) hit_run_publish_kernel(u64* __restrict__ read_run_offsets, u64 n_reads, const Hit* __restrict__ local_runs,
                                                              const u64* __restrict__ chunk_total, const u64* __restrict__ base_ptr, u64 cap,
                                                              Hit* __restrict__ runs) {
  const u64 base = *base_ptr, total = *chunk_total;
  const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x, nt = (u64)gridDim.x * blockDim.x;
  for (u64 r = tid; r < n_reads; r += nt) read_run_offsets[r] += base;
  for (u64 i = tid; i < total; i += nt)
    if (base + i < cap) store_hit(runs + base + i, local_runs[i]);
}
__global__ void hit_run_advance_kernel(u64* __restrict__ base_ptr, const u64* __restrict__ chunk_total) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *base_ptr += *chunk_total;
}

// 2-bit packed reads (kmers::SeqVector layout: base j of a read at bits [2j, 2j+2) of its words, A0 C1 G2 T3) -> ASCII, for
// callers that ship reads 2-bit packed over PCIe (0.25 byte per base instead of 1).  n_mask (optional, 1 bit per base, word
// stride ceil(read_len / 64) per read): set bit = the base is not ACGT and becomes 'N'.
__global__ void __launch_bounds__(256) unpack_reads_kernel(const u64* __restrict__ words, const u64* __restrict__ n_mask, u64 n_reads, u64 read_len,
                                                           u8* __restrict__ bases) {
  const u64 wpr = (read_len + 31) / 32, mpr = (read_len + 63) / 64, total = n_reads * read_len;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (u64)gridDim.x * blockDim.x) {
    const u64 r = i / read_len, j = i - r * read_len;
    const u32 c = (u32)(words[r * wpr + (j >> 5)] >> (2 * (j & 31))) & 3u;
    u8 ch = (u8)((0x54474341u >> (8 * c)) & 0xFFu);  // "ACGT"
    if (n_mask && ((n_mask[r * mpr + (j >> 6)] >> (j & 63)) & 1ULL)) ch = (u8)'N';
    bases[i] = ch;
  }
}
// run codes, one byte per slot -> 2 bits per slot (slot s at bits [2 (s & 3), 2 (s & 3) + 2) of byte s >> 2)
__global__ void __launch_bounds__(256) pack_codes_kernel(const u8* __restrict__ codes, u64 n_slots, u8* __restrict__ out) {
  const u64 nb = (n_slots + 3) / 4;
  for (u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (u64)gridDim.x * blockDim.x) {
    u32 x = 0;
#pragma unroll
    for (u32 q = 0; q < 4; ++q)
      if (4 * b + q < n_slots) x |= (u32)(codes[4 * b + q] & 3u) << (2 * q);
    out[b] = (u8)x;
  }
}

// ---------------------------------------------------------------------------------------------
// K4: U2Pos decode / projection
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void occ_range(const IndexView& ix, u32 uid, u64& s, u64& e) {
  s = packed_get(ix.contig_offsets, uid);  // dense_unitig_table.rs:58-63 / :130-135
  e = packed_get(ix.contig_offsets, (u64)uid + 1);
}
__device__ __forceinline__ OccRec occ_decode(const IndexView& ix, u64 i) {
  OccRec o;
  if (ix.u2pos_kind == MAZU_U2POS_DENSE) {  // UnitigOcc::decode_pf1 (index.rs:335-346)
    u64 word = __ldg(ix.ctable_words + i);
    o.ref_id = (u32)(word & 0xFFFFFFFFULL);
    u64 pw = word >> 32;
    o.pos = (u32)(pw & 0x7FFFFFFFULL);
    o.fw = (pw & 0x80000000ULL) ? 1u : 0u;
  } else {  // UnitigOcc::decode_piscem (spt_compact.rs:99-110)
    PackedVecView v{ix.ctable_words, ix.n_occs, ix.ctable_width, 0};
    u64 enc = packed_get(v, i);
    o.ref_id = (u32)(enc >> ix.ref_shift);
    o.pos = (u32)((enc >> 1) & ix.pos_mask);
    o.fw = (u32)(enc & 1ULL);
  }
  return o;
}
// project_onto_u_occ (index.rs:194-216)
__device__ __forceinline__ OccRec project_occ(u32 k, const Hit& h, const OccRec& occ) {
  OccRec m;
  m.ref_id = occ.ref_id;
  m.pos = occ.fw ? h.pos + occ.pos : occ.pos + (h.unitig_len - h.pos) - k;
  u32 o = h.match == IDENTITY_MATCH ? 1u : 0u;
  m.fw = occ.fw ? o : (o ^ 1u);
  return m;
}

// ---------------------------------------------------------------------------------------------
// GetRefPos::get_ref_pos over reads (src/index.rs:156-216 + the read loop of validate_ckmers): reads -> K2UPos ->
// occurrence list -> MappedRefPos, organised by TILES (one chunk of <= QR_CHUNK k-mer positions of one read) so that
// the variable-length output needs a scan over tiles only, not over slots, and no search:
//   pass 1  get_ref_pos_pass1_kernel   the lookups of query_reads_kernel; the warp also sums the occurrence-list lengths
//                                      of its tile's hits while they are in its shared-memory stage -> tile_totals[tile]
//   scan    exclusive scan of tile_totals (n_tiles items: 1/120 of the per-slot scan of the unfused chain)
//   pass 2  get_ref_pos_pass2_kernel   a warp re-reads its tile's hit records (coalesced), scans the list lengths with
//                                      shuffles, and writes the slot offsets and the projected records at tile_base + local
// against the unfused chain (occ_lens_kernel over every slot -> per-slot scan -> fill with a search per tile of the output).
// (A single-kernel version with a ticket counter and decoupled look-back was built first and measured SLOWER than the
// unfused chain, 32 vs 25 ms for 4.8e8 slots: with 4,736 warps in flight a tile walks back through thousands of
// predecessors that have published an aggregate but not yet an inclusive prefix.)
// ---------------------------------------------------------------------------------------------
struct TileMap {  // tile -> (read, first k-mer position): arithmetic for uniform reads, seg_offsets for ragged ones
  const u64* read_offsets;
  const u64* kmer_offsets;
  const u64* seg_offsets;
  u64 n_reads, uniform_len, n_tiles;
  u64 flat_n;  // != 0: no reads at all, the tiles cut a flat array of flat_n hit records (mazu_b200_project_hits)
};
__device__ __forceinline__ void tile_locate(const TileMap& tm, u32 k, u64 tile, u64& beg, u64& len, u64& slot0, u64& c0, u32& n_c) {
  u64 r;
  if (tm.flat_n) {
    beg = len = c0 = 0;
    slot0 = tile * QR_CHUNK;
    n_c = slot0 < tm.flat_n ? (u32)min((u64)QR_CHUNK, tm.flat_n - slot0) : 0u;
    return;
  }
  if (tm.uniform_len) {
    const u64 nk = tm.uniform_len >= k ? tm.uniform_len - k + 1 : 0;
    const u64 cpr = nk <= (u64)QR_CHUNK ? 1 : (nk + QR_CHUNK - 1) / QR_CHUNK;
    r = tile / cpr;
    c0 = (tile - r * cpr) * QR_CHUNK;
    beg = r * tm.uniform_len;
    len = tm.uniform_len;
    slot0 = r * nk;
  } else {
    u64 lo = 0, hi = tm.n_reads;  // largest r with seg_offsets[r] <= tile
    while (hi - lo > 1) {
      u64 mid = (lo + hi) >> 1;
      if (tm.seg_offsets[mid] <= tile) lo = mid; else hi = mid;
    }
    r = lo;
    c0 = (tile - tm.seg_offsets[r]) * QR_CHUNK;
    beg = tm.read_offsets[r];
    len = tm.read_offsets[r + 1] - beg;
    slot0 = tm.kmer_offsets[r];
  }
  const u64 nk = len >= k ? len - k + 1 : 0;
  n_c = c0 < nk ? (u32)min((u64)QR_CHUNK, nk - c0) : 0u;
}

template <int KIND, u32 FAMILY, u32 KW = 0>
__global__ void __launch_bounds__(QR_WARPS * 32, 4) get_ref_pos_pass1_kernel(const __grid_constant__ IndexView ix, const u8* __restrict__ bases,
                                                                            const TileMap tm, Hit* __restrict__ out_hits,
                                                                            unsigned long long* __restrict__ counts, u64* __restrict__ tile_totals) {
  __shared__ WarpStage s_stage[QR_WARPS];
  const u32 lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  WarpStage& S = s_stage[wib];
  const u32 k = kw_k<KW>(ix.unitigs.k);
  constexpr bool SS = KIND == MAZU_K2U_SSHASH;
  u32 n_valid = 0, n_hit = 0;
  for (u64 tile = (u64)blockIdx.x * QR_WARPS + wib; tile < tm.n_tiles; tile += (u64)gridDim.x * QR_WARPS) {
    u64 beg, len, slot0, c0;
    u32 n_c;
    tile_locate(tm, k, tile, beg, len, slot0, c0, n_c);
    __syncwarp();
    u32 total = 0;
    if (n_c) {
      ChunkInfo ci;
      stage_encode(bases + beg, len, c0, n_c, k, lane, S, ci);
      __syncwarp();
      if (SS) stage_buckets<FAMILY, false, KW>(ix, ci, lane, S);
#pragma unroll 1
      for (u32 p = lane; p < n_c; p += 32) {
        Hit h = hit_none(SKIPPED);
        if (chunk_valid(ci, p)) {
          u64 fw = S.fw[p], rc = SS ? S.rc[p] : revcomp(fw, k);
          bool ok = SS ? verify_sshash<FAMILY, KW>(ix, S, p, fw, rc, h)
                       : (KIND == MAZU_K2U_SAMPLED_PFHASH ? sampled_pfhash_k2u_t<FAMILY>(ix, fw, rc, h, nullptr) : pfhash_k2u_t<FAMILY>(ix, fw, rc, h));
          ++n_valid;
          if (ok) {
            ++n_hit;
            u64 s, e;
            occ_range(ix, h.unitig_id, s, e);  // U2Pos::encoded_unitig_occs
            total += (u32)(e - s);
          } else {
            h = hit_none(NO_MATCH);
          }
        }
        store_hit(out_hits + slot0 + c0 + p, h);
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
    if (lane == 0) tile_totals[tile] = total;
  }
  if (counts) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      n_valid += __shfl_xor_sync(0xffffffffu, n_valid, d);
      n_hit += __shfl_xor_sync(0xffffffffu, n_hit, d);
    }
    if (lane == 0 && n_valid) {
      atomicAdd(counts + 0, (unsigned long long)n_valid);
      atomicAdd(counts + 1, (unsigned long long)n_hit);
      atomicAdd(counts + 2, (unsigned long long)(n_valid - n_hit));
    }
  }
}

// project_hits over a flat array of hit records, short lists (GetRefPos::project_hits, src/index.rs:174-216): the same two passes
// without the lookups -- this kernel sums the list lengths of every tile of QR_CHUNK records, the scan runs over tiles, pass 2
// below emits.  Replaces occ_lens_kernel -> per-record scan -> occ_fill_kernel (search per output tile) for hit batches.
__global__ void __launch_bounds__(256) project_tile_totals_kernel(const __grid_constant__ IndexView ix, const Hit* __restrict__ hits, u64 n,
                                                                  u64 n_tiles, u64* __restrict__ tile_totals) {
  const u32 lane = threadIdx.x & 31;
  const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((u64)gridDim.x * blockDim.x) >> 5;
  for (u64 tile = warp; tile < n_tiles; tile += n_warps) {
    const u64 slot0 = tile * QR_CHUNK;
    Hit hs[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const u64 i = slot0 + 32u * t + lane;
      hs[t] = hit_none(NO_MATCH);
      if (i < n) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(hits + i));
        hs[t] = Hit{q.x, q.y, q.z, q.w};
      }
    }
    u32 total = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if ((hs[t].match == IDENTITY_MATCH || hs[t].match == TWIN_MATCH) && (u64)hs[t].unitig_id < ix.unitigs.n_unitigs) {  // ids outside the table: empty lists
        u64 a, e;
        occ_range(ix, hs[t].unitig_id, a, e);
        total += (u32)(e - a);
      }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
    if (lane == 0) tile_totals[tile] = total;
  }
}

// pass 2: tile_base = exclusive scan of tile_totals (n_tiles + 1 entries, the last one the total)
#ifndef MAZU_P2_OCC
#define MAZU_P2_OCC 4  // resident CTAs the tile-wise emit is compiled for (64 registers, no spills): 1 / 4 / 5 / 6 / 8 measured, profiles/experiments/README.md
#endif
__global__ void __launch_bounds__(256, MAZU_P2_OCC) get_ref_pos_pass2_kernel(const __grid_constant__ IndexView ix, const TileMap tm, const Hit* __restrict__ hits,
                                                                const u64* __restrict__ tile_base, u64 n_slots, u64* __restrict__ out_offsets,
                                                                OccRec* __restrict__ out, u64 cap, u64* __restrict__ out_total) {
  const u32 lane = threadIdx.x & 31;
  const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((u64)gridDim.x * blockDim.x) >> 5;
  const u32 k = ix.unitigs.k;
  if (warp == 0 && lane == 0) {
    out_offsets[n_slots] = tile_base[tm.n_tiles];
    if (out_total) *out_total = tile_base[tm.n_tiles];
  }
  for (u64 tile = warp; tile < tm.n_tiles; tile += n_warps) {
    u64 beg, len, slot0, c0;
    u32 n_c;
    tile_locate(tm, k, tile, beg, len, slot0, c0, n_c);
    u64 o = tile_base[tile];
    // the tile's four groups of 32 slots: every group's loads (hit record -> list bounds) are issued before any group is
    // scanned or written, so a warp has four independent dependent-load chains in flight instead of one
    Hit hs[4];
    u64 firsts[4];
    u32 cnts[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const u32 p = 32u * t + lane;
      hs[t] = hit_none(NO_MATCH);
      if (p < n_c) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(hits + slot0 + c0 + p));
        hs[t] = Hit{q.x, q.y, q.z, q.w};
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      firsts[t] = 0;
      cnts[t] = 0;
      if (32u * t + lane < n_c && (hs[t].match == IDENTITY_MATCH || hs[t].match == TWIN_MATCH) && (u64)hs[t].unitig_id < ix.unitigs.n_unitigs) {
        u64 e;
        occ_range(ix, hs[t].unitig_id, firsts[t], e);
        cnts[t] = (u32)(e - firsts[t]);
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const u32 p = 32u * t + lane;
      if (32u * t >= n_c) break;  // (warp-uniform)
      const bool mine = p < n_c;
      const Hit h = hs[t];
      const u64 first = firsts[t];
      const u32 cnt = cnts[t];
      u32 inc = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const u32 y = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (u32)d) inc += y;
      }
      const u64 o0 = o + inc - cnt;
      if (mine) out_offsets[slot0 + c0 + p] = o0;
      // short lists are written by their lane, long ones by the whole warp (project_onto_u_occ, index.rs:194-216)
      u32 big = out ? __ballot_sync(0xffffffffu, cnt >= 32u) : 0u;  // out == nullptr: offsets only (a sizing call)
      if (out && cnt < 32u)
        for (u32 j = 0; j < cnt; ++j)
          if (o0 + j < cap) out[o0 + j] = project_occ(k, h, occ_decode(ix, first + j));
      while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        Hit hb;
        hb.unitig_id = __shfl_sync(0xffffffffu, h.unitig_id, src);
        hb.unitig_len = __shfl_sync(0xffffffffu, h.unitig_len, src);
        hb.pos = __shfl_sync(0xffffffffu, h.pos, src);
        hb.match = __shfl_sync(0xffffffffu, h.match, src);
        const u64 fb = __shfl_sync(0xffffffffu, first, src), ob = __shfl_sync(0xffffffffu, o0, src);
        const u32 nb = __shfl_sync(0xffffffffu, cnt, src);
        for (u32 j = lane; j < nb; j += 32)
          if (ob + j < cap) out[ob + j] = project_occ(k, hb, occ_decode(ix, fb + j));
      }
      o += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fused hit runs: reads -> K2UPos -> run codes + run records in ONE kernel (the host result of a read batch, see "Hit runs"
// above), for batches whose reads all fit one chunk (<= QR_CHUNK k-mer positions: every short-read batch).  A warp owns a
// read: it derives every slot's code from its neighbour's hit with shuffles, counts the run starts, reserves that many
// records with ONE atomicAdd on the batch's run cursor, and writes one code byte per slot + the 16-byte record of every
// run start.  The runs of a read are consecutive; the order of different reads' runs in the array is whatever order the
// warps reserved them in -- read_run_offsets[r] says where read r's runs start, which is all the decoder needs.
// Replaces query_reads_kernel (16 B per slot written) + hit_run_codes_kernel (16 B per slot read) + scan + hit_run_fill_kernel.
// (An ordered single-pass layout was built first -- tickets + decoupled look-back over one status word per tile -- and
// measured 2x slower than the kernels it replaced: with 4,736 warps in flight a tile walks back through thousands of
// predecessors that have published an aggregate but not yet an inclusive prefix.)
// ---------------------------------------------------------------------------------------------
struct RunsTileOut {
  u8* codes;                   // one byte per k-mer slot ...
  u8* codes2;                  // ... or, when set, 2 bits per slot (slot s at bits [2 (s & 3), +2) of byte s >> 2; uniform reads whose
                               // slot count is a multiple of 4, so every read owns whole bytes)
  Hit* runs;                   // run records, chunk-local indexes
  u64* read_run_offsets;       // n_reads entries: index of every read's first run
  unsigned long long* cursor;  // next free run record (zeroed before the launch; the batch's run count afterwards)
  u64 cap;                     // capacity of `runs`
  // interval mode (mazu_b200_query_reads_intervals_packed): when set, NOTHING above but cursor / cap is written; every run
  // leaves as one self-contained 16-byte record {unitig id, position | twin << 31, read, first slot | length << 16}
  uint4* intervals;
  u64 read_base;               // index of the chunk's first read in the caller's batch
};

// IO: 0 = ASCII reads in, one code byte per slot out; 1 = 2-bit packed reads in, 2-bit codes out; 2 = packed reads in, interval
// records out; 3 = ASCII reads in, interval records out.  One instantiation per mode: with all of them in one body the kernel
// grew to 4,000 SASS instructions and every mode lost 20 % to instruction fetch (profiles/experiments/README.md).
enum { RUNS_IO_ASCII = 0, RUNS_IO_PACKED = 1, RUNS_IO_INTERVALS = 2, RUNS_IO_INTERVALS_ASCII = 3 };
template <int KIND, u32 FAMILY, int IO, u32 KW = 0>
__global__ void __launch_bounds__(QR_WARPS * 32, 4) query_reads_runs_kernel(const __grid_constant__ IndexView ix, const u8* __restrict__ bases,
                                                                           const u64* __restrict__ read_offsets, u64 n_reads, u64 uniform_len,
                                                                           const u64* __restrict__ kmer_offsets, unsigned long long* __restrict__ counts,
                                                                           const RunsTileOut ro, const u64* __restrict__ packed_words,
                                                                           const u64* __restrict__ packed_nmask) {
  __shared__ WarpStage s_stage[QR_WARPS];
  const u32 lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  WarpStage& S = s_stage[wib];
  const u32 k = kw_k<KW>(ix.unitigs.k);
  constexpr bool SS = KIND == MAZU_K2U_SSHASH;
  u32 n_valid = 0, n_hit = 0;
  for (u64 r = (u64)blockIdx.x * QR_WARPS + wib; r < n_reads; r += (u64)gridDim.x * QR_WARPS) {
    u64 beg, len, slot0;
    if (uniform_len) {
      beg = r * uniform_len;
      len = uniform_len;
      slot0 = uniform_len >= k ? r * (uniform_len - k + 1) : 0;
    } else {
      beg = read_offsets[r];
      len = read_offsets[r + 1] - beg;
      slot0 = kmer_offsets[r];
    }
    const u32 n_c = len >= k ? (u32)min((u64)QR_CHUNK, len - k + 1) : 0u;  // the launcher guarantees len - k + 1 <= QR_CHUNK
    __syncwarp();
    Hit hh[4];
    u32 code[4], n_runs = 0;
    if (n_c) {
      ChunkInfo ci;
      if (IO == RUNS_IO_PACKED || IO == RUNS_IO_INTERVALS)  // uniform reads, 2-bit packed
        stage_encode_packed(packed_words + r * ((len + 31) / 32), packed_nmask ? packed_nmask + r * ((len + 63) / 64) : nullptr, (u32)len, n_c, k, lane, S, ci);
      else
        stage_encode(bases + beg, len, 0, n_c, k, lane, S, ci);
      __syncwarp();
      if (SS) stage_buckets<FAMILY, false, KW>(ix, ci, lane, S);
      // one copy of the lookup (the loop is not unrolled: instruction-cache footprint); hits are parked in the warp's stage
#pragma unroll 1
      for (u32 p = lane; p < n_c; p += 32) {
        Hit h = hit_none(SKIPPED);
        if (chunk_valid(ci, p)) {
          u64 fw = S.fw[p], rc = SS ? S.rc[p] : revcomp(fw, k);
          bool ok = SS ? verify_sshash<FAMILY, KW>(ix, S, p, fw, rc, h)
                       : (KIND == MAZU_K2U_SAMPLED_PFHASH ? sampled_pfhash_k2u_t<FAMILY>(ix, fw, rc, h, nullptr) : pfhash_k2u_t<FAMILY>(ix, fw, rc, h));
          ++n_valid;
          if (ok) ++n_hit; else h = hit_none(NO_MATCH);
        }
        S.fw[p] = (u64)h.unitig_id | ((u64)h.pos << 32);
        S.rc[p] = (u64)h.unitig_len | ((u64)h.match << 32);
      }
      __syncwarp();
      Hit carry = hit_none(NO_MATCH);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const u32 p = 32 * t + lane;
        Hit h = hit_none(SKIPPED);
        if (p < n_c) {
          const u64 a = S.fw[p], b = S.rc[p];
          h = Hit{(u32)a, (u32)b, (u32)(a >> 32), (u32)(b >> 32)};
        }
        Hit pv;
        pv.unitig_id = __shfl_up_sync(0xffffffffu, h.unitig_id, 1);
        pv.unitig_len = 0;
        pv.pos = __shfl_up_sync(0xffffffffu, h.pos, 1);
        pv.match = __shfl_up_sync(0xffffffffu, h.match, 1);
        if (lane == 0) pv = carry;
        code[t] = run_code_of(h, pv, p > 0);
        hh[t] = h;
        n_runs += __popc(__ballot_sync(0xffffffffu, p < n_c && code[t] == RUN_START));
        carry.unitig_id = __shfl_sync(0xffffffffu, h.unitig_id, 31);
        carry.pos = __shfl_sync(0xffffffffu, h.pos, 31);
        carry.match = __shfl_sync(0xffffffffu, h.match, 31);
      }
    }
    unsigned long long E = 0;
    if (lane == 0) {
      E = n_runs ? atomicAdd(ro.cursor, (unsigned long long)n_runs) : 0ULL;
      if (IO < RUNS_IO_INTERVALS) ro.read_run_offsets[r] = E;
    }
    if (IO >= RUNS_IO_INTERVALS) {
      if (n_runs) {  // (warp-uniform)
        // a run = its start slot + the CONT slots that follow it: 128 continuation flags as two 64-bit words, inverted so the
        // run's end is a find-first-set
        u32 cm[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) cm[t] = __ballot_sync(0xffffffffu, 32u * t + lane < n_c && code[t] == RUN_CONT);
        const u64 stop_lo = ~((u64)cm[0] | ((u64)cm[1] << 32)), stop_hi = ~((u64)cm[2] | ((u64)cm[3] << 32));
        u64 o = __shfl_sync(0xffffffffu, E, 0);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const u32 p = 32 * t + lane;
          const bool start = p < n_c && code[t] == RUN_START;
          const u32 m = __ballot_sync(0xffffffffu, start);
          if (start) {
            const u32 q = p + 1;  // first slot that may end the run
            u32 end = QR_CHUNK;
            if (q < 64) {
              const u64 a = stop_lo >> q;
              if (a) end = q + (u32)__ffsll((long long)a) - 1;
              else if (stop_hi) end = 64 + (u32)__ffsll((long long)stop_hi) - 1;
            } else if (q < 128) {
              const u64 a = stop_hi >> (q - 64);
              if (a) end = q + (u32)__ffsll((long long)a) - 1;
            }
            const u64 dst = o + __popc(m & ((1u << lane) - 1u));
            if (dst < ro.cap)
              ro.intervals[dst] = make_uint4(hh[t].unitig_id, hh[t].pos | (hh[t].match == TWIN_MATCH ? 0x80000000u : 0u), (u32)(ro.read_base + r),
                                             p | ((end - p) << 16));
          }
          o += __popc(m);
        }
      }
      continue;
    }
    if (n_c) {  // the codes do not depend on the reservation: they go out while the atomic is in flight
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const u32 p = 32 * t + lane;
        if (IO == RUNS_IO_PACKED) {  // four neighbouring lanes share a byte
          u32 v = p < n_c ? (code[t] & 3u) << (2 * (lane & 3u)) : 0u;
          v |= __shfl_xor_sync(0xffffffffu, v, 1);
          v |= __shfl_xor_sync(0xffffffffu, v, 2);
          if ((lane & 3u) == 0 && p < n_c) ro.codes2[(slot0 + p) >> 2] = (u8)v;
        } else if (p < n_c) {
          ro.codes[slot0 + p] = (u8)code[t];
        }
      }
    }
    E = __shfl_sync(0xffffffffu, E, 0);
    if (n_c) {
      u64 o = E;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const u32 p = 32 * t + lane;
        const bool start = p < n_c && code[t] == RUN_START;
        const u32 m = __ballot_sync(0xffffffffu, start);
        if (start) {
          const u64 dst = o + __popc(m & ((1u << lane) - 1u));
          if (dst < ro.cap) store_hit(ro.runs + dst, hh[t]);
        }
        o += __popc(m);
      }
    }
  }
  if (counts) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      n_valid += __shfl_xor_sync(0xffffffffu, n_valid, d);
      n_hit += __shfl_xor_sync(0xffffffffu, n_hit, d);
    }
    if (lane == 0 && n_valid) {
      atomicAdd(counts + 0, (unsigned long long)n_valid);
      atomicAdd(counts + 1, (unsigned long long)n_hit);
      atomicAdd(counts + 2, (unsigned long long)(n_valid - n_hit));
    }
  }
}

// list lengths: from unitig ids (hits == nullptr) or from hit records
__global__ void occ_lens_kernel(const __grid_constant__ IndexView ix, const u32* __restrict__ uids, const Hit* __restrict__ hits, u64 n,
                                u64* __restrict__ lens) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    u32 uid;
    if (hits) {
      Hit h = hits[i];
      uid = (h.match == IDENTITY_MATCH || h.match == TWIN_MATCH) ? h.unitig_id : ~0u;
    } else {
      uid = uids[i];
    }
    u64 len = 0;
    if ((u64)uid < ix.unitigs.n_unitigs) {  // ~0 (a miss) and any other id outside the table: empty list
      u64 s, e;
      occ_range(ix, uid, s, e);
      len = e - s;
    }
    lens[i] = len;
  }
}
// largest q in [0, n) with off[q] <= x, found by the whole warp with a 32-ary search (off[0] == 0 <= x)
__device__ __forceinline__ u64 warp_last_le(const u64* __restrict__ off, u64 n, u64 x, u32 lane) {
  u64 lo = 0, hi = n;  // answer in [lo, hi)
  while (hi - lo > 1) {
    u64 step = (hi - lo + 31) / 32;
    u64 probe = lo + (u64)(lane + 1) * step;
    bool le = probe < hi && __ldg(off + probe) <= x;
    u32 c = __popc(__ballot_sync(0xffffffffu, le));  // off is monotone: the lanes that pass form a prefix
    lo += (u64)c * step;
    hi = min(hi, lo + step);
  }
  return lo;
}

// fill, load-balanced over the OUTPUT: a warp owns a tile of 1024 consecutive output records, locates the
// queries overlapping the tile once (32-ary search over the scanned offsets), then every lane maps its
// records to (query, element) with a short binary search inside that range.  A 65,536-entry list is
// therefore spread over 64 warps instead of serialising one, and consecutive lanes read consecutive
// occurrence words and write consecutive 12-byte records.
static const u64 OCC_TILE = 1024;
// one tile [t0, t1) of the output, overlapping the queries qlo..qhi (generic path: plain loads and stores)
template <bool PROJECT>
__device__ __forceinline__ void occ_fill_tile(const IndexView& ix, const u32* __restrict__ uids, const Hit* __restrict__ hits,
                                              const u64* __restrict__ out_offsets, OccRec* __restrict__ out, u64 t0, u64 t1, u64 qlo, u64 qhi,
                                              u32 lane, u32 k, bool out_aligned) {
  // callers clip t1 to the capacity of `out`: records beyond it are never computed
  if (qhi - qlo <= 32) {
    // few, long lists in this tile: walk the overlapping queries (warp-uniform), lanes stride over each
    // segment with plain arithmetic -- no per-record search; 4 independent records in flight per lane
    for (u64 q = qlo; q <= qhi; ++q) {
      const u64 ob = __ldg(out_offsets + q), oe = __ldg(out_offsets + q + 1);
      const u64 sb = max(ob, t0), se = min(oe, t1);
      if (sb >= se) continue;
      Hit h = hit_none(NO_MATCH);
      u32 uid;
      if (PROJECT) {
        h = hits[q];
        uid = h.unitig_id;
      } else {
        uid = uids[q];
      }
      const u64 e0 = packed_get(ix.contig_offsets, uid) - ob;  // element index = e0 + rec
      // body: every lane owns 4 CONSECUTIVE records = 48 contiguous output bytes = three 16-byte stores
      // (record index a multiple of 4 <=> byte offset a multiple of 16); head/tail records go one by one
      u64 body_b = (sb + 3) & ~3ULL, body_e = body_b + ((se > body_b ? se - body_b : 0) & ~3ULL);
      if (!out_aligned || body_b >= se) body_b = body_e = sb;
      for (u64 rec = sb + lane; rec < body_b; rec += 32) {
        OccRec o = occ_decode(ix, e0 + rec);
        out[rec] = PROJECT ? project_occ(k, h, o) : o;
      }
      for (u64 base = body_b + 4 * lane; base < body_e; base += 128) {
        OccRec o[4];
        if (ix.u2pos_kind == MAZU_U2POS_PISCEM && ix.ctable_width <= 48) {
          // four consecutive packed fields span at most 4 words (4 * 48 + 63 < 256 bits): 4 loads instead of 8
          const u32 wd = ix.ctable_width;
          const u64 bit0 = (e0 + base) * wd, wi = bit0 >> 6;
          const u64 w0 = __ldg(ix.ctable_words + wi), w1 = __ldg(ix.ctable_words + wi + 1), w2 = __ldg(ix.ctable_words + wi + 2),
                    w3 = __ldg(ix.ctable_words + wi + 3);
          const u64 fmask = (1ULL << wd) - 1ULL;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const u32 b = (u32)(bit0 & 63) + (u32)j * wd;  // < 64 + 3 * 48 = 208
            const u32 q = b >> 6, sh = b & 63u;
            const u64 lo = q == 0 ? w0 : (q == 1 ? w1 : (q == 2 ? w2 : w3)), hi = q == 0 ? w1 : (q == 1 ? w2 : w3);
            const u64 enc = (sh ? (lo >> sh) | (hi << (64 - sh)) : lo) & fmask;
            o[j].ref_id = (u32)(enc >> ix.ref_shift);
            o[j].pos = (u32)((enc >> 1) & ix.pos_mask);
            o[j].fw = (u32)(enc & 1ULL);
            if (PROJECT) o[j] = project_occ(k, h, o[j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            o[j] = occ_decode(ix, e0 + base + j);
            if (PROJECT) o[j] = project_occ(k, h, o[j]);
          }
        }
        uint4* dst = reinterpret_cast<uint4*>(out + base);
        dst[0] = make_uint4(o[0].ref_id, o[0].pos, o[0].fw, o[1].ref_id);
        dst[1] = make_uint4(o[1].pos, o[1].fw, o[2].ref_id, o[2].pos);
        dst[2] = make_uint4(o[2].fw, o[3].ref_id, o[3].pos, o[3].fw);
      }
      for (u64 rec = body_e + lane; rec < se; rec += 32) {
        OccRec o = occ_decode(ix, e0 + rec);
        out[rec] = PROJECT ? project_occ(k, h, o) : o;
      }
    }
    return;
  }
  for (u64 rec = t0 + lane; rec < t1; rec += 32) {
    u64 a = qlo, b = qhi;  // off[a] <= rec, answer in [a, b]
    while (a < b) {
      u64 m = (a + b + 1) >> 1;
      if (__ldg(out_offsets + m) <= rec) a = m; else b = m - 1;
    }
    Hit h = hit_none(NO_MATCH);
    u32 uid;
    if (PROJECT) {
      h = hits[a];
      uid = h.unitig_id;
    } else {
      uid = uids[a];
    }
    u64 s = packed_get(ix.contig_offsets, uid);  // dense_unitig_table.rs:58-63 / :130-135
    OccRec o = occ_decode(ix, s + (rec - __ldg(out_offsets + a)));
    if (PROJECT) o = project_occ(k, h, o);
    out[rec] = o;
  }
}

template <bool PROJECT>
// (256, 1): ptxas takes ~51 registers and keeps a lane's four packed-word loads in flight; forcing 6 or 8 resident CTAs
// (40 / 32 registers, spills) measured 5 % / 18 % slower
__global__ void __launch_bounds__(256, 1) occ_fill_kernel(const __grid_constant__ IndexView ix, const u32* __restrict__ uids,
                                                       const Hit* __restrict__ hits, u64 n, const u64* __restrict__ out_offsets,
                                                       OccRec* __restrict__ out, u64 cap) {
  const u32 lane = threadIdx.x & 31;
  const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u64 n_warps = ((u64)gridDim.x * blockDim.x) >> 5;
  const u32 k = ix.unitigs.k;
  const u64 total = min(__ldg(out_offsets + n), cap);  // an undersized buffer is filled up to its capacity, never beyond
  const bool out_aligned = (reinterpret_cast<unsigned long long>(out) & 15ULL) == 0;
  for (u64 t0 = warp * OCC_TILE; t0 < total; t0 += n_warps * OCC_TILE) {
    const u64 t1 = min(t0 + OCC_TILE, total);
    const u64 qlo = warp_last_le(out_offsets, n, t0, lane);
    const u64 qhi = warp_last_le(out_offsets, n, t1 - 1, lane);
    occ_fill_tile<PROJECT>(ix, uids, hits, out_offsets, out, t0, t1, qlo, qhi, lane, k, out_aligned);
  }
}

// ---------------------------------------------------------------------------------------------
// K4 with bulk-copy (TMA) staging.  A full tile that lies inside ONE occurrence list -- the case that carries almost all
// bytes when lists are long -- reads a contiguous bit range of the table and writes contiguous records.  Each warp owns
// two input buffers and one output buffer in shared memory:
//   lane 0   cp.async.bulk global -> shared of the NEXT tile's packed words into the idle input buffer (completion on that
//            buffer's mbarrier) before the warp touches the current tile: a load is always in flight per warp
//   warp     waits on the current buffer's mbarrier, decodes fields out of shared memory into 12-byte records
//   lane 0   fence.proxy.async + cp.async.bulk shared -> global of the records (bulk group); the output buffer is reused
//            after cp.async.bulk.wait_group.read
// No register staging, no per-lane global addresses.  A warp walks runs of OCC_TMA_RUN consecutive tiles, so the search
// that maps a tile to its list is done once per run while the list lasts.  Tiles that straddle lists, the last partial
// tile, and tables wider than 48 bits take occ_fill_tile().  One CTA of 24 warps per SM (174 KB of dynamic shared memory);
// 12 warps with 512-record tiles measured 2 % (decode) / 3.5 % (projection) slower.
// ---------------------------------------------------------------------------------------------
constexpr u32 OCC_TMA_WARPS = 24;
constexpr u32 OCC_TMA_TILE = 256;        // records per staged tile
constexpr u32 OCC_TMA_RUN = 16;          // consecutive tiles a warp takes at a time
constexpr u32 OCC_TMA_IN_BYTES = 2176;   // 256 x 64 bit (pf1 words) + alignment slack, a multiple of 128
constexpr u32 OCC_TMA_OUT_BYTES = OCC_TMA_TILE * 12;
constexpr u32 OCC_TMA_WARP_BYTES = 2 * OCC_TMA_IN_BYTES + OCC_TMA_OUT_BYTES;
constexpr u32 OCC_TMA_SMEM = OCC_TMA_WARPS * OCC_TMA_WARP_BYTES + OCC_TMA_WARPS * 16 + 128;

struct OccTilePlan {
  u64 t0, t1;
  u64 qlo, qhi;
  const char* src;  // 16-byte aligned start of the bulk load
  u32 bytes;        // multiple of 16
  u32 shift;        // bit offset of the tile's first field inside the staged bytes
  u32 tma;          // 1: staged path, 0: generic path
  u32 valid;
};
struct OccListCache {  // the list the previous tile of this warp was in
  u64 q, ob, oe, e0;
  u32 valid;
};

template <bool PROJECT>
__device__ __forceinline__ void occ_plan_tile(const IndexView& ix, const u32* __restrict__ uids, const Hit* __restrict__ hits, u64 n,
                                              const u64* __restrict__ out_offsets, u64 total, u32 lane, bool stage_ok, OccListCache& c,
                                              OccTilePlan& p) {
  p.t1 = min(p.t0 + OCC_TMA_TILE, total);
  p.src = nullptr;
  p.bytes = p.shift = p.tma = 0;
  if (c.valid && p.t0 >= c.ob && p.t1 <= c.oe) {
    p.qlo = p.qhi = c.q;
  } else {
    p.qlo = warp_last_le(out_offsets, n, p.t0, lane);
    p.qhi = warp_last_le(out_offsets, n, p.t1 - 1, lane);
    c.valid = 0;
    if (p.qlo == p.qhi) {
      const u32 uid = PROJECT ? hits[p.qlo].unitig_id : uids[p.qlo];
      c.q = p.qlo;
      c.ob = __ldg(out_offsets + p.qlo);
      c.oe = __ldg(out_offsets + p.qlo + 1);
      c.e0 = packed_get(ix.contig_offsets, uid);
      c.valid = 1;
    }
  }
  if (stage_ok && c.valid && p.t1 - p.t0 == OCC_TMA_TILE) {
    const u64 e_first = c.e0 + (p.t0 - c.ob);
    const u32 wd = ix.u2pos_kind == MAZU_U2POS_DENSE ? 64u : ix.ctable_width;
    const u64 bit0 = e_first * wd, bit1 = bit0 + (u64)OCC_TMA_TILE * wd;
    const u64 byte0 = (bit0 >> 3) & ~15ULL, byte1 = (((bit1 + 7) >> 3) + 8 + 15) & ~15ULL;  // + 8: the decoder reads one word past a field
    p.src = reinterpret_cast<const char*>(ix.ctable_words) + byte0;
    p.bytes = (u32)(byte1 - byte0);
    p.shift = (u32)(bit0 - 8 * byte0);
    p.tma = p.bytes <= OCC_TMA_IN_BYTES ? 1u : 0u;
  }
}

template <bool PROJECT>
__global__ void __launch_bounds__(OCC_TMA_WARPS * 32, 1) occ_fill_tma_kernel(const __grid_constant__ IndexView ix, const u32* __restrict__ uids,
                                                                              const Hit* __restrict__ hits, u64 n,
                                                                              const u64* __restrict__ out_offsets, OccRec* __restrict__ out, u64 cap) {
  extern __shared__ __align__(128) unsigned char occ_smem[];
  namespace ptx = cuda::ptx;
  const u32 lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const u64 warp = (u64)blockIdx.x * OCC_TMA_WARPS + wib, n_warps = (u64)gridDim.x * OCC_TMA_WARPS;
  const u32 k = ix.unitigs.k;
  const u64 total = min(__ldg(out_offsets + n), cap);  // an undersized buffer is filled up to its capacity, never beyond
  const bool out_aligned = (reinterpret_cast<unsigned long long>(out) & 15ULL) == 0;
  const bool stage_ok = out_aligned && (ix.u2pos_kind == MAZU_U2POS_DENSE || ix.ctable_width <= 48);
  unsigned char* in_buf[2];
  in_buf[0] = occ_smem + wib * OCC_TMA_WARP_BYTES;
  in_buf[1] = in_buf[0] + OCC_TMA_IN_BYTES;
  unsigned char* out_buf = in_buf[1] + OCC_TMA_IN_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(occ_smem + OCC_TMA_WARPS * OCC_TMA_WARP_BYTES) + 2 * wib;
  if (lane == 0) {
    ptx::mbarrier_init(bar, 1);
    ptx::mbarrier_init(bar + 1, 1);
  }
  ptx::fence_proxy_async();  // the initialised barriers become visible to the async proxy
  __syncwarp();
  u32 parity0 = 0, parity1 = 0, buf = 0;  // buf: input buffer the CURRENT tile uses
  auto issue_load = [&](const OccTilePlan& p, u32 b) {
    if (lane == 0) {
      ptx::mbarrier_arrive_expect_tx(ptx::sem_release, ptx::scope_cta, ptx::space_shared, bar + b, p.bytes);
      ptx::cp_async_bulk(ptx::space_cluster, ptx::space_global, in_buf[b], p.src, p.bytes, bar + b);
    }
  };
  // this warp's tiles: runs of OCC_TMA_RUN consecutive tiles, runs strided over all warps
  const u64 run_bytes = (u64)OCC_TMA_RUN * OCC_TMA_TILE;
  u64 run0 = warp * run_bytes;
  u32 j = 0;
  auto next_t0 = [&](u64& t0) -> bool {  // successor of the tile starting at t0
    ++j;
    if (j == OCC_TMA_RUN) {
      j = 0;
      run0 += n_warps * run_bytes;
    }
    t0 = run0 + (u64)j * OCC_TMA_TILE;
    return t0 < total;
  };
  OccListCache cache;
  cache.valid = 0;
  OccTilePlan cur, nxt;
  cur.t0 = run0;
  cur.valid = cur.t0 < total ? 1u : 0u;
  if (!cur.valid) return;
  occ_plan_tile<PROJECT>(ix, uids, hits, n, out_offsets, total, lane, stage_ok, cache, cur);
  if (cur.tma) issue_load(cur, buf);
  while (true) {
    nxt.t0 = cur.t0;
    nxt.valid = next_t0(nxt.t0) ? 1u : 0u;
    nxt.tma = 0;
    if (nxt.valid) {
      occ_plan_tile<PROJECT>(ix, uids, hits, n, out_offsets, total, lane, stage_ok, cache, nxt);
      if (nxt.tma) issue_load(nxt, cur.tma ? (buf ^ 1u) : buf);  // the idle input buffer
    }
    if (cur.tma) {
      if (buf == 0) {
        while (!ptx::mbarrier_try_wait_parity(bar, parity0)) {
        }
        parity0 ^= 1u;
      } else {
        while (!ptx::mbarrier_try_wait_parity(bar + 1, parity1)) {
        }
        parity1 ^= 1u;
      }
      if (lane == 0) ptx::cp_async_bulk_wait_group_read(ptx::n32_t<0>{});  // the previous store has finished reading out_buf
      __syncwarp();
      Hit h = hit_none(NO_MATCH);
      if (PROJECT) h = hits[cur.qlo];
      const u64* in64 = reinterpret_cast<const u64*>(in_buf[buf]);
      u32* o32 = reinterpret_cast<u32*>(out_buf);
      const bool dense = ix.u2pos_kind == MAZU_U2POS_DENSE;
      const u32 wd = dense ? 64u : ix.ctable_width;
      const u64 fmask = wd >= 64 ? ~0ULL : ((1ULL << wd) - 1ULL);
#pragma unroll 4
      for (u32 r = lane; r < OCC_TMA_TILE; r += 32) {
        const u32 bit = cur.shift + r * wd, w = bit >> 6, sh = bit & 63u;
        const u64 lo = in64[w], hi = in64[w + 1];
        const u64 enc = (sh ? (lo >> sh) | (hi << (64 - sh)) : lo) & fmask;
        OccRec oc;
        if (dense) {  // UnitigOcc::decode_pf1 (index.rs:335-346)
          oc.ref_id = (u32)(enc & 0xFFFFFFFFULL);
          oc.pos = (u32)((enc >> 32) & 0x7FFFFFFFULL);
          oc.fw = (u32)(enc >> 63);
        } else {  // UnitigOcc::decode_piscem (spt_compact.rs:99-110)
          oc.ref_id = (u32)(enc >> ix.ref_shift);
          oc.pos = (u32)((enc >> 1) & ix.pos_mask);
          oc.fw = (u32)(enc & 1ULL);
        }
        if (PROJECT) oc = project_occ(k, h, oc);
        o32[3 * r] = oc.ref_id;  // stride of 3 words across lanes: conflict-free
        o32[3 * r + 1] = oc.pos;
        o32[3 * r + 2] = oc.fw;
      }
      ptx::fence_proxy_async();  // this lane's generic-proxy writes to out_buf -> visible to the async proxy (the bulk store)
      __syncwarp();              // every lane is done with the input buffer and has written (and fenced) out_buf
      if (lane == 0) {
        ptx::cp_async_bulk(ptx::space_global, ptx::space_shared, reinterpret_cast<char*>(out + cur.t0), out_buf, u32(OCC_TMA_OUT_BYTES));
        ptx::cp_async_bulk_commit_group();
      }
      if (nxt.tma) buf ^= 1u;  // the next staged tile was loaded into the other buffer
    } else {
      occ_fill_tile<PROJECT>(ix, uids, hits, out_offsets, out, cur.t0, cur.t1, cur.qlo, cur.qhi, lane, k, out_aligned);
    }
    if (!nxt.valid) break;
    cur = nxt;
  }
  if (lane == 0) ptx::cp_async_bulk_wait_group_read(ptx::n32_t<0>{});  // shared memory must outlive the last store's reads
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// GPU validate drivers.  counts = {n_queries, n_identity, n_twin, n_projected, n_fail}
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 packed2_window(const u64* __restrict__ words, u64 pos, u32 k) {
  u64 bit = 2 * pos, wi = bit >> 6;
  u32 sh = (u32)(bit & 63);
  u64 x = __ldg(words + wi) >> sh;
  if (sh + 2 * k > 64) x |= __ldg(words + wi + 1) << (64 - sh);
  return x & kmer_mask(k);
}
__device__ __forceinline__ void block_accumulate(unsigned long long* counts, u32 v[5]) {
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    u32 x = v[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0 && x) atomicAdd(counts + j, (unsigned long long)x);
  }
}
// Validate::validate_self (src/index/validate.rs:24-52): thread per reference position
__global__ void __launch_bounds__(256) validate_self_kernel(const __grid_constant__ IndexView ix, unsigned long long* counts) {
  const u32 k = ix.unitigs.k;
  const u64 total = ix.ref_prefix[ix.n_refs];
  u32 v[5] = {0, 0, 0, 0, 0};
  for (u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (u64)gridDim.x * blockDim.x) {
    // reference containing p (few references: binary search)
    u64 lo = 0, hi = ix.n_refs;
    while (hi - lo > 1) {
      u64 mid = (lo + hi) >> 1;
      if (ix.ref_prefix[mid] <= p) lo = mid; else hi = mid;
    }
    u64 rs = ix.ref_prefix[lo], re = ix.ref_prefix[lo + 1];
    if (p + k > re) continue;
    u64 fw = packed2_window(ix.refseq, p, k), rc = revcomp(fw, k);
    Hit h;
    ++v[0];
    if (!k2u_any(ix, fw, rc, h)) {
      ++v[4];
      continue;
    }
    ++v[h.match == IDENTITY_MATCH ? 1 : 2];
    u64 s, e;
    occ_range(ix, h.unitig_id, s, e);
    bool found = false;
    for (u64 j = s; j < e; ++j) {
      OccRec m = project_occ(k, h, occ_decode(ix, j));
      found |= (m.pos == (u32)(p - rs)) && (m.ref_id == (u32)lo);
    }
    v[3] += (u32)(e - s);
    if (!found) ++v[4];
  }
  block_accumulate(counts, v);
}
// Validate::validate_ckmers over a batch of records (src/index/validate.rs:54-81, src/index/caching.rs:175-201): thread per
// k-mer slot of the hit records query_reads produced; record r of the batch is reference ref0 + r.  Every yielded k-mer must
// have a projected position (ref0 + r, position of the slot in its record).
__global__ void __launch_bounds__(256) validate_reads_kernel(const __grid_constant__ IndexView ix, const Hit* __restrict__ hits,
                                                             const u64* __restrict__ kmer_offsets, u64 n_reads, u64 ref0,
                                                             unsigned long long* counts) {
  const u32 k = ix.unitigs.k;
  const u64 total = kmer_offsets[n_reads];
  u32 v[5] = {0, 0, 0, 0, 0};
  for (u64 slot = (u64)blockIdx.x * blockDim.x + threadIdx.x; slot < total; slot += (u64)gridDim.x * blockDim.x) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(hits + slot));
    const Hit h{q.x, q.y, q.z, q.w};
    if (h.match == SKIPPED) continue;  // a window the CanonicalKmerIterator does not yield
    ++v[0];
    if (h.match == NO_MATCH) {  // "No MRPs found for true +ve kmer"
      ++v[4];
      continue;
    }
    ++v[h.match == IDENTITY_MATCH ? 1 : 2];
    u64 lo = 0, hi = n_reads;  // largest r with kmer_offsets[r] <= slot
    while (hi - lo > 1) {
      u64 mid = (lo + hi) >> 1;
      if (kmer_offsets[mid] <= slot) lo = mid; else hi = mid;
    }
    const u64 pos = slot - kmer_offsets[lo], ref_id = ref0 + lo;
    u64 s, e;
    occ_range(ix, h.unitig_id, s, e);
    bool found = false;
    for (u64 j = s; j < e; ++j) {
      OccRec m = project_occ(k, h, occ_decode(ix, j));
      found |= ((u64)m.pos == pos) && ((u64)m.ref_id == ref_id);
    }
    v[3] += (u32)(e - s);
    if (!found) ++v[4];
  }
  block_accumulate(counts, v);
}
// K2U::k2u for every k-mer start of one reference (input of the iter_unitigs_on_ref walk, src/index.rs:396-423)
__global__ void __launch_bounds__(256) ref_hits_kernel(const __grid_constant__ IndexView ix, u64 ref_begin, u64 n_pos, Hit* __restrict__ out) {
  const u32 k = ix.unitigs.k;
  for (u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x; p < n_pos; p += (u64)gridDim.x * blockDim.x) {
    u64 fw = packed2_window(ix.refseq, ref_begin + p, k), rc = revcomp(fw, k);
    Hit h;
    if (!k2u_any(ix, fw, rc, h)) h = hit_none(NO_MATCH);
    store_hit(out + p, h);
  }
}

// K2U::validate_self (src/kphf/mod.rs:69-103): thread per useq position, fw then swapped
__global__ void __launch_bounds__(256) k2u_validate_self_kernel(const __grid_constant__ IndexView ix, unsigned long long* counts) {
  const u32 k = ix.unitigs.k;
  const u64 total = ix.unitigs.total_len;
  u32 v[5] = {0, 0, 0, 0, 0};
  for (u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x; p + k <= total; p += (u64)gridDim.x * blockDim.x) {
    u64 id, s, e;
    unitig_locate(ix.unitigs, p, id, s, e);
    if (p + k > e) continue;  // window straddles a unitig boundary: not a k-mer of the set
    u64 fw = useq_window(ix.unitigs, p), rc = revcomp(fw, k);
    for (int t = 0; t < 2; ++t) {
      Hit h;
      ++v[0];
      u32 want = t == 0 ? IDENTITY_MATCH : TWIN_MATCH;
      bool ok = t == 0 ? k2u_any(ix, fw, rc, h) : k2u_any(ix, rc, fw, h);
      if (!ok) {
        ++v[3];  // not found at all (counts[3] is otherwise unused by this driver): can never happen for a k-mer of the set
        ++v[4];
      } else if (h.unitig_id != (u32)id || h.unitig_len != (u32)(e - s) || h.pos != (u32)(p - s) || h.match != want) {
        ++v[4];  // found, verified, but at another position: the unitig set holds this canonical k-mer twice
      } else {
        ++v[t == 0 ? 1 : 2];
      }
    }
  }
  block_accumulate(counts, v);
}

// Are the canonical k-mers of the unitig set pairwise distinct (true for every compacted de Bruijn graph)?  Thread per useq
// position: the k-mer there must be found AT that position; a hit elsewhere means the set holds the k-mer twice.
//  * No duplicate: StreamingK2U (src/index/caching.rs:65-103) answers exactly what K2U::k2u answers -- a warm hit at
//    (unitig, pos + 1) is an occurrence of the k-mer, and there is only one -- so the launcher serves streaming queries with
//    the random-access kernel.
//  * Duplicates: both the position at hand and the position the lookup returned get their unitig line flagged (ULINE_DUP):
//    every occurrence of a duplicated k-mer then lies in a flagged line, and the walk kernel settles cursors only in groups
//    that touch one.
__global__ void __launch_bounds__(256) flag_duplicated_kmers_kernel(const __grid_constant__ IndexView ix, UnitigLine* __restrict__ lines,
                                                                    unsigned long long* n_elsewhere) {
  const u32 k = ix.unitigs.k;
  const u64 total = ix.unitigs.total_len;
  u32 bad = 0;
  for (u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x; p + k <= total; p += (u64)gridDim.x * blockDim.x) {
    u64 id, s, e;
    unitig_locate(ix.unitigs, p, id, s, e);
    if (p + k > e) continue;
    const u64 fw = useq_window(ix.unitigs, p), rc = revcomp(fw, k);
    Hit h;
    const bool found = k2u_any(ix, fw, rc, h);
    if (!found || h.unitig_id != (u32)id || h.pos != (u32)(p - s)) {
      ++bad;
      atomicOr(&lines[p >> ULINE_SHIFT].end_delta, ULINE_DUP);
      if (found) atomicOr(&lines[(ix.unitigs.starts[h.unitig_id] + h.pos) >> ULINE_SHIFT].end_delta, ULINE_DUP);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(n_elsewhere, (unsigned long long)bad);
}

// ---------------------------------------------------------------------------------------------
// Roofline probe: independent random reads of aligned granules (16 / 32 / 64 / 128 bytes), swept over granule size,
// loads in flight per thread and resident CTAs per SM by profiles/measure_prand.py.  A granule is read by LANES adjacent
// lanes with one 16-byte load each (LANES = 8: a warp instruction touches four random 128-byte lines).
// ---------------------------------------------------------------------------------------------
template <int LANES, int ILP>
__global__ void __launch_bounds__(256) gather_probe_kernel(const uint4* __restrict__ table, u64 n_granules, u64 n_items, u64 seed,
                                                           unsigned long long* __restrict__ sink) {
  const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x, nt = (u64)gridDim.x * blockDim.x;
  const u64 group = tid / LANES, n_groups = nt / LANES;
  const u32 sub = (u32)(tid % LANES);
  u64 acc = 0;
  for (u64 i = group * ILP; i < n_items; i += n_groups * ILP) {
    uint4 v[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      const u64 g = mulhi64(fmix64((i + j) * 0x9E3779B97F4A7C15ULL + seed), n_granules);
      v[j] = __ldg(table + g * LANES + sub);
    }
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc += v[j].x + v[j].w;
  }
  if (acc == 0x1234567887654321ULL) atomicAdd(sink, 1ULL);
}

}  // namespace mazu
