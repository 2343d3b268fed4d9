// 2-bit k-mer arithmetic shared by host builders and device kernels.
//
// Encoding (the `kmers` crate the reference builds on): A=0 C=1 G=2 T=3,
// base i of a k-mer at bits [2i, 2i+2) of a u64, canonical form = min(fw, rc) as integers.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define MZ_HD __host__ __device__ __forceinline__
#else
#define MZ_HD inline
#endif

namespace mazu {

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint8_t u8;

enum : u32 { NO_MATCH = 0, IDENTITY_MATCH = 1, TWIN_MATCH = 2, SKIPPED = 3 };

MZ_HD u64 kmer_mask(u32 k) { return k >= 32 ? ~0ULL : ((1ULL << (2 * k)) - 1ULL); }

// ASCII -> 2-bit code, or 4 for anything that is not ACGTacgt
MZ_HD u32 base_code(u32 c) {
  c &= 0xDFu;  // fold case
  u32 code = (c >> 1) & 3u;  // A(0x41)->0  C(0x43)->1  G(0x47)->3  T(0x54)->2
  code ^= code >> 1;         // swap 2<->3 : A0 C1 G2 T3
  bool ok = (c == 'A') | (c == 'C') | (c == 'G') | (c == 'T');
  return ok ? code : 4u;
}

// Eight ASCII bases at once on the host (v little-endian: base j in byte j): their 2-bit codes in 16 bits (0 where the base is
// not ACGTacgt) and one invalid flag per base -- base_code() on eight bytes with 64-bit arithmetic (mazu_b200_pack_reads).
inline void pack8_ascii(u64 v, u32& codes16, u32& inv8) {
  const u64 ONES = 0x0101010101010101ULL, LOW7 = 0x7F7F7F7F7F7F7F7FULL, HI = 0x8080808080808080ULL;
  const u64 u = v & 0xDFDFDFDFDFDFDFDFULL;  // fold case
  // 0x80 in every byte of u that differs from `letter` (exact: no carry leaves a byte)
#define MZ_NE8(letter) (((((u ^ (ONES * (u64)(letter))) & LOW7) + LOW7) | (u ^ (ONES * (u64)(letter)))) & HI)
  const u64 bad = (MZ_NE8('A') & MZ_NE8('C') & MZ_NE8('G') & MZ_NE8('T')) >> 7;  // bit 0 of every byte that is none of the four
#undef MZ_NE8
  u64 t = (u >> 1) & (ONES * 3);  // A->0 C->1 G->3 T->2
  t ^= (t >> 1) & ONES;           // A0 C1 G2 T3
  t &= ~(bad * 3);
  t = (t | (t >> 6)) & 0x000F000F000F000FULL;  // gather the eight 2-bit fields
  t = (t | (t >> 12)) & 0x000000FF000000FFULL;
  t = (t | (t >> 24)) & 0xFFFFULL;
  codes16 = (u32)t;
  inv8 = (u32)((bad * 0x0102040810204080ULL) >> 56);  // bit 0 of byte j -> bit j
}

// reverse complement of the low 2k bits
MZ_HD u64 revcomp(u64 x, u32 k) {
  x = ~x;  // complement: 3 - b
#if defined(__CUDA_ARCH__)
  x = __brevll(x);
#else
  x = ((x >> 32) | (x << 32));
  x = ((x & 0xFFFF0000FFFF0000ULL) >> 16) | ((x & 0x0000FFFF0000FFFFULL) << 16);
  x = ((x & 0xFF00FF00FF00FF00ULL) >> 8) | ((x & 0x00FF00FF00FF00FFULL) << 8);
  x = ((x & 0xF0F0F0F0F0F0F0F0ULL) >> 4) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
  x = ((x & 0xCCCCCCCCCCCCCCCCULL) >> 2) | ((x & 0x3333333333333333ULL) << 2);
  x = ((x & 0xAAAAAAAAAAAAAAAAULL) >> 1) | ((x & 0x5555555555555555ULL) << 1);
#endif
  // full bit reversal also reversed the two bits inside each base: swap them back
  x = ((x & 0xAAAAAAAAAAAAAAAAULL) >> 1) | ((x & 0x5555555555555555ULL) << 1);
  return k >= 32 ? x : (x >> (64 - 2 * k));
}

MZ_HD u64 mulhi64(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}

// Minimizer order v3 (DESIGN.md "Minimizer order"): a seeded 32-bit multiplicative mixer of the
// STRAND-SYMMETRIC w-mer word min(wmer, revcomp(wmer)).  The w-mer at offset ci (0..k-w, k-w <= 31) of the
// CANONICAL k-mer gets the key
//     (mm_hash32(min(wmer, rc(wmer))) & 0xFFFFFFE0) | ci
// the w-mer with the smallest key wins (top 27 hash bits, ties -> leftmost in the canonical k-mer), and the
// minimizer WORD is min(wmer, rc(wmer)).  A k-mer and its neighbour therefore keep their minimizer when the
// canonical strand flips between them (it flips every other k-mer on average): a 150 bp read has ~18
// super-k-mers instead of ~60 under v2, which hashed the w-mer as read off the canonical strand.
// Stands in for kmers::canonical_minimizer + wyhash 0.5.0 (neither is in the reference tree, so
// no hash could be pinned); k-mer -> unitig results do not depend on it.
MZ_HD u32 mm_hash32(u64 x, u64 seed) {
  u32 h = ((u32)x ^ (u32)seed) * 0x85EBCA6Bu;
  h ^= ((u32)(x >> 32) ^ (u32)(seed >> 32)) * 0xC2B2AE35u;
  h ^= h >> 16;
  h *= 0x7FEB352Du;
  h ^= h >> 15;
  h *= 0x846CA68Bu;
  h ^= h >> 16;
  return h;
}
static const u32 MM_KEY_MASK = 0xFFFFFFE0u;

// murmur3 finaliser: the per-level hash of the native MPHF ("kphf" tables)
MZ_HD u64 fmix64(u64 x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

// pufferfish BooPHF hashing (src/pf1/boophf/hash.rs:33-49,111-135)
MZ_HD u64 boophf_hash64(u64 key, u64 seed) {
  u64 hash = seed;
  hash ^= (hash << 7) ^ (key * (hash >> 3)) ^ (~((hash << 11) + (key ^ (hash >> 5))));
  hash = (~hash) + (hash << 21);
  hash = hash ^ (hash >> 24);
  hash = (hash + (hash << 3)) + (hash << 8);
  hash = hash ^ (hash >> 14);
  hash = (hash + (hash << 2)) + (hash << 4);
  hash = hash ^ (hash >> 28);
  hash = hash + (hash << 31);
  return hash;
}
static const u64 BOOPHF_SEED0 = 0xAAAAAAAA55555555ULL;
static const u64 BOOPHF_SEED1 = 0x33333333CCCCCCCCULL;

// minimizer of a k-mer given as fw/rc words: argmin over the canonical k-mer's w-mers, leftmost
// wins; `offset` is reported in the coordinates of the fw k-mer (sshash.rs:563-624 pins this).
struct MinimizerResult {
  u64 word;
  u32 offset;
};
MZ_HD u32 mm_key(u64 wf, u64 wr, u64 seed) { return mm_hash32(wf <= wr ? wf : wr, seed) & MM_KEY_MASK; }
MZ_HD MinimizerResult canonical_minimizer_naive(u64 fw, u64 rc, u32 k, u32 w, u64 seed) {
  const bool fw_canon = fw <= rc;
  const u64 c = fw_canon ? fw : rc, d = fw_canon ? rc : fw;  // canonical strand, other strand
  const u64 wmask = kmer_mask(w);
  const u32 span = k - w;
  u32 best = 0xFFFFFFFFu;
  for (u32 i = 0; i <= span; ++i) {  // the w-mer at offset i of c is the reverse complement of the one at offset span - i of d
    u32 key = mm_key((c >> (2 * i)) & wmask, (d >> (2 * (span - i))) & wmask, seed) | i;
    best = key < best ? key : best;
  }
  const u32 best_i = best & 31u;
  const u64 a = (c >> (2 * best_i)) & wmask, b = (d >> (2 * (span - best_i))) & wmask;
  MinimizerResult r;
  r.word = a <= b ? a : b;
  r.offset = fw_canon ? best_i : (span - best_i);
  return r;
}

}  // namespace mazu
