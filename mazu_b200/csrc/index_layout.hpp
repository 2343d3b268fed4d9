// Device-resident index layout (what lives in HBM) and the lookup primitives over it.
//
// Everything here is laid out for 32-byte DRAM sectors: each logically distinct step of the
// reference's dependent chain (MPHF level probe, bucket bounds, bucket positions, k-mer window,
// unitig id/bounds) touches ONE aligned 32-byte block wherever possible.  See DESIGN.md
// "Data layout in HBM".  The structs hold raw pointers; they are filled by DeviceIndex (device
// pointers) for kernels, and by the host builders (host pointers) where a builder needs to
// evaluate its own MPHF.
#pragma once
#include "kmer.hpp"

namespace mazu {

#if defined(__CUDA_ARCH__)
#define MZ_POPC(x) __popc(x)
#define MZ_POPCLL(x) __popcll(x)
#define MZ_LDG(p) __ldg(p)
#else
#define MZ_POPC(x) __builtin_popcount(x)
#define MZ_POPCLL(x) __builtin_popcountll(x)
#define MZ_LDG(p) (*(p))
#endif

static const u32 MPHF_MAX_LEVELS = 32;
static const u32 MPHF_BLOCK_BITS = 224;  // payload bits per 32-byte block (7 x u32) + 1 x u32 rank
enum : u32 { MPHF_FAMILY_BOOPHF = 0, MPHF_FAMILY_NATIVE = 1, MPHF_FAMILY_CASCADE = 2 };

// ---------------------------------------------------------------------------------------------
// Ranked bitset blocks: the storage of both MPHF flavours.
//   block (32 B, aligned) = { u32 ones_before_block_in_level ; u32 bits[7] }   (224 slots)
// A level probe reads the slot's u32; on a set bit it reads the block's other words to get the
// rank -- one sector per probed level.  The reference BooPHF keeps bits and 512-bit rank samples
// in separate arrays (src/pf1/boophf/mod.rs:250-266: 1 bit word + 1 sample + <=8 words).
// ---------------------------------------------------------------------------------------------
struct RankedLevels {
  const u32* blocks;
  u32 family;
  u32 n_levels;
  u64 n_keys;
  u64 size[MPHF_MAX_LEVELS];       // BOOPHF: n_bits of the level (fastrange modulus); NATIVE: n_blocks; CASCADE: slots of the level
  u64 block_off[MPHF_MAX_LEVELS];  // first block of the level (CASCADE: first global slot of the level)
  u64 rank_base[MPHF_MAX_LEVELS];  // keys placed in earlier levels
  const u64* fb_keys;              // fallback ("final hash"): sorted keys ...
  const u64* fb_vals;              // ... and their hash values
  u32 n_fb;
  u32 _pad;
};

MZ_HD u32 mulhi32(u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (u32)(((u64)a * b) >> 32);
#endif
}

// slot -> (block, bit) for the native family.  One 64-bit mix of the key per lookup, then a cheap
// 32-bit remix per level (the level loop is the divergent part of a lookup, so it is kept short).
MZ_HD void native_slot_from(u64 h, u32 level, u64 n_blocks, u64& blk, u32& bit) {
  u32 a = (u32)(h >> 32) + level * 0x9E3779B1u;
  u32 b = (u32)h ^ (level * 0x85EBCA77u);
  a ^= b;
  a *= 0x2C1B3C6Du;
  a ^= a >> 15;
  b += a;
  b *= 0x297A2D39u;
  b ^= b >> 16;
  blk = mulhi32(a, (u32)n_blocks);
  bit = mulhi32(b, MPHF_BLOCK_BITS);
}
MZ_HD void native_slot(u64 key, u32 level, u64 n_blocks, u64& blk, u32& bit) {
  native_slot_from(fmix64(key), level, n_blocks, blk, bit);
}

// bit test of one slot: reads a single u32 of the block
MZ_HD bool ranked_test(const RankedLevels& m, u32 level, u64 blk, u32 bit) {
  const u32* b = m.blocks + (m.block_off[level] + blk) * 8;
  return (MZ_LDG(b + 1 + (bit >> 5)) >> (bit & 31)) & 1u;
}
// rank of a set slot: block header + popcounts inside the same 32-byte block
MZ_HD u64 ranked_rank(const RankedLevels& m, u32 level, u64 blk, u32 bit) {
  const u32* b = m.blocks + (m.block_off[level] + blk) * 8;
#if defined(__CUDA_ARCH__)
  const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(b));
  const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(b) + 1);
  const u32 w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#else
  const u32* w = b;
#endif
  const u32 wi = 1 + (bit >> 5), sh = bit & 31;
  u32 r = w[0];
#pragma unroll
  for (u32 j = 1; j < 8; ++j) {
    u32 x = w[j];
    if (j == wi) x &= (1u << sh) - 1u;
    if (j <= wi) r += MZ_POPC(x);
  }
  return m.rank_base[level] + r;
}

// Load factor of level `l` of the native cascade.  The first two levels use gamma (2 slots per key) and place
// ~85 % of the keys; after that gamma escalates (x4, then x32) so the few keys left are placed almost without
// collisions and the cascade ends after 5-6 levels instead of ~9 (yeast minimizers) / ~20 (human scale).  The
// level loop of a warp runs until its slowest lane is done, and a NON-member key walks all levels unless it
// meets a set bit (30 % per level at gamma = 2), so the depth of the cascade -- not the average level of a
// member -- is what a warp of mixed lookups pays for.  Costs ~5.2 instead of 3.3 bits per key.
MZ_HD double native_level_gamma(double gamma, u32 level) { return level < 2 ? gamma : (level == 2 ? 4.0 * gamma : 32.0 * gamma); }

// fast_range_64 (mod.rs:136-144): slot = (h * n_bits) >> 64, then (block, bit) in the re-blocked level.  Levels of fewer than
// 2^32 slots (every pufferfish index below 4e9 k-mers, and all fixtures) take two 32x32 multiplies and a 32-bit division
// instead of the 64-bit ones.
MZ_HD void boophf_slot(u64 h, u64 nb, u64& blk, u32& bit) {
  if ((nb >> 32) == 0) {
    const u64 hi = (h >> 32) * nb, lo = (h & 0xFFFFFFFFULL) * nb;
    const u32 pos = (u32)((hi + (lo >> 32)) >> 32);
    const u32 b32 = pos / MPHF_BLOCK_BITS;
    blk = b32;
    bit = pos - b32 * MPHF_BLOCK_BITS;
  } else {
    const u64 pos = mulhi64(h, nb);
    blk = pos / MPHF_BLOCK_BITS;
    bit = (u32)(pos - blk * MPHF_BLOCK_BITS);
  }
}

// MPHF::try_hash_u64 (src/kphf/mod.rs:54-56).  BOOPHF reproduces BooPHF<u64>::lookup
// (src/pf1/boophf/mod.rs:96-181) bit-exactly; NATIVE is this library's own BBHash-style MPHF.
// Like boomphf::try_hash, a non-member key may return a false-positive value < n_keys.
// Templated on the family so a kernel only carries the code of the family it serves.  The level
// loop only tests bits (divergent lanes leave at different levels); the rank is taken once after it.
template <u32 FAMILY>
MZ_HD bool mphf_lookup_t(const RankedLevels& m, u64 key, u64& out) {
  u32 hit_level = MPHF_MAX_LEVELS, hit_bit = 0;
  u64 hit_blk = 0;
  if (FAMILY == MPHF_FAMILY_NATIVE) {
    const u64 hk = fmix64(key);
#pragma unroll 1
    for (u32 l = 0; l < m.n_levels; ++l) {
      u64 blk;
      u32 bit;
      native_slot_from(hk, l, m.size[l], blk, bit);
      if (ranked_test(m, l, blk, bit)) {
        hit_level = l;
        hit_blk = blk;
        hit_bit = bit;
        break;
      }
    }
  } else {
    u64 s0 = BOOPHF_SEED0, s1 = BOOPHF_SEED1;  // MultiHashState (src/pf1/boophf/hash.rs:91-97)
#pragma unroll 1
    for (u32 l = 0; l < m.n_levels; ++l) {
      u64 h;
      if (l == 0) {
        h = boophf_hash64(key, BOOPHF_SEED0);
        s0 = h;
      } else if (l == 1) {
        h = boophf_hash64(key, BOOPHF_SEED1);
        s1 = h;
      } else {
        u64 a = s0, b = s1;
        a ^= a << 23;
        a = a ^ b ^ (a >> 17) ^ (b >> 26);
        h = a + b;
        s0 = b;
        s1 = a;
      }
      u64 blk;
      u32 bit;
      boophf_slot(h, m.size[l], blk, bit);
      if (ranked_test(m, l, blk, bit)) {
        hit_level = l;
        hit_blk = blk;
        hit_bit = bit;
        break;
      }
    }
  }
  if (hit_level < MPHF_MAX_LEVELS) {
    out = ranked_rank(m, hit_level, hit_blk, hit_bit);
    return true;
  }
  // lookup_in_final_hash (mod.rs:177-181) / native leftovers
  u32 lo = 0, hi = m.n_fb;
  while (lo < hi) {
    u32 mid = (lo + hi) >> 1;
    u64 kk = MZ_LDG(m.fb_keys + mid);
    if (kk == key) {
      out = MZ_LDG(m.fb_vals + mid);
      return true;
    }
    if (kk < key) lo = mid + 1; else hi = mid;
  }
  return false;
}
// lookup_in_final_hash (mod.rs:177-181) / native leftovers
MZ_HD bool mphf_fallback(const RankedLevels& m, u64 key, u64& out) {
  u32 lo = 0, hi = m.n_fb;
  while (lo < hi) {
    u32 mid = (lo + hi) >> 1;
    u64 kk = MZ_LDG(m.fb_keys + mid);
    if (kk == key) {
      out = MZ_LDG(m.fb_vals + mid);
      return true;
    }
    if (kk < key) lo = mid + 1; else hi = mid;
  }
  return false;
}

#if defined(__CUDACC__)
// The level loop of mphf_lookup_t for SEVERAL keys per lane, the lane refilling itself: a lane whose key is placed moves on to
// its next key in the same loop iteration count as its neighbours keep probing levels, so the warp's loop runs
// max-over-lanes(sum of levels of the lane's keys) times instead of (number of keys) x max-over-lanes(levels of one key) --
// 61-69 % of the lanes busy per iteration for 4-8 keys per lane, against 28 % for one key per lane (a level places ~61 % of
// its keys, the slowest of 32 lanes needs ~6 levels).
// In : a[j * stride] (and b[j * stride] for BOOPHF) = the level-0 (level-1) hash of the lane's key j < n_mine -- BOOPHF:
//      boophf_hash64(key, SEED0 / SEED1), the MultiHashState of src/pf1/boophf/hash.rs:91-97; NATIVE: fmix64(key).
// Out: a[j * stride] = MPHF_MULTI_NONE, or level << 56 | block << 8 | bit of the set slot (rank it with mphf_multi_rank).
// The arrays live in shared memory (they are indexed by the lane's own progress).
static const u64 MPHF_MULTI_NONE = ~0ULL;
template <u32 FAMILY>
__device__ __forceinline__ void mphf_levels_multi(const RankedLevels& m, u64* a, u64* b, u32 stride, u32 n_mine) {
  if (n_mine == 0) return;
  u32 j = 0, l = 0;
  u64 s0 = a[0], s1 = FAMILY == MPHF_FAMILY_NATIVE ? 0ULL : b[0];
  const u32 n_levels = m.n_levels;
#pragma unroll 1
  while (j < n_mine) {
    u64 blk;
    u32 bit;
    if (FAMILY == MPHF_FAMILY_NATIVE) {
      native_slot_from(s0, l, m.size[l], blk, bit);
    } else {
      u64 h = l == 0 ? s0 : s1;
      if (l >= 2) {  // xorshift128+ step of the MultiHashState (hash.rs:111-135)
        u64 x = s0, y = s1;
        x ^= x << 23;
        x = x ^ y ^ (x >> 17) ^ (y >> 26);
        h = x + y;
        s0 = y;
        s1 = x;
      }
      boophf_slot(h, m.size[l], blk, bit);
    }
    const bool hit = ranked_test(m, l, blk, bit);
    if (hit || l + 1 >= n_levels) {
      a[j * stride] = hit ? ((u64)l << 56) | (blk << 8) | bit : MPHF_MULTI_NONE;
      ++j;
      l = 0;
      if (j < n_mine) {
        s0 = a[j * stride];
        if (FAMILY != MPHF_FAMILY_NATIVE) s1 = b[j * stride];
      }
    } else {
      ++l;
    }
  }
}
__device__ __forceinline__ bool mphf_multi_rank(const RankedLevels& m, u64 packed, u64 key, u64& out) {
  if (packed == MPHF_MULTI_NONE) return mphf_fallback(m, key, out);
  out = ranked_rank(m, (u32)(packed >> 56), (packed >> 8) & 0xFFFFFFFFFFFFULL, (u32)(packed & 255u));
  return true;
}
#endif

MZ_HD bool mphf_lookup(const RankedLevels& m, u64 key, u64& out) {
  return m.family == MPHF_FAMILY_NATIVE ? mphf_lookup_t<MPHF_FAMILY_NATIVE>(m, key, out) : mphf_lookup_t<MPHF_FAMILY_BOOPHF>(m, key, out);
}

// ---------------------------------------------------------------------------------------------
// Packed integer vector (simple-sds IntVector / pufferfish compact vector), LSB-first.
// ---------------------------------------------------------------------------------------------
struct PackedVecView {
  const u64* words;  // padded by >= 1 word
  u64 len;
  u32 width;
  u32 _pad;
};
MZ_HD u64 packed_get(const PackedVecView& v, u64 i) {
  u64 bit = i * v.width, wi = bit >> 6;
  u32 sh = (u32)(bit & 63);
  u64 x = MZ_LDG(v.words + wi) >> sh;
  if (sh + v.width > 64) x |= MZ_LDG(v.words + wi + 1) << (64 - sh);
  return v.width >= 64 ? x : (x & ((1ULL << v.width) - 1ULL));
}

// ---------------------------------------------------------------------------------------------
// Blocked Elias-Fano: EFVector (src/elias_fano.rs) re-partitioned so that get(i) and get(i+1)
// -- the bucket bounds of SSHash::k2u (src/kphf/sshash.rs:482-483) -- come from ONE 32-byte block.
//   block b covers elements [b*S, b*S + S] (S+1 elements, neighbouring blocks overlap by one):
//     word0          : value of element b*S (the block base); bit 63 set => exception block,
//                      then bits 0..62 = index of S+1 plain u64 values in `exceptions`
//     word1, word2   : upper-bits bucket in negated-unary form, relative to (base >> l):
//                      element j has its bit at position ((x_j >> l) - (x_0 >> l)) + j   (<128)
//     word3          : lower bits, element j at bits [j*l, (j+1)*l)                      (<=64)
//   l follows the reference rule l = max(1, msb(u / n)) (elias_fano.rs:63-75).
// ---------------------------------------------------------------------------------------------
struct BlockedEFView {
  const u64* blocks;      // wpb words per block: 4 (Elias-Fano only) or 8 (+ one fingerprint byte per element, words 4..7)
  const u64* exceptions;  // (S+1) words per exception block
  u64 n;                  // number of elements
  u32 l;
  u32 log_s;
  u32 wpb;
  u32 _pad;
};

// the state / fingerprint byte of element i (words 4..7 of its block; see the fingerprinted cascade below)
MZ_HD u32 blocked_ef_fp(const BlockedEFView& ef, u64 i) {
  u64 blk = i >> ef.log_s;
  u32 j = (u32)(i & ((1ULL << ef.log_s) - 1ULL));
  return (u32)((MZ_LDG(ef.blocks + blk * ef.wpb + 4 + (j >> 3)) >> (8 * (j & 7))) & 0xFFULL);
}

// ---------------------------------------------------------------------------------------------
// Fingerprinted cascade (MPHF_FAMILY_CASCADE): the perfect hash of the SSHash minimizers.
// A BBHash-style cascade whose VALUE is the slot itself (level offset + slot in level) instead of the slot's rank, so it
// needs no bit tables and no rank: the only per-slot state is one byte, and that byte is the fingerprint byte the bucket
// bounds block of the slot already carries (words 4..7 of the 64-byte blocked-Elias-Fano block):
//     0 EMPTY      no key of the level maps here        -> a queried key is provably not a member
//     1 COLLIDED   two or more keys map here            -> they were all sent to the next level: go on
//     2..255       exactly one key: its fingerprint     -> equal: found, the bounds are in the block just read; else not a member
// A member pays 1.28 probes on average (each the 128-byte line that also holds its bucket bounds), a non-member 1.03 --
// the reference's chain (MPHF levels, then Elias-Fano bounds, src/kphf/sshash.rs:478-483) pays both and learns that a
// minimizer is foreign only at the k-mer compare.  The hash is not minimal: the bounds array has one (possibly empty)
// bucket per slot, ~5.3 slots per key at 2 bytes each.  Leftover keys (after MPHF_MAX_LEVELS levels) sit in the sorted
// fallback list with slots behind the last level.
// ---------------------------------------------------------------------------------------------
static const u32 CASCADE_EMPTY = 0, CASCADE_COLLIDED = 1;
MZ_HD u32 cascade_fp(u64 hk) { return 2u + mulhi32((u32)(hk ^ (hk >> 29)) * 0x9E3779B1u, 254u); }
MZ_HD u64 cascade_slot(u64 hk, u32 level, u64 size) {
  if (size >> 32) {  // levels of more than 2^32 slots (> 2e9 minimizers): 64-bit arithmetic
    u64 x = (hk ^ ((u64)(level + 1) * 0x9E3779B97F4A7C15ULL)) * 0xD6E8FEB86659FD93ULL;
    x ^= x >> 32;
    return mulhi64(x * 0xFF51AFD7ED558CCDULL, size);
  }
  // hk is already a full 64-bit mix of the key: a cheap 32-bit remix per level picks the slot (32 bits of hash are enough to
  // address < 2^32 slots uniformly; the level loop is the divergent part of a lookup, so it is kept short)
  u32 a = (u32)(hk >> 32) + level * 0x9E3779B1u;
  const u32 b = (u32)hk ^ (level * 0x85EBCA77u);
  a ^= b;
  a *= 0x2C1B3C6Du;
  a ^= a >> 15;
  a *= 0x297A2D39u;
  a ^= a >> 16;
  return mulhi32(a, (u32)size);
}
// slots of level `l` for `n` keys: 4 slots per key on the first two levels (78 % of the keys of a level are alone in their
// slot, a member pays 1.28 probes on average, and 78 % of the level-0 slots are empty so a non-member usually stops at once),
// then 8.  Measured against 2/2/4/8 (profiles/experiments/README.md, round 2): config 5 reads +3.4 %, flat k-mers +8.8 %,
// for 1.2 GB more index at human scale (5.3 slots of 2 bytes per key instead of 3.7).
MZ_HD u64 cascade_level_size(u64 n, u32 level) {
#ifndef MAZU_CASCADE_G0
#define MAZU_CASCADE_G0 4
#define MAZU_CASCADE_G1 4
#define MAZU_CASCADE_G2 8
#endif
  const u64 g = level == 0 ? MAZU_CASCADE_G0 : (level == 1 ? MAZU_CASCADE_G1 : (level == 2 ? MAZU_CASCADE_G2 : 8));
  const u64 s = g * n;
  return s < 32 ? 32 : s;
}
MZ_HD bool cascade_lookup(const RankedLevels& m, const BlockedEFView& ef, u64 key, u64& out) {
  const u64 hk = fmix64(key);
  const u32 fp = cascade_fp(hk);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (u32 l = 0; l < m.n_levels; ++l) {
    const u64 g = m.block_off[l] + cascade_slot(hk, l, m.size[l]);
    const u32 st = blocked_ef_fp(ef, g);
    if (st == fp) {
      out = g;
      return true;
    }
    if (st != CASCADE_COLLIDED) return false;
  }
  u32 lo = 0, hi = m.n_fb;
  while (lo < hi) {
    u32 mid = (lo + hi) >> 1;
    u64 kk = MZ_LDG(m.fb_keys + mid);
    if (kk == key) {
      out = MZ_LDG(m.fb_vals + mid);
      return true;
    }
    if (kk < key) lo = mid + 1; else hi = mid;
  }
  return false;
}

// position of the j-th (0-based) set bit of the 128-bit value hi:lo; caller guarantees it exists
MZ_HD u32 select128(u64 lo, u64 hi, u32 j) {
  u32 c = (u32)MZ_POPCLL(lo);
  u64 w = lo;
  u32 base = 0;
  if (j >= c) {
    j -= c;
    w = hi;
    base = 64;
  }
  u32 wl = (u32)w, wh = (u32)(w >> 32);
  u32 cl = (u32)MZ_POPC(wl);
  u32 x = wl;
  if (j >= cl) {
    j -= cl;
    x = wh;
    base += 32;
  }
  // j-th set bit of the 32-bit word x by halving (popcount binary search; __fns is a 32-step loop)
  u32 pos = 0;
#pragma unroll
  for (u32 width = 16; width >= 1; width >>= 1) {
    u32 c2 = (u32)MZ_POPC((x >> pos) & ((1u << width) - 1u));
    if (j >= c2) {
      j -= c2;
      pos += width;
    }
  }
  return base + pos;
}

// returns x_i in `a` and x_{i+1} in `b`; requires i + 1 < n
MZ_HD void blocked_ef_get2(const BlockedEFView& ef, u64 i, u64& a, u64& b) {
  u64 blk = i >> ef.log_s;
  u32 j = (u32)(i & ((1ULL << ef.log_s) - 1ULL));
  const u64* p = ef.blocks + blk * ef.wpb;
#if defined(__CUDA_ARCH__)
  const ulonglong2 q0 = __ldg(reinterpret_cast<const ulonglong2*>(p));
  const ulonglong2 q1 = __ldg(reinterpret_cast<const ulonglong2*>(p) + 1);
  u64 w0 = q0.x, w1 = q0.y, w2 = q1.x, w3 = q1.y;
#else
  u64 w0 = p[0], w1 = p[1], w2 = p[2], w3 = p[3];
#endif
  if (w0 >> 63) {
    const u64* e = ef.exceptions + (w0 & ~(1ULL << 63)) * ((1ULL << ef.log_s) + 1ULL);
    a = MZ_LDG(e + j);
    b = MZ_LDG(e + j + 1);
    return;
  }
  u32 pj = select128(w1, w2, j);
  // next set bit after pj
  u32 pn;
  if (pj < 63) {
    u64 rest = w1 & (~0ULL << (pj + 1));
#if defined(__CUDA_ARCH__)
    pn = rest ? (u32)(__ffsll((long long)rest) - 1) : 64u + (u32)(__ffsll((long long)w2) - 1);
#else
    pn = rest ? (u32)__builtin_ctzll(rest) : 64u + (u32)__builtin_ctzll(w2);
#endif
  } else {
    u64 rest = pj == 63 ? w2 : (pj == 127 ? 0ULL : (w2 & (~0ULL << (pj - 63))));
#if defined(__CUDA_ARCH__)
    pn = 64u + (u32)(__ffsll((long long)rest) - 1);
#else
    pn = 64u + (u32)__builtin_ctzll(rest);
#endif
  }
  u64 lmask = (1ULL << ef.l) - 1ULL;
  u64 hb = w0 >> ef.l;
  a = ((hb + (u64)(pj - j)) << ef.l) | ((w3 >> (j * ef.l)) & lmask);
  b = ((hb + (u64)(pn - (j + 1))) << ef.l) | ((w3 >> ((j + 1) * ef.l)) & lmask);
}

// ---------------------------------------------------------------------------------------------
// UnitigSet (src/unitig_set.rs:31-36) on the device.
//   useq    2-bit packed concatenated unitigs (+2 pad words: get_kmer_u64 may read past the end)
//   dir     dir[p >> dir_shift] = id of the unitig containing position (p >> dir_shift) << dir_shift
//   starts  n_unitigs+1 plain u64 prefix lengths
// locate(pos) replaces the reference's rank over an L-bit end-marker vector plus 3-4 Elias-Fano
// get()s (unitig_set.rs:178-209) with one directory sector + one or two `starts` sectors.
// ---------------------------------------------------------------------------------------------
// Unitig line: everything a verified candidate needs, in ONE 128-byte DRAM line (the granularity a random access
// pays for on this part: profiles/r02_prand.json).  Line i covers bases [256 i, 256 i + 256), base = 256 i:
//   seq[9]       the line's 256 bases + the 32 bases that follow, so a k-mer window (k <= 32) never straddles lines
//   first_id     unitig containing `base`                start_delta  base - start of that unitig
//   end_delta    (end of the unitig containing base + 255) - base, in the low 31 bits; bit 31 = DUP flag: some k-mer that starts
//                in this line occurs more than once in the unitig set (set at creation, flag_duplicated_kmers_kernel).  A
//                streaming answer can differ from K2U::k2u only for such k-mers, so the walk kernel settles cursors only there
//   ends[4]      bit j <=> base + j is the last base of a unitig (the reference's end-marker bit-vector,
//                unitig_set.rs:146-149, cut into the line it belongs to)
//   per 64-base word w of `ends` (one byte each, packed in a u32):
//     cnt_before   markers in words < w
//     prev_end     offset of the last marker in words < w   (0xFF: none -> the unitig of `base`)
//     next_end     offset of the first marker in words > w  (0: none -> end_delta)
//   so a lookup reads ONE word of `ends`, whatever the offset.
// get_kmer_u64_from_useq_pos + pos_to_id + unitig_start_pos + unitig_end_pos (unitig_set.rs:185-229) = this one line;
// the flat arrays above cost three (window, directory, starts).  Built on the device from useq / dir / starts
// (build_unitig_lines_kernel); the flat arrays stay for the builders, the query kernels only touch lines.
static const u32 ULINE_SHIFT = 8;
struct alignas(128) UnitigLine {
  u64 seq[9];
  u32 first_id;
  u32 start_delta;
  u32 end_delta;
  u32 cnt_before;
  u32 prev_end;
  u32 next_end;
  u64 ends[4];
};
static_assert(sizeof(UnitigLine) == 128, "UnitigLine must be one 128-byte line");

struct UnitigsView {
  const u64* useq;
  const u32* dir;
  const u64* starts;
  const UnitigLine* lines;
  u64 total_len;
  u64 n_unitigs;
  u32 k;
  u32 dir_shift;
};
// Compile-time (k, w) of a read-kernel instantiation: KW = MZ_KW(k, w), or 0 = take them from the view.  The read kernels are
// issue-bound and k, w feed every mask, shift and window bound of stages M, B and V; with the two (k, w) pairs the named
// configurations use folded into the code the SSHash read kernel issues 11 % fewer instructions per lookup (profiles/experiments).
#define MZ_KW(k, w) (((u32)(k) << 8) | (u32)(w))
template <u32 KW> MZ_HD u32 kw_k(u32 runtime_k) { return KW ? (KW >> 8) : runtime_k; }
template <u32 KW> MZ_HD u32 kw_w(u32 runtime_w) { return KW ? (KW & 255u) : runtime_w; }
// SeqVector::get_kmer_u64(pos, k): 2k bits at bit 2*pos (unitig_set.rs:226-229)
MZ_HD u64 useq_window(const UnitigsView& u, u64 pos) {
  u64 bit = 2 * pos, wi = bit >> 6;
  u32 sh = (u32)(bit & 63);
  u64 x = MZ_LDG(u.useq + wi) >> sh;
  if (sh + 2 * u.k > 64) x |= MZ_LDG(u.useq + wi + 1) << (64 - sh);
  return x & kmer_mask(u.k);
}
// pos_to_id + unitig_start_pos + unitig_end_pos (unitig_set.rs:185-204)
MZ_HD void unitig_locate(const UnitigsView& u, u64 pos, u64& id, u64& start, u64& end) {
  u64 i = MZ_LDG(u.dir + (pos >> u.dir_shift));
  u64 e = MZ_LDG(u.starts + i + 1);
  while (e <= pos) {
    ++i;
    e = MZ_LDG(u.starts + i + 1);
  }
  id = i;
  start = MZ_LDG(u.starts + i);
  end = e;
}

#if defined(__CUDA_ARCH__) || defined(__CUDACC__)
// the same two primitives over unitig lines (query kernels): one DRAM line per verified candidate
template <u32 KW = 0>
__device__ __forceinline__ u64 line_window(const UnitigsView& u, u64 pos) {
  const u32 k = kw_k<KW>(u.k);
  const u64* ln = reinterpret_cast<const u64*>(u.lines + (pos >> ULINE_SHIFT));
  const u32 off = (u32)pos & 255u, wi = off >> 5, sh = 2 * (off & 31u);
  u64 x = __ldg(ln + wi) >> sh;
  if (sh + 2 * k > 64) x |= __ldg(ln + wi + 1) << (64 - sh);
  return x & kmer_mask(k);
}
static const u32 ULINE_DUP = 0x80000000u;
__device__ __forceinline__ void line_locate(const UnitigsView& u, u64 pos, u64& id, u64& start, u64& end, u32* dup = nullptr) {
  const UnitigLine* L = u.lines + (pos >> ULINE_SHIFT);
  const u32 off = (u32)pos & 255u, wq = off >> 6;
  const u64 e = __ldg(L->ends + wq);
  const uint2 a = __ldg(reinterpret_cast<const uint2*>(&L->first_id));  // first_id, start_delta
  const uint4 b = __ldg(reinterpret_cast<const uint4*>(&L->end_delta));  // end_delta, cnt_before, prev_end, next_end
  const u64 below = (1ULL << (off & 63u)) - 1ULL;
  const u64 lo = e & below, hi = e & ~below;  // markers before `off` / at or after `off` inside this word
  const u32 sh = 8 * wq;
  const u32 prev = (b.z >> sh) & 255u, next = (b.w >> sh) & 255u;
  const u64 base = pos & ~255ULL;
  id = (u64)a.x + ((b.y >> sh) & 255u) + (u32)__popcll(lo);
  if (lo) start = base + (u64)(64 * wq + 64 - __clzll((long long)lo));
  else start = prev != 255u ? base + prev + 1 : base - a.y;
  if (hi) end = base + (u64)(64 * wq + __ffsll((long long)hi));
  else end = next != 0u ? base + next + 1 : base + (b.x & ~ULINE_DUP);
  if (dup) *dup = b.x >> 31;
}
#endif

// ---------------------------------------------------------------------------------------------
// The whole index as the kernels see it (passed by value as a __grid_constant__ parameter).
// ---------------------------------------------------------------------------------------------
struct IndexView {
  u32 k2u_kind;  // MAZU_K2U_PFHASH / MAZU_K2U_SSHASH
  u32 u2pos_kind;
  UnitigsView unitigs;
  RankedLevels mphf;  // PFHash: over canonical k-mers; SSHash: over minimizers
  PackedVecView pos;  // PFHash: k-mer positions; SSHash: minimizer occurrence positions ("Offsets")
  // SSHash only
  BlockedEFView sizes;  // occs_prefix_sum ("Sizes"), n_minimizers + 1 elements
  u32 w;
  u32 has_skew;
  u64 seed;
  u64 skew_param;
  RankedLevels skew_mphf;
  PackedVecView skew_pos;
  // SampledPFHash only (src/kphf/pfhash.rs:139-151): pos is `sampled_pos`
  RankedLevels sampled;        // `sampled_vec` as one level of ranked blocks (bit test + rank in one sector)
  const u64* canonical_bits;   // `canonical_vec`
  const u64* direction_bits;   // `direction_vec`
  PackedVecView ext_sizes;
  PackedVecView ext_bases;
  u32 extension_size;
  u32 _pad_sampled;
  // U2Pos
  const u64* ctable_words;  // DENSE: one u64 per occurrence; PISCEM: packed `ctable_width`-bit fields
  u64 n_occs;
  u32 ctable_width;
  u32 ref_shift;
  u64 pos_mask;
  PackedVecView contig_offsets;
  // references (validate_self)
  const u64* refseq;
  const u64* ref_prefix;
  u64 n_refs;
};

}  // namespace mazu
