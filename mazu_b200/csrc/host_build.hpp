// Host side of the index: the containers mazu's Rust host would hand over (UnitigSet, packed
// vectors, BooPHF, occurrence tables), the one-time builders that mirror
// SSHash::from_unitig_set / PFHash::from_unitig_set, and the conversion into the device layout
// of index_layout.hpp.  Multi-threaded (std::thread) where the reference uses rayon.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mazu_b200.h"
#include "index_layout.hpp"

namespace mazu {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline unsigned host_threads() {
  unsigned n = std::thread::hardware_concurrency();
  if (n == 0) n = 1;
  if (const char* e = getenv("MAZU_B200_THREADS")) {
    int v = atoi(e);
    if (v > 0) n = (unsigned)v;
  }
  return std::min(n, 64u);
}

// run f(t, lo, hi) over [0, n) split into contiguous ranges, one per thread
template <class F>
inline void parallel_ranges(u64 n, unsigned threads, F&& f) {
  if (threads <= 1 || n < 4096) {
    f(0u, (u64)0, n);
    return;
  }
  std::vector<std::thread> ts;
  for (unsigned t = 0; t < threads; ++t) ts.emplace_back([&, t] { f(t, n * t / threads, n * (t + 1) / threads); });
  for (auto& t : ts) t.join();
}

inline u64 msb(u64 n) { return n == 0 ? 0 : 63 - (u64)__builtin_clzll(n); }  // src/util.rs:48-55

// ---------------------------------------------------------------------------------------------
// simple-sds IntVector / pufferfish compact vector
// ---------------------------------------------------------------------------------------------
struct PackedVec {
  u64 width = 1, len = 0;
  std::vector<u64> words;
  PackedVec() {}
  PackedVec(u64 n, u64 w) : width(w), len(n), words((n * w + 63) / 64 + 2, 0) {}
  u64 get(u64 i) const {
    PackedVecView v = view();
    return packed_get(v, i);
  }
  void set(u64 i, u64 x) {
    u64 bit = i * width, wi = bit >> 6, sh = bit & 63;
    u64 m = width >= 64 ? ~0ULL : ((1ULL << width) - 1);
    x &= m;
    words[wi] = (words[wi] & ~(m << sh)) | (x << sh);
    if (sh + width > 64) {
      u64 hi = sh + width - 64, mh = (1ULL << hi) - 1;
      words[wi + 1] = (words[wi + 1] & ~mh) | (x >> (64 - sh));
    }
  }
  // IntVector::from(Vec<usize>) followed by pack(): width = bit length of the maximum (>= 1)
  static PackedVec packed(const std::vector<u64>& xs) {
    u64 mx = 0;
    for (u64 x : xs) mx = std::max(mx, x);
    PackedVec v(xs.size(), mx == 0 ? 1 : msb(mx) + 1);
    for (u64 i = 0; i < xs.size(); ++i) v.set(i, xs[i]);
    return v;
  }
  static PackedVec from_desc(const mazu_packed_vec_desc_t& d) {
    if (d.width == 0 || d.width > 64) throw Error(MAZU_ERR_INVALID_DATA, "packed vector width must be in [1,64]");
    PackedVec v(d.len, d.width);
    memcpy(v.words.data(), d.words, ((d.len * d.width + 63) / 64) * 8);
    return v;
  }
  PackedVecView view() const { return PackedVecView{words.data(), len, (u32)width, 0}; }
};

// ---------------------------------------------------------------------------------------------
// UnitigSet (src/unitig_set.rs:31-250)
// ---------------------------------------------------------------------------------------------
struct UnitigSetHost {
  u32 k = 0;
  std::vector<u64> useq;  // padded with >= 2 zero words
  u64 n_bases = 0;
  std::vector<u64> accum;  // n_unitigs + 1
  u64 n_unitigs() const { return accum.size() - 1; }
  u64 total_len() const { return accum.back(); }
  u64 n_kmers() const { return total_len() - (u64)k * n_unitigs() + n_unitigs(); }  // unitig_set.rs:212-214
  u64 unitig_len(u64 i) const { return accum[i + 1] - accum[i]; }
  u64 window(u64 pos) const {  // get_kmer_u64_from_useq_pos, unitig_set.rs:226-229
    u64 bit = 2 * pos, wi = bit >> 6, sh = bit & 63;
    u64 x = useq[wi] >> sh;
    if (sh) x |= useq[wi + 1] << (64 - sh);
    return x & kmer_mask(k);
  }
  u64 pos_to_id(u64 pos) const {  // unitig_set.rs:185-187
    return (u64)(std::upper_bound(accum.begin(), accum.end(), pos) - accum.begin()) - 1;
  }
  bool is_valid_useq_pos(u64 pos) const {  // unitig_set.rs:235-245
    if (total_len() < k || pos > total_len() - k) return false;
    return accum[pos_to_id(pos) + 1] >= pos + k;
  }
  void validate() const {
    if (k == 0 || k > 32) throw Error(MAZU_ERR_INVALID_ARG, "k must be in [1,32]");
    if (accum.empty() || accum[0] != 0) throw Error(MAZU_ERR_INVALID_DATA, "accum_lens must start at 0");
    for (size_t i = 1; i < accum.size(); ++i)
      if (accum[i] < accum[i - 1]) throw Error(MAZU_ERR_EF_NOT_MONOTONE, "unitig prefix lengths are not monotone");
    if (accum.size() < 2) throw Error(MAZU_ERR_EF_EMPTY, "empty unitig set");
    if (accum.back() != n_bases) throw Error(MAZU_ERR_INVALID_DATA, "accum_lens.back() != n_bases");
  }
  static UnitigSetHost from_desc(const mazu_unitig_set_desc_t& d) {
    UnitigSetHost u;
    u.k = d.k;
    u.n_bases = d.n_bases;
    u64 nw = (2 * d.n_bases + 63) / 64;
    u.useq.assign(nw + 2, 0);
    memcpy(u.useq.data(), d.useq_words, nw * 8);
    if ((2 * d.n_bases) & 63) u.useq[nw - 1] &= (1ULL << ((2 * d.n_bases) & 63)) - 1;
    u.accum.assign(d.accum_lens, d.accum_lens + d.n_unitigs + 1);
    u.validate();
    return u;
  }
  // push ASCII bases (UnitigSet::from_seqs / from_cf_reduced_gfa, unitig_set.rs:74-165)
  void push_seq(const char* s, u64 n) {
    if (accum.empty()) accum.push_back(0);
    u64 need = (2 * (n_bases + n) + 63) / 64 + 2;
    if (useq.size() < need) useq.resize(std::max<u64>(need, useq.size() * 2), 0);
    for (u64 i = 0; i < n; ++i) {
      u32 c = base_code((u8)s[i]);
      if (c > 3) throw Error(MAZU_ERR_INVALID_DATA, "non-ACGT base in unitig sequence");
      u64 bit = 2 * (n_bases + i);
      useq[bit >> 6] |= (u64)c << (bit & 63);
    }
    n_bases += n;
    accum.push_back(n_bases);
  }
};

// ---------------------------------------------------------------------------------------------
// Blocked Elias-Fano builder (layout: index_layout.hpp BlockedEFView; rules: src/elias_fano.rs:55-114)
// ---------------------------------------------------------------------------------------------
struct BlockedEF {
  u32 l = 1, log_s = 5;
  u64 n = 0;
  std::vector<u64> blocks, exceptions;
  u64 n_exception_blocks = 0;
  u32 wpb = 4;  // words per block: 4, or 8 when fingerprints ride along (element i's byte at word 4 + j/8 of its block)
  // fps: optional, one byte per element i in [0, n-1) (the fingerprint of the key whose MPHF value is i)
  static BlockedEF build(const std::vector<u64>& xs, const std::vector<u8>* fps = nullptr) {
    if (xs.empty()) throw Error(MAZU_ERR_EF_EMPTY, "EFVector: empty input");  // elias_fano.rs:124-140
    for (size_t i = 1; i < xs.size(); ++i)
      if (xs[i] < xs[i - 1]) throw Error(MAZU_ERR_EF_NOT_MONOTONE, "EFVector: sequence is not monotone");  // :89-91
    BlockedEF ef;
    u64 n = xs.size(), u = xs.back();
    ef.n = n;
    u64 l = msb(u / n);
    if (l == 0) l = 1;  // elias_fano.rs:64-75
    ef.l = (u32)l;
    u32 log_s = 5;
    while (log_s > 0 && ((1ULL << log_s) + 1) * l > 64) --log_s;
    bool all_exc = ((1ULL << log_s) + 1) * l > 64 || xs.back() >> 63;
    ef.log_s = log_s;
    u64 S = 1ULL << log_s;
    u64 nb = (n + S - 1) / S;  // block b covers [b*S, b*S+S]
    ef.wpb = fps ? 8 : 4;
    const u64 wpb = ef.wpb;
    ef.blocks.assign(nb * wpb, 0);
    for (u64 b = 0; b < nb; ++b) {
      u64 i0 = b * S, cnt = std::min<u64>(S + 1, n - i0);
      u64 hb = xs[i0] >> l;
      bool exc = all_exc || ((xs[i0 + cnt - 1] >> l) - hb) + (cnt - 1) >= 128;
      u64* w = &ef.blocks[b * wpb];
      if (fps)
        for (u64 j = 0; j < S && i0 + j < fps->size(); ++j) w[4 + (j >> 3)] |= (u64)(*fps)[i0 + j] << (8 * (j & 7));
      if (exc) {
        w[0] = (1ULL << 63) | ef.n_exception_blocks++;
        for (u64 j = 0; j <= S; ++j) ef.exceptions.push_back(j < cnt ? xs[i0 + j] : xs[i0 + cnt - 1]);
        continue;
      }
      w[0] = xs[i0];
      u64 lmask = (1ULL << l) - 1;
      for (u64 j = 0; j < cnt; ++j) {
        u64 x = xs[i0 + j];
        u64 p = ((x >> l) - hb) + j;
        w[1 + (p >> 6)] |= 1ULL << (p & 63);
        w[3] |= (x & lmask) << (j * l);
      }
    }
    if (ef.exceptions.empty()) ef.exceptions.push_back(0);
    return ef;
  }
  BlockedEFView view() const { return BlockedEFView{blocks.data(), exceptions.data(), n, l, log_s, wpb, 0}; }
  u64 get(u64 i) const {  // EFVector::get, elias_fano.rs:116-122
    u64 a, b;
    BlockedEFView v = view();
    if (n == 1) return (blocks[0] >> 63) ? exceptions[0] : blocks[0];
    if (i + 1 < n) {
      blocked_ef_get2(v, i, a, b);
      return a;
    }
    blocked_ef_get2(v, i - 1, a, b);
    return b;
  }
};

// ---------------------------------------------------------------------------------------------
// MPHF host side: ranked bitset blocks (index_layout.hpp) + builders
// ---------------------------------------------------------------------------------------------
struct MphfHost {
  std::vector<u32> blocks;
  RankedLevels meta{};  // pointers filled by view()
  std::vector<u64> fb_keys, fb_vals;

  RankedLevels view() const {
    RankedLevels m = meta;
    m.blocks = blocks.data();
    m.fb_keys = fb_keys.data();
    m.fb_vals = fb_vals.data();
    m.n_fb = (u32)n_fb_real_;
    return m;
  }
  bool lookup(u64 key, u64& out) const {
    RankedLevels v = view();
    return mphf_lookup(v, key, out);
  }
  u64 hash(u64 key) const {
    u64 h;
    if (!lookup(key, h)) throw Error(MAZU_ERR_OTHER, "internal: MPHF miss on a member key");
    return h;
  }

  // turn a level's plain bit array (bit i = slot i) into ranked blocks appended to `blocks`
  void append_level_from_bits(const std::vector<u64>& bits, u64 n_slots, u64 size_field) {
    u32 lvl = meta.n_levels++;
    if (lvl >= MPHF_MAX_LEVELS) throw Error(MAZU_ERR_INVALID_DATA, "MPHF has more than 32 levels");
    u64 nb = (n_slots + MPHF_BLOCK_BITS - 1) / MPHF_BLOCK_BITS;
    if (nb == 0) nb = 1;
    u64 off = blocks.size() / 8;
    blocks.resize((off + nb) * 8, 0);
    meta.size[lvl] = size_field ? size_field : nb;
    meta.block_off[lvl] = off;
    meta.rank_base[lvl] = lvl == 0 ? 0 : meta.rank_base[lvl - 1] + level_ones_[lvl - 1];
    u64 ones = 0;
    auto getbits = [&](u64 pos, u32 cnt) -> u32 {  // cnt <= 32 bits at bit offset pos of `bits`
      u64 wi = pos >> 6, sh = pos & 63;
      u64 x = wi < bits.size() ? bits[wi] >> sh : 0;
      if (sh + cnt > 64 && wi + 1 < bits.size()) x |= bits[wi + 1] << (64 - sh);
      return (u32)(x & (cnt >= 32 ? 0xFFFFFFFFULL : ((1ULL << cnt) - 1)));
    };
    for (u64 b = 0; b < nb; ++b) {
      u32* blk = &blocks[(off + b) * 8];
      if (ones >> 32) throw Error(MAZU_ERR_INVALID_DATA, "MPHF level holds more than 2^32 keys");
      blk[0] = (u32)ones;
      for (u32 j = 0; j < 7; ++j) {
        u64 pos = b * MPHF_BLOCK_BITS + 32ULL * j;
        u32 x = 0;
        if (pos < n_slots) {
          u32 cnt = (u32)std::min<u64>(32, n_slots - pos);
          x = getbits(pos, cnt);
        }
        blk[1 + j] = x;
        ones += (u64)__builtin_popcount(x);
      }
    }
    level_ones_.push_back(ones);
  }
  u64 total_level_ones() const {
    u64 s = 0;
    for (u64 o : level_ones_) s += o;
    return s;
  }

  // BooPHF<u64> as parsed from pufferfish's mphf.bin (src/pf1/boophf/mod.rs:50-86,269-293).
  // The bits are re-blocked; hash values are unchanged (rank is layout-independent).
  static MphfHost from_boophf(const mazu_boophf_desc_t& d) {
    MphfHost m;
    m.meta.family = MPHF_FAMILY_BOOPHF;
    m.meta.n_keys = d.n_elem;
    for (u32 l = 0; l < d.n_levels; ++l) {
      u64 nbits = d.level_n_bits[l];
      std::vector<u64> bits(d.level_words[l], d.level_words[l] + (nbits + 63) / 64);
      m.append_level_from_bits(bits, nbits, nbits);
    }
    if (m.total_level_ones() != d.last_bitset_rank && d.n_final)
      throw Error(MAZU_ERR_INVALID_DATA, "BooPHF: last_bitset_rank does not match the level popcounts");
    std::vector<std::pair<u64, u64>> kv;
    for (u64 i = 0; i < d.n_final; ++i) kv.push_back({d.final_keys[i], d.final_vals[i] + d.last_bitset_rank});  // mod.rs:177-181
    std::sort(kv.begin(), kv.end());
    for (auto& p : kv) {
      m.fb_keys.push_back(p.first);
      m.fb_vals.push_back(p.second);
    }
    if (m.fb_keys.empty()) {  // keep the arrays non-empty so they can be uploaded
      m.fb_keys.push_back(0);
      m.fb_vals.push_back(0);
    }
    m.n_fb_real_ = kv.size();
    return m;
  }

  // Native MPHF over distinct u64 keys (stands where the reference uses boomphf::Mphf::new_parallel(1.7, ..),
  // src/kphf/sshash.rs:177,280 / src/kphf/pfhash.rs:43-49).  BBHash cascade, multi-threaded.
  static MphfHost build_native(const std::vector<u64>& keys_in, double gamma, unsigned threads) {
    MphfHost m;
    m.meta.family = MPHF_FAMILY_NATIVE;
    std::vector<u64> keys = keys_in, next;
    for (u32 lvl = 0; lvl < MPHF_MAX_LEVELS && !keys.empty(); ++lvl) {
      u64 n = keys.size();
      u64 nb = (u64)((native_level_gamma(gamma, lvl) * (double)n) / MPHF_BLOCK_BITS) + 1;
      if (nb >> 32) throw Error(MAZU_ERR_INVALID_ARG, "MPHF level too large");
      u64 n_slots = nb * MPHF_BLOCK_BITS;
      std::vector<u64> seen((n_slots + 63) / 64, 0), coll((n_slots + 63) / 64, 0);
      parallel_ranges(n, threads, [&](unsigned, u64 lo, u64 hi) {
        for (u64 i = lo; i < hi; ++i) {
          u64 blk;
          u32 bit;
          native_slot(keys[i], lvl, nb, blk, bit);
          u64 s = blk * MPHF_BLOCK_BITS + bit, mask = 1ULL << (s & 63);
          u64 old = __atomic_fetch_or(&seen[s >> 6], mask, __ATOMIC_RELAXED);
          if (old & mask) __atomic_fetch_or(&coll[s >> 6], mask, __ATOMIC_RELAXED);
        }
      });
      std::vector<std::vector<u64>> nexts(threads ? threads : 1);
      parallel_ranges(n, threads, [&](unsigned t, u64 lo, u64 hi) {
        for (u64 i = lo; i < hi; ++i) {
          u64 blk;
          u32 bit;
          native_slot(keys[i], lvl, nb, blk, bit);
          u64 s = blk * MPHF_BLOCK_BITS + bit;
          if ((coll[s >> 6] >> (s & 63)) & 1) nexts[t].push_back(keys[i]);
        }
      });
      for (size_t i = 0; i < seen.size(); ++i) seen[i] &= ~coll[i];
      m.append_level_from_bits(seen, n_slots, nb);
      next.clear();
      for (auto& v : nexts) next.insert(next.end(), v.begin(), v.end());
      keys.swap(next);
    }
    // leftovers (duplicates or astronomically unlucky keys): sorted fallback
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    u64 base = m.total_level_ones();
    for (u64 i = 0; i < keys.size(); ++i) {
      m.fb_keys.push_back(keys[i]);
      m.fb_vals.push_back(base + i);
    }
    m.meta.n_keys = base + keys.size();
    m.n_fb_real_ = m.fb_keys.size();
    if (m.fb_keys.empty()) {
      m.fb_keys.push_back(0);
      m.fb_vals.push_back(0);
    }
    return m;
  }
  u64 n_fb_real() const { return n_fb_real_; }
  // Fingerprinted cascade over distinct u64 keys (index_layout.hpp, MPHF_FAMILY_CASCADE): the perfect hash of the SSHash
  // minimizers (stands where the reference uses boomphf::Mphf::new_parallel(1.7, ..), src/kphf/sshash.rs:177).
  // states[g] = the byte of slot g (EMPTY / COLLIDED / fingerprint), slots[i] = the slot of keys[i].
  static MphfHost build_cascade(const std::vector<u64>& keys_in, unsigned threads, std::vector<u8>& states, std::vector<u64>& slots) {
    MphfHost m;
    m.meta.family = MPHF_FAMILY_CASCADE;
    const u64 n_in = keys_in.size();
    slots.assign(n_in, ~0ULL);
    states.clear();
    std::vector<u64> idx(n_in), next_idx;
    for (u64 i = 0; i < n_in; ++i) idx[i] = i;
    u64 off = 0;
    for (u32 lvl = 0; lvl < MPHF_MAX_LEVELS && !idx.empty(); ++lvl) {
      const u64 n = idx.size(), size = cascade_level_size(n, lvl);
      std::vector<u64> seen((size + 63) / 64, 0), coll((size + 63) / 64, 0);
      parallel_ranges(n, threads, [&](unsigned, u64 lo, u64 hi) {
        for (u64 i = lo; i < hi; ++i) {
          const u64 sl = cascade_slot(fmix64(keys_in[idx[i]]), lvl, size), mask = 1ULL << (sl & 63);
          const u64 old = __atomic_fetch_or(&seen[sl >> 6], mask, __ATOMIC_RELAXED);
          if (old & mask) __atomic_fetch_or(&coll[sl >> 6], mask, __ATOMIC_RELAXED);
        }
      });
      states.resize(off + size, (u8)CASCADE_EMPTY);
      std::vector<std::vector<u64>> nexts(threads ? threads : 1);
      parallel_ranges(n, threads, [&](unsigned t, u64 lo, u64 hi) {
        for (u64 i = lo; i < hi; ++i) {
          const u64 hk = fmix64(keys_in[idx[i]]), sl = cascade_slot(hk, lvl, size);
          if ((coll[sl >> 6] >> (sl & 63)) & 1) {
            states[off + sl] = (u8)CASCADE_COLLIDED;  // every collider writes the same byte
            nexts[t].push_back(idx[i]);
          } else {
            states[off + sl] = (u8)cascade_fp(hk);
            slots[idx[i]] = off + sl;
          }
        }
      });
      m.meta.size[lvl] = size;
      m.meta.block_off[lvl] = off;
      m.meta.rank_base[lvl] = 0;
      m.meta.n_levels = lvl + 1;
      off += size;
      next_idx.clear();
      for (auto& v : nexts) next_idx.insert(next_idx.end(), v.begin(), v.end());
      idx.swap(next_idx);
    }
    // leftovers (duplicated keys, or keys that collided on every level): sorted fallback, slots behind the last level
    std::vector<std::pair<u64, u64>> left;
    for (u64 i : idx) left.push_back({keys_in[i], i});
    std::sort(left.begin(), left.end());
    u64 n_fb = 0;
    for (size_t i = 0; i < left.size(); ++i) {
      if (i == 0 || left[i].first != left[i - 1].first) {
        m.fb_keys.push_back(left[i].first);
        m.fb_vals.push_back(off + n_fb);
        states.push_back((u8)cascade_fp(fmix64(left[i].first)));
        ++n_fb;
      }
      slots[left[i].second] = off + n_fb - 1;
    }
    m.meta.n_keys = off + n_fb;  // the RANGE of the hash (one bucket per slot), not the number of keys
    m.n_fb_real_ = n_fb;
    if (m.fb_keys.empty()) {
      m.fb_keys.push_back(0);
      m.fb_vals.push_back(0);
    }
    return m;
  }
  bool cascade_lookup_host(const BlockedEFView& ef, u64 key, u64& out) const {
    RankedLevels v = view();
    return cascade_lookup(v, ef, key, out);
  }

 private:
  std::vector<u64> level_ones_;
  u64 n_fb_real_ = 0;
};

// ---------------------------------------------------------------------------------------------
// stable sort of records by a small-range key, multi-threaded: stable counting scatter on the top
// key bits, then an independent std::stable_sort per bucket (the reference uses rayon
// par_sort_by_key, which is stable -- bucket order decides first-match order in SSHash::k2u).
// ---------------------------------------------------------------------------------------------
template <class T, class KeyFn>
inline void parallel_stable_sort_by_key(std::vector<T>& v, KeyFn key, u32 key_bits, unsigned threads) {
  u64 n = v.size();
  if (n < (1u << 16) || threads <= 1) {
    std::stable_sort(v.begin(), v.end(), [&](const T& a, const T& b) { return key(a) < key(b); });
    return;
  }
  u32 rb = std::min<u32>(12, key_bits);
  u32 shift = key_bits - rb;
  u64 nbk = 1ULL << rb;
  std::vector<u64> start(nbk + 1, 0);
  for (u64 i = 0; i < n; ++i) start[(key(v[i]) >> shift) + 1]++;
  for (u64 b = 0; b < nbk; ++b) start[b + 1] += start[b];
  std::vector<T> tmp(n);
  {
    std::vector<u64> cur(start.begin(), start.end() - 1);
    for (u64 i = 0; i < n; ++i) tmp[cur[key(v[i]) >> shift]++] = v[i];
  }
  v.swap(tmp);
  std::atomic<u64> nextb{0};
  std::vector<std::thread> ts;
  for (unsigned t = 0; t < threads; ++t)
    ts.emplace_back([&] {
      for (;;) {
        u64 b = nextb.fetch_add(1);
        if (b >= nbk) break;
        std::stable_sort(v.begin() + start[b], v.begin() + start[b + 1], [&](const T& a, const T& c) { return key(a) < key(c); });
      }
    });
  for (auto& t : ts) t.join();
}

// ---------------------------------------------------------------------------------------------
// K2U host parts
// ---------------------------------------------------------------------------------------------
struct K2UHost {
  int kind = MAZU_K2U_PFHASH;
  std::shared_ptr<const UnitigSetHost> unitigs;
  MphfHost mphf;
  PackedVec pos;
  // SSHash
  u32 w = 0;
  u64 seed = 0, skew_param = MAZU_SKEW_NONE;
  BlockedEF sizes;
  u64 n_minimizers = 0;      // distinct minimizers
  u64 n_minimizer_occs = 0;  // |pos|
  bool has_skew = false;
  MphfHost skew_mphf;
  PackedVec skew_pos;
  u64 n_skew_kmers = 0;
  // SampledPFHash (pufferfish sparse index): `pos` holds sampled_pos
  MphfHost sampled;  // presence bits as a single ranked level
  std::vector<u64> canonical_bits, direction_bits;
  PackedVec ext_sizes, ext_bases;
  u64 sample_size = 0, extension_size = 0;
};

struct MinOcc {
  u64 word, pos;
};

// minimizer occurrences of one unitig in the reference's order (src/kphf/sshash.rs:100-143):
// the fw-canonical k-mers' stream then the rc-canonical k-mers' stream, each run-length deduped.
inline void collect_unitig_minimizers(const UnitigSetHost& us, u64 ui, u32 w, u64 seed, std::vector<u32>& hf, std::vector<u32>& hr,
                                      std::vector<MinOcc>& out) {
  const u32 k = us.k;
  const u64 s = us.accum[ui], e = us.accum[ui + 1];
  if (e - s < k) return;
  const u64 n_w = e - s - w + 1;  // w-mers of this unitig
  const u64 wmask = kmer_mask(w);
  hf.resize(n_w);
  hr.resize(n_w);
  auto base = [&](u64 p) { return (us.useq[(2 * p) >> 6] >> ((2 * p) & 63)) & 3ULL; };
  {
    u64 f = 0, r = 0;
    for (u64 i = 0; i < e - s; ++i) {
      u64 c = base(s + i);
      f = ((f >> 2) | (c << (2 * (w - 1)))) & wmask;
      r = ((r << 2) | (3 - c)) & wmask;
      if (i + 1 >= w) {
        hf[i + 1 - w] = hr[i + 1 - w] = mm_key(f, r, seed);  // strand-symmetric key (minimizer order v3)
      }
    }
  }
  const u32 span = k - w;  // window holds span+1 w-mers
  for (int pass = 0; pass < 2; ++pass) {
    bool have_prev = false;
    MinOcc prev{0, 0};
    u64 fw = 0, rc = 0;
    const u64 kmask = kmer_mask(k);
    for (u64 i = 0; i < e - s; ++i) {
      u64 c = base(s + i);
      fw = ((fw >> 2) | (c << (2 * (k - 1)))) & kmask;
      rc = ((rc << 2) | (3 - c)) & kmask;
      if (i + 1 < k) continue;
      u64 p = i + 1 - k;  // k-mer start inside the unitig
      bool fw_canon = fw <= rc;
      if (fw_canon != (pass == 0)) continue;
      // smallest (hash key | offset in the canonical k-mer): minimizer order v3 (kmer.hpp)
      u32 best = 0xFFFFFFFFu;
      for (u32 ci = 0; ci <= span; ++ci) {
        u32 key = (fw_canon ? hf[p + ci] : hr[p + span - ci]) | ci;
        best = key < best ? key : best;
      }
      u32 best_i = best & 31u;
      u64 off_fw = fw_canon ? best_i : (span - best_i);  // offset in fw-mer coordinates
      const u64 wf = (fw >> (2 * off_fw)) & wmask, wr = (rc >> (2 * (span - off_fw))) & wmask;
      u64 word = wf <= wr ? wf : wr;
      MinOcc cur{word, s + p + off_fw};
      if (!have_prev || cur.word != prev.word || cur.pos != prev.pos) out.push_back(cur);
      prev = cur;
      have_prev = true;
    }
  }
}

// SSHashBuilder::from_unitig_set + finish (src/kphf/sshash.rs:86-329)
inline std::shared_ptr<K2UHost> build_sshash(std::shared_ptr<const UnitigSetHost> usp, u32 w, u64 skew_param, u64 seed, double gamma = 2.0) {
  const UnitigSetHost& us = *usp;
  const u32 k = us.k;
  if (w == 0 || w > k) throw Error(MAZU_ERR_INVALID_ARG, "minimizer length w must satisfy 1 <= w <= k");  // sshash.rs:92
  if (us.n_kmers() == 0 || us.total_len() < k) throw Error(MAZU_ERR_INVALID_DATA, "unitig set holds no k-mer");
  const unsigned T = host_threads();
  auto H = std::make_shared<K2UHost>();
  H->kind = MAZU_K2U_SSHASH;
  H->unitigs = usp;
  H->w = w;
  H->seed = seed;
  H->skew_param = skew_param;
  // 1. collect (parallel over contiguous unitig ranges, concatenated in unitig order)
  const u64 U = us.n_unitigs();
  std::vector<MinOcc> minimizers;
  {
    unsigned parts = (U < 64) ? 1 : T;
    std::vector<std::vector<MinOcc>> per(parts);
    // split by cumulative length so long unitigs do not unbalance threads
    std::vector<u64> cut(parts + 1, U);
    cut[0] = 0;
    for (unsigned t = 1; t < parts; ++t) {
      u64 target = us.total_len() / parts * t;
      cut[t] = (u64)(std::lower_bound(us.accum.begin(), us.accum.end(), target) - us.accum.begin());
      cut[t] = std::min(cut[t], U);
    }
    std::vector<std::thread> ts;
    for (unsigned t = 0; t < parts; ++t)
      ts.emplace_back([&, t] {
        std::vector<u32> hf, hr;
        for (u64 ui = cut[t]; ui < cut[t + 1]; ++ui) collect_unitig_minimizers(us, ui, w, seed, hf, hr, per[t]);
      });
    for (auto& t : ts) t.join();
    u64 tot = 0;
    for (auto& v : per) tot += v.size();
    minimizers.reserve(tot);
    for (auto& v : per) {
      minimizers.insert(minimizers.end(), v.begin(), v.end());
      std::vector<MinOcc>().swap(v);
    }
  }
  // 2. stable sort by word, group (sshash.rs:150-172)
  parallel_stable_sort_by_key(minimizers, [](const MinOcc& m) { return m.word; }, 2 * w, T);
  // 2b. A super-k-mer that spans a flip of the canonical strand is pushed by both of the reference's streams (fw-canonical
  // k-mers, then rc-canonical ones: sshash.rs:100-143), so after the stable sort its (word, position) sits twice in a row in
  // its bucket -- under the strand-symmetric minimizer order that is the commonest bucket there is.  The second copy offers
  // the two candidates of the first, which have just failed when the bucket loop gets to it (sshash.rs:494-552 returns at the
  // first match): dropping it cannot change an answer, and a k-mer that misses stops one dependent access earlier.  Buckets
  // that the reference sends to the skew index (more than skew_param entries, sshash.rs:232,486-490) are left as they are,
  // so every bucket takes the same path as in the reference.
  {
    u64 out = 0;
    for (u64 i = 0; i < minimizers.size();) {
      u64 j = i;
      while (j < minimizers.size() && minimizers[j].word == minimizers[i].word) ++j;
      const bool heavy = j - i > skew_param;
      for (u64 t = i; t < j; ++t)
        if (heavy || t == i || minimizers[t].pos != minimizers[t - 1].pos) minimizers[out++] = minimizers[t];
      i = j;
    }
    minimizers.resize(out);
  }
  std::vector<u64> mm_set, ranges;  // ranges = prefix sum of mm_occs
  for (u64 i = 0; i < minimizers.size(); ++i)
    if (i == 0 || minimizers[i].word != minimizers[i - 1].word) {
      mm_set.push_back(minimizers[i].word);
      ranges.push_back(i);
    }
  ranges.push_back(minimizers.size());
  const u64 M = mm_set.size();
  H->n_minimizers = M;
  H->n_minimizer_occs = minimizers.size();
  // 3. perfect hash over the minimizer set (sshash.rs:177): the fingerprinted cascade; its value is the minimizer's slot
  std::vector<u8> states;
  std::vector<u64> hashes;
  H->mphf = MphfHost::build_cascade(mm_set, T, states, hashes);
  const u64 R = H->mphf.meta.n_keys;  // slots: one (possibly empty) bucket each
  // 4. bucket sizes in slot order -> prefix sum (sshash.rs:181-189)
  std::vector<u64> prefix(R + 1, 0);
  for (u64 i = 0; i < M; ++i) {
    if (hashes[i] >= R) throw Error(MAZU_ERR_OTHER, "internal: perfect-hash value out of range");
    prefix[hashes[i] + 1] = ranges[i + 1] - ranges[i];
  }
  for (u64 i = 0; i < R; ++i) prefix[i + 1] += prefix[i];
  // 5. scatter positions (sshash.rs:196-219)
  std::vector<u64> pos(minimizers.size());
  parallel_ranges(M, T, [&](unsigned, u64 lo, u64 hi) {
    for (u64 i = lo; i < hi; ++i) {
      u64 sh = prefix[hashes[i]];
      for (u64 j = ranges[i]; j < ranges[i + 1]; ++j) pos[sh + (j - ranges[i])] = minimizers[j].pos;
    }
  });
  // 6. skew index (sshash.rs:222-296)
  if (skew_param != MAZU_SKEW_NONE) {
    std::vector<MinOcc> tuples;  // (canonical k-mer word, position)
    for (u64 i = 0; i < M; ++i) {
      if (ranges[i + 1] - ranges[i] <= skew_param) continue;  // sshash.rs:232
      for (u64 j = ranges[i]; j < ranges[i + 1]; ++j) {
        u64 mp = minimizers[j].pos;
        u64 start_pos = mp < (u64)(k - w) ? 0 : mp - (u64)(k - w);  // sshash.rs:241-247
        for (u64 off = 0; off < (u64)(k - w + 1); ++off) {
          u64 p = start_pos + off;
          if (us.is_valid_useq_pos(p)) {
            u64 fw = us.window(p);
            tuples.push_back(MinOcc{std::min(fw, revcomp(fw, k)), p});
          }
        }
      }
    }
    parallel_stable_sort_by_key(tuples, [](const MinOcc& m) { return m.word; }, 2 * k, T);
    std::vector<u64> km_set, km_pos;
    for (u64 i = 0; i < tuples.size(); ++i)
      if (i == 0 || tuples[i].word != tuples[i - 1].word) {  // dedup_by_key keeps the first (sshash.rs:273-274)
        km_set.push_back(tuples[i].word);
        km_pos.push_back(tuples[i].pos);
      }
    H->has_skew = true;
    H->n_skew_kmers = km_set.size();
    H->skew_mphf = MphfHost::build_native(km_set, gamma, T);
    std::vector<u64> sp(std::max<u64>(km_set.size(), 1), 0);
    for (u64 i = 0; i < km_set.size(); ++i) sp[H->skew_mphf.hash(km_set[i])] = km_pos[i];
    sp.resize(km_set.size());
    H->skew_pos = PackedVec::packed(sp);
  }
  // finish (sshash.rs:310-329): Elias-Fano encode the prefix sums (the slot states ride in the blocks), bit-pack the positions
  H->sizes = BlockedEF::build(prefix, &states);
  H->pos = PackedVec::packed(pos);
  return H;
}

// PFHash::from_unitig_set (src/kphf/pfhash.rs:40-73)
inline std::shared_ptr<K2UHost> build_pfhash(std::shared_ptr<const UnitigSetHost> usp, double gamma = 2.0) {
  const UnitigSetHost& us = *usp;
  const u32 k = us.k;
  const unsigned T = host_threads();
  auto H = std::make_shared<K2UHost>();
  H->kind = MAZU_K2U_PFHASH;
  H->unitigs = usp;
  const u64 U = us.n_unitigs(), N = us.n_kmers();
  // k-mer index base per unitig (prefix of per-unitig k-mer counts)
  std::vector<u64> kbase(U + 1, 0);
  for (u64 ui = 0; ui < U; ++ui) kbase[ui + 1] = kbase[ui] + (us.unitig_len(ui) >= k ? us.unitig_len(ui) - k + 1 : 0);
  if (kbase[U] != N) throw Error(MAZU_ERR_INVALID_DATA, "a unitig is shorter than k");
  std::vector<u64> keys(N);
  parallel_ranges(U, T, [&](unsigned, u64 lo, u64 hi) {
    for (u64 ui = lo; ui < hi; ++ui) {
      u64 s = us.accum[ui];
      for (u64 j = 0; j < kbase[ui + 1] - kbase[ui]; ++j) {
        u64 fw = us.window(s + j);
        keys[kbase[ui] + j] = std::min(fw, revcomp(fw, k));
      }
    }
  });
  H->mphf = MphfHost::build_native(keys, gamma, T);
  std::vector<u64> pos(H->mphf.meta.n_keys, 0);
  parallel_ranges(U, T, [&](unsigned, u64 lo, u64 hi) {
    for (u64 ui = lo; ui < hi; ++ui) {
      u64 s = us.accum[ui];
      for (u64 j = 0; j < kbase[ui + 1] - kbase[ui]; ++j) pos[H->mphf.hash(keys[kbase[ui] + j])] = s + j;  // pfhash.rs:60-68
    }
  });
  H->pos = PackedVec(pos.size(), std::max<u64>(1, msb(std::max<u64>(us.total_len(), 1)) + 1));
  for (u64 i = 0; i < pos.size(); ++i) H->pos.set(i, pos[i]);
  return H;
}

// PFHash::from_parts(unitigs, BooPHF, pos) (src/kphf/pfhash.rs:34-36)
inline std::shared_ptr<K2UHost> pfhash_from_parts(std::shared_ptr<const UnitigSetHost> usp, const mazu_boophf_desc_t& mphf,
                                                  const mazu_packed_vec_desc_t& pos) {
  auto H = std::make_shared<K2UHost>();
  H->kind = MAZU_K2U_PFHASH;
  H->unitigs = usp;
  H->mphf = MphfHost::from_boophf(mphf);
  H->pos = PackedVec::from_desc(pos);
  if (H->pos.len != usp->n_kmers()) throw Error(MAZU_ERR_INVALID_DATA, "pos.len() != unitigs.n_kmers()");  // dense_index.rs:80
  return H;
}

// ---------------------------------------------------------------------------------------------
// U2Pos / references host parts
// ---------------------------------------------------------------------------------------------
struct U2PosHost {
  int kind = MAZU_U2POS_NONE;
  std::vector<u64> ctable_words;  // DENSE: u64 per occurrence; PISCEM: packed
  u64 n_occs = 0;
  u32 ctable_width = 64;
  u64 ref_shift = 0, pos_mask = 0;
  PackedVec contig_offsets;
  std::vector<std::string> ref_names;
};
struct RefSeqHost {
  bool has_seq = false;
  std::vector<u64> seq_words;
  std::vector<u64> prefix;  // n_refs + 1
  u64 n_refs() const { return prefix.empty() ? 0 : prefix.size() - 1; }
};

}  // namespace mazu
