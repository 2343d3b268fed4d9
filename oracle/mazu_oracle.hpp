// =====================================================================================
//  mazu_oracle.hpp  --  TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT.
//
//  A plain CPU restatement (queries single-threaded per call site, the one-time builder
//  optionally threaded like the reference's rayon builder) of the batched k-mer query path of
//  COMBINE-lab/mazu (Rust), used as the bit-exactness checker for the CUDA path and as the
//  timed "port" CPU baseline in bench.py.  Only tests/, __graft_entry__.smoke() and
//  bench.py's cpu_baseline / --impl reference legs may load this code.  The product
//  (mazu_b200/csrc, libmazu_b200.so) never includes, links or calls anything in oracle/.
//
//  Every function cites the reference file:line (under /root/reference) it restates.
//  Data layouts deliberately follow the *reference's* structures (rank bit-vector for
//  pos_to_id, Elias-Fano get via select, separate rank samples in BooPHF, ...) -- the CUDA
//  path uses different, device-oriented layouts, so agreement between the two is a real check.
//
//  PARITY PINNING (see DESIGN.md "Oracle"):
//   * pinned by the reference's own golden vectors / C++-pufferfish fixtures (tests/test_oracle_*.py):
//       SimpleHash, MultiHash, BooPHF lookup (bbhash_n=10 + example_*.json), compact vectors,
//       DenseIndex tiny MappedRefPos answers, SSHash K2UPos answers, SPT occurrences,
//       occ packing, Elias-Fano, validate_self counts on yeast_chr01.
//   * parity UNPINNED (third-party crates absent from /root/reference, Cargo.lock git-ignored):
//       the minimizer order (`kmers::canonical_minimizer` + `wyhash` 0.5.0) and the SSHash MPHF
//       (`boomphf` 0.5.9).  k-mer -> (unitig,pos,orientation) results are provably independent
//       of both (every candidate is verified against the packed sequence, sshash.rs:500-504),
//       so result parity is unaffected; see mm_hash64 / OracleMphf below.
// =====================================================================================
#pragma once
#include <algorithm>
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <chrono>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace mazu_oracle {

using u8 = uint8_t;
using u32 = uint32_t;
using u64 = uint64_t;
using usize = uint64_t;
static const u64 USIZE_MAX = ~0ULL;

struct OracleError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// ------------------------------------------------------------------------------------
// util.rs:26-38 prefix_sum, util.rs:48-55 msb
// ------------------------------------------------------------------------------------
inline std::vector<u64> prefix_sum(const std::vector<u64>& xs) {
  std::vector<u64> res;
  res.reserve(xs.size() + 1);
  u64 accum = 0;
  for (u64 x : xs) {
    res.push_back(accum);
    accum += x;
  }
  res.push_back(accum);
  return res;
}
inline u64 msb(u64 n) { return n == 0 ? 0 : 63 - (u64)__builtin_clzll(n); }

// ------------------------------------------------------------------------------------
// `kmers` crate (not in tree): 2-bit code A=0,C=1,G=2,T=3, base i at
// bits [2i,2i+2) of the word, canonical = min(fw, rc) as u64, case-insensitive.
// Pinned indirectly by the pufferfish fixtures (pos.bin is addressed by BooPHF(min(fw,rc))).
// ------------------------------------------------------------------------------------
inline int base_code(u8 c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
  }
}
inline u64 kmer_mask(int k) { return k >= 32 ? ~0ULL : ((1ULL << (2 * k)) - 1); }
// reverse complement of the low 2k bits: complement (3 - b), reverse the order of the 2-bit bases, shift down.
// (Same value as the per-base loop `rc |= (3 - b_i) << 2(k-1-i)`; tests/test_oracle_golden.py checks it against that loop.)
inline u64 revcomp_word(u64 fw, int k) {
  u64 x = ~fw;
  x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
  x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
  x = __builtin_bswap64(x);
  return k >= 32 ? x : (x >> (64 - 2 * k));
}
inline u64 revcomp_word_loop(u64 fw, int k) {
  u64 rc = 0;
  for (int i = 0; i < k; ++i) {
    u64 b = (fw >> (2 * i)) & 3;
    rc |= (3 - b) << (2 * (k - 1 - i));
  }
  return rc;
}
struct BuildTimer {  // MAZU_ORACLE_TIMING=1: phase times of the one-time builders on stderr
  bool on = getenv("MAZU_ORACLE_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void lap(const char* what) {
    auto n = std::chrono::steady_clock::now();
    if (on) fprintf(stderr, "[oracle build] %-24s %.2f s\n", what, std::chrono::duration<double>(n - t).count());
    t = n;
  }
};
// threads the one-time builders may use (the reference builds with rayon: par_sort_by_key sshash.rs:150,202,273)
inline int& build_threads() {
  static int t = 1;
  return t;
}
template <class F>
inline void parallel_blocks(u64 n, F&& f) {  // f(block index, begin, end) over `build_threads()` contiguous blocks
  const int T = (int)std::max<u64>(1, std::min<u64>((u64)build_threads(), n));
  if (T <= 1) {
    f(0, (u64)0, n);
    return;
  }
  std::vector<std::thread> ts;
  for (int t = 0; t < T; ++t) ts.emplace_back([&, t] { f(t, n * t / T, n * (t + 1) / T); });
  for (auto& th : ts) th.join();
}
// stable sort by key with the builder's threads: blocks sorted independently, then merged pairwise (stable)
template <class T, class Less>
inline void parallel_stable_sort(std::vector<T>& v, Less less) {
  const int nb = (int)std::max<u64>(1, std::min<u64>((u64)build_threads(), v.size() / 4096 + 1));
  if (nb <= 1) {
    std::stable_sort(v.begin(), v.end(), less);
    return;
  }
  std::vector<u64> cut(nb + 1);
  for (int b = 0; b <= nb; ++b) cut[b] = v.size() * (u64)b / nb;
  {
    std::vector<std::thread> ts;
    for (int b = 0; b < nb; ++b) ts.emplace_back([&, b] { std::stable_sort(v.begin() + cut[b], v.begin() + cut[b + 1], less); });
    for (auto& th : ts) th.join();
  }
  for (int width = 1; width < nb; width *= 2) {
    std::vector<std::thread> ts;
    for (int b = 0; b + width < nb; b += 2 * width)
      ts.emplace_back([&, b, width] { std::inplace_merge(v.begin() + cut[b], v.begin() + cut[b + width], v.begin() + cut[std::min(nb, b + 2 * width)], less); });
    for (auto& th : ts) th.join();
  }
}
inline std::string word_to_string(u64 w, int k) {
  std::string s(k, 'A');
  for (int i = 0; i < k; ++i) s[i] = "ACGT"[(w >> (2 * i)) & 3];
  return s;
}

enum MatchType : u32 { NoMatch = 0, IdentityMatch = 1, TwinMatch = 2 };
// record-only marker for windows the CanonicalKmerIterator skips (non-ACGT); never a K2U result
static const u32 MATCH_SKIPPED = 3;

struct CanonicalKmer {
  u64 fw = 0, rc = 0;
  int k = 0;
  static CanonicalKmer from_u64(u64 fw, int k) {
    CanonicalKmer c;
    c.fw = fw & kmer_mask(k);
    c.rc = revcomp_word(c.fw, k);
    c.k = k;
    return c;
  }
  static CanonicalKmer from_str(const std::string& s) {
    u64 w = 0;
    for (size_t i = 0; i < s.size(); ++i) {
      int c = base_code((u8)s[i]);
      if (c < 0) throw OracleError("non-ACGT base in k-mer string");
      w |= (u64)c << (2 * i);
    }
    return from_u64(w, (int)s.size());
  }
  void swap() { std::swap(fw, rc); }
  bool is_fw_canonical() const { return fw <= rc; }
  u64 canonical_word() const { return fw <= rc ? fw : rc; }
  int len() const { return k; }
  // kmers::CanonicalKmer::get_word_equivalency / get_kmer_equivalency
  MatchType word_equivalency(u64 kw) const {
    if (kw == fw) return IdentityMatch;
    if (kw == rc) return TwinMatch;
    return NoMatch;
  }
};

// CanonicalKmerIterator::from_u8_slice(seq,k): yields (pos, km) for every window of k valid
// bases; windows containing a non-ACGT byte are skipped, pos stays in read coordinates
// (inferred from cf/tiny/tiny.fa + piscem_index.rs:63-72).
template <class F>
inline void for_each_canonical_kmer(const u8* seq, u64 len, int k, F&& f) {
  u64 fw = 0, rc = 0;
  const u64 mask = kmer_mask(k);
  int valid = 0;  // number of consecutive valid bases ending at i
  for (u64 i = 0; i < len; ++i) {
    int c = base_code(seq[i]);
    if (c < 0) {
      valid = 0;
      fw = rc = 0;
      continue;
    }
    fw = ((fw >> 2) | ((u64)c << (2 * (k - 1)))) & mask;
    rc = ((rc << 2) | (u64)(3 - c)) & mask;
    if (++valid >= k) {
      CanonicalKmer km;
      km.fw = fw;
      km.rc = rc;
      km.k = k;
      f(i + 1 - k, km);
    }
  }
}

// ------------------------------------------------------------------------------------
// Minimizer order.  PARITY UNPINNED: stands in for kmers::Kmer::canonical_minimizer(w,&bh)
// with bh = WyHashState(seed) (kphf/mod.rs:32-52); neither `kmers` nor `wyhash` is in the
// reference tree.  Order v3 (shared definition): the w-mer at offset ci of the
// CANONICAL k-mer has key (mm_hash32(min(wmer, rc(wmer))) & 0xFFFFFFE0) | ci; the smallest key
// wins (top 27 hash bits, ties -> leftmost) and the minimizer word is min(wmer, rc(wmer)) -- the
// order is symmetric in the strand, so neighbouring k-mers keep their minimizer when the
// canonical strand flips.  Definition of "canonical minimizer" (sshash.rs:32-37):
// mini(g*) = mini(min(g, g')): still a function of the canonical k-mer only.
// ------------------------------------------------------------------------------------
inline u32 mm_hash32(u64 x, u64 seed) {
  u32 h = ((u32)x ^ (u32)seed) * 0x85EBCA6Bu;
  h ^= ((u32)(x >> 32) ^ (u32)(seed >> 32)) * 0xC2B2AE35u;
  h ^= h >> 16;
  h *= 0x7FEB352Du;
  h ^= h >> 15;
  h *= 0x846CA68Bu;
  h ^= h >> 16;
  return h;
}
struct Minimizer {
  u64 word;    // w-mer word as read off the canonical k-mer
  u64 offset;  // offset of that w-mer's occurrence inside the QUERIED (fw) k-mer, in [0, k-w].
               // Pinned by sshash.rs:563-624 + its tests (sshash.rs:645-763): k2u_fw finds
               // rc-canonical k-mers at mm_pos - offset, so offset is in fw-mer coordinates.
};
inline Minimizer canonical_minimizer(u64 fw_word, int k, int w, u64 seed) {
  u64 rc = revcomp_word(fw_word, k);
  const bool fw_canon = fw_word <= rc;
  const u64 c = fw_canon ? fw_word : rc, d = fw_canon ? rc : fw_word;  // canonical strand / the other one
  const u64 wmask = kmer_mask(w);
  Minimizer best{0, 0};
  u32 best_key = 0;
  for (int i = 0; i + w <= k; ++i) {  // i = offset inside the canonical k-mer
    u64 wm = (c >> (2 * i)) & wmask;
    u64 wm_rc = (d >> (2 * (k - w - i))) & wmask;  // the same w-mer read off the other strand
    u64 sym = wm <= wm_rc ? wm : wm_rc;
    u32 key = (mm_hash32(sym, seed) & 0xFFFFFFE0u) | (u32)i;
    if (i == 0 || key < best_key) {
      best_key = key;
      best = Minimizer{sym, fw_canon ? (u64)i : (u64)(k - i - w)};
    }
  }
  return best;
}

// ------------------------------------------------------------------------------------
// simple-sds stand-ins (layout is ours; semantics pinned by wm.rs:521-537,
// pf1/boophf/mod.rs:359-382): IntVector (packed, LSB-first), BitVector with
// rank(i) = #ones in [0,i) and select(i) = position of the i-th one (0-based).
// ------------------------------------------------------------------------------------
struct IntVector {
  u64 width = 64, len = 0;
  std::vector<u64> words;
  IntVector() {}
  IntVector(u64 n, u64 w) : width(w), len(n), words((n * w + 63) / 64 + 1, 0) {}
  u64 get(u64 i) const {
    u64 bit = i * width, wi = bit >> 6, sh = bit & 63;
    u64 v = words[wi] >> sh;
    if (sh + width > 64) v |= words[wi + 1] << (64 - sh);
    return width == 64 ? v : (v & ((1ULL << width) - 1));
  }
  void set(u64 i, u64 v) {
    u64 bit = i * width, wi = bit >> 6, sh = bit & 63;
    u64 m = width == 64 ? ~0ULL : ((1ULL << width) - 1);
    v &= m;
    words[wi] = (words[wi] & ~(m << sh)) | (v << sh);
    if (sh + width > 64) {
      u64 hi = sh + width - 64;
      u64 mh = (1ULL << hi) - 1;
      words[wi + 1] = (words[wi + 1] & ~mh) | (v >> (64 - sh));
    }
  }
  // IntVector::from(Vec<usize>) + pack(): width = bit length of the maximum (>= 1)
  static IntVector packed(const std::vector<u64>& xs) {
    u64 mx = 0;
    for (u64 x : xs) mx = std::max(mx, x);
    u64 w = mx == 0 ? 1 : msb(mx) + 1;
    IntVector iv(xs.size(), w);
    for (u64 i = 0; i < xs.size(); ++i) iv.set(i, xs[i]);
    return iv;
  }
};

struct BitVector {
  u64 len = 0;
  std::vector<u64> words;
  std::vector<u64> cum;  // #ones before word i
  BitVector() {}
  explicit BitVector(u64 n) : len(n), words((n + 63) / 64 + 1, 0) {}
  void set_bit(u64 i) { words[i >> 6] |= 1ULL << (i & 63); }
  bool bit(u64 i) const { return (words[i >> 6] >> (i & 63)) & 1; }
  void enable_rank() {
    cum.assign(words.size() + 1, 0);
    for (size_t i = 0; i < words.size(); ++i) cum[i + 1] = cum[i] + (u64)__builtin_popcountll(words[i]);
  }
  u64 count_ones() const { return cum.back(); }
  u64 rank(u64 i) const {  // #ones in [0, i)
    u64 wi = i >> 6, off = i & 63;
    u64 r = cum[wi];
    if (off) r += (u64)__builtin_popcountll(words[wi] & ((1ULL << off) - 1));
    return r;
  }
  // position of the i-th (0-based) one
  u64 select(u64 i) const {
    if (i >= count_ones()) throw OracleError("select out of range");
    size_t wi = std::upper_bound(cum.begin(), cum.end(), i) - cum.begin() - 1;
    u64 r = i - cum[wi];
    u64 w = words[wi];
    for (u64 j = 0; j < r; ++j) w &= w - 1;
    return (u64)wi * 64 + (u64)__builtin_ctzll(w);
  }
  // RawVector::int(pos, n): n (<=64) bits starting at bit pos
  u64 int_at(u64 pos, u64 n) const {
    if (n == 0) return 0;
    u64 wi = pos >> 6, sh = pos & 63;
    u64 v = words[wi] >> sh;
    if (sh + n > 64 && wi + 1 < words.size()) v |= words[wi + 1] << (64 - sh);
    return n == 64 ? v : (v & ((1ULL << n) - 1));
  }
};

// kmers::SeqVector: 2-bit bases, base i at bits [2i,2i+2) LSB-first
struct SeqVector {
  u64 len = 0;  // bases
  std::vector<u64> words;
  void ensure(u64 nbases) {
    u64 need = (2 * nbases + 63) / 64 + 2;  // +2 pad words: get_kmer_u64 may read past the end (unitig_set.rs:226-229)
    if (words.size() < need) words.resize(need, 0);
  }
  void push_chars(const u8* s, u64 n) {
    ensure(len + n);
    for (u64 i = 0; i < n; ++i) {
      int c = base_code(s[i]);
      if (c < 0) throw OracleError("non-ACGT base in unitig sequence");
      u64 bit = 2 * (len + i);
      words[bit >> 6] |= (u64)c << (bit & 63);
    }
    len += n;
  }
  u64 get_kmer_u64(u64 pos, int k) const {
    u64 bit = 2 * pos, wi = bit >> 6, sh = bit & 63;
    u64 v = words[wi] >> sh;
    if (sh) v |= words[wi + 1] << (64 - sh);
    return v & kmer_mask(k);
  }
  u64 get_base(u64 pos) const { return (words[(2 * pos) >> 6] >> ((2 * pos) & 63)) & 3; }
};

// ------------------------------------------------------------------------------------
// elias_fano.rs:55-122
// ------------------------------------------------------------------------------------
struct EFVector {
  u64 u = 0, l = 0, n = 0;
  BitVector high_bits;
  IntVector low_bits;
  static EFVector from_slice(const std::vector<u64>& xs) {  // elias_fano.rs:124-140
    if (xs.empty()) throw OracleError("EFEmpty");
    return from_iter(xs, xs.back());
  }
  static EFVector from_iter(const std::vector<u64>& xs, u64 u) {  // elias_fano.rs:55-114
    EFVector ef;
    u64 n = xs.size();
    ef.n = n;
    ef.u = u;
    u64 l = msb(u / n);
    if (l == 0) l = 1;  // elias_fano.rs:64-75 "hack to avoid l == 0"
    ef.l = l;
    u64 low_mask = (1ULL << l) - 1;
    ef.low_bits = IntVector(n, l);
    u64 hb_len = n + (u >> l);
    ef.high_bits = BitVector(hb_len + 1);
    u64 last = 0, h0s = 0;
    for (u64 i = 0; i < n; ++i) {
      u64 x = xs[i];
      if (x < last) throw OracleError("EFNotMonotone");
      ef.low_bits.set(i, x & low_mask);
      u64 gap = (x >> l) - (last >> l);
      ef.high_bits.set_bit(h0s + i + gap);
      last = x;
      h0s += gap;
    }
    ef.high_bits.enable_rank();
    return ef;
  }
  u64 len() const { return n; }
  u64 get(u64 i) const {  // elias_fano.rs:116-122
    u64 high = high_bits.select(i) - i;
    return (high << l) | low_bits.get(i);
  }
};

// ------------------------------------------------------------------------------------
// unitig_set.rs:31-250
// ------------------------------------------------------------------------------------
struct UnitigSet {
  int k = 0;
  SeqVector useq;
  EFVector accum_lens;
  BitVector bv;

  static UnitigSet from_accum(int k, SeqVector useq, const std::vector<u64>& accum) {
    UnitigSet us;
    us.k = k;
    us.useq = std::move(useq);
    u64 total = accum.back();
    us.bv = BitVector(total);
    for (size_t i = 1; i < accum.size(); ++i) us.bv.set_bit(accum[i] - 1);  // unitig_set.rs:90-93,146-149
    us.bv.enable_rank();
    us.accum_lens = EFVector::from_slice(accum);
    return us;
  }
  static UnitigSet from_seqs(const std::vector<std::string>& seqs, int k) {  // unitig_set.rs:74-106
    SeqVector useq;
    std::vector<u64> accum;
    u64 ps = 0;
    for (auto& s : seqs) {
      useq.push_chars((const u8*)s.data(), s.size());
      accum.push_back(ps);
      ps += s.size();
    }
    accum.push_back(ps);
    return from_accum(k, std::move(useq), accum);
  }
  u64 n_unitigs() const { return accum_lens.len() - 1; }                                   // :168-170
  u64 unitig_len(u64 i) const { return accum_lens.get(i + 1) - accum_lens.get(i); }         // :178-180
  u64 pos_to_id(u64 pos) const { return bv.rank(pos); }                                     // :185-187
  u64 unitig_start_pos(u64 i) const { return accum_lens.get(i); }                           // :197-199
  u64 unitig_end_pos(u64 i) const { return accum_lens.get(i + 1); }                         // :202-204
  u64 total_len() const { return accum_lens.get(n_unitigs()); }                             // :207-209
  u64 n_kmers() const { return total_len() - (u64)k * n_unitigs() + n_unitigs(); }          // :212-214
  u64 get_kmer_u64_from_useq_pos(u64 pos) const { return useq.get_kmer_u64(pos, k); }       // :226-229
  bool is_valid_useq_pos(u64 pos) const {                                                  // :235-245
    u64 last = total_len() - (u64)k;
    if (pos > last) return false;
    return bv.int_at(pos, (u64)k - 1) == 0;
  }
};

// ------------------------------------------------------------------------------------
// K2UPos (kphf/mod.rs:13-19) and the K2U trait (kphf/mod.rs:58-67)
// ------------------------------------------------------------------------------------
struct K2UPos {
  u64 unitig_id = USIZE_MAX, unitig_len = USIZE_MAX, pos = USIZE_MAX;
  MatchType o = NoMatch;
  bool operator==(const K2UPos& b) const {
    return unitig_id == b.unitig_id && unitig_len == b.unitig_len && pos == b.pos && o == b.o;
  }
};
struct K2U {
  virtual ~K2U() {}
  virtual const UnitigSet& unitigs() const = 0;
  virtual bool k2u(const CanonicalKmer& km, K2UPos& out) const = 0;  // Option<K2UPos>: false == None
  int k() const { return unitigs().k; }
  void check_k(const CanonicalKmer& km) const {
    // contract panic: kphf/pfhash.rs:109, kphf/sshash.rs:473, index.rs:157-163
    if (km.len() != k()) throw OracleError("Got query k-mer size k=" + std::to_string(km.len()) + ", expected k=" + std::to_string(k()));
  }
};

// ------------------------------------------------------------------------------------
// pf1/boophf/hash.rs:9-135 -- SimpleHash (SingleHashFunctor<uint64_t>) and the
// XorshiftHashFunctors chain; golden-pinned (hash.rs:152-253).
// ------------------------------------------------------------------------------------
static const u64 HASH_PAIR_SEED0 = 0xAAAAAAAA55555555ULL, HASH_PAIR_SEED1 = 0x33333333CCCCCCCCULL;
inline u64 simple_hash64(u64 key, u64 seed) {  // hash.rs:33-49
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
struct MultiHashState {
  u64 s0 = HASH_PAIR_SEED0, s1 = HASH_PAIR_SEED1;
};
inline u64 multihash_h0(MultiHashState& st, u64 key) {  // hash.rs:111-115
  u64 h = simple_hash64(key, HASH_PAIR_SEED0);
  st.s0 = h;
  st.s1 = HASH_PAIR_SEED1;
  return h;
}
inline u64 multihash_h1(MultiHashState& st, u64 key) {  // hash.rs:117-121
  u64 h = simple_hash64(key, HASH_PAIR_SEED1);
  st.s1 = h;
  return h;
}
inline u64 multihash_next(MultiHashState& st) {  // hash.rs:123-135
  u64 s1 = st.s0, s0 = st.s1;
  s1 ^= s1 << 23;
  s1 = s1 ^ s0 ^ (s1 >> 17) ^ (s0 >> 26);
  u64 hash = s1 + s0;
  st.s0 = s0;
  st.s1 = s1;
  return hash;
}

// ------------------------------------------------------------------------------------
// pf1/boophf/mod.rs:23-293 -- load-only C++ BooPHF
// ------------------------------------------------------------------------------------
struct Reader {
  std::vector<u8> buf;
  size_t p = 0;
  explicit Reader(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw OracleError("cannot open " + path);
    buf.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
  }
  template <class T>
  T rd() {
    if (p + sizeof(T) > buf.size()) throw OracleError("unexpected EOF");
    T v;
    memcpy(&v, &buf[p], sizeof(T));
    p += sizeof(T);
    return v;
  }
  std::vector<u64> rd_u64s(u64 n) {
    std::vector<u64> v(n);
    if (p + 8 * n > buf.size()) throw OracleError("unexpected EOF");
    memcpy(v.data(), &buf[p], 8 * n);
    p += 8 * n;
    return v;
  }
  size_t remaining() const { return buf.size() - p; }
};

struct BoophfBitVec {  // boophf/mod.rs:183-293
  u64 n_bits = 0;
  std::vector<u64> data;
  std::vector<u64> ranks;
  bool bit(u64 pos) const { return (data[pos >> 6] >> (pos & 63)) & 1; }
  u64 rank(u64 pos) const {  // boophf/mod.rs:250-266
    u64 word_idx = pos / 64, word_offset = pos % 64, block = pos / 512;
    u64 r = ranks[block];
    for (u64 i = block * 512 / 64; i < word_idx; ++i) r += (u64)__builtin_popcountll(data[i]);
    u64 mask = (1ULL << word_offset) - 1;
    r += (u64)__builtin_popcountll(data[word_idx] & mask);
    return r;
  }
};
struct BooPHF {
  double gamma = 0;
  u64 last_bitset_rank = 0, n_elem = 0;
  std::vector<BoophfBitVec> levels;
  std::unordered_map<u64, u64> final_hash;

  static BooPHF load(const std::string& path) {  // boophf/mod.rs:50-86,269-293
    Reader r(path);
    BooPHF m;
    m.gamma = r.rd<double>();
    int nb_levels = r.rd<int32_t>();
    m.last_bitset_rank = r.rd<u64>();
    m.n_elem = r.rd<u64>();
    for (int i = 0; i < nb_levels; ++i) {
      BoophfBitVec bv;
      bv.n_bits = r.rd<u64>();
      u64 n_words = r.rd<u64>();
      bv.data = r.rd_u64s(n_words);
      u64 rs = r.rd<u64>();
      bv.ranks = r.rd_u64s(rs);
      m.levels.push_back(std::move(bv));
    }
    u64 fh = r.rd<u64>();
    for (u64 i = 0; i < fh; ++i) {
      u64 k = r.rd<u64>();
      u64 v = r.rd<u64>();
      m.final_hash[k] = v;
    }
    return m;
  }
  static u64 fast_range_64(u64 word, u64 p) { return (u64)(((__uint128_t)word * p) >> 64); }  // :136-144
  bool lookup_in_levels(u64 item, u64& out) const {  // :146-175
    MultiHashState st;
    for (size_t li = 0; li < levels.size(); ++li) {
      u64 h = li == 0 ? multihash_h0(st, item) : li == 1 ? multihash_h1(st, item) : multihash_next(st);
      u64 pos = fast_range_64(h, levels[li].n_bits);
      if (levels[li].bit(pos)) {
        out = levels[li].rank(pos);
        return true;
      }
    }
    return false;
  }
  bool lookup(u64 item, u64& out) const {  // :96-102,177-181
    if (lookup_in_levels(item, out)) return true;
    auto it = final_hash.find(item);
    if (it == final_hash.end()) return false;
    out = it->second + last_bitset_rank;
    return true;
  }
};

// ------------------------------------------------------------------------------------
// Stand-in for boomphf 0.5.9 `Mphf<u64>` (PARITY UNPINNED; results are invariant to the
// MPHF).  Plain serial BBHash: level bitsets of gamma*n slots, collision-free
// keys placed, colliders cascade; try_hash = first level whose bit is set, rank over all
// levels.  Like boomphf::try_hash it may return a false-positive slot for non-members.
// ------------------------------------------------------------------------------------
struct OracleMphf {
  struct Level {
    u64 n_bits = 0;
    BitVector bv;
    u64 rank_base = 0;
  };
  std::vector<Level> levels;
  std::map<u64, u64> leftovers;  // keys that survived all levels (practically never)
  u64 n = 0;
  static u64 level_hash(u64 key, u64 level) {
    u64 x = key + 0x9E3779B97F4A7C15ULL * (level + 1);
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return x;
  }
  static OracleMphf build(std::vector<u64> keys, double gamma) {
    OracleMphf m;
    m.n = keys.size();
    u64 rank_base = 0;
    for (u64 lvl = 0; lvl < 40 && !keys.empty(); ++lvl) {
      Level L;
      L.n_bits = std::max<u64>(64, (u64)(gamma * (double)keys.size()));
      BitVector seen(L.n_bits), coll(L.n_bits);
      for (u64 key : keys) {
        u64 s = (u64)(((__uint128_t)level_hash(key, lvl) * L.n_bits) >> 64);
        if (seen.bit(s)) coll.set_bit(s); else seen.set_bit(s);
      }
      L.bv = BitVector(L.n_bits);
      std::vector<u64> next;
      for (u64 key : keys) {
        u64 s = (u64)(((__uint128_t)level_hash(key, lvl) * L.n_bits) >> 64);
        if (coll.bit(s)) next.push_back(key); else L.bv.set_bit(s);
      }
      L.bv.enable_rank();
      L.rank_base = rank_base;
      rank_base += L.bv.count_ones();
      m.levels.push_back(std::move(L));
      keys.swap(next);
    }
    std::sort(keys.begin(), keys.end());
    for (u64 key : keys) m.leftovers[key] = rank_base++;
    return m;
  }
  bool try_hash(u64 key, u64& out) const {
    for (u64 lvl = 0; lvl < levels.size(); ++lvl) {
      const Level& L = levels[lvl];
      u64 s = (u64)(((__uint128_t)level_hash(key, lvl) * L.n_bits) >> 64);
      if (L.bv.bit(s)) {
        out = L.rank_base + L.bv.rank(s);
        return true;
      }
    }
    auto it = leftovers.find(key);
    if (it == leftovers.end()) return false;
    out = it->second;
    return true;
  }
  u64 hash(u64 key) const {
    u64 h;
    if (!try_hash(key, h)) throw OracleError("mphf.hash on non-member");
    return h;
  }
};

// ------------------------------------------------------------------------------------
// kphf/pfhash.rs:20-135 -- PFHash<MPHF>
// ------------------------------------------------------------------------------------
template <class MPHF>
struct PFHash : K2U {
  UnitigSet us;
  MPHF mphf;
  IntVector pos;
  const UnitigSet& unitigs() const override { return us; }
  bool mphf_try(u64 w, u64& h) const;
  bool k2u(const CanonicalKmer& km, K2UPos& out) const override {  // pfhash.rs:108-134
    check_k(km);
    u64 word = km.canonical_word();
    u64 h;
    if (!mphf_try(word, h)) return false;
    if (h >= pos.len) return false;  // (guard only reachable through an MPHF false positive beyond n; boomphf/BooPHF never exceed n)
    u64 km_pos = pos.get(h);
    u64 kw = us.get_kmer_u64_from_useq_pos(km_pos);
    MatchType mt = km.word_equivalency(kw);
    if (mt == NoMatch) return false;
    u64 uid = us.pos_to_id(km_pos);
    out.unitig_id = uid;
    out.unitig_len = us.unitig_len(uid);
    out.pos = km_pos - us.unitig_start_pos(uid);
    out.o = mt;
    return true;
  }
};
template <>
inline bool PFHash<BooPHF>::mphf_try(u64 w, u64& h) const { return mphf.lookup(w, h); }
template <>
inline bool PFHash<OracleMphf>::mphf_try(u64 w, u64& h) const { return mphf.try_hash(w, h); }

// pfhash.rs:40-73 PFHash::from_unitig_set
inline std::unique_ptr<PFHash<OracleMphf>> pfhash_from_unitig_set(UnitigSet us) {
  auto h = std::make_unique<PFHash<OracleMphf>>();
  int k = us.k;
  std::vector<u64> keys;
  keys.reserve(us.n_kmers());
  u64 U = us.n_unitigs();
  for (u64 ui = 0; ui < U; ++ui) {
    u64 s = us.unitig_start_pos(ui), e = us.unitig_end_pos(ui);
    for (u64 p = s; p + k <= e; ++p) {
      u64 fw = us.useq.get_kmer_u64(p, k);
      keys.push_back(std::min(fw, revcomp_word(fw, k)));
    }
  }
  h->mphf = OracleMphf::build(keys, 1.7);
  std::vector<u64> pos(us.n_kmers(), USIZE_MAX);
  size_t idx = 0;
  for (u64 ui = 0; ui < U; ++ui) {
    u64 s = us.unitig_start_pos(ui), e = us.unitig_end_pos(ui);
    for (u64 p = s; p + k <= e; ++p) pos[h->mphf.hash(keys[idx++])] = p;
  }
  h->pos = IntVector::packed(pos);
  h->us = std::move(us);
  return h;
}

// ------------------------------------------------------------------------------------
// kphf/sshash.rs:20-555 -- SSHashBuilder::from_unitig_set, finish, SSHash::k2u, skew index
// ------------------------------------------------------------------------------------
struct MinimizerOcc {
  u64 word;
  u64 pos;  // position of the minimizer w-mer on the forward strand of the concatenated useq
  bool operator==(const MinimizerOcc& b) const { return word == b.word && pos == b.pos; }
};

struct SSHash : K2U {
  u64 seed = 0;
  int w = 0;
  UnitigSet us;
  OracleMphf mphf;
  EFVector occs_prefix_sum;
  IntVector pos;
  u64 skew_param = USIZE_MAX;
  bool has_skew = false;
  OracleMphf skew_mphf;
  IntVector skew_pos;
  u64 n_minimizer_occs = 0;

  const UnitigSet& unitigs() const override { return us; }
  u64 n_minimizers() const { return occs_prefix_sum.len(); }  // sshash.rs:333-335 (sic: len of the prefix sum)
  u64 n_kmers_in_skew_index() const { return has_skew ? skew_pos.len : 0; }

  // Restates `SeqVectorSlice::iter_canonical_minimizers(k,w,bh)` of one unitig + the
  // fw / rc stream split and run-length dedup of sshash.rs:100-143.
  static void collect_unitig(const UnitigSet& us, u64 ui, int w, u64 seed, std::vector<MinimizerOcc>& out) {
    int k = us.k;
    u64 s = us.unitig_start_pos(ui), e = us.unitig_end_pos(ui);
    // the fw-canonical stream goes straight to `out`, the rc-canonical stream is buffered and appended after it: the same
    // two run-length-deduped streams, in the same order, as walking the unitig twice
    std::vector<MinimizerOcc> rc_stream;
    bool have_prev[2] = {false, false};
    MinimizerOcc prev[2] = {{0, 0}, {0, 0}};
    for (u64 p = s; p + k <= e; ++p) {
      u64 fw = us.useq.get_kmer_u64(p, k);
      u64 rc = revcomp_word(fw, k);
      const int pass = fw <= rc ? 0 : 1;
      Minimizer mm = canonical_minimizer(fw, k, w, seed);
      // position of the minimizer occurrence on the forward strand (offset is in fw-mer coordinates)
      MinimizerOcc cur{mm.word, p + mm.offset};
      if (!have_prev[pass] || !(cur == prev[pass])) (pass == 0 ? out : rc_stream).push_back(cur);
      prev[pass] = cur;
      have_prev[pass] = true;
    }
    out.insert(out.end(), rc_stream.begin(), rc_stream.end());
  }

  static std::unique_ptr<SSHash> from_unitig_set(UnitigSet us_in, int w, u64 skew_param, u64 seed) {
    auto H = std::make_unique<SSHash>();
    SSHash& h = *H;
    h.us = std::move(us_in);
    const UnitigSet& us = h.us;
    int k = us.k;
    if (w > k) throw OracleError("assert w <= k");  // sshash.rs:92
    h.w = w;
    h.seed = seed;
    h.skew_param = skew_param;
    BuildTimer timer;
    // 1. collect minimizers (sshash.rs:97-143)
    std::vector<MinimizerOcc> minimizers;
    {
      std::vector<std::vector<MinimizerOcc>> parts((size_t)std::max(1, build_threads()));
      parallel_blocks(us.n_unitigs(), [&](int t, u64 lo, u64 hi) {
        for (u64 ui = lo; ui < hi; ++ui) collect_unitig(us, ui, w, seed, parts[t]);
      });
      for (auto& part : parts) {  // blocks are contiguous unitig ranges: concatenation keeps the unitig order
        minimizers.insert(minimizers.end(), part.begin(), part.end());
        std::vector<MinimizerOcc>().swap(part);
      }
    }
    timer.lap("collect");
    // 2. sort (stable, as rayon par_sort_by_key) and group (sshash.rs:150-172)
    parallel_stable_sort(minimizers, [](const MinimizerOcc& a, const MinimizerOcc& b) { return a.word < b.word; });
    std::vector<u64> mm_occs, mm_set;
    {
      u64 cur = minimizers[0].word, occs = 0;
      for (auto& mm : minimizers) {
        if (cur != mm.word) {
          mm_occs.push_back(occs);
          mm_set.push_back(cur);
          cur = mm.word;
          occs = 0;
        }
        ++occs;
      }
      mm_occs.push_back(occs);
      mm_set.push_back(cur);
    }
    timer.lap("sort + group");
    // 3. MPHF over minimizers (sshash.rs:177)
    h.mphf = OracleMphf::build(mm_set, 1.7);
    timer.lap("mphf");
    // 4. bucket sizes in MPHF order -> prefix sum (sshash.rs:181-189)
    std::vector<u64> n_occs(mm_set.size(), USIZE_MAX), mm_hash(mm_set.size());
    parallel_blocks(mm_set.size(), [&](int, u64 lo, u64 hi) {  // the MPHF is a bijection: every thread writes its own slots
      for (u64 i = lo; i < hi; ++i) {
        mm_hash[i] = h.mphf.hash(mm_set[i]);
        n_occs[mm_hash[i]] = mm_occs[i];
      }
    });
    std::vector<u64> occs_prefix_sum = prefix_sum(n_occs);
    // 5. scatter positions (sshash.rs:196-219; the reference scatters through an UnsafeSlice from a rayon loop)
    std::vector<u64> pos(minimizers.size(), USIZE_MAX);
    std::vector<u64> ranges = prefix_sum(mm_occs);
    parallel_blocks(mm_set.size(), [&](int, u64 lo, u64 hi) {
      for (u64 i = lo; i < hi; ++i) {
        u64 sh = occs_prefix_sum[mm_hash[i]];
        for (u64 j = ranges[i]; j < ranges[i + 1]; ++j) pos[sh + (j - ranges[i])] = minimizers[j].pos;
      }
    });
    timer.lap("bucket sizes + scatter");
    // 6. skew index (sshash.rs:222-296)
    if (skew_param != USIZE_MAX) {
      std::vector<std::pair<u64, u64>> skew_tuples;
      for (size_t i = 0; i < mm_occs.size(); ++i) {
        if (mm_occs[i] <= skew_param) continue;  // sshash.rs:232
        for (u64 j = ranges[i]; j < ranges[i + 1]; ++j) {
          u64 mp = minimizers[j].pos;
          u64 start_pos = mp < (u64)(k - w) ? 0 : mp - (u64)(k - w);  // sshash.rs:241-247
          u64 n_kmers = (u64)(k - w + 1);
          for (u64 off = 0; off < n_kmers; ++off) {
            u64 p = start_pos + off;
            if (us.is_valid_useq_pos(p)) {
              u64 fw = us.get_kmer_u64_from_useq_pos(p);
              skew_tuples.push_back({std::min(fw, revcomp_word(fw, k)), p});
            }
          }
        }
      }
      parallel_stable_sort(skew_tuples, [](const std::pair<u64, u64>& a, const std::pair<u64, u64>& b) { return a.first < b.first; });
      // dedup_by_key keeps the first of each run (sshash.rs:273-274)
      std::vector<std::pair<u64, u64>> ded;
      for (auto& t : skew_tuples)
        if (ded.empty() || ded.back().first != t.first) ded.push_back(t);
      std::vector<u64> km_set;
      for (auto& t : ded) km_set.push_back(t.first);
      h.has_skew = true;
      h.skew_mphf = OracleMphf::build(km_set, 1.7);
      std::vector<u64> sp(km_set.size(), 0);
      for (auto& t : ded) sp[h.skew_mphf.hash(t.first)] = t.second;
      h.skew_pos = IntVector::packed(sp);
    }
    timer.lap("skew index");
    // finish (sshash.rs:310-329)
    h.occs_prefix_sum = EFVector::from_slice(occs_prefix_sum);
    h.pos = IntVector::packed(pos);
    h.n_minimizer_occs = pos.size();
    timer.lap("elias-fano + packing");
    return H;
  }

  bool make_hit(u64 km_pos, MatchType mt, bool boundary_check, K2UPos& out) const {
    u64 uid = us.pos_to_id(km_pos);
    u64 ulen = us.unitig_len(uid);
    u64 pos_in = km_pos - us.unitig_start_pos(uid);
    if (boundary_check) {
      u64 end_pos = km_pos + (u64)us.k;  // sshash.rs:513-514,539-541
      if (end_pos > us.unitig_end_pos(uid)) return false;
    }
    out.unitig_id = uid;
    out.unitig_len = ulen;
    out.pos = pos_in;
    out.o = mt;
    return true;
  }

  bool k2u_skew_index(const CanonicalKmer& km, K2UPos& out) const {  // sshash.rs:415-433, 59-66
    if (!has_skew) return false;
    u64 hh;
    if (!skew_mphf.try_hash(km.canonical_word(), hh)) return false;
    if (hh >= skew_pos.len) return false;
    u64 p = skew_pos.get(hh);
    u64 kw = us.get_kmer_u64_from_useq_pos(p);
    MatchType mt = km.word_equivalency(kw);
    if (mt == NoMatch) return false;
    return make_hit(p, mt, false, out);
  }

  bool k2u(const CanonicalKmer& km, K2UPos& out) const override {  // sshash.rs:471-555
    check_k(km);
    int k = us.k;
    Minimizer mm = canonical_minimizer(km.fw, k, w, seed);
    u64 offset = mm.offset;
    u64 hh;
    if (!mphf.try_hash(mm.word, hh)) return false;
    if (hh + 1 >= occs_prefix_sum.len()) return false;  // unreachable for a valid MPHF (hash < n)
    u64 pos_start = occs_prefix_sum.get(hh);
    u64 pos_end = occs_prefix_sum.get(hh + 1);
    u64 n_occs = pos_end - pos_start;
    if (n_occs > skew_param) return k2u_skew_index(km, out);  // sshash.rs:486-490
    u64 last_km_start_pos = us.total_len() - (u64)k;
    u64 rc_offset = (u64)k - offset - (u64)w;
    for (u64 pi = pos_start; pi < pos_end; ++pi) {
      u64 mm_pos = pos.get(pi);
      if (mm_pos >= offset && (mm_pos - offset) <= last_km_start_pos) {  // sshash.rs:498
        u64 km_pos = mm_pos - offset;
        MatchType mt = km.word_equivalency(us.get_kmer_u64_from_useq_pos(km_pos));
        if (mt != NoMatch && make_hit(km_pos, mt, true, out)) return true;
      }
      if (mm_pos >= rc_offset && (mm_pos - rc_offset) <= last_km_start_pos) {  // sshash.rs:527
        u64 km_pos = mm_pos - rc_offset;
        MatchType mt = km.word_equivalency(us.get_kmer_u64_from_useq_pos(km_pos));
        if (mt != NoMatch && make_hit(km_pos, mt, true, out)) return true;
      }
    }
    return false;
  }

  // sshash.rs:563-624 -- forward-orientation-only lookup kept for the reference's unit tests
  bool k2u_fw(u64 fw_word, int qk, K2UPos& out) const {
    int k = us.k;
    if (qk != k) throw OracleError("assert km.k == self.k()");
    Minimizer mm = canonical_minimizer(fw_word, k, w, seed);
    u64 hh;
    if (!mphf.try_hash(mm.word, hh)) return false;
    if (hh + 1 >= occs_prefix_sum.len()) return false;
    u64 pos_start = occs_prefix_sum.get(hh), pos_end = occs_prefix_sum.get(hh + 1);
    for (u64 pi = pos_start; pi < pos_end; ++pi) {
      u64 mm_pos = pos.get(pi);
      if (mm_pos < mm.offset) continue;
      u64 km_pos = mm_pos - mm.offset;
      if (km_pos > us.total_len() - (u64)k) continue;
      if (us.get_kmer_u64_from_useq_pos(km_pos) == (fw_word & kmer_mask(k))) {
        if (make_hit(km_pos, IdentityMatch, true, out)) return true;
      }
    }
    return false;
  }
};

// ------------------------------------------------------------------------------------
// index/caching.rs:13-103 -- StreamingK2U
// ------------------------------------------------------------------------------------
struct StreamingK2U {
  bool is_warm = false;
  K2UPos prev;
  const K2U* k2u;
  explicit StreamingK2U(const K2U* h) : k2u(h) {}
  void reset() {
    is_warm = false;
    prev = K2UPos();
  }
  bool k2u_streaming(const CanonicalKmer& km, K2UPos& out) {  // caching.rs:65-71
    return is_warm ? k2u_warm(km, out) : k2u_cold(km, out);
  }
  bool k2u_warm(const CanonicalKmer& km, K2UPos& out) {  // caching.rs:73-97
    const UnitigSet& us = k2u->unitigs();
    u64 k = (u64)us.k;
    u64 next_pos = prev.pos + 1;
    u64 last_km_pos = prev.unitig_len - k;
    if (next_pos > last_km_pos) return k2u_cold(km, out);
    u64 kw = us.useq.get_kmer_u64(us.unitig_start_pos(prev.unitig_id) + next_pos, (int)k);
    MatchType mt = km.word_equivalency(kw);
    if (mt == NoMatch) return k2u_cold(km, out);
    prev.pos = next_pos;
    prev.o = mt;
    out = prev;
    return true;
  }
  bool k2u_cold(const CanonicalKmer& km, K2UPos& out) {  // caching.rs:99-103
    K2UPos r;
    if (!k2u->k2u(km, r)) return false;  // `?` returns before touching prev / is_warm
    prev = r;
    is_warm = true;
    out = prev;
    return true;
  }
};

// ------------------------------------------------------------------------------------
// lib.rs:36-85 Orientation (Forward -> 1, Backward -> 0); index.rs:304-346 UnitigOcc
// ------------------------------------------------------------------------------------
struct UnitigOcc {
  u64 ref_id, pos;
  u32 fw;  // 1 = Forward, 0 = Backward
  bool operator==(const UnitigOcc& b) const { return ref_id == b.ref_id && pos == b.pos && fw == b.fw; }
};
inline u64 encode_pf1(const UnitigOcc& o) {  // index.rs:320-332
  u64 word = o.pos;
  if (o.fw) word |= 0x80000000ULL;
  word <<= 32;
  word |= o.ref_id;
  return word;
}
inline UnitigOcc decode_pf1(u64 word) {  // index.rs:335-346
  UnitigOcc o;
  o.ref_id = word & 0xFFFFFFFFULL;
  u64 pw = word >> 32;
  o.pos = pw & 0x7FFFFFFFULL;
  o.fw = (pw & 0x80000000ULL) ? 1 : 0;
  return o;
}
inline u64 encode_piscem(const UnitigOcc& o, u64 ref_shift) {  // spt_compact.rs:90-97
  u64 e = o.ref_id;
  e <<= ref_shift;
  e |= o.pos << 1;
  e |= o.fw ? 1 : 0;
  return e;
}
inline UnitigOcc decode_piscem(u64 ref_shift, u64 pos_mask, u64 enc) {  // spt_compact.rs:99-110
  UnitigOcc o;
  o.ref_id = enc >> ref_shift;
  o.pos = (enc >> 1) & pos_mask;
  o.fw = (enc & 1) ? 1 : 0;
  return o;
}
// spt_compact.rs:221-242 required_num_bits -> (pos_bits, ref_bits, total_bits)
inline void required_num_bits(u64 longest_ref, u64 num_refs, u32& pos_bits, u32& ref_bits, u32& total_bits) {
  if (longest_ref == 0 || num_refs == 0) throw OracleError("Could not obtain log2");
  pos_bits = (u32)msb(longest_ref) + 1;
  ref_bits = (u32)msb(num_refs) + 1;
  total_bits = 1 + ref_bits + pos_bits;
  if (total_bits > 255) throw OracleError("Could not compute number of bits required for each occ entry");
}

struct MappedRefPos {
  u64 ref_id, pos;
  u32 fw;
  bool operator==(const MappedRefPos& b) const { return ref_id == b.ref_id && pos == b.pos && fw == b.fw; }
};
// index.rs:194-216 project_onto_u_occ
inline MappedRefPos project_onto_u_occ(u64 k, const K2UPos& h, const UnitigOcc& occ) {
  MappedRefPos m;
  m.ref_id = occ.ref_id;
  m.pos = occ.fw ? h.pos + occ.pos : occ.pos + (h.unitig_len - h.pos) - k;
  u32 o = h.o == IdentityMatch ? 1 : 0;
  m.fw = occ.fw ? o : (o ^ 1);
  return m;
}

// ------------------------------------------------------------------------------------
// index/dense_unitig_table.rs:13-153 -- U2Pos tables
// ------------------------------------------------------------------------------------
struct U2Pos {
  virtual ~U2Pos() {}
  virtual void encoded_unitig_occs(u64 ui, u64& start, u64& end) const = 0;
  virtual UnitigOcc decode_at(u64 i) const = 0;
  virtual u64 n_total_occs() const = 0;
  virtual u64 n_unitigs() const = 0;
  std::vector<UnitigOcc> decode_unitig_occs(u64 ui) const {
    u64 s, e;
    encoded_unitig_occs(ui, s, e);
    std::vector<UnitigOcc> v;
    for (u64 i = s; i < e; ++i) v.push_back(decode_at(i));
    return v;
  }
};
struct DenseUnitigTable : U2Pos {  // dense_unitig_table.rs:13-76
  std::vector<u64> ctable;
  IntVector contig_offsets;
  std::vector<std::string> ref_names;
  std::vector<u32> ref_exts;
  void encoded_unitig_occs(u64 ui, u64& s, u64& e) const override {
    s = contig_offsets.get(ui);
    e = contig_offsets.get(ui + 1);
  }
  UnitigOcc decode_at(u64 i) const override { return decode_pf1(ctable[i]); }
  u64 n_total_occs() const override { return ctable.size(); }
  u64 n_unitigs() const override { return contig_offsets.len - 1; }
};
struct PiscemUnitigTable : U2Pos {  // dense_unitig_table.rs:109-153
  u64 ref_shift = 0, pos_mask = 0;
  IntVector ctable;
  IntVector contig_offsets;
  std::vector<std::string> ref_names;
  void encoded_unitig_occs(u64 ui, u64& s, u64& e) const override {
    s = contig_offsets.get(ui);
    e = contig_offsets.get(ui + 1);
  }
  UnitigOcc decode_at(u64 i) const override { return decode_piscem(ref_shift, pos_mask, ctable.get(i)); }
  u64 n_total_occs() const override { return ctable.len; }
  u64 n_unitigs() const override { return contig_offsets.len - 1; }
};

// ------------------------------------------------------------------------------------
// pf1/cpp.rs:124-237 readers: compact vectors and cereal vectors
// ------------------------------------------------------------------------------------
inline IntVector read_compact_vector(const std::string& path) {  // cpp.rs:217-237
  Reader r(path);
  (void)r.rd<u64>();  // static flag
  u64 width = r.rd<u64>();
  if (width == 0) throw OracleError("assert width > 0");
  u64 len = r.rd<u64>();
  (void)r.rd<u64>();  // capacity
  if (r.remaining() % 8) throw OracleError("InvalidData: bytes not divisible by 8");
  IntVector iv;
  iv.width = width;
  iv.len = len;
  iv.words = r.rd_u64s(r.remaining() / 8);
  iv.words.resize(std::max<size_t>(iv.words.size(), (len * width + 63) / 64) + 2, 0);
  return iv;
}
inline std::vector<std::string> read_cereal_strings(Reader& r) {  // cpp.rs:138-154
  u64 n = r.rd<u64>();
  std::vector<std::string> v;
  for (u64 i = 0; i < n; ++i) {
    u64 nc = r.rd<u64>();
    std::string s(nc, '\0');
    for (u64 j = 0; j < nc; ++j) s[j] = (char)r.rd<u8>();
    v.push_back(s);
  }
  return v;
}
inline std::vector<u32> read_cereal_u32s(Reader& r) {  // cpp.rs:156-163
  u64 n = r.rd<u64>();
  std::vector<u32> v(n);
  for (u64 i = 0; i < n; ++i) v[i] = r.rd<u32>();
  return v;
}
inline std::vector<u64> read_cereal_u64s(Reader& r) {  // cpp.rs:165-172
  u64 n = r.rd<u64>();
  return r.rd_u64s(n);
}

// refseq.rs:16-20 + pf1/mod.rs:213-236
struct RefSeqCollection {
  bool has_seq = false;
  SeqVector seq;
  std::vector<u64> prefix;
  u64 n_refs() const { return prefix.size() - 1; }
  u64 ref_len(u64 i) const { return prefix[i + 1] - prefix[i]; }
};

// tiny JSON number extraction, enough for info.json / cuttlefish .json
inline u64 json_u64(const std::string& text, const std::string& key) {
  size_t p = text.find("\"" + key + "\"");
  if (p == std::string::npos) throw OracleError("json key not found: " + key);
  p = text.find(':', p);
  ++p;
  while (p < text.size() && (text[p] == ' ' || text[p] == '\t')) ++p;
  return strtoull(text.c_str() + p, nullptr, 10);
}
inline std::string slurp(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw OracleError("cannot open " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

// ------------------------------------------------------------------------------------
// index.rs:49-216 ModIndex + validate.rs:24-100 + caching.rs:150-218
// ------------------------------------------------------------------------------------
struct ValidateCounts {
  u64 n_queries = 0, n_identity = 0, n_twin = 0, n_projected = 0, n_fail = 0;
};
struct ModIndex {
  std::unique_ptr<K2U> k2u;
  std::shared_ptr<U2Pos> u2pos;
  std::shared_ptr<RefSeqCollection> refs;
  int k() const { return k2u->k(); }

  // index.rs:156-179 get_ref_pos + project_hits (eager)
  bool get_ref_pos_eager(const CanonicalKmer& km, K2UPos& hit, std::vector<MappedRefPos>& out) const {
    if (km.len() != k()) throw OracleError("Got query k-mer size k=" + std::to_string(km.len()) + ", expected k=" + std::to_string(k()));
    if (!k2u->k2u(km, hit)) return false;
    project(hit, out);
    return true;
  }
  void project(const K2UPos& hit, std::vector<MappedRefPos>& out) const {
    out.clear();
    u64 s, e;
    u2pos->encoded_unitig_occs(hit.unitig_id, s, e);
    for (u64 i = s; i < e; ++i) out.push_back(project_onto_u_occ((u64)k(), hit, u2pos->decode_at(i)));
  }
  // validate.rs:24-52 (counts instead of panics)
  ValidateCounts validate_self() const {
    if (!refs || !refs->has_seq) throw OracleError("assert has_refseq");
    ValidateCounts c;
    std::vector<MappedRefPos> mrps;
    int kk = k();
    for (u64 r = 0; r < refs->n_refs(); ++r) {
      u64 s = refs->prefix[r], e = refs->prefix[r + 1];
      for (u64 p = s; p + kk <= e; ++p) {
        CanonicalKmer km = CanonicalKmer::from_u64(refs->seq.get_kmer_u64(p, kk), kk);
        K2UPos hit;
        ++c.n_queries;
        if (!get_ref_pos_eager(km, hit, mrps)) {
          ++c.n_fail;
          continue;
        }
        (hit.o == IdentityMatch ? c.n_identity : c.n_twin)++;
        c.n_projected += mrps.size();
        bool found = false;
        for (auto& m : mrps) found |= (m.pos == p - s) && (m.ref_id == r);
        if (!found) ++c.n_fail;
      }
    }
    return c;
  }
  // validate.rs:54-81 / caching.rs:175-201 validate_ckmers (streaming optional)
  ValidateCounts validate_ckmers(u64 ref_id, const u8* seq, u64 len, StreamingK2U* st) const {
    ValidateCounts c;
    std::vector<MappedRefPos> mrps;
    for_each_canonical_kmer(seq, len, k(), [&](u64 pos, const CanonicalKmer& km) {
      ++c.n_queries;
      K2UPos hit;
      bool ok = st ? st->k2u_streaming(km, hit) : k2u->k2u(km, hit);
      if (!ok) {
        ++c.n_fail;
        return;
      }
      project(hit, mrps);
      (hit.o == IdentityMatch ? c.n_identity : c.n_twin)++;
      c.n_projected += mrps.size();
      bool found = false;
      for (auto& m : mrps) found |= (m.pos == pos) && (m.ref_id == ref_id);
      if (!found) ++c.n_fail;
    });
    return c;
  }
};

// index.rs:363-424 iter_unitigs_on_ref / RefSeqContigIterator: the unitig tiling of one reference
struct RefSeqUnitigOcc {
  u64 unitig_id, unitig_len, pos;
  u32 fw;
};
inline std::vector<RefSeqUnitigOcc> iter_unitigs_on_ref(const ModIndex& idx, u64 ref_id) {
  if (!idx.refs || !idx.refs->has_seq) throw OracleError("Refseq is None");
  std::vector<RefSeqUnitigOcc> out;
  int k = idx.k();
  u64 s = idx.refs->prefix[ref_id], len = idx.refs->ref_len(ref_id);
  u64 end_pos = len + 1 - (u64)k;  // index.rs:368
  u64 pos = 0;
  while (pos < end_pos) {
    CanonicalKmer km = CanonicalKmer::from_u64(idx.refs->seq.get_kmer_u64(s + pos, k), k);
    K2UPos hit;
    if (!idx.k2u->k2u(km, hit)) throw OracleError("called `Option::unwrap()` on a `None` value");  // index.rs:403
    out.push_back(RefSeqUnitigOcc{hit.unitig_id, hit.unitig_len, pos, hit.o == IdentityMatch ? 1u : 0u});
    pos += hit.unitig_len - (u64)k + 1;
  }
  return out;
}

// kphf/mod.rs:69-103 K2U::validate_self (counts instead of panics)
inline ValidateCounts k2u_validate_self(const K2U& h) {
  ValidateCounts c;
  const UnitigSet& us = h.unitigs();
  int k = us.k;
  for (u64 ui = 0; ui < us.n_unitigs(); ++ui) {
    u64 s = us.unitig_start_pos(ui), e = us.unitig_end_pos(ui), ulen = e - s;
    for (u64 p = s; p + k <= e; ++p) {
      CanonicalKmer km = CanonicalKmer::from_u64(us.useq.get_kmer_u64(p, k), k);
      for (int t = 0; t < 2; ++t) {
        K2UPos r;
        ++c.n_queries;
        MatchType want = t == 0 ? IdentityMatch : TwinMatch;
        if (!h.k2u(km, r) || r.unitig_id != ui || r.unitig_len != ulen || r.pos != p - s || r.o != want) ++c.n_fail;
        else (t == 0 ? c.n_identity : c.n_twin)++;
        km.swap();
      }
    }
  }
  return c;
}

// ------------------------------------------------------------------------------------
// pf1/dense_index.rs:33-97 DenseIndex::deserialize_from_cpp, pf1/unitig_table.rs:28-49
// ------------------------------------------------------------------------------------
inline UnitigSet unitig_set_from_pf1(const std::string& dir, int k) {
  IntVector seq = read_compact_vector(dir + "/seq.bin");
  if (seq.width != 2) throw OracleError("seq.bin width != 2");
  IntVector rank = read_compact_vector(dir + "/rank.bin");
  if (rank.width != 1) throw OracleError("rank.bin width != 1");
  SeqVector useq;
  useq.len = seq.len;
  useq.words = seq.words;
  useq.ensure(seq.len);
  // mask stale bits beyond len (capacity padding)
  {
    u64 bits = 2 * seq.len;
    for (u64 wi = (bits + 63) / 64; wi < useq.words.size(); ++wi) useq.words[wi] = 0;
    if (bits & 63) useq.words[bits >> 6] &= (1ULL << (bits & 63)) - 1;
  }
  // dense_index.rs:55-66: select every 1 -> prefix lengths
  std::vector<u64> accum{0};
  for (u64 i = 0; i < rank.len; ++i)
    if (rank.get(i)) accum.push_back(i + 1);
  return UnitigSet::from_accum(k, std::move(useq), accum);
}
inline std::shared_ptr<DenseUnitigTable> dense_unitig_table_from_pf1(const std::string& dir) {
  auto t = std::make_shared<DenseUnitigTable>();
  Reader r(dir + "/ctable.bin");
  t->ref_names = read_cereal_strings(r);
  t->ref_exts = read_cereal_u32s(r);
  t->ctable = read_cereal_u64s(r);
  if (r.remaining() != 0) throw OracleError("ctable.bin: trailing bytes");  // unitig_table.rs:44-46
  t->contig_offsets = read_compact_vector(dir + "/ctg_offsets.bin");
  return t;
}
inline std::shared_ptr<RefSeqCollection> refseq_from_pf1(const std::string& dir) {  // pf1/mod.rs:213-236
  auto rs = std::make_shared<RefSeqCollection>();
  std::ifstream probe(dir + "/refseq.bin", std::ios::binary);
  if (probe.good()) {
    IntVector seq = read_compact_vector(dir + "/refseq.bin");
    rs->has_seq = true;
    rs->seq.len = seq.len;
    rs->seq.words = seq.words;
    rs->seq.ensure(seq.len);
  }
  Reader r(dir + "/refAccumLengths.bin");
  std::vector<u64> acc = read_cereal_u64s(r);
  rs->prefix.push_back(0);
  for (u64 x : acc) rs->prefix.push_back(x);
  return rs;
}
inline std::unique_ptr<ModIndex> dense_index_from_pf1(const std::string& dir) {
  std::string info = slurp(dir + "/info.json");
  int k = (int)json_u64(info, "k");
  auto idx = std::make_unique<ModIndex>();
  auto h = std::make_unique<PFHash<BooPHF>>();
  h->us = unitig_set_from_pf1(dir, k);
  h->mphf = BooPHF::load(dir + "/mphf.bin");
  h->pos = read_compact_vector(dir + "/pos.bin");
  if (h->pos.len != h->us.n_kmers()) throw OracleError("assert pos.len() == unitigs.n_kmers()");  // dense_index.rs:80
  idx->k2u = std::move(h);
  idx->u2pos = dense_unitig_table_from_pf1(dir);
  idx->refs = refseq_from_pf1(dir);
  return idx;
}

// ------------------------------------------------------------------------------------
// kphf/pfhash.rs:137-285 SampledPFHash (pufferfish sparse index, load-only) and its loader
// pf1/sparse_index.rs:32-110.  CanonicalKmer::append_base / prepend_base belong to the `kmers`
// crate (not in tree); their meaning is pinned by validate_self on small_txome_index_sparse
// (sparse_index.rs:163-192): append = drop the first base and add one at the end of the fw
// k-mer, prepend = drop the last base and add one at the front.
// ------------------------------------------------------------------------------------
inline BitVector bitvector_from_compact(const std::string& path) {
  IntVector iv = read_compact_vector(path);
  if (iv.width != 1) throw OracleError(path + ": width != 1");
  BitVector bv(iv.len);
  for (size_t i = 0; i < bv.words.size() && i < iv.words.size(); ++i) bv.words[i] = iv.words[i];
  if (iv.len & 63) bv.words[iv.len >> 6] &= (1ULL << (iv.len & 63)) - 1;
  for (size_t i = (iv.len + 63) / 64; i < bv.words.size(); ++i) bv.words[i] = 0;
  bv.enable_rank();
  return bv;
}
struct SampledPFHash : K2U {
  UnitigSet us;
  BooPHF mphf;
  IntVector sampled_pos, ext_sizes, ext_bases;
  BitVector sampled_vec, canonical_vec, direction_vec;
  u64 sample_size = 0, extension_size = 0;
  const UnitigSet& unitigs() const override { return us; }
  bool k2u_w_pos(const CanonicalKmer& km, u64 pos, K2UPos& out) const {  // pfhash.rs:262-285
    u64 kw = us.get_kmer_u64_from_useq_pos(pos);
    MatchType mt = km.word_equivalency(kw);
    if (mt == NoMatch) return false;
    u64 uid = us.pos_to_id(pos);
    out.unitig_id = uid;
    out.unitig_len = us.unitig_len(uid);
    out.pos = pos - us.unitig_start_pos(uid);
    out.o = mt;
    return true;
  }
  bool k2u(const CanonicalKmer& kmer, K2UPos& out) const override {  // pfhash.rs:190-259
    u64 idx;
    if (!mphf.lookup(kmer.canonical_word(), idx)) return false;
    if (idx >= sampled_vec.len) return false;
    if (sampled_vec.bit(idx)) {
      u64 pos = sampled_pos.get(sampled_vec.rank(idx));
      return k2u_w_pos(kmer, pos, out);
    }
    int64_t signed_shift = 0;
    u64 current_rank = sampled_vec.rank(idx);
    u64 extension_pos = idx - current_rank;
    u64 extension_word = ext_bases.get(extension_pos);
    CanonicalKmer km = kmer;
    if ((!canonical_vec.bit(extension_pos)) ^ (!km.is_fw_canonical())) km.swap();
    bool shift_fw = direction_vec.bit(extension_pos);
    u64 llimit = extension_size - (ext_sizes.get(extension_pos) + 1);
    const u64 mask = kmer_mask(km.k);
    for (u64 i = extension_size; i >= llimit + 1; --i) {
      u64 ssize = 2 * (i - 1);
      u64 code = (extension_word >> ssize) & 3;
      if (shift_fw) {  // append_base
        km.fw = ((km.fw >> 2) | (code << (2 * (km.k - 1)))) & mask;
        km.rc = ((km.rc << 2) | (3 - code)) & mask;
        signed_shift -= 1;
      } else {  // prepend_base
        km.fw = ((km.fw << 2) | code) & mask;
        km.rc = (km.rc >> 2) | ((3 - code) << (2 * (km.k - 1)));
        signed_shift += 1;
      }
    }
    if (!mphf.lookup(km.canonical_word(), idx)) return false;
    if (idx >= sampled_vec.len || !sampled_vec.bit(idx)) return false;
    u64 sample_pos = sampled_pos.get(sampled_vec.rank(idx));
    u64 pos = (u64)((int64_t)sample_pos + signed_shift);
    if (!us.is_valid_useq_pos(pos)) return false;
    return k2u_w_pos(kmer, pos, out);
  }
};
inline std::unique_ptr<ModIndex> sparse_index_from_pf1(const std::string& dir) {  // pf1/sparse_index.rs:32-110
  std::string info = slurp(dir + "/info.json");
  int k = (int)json_u64(info, "k");
  auto idx = std::make_unique<ModIndex>();
  auto h = std::make_unique<SampledPFHash>();
  h->us = unitig_set_from_pf1(dir, k);
  h->mphf = BooPHF::load(dir + "/mphf.bin");
  h->sampled_pos = read_compact_vector(dir + "/sample_pos.bin");
  h->canonical_vec = bitvector_from_compact(dir + "/canonical.bin");
  h->direction_vec = bitvector_from_compact(dir + "/direction.bin");
  h->ext_sizes = read_compact_vector(dir + "/extensionSize.bin");
  h->ext_bases = read_compact_vector(dir + "/extension.bin");
  h->sampled_vec = bitvector_from_compact(dir + "/presence.bin");
  h->sample_size = json_u64(info, "sample_size");
  h->extension_size = json_u64(info, "extension_size");
  idx->k2u = std::move(h);
  idx->u2pos = dense_unitig_table_from_pf1(dir);
  idx->refs = refseq_from_pf1(dir);
  return idx;
}

// ------------------------------------------------------------------------------------
// cuttlefish.rs:11-183, unitig_set.rs:119-165, spt.rs:67-140, spt_compact.rs:287-389
// ------------------------------------------------------------------------------------
struct CfToken {
  bool is_n;
  u64 n_or_id;
  u32 fw;
};
inline std::vector<std::pair<std::string, std::vector<CfToken>>> read_cf_seq(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw OracleError("cannot open " + path);
  std::vector<std::pair<std::string, std::vector<CfToken>>> out;
  std::string line;
  while (std::getline(f, line)) {
    if (line.empty()) continue;
    size_t tab = line.find('\t');
    std::string name = line.substr(0, tab);
    std::vector<CfToken> toks;
    std::stringstream ss(line.substr(tab + 1));
    std::string t;
    while (std::getline(ss, t, ' ')) {
      if (t.empty()) continue;
      CfToken tok;
      if (t[0] == 'N') {  // cuttlefish.rs:117-122
        tok.is_n = true;
        tok.n_or_id = strtoull(t.c_str() + 1, nullptr, 10);
        tok.fw = 0;
      } else {  // cuttlefish.rs:123-135
        tok.is_n = false;
        tok.fw = t.back() == '+' ? 1 : 0;
        tok.n_or_id = strtoull(t.substr(0, t.size() - 1).c_str(), nullptr, 10);
      }
      toks.push_back(tok);
    }
    out.push_back({name, toks});
  }
  return out;
}
struct CfLoaded {
  UnitigSet us;
  std::unordered_map<u64, u64> cfid2uid;
};
inline CfLoaded unitig_set_from_cf(const std::string& prefix) {  // unitig_set.rs:119-165
  std::string info = slurp(prefix + ".json");
  int k = (int)json_u64(info, "k");
  std::ifstream f(prefix + ".cf_seg");
  if (!f) throw OracleError("cannot open " + prefix + ".cf_seg");
  CfLoaded L;
  SeqVector useq;
  std::vector<u64> accum;
  u64 ps = 0, i = 0;
  std::string line;
  while (std::getline(f, line)) {
    if (line.empty()) continue;
    size_t tab = line.find('\t');
    u64 id = strtoull(line.substr(0, tab).c_str(), nullptr, 10);
    std::string seq = line.substr(tab + 1);
    while (!seq.empty() && (seq.back() == '\r' || seq.back() == '\n')) seq.pop_back();
    useq.push_chars((const u8*)seq.data(), seq.size());
    L.cfid2uid[id] = i++;
    accum.push_back(ps);
    ps += seq.size();
  }
  accum.push_back(ps);
  L.us = UnitigSet::from_accum(k, std::move(useq), accum);
  return L;
}
struct SPTParts {
  std::vector<std::string> ref_names;
  std::vector<u64> offsets;      // prefix sum of per-unitig occurrence counts
  std::vector<UnitigOcc> occs;   // inverted lists, in insertion order
  std::vector<u64> ref_lens;
  u64 max_ref_len = 0;
};
// spt.rs:67-140 and spt_compact.rs:287-389 (identical position bookkeeping)
inline SPTParts spt_from_cf(const std::string& prefix, const CfLoaded& L) {
  const UnitigSet& us = L.us;
  u64 k = (u64)us.k, U = us.n_unitigs();
  auto tilings = read_cf_seq(prefix + ".cf_seq");
  SPTParts P;
  std::vector<u64> ufreq(U, 0);
  for (auto& t : tilings) {
    P.ref_names.push_back(t.first);
    for (auto& tok : t.second)
      if (!tok.is_n) ufreq[L.cfid2uid.at(tok.n_or_id)]++;
  }
  P.offsets = prefix_sum(ufreq);
  P.occs.resize(P.offsets[U]);
  std::vector<u64> ptrs = P.offsets;
  u64 ref_id = 0;
  for (auto& t : tilings) {
    bool prev_was_unitig = false;
    u64 pos = 0;
    for (auto& tok : t.second) {
      if (tok.is_n) {
        pos += tok.n_or_id;
        if (prev_was_unitig) pos += k - 1;
        prev_was_unitig = false;
      } else {
        u64 id = L.cfid2uid.at(tok.n_or_id);
        u64 len = us.unitig_len(id);
        P.occs[ptrs[id]++] = UnitigOcc{ref_id, pos, tok.fw};
        pos += len - k + 1;
        prev_was_unitig = true;
      }
    }
    u64 len = prev_was_unitig ? pos + k - 1 : pos;
    P.ref_lens.push_back(len);
    P.max_ref_len = std::max(P.max_ref_len, len);
    ++ref_id;
  }
  return P;
}
inline std::shared_ptr<DenseUnitigTable> dense_table_from_spt(const SPTParts& P) {  // defaults.rs:26-52
  auto t = std::make_shared<DenseUnitigTable>();
  for (auto& o : P.occs) t->ctable.push_back(encode_pf1(o));
  t->contig_offsets = IntVector::packed(P.offsets);
  t->ref_names = P.ref_names;
  return t;
}
inline std::shared_ptr<PiscemUnitigTable> piscem_table_from_spt(const SPTParts& P) {  // spt_compact.rs:132-147, piscem_index.rs:24-52
  auto t = std::make_shared<PiscemUnitigTable>();
  u32 pos_bits, ref_bits, total_bits;
  required_num_bits(P.max_ref_len, P.ref_names.size(), pos_bits, ref_bits, total_bits);
  t->ref_shift = pos_bits + 1;
  t->pos_mask = pos_bits >= 64 ? ~0ULL : ((1ULL << pos_bits) - 1);
  t->ctable = IntVector(P.occs.size(), total_bits);
  for (size_t i = 0; i < P.occs.size(); ++i) t->ctable.set(i, encode_piscem(P.occs[i], t->ref_shift));
  t->contig_offsets = IntVector::packed(P.offsets);
  t->ref_names = P.ref_names;
  return t;
}
inline std::shared_ptr<RefSeqCollection> refs_from_spt(const SPTParts& P) {  // spt.rs:142-147
  auto rs = std::make_shared<RefSeqCollection>();
  rs->prefix = prefix_sum(P.ref_lens);
  return rs;
}

// util.rs:93-149 FastaReader (records = header line + concatenated sequence lines)
inline std::vector<std::pair<std::string, std::string>> read_fasta(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw OracleError("cannot open " + path);
  std::vector<std::pair<std::string, std::string>> recs;
  std::string line;
  while (std::getline(f, line)) {
    while (!line.empty() && (line.back() == '\r')) line.pop_back();
    if (!line.empty() && line[0] == '>') recs.push_back({line.substr(1), ""});
    else if (!recs.empty()) recs.back().second += line;
  }
  return recs;
}

}  // namespace mazu_oracle
