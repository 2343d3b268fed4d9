// CPU-side consistency checks of the PRODUCT's host builders and device-layout primitives
// (mazu_b200/csrc/{host_build,formats,index_layout}.hpp).  No oracle, no CUDA: this verifies that
// the tables the library uploads are self-consistent (MPHF is a bijection, blocked Elias-Fano decodes
// to its input, every unitig k-mer is reachable through minimizer -> MPHF -> bucket -> position,
// the unitig directory locates every position) before any GPU time is spent.
// Usage: host_check <pf1_dir> <cf_prefix>
#include <cstdio>
#include <random>
#include <set>

#include "../mazu_b200/csrc/formats.hpp"

using namespace mazu;

static int g_fail = 0;
#define CHECK(c, ...)                      \
  do {                                     \
    if (!(c)) {                            \
      if (g_fail < 20) {                   \
        printf("FAIL %s:%d: ", __FILE__, __LINE__); \
        printf(__VA_ARGS__);               \
        printf("\n");                      \
      }                                    \
      ++g_fail;                            \
    }                                      \
  } while (0)

static void check_blocked_ef() {
  std::mt19937_64 g(5);
  struct Case { u64 n, maxgap; };
  for (Case c : {Case{1, 1}, Case{2, 3}, Case{33, 2}, Case{1000, 2}, Case{1000, 3}, Case{5000, 40}, Case{5000, 100000}, Case{300, 1u << 20}, Case{4097, 1}}) {
    std::vector<u64> xs(c.n);
    u64 acc = 0;
    for (u64 i = 0; i < c.n; ++i) {
      acc += g() % c.maxgap;
      if (g() % 97 == 0) acc += g() % (c.maxgap * 300 + 1);  // occasional heavy bucket -> exception blocks
      xs[i] = acc;
    }
    if (xs.back() == 0) xs.back() = 1;
    BlockedEF ef = BlockedEF::build(xs);
    for (u64 i = 0; i < c.n; ++i) CHECK(ef.get(i) == xs[i], "ef.get(%llu) n=%llu gap=%llu", (unsigned long long)i, (unsigned long long)c.n, (unsigned long long)c.maxgap);
    BlockedEFView v = ef.view();
    for (u64 i = 0; i + 1 < c.n; ++i) {
      u64 a, b;
      blocked_ef_get2(v, i, a, b);
      CHECK(a == xs[i] && b == xs[i + 1], "get2(%llu)", (unsigned long long)i);
    }
  }
  bool threw = false;
  try { BlockedEF::build({5, 8, 7, 15, 32}); } catch (const Error& e) { threw = e.code == MAZU_ERR_EF_NOT_MONOTONE; }
  CHECK(threw, "EFNotMonotone");
  threw = false;
  try { BlockedEF::build({}); } catch (const Error& e) { threw = e.code == MAZU_ERR_EF_EMPTY; }
  CHECK(threw, "EFEmpty");
  std::vector<u64> vigna{5, 8, 8, 15, 32};
  BlockedEF ef = BlockedEF::build(vigna);
  for (u64 i = 0; i < 5; ++i) CHECK(ef.get(i) == vigna[i], "vigna");
}

static void check_native_mphf() {
  std::mt19937_64 g(11);
  for (u64 n : {1ull, 2ull, 100ull, 5000ull, 300000ull}) {
    std::set<u64> s;
    while (s.size() < n) s.insert(g());
    std::vector<u64> keys(s.begin(), s.end());
    MphfHost m = MphfHost::build_native(keys, 2.0, 4);
    std::vector<u8> seen(n, 0);
    CHECK(m.meta.n_keys == n, "n_keys");
    for (u64 key : keys) {
      u64 h;
      bool ok = m.lookup(key, h);
      CHECK(ok && h < n, "member lookup");
      if (ok && h < n) {
        CHECK(!seen[h], "collision");
        seen[h] = 1;
      }
    }
    // non-members: either None or some value < n (false positive), never out of range
    for (int i = 0; i < 1000; ++i) {
      u64 h;
      if (m.lookup(g(), h)) CHECK(h < n, "non-member out of range");
    }
  }
}

static void check_unitig_dir(const UnitigSetHost& us) {
  // mirror of upload_unitigs' directory, then unitig_locate on host pointers
  const u32 shift = 6;
  u64 L = us.total_len(), U = us.n_unitigs();
  std::vector<u32> dir((L >> shift) + 2, (u32)(U - 1));
  u64 ui = 0;
  for (u64 b = 0; b < dir.size(); ++b) {
    u64 p = b << shift;
    if (p >= L) break;
    while (us.accum[ui + 1] <= p) ++ui;
    dir[b] = (u32)ui;
  }
  UnitigsView v{us.useq.data(), dir.data(), us.accum.data(), nullptr, L, U, us.k, shift};
  for (u64 p = 0; p < L; p += (L > 2000000 ? 7 : 1)) {
    u64 id, s, e;
    unitig_locate(v, p, id, s, e);
    CHECK(id == us.pos_to_id(p) && s == us.accum[id] && e == us.accum[id + 1], "locate(%llu)", (unsigned long long)p);
    if (p + us.k <= L) CHECK(useq_window(v, p) == us.window(p), "window");
  }
}

// every unitig k-mer (both orientations) must be reachable exactly like SSHash::k2u walks
static void check_sshash_tables(const K2UHost& h) {
  const UnitigSetHost& us = *h.unitigs;
  const u32 k = us.k, w = h.w;
  RankedLevels mv = h.mphf.view();
  BlockedEFView ev = h.sizes.view();
  PackedVecView pv = h.pos.view();
  u64 n_skew = 0, n_checked = 0;
  for (u64 ui = 0; ui < us.n_unitigs(); ++ui) {
    for (u64 p = us.accum[ui]; p + k <= us.accum[ui + 1]; ++p) {
      u64 fw = us.window(p), rc = revcomp(fw, k);
      for (int t = 0; t < 2; ++t) {
        u64 qf = t ? rc : fw, qr = t ? fw : rc;
        MinimizerResult mm = canonical_minimizer_naive(qf, qr, k, w, h.seed);
        u64 hh;
        bool ok = mv.family == MPHF_FAMILY_CASCADE && cascade_lookup(mv, ev, mm.word, hh);
        CHECK(ok && hh + 1 < ev.n, "minimizer not in the cascade");
        if (!ok) continue;
        CHECK(ev.wpb == 8 && blocked_ef_fp(ev, hh) == cascade_fp(fmix64(mm.word)), "fingerprint of a member minimizer");
        u64 a, b;
        blocked_ef_get2(ev, hh, a, b);
        CHECK(b > a, "empty bucket");
        if (b - a > h.skew_param) {
          CHECK(h.has_skew, "skew bucket without skew index");
          u64 hs;
          bool oks = h.skew_mphf.lookup(std::min(qf, qr), hs);
          CHECK(oks && hs < h.skew_pos.len && h.skew_pos.get(hs) == p, "skew lookup");
          ++n_skew;
          continue;
        }
        bool found = false;
        u64 off = mm.offset, rco = k - mm.offset - w;
        for (u64 pi = a; pi < b && !found; ++pi) {
          u64 mp = packed_get(pv, pi);
          if (mp >= off && mp - off == p && t == 0) found = true;   // fw k-mer sits forward
          if (mp >= rco && mp - rco == p && t == 1) found = true;   // swapped query: its rc sits forward
        }
        CHECK(found, "k-mer at %llu not reachable (t=%d)", (unsigned long long)p, t);
        ++n_checked;
      }
    }
  }
  printf("  sshash: %llu lookups via buckets, %llu via skew index, %llu minimizers, %llu occs, ef l=%u log_s=%u exc=%llu, mphf levels=%u\n",
         (unsigned long long)n_checked, (unsigned long long)n_skew, (unsigned long long)h.n_minimizers, (unsigned long long)h.n_minimizer_occs,
         h.sizes.l, h.sizes.log_s, (unsigned long long)h.sizes.n_exception_blocks, h.mphf.meta.n_levels);
}

static void check_pfhash_tables(const K2UHost& h) {
  const UnitigSetHost& us = *h.unitigs;
  const u32 k = us.k;
  RankedLevels mv = h.mphf.view();
  PackedVecView pv = h.pos.view();
  for (u64 ui = 0; ui < us.n_unitigs(); ++ui)
    for (u64 p = us.accum[ui]; p + k <= us.accum[ui + 1]; ++p) {
      u64 fw = us.window(p), rc = revcomp(fw, k), hh;
      bool ok = mphf_lookup(mv, std::min(fw, rc), hh);
      CHECK(ok && hh < pv.len && packed_get(pv, hh) == p, "pfhash pos[h(kmer)] != pos at %llu", (unsigned long long)p);
    }
}

int main(int argc, char** argv) {
  if (argc < 3) {
    printf("usage: host_check <pf1_dir> <cf_prefix>\n");
    return 2;
  }
  try {
    check_blocked_ef();
    printf("blocked EF ok (%d failures so far)\n", g_fail);
    check_native_mphf();
    printf("native MPHF ok (%d failures so far)\n", g_fail);
    {
      LoadedIndex L = load_pf1_dense(argv[1]);
      printf("pf1: k=%u unitigs=%llu kmers=%llu occs=%llu levels=%u\n", L.unitigs->k, (unsigned long long)L.unitigs->n_unitigs(),
             (unsigned long long)L.unitigs->n_kmers(), (unsigned long long)L.u2pos->n_occs, L.k2u->mphf.meta.n_levels);
      check_unitig_dir(*L.unitigs);
      check_pfhash_tables(*L.k2u);  // BooPHF re-blocked: pos.bin is addressed by the C++ hash values
      auto ss = build_sshash(L.unitigs, 15, 32, 0);
      check_sshash_tables(*ss);
      auto ss2 = build_sshash(L.unitigs, 15, MAZU_SKEW_NONE, 0);
      check_sshash_tables(*ss2);
      auto pf = build_pfhash(L.unitigs);
      check_pfhash_tables(*pf);
    }
    {
      LoadedIndex L = load_cf_prefix(argv[2], MAZU_INDEX_PISCEM, 3, 0, 0);
      printf("cf: k=%u unitigs=%llu kmers=%llu occs=%llu skew kmers=%llu\n", L.unitigs->k, (unsigned long long)L.unitigs->n_unitigs(),
             (unsigned long long)L.unitigs->n_kmers(), (unsigned long long)L.u2pos->n_occs, (unsigned long long)L.k2u->n_skew_kmers);
      check_unitig_dir(*L.unitigs);
      check_sshash_tables(*L.k2u);
      for (u32 w = 1; w <= L.unitigs->k; ++w) {
        auto ss = build_sshash(L.unitigs, w, w % 2 ? MAZU_SKEW_NONE : 1, w);
        check_sshash_tables(*ss);
      }
      LoadedIndex P = load_cf_prefix(argv[2], MAZU_INDEX_PUFFERFISH_DENSE, 0, 0, 0);
      check_pfhash_tables(*P.k2u);
    }
  } catch (const std::exception& e) {
    printf("EXCEPTION: %s\n", e.what());
    return 1;
  }
  printf("host_check: %d failures\n", g_fail);
  return g_fail ? 1 : 0;
}
