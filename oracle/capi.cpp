// TEST INFRASTRUCTURE ONLY -- C entry points of the CPU oracle (see mazu_oracle.hpp header).
// Loaded with ctypes from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs.  The product never links or calls this library.
#include <atomic>
#include <thread>

#include "mazu_oracle.hpp"

using namespace mazu_oracle;

namespace {
thread_local std::string g_err;

struct Hit {  // same 16-byte record the product writes (include/mazu_b200.h mazu_hit_t)
  u32 unitig_id, unitig_len, pos, match;
};
struct Occ {  // 12-byte decoded occurrence / projected position (mazu_occ_t)
  u32 ref_id, pos, fw;
};
inline Hit miss_hit(u32 match = 0) { return Hit{~0u, ~0u, ~0u, match}; }
inline Hit to_hit(const K2UPos& p) { return Hit{(u32)p.unitig_id, (u32)p.unitig_len, (u32)p.pos, (u32)p.o}; }

struct Handle {
  std::unique_ptr<ModIndex> idx;
  SSHash* sshash = nullptr;  // non-owning view when k2u is an SSHash
};

template <class F>
int guard(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
template <class F>
void* guard_ptr(F&& f) {
  try {
    return f();
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

Handle* make_cf(const std::string& prefix, int kind, int w, u64 skew, u64 seed) {
  CfLoaded L = unitig_set_from_cf(prefix);
  SPTParts P = spt_from_cf(prefix, L);
  auto h = new Handle();
  h->idx = std::make_unique<ModIndex>();
  h->idx->refs = refs_from_spt(P);
  if (kind == 0) {  // index/defaults.rs:17-58 PufferfishDenseIndexDefault::from_cf_prefix
    h->idx->u2pos = dense_table_from_spt(P);
    h->idx->k2u = pfhash_from_unitig_set(std::move(L.us));
  } else {  // index/piscem_index.rs:14-58 PiscemIndex::from_cf_prefix
    h->idx->u2pos = piscem_table_from_spt(P);
    auto s = SSHash::from_unitig_set(std::move(L.us), w, skew, seed);
    h->sshash = s.get();
    h->idx->k2u = std::move(s);
  }
  return h;
}
}  // namespace

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

// ---- unit-level entry points (golden-vector tests) --------------------------------------
u64 orc_simple_hash64(u64 key, u64 seed) { return simple_hash64(key, seed); }
void orc_multihash_chain(u64 key, u64 n, u64* out) {  // h0, h1, next, next, ...
  MultiHashState st;
  for (u64 i = 0; i < n; ++i) out[i] = i == 0 ? multihash_h0(st, key) : i == 1 ? multihash_h1(st, key) : multihash_next(st);
}
void* orc_boophf_load(const char* path) {
  return guard_ptr([&]() -> void* { return new BooPHF(BooPHF::load(path)); });
}
void orc_boophf_free(void* p) { delete (BooPHF*)p; }
int orc_boophf_lookup(void* p, u64 key, u64* out) { return ((BooPHF*)p)->lookup(key, *out) ? 1 : 0; }
int orc_boophf_final_lookup(void* p, u64 key, u64* out) {
  BooPHF* m = (BooPHF*)p;
  auto it = m->final_hash.find(key);
  if (it == m->final_hash.end()) return 0;
  *out = it->second + m->last_bitset_rank;
  return 1;
}
// what: 0 n_elem, 1 n_levels, 2 final_hash size, 3 level0 word0, 4 level `arg` n_bits
u64 orc_boophf_info(void* p, int what, u64 arg) {
  BooPHF* m = (BooPHF*)p;
  switch (what) {
    case 0: return m->n_elem;
    case 1: return m->levels.size();
    case 2: return m->final_hash.size();
    case 3: return m->levels[0].data[0];
    case 4: return m->levels[arg].n_bits;
  }
  return 0;
}
// boophf/mod.rs:359-382 test_levels_ranks: BooPHF block rank == plain rank + offset of previous levels
int orc_boophf_check_ranks(void* p) {
  BooPHF* m = (BooPHF*)p;
  u64 offset = 0;
  for (auto& lv : m->levels) {
    BitVector bv(lv.n_bits);
    for (u64 i = 0; i < lv.n_bits; ++i)
      if (lv.bit(i)) bv.set_bit(i);
    bv.enable_rank();
    u64 last = 0;
    for (u64 i = 0; i < lv.n_bits; ++i) {
      if (bv.rank(i) + offset != lv.rank(i)) return 0;
      last = lv.rank(i);
    }
    (void)last;
    offset += bv.count_ones();
  }
  return 1;
}

// Elias-Fano: returns 0 ok, -1 error (EFNotMonotone / EFEmpty); out[i] = get(i); meta = {l, high_bits len}
int orc_ef_roundtrip(const u64* xs, u64 n, u64 u, int use_last_as_u, u64* out, u64* meta) {
  return guard([&] {
    std::vector<u64> v(xs, xs + n);
    EFVector ef = use_last_as_u ? EFVector::from_slice(v) : EFVector::from_iter(v, u);
    for (u64 i = 0; i < n; ++i) out[i] = ef.get(i);
    if (meta) {
      meta[0] = ef.l;
      meta[1] = ef.high_bits.len;
    }
  });
}
int orc_compact_vector_read(const char* path, u64* width, u64* len, u64* out, u64 out_cap) {
  return guard([&] {
    IntVector iv = read_compact_vector(path);
    *width = iv.width;
    *len = iv.len;
    for (u64 i = 0; i < iv.len && i < out_cap; ++i) out[i] = iv.get(i);
  });
}
u64 orc_encode_pf1(u32 ref_id, u32 pos, u32 fw) { return encode_pf1(UnitigOcc{ref_id, pos, fw}); }
void orc_decode_pf1(u64 word, u32* out3) {
  UnitigOcc o = decode_pf1(word);
  out3[0] = (u32)o.ref_id; out3[1] = (u32)o.pos; out3[2] = o.fw;
}
u64 orc_encode_piscem(u32 ref_id, u32 pos, u32 fw, u64 ref_shift) { return encode_piscem(UnitigOcc{ref_id, pos, fw}, ref_shift); }
void orc_decode_piscem(u64 ref_shift, u64 pos_mask, u64 word, u32* out3) {
  UnitigOcc o = decode_piscem(ref_shift, pos_mask, word);
  out3[0] = (u32)o.ref_id; out3[1] = (u32)o.pos; out3[2] = o.fw;
}
int orc_required_num_bits(u64 longest_ref, u64 num_refs, u32* out3) {
  return guard([&] { required_num_bits(longest_ref, num_refs, out3[0], out3[1], out3[2]); });
}
u64 orc_revcomp(u64 w, int k) { return revcomp_word(w, k); }
u64 orc_revcomp_loop(u64 w, int k) { return revcomp_word_loop(w, k); }
// threads for the one-time builders (default 1); returns the previous value
int orc_set_build_threads(int n) {
  int prev = build_threads();
  build_threads() = n < 1 ? 1 : n;
  return prev;
}
u64 orc_mm_hash32(u64 x, u64 seed) { return mm_hash32(x, seed); }
void orc_canonical_minimizer(u64 fw, int k, int w, u64 seed, u64* word, u64* offset) {
  Minimizer m = canonical_minimizer(fw, k, w, seed);
  *word = m.word;
  *offset = m.offset;
}
// K1 oracle: per k-mer slot of one read: fw, rc words, minimizer word + offset, valid flag.
// Slot p corresponds to read position p, p in [0, len-k+1).
u64 orc_encode_read(const u8* seq, u64 len, int k, int w, u64 seed, u64* fw, u64* rc, u64* mm_word, u32* mm_off, u8* valid) {
  u64 n = len >= (u64)k ? len - k + 1 : 0;
  for (u64 i = 0; i < n; ++i) {
    valid[i] = 0;
    fw[i] = rc[i] = mm_word[i] = 0;
    mm_off[i] = 0;
  }
  u64 cnt = 0;
  for_each_canonical_kmer(seq, len, k, [&](u64 pos, const CanonicalKmer& km) {
    valid[pos] = 1;
    fw[pos] = km.fw;
    rc[pos] = km.rc;
    if (w > 0) {
      Minimizer m = canonical_minimizer(km.fw, k, w, seed);
      mm_word[pos] = m.word;
      mm_off[pos] = (u32)m.offset;
    }
    ++cnt;
  });
  return cnt;
}

// ---- index construction -------------------------------------------------------------------
void* orc_dense_index_load(const char* dir) {  // pf1/dense_index.rs:33-97
  return guard_ptr([&]() -> void* {
    auto h = new Handle();
    h->idx = dense_index_from_pf1(dir);
    return h;
  });
}
void* orc_sparse_index_load(const char* dir) {  // pf1/sparse_index.rs:32-110
  return guard_ptr([&]() -> void* {
    auto h = new Handle();
    h->idx = sparse_index_from_pf1(dir);
    return h;
  });
}
void* orc_index_from_cf(const char* prefix, int kind, int w, u64 skew, u64 seed) {
  return guard_ptr([&]() -> void* { return make_cf(prefix, kind, w, skew, seed); });
}
// k2u_kind: 0 = PFHash::from_unitig_set, 1 = SSHash::from_unitig_set(w, skew, seed); no U2Pos.
void* orc_index_from_packed(int k, const u64* words, u64 n_bases, const u64* accum_lens, u64 n_unitigs, int k2u_kind, int w, u64 skew, u64 seed) {
  return guard_ptr([&]() -> void* {
    SeqVector sv;
    sv.len = n_bases;
    sv.words.assign(words, words + (2 * n_bases + 63) / 64);
    sv.ensure(n_bases);
    std::vector<u64> accum(accum_lens, accum_lens + n_unitigs + 1);
    UnitigSet us = UnitigSet::from_accum(k, std::move(sv), accum);
    auto h = new Handle();
    h->idx = std::make_unique<ModIndex>();
    if (k2u_kind == 0) h->idx->k2u = pfhash_from_unitig_set(std::move(us));
    else {
      auto s = SSHash::from_unitig_set(std::move(us), w, skew, seed);
      h->sshash = s.get();
      h->idx->k2u = std::move(s);
    }
    return h;
  });
}
void* orc_index_from_seqs(const char* concat, const u64* offsets, u64 n, int k, int k2u_kind, int w, u64 skew, u64 seed) {
  return guard_ptr([&]() -> void* {
    std::vector<std::string> seqs;
    for (u64 i = 0; i < n; ++i) seqs.emplace_back(concat + offsets[i], concat + offsets[i + 1]);
    UnitigSet us = UnitigSet::from_seqs(seqs, k);
    auto h = new Handle();
    h->idx = std::make_unique<ModIndex>();
    if (k2u_kind == 0) h->idx->k2u = pfhash_from_unitig_set(std::move(us));
    else {
      auto s = SSHash::from_unitig_set(std::move(us), w, skew, seed);
      h->sshash = s.get();
      h->idx->k2u = std::move(s);
    }
    return h;
  });
}
// ModIndex::from_parts(base, SSHash::from_unitig_set(unitigs.clone(), w, skew), u2pos.clone(), refs.clone())
// as in pf1/dense_index.rs:315-328; k2u_kind as above.
void* orc_index_rebuild_k2u(void* hp, int k2u_kind, int w, u64 skew, u64 seed) {
  return guard_ptr([&]() -> void* {
    Handle* src = (Handle*)hp;
    UnitigSet us = src->idx->k2u->unitigs();
    auto h = new Handle();
    h->idx = std::make_unique<ModIndex>();
    h->idx->u2pos = src->idx->u2pos;
    h->idx->refs = src->idx->refs;
    if (k2u_kind == 0) h->idx->k2u = pfhash_from_unitig_set(std::move(us));
    else {
      auto s = SSHash::from_unitig_set(std::move(us), w, skew, seed);
      h->sshash = s.get();
      h->idx->k2u = std::move(s);
    }
    return h;
  });
}
// synthetic U2Pos attach (config 4): occurrences given decoded, offsets per unitig; kind 0 dense(pf1 words), 1 piscem
int orc_attach_u2pos(void* hp, int kind, const u64* offsets, u64 n_unitigs, const u32* ref_ids, const u32* poss, const u8* fws, u64 n_occs, u64 max_ref_len, u64 n_refs) {
  return guard([&] {
    Handle* h = (Handle*)hp;
    SPTParts P;
    P.offsets.assign(offsets, offsets + n_unitigs + 1);
    P.occs.resize(n_occs);
    for (u64 i = 0; i < n_occs; ++i) P.occs[i] = UnitigOcc{ref_ids[i], poss[i], fws[i]};
    P.ref_names.resize(n_refs);
    P.max_ref_len = max_ref_len;
    if (kind == 0) h->idx->u2pos = dense_table_from_spt(P);
    else h->idx->u2pos = piscem_table_from_spt(P);
  });
}
void orc_index_free(void* hp) { delete (Handle*)hp; }

// what: 0 k, 1 n_unitigs, 2 n_kmers, 3 total_len, 4 n_minimizers (len of prefix sum), 5 n_kmers_in_skew,
//       6 n_refs, 7 n_total_occs, 8 n_minimizer_occs, 9 piscem ref_shift, 10 piscem pos_mask, 11 has_refseq,
//       12 sparse sample_size, 13 sparse extension_size
u64 orc_index_info(void* hp, int what) {
  Handle* h = (Handle*)hp;
  const UnitigSet& us = h->idx->k2u->unitigs();
  switch (what) {
    case 0: return (u64)us.k;
    case 1: return us.n_unitigs();
    case 2: return us.n_kmers();
    case 3: return us.total_len();
    case 4: return h->sshash ? h->sshash->n_minimizers() : 0;
    case 5: return h->sshash ? h->sshash->n_kmers_in_skew_index() : 0;
    case 6: return h->idx->refs ? h->idx->refs->n_refs() : 0;
    case 7: return h->idx->u2pos ? h->idx->u2pos->n_total_occs() : 0;
    case 8: return h->sshash ? h->sshash->n_minimizer_occs : 0;
    case 9: { auto* p = dynamic_cast<PiscemUnitigTable*>(h->idx->u2pos.get()); return p ? p->ref_shift : 0; }
    case 10: { auto* p = dynamic_cast<PiscemUnitigTable*>(h->idx->u2pos.get()); return p ? p->pos_mask : 0; }
    case 11: return h->idx->refs && h->idx->refs->has_seq;
    case 12: { auto* p = dynamic_cast<SampledPFHash*>(h->idx->k2u.get()); return p ? p->sample_size : 0; }
    case 13: { auto* p = dynamic_cast<SampledPFHash*>(h->idx->k2u.get()); return p ? p->extension_size : 0; }
  }
  return 0;
}
u64 orc_unitig_len(void* hp, u64 ui) { return ((Handle*)hp)->idx->k2u->unitigs().unitig_len(ui); }
u64 orc_unitig_start(void* hp, u64 ui) { return ((Handle*)hp)->idx->k2u->unitigs().unitig_start_pos(ui); }
u64 orc_pos_to_id(void* hp, u64 pos) { return ((Handle*)hp)->idx->k2u->unitigs().pos_to_id(pos); }
u64 orc_ref_len(void* hp, u64 ri) { return ((Handle*)hp)->idx->refs->ref_len(ri); }
// copy the unitig sequence words / reference sequence words out (for building query sets in tests)
void orc_copy_useq_words(void* hp, u64* out, u64 n_words) {
  const auto& w = ((Handle*)hp)->idx->k2u->unitigs().useq.words;
  for (u64 i = 0; i < n_words; ++i) out[i] = i < w.size() ? w[i] : 0;
}
u64 orc_ref_total_len(void* hp) { return ((Handle*)hp)->idx->refs->prefix.back(); }
void orc_copy_refseq_words(void* hp, u64* out, u64 n_words) {
  const auto& w = ((Handle*)hp)->idx->refs->seq.words;
  for (u64 i = 0; i < n_words; ++i) out[i] = i < w.size() ? w[i] : 0;
}

// ---- queries --------------------------------------------------------------------------------
// K2U::k2u on a batch of forward k-mer words (query k = qk; mismatch -> error, as the reference panics)
int orc_k2u_batch(void* hp, const u64* fw_words, u64 n, int qk, void* out_hits, int n_threads) {
  return guard([&] {
    Handle* h = (Handle*)hp;
    Hit* out = (Hit*)out_hits;
    const K2U& k2u = *h->idx->k2u;
    if (qk != k2u.k()) throw OracleError("Got query k-mer size k=" + std::to_string(qk) + ", expected k=" + std::to_string(k2u.k()));
    auto work = [&](u64 lo, u64 hi) {
      for (u64 i = lo; i < hi; ++i) {
        K2UPos p;
        out[i] = k2u.k2u(CanonicalKmer::from_u64(fw_words[i], qk), p) ? to_hit(p) : miss_hit();
      }
    };
    if (n_threads <= 1) work(0, n);
    else {
      std::vector<std::thread> ts;
      for (int t = 0; t < n_threads; ++t) ts.emplace_back(work, n * t / n_threads, n * (t + 1) / n_threads);
      for (auto& t : ts) t.join();
    }
  });
}
int orc_k2u_fw_batch(void* hp, const u64* fw_words, u64 n, int qk, void* out_hits) {  // sshash.rs:563-624
  return guard([&] {
    Handle* h = (Handle*)hp;
    if (!h->sshash) throw OracleError("k2u_fw: not an SSHash");
    Hit* out = (Hit*)out_hits;
    for (u64 i = 0; i < n; ++i) {
      K2UPos p;
      out[i] = h->sshash->k2u_fw(fw_words[i], qk, p) ? to_hit(p) : miss_hit();
    }
  });
}
// total number of k-mer slots for a read batch: sum over reads of max(len-k+1, 0); fills kmer_offsets[n_reads+1]
u64 orc_kmer_offsets(const u64* read_offsets, u64 n_reads, int k, u64* kmer_offsets) {
  u64 acc = 0;
  for (u64 r = 0; r < n_reads; ++r) {
    kmer_offsets[r] = acc;
    u64 len = read_offsets[r + 1] - read_offsets[r];
    if (len >= (u64)k) acc += len - k + 1;
  }
  kmer_offsets[n_reads] = acc;
  return acc;
}
// The shape of `kphf bench` (bin/kphf/main.rs:273-339) and validate_ckmers: per read, iterate
// canonical k-mers, query random-access (streaming=0) or via StreamingK2U (streaming=1).
// reset_per_read=1 creates a fresh StreamingK2U per read (the GPU contract); 0 keeps one
// cursor across all reads of a thread's shard like the reference drivers (caching.rs:204-218).
// out_hits may be null (count only).  counts = {n_kmers (valid), n_hit, n_miss}.
int orc_query_reads(void* hp, const u8* bases, const u64* read_offsets, u64 n_reads, int streaming, int reset_per_read,
                    const u64* kmer_offsets, void* out_hits, u64* counts, int n_threads) {
  return guard([&] {
    Handle* h = (Handle*)hp;
    Hit* out = (Hit*)out_hits;
    const K2U& k2u = *h->idx->k2u;
    int k = k2u.k();
    std::atomic<u64> a_k{0}, a_h{0}, a_m{0};
    auto work = [&](u64 lo, u64 hi) {
      StreamingK2U st(&k2u);
      u64 nk = 0, nh = 0, nm = 0;
      for (u64 r = lo; r < hi; ++r) {
        const u8* seq = bases + read_offsets[r];
        u64 len = read_offsets[r + 1] - read_offsets[r];
        u64 nslots = len >= (u64)k ? len - k + 1 : 0;
        Hit* o = out ? out + kmer_offsets[r] : nullptr;
        if (o) for (u64 i = 0; i < nslots; ++i) o[i] = miss_hit(MATCH_SKIPPED);
        if (reset_per_read) st.reset();
        for_each_canonical_kmer(seq, len, k, [&](u64 pos, const CanonicalKmer& km) {
          K2UPos p;
          bool ok = streaming ? st.k2u_streaming(km, p) : k2u.k2u(km, p);
          ++nk;
          if (ok) ++nh; else ++nm;
          if (o) o[pos] = ok ? to_hit(p) : miss_hit();
        });
      }
      a_k += nk; a_h += nh; a_m += nm;
    };
    if (n_threads <= 1) work(0, n_reads);
    else {
      std::vector<std::thread> ts;
      for (int t = 0; t < n_threads; ++t) ts.emplace_back(work, n_reads * t / n_threads, n_reads * (t + 1) / n_threads);
      for (auto& t : ts) t.join();
    }
    if (counts) { counts[0] = a_k; counts[1] = a_h; counts[2] = a_m; }
  });
}

// U2Pos::encoded_unitig_occs + decode_unitig_occs for a batch of unitig ids.
// out_offsets[n+1] = prefix of list lengths; out_occs may be null (count pass).
int orc_decode_occs(void* hp, const u32* unitig_ids, u64 n, u64* out_offsets, void* out_occs) {
  return guard([&] {
    Handle* h = (Handle*)hp;
    if (!h->idx->u2pos) throw OracleError("index has no U2Pos");
    Occ* out = (Occ*)out_occs;
    u64 acc = 0;
    for (u64 i = 0; i < n; ++i) {
      out_offsets[i] = acc;
      if (unitig_ids[i] == ~0u) continue;
      u64 s, e;
      h->idx->u2pos->encoded_unitig_occs(unitig_ids[i], s, e);
      if (out)
        for (u64 j = s; j < e; ++j) {
          UnitigOcc o = h->idx->u2pos->decode_at(j);
          out[acc + (j - s)] = Occ{(u32)o.ref_id, (u32)o.pos, o.fw};
        }
      acc += e - s;
    }
    out_offsets[n] = acc;
  });
}
// GetRefPos::get_ref_pos_eager's projection (index.rs:174-216) for a batch of hits (misses -> empty list).
int orc_project_hits(void* hp, const void* hits_v, u64 n, u64* out_offsets, void* out_mrps) {
  return guard([&] {
    Handle* h = (Handle*)hp;
    if (!h->idx->u2pos) throw OracleError("index has no U2Pos");
    const Hit* hits = (const Hit*)hits_v;
    Occ* out = (Occ*)out_mrps;
    u64 acc = 0, k = (u64)h->idx->k();
    for (u64 i = 0; i < n; ++i) {
      out_offsets[i] = acc;
      if (hits[i].match != IdentityMatch && hits[i].match != TwinMatch) continue;
      K2UPos p;
      p.unitig_id = hits[i].unitig_id; p.unitig_len = hits[i].unitig_len; p.pos = hits[i].pos; p.o = (MatchType)hits[i].match;
      u64 s, e;
      h->idx->u2pos->encoded_unitig_occs(p.unitig_id, s, e);
      if (out)
        for (u64 j = s; j < e; ++j) {
          MappedRefPos m = project_onto_u_occ(k, p, h->idx->u2pos->decode_at(j));
          out[acc + (j - s)] = Occ{(u32)m.ref_id, (u32)m.pos, m.fw};
        }
      acc += e - s;
    }
    out_offsets[n] = acc;
  });
}

// ---- validation drivers (counts = {n_queries, n_identity, n_twin, n_projected, n_fail}) -----
int orc_validate_self(void* hp, u64* counts) {  // index/validate.rs:24-52
  return guard([&] {
    ValidateCounts c = ((Handle*)hp)->idx->validate_self();
    counts[0] = c.n_queries; counts[1] = c.n_identity; counts[2] = c.n_twin; counts[3] = c.n_projected; counts[4] = c.n_fail;
  });
}
int orc_k2u_validate_self(void* hp, u64* counts) {  // kphf/mod.rs:69-103
  return guard([&] {
    ValidateCounts c = k2u_validate_self(*((Handle*)hp)->idx->k2u);
    counts[0] = c.n_queries; counts[1] = c.n_identity; counts[2] = c.n_twin; counts[3] = c.n_projected; counts[4] = c.n_fail;
  });
}
int orc_validate_fasta(void* hp, const char* path, int streaming, u64* counts) {  // validate.rs:83-100, caching.rs:204-218
  return guard([&] {
    Handle* h = (Handle*)hp;
    auto recs = read_fasta(path);
    ValidateCounts tot;
    StreamingK2U st(h->idx->k2u.get());  // one cursor across ALL records (caching.rs:204-218)
    for (size_t ri = 0; ri < recs.size(); ++ri) {
      ValidateCounts c = h->idx->validate_ckmers(ri, (const u8*)recs[ri].second.data(), recs[ri].second.size(), streaming ? &st : nullptr);
      tot.n_queries += c.n_queries; tot.n_identity += c.n_identity; tot.n_twin += c.n_twin; tot.n_projected += c.n_projected; tot.n_fail += c.n_fail;
    }
    counts[0] = tot.n_queries; counts[1] = tot.n_identity; counts[2] = tot.n_twin; counts[3] = tot.n_projected; counts[4] = tot.n_fail;
  });
}
// ModIndex::iter_unitigs_on_ref (index.rs:363-424): out = 4 x u32 per tile {unitig_id, unitig_len, ref pos, fw}; returns #tiles or -1
long long orc_iter_unitigs_on_ref(void* hp, u64 ref_id, u32* out, u64 cap) {
  long long n = -1;
  guard([&] {
    auto v = iter_unitigs_on_ref(*((Handle*)hp)->idx, ref_id);
    for (u64 i = 0; i < v.size() && i < cap; ++i) {
      out[4 * i] = (u32)v[i].unitig_id; out[4 * i + 1] = (u32)v[i].unitig_len; out[4 * i + 2] = (u32)v[i].pos; out[4 * i + 3] = v[i].fw;
    }
    n = (long long)v.size();
  });
  return n;
}

// single eager query on a k-mer string (index.rs:139-142); returns #MappedRefPos or -1 (None) or -2 (error/panic)
int orc_get_ref_pos_eager_str(void* hp, const char* kmer, void* out_mrps, int cap, void* out_hit) {
  int n = -2;
  guard([&] {
    Handle* h = (Handle*)hp;
    CanonicalKmer km = CanonicalKmer::from_str(kmer);
    K2UPos hit;
    std::vector<MappedRefPos> v;
    if (!h->idx->u2pos) {
      if (km.len() != h->idx->k()) throw OracleError("Got query k-mer size mismatch");
      bool ok = h->idx->k2u->k2u(km, hit);
      n = ok ? 0 : -1;
    } else {
      bool ok = h->idx->get_ref_pos_eager(km, hit, v);
      n = ok ? (int)v.size() : -1;
    }
    if (n >= 0 && out_hit) *(Hit*)out_hit = to_hit(hit);
    Occ* out = (Occ*)out_mrps;
    for (int i = 0; i < n && i < cap; ++i) out[i] = Occ{(u32)v[i].ref_id, (u32)v[i].pos, v[i].fw};
  });
  return n;
}

}  // extern "C"
