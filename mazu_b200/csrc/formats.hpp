// On-disk inputs of the path: pufferfish (C++) index directories and cuttlefish reduced-GFA
// prefixes, parsed into the host containers of host_build.hpp.  Byte layouts: SURVEY.md appendix A
// (restating src/pf1/cpp.rs:124-237, src/pf1/boophf/mod.rs:50-86,269-293, src/pf1/unitig_table.rs:28-49,
// src/pf1/mod.rs:213-236, src/cuttlefish.rs:11-183, src/spt.rs:67-140, src/spt_compact.rs:221-389).
#pragma once
#include <fstream>
#include <iterator>
#include <sstream>
#include <unordered_map>

#include "host_build.hpp"

namespace mazu {

class ByteReader {
 public:
  explicit ByteReader(const std::string& path) : path_(path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error(MAZU_ERR_IO, "cannot open " + path);
    f.seekg(0, std::ios::end);
    buf_.resize((size_t)f.tellg());
    f.seekg(0);
    f.read((char*)buf_.data(), (std::streamsize)buf_.size());
  }
  template <class T>
  T get() {
    need(sizeof(T));
    T v;
    memcpy(&v, &buf_[p_], sizeof(T));
    p_ += sizeof(T);
    return v;
  }
  void get_bytes(void* dst, size_t n) {
    need(n);
    memcpy(dst, &buf_[p_], n);
    p_ += n;
  }
  size_t remaining() const { return buf_.size() - p_; }

 private:
  void need(size_t n) {
    if (p_ + n > buf_.size()) throw Error(MAZU_ERR_IO, "unexpected end of file in " + path_);
  }
  std::string path_;
  std::vector<u8> buf_;
  size_t p_ = 0;
};

inline bool file_exists(const std::string& p) {
  std::ifstream f(p);
  return f.good();
}

// pufferfish compact vector: u64 static_flag, u64 width, u64 len, u64 capacity, words to EOF
inline PackedVec load_compact_vector(const std::string& path) {
  ByteReader r(path);
  (void)r.get<u64>();
  u64 width = r.get<u64>();
  if (width == 0 || width > 64) throw Error(MAZU_ERR_INVALID_DATA, path + ": compact vector width must be in [1,64]");
  u64 len = r.get<u64>();
  (void)r.get<u64>();
  if (r.remaining() % 8) throw Error(MAZU_ERR_INVALID_DATA, path + ": trailing bytes not divisible by 8");
  u64 nw = r.remaining() / 8;
  if (nw * 64 < len * width) throw Error(MAZU_ERR_INVALID_DATA, path + ": fewer words than len*width bits");
  PackedVec v(len, width);
  if (v.words.size() < nw + 2) v.words.resize(nw + 2, 0);
  r.get_bytes(v.words.data(), nw * 8);
  return v;
}

struct BooPHFFile {
  double gamma = 0;
  u64 last_bitset_rank = 0, n_elem = 0;
  std::vector<std::vector<u64>> level_words;
  std::vector<u64> level_n_bits;
  std::vector<u64> final_keys, final_vals;
  std::vector<const u64*> ptrs;
  mazu_boophf_desc_t desc() {
    ptrs.clear();
    for (auto& w : level_words) ptrs.push_back(w.data());
    mazu_boophf_desc_t d;
    d.n_levels = (u32)level_words.size();
    d.level_words = ptrs.data();
    d.level_n_bits = level_n_bits.data();
    d.last_bitset_rank = last_bitset_rank;
    d.n_elem = n_elem;
    d.final_keys = final_keys.data();
    d.final_vals = final_vals.data();
    d.n_final = final_keys.size();
    return d;
  }
};
inline BooPHFFile load_boophf(const std::string& path) {
  ByteReader r(path);
  BooPHFFile m;
  m.gamma = r.get<double>();
  int32_t nl = r.get<int32_t>();
  if (nl < 0 || nl > (int32_t)MPHF_MAX_LEVELS) throw Error(MAZU_ERR_INVALID_DATA, path + ": unsupported BooPHF level count");
  m.last_bitset_rank = r.get<u64>();
  m.n_elem = r.get<u64>();
  for (int32_t l = 0; l < nl; ++l) {
    u64 n_bits = r.get<u64>(), n_words = r.get<u64>();
    if (n_words * 64 < n_bits) throw Error(MAZU_ERR_INVALID_DATA, path + ": BooPHF level has fewer words than bits");
    std::vector<u64> w(n_words);
    r.get_bytes(w.data(), n_words * 8);
    u64 rs = r.get<u64>();
    std::vector<u64> ranks(rs);  // recomputed from the bits on load; parsed only to advance
    r.get_bytes(ranks.data(), rs * 8);
    m.level_words.push_back(std::move(w));
    m.level_n_bits.push_back(n_bits);
  }
  u64 fh = r.get<u64>();
  for (u64 i = 0; i < fh; ++i) {
    m.final_keys.push_back(r.get<u64>());
    m.final_vals.push_back(r.get<u64>());
  }
  return m;
}

inline u64 json_u64_field(const std::string& text, const std::string& key, const std::string& path) {
  size_t p = text.find("\"" + key + "\"");
  if (p == std::string::npos) throw Error(MAZU_ERR_INVALID_DATA, path + ": missing field '" + key + "'");
  p = text.find(':', p);
  if (p == std::string::npos) throw Error(MAZU_ERR_INVALID_DATA, path + ": malformed json");
  return strtoull(text.c_str() + p + 1, nullptr, 10);
}
inline std::string read_text(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw Error(MAZU_ERR_IO, "cannot open " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

// FastaReader (src/util.rs:93-149) over a whole file: the first line and every later line starting with '>' is a header
// (name = the line without its first character, verbatim), all other lines are appended to the current record's sequence.
// A file whose first byte is '@' is read as FASTQ with four lines per record.  Sequences are kept as written (case, N).
struct FastaHost {
  std::vector<u8> bases;
  std::vector<u64> offsets{0};
  std::vector<std::string> names;
  u64 n_records() const { return offsets.size() - 1; }
};
inline FastaHost read_fasta_file(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw Error(MAZU_ERR_IO, "cannot open " + path);
  std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  FastaHost out;
  out.bases.reserve(text.size());
  const bool fastq = !text.empty() && text[0] == '@';
  size_t p = 0, line_no = 0;
  bool first = true;
  auto next_line = [&](size_t& b, size_t& e) -> bool {  // BufRead::lines: split at '\n', drop one trailing '\r'
    if (p >= text.size()) return false;
    size_t nl = text.find('\n', p);
    b = p;
    e = nl == std::string::npos ? text.size() : nl;
    p = nl == std::string::npos ? text.size() : nl + 1;
    if (e > b && text[e - 1] == '\r') --e;
    ++line_no;
    return true;
  };
  size_t b, e;
  if (fastq) {
    while (next_line(b, e)) {
      if (e == b) continue;  // blank line between records
      if (text[b] != '@') throw Error(MAZU_ERR_INVALID_DATA, "FASTQ: record header expected at line " + std::to_string(line_no));
      out.names.emplace_back(text, b + 1, e - b - 1);
      size_t sb, se, xb, xe;
      if (!next_line(sb, se)) throw Error(MAZU_ERR_INVALID_DATA, "FASTQ: truncated record");
      out.bases.insert(out.bases.end(), text.begin() + sb, text.begin() + se);
      out.offsets.push_back(out.bases.size());
      if (!next_line(xb, xe) || xe == xb || text[xb] != '+') throw Error(MAZU_ERR_INVALID_DATA, "FASTQ: '+' line expected at line " + std::to_string(line_no));
      if (!next_line(xb, xe) || xe - xb != se - sb) throw Error(MAZU_ERR_INVALID_DATA, "FASTQ: quality line does not match the sequence at line " + std::to_string(line_no));
    }
    return out;
  }
  while (next_line(b, e)) {
    if (first || (e > b && text[b] == '>')) {  // util.rs:131-133: the first line is a header whatever it starts with
      if (!first) out.offsets.push_back(out.bases.size());
      out.names.emplace_back(text, e > b ? b + 1 : b, e > b ? e - b - 1 : 0);
      first = false;
    } else {
      out.bases.insert(out.bases.end(), text.begin() + b, text.begin() + e);
    }
  }
  if (!first) out.offsets.push_back(out.bases.size());
  return out;
}

struct LoadedIndex {
  std::shared_ptr<UnitigSetHost> unitigs;
  std::shared_ptr<K2UHost> k2u;
  std::shared_ptr<U2PosHost> u2pos;
  std::shared_ptr<RefSeqHost> refs;
};

// the parts every pufferfish index directory shares: unitigs (seq.bin + rank.bin), ctable + ctg_offsets, references
inline void load_pf1_unitigs_u2pos_refs(const std::string& dir, LoadedIndex& L, u32& k_out) {
  std::string info_path = dir + "/info.json";
  std::string info = read_text(info_path);
  u32 k = (u32)json_u64_field(info, "k", info_path);
  k_out = k;
  // unitigs: seq.bin (2-bit) + rank.bin (1 at the last base of each unitig)
  PackedVec seq = load_compact_vector(dir + "/seq.bin");
  if (seq.width != 2) throw Error(MAZU_ERR_INVALID_DATA, "seq.bin: width != 2");
  PackedVec rank = load_compact_vector(dir + "/rank.bin");
  if (rank.width != 1) throw Error(MAZU_ERR_INVALID_DATA, "rank.bin: width != 1");
  auto us = std::make_shared<UnitigSetHost>();
  us->k = k;
  us->n_bases = seq.len;
  u64 nw = (2 * seq.len + 63) / 64;
  us->useq.assign(nw + 2, 0);
  memcpy(us->useq.data(), seq.words.data(), nw * 8);
  if ((2 * seq.len) & 63) us->useq[nw - 1] &= (1ULL << ((2 * seq.len) & 63)) - 1;  // drop capacity padding garbage
  us->accum.push_back(0);
  for (u64 wi = 0; wi * 64 < rank.len; ++wi) {  // select every 1 (dense_index.rs:55-66)
    u64 x = rank.words[wi];
    while (x) {
      u64 p = wi * 64 + (u64)__builtin_ctzll(x);
      if (p < rank.len) us->accum.push_back(p + 1);
      x &= x - 1;
    }
  }
  us->validate();
  L.unitigs = us;
  // ctable.bin: Vec<String> ref_names, Vec<u32> ref_exts, Vec<u64> ctable, EOF (unitig_table.rs:28-49)
  auto up = std::make_shared<U2PosHost>();
  {
    ByteReader r(dir + "/ctable.bin");
    u64 n = r.get<u64>();
    for (u64 i = 0; i < n; ++i) {
      u64 nc = r.get<u64>();
      std::string s(nc, '\0');
      r.get_bytes(&s[0], nc);
      up->ref_names.push_back(s);
    }
    u64 ne = r.get<u64>();
    std::vector<u32> exts(ne);
    r.get_bytes(exts.data(), ne * 4);
    u64 nc = r.get<u64>();
    up->ctable_words.resize(nc);
    r.get_bytes(up->ctable_words.data(), nc * 8);
    if (r.remaining() != 0) throw Error(MAZU_ERR_INVALID_DATA, "ctable.bin: trailing bytes");
    up->n_occs = nc;
  }
  up->kind = MAZU_U2POS_DENSE;
  up->ctable_width = 64;
  up->contig_offsets = load_compact_vector(dir + "/ctg_offsets.bin");
  if (up->contig_offsets.len != us->n_unitigs() + 1) throw Error(MAZU_ERR_INVALID_DATA, "ctg_offsets.bin: len != n_unitigs + 1");
  L.u2pos = up;
  // references (pf1/mod.rs:213-236)
  auto rs = std::make_shared<RefSeqHost>();
  if (file_exists(dir + "/refseq.bin")) {
    PackedVec rseq = load_compact_vector(dir + "/refseq.bin");
    if (rseq.width != 2) throw Error(MAZU_ERR_INVALID_DATA, "refseq.bin: width != 2");
    rs->has_seq = true;
    rs->seq_words = rseq.words;
  }
  {
    ByteReader r(dir + "/refAccumLengths.bin");
    u64 n = r.get<u64>();
    rs->prefix.push_back(0);
    for (u64 i = 0; i < n; ++i) rs->prefix.push_back(r.get<u64>());
  }
  L.refs = rs;
}


// DenseIndex::deserialize_from_cpp (src/pf1/dense_index.rs:33-97)
inline LoadedIndex load_pf1_dense(const std::string& dir) {
  LoadedIndex L;
  u32 k = 0;
  load_pf1_unitigs_u2pos_refs(dir, L, k);
  BooPHFFile mphf = load_boophf(dir + "/mphf.bin");
  PackedVec pos = load_compact_vector(dir + "/pos.bin");
  mazu_packed_vec_desc_t pd{pos.words.data(), pos.width, pos.len};
  L.k2u = pfhash_from_parts(L.unitigs, mphf.desc(), pd);
  return L;
}

// SparseIndex::deserialize_from_cpp (src/pf1/sparse_index.rs:32-110): the dense loader's unitigs / U2Pos / references
// plus the sampled position table, presence / canonical / direction bits and the extension words
inline LoadedIndex load_pf1_sparse(const std::string& dir) {
  LoadedIndex L;
  u32 k = 0;
  load_pf1_unitigs_u2pos_refs(dir, L, k);
  std::string info_path = dir + "/info.json";
  std::string info = read_text(info_path);
  auto H = std::make_shared<K2UHost>();
  H->kind = MAZU_K2U_SAMPLED_PFHASH;
  H->unitigs = L.unitigs;
  BooPHFFile mphf = load_boophf(dir + "/mphf.bin");
  H->mphf = MphfHost::from_boophf(mphf.desc());
  H->pos = load_compact_vector(dir + "/sample_pos.bin");
  H->ext_sizes = load_compact_vector(dir + "/extensionSize.bin");
  H->ext_bases = load_compact_vector(dir + "/extension.bin");
  H->sample_size = json_u64_field(info, "sample_size", info_path);
  H->extension_size = json_u64_field(info, "extension_size", info_path);
  if (H->extension_size == 0 || H->extension_size > 32) throw Error(MAZU_ERR_INVALID_DATA, info_path + ": extension_size out of range");
  auto bits = [&](const char* name, u64& len) {
    PackedVec v = load_compact_vector(dir + "/" + name);
    if (v.width != 1) throw Error(MAZU_ERR_INVALID_DATA, std::string(name) + ": width != 1");
    len = v.len;
    std::vector<u64> w((v.len + 63) / 64 + 2, 0);
    for (size_t i = 0; i < (v.len + 63) / 64; ++i) w[i] = v.words[i];
    if (v.len & 63) w[v.len >> 6] &= (1ULL << (v.len & 63)) - 1;
    return w;
  };
  u64 n_presence = 0, n_c = 0, n_d = 0;
  std::vector<u64> presence = bits("presence.bin", n_presence);
  H->canonical_bits = bits("canonical.bin", n_c);
  H->direction_bits = bits("direction.bin", n_d);
  if (n_presence != L.unitigs->n_kmers()) throw Error(MAZU_ERR_INVALID_DATA, "presence.bin: len != n_kmers");
  H->sampled.meta.family = MPHF_FAMILY_BOOPHF;
  H->sampled.meta.n_keys = n_presence;
  H->sampled.append_level_from_bits(presence, n_presence, n_presence);
  H->sampled.fb_keys.push_back(0);
  H->sampled.fb_vals.push_back(0);
  if (H->sampled.total_level_ones() != H->pos.len) throw Error(MAZU_ERR_INVALID_DATA, "presence.bin: #ones != sample_pos.len");
  if (H->ext_sizes.len != n_presence - H->pos.len || H->ext_bases.len != H->ext_sizes.len || n_c < H->ext_sizes.len || n_d < H->ext_sizes.len)
    throw Error(MAZU_ERR_INVALID_DATA, "sparse index: extension tables do not match the number of unsampled k-mers");
  L.k2u = H;
  return L;
}

// cuttlefish reduced GFA -> UnitigSet + SPT / SPTCompact (unitig_set.rs:119-165, spt.rs:67-140,
// spt_compact.rs:287-389) -> ModIndex as in index/defaults.rs:17-58 / index/piscem_index.rs:14-58
inline LoadedIndex load_cf_prefix(const std::string& prefix, int index_kind, u32 w, u64 skew_param, u64 seed) {
  LoadedIndex L;
  std::string jpath = prefix + ".json";
  std::string info = read_text(jpath);
  u32 k = (u32)json_u64_field(info, "k", jpath);
  auto us = std::make_shared<UnitigSetHost>();
  us->k = k;
  std::unordered_map<u64, u64> cfid2uid;
  {
    std::ifstream f(prefix + ".cf_seg");
    if (!f) throw Error(MAZU_ERR_IO, "cannot open " + prefix + ".cf_seg");
    std::string line;
    u64 i = 0;
    while (std::getline(f, line)) {
      while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
      if (line.empty()) continue;
      size_t tab = line.find('\t');
      if (tab == std::string::npos) throw Error(MAZU_ERR_INVALID_DATA, "cannot split .cf_seg line");
      cfid2uid[strtoull(line.c_str(), nullptr, 10)] = i++;
      us->push_seq(line.c_str() + tab + 1, line.size() - tab - 1);
    }
  }
  us->validate();
  L.unitigs = us;
  // tilings
  struct Tok {
    bool is_n;
    u64 v;
    u32 fw;
  };
  std::vector<std::vector<Tok>> tilings;
  auto up = std::make_shared<U2PosHost>();
  {
    std::ifstream f(prefix + ".cf_seq");
    if (!f) throw Error(MAZU_ERR_IO, "cannot open " + prefix + ".cf_seq");
    std::string line;
    while (std::getline(f, line)) {
      while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
      if (line.empty()) continue;
      size_t tab = line.find('\t');
      if (tab == std::string::npos) throw Error(MAZU_ERR_INVALID_DATA, "cannot split .cf_seq line");
      up->ref_names.push_back(line.substr(0, tab));
      std::vector<Tok> toks;
      size_t p = tab + 1;
      while (p < line.size()) {
        size_t e = line.find(' ', p);
        if (e == std::string::npos) e = line.size();
        if (e > p) {
          std::string t = line.substr(p, e - p);
          if (t[0] == 'N') toks.push_back(Tok{true, strtoull(t.c_str() + 1, nullptr, 10), 0});
          else {
            char o = t.back();
            if (o != '+' && o != '-') throw Error(MAZU_ERR_INVALID_DATA, "CfSeqTokenParseError");
            u64 id = strtoull(t.c_str(), nullptr, 10);
            auto it = cfid2uid.find(id);
            if (it == cfid2uid.end()) throw Error(MAZU_ERR_INVALID_DATA, "tiling references an unknown unitig id");
            toks.push_back(Tok{false, it->second, o == '+' ? 1u : 0u});
          }
        }
        p = e + 1;
      }
      tilings.push_back(std::move(toks));
    }
  }
  const u64 U = us->n_unitigs();
  std::vector<u64> offsets(U + 1, 0);
  for (auto& t : tilings)
    for (auto& tok : t)
      if (!tok.is_n) offsets[tok.v + 1]++;
  for (u64 i = 0; i < U; ++i) offsets[i + 1] += offsets[i];
  std::vector<mazu_occ_t> occs(offsets[U]);
  std::vector<u64> ptrs(offsets.begin(), offsets.end() - 1);
  std::vector<u64> ref_lens;
  u64 max_ref_len = 0;
  for (u64 ref_id = 0; ref_id < tilings.size(); ++ref_id) {
    bool prev_was_unitig = false;
    u64 pos = 0;
    for (auto& tok : tilings[ref_id]) {
      if (tok.is_n) {
        pos += tok.v;
        if (prev_was_unitig) pos += k - 1;
        prev_was_unitig = false;
      } else {
        occs[ptrs[tok.v]++] = mazu_occ_t{(u32)ref_id, (u32)pos, tok.fw};
        pos += us->unitig_len(tok.v) - k + 1;
        prev_was_unitig = true;
      }
    }
    u64 len = prev_was_unitig ? pos + k - 1 : pos;
    ref_lens.push_back(len);
    max_ref_len = std::max(max_ref_len, len);
  }
  up->n_occs = occs.size();
  up->contig_offsets = PackedVec::packed(offsets);
  if (index_kind == MAZU_INDEX_PUFFERFISH_DENSE) {
    up->kind = MAZU_U2POS_DENSE;
    up->ctable_width = 64;
    up->ctable_words.resize(occs.size() + 1, 0);
    for (u64 i = 0; i < occs.size(); ++i)  // UnitigOcc::encode_pf1 (index.rs:320-332)
      up->ctable_words[i] = (((u64)occs[i].pos | (occs[i].fw ? 0x80000000ULL : 0ULL)) << 32) | occs[i].ref_id;
    L.k2u = build_pfhash(us);
  } else if (index_kind == MAZU_INDEX_PISCEM) {
    if (max_ref_len == 0 || tilings.empty()) throw Error(MAZU_ERR_OTHER, "Could not construct tile occurrence table.");
    u32 pos_bits = (u32)msb(max_ref_len) + 1, ref_bits = (u32)msb(tilings.size()) + 1;  // spt_compact.rs:221-242
    u32 total = 1 + pos_bits + ref_bits;
    if (total > 64) throw Error(MAZU_ERR_OTHER, "Could not compute number of bits required for each occ entry");
    up->kind = MAZU_U2POS_PISCEM;
    up->ctable_width = total;
    up->ref_shift = pos_bits + 1;
    up->pos_mask = (1ULL << pos_bits) - 1;
    PackedVec ct(occs.size(), total);
    for (u64 i = 0; i < occs.size(); ++i)  // UnitigOcc::encode_piscem (spt_compact.rs:90-97)
      ct.set(i, ((u64)occs[i].ref_id << up->ref_shift) | ((u64)occs[i].pos << 1) | (occs[i].fw ? 1 : 0));
    up->ctable_words = ct.words;
    L.k2u = build_sshash(us, w, skew_param, seed);
  } else {
    throw Error(MAZU_ERR_INVALID_ARG, "unknown index kind");
  }
  L.u2pos = up;
  auto rs = std::make_shared<RefSeqHost>();  // SPT::get_ref_seq_collection (spt.rs:142-147): lengths only
  rs->prefix.push_back(0);
  for (u64 l : ref_lens) rs->prefix.push_back(rs->prefix.back() + l);
  L.refs = rs;
  return L;
}

}  // namespace mazu
