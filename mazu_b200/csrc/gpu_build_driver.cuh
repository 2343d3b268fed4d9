// Host-side orchestration of the device builders (kernels: gpu_build.cuh): SSHashBuilder::from_unitig_set and
// PFHash::from_unitig_set with every stage in HBM, bit-identical to host_build.hpp (src/kphf/sshash.rs:86-329,
// src/kphf/pfhash.rs:40-73).  Included by capi.cu after device_index.cuh.
#pragma once
#include "gpu_build.cuh"
namespace {

template <class T>
T d2h_value(const T* p) {
  T v;
  MZ_CUDA(cudaMemcpy(&v, p, sizeof(T), cudaMemcpyDeviceToHost));
  return v;
}
int grid_1d(u64 n, int sm_count, int block = 256) { return (int)std::max<u64>(1, std::min<u64>((n + block - 1) / block, (u64)sm_count * 16)); }

struct CubTemp {  // grow-only scratch for cub primitives
  void* p = nullptr;
  size_t bytes = 0;
  void need(size_t n) {
    if (n > bytes) {
      if (p) cudaFree(p);
      MZ_CUDA(cudaMalloc(&p, n));
      bytes = n;
    }
  }
  ~CubTemp() {
    if (p) cudaFree(p);
  }
};
void exclusive_scan_u64(CubTemp& t, const u64* in, u64* out, u64 n_items) {
  size_t b = 0;
  MZ_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, b, in, out, (size_t)n_items));
  t.need(b ? b : 1);
  MZ_CUDA(cub::DeviceScan::ExclusiveSum(t.p, b, in, out, (size_t)n_items));
}
void inclusive_scan_u64(CubTemp& t, const u64* in, u64* out, u64 n_items) {
  size_t b = 0;
  MZ_CUDA(cub::DeviceScan::InclusiveSum(nullptr, b, in, out, (size_t)n_items));
  t.need(b ? b : 1);
  MZ_CUDA(cub::DeviceScan::InclusiveSum(t.p, b, in, out, (size_t)n_items));
}
void sort_pairs_u64(CubTemp& t, const u64* kin, u64* kout, const u64* vin, u64* vout, u64 n, int end_bit) {
  size_t b = 0;
  MZ_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b, kin, kout, vin, vout, (size_t)n, 0, end_bit));
  t.need(b ? b : 1);
  MZ_CUDA(cub::DeviceRadixSort::SortPairs(t.p, b, kin, kout, vin, vout, (size_t)n, 0, end_bit));
}
u64 reduce_max_u64(CubTemp& t, const u64* in, u64 n, int dev) {
  if (n == 0) return 0;
  DevBuf out(8, dev);
  size_t b = 0;
  MZ_CUDA(cub::DeviceReduce::Max(nullptr, b, in, (u64*)out.p, (size_t)n));
  t.need(b ? b : 1);
  MZ_CUDA(cub::DeviceReduce::Max(t.p, b, in, (u64*)out.p, (size_t)n));
  return d2h_value((const u64*)out.p);
}

struct DevMphf {
  RankedLevels view{};
  std::vector<DevBufP> bufs;
  size_t bytes = 0;
  u64 block_bytes = 0, n_fb = 0;
};
// MphfHost::build_native on the device: same level sizes, same slots, hence the same bits
DevMphf build_mphf_gpu(CubTemp& tmp, const u64* d_keys, u64 n, double gamma, int dev, int sm) {
  DevMphf m;
  m.view.family = MPHF_FAMILY_NATIVE;
  auto cur = std::make_shared<DevBuf>(std::max<u64>(n, 1) * 8, dev), nxt = std::make_shared<DevBuf>(std::max<u64>(n, 1) * 8, dev);
  if (n) MZ_CUDA(cudaMemcpy(cur->p, d_keys, n * 8, cudaMemcpyDeviceToDevice));
  DevBuf counter(8, dev);
  std::vector<DevBufP> level_blocks;
  std::vector<u64> level_nb;
  u64 n_cur = n, total_ones = 0, total_nb = 0;
  for (u32 lvl = 0; lvl < MPHF_MAX_LEVELS && n_cur > 0; ++lvl) {
    const u64 nb = (u64)((native_level_gamma(gamma, lvl) * (double)n_cur) / MPHF_BLOCK_BITS) + 1;
    if (nb >> 32) throw Error(MAZU_ERR_INVALID_ARG, "MPHF level too large");
    auto seen = std::make_shared<DevBuf>(nb * 32, dev);
    DevBuf coll(nb * 32, dev), ones((nb + 1) * 4, dev), prefix((nb + 1) * 4, dev);
    MZ_CUDA(cudaMemset(seen->p, 0, nb * 32));
    MZ_CUDA(cudaMemset(coll.p, 0, nb * 32));
    MZ_CUDA(cudaMemset(ones.p, 0, (nb + 1) * 4));
    mphf_mark_kernel<<<grid_1d(n_cur, sm), 256>>>((const u64*)cur->p, n_cur, lvl, nb, (u32*)seen->p, (u32*)coll.p);
    mphf_finalize_kernel<<<grid_1d(nb, sm), 256>>>((u32*)seen->p, (const u32*)coll.p, nb, (u32*)ones.p);
    {
      size_t b = 0;
      MZ_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, b, (const u32*)ones.p, (u32*)prefix.p, (size_t)(nb + 1)));
      tmp.need(b ? b : 1);
      MZ_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, b, (const u32*)ones.p, (u32*)prefix.p, (size_t)(nb + 1)));
    }
    mphf_set_ranks_kernel<<<grid_1d(nb, sm), 256>>>((u32*)seen->p, (const u32*)prefix.p, nb);
    const u64 lvl_ones = d2h_value((const u32*)prefix.p + nb);
    MZ_CUDA(cudaMemset(counter.p, 0, 8));
    mphf_filter_kernel<<<grid_1d(n_cur, sm), 256>>>((const u64*)cur->p, n_cur, lvl, nb, (const u32*)coll.p, (u64*)nxt->p,
                                                     (unsigned long long*)counter.p);
    MZ_CUDA(cudaGetLastError());
    const u64 n_next = d2h_value((const u64*)counter.p);
    m.view.size[lvl] = nb;
    m.view.block_off[lvl] = total_nb;
    m.view.rank_base[lvl] = total_ones;
    m.view.n_levels = lvl + 1;
    total_nb += nb;
    total_ones += lvl_ones;
    level_blocks.push_back(seen);
    level_nb.push_back(nb);
    std::swap(cur, nxt);
    n_cur = n_next;
  }
  // leftovers -> sorted fallback
  std::vector<u64> left(n_cur);
  if (n_cur) MZ_CUDA(cudaMemcpy(left.data(), cur->p, n_cur * 8, cudaMemcpyDeviceToHost));
  std::sort(left.begin(), left.end());
  left.erase(std::unique(left.begin(), left.end()), left.end());
  std::vector<u64> fbv(left.size());
  for (u64 i = 0; i < left.size(); ++i) fbv[i] = total_ones + i;
  m.n_fb = left.size();
  m.view.n_keys = total_ones + left.size();
  m.view.n_fb = (u32)left.size();
  if (left.empty()) {
    left.push_back(0);
    fbv.push_back(0);
  }
  auto fk = upload(left, dev), fv = upload(fbv, dev);
  auto blocks = std::make_shared<DevBuf>((total_nb + 1) * 32, dev);
  MZ_CUDA(cudaMemset(blocks->p, 0, (total_nb + 1) * 32));
  u64 off = 0;
  for (size_t l = 0; l < level_blocks.size(); ++l) {
    MZ_CUDA(cudaMemcpy((char*)blocks->p + off * 32, level_blocks[l]->p, level_nb[l] * 32, cudaMemcpyDeviceToDevice));
    off += level_nb[l];
  }
  m.view.blocks = (const u32*)blocks->p;
  m.view.fb_keys = (const u64*)fk->p;
  m.view.fb_vals = (const u64*)fv->p;
  m.bufs = {blocks, fk, fv};
  m.bytes = blocks->bytes + fk->bytes + fv->bytes;
  m.block_bytes = total_nb * 32;
  return m;
}

struct DevCascade {
  RankedLevels view{};
  std::vector<DevBufP> bufs;  // fallback keys / values
  size_t bytes = 0;
  DevBufP states;  // one byte per slot (R of them)
  DevBufP slots;   // slot of every input key
  u64 R = 0, n_fb = 0;
};
// MphfHost::build_cascade on the device: same level sizes, same slots, hence the same states
DevCascade build_cascade_gpu(const u64* d_keys, u64 n, int dev, int sm) {
  DevCascade c;
  c.view.family = MPHF_FAMILY_CASCADE;
  c.slots = std::make_shared<DevBuf>(std::max<u64>(n, 1) * 8, dev);
  MZ_CUDA(cudaMemset(c.slots->p, 0xFF, std::max<u64>(n, 1) * 8));
  auto cur_k = std::make_shared<DevBuf>(std::max<u64>(n, 1) * 8, dev), nxt_k = std::make_shared<DevBuf>(std::max<u64>(n, 1) * 8, dev);
  auto cur_i = std::make_shared<DevBuf>(std::max<u64>(n, 1) * 8, dev), nxt_i = std::make_shared<DevBuf>(std::max<u64>(n, 1) * 8, dev);
  if (n) MZ_CUDA(cudaMemcpy(cur_k->p, d_keys, n * 8, cudaMemcpyDeviceToDevice));
  DevBuf counter(8, dev);
  std::vector<DevBufP> level_states;
  std::vector<u64> level_size;
  u64 n_cur = n, off = 0;
  for (u32 lvl = 0; lvl < MPHF_MAX_LEVELS && n_cur > 0; ++lvl) {
    const u64 size = cascade_level_size(n_cur, lvl), nw = (size + 31) / 32;
    DevBuf seen(nw * 4, dev), coll(nw * 4, dev);
    auto st = std::make_shared<DevBuf>(size, dev);
    MZ_CUDA(cudaMemset(seen.p, 0, nw * 4));
    MZ_CUDA(cudaMemset(coll.p, 0, nw * 4));
    MZ_CUDA(cudaMemset(st->p, 0, size));
    MZ_CUDA(cudaMemset(counter.p, 0, 8));
    cascade_mark_kernel<<<grid_1d(n_cur, sm), 256>>>((const u64*)cur_k->p, n_cur, lvl, size, (u32*)seen.p, (u32*)coll.p);
    cascade_place_kernel<<<grid_1d(n_cur, sm), 256>>>((const u64*)cur_k->p, lvl == 0 ? nullptr : (const u64*)cur_i->p, n_cur, lvl, size, off,
                                                       (const u32*)coll.p, (u8*)st->p, (u64*)c.slots->p, (u64*)nxt_k->p, (u64*)nxt_i->p,
                                                       (unsigned long long*)counter.p);
    MZ_CUDA(cudaGetLastError());
    const u64 n_next = d2h_value((const u64*)counter.p);
    c.view.size[lvl] = size;
    c.view.block_off[lvl] = off;
    c.view.rank_base[lvl] = 0;
    c.view.n_levels = lvl + 1;
    level_states.push_back(st);
    level_size.push_back(size);
    off += size;
    std::swap(cur_k, nxt_k);
    std::swap(cur_i, nxt_i);
    n_cur = n_next;
  }
  // leftovers -> sorted fallback, slots behind the last level
  std::vector<u64> lk(n_cur), li(n_cur);
  if (n_cur) {
    MZ_CUDA(cudaMemcpy(lk.data(), cur_k->p, n_cur * 8, cudaMemcpyDeviceToHost));
    MZ_CUDA(cudaMemcpy(li.data(), cur_i->p, n_cur * 8, cudaMemcpyDeviceToHost));
  }
  std::vector<std::pair<u64, u64>> left(n_cur);
  for (u64 i = 0; i < n_cur; ++i) left[i] = {lk[i], li[i]};
  std::sort(left.begin(), left.end());
  std::vector<u64> fbk, fbv;
  std::vector<u8> fb_states;
  for (size_t i = 0; i < left.size(); ++i) {
    if (i == 0 || left[i].first != left[i - 1].first) {
      fbk.push_back(left[i].first);
      fbv.push_back(off + fbk.size() - 1);
      fb_states.push_back((u8)cascade_fp(fmix64(left[i].first)));
    }
    const u64 slot = off + fbk.size() - 1;
    MZ_CUDA(cudaMemcpy((u64*)c.slots->p + left[i].second, &slot, 8, cudaMemcpyHostToDevice));
  }
  c.n_fb = fbk.size();
  c.R = off + c.n_fb;
  c.view.n_keys = c.R;
  c.view.n_fb = (u32)c.n_fb;
  if (fbk.empty()) {
    fbk.push_back(0);
    fbv.push_back(0);
  }
  auto fk = upload(fbk, dev), fv = upload(fbv, dev);
  c.view.fb_keys = (const u64*)fk->p;
  c.view.fb_vals = (const u64*)fv->p;
  c.view.blocks = nullptr;
  c.bufs = {fk, fv};
  c.bytes = fk->bytes + fv->bytes;
  c.states = std::make_shared<DevBuf>(std::max<u64>(c.R, 1), dev);
  u64 o = 0;
  for (size_t l = 0; l < level_states.size(); ++l) {
    MZ_CUDA(cudaMemcpy((char*)c.states->p + o, level_states[l]->p, level_size[l], cudaMemcpyDeviceToDevice));
    o += level_size[l];
  }
  if (!fb_states.empty()) MZ_CUDA(cudaMemcpy((char*)c.states->p + o, fb_states.data(), fb_states.size(), cudaMemcpyHostToDevice));
  return c;
}

// bit-pack a device array of u64 values; returns the buffer and the width chosen like PackedVec::packed
DevBufP pack_gpu(CubTemp& tmp, const u64* d_vals, u64 n, int dev, int sm, u32& width_out, u64& logical_bytes, u32 fixed_width = 0) {
  u32 width = fixed_width;
  if (!width) {
    u64 mx = reduce_max_u64(tmp, d_vals, n, dev);
    width = mx == 0 ? 1 : (u32)msb(mx) + 1;
  }
  width_out = width;
  u64 nw = (n * width + 63) / 64;
  auto b = std::make_shared<DevBuf>((nw + 2) * 8, dev);
  MZ_CUDA(cudaMemset(b->p, 0, (nw + 2) * 8));
  if (nw) pack_kernel<<<grid_1d(nw, sm), 256>>>(d_vals, n, width, (u64*)b->p, nw);
  MZ_CUDA(cudaGetLastError());
  logical_bytes = nw * 8;
  return b;
}

// distinct keys of a sorted key array: returns M and fills set / ranges (M+1 entries) / optional first values / gid
u64 group_sorted_gpu(CubTemp& tmp, const u64* d_keys, const u64* d_vals, u64 n, int dev, int sm, DevBufP& set, DevBufP& ranges, DevBufP* first_vals,
                     DevBufP& gid) {
  DevBuf flags(std::max<u64>(n, 1) * 8, dev);
  gid = std::make_shared<DevBuf>(std::max<u64>(n, 1) * 8, dev);
  mark_heads_kernel<<<grid_1d(n, sm), 256>>>(d_keys, n, (u64*)flags.p);
  inclusive_scan_u64(tmp, (const u64*)flags.p, (u64*)gid->p, n);
  const u64 M = d2h_value((const u64*)gid->p + (n - 1));
  set = std::make_shared<DevBuf>(M * 8, dev);
  ranges = std::make_shared<DevBuf>((M + 1) * 8, dev);
  if (first_vals) *first_vals = std::make_shared<DevBuf>(M * 8, dev);
  scatter_groups_kernel<<<grid_1d(n, sm), 256>>>(d_keys, d_vals, (const u64*)flags.p, (const u64*)gid->p, n, (u64*)set->p,
                                                 first_vals ? (u64*)(*first_vals)->p : nullptr, (u64*)ranges->p);
  MZ_CUDA(cudaGetLastError());
  MZ_CUDA(cudaMemcpy((u64*)ranges->p + M, &n, 8, cudaMemcpyHostToDevice));
  return M;
}

// PFHash::from_unitig_set on the device (twin of build_pfhash in host_build.hpp)
void build_pfhash_gpu(mazu_index& ix, double gamma = 2.0) {
  const UnitigSetHost& us = *ix.unitigs;
  const u32 k = us.k;
  const int dev = ix.device, sm = ix.sm_count;
  const UnitigsView uv = ix.d_unitigs->view;
  const u64 N = us.n_kmers();
  for (u64 ui = 0; ui < us.n_unitigs(); ++ui)
    if (us.unitig_len(ui) < k) throw Error(MAZU_ERR_INVALID_DATA, "a unitig is shorter than k");
  CubTemp tmp;
  auto d = std::make_shared<K2UDev>();
  DevBuf keys(std::max<u64>(N, 1) * 8, dev), positions(std::max<u64>(N, 1) * 8, dev), bad(8, dev);
  pfhash_keys_kernel<<<grid_1d(us.total_len(), sm), 256>>>(uv, (u64*)keys.p, (u64*)positions.p);
  MZ_CUDA(cudaGetLastError());
  DevMphf mphf = build_mphf_gpu(tmp, (const u64*)keys.p, N, gamma, dev, sm);
  for (auto& b : mphf.bufs) {
    d->bufs.push_back(b);
    d->bytes += b->bytes;
  }
  const u64 n_slots = mphf.view.n_keys;
  DevBuf vals(std::max<u64>(n_slots, 1) * 8, dev);
  MZ_CUDA(cudaMemset(vals.p, 0, std::max<u64>(n_slots, 1) * 8));
  MZ_CUDA(cudaMemset(bad.p, 0, 8));
  pfhash_scatter_kernel<<<grid_1d(N, sm), 256>>>(mphf.view, (const u64*)keys.p, (const u64*)positions.p, N, n_slots, (u64*)vals.p,
                                                 (unsigned long long*)bad.p);
  MZ_CUDA(cudaGetLastError());
  if (d2h_value((const u64*)bad.p) != 0) throw Error(MAZU_ERR_OTHER, "internal: GPU-built MPHF misses one of its keys");
  u32 width = 1;
  u64 bytes = 0;
  DevBufP pos = pack_gpu(tmp, (const u64*)vals.p, n_slots, dev, sm, width, bytes, (u32)std::max<u64>(1, msb(std::max<u64>(us.total_len(), 1)) + 1));
  d->bufs.push_back(pos);
  d->bytes += pos->bytes;
  auto H = std::make_shared<K2UHost>();
  H->kind = MAZU_K2U_PFHASH;
  H->unitigs = ix.unitigs;
  IndexView& v = ix.view;
  v.k2u_kind = MAZU_K2U_PFHASH;
  v.mphf = mphf.view;
  v.pos = PackedVecView{(const u64*)pos->p, n_slots, width, 0};
  v.w = 0;
  v.has_skew = 0;
  v.skew_param = MAZU_SKEW_NONE;
  ix.tables[0] = {v.mphf.blocks, mphf.block_bytes};
  ix.tables[3] = {v.pos.words, bytes};
  ix.tables[6] = {v.mphf.fb_keys, mphf.n_fb * 8};
  MZ_CUDA(cudaDeviceSynchronize());
  ix.k2u = H;
  ix.d_k2u = d;
}

// SSHashBuilder::from_unitig_set + finish on the device; fills ix.view / ix.k2u metadata / ix.d_k2u
void build_sshash_gpu(mazu_index& ix, u32 w, u64 skew_param, u64 seed, double gamma = 2.0) {
  const UnitigSetHost& us = *ix.unitigs;
  const u32 k = us.k;
  if (w == 0 || w > k) throw Error(MAZU_ERR_INVALID_ARG, "minimizer length w must satisfy 1 <= w <= k");
  if (us.n_kmers() == 0 || us.total_len() < k) throw Error(MAZU_ERR_INVALID_DATA, "unitig set holds no k-mer");
  const int dev = ix.device, sm = ix.sm_count;
  const UnitigsView uv = ix.d_unitigs->view;
  const u64 U = us.n_unitigs();
  CubTemp tmp;
  auto d = std::make_shared<K2UDev>();
  auto keep = [&](const DevBufP& b) {
    d->bufs.push_back(b);
    d->bytes += b->bytes;
  };
  // 1. collect
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, collect_minimizers_kernel<true>, QR_WARPS * 32, 0);
  const int cgrid = (int)std::max<u64>(1, std::min<u64>((U + QR_WARPS - 1) / QR_WARPS, (u64)sm * std::max(occ, 1)));
  DevBuf counts((2 * U + 1) * 8, dev), bases((2 * U + 1) * 8, dev);
  MZ_CUDA(cudaMemset(counts.p, 0, (2 * U + 1) * 8));
  collect_minimizers_kernel<false><<<cgrid, QR_WARPS * 32>>>(uv, w, seed, (u64*)counts.p, nullptr, nullptr, nullptr);
  MZ_CUDA(cudaGetLastError());
  exclusive_scan_u64(tmp, (const u64*)counts.p, (u64*)bases.p, 2 * U + 1);
  u64 n = d2h_value((const u64*)bases.p + 2 * U);
  if (n == 0) throw Error(MAZU_ERR_INVALID_DATA, "no minimizer occurrence collected");
  DevBufP words = std::make_shared<DevBuf>(n * 8, dev), poss = std::make_shared<DevBuf>(n * 8, dev);
  collect_minimizers_kernel<true><<<cgrid, QR_WARPS * 32>>>(uv, w, seed, nullptr, (const u64*)bases.p, (u64*)words->p, (u64*)poss->p);
  MZ_CUDA(cudaGetLastError());
  // 2. stable sort by minimizer word
  DevBufP words_s = std::make_shared<DevBuf>(n * 8, dev), poss_s = std::make_shared<DevBuf>(n * 8, dev);
  sort_pairs_u64(tmp, (const u64*)words->p, (u64*)words_s->p, (const u64*)poss->p, (u64*)poss_s->p, n, (int)std::min<u32>(64, 2 * w));
  words.reset();
  poss.reset();
  // 3. group
  DevBufP mm_set, ranges, gid;
  u64 M = group_sorted_gpu(tmp, (const u64*)words_s->p, nullptr, n, dev, sm, mm_set, ranges, nullptr, gid);
  // 3b. drop the second copy of an entry pushed by both streams (light buckets only), then group what is left
  {
    DevBuf keepf(n * 8, dev), kidx(n * 8, dev);
    mark_repeats_kernel<<<grid_1d(n, sm), 256>>>((const u64*)words_s->p, (const u64*)poss_s->p, (const u64*)gid->p, (const u64*)ranges->p, n,
                                                 skew_param, (u64*)keepf.p);
    MZ_CUDA(cudaGetLastError());
    inclusive_scan_u64(tmp, (const u64*)keepf.p, (u64*)kidx.p, n);
    const u64 n_kept = d2h_value((const u64*)kidx.p + (n - 1));
    if (n_kept != n) {
      DevBufP words_c = std::make_shared<DevBuf>(n_kept * 8, dev), poss_c = std::make_shared<DevBuf>(n_kept * 8, dev);
      compact_pairs_kernel<<<grid_1d(n, sm), 256>>>((const u64*)words_s->p, (const u64*)poss_s->p, (const u64*)keepf.p, (const u64*)kidx.p, n,
                                                    (u64*)words_c->p, (u64*)poss_c->p);
      MZ_CUDA(cudaGetLastError());
      words_s = words_c;
      poss_s = poss_c;
      n = n_kept;
      mm_set.reset();
      ranges.reset();
      gid.reset();
      const u64 M2 = group_sorted_gpu(tmp, (const u64*)words_s->p, nullptr, n, dev, sm, mm_set, ranges, nullptr, gid);
      if (M2 != M) throw Error(MAZU_ERR_OTHER, "internal: dropping repeated entries changed the minimizer set");
    }
  }
  words_s.reset();
  // 4. perfect hash over the minimizer set: the fingerprinted cascade; its value is the minimizer's slot
  DevCascade casc = build_cascade_gpu((const u64*)mm_set->p, M, dev, sm);
  for (auto& b : casc.bufs) keep(b);
  const u64 R = casc.R;
  // 5. bucket sizes in slot order (one, possibly empty, bucket per slot), prefix sum
  DevBuf sizes_by_h((R + 1) * 8, dev), prefix((R + 1) * 8, dev), bad(8, dev);
  MZ_CUDA(cudaMemset(sizes_by_h.p, 0, (R + 1) * 8));
  MZ_CUDA(cudaMemset(bad.p, 0, 8));
  cascade_sizes_kernel<<<grid_1d(M, sm), 256>>>((const u64*)casc.slots->p, (const u64*)ranges->p, M, R, (u64*)sizes_by_h.p, (unsigned long long*)bad.p);
  MZ_CUDA(cudaGetLastError());
  if (d2h_value((const u64*)bad.p) != 0) throw Error(MAZU_ERR_OTHER, "internal: a minimizer was not placed by the GPU-built cascade");
  exclusive_scan_u64(tmp, (const u64*)sizes_by_h.p, (u64*)prefix.p, R + 1);
  const DevBuf& hashes = *casc.slots;
  const DevBuf& fps = *casc.states;
  // 6. scatter positions into bucket order
  DevBuf pos_out(n * 8, dev);
  scatter_positions_kernel<<<grid_1d(n, sm), 256>>>((const u64*)poss_s->p, (const u64*)gid->p, (const u64*)ranges->p, (const u64*)hashes.p,
                                                    (const u64*)prefix.p, n, (u64*)pos_out.p);
  MZ_CUDA(cudaGetLastError());
  // 9a. pack the positions
  u32 pos_width = 1;
  u64 pos_bytes = 0;
  DevBufP pos_packed = pack_gpu(tmp, (const u64*)pos_out.p, n, dev, sm, pos_width, pos_bytes);
  keep(pos_packed);
  // 8. blocked Elias-Fano of the prefix sums (+ the slot states / fingerprints)
  const u64 nE = R + 1, u = d2h_value((const u64*)prefix.p + R);
  u64 l = msb(u / nE);
  if (l == 0) l = 1;
  u32 log_s = 5;
  while (log_s > 0 && ((1ULL << log_s) + 1) * l > 64) --log_s;
  const bool all_exc = ((1ULL << log_s) + 1) * l > 64 || (u >> 63);
  const u64 S = 1ULL << log_s, nbE = (nE + S - 1) / S;
  DevBuf exc_flags((nbE + 1) * 8, dev), exc_index((nbE + 1) * 8, dev);
  MZ_CUDA(cudaMemset(exc_flags.p, 0, (nbE + 1) * 8));
  ef_blocks_kernel<false><<<grid_1d(nbE, sm), 256>>>((const u64*)prefix.p, nE, (u32)l, log_s, 8, all_exc, (const u8*)fps.p, R, (u64*)exc_flags.p, nullptr,
                                                      nullptr, nullptr);
  MZ_CUDA(cudaGetLastError());
  exclusive_scan_u64(tmp, (const u64*)exc_flags.p, (u64*)exc_index.p, nbE + 1);
  const u64 n_exc = d2h_value((const u64*)exc_index.p + nbE);
  DevBufP ef_blocks = std::make_shared<DevBuf>((nbE * 8 + 8) * 8, dev), ef_exc = std::make_shared<DevBuf>(std::max<u64>(1, n_exc * (S + 1)) * 8, dev);
  MZ_CUDA(cudaMemset(ef_blocks->p, 0, (nbE * 8 + 8) * 8));
  MZ_CUDA(cudaMemset(ef_exc->p, 0, std::max<u64>(1, n_exc * (S + 1)) * 8));
  ef_blocks_kernel<true><<<grid_1d(nbE, sm), 256>>>((const u64*)prefix.p, nE, (u32)l, log_s, 8, all_exc, (const u8*)fps.p, R, nullptr,
                                                     (const u64*)exc_index.p, (u64*)ef_blocks->p, (u64*)ef_exc->p);
  MZ_CUDA(cudaGetLastError());
  keep(ef_blocks);
  keep(ef_exc);
  // host-side metadata
  auto H = std::make_shared<K2UHost>();
  H->kind = MAZU_K2U_SSHASH;
  H->unitigs = ix.unitigs;
  H->w = w;
  H->seed = seed;
  H->skew_param = skew_param;
  H->n_minimizers = M;
  H->n_minimizer_occs = n;
  H->sizes.n = nE;
  H->sizes.l = (u32)l;
  H->sizes.log_s = log_s;
  H->sizes.wpb = 8;
  H->sizes.n_exception_blocks = n_exc;
  IndexView& v = ix.view;
  v.k2u_kind = MAZU_K2U_SSHASH;
  v.mphf = casc.view;
  v.pos = PackedVecView{(const u64*)pos_packed->p, n, pos_width, 0};
  v.sizes = BlockedEFView{(const u64*)ef_blocks->p, (const u64*)ef_exc->p, nE, (u32)l, log_s, 8, 0};
  v.w = w;
  v.seed = seed;
  v.skew_param = skew_param;
  v.has_skew = 0;
  ix.tables[0] = {nullptr, 0};  // the cascade has no table of its own: its states ride in the bucket-bound blocks
  ix.tables[1] = {v.sizes.blocks, nbE * 8 * 8};
  ix.tables[2] = {v.sizes.exceptions, n_exc * (S + 1) * 8};
  ix.tables[3] = {v.pos.words, pos_bytes};
  ix.tables[6] = {v.mphf.fb_keys, casc.n_fb * 8};
  // 7. skew index
  if (skew_param != MAZU_SKEW_NONE) {
    H->has_skew = true;
    v.has_skew = 1;
    DevBuf cnts((n + 1) * 8, dev), sb((n + 1) * 8, dev);
    MZ_CUDA(cudaMemset(cnts.p, 0, (n + 1) * 8));
    skew_tuples_kernel<false><<<grid_1d(n, sm), 256>>>(uv, w, skew_param, (const u64*)poss_s->p, (const u64*)gid->p, (const u64*)ranges->p, n,
                                                        (u64*)cnts.p, nullptr, nullptr, nullptr);
    MZ_CUDA(cudaGetLastError());
    exclusive_scan_u64(tmp, (const u64*)cnts.p, (u64*)sb.p, n + 1);
    const u64 T = d2h_value((const u64*)sb.p + n);
    DevMphf smphf;
    smphf.view.family = MPHF_FAMILY_NATIVE;
    u64 Ms = 0, spos_bytes = 0;
    u32 swidth = 1;
    DevBufP spos_packed;
    if (T > 0) {
      DevBufP sw = std::make_shared<DevBuf>(T * 8, dev), sp = std::make_shared<DevBuf>(T * 8, dev);
      skew_tuples_kernel<true><<<grid_1d(n, sm), 256>>>(uv, w, skew_param, (const u64*)poss_s->p, (const u64*)gid->p, (const u64*)ranges->p, n, nullptr,
                                                         (const u64*)sb.p, (u64*)sw->p, (u64*)sp->p);
      MZ_CUDA(cudaGetLastError());
      DevBufP sw_s = std::make_shared<DevBuf>(T * 8, dev), sp_s = std::make_shared<DevBuf>(T * 8, dev);
      sort_pairs_u64(tmp, (const u64*)sw->p, (u64*)sw_s->p, (const u64*)sp->p, (u64*)sp_s->p, T, (int)std::min<u32>(64, 2 * k));
      sw.reset();
      sp.reset();
      DevBufP km_set, ranges2, first_pos, gid2;
      Ms = group_sorted_gpu(tmp, (const u64*)sw_s->p, (const u64*)sp_s->p, T, dev, sm, km_set, ranges2, &first_pos, gid2);  // dedup keeps the first
      smphf = build_mphf_gpu(tmp, (const u64*)km_set->p, Ms, gamma, dev, sm);
      DevBuf hashes2(Ms * 8, dev), svals(Ms * 8, dev);
      MZ_CUDA(cudaMemset(bad.p, 0, 8));
      group_hash_kernel<<<grid_1d(Ms, sm), 256>>>(smphf.view, (const u64*)km_set->p, (const u64*)ranges2->p, Ms, (u64*)hashes2.p, nullptr,
                                                  (unsigned long long*)bad.p);
      MZ_CUDA(cudaGetLastError());
      if (d2h_value((const u64*)bad.p) != 0) throw Error(MAZU_ERR_OTHER, "internal: GPU-built skew MPHF is not a bijection on its keys");
      scatter_by_hash_kernel<<<grid_1d(Ms, sm), 256>>>((const u64*)first_pos->p, (const u64*)hashes2.p, Ms, (u64*)svals.p);
      MZ_CUDA(cudaGetLastError());
      spos_packed = pack_gpu(tmp, (const u64*)svals.p, Ms, dev, sm, swidth, spos_bytes);
    } else {
      smphf = build_mphf_gpu(tmp, nullptr, 0, gamma, dev, sm);
      spos_packed = std::make_shared<DevBuf>(16, dev);
      MZ_CUDA(cudaMemset(spos_packed->p, 0, 16));
    }
    for (auto& b : smphf.bufs) keep(b);
    keep(spos_packed);
    H->n_skew_kmers = Ms;
    v.skew_mphf = smphf.view;
    v.skew_pos = PackedVecView{(const u64*)spos_packed->p, Ms, swidth, 0};
    ix.tables[4] = {v.skew_mphf.blocks, smphf.block_bytes};
    ix.tables[5] = {v.skew_pos.words, spos_bytes};
  }
  MZ_CUDA(cudaDeviceSynchronize());
  ix.k2u = H;
  ix.d_k2u = d;
}

}  // namespace
