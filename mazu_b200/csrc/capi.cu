// extern "C" surface of libmazu_b200.so (include/mazu_b200.h): index upload, kernel launches, and the
// chunked host<->device pipelines behind MAZU_MEM_HOST calls.  There is no CPU fallback: every query
// entry point launches the CUDA kernels of kernels.cuh or fails with MAZU_ERR_CUDA.
#include <cub/device/device_scan.cuh>

#include <functional>
#include <mutex>

#include "formats.hpp"
#include "kernels.cuh"

using namespace mazu;

namespace {

thread_local std::string g_err;

#define MZ_CUDA(expr)                                                                                                   \
  do {                                                                                                                  \
    cudaError_t _e = (expr);                                                                                            \
    if (_e != cudaSuccess) throw Error(MAZU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));              \
  } while (0)

template <class F>
mazu_status_t guarded(F&& f) {
  try {
    f();
    return MAZU_OK;
  } catch (const Error& e) {
    g_err = e.what();
    return e.code;
  } catch (const std::bad_alloc&) {
    g_err = "out of host memory";
    return MAZU_ERR_OTHER;
  } catch (const std::exception& e) {
    g_err = e.what();
    return MAZU_ERR_OTHER;
  }
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    MZ_CUDA(cudaSetDevice(dev));
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int device = 0;
  DevBuf(size_t n, int dev) : bytes(n), device(dev) { MZ_CUDA(cudaMalloc(&p, n ? n : 1)); }
  ~DevBuf() {
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    cudaFree(p);
    if (prev >= 0) cudaSetDevice(prev);
  }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};
using DevBufP = std::shared_ptr<DevBuf>;

template <class T>
DevBufP upload(const T* host, size_t n, int dev, size_t pad_elems = 0) {
  auto b = std::make_shared<DevBuf>((n + pad_elems) * sizeof(T), dev);
  if (pad_elems) MZ_CUDA(cudaMemset(b->p, 0, (n + pad_elems) * sizeof(T)));
  if (n) MZ_CUDA(cudaMemcpy(b->p, host, n * sizeof(T), cudaMemcpyHostToDevice));
  return b;
}
template <class T>
DevBufP upload(const std::vector<T>& v, int dev, size_t pad_elems = 0) {
  return upload(v.data(), v.size(), dev, pad_elems);
}

// a group of device buffers + the view fields they back; groups are shared between handles
// created by rebuild_k2u (the reference clones u2pos / refs there; they are immutable so we share)
struct UnitigsDev {
  std::vector<DevBufP> bufs;
  UnitigsView view{};
  size_t bytes = 0;
};
struct K2UDev {
  std::vector<DevBufP> bufs;
  size_t bytes = 0;
};
struct U2PosDev {
  std::vector<DevBufP> bufs;
  size_t bytes = 0;
};
struct RefsDev {
  std::vector<DevBufP> bufs;
  size_t bytes = 0;
};

}  // namespace

struct mazu_index {
  int device = 0;
  int sm_count = 148;
  std::shared_ptr<const UnitigSetHost> unitigs;
  std::shared_ptr<K2UHost> k2u;
  std::shared_ptr<U2PosHost> u2pos;
  std::shared_ptr<RefSeqHost> refs;
  std::shared_ptr<UnitigsDev> d_unitigs;
  std::shared_ptr<K2UDev> d_k2u;
  std::shared_ptr<U2PosDev> d_u2pos;
  std::shared_ptr<RefsDev> d_refs;
  IndexView view{};
  // named device tables (pointer, logical bytes) for mazu_b200_debug_table_digest: lets a test compare the
  // host-built and the GPU-built index table by table
  std::function<void(mazu_index&)> gpu_builder;  // set when the K2U tables are to be built on the device
  std::vector<std::pair<const void*, u64>> tables = std::vector<std::pair<const void*, u64>>(8, {nullptr, 0});
  // per-handle stream-ordered pool for per-call scratch (scan temporaries, list lengths).  It keeps what it has freed
  // (release threshold = max), so a call after a synchronisation does not pay for fresh physical memory again
  // (measured: 1.5-4 ms on the first decode call after every sync with the default pool's threshold of 0).
  cudaMemPool_t pool = nullptr;
  ~mazu_index() {
    if (pool) cudaMemPoolDestroy(pool);
  }
  u64 device_bytes() const {
    return (d_unitigs ? d_unitigs->bytes : 0) + (d_k2u ? d_k2u->bytes : 0) + (d_u2pos ? d_u2pos->bytes : 0) + (d_refs ? d_refs->bytes : 0);
  }
};

namespace {

static const u32 DIR_SHIFT = 6;  // one directory entry per 64 bases

std::shared_ptr<UnitigsDev> upload_unitigs(const UnitigSetHost& us, int dev) {
  auto d = std::make_shared<UnitigsDev>();
  const u64 L = us.total_len(), U = us.n_unitigs();
  if (U >> 32) throw Error(MAZU_ERR_INVALID_ARG, "more than 2^32 unitigs");
  u64 nw = (2 * L + 63) / 64;
  auto b_seq = upload(us.useq.data(), std::min<u64>(nw, us.useq.size()), dev, 4);
  // directory: unitig containing the first base of every 2^DIR_SHIFT block
  u64 nd = (L >> DIR_SHIFT) + 2;
  std::vector<u32> dir(nd, (u32)(U ? U - 1 : 0));
  {
    u64 ui = 0;
    for (u64 b = 0; b < nd; ++b) {
      u64 p = b << DIR_SHIFT;
      if (p >= L) break;
      while (us.accum[ui + 1] <= p) ++ui;
      dir[b] = (u32)ui;
    }
  }
  auto b_dir = upload(dir, dev);
  auto b_starts = upload(us.accum, dev, 4);
  d->bufs = {b_seq, b_dir, b_starts};
  d->bytes = b_seq->bytes + b_dir->bytes + b_starts->bytes;
  d->view.useq = (const u64*)b_seq->p;
  d->view.dir = (const u32*)b_dir->p;
  d->view.starts = (const u64*)b_starts->p;
  d->view.total_len = L;
  d->view.n_unitigs = U;
  d->view.k = us.k;
  d->view.dir_shift = DIR_SHIFT;
  return d;
}

RankedLevels upload_mphf(const MphfHost& m, int dev, K2UDev& d) {
  RankedLevels v = m.view();
  auto b = upload(m.blocks, dev, 8);
  auto fk = upload(m.fb_keys, dev), fv = upload(m.fb_vals, dev);
  d.bufs.insert(d.bufs.end(), {b, fk, fv});
  d.bytes += b->bytes + fk->bytes + fv->bytes;
  v.blocks = (const u32*)b->p;
  v.fb_keys = (const u64*)fk->p;
  v.fb_vals = (const u64*)fv->p;
  return v;
}
PackedVecView upload_packed(const PackedVec& pv, int dev, std::vector<DevBufP>& bufs, size_t& bytes) {
  auto b = upload(pv.words, dev, 2);
  bufs.push_back(b);
  bytes += b->bytes;
  return PackedVecView{(const u64*)b->p, pv.len, (u32)pv.width, 0};
}

void upload_k2u(mazu_index& ix) {
  auto d = std::make_shared<K2UDev>();
  const K2UHost& h = *ix.k2u;
  IndexView& v = ix.view;
  v.k2u_kind = (u32)h.kind;
  v.mphf = upload_mphf(h.mphf, ix.device, *d);
  v.pos = upload_packed(h.pos, ix.device, d->bufs, d->bytes);
  v.w = h.w;
  v.seed = h.seed;
  v.skew_param = h.skew_param;
  v.has_skew = h.has_skew ? 1u : 0u;
  if (h.kind == MAZU_K2U_SSHASH) {
    auto bb = upload(h.sizes.blocks, ix.device, 8);
    auto be = upload(h.sizes.exceptions, ix.device);
    d->bufs.insert(d->bufs.end(), {bb, be});
    d->bytes += bb->bytes + be->bytes;
    v.sizes = BlockedEFView{(const u64*)bb->p, (const u64*)be->p, h.sizes.n, h.sizes.l, h.sizes.log_s, h.sizes.wpb, 0};
    if (h.has_skew) {
      v.skew_mphf = upload_mphf(h.skew_mphf, ix.device, *d);
      v.skew_pos = upload_packed(h.skew_pos, ix.device, d->bufs, d->bytes);
    }
  }
  if (h.kind == MAZU_K2U_SAMPLED_PFHASH) {
    v.sampled = upload_mphf(h.sampled, ix.device, *d);
    auto bc = upload(h.canonical_bits, ix.device, 2), bd = upload(h.direction_bits, ix.device, 2);
    d->bufs.insert(d->bufs.end(), {bc, bd});
    d->bytes += bc->bytes + bd->bytes;
    v.canonical_bits = (const u64*)bc->p;
    v.direction_bits = (const u64*)bd->p;
    v.ext_sizes = upload_packed(h.ext_sizes, ix.device, d->bufs, d->bytes);
    v.ext_bases = upload_packed(h.ext_bases, ix.device, d->bufs, d->bytes);
    v.extension_size = (u32)h.extension_size;
  }
  ix.d_k2u = d;
  auto packed_bytes = [](const PackedVec& pv) { return ((pv.len * pv.width + 63) / 64) * 8; };
  ix.tables[0] = {v.mphf.blocks, h.mphf.blocks.size() * 4};
  ix.tables[3] = {v.pos.words, packed_bytes(h.pos)};
  ix.tables[6] = {v.mphf.fb_keys, h.mphf.n_fb_real() * 8};
  if (h.kind == MAZU_K2U_SSHASH) {
    ix.tables[1] = {v.sizes.blocks, h.sizes.blocks.size() * 8};
    ix.tables[2] = {v.sizes.exceptions, h.sizes.n_exception_blocks * ((1ULL << h.sizes.log_s) + 1) * 8};
    if (h.has_skew) {
      ix.tables[4] = {v.skew_mphf.blocks, h.skew_mphf.blocks.size() * 4};
      ix.tables[5] = {v.skew_pos.words, packed_bytes(h.skew_pos)};
    }
  }
}
void upload_u2pos(mazu_index& ix) {
  IndexView& v = ix.view;
  v.u2pos_kind = MAZU_U2POS_NONE;
  if (!ix.u2pos || ix.u2pos->kind == MAZU_U2POS_NONE) return;
  if (!ix.d_u2pos) {
    auto d = std::make_shared<U2PosDev>();
    auto b = upload(ix.u2pos->ctable_words, ix.device, 4);
    d->bufs.push_back(b);
    d->bytes += b->bytes;
    upload_packed(ix.u2pos->contig_offsets, ix.device, d->bufs, d->bytes);
    ix.d_u2pos = d;
  }
  const U2PosHost& u = *ix.u2pos;
  v.u2pos_kind = (u32)u.kind;
  v.ctable_words = (const u64*)ix.d_u2pos->bufs[0]->p;
  v.n_occs = u.n_occs;
  v.ctable_width = u.ctable_width;
  v.ref_shift = (u32)u.ref_shift;
  v.pos_mask = u.pos_mask;
  v.contig_offsets = PackedVecView{(const u64*)ix.d_u2pos->bufs[1]->p, u.contig_offsets.len, (u32)u.contig_offsets.width, 0};
}
void upload_refs(mazu_index& ix) {
  IndexView& v = ix.view;
  v.refseq = nullptr;
  v.ref_prefix = nullptr;
  v.n_refs = ix.refs ? ix.refs->n_refs() : 0;
  if (!ix.refs || !ix.refs->has_seq) return;
  if (!ix.d_refs) {
    auto d = std::make_shared<RefsDev>();
    auto bs = upload(ix.refs->seq_words, ix.device, 2);
    auto bp = upload(ix.refs->prefix, ix.device);
    d->bufs = {bs, bp};
    d->bytes = bs->bytes + bp->bytes;
    ix.d_refs = d;
  }
  v.refseq = (const u64*)ix.d_refs->bufs[0]->p;
  v.ref_prefix = (const u64*)ix.d_refs->bufs[1]->p;
}

}  // namespace
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
  const u64 n = d2h_value((const u64*)bases.p + 2 * U);
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
  const u64 M = group_sorted_gpu(tmp, (const u64*)words_s->p, nullptr, n, dev, sm, mm_set, ranges, nullptr, gid);
  words_s.reset();
  // 4. MPHF over the minimizer set
  DevMphf mphf = build_mphf_gpu(tmp, (const u64*)mm_set->p, M, gamma, dev, sm);
  for (auto& b : mphf.bufs) keep(b);
  // 5. bucket sizes in MPHF order, fingerprints, prefix sum
  DevBuf hashes(M * 8, dev), sizes_by_h((M + 1) * 8, dev), prefix((M + 1) * 8, dev), fps(M + 1, dev), bad(8, dev);
  MZ_CUDA(cudaMemset(sizes_by_h.p, 0, (M + 1) * 8));
  MZ_CUDA(cudaMemset(bad.p, 0, 8));
  group_hash_kernel<<<grid_1d(M, sm), 256>>>(mphf.view, (const u64*)mm_set->p, (const u64*)ranges->p, M, (u64*)hashes.p, (u64*)sizes_by_h.p,
                                             (u8*)fps.p, (unsigned long long*)bad.p);
  MZ_CUDA(cudaGetLastError());
  if (d2h_value((const u64*)bad.p) != 0) throw Error(MAZU_ERR_OTHER, "internal: GPU-built MPHF is not a bijection on its keys");
  exclusive_scan_u64(tmp, (const u64*)sizes_by_h.p, (u64*)prefix.p, M + 1);
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
  // 8. blocked Elias-Fano of the prefix sums (+ fingerprints)
  const u64 nE = M + 1, u = d2h_value((const u64*)prefix.p + M);
  u64 l = msb(u / nE);
  if (l == 0) l = 1;
  u32 log_s = 5;
  while (log_s > 0 && ((1ULL << log_s) + 1) * l > 64) --log_s;
  const bool all_exc = ((1ULL << log_s) + 1) * l > 64 || (u >> 63);
  const u64 S = 1ULL << log_s, nbE = (nE + S - 1) / S;
  DevBuf exc_flags((nbE + 1) * 8, dev), exc_index((nbE + 1) * 8, dev);
  MZ_CUDA(cudaMemset(exc_flags.p, 0, (nbE + 1) * 8));
  ef_blocks_kernel<false><<<grid_1d(nbE, sm), 256>>>((const u64*)prefix.p, nE, (u32)l, log_s, 8, all_exc, (const u8*)fps.p, M, (u64*)exc_flags.p, nullptr,
                                                      nullptr, nullptr);
  MZ_CUDA(cudaGetLastError());
  exclusive_scan_u64(tmp, (const u64*)exc_flags.p, (u64*)exc_index.p, nbE + 1);
  const u64 n_exc = d2h_value((const u64*)exc_index.p + nbE);
  DevBufP ef_blocks = std::make_shared<DevBuf>((nbE * 8 + 8) * 8, dev), ef_exc = std::make_shared<DevBuf>(std::max<u64>(1, n_exc * (S + 1)) * 8, dev);
  MZ_CUDA(cudaMemset(ef_blocks->p, 0, (nbE * 8 + 8) * 8));
  MZ_CUDA(cudaMemset(ef_exc->p, 0, std::max<u64>(1, n_exc * (S + 1)) * 8));
  ef_blocks_kernel<true><<<grid_1d(nbE, sm), 256>>>((const u64*)prefix.p, nE, (u32)l, log_s, 8, all_exc, (const u8*)fps.p, M, nullptr,
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
  v.mphf = mphf.view;
  v.pos = PackedVecView{(const u64*)pos_packed->p, n, pos_width, 0};
  v.sizes = BlockedEFView{(const u64*)ef_blocks->p, (const u64*)ef_exc->p, nE, (u32)l, log_s, 8, 0};
  v.w = w;
  v.seed = seed;
  v.skew_param = skew_param;
  v.has_skew = 0;
  ix.tables[0] = {v.mphf.blocks, mphf.block_bytes};
  ix.tables[1] = {v.sizes.blocks, nbE * 8 * 8};
  ix.tables[2] = {v.sizes.exceptions, n_exc * (S + 1) * 8};
  ix.tables[3] = {v.pos.words, pos_bytes};
  ix.tables[6] = {v.mphf.fb_keys, mphf.n_fb * 8};
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
      group_hash_kernel<<<grid_1d(Ms, sm), 256>>>(smphf.view, (const u64*)km_set->p, (const u64*)ranges2->p, Ms, (u64*)hashes2.p, nullptr, nullptr,
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

mazu_index* finalize_index(std::unique_ptr<mazu_index> ix) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) throw Error(MAZU_ERR_CUDA, "no CUDA device visible: libmazu_b200 has no CPU fallback");
  if (ix->device < 0 || ix->device >= n) throw Error(MAZU_ERR_INVALID_ARG, "device ordinal out of range");
  DeviceGuard g(ix->device);
  cudaDeviceProp prop;
  MZ_CUDA(cudaGetDeviceProperties(&prop, ix->device));
  ix->sm_count = prop.multiProcessorCount;
  {
    cudaMemPoolProps pp{};
    pp.allocType = cudaMemAllocationTypePinned;
    pp.handleTypes = cudaMemHandleTypeNone;
    pp.location.type = cudaMemLocationTypeDevice;
    pp.location.id = ix->device;
    MZ_CUDA(cudaMemPoolCreate(&ix->pool, &pp));
    unsigned long long keep = ~0ULL;
    MZ_CUDA(cudaMemPoolSetAttribute(ix->pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  if (!ix->d_unitigs) ix->d_unitigs = upload_unitigs(*ix->unitigs, ix->device);
  ix->view.unitigs = ix->d_unitigs->view;
  if (ix->gpu_builder) ix->gpu_builder(*ix);  // tables are produced in HBM (gpu_build.cuh)
  else upload_k2u(*ix);
  upload_u2pos(*ix);
  upload_refs(*ix);
  MZ_CUDA(cudaDeviceSynchronize());
  return ix.release();
}

mazu_index* index_from_loaded(LoadedIndex&& L, int device) {
  auto ix = std::make_unique<mazu_index>();
  ix->device = device;
  ix->unitigs = L.unitigs;
  ix->k2u = L.k2u;
  ix->u2pos = L.u2pos;
  ix->refs = L.refs;
  return finalize_index(std::move(ix));
}

template <class K>
int grid_for(K kernel, int block, const mazu_index* ix, u64 work_items_per_block_hint, u64 n_items) {
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, 0) != cudaSuccess || occ < 1) occ = 1;
  u64 want = (n_items + work_items_per_block_hint - 1) / work_items_per_block_hint;
  u64 cap = (u64)ix->sm_count * (u64)occ;
  return (int)std::max<u64>(1, std::min(want, cap));
}

void check_k(const mazu_index* ix, u32 k) {
  if (k != ix->unitigs->k)
    throw Error(MAZU_ERR_K_MISMATCH, "Got query k-mer size k=" + std::to_string(k) + ", expected k=" + std::to_string(ix->unitigs->k) + ".");
}

void launch_k2u_batch(const mazu_index* ix, const u64* d_words, u64 n, Hit* d_out, cudaStream_t s) {
  if (n == 0) return;
  int grid = grid_for(k2u_batch_kernel, 256, ix, 256, n);
  k2u_batch_kernel<<<grid, 256, 0, s>>>(ix->view, d_words, n, d_out);
  MZ_CUDA(cudaGetLastError());
}

void device_exclusive_scan(const u64* d_in, u64* d_out, u64 n, cudaMemPool_t pool, cudaStream_t s);

template <int MODE, int KIND, u32 FAMILY, int OCC>
void launch_qr_occ(const mazu_index* ix, const u8* d_bases, const u64* d_read_offsets, u64 n_reads, u64 uniform_len, const u64* d_kmer_offsets,
                   void* d_out, u32 compact, u64* d_counts, const u64* d_seg_offsets, cudaStream_t s) {
  auto kern = query_reads_kernel<MODE, KIND, FAMILY, OCC>;
  // with a segment table the number of work items is only known on the device: launch every resident CTA, idle warps leave at once
  int grid = grid_for(kern, QR_WARPS * 32, ix, QR_WARPS, d_seg_offsets ? ~0ULL >> 8 : n_reads);
  kern<<<grid, QR_WARPS * 32, 0, s>>>(ix->view, d_bases, d_read_offsets, n_reads, uniform_len, d_kmer_offsets, d_out, compact,
                                      (unsigned long long*)d_counts, d_seg_offsets);
}
template <int MODE, int KIND, u32 FAMILY>
void launch_qr(const mazu_index* ix, const u8* d_bases, const u64* d_read_offsets, u64 n_reads, u64 uniform_len, const u64* d_kmer_offsets,
               void* d_out, u32 compact, u64* d_counts, cudaStream_t s) {
  // streaming walk: 3 resident CTAs (80 registers) while the index sits in the 126 MB L2, 4 (64 registers) once it does not
  if constexpr (MODE == 1) {
    if (ix->device_bytes() <= (96ull << 20))
      launch_qr_occ<MODE, KIND, FAMILY, 3>(ix, d_bases, d_read_offsets, n_reads, uniform_len, d_kmer_offsets, d_out, compact, d_counts, nullptr, s);
    else
      launch_qr_occ<MODE, KIND, FAMILY, 4>(ix, d_bases, d_read_offsets, n_reads, uniform_len, d_kmer_offsets, d_out, compact, d_counts, nullptr, s);
  } else {
    // ragged batches: per-read work-item counts -> scan; the kernel cuts long reads into segments (kernels.cuh, QR_SEGMENT)
    void *cnt = nullptr, *seg = nullptr;
    if (!uniform_len) {
      MZ_CUDA(cudaMallocFromPoolAsync(&cnt, (n_reads + 1) * 8, ix->pool, s));
      MZ_CUDA(cudaMallocFromPoolAsync(&seg, (n_reads + 1) * 8, ix->pool, s));
      MZ_CUDA(cudaMemsetAsync(cnt, 0, (n_reads + 1) * 8, s));
      segment_counts_kernel<<<(int)std::min<u64>((n_reads + 255) / 256, 4096), 256, 0, s>>>(d_read_offsets, n_reads, ix->unitigs->k, (u64*)cnt);
      MZ_CUDA(cudaGetLastError());
      device_exclusive_scan((const u64*)cnt, (u64*)seg, n_reads, ix->pool, s);
    }
    launch_qr_occ<MODE, KIND, FAMILY, 0>(ix, d_bases, d_read_offsets, n_reads, uniform_len, d_kmer_offsets, d_out, compact, d_counts, (const u64*)seg, s);
    if (cnt) MZ_CUDA(cudaFreeAsync(cnt, s));
    if (seg) MZ_CUDA(cudaFreeAsync(seg, s));
  }
}

void launch_query_reads(const mazu_index* ix, const u8* d_bases, const u64* d_read_offsets, u64 n_reads, u64 uniform_len, int mode,
                        const u64* d_kmer_offsets, void* d_out, u32 compact, u64* d_counts, cudaStream_t s) {
  if (n_reads == 0) return;
  const bool ss = ix->view.k2u_kind == MAZU_K2U_SSHASH;
  const bool native = ix->view.mphf.family == MPHF_FAMILY_NATIVE;
  const bool st = mode == MAZU_MODE_STREAMING;
#define MZ_QR(M, K, F) launch_qr<M, K, F>(ix, d_bases, d_read_offsets, n_reads, uniform_len, d_kmer_offsets, d_out, compact, d_counts, s)
  if (ss) {
    if (!native) throw Error(MAZU_ERR_OTHER, "internal: SSHash index without a native MPHF");
    if (st) MZ_QR(1, MAZU_K2U_SSHASH, MPHF_FAMILY_NATIVE); else MZ_QR(0, MAZU_K2U_SSHASH, MPHF_FAMILY_NATIVE);
  } else if (ix->view.k2u_kind == MAZU_K2U_SAMPLED_PFHASH) {
    if (st) MZ_QR(1, MAZU_K2U_SAMPLED_PFHASH, MPHF_FAMILY_BOOPHF); else MZ_QR(0, MAZU_K2U_SAMPLED_PFHASH, MPHF_FAMILY_BOOPHF);
  } else if (native) {
    if (st) MZ_QR(1, MAZU_K2U_PFHASH, MPHF_FAMILY_NATIVE); else MZ_QR(0, MAZU_K2U_PFHASH, MPHF_FAMILY_NATIVE);
  } else {
    if (st) MZ_QR(1, MAZU_K2U_PFHASH, MPHF_FAMILY_BOOPHF); else MZ_QR(0, MAZU_K2U_PFHASH, MPHF_FAMILY_BOOPHF);
  }
#undef MZ_QR
  MZ_CUDA(cudaGetLastError());
}

// exclusive scan with the total appended: out[0..n] from in[0..n)
void device_exclusive_scan(const u64* d_in, u64* d_out, u64 n, cudaMemPool_t pool, cudaStream_t s) {
  // scan n+1 items where the last input is ignored: simplest is an exclusive scan over n items plus one tail kernel-free trick:
  // run ExclusiveSum over n+1 elements with in[n] readable (callers allocate n+1).
  size_t tmp_bytes = 0;
  MZ_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_in, d_out, (size_t)(n + 1), s));
  void* tmp = nullptr;
  MZ_CUDA(cudaMallocFromPoolAsync(&tmp, tmp_bytes ? tmp_bytes : 1, pool, s));
  MZ_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_in, d_out, (size_t)(n + 1), s));
  MZ_CUDA(cudaFreeAsync(tmp, s));
}

struct StreamPair {
  cudaStream_t s[2] = {nullptr, nullptr};
  StreamPair() {
    for (auto& x : s) MZ_CUDA(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
  }
  ~StreamPair() {
    for (auto& x : s)
      if (x) cudaStreamDestroy(x);
  }
};

// Per-call device scratch of the host-buffer paths, taken from the handle's stream-ordered pool on stream `s` and given
// back on destruction (after the caller has synchronised its streams).  No cudaMalloc / cudaFree per call: both
// synchronise the whole device, which serialised concurrent callers of one handle and cost milliseconds per batch.
struct PoolScratch {
  cudaMemPool_t pool;
  cudaStream_t s;
  std::vector<void*> ptrs;
  PoolScratch(cudaMemPool_t p, cudaStream_t st) : pool(p), s(st) {}
  void* get(size_t n) {
    void* p = nullptr;
    MZ_CUDA(cudaMallocFromPoolAsync(&p, n ? n : 1, pool, s));
    ptrs.push_back(p);
    return p;
  }
  // make allocations done so far usable on `other`
  void publish(cudaStream_t other) {
    cudaEvent_t e;
    MZ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    MZ_CUDA(cudaEventRecord(e, s));
    MZ_CUDA(cudaStreamWaitEvent(other, e, 0));
    cudaEventDestroy(e);
  }
  ~PoolScratch() {
    for (void* p : ptrs) cudaFreeAsync(p, s);
  }
  PoolScratch(const PoolScratch&) = delete;
  PoolScratch& operator=(const PoolScratch&) = delete;
};

}  // namespace

// =============================================================================================
extern "C" {

const char* mazu_b200_last_error(void) { return g_err.c_str(); }

int32_t mazu_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

mazu_status_t mazu_b200_dense_index_deserialize_from_cpp(const char* dir, int32_t device, mazu_index_t** out) {
  return guarded([&] {
    if (!dir || !out) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    *out = index_from_loaded(load_pf1_dense(dir), device);
  });
}

mazu_status_t mazu_b200_sparse_index_deserialize_from_cpp(const char* dir, int32_t device, mazu_index_t** out) {
  return guarded([&] {
    if (!dir || !out) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    *out = index_from_loaded(load_pf1_sparse(dir), device);
  });
}

mazu_status_t mazu_b200_index_from_cf_prefix(const char* prefix, int32_t index_kind, uint32_t w, uint64_t skew_param, uint64_t hash_seed,
                                             int32_t device, mazu_index_t** out) {
  return guarded([&] {
    if (!prefix || !out) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    *out = index_from_loaded(load_cf_prefix(prefix, index_kind, w, skew_param, hash_seed), device);
  });
}

mazu_status_t mazu_b200_index_create_sshash(const mazu_unitig_set_desc_t* unitigs, uint32_t w, uint64_t skew_param, uint64_t hash_seed,
                                            int32_t device, mazu_index_t** out) {
  return guarded([&] {
    if (!unitigs || !out) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    LoadedIndex L;
    L.unitigs = std::make_shared<UnitigSetHost>(UnitigSetHost::from_desc(*unitigs));
    L.k2u = build_sshash(L.unitigs, w, skew_param, hash_seed);
    *out = index_from_loaded(std::move(L), device);
  });
}

mazu_status_t mazu_b200_index_create_sshash_gpu(const mazu_unitig_set_desc_t* unitigs, uint32_t w, uint64_t skew_param, uint64_t hash_seed,
                                                int32_t device, mazu_index_t** out) {
  return guarded([&] {
    if (!unitigs || !out) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    auto ix = std::make_unique<mazu_index>();
    ix->device = device;
    ix->unitigs = std::make_shared<UnitigSetHost>(UnitigSetHost::from_desc(*unitigs));
    ix->gpu_builder = [=](mazu_index& m) { build_sshash_gpu(m, w, skew_param, hash_seed); };
    *out = finalize_index(std::move(ix));
  });
}

mazu_status_t mazu_b200_index_create_pfhash_gpu(const mazu_unitig_set_desc_t* unitigs, int32_t device, mazu_index_t** out) {
  return guarded([&] {
    if (!unitigs || !out) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    auto ix = std::make_unique<mazu_index>();
    ix->device = device;
    ix->unitigs = std::make_shared<UnitigSetHost>(UnitigSetHost::from_desc(*unitigs));
    ix->gpu_builder = [](mazu_index& m) { build_pfhash_gpu(m); };
    *out = finalize_index(std::move(ix));
  });
}

mazu_status_t mazu_b200_debug_table_digest(const mazu_index_t* idx, int32_t which, uint64_t* digest, uint64_t* n_bytes) {
  return guarded([&] {
    if (!idx || !digest || !n_bytes || which < 0 || which >= (int)idx->tables.size()) throw Error(MAZU_ERR_INVALID_ARG, "bad argument");
    DeviceGuard g(idx->device);
    const auto& t = idx->tables[which];
    std::vector<u8> host(t.second);
    if (t.second) MZ_CUDA(cudaMemcpy(host.data(), t.first, t.second, cudaMemcpyDeviceToHost));
    u64 hsh = 0xcbf29ce484222325ULL;  // FNV-1a over the table bytes
    for (u8 b : host) hsh = (hsh ^ b) * 0x100000001b3ULL;
    *digest = hsh;
    *n_bytes = t.second;
  });
}

mazu_status_t mazu_b200_index_create_pfhash(const mazu_unitig_set_desc_t* unitigs, int32_t device, mazu_index_t** out) {
  return guarded([&] {
    if (!unitigs || !out) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    LoadedIndex L;
    L.unitigs = std::make_shared<UnitigSetHost>(UnitigSetHost::from_desc(*unitigs));
    L.k2u = build_pfhash(L.unitigs);
    *out = index_from_loaded(std::move(L), device);
  });
}

mazu_status_t mazu_b200_index_create_pfhash_from_parts(const mazu_unitig_set_desc_t* unitigs, const mazu_boophf_desc_t* mphf,
                                                       const mazu_packed_vec_desc_t* pos, int32_t device, mazu_index_t** out) {
  return guarded([&] {
    if (!unitigs || !mphf || !pos || !out) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    LoadedIndex L;
    L.unitigs = std::make_shared<UnitigSetHost>(UnitigSetHost::from_desc(*unitigs));
    L.k2u = pfhash_from_parts(L.unitigs, *mphf, *pos);
    *out = index_from_loaded(std::move(L), device);
  });
}

mazu_status_t mazu_b200_index_rebuild_k2u(const mazu_index_t* src, int32_t k2u_kind, uint32_t w, uint64_t skew_param, uint64_t hash_seed,
                                          mazu_index_t** out) {
  return guarded([&] {
    if (!src || !out) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    auto ix = std::make_unique<mazu_index>();
    ix->device = src->device;
    ix->unitigs = src->unitigs;
    ix->u2pos = src->u2pos;
    ix->refs = src->refs;
    ix->d_unitigs = src->d_unitigs;
    ix->d_u2pos = src->d_u2pos;
    ix->d_refs = src->d_refs;
    if (k2u_kind == MAZU_K2U_SSHASH) ix->k2u = build_sshash(ix->unitigs, w, skew_param, hash_seed);
    else if (k2u_kind == MAZU_K2U_PFHASH) ix->k2u = build_pfhash(ix->unitigs);
    else throw Error(MAZU_ERR_INVALID_ARG, "unknown k2u kind");
    *out = finalize_index(std::move(ix));
  });
}

mazu_status_t mazu_b200_index_attach_u2pos_dense(mazu_index_t* idx, const uint64_t* ctable, uint64_t n_occs,
                                                 const mazu_packed_vec_desc_t* contig_offsets) {
  return guarded([&] {
    if (!idx || !ctable || !contig_offsets) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    if (contig_offsets->len != idx->unitigs->n_unitigs() + 1) throw Error(MAZU_ERR_INVALID_DATA, "contig_offsets.len != n_unitigs + 1");
    auto u = std::make_shared<U2PosHost>();
    u->kind = MAZU_U2POS_DENSE;
    u->ctable_words.assign(ctable, ctable + n_occs);
    u->n_occs = n_occs;
    u->ctable_width = 64;
    u->contig_offsets = PackedVec::from_desc(*contig_offsets);
    DeviceGuard g(idx->device);
    idx->u2pos = u;
    idx->d_u2pos.reset();
    upload_u2pos(*idx);
  });
}

mazu_status_t mazu_b200_index_attach_u2pos_piscem(mazu_index_t* idx, const mazu_packed_vec_desc_t* ctable, uint64_t ref_shift,
                                                  uint64_t pos_mask, const mazu_packed_vec_desc_t* contig_offsets) {
  return guarded([&] {
    if (!idx || !ctable || !contig_offsets) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    if (contig_offsets->len != idx->unitigs->n_unitigs() + 1) throw Error(MAZU_ERR_INVALID_DATA, "contig_offsets.len != n_unitigs + 1");
    if (ref_shift >= 64) throw Error(MAZU_ERR_INVALID_ARG, "ref_shift must be < 64");
    auto u = std::make_shared<U2PosHost>();
    u->kind = MAZU_U2POS_PISCEM;
    PackedVec ct = PackedVec::from_desc(*ctable);
    u->ctable_words = ct.words;
    u->n_occs = ct.len;
    u->ctable_width = (u32)ct.width;
    u->ref_shift = ref_shift;
    u->pos_mask = pos_mask;
    u->contig_offsets = PackedVec::from_desc(*contig_offsets);
    DeviceGuard g(idx->device);
    idx->u2pos = u;
    idx->d_u2pos.reset();
    upload_u2pos(*idx);
  });
}

mazu_status_t mazu_b200_index_attach_refseq(mazu_index_t* idx, const uint64_t* seq_words, const uint64_t* prefix_sum, uint64_t n_refs) {
  return guarded([&] {
    if (!idx || !prefix_sum) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    auto r = std::make_shared<RefSeqHost>();
    r->prefix.assign(prefix_sum, prefix_sum + n_refs + 1);
    if (seq_words) {
      r->has_seq = true;
      u64 nw = (2 * r->prefix.back() + 63) / 64;
      r->seq_words.assign(seq_words, seq_words + nw);
    }
    DeviceGuard g(idx->device);
    idx->refs = r;
    idx->d_refs.reset();
    upload_refs(*idx);
  });
}

void mazu_b200_index_destroy(mazu_index_t* idx) { delete idx; }

mazu_status_t mazu_b200_alloc_pinned(uint64_t bytes, void** out) {
  return guarded([&] {
    if (!out) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    MZ_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
  });
}
void mazu_b200_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}

mazu_status_t mazu_b200_index_release_scratch(mazu_index_t* idx, uint64_t* released) {
  return guarded([&] {
    if (!idx) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    DeviceGuard g(idx->device);
    unsigned long long before = 0, after = 0;
    MZ_CUDA(cudaMemPoolGetAttribute(idx->pool, cudaMemPoolAttrReservedMemCurrent, &before));
    MZ_CUDA(cudaDeviceSynchronize());  // frees queued by earlier calls become visible to the trim
    MZ_CUDA(cudaMemPoolTrimTo(idx->pool, 0));
    MZ_CUDA(cudaMemPoolGetAttribute(idx->pool, cudaMemPoolAttrReservedMemCurrent, &after));
    if (released) *released = before > after ? before - after : 0;
  });
}

uint64_t mazu_b200_index_info(const mazu_index_t* idx, int32_t what) {
  if (!idx) return 0;
  switch (what) {
    case MAZU_INFO_K: return idx->unitigs->k;
    case MAZU_INFO_N_UNITIGS: return idx->unitigs->n_unitigs();
    case MAZU_INFO_N_KMERS: return idx->unitigs->n_kmers();
    case MAZU_INFO_SUM_UNITIGS_LEN: return idx->unitigs->total_len();
    case MAZU_INFO_N_MINIMIZERS: return idx->k2u->kind == MAZU_K2U_SSHASH ? idx->k2u->sizes.n : 0;  // len of the prefix sum (sshash.rs:333-335)
    case MAZU_INFO_N_KMERS_IN_SKEW_INDEX: return idx->k2u->n_skew_kmers;
    case MAZU_INFO_N_REFS: return idx->refs ? idx->refs->n_refs() : 0;
    case MAZU_INFO_N_TOTAL_OCCS: return idx->u2pos ? idx->u2pos->n_occs : 0;
    case MAZU_INFO_K2U_KIND: return (u64)idx->k2u->kind;
    case MAZU_INFO_U2POS_KIND: return idx->u2pos ? (u64)idx->u2pos->kind : 0;
    case MAZU_INFO_DEVICE_BYTES: return idx->device_bytes();
    case MAZU_INFO_W: return idx->k2u->w;
    case MAZU_INFO_N_MINIMIZER_OCCS: return idx->k2u->n_minimizer_occs;
    case MAZU_INFO_MPHF_LEVELS: return idx->view.mphf.n_levels;
    case MAZU_INFO_DEVICE: return (u64)idx->device;
    case MAZU_INFO_SAMPLE_SIZE: return idx->k2u->sample_size;
    case MAZU_INFO_EXTENSION_SIZE: return idx->k2u->extension_size;
  }
  return 0;
}

mazu_status_t mazu_b200_unitig_len(const mazu_index_t* idx, uint64_t unitig_id, uint64_t* len, uint64_t* start_pos) {
  return guarded([&] {
    if (!idx) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    if (unitig_id >= idx->unitigs->n_unitigs()) throw Error(MAZU_ERR_INVALID_ARG, "unitig id out of range");
    if (len) *len = idx->unitigs->unitig_len(unitig_id);
    if (start_pos) *start_pos = idx->unitigs->accum[unitig_id];
  });
}

// ---------------------------------------------------------------------------------------------
mazu_status_t mazu_b200_k2u_batch(const mazu_index_t* idx, const uint64_t* fw_words, uint64_t n, uint32_t k, mazu_hit_t* out_hits,
                                  int32_t mem, void* stream) {
  return guarded([&] {
    if (!idx || (n && (!fw_words || !out_hits))) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    check_k(idx, k);
    DeviceGuard g(idx->device);
    if (mem != MAZU_MEM_HOST && mem != MAZU_MEM_DEVICE) throw Error(MAZU_ERR_INVALID_ARG, "unknown mem mode");
    if (mem == MAZU_MEM_DEVICE) {
      launch_k2u_batch(idx, fw_words, n, (Hit*)out_hits, (cudaStream_t)stream);
      return;
    }
    // host buffers: double-buffered chunks, H2D -> kernel -> D2H on two streams
    const u64 CH = 1ull << 22;
    StreamPair sp;
    PoolScratch scratch(idx->pool, sp.s[0]);
    void* din[2] = {scratch.get(std::min(n, CH) * 8), scratch.get(std::min(n, CH) * 8)};
    void* dout[2] = {scratch.get(std::min(n, CH) * 16), scratch.get(std::min(n, CH) * 16)};
    scratch.publish(sp.s[1]);
    int b = 0;
    for (u64 o = 0; o < n; o += CH, b ^= 1) {
      u64 m = std::min(CH, n - o);
      MZ_CUDA(cudaMemcpyAsync(din[b], fw_words + o, m * 8, cudaMemcpyHostToDevice, sp.s[b]));
      launch_k2u_batch(idx, (const u64*)din[b], m, (Hit*)dout[b], sp.s[b]);
      MZ_CUDA(cudaMemcpyAsync(out_hits + o, dout[b], m * 16, cudaMemcpyDeviceToHost, sp.s[b]));
    }
    MZ_CUDA(cudaStreamSynchronize(sp.s[1]));
    MZ_CUDA(cudaStreamSynchronize(sp.s[0]));
  });
}

uint64_t mazu_b200_count_kmer_slots(const mazu_index_t* idx, const uint64_t* read_offsets, uint64_t n_reads, uint64_t uniform_read_len) {
  if (!idx) return 0;
  const u64 k = idx->unitigs->k;
  if (uniform_read_len) return uniform_read_len >= k ? n_reads * (uniform_read_len - k + 1) : 0;
  if (!read_offsets) return 0;
  u64 acc = 0;
  for (u64 r = 0; r < n_reads; ++r) {
    u64 len = read_offsets[r + 1] - read_offsets[r];
    if (len >= k) acc += len - k + 1;
  }
  return acc;
}

static mazu_status_t query_reads_impl(const mazu_index_t* idx, const uint8_t* bases, const uint64_t* read_offsets, uint64_t n_reads,
                                      uint64_t uniform_read_len, int32_t mode, uint64_t* kmer_offsets, void* out_hits, uint32_t compact,
                                      uint64_t* counts, int32_t mem, void* stream) {
  const u64 rec = compact ? 8 : 16;
  return guarded([&] {
    if (!idx) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    if (mode != MAZU_MODE_RANDOM && mode != MAZU_MODE_STREAMING) throw Error(MAZU_ERR_INVALID_ARG, "unknown query mode");
    if (n_reads && !bases) throw Error(MAZU_ERR_INVALID_ARG, "null bases");
    if (n_reads && !uniform_read_len && !read_offsets) throw Error(MAZU_ERR_INVALID_ARG, "read_offsets is required for ragged reads");
    const u32 k = idx->unitigs->k;
    DeviceGuard g(idx->device);
    if (mem == MAZU_MEM_DEVICE) {
      cudaStream_t s = (cudaStream_t)stream;
      u64* d_koffs = kmer_offsets;
      void* tmp_koffs = nullptr;
      if (!uniform_read_len && n_reads) {
        if (!d_koffs) {
          MZ_CUDA(cudaMallocFromPoolAsync(&tmp_koffs, (n_reads + 1) * 8, idx->pool, s));
          d_koffs = (u64*)tmp_koffs;
        }
        void* lens = nullptr;
        MZ_CUDA(cudaMallocFromPoolAsync(&lens, (n_reads + 1) * 8, idx->pool, s));
        MZ_CUDA(cudaMemsetAsync(lens, 0, (n_reads + 1) * 8, s));
        kmer_counts_kernel<<<(int)std::min<u64>((n_reads + 255) / 256, 4096), 256, 0, s>>>(read_offsets, n_reads, k, (u64*)lens);
        MZ_CUDA(cudaGetLastError());
        device_exclusive_scan((const u64*)lens, d_koffs, n_reads, idx->pool, s);
        MZ_CUDA(cudaFreeAsync(lens, s));
      }
      launch_query_reads(idx, bases, read_offsets, n_reads, uniform_read_len, mode, d_koffs, out_hits, compact, counts, s);
      if (tmp_koffs) MZ_CUDA(cudaFreeAsync(tmp_koffs, s));
      return;
    }
    if (mem != MAZU_MEM_HOST && mem != MAZU_MEM_HOST_IN_DEVICE_OUT) throw Error(MAZU_ERR_INVALID_ARG, "unknown mem mode");
    const bool dev_out = mem == MAZU_MEM_HOST_IN_DEVICE_OUT;  // records are written straight into the caller's device buffer
    // ---- host buffers: chunk the batch, overlap H2D / kernel / D2H on two streams ----
    std::vector<u64> koffs_local;
    const u64* koffs = nullptr;
    if (!uniform_read_len) {
      u64* dst = kmer_offsets;
      if (!dst) {
        koffs_local.resize(n_reads + 1);
        dst = koffs_local.data();
      }
      u64 acc = 0;
      for (u64 r = 0; r < n_reads; ++r) {
        dst[r] = acc;
        u64 len = read_offsets[r + 1] - read_offsets[r];
        if (len >= k) acc += len - k + 1;
      }
      dst[n_reads] = acc;
      koffs = dst;
    } else if (kmer_offsets) {
      u64 per = uniform_read_len >= k ? uniform_read_len - k + 1 : 0;
      for (u64 r = 0; r <= n_reads; ++r) kmer_offsets[r] = r * per;
    }
    auto base_off = [&](u64 r) { return uniform_read_len ? r * uniform_read_len : read_offsets[r]; };
    auto slot_off = [&](u64 r) { return uniform_read_len ? r * (uniform_read_len >= k ? uniform_read_len - k + 1 : 0) : koffs[r]; };
    // chunk boundaries: ~32 MiB of bases per chunk
    u64 TARGET = 32ull << 20;
    if (const char* e = getenv("MAZU_B200_CHUNK_MIB")) {  // tuning knob: bases per pipeline chunk
      long v = atol(e);
      if (v > 0) TARGET = (u64)v << 20;
    }
    std::vector<u64> cuts{0};
    {
      u64 r = 0;
      while (r < n_reads) {
        u64 r1;
        if (uniform_read_len) r1 = std::min(n_reads, r + std::max<u64>(1, TARGET / uniform_read_len));
        else {
          u64 lim = read_offsets[r] + TARGET;
          r1 = (u64)(std::upper_bound(read_offsets + r + 1, read_offsets + n_reads + 1, lim) - read_offsets) - 1;
          r1 = std::max(r1, r + 1);
        }
        cuts.push_back(r1);
        r = r1;
      }
    }
    u64 max_bases = 0, max_slots = 0, max_reads = 0;
    for (size_t c = 0; c + 1 < cuts.size(); ++c) {
      max_bases = std::max(max_bases, base_off(cuts[c + 1]) - base_off(cuts[c]));
      max_slots = std::max(max_slots, slot_off(cuts[c + 1]) - slot_off(cuts[c]));
      max_reads = std::max(max_reads, cuts[c + 1] - cuts[c]);
    }
    StreamPair sp;
    PoolScratch scratch(idx->pool, sp.s[0]);
    void *d_bases[2] = {nullptr, nullptr}, *d_ro[2] = {nullptr, nullptr}, *d_ko[2] = {nullptr, nullptr}, *d_hits[2] = {nullptr, nullptr};
    void* d_counts = scratch.get(3 * 8);
    MZ_CUDA(cudaMemsetAsync(d_counts, 0, 24, sp.s[0]));
    for (int b = 0; b < 2; ++b) {
      d_bases[b] = scratch.get(max_bases + 16);
      if (!uniform_read_len) {
        d_ro[b] = scratch.get((max_reads + 1) * 8);
        d_ko[b] = scratch.get((max_reads + 1) * 8);
      }
      if (out_hits && !dev_out) d_hits[b] = scratch.get(max_slots * rec + 16);
    }
    scratch.publish(sp.s[1]);
    int b = 0;
    for (size_t c = 0; c + 1 < cuts.size(); ++c, b ^= 1) {
      u64 r0 = cuts[c], r1 = cuts[c + 1];
      u64 b0 = base_off(r0), nb = base_off(r1) - b0;
      u64 s0 = slot_off(r0), ns = slot_off(r1) - s0;
      cudaStream_t s = sp.s[b];
      MZ_CUDA(cudaMemcpyAsync(d_bases[b], bases + b0, nb, cudaMemcpyHostToDevice, s));
      const u64* dro = nullptr;
      const u64* dko = nullptr;
      if (!uniform_read_len) {
        MZ_CUDA(cudaMemcpyAsync(d_ro[b], read_offsets + r0, (r1 - r0 + 1) * 8, cudaMemcpyHostToDevice, s));
        MZ_CUDA(cudaMemcpyAsync(d_ko[b], koffs + r0, (r1 - r0 + 1) * 8, cudaMemcpyHostToDevice, s));
        dro = (const u64*)d_ro[b];
        dko = (const u64*)d_ko[b];
      }
      // offsets uploaded are absolute: rebase the data pointers instead of the offset arrays
      const u8* dbases = (const u8*)d_bases[b] - (uniform_read_len ? 0 : b0);
      void* dh = nullptr;
      if (out_hits && dev_out) dh = (char*)out_hits + (uniform_read_len ? s0 * rec : 0);
      else if (out_hits) dh = (char*)d_hits[b] - (uniform_read_len ? 0 : s0 * rec);
      launch_query_reads(idx, dbases, dro, r1 - r0, uniform_read_len, mode, dko, dh, compact, (u64*)d_counts, s);
      if (out_hits && !dev_out && ns) MZ_CUDA(cudaMemcpyAsync((char*)out_hits + s0 * rec, d_hits[b], ns * rec, cudaMemcpyDeviceToHost, s));
    }
    MZ_CUDA(cudaStreamSynchronize(sp.s[1]));
    if (counts) MZ_CUDA(cudaMemcpyAsync(counts, d_counts, 24, cudaMemcpyDeviceToHost, sp.s[0]));
    MZ_CUDA(cudaStreamSynchronize(sp.s[0]));
  });
}

mazu_status_t mazu_b200_query_reads(const mazu_index_t* idx, const uint8_t* bases, const uint64_t* read_offsets, uint64_t n_reads,
                                    uint64_t uniform_read_len, int32_t mode, uint64_t* kmer_offsets, mazu_hit_t* out_hits,
                                    uint64_t* counts, int32_t mem, void* stream) {
  return query_reads_impl(idx, bases, read_offsets, n_reads, uniform_read_len, mode, kmer_offsets, out_hits, 0, counts, mem, stream);
}

mazu_status_t mazu_b200_query_reads_compact(const mazu_index_t* idx, const uint8_t* bases, const uint64_t* read_offsets, uint64_t n_reads,
                                            uint64_t uniform_read_len, int32_t mode, uint64_t* kmer_offsets, mazu_hit8_t* out_hits,
                                            uint64_t* counts, int32_t mem, void* stream) {
  if (idx) {
    // pos shares a word with the match type: every unitig must be shorter than 2^30 bases
    const auto& acc = idx->unitigs->accum;
    static thread_local const mazu_index* checked = nullptr;
    if (checked != idx) {
      for (size_t i = 0; i + 1 < acc.size(); ++i)
        if (acc[i + 1] - acc[i] >= (1ULL << 30)) {
          g_err = "compact hit records need every unitig shorter than 2^30 bases";
          return MAZU_ERR_INVALID_ARG;
        }
      checked = idx;
    }
  }
  return query_reads_impl(idx, bases, read_offsets, n_reads, uniform_read_len, mode, kmer_offsets, out_hits, 1, counts, mem, stream);
}

mazu_status_t mazu_b200_encode_reads(const mazu_index_t* idx, const uint8_t* bases, const uint64_t* read_offsets, uint64_t n_reads,
                                     uint64_t uniform_read_len, const uint64_t* kmer_offsets, uint64_t* out_fw, uint64_t* out_rc,
                                     uint64_t* out_mm_word, uint32_t* out_mm_offset, uint8_t* out_valid, void* stream) {
  return guarded([&] {
    if (!idx || (n_reads && !bases)) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    if (n_reads && !uniform_read_len && (!read_offsets || !kmer_offsets)) throw Error(MAZU_ERR_INVALID_ARG, "ragged reads need read_offsets and kmer_offsets");
    if (!n_reads) return;
    DeviceGuard g(idx->device);
    int grid = grid_for(encode_reads_kernel, QR_WARPS * 32, idx, QR_WARPS, n_reads);
    encode_reads_kernel<<<grid, QR_WARPS * 32, 0, (cudaStream_t)stream>>>(idx->view, bases, read_offsets, n_reads, uniform_read_len, kmer_offsets,
                                                                          out_fw, out_rc, out_mm_word, out_mm_offset, out_valid);
    MZ_CUDA(cudaGetLastError());
  });
}

// shared driver of decode_occs / project_hits
static void occ_driver(const mazu_index_t* idx, const uint32_t* uids, const mazu_hit_t* hits, uint64_t n, uint64_t* out_offsets,
                       mazu_occ_t* out, uint64_t cap, uint64_t* out_total, int32_t mem, void* stream) {
  if (!idx || !out_offsets) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
  if (idx->view.u2pos_kind == MAZU_U2POS_NONE) throw Error(MAZU_ERR_NO_U2POS, "index has no U2Pos table");
  if (mem != MAZU_MEM_HOST && mem != MAZU_MEM_DEVICE) throw Error(MAZU_ERR_INVALID_ARG, "unknown mem mode");
  DeviceGuard g(idx->device);
  const bool project = hits != nullptr;
  cudaStream_t s = (cudaStream_t)stream;
  std::unique_ptr<StreamPair> sp;
  std::unique_ptr<PoolScratch> scratch;  // declared after `sp`: released before the streams go away
  void *d_in = nullptr, *d_offs = nullptr, *d_out = nullptr;
  const u32* d_uids = uids;
  const Hit* d_hits = (const Hit*)hits;
  u64* d_offsets = out_offsets;
  if (mem == MAZU_MEM_HOST) {
    sp = std::make_unique<StreamPair>();
    s = sp->s[0];
    scratch = std::make_unique<PoolScratch>(idx->pool, s);
    d_in = scratch->get(n * (project ? 16 : 4) + 16);
    MZ_CUDA(cudaMemcpyAsync(d_in, project ? (const void*)hits : (const void*)uids, n * (project ? 16 : 4), cudaMemcpyHostToDevice, s));
    d_uids = project ? nullptr : (const u32*)d_in;
    d_hits = project ? (const Hit*)d_in : nullptr;
    d_offs = scratch->get((n + 1) * 8);
    d_offsets = (u64*)d_offs;
  }
  void* lens = nullptr;
  MZ_CUDA(cudaMallocFromPoolAsync(&lens, (n + 1) * 8, idx->pool, s));
  MZ_CUDA(cudaMemsetAsync(lens, 0, (n + 1) * 8, s));
  if (n) {
    occ_lens_kernel<<<(int)std::min<u64>((n + 255) / 256, (u64)idx->sm_count * 8), 256, 0, s>>>(idx->view, d_uids, d_hits, n, (u64*)lens);
    MZ_CUDA(cudaGetLastError());
  }
  device_exclusive_scan((const u64*)lens, d_offsets, n, idx->pool, s);
  MZ_CUDA(cudaFreeAsync(lens, s));
  u64 total = 0;
  bool need_total = mem == MAZU_MEM_HOST || out_total != nullptr;
  if (need_total) {
    MZ_CUDA(cudaMemcpyAsync(&total, d_offsets + n, 8, cudaMemcpyDeviceToHost, s));
    MZ_CUDA(cudaStreamSynchronize(s));
    if (out_total) *out_total = total;
  }
  if (mem == MAZU_MEM_HOST) MZ_CUDA(cudaMemcpyAsync(out_offsets, d_offsets, (n + 1) * 8, cudaMemcpyDeviceToHost, s));
  if (!out) {
    if (mem == MAZU_MEM_HOST) MZ_CUDA(cudaStreamSynchronize(s));
    return;
  }
  if (need_total && total > cap) {
    if (mem == MAZU_MEM_HOST) MZ_CUDA(cudaStreamSynchronize(s));
    throw Error(MAZU_ERR_INVALID_ARG, "output capacity too small: need " + std::to_string(total) + " records");
  }
  OccRec* d_o = (OccRec*)out;
  if (mem == MAZU_MEM_HOST) {
    d_out = scratch->get(total * 12 + 16);
    d_o = (OccRec*)d_out;
  }
  if (n) {
    int grid = idx->sm_count * 8;  // tiles of the OUTPUT are grid-strided; the kernel reads the total from out_offsets[n]
    if (project) occ_fill_kernel<true><<<grid, 256, 0, s>>>(idx->view, d_uids, d_hits, n, d_offsets, d_o);
    else occ_fill_kernel<false><<<grid, 256, 0, s>>>(idx->view, d_uids, d_hits, n, d_offsets, d_o);
    MZ_CUDA(cudaGetLastError());
  }
  if (mem == MAZU_MEM_HOST) {
    if (total) MZ_CUDA(cudaMemcpyAsync(out, d_o, total * 12, cudaMemcpyDeviceToHost, s));
    MZ_CUDA(cudaStreamSynchronize(s));
  }
}

mazu_status_t mazu_b200_decode_occs(const mazu_index_t* idx, const uint32_t* unitig_ids, uint64_t n, uint64_t* out_offsets,
                                    mazu_occ_t* out_occs, uint64_t cap, uint64_t* out_total, int32_t mem, void* stream) {
  return guarded([&] {
    if (n && !unitig_ids) throw Error(MAZU_ERR_INVALID_ARG, "null unitig_ids");
    occ_driver(idx, unitig_ids, nullptr, n, out_offsets, out_occs, cap, out_total, mem, stream);
  });
}

mazu_status_t mazu_b200_project_hits(const mazu_index_t* idx, const mazu_hit_t* hits, uint64_t n, uint64_t* out_offsets,
                                     mazu_occ_t* out_mrps, uint64_t cap, uint64_t* out_total, int32_t mem, void* stream) {
  return guarded([&] {
    if (!hits) throw Error(MAZU_ERR_INVALID_ARG, "null hits");
    occ_driver(idx, nullptr, hits, n, out_offsets, out_mrps, cap, out_total, mem, stream);
  });
}

mazu_status_t mazu_b200_iter_unitigs_on_ref(const mazu_index_t* idx, uint64_t ref_id, mazu_hit_t* out, uint64_t cap, uint64_t* n_out) {
  return guarded([&] {
    if (!idx || !n_out) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    if (!idx->refs || !idx->refs->has_seq) throw Error(MAZU_ERR_NO_REFSEQ, "Refseq is None");
    if (ref_id >= idx->refs->n_refs()) throw Error(MAZU_ERR_INVALID_ARG, "reference id out of range");
    const u64 k = idx->unitigs->k;
    const u64 begin = idx->refs->prefix[ref_id], len = idx->refs->prefix[ref_id + 1] - begin;
    *n_out = 0;
    if (len < k) return;
    const u64 n_pos = len - k + 1;  // index.rs:368
    DeviceGuard g(idx->device);
    DevBuf d(n_pos * 16, idx->device);
    int grid = (int)std::max<u64>(1, std::min<u64>((n_pos + 255) / 256, (u64)idx->sm_count * 8));
    ref_hits_kernel<<<grid, 256>>>(idx->view, begin, n_pos, (Hit*)d.p);
    MZ_CUDA(cudaGetLastError());
    std::vector<Hit> hits(n_pos);
    MZ_CUDA(cudaMemcpy(hits.data(), d.p, n_pos * 16, cudaMemcpyDeviceToHost));
    u64 n = 0;
    for (u64 pos = 0; pos < n_pos;) {
      const Hit& h = hits[pos];
      if (h.match == NO_MATCH) throw Error(MAZU_ERR_INVALID_DATA, "iter_unitigs_on_ref: reference k-mer at position " + std::to_string(pos) + " is not in the index");
      if (out && n < cap) out[n] = mazu_hit_t{h.unitig_id, h.unitig_len, (uint32_t)pos, h.match == IDENTITY_MATCH ? 1u : 0u};
      ++n;
      pos += (u64)h.unitig_len - k + 1;
    }
    *n_out = n;
    if (out && n > cap) throw Error(MAZU_ERR_INVALID_ARG, "output capacity too small: need " + std::to_string(n) + " records");
  });
}

static void run_validate(const mazu_index_t* idx, bool k2u_only, uint64_t counts[5]) {
  if (!idx || !counts) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
  DeviceGuard g(idx->device);
  DevBuf d(5 * 8, idx->device);
  MZ_CUDA(cudaMemset(d.p, 0, 40));
  if (k2u_only) {
    u64 n = idx->unitigs->total_len();
    int grid = (int)std::max<u64>(1, std::min<u64>((n + 255) / 256, (u64)idx->sm_count * 8));
    k2u_validate_self_kernel<<<grid, 256>>>(idx->view, (unsigned long long*)d.p);
  } else {
    if (!idx->refs || !idx->refs->has_seq) throw Error(MAZU_ERR_NO_REFSEQ, "validate_self: index has no reference sequence (assert has_refseq)");
    if (idx->view.u2pos_kind == MAZU_U2POS_NONE) throw Error(MAZU_ERR_NO_U2POS, "validate_self: index has no U2Pos table");
    u64 n = idx->refs->prefix.back();
    int grid = (int)std::max<u64>(1, std::min<u64>((n + 255) / 256, (u64)idx->sm_count * 8));
    validate_self_kernel<<<grid, 256>>>(idx->view, (unsigned long long*)d.p);
  }
  MZ_CUDA(cudaGetLastError());
  MZ_CUDA(cudaMemcpy(counts, d.p, 40, cudaMemcpyDeviceToHost));
}

mazu_status_t mazu_b200_validate_self(const mazu_index_t* idx, uint64_t counts[5]) {
  return guarded([&] { run_validate(idx, false, counts); });
}
mazu_status_t mazu_b200_k2u_validate_self(const mazu_index_t* idx, uint64_t counts[5]) {
  return guarded([&] { run_validate(idx, true, counts); });
}

mazu_status_t mazu_b200_measure_random_gather(uint64_t table_bytes, uint64_t n_gathers, int32_t iters, int32_t device, double* sectors_per_s) {
  return guarded([&] {
    if (!sectors_per_s || table_bytes < 64) throw Error(MAZU_ERR_INVALID_ARG, "bad argument");
    DeviceGuard g(device);
    cudaDeviceProp prop;
    MZ_CUDA(cudaGetDeviceProperties(&prop, device));
    DevBuf table(table_bytes, device), sink(8, device);
    MZ_CUDA(cudaMemset(table.p, 1, table_bytes));
    MZ_CUDA(cudaMemset(sink.p, 0, 8));
    cudaEvent_t e0, e1;
    MZ_CUDA(cudaEventCreate(&e0));
    MZ_CUDA(cudaEventCreate(&e1));
    double best = 0;
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, random_gather_kernel, 256, 0);
    int grid = prop.multiProcessorCount * std::max(occ, 1);
    for (int it = 0; it < iters + 1; ++it) {
      MZ_CUDA(cudaEventRecord(e0));
      random_gather_kernel<<<grid, 256>>>((const uint4*)table.p, table_bytes / 32, n_gathers, 0x1234 + it, (unsigned long long*)sink.p);
      MZ_CUDA(cudaEventRecord(e1));
      MZ_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      MZ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (it > 0) best = std::max(best, (double)n_gathers / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *sectors_per_s = best;
  });
}

}  // extern "C"
