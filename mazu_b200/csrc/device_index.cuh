// Device-resident index: the opaque handle behind mazu_index_t, device buffers, and the one-time upload of the
// host-built tables (UnitigSet directory + starts, MPHF blocks, bucket bounds, packed positions, U2Pos, references).
#pragma once
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
  // function attributes are per device: the staged decode kernel's dynamic shared memory limit is raised once per handle
  mutable std::once_flag occ_attr_once;
  bool compact_ok = true;  // every unitig shorter than 2^30 bases: mazu_hit8_t can hold pos | match << 30
  // canonical k-mers pairwise distinct (flag_duplicated_kmers_kernel at creation): streaming queries then equal random-access
  // queries record for record and are served by the random-access kernel
  bool kmers_unique = false;
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
  for (u64 i = 0; i < U; ++i)  // hit records and unitig lines carry 32-bit lengths
    if (us.accum[i + 1] - us.accum[i] >= (1ULL << 31) - 512) throw Error(MAZU_ERR_INVALID_ARG, "a unitig is longer than 2^31 - 512 bases");
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
  d->view.useq = (const u64*)b_seq->p;
  d->view.dir = (const u32*)b_dir->p;
  d->view.starts = (const u64*)b_starts->p;
  d->view.total_len = L;
  d->view.n_unitigs = U;
  d->view.k = us.k;
  d->view.dir_shift = DIR_SHIFT;
  // unitig lines: what the query kernels read (window + id + bounds of a candidate in one 128-byte line)
  const u64 n_lines = (L >> ULINE_SHIFT) + 2;
  auto b_lines = std::make_shared<DevBuf>(n_lines * sizeof(UnitigLine), dev);
  d->view.lines = (const UnitigLine*)b_lines->p;
  build_unitig_lines_kernel<<<(int)std::min<u64>((n_lines + 255) / 256, 148 * 8), 256>>>(d->view, n_lines, std::min<u64>(nw, us.useq.size()) + 4,
                                                                                        (UnitigLine*)b_lines->p);
  MZ_CUDA(cudaGetLastError());
  MZ_CUDA(cudaDeviceSynchronize());
  d->bufs = {b_seq, b_dir, b_starts, b_lines};
  d->bytes = b_seq->bytes + b_dir->bytes + b_starts->bytes + b_lines->bytes;
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
