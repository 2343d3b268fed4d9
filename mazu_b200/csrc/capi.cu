// extern "C" surface of libmazu_b200.so (include/mazu_b200.h): index upload, kernel launches, and the
// chunked host<->device pipelines behind MAZU_MEM_HOST calls.  There is no CPU fallback: every query
// entry point launches the CUDA kernels of kernels.cuh or fails with MAZU_ERR_CUDA.
#include <array>

#include "device_index.cuh"
#include "gpu_build_driver.cuh"
#include "../../include/mazu_b200_debug.h"

namespace {

// every unitig k-mer looked up once: found anywhere but at its own position <=> the set holds a canonical k-mer twice; the unitig
// lines of both positions are flagged (the walk kernel settles cursors only around flagged lines)
void detect_unique_kmers(mazu_index& ix) {
  static const bool force_walk = getenv("MAZU_B200_FORCE_WALK") != nullptr;  // measurement knob: always take the cursor-walk kernel
  ix.kmers_unique = false;
  if (ix.unitigs->total_len() < ix.unitigs->k) return;
  DevBuf d(8, ix.device);
  MZ_CUDA(cudaMemset(d.p, 0, 8));
  const u64 n = ix.unitigs->total_len();
  const int grid = (int)std::max<u64>(1, std::min<u64>((n + 255) / 256, (u64)ix.sm_count * 8));
  // the lines are written once more here (flags), before the handle is handed out; they are read-only from then on
  flag_duplicated_kmers_kernel<<<grid, 256>>>(ix.view, const_cast<UnitigLine*>(ix.view.unitigs.lines), (unsigned long long*)d.p);
  MZ_CUDA(cudaGetLastError());
  u64 bad = 0;
  MZ_CUDA(cudaMemcpy(&bad, d.p, 8, cudaMemcpyDeviceToHost));
  ix.kmers_unique = bad == 0 && !force_walk;
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
  for (size_t i = 0; i + 1 < ix->unitigs->accum.size(); ++i)
    if (ix->unitigs->accum[i + 1] - ix->unitigs->accum[i] >= (1ULL << 30)) ix->compact_ok = false;
  ix->view.unitigs = ix->d_unitigs->view;
  ix->tables[7] = {ix->view.unitigs.lines, ((ix->unitigs->total_len() >> ULINE_SHIFT) + 2) * sizeof(UnitigLine)};
  if (ix->gpu_builder) ix->gpu_builder(*ix);  // tables are produced in HBM (gpu_build.cuh)
  else upload_k2u(*ix);
  upload_u2pos(*ix);
  upload_refs(*ix);
  MZ_CUDA(cudaDeviceSynchronize());
  detect_unique_kmers(*ix);
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
  if (ix->view.k2u_kind == MAZU_K2U_PFHASH) {  // several keys per lane through one MPHF level loop
    const bool native = ix->view.mphf.family == MPHF_FAMILY_NATIVE;
    auto kern = native ? k2u_batch_pfhash_kernel<MPHF_FAMILY_NATIVE> : k2u_batch_pfhash_kernel<MPHF_FAMILY_BOOPHF>;
    int grid = grid_for(kern, 256, ix, (u64)KB_N * 256, n);
    kern<<<grid, 256, 0, s>>>(ix->view, d_words, n, d_out);
    MZ_CUDA(cudaGetLastError());
    return;
  }
  int grid = grid_for(k2u_batch_kernel, 256, ix, 256, n);
  k2u_batch_kernel<<<grid, 256, 0, s>>>(ix->view, d_words, n, d_out);
  MZ_CUDA(cudaGetLastError());
}

void device_exclusive_scan(const u64* d_in, u64* d_out, u64 n, cudaMemPool_t pool, cudaStream_t s);

// one stream-ordered allocation from the handle's pool, given back on the same stream when the scope ends (also when it
// ends by an exception: the pool's release threshold is "keep everything", so a leak would never be returned)
struct PoolBuf {
  void* p = nullptr;
  cudaStream_t s;
  PoolBuf(cudaMemPool_t pool, size_t bytes, cudaStream_t st) : s(st) { MZ_CUDA(cudaMallocFromPoolAsync(&p, bytes ? bytes : 1, pool, st)); }
  ~PoolBuf() {
    if (p) cudaFreeAsync(p, s);
  }
  PoolBuf(const PoolBuf&) = delete;
  PoolBuf& operator=(const PoolBuf&) = delete;
};

// resident CTAs per SM the random-access read kernel is compiled for (4 -> <= 64 registers; 3 -> <= 80); A/B knob
#ifndef MAZU_QR_RANDOM_OCC
#define MAZU_QR_RANDOM_OCC 4
#endif

// (k, w) pairs the SSHash read kernels are ALSO compiled for with k and w folded into the code (kernels.cuh, KW): the pairs of
// the named configurations (k31/w19: human-scale index, `index build --skew 64`; k31/w15: the yeast SSHash fixtures).  Any other
// pair, and every PFHash index, runs the instantiation that reads k and w from the view.  MAZU_B200_GENERIC_KW=1 forces that
// instantiation (A/B; test_read_kernels_kw_specialisations runs both).
static u32 kw_code(const mazu_index* ix) {
  const char* e = getenv("MAZU_B200_GENERIC_KW");  // read per call: the parity test flips it inside one process
  const bool generic = e && *e && *e != '0';
  if (generic || ix->view.k2u_kind != MAZU_K2U_SSHASH || ix->view.unitigs.k != 31) return 0;
  return ix->view.w == 19 ? MZ_KW(31, 19) : ix->view.w == 15 ? MZ_KW(31, 15) : 0;
}
// MZ_KW_DISPATCH(code, CALL): CALL(KW) with KW a constant expression
#define MZ_KW_DISPATCH(code, CALL)            \
  switch (code) {                             \
    case MZ_KW(31, 19): CALL(MZ_KW(31, 19)); break; \
    case MZ_KW(31, 15): CALL(MZ_KW(31, 15)); break; \
    default: CALL(0); break;                  \
  }

template <int MODE, int KIND, u32 FAMILY, int OCC, u32 KW>
void launch_qr_occ(const mazu_index* ix, const u8* d_bases, const u64* d_read_offsets, u64 n_reads, u64 uniform_len, const u64* d_kmer_offsets,
                   void* d_out, u32 compact, u64* d_counts, const u64* d_seg_offsets, cudaStream_t s) {
  auto kern = query_reads_kernel<MODE, KIND, FAMILY, OCC, KW>;
  // with a segment table the number of work items is only known on the device: launch every resident CTA, idle warps leave at once
  int grid = grid_for(kern, QR_WARPS * 32, ix, QR_WARPS, d_seg_offsets ? ~0ULL >> 8 : n_reads);
  kern<<<grid, QR_WARPS * 32, 0, s>>>(ix->view, d_bases, d_read_offsets, n_reads, uniform_len, d_kmer_offsets, d_out, compact,
                                      (unsigned long long*)d_counts, d_seg_offsets);
}
template <int MODE, int KIND, u32 FAMILY, u32 KW = 0>
void launch_qr(const mazu_index* ix, const u8* d_bases, const u64* d_read_offsets, u64 n_reads, u64 uniform_len, const u64* d_kmer_offsets,
               void* d_out, u32 compact, u64* d_counts, cudaStream_t s) {
  // streaming walk: 3 resident CTAs (80 registers) while the index sits in the 126 MB L2, 4 (64 registers) once it does not
  if constexpr (MODE == 1) {
    if (ix->device_bytes() <= (96ull << 20))
      launch_qr_occ<MODE, KIND, FAMILY, 3, KW>(ix, d_bases, d_read_offsets, n_reads, uniform_len, d_kmer_offsets, d_out, compact, d_counts, nullptr, s);
    else
      launch_qr_occ<MODE, KIND, FAMILY, 4, KW>(ix, d_bases, d_read_offsets, n_reads, uniform_len, d_kmer_offsets, d_out, compact, d_counts, nullptr, s);
  } else {
    // ragged batches: per-read work-item counts -> scan; the kernel cuts long reads into segments (kernels.cuh, QR_SEGMENT)
    std::unique_ptr<PoolBuf> cnt, seg;
    if (!uniform_len) {
      cnt = std::make_unique<PoolBuf>(ix->pool, (n_reads + 1) * 8, s);
      seg = std::make_unique<PoolBuf>(ix->pool, (n_reads + 1) * 8, s);
      MZ_CUDA(cudaMemsetAsync(cnt->p, 0, (n_reads + 1) * 8, s));
      segment_counts_kernel<<<(int)std::min<u64>((n_reads + 255) / 256, 4096), 256, 0, s>>>(d_read_offsets, n_reads, ix->unitigs->k, QR_SEGMENT, (u64*)cnt->p);
      MZ_CUDA(cudaGetLastError());
      device_exclusive_scan((const u64*)cnt->p, (u64*)seg->p, n_reads, ix->pool, s);
    }
    launch_qr_occ<MODE, KIND, FAMILY, MAZU_QR_RANDOM_OCC, KW>(ix, d_bases, d_read_offsets, n_reads, uniform_len, d_kmer_offsets, d_out, compact, d_counts,
                                                          seg ? (const u64*)seg->p : nullptr, s);
  }
}

void launch_query_reads(const mazu_index* ix, const u8* d_bases, const u64* d_read_offsets, u64 n_reads, u64 uniform_len, int mode,
                        const u64* d_kmer_offsets, void* d_out, u32 compact, u64* d_counts, cudaStream_t s) {
  if (n_reads == 0) return;
  const bool ss = ix->view.k2u_kind == MAZU_K2U_SSHASH;
  const bool native = ix->view.mphf.family != MPHF_FAMILY_BOOPHF;  // SSHash: fingerprinted cascade for the minimizers, native MPHF for the skew index
  // distinct canonical k-mers: the cursor walk cannot answer anything the lookup does not (kernels.cuh, flag_duplicated_kmers_kernel)
  const bool st = mode == MAZU_MODE_STREAMING && !ix->kmers_unique;
#define MZ_QR(M, K, F) launch_qr<M, K, F>(ix, d_bases, d_read_offsets, n_reads, uniform_len, d_kmer_offsets, d_out, compact, d_counts, s)
  if (ss) {
    if (!native) throw Error(MAZU_ERR_OTHER, "internal: SSHash index without a native MPHF");
#define MZ_QR_SS(KW)                                                                                                               \
  if (st) launch_qr<1, MAZU_K2U_SSHASH, MPHF_FAMILY_NATIVE, KW>(ix, d_bases, d_read_offsets, n_reads, uniform_len, d_kmer_offsets, d_out, compact, d_counts, s); \
  else launch_qr<0, MAZU_K2U_SSHASH, MPHF_FAMILY_NATIVE, KW>(ix, d_bases, d_read_offsets, n_reads, uniform_len, d_kmer_offsets, d_out, compact, d_counts, s)
    MZ_KW_DISPATCH(kw_code(ix), MZ_QR_SS)
#undef MZ_QR_SS
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
  PoolBuf tmp(pool, tmp_bytes, s);
  MZ_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, d_in, d_out, (size_t)(n + 1), s));
}

// Tiles of get_ref_pos over reads: one per chunk of QR_CHUNK k-mer positions, at least one per read.  Uniform reads need no
// table; for ragged reads seg (n_reads + 1 first-tile indexes) is built on the device and the tile count read back (one
// synchronisation per call).
struct TileTable {
  std::unique_ptr<PoolBuf> cnt, seg;
  u64 n_tiles = 0;
};
void make_tiles(const mazu_index* idx, const u64* d_read_offsets, u64 n_reads, u64 uniform_len, cudaStream_t s, TileTable& t) {
  const u32 k = idx->unitigs->k;
  if (uniform_len) {
    const u64 nk = uniform_len >= k ? uniform_len - k + 1 : 0;
    t.n_tiles = n_reads * (nk <= (u64)QR_CHUNK ? 1 : (nk + QR_CHUNK - 1) / QR_CHUNK);
    return;
  }
  t.cnt = std::make_unique<PoolBuf>(idx->pool, (n_reads + 1) * 8, s);
  t.seg = std::make_unique<PoolBuf>(idx->pool, (n_reads + 1) * 8, s);
  MZ_CUDA(cudaMemsetAsync(t.cnt->p, 0, (n_reads + 1) * 8, s));
  segment_counts_kernel<<<(int)std::min<u64>((n_reads + 255) / 256, 4096), 256, 0, s>>>(d_read_offsets, n_reads, k, (u64)QR_CHUNK, (u64*)t.cnt->p);
  MZ_CUDA(cudaGetLastError());
  device_exclusive_scan((const u64*)t.cnt->p, (u64*)t.seg->p, n_reads, idx->pool, s);
  MZ_CUDA(cudaMemcpyAsync(&t.n_tiles, (u64*)t.seg->p + n_reads, 8, cudaMemcpyDeviceToHost, s));
  MZ_CUDA(cudaStreamSynchronize(s));
}

// fused reads -> hit runs (query_reads_runs_kernel): codes per slot, chunk-local run records, per-read run offsets; d_rro[n_reads]
// is the run cursor (the chunk's run count afterwards).  Only for chunks whose reads all fit one tile.
void launch_query_reads_runs(const mazu_index* idx, const u8* d_bases, const u64* d_read_offsets, u64 n_reads, u64 uniform_len, const u64* d_kmer_offsets,
                             u64* d_counts, u8* d_codes, Hit* d_runs, u64* d_rro, u64 cap, cudaStream_t s, const u64* d_packed_words = nullptr,
                             const u64* d_packed_nmask = nullptr, u8* d_codes2 = nullptr, uint4* d_intervals = nullptr, u64 read_base = 0,
                             u64* d_cursor = nullptr) {
  if (!d_cursor) d_cursor = d_rro + n_reads;  // run format: the entry behind the last read's offset doubles as the cursor
  MZ_CUDA(cudaMemsetAsync(d_cursor, 0, 8, s));
  if (!n_reads) return;
  RunsTileOut ro{d_codes, d_codes2, d_runs, d_rro, (unsigned long long*)d_cursor, cap, d_intervals, read_base};
  const bool ss = idx->view.k2u_kind == MAZU_K2U_SSHASH, boophf = idx->view.mphf.family == MPHF_FAMILY_BOOPHF;
  const int io = d_intervals ? (d_packed_words ? RUNS_IO_INTERVALS : RUNS_IO_INTERVALS_ASCII)
                             : (d_packed_words && d_codes2 ? RUNS_IO_PACKED : RUNS_IO_ASCII);
  if (io == RUNS_IO_ASCII && (d_packed_words || d_codes2)) throw Error(MAZU_ERR_INVALID_ARG, "packed reads and 2-bit codes come together");
#define MZ_QRR(K, F, KW)                                                                                                      \
  {                                                                                                                            \
    auto kern = io == RUNS_IO_INTERVALS         ? query_reads_runs_kernel<K, F, RUNS_IO_INTERVALS, KW>                         \
                : io == RUNS_IO_INTERVALS_ASCII ? query_reads_runs_kernel<K, F, RUNS_IO_INTERVALS_ASCII, KW>                   \
                : io == RUNS_IO_PACKED          ? query_reads_runs_kernel<K, F, RUNS_IO_PACKED, KW>                            \
                                                : query_reads_runs_kernel<K, F, RUNS_IO_ASCII, KW>;                            \
    int grid = grid_for(kern, QR_WARPS * 32, idx, QR_WARPS, n_reads);                                                          \
    kern<<<grid, QR_WARPS * 32, 0, s>>>(idx->view, d_bases, d_read_offsets, n_reads, uniform_len, d_kmer_offsets,              \
                                        (unsigned long long*)d_counts, ro, d_packed_words, d_packed_nmask);                    \
  }
#define MZ_QRR_SS(KW) MZ_QRR(MAZU_K2U_SSHASH, MPHF_FAMILY_NATIVE, KW)
  if (ss) {
    MZ_KW_DISPATCH(kw_code(idx), MZ_QRR_SS)
  } else if (idx->view.k2u_kind == MAZU_K2U_SAMPLED_PFHASH) MZ_QRR(MAZU_K2U_SAMPLED_PFHASH, MPHF_FAMILY_BOOPHF, 0)
  else if (boophf) MZ_QRR(MAZU_K2U_PFHASH, MPHF_FAMILY_BOOPHF, 0)
  else MZ_QRR(MAZU_K2U_PFHASH, MPHF_FAMILY_NATIVE, 0)
#undef MZ_QRR_SS
#undef MZ_QRR
  MZ_CUDA(cudaGetLastError());
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
  std::vector<cudaStream_t> users;  // other streams the buffers were published to
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
    users.push_back(other);
  }
  ~PoolScratch() {
    // the frees are ordered on `s` only: when the scope ends early (an exception in a multi-stream pipeline) work on the
    // other streams may still be using the buffers, so wait for them first (a no-op on the normal path, which has synchronised)
    for (cudaStream_t u : users) cudaStreamSynchronize(u);
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

mazu_status_t mazu_b200_debug_probe_key(const mazu_index_t* idx, const uint64_t* fw_words, uint64_t n, uint32_t* out_block, void* stream) {
  return guarded([&] {
    if (!idx || (n && (!fw_words || !out_block))) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    if (!n) return;
    DeviceGuard g(idx->device);
    probe_key_kernel<<<idx->sm_count * 8, 256, 0, (cudaStream_t)stream>>>(idx->view, fw_words, n, out_block);
    MZ_CUDA(cudaGetLastError());
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
    case MAZU_INFO_N_MINIMIZERS: return idx->k2u->kind == MAZU_K2U_SSHASH ? idx->k2u->n_minimizers + 1 : 0;  // len of the reference's prefix sum (sshash.rs:333-335)
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
    case MAZU_INFO_KMERS_UNIQUE: return idx->kmers_unique ? 1 : 0;
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

// hit-run output of mazu_b200_query_reads_runs (host buffers)
struct RunsOut {
  uint8_t* codes;
  mazu_hit_t* runs;            // the caller's whole run array
  uint64_t cap_runs;           // end (exclusive) of the part of `runs` this call may fill
  uint64_t* read_run_offsets;  // n_reads + 1 (the last entry only if write_end)
  uint64_t n_runs = 0;         // runs this call produced (it keeps counting past the capacity)
  uint64_t base = 0;           // index in `runs` of this call's first run (sharded calls: the shard's region)
  bool write_end = true;       // write read_run_offsets[n_reads] (sharded calls: that entry belongs to the next shard)
  // packed I/O (mazu_b200_query_reads_runs_packed): uniform reads arrive as 2-bit words (+ optional N mask), codes leave 2 bits per slot
  const uint64_t* packed_words = nullptr;
  const uint64_t* packed_nmask = nullptr;
  bool codes2 = false;
};

static mazu_status_t query_reads_impl(const mazu_index_t* idx, const uint8_t* bases, const uint64_t* read_offsets, uint64_t n_reads,
                                      uint64_t uniform_read_len, int32_t mode, uint64_t* kmer_offsets, void* out_hits, uint32_t compact,
                                      uint64_t* counts, int32_t mem, void* stream, RunsOut* ro = nullptr) {
  const u64 rec = compact ? 8 : 16;
  return guarded([&] {
    if (!idx) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    if (ro && (mem != MAZU_MEM_HOST || compact || out_hits)) throw Error(MAZU_ERR_INVALID_ARG, "hit runs are a host-buffer output");
    if (ro && n_reads && (!ro->codes || !ro->read_run_offsets || (ro->cap_runs && !ro->runs))) throw Error(MAZU_ERR_INVALID_ARG, "null run buffers");
    if (mode != MAZU_MODE_RANDOM && mode != MAZU_MODE_STREAMING) throw Error(MAZU_ERR_INVALID_ARG, "unknown query mode");
    const bool packed_in = ro && ro->packed_words;
    if (n_reads && !bases && !packed_in) throw Error(MAZU_ERR_INVALID_ARG, "null bases");
    if (n_reads && !uniform_read_len && !read_offsets) throw Error(MAZU_ERR_INVALID_ARG, "read_offsets is required for ragged reads");
    if (ro && (packed_in || ro->codes2) && !uniform_read_len) throw Error(MAZU_ERR_INVALID_ARG, "packed reads / packed codes need uniform reads");
    const u32 k = idx->unitigs->k;
    DeviceGuard g(idx->device);
    if (mem == MAZU_MEM_DEVICE) {
      cudaStream_t s = (cudaStream_t)stream;
      u64* d_koffs = kmer_offsets;
      std::unique_ptr<PoolBuf> tmp_koffs;
      if (!uniform_read_len && n_reads) {
        if (!d_koffs) {
          tmp_koffs = std::make_unique<PoolBuf>(idx->pool, (n_reads + 1) * 8, s);
          d_koffs = (u64*)tmp_koffs->p;
        }
        PoolBuf lens(idx->pool, (n_reads + 1) * 8, s);
        MZ_CUDA(cudaMemsetAsync(lens.p, 0, (n_reads + 1) * 8, s));
        kmer_counts_kernel<<<(int)std::min<u64>((n_reads + 255) / 256, 4096), 256, 0, s>>>(read_offsets, n_reads, k, (u64*)lens.p);
        MZ_CUDA(cudaGetLastError());
        device_exclusive_scan((const u64*)lens.p, d_koffs, n_reads, idx->pool, s);
      }
      launch_query_reads(idx, bases, read_offsets, n_reads, uniform_read_len, mode, d_koffs, out_hits, compact, counts, s);
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
    // chunk boundaries: ~16 MiB of bases per chunk (8 / 16 / 32 / 64 measured on the final kernels: profiles/experiments/README.md)
    u64 TARGET = 16ull << 20;
    if (const char* e = getenv("MAZU_B200_CHUNK_MIB")) {  // tuning knob: bases per pipeline chunk
      long v = atol(e);
      if (v > 0) TARGET = (u64)v << 20;
    }
    std::vector<u64> cuts{0};
    {
      u64 r = 0;
      while (r < n_reads) {
        u64 r1;
        if (uniform_read_len) r1 = std::min(n_reads, r + std::max<u64>(4, (TARGET / uniform_read_len) & ~3ULL));  // a multiple of 4 reads: packed codes stay byte aligned
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
    // run buffers the device can address (pinned host memory: torch pin_memory, mazu_b200_alloc_pinned, cudaHostAlloc) take the
    // sync-free path: run records are published into them at a device-side running base; otherwise the host waits for each
    // chunk's total to place its runs (below)
    Hit* runs_dev = nullptr;
    if (ro && ro->cap_runs > ro->base) {
      cudaPointerAttributes at{};
      if (cudaPointerGetAttributes(&at, ro->runs) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) runs_dev = (Hit*)at.devicePointer;
      else cudaGetLastError();
    }
    // hit runs come straight out of the fused kernel unless the cursor walk is needed (streaming on a set with duplicated k-mers)
    static const bool unfused_knob = getenv("MAZU_B200_UNFUSED_RUNS") != nullptr;  // A/B knob: the round-1 chain of four kernels
    bool fused_runs = ro && !(mode == MAZU_MODE_STREAMING && !idx->kmers_unique) && !unfused_knob;
    if (fused_runs) {  // the fused kernel owns one read per warp and one chunk per read: every read must fit QR_CHUNK k-mer positions
      const u64 max_len = (u64)QR_CHUNK + k - 1;
      if (uniform_read_len) fused_runs = uniform_read_len <= max_len;
      else
        for (u64 r = 0; r < n_reads && fused_runs; ++r) fused_runs = read_offsets[r + 1] - read_offsets[r] <= max_len;
    }
    if (fused_runs && packed_in != ro->codes2) fused_runs = false;  // the fused kernel is built for ASCII in / byte codes out and packed in / 2-bit codes out
    // two buffers / streams alternate; the sync-free run path uses three, because its per-chunk chain (H2D, lookups, encode, two
    // D2H copies) is 2.3x as long as its compute and two streams leave the SMs idle a third of the time
    static const int MAXB = 6;
    static const int nb_knob = [] {  // buffers / streams of the sync-free run pipeline.  Measured on config 5 (ASCII reads in, byte codes
      // out): 3 -> 1.93e10, 4 -> 2.07e10, 6 -> 2.39e10 lookups/s; a chunk's chain of copies and kernels is several times as
      // long as its compute, so the pipeline has to be that deep to keep the SMs busy
      const char* e = getenv("MAZU_B200_RUN_STREAMS");
      const int v = e ? atoi(e) : 6;
      return v < 2 ? 2 : (v > MAXB ? MAXB : v);
    }();
    const int NB = runs_dev ? nb_knob : 2;
    StreamPair sp;
    struct ExtraStreams {
      cudaStream_t s[MAXB - 2] = {};
      ~ExtraStreams() {
        for (auto x : s)
          if (x) cudaStreamDestroy(x);
      }
    } extra;
    for (int b = 2; b < NB; ++b) MZ_CUDA(cudaStreamCreateWithFlags(&extra.s[b - 2], cudaStreamNonBlocking));
    cudaStream_t st[MAXB] = {sp.s[0], sp.s[1], extra.s[0], extra.s[1], extra.s[2], extra.s[3]};
    PoolScratch scratch(idx->pool, sp.s[0]);
    void *d_bases[MAXB] = {}, *d_ro[MAXB] = {}, *d_ko[MAXB] = {}, *d_hits[MAXB] = {};
    void* d_counts = scratch.get(3 * 8);
    MZ_CUDA(cudaMemsetAsync(d_counts, 0, 24, sp.s[0]));
    for (int b = 0; b < NB; ++b) {
      d_bases[b] = scratch.get(max_bases + 16);
      if (!uniform_read_len) {
        d_ro[b] = scratch.get((max_reads + 1) * 8);
        d_ko[b] = scratch.get((max_reads + 1) * 8);
      }
      if ((out_hits && !dev_out) || (ro && !fused_runs)) d_hits[b] = scratch.get(max_slots * rec + 16);
    }
    const u64 wpr = uniform_read_len ? (uniform_read_len + 31) / 32 : 0, mpr = uniform_read_len ? (uniform_read_len + 63) / 64 : 0;
    void *d_pw[MAXB] = {}, *d_pm[MAXB] = {}, *d_codes2[MAXB] = {};
    if (packed_in)
      for (int b = 0; b < NB; ++b) {
        d_pw[b] = scratch.get(max_reads * wpr * 8 + 16);
        if (ro->packed_nmask) d_pm[b] = scratch.get(max_reads * mpr * 8 + 16);
      }
    if (ro && ro->codes2)
      for (int b = 0; b < NB; ++b) d_codes2[b] = scratch.get(max_slots / 4 + 16);
    u64* d_base = nullptr;
    u64* d_total_copy[MAXB] = {};
    cudaEvent_t base_ev = nullptr;
    struct EvGuard {
      cudaEvent_t* e;
      ~EvGuard() {
        if (*e) cudaEventDestroy(*e);
      }
    } ev_guard{&base_ev};
    if (runs_dev) {
      d_base = (u64*)scratch.get(8);
      MZ_CUDA(cudaMemcpyAsync(d_base, &ro->base, 8, cudaMemcpyHostToDevice, sp.s[0]));  // the running index starts at this call's first run
      for (int b = 0; b < NB; ++b) d_total_copy[b] = (u64*)scratch.get(8);
      MZ_CUDA(cudaEventCreateWithFlags(&base_ev, cudaEventDisableTiming));
      MZ_CUDA(cudaEventRecord(base_ev, sp.s[0]));
    }
    // hit runs: codes, per-read run counts / offsets and (worst case: every slot starts a run) the run records, per buffer;
    // the offsets come back through a small pinned array because the host needs each chunk's total to place its runs
    void *d_codes[MAXB] = {}, *d_rc[MAXB] = {}, *d_rro[MAXB] = {}, *d_runs[MAXB] = {};
    u64* h_rro[2] = {nullptr, nullptr};
    struct PinnedPair {
      u64** p;
      ~PinnedPair() {
        for (int i = 0; i < 2; ++i)
          if (p[i]) cudaFreeHost(p[i]);
      }
    } pinned_guard{h_rro};
    if (ro) {
      for (int b = 0; b < NB; ++b) {
        d_codes[b] = scratch.get(max_slots + 16);
        d_rc[b] = scratch.get((max_reads + 1) * 8);
        d_rro[b] = scratch.get((max_reads + 1) * 8);
        d_runs[b] = scratch.get(max_slots * 16 + 16);
        if (!runs_dev) MZ_CUDA(cudaHostAlloc((void**)&h_rro[b], (max_reads + 1) * 8, cudaHostAllocDefault));
      }
      ro->n_runs = 0;
      if (n_reads == 0 && ro->read_run_offsets && ro->write_end) ro->read_run_offsets[0] = ro->base;
    }
    const u64 uniform_slots = uniform_read_len >= k ? uniform_read_len - k + 1 : 0;
    // second half of a chunk in hit-run mode: wait for its offsets, place its runs, publish its offsets
    auto finish_runs = [&](size_t c, int bb) {
      const u64 r0 = cuts[c], nr = cuts[c + 1] - cuts[c];
      MZ_CUDA(cudaStreamSynchronize(sp.s[bb]));
      const u64 total = h_rro[bb][nr];
      if (ro->base + ro->n_runs + total > ro->cap_runs) {
        ro->n_runs += total;  // keep counting so the caller learns the capacity it needs
      } else {
        const u64 at = ro->base + ro->n_runs;
        if (total) MZ_CUDA(cudaMemcpyAsync(ro->runs + at, d_runs[bb], total * 16, cudaMemcpyDeviceToHost, sp.s[bb]));
        for (u64 i = 0; i < nr; ++i) ro->read_run_offsets[r0 + i] = at + h_rro[bb][i];
        ro->n_runs += total;
        if (ro->write_end || r0 + nr < n_reads) ro->read_run_offsets[r0 + nr] = at + total;
      }
    };
    // codes of one chunk back to the host: one byte per slot, or 2 bits per slot (chunks start at a multiple of 4 slots)
    auto copy_codes = [&](int bb, u64 s0, u64 ns, cudaStream_t s) {
      if (!ns) return;
      if (ro->codes2) {
        if (!fused_runs) {
          pack_codes_kernel<<<idx->sm_count * 4, 256, 0, s>>>((const u8*)d_codes[bb], ns, (u8*)d_codes2[bb]);
          MZ_CUDA(cudaGetLastError());
        }
        MZ_CUDA(cudaMemcpyAsync(ro->codes + s0 / 4, d_codes2[bb], (ns + 3) / 4, cudaMemcpyDeviceToHost, s));
      } else {
        MZ_CUDA(cudaMemcpyAsync(ro->codes + s0, d_codes[bb], ns, cudaMemcpyDeviceToHost, s));
      }
    };
    for (int bb = 1; bb < NB; ++bb) scratch.publish(st[bb]);
    int b = 0;
    for (size_t c = 0; c + 1 < cuts.size(); ++c, b = (b + 1) % NB) {
      u64 r0 = cuts[c], r1 = cuts[c + 1];
      u64 b0 = base_off(r0), nb = base_off(r1) - b0;
      u64 s0 = slot_off(r0), ns = slot_off(r1) - s0;
      cudaStream_t s = st[b];
      if (packed_in) {  // 2-bit words over PCIe, ASCII only ever exists in HBM
        const u64 nrc = r1 - r0;
        MZ_CUDA(cudaMemcpyAsync(d_pw[b], ro->packed_words + r0 * wpr, nrc * wpr * 8, cudaMemcpyHostToDevice, s));
        if (ro->packed_nmask) MZ_CUDA(cudaMemcpyAsync(d_pm[b], ro->packed_nmask + r0 * mpr, nrc * mpr * 8, cudaMemcpyHostToDevice, s));
        if (!fused_runs) {  // the fused run kernel reads the packed words itself (stage_encode_packed)
          unpack_reads_kernel<<<idx->sm_count * 8, 256, 0, s>>>((const u64*)d_pw[b], (const u64*)d_pm[b], nrc, uniform_read_len, (u8*)d_bases[b]);
          MZ_CUDA(cudaGetLastError());
        }
      } else {
        MZ_CUDA(cudaMemcpyAsync(d_bases[b], bases + b0, nb, cudaMemcpyHostToDevice, s));
      }
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
      else if (out_hits || ro) dh = (char*)d_hits[b] - (uniform_read_len ? 0 : s0 * rec);
      if (!fused_runs) launch_query_reads(idx, dbases, dro, r1 - r0, uniform_read_len, mode, dko, dh, compact, (u64*)d_counts, s);
      if (out_hits && !dev_out && ns) MZ_CUDA(cudaMemcpyAsync((char*)out_hits + s0 * rec, d_hits[b], ns * rec, cudaMemcpyDeviceToHost, s));
      if (ro) {
        const u64 nr = r1 - r0;
        const Hit* hh = (const Hit*)dh;
        u8* cc = (u8*)d_codes[b] - (uniform_read_len ? 0 : s0);
        const int grid = (int)std::min<u64>((nr + 7) / 8, (u64)idx->sm_count * 8);
        if (fused_runs) {  // one kernel: lookups, run codes, run records (chunk-local), per-read run offsets
          launch_query_reads_runs(idx, dbases, dro, nr, uniform_read_len, dko, (u64*)d_counts, cc, (Hit*)d_runs[b], (u64*)d_rro[b], max_slots, s,
                                  packed_in ? (const u64*)d_pw[b] : nullptr, packed_in ? (const u64*)d_pm[b] : nullptr,
                                  ro->codes2 ? (u8*)d_codes2[b] : nullptr);  // 2-bit codes straight from the kernel
        } else {
          MZ_CUDA(cudaMemsetAsync(d_rc[b], 0, (nr + 1) * 8, s));
          hit_run_codes_kernel<<<grid, 256, 0, s>>>(hh, dko, nr, uniform_slots, cc, (u64*)d_rc[b]);
          MZ_CUDA(cudaGetLastError());
          device_exclusive_scan((const u64*)d_rc[b], (u64*)d_rro[b], nr, idx->pool, s);
        }
        if (runs_dev) {
          // the running base lives on the device: this chunk reads it after the previous chunk (other stream) has advanced it
          MZ_CUDA(cudaMemcpyAsync(d_total_copy[b], (u64*)d_rro[b] + nr, 8, cudaMemcpyDeviceToDevice, s));  // the fill kernel overwrites offsets in place
          MZ_CUDA(cudaStreamWaitEvent(s, base_ev, 0));
          if (!fused_runs) hit_run_fill_kernel<<<grid, 256, 0, s>>>(hh, cc, dko, nr, uniform_slots, (const u64*)d_rro[b], 0, max_slots, (Hit*)d_runs[b]);
          hit_run_publish_kernel<<<idx->sm_count * 2, 256, 0, s>>>((u64*)d_rro[b], nr, (const Hit*)d_runs[b], d_total_copy[b], d_base, ro->cap_runs,
                                                                   runs_dev);
          hit_run_advance_kernel<<<1, 32, 0, s>>>(d_base, d_total_copy[b]);
          MZ_CUDA(cudaGetLastError());
          MZ_CUDA(cudaEventRecord(base_ev, s));
          copy_codes(b, s0, ns, s);
          if (nr) MZ_CUDA(cudaMemcpyAsync(ro->read_run_offsets + r0, d_rro[b], nr * 8, cudaMemcpyDeviceToHost, s));
        } else {
          if (!fused_runs) hit_run_fill_kernel<<<grid, 256, 0, s>>>(hh, cc, dko, nr, uniform_slots, (const u64*)d_rro[b], 0, max_slots, (Hit*)d_runs[b]);
          MZ_CUDA(cudaGetLastError());
          copy_codes(b, s0, ns, s);
          MZ_CUDA(cudaMemcpyAsync(h_rro[b], d_rro[b], (nr + 1) * 8, cudaMemcpyDeviceToHost, s));
          if (c > 0) finish_runs(c - 1, b ^ 1);  // the previous chunk finishes while this one runs
        }
      }
    }
    if (ro && !runs_dev && cuts.size() > 1) finish_runs(cuts.size() - 2, b ^ 1);
    if (ro && runs_dev) {
      for (int bb = 0; bb < NB; ++bb) MZ_CUDA(cudaStreamSynchronize(st[bb]));
      u64 end = 0;
      MZ_CUDA(cudaMemcpy(&end, d_base, 8, cudaMemcpyDeviceToHost));
      ro->n_runs = end - ro->base;
      if (ro->write_end) ro->read_run_offsets[n_reads] = end;
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

mazu_status_t mazu_b200_query_reads_runs(const mazu_index_t* idx, const uint8_t* bases, const uint64_t* read_offsets, uint64_t n_reads,
                                         uint64_t uniform_read_len, int32_t mode, uint64_t* kmer_offsets, uint8_t* out_codes,
                                         mazu_hit_t* out_runs, uint64_t cap_runs, uint64_t* out_read_run_offsets, uint64_t* out_n_runs,
                                         uint64_t* counts) {
  RunsOut ro{out_codes, out_runs, cap_runs, out_read_run_offsets};
  mazu_status_t rc = query_reads_impl(idx, bases, read_offsets, n_reads, uniform_read_len, mode, kmer_offsets, nullptr, 0, counts, MAZU_MEM_HOST,
                                      nullptr, &ro);
  if (out_n_runs) *out_n_runs = ro.n_runs;
  if (rc == MAZU_OK && ro.n_runs > cap_runs) {
    g_err = "run capacity too small: need " + std::to_string(ro.n_runs) + " records";
    return MAZU_ERR_INVALID_ARG;
  }
  return rc;
}

mazu_status_t mazu_b200_query_reads_runs_packed(const mazu_index_t* idx, const uint64_t* packed_reads, const uint64_t* n_mask, uint64_t n_reads,
                                                uint64_t read_len, int32_t mode, uint8_t* out_codes2, mazu_hit_t* out_runs, uint64_t cap_runs,
                                                uint64_t* out_read_run_offsets, uint64_t* out_n_runs, uint64_t* counts) {
  if (!read_len || (n_reads && !packed_reads)) {
    g_err = "packed reads are uniform: read_len > 0 and packed_reads are required";
    return MAZU_ERR_INVALID_ARG;
  }
  if (idx && ((read_len >= idx->unitigs->k ? read_len - idx->unitigs->k + 1 : 0) & 3)) {
    g_err = "packed codes need a multiple of 4 k-mer slots per read (read_len - k + 1)";
    return MAZU_ERR_INVALID_ARG;
  }
  RunsOut ro{out_codes2, out_runs, cap_runs, out_read_run_offsets};
  ro.packed_words = packed_reads;
  ro.packed_nmask = n_mask;
  ro.codes2 = true;
  mazu_status_t rc = query_reads_impl(idx, nullptr, nullptr, n_reads, read_len, mode, nullptr, nullptr, 0, counts, MAZU_MEM_HOST, nullptr, &ro);
  if (out_n_runs) *out_n_runs = ro.n_runs;
  if (rc == MAZU_OK && ro.n_runs > cap_runs) {
    g_err = "run capacity too small: need " + std::to_string(ro.n_runs) + " records";
    return MAZU_ERR_INVALID_ARG;
  }
  return rc;
}

// Hit intervals (include/mazu_b200.h): packed uniform short reads in, one 16-byte record per run out.  Its own small pipeline:
// per chunk H2D of the 2-bit words, the fused run kernel in interval mode (chunk-local records + count), then the records go
// to the caller's buffer at a running base -- published by a kernel when the buffer is pinned (no host synchronisation at
// all), or copied after a per-chunk synchronisation when it is pageable.
// `bases` set: the reads arrive as ASCII (n_reads x read_len bytes, the form the reference takes them in); else 2-bit packed.
static mazu_status_t query_reads_intervals_impl(const mazu_index_t* idx, const uint8_t* bases, const uint64_t* packed_reads, const uint64_t* n_mask,
                                                uint64_t n_reads, uint64_t read_len, int32_t mode, mazu_hit_interval_t* out_intervals,
                                                uint64_t cap, uint64_t* out_n, uint64_t* counts) {
  static_assert(sizeof(mazu_hit_interval_t) == 16, "interval records are 16 bytes");
  u64 n_total = 0;
  mazu_status_t rc = guarded([&] {
    if (!idx || !read_len || (n_reads && !packed_reads && !bases) || (cap && !out_intervals)) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    if (mode != MAZU_MODE_RANDOM && mode != MAZU_MODE_STREAMING) throw Error(MAZU_ERR_INVALID_ARG, "unknown query mode");
    if (mode == MAZU_MODE_STREAMING && !idx->kmers_unique)
      throw Error(MAZU_ERR_INVALID_ARG, "hit intervals come from the random-access kernel: streaming mode needs an index with unique k-mers");
    const u32 k = idx->unitigs->k;
    const u64 slots = read_len >= k ? read_len - k + 1 : 0;
    if (slots > (u64)QR_CHUNK) throw Error(MAZU_ERR_INVALID_ARG, "hit intervals are for short reads: read_len - k + 1 <= 128");
    if (n_reads >> 32) throw Error(MAZU_ERR_INVALID_ARG, "more than 2^32 reads in one call");
    if (counts) counts[0] = counts[1] = counts[2] = 0;
    if (!n_reads || !slots) return;
    DeviceGuard g(idx->device);
    const u64 wpr = (read_len + 31) / 32, mpr = (read_len + 63) / 64;
    u64 TARGET = 16ull << 20;
    if (const char* e = getenv("MAZU_B200_CHUNK_MIB")) {
      long v = atol(e);
      if (v > 0) TARGET = (u64)v << 20;
    }
    const u64 rpc = std::max<u64>(4, TARGET / read_len);  // reads per chunk
    static const int NB = 6;
    uint4* out_dev = nullptr;  // the caller's buffer as the device sees it (pinned memory), or null
    {
      cudaPointerAttributes at{};
      if (cap && cudaPointerGetAttributes(&at, out_intervals) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) out_dev = (uint4*)at.devicePointer;
      else cudaGetLastError();
    }
    struct Streams {
      cudaStream_t s[NB] = {};
      cudaEvent_t base_ev = nullptr;
      ~Streams() {
        for (auto x : s)
          if (x) {
            cudaStreamSynchronize(x);
            cudaStreamDestroy(x);
          }
        if (base_ev) cudaEventDestroy(base_ev);
      }
    } st;
    for (auto& x : st.s) MZ_CUDA(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
    MZ_CUDA(cudaEventCreateWithFlags(&st.base_ev, cudaEventDisableTiming));
    PoolScratch scratch(idx->pool, st.s[0]);
    const u64 max_reads = std::min(rpc, n_reads), max_slots = max_reads * slots;
    void *d_pw[NB], *d_pm[NB] = {}, *d_iv[NB];
    u64* d_cur[NB];  // run cursor of the chunk (its run count afterwards)
    u64* d_counts = (u64*)scratch.get(24);
    u64* d_base = (u64*)scratch.get(8);
    MZ_CUDA(cudaMemsetAsync(d_counts, 0, 24, st.s[0]));
    MZ_CUDA(cudaMemsetAsync(d_base, 0, 8, st.s[0]));
    for (int b = 0; b < NB; ++b) {
      d_pw[b] = scratch.get(bases ? max_reads * read_len + 16 : max_reads * wpr * 8 + 16);  // ASCII bases, or 2-bit words
      if (n_mask && !bases) d_pm[b] = scratch.get(max_reads * mpr * 8 + 16);
      d_iv[b] = scratch.get(max_slots * 16 + 16);  // worst case: every slot starts a run
      d_cur[b] = (u64*)scratch.get(8);
    }
    for (int b = 1; b < NB; ++b) scratch.publish(st.s[b]);
    MZ_CUDA(cudaEventRecord(st.base_ev, st.s[0]));
    u64 host_base = 0;  // pageable path: records placed so far
    int b = 0;
    for (u64 r0 = 0; r0 < n_reads; r0 += rpc, b = (b + 1) % NB) {
      const u64 nr = std::min(rpc, n_reads - r0);
      cudaStream_t s = st.s[b];
      if (bases) {
        MZ_CUDA(cudaMemcpyAsync(d_pw[b], bases + r0 * read_len, nr * read_len, cudaMemcpyHostToDevice, s));
        launch_query_reads_runs(idx, (const u8*)d_pw[b], nullptr, nr, read_len, nullptr, d_counts, nullptr, nullptr, nullptr, max_slots, s, nullptr,
                                nullptr, nullptr, (uint4*)d_iv[b], r0, d_cur[b]);
      } else {
        MZ_CUDA(cudaMemcpyAsync(d_pw[b], packed_reads + r0 * wpr, nr * wpr * 8, cudaMemcpyHostToDevice, s));
        if (n_mask) MZ_CUDA(cudaMemcpyAsync(d_pm[b], n_mask + r0 * mpr, nr * mpr * 8, cudaMemcpyHostToDevice, s));
        launch_query_reads_runs(idx, nullptr, nullptr, nr, read_len, nullptr, d_counts, nullptr, nullptr, nullptr, max_slots, s, (const u64*)d_pw[b],
                                (const u64*)d_pm[b], nullptr, (uint4*)d_iv[b], r0, d_cur[b]);
      }
      if (out_dev) {
        MZ_CUDA(cudaStreamWaitEvent(s, st.base_ev, 0));  // the running base: after the previous chunk has advanced it
        hit_run_publish_kernel<<<idx->sm_count * 2, 256, 0, s>>>(nullptr, 0, (const Hit*)d_iv[b], d_cur[b], d_base, cap, (Hit*)out_dev);
        hit_run_advance_kernel<<<1, 32, 0, s>>>(d_base, d_cur[b]);
        MZ_CUDA(cudaGetLastError());
        MZ_CUDA(cudaEventRecord(st.base_ev, s));
      } else {
        u64 total = 0;
        MZ_CUDA(cudaMemcpyAsync(&total, d_cur[b], 8, cudaMemcpyDeviceToHost, s));
        MZ_CUDA(cudaStreamSynchronize(s));
        if (host_base < cap && total) {
          const u64 n_fit = std::min(total, cap - host_base);
          MZ_CUDA(cudaMemcpyAsync((uint4*)out_intervals + host_base, d_iv[b], n_fit * 16, cudaMemcpyDeviceToHost, s));
        }
        host_base += total;
      }
    }
    for (auto x : st.s) MZ_CUDA(cudaStreamSynchronize(x));
    if (out_dev) MZ_CUDA(cudaMemcpy(&n_total, d_base, 8, cudaMemcpyDeviceToHost));
    else n_total = host_base;
    if (counts) MZ_CUDA(cudaMemcpy(counts, d_counts, 24, cudaMemcpyDeviceToHost));
  });
  if (out_n) *out_n = n_total;
  if (rc == MAZU_OK && n_total > cap) {
    g_err = "interval capacity too small: need " + std::to_string(n_total) + " records";
    return MAZU_ERR_INVALID_ARG;
  }
  return rc;
}

mazu_status_t mazu_b200_query_reads_intervals_packed(const mazu_index_t* idx, const uint64_t* packed_reads, const uint64_t* n_mask,
                                                     uint64_t n_reads, uint64_t read_len, int32_t mode, mazu_hit_interval_t* out_intervals,
                                                     uint64_t cap, uint64_t* out_n, uint64_t* counts) {
  if (n_reads && !packed_reads) {
    g_err = "null argument";
    return MAZU_ERR_INVALID_ARG;
  }
  return query_reads_intervals_impl(idx, nullptr, packed_reads, n_mask, n_reads, read_len, mode, out_intervals, cap, out_n, counts);
}
mazu_status_t mazu_b200_query_reads_intervals(const mazu_index_t* idx, const uint8_t* bases, uint64_t n_reads, uint64_t read_len, int32_t mode,
                                              mazu_hit_interval_t* out_intervals, uint64_t cap, uint64_t* out_n, uint64_t* counts) {
  if (n_reads && !bases) {
    g_err = "null argument";
    return MAZU_ERR_INVALID_ARG;
  }
  return query_reads_intervals_impl(idx, bases, nullptr, nullptr, n_reads, read_len, mode, out_intervals, cap, out_n, counts);
}

static mazu_status_t expand_hit_intervals_impl(const mazu_index_t* idx, const mazu_hit_interval_t* intervals, uint64_t n_intervals,
                                               const uint64_t* n_mask, const uint8_t* bases, uint64_t n_reads, uint64_t read_len,
                                               mazu_hit_t* out_hits) {
  return guarded([&] {
    if (!idx || (n_intervals && !intervals) || !read_len) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    const u32 k = idx->unitigs->k;
    const u64 slots = read_len >= k ? read_len - k + 1 : 0, mpr = (read_len + 63) / 64;
    if (n_reads && slots && !out_hits) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    const UnitigSetHost& us = *idx->unitigs;
    const unsigned T = host_threads();
    // every slot a miss, or skipped where its window holds a masked base
    parallel_ranges(n_reads, T, [&](unsigned, u64 lo, u64 hi) {
      for (u64 r = lo; r < hi; ++r) {
        const u64* m = n_mask ? n_mask + r * mpr : nullptr;
        const u8* bs = bases ? bases + r * read_len : nullptr;  // ASCII reads: a base is masked when it is not ACGTacgt
        auto masked = [&](u64 j) { return bs ? base_code(bs[j]) > 3 : (m && ((m[j >> 6] >> (j & 63)) & 1ULL)); };
        long long last_bad = -1;  // most recent masked base at or before the window's last base
        for (u64 j = 0; j + 1 < k && j < read_len; ++j)
          if (masked(j)) last_bad = (long long)j;
        for (u64 i = 0; i < slots; ++i) {
          const u64 j = i + k - 1;
          if (masked(j)) last_bad = (long long)j;
          out_hits[r * slots + i] = mazu_hit_t{~0u, ~0u, ~0u, last_bad >= (long long)i ? (u32)MAZU_SKIPPED : (u32)MAZU_NO_MATCH};
        }
      }
    });
    std::atomic<bool> bad{false};
    parallel_ranges(n_intervals, T, [&](unsigned, u64 lo, u64 hi) {
      for (u64 i = lo; i < hi; ++i) {
        const mazu_hit_interval_t& v = intervals[i];
        if (v.read >= n_reads || (u64)v.start + v.len > slots || v.unitig_id >= us.n_unitigs()) {
          bad = true;
          continue;
        }
        const bool twin = v.pos_o >> 31;
        const u32 ulen = (u32)(us.accum[v.unitig_id + 1] - us.accum[v.unitig_id]), pos0 = v.pos_o & 0x7FFFFFFFu;
        mazu_hit_t* o = out_hits + (u64)v.read * slots + v.start;
        for (u32 j = 0; j < v.len; ++j) o[j] = mazu_hit_t{v.unitig_id, ulen, twin ? pos0 - j : pos0 + j, twin ? (u32)MAZU_TWIN_MATCH : (u32)MAZU_IDENTITY_MATCH};
      }
    });
    if (bad) throw Error(MAZU_ERR_INVALID_DATA, "an interval record points outside the batch or the unitig set");
  });
}
mazu_status_t mazu_b200_expand_hit_intervals(const mazu_index_t* idx, const mazu_hit_interval_t* intervals, uint64_t n_intervals,
                                             const uint64_t* n_mask, uint64_t n_reads, uint64_t read_len, mazu_hit_t* out_hits) {
  return expand_hit_intervals_impl(idx, intervals, n_intervals, n_mask, nullptr, n_reads, read_len, out_hits);
}
mazu_status_t mazu_b200_expand_hit_intervals_ascii(const mazu_index_t* idx, const mazu_hit_interval_t* intervals, uint64_t n_intervals,
                                                   const uint8_t* bases, uint64_t n_reads, uint64_t read_len, mazu_hit_t* out_hits) {
  if (n_reads && !bases) {
    g_err = "null argument";
    return MAZU_ERR_INVALID_ARG;
  }
  return expand_hit_intervals_impl(idx, intervals, n_intervals, nullptr, bases, n_reads, read_len, out_hits);
}

mazu_status_t mazu_b200_pack_reads(const uint8_t* bases, uint64_t n_reads, uint64_t read_len, uint64_t* out_words, uint64_t* out_n_mask,
                                   uint64_t* n_non_acgt) {
  return guarded([&] {
    if (!read_len || (n_reads && (!bases || !out_words))) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    const u64 wpr = (read_len + 31) / 32, mpr = (read_len + 63) / 64;
    std::atomic<u64> bad{0};
    parallel_ranges(n_reads, host_threads(), [&](unsigned, u64 lo, u64 hi) {
      u64 nb = 0;
      for (u64 r = lo; r < hi; ++r) {
        u64* w = out_words + r * wpr;
        u64* m = out_n_mask ? out_n_mask + r * mpr : nullptr;
        for (u64 j = 0; j < wpr; ++j) w[j] = 0;
        if (m)
          for (u64 j = 0; j < mpr; ++j) m[j] = 0;
        const u8* b = bases + r * read_len;
        u64 j = 0;
        for (; j + 8 <= read_len; j += 8) {  // eight bases per step (kmer.hpp: pack8_ascii); j is a multiple of 8: no word is straddled
          u64 v;
          memcpy(&v, b + j, 8);
          u32 c16, inv;
          pack8_ascii(v, c16, inv);
          w[j >> 5] |= (u64)c16 << (2 * (j & 31));
          if (inv) {
            nb += (u64)__builtin_popcount(inv);
            if (m) m[j >> 6] |= (u64)inv << (j & 63);
          }
        }
        for (; j < read_len; ++j) {
          const u32 c = base_code(b[j]);
          if (c > 3) {
            ++nb;
            if (m) m[j >> 6] |= 1ULL << (j & 63);
          } else {
            w[j >> 5] |= (u64)c << (2 * (j & 31));
          }
        }
      }
      bad += nb;
    });
    if (n_non_acgt) *n_non_acgt = bad.load();
  });
}

mazu_status_t mazu_b200_expand_hit_runs_packed(const uint8_t* codes2, const mazu_hit_t* runs, const uint64_t* read_run_offsets, uint64_t n_reads,
                                               uint64_t uniform_slots, mazu_hit_t* out_hits) {
  return guarded([&] {
    if (n_reads && (!codes2 || !read_run_offsets || !out_hits)) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    parallel_ranges(n_reads, host_threads(), [&](unsigned, u64 lo, u64 hi) {
      for (u64 r = lo; r < hi; ++r) {
        const u64 s0 = r * uniform_slots;
        u64 ri = read_run_offsets[r];
        mazu_hit_t prev{~0u, ~0u, ~0u, MAZU_NO_MATCH};
        for (u64 i = 0; i < uniform_slots; ++i) {
          const u64 s = s0 + i;
          mazu_hit_t h{~0u, ~0u, ~0u, MAZU_NO_MATCH};
          switch ((codes2[s >> 2] >> (2 * (s & 3))) & 3u) {
            case 1:
              h = prev;
              h.pos = prev.match == MAZU_IDENTITY_MATCH ? prev.pos + 1u : prev.pos - 1u;
              break;
            case 2:
              h = runs[ri++];
              break;
            case 3:
              h.match = MAZU_SKIPPED;
              break;
            default:
              break;
          }
          out_hits[s] = h;
          prev = h;
        }
      }
    });
  });
}

mazu_status_t mazu_b200_expand_hit_runs(const uint8_t* codes, const mazu_hit_t* runs, const uint64_t* read_run_offsets,
                                        const uint64_t* kmer_offsets, uint64_t n_reads, uint64_t uniform_slots, mazu_hit_t* out_hits) {
  return guarded([&] {
    if (n_reads && (!codes || !read_run_offsets || !out_hits)) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    const unsigned T = host_threads();
    parallel_ranges(n_reads, T, [&](unsigned, u64 lo, u64 hi) {
      for (u64 r = lo; r < hi; ++r) {
        const u64 s0 = kmer_offsets ? kmer_offsets[r] : r * uniform_slots;
        const u64 ns = kmer_offsets ? kmer_offsets[r + 1] - s0 : uniform_slots;
        u64 ri = read_run_offsets[r];
        mazu_hit_t prev{~0u, ~0u, ~0u, MAZU_NO_MATCH};
        for (u64 i = 0; i < ns; ++i) {
          mazu_hit_t h{~0u, ~0u, ~0u, MAZU_NO_MATCH};
          switch (codes[s0 + i]) {
            case 1:  // continues the previous slot's run
              h = prev;
              h.pos = prev.match == MAZU_IDENTITY_MATCH ? prev.pos + 1u : prev.pos - 1u;
              break;
            case 2:
              h = runs[ri++];
              break;
            case 3:
              h.match = MAZU_SKIPPED;
              break;
            default:
              break;
          }
          out_hits[s0 + i] = h;
          prev = h;
        }
      }
    });
  });
}

mazu_status_t mazu_b200_query_reads_compact(const mazu_index_t* idx, const uint8_t* bases, const uint64_t* read_offsets, uint64_t n_reads,
                                            uint64_t uniform_read_len, int32_t mode, uint64_t* kmer_offsets, mazu_hit8_t* out_hits,
                                            uint64_t* counts, int32_t mem, void* stream) {
  if (idx && !idx->compact_ok) {  // pos shares a word with the match type
    g_err = "compact hit records need every unitig shorter than 2^30 bases";
    return MAZU_ERR_INVALID_ARG;
  }
  return query_reads_impl(idx, bases, read_offsets, n_reads, uniform_read_len, mode, kmer_offsets, out_hits, 1, counts, mem, stream);
}

// ---------------------------------------------------------------------------------------------
// Multi-GPU: replicate the device tables, shard the reads
// ---------------------------------------------------------------------------------------------
mazu_status_t mazu_b200_index_replicate(const mazu_index_t* src, const int32_t* devices, int32_t n_devices, mazu_index_t** out) {
  return guarded([&] {
    if (!src || !devices || !out || n_devices <= 0) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    int n_dev = 0;
    MZ_CUDA(cudaGetDeviceCount(&n_dev));
    for (int i = 0; i < n_devices; ++i) {
      out[i] = nullptr;
      if (devices[i] < 0 || devices[i] >= n_dev) throw Error(MAZU_ERR_INVALID_ARG, "device ordinal out of range");
    }
    {
      DeviceGuard gs(src->device);
      MZ_CUDA(cudaDeviceSynchronize());  // everything that built `src` has finished
    }
    std::vector<std::unique_ptr<mazu_index>> made;
    for (int i = 0; i < n_devices; ++i) {
      const int dev = devices[i];
      DeviceGuard g(dev);
      if (dev != src->device) {  // direct peer copies (NVLink) when the two devices can reach each other; staged by the driver otherwise
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, dev, src->device) == cudaSuccess && can) {
          cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) MZ_CUDA(e);
          cudaGetLastError();
        }
      }
      auto ix = std::make_unique<mazu_index>();
      ix->device = dev;
      ix->unitigs = src->unitigs;
      ix->k2u = src->k2u;
      ix->u2pos = src->u2pos;
      ix->refs = src->refs;
      ix->compact_ok = src->compact_ok;
      ix->kmers_unique = src->kmers_unique;
      ix->view = src->view;
      ix->tables = src->tables;
      cudaDeviceProp prop;
      MZ_CUDA(cudaGetDeviceProperties(&prop, dev));
      ix->sm_count = prop.multiProcessorCount;
      {
        cudaMemPoolProps pp{};
        pp.allocType = cudaMemAllocationTypePinned;
        pp.handleTypes = cudaMemHandleTypeNone;
        pp.location.type = cudaMemLocationTypeDevice;
        pp.location.id = dev;
        MZ_CUDA(cudaMemPoolCreate(&ix->pool, &pp));
        unsigned long long keep = ~0ULL;
        MZ_CUDA(cudaMemPoolSetAttribute(ix->pool, cudaMemPoolAttrReleaseThreshold, &keep));
      }
      // copy every buffer of every group, remember where it went
      struct Span {
        const char* lo;
        const char* hi;
        char* to;
      };
      std::vector<Span> spans;
      auto copy_bufs = [&](const std::vector<DevBufP>& from, std::vector<DevBufP>& to, size_t& bytes) {
        for (const DevBufP& b : from) {
          auto nb = std::make_shared<DevBuf>(b->bytes, dev);
          if (b->bytes) MZ_CUDA(cudaMemcpyPeerAsync(nb->p, dev, b->p, src->device, b->bytes, 0));
          spans.push_back(Span{(const char*)b->p, (const char*)b->p + std::max<size_t>(b->bytes, 1), (char*)nb->p});
          to.push_back(nb);
          bytes += nb->bytes;
        }
      };
      if (src->d_unitigs) {
        ix->d_unitigs = std::make_shared<UnitigsDev>();
        copy_bufs(src->d_unitigs->bufs, ix->d_unitigs->bufs, ix->d_unitigs->bytes);
      }
      if (src->d_k2u) {
        ix->d_k2u = std::make_shared<K2UDev>();
        copy_bufs(src->d_k2u->bufs, ix->d_k2u->bufs, ix->d_k2u->bytes);
      }
      if (src->d_u2pos) {
        ix->d_u2pos = std::make_shared<U2PosDev>();
        copy_bufs(src->d_u2pos->bufs, ix->d_u2pos->bufs, ix->d_u2pos->bytes);
      }
      if (src->d_refs) {
        ix->d_refs = std::make_shared<RefsDev>();
        copy_bufs(src->d_refs->bufs, ix->d_refs->bufs, ix->d_refs->bytes);
      }
      auto remap = [&](auto& ptr) {
        using P = std::remove_reference_t<decltype(ptr)>;
        const char* p = (const char*)ptr;
        if (!p) return;
        for (const Span& sp : spans)
          if (p >= sp.lo && p < sp.hi) {
            ptr = (P)(sp.to + (p - sp.lo));
            return;
          }
        throw Error(MAZU_ERR_OTHER, "internal: an index pointer lies outside every device buffer of its handle");
      };
      IndexView& v = ix->view;
      auto remap_mphf = [&](RankedLevels& m) {
        remap(m.blocks);
        remap(m.fb_keys);
        remap(m.fb_vals);
      };
      remap(v.unitigs.useq);
      remap(v.unitigs.dir);
      remap(v.unitigs.starts);
      remap(v.unitigs.lines);
      remap_mphf(v.mphf);
      remap(v.pos.words);
      remap(v.sizes.blocks);
      remap(v.sizes.exceptions);
      remap_mphf(v.skew_mphf);
      remap(v.skew_pos.words);
      remap_mphf(v.sampled);
      remap(v.canonical_bits);
      remap(v.direction_bits);
      remap(v.ext_sizes.words);
      remap(v.ext_bases.words);
      remap(v.ctable_words);
      remap(v.contig_offsets.words);
      remap(v.refseq);
      remap(v.ref_prefix);
      for (auto& t : ix->tables) remap(t.first);
      if (ix->d_unitigs) ix->d_unitigs->view = v.unitigs;
      MZ_CUDA(cudaDeviceSynchronize());
      made.push_back(std::move(ix));
    }
    for (int i = 0; i < n_devices; ++i) out[i] = made[i].release();
  });
}

namespace {
struct ShardPlan {
  std::vector<u64> cut;   // n + 1 read boundaries
  std::vector<u64> slot;  // first k-mer slot of every shard (+ total)
};
// contiguous blocks of about equal size in bases
ShardPlan plan_shards(int n, const uint64_t* read_offsets, uint64_t n_reads, uint64_t uniform_read_len, u32 k, uint64_t* kmer_offsets) {
  ShardPlan p;
  p.cut.assign(n + 1, n_reads);
  p.cut[0] = 0;
  for (int i = 1; i < n; ++i) {
    if (uniform_read_len) p.cut[i] = n_reads * (u64)i / n;
    else {
      const u64 target = read_offsets[0] + (read_offsets[n_reads] - read_offsets[0]) * (u64)i / n;
      p.cut[i] = std::max<u64>(p.cut[i - 1], (u64)(std::lower_bound(read_offsets, read_offsets + n_reads + 1, target) - read_offsets));
      p.cut[i] = std::min<u64>(p.cut[i], n_reads);
    }
  }
  p.slot.assign(n + 1, 0);
  if (uniform_read_len) {
    const u64 per = uniform_read_len >= k ? uniform_read_len - k + 1 : 0;
    for (int i = 0; i <= n; ++i) p.slot[i] = p.cut[i] * per;
    if (kmer_offsets)
      for (u64 r = 0; r <= n_reads; ++r) kmer_offsets[r] = r * per;
  } else {
    u64 acc = 0;
    int sh = 0;
    for (u64 r = 0; r <= n_reads; ++r) {
      while (sh <= n && p.cut[sh] == r) p.slot[sh++] = acc;
      if (kmer_offsets) kmer_offsets[r] = acc;
      if (r < n_reads) {
        const u64 len = read_offsets[r + 1] - read_offsets[r];
        if (len >= k) acc += len - k + 1;
      }
    }
  }
  return p;
}
void check_replicas(const mazu_index_t* const* handles, int32_t n) {
  if (!handles || n <= 0) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
  for (int i = 0; i < n; ++i) {
    if (!handles[i]) throw Error(MAZU_ERR_INVALID_ARG, "null handle");
    if (handles[i]->unitigs->k != handles[0]->unitigs->k || handles[i]->unitigs->total_len() != handles[0]->unitigs->total_len() ||
        handles[i]->view.k2u_kind != handles[0]->view.k2u_kind)
      throw Error(MAZU_ERR_INVALID_ARG, "the handles of a sharded call must hold the same index");
  }
}
}  // namespace

mazu_status_t mazu_b200_query_reads_sharded(const mazu_index_t* const* handles, int32_t n_handles, const uint8_t* bases,
                                            const uint64_t* read_offsets, uint64_t n_reads, uint64_t uniform_read_len, int32_t mode,
                                            uint64_t* kmer_offsets, mazu_hit_t* out_hits, uint64_t* counts) {
  return guarded([&] {
    check_replicas(handles, n_handles);
    if (n_reads && !uniform_read_len && !read_offsets) throw Error(MAZU_ERR_INVALID_ARG, "read_offsets is required for ragged reads");
    const ShardPlan p = plan_shards(n_handles, read_offsets, n_reads, uniform_read_len, handles[0]->unitigs->k, kmer_offsets);
    std::vector<mazu_status_t> rc(n_handles, MAZU_OK);
    std::vector<std::string> err(n_handles);
    std::vector<std::array<u64, 3>> cnt(n_handles, std::array<u64, 3>{0, 0, 0});
    std::vector<std::thread> ts;
    for (int i = 0; i < n_handles; ++i)
      ts.emplace_back([&, i] {  // one host thread + its own streams per device
        const u64 r0 = p.cut[i], nr = p.cut[i + 1] - r0;
        if (!nr) return;
        const uint8_t* b = uniform_read_len ? bases + r0 * uniform_read_len : bases;
        rc[i] = query_reads_impl(handles[i], b, uniform_read_len ? nullptr : read_offsets + r0, nr, uniform_read_len, mode, nullptr,
                                 out_hits ? out_hits + p.slot[i] : nullptr, 0, cnt[i].data(), MAZU_MEM_HOST, nullptr, nullptr);
        if (rc[i] != MAZU_OK) err[i] = g_err;
      });
    for (auto& t : ts) t.join();
    for (int i = 0; i < n_handles; ++i)
      if (rc[i] != MAZU_OK) throw Error(rc[i], "shard " + std::to_string(i) + ": " + err[i]);
    if (counts)
      for (int j = 0; j < 3; ++j) {
        counts[j] = 0;
        for (int i = 0; i < n_handles; ++i) counts[j] += cnt[i][j];
      }
  });
}

mazu_status_t mazu_b200_query_reads_runs_sharded(const mazu_index_t* const* handles, int32_t n_handles, const uint8_t* bases,
                                                 const uint64_t* read_offsets, uint64_t n_reads, uint64_t uniform_read_len, int32_t mode,
                                                 uint64_t* kmer_offsets, uint8_t* out_codes, mazu_hit_t* out_runs, uint64_t cap_runs,
                                                 uint64_t* out_read_run_offsets, uint64_t* out_n_runs, uint64_t* counts) {
  return guarded([&] {
    check_replicas(handles, n_handles);
    if (n_reads && !uniform_read_len && !read_offsets) throw Error(MAZU_ERR_INVALID_ARG, "read_offsets is required for ragged reads");
    if (n_reads && (!out_codes || !out_read_run_offsets || (cap_runs && !out_runs))) throw Error(MAZU_ERR_INVALID_ARG, "null run buffers");
    const ShardPlan p = plan_shards(n_handles, read_offsets, n_reads, uniform_read_len, handles[0]->unitigs->k, kmer_offsets);
    const u64 region = cap_runs / (u64)n_handles;
    std::vector<mazu_status_t> rc(n_handles, MAZU_OK);
    std::vector<std::string> err(n_handles);
    std::vector<std::array<u64, 3>> cnt(n_handles, std::array<u64, 3>{0, 0, 0});
    std::vector<u64> n_runs(n_handles, 0);
    std::vector<std::thread> ts;
    for (int i = 0; i < n_handles; ++i)
      ts.emplace_back([&, i] {
        const u64 r0 = p.cut[i], nr = p.cut[i + 1] - r0;
        if (!nr) return;
        const uint8_t* b = uniform_read_len ? bases + r0 * uniform_read_len : bases;
        // shard i fills out_runs[i * region, (i + 1) * region); its offsets are global indexes; entry r0 + nr belongs to the next shard
        RunsOut ro{out_codes + p.slot[i], out_runs, (u64)(i + 1) * region, out_read_run_offsets + r0};
        ro.base = (u64)i * region;
        ro.write_end = false;
        rc[i] = query_reads_impl(handles[i], b, uniform_read_len ? nullptr : read_offsets + r0, nr, uniform_read_len, mode, nullptr, nullptr, 0,
                                 cnt[i].data(), MAZU_MEM_HOST, nullptr, &ro);
        n_runs[i] = ro.n_runs;
        if (rc[i] != MAZU_OK) err[i] = g_err;
      });
    for (auto& t : ts) t.join();
    for (int i = 0; i < n_handles; ++i)
      if (rc[i] != MAZU_OK) throw Error(rc[i], "shard " + std::to_string(i) + ": " + err[i]);
    u64 total = 0, worst = 0;
    for (int i = 0; i < n_handles; ++i) {
      total += n_runs[i];
      worst = std::max(worst, n_runs[i]);
    }
    if (out_n_runs) *out_n_runs = total;
    if (worst > region) {
      if (out_n_runs) *out_n_runs = worst * (u64)n_handles;
      throw Error(MAZU_ERR_INVALID_ARG, "run capacity too small: need " + std::to_string(worst * (u64)n_handles) + " records");
    }
    if (counts)
      for (int j = 0; j < 3; ++j) {
        counts[j] = 0;
        for (int i = 0; i < n_handles; ++i) counts[j] += cnt[i][j];
      }
  });
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
  // the staged kernel pays off when tiles lie inside long lists (config 4: 92 % vs 85 % of the copy peak); with short lists
  // (a read batch's hits: ~1 occurrence each) every tile takes the generic path and the plain kernel's larger grid is 17 % faster
  const bool long_lists = idx->view.n_occs >= 64 * std::max<u64>(1, idx->unitigs->n_unitigs());
  // hit batches with short lists go by TILES of QR_CHUNK records (project_tile_totals_kernel -> scan over tiles ->
  // get_ref_pos_pass2_kernel, which writes offsets and records): 1/128 of the scan, no search, no list-length array
  const bool by_tiles = project && n > 0 && !long_lists;
  const u64 n_tiles = by_tiles ? (n + QR_CHUNK - 1) / QR_CHUNK : 0;
  std::unique_ptr<PoolBuf> tile_base;
  TileMap tm{nullptr, nullptr, nullptr, 0, 0, n_tiles, n};
  const int grid_tiles = (int)std::max<u64>(1, std::min<u64>((n_tiles + 7) / 8, (u64)idx->sm_count * 8));
  const u64* d_total_src = d_offsets + n;
  if (by_tiles) {
    PoolBuf totals(idx->pool, (n_tiles + 1) * 8, s);
    tile_base = std::make_unique<PoolBuf>(idx->pool, (n_tiles + 1) * 8, s);
    MZ_CUDA(cudaMemsetAsync((u64*)totals.p + n_tiles, 0, 8, s));
    project_tile_totals_kernel<<<grid_tiles, 256, 0, s>>>(idx->view, d_hits, n, n_tiles, (u64*)totals.p);
    MZ_CUDA(cudaGetLastError());
    device_exclusive_scan((const u64*)totals.p, (u64*)tile_base->p, n_tiles, idx->pool, s);
    d_total_src = (const u64*)tile_base->p + n_tiles;
  } else {
    PoolBuf lens(idx->pool, (n + 1) * 8, s);
    MZ_CUDA(cudaMemsetAsync(lens.p, 0, (n + 1) * 8, s));
    if (n) {
      occ_lens_kernel<<<(int)std::min<u64>((n + 255) / 256, (u64)idx->sm_count * 8), 256, 0, s>>>(idx->view, d_uids, d_hits, n, (u64*)lens.p);
      MZ_CUDA(cudaGetLastError());
    }
    device_exclusive_scan((const u64*)lens.p, d_offsets, n, idx->pool, s);
  }
  u64 total = 0;
  bool need_total = mem == MAZU_MEM_HOST || out_total != nullptr;
  if (need_total) {
    MZ_CUDA(cudaMemcpyAsync(&total, d_total_src, 8, cudaMemcpyDeviceToHost, s));
    MZ_CUDA(cudaStreamSynchronize(s));
    if (out_total) *out_total = total;
  }
  if (mem == MAZU_MEM_HOST && !by_tiles) MZ_CUDA(cudaMemcpyAsync(out_offsets, d_offsets, (n + 1) * 8, cudaMemcpyDeviceToHost, s));
  if (!out || (need_total && total > cap)) {
    if (by_tiles) {  // the offsets come out of the emit pass: run it without records
      get_ref_pos_pass2_kernel<<<grid_tiles, 256, 0, s>>>(idx->view, tm, d_hits, (const u64*)tile_base->p, n, d_offsets, nullptr, 0, nullptr);
      MZ_CUDA(cudaGetLastError());
      if (mem == MAZU_MEM_HOST) MZ_CUDA(cudaMemcpyAsync(out_offsets, d_offsets, (n + 1) * 8, cudaMemcpyDeviceToHost, s));
    }
    if (!out) {
      if (mem == MAZU_MEM_HOST) MZ_CUDA(cudaStreamSynchronize(s));
      return;
    }
  }
  if (need_total && total > cap) {
    if (mem == MAZU_MEM_HOST) MZ_CUDA(cudaStreamSynchronize(s));
    throw Error(MAZU_ERR_INVALID_ARG, "output capacity too small: need " + std::to_string(total) + " records");
  }
  OccRec* d_o = (OccRec*)out;
  // the fill kernels clip to this many records: with device buffers and no out_total the host never learns the total, so
  // an undersized `out` is filled up to `cap` and out_offsets[n] tells the caller what the full output needs
  u64 fill_cap = cap;
  if (mem == MAZU_MEM_HOST) {
    d_out = scratch->get(total * 12 + 16);
    d_o = (OccRec*)d_out;
    fill_cap = total;
  }
  if (n) {
    // tiles of the OUTPUT are grid-strided; the kernels read the total from out_offsets[n]
    static const bool use_tma = [] {
      const char* e = getenv("MAZU_B200_OCC_TMA");
      return e ? atoi(e) != 0 : true;
    }();
    if (by_tiles) {
      get_ref_pos_pass2_kernel<<<grid_tiles, 256, 0, s>>>(idx->view, tm, d_hits, (const u64*)tile_base->p, n, d_offsets, d_o, fill_cap, nullptr);
      if (mem == MAZU_MEM_HOST) MZ_CUDA(cudaMemcpyAsync(out_offsets, d_offsets, (n + 1) * 8, cudaMemcpyDeviceToHost, s));
    } else if (use_tma && long_lists) {  // bulk-copy staged fill: one CTA of 24 warps per SM, 174 KB of dynamic shared memory
      std::call_once(idx->occ_attr_once, [] {
        cudaFuncSetAttribute(occ_fill_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OCC_TMA_SMEM);
        cudaFuncSetAttribute(occ_fill_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OCC_TMA_SMEM);
      });
      int grid = idx->sm_count;
      if (project) occ_fill_tma_kernel<true><<<grid, OCC_TMA_WARPS * 32, OCC_TMA_SMEM, s>>>(idx->view, d_uids, d_hits, n, d_offsets, d_o, fill_cap);
      else occ_fill_tma_kernel<false><<<grid, OCC_TMA_WARPS * 32, OCC_TMA_SMEM, s>>>(idx->view, d_uids, d_hits, n, d_offsets, d_o, fill_cap);
    } else {
      int grid = idx->sm_count * 8;
      if (project) occ_fill_kernel<true><<<grid, 256, 0, s>>>(idx->view, d_uids, d_hits, n, d_offsets, d_o, fill_cap);
      else occ_fill_kernel<false><<<grid, 256, 0, s>>>(idx->view, d_uids, d_hits, n, d_offsets, d_o, fill_cap);
    }
    MZ_CUDA(cudaGetLastError());
  }
  if (mem == MAZU_MEM_HOST) {
    if (total) MZ_CUDA(cudaMemcpyAsync(out, d_o, total * 12, cudaMemcpyDeviceToHost, s));
    MZ_CUDA(cudaStreamSynchronize(s));
  }
}

// reads -> MappedRefPos by tiles (get_ref_pos_pass1_kernel -> scan over tiles -> get_ref_pos_pass2_kernel); device pointers, work on `s`
static void launch_get_ref_pos_reads(const mazu_index_t* idx, const u8* d_bases, const u64* d_read_offsets, u64 n_reads, u64 uniform_len,
                                     const u64* d_kmer_offsets, Hit* d_hits, u64* d_counts, u64 n_slots, u64* d_offsets, OccRec* d_out, u64 cap,
                                     u64* d_total, cudaStream_t s) {
  TileTable tt;
  make_tiles(idx, d_read_offsets, n_reads, uniform_len, s, tt);
  const u64 n_tiles = tt.n_tiles;
  if (n_tiles == 0) {
    MZ_CUDA(cudaMemsetAsync(d_offsets, 0, 8, s));
    if (d_total) MZ_CUDA(cudaMemsetAsync(d_total, 0, 8, s));
    return;
  }
  std::unique_ptr<PoolBuf> tmp_hits;
  if (!d_hits) {  // the caller does not want the K2UPos records: they still carry pass 1's answers to pass 2
    tmp_hits = std::make_unique<PoolBuf>(idx->pool, n_slots * 16 + 16, s);
    d_hits = (Hit*)tmp_hits->p;
  }
  PoolBuf totals(idx->pool, (n_tiles + 1) * 8, s), base(idx->pool, (n_tiles + 1) * 8, s);
  MZ_CUDA(cudaMemsetAsync((u64*)totals.p + n_tiles, 0, 8, s));
  TileMap tm{d_read_offsets, d_kmer_offsets, tt.seg ? (const u64*)tt.seg->p : nullptr, n_reads, uniform_len, n_tiles, 0};
  const bool ss = idx->view.k2u_kind == MAZU_K2U_SSHASH, boophf = idx->view.mphf.family == MPHF_FAMILY_BOOPHF;
#define MZ_GRP(K, F, KW)                                                                                                 \
  {                                                                                                                       \
    auto kern = get_ref_pos_pass1_kernel<K, F, KW>;                                                                       \
    int grid = grid_for(kern, QR_WARPS * 32, idx, QR_WARPS, n_tiles);                                                     \
    kern<<<grid, QR_WARPS * 32, 0, s>>>(idx->view, d_bases, tm, d_hits, (unsigned long long*)d_counts, (u64*)totals.p);   \
  }
#define MZ_GRP_SS(KW) MZ_GRP(MAZU_K2U_SSHASH, MPHF_FAMILY_NATIVE, KW)
  if (ss) {
    MZ_KW_DISPATCH(kw_code(idx), MZ_GRP_SS)
  } else if (idx->view.k2u_kind == MAZU_K2U_SAMPLED_PFHASH) MZ_GRP(MAZU_K2U_SAMPLED_PFHASH, MPHF_FAMILY_BOOPHF, 0)
  else if (boophf) MZ_GRP(MAZU_K2U_PFHASH, MPHF_FAMILY_BOOPHF, 0)
  else MZ_GRP(MAZU_K2U_PFHASH, MPHF_FAMILY_NATIVE, 0)
#undef MZ_GRP_SS
#undef MZ_GRP
  MZ_CUDA(cudaGetLastError());
  device_exclusive_scan((const u64*)totals.p, (u64*)base.p, n_tiles, idx->pool, s);
  const int grid2 = (int)std::max<u64>(1, std::min<u64>((n_tiles + 7) / 8, (u64)idx->sm_count * 8));
  get_ref_pos_pass2_kernel<<<grid2, 256, 0, s>>>(idx->view, tm, d_hits, (const u64*)base.p, n_slots, d_offsets, d_out, cap, d_total);
  MZ_CUDA(cudaGetLastError());
}

mazu_status_t mazu_b200_get_ref_pos_reads(const mazu_index_t* idx, const uint8_t* bases, const uint64_t* read_offsets, uint64_t n_reads,
                                          uint64_t uniform_read_len, int32_t mode, uint64_t n_slots, uint64_t* kmer_offsets, mazu_hit_t* out_hits,
                                          uint64_t* out_offsets, mazu_occ_t* out_mrps, uint64_t cap, uint64_t* out_total, uint64_t* counts,
                                          int32_t mem, void* stream) {
  return guarded([&] {
    if (!idx || !out_offsets || (n_reads && !bases) || (cap && !out_mrps)) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    if (idx->view.u2pos_kind == MAZU_U2POS_NONE) throw Error(MAZU_ERR_NO_U2POS, "index has no U2Pos table");
    if (mode != MAZU_MODE_RANDOM && mode != MAZU_MODE_STREAMING) throw Error(MAZU_ERR_INVALID_ARG, "unknown query mode");
    if (mem != MAZU_MEM_HOST && mem != MAZU_MEM_DEVICE) throw Error(MAZU_ERR_INVALID_ARG, "unknown mem mode");
    if (n_reads && !uniform_read_len && !read_offsets) throw Error(MAZU_ERR_INVALID_ARG, "read_offsets is required for ragged reads");
    const u32 k = idx->unitigs->k;
    DeviceGuard g(idx->device);
    if (mem == MAZU_MEM_HOST) {
      if (n_slots != mazu_b200_count_kmer_slots(idx, read_offsets, n_reads, uniform_read_len)) throw Error(MAZU_ERR_INVALID_ARG, "n_slots does not match the reads");
      StreamPair sp;
      cudaStream_t s = sp.s[0];
      const u64 nb = uniform_read_len ? n_reads * uniform_read_len : (n_reads ? read_offsets[n_reads] - read_offsets[0] : 0);
      PoolBuf d_bases(idx->pool, nb + 16, s), d_ro(idx->pool, (n_reads + 1) * 8, s), d_ko(idx->pool, (n_reads + 1) * 8, s),
          d_hits(idx->pool, out_hits ? n_slots * 16 + 16 : 16, s), d_offs(idx->pool, (n_slots + 1) * 8, s), d_out(idx->pool, cap * 12 + 16, s),
          d_cnt(idx->pool, 24, s), d_tot(idx->pool, 8, s);
      if (nb) MZ_CUDA(cudaMemcpyAsync(d_bases.p, bases + (uniform_read_len ? 0 : read_offsets[0]), nb, cudaMemcpyHostToDevice, s));
      std::vector<u64> rel;
      if (!uniform_read_len) {
        rel.resize(n_reads + 1);
        for (u64 i = 0; i <= n_reads; ++i) rel[i] = read_offsets[i] - read_offsets[0];
        MZ_CUDA(cudaMemcpyAsync(d_ro.p, rel.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, s));
      }
      MZ_CUDA(cudaMemsetAsync(d_cnt.p, 0, 24, s));
      u64 total = 0;
      mazu_status_t rc = mazu_b200_get_ref_pos_reads(idx, (const uint8_t*)d_bases.p, uniform_read_len ? nullptr : (const uint64_t*)d_ro.p, n_reads,
                                                     uniform_read_len, mode, n_slots, (uint64_t*)d_ko.p, out_hits ? (mazu_hit_t*)d_hits.p : nullptr,
                                                     (uint64_t*)d_offs.p, (mazu_occ_t*)d_out.p, cap, &total, (uint64_t*)d_cnt.p, MAZU_MEM_DEVICE, s);
      if (out_total) *out_total = total;
      if (rc != MAZU_OK && !(rc == MAZU_ERR_INVALID_ARG && total > cap)) throw Error(rc, g_err);
      MZ_CUDA(cudaMemcpyAsync(out_offsets, d_offs.p, (n_slots + 1) * 8, cudaMemcpyDeviceToHost, s));
      if (out_hits && n_slots) MZ_CUDA(cudaMemcpyAsync(out_hits, d_hits.p, n_slots * 16, cudaMemcpyDeviceToHost, s));
      if (kmer_offsets) {
        if (uniform_read_len) {
          const u64 per = uniform_read_len >= k ? uniform_read_len - k + 1 : 0;
          for (u64 r = 0; r <= n_reads; ++r) kmer_offsets[r] = r * per;
        } else {
          MZ_CUDA(cudaMemcpyAsync(kmer_offsets, d_ko.p, (n_reads + 1) * 8, cudaMemcpyDeviceToHost, s));
        }
      }
      if (counts) MZ_CUDA(cudaMemcpyAsync(counts, d_cnt.p, 24, cudaMemcpyDeviceToHost, s));
      if (total && total <= cap) MZ_CUDA(cudaMemcpyAsync(out_mrps, d_out.p, total * 12, cudaMemcpyDeviceToHost, s));
      MZ_CUDA(cudaStreamSynchronize(s));
      if (total > cap) throw Error(MAZU_ERR_INVALID_ARG, "output capacity too small: need " + std::to_string(total) + " records");
      return;
    }
    // ---- device buffers ----
    cudaStream_t s = (cudaStream_t)stream;
    u64* d_koffs = kmer_offsets;
    std::unique_ptr<PoolBuf> tmp_koffs, d_tot;
    if (!uniform_read_len && n_reads) {
      if (!d_koffs) {
        tmp_koffs = std::make_unique<PoolBuf>(idx->pool, (n_reads + 1) * 8, s);
        d_koffs = (u64*)tmp_koffs->p;
      }
      PoolBuf lens(idx->pool, (n_reads + 1) * 8, s);
      MZ_CUDA(cudaMemsetAsync(lens.p, 0, (n_reads + 1) * 8, s));
      kmer_counts_kernel<<<(int)std::min<u64>((n_reads + 255) / 256, 4096), 256, 0, s>>>(read_offsets, n_reads, k, (u64*)lens.p);
      MZ_CUDA(cudaGetLastError());
      device_exclusive_scan((const u64*)lens.p, d_koffs, n_reads, idx->pool, s);
    }
    if (n_reads == 0) {  // nothing to look up: an empty prefix
      MZ_CUDA(cudaMemsetAsync(out_offsets, 0, 8, s));
      if (out_total) *out_total = 0;
      return;
    }
    if (out_total) d_tot = std::make_unique<PoolBuf>(idx->pool, 8, s);
    if (mode == MAZU_MODE_STREAMING && !idx->kmers_unique) {
      // the cursor walk is sequential per read and its answers can differ from the lookups' (duplicated k-mers): unfused chain
      std::unique_ptr<PoolBuf> tmp_hits;
      Hit* d_hits = (Hit*)out_hits;
      if (!d_hits) {
        tmp_hits = std::make_unique<PoolBuf>(idx->pool, n_slots * 16 + 16, s);
        d_hits = (Hit*)tmp_hits->p;
      }
      launch_query_reads(idx, bases, read_offsets, n_reads, uniform_read_len, mode, d_koffs, d_hits, 0, counts, s);
      occ_driver(idx, nullptr, (const mazu_hit_t*)d_hits, n_slots, out_offsets, out_mrps, cap, out_total, MAZU_MEM_DEVICE, s);
      return;
    }
    launch_get_ref_pos_reads(idx, bases, read_offsets, n_reads, uniform_read_len, d_koffs, (Hit*)out_hits, counts, n_slots, out_offsets,
                             (OccRec*)out_mrps, cap, d_tot ? (u64*)d_tot->p : nullptr, s);
    if (out_total) {
      MZ_CUDA(cudaMemcpyAsync(out_total, d_tot->p, 8, cudaMemcpyDeviceToHost, s));
      MZ_CUDA(cudaStreamSynchronize(s));
      if (*out_total > cap) throw Error(MAZU_ERR_INVALID_ARG, "output capacity too small: need " + std::to_string(*out_total) + " records");
    }
  });
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

// Validate::validate_ckmers for a batch of records: lookups through the read kernels, then the check of the projected
// positions on the device (validate_reads_kernel).  Records are processed in chunks of ~64 M bases.
static void validate_reads_impl(const mazu_index_t* idx, const uint8_t* bases, const uint64_t* read_offsets, uint64_t n_reads, int32_t mode,
                                uint64_t counts[5]) {
  if (!idx || !counts || (n_reads && (!bases || !read_offsets))) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
  if (mode != MAZU_MODE_RANDOM && mode != MAZU_MODE_STREAMING) throw Error(MAZU_ERR_INVALID_ARG, "unknown query mode");
  if (idx->view.u2pos_kind == MAZU_U2POS_NONE) throw Error(MAZU_ERR_NO_U2POS, "validate: index has no U2Pos table");
  const u32 k = idx->unitigs->k;
  DeviceGuard g(idx->device);
  StreamPair sp;
  cudaStream_t s = sp.s[0];
  PoolBuf d_counts(idx->pool, 40, s);
  MZ_CUDA(cudaMemsetAsync(d_counts.p, 0, 40, s));
  const u64 CHUNK = 64ull << 20;
  std::vector<u64> rel;
  for (u64 r0 = 0; r0 < n_reads;) {
    u64 r1 = (u64)(std::upper_bound(read_offsets + r0 + 1, read_offsets + n_reads + 1, read_offsets[r0] + CHUNK) - read_offsets) - 1;
    r1 = std::max(r1, r0 + 1);
    const u64 nr = r1 - r0, b0 = read_offsets[r0], nb = read_offsets[r1] - b0;
    rel.resize(nr + 1);
    u64 slots = 0;
    for (u64 i = 0; i <= nr; ++i) {
      rel[i] = read_offsets[r0 + i] - b0;
      if (i < nr) {
        const u64 len = read_offsets[r0 + i + 1] - read_offsets[r0 + i];
        if (len >= k) slots += len - k + 1;
      }
    }
    PoolBuf d_bases(idx->pool, nb + 16, s), d_ro(idx->pool, (nr + 1) * 8, s), d_ko(idx->pool, (nr + 1) * 8, s), d_hits(idx->pool, slots * 16 + 16, s);
    MZ_CUDA(cudaMemcpyAsync(d_bases.p, bases + b0, nb, cudaMemcpyHostToDevice, s));
    MZ_CUDA(cudaMemcpyAsync(d_ro.p, rel.data(), (nr + 1) * 8, cudaMemcpyHostToDevice, s));
    mazu_status_t rc = mazu_b200_query_reads(idx, (const uint8_t*)d_bases.p, (const uint64_t*)d_ro.p, nr, 0, mode, (uint64_t*)d_ko.p, (mazu_hit_t*)d_hits.p,
                                             nullptr, MAZU_MEM_DEVICE, s);
    if (rc != MAZU_OK) throw Error(rc, g_err);
    if (slots) {
      const int grid = (int)std::max<u64>(1, std::min<u64>((slots + 255) / 256, (u64)idx->sm_count * 8));
      validate_reads_kernel<<<grid, 256, 0, s>>>(idx->view, (const Hit*)d_hits.p, (const u64*)d_ko.p, nr, r0, (unsigned long long*)d_counts.p);
      MZ_CUDA(cudaGetLastError());
    }
    MZ_CUDA(cudaStreamSynchronize(s));  // `rel` is reused by the next chunk
    r0 = r1;
  }
  MZ_CUDA(cudaMemcpyAsync(counts, d_counts.p, 40, cudaMemcpyDeviceToHost, s));
  MZ_CUDA(cudaStreamSynchronize(s));
}

struct mazu_fasta {
  FastaHost h;
};

mazu_status_t mazu_b200_fasta_open(const char* path, mazu_fasta_t** out) {
  return guarded([&] {
    if (!path || !out) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    auto f = std::make_unique<mazu_fasta>();
    f->h = read_fasta_file(path);
    *out = f.release();
  });
}
void mazu_b200_fasta_close(mazu_fasta_t* f) { delete f; }
uint64_t mazu_b200_fasta_n_records(const mazu_fasta_t* f) { return f ? f->h.n_records() : 0; }
const uint8_t* mazu_b200_fasta_bases(const mazu_fasta_t* f) { return f ? f->h.bases.data() : nullptr; }
const uint64_t* mazu_b200_fasta_offsets(const mazu_fasta_t* f) { return f ? f->h.offsets.data() : nullptr; }
const char* mazu_b200_fasta_name(const mazu_fasta_t* f, uint64_t record) { return f && record < f->h.names.size() ? f->h.names[record].c_str() : nullptr; }

mazu_status_t mazu_b200_validate_reads(const mazu_index_t* idx, const uint8_t* bases, const uint64_t* read_offsets, uint64_t n_reads, int32_t mode,
                                       uint64_t counts[5]) {
  return guarded([&] { validate_reads_impl(idx, bases, read_offsets, n_reads, mode, counts); });
}
mazu_status_t mazu_b200_validate_fasta(const mazu_index_t* idx, const char* path, int32_t mode, uint64_t counts[5]) {
  return guarded([&] {
    if (!path) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    FastaHost f = read_fasta_file(path);
    validate_reads_impl(idx, f.bases.data(), f.offsets.data(), f.n_records(), mode, counts);
  });
}

mazu_status_t mazu_b200_unitig_seq(const mazu_index_t* idx, uint64_t unitig_id, uint64_t* out_words, uint64_t cap_words, uint64_t* len) {
  return guarded([&] {
    if (!idx || !len) throw Error(MAZU_ERR_INVALID_ARG, "null argument");
    const UnitigSetHost& us = *idx->unitigs;
    if (unitig_id >= us.n_unitigs()) throw Error(MAZU_ERR_INVALID_ARG, "unitig id out of range");
    const u64 s0 = us.accum[unitig_id], n = us.accum[unitig_id + 1] - s0, nw = (n + 31) / 32;
    *len = n;
    if (!out_words) return;
    if (cap_words < nw) throw Error(MAZU_ERR_INVALID_ARG, "output capacity too small: need " + std::to_string(nw) + " words");
    for (u64 j = 0; j < nw; ++j) {
      const u64 bit = 2 * (s0 + 32 * j), wi = bit >> 6, sh = bit & 63;
      u64 x = us.useq[wi] >> sh;
      if (sh) x |= us.useq[wi + 1] << (64 - sh);
      const u64 left = n - 32 * j;
      out_words[j] = left >= 32 ? x : (x & ((1ULL << (2 * left)) - 1ULL));
    }
  });
}

mazu_status_t mazu_b200_validate_self(const mazu_index_t* idx, uint64_t counts[5]) {
  return guarded([&] { run_validate(idx, false, counts); });
}
mazu_status_t mazu_b200_k2u_validate_self(const mazu_index_t* idx, uint64_t counts[5]) {
  return guarded([&] { run_validate(idx, true, counts); });
}

}  // extern "C"

namespace {
struct ProbeTable {  // kept between calls of the same size: a sweep point then costs its launches only
  void* p = nullptr;
  u64 bytes = 0;
  int device = -1;
  void release() {
    if (p) {
      cudaSetDevice(device);
      cudaFree(p);
    }
    p = nullptr;
    bytes = 0;
  }
};
ProbeTable g_probe;
std::mutex g_probe_mu;

template <int LANES>
void launch_gather_probe(int ilp, int grid, size_t smem, const uint4* table, u64 n_granules, u64 n_items, u64 seed, unsigned long long* sink) {
#define MZ_GP(I)                                                                                              \
  {                                                                                                           \
    MZ_CUDA(cudaFuncSetAttribute(gather_probe_kernel<LANES, I>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    gather_probe_kernel<LANES, I><<<grid, 256, smem>>>(table, n_granules, n_items, seed, sink);              \
  }
  switch (ilp) {
    case 1: MZ_GP(1) break;
    case 2: MZ_GP(2) break;
    case 4: MZ_GP(4) break;
    case 8: MZ_GP(8) break;
    default: throw Error(MAZU_ERR_INVALID_ARG, "ilp must be 1, 2, 4 or 8");
  }
#undef MZ_GP
}
}  // namespace

extern "C" {

mazu_status_t mazu_b200_debug_gather_probe(uint64_t table_bytes, uint64_t n_items, int32_t granule_bytes, int32_t ilp, int32_t blocks_per_sm,
                                           int32_t iters, int32_t device, double* items_per_s) {
  return guarded([&] {
    std::lock_guard<std::mutex> lk(g_probe_mu);
    if (table_bytes == 0) {
      g_probe.release();
      return;
    }
    if (!items_per_s || table_bytes < 4096 || blocks_per_sm < 1 || blocks_per_sm > 8) throw Error(MAZU_ERR_INVALID_ARG, "bad argument");
    DeviceGuard g(device);
    cudaDeviceProp prop;
    MZ_CUDA(cudaGetDeviceProperties(&prop, device));
    if (g_probe.bytes != table_bytes || g_probe.device != device) {
      g_probe.release();
      MZ_CUDA(cudaMalloc(&g_probe.p, table_bytes));
      g_probe.bytes = table_bytes;
      g_probe.device = device;
      MZ_CUDA(cudaMemset(g_probe.p, 1, table_bytes));
    }
    DevBuf sink(8, device);
    MZ_CUDA(cudaMemset(sink.p, 0, 8));
    // dynamic shared memory as ballast: exactly blocks_per_sm CTAs fit an SM
    const size_t smem_total = prop.sharedMemPerMultiprocessor, reserved = prop.reservedSharedMemPerBlock;
    size_t smem = blocks_per_sm >= 8 ? 0 : smem_total / blocks_per_sm - reserved - 256;
    smem = std::min<size_t>(smem, prop.sharedMemPerBlockOptin);
    const int grid = prop.multiProcessorCount * blocks_per_sm;
    const u64 n_granules = table_bytes / (u64)granule_bytes;
    cudaEvent_t e0, e1;
    MZ_CUDA(cudaEventCreate(&e0));
    MZ_CUDA(cudaEventCreate(&e1));
    double best = 0;
    for (int it = 0; it < iters + 1; ++it) {
      MZ_CUDA(cudaEventRecord(e0));
      const uint4* t = (const uint4*)g_probe.p;
      unsigned long long* sk = (unsigned long long*)sink.p;
      switch (granule_bytes) {
        case 16: launch_gather_probe<1>(ilp, grid, smem, t, n_granules, n_items, 0x1234 + it, sk); break;
        case 32: launch_gather_probe<2>(ilp, grid, smem, t, n_granules, n_items, 0x1234 + it, sk); break;
        case 64: launch_gather_probe<4>(ilp, grid, smem, t, n_granules, n_items, 0x1234 + it, sk); break;
        case 128: launch_gather_probe<8>(ilp, grid, smem, t, n_granules, n_items, 0x1234 + it, sk); break;
        default: throw Error(MAZU_ERR_INVALID_ARG, "granule_bytes must be 16, 32, 64 or 128");
      }
      MZ_CUDA(cudaGetLastError());
      MZ_CUDA(cudaEventRecord(e1));
      MZ_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      MZ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (it > 0) best = std::max(best, (double)n_items / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *items_per_s = best;
  });
}

}  // extern "C"
