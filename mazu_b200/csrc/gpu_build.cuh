// SSHashBuilder::from_unitig_set + finish (src/kphf/sshash.rs:86-329) on the device.
//
// Produces, directly in HBM and bit-identical to the host builder of host_build.hpp, the tables the
// lookup kernels read: native MPHF blocks, blocked Elias-Fano bucket bounds (+ fingerprints), packed
// minimizer positions, and the skew index.  Stages (each a kernel or a cub primitive):
//   1 collect   warp per unitig: k-mer words from the packed sequence, w-mer hash keys, window minimum,
//               the reference's two streams (fw-canonical k-mers then rc-canonical k-mers) with run-length
//               dedup; two passes (count, scan, fill) so tuples land in the reference's order
//   2 sort      cub::DeviceRadixSort::SortPairs on the minimizer word (stable, like rayon par_sort_by_key)
//   3 group     heads -> scan -> distinct minimizers + ranges
//   4 mphf      BBHash cascade with atomicOr on the final block layout; colliders filtered to the next level
//   5 sizes     MPHF value of every minimizer -> bucket sizes -> exclusive scan; fingerprints
//   6 scatter   positions into bucket order
//   7 skew      k-mers around the occurrences of heavy buckets -> sort -> first of each run -> second MPHF
//   8 encode    blocked Elias-Fano blocks, exception blocks compacted by a scan
//   9 pack      positions bit-packed, one thread per output word
// Included by capi.cu after its helpers (DevBuf, MZ_CUDA).
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_scan.cuh>

namespace mazu {

struct CollectStage {
  u64 fw[QR_CHUNK];
  u32 hf[QR_BASES];
  u32 hr[QR_BASES];
};

// stage 1.  WRITE = false: counts[2u], counts[2u+1] = tuples of the fw / rc stream of unitig u.
//           WRITE = true : tuples written at bases[2u], bases[2u+1] (exclusive scan of the counts).
template <bool WRITE>
__global__ void __launch_bounds__(QR_WARPS * 32) collect_minimizers_kernel(const __grid_constant__ UnitigsView uv, u32 w, u64 seed,
                                                                           u64* __restrict__ counts, const u64* __restrict__ bases,
                                                                           u64* __restrict__ out_word, u64* __restrict__ out_pos) {
  __shared__ CollectStage s_stage[QR_WARPS];
  const u32 lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  CollectStage& S = s_stage[wib];
  const u32 k = uv.k, span = k - w;
  const u64 wmask = kmer_mask(w);
  const u32 lt_mask = (1u << lane) - 1u;
  for (u64 ui = (u64)blockIdx.x * QR_WARPS + wib; ui < uv.n_unitigs; ui += (u64)gridDim.x * QR_WARPS) {
    const u64 s = __ldg(uv.starts + ui), e = __ldg(uv.starts + ui + 1);
    const u64 nk = e - s >= k ? e - s - k + 1 : 0;
    u32 have[2] = {0, 0};
    u64 pw[2] = {0, 0}, pp[2] = {0, 0}, cnt[2] = {0, 0};
    u64 base[2] = {0, 0};
    if (WRITE) {
      base[0] = __ldg(bases + 2 * ui);
      base[1] = __ldg(bases + 2 * ui + 1);
    }
    for (u64 c0 = 0; c0 < nk; c0 += QR_CHUNK) {
      const u32 n_c = (u32)min((u64)QR_CHUNK, nk - c0);
      __syncwarp();
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        u64 p = s + c0 + 32u * t + lane;
        S.fw[32 * t + lane] = p <= uv.total_len ? useq_window(uv, p) : 0ULL;
      }
      __syncwarp();
#pragma unroll 1
      for (int t = 0; t < 5; ++t) {
        u32 q = 32 * t + lane;
        u64 x;
        if (q < QR_CHUNK) x = S.fw[q];
        else {
          u32 d = q - (QR_CHUNK - 1);
          x = d < 32 ? (S.fw[QR_CHUNK - 1] >> (2 * d)) : 0ULL;
        }
        u64 wf = x & wmask;
        const u32 key = mm_key(wf, revcomp(wf, w), seed);
        S.hf[q] = key;
        S.hr[QR_BASES - 1 - q] = key;
      }
      __syncwarp();
#pragma unroll 1
      for (int t = 0; t < 4; ++t) {
        const u32 p = 32 * t + lane;
        const bool valid = p < n_c;
        u64 mmw = 0, apos = 0;
        u32 stream = 0;
        if (valid) {
          u64 fw = S.fw[p], rc = revcomp(fw, k);
          bool fw_canon = fw <= rc;
          const u32* h = fw_canon ? S.hf + p : S.hr + (QR_BASES - 1 - p - span);
          u32 bi = window_min(h, span) & 31u;
          const u32 off_fw = fw_canon ? bi : span - bi;
          mmw = mm_word_of(fw, rc, off_fw, k, w);
          apos = s + c0 + p + off_fw;  // occurrence position on the forward strand
          stream = fw_canon ? 0u : 1u;
        }
#pragma unroll
        for (u32 st = 0; st < 2; ++st) {
          const bool in_s = valid && stream == st;
          const u32 mask = __ballot_sync(0xffffffffu, in_s);
          const u32 lower = mask & lt_mask;
          const int j = lower ? 31 - __clz(lower) : 0;
          u64 prev_w = __shfl_sync(0xffffffffu, mmw, j), prev_p = __shfl_sync(0xffffffffu, apos, j);
          u32 prev_have = 1;
          if (!lower) {
            prev_w = pw[st];
            prev_p = pp[st];
            prev_have = have[st];
          }
          const bool emit = in_s && (!prev_have || prev_w != mmw || prev_p != apos);  // sshash.rs:109-117 / :126-134
          const u32 emask = __ballot_sync(0xffffffffu, emit);
          if (WRITE && emit) {
            u64 o = base[st] + cnt[st] + __popc(emask & lt_mask);
            out_word[o] = mmw;
            out_pos[o] = apos;
          }
          cnt[st] += __popc(emask);
          if (mask) {
            const int last = 31 - __clz(mask);
            pw[st] = __shfl_sync(0xffffffffu, mmw, last);
            pp[st] = __shfl_sync(0xffffffffu, apos, last);
            have[st] = 1;
          }
        }
      }
    }
    if (!WRITE && lane == 0) {
      counts[2 * ui] = cnt[0];
      counts[2 * ui + 1] = cnt[1];
    }
  }
}

// stage 3
__global__ void mark_heads_kernel(const u64* __restrict__ keys, u64 n, u64* __restrict__ flags) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
    flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1ULL : 0ULL;
}
// stage 3b (host twin: build_sshash step 2b): an entry equal to its predecessor -- the same (word, position) pushed by both of
// the reference's streams -- is dropped unless its bucket goes to the skew index (more than skew_param entries)
__global__ void mark_repeats_kernel(const u64* __restrict__ words, const u64* __restrict__ poss, const u64* __restrict__ gid,
                                    const u64* __restrict__ ranges, u64 n, u64 skew_param, u64* __restrict__ keep) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    bool rep = i > 0 && words[i] == words[i - 1] && poss[i] == poss[i - 1];
    if (rep) {
      const u64 g = gid[i] - 1;
      if (ranges[g + 1] - ranges[g] > skew_param) rep = false;
    }
    keep[i] = rep ? 0ULL : 1ULL;
  }
}
__global__ void compact_pairs_kernel(const u64* __restrict__ words, const u64* __restrict__ poss, const u64* __restrict__ keep,
                                     const u64* __restrict__ kidx, u64 n, u64* __restrict__ words_out, u64* __restrict__ poss_out) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
    if (keep[i]) {
      const u64 o = kidx[i] - 1;  // kidx = inclusive scan of keep
      words_out[o] = words[i];
      poss_out[o] = poss[i];
    }
}
// gid = inclusive scan of flags (1-based group number); heads write their key / first index
__global__ void scatter_groups_kernel(const u64* __restrict__ keys, const u64* __restrict__ vals, const u64* __restrict__ flags,
                                      const u64* __restrict__ gid, u64 n, u64* __restrict__ set, u64* __restrict__ first_val,
                                      u64* __restrict__ ranges) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
    if (flags[i]) {
      u64 g = gid[i] - 1;
      set[g] = keys[i];
      if (first_val) first_val[g] = vals[i];
      ranges[g] = i;
    }
}

// stage 4
__global__ void mphf_mark_kernel(const u64* __restrict__ keys, u64 n, u32 level, u64 nb, u32* __restrict__ seen, u32* __restrict__ coll) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    u64 blk;
    u32 bit;
    native_slot(keys[i], level, nb, blk, bit);
    u64 wi = blk * 8 + 1 + (bit >> 5);
    u32 m = 1u << (bit & 31);
    u32 old = atomicOr(seen + wi, m);
    if (old & m) atomicOr(coll + wi, m);
  }
}
__global__ void mphf_finalize_kernel(u32* __restrict__ seen, const u32* __restrict__ coll, u64 nb, u32* __restrict__ ones) {
  for (u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (u64)gridDim.x * blockDim.x) {
    u32 c = 0;
    for (int j = 1; j < 8; ++j) {
      u32 x = seen[b * 8 + j] & ~coll[b * 8 + j];
      seen[b * 8 + j] = x;
      c += __popc(x);
    }
    ones[b] = c;
  }
}
__global__ void mphf_set_ranks_kernel(u32* __restrict__ blocks, const u32* __restrict__ prefix, u64 nb) {
  for (u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (u64)gridDim.x * blockDim.x) blocks[b * 8] = prefix[b];
}
__global__ void mphf_filter_kernel(const u64* __restrict__ keys, u64 n, u32 level, u64 nb, const u32* __restrict__ coll,
                                   u64* __restrict__ next, unsigned long long* __restrict__ next_n) {
  for (u64 i0 = (u64)blockIdx.x * blockDim.x; i0 < n; i0 += (u64)gridDim.x * blockDim.x) {
    u64 i = i0 + threadIdx.x;
    bool keep = false;
    u64 key = 0;
    if (i < n) {
      key = keys[i];
      u64 blk;
      u32 bit;
      native_slot(key, level, nb, blk, bit);
      keep = (coll[blk * 8 + 1 + (bit >> 5)] >> (bit & 31)) & 1u;
    }
    u32 m = __ballot_sync(0xffffffffu, keep);
    u64 basei = 0;
    if ((threadIdx.x & 31) == 0 && m) basei = atomicAdd(next_n, (unsigned long long)__popc(m));
    basei = __shfl_sync(0xffffffffu, basei, 0);
    if (keep) next[basei + __popc(m & ((1u << (threadIdx.x & 31)) - 1u))] = key;
  }
}

// stage 4': the fingerprinted cascade (index_layout.hpp, MPHF_FAMILY_CASCADE; host twin: MphfHost::build_cascade).
// seen / coll are plain bitmaps of `size` bits; keys carry the index of their group so the slot can be recorded per group.
__global__ void cascade_mark_kernel(const u64* __restrict__ keys, u64 n, u32 level, u64 size, u32* __restrict__ seen, u32* __restrict__ coll) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    const u64 sl = cascade_slot(fmix64(keys[i]), level, size);
    const u32 m = 1u << (sl & 31);
    const u32 old = atomicOr(seen + (sl >> 5), m);
    if (old & m) atomicOr(coll + (sl >> 5), m);
  }
}
// keys alone in their slot are placed (state = fingerprint, slot recorded); colliders mark the slot and move on to the next level
__global__ void cascade_place_kernel(const u64* __restrict__ keys, const u64* __restrict__ idx, u64 n, u32 level, u64 size, u64 off,
                                     const u32* __restrict__ coll, u8* __restrict__ states, u64* __restrict__ slots, u64* __restrict__ next_keys,
                                     u64* __restrict__ next_idx, unsigned long long* __restrict__ next_n) {
  for (u64 i0 = (u64)blockIdx.x * blockDim.x; i0 < n; i0 += (u64)gridDim.x * blockDim.x) {
    const u64 i = i0 + threadIdx.x;
    bool keep = false;
    u64 key = 0, id = 0;
    if (i < n) {
      key = keys[i];
      id = idx ? idx[i] : i;
      const u64 hk = fmix64(key), sl = cascade_slot(hk, level, size);
      keep = (coll[sl >> 5] >> (sl & 31)) & 1u;
      states[sl] = keep ? (u8)CASCADE_COLLIDED : (u8)cascade_fp(hk);  // every collider of a slot writes the same byte
      if (!keep) slots[id] = off + sl;
    }
    const u32 m = __ballot_sync(0xffffffffu, keep);
    u64 basei = 0;
    if ((threadIdx.x & 31) == 0 && m) basei = atomicAdd(next_n, (unsigned long long)__popc(m));
    basei = __shfl_sync(0xffffffffu, basei, 0);
    if (keep) {
      const u64 o = basei + __popc(m & ((1u << (threadIdx.x & 31)) - 1u));
      next_keys[o] = key;
      next_idx[o] = id;
    }
  }
}
__global__ void cascade_sizes_kernel(const u64* __restrict__ slots, const u64* __restrict__ ranges, u64 M, u64 R, u64* __restrict__ sizes_by_slot,
                                     unsigned long long* __restrict__ bad) {
  for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < M; g += (u64)gridDim.x * blockDim.x) {
    const u64 h = slots[g];
    if (h >= R) {
      atomicAdd(bad, 1ULL);
      continue;
    }
    sizes_by_slot[h] = ranges[g + 1] - ranges[g];
  }
}

// stage 5
__global__ void group_hash_kernel(const __grid_constant__ RankedLevels m, const u64* __restrict__ set, const u64* __restrict__ ranges, u64 M,
                                  u64* __restrict__ hashes, u64* __restrict__ sizes_by_h, unsigned long long* __restrict__ bad) {
  for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < M; g += (u64)gridDim.x * blockDim.x) {
    u64 h;
    if (!mphf_lookup_t<MPHF_FAMILY_NATIVE>(m, set[g], h) || h >= M) {
      atomicAdd(bad, 1ULL);
      continue;
    }
    hashes[g] = h;
    if (sizes_by_h) sizes_by_h[h] = ranges[g + 1] - ranges[g];
  }
}
// stage 6: tuple j of group g goes to prefix[hash(g)] + (j - ranges[g])
__global__ void scatter_positions_kernel(const u64* __restrict__ vals, const u64* __restrict__ gid, const u64* __restrict__ ranges,
                                         const u64* __restrict__ hashes, const u64* __restrict__ prefix, u64 n, u64* __restrict__ out) {
  for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (u64)gridDim.x * blockDim.x) {
    u64 g = gid[j] - 1;
    out[prefix[hashes[g]] + (j - ranges[g])] = vals[j];
  }
}
// skew: scatter position of the first tuple of every distinct k-mer to its MPHF slot
__global__ void scatter_by_hash_kernel(const u64* __restrict__ first_val, const u64* __restrict__ hashes, u64 M, u64* __restrict__ out) {
  for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < M; g += (u64)gridDim.x * blockDim.x) out[hashes[g]] = first_val[g];
}

// stage 7: k-mers overlapping the occurrences of heavy buckets (sshash.rs:231-262)
__device__ __forceinline__ bool is_valid_useq_pos_dev(const UnitigsView& uv, u64 p) {  // unitig_set.rs:235-245
  if (uv.total_len < uv.k || p > uv.total_len - uv.k) return false;
  u64 id, s, e;
  unitig_locate(uv, p, id, s, e);
  return p + uv.k <= e;
}
template <bool WRITE>
__global__ void skew_tuples_kernel(const __grid_constant__ UnitigsView uv, u32 w, u64 skew_param, const u64* __restrict__ pos_sorted,
                                   const u64* __restrict__ gid, const u64* __restrict__ ranges, u64 n, u64* __restrict__ counts,
                                   const u64* __restrict__ bases, u64* __restrict__ out_word, u64* __restrict__ out_pos) {
  const u32 k = uv.k;
  for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (u64)gridDim.x * blockDim.x) {
    u64 g = gid[j] - 1;
    u64 c = 0;
    if (ranges[g + 1] - ranges[g] > skew_param) {
      u64 mp = pos_sorted[j];
      u64 start = mp < (u64)(k - w) ? 0 : mp - (u64)(k - w);  // sshash.rs:241-247
      u64 o = WRITE ? bases[j] : 0;
      for (u32 off = 0; off <= k - w; ++off) {
        u64 p = start + off;
        if (is_valid_useq_pos_dev(uv, p)) {
          if (WRITE) {
            u64 fw = useq_window(uv, p), rc = revcomp(fw, k);
            out_word[o + c] = fw <= rc ? fw : rc;
            out_pos[o + c] = p;
          }
          ++c;
        }
      }
    }
    if (!WRITE) counts[j] = c;
  }
}

// PFHash::from_unitig_set (src/kphf/pfhash.rs:40-73): canonical k-mer of every valid start position, in unitig order
// (k-mer index of position p inside unitig ui = p - ui * (k - 1), every unitig being at least k long)
__global__ void pfhash_keys_kernel(const __grid_constant__ UnitigsView uv, u64* __restrict__ keys, u64* __restrict__ positions) {
  const u32 k = uv.k;
  for (u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x; p + k <= uv.total_len; p += (u64)gridDim.x * blockDim.x) {
    u64 id, s, e;
    unitig_locate(uv, p, id, s, e);
    if (p + k > e) continue;
    u64 fw = useq_window(uv, p), rc = revcomp(fw, k);
    u64 i = p - id * (u64)(k - 1);
    keys[i] = fw <= rc ? fw : rc;
    positions[i] = p;
  }
}
// pos[h(kmer)] = position (pfhash.rs:60-68)
__global__ void pfhash_scatter_kernel(const __grid_constant__ RankedLevels m, const u64* __restrict__ keys, const u64* __restrict__ positions, u64 n,
                                      u64 n_slots, u64* __restrict__ out, unsigned long long* __restrict__ bad) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    u64 h;
    if (!mphf_lookup_t<MPHF_FAMILY_NATIVE>(m, keys[i], h) || h >= n_slots) {
      atomicAdd(bad, 1ULL);
      continue;
    }
    out[h] = positions[i];
  }
}

// stage 8: blocked Elias-Fano (host twin: BlockedEF::build)
template <bool WRITE>
__global__ void ef_blocks_kernel(const u64* __restrict__ xs, u64 n, u32 l, u32 log_s, u32 wpb, bool all_exc, const u8* __restrict__ fps, u64 n_fps,
                                 u64* __restrict__ exc_flags, const u64* __restrict__ exc_index, u64* __restrict__ blocks,
                                 u64* __restrict__ exceptions) {
  const u64 S = 1ULL << log_s, nb = (n + S - 1) / S;
  for (u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (u64)gridDim.x * blockDim.x) {
    const u64 i0 = b * S, cnt = min(S + 1, n - i0);
    const u64 hb = xs[i0] >> l;
    const bool exc = all_exc || ((xs[i0 + cnt - 1] >> l) - hb) + (cnt - 1) >= 128;
    if (!WRITE) {
      exc_flags[b] = exc ? 1ULL : 0ULL;
      continue;
    }
    u64 w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (fps)
      for (u64 j = 0; j < S && i0 + j < n_fps; ++j) w[4 + (j >> 3)] |= (u64)fps[i0 + j] << (8 * (j & 7));
    if (exc) {
      const u64 ei = exc_index[b];
      w[0] = (1ULL << 63) | ei;
      for (u64 j = 0; j <= S; ++j) exceptions[ei * (S + 1) + j] = j < cnt ? xs[i0 + j] : xs[i0 + cnt - 1];
    } else {
      w[0] = xs[i0];
      const u64 lmask = (1ULL << l) - 1;
      for (u64 j = 0; j < cnt; ++j) {
        u64 x = xs[i0 + j];
        u64 p = ((x >> l) - hb) + j;
        w[1 + (p >> 6)] |= 1ULL << (p & 63);
        w[3] |= (x & lmask) << (j * l);
      }
    }
    for (u32 j = 0; j < wpb; ++j) blocks[b * wpb + j] = w[j];
  }
}

// stage 9: bit-pack `width`-bit fields, one thread per output word (no atomics)
__global__ void pack_kernel(const u64* __restrict__ vals, u64 n, u32 width, u64* __restrict__ words, u64 n_words) {
  for (u64 wi = (u64)blockIdx.x * blockDim.x + threadIdx.x; wi < n_words; wi += (u64)gridDim.x * blockDim.x) {
    const u64 bit0 = wi * 64;
    u64 i = bit0 / width;  // first field overlapping this word
    u64 acc = 0;
    for (; i < n; ++i) {
      u64 fb = i * width;
      if (fb >= bit0 + 64) break;
      u64 v = vals[i] & (width >= 64 ? ~0ULL : ((1ULL << width) - 1ULL));
      if (fb >= bit0) acc |= v << (fb - bit0);
      else acc |= v >> (bit0 - fb);
    }
    words[wi] = acc;
  }
}

}  // namespace mazu
