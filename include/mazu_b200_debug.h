/*
 * mazu_b200_debug.h -- test and measurement hooks of libmazu_b200.so.  NOT part of the drop-in boundary
 * (include/mazu_b200.h): nothing here stands for a mazu interface; the parity tests and profiles/ scripts use them.
 */
#ifndef MAZU_B200_DEBUG_H
#define MAZU_B200_DEBUG_H

#include "mazu_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* test hook: FNV-1a digest and logical size of one device table (0 MPHF blocks, 1 bucket-bound blocks, 2 their exceptions,
 * 3 packed positions, 4 skew MPHF blocks, 5 skew positions, 6 MPHF fallback keys, 7 unitig lines) */
mazu_status_t mazu_b200_debug_table_digest(const mazu_index_t* idx, int32_t which, uint64_t* digest, uint64_t* n_bytes);

/* measurement hook: level-0 MPHF block of every query's key (the minimizer for SSHash, the canonical k-mer for PFHash);
 * device pointers.  Sorting a flat batch by this key makes its MPHF / bounds / positions accesses sequential: used by
 * profiles/sorted_probe_experiment.py to bound what "sorted probe batches" could gain. */
mazu_status_t mazu_b200_debug_probe_key(const mazu_index_t* idx, const uint64_t* fw_words, uint64_t n, uint32_t* out_block, void* stream);

/* Random-access roofline probe (the P_rand denominator of BASELINE.md section 2), one point of the sweep:
 * n_items independent random reads of one aligned granule of granule_bytes (16, 32, 64 or 128) each from a table of
 * table_bytes (allocated internally, kept between calls of the same size; table_bytes = 0 frees it); a granule is read
 * by granule_bytes / 16 adjacent lanes with one 16-byte load each; `ilp` (1, 2, 4, 8) independent granules are in flight
 * per thread; blocks_per_sm (1..8) CTAs of 256 threads are resident per SM.  Best of `iters` launches, in granules per
 * second.  profiles/measure_prand.py sweeps it; DRAM bytes per granule come from an ncu capture of the same launch. */
mazu_status_t mazu_b200_debug_gather_probe(uint64_t table_bytes, uint64_t n_items, int32_t granule_bytes, int32_t ilp,
                                           int32_t blocks_per_sm, int32_t iters, int32_t device, double* items_per_s);

#ifdef __cplusplus
}
#endif
#endif /* MAZU_B200_DEBUG_H */
