/* Minimal C caller of libmazu_b200.so: load a pufferfish dense index, rebuild its K2U as an SSHash
 * (src/pf1/dense_index.rs:315-328), query a few reads in random-access and streaming mode, project the hits.
 *   gcc -std=c99 -Iinclude examples/query_reads.c -Lmazu_b200 -lmazu_b200 -Wl,-rpath,$PWD/mazu_b200 -o query_reads
 *   ./query_reads tests/data/pf1/yeast_chr01_index */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mazu_b200.h"

#define CHECK(call)                                                                   \
  do {                                                                                \
    mazu_status_t rc_ = (call);                                                       \
    if (rc_ != MAZU_OK) {                                                             \
      fprintf(stderr, "%s failed (%d): %s\n", #call, (int)rc_, mazu_b200_last_error()); \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s <pufferfish index dir>\n", argv[0]);
    return 2;
  }
  if (mazu_b200_device_count() <= 0) {
    fprintf(stderr, "no CUDA device: libmazu_b200 has no CPU fallback\n");
    return 3;
  }
  mazu_index_t *dense = NULL, *idx = NULL;
  CHECK(mazu_b200_dense_index_deserialize_from_cpp(argv[1], 0, &dense));
  CHECK(mazu_b200_index_rebuild_k2u(dense, MAZU_K2U_SSHASH, 15, 32, 0, &idx));
  uint64_t counts5[5];
  CHECK(mazu_b200_validate_self(idx, counts5));
  printf("validate_self: %llu queries, %llu identity, %llu twin, %llu projected, %llu failures\n", (unsigned long long)counts5[0],
         (unsigned long long)counts5[1], (unsigned long long)counts5[2], (unsigned long long)counts5[3], (unsigned long long)counts5[4]);

  /* two reads of 62 bases, one with an N */
  const char* r0 = "ACGTTGCAAGGCTTAACCGGTTAAGGCCTTAACCGGATATCGCGATATCGCGAATTCCGGAA";
  const char* r1 = "TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTNTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT";
  uint64_t offs[3] = {0, 0, 0};
  offs[1] = strlen(r0);
  offs[2] = offs[1] + strlen(r1);
  uint8_t* bases = NULL;
  CHECK(mazu_b200_alloc_pinned(offs[2], (void**)&bases)); /* pinned: copies overlap with the kernels */
  memcpy(bases, r0, offs[1]);
  memcpy(bases + offs[1], r1, offs[2] - offs[1]);
  uint64_t n_slots = mazu_b200_count_kmer_slots(idx, offs, 2, 0);
  mazu_hit_t* hits = NULL;
  CHECK(mazu_b200_alloc_pinned(n_slots * sizeof(mazu_hit_t), (void**)&hits));
  uint64_t koffs[3], counts[3];
  for (int mode = MAZU_MODE_RANDOM; mode <= MAZU_MODE_STREAMING; ++mode) {
    CHECK(mazu_b200_query_reads(idx, bases, offs, 2, 0, mode, koffs, hits, counts, MAZU_MEM_HOST, NULL));
    printf("mode %d: %llu k-mers, %llu hits, %llu misses, %llu slots\n", mode, (unsigned long long)counts[0], (unsigned long long)counts[1],
           (unsigned long long)counts[2], (unsigned long long)n_slots);
  }
  /* GetRefPos::project_hits: sizes first, then the records */
  uint64_t* occ_offs = (uint64_t*)malloc((n_slots + 1) * sizeof(uint64_t));
  uint64_t total = 0;
  CHECK(mazu_b200_project_hits(idx, hits, n_slots, occ_offs, NULL, 0, &total, MAZU_MEM_HOST, NULL));
  mazu_occ_t* occs = (mazu_occ_t*)malloc((total ? total : 1) * sizeof(mazu_occ_t));
  if (total) CHECK(mazu_b200_project_hits(idx, hits, n_slots, occ_offs, occs, total, &total, MAZU_MEM_HOST, NULL));
  printf("projected reference positions: %llu\n", (unsigned long long)total);
  free(occs);
  free(occ_offs);
  mazu_b200_free_pinned(hits);
  mazu_b200_free_pinned(bases);
  mazu_b200_index_destroy(idx);
  mazu_b200_index_destroy(dense);
  return counts5[4] == 0 ? 0 : 4;
}
