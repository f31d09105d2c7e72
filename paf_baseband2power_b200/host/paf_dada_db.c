/*
 * paf_dada_db — create or destroy the shared-memory rings of one key: the role of
 * PSRDADA's `dada_db -l -p -k <key> -b <bufsz> -n <nbufs> -r <nreaders>` and
 * `dada_db -d -k <key>` in the reference launcher (paf-baseband2power.py:114-115,
 * 129-130).  -l (lock in RAM) and -p (page) are accepted for compatibility; the
 * stage page-locks the input ring itself through CUDA.
 */
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>

#include "dada/dada_hdu.h"

int main(int argc, char **argv)
{
  key_t key = DADA_DEFAULT_BLOCK_KEY;
  unsigned long long bufsz = 524288, nbufs = 4, hdrsz = DADA_DEFAULT_HEADER_SIZE, nhdr = DADA_DEFAULT_HDR_NBUFS;
  unsigned nreaders = 1;
  int destroy = 0, arg;
  while ((arg = getopt(argc, argv, "k:b:n:r:a:lpdh")) != -1) {
    switch (arg) {
      case 'k':
        if (sscanf(optarg, "%x", (unsigned *)&key) != 1) return EXIT_FAILURE;
        break;
      case 'b': bufsz = strtoull(optarg, NULL, 10); break;
      case 'n': nbufs = strtoull(optarg, NULL, 10); break;
      case 'r': nreaders = (unsigned)atoi(optarg); break;
      case 'a': hdrsz = strtoull(optarg, NULL, 10); break;
      case 'l':
      case 'p': break;
      case 'd': destroy = 1; break;
      default:
        fprintf(stdout,
                "paf_dada_db - create/destroy shared memory ring buffers\n"
                " -k key [dada]  -b buffer bytes  -n buffers  -r readers  -a header bytes [4096]\n"
                " -l lock  -p page  -d destroy\n");
        return EXIT_FAILURE;
    }
  }
  if (destroy) {
    if (dada_hdu_remove_rings(key) < 0) {
      fprintf(stderr, "paf_dada_db: could not remove ring %x\n", (unsigned)key);
      return EXIT_FAILURE;
    }
    printf("Destroyed DADA data and header blocks, key %x\n", (unsigned)key);
    return EXIT_SUCCESS;
  }
  if (dada_hdu_create_rings(key, nbufs, bufsz, nhdr, hdrsz, nreaders) < 0) {
    fprintf(stderr, "paf_dada_db: could not create ring %x\n", (unsigned)key);
    return EXIT_FAILURE;
  }
  printf("Created DADA data block with nbufs=%llu bufsz=%llu nread=%u, key %x\n", nbufs, bufsz, nreaders,
         (unsigned)key);
  return EXIT_SUCCESS;
}
