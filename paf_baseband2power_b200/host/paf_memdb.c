/*
 * paf_memdb — memory-resident synthetic producer for the input ring.
 *
 * BASELINE.json configs[1] asks for a continuous stream of integrations through
 * the DADA ring with the producer out of the way ("rotate a few distinct
 * pre-generated blocks").  This tool is that producer: it generates one distinct
 * block of the b2p_synth stream in place into each ring buffer on its first use
 * (block i gets stream position i*words_per_block, seed s) and afterwards only
 * re-publishes the buffers — block i carries the data of block i % nbufs — so it
 * can feed the reader at any rate the reader sustains.  The header is written
 * from the template like paf_diskdb does (diskdb.cu:79-93).
 *
 *   -k key  -n blocks to publish  -s seed  -m mode  -H header template  -e sod
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "../../include/b2p_synth.h"
#include "dada/ascii_header.h"
#include "dada/dada_hdu.h"
#include "dada/futils.h"

int main(int argc, char **argv)
{
  key_t key = 0xdada;
  uint64_t nblocks = 8, seed = 1;
  int mode = B2P_SYNTH_GAUSS, sod = 1, arg;
  char tmpl[1024] = "";
  const int nchunk = 48, nch = 7, nsamp = 128, nchan = nchunk * nch;
  while ((arg = getopt(argc, argv, "k:n:s:m:H:e:h")) != -1) {
    switch (arg) {
      case 'k':
        if (sscanf(optarg, "%x", (unsigned *)&key) != 1) return EXIT_FAILURE;
        break;
      case 'n': nblocks = strtoull(optarg, NULL, 10); break;
      case 's': seed = strtoull(optarg, NULL, 10); break;
      case 'm': mode = atoi(optarg); break;
      case 'H': snprintf(tmpl, sizeof(tmpl), "%s", optarg); break;
      case 'e': sod = atoi(optarg); break;
      default:
        fprintf(stdout, "paf_memdb -k key -n blocks -s seed -m mode -H header_template -e sod\n");
        return EXIT_FAILURE;
    }
  }
  multilog_t *log = multilog_open("paf_memdb", 0);
  multilog_add(log, stderr);
  dada_hdu_t *hdu = dada_hdu_create(log);
  dada_hdu_set_key(hdu, key);
  if (dada_hdu_connect(hdu) < 0 || dada_hdu_lock_write(hdu) < 0) {
    fprintf(stderr, "paf_memdb: can not connect to / lock ring %x\n", (unsigned)key);
    return EXIT_FAILURE;
  }
  ipcbuf_t *db = (ipcbuf_t *)hdu->data_block;
  const uint64_t bufsz = ipcbuf_get_bufsz(db), nbufs = ipcbuf_get_nbufs(db);
  const uint64_t frame = (uint64_t)nchunk * nsamp * nch * 8;
  if (bufsz % frame) {
    fprintf(stderr, "paf_memdb: ring block %lu is not a whole number of data frames\n", (unsigned long)bufsz);
    return EXIT_FAILURE;
  }
  if (sod) ipcbuf_enable_sod(db, 0, 0);
  char *hdr = ipcbuf_get_next_write(hdu->header_block);
  if (!hdr || !tmpl[0] || fileread(tmpl, hdr, DADA_DEFAULT_HEADER_SIZE) < 0) {
    fprintf(stderr, "paf_memdb: can not read header template '%s'\n", tmpl);
    return EXIT_FAILURE;
  }
  ascii_header_set(hdr, "UTC_START", "%s", "2026-10-18-00:00:00");
  ascii_header_set(hdr, "PICOSECONDS", "%d", 0);
  ascii_header_set(hdr, "FREQ", "%.1f", 1340.5);
  ipcbuf_mark_filled(hdu->header_block, DADA_DEFAULT_HEADER_SIZE);

  const uint64_t wpb = bufsz / 8;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (uint64_t b = 0; b < nblocks; ++b) {
    uint64_t id = 0;
    uint64_t *buf = (uint64_t *)ipcio_open_block_write(hdu->data_block, &id);
    if (!buf) return EXIT_FAILURE;
    if (b < nbufs) { /* first use of this ring buffer: generate its block in place */
#pragma omp parallel for schedule(static)
      for (int64_t w = 0; w < (int64_t)wpb; ++w) {
        int16_t v[4];
        b2p_synth_word(seed, b * wpb + (uint64_t)w, b2p_synth_chan((uint64_t)w, nchunk, nch, nsamp), nchan, mode, v);
        buf[w] = b2p_synth_pack(v, 1);
      }
    }
    ipcio_close_block_write(hdu->data_block, bufsz);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  multilog(log, LOG_INFO, "published %lu blocks of %lu bytes (%lu distinct) in %.3f s\n", (unsigned long)nblocks,
           (unsigned long)bufsz, (unsigned long)(nblocks < nbufs ? nblocks : nbufs),
           (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec));
  dada_hdu_unlock_write(hdu);
  dada_hdu_disconnect(hdu);
  dada_hdu_destroy(hdu);
  multilog_close(log);
  return EXIT_SUCCESS;
}
