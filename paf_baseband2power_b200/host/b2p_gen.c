/*
 * b2p_gen — write a synthetic BMF DADA file: 4096-byte ASCII header (from a
 * template, UTC_START / FREQ / BW filled) + `ndf` data frames of the
 * counter-based stream of include/b2p_synth.h in ring-block layout
 * (capture.c:540-542).  The reference ships no sample data; this
 * is the input for config 1: paf_diskdb -> paf_baseband2power.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "../../include/b2p_synth.h"
#include "dada/ascii_header.h"
#include "dada/dada_def.h"
#include "dada/futils.h"

int main(int argc, char **argv)
{
  char out[1024] = "synthetic.dada", tmpl[1024] = "";
  uint64_t ndf = 64, seed = 1;
  int mode = B2P_SYNTH_GAUSS, be = 1, nchunk = 48, nch = 7, nsamp = 128, arg;
  while ((arg = getopt(argc, argv, "o:n:s:m:H:e:h")) != -1) {
    switch (arg) {
      case 'o': snprintf(out, sizeof(out), "%s", optarg); break;
      case 'n': ndf = strtoull(optarg, NULL, 10); break;
      case 's': seed = strtoull(optarg, NULL, 10); break;
      case 'm': mode = atoi(optarg); break;
      case 'H': snprintf(tmpl, sizeof(tmpl), "%s", optarg); break;
      case 'e': be = atoi(optarg); break;
      default:
        fprintf(stdout, "b2p_gen -o file -n ndf -s seed -m mode(0 uniform,1 gauss) -H header_template -e big_endian\n");
        return EXIT_FAILURE;
    }
  }
  char header[DADA_DEFAULT_HEADER_SIZE];
  memset(header, 0, sizeof(header));
  if (tmpl[0]) {
    if (fileread(tmpl, header, sizeof(header)) < 0) {
      fprintf(stderr, "b2p_gen: can not read %s\n", tmpl);
      return EXIT_FAILURE;
    }
  } else {
    strcpy(header, "HEADER       DADA\nHDR_VERSION  1.0\nHDR_SIZE     4096\n");
  }
  ascii_header_set(header, "UTC_START", "%s", "2026-10-18-00:00:00");
  ascii_header_set(header, "PICOSECONDS", "%d", 0);
  ascii_header_set(header, "FREQ", "%.1f", 1340.5);
  ascii_header_set(header, "BW", "%d", 336);
  ascii_header_set(header, "NBIT", "%d", 16);
  ascii_header_set(header, "NDIM", "%d", 2);
  ascii_header_set(header, "NPOL", "%d", 2);
  ascii_header_set(header, "TSAMP", "%.5f", 27.0 / 32.0);
  const uint64_t wpf = (uint64_t)nchunk * nsamp * nch;
  ascii_header_set(header, "FILE_SIZE", "%lu", (unsigned long)(ndf * wpf * 8));
  size_t hl = strlen(header);
  memset(header + hl, 0, sizeof(header) - hl);

  FILE *fp = fopen(out, "wb");
  if (!fp) {
    fprintf(stderr, "b2p_gen: can not create %s\n", out);
    return EXIT_FAILURE;
  }
  fwrite(header, 1, sizeof(header), fp);
  uint64_t *frame = (uint64_t *)malloc(wpf * 8);
  const int nchan = nchunk * nch;
  for (uint64_t f = 0; f < ndf; ++f) {
    for (uint64_t k = 0; k < wpf; ++k) {
      int16_t v[4];
      const uint64_t w = f * wpf + k;
      b2p_synth_word(seed, w, b2p_synth_chan(w, nchunk, nch, nsamp), nchan, mode, v);
      frame[k] = b2p_synth_pack(v, be);
    }
    if (fwrite(frame, 8, wpf, fp) != wpf) {
      fprintf(stderr, "b2p_gen: short write\n");
      return EXIT_FAILURE;
    }
  }
  free(frame);
  fclose(fp);
  printf("wrote %s: 4096-byte header + %lu frames (%lu bytes)\n", out, (unsigned long)ndf,
         (unsigned long)(ndf * wpf * 8));
  return EXIT_SUCCESS;
}
