/*
 * paf_dbdisk — drain a ring to a DADA file: what `dada_dbdisk -k <key> -D <dir> -W`
 * does at the end of the reference pipeline (paf-baseband2power.py:94-95).
 * Writes <dir>/<UTC_START>_<OBS_OFFSET 16 digits>.000000.dada = the 4096-byte
 * header (FILE_SIZE / OBS_OFFSET set) followed by every block until end of data.
 * -f <name> overrides the file name.
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "dada/ascii_header.h"
#include "dada/dada_hdu.h"

static void usage(void)
{
  fprintf(stdout,
          "paf_dbdisk - write a ring buffer to a DADA file\n"
          " -k  hexadecimal shared memory key [default dada]\n"
          " -D  output directory [default .]\n"
          " -f  output file name [default <UTC_START>_<offset>.000000.dada]\n"
          " -W  over-write an existing file\n"
          " -h  show help\n");
}

int main(int argc, char **argv)
{
  key_t key = 0xdada;
  char dir[512] = ".", name[512] = "";
  int overwrite = 0, arg;
  while ((arg = getopt(argc, argv, "k:D:f:Wb:h")) != -1) {
    switch (arg) {
      case 'k':
        if (sscanf(optarg, "%x", (unsigned *)&key) != 1) return EXIT_FAILURE;
        break;
      case 'D': snprintf(dir, sizeof(dir), "%s", optarg); break;
      case 'f': snprintf(name, sizeof(name), "%s", optarg); break;
      case 'W': overwrite = 1; break;
      case 'b': break; /* core binding of dada_dbdisk: accepted, ignored */
      default: usage(); return EXIT_FAILURE;
    }
  }
  multilog_t *log = multilog_open("paf_dbdisk", 0);
  multilog_add(log, stderr);
  dada_hdu_t *hdu = dada_hdu_create(log);
  dada_hdu_set_key(hdu, key);
  if (dada_hdu_connect(hdu) < 0 || dada_hdu_lock_read(hdu) < 0) {
    fprintf(stderr, "paf_dbdisk: can not connect to / lock ring %x\n", (unsigned)key);
    return EXIT_FAILURE;
  }
  uint64_t hbytes = 0;
  char *hdr = ipcbuf_get_next_read(hdu->header_block, &hbytes);
  if (!hdr) {
    fprintf(stderr, "paf_dbdisk: no header on ring %x\n", (unsigned)key);
    return EXIT_FAILURE;
  }
  char header[DADA_DEFAULT_HEADER_SIZE];
  memcpy(header, hdr, DADA_DEFAULT_HEADER_SIZE);
  header[DADA_DEFAULT_HEADER_SIZE - 1] = 0;
  ipcbuf_mark_cleared(hdu->header_block);

  if (!name[0]) {
    char utc[64] = "unset";
    ascii_header_get(header, "UTC_START", "%63s", utc);
    snprintf(name, sizeof(name), "%s_%016d.000000.dada", utc, 0);
  }
  char path[1100];
  snprintf(path, sizeof(path), "%s/%s", dir, name);
  if (!overwrite && access(path, F_OK) == 0) {
    fprintf(stderr, "paf_dbdisk: %s exists (use -W)\n", path);
    return EXIT_FAILURE;
  }
  FILE *fp = fopen(path, "wb");
  if (!fp) {
    fprintf(stderr, "paf_dbdisk: can not create %s\n", path);
    return EXIT_FAILURE;
  }
  ascii_header_set(header, "OBS_OFFSET", "%d", 0);
  /* the header is rewritten with the final FILE_SIZE once the data are in */
  fwrite(header, 1, DADA_DEFAULT_HEADER_SIZE, fp);
  unsigned long total = 0, nblk = 0;
  for (;;) {
    uint64_t bytes = 0, id = 0;
    char *blk = ipcio_open_block_read(hdu->data_block, &bytes, &id);
    if (!blk) break;
    if (fwrite(blk, 1, bytes, fp) != bytes) {
      fprintf(stderr, "paf_dbdisk: short write to %s\n", path);
      return EXIT_FAILURE;
    }
    total += bytes;
    nblk++;
    ipcio_close_block_read(hdu->data_block, bytes);
  }
  ascii_header_set(header, "FILE_SIZE", "%lu", total);
  size_t hl = strlen(header);
  memset(header + hl, 0, DADA_DEFAULT_HEADER_SIZE - hl);
  fseek(fp, 0, SEEK_SET);
  fwrite(header, 1, DADA_DEFAULT_HEADER_SIZE, fp);
  fclose(fp);
  multilog(log, LOG_INFO, "wrote %lu bytes in %lu blocks to %s\n", total, nblk, path);
  dada_hdu_unlock_read(hdu);
  dada_hdu_disconnect(hdu);
  dada_hdu_destroy(hdu);
  multilog_close(log);
  return EXIT_SUCCESS;
}
