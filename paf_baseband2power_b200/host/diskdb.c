#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif

#include "diskdb.h"

#include <stdlib.h>
#include <string.h>

#include "dada/futils.h"

#define DISKDB_ERR(conf, ...)                                                       \
  do {                                                                              \
    if ((conf)->log) multilog((conf)->log, LOG_ERR, __VA_ARGS__);                   \
    fprintf(stderr, __VA_ARGS__);                                                   \
    fprintf(stderr, "  which happens at \"%s\", line [%d].\n", __FILE__, __LINE__); \
  } while (0)

int init_diskdb(diskdb_conf_t *conf)
{
  conf->nblocks = conf->nbytes = 0;
  conf->fp = fopen(conf->fname, "rb");
  if (!conf->fp) {
    DISKDB_ERR(conf, "Can not open file: %s\n", conf->fname);
    return EXIT_FAILURE;
  }
  conf->hdu = dada_hdu_create(conf->log);
  dada_hdu_set_key(conf->hdu, conf->key);
  if (dada_hdu_connect(conf->hdu) < 0) {
    DISKDB_ERR(conf, "Can not connect to hdu %x\n", (unsigned)conf->key);
    return EXIT_FAILURE;
  }
  ipcbuf_t *db = (ipcbuf_t *)conf->hdu->data_block;
  conf->rbufsz = ipcbuf_get_bufsz(db);
  conf->hdrsz = ipcbuf_get_bufsz(conf->hdu->header_block);
  if (conf->hdrsz != DADA_HDR_SIZE) { /* the ring must have been created for 4096-byte headers */
    DISKDB_ERR(conf, "Header buffer size mismatch: %zu, expected %d\n", conf->hdrsz, DADA_HDR_SIZE);
    return EXIT_FAILURE;
  }
  if (dada_hdu_lock_write(conf->hdu) < 0) { /* make ourselves the write client */
    DISKDB_ERR(conf, "Error locking HDU\n");
    return EXIT_FAILURE;
  }
  const int rc = conf->sod ? ipcbuf_enable_sod(db, 0, 0) : ipcbuf_disable_sod(db);
  if (rc < 0) {
    DISKDB_ERR(conf, "Can not set start-of-data\n");
    return EXIT_FAILURE;
  }
  /* the payload starts after the file's own 4096-byte header */
  if (fseek(conf->fp, DADA_HDR_SIZE, SEEK_SET) != 0) {
    DISKDB_ERR(conf, "Can not skip the file header of %s\n", conf->fname);
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}

int do_diskdb(diskdb_conf_t *conf)
{
  /* the ring's header comes from the separate template file, not from the data file */
  char *hdrbuf = ipcbuf_get_next_write(conf->hdu->header_block);
  if (!hdrbuf || fileread(conf->hfname, hdrbuf, DADA_HDR_SIZE) < 0) {
    DISKDB_ERR(conf, "Error reading header file %s\n", conf->hfname);
    return EXIT_FAILURE;
  }
  if (ipcbuf_mark_filled(conf->hdu->header_block, DADA_HDR_SIZE) < 0) {
    DISKDB_ERR(conf, "Could not mark filled header block\n");
    return EXIT_FAILURE;
  }
  /* block by block; a short (or empty) last block ends the data */
  for (;;) {
    uint64_t block_id = 0;
    char *cur = ipcio_open_block_write(conf->hdu->data_block, &block_id);
    if (!cur) {
      DISKDB_ERR(conf, "Can not open a ring block for writing\n");
      return EXIT_FAILURE;
    }
    const size_t got = fread(cur, 1, conf->rbufsz, conf->fp);
    if (ipcio_close_block_write(conf->hdu->data_block, got) < 0) {
      DISKDB_ERR(conf, "Can not close the ring block\n");
      return EXIT_FAILURE;
    }
    conf->nbytes += got;
    if (got) conf->nblocks++;
    if (got < conf->rbufsz) break;
  }
  if (conf->log)
    multilog(conf->log, LOG_INFO, "diskdb: %lu bytes in %lu blocks from %s\n", conf->nbytes,
             conf->nblocks, conf->fname);
  return EXIT_SUCCESS;
}

int destroy_diskdb(diskdb_conf_t *conf)
{
  if (conf->hdu) {
    dada_hdu_unlock_write(conf->hdu);
    dada_hdu_disconnect(conf->hdu);
    dada_hdu_destroy(conf->hdu);
    conf->hdu = NULL;
  }
  if (conf->fp) fclose(conf->fp);
  conf->fp = NULL;
  return EXIT_SUCCESS;
}
