#include "dada_hdu.h"

#include <stdio.h>
#include <stdlib.h>

dada_hdu_t *dada_hdu_create(multilog_t *log)
{
  dada_hdu_t *h = (dada_hdu_t *)calloc(1, sizeof(dada_hdu_t));
  if (!h) return NULL;
  h->log = log;
  dada_hdu_set_key(h, DADA_DEFAULT_BLOCK_KEY);
  return h;
}

void dada_hdu_set_key(dada_hdu_t *h, key_t key)
{
  h->data_block_key = key;
  h->header_block_key = key + 1;
}

int dada_hdu_connect(dada_hdu_t *h)
{
  if (!h || h->data_block) return -1;
  ipcio_t *d = (ipcio_t *)calloc(1, sizeof(ipcio_t));
  ipcbuf_t *hb = (ipcbuf_t *)calloc(1, sizeof(ipcbuf_t));
  if (!d || !hb) return -1;
  if (ipcbuf_connect(hb, h->header_block_key) < 0) {
    if (h->log) multilog(h->log, LOG_ERR, "cannot connect to header block %x\n", (unsigned)h->header_block_key);
    free(d);
    free(hb);
    return -1;
  }
  if (ipcio_connect(d, h->data_block_key) < 0) {
    if (h->log) multilog(h->log, LOG_ERR, "cannot connect to data block %x\n", (unsigned)h->data_block_key);
    ipcbuf_disconnect(hb);
    free(d);
    free(hb);
    return -1;
  }
  h->data_block = d;
  h->header_block = hb;
  h->header_size = ipcbuf_get_bufsz(hb);
  return 0;
}

int dada_hdu_disconnect(dada_hdu_t *h)
{
  if (!h || !h->data_block) return -1;
  ipcio_disconnect(h->data_block);
  ipcbuf_disconnect(h->header_block);
  free(h->data_block);
  free(h->header_block);
  h->data_block = NULL;
  h->header_block = NULL;
  return 0;
}

int dada_hdu_lock_write(dada_hdu_t *h)
{
  if (!h || !h->data_block) return -1;
  if (ipcbuf_lock_write(h->header_block) < 0) return -1;
  if (ipcio_open(h->data_block, 'W') < 0) {
    ipcbuf_unlock_write(h->header_block);
    return -1;
  }
  return 0;
}

int dada_hdu_unlock_write(dada_hdu_t *h)
{
  if (!h || !h->data_block) return -1;
  int a = ipcio_close(h->data_block);
  int b = ipcbuf_unlock_write(h->header_block);
  return (a < 0 || b < 0) ? -1 : 0;
}

int dada_hdu_lock_read(dada_hdu_t *h)
{
  if (!h || !h->data_block) return -1;
  if (ipcbuf_lock_read(h->header_block) < 0) return -1;
  if (ipcio_open(h->data_block, 'R') < 0) {
    ipcbuf_unlock_read(h->header_block);
    return -1;
  }
  return 0;
}

int dada_hdu_unlock_read(dada_hdu_t *h)
{
  if (!h || !h->data_block) return -1;
  int a = ipcio_close(h->data_block);
  int b = ipcbuf_unlock_read(h->header_block);
  return (a < 0 || b < 0) ? -1 : 0;
}

void dada_hdu_destroy(dada_hdu_t *h)
{
  if (!h) return;
  if (h->data_block) dada_hdu_disconnect(h);
  free(h);
}

int dada_hdu_create_rings(key_t key, uint64_t nbufs, uint64_t bufsz, uint64_t hdr_nbufs,
                          uint64_t hdr_bufsz, unsigned nreaders)
{
  ipcbuf_t data = IPCBUF_INIT, hdr = IPCBUF_INIT;
  if (ipcbuf_create(&data, key, nbufs, bufsz, nreaders) < 0) return -1;
  if (ipcbuf_create(&hdr, key + 1, hdr_nbufs, hdr_bufsz, nreaders) < 0) {
    ipcbuf_destroy(&data);
    return -1;
  }
  ipcbuf_disconnect(&data);
  ipcbuf_disconnect(&hdr);
  return 0;
}

int dada_hdu_remove_rings(key_t key)
{
  int rc = 0;
  ipcbuf_t data = IPCBUF_INIT, hdr = IPCBUF_INIT;
  if (ipcbuf_connect(&data, key) == 0) ipcbuf_destroy(&data); else rc = -1;
  if (ipcbuf_connect(&hdr, key + 1) == 0) ipcbuf_destroy(&hdr); else rc = -1;
  return rc;
}
