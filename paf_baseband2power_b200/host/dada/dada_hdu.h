/*
 * dada_hdu.h — "header + data unit": a header ring and a data ring under one key
 * (PSRDADA-named shim).  The reference drives it exactly like this:
 * dada_hdu_create / set_key / connect / lock_write ... unlock_write /
 * disconnect / destroy (diskdb.cu:24-50,128-130; capture.c:590-633).
 * The header ring lives at key+1, as in PSRDADA.
 */
#ifndef B2P_DADA_HDU_H
#define B2P_DADA_HDU_H

#include "dada_def.h"
#include "ipcio.h"
#include "multilog.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dada_hdu_t {
  multilog_t *log;
  ipcio_t *data_block;
  ipcbuf_t *header_block;
  char *header;
  uint64_t header_size;
  key_t data_block_key, header_block_key;
} dada_hdu_t;

dada_hdu_t *dada_hdu_create(multilog_t *log);
void dada_hdu_set_key(dada_hdu_t *hdu, key_t key);
int dada_hdu_connect(dada_hdu_t *hdu);
int dada_hdu_disconnect(dada_hdu_t *hdu);
int dada_hdu_lock_write(dada_hdu_t *hdu);
int dada_hdu_unlock_write(dada_hdu_t *hdu);
int dada_hdu_lock_read(dada_hdu_t *hdu);
int dada_hdu_unlock_read(dada_hdu_t *hdu);
void dada_hdu_destroy(dada_hdu_t *hdu);

/* dada_db's job: create / remove both rings of a key */
int dada_hdu_create_rings(key_t key, uint64_t nbufs, uint64_t bufsz, uint64_t hdr_nbufs,
                          uint64_t hdr_bufsz, unsigned nreaders);
int dada_hdu_remove_rings(key_t key);

#ifdef __cplusplus
}
#endif
#endif
