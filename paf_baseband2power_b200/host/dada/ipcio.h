/*
 * ipcio.h — block I/O over an ipcbuf ring (PSRDADA-named shim).
 *
 * ipcio_open_block_write / ipcio_close_block_write are what the reference's
 * producers use (diskdb.cu:105-109; capture.c:316; sync.c:101,109); the read
 * pair is what the baseband2power stage needs.  An ipcio_t starts with its
 * ipcbuf_t so the reference's cast `(ipcbuf_t *) hdu->data_block`
 * (diskdb.cu:33) keeps working.
 */
#ifndef B2P_IPCIO_H
#define B2P_IPCIO_H

#include <sys/types.h>

#include "ipcbuf.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ipcio_t {
  ipcbuf_t buf;     /* must stay first */
  char rdwrt;       /* 0 closed, 'W' writer, 'R' reader */
  char *curbuf;     /* block currently borrowed */
  uint64_t curbufsz;
  int block_open;
} ipcio_t;

#define IPCIO_INIT {IPCBUF_INIT, 0, 0, 0, 0}

int ipcio_create(ipcio_t *ipc, key_t key, uint64_t nbufs, uint64_t bufsz, unsigned nreaders);
int ipcio_connect(ipcio_t *ipc, key_t key);
int ipcio_disconnect(ipcio_t *ipc);
int ipcio_destroy(ipcio_t *ipc);

/* 'W'/'w' lock as writer, 'R'/'r' lock as reader */
int ipcio_open(ipcio_t *ipc, char rdwrt);
/* writer: raise end-of-data if not raised yet, then unlock; reader: unlock */
int ipcio_close(ipcio_t *ipc);

char *ipcio_open_block_write(ipcio_t *ipc, uint64_t *block_id);
/* bytes < block size ends the data stream */
ssize_t ipcio_close_block_write(ipcio_t *ipc, uint64_t bytes);

/* NULL at end of data */
char *ipcio_open_block_read(ipcio_t *ipc, uint64_t *bytes, uint64_t *block_id);
ssize_t ipcio_close_block_read(ipcio_t *ipc, uint64_t bytes);

#ifdef __cplusplus
}
#endif
#endif
