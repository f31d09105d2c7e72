#include "ipcio.h"

#include <stdio.h>

int ipcio_create(ipcio_t *ipc, key_t key, uint64_t nbufs, uint64_t bufsz, unsigned nreaders)
{
  if (!ipc) return -1;
  ipc->rdwrt = 0;
  ipc->curbuf = 0;
  ipc->block_open = 0;
  return ipcbuf_create(&ipc->buf, key, nbufs, bufsz, nreaders);
}

int ipcio_connect(ipcio_t *ipc, key_t key)
{
  if (!ipc) return -1;
  ipc->rdwrt = 0;
  ipc->curbuf = 0;
  ipc->block_open = 0;
  return ipcbuf_connect(&ipc->buf, key);
}

int ipcio_disconnect(ipcio_t *ipc) { return ipc ? ipcbuf_disconnect(&ipc->buf) : -1; }
int ipcio_destroy(ipcio_t *ipc) { return ipc ? ipcbuf_destroy(&ipc->buf) : -1; }

int ipcio_open(ipcio_t *ipc, char rdwrt)
{
  if (!ipc || ipc->rdwrt) return -1;
  if (rdwrt == 'W' || rdwrt == 'w') {
    if (ipcbuf_lock_write(&ipc->buf) < 0) return -1;
    ipc->rdwrt = 'W';
  } else if (rdwrt == 'R' || rdwrt == 'r') {
    if (ipcbuf_lock_read(&ipc->buf) < 0) return -1;
    ipc->rdwrt = 'R';
  } else {
    return -1;
  }
  return 0;
}

int ipcio_close(ipcio_t *ipc)
{
  if (!ipc || !ipc->rdwrt) return -1;
  int r;
  if (ipc->rdwrt == 'W') {
    if (ipc->block_open) ipcio_close_block_write(ipc, 0);
    ipcbuf_enable_eod(&ipc->buf);
    r = ipcbuf_unlock_write(&ipc->buf);
  } else {
    if (ipc->block_open) ipcio_close_block_read(ipc, ipc->curbufsz);
    r = ipcbuf_unlock_read(&ipc->buf);
  }
  ipc->rdwrt = 0;
  return r;
}

char *ipcio_open_block_write(ipcio_t *ipc, uint64_t *block_id)
{
  if (!ipc || ipc->rdwrt != 'W' || ipc->block_open) {
    fprintf(stderr, "ipcio_open_block_write: ring not open for writing or block already open\n");
    return NULL;
  }
  ipc->curbuf = ipcbuf_get_next_write(&ipc->buf);
  if (!ipc->curbuf) return NULL;
  if (block_id) *block_id = ipcbuf_get_write_count(&ipc->buf) % ipcbuf_get_nbufs(&ipc->buf);
  ipc->block_open = 1;
  return ipc->curbuf;
}

ssize_t ipcio_close_block_write(ipcio_t *ipc, uint64_t bytes)
{
  if (!ipc || ipc->rdwrt != 'W' || !ipc->block_open) return -1;
  if (ipcbuf_mark_filled(&ipc->buf, bytes) < 0) return -1;
  ipc->block_open = 0;
  ipc->curbuf = 0;
  return 0;
}

char *ipcio_open_block_read(ipcio_t *ipc, uint64_t *bytes, uint64_t *block_id)
{
  if (!ipc || ipc->rdwrt != 'R' || ipc->block_open) {
    fprintf(stderr, "ipcio_open_block_read: ring not open for reading or block already open\n");
    return NULL;
  }
  for (;;) {
    uint64_t n = 0;
    const uint64_t id = ipcbuf_get_read_count(&ipc->buf) % ipcbuf_get_nbufs(&ipc->buf);
    char *p = ipcbuf_get_next_read(&ipc->buf, &n);
    if (!p) return NULL;
    if (n == 0) { /* an empty terminating block (file size was a multiple of the block size) */
      ipcbuf_mark_cleared(&ipc->buf);
      continue;
    }
    ipc->curbuf = p;
    ipc->curbufsz = n;
    ipc->block_open = 1;
    if (bytes) *bytes = n;
    if (block_id) *block_id = id;
    return p;
  }
}

ssize_t ipcio_close_block_read(ipcio_t *ipc, uint64_t bytes)
{
  (void)bytes;
  if (!ipc || ipc->rdwrt != 'R' || !ipc->block_open) return -1;
  if (ipcbuf_mark_cleared(&ipc->buf) < 0) return -1;
  ipc->block_open = 0;
  ipc->curbuf = 0;
  return 0;
}
