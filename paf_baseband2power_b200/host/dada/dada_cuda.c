#include "dada_cuda.h"

#include "../../../include/b2p.h"

int dada_cuda_dbregister(dada_hdu_t *hdu)
{
  if (!hdu || !hdu->data_block) return -1;
  ipcbuf_t *db = (ipcbuf_t *)hdu->data_block;
  const uint64_t n = ipcbuf_get_nbufs(db), sz = ipcbuf_get_bufsz(db);
  for (uint64_t i = 0; i < n; ++i)
    if (b2p_host_register(ipcbuf_get_buffer(db, i), sz) != B2P_OK) {
      if (hdu->log) multilog(hdu->log, LOG_ERR, "dada_cuda_dbregister: %s\n", b2p_last_error(0));
      for (uint64_t j = 0; j < i; ++j) b2p_host_unregister(ipcbuf_get_buffer(db, j));
      return -1;
    }
  return 0;
}

int dada_cuda_dbunregister(dada_hdu_t *hdu)
{
  if (!hdu || !hdu->data_block) return -1;
  ipcbuf_t *db = (ipcbuf_t *)hdu->data_block;
  const uint64_t n = ipcbuf_get_nbufs(db);
  int rc = 0;
  for (uint64_t i = 0; i < n; ++i)
    if (b2p_host_unregister(ipcbuf_get_buffer(db, i)) != B2P_OK) rc = -1;
  return rc;
}
