/*
 * dada_cuda.h — page-lock a connected ring for CUDA (PSRDADA-named shim).
 * The reference includes dada_cuda.h in the stage header (baseband2power.cuh:9)
 * but never calls it; dada_cuda_dbregister is what it exists for: after it, the
 * blocks returned by ipcio_open_block_read are pinned, so cudaMemcpyAsync from
 * them is truly asynchronous.  Implemented over b2p_host_register (include/b2p.h).
 */
#ifndef B2P_DADA_CUDA_H
#define B2P_DADA_CUDA_H
#include "dada_hdu.h"
#ifdef __cplusplus
extern "C" {
#endif
int dada_cuda_dbregister(dada_hdu_t *hdu);
int dada_cuda_dbunregister(dada_hdu_t *hdu);
#ifdef __cplusplus
}
#endif
#endif
