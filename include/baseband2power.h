/*
 * baseband2power.h — the baseband2power stage: ring in -> GPU -> ring out.
 *
 * This is the stage the reference declares and never implements:
 * baseband2power.cuh:18-23 holds only conf_t {device_id, dir, key_in, key_out},
 * baseband2power.cu:1-16 is empty and paf_baseband2power.cu:92 returns before
 * any work.  The entry points keep the init_/do_/destroy_ convention of the
 * sibling stage (diskdb.cuh:32-34) and its EXIT_SUCCESS / EXIT_FAILURE returns.
 *
 *   init_baseband2power    connect + lock_read key_in, connect + lock_write
 *                          key_out, check the block geometry (in = ndf *
 *                          NCHK_NIC * 7168, paf-baseband2power.py:67; out =
 *                          NCHAN * NBYTE = 1344, :79; header 4096,
 *                          diskdb.cuh:17), page-lock the input ring, create the
 *                          b2p context on conf->device_id
 *   do_baseband2power      pass the DADA header through (setting the keys this
 *                          stage changes), then per input block:
 *                          b2p_accumulate_host -> every `ndf_integration`
 *                          frames b2p_finish into an output block
 *   destroy_baseband2power unlock, disconnect, free
 */
#ifndef BASEBAND2POWER_H
#define BASEBAND2POWER_H

#include <stdint.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSTR_LEN 512 /* paf_baseband2power.cuh:4 */

struct dada_hdu_t;
struct multilog_t;
struct b2p_ctx;
struct b2p_group;

#define B2P_STAGE_MAX_GPUS 16

typedef struct conf_t {
  /* the reference's fields (baseband2power.cuh:18-23) */
  int device_id;
  char dir[MSTR_LEN];
  key_t key_in, key_out;
  /* geometry: defaults from paf-baseband2power.conf:2-5,24 */
  int nchunk, nch_per_chunk, nsamp_df;
  int big_endian;
  int average;              /* 0: integral (README.md:2)  1: time average (paf_baseband2power.cu:20) */
  uint64_t ndf_integration; /* frames per integration; 0 = frames of one input block */
  int kernel;               /* B2P_KERNEL_* */
  int pin_ring;             /* page-lock the input ring (default 1) */
  /* channel-group sharding of this beam over several GPUs (-d 0,1,2,3): GPU gpus[i] takes
     gpu_chunks[i] consecutive chunks; all zero = split in proportion to the measured
     host-link rate of each GPU.  ngpu <= 1: the reference's one GPU per stage process
     (paf_baseband2power.cu:23-26). */
  int ngpu;
  int gpus[B2P_STAGE_MAX_GPUS];
  int gpu_chunks[B2P_STAGE_MAX_GPUS];
  /* runtime state */
  struct dada_hdu_t *hdu_in, *hdu_out;
  struct multilog_t *log;
  struct b2p_ctx *ctx;      /* ngpu <= 1 */
  struct b2p_group *grp;    /* ngpu > 1 */
  uint64_t rbufsz_in, rbufsz_out, ndf_block;
  int ring_pinned;
  /* statistics */
  uint64_t nblocks_in, nblocks_out, nframes_dropped;
  double seconds_busy;
  /* the same for the blocks after the split has settled (group mode tunes on the first 8) */
  uint64_t nblocks_steady;
  double seconds_busy_steady, seconds_block_max;
} conf_t;

void default_baseband2power(conf_t *conf);
int init_baseband2power(conf_t *conf);
int do_baseband2power(conf_t *conf);
int destroy_baseband2power(conf_t *conf);

#ifdef __cplusplus
}
#endif
#endif
