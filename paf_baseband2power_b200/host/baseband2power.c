/*
 * baseband2power.c — host side of the stage (see include/baseband2power.h for what
 * of the reference each function stands in for).  Plain C over two ABIs: the
 * ring-buffer shim (PSRDADA names) and libb2p (include/b2p.h).  No CUDA here.
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif

#include "../../include/baseband2power.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/b2p.h"
#include "dada/ascii_header.h"
#include "dada/dada_cuda.h"
#include "dada/dada_hdu.h"
#include "dada/multilog.h"

#define PKT_BYTES(c) ((uint64_t)(c)->nsamp_df * (c)->nch_per_chunk * 8u)
#define TSAMP_US (27.0 / 32.0) /* README.md:2 */

#define STAGE_ERR(conf, ...)                                                              \
  do {                                                                                    \
    if ((conf)->log) multilog((conf)->log, LOG_ERR, __VA_ARGS__);                         \
    fprintf(stderr, __VA_ARGS__);                                                         \
    fprintf(stderr, "  which happens at \"%s\", line [%d].\n", __FILE__, __LINE__);       \
  } while (0)

void default_baseband2power(conf_t *conf)
{
  memset(conf, 0, sizeof(*conf));
  conf->device_id = 0;
  strcpy(conf->dir, ".");
  conf->key_in = 0xdada;  /* paf-baseband2power.conf:13 */
  conf->key_out = 0xadad; /* paf-baseband2power.conf:20 */
  conf->nchunk = 48;
  conf->nch_per_chunk = 7;
  conf->nsamp_df = 128;
  conf->big_endian = 1;
  conf->average = 0;
  conf->ndf_integration = 0;
  conf->kernel = B2P_KERNEL_AUTO;
  conf->pin_ring = 1;
  conf->ngpu = 0;
}

static double now_s(void)
{
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

int init_baseband2power(conf_t *conf)
{
  const uint64_t frame_bytes = (uint64_t)conf->nchunk * PKT_BYTES(conf);
  const uint64_t nchan = (uint64_t)conf->nchunk * conf->nch_per_chunk;

  /* input ring: we are its reader */
  conf->hdu_in = dada_hdu_create(conf->log);
  dada_hdu_set_key(conf->hdu_in, conf->key_in);
  if (dada_hdu_connect(conf->hdu_in) < 0) {
    STAGE_ERR(conf, "Can not connect to input hdu %x\n", (unsigned)conf->key_in);
    return EXIT_FAILURE;
  }
  /* output ring: we are its writer */
  conf->hdu_out = dada_hdu_create(conf->log);
  dada_hdu_set_key(conf->hdu_out, conf->key_out);
  if (dada_hdu_connect(conf->hdu_out) < 0) {
    STAGE_ERR(conf, "Can not connect to output hdu %x\n", (unsigned)conf->key_out);
    return EXIT_FAILURE;
  }

  conf->rbufsz_in = ipcbuf_get_bufsz((ipcbuf_t *)conf->hdu_in->data_block);
  conf->rbufsz_out = ipcbuf_get_bufsz((ipcbuf_t *)conf->hdu_out->data_block);
  if (conf->rbufsz_in == 0 || conf->rbufsz_in % frame_bytes) {
    STAGE_ERR(conf, "Input block size %lu is not a whole number of %lu-byte data frames\n",
              (unsigned long)conf->rbufsz_in, (unsigned long)frame_bytes);
    return EXIT_FAILURE;
  }
  if (conf->rbufsz_out != nchan * sizeof(float)) {
    STAGE_ERR(conf, "Output block size %lu does not match NCHAN*NBYTE = %lu\n",
              (unsigned long)conf->rbufsz_out, (unsigned long)(nchan * sizeof(float)));
    return EXIT_FAILURE;
  }
  if (ipcbuf_get_bufsz(conf->hdu_in->header_block) != DADA_DEFAULT_HEADER_SIZE ||
      ipcbuf_get_bufsz(conf->hdu_out->header_block) != DADA_DEFAULT_HEADER_SIZE) {
    STAGE_ERR(conf, "Header block size mismatch, expected %d\n", DADA_DEFAULT_HEADER_SIZE);
    return EXIT_FAILURE;
  }
  conf->ndf_block = conf->rbufsz_in / frame_bytes;
  if (conf->ndf_integration == 0) conf->ndf_integration = conf->ndf_block;

  if (dada_hdu_lock_read(conf->hdu_in) < 0) {
    STAGE_ERR(conf, "Error locking input HDU for reading\n");
    return EXIT_FAILURE;
  }
  if (dada_hdu_lock_write(conf->hdu_out) < 0) {
    STAGE_ERR(conf, "Error locking output HDU for writing\n");
    return EXIT_FAILURE;
  }

  /* the GPU side */
  b2p_params p;
  b2p_default_params(&p);
  p.device_id = conf->device_id;
  p.nchunk = conf->nchunk;
  p.nch_per_chunk = conf->nch_per_chunk;
  p.nsamp_df = conf->nsamp_df;
  p.big_endian = conf->big_endian;
  p.kernel = conf->kernel;
  p.nbeam = 1;
  p.scale = conf->average ? (float)(1.0 / ((double)conf->ndf_integration * conf->nsamp_df)) : 1.0f;
  if (conf->ngpu > 1) {
    /* one beam over several GPUs by channel group: split the chunks in proportion to what
       each GPU's host link delivers with all of them copying at once, unless told */
    int given = 0;
    for (int i = 0; i < conf->ngpu; ++i) given += conf->gpu_chunks[i];
    if (given == 0) {
      double rate[B2P_STAGE_MAX_GPUS];
      if (b2p_probe_h2d(conf->gpus, conf->ngpu, (size_t)256 << 20, 4, rate) != B2P_OK ||
          b2p_split_chunks(rate, conf->ngpu, conf->nchunk, conf->gpu_chunks) != B2P_OK) {
        STAGE_ERR(conf, "host-link probe failed: %s\n", b2p_last_error(NULL));
        return EXIT_FAILURE;
      }
      if (conf->log)
        for (int i = 0; i < conf->ngpu; ++i)
          multilog(conf->log, LOG_INFO, "gpu %d: host link %.1f GB/s -> %d chunks\n", conf->gpus[i],
                   rate[i], conf->gpu_chunks[i]);
    } else if (given != conf->nchunk) {
      STAGE_ERR(conf, "Chunk counts per GPU add up to %d, not %d\n", given, conf->nchunk);
      return EXIT_FAILURE;
    }
    if (b2p_group_create(&conf->grp, &p, conf->gpus, conf->gpu_chunks, conf->ngpu) != B2P_OK) {
      STAGE_ERR(conf, "b2p_group_create failed: %s\n", b2p_last_error(NULL));
      return EXIT_FAILURE;
    }
  } else if (b2p_create(&conf->ctx, &p) != B2P_OK) {
    STAGE_ERR(conf, "b2p_create failed: %s\n", b2p_last_error(NULL));
    return EXIT_FAILURE;
  }
  conf->ring_pinned = 0;
  if (conf->pin_ring) {
    if (dada_cuda_dbregister(conf->hdu_in) == 0)
      conf->ring_pinned = 1;
    else if (conf->log)
      multilog(conf->log, LOG_WARNING, "input ring could not be page-locked, copies will be staged\n");
  }
  if (conf->log)
    multilog(conf->log, LOG_INFO,
             "baseband2power ready: gpu %d (%d gpu%s), %lu frames/block, %lu frames/integration, ring %s\n",
             conf->ngpu > 1 ? conf->gpus[0] : conf->device_id, conf->ngpu > 1 ? conf->ngpu : 1,
             conf->ngpu > 1 ? "s, channel groups" : "", (unsigned long)conf->ndf_block,
             (unsigned long)conf->ndf_integration, conf->ring_pinned ? "pinned" : "pageable");
  return EXIT_SUCCESS;
}

/* The keys this stage changes on the way through: the stream becomes one float32 per
   channel per integration (header_baseband2power.txt:36-42). */
static int rewrite_header(conf_t *conf, char *hdr)
{
  const double tsamp = (double)conf->ndf_integration * conf->nsamp_df * TSAMP_US;
  const uint64_t nchan = (uint64_t)conf->nchunk * conf->nch_per_chunk;
  int rc = 0;
  rc |= ascii_header_set(hdr, "NBIT", "%d", 32);
  rc |= ascii_header_set(hdr, "NDIM", "%d", 1);
  rc |= ascii_header_set(hdr, "NPOL", "%d", 1);
  rc |= ascii_header_set(hdr, "NCHAN", "%lu", (unsigned long)nchan);
  /* the template carries 88473.6, a factor-10 slip of 1024*1024*27/32 us = 884736 us */
  rc |= ascii_header_set(hdr, "TSAMP", "%.4f", tsamp);
  rc |= ascii_header_set(hdr, "BYTES_PER_SECOND", "%.6f", (double)(nchan * sizeof(float)) / (tsamp * 1e-6));
  return rc;
}

/* Write the `n` oldest finished spectra into output ring blocks (each b2p_wait_output returns
   the oldest integration not yet collected); in group mode the first integrations also tune
   the chunk split to what the links deliver together. */
static int collect_spectra(conf_t *conf, unsigned *due, unsigned n)
{
  for (; n > 0 && *due > 0; --n) {
    uint64_t out_id = 0;
    char *out = ipcio_open_block_write(conf->hdu_out->data_block, &out_id);
    if (!out) {
      STAGE_ERR(conf, "Can not open an output block\n");
      return EXIT_FAILURE;
    }
    const int rc = conf->grp ? b2p_group_wait_output(conf->grp, (float *)out)
                             : b2p_wait_output(conf->ctx, (float *)out);
    if (rc != B2P_OK) {
      STAGE_ERR(conf, "collecting a spectrum failed: %s\n",
                conf->grp ? b2p_group_last_error(conf->grp) : b2p_last_error(conf->ctx));
      return EXIT_FAILURE;
    }
    ipcio_close_block_write(conf->hdu_out->data_block, conf->rbufsz_out);
    conf->nblocks_out++;
    *due -= 1;
  }
  return EXIT_SUCCESS;
}

int do_baseband2power(conf_t *conf)
{
  /* ---- header: in -> out ---- */
  uint64_t hbytes = 0;
  char *hin = ipcbuf_get_next_read(conf->hdu_in->header_block, &hbytes);
  if (!hin) {
    STAGE_ERR(conf, "No header on the input ring\n");
    return EXIT_FAILURE;
  }
  char *hout = ipcbuf_get_next_write(conf->hdu_out->header_block);
  if (!hout) {
    STAGE_ERR(conf, "Can not get the output header block\n");
    return EXIT_FAILURE;
  }
  memcpy(hout, hin, DADA_DEFAULT_HEADER_SIZE);
  hout[DADA_DEFAULT_HEADER_SIZE - 1] = 0;
  ipcbuf_mark_cleared(conf->hdu_in->header_block);
  if (rewrite_header(conf, hout) != 0) {
    STAGE_ERR(conf, "Can not update the output header\n");
    return EXIT_FAILURE;
  }
  if (ipcbuf_mark_filled(conf->hdu_out->header_block, DADA_DEFAULT_HEADER_SIZE) < 0) {
    STAGE_ERR(conf, "Could not mark filled header block\n");
    return EXIT_FAILURE;
  }
  ipcbuf_enable_sod((ipcbuf_t *)conf->hdu_out->data_block, 0, 0);

  /* ---- data ---- */
  /*
   * One block ahead: the copies and kernels of a block are queued without waiting, the ring
   * block goes back as soon as its bytes have left the host, and the spectrum of an integration
   * is collected only after the NEXT block has been queued (when one is already waiting in the
   * ring) — so the host links never idle between blocks.  When the ring is empty (a live stream
   * at 1x) the spectra are written out at once.
   */
  const uint64_t frame_bytes = (uint64_t)conf->nchunk * PKT_BYTES(conf);
  ipcbuf_t *db_in = (ipcbuf_t *)conf->hdu_in->data_block;
  uint64_t in_integration = 0;
  unsigned due = 0; /* integrations closed on the GPU whose spectra are not in the output ring yet */
  for (;;) {
    uint64_t bytes = 0, block_id = 0;
    char *blk = ipcio_open_block_read(conf->hdu_in->data_block, &bytes, &block_id);
    if (!blk) break; /* end of data */
    const double t0 = now_s();
    uint64_t ndf = bytes / frame_bytes, done = 0;
    if (bytes % frame_bytes) conf->nframes_dropped += 1; /* a torn trailing frame */
    while (done < ndf) {
      uint64_t n = conf->ndf_integration - in_integration;
      if (n > ndf - done) n = ndf - done;
      const void *ptr = blk + done * frame_bytes;
      const int closes = in_integration + n == conf->ndf_integration;
      if (closes && due >= 3 && collect_spectra(conf, &due, 1) != EXIT_SUCCESS) return EXIT_FAILURE;
      /* when these frames complete an integration, the kernel of the last staging piece emits
         the spectrum itself (one launch per piece, no separate finish) */
      const int rc = conf->grp ? b2p_group_issue_host(conf->grp, &ptr, n, closes)
                               : b2p_accumulate_host_async(conf->ctx, &ptr, n, closes);
      if (rc != B2P_OK) {
        STAGE_ERR(conf, "queueing a block failed: %s\n",
                  conf->grp ? b2p_group_last_error(conf->grp) : b2p_last_error(conf->ctx));
        return EXIT_FAILURE;
      }
      if (closes) due++;
      done += n;
      in_integration = closes ? 0 : in_integration + n;
    }
    /* spectra of earlier blocks: their kernels ran while this block was being queued */
    const unsigned just_queued = (in_integration == 0 && ndf) ? 1 : 0;
    if (due > just_queued && collect_spectra(conf, &due, due - just_queued) != EXIT_SUCCESS) return EXIT_FAILURE;
    if ((conf->grp ? b2p_group_wait_input(conf->grp) : b2p_wait_input(conf->ctx)) != B2P_OK) {
      STAGE_ERR(conf, "waiting for the H2D copies failed: %s\n",
                conf->grp ? b2p_group_last_error(conf->grp) : b2p_last_error(conf->ctx));
      return EXIT_FAILURE;
    }
    ipcio_close_block_read(conf->hdu_in->data_block, bytes);
    conf->nblocks_in++;
    /* nothing waiting in the ring: do not sit on a finished spectrum until the next block comes.
       Several GPUs: the first integrations are not run ahead either — each one's measured copy
       times move chunks towards the faster links (b2p_group_rebalance) before the next starts. */
    const int tuning = conf->grp && conf->nblocks_out + due <= 8;
    if (due && (tuning || ipcbuf_get_write_count(db_in) <= ipcbuf_get_read_count(db_in)) &&
        collect_spectra(conf, &due, due) != EXIT_SUCCESS)
      return EXIT_FAILURE;
    if (tuning && due == 0 && in_integration == 0) {
      int moved = 0;
      if (b2p_group_rebalance(conf->grp, &moved) == B2P_OK && moved && conf->log) {
        char txt[256] = "";
        for (int i = 0, n = b2p_group_size(conf->grp); i < n; ++i) {
          int dev = 0, first = 0, cnt = 0;
          b2p_group_shard(conf->grp, i, &dev, &first, &cnt);
          snprintf(txt + strlen(txt), sizeof(txt) - strlen(txt), " gpu%d:%d", dev, cnt);
        }
        multilog(conf->log, LOG_INFO, "rebalanced chunks per gpu:%s\n", txt);
      }
    }
    const double dt = now_s() - t0;
    conf->seconds_busy += dt;
    if (!tuning) {
      conf->nblocks_steady++;
      conf->seconds_busy_steady += dt;
      if (dt > conf->seconds_block_max) conf->seconds_block_max = dt;
    }
  }
  if (due && collect_spectra(conf, &due, due) != EXIT_SUCCESS) return EXIT_FAILURE;
  if (in_integration) { /* an incomplete integration has the wrong scale: do not emit it */
    conf->nframes_dropped += in_integration;
    if (conf->grp)
      b2p_group_reset(conf->grp);
    else
      b2p_reset(conf->ctx);
    if (conf->log)
      multilog(conf->log, LOG_WARNING, "dropped a trailing partial integration of %lu frames\n",
               (unsigned long)in_integration);
  }
  if (conf->log)
    multilog(conf->log, LOG_INFO, "END: %lu blocks in, %lu spectra out, %.3f s busy\n",
             (unsigned long)conf->nblocks_in, (unsigned long)conf->nblocks_out, conf->seconds_busy);
  if (conf->log && conf->nblocks_steady)
    multilog(conf->log, LOG_INFO, "STEADY: %lu blocks after the split settled, %.3f s busy, slowest block %.4f s\n",
             (unsigned long)conf->nblocks_steady, conf->seconds_busy_steady, conf->seconds_block_max);
  return EXIT_SUCCESS;
}

int destroy_baseband2power(conf_t *conf)
{
  if (conf->ring_pinned && conf->hdu_in) dada_cuda_dbunregister(conf->hdu_in);
  if (conf->ctx) b2p_destroy(conf->ctx);
  conf->ctx = NULL;
  if (conf->grp) b2p_group_destroy(conf->grp);
  conf->grp = NULL;
  if (conf->hdu_out) {
    dada_hdu_unlock_write(conf->hdu_out); /* raises end-of-data for the downstream reader */
    dada_hdu_disconnect(conf->hdu_out);
    dada_hdu_destroy(conf->hdu_out);
    conf->hdu_out = NULL;
  }
  if (conf->hdu_in) {
    dada_hdu_unlock_read(conf->hdu_in);
    dada_hdu_disconnect(conf->hdu_in);
    dada_hdu_destroy(conf->hdu_in);
    conf->hdu_in = NULL;
  }
  return EXIT_SUCCESS;
}
