/*
 * b2p.h — C ABI of the B200-native baseband -> power kernel library (libb2p.so).
 *
 * This is the drop-in boundary for the hot path of paf-baseband2power: the
 * work the reference planned for kernel.cu / baseband2power.cu and never wrote
 * (kernel.cu:1-7 and baseband2power.cu:1-16 are #includes only; the stage's
 * conf_t at baseband2power.cuh:18-23 is the one declaration that exists).  The
 * reference declares no function-level interface for this path, so each entry
 * point cites the reference artefact whose contract it implements.
 *
 * Plain C: pointers and sizes only, no CUDA or torch types.  `stream`
 * arguments are a cudaStream_t passed as void* (NULL = the context's own
 * compute stream).  Every int-returning call returns 0 (EXIT_SUCCESS) or a
 * non-zero B2P_E* code, mirroring the EXIT_SUCCESS/EXIT_FAILURE convention of
 * init_diskdb/do_diskdb (diskdb.cuh:32-34); the message is kept per context
 * (b2p_last_error).  A context is not thread-safe: one context per
 * (GPU, group of beam streams), used from one host thread at a time.
 *
 * There is no CPU fallback: every compute entry point fails when no CUDA
 * device is usable.
 */
#ifndef B2P_H
#define B2P_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2P_VERSION "0.2.0"

/* error codes */
#define B2P_OK        0
#define B2P_EINVAL    1 /* bad argument / unsupported geometry */
#define B2P_ECUDA     2 /* a CUDA runtime call failed (message holds file:line) */
#define B2P_ENOMEM    3
#define B2P_ESTATE    4 /* call made in the wrong state */

/* accumulation modes (BASELINE.json north_star: exact int64 mode + float mode) */
#define B2P_MODE_EXACT 0 /* uint64 sums, result independent of order: bit-exact */
#define B2P_MODE_FLOAT 1 /* fp32 detect + fp64 integrate, <= 1e-6 relative */

/* fused-kernel variants */
#define B2P_KERNEL_AUTO 0
#define B2P_KERNEL_LDG  1 /* 256-bit coalesced evict-first streaming loads, register pipeline */
#define B2P_KERNEL_TMA  2 /* cp.async.bulk + mbarrier multi-stage shared-memory ring */

#define B2P_MAX_BEAMS 64
#define B2P_MAX_GROUP 16 /* GPUs one beam's channel groups can be spread over */

typedef struct b2p_ctx b2p_ctx;

/*
 * Geometry defaults come from paf-baseband2power.conf: NCHK_NIC 48 (:5),
 * NSAMP_DF 128 (:2), NCHAN 336 (:24) -> 7 channels per chunk; the payload of a
 * data frame is nsamp_df*nch_per_chunk*8 = 7168 B (capture.h:28).
 */
typedef struct b2p_params {
  int device_id;        /* conf_t.device_id, baseband2power.cuh:20 (-d flag) */
  int nchunk;           /* frequency chunks per data frame (48) */
  int nch_per_chunk;    /* channels per chunk (7) */
  int nsamp_df;         /* time samples per data frame (128) */
  int big_endian;       /* 1: BMF wire order (hdr.c:15 byte-swaps; cudautil.cuh:118) */
  float scale;          /* output = (float)sum * scale; 1 = integral (README.md:2),
                           2^-20 = time average (paf_baseband2power.cu:20) */
  int mode;             /* B2P_MODE_* */
  int nbeam;            /* independent beam streams batched in this context (>=1) */
  int kernel;           /* B2P_KERNEL_* */
  int nsplit;           /* time splits per chunk (0 = auto from SM count) */
  uint64_t stage_ndf;   /* data frames per H2D staging piece (0 = default 256) */
  int nstage_bufs;      /* device staging buffers for the host path (0 = default 3) */
  /* channel-group shard (north_star: "beams and channel groups shard independently"): the
     context covers chunks [first_chunk, first_chunk + nchunk) of a stream whose data frames
     hold nchunk_total chunks (0 = nchunk: the whole frame).  Chunks interleave inside every
     frame (capture.c:540-542), so a shard reads rows of nchunk*7168 B at a pitch of
     nchunk_total*7168 B (344 064).  Pointers given to accumulate calls always address frame
     0 of the full stream — the ring block as ipcio_open_block_read returns it. */
  int first_chunk;
  int nchunk_total;
  int resizable;        /* 1: size the buffers for any chunk range of nchunk_total, so that
                           b2p_set_chunk_range can move the range later without reallocating */
} b2p_params;

/* Fill *p with the reference pipeline's defaults (one beam, exact mode, scale 1). */
void b2p_default_params(b2p_params *p);

/*
 * Create / destroy a context: selects the device, allocates the per-beam
 * uint64 accumulators, the partial-sum scratch, streams, events and (lazily)
 * the staging buffers.  Replaces the never-written init_baseband2power /
 * destroy_baseband2power (siblings: diskdb.cuh:32-33).
 */
int  b2p_create(b2p_ctx **out, const b2p_params *p);
void b2p_destroy(b2p_ctx *ctx);

/* Message of the last failure on ctx (ctx == NULL: last b2p_create failure). */
const char *b2p_last_error(const b2p_ctx *ctx);

/*
 * Unpack + detect + integrate `ndf` data frames per beam that already sit in
 * device memory: dptrs[b] points at beam b's frames in ring-block layout
 * block[idf][chunk][t][ch][pol][re,im] (capture.c:540-542).  Adds into the
 * context's accumulators; asynchronous on `stream`.  An integration is any
 * sequence of accumulate calls closed by b2p_finish*.
 *
 * Stream rule: on a caller's stream the kernels obey ordinary stream order (the
 * input may be produced by the caller's previous kernel on that stream).  With
 * stream == NULL they run on the context's own stream, where a fused kernel is
 * allowed to start reading its input while the context's previous kernel is
 * still finishing (programmatic dependent launch) — so the input must be
 * complete when the call is made (synchronise the producer first).
 * On a caller's stream, accumulate_device + finish_device neither synchronise nor
 * allocate, so the pair can be captured into a CUDA graph and replayed.
 */
int b2p_accumulate_device(b2p_ctx *ctx, const void *const *dptrs, uint64_t ndf, void *stream);

/*
 * One whole integration in ONE kernel launch: b2p_accumulate_device of `ndf` frames and
 * b2p_finish_device fused — the last CTA to complete each (beam, chunk) column adds the
 * column's partial sums in fixed order to whatever earlier accumulate calls left in the
 * accumulators, writes out_dev[b*nchan + k] and clears them.  The steady-state call of a
 * stage whose integration is one ring block (NDF 8192, paf-baseband2power.conf:9).
 */
int b2p_integrate_device(b2p_ctx *ctx, const void *const *dptrs, uint64_t ndf, float *out_dev,
                         void *stream);

/*
 * Same, from host memory (a ring-buffer block returned by
 * ipcio_open_block_read): frames are cut into pieces of stage_ndf, copied
 * H2D on a copy stream into rotating staging buffers and reduced on the
 * compute stream, copy k+1 overlapping kernel k.  hptrs[b] should be pinned
 * (b2p_host_alloc / b2p_host_register, the role dada_cuda_dbregister from
 * dada_cuda.h — included at baseband2power.cuh:9 — was to play); pageable
 * memory works but the copies serialise.  Returns once every piece has been
 * consumed (the caller may then release the ring block).
 */
int b2p_accumulate_host(b2p_ctx *ctx, const void *const *hptrs, uint64_t ndf);

/*
 * b2p_accumulate_host and b2p_finish in one call: the kernel of the last staging piece
 * closes the integration (no separate finish launch), then the spectrum is copied to
 * out_host[nbeam*nchan].  What do_baseband2power calls when a block completes an integration.
 */
int b2p_integrate_host(b2p_ctx *ctx, const void *const *hptrs, uint64_t ndf, float *out_host);

/*
 * The same in asynchronous form, so that one host thread can keep several GPUs busy (one
 * context per GPU, e.g. the channel-group shards of a beam): _async queues the copies and
 * kernels and returns at once (finish != 0: the last piece closes the integration and the
 * spectrum's D2H is queued; ndf == 0 with finish: close only); b2p_wait_input returns when
 * the host blocks have been read (the ring block may be released); b2p_wait_output returns
 * the spectrum of the OLDEST finished integration not yet collected — up to 4 may be queued,
 * so a stage can issue the next ring block before it collects the previous spectrum and the
 * host link never idles between blocks.
 */
int b2p_accumulate_host_async(b2p_ctx *ctx, const void *const *hptrs, uint64_t ndf, int finish);
int b2p_wait_input(b2p_ctx *ctx);
int b2p_wait_output(b2p_ctx *ctx, float *out_host);
/* Duration of the H2D copies of the last host call (first copy issued -> last copy done),
   from CUDA events on the copy stream; synchronises with the last copy. */
int b2p_last_h2d_ms(b2p_ctx *ctx, double *ms);

/*
 * Zero-copy variant: the fused kernel reads the pinned, device-mapped host
 * block directly over PCIe (no staging, one launch).  hptrs[b] must be pinned
 * with b2p_host_alloc/b2p_host_register.  Synchronous like b2p_accumulate_host.
 */
int b2p_accumulate_host_mapped(b2p_ctx *ctx, const void *const *hptrs, uint64_t ndf);

/*
 * Close the integration: out[b*nchan + k] = (float)sum[b][k] * scale for
 * k = chunk*nch_per_chunk + ch ascending — the 1344-byte output ring block
 * (NCHAN*NBYTE, paf-baseband2power.conf:24-25; NBIT 32 / NDIM 1 / NPOL 1,
 * header_baseband2power.txt:39-41) — and reset the accumulators.
 * b2p_finish copies to host memory and synchronises; b2p_finish_device writes
 * device memory asynchronously on `stream`.
 */
int b2p_finish(b2p_ctx *ctx, float *out_host);
int b2p_finish_device(b2p_ctx *ctx, float *out_dev, void *stream);

/* Move a resizable context to chunks [first_chunk, first_chunk + nchunk) — between
   integrations only (the accumulators are cleared); no allocation, a few microseconds. */
int b2p_set_chunk_range(b2p_ctx *ctx, int first_chunk, int nchunk);

/* Copy the exact uint64 sums [nbeam*nchan] to the host without resetting (exact mode). */
int b2p_read_sums(b2p_ctx *ctx, uint64_t *sums_host);
/* Zero the accumulators without producing output. */
int b2p_reset(b2p_ctx *ctx);

/* introspection */
int      b2p_nchan(const b2p_ctx *ctx);            /* nchunk*nch_per_chunk */
uint64_t b2p_frame_bytes(const b2p_ctx *ctx);      /* bytes of one data frame of this context's chunks */
uint64_t b2p_source_frame_bytes(const b2p_ctx *ctx); /* bytes of one data frame of the source stream */
int      b2p_first_chunk(const b2p_ctx *ctx);
int      b2p_kernel_in_use(const b2p_ctx *ctx);    /* B2P_KERNEL_LDG / _TMA after AUTO */
int      b2p_nsplit_in_use(const b2p_ctx *ctx);
uint64_t b2p_launch_count(const b2p_ctx *ctx);     /* kernels launched by this context */
void    *b2p_stream(const b2p_ctx *ctx);           /* the context's compute stream */
const char *b2p_version(void);
int      b2p_device_count(void);                   /* cudaGetDeviceCount, paf_baseband2power.cu:88 */
int      b2p_device_info(int device, char *name, size_t name_len, int *sm_count,
                         int *cc_major, int *cc_minor, uint64_t *mem_bytes);

/*
 * Per-launch timing of the fused kernel with CUDA events on the launching
 * stream.  When enabled every fused launch is bracketed by an event pair;
 * b2p_fused_time_ms synchronises, returns the summed duration and the number
 * of launches since the last call, and clears the record.
 */
int b2p_set_timing(b2p_ctx *ctx, int enabled);
int b2p_fused_time_ms(b2p_ctx *ctx, double *sum_ms, uint64_t *launches);

/*
 * Channel-group sharding of beam streams over several GPUs (SURVEY §8e "secondary";
 * one GPU per stage process in the reference, paf_baseband2power.cu:23-26, so a beam
 * there can never use more than one host link).  A group owns one context per GPU;
 * GPU i covers nchunks[i] consecutive chunks (sum = params.nchunk; 0 = GPU unused), the
 * split typically made proportional to each GPU's host-link rate.  Calls take the same
 * full-frame host blocks as b2p_accumulate_host; every GPU copies only its own chunk
 * columns (strided H2D) and the spectra come back as one [nbeam][nchan] array.  Issue on
 * all GPUs first, wait afterwards: one host thread drives all links.
 */
typedef struct b2p_group b2p_group;
int  b2p_group_create(b2p_group **out, const b2p_params *params, const int *devices,
                      const int *nchunks, int ndev);
void b2p_group_destroy(b2p_group *g);
int  b2p_group_accumulate_host(b2p_group *g, const void *const *hptrs, uint64_t ndf);
int  b2p_group_integrate_host(b2p_group *g, const void *const *hptrs, uint64_t ndf, float *out_host);
int  b2p_group_finish(b2p_group *g, float *out_host);
/* the asynchronous trio, as for a single context (b2p_accumulate_host_async & co.) */
int  b2p_group_issue_host(b2p_group *g, const void *const *hptrs, uint64_t ndf, int finish);
int  b2p_group_wait_input(b2p_group *g);
int  b2p_group_wait_output(b2p_group *g, float *out_host);
int  b2p_group_reset(b2p_group *g);
/* Between integrations: move chunks towards the GPUs whose links delivered more during the
   last host call (H2D time per shard, CUDA events), half way per call.  *changed = 1 when the
   split moved (the shard contexts are resizable: the ranges move in place, nothing is rebuilt). */
int  b2p_group_rebalance(b2p_group *g, int *changed);
int  b2p_group_size(const b2p_group *g);                  /* shards with >= 1 chunk */
b2p_ctx *b2p_group_ctx(const b2p_group *g, int i);
int  b2p_group_shard(const b2p_group *g, int i, int *device, int *first_chunk, int *nchunk);
const char *b2p_group_last_error(const b2p_group *g);

/* Pinned host -> device copy rate of each listed GPU with all of them copying at once: the
   link weights for b2p_split_chunks.  Two passes of `reps` copies: the first moves `bytes` per
   copy over every link, the second bytes in proportion to the first pass's rates, so that all
   links stay loaded to the end — what is returned is what they deliver TOGETHER (equal byte
   counts let the slow links finish alone and read too high). */
int b2p_probe_h2d(const int *devices, int n, size_t bytes, int reps, double *gbps_out);
/* Split nchunk chunks over n parts in proportion to weights (NULL = equal), largest
   remainder; counts[] sums to nchunk exactly. */
int b2p_split_chunks(const double *weights, int n, int nchunk, int *counts);

/* host / device memory helpers (for ring registration, tests and the bench) */
int b2p_host_alloc(void **p, size_t bytes);        /* pinned + mapped */
int b2p_host_free(void *p);
int b2p_host_register(void *p, size_t bytes);      /* dada_cuda_dbregister's role */
int b2p_host_unregister(void *p);
int b2p_device_alloc(int device, void **p, size_t bytes);
int b2p_device_free(int device, void *p);
int b2p_memcpy_h2d(int device, void *dst, const void *src, size_t bytes);
int b2p_memcpy_d2h(int device, void *dst, const void *src, size_t bytes);
int b2p_device_sync(int device);

/*
 * Fill device memory with the counter-based synthetic BMF stream of
 * b2p_synth.h (byte-identical to the host generator).  Test/bench data only.
 */
int b2p_synth_fill_device(int device, void *dptr, uint64_t ndf, int nchunk, int nch_per_chunk,
                          int nsamp_df, int big_endian, uint64_t seed, uint64_t first_word,
                          int mode, void *stream);

/* Unit-test hook: unpack all 65536 16-bit patterns with the kernel's PRMT path.
   out_host[v] = sign-extended value decoded from pattern v (both lanes checked). */
int b2p_selftest_unpack(int device, int big_endian, int32_t *out_host /* [65536] */);

#ifdef __cplusplus
}
#endif
#endif /* B2P_H */
