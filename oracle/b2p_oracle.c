/*
 * b2p_oracle.c — CPU oracle for the baseband -> power hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing on the product path may link, import or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker or as the
 * reported CPU baseline.
 *
 * PARITY UNPINNED.  The reference never implemented this path: kernel.cu:1-7
 * and baseband2power.cu:1-16 hold only #includes, paf_baseband2power.cu:92
 * returns before touching data, and the repository has no tests, fixtures or
 * golden vectors.  This file therefore restates the
 * SPECIFICATION assembled from what the reference does declare:
 *
 *   input layout   block[idf][chunk][t][ch][pol][re,im], byte offset of a
 *                  packet payload = (idf*NCHK_NIC + ifreq)*pkt_size
 *                                                   capture.c:540-542, sync.c:157
 *   geometry       NSAMP_DF 128, NPOL_SAMP 2, NDIM_POL 2, NCHK_NIC 48, NDF 8192,
 *                  NCHAN 336, NBYTE 4                paf-baseband2power.conf:2-5,9,24-25
 *                  payload 7168 B                    capture.h:28
 *   sample format  16-bit two's-complement components, big-endian (header words
 *                  are byte-swapped in hdr.c:15,20,23; BSWAP_64 sits unused in
 *                  cudautil.cuh:118-125) — little-endian selectable
 *   output         one float32 per channel, NPOL 1, NDIM 1, NCHAN 336
 *                                                   header_baseband2power.txt:39-42
 *   semantics      detect (|X|^2+|Y|^2) and integrate/average in time
 *                                                   README.md:2, paf_baseband2power.cu:20
 *
 * Exactness: a component square is <= 2^30, a (t,ch) word contributes <= 2^32,
 * one integration has 2^20 words per channel, so a channel sum is <= 2^52 and
 * is exact in uint64.  The float32 output is one round-to-nearest conversion of
 * that exact integer times `scale` (1 for the integral, 2^-20 for the mean), so
 * it does not depend on accumulation order.
 *
 * What pins it in place of reference vectors: the hand-computable known-answer
 * tests of SURVEY.md §8c (tests/test_oracle_kat.py), an independent numpy
 * restatement (oracle/b2p_oracle_np.py), committed fixtures generated from
 * both (tests/golden/), and — for the byte order alone — vectors produced by the
 * reference's own BSWAP_64 macro (tests/golden/bswap64_vectors.json).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/b2p_synth.h"

typedef struct {
  int nchunk;        /* NCHK_NIC, 48 */
  int nch_per_chunk; /* NCHAN / NCHK_NIC, 7 */
  int nsamp_df;      /* NSAMP_DF, 128 */
  int big_endian;    /* 1 = BMF wire order */
} b2p_oracle_geom;

#define B2P_ORACLE_MAXCH 4096

static inline int32_t load_i16(const uint8_t *p, int big_endian)
{
  uint16_t u = big_endian ? (uint16_t)((p[0] << 8) | p[1]) : (uint16_t)((p[1] << 8) | p[0]);
  return (int32_t)(int16_t)u;
}

/* Sum of the four squared components of one 8-byte (t,ch) word. */
static inline uint64_t word_power(const uint8_t *p, int big_endian)
{
  int32_t a = load_i16(p, big_endian), b = load_i16(p + 2, big_endian);
  int32_t c = load_i16(p + 4, big_endian), d = load_i16(p + 6, big_endian);
  return (uint64_t)(uint32_t)(a * a) + (uint64_t)(uint32_t)(b * b) +
         (uint64_t)(uint32_t)(c * c) + (uint64_t)(uint32_t)(d * d);
}

/* One packet payload: [t][ch] words; adds into sums7[ch]. */
static inline void packet_power(const uint8_t *pkt, const b2p_oracle_geom *g, uint64_t *sums_ch)
{
  const int nch = g->nch_per_chunk;
  for (int t = 0; t < g->nsamp_df; ++t)
    for (int ch = 0; ch < nch; ++ch)
      sums_ch[ch] += word_power(pkt + ((size_t)t * nch + ch) * 8, g->big_endian);
}

int b2p_oracle_nchan(const b2p_oracle_geom *g) { return g->nchunk * g->nch_per_chunk; }

uint64_t b2p_oracle_frame_bytes(const b2p_oracle_geom *g)
{
  return (uint64_t)g->nchunk * g->nsamp_df * g->nch_per_chunk * 8u;
}

/*
 * sums[nchan] += per-channel power of `ndf` data frames.  Single thread, loops
 * in the layout's own order [idf][chunk][t][ch].
 */
int b2p_oracle_accumulate(const uint8_t *block, uint64_t ndf, const b2p_oracle_geom *g,
                          uint64_t *sums)
{
  if (!block || !g || !sums || g->nchunk <= 0 || g->nch_per_chunk <= 0 || g->nsamp_df <= 0)
    return 1;
  const size_t pkt = (size_t)g->nsamp_df * g->nch_per_chunk * 8;
  for (uint64_t idf = 0; idf < ndf; ++idf)
    for (int c = 0; c < g->nchunk; ++c)
      packet_power(block + (idf * g->nchunk + c) * pkt, g, sums + (size_t)c * g->nch_per_chunk);
  return 0;
}

/*
 * Same result with all host threads: frames are split over threads, each thread
 * keeps private uint64 sums, the join adds them in thread order (integers, so
 * the order is immaterial).  This is the CPU baseline bench.py times.
 */
int b2p_oracle_accumulate_omp(const uint8_t *block, uint64_t ndf, const b2p_oracle_geom *g,
                              uint64_t *sums, int nthreads)
{
  if (!block || !g || !sums || g->nchunk <= 0 || g->nch_per_chunk <= 0 || g->nsamp_df <= 0)
    return 1;
  const int nchan = b2p_oracle_nchan(g);
  if (nchan > B2P_ORACLE_MAXCH) return 2;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  uint64_t *priv = (uint64_t *)calloc((size_t)nthreads * nchan, sizeof(uint64_t));
  if (!priv) return 3;
  const uint64_t fb = b2p_oracle_frame_bytes(g);
  /* runs of 32 frames handed out dynamically: a vCPU that loses its time slice (shared
     hosts) delays one run, not a whole 1/nthreads share of the block */
  const uint64_t run = 32, nruns = (ndf + run - 1) / run;
#pragma omp parallel num_threads(nthreads)
  {
    const int tid = omp_get_thread_num();
#pragma omp for schedule(dynamic, 1)
    for (uint64_t r = 0; r < nruns; ++r) {
      const uint64_t f0 = r * run, n = (ndf - f0 < run) ? ndf - f0 : run;
      b2p_oracle_accumulate(block + f0 * fb, n, g, priv + (size_t)tid * nchan);
    }
  }
  for (int t = 0; t < nthreads; ++t)
    for (int k = 0; k < nchan; ++k) sums[k] += priv[(size_t)t * nchan + k];
  free(priv);
  return 0;
#else
  (void)nthreads;
  return b2p_oracle_accumulate(block, ndf, g, sums);
#endif
}

int b2p_oracle_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* out[k] = (float)sums[k] * scale — one RN conversion, then an fp32 multiply. */
void b2p_oracle_finish(const uint64_t *sums, int nchan, float scale, float *out)
{
  for (int k = 0; k < nchan; ++k) {
    volatile float f = (float)sums[k]; /* volatile: forbid fused/extended evaluation */
    out[k] = f * scale;
  }
}

/*
 * "Float accumulation" restatement: double accumulators in layout order.  With
 * every partial sum < 2^53 this is exact too, which is the point — it shows the
 * answer is order independent.  The float-mode GPU kernel is checked against
 * the exact value to <= 1e-6 relative.
 */
int b2p_oracle_accumulate_f64(const uint8_t *block, uint64_t ndf, const b2p_oracle_geom *g,
                              double *sums)
{
  if (!block || !g || !sums) return 1;
  const int nch = g->nch_per_chunk;
  const size_t pkt = (size_t)g->nsamp_df * nch * 8;
  for (uint64_t idf = 0; idf < ndf; ++idf)
    for (int c = 0; c < g->nchunk; ++c) {
      const uint8_t *p = block + (idf * g->nchunk + c) * pkt;
      for (int t = 0; t < g->nsamp_df; ++t)
        for (int ch = 0; ch < nch; ++ch) {
          const uint8_t *w = p + ((size_t)t * nch + ch) * 8;
          double s = 0.0;
          for (int k = 0; k < 4; ++k) {
            double x = (double)load_i16(w + 2 * k, g->big_endian);
            s += x * x;
          }
          sums[(size_t)c * nch + ch] += s;
        }
    }
  return 0;
}

/*
 * Naive fp32 running sum in layout order — informational only: it documents
 * how far a careless single-accumulator float kernel would drift (~1e-4), which
 * is why the product's float mode keeps short fp32 runs and joins them wider.
 */
int b2p_oracle_accumulate_f32_naive(const uint8_t *block, uint64_t ndf, const b2p_oracle_geom *g,
                                    float *sums)
{
  if (!block || !g || !sums) return 1;
  const int nch = g->nch_per_chunk;
  const size_t pkt = (size_t)g->nsamp_df * nch * 8;
  for (uint64_t idf = 0; idf < ndf; ++idf)
    for (int c = 0; c < g->nchunk; ++c) {
      const uint8_t *p = block + (idf * g->nchunk + c) * pkt;
      for (int t = 0; t < g->nsamp_df; ++t)
        for (int ch = 0; ch < nch; ++ch) {
          const uint8_t *w = p + ((size_t)t * nch + ch) * 8;
          for (int k = 0; k < 4; ++k) {
            float x = (float)load_i16(w + 2 * k, g->big_endian);
            sums[(size_t)c * nch + ch] += x * x;
          }
        }
    }
  return 0;
}

/*
 * Fill `ndf` frames with the counter-based synthetic stream of
 * include/b2p_synth.h.  `first_word` is the absolute index of the block's
 * first word in the stream (block_index * words_per_block for a stream of
 * equal blocks).  Threaded over frames when built with OpenMP.
 */
int b2p_oracle_synth_fill(uint8_t *block, uint64_t ndf, const b2p_oracle_geom *g, uint64_t seed,
                          uint64_t first_word, int mode)
{
  if (!block || !g) return 1;
  const int nchan = b2p_oracle_nchan(g);
  const uint64_t wpf = b2p_oracle_frame_bytes(g) / 8;
  uint64_t *out = (uint64_t *)block;
#pragma omp parallel for schedule(static)
  for (int64_t idf = 0; idf < (int64_t)ndf; ++idf)
    for (uint64_t k = 0; k < wpf; ++k) {
      uint64_t w = (uint64_t)idf * wpf + k;
      int16_t v[4];
      int chan = b2p_synth_chan(w, g->nchunk, g->nch_per_chunk, g->nsamp_df);
      b2p_synth_word(seed, first_word + w, chan, nchan, mode, v);
      uint64_t packed = b2p_synth_pack(v, g->big_endian);
      memcpy(&out[w], &packed, 8);
    }
  return 0;
}
