/*
 * bmf_packet.h — the 64-byte BMF data-frame header and the stream constants.
 *
 * Bit layout from the reference decoder hdr.c:10-28 (three big-endian 64-bit
 * words; the rest of the 64 bytes is unused by the pipeline):
 *   word0  bit 63 valid | bits 61..32 sec (30 bits) | bits 31..0 idf
 *   word1  bits 31..26 epoch
 *   word2  bits 31..16 freq (integer MHz of the chunk) | bits 15..0 beam
 * Decoding is pinned by tests/golden/bmf_hdr_vectors.json, produced by the
 * reference's own hdr.c.  Stream constants: capture.h:19-32.
 */
#ifndef BMF_PACKET_H
#define BMF_PACKET_H

#include <stdint.h>
#include <string.h>

#define BMF_DF_SIZE   7232    /* data frame with header   (capture.h:27) */
#define BMF_DT_SIZE   7168    /* payload                  (capture.h:28) */
#define BMF_HDR_SIZE  64      /*                          (capture.h:29) */
#define BMF_TDF_SEC   1.08E-4 /* one data frame in time   (capture.h:30) */
#define BMF_PRD_SEC   27      /* streaming period         (capture.h:31) */
#define BMF_NDF_PRD   250000  /* data frames per period   (capture.h:32) */
#define BMF_NCHK_NIC  48      /* chunks per NIC           (capture.h:20) */
#define BMF_NCHK_BMF  6       /* chunks per BMF board     (capture.h:21) */
#define BMF_NPORT_NIC 6       /*                          (capture.h:23) */
#define BMF_PORT_BASE 17100   /*                          (capture.h:24) */

typedef struct bmf_hdr_t {
  int valid;
  uint64_t idf;
  uint64_t sec;
  int epoch;
  int beam;
  double freq;
} bmf_hdr_t;

static inline uint64_t bmf_be64(const unsigned char *p)
{
  uint64_t v = 0;
  for (int i = 0; i < 8; ++i) v = (v << 8) | p[i];
  return v;
}

static inline void bmf_put_be64(unsigned char *p, uint64_t v)
{
  for (int i = 7; i >= 0; --i) {
    p[i] = (unsigned char)(v & 0xFF);
    v >>= 8;
  }
}

static inline void bmf_hdr_decode(const void *df, bmf_hdr_t *h)
{
  const unsigned char *p = (const unsigned char *)df;
  const uint64_t w0 = bmf_be64(p), w1 = bmf_be64(p + 8), w2 = bmf_be64(p + 16);
  h->idf = w0 & 0xFFFFFFFFull;
  h->sec = (w0 >> 32) & 0x3FFFFFFFull;
  h->valid = (int)(w0 >> 63);
  h->epoch = (int)((w1 & 0xFC000000ull) >> 26);
  h->freq = (double)((w2 & 0xFFFF0000ull) >> 16);
  h->beam = (int)(w2 & 0xFFFFull);
}

static inline void bmf_hdr_encode(void *df, const bmf_hdr_t *h)
{
  unsigned char *p = (unsigned char *)df;
  memset(p, 0, BMF_HDR_SIZE);
  bmf_put_be64(p, ((uint64_t)(h->valid & 1) << 63) | ((h->sec & 0x3FFFFFFFull) << 32) | (h->idf & 0xFFFFFFFFull));
  bmf_put_be64(p + 8, ((uint64_t)(h->epoch & 0x3F)) << 26);
  bmf_put_be64(p + 16, (((uint64_t)h->freq & 0xFFFFull) << 16) | ((uint64_t)h->beam & 0xFFFFull));
}

/* Frames elapsed from (sec0, idf0) to (sec, idf): capture.c:562-568 (acquire_idf), which
   divides the second difference by TDF_SEC in floating point; sec advances in whole
   periods of 27 s = 250000 frames, so the same value is computed here in integers. */
static inline int64_t bmf_frames_since(uint64_t sec, uint64_t idf, uint64_t sec0, uint64_t idf0)
{
  const int64_t dsec = (int64_t)sec - (int64_t)sec0;
  return (int64_t)idf + dsec * BMF_NDF_PRD / BMF_PRD_SEC - (int64_t)idf0;
}

/* Chunk index of a packet from its source address a.b.X.Y: X = BMF board 1..8, Y = link
   1..12, odd links carry the first beam set: capture.c:570-584 (acquire_ifreq). */
static inline int bmf_chunk_of_source(unsigned char x, unsigned char y)
{
  return ((int)x - 1) * BMF_NCHK_BMF + ((int)y + 1) / 2 - 1;
}

/* The inverse, for the replayer: source address bytes of chunk c (odd links). */
static inline void bmf_source_of_chunk(int c, unsigned char *x, unsigned char *y)
{
  *x = (unsigned char)(c / BMF_NCHK_BMF + 1);
  *y = (unsigned char)(2 * (c % BMF_NCHK_BMF) + 1);
}

#endif
