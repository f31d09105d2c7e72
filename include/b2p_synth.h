/*
 * b2p_synth.h — counter-based synthetic BMF payload generator (test/bench data only).
 *
 * The reference ships no sample data and no generator, so the
 * synthetic streams used by the tests, bench.py and the b2p_gen tool are
 * defined here, once, as pure integer functions that compile as C, C++ and
 * CUDA.  Every 64-bit payload word (one (t,ch) sample of both polarisations,
 * four int16 components) is a function of (seed, absolute word index, output
 * channel) only, so a host thread, the CPU oracle's caller and a device
 * kernel all produce byte-identical blocks without exchanging data.
 *
 * Block layout being filled (capture.c:540-542, "TFTFP order"):
 *   block[idf][chunk][t][ch][pol][re,im]  int16, big-endian by default.
 */
#ifndef B2P_SYNTH_H
#define B2P_SYNTH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define B2P_HD __host__ __device__ __forceinline__
#else
#define B2P_HD static inline
#endif

/* generator modes */
#define B2P_SYNTH_UNIFORM 0 /* raw hash bits: full int16 range incl. -32768       */
#define B2P_SYNTH_GAUSS   1 /* ~Gaussian, sigma ~512 at channel 0, gain 1+ch/nchan */

B2P_HD uint64_t b2p_synth_mix64(uint64_t z)
{
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

/* Sum of the four bytes of a 32-bit lane, centred: range [-510, 510], sigma 147.8 */
B2P_HD int32_t b2p_synth_ih4(uint32_t u)
{
  return (int32_t)((u & 0xFF) + ((u >> 8) & 0xFF) + ((u >> 16) & 0xFF) + (u >> 24)) - 510;
}

/*
 * The four int16 components of payload word `word` (absolute index in the
 * stream, i.e. block_index*words_per_block + index inside the block) destined
 * for output channel `chan` of `nchan`.
 */
B2P_HD void b2p_synth_word(uint64_t seed, uint64_t word, int chan, int nchan, int mode,
                           int16_t v[4])
{
  uint64_t h1 = b2p_synth_mix64(seed ^ (word * 0xD1342543DE82EF95ULL));
  if (mode == B2P_SYNTH_UNIFORM) {
    v[0] = (int16_t)(h1 & 0xFFFF);
    v[1] = (int16_t)((h1 >> 16) & 0xFFFF);
    v[2] = (int16_t)((h1 >> 32) & 0xFFFF);
    v[3] = (int16_t)((h1 >> 48) & 0xFFFF);
  } else {
    uint64_t h2 = b2p_synth_mix64(h1 ^ 0xA0761D6478BD642FULL);
    /* gain/64: 222/64 = 3.47 -> sigma 512.7 at chan 0, doubling towards chan nchan */
    int32_t g = 222 + (222 * chan) / nchan;
    v[0] = (int16_t)((b2p_synth_ih4((uint32_t)h1) * g) / 64);
    v[1] = (int16_t)((b2p_synth_ih4((uint32_t)(h1 >> 32)) * g) / 64);
    v[2] = (int16_t)((b2p_synth_ih4((uint32_t)h2) * g) / 64);
    v[3] = (int16_t)((b2p_synth_ih4((uint32_t)(h2 >> 32)) * g) / 64);
  }
}

/* Pack the four components into the 8 payload bytes, as a little-endian-host uint64. */
B2P_HD uint64_t b2p_synth_pack(const int16_t v[4], int big_endian)
{
  uint64_t w = 0;
  for (int k = 0; k < 4; ++k) {
    uint16_t u = (uint16_t)v[k];
    if (big_endian) u = (uint16_t)((u >> 8) | (u << 8));
    w |= (uint64_t)u << (16 * k);
  }
  return w;
}

/* Output channel of word `w_in_block` for a geometry (nchunk, nch, nsamp). */
B2P_HD int b2p_synth_chan(uint64_t w_in_block, int nchunk, int nch, int nsamp)
{
  uint64_t wpp = (uint64_t)nsamp * (uint64_t)nch; /* words per packet */
  uint64_t pkt = w_in_block / wpp;
  return (int)(pkt % (uint64_t)nchunk) * nch + (int)((w_in_block % wpp) % (uint64_t)nch);
}

#endif
