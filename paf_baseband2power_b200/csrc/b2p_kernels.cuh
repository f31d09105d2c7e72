/*
 * b2p_kernels.cuh — internal launch interface between the C ABI (b2p_api.cu)
 * and the sm_100a kernels (b2p_kernels.cu).  Not installed; the public
 * boundary is include/b2p.h.
 */
#ifndef B2P_KERNELS_CUH
#define B2P_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b2p.h"

/* Kernel-parameter block: beam b of this launch reads ptr[b] and adds into
   accumulator row slot[b] of the context. */
struct B2pBeams {
  const void *ptr[B2P_MAX_BEAMS];
  int slot[B2P_MAX_BEAMS];
};

/* Epilogue of a fused launch: the last CTA of each (beam, column) adds the column's
   partial sums to accumulator row slot[beam] and stores the total back (finish = 0) or
   emits (float)total*scale to `out` and clears the accumulator (finish = 1). */
struct B2pFold {
  void *acc;            /* [ctx nbeam][nchan] uint64 (exact) or double (float mode) */
  float *out;           /* [ctx nbeam][nchan] float32, written when finish */
  unsigned int *colcnt; /* [launch nbeam][columns] arrival counters, all zero between launches */
  float scale;
  int finish;
};

struct B2pLaunch {
  B2pBeams beams;
  int nbeam;       /* beams in this launch */
  int nchunk, nch, nsamp; /* nchunk: chunks covered by this launch (a shard's local count) */
  uint64_t fpitch;        /* bytes between successive data frames of the source */
  int big_endian;
  int mode;        /* B2P_MODE_* */
  int kernel;      /* B2P_KERNEL_LDG / B2P_KERNEL_TMA (resolved, not AUTO) */
  int nsplit;      /* time splits per chunk */
  int sm_count;
  int variant;     /* tuning variant of the selected kernel (0 = default) */
  int calib;       /* B2P_CALIB=1: allow the bandwidth-calibration variants (wrong sums) */
  int pdl;         /* launch with programmatic stream serialization */
  int early;       /* input independent of the stream's previous kernel: start before it ends */
  uint64_t ndf;    /* frames per beam in this launch */
  unsigned int *ticket; /* zeroed work-item counter of this launch (TMA kernel); the kernel
                           puts it back to zero itself */
  void *partials;  /* [nbeam][nsplit][nchan] uint64 (exact) or double (float mode) */
  B2pFold fold;
};

/* true when (nch, nsamp) is the BMF geometry the specialised kernels cover */
static inline bool b2p_is_bmf_geometry(int nch, int nsamp) { return nch == 7 && nsamp == 128; }

/* chunks per TMA stage for a given nchunk (0: TMA kernel not applicable) */
int b2p_tma_group(int nchunk);

cudaError_t b2p_launch_fused(const B2pLaunch &L, cudaStream_t st);
/* Stand-alone finish (no fused launch to ride on): out = (float)acc*scale, acc = 0. */
struct B2pFinish {
  int nrows, nchan, mode, pdl;
  void *acc;
  float *out;
  float scale;
};

cudaError_t b2p_launch_finish(const B2pFinish &R, cudaStream_t st);
cudaError_t b2p_launch_synth(void *dptr, uint64_t ndf, int nchunk, int nch, int nsamp,
                             int big_endian, uint64_t seed, uint64_t first_word, int mode,
                             cudaStream_t st);
cudaError_t b2p_launch_selftest_unpack(int big_endian, int32_t *out_dev, cudaStream_t st);
cudaError_t b2p_kernels_configure(void); /* one-time cudaFuncSetAttribute calls */

#endif
