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

/* For the cross-CTA reduce: lb[row] = index of accumulator row `row` among the beams
   of the pending fused launch, or -1 when the row has no pending partial sums. */
struct B2pSlots {
  int lb[B2P_MAX_BEAMS];
};

struct B2pLaunch {
  B2pBeams beams;
  int nbeam;       /* beams in this launch */
  int nchunk, nch, nsamp;
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
  unsigned int *ticket; /* zeroed work-item counter of this launch (TMA kernel) */
  void *partials;  /* [nbeam][nsplit][nchan] uint64 (exact) or double (float mode) */
  void *acc;       /* [ctx nbeam][nchan]     uint64 (exact) or double (float mode) */
};

/* true when (nch, nsamp) is the BMF geometry the specialised kernels cover */
static inline bool b2p_is_bmf_geometry(int nch, int nsamp) { return nch == 7 && nsamp == 128; }

/* chunks per TMA stage for a given nchunk (0: TMA kernel not applicable) */
int b2p_tma_group(int nchunk);

cudaError_t b2p_launch_fused(const B2pLaunch &L, cudaStream_t st);
struct B2pReduce {
  B2pSlots slots;
  int nrows;       /* accumulator rows of the context (its nbeam) */
  int nsplit;      /* splits of the pending launch */
  int nchan;
  int mode;
  int finish;      /* 1: emit float32 spectrum and clear; 0: fold partials into acc */
  int pdl;
  const void *partials;
  void *acc;
  float *out;
  float scale;
  unsigned int *ticket; /* counter of the pending fused launch, reset here (NULL: none) */
};

cudaError_t b2p_launch_reduce(const B2pReduce &R, cudaStream_t st);
cudaError_t b2p_launch_synth(void *dptr, uint64_t ndf, int nchunk, int nch, int nsamp,
                             int big_endian, uint64_t seed, uint64_t first_word, int mode,
                             cudaStream_t st);
cudaError_t b2p_launch_selftest_unpack(int big_endian, int32_t *out_dev, cudaStream_t st);
cudaError_t b2p_kernels_configure(void); /* one-time cudaFuncSetAttribute calls */

#endif
