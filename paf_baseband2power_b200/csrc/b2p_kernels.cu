/*
 * b2p_kernels.cu — fused unpack + detect + integrate for PAF BMF baseband on B200 (sm_100a).
 *
 * The reference planned these kernels in kernel.cu and never wrote them
 * (kernel.cu:1-7 is three #includes).  What follows is a new design against the
 * specification in DESIGN.md §1 / SURVEY.md §8a-spec:
 *
 *   in   block[idf][chunk][t][ch][pol][re,im], 16-bit big-endian components,
 *        packet payload (one idf, one chunk) = nsamp*nch*8 B = 7168 B,
 *        packet offset (idf*nchunk + chunk)*7168        (capture.c:540-542)
 *   out  per channel k = chunk*nch + ch: sum over idf,t of Xre^2+Xim^2+Yre^2+Yim^2
 *
 * The op is a streaming reduction, ~1 integer op per input byte, no reuse:
 * HBM-bound, tensor cores irrelevant.  Two fused variants:
 *
 *   LDG  one CTA = (time split, chunk, beam), 448 threads; thread j owns the
 *        j-th 16-byte unit of every packet of its chunk, i.e. payload words 2j
 *        and 2j+1, whose channels (2j)%7 and (2j+1)%7 never change -> two
 *        register accumulators, fully coalesced 7168-B rows, UNROLL independent
 *        128-bit streaming loads in flight per thread.
 *   TMA  persistent CTAs (one per SM): a producer lane streams G consecutive
 *        packets (G*7168 contiguous bytes, one cp.async.bulk) per stage into an
 *        NSTAGE-deep shared-memory ring guarded by full/empty mbarriers; 448
 *        consumer threads read their 16-byte unit of each packet with LDS.128.
 *        The ring keeps ~200 KB per SM in flight without spending registers.
 *
 * Unpack: one PRMT per component does byte swap and sign extension at once.
 * Detect/integrate (exact mode): IMAD squares, a pair of squares fits uint32
 * (<= 2^31), words are accumulated in uint64 — a channel total is <= 2^52, so
 * the sum is exact and independent of order; CTA partials go to global memory
 * and a second tiny kernel adds them in index order (no atomics anywhere).
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b2p_synth.h"
#include "b2p_kernels.cuh"

namespace {

constexpr int kUnitsBmf = 448;        /* 16-byte units per 7168-byte packet */
constexpr int kPktBytes = 7168;
constexpr int kNchBmf = 7;
constexpr int kLdgUnroll = 8;
constexpr int kTmaConsumers = kUnitsBmf;           /* 14 warps */
constexpr int kTmaThreads = kTmaConsumers + 32;    /* + 1 producer warp */
constexpr int kTmaConsumerWarps = kTmaConsumers / 32;

/* ------------------------------------------------------------------ unpack */

__device__ __forceinline__ int32_t prmt(uint32_t a, uint32_t sel)
{
  int32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0u), "r"(sel));
  return d;
}

/* A 32-bit register holds memory bytes m0..m3 in byte lanes 0..3.  Component 0 is
   (m0,m1), component 1 is (m2,m3).  Selector nibble bit 3 replicates the sign of
   the chosen byte, so one PRMT yields the sign-extended int32. */
template <bool BE> __device__ __forceinline__ int32_t unpack_lo(uint32_t r)
{
  return prmt(r, BE ? 0x8801u : 0x9910u);
}
template <bool BE> __device__ __forceinline__ int32_t unpack_hi(uint32_t r)
{
  return prmt(r, BE ? 0xAA23u : 0xBB32u);
}

/* ------------------------------------------------------------- accumulators */

struct AccExact {
  typedef unsigned long long type;
  /* power of one polarisation (two components): <= 2^31, fits uint32 */
  template <bool BE> static __device__ __forceinline__ uint32_t pol(uint32_t r)
  {
    int32_t a = unpack_lo<BE>(r), b = unpack_hi<BE>(r);
    return (uint32_t)(a * a) + (uint32_t)(b * b);
  }
  template <bool BE> static __device__ __forceinline__ void add(type &acc, uint32_t x, uint32_t y)
  {
    acc += (type)pol<BE>(x);
    acc += (type)pol<BE>(y);
  }
};

struct AccFloat {
  typedef double type;
  /* fp32 detect (4 squares, 3 adds: <= 4 ulp_rel = 2.4e-7), fp64 integrate */
  template <bool BE> static __device__ __forceinline__ void add(type &acc, uint32_t x, uint32_t y)
  {
    float a = (float)unpack_lo<BE>(x), b = (float)unpack_hi<BE>(x);
    float c = (float)unpack_lo<BE>(y), d = (float)unpack_hi<BE>(y);
    float p = fmaf(a, a, b * b) + fmaf(c, c, d * d);
    acc += (double)p;
  }
};

template <typename T> __device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p)
{
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

/* frames [f0,f1) of time split `s` out of `n` */
__device__ __forceinline__ void split_range(uint64_t ndf, uint32_t s, uint32_t n, uint64_t &f0,
                                            uint64_t &f1)
{
  f0 = ndf * s / n;
  f1 = ndf * (s + 1) / n;
}

/*
 * CTA reduction of per-thread accumulators a0 (channel c0) and a1 (channel c1)
 * into nch channel totals, written to dst[0..nch).  `red` is [nwarps][nch].
 * Every add is integer (exact mode) so the order is immaterial; in float mode
 * the order is fixed by construction (xor tree, then warps ascending).
 */
template <typename T>
__device__ __forceinline__ void cta_reduce_channels(T a0, T a1, int c0, int c1, int nch, T *red,
                                                    T *dst, int tid, int nthreads)
{
  const int lane = tid & 31, warp = tid >> 5, nwarps = (nthreads + 31) >> 5;
  for (int ch = 0; ch < nch; ++ch) {
    T v = (c0 == ch ? a0 : (T)0) + (c1 == ch ? a1 : (T)0);
    v = warp_sum(v);
    if (lane == 0) red[warp * nch + ch] = v;
  }
  __syncthreads();
  if (tid < nch) {
    T s = 0;
    for (int w = 0; w < nwarps; ++w) s += red[w * nch + tid];
    dst[tid] = s;
  }
}

/* ------------------------------------------------ LDG kernel, BMF geometry */

template <typename Acc, bool BE, int UNROLL>
__global__ void __launch_bounds__(kUnitsBmf, 2)
b2p_fused_ldg_bmf(const B2pBeams beams, const uint32_t nchunk, const uint64_t ndf,
                  typename Acc::type *__restrict__ partials)
{
  typedef typename Acc::type T;
  __shared__ T red[(kUnitsBmf / 32) * kNchBmf];
  const int j = threadIdx.x;
  const uint32_t split = blockIdx.x, nsplit = gridDim.x, chunk = blockIdx.y, beam = blockIdx.z;
  uint64_t f0, f1;
  split_range(ndf, split, nsplit, f0, f1);

  const size_t fstride = (size_t)nchunk * kUnitsBmf; /* uint4 units per data frame */
  const uint4 *p = (const uint4 *)beams.ptr[beam] + (f0 * nchunk + chunk) * kUnitsBmf + j;
  T a0 = 0, a1 = 0;
  uint64_t f = f0;
  for (; f + UNROLL <= f1; f += UNROLL) {
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = ldg_stream(p + u * fstride);
    p += UNROLL * fstride;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      Acc::template add<BE>(a0, v[u].x, v[u].y);
      Acc::template add<BE>(a1, v[u].z, v[u].w);
    }
  }
  for (; f < f1; ++f) {
    uint4 v = ldg_stream(p);
    p += fstride;
    Acc::template add<BE>(a0, v.x, v.y);
    Acc::template add<BE>(a1, v.z, v.w);
  }
  const size_t nchan = (size_t)nchunk * kNchBmf;
  T *dst = partials + ((size_t)beam * nsplit + split) * nchan + (size_t)chunk * kNchBmf;
  cta_reduce_channels<T>(a0, a1, (2 * j) % kNchBmf, (2 * j + 1) % kNchBmf, kNchBmf, red, dst, j,
                         kUnitsBmf);
}

/* ------------------------------------------------ LDG kernel, any geometry */
/* blockDim.x = 32*nch so that 2*blockDim.x is a multiple of nch: the channels of
   a thread's two words are fixed although it walks several units per packet. */
template <typename Acc, bool BE>
__global__ void b2p_fused_ldg_any(const B2pBeams beams, const uint32_t nchunk, const uint32_t nch,
                                  const uint32_t units_per_pkt, const uint64_t ndf,
                                  typename Acc::type *__restrict__ partials)
{
  typedef typename Acc::type T;
  extern __shared__ __align__(16) unsigned char smem_any[];
  T *red = (T *)smem_any;
  const int j = threadIdx.x, nt = blockDim.x;
  const uint32_t split = blockIdx.x, nsplit = gridDim.x, chunk = blockIdx.y, beam = blockIdx.z;
  uint64_t f0, f1;
  split_range(ndf, split, nsplit, f0, f1);
  const uint4 *base = (const uint4 *)beams.ptr[beam];
  T a0 = 0, a1 = 0;
  for (uint64_t f = f0; f < f1; ++f) {
    const uint4 *pkt = base + (f * nchunk + chunk) * units_per_pkt;
#pragma unroll 4
    for (uint32_t u = j; u < units_per_pkt; u += nt) {
      uint4 v = ldg_stream(pkt + u);
      Acc::template add<BE>(a0, v.x, v.y);
      Acc::template add<BE>(a1, v.z, v.w);
    }
  }
  const size_t nchan = (size_t)nchunk * nch;
  T *dst = partials + ((size_t)beam * nsplit + split) * nchan + (size_t)chunk * nch;
  cta_reduce_channels<T>(a0, a1, (2 * j) % nch, (2 * j + 1) % nch, nch, red, dst, j, nt);
}

/* --------------------------------------------------- TMA / mbarrier helpers */

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
/* 1-D bulk async copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP) */
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ uint4 lds128(const void *p)
{
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "r"(smem_u32(p)));
  return r;
}
__device__ __forceinline__ void consumer_bar()
{
  asm volatile("bar.sync 1, %0;" ::"n"(kTmaConsumers) : "memory");
}

/* ------------------------------------------------ TMA kernel, BMF geometry */
/*
 * Work item = (beam, time split, chunk group of G).  Item i belongs to CTA
 * i % gridDim.x.  One stage = the G packets of one data frame of the group:
 * G*7168 contiguous bytes.  Consumer thread j reads unit j of each packet, so
 * it keeps 2 accumulators per packet slot g.
 */
template <int G, int NSTAGE> struct TmaSmem {
  static constexpr int kStageBytes = G * kPktBytes;
  static constexpr int kBarOff = NSTAGE * kStageBytes;
  static constexpr int kRedOff = kBarOff + 2 * NSTAGE * 8;
  static constexpr int kBytes = kRedOff + kTmaConsumerWarps * G * kNchBmf * 8;
};

template <typename Acc, bool BE, int G, int NSTAGE>
__global__ void __launch_bounds__(kTmaThreads, 1)
b2p_fused_tma_bmf(const B2pBeams beams, const uint32_t nchunk, const uint64_t ndf,
                  const uint32_t nsplit, const uint32_t nitems,
                  typename Acc::type *__restrict__ partials)
{
  typedef typename Acc::type T;
  typedef TmaSmem<G, NSTAGE> S;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *full = (uint64_t *)(smem + S::kBarOff);
  uint64_t *empty = full + NSTAGE;
  T *red = (T *)(smem + S::kRedOff);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ngroups = nchunk / G;
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kTmaConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == kTmaConsumerWarps) {
    /* ---- producer: one lane issues every bulk copy of this CTA ---- */
    if (lane == 0) {
      uint32_t it = 0;
      for (uint32_t item = blockIdx.x; item < nitems; item += gridDim.x) {
        const uint32_t group = item % ngroups, rest = item / ngroups;
        const uint32_t split = rest % nsplit, beam = rest / nsplit;
        uint64_t f0, f1;
        split_range(ndf, split, nsplit, f0, f1);
        const unsigned char *src = (const unsigned char *)beams.ptr[beam] +
                                   (f0 * nchunk + (uint64_t)group * G) * kPktBytes;
        const size_t fstride = (size_t)nchunk * kPktBytes;
        for (uint64_t f = f0; f < f1; ++f, ++it, src += fstride) {
          const uint32_t s = it % NSTAGE, k = it / NSTAGE;
          if (k > 0) mbar_wait(&empty[s], (k - 1) & 1);
          mbar_arrive_expect_tx(&full[s], S::kStageBytes);
          bulk_g2s(smem + s * S::kStageBytes, src, S::kStageBytes, &full[s]);
        }
      }
    }
    return;
  }

  /* ---- consumers ---- */
  const int c0 = (2 * tid) % kNchBmf, c1 = (2 * tid + 1) % kNchBmf;
  const size_t nchan = (size_t)nchunk * kNchBmf;
  uint32_t it = 0;
  for (uint32_t item = blockIdx.x; item < nitems; item += gridDim.x) {
    const uint32_t group = item % ngroups, rest = item / ngroups;
    const uint32_t split = rest % nsplit, beam = rest / nsplit;
    uint64_t f0, f1;
    split_range(ndf, split, nsplit, f0, f1);
    T a[G][2];
#pragma unroll
    for (int g = 0; g < G; ++g) a[g][0] = a[g][1] = 0;

    for (uint64_t f = f0; f < f1; ++f, ++it) {
      const uint32_t s = it % NSTAGE, k = it / NSTAGE;
      mbar_wait(&full[s], k & 1);
      const unsigned char *st = smem + s * S::kStageBytes + tid * 16;
      uint4 v[G];
#pragma unroll
      for (int g = 0; g < G; ++g) v[g] = lds128(st + g * kPktBytes);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]); /* data is in registers: free the slot */
#pragma unroll
      for (int g = 0; g < G; ++g) {
        Acc::template add<BE>(a[g][0], v[g].x, v[g].y);
        Acc::template add<BE>(a[g][1], v[g].z, v[g].w);
      }
    }

    /* item done: G*7 channel totals of this (beam, split, group) */
    T *dst = partials + ((size_t)beam * nsplit + split) * nchan + (size_t)group * G * kNchBmf;
#pragma unroll
    for (int g = 0; g < G; ++g)
      for (int ch = 0; ch < kNchBmf; ++ch) {
        T v = (c0 == ch ? a[g][0] : (T)0) + (c1 == ch ? a[g][1] : (T)0);
        v = warp_sum(v);
        if (lane == 0) red[warp * (G * kNchBmf) + g * kNchBmf + ch] = v;
      }
    consumer_bar();
    if (tid < G * kNchBmf) {
      T sum = 0;
      for (int w = 0; w < kTmaConsumerWarps; ++w) sum += red[w * (G * kNchBmf) + tid];
      dst[tid] = sum;
    }
    consumer_bar(); /* red is reused by the next item */
  }
}

/* ----------------------------------------------------- finalize and finish */

/* acc[slot[b]][k] += sum over splits (ascending) of partials[b][split][k] */
template <typename T>
__global__ void b2p_finalize(const B2pBeams beams, const uint32_t nsplit, const uint32_t nchan,
                             const T *__restrict__ partials, T *__restrict__ acc)
{
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x, beam = blockIdx.y;
  if (k >= nchan) return;
  const T *p = partials + (size_t)beam * nsplit * nchan + k;
  T s = 0;
  for (uint32_t i = 0; i < nsplit; ++i) s += p[(size_t)i * nchan];
  acc[(size_t)beams.slot[beam] * nchan + k] += s;
}

/* out = (float)sum * scale (one RN conversion, one fp32 multiply), then clear */
__device__ __forceinline__ float to_f32_rn(unsigned long long v) { return __ull2float_rn(v); }
__device__ __forceinline__ float to_f32_rn(double v) { return __double2float_rn(v); }

template <typename T>
__global__ void b2p_finish_k(T *__restrict__ acc, float *__restrict__ out, const int n,
                             const float scale)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const T v = acc[i];
  acc[i] = 0;
  out[i] = __fmul_rn(to_f32_rn(v), scale);
}

/* ------------------------------------------------------- synthetic stream */

__global__ void b2p_synth_k(uint64_t *__restrict__ out, const uint64_t nwords, const int nchunk,
                            const int nch, const int nsamp, const int big_endian,
                            const uint64_t seed, const uint64_t first_word, const int mode)
{
  const int nchan = nchunk * nch;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += stride) {
    int16_t v[4];
    const int chan = b2p_synth_chan(w, nchunk, nch, nsamp);
    b2p_synth_word(seed, first_word + w, chan, nchan, mode, v);
    out[w] = b2p_synth_pack(v, big_endian);
  }
}

/* --------------------------------------------------------- unpack self-test */

template <bool BE> __global__ void b2p_selftest_unpack_k(int32_t *out)
{
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; /* 16-bit memory pattern */
  if (v >= 65536u) return;
  /* lane 0 holds the pattern, lane 1 its complement, and the other way round */
  const uint32_t r0 = v | ((~v & 0xFFFFu) << 16), r1 = (~v & 0xFFFFu) | (v << 16);
  const int32_t lo = unpack_lo<BE>(r0), hi = unpack_hi<BE>(r1);
  out[v] = (lo == hi) ? lo : (int32_t)0x7FFFFFFF;
}

} /* namespace */

/* ======================================================================== host */

int b2p_tma_group(int nchunk)
{
  if (nchunk % 4 == 0) return 4;
  if (nchunk % 2 == 0) return 2;
  return 1;
}

namespace {
constexpr int kTmaStagesG4 = 7, kTmaStagesG2 = 14, kTmaStagesG1 = 28;

template <typename Acc, bool BE, int G, int NSTAGE>
cudaError_t launch_tma(const B2pLaunch &L, cudaStream_t st)
{
  typedef TmaSmem<G, NSTAGE> S;
  const uint32_t ngroups = L.nchunk / G;
  const uint32_t nitems = (uint32_t)L.nbeam * L.nsplit * ngroups;
  uint32_t grid = (uint32_t)L.sm_count;
  if (grid > nitems) grid = nitems;
  b2p_fused_tma_bmf<Acc, BE, G, NSTAGE><<<grid, kTmaThreads, S::kBytes, st>>>(
      L.beams, (uint32_t)L.nchunk, L.ndf, (uint32_t)L.nsplit, nitems,
      (typename Acc::type *)L.partials);
  return cudaGetLastError();
}

template <typename Acc, bool BE> cudaError_t launch_fused_t(const B2pLaunch &L, cudaStream_t st)
{
  typedef typename Acc::type T;
  const bool bmf = b2p_is_bmf_geometry(L.nch, L.nsamp);
  if (L.kernel == B2P_KERNEL_TMA && bmf) {
    switch (b2p_tma_group(L.nchunk)) {
      case 4: return launch_tma<Acc, BE, 4, kTmaStagesG4>(L, st);
      case 2: return launch_tma<Acc, BE, 2, kTmaStagesG2>(L, st);
      default: return launch_tma<Acc, BE, 1, kTmaStagesG1>(L, st);
    }
  }
  dim3 grid((unsigned)L.nsplit, (unsigned)L.nchunk, (unsigned)L.nbeam);
  if (bmf) {
    b2p_fused_ldg_bmf<Acc, BE, kLdgUnroll><<<grid, kUnitsBmf, 0, st>>>(
        L.beams, (uint32_t)L.nchunk, L.ndf, (T *)L.partials);
  } else {
    const int nt = 32 * L.nch;
    const size_t sh = (size_t)(nt / 32) * L.nch * sizeof(T);
    b2p_fused_ldg_any<Acc, BE><<<grid, nt, sh, st>>>(L.beams, (uint32_t)L.nchunk, (uint32_t)L.nch,
                                                     (uint32_t)(L.nsamp * L.nch / 2), L.ndf,
                                                     (T *)L.partials);
  }
  return cudaGetLastError();
}

template <typename Acc, bool BE, int G, int NSTAGE> cudaError_t configure_tma()
{
  return cudaFuncSetAttribute(b2p_fused_tma_bmf<Acc, BE, G, NSTAGE>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                              TmaSmem<G, NSTAGE>::kBytes);
}
template <typename Acc, bool BE> cudaError_t configure_all()
{
  cudaError_t e;
  if ((e = configure_tma<Acc, BE, 4, kTmaStagesG4>()) != cudaSuccess) return e;
  if ((e = configure_tma<Acc, BE, 2, kTmaStagesG2>()) != cudaSuccess) return e;
  return configure_tma<Acc, BE, 1, kTmaStagesG1>();
}
} /* namespace */

cudaError_t b2p_kernels_configure(void)
{
  cudaError_t e;
  if ((e = configure_all<AccExact, true>()) != cudaSuccess) return e;
  if ((e = configure_all<AccExact, false>()) != cudaSuccess) return e;
  if ((e = configure_all<AccFloat, true>()) != cudaSuccess) return e;
  return configure_all<AccFloat, false>();
}

cudaError_t b2p_launch_fused(const B2pLaunch &L, cudaStream_t st)
{
  if (L.mode == B2P_MODE_FLOAT)
    return L.big_endian ? launch_fused_t<AccFloat, true>(L, st) : launch_fused_t<AccFloat, false>(L, st);
  return L.big_endian ? launch_fused_t<AccExact, true>(L, st) : launch_fused_t<AccExact, false>(L, st);
}

cudaError_t b2p_launch_finalize(const B2pLaunch &L, cudaStream_t st)
{
  const uint32_t nchan = (uint32_t)(L.nchunk * L.nch);
  dim3 grid((nchan + 127) / 128, (unsigned)L.nbeam);
  if (L.mode == B2P_MODE_FLOAT)
    b2p_finalize<double><<<grid, 128, 0, st>>>(L.beams, (uint32_t)L.nsplit, nchan,
                                               (const double *)L.partials, (double *)L.acc);
  else
    b2p_finalize<unsigned long long><<<grid, 128, 0, st>>>(
        L.beams, (uint32_t)L.nsplit, nchan, (const unsigned long long *)L.partials,
        (unsigned long long *)L.acc);
  return cudaGetLastError();
}

cudaError_t b2p_launch_finish(void *acc, float *out, int n, float scale, int mode, cudaStream_t st)
{
  const int grid = (n + 127) / 128;
  if (mode == B2P_MODE_FLOAT)
    b2p_finish_k<double><<<grid, 128, 0, st>>>((double *)acc, out, n, scale);
  else
    b2p_finish_k<unsigned long long><<<grid, 128, 0, st>>>((unsigned long long *)acc, out, n, scale);
  return cudaGetLastError();
}

cudaError_t b2p_launch_synth(void *dptr, uint64_t ndf, int nchunk, int nch, int nsamp,
                             int big_endian, uint64_t seed, uint64_t first_word, int mode,
                             cudaStream_t st)
{
  const uint64_t nwords = ndf * (uint64_t)nchunk * nsamp * nch;
  if (nwords == 0) return cudaSuccess;
  uint64_t blocks = (nwords + 255) / 256;
  if (blocks > 148u * 32u) blocks = 148u * 32u;
  b2p_synth_k<<<(unsigned)blocks, 256, 0, st>>>((uint64_t *)dptr, nwords, nchunk, nch, nsamp,
                                                big_endian, seed, first_word, mode);
  return cudaGetLastError();
}

cudaError_t b2p_launch_selftest_unpack(int big_endian, int32_t *out_dev, cudaStream_t st)
{
  if (big_endian)
    b2p_selftest_unpack_k<true><<<256, 256, 0, st>>>(out_dev);
  else
    b2p_selftest_unpack_k<false><<<256, 256, 0, st>>>(out_dev);
  return cudaGetLastError();
}
