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
 *   LDG  one CTA = (chunk, time split, beam), 224 threads; thread j owns the
 *        j-th 32-byte unit of every packet of its chunk, i.e. payload words
 *        4j..4j+3, whose channels (4j+k)%7 never change -> four register
 *        accumulators, fully coalesced 7168-B rows, 4 independent 256-bit
 *        streaming loads (LDG.E.256, L2 evict-first) in flight per thread, 4
 *        CTAs per SM.  A 128-bit form of the same scheme serves other
 *        geometries and is kept as a tuning point (B2P_VARIANT=1).
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
 * and the LAST CTA to finish a (beam, chunk) column — found with one arrival
 * counter per column — adds the column's partial sums in fixed split order and
 * either folds them into the running accumulator or emits the float32 spectrum:
 * one launch per integration.  Sums never go through atomics.
 *
 * Launch chaining: every kernel is launched with programmatic stream
 * serialization (PDL).  A fused kernel releases its dependents at once and only
 * waits for its predecessor (griddepcontrol.wait) right before it writes its
 * partial sums, so on the context's own stream the head of integration N+1
 * overlaps the tail of integration N and the launch gaps disappear.
 *
 * Channel-group shards: `nchunk` is the number of chunks this launch covers and
 * `fpitch` the byte distance between successive data frames of the source, so a
 * launch can read chunks [c0, c0+nchunk) straight out of full 48-chunk frames
 * (pitch 344 064 B) or out of a compact staging copy (pitch nchunk*7168).
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/b2p_synth.h"
#include "b2p_kernels.cuh"

namespace {

constexpr int kUnitsBmf = 448;        /* 16-byte units per 7168-byte packet */
constexpr int kPktBytes = 7168;
constexpr int kNchBmf = 7;
constexpr int kTmaConsumers = kUnitsBmf;           /* 14 warps */
constexpr int kTmaThreads = kTmaConsumers + 32;    /* + 1 producer warp */
constexpr int kTmaConsumerWarps = kTmaConsumers / 32;

/* ------------------------------------------------- programmatic launch (PDL) */

__device__ __forceinline__ void pdl_release_dependents()
{
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait_predecessor()
{
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

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

/* Blackwell's 256-bit global load carries an L2 eviction priority: the block is
   read exactly once, so its lines are marked evict-first (SASS LDG.E.NA.EFL2.256). */
struct u32x8 {
  uint32_t r[8];
};
__device__ __forceinline__ u32x8 ldg256_stream(const void *p)
{
  u32x8 v;
  asm volatile(
      "ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(v.r[0]), "=r"(v.r[1]), "=r"(v.r[2]), "=r"(v.r[3]), "=r"(v.r[4]), "=r"(v.r[5]),
        "=r"(v.r[6]), "=r"(v.r[7])
      : "l"(p));
  return v;
}

/*
 * Time splits are tapered: the first 3/4 of the splits get ranges of weight 4, the next
 * 1/8 weight 2, the last 1/8 weight 1.  CTAs (and TMA work items) are issued in split
 * order, so the last wave consists of short ranges and the SMs run dry together: +0.3 %
 * on a chained stream, +0.4 % on an isolated launch (profiles/r01_variant_sweep.md).
 */
__device__ __forceinline__ uint64_t taper_cum(uint32_t s, uint32_t a, uint32_t b)
{
  const uint32_t s1 = s < a ? s : a;
  const uint32_t s2 = s > a ? (s - a < b ? s - a : b) : 0;
  const uint32_t s3 = s > a + b ? s - a - b : 0;
  return 4ull * s1 + 2ull * s2 + s3;
}

/* frames [f0,f1) of time split `s` out of `n` */
__device__ __forceinline__ void split_range(uint64_t ndf, uint32_t s, uint32_t n, uint64_t &f0,
                                            uint64_t &f1)
{
  if (n >= 16) {
    const uint32_t a = n - n / 4, b = n / 8;
    const uint64_t W = taper_cum(n, a, b);
    f0 = ndf * taper_cum(s, a, b) / W;
    f1 = ndf * taper_cum(s + 1, a, b) / W;
    return;
  }
  f0 = ndf * s / n;
  f1 = ndf * (s + 1) / n;
}

/*
 * CTA reduction of NA per-thread accumulators a[i] (channel c[i]) into nch
 * channel totals, written to dst[0..nch).  `red` is [nwarps][nch].  Every add
 * is integer (exact mode) so the order is immaterial; in float mode the order is
 * fixed by construction (xor tree, then warps ascending).
 */
template <typename T, int NA>
__device__ __forceinline__ void cta_reduce_n(const T (&a)[NA], const int (&c)[NA], int nch, T *red,
                                             T *dst, int tid, int nthreads)
{
  const int lane = tid & 31, warp = tid >> 5, nwarps = (nthreads + 31) >> 5;
  for (int ch = 0; ch < nch; ++ch) {
    T v = 0;
#pragma unroll
    for (int i = 0; i < NA; ++i) v += (c[i] == ch ? a[i] : (T)0);
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

/* ------------------------------------------- column epilogue (cross-CTA reduce) */

__device__ __forceinline__ float to_f32_rn(unsigned long long v) { return __ull2float_rn(v); }
__device__ __forceinline__ float to_f32_rn(double v) { return __double2float_rn(v); }

/* barrier over the threads that take part in an epilogue: the whole CTA (ID 0) or the
   consumer threads of the TMA kernel (named barrier ID, COUNT threads) */
template <int ID, int COUNT> __device__ __forceinline__ void epi_bar()
{
  if (ID == 0)
    __syncthreads();
  else
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}

/*
 * Arrival: every CTA that has stored the partial sums of one (beam, column) item calls
 * this with all participating threads; the partial stores must precede it in program
 * order in the threads that made them.  Returns true (to all threads) in the one CTA that
 * arrived last — the partial sums of all nsplit splits of the column are then visible.
 * One acq_rel atomic by one thread after a CTA barrier: the barrier puts the CTA's stores
 * before the atomic (release, cumulative), the atomic's acquire side and the second barrier
 * put the other CTAs' stores before the reads of the last CTA — no per-thread fences.
 */
template <int ID, int COUNT>
__device__ __forceinline__ bool column_arrive(unsigned int *cnt, uint32_t nsplit, uint32_t *flag,
                                              int tid)
{
  epi_bar<ID, COUNT>();
  if (tid == 0) {
    unsigned int old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(cnt) : "memory");
    *flag = (old == nsplit - 1u) ? 1u : 0u;
  }
  epi_bar<ID, COUNT>();
  const bool last = *flag != 0u;
  if (last) __threadfence(); /* once per column: belt and braces on the acquire side */
  return last;
}

/*
 * Fold of one column by its last CTA: `ncol` channels starting at channel col0.  Thread t <
 * LP*ncol takes channel t % ncol and sums splits l, l+LP, ... (l = t / ncol) in ascending
 * order; ncol threads then add the LP lane sums in ascending order and the running
 * accumulator — a fixed order whoever arrives last.  finish: emit (float)total*scale (one RN
 * conversion, one fp32 multiply) and clear the accumulator; otherwise store the total back.
 * Partial sums were written by other SMs: read them at L2 (ld.global.cg), eight loads in
 * flight per thread so the fold costs about one L2 round trip, not one per split.
 */
template <typename T, int ID, int COUNT>
__device__ __forceinline__ void column_fold(const T *__restrict__ partials, T *fold, const B2pFold &F,
                                            uint32_t lbeam, int row, uint32_t nsplit, size_t nchan,
                                            uint32_t col0, int ncol, int tid, int nt,
                                            unsigned int *cnt)
{
  int LP = nt / ncol;
  if (LP > 32) LP = 32;
  T *acc = (T *)F.acc;
  T prev = 0;
  if (tid < ncol) prev = acc[(size_t)row * nchan + col0 + tid]; /* in flight under the fold */
  if (tid < LP * ncol) {
    const int ch = tid % ncol, l = tid / ncol;
    const T *p = partials + (size_t)lbeam * nsplit * nchan + col0 + ch;
    T v = 0;
    for (uint32_t sp = l; sp < nsplit; sp += 8u * LP) {
      T x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t q = sp + (uint32_t)u * LP;
        x[u] = q < nsplit ? __ldcg(p + (size_t)q * nchan) : (T)0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) v += x[u]; /* ascending split order */
    }
    fold[tid] = v;
  }
  epi_bar<ID, COUNT>();
  if (tid < ncol) {
    T tot = 0;
    for (int l = 0; l < LP; ++l) tot += fold[l * ncol + tid];
    const size_t idx = (size_t)row * nchan + col0 + tid;
    tot += prev;
    if (F.finish) {
      F.out[idx] = __fmul_rn(to_f32_rn(tot), F.scale);
      acc[idx] = 0;
    } else {
      acc[idx] = tot;
    }
  }
  if (tid == 0) *cnt = 0u; /* clean for the next launch (and for a CUDA-graph replay) */
  epi_bar<ID, COUNT>();    /* `fold` may be reused by the caller */
}

/* Out of line for the persistent TMA kernel: the fold runs a handful of times per CTA and must
   not cost the streaming loop registers (480 threads leave 128 per thread). */
template <typename T, int ID, int COUNT>
__device__ __noinline__ void column_fold_call(const T *partials, T *fold, const B2pFold *F,
                                              uint32_t lbeam, int row, uint32_t nsplit, uint32_t nchan,
                                              uint32_t col0, int ncol, int tid, unsigned int *cnt)
{
  __threadfence();
  column_fold<T, ID, COUNT>(partials, fold, *F, lbeam, row, nsplit, (size_t)nchan, col0, ncol, tid, COUNT, cnt);
}

/* ------------------------------------- LDG.256 kernel, BMF geometry (default) */
/*
 * 224 threads cover a 7168-byte packet with 32 bytes each = payload words
 * 4j..4j+3, channels (4j+k)%7: four register accumulators per thread.  UF frames
 * are unrolled, so UF independent 256-bit loads are in flight per thread.  The
 * chunk index is blockIdx.x, so CTAs that run together read adjacent packets of
 * the same frames.  `early`: the input does not depend on the preceding kernel of
 * the stream, so only the partial-sum store has to wait for it.
 */
template <typename Acc, bool BE, int UF, int MINB>
__global__ void __launch_bounds__(224, MINB)
b2p_fused_ldg256_bmf(const B2pBeams beams, const uint32_t nchunk, const uint64_t fpitch,
                     const uint64_t ndf, const int early, typename Acc::type *__restrict__ partials,
                     const B2pFold fold)
{
  typedef typename Acc::type T;
  constexpr int THREADS = 224;
  __shared__ T red[THREADS]; /* [7 warps][7 ch] for the CTA sum, [32 lanes][7 ch] for the fold */
  __shared__ uint32_t last_flag;
  pdl_release_dependents();
  if (!early) pdl_wait_predecessor();
  const int j = threadIdx.x;
  const uint32_t chunk = blockIdx.x, split = blockIdx.y, nsplit = gridDim.y, beam = blockIdx.z;
  uint64_t f0, f1;
  split_range(ndf, split, nsplit, f0, f1);
  const size_t fstride = (size_t)fpitch; /* bytes per data frame of the source */
  const unsigned char *p =
      (const unsigned char *)beams.ptr[beam] + f0 * fstride + (size_t)chunk * kPktBytes + j * 32;
  T a[4] = {0, 0, 0, 0};
  uint64_t f = f0;
  for (; f + UF <= f1; f += UF) {
    u32x8 v[UF];
#pragma unroll
    for (int u = 0; u < UF; ++u) v[u] = ldg256_stream(p + u * fstride);
    p += UF * fstride;
#pragma unroll
    for (int u = 0; u < UF; ++u)
#pragma unroll
      for (int k = 0; k < 4; ++k) Acc::template add<BE>(a[k], v[u].r[2 * k], v[u].r[2 * k + 1]);
  }
  for (; f < f1; ++f) {
    u32x8 v = ldg256_stream(p);
    p += fstride;
#pragma unroll
    for (int k = 0; k < 4; ++k) Acc::template add<BE>(a[k], v.r[2 * k], v.r[2 * k + 1]);
  }
  const size_t nchan = (size_t)nchunk * kNchBmf;
  T *dst = partials + ((size_t)beam * nsplit + split) * nchan + (size_t)chunk * kNchBmf;
  const int c[4] = {(4 * j) % kNchBmf, (4 * j + 1) % kNchBmf, (4 * j + 2) % kNchBmf,
                    (4 * j + 3) % kNchBmf};
  /* the previous kernel of the stream has finished with `partials`, the counters and acc */
  if (early) pdl_wait_predecessor();
  cta_reduce_n<T, 4>(a, c, kNchBmf, red, dst, j, THREADS);
  unsigned int *cnt = fold.colcnt + beam * nchunk + chunk;
  if (!column_arrive<0, 0>(cnt, nsplit, &last_flag, j)) return;
  column_fold<T, 0, 0>(partials, red, fold, beam, beams.slot[beam], nsplit, nchan, chunk * kNchBmf,
                       kNchBmf, j, THREADS, cnt);
}

/* ------------------------------------------------ LDG.128 kernel, any geometry */
/*
 * blockDim.x is a multiple of nch (32*nch, or 448 for the BMF packet), so
 * 2*blockDim.x is too: the channels of a thread's two words are fixed although it
 * may walk several 16-byte units per packet.  UF frames unrolled.  CALIB replaces
 * unpack/detect by an XOR (bandwidth calibration only, B2P_CALIB=1).
 */
template <typename Acc, bool BE, int UF, bool CALIB>
__global__ void b2p_fused_ldg128(const B2pBeams beams, const uint32_t nchunk, const uint32_t nch,
                                 const uint32_t units_per_pkt, const uint64_t fpitch,
                                 const uint64_t ndf, const int early,
                                 typename Acc::type *__restrict__ partials, const B2pFold fold)
{
  typedef typename Acc::type T;
  extern __shared__ __align__(16) unsigned char smem_any[];
  T *red = (T *)smem_any; /* max([nwarps][nch], [32][nch]) elements */
  __shared__ uint32_t last_flag;
  pdl_release_dependents();
  if (!early) pdl_wait_predecessor();
  const int j = threadIdx.x, nt = blockDim.x;
  const uint32_t chunk = blockIdx.x, split = blockIdx.y, nsplit = gridDim.y, beam = blockIdx.z;
  uint64_t f0, f1;
  split_range(ndf, split, nsplit, f0, f1);
  const size_t fstride = (size_t)(fpitch / 16u); /* uint4 units per data frame of the source */
  const uint4 *base =
      (const uint4 *)beams.ptr[beam] + f0 * fstride + (size_t)chunk * units_per_pkt;
  T a0 = 0, a1 = 0;
  for (uint32_t u = j; u < units_per_pkt; u += nt) {
    const uint4 *p = base + u;
    uint64_t f = f0;
    for (; f + UF <= f1; f += UF) {
      uint4 v[UF];
#pragma unroll
      for (int k = 0; k < UF; ++k) v[k] = ldg_stream(p + k * fstride);
      p += UF * fstride;
#pragma unroll
      for (int k = 0; k < UF; ++k) {
        if (CALIB) {
          a0 += (T)(v[k].x ^ v[k].y);
          a1 += (T)(v[k].z ^ v[k].w);
        } else {
          Acc::template add<BE>(a0, v[k].x, v[k].y);
          Acc::template add<BE>(a1, v[k].z, v[k].w);
        }
      }
    }
    for (; f < f1; ++f) {
      uint4 v = ldg_stream(p);
      p += fstride;
      Acc::template add<BE>(a0, v.x, v.y);
      Acc::template add<BE>(a1, v.z, v.w);
    }
  }
  const size_t nchan = (size_t)nchunk * nch;
  T *dst = partials + ((size_t)beam * nsplit + split) * nchan + (size_t)chunk * nch;
  const T a[2] = {a0, a1};
  const int c[2] = {(int)((2 * j) % nch), (int)((2 * j + 1) % nch)};
  if (early) pdl_wait_predecessor();
  cta_reduce_n<T, 2>(a, c, (int)nch, red, dst, j, nt);
  unsigned int *cnt = fold.colcnt + beam * nchunk + chunk;
  if (!column_arrive<0, 0>(cnt, nsplit, &last_flag, j)) return;
  column_fold<T, 0, 0>(partials, red, fold, beam, beams.slot[beam], nsplit, nchan, chunk * nch,
                       (int)nch, j, nt, cnt);
}

/* Bandwidth calibration only (B2P_CALIB=1): flat grid-stride read of the block,
   no unpack/detect; tells how close the real kernels sit to a plain streaming read. */
template <typename T>
__global__ void __launch_bounds__(256, 4)
b2p_calib_flat(const B2pBeams beams, const uint64_t nunits, T *__restrict__ partials)
{
  const uint4 *p = (const uint4 *)beams.ptr[blockIdx.y];
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t x = 0;
  for (; i + 7 * stride < nunits; i += 8 * stride) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = ldg_stream(p + i + u * stride);
#pragma unroll
    for (int u = 0; u < 8; ++u) x ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  for (; i < nunits; i += stride) {
    uint4 v = ldg_stream(p + i);
    x ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (x == 0x12345678u) partials[0] = (T)x; /* keep the loads alive */
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
 * Work item = (beam, time split, chunk group of G).  Persistent CTAs (one per SM)
 * draw items from a ticket counter, so SMs that stream faster (nearer L2 slices)
 * simply take more items and all of them finish together — with a static
 * round-robin ncu showed SMs only 89 % active (a long tail).  One stage = the G
 * packets of one data frame of the group: G*7168 contiguous bytes, one bulk
 * copy.  The producer tags every stage with its item (and whether it is the
 * item's last frame); the consumers just follow the tags, so they never need
 * the schedule.  Consumer thread j reads unit j of each packet: 2 accumulators
 * per packet slot g.
 */
template <int G, int NSTAGE> struct TmaSmem {
  static constexpr int kStageBytes = G * kPktBytes;
  static constexpr int kBarOff = NSTAGE * kStageBytes;
  static constexpr int kTagOff = kBarOff + 2 * NSTAGE * 8;
  static constexpr int kRedOff = kTagOff + NSTAGE * 8;
  /* [14 warps][G*7] for the item sum; the column fold needs [lanes][G*7] <= 448 elements */
  static constexpr int kRedElems = kTmaConsumerWarps * G * kNchBmf > kTmaConsumers
                                       ? kTmaConsumerWarps * G * kNchBmf
                                       : kTmaConsumers;
  static constexpr int kFoldOff = kRedOff + kRedElems * 8; /* scratch of the column fold */
  static constexpr int kFlagOff = kFoldOff + kTmaConsumers * 8;
  static constexpr int kBytes = kFlagOff + 16;
};

constexpr uint32_t kTagLast = 0x80000000u; /* last frame of the item */
constexpr uint32_t kTagEnd = 0xFFFFFFFFu;  /* no more work for this CTA */

template <typename Acc, bool BE, int G, int NSTAGE, int MINB>
__global__ void __launch_bounds__(kTmaThreads, MINB)
b2p_fused_tma_bmf(const B2pBeams beams, const uint32_t nchunk, const uint64_t fpitch,
                  const uint64_t ndf, const uint32_t nsplit, const uint32_t nitems, const int early,
                  unsigned int *__restrict__ ticket, typename Acc::type *__restrict__ partials,
                  const B2pFold fold)
{
  typedef typename Acc::type T;
  typedef TmaSmem<G, NSTAGE> S;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *full = (uint64_t *)(smem + S::kBarOff);
  uint64_t *empty = full + NSTAGE;
  volatile uint32_t *tag = (volatile uint32_t *)(smem + S::kTagOff);
  T *red = (T *)(smem + S::kRedOff);
  T *foldbuf = (T *)(smem + S::kFoldOff);
  /* [0] != 0: the column of item [1] is complete, fold it at the next item boundary */
  volatile uint32_t *fold_due = (volatile uint32_t *)(smem + S::kFlagOff);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ngroups = nchunk / G;
  pdl_release_dependents();
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kTmaConsumerWarps);
    }
    mbar_fence_init();
    fold_due[0] = 0u;
  }
  if (!early) pdl_wait_predecessor();
  __syncthreads();

  if (warp == kTmaConsumerWarps) {
    /* ---- producer: one lane draws tickets and issues every bulk copy of this CTA ---- */
    if (lane == 0) {
      uint32_t it = 0;
      for (;;) {
        const uint32_t item = atomicAdd(ticket, 1u);
        const bool end = item >= nitems;
        /* every CTA draws exactly one ticket past the end: the last such draw of the launch
           leaves the counter at zero again for the launch that reuses it (or a graph replay) */
        if (item == nitems + gridDim.x - 1u) atomicExch(ticket, 0u);
        uint64_t f0 = 0, f1 = 1;
        const unsigned char *src = nullptr;
        if (!end) {
          const uint32_t group = item % ngroups, rest = item / ngroups;
          const uint32_t split = rest % nsplit, beam = rest / nsplit;
          split_range(ndf, split, nsplit, f0, f1);
          src = (const unsigned char *)beams.ptr[beam] + f0 * fpitch + (uint64_t)group * G * kPktBytes;
          if (f1 == f0) { /* an empty split still owes its (zero) partial sums */
            const uint32_t s = it % NSTAGE, k = it / NSTAGE;
            if (k > 0) mbar_wait(&empty[s], (k - 1) & 1);
            tag[s] = item | kTagLast | 0x40000000u; /* bit 30: no data in this stage */
            mbar_arrive(&full[s]);
            ++it;
            continue;
          }
        }
        const size_t fstride = (size_t)fpitch;
        for (uint64_t f = f0; f < f1; ++f, ++it, src += fstride) {
          const uint32_t s = it % NSTAGE, k = it / NSTAGE;
          if (k > 0) mbar_wait(&empty[s], (k - 1) & 1);
          if (end) {
            tag[s] = kTagEnd;
            mbar_arrive(&full[s]);
          } else {
            tag[s] = item | (f + 1 == f1 ? kTagLast : 0u);
            mbar_arrive_expect_tx(&full[s], S::kStageBytes);
            bulk_g2s(smem + s * S::kStageBytes, src, S::kStageBytes, &full[s]);
          }
        }
        if (end) break;
      }
    }
    return;
  }

  /* ---- consumers: follow the stage tags ---- */
  const int c0 = (2 * tid) % kNchBmf, c1 = (2 * tid + 1) % kNchBmf;
  const size_t nchan = (size_t)nchunk * kNchBmf;
  bool waited = !early;
  T a[G][2];
#pragma unroll
  for (int g = 0; g < G; ++g) a[g][0] = a[g][1] = 0;
  for (uint32_t it = 0;; ++it) {
    const uint32_t s = it % NSTAGE, k = it / NSTAGE;
    mbar_wait(&full[s], k & 1);
    const uint32_t t = tag[s];
    if (t == kTagEnd) break;
    if (!(t & 0x40000000u)) {
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
    } else {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (!(t & kTagLast)) continue;

    /* item done: G*7 channel totals of this (beam, split, group) */
    const uint32_t item = t & 0x3FFFFFFFu;
    const uint32_t group = item % ngroups, rest = item / ngroups;
    const uint32_t split = rest % nsplit, beam = rest / nsplit;
    T *dst = partials + ((size_t)beam * nsplit + split) * nchan + (size_t)group * G * kNchBmf;
#pragma unroll
    for (int g = 0; g < G; ++g)
      for (int ch = 0; ch < kNchBmf; ++ch) {
        T v = (c0 == ch ? a[g][0] : (T)0) + (c1 == ch ? a[g][1] : (T)0);
        v = warp_sum(v);
        if (lane == 0) red[warp * (G * kNchBmf) + g * kNchBmf + ch] = v;
      }
    consumer_bar(); /* A: the warp sums of this item are in `red` */
    /*
     * Column bookkeeping without stalling the consumers: thread 0 alone does the arrival
     * atomic of an item (after barrier B below) while the others go on streaming; if that
     * made a column complete it leaves a note, and everybody folds that column here, at the
     * next item boundary (or after the last stage).
     */
    if (fold_due[0]) {
      const uint32_t it2 = fold_due[1];
      const uint32_t g2 = it2 % ngroups, b2 = (it2 / ngroups) / nsplit;
      column_fold_call<T, 1, kTmaConsumers>(partials, foldbuf, &fold, b2, beams.slot[b2], nsplit,
                                            (uint32_t)nchan, g2 * G * kNchBmf, G * kNchBmf, tid,
                                            fold.colcnt + b2 * ngroups + g2);
      if (tid == 0) fold_due[0] = 0u;
    }
    /* the previous kernel of the stream has finished with `partials`, the counters and acc */
    if (!waited) pdl_wait_predecessor();
    waited = true;
    if (tid < G * kNchBmf) {
      T sum = 0;
      for (int w = 0; w < kTmaConsumerWarps; ++w) sum += red[w * (G * kNchBmf) + tid];
      dst[tid] = sum;
    }
    consumer_bar(); /* B: `red` is free again; the partial sums precede thread 0's arrival */
    if (tid == 0) {
      unsigned int old;
      asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;"
                   : "=r"(old)
                   : "l"(fold.colcnt + beam * ngroups + group)
                   : "memory");
      if (old == nsplit - 1u) { /* last item of its column: all nsplit partial sums are there */
        fold_due[1] = item;
        __threadfence_block();
        fold_due[0] = 1u;
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) a[g][0] = a[g][1] = 0;
  }
  /* every consumer has seen the end tag; a column completed by this CTA's last item is still due */
  consumer_bar();
  if (fold_due[0]) {
    const uint32_t it2 = fold_due[1];
    const uint32_t g2 = it2 % ngroups, b2 = (it2 / ngroups) / nsplit;
    column_fold_call<T, 1, kTmaConsumers>(partials, foldbuf, &fold, b2, beams.slot[b2], nsplit,
                                          (uint32_t)nchan, g2 * G * kNchBmf, G * kNchBmf, tid,
                                          fold.colcnt + b2 * ngroups + g2);
  }
}

/* ------------------------------------------------------- stand-alone finish */
/*
 * Used only when an integration is closed without a fused launch to ride on (b2p_finish
 * after plain accumulate calls): out = (float)acc * scale, acc = 0.  The steady-state path
 * (b2p_integrate_*) finishes inside the fused kernel and never launches this.
 */
template <typename T>
__global__ void __launch_bounds__(256)
b2p_finish_k(const uint32_t n, T *__restrict__ acc, float *__restrict__ out, const float scale)
{
  pdl_release_dependents();
  pdl_wait_predecessor();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = __fmul_rn(to_f32_rn(acc[i]), scale);
  acc[i] = 0;
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

/* ------------------------------------------------------------ launch helper */

template <typename... KArgs, typename... Args>
cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                     bool pdl, Args... args)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
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

template <typename Acc, bool BE, int G, int NSTAGE, int MINB = 1>
cudaError_t launch_tma(const B2pLaunch &L, cudaStream_t st)
{
  typedef TmaSmem<G, NSTAGE> S;
  const uint32_t ngroups = L.nchunk / G;
  const uint32_t nitems = (uint32_t)L.nbeam * L.nsplit * ngroups;
  uint32_t grid = (uint32_t)L.sm_count * MINB;
  if (grid > nitems) grid = nitems;
  return launch_k(b2p_fused_tma_bmf<Acc, BE, G, NSTAGE, MINB>, dim3(grid), dim3(kTmaThreads), S::kBytes,
                  st, L.pdl != 0, L.beams, (uint32_t)L.nchunk, L.fpitch, L.ndf, (uint32_t)L.nsplit,
                  nitems, L.early, L.ticket, (typename Acc::type *)L.partials, L.fold);
}

template <typename Acc, bool BE> cudaError_t launch_fused_t(const B2pLaunch &L, cudaStream_t st)
{
  typedef typename Acc::type T;
  const bool bmf = b2p_is_bmf_geometry(L.nch, L.nsamp);
  const bool pdl = L.pdl != 0;
  if (L.kernel == B2P_KERNEL_TMA && bmf) {
    if (L.variant == 1 && L.nchunk % 2 == 0) return launch_tma<Acc, BE, 2, kTmaStagesG2>(L, st);
    if (L.variant == 2) return launch_tma<Acc, BE, 1, kTmaStagesG1>(L, st);
    if (L.variant == 3) return launch_tma<Acc, BE, 1, 14, 2>(L, st); /* 2 CTAs/SM, 100 KB rings */
    switch (b2p_tma_group(L.nchunk)) {
      case 4: return launch_tma<Acc, BE, 4, kTmaStagesG4>(L, st);
      case 2: return launch_tma<Acc, BE, 2, kTmaStagesG2>(L, st);
      default: return launch_tma<Acc, BE, 1, kTmaStagesG1>(L, st);
    }
  }
  const dim3 grid((unsigned)L.nchunk, (unsigned)L.nsplit, (unsigned)L.nbeam);
  T *part = (T *)L.partials;
  const size_t sh448 = 32 * 7 * sizeof(T); /* >= [14 warps][7] and = [32 lanes][7] */
  if (bmf && L.calib && L.variant == 9) /* calibration only: flat streaming read */
    return launch_k(b2p_calib_flat<T>, dim3(4 * L.sm_count, L.nbeam), dim3(256), 0, st, false,
                    L.beams, L.ndf * (uint64_t)L.nchunk * kUnitsBmf, part);
  if (bmf && L.calib && L.variant == 7) /* calibration only: same pattern, loads + XOR */
    return launch_k(b2p_fused_ldg128<Acc, BE, 8, true>, grid, dim3(448), sh448, st, pdl, L.beams,
                    (uint32_t)L.nchunk, 7u, (uint32_t)kUnitsBmf, L.fpitch, L.ndf, L.early, part,
                    L.fold);
  if (bmf && L.variant == 1) /* tuning point: 128-bit loads, 448 threads, 8 frames in flight */
    return launch_k(b2p_fused_ldg128<Acc, BE, 8, false>, grid, dim3(448), sh448, st, pdl, L.beams,
                    (uint32_t)L.nchunk, 7u, (uint32_t)kUnitsBmf, L.fpitch, L.ndf, L.early, part,
                    L.fold);
  if (bmf) {
    if (L.variant == 2)
      return launch_k(b2p_fused_ldg256_bmf<Acc, BE, 2, 6>, grid, dim3(224), 0, st, pdl, L.beams,
                      (uint32_t)L.nchunk, L.fpitch, L.ndf, L.early, part, L.fold);
    if (L.variant == 3)
      return launch_k(b2p_fused_ldg256_bmf<Acc, BE, 8, 2>, grid, dim3(224), 0, st, pdl, L.beams,
                      (uint32_t)L.nchunk, L.fpitch, L.ndf, L.early, part, L.fold);
    /* default: 4 x 256-bit loads in flight per thread, 4 CTAs/SM — fastest measured */
    return launch_k(b2p_fused_ldg256_bmf<Acc, BE, 4, 4>, grid, dim3(224), 0, st, pdl, L.beams,
                    (uint32_t)L.nchunk, L.fpitch, L.ndf, L.early, part, L.fold);
  }
  const int nt = 32 * L.nch;
  const size_t sh = (size_t)32 * L.nch * sizeof(T); /* nwarps == nch <= 32 */
  return launch_k(b2p_fused_ldg128<Acc, BE, 4, false>, grid, dim3(nt), sh, st, pdl, L.beams,
                  (uint32_t)L.nchunk, (uint32_t)L.nch, (uint32_t)(L.nsamp * L.nch / 2), L.fpitch,
                  L.ndf, L.early, part, L.fold);
}

template <typename Acc, bool BE, int G, int NSTAGE, int MINB = 1> cudaError_t configure_tma()
{
  return cudaFuncSetAttribute(b2p_fused_tma_bmf<Acc, BE, G, NSTAGE, MINB>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                              TmaSmem<G, NSTAGE>::kBytes);
}
template <typename Acc, bool BE> cudaError_t configure_all()
{
  cudaError_t e;
  if ((e = configure_tma<Acc, BE, 4, kTmaStagesG4>()) != cudaSuccess) return e;
  if ((e = configure_tma<Acc, BE, 2, kTmaStagesG2>()) != cudaSuccess) return e;
  if ((e = configure_tma<Acc, BE, 1, 14, 2>()) != cudaSuccess) return e;
  return configure_tma<Acc, BE, 1, kTmaStagesG1>();
}

template <typename T> cudaError_t launch_finish_t(const B2pFinish &R, cudaStream_t st)
{
  const uint32_t n = (uint32_t)R.nrows * (uint32_t)R.nchan;
  return launch_k(b2p_finish_k<T>, dim3((n + 255) / 256), dim3(256), 0, st, R.pdl != 0, n,
                  (T *)R.acc, R.out, R.scale);
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

cudaError_t b2p_launch_finish(const B2pFinish &R, cudaStream_t st)
{
  if (R.mode == B2P_MODE_FLOAT) return launch_finish_t<double>(R, st);
  return launch_finish_t<unsigned long long>(R, st);
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
