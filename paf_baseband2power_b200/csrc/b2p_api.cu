/*
 * b2p_api.cu — C ABI (include/b2p.h) over the sm_100a kernels.
 *
 * Host-side role: what init_baseband2power / do_baseband2power /
 * destroy_baseband2power would have done around the kernels had the reference
 * written them (baseband2power.cu:1-16 is empty; the sibling stage
 * diskdb.cu:12-134 shows the init/do/destroy convention).  Error convention:
 * the reference's CudaSafeCall prints file:line and exit(-1)s
 * (cudautil.cuh:29-41); a library must not exit, so the same file:line message
 * is stored in the context and the call returns B2P_ECUDA.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "b2p_kernels.cuh"

#define B2P_MAX_STAGE_BUFS 8
#define B2P_NTICKETS 4096
#define B2P_OUT_DEPTH 4 /* finished integrations that may wait to be collected (b2p_wait_output) */

struct b2p_ctx {
  b2p_params p;
  int nchan;            /* channels this context produces: nchunk * nch_per_chunk */
  uint64_t pkt_bytes;   /* payload of one packet (one frame of one chunk) */
  uint64_t frame_bytes; /* one data frame of this context's chunks (compact) */
  uint64_t src_pitch;   /* one data frame of the source stream: nchunk_total * pkt_bytes */
  uint64_t src_offset;  /* first_chunk * pkt_bytes */
  int sm_count;
  int kernel;  /* resolved */
  int nsplit;  /* resolved: splits of a full-length launch */
  int split_base; /* smallest split count that fills whole waves */
  int cap_nchunk; /* chunks the buffers are sized for (nchunk_total when the range may move) */
  int cap_nsplit;
  uint64_t user_stage_ndf; /* stage_ndf as given at creation (0 = derive from the chunk range) */
  int variant; /* tuning variant, B2P_VARIANT env (experiments) */
  int no_early; /* B2P_NO_EARLY=1: never start a fused kernel before its predecessor ends */
  int calib;
  size_t acc_elem;
  cudaStream_t compute, copy;
  void *acc;
  void *partials;
  float *out_dev;    /* [B2P_OUT_DEPTH][nbeam][nchan] */
  float *out_pinned; /* [B2P_OUT_DEPTH][nbeam][nchan] */
  cudaEvent_t out_ready[B2P_OUT_DEPTH];
  uint64_t out_head, out_tail; /* spectra queued / collected over the context's life */
  unsigned int *colcnt; /* per (launch beam, column) arrival counters; zero between launches */
  /* ticket counters of the persistent (TMA) kernel: a ring, one per fused launch in flight;
     a launch leaves its counter at zero (the last draw past the end resets it) */
  unsigned int *tickets;
  uint64_t fused_seq;
  /* host-path staging */
  int nbufs;
  void *stage[B2P_MAX_STAGE_BUFS];
  cudaEvent_t copied[B2P_MAX_STAGE_BUFS], consumed[B2P_MAX_STAGE_BUFS];
  uint64_t pieces; /* pieces issued so far over the context's life */
  cudaEvent_t h2d_begin, h2d_end; /* around the H2D copies of the last host call (copy stream) */
  int h2d_timed;
  /* per-launch timing */
  int timing;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used;
  uint64_t launches;
  /* stream of the most recent launch: a launch on another stream is ordered behind it */
  cudaStream_t last_stream;
  cudaEvent_t xev;
  char err[512];
};

static thread_local char g_create_err[512] = "";

static int set_err(b2p_ctx *c, int code, const char *fmt, const char *a, const char *file, int line)
{
  char *dst = c ? c->err : g_create_err;
  snprintf(dst, 512, fmt, a, file, line);
  return code;
}

#define CK(c, call)                                                                          \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      return set_err((c), B2P_ECUDA, "CUDA error: %s, which happens at \"%s\", line [%d].",  \
                     cudaGetErrorString(e__), __FILE__, __LINE__);                           \
  } while (0)

#define FAIL(c, code, msg) \
  return set_err((c), (code), "%s, which happens at \"%s\", line [%d].", (msg), __FILE__, __LINE__)

static unsigned gcd_u(unsigned a, unsigned b)
{
  while (b) {
    unsigned t = a % b;
    a = b;
    b = t;
  }
  return a;
}

extern "C" {

const char *b2p_version(void) { return B2P_VERSION; }

void b2p_default_params(b2p_params *p)
{
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->device_id = 0;
  p->nchunk = 48;
  p->nch_per_chunk = 7;
  p->nsamp_df = 128;
  p->big_endian = 1;
  p->scale = 1.0f;
  p->mode = B2P_MODE_EXACT;
  p->nbeam = 1;
  p->kernel = B2P_KERNEL_AUTO;
  p->nsplit = 0;
  p->stage_ndf = 0;
  p->nstage_bufs = 0;
  p->first_chunk = 0;
  p->nchunk_total = 0;
  p->resizable = 0;
}

const char *b2p_last_error(const b2p_ctx *ctx) { return ctx ? ctx->err : g_create_err; }

int b2p_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int b2p_device_info(int device, char *name, size_t name_len, int *sm_count, int *cc_major,
                    int *cc_minor, uint64_t *mem_bytes)
{
  cudaDeviceProp prop;
  CK(NULL, cudaGetDeviceProperties(&prop, device));
  if (name && name_len) snprintf(name, name_len, "%s", prop.name);
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (mem_bytes) *mem_bytes = (uint64_t)prop.totalGlobalMem;
  return B2P_OK;
}

static int resolve_kernel(const b2p_params *p)
{
  int k = p->kernel;
  const char *env = getenv("B2P_KERNEL");
  if (k == B2P_KERNEL_AUTO && env) {
    if (!strcmp(env, "ldg")) k = B2P_KERNEL_LDG;
    if (!strcmp(env, "tma")) k = B2P_KERNEL_TMA;
  }
  const bool bmf = b2p_is_bmf_geometry(p->nch_per_chunk, p->nsamp_df);
  if (k == B2P_KERNEL_AUTO) k = B2P_KERNEL_LDG;
  if (k == B2P_KERNEL_TMA && !bmf) k = B2P_KERNEL_LDG; /* TMA variant is BMF-only */
  return k;
}

static int resolve_nsplit(const b2p_params *p, int kernel, int sm_count, int *base_out)
{
  *base_out = 0;
  if (p->nsplit > 0) return p->nsplit > 1024 ? 1024 : p->nsplit;
  /* base = smallest split count that makes the grid a whole number of waves.  With the
     cross-CTA fold inside the kernel every CTA costs a prologue, a block reduction and an
     arrival, so fewer, longer CTAs win: 74 splits (6 waves of 4 CTAs/SM for one beam, ~110
     frames per CTA) beat round 1's 222 by 0.8 % chained and 0.9 % isolated; 37 and 111 are
     within 0.3 % of 74, 444 is 4 % behind (profiles/r02_nsplit_sweep.txt). */
  unsigned slots, units, target;
  if (kernel == B2P_KERNEL_TMA) {
    slots = (unsigned)sm_count;
    units = (unsigned)(p->nchunk / b2p_tma_group(p->nchunk)) * (unsigned)p->nbeam;
    target = 74; /* items are drawn dynamically; 74: +3.4 % chained, +2.6 % isolated over 222 */
  } else {
    slots = 4u * (unsigned)sm_count;
    units = (unsigned)p->nchunk * (unsigned)p->nbeam;
    target = 74; /* 8192 frames / 74 = 110 frames per CTA (tapered 131 / 65 / 33) */
  }
  const unsigned base = slots / gcd_u(slots, units); /* smallest n with n*units % slots == 0 */
  *base_out = (int)(base > 1024 ? 1024 : base);
  unsigned k = (target + base / 2) / base;
  if (k < 1) k = 1;
  unsigned n = base * k;
  if (n > 1024) n = 1024;
  if (n < 1) n = 1;
  return (int)n;
}

int b2p_create(b2p_ctx **out, const b2p_params *p)
{
  if (!out || !p) FAIL(NULL, B2P_EINVAL, "b2p_create: NULL argument");
  *out = NULL;
  if (p->nchunk <= 0 || p->nch_per_chunk <= 0 || p->nsamp_df <= 0)
    FAIL(NULL, B2P_EINVAL, "b2p_create: geometry must be positive");
  if (p->nch_per_chunk > 32) FAIL(NULL, B2P_EINVAL, "b2p_create: nch_per_chunk > 32 unsupported");
  if ((p->nsamp_df * p->nch_per_chunk) % 2)
    FAIL(NULL, B2P_EINVAL, "b2p_create: packet payload must be a multiple of 16 bytes");
  if (p->nbeam < 1 || p->nbeam > B2P_MAX_BEAMS)
    FAIL(NULL, B2P_EINVAL, "b2p_create: nbeam out of range");
  if (p->mode != B2P_MODE_EXACT && p->mode != B2P_MODE_FLOAT)
    FAIL(NULL, B2P_EINVAL, "b2p_create: unknown mode");
  if (p->kernel < B2P_KERNEL_AUTO || p->kernel > B2P_KERNEL_TMA)
    FAIL(NULL, B2P_EINVAL, "b2p_create: unknown kernel variant");
  const int ntotal = p->nchunk_total > 0 ? p->nchunk_total : p->nchunk;
  if (p->first_chunk < 0 || p->first_chunk + p->nchunk > ntotal)
    FAIL(NULL, B2P_EINVAL, "b2p_create: chunk range [first_chunk, first_chunk+nchunk) exceeds nchunk_total");

  int ndev = 0;
  CK(NULL, cudaGetDeviceCount(&ndev));
  if (ndev <= 0) FAIL(NULL, B2P_ECUDA, "b2p_create: no CUDA device (there is no CPU fallback)");
  int dev = p->device_id;
  /* one visible GPU (container) -> its index is 0: paf_baseband2power.cu:86-90 */
  if (ndev == 1) dev = 0;
  if (dev < 0 || dev >= ndev) FAIL(NULL, B2P_EINVAL, "b2p_create: device_id out of range");
  CK(NULL, cudaSetDevice(dev));

  b2p_ctx *c = new (std::nothrow) b2p_ctx();
  if (!c) FAIL(NULL, B2P_ENOMEM, "b2p_create: out of host memory");
  c->p = *p;
  c->p.device_id = dev;
  c->p.nchunk_total = ntotal;
  c->nchan = p->nchunk * p->nch_per_chunk;
  c->pkt_bytes = (uint64_t)p->nsamp_df * p->nch_per_chunk * 8u;
  c->frame_bytes = (uint64_t)p->nchunk * c->pkt_bytes;
  c->src_pitch = (uint64_t)ntotal * c->pkt_bytes;
  c->src_offset = (uint64_t)p->first_chunk * c->pkt_bytes;
  c->acc_elem = 8;
  c->err[0] = 0;
  c->pieces = 0;
  c->out_head = c->out_tail = 0;
  for (int i = 0; i < B2P_OUT_DEPTH; ++i) c->out_ready[i] = NULL;
  c->timing = 0;
  c->ev_used = 0;
  c->launches = 0;
  c->nbufs = 0;
  c->acc = c->partials = NULL;
  c->out_dev = c->out_pinned = NULL;
  c->tickets = c->colcnt = NULL;
  c->fused_seq = 0;
  c->compute = c->copy = NULL;
  c->last_stream = NULL;
  c->xev = NULL;
  c->h2d_begin = c->h2d_end = NULL;
  c->h2d_timed = 0;
  for (int i = 0; i < B2P_MAX_STAGE_BUFS; ++i) {
    c->stage[i] = NULL;
    c->copied[i] = c->consumed[i] = NULL;
  }

#define CKC(call)                                                                             \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      set_err(NULL, B2P_ECUDA, "CUDA error: %s, which happens at \"%s\", line [%d].",         \
              cudaGetErrorString(e__), __FILE__, __LINE__);                                   \
      b2p_destroy(c);                                                                         \
      return B2P_ECUDA;                                                                       \
    }                                                                                         \
  } while (0)

  CKC(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, dev));
  CKC(b2p_kernels_configure());
  c->kernel = resolve_kernel(p);
  c->nsplit = resolve_nsplit(p, c->kernel, c->sm_count, &c->split_base);
  c->user_stage_ndf = p->stage_ndf;
  c->cap_nchunk = p->resizable ? ntotal : p->nchunk;
  c->cap_nsplit = c->nsplit;
  if (p->resizable) /* the chunk range may be moved later: size the scratch for any range */
    for (int n = 1; n <= ntotal; ++n) {
      b2p_params q = *p;
      q.nchunk = n;
      int base = 0;
      const int ns = resolve_nsplit(&q, resolve_kernel(&q), c->sm_count, &base);
      if (ns > c->cap_nsplit) c->cap_nsplit = ns;
    }
  c->variant = getenv("B2P_VARIANT") ? atoi(getenv("B2P_VARIANT")) : 0;
  c->no_early = getenv("B2P_NO_EARLY") ? atoi(getenv("B2P_NO_EARLY")) : 0;
  c->calib = getenv("B2P_CALIB") ? atoi(getenv("B2P_CALIB")) : 0;
  CKC(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
  CKC(cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
  CKC(cudaEventCreateWithFlags(&c->xev, cudaEventDisableTiming));
  CKC(cudaEventCreate(&c->h2d_begin));
  CKC(cudaEventCreate(&c->h2d_end));
  const size_t nacc = (size_t)p->nbeam * c->cap_nchunk * p->nch_per_chunk;
  const size_t ncnt = (size_t)p->nbeam * c->cap_nchunk;
  CKC(cudaMalloc(&c->acc, nacc * c->acc_elem));
  CKC(cudaMemset(c->acc, 0, nacc * c->acc_elem));
  CKC(cudaMalloc(&c->partials, nacc * (size_t)c->cap_nsplit * c->acc_elem));
  CKC(cudaMalloc((void **)&c->out_dev, B2P_OUT_DEPTH * nacc * sizeof(float)));
  for (int i = 0; i < B2P_OUT_DEPTH; ++i) CKC(cudaEventCreateWithFlags(&c->out_ready[i], cudaEventDisableTiming));
  CKC(cudaMalloc((void **)&c->tickets, B2P_NTICKETS * sizeof(unsigned int)));
  CKC(cudaMemset(c->tickets, 0, B2P_NTICKETS * sizeof(unsigned int)));
  CKC(cudaMalloc((void **)&c->colcnt, ncnt * sizeof(unsigned int)));
  CKC(cudaMemset(c->colcnt, 0, ncnt * sizeof(unsigned int)));
  CKC(cudaHostAlloc((void **)&c->out_pinned, B2P_OUT_DEPTH * nacc * sizeof(float), cudaHostAllocDefault));
  CKC(cudaDeviceSynchronize());
#undef CKC
  *out = c;
  return B2P_OK;
}

void b2p_destroy(b2p_ctx *c)
{
  if (!c) return;
  cudaSetDevice(c->p.device_id);
  if (c->last_stream && c->last_stream != c->compute) cudaStreamSynchronize(c->last_stream);
  if (c->compute) cudaStreamSynchronize(c->compute);
  if (c->copy) cudaStreamSynchronize(c->copy);
  for (int i = 0; i < B2P_MAX_STAGE_BUFS; ++i) {
    if (c->stage[i]) cudaFree(c->stage[i]);
    if (c->copied[i]) cudaEventDestroy(c->copied[i]);
    if (c->consumed[i]) cudaEventDestroy(c->consumed[i]);
  }
  for (size_t i = 0; i < c->ev_pool.size(); ++i) cudaEventDestroy(c->ev_pool[i]);
  if (c->xev) cudaEventDestroy(c->xev);
  for (int i = 0; i < B2P_OUT_DEPTH; ++i)
    if (c->out_ready[i]) cudaEventDestroy(c->out_ready[i]);
  if (c->h2d_begin) cudaEventDestroy(c->h2d_begin);
  if (c->h2d_end) cudaEventDestroy(c->h2d_end);
  if (c->acc) cudaFree(c->acc);
  if (c->partials) cudaFree(c->partials);
  if (c->out_dev) cudaFree(c->out_dev);
  if (c->tickets) cudaFree(c->tickets);
  if (c->colcnt) cudaFree(c->colcnt);
  if (c->out_pinned) cudaFreeHost(c->out_pinned);
  if (c->compute) cudaStreamDestroy(c->compute);
  if (c->copy) cudaStreamDestroy(c->copy);
  delete c;
}

int b2p_nchan(const b2p_ctx *c) { return c ? c->nchan : 0; }
uint64_t b2p_frame_bytes(const b2p_ctx *c) { return c ? c->frame_bytes : 0; }
uint64_t b2p_source_frame_bytes(const b2p_ctx *c) { return c ? c->src_pitch : 0; }
int b2p_first_chunk(const b2p_ctx *c) { return c ? c->p.first_chunk : 0; }
int b2p_kernel_in_use(const b2p_ctx *c) { return c ? c->kernel : 0; }
int b2p_nsplit_in_use(const b2p_ctx *c) { return c ? c->nsplit : 0; }
uint64_t b2p_launch_count(const b2p_ctx *c) { return c ? c->launches : 0; }
void *b2p_stream(const b2p_ctx *c) { return c ? (void *)c->compute : NULL; }

/* Successive launches of a context share partial sums, counters and accumulators, so they
   must execute in issue order.  Same stream: stream order.  Another stream: it waits for
   everything the context has put on the stream it used last. */
static int order_behind_last(b2p_ctx *c, cudaStream_t st)
{
  if (c->last_stream && c->last_stream != st) {
    CK(c, cudaEventRecord(c->xev, c->last_stream));
    CK(c, cudaStreamWaitEvent(st, c->xev, 0));
  }
  c->last_stream = st;
  return B2P_OK;
}

/* time splits of a launch of `ndf` frames: whole waves, never fewer than ~8 frames per CTA
   when that can be helped; a full-length launch gets the context's nsplit */
static int splits_for(const b2p_ctx *c, uint64_t ndf)
{
  uint64_t ns;
  if (c->split_base > 0) {
    uint64_t k = ndf / (8u * (uint64_t)c->split_base);
    const uint64_t kmax = (uint64_t)(c->nsplit / c->split_base);
    if (k > kmax) k = kmax;
    if (k < 1) k = 1;
    ns = k * (uint64_t)c->split_base;
  } else {
    ns = (uint64_t)c->nsplit; /* caller-chosen split count */
  }
  if (ns > ndf) ns = ndf;
  if (ns > (uint64_t)c->nsplit) ns = (uint64_t)c->nsplit;
  if (ns < 1) ns = 1;
  return (int)ns;
}

/*
 * One fused launch for `n` beams on `st`: unpack + detect + integrate `ndf` frames per beam
 * read with frame pitch `fpitch`, then (inside the same kernel) fold the column sums into
 * the accumulators, or — finish — emit the spectrum to `out` and clear them.
 */
static int launch_fused(b2p_ctx *c, const void *const *ptrs, const int *slots, int n, uint64_t ndf,
                        uint64_t fpitch, int kernel, cudaStream_t st, int finish, float *out)
{
  B2pLaunch L;
  memset(&L, 0, sizeof(L));
  const uintptr_t align_mask = b2p_is_bmf_geometry(c->p.nch_per_chunk, c->p.nsamp_df) ? 31u : 15u;
  for (int b = 0; b < n; ++b) {
    if (!ptrs[b]) FAIL(c, B2P_EINVAL, "accumulate: NULL beam pointer");
    if (((uintptr_t)ptrs[b]) & align_mask)
      FAIL(c, B2P_EINVAL, "accumulate: beam pointer misaligned (32 bytes for the BMF geometry, else 16)");
    L.beams.ptr[b] = ptrs[b];
    L.beams.slot[b] = slots ? slots[b] : b;
  }
  int rc = order_behind_last(c, st);
  if (rc) return rc;
  L.nbeam = n;
  L.nchunk = c->p.nchunk;
  L.nch = c->p.nch_per_chunk;
  L.nsamp = c->p.nsamp_df;
  L.fpitch = fpitch;
  L.big_endian = c->p.big_endian;
  L.mode = c->p.mode;
  L.kernel = kernel;
  L.sm_count = c->sm_count;
  L.variant = c->variant;
  L.calib = c->calib;
  L.pdl = 1;
  /* Only on the context's own stream is it known that the input was complete before
     the call: there the kernel may start while its predecessor is still running. */
  L.early = (st == c->compute && !c->no_early) ? 1 : 0;
  L.ndf = ndf;
  L.partials = c->partials;
  L.nsplit = splits_for(c, ndf);
  L.fold.acc = c->acc;
  L.fold.out = out;
  L.fold.colcnt = c->colcnt;
  L.fold.scale = c->p.scale;
  L.fold.finish = finish;

  /* a counter of its own for every launch that can be in flight at once */
  const uint64_t seq = c->fused_seq++;
  L.ticket = c->tickets + (seq % B2P_NTICKETS);

  cudaEvent_t e0 = NULL, e1 = NULL;
  if (c->timing) {
    while (c->ev_pool.size() < c->ev_used + 2) {
      cudaEvent_t e;
      CK(c, cudaEventCreate(&e));
      c->ev_pool.push_back(e);
    }
    e0 = c->ev_pool[c->ev_used];
    e1 = c->ev_pool[c->ev_used + 1];
    c->ev_used += 2;
    CK(c, cudaEventRecord(e0, st));
  }
  CK(c, b2p_launch_fused(L, st));
  if (c->timing) CK(c, cudaEventRecord(e1, st));
  c->launches += 1;
  return B2P_OK;
}

/* the caller's pointers address frame 0 of the source stream; a shard starts first_chunk in */
static void shard_ptrs(const b2p_ctx *c, const void *const *in, const void **out)
{
  for (int b = 0; b < c->p.nbeam; ++b)
    out[b] = in[b] ? (const void *)((const unsigned char *)in[b] + c->src_offset) : NULL;
}

static int device_call(b2p_ctx *c, const void *const *dptrs, uint64_t ndf, void *stream, int finish,
                       float *out_dev)
{
  if (!c) return B2P_EINVAL;
  if (!dptrs) FAIL(c, B2P_EINVAL, "b2p_accumulate_device: NULL dptrs");
  CK(c, cudaSetDevice(c->p.device_id));
  cudaStream_t st = stream ? (cudaStream_t)stream : c->compute;
  if (ndf == 0) {
    if (!finish) return B2P_OK;
    return b2p_finish_device(c, out_dev, stream);
  }
  const void *ptrs[B2P_MAX_BEAMS];
  shard_ptrs(c, dptrs, ptrs);
  return launch_fused(c, ptrs, NULL, c->p.nbeam, ndf, c->src_pitch, c->kernel, st, finish, out_dev);
}

int b2p_accumulate_device(b2p_ctx *c, const void *const *dptrs, uint64_t ndf, void *stream)
{
  return device_call(c, dptrs, ndf, stream, 0, NULL);
}

int b2p_integrate_device(b2p_ctx *c, const void *const *dptrs, uint64_t ndf, float *out_dev,
                         void *stream)
{
  if (!c) return B2P_EINVAL;
  if (!out_dev) FAIL(c, B2P_EINVAL, "b2p_integrate_device: NULL output");
  return device_call(c, dptrs, ndf, stream, 1, out_dev);
}

static int ensure_staging(b2p_ctx *c)
{
  if (c->nbufs) return B2P_OK;
  int nb = c->p.nstage_bufs > 0 ? c->p.nstage_bufs : 3;
  if (nb < 2) nb = 2;
  if (nb > B2P_MAX_STAGE_BUFS) nb = B2P_MAX_STAGE_BUFS;
  size_t bytes;
  if (c->user_stage_ndf == 0) {
    /* ~88 MB per piece whatever the shard width: 256 frames of 48 chunks */
    uint64_t n = 256u * (uint64_t)c->p.nchunk_total / (uint64_t)c->p.nchunk;
    c->p.stage_ndf = n ? n : 1;
    bytes = (size_t)256u * c->p.nchunk_total * c->pkt_bytes;
    if (bytes < (size_t)c->p.stage_ndf * c->frame_bytes) bytes = (size_t)c->p.stage_ndf * c->frame_bytes;
  } else {
    bytes = (size_t)c->user_stage_ndf * c->cap_nchunk * c->pkt_bytes;
  }
  for (int i = 0; i < nb; ++i) {
    CK(c, cudaMalloc(&c->stage[i], bytes));
    CK(c, cudaEventCreateWithFlags(&c->copied[i], cudaEventDisableTiming));
    CK(c, cudaEventCreateWithFlags(&c->consumed[i], cudaEventDisableTiming));
  }
  c->nbufs = nb;
  return B2P_OK;
}

/* queue: per beam, pieces of stage_ndf frames -> H2D (2-D when the context is a chunk-group
   shard of wider frames) into the staging ring -> fused kernel; the last piece of a beam
   finishes that beam's row when `finish`.  Nothing here waits for the device. */
static int host_issue(b2p_ctx *c, const void *const *hptrs, uint64_t ndf, int finish)
{
  if (ndf) {
    if (!hptrs) FAIL(c, B2P_EINVAL, "b2p_accumulate_host: NULL hptrs");
    for (int b = 0; b < c->p.nbeam; ++b)
      if (!hptrs[b]) FAIL(c, B2P_EINVAL, "b2p_accumulate_host: NULL beam pointer");
  }
  CK(c, cudaSetDevice(c->p.device_id));
  const size_t nacc = (size_t)c->p.nbeam * c->nchan;
  float *out_slot = NULL;
  if (finish) {
    if (c->out_head - c->out_tail >= B2P_OUT_DEPTH)
      FAIL(c, B2P_ESTATE, "too many finished integrations wait to be collected (b2p_wait_output)");
    out_slot = c->out_dev + (c->out_head % B2P_OUT_DEPTH) * nacc;
  }
  if (ndf == 0) {
    if (!finish) return B2P_OK;
    int rc0 = b2p_finish_device(c, out_slot, c->compute);
    if (rc0) return rc0;
  }
  int rc = ensure_staging(c);
  if (rc) return rc;
  const uint64_t piece = c->p.stage_ndf;
  const bool compact = c->src_pitch == c->frame_bytes;
  if (ndf) CK(c, cudaEventRecord(c->h2d_begin, c->copy));
  for (int b = 0; b < c->p.nbeam && ndf; ++b) {
    const unsigned char *src = (const unsigned char *)hptrs[b] + c->src_offset;
    for (uint64_t f0 = 0; f0 < ndf; f0 += piece) {
      const uint64_t n = (ndf - f0 < piece) ? ndf - f0 : piece;
      const int buf = (int)(c->pieces % (uint64_t)c->nbufs);
      if (c->pieces >= (uint64_t)c->nbufs) CK(c, cudaStreamWaitEvent(c->copy, c->consumed[buf], 0));
      if (compact)
        CK(c, cudaMemcpyAsync(c->stage[buf], src + f0 * c->src_pitch, n * c->frame_bytes,
                              cudaMemcpyHostToDevice, c->copy));
      else /* chunks [first_chunk, +nchunk) of every frame: rows of frame_bytes, pitch src_pitch */
        CK(c, cudaMemcpy2DAsync(c->stage[buf], c->frame_bytes, src + f0 * c->src_pitch, c->src_pitch,
                                c->frame_bytes, n, cudaMemcpyHostToDevice, c->copy));
      CK(c, cudaEventRecord(c->copied[buf], c->copy));
      CK(c, cudaStreamWaitEvent(c->compute, c->copied[buf], 0));
      const void *ptr = c->stage[buf];
      const int fin = finish && f0 + n == ndf;
      rc = launch_fused(c, &ptr, &b, 1, n, c->frame_bytes, c->kernel, c->compute, fin, out_slot);
      if (rc) return rc;
      CK(c, cudaEventRecord(c->consumed[buf], c->compute));
      c->pieces++;
    }
  }
  if (ndf) {
    CK(c, cudaEventRecord(c->h2d_end, c->copy));
    c->h2d_timed = 1;
  }
  if (finish) {
    const int slot = (int)(c->out_head % B2P_OUT_DEPTH);
    CK(c, cudaMemcpyAsync(c->out_pinned + slot * nacc, out_slot, nacc * sizeof(float), cudaMemcpyDeviceToHost,
                          c->compute));
    CK(c, cudaEventRecord(c->out_ready[slot], c->compute));
    c->out_head++;
  }
  return B2P_OK;
}

int b2p_last_h2d_ms(b2p_ctx *c, double *ms)
{
  if (!c || !ms) return B2P_EINVAL;
  if (!c->h2d_timed) FAIL(c, B2P_ESTATE, "b2p_last_h2d_ms: no host call has been made yet");
  CK(c, cudaSetDevice(c->p.device_id));
  float f = 0.f;
  CK(c, cudaEventSynchronize(c->h2d_end));
  CK(c, cudaEventElapsedTime(&f, c->h2d_begin, c->h2d_end));
  *ms = (double)f;
  return B2P_OK;
}

int b2p_accumulate_host_async(b2p_ctx *c, const void *const *hptrs, uint64_t ndf, int finish)
{
  if (!c) return B2P_EINVAL;
  return host_issue(c, hptrs, ndf, finish);
}

int b2p_wait_input(b2p_ctx *c)
{
  if (!c) return B2P_EINVAL;
  CK(c, cudaSetDevice(c->p.device_id));
  /* every byte has left the host block once the copy stream drains */
  CK(c, cudaStreamSynchronize(c->copy));
  return B2P_OK;
}

int b2p_wait_output(b2p_ctx *c, float *out_host)
{
  if (!c) return B2P_EINVAL;
  if (!out_host) FAIL(c, B2P_EINVAL, "b2p_wait_output: NULL output");
  if (c->out_head == c->out_tail) FAIL(c, B2P_ESTATE, "b2p_wait_output: no finished integration is queued");
  CK(c, cudaSetDevice(c->p.device_id));
  const size_t nacc = (size_t)c->p.nbeam * c->nchan;
  const int slot = (int)(c->out_tail % B2P_OUT_DEPTH);
  CK(c, cudaEventSynchronize(c->out_ready[slot])); /* the oldest one; later work keeps running */
  memcpy(out_host, c->out_pinned + slot * nacc, nacc * sizeof(float));
  c->out_tail++;
  return B2P_OK;
}

int b2p_accumulate_host(b2p_ctx *c, const void *const *hptrs, uint64_t ndf)
{
  if (!c) return B2P_EINVAL;
  int rc = host_issue(c, hptrs, ndf, 0);
  if (rc) return rc;
  return b2p_wait_input(c);
}

int b2p_integrate_host(b2p_ctx *c, const void *const *hptrs, uint64_t ndf, float *out_host)
{
  if (!c) return B2P_EINVAL;
  if (!out_host) FAIL(c, B2P_EINVAL, "b2p_integrate_host: NULL output");
  int rc = host_issue(c, hptrs, ndf, 1);
  if (rc) return rc;
  return b2p_wait_output(c, out_host); /* the compute stream runs behind the copy stream */
}

int b2p_accumulate_host_mapped(b2p_ctx *c, const void *const *hptrs, uint64_t ndf)
{
  if (!c) return B2P_EINVAL;
  if (!hptrs) FAIL(c, B2P_EINVAL, "b2p_accumulate_host_mapped: NULL hptrs");
  if (ndf == 0) return B2P_OK;
  CK(c, cudaSetDevice(c->p.device_id));
  const void *dptrs[B2P_MAX_BEAMS];
  for (int b = 0; b < c->p.nbeam; ++b) {
    if (!hptrs[b]) FAIL(c, B2P_EINVAL, "b2p_accumulate_host_mapped: NULL beam pointer");
    void *d = NULL;
    CK(c, cudaHostGetDevicePointer(&d, (void *)hptrs[b], 0));
    dptrs[b] = (const unsigned char *)d + c->src_offset;
  }
  /* the kernel walks the host block itself, frame pitch = that of the source stream */
  int rc = launch_fused(c, dptrs, NULL, c->p.nbeam, ndf, c->src_pitch, B2P_KERNEL_LDG, c->compute, 0,
                        NULL);
  if (rc) return rc;
  CK(c, cudaStreamSynchronize(c->compute));
  return B2P_OK;
}

int b2p_finish_device(b2p_ctx *c, float *out_dev, void *stream)
{
  if (!c) return B2P_EINVAL;
  if (!out_dev) FAIL(c, B2P_EINVAL, "b2p_finish_device: NULL output");
  CK(c, cudaSetDevice(c->p.device_id));
  cudaStream_t st = stream ? (cudaStream_t)stream : c->compute;
  int rc = order_behind_last(c, st);
  if (rc) return rc;
  B2pFinish R;
  memset(&R, 0, sizeof(R));
  R.nrows = c->p.nbeam;
  R.nchan = c->nchan;
  R.mode = c->p.mode;
  R.pdl = 1;
  R.acc = c->acc;
  R.out = out_dev;
  R.scale = c->p.scale;
  CK(c, b2p_launch_finish(R, st));
  c->launches += 1;
  return B2P_OK;
}

int b2p_finish(b2p_ctx *c, float *out_host)
{
  if (!c) return B2P_EINVAL;
  if (!out_host) FAIL(c, B2P_EINVAL, "b2p_finish: NULL output");
  if (c->out_head != c->out_tail)
    FAIL(c, B2P_ESTATE, "b2p_finish: collect the queued integrations first (b2p_wait_output)");
  int rc = host_issue(c, NULL, 0, 1); /* finish kernel + D2H into the output queue */
  if (rc) return rc;
  return b2p_wait_output(c, out_host);
}

/* wait for everything the context has launched, wherever it launched it */
static int drain(b2p_ctx *c)
{
  if (c->last_stream && c->last_stream != c->compute) CK(c, cudaStreamSynchronize(c->last_stream));
  CK(c, cudaStreamSynchronize(c->copy));
  CK(c, cudaStreamSynchronize(c->compute));
  return B2P_OK;
}

int b2p_read_sums(b2p_ctx *c, uint64_t *sums_host)
{
  if (!c) return B2P_EINVAL;
  if (!sums_host) FAIL(c, B2P_EINVAL, "b2p_read_sums: NULL output");
  if (c->p.mode != B2P_MODE_EXACT) FAIL(c, B2P_ESTATE, "b2p_read_sums: exact mode only");
  CK(c, cudaSetDevice(c->p.device_id));
  int rc = drain(c);
  if (rc) return rc;
  CK(c, cudaMemcpy(sums_host, c->acc, (size_t)c->p.nbeam * c->nchan * 8, cudaMemcpyDeviceToHost));
  return B2P_OK;
}

int b2p_reset(b2p_ctx *c)
{
  if (!c) return B2P_EINVAL;
  CK(c, cudaSetDevice(c->p.device_id));
  int rc = drain(c);
  if (rc) return rc;
  /* nothing else to put right: every launch leaves its counters (columns, TMA ticket) at zero */
  CK(c, cudaMemsetAsync(c->acc, 0, (size_t)c->p.nbeam * c->nchan * c->acc_elem, c->compute));
  CK(c, cudaStreamSynchronize(c->compute));
  c->out_tail = c->out_head; /* drop spectra nobody collected */
  return B2P_OK;
}

int b2p_set_chunk_range(b2p_ctx *c, int first_chunk, int nchunk)
{
  if (!c) return B2P_EINVAL;
  if (!c->p.resizable) FAIL(c, B2P_ESTATE, "b2p_set_chunk_range: the context was not created resizable");
  if (nchunk < 1 || first_chunk < 0 || first_chunk + nchunk > c->p.nchunk_total)
    FAIL(c, B2P_EINVAL, "b2p_set_chunk_range: range exceeds nchunk_total");
  if (c->out_head != c->out_tail)
    FAIL(c, B2P_ESTATE, "b2p_set_chunk_range: finished integrations wait to be collected");
  CK(c, cudaSetDevice(c->p.device_id));
  int rc = drain(c);
  if (rc) return rc;
  c->p.nchunk = nchunk;
  c->p.first_chunk = first_chunk;
  c->nchan = nchunk * c->p.nch_per_chunk;
  c->frame_bytes = (uint64_t)nchunk * c->pkt_bytes;
  c->src_offset = (uint64_t)first_chunk * c->pkt_bytes;
  c->nsplit = resolve_nsplit(&c->p, c->kernel, c->sm_count, &c->split_base);
  if (c->nsplit > c->cap_nsplit) c->nsplit = c->cap_nsplit;
  if (c->user_stage_ndf == 0) {
    uint64_t n = 256u * (uint64_t)c->p.nchunk_total / (uint64_t)nchunk;
    c->p.stage_ndf = n ? n : 1;
  }
  /* between integrations the accumulators are zero anyway; their row pitch has changed */
  CK(c, cudaMemsetAsync(c->acc, 0, (size_t)c->p.nbeam * c->cap_nchunk * c->p.nch_per_chunk * c->acc_elem, c->compute));
  CK(c, cudaStreamSynchronize(c->compute));
  return B2P_OK;
}

int b2p_set_timing(b2p_ctx *c, int enabled)
{
  if (!c) return B2P_EINVAL;
  c->timing = enabled ? 1 : 0;
  return B2P_OK;
}

int b2p_fused_time_ms(b2p_ctx *c, double *sum_ms, uint64_t *launches)
{
  if (!c) return B2P_EINVAL;
  CK(c, cudaSetDevice(c->p.device_id));
  double sum = 0.0;
  for (size_t i = 0; i + 1 < c->ev_used; i += 2) {
    float ms = 0.f;
    CK(c, cudaEventSynchronize(c->ev_pool[i + 1]));
    CK(c, cudaEventElapsedTime(&ms, c->ev_pool[i], c->ev_pool[i + 1]));
    sum += ms;
  }
  if (sum_ms) *sum_ms = sum;
  if (launches) *launches = c->ev_used / 2;
  c->ev_used = 0;
  return B2P_OK;
}

/* ------------------------------------------------------------ memory helpers */

int b2p_host_alloc(void **p, size_t bytes)
{
  if (!p) return B2P_EINVAL;
  CK(NULL, cudaHostAlloc(p, bytes, cudaHostAllocMapped | cudaHostAllocPortable));
  return B2P_OK;
}
int b2p_host_free(void *p)
{
  CK(NULL, cudaFreeHost(p));
  return B2P_OK;
}
int b2p_host_register(void *p, size_t bytes)
{
  CK(NULL, cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
  return B2P_OK;
}
int b2p_host_unregister(void *p)
{
  CK(NULL, cudaHostUnregister(p));
  return B2P_OK;
}
int b2p_device_alloc(int device, void **p, size_t bytes)
{
  if (!p) return B2P_EINVAL;
  CK(NULL, cudaSetDevice(device));
  CK(NULL, cudaMalloc(p, bytes));
  return B2P_OK;
}
int b2p_device_free(int device, void *p)
{
  CK(NULL, cudaSetDevice(device));
  CK(NULL, cudaFree(p));
  return B2P_OK;
}
int b2p_memcpy_h2d(int device, void *dst, const void *src, size_t bytes)
{
  CK(NULL, cudaSetDevice(device));
  CK(NULL, cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
  return B2P_OK;
}
int b2p_memcpy_d2h(int device, void *dst, const void *src, size_t bytes)
{
  CK(NULL, cudaSetDevice(device));
  CK(NULL, cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return B2P_OK;
}
int b2p_device_sync(int device)
{
  CK(NULL, cudaSetDevice(device));
  CK(NULL, cudaDeviceSynchronize());
  return B2P_OK;
}

/* ----------------------------------------------- link probe and chunk split */

int b2p_probe_h2d(const int *devices, int n, size_t bytes, int reps, double *gbps_out)
{
  if (!devices || !gbps_out || n < 1 || n > B2P_MAX_GROUP || bytes == 0)
    FAIL(NULL, B2P_EINVAL, "b2p_probe_h2d: bad argument");
  if (reps < 1) reps = 1;
  void *host = NULL;
  void *dev[B2P_MAX_GROUP] = {0};
  cudaStream_t st[B2P_MAX_GROUP] = {0};
  cudaEvent_t e0[B2P_MAX_GROUP] = {0}, e1[B2P_MAX_GROUP] = {0};
  cudaError_t e = cudaHostAlloc(&host, bytes, cudaHostAllocPortable);
  for (int i = 0; i < n && e == cudaSuccess; ++i) {
    if ((e = cudaSetDevice(devices[i])) != cudaSuccess) break;
    if ((e = cudaMalloc(&dev[i], bytes)) != cudaSuccess) break;
    if ((e = cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking)) != cudaSuccess) break;
    if ((e = cudaEventCreate(&e0[i])) != cudaSuccess) break;
    if ((e = cudaEventCreate(&e1[i])) != cudaSuccess) break;
    e = cudaMemcpyAsync(dev[i], host, bytes, cudaMemcpyHostToDevice, st[i]); /* warm-up */
  }
  for (int i = 0; i < n && e == cudaSuccess; ++i) {
    cudaSetDevice(devices[i]);
    e = cudaStreamSynchronize(st[i]);
  }
  /* all links at once: every device's copies are queued before any is waited for.  Pass 1
     moves the same bytes over every link, so the fast links finish first and the slow ones
     then run with less company — their rate reads too high.  Pass 2 hands every link bytes in
     proportion to its pass-1 rate, so that all stay busy until the end: its rates are what the
     links deliver TOGETHER, which is what a split over them needs. */
  size_t nbytes[B2P_MAX_GROUP];
  for (int i = 0; i < n; ++i) nbytes[i] = bytes;
  for (int pass = 0; pass < (n > 1 ? 2 : 1) && e == cudaSuccess; ++pass) {
    for (int i = 0; i < n && e == cudaSuccess; ++i) {
      cudaSetDevice(devices[i]);
      e = cudaEventRecord(e0[i], st[i]);
    }
    for (int r = 0; r < reps && e == cudaSuccess; ++r)
      for (int i = 0; i < n && e == cudaSuccess; ++i) {
        cudaSetDevice(devices[i]);
        e = cudaMemcpyAsync(dev[i], host, nbytes[i], cudaMemcpyHostToDevice, st[i]);
      }
    for (int i = 0; i < n && e == cudaSuccess; ++i) {
      cudaSetDevice(devices[i]);
      e = cudaEventRecord(e1[i], st[i]);
    }
    double best = 0.0;
    for (int i = 0; i < n && e == cudaSuccess; ++i) {
      cudaSetDevice(devices[i]);
      float ms = 0.f;
      if ((e = cudaEventSynchronize(e1[i])) != cudaSuccess) break;
      if ((e = cudaEventElapsedTime(&ms, e0[i], e1[i])) != cudaSuccess) break;
      gbps_out[i] = ms > 0.f ? (double)reps * (double)nbytes[i] / (ms * 1e-3) / 1e9 : 0.0;
      if (gbps_out[i] > best) best = gbps_out[i];
    }
    for (int i = 0; i < n && best > 0.0; ++i) {
      size_t b = (size_t)((double)bytes * gbps_out[i] / best);
      b &= ~(size_t)255;
      nbytes[i] = b < 4096 ? (bytes < 4096 ? bytes : 4096) : b;
    }
  }
  for (int i = 0; i < n; ++i) {
    if (devices[i] >= 0) cudaSetDevice(devices[i]);
    if (e0[i]) cudaEventDestroy(e0[i]);
    if (e1[i]) cudaEventDestroy(e1[i]);
    if (st[i]) cudaStreamDestroy(st[i]);
    if (dev[i]) cudaFree(dev[i]);
  }
  if (host) cudaFreeHost(host);
  CK(NULL, e);
  return B2P_OK;
}

int b2p_split_chunks(const double *weights, int n, int nchunk, int *counts)
{
  if (!counts || n < 1 || nchunk < 0) FAIL(NULL, B2P_EINVAL, "b2p_split_chunks: bad argument");
  double sum = 0.0;
  for (int i = 0; i < n; ++i) {
    const double w = weights ? weights[i] : 1.0;
    if (!(w >= 0.0)) FAIL(NULL, B2P_EINVAL, "b2p_split_chunks: weights must be >= 0");
    sum += w;
  }
  if (sum <= 0.0) FAIL(NULL, B2P_EINVAL, "b2p_split_chunks: all weights are zero");
  /* largest remainder: floor of the exact share, then the leftover chunks to the largest
     fractional parts (ties to the lower index) */
  double frac[B2P_MAX_GROUP];
  if (n > B2P_MAX_GROUP) FAIL(NULL, B2P_EINVAL, "b2p_split_chunks: too many parts");
  int given = 0;
  for (int i = 0; i < n; ++i) {
    const double share = (weights ? weights[i] : 1.0) / sum * nchunk;
    counts[i] = (int)share;
    frac[i] = share - counts[i];
    given += counts[i];
  }
  while (given < nchunk) {
    int best = 0;
    for (int i = 1; i < n; ++i)
      if (frac[i] > frac[best]) best = i;
    counts[best] += 1;
    frac[best] = -1.0;
    given += 1;
  }
  return B2P_OK;
}

/* ------------------------------------------------------- chunk-group shards */

struct b2p_group {
  int n;                       /* shards with at least one chunk */
  b2p_ctx *ctx[B2P_MAX_GROUP];
  int device[B2P_MAX_GROUP], first[B2P_MAX_GROUP], count[B2P_MAX_GROUP];
  int nbeam, nch, nchan_total;
  /* as given to b2p_group_create, for b2p_group_rebalance */
  b2p_params base;
  int ndev, all_devices[B2P_MAX_GROUP], all_counts[B2P_MAX_GROUP];
  double share[B2P_MAX_GROUP];
  int open_integration; /* frames accumulated and not yet finished */
  std::vector<float> tmp;
  char err[512];
};

static int group_fail(b2p_group *g, int i, int rc)
{
  snprintf(g->err, sizeof(g->err), "shard %d (gpu %d, chunks %d..%d): %s", i, g->device[i], g->first[i],
           g->first[i] + g->count[i] - 1, b2p_last_error(g->ctx[i]));
  return rc;
}

/* (re)build the shard contexts of g from g->all_counts */
static int group_build(b2p_group *g)
{
  for (int i = 0; i < g->n; ++i) b2p_destroy(g->ctx[i]);
  g->n = 0;
  int first = 0;
  for (int i = 0; i < g->ndev; ++i) {
    if (g->all_counts[i] == 0) continue;
    b2p_params p = g->base;
    p.device_id = g->all_devices[i];
    p.nchunk = g->all_counts[i];
    p.first_chunk = first;
    p.nchunk_total = g->base.nchunk;
    p.nsplit = 0;
    p.stage_ndf = 0;
    p.resizable = 1; /* b2p_group_rebalance moves the range in place */
    const int k = g->n;
    g->device[k] = g->all_devices[i];
    g->first[k] = first;
    g->count[k] = g->all_counts[i];
    int rc = b2p_create(&g->ctx[k], &p);
    if (rc) return rc; /* g_create_err holds the message */
    g->n = k + 1;
    first += g->all_counts[i];
  }
  return B2P_OK;
}

int b2p_group_create(b2p_group **out, const b2p_params *base, const int *devices, const int *nchunks,
                     int ndev)
{
  if (!out || !base || !devices || !nchunks || ndev < 1 || ndev > B2P_MAX_GROUP)
    FAIL(NULL, B2P_EINVAL, "b2p_group_create: bad argument");
  *out = NULL;
  int total = 0;
  for (int i = 0; i < ndev; ++i) {
    if (nchunks[i] < 0) FAIL(NULL, B2P_EINVAL, "b2p_group_create: negative chunk count");
    total += nchunks[i];
  }
  if (total != base->nchunk)
    FAIL(NULL, B2P_EINVAL, "b2p_group_create: chunk counts must add up to params.nchunk");
  b2p_group *g = new (std::nothrow) b2p_group();
  if (!g) FAIL(NULL, B2P_ENOMEM, "b2p_group_create: out of host memory");
  g->n = 0;
  g->err[0] = 0;
  g->nbeam = base->nbeam;
  g->nch = base->nch_per_chunk;
  g->nchan_total = base->nchunk * base->nch_per_chunk;
  g->base = *base;
  g->ndev = ndev;
  g->open_integration = 0;
  for (int i = 0; i < ndev; ++i) {
    g->all_devices[i] = devices[i];
    g->all_counts[i] = nchunks[i];
    g->share[i] = (double)nchunks[i] / (double)total;
  }
  int rc = group_build(g);
  if (rc) {
    for (int j = 0; j < g->n; ++j) b2p_destroy(g->ctx[j]);
    delete g;
    return rc;
  }
  g->tmp.resize((size_t)g->nbeam * g->nchan_total);
  *out = g;
  return B2P_OK;
}

/*
 * Move chunks between the GPUs so that their host links finish together: the H2D time of each
 * shard's last host call (CUDA events on its copy stream) gives the rate its link really
 * delivered with all links busy — links behind one host bridge slow each other down, which no
 * solo probe sees.  Only between integrations.  Returns 1 in *changed when the split moved.
 */
int b2p_group_rebalance(b2p_group *g, int *changed)
{
  if (!g) return B2P_EINVAL;
  if (changed) *changed = 0;
  if (g->open_integration) {
    snprintf(g->err, sizeof(g->err), "b2p_group_rebalance: an integration is open");
    return B2P_ESTATE;
  }
  for (int i = 0; i < g->n; ++i)
    if (g->ctx[i]->out_head != g->ctx[i]->out_tail) {
      snprintf(g->err, sizeof(g->err), "b2p_group_rebalance: finished integrations wait to be collected");
      return B2P_ESTATE;
    }
  double rate[B2P_MAX_GROUP], sum_rate = 0.0, sum_share = 0.0, ms_min = 1e300, ms_max = 0.0;
  int k = 0;
  for (int i = 0; i < g->ndev; ++i) {
    rate[i] = 0.0;
    if (g->all_counts[i] == 0) continue;
    double ms = 0.0;
    int rc = b2p_last_h2d_ms(g->ctx[k], &ms);
    if (rc) return group_fail(g, k, rc);
    rate[i] = ms > 0.0 ? (double)g->all_counts[i] / ms : 0.0;
    sum_rate += rate[i];
    sum_share += g->share[i];
    if (ms < ms_min) ms_min = ms;
    if (ms > ms_max) ms_max = ms;
    ++k;
  }
  if (sum_rate <= 0.0) return B2P_OK;
  if (ms_max <= 1.04 * ms_min) return B2P_OK; /* the links already finish together */
  /* damped: half way from the present shares to the measured rates (shares of unused GPUs stay 0) */
  double target[B2P_MAX_GROUP];
  for (int i = 0; i < g->ndev; ++i) {
    target[i] = g->all_counts[i] ? 0.5 * g->share[i] / sum_share + 0.5 * rate[i] / sum_rate : 0.0;
    g->share[i] = target[i];
  }
  int counts[B2P_MAX_GROUP];
  if (b2p_split_chunks(target, g->ndev, g->base.nchunk, counts) != B2P_OK) return B2P_EINVAL;
  /* a GPU that takes part keeps at least one chunk (its context stays) */
  for (int i = 0; i < g->ndev; ++i)
    if (g->all_counts[i] > 0 && counts[i] == 0) {
      int big = 0;
      for (int j = 1; j < g->ndev; ++j)
        if (counts[j] > counts[big]) big = j;
      counts[big] -= 1;
      counts[i] = 1;
    }
  bool same = true;
  for (int i = 0; i < g->ndev; ++i) same = same && counts[i] == g->all_counts[i];
  if (same) return B2P_OK;
  /* move the ranges in place: no allocation, no context rebuilt */
  int first = 0;
  k = 0;
  for (int i = 0; i < g->ndev; ++i) {
    if (g->all_counts[i] == 0) continue;
    int rc = b2p_set_chunk_range(g->ctx[k], first, counts[i]);
    if (rc) return group_fail(g, k, rc);
    g->first[k] = first;
    g->count[k] = counts[i];
    g->all_counts[i] = counts[i];
    first += counts[i];
    ++k;
  }
  if (changed) *changed = 1;
  return B2P_OK;
}

void b2p_group_destroy(b2p_group *g)
{
  if (!g) return;
  for (int i = 0; i < g->n; ++i) b2p_destroy(g->ctx[i]);
  delete g;
}

const char *b2p_group_last_error(const b2p_group *g) { return g ? g->err : g_create_err; }
int b2p_group_size(const b2p_group *g) { return g ? g->n : 0; }
b2p_ctx *b2p_group_ctx(const b2p_group *g, int i) { return (g && i >= 0 && i < g->n) ? g->ctx[i] : NULL; }
int b2p_group_shard(const b2p_group *g, int i, int *device, int *first_chunk, int *nchunk)
{
  if (!g || i < 0 || i >= g->n) return B2P_EINVAL;
  if (device) *device = g->device[i];
  if (first_chunk) *first_chunk = g->first[i];
  if (nchunk) *nchunk = g->count[i];
  return B2P_OK;
}

/* shard i's [nbeam][count*nch] spectrum -> channels [first*nch, ...) of out[nbeam][nchan_total] */
static void group_scatter(const b2p_group *g, int i, const float *part, float *out)
{
  const int w = g->count[i] * g->nch;
  for (int b = 0; b < g->nbeam; ++b)
    memcpy(out + (size_t)b * g->nchan_total + (size_t)g->first[i] * g->nch, part + (size_t)b * w,
           (size_t)w * sizeof(float));
}

/* queue every GPU's copies and kernels (one host thread keeps all links busy); nothing waits */
int b2p_group_issue_host(b2p_group *g, const void *const *hptrs, uint64_t ndf, int finish)
{
  if (!g) return B2P_EINVAL;
  for (int i = 0; i < g->n; ++i) {
    int rc = b2p_accumulate_host_async(g->ctx[i], hptrs, ndf, finish);
    if (rc) return group_fail(g, i, rc);
  }
  g->open_integration = finish ? 0 : (ndf ? 1 : g->open_integration);
  return B2P_OK;
}

int b2p_group_wait_input(b2p_group *g)
{
  if (!g) return B2P_EINVAL;
  for (int i = 0; i < g->n; ++i) {
    int rc = b2p_wait_input(g->ctx[i]);
    if (rc) return group_fail(g, i, rc);
  }
  return B2P_OK;
}

int b2p_group_wait_output(b2p_group *g, float *out_host)
{
  if (!g || !out_host) return B2P_EINVAL;
  for (int i = 0; i < g->n; ++i) {
    int rc = b2p_wait_output(g->ctx[i], g->tmp.data());
    if (rc) return group_fail(g, i, rc);
    group_scatter(g, i, g->tmp.data(), out_host);
  }
  return B2P_OK;
}

int b2p_group_accumulate_host(b2p_group *g, const void *const *hptrs, uint64_t ndf)
{
  int rc = b2p_group_issue_host(g, hptrs, ndf, 0);
  return rc ? rc : b2p_group_wait_input(g);
}

int b2p_group_integrate_host(b2p_group *g, const void *const *hptrs, uint64_t ndf, float *out_host)
{
  if (!g || !out_host) return B2P_EINVAL;
  int rc = b2p_group_issue_host(g, hptrs, ndf, 1);
  return rc ? rc : b2p_group_wait_output(g, out_host);
}

int b2p_group_finish(b2p_group *g, float *out_host)
{
  if (!g || !out_host) return B2P_EINVAL;
  int rc = b2p_group_issue_host(g, NULL, 0, 1); /* ndf 0: finish only */
  return rc ? rc : b2p_group_wait_output(g, out_host);
}

int b2p_group_reset(b2p_group *g)
{
  if (!g) return B2P_EINVAL;
  for (int i = 0; i < g->n; ++i) {
    int rc = b2p_reset(g->ctx[i]);
    if (rc) return group_fail(g, i, rc);
  }
  g->open_integration = 0;
  return B2P_OK;
}

int b2p_synth_fill_device(int device, void *dptr, uint64_t ndf, int nchunk, int nch_per_chunk,
                          int nsamp_df, int big_endian, uint64_t seed, uint64_t first_word,
                          int mode, void *stream)
{
  if (!dptr || nchunk <= 0 || nch_per_chunk <= 0 || nsamp_df <= 0)
    FAIL(NULL, B2P_EINVAL, "b2p_synth_fill_device: bad argument");
  CK(NULL, cudaSetDevice(device));
  CK(NULL, b2p_launch_synth(dptr, ndf, nchunk, nch_per_chunk, nsamp_df, big_endian, seed,
                            first_word, mode, (cudaStream_t)stream));
  if (!stream) CK(NULL, cudaStreamSynchronize(0));
  return B2P_OK;
}

int b2p_selftest_unpack(int device, int big_endian, int32_t *out_host)
{
  if (!out_host) return B2P_EINVAL;
  CK(NULL, cudaSetDevice(device));
  int32_t *d = NULL;
  CK(NULL, cudaMalloc((void **)&d, 65536 * sizeof(int32_t)));
  cudaError_t e = b2p_launch_selftest_unpack(big_endian, d, 0);
  if (e == cudaSuccess) e = cudaMemcpy(out_host, d, 65536 * sizeof(int32_t), cudaMemcpyDeviceToHost);
  cudaFree(d);
  CK(NULL, e);
  return B2P_OK;
}

} /* extern "C" */
