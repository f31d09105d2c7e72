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

struct b2p_ctx {
  b2p_params p;
  int nchan;
  uint64_t frame_bytes;
  int sm_count;
  int kernel; /* resolved */
  int nsplit; /* resolved */
  int variant; /* tuning variant, B2P_VARIANT env (experiments) */
  int no_early; /* B2P_NO_EARLY=1: never start a fused kernel before its predecessor ends */
  size_t acc_elem;
  cudaStream_t compute, copy;
  void *acc;
  void *partials;
  float *out_dev;
  float *out_pinned;
  /* ticket counters of the persistent (TMA) kernel: a ring of counters, one per fused
     launch; the reduce kernel that follows a launch puts its counter back to zero */
  unsigned int *tickets;
  uint64_t fused_seq;
  /* host-path staging */
  int nbufs;
  void *stage[B2P_MAX_STAGE_BUFS];
  cudaEvent_t copied[B2P_MAX_STAGE_BUFS], consumed[B2P_MAX_STAGE_BUFS];
  uint64_t pieces; /* pieces issued so far over the context's life */
  /* per-launch timing */
  int timing;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used;
  uint64_t launches;
  /* partial sums of the last fused launch not yet folded into acc */
  int pending;
  B2pSlots pend_slots;
  int pend_nsplit;
  cudaStream_t pend_stream;
  unsigned int *pend_ticket;
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

static int resolve_nsplit(const b2p_params *p, int kernel, int sm_count)
{
  if (p->nsplit > 0) return p->nsplit > 1024 ? 1024 : p->nsplit;
  /* base = smallest split count that makes the grid a whole number of waves;
     LDG: 4 CTAs/SM and ~37 frames per CTA measured best (18 waves for one beam). */
  unsigned slots, units, target;
  if (kernel == B2P_KERNEL_TMA) {
    slots = (unsigned)sm_count;
    units = (unsigned)(p->nchunk / b2p_tma_group(p->nchunk)) * (unsigned)p->nbeam;
    target = 222; /* items are drawn dynamically: many small ones balance the SMs */
  } else {
    slots = 4u * (unsigned)sm_count;
    units = (unsigned)p->nchunk * (unsigned)p->nbeam;
    target = 222; /* 8192 frames / 222 = 37 frames per CTA */
  }
  const unsigned base = slots / gcd_u(slots, units); /* smallest n with n*units % slots == 0 */
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
  c->nchan = p->nchunk * p->nch_per_chunk;
  c->frame_bytes = (uint64_t)p->nchunk * p->nsamp_df * p->nch_per_chunk * 8u;
  c->acc_elem = 8;
  c->err[0] = 0;
  c->pieces = 0;
  c->timing = 0;
  c->ev_used = 0;
  c->launches = 0;
  c->pending = 0;
  c->pend_nsplit = 0;
  c->pend_stream = NULL;
  c->pend_ticket = NULL;
  c->nbufs = 0;
  c->acc = c->partials = NULL;
  c->out_dev = c->out_pinned = NULL;
  c->tickets = NULL;
  c->fused_seq = 0;
  c->compute = c->copy = NULL;
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
  c->nsplit = resolve_nsplit(p, c->kernel, c->sm_count);
  c->variant = getenv("B2P_VARIANT") ? atoi(getenv("B2P_VARIANT")) : 0;
  c->no_early = getenv("B2P_NO_EARLY") ? atoi(getenv("B2P_NO_EARLY")) : 0;
  CKC(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
  CKC(cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
  const size_t nacc = (size_t)p->nbeam * c->nchan;
  CKC(cudaMalloc(&c->acc, nacc * c->acc_elem));
  CKC(cudaMemset(c->acc, 0, nacc * c->acc_elem));
  CKC(cudaMalloc(&c->partials, nacc * (size_t)c->nsplit * c->acc_elem));
  CKC(cudaMalloc((void **)&c->out_dev, nacc * sizeof(float)));
  CKC(cudaMalloc((void **)&c->tickets, B2P_NTICKETS * sizeof(unsigned int)));
  CKC(cudaMemset(c->tickets, 0, B2P_NTICKETS * sizeof(unsigned int)));
  CKC(cudaHostAlloc((void **)&c->out_pinned, nacc * sizeof(float), cudaHostAllocDefault));
  CKC(cudaDeviceSynchronize());
#undef CKC
  *out = c;
  return B2P_OK;
}

void b2p_destroy(b2p_ctx *c)
{
  if (!c) return;
  cudaSetDevice(c->p.device_id);
  if (c->compute) cudaStreamSynchronize(c->compute);
  if (c->copy) cudaStreamSynchronize(c->copy);
  for (int i = 0; i < B2P_MAX_STAGE_BUFS; ++i) {
    if (c->stage[i]) cudaFree(c->stage[i]);
    if (c->copied[i]) cudaEventDestroy(c->copied[i]);
    if (c->consumed[i]) cudaEventDestroy(c->consumed[i]);
  }
  for (size_t i = 0; i < c->ev_pool.size(); ++i) cudaEventDestroy(c->ev_pool[i]);
  if (c->acc) cudaFree(c->acc);
  if (c->partials) cudaFree(c->partials);
  if (c->out_dev) cudaFree(c->out_dev);
  if (c->tickets) cudaFree(c->tickets);
  if (c->out_pinned) cudaFreeHost(c->out_pinned);
  if (c->compute) cudaStreamDestroy(c->compute);
  if (c->copy) cudaStreamDestroy(c->copy);
  delete c;
}

int b2p_nchan(const b2p_ctx *c) { return c ? c->nchan : 0; }
uint64_t b2p_frame_bytes(const b2p_ctx *c) { return c ? c->frame_bytes : 0; }
int b2p_kernel_in_use(const b2p_ctx *c) { return c ? c->kernel : 0; }
int b2p_nsplit_in_use(const b2p_ctx *c) { return c ? c->nsplit : 0; }
uint64_t b2p_launch_count(const b2p_ctx *c) { return c ? c->launches : 0; }
void *b2p_stream(const b2p_ctx *c) { return c ? (void *)c->compute : NULL; }

/* Fold the pending partial sums into the accumulators (finish=0) or emit the
   spectrum and clear (finish=1).  One kernel either way. */
static int launch_reduce(b2p_ctx *c, int finish, float *out, cudaStream_t st)
{
  B2pReduce R;
  memset(&R, 0, sizeof(R));
  if (c->pending)
    R.slots = c->pend_slots;
  else
    for (int i = 0; i < B2P_MAX_BEAMS; ++i) R.slots.lb[i] = -1;
  R.nrows = c->p.nbeam;
  R.nsplit = c->pending ? c->pend_nsplit : 0;
  R.nchan = c->nchan;
  R.mode = c->p.mode;
  R.finish = finish;
  R.pdl = 1;
  R.partials = c->partials;
  R.acc = c->acc;
  R.out = out;
  R.scale = c->p.scale;
  R.ticket = c->pending ? c->pend_ticket : NULL;
  CK(c, b2p_launch_reduce(R, st));
  c->pending = 0;
  c->launches += 1;
  return B2P_OK;
}

static int flush_pending(b2p_ctx *c, cudaStream_t st)
{
  if (!c->pending) return B2P_OK;
  if (c->pend_stream != st) CK(c, cudaStreamSynchronize(c->pend_stream));
  return launch_reduce(c, 0, NULL, st);
}

/* One fused launch for `n` beams on `st`; its partial sums stay pending until the
   next accumulate (folded into acc first) or finish (folded and converted at once). */
static int launch_fused(b2p_ctx *c, const void *const *ptrs, const int *slots, int n, uint64_t ndf,
                        int kernel, cudaStream_t st)
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
  int rc = flush_pending(c, st);
  if (rc) return rc;
  L.nbeam = n;
  L.nchunk = c->p.nchunk;
  L.nch = c->p.nch_per_chunk;
  L.nsamp = c->p.nsamp_df;
  L.big_endian = c->p.big_endian;
  L.mode = c->p.mode;
  L.kernel = kernel;
  L.sm_count = c->sm_count;
  L.variant = c->variant;
  L.calib = getenv("B2P_CALIB") ? atoi(getenv("B2P_CALIB")) : 0;
  L.pdl = 1;
  /* Only on the context's own stream is it known that the input was complete before
     the call: there the kernel may start while its predecessor is still running. */
  L.early = (st == c->compute && !c->no_early) ? 1 : 0;
  L.ndf = ndf;
  L.partials = c->partials;
  L.acc = c->acc;
  /* short launches: never give a CTA fewer than 16 frames if it can be helped */
  uint64_t ns = ndf / 16;
  if (ns < 1) ns = 1;
  if (ns > (uint64_t)c->nsplit) ns = (uint64_t)c->nsplit;
  L.nsplit = (int)ns;

  /* a counter of its own for every launch in flight (the ring is far longer than any chain) */
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
  c->pending = 1;
  c->pend_nsplit = L.nsplit;
  c->pend_stream = st;
  c->pend_ticket = L.ticket;
  for (int i = 0; i < B2P_MAX_BEAMS; ++i) c->pend_slots.lb[i] = -1;
  for (int b = 0; b < n; ++b) c->pend_slots.lb[L.beams.slot[b]] = b;
  return B2P_OK;
}

int b2p_accumulate_device(b2p_ctx *c, const void *const *dptrs, uint64_t ndf, void *stream)
{
  if (!c) return B2P_EINVAL;
  if (!dptrs) FAIL(c, B2P_EINVAL, "b2p_accumulate_device: NULL dptrs");
  if (ndf == 0) return B2P_OK;
  CK(c, cudaSetDevice(c->p.device_id));
  cudaStream_t st = stream ? (cudaStream_t)stream : c->compute;
  return launch_fused(c, dptrs, NULL, c->p.nbeam, ndf, c->kernel, st);
}

static int ensure_staging(b2p_ctx *c)
{
  if (c->nbufs) return B2P_OK;
  int nb = c->p.nstage_bufs > 0 ? c->p.nstage_bufs : 3;
  if (nb < 2) nb = 2;
  if (nb > B2P_MAX_STAGE_BUFS) nb = B2P_MAX_STAGE_BUFS;
  if (c->p.stage_ndf == 0) c->p.stage_ndf = 256;
  const size_t bytes = (size_t)c->p.stage_ndf * c->frame_bytes;
  for (int i = 0; i < nb; ++i) {
    CK(c, cudaMalloc(&c->stage[i], bytes));
    CK(c, cudaEventCreateWithFlags(&c->copied[i], cudaEventDisableTiming));
    CK(c, cudaEventCreateWithFlags(&c->consumed[i], cudaEventDisableTiming));
  }
  c->nbufs = nb;
  return B2P_OK;
}

int b2p_accumulate_host(b2p_ctx *c, const void *const *hptrs, uint64_t ndf)
{
  if (!c) return B2P_EINVAL;
  if (!hptrs) FAIL(c, B2P_EINVAL, "b2p_accumulate_host: NULL hptrs");
  for (int b = 0; b < c->p.nbeam; ++b)
    if (!hptrs[b]) FAIL(c, B2P_EINVAL, "b2p_accumulate_host: NULL beam pointer");
  if (ndf == 0) return B2P_OK;
  CK(c, cudaSetDevice(c->p.device_id));
  int rc = ensure_staging(c);
  if (rc) return rc;
  const uint64_t piece = c->p.stage_ndf;
  for (int b = 0; b < c->p.nbeam; ++b) {
    const unsigned char *src = (const unsigned char *)hptrs[b];
    for (uint64_t f0 = 0; f0 < ndf; f0 += piece) {
      const uint64_t n = (ndf - f0 < piece) ? ndf - f0 : piece;
      const int buf = (int)(c->pieces % (uint64_t)c->nbufs);
      if (c->pieces >= (uint64_t)c->nbufs) CK(c, cudaStreamWaitEvent(c->copy, c->consumed[buf], 0));
      CK(c, cudaMemcpyAsync(c->stage[buf], src + f0 * c->frame_bytes, n * c->frame_bytes,
                            cudaMemcpyHostToDevice, c->copy));
      CK(c, cudaEventRecord(c->copied[buf], c->copy));
      CK(c, cudaStreamWaitEvent(c->compute, c->copied[buf], 0));
      const void *ptr = c->stage[buf];
      rc = launch_fused(c, &ptr, &b, 1, n, c->kernel, c->compute);
      if (rc) return rc;
      CK(c, cudaEventRecord(c->consumed[buf], c->compute));
      c->pieces++;
    }
  }
  /* every byte has left the host block once the copy stream drains */
  CK(c, cudaStreamSynchronize(c->copy));
  return B2P_OK;
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
    dptrs[b] = d;
  }
  int rc = launch_fused(c, dptrs, NULL, c->p.nbeam, ndf, B2P_KERNEL_LDG, c->compute);
  if (rc) return rc;
  CK(c, cudaStreamSynchronize(c->compute)); /* the kernel reads the host block itself */
  return B2P_OK;
}

int b2p_finish_device(b2p_ctx *c, float *out_dev, void *stream)
{
  if (!c) return B2P_EINVAL;
  if (!out_dev) FAIL(c, B2P_EINVAL, "b2p_finish_device: NULL output");
  CK(c, cudaSetDevice(c->p.device_id));
  cudaStream_t st = stream ? (cudaStream_t)stream : c->compute;
  if (c->pending && c->pend_stream != st) CK(c, cudaStreamSynchronize(c->pend_stream));
  return launch_reduce(c, 1, out_dev, st);
}

int b2p_finish(b2p_ctx *c, float *out_host)
{
  if (!c) return B2P_EINVAL;
  if (!out_host) FAIL(c, B2P_EINVAL, "b2p_finish: NULL output");
  int rc = b2p_finish_device(c, c->out_dev, c->compute);
  if (rc) return rc;
  const size_t bytes = (size_t)c->p.nbeam * c->nchan * sizeof(float);
  CK(c, cudaMemcpyAsync(c->out_pinned, c->out_dev, bytes, cudaMemcpyDeviceToHost, c->compute));
  CK(c, cudaStreamSynchronize(c->compute));
  memcpy(out_host, c->out_pinned, bytes);
  return B2P_OK;
}

int b2p_read_sums(b2p_ctx *c, uint64_t *sums_host)
{
  if (!c) return B2P_EINVAL;
  if (!sums_host) FAIL(c, B2P_EINVAL, "b2p_read_sums: NULL output");
  if (c->p.mode != B2P_MODE_EXACT) FAIL(c, B2P_ESTATE, "b2p_read_sums: exact mode only");
  CK(c, cudaSetDevice(c->p.device_id));
  if (c->pending) {
    cudaStream_t ps = c->pend_stream;
    int rc = flush_pending(c, ps);
    if (rc) return rc;
    CK(c, cudaStreamSynchronize(ps));
  }
  CK(c, cudaStreamSynchronize(c->compute));
  CK(c, cudaMemcpy(sums_host, c->acc, (size_t)c->p.nbeam * c->nchan * 8, cudaMemcpyDeviceToHost));
  return B2P_OK;
}

int b2p_reset(b2p_ctx *c)
{
  if (!c) return B2P_EINVAL;
  CK(c, cudaSetDevice(c->p.device_id));
  if (c->pending) {
    CK(c, cudaStreamSynchronize(c->pend_stream));
    c->pending = 0;
  }
  CK(c, cudaMemsetAsync(c->acc, 0, (size_t)c->p.nbeam * c->nchan * c->acc_elem, c->compute));
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
