/*
 * paf_capture — receive one beam's BMF UDP stream into the input ring.
 *
 * Program surface of the reference (paf_capture.c:27-44,59-112): -a key, -b
 * start-of-data, -c data frames per ring block, -d keep the 64-byte frame
 * header, -e NIC index (IP 10.17.<node>.<nic>, :115-118), -f DADA header
 * template, -g epoch file, -i centre frequency, -j length in seconds, -k
 * directory (log <dir>/paf_capture.log).  Ring layout as the reference writes
 * it: payload of frame idf, chunk ifreq at (idf*NCHK_NIC + ifreq)*pkt_size
 * (capture.c:540-542); chunk from the source address (capture.c:570-584); frame
 * index relative to the first frame seen (capture.c:562-568).
 *
 * New design (the reference's is known to drop frames, capture.c:20-29, and
 * swaps its block pointer under running memcpys, sync.c:109 vs capture.c:542):
 * one thread per port pulling batches with recvmmsg — with UDP_GRO where the
 * kernel offers it, so one message carries up to 8 frames of a source — and a
 * window of the current ring block plus the start of the next one, so packets
 * that straddle a block boundary are never dropped; the window moves on under
 * a write lock; packets that never arrived are zero-filled and counted, so a
 * block never carries stale data from its previous use.
 *
 * Two ring back-ends, chosen at compile time:
 *   default              two ring blocks open at once (ipcbuf_get_write_ahead, an
 *                        extension of this repo's ring shim): straddlers land in the
 *                        ring directly, no side buffer, no copy.
 *   -DB2P_STOCK_PSRDADA  only ipcio_open_block_write / ipcio_close_block_write — the
 *                        two calls the reference's capture makes (capture.c:316,
 *                        sync.c:101-109) — so it links against a stock -lpsrdada: one
 *                        block open, the head of the next block kept in a small spill
 *                        buffer (-w frames, default 256) and copied in when the block
 *                        opens (the role of the reference's tbuf, sync.c:150-165).
 *
 * Extra flags: -I bind address, -p first port, -n ports, -t socket timeout [s],
 * -w spill window in frames (stock back-end), -G 0 switches UDP_GRO off, -x the longest
 * time [ms] a port that has run past the window waits for the other ports before the oldest
 * block is retired (the port threads are not in step; without the wait a fast port would turn
 * the slower ports' frames into late ones).
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif

#include <arpa/inet.h>
#include <errno.h>
#include <math.h>
#include <netinet/in.h>
#include <netinet/udp.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <time.h>
#include <unistd.h>

#include "bmf_packet.h"
#include "dada/ascii_header.h"
#include "dada/dada_hdu.h"
#include "dada/futils.h"
#include "dada/multilog.h"

#define MSTR_LEN 512
#define HN_LEN 8
#define BATCH 16
#define SECDAY 86400.0
#ifndef UDP_GRO
#define UDP_GRO 104
#endif
#define GRO_MAX 65536 /* largest coalesced datagram the kernel hands over */

multilog_t *runtime_log;

typedef struct open_block_t {
  char *buf;
  unsigned char *seen; /* [ndf_block * nchunk] */
  int64_t index;       /* block number since the reference frame, -1 = none */
} open_block_t;

typedef struct capture_t {
  /* command line */
  key_t key;
  int sod, keep_hdr, nic, nports, port_base, nchunk, timeout_s, want_gro, gro_on;
  uint64_t ndf_block;
  uint64_t ahead_ndf; /* frames of the next block the window covers: the whole block with the
                         write-ahead back-end, the spill window with the stock one */
  double freq, length;
  char hfname[MSTR_LEN], efname[MSTR_LEN], dir[MSTR_LEN], ip[64];
  /* derived */
  int pkt_size, pkt_offset;
  uint64_t rbufsz;
  int64_t nframes_total;
  dada_hdu_t *hdu;
  int socks[16];
  /* reference frame = first frame seen */
  int have_ref;
  bmf_hdr_t ref;
  /* window of two open blocks, guarded by win (readers: port threads, writer: rotation) */
  pthread_rwlock_t win;
  open_block_t blk[2];
  int64_t base; /* index of blk[0] */
  int header_done;
  /* statistics */
  atomic_ullong n_recv, n_late, n_early, n_invalid, n_missing, n_blocks, n_msgs, ns_blocked, ns_held;
  atomic_ullong port_recv[16];
  atomic_int port_done[16];    /* the port has seen a frame past the requested length */
  atomic_llong port_block[16]; /* newest block a frame of this port belonged to */
  atomic_int port_silent[16];  /* the port let a retirement wait run out: not waited for again until it speaks */
  int64_t lag_wait_ns;         /* longest a port that is ahead holds back for the others */
  atomic_int quit, ndone;
} capture_t;

#define CAP_ERR(...)                                                                \
  do {                                                                              \
    if (runtime_log) multilog(runtime_log, LOG_ERR, __VA_ARGS__);                   \
    fprintf(stderr, __VA_ARGS__);                                                   \
    fprintf(stderr, "  which happens at \"%s\", line [%d].\n", __FILE__, __LINE__); \
  } while (0)

static void usage(void)
{
  fprintf(stdout,
          "paf_capture - capture PAF BMF raw data from NiC\n"
          "\n"
          "Usage: paf_capture [options]\n"
          " -a Hexadecimal shared memory key for capture \n"
          " -b Enable start-of-data or not \n"
          " -c Data frames (of all chunks) per ring buffer block \n"
          " -d Record the 64-byte header of every data frame or not \n"
          " -e NiC index, the IP is 10.17.<node>.<nic> \n"
          " -f DADA header template \n"
          " -g Epoch file: lines of <epoch> <days since 1970-01-01 of the epoch> \n"
          " -i Centre frequency in MHz \n"
          " -j Length of the capture in seconds \n"
          " -k Directory for the log file \n"
          " -h Show help \n"
          "extensions: -I bind address  -p first port [17100]  -n ports [6]  -t socket timeout s [27]\n"
          "            -w spill window in frames (stock PSRDADA back-end) [256]  -G 0 no UDP_GRO\n"
          "            -x longest wait [ms] of a port that is a block ahead for the others [200]\n");
}

/* UTC_START / PICOSECONDS of the reference frame: capture.c:791-843 (acquire_start_time) */
static int start_time(const capture_t *c, char utc[64], uint64_t *picoseconds)
{
  double days_epoch = 0.0;
  int found = 0;
  FILE *fp = fopen(c->efname, "r");
  if (fp) {
    char line[MSTR_LEN];
    while (fgets(line, sizeof(line), fp)) {
      int e;
      double d;
      if (line[0] != '#' && sscanf(line, "%d %lf", &e, &d) == 2 && e == c->ref.epoch) {
        days_epoch = d;
        found = 1;
        break;
      }
    }
    fclose(fp);
  }
  if (!found) {
    /* no epoch file: the header comment says half-years since 2000-01-01 (hdr.h:11) */
    days_epoch = 10957.0 + c->ref.epoch * 182.625;
    if (runtime_log) multilog(runtime_log, LOG_WARNING, "epoch %d not found in '%s', using 2000-01-01 + epoch*182.625 d\n", c->ref.epoch, c->efname);
  }
  const double sec_prd = (double)c->ref.idf * BMF_TDF_SEC;
  time_t sec = (time_t)(SECDAY * days_epoch + (double)c->ref.sec + floor(sec_prd));
  struct tm tmv;
  gmtime_r(&sec, &tmv);
  strftime(utc, 64, DADA_TIMESTR, &tmv);
  *picoseconds = (uint64_t)(1E6 * round(1.0E6 * (sec_prd - floor(sec_prd))));
  return 0;
}

static int register_header(capture_t *c)
{
  char *hdr = ipcbuf_get_next_write(c->hdu->header_block);
  if (!hdr) {
    CAP_ERR("Error getting header_buf\n");
    return -1;
  }
  if (fileread(c->hfname, hdr, DADA_DEFAULT_HEADER_SIZE) < 0) {
    CAP_ERR("Error reading header file %s\n", c->hfname);
    return -1;
  }
  char utc[64];
  uint64_t ps = 0;
  start_time(c, utc, &ps);
  if (ascii_header_set(hdr, "UTC_START", "%s", utc) < 0 ||
      ascii_header_set(hdr, "PICOSECONDS", "%lu", (unsigned long)ps) < 0 ||
      ascii_header_set(hdr, "FREQ", "%.1lf", c->freq) < 0) {
    CAP_ERR("Error setting UTC_START / PICOSECONDS / FREQ\n");
    return -1;
  }
  if (ipcbuf_mark_filled(c->hdu->header_block, DADA_DEFAULT_HEADER_SIZE) < 0) {
    CAP_ERR("Error header_fill\n");
    return -1;
  }
  multilog(runtime_log, LOG_INFO, "UTC_START:\t%s\tPICOSECONDS:\t%lu\tSEC_START:\t%lu\tIDF_START:\t%lu\n", utc,
           (unsigned long)ps, (unsigned long)c->ref.sec, (unsigned long)c->ref.idf);
  return 0;
}

static void stop_all(capture_t *c);

static uint64_t now_ns(void)
{
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (uint64_t)t.tv_sec * 1000000000ull + (uint64_t)t.tv_nsec;
}

#ifdef B2P_STOCK_PSRDADA
/* ---- stock PSRDADA: one open ring block + a spill buffer for the head of the next ---- */
static int open_first(capture_t *c)
{
  uint64_t id = 0;
  c->blk[0].buf = ipcio_open_block_write(c->hdu->data_block, &id); /* capture.c:316 */
  c->blk[1].buf = (char *)malloc(c->ahead_ndf * (uint64_t)c->nchunk * (uint64_t)c->pkt_size);
  if (!c->blk[0].buf || !c->blk[1].buf) return -1;
  c->blk[0].index = 0;
  c->blk[1].index = 1;
  memset(c->blk[0].seen, 0, c->ndf_block * (uint64_t)c->nchunk);
  memset(c->blk[1].seen, 0, c->ndf_block * (uint64_t)c->nchunk);
  return 0;
}

/* close the current block and (open_next) open the next one; what the spill buffer caught moves in */
static int advance(capture_t *c, int open_next)
{
  const uint64_t t0 = now_ns();
  if (ipcio_close_block_write(c->hdu->data_block, c->rbufsz) < 0) return -1; /* sync.c:101 */
  if (!open_next) {
    c->blk[0].buf = NULL;
    return 0;
  }
  uint64_t id = 0;
  char *next = ipcio_open_block_write(c->hdu->data_block, &id);               /* sync.c:109 */
  atomic_fetch_add(&c->ns_blocked, now_ns() - t0);
  if (!next) return -1;
  const uint64_t nspill = c->ahead_ndf * (uint64_t)c->nchunk;
  for (uint64_t i = 0; i < nspill; ++i)
    if (c->blk[1].seen[i])
      memcpy(next + i * (uint64_t)c->pkt_size, c->blk[1].buf + i * (uint64_t)c->pkt_size, (size_t)c->pkt_size);
  unsigned char *seen0 = c->blk[0].seen;
  c->blk[0].seen = c->blk[1].seen; /* the spill's marks are the new block's marks */
  c->blk[0].buf = next;
  c->blk[1].seen = seen0;
  memset(c->blk[1].seen, 0, c->ndf_block * (uint64_t)c->nchunk);
  return 0;
}
#else
/* ---- ring shim: the next block is open in the ring as well (ipcbuf_get_write_ahead) ---- */
static int open_slot(capture_t *c, int slot, int64_t index)
{
  ipcbuf_t *db = (ipcbuf_t *)c->hdu->data_block;
  const uint64_t t0 = now_ns();
  c->blk[slot].buf = ipcbuf_get_write_ahead(db, (unsigned)slot); /* waits while the ring is full */
  atomic_fetch_add(&c->ns_blocked, now_ns() - t0);
  c->blk[slot].index = index;
  memset(c->blk[slot].seen, 0, c->ndf_block * (uint64_t)c->nchunk);
  return c->blk[slot].buf ? 0 : -1; /* NULL: the ring was destroyed under us */
}

static int open_first(capture_t *c)
{
  if (open_slot(c, 0, 0) < 0) return -1;
  return open_slot(c, 1, 1);
}

static int advance(capture_t *c, int open_next)
{
  if (ipcbuf_mark_filled((ipcbuf_t *)c->hdu->data_block, c->rbufsz) < 0) return -1;
  unsigned char *seen0 = c->blk[0].seen;
  c->blk[0] = c->blk[1]; /* after mark_filled, "ahead 1" has become "ahead 0" */
  c->blk[1].seen = seen0;
  if (!open_next) { /* end of the capture: nothing will be written beyond blk[0] */
    c->blk[1].buf = NULL;
    memset(c->blk[1].seen, 0, c->ndf_block * (uint64_t)c->nchunk);
    return 0;
  }
  return open_slot(c, 1, c->base + 2);
}
#endif

static int any_seen(const capture_t *c, int slot)
{
  const uint64_t n = (slot ? c->ahead_ndf : c->ndf_block) * (uint64_t)c->nchunk;
  for (uint64_t i = 0; i < n; ++i)
    if (c->blk[slot].seen[i]) return 1;
  return 0;
}

/* Retire blk[0] (zero-fill what never arrived) and move the window one block on.
   Caller holds the write lock.  A ring that can not deliver the next block (destroyed, or an
   error from the ring library) ends the capture instead of writing through a NULL pointer. */
static void rotate(capture_t *c, int open_next)
{
  open_block_t *b = &c->blk[0];
  const uint64_t npkt = c->ndf_block * (uint64_t)c->nchunk;
  unsigned long long missing = 0;
  for (uint64_t i = 0; i < npkt; ++i)
    if (!b->seen[i]) {
      memset(b->buf + i * (uint64_t)c->pkt_size, 0, (size_t)c->pkt_size);
      ++missing;
    }
  atomic_fetch_add(&c->n_missing, missing);
  atomic_fetch_add(&c->n_blocks, 1);
  if (advance(c, open_next) < 0) {
    CAP_ERR("The ring gave no next block (destroyed?), stopping\n");
    c->blk[0].buf = c->blk[1].buf = NULL;
    stop_all(c);
  }
  c->base += 1;
}

/* Stop every port thread now: a blocked recvmmsg returns once its socket is shut down. */
static void stop_all(capture_t *c)
{
  if (atomic_exchange(&c->quit, 1)) return;
  for (int i = 0; i < c->nports; ++i) shutdown(c->socks[i], SHUT_RDWR);
}

typedef struct port_arg_t {
  capture_t *c;
  int iport;
} port_arg_t;

/* a port other than `self` that still delivers into the oldest open block, or -1 */
static int laggard(const capture_t *c, int self)
{
  for (int k = 0; k < c->nports; ++k)
    if (k != self && !c->port_done[k] && !atomic_load(&c->port_silent[k]) && atomic_load(&c->port_block[k]) <= c->base) return k;
  return -1;
}

/* one BMF frame (header + payload) from source address `from` into the window; the caller
   holds the read lock and gets it back */
static void place_frame(capture_t *c, int iport, const unsigned char *frame, const struct sockaddr_in *from)
{
  bmf_hdr_t h;
  bmf_hdr_decode(frame, &h);
  const unsigned char *ip = (const unsigned char *)&from->sin_addr.s_addr;
  const int ifreq = bmf_chunk_of_source(ip[2], ip[3]);
  if (ifreq < 0 || ifreq >= c->nchunk) {
    atomic_fetch_add(&c->n_invalid, 1);
    return;
  }
  if (!c->have_ref) { /* first frame of the stream: it becomes frame 0 */
    pthread_rwlock_unlock(&c->win);
    pthread_rwlock_wrlock(&c->win);
    if (!c->have_ref) {
      c->ref = h;
      c->have_ref = 1;
      c->base = 0;
      if (register_header(c) < 0 || open_first(c) < 0) {
        c->blk[0].buf = c->blk[1].buf = NULL;
        stop_all(c);
      }
    }
    pthread_rwlock_unlock(&c->win);
    pthread_rwlock_rdlock(&c->win);
  }
  const int64_t f = bmf_frames_since(h.sec, h.idf, c->ref.sec, c->ref.idf);
  if (f < 0) {
    atomic_fetch_add(&c->n_early, 1);
    return;
  }
  if (f >= c->nframes_total) { /* the requested length is in on this port */
    c->port_done[iport] = 1;
    return;
  }
  const int64_t bi = f / (int64_t)c->ndf_block;
  const uint64_t fin = (uint64_t)(f % (int64_t)c->ndf_block);
  if (bi > atomic_load(&c->port_block[iport])) atomic_store(&c->port_block[iport], bi);
  if (atomic_load(&c->port_silent[iport])) atomic_store(&c->port_silent[iport], 0); /* read-mostly: no line ping-pong between the ports */
  /* beyond the window (two blocks ahead, or past the spill window of the next block):
     retire the oldest block — once the other ports are through with it.  The port threads do
     not run in step; one that is ahead holds back here (its packets queue in the socket
     buffer) for at most lag_wait_ns instead of turning the others' frames into late ones. */
  uint64_t waited = 0;
  while ((bi > c->base + 1 || (bi == c->base + 1 && fin >= c->ahead_ndf)) && !atomic_load(&c->quit)) {
    const int lag = laggard(c, iport);
    if (lag >= 0 && waited < (uint64_t)c->lag_wait_ns) {
      const uint64_t t0 = now_ns();
      pthread_rwlock_unlock(&c->win);
      usleep(50);
      pthread_rwlock_rdlock(&c->win);
      const uint64_t dt = now_ns() - t0;
      waited += dt;
      atomic_fetch_add(&c->ns_held, dt);
      continue;
    }
    if (lag >= 0) /* the wait ran out: do not wait for these again until they deliver */
      for (int k = 0; k < c->nports; ++k)
        if (k != iport && !c->port_done[k] && atomic_load(&c->port_block[k]) <= c->base) atomic_store(&c->port_silent[k], 1);
    pthread_rwlock_unlock(&c->win);
    pthread_rwlock_wrlock(&c->win);
    if (bi > c->base + 1 || (bi == c->base + 1 && fin >= c->ahead_ndf)) rotate(c, 1);
    pthread_rwlock_unlock(&c->win);
    pthread_rwlock_rdlock(&c->win);
  }
  if (bi < c->base) {
    atomic_fetch_add(&c->n_late, 1);
    return;
  }
  if (bi > c->base + 1 || (bi == c->base + 1 && fin >= c->ahead_ndf)) return; /* quitting */
  open_block_t *b = &c->blk[bi - c->base];
  if (!b->buf) return; /* the ring went away */
  const uint64_t slot = fin * (uint64_t)c->nchunk + (uint64_t)ifreq;
  memcpy(b->buf + slot * (uint64_t)c->pkt_size, frame + c->pkt_offset, (size_t)c->pkt_size);
  b->seen[slot] = 1;
  atomic_fetch_add(&c->n_recv, 1);
  atomic_fetch_add(&c->port_recv[iport], 1);
}

static void *port_thread(void *argp)
{
  port_arg_t *pa = (port_arg_t *)argp;
  capture_t *c = pa->c;
  const int sock = c->socks[pa->iport];
  /* with UDP_GRO one message is up to GRO_MAX bytes: several frames of one source, back to back */
  const size_t cap = c->gro_on ? GRO_MAX : BMF_DF_SIZE;
  unsigned char *frames = (unsigned char *)malloc((size_t)BATCH * cap);
  struct mmsghdr msgs[BATCH];
  struct iovec iov[BATCH];
  struct sockaddr_in from[BATCH];
  char ctl[BATCH][CMSG_SPACE(sizeof(int)) + 32];
  if (!frames) {
    stop_all(c);
    return NULL;
  }

  while (!atomic_load(&c->quit)) {
    for (int i = 0; i < BATCH; ++i) {
      iov[i].iov_base = frames + (size_t)i * cap;
      iov[i].iov_len = cap;
      memset(&msgs[i], 0, sizeof(msgs[i]));
      msgs[i].msg_hdr.msg_iov = &iov[i];
      msgs[i].msg_hdr.msg_iovlen = 1;
      msgs[i].msg_hdr.msg_name = &from[i];
      msgs[i].msg_hdr.msg_namelen = sizeof(from[i]);
      if (c->gro_on) {
        msgs[i].msg_hdr.msg_control = ctl[i];
        msgs[i].msg_hdr.msg_controllen = sizeof(ctl[i]);
      }
    }
    const int n = recvmmsg(sock, msgs, BATCH, MSG_WAITFORONE, NULL);
    if (n <= 0) {
      if (n < 0 && (errno == EINTR)) continue;
      /* timeout: this port is silent — end of the stream (capture.c:438-456) */
      if (!atomic_load(&c->quit))
        multilog(runtime_log, LOG_WARNING, "port %d: no data for %d s, stopping\n", c->port_base + pa->iport, c->timeout_s);
      stop_all(c);
      break;
    }
    atomic_fetch_add(&c->n_msgs, (unsigned long long)n);
    pthread_rwlock_rdlock(&c->win);
    for (int i = 0; i < n; ++i) {
      const unsigned char *m = frames + (size_t)i * cap;
      size_t len = msgs[i].msg_len, seg = len;
      if (c->gro_on) /* the segment size of a coalesced message comes as a control message */
        for (struct cmsghdr *cm = CMSG_FIRSTHDR(&msgs[i].msg_hdr); cm; cm = CMSG_NXTHDR(&msgs[i].msg_hdr, cm))
          if (cm->cmsg_level == SOL_UDP && cm->cmsg_type == UDP_GRO) {
            int v = 0;
            memcpy(&v, CMSG_DATA(cm), sizeof(v));
            if (v > 0) seg = (size_t)v;
          }
      if (seg != BMF_DF_SIZE || len % BMF_DF_SIZE) {
        if (!atomic_load(&c->quit)) atomic_fetch_add(&c->n_invalid, 1);
        continue;
      }
      for (size_t off = 0; off < len; off += BMF_DF_SIZE) place_frame(c, pa->iport, m + off, &from[i]);
    }
    pthread_rwlock_unlock(&c->win);
    if (c->port_done[pa->iport]) {
      /* This port is through.  The others may still have frames of the last block queued in
         their socket buffers: the first port to finish gives them up to 2 s to get there
         before everything is stopped (a port whose last packets were lost never would). */
      if (atomic_fetch_add(&c->ndone, 1) == 0) {
        for (int k = 0; k < 200 && atomic_load(&c->ndone) < c->nports && !atomic_load(&c->quit); ++k) usleep(10000);
        stop_all(c);
      }
      break;
    }
  }
  free(frames);
  return NULL;
}

static int init_sockets(capture_t *c)
{
  for (int i = 0; i < c->nports; ++i) {
    int s = socket(AF_INET, SOCK_DGRAM, 0);
    if (s < 0) return -1;
    int one = 1, rcv = 256 << 20;
    setsockopt(s, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    /* a deep socket buffer rides out scheduling hiccups; root may go past net.core.rmem_max */
    if (setsockopt(s, SOL_SOCKET, SO_RCVBUFFORCE, &rcv, sizeof(rcv)) < 0)
      setsockopt(s, SOL_SOCKET, SO_RCVBUF, &rcv, sizeof(rcv));
    struct timeval tv = {c->timeout_s, 0}; /* SO_RCVTIMEO, capture.c:149,158 */
    setsockopt(s, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
    if (c->want_gro && i == 0) c->gro_on = 1;
    if (c->gro_on && setsockopt(s, IPPROTO_UDP, UDP_GRO, &one, sizeof(one)) < 0) {
      if (i == 0)
        c->gro_on = 0; /* kernel without UDP_GRO: plain datagrams */
      else {
        CAP_ERR("UDP_GRO accepted on the first port only (%s)\n", strerror(errno));
        return -1;
      }
    }
    struct sockaddr_in sa;
    memset(&sa, 0, sizeof(sa));
    sa.sin_family = AF_INET;
    sa.sin_port = htons((uint16_t)(c->port_base + i));
    if (inet_pton(AF_INET, c->ip, &sa.sin_addr) != 1 || bind(s, (struct sockaddr *)&sa, sizeof(sa)) < 0) {
      CAP_ERR("Can not bind to %s:%d (%s)\n", c->ip, c->port_base + i, strerror(errno));
      return -1;
    }
    c->socks[i] = s;
  }
  return 0;
}

int main(int argc, char **argv)
{
  capture_t *c = (capture_t *)calloc(1, sizeof(capture_t));
  c->key = 0xdada;
  c->sod = 1;
  c->ndf_block = 8192;
  c->nic = 1;
  c->length = 36.0; /* paf_capture.c:53 */
  c->nports = BMF_NPORT_NIC;
  c->port_base = BMF_PORT_BASE;
  c->nchunk = BMF_NCHK_NIC;
  c->timeout_s = BMF_PRD_SEC;
  c->want_gro = 1;
  c->ahead_ndf = 256;
  c->lag_wait_ns = 200 * 1000000ll;
  strcpy(c->dir, ".");
  int arg;
  while ((arg = getopt(argc, argv, "a:b:c:d:e:f:g:hi:j:k:I:p:n:t:w:G:x:")) != -1) {
    switch (arg) {
      case 'h': usage(); return EXIT_FAILURE;
      case 'a':
        if (sscanf(optarg, "%x", (unsigned *)&c->key) != 1) {
          fprintf(stderr, "Could not parse key from %s, which happens at \"%s\", line [%d].\n", optarg, __FILE__, __LINE__);
          return EXIT_FAILURE;
        }
        break;
      case 'b': c->sod = atoi(optarg); break;
      case 'c': c->ndf_block = strtoull(optarg, NULL, 10); break;
      case 'd': c->keep_hdr = atoi(optarg); break;
      case 'e': c->nic = atoi(optarg); break;
      case 'f': snprintf(c->hfname, MSTR_LEN, "%s", optarg); break;
      case 'g': snprintf(c->efname, MSTR_LEN, "%s", optarg); break;
      case 'i': c->freq = atof(optarg); break;
      case 'j': c->length = atof(optarg); break;
      case 'k': snprintf(c->dir, MSTR_LEN, "%s", optarg); break;
      case 'I': snprintf(c->ip, sizeof(c->ip), "%s", optarg); break;
      case 'p': c->port_base = atoi(optarg); break;
      case 'n': c->nports = atoi(optarg); break;
      case 't': c->timeout_s = atoi(optarg); break;
      case 'w': c->ahead_ndf = strtoull(optarg, NULL, 10); break;
      case 'G': c->want_gro = atoi(optarg) != 0; break;
      case 'x': c->lag_wait_ns = (int64_t)(atof(optarg) * 1e6); break;
      default: usage(); return EXIT_FAILURE;
    }
  }
  if (c->nports < 1 || c->nports > 16 || c->ndf_block == 0) {
    fprintf(stderr, "paf_capture: bad -n or -c\n");
    return EXIT_FAILURE;
  }
  if (!c->ip[0]) { /* 10.17.<last digit of the host name>.<nic>, paf_capture.c:115-118 */
    char hostname[HN_LEN + 1] = "";
    gethostname(hostname, HN_LEN + 1);
    hostname[HN_LEN] = 0;
    size_t l = strlen(hostname);
    int node = (l && hostname[l - 1] >= '0' && hostname[l - 1] <= '9') ? hostname[l - 1] - '0' : 0;
    snprintf(c->ip, sizeof(c->ip), "10.17.%d.%d", node, c->nic);
  }

  char log_fname[MSTR_LEN + 32];
  snprintf(log_fname, sizeof(log_fname), "%s/paf_capture.log", c->dir);
  FILE *fp_log = fopen(log_fname, "ab+");
  if (!fp_log) {
    fprintf(stderr, "Can not open log file %s\n", log_fname);
    return EXIT_FAILURE;
  }
  runtime_log = multilog_open("paf_capture", 1);
  multilog_add(runtime_log, fp_log);
  multilog(runtime_log, LOG_INFO, "START PAF_CAPTURE\n");

  c->pkt_size = c->keep_hdr ? BMF_DF_SIZE : BMF_DT_SIZE; /* capture.c:216,222 */
  c->pkt_offset = c->keep_hdr ? 0 : BMF_HDR_SIZE;
  c->rbufsz = c->ndf_block * (uint64_t)c->nchunk * (uint64_t)c->pkt_size;
  c->nframes_total = (int64_t)ceil(c->length / BMF_TDF_SEC - 1e-9);
#ifdef B2P_STOCK_PSRDADA
  if (c->ahead_ndf < 1) c->ahead_ndf = 1;
  if (c->ahead_ndf > c->ndf_block) c->ahead_ndf = c->ndf_block;
#else
  c->ahead_ndf = c->ndf_block; /* the whole next block is open in the ring */
#endif
  pthread_rwlock_init(&c->win, NULL);
  for (int s = 0; s < 2; ++s) {
    c->blk[s].seen = (unsigned char *)malloc(c->ndf_block * (uint64_t)c->nchunk);
    c->blk[s].index = -1;
  }

  /* ring: connect, check sizes, become the writer (capture.c:586-642, init_rbuf) */
  c->hdu = dada_hdu_create(runtime_log);
  dada_hdu_set_key(c->hdu, c->key);
  if (dada_hdu_connect(c->hdu) < 0) {
    CAP_ERR("Can not connect to hdu %x\n", (unsigned)c->key);
    return EXIT_FAILURE;
  }
  ipcbuf_t *db = (ipcbuf_t *)c->hdu->data_block;
  if (c->rbufsz != ipcbuf_get_bufsz(db) || ipcbuf_get_bufsz(c->hdu->header_block) != DADA_DEFAULT_HEADER_SIZE) {
    CAP_ERR("Buffer size mismatch: ring block %lu, expected %lu\n", (unsigned long)ipcbuf_get_bufsz(db), (unsigned long)c->rbufsz);
    return EXIT_FAILURE;
  }
#ifndef B2P_STOCK_PSRDADA
  if (ipcbuf_get_nbufs(db) < 3) {
    CAP_ERR("The ring needs at least 3 blocks (two are open at once)\n");
    return EXIT_FAILURE;
  }
#endif
  if (dada_hdu_lock_write(c->hdu) < 0) {
    CAP_ERR("Error locking HDU\n");
    return EXIT_FAILURE;
  }
  if ((c->sod ? ipcbuf_enable_sod(db, 0, 0) : ipcbuf_disable_sod(db)) < 0) {
    CAP_ERR("Can not set start-of-data\n");
    return EXIT_FAILURE;
  }
  if (init_sockets(c) < 0) return EXIT_FAILURE;

  pthread_t th[16];
  port_arg_t pa[16];
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int i = 0; i < c->nports; ++i) {
    pa[i].c = c;
    pa[i].iport = i;
    pthread_create(&th[i], NULL, port_thread, &pa[i]);
  }
  for (int i = 0; i < c->nports; ++i) pthread_join(th[i], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);

  /* flush: a block that received anything is delivered whole (zero-filled), an untouched one is not */
  if (c->have_ref && c->blk[0].buf) {
    if (any_seen(c, 0) || any_seen(c, 1)) {
      const int more = any_seen(c, 1);
      rotate(c, more);                                 /* the current block */
      if (more && c->blk[0].buf) rotate(c, 0);         /* and the one that had been started */
    } else {
#ifdef B2P_STOCK_PSRDADA
      ipcio_close_block_write(c->hdu->data_block, 0);  /* opened, never written: hand it back empty */
#endif
    }
  }

  /* statistics (capture.c:700-725) */
  const unsigned long long expected = (unsigned long long)atomic_load(&c->n_blocks) * c->ndf_block * (unsigned long long)c->nchunk;
  multilog(runtime_log, LOG_INFO,
           "blocks %llu  frames received %llu  expected %llu  missing(zero-filled) %llu  late %llu  early %llu  invalid %llu  in %.3f s\n",
           (unsigned long long)atomic_load(&c->n_blocks), (unsigned long long)atomic_load(&c->n_recv), expected,
           (unsigned long long)atomic_load(&c->n_missing), (unsigned long long)atomic_load(&c->n_late),
           (unsigned long long)atomic_load(&c->n_early), (unsigned long long)atomic_load(&c->n_invalid),
           (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec));
  multilog(runtime_log, LOG_INFO, "udp_gro %s  messages %llu (%.2f frames per message)  blocked on a full ring %.3f s  held for slower ports %.3f s\n",
           c->gro_on ? "on" : "off", (unsigned long long)atomic_load(&c->n_msgs),
           (double)atomic_load(&c->n_recv) / (double)(atomic_load(&c->n_msgs) ? atomic_load(&c->n_msgs) : 1),
           1e-9 * (double)atomic_load(&c->ns_blocked), 1e-9 * (double)atomic_load(&c->ns_held));
  for (int i = 0; i < c->nports; ++i)
    multilog(runtime_log, LOG_INFO, "port %d: %llu frames\n", c->port_base + i, (unsigned long long)atomic_load(&c->port_recv[i]));

  for (int i = 0; i < c->nports; ++i) close(c->socks[i]);
  dada_hdu_unlock_write(c->hdu); /* end of data */
  dada_hdu_disconnect(c->hdu);
  dada_hdu_destroy(c->hdu);
  multilog(runtime_log, LOG_INFO, "FINISH PAF_CAPTURE\n\n");
  multilog_close(runtime_log);
  fclose(fp_log);
  return EXIT_SUCCESS;
}
