/*
 * bmf_replay — synthetic BMF beamformer: emits one beam's UDP stream.
 *
 * For every data frame it sends NCHK_NIC packets of 7232 bytes (64-byte header
 * hdr.c:10-28 + 7168-byte payload, capture.h:27-29).  Chunk c is sent from
 * source address <a>.<b>.X.Y with X = c/6 + 1, Y = 2*(c%6) + 1 — the addressing
 * the capture stage decodes (capture.c:570-584) — to port base + c/8 (6 ports x
 * 8 chunks, capture.h:19-24).  The payload is the counter-based stream of
 * include/b2p_synth.h, so the ring block a correct capture assembles is byte
 * for byte the block b2p_gen / the oracle generate for the same seed.
 * On loopback the source addresses are 127.0.X.Y (all of 127/8 is local).
 *
 *  -D dest ip [127.0.0.1]  -p first port [17100]  -P ports [6]  -n frames  -s seed  -m mode
 *  -r frames per second (0 = unpaced; line rate is 9259.26)  -S sec  -i idf  -e epoch  -b beam
 *  -f first chunk frequency MHz  -L drop every L-th packet (loss injection)  -A source prefix [127.0]
 *  -C cache this many generated frames and repeat them (0 = all)  -T sender threads [1]
 *  -G consecutive frames of a chunk per send call [1]; > 1 uses UDP generic segmentation
 *     offload (UDP_SEGMENT): one system call and one trip through the stack for up to 8 packets.
 *     On the wire (or on loopback towards a UDP_GRO receiver) the packets are the same 7232-byte
 *     datagrams.
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif

#include <arpa/inet.h>
#include <errno.h>
#include <netinet/in.h>
#include <netinet/udp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <time.h>
#include <omp.h>
#include <unistd.h>

#include "../../include/b2p_synth.h"
#include "bmf_packet.h"

#ifndef UDP_SEGMENT
#define UDP_SEGMENT 103
#endif
#define GSO_MAX 8 /* 8 x 7232 = 57856 B, under the 65507-byte datagram limit */

static double now_s(void)
{
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

int main(int argc, char **argv)
{
  char dest[64] = "127.0.0.1", prefix[32] = "127.0";
  int port_base = BMF_PORT_BASE, nports = BMF_NPORT_NIC, nchunk = BMF_NCHK_NIC, mode = B2P_SYNTH_GAUSS;
  int epoch = 37, beam = 0, drop_every = 0, arg;
  uint64_t nframes = 64, seed = 1, sec0 = 27 * 1000, idf0 = 0;
  double rate = 2000.0, freq0 = 1173.0;
  uint64_t cache = 0;
  int nthreads = 1, gso = 1;
  while ((arg = getopt(argc, argv, "D:p:P:n:s:m:r:S:i:e:b:f:L:A:C:T:G:h")) != -1) {
    switch (arg) {
      case 'D': snprintf(dest, sizeof(dest), "%s", optarg); break;
      case 'p': port_base = atoi(optarg); break;
      case 'P': nports = atoi(optarg); break;
      case 'n': nframes = strtoull(optarg, NULL, 10); break;
      case 's': seed = strtoull(optarg, NULL, 10); break;
      case 'm': mode = atoi(optarg); break;
      case 'r': rate = atof(optarg); break;
      case 'S': sec0 = strtoull(optarg, NULL, 10); break;
      case 'i': idf0 = strtoull(optarg, NULL, 10); break;
      case 'e': epoch = atoi(optarg); break;
      case 'b': beam = atoi(optarg); break;
      case 'f': freq0 = atof(optarg); break;
      case 'L': drop_every = atoi(optarg); break;
      case 'A': snprintf(prefix, sizeof(prefix), "%s", optarg); break;
      case 'C': cache = strtoull(optarg, NULL, 10); break;
      case 'T': nthreads = atoi(optarg); break;
      case 'G': gso = atoi(optarg); break;
      default:
        fprintf(stdout, "bmf_replay -D dest -p port -P nports -n frames -s seed -m mode -r fps -S sec -i idf -e epoch -b beam -L drop_every\n");
        return EXIT_FAILURE;
    }
  }
  const int chunks_per_port = (nchunk + nports - 1) / nports;
  int *socks = (int *)malloc(sizeof(int) * (size_t)nchunk);
  struct sockaddr_in *dst = (struct sockaddr_in *)calloc((size_t)nchunk, sizeof(struct sockaddr_in));
  for (int c = 0; c < nchunk; ++c) {
    socks[c] = socket(AF_INET, SOCK_DGRAM, 0);
    int snd = 8 << 20;
    setsockopt(socks[c], SOL_SOCKET, SO_SNDBUF, &snd, sizeof(snd));
    unsigned char x, y;
    bmf_source_of_chunk(c, &x, &y);
    char src[64];
    snprintf(src, sizeof(src), "%s.%u.%u", prefix, x, y);
    struct sockaddr_in sa;
    memset(&sa, 0, sizeof(sa));
    sa.sin_family = AF_INET;
    if (inet_pton(AF_INET, src, &sa.sin_addr) != 1 || bind(socks[c], (struct sockaddr *)&sa, sizeof(sa)) < 0) {
      fprintf(stderr, "bmf_replay: can not bind source %s: %s\n", src, strerror(errno));
      return EXIT_FAILURE;
    }
    dst[c].sin_family = AF_INET;
    dst[c].sin_port = htons((uint16_t)(port_base + c / chunks_per_port));
    inet_pton(AF_INET, dest, &dst[c].sin_addr);
  }

  const int nch = 7, nsamp = 128, nchan = nchunk * nch;
  const uint64_t wpp = (uint64_t)nch * nsamp; /* words per packet */
  /* Payloads are generated up front (generation runs at ~0.2 GB/s, line rate is 3.2 GB/s):
     frame f carries cached payload f % ncache, i.e. the stream repeats with that period
     when more frames are sent than cached. */
  const uint64_t ncache = (cache && cache < nframes) ? cache : nframes;
  unsigned char *pool = (unsigned char *)malloc(ncache * (uint64_t)nchunk * BMF_DT_SIZE);
  if (!pool) {
    fprintf(stderr, "bmf_replay: can not allocate the payload cache\n");
    return EXIT_FAILURE;
  }
#pragma omp parallel for schedule(static)
  for (int64_t f = 0; f < (int64_t)ncache; ++f)
    for (int c = 0; c < nchunk; ++c) {
      uint64_t *pay = (uint64_t *)(pool + ((uint64_t)f * nchunk + c) * BMF_DT_SIZE);
      const uint64_t w0 = ((uint64_t)f * (uint64_t)nchunk + (uint64_t)c) * wpp;
      for (uint64_t k = 0; k < wpp; ++k) {
        int16_t v[4];
        b2p_synth_word(seed, w0 + k, c * nch + (int)(k % (uint64_t)nch), nchan, mode, v);
        pay[k] = b2p_synth_pack(v, 1);
      }
    }

  if (nthreads < 1) nthreads = 1;
  if (nthreads > nchunk) nthreads = nchunk;
  uint64_t sent = 0, dropped = 0;
  int failed = 0;
  const double t0 = now_s();
#pragma omp parallel num_threads(nthreads) reduction(+ : sent, dropped)
  {
    const int tid = omp_get_thread_num(), nt = omp_get_num_threads();
    const int c_lo = nchunk * tid / nt, c_hi = nchunk * (tid + 1) / nt;
    unsigned char *pkt = (unsigned char *)malloc((size_t)GSO_MAX * BMF_DF_SIZE);
    const int G = gso < 1 ? 1 : (gso > GSO_MAX ? GSO_MAX : gso);
    for (uint64_t f = 0; f < nframes && !failed; f += (uint64_t)G) {
      const int ng = (nframes - f < (uint64_t)G) ? (int)(nframes - f) : G;
      for (int c = c_lo; c < c_hi; ++c) {
        int npk = 0; /* packets of chunk c, frames f .. f+ng-1, back to back in pkt */
        for (int k = 0; k < ng; ++k) {
          const uint64_t fk = f + (uint64_t)k;
          uint64_t idf = idf0 + fk, sec = sec0;
          sec += (idf / BMF_NDF_PRD) * BMF_PRD_SEC; /* the frame counter wraps every period of 27 s */
          idf %= BMF_NDF_PRD;
          const uint64_t counter = fk * (uint64_t)nchunk + (uint64_t)c + 1;
          if (drop_every > 0 && counter % (uint64_t)drop_every == 0) {
            ++dropped;
            continue;
          }
          unsigned char *q = pkt + (size_t)npk * BMF_DF_SIZE;
          bmf_hdr_t h = {1, idf, sec, epoch, beam, freq0 + 7.0 * c};
          bmf_hdr_encode(q, &h);
          memcpy(q + BMF_HDR_SIZE, pool + ((fk % ncache) * (uint64_t)nchunk + (uint64_t)c) * BMF_DT_SIZE, BMF_DT_SIZE);
          ++npk;
        }
        if (!npk) continue;
        struct iovec iov = {pkt, (size_t)npk * BMF_DF_SIZE};
        struct msghdr mh;
        char ctl[CMSG_SPACE(sizeof(uint16_t))];
        memset(&mh, 0, sizeof(mh));
        mh.msg_name = &dst[c];
        mh.msg_namelen = sizeof(dst[c]);
        mh.msg_iov = &iov;
        mh.msg_iovlen = 1;
        if (npk > 1) { /* one datagram train: the kernel cuts it every BMF_DF_SIZE bytes */
          memset(ctl, 0, sizeof(ctl));
          mh.msg_control = ctl;
          mh.msg_controllen = sizeof(ctl);
          struct cmsghdr *cm = CMSG_FIRSTHDR(&mh);
          cm->cmsg_level = SOL_UDP;
          cm->cmsg_type = UDP_SEGMENT;
          cm->cmsg_len = CMSG_LEN(sizeof(uint16_t));
          const uint16_t seg = BMF_DF_SIZE;
          memcpy(CMSG_DATA(cm), &seg, sizeof(seg));
        }
        while (sendmsg(socks[c], &mh, 0) < 0) {
          if (errno == ENOBUFS || errno == EAGAIN || errno == EINTR) {
            usleep(20);
            continue;
          }
          fprintf(stderr, "bmf_replay: sendmsg: %s\n", strerror(errno));
          failed = 1;
          break;
        }
        sent += (uint64_t)npk;
      }
      if (rate > 0) { /* pace on absolute time so the average rate holds */
        const double due = t0 + (double)(f + (uint64_t)ng) / rate;
        double dt = due - now_s();
        if (dt > 0) {
          struct timespec ts = {(time_t)dt, (long)((dt - (double)(time_t)dt) * 1e9)};
          nanosleep(&ts, NULL);
        }
      }
    }
    free(pkt);
  }
  if (failed) return EXIT_FAILURE;
  const double el = now_s() - t0;
  printf("bmf_replay: %lu packets sent, %lu dropped on purpose, %lu frames in %.3f s (%.1f frames/s, %.3f GB/s, %.2fx line rate)\n",
         (unsigned long)sent, (unsigned long)dropped, (unsigned long)nframes, el, (double)nframes / el,
         (double)sent * BMF_DF_SIZE / el / 1e9, (double)nframes / el * BMF_TDF_SEC);
  return EXIT_SUCCESS;
}
