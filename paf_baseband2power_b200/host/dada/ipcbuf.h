/*
 * ipcbuf.h — ring of fixed-size buffers in SysV shared memory (PSRDADA-named shim).
 *
 * Names and call semantics follow the PSRDADA calls used by the reference
 * (ipcbuf_get_bufsz, ipcbuf_enable_sod/disable_sod, ipcbuf_get_next_write,
 * ipcbuf_mark_filled: diskdb.cu:34-36,54,62,79,88; capture.c:590-633) plus the
 * reader side (ipcbuf_get_next_read, ipcbuf_mark_cleared, ipcbuf_eod).  One
 * writer, one reader (NREADER 1, paf-baseband2power.conf:15,22).
 *
 * Layout: segment `key` holds an ipcsync_t (counters + a process-shared robust
 * mutex and condition variable); every buffer is its own segment (ids kept in
 * the sync block) so that single buffers can be page-locked for CUDA.
 */
#ifndef B2P_IPCBUF_H
#define B2P_IPCBUF_H

#include <pthread.h>
#include <stdint.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPCBUF_MAX_BUFS 256
#define IPCBUF_MAGIC    0x42325042u /* "B2PB" */

typedef struct ipcsync_t {
  uint32_t magic;
  uint32_t nreaders;
  uint64_t nbufs, bufsz;
  pthread_mutex_t mtx;
  pthread_cond_t cv;
  uint64_t w_count;   /* buffers marked filled since creation */
  uint64_t r_count;   /* buffers marked cleared since creation */
  int sod;            /* start of data raised */
  int eod;            /* end of data raised by the writer ... */
  uint64_t eod_count; /* ... valid data ends after this many filled buffers */
  int writer_locked, reader_locked;
  int shmid[IPCBUF_MAX_BUFS];
  uint64_t fill[IPCBUF_MAX_BUFS]; /* valid bytes in each buffer */
} ipcsync_t;

typedef struct ipcbuf_t {
  key_t key;
  int syncid;
  ipcsync_t *sync;
  char **buffer;
  int is_writer, is_reader;
  uint64_t last_read_bytes;
} ipcbuf_t;

#define IPCBUF_INIT {0, -1, 0, 0, 0, 0, 0}

int ipcbuf_create(ipcbuf_t *id, key_t key, uint64_t nbufs, uint64_t bufsz, unsigned nreaders);
int ipcbuf_connect(ipcbuf_t *id, key_t key);
int ipcbuf_disconnect(ipcbuf_t *id);
int ipcbuf_destroy(ipcbuf_t *id); /* connected: remove every segment */

uint64_t ipcbuf_get_bufsz(ipcbuf_t *id);
uint64_t ipcbuf_get_nbufs(ipcbuf_t *id);
uint64_t ipcbuf_get_write_count(ipcbuf_t *id);
uint64_t ipcbuf_get_read_count(ipcbuf_t *id);

int ipcbuf_lock_write(ipcbuf_t *id);
int ipcbuf_unlock_write(ipcbuf_t *id);
int ipcbuf_lock_read(ipcbuf_t *id);
int ipcbuf_unlock_read(ipcbuf_t *id);

int ipcbuf_enable_sod(ipcbuf_t *id, uint64_t st_buf, uint64_t st_byte);
int ipcbuf_disable_sod(ipcbuf_t *id);
int ipcbuf_enable_eod(ipcbuf_t *id);
int ipcbuf_eod(ipcbuf_t *id); /* reader: 1 once every valid buffer has been cleared */
int ipcbuf_reset(ipcbuf_t *id); /* writer: forget a finished observation */

char *ipcbuf_get_next_write(ipcbuf_t *id);
/* Shim extension: the buffer `ahead` places after the next one to fill (0 = same as
   ipcbuf_get_next_write); waits until the reader has freed it.  Lets the capture stage keep
   two blocks open so late packets of block k and early ones of k+1 both land in the ring
   (the reference spills into a malloc'ed side buffer and copies, capture.c:527-533). */
#ifndef B2P_NO_SHIM_EXTENSIONS /* define to compile a client as against a stock PSRDADA */
char *ipcbuf_get_write_ahead(ipcbuf_t *id, unsigned ahead);
#endif
int ipcbuf_mark_filled(ipcbuf_t *id, uint64_t nbytes);
char *ipcbuf_get_next_read(ipcbuf_t *id, uint64_t *bytes);
int ipcbuf_mark_cleared(ipcbuf_t *id);

/* every buffer of the ring, for page-locking (dada_cuda_dbregister) */
char *ipcbuf_get_buffer(ipcbuf_t *id, uint64_t index);

#ifdef __cplusplus
}
#endif
#endif
