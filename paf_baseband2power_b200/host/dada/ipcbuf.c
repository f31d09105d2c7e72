#include "ipcbuf.h"

#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/ipc.h>
#include <sys/shm.h>
#include <time.h>

static int lock(ipcsync_t *s)
{
  int r = pthread_mutex_lock(&s->mtx);
  if (r == EOWNERDEAD) { /* the peer died holding the lock: state is still consistent for us */
    pthread_mutex_consistent(&s->mtx);
    r = 0;
  }
  return r;
}
static void unlock(ipcsync_t *s) { pthread_mutex_unlock(&s->mtx); }

/* Wait for a state change, but wake at least every 200 ms so that a waiter notices when the
   ring has been destroyed under it (magic cleared) instead of sleeping forever.
   Returns 0 while the ring is alive. */
static int wait_change(ipcsync_t *s)
{
  struct timespec ts;
  clock_gettime(CLOCK_REALTIME, &ts);
  ts.tv_nsec += 200000000L;
  if (ts.tv_nsec >= 1000000000L) {
    ts.tv_sec += 1;
    ts.tv_nsec -= 1000000000L;
  }
  int r = pthread_cond_timedwait(&s->cv, &s->mtx, &ts);
  if (r == EOWNERDEAD) pthread_mutex_consistent(&s->mtx);
  return s->magic == IPCBUF_MAGIC ? 0 : -1;
}

static int attach_buffers(ipcbuf_t *id)
{
  ipcsync_t *s = id->sync;
  id->buffer = (char **)calloc(s->nbufs, sizeof(char *));
  if (!id->buffer) return -1;
  for (uint64_t i = 0; i < s->nbufs; ++i) {
    void *p = shmat(s->shmid[i], NULL, 0);
    if (p == (void *)-1) {
      fprintf(stderr, "ipcbuf: shmat buffer %lu: %s\n", (unsigned long)i, strerror(errno));
      return -1;
    }
    id->buffer[i] = (char *)p;
  }
  return 0;
}

int ipcbuf_create(ipcbuf_t *id, key_t key, uint64_t nbufs, uint64_t bufsz, unsigned nreaders)
{
  if (!id || nbufs == 0 || nbufs > IPCBUF_MAX_BUFS || bufsz == 0) return -1;
  if (nreaders != 1) {
    fprintf(stderr, "ipcbuf: this shim supports exactly one reader (asked for %u)\n", nreaders);
    return -1;
  }
  int sid = shmget(key, sizeof(ipcsync_t), IPC_CREAT | IPC_EXCL | 0666);
  if (sid < 0) {
    fprintf(stderr, "ipcbuf: shmget key %x: %s\n", (unsigned)key, strerror(errno));
    return -1;
  }
  ipcsync_t *s = (ipcsync_t *)shmat(sid, NULL, 0);
  if (s == (void *)-1) return -1;
  memset(s, 0, sizeof(*s));
  s->nbufs = nbufs;
  s->bufsz = bufsz;
  s->nreaders = nreaders;
  pthread_mutexattr_t ma;
  pthread_mutexattr_init(&ma);
  pthread_mutexattr_setpshared(&ma, PTHREAD_PROCESS_SHARED);
  pthread_mutexattr_setrobust(&ma, PTHREAD_MUTEX_ROBUST);
  pthread_mutex_init(&s->mtx, &ma);
  pthread_mutexattr_destroy(&ma);
  pthread_condattr_t ca;
  pthread_condattr_init(&ca);
  pthread_condattr_setpshared(&ca, PTHREAD_PROCESS_SHARED);
  pthread_cond_init(&s->cv, &ca);
  pthread_condattr_destroy(&ca);
  for (uint64_t i = 0; i < nbufs; ++i) {
    s->shmid[i] = shmget(IPC_PRIVATE, bufsz, IPC_CREAT | 0666);
    if (s->shmid[i] < 0) {
      fprintf(stderr, "ipcbuf: shmget buffer %lu (%lu B): %s\n", (unsigned long)i,
              (unsigned long)bufsz, strerror(errno));
      for (uint64_t j = 0; j < i; ++j) shmctl(s->shmid[j], IPC_RMID, NULL);
      shmdt(s);
      shmctl(sid, IPC_RMID, NULL);
      return -1;
    }
  }
  s->magic = IPCBUF_MAGIC;
  id->key = key;
  id->syncid = sid;
  id->sync = s;
  id->is_writer = id->is_reader = 0;
  return attach_buffers(id);
}

int ipcbuf_connect(ipcbuf_t *id, key_t key)
{
  if (!id) return -1;
  int sid = shmget(key, sizeof(ipcsync_t), 0666);
  if (sid < 0) {
    fprintf(stderr, "ipcbuf: no ring with key %x: %s\n", (unsigned)key, strerror(errno));
    return -1;
  }
  ipcsync_t *s = (ipcsync_t *)shmat(sid, NULL, 0);
  if (s == (void *)-1) return -1;
  if (s->magic != IPCBUF_MAGIC) {
    fprintf(stderr, "ipcbuf: segment %x is not a ring of this shim\n", (unsigned)key);
    shmdt(s);
    return -1;
  }
  id->key = key;
  id->syncid = sid;
  id->sync = s;
  id->is_writer = id->is_reader = 0;
  return attach_buffers(id);
}

int ipcbuf_disconnect(ipcbuf_t *id)
{
  if (!id || !id->sync) return -1;
  for (uint64_t i = 0; id->buffer && i < id->sync->nbufs; ++i)
    if (id->buffer[i]) shmdt(id->buffer[i]);
  free(id->buffer);
  id->buffer = NULL;
  shmdt(id->sync);
  id->sync = NULL;
  return 0;
}

int ipcbuf_destroy(ipcbuf_t *id)
{
  if (!id || !id->sync) return -1;
  ipcsync_t *s = id->sync;
  lock(s);
  s->magic = 0; /* tell every attached waiter that the ring is gone */
  pthread_cond_broadcast(&s->cv);
  unlock(s);
  for (uint64_t i = 0; i < s->nbufs; ++i) shmctl(s->shmid[i], IPC_RMID, NULL);
  int sid = id->syncid;
  ipcbuf_disconnect(id);
  shmctl(sid, IPC_RMID, NULL);
  return 0;
}

uint64_t ipcbuf_get_bufsz(ipcbuf_t *id) { return id && id->sync ? id->sync->bufsz : 0; }
uint64_t ipcbuf_get_nbufs(ipcbuf_t *id) { return id && id->sync ? id->sync->nbufs : 0; }
uint64_t ipcbuf_get_write_count(ipcbuf_t *id) { return id && id->sync ? id->sync->w_count : 0; }
uint64_t ipcbuf_get_read_count(ipcbuf_t *id) { return id && id->sync ? id->sync->r_count : 0; }
char *ipcbuf_get_buffer(ipcbuf_t *id, uint64_t i)
{
  return (id && id->sync && i < id->sync->nbufs) ? id->buffer[i] : NULL;
}

int ipcbuf_lock_write(ipcbuf_t *id)
{
  if (!id || !id->sync) return -1;
  ipcsync_t *s = id->sync;
  lock(s);
  if (s->writer_locked) {
    unlock(s);
    fprintf(stderr, "ipcbuf: ring %x already has a writer\n", (unsigned)id->key);
    return -1;
  }
  /* a finished, fully drained observation is forgotten when a new writer arrives */
  if (s->eod && s->r_count >= s->eod_count) {
    s->eod = 0;
    s->sod = 0;
  }
  s->writer_locked = 1;
  unlock(s);
  id->is_writer = 1;
  return 0;
}

int ipcbuf_unlock_write(ipcbuf_t *id)
{
  if (!id || !id->sync || !id->is_writer) return -1;
  lock(id->sync);
  id->sync->writer_locked = 0;
  pthread_cond_broadcast(&id->sync->cv);
  unlock(id->sync);
  id->is_writer = 0;
  return 0;
}

int ipcbuf_lock_read(ipcbuf_t *id)
{
  if (!id || !id->sync) return -1;
  ipcsync_t *s = id->sync;
  lock(s);
  if (s->reader_locked) {
    unlock(s);
    fprintf(stderr, "ipcbuf: ring %x already has a reader\n", (unsigned)id->key);
    return -1;
  }
  s->reader_locked = 1;
  unlock(s);
  id->is_reader = 1;
  return 0;
}

int ipcbuf_unlock_read(ipcbuf_t *id)
{
  if (!id || !id->sync || !id->is_reader) return -1;
  lock(id->sync);
  id->sync->reader_locked = 0;
  pthread_cond_broadcast(&id->sync->cv);
  unlock(id->sync);
  id->is_reader = 0;
  return 0;
}

int ipcbuf_enable_sod(ipcbuf_t *id, uint64_t st_buf, uint64_t st_byte)
{
  (void)st_buf;
  (void)st_byte; /* the reference always starts at the beginning (diskdb.cu:54) */
  if (!id || !id->sync || !id->is_writer) return -1;
  lock(id->sync);
  id->sync->sod = 1;
  pthread_cond_broadcast(&id->sync->cv);
  unlock(id->sync);
  return 0;
}

int ipcbuf_disable_sod(ipcbuf_t *id)
{
  if (!id || !id->sync || !id->is_writer) return -1;
  lock(id->sync);
  id->sync->sod = 0;
  unlock(id->sync);
  return 0;
}

int ipcbuf_enable_eod(ipcbuf_t *id)
{
  if (!id || !id->sync || !id->is_writer) return -1;
  ipcsync_t *s = id->sync;
  lock(s);
  if (!s->eod) {
    s->eod = 1;
    s->eod_count = s->w_count;
  }
  pthread_cond_broadcast(&s->cv);
  unlock(s);
  return 0;
}

int ipcbuf_reset(ipcbuf_t *id)
{
  if (!id || !id->sync || !id->is_writer) return -1;
  ipcsync_t *s = id->sync;
  lock(s);
  s->eod = 0;
  s->sod = 0;
  s->r_count = s->w_count;
  unlock(s);
  return 0;
}

int ipcbuf_eod(ipcbuf_t *id)
{
  if (!id || !id->sync) return -1;
  ipcsync_t *s = id->sync;
  lock(s);
  int e = s->eod && s->r_count >= s->eod_count;
  unlock(s);
  return e;
}

char *ipcbuf_get_next_write(ipcbuf_t *id)
{
  if (!id || !id->sync || !id->is_writer) return NULL;
  ipcsync_t *s = id->sync;
  lock(s);
  while (s->w_count - s->r_count >= s->nbufs)
    if (wait_change(s) < 0) {
      unlock(s);
      return NULL;
    }
  char *p = id->buffer[s->w_count % s->nbufs];
  unlock(s);
  return p;
}

char *ipcbuf_get_write_ahead(ipcbuf_t *id, unsigned ahead)
{
  if (!id || !id->sync || !id->is_writer) return NULL;
  ipcsync_t *s = id->sync;
  if ((uint64_t)ahead + 1 > s->nbufs) return NULL;
  lock(s);
  while (s->w_count + ahead - s->r_count >= s->nbufs)
    if (wait_change(s) < 0) {
      unlock(s);
      return NULL;
    }
  char *p = id->buffer[(s->w_count + ahead) % s->nbufs];
  unlock(s);
  return p;
}

int ipcbuf_mark_filled(ipcbuf_t *id, uint64_t nbytes)
{
  if (!id || !id->sync || !id->is_writer) return -1;
  ipcsync_t *s = id->sync;
  if (nbytes > s->bufsz) return -1;
  lock(s);
  s->fill[s->w_count % s->nbufs] = nbytes;
  s->w_count++;
  if (nbytes < s->bufsz && !s->eod) { /* a short buffer ends the data */
    s->eod = 1;
    s->eod_count = s->w_count;
  }
  pthread_cond_broadcast(&s->cv);
  unlock(s);
  return 0;
}

char *ipcbuf_get_next_read(ipcbuf_t *id, uint64_t *bytes)
{
  if (!id || !id->sync || !id->is_reader) return NULL;
  ipcsync_t *s = id->sync;
  lock(s);
  for (;;) {
    if (s->r_count < s->w_count && (!s->eod || s->r_count < s->eod_count)) break;
    if (s->eod && s->r_count >= s->eod_count) {
      unlock(s);
      if (bytes) *bytes = 0;
      return NULL;
    }
    if (wait_change(s) < 0) {
      unlock(s);
      if (bytes) *bytes = 0;
      return NULL;
    }
  }
  const uint64_t i = s->r_count % s->nbufs;
  id->last_read_bytes = s->fill[i];
  if (bytes) *bytes = s->fill[i];
  char *p = id->buffer[i];
  unlock(s);
  return p;
}

int ipcbuf_mark_cleared(ipcbuf_t *id)
{
  if (!id || !id->sync || !id->is_reader) return -1;
  ipcsync_t *s = id->sync;
  lock(s);
  if (s->r_count >= s->w_count) {
    unlock(s);
    return -1;
  }
  s->r_count++;
  pthread_cond_broadcast(&s->cv);
  unlock(s);
  return 0;
}
