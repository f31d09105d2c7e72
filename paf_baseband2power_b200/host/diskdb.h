/*
 * diskdb.h — read a DADA data file into the input ring (the paf_diskdb stage).
 * Same entry points and conf fields as the reference (diskdb.cuh:19-34):
 * init_diskdb / do_diskdb / destroy_diskdb, returning EXIT_SUCCESS/EXIT_FAILURE.
 * Differences, all defect fixes: the log is initialised
 * before use, return codes are honoured by main, conf is passed by pointer.
 */
#ifndef DISKDB_H
#define DISKDB_H

#include <stdio.h>
#include <sys/types.h>

#include "dada/dada_hdu.h"
#include "dada/multilog.h"

#define DADA_HDR_SIZE 4096 /* diskdb.cuh:17 */
#ifndef MSTR_LEN
#define MSTR_LEN 512
#endif

typedef struct diskdb_conf_t {
  key_t key;
  int sod;
  char fname[2 * MSTR_LEN], hfname[MSTR_LEN];
  FILE *fp;
  dada_hdu_t *hdu;
  multilog_t *log;
  size_t hdrsz;
  size_t rbufsz;
  unsigned long nblocks, nbytes;
} diskdb_conf_t;

int init_diskdb(diskdb_conf_t *conf);
int do_diskdb(diskdb_conf_t *conf);
int destroy_diskdb(diskdb_conf_t *conf);

#endif
