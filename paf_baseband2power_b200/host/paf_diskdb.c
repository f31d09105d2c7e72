/*
 * paf_diskdb — read a DADA data file into shared memory.
 * Flags as in the reference (paf_diskdb.cu:10-22,30-63): -a key, -b directory,
 * -c data file, -d header file, -e start-of-data, -h help.
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "diskdb.h"

static void usage(void)
{
  fprintf(stdout,
          "paf_diskdb - read dada data file into shared memory \n"
          "\n"
          "Usage: paf_diskdb [options]\n"
          " -a Hexadecimal shared memory key for capture \n"
          " -b Directory with data file \n"
          " -c The name of data file    \n"
          " -d The name of header file  \n"
          " -e Enable start-of-data or not \n"
          " -h Show help    \n");
}

int main(int argc, char **argv)
{
  int arg;
  char fdir[MSTR_LEN] = ".", fname[MSTR_LEN] = "";
  diskdb_conf_t conf;
  memset(&conf, 0, sizeof(conf));
  conf.key = 0xdada;
  conf.sod = 1;

  while ((arg = getopt(argc, argv, "a:b:c:d:e:h")) != -1) {
    switch (arg) {
      case 'h':
        usage();
        return EXIT_FAILURE;
      case 'a':
        if (sscanf(optarg, "%x", (unsigned *)&conf.key) != 1) {
          fprintf(stderr, "Could not parse key from %s, which happens at \"%s\", line [%d].\n", optarg, __FILE__, __LINE__);
          return EXIT_FAILURE;
        }
        break;
      case 'b':
        snprintf(fdir, MSTR_LEN, "%s", optarg);
        break;
      case 'c':
        snprintf(fname, MSTR_LEN, "%s", optarg);
        break;
      case 'd':
        snprintf(conf.hfname, MSTR_LEN, "%s", optarg);
        break;
      case 'e':
        conf.sod = atoi(optarg);
        break;
      default:
        usage();
        return EXIT_FAILURE;
    }
  }
  snprintf(conf.fname, sizeof(conf.fname), "%s/%s", fdir, fname);
  conf.log = multilog_open("paf_diskdb", 0);
  multilog_add(conf.log, stderr);

  int rc = init_diskdb(&conf);
  if (rc == EXIT_SUCCESS) rc = do_diskdb(&conf);
  destroy_diskdb(&conf);
  multilog_close(conf.log);
  return rc;
}
