/*
 * paf_diskdb — read a DADA data file into shared memory.
 * Flags as in the reference (paf_diskdb.cu:10-22,30-63): -a key, -b directory,
 * -c data file, -d header file, -e start-of-data, -h help.
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif

#include <unistd.h>

#include "cli_util.h"
#include "diskdb.h"

static const char *const kUsage[] = {
    "paf_diskdb - read dada data file into shared memory",
    "",
    "Usage: paf_diskdb [options]",
    " -a Hexadecimal shared memory key for capture",
    " -b Directory with data file",
    " -c The name of data file",
    " -d The name of header file",
    " -e Enable start-of-data or not",
    " -h Show help",
    NULL};

int main(int argc, char **argv)
{
  char directory[MSTR_LEN] = ".", datafile[MSTR_LEN] = "";
  diskdb_conf_t conf;
  memset(&conf, 0, sizeof(conf));
  conf.key = 0xdada; /* paf-baseband2power.conf:13 */
  conf.sod = 1;      /* paf-baseband2power.conf:16 */

  for (int opt; (opt = getopt(argc, argv, "a:b:c:d:e:h")) != -1;) {
    if (opt == 'a') {
      if (cli_hex_key(optarg, &conf.key, __FILE__, __LINE__)) return EXIT_FAILURE;
    } else if (opt == 'b') {
      cli_copy(directory, MSTR_LEN, optarg);
    } else if (opt == 'c') {
      cli_copy(datafile, MSTR_LEN, optarg);
    } else if (opt == 'd') {
      cli_copy(conf.hfname, MSTR_LEN, optarg);
    } else if (opt == 'e') {
      conf.sod = atoi(optarg);
    } else {
      cli_print_lines(stdout, kUsage);
      return EXIT_FAILURE;
    }
  }
  snprintf(conf.fname, sizeof(conf.fname), "%s/%s", directory, datafile);
  conf.log = multilog_open("paf_diskdb", 0);
  multilog_add(conf.log, stderr);

  int status = init_diskdb(&conf);
  if (status == EXIT_SUCCESS) status = do_diskdb(&conf);
  destroy_diskdb(&conf);
  multilog_close(conf.log);
  return status;
}
