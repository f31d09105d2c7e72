/*
 * paf_baseband2power — detect baseband data with the original channels and
 * integrate the detected data in time.
 *
 * Program surface of the reference executable (paf_baseband2power.cu:17-28,
 * 40-72, 75-90): -a/-b hexadecimal ring keys, -c directory (log file
 * <dir>/paf_baseband2power.log), -d GPU index (forced to 0 when only one GPU is
 * visible), -h help.  The reference main returns right after that; this one
 * runs the stage.  Extra, optional flags: -s, -n, -k, -e, -p.
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "../../include/b2p.h"
#include "../../include/baseband2power.h"
#include "dada/multilog.h"

multilog_t *runtime_log;

static void usage(void)
{
  fprintf(stdout,
          "paf_baseband2power - To detect baseband data with original channels and integrate the "
          "detected data in time\n"
          "\n"
          "Usage: paf_baseband2power [options]\n"
          " -a  Hexadecimal shared memory key for incoming ring buffer\n"
          " -b  Hexadecimal shared memory key for outcoming ring buffer\n"
          " -c  The name of the directory in which we will record the data\n"
          " -d  The index of GPU\n"
          " -h  show help\n"
          "extensions:\n"
          " -s  0 integral over the integration (default), 1 average in time\n"
          " -n  data frames per integration (default: frames of one input block)\n"
          " -k  kernel: auto | ldg | tma\n"
          " -e  1 big-endian samples (default), 0 little-endian\n"
          " -p  1 page-lock the input ring (default), 0 leave it pageable\n");
}

int main(int argc, char *argv[])
{
  int arg;
  conf_t conf;
  default_baseband2power(&conf);

  while ((arg = getopt(argc, argv, "a:b:c:d:hs:n:k:e:p:")) != -1) {
    switch (arg) {
      case 'h':
        usage();
        return EXIT_FAILURE;
      case 'a':
        if (sscanf(optarg, "%x", (unsigned *)&conf.key_in) != 1) {
          fprintf(stderr, "Could not parse key from %s, which happens at \"%s\", line [%d].\n", optarg, __FILE__, __LINE__);
          return EXIT_FAILURE;
        }
        break;
      case 'b':
        if (sscanf(optarg, "%x", (unsigned *)&conf.key_out) != 1) {
          fprintf(stderr, "Could not parse key from %s, which happens at \"%s\", line [%d].\n", optarg, __FILE__, __LINE__);
          return EXIT_FAILURE;
        }
        break;
      case 'c':
        snprintf(conf.dir, MSTR_LEN, "%s", optarg);
        break;
      case 'd':
        conf.device_id = atoi(optarg);
        break;
      case 's':
        conf.average = atoi(optarg) ? 1 : 0;
        break;
      case 'n':
        conf.ndf_integration = strtoull(optarg, NULL, 10);
        break;
      case 'k':
        conf.kernel = !strcmp(optarg, "tma") ? B2P_KERNEL_TMA : (!strcmp(optarg, "ldg") ? B2P_KERNEL_LDG : B2P_KERNEL_AUTO);
        break;
      case 'e':
        conf.big_endian = atoi(optarg) ? 1 : 0;
        break;
      case 'p':
        conf.pin_ring = atoi(optarg) ? 1 : 0;
        break;
      default:
        usage();
        return EXIT_FAILURE;
    }
  }

  /* Setup log interface */
  char log_fname[MSTR_LEN + 64];
  snprintf(log_fname, sizeof(log_fname), "%s/paf_baseband2power.log", conf.dir);
  FILE *fp_log = fopen(log_fname, "ab+");
  if (fp_log == NULL) {
    fprintf(stderr, "Can not open log file %s\n", log_fname);
    return EXIT_FAILURE;
  }
  runtime_log = multilog_open("paf_baseband2power", 1);
  multilog_add(runtime_log, fp_log);
  multilog(runtime_log, LOG_INFO, "START PAF_PROCESS\n");
  conf.log = runtime_log;

  /* one GPU exposed to the container: its index is 0 whatever -d says */
  if (b2p_device_count() == 1) conf.device_id = 0;

  int rc = init_baseband2power(&conf);
  if (rc == EXIT_SUCCESS) rc = do_baseband2power(&conf);
  destroy_baseband2power(&conf);

  multilog(runtime_log, LOG_INFO, "FINISH PAF_PROCESS\n");
  multilog_close(runtime_log);
  fclose(fp_log);
  return rc;
}
