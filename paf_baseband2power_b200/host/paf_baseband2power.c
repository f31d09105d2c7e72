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

#include <unistd.h>

#include "../../include/b2p.h"
#include "../../include/baseband2power.h"
#include "cli_util.h"

multilog_t *runtime_log;

/* the first eight lines are the reference's usage text as it stands (paf_baseband2power.cu:17-28,
   its spelling included): operators and wrapper scripts know it */
static const char *const kUsage[] = {
    "paf_baseband2power - To detect baseband data with original channels and average the detected data in time",
    "",
    "Usage: paf_process [options]",
    " -a  Hexacdecimal shared memory key for incoming ring buffer",
    " -b  Hexacdecimal shared memory key for outcoming ring buffer",
    " -c  The name of the directory in which we will record the data",
    " -d  The index of GPU",
    " -h  show help",
    "extensions:",
    " -d  may be a list (0,1,2,3): the beam's channel groups are spread over those GPUs",
    " -g  chunks per GPU for a -d list, e.g. 5,5,7,7 (default: in proportion to the host-link rates)",
    " -s  0 integral over the integration (default), 1 average in time",
    " -n  data frames per integration (default: frames of one input block)",
    " -k  kernel: auto | ldg | tma",
    " -e  1 big-endian samples (default), 0 little-endian",
    " -p  1 page-lock the input ring (default), 0 leave it pageable",
    NULL};

static int kernel_by_name(const char *name)
{
  if (!strcmp(name, "tma")) return B2P_KERNEL_TMA;
  if (!strcmp(name, "ldg")) return B2P_KERNEL_LDG;
  return B2P_KERNEL_AUTO;
}

/* returns 0 to go on, 1 to leave with EXIT_FAILURE (help or a bad option) */
static int parse_args(int argc, char *argv[], conf_t *conf)
{
  for (int opt; (opt = getopt(argc, argv, "a:b:c:d:h::s:n:k:e:p:g:")) != -1;) {
    if (opt == 'a' || opt == 'b') {
      if (cli_hex_key(optarg, opt == 'a' ? &conf->key_in : &conf->key_out, __FILE__, __LINE__)) return 1;
    } else if (opt == 'c') {
      cli_copy(conf->dir, MSTR_LEN, optarg);
    } else if (opt == 'd') {
      conf->device_id = atoi(optarg);
      conf->ngpu = cli_int_list(optarg, conf->gpus, B2P_STAGE_MAX_GPUS);
    } else if (opt == 'g') {
      cli_int_list(optarg, conf->gpu_chunks, B2P_STAGE_MAX_GPUS);
    } else if (opt == 's') {
      conf->average = atoi(optarg) != 0;
    } else if (opt == 'n') {
      conf->ndf_integration = strtoull(optarg, NULL, 10);
    } else if (opt == 'k') {
      conf->kernel = kernel_by_name(optarg);
    } else if (opt == 'e') {
      conf->big_endian = atoi(optarg) != 0;
    } else if (opt == 'p') {
      conf->pin_ring = atoi(optarg) != 0;
    } else { /* -h and anything unknown */
      cli_print_lines(stdout, kUsage);
      return 1;
    }
  }
  return 0;
}

int main(int argc, char *argv[])
{
  conf_t conf;
  default_baseband2power(&conf);
  if (parse_args(argc, argv, &conf)) return EXIT_FAILURE;

  FILE *fp_log = NULL;
  runtime_log = cli_open_log(conf.dir, "paf_baseband2power", &fp_log);
  if (!runtime_log) return EXIT_FAILURE;
  multilog(runtime_log, LOG_INFO, "START PAF_PROCESS\n");
  conf.log = runtime_log;

  /* a container that exposes a single GPU numbers it 0 whatever -d says */
  if (b2p_device_count() == 1) {
    conf.device_id = 0;
    for (int i = 0; i < conf.ngpu; ++i) conf.gpus[i] = 0;
  }

  int status = init_baseband2power(&conf);
  if (status == EXIT_SUCCESS) status = do_baseband2power(&conf);
  destroy_baseband2power(&conf);

  multilog(runtime_log, LOG_INFO, "FINISH PAF_PROCESS\n");
  multilog_close(runtime_log);
  fclose(fp_log);
  return status;
}
