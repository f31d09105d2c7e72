/* cli_util.h — small helpers shared by the executables' mains. */
#ifndef B2P_CLI_UTIL_H
#define B2P_CLI_UTIL_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>

#include "dada/multilog.h"

/* "-a dada": ring keys are given in hexadecimal, as for every PSRDADA tool. */
static inline int cli_hex_key(const char *text, key_t *key, const char *file, int line)
{
  unsigned v = 0;
  char tail = 0;
  if (!text || sscanf(text, "%x%c", &v, &tail) != 1) {
    fprintf(stderr, "Could not parse key from %s, which happens at \"%s\", line [%d].\n", text ? text : "(null)", file, line);
    return -1;
  }
  *key = (key_t)v;
  return 0;
}

static inline void cli_copy(char *dst, size_t cap, const char *src)
{
  snprintf(dst, cap, "%s", src ? src : "");
}

/* "0,1,2,3" -> out[]; returns how many integers were read (at most cap) */
static inline int cli_int_list(const char *text, int *out, int cap)
{
  int n = 0;
  for (const char *p = text; p && *p && n < cap;) {
    char *end = NULL;
    const long v = strtol(p, &end, 10);
    if (end == p) break;
    out[n++] = (int)v;
    p = (*end == ',' || *end == ':') ? end + 1 : end;
    if (end == p && *end) break;
  }
  return n;
}

static inline void cli_print_lines(FILE *fp, const char *const *lines)
{
  for (; *lines; ++lines) fprintf(fp, "%s\n", *lines);
}

/* <dir>/<program>.log opened for appending and attached to a new multilog. */
static inline multilog_t *cli_open_log(const char *dir, const char *program, FILE **fp_out)
{
  char path[1200];
  snprintf(path, sizeof(path), "%s/%s.log", dir, program);
  FILE *fp = fopen(path, "ab+");
  if (!fp) {
    fprintf(stderr, "Can not open log file %s\n", path);
    return NULL;
  }
  multilog_t *log = multilog_open(program, 1);
  multilog_add(log, fp);
  *fp_out = fp;
  return log;
}

#endif
