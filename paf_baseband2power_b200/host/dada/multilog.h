/* multilog.h — shim of PSRDADA's multilog: one message to several FILE* sinks.
   Used by the reference at paf_baseband2power.cu:82-84, paf_capture.c:140-142. */
#ifndef B2P_MULTILOG_H
#define B2P_MULTILOG_H

#include <stdio.h>
#include <syslog.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MULTILOG_MAX_SINKS 8

typedef struct multilog_t {
  char name[64];
  int syslog;
  int nsinks;
  FILE *sinks[MULTILOG_MAX_SINKS];
} multilog_t;

multilog_t *multilog_open(const char *program_name, char syslog);
int multilog_add(multilog_t *m, FILE *fptr);
int multilog(multilog_t *m, int priority, const char *format, ...);
int multilog_close(multilog_t *m);

#ifdef __cplusplus
}
#endif
#endif
