#include "multilog.h"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

multilog_t *multilog_open(const char *program_name, char use_syslog)
{
  multilog_t *m = (multilog_t *)calloc(1, sizeof(multilog_t));
  if (!m) return NULL;
  snprintf(m->name, sizeof(m->name), "%s", program_name ? program_name : "");
  m->syslog = use_syslog; /* recorded only: the shim never talks to syslogd */
  return m;
}

int multilog_add(multilog_t *m, FILE *fptr)
{
  if (!m || !fptr || m->nsinks >= MULTILOG_MAX_SINKS) return -1;
  m->sinks[m->nsinks++] = fptr;
  return 0;
}

int multilog(multilog_t *m, int priority, const char *format, ...)
{
  if (!m) return -1;
  char stamp[32];
  time_t now = time(NULL);
  struct tm tmv;
  gmtime_r(&now, &tmv);
  strftime(stamp, sizeof(stamp), "%Y-%m-%d-%H:%M:%S", &tmv);
  const char *lvl = priority <= LOG_ERR ? "ERR" : (priority == LOG_WARNING ? "WARN" : "INFO");
  for (int i = 0; i < m->nsinks; ++i) {
    va_list ap;
    va_start(ap, format);
    fprintf(m->sinks[i], "[%s] %s %s: ", stamp, m->name, lvl);
    vfprintf(m->sinks[i], format, ap);
    va_end(ap);
    fflush(m->sinks[i]);
  }
  return 0;
}

int multilog_close(multilog_t *m)
{
  free(m);
  return 0;
}
