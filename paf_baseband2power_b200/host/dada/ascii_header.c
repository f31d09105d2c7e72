#include "ascii_header.h"

#include <ctype.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "dada_def.h"

/* Start of the line whose first token is exactly `keyword`, or NULL. */
static char *find_key(const char *header, const char *keyword)
{
  const size_t klen = strlen(keyword);
  const char *p = header;
  while (p && *p) {
    const char *line = p;
    while (*line == ' ' || *line == '\t') ++line;
    if (strncmp(line, keyword, klen) == 0 && (line[klen] == ' ' || line[klen] == '\t'))
      return (char *)p;
    p = strchr(p, '\n');
    if (p) ++p;
  }
  return NULL;
}

int ascii_header_get(const char *header, const char *keyword, const char *format, ...)
{
  if (!header || !keyword || !format) return -1;
  const char *line = find_key(header, keyword);
  if (!line) return -1;
  while (*line == ' ' || *line == '\t') ++line;
  const char *val = line + strlen(keyword);
  while (*val == ' ' || *val == '\t') ++val;
  char tmp[1024];
  size_t n = 0;
  while (val[n] && val[n] != '\n' && val[n] != '#' && n < sizeof(tmp) - 1) {
    tmp[n] = val[n];
    ++n;
  }
  while (n > 0 && isspace((unsigned char)tmp[n - 1])) --n;
  tmp[n] = 0;
  va_list ap;
  va_start(ap, format);
  int r = vsscanf(tmp, format, ap);
  va_end(ap);
  return r;
}

int ascii_header_del(char *header, const char *keyword)
{
  char *line = header ? find_key(header, keyword) : NULL;
  if (!line) return -1;
  char *end = strchr(line, '\n');
  end = end ? end + 1 : line + strlen(line);
  memmove(line, end, strlen(end) + 1);
  return 0;
}

int ascii_header_set(char *header, const char *keyword, const char *format, ...)
{
  if (!header || !keyword || !format) return -1;
  char value[512];
  va_list ap;
  va_start(ap, format);
  vsnprintf(value, sizeof(value), format, ap);
  va_end(ap);

  /* keep the trailing comment of an existing line */
  char comment[256] = "";
  char *old = find_key(header, keyword);
  if (old) {
    char *eol = strchr(old, '\n');
    size_t len = eol ? (size_t)(eol - old) : strlen(old);
    char *hash = (char *)memchr(old, '#', len);
    if (hash) snprintf(comment, sizeof(comment), "%.*s", (int)(len - (size_t)(hash - old)), hash);
  }
  char line[1024];
  if (comment[0])
    snprintf(line, sizeof(line), "%-12s %-20s %s\n", keyword, value, comment);
  else
    snprintf(line, sizeof(line), "%-12s %s\n", keyword, value);

  const size_t hlen = strlen(header), llen = strlen(line);
  if (old) {
    char *eol = strchr(old, '\n');
    char *rest = eol ? eol + 1 : old + strlen(old);
    const size_t oldlen = (size_t)(rest - old);
    if (hlen - oldlen + llen >= DADA_DEFAULT_HEADER_SIZE) return -2;
    memmove(old + llen, rest, strlen(rest) + 1);
    memcpy(old, line, llen);
  } else {
    if (hlen + llen + 1 >= DADA_DEFAULT_HEADER_SIZE) return -2;
    if (hlen && header[hlen - 1] != '\n') {
      header[hlen] = '\n';
      header[hlen + 1] = 0;
    }
    strcat(header, line);
  }
  return 0;
}
