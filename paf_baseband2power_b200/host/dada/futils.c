#include "futils.h"

#include <stdio.h>
#include <string.h>
#include <sys/stat.h>

long filesize(const char *filename)
{
  struct stat st;
  if (!filename || stat(filename, &st) < 0) return -1;
  return (long)st.st_size;
}

long fileread(const char *filename, char *buffer, unsigned bufsz)
{
  if (!filename || !buffer || bufsz == 0) return -1;
  FILE *fp = fopen(filename, "r");
  if (!fp) return -1;
  memset(buffer, 0, bufsz);
  size_t n = fread(buffer, 1, bufsz - 1, fp);
  fclose(fp);
  return (long)n;
}
