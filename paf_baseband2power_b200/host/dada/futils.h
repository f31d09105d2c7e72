/* futils.h — shim of the PSRDADA file helpers the reference uses (fileread, diskdb.cu:80). */
#ifndef B2P_FUTILS_H
#define B2P_FUTILS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
/* Read at most bufsz-1 bytes of `filename` into buffer, NUL-pad the rest; <0 on error. */
long fileread(const char *filename, char *buffer, unsigned bufsz);
long filesize(const char *filename);
#ifdef __cplusplus
}
#endif
#endif
