/*
 * dada_def.h — constants of the ring-buffer shim.
 *
 * PSRDADA is an external, un-vendored dependency of the reference
 * (-lpsrdada, makefile:27) and is absent from this environment, so the
 * executables are written against this in-repo shim, which exposes the subset of
 * PSRDADA names the reference's code calls (diskdb.cu:24-50,79-88,105-109,
 * 128-130; capture.c:316,590-633,733-781; sync.c:101,109) plus the reader side
 * the unwritten stage needs.  The implementation is new (process-shared pthread
 * primitives in a SysV segment); only the names and call semantics follow
 * PSRDADA so that a site with the real library can link against it instead.
 */
#ifndef B2P_DADA_DEF_H
#define B2P_DADA_DEF_H

#define DADA_DEFAULT_BLOCK_KEY   0x0000dada
#define DADA_DEFAULT_HEADER_SIZE 4096
#define DADA_DEFAULT_HDR_NBUFS   8
#define DADA_TIMESTR             "%Y-%m-%d-%H:%M:%S"

#endif
