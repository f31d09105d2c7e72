/* ascii_header.h — shim of PSRDADA's "KEY value" header accessors
   (capture.c:758-781 sets UTC_START, PICOSECONDS, FREQ this way). */
#ifndef B2P_ASCII_HEADER_H
#define B2P_ASCII_HEADER_H
#ifdef __cplusplus
extern "C" {
#endif
/* Replace the value of `keyword` (or append a new line); returns 0, <0 on error. */
int ascii_header_set(char *header, const char *keyword, const char *format, ...);
/* sscanf the value of `keyword`; returns the number of items parsed, <0 if absent. */
int ascii_header_get(const char *header, const char *keyword, const char *format, ...);
/* Remove the line holding `keyword`; returns 0, <0 if absent. */
int ascii_header_del(char *header, const char *keyword);
#ifdef __cplusplus
}
#endif
#endif
