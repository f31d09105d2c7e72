/* daemon.h — placeholder: the reference includes it (capture.h:12, diskdb.cuh:11) and calls nothing from it. */
#ifndef B2P_DAEMON_H
#define B2P_DAEMON_H
#ifdef __cplusplus
extern "C" {
#endif
void be_a_daemon(void);
#ifdef __cplusplus
}
#endif
#endif
