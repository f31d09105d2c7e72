#!/bin/bash
# Run the host-side CPU tests against sanitizer builds of the executables and the ring shim.
# Works on a scratch copy (the in-tree build is left alone).  No GPU needed.
#   pass 1: AddressSanitizer + UndefinedBehaviorSanitizer on everything under host/
#   pass 2: ThreadSanitizer on paf_capture / paf_capture_stock (the one multi-threaded program;
#           bmf_replay and paf_memdb use OpenMP, whose runtime TSan cannot see into)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
S=${1:-/tmp/b2p_sanitize}
rm -rf "$S" && mkdir -p "$S" && cp -r "$ROOT"/. "$S"/ && rm -rf "$S/.git" "$S/gpurun_out"
cd "$S"
H=paf_baseband2power_b200/host
TESTS="tests/test_capture.py tests/test_host_ring.py tests/test_reference_producers.py tests/test_launcher.py tests/test_reference_main.py"

make -s -C $H clean >/dev/null
make -s -C $H all HOSTCC="gcc -fsanitize=address,undefined -fno-omit-frame-pointer" CFLAGS="-O1 -g -Wall -Wextra -std=gnu11 -fPIC"
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1:abort_on_error=1 UBSAN_OPTIONS=halt_on_error=1:print_stacktrace=1 \
LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) \
  python -m pytest $TESTS -x -q -m "not gpu" -p no:cacheprovider

make -s -C $H clean >/dev/null
make -s -C $H all
( cd $H
  SH="dada/ipcbuf.c dada/ipcio.c dada/dada_hdu.c dada/ascii_header.c dada/multilog.c dada/futils.c"
  gcc -fsanitize=thread -O1 -g -std=gnu11 -o ../bin/paf_capture paf_capture.c $SH -lpthread -lm
  gcc -fsanitize=thread -O1 -g -std=gnu11 -DB2P_STOCK_PSRDADA -DB2P_NO_SHIM_EXTENSIONS -o ../bin/paf_capture_stock paf_capture.c $SH -lpthread -lm )
rm -f "$S"/tsan.*
TSAN_OPTIONS="halt_on_error=0 exitcode=66 log_path=$S/tsan" \
  python -m pytest tests/test_capture.py -x -q -m "not gpu" -p no:cacheprovider -k "not header_decode and not stock_capture_uses"
n=$(ls "$S"/tsan.* 2>/dev/null | wc -l)
echo "ThreadSanitizer reports: $n"
[ "$n" -eq 0 ]
