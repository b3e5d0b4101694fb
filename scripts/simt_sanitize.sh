#!/bin/bash
# Memory and race checking of the plain-SIMT kernel SOURCES without a GPU: the host build of csrc/{geo,matcher,score,dense,pixel,backbone,
# evaluate,planes}.cu (tests/simt_host: one OS thread per CUDA thread, pthread barriers for __syncthreads / warp intrinsics) is
# compiled with -fsanitize=address or -fsanitize=thread and the host-execution tests are run under the sanitizer runtime.
#   address: out-of-bounds / use-after-free on global buffers, static and dynamic shared memory (report names the kernel line)
#   thread : data races between CUDA threads that no barrier / atomic orders (missing __syncthreads / __syncwarp)
# usage: scripts/simt_sanitize.sh [address|thread]   (default: both)      compute-sanitizer needs a GPU; this does not.
set -u
cd "$(dirname "$0")/.."
TESTS="tests/test_simt_host_kernels.py tests/test_simt_host_planes.py tests/test_host_reruns_gpu_tests.py"
rc=0
for san in ${1:-address thread}; do
  case $san in
    address) lib=$(gcc -print-file-name=libasan.so); export ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 ;;
    thread)  lib=$(gcc -print-file-name=libtsan.so); export TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0"
             export OMP_NUM_THREADS=1 ;;      # the OpenMP loops of the test-side stand-ins are not what is being checked
    *) echo "unknown sanitizer $san"; exit 2 ;;
  esac
  log=$(mktemp)
  NSAC_SIMT_SANITIZE=$san LD_PRELOAD=$lib python -m pytest $TESTS -q -p no:cacheprovider >"$log" 2>&1
  # reports are blocks between "==================" lines; only those with a frame in the generated kernel sources
  # (/tmp/nsac_simt_host/*.cpp) count — libtorch's own OpenMP threads produce TSan noise (uninstrumented libgomp)
  n=$(awk 'BEGIN{RS="=================="} /AddressSanitizer|ThreadSanitizer: data race/ && /nsac_simt_host/ {c++} END{print c+0}' "$log")
  other=$(awk 'BEGIN{RS="=================="} /ThreadSanitizer: data race/ && !/nsac_simt_host/ {c++} END{print c+0}' "$log")
  echo "[$san] $(grep -E '[0-9]+ (passed|failed)' "$log" | tail -1); reports in kernel sources: $n; elsewhere (libtorch / libgomp internals): $other"
  awk 'BEGIN{RS="=================="} /nsac_simt_host/ && /SUMMARY/ {print}' "$log" | grep "SUMMARY: " | sort | uniq -c | head -20
  grep -qE '[0-9]+ failed' "$log" && rc=1
  [ "$n" -ne 0 ] && rc=1
  rm -f "$log"
done
exit $rc
