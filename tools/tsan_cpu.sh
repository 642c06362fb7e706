#!/bin/bash
# ThreadSanitizer over the drop-in bodies used from several threads at once, on the CPU: an instrumented copy of HEAD is
# built under /tmp/tsan_tree (oracle, the reference's sources compiled in place, both stand-in shim worlds with the ABI
# answered by the oracle) and the three-thread tests run under LD_PRELOAD=libtsan. What it covers: the per-thread matcher
# context (shim/orbx_thread_matcher.h), the per-thread vocabulary record of shim/FrameBoW_orbx.cc, and that the bodies
# keep no other shared state. (The CUDA side of the threading contract is tests/thread_check.cpp, on the GPU.)
# Reports would land in /tmp/tsan_log.*; the summary of the last run is in profiles/r02_sanitizer.txt.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
T=/tmp/tsan_tree
rm -rf $T /tmp/tsan_log.* && mkdir $T && (cd $ROOT && git archive HEAD) | tar -x -C $T
cp $ROOT/orb_slam3_fast_b200/liborbx.so $T/orb_slam3_fast_b200/   # link-time dependency only
make -C $T/oracle -s CXX=/usr/bin/g++ CXXFLAGS="-O1 -g -std=c++17 -fPIC -Wall -ffp-contract=off -fno-fast-math -pthread -fsanitize=thread -fno-omit-frame-pointer" all
cd $T
TSAN_OPTIONS="halt_on_error=0:log_path=/tmp/tsan_log:report_signal_unsafe=0" LD_PRELOAD=$(/usr/bin/g++ -print-file-name=libtsan.so) \
  python -m pytest -q -p no:cacheprovider -k three_threads tests/test_shim_bow_vs_reference_source.py tests/test_shim_bodies_vs_reference_source.py
ls /tmp/tsan_log.* 2>/dev/null && echo "REPORTS (see the files above)" || echo "no ThreadSanitizer reports"
