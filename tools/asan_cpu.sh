#!/bin/bash
# AddressSanitizer + UBSan over everything of the boundary that runs on the CPU: the oracle, the reference's own sources
# compiled in place, the drop-in shim bodies in both stand-in worlds (ABI answered by the oracle) and the product's
# __host__ __device__ arithmetic twin (tests/hostcheck.cpp). Builds an instrumented copy of HEAD under /tmp/asan_tree (the
# working tree is not touched) and runs the CPU test files that load those libraries under LD_PRELOAD=libasan.
# Findings would land in /tmp/asan_log.* / /tmp/ubsan_log.*; the summary of the last run is profiles/r02_sanitizer.txt.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
T=/tmp/asan_tree
SAN="-fsanitize=address,undefined -fno-omit-frame-pointer"
rm -rf $T /tmp/asan_log.* /tmp/ubsan_log.* && mkdir $T && (cd $ROOT && git archive HEAD) | tar -x -C $T
cp $ROOT/orb_slam3_fast_b200/liborbx.so $T/orb_slam3_fast_b200/   # link-time dependency of tests/Makefile only
make -C $T/oracle -s CXX=/usr/bin/g++ CXXFLAGS="-O1 -g -std=c++17 -fPIC -Wall -ffp-contract=off -fno-fast-math -pthread $SAN" all
make -C $T/tests -s
/usr/bin/g++ -O1 -g -std=c++17 -fPIC -shared -ffp-contract=off -Wno-unknown-pragmas $SAN -o $T/tests/libhostcheck.so $T/tests/hostcheck.cpp
cd $T
ASAN_OPTIONS=detect_leaks=0:halt_on_error=0:log_path=/tmp/asan_log UBSAN_OPTIONS=print_stacktrace=1:log_path=/tmp/ubsan_log \
LD_PRELOAD=$(/usr/bin/g++ -print-file-name=libasan.so) python -m pytest -q -m "not gpu" -p no:cacheprovider \
  tests/test_shim_bodies_vs_reference_source.py tests/test_shim_bow_vs_reference_source.py \
  tests/test_oracle_matchers_vs_reference_source.py tests/test_oracle_vs_reference_source.py tests/test_oracle_matchers.py \
  tests/test_oracle_pipeline.py tests/test_host_math.py
ls /tmp/asan_log.* /tmp/ubsan_log.* 2>/dev/null && echo "FINDINGS (see the files above)" || echo "no sanitizer findings"
