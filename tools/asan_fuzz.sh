#!/bin/bash
# Builds the host code (pipeline, chaining, scoring, CLI) and the CPU oracle with AddressSanitizer + UndefinedBehaviorSanitizer into /tmp/asan and
# runs tools/diff_fuzz.py with that CLI: every run must still be byte-identical to the reference, and a sanitizer report aborts the run (rc != 0
# shows up as a DIFF line).  CPU only -- the device kernels are covered by compute-sanitizer on the GPU box (profiles/memcheck_r1.txt).
#   tools/asan_fuzz.sh [diff_fuzz.py options]      e.g.  tools/asan_fuzz.sh --runs 60 --seed 1401 --filters 0.4 --blocks 0.3
set -e
cd "$(dirname "$0")/.."
H=diamond_b200/csrc/host
O=/tmp/asan
mkdir -p $O
FL="-O1 -g -std=c++17 -ffp-contract=off -fPIC -pthread -fsanitize=address,undefined -fno-omit-frame-pointer"
for f in pipeline chaining scoring; do g++ $FL -c $H/$f.cpp -o $O/$f.o; done
gcc -O1 -g -ffp-contract=off -fPIC -fsanitize=address,undefined -c oracle/dmnd_oracle.c -o $O/oracle.o
g++ -shared -pthread -fsanitize=address,undefined -o $O/libdmnd_oracle.so $O/pipeline.o $O/chaining.o $O/scoring.o $O/oracle.o -lm
g++ $FL $H/cli.cpp -o $O/dmnd-asan-cli -L$O -ldmnd_oracle -lz -Wl,-rpath,$O
export ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1
exec python tools/diff_fuzz.py --cli $O/dmnd-asan-cli "$@"
