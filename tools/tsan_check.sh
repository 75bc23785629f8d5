#!/bin/bash
# ThreadSanitizer over the host pipeline (query lanes, worker threads, the seed-turn hand-over) with the CPU oracle underneath:
# builds into /tmp/tsan and runs the CLI with several lanes on the files given.   tools/tsan_check.sh blastp -q Q -d D [options]
set -e
cd "$(dirname "$0")/.."
H=diamond_b200/csrc/host; O=/tmp/tsan
mkdir -p $O
FL="-O1 -g -std=c++17 -ffp-contract=off -fPIC -pthread -fsanitize=thread"
for f in pipeline chaining scoring; do g++ $FL -c $H/$f.cpp -o $O/$f.o; done
gcc -O1 -g -ffp-contract=off -fPIC -fsanitize=thread -c oracle/dmnd_oracle.c -o $O/oracle.o
g++ -shared -pthread -fsanitize=thread -o $O/libdmnd_oracle.so $O/pipeline.o $O/chaining.o $O/scoring.o $O/oracle.o -lm
g++ $FL $H/cli.cpp -o $O/dmnd-tsan-cli -L$O -ldmnd_oracle -lz -Wl,-rpath,$O
DMND_LANES=${DMND_LANES:-4} $O/dmnd-tsan-cli "$@" -o $O/out -p 8 2>&1 | grep -E "WARNING|SUMMARY" | sort | uniq -c
echo "tsan run finished (no lines above = no report)"
