#!/bin/bash
# N parse-only HM decoder processes on one stream at once, packed syntax to /dev/null: the host producers' own rate (no GPU consumer).
# usage: producers_alone.sh <N> <stream.bin> <pictures per stream>
N=$1; S=$2; P=$3
t0=$(date +%s.%N)
for i in $(seq 1 $N); do
  ( exec 9>/dev/null; HM_PACKED_FD=9 CDS_PARSE_ONLY=1 "$(dirname "$0")/../oracle/_ref/hm_syntax_dump" -b "$S" >/dev/null 2>&1 ) &
done
wait
t1=$(date +%s.%N)
python3 - <<PY
n, p, dt = $N, $P, $t1 - $t0
print('{"producers": %d, "pictures": %d, "seconds": %.3f, "pictures_per_s": %.1f, "ms_per_picture_and_process": %.2f}' % (n, n * p, dt, n * p / dt, 1e3 * dt / p))
PY
