#!/bin/bash
# A/B runs of the Monte-Carlo tile kernel: register-limited builds and CTA sizes
run() { # lib tpb n
  local out
  out=$(CARDOON_B200_LIB=$1 python bench.py --workload mc --steps 5 --warmup 3 --tile-tpb $2 --n $3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['config'].get('failed_instances'))")
  echo "lib=$(basename "$1") tpb=$2 n=$3 -> $out"
}
V=cardoon_b200/variants
run "" 0 4096
run $V/lib_t384.so 384 4096
run $V/lib_t384.so 320 4096
run $V/lib_t512.so 512 4096
run $V/lib_t512.so 448 4096
run $V/lib_t192x2.so 192 4096
run $V/lib_t192x2.so 128 4096
run "" 0 512
run "" 32 512
run "" 64 512
run "" 128 512
run "" 32 1024
run "" 64 1024
run "" 64 2048
run "" 128 2048
