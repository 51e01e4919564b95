#!/bin/bash
# A/B runs of the soliton block kernel: CTA sizes
run() {
  local out
  out=$(CARDOON_B200_LIB=$1 python bench.py --workload soliton --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])")
  echo "lib=$(basename "$1") -> $out"
}
V=cardoon_b200/variants
run ""
for t in "$@"; do run $V/lib_$t.so; done
