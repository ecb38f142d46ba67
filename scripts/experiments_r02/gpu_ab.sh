#!/bin/bash
# same-box A/B of two builds (ab_build/libA.so = HEAD, libB.so = general integrand path as of 572af50)
for r in 1 2; do for L in A B; do
  echo "== lib$L WS N=1"; PNJL_LIB=$PWD/ab_build/lib$L.so timeout 300 python scripts/dev_bench.py --workload cfg5 --schedule 2 2>&1 | tail -1 | sed 's/ | passes.*//'
done; done
for L in A B; do
  echo "== lib$L march N=1"; PNJL_LIB=$PWD/ab_build/lib$L.so timeout 300 python scripts/dev_bench.py --workload cfg5 --schedule 3 2>&1 | tail -1 | sed 's/ | passes.*//'
  echo "== lib$L cfg3 WS"; PNJL_LIB=$PWD/ab_build/lib$L.so timeout 300 python scripts/dev_bench.py --workload cfg3 --schedule 2 --n-t 60000 2>&1 | tail -1 | sed 's/ | passes.*//'
done
