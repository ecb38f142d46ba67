# Development aid: one GPU carrying the slab a GPU gets under strong scaling of cfg5 (n_mu = 1024 / N), for several pass splits.
for cfg in "512 2" "256 4" "256 8" "128 2" "128 4" "128 8" "128 16"; do set -- $cfg
  echo -n "n_mu=$1 parts=$2: "; env PNJL_WS_PARTS=$2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --n-mu $1 2>&1 | grep -o '"value": [0-9.]*\|"converged": [0-9]*' | tr '\n' ' '; echo
done
