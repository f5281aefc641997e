#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -k "chains or resident or staged or vaat or accept_local or debug or cpp" > $O/r02_p_pytest.log 2>&1; tail -5 $O/r02_p_pytest.log
for rep in 1 2; do
  (cd _r01 && timeout 300 python scripts/configs_bench.py c1 c3 2>/dev/null | grep '"mode": "per-chain"\|"config": "C1"' | cut -c1-260 | sed 's/^/r01: /')
  timeout 300 python scripts/configs_bench.py c1 c3 2>/dev/null | grep '"mode": "per-chain"\|"config": "C1"' | cut -c1-260 | sed 's/^/r02: /'
done > $O/r02_p_ab.txt 2>&1
cat $O/r02_p_ab.txt
