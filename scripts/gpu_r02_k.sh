#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -k "fake or small or stream or fullsize or graph or multi" > $O/r02_k_pytest.log 2>&1; tail -6 $O/r02_k_pytest.log
timeout 300 python scripts/configs_bench.py stream > $O/r02_k_stream.jsonl 2>&1; cut -c1-330 $O/r02_k_stream.jsonl
timeout 200 python scripts/sanitize_small.py > $O/r02_k_sanitize_plain.log 2>&1; tail -2 $O/r02_k_sanitize_plain.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > $O/r02_k_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 $O/r02_k_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_small.py > $O/r02_k_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 $O/r02_k_racecheck.log
