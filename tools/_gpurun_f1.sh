python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|Error|assert |FAILED" | head -n 12
timeout 400 python bench.py > gpurun_out/f1_bench.json 2> gpurun_out/f1_bench.err; cut -c1-330 gpurun_out/f1_bench.json; tail -n 2 gpurun_out/f1_bench.err
