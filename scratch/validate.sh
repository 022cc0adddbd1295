#!/bin/bash
# one validation pass of the restored tree: GPU parity tests, smoke, both bench arms
s=$(date +%s)
timeout 420 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16
echo "pytest took $(( $(date +%s) - s )) s"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
s=$(date +%s)
python bench.py > gpurun_out/bench_validate.json 2> gpurun_out/bench_validate.err; echo "bench rc=$? $(( $(date +%s) - s )) s"
s=$(date +%s)
python bench.py --impl reference > gpurun_out/bench_validate_ref.json 2> gpurun_out/bench_validate_ref.err; echo "ref rc=$? $(( $(date +%s) - s )) s"
cat gpurun_out/bench_validate.json | cut -c1-600
cat gpurun_out/bench_validate_ref.json | cut -c1-600
