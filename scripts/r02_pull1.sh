#!/bin/bash
# first GPU run of the pull ray cast: parity suite, memcheck of the smoke run, bench A/B against the per-beam walk
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pull1_pytest.txt 2>&1
tail -15 gpurun_out/r02_pull1_pytest.txt
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r02_pull1_memcheck.txt 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/r02_pull1_memcheck.txt
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/r02_pull1_bench.json 2> gpurun_out/r02_pull1_bench.err
echo "bench rc=$?"; cat gpurun_out/r02_pull1_bench.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['e2e']['value'], d.get('kernel_ms_per_step'), d['counters_per_step'])"
LAMA_NO_PULL=1 timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/r02_pull1_bench_nopull.json 2> gpurun_out/r02_pull1_bench_nopull.err
cat gpurun_out/r02_pull1_bench_nopull.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['e2e']['value'], d.get('kernel_ms_per_step'))"
