#!/bin/bash
# round 2, call A: at-size parity tests, compute-sanitizer (memcheck + racecheck) on smoke(), starting-point bench
mkdir -p gpurun_out
python -m pytest tests/test_gpu_full_size.py -x -q > gpurun_out/r02_fullsize.log 2>&1; echo "fullsize rc=$?" >> gpurun_out/r02_fullsize.log
tail -5 gpurun_out/r02_fullsize.log
timeout 700 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_memcheck_smoke.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_memcheck_stdout.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/r02_memcheck_smoke.log
timeout 700 compute-sanitizer --tool racecheck --log-file gpurun_out/r02_racecheck_smoke.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_racecheck_stdout.log 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/r02_racecheck_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_start.json 2> gpurun_out/r02_bench_start.err; echo "bench rc=$?"
cat gpurun_out/r02_bench_start.json | cut -c1-400
