#!/bin/bash
# One-GPU evidence run (on the box): smoke, GPU tests, both bench arms, launch list, ncu --set full captures of the
# bench kernels, drop-in timing and the range sweeps.  Everything lands in gpurun_out/ under the given prefix (default r2).
P=${1:-r2}
O=gpurun_out
mkdir -p $O
set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${P}_smoke.log 2>&1; tail -1 $O/${P}_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/pytest_gpu.log; cat $O/pytest_gpu.log
grep -q passed $O/pytest_gpu.log || exit 1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/${P}_bench_n1_reference.json 2> $O/${P}_bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 3 > $O/${P}_bench_n1.json 2> $O/${P}_bench.err; tail -c 300 $O/${P}_bench.err
timeout 600 python tools/dropin_time.py > $O/${P}_dropin_c1.json 2>/dev/null
timeout 600 python tools/sweep.py --metric sad > $O/${P}_sweep_sad.json 2>/dev/null
timeout 600 python tools/sweep.py --metric wrap8 > $O/${P}_sweep_wrap8.json 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${P}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --skip-c3 --skip-sad --skip-cpu > $O/${P}_ncu_launches.log 2>&1
FULL="--set full --clock-control none --import-source on -f"
timeout 600 ncu $FULL -k regex:me_tiled -c 1 -o $O/${P}_me python bench.py --steps 1 --warmup 1 --skip-c3 --skip-sad --skip-cpu > /dev/null 2>&1
R=16 T=24 timeout 600 ncu $FULL -k regex:me_tiled -c 1 -o $O/${P}_me_sad python tools/prof_me.py > /dev/null 2>&1
timeout 600 ncu $FULL -k regex:dct_stage -s 1 -c 1 -o $O/${P}_dct_fi python tools/prof_dct.py > /dev/null 2>&1
timeout 600 ncu $FULL --kernel-name-base demangled -k 'regex:dct_stage_kernel<\(int\)3, \(int\)0>' -s 12 -c 1 -o $O/${P}_dct_fwd python tools/e2e_sched.py "" > /dev/null 2>&1
timeout 600 ncu $FULL -k regex:pack_write -s 12 -c 1 -o $O/${P}_pack_write python tools/e2e_sched.py "" > /dev/null 2>&1
ls -la $O/*.ncu-rep
