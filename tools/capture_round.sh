#!/bin/bash
# Round-end captures on the GPU box (run from the repo root under gpurun): GPU tests, smoke, both bench arms, the batch
# sweep and the ncu per-launch lists of one 1024-stream and one 64-stream step.  Everything lands in gpurun_out/;
# tools/summarize_step.py / summarize_dram.py / roofline_table.py turn the CSVs into the tables under profiles/.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_final.log 2>&1; echo "smoke rc=$?" >> $O/smoke_final.log
( time python bench.py > $O/bench_final.json 2> $O/bench_final.err ) 2> $O/bench_final.time
( time python bench.py --impl reference > $O/bench_ref_final.json 2> $O/bench_ref_final.err ) 2> $O/bench_ref_final.time
python tools/gpu_sweep.py --out $O/r02_sweep_300ms.json 1 64 256 512 1024 2048 4096 > $O/sweep_final.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $O/r02_step_B1024.csv python tools/gpu_one.py 1024 steps=1 > $O/ncu_step1024.log 2>&1
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $O/r02_step_B64.csv python tools/gpu_one.py 64 steps=1 > $O/ncu_step64.log 2>&1
tail -2 $O/pytest_final.log; tail -1 $O/smoke_final.log; cat $O/bench_final.time | grep real; head -c 400 $O/bench_final.json
