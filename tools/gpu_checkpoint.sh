#!/bin/bash
# One GPU-box call that produces everything a round's evidence needs (run under gpurun from the repo root):
#   the -m gpu suite in both pipelines and with both solvers of the split pipeline, the default bench line, the reference arm, the in-kernel phase
#   breakdown, the ncu launch list and one `ncu --set full` capture of one launch of each kernel.
#   tools/gpu_checkpoint.sh <tag>      -> gpurun_out/<tag>_*
# Afterwards, here:  python profiles/summarize.py gpurun_out/<tag>_full.ncu-rep <tag> --worlds 4096
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; tail -3 $out/${tag}_pytest_gpu.log
PX_PIPELINE=fused timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu_fused.log 2>&1; tail -3 $out/${tag}_pytest_gpu_fused.log
PX_SOLVER=dynamic timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu_dynamic.log 2>&1; tail -3 $out/${tag}_pytest_gpu_dynamic.log
python bench.py --steps 20 --warmup 3 > $out/${tag}_bench_line.json 2> $out/${tag}_bench.err; tail -c 300 $out/${tag}_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > $out/${tag}_bench_reference_line.json 2> $out/${tag}_bench_reference.err
python bench.py --steps 10 --warmup 3 --profile --no-cpu --no-e2e > $out/${tag}_phase_breakdown.json 2> $out/${tag}_phase.err
# launch list: skip the settle + warm-up launches.  Large batches are stepped as two halves on two streams: 2 x (5 sweep +
# 5 solver + 1 finish) = 22 launches per step of 5 substeps, 4 040 for the 1 000 settle substeps, 88 for warm-up: the timed
# region starts at launch 4 128 (a half's chain: sweep, solve, sweep, solve, ..., finish; launch 4 130 is a sweep that also
# finishes the previous substep).  Every launch then covers HALF the batch: summarize with --worlds 4096.
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $out/${tag}_plain.json 2> $out/${tag}_plain.err &&
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"px(Step|Solve|Replay|Finish)" -s 4130 -c 44 --csv \
      --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"px(Step|Solve|Replay|Finish)" -s 4130 -c 2 -o $out/${tag}_full -f \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $out/${tag}_ncu_full.log 2>&1
echo checkpoint done
