#!/bin/bash
# BASELINE.json configs 2-5 as stated, one bench.py line each (run under gpurun from the repo root):
#   tools/gpu_configs.sh <tag>   -> gpurun_out/<tag>_config{2,3,4,5_k1,5_k8}.json
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
set -x
python bench.py --config 2 --steps 10 --warmup 3 --no-cpu --no-aos > $out/${tag}_config2.json 2> $out/${tag}_config2.err
python bench.py --config 3 --steps 10 --warmup 3 --no-cpu --no-aos > $out/${tag}_config3.json 2> $out/${tag}_config3.err
python bench.py --config 4 --steps 10 --warmup 3 --no-cpu --no-aos > $out/${tag}_config4.json 2> $out/${tag}_config4.err
python bench.py --config 5 --fused 1 --steps 10 --warmup 3 --no-cpu --no-aos > $out/${tag}_config5_k1.json 2> $out/${tag}_config5_k1.err
python bench.py --config 5 --fused 8 --steps 10 --warmup 3 --no-cpu --no-aos > $out/${tag}_config5_k8.json 2> $out/${tag}_config5_k8.err
tail -c 400 $out/${tag}_config*.err
echo configs done
