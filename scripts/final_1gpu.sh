#!/bin/bash
# Round-end validation on one B200: GPU tests, every bench configuration, the reference arm, an ncu launch list, smoke().
# Usage (from the repo root): gpurun --timeout 1500 -- 'bash scripts/final_1gpu.sh r3d'
out=gpurun_out/${1:-final}
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $out/tests.log
cat $out/tests.log
python bench.py > $out/bench_train_bf16.json 2> $out/bench.err
python bench.py --no-cpu-baseline --no-gpu-baseline > $out/bench_train_bf16_b.json 2>> $out/bench.err
for c in fwd_fp32 attn_bf16 clap_infer; do
  timeout 300 python bench.py --config $c --no-cpu-baseline > $out/bench_$c.json 2>> $out/bench.err
done
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2>> $out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 480 --csv \
  --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --settle 0 > $out/ncu_bench.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ls -la $out
for f in $out/bench_*.json; do python profiles/show_bench.py $f 2>/dev/null | head -1; done
