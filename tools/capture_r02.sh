#!/usr/bin/env bash
# One-GPU capture of everything profiles/r02_* is built from, in two calls because gpurun brings back at most 64 MiB:
#   gpurun -- 'bash tools/capture_r02.sh a'   tests, bench lines, launch list, ncu --set full of the feature kernel
#   gpurun -- 'bash tools/capture_r02.sh b'   ncu --set full of the retrieval kernel, secondary benchmarks
# Every profiler pass runs only after the plain command exited 0; nothing printed under ncu is a bench value.
set -u
OUT=gpurun_out
mkdir -p $OUT
PART=${1:-a}
BENCH_ARGS="--steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --stft-steps 2 --retrieval-steps 1"
if [ "$PART" = a ]; then
python -m pytest tests -m gpu -x -q > $OUT/r02f_gputest.log 2>&1; echo "gpu tests rc=$?"; tail -2 $OUT/r02f_gputest.log
python bench.py > $OUT/r02f_bench_1gpu.log 2> $OUT/r02f_bench_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference > $OUT/r02f_bench_reference.log 2> $OUT/r02f_bench_reference.err; echo "reference arm rc=$?"
python bench.py $BENCH_ARGS > $OUT/r02f_bench_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r02f_launches_raw.csv python bench.py $BENCH_ARGS > $OUT/r02f_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:feat_warp8_kernel -s 3 -c 1 -o $OUT/r02f_feat_warp8 -f python bench.py $BENCH_ARGS > $OUT/r02f_ncu_feat.log 2>&1
else
python bench.py $BENCH_ARGS > $OUT/r02f_bench_plain_b.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:cosine_topk_tc_kernel -c 1 -o $OUT/r02f_topk_tc -f python bench.py $BENCH_ARGS > $OUT/r02f_ncu_topk.log 2>&1
python benchmarks/bench_extra.py --sweep --logmel128 --clips 1000 > $OUT/r02f_bench_extra.jsonl 2>&1
python benchmarks/odd_configs_bench.py > $OUT/r02f_odd_configs.jsonl 2>&1
python benchmarks/nchw_bench.py > $OUT/r02f_nchw.jsonl 2>&1
python benchmarks/retr_quick.py > $OUT/r02f_retr_quick.txt 2>&1
fi
ls -la $OUT | grep r02f
