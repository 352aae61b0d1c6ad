#!/usr/bin/env bash
# One-GPU A/B of the retrieval filter's mixed TF32 + fp16 split (last session of round 2):
#   gpurun -- 'bash tools/capture_r02g.sh ab'      error micro-benchmark, retrieval GPU tests, retr_quick on three libraries
#   gpurun -- 'bash tools/capture_r02g.sh final'   whole GPU suite, bench lines, launch list, ncu of the retrieval kernel
#   gpurun -- 'bash tools/capture_r02g.sh final_b' ncu of the feature kernel, frame x hop sweep
# The A/B libraries live in gpurun_scratch/ (git-ignored, travels with the snapshot): libdspx_head.so = the previous
# commit, libdspx_mixed_4vote.so = mixed split with the per-block votes, libdspx_prof.so = -DDSPX_TC_PROFILE.
set -u
OUT=gpurun_out
mkdir -p $OUT
PART=${1:-ab}
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/cuda_runtime/lib:${LD_LIBRARY_PATH:-}
BENCH_ARGS="--steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --stft-steps 2 --retrieval-steps 1"
if [ "$PART" = ab ]; then
timeout 60 benchmarks/micro/umma_f16x > $OUT/r02g_umma_f16x.txt 2>&1; echo "umma_f16x rc=$?"; grep -c "max err" $OUT/r02g_umma_f16x.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "retrieval or topk or sharded or pipeline or embed" > $OUT/r02g_gputest_retr.log 2>&1; echo "retrieval gpu tests rc=$?"; tail -2 $OUT/r02g_gputest_retr.log
timeout 300 python benchmarks/retr_quick.py > $OUT/r02g_retr_quick_new.txt 2>&1; echo "quick new rc=$?"; cat $OUT/r02g_retr_quick_new.txt
for lib in head mixed_4vote; do
  if [ -f gpurun_scratch/libdspx_$lib.so ]; then
    DSPX_LIBRARY=$PWD/gpurun_scratch/libdspx_$lib.so timeout 300 python benchmarks/retr_quick.py > $OUT/r02g_retr_quick_$lib.txt 2>&1; echo "quick $lib rc=$?"; cat $OUT/r02g_retr_quick_$lib.txt
  fi
done
for sp in 5 9 11 13; do DSPX_TOPK_SPLITS=$sp timeout 120 python benchmarks/retr_time.py 2>&1 | tail -1 | sed "s/^/splits $sp: /" >> $OUT/r02g_retr_splits.txt; done; cat $OUT/r02g_retr_splits.txt
if [ -f gpurun_scratch/libdspx_prof.so ]; then
  DSPX_LIBRARY=$PWD/gpurun_scratch/libdspx_prof.so timeout 300 python benchmarks/retr_tc_roles_steady.py > $OUT/r02g_roles_steady.txt 2>&1; echo "roles rc=$?"; cat $OUT/r02g_roles_steady.txt
fi
elif [ "$PART" = f16 ]; then
# all-fp16 split + bulk-copy operands: error micro-benchmark, retrieval tests, timings, steady-state roles
timeout 60 benchmarks/micro/umma_f16x > $OUT/r02h_umma_f16x.txt 2>&1; echo "umma_f16x rc=$?"; grep "status\|max err" $OUT/r02h_umma_f16x.txt | sort | uniq -c | sort -rn | head -30
timeout 600 python -m pytest tests -m gpu -x -q -k "retrieval or topk or sharded or pipeline or embed" > $OUT/r02h_gputest_retr.log 2>&1; echo "retrieval gpu tests rc=$?"; tail -5 $OUT/r02h_gputest_retr.log
timeout 300 python benchmarks/retr_quick.py > $OUT/r02h_retr_quick.txt 2>&1; echo "quick rc=$?"; cat $OUT/r02h_retr_quick.txt
timeout 120 python benchmarks/retr_time.py 2>&1 | tail -1 > $OUT/r02h_retr_time.txt; cat $OUT/r02h_retr_time.txt
for sp in 5 9 11 13; do DSPX_TOPK_SPLITS=$sp timeout 120 python benchmarks/retr_time.py 2>&1 | tail -1 | sed "s/^/splits $sp: /" >> $OUT/r02h_retr_splits.txt; done; cat $OUT/r02h_retr_splits.txt
for f in gpurun_scratch/libdspx_prof*.so; do
  [ -f $f ] || continue
  echo "## $f (DSPX_EXPERIMENT_KEEP_THR=1: calls 1, 2 are steady state)" >> $OUT/r02h_roles_keepthr.txt
  DSPX_EXPERIMENT_KEEP_THR=1 DSPX_LIBRARY=$PWD/$f timeout 300 python benchmarks/retr_tc_roles_steady.py >> $OUT/r02h_roles_keepthr.txt 2>&1
done
cat $OUT/r02h_roles_keepthr.txt
elif [ "$PART" = seed ]; then
# threshold seed: retrieval tests, size sweep (DSPX_TOPK_SEED_TILES, 0 = off), three shapes with and without
timeout 600 python -m pytest tests -m gpu -x -q -k "retrieval or topk or sharded or pipeline or embed" > $OUT/r02i_gputest_retr.log 2>&1; echo "retrieval gpu tests rc=$?"; tail -3 $OUT/r02i_gputest_retr.log
for sd in 0 16 32 64 128 256; do DSPX_TOPK_SEED_TILES=$sd timeout 120 python benchmarks/retr_time.py 2>&1 | tail -1 | sed "s/^/seed tiles $sd: /" >> $OUT/r02i_seed_sweep.txt; done; cat $OUT/r02i_seed_sweep.txt
echo "## default (64 seed tiles)" > $OUT/r02i_retr_quick.txt; timeout 300 python benchmarks/retr_quick.py >> $OUT/r02i_retr_quick.txt 2>&1
echo "## DSPX_TOPK_SEED_TILES=0" >> $OUT/r02i_retr_quick.txt; DSPX_TOPK_SEED_TILES=0 timeout 300 python benchmarks/retr_quick.py >> $OUT/r02i_retr_quick.txt 2>&1
echo "## DSPX_TOPK_SEED_TILES=256" >> $OUT/r02i_retr_quick.txt; DSPX_TOPK_SEED_TILES=256 timeout 300 python benchmarks/retr_quick.py >> $OUT/r02i_retr_quick.txt 2>&1
cat $OUT/r02i_retr_quick.txt
for f in gpurun_scratch/libdspx_prof*.so; do
  [ -f $f ] || continue
  echo "## $f, cold calls (first-wave CTA (0, 0) after the seed launch)" >> $OUT/r02i_roles.txt
  DSPX_LIBRARY=$PWD/$f timeout 300 python benchmarks/retr_tc_roles_steady.py 2>&1 | grep "^call" >> $OUT/r02i_roles.txt
done
cat $OUT/r02i_roles.txt
elif [ "$PART" = var ]; then
# tuning variants: every gpurun_scratch/libdspx_<name>.so except the profiling build; then the steady-state role profile
: > $OUT/r02g_variants.txt
echo "## default library" >> $OUT/r02g_variants.txt
timeout 300 python benchmarks/retr_quick.py >> $OUT/r02g_variants.txt 2>&1
for f in gpurun_scratch/libdspx_*.so; do
  name=$(basename $f .so); name=${name#libdspx_}
  case $name in prof*) continue;; esac
  echo "## $name" >> $OUT/r02g_variants.txt
  DSPX_LIBRARY=$PWD/$f timeout 300 python benchmarks/retr_quick.py >> $OUT/r02g_variants.txt 2>&1
done
cat $OUT/r02g_variants.txt
for f in gpurun_scratch/libdspx_prof*.so; do
  [ -f $f ] || continue
  echo "## $f (DSPX_EXPERIMENT_KEEP_THR=1: calls 1, 2 are steady state)" >> $OUT/r02g_roles_keepthr.txt
  DSPX_EXPERIMENT_KEEP_THR=1 DSPX_LIBRARY=$PWD/$f timeout 300 python benchmarks/retr_tc_roles_steady.py >> $OUT/r02g_roles_keepthr.txt 2>&1
done
cat $OUT/r02g_roles_keepthr.txt
elif [ "$PART" = final ]; then
# gpurun brings back at most 64 MiB: one ncu --set full report per call (final = retrieval kernel, final_b = feature kernel)
python -m pytest tests -m gpu -x -q > $OUT/r02g_gputest.log 2>&1; echo "gpu tests rc=$?"; tail -2 $OUT/r02g_gputest.log
python bench.py > $OUT/r02g_bench_1gpu.log 2> $OUT/r02g_bench_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference > $OUT/r02g_bench_reference.log 2> $OUT/r02g_bench_reference.err; echo "reference arm rc=$?"
python bench.py $BENCH_ARGS > $OUT/r02g_bench_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r02g_launches_raw.csv python bench.py $BENCH_ARGS > $OUT/r02g_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cosine_topk_tc_kernel -c 1 -o $OUT/r02g_topk_tc -f python bench.py $BENCH_ARGS > $OUT/r02g_ncu_topk.log 2>&1
python benchmarks/retr_quick.py > $OUT/r02g_retr_quick_final.txt 2>&1
else
python bench.py $BENCH_ARGS > $OUT/r02g_bench_plain_b.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:feat_warp8_kernel -s 3 -c 1 -o $OUT/r02g_feat_warp8 -f python bench.py $BENCH_ARGS > $OUT/r02g_ncu_feat.log 2>&1
python benchmarks/bench_extra.py --sweep --logmel128 --clips 1000 > $OUT/r02g_bench_extra.jsonl 2>&1
fi
ls -la $OUT | grep 'r02[ghi]'
