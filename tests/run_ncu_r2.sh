#!/bin/bash
# Round-2 ncu captures on one B200 (gpurun): launch lists of one step per codec and `--set full`
# captures of every dominant kernel.  usage: bash tests/run_ncu_r2.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# launch lists (serialised, cold cache: shares only)
for c in hsq qsgd terngrad sign topk; do
  $NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file gpurun_out/launches_${c}_$TAG.csv \
      python bench.py --codec $c --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "launches $c rc=$?"
done
full() {  # name, kernel regex, skip, count, command...
  n=$1; k=$2; s=$3; c=$4; shift 4
  $NCU --set full --import-source on -k regex:$k -s $s -c $c -f -o gpurun_out/prof_${n}_$TAG "$@" > gpurun_out/ncu_${n}_$TAG.log 2>&1
  echo "full $n rc=$?"
  # the reports are too large to travel (gpurun_out is capped at 64 MiB): keep the raw-metric page and the
  # per-instruction hot-spot summary, drop the report
  ncu -i gpurun_out/prof_${n}_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${n}_${TAG}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_${n}_$TAG.ncu-rep --page source --csv --print-source sass > /tmp/src_$n.csv 2>/dev/null
  python profiles/sass_hot.py /tmp/src_$n.csv > gpurun_out/prof_${n}_${TAG}_hot.txt 2>&1
  [ "$n" = "enc" ] || rm -f gpurun_out/prof_${n}_$TAG.ncu-rep
}
full enc  hsq_encode_tc2            4 1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
full dec1 hsq_decode_reduce_staged  4 1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
full dec8 hsq_decode_reduce_staged  3 1 python tests/dec_time.py 8
full qsgd_enc  qsgd_encode_chunks_kernel.*1,32,4   3 1 python bench.py --codec qsgd --steps 3 --warmup 3 --no-cpu-baseline
full qsgd_dec  qsgd_decode_reduce8  6 1 python bench.py --codec qsgd --steps 3 --warmup 3 --no-cpu-baseline
full tern_max  seg_absmax_ranges    3 1 python bench.py --codec terngrad --steps 3 --warmup 3 --no-cpu-baseline
full tern_q    qsgd_quantize_ranges 3 1 python bench.py --codec terngrad --steps 3 --warmup 3 --no-cpu-baseline
full sign_enc  sign_encode          3 1 python bench.py --codec sign --steps 3 --warmup 3 --no-cpu-baseline
full sign_dec  sign_decode_reduce   3 1 python bench.py --codec sign --steps 3 --warmup 3 --no-cpu-baseline
full topk_hist topk_hist0           3 1 python bench.py --codec topk --steps 3 --warmup 3 --no-cpu-baseline
full topk_cls  topk_classify        3 1 python bench.py --codec topk --steps 3 --warmup 3 --no-cpu-baseline
full enc_d8  hsq_encode_tc2         2 1 python bench.py --codec hsq --c-dim 8 --workload flat --steps 3 --warmup 3 --no-cpu-baseline
full enc_d32 hsq_encode_tc2         2 1 python bench.py --codec hsq --c-dim 32 --workload flat --steps 3 --warmup 3 --no-cpu-baseline
full tck     hsq_search_tck         2 1 python bench.py --codec hsq --k-bit 12 --workload flat --steps 3 --warmup 3 --no-cpu-baseline
ls -la gpurun_out/prof_*_$TAG* | awk '{print $5, $9}'
