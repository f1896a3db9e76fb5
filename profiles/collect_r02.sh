#!/bin/bash
# ncu evidence of round 2 (run under gpurun, outputs to gpurun_out/; profiles/postprocess_r02.py turns them into the committed
# summaries): launch list of the bench command, --set full captures of every kernel of the step at C4 / C2 / C3 and of the
# generic-path kernels, then plain bench lines of every configuration and the reference arm.
R=${1:-r02}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_$R.log 2>&1
$NCU --set full --import-source on -k 'regex:k_seg_decide|k_cell_scatter|k_cell_init|k_scan_tiles|k_scan_apply|k_step_end' -s 12 -c 6 -f -o gpurun_out/prof_${R}_c4 \
    python bench.py --config C4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_full_${R}_c4.log 2>&1
$NCU --set full --import-source on -k 'regex:k_seg_decide|k_cell_scatter' -s 4 -c 2 -f -o gpurun_out/prof_${R}_c2 \
    python bench.py --config C2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_full_${R}_c2.log 2>&1
$NCU --set full --import-source on -k 'regex:k_seg_decide|k_cell_scatter|k_make_offspring|k_free_genomes' -s 8 -c 4 -f -o gpurun_out/prof_${R}_c3 \
    python bench.py --config C3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_full_${R}_c3.log 2>&1
QHG_B200_PATH=generic QHG_GEN_FAST=0 $NCU --set full --import-source on -k 'regex:k_actions|k_scatter|k_pair|k_make_offspring|k_free_genomes' -s 8 -c 8 -f -o gpurun_out/prof_${R}_generic \
    python profiles/prof_generic.py > gpurun_out/prof_generic_$R.log 2>&1
# summaries on the box (the reports themselves are 15-25 MB each and gpurun brings back 64 MB at most: only C4's travels)
for x in c4 c2 c3 generic; do
  python profiles/ncu_summary.py gpurun_out/prof_${R}_$x.ncu-rep > gpurun_out/ncu_full_${R}_${x}_summary.txt 2>&1
  ncu -i gpurun_out/prof_${R}_$x.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/ncu_src_${R}_$x.csv 2>/dev/null
  python profiles/ncu_lines.py gpurun_out/ncu_src_${R}_$x.csv 40 > gpurun_out/ncu_lines_${R}_$x.txt 2>&1
  ncu -i gpurun_out/prof_${R}_$x.ncu-rep --page raw --csv > gpurun_out/ncu_raw_${R}_$x.csv 2>/dev/null
  [ $x == c4 ] || rm -f gpurun_out/prof_${R}_$x.ncu-rep gpurun_out/ncu_src_${R}_$x.csv
done
du -sh gpurun_out
for c in C4 C2 C3 C5; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 $( [ $c == C4 ] || echo --no-cpu-baseline ) > gpurun_out/bench_${c}_$R.json 2> gpurun_out/bench_${c}_$R.err
  tail -c 300 gpurun_out/bench_${c}_$R.json; echo
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$R.json 2>&1; tail -c 400 gpurun_out/bench_ref_$R.json
ls -la gpurun_out | head -40
