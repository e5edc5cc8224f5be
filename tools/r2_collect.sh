#!/bin/bash
# After a GPU call: turn what came back in gpurun_out/ into the tracked evidence under profiles/ (run here, on the CPU box).
set -u
cd "$(dirname "$0")/.."
for f in r2_tests.log r2_kbench.txt r2_kbench_lc_old.txt r2_vs_reference_kernels.md r2_bench_n1.json r2_bench_ref_n1.json r2_bench_yolo_n1.json r2_launches.csv; do
  [ -s gpurun_out/$f ] && cp gpurun_out/$f profiles/${f/r2_/r02_}
done
for f in gpurun_out/r2_scale_*.json; do [ -s "$f" ] && cp "$f" profiles/$(basename ${f/r2_/r02_}); done
summ() { [ -s gpurun_out/$1.ncu-rep ] && python tools/ncu_summary.py gpurun_out/$1.ncu-rep "$2" > profiles/${1/r2_prof_/r02_}.md 2>/dev/null && echo "profiles/${1/r2_prof_/r02_}.md"; }
summ r2_prof_select "Radix select, single tensor (50.3 M elements, q = 0.9999): pass 0 (digit histogram), pass 1 (filter + compaction), pass 2 (early exit)"
summ r2_prof_select_spec "Radix select, table form with speculation: pass 0 (digit histogram + compaction of the candidates beyond the previous call's thresholds)"
summ r2_prof_lc_table "Per-channel fake-quant, short rows: shared-memory operator table kernel"
summ r2_prof_multi_hist "multi_histogram_t_kernel in bench.py (ResNet-50 activation set, batch 32)"
summ r2_prof_kl "KL search, one warp per candidate"
python tools/sass_summary.py > profiles/r02_sass_opcodes.md 2>/dev/null
ls profiles | grep r02_
