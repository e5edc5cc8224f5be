#!/bin/bash
# After a GPU call: turn what came back in gpurun_out/ into the tracked evidence under profiles/ (run here, on the CPU box).
set -u
cd "$(dirname "$0")/.."
# final build (calls D / E) wins over the first pass (call A), which is kept as r02_first_pass_* where the numbers changed
for f in r2_kbench_lc_old.txt; do [ -s gpurun_out/$f ] && cp gpurun_out/$f profiles/${f/r2_/r02_}; done
[ -s gpurun_out/r2_kbench.txt ] && cp gpurun_out/r2_kbench.txt profiles/r02_first_pass_kbench.txt
for f in r2d_kbench.txt r2d_kbench_kl_old.txt r2d_vs_reference_kernels.md r2d_bench_ref_n1.json r2d_launches.csv r2e_kbench_quantile.txt r2e_bench_e2e10.json r2b_bench_yolo_ref.json; do
  [ -s gpurun_out/$f ] && cp gpurun_out/$f profiles/$(echo $f | sed 's/^r2[a-e]_/r02_/')
done
# the bench lines of the final build (call F: inputs through the device ring)
[ -s gpurun_out/r2f_bench_n1.json ] && cp gpurun_out/r2f_bench_n1.json profiles/r02_bench_n1_callF.json
for f in r2f_bench_yolo_n1.json r2f_bench_e2e_1.json r2f_bench_e2e_2.json r2g_bench_n1.json r2g_bench_n2.json; do
  [ -s gpurun_out/$f ] && cp gpurun_out/$f profiles/$(echo $f | sed 's/^r2[a-g]_/r02_/')
done
[ -s gpurun_out/r2e_bench_e2e10.json ] && cp gpurun_out/r2e_bench_e2e10.json profiles/r02_before_ring_bench_e2e10.json
rm -f profiles/r02_bench_e2e10.json
# call J: whole suite + driver command on the final build
[ -s gpurun_out/r2j_tests.log ] && cp gpurun_out/r2j_tests.log profiles/r02_tests.log
[ -s gpurun_out/r2j_bench_n1.json ] && cp gpurun_out/r2j_bench_n1.json profiles/r02_bench_n1_final.json
[ -s gpurun_out/r2d_vs_reference_kernels.md ] && cp gpurun_out/r2d_vs_reference_kernels.md profiles/r02_vs_reference_kernels_callD.md
[ -s gpurun_out/r2j_vs_reference_kernels.md ] && cp gpurun_out/r2j_vs_reference_kernels.md profiles/r02_vs_reference_kernels.md
[ -s gpurun_out/r2k_kbench_quantile.txt ] && cp gpurun_out/r2k_kbench_quantile.txt profiles/r02_kbench_quantile.txt
[ -s gpurun_out/r2e_kbench_quantile.txt ] && cp gpurun_out/r2e_kbench_quantile.txt profiles/r02_kbench_quantile_before_sampling.txt
[ -s gpurun_out/r2i_select_ab.txt ] && cp gpurun_out/r2i_select_ab.txt profiles/r02_select_ab.txt
[ -s gpurun_out/r2h_select_ab.txt ] && cp gpurun_out/r2h_select_ab.txt profiles/r02_select_launch_shapes.txt
python tools/launch_list_summary.py profiles/r02_launches.csv > profiles/r02_launches.md 2>/dev/null
for f in gpurun_out/r2_scale_*.json; do [ -s "$f" ] && cp "$f" profiles/$(basename ${f/r2_/r02_}); done
summ() { [ -s gpurun_out/$1.ncu-rep ] && python tools/ncu_summary.py gpurun_out/$1.ncu-rep "$2" > profiles/${1/r2_prof_/r02_}.md 2>/dev/null && echo "profiles/${1/r2_prof_/r02_}.md"; }
summ r2_prof_select "Radix select, single tensor (50.3 M elements, q = 0.9999), cold call with thresholds from a sample: sample, init (+ select among the samples), speculative pass 0, passes 1-2 (early exit), finish"
summ r2_prof_select_spec "Radix select, table form with speculation: pass 0 (digit histogram + compaction of the candidates beyond the previous call's thresholds)"
summ r2_prof_lc_table "Per-channel fake-quant, short rows: shared-memory operator table kernel"
summ r2_prof_multi_hist "multi_histogram_t_kernel in bench.py (ResNet-50 activation set, batch 32)"
summ r2_prof_kl "KL search, one warp per candidate"
python tools/sass_summary.py > profiles/r02_sass_opcodes.md 2>/dev/null
ls profiles | grep r02_
