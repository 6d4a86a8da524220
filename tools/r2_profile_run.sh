set -x
python tools/config_table.py > gpurun_out/r2_config_table.jsonl 2> gpurun_out/r2_config_table.err
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
python tools/prof_kpm_cg.py cfg4 4 > gpurun_out/r2_prof_kpm_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_kpm_cg_cfg4.csv python tools/prof_kpm_cg.py cfg4 4 > gpurun_out/r2_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_kpm_cheb_reg -s 2 -c 1 -f -o gpurun_out/prof_r2_cheb_reg python tools/prof_kpm_cg.py cfg4 4 > gpurun_out/r2_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tau_fft -s 4 -c 1 -f -o gpurun_out/prof_r2_fft python tools/prof_kpm_cg.py cfg4 4 > gpurun_out/r2_ncu_c.log 2>&1
python bench.py --steps 1 --warmup 3 --precond off --no-cpu > gpurun_out/r2_bench_plain_chk.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_cfg4.csv python bench.py --steps 1 --warmup 3 --precond off --no-cpu > gpurun_out/r2_ncu_d.log 2>&1
ls -la gpurun_out | tail -15
