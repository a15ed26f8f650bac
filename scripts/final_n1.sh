# final single-GPU evidence: bench line, ncu launch list, ncu --set full of one step (each after the plain run exited 0)
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err || exit 1
K='regex:axis_pass|axis_tma|zfused|relax_|accumulate|compact|transpose'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-compact > gpurun_out/ncu_l_final.log 2>&1
ncu --set full --clock-control none -k "$K" -s 48 -c 16 -o /tmp/full_final -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-compact > gpurun_out/ncu_f_final.log 2>&1
python scripts/ncu_summary.py /tmp/full_final.ncu-rep > gpurun_out/full_final_summary.txt 2>&1
ls -la /tmp/full_final.ncu-rep; du -sh gpurun_out
tail -2 gpurun_out/bench_n1_final.err
