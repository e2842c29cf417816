O=gpurun_out
python bench.py > $O/r02_bench_n1.json 2>$O/b.err; tail -2 $O/b.err
ncu --clock-control none --metrics gpu__time_duration.sum -k regex:"^k_" -c 1300 --csv --log-file $O/r02_launches_bench.csv python bench.py --no-extra > $O/r02_bench_under_ncu.json 2> /dev/null
python tools/launch_summary.py $O/r02_launches_bench.csv > $O/r02_bench_launch_summary.txt; cat $O/r02_bench_launch_summary.txt
ncu --clock-control none --set full --import-source on -k regex:"k_(warp|pyramid|fast|distribute|describe)" --launch-skip 37 -c 37 -o /tmp/r02_frontend -f python tools/run_one_batch.py 128 2 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/r02_frontend.ncu-rep --lines k_fast,k_describe,k_pyramid > $O/r02_frontend_ncu_full.txt
python -c "
import json; d=json.loads(open('$O/r02_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['clocks']); e=d['extra']; print(e['local_ba']['lm_iters_per_s'], e['local_ba_dense']['lm_iters_per_s'], e['tracking']['frames_per_s_per_gpu']); print({k:(v['share_of_step'], v['ms_per_batch']) for k,v in d['roofline']['per_kernel'].items()}); print(d['roofline']['frac'], d['roofline']['pipeline_frac'], d['roofline']['per_kernel']['k_fast']['frac_of_minmax3_peak'])"
