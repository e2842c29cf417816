# ncu captures behind profiles/r02_*: run on the GPU box (gpurun -- 'bash tools/profile_r02.sh'). The .ncu-rep files are summarised on the
# box (gpurun_out/ is capped at 64 MiB) and only the BA report is kept.
NCU="ncu --clock-control none"
O=gpurun_out
$NCU --metrics gpu__time_duration.sum -k regex:"^k_" -c 1300 --csv --log-file $O/r02_launches_bench.csv python bench.py > $O/r02_bench_under_ncu.json 2> /dev/null
python tools/launch_summary.py $O/r02_launches_bench.csv > $O/r02_bench_launch_summary.txt
$NCU --set full --import-source on -k regex:"k_(warp|pyramid|fast|distribute|describe)" --launch-skip 37 -c 37 -o /tmp/r02_frontend -f python tools/run_one_batch.py 128 2 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/r02_frontend.ncu-rep --lines k_fast,k_describe,k_pyramid > $O/r02_frontend_ncu_full.txt
$NCU --set full --import-source on -k regex:"k_match_bruteforce|k_search_by_bow|k_search_by_projection|k_frame_index|k_pose_opt|k_gather_pose" -c 24 -o /tmp/r02_legs -f python tools/run_legs.py > /dev/null 2>&1
python tools/ncu_summary.py /tmp/r02_legs.ncu-rep --lines k_match_bruteforce,k_search_by_bow,k_search_by_projection,k_pose_opt > $O/r02_legs_ncu_full.txt
CSLAM_BA_GRAPH=0 $NCU --set full --import-source on -k regex:"k_ba_" --launch-skip 40 -c 18 -o $O/r02_ba -f python tools/run_ba.py 0 > /dev/null 2>&1
python tools/ncu_summary.py $O/r02_ba.ncu-rep --lines k_ba_solve,k_ba_schur,k_ba_lin_points,k_ba_lin_poses > $O/r02_ba_ncu_full.txt
CSLAM_BA_GRAPH=0 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/r02_launches_ba.csv python tools/run_ba.py 0 > /dev/null 2>&1
python tools/launch_summary.py $O/r02_launches_ba.csv > $O/r02_ba_launch_summary.txt
ls -la $O /tmp/*.ncu-rep
