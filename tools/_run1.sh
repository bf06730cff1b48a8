python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/dfaust_layer_run.py seg_head 6 1; python tools/dfaust_layer_run.py enc1_block0 6 1; python tools/dfaust_layer_run.py dec2 6 1
python tools/dfaust_layer_run.py seg_head 4 1 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_seghead_v5.csv python tools/dfaust_layer_run.py seg_head 4 1 > gpurun_out/ncu_l.log 2>&1
python tools/dfaust_layer_run.py seg_head 4 1 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_agg_tc|k_edge_tc<" -s 9 -c 3 -o gpurun_out/prof_seghead_v5 python tools/dfaust_layer_run.py seg_head 4 1 > gpurun_out/ncu_f.log 2>&1; tail -2 gpurun_out/ncu_f.log
