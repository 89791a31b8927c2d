#!/bin/bash
# Round evidence on a GPU box: tests, bench (both arms), ncu launch list, DRAM traffic of the evaluation
# launches at the bench workload, full ncu captures of the two dominant kernels. Outputs under gpurun_out/.
R=${1:-r02}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${R}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${R}_bench.log 2>&1; tail -1 gpurun_out/${R}_bench.log > gpurun_out/${R}_bench.json; cut -c1-300 gpurun_out/${R}_bench.json
timeout 900 python bench.py --impl reference > gpurun_out/${R}_bench_ref.log 2>&1; tail -1 gpurun_out/${R}_bench_ref.log > gpurun_out/${R}_bench_reference.json; cut -c1-300 gpurun_out/${R}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --grid 128 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_ncu_launches.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_eval_v3 -c 16 --csv --log-file gpurun_out/${R}_traffic.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${R}_ncu_traffic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eval_v3 -c 2 -o gpurun_out/${R}_k_eval_v3 -f python bench.py --grid 96 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${R}_ncu_eval.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_leaf_train -s 30 -c 1 -o gpurun_out/${R}_k_leaf_train -f python bench.py --grid 32 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${R}_ncu_train.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_obs_test_grouped|k_ball|k_frame_numerics|k_reeval_numerics|k_dirty_mark|k_obs_train" -s 16 -c 14 -o gpurun_out/${R}_other -f python bench.py --grid 32 --frames 6 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${R}_ncu_other.log 2>&1
timeout 600 python scripts/hbm_regime.py > gpurun_out/${R}_hbm_regime.json 2> gpurun_out/${R}_hbm_regime.err; cut -c1-400 gpurun_out/${R}_hbm_regime.json
timeout 600 python scripts/update_profile.py 40 > gpurun_out/${R}_update_profile.txt 2>&1
timeout 600 python scripts/train_bench.py 592 > gpurun_out/${R}_train_bench.txt 2>&1
ls -la gpurun_out | tail -16
