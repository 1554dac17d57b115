cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6400 -c 2250 --csv --log-file gpurun_out/launches_step_b8_final.csv python bench.py --batch 8 --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_final.log 2>&1
python tools/ncu_summarize.py gpurun_out/launches_step_b8_final.csv | tee gpurun_out/launches_step_b8_final_summary.txt | head -40
