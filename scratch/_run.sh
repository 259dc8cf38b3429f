bash scratch/qb.sh "--workload encoder6_bf16" "--workload decoder6_f32" "--workload encoder1_hr2000 --deterministic" "--workload encoder1" "--workload encoder6 --graph"
timeout 300 python bench.py --workload encoder_layer_ddp --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-300
timeout 300 python bench.py --workload encoder_stack6 --tf32 --fuse-prologue --padding --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-200
