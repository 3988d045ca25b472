# Round-2 evidence: run on the GPU box (gpurun -- 'bash tools/profile_round2.sh').  The .ncu-rep files stay in /tmp on the
# box (gpurun brings back at most 64 MiB); only the text summaries travel.
mkdir -p gpurun_out /tmp/rep
NCU="ncu --set full --clock-control none --import-source on"
SUM="python tools/ncu_summary.py"
# 1. bench launch list (headline only, 2 steps)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --secondary none > gpurun_out/p1.log 2>&1
python tools/launch_summary.py gpurun_out/r2_bench_launches.csv > gpurun_out/r2_bench_launches_summary.txt 2>&1
# 2. scorer inside the bench (grouped launch of 12 images)
timeout 400 $NCU -k regex:bvsb_stats_tma --profile-from-start off -s 1 -c 2 -o /tmp/rep/scorer python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --secondary none > gpurun_out/p2.log 2>&1
$SUM /tmp/rep/scorer.ncu-rep gpurun_out/r2_scorer_tma_c19_12img_full.txt > /dev/null 2>&1
# 3. loss kernels at rho 0.02 and 1.0
for rho in 0.02 1.0; do
timeout 400 $NCU -k regex:"multihot_loss_|multihot_dense_|tile_scan|grad_zero" -s 24 -c 6 -o /tmp/rep/losses_$rho python tools/bench_stage.py --profile --only losses --rho $rho --exact 0 > gpurun_out/p3.log 2>&1
$SUM /tmp/rep/losses_$rho.ncu-rep gpurun_out/r2_losses_rho${rho}_full.txt > /dev/null 2>&1
done
# 4. labeller
timeout 400 $NCU -k regex:proto_ -s 8 -c 4 -o /tmp/rep/labeller python tools/bench_stage.py --profile --only labeller > gpurun_out/p4.log 2>&1
$SUM /tmp/rep/labeller.ncu-rep gpurun_out/r2_labeller_full.txt > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 40 -c 24 --csv --log-file gpurun_out/r2_labeller_launches.csv python tools/bench_stage.py --profile --only labeller > gpurun_out/p5.log 2>&1
# 5. lowres scorer (x4 fast path)
timeout 300 $NCU -k regex:abreast -s 30 -c 1 -o /tmp/rep/lowres python tools/kbench_paths.py lowres > gpurun_out/p6.log 2>&1
$SUM /tmp/rep/lowres.ncu-rep gpurun_out/r2_lowres_x4_full.txt > /dev/null 2>&1
# 6. dense loss kernels (TMA strip walk) at rho = 1, N=16 x 20 x 768 x 768, int64 ids
timeout 400 $NCU -k regex:multihot_dense -s 2 -c 2 -o /tmp/rep/losses_dense python tools/probes/dense_probe.py --one > gpurun_out/p7.log 2>&1
$SUM /tmp/rep/losses_dense.ncu-rep gpurun_out/r2_losses_dense_rho1.0_full.txt > /dev/null 2>&1
# 7. scorer with bf16 logits (packed top-2 scan)
timeout 400 $NCU -k regex:bvsb_stats_tma --profile-from-start off -s 1 -c 1 -o /tmp/rep/scorer_bf16 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --secondary none --workload cityscapes_bf16 > gpurun_out/p8.log 2>&1
$SUM /tmp/rep/scorer_bf16.ncu-rep gpurun_out/r2_scorer_tma_c19_bf16_full.txt > /dev/null 2>&1
ls -la gpurun_out/ /tmp/rep
