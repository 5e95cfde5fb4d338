set -x
cd $GRAFT_REPO_ROOT
python tools/profile_forward.py 32 > gpurun_out/v12_pf.log 2>&1 || exit 1
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
$NCU -k regex:conv_igemm2 -c 14 -f -o gpurun_out/v12_conv python tools/profile_forward.py 32 > gpurun_out/v12_ncu_conv.log 2>&1
$NCU -k regex:"gn_apply|attention_tc|conv_in_kernel|conv_out_kernel" -c 12 -f -o gpurun_out/v12_elem python tools/profile_forward.py 32 > gpurun_out/v12_ncu_elem.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"median3d|residual_erode|edt_axis|filter_small|edge_seed|row_stats|threshold_counts" -c 10 -f -o gpurun_out/v12_tail python tools/time_volume.py 50 2 > gpurun_out/v12_ncu_tail.log 2>&1
ls -la gpurun_out/*.ncu-rep

# --- v15 / v17 (end of round 1): launch list of the bench command itself and the conv kernel after the epilogue change.
# Each ncu command ran only after the same program had exited 0 without ncu in the same gpurun call.
#   python bench.py --steps 1 --warmup 1 --start-t 3 --no-cpu-baseline
#   ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file gpurun_out/v15_launches_bench.csv \
#       python bench.py --steps 1 --warmup 1 --start-t 3 --no-cpu-baseline        -> profiles/r01_v15_launches_bench.csv
#   python tools/profile_forward.py 32
#   ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_igemm2 -c 14 -f \
#       -o gpurun_out/v17_conv python tools/profile_forward.py 32
#   python tools/ncu_extract.py gpurun_out/v17_conv.ncu-rep > profiles/r01_v17_ncu_conv.csv
# A/B switches used for the timing logs: CDDPM_FORK_EMBED=0|1, CDDPM_EMBED_SIMT=0|1, CDDPM_FUSE_GN=0|1 (+ _MINK),
# CDDPM_CONV_BENCH_STATS=0|1 (tools/bench_conv.py), TIME_FORWARD_ITERS=100 (tools/time_forward.py).
