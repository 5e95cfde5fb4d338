set -x
cd $GRAFT_REPO_ROOT
python tools/profile_forward.py 32 > gpurun_out/v12_pf.log 2>&1 || exit 1
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
$NCU -k regex:conv_igemm2 -c 14 -f -o gpurun_out/v12_conv python tools/profile_forward.py 32 > gpurun_out/v12_ncu_conv.log 2>&1
$NCU -k regex:"gn_apply|attention_tc|conv_in_kernel|conv_out_kernel" -c 12 -f -o gpurun_out/v12_elem python tools/profile_forward.py 32 > gpurun_out/v12_ncu_elem.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"median3d|residual_erode|edt_axis|filter_small|edge_seed|row_stats|threshold_counts" -c 10 -f -o gpurun_out/v12_tail python tools/time_volume.py 50 2 > gpurun_out/v12_ncu_tail.log 2>&1
ls -la gpurun_out/*.ncu-rep
