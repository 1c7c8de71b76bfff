#!/bin/bash
# ncu evidence of one build (run on the GPU box: gpurun --timeout 1500 -- 'bash tools/gpu_profile.sh'):
#   * launch lists (gpu__time_duration.sum, --clock-control none) of one C3 forward, one C1 synthesis, front-end, train
#   * ncu --set full of the decoder conv layer, the IMV chain kernels, the length regulator, the resident stack kernels
# Reports land in gpurun_out/ncu/; tools/make_roofline_json.py turns the raw CSV pages into profiles/roofline_traffic.json.
set -u
O=gpurun_out/ncu
mkdir -p $O
NCU="ncu --clock-control none"
LL="--metrics gpu__time_duration.sum --csv"
$NCU $LL --log-file $O/launches_c3_forward.csv python tools/profile_workloads.py c3 2 > /dev/null 2>&1
$NCU $LL --log-file $O/launches_c1_inference.csv python tools/profile_workloads.py c1 3 > /dev/null 2>&1
$NCU $LL --log-file $O/launches_frontend.csv python tools/profile_workloads.py frontend 2 > /dev/null 2>&1
$NCU $LL --log-file $O/launches_train_slice.csv python tools/profile_workloads.py train 1 > /dev/null 2>&1
FULL="--set full --import-source on --kernel-name-base demangled"
# decoder conv: second forward, decoder layer 2
# gemm2 launches of a forward: text x5, key, value, prenet, mel x3, energy, expand, decoder x6, mel head, duration x2 = 22
$NCU --set full --import-source on -k regex:gemm2_kernel -s 36 -c 1 -o $O/dec_conv -f python tools/profile_workloads.py c3 3 > /dev/null 2>&1
$NCU $FULL -k regex:"reconstruct_alignment_rows_kernel|imv_scan_block_kernel|aligned_positions_block_kernel" -s 3 -c 3 -o $O/imv_chain -f python tools/profile_workloads.py c3 3 > /dev/null 2>&1
$NCU $FULL -k regex:"length_regulator_fwd_kernel" -s 1 -c 1 -o $O/length_regulator -f python tools/profile_workloads.py lr 3 > /dev/null 2>&1
$NCU $FULL -k regex:"stack_kernel" -s 2 -c 2 -o $O/stack -f python tools/profile_workloads.py c1 3 > /dev/null 2>&1
# weight-gradient GEMM (both operands MN-major): gemm2 launches of one forward + backward of the 2-layer stack are
# fwd x2, then per layer 5 weight-gradient launches + the data-gradient launch; the STFT GEMM with the magnitude epilogue
# is the first gemm2 launch of a front-end call
$NCU --set full --import-source on -k regex:gemm2_kernel -s 4 -c 1 -o $O/wgrad -f python tools/profile_workloads.py train 1 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:gemm2_kernel -s 2 -c 1 -o $O/stft_mag -f python tools/profile_workloads.py frontend 2 > /dev/null 2>&1
for r in dec_conv imv_chain length_regulator stack wgrad stft_mag; do
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2> /dev/null
done
ls -la $O
