#!/bin/bash
# Refresh of the HBM-row evidence (light sections) for the final kernels + racecheck of the mbarrier / shared-memory kernels added in round 2.
set -x
REST='gs_|gn_seg|gn_tile|maxpool|upsample|patch_scores|knn_select|p2n_assign|kpconv_c1|row_positive|sinkhorn|topk_flat|nc_overlap|nc_compact|spm_|lgr_|evaluate|cloud_max|seg_exclusive|add_layernorm|l2_norm|gather_patches|splitk|linear_tc_kernel|linear_kernel|head_bias|rs_|att_scores|att_softmax|att_qk|att_pv'
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none --kernel-name regex:"$REST" --launch-skip 380 --launch-count 380 -f -o /tmp/r02_rest python tools/ncu_target.py 4 2 > gpurun_out/r02_ncu_d.log 2>&1
ncu -i /tmp/r02_rest.ncu-rep --page raw --csv > gpurun_out/r02_rest_raw.csv 2>/dev/null
ls -la gpurun_out/r02_rest_raw.csv
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/att_bench.py 2 100 1 > gpurun_out/r02_racecheck_att.txt 2>&1; echo "racecheck att exit $?" >> gpurun_out/r02_racecheck_att.txt; tail -4 gpurun_out/r02_racecheck_att.txt
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_batch.py tests/test_gpu_collate.py -q -x -k "per_pair or pyramid or demo2k" > gpurun_out/r02_racecheck_gn_rs.txt 2>&1; echo "racecheck gn/rs exit $?" >> gpurun_out/r02_racecheck_gn_rs.txt; tail -4 gpurun_out/r02_racecheck_gn_rs.txt
