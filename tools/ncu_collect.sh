#!/bin/bash
# Evidence collection on the GPU box (run through gpurun): launch list, --set full of the GPU-filling kernels, light sections for
# the HBM-bound rows.  Outputs under gpurun_out/ (condensed afterwards with tools/ncu_select.py into profiles/).
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_final.csv python tools/ncu_target.py 8 3 > gpurun_out/r02_ncu_a.log 2>&1
TOP='linear_tc_persistent_kernel|gse_embed_f16_kernel|table_embed_kernel|att_stream_kernel|kpconv_gather_kernel|rs_query_kernel|att_qk_kernel|att_pv_kernel'
timeout 700 ncu --set full --clock-control none --kernel-name regex:"$TOP" --launch-skip 70 --launch-count 60 -f -o /tmp/r02_top_full python tools/ncu_target.py 4 2 > gpurun_out/r02_ncu_b.log 2>&1
ncu -i /tmp/r02_top_full.ncu-rep --page raw --csv > gpurun_out/r02_top_full_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:"linear_tc_persistent_kernel|gse_embed_f16_kernel|att_stream_kernel" --launch-skip 50 --launch-count 3 -f -o gpurun_out/r02_tc_source python tools/ncu_target.py 4 2 > gpurun_out/r02_ncu_c.log 2>&1
REST='gs_|gn_seg|gn_tile|maxpool|upsample|patch_scores|knn_select|p2n_assign|kpconv_c1|row_positive|sinkhorn|topk_flat|nc_overlap|nc_compact|spm_|lgr_|evaluate|cloud_max|seg_exclusive|add_layernorm|l2_norm|gather_patches|splitk|linear_tc_kernel|linear_kernel|head_bias|rs_bounds|rs_count|rs_scatter|att_scores|att_softmax'
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none --kernel-name regex:"$REST" --launch-skip 330 --launch-count 330 -f -o /tmp/r02_rest python tools/ncu_target.py 4 2 > gpurun_out/r02_ncu_d.log 2>&1
ncu -i /tmp/r02_rest.ncu-rep --page raw --csv > gpurun_out/r02_rest_raw.csv 2>/dev/null
ls -la gpurun_out/*.csv gpurun_out/*.ncu-rep
tail -2 gpurun_out/r02_ncu_b.log gpurun_out/r02_ncu_d.log
