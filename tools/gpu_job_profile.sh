M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum
python tools/profile_case.py bulk 151552 1 > gpurun_out/plain_bulk.log 2>&1 && \
ncu --set full --metrics $M --clock-control none --import-source on -k regex:k_run_bulk -c 1 -o gpurun_out/prof_bulk -f python tools/profile_case.py bulk 151552 1 > gpurun_out/ncu_bulk.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/ncu_bulk.log | cut -c1-200
