PFB200_TRACE=1 python tools/trace_run.py 2 2> gpurun_out/trace_c2.log; echo rc=$?; tail -26 gpurun_out/trace_c2.log
