timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/memcheck.log 2>&1; echo memcheck rc=$?; tail -8 gpurun_out/memcheck.log | cut -c1-300
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/racecheck.log 2>&1; echo racecheck rc=$?; tail -8 gpurun_out/racecheck.log | cut -c1-300
