python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_j.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_j.log
python tools/sanitize_case.py > gpurun_out/r02_sanitize_plain.log 2>&1 && \
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_case.py > gpurun_out/r02_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/r02_memcheck.log
