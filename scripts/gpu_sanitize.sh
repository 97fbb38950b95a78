cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_driver.py > gpurun_out/sanitize_$tool.log 2>&1; echo "$tool rc=$?"
  tail -4 gpurun_out/sanitize_$tool.log
done
