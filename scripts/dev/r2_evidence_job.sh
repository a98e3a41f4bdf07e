# Round-2 evidence run (one B200): ncu capture of the contraction kernel, launch list of a bench step,
# compute-sanitizer logs, the reference's production binning at 1024^3.  Outputs land in gpurun_out/.
set -x
ncu --set full --clock-control none --import-source on -k regex:tc_contract_kernel -s 1 -c 1 -f -o gpurun_out/r2_tc_final ./build/tc_contract_check 0 27 40 > gpurun_out/r2_tc_final_ncu_log.txt 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2e_launches.csv python bench.py --profile --steps 1 --warmup 1 > gpurun_out/r2e_profile_stdout.txt 2>&1
compute-sanitizer --tool memcheck ./build/tc_contract_check 14 17 40 > gpurun_out/r2_sanitizer_memcheck_contract.txt 2>&1
compute-sanitizer --tool racecheck ./build/tc_contract_check 14 16 40 > gpurun_out/r2_sanitizer_racecheck_contract.txt 2>&1
compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_memcheck_smoke.txt 2>&1
timeout 900 python bench.py --nmesh 1024 --scheme paper80 --mesh device --no-cpu-baseline --steps 2 --warmup 1 --oracle-shells 45,50,55 > gpurun_out/r2_paper80_1024_n1.json 2> gpurun_out/r2_paper80_1024_n1.err
tail -c 200 gpurun_out/r2_sanitizer_memcheck_contract.txt; tail -c 200 gpurun_out/r2_sanitizer_racecheck_contract.txt; tail -c 300 gpurun_out/r2_sanitizer_memcheck_smoke.txt; head -c 300 gpurun_out/r2_paper80_1024_n1.json; tail -c 600 gpurun_out/r2_paper80_1024_n1.err
