set -x
# end-of-round-2 evidence: GPU tests, launch list of the bench command, DRAM traffic of the step's 32 scans,
# full captures of the scans that changed (1 and 2 bit ring-less + programmatic launch; the 17-bit scan that
# also fills the position list's segment counts)
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_pytest_gpu_v9.log; cat gpurun_out/r02_pytest_gpu_v9.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_v2.csv python bench.py --logn 28 --steps 2 --warmup 1 --no-extras --no-cpu > gpurun_out/r02_ncu_launches_v2.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:scan_(range|direct)_kernel' --csv --log-file gpurun_out/scan_traffic.csv python tools/scan_traffic.py > /dev/null 2>&1
python tools/summarise_traffic.py gpurun_out/scan_traffic.csv gpurun_out/r02_scan_traffic_v2.json
for w in 1 2; do
  ncu --set full --clock-control none --import-source on -k 'regex:scan_(range|direct)_kernel' -s 2 -c 1 -o gpurun_out/r02_ncu_scan_w${w}_v2 -f python tools/ncu_family_driver.py scan $w 30 > /dev/null 2>&1
  python tools/ncu_brief.py gpurun_out/r02_ncu_scan_w${w}_v2.ncu-rep $((1<<30)) > gpurun_out/r02_ncu_scan_w${w}_v2.txt
  rm -f gpurun_out/r02_ncu_scan_w${w}_v2.ncu-rep
done
SEL=0.01 ncu --set full --clock-control none --import-source on -k 'regex:scan_range_kernel' -s 3 -c 1 -o gpurun_out/r02_ncu_scanseg_w17 -f python tools/ncu_family_driver.py scanpos 17 30 > /dev/null 2>&1
python tools/ncu_brief.py gpurun_out/r02_ncu_scanseg_w17.ncu-rep $((1<<30)) > gpurun_out/r02_ncu_scanseg_w17.txt; rm -f gpurun_out/r02_ncu_scanseg_w17.ncu-rep
head -12 gpurun_out/r02_ncu_scanseg_w17.txt; head -3 gpurun_out/r02_ncu_scan_w1_v2.txt
